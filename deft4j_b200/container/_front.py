"""Calls into the native file front-ends (`deft4cu_png_optimise_batch`, `deft4cu_zip_optimise_batch`)."""
import ctypes as C

from .. import _native as N


def front_call(entry, datas, merge_blocks, lib=None):
    """The C-ABI call alone: (library, result array, n); the caller frees with deft4cu_free_file_results.
    `lib`: another library exporting the same entry points (the CPU tests pass a build of the front-ends over the oracle)."""
    L = lib if lib is not None else N.lib()
    n = len(datas)
    ptrs, lens = N.make_ptr_arrays(datas)
    res = (N.FileResult * max(n, 1))()
    fn = getattr(L, entry)
    if lib is not None:
        fn.restype = C.c_int
        fn.argtypes = N.SYMBOLS[entry][1]
        L.deft4cu_free_file_results.restype = None
        L.deft4cu_free_file_results.argtypes = N.SYMBOLS["deft4cu_free_file_results"][1]
    rc = fn(ptrs, lens, n, N.MERGE_BLOCKS if merge_blocks else 0, res)
    if rc != N.OK:
        raise N.Deft4cuError("%s failed (%d): %s" % (entry, rc, N.last_error() if lib is None else ""))
    return L, res, n


def front_optimise(entry, datas, merge_blocks, lib=None, name_encoding="latin-1"):
    """One dict per file: status (0 = read, optimised and written; 1 = the container's `read` is False; 2 = `write` raises;
    3 = a stream hit an internal limit), out (the bytes `write()` returns), saved_bits, streams = [(name, saved bits)] in
    `getDeflateStreams()` order — everything `CMDUtil.optimiseFile` prints or writes for the file."""
    L, res, n = front_call(entry, datas, merge_blocks, lib)
    out = []
    try:
        for i in range(n):
            r = res[i]
            d = {"status": r.status, "saved_bits": r.saved_bits, "out": None, "streams": []}
            if r.status == N.OK:
                d["out"] = C.string_at(r.out, r.out_len)
                d["streams"] = [(r.stream_name[k].decode(name_encoding, "replace"), r.stream_saved[k]) for k in range(r.n_streams)]
            out.append(d)
    finally:
        L.deft4cu_free_file_results(res, n)
    return out
