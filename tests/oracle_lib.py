"""ctypes binding of the CPU oracle (oracle/libdeft_oracle.so) — TEST INFRASTRUCTURE.

`OracleDeflateStream` has the same method surface as `deft4j_b200.DeflateStream`, so the container
mirrors can be driven by either.  Nothing under deft4j_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import zlib

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libdeft_oracle.so")


class BlockInfo(C.Structure):
    _fields_ = [("type", C.c_int32), ("size_bits", C.c_int64), ("position", C.c_int64),
                ("uncompressed_len", C.c_uint64), ("n_symbols", C.c_uint32), ("n_rle_pairs", C.c_uint32),
                ("num_litlen_lens", C.c_int32), ("num_dist_lens", C.c_int32), ("num_codelen_lens", C.c_int32),
                ("litlen_size_bits", C.c_int64), ("header_size_bits", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("optimise_block_calls", "candidates", "header_rewrites",
                                           "tree_builds_small", "tree_builds_big", "symbol_passes")]


def build():
    src = os.path.join(_ROOT, "oracle", "deft_oracle.cpp")
    if (not os.path.exists(_SO)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ora_parse.restype = C.c_void_p
        L.ora_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_optimise.restype = C.c_int64
        L.ora_optimise.argtypes = [C.c_void_p, C.c_int]
        L.ora_size_bits.restype = C.c_int64
        L.ora_size_bits.argtypes = [C.c_void_p]
        L.ora_uncompressed_len.restype = C.c_size_t
        L.ora_uncompressed_len.argtypes = [C.c_void_p]
        L.ora_uncompressed.argtypes = [C.c_void_p, C.c_char_p]
        L.ora_write.restype = C.c_size_t
        L.ora_write.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ora_block_count.restype = C.c_uint32
        L.ora_block_count.argtypes = [C.c_void_p]
        L.ora_block_info_get.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(BlockInfo)]
        L.ora_block_symbols.restype = C.c_uint32
        L.ora_block_symbols.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32]
        L.ora_block_rle_pairs.restype = C.c_uint32
        L.ora_block_rle_pairs.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32]
        L.ora_block_codelens.restype = C.c_uint32
        L.ora_block_codelens.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.c_uint32]
        L.ora_huffman_tree.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32)]
        L.ora_pack_code_lengths.restype = C.c_int
        L.ora_pack_code_lengths.argtypes = [C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int,
                                            C.POINTER(C.c_int32), C.c_int]
        L.ora_get_stats.argtypes = [C.POINTER(Stats)]
        _lib = L
    return _lib


class OracleDeflateStream:
    """DeflateStream surface (base/deflate/DeflateStream.java) backed by the CPU oracle."""

    def __init__(self, name=None):
        self.name = name if name is not None else "unnamed stream"
        self.h = None

    def __del__(self):
        if self.h is not None and _lib is not None:
            _lib.ora_free(self.h)
            self.h = None

    def getName(self): return self.name
    def setName(self, n): self.name = n

    def parse(self, src):
        """parse(InputStream | byte[]): a reader is left positioned after the last consumed byte."""
        from deft4j_b200.container._io import ByteReader
        if isinstance(src, ByteReader):
            data = src.remaining()
        else:
            data = bytes(src)
        consumed = C.c_size_t(0)
        self.h = lib().ora_parse(data, len(data), C.byref(consumed))
        if isinstance(src, ByteReader):
            src.pos += consumed.value
        self.consumed = consumed.value
        return self.h is not None

    def optimise(self, mergeBlocks=True):
        return lib().ora_optimise(self.h, 1 if mergeBlocks else 0)

    def getSizeBits(self):
        return lib().ora_size_bits(self.h)

    def getUncompressedData(self):
        n = lib().ora_uncompressed_len(self.h)
        buf = C.create_string_buffer(n if n else 1)
        lib().ora_uncompressed(self.h, buf)
        return buf.raw[:n]

    def getChecksums(self):
        d = self.getUncompressedData()
        return zlib.crc32(d) & 0xffffffff, zlib.adler32(d) & 0xffffffff, len(d)

    def asBytes(self):
        n = lib().ora_write(self.h, None, 0)
        buf = C.create_string_buffer(n if n else 1)
        lib().ora_write(self.h, buf, n)
        return buf.raw[:n]

    def blockCount(self):
        return lib().ora_block_count(self.h)

    def blockInfo(self, i):
        bi = BlockInfo()
        if lib().ora_block_info_get(self.h, i, C.byref(bi)) != 0:
            raise IndexError(i)
        return bi

    def blockSymbols(self, i):
        n = self.blockInfo(i).n_symbols
        arr = (C.c_int32 * (3 * max(n, 1)))()
        lib().ora_block_symbols(self.h, i, arr, n)
        return [(arr[3 * k], arr[3 * k + 1], arr[3 * k + 2]) for k in range(n)]

    def blockRlePairs(self, i):
        n = self.blockInfo(i).n_rle_pairs
        arr = (C.c_int32 * (2 * max(n, 1)))()
        lib().ora_block_rle_pairs(self.h, i, arr, n)
        return [(arr[2 * k], arr[2 * k + 1]) for k in range(n)]

    def blockCodelens(self, i, which):
        arr = (C.c_int32 * 320)()
        n = lib().ora_block_codelens(self.h, i, which, arr, 320)
        return list(arr[:n])

    def printBlockInfo(self):
        names = ["STORED", "FIXED", "DYNAMIC"]
        s = ""
        n = self.blockCount()
        for i in range(n):
            bi = self.blockInfo(i)
            s += "\nBlock %d position %d size %d type %s" % (i, bi.position, bi.size_bits + 3, names[bi.type])
        return "Stream name: " + self.name + "\nBlock info:" + s + "\nTotal blocks: %d" % n


def huffman_tree(freq, limit):
    n = len(freq)
    f = (C.c_int32 * n)(*freq)
    code = (C.c_int32 * n)()
    ln = (C.c_int32 * n)()
    lib().ora_huffman_tree(f, n, limit, code, ln)
    return list(code), list(ln)


def pack_code_lengths(lit, dist, flags):
    a = (C.c_int32 * len(lit))(*lit)
    b = (C.c_int32 * max(len(dist), 1))(*dist)
    out = (C.c_int32 * 1024)()
    n = lib().ora_pack_code_lengths(a, len(lit), b, len(dist), flags, out, 1024)
    return list(out[:n])


def stats():
    s = Stats()
    lib().ora_get_stats(C.byref(s))
    return {n: getattr(s, n) for n, _ in Stats._fields_}


def header_trial(lit, dist, flags, post_op=0):
    """optimiseBlockDynBlock on a bare dynamic block (see ora_header_trial)."""
    L = lib()
    L.ora_header_trial.restype = C.c_int64
    L.ora_header_trial.argtypes = [C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    a = (C.c_int32 * len(lit))(*lit)
    b = (C.c_int32 * len(dist))(*dist)
    pairs = (C.c_int32 * 700)()
    np_, ncl = C.c_int32(0), C.c_int32(0)
    cl = (C.c_int32 * 19)()
    bits = L.ora_header_trial(a, len(lit), b, len(dist), flags, post_op, pairs, C.byref(np_), cl, C.byref(ncl))
    return bits, [(pairs[2 * i], pairs[2 * i + 1]) for i in range(np_.value)], list(cl), ncl.value
