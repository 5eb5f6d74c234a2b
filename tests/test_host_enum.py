"""CPU pin of the symbolic candidate enumerator (deft4j_b200/csrc/enum.cuh, compiled as host code): the enumeration
order, memo / request / sweep protocol, done-node skipping, segmented rounds and the index bookkeeping of the skipped
sweep, driven by a plain serial executor (hosttest.cu) and compared with the oracle's candidate trace of
DeflateStream.optimiseBlock (DeflateStream.java:343-490), call by call and candidate by candidate."""
import ctypes as C
import zlib

import numpy as np
import pytest

import hosttest_lib as H
import workloads as W
from conftest import read_golden

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]


def _lensym(length, edge):
    if edge:
        return 284
    if length == 258:
        return 285
    return 257 + max(k for k in range(28) if LEN_BASE[k] <= length)


def _pack_block(triples):
    """oracle {dist, litlen, edge} triples -> packed symbols (common.cuh) + block-relative decoded offsets"""
    sym, so, off = [], [], 0
    for dist, ll, edge in triples:
        so.append(off)
        if dist == 0:
            sym.append(ll)
            off += 1 if ll < 256 else 0
        else:
            sym.append(0x80000000 | (ll - 3) | ((dist - 1) << 9) | (edge << 24) | ((_lensym(ll, edge) - 257) << 25))
            off += ll
    return np.array(sym, dtype=np.uint32), np.array(so, dtype=np.uint32), off


def _engine():
    L = H.lib()
    L.host_engine_load.restype = C.c_void_p
    L.host_engine_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_int,
                                   C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_longlong, C.c_int]
    L.host_engine_free.argtypes = [C.c_void_p]
    L.host_engine_round.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                    C.c_uint, C.POINTER(C.c_uint)]
    L.host_engine_advance.argtypes = [C.c_void_p, C.c_int]
    return L


def _oracle_calls(oracle, raw):
    OL = oracle.lib()
    OL.ora_trace_begin.argtypes = [C.POINTER(C.c_int64), C.c_size_t]
    OL.ora_trace_end.restype = C.c_size_t
    CAP = 3_000_000
    o = oracle.OracleDeflateStream()
    assert o.parse(raw)
    blocks = []
    data = o.getUncompressedData()
    pos = 0
    for i in range(o.blockCount()):
        bi = o.blockInfo(i)
        blk = {"type": bi.type, "ulen": bi.uncompressed_len, "data": data[pos:pos + bi.uncompressed_len], "payload": bi.litlen_size_bits,
               "hbits": bi.header_size_bits, "ncl": bi.num_codelen_lens}
        pos += bi.uncompressed_len
        if bi.type != 0:
            blk["sym"] = o.blockSymbols(i)
            if bi.type == 2:
                blk["L"], blk["D"], blk["CL"] = o.blockCodelens(i, 0), o.blockCodelens(i, 1), o.blockCodelens(i, 2)
                blk["pairs"] = o.blockRlePairs(i)
        blocks.append(blk)
    buf = (C.c_int64 * (2 * CAP))()
    OL.ora_trace_begin(buf, CAP)
    o.optimise(False)
    n = OL.ora_trace_end()
    assert n < CAP
    a = np.ctypeslib.as_array(buf)[:2 * n].reshape(-1, 2)
    calls = []
    for idx, sz in a.tolist():
        if idx == -1:
            calls.append([sz, {}])
        else:
            calls[-1][1][idx] = sz
    return blocks, calls


def _load(L, blk, to_fixed=False):
    sym, so, off = _pack_block(blk["sym"])
    assert off == blk["ulen"]
    if blk["type"] == 2:
        pairs, prev = [], 0
        for run, s in blk["pairs"]:   # {run, sym}: the repeated value is the previous length for 16, 0 for 17 / 18
            val = s if s <= 15 else (prev if s == 16 else 0)
            pairs += [s, run, val]
            prev = val
        pa = (C.c_int32 * max(1, len(pairs)))(*pairs)
        Lb, Db, CLb = bytes(blk["L"]), bytes(blk["D"]), bytes(blk["CL"])
        return L.host_engine_load(sym.ctypes.data, so.ctypes.data, len(sym), blk["data"], blk["ulen"], 2, Lb, len(Lb), Db, len(Db),
                                  pa, len(pairs) // 3, CLb, blk["ncl"], blk["hbits"], blk["payload"], int(to_fixed)), (sym, so)
    return L.host_engine_load(sym.ctypes.data, so.ctypes.data, len(sym), blk["data"], blk["ulen"], 1, None, 0, None, 0, None, 0, None, 0, 0,
                              blk["payload"], int(to_fixed)), (sym, so)


def _check_stream(oracle, raw, segmented=False, keep_pools=True):
    L = _engine()
    blocks, calls = _oracle_calls(oracle, raw)
    ci = 0
    stats = {"rounds": 0, "sweeps": 0, "segmented": 0, "masks": 0, "slow": 0}
    for bi, blk in enumerate(blocks):
        if blk["ulen"] == 0 and len(blocks) > 1:
            break   # removed; the reference's loop ends here
        if blk["type"] == 0:
            ci += 1   # optimiseBlock on a stored block compares nothing
            continue
        h, keep = _load(L, blk)
        try:
            while True:
                incumbent, ref = calls[ci]
                ci += 1
                res = (C.c_longlong * 16)()
                CAP = 600_000
                tr = (C.c_longlong * (2 * CAP))()
                tn = C.c_uint(0)
                assert L.host_engine_round(h, -1, int(segmented), res, tr, CAP, C.byref(tn)) == 0, (bi, list(res))
                assert res[3] == incumbent, (bi, "incumbent", res[3], incumbent)
                got = np.ctypeslib.as_array(tr)[:2 * tn.value].reshape(-1, 2).tolist()
                assert got[0] == [-1, incumbent]
                hasO = res[4] != res[3]
                stored_idx = (1 if hasO else 0) if blk["ulen"] <= 65535 else None
                seen = set()
                for idx, sz in got[1:]:
                    assert ref.get(idx) == sz, (bi, ci, idx, ref.get(idx), sz)
                    seen.add(idx)
                # everything the oracle logged that we did not: the stored candidate and the skipped sweeps (which never
                # hold a strict minimum)
                best = min(ref.values()) if ref else incumbent
                first = min(i for i, v in ref.items() if v == best) if ref else None
                assert res[6] == (max(ref) + 1 if ref else 0), (bi, "candidate count", res[6], max(ref) + 1)
                if stored_idx is not None and ref:
                    # again with the stored size the oracle saw at that index: the complete selection
                    res2 = (C.c_longlong * 16)()
                    assert L.host_engine_round(h, ref[stored_idx], int(segmented), res2, None, 0, None) == 0
                    res = res2
                    assert stored_idx not in seen
                stats["rounds"] += 1; stats["sweeps"] += res[8]; stats["segmented"] += res[9]
                stats["masks"] = max(stats["masks"], res[10]); stats["slow"] = res[14]
                if best < incumbent:
                    assert (res[0], res[1]) == (best, first), (bi, ci, list(res)[:8], best, first)
                    if res[2]:
                        break   # the block became STORED
                    assert res[15] == best, (bi, "materialised winner", res[15], best)
                    L.host_engine_advance(h, int(keep_pools))
                else:
                    assert res[0] == incumbent and res[1] == 0xffffffff
                    break
        finally:
            L.host_engine_free(h)
    return stats


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(data) + co.flush()


def _fixture_streams(oracle, inp, take):
    from deft4j_b200.container import getContainerForBytes
    data = read_golden(inp)
    co = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
    assert co.read(data)
    return [s.asBytes() for s in co.getDeflateStreams()][:take]


@pytest.mark.parametrize("mode", ["all", "segmented", "reset_every_round"])
def test_enumerator_matches_oracle_trace(oracle, mode):
    raws = _fixture_streams(oracle, "text.png", 3) + _fixture_streams(oracle, "apng/ball.png", 2)
    raws += [_deflate(W.c2_text(6000, seed=3), 6, zlib.Z_FIXED), _deflate(W.c2_text(30_000, seed=4)),
             W.handmade_streams()["edge284"], W.handmade_streams()["dyn_partialflush_dyn"], W.handmade_streams()["rle_two_fixed"],
             _deflate(bytes(range(256)) * 8), _deflate(b"ab" * 3000 + bytes(range(64)))]
    raws += [s for s in W.c5_streams() if len(s) < 3000][:3]
    total = {"rounds": 0, "sweeps": 0, "segmented": 0}
    for raw in raws:
        st = _check_stream(oracle, raw, segmented=(mode == "segmented"), keep_pools=(mode != "reset_every_round"))
        for k in total:
            total[k] += st[k]
    assert total["rounds"] >= len(raws)
    if mode == "segmented":
        assert total["segmented"] == total["rounds"]


@pytest.mark.slow
def test_enumerator_on_the_asyoulik_fixture(oracle):
    for raw in _fixture_streams(oracle, "asyoulik/asyoulik-gzip.txt.gz", 1):
        _check_stream(oracle, raw)
