"""Locate the first candidate whose size differs between the CUDA enumerator and the oracle.

usage: python scripts/gpu_trace_diff.py <golden file> [stream index ...]
Single-block streams only (the traces of several blocks interleave on the device).
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from oracle_lib import OracleDeflateStream
from conftest import read_golden
from deft4j_b200 import DeflateStream, _native
from deft4j_b200.container import getContainerForBytes

CAP = 4_000_000
raws = []


class Capture(OracleDeflateStream):
    def parse(self, src):
        from deft4j_b200.container._io import ByteReader
        data = src.remaining() if isinstance(src, ByteReader) else bytes(src)
        ok = super().parse(src)
        raws.append(data[:self.consumed])
        return ok


def split_calls(pairs):
    calls = []
    for idx, sz in pairs:
        if idx == -1:
            calls.append({"incumbent": sz, "c": {}})
        else:
            calls[-1]["c"][idx] = sz
    return calls


def oracle_trace(raw):
    L = oracle_lib.lib()
    L.ora_trace_begin.argtypes = [C.POINTER(C.c_int64), C.c_size_t]
    L.ora_trace_end.restype = C.c_size_t
    buf = (C.c_int64 * (2 * CAP))()
    s = OracleDeflateStream()
    assert s.parse(raw)
    L.ora_trace_begin(buf, CAP)
    saved = s.optimise(False)
    n = L.ora_trace_end()
    return saved, split_calls([(buf[2 * i], buf[2 * i + 1]) for i in range(min(n, CAP))]), s


def gpu_trace(raw):
    L = _native.lib()
    s = DeflateStream()
    assert s.parse(raw)
    assert L.deft4cu_debug_trace_begin(CAP) == 0
    saved = s.optimise(False)
    buf = (C.c_int64 * (2 * CAP))()
    n = C.c_uint32(0)
    assert L.deft4cu_debug_trace_end(buf, CAP, C.byref(n)) == 0
    return saved, split_calls([(buf[2 * i], buf[2 * i + 1]) for i in range(min(n.value, CAP))]), s


name = sys.argv[1]
sel = [int(x) for x in sys.argv[2:]]
data = read_golden(name)
c = getContainerForBytes(data, name, Capture)
c.read(data)
for k, raw in enumerate(raws):
    if sel and k not in sel:
        continue
    so, co, os_ = oracle_trace(raw)
    sg, cg, gs_ = gpu_trace(raw)
    print("stream %d: saved oracle %d gpu %d; calls oracle %d gpu %d" % (k, so, sg, len(co), len(cg)))
    for r, (a, b) in enumerate(zip(co, cg)):
        if a["incumbent"] != b["incumbent"]:
            print("  call %d incumbent oracle %d gpu %d" % (r, a["incumbent"], b["incumbent"]))
        bad = [i for i in sorted(b["c"]) if i in a["c"] and a["c"][i] != b["c"][i]]
        missing = [i for i in sorted(b["c"]) if i not in a["c"]]
        print("  call %d: oracle %d candidates, gpu logged %d, %d differ, %d not in oracle; oracle max idx %d gpu max idx %d" % (
            r, len(a["c"]), len(b["c"]), len(bad), len(missing), max(a["c"]) if a["c"] else -1, max(b["c"]) if b["c"] else -1))
        for i in bad[:12]:
            print("     idx %d oracle %d gpu %d" % (i, a["c"][i], b["c"][i]))
        if bad or missing:
            break
