"""Per-kernel-family device time of one batch.
usage: gpu_timing.py <merge 0|1> golden:<file>[*rep] | c2:<MiB> | c3:<count> | c4:<count> ..."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W
from deft4j_b200 import _native as N


def golden_raws(nm):
    from oracle_lib import OracleDeflateStream
    from conftest import read_golden
    from deft4j_b200.container import getContainerForBytes
    raws = []
    class Capture(OracleDeflateStream):
        def parse(self, src):
            from deft4j_b200.container._io import ByteReader
            data = src.remaining() if isinstance(src, ByteReader) else bytes(src)
            ok = super().parse(src)
            raws.append(data[:self.consumed])
            return ok
    data = read_golden(nm)
    getContainerForBytes(data, nm, Capture).read(data)
    return raws


merge = int(sys.argv[1])
if os.environ.get('D4_TORCH'):
    import torch; torch.cuda.init(); torch.zeros(1, device='cuda')
bufs = []
for spec in sys.argv[2:]:
    kind, arg = spec.split(":")
    if kind == "golden":
        nm, _, rep = arg.partition("*")
        bufs += golden_raws(nm) * int(rep or 1)
    elif kind == "c2":
        bufs.append(W.c2_stream(int(float(arg) * (1 << 20))))
    elif kind == "c3":
        bufs += W.c3_streams(int(arg))
    elif kind == "c4":
        bufs += W.c4_streams(int(arg))
L = N.lib()
ptrs, lens = N.make_ptr_arrays(bufs)
h = C.c_void_p()
assert L.deft4cu_device_batch_create(ptrs, lens, len(bufs), C.byref(h)) == 0
for it in range(int(os.environ.get("D4_ITERS", "2"))):
    launches = C.c_uint64(0)
    t0 = time.time()
    rc = L.deft4cu_device_batch_run(h, merge, C.byref(launches), None)
    dt = time.time() - t0
    ms = (C.c_float * 8)()
    L.deft4cu_device_batch_timings(h, ms, 8)
    tot = sum(len(b) for b in bufs)
    print("rc %d streams %d bytes %d wall %.3fs launches %d  %.3f MB/s  ms: count %.1f emit %.1f lz %.1f opt %.1f finish %.1f write %.1f sums %.1f" % (
        rc, len(bufs), tot, dt, launches.value, tot / dt / 1e6, *list(ms)[:7]), flush=True)
    if rc:
        print(N.last_error())
