"""Engine phase breakdown (needs the -DD4_PROF build: DEFT4CU_LIB=.../libdeft4cu_prof.so).
usage: DEFT4CU_LIB=deft4j_b200/libdeft4cu_prof.so python scripts/gpu_prof.py [MiB of C2 stream]
Cycles are thread 0's clock64 deltas per CTA, summed over CTAs; categories nest (round contains sweeps, passes, ...)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200 import optimise_batch, _native as N
NAMES = ["block", "round", "sweep", "select", "pass", "dc_build", "recode", "trees(warp0)", "hdr_default(warp0)", "hdr_ops", "trials",
         "load", "materialise", "advance", "intern_mask", "intern_tab", "fixed", "slow_tree", "segmented", "hist_full",
         "replace_main", "replace_hq", "least_apply", "least_stats", "sign_masks"]
mib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
merge = len(sys.argv) > 2 and sys.argv[2] == "merge"
raw = W.c2_stream(int(mib * (1 << 20)))
L = N.lib()
buf = (C.c_uint64 * 64)()
optimise_batch([raw], merge)
L.deft4cu_debug_prof(buf, 64, 1)
t = time.time(); r = optimise_batch([raw], merge)[0]; dt = time.time() - t
assert L.deft4cu_debug_prof(buf, 64, 0) == 0, N.last_error()
tot = buf[0]
print("wall %.3fs saved %d; block-cycles total %.3e (blocks %d, rounds %d)" % (dt, r["saved_bits"], tot, buf[32], buf[33]))
for i, nm in enumerate(NAMES):
    if buf[32 + i]:
        print("%-20s %6.2f%%  calls %9d  cyc/call %9.0f  calls/block %.2f" % (nm, 100.0 * buf[i] / tot, buf[32 + i], buf[i] / buf[32 + i], buf[32 + i] / buf[32]))
