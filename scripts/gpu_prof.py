"""Engine phase breakdown (needs the -DD4_PROF build: DEFT4CU_LIB=.../libdeft4cu_prof.so)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200 import optimise_batch, _native as N
NAMES = ["block", "repl_hit", "repl_miss", "least_hit", "least_miss", "hist", "recode_hit", "recode_miss", "trees",
         "hdr_default", "intern_tab", "intern_mask", "copy", "cb", "trials", "trials_eval", "hdr_opt", "hdr_recode",
         "to_fixed", "flush", "payload"]
mib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
raw = W.c2_stream(int(mib * (1 << 20)))
L = N.lib()
buf = (C.c_uint64 * 64)()
optimise_batch([raw[:200000] if False else raw], False)
L.deft4cu_debug_prof(buf, 64, 1)
t = time.time(); r = optimise_batch([raw], False)[0]; dt = time.time() - t
assert L.deft4cu_debug_prof(buf, 64, 0) == 0, N.last_error()
tot = buf[0]
print("wall %.3fs saved %d; block-cycles total %.3e (calls %d)" % (dt, r["saved_bits"], tot, buf[32]))
for i, nm in enumerate(NAMES):
    if buf[32 + i]:
        print("%-12s %6.2f%%  calls %9d  cyc/call %9.0f  calls/block-round %.1f" % (nm, 100.0 * buf[i] / tot, buf[32 + i], buf[i] / buf[32 + i], buf[32 + i] / buf[32]))
