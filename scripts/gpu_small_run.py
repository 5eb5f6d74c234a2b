"""Small optimise run for compute-sanitizer: two dynamic streams + a fixed one, merge on."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200 import optimise_batch
def dfl(d, strat=0):
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, strat); return co.compress(d) + co.flush()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
streams = [dfl(W.c2_text(n)), dfl(W.c2_text(n // 4, seed=3)), dfl(W.c2_text(3000, seed=4), zlib.Z_FIXED)]
res = optimise_batch(streams, True)
print([(r["status"], r["saved_bits"]) for r in res])
