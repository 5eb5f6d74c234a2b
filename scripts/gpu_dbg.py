"""Debug run: C2 stream of argv[1] MiB through the library selected by DEFT4CU_LIB; prints the outcome."""
import os, sys, zlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200 import optimise_batch
mib = float(sys.argv[1]) if len(sys.argv) > 1 else 2
raw = W.c2_stream(int(mib * (1 << 20)))
t = time.time()
try:
    r = optimise_batch([raw], False)[0]
    print("ok", r["status"], r["saved_bits"], zlib.decompress(r["out"], -15) == zlib.decompress(raw, -15), round(time.time() - t, 2), "s")
except Exception as e:
    print("FAIL", e)
