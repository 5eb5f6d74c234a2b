"""One optimise pass over a C2 stream (size in MiB, default 16) through the batch entry — the command ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200 import optimise_batch
mib = float(sys.argv[1]) if len(sys.argv) > 1 else 16
merge = len(sys.argv) > 2 and sys.argv[2] == "merge"
raw = W.c2_stream(int(mib * (1 << 20)))
r = optimise_batch([raw], merge)[0]
print("saved", r["saved_bits"], "status", r["status"])
