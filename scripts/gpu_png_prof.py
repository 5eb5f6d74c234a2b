"""Host-side profile of the PNG container path (read_containers -> optimise_containers -> write) on the GPU box.
usage: python scripts/gpu_png_prof.py [files]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as W
from deft4j_b200.container import read_containers, optimise_containers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
files = W.c3_png_files(n, first=10_000)
names = ["f%d.png" % i for i in range(n)]


def run(tm=None):
    t0 = time.perf_counter()
    c = read_containers(files, names)
    t1 = time.perf_counter()
    optimise_containers(c, True)
    t2 = time.perf_counter()
    o = [x.write() for x in c]
    t3 = time.perf_counter()
    if tm is not None:
        tm.append((t1 - t0, t2 - t1, t3 - t2))
    return o


from deft4j_b200.container import optimise_png_files
for _ in range(8):
    print("---- native pass", file=sys.stderr, flush=True)
    t = time.perf_counter()
    r = optimise_png_files(files, True)
    print("native total %.1f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr, flush=True)
if len(sys.argv) > 2 and sys.argv[2] == "native":
    sys.exit(0)
run()
tm = []
for _ in range(6):
    print("---- pass", file=sys.stderr, flush=True)
    run(tm)
for t in tm:
    print("read %.1f ms  optimise %.1f ms  write %.1f ms  total %.1f ms  -> %.0f MB/s" %
          (t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, sum(t) * 1e3, sum(map(len, files)) / sum(t) / 1e6))
if len(sys.argv) > 2:
    pr = cProfile.Profile()
    pr.enable()
    run()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
