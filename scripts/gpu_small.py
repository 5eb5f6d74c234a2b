"""Small-pool stress build on a few inputs with timing (debug)."""
import io, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W
from conftest import GOLDEN_PAIRS, read_golden
import deft4j_b200
from deft4j_b200.container import getContainerForBytes
for inp, gold, merge in GOLDEN_PAIRS:
    t = time.time()
    data = read_golden(inp)
    cont = getContainerForBytes(data, inp, deft4j_b200.DeflateStream)
    assert cont.read(data)
    cont.optimise(merge, io.StringIO())
    ok = cont.write() == read_golden(gold)
    print(inp, ok, round(time.time() - t, 2), flush=True)
co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
raw = co.compress(W.c2_text(300000)) + co.flush()
for merge in (False, True):
    t = time.time()
    r = deft4j_b200.optimise_batch([raw], merge)[0]
    print("c2", merge, r["saved_bits"], round(time.time() - t, 2), flush=True)
