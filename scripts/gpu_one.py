"""One golden input through the library selected by DEFT4CU_LIB (debug)."""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import read_golden
import deft4j_b200
from deft4j_b200.container import getContainerForBytes
inp = sys.argv[1]; merge = int(sys.argv[2])
data = read_golden(inp)
cont = getContainerForBytes(data, inp, deft4j_b200.DeflateStream)
assert cont.read(data)
t = time.time()
try:
    cont.optimise(bool(merge), io.StringIO())
    print(inp, "done", round(time.time() - t, 2), flush=True)
except Exception as e:
    print("FAIL", e, flush=True)
