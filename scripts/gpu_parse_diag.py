"""Parse every deflate stream of the fixture set with the GPU library and with the oracle; report."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import OracleDeflateStream
from conftest import GOLDEN_PAIRS, UNPAIRED_INPUTS, read_golden
from deft4j_b200 import DeflateStream, _native
from deft4j_b200.container import getContainerForBytes

raws = []
class Capture(OracleDeflateStream):
    def parse(self, src):
        from deft4j_b200.container._io import ByteReader
        data = src.remaining() if isinstance(src, ByteReader) else bytes(src)
        ok = super().parse(src)
        raws.append((data, ok, self))
        return ok

names = [p[0] for p in GOLDEN_PAIRS] + UNPAIRED_INPUTS
for nm in names:
    if sys.argv[1:] and nm not in sys.argv[1:]:
        continue
    data = read_golden(nm)
    del raws[:]
    c = getContainerForBytes(data, nm, Capture)
    c.read(data)
    for k, (raw, ok, o) in enumerate(raws):
        g = DeflateStream()
        okg = g.parse(raw)
        line = "%-40s stream %d len %d oracle %s gpu %s consumed %d/%d" % (nm, k, len(raw), ok, okg, g.consumed, o.consumed)
        if ok and okg:
            same = g.getUncompressedData() == o.getUncompressedData()
            line += " data %s blocks %d/%d bits %d/%d" % (same, g.blockCount(), o.blockCount(), g.getSizeBits(), o.getSizeBits())
        else:
            line += " err=" + _native.last_error()
            if ok:
                line += " oracle blocks: " + ",".join("%d:%d" % (o.blockInfo(i).type, o.blockInfo(i).n_symbols) for i in range(o.blockCount()))
        print(line)
