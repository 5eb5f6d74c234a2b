"""First-contact GPU diagnostics: parse + optimise parity against the oracle with verbose output."""
import io, os, sys, time, traceback, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from oracle_lib import OracleDeflateStream
from conftest import GOLDEN_PAIRS, UNPAIRED_INPUTS, read_golden
from deft4j_b200 import DeflateStream
from deft4j_b200.container import getContainerForBytes

sel = sys.argv[1:]


def cmp_parse(name, g, o):
    ok = True
    ng, no = g.blockCount(), o.blockCount()
    if ng != no:
        print("  [%s] block count gpu %d oracle %d" % (name, ng, no)); return False
    for i in range(no):
        a, b = g.blockInfo(i), o.blockInfo(i)
        for f in ("type", "size_bits", "position", "uncompressed_len", "n_symbols", "n_rle_pairs", "num_litlen_lens",
                  "num_dist_lens", "num_codelen_lens", "litlen_size_bits", "header_size_bits"):
            if getattr(a, f) != getattr(b, f):
                print("  [%s] block %d %s gpu %d oracle %d" % (name, i, f, getattr(a, f), getattr(b, f))); ok = False
        if a.type != 0 and ok:
            sa, sb = g.blockSymbols(i), o.blockSymbols(i)
            if sa != sb:
                k = next((k for k in range(min(len(sa), len(sb))) if sa[k] != sb[k]), None)
                print("  [%s] block %d symbols differ at %s: %s vs %s" % (name, i, k, sa[k] if k is not None else None, sb[k] if k is not None else None)); ok = False
            for w in (0, 1, 2):
                if g.blockCodelens(i, w) != o.blockCodelens(i, w):
                    print("  [%s] block %d codelens[%d] differ\n   gpu %s\n   ora %s" % (name, i, w, g.blockCodelens(i, w), o.blockCodelens(i, w))); ok = False
            if g.blockRlePairs(i) != o.blockRlePairs(i):
                print("  [%s] block %d rle pairs differ" % (name, i)); ok = False
    if g.getUncompressedData() != o.getUncompressedData():
        print("  [%s] uncompressed data differs" % name); ok = False
    return ok


def streams_of(data, name, cls):
    c = getContainerForBytes(data, name, cls)
    assert c.read(data), name
    return c


allok = True
for idx, (inp, gold, merge) in enumerate(GOLDEN_PAIRS):
    if sel and str(idx) not in sel:
        continue
    data = read_golden(inp)
    try:
        t0 = time.time()
        cg = streams_of(data, inp, DeflateStream)
        tp = time.time() - t0
        co = streams_of(data, inp, OracleDeflateStream)
        ok = True
        for k, (g, o) in enumerate(zip(cg.getDeflateStreams(), co.getDeflateStreams())):
            ok &= cmp_parse("%s#%d" % (inp, k), g, o)
            cs = g.getChecksums()
            d = o.getUncompressedData()
            if cs != (zlib.crc32(d) & 0xffffffff, zlib.adler32(d) & 0xffffffff, len(d)):
                print("  checksums differ", cs); ok = False
        print("PARSE %-40s %s (%.2fs)" % (inp, "ok" if ok else "MISMATCH", tp))
        t0 = time.time()
        log = io.StringIO()
        saved = cg.optimise(merge, log)
        out = cg.write()
        tg = time.time() - t0
        good = out == read_golden(gold)
        print("OPT   %-40s saved %d  %s  (%.2fs)" % (inp, saved, "BYTE-EXACT" if good else "MISMATCH", tg))
        if not good:
            allok = False
            so = co.optimise(merge, None)
            print("   oracle saved", so, "gpu log:", log.getvalue().strip().replace("\n", " | "))
            for k, (g, o) in enumerate(zip(cg.getDeflateStreams(), co.getDeflateStreams())):
                cmp_parse("%s#%d(after)" % (inp, k), g, o)
        allok &= ok
    except Exception:
        traceback.print_exc()
        allok = False
print("ALL OK" if allok else "FAILURES")
