"""The native gzip / zlib front-ends (deft4j_b200/csrc/gz_front.cpp) on the CPU: the shipped header code linked over the
oracle (tests/front_oracle_shim.cpp), against the reference's golden files and the Python mirrors of GZFile / ZLibFile."""
import gzip
import io
import zlib

import pytest

import workloads as W
from conftest import GOLDEN_PAIRS, UNPAIRED_INPUTS, read_golden


@pytest.fixture(scope="module")
def front():
    import hosttest_lib
    return hosttest_lib.front_oracle_lib()


def _gz(data, name=b"", comment=b"", extra=b"", hcrc=False, level=6, mtime=0x12345678, xfl=2, os_=3):
    flags = 0
    if extra:
        flags |= 4
    if name is not None:
        flags |= 8
    if comment:
        flags |= 16
    if hcrc:
        flags |= 2
    out = bytearray([0x1f, 0x8b, 8, flags]) + mtime.to_bytes(4, "little") + bytes([xfl, os_])
    if extra:
        out += len(extra).to_bytes(2, "little") + extra
    if name is not None:
        out += name + b"\0"
    if comment:
        out += comment + b"\0"
    if hcrc:
        out += b"\xab\xcd"
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    out += co.compress(data) + co.flush()
    out += (zlib.crc32(data) & 0xffffffff).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")
    return bytes(out)


def gz_files():
    text = W.c2_text(30000, seed=9)
    good = [_gz(text[:8000], name=None), _gz(text[:9000], name=b"a.txt"), _gz(text[:7000], name=b""),      # FNAME with an empty name
            _gz(text[100:5000], name=b"n", comment=b"a comment"), _gz(text[:4000], name=None, extra=b"EXTRA!", hcrc=True),
            _gz(text[:6000], name=b"x.bin", comment=b"c", extra=b"12", hcrc=True, level=1),
            _gz(b"", name=b"empty.txt"), _gz(text[:3000], name=b"two") + _gz(text[:100], name=b"second member")]
    g = good[1]
    bad = [b"", b"\x1f", g[:3], g[:9], g[:12], g[:14], bytes([0x1f, 0x8b, 7]) + g[3:], g[:3] + bytes([0x20 | g[3]]) + g[4:],
           good[4][:13], good[3][:14], g[:-9], g[:-8], g[:-3], b"PK\x03\x04 not gzip", g[:16] + b"\xff" + g[17:]]
    return good + bad


def zlib_files():
    text = W.c2_text(20000, seed=4)
    good = [zlib.compress(text[:9000], 6), zlib.compress(text[:5000], 1), zlib.compress(text[:7000], 9), zlib.compress(b"")]
    z = good[0]
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY, b"dictionary")
    with_dict = co.compress(text[:3000]) + co.flush()
    bad = [b"", z[:1], z[:2], z[:20], bytes([0x79]) + z[1:], bytes([z[0], z[1] ^ 1]) + z[2:], with_dict, z[:-4], z[:-1]]
    return good + bad


def mirror_outcome(cls, data, stream_cls):
    c = cls(stream_cls)
    if not c.read(data):
        return 1, None, 0, []
    saved = c.optimise(True, None)
    return 0, c.write(), saved, [s.getName() for s in c.getDeflateStreams()]


def check(files, res, cls, stream_cls):
    seen = {}
    for k, (data, r) in enumerate(zip(files, res)):
        st, out, saved, names = mirror_outcome(cls, data, stream_cls)
        assert r["status"] == st, k
        seen[st] = seen.get(st, 0) + 1
        if st == 0:
            assert r["out"] == out and r["saved_bits"] == saved and [n for n, _ in r["streams"]] == names, k
    return seen


def test_gzip_goldens_and_mutants(front, oracle):
    from deft4j_b200.container import GZFile, optimise_gz_files
    pairs = [(a, g) for a, g, fast in GOLDEN_PAIRS if a.endswith(".gz") and fast and "zopfli" not in a]
    files = [read_golden(a) for a, _ in pairs]
    res = optimise_gz_files(files, True, lib=front)
    for (a, g), r in zip(pairs, res):
        assert r["status"] == 0 and r["out"] == read_golden(g), a
    assert len(pairs) >= 3
    files = gz_files() + [read_golden(x) for x in UNPAIRED_INPUTS if x.endswith(".gz") and "asyoulik" not in x]
    res = optimise_gz_files(files, True, lib=front)
    seen = check(files, res, GZFile, oracle.OracleDeflateStream)
    assert seen.get(0, 0) >= 10 and seen.get(1, 0) >= 8
    # what gzip itself makes of the rewritten members (not the ones with FCOMMENT: the reference writes the comment back
    # without its NUL, GZFile.java:117-119; not the two-member file: only the first member is rewritten)
    for k in (0, 1, 2, 4, 6):
        assert gzip.GzipFile(fileobj=io.BytesIO(res[k]["out"])).read() == gzip.GzipFile(fileobj=io.BytesIO(files[k])).read()


def test_zlib_files_follow_the_mirror(front, oracle):
    from deft4j_b200.container import ZLibFile, optimise_zlib_files
    files = zlib_files() + [read_golden("deflate-fixed.txt.zz")]
    res = optimise_zlib_files(files, True, lib=front)
    seen = check(files, res, ZLibFile, oracle.OracleDeflateStream)
    assert seen.get(0, 0) >= 5 and seen.get(1, 0) >= 5
    for data, r in zip(files[:4], res[:4]):
        assert zlib.decompress(r["out"]) == zlib.decompress(data)
