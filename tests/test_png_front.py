"""The native PNG / APNG front-end (deft4j_b200/csrc/png_front.cpp, `deft4cu_png_optimise_batch`) on the CPU: the shipped
chunk-model code linked over the oracle (tests/front_oracle_shim.cpp) instead of the device batch entry, checked against
the reference's golden files and, on structurally mutated files, against the Python mirror of PNGFile."""
import os
import random
import zlib

import pytest

import workloads as W
from conftest import read_golden

SIG = bytes([137, 80, 78, 71, 13, 10, 26, 10])


@pytest.fixture(scope="module")
def front():
    import hosttest_lib
    return hosttest_lib.front_oracle_lib()


def _chunks(b):
    pos, out = 8, []
    while pos < len(b):
        n = int.from_bytes(b[pos:pos + 4], "big")
        out.append((b[pos + 4:pos + 8], b[pos + 8:pos + 8 + n]))
        pos += 12 + n
    return out


def _ser(ch):
    o = bytearray(SIG)
    for t, d in ch:
        o += len(d).to_bytes(4, "big") + t + d + (zlib.crc32(d, zlib.crc32(t)) & 0xffffffff).to_bytes(4, "big")
    return bytes(o)


def mutants(per_base=30, seed=11):
    """Files with valid chunk CRCs whose chunk ORDER / content breaks (or does not break) PNGChunkHelper's rules."""
    rnd = random.Random(seed)
    ball = _chunks(read_golden("apng/ball.png"))
    fctl = [i for i, (t, _) in enumerate(ball) if t == b"fcTL"]
    small_apng = ball[:fctl[3]] + [ball[-1]]          # IDAT frame + two fdAT frames + IEND: cheap for the oracle
    bases = [(small_apng, per_base), (_chunks(read_golden("text.png")), per_base), (_chunks(W.c3_png_files(1)[0]), per_base // 2),
             (ball, 2), (_chunks(read_golden("284-edge-case/284.png")), 1)]
    out = []
    for ch, count in bases:
        out.append(_ser(ch))
        for _ in range(count):
            c = list(ch)
            r = rnd.randrange(12)
            i = rnd.randrange(len(c))
            if r == 0:
                del c[i]
            elif r == 1:
                c.insert(rnd.randrange(len(c) + 1), c[i])
            elif r == 2:
                j = rnd.randrange(len(c))
                c[i], c[j] = c[j], c[i]
            elif r == 3:
                t, d = c[i]
                if t == b"IDAT" and len(d) > 2:
                    k = rnd.randrange(1, len(d))
                    c[i:i + 1] = [(t, d[:k]), (t, d[k:])]
                elif t == b"fdAT" and len(d) > 8:
                    k = rnd.randrange(5, len(d))
                    c[i:i + 1] = [(t, d[:k]), (t, d[:4] + d[k:])]
            elif r == 4:
                c[i] = (c[i][0], b"")
            elif r == 5:
                t, d = c[i]
                if len(d) > 4:
                    d = bytearray(d)
                    d[rnd.randrange(len(d))] ^= 1 << rnd.randrange(8)
                    c[i] = (t, bytes(d))
            elif r == 6:
                c.insert(i, (b"zTXt", b"key\0\0" + zlib.compress(b"hello world " * 20)))
            elif r == 7:
                c.insert(i, (b"iTXt", b"key\0\1\0lang\0tr\0" + zlib.compress(b"hello world " * 20)))
            elif r == 8:
                c.insert(i, (b"iTXt", b"key\0\0\0lang\0tr\0plain text"))
            elif r == 9:
                c.insert(i, (b"zTXt", b"key\0\1" + zlib.compress(b"abc")))
            elif r == 10:
                c.insert(i, (b"iCCP", b"name\0\0" + zlib.compress(bytes(range(256)) * 3)[:-3]))
            else:
                c.insert(i, (b"zTXt", b"nonul"))
            out.append(_ser(c))
    return out


def mirror_outcome(data, stream_cls):
    """(status, bytes, saved, stream names) of the Python mirror of PNGFile (itself pinned on the goldens)."""
    from deft4j_b200.container import PNGFile
    c = PNGFile(stream_cls)
    try:
        ok = c.read(data)
    except Exception:  # noqa: BLE001 - the reference throws on the same inputs (array index)
        ok = False
    if not ok:
        return 1, None, 0, []
    saved = c.optimise(True, None)
    names = [s.getName() for s in c.getDeflateStreams()]
    try:
        return 0, c.write(), saved, names
    except IOError:
        return 2, None, saved, names


def check_against_mirror(files, res, stream_cls):
    seen = {}
    for data, r in zip(files, res):
        st, out, saved, names = mirror_outcome(data, stream_cls)
        assert r["status"] == st
        seen[st] = seen.get(st, 0) + 1
        if st == 0:
            assert r["out"] == out and r["saved_bits"] == saved and [n for n, _ in r["streams"]] == names
    return seen


def test_crc32_matches_zlib(front):
    rnd = random.Random(3)
    for n in [0, 1, 5, 15, 16, 63, 64, 127, 128, 129, 143, 144, 255, 256, 1000, 4096, 65537, 1 << 20]:
        d = rnd.randbytes(n)
        for off in (0, 1, 3):
            assert front.deft4cu_crc32(0, d[off:], max(0, n - off)) == zlib.crc32(d[off:])
        a = front.deft4cu_crc32(0, d[:n // 3], n // 3)
        assert front.deft4cu_crc32(a, d[n // 3:], n - n // 3) == zlib.crc32(d)


def test_reference_goldens_through_the_native_front_end(front):
    """runTestOpt.sh's PNG fixtures: ball.png (APNG, 20 streams), text.png (zTXt + iTXt), 284.png."""
    from deft4j_b200.container import optimise_png_files
    pairs = [("apng/ball.png", "apng/ball-opt.png"), ("text.png", "text-opt.png"), ("284-edge-case/284.png", "284-edge-case/284-opt.png")]
    files = [read_golden(a) for a, _ in pairs] + [b"junk", read_golden("text.png")[:200], b""]
    res = optimise_png_files(files, True, lib=front)
    for (a, g), r in zip(pairs, res):
        assert r["status"] == 0 and r["out"] == read_golden(g), a
    assert res[0]["streams"][0][0] == "IDAT chunk" and res[0]["streams"][1][0] == "fdAT chunk 1"
    assert [n for n, _ in res[1]["streams"]] == ["IDAT chunk", "zTXt chunk", "iTXt chunk"]
    assert [r["status"] for r in res[3:]] == [1, 1, 1]


def test_mutated_files_follow_the_mirror(front, oracle):
    from deft4j_b200.container import optimise_png_files
    files = mutants()
    res = optimise_png_files(files, True, lib=front)
    seen = check_against_mirror(files, res, oracle.OracleDeflateStream)
    assert seen.get(0, 0) >= 40 and seen.get(1, 0) >= 10


def test_pillow_files(front, oracle):
    from deft4j_b200.container import optimise_png_files
    files = W.c3_png_files(6, first=50)
    res = optimise_png_files(files, True, lib=front)
    assert check_against_mirror(files, res, oracle.OracleDeflateStream) == {0: 6}
