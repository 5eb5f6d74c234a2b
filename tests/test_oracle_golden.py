"""Pins the CPU oracle against the reference's own golden files.

Each pair is (input, reference-optimised output) produced by the reference's runTestOpt.sh; the
`.txt` next to each golden is the reference's stdout (saved-bits lines).  The oracle, driven through
the container mirrors, must reproduce both byte for byte.
"""
import io
import zlib

import pytest

from conftest import GOLDEN_PAIRS, UNPAIRED_INPUTS, read_golden
from deft4j_b200.container import getContainerForBytes


@pytest.mark.parametrize("inp,gold,merge", GOLDEN_PAIRS, ids=[p[0] for p in GOLDEN_PAIRS])
def test_oracle_reproduces_reference_golden(oracle, inp, gold, merge):
    data = read_golden(inp)
    cont = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
    assert cont.read(data)
    log = io.StringIO()
    cont.optimise(merge, log)
    out = cont.write()
    assert out == read_golden(gold)
    ref_log = read_golden(gold + ".txt").decode()
    # the reference log has "File type recognised as ..." first and "Saved N bits with optimisation" last
    for line in log.getvalue().strip().splitlines():
        assert line in ref_log.splitlines()
    ours = [l for l in log.getvalue().strip().splitlines()]
    theirs = [l for l in ref_log.splitlines() if "bits saved" in l]
    assert ours == theirs


@pytest.mark.parametrize("inp", UNPAIRED_INPUTS)
def test_oracle_roundtrip_unpaired(oracle, inp):
    """Inputs without a golden: the optimised output must inflate to the same bytes and be no larger."""
    data = read_golden(inp)
    cont = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
    assert cont.read(data)
    before = [s.getUncompressedData() for s in cont.getDeflateStreams()]
    sizes = [s.getSizeBits() for s in cont.getDeflateStreams()]
    saved = cont.optimise(True, None)
    assert saved >= 0
    for s, b, sz in zip(cont.getDeflateStreams(), before, sizes):
        raw = s.asBytes()
        assert zlib.decompress(raw, -15) == b
        assert s.getSizeBits() <= sz
    cont.write()


def test_oracle_parse_matches_zlib(oracle):
    import random
    rnd = random.Random(1)
    words = [bytes(rnd.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rnd.randint(2, 9))) for _ in range(500)]
    text = b" ".join(rnd.choice(words) for _ in range(20000))
    for level, strategy in [(0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        raw = co.compress(text) + co.flush()
        s = oracle.OracleDeflateStream()
        assert s.parse(raw + b"trailer")
        assert s.consumed == len(raw)
        assert s.getUncompressedData() == text
        assert s.asBytes() == raw or zlib.decompress(s.asBytes(), -15) == text
        assert s.getSizeBits() <= len(raw) * 8 and s.getSizeBits() > len(raw) * 8 - 8


def test_oracle_rejects_garbage(oracle):
    s = oracle.OracleDeflateStream()
    assert not s.parse(b"\x07")          # reserved block type
    assert not s.parse(b"")              # EOF in header
    assert not s.parse(b"\x01\x05\x00\x00\x00")  # stored LEN/NLEN mismatch
    # backref before the start of the stream: fixed block, length 3 distance 1 as first symbol
    s2 = oracle.OracleDeflateStream()
    assert not s2.parse(bytes([0x03, 0x02, 0x00]) + b"\0" * 4)
