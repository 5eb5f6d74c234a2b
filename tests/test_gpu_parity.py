"""Parity tests proper (run on the B200): the CUDA path, called through the C ABI (ctypes -> libdeft4cu.so),
against the CPU oracle and the reference's committed golden files.  Integer/byte work: the bar is bit-exact.
"""
import io
import os
import zlib

import pytest

import workloads as W
from conftest import GOLDEN_PAIRS, UNPAIRED_INPUTS, read_golden

pytestmark = pytest.mark.gpu

INFO_FIELDS = ("type", "size_bits", "position", "uncompressed_len", "n_symbols", "n_rle_pairs", "num_litlen_lens",
               "num_dist_lens", "num_codelen_lens", "litlen_size_bits", "header_size_bits")


@pytest.fixture(scope="module")
def gpu():
    import deft4j_b200
    from deft4j_b200 import _native
    _native.lib()  # raises when the CUDA library or a device is missing: no fallback
    return deft4j_b200


def assert_same_model(g, o, symbols=True):
    assert g.blockCount() == o.blockCount()
    for i in range(o.blockCount()):
        a, b = g.blockInfo(i), o.blockInfo(i)
        for f in INFO_FIELDS:
            if f in ("num_litlen_lens", "num_dist_lens", "num_codelen_lens") and b.type != 2:
                continue
            assert getattr(a, f) == getattr(b, f), (i, f, getattr(a, f), getattr(b, f))
        if a.type != 0 and symbols:
            assert g.blockSymbols(i) == o.blockSymbols(i), i
            if a.type == 2:
                for w in (0, 1, 2):
                    assert g.blockCodelens(i, w) == o.blockCodelens(i, w), (i, w)
                assert g.blockRlePairs(i) == o.blockRlePairs(i), i


def compare_stream(gpu, oracle, raw, merge, check_model=True):
    g = gpu.DeflateStream()
    o = oracle.OracleDeflateStream()
    okg, oko = g.parse(raw + b"TRAILER"), o.parse(raw + b"TRAILER")
    assert okg == oko
    if not oko:
        return None
    assert g.consumed == o.consumed
    assert g.getSizeBits() == o.getSizeBits()
    data = o.getUncompressedData()
    assert g.getUncompressedData() == data
    assert g.getChecksums() == (zlib.crc32(data) & 0xffffffff, zlib.adler32(data) & 0xffffffff, len(data))
    if check_model:
        assert_same_model(g, o)
    sg, so = g.optimise(merge), o.optimise(merge)
    assert sg == so
    out = g.asBytes()
    assert out == o.asBytes()
    assert g.getSizeBits() == o.getSizeBits()
    if check_model:
        assert_same_model(g, o)
    assert zlib.decompress(out, -15) == data
    return sg


# ---- C1: the reference's own fixtures --------------------------------------------------------------------------
@pytest.mark.parametrize("inp,gold,merge", GOLDEN_PAIRS, ids=[p[0] for p in GOLDEN_PAIRS])
def test_reference_golden_pairs(gpu, inp, gold, merge):
    """`deft4j optimise -m NONE` on the repo's fixtures with runTestOpt.sh's flags: container bytes and the
    saved-bits log lines identical to the reference's committed outputs."""
    from deft4j_b200.container import getContainerForBytes
    data = read_golden(inp)
    cont = getContainerForBytes(data, inp, gpu.DeflateStream)
    assert cont.read(data)
    log = io.StringIO()
    cont.optimise(merge, log)
    assert cont.write() == read_golden(gold)
    ref_log = [l for l in read_golden(gold + ".txt").decode().splitlines() if "bits saved" in l]
    assert log.getvalue().strip().splitlines() == ref_log


@pytest.mark.parametrize("inp", [p[0] for p in GOLDEN_PAIRS] + UNPAIRED_INPUTS)
def test_parse_model_matches_oracle(gpu, oracle, inp):
    """Decode parity on every fixture: block list, per-symbol LitLen records, code tables, header RLE pairs,
    bit sizes, decoded bytes, CRC-32/Adler-32 (computed on the device)."""
    from deft4j_b200.container import getContainerForBytes
    data = read_golden(inp)
    cg = getContainerForBytes(data, inp, gpu.DeflateStream)
    co = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
    assert cg.read(data) and co.read(data)
    assert len(cg.getDeflateStreams()) == len(co.getDeflateStreams())
    for g, o in zip(cg.getDeflateStreams(), co.getDeflateStreams()):
        assert_same_model(g, o)
        d = o.getUncompressedData()
        assert g.getUncompressedData() == d
        assert g.getChecksums() == (zlib.crc32(d) & 0xffffffff, zlib.adler32(d) & 0xffffffff, len(d))


@pytest.mark.parametrize("inp", ["ban.txt.gz", "lz.txt.gz", "deflate-store.txt.gz", "deflate-fixed.txt.zz",
                                 "deflate-dynamic.txt.gz", "asyoulik/asyoulik-zopfli-extopt.txt.gz"])
def test_unpaired_fixtures_match_oracle(gpu, oracle, inp):
    from deft4j_b200.container import getContainerForBytes
    data = read_golden(inp)
    cg = getContainerForBytes(data, inp, gpu.DeflateStream)
    co = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
    assert cg.read(data) and co.read(data)
    assert cg.optimise(True, None) == co.optimise(True, None)
    assert cg.write() == co.write()


# ---- C2..C5 at sizes the oracle finishes in seconds ---------------------------------------------------------------
def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return co.compress(data) + co.flush()


@pytest.mark.parametrize("merge", [False, True])
def test_c2_text_stream(gpu, oracle, merge):
    raw = _deflate(W.c2_text(200_000))
    assert compare_stream(gpu, oracle, raw, merge) > 0


def _oracle_batch(streams, merge):
    """The oracle on every host core (it is the checker here, never the path under test)."""
    import multiprocessing as mp
    from bench import _oracle_out
    with mp.get_context("spawn").Pool(min(len(streams), os.cpu_count() or 1)) as pool:
        return pool.map(_oracle_out, [(s, merge) for s in streams], chunksize=1)


def _compare_batch(gpu, streams, merge):
    """One device batch against the oracle, stream by stream: saved bits, output bytes, consumed bytes, checksums."""
    res = gpu.optimise_batch(streams, merge)
    ref = _oracle_batch(streams, merge)
    for k, (raw, r, (saved, out, consumed, crc, adler, ulen)) in enumerate(zip(streams, res, ref)):
        assert r["status"] == 0, k
        assert (r["saved_bits"], r["consumed"]) == (saved, consumed), k
        assert r["out"] == out, k
        assert (r["crc32"], r["adler32"], r["uncompressed_len"]) == (crc, adler, ulen), k
        assert zlib.decompress(r["out"], -15) == zlib.decompress(raw, -15), k


def test_c3_png_idat_streams(gpu, oracle):
    for raw in W.c3_streams(3):
        compare_stream(gpu, oracle, raw, True, check_model=False)
    _compare_batch(gpu, W.c3_streams(64, first=2000), True)


def test_c4_entry_mix(gpu, oracle):
    for raw in W.c4_streams(12, seed=44):
        compare_stream(gpu, oracle, raw, True, check_model=False)
    _compare_batch(gpu, W.c4_streams(128, seed=45), True)


def test_c5_adversarial(gpu, oracle):
    streams = [s for s in W.c5_streams() if len(s) < 200_000]
    assert len(streams) >= 10
    for raw in streams:
        compare_stream(gpu, oracle, raw, True, check_model=False)


def test_c5_full_size_shapes(gpu, oracle):
    """The 64 MiB long-match shapes of SURVEY.md 8(d) C5(i): len 258 / dist 1 and dist 32768.  The whole stream is
    checked through inflate + device checksums + "no larger"; the first 1 MiB of text through the oracle."""
    rng_period = __import__("numpy").random.default_rng(5).integers(0, 256, 32768, dtype="uint8").tobytes()
    n = 64 << 20
    for data in (b"\x41" * n, (rng_period * (n // 32768 + 1))[:n]):
        for merge in (False, True):
            raw = _deflate(data, 6, memlevel=9)
            r = gpu.optimise_batch([raw], merge)[0]
            assert r["status"] == 0 and r["consumed"] == len(raw)
            assert len(r["out"]) <= len(raw)
            assert (r["uncompressed_len"], r["crc32"], r["adler32"]) == (n, zlib.crc32(data) & 0xffffffff, zlib.adler32(data) & 0xffffffff)
            assert zlib.decompress(r["out"], -15) == data
        compare_stream(gpu, oracle, _deflate(data[:1 << 20], 6, memlevel=9), True, check_model=False)


# ---- C2 at sizes the oracle needs minutes for: its outputs are committed as hashes (scripts/make_c2_golden.py) ---------
@pytest.mark.parametrize("case", ["c2_4MiB_nomerge", "c2_8MiB_nomerge_seed2", "c2_640KiB_merge", "c2_640KiB_nomerge"])
def test_c2_against_committed_oracle_hashes(gpu, case):
    import hashlib, json
    with open(os.path.join(os.path.dirname(__file__), "golden", "c2_oracle_hashes.json")) as f:
        g = json.load(f)[case]
    raw = W._c2_stream(g["target_bytes"], g["seed"])
    assert (len(raw), hashlib.sha256(raw).hexdigest()) == (g["in_len"], g["in_sha256"]), "generator drifted: regenerate the golden"
    r = gpu.optimise_batch([raw], g["merge"])[0]
    assert r["status"] == 0
    assert r["saved_bits"] == g["saved_bits"]
    assert (len(r["out"]), hashlib.sha256(r["out"]).hexdigest()) == (g["out_len"], g["out_sha256"])
    assert zlib.decompress(r["out"], -15) == zlib.decompress(raw, -15)


# ---- H10: code-length sets zlib rejects but the reference's first-match decoder accepts ----------------------------------
@pytest.mark.parametrize("name", sorted(W.odd_code_streams()))
@pytest.mark.parametrize("merge", [False, True])
def test_odd_code_sets_follow_the_reference(gpu, oracle, name, merge):
    """Incomplete and over-subscribed code-length sets decode exactly like Huffman.readSymbol (Huffman.java:170-197:
    by length, first match wins, no validity check), and a block that keeps such a header is written back as is."""
    raw = W.odd_code_streams()[name]
    g, o = gpu.DeflateStream(), oracle.OracleDeflateStream()
    assert o.parse(raw + b"xx") and g.parse(raw + b"xx")
    assert g.consumed == o.consumed and g.getSizeBits() == o.getSizeBits()
    assert g.getUncompressedData() == o.getUncompressedData()
    assert_same_model(g, o)
    assert g.optimise(merge) == o.optimise(merge)
    assert g.asBytes() == o.asBytes()
    assert_same_model(g, o)


# ---- DeflateStream.optimise called again on the same object ------------------------------------------------------------
@pytest.mark.parametrize("seq", [(False, True), (True, True), (False, False, True)])
def test_repeated_optimise_calls(gpu, oracle, seq):
    streams = [_deflate(W.c2_text(90_000, seed=31)), W.handmade_streams()["dyn_partialflush_x3"],
               W.handmade_streams()["empty_blocks_midstream"], W.handmade_streams()["stored_around_65535"],
               W.c4_streams(3, seed=12)[1]]
    for raw in streams:
        g, o = gpu.DeflateStream(), oracle.OracleDeflateStream()
        assert g.parse(raw) and o.parse(raw)
        for merge in seq:
            assert g.optimise(merge) == o.optimise(merge)
            assert g.asBytes() == o.asBytes()
            assert g.getSizeBits() == o.getSizeBits()
        assert_same_model(g, o)


# ---- the tie-break contract, candidate by candidate ----------------------------------------------------------------------
def _split_calls(pairs):
    calls = []
    for idx, sz in pairs:
        if idx == -1:
            calls.append([sz, {}])
        else:
            calls[-1][1][idx] = sz
    return calls


def test_candidate_trace_matches_oracle(gpu, oracle):
    """Every candidate the selection callback of DeflateStream.optimiseBlock sees (DeflateStream.java:349-368), as
    (enumeration index, size in bits), for every optimiseBlock call of the stream: the oracle's list and the CUDA
    enumerator's must agree index by index (H7), including the index bookkeeping of the skipped
    addOptimisedRecoded(prune) sweep, which logs nothing but advances the index."""
    import ctypes as C
    from deft4j_b200 import _native
    from deft4j_b200.container import getContainerForBytes
    CAP = 3_000_000
    OL, GL = oracle.lib(), _native.lib()
    OL.ora_trace_begin.argtypes = [C.POINTER(C.c_int64), C.c_size_t]
    OL.ora_trace_end.restype = C.c_size_t
    raws = []
    for inp, take in (("asyoulik/asyoulik-gzip.txt.gz", 1), ("apng/ball.png", 3), ("text.png", 3)):
        data = read_golden(inp)
        co = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
        assert co.read(data)
        raws += [s.asBytes() for s in co.getDeflateStreams()][:take]
    raws += [_deflate(W.c2_text(5000, seed=3), 6, zlib.Z_FIXED), W.handmade_streams()["edge284"],
             W.handmade_streams()["dyn_partialflush_dyn"]]
    for k, raw in enumerate(raws):
        o = oracle.OracleDeflateStream()
        assert o.parse(raw)
        buf = (C.c_int64 * (2 * CAP))()
        OL.ora_trace_begin(buf, CAP)
        so = o.optimise(False)
        n = OL.ora_trace_end()
        assert n < CAP
        ref = _split_calls([(buf[2 * i], buf[2 * i + 1]) for i in range(n)])
        g = gpu.DeflateStream()
        assert g.parse(raw)
        assert GL.deft4cu_debug_trace_begin(CAP) == 0
        sg = g.optimise(False)
        gbuf = (C.c_int64 * (2 * CAP))()
        gn = C.c_uint32(0)
        assert GL.deft4cu_debug_trace_end(gbuf, CAP, C.byref(gn)) == 0
        assert gn.value < CAP
        got = _split_calls([(gbuf[2 * i], gbuf[2 * i + 1]) for i in range(gn.value)])
        assert sg == so
        assert len(got) == len(ref), (k, len(got), len(ref))
        for call, ((ri, rc), (gi, gc)) in enumerate(zip(ref, got)):
            assert ri == gi, (k, call, "incumbent")
            # the device logs every candidate except the skipped sweep's (whose indices it must still skip over)
            assert gc, (k, call)
            for idx, sz in gc.items():
                assert rc.get(idx) == sz, (k, call, idx, rc.get(idx), sz)
            assert max(gc) <= max(rc) and len(rc) - len(gc) <= 56 * 4 * 16, (k, call, len(rc), len(gc))
            # the chosen candidate: first strict minimum over the oracle's full list is in the device's list too
            best = min(rc.values())
            first = min(i for i, v in rc.items() if v == best)
            if best < ri:
                assert gc.get(first) == best, (k, call, first)


@pytest.mark.parametrize("name", sorted(W.handmade_streams()))
@pytest.mark.parametrize("merge", [False, True])
def test_handmade_shapes(gpu, oracle, name, merge):
    """distance 32768 / length 258, the 284+31 spelling of 258 (H9), empty blocks mid-stream (H6), stored blocks
    around 65535 (H5/H12), lone EOB, empty stored block."""
    compare_stream(gpu, oracle, W.handmade_streams()[name], merge)


@pytest.mark.parametrize("level,strategy", [(0, 0), (1, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_zlib_flavours(gpu, oracle, level, strategy):
    compare_stream(gpu, oracle, _deflate(W.c2_text(70_000, seed=level * 10 + strategy + 1), level, strategy), True,
                   check_model=False)


def test_sync_flush_empty_stored_blocks(gpu, oracle):
    """Z_SYNC_FLUSH / Z_FULL_FLUSH leave empty stored blocks mid-stream: the reference's loop stops at the first one
    it removes."""
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    t = W.c2_text(30_000, seed=9)
    raw = co.compress(t[:10_000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(t[10_000:20_000]) + \
        co.flush(zlib.Z_FULL_FLUSH) + co.compress(t[20_000:]) + co.flush()
    for merge in (False, True):
        compare_stream(gpu, oracle, raw, merge)


# ---- parse failures and edge inputs --------------------------------------------------------------------------------
@pytest.mark.parametrize("raw", [b"", b"\x07", b"\x01\x05\x00\x00\x00", bytes([0x03, 0x02, 0x00]) + b"\0" * 4,
                                 b"\x05", b"\xff" * 40], ids=["empty", "btype3", "len_nlen", "dist_before_start",
                                                             "truncated_dynamic", "garbage"])
def test_parse_failures(gpu, oracle, raw):
    g, o = gpu.DeflateStream(), oracle.OracleDeflateStream()
    assert o.parse(raw) is False
    assert g.parse(raw) is False


def test_truncated_stream_fails(gpu, oracle):
    raw = _deflate(W.c2_text(20_000))
    for cut in (len(raw) // 2, len(raw) - 1, 5):
        g, o = gpu.DeflateStream(), oracle.OracleDeflateStream()
        assert o.parse(raw[:cut]) is False
        assert g.parse(raw[:cut]) is False


# ---- the entry points of include/deft4cu.h -----------------------------------------------------------------------
def test_batch_entry_matches_handles(gpu, oracle):
    streams = W.c4_streams(9, seed=7) + [b"\x07", W.handmade_streams()["edge284"]]
    res = gpu.optimise_batch(streams, True)
    for raw, r in zip(streams, res):
        o = oracle.OracleDeflateStream()
        if not o.parse(raw):
            assert r["status"] == 1
            continue
        assert r["status"] == 0 and r["consumed"] == o.consumed
        before = o.getSizeBits()
        assert r["size_bits_in"] == before
        assert r["saved_bits"] == o.optimise(True)
        assert r["out"] == o.asBytes()
        assert r["size_bits_out"] == o.getSizeBits()
        d = o.getUncompressedData()
        assert (r["uncompressed_len"], r["crc32"], r["adler32"]) == (len(d), zlib.crc32(d) & 0xffffffff, zlib.adler32(d) & 0xffffffff)


def test_facade_identity_semantics(gpu):
    """Deft.optimiseDeflateStream returns the SAME array when nothing is saved or the stream does not parse (Deft.java:25-33)."""
    already = gpu.Deft.optimiseDeflateStream(_deflate(W.c2_text(30_000)), True)
    assert gpu.Deft.optimiseDeflateStream(already, True) is already
    junk = b"\x07junk"
    assert gpu.Deft.optimiseDeflateStream(junk, True) is junk
    raw = _deflate(W.c2_text(30_000))
    out = gpu.Deft.optimiseDeflateStream(raw, True)
    assert out is not raw and len(out) <= len(raw) and zlib.decompress(out, -15) == zlib.decompress(raw, -15)
    assert gpu.Deft.getSizeBitsFallback(junk) == len(junk) * 8
    assert len(raw) * 8 - 8 < gpu.Deft.getSizeBitsFallback(raw) <= len(raw) * 8


def test_parse_batch_and_partial_optimise(gpu, oracle):
    streams = W.c3_streams(4, first=100)
    hs = gpu.DeflateStream.parse_batch(streams)
    assert all(h is not None for h in hs)
    saved = gpu.DeflateStream.optimise_batch(hs[1:3], True)
    for k, raw in enumerate(streams):
        o = oracle.OracleDeflateStream()
        assert o.parse(raw)
        if k in (1, 2):
            assert o.optimise(True) == saved[k - 1]
        assert hs[k].asBytes() == o.asBytes()


# ---- full-size properties (no oracle: size-independent invariants) ----------------------------------------------------
def test_c2_large_stream_properties(gpu):
    """16 MiB single stream (about 650 dynamic blocks), mergeBlocks=false like the benchmark: the rewritten stream
    inflates to the same bytes, is no larger, and the device checksums match zlib's."""
    raw = W.c2_stream(16 << 20)
    data = zlib.decompress(raw, -15)
    r = gpu.optimise_batch([raw], False)[0]
    assert r["status"] == 0 and r["consumed"] == len(raw)
    assert zlib.decompress(r["out"], -15) == data
    assert r["size_bits_out"] == r["size_bits_in"] - r["saved_bits"]
    assert 0 < r["saved_bits"] and len(r["out"]) < len(raw)
    assert (r["uncompressed_len"], r["crc32"], r["adler32"]) == (len(data), zlib.crc32(data) & 0xffffffff, zlib.adler32(data) & 0xffffffff)
    # a second pass over the rewritten stream never grows it and still round-trips (on the samples the oracle
    # can finish it saves exactly 0: every block already sits at the enumerator's fix point)
    r2 = gpu.optimise_batch([r["out"]], False)[0]
    assert r2["saved_bits"] >= 0 and len(r2["out"]) <= len(r["out"]) and zlib.decompress(r2["out"], -15) == data


def test_c3_batch_properties(gpu):
    streams = W.c3_streams(300, first=1000)
    res = gpu.optimise_batch(streams, True)
    for raw, r in zip(streams, res):
        assert r["status"] == 0
        assert zlib.decompress(r["out"], -15) == zlib.decompress(raw, -15)
        assert len(r["out"]) <= len(raw)


# ---- engine pool overflow: the same library built with tiny pools (-DD4_SMALL_POOLS) -----------------------------
_SMALLPOOL_SCRIPT = r"""
import io, sys, zlib
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import workloads as W, oracle_lib
from conftest import GOLDEN_PAIRS, read_golden
import deft4j_b200
from deft4j_b200.container import getContainerForBytes
for inp, gold, merge in GOLDEN_PAIRS:
    data = read_golden(inp)
    cont = getContainerForBytes(data, inp, deft4j_b200.DeflateStream)
    assert cont.read(data)
    cont.optimise(merge, io.StringIO())
    assert cont.write() == read_golden(gold), inp
co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
raw = co.compress(W.c2_text(300000)) + co.flush()
for merge in (False, True):
    r = deft4j_b200.optimise_batch([raw], merge)[0]
    o = oracle_lib.OracleDeflateStream(); assert o.parse(raw)
    assert r["saved_bits"] == o.optimise(merge) and r["out"] == o.asBytes()
print("SMALLPOOLS-OK")
"""


def test_engine_pool_overflow_path(gpu):
    """The engine's hash-consed mask/table pools and pass memo are flushed when full; with pools of 20 entries
    every block of the fixtures overflows them many times, and the output must not change."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "deft4j_b200", "libdeft4cu_smallpools.so")
    assert os.path.exists(lib), "run __graft_entry__.build()"
    env = dict(os.environ, DEFT4CU_LIB=lib)
    p = subprocess.run([sys.executable, "-c", _SMALLPOOL_SCRIPT % {"root": root, "tests": os.path.join(root, "tests")}],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "SMALLPOOLS-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


# ---- intra-stream parallel parse: tiny segments force many walkers, wrong guesses and re-walks --------------------
@pytest.mark.parametrize("seg", [64, 200, 1000, 4096])
def test_segmented_parse_matches_oracle(gpu, oracle, seg, monkeypatch):
    """The stream is cut into `seg`-byte segments (default 128 KiB), every segment's walker guesses its first block
    boundary on the device and the host follows the chain from bit 0.  Whatever the guesses, the parsed model and
    the optimised bytes must equal the single-walker result: dynamic blocks (found), fixed / stored blocks (never
    guessed: re-walked), empty blocks, blocks spanning many segments, truncated streams."""
    monkeypatch.setenv("D4_SEG_BYTES", str(seg))
    from deft4j_b200.container import getContainerForBytes
    streams = []
    for inp in ["asyoulik/asyoulik-gzip.txt.gz", "apng/ball.png", "deflate-store-2.txt.gz", "deflate-fixed.txt.gz"]:
        data = read_golden(inp)
        co = getContainerForBytes(data, inp, oracle.OracleDeflateStream)
        assert co.read(data)
        streams += [s.asBytes() for s in co.getDeflateStreams()][:3]
    streams += [_deflate(W.c2_text(150_000, seed=5), 1), _deflate(W.c2_text(60_000, seed=6), 6, zlib.Z_FIXED),
                _deflate(W.c2_text(200_000, seed=7), 0), _deflate(W.c2_text(120_000, seed=8), 9)]
    streams += list(W.handmade_streams().values())
    streams += [s for s in W.c5_streams() if len(s) < 60_000][:6]
    for raw in streams:
        compare_stream(gpu, oracle, raw, False, check_model=True)
    # truncated input: still a parse failure, never a hang or a wrong chain
    raw = _deflate(W.c2_text(50_000, seed=11))
    for cut in (len(raw) // 3, len(raw) - 2):
        assert gpu.DeflateStream().parse(raw[:cut]) is False
    # one batch with everything: walkers of different streams side by side
    res = gpu.optimise_batch(streams, True)
    monkeypatch.delenv("D4_SEG_BYTES")
    ref = gpu.optimise_batch(streams, True)
    for a, b in zip(res, ref):
        assert (a["status"], a["saved_bits"], a["out"], a["crc32"]) == (b["status"], b["saved_bits"], b["out"], b["crc32"])


def test_literal_cost_paths_agree(gpu, oracle, monkeypatch):
    """The engine takes a match's literal cost from prefix sums over the block's decoded bytes when matches are long
    (>= 24 decoded bytes per symbol) and the scratch fits, and walks the match's bytes otherwise: forced either way,
    both must give the oracle's bytes."""
    raw = _deflate(W.c2_text(120_000, seed=21))
    batch = [raw] + W.c3_streams(2, first=7)
    monkeypatch.setenv("D4_NO_PREFIX", "1")
    compare_stream(gpu, oracle, raw, True, check_model=False)
    loops = gpu.optimise_batch(batch, True)
    monkeypatch.delenv("D4_NO_PREFIX")
    monkeypatch.setenv("D4_PREFIX_RATIO", "0")   # prefix sums for every block
    compare_stream(gpu, oracle, raw, True, check_model=False)
    sums = gpu.optimise_batch(batch, True)
    monkeypatch.delenv("D4_PREFIX_RATIO")
    auto = gpu.optimise_batch(batch, True)
    key = lambda rs: [(r["saved_bits"], r["out"]) for r in rs]
    assert key(loops) == key(sums) == key(auto)


def test_zip_container_and_cli_on_the_gpu(gpu, oracle, tmp_path, capsys):
    """SURVEY.md §8f rows 3-4 with the CUDA stream engine: the ZIP container mirror batches its entries through
    deft4cu_stream_optimise_batch, the CLI mirror reproduces a reference golden byte for byte."""
    import zipfile
    from deft4j_b200.container import getContainerForBytes
    from deft4j_b200.__main__ import main
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as z:
        for k in range(5):
            zi = zipfile.ZipInfo("dir/f%d.txt" % k, date_time=(2024, 5, 6, 7, 8, 10))
            zi.compress_type = zipfile.ZIP_DEFLATED if k != 3 else zipfile.ZIP_STORED
            z.writestr(zi, W.c2_text(9000 + 7000 * k, seed=40 + k), compresslevel=[6, 1, 9, None, 6][k])
    data = buf.getvalue()
    cg = getContainerForBytes(data, "a.zip", gpu.DeflateStream)
    co = getContainerForBytes(data, "a.zip", oracle.OracleDeflateStream)
    assert cg.read(data) and co.read(data)
    assert cg.optimise(True, None) == co.optimise(True, None)
    out = cg.write()
    assert out == co.write()
    with zipfile.ZipFile(io.BytesIO(out)) as z, zipfile.ZipFile(io.BytesIO(data)) as z0:
        assert z.testzip() is None and [z.read(n) for n in z.namelist()] == [z0.read(n) for n in z0.namelist()]
    dst = tmp_path / "o.gz"
    assert main(["optimise", os.path.join(os.path.dirname(__file__), "golden", "asyoulik", "asyoulik-gzip.txt.gz"), str(dst)]) == 0
    assert dst.read_bytes() == read_golden("asyoulik/asyoulik-gzip-opt.txt.gz")
    assert "167 bits saved in stream 0" in capsys.readouterr().out


# ---- the stream LIST is the batch: one launch of the candidate engine per container / folder / list -------------------
def test_container_list_is_one_device_batch(gpu, tmp_path, capsys):
    """DeflateFilesContainer.optimise hands a list of streams over (DeflateFilesContainer.java:18-43): ball.png's 20
    zlib streams, parsed one by one by the PNG reader, must reach the candidate engine as ONE launch; the same for the
    files of `optimise-folder`; output identical to the reference's goldens either way."""
    import shutil
    from deft4j_b200 import _native
    from deft4j_b200.container import getContainerForBytes
    from deft4j_b200.__main__ import main
    L = _native.lib()
    data = read_golden("apng/ball.png")
    cont = getContainerForBytes(data, "ball.png", gpu.DeflateStream)
    assert cont.read(data) and len(cont.getDeflateStreams()) >= 20
    n0 = L.deft4cu_debug_engine_launches()
    cont.optimise(True, None)
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    assert cont.write() == read_golden("apng/ball-opt.png")
    # a folder with several kinds of files: one launch for all of them, every file as the reference writes it
    pairs = [("apng/ball.png", "apng/ball-opt.png"), ("text.png", "text-opt.png"), ("lz-twice-twice.txt.gz", "lz-twice-twice-opt.txt.gz"),
             ("asyoulik/asyoulik-gzip.txt.gz", "asyoulik/asyoulik-gzip-opt.txt.gz"), ("284-edge-case/284.png", "284-edge-case/284-opt.png")]
    for inp, _ in pairs:
        shutil.copyfile(os.path.join(os.path.dirname(__file__), "golden", inp), tmp_path / os.path.basename(inp))
    n0 = L.deft4cu_debug_engine_launches()
    assert main(["optimise-folder", str(tmp_path)]) == 0
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    for inp, gold in pairs:
        assert (tmp_path / os.path.basename(inp)).read_bytes() == read_golden(gold), inp
    out = capsys.readouterr().out
    assert "167 bits saved in stream 0" in out and out.count("Optimising file") == len(pairs)


def test_batches_beyond_the_decoded_limit_are_split(gpu, oracle, monkeypatch):
    """Decoded offsets are 32-bit, so one device batch holds less than 4 GiB of decoded data; a larger list is split
    (streams never interact).  The limit is lowered to force the split on a small list."""
    streams = W.c4_streams(9, seed=21)
    ref = gpu.optimise_batch(streams, True)
    monkeypatch.setenv("D4_MAX_DECODED", "600000")   # every stream fits (<= 256 KiB decoded), the list of nine does not
    res = gpu.optimise_batch(streams, True)
    hs = gpu.DeflateStream.parse_batch(streams[:4])
    monkeypatch.delenv("D4_MAX_DECODED")
    for a, b in zip(res, ref):
        assert (a["status"], a["saved_bits"], a["out"], a["crc32"]) == (b["status"], b["saved_bits"], b["out"], b["crc32"])
    assert any(r["status"] == 0 for r in res)


def test_png_files_read_as_one_device_batch(gpu, oracle):
    """read_containers: the IDAT streams of many PNG files (Pillow-written) are parsed by ONE batch
    parse and optimised by ONE engine launch; every rewritten file equals what the oracle's stream class produces
    through the same PNGFile mirror, and the fixtures' goldens still come out."""
    from deft4j_b200 import _native
    from deft4j_b200.container import getContainerForBytes, read_containers, optimise_containers
    L = _native.lib()
    files = W.c3_png_files(12, first=400) + [read_golden("apng/ball.png"), read_golden("text.png"), b"not a container"]
    names = ["f%d.png" % i for i in range(len(files) - 1)] + ["junk.bin"]
    n0 = L.deft4cu_debug_engine_launches()
    conts = read_containers(files, names, gpu.DeflateStream)
    assert conts[-1] is None and all(c is not None for c in conts[:-1])
    saved = optimise_containers(conts[:-1], True)
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    assert conts[12].write() == read_golden("apng/ball-opt.png")
    assert conts[13].write() == read_golden("text-opt.png")
    for k in range(12):
        ref = getContainerForBytes(files[k], names[k], oracle.OracleDeflateStream)
        assert ref.read(files[k])
        assert ref.optimise(True, None) == saved[k]
        assert conts[k].write() == ref.write(), k


def test_native_png_front_end_on_the_device(gpu, oracle):
    """deft4cu_png_optimise_batch (csrc/png_front.cpp over the device batch entry): ONE engine launch for a list of
    files, the reference's golden outputs for its PNG fixtures, and for Pillow files and structurally mutated files
    exactly what the same front-end produces over the oracle (tests/front_oracle_shim.cpp)."""
    import hosttest_lib
    from test_png_front import mutants
    from deft4j_b200 import _native
    from deft4j_b200.container import optimise_png_files
    L = _native.lib()
    pairs = [("apng/ball.png", "apng/ball-opt.png"), ("text.png", "text-opt.png"), ("284-edge-case/284.png", "284-edge-case/284-opt.png")]
    files = [read_golden(a) for a, _ in pairs] + W.c3_png_files(40, first=700) + [b"junk"]
    n0 = L.deft4cu_debug_engine_launches()
    res = optimise_png_files(files, True)
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    for (a, g), r in zip(pairs, res):
        assert r["status"] == 0 and r["out"] == read_golden(g), a
    assert res[-1]["status"] == 1
    ref = optimise_png_files(files[3:15], True, lib=hosttest_lib.front_oracle_lib())
    assert res[3:15] == ref
    muts = mutants(per_base=12, seed=5)
    assert optimise_png_files(muts, True) == optimise_png_files(muts, True, lib=hosttest_lib.front_oracle_lib())
    assert optimise_png_files([], True) == []


def test_cli_folder_of_pngs_uses_the_native_front_end(gpu, oracle, tmp_path, capsys):
    """`optimise-folder` on a folder that holds only PNG files: one engine launch, the lines CMDUtil.optimiseFile prints,
    and the files the Python mirror of PNGFile writes with the oracle's streams."""
    from deft4j_b200 import _native
    from deft4j_b200.container import PNGFile
    from deft4j_b200.__main__ import main
    L = _native.lib()
    files = {"a%d.png" % i: d for i, d in enumerate(W.c3_png_files(5, first=900))}
    files["ball.png"] = read_golden("apng/ball.png")
    files["broken.png"] = read_golden("text.png")[:300]
    for name, d in files.items():
        (tmp_path / name).write_bytes(d)
    n0 = L.deft4cu_debug_engine_launches()
    assert main(["optimise-folder", str(tmp_path)]) == 1      # broken.png fails, the rest is optimised
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    cap = capsys.readouterr()
    assert cap.out.count("File type recognised as PNG") == 6 and "Invalid file" in cap.err
    assert "bits saved in stream 1 (fdAT chunk 1)" in cap.out
    assert (tmp_path / "ball.png").read_bytes() == read_golden("apng/ball-opt.png")
    assert (tmp_path / "broken.png").read_bytes() == files["broken.png"]
    for i in range(5):
        c = PNGFile(oracle.OracleDeflateStream)
        assert c.read(files["a%d.png" % i])
        saved = c.optimise(True, None)
        assert (tmp_path / ("a%d.png" % i)).read_bytes() == c.write()
        assert ("Saved %d bits with optimisation" % saved) in cap.out


def test_zip_archive_entries_are_one_device_batch(gpu, oracle):
    """A ZIP archive with many method-8 entries (workloads.c4_zip_archive) read through read_containers: every entry's
    stream in ONE batch parse and ONE engine launch; the rewritten archive equals the mirror's output over the oracle."""
    import io
    import zipfile
    from deft4j_b200 import _native
    from deft4j_b200.container import getContainerForBytes, read_containers, optimise_containers
    L = _native.lib()
    arch = W.c4_zip_archive(24, seed=31)
    n0 = L.deft4cu_debug_engine_launches()
    conts = read_containers([arch], ["a.zip"], gpu.DeflateStream)
    saved = optimise_containers(conts, True)[0]
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    out = conts[0].write()
    ref = getContainerForBytes(arch, "a.zip", oracle.OracleDeflateStream)
    assert ref.read(arch) and ref.optimise(True, None) == saved
    assert out == ref.write()
    zo = zipfile.ZipFile(io.BytesIO(out))
    assert zo.testzip() is None and len(zo.infolist()) == 24


def test_native_zip_front_end_on_the_device(gpu, oracle, tmp_path, capsys):
    """deft4cu_zip_optimise_batch (csrc/zip_front.cpp over the device batch entry): one engine launch for a list of
    archives, identical to the same front-end over the oracle, and `optimise-folder` on a folder of archives."""
    import hosttest_lib
    from test_zip_front import archives
    from deft4j_b200 import _native
    from deft4j_b200.container import optimise_zip_files
    from deft4j_b200.__main__ import main
    L = _native.lib()
    files = archives() + [W.c4_zip_archive(40, seed=77)]
    n0 = L.deft4cu_debug_engine_launches()
    res = optimise_zip_files(files, True)
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    assert res == optimise_zip_files(files, True, lib=hosttest_lib.front_oracle_lib())
    good = [k for k, r in enumerate(res) if r["status"] == 0]
    for k in good[:3]:
        (tmp_path / ("z%d.zip" % k)).write_bytes(files[k])
    n0 = L.deft4cu_debug_engine_launches()
    assert main(["optimise-folder", str(tmp_path)]) == 0
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    assert capsys.readouterr().out.count("File type recognised as Zip") == 3
    for k in good[:3]:
        assert (tmp_path / ("z%d.zip" % k)).read_bytes() == res[k]["out"]


def test_native_gz_and_zlib_front_ends_on_the_device(gpu, oracle, tmp_path, capsys):
    """deft4cu_gz_optimise_batch / deft4cu_zlib_optimise_batch over the device batch entry: the reference's gzip goldens, the
    same results as the front-ends over the oracle, one engine launch per list, and `optimise-folder` on a folder of .gz."""
    import hosttest_lib
    from test_gz_front import gz_files, zlib_files
    from deft4j_b200 import _native
    from deft4j_b200.container import optimise_gz_files, optimise_zlib_files
    from deft4j_b200.__main__ import main
    L = _native.lib()
    ref = hosttest_lib.front_oracle_lib()
    pairs = [("lz-twice-twice.txt.gz", "lz-twice-twice-opt.txt.gz"), ("asyoulik/asyoulik-gzip.txt.gz", "asyoulik/asyoulik-gzip-opt.txt.gz"),
             ("deflate-store-2.txt.gz", "deflate-store-2-opt.txt.gz")]
    files = [read_golden(a) for a, _ in pairs] + gz_files()
    n0 = L.deft4cu_debug_engine_launches()
    res = optimise_gz_files(files, True)
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    for (a, g), r in zip(pairs, res):
        assert r["status"] == 0 and r["out"] == read_golden(g), a
    assert res[3:] == optimise_gz_files(files[3:], True, lib=ref)
    zf = zlib_files()
    assert optimise_zlib_files(zf, True) == optimise_zlib_files(zf, True, lib=ref)
    for a, _ in pairs:
        (tmp_path / os.path.basename(a)).write_bytes(read_golden(a))
    n0 = L.deft4cu_debug_engine_launches()
    assert main(["optimise-folder", str(tmp_path)]) == 0
    assert L.deft4cu_debug_engine_launches() - n0 == 1
    assert capsys.readouterr().out.count("File type recognised as GZip") == 3
    for a, g in pairs:
        assert (tmp_path / os.path.basename(a)).read_bytes() == read_golden(g)
