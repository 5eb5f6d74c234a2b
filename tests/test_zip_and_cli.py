"""SURVEY.md §8f rows 3-4 on CPU: the ZIP container mirror and the CLI mirror, driven with the oracle as the stream
engine (the CUDA engine is covered by test_gpu_parity.py; containers and CLI are host-side code either way)."""
import io
import os
import zipfile
import zlib

import pytest

import workloads as W
from conftest import GOLDEN_PAIRS, golden_path, read_golden


def _make_zip(entries, comment=b""):
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as z:
        z.comment = comment
        for name, data, method, level in entries:
            zi = zipfile.ZipInfo(name, date_time=(2024, 1, 2, 3, 4, 6))
            zi.compress_type = method
            z.writestr(zi, data, compresslevel=level)
    return buf.getvalue()


def _entries(z):
    with zipfile.ZipFile(io.BytesIO(z)) as f:
        assert f.testzip() is None
        return {i.filename: (f.read(i.filename), i.compress_type, i.compress_size, i.CRC, i.date_time) for i in f.infolist()}, f.comment


def test_zip_container_optimises_deflated_entries(oracle):
    from deft4j_b200.container import ZipFile, getContainerForBytes
    text = W.c2_text(40_000, seed=3)
    z = _make_zip([("a/text.txt", text, zipfile.ZIP_DEFLATED, 6), ("stored.bin", b"\x00\x01" * 300, zipfile.ZIP_STORED, None),
                   ("fast.txt", text[:9000], zipfile.ZIP_DEFLATED, 1), ("empty", b"", zipfile.ZIP_DEFLATED, 6)], b"archive comment")
    cont = getContainerForBytes(z, "x.zip", oracle.OracleDeflateStream)
    assert isinstance(cont, ZipFile) and cont.fileType() == "Zip"
    assert cont.read(z)
    assert len(cont.getDeflateStreams()) == 3            # the stored entry is carried through (ZipFile.java:97-99)
    log = io.StringIO()
    saved = cont.optimise(True, log)
    out = cont.write()
    before, c0 = _entries(z)
    after, c1 = _entries(out)
    assert c0 == c1 == b"archive comment"
    assert list(before) == list(after)
    for name in before:
        assert before[name][0] == after[name][0] and before[name][1] == after[name][1]
        assert before[name][3:] == after[name][3:]
        assert after[name][2] <= before[name][2]
    assert saved > 0 and len(out) < len(z) and "Total bits saved %d" % saved in log.getvalue()
    # every deflated payload is exactly the oracle's optimised stream of the original payload
    with zipfile.ZipFile(io.BytesIO(z)) as f0, zipfile.ZipFile(io.BytesIO(out)) as f1:
        for i0, i1 in zip(f0.infolist(), f1.infolist()):
            if i0.compress_type != zipfile.ZIP_DEFLATED:
                continue
            raw0 = z[i0.header_offset + 30 + len(i0.filename.encode()) + len(i0.extra):][:i0.compress_size]
            raw1 = out[i1.header_offset + 30 + len(i1.filename.encode()) + len(i1.extra):][:i1.compress_size]
            s = oracle.OracleDeflateStream()
            assert s.parse(raw0)
            s.optimise(True)
            assert s.asBytes() == raw1
    # idempotent layout: a second pass over the output changes nothing but may not grow it
    c2 = getContainerForBytes(out, "x.zip", oracle.OracleDeflateStream)
    assert c2.read(out)
    c2.optimise(True, None)
    assert len(c2.write()) <= len(out)


@pytest.mark.parametrize("name,sibling", [("deflate-dynamic.txt.zip", "deflate-dynamic.txt.gz"), ("deflate-store.txt.zip", "deflate-store.txt.gz")])
def test_reference_zip_fixtures(oracle, name, sibling):
    """The reference's two ZIP inputs (no golden output exists: parity unpinned).  Their single entry holds the same
    deflate stream as the .gz sibling, so the optimised payload must equal the optimised gzip member's stream."""
    from deft4j_b200.container import getContainerForBytes
    z = read_golden(name)
    cont = getContainerForBytes(z, name, oracle.OracleDeflateStream)
    assert cont.read(z)
    cont.optimise(True, None)
    out = cont.write()
    (data, *_), = _entries(out)[0].values()
    g = getContainerForBytes(read_golden(sibling), sibling, oracle.OracleDeflateStream)
    assert g.read(read_golden(sibling))
    g.optimise(True, None)
    assert zlib.decompress(g.getDeflateStreams()[0].asBytes(), -15) == data
    assert cont.getDeflateStreams()[0].asBytes() == g.getDeflateStreams()[0].asBytes()


def test_zip_refuses_what_the_reference_refuses(oracle):
    from deft4j_b200.container import ZipFile
    assert ZipFile(oracle.OracleDeflateStream).read(b"PK\x03\x04" + b"\0" * 40) is False
    assert ZipFile(oracle.OracleDeflateStream).read(b"") is False


def test_cli_optimise_matches_reference_goldens(oracle, tmp_path, capsys):
    """`deft4j optimise -m NONE` on three of the reference's fixtures with runTestOpt.sh's flags: same output bytes,
    same stdout lines (CMDUtil.java:64-72, DeflateFilesContainer.java:31-39)."""
    from deft4j_b200.__main__ import main
    for inp, gold, merge in [GOLDEN_PAIRS[0], GOLDEN_PAIRS[1], GOLDEN_PAIRS[2]]:
        out = tmp_path / ("out-" + os.path.basename(inp))
        argv = ["optimise", golden_path(inp), str(out)] + ([] if merge else ["--no-merge-blocks"])
        assert main(argv, stream_cls=oracle.OracleDeflateStream) == 0
        assert out.read_bytes() == read_golden(gold)
        printed = capsys.readouterr().out.splitlines()
        ref = [l for l in read_golden(gold + ".txt").decode().splitlines() if l.strip()]
        assert printed == ref, (printed, ref)


def test_cli_errors_and_overwrite(oracle, tmp_path, capsys):
    from deft4j_b200.__main__ import main
    cls = oracle.OracleDeflateStream
    assert main(["optimise", str(tmp_path / "missing.gz"), str(tmp_path / "o.gz")], stream_cls=cls) == 1
    assert "Error: Input file does not exist" in capsys.readouterr().err
    bad = tmp_path / "bad.gz"
    bad.write_bytes(b"\x1f\x8b\x08\x00" + b"\xff" * 30)
    assert main(["optimise", str(bad), str(tmp_path / "o.gz")], stream_cls=cls) == 1
    assert not (tmp_path / "o.gz").exists()
    # in-place overwrite goes through a temp file and only happens on success
    src = tmp_path / "lz.gz"
    src.write_bytes(read_golden("lz-twice-twice.txt.gz"))
    assert main(["optimise", str(src), str(src)], stream_cls=cls) == 0
    assert src.read_bytes() == read_golden("lz-twice-twice-opt.txt.gz")
    bad_before = bad.read_bytes()
    assert main(["optimise", str(bad), str(bad)], stream_cls=cls) == 1 and bad.read_bytes() == bad_before


def test_cli_optimise_folder(oracle, tmp_path, capsys):
    from deft4j_b200.__main__ import main
    d = tmp_path / "tree" / "sub"
    d.mkdir(parents=True)
    (d / "a.txt.gz").write_bytes(read_golden("lz-twice-twice.txt.gz"))
    (d / "notes.txt").write_bytes(b"plain text, not a container")       # RawDeflateFile by extension: skipped
    (tmp_path / "tree" / "t.png").write_bytes(read_golden("text.png"))
    assert main(["optimise-folder", str(tmp_path / "tree")], stream_cls=oracle.OracleDeflateStream) == 0
    assert (d / "a.txt.gz").read_bytes() == read_golden("lz-twice-twice-opt.txt.gz")
    assert (tmp_path / "tree" / "t.png").read_bytes() == read_golden("text-opt.png")
    assert (d / "notes.txt").read_bytes() == b"plain text, not a container"
    assert capsys.readouterr().out.count("Optimising file ") == 2


def test_cli_has_no_cpu_fallback(tmp_path):
    """Without a CUDA device the command line must fail loudly (no silent CPU path): the default stream engine is the
    CUDA one and its first call raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from deft4j_b200.__main__ import main
    from deft4j_b200._native import Deft4cuError
    with pytest.raises(Deft4cuError):
        main(["optimise", golden_path("lz-twice-twice.txt.gz"), str(tmp_path / "o.gz")])
    assert not (tmp_path / "o.gz").exists()


def test_read_containers_record_and_replay(oracle):
    """The two-pass container read (deft4j_b200.container.read_containers) with a batching stand-in built on the
    oracle: every stream goes through ONE parse_batch call and the rewritten files are the reference's goldens."""
    from deft4j_b200.container import read_containers, optimise_containers
    calls = []

    class Batching(oracle.OracleDeflateStream):
        @classmethod
        def parse_batch(cls, buffers, names=None):
            calls.append(len(buffers))
            out = []
            for i, b in enumerate(buffers):
                s = cls(names[i] if names else None)
                out.append(s if oracle.OracleDeflateStream.parse(s, b) else None)
            return out

    pairs = [(a, g) for a, g, fast in GOLDEN_PAIRS if fast and "zopfli" not in a]
    files = [read_golden(a) for a, _ in pairs] + [b"\xff\xff not deflate"]
    names = [os.path.basename(a) for a, _ in pairs] + ["junk.bin"]
    conts = read_containers(files, names, Batching)
    assert len(calls) == 1 and calls[0] >= len(pairs)
    assert conts[-1] is None
    optimise_containers(conts[:-1], True)
    for (a, gold), c in zip(pairs, conts):
        assert c.write() == read_golden(gold), a
