"""The native ZIP front-end (deft4j_b200/csrc/zip_front.cpp, `deft4cu_zip_optimise_batch`) on the CPU: the shipped archive
code linked over the oracle (tests/front_oracle_shim.cpp), against the Python mirror of ZipFile (both restate the
un-vendored lljzip reader: parity with the reference itself is unpinned)."""
import io
import random
import struct
import zipfile
import zlib

import pytest

import workloads as W


@pytest.fixture(scope="module")
def front():
    import hosttest_lib
    return hosttest_lib.front_oracle_lib()


def _zip(entries, comment=b"", descriptor=False):
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as z:
        z.comment = comment
        for name, data, method, level in entries:
            zi = zipfile.ZipInfo(name, date_time=(2024, 1, 2, 3, 4, 6))
            zi.compress_type = method
            if descriptor:
                with z.open(zi, "w") as f:     # streamed write: sizes go into a data descriptor, flag bit 3 set
                    f.write(data)
            else:
                z.writestr(zi, data, compresslevel=level)
    return buf.getvalue()


def archives():
    rnd = random.Random(7)
    text = W.c2_text(40000, seed=3)
    ent = [("a.txt", text[:9000], zipfile.ZIP_DEFLATED, 6), ("dir/b.bin", bytes(rnd.randrange(256) for _ in range(700)), zipfile.ZIP_STORED, None),
           ("c.txt", text[9000:30000], zipfile.ZIP_DEFLATED, 9), ("empty", b"", zipfile.ZIP_DEFLATED, 6), ("d.txt", text[100:900], zipfile.ZIP_DEFLATED, 1)]
    out = [_zip(ent), _zip(ent, comment=b"a comment with PK\x05\x06 inside"), _zip(ent[:2], descriptor=True), _zip([ent[1]]),
           W.c4_zip_archive(7, seed=2)]
    base = out[0]
    # structural damage the reader must treat like the mirror does
    out.append(base[:-30])                                     # end record cut off
    out.append(base[:len(base) // 2])                          # no end record at all
    out.append(b"PK\x05\x06" + b"\0" * 18)                     # empty archive: no local headers
    cd = base.rfind(b"PK\x01\x02")
    broken = bytearray(base)
    broken[struct.unpack_from("<I", base, cd + 42)[0]] ^= 0xFF  # last entry's local signature destroyed -> write raises
    out.append(bytes(broken))
    bad = bytearray(base)
    off0 = 30 + len("a.txt")
    bad[off0] |= 0x06                                          # first entry's deflate stream: block type 3
    out.append(bytes(bad))
    dup = bytearray(base)                                      # two central entries pointing at the same local header
    first = base.find(b"PK\x01\x02")
    second = base.find(b"PK\x01\x02", first + 4)
    dup[second + 42:second + 46] = base[first + 42:first + 46]
    out.append(bytes(dup))
    out.append(b"not a zip at all")
    return out


def mirror_outcome(data, stream_cls):
    from deft4j_b200.container import ZipFile
    c = ZipFile(stream_cls)
    if not c.read(data):
        return 1, None, 0, []
    saved = c.optimise(True, None)
    names = [s.getName() for s in c.getDeflateStreams()]
    try:
        return 0, c.write(), saved, names
    except IOError:
        return 2, None, saved, names


def test_native_zip_front_end_follows_the_mirror(front, oracle):
    from deft4j_b200.container import optimise_zip_files
    files = archives()
    res = optimise_zip_files(files, True, lib=front)
    seen = {}
    for k, (data, r) in enumerate(zip(files, res)):
        st, out, saved, names = mirror_outcome(data, oracle.OracleDeflateStream)
        assert r["status"] == st, k
        seen[st] = seen.get(st, 0) + 1
        if st == 0:
            assert r["out"] == out and r["saved_bits"] == saved and [n for n, _ in r["streams"]] == names, k
            with zipfile.ZipFile(io.BytesIO(r["out"])) as zo, zipfile.ZipFile(io.BytesIO(data)) as zi:
                if k != 10:   # (the archive with two directory entries for one file is not a valid ZIP to begin with)
                    assert zo.testzip() is None
                    assert [zo.read(i.filename) for i in zo.infolist()] == [zi.read(i.filename) for i in zi.infolist()]
    assert seen.get(0, 0) >= 5 and seen.get(1, 0) >= 4 and seen.get(2, 0) >= 1
    assert optimise_zip_files([], True, lib=front) == []
