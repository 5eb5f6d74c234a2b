"""Synthetic workloads of BASELINE.json's configs, shared by bench.py and the tests.

Everything is generated from fixed seeds with numpy + the system zlib (the compressor that *produces* the
input streams is not part of the measured path).

  C2  single raw deflate stream, zlib level 6 (dynamic blocks) over Zipf-distributed synthetic text
  C3  batch of 256x256 RGBA PNG IDAT streams (zlib level 6 over adaptively filtered synthetic images)
  C4  stream mix of stored / fixed / dynamic entries (the payloads of the ZIP config)
  C5  adversarial streams: len-258 / dist-1 and dist-32768 matches, RLE-heavy headers
"""
import zlib

import numpy as np

VOCAB = 20000


def _vocab(seed=0xDEF7):
    rng = np.random.default_rng(seed)
    lens = rng.integers(2, 10, VOCAB)
    table = rng.integers(97, 123, (VOCAB, 10), dtype=np.uint8)
    table[np.arange(VOCAB), lens] = 32  # the separating space
    p = 1.0 / np.arange(1, VOCAB + 1)
    return table, (lens + 1).astype(np.int64), p / p.sum(), rng


def text_chunks(seed=0xDEF7, words_per_chunk=1 << 18):
    """Endless generator of text chunks (bytes): 20 000-word vocabulary, Zipf(1.0) draws, single spaces."""
    table, wl, p, rng = _vocab(seed)
    cdf = np.cumsum(p)
    while True:
        idx = np.searchsorted(cdf, rng.random(words_per_chunk)).clip(0, VOCAB - 1)
        l = wl[idx]
        ends = np.cumsum(l)
        total = int(ends[-1])
        starts = ends - l
        rows = np.repeat(idx, l)
        cols = np.arange(total, dtype=np.int64) - np.repeat(starts, l)
        yield table[rows, cols].tobytes()


def c2_stream(target_bytes, seed=0xDEF7):
    """ONE raw deflate stream (zlib.compressobj(6, DEFLATED, -15, 8)) of at least target_bytes.  Generating 1 GiB
    takes about two minutes of single-thread zlib, so large streams are cached under the temp directory (the
    generator is deterministic: the cache only saves time when bench.py runs more than once on a box)."""
    import os
    import tempfile
    cache = None
    if target_bytes >= (64 << 20):
        cache = os.path.join(tempfile.gettempdir(), "deft4cu_c2_%d_%x.bin" % (target_bytes, seed))
        try:
            with open(cache, "rb") as f:
                data = f.read()
            if len(data) >= target_bytes:
                return data
        except OSError:
            pass
    data = _c2_stream(target_bytes, seed)
    if cache:
        try:
            tmp = "%s.%d.tmp" % (cache, os.getpid())
            with open(tmp, "wb") as f:
                f.write(data)
            os.replace(tmp, cache)
        except OSError:
            pass
    return data


def _c2_stream(target_bytes, seed):
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    parts, n = [], 0
    # ~2.6 bytes of text per compressed byte; feed text until the compressor has emitted enough
    for chunk in text_chunks(seed):
        out = co.compress(chunk)
        parts.append(out)
        n += len(out)
        if n >= target_bytes:
            break
    parts.append(co.flush())
    return b"".join(parts)


def c2_text(nbytes, seed=0xDEF7):
    parts, n = [], 0
    for chunk in text_chunks(seed, 1 << 16):
        parts.append(chunk)
        n += len(chunk)
        if n >= nbytes:
            break
    return b"".join(parts)[:nbytes]


def _png_image(index, w=256, h=256):
    """Synthetic 256x256 RGBA image: a mix of flat regions, linear gradients and 5 % uniform noise."""
    rng = np.random.default_rng(index)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    # flat regions
    for _ in range(6):
        x0, y0 = rng.integers(0, w), rng.integers(0, h)
        x1, y1 = rng.integers(x0, w + 1), rng.integers(y0, h + 1)
        img[y0:y1, x0:x1] = rng.integers(0, 256, 4, dtype=np.uint8)
    # linear gradients
    for _ in range(3):
        y0 = rng.integers(0, h - 16)
        y1 = rng.integers(y0 + 8, h + 1)
        g = (np.arange(w) * rng.integers(1, 4) + rng.integers(0, 256)) & 255
        img[y0:y1, :, rng.integers(0, 3)] = g.astype(np.uint8)[None, :]
    img[..., 3] = 255
    # 5 % noise
    m = rng.random((h, w)) < 0.05
    img[m] = rng.integers(0, 256, (int(m.sum()), 4), dtype=np.uint8)
    return img


def c3_png_files(count, first=0):
    """`count` PNG files as SURVEY.md 8(d) specifies them: Pillow `save(compress_level=6)` (adaptive filter, zlib IDAT),
    seed = index."""
    import io
    from PIL import Image
    out = []
    for i in range(first, first + count):
        b = io.BytesIO()
        Image.fromarray(_png_image(i), "RGBA").save(b, format="PNG", compress_level=6)
        out.append(b.getvalue())
    return out


def _idat_payload(png):
    pos, z = 8, b""
    while pos + 12 <= len(png):
        n = int.from_bytes(png[pos:pos + 4], "big")
        if png[pos + 4:pos + 8] == b"IDAT":
            z += png[pos + 8:pos + 8 + n]
        pos += 12 + n
    return z


def c3_streams(count, first=0):
    """Raw deflate payloads of the IDAT streams of c3_png_files (the zlib wrapper is the container's business)."""
    return [_idat_payload(p)[2:-4] for p in c3_png_files(count, first)]


def c4_streams(count, seed=4):
    """Method-8 entry payloads: k%3==0 stored (level 0), ==1 Z_FIXED level 6, ==2 default level 6; payload sizes
    log-uniform 1 KiB..256 KiB of C2 text."""
    rng = np.random.default_rng(seed)
    text = c2_text(4 << 20, seed=seed)
    out = []
    for k in range(count):
        n = int(np.exp(rng.uniform(np.log(1024), np.log(256 * 1024))))
        off = int(rng.integers(0, len(text) - n))
        payload = text[off:off + n]
        if k % 3 == 0:
            co = zlib.compressobj(0, zlib.DEFLATED, -15, 8)
        elif k % 3 == 1:
            co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
        else:
            co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
        out.append(co.compress(payload) + co.flush())
    return out


def c4_zip_archive(count, seed=4):
    """One ZIP archive whose `count` method-8 entries carry the payloads of c4_streams: local headers,
    central directory and end record written by hand so that the entry data is exactly those raw deflate streams."""
    import struct
    streams = c4_streams(count, seed)
    out = bytearray()
    central = bytearray()
    for k, raw in enumerate(streams):
        data = zlib.decompress(raw, -15)
        name = ("dir%d/entry%05d.txt" % (k % 7, k)).encode()
        crc = zlib.crc32(data) & 0xffffffff
        off = len(out)
        # version needed 20, flags 0, method 8, time / date, crc, sizes, name length, extra length
        out += struct.pack("<IHHHHHIIIHH", 0x04034b50, 20, 0, 8, 0x6000, 0x5821, crc, len(raw), len(data), len(name), 0) + name + raw
        central += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014b50, 20, 20, 0, 8, 0x6000, 0x5821, crc, len(raw), len(data),
                               len(name), 0, 0, 0, 0, 0, off) + name
    cd_off = len(out)
    out += central
    out += struct.pack("<IHHHHIIH", 0x06054b50, 0, 0, count, count, len(central), cd_off, 0)
    return bytes(out)


def c5_streams(scale=1):
    """Adversarial streams (scaled: `scale` MiB per long-match stream)."""
    rng = np.random.default_rng(5)
    n = scale << 20
    period = rng.integers(0, 256, 32768, dtype=np.uint8).tobytes()
    sparse = rng.choice(np.array([0, 255], dtype=np.uint8), 1 << 16).tobytes()
    two = rng.choice(np.array([7, 9], dtype=np.uint8), 1 << 15, p=[0.9, 0.1]).tobytes()
    datas = [b"\x41" * n, (period * (n // 32768 + 1))[:n], sparse, two, bytes(range(256)) * 64]
    out = []
    for d in datas:
        for level, strat in ((6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_RLE)):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strat)
            out.append(co.compress(d) + co.flush())
    return out


# ---- hand-made streams (fixed-code blocks) for shapes zlib never emits -----------------------------------------
_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EB = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
              6145, 8193, 12289, 16385, 24577]
_DIST_EB = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


class BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def bits(self, value, count):  # LSB first
        self.acc |= value << self.n
        self.n += count
        while self.n >= 8:
            self.out.append(self.acc & 255)
            self.acc >>= 8
            self.n -= 8

    def code(self, code, length):  # Huffman codes go MSB first
        self.bits(int(format(code, "0%db" % length)[::-1], 2), length)

    def align(self):
        if self.n:
            self.bits(0, 8 - self.n)

    def done(self):
        self.align()
        return bytes(self.out)


def _fixed_litlen(bw, sym):
    if sym <= 143: bw.code(0x30 + sym, 8)
    elif sym <= 255: bw.code(0x190 + sym - 144, 9)
    elif sym <= 279: bw.code(sym - 256, 7)
    else: bw.code(0xC0 + sym - 280, 8)


def fixed_block(bw, symbols, final):
    """symbols: int literal | (length, distance) | (258, distance, 'edge') for the symbol-284+31 spelling of 258."""
    bw.bits(1 if final else 0, 1)
    bw.bits(1, 2)
    for s in symbols:
        if isinstance(s, int):
            _fixed_litlen(bw, s)
            continue
        length, dist = s[0], s[1]
        if len(s) > 2:
            _fixed_litlen(bw, 284); bw.bits(31, 5)
        else:
            i = max(k for k in range(29) if _LEN_BASE[k] <= length and (k < 28 or length == 258))
            if length == 258: i = 28
            _fixed_litlen(bw, 257 + i); bw.bits(length - _LEN_BASE[i], _LEN_EB[i])
        d = max(k for k in range(30) if _DIST_BASE[k] <= dist)
        bw.code(d, 5); bw.bits(dist - _DIST_BASE[d], _DIST_EB[d])
    _fixed_litlen(bw, 256)


def stored_block(bw, data, final):
    bw.bits(1 if final else 0, 1)
    bw.bits(0, 2)
    bw.align()
    bw.bits(len(data), 16); bw.bits(len(data) ^ 0xffff, 16)
    bw.out += data


def _canon_codes(lens):
    """Huffman.buildCodes (base/huffman/Huffman.java:35-64): canonical codes, no validity check — an over-subscribed
    length set yields codes that do not fit their length (they can never be matched by the first-match decoder)."""
    count = [0] * 17
    for l in lens:
        count[l] += 1
    count[0] = 0
    code, nxt = 0, [0] * 17
    for l in range(1, 17):
        code = (code + count[l - 1]) << 1
        nxt[l] = code
    out = []
    for l in lens:
        out.append(nxt[l] if l else 0)
        if l:
            nxt[l] += 1
    return out


def dynamic_block(bw, litlen_lens, dist_lens, symbols, final):
    """A dynamic block with an explicit (possibly incomplete or over-subscribed) pair of code-length sets.  The
    header spells every length plainly (no 16/17/18 runs) with a flat 4-bit code over lengths 0..15."""
    bw.bits(1 if final else 0, 1)
    bw.bits(2, 2)
    nl, nd = max(257, len(litlen_lens)), max(1, len(dist_lens))
    L = list(litlen_lens) + [0] * (nl - len(litlen_lens))
    D = list(dist_lens) + [0] * (nd - len(dist_lens))
    bw.bits(nl - 257, 5); bw.bits(nd - 1, 5); bw.bits(19 - 4, 4)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    for s in order:
        bw.bits(4 if s < 16 else 0, 3)
    for l in L + D:
        bw.code(l, 4)
    lc, dc = _canon_codes(L), _canon_codes(D)
    for s in symbols + [256]:
        if isinstance(s, int):
            bw.code(lc[s], L[s])
            continue
        length, dist = s
        i = 28 if length == 258 else max(k for k in range(28) if _LEN_BASE[k] <= length)
        bw.code(lc[257 + i], L[257 + i]); bw.bits(length - _LEN_BASE[i], _LEN_EB[i])
        d = max(k for k in range(30) if _DIST_BASE[k] <= dist)
        bw.code(dc[d], D[d]); bw.bits(dist - _DIST_BASE[d], _DIST_EB[d])


def odd_code_streams():
    """Code-length sets zlib rejects but the reference's first-match decoder accepts: an incomplete
    litlen code, an over-subscribed litlen code (the surplus code never matches), an over-subscribed distance code."""
    out = {}
    L = [0] * 257
    L[65], L[66], L[256] = 1, 3, 2                      # Kraft 1/2 + 1/8 + 1/4 < 1
    bw = BitWriter(); dynamic_block(bw, L, [0], [65, 66, 65, 65, 66] * 6, True); out["incomplete_litlen"] = bw.done()
    L = [0] * 258
    L[256], L[65], L[66], L[67], L[257] = 1, 2, 2, 2, 3  # Kraft 1/2 + 3/4 + 1/8 > 1: 67 and 257 get codes that never match
    bw = BitWriter(); dynamic_block(bw, L, [0], [65, 66, 66, 65] * 9, True); out["oversubscribed_litlen"] = bw.done()
    L = [0] * 258
    L[65], L[66], L[256], L[257] = 2, 2, 2, 2
    D = [1, 1, 1]                                        # three 1-bit distance codes: the third never matches
    bw = BitWriter(); dynamic_block(bw, L, D, [65, 66, 65, (3, 1), (3, 2), 66, (3, 1)] * 5, True); out["oversubscribed_dist"] = bw.done()
    bw = BitWriter()
    fixed_block(bw, [72, 105] * 30, False); dynamic_block(bw, L, D, [65, (3, 2), 66] * 8, False); fixed_block(bw, [33] * 9, True)
    out["odd_code_midstream"] = bw.done()
    return out


def handmade_streams():
    """Shapes outside zlib's repertoire: distance 32768 / length 258, the 284+31 edge case, empty blocks mid-stream
    (SURVEY.md H6), stored blocks around 65535 (H5/H12), a lone EOB."""
    rng = np.random.default_rng(55)
    out = {}
    base = [int(x) for x in rng.integers(0, 256, 32768)]
    bw = BitWriter(); fixed_block(bw, base + [(258, 32768)] * 130 + [(3, 1), (258, 1)], True); out["dist32768_len258"] = bw.done()
    bw = BitWriter(); fixed_block(bw, [65, 66, 67] + [(258, 3, "edge")] * 40 + [(258, 3)] * 3, True); out["edge284"] = bw.done()
    bw = BitWriter()
    fixed_block(bw, [72, 105, 32] * 20, False); stored_block(bw, b"", False); fixed_block(bw, [(30, 60), 33], False)
    fixed_block(bw, [], False); fixed_block(bw, [10] * 50, True); out["empty_blocks_midstream"] = bw.done()
    bw = BitWriter(); stored_block(bw, bytes(base[:3000]) * 10, False); stored_block(bw, bytes(base) + bytes(base[:2766]), False)
    stored_block(bw, b"xyz" * 11000, False); stored_block(bw, b"tail", True); out["stored_around_65535"] = bw.done()
    bw = BitWriter(); fixed_block(bw, [], True); out["lone_eob"] = bw.done()
    bw = BitWriter(); stored_block(bw, b"", True); out["empty_stored_only"] = bw.done()
    bw = BitWriter(); fixed_block(bw, [0] + [(258, 1)] * 300, False); fixed_block(bw, [(258, 1)] * 300 + [255], True); out["rle_two_fixed"] = bw.done()
    # an empty FIXED block (what Z_PARTIAL_FLUSH emits) first in the list of empties, between two blocks the merge
    # phase then joins: its EOB sits between their symbols in any pooled layout
    words = [int(x) for x in rng.integers(97, 110, 400)]
    bw = BitWriter(); fixed_block(bw, words[:200] + [(20, 150)] * 4, False); fixed_block(bw, [], False)
    fixed_block(bw, words[200:] + [(9, 33)] * 6, True); out["fixed_emptyfixed_fixed"] = bw.done()
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    t = c2_text(24_000, seed=77)
    out["dyn_partialflush_dyn"] = co.compress(t[:12_000]) + co.flush(zlib.Z_PARTIAL_FLUSH) + co.compress(t[12_000:]) + co.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    out["dyn_partialflush_x3"] = (co.compress(t[:6_000]) + co.flush(zlib.Z_PARTIAL_FLUSH) + co.compress(t[6_000:12_000]) +
                                   co.flush(zlib.Z_PARTIAL_FLUSH) + co.compress(t[12_000:]) + co.flush())
    return out
