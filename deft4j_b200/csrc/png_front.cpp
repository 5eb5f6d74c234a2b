// png_front.cpp — native batch front-end for PNG / APNG files (host code).
//
// Replaces, for a LIST of files, what deft4j-container's PNGFile does per file
// (deft4j-container/src/main/java/com/github/NeRdTheNed/deft4j/container/PNGFile.java):
//   read      :574-605  (chunk reader :162-215, chunk state machine PNGChunkHelper :413-572)
//   optimise  : DeflateFilesContainer.java:18-43 over getDeflateStreams() (:376-389)
//   write     :391-411  (syncStreams :262-369: one IDAT chunk, one fdAT chunk per frame, renumbered sequence numbers,
//                         zTXt / iCCP / iTXt payloads put back behind their keyword prefix)
// The chunk work is format-driven host code and runs on host threads, one file per task; the zlib streams of ALL
// files go to the device as ONE list through deft4cu_optimise_batch (the only thing this file calls), so a folder of
// PNGs is one launch of every kernel.  A file is read in two steps because a container only learns whether a stream
// parses when the device has seen it: step 1 assumes every stream parses, and a file whose IDAT / fdAT stream did not
// is reported unreadable afterwards — which is what PNGFile.read returns for it (flush() fails, :468-491).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/deft4cu.h"
#include "front_util.h"

namespace {
using d4front::parallel_for;

// ---- CRC-32 (IEEE, reflected) of chunk type + data (PNGFile.java:140-158 uses java.util.zip.CRC32) -----------------
uint32_t g_tab[8][256];
std::once_flag g_tab_once;
void crc_tables() {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1)));
        g_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xff];
}
// state in, state out (the caller inverts at both ends)
uint32_t crc_bytes(uint32_t c, const uint8_t* p, size_t n) {
    while (n && ((uintptr_t)p & 7)) { c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xff]; n--; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = g_tab[7][w & 0xff] ^ g_tab[6][(w >> 8) & 0xff] ^ g_tab[5][(w >> 16) & 0xff] ^ g_tab[4][(w >> 24) & 0xff] ^
            g_tab[3][(w >> 32) & 0xff] ^ g_tab[2][(w >> 40) & 0xff] ^ g_tab[1][(w >> 48) & 0xff] ^ g_tab[0][w >> 56];
        p += 8; n -= 8;
    }
    while (n--) c = (c >> 8) ^ g_tab[0][(c ^ *p++) & 0xff];
    return c;
}
#if defined(__x86_64__)
// carry-less-multiply folding (Gopal et al., "Fast CRC computation for generic polynomials using PCLMULQDQ", Intel
// 2009): four 128-bit lanes folded 64 bytes at a time, then 128 -> 64 -> 32 bits by Barrett reduction.
// n is a multiple of 16 and at least 64.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc_fold(uint32_t c, const uint8_t* p, size_t n) {
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124ll);
    const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)c));
    p += 64; n -= 64;
    while (n >= 64) {
        __m128i a = _mm_clmulepi64_si128(x1, k1k2, 0x00), b = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        __m128i d = _mm_clmulepi64_si128(x3, k1k2, 0x00), e = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11); x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11); x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, a), _mm_loadu_si128((const __m128i*)(p + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, b), _mm_loadu_si128((const __m128i*)(p + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, d), _mm_loadu_si128((const __m128i*)(p + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, e), _mm_loadu_si128((const __m128i*)(p + 48)));
        p += 64; n -= 64;
    }
#define D4_FOLD1(acc, next)                                         \
    do {                                                            \
        const __m128i lo_ = _mm_clmulepi64_si128(acc, k3k4, 0x00);  \
        acc = _mm_clmulepi64_si128(acc, k3k4, 0x11);                \
        acc = _mm_xor_si128(_mm_xor_si128(acc, next), lo_);         \
    } while (0)
    D4_FOLD1(x1, x2); D4_FOLD1(x1, x3); D4_FOLD1(x1, x4);
    while (n >= 16) { const __m128i nx = _mm_loadu_si128((const __m128i*)p); D4_FOLD1(x1, nx); p += 16; n -= 16; }
#undef D4_FOLD1
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_clmulepi64_si128(_mm_and_si128(x1, mask32), k5, 0x00);
    x1 = _mm_xor_si128(x1, t);
    t = _mm_clmulepi64_si128(_mm_and_si128(x1, mask32), poly, 0x10);
    t = _mm_clmulepi64_si128(_mm_and_si128(t, mask32), poly, 0x00);
    x1 = _mm_xor_si128(x1, t);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
bool have_clmul() {
    static const bool v = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return v;
}
#endif
uint32_t crc_update(uint32_t c, const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    if (n >= 128 && have_clmul()) {
        const size_t m = n & ~(size_t)15;
        c = crc_fold(c, p, m);
        p += m; n -= m;
    }
#endif
    return crc_bytes(c, p, n);
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put_be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
constexpr uint32_t tag(char a, char b, char c, char d) { return ((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint8_t)d; }
constexpr uint32_t IDAT = tag('I', 'D', 'A', 'T'), IEND = tag('I', 'E', 'N', 'D'), zTXt = tag('z', 'T', 'X', 't'),
                   iCCP = tag('i', 'C', 'C', 'P'), iTXt = tag('i', 'T', 'X', 't'), acTL = tag('a', 'c', 'T', 'L'),
                   fcTL = tag('f', 'c', 'T', 'L'), fdAT = tag('f', 'd', 'A', 'T');
constexpr uint64_t INT_MAX_ = 2147483647ull;

// java.io.InputStream over the file: read() past the end is -1, which the readers mask to 255
struct Reader {
    const uint8_t* p; uint64_t n, pos;
    uint32_t rd() { return pos < n ? p[pos++] : 255u; }
    uint32_t rd32() { uint32_t a = rd(), b = rd(), c = rd(), d = rd(); return (a << 24) | (b << 16) | (c << 8) | d; }
};

struct Chunk {
    uint32_t type = 0;
    const uint8_t* data = nullptr;   // into the file (or into File::owned)
    uint32_t len = 0;
    const uint8_t* raw = nullptr;    // the chunk as it stands in the file (length .. CRC), when it can be copied verbatim
    uint32_t seq = 0;
    bool new_seq = false;            // setSeq: the first four data bytes are replaced by `seq` (:67-73)
    int gen = -1;                    // a chunk made by syncStreams: index into File::zs of the IDAT / fdAT stream it carries
    int non_id = -1;                 // zTXt / iCCP / iTXt whose payload is a stream of this file: index among those
    uint32_t keep = 0;               //   ... and the bytes of `data` that stay in front of the rewritten payload
    bool hasSeq() const { return type == fdAT || type == fcTL; }
    bool zlibNonIdat() const { return type == zTXt || type == iCCP || type == iTXt; }
};

// one zlib stream of a file (ZLibFile.java:14-109)
struct ZS {
    int role = 0;                    // 0 IDAT, 1 fdAT frame, 2 zTXt / iCCP / iTXt
    std::vector<uint8_t> joined;     // the chunk payloads of the stream, when it spans more than one chunk
    const uint8_t* z = nullptr;
    uint64_t zlen = 0;
    uint8_t cmf = 0, flg = 0;
    char name[24] = {0};
    int slot = -1;                   // position in the flat list handed to the device
    bool dropped = false;            // role 2 whose stream did not parse: the chunk stays as it is (:532-543)
    void add(const uint8_t* p, uint64_t n) {
        if (!n) return;
        if (!z && joined.empty()) { z = p; zlen = n; return; }
        if (joined.empty()) joined.assign(z, z + zlen);
        joined.insert(joined.end(), p, p + n);
        z = joined.data(); zlen = joined.size();
    }
    // ZLibFile.read up to the deflate stream (:59-80)
    bool header() {
        if (zlen < 2) return false;   // CMF or FLG is -1: fails the method / FCHECK test for every value of the other
        cmf = z[0]; flg = z[1];
        if ((cmf & 0xF) != 8) return false;
        if ((((uint32_t)cmf << 8) + flg) % 31 != 0) return false;
        if (flg & 0x20) return false;
        return true;
    }
};

struct File {
    bool ok = false;
    int status = DEFT4CU_ERR_PARSE;
    std::vector<Chunk> chunks;
    std::vector<ZS> zs;              // IDAT first, then the fdAT frames, then the others: the order of getDeflateStreams (:376-389)
    std::vector<std::vector<uint8_t>> owned;
    bool has_fdats = false;
};

size_t strlen_at(const uint8_t* d, size_t n, size_t off) {   // Util.strlen
    size_t i = off;
    while (i < n && d[i] != 0) i++;
    return i - off;
}

// PNGChunk.read (:162-215)
bool read_chunk(Reader& r, File& f, Chunk& c) {
    const uint64_t start = r.pos;
    const uint64_t length = r.rd32();
    if (length > INT_MAX_) return false;
    c.type = r.rd32();
    const uint64_t avail = r.pos < r.n ? std::min<uint64_t>(length, r.n - r.pos) : 0;
    c.data = r.p + std::min(r.pos, r.n);
    c.len = (uint32_t)length;
    r.pos += avail;
    const uint64_t pad = length - avail;   // Util.readFromInputStream leaves (byte) -1 where the file ended
    uint8_t ty[4];
    put_be32(ty, c.type);
    uint32_t crc = crc_update(~0u, ty, 4);
    crc = crc_update(crc, c.data, avail);
    if (pad) {
        static const std::vector<uint8_t> ff(1 << 16, 0xff);
        for (uint64_t left = pad; left;) { const uint64_t k = std::min<uint64_t>(left, ff.size()); crc = crc_update(crc, ff.data(), k); left -= k; }
    }
    crc = ~crc;
    const uint32_t want = r.rd32();
    if (crc != want) return false;
    if (pad) {   // (a truncated chunk whose CRC still matches)
        f.owned.emplace_back(length, 0xff);
        memcpy(f.owned.back().data(), c.data, avail);
        c.data = f.owned.back().data();
    } else if (start + 12 + length <= r.n) {
        c.raw = r.p + start;
    }
    if (c.hasSeq()) {
        if (c.len < 4) return false;   // the reference indexes data[0..3] and throws
        c.seq = be32(c.data);
    }
    return true;
}

// PNGFile.read with PNGChunkHelper's state machine; every zlib stream is assumed to parse (checked after the batch)
void read_file(const uint8_t* p, uint64_t n, File& f) {
    static const uint8_t SIG[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (n < 8 || memcmp(p, SIG, 8) != 0) return;
    Reader r{p, n, 8};
    bool readingIDAT = false, readingfdAT = false, seenacTL = false, seenIDAT = false, seenIEND = false, outOfOrder = false;
    uint32_t seq = 0;
    ZS idat, cur;
    std::vector<ZS> fdats, non;
    bool any_fdat = false, failed = false;
    auto flush = [&]() {
        if (readingIDAT) {
            idat = std::move(cur);
            cur = ZS();
            if (!idat.header()) return false;
            idat.role = 0;
            snprintf(idat.name, sizeof idat.name, "IDAT chunk");
            seenIDAT = true; readingIDAT = false;
        } else {
            any_fdat = true;
            fdats.push_back(std::move(cur));
            cur = ZS();
            ZS& z = fdats.back();
            if (!z.header()) return false;
            z.role = 1;
            snprintf(z.name, sizeof z.name, "fdAT chunk %d", (int)fdats.size());
            readingfdAT = false;
        }
        return true;
    };
    while (true) {
        Chunk c;
        if (!read_chunk(r, f, c)) { failed = true; break; }
        // submitChunkImpl (:493-556)
        if (seenIEND) { outOfOrder = true; failed = true; break; }
        if (c.hasSeq()) {
            if ((seenIDAT && !seenacTL) || c.seq != seq) { outOfOrder = true; failed = true; break; }
            seq++;
        }
        const bool should_flush = (readingIDAT && c.type != IDAT) || (readingfdAT && (c.type == fcTL || c.type == IEND));
        if (should_flush && !flush()) { failed = true; break; }
        if (c.type == IEND) {
            seenIEND = true;
            f.chunks.push_back(c);
            break;
        }
        if (c.type == acTL) {
            if (seenIDAT || seenacTL) { failed = true; break; }
            seenacTL = true;
            f.chunks.push_back(c);
            continue;
        }
        if (c.type == fdAT) {
            if (!seenIDAT || readingIDAT) { failed = true; break; }
            readingfdAT = true;
        } else if (c.type == IDAT) {
            if (seenIDAT || readingfdAT) { failed = true; break; }
            readingIDAT = true;
        }
        if (c.len > 0) {
            if (readingIDAT) cur.add(c.data, c.len);
            else if (readingfdAT && c.type == fdAT) cur.add(c.data + 4, c.len - 4);
            else if (c.zlibNonIdat()) {
                // getZLibCompressedNonIdat (:83-115)
                size_t off = strlen_at(c.data, c.len, 0) + 2;
                bool take = true;
                if (c.type == iTXt) {
                    if (off - 1 >= c.len) { failed = true; break; }   // the reference indexes past the array and throws
                    if (c.data[off - 1] != 1) take = false;
                    off += 1;
                }
                if (take) {
                    if (off - 1 >= c.len) { failed = true; break; }
                    if (c.data[off - 1] != 0) take = false;   // "Only deflate compression is currently supported"
                }
                if (take && c.type == iTXt) {
                    off += strlen_at(c.data, c.len, off) + 1;
                    off += strlen_at(c.data, c.len, off) + 1;
                }
                if (take) {
                    ZS z;
                    z.role = 2;
                    if (off < c.len) z.add(c.data + off, c.len - off);
                    if (z.header()) {
                        char ty[5] = {(char)(c.type >> 24), (char)(c.type >> 16), (char)(c.type >> 8), (char)c.type, 0};
                        snprintf(z.name, sizeof z.name, "%s chunk", ty);
                        c.non_id = (int)non.size();
                        // setZLibCompressedNonIdat's prefix (:117-138)
                        size_t keep = strlen_at(c.data, c.len, 0) + 2;
                        if (c.type == iTXt) {
                            keep += 1;
                            keep += strlen_at(c.data, c.len, keep) + 1;
                            keep += strlen_at(c.data, c.len, keep) + 1;
                        }
                        c.keep = (uint32_t)std::min<size_t>(keep, c.len);
                        non.push_back(std::move(z));
                    }
                }
            }
        }
        f.chunks.push_back(c);
    }
    // goodEndState (:558-571)
    const bool good = !failed && seenIEND && seenIDAT && !readingIDAT && !readingfdAT && !outOfOrder && (!any_fdat || seenacTL);
    if (!good) return;
    f.has_fdats = any_fdat;
    f.zs.push_back(std::move(idat));
    for (auto& z : fdats) f.zs.push_back(std::move(z));
    for (auto& z : non) f.zs.push_back(std::move(z));
    // a moved ZS keeps pointing at its own `joined` buffer
    for (auto& z : f.zs) if (!z.joined.empty()) { z.z = z.joined.data(); z.zlen = z.joined.size(); }
    f.ok = true;
}

// syncStreams (:262-369): the chunk list as it will be written.  false: "No IDAT chunk found" / "Incorrect chunk order in APNG"
bool sync_streams(File& f) {
    std::vector<Chunk>& ch = f.chunks;
    int index = -1;
    for (size_t i = 0; i < ch.size(); i++) if (ch[i].type == IDAT) { index = (int)i; break; }
    if (index < 0) return false;
    {
        std::vector<Chunk> keep;
        keep.reserve(ch.size());
        for (size_t i = 0; i < ch.size(); i++) {
            if ((int)i == index) {
                Chunk g;
                g.type = IDAT; g.gen = 0;
                keep.push_back(g);
            }
            if (ch[i].type != IDAT) keep.push_back(ch[i]);
        }
        ch.swap(keep);
    }
    if (f.has_fdats) {
        size_t i = 0;   // the ListIterator's cursor
        for (int zi = 1; zi < (int)f.zs.size() && f.zs[zi].role == 1; zi++) {
            int fdat_index = -1;
            while (i < ch.size()) {
                const uint32_t type = ch[i].type;
                i++;
                if (fdat_index == -1) {
                    if (type == fdAT) {
                        fdat_index = (int)i - 1;
                        Chunk g;
                        g.type = fdAT; g.gen = zi;
                        ch[i - 1] = g;      // remove() then add() at the cursor
                    }
                } else {
                    if (type == fdAT) { ch.erase(ch.begin() + (i - 1)); i--; }
                    if (type == fcTL) break;
                }
            }
            if (fdat_index == -1) return false;
        }
        uint32_t seq = 0;
        for (auto& c : ch) if (c.hasSeq()) { c.seq = seq++; c.new_seq = true; }
    }
    return true;
}

// PNGFile.write (:391-411) + PNGChunk.write (:140-158) + ZLibFile.write (:33-57, Adler-32 recalculated)
void write_file(const File& f, const deft4cu_result* res, uint8_t* out, uint64_t* out_len) {
    static const uint8_t SIG[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    const size_t first_non = f.zs.size() - [&] { size_t k = 0; for (auto& z : f.zs) k += z.role == 2; return k; }();
    uint8_t* w = out;
    if (w) memcpy(w, SIG, 8);
    uint64_t pos = 8;
    for (const Chunk& c : f.chunks) {
        const ZS* z = nullptr;
        uint32_t keep = 0;
        if (c.gen >= 0) z = &f.zs[c.gen];
        else if (c.non_id >= 0 && !f.zs[first_non + c.non_id].dropped) { z = &f.zs[first_non + c.non_id]; keep = c.keep; }
        if (!z) {
            if (w) {
                if (c.raw && !c.new_seq) memcpy(w + pos, c.raw, 12ull + c.len);
                else {
                    uint8_t* q = w + pos;
                    put_be32(q, c.len); put_be32(q + 4, c.type);
                    if (c.len) memcpy(q + 8, c.data, c.len);
                    if (c.new_seq) put_be32(q + 8, c.seq);
                    put_be32(q + 8 + c.len, ~crc_update(~0u, q + 4, 4ull + c.len));
                }
            }
            pos += 12ull + c.len;
            continue;
        }
        const deft4cu_result& r = res[z->slot];
        const uint32_t lead = c.gen >= 0 ? (c.type == fdAT ? 4u : 0u) : keep;
        const uint64_t len = lead + 2 + r.out_len + 4;
        if (w) {
            uint8_t* q = w + pos;
            put_be32(q, (uint32_t)len); put_be32(q + 4, c.type);
            uint8_t* d = q + 8;
            if (c.gen >= 0) { if (lead) put_be32(d, c.new_seq ? c.seq : 0); }
            else if (lead) memcpy(d, c.data, lead);
            d += lead;
            d[0] = z->cmf; d[1] = z->flg;
            if (r.out_len) memcpy(d + 2, r.out, r.out_len);
            put_be32(d + 2 + r.out_len, r.adler32);
            put_be32(q + 8 + len, ~crc_update(~0u, q + 4, 4 + len));
        }
        pos += 12 + len;
    }
    *out_len = pos;
}

}  // namespace

extern "C" {

uint32_t deft4cu_crc32(uint32_t crc, const uint8_t* data, uint64_t len) {
    std::call_once(g_tab_once, crc_tables);
    return ~crc_update(~crc, data, len);
}

int deft4cu_png_optimise_batch(const uint8_t* const* files, const uint64_t* lens, uint32_t n, uint32_t flags,
                               deft4cu_file_result* results) {
    if ((n && (!files || !lens)) || !results) return DEFT4CU_ERR_ARG;
    std::call_once(g_tab_once, crc_tables);
    static const bool timing = getenv("D4_HOST_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    auto lap = [&](const char* what) {
        if (timing) { auto t1 = now(); fprintf(stderr, "[deft4cu] png %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count()); t0 = t1; }
    };
    std::vector<File> F(n);
    for (uint32_t i = 0; i < n; i++) memset(&results[i], 0, sizeof results[i]);
    parallel_for(n, [&](uint32_t i) { read_file(files[i], lens[i], F[i]); });
    lap("read");
    // the zlib streams of every readable file: ONE list for the device (DeflateFilesContainer.java:18-43)
    std::vector<const uint8_t*> ptr;
    std::vector<uint64_t> len;
    for (auto& f : F) {
        if (!f.ok) continue;
        for (auto& z : f.zs) { z.slot = (int)ptr.size(); ptr.push_back(z.z + 2); len.push_back(z.zlen - 2); }
    }
    std::vector<deft4cu_result> R(ptr.size());
    int rc = DEFT4CU_OK;
    if (!ptr.empty()) {
        rc = deft4cu_optimise_batch(ptr.data(), len.data(), (uint32_t)ptr.size(), flags, R.data());
        if (rc != DEFT4CU_OK && rc != DEFT4CU_ERR_UNSUPPORTED && rc != DEFT4CU_ERR_PARSE) {
            deft4cu_free_results(R.data(), (uint32_t)R.size());
            return rc;
        }
    }
    lap("device");
    std::atomic<int> oom{0};
    parallel_for(n, [&](uint32_t i) {
        File& f = F[i];
        deft4cu_file_result& fr = results[i];
        fr.status = DEFT4CU_ERR_PARSE;
        if (!f.ok) return;
        // what the batch parse says about the streams step 1 took on trust
        for (auto& z : f.zs) {
            const int st = R[z.slot].status;
            if (st == DEFT4CU_OK) continue;
            if (z.role == 2 && st == DEFT4CU_ERR_PARSE) { z.dropped = true; continue; }   // helperNonIDAT never sees it (:532-543)
            fr.status = st == DEFT4CU_ERR_PARSE ? DEFT4CU_ERR_PARSE : DEFT4CU_ERR_UNSUPPORTED;
            return;
        }
        std::vector<std::string> names;
        std::vector<int64_t> saved;
        for (auto& z : f.zs) {
            if (z.dropped) continue;
            names.emplace_back(z.name);
            saved.push_back(R[z.slot].saved_bits);
        }
        if (!d4front::set_streams(fr, names, saved)) { oom = 1; return; }
        if (!sync_streams(f)) { fr.status = DEFT4CU_ERR_WRITE; return; }
        uint64_t need = 0;
        write_file(f, R.data(), nullptr, &need);
        fr.out = (uint8_t*)malloc(need ? need : 1);
        if (!fr.out) { oom = 1; return; }
        write_file(f, R.data(), fr.out, &fr.out_len);
        fr.status = DEFT4CU_OK;
    });
    lap("write");
    deft4cu_free_results(R.data(), (uint32_t)R.size());
    F.clear();
    lap("free");
    if (oom) { deft4cu_free_file_results(results, n); return DEFT4CU_ERR_ARG; }
    return DEFT4CU_OK;
}

void deft4cu_free_file_results(deft4cu_file_result* results, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) d4front::free_result(results[i]);
}

}  // extern "C"
