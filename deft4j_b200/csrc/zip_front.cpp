// zip_front.cpp — native batch front-end for ZIP archives (host code).
//
// Replaces, for a LIST of archives, deft4j-container's ZipFile (deft4j-container/src/main/java/com/github/NeRdTheNed/
// deft4j/container/ZipFile.java): read :82-127 (entries with compression method 8 become DeflateStreams, everything
// else is carried through), the container's optimise (DeflateFilesContainer.java:18-43) and write :46-79 through
// RecalculatingZipWriter (container/lljzip/RecalculatingZipWriter.java:23-136: local headers rewritten with the CRC and
// sizes of their central directory entries, central directory with recomputed offsets, end record with recomputed
// counts, size and offset).
// The reference reads archives with the third-party lljzip 2.3.0 (`ZipIO.readStandard`), which is not vendored: the
// reader here restates the standard strategy of the ZIP application note exactly as the Python mirror
// (deft4j_b200/container/zip_file.py) does — PARITY UNPINNED, like the mirror.  The entry streams of
// ALL archives go to the device as ONE list through deft4cu_optimise_batch, the only thing this file calls.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/deft4cu.h"
#include "front_util.h"

namespace {

constexpr uint32_t SIG_LOCAL = 0x04034B50u, SIG_CENTRAL = 0x02014B50u, SIG_END = 0x06054B50u;

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void put16(uint8_t*& w, uint32_t v) { w[0] = (uint8_t)v; w[1] = (uint8_t)(v >> 8); w += 2; }
inline void put32(uint8_t*& w, uint32_t v) { w[0] = (uint8_t)v; w[1] = (uint8_t)(v >> 8); w[2] = (uint8_t)(v >> 16); w[3] = (uint8_t)(v >> 24); w += 4; }

// a byte range of the archive, clamped to it (the mirror slices; a slice past the end is short, not an error)
struct Span {
    const uint8_t* p = nullptr;
    uint64_t n = 0;
};
inline Span slice(const uint8_t* base, uint64_t len, uint64_t from, uint64_t count) {
    Span s;
    if (from >= len) { s.p = base + len; s.n = 0; return s; }
    s.p = base + from;
    s.n = std::min<uint64_t>(count, len - from);
    return s;
}

struct Central;
struct Local {
    uint16_t version, flags, method, mtime, mdate;
    uint32_t crc32, csize, usize;
    Span name, extra, data;
    uint64_t offset;
    int central;       // index into Archive::centrals
    int slot = -1;     // deflated entry: position in the flat list handed to the device
};
struct Central {
    uint16_t made_by, version, flags, method, mtime, mdate, disk, iattr;
    uint32_t crc32, csize, usize, eattr, offset;
    Span name, extra, comment;
    int local = -1;    // index into Archive::locals (before sorting: see `order`)
};
struct Archive {
    bool ok = false;
    uint16_t end_disk = 0, end_start_disk = 0;
    Span comment;
    std::vector<Central> centrals;
    std::vector<Local> locals;           // in creation (central directory) order
    std::vector<int> order;              // locals sorted by offset (stable): the order of the written file and of the streams
};

// ZipFile.read (:82-127) over the standard reading strategy
void read_archive(const uint8_t* d, uint64_t len, Archive& a) {
    // last occurrence of the end-of-central-directory signature
    int64_t end = -1;
    if (len >= 4)
        for (int64_t i = (int64_t)len - 4; i >= 0; i--)
            if (d[i] == 0x50 && d[i + 1] == 0x4B && d[i + 2] == 0x05 && d[i + 3] == 0x06) { end = i; break; }
    if (end < 0 || (uint64_t)end + 22 > len) return;
    const uint8_t* e = d + end;
    a.end_disk = le16(e + 4);
    a.end_start_disk = le16(e + 6);
    const uint32_t n_total = le16(e + 10), cd_size = le32(e + 12), cd_off = le32(e + 16), clen = le16(e + 20);
    if (n_total == 0xFFFF || cd_off == 0xFFFFFFFFu || cd_size == 0xFFFFFFFFu) return;   // Zip64: refused
    a.comment = slice(d, len, (uint64_t)end + 22, clen);
    uint64_t p = cd_off;
    for (uint32_t k = 0; k < n_total; k++) {
        if (p + 46 > len || le32(d + p) != SIG_CENTRAL) break;
        const uint8_t* c = d + p;
        Central x;
        x.made_by = le16(c + 4); x.version = le16(c + 6); x.flags = le16(c + 8); x.method = le16(c + 10);
        x.mtime = le16(c + 12); x.mdate = le16(c + 14); x.crc32 = le32(c + 16); x.csize = le32(c + 20); x.usize = le32(c + 24);
        const uint32_t nlen = le16(c + 28), xlen = le16(c + 30), klen = le16(c + 32);
        x.disk = le16(c + 34); x.iattr = le16(c + 36); x.eattr = le32(c + 38); x.offset = le32(c + 42);
        x.name = slice(d, len, p + 46, nlen);
        x.extra = slice(d, len, p + 46 + nlen, xlen);
        x.comment = slice(d, len, p + 46 + nlen + xlen, klen);
        p += 46ull + nlen + xlen + klen;
        a.centrals.push_back(x);
    }
    for (size_t ci = 0; ci < a.centrals.size(); ci++) {
        Central& c = a.centrals[ci];
        const uint64_t o = c.offset;
        if (o + 30 > len || le32(d + o) != SIG_LOCAL) continue;
        const uint8_t* h = d + o;
        Local l;
        l.version = le16(h + 4); l.flags = le16(h + 6); l.method = le16(h + 8); l.mtime = le16(h + 10); l.mdate = le16(h + 12);
        l.crc32 = le32(h + 14); l.csize = le32(h + 18); l.usize = le32(h + 22);
        const uint32_t nlen = le16(h + 26), xlen = le16(h + 28);
        l.name = slice(d, len, o + 30, nlen);
        l.extra = slice(d, len, o + 30 + nlen, xlen);
        l.offset = o;
        l.central = (int)ci;
        // data descriptors leave the local sizes zero: take the central directory's (ZipFile.java:104-107)
        if (l.csize == 0 && !(l.csize == c.csize && l.usize == c.usize && l.crc32 == c.crc32)) {
            l.csize = c.csize; l.usize = c.usize; l.crc32 = c.crc32;
        }
        l.data = slice(d, len, o + 30ull + nlen + xlen, l.csize);
        c.local = (int)a.locals.size();
        a.locals.push_back(l);
    }
    if (a.locals.empty()) return;
    a.order.resize(a.locals.size());
    for (size_t i = 0; i < a.order.size(); i++) a.order[i] = (int)i;
    std::stable_sort(a.order.begin(), a.order.end(), [&](int x, int y) { return a.locals[x].offset < a.locals[y].offset; });
    a.ok = true;
}

// ZipFile.write (:46-79) + RecalculatingZipWriter (:23-136).  out == nullptr: size only.  false: a central directory
// entry whose local header was not found ("could not find old offset")
bool write_archive(const Archive& a, const deft4cu_result* res, uint8_t* out, uint64_t* out_len) {
    // sizes after syncStreams: a deflated entry carries its rewritten stream
    auto data_len = [&](const Local& l) -> uint64_t { return l.slot >= 0 ? res[l.slot].out_len : l.data.n; };
    std::map<uint32_t, uint64_t> new_off;
    uint8_t* w = out;
    uint64_t pos = 0;
    for (int li : a.order) {
        const Local& l = a.locals[li];
        const Central& c = a.centrals[l.central];
        const uint32_t csize = l.slot >= 0 ? (uint32_t)res[l.slot].out_len : c.csize;
        new_off[c.offset] = pos;
        const uint64_t n = 30 + l.name.n + l.extra.n + data_len(l);
        if (w) {
            uint8_t* q = w + pos;
            put32(q, SIG_LOCAL); put16(q, l.version); put16(q, l.flags); put16(q, l.method); put16(q, l.mtime); put16(q, l.mdate);
            put32(q, c.crc32); put32(q, csize); put32(q, c.usize); put16(q, (uint32_t)l.name.n); put16(q, (uint32_t)l.extra.n);
            if (l.name.n) memcpy(q, l.name.p, l.name.n);
            q += l.name.n;
            if (l.extra.n) memcpy(q, l.extra.p, l.extra.n);
            q += l.extra.n;
            if (l.slot >= 0) { if (res[l.slot].out_len) memcpy(q, res[l.slot].out, res[l.slot].out_len); }
            else if (l.data.n) memcpy(q, l.data.p, l.data.n);
        }
        pos += n;
    }
    const uint64_t start_central = pos;
    uint32_t count = 0;
    for (const Central& c : a.centrals) {
        auto it = new_off.find(c.offset);
        if (it == new_off.end()) return false;
        uint32_t csize = c.csize, usize = c.usize;
        if (c.local >= 0) {
            const Local& l = a.locals[c.local];
            csize = l.slot >= 0 ? (uint32_t)res[l.slot].out_len : l.csize;
            usize = l.usize;
        }
        if (w) {
            uint8_t* q = w + pos;
            put32(q, SIG_CENTRAL); put16(q, c.made_by); put16(q, c.version); put16(q, c.flags); put16(q, c.method); put16(q, c.mtime);
            put16(q, c.mdate); put32(q, c.crc32); put32(q, csize); put32(q, usize); put16(q, (uint32_t)c.name.n);
            put16(q, (uint32_t)c.extra.n); put16(q, (uint32_t)c.comment.n); put16(q, c.disk); put16(q, c.iattr); put32(q, c.eattr);
            put32(q, (uint32_t)it->second);
            if (c.name.n) memcpy(q, c.name.p, c.name.n);
            q += c.name.n;
            if (c.extra.n) memcpy(q, c.extra.p, c.extra.n);
            q += c.extra.n;
            if (c.comment.n) memcpy(q, c.comment.p, c.comment.n);
        }
        pos += 46 + c.name.n + c.extra.n + c.comment.n;
        count++;
    }
    const uint64_t central_size = pos - start_central;
    if (w) {
        uint8_t* q = w + pos;
        put32(q, SIG_END); put16(q, a.end_disk); put16(q, a.end_start_disk); put16(q, count); put16(q, count);
        put32(q, (uint32_t)central_size); put32(q, (uint32_t)start_central); put16(q, (uint32_t)a.comment.n);
        if (a.comment.n) memcpy(q, a.comment.p, a.comment.n);
    }
    pos += 22 + a.comment.n;
    *out_len = pos;
    return true;
}

}  // namespace

extern "C" int deft4cu_zip_optimise_batch(const uint8_t* const* files, const uint64_t* lens, uint32_t n, uint32_t flags,
                                          deft4cu_file_result* results) {
    if ((n && (!files || !lens)) || !results) return DEFT4CU_ERR_ARG;
    std::vector<Archive> A(n);
    for (uint32_t i = 0; i < n; i++) memset(&results[i], 0, sizeof results[i]);
    d4front::parallel_for(n, [&](uint32_t i) { read_archive(files[i], lens[i], A[i]); }, 1);
    // the deflated entries of every archive, in local-file order: ONE list for the device
    std::vector<const uint8_t*> ptr;
    std::vector<uint64_t> len;
    for (auto& a : A) {
        if (!a.ok) continue;
        for (int li : a.order) {
            Local& l = a.locals[li];
            if (l.method != 8) continue;        // stored / other methods are carried through untouched (:97-99)
            l.slot = (int)ptr.size();
            ptr.push_back(l.data.p);
            len.push_back(l.data.n);
        }
    }
    std::vector<deft4cu_result> R(ptr.size());
    if (!ptr.empty()) {
        const int rc = deft4cu_optimise_batch(ptr.data(), len.data(), (uint32_t)ptr.size(), flags, R.data());
        if (rc != DEFT4CU_OK && rc != DEFT4CU_ERR_UNSUPPORTED && rc != DEFT4CU_ERR_PARSE) {
            deft4cu_free_results(R.data(), (uint32_t)R.size());
            return rc;
        }
    }
    std::atomic<int> oom{0};
    d4front::parallel_for(n, [&](uint32_t i) {
        Archive& a = A[i];
        deft4cu_file_result& fr = results[i];
        fr.status = DEFT4CU_ERR_PARSE;
        if (!a.ok) return;
        std::vector<std::string> names;
        std::vector<int64_t> saved;
        for (int li : a.order) {
            const Local& l = a.locals[li];
            if (l.slot < 0) continue;
            const int st = R[l.slot].status;
            if (st != DEFT4CU_OK) {   // "Failed to parse stream for file ..." -> read returns false (:113-117)
                fr.status = st == DEFT4CU_ERR_PARSE ? DEFT4CU_ERR_PARSE : DEFT4CU_ERR_UNSUPPORTED;
                return;
            }
            names.emplace_back(l.name.n ? std::string((const char*)l.name.p, l.name.n) : std::string("unnamed stream"));
            saved.push_back(R[l.slot].saved_bits);
        }
        if (!d4front::set_streams(fr, names, saved)) { oom = 1; return; }
        uint64_t need = 0;
        if (!write_archive(a, R.data(), nullptr, &need)) { fr.status = DEFT4CU_ERR_WRITE; return; }
        fr.out = (uint8_t*)malloc(need ? need : 1);
        if (!fr.out) { oom = 1; return; }
        write_archive(a, R.data(), fr.out, &fr.out_len);
        fr.status = DEFT4CU_OK;
    }, 1);
    deft4cu_free_results(R.data(), (uint32_t)R.size());
    if (oom) { deft4cu_free_file_results(results, n); return DEFT4CU_ERR_ARG; }
    return DEFT4CU_OK;
}
