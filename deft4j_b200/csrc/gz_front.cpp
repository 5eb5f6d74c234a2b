// gz_front.cpp — native batch front-ends for gzip members and zlib streams (host code).
//
// Replaces, for a LIST of files, deft4j-container's GZFile (deft4j-container/src/main/java/com/github/NeRdTheNed/deft4j/
// container/GZFile.java: read :42-87, write :92-152 with CRC-32 and ISIZE recalculated, setFilename :158-169) and ZLibFile
// (container/ZLibFile.java: read :59-95, write :33-57 with Adler-32 recalculated), around the container's optimise
// (DeflateFilesContainer.java:18-43).  One stream per file; the streams of all files go to the device as ONE list through
// deft4cu_optimise_batch, the only thing this file calls.  Checksums and lengths come from the device (they are results of
// the batch entry): the decoded data never crosses PCIe.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/deft4cu.h"
#include "front_util.h"

namespace {

constexpr int FHCRC = 2, FEXTRA = 4, FNAME = 8, FCOMMENT = 16;

struct Member {
    bool ok = false;
    bool zlib = false;
    // gzip header fields as they are written back (GZFile.java:92-127)
    uint8_t flags = 0, xfl = 0, os = 0;
    uint8_t mtime[4] = {0, 0, 0, 0};
    const uint8_t* extra = nullptr; uint32_t extra_len = 0;
    const uint8_t* name = nullptr; uint32_t name_len = 0;
    const uint8_t* comment = nullptr; uint32_t comment_len = 0;
    uint8_t crc16[2] = {0, 0};
    uint8_t cmf = 0, flg = 0;     // zlib
    const uint8_t* body = nullptr; uint64_t body_len = 0;
    int slot = -1;
};

// GZFile.read up to the deflate stream.  Running out of bytes anywhere in the header leaves nothing for the stream to
// parse, so every such file is unreadable (the mirror reads -1s and then fails in DeflateStream.parse).
void read_gz(const uint8_t* d, uint64_t n, Member& m) {
    uint64_t p = 0;
    if (n < 10 || d[0] != 0x1f || d[1] != 0x8b || d[2] != 8) return;
    m.flags = d[3];
    if (m.flags & 0xe0) return;
    memcpy(m.mtime, d + 4, 4);
    m.xfl = d[8]; m.os = d[9];
    p = 10;
    if (m.flags & FEXTRA) {
        if (p + 2 > n) return;
        const uint32_t xlen = d[p] | (d[p + 1] << 8);
        p += 2;
        if (p + xlen > n) return;
        m.extra = d + p; m.extra_len = xlen;
        p += xlen;
    }
    auto cstr = [&](const uint8_t*& s, uint32_t& len) {   // Util.readStr: up to the NUL
        const uint64_t s0 = p;
        while (p < n && d[p] != 0) p++;
        s = d + s0; len = (uint32_t)(p - s0);
        if (p >= n) return false;   // no terminator: nothing follows
        p++;
        return true;
    };
    if (m.flags & FNAME) {
        if (!cstr(m.name, m.name_len)) return;
        if (m.name_len == 0) m.flags &= (uint8_t)~FNAME;   // setFilename("") clears the flag (:158-169)
    }
    if (m.flags & FCOMMENT) { if (!cstr(m.comment, m.comment_len)) return; }
    if (m.flags & FHCRC) {
        if (p + 2 > n) return;
        m.crc16[0] = d[p]; m.crc16[1] = d[p + 1];
        p += 2;
    }
    m.body = d + p; m.body_len = n - p;
    m.ok = true;
}

// ZLibFile.read up to the deflate stream (:59-87)
void read_zlib(const uint8_t* d, uint64_t n, Member& m) {
    m.zlib = true;
    if (n < 2) return;
    m.cmf = d[0]; m.flg = d[1];
    if ((m.cmf & 0xF) != 8) return;
    if ((((uint32_t)m.cmf << 8) + m.flg) % 31 != 0) return;
    if (m.flg & 0x20) return;     // preset dictionary
    m.body = d + 2; m.body_len = n - 2;
    m.ok = true;
}

uint64_t write_member(const Member& m, const deft4cu_result& r, uint8_t* out) {
    uint64_t pos = 0;
    auto put = [&](const void* s, uint64_t k) { if (out && k) memcpy(out + pos, s, k); pos += k; };
    auto byte = [&](uint8_t b) { if (out) out[pos] = b; pos++; };
    if (m.zlib) {
        byte(m.cmf); byte(m.flg);
        put(r.out, r.out_len);
        byte((uint8_t)(r.adler32 >> 24)); byte((uint8_t)(r.adler32 >> 16)); byte((uint8_t)(r.adler32 >> 8)); byte((uint8_t)r.adler32);
        return pos;
    }
    byte(0x1f); byte(0x8b); byte(8); byte(m.flags);
    put(m.mtime, 4);
    byte(m.xfl); byte(m.os);
    if (m.flags & FEXTRA) { byte((uint8_t)m.extra_len); byte((uint8_t)(m.extra_len >> 8)); put(m.extra, m.extra_len); }
    if (m.flags & FNAME) { put(m.name, m.name_len); byte(0); }
    if (m.flags & FCOMMENT) put(m.comment, m.comment_len);      // written WITHOUT its NUL terminator (GZFile.java:117-119)
    if (m.flags & FHCRC) put(m.crc16, 2);
    put(r.out, r.out_len);
    const uint32_t crc = r.crc32, isize = (uint32_t)r.uncompressed_len;
    byte((uint8_t)crc); byte((uint8_t)(crc >> 8)); byte((uint8_t)(crc >> 16)); byte((uint8_t)(crc >> 24));
    byte((uint8_t)isize); byte((uint8_t)(isize >> 8)); byte((uint8_t)(isize >> 16)); byte((uint8_t)(isize >> 24));
    return pos;
}

int run(const uint8_t* const* files, const uint64_t* lens, uint32_t n, uint32_t flags, deft4cu_file_result* results, bool zlib) {
    if ((n && (!files || !lens)) || !results) return DEFT4CU_ERR_ARG;
    std::vector<Member> M(n);
    for (uint32_t i = 0; i < n; i++) memset(&results[i], 0, sizeof results[i]);
    std::vector<const uint8_t*> ptr;
    std::vector<uint64_t> len;
    for (uint32_t i = 0; i < n; i++) {
        if (zlib) read_zlib(files[i], lens[i], M[i]); else read_gz(files[i], lens[i], M[i]);
        if (!M[i].ok) continue;
        M[i].slot = (int)ptr.size();
        ptr.push_back(M[i].body);
        len.push_back(M[i].body_len);
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
        const Member& m = M[i];
        deft4cu_file_result& fr = results[i];
        fr.status = DEFT4CU_ERR_PARSE;
        if (!m.ok) return;
        const deft4cu_result& r = R[m.slot];
        if (r.status != DEFT4CU_OK) { fr.status = r.status == DEFT4CU_ERR_PARSE ? DEFT4CU_ERR_PARSE : DEFT4CU_ERR_UNSUPPORTED; return; }
        std::string name = "unnamed stream";    // DeflateStream.java:19; a gzip member's stream is named after FNAME (:77-79)
        if (!m.zlib && m.name_len) name.assign((const char*)m.name, m.name_len);
        if (!d4front::set_streams(fr, {name}, {r.saved_bits})) { oom = 1; return; }
        const uint64_t need = write_member(m, r, nullptr);
        fr.out = (uint8_t*)malloc(need ? need : 1);
        if (!fr.out) { oom = 1; return; }
        fr.out_len = write_member(m, r, fr.out);
        fr.status = DEFT4CU_OK;
    }, 4);
    deft4cu_free_results(R.data(), (uint32_t)R.size());
    if (oom) { deft4cu_free_file_results(results, n); return DEFT4CU_ERR_ARG; }
    return DEFT4CU_OK;
}

}  // namespace

extern "C" int deft4cu_gz_optimise_batch(const uint8_t* const* files, const uint64_t* lens, uint32_t n, uint32_t flags,
                                         deft4cu_file_result* results) {
    return run(files, lens, n, flags, results, false);
}
extern "C" int deft4cu_zlib_optimise_batch(const uint8_t* const* files, const uint64_t* lens, uint32_t n, uint32_t flags,
                                           deft4cu_file_result* results) {
    return run(files, lens, n, flags, results, true);
}
