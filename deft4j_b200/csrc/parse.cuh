// parse.cuh — Huffman decode of deflate streams on the GPU.
//
// Pass 0  k_find  : one CTA per 128 KiB segment of a long stream: first well-formed dynamic block header in it.
// Pass 1  k_count : one CTA per walker (a stream, or a segment of one starting at the header k_find guessed;
//                   the host validates the chain, deft4cu.cu Batch::parse).  Blocks are walked in order; inside
//                   a Huffman block the CTA decodes a window of NT*CHUNK_BITS bits
//                   speculatively: thread t starts at bit t*CHUNK_BITS assuming a symbol boundary,
//                   then a fix-up loop re-decodes every chunk whose predecessor ended elsewhere until
//                   the chain from the (known) window start is consistent (self-synchronisation).
//                   Output: BlockRec per block (code tables, header pairs, exact bit sizes), one
//                   ChunkRec per valid chunk (start bit, symbol index, decoded offset), stream totals.
// Pass 2  k_emit  : one CTA per Huffman block, one thread per ChunkRec: decode again, now writing the
//                   packed symbols and their decoded offsets.  Stored blocks are copied.
// Pass 3  k_lz_*  : LZ77 resolution by pointer jumping over the decoded bytes (every byte of a match
//                   points at its source byte; literals are roots; up to 8 links per pass), replacing the reference's
//                   byte-serial readSlice (DeflateBlock.java:147-222).
//
// Decoding restates the reference decoder's semantics (Huffman.java:170-197): codes are matched one
// length at a time, first match wins, no validity check on the length set, failure after 15 bits.
#pragma once
#include "common.cuh"

namespace d4 {

constexpr int PARSE_NT = 1024;       // threads per stream CTA in k_count
constexpr int CHUNK_BITS = 256;      // speculative chunk; must exceed the longest unit (48 bits)

struct DecTab {
    uint32_t first[16];   // canonical code value of the first code of each length
    uint16_t count[16];
    uint16_t offs[16];
    uint16_t symtab[MAX_LL];
};

// Huffman.buildCodes (Huffman.java:35-64) turned into a by-length decode table.
__device__ inline void build_dectab(const uint8_t* lens, int n, DecTab& t) {
    for (int l = 0; l < 16; l++) { t.count[l] = 0; t.first[l] = 0; t.offs[l] = 0; }
    for (int i = 0; i < n; i++) if (lens[i] > 0 && lens[i] < 16) t.count[lens[i]]++;
    uint32_t next = 0;
    int lastShift = 0, off = 0;
    for (int l = 1; l < 16; l++) {
        t.offs[l] = (uint16_t)off;
        off += t.count[l];
        if (t.count[l] == 0) continue;
        next <<= (l - lastShift);
        lastShift = l;
        t.first[l] = next;
        next += t.count[l];
    }
    uint16_t fill[16];
    for (int l = 0; l < 16; l++) fill[l] = t.offs[l];
    for (int i = 0; i < n; i++) if (lens[i] > 0 && lens[i] < 16) t.symtab[fill[lens[i]]++] = (uint16_t)i;
}

// returns symbol (>= 0) and its length through *len, or -1 (no code within 15 bits)
__device__ __forceinline__ int decode_sym(const DecTab& t, uint64_t w, int* len) {
    uint32_t code = 0;
#pragma unroll 1
    for (int l = 1; l <= 15; l++) {
        code = (code << 1) | (uint32_t)(w & 1);
        w >>= 1;
        uint32_t idx = code - t.first[l];
        if (idx < t.count[l]) { *len = l; return t.symtab[t.offs[l] + idx]; }
    }
    return -1;
}

// Multi-bit lookup in front of the by-length decoder: slot i of the table holds what decode_sym finds in the bit pattern i
// when a code of at most LUT bits matches ((len << 9) | symbol; 0: no such code).  Every slot is filled by running the
// reference's first-match-by-length rule on its own index, so incomplete and over-subscribed length sets decode exactly
// as Huffman.readSymbol does (Huffman.java:170-197); codes longer than the table fall through to decode_sym.
constexpr int LUT_L_BITS = 10, LUT_D_BITS = 9;
struct DecLut {
    uint16_t lit[1 << LUT_L_BITS];
    uint16_t dst[1 << LUT_D_BITS];
};
__device__ __forceinline__ void build_lut(const DecTab& t, uint16_t* lut, int bits, int tid, int nthreads) {
    for (int i = tid; i < (1 << bits); i += nthreads) {
        uint32_t code = 0, w = (uint32_t)i;
        uint16_t e = 0;
#pragma unroll 1
        for (int l = 1; l <= bits; l++) {
            code = (code << 1) | (w & 1u);
            w >>= 1;
            const uint32_t idx = code - t.first[l];
            if (idx < t.count[l]) { e = (uint16_t)((l << 9) | t.symtab[t.offs[l] + idx]); break; }
        }
        lut[i] = e;
    }
}
__device__ __forceinline__ int decode_sym_lut(const DecTab& t, const uint16_t* lut, int bits, uint64_t w, int* len) {
    const uint32_t e = lut[(uint32_t)w & ((1u << bits) - 1u)];
    if (e) { *len = (int)(e >> 9); return (int)(e & 511u); }
    return decode_sym(t, w, len);
}

// ---- staging of the bitstream window in shared memory --------------------------------------------------------------
// k_count decodes a window of PARSE_NT chunks at a time, speculatively and then again in the fix-up rounds: the
// window's bytes are copied into shared memory once with cp.async (LDGSTS: global -> shared without passing through
// registers) and every peek reads shared memory.  `avail` bytes of the source are valid; the rest is zero-filled.
constexpr int WIN_BYTES = PARSE_NT * CHUNK_BITS / 8;      // 32 KiB
constexpr int WIN_SLACK = 64;                              // a unit is <= 48 bits and a peek reads 16 bytes
constexpr int HDR_STAGE = 1024;                            // a dynamic header is at most 4551 bits (+ 16 bytes of alignment and a peek's over-read)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
// >= 57 valid bits starting at bit `rel` of the staged window
__device__ __forceinline__ uint64_t peek_bits_smem(const uint8_t* win, uint32_t rel) {
    const uint32_t byte = rel >> 3;
    const uint64_t* q = (const uint64_t*)(win + (byte & ~7u));
    const int sh = (int)(byte & 7u) * 8;
    const uint64_t lo = q[0], hi = q[1];
    const uint64_t w = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    return w >> (rel & 7u);
}

struct Unit {
    int nbits;       // bits consumed, 0 = decode error
    int outlen;      // decoded bytes produced
    uint32_t packed; // packed symbol
    int eob;
};

// One literal / EOB / match unit (DeflateBlockHuffman.decodeStream, :778-890).
__device__ __forceinline__ Unit decode_unit_w(const DecTab& lit, const DecTab& dst, const DecLut& lut, uint64_t w);
__device__ __forceinline__ Unit decode_unit(const DecTab& lit, const DecTab& dst, const DecLut& lut, const uint8_t* in, uint64_t pos) {
    return decode_unit_w(lit, dst, lut, peek_bits(in, pos));
}
// the same from the (>= 57) bits at the unit's position
__device__ __forceinline__ Unit decode_unit_w(const DecTab& lit, const DecTab& dst, const DecLut& lut, uint64_t w) {
    Unit u;
    u.nbits = 0; u.outlen = 0; u.packed = 0; u.eob = 0;
    int l;
    int s = decode_sym_lut(lit, lut.lit, LUT_L_BITS, w, &l);
    if (s < 0 || s > 285) return u;
    if (s < 256) { u.nbits = l; u.outlen = 1; u.packed = (uint32_t)s; return u; }
    if (s == 256) { u.nbits = l; u.eob = 1; u.packed = 256; return u; }
    int nb = l;
    w >>= l;
    int eb = c_len_ebits[s - 257];
    int len = c_len_base[s - 257] + (int)(w & ((1u << eb) - 1));
    w >>= eb; nb += eb;
    int edge = (len == 258 && s == 284) ? 1 : 0;
    int dl;
    int ds = decode_sym_lut(dst, lut.dst, LUT_D_BITS, w, &dl);
    if (ds < 0 || ds > 29) return u;
    w >>= dl; nb += dl;
    int deb = c_dist_ebits[ds];
    int dist = c_dist_base[ds] + (int)(w & ((1u << deb) - 1));
    nb += deb;
    u.nbits = nb; u.outlen = len;
    u.packed = sym_pack_match(len, dist, edge, s);
    return u;
}

__device__ inline void fixed_lens(uint8_t* L, uint8_t* D) {  // HuffmanTable.java:166-209 (286 / 30 entries)
    for (int i = 0; i < 286; i++) L[i] = (i <= 143) ? 8 : (i <= 255) ? 9 : (i <= 279) ? 7 : 8;
    for (int i = 0; i < 30; i++) D[i] = 5;
}

// DeflateBlockHuffman.initDynamicDecoder (:892-1010), run by one thread.  Returns bits consumed or 0.
// `peek(p)` returns the (>= 57) bits at stream bit position p; `lens`: MAX_LL + MAX_D bytes of scratch.
template <class Peek>
__device__ inline uint32_t parse_dynamic_header(Peek peek, uint64_t pos, uint64_t total_bits, BlockRec& b, DecTab& scratch,
                                                uint8_t* lens) {
    uint64_t p = pos;
    if (p + 14 > total_bits) return 0;
    uint64_t w = peek(p);
    int nL = (int)(w & 31) + 257, nD = (int)((w >> 5) & 31) + 1, ncl = (int)((w >> 10) & 15) + 4;
    p += 14;
    if (p + 3ull * ncl > total_bits) return 0;
    for (int i = 0; i < 19; i++) b.hdr.CL[i] = 0;
    for (int i = 0; i < ncl; i++) {
        b.hdr.CL[c_codelen_order[i]] = (uint8_t)(peek(p) & 7);
        p += 3;
    }
    int hbits = 14 + 3 * ncl;
    build_dectab(b.hdr.CL, 19, scratch);
    // HLIT up to 288 is accepted by the reference parser (asserts only); the tables then have 287/288
    // entries.  Our Tab holds 288.
    int i = 0, np = 0;
    const int combined = nL + nD;
    while (i < combined) {
        if (p >= total_bits) return 0;
        w = peek(p);
        int l;
        int s = decode_sym(scratch, w, &l);
        if (s < 0) return 0;
        p += l; hbits += l; w >>= l;
        int run = 0, val;
        if (s <= 15) { lens[i++] = (uint8_t)s; val = s; }
        else if (s == 16) {
            if (i < 1) return 0;
            run = (int)(w & 3) + 3; p += 2; hbits += 2;
            if (i + run > combined) return 0;
            val = lens[i - 1];
            for (int k = 0; k < run; k++) lens[i++] = (uint8_t)val;
        } else if (s == 17) {
            run = (int)(w & 7) + 3; p += 3; hbits += 3;
            if (i + run > combined) return 0;
            val = 0;
            for (int k = 0; k < run; k++) lens[i++] = 0;
        } else {
            run = (int)(w & 127) + 11; p += 7; hbits += 7;
            if (i + run > combined) return 0;
            val = 0;
            for (int k = 0; k < run; k++) lens[i++] = 0;
        }
        if (p > total_bits) return 0;
        b.hdr.pairs[np++] = pair_pack(s, run, val);
    }
    b.hdr.np = (uint16_t)np;
    b.hdr.ncl = (uint8_t)ncl;
    b.hdr.bits = hbits;
    b.tab.nL = (uint16_t)nL; b.tab.nD = (uint16_t)nD; b.tab.type = 2;
    for (int k = 0; k < MAX_LL; k++) b.tab.L[k] = k < nL ? lens[k] : 0;
    for (int k = 0; k < MAX_D; k++) b.tab.D[k] = k < nD ? lens[nL + k] : 0;
    return (uint32_t)(p - pos);
}

// status codes shared with the host (mirror DEFT4CU_*); ST_NONE / ST_ABORT only ever describe speculative walkers
constexpr int ST_OK = 0, ST_PARSE = 1, ST_UNSUPPORTED = 3, ST_NONE = 100, ST_ABORT = 101;

// --------------------------------------------------------------------------------------------------
// k_find: block boundaries inside a stream, found on the device.  One CTA per speculative walker scans
// its segment for the first bit position that holds a well-formed dynamic block header: BTYPE = 2,
// HLIT <= 29, HDIST <= 29, a COMPLETE code-length code (Kraft sum exactly 1), code lengths that decode
// to exactly HLIT + HDIST entries, an end-of-block code, a complete litlen code and a complete (or
// empty / single-code) distance code.  The test only has to be a good guess: a walker's result is used
// only if the chain from bit 0 ends exactly at its start, and a boundary the test misses (fixed or
// stored blocks, incomplete codes the reference accepts) is re-walked from the known
// position.  Stage 1 (cheap, every bit position) runs on all threads; survivors are queued in shared
// memory and stage 2 (full header decode) is spread over the warps.
// --------------------------------------------------------------------------------------------------
constexpr int FIND_NT = 256;
constexpr int FIND_TILE = FIND_NT * 8;   // bit positions per tile

__device__ __forceinline__ uint64_t funnel128(uint64_t lo, uint64_t hi, int k) {  // bits [k, k+64) of hi:lo, 0 <= k < 64
    return k ? ((lo >> k) | (hi << (64 - k))) : lo;
}

__device__ inline bool find_full_check(const uint8_t* in, uint64_t p, uint64_t total_bits) {
    uint64_t w = peek_bits(in, p);
    const int nL = (int)((w >> 3) & 31) + 257, nD = (int)((w >> 8) & 31) + 1, ncl = (int)((w >> 13) & 15) + 4;
    uint64_t q = p + 17;
    if (q + 3ull * ncl > total_bits) return false;
    uint64_t cw = peek_bits(in, q);
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (int i = 0; i < ncl; i++) cl[c_codelen_order[i]] = (uint8_t)((cw >> (3 * i)) & 7);
    q += 3ull * ncl;
    // canonical decode by length (<= 7 bits)
    int count[8], first[8], offs[8];
    uint8_t symtab[19];
    for (int l = 0; l < 8; l++) count[l] = 0;
    for (int s = 0; s < 19; s++) count[cl[s]]++;
    count[0] = 0;
    int code = 0, off = 0;
    for (int l = 1; l < 8; l++) { code = (code + count[l - 1]) << 1; first[l] = code; offs[l] = off; off += count[l]; }
    {
        int fill[8];
        for (int l = 0; l < 8; l++) fill[l] = offs[l];
        for (int s = 0; s < 19; s++) if (cl[s]) symtab[fill[cl[s]]++] = (uint8_t)s;
    }
    const int combined = nL + nD;
    int i = 0, prev = -1, len256 = 0, nzL = 0, nzD = 0;
    uint32_t kl = 0, kd = 0;  // Kraft sums in units of 2^-15
    while (i < combined) {
        if (q >= total_bits) return false;
        w = peek_bits(in, q);
        int c = 0, l = 0, sym = -1;
        for (l = 1; l <= 7; l++) {
            c = (c << 1) | (int)(w & 1);
            w >>= 1;
            int idx = c - first[l];
            if (idx >= 0 && idx < count[l]) { sym = symtab[offs[l] + idx]; break; }
        }
        if (sym < 0) return false;
        q += l;
        int run = 1, val = sym;
        if (sym == 16) { if (prev < 0) return false; run = (int)(w & 3) + 3; q += 2; val = prev; }
        else if (sym == 17) { run = (int)(w & 7) + 3; q += 3; val = 0; }
        else if (sym == 18) { run = (int)(w & 127) + 11; q += 7; val = 0; }
        if (i + run > combined) return false;
        if (val) {
            for (int k = 0; k < run; k++) {
                const int idx = i + k;
                if (idx < nL) { kl += 32768u >> val; nzL++; if (idx == 256) len256 = val; }
                else { kd += 32768u >> val; nzD++; }
            }
            if (kl > 32768u || kd > 32768u) return false;
        }
        i += run;
        prev = val;
    }
    if (q > total_bits) return false;
    if (!len256) return false;
    if (!(kl == 32768u || (nzL == 1 && kl == 16384u))) return false;
    if (!(kd == 32768u || nzD == 0 || (nzD == 1 && kd == 16384u))) return false;
    return true;
}

__global__ void __launch_bounds__(FIND_NT)
k_find(const uint8_t* __restrict__ d_in, StreamDesc* __restrict__ descs, const uint32_t* __restrict__ list, uint64_t seg_bits) {
    const uint32_t wid = list[blockIdx.x];
    StreamDesc& sd = descs[wid];
    const uint8_t* in = d_in + sd.in_off;
    const uint64_t total_bits = sd.in_len * 8;
    const uint64_t seg_begin = (uint64_t)(wid - sd.walker0) * seg_bits;
    uint64_t seg_end = seg_begin + seg_bits;
    if (seg_end > total_bits) seg_end = total_bits;
    const int t = threadIdx.x;
    __shared__ uint16_t s_q[FIND_TILE];
    __shared__ unsigned s_n, s_best;
    uint64_t found = BIT_NONE;
    for (uint64_t tile = seg_begin; tile < seg_end && found == BIT_NONE; tile += FIND_TILE) {
        if (t == 0) { s_n = 0; s_best = 0xFFFFFFFFu; }
        __syncthreads();
        {   // stage 1: 8 consecutive bit positions per thread out of one 128-bit window
            const uint8_t* a = in + (tile >> 3) + t;
            const uint64_t* al = (const uint64_t*)((uintptr_t)a & ~(uintptr_t)7);
            const int sh = (int)((uintptr_t)a & 7) * 8;
            const uint64_t x0 = __ldg(al), x1 = __ldg(al + 1), x2 = __ldg(al + 2);
            const uint64_t lo = funnel128(x0, x1, sh), hi = funnel128(x1, x2, sh);
#pragma unroll
            for (int s8 = 0; s8 < 8; s8++) {
                const uint64_t p = tile + (uint64_t)t * 8 + s8;
                if (p >= seg_end || p + 17 + 12 > total_bits) continue;
                const uint64_t w1 = funnel128(lo, hi, s8);
                if (((w1 >> 1) & 3) != 2) continue;
                if (((w1 >> 3) & 31) > 29 || ((w1 >> 8) & 31) > 29) continue;
                const int ncl = (int)((w1 >> 13) & 15) + 4;
                const uint64_t cw = funnel128(lo, hi, s8 + 17);
                int kraft = 0;
                for (int i = 0; i < ncl; i++) {
                    const int v = (int)((cw >> (3 * i)) & 7);
                    kraft += v ? (128 >> v) : 0;
                }
                if (kraft != 128) continue;
                s_q[atomicAdd(&s_n, 1u)] = (uint16_t)(t * 8 + s8);
            }
        }
        __syncthreads();
        const unsigned cnt = s_n;
        // stage 2: queue entry e -> lane e / 8 of warp e % 8, so few survivors land in different warps
        for (unsigned e = (unsigned)(t & 31) * 8 + (unsigned)(t >> 5); e < cnt; e += FIND_NT) {
            const unsigned rel = s_q[e];
            if (rel < s_best && find_full_check(in, tile + rel, total_bits)) atomicMin(&s_best, rel);
        }
        __syncthreads();
        if (s_best != 0xFFFFFFFFu) found = tile + s_best;
        __syncthreads();
    }
    if (t == 0) sd.start_bit = found;
}

// --------------------------------------------------------------------------------------------------
// k_count: one CTA per stream.
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PARSE_NT, 1)
k_count(const uint8_t* __restrict__ d_in, const StreamDesc* __restrict__ descs, StreamInfo* __restrict__ infos,
        BlockRec* __restrict__ blocks, ChunkRec* __restrict__ chunks, const uint32_t* __restrict__ list,
        uint64_t seg_bits, uint64_t spec_max_bits) {
    const int sid = (int)list[blockIdx.x], t = threadIdx.x;
    const StreamDesc sd = descs[sid];
    const uint8_t* in = d_in + sd.in_off;
    const uint64_t total_bits = sd.in_len * 8;
    if (sd.start_bit == BIT_NONE) {  // speculative walker without a candidate header in its segment
        if (t == 0) {
            StreamInfo si;
            si.status = ST_NONE; si.n_blocks = 0; si.n_syms = 0; si.out_len = 0; si.consumed = 0; si.n_chunks = 0;
            si.total_bits = 0; si.end_bit = BIT_NONE; si.final_seen = 0; si.pad = 0;
            infos[sid] = si;
        }
        return;
    }

    __shared__ DecTab s_lit, s_dst, s_fixlit, s_fixdst;
    __shared__ DecLut s_lut, s_fixlut;
    __shared__ BlockRec s_blk;
    __shared__ uint32_t s_start[PARSE_NT], s_end[PARSE_NT], s_cnt[PARSE_NT], s_outb[PARSE_NT];
    __shared__ uint8_t s_flag[PARSE_NT];
    __shared__ uint32_t s_scan_cnt[PARSE_NT / 32], s_scan_out[PARSE_NT / 32];
    __shared__ int s_status, s_type, s_final, s_stop;
    __shared__ uint64_t s_pos;
    extern __shared__ __align__(16) uint8_t s_win[];   // WIN_BYTES + WIN_SLACK + 16: the window being decoded

    if (t == 0) {
        // the fixed litlen CODES are those of the 288-entry RFC table (HuffmanTable.java:183-186 skips two
        // code values), although the reference's table object has 286 entries
        uint8_t L[MAX_LL], D[MAX_D];
        fixed_lens(L, D);
        L[286] = L[287] = 8;
        build_dectab(L, 288, s_fixlit);
        build_dectab(D, 30, s_fixdst);
        s_status = ST_OK;
        s_pos = sd.start_bit;
    }
    __syncthreads();
    build_lut(s_fixlit, s_fixlut.lit, LUT_L_BITS, t, PARSE_NT);
    build_lut(s_fixdst, s_fixlut.dst, LUT_D_BITS, t, PARSE_NT);
    __syncthreads();

    uint64_t n_syms = 0, out_total = 0, n_chunks = 0;
    uint32_t n_blocks = 0;
    bool done = false;

    while (!done) {
        // ---- block header (thread 0), from a staged copy of its bytes: a dynamic header is a few hundred dependent
        //      reads, far too slow from global memory --------------------------------------------------------------
        const uint64_t hbyte = (s_pos >> 3) & ~15ull;
        {
            const int64_t valid = (int64_t)(sd.in_len + 32) - (int64_t)hbyte;
            for (int o = t * 16; o < HDR_STAGE; o += PARSE_NT * 16) {
                const int64_t left = valid - o;
                const int nb = left >= 16 ? 16 : left > 0 ? (int)left : 0;
                cp_async16(s_win + o, in + hbyte + (nb ? o : 0), nb);
            }
            cp_async_wait_all();
            __syncthreads();
        }
        if (t == 0) {
            auto peek = [&](uint64_t p) { return peek_bits_smem(s_win, (uint32_t)(p - hbyte * 8)); };
            uint64_t pos = s_pos;
            s_blk.hdr_bit = pos;
            s_blk.hdr.np = 0; s_blk.hdr.ncl = 0; s_blk.hdr.bits = 0;
            s_blk.tab.nL = 0; s_blk.tab.nD = 0;
            if (pos + 3 > total_bits) { s_status = ST_PARSE; }
            else {
                uint64_t w = peek(pos);
                s_final = (int)(w & 1);
                int type = (int)((w >> 1) & 3);
                s_type = type;
                pos += 3;
                if (type == 3) s_status = ST_PARSE;
                else if (type == 0) {  // DeflateBlockUncompressed.parse (:23-36)
                    pos = (pos + 7) & ~7ull;
                    if (pos + 32 > total_bits) s_status = ST_PARSE;
                    else {
                        uint64_t v = peek(pos);
                        uint32_t len = (uint32_t)(v & 0xffff), nlen = (uint32_t)((v >> 16) & 0xffff);
                        pos += 32;
                        if (nlen != (~len & 0xffff)) s_status = ST_PARSE;
                        // truncated stored data: the reference pads with 0xFF; we refuse
                        else if (pos + 8ull * len > total_bits) s_status = ST_UNSUPPORTED;
                        else {
                            s_blk.data_bit = pos;
                            s_blk.out_len = len;
                            s_blk.n_sym = 0;
                            s_blk.payload_bits = 0;
                            s_blk.tab.type = 0;
                            pos += 8ull * len;
                            s_blk.end_bit = pos;
                        }
                    }
                } else if (type == 1) {
                    s_blk.tab.type = 1;
                    s_blk.tab.nL = 286; s_blk.tab.nD = 30;
                    fixed_lens(s_blk.tab.L, s_blk.tab.D);
                    for (int k = 286; k < MAX_LL; k++) s_blk.tab.L[k] = 0;
                    for (int k = 30; k < MAX_D; k++) s_blk.tab.D[k] = 0;
                    s_blk.data_bit = pos;
                } else {
                    uint32_t used = parse_dynamic_header(peek, pos, total_bits, s_blk, s_lit, s_win + HDR_STAGE);
                    if (used == 0) s_status = ST_PARSE;
                    else {
                        pos += used;
                        s_blk.data_bit = pos;
                        build_dectab(s_blk.tab.L, s_blk.tab.nL, s_lit);
                        build_dectab(s_blk.tab.D, s_blk.tab.nD, s_dst);
                    }
                }
            }
            s_blk.type = (uint8_t)s_type;
            s_blk.bfinal = (uint8_t)s_final;
            s_pos = pos;
        }
        __syncthreads();
        if (s_status != ST_OK) break;
        const int type = s_type;
        uint32_t blk_syms = 0, blk_out = 0, blk_chunks = 0;
        const uint64_t chunk_base = n_chunks;

        if (type != 0) {
            const DecTab& lit = (type == 1) ? s_fixlit : s_lit;
            const DecTab& dst = (type == 1) ? s_fixdst : s_dst;
            if (type == 2) {   // the block's own lookup tables
                build_lut(s_lit, s_lut.lit, LUT_L_BITS, t, PARSE_NT);
                build_lut(s_dst, s_lut.dst, LUT_D_BITS, t, PARSE_NT);
                __syncthreads();
            }
            const DecLut& lut = (type == 1) ? s_fixlut : s_lut;
            const uint64_t data_bit = s_blk.data_bit;
            uint64_t win = data_bit;  // true symbol boundary
            int64_t payload = 0;
            bool eob_seen = false;
            while (!eob_seen) {
                // ---- the window's bytes -> shared memory (cp.async), 16-byte aligned source --------------
                const uint64_t wbyte = (win >> 3) & ~15ull;
                {
                    const int64_t valid = (int64_t)(sd.in_len + 32) - (int64_t)wbyte;   // bytes of the stream (+ its zero padding) from here
                    for (int o = t * 16; o < WIN_BYTES + WIN_SLACK + 16; o += PARSE_NT * 16) {
                        const int64_t left = valid - o;
                        const int nb = left >= 16 ? 16 : left > 0 ? (int)left : 0;
                        cp_async16(s_win + o, in + wbyte + (nb ? o : 0), nb);
                    }
                    cp_async_wait_all();
                    __syncthreads();
                }
                const uint32_t wrel = (uint32_t)(win - wbyte * 8);   // bit offset of the window start inside the staged bytes
                // ---- speculative decode of one window ----------------------------------------------
                uint32_t start = (uint32_t)t * CHUNK_BITS;
                const uint32_t limit = (uint32_t)(t + 1) * CHUNK_BITS;
                bool need = true, changed;
                uint32_t end = 0, cnt = 0, outb = 0;
                int flag = 0;
                do {
                    if (need) {
                        uint32_t p = start;
                        cnt = 0; outb = 0; flag = 0;
                        while (p < limit) {
                            uint64_t abs = win + p;
                            if (abs >= total_bits) { flag = 2; break; }
                            Unit u = decode_unit_w(lit, dst, lut, peek_bits_smem(s_win, wrel + p));
                            if (u.nbits == 0 || abs + u.nbits > total_bits) { flag = 2; break; }
                            p += u.nbits; cnt++; outb += u.outlen;
                            if (u.eob) { flag = 1; break; }
                        }
                        end = p;
                        s_end[t] = end; s_flag[t] = (uint8_t)flag;
                    }
                    __syncthreads();
                    changed = false;
                    need = false;
                    if (t > 0 && s_flag[t - 1] == 0) {
                        uint32_t pe = s_end[t - 1];
                        if (pe != start) { start = pe; need = true; changed = true; }
                    }
                    changed = __syncthreads_or(changed);
                } while (changed);
                // ---- first stopping chunk on the valid chain ----------------------------------------
                if (t == 0) s_stop = PARSE_NT;
                __syncthreads();
                if (flag != 0) atomicMin(&s_stop, t);
                __syncthreads();
                const int stop = s_stop;  // chunks 0..min(stop, NT-1) are valid
                const bool valid = t <= stop;
                if (stop < PARSE_NT && s_flag[stop] == 2) { if (t == 0) s_status = ST_PARSE; }
                // ---- exclusive scans of symbol counts / decoded bytes over the valid chunks ----------
                uint32_t c = valid ? cnt : 0, o = valid ? outb : 0;
                uint32_t ci = c, oi = o;
                const int lane = t & 31, wid = t >> 5;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t a = __shfl_up_sync(0xffffffffu, ci, d), b2 = __shfl_up_sync(0xffffffffu, oi, d);
                    if (lane >= d) { ci += a; oi += b2; }
                }
                if (lane == 31) { s_scan_cnt[wid] = ci; s_scan_out[wid] = oi; }
                __syncthreads();
                if (wid == 0) {
                    uint32_t a = s_scan_cnt[lane], b2 = s_scan_out[lane];
                    uint32_t ai = a, bi = b2;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t x = __shfl_up_sync(0xffffffffu, ai, d), y = __shfl_up_sync(0xffffffffu, bi, d);
                        if (lane >= d) { ai += x; bi += y; }
                    }
                    s_scan_cnt[lane] = ai - a; s_scan_out[lane] = bi - b2;  // exclusive warp bases
                    if (lane == 31) { s_cnt[0] = ai; s_outb[0] = bi; }      // window totals
                }
                __syncthreads();
                const uint32_t cbase = s_scan_cnt[wid] + ci - c, obase = s_scan_out[wid] + oi - o;
                const uint32_t wcnt = s_cnt[0], wout = s_outb[0];
                if (valid) {
                    uint64_t slot = n_chunks + (uint64_t)t;
                    if (slot < sd.chunk_cap) {
                        ChunkRec r;
                        r.start_rel = (uint32_t)(win - data_bit) + start;
                        r.sym_idx = blk_syms + cbase;
                        r.out_off = blk_out + obase;
                        r.n = cnt;
                        chunks[sd.chunk_base + slot] = r;
                    }
                }
                const int nvalid = (stop < PARSE_NT ? stop : PARSE_NT - 1) + 1;
                // payload bits of this window = end of the last valid chunk
                if (t == nvalid - 1) s_start[0] = end;
                __syncthreads();
                const uint32_t wend = s_start[0];
                payload += wend;
                win += wend;
                n_chunks += nvalid; blk_chunks += nvalid;
                blk_syms += wcnt; blk_out += wout;
                if (stop < PARSE_NT) eob_seen = true;  // EOB (or error, handled by status)
                if (win - data_bit > 0xFFFF0000ull || blk_syms > 0x7FFF0000u) { if (t == 0) s_status = ST_UNSUPPORTED; eob_seen = true; }
                // a speculative walker that runs far past its segment is decoding garbage (or a block too long to
                // be worth guessing): give up, the chain re-walks it from a known boundary if it is needed
                if (sd.spec && !eob_seen && win > sd.stop_bit + spec_max_bits) { if (t == 0) s_status = ST_ABORT; eob_seen = true; }
                __syncthreads();
                if (s_status != ST_OK) break;
            }
            if (s_status != ST_OK) break;
            if (t == 0) {
                s_blk.end_bit = win;
                s_blk.n_sym = blk_syms;
                s_blk.out_len = blk_out;
                s_blk.payload_bits = payload;
                s_pos = win;
            }
        }
        // ---- record the block ------------------------------------------------------------------------
        if (t == 0) {
            s_blk.sym_base = n_syms;
            s_blk.out_base = out_total;
            s_blk.chunk_base = (uint32_t)chunk_base;
            s_blk.n_chunks = blk_chunks;
        }
        __syncthreads();
        if (n_blocks < sd.blk_cap) {
            // cooperative copy of the record
            const uint32_t* src = (const uint32_t*)&s_blk;
            uint32_t* dstp = (uint32_t*)&blocks[sd.blk_base + n_blocks];
            for (int k = t; k < (int)(sizeof(BlockRec) / 4); k += PARSE_NT) dstp[k] = src[k];
        }
        n_syms += (type != 0) ? s_blk.n_sym : 0;
        out_total += s_blk.out_len;
        n_blocks++;
        done = s_final != 0;
        if (!done && s_pos >= sd.stop_bit) {
            // block boundary at or past the segment end: hand over to the walker of the segment the boundary is in.
            // A walker on a known boundary keeps going when that walker did not start here (it would have to be
            // re-walked anyway), as long as its record slots last.
            done = true;
            if (!sd.spec) {
                uint64_t j = s_pos / seg_bits;
                if (j >= sd.nseg) j = sd.nseg - 1;
                const uint64_t nxt = descs[sd.walker0 + j].start_bit;
                if (nxt != s_pos && (uint64_t)n_blocks * 2 < sd.blk_cap && n_chunks * 2 < sd.chunk_cap) done = false;
            }
        }
        if (out_total > 0xF0000000ull) { if (t == 0) s_status = ST_UNSUPPORTED; done = true; }
        __syncthreads();
    }
    if (t == 0) {
        StreamInfo si;
        si.status = s_status;
        si.n_blocks = n_blocks;
        si.n_syms = n_syms;
        si.out_len = out_total;
        si.consumed = (s_pos + 7) >> 3;
        si.n_chunks = n_chunks;
        si.total_bits = s_pos;
        si.end_bit = s_pos;
        si.final_seen = (s_status == ST_OK && s_final != 0) ? 1u : 0u;
        si.pad = 0;
        infos[sid] = si;
    }
}

// --------------------------------------------------------------------------------------------------
// k_emit: one CTA per block.  Huffman blocks: thread per chunk re-decodes and writes symbols.
// Stored blocks: copy bytes into `out` (DeflateBlockUncompressed.parse :23-36).
// --------------------------------------------------------------------------------------------------
struct EmitJob { uint32_t stream, block; };

__global__ void __launch_bounds__(256)
k_emit(const uint8_t* __restrict__ d_in, const StreamDesc* __restrict__ descs, const BlockRec* __restrict__ blocks,
       const ChunkRec* __restrict__ chunks, const EmitJob* __restrict__ jobs, StreamInfo* __restrict__ infos,
       uint32_t* __restrict__ sym, uint32_t* __restrict__ symout, uint8_t* __restrict__ out) {
    const EmitJob job = jobs[blockIdx.x];
    const StreamDesc sd = descs[job.stream];
    const BlockRec& b = blocks[sd.blk_base + job.block];
    const uint8_t* in = d_in + sd.in_off;
    const int t = threadIdx.x;
    if (b.type == 0) {
        const uint8_t* src = in + (b.data_bit >> 3);
        uint8_t* dst = out + sd.out_base + b.out_base;
        for (uint32_t k = t; k < b.out_len; k += blockDim.x) dst[k] = src[k];
        return;
    }
    __shared__ DecTab s_lit, s_dst;
    __shared__ DecLut s_lut;
    if (t == 0) {
        if (b.type == 1) {
            uint8_t L[MAX_LL], D[MAX_D];
            fixed_lens(L, D);
            L[286] = L[287] = 8;
            build_dectab(L, 288, s_lit);
        } else {
            build_dectab(b.tab.L, b.tab.nL, s_lit);
        }
    }
    if (t == 32) build_dectab(b.tab.D, b.tab.nD, s_dst);
    __syncthreads();
    build_lut(s_lit, s_lut.lit, LUT_L_BITS, t, (int)blockDim.x);
    build_lut(s_dst, s_lut.dst, LUT_D_BITS, t, (int)blockDim.x);
    __syncthreads();
    const uint64_t symb = sd.sym_base + b.sym_base, outb = sd.out_base + b.out_base;
    for (uint32_t c = t; c < b.n_chunks; c += blockDim.x) {
        const ChunkRec r = chunks[sd.chunk_base + b.chunk_base + c];
        uint64_t pos = b.data_bit + r.start_rel;
        uint64_t si = symb + r.sym_idx;
        uint32_t oo = (uint32_t)(outb + r.out_off);
        for (uint32_t k = 0; k < r.n; k++) {
            Unit u = decode_unit(s_lit, s_dst, s_lut, in, pos);
            pos += u.nbits;
            sym[si + k] = u.packed;
            symout[si + k] = oo;
            // a match reaching before the start of its stream: the reference dereferences a null
            // prevBlock (DeflateBlock.java:175-181); reported as a parse failure
            if (sym_is_match(u.packed) && (uint64_t)oo - sd.stream_out_base < (uint64_t)sym_dist(u.packed))
                infos[job.stream].status = ST_PARSE;
            oo += u.outlen;
        }
    }
}

// --------------------------------------------------------------------------------------------------
// LZ77 resolution (replaces DeflateBlock.readSlice, DeflateBlock.java:147-222)
//   k_lz_fill : per symbol, literal -> out byte + root pointer; match -> every byte points at
//               (position - distance), folded into the non-overlapping source for overlapped copies.
//   k_lz_jump : pointer doubling; a byte that reaches a root copies its value and becomes a root itself.
// A match reaching before the start of its stream is a parse failure (the reference dereferences a
// null prevBlock there).
// --------------------------------------------------------------------------------------------------
__global__ void k_lz_fill(const uint32_t* __restrict__ sym, const uint32_t* __restrict__ symout, uint64_t n_sym,
                          uint8_t* __restrict__ out, uint32_t* __restrict__ ptr) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sym) return;
    uint32_t s = sym[i];
    uint32_t o = symout[i];
    if (!sym_is_match(s)) {
        if (s < 256) { out[o] = (uint8_t)s; ptr[o] = o; }
        return;
    }
    const int len = sym_len(s), dist = sym_dist(s);
    if (o < (uint32_t)dist) {  // only reachable in a stream already flagged by k_emit; keep pointers in bounds
        for (int k = 0; k < len; k++) ptr[o + k] = o + k;
        return;
    }
    const uint32_t src = o - dist;
    for (int k = 0; k < len; k++) ptr[o + k] = src + (dist < len ? (uint32_t)(k % dist) : (uint32_t)k);
}

// stored blocks: their bytes were copied by k_emit and are roots
__global__ void k_lz_root_stored(const StreamDesc* __restrict__ descs, const BlockRec* __restrict__ blocks,
                                 const EmitJob* __restrict__ jobs, uint32_t njobs, uint32_t* __restrict__ ptr) {
    if (blockIdx.x >= njobs) return;
    const EmitJob job = jobs[blockIdx.x];
    const StreamDesc sd = descs[job.stream];
    const BlockRec& b = blocks[sd.blk_base + job.block];
    if (b.type != 0) return;
    const uint32_t o = (uint32_t)(sd.out_base + b.out_base);
    for (uint32_t k = threadIdx.x; k < b.out_len; k += blockDim.x) ptr[o + k] = o + k;
}

// gather the per-stream BlockRec slots into one gap-free array in stream order (warp per record)
__global__ void k_compact_blocks(const StreamDesc* __restrict__ descs, const BlockRec* __restrict__ blocks,
                                 const EmitJob* __restrict__ jobs, uint64_t n, BlockRec* __restrict__ dst) {
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    const EmitJob job = jobs[w];
    const StreamDesc& sd = descs[job.stream];
    const uint32_t* s = (const uint32_t*)&blocks[sd.blk_base + job.block];
    uint32_t* d = (uint32_t*)&dst[w];
    for (int k = lane; k < (int)(sizeof(BlockRec) / 4); k += 32) d[k] = s[k];
    __syncwarp();
    if (lane == 0) { dst[w].sym_base += sd.sym_base; dst[w].out_base += sd.out_base; }  // walker relative -> pool index
}

// First pass, in output order inside a tile: a CTA walks its tile chunk by chunk, so by the time a byte is looked at,
// whatever it copies from further back in the same tile is already resolved (`done`, value in `out`) and costs one hop;
// only sources in the same chunk or in another CTA's tile are followed link by link (at most 8, as in k_lz_jump).
// A `done` flag of ANOTHER tile is never trusted here (its value may not be visible yet); the pointer chain is.
constexpr int LZJ_NT = 512, LZJ_TILE = 32768;
__global__ void __launch_bounds__(LZJ_NT)
k_lz_jump_tiles(uint32_t* ptr, uint8_t* out, uint8_t* done, uint64_t n, uint32_t tile, int* __restrict__ changed) {
    const uint64_t t0 = (uint64_t)blockIdx.x * tile;
    const uint64_t t1 = t0 + tile < n ? t0 + tile : n;
    bool ch = false;
    uint32_t nextp = t0 + threadIdx.x < t1 ? ptr[t0 + threadIdx.x] : 0u;
#pragma unroll 1
    for (uint64_t c = t0; c < t1; c += LZJ_NT) {
        const uint64_t i = c + threadIdx.x;
        // the byte's own pointer does not depend on the chunks before it: the next chunk's load is in flight while this
        // one follows its links
        const uint32_t mine = nextp;
        if (i + LZJ_NT < t1) nextp = ptr[i + LZJ_NT];
        if (i < t1) {
            uint32_t p = mine;
            bool root = p == (uint32_t)i;
            if (!root) {
                const uint32_t p0 = p;
                uint32_t to = p;
#pragma unroll 1
                for (int h = 0; h < 8; h++) {
                    // (loading ptr / out / done of p together, one round trip per hop, was measured slower: the pass is
                    //  bound by the number of scattered accesses, not by their latency)
                    if (p >= t0 && p < c && done[p]) { root = true; to = ptr[p]; break; }   // resolved earlier by this CTA
                    const uint32_t q = ptr[p];
                    if (q == p) { root = true; to = p; break; }
                    p = q;
                    to = p;
                }
                if (root) out[i] = out[p];
                if (to != p0) ptr[i] = to;
                ch |= !root;
            }
            if (root) done[i] = 1;
        }
        __syncthreads();
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) *changed = 1;
}

// One pass: follow up to 8 links.  Pointers only ever move towards the root and a root's pointer and value never
// change during the passes, so plain (cached, possibly stale) loads are safe: whatever is read on the way is an ancestor.
// A byte that reaches a root takes the root's value, points at that root and sets its `done` flag: later passes skip it
// after a one-byte read, and no gather pass is needed at the end.  (It does not become a root itself: a walker running
// in the same pass could then read its value before it is written.)
__global__ void k_lz_jump(uint32_t* ptr, uint8_t* out, uint8_t* done, uint64_t n, int* __restrict__ changed) {
    // four bytes per thread: one 32-bit read of their flags decides whether there is anything to do
    const uint64_t i4 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    bool ch = false;
    if (i4 < n) {
        const uint32_t flags = i4 + 4 <= n ? *(const uint32_t*)(done + i4) : 0u;
        if (flags != 0x01010101u) {
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                const uint64_t i = i4 + b;
                if (i >= n || done[i]) continue;
                uint32_t p = ptr[i];
                bool root = p == (uint32_t)i;
                if (!root) {
                    const uint32_t p0 = p;
#pragma unroll 1
                    for (int h = 0; h < 8; h++) {
                        const uint32_t q = ptr[p];
                        if (q == p) { root = true; break; }
                        p = q;
                    }
                    if (root) out[i] = out[p];
                    if (p != p0) ptr[i] = p;
                    ch |= !root;
                }
                if (root) done[i] = 1;
            }
        }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) *changed = 1;
}

}  // namespace d4
