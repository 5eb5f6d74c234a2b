// write.cuh — scan-and-scatter bitstream writer (DeflateStream.write, DeflateStream.java:128-145;
// DeflateBlockHuffman.write*, DeflateBlockHuffman.java:1033-1156; DeflateBlockUncompressed.write,
// DeflateBlockUncompressed.java:39-56).
//
// One CTA per block of the final model.  Per tile of symbols: every thread sizes its symbol, a CTA-wide
// exclusive scan gives each symbol its bit offset, then the code bits are OR-ed into the (zeroed) output
// with 64-bit atomics — neighbouring symbols share words, nothing else does.  The reference writes one
// bit per loop iteration (BitOutputStream.java:41-45).
#pragma once
#include "optimise.cuh"

namespace d4 {

constexpr int WR_NT = 256;

__device__ __forceinline__ void put_bits(unsigned long long* outw, uint64_t bitpos, unsigned long long bits, int n) {
    if (n == 0) return;
    const uint64_t w = bitpos >> 6;
    const int sh = (int)(bitpos & 63);
    atomicOr(&outw[w], bits << sh);
    if (sh + n > 64) atomicOr(&outw[w + 1], bits >> (64 - sh));
}

__device__ __forceinline__ uint32_t rev_bits(uint32_t code, int len) { return len ? (__brev(code) >> (32 - len)) : 0; }

// canonical codes of a length set (Huffman.buildCodes, Huffman.java:35-64), already bit-reversed for
// LSB-first emission (Huffman.getSym -> Util.rev, Util.java:244-254)
__device__ inline void build_codes(const uint8_t* lens, int n, uint16_t* codes) {
    int count[16];
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int i = 0; i < n; i++) count[lens[i] & 15]++;
    uint32_t next[16];
    uint32_t nc = 0;
    int lastShift = 0;
    for (int l = 1; l < 16; l++) {
        next[l] = 0;
        if (count[l] == 0) continue;
        nc <<= (l - lastShift);
        lastShift = l;
        next[l] = nc;
        nc += count[l];
    }
    for (int i = 0; i < n; i++) {
        int l = lens[i];
        codes[i] = l ? (uint16_t)rev_bits(next[l]++, l) : 0;
    }
}

struct WriteSmem {
    uint16_t codeL[MAX_LL], codeD[MAX_D], codeCL[MAX_CL];
    uint8_t L[MAX_LL], D[MAX_D];
    uint32_t warpsum[WR_NT / 32];
    uint64_t base;
};

__global__ void __launch_bounds__(WR_NT)
k_write(const StreamState* __restrict__ streams, const uint32_t* __restrict__ blk_stream, const BlkState* __restrict__ bs,
        const uint32_t* __restrict__ sym, const uint32_t* __restrict__ symout, const uint8_t* __restrict__ out,
        const uint32_t* __restrict__ maskpool, const uint64_t* __restrict__ dst_off, unsigned long long* __restrict__ dstw,
        int* __restrict__ gerr) {
    const BlkState& b = bs[blockIdx.x];
    const uint32_t sid = blk_stream[blockIdx.x];
    if (!b.alive || streams[sid].status != ST_OK) return;
    const int tid = threadIdx.x;
    unsigned long long* dw = dstw + (dst_off[sid] >> 3);  // stream output base (8-byte aligned)
    const uint64_t pos0 = b.bit_pos;
    const int type = b.cand.tab.type;

    if (type == 0) {
        const uint64_t dataByte = ((pos0 + 3 + 7) >> 3) + 4;
        const uint8_t* src = out + b.out_off;
        if (tid == 0) {
            put_bits(dw, pos0, (unsigned long long)(b.bfinal ? 1 : 0), 3);
            uint32_t len = (uint32_t)b.out_len & 0xffff;
            uint32_t hdr = len | ((~len & 0xffff) << 16);
            put_bits(dw, (dataByte - 4) * 8, hdr, 32);
        }
        // data bytes: 8 at a time (byte-aligned destination; edges and interior alike go through OR)
        const uint64_t n = b.out_len;
        for (uint64_t k = (uint64_t)tid * 8; k < n; k += (uint64_t)WR_NT * 8) {
            unsigned long long v = 0;
            int cnt = (int)((n - k) < 8 ? (n - k) : 8);
            for (int j = 0; j < cnt; j++) v |= (unsigned long long)src[k + j] << (8 * j);
            put_bits(dw, (dataByte + k) * 8, v, cnt * 8);
        }
        return;
    }

    __shared__ WriteSmem W;
    for (int k = tid; k < MAX_LL; k += WR_NT) W.L[k] = b.cand.tab.L[k];
    if (tid < MAX_D) W.D[tid] = b.cand.tab.D[tid];
    __syncthreads();
    // FIXED: codes of the 288-entry RFC table (see parse.cuh); lengths of 286/287 are only used for codes
    if (tid == 0) {
        if (type == 1) { W.L[286] = 8; W.L[287] = 8; }
        build_codes(W.L, type == 1 ? 288 : b.cand.tab.nL, W.codeL);
    }
    if (tid == 32) build_codes(W.D, b.cand.tab.nD, W.codeD);
    if (tid == 64 && type == 2) build_codes(b.cand.hdr.CL, 19, W.codeCL);
    __syncthreads();

    // ---- block header (writeProlog / writeHuffCode), serial ------------------------------------------
    if (tid == 0) {
        uint64_t p = pos0;
        put_bits(dw, p, (unsigned long long)((type << 1) | (b.bfinal ? 1 : 0)), 3);
        p += 3;
        if (type == 2) {
            const Hdr& h = b.cand.hdr;
            put_bits(dw, p, (unsigned long long)((b.cand.tab.nL - 257) & 31), 5); p += 5;
            put_bits(dw, p, (unsigned long long)((b.cand.tab.nD - 1) & 31), 5); p += 5;
            put_bits(dw, p, (unsigned long long)((h.ncl - 4) & 15), 4); p += 4;
            for (int i = 0; i < h.ncl; i++) { put_bits(dw, p, h.CL[c_codelen_order[i]], 3); p += 3; }
            for (int i = 0; i < h.np; i++) {
                uint16_t pr = h.pairs[i];
                int s = pair_sym(pr), run = pair_run(pr);
                put_bits(dw, p, W.codeCL[s], h.CL[s]); p += h.CL[s];
                if (run > 0) {
                    int off = s == 18 ? 11 : 3, sz = pair_extra_bits(s);
                    put_bits(dw, p, (unsigned long long)(run - off), sz); p += sz;
                }
            }
            if ((long long)(p - pos0 - 3) != (long long)h.bits) {  // model/writer disagree
                if (atomicMax(gerr, 2) == 0) { gerr[1] = (int)blockIdx.x; gerr[2] = (int)(p - pos0 - 3); gerr[3] = h.bits; gerr[4] = 1; }
            }
        }
        W.base = p;
    }
    __syncthreads();

    // ---- symbols (writeDefBlock) -----------------------------------------------------------------------
    const uint32_t* S = sym + b.sym_off;
    const uint32_t* SO = symout + b.sym_off;
    const uint32_t* M = maskpool + b.mask_off;
    const int lane = tid & 31, wid = tid >> 5;
    for (uint32_t base = 0; base < b.n_sym; base += WR_NT) {
        const uint32_t i = base + tid;
        uint32_t s = 0;
        int nbits = 0;
        bool replaced = false;
        if (i < b.n_sym) {
            s = S[i];
            if (!sym_is_match(s)) {
                nbits = (s <= 256) ? W.L[s] : 0;  // NOP writes nothing
            } else {
                replaced = (M[i >> 5] >> (i & 31)) & 1;
                if (!replaced) {
                    int ls = sym_lensym(s), ds = dist_sym(sym_dist(s));
                    nbits = W.L[ls] + len_ebits_of(ls) + W.D[ds] + dist_ebits_of(ds);
                } else {
                    const uint8_t* p = out + SO[i];
                    int len = sym_len(s);
                    for (int k = 0; k < len; k++) nbits += W.L[p[k]];
                }
            }
        }
        // exclusive scan over the tile
        uint32_t incl = (uint32_t)nbits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t a = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += a;
        }
        if (lane == 31) W.warpsum[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (int k = 0; k < WR_NT / 32; k++) { uint32_t x = W.warpsum[k]; if (k < wid) wbase += x; total += x; }
        uint64_t p = W.base + wbase + incl - (uint32_t)nbits;
        if (i < b.n_sym && nbits) {
            if (!sym_is_match(s)) {
                put_bits(dw, p, W.codeL[s], nbits);
            } else if (!replaced) {  // writeBackref (:1110-1130)
                int ls = sym_lensym(s), dist = sym_dist(s), ds = dist_sym(dist), len = sym_len(s);
                unsigned long long bits = W.codeL[ls];
                int nb = W.L[ls];
                bits |= (unsigned long long)(len - c_len_base[ls - 257]) << nb;
                nb += len_ebits_of(ls);
                bits |= (unsigned long long)W.codeD[ds] << nb;
                nb += W.D[ds];
                bits |= (unsigned long long)(dist - c_dist_base[ds]) << nb;
                nb += dist_ebits_of(ds);
                put_bits(dw, p, bits, nb);
            } else {  // the match's bytes as literals
                const uint8_t* q = out + SO[i];
                int len = sym_len(s);
                unsigned long long acc = 0;
                int na = 0;
                for (int k = 0; k < len; k++) {
                    int v = q[k], l = W.L[v];
                    if (na + l > 64) { put_bits(dw, p, acc, na); p += na; acc = 0; na = 0; }
                    acc |= (unsigned long long)W.codeL[v] << na;
                    na += l;
                }
                put_bits(dw, p, acc, na);
            }
        }
        __syncthreads();
        if (tid == 0) W.base += total;
        __syncthreads();
    }
    if (tid == 0) {
        if ((long long)(W.base - pos0 - 3) != b.size_bits) {
            if (atomicMax(gerr, 2) == 0) { gerr[1] = (int)blockIdx.x; gerr[2] = (int)(W.base - pos0 - 3); gerr[3] = (int)b.size_bits; gerr[4] = 2; }
        }
    }
}

}  // namespace d4
