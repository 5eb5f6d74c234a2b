// common.cuh — shared definitions for the deft4cu kernels (sm_100a).
//
// Data model in HBM (per batch of streams; "pool" arrays are batch-wide):
//   in      : all input streams back to back, each 16-byte aligned and followed by >= 16 zero bytes
//   sym     : u32 per deflate symbol (literal / EOB / match), in stream order          (SymPool)
//   symout  : u32 per symbol = offset of the symbol's first decoded byte in `out`
//   out     : decoded bytes of all streams back to back
//   BlockRec: one record per deflate block as parsed (tables, header pairs, counters)
//   BlkState: one record per block of the *current* model (after optimisation rounds)
//
// The reference's LitLen list (deft4j-base .../deflate/LitLen.java:29-47) becomes sym/symout plus a
// per-candidate bit mask "this match has been replaced by its literals": the optimiser only ever turns
// matches into literals (DeflateBlockHuffman.java:222-296,373-458), so (mask, code lengths) identify a
// candidate's symbol list exactly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// D4_HOST_TEST builds the single-thread building blocks (huff.cuh) as host code so that CPU-only tests
// can pin the very source the kernels run against the oracle (tests/test_host_units.py).
#ifdef D4_HOST_TEST
#define D4_DEV inline
#define D4_DEV_BIG inline
#define D4_CONST static const
#else
#define D4_DEV __device__ inline
#define D4_DEV_BIG __device__ __noinline__   // one copy of the big serial routines: the engine kernel is i-cache bound
#define D4_CONST __constant__
#endif

namespace d4 {

// ---- packed symbol ------------------------------------------------------------------------------
// bit 31 = 1: match   bits 0-8 len-3 | bits 9-23 dist-1 | bit 24 edgecase (len 258 coded as 284+31,
//                     DeflateBlockHuffman.java:840-843) | bits 25-29 length symbol - 257
// bit 31 = 0: bits 0-8 literal value (0..255) or 256 = EOB; 0x1FF = NOP (EOB dropped by a merge,
//                     DeflateBlockHuffman.java:1257)
constexpr uint32_t SYM_MATCH = 0x80000000u;
constexpr uint32_t SYM_NOP = 0x1FFu;
__host__ __device__ inline bool sym_is_match(uint32_t s) { return (s & SYM_MATCH) != 0; }
__host__ __device__ inline int sym_len(uint32_t s) { return (int)(s & 0x1FF) + 3; }
__host__ __device__ inline int sym_dist(uint32_t s) { return (int)((s >> 9) & 0x7FFF) + 1; }
__host__ __device__ inline int sym_edge(uint32_t s) { return (int)((s >> 24) & 1); }
__host__ __device__ inline int sym_lensym(uint32_t s) { return (int)((s >> 25) & 31) + 257; }
__host__ __device__ inline uint32_t sym_pack_match(int len, int dist, int edge, int lensym) {
    return SYM_MATCH | (uint32_t)(len - 3) | ((uint32_t)(dist - 1) << 9) | ((uint32_t)edge << 24) |
           ((uint32_t)(lensym - 257) << 25);
}

// ---- RFC 1951 tables (Constants.java:63-128) ------------------------------------------------------
D4_CONST uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                        67, 83, 99, 115, 131, 163, 195, 227, 258};
D4_CONST uint8_t c_len_ebits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                        4, 4, 4, 4, 5, 5, 5, 5, 0};
D4_CONST uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
D4_CONST uint8_t c_dist_ebits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                         9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
D4_CONST uint8_t c_codelen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Constants.distance2dist (Constants.java:9-13) in closed form.
__host__ __device__ inline int dist_sym(int distance) {
    unsigned d = (unsigned)distance - 1;
    if (d < 4) return (int)d;
#ifdef __CUDA_ARCH__
    int hb = 31 - __clz(d);
#else
    int hb = 31 - __builtin_clz(d);
#endif
    return 2 * hb + (int)((d >> (hb - 1)) & 1);
}
__host__ __device__ inline int dist_ebits_of(int dsym) { return dsym < 4 ? 0 : (dsym >> 1) - 1; }
__host__ __device__ inline int len_ebits_of(int lensym) {  // lensym 257..285
    int i = lensym - 257;
    return (i < 8 || i == 28) ? 0 : (i >> 2) - 1;
}

// ---- limits ---------------------------------------------------------------------------------------
constexpr int MAX_LL = 288, MAX_D = 32, MAX_CL = 19, MAX_PAIRS = 320;

// header RLE pair (the reference stores these as LitLen(dist=run, litlen=sym), DeflateBlockHuffman
// .java:997-999) in 16 bits: sym 5 bits | run field 7 bits << 5 | repeated value 4 bits << 12.  The run
// field holds the run length (0 = plain length, 3..10 for 16/17) except for sym 18, where it holds run - 11.
__host__ __device__ inline uint16_t pair_pack(int sym, int run, int val) {
    return (uint16_t)(sym | ((sym == 18 ? run - 11 : run) << 5) | (val << 12));
}
__host__ __device__ inline int pair_sym(uint16_t p) { return p & 31; }
__host__ __device__ inline int pair_run(uint16_t p) { return (p & 31) == 18 ? ((p >> 5) & 127) + 11 : (p >> 5) & 127; }
__host__ __device__ inline int pair_val(uint16_t p) { return (p >> 12) & 15; }

// Code tables of a Huffman block: lengths only — every code the reference ever writes is the canonical
// code of its length set (Huffman.java:35-64, HuffmanTree.java:164-192, HuffmanTable.java:166-209).
struct Tab {
    uint8_t L[MAX_LL];
    uint8_t D[MAX_D];
    uint16_t nL, nD;   // table lengths (numLitlenLens/numDistLens after rewriteHeader, :495-496)
    uint8_t type;      // 1 FIXED, 2 DYNAMIC
    uint8_t pad[3];
};
struct Hdr {
    uint16_t pairs[MAX_PAIRS];
    uint16_t np;
    uint8_t CL[MAX_CL];
    uint8_t ncl;       // numCodelenLens
    int32_t bits;      // dynamicHeaderSizeBits
};

// Parsed block record (written by the decode kernels, read by everything else).
struct BlockRec {
    uint64_t hdr_bit;     // bit position of the 3-bit block header, stream relative
    uint64_t data_bit;    // first bit of the symbol data (or of the stored bytes)
    uint64_t end_bit;     // first bit after the block
    uint64_t sym_base;    // first symbol (stream relative until the host rebases; then pool index)
    uint64_t out_base;    // first decoded byte (same)
    uint32_t n_sym;
    uint32_t out_len;
    uint32_t chunk_base, n_chunks;
    int64_t payload_bits; // litlenSizeBits
    uint8_t type;         // 0 STORED 1 FIXED 2 DYNAMIC
    uint8_t bfinal;
    uint8_t pad[6];
    Tab tab;
    Hdr hdr;
};

// one speculative-decode chunk that turned out valid: where to start and where its symbols go
struct ChunkRec {
    uint32_t start_rel;   // bit offset from the block's data_bit
    uint32_t sym_idx;     // block-relative index of its first symbol
    uint32_t out_off;     // block-relative decoded offset of its first symbol
    uint32_t n;           // symbols in the chunk
};

// A stream is parsed by one or more "walkers": the input is cut into segments, walker j of a stream
// starts at the first plausible block header inside segment j (k_find; walker 0 at bit 0) and walks
// blocks until it reaches a block boundary at or beyond its segment's end.  The host then follows the
// chain from walker 0: a walker is valid iff its predecessor ended exactly where it started; a missing
// or wrong start is re-walked from the known position (deft4cu.cu, Batch::parse).
constexpr uint64_t BIT_NONE = ~0ull;
struct StreamDesc {                 // one per walker
    uint64_t in_off, in_len;        // the whole stream: bytes in the batch input buffer
    uint64_t start_bit, stop_bit;   // first block header of this walker (BIT_NONE: nothing to do); segment end
    uint64_t blk_base, blk_cap;     // BlockRec slots
    uint64_t chunk_base, chunk_cap; // ChunkRec slots
    uint64_t sym_base, out_base;    // pool bases (filled before emit)
    uint64_t stream_out_base;       // pool offset of the stream's first decoded byte
    uint32_t walker0, nseg;         // the walkers of this stream are [walker0, walker0 + nseg)
    uint32_t spec;                  // 1: speculative start (from k_find); 0: known block boundary
    uint32_t pad;
};
struct StreamInfo {                 // result of the count pass (per walker); the host aggregates per stream
    int32_t status;
    uint32_t n_blocks;
    uint64_t n_syms, out_len, consumed, n_chunks, total_bits;
    uint64_t end_bit;               // walker: bit position where it stopped (next block header / end of stream)
    uint32_t final_seen;            // walker: stopped at BFINAL
    uint32_t pad;
};

// ---- little bit reader ----------------------------------------------------------------------------
// >= 57 valid bits starting at bit position `bitpos`; the input buffer is padded so the 16-byte
// over-read is always in bounds.
#ifndef D4_HOST_TEST
__device__ __forceinline__ uint64_t peek_bits(const uint8_t* in, uint64_t bitpos) {
    const uint8_t* p = in + (bitpos >> 3);
    uintptr_t a = (uintptr_t)p;
    const uint64_t* q = (const uint64_t*)(a & ~(uintptr_t)7);
    int sh = (int)(a & 7) * 8;
    uint64_t lo = __ldg(q), hi = __ldg(q + 1);
    uint64_t w = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    return w >> (bitpos & 7);
}
#endif

#define D4_CUDA_CHECK(x)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            d4::set_error(std::string(#x) + ": " + cudaGetErrorString(e_));                   \
            return DEFT4CU_ERR_CUDA;                                                          \
        }                                                                                     \
    } while (0)

}  // namespace d4
