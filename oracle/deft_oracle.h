/*
 * deft_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a literal CPU restatement of deft4j-base (the `optimise -m NONE` hot path).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libdeft4cu.so) never links, loads or calls it.
 *
 * Parity status: PINNED — the restatement reproduces all nine golden pairs of the reference
 * (test/ *-opt.* files, runTestOpt.sh:3-11) byte for byte, see tests/test_oracle_golden.py.
 */
#ifndef DEFT_ORACLE_H
#define DEFT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_stream ora_stream;

/* DeflateStream.parse(InputStream)  (base/deflate/DeflateStream.java:72-126).
 * Returns NULL on parse failure.  *consumed = bytes pulled from the input (BitInputStream.pos). */
ora_stream* ora_parse(const uint8_t* data, size_t len, size_t* consumed);
void        ora_free(ora_stream* s);

/* DeflateStream.optimise(boolean) (DeflateStream.java:496-566) → bits saved. */
int64_t ora_optimise(ora_stream* s, int merge_blocks);
/* DeflateStream.getSizeBits() (DeflateStream.java:171-182). */
int64_t ora_size_bits(const ora_stream* s);
/* DeflateStream.getUncompressedData() (DeflateStream.java:159-169): length, then copy. */
size_t  ora_uncompressed_len(const ora_stream* s);
void    ora_uncompressed(const ora_stream* s, uint8_t* dst);
/* DeflateStream.write(OutputStream) (DeflateStream.java:128-145).  Returns bytes needed;
 * writes only if cap is large enough. */
size_t  ora_write(const ora_stream* s, uint8_t* dst, size_t cap);

/* Block model inspection (printBlockInfo, DeflateStream.java:35-51, plus symbol dumps used by the
 * GPU decode parity tests). */
uint32_t ora_block_count(const ora_stream* s);
typedef struct ora_block_info {
    int32_t  type;            /* 0 stored, 1 fixed, 2 dynamic (DeflateBlockType ordinal) */
    int64_t  size_bits;       /* getSizeBits(pos) at its stream position (without the 3 header bits) */
    int64_t  position;        /* bit position of the 3-bit block header */
    uint64_t uncompressed_len;
    uint32_t n_symbols;       /* litlens.size() (0 for stored) */
    uint32_t n_rle_pairs;     /* rlePairs.size() (dynamic only) */
    int32_t  num_litlen_lens, num_dist_lens, num_codelen_lens;
    int64_t  litlen_size_bits, header_size_bits;
} ora_block_info;
int ora_block_info_get(const ora_stream* s, uint32_t block, ora_block_info* out);
/* symbols: 3 ints per symbol {dist, litlen, edgecase}; returns count */
uint32_t ora_block_symbols(const ora_stream* s, uint32_t block, int32_t* dst, uint32_t cap_syms);
/* rle pairs: 2 ints per pair {dist(run length or 0), sym}; returns count */
uint32_t ora_block_rle_pairs(const ora_stream* s, uint32_t block, int32_t* dst, uint32_t cap_pairs);
/* code length tables: which = 0 litlen, 1 dist, 2 codelen; returns table length */
uint32_t ora_block_codelens(const ora_stream* s, uint32_t block, int which, int32_t* dst, uint32_t cap);

/* One-shot raw-stream helper: parse + optimise + write.  Returns 0 OK, 1 parse failed.
 * *out is malloc'd (free with ora_free_buf). */
int  ora_optimise_stream(const uint8_t* data, size_t len, int merge_blocks,
                         uint8_t** out, size_t* out_len, int64_t* saved_bits, size_t* consumed);
void ora_free_buf(uint8_t* p);

/* Unit-level entry points so tests can pin sub-steps (HuffmanTree, pack). */
/* HuffmanTree(freq, limit).getTable() (base/huffman/HuffmanTree.java:36-128,164-192). */
void ora_huffman_tree(const int32_t* freq, int n, int limit, int32_t* code_out, int32_t* len_out);
/* HuffmanTable.packCodeLengths (base/huffman/HuffmanTable.java:42-159); flags bit0 ohh,1 use8,2 use7,
 * 3 alt8,4 noRep,5 noZRep,6 noZRep2,7 noRepZeros.  Returns number of ints written (flat list). */
int  ora_pack_code_lengths(const int32_t* lit, int nlit, const int32_t* dist, int ndist, int flags,
                           int32_t* dst, int cap);

/* One header-strategy trial (DeflateStream.optimiseBlockDynBlock, DeflateStream.java:184-198) on a bare dynamic
 * block holding only the two code-length tables; flags as above plus bit8 = prune.  post_op applies one more
 * mutator afterwards: 0 none, 1 recodeHeader, 2 recodeHeaderToLessRLEMatches, 3 optimiseHeader.
 * Returns dynamicHeaderSizeBits; pairs_out = {run, sym} per RLE pair (cap 320), cl_out = 19 header code lengths. */
int64_t ora_header_trial(const int32_t* lit, int nlit, const int32_t* dist, int ndist, int flags, int post_op,
                         int32_t* pairs_out, int32_t* np_out, int32_t* cl_out, int32_t* ncl_out);

/* candidate trace for locating divergences: while armed, every candidate the selection callback of
 * optimiseBlock compares is logged as {index within the call, size}; {-1, incumbent size} opens a call. */
void   ora_trace_begin(int64_t* buf_pairs, size_t cap_pairs);
size_t ora_trace_end(void);

/* counters for the last ora_optimise call (instrumentation for DESIGN.md sizing) */
typedef struct ora_stats {
    uint64_t optimise_block_calls, candidates, header_rewrites, tree_builds_small, tree_builds_big,
             symbol_passes;
} ora_stats;
void ora_get_stats(ora_stats* out);
void ora_reset_stats(void);

#ifdef __cplusplus
}
#endif
#endif
