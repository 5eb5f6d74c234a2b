// hosttest.cu — the single-thread building blocks of the cost model (huff.cuh) compiled as HOST code.
// TEST INFRASTRUCTURE: lets the CPU-only test suite pin the exact source the kernels execute
// (Huffman tree with java.util.PriorityQueue mechanics, header RLE packing, header trials) against the
// oracle.  Never loaded by the product.
#define D4_HOST_TEST 1
#include "huff.cuh"

using namespace d4;

extern "C" {

int host_huff_tree(const uint32_t* freq, int n, int limit, uint8_t* lens) {
    if (n <= 19) { static TreeWsCL ws; return huff_tree<21, 46>(freq, n, limit, lens, ws); }
    if (n <= 30) { static TreeWs<32, 68> ws; return huff_tree<32, 68>(freq, n, limit, lens, ws); }
    static TreeWs<290, 584> ws;
    return huff_tree<290, 584>(freq, n, limit, lens, ws);
}

// flags: bits 0-7 strategy, bit 8 prune; post_op: 0 none, 1 hdr_recode, 2 hdr_recode_less, 3 hdr_optimise
long long host_header_trial(const uint8_t* lit, int nlit, const uint8_t* dist, int ndist, int flags, int post_op,
                            int32_t* pairs_out, int32_t* np_out, int32_t* cl_out, int32_t* ncl_out) {
    static Tab t;
    static Hdr h;
    static TreeWsCL ws;
    for (int i = 0; i < MAX_LL; i++) t.L[i] = i < nlit ? lit[i] : 0;
    for (int i = 0; i < MAX_D; i++) t.D[i] = i < ndist ? dist[i] : 0;
    t.nL = (uint16_t)nlit; t.nD = (uint16_t)ndist; t.type = 2;
    if (hdr_trial(t, flags, h, ws)) return -1;
    if (post_op == 1) { if (hdr_recode(h, ws)) return -1; }
    else if (post_op == 2) { if (hdr_recode_less(h, ws)) return -1; }
    else if (post_op == 3) hdr_optimise(h);
    for (int i = 0; i < h.np; i++) { pairs_out[2 * i] = pair_run(h.pairs[i]); pairs_out[2 * i + 1] = pair_sym(h.pairs[i]); }
    *np_out = h.np;
    for (int i = 0; i < 19; i++) cl_out[i] = h.CL[i];
    *ncl_out = h.ncl;
    return h.bits;
}

int host_trial_flags(int k) { return c_trial_flags[k]; }

// the compact header-code workspace (weights must add up to less than 1024)
int host_huff_tree_compact(const uint32_t* freq, int n, int limit, uint8_t* lens) {
    static TreeWsCLc ws;
    return huff_tree_ws(freq, n, limit, lens, ws);
}

// size-only evaluation of one rewrite strategy for both prune values (trial_sizes over the run list)
int host_trial_sizes(const uint8_t* lit, int nlit, const uint8_t* dist, int ndist, int flags, int32_t* no_prune, int32_t* prune) {
    static Tab t;
    static RunList rl;
    static TreeWsCL ws;    // the workspace the engine's trial threads use
    for (int i = 0; i < MAX_LL; i++) t.L[i] = i < nlit ? lit[i] : 0;
    for (int i = 0; i < MAX_D; i++) t.D[i] = i < ndist ? dist[i] : 0;
    t.nL = (uint16_t)nlit; t.nD = (uint16_t)ndist; t.type = 2;
    runlist_build(t, rl);
    int a = 0, b = 0;
    if (trial_sizes(rl, flags, &a, &b, ws)) return 1;
    *no_prune = a; *prune = b;
    return 0;
}

}  // extern "C"
