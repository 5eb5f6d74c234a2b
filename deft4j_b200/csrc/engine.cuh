// engine.cuh — the block candidate enumerator (DeflateStream.optimiseBlock, DeflateStream.java:343-490)
// as CTA-cooperative device code.  One CTA owns one block; every function below is called by ALL
// threads of the CTA with uniform arguments.
//
// A candidate ("Cand") is the reference's DeflateBlockHuffman copy reduced to what determines its bytes:
//   Tab (code lengths + table lengths + type), Hdr (header RLE pairs, header code, numCodelenLens),
//   payload = litlenSizeBits, and a bit mask over the block's symbols (1 = match replaced by literals).
// Masks are immutable entries of a per-CTA pool in global scratch; a candidate holds a mask id.
//
// O(n) work (n = symbols of the block) is done by CTA-wide passes:
//   pass_replace : replaceBackrefsWithLiteralsIfSmaller (DeflateBlockHuffman.java:222-319)
//   pass_least   : removeDistLitLeastExpensive          (:373-458)
//   pass_hist_full / hist deltas : the histogram loop of recodeHuffman (:671-681); payload = hist . (len + extra bits)
// Serial work runs in single threads (huff.cuh): the litlen / distance Huffman trees (exact PriorityQueue order),
// the header code, run replacement in headers.  The 56 header strategy trials are evaluated size-only, one
// thread per rewrite strategy, over a run list of the code lengths; the default header rewrite and the header
// recode are spread over the CTA (run list -> scan -> parallel pair emission).
//
// Exact shortcuts (pure-function memoisation; none changes any candidate's size or the order in which
// candidates are compared).  The enumeration revisits the same (symbol list, code tables) states many
// times (e.g. recoded(e) inside addOptimisedRecoded(post(e)) is the seed of the next Run), so states
// are hash-consed into ids and every O(n) pass is a function of ids:
//   * masks are immutable and live in a per-CTA pool; a candidate holds a mask id `mid`; Tabs are interned
//     into `tabid`s.  replace/least passes are memoised on (mid, tabid, op) -> (mid', payload delta);
//     recodeHuffman's result (tables, payload, default header) depends only on the mask -> cached per mid.
//     When a pool fills up, everything not referenced by a candidate slot is dropped (flush_all).
//   * the 56 header-strategy trials of a base depend only on (Tab, payload) and their sizes are
//     payload + f(Tab, strategy) -> the first-minimum strategy per tabid is cached, and the
//     addOptimisedRecoded(prune) sweep (DeflateStream.java:431) is skipped: it re-evaluates candidates
//     with exactly the sizes of the sweep on `post` (:418) — both are copies of the same block that
//     differ only in the header, which every base/trial discards — so under the strict `<` of the
//     selection callback (:357) none of them can ever be chosen.
#pragma once
#include "huff.cuh"
#include "parse.cuh"

namespace d4 {

// threads per block-in-flight and CTAs per SM the kernels are built for (A/B knobs: -DD4_ENG_NT=128 -DD4_ENG_MINB=7)
#ifndef D4_ENG_NT
#define D4_ENG_NT 256
#endif
#ifndef D4_ENG_MINB
#define D4_ENG_MINB 4
#endif
constexpr int ENG_NT = D4_ENG_NT;
constexpr int ERR_TREE = 11, ERR_ROUNDS = 12, ERR_WRITER = 2;  // internal-limit codes reported through gerr

// parity-debug instrumentation (deft4cu_debug_trace): when armed, every candidate the selection callback sees
// is logged as (candidate index, size) and the header-strategy memo is bypassed so all 56 trials are logged
__device__ long long* g_trace = nullptr;
__device__ unsigned g_trace_cap = 0;
__device__ unsigned g_trace_n = 0;
__device__ __forceinline__ void trace_put(long long idx, long long sz) {
    unsigned k = atomicAdd(&g_trace_n, 1u);
    if (k < g_trace_cap) { g_trace[2 * k] = idx; g_trace[2 * k + 1] = sz; }
}
constexpr int NCAND = 16;
#ifdef D4_SMALL_POOLS           // stress build: forces the pool-overflow path (flush_all) on ordinary inputs
constexpr int MAXM = 20, MAXT = 20, MEMO_P = 64, DCN_MAX = 3;
#else
constexpr int DCN_MAX = 64;   // per-table literal-minus-match cost arrays kept per block (round-robin eviction)
constexpr int MAXM = 256;     // distinct symbol-list masks kept per block
constexpr int MAXT = 256;     // distinct code tables kept per block
constexpr int MEMO_P = 1024;  // pass memo slots (open addressing, kept under 3/4 full)
#endif
constexpr int ERR_POOL = 15;
constexpr int TRIAL_UNSET = (int)0x80000000;

struct Cand {
    Tab tab;
    Hdr hdr;
    long long payload;  // litlenSizeBits
    uint16_t mid;       // mask id (engine-internal: index into the CTA's mask pool)
    uint16_t tabid;     // interned Tab id (engine-internal)
    uint32_t pad2;
};
struct PVal { uint32_t mid; uint32_t pad; long long delta; };
// memo tables of one CTA in global scratch (kept out of shared memory so that more CTAs fit on an SM; they are
// probed by thread 0 or scanned by all threads a few hundred times per round)
struct EngG {
    unsigned long long pkey[MEMO_P];   // pass memo keys (open addressing); values in Eng::pvals
    unsigned long long maskHash[MAXM];
    unsigned long long tabHash[MAXT];
    int tabTrialBits[MAXT];
    unsigned char tabTrialArg[MAXT];
};
__device__ __forceinline__ long long cand_size(const Cand& c) { return c.payload + (c.tab.type == 2 ? c.hdr.bits : 0); }

struct BlkView {
    const uint32_t* sym;
    const uint32_t* symout;
    const uint8_t* out;
    uint32_t n;       // symbols (including a NOP left by a merge)
    uint32_t nwords;  // mask words
    uint64_t ulen;    // decoded length
    uint64_t out_off; // pool offset of the block's first decoded byte
};

struct EngSmem {
    Cand c[NCAND];
    uint32_t hist[320];  // [0,286) litlen, [288,318) dist
    int leastSum[32], leastCnt[32];
    unsigned leastBlocked, leastSeen;
    unsigned long long red;
    int redAny;
    int err;
    TreeWs<290, 584> tl;
    TreeWs<32, 68> td;
    TreeWsCL wsCL;       // header-code tree workspace of thread 0 (local memory costs an L2 round trip per access)
    // selection state
    long long bestSize;
    int bestStored;
    long long sizeI, sizeC1, restMin;
    unsigned candIndex, bestIndex;
    // pools and memo tables (the tables themselves live in global scratch, EngG)
    unsigned char recodeValid[MAXM];
    int nMasks;
    int nTabs, fixedTab;
    int nP;
    unsigned char tabDc[MAXT];        // tabid -> cost-array slot (0xFF: none)
    unsigned short dcOwner[DCN_MAX];  // slot -> tabid (0xFFFF: free)
    int dcNext;
    int remap[NCAND], uniq[NCAND];
    unsigned long long uh[NCAND];
    int tmpIdx, redAny2;
    int trialBits[4 * 56];
    unsigned long long hred[ENG_NT / 32];
};

enum { C_B = 0, C_BEST, C_O, C_H, C_E, C_X, C_CHK, C_T, C_Y, C_B1, C_B2, C_B3, C_B4, C_PP, C_CHK2, C_TMP };

// cycle accounting per engine phase (-DD4_PROF builds only; read back with deft4cu_debug_prof): thread 0's
// clock64 deltas, [cat] = cycles, [32 + cat] = calls.  Categories nest (recode-miss contains hist, trees, ...).
#ifdef D4_PROF
__device__ unsigned long long g_prof[64];
#define P0() const long long p0_ = clock64()
#define P1(cat) do { if (tid == 0) { atomicAdd(&g_prof[cat], (unsigned long long)(clock64() - p0_)); atomicAdd(&g_prof[32 + (cat)], 1ull); } } while (0)
#else
#define P0()
#define P1(cat)
#endif
enum { PR_BLOCK = 0, PR_REPL_HIT, PR_REPL_MISS, PR_LEAST_HIT, PR_LEAST_MISS, PR_HIST, PR_RECODE_HIT, PR_RECODE_MISS, PR_TREES,
       PR_HDR_DEFAULT, PR_INTERN_TAB, PR_INTERN_MASK, PR_COPY, PR_CB, PR_TRIALS, PR_TRIALS_EVAL, PR_HDR_OPT, PR_HDR_RECODE,
       PR_TO_FIXED, PR_FLUSH, PR_PAYLOAD };

#ifdef D4_VERIFY
#define D4V(c, op) verify(c, op)
#else
#define D4V(c, op)
#endif

struct Eng {
    EngSmem* S;
    BlkView v;
    uint32_t* masks;      // (MAXM + NCAND) slots of maxwords: the mask pool + the evacuation area of flush_all
    uint32_t maxwords;
    Tab* tabs;            // MAXT interned code tables
    Cand* recode;         // MAXM: recodeHuffman result per mask id
    PVal* pvals;          // MEMO_P pass memo values
    EngG* G;
    short* dc;            // dcn arrays of maxn: per symbol, (literal cost - match cost) under one Tab
    uint32_t* hists;      // MAXM * 320: symbol histogram per mask id
    uint8_t* kind;        // maxn: 0 = not a match, else length symbol - 256
    uint32_t* meta;       // maxn: match only: len-3 | dist symbol << 9 | extra bits of the match << 14
    uint32_t* P;          // maxp: exclusive prefix sums of per-byte literal costs (nullptr: blocks too long, byte loops)
    uint32_t maxp;
    uint32_t prefixRatio; // use P when decoded bytes per symbol >= this
    uint32_t maxn;        // maxwords * 32
    int dcn;
    int tid;

    __device__ uint32_t* maskp(int id) const { return masks + (size_t)id * maxwords; }

    // ---- candidate copy (DeflateBlockHuffman.copy, :1174-1205, with value semantics; masks are immutable
    //      pool entries, so a copy shares its source's mask id) --------------------------------------------
    __device__ __noinline__ void copy(int dst, int src) {
        if (dst == src) return;
        P0();
        const uint32_t* s = (const uint32_t*)&S->c[src];
        uint32_t* d = (uint32_t*)&S->c[dst];
        for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
        __syncthreads();
        P1(PR_COPY);
        D4V(dst, 7);
    }

    // ---- selection callback (DeflateStream.java:349-368) ------------------------------------------
    __device__ __noinline__ void cb(int c, bool isRest = true) {
        P0();
        long long sz = cand_size(S->c[c]);
        bool better = sz < S->bestSize;
        __syncthreads();
        if (tid == 0) {
            if (g_trace) trace_put(S->candIndex, sz);
            if (isRest && sz < S->restMin) S->restMin = sz;
            if (better) { S->bestSize = sz; S->bestStored = 0; S->bestIndex = S->candIndex; }
            S->candIndex++;
        }
        if (better) copy(C_BEST, c); else __syncthreads();
        P1(PR_CB);
    }

    __device__ __noinline__ unsigned long long hash_words(const uint32_t* p, int nwords32) {
        unsigned long long h = 0;
        for (int k = tid; k < nwords32; k += ENG_NT) {
            unsigned long long x = (unsigned long long)p[k] + 0x9E3779B97F4A7C15ull * (unsigned long long)(k + 1);
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
            h += x;
        }
        for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
        if ((tid & 31) == 0) S->hred[tid >> 5] = h;
        __syncthreads();
        unsigned long long r = 0;
        for (int k = 0; k < ENG_NT / 32; k++) r += S->hred[k];
        __syncthreads();
        return r | 1ull;
    }

    // ---- pools --------------------------------------------------------------------------------------------
    // start of a new block (new symbol view): everything is forgotten
    __device__ __noinline__ void begin_block() {
        __syncthreads();
        for (int k = tid; k < MEMO_P; k += ENG_NT) G->pkey[k] = 0;
        for (int k = tid; k < MAXM; k += ENG_NT) S->recodeValid[k] = 0;
        if (tid < NCAND) { S->c[tid].mid = 0; S->c[tid].tabid = 0; S->c[tid].pad2 = 0; }
        for (int k = tid; k < MAXT; k += ENG_NT) S->tabDc[k] = 0xFF;
        if (tid < DCN_MAX) S->dcOwner[tid] = 0xFFFF;
        if (tid == 0) { S->nMasks = 0; S->nTabs = 0; S->nP = 0; S->fixedTab = -1; S->dcNext = 0; }
        __syncthreads();
    }

    // Tab of candidate c -> S->c[c].tabid (hash-consed; a hash hit is confirmed by a full comparison)
    __device__ __noinline__ void intern_tab(int c) {
        P0();
        const uint32_t* q = (const uint32_t*)&S->c[c].tab;
        const unsigned long long h = hash_words(q, (int)(sizeof(Tab) / 4));
        if (tid == 0) { S->tmpIdx = -1; S->redAny2 = 0; }
        __syncthreads();
        const int nT = S->nTabs;
        for (int k = tid; k < nT; k += ENG_NT)
            if (G->tabHash[k] == h) atomicMax(&S->tmpIdx, k);
        __syncthreads();
        int hit = S->tmpIdx;
        if (hit >= 0) {
            const uint32_t* a = (const uint32_t*)&tabs[hit];
            bool diff = false;
            for (int k = tid; k < (int)(sizeof(Tab) / 4); k += ENG_NT) diff |= a[k] != q[k];
            if (diff) S->redAny2 = 1;
            __syncthreads();
            if (S->redAny2) hit = -1;
        }
        __syncthreads();
        if (hit < 0) {
            if (tid == 0) {
                int slot = S->nTabs;
                if (slot >= MAXT) { S->err = ERR_POOL; slot = MAXT - 1; } else S->nTabs = slot + 1;
                G->tabHash[slot] = h;
                G->tabTrialBits[slot] = TRIAL_UNSET;
                S->tmpIdx = slot;
            }
            __syncthreads();
            hit = S->tmpIdx;
            uint32_t* a = (uint32_t*)&tabs[hit];
            for (int k = tid; k < (int)(sizeof(Tab) / 4); k += ENG_NT) a[k] = q[k];
        }
        if (tid == 0) S->c[c].tabid = (uint16_t)hit;
        __syncthreads();
        P1(PR_INTERN_TAB);
    }

    // the mask just written into pool slot nMasks -> its id (an equal older mask wins, so equal symbol lists
    // reached along different paths share their memo entries)
    __device__ __noinline__ int intern_mask() {
        P0();
        const int fresh = S->nMasks;
        const uint32_t* q = maskp(fresh);
        const unsigned long long h = hash_words(q, (int)v.nwords);
        if (tid == 0) { S->tmpIdx = -1; S->redAny2 = 0; }
        __syncthreads();
        for (int k = tid; k < fresh; k += ENG_NT)
            if (G->maskHash[k] == h) atomicMax(&S->tmpIdx, k);
        __syncthreads();
        int hit = S->tmpIdx;
        if (hit >= 0) {
            const uint32_t* a = maskp(hit);
            bool diff = false;
            for (uint32_t k = tid; k < v.nwords; k += ENG_NT) diff |= a[k] != q[k];
            if (diff) S->redAny2 = 1;
            __syncthreads();
            if (S->redAny2) hit = -1;
        }
        __syncthreads();
        if (hit < 0) {
            hit = fresh;
            if (tid == 0) { G->maskHash[fresh] = h; S->recodeValid[fresh] = 0; S->nMasks = fresh + 1; }
        }
        __syncthreads();
        P1(PR_INTERN_MASK);
        return hit;
    }

    // a pool is full: keep only what the candidate slots reference
    __device__ __noinline__ void flush_all() {
        P0();
        __syncthreads();
        if (tid == 0) {
            int nu = 0;
            for (int c = 0; c < NCAND; c++) {
                const int m = S->c[c].mid;
                int k = 0;
                while (k < nu && S->uniq[k] != m) k++;
                if (k == nu) { S->uniq[nu] = m; S->uh[nu] = G->maskHash[m]; nu++; }
                S->remap[c] = k;
            }
            S->tmpIdx = nu;
        }
        __syncthreads();
        const int nu = S->tmpIdx;
        for (int k = 0; k < nu; k++) {
            const uint32_t* s = maskp(S->uniq[k]);
            uint32_t* d = maskp(MAXM + k);
            for (uint32_t w = tid; w < v.nwords; w += ENG_NT) d[w] = s[w];
            const uint32_t* hs = hists + (size_t)S->uniq[k] * 320;
            uint32_t* hd = hists + (size_t)(MAXM + k) * 320;
            for (int w = tid; w < 320; w += ENG_NT) hd[w] = hs[w];
        }
        __syncthreads();
        for (int k = 0; k < nu; k++) {
            const uint32_t* s = maskp(MAXM + k);
            uint32_t* d = maskp(k);
            for (uint32_t w = tid; w < v.nwords; w += ENG_NT) d[w] = s[w];
            const uint32_t* hs = hists + (size_t)(MAXM + k) * 320;
            uint32_t* hd = hists + (size_t)k * 320;
            for (int w = tid; w < 320; w += ENG_NT) hd[w] = hs[w];
        }
        if (tid < NCAND) S->c[tid].mid = (uint16_t)S->remap[tid];
        if (tid < nu) G->maskHash[tid] = S->uh[tid];
        for (int k = tid; k < MAXM; k += ENG_NT) S->recodeValid[k] = 0;
        for (int k = tid; k < MEMO_P; k += ENG_NT) G->pkey[k] = 0;
        for (int k = tid; k < MAXT; k += ENG_NT) S->tabDc[k] = 0xFF;
        if (tid < DCN_MAX) S->dcOwner[tid] = 0xFFFF;
        __syncthreads();
        if (tid == 0) { S->nMasks = nu; S->nTabs = 0; S->nP = 0; S->fixedTab = -1; S->dcNext = 0; }
        __syncthreads();
        for (int c = 0; c < NCAND; c++) intern_tab(c);
        P1(PR_FLUSH);
    }
    // every op creates at most one mask, one Tab and one memo entry
    __device__ __forceinline__ void maybe_flush() {
        const bool need = S->nMasks >= MAXM || S->nTabs >= MAXT || S->nP >= MEMO_P * 3 / 4;
        if (need) flush_all();
    }

    // pass memo (thread 0 only): slot of `key`, or -1 - (insert position)
    __device__ __forceinline__ int pm_find(unsigned long long key) const {
        unsigned long long x = key;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        unsigned h = (unsigned)x & (MEMO_P - 1);
        while (true) {
            const unsigned long long k = G->pkey[h];
            if (k == key) return (int)h;
            if (k == 0) return -1 - (int)h;
            h = (h + 1) & (MEMO_P - 1);
        }
    }
    __device__ __forceinline__ unsigned long long pm_key(int mid, int tabid, int op) const {
        return (1ull << 63) | ((unsigned long long)op << 32) | ((unsigned long long)tabid << 16) | (unsigned long long)mid;
    }

    // ---- CTA-wide passes ----------------------------------------------------------------------------
    // literal cost of the bytes a match produces; returns -1 when a byte has no code.  Loads are issued
    // eight at a time so their latencies overlap.
    __device__ __forceinline__ int lit_cost(const uint8_t* L, uint32_t off, int len) const {
        const uint8_t* p = v.out + off;
        int tot = 0, bad = 0;
        for (int k = 0; k < len; k += 8) {
            uint32_t b[8];
#pragma unroll
            for (int j = 0; j < 8; j++) b[j] = (k + j < len) ? (uint32_t)p[k + j] : 256u;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (b[j] < 256u) { const int cbits = L[b[j]]; bad |= (cbits < 1); tot += cbits; }
        }
        return bad ? -1 : tot;
    }
    // getLitLenSize for a match (:112-128)
    __device__ __forceinline__ int ref_cost(const Tab& t, uint32_t s) const {
        int ls = sym_lensym(s), ds = dist_sym(sym_dist(s));
        return t.L[ls] + len_ebits_of(ls) + t.D[ds] + dist_ebits_of(ds);
    }

    // Per-Tab cost array: dc[i] = (literal cost of match i's bytes) - (cost of the match) under candidate c's
    // tables, DC_BLOCKED when a byte has no code (DeflateBlockHuffman.java:238-246).  It does not depend on the
    // mask, so every replace / least pass under the same tables reads it instead of walking the bytes again.
    static constexpr short DC_BLOCKED = 0x7FFF, DC_NOT_MATCH = 0x7FFE;
    static constexpr uint32_t UNC = 1u << 20;  // prefix-sum cost of a byte without a code (a match has <= 258 bytes)

    // P[j] = sum of literal costs of the block's decoded bytes before position j (positions count from the
    // 16-byte boundary at or below the block's first byte), so a match's literal cost is P[end] - P[start].
    // Coalesced 128-bit loads, one tile of ENG_NT * 16 bytes per step, CTA-wide scan.
    __device__ __noinline__ void build_prefix(const uint8_t* L) {
        __shared__ uint32_t s_wt[ENG_NT / 32];
        for (int k = tid; k < 256; k += ENG_NT) S->hist[k] = L[k] ? (uint32_t)L[k] : UNC;
        __syncthreads();
        const uint64_t a0 = v.out_off & ~15ull;
        const uint32_t head = (uint32_t)(v.out_off - a0);
        const uint32_t endRel = head + (uint32_t)v.ulen;
        const uint8_t* base = v.out + a0;
        const int lane = tid & 31, wid = tid >> 5;
        uint32_t carry = 0;
        for (uint32_t T = 0; T <= endRel; T += ENG_NT * 16) {
            const uint32_t idx = T + 16u * tid;
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            uint32_t tot = 0;
            if (idx < endRel) {
                const uint4 q = *(const uint4*)(base + idx);
                w0 = q.x; w1 = q.y; w2 = q.z; w3 = q.w;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t wk = k < 4 ? w0 : k < 8 ? w1 : k < 12 ? w2 : w3;
                    const uint32_t j = idx + k;
                    tot += (j >= head && j < endRel) ? S->hist[(wk >> (8 * (k & 3))) & 0xff] : 0u;
                }
            }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += x;
            }
            if (lane == 31) s_wt[wid] = incl;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int k = 0; k < ENG_NT / 32; k++) { const uint32_t x = s_wt[k]; if (k < wid) wbase += x; total += x; }
            if (idx <= endRel) {  // second walk over the 16 bytes: running exclusive prefix, stored 4 at a time
                uint32_t run = carry + wbase + incl - tot;
                uint4* dst = (uint4*)(P + idx);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint32_t wk = g == 0 ? w0 : g == 1 ? w1 : g == 2 ? w2 : w3;
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t j = idx + 4 * g + k;
                        o[k] = run;
                        run += (j >= head && j < endRel) ? S->hist[(wk >> (8 * k)) & 0xff] : 0u;
                    }
                    dst[g] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            carry += total;
            __syncthreads();
        }
    }

    __device__ __noinline__ const short* ensure_dc(int c) {
        const Cand& cd = S->c[c];
        const int t = cd.tabid;
        int slot = S->tabDc[t];
        if (slot != 0xFF) return dc + (size_t)slot * maxn;
        __syncthreads();  // every thread has seen the miss before thread 0 records the new slot
        if (tid == 0) {
            slot = S->dcNext;
            S->dcNext = (slot + 1) % dcn;
            const int owner = S->dcOwner[slot];
            if (owner != 0xFFFF) S->tabDc[owner] = 0xFF;
            S->dcOwner[slot] = (unsigned short)t;
            S->tabDc[t] = (unsigned char)slot;
            S->tmpIdx = slot;
        }
        __syncthreads();
        slot = S->tmpIdx;
        short* d = dc + (size_t)slot * maxn;
        // prefix sums pay off when matches are long (one pass over the bytes instead of one per match byte); on
        // ordinary text (about 4 decoded bytes per symbol) the per-match loops move far fewer bytes through DRAM
        const bool prefix = P != nullptr && v.ulen + 64 <= (uint64_t)maxp && v.ulen >= (uint64_t)prefixRatio * v.n;
        if (prefix) {
            build_prefix(cd.tab.L);
            const uint32_t a0 = (uint32_t)(v.out_off & ~15ull);
            // 8 consecutive symbols per thread, every load of the batch issued before the first use
            for (uint32_t i0 = (uint32_t)tid * 8; i0 < v.nwords * 32; i0 += ENG_NT * 8) {
                const uint2 kk = *(const uint2*)(kind + i0);
                const uint4 ma = *(const uint4*)(meta + i0), mb = *(const uint4*)(meta + i0 + 4);
                const uint32_t mt[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                uint32_t st[8], lit[8];
                int kq[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    kq[u] = (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff);
                    if (i0 + u >= v.n) kq[u] = 0;
                    st[u] = kq[u] ? v.symout[i0 + u] - a0 : 0u;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) lit[u] = kq[u] ? P[st[u] + (mt[u] & 0x1FF) + 3] - P[st[u]] : 0u;
                uint32_t o[4];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    short x = DC_NOT_MATCH;
                    if (kq[u]) {
                        const int ref = cd.tab.L[256 + kq[u]] + cd.tab.D[(mt[u] >> 9) & 31] + (int)((mt[u] >> 14) & 31);
                        x = lit[u] >= UNC ? DC_BLOCKED : (short)((int)lit[u] - ref);
                    }
                    if (u & 1) o[u >> 1] |= (uint32_t)(unsigned short)x << 16; else o[u >> 1] = (uint32_t)(unsigned short)x;
                }
                *(uint4*)(d + i0) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        } else {
            for (uint32_t i = tid; i < v.n; i += ENG_NT) {
                short x = DC_NOT_MATCH;
                if (kind[i]) {
                    const uint32_t s = v.sym[i];
                    const int lit = lit_cost(cd.tab.L, v.symout[i], sym_len(s));
                    x = lit < 0 ? DC_BLOCKED : (short)(lit - ref_cost(cd.tab, s));
                }
                d[i] = x;
            }
        }
        __syncthreads();
        return d;
    }

    // match i leaves the symbol list and its bytes enter it as literals: histogram delta in S->hist
    __device__ __forceinline__ void hist_delta_replace(uint32_t i) {
        const uint32_t s = v.sym[i];
        atomicSub(&S->hist[sym_lensym(s)], 1u);
        atomicSub(&S->hist[288 + dist_sym(sym_dist(s))], 1u);
        const uint8_t* p = v.out + v.symout[i];
        const int len = sym_len(s);
        for (int k = 0; k < len; k++) atomicAdd(&S->hist[p[k]], 1u);
    }
    // hists[dst] = hists[src] + S->hist (the delta of the matches that were just replaced)
    __device__ __forceinline__ void hist_store_delta(int dst, int src) {
        const uint32_t* hs = hists + (size_t)src * 320;
        uint32_t* hd = hists + (size_t)dst * 320;
        for (int k = tid; k < 320; k += ENG_NT) hd[k] = hs[k] + S->hist[k];
    }

    // replaceBackrefsWithLiteralsIfSmaller(prune) on candidate c (in place)
    __device__ __noinline__ void pass_replace(int c, bool prune) {
        maybe_flush();
        P0();
        Cand& cd = S->c[c];
        const int mid = cd.mid;
        const unsigned long long key = pm_key(mid, cd.tabid, prune ? 1 : 0);
        if (tid == 0) { S->tmpIdx = pm_find(key); S->red = 0; S->redAny = 0; }
        __syncthreads();
        int slot = S->tmpIdx;
        __syncthreads();
        if (slot >= 0) {
            if (tid == 0) { const PVal pv = pvals[slot]; cd.mid = (uint16_t)pv.mid; cd.payload -= pv.delta; }
            __syncthreads();
            P1(PR_REPL_HIT);
            D4V(c, prune ? 2 : 1);
            return;
        }
        slot = -1 - slot;
        const short* d = ensure_dc(c);
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = 0;
        __syncthreads();
        const uint32_t* m = maskp(mid);
        const int fresh = S->nMasks;
        uint32_t* md = maskp(fresh);
        long long saved = 0;
        const int lane = tid & 31;
        {   // 8 consecutive symbols (one mask byte) per thread; loads batched
            const uint8_t* mb = (const uint8_t*)m;
            uint8_t* mdb = (uint8_t*)md;
            bool any = false;
            for (uint32_t i0 = (uint32_t)tid * 8; i0 < v.nwords * 32; i0 += ENG_NT * 8) {
                const uint2 kk = *(const uint2*)(kind + i0);
                const uint4 dq = *(const uint4*)(d + i0);
                const uint32_t ob = mb[i0 >> 3];
                const uint32_t dw[4] = {dq.x, dq.y, dq.z, dq.w};
                uint32_t nbits = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int k = (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff);
                    const int x = (int)(short)((u & 1) ? (dw[u >> 1] >> 16) : (dw[u >> 1] & 0xffff));
                    if (i0 + u < v.n && k && !((ob >> u) & 1) && (prune ? x <= 0 : x < 0)) { saved -= x; nbits |= 1u << u; }
                }
                mdb[i0 >> 3] = (uint8_t)(ob | nbits);
                for (uint32_t b = nbits; b; b &= b - 1) hist_delta_replace(i0 + (uint32_t)__ffs((int)b) - 1);
                any |= nbits != 0;
            }
            if (any) S->redAny = 1;
        }
        for (int dd = 16; dd > 0; dd >>= 1) saved += __shfl_xor_sync(0xffffffffu, saved, dd);
        if (lane == 0 && saved) atomicAdd(&S->red, (unsigned long long)saved);
        __syncthreads();
        int newmid = mid;
        if (S->redAny) {
            newmid = intern_mask();
            if (newmid == fresh) hist_store_delta(newmid, mid);
        }
        if (tid == 0) {
            PVal pv; pv.mid = (uint32_t)newmid; pv.pad = 0; pv.delta = (long long)S->red;
            pvals[slot] = pv;
            G->pkey[slot] = key;
            S->nP++;
            cd.mid = (uint16_t)newmid;
            cd.payload -= pv.delta;
        }
        __syncthreads();
        P1(PR_REPL_MISS);
        D4V(c, prune ? 2 : 1);
    }

    // removeDistLitLeastExpensive(mode) on candidate c (in place); no-op unless DYNAMIC
    __device__ __noinline__ void pass_least(int c, int mode) {
        Cand& cd = S->c[c];
        if (cd.tab.type != 2) return;
        maybe_flush();
        P0();
        const int mid = cd.mid;
        const unsigned long long key = pm_key(mid, cd.tabid, 2 + mode);
        if (tid == 0) S->tmpIdx = pm_find(key);
        if (tid < 32) { S->leastSum[tid] = 0; S->leastCnt[tid] = 0; }
        if (tid == 0) { S->leastBlocked = 0; S->leastSeen = 0; }
        __syncthreads();
        int slot = S->tmpIdx;
        __syncthreads();
        if (slot >= 0) {
            if (tid == 0) { const PVal pv = pvals[slot]; cd.mid = (uint16_t)pv.mid; cd.payload -= pv.delta; }
            __syncthreads();
            P1(PR_LEAST_HIT);
            D4V(c, 3 + mode);
            return;
        }
        slot = -1 - slot;
        const short* d = ensure_dc(c);
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = 0;
        const uint32_t* m = maskp(mid);
        const int lane = tid & 31;
        // per length symbol: sum of (literal - match) cost, count, blocked (:386-420); lanes of a warp that hold
        // the same length symbol are summed with one shared-memory atomic
        const uint8_t* mbytes = (const uint8_t*)m;
        // the warp-wide votes below need every lane of a warp in the loop: the bound is per warp, loads are guarded
        for (uint32_t wb = (uint32_t)(tid >> 5) * 256; wb < v.nwords * 32; wb += ENG_NT * 8) {
            const uint32_t i0 = wb + (uint32_t)lane * 8;
            const bool inr = i0 < v.nwords * 32;
            const uint2 kk = inr ? *(const uint2*)(kind + i0) : make_uint2(0, 0);
            const uint4 dq = inr ? *(const uint4*)(d + i0) : make_uint4(0, 0, 0, 0);
            const uint32_t ob = inr ? mbytes[i0 >> 3] : 0u;
            const uint32_t dw[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = (i0 + u < v.n) ? (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff) : 0;
                const bool live = k && !((ob >> u) & 1);
                const int x = live ? (int)(short)((u & 1) ? (dw[u >> 1] >> 16) : (dw[u >> 1] & 0xffff)) : 0;
                const bool blocked = live && x == DC_BLOCKED;
                const int bin = live ? k - 1 : 31;
                const unsigned grp = __match_any_sync(0xffffffffu, bin);
                // one group reduction carries both the count (bits 24+) and the biased cost sum (x >= -64, 32 lanes)
                const unsigned packed = __reduce_add_sync(grp, (live && !blocked) ? (1u << 24) + (unsigned)(x + 64) : 0u);
                const int cn = (int)(packed >> 24);
                const int xs = (int)(packed & 0xFFFFFFu) - 64 * cn;
                const unsigned anyBlocked = __ballot_sync(0xffffffffu, blocked) & grp;
                if (live && lane == __ffs(grp) - 1) {
                    atomicOr(&S->leastSeen, 1u << bin);
                    if (anyBlocked) atomicOr(&S->leastBlocked, 1u << bin);
                    if (cn) { atomicAdd(&S->leastSum[bin], xs); atomicAdd(&S->leastCnt[bin], cn); }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            int rem = -1, remSize = 0, remFreq = 0;
            for (int i = 0; i < 32; i++) {
                if (!((S->leastBlocked >> i) & 1) && ((S->leastSeen >> i) & 1)) {
                    bool doRem = mode == 1 ? S->leastCnt[i] < remFreq : S->leastSum[i] < remSize;
                    if (rem == -1 || doRem) { rem = i; remSize = S->leastSum[i]; remFreq = S->leastCnt[i]; }
                }
            }
            S->tmpIdx = rem;
            S->red = (unsigned long long)(long long)remSize;
        }
        __syncthreads();
        const int rem = S->tmpIdx;
        int newmid = mid;
        if (rem >= 0) {
            const int fresh = S->nMasks;
            uint32_t* md = maskp(fresh);
            uint8_t* mdb = (uint8_t*)md;
            for (uint32_t i0 = (uint32_t)tid * 8; i0 < v.nwords * 32; i0 += ENG_NT * 8) {
                const uint2 kk = *(const uint2*)(kind + i0);
                const uint32_t ob = mbytes[i0 >> 3];
                uint32_t nbits = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int k = (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff);
                    if (i0 + u < v.n && k == rem + 1) nbits |= 1u << u;
                }
                mdb[i0 >> 3] = (uint8_t)(ob | nbits);
                for (uint32_t b = nbits & ~ob; b; b &= b - 1) hist_delta_replace(i0 + (uint32_t)__ffs((int)b) - 1);
            }
            __syncthreads();
            newmid = intern_mask();
            if (newmid == fresh) hist_store_delta(newmid, mid);
        }
        if (tid == 0) {
            PVal pv; pv.mid = (uint32_t)newmid; pv.pad = 0; pv.delta = -(long long)S->red;
            pvals[slot] = pv;
            G->pkey[slot] = key;
            S->nP++;
            cd.mid = (uint16_t)newmid;
            cd.payload -= pv.delta;
        }
        __syncthreads();
        P1(PR_LEAST_MISS);
        D4V(c, 3 + mode);
    }

    // histogram of the symbol list with mask `mid` into S->hist, from the symbols (block start, checks)
    __device__ __noinline__ void pass_hist_full(int mid) {
        P0();
        const uint32_t* m = maskp(mid);
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < v.n; i += ENG_NT) {
            uint32_t s = v.sym[i];
            if (!sym_is_match(s)) {
                if (s <= 256) atomicAdd(&S->hist[s], 1u);
            } else if (!((m[i >> 5] >> (i & 31)) & 1)) {
                atomicAdd(&S->hist[sym_lensym(s)], 1u);
                atomicAdd(&S->hist[288 + dist_sym(sym_dist(s))], 1u);
            } else {
                const uint8_t* p = v.out + v.symout[i];
                int len = sym_len(s);
                for (int k = 0; k < len; k++) atomicAdd(&S->hist[p[k]], 1u);
            }
        }
        __syncthreads();
        P1(PR_HIST);
    }
    // the same from the per-mask cache (every mask's histogram is derived from its parent's when it is created)
    __device__ __forceinline__ void load_hist(int mid) {
        const uint32_t* hs = hists + (size_t)mid * 320;
        __syncthreads();
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = hs[k];
        __syncthreads();
    }

    // payload of the symbol list described by S->hist under table t (recodeToHuffmanInternal, :759-770)
    __device__ __noinline__ long long hist_payload(const Tab& t) {
        P0();
        long long acc = 0;
        for (int k = tid; k < 318; k += ENG_NT) {
            uint32_t f = S->hist[k];
            if (!f) continue;
            int bits;
            if (k < 257) bits = t.L[k];
            else if (k < 286) bits = t.L[k] + len_ebits_of(k);
            else if (k >= 288) bits = t.D[k - 288] + dist_ebits_of(k - 288);
            else bits = 0;
            acc += (long long)f * bits;
        }
        if (tid == 0) S->red = 0;
        __syncthreads();
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if ((tid & 31) == 0 && acc) atomicAdd(&S->red, (unsigned long long)acc);
        __syncthreads();
        long long r = (long long)S->red;
        __syncthreads();
        P1(PR_PAYLOAD);
        return r;
    }

#ifdef D4_VERIFY
    int* vgerr = nullptr;
    int vjob = -1;
    // debug: payload of candidate c recomputed from its mask and tables; first mismatch is recorded
    __device__ __noinline__ void verify(int c, int opcode) {
        __syncthreads();
        // pass_hist_full/hist_payload clobber S->hist and S->red only
        pass_hist_full(S->c[c].mid);
        long long t = hist_payload(S->c[c].tab);
        if (tid == 0 && t != S->c[c].payload) {
            if (atomicMax(vgerr, 14) < 13) {
                vgerr[1] = vjob; vgerr[2] = opcode; vgerr[3] = c; vgerr[4] = (int)S->c[c].payload; vgerr[5] = (int)t;
                vgerr[6] = (int)blockIdx.x; vgerr[7] = (int)S->candIndex;
            }
        }
        __syncthreads();
    }
#endif

    // ---- recodeHuffman (:670-743) on candidate c: tables from the histogram, payload, default header.
    //      The result is a function of the symbol list alone -> cached per mask id.
    __device__ __noinline__ void op_recode(int c) {
        maybe_flush();
        P0();
        Cand& cd = S->c[c];
        const int mid = cd.mid;
        if (S->recodeValid[mid]) {
            const uint32_t* s = (const uint32_t*)&recode[mid];
            uint32_t* d = (uint32_t*)&S->c[c];
            __syncthreads();
            for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
            __syncthreads();
            P1(PR_RECODE_HIT);
            D4V(c, 5);
            return;
        }
        load_hist(mid);
        // trailing zero-frequency trimming + the distance special cases (:683-740)
        const long long p0_trees = clock64();
        if (tid == 0) {
            int nl = 286;
            while (nl > 0 && S->hist[nl - 1] == 0) nl--;
            cd.tab.nL = (uint16_t)nl;
            if (huff_tree<290, 584>(S->hist, nl, 15, cd.tab.L, S->tl)) S->err = ERR_TREE;
            for (int k = nl; k < MAX_LL; k++) cd.tab.L[k] = 0;
        }
        if (tid == 32) {
            const uint32_t* df = S->hist + 288;
            int nd = 30;
            while (nd > 0 && df[nd - 1] == 0) nd--;
            int nz = 0;
            for (int k = 0; k < nd; k++) nz += df[k] != 0;
            for (int k = 0; k < MAX_D; k++) cd.tab.D[k] = 0;
            if (nd == 0) { cd.tab.nD = 1; }                                   // handleZero: one entry, length 0
            else if (nz <= 1) { cd.tab.nD = (uint16_t)nd; cd.tab.D[nd - 1] = 1; }  // handleOne
            else {
                cd.tab.nD = (uint16_t)nd;
                if (huff_tree<32, 68>(df, nd, 15, cd.tab.D, S->td)) S->err = ERR_TREE;
            }
        }
        __syncthreads();
        { const long long p0_ = p0_trees; (void)p0_; P1(PR_TREES); }
        if (tid == 0) { cd.tab.type = 2; cd.tab.pad[0] = cd.tab.pad[1] = cd.tab.pad[2] = 0; }
        __syncthreads();
        long long pay = hist_payload(cd.tab);
        {
            P0();
            if (tid == 0) cd.payload = pay;
            hdr_default_parallel(c);
            P1(PR_HDR_DEFAULT);
        }
        intern_tab(c);
        {
            const uint32_t* s = (const uint32_t*)&S->c[c];
            uint32_t* d = (uint32_t*)&recode[mid];
            for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
            if (tid == 0) S->recodeValid[mid] = 1;
        }
        __syncthreads();
        P1(PR_RECODE_MISS);
        D4V(c, 6);
    }
    // recodeHuffmanLessMatches (:655-658)
    __device__ void op_recode_less(int c) { pass_replace(c, true); op_recode(c); }

    // recodeToFixedHuffman (:637-653); the fixed-code payload is a function of the symbol list alone
    __device__ __noinline__ void op_to_fixed(int c) {
        Cand& cd = S->c[c];
        if (cd.tab.type == 1) return;
        __syncthreads();  // every thread has read the type before thread 0 rewrites it below
        maybe_flush();
        P0();
        const int mid = cd.mid;
        const unsigned long long key = pm_key(mid, 0xFFFF, 4);
        if (tid == 0) {
            S->tmpIdx = pm_find(key);
            cd.tab.type = 1; cd.tab.nL = 286; cd.tab.nD = 30;
            cd.tab.pad[0] = cd.tab.pad[1] = cd.tab.pad[2] = 0;
            fixed_lens(cd.tab.L, cd.tab.D);
            for (int k = 286; k < MAX_LL; k++) cd.tab.L[k] = 0;
            for (int k = 30; k < MAX_D; k++) cd.tab.D[k] = 0;
            cd.hdr.np = 0; cd.hdr.ncl = 0; cd.hdr.bits = 0;
        }
        __syncthreads();
        int slot = S->tmpIdx;
        __syncthreads();
        if (slot >= 0) {
            if (tid == 0) cd.payload = pvals[slot].delta;
        } else {
            slot = -1 - slot;
            load_hist(mid);
            long long pay = hist_payload(cd.tab);
            if (tid == 0) {
                cd.payload = pay;
                PVal pv; pv.mid = (uint32_t)mid; pv.pad = 0; pv.delta = pay;
                pvals[slot] = pv;
                G->pkey[slot] = key;
                S->nP++;
            }
        }
        __syncthreads();
        if (S->fixedTab >= 0) {
            if (tid == 0) cd.tabid = (uint16_t)S->fixedTab;
            __syncthreads();
        } else {
            intern_tab(c);
            if (tid == 0) S->fixedTab = cd.tabid;
            __syncthreads();
        }
        P1(PR_TO_FIXED);
    }

    // DeflateBlockHuffman.optimise (:460-469): returns bits saved
    __device__ __noinline__ long long op_optimise(int c) {
        long long before = cand_size(S->c[c]);
        __syncthreads();
        pass_replace(c, false);
        P0();
        if (tid == 0 && S->c[c].tab.type == 2) hdr_optimise(S->c[c].hdr);
        __syncthreads();
        P1(PR_HDR_OPT);
        long long after = cand_size(S->c[c]);
        __syncthreads();
        return before - after;
    }
    // optimiseBlockNormal (DeflateStream.java:319-327): dst = copy(src).optimise(); returns saved > 0
    __device__ bool op_optimise_normal(int dst, int src) {
        copy(dst, src);
        return op_optimise(dst) > 0;
    }
    __device__ void op_recode_header(int c) {
        P0();
        if (S->c[c].tab.type == 2) hdr_recode_parallel(c);
        P1(PR_HDR_RECODE);
    }
    __device__ void op_recode_header_less(int c) {
        P0();
        if (S->c[c].tab.type == 2) {
            if (tid == 0) hdr_replace_runs(S->c[c].hdr, true);   // recodeHeaderToLessRLEMatches (:632-635)
            __syncthreads();
            hdr_recode_parallel(c);
        }
        P1(PR_HDR_RECODE);
    }

    // runs of equal code lengths of a Tab, cut by one warp: ballot + popcount compaction (same result as runlist_build)
    struct RunListW : RunList { uint16_t start[MAX_PAIRS]; };
    __device__ __forceinline__ void runlist_warp(const Tab& t, RunListW& rl, int lane) {
        const int nL = t.nL, n = t.nL + t.nD;
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int v = i < n ? (i < nL ? t.L[i] : t.D[i - nL]) : -1;
            const int pv = (i > 0 && i < n) ? (i - 1 < nL ? t.L[i - 1] : t.D[i - 1 - nL]) : -2;
            const bool st = i < n && v != pv;
            const unsigned bal = __ballot_sync(0xffffffffu, st);
            if (st) {
                const int k = cnt + __popc(bal & ((1u << lane) - 1u));
                rl.val[k] = (uint8_t)v;
                rl.start[k] = (uint16_t)i;
            }
            cnt += __popc(bal);
        }
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) rl.len[k] = (uint16_t)((k + 1 < cnt ? rl.start[k + 1] : n) - rl.start[k]);
        if (lane == 0) rl.n = (uint16_t)cnt;
    }

    // header code of the pair frequencies in S->hist[0..19) (Huffman.ofRLEPacked), trimmed on from `ncl`, and the
    // header size: sum of pair bits = sum_s freq[s] * CL[s] + 2 f16 + 3 f17 + 7 f18.  Thread 0 only.
    __device__ __forceinline__ void hdr_code_from_freq(Hdr& h, int ncl) {
        if (huff_tree<21, 46>(S->hist, 19, 7, h.CL, S->wsCL)) S->err = ERR_TREE;
        ncl = trim_ncl(h.CL, ncl);
        int bits = 5 + 5 + 4 + 3 * ncl + 2 * (int)S->hist[16] + 3 * (int)S->hist[17] + 7 * (int)S->hist[18];
        for (int k = 0; k < 19; k++) bits += (int)S->hist[k] * h.CL[k];
        h.ncl = (uint8_t)ncl;
        h.bits = bits;
    }

    // rewriteHeader with the default strategy (DeflateBlockHuffman.java:480-577) for candidate c, by the whole CTA:
    // runs -> pairs per run -> exclusive scan -> every run writes its pairs; same bytes as hdr_rewrite(FLAGS_DEFAULT)
    __device__ __noinline__ void hdr_default_parallel(int c) {
        struct HdrWs { RunListW rl; uint16_t off[MAX_PAIRS + 2]; };
        static_assert(sizeof(HdrWs) <= sizeof(S->tl), "overlays the litlen tree workspace (idle once the trees are built)");
        HdrWs& W = *reinterpret_cast<HdrWs*>(&S->tl);
        Cand& cd = S->c[c];
        const int w = tid >> 5, lane = tid & 31;
        if (tid < 19) S->hist[tid] = 0;
        if (w == 0) runlist_warp(cd.tab, W.rl, lane);
        __syncthreads();
        const int R = W.rl.n;
        for (int r = tid; r < R; r += ENG_NT) {
            int cnt = 0;
            emit_run(W.rl.val[r], W.rl.len[r], FLAGS_DEFAULT, [&](int, int, int, int k) { cnt += k; });
            W.off[r] = (uint16_t)cnt;
        }
        __syncthreads();
        if (w == 0) {
            int carry = 0;
            for (int i0 = 0; i0 < R; i0 += 32) {
                const int i = i0 + lane;
                const int x = i < R ? W.off[i] : 0;
                int incl = x;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                if (i < R) W.off[i] = (uint16_t)(carry + incl - x);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) W.off[R] = (uint16_t)carry;
        }
        __syncthreads();
        for (int r = tid; r < R; r += ENG_NT) {
            int o = W.off[r];
            emit_run(W.rl.val[r], W.rl.len[r], FLAGS_DEFAULT, [&](int sym, int run, int val, int k) {
                const uint16_t p = pair_pack(sym, run, val);
                for (int q = 0; q < k; q++) cd.hdr.pairs[o++] = p;
                atomicAdd(&S->hist[sym], (uint32_t)k);
            });
        }
        __syncthreads();
        if (tid == 0) {
            cd.hdr.np = W.off[R];
            hdr_code_from_freq(cd.hdr, 19);
        }
        __syncthreads();
    }

    // recodeHeader (:579-629) for candidate c: pair frequencies by all threads, the code by thread 0
    __device__ __noinline__ void hdr_recode_parallel(int c) {
        Cand& cd = S->c[c];
        if (tid < 19) S->hist[tid] = 0;
        __syncthreads();
        const int np = cd.hdr.np;
        for (int i = tid; i < np; i += ENG_NT) atomicAdd(&S->hist[pair_sym(cd.hdr.pairs[i])], 1u);
        __syncthreads();
        if (tid == 0) hdr_code_from_freq(cd.hdr, cd.hdr.ncl);
        __syncthreads();
    }

    // ---- the 56 header strategy trials of up to 4 bases (addOptimisedRecoded, :277-316) -------------
    // bases are candidates C_B1.. (nb of them).  For each base the first-minimum strategy is looked up
    // in / added to the per-tabid memo, then the virtual candidates are fed to the selection callback in
    // the reference's order; only a winning trial is materialised.
    __device__ __noinline__ void trials(int nb) {
        P0();
        __shared__ int s_miss[4];
        if (tid < nb) s_miss[tid] = (G->tabTrialBits[S->c[C_B1 + tid].tabid] == TRIAL_UNSET) || g_trace != nullptr;
        __syncthreads();
        // Evaluate the misses.  Sizes only (huff.cuh trial_sizes): warp b cuts base b's code lengths into runs of equal
        // values (ballot + popcount compaction into shared memory, overlaying the litlen tree workspace, which is idle
        // here); then thread j = 28 b + c evaluates rewrite strategy c of base b for both prune values, regenerating
        // the RLE pair stream from the runs instead of storing it.
        {
            static_assert(4 * sizeof(RunListW) <= sizeof(S->tl), "run lists overlay the litlen tree workspace");
            RunListW* rls = reinterpret_cast<RunListW*>(&S->tl);
            const int w = tid >> 5, lane = tid & 31;
            for (int b = w; b < nb; b += ENG_NT / 32) {   // warp-uniform
                if (!s_miss[b]) continue;
                runlist_warp(S->c[C_B1 + b].tab, rls[b], lane);
            }
            __syncthreads();
            for (int j = tid; j < nb * 28; j += ENG_NT) {
                const int b = j / 28, c = j % 28;
                if (!s_miss[b]) continue;
                // c_trial_flags order: [0,20) and [40,48) are the rewrite strategies without prune, +20 / +8 with it
                const int kF = c < 20 ? c : 40 + (c - 20), kT = c < 20 ? 20 + c : 48 + (c - 20);
                TreeWsCL ws;
                int a = 0, bp = 0;
                if (trial_sizes(rls[b], c_trial_flags[kF], &a, &bp, ws)) S->err = ERR_TREE;
                S->trialBits[b * 56 + kF] = a;
                S->trialBits[b * 56 + kT] = bp;
            }
        }
        __syncthreads();
        P1(PR_TRIALS_EVAL);
        if (g_trace && tid == 0)
            for (int b = 0; b < nb; b++)
                for (int k = 0; k < 56; k++) trace_put(S->candIndex + b * 56 + k, S->c[C_B1 + b].payload + S->trialBits[b * 56 + k]);
        __syncthreads();
        if (tid < nb && s_miss[tid]) {
            int best = 0x7fffffff, arg = 0;
            for (int k = 0; k < 56; k++) { int bts = S->trialBits[tid * 56 + k]; if (bts < best) { best = bts; arg = k; } }
            const int t = S->c[C_B1 + tid].tabid;   // two bases with one Tab write the same values
            G->tabTrialBits[t] = best;
            G->tabTrialArg[t] = (unsigned char)arg;
        }
        __syncthreads();
        // selection in reference order: base b's 56 candidates; the first minimum is the only one that
        // can replace the incumbent
        for (int b = 0; b < nb; b++) {
            const int t = S->c[C_B1 + b].tabid;
            const int bits = G->tabTrialBits[t], arg = G->tabTrialArg[t];
            const long long sz = S->c[C_B1 + b].payload + bits;
            const bool better = sz < S->bestSize;
            __syncthreads();
            if (tid == 0) {
                if (sz < S->restMin) S->restMin = sz;
                if (better) { S->bestSize = sz; S->bestStored = 0; S->bestIndex = S->candIndex + arg; }
                S->candIndex += 56;
            }
            if (better) {
                copy(C_BEST, C_B1 + b);
                if (tid == 0) {
                    if (hdr_trial(S->c[C_BEST].tab, c_trial_flags[arg], S->c[C_BEST].hdr, S->wsCL)) S->err = ERR_TREE;
                }
            }
            __syncthreads();
        }
        P1(PR_TRIALS);
    }

    // recodedHuffmanFull (DeflateStream.java:212-229): cur (slot a) is replaced while a further
    // recodeHuffmanLessMatches shrinks it.  Returns true when the result differs from the start
    // (the reference's `prunedFull != pruned` identity test).
    __device__ __noinline__ bool op_recoded_full(int a, int tmp) {
        bool changed = false;
        while (true) {
            copy(tmp, a);
            op_recode_less(tmp);
            long long s1 = cand_size(S->c[tmp]), s0 = cand_size(S->c[a]);
            __syncthreads();
            if (s1 >= s0) break;
            copy(a, tmp);
            changed = true;
        }
        return changed;
    }

    // addOptimisedRecoded (DeflateStream.java:265-317) for base block y
    __device__ __noinline__ int aor(int y) {
        // The four bases are only ever read by trials(): their Tab and payload.  Every trial rewrites the header from
        // the Tab (optimiseBlockDynBlock -> rewriteHeader, DeflateStream.java:184-198), so the header half of
        // DeflateBlockHuffman.optimise() (optimiseHeader, :471-476) cannot influence any candidate here and is not run.
        copy(C_B1, y); pass_replace(C_B1, false);                      // optimiseBlockCopyHelper(toOptimise)
        copy(C_B2, y); op_recode(C_B2); pass_replace(C_B2, false);     // optimiseBlockHelper(recodedHuffman(.., false))
        copy(C_PP, y); op_recode_less(C_PP);                // pruned
        copy(C_B3, C_PP); pass_replace(C_B3, false);                   // optimiseBlockCopyHelper(pruned)
        copy(C_B4, C_PP);
        bool full = op_recoded_full(C_B4, C_CHK2);          // prunedFull
        if (full) pass_replace(C_B4, false);
        trials(full ? 4 : 3);
        return full ? 4 : 3;
    }

    // runOptimisationsCallback (DeflateStream.java:400-442) for block x
    __device__ __noinline__ void run(int x) {
        copy(C_T, x); op_recode_header(C_T); cb(C_T);                 // post
        if (op_optimise_normal(C_TMP, C_T)) cb(C_TMP);                // post optimised
        const int nb = aor(C_T);
        copy(C_T, x); op_recode_header_less(C_T); cb(C_T);            // pruned header
        if (op_optimise_normal(C_TMP, C_T)) cb(C_TMP);
        if (tid == 0) S->candIndex += 56 * nb;                        // aor(prune): see file header
        __syncthreads();
        copy(C_Y, x); pass_least(C_Y, 0); aor(C_Y);
        copy(C_Y, x); pass_least(C_Y, 1); aor(C_Y);
    }

    // runOptimisationsCallbackMulti (DeflateStream.java:443-463) for seed e
    __device__ __noinline__ void multi(int e) {
        cb(e); run(e);
        copy(C_X, e); op_recode(C_X); cb(C_X); run(C_X);
        copy(C_X, e); op_recode_less(C_X); cb(C_X); run(C_X);
        bool full = op_recoded_full(C_X, C_CHK);
        if (full) { cb(C_X); run(C_X); }
    }

    // DeflateStream.optimiseBlock (:343-490) for the Huffman block in C_B.  storedSize < 0: no stored
    // candidate is compared (phase A resolves it afterwards from sizeI / sizeC1 / restMin).
    // Result: C_BEST (or stored when S->bestStored).
    __device__ void optimise_block(long long storedSize) {
        P0();
        if (tid == 0) {
            if (g_trace) trace_put(-1, cand_size(S->c[C_B]));
            S->bestSize = cand_size(S->c[C_B]);
            S->bestStored = 0;
            S->sizeI = S->bestSize;
            S->sizeC1 = S->bestSize;
            S->restMin = 0x7fffffffffffffffll;
            S->candIndex = 0;
            S->bestIndex = 0xffffffffu;
        }
        __syncthreads();
        copy(C_BEST, C_B);
        const int type = S->c[C_B].tab.type;
        bool hasO = op_optimise_normal(C_O, C_B);
        if (hasO) {
            cb(C_O, false);
            if (tid == 0) S->sizeC1 = cand_size(S->c[C_O]);
            __syncthreads();
        }
        if (v.ulen <= 65535) {
            if (storedSize >= 0 && storedSize < S->bestSize) {
                __syncthreads();
                if (tid == 0) { S->bestSize = storedSize; S->bestStored = 1; S->bestIndex = S->candIndex; }
            }
            __syncthreads();
            if (tid == 0) S->candIndex++;
            __syncthreads();
        }
        int H = C_B;
        bool hasOh = hasO;
        if (type == 1) {
            copy(C_H, C_B); op_recode(C_H);
            H = C_H;
            hasOh = op_optimise_normal(C_O, C_H);
        }
        multi(H);
        if (hasOh) multi(C_O);
        if (type != 1) {
            copy(C_E, H); op_to_fixed(C_E); op_optimise(C_E); cb(C_E);
        }
        copy(C_E, H); pass_least(C_E, 0); multi(C_E);
        copy(C_E, H); pass_least(C_E, 1); multi(C_E);
        P1(PR_BLOCK);
    }
};

}  // namespace d4
