// engine.cuh — the block candidate enumerator (DeflateStream.optimiseBlock, DeflateStream.java:343-490)
// as CTA-cooperative device code.  One CTA owns one block; every function below is called by ALL
// threads of the CTA with uniform arguments.
//
// A candidate ("Cand") is the reference's DeflateBlockHuffman copy reduced to what determines its bytes:
//   Tab (code lengths + table lengths + type), Hdr (header RLE pairs, header code, numCodelenLens),
//   payload = litlenSizeBits, and a bit mask over the block's symbols (1 = match replaced by literals).
// Candidate i's mask lives in mask slot i of the CTA's global scratch.
//
// O(n) work (n = symbols of the block) is done by CTA-wide passes:
//   pass_replace : replaceBackrefsWithLiteralsIfSmaller (DeflateBlockHuffman.java:222-319)
//   pass_least   : removeDistLitLeastExpensive          (:373-458)
//   pass_hist    : the histogram loop of recodeHuffman   (:671-681); payload = hist . (len + extra bits)
// Small serial work (Huffman trees, header model) runs in single threads (huff.cuh).
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

constexpr int ENG_NT = 256;
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
constexpr int MAXM = 20, MAXT = 20, MEMO_P = 64;
#else
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
__device__ __forceinline__ long long cand_size(const Cand& c) { return c.payload + (c.tab.type == 2 ? c.hdr.bits : 0); }

struct BlkView {
    const uint32_t* sym;
    const uint32_t* symout;
    const uint8_t* out;
    uint32_t n;       // symbols (including a NOP left by a merge)
    uint32_t nwords;  // mask words
    uint64_t ulen;    // decoded length
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
    // selection state
    long long bestSize;
    int bestStored;
    long long sizeI, sizeC1, restMin;
    unsigned candIndex, bestIndex;
    // pools and memo tables
    unsigned long long maskHash[MAXM];
    unsigned char recodeValid[MAXM];
    int nMasks;
    unsigned long long tabHash[MAXT];
    int tabTrialBits[MAXT];
    unsigned char tabTrialArg[MAXT];
    int nTabs, fixedTab;
    unsigned long long pkey[MEMO_P];
    int nP;
    int remap[NCAND], uniq[NCAND];
    unsigned long long uh[NCAND];
    int tmpIdx, redAny2;
    int trialBits[4 * 56];
    unsigned long long hred[ENG_NT / 32];
};

enum { C_B = 0, C_BEST, C_O, C_H, C_E, C_X, C_CHK, C_T, C_Y, C_B1, C_B2, C_B3, C_B4, C_PP, C_CHK2, C_TMP };

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
    int tid;

    __device__ uint32_t* maskp(int id) const { return masks + (size_t)id * maxwords; }

    // ---- candidate copy (DeflateBlockHuffman.copy, :1174-1205, with value semantics; masks are immutable
    //      pool entries, so a copy shares its source's mask id) --------------------------------------------
    __device__ __noinline__ void copy(int dst, int src) {
        if (dst == src) return;
        const uint32_t* s = (const uint32_t*)&S->c[src];
        uint32_t* d = (uint32_t*)&S->c[dst];
        for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
        __syncthreads();
        D4V(dst, 7);
    }

    // ---- selection callback (DeflateStream.java:349-368) ------------------------------------------
    __device__ __noinline__ void cb(int c, bool isRest = true) {
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
        for (int k = tid; k < MEMO_P; k += ENG_NT) S->pkey[k] = 0;
        for (int k = tid; k < MAXM; k += ENG_NT) S->recodeValid[k] = 0;
        if (tid < NCAND) { S->c[tid].mid = 0; S->c[tid].tabid = 0; S->c[tid].pad2 = 0; }
        if (tid == 0) { S->nMasks = 0; S->nTabs = 0; S->nP = 0; S->fixedTab = -1; }
        __syncthreads();
    }

    // Tab of candidate c -> S->c[c].tabid (hash-consed; a hash hit is confirmed by a full comparison)
    __device__ __noinline__ void intern_tab(int c) {
        const uint32_t* q = (const uint32_t*)&S->c[c].tab;
        const unsigned long long h = hash_words(q, (int)(sizeof(Tab) / 4));
        if (tid == 0) { S->tmpIdx = -1; S->redAny2 = 0; }
        __syncthreads();
        const int nT = S->nTabs;
        for (int k = tid; k < nT; k += ENG_NT)
            if (S->tabHash[k] == h) atomicMax(&S->tmpIdx, k);
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
                S->tabHash[slot] = h;
                S->tabTrialBits[slot] = TRIAL_UNSET;
                S->tmpIdx = slot;
            }
            __syncthreads();
            hit = S->tmpIdx;
            uint32_t* a = (uint32_t*)&tabs[hit];
            for (int k = tid; k < (int)(sizeof(Tab) / 4); k += ENG_NT) a[k] = q[k];
        }
        if (tid == 0) S->c[c].tabid = (uint16_t)hit;
        __syncthreads();
    }

    // the mask just written into pool slot nMasks -> its id (an equal older mask wins, so equal symbol lists
    // reached along different paths share their memo entries)
    __device__ __noinline__ int intern_mask() {
        const int fresh = S->nMasks;
        const uint32_t* q = maskp(fresh);
        const unsigned long long h = hash_words(q, (int)v.nwords);
        if (tid == 0) { S->tmpIdx = -1; S->redAny2 = 0; }
        __syncthreads();
        for (int k = tid; k < fresh; k += ENG_NT)
            if (S->maskHash[k] == h) atomicMax(&S->tmpIdx, k);
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
            if (tid == 0) { S->maskHash[fresh] = h; S->recodeValid[fresh] = 0; S->nMasks = fresh + 1; }
        }
        __syncthreads();
        return hit;
    }

    // a pool is full: keep only what the candidate slots reference
    __device__ __noinline__ void flush_all() {
        __syncthreads();
        if (tid == 0) {
            int nu = 0;
            for (int c = 0; c < NCAND; c++) {
                const int m = S->c[c].mid;
                int k = 0;
                while (k < nu && S->uniq[k] != m) k++;
                if (k == nu) { S->uniq[nu] = m; S->uh[nu] = S->maskHash[m]; nu++; }
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
        }
        __syncthreads();
        for (int k = 0; k < nu; k++) {
            const uint32_t* s = maskp(MAXM + k);
            uint32_t* d = maskp(k);
            for (uint32_t w = tid; w < v.nwords; w += ENG_NT) d[w] = s[w];
        }
        if (tid < NCAND) S->c[tid].mid = (uint16_t)S->remap[tid];
        if (tid < nu) S->maskHash[tid] = S->uh[tid];
        for (int k = tid; k < MAXM; k += ENG_NT) S->recodeValid[k] = 0;
        for (int k = tid; k < MEMO_P; k += ENG_NT) S->pkey[k] = 0;
        __syncthreads();
        if (tid == 0) { S->nMasks = nu; S->nTabs = 0; S->nP = 0; S->fixedTab = -1; }
        __syncthreads();
        for (int c = 0; c < NCAND; c++) intern_tab(c);
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
            const unsigned long long k = S->pkey[h];
            if (k == key) return (int)h;
            if (k == 0) return -1 - (int)h;
            h = (h + 1) & (MEMO_P - 1);
        }
    }
    __device__ __forceinline__ unsigned long long pm_key(int mid, int tabid, int op) const {
        return (1ull << 63) | ((unsigned long long)op << 32) | ((unsigned long long)tabid << 16) | (unsigned long long)mid;
    }

    // ---- CTA-wide passes ----------------------------------------------------------------------------
    // literal cost of the bytes a match produces; returns -1 when a byte has no code
    __device__ __forceinline__ int lit_cost(const uint8_t* L, uint32_t off, int len) const {
        int tot = 0;
        const uint8_t* p = v.out + off;
        for (int k = 0; k < len; k++) {
            int b = L[p[k]];
            if (b < 1) return -1;
            tot += b;
        }
        return tot;
    }
    // getLitLenSize for a match (:112-128)
    __device__ __forceinline__ int ref_cost(const Tab& t, uint32_t s) const {
        int ls = sym_lensym(s), ds = dist_sym(sym_dist(s));
        return t.L[ls] + len_ebits_of(ls) + t.D[ds] + dist_ebits_of(ds);
    }

    // replaceBackrefsWithLiteralsIfSmaller(prune) on candidate c (in place)
    __device__ __noinline__ void pass_replace(int c, bool prune) {
        maybe_flush();
        Cand& cd = S->c[c];
        const int mid = cd.mid;
        const unsigned long long key = pm_key(mid, cd.tabid, prune ? 1 : 0);
        if (tid == 0) { S->tmpIdx = pm_find(key); S->red = 0; S->redAny = 0; }
        __syncthreads();
        int slot = S->tmpIdx;
        if (slot >= 0) {
            if (tid == 0) { const PVal pv = pvals[slot]; cd.mid = (uint16_t)pv.mid; cd.payload -= pv.delta; }
            __syncthreads();
            D4V(c, prune ? 2 : 1);
            return;
        }
        slot = -1 - slot;
        const uint32_t* m = maskp(mid);
        uint32_t* md = maskp(S->nMasks);
        long long saved = 0;
        const int lane = tid & 31;
        for (uint32_t base = (tid >> 5) * 32; base < v.n; base += ENG_NT) {
            uint32_t i = base + lane;
            bool rep = false;
            uint32_t word = m[base >> 5];
            if (i < v.n) {
                uint32_t s = v.sym[i];
                if (sym_is_match(s) && !((word >> lane) & 1)) {
                    int ref = ref_cost(cd.tab, s);
                    int lit = lit_cost(cd.tab.L, v.symout[i], sym_len(s));
                    if (lit >= 0 && (prune ? lit <= ref : lit < ref)) { rep = true; saved += ref - lit; }
                }
            }
            unsigned bal = __ballot_sync(0xffffffffu, rep);
            if (lane == 0) {
                md[base >> 5] = word | bal;
                if (bal) S->redAny = 1;
            }
        }
        for (int d = 16; d > 0; d >>= 1) saved += __shfl_xor_sync(0xffffffffu, saved, d);
        if (lane == 0 && saved) atomicAdd(&S->red, (unsigned long long)saved);
        __syncthreads();
        int newmid = mid;
        if (S->redAny) newmid = intern_mask();
        if (tid == 0) {
            PVal pv; pv.mid = (uint32_t)newmid; pv.pad = 0; pv.delta = (long long)S->red;
            pvals[slot] = pv;
            S->pkey[slot] = key;
            S->nP++;
            cd.mid = (uint16_t)newmid;
            cd.payload -= pv.delta;
        }
        __syncthreads();
        D4V(c, prune ? 2 : 1);
    }

    // removeDistLitLeastExpensive(mode) on candidate c (in place); no-op unless DYNAMIC
    __device__ __noinline__ void pass_least(int c, int mode) {
        Cand& cd = S->c[c];
        if (cd.tab.type != 2) return;
        maybe_flush();
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
            D4V(c, 3 + mode);
            return;
        }
        slot = -1 - slot;
        const uint32_t* m = maskp(mid);
        for (uint32_t i = tid; i < v.n; i += ENG_NT) {
            uint32_t s = v.sym[i];
            if (sym_is_match(s) && !((m[i >> 5] >> (i & 31)) & 1)) {
                int bin = sym_lensym(s) - 257;
                atomicOr(&S->leastSeen, 1u << bin);
                int lit = lit_cost(cd.tab.L, v.symout[i], sym_len(s));
                if (lit < 0) atomicOr(&S->leastBlocked, 1u << bin);
                else { atomicAdd(&S->leastSum[bin], lit - ref_cost(cd.tab, s)); atomicAdd(&S->leastCnt[bin], 1); }
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
            uint32_t* md = maskp(S->nMasks);
            const int lane = tid & 31;
            for (uint32_t base = (tid >> 5) * 32; base < v.n; base += ENG_NT) {
                uint32_t i = base + lane;
                bool rep = false;
                if (i < v.n) {
                    uint32_t s = v.sym[i];
                    rep = sym_is_match(s) && (sym_lensym(s) - 257 == rem);
                }
                unsigned bal = __ballot_sync(0xffffffffu, rep);
                if (lane == 0) md[base >> 5] = m[base >> 5] | bal;
            }
            __syncthreads();
            newmid = intern_mask();
        }
        if (tid == 0) {
            PVal pv; pv.mid = (uint32_t)newmid; pv.pad = 0; pv.delta = -(long long)S->red;
            pvals[slot] = pv;
            S->pkey[slot] = key;
            S->nP++;
            cd.mid = (uint16_t)newmid;
            cd.payload -= pv.delta;
        }
        __syncthreads();
        D4V(c, 3 + mode);
    }

    // histogram of the symbol list with mask `mid` into S->hist
    __device__ __noinline__ void pass_hist(int mid) {
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
    }

    // payload of the symbol list described by S->hist under table t (recodeToHuffmanInternal, :759-770)
    __device__ __noinline__ long long hist_payload(const Tab& t) {
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
        return r;
    }

#ifdef D4_VERIFY
    int* vgerr = nullptr;
    int vjob = -1;
    // debug: payload of candidate c recomputed from its mask and tables; first mismatch is recorded
    __device__ __noinline__ void verify(int c, int opcode) {
        __syncthreads();
        // pass_hist/hist_payload clobber S->hist and S->red only
        pass_hist(S->c[c].mid);
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
        Cand& cd = S->c[c];
        const int mid = cd.mid;
        if (S->recodeValid[mid]) {
            const uint32_t* s = (const uint32_t*)&recode[mid];
            uint32_t* d = (uint32_t*)&S->c[c];
            __syncthreads();
            for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
            __syncthreads();
            D4V(c, 5);
            return;
        }
        pass_hist(mid);
        // trailing zero-frequency trimming + the distance special cases (:683-740)
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
        if (tid == 0) { cd.tab.type = 2; cd.tab.pad[0] = cd.tab.pad[1] = cd.tab.pad[2] = 0; }
        __syncthreads();
        long long pay = hist_payload(cd.tab);
        if (tid == 0) {
            cd.payload = pay;
            TreeWsCL ws;
            if (hdr_rewrite(cd.tab, FLAGS_DEFAULT, cd.hdr, ws)) S->err = ERR_TREE;
        }
        __syncthreads();
        intern_tab(c);
        {
            const uint32_t* s = (const uint32_t*)&S->c[c];
            uint32_t* d = (uint32_t*)&recode[mid];
            for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
            if (tid == 0) S->recodeValid[mid] = 1;
        }
        __syncthreads();
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
            pass_hist(mid);
            long long pay = hist_payload(cd.tab);
            if (tid == 0) {
                cd.payload = pay;
                PVal pv; pv.mid = (uint32_t)mid; pv.pad = 0; pv.delta = pay;
                pvals[slot] = pv;
                S->pkey[slot] = key;
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
    }

    // DeflateBlockHuffman.optimise (:460-469): returns bits saved
    __device__ __noinline__ long long op_optimise(int c) {
        long long before = cand_size(S->c[c]);
        __syncthreads();
        pass_replace(c, false);
        if (tid == 0 && S->c[c].tab.type == 2) hdr_optimise(S->c[c].hdr);
        __syncthreads();
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
        if (tid == 0 && S->c[c].tab.type == 2) { TreeWsCL ws; if (hdr_recode(S->c[c].hdr, ws)) S->err = ERR_TREE; }
        __syncthreads();
    }
    __device__ void op_recode_header_less(int c) {
        if (tid == 0 && S->c[c].tab.type == 2) { TreeWsCL ws; if (hdr_recode_less(S->c[c].hdr, ws)) S->err = ERR_TREE; }
        __syncthreads();
    }

    // ---- the 56 header strategy trials of up to 4 bases (addOptimisedRecoded, :277-316) -------------
    // bases are candidates C_B1.. (nb of them).  For each base the first-minimum strategy is looked up
    // in / added to the per-tabid memo, then the virtual candidates are fed to the selection callback in
    // the reference's order; only a winning trial is materialised.
    __device__ __noinline__ void trials(int nb) {
        __shared__ int s_miss[4];
        if (tid < nb) s_miss[tid] = (S->tabTrialBits[S->c[C_B1 + tid].tabid] == TRIAL_UNSET) || g_trace != nullptr;
        __syncthreads();
        // evaluate the misses: thread j -> (base j / 56, strategy j % 56)
        {
            int j = tid;
            if (j < nb * 56 && s_miss[j / 56]) {
                Hdr h;
                TreeWsCL ws;
                if (hdr_trial(S->c[C_B1 + j / 56].tab, c_trial_flags[j % 56], h, ws)) S->err = ERR_TREE;
                S->trialBits[j] = h.bits;
            }
        }
        __syncthreads();
        if (g_trace && tid == 0)
            for (int b = 0; b < nb; b++)
                for (int k = 0; k < 56; k++) trace_put(S->candIndex + b * 56 + k, S->c[C_B1 + b].payload + S->trialBits[b * 56 + k]);
        __syncthreads();
        if (tid < nb && s_miss[tid]) {
            int best = 0x7fffffff, arg = 0;
            for (int k = 0; k < 56; k++) { int bts = S->trialBits[tid * 56 + k]; if (bts < best) { best = bts; arg = k; } }
            const int t = S->c[C_B1 + tid].tabid;   // two bases with one Tab write the same values
            S->tabTrialBits[t] = best;
            S->tabTrialArg[t] = (unsigned char)arg;
        }
        __syncthreads();
        // selection in reference order: base b's 56 candidates; the first minimum is the only one that
        // can replace the incumbent
        for (int b = 0; b < nb; b++) {
            const int t = S->c[C_B1 + b].tabid;
            const int bits = S->tabTrialBits[t], arg = S->tabTrialArg[t];
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
                    TreeWsCL ws;
                    if (hdr_trial(S->c[C_BEST].tab, c_trial_flags[arg], S->c[C_BEST].hdr, ws)) S->err = ERR_TREE;
                }
            }
            __syncthreads();
        }
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
        op_optimise_normal(C_B1, y);                       // optimiseBlockCopyHelper(toOptimise)
        copy(C_B2, y); op_recode(C_B2); op_optimise(C_B2);  // optimiseBlockHelper(recodedHuffman(.., false))
        copy(C_PP, y); op_recode_less(C_PP);                // pruned
        op_optimise_normal(C_B3, C_PP);                     // optimiseBlockCopyHelper(pruned)
        copy(C_B4, C_PP);
        bool full = op_recoded_full(C_B4, C_CHK2);          // prunedFull
        if (full) op_optimise(C_B4);
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
    }
};

}  // namespace d4
