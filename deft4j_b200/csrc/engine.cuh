// engine.cuh — the block candidate engine: executes what the symbolic enumerator (enum.cuh) asks for.
//
// One CTA owns one block.  Thread 0 sweeps the enumeration of DeflateStream.optimiseBlock (DeflateStream.java:343-490)
// over memo tables in shared memory; every miss becomes a request.  After a sweep the whole CTA computes the requests:
//   passes   replaceBackrefsWithLiteralsIfSmaller / removeDistLitLeastExpensive (DeflateBlockHuffman.java:222-319,
//            373-458): CTA-wide, 8 symbols per thread, against a per-table cost array dc[i] = literal cost - match cost.
//            The cost array of a table is built in ONE coalesced pass over the block's decoded bytes: 2 KiB tiles, per-byte
//            code lengths looked up in shared memory, a CTA-wide prefix sum per tile, every match takes P[end] - P[start].
//   recodes  recodeHuffman (:670-743): one WARP per request - histogram row -> litlen and distance Huffman trees with the
//            exact java.util.PriorityQueue mechanics (lane 0; the trees of up to 8 requests run side by side) -> payload
//            (dot product) -> default header (runs -> pairs -> header code), all inside the warp.
//   headers  recodeHeader / recodeHeaderToLessRLEMatches / optimiseHeader (:471-476,579-635): one warp per request.
//   trials   the 56 header strategies of a table (DeflateStream.java:277-316): 28 threads per table, 8 tables at a time,
//            sizes only (huff.cuh trial_sizes); the winner alone is materialised.
// Masks (symbol lists), Tabs and Hdrs are immutable entries of per-CTA pools in global scratch, hash-consed where it
// pays (masks, Tabs); a pool that fills up makes the round restart in segments with a reset between them.
#pragma once
#include "enum.cuh"
#include "parse.cuh"

namespace d4 {

#ifndef D4_ENG_NT
#define D4_ENG_NT 256
#endif
#ifndef D4_ENG_MINB
#define D4_ENG_MINB 4
#endif
constexpr int ENG_NT = D4_ENG_NT;
constexpr int ENG_NW = ENG_NT / 32;
constexpr int ERR_TREE = 11, ERR_ROUNDS = 12, ERR_WRITER = 2, ERR_POOL = 15, ERR_INTERNAL = 16;  // reported through gerr

// parity-debug instrumentation (deft4cu_debug_trace): when armed, every candidate the selection callback sees is logged
// as (candidate index, size)
__device__ long long* g_trace = nullptr;
__device__ unsigned g_trace_cap = 0;
__device__ unsigned g_trace_n = 0;

struct Cand {               // a materialised candidate (BlkState, the records of the current and the best block)
    Tab tab;
    Hdr hdr;
    long long payload;      // litlenSizeBits
    uint16_t mid, tabid;    // unused outside the engine
    uint32_t pad2;
};
__device__ __forceinline__ long long cand_size(const Cand& c) { return c.payload + (c.tab.type == 2 ? c.hdr.bits : 0); }

struct BlkView {
    const uint32_t* sym;
    const uint32_t* symout;
    const uint8_t* out;
    uint32_t n;       // symbols (including NOPs left by a merge)
    uint32_t nwords;  // mask words
    uint64_t ulen;    // decoded length
    uint64_t out_off; // pool offset of the block's first decoded byte
};

constexpr int DC_TILE = 2048;     // decoded bytes per cost-array tile (8 per thread)
constexpr int DCN = 8;            // cost arrays kept per block (round-robin eviction)
constexpr int WS_BYTES = 3072;    // per-warp workspace for trees / header work
constexpr int SLOT_B = MAXM, SLOT_BEST = MAXM + 1;   // extra mask / histogram slots: the records of B and of the winner
static_assert(DC_TILE == ENG_NT * 8, "one 8-byte load per thread and tile");

// cycle accounting per engine phase (-DD4_PROF builds only; read back with deft4cu_debug_prof): thread 0's clock64
// deltas, [cat] = cycles, [32 + cat] = calls
#ifdef D4_PROF
__device__ unsigned long long g_prof[64];
#define P0() const long long p0_ = clock64()
#define P1(cat) do { if (threadIdx.x == 0) { atomicAdd(&g_prof[cat], (unsigned long long)(clock64() - p0_)); atomicAdd(&g_prof[32 + (cat)], 1ull); } } while (0)
#define PCOUNT(cat, k) do { if (threadIdx.x == 0) atomicAdd(&g_prof[32 + (cat)], (unsigned long long)(k)); } while (0)
#else
#define P0()
#define P1(cat)
#define PCOUNT(cat, k)
#endif
enum { PR_BLOCK = 0, PR_ROUND, PR_SWEEP, PR_SELECT, PR_PASS, PR_DC, PR_RECODE, PR_TREES, PR_HDR_DEFAULT, PR_HDROP, PR_TRIALS,
       PR_LOAD, PR_MATERIAL, PR_REBASE, PR_INTERN_MASK, PR_INTERN_TAB, PR_FIXED, PR_SLOWTREE, PR_SEGMENTED, PR_HIST };

struct EngSmem {
    SymState sym;
    Enumer en;
    TraceSink tsink;
    unsigned long long maskHash[MAXM];
    uint32_t hist[320];          // pass histogram delta / block histogram; [0,19) header pair frequencies
    int leastSum[32], leastCnt[32];
    unsigned leastBlocked, leastSeen;
    unsigned long long red;
    unsigned long long hred[ENG_NW];
    long long recPay[ENG_NW];
    uint32_t wt[ENG_NW];
    int redAny, redAny2, tmpIdx, err;
    int sweepDone, segImproved, segmentedRound;
    int carryIdx[2], carryRef[2];   // cost-array build: the match that straddles a tile boundary (double buffered by tile parity)
    uint32_t carryPart[2];
    unsigned char tabDc[MAXT];   // tabid -> cost-array slot (0xFF: none)
    unsigned short dcOwner[DCN]; // slot -> tabid (0xFFFF: free)
    int dcNext;
    uint8_t ctab[256 + 32 + 32]; // cost-array build: literal lengths, length-symbol lengths, distance lengths
    union {
        unsigned char ws[ENG_NW][WS_BYTES];
        uint32_t P[DC_TILE + 1]; // cost-array build: packed (cost | uncodable count << 16) exclusive prefix of one tile
        struct { Hdr hdr; TreeWsCL ws; } mat;   // winner materialisation (thread 0)
    } u;
};

struct EngScratch {       // global scratch, one slice per CTA
    uint32_t* masks;      // (MAXM + 2) * maxwords
    Tab* tabs;            // MAXT + ENG_NW (staging)
    Hdr* hdrs;            // MAXH
    uint32_t* hists;      // (MAXM + 2) * 320
    unsigned long long* tabHash;   // MAXT
    short* dc;            // DCN * maxwords * 32
    uint8_t* kind;        // maxwords * 32
    uint32_t* minfo;      // maxwords * 32
    uint32_t* tileFirst;  // maxtiles
    int* trialAll;        // MAXT * 56
    Cand* recs;           // 2: B and the winner
    TreeWs<290, 584>* slowWs;   // ENG_NW
    uint32_t maxwords, maxtiles;
};

struct Eng {
    EngSmem* S;
    BlkView v;
    uint32_t* masks;
    uint32_t maxwords, maxn;
    Tab* tabs;
    Hdr* hdrs;
    uint32_t* hists;
    unsigned long long* tabHash;
    short* dc;
    uint8_t* kind;
    uint32_t* minfo;
    uint32_t* tileFirst;
    int* trialAll;
    Cand* recs;
    TreeWs<290, 584>* slowWs;
    int tid;
    bool bigWeights;     // the histogram total may not fit the fast tree's 22-bit weights

    __device__ uint32_t* maskp(int id) const { return masks + (size_t)id * maxwords; }
    __device__ uint32_t* histp(int id) const { return hists + (size_t)id * 320; }

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
        for (int k = 0; k < ENG_NW; k++) r += S->hred[k];
        __syncthreads();
        return r | 1ull;
    }
    static __device__ __forceinline__ unsigned long long hash_words_warp(const uint32_t* p, int nwords32, int lane) {
        unsigned long long h = 0;
        for (int k = lane; k < nwords32; k += 32) {
            unsigned long long x = (unsigned long long)p[k] + 0x9E3779B97F4A7C15ull * (unsigned long long)(k + 1);
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
            h += x;
        }
        for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
        return h | 1ull;
    }

    // ---- pools ----------------------------------------------------------------------------------------------------
    // the Tab in staging slot `st` -> its id (hash-consed; a hash hit is confirmed by a full comparison).  One warp.
    __device__ __noinline__ int intern_tab_warp(int st, int lane) {
        const uint32_t* q = (const uint32_t*)&tabs[MAXT + st];
        const unsigned long long h = hash_words_warp(q, (int)(sizeof(Tab) / 4), lane);
        const int nT = S->sym.nTabs;
        int hit = -1;
        for (int k0 = 0; k0 < nT && hit < 0; k0 += 32) {
            const int k = k0 + lane;
            unsigned m = __ballot_sync(0xffffffffu, k < nT && tabHash[k] == h);
            while (m && hit < 0) {
                const int cand = k0 + __ffs((int)m) - 1;
                m &= m - 1;
                const uint32_t* a = (const uint32_t*)&tabs[cand];
                bool diff = false;
                for (int w = lane; w < (int)(sizeof(Tab) / 4); w += 32) diff |= a[w] != q[w];
                if (!__any_sync(0xffffffffu, diff)) hit = cand;
            }
        }
        if (hit < 0) {
            if (nT >= MAXT) { if (lane == 0) S->sym.overflow = 1; return 0; }
            hit = nT;
            uint32_t* a = (uint32_t*)&tabs[hit];
            for (int w = lane; w < (int)(sizeof(Tab) / 4); w += 32) a[w] = q[w];
            if (lane == 0) { tabHash[hit] = h; S->sym.trialState[hit] = ST_EMPTY; S->sym.nTabs = nT + 1; }
            __syncwarp();
        }
        return hit;
    }

    // the mask just written into pool slot nMasks -> its id (an equal older mask wins, so equal symbol lists reached
    // along different paths share their memo entries)
    __device__ __noinline__ int intern_mask() {
        P0();
        const int fresh = S->sym.nMasks;
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
            if (tid == 0) { S->maskHash[fresh] = h; S->sym.rc[fresh].state = ST_EMPTY; S->sym.nMasks = fresh + 1; }
        }
        __syncthreads();
        P1(PR_INTERN_MASK);
        return hit;
    }

    // ---- block load -----------------------------------------------------------------------------------------------
    // per-symbol views the passes read: kind (0 = not a match, else length symbol - 256) and, for matches,
    // minfo = len-3 | dist symbol << 9 | extra bits << 14 | (start offset in its cost tile) << 19; tileFirst[t] = first
    // symbol that starts in tile t.  Tiles count from the 8-byte boundary at or below the block's first decoded byte.
    __device__ __noinline__ void build_views() {
        const uint32_t a0 = (uint32_t)(v.out_off & ~7ull);
        const uint32_t ntiles = (uint32_t)(((v.out_off - a0) + v.ulen) / DC_TILE) + 2;
        for (uint32_t t = tid; t < ntiles; t += ENG_NT) tileFirst[t] = v.n;
        __syncthreads();
        for (uint32_t i = tid; i < v.n; i += ENG_NT) {
            const uint32_t s = v.sym[i];
            const uint32_t rel = v.symout[i] - a0;
            const bool mt = sym_is_match(s);
            kind[i] = mt ? (uint8_t)(sym_lensym(s) - 256) : (uint8_t)0;
            if (mt) {
                const int ds = dist_sym(sym_dist(s));
                minfo[i] = (s & 0x1FF) | ((uint32_t)ds << 9) | ((uint32_t)(len_ebits_of(sym_lensym(s)) + dist_ebits_of(ds)) << 14) |
                           ((rel & (DC_TILE - 1)) << 19);
            }
            const uint32_t ti = rel / DC_TILE;
            const int tprev = i ? (int)((v.symout[i - 1] - a0) / DC_TILE) : -1;
            if ((int)ti != tprev) tileFirst[ti] = i;
        }
        for (uint32_t i = v.n + tid; i < v.nwords * 32; i += ENG_NT) kind[i] = 0;
        __syncthreads();
    }

    // histogram of the symbol list with mask `mid` into S->hist, from the symbols
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

    // payload of the symbol list described by histogram h (global or shared) under table t (recodeToHuffmanInternal,
    // DeflateBlockHuffman.java:759-770); CTA-wide
    __device__ __noinline__ long long hist_payload(const uint32_t* h, const Tab& t) {
        long long acc = 0;
        for (int k = tid; k < 318; k += ENG_NT) {
            const uint32_t f = h[k];
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
        const long long r = (long long)S->red;
        __syncthreads();
        return r;
    }

    // pools forgotten; tab 0 = the fixed code
    __device__ __noinline__ void reset_pools() {
        __syncthreads();
        sym_reset(S->sym, tid, ENG_NT);
        for (int k = tid; k < MAXT; k += ENG_NT) S->tabDc[k] = 0xFF;
        if (tid < DCN) S->dcOwner[tid] = 0xFFFF;
        if (tid == 0) S->dcNext = 0;
        Tab& f = tabs[TAB_FIXED];
        for (int k = tid; k < MAX_LL; k += ENG_NT) f.L[k] = (k < 286) ? ((k <= 143) ? 8 : (k <= 255) ? 9 : (k <= 279) ? 7 : 8) : 0;
        if (tid < MAX_D) f.D[tid] = tid < 30 ? 5 : 0;
        if (tid == 0) { f.nL = 286; f.nD = 30; f.type = 1; f.pad[0] = f.pad[1] = f.pad[2] = 0; }
        __syncthreads();
        const unsigned long long h = hash_words((const uint32_t*)&f, (int)(sizeof(Tab) / 4));
        if (tid == 0) { tabHash[0] = h; S->sym.nTabs = 1; }
        __syncthreads();
    }

    // B := (mask in slot SLOT_B with its histogram in hists[SLOT_B], tables / header / payload of recs[0]); toFixed: the
    // block is first recoded to the fixed code (DeflateBlockHuffman.merge, :1233-1271), payload from the histogram
    __device__ __noinline__ void adopt_B(bool toFixed) {
        reset_pools();
        const uint32_t* ms = maskp(SLOT_B);
        uint32_t* m0 = maskp(0);
        for (uint32_t k = tid; k < v.nwords; k += ENG_NT) m0[k] = ms[k];
        const uint32_t* hs = histp(SLOT_B);
        uint32_t* h0 = histp(0);
        for (int k = tid; k < 320; k += ENG_NT) h0[k] = hs[k];
        __syncthreads();
        const unsigned long long h = hash_words(m0, (int)v.nwords);
        if (tid == 0) { S->maskHash[0] = h; S->sym.nMasks = 1; }
        const Cand& src = recs[0];
        SC b;
        b.mid = 0; b.ok = 1; b.hid = -1; b.tabid = TAB_FIXED; b.type = 1;
        long long pay = src.payload;
        if (toFixed) pay = hist_payload(h0, tabs[TAB_FIXED]);
        else if (src.tab.type == 2) {
            uint32_t* st = (uint32_t*)&tabs[MAXT];
            const uint32_t* q = (const uint32_t*)&src.tab;
            for (int k = tid; k < (int)(sizeof(Tab) / 4); k += ENG_NT) st[k] = q[k];
            uint32_t* hd = (uint32_t*)&hdrs[0];
            const uint32_t* hq = (const uint32_t*)&src.hdr;
            for (int k = tid; k < (int)(sizeof(Hdr) / 4); k += ENG_NT) hd[k] = hq[k];
            __syncthreads();
            if (tid < 32) {
                const int t = intern_tab_warp(0, tid);
                if (tid == 0) S->tmpIdx = t;
            }
            __syncthreads();
            b.tabid = (short)S->tmpIdx; b.type = 2; b.hid = 0;
            if (tid == 0) { S->sym.hbits[0] = src.hdr.bits; S->sym.nHdrs = 1; }
        }
        b.payload = pay;
        if (tid == 0) { S->en.B = b; S->en.blockType = b.type; }
        __syncthreads();
    }

    // a new block: symbol view set by the caller, candidate `src` (global), mask words `maskSrc`
    __device__ __noinline__ void load_block(const Cand& src, const uint32_t* maskSrc, bool toFixed) {
        P0();
        __syncthreads();
        bigWeights = v.ulen + (uint64_t)v.n + 4 >= (1ull << 22);
        uint32_t* mb = maskp(SLOT_B);
        for (uint32_t k = tid; k < v.nwords; k += ENG_NT) mb[k] = maskSrc[k];
        uint32_t* d = (uint32_t*)&recs[0];
        const uint32_t* s = (const uint32_t*)&src;
        for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
        build_views();
        pass_hist_full(SLOT_B);
        uint32_t* hb = histp(SLOT_B);
        for (int k = tid; k < 320; k += ENG_NT) hb[k] = S->hist[k];
        __syncthreads();
        adopt_B(toFixed);
        if (tid == 0) {
            S->en.S = &S->sym;
            S->en.trialAll = trialAll;
            S->tsink.buf = g_trace; S->tsink.cap = g_trace_cap; S->tsink.n = &g_trace_n;
            S->en.trace = g_trace ? &S->tsink : nullptr;
            S->en.storedOK = v.ulen <= 65535;
        }
        __syncthreads();
        P1(PR_LOAD);
    }

    // ---- cost arrays ------------------------------------------------------------------------------------------------
    static constexpr short DC_BLOCKED = 0x7FFF;

    // dc[i] = (literal cost of match i's bytes) - (cost of the match) under table `t`, DC_BLOCKED when a byte has no code
    // (DeflateBlockHuffman.java:238-246).  It does not depend on the mask, so every replace / least pass under the same
    // tables reads it.  One coalesced pass over the decoded bytes, see the file header.
    __device__ __noinline__ const short* ensure_dc(int t) {
        int slot = S->tabDc[t];
        if (slot != 0xFF) return dc + (size_t)slot * maxn;
        P0();
        __syncthreads();  // every thread has seen the miss before thread 0 records the new slot
        if (tid == 0) {
            slot = S->dcNext;
            S->dcNext = (slot + 1) % DCN;
            const int owner = S->dcOwner[slot];
            if (owner != 0xFFFF) S->tabDc[owner] = 0xFF;
            S->dcOwner[slot] = (unsigned short)t;
            S->tabDc[t] = (unsigned char)slot;
            S->tmpIdx = slot;
            S->carryIdx[0] = S->carryIdx[1] = -1;
        }
        const Tab& tb = tabs[t];
        for (int k = tid; k < 256 + 32 + 32; k += ENG_NT) S->ctab[k] = k < 256 ? tb.L[k] : k < 288 ? (k - 256 + 256 < MAX_LL ? tb.L[k] : 0) : tb.D[k - 288];
        __syncthreads();
        slot = S->tmpIdx;
        short* d = dc + (size_t)slot * maxn;
        const uint32_t a0 = (uint32_t)(v.out_off & ~7ull);
        const uint32_t head = (uint32_t)(v.out_off - a0);
        const uint32_t endRel = head + (uint32_t)v.ulen;
        const uint8_t* base = v.out + a0;
        const int lane = tid & 31, wid = tid >> 5;
        uint32_t* P = S->u.P;
        const uint32_t ntiles = endRel / DC_TILE + 1;
        for (uint32_t T = 0; T < ntiles; T++) {
            const uint32_t idx = T * DC_TILE + 8u * (uint32_t)tid;
            uint32_t x[8];
            uint32_t tot = 0;
            if (idx < endRel) {
                const uint2 q = *(const uint2*)(base + idx);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t bt = ((k < 4 ? q.x : q.y) >> (8 * (k & 3))) & 0xffu;
                    const uint32_t j = idx + k;
                    const uint32_t c = S->ctab[bt];
                    x[k] = (j >= head && j < endRel) ? (c ? c : 0x10000u) : 0u;
                    tot += x[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = 0;
            }
            uint32_t incl = tot;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, dd);
                if (lane >= dd) incl += y;
            }
            if (lane == 31) S->wt[wid] = incl;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int k = 0; k < ENG_NW; k++) { const uint32_t y = S->wt[k]; if (k < wid) wbase += y; total += y; }
            uint32_t run = wbase + incl - tot;
#pragma unroll
            for (int k = 0; k < 8; k++) { P[8 * tid + k] = run; run += x[k]; }
            if (tid == ENG_NT - 1) P[DC_TILE] = total;
            __syncthreads();
            // the match that started in the previous tile and ends in this one
            const int cp = (int)((T + 1) & 1), cq = (int)(T & 1);   // previous tile's carry slot, this tile's
            if (tid == 0 && S->carryIdx[cp] >= 0) {
                const int ci = S->carryIdx[cp];
                const uint32_t mi = minfo[ci];
                const uint32_t e = ((mi >> 19) & (DC_TILE - 1)) + (mi & 0x1FF) + 3 - DC_TILE;
                const uint32_t y = S->carryPart[cp] + P[e];
                d[ci] = (y >> 16) ? DC_BLOCKED : (short)((int)(y & 0xffffu) - S->carryRef[cp]);
                S->carryIdx[cp] = -1;
            }
            const uint32_t i0 = tileFirst[T], i1 = tileFirst[T + 1];
            for (uint32_t i = i0 + tid; i < i1; i += ENG_NT) {
                const int kq = kind[i];
                if (!kq) continue;
                const uint32_t mi = minfo[i];
                const uint32_t s = (mi >> 19) & (DC_TILE - 1), e = s + (mi & 0x1FF) + 3;
                const int ref = S->ctab[256 + kq] + S->ctab[288 + ((mi >> 9) & 31)] + (int)((mi >> 14) & 31);
                if (e <= DC_TILE) {
                    const uint32_t y = P[e] - P[s];
                    d[i] = (y >> 16) ? DC_BLOCKED : (short)((int)(y & 0xffffu) - ref);
                } else {   // at most one match per tile crosses its end
                    S->carryIdx[cq] = (int)i; S->carryRef[cq] = ref; S->carryPart[cq] = P[DC_TILE] - P[s];
                }
            }
            __syncthreads();
        }
        P1(PR_DC);
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
        const uint32_t* hs = histp(src);
        uint32_t* hd = histp(dst);
        for (int k = tid; k < 320; k += ENG_NT) hd[k] = hs[k] + S->hist[k];
    }

    // replaceBackrefsWithLiteralsIfSmaller(prune) for memo slot `slot` = (mid, tabid)
    __device__ __noinline__ void pass_replace(int slot, int mid, int tabid, bool prune) {
        if (S->sym.nMasks >= MAXM) { if (tid == 0) S->sym.overflow = 1; __syncthreads(); return; }
        const short* d = ensure_dc(tabid);
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = 0;
        if (tid == 0) { S->red = 0; S->redAny = 0; }
        __syncthreads();
        const uint32_t* m = maskp(mid);
        const int fresh = S->sym.nMasks;
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
                    if (k && !((ob >> u) & 1) && (prune ? x <= 0 : x < 0)) { saved -= x; nbits |= 1u << u; }
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
            PSlot& p = S->sym.pm[slot];
            p.mid = (unsigned short)newmid; p.delta = (long long)S->red; p.state = ST_DONE;
        }
        __syncthreads();
    }

    // removeDistLitLeastExpensive(mode) for memo slot `slot`
    __device__ __noinline__ void pass_least(int slot, int mid, int tabid, int mode) {
        if (S->sym.nMasks >= MAXM) { if (tid == 0) S->sym.overflow = 1; __syncthreads(); return; }
        const short* d = ensure_dc(tabid);
        if (tid < 32) { S->leastSum[tid] = 0; S->leastCnt[tid] = 0; }
        if (tid == 0) { S->leastBlocked = 0; S->leastSeen = 0; }
        for (int k = tid; k < 320; k += ENG_NT) S->hist[k] = 0;
        __syncthreads();
        const uint32_t* m = maskp(mid);
        const int lane = tid & 31;
        // per length symbol: sum of (literal - match) cost, count, blocked (:386-420); lanes of a warp that hold the same
        // length symbol are summed with one shared-memory atomic
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
                const int k = (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff);
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
            const int fresh = S->sym.nMasks;
            uint32_t* md = maskp(fresh);
            uint8_t* mdb = (uint8_t*)md;
            for (uint32_t i0 = (uint32_t)tid * 8; i0 < v.nwords * 32; i0 += ENG_NT * 8) {
                const uint2 kk = *(const uint2*)(kind + i0);
                const uint32_t ob = mbytes[i0 >> 3];
                uint32_t nbits = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int k = (int)(((u < 4 ? kk.x : kk.y) >> (8 * (u & 3))) & 0xff);
                    if (k == rem + 1) nbits |= 1u << u;
                }
                mdb[i0 >> 3] = (uint8_t)(ob | nbits);
                for (uint32_t b = nbits & ~ob; b; b &= b - 1) hist_delta_replace(i0 + (uint32_t)__ffs((int)b) - 1);
            }
            __syncthreads();
            newmid = intern_mask();
            if (newmid == fresh) hist_store_delta(newmid, mid);
        }
        if (tid == 0) {
            PSlot& p = S->sym.pm[slot];
            p.mid = (unsigned short)newmid; p.delta = -(long long)S->red; p.state = ST_DONE;
        }
        __syncthreads();
    }

    __device__ __noinline__ void exec_passes() {
        const int nq = S->sym.nqPass;
        if (!nq) return;
        P0();
        for (int q = 0; q < nq; q++) {
            const int slot = S->sym.qPass[q];
            const unsigned key = S->sym.pm[slot].key;
            const int mid = pm_key_mid(key), tabid = pm_key_tab(key), op = pm_key_op(key);
            if (op == OP_FIXED) {   // recodeToFixedHuffman: the fixed-code payload is a function of the symbol list alone
                const long long pay = hist_payload(histp(mid), tabs[TAB_FIXED]);
                if (tid == 0) { PSlot& p = S->sym.pm[slot]; p.delta = pay; p.mid = (unsigned short)mid; p.state = ST_DONE; }
                __syncthreads();
            } else if (op <= OP_REPLACE_PRUNE) pass_replace(slot, mid, tabid, op == OP_REPLACE_PRUNE);
            else pass_least(slot, mid, tabid, op - OP_LEAST0);
        }
        // a pass that could not run (mask pool full) leaves its slot pending: the round restarts after a reset
        PCOUNT(PR_PASS, nq - 1);
        P1(PR_PASS);
    }

    // ---- recodeHuffman, one warp per request ------------------------------------------------------------------------
    // runs of equal code lengths of a Tab, cut by one warp: ballot + popcount compaction
    struct RunListW : RunList { uint16_t start[MAX_PAIRS]; };
    static __device__ __forceinline__ void runlist_warp(const Tab& t, RunListW& rl, int lane) {
        const int nL = t.nL, n = t.nL + t.nD;
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int vv = i < n ? (i < nL ? t.L[i] : t.D[i - nL]) : -1;
            const int pv = (i > 0 && i < n) ? (i - 1 < nL ? t.L[i - 1] : t.D[i - 1 - nL]) : -2;
            const bool st = i < n && vv != pv;
            const unsigned bal = __ballot_sync(0xffffffffu, st);
            if (st) {
                const int k = cnt + __popc(bal & ((1u << lane) - 1u));
                rl.val[k] = (uint8_t)vv;
                rl.start[k] = (uint16_t)i;
            }
            cnt += __popc(bal);
        }
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) rl.len[k] = (uint16_t)((k + 1 < cnt ? rl.start[k + 1] : n) - rl.start[k]);
        if (lane == 0) rl.n = (uint16_t)cnt;
        __syncwarp();
    }

    // header code of the pair frequencies f[0..19) (Huffman.ofRLEPacked), trimmed on from `ncl`, and the header size:
    // sum of pair bits = sum_s freq[s] * CL[s] + 2 f16 + 3 f17 + 7 f18.  One thread.
    __device__ __forceinline__ void hdr_code_from_freq(const uint32_t* f, uint8_t* CL, int ncl, int* nclOut, int* bitsOut, TreeWsCLc& ws) {
        if (huff_tree_ws(f, 19, 7, CL, ws)) S->err = ERR_TREE;
        ncl = trim_ncl(CL, ncl);
        int bits = 5 + 5 + 4 + 3 * ncl + 2 * (int)f[16] + 3 * (int)f[17] + 7 * (int)f[18];
        for (int k = 0; k < 19; k++) bits += (int)f[k] * CL[k];
        *nclOut = ncl;
        *bitsOut = bits;
    }

    // rewriteHeader with the default strategy (DeflateBlockHuffman.java:480-577) for table `t` into header `h` (both
    // global), by one warp: runs -> pairs per run -> exclusive scan -> every run writes its pairs
    __device__ __noinline__ void hdr_default_warp(const Tab& t, Hdr& h, unsigned char* wsb, int lane) {
        RunListW& rl = *reinterpret_cast<RunListW*>(wsb);
        uint16_t* off = reinterpret_cast<uint16_t*>(wsb + 1608);          // MAX_PAIRS + 2
        uint32_t* f19 = reinterpret_cast<uint32_t*>(wsb + 2256);          // 19 (+ CL staging)
        uint8_t* cl = wsb + 2336;                                          // 19
        TreeWsCLc& tw = *reinterpret_cast<TreeWsCLc*>(wsb + 2368);
        static_assert(sizeof(RunListW) <= 1608 && 2368 + sizeof(TreeWsCLc) <= WS_BYTES, "warp workspace layout");
        if (lane < 19) f19[lane] = 0;
        runlist_warp(t, rl, lane);
        const int R = rl.n;
        for (int r = lane; r < R; r += 32) {
            int cnt = 0;
            emit_run(rl.val[r], rl.len[r], FLAGS_DEFAULT, [&](int, int, int, int k) { cnt += k; });
            off[r] = (uint16_t)cnt;
        }
        __syncwarp();
        int carry = 0;
        for (int i0 = 0; i0 < R; i0 += 32) {
            const int i = i0 + lane;
            const int x = i < R ? off[i] : 0;
            int incl = x;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += y; }
            if (i < R) off[i] = (uint16_t)(carry + incl - x);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            int o = off[r];
            emit_run(rl.val[r], rl.len[r], FLAGS_DEFAULT, [&](int sym, int run, int val, int k) {
                const uint16_t p = pair_pack(sym, run, val);
                for (int q = 0; q < k; q++) h.pairs[o++] = p;
                atomicAdd(&f19[sym], (uint32_t)k);
            });
        }
        __syncwarp();
        if (lane == 0) {
            int ncl, bits;
            hdr_code_from_freq(f19, cl, 19, &ncl, &bits, tw);
            h.np = (uint16_t)carry; h.ncl = (uint8_t)ncl; h.bits = bits;
            for (int k = 0; k < 19; k++) h.CL[k] = cl[k];
        }
        __syncwarp();
    }

    __device__ __noinline__ void recode_one(int mid, int w, int lane, int hid) {
        unsigned char* wsb = S->u.ws[w];
        uint32_t* heap = reinterpret_cast<uint32_t*>(wsb);                 // 294 words (also the distance tree workspace)
        uint32_t* freq = reinterpret_cast<uint32_t*>(wsb + 1184);          // 320 words; the litlen tree's parent[] later
        uint16_t* value = reinterpret_cast<uint16_t*>(wsb + 2464);         // 292
        static_assert(sizeof(TreeWs<32, 68>) <= 1184 && 2464 + 292 * 2 <= WS_BYTES, "warp workspace layout");
        const uint32_t* hs = histp(mid);
        for (int k = lane; k < 320; k += 32) freq[k] = hs[k];
        Tab& T = tabs[MAXT + w];
        for (int k = lane; k < MAX_LL; k += 32) T.L[k] = 0;
        T.D[lane] = 0;
        __syncwarp();
        if (lane == 0) {
            // trailing zero-frequency trimming + the distance special cases (DeflateBlockHuffman.java:683-740)
            const uint32_t* df = freq + 288;
            int nd = 30;
            while (nd > 0 && df[nd - 1] == 0) nd--;
            int nz = 0;
            for (int k = 0; k < nd; k++) nz += df[k] != 0;
            if (nd == 0) { T.nD = 1; }                                          // handleZero: one entry, length 0
            else if (nz <= 1) { T.nD = (uint16_t)nd; T.D[nd - 1] = 1; }         // handleOne
            else {
                T.nD = (uint16_t)nd;
                if (huff_tree<32, 68>(df, nd, 15, T.D, *reinterpret_cast<TreeWs<32, 68>*>(wsb))) S->err = ERR_TREE;
            }
            int nl = 286;
            while (nl > 0 && freq[nl - 1] == 0) nl--;
            T.nL = (uint16_t)nl;
            int rc = bigWeights ? 2 : huff_tree_fast(freq, nl, 15, T.L, heap, value);
            if (rc == 2) {   // deeper than 15 (or weights too large for the fast keys): the full algorithm with its limiter
                PCOUNT(PR_SLOWTREE, 1);
                for (int k = 0; k < nl; k++) T.L[k] = 0;
                if (huff_tree<290, 584>(hs, nl, 15, T.L, slowWs[w])) S->err = ERR_TREE;
            }
            T.type = 2; T.pad[0] = T.pad[1] = T.pad[2] = 0;
        }
        __syncwarp();
        // payload = histogram . (code length + extra bits) (recodeToHuffmanInternal, :759-770)
        long long acc = 0;
        for (int k = lane; k < 318; k += 32) {
            const uint32_t f = hs[k];
            if (!f) continue;
            int bits;
            if (k < 257) bits = T.L[k];
            else if (k < 286) bits = T.L[k] + len_ebits_of(k);
            else if (k >= 288) bits = T.D[k - 288] + dist_ebits_of(k - 288);
            else bits = 0;
            acc += (long long)f * bits;
        }
        for (int dd = 16; dd > 0; dd >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, dd);
        if (lane == 0) S->recPay[w] = acc;
        hdr_default_warp(T, hdrs[hid], wsb, lane);
        if (lane == 0) S->sym.hbits[hid] = hdrs[hid].bits;
        __syncwarp();
    }

    __device__ __noinline__ void exec_recodes() {
        const int nq = S->sym.nqRec;
        if (!nq) return;
        P0();
        const int w = tid >> 5, lane = tid & 31;
        const int h0 = S->sym.nHdrs;
        __syncthreads();
        for (int base = 0; base < nq; base += ENG_NW) {
            const int r = base + w;
            const bool mine = r < nq && h0 + r < MAXH;
            if (mine) recode_one(S->sym.qRec[r], w, lane, h0 + r);
            __syncthreads();
            if (w == 0) {   // intern the staged tables one after the other (two requests may produce the same table)
                for (int k = 0; k < ENG_NW && base + k < nq; k++) {
                    if (h0 + base + k >= MAXH) { if (lane == 0) S->sym.overflow = 1; continue; }
                    const int t = intern_tab_warp(k, lane);
                    if (lane == 0 && !S->sym.overflow) {
                        RSlot& rs = S->sym.rc[S->sym.qRec[base + k]];
                        rs.tabid = (unsigned short)t; rs.hid = (unsigned short)(h0 + base + k); rs.payload = S->recPay[k]; rs.state = ST_DONE;
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (tid == 0) S->sym.nHdrs = min(MAXH, h0 + nq);
        __syncthreads();
        PCOUNT(PR_RECODE, nq - 1);
        P1(PR_RECODE);
    }

    // ---- header mutators, one warp per request ------------------------------------------------------------------------
    // replaceRLERunsWithLiteralsIfSmaller (DeflateBlockHuffman.java:321-332): in[0..np) -> out, returns the new count;
    // *saved = bits saved under CL
    static __device__ __forceinline__ int replace_runs_warp(const uint16_t* in, int np, const uint8_t* CL, bool prune, uint16_t* out,
                                                           int* saved, int lane) {
        int base = 0, sv = 0;
        for (int c0 = 0; c0 < np; c0 += 32) {
            const int i = c0 + lane;
            const uint16_t p = i < np ? in[i] : 0;
            const int run = i < np ? pair_run(p) : 0;
            bool rep = false;
            if (run > 0) {
                const int size = pair_size(p, CL);
                const int b = CL[pair_val(p)];
                const int tot = b * run;
                rep = b >= 1 && (prune ? tot <= size : tot < size);
                if (rep) sv += size - tot;
            }
            const int cnt = i < np ? (rep ? run : 1) : 0;
            int incl = cnt;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += y; }
            int o = base + incl - cnt;
            if (rep) { const int vv = pair_val(p); const uint16_t q = pair_pack(vv, 0, vv); for (int k = 0; k < cnt; k++) out[o++] = q; }
            else if (cnt) out[o] = p;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (int dd = 16; dd > 0; dd >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, dd);
        *saved = sv;
        __syncwarp();
        return base;
    }

    __device__ __noinline__ void hdrop_one(int src, int op, int dst, int w, int lane) {
        unsigned char* wsb = S->u.ws[w];
        uint16_t* in = reinterpret_cast<uint16_t*>(wsb);                  // 320
        uint16_t* out = reinterpret_cast<uint16_t*>(wsb + 640);           // 320
        uint32_t* f19 = reinterpret_cast<uint32_t*>(wsb + 1280);          // 19
        uint8_t* cl = wsb + 1360;                                          // 19 old
        uint8_t* cl2 = wsb + 1392;                                         // 19 new
        TreeWsCLc& tw = *reinterpret_cast<TreeWsCLc*>(wsb + 1424);
        const Hdr& hs = hdrs[src];
        Hdr& hd = hdrs[dst];
        const int np = hs.np;
        for (int k = lane; k < np; k += 32) in[k] = hs.pairs[k];
        if (lane < 19) { cl[lane] = hs.CL[lane]; f19[lane] = 0; }
        int ncl = hs.ncl, bits = hs.bits;
        __syncwarp();
        const uint16_t* pairs = in;
        int npo = np;
        if (op == HOP_OPT) {   // optimiseHeader (:471-476): trailing zero code-length trim, then runs -> literals
            const int n2 = trim_ncl(cl, ncl);
            bits -= 3 * (ncl - n2);
            ncl = n2;
            int saved;
            npo = replace_runs_warp(in, np, cl, false, out, &saved, lane);
            bits -= saved;
            pairs = out;
            if (lane < 19) cl2[lane] = cl[lane];
        } else {               // recodeHeader (:579-629), after replaceRLERuns(prune) for the LessRLE variant (:632-635)
            if (op == HOP_RECODE_LESS) {
                int saved;
                npo = replace_runs_warp(in, np, cl, true, out, &saved, lane);
                pairs = out;
            }
            for (int k = lane; k < npo; k += 32) atomicAdd(&f19[pair_sym(pairs[k])], 1u);
            __syncwarp();
            if (lane == 0) hdr_code_from_freq(f19, cl2, ncl, &ncl, &bits, tw);   // numCodelenLens is NOT reset (H8)
            ncl = __shfl_sync(0xffffffffu, ncl, 0);
            bits = __shfl_sync(0xffffffffu, bits, 0);
        }
        __syncwarp();
        for (int k = lane; k < npo; k += 32) hd.pairs[k] = pairs[k];
        if (lane < 19) hd.CL[lane] = cl2[lane];
        if (lane == 0) { hd.np = (uint16_t)npo; hd.ncl = (uint8_t)ncl; hd.bits = bits; S->sym.hbits[dst] = bits; }
        __syncwarp();
    }

    __device__ __noinline__ void exec_hdrops() {
        const int nq = S->sym.nqHdr;
        if (!nq) return;
        P0();
        const int w = tid >> 5, lane = tid & 31;
        const int h0 = S->sym.nHdrs;
        __syncthreads();
        for (int r = w; r < nq; r += ENG_NW) {
            const int src = S->sym.qHdr[r] >> 2, op = S->sym.qHdr[r] & 3;
            if (h0 + r >= MAXH) { if (lane == 0) S->sym.overflow = 1; continue; }
            hdrop_one(src, op, h0 + r, w, lane);
            if (lane == 0) S->sym.hop[src][op] = (unsigned short)(h0 + r + 1);
        }
        __syncthreads();
        if (tid == 0) S->sym.nHdrs = min(MAXH, h0 + nq);
        __syncthreads();
        PCOUNT(PR_HDROP, nq - 1);
        P1(PR_HDROP);
    }

    // ---- the 56 header strategy trials of the queued tables ----------------------------------------------------------
    __device__ __noinline__ void exec_trials() {
        const int nq = S->sym.nqTrial;
        if (!nq) return;
        P0();
        const int w = tid >> 5, lane = tid & 31;
        constexpr int GROUP = (ENG_NT / 28) < ENG_NW ? (ENG_NT / 28) : ENG_NW;
        for (int base = 0; base < nq; base += GROUP) {
            const int cnt = nq - base < GROUP ? nq - base : GROUP;
            if (w < cnt) runlist_warp(tabs[S->sym.qTrial[base + w]], *reinterpret_cast<RunListW*>(S->u.ws[w]), lane);
            __syncthreads();
            if (tid < cnt * 28) {
                const int b = tid / 28, c = tid % 28;
                const int t = S->sym.qTrial[base + b];
                // c_trial_flags order: [0,20) and [40,48) are the rewrite strategies without prune, +20 / +8 with it
                const int kF = c < 20 ? c : 40 + (c - 20), kT = c < 20 ? 20 + c : 48 + (c - 20);
                TreeWsCLc ws;
                int a = 0, bp = 0;
                if (trial_sizes(*reinterpret_cast<RunListW*>(S->u.ws[b]), c_trial_flags[kF], &a, &bp, ws)) S->err = ERR_TREE;
                trialAll[t * 56 + kF] = a;
                trialAll[t * 56 + kT] = bp;
            }
            __syncthreads();
            if (tid < cnt) {
                const int t = S->sym.qTrial[base + tid];
                int best = 0x7fffffff, arg = 0;
                for (int k = 0; k < 56; k++) { const int bts = trialAll[t * 56 + k]; if (bts < best) { best = bts; arg = k; } }
                S->sym.trialBits[t] = best; S->sym.trialArg[t] = (unsigned char)arg; S->sym.trialState[t] = ST_DONE;
            }
            __syncthreads();
        }
        PCOUNT(PR_TRIALS, nq - 1);
        P1(PR_TRIALS);
    }

    // everything the last sweep asked for.  Header trials feed nothing but the selection, so they wait until a batch is
    // worth the threads (or nothing else is left to do).
    __device__ __noinline__ void execute() {
        __syncthreads();
        exec_passes();
        exec_recodes();
        exec_hdrops();
        const bool others = S->sym.nqPass + S->sym.nqRec + S->sym.nqHdr > 0;
        const bool doTrials = S->sym.nqTrial && (!others || S->sym.nqTrial >= 2 * ENG_NW);
        __syncthreads();
        if (doTrials) exec_trials();
        if (tid == 0) {
            S->sym.nqPass = S->sym.nqRec = S->sym.nqHdr = 0;
            if (doTrials) S->sym.nqTrial = 0;
        }
        __syncthreads();
    }

    // ---- materialisation --------------------------------------------------------------------------------------------
    // candidate c (with header strategy `arg` when it is a trial winner) -> recs[which] + mask / histogram slot
    __device__ __noinline__ void materialise(const SC c, int arg, int which) {
        P0();
        __syncthreads();
        Cand& dst = recs[which];
        const int slot = which == 0 ? SLOT_B : SLOT_BEST;
        if (c.mid != slot) {
            const uint32_t* ms = maskp(c.mid);
            uint32_t* md = maskp(slot);
            for (uint32_t k = tid; k < v.nwords; k += ENG_NT) md[k] = ms[k];
            const uint32_t* hs = histp(c.mid);
            uint32_t* hd = histp(slot);
            for (int k = tid; k < 320; k += ENG_NT) hd[k] = hs[k];
        }
        const uint32_t* ts = (const uint32_t*)&tabs[c.tabid];
        uint32_t* td = (uint32_t*)&dst.tab;
        for (int k = tid; k < (int)(sizeof(Tab) / 4); k += ENG_NT) td[k] = ts[k];
        __syncthreads();
        if (c.type == 2) {
            if (arg >= 0) {   // a header strategy trial won: optimiseBlockDynBlock (DeflateStream.java:184-198) for real
                if (tid == 0) {
                    if (hdr_trial(dst.tab, c_trial_flags[arg], S->u.mat.hdr, S->u.mat.ws)) S->err = ERR_TREE;
                }
                __syncthreads();
                const uint32_t* hs = (const uint32_t*)&S->u.mat.hdr;
                uint32_t* hd = (uint32_t*)&dst.hdr;
                for (int k = tid; k < (int)(sizeof(Hdr) / 4); k += ENG_NT) hd[k] = hs[k];
            } else {
                const uint32_t* hs = (const uint32_t*)&hdrs[c.hid];
                uint32_t* hd = (uint32_t*)&dst.hdr;
                for (int k = tid; k < (int)(sizeof(Hdr) / 4); k += ENG_NT) hd[k] = hs[k];
            }
        } else if (tid == 0) { dst.hdr.np = 0; dst.hdr.ncl = 0; dst.hdr.bits = 0; }
        if (tid == 0) { dst.payload = c.payload; dst.mid = 0; dst.tabid = 0; dst.pad2 = 0; }
        __syncthreads();
        P1(PR_MATERIAL);
    }

    // ---- one optimiseBlock call ---------------------------------------------------------------------------------------
    // sweeps + execution until the part `seg` of the enumeration is resolved, then the selection sweep.  Returns false on
    // a pool overflow (nothing selected).
    __device__ __noinline__ bool run_segment(unsigned seg) {
        if (tid == 0) { S->sym.doneMulti = 0; S->sym.doneRun = 0; S->sym.doneAor = 0; }
        __syncthreads();
        for (int it = 0;; it++) {
            if (tid == 0) {
                P0();
                S->sweepDone = S->en.sweep(false, seg) ? 1 : 0;
                P1(PR_SWEEP);
            }
            __syncthreads();
            if (S->sym.overflow) return false;
            if (S->sweepDone) break;
            const bool nothing = S->sym.nqPass + S->sym.nqRec + S->sym.nqHdr + S->sym.nqTrial == 0;
            if (nothing || it > 4096 || S->en.internalError) { if (tid == 0) S->err = ERR_INTERNAL; __syncthreads(); return false; }
            execute();
            if (S->sym.overflow) return false;
        }
        if (tid == 0) {
            P0();
            const unsigned before = S->en.bestIndex;
            S->en.sweep(true, seg);
            S->segImproved = S->en.bestIndex != before && !S->en.bestStored;
            if (S->en.internalError || S->en.poisoned) S->err = ERR_INTERNAL;
            P1(PR_SELECT);
        }
        __syncthreads();
        return true;
    }

    // DeflateStream.optimiseBlock (:343-490) for B.  storedSize < 0: no stored candidate is compared (phase A resolves
    // it afterwards from sizeI / sizeC1 / restMin).  Result: S->en.bestSize / bestStored / sizeI / sizeC1 / restMin /
    // bestIndex, and the winning Huffman candidate (B itself when nothing is smaller) materialised in recs[1].
    __device__ __noinline__ void optimise_block(long long storedSize) {
        P0();
        __syncthreads();
        if (tid == 0) { S->en.storedSize = storedSize; S->en.begin_round(); S->segmentedRound = 0; }
        __syncthreads();
        bool ok = run_segment(Enumer::SEG_ALL);
        if (ok) {
            materialise(S->en.best, S->en.bestArg, 1);
        } else if (!S->err) {
            // a pool filled up: the same round again in four segments, with the pools reset (and B re-adopted) in between;
            // the running best lives in recs[1]
            PCOUNT(PR_SEGMENTED, 1);
            materialise(S->en.B, -1, 0);
            __syncthreads();
            if (tid == 0) { S->en.begin_round(); S->segmentedRound = 1; }   // the winner's ids belong to pools that are gone
            __syncthreads();
            const unsigned segs[4] = {Enumer::SEG_HEAD | Enumer::SEG_MULTI_H, Enumer::SEG_MULTI_O, Enumer::SEG_FIXED | Enumer::SEG_LEAST0,
                                      Enumer::SEG_LEAST1};
            bool haveBest = false;
            for (int k = 0; k < 4; k++) {
                const long long bs = S->en.bestSize;
                const unsigned bi = S->en.bestIndex, ci = S->en.candIndex;
                const int bst = S->en.bestStored;
                const long long rm = S->en.restMin, c1 = S->en.sizeC1;
                __syncthreads();
                adopt_B(false);
                if (tid == 0) {   // adopt_B rewrote B's ids; the selection state carries over
                    S->en.bestSize = bs; S->en.bestIndex = bi; S->en.candIndex = ci; S->en.bestStored = bst;
                    S->en.restMin = rm; S->en.sizeC1 = c1;
                }
                __syncthreads();
                if (!run_segment(segs[k])) { if (tid == 0 && !S->err) S->err = ERR_POOL; __syncthreads(); break; }
                if (S->segImproved) { materialise(S->en.best, S->en.bestArg, 1); haveBest = true; }
                __syncthreads();
            }
            if (!haveBest && !S->err) { adopt_B(false); materialise(S->en.B, -1, 1); }
        }
        __syncthreads();
        P1(PR_ROUND);
    }

    // the winner of the last round becomes B (DeflateStream.java:514-530): pools and memo tables are kept while there is
    // room, since the next round revisits most of this round's states
    __device__ __noinline__ void advance_to_best() {
        P0();
        __syncthreads();
        const bool keep = !S->segmentedRound && !S->sym.overflow && S->sym.nMasks <= MAXM / 2 && S->sym.nTabs <= MAXT / 2 && S->sym.nHdrs + 1 <= MAXH / 2 &&
                          S->sym.nP <= PMEMO * 3 / 8 && S->en.bestIndex != 0xffffffffu;
        // recs[0] := recs[1], slot B := slot BEST
        {
            const uint32_t* s = (const uint32_t*)&recs[1];
            uint32_t* d = (uint32_t*)&recs[0];
            for (int k = tid; k < (int)(sizeof(Cand) / 4); k += ENG_NT) d[k] = s[k];
            const uint32_t* ms = maskp(SLOT_BEST);
            uint32_t* md = maskp(SLOT_B);
            for (uint32_t k = tid; k < v.nwords; k += ENG_NT) md[k] = ms[k];
            const uint32_t* hs = histp(SLOT_BEST);
            uint32_t* hd = histp(SLOT_B);
            for (int k = tid; k < 320; k += ENG_NT) hd[k] = hs[k];
        }
        __syncthreads();
        if (keep) {
            SC b = S->en.best;
            if (b.type == 2 && S->en.bestArg >= 0) {   // the trial header only exists in the record: give it an id
                const int hid = S->sym.nHdrs;
                const uint32_t* hs = (const uint32_t*)&recs[0].hdr;
                uint32_t* hd = (uint32_t*)&hdrs[hid];
                for (int k = tid; k < (int)(sizeof(Hdr) / 4); k += ENG_NT) hd[k] = hs[k];
                __syncthreads();
                if (tid == 0) {
                    S->sym.hbits[hid] = recs[0].hdr.bits;
                    S->sym.nHdrs = hid + 1;
                    S->sym.hop[hid][0] = S->sym.hop[hid][1] = S->sym.hop[hid][2] = 0;
                }
                b.hid = (short)hid;
            }
            __syncthreads();
            if (tid == 0) { S->en.B = b; S->en.blockType = b.type; }
            __syncthreads();
        } else {
            adopt_B(false);
        }
        P1(PR_REBASE);
    }
};

__device__ inline void eng_init(Eng& e, EngSmem* S, const EngScratch& sc, int cta) {
    e.S = S;
    e.tid = (int)threadIdx.x;
    e.maxwords = sc.maxwords;
    e.maxn = sc.maxwords * 32;
    e.masks = sc.masks + (size_t)cta * (MAXM + 2) * sc.maxwords;
    e.tabs = sc.tabs + (size_t)cta * (MAXT + ENG_NW);
    e.hdrs = sc.hdrs + (size_t)cta * MAXH;
    e.hists = sc.hists + (size_t)cta * (MAXM + 2) * 320;
    e.tabHash = sc.tabHash + (size_t)cta * MAXT;
    e.dc = sc.dc + (size_t)cta * DCN * e.maxn;
    e.kind = sc.kind + (size_t)cta * e.maxn;
    e.minfo = sc.minfo + (size_t)cta * e.maxn;
    e.tileFirst = sc.tileFirst + (size_t)cta * sc.maxtiles;
    e.trialAll = sc.trialAll + (size_t)cta * MAXT * 56;
    e.recs = sc.recs + (size_t)cta * 2;
    e.slowWs = sc.slowWs + (size_t)cta * ENG_NW;
    e.bigWeights = false;
    if (threadIdx.x == 0) S->err = 0;
    __syncthreads();
}

}  // namespace d4
