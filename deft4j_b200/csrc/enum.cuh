// enum.cuh — the candidate enumerator of DeflateStream.optimiseBlock (DeflateStream.java:343-490) as a SYMBOLIC
// program over ids, separated from the data-parallel work it asks for.
//
// A candidate of the reference (a DeflateBlockHuffman copy) is determined by
//     mid    id of its symbol list  = bit mask "this match has been replaced by its literals" (common.cuh)
//     tabid  id of its code tables  (Tab, hash-consed)
//     hid    id of its dynamic header (Hdr: RLE pairs, header code, numCodelenLens, size)
//     payload = litlenSizeBits
// and every mutator of the reference is a pure function of those ids:
//     replace / least passes   (mid, tabid, op) -> (mid', payload delta)        DeflateBlockHuffman.java:222-319,373-458
//     recodeHuffman            mid -> (tabid, default hid, payload)               :670-743
//     recodeToFixedHuffman     mid -> payload                                     :637-653
//     header mutators          (hid, op) -> hid'                                  :471-476,579-635
//     56 header strategy trials  tabid -> (first-minimum bits, its index)         DeflateStream.java:277-316
// ONE thread walks the reference's enumeration (multi -> run -> aor -> trials, the tie-break contract of SURVEY.md §8a)
// over memo tables of those functions.  A miss whose inputs are known is queued as a request and poisons everything
// that depends on it; after the sweep the whole CTA computes the queued requests side by side (engine.cuh: Huffman
// trees of different requests in different warps, passes CTA-wide, header trials 28 threads per table) and the sweep
// is repeated, skipping the sub-trees (aor / run / multi nodes) that completed earlier.  The enumeration does not depend
// on which candidate is winning, so when a sweep raises no request a final sweep feeds every candidate to the selection
// callback (DeflateStream.java:349-368) in the reference's order, with the reference's strict `<`.
//
// The file compiles as host code too (D4_HOST_TEST): tests/test_host_units.py drives this very enumerator with a plain
// serial executor (hosttest.cu) against the oracle's candidate trace.
#pragma once
#include "huff.cuh"

namespace d4 {

#ifdef D4_HOST_TEST
#define D4_HD inline
#define D4_HD_BIG inline
#else
#define D4_HD __device__ __forceinline__
#define D4_HD_BIG __device__ __noinline__   // one copy of each enumerator routine (they nest four deep)
#endif

#ifdef D4_SMALL_POOLS           // stress build: ordinary inputs overflow the pools, which forces the segmented rounds
constexpr int MAXM = 96, MAXT = 96, MAXH = 192, PMEMO = 256;
#else
constexpr int MAXM = 256;     // distinct symbol-list masks kept per block
constexpr int MAXT = 256;     // distinct code tables kept per block
constexpr int MAXH = 512;     // dynamic headers kept per block
constexpr int PMEMO = 512;    // pass memo slots (open addressing)
#endif
constexpr int QPASS = 64, QREC = 24, QHDR = 64, QTRIAL = 64;
constexpr int TAB_FIXED = 0;  // tab id 0 is always the fixed code (HuffmanTable.java:166-209)

// ops of the pass memo
enum { OP_REPLACE = 0, OP_REPLACE_PRUNE = 1, OP_LEAST0 = 2, OP_LEAST1 = 3, OP_FIXED = 4 };
// header mutators
enum { HOP_RECODE = 0, HOP_RECODE_LESS = 1, HOP_OPT = 2 };
enum { ST_EMPTY = 0, ST_PENDING = 1, ST_DONE = 2 };

struct SC {                    // symbolic candidate
    short mid, tabid, hid;     // hid < 0: none (FIXED)
    signed char type;          // 1 FIXED, 2 DYNAMIC
    signed char ok;            // 0: depends on a request that has not been computed yet
    long long payload;
};

struct PSlot { unsigned key; unsigned short mid; unsigned short state; long long delta; };
struct RSlot { unsigned short tabid, hid; unsigned state; long long payload; };

// memo tables + request queues of one block (shared memory on the device)
struct SymState {
    PSlot pm[PMEMO];
    RSlot rc[MAXM];
    unsigned short hop[MAXH][3];      // 0 absent, 0xFFFF pending, else hid' + 1
    int hbits[MAXH];                   // dynamicHeaderSizeBits of every header
    int trialBits[MAXT];               // first-minimum header bits over the 56 strategies
    unsigned char trialArg[MAXT];
    unsigned char trialState[MAXT];
    unsigned short qPass[QPASS], qRec[QREC], qTrial[QTRIAL];
    unsigned short qHdr[QHDR];         // hid << 2 | op
    int nqPass, nqRec, nqHdr, nqTrial;
    int nP;                            // used pass memo slots
    int nMasks, nTabs, nHdrs;
    int overflow;                      // a pool filled up during this round
    unsigned doneMulti, doneRun;       // sub-trees that completed in an earlier sweep of this round
    unsigned long long doneAor;
};

D4_HD void sym_reset(SymState& S, int tid, int nthreads) {
    for (int k = tid; k < PMEMO; k += nthreads) { S.pm[k].key = 0; S.pm[k].state = ST_EMPTY; }
    for (int k = tid; k < MAXM; k += nthreads) S.rc[k].state = ST_EMPTY;
    for (int k = tid; k < MAXH * 3; k += nthreads) S.hop[k / 3][k % 3] = 0;
    for (int k = tid; k < MAXT; k += nthreads) S.trialState[k] = ST_EMPTY;
    if (tid == 0) {
        S.nqPass = S.nqRec = S.nqHdr = S.nqTrial = 0;
        S.nP = 0; S.nMasks = 0; S.nTabs = 0; S.nHdrs = 0; S.overflow = 0;
        S.doneMulti = 0; S.doneRun = 0; S.doneAor = 0;
    }
}

D4_HD unsigned pm_key(int mid, int tabid, int op) { return 0x80000000u | ((unsigned)op << 24) | ((unsigned)tabid << 12) | (unsigned)mid; }
D4_HD int pm_key_mid(unsigned k) { return (int)(k & 0xFFF); }
D4_HD int pm_key_tab(unsigned k) { return (int)((k >> 12) & 0xFFF); }
D4_HD int pm_key_op(unsigned k) { return (int)((k >> 24) & 0x7F); }
// slot of `key`, or -1 - (insert position)
D4_HD int pm_find(const SymState& S, unsigned key) {
    unsigned x = key * 0x9E3779B1u;
    x ^= x >> 15;
    unsigned h = x & (PMEMO - 1);
    while (true) {
        const unsigned k = S.pm[h].key;
        if (k == key) return (int)h;
        if (k == 0) return -1 - (int)h;
        h = (h + 1) & (PMEMO - 1);
    }
}

// trace sink of the selection sweep (parity debugging): host tests pass a vector, the device a global buffer
struct TraceSink {
    long long* buf;
    unsigned cap;
    unsigned* n;
};

struct Enumer {
    SymState* S;
    const int* trialAll;       // [tabid * 56 + k] header bits of every strategy (trace only; may be null)
    TraceSink* trace;          // null: no trace
    // the block
    SC B;
    int blockType;             // type of B
    bool storedOK;             // uncompressed length <= 65535: the stored candidate exists (DeflateStream.java:376-383)
    long long storedSize;      // < 0: not compared here (phase A resolves it in the replay)
    unsigned segMask;          // which top-level parts this sweep covers (SEG_*)
    bool select;               // final sweep: every op is resolved, run the selection callback
    // sweep state
    bool poisoned;
    int internalError;
    // selection state (carried across the segments of a segmented round)
    long long bestSize, sizeI, sizeC1, restMin;
    int bestStored, bestArg;
    unsigned candIndex, bestIndex;
    SC best;

    // ---- memo access ---------------------------------------------------------------------------------------
    D4_HD SC bad(SC c) { c.ok = 0; poisoned = true; return c; }
    D4_HD long long size(const SC& c) const { return c.payload + (c.type == 2 ? (long long)S->hbits[c.hid] : 0); }

    D4_HD_BIG SC op_pass(SC c, int op) {
        if (!c.ok) return bad(c);
        if (op >= OP_LEAST0 && op <= OP_LEAST1 && c.type != 2) return c;   // removeDistLitLeastExpensive: DYNAMIC only
        const unsigned key = pm_key(c.mid, c.tabid, op);
        int slot = pm_find(*S, key);
        if (slot >= 0) {
            const PSlot& p = S->pm[slot];
            if (p.state != ST_DONE) return bad(c);
            c.mid = (short)p.mid;
            c.payload -= p.delta;
            return c;
        }
        if (select) { internalError = 1; return bad(c); }
        slot = -1 - slot;
        if (S->nqPass < QPASS && S->nP < PMEMO * 3 / 4) {
            S->pm[slot].key = key; S->pm[slot].state = ST_PENDING;
            S->qPass[S->nqPass++] = (unsigned short)slot;
            S->nP++;
        } else if (S->nP >= PMEMO * 3 / 4) S->overflow = 1;
        return bad(c);
    }
    D4_HD_BIG SC op_recode(SC c) {   // recodeHuffman (:670-743)
        if (!c.ok) return bad(c);
        RSlot& r = S->rc[c.mid];
        if (r.state == ST_DONE) { c.tabid = (short)r.tabid; c.hid = (short)r.hid; c.payload = r.payload; c.type = 2; return c; }
        if (r.state == ST_EMPTY) {
            if (select) { internalError = 2; return bad(c); }
            if (S->nqRec < QREC) { r.state = ST_PENDING; S->qRec[S->nqRec++] = (unsigned short)c.mid; }
        }
        return bad(c);
    }
    D4_HD SC op_recode_less(SC c) { return op_recode(op_pass(c, OP_REPLACE_PRUNE)); }   // recodeHuffmanLessMatches (:655-658)
    D4_HD_BIG SC op_to_fixed(SC c) {   // recodeToFixedHuffman (:637-653)
        if (!c.ok) return bad(c);
        if (c.type == 1) return c;
        const unsigned key = pm_key(c.mid, 0xFFF, OP_FIXED);
        int slot = pm_find(*S, key);
        if (slot >= 0) {
            const PSlot& p = S->pm[slot];
            if (p.state != ST_DONE) return bad(c);
            c.payload = p.delta; c.type = 1; c.tabid = TAB_FIXED; c.hid = -1;
            return c;
        }
        if (select) { internalError = 3; return bad(c); }
        slot = -1 - slot;
        if (S->nqPass < QPASS && S->nP < PMEMO * 3 / 4) {
            S->pm[slot].key = key; S->pm[slot].state = ST_PENDING;
            S->qPass[S->nqPass++] = (unsigned short)slot;
            S->nP++;
        } else if (S->nP >= PMEMO * 3 / 4) S->overflow = 1;
        return bad(c);
    }
    D4_HD_BIG SC op_hdr(SC c, int hop) {
        if (!c.ok) return bad(c);
        if (c.type != 2) return c;
        unsigned short& h = S->hop[c.hid][hop];
        if (h != 0 && h != 0xFFFF) { c.hid = (short)(h - 1); return c; }
        if (h == 0) {
            if (select) { internalError = 4; return bad(c); }
            if (S->nqHdr < QHDR) { h = 0xFFFF; S->qHdr[S->nqHdr++] = (unsigned short)((c.hid << 2) | hop); }
        }
        return bad(c);
    }
    // DeflateBlockHuffman.optimise (:460-469): replace pass, then the header half; *saved = bits saved
    D4_HD_BIG SC op_optimise(SC c, long long* saved) {
        if (!c.ok) return bad(c);
        const long long before = size(c);
        c = op_pass(c, OP_REPLACE);
        if (c.ok && c.type == 2) c = op_hdr(c, HOP_OPT);
        if (c.ok) *saved = before - size(c);
        return c;
    }

    // ---- selection callback (DeflateStream.java:349-368) ---------------------------------------------------
    D4_HD void trace_put(long long idx, long long sz) {
        if (!trace) return;
        const unsigned k = (*trace->n)++;
        if (k < trace->cap) { trace->buf[2 * k] = idx; trace->buf[2 * k + 1] = sz; }
    }
    D4_HD_BIG void cb(const SC& c, bool isRest = true) {
        if (!c.ok) { poisoned = true; return; }
        if (!select) return;
        const long long sz = size(c);
        trace_put(candIndex, sz);
        if (isRest && sz < restMin) restMin = sz;
        if (sz < bestSize) { bestSize = sz; bestStored = 0; bestIndex = candIndex; best = c; bestArg = -1; }
        candIndex++;
    }

    // the 56 header strategy trials of up to 4 bases (addOptimisedRecoded, DeflateStream.java:277-316): per base the
    // first-minimum strategy is the only one the strict `<` of the callback can accept
    D4_HD_BIG void trials(const SC* base, int nb) {
        for (int b = 0; b < nb; b++) {
            const SC& c = base[b];
            if (!c.ok) { poisoned = true; continue; }
            const int t = c.tabid;
            if (S->trialState[t] != ST_DONE) {
                if (select) { internalError = 5; poisoned = true; continue; }
                if (S->trialState[t] == ST_EMPTY) {
                    if (S->nqTrial < QTRIAL) { S->trialState[t] = ST_PENDING; S->qTrial[S->nqTrial++] = (unsigned short)t; }
                    else poisoned = true;   // not queued: this node is not complete yet
                }
                continue;
            }
            if (!select) continue;
            if (trace && trialAll)
                for (int k = 0; k < 56; k++) trace_put(candIndex + k, c.payload + trialAll[t * 56 + k]);
            const long long sz = c.payload + S->trialBits[t];
            if (sz < restMin) restMin = sz;
            if (sz < bestSize) { bestSize = sz; bestStored = 0; bestIndex = candIndex + S->trialArg[t]; best = c; bestArg = S->trialArg[t]; }
            candIndex += 56;
        }
    }

    // recodedHuffmanFull (DeflateStream.java:212-229): x is replaced while a further recodeHuffmanLessMatches shrinks it
    D4_HD_BIG SC recoded_full(SC x, bool* changed) {
        *changed = false;
        if (!x.ok) return bad(x);
        while (true) {
            const SC t = op_recode_less(x);
            if (!t.ok) return bad(x);
            if (size(t) >= size(x)) break;
            x = t;
            *changed = true;
        }
        return x;
    }

    // addOptimisedRecoded (DeflateStream.java:265-317) for base block y; returns the number of bases (0: unknown)
    D4_HD_BIG int aor(const SC& y, int node) {
        if (!select && ((S->doneAor >> node) & 1)) return 0;
        const bool outer = poisoned;
        poisoned = false;
        // the four bases are only ever read by trials(): their Tab and payload.  Every trial rewrites the header from the
        // Tab (optimiseBlockDynBlock -> rewriteHeader, DeflateStream.java:184-198), so the header half of
        // DeflateBlockHuffman.optimise() cannot influence any candidate here and is not run.
        SC base[4];
        base[0] = op_pass(y, OP_REPLACE);                      // optimiseBlockCopyHelper(toOptimise)
        base[1] = op_pass(op_recode(y), OP_REPLACE);           // optimiseBlockHelper(recodedHuffman(.., false))
        const SC pp = op_recode_less(y);                       // pruned
        base[2] = op_pass(pp, OP_REPLACE);                     // optimiseBlockCopyHelper(pruned)
        bool full = false;
        base[3] = recoded_full(pp, &full);                     // prunedFull
        int nb = 0;
        if (base[3].ok) {
            nb = full ? 4 : 3;
            if (full) base[3] = op_pass(base[3], OP_REPLACE);
            trials(base, nb);
        } else poisoned = true;
        if (!poisoned) S->doneAor |= 1ull << node;
        poisoned = poisoned || outer;
        return nb;
    }

    // runOptimisationsCallback (DeflateStream.java:400-442) for block x
    D4_HD_BIG void run(const SC& x, int node) {
        if (!select && ((S->doneRun >> node) & 1)) return;
        const bool outer = poisoned;
        poisoned = false;
        long long saved = 0;
        SC t = op_hdr(x, HOP_RECODE); cb(t);                                       // post
        SC o = op_optimise(t, &saved); if (o.ok) { if (saved > 0) cb(o); } else poisoned = true;   // post optimised
        const int nb = aor(t, node * 3 + 0);
        t = op_hdr(x, HOP_RECODE_LESS); cb(t);                                     // pruned header
        o = op_optimise(t, &saved); if (o.ok) { if (saved > 0) cb(o); } else poisoned = true;
        // addOptimisedRecoded(prune) (:431) re-evaluates candidates with exactly the sizes of the sweep on `post` (both are
        // copies of the same block that differ only in the header, which every base / trial discards), so under the
        // strict `<` none of them can ever be chosen: only the candidate index advances
        if (select) candIndex += 56u * (unsigned)nb;
        aor(op_pass(x, OP_LEAST0), node * 3 + 1);
        aor(op_pass(x, OP_LEAST1), node * 3 + 2);
        if (!poisoned) S->doneRun |= 1u << node;
        poisoned = poisoned || outer;
    }

    // runOptimisationsCallbackMulti (DeflateStream.java:443-463) for seed e
    D4_HD_BIG void multi(const SC& e, int node) {
        if (!select && ((S->doneMulti >> node) & 1)) return;
        const bool outer = poisoned;
        poisoned = false;
        cb(e); run(e, node * 4 + 0);
        SC x = op_recode(e); cb(x); run(x, node * 4 + 1);
        x = op_recode_less(e); cb(x); run(x, node * 4 + 2);
        bool full = false;
        x = recoded_full(x, &full);
        if (!x.ok) poisoned = true;
        else if (full) { cb(x); run(x, node * 4 + 3); }
        if (!poisoned) S->doneMulti |= 1u << node;
        poisoned = poisoned || outer;
    }

    enum { SEG_HEAD = 1, SEG_MULTI_H = 2, SEG_MULTI_O = 4, SEG_FIXED = 8, SEG_LEAST0 = 16, SEG_LEAST1 = 32, SEG_ALL = 63 };

    // start of a round (DeflateStream.optimiseBlock entry): the incumbent is B
    D4_HD void begin_round() {
        bestSize = size(B);
        bestStored = 0; bestArg = -1;
        sizeI = bestSize; sizeC1 = bestSize;
        restMin = 0x7fffffffffffffffll;
        candIndex = 0; bestIndex = 0xffffffffu;
        best = B;
        internalError = 0;
        S->doneMulti = 0; S->doneRun = 0; S->doneAor = 0;
    }

    // DeflateStream.optimiseBlock (:343-490).  Returns true when the sweep raised no request and met no pending op.
    D4_HD_BIG bool sweep(bool sel, unsigned seg) {
        select = sel;
        segMask = seg;
        poisoned = false;
        if (sel && (seg & SEG_HEAD)) trace_put(-1, size(B));
        long long saved = 0;
        SC O = op_optimise(B, &saved);
        bool hasO = O.ok && saved > 0;
        if (!O.ok) poisoned = true;
        if (seg & SEG_HEAD) {
            if (hasO) { cb(O, false); if (sel) sizeC1 = size(O); }
            if (storedOK && sel) {
                if (storedSize >= 0 && storedSize < bestSize) { bestSize = storedSize; bestStored = 1; bestIndex = candIndex; }
                candIndex++;
            }
        }
        SC H = B;
        bool hasOh = hasO;
        if (blockType == 1) {
            H = op_recode(B);
            O = op_optimise(H, &saved);
            hasOh = O.ok && saved > 0;
            if (!O.ok) poisoned = true;
        }
        if (seg & SEG_MULTI_H) multi(H, 0);
        if ((seg & SEG_MULTI_O) && hasOh) multi(O, 1);
        if ((seg & SEG_FIXED) && blockType != 1) {
            long long s2 = 0;
            SC E = op_optimise(op_to_fixed(H), &s2);
            cb(E);
        }
        if (seg & SEG_LEAST0) multi(op_pass(H, OP_LEAST0), 2);
        if (seg & SEG_LEAST1) multi(op_pass(H, OP_LEAST1), 3);
        return !poisoned && S->nqPass == 0 && S->nqRec == 0 && S->nqHdr == 0 && S->nqTrial == 0;
    }
};

}  // namespace d4
