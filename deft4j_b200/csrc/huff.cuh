// huff.cuh — single-thread building blocks of the block cost model:
//   * huff_tree      : HuffmanTree (huffman/HuffmanTree.java:36-128,134-192) including the exact
//                      java.util.PriorityQueue heap mechanics and the bespoke depth
//                      limiter; output = code lengths (codes are canonical, see common.cuh)
//   * hdr_*          : dynamic-header model: RLE packing with the 8 strategy flags
//                      (HuffmanTable.java:42-159), header code build (Huffman.java:117-134), RLE
//                      run -> literal replacement (DeflateBlockHuffman.java:222-296,321-332), trailing
//                      zero code-length trimming (:335-370), recodeHeader (:579-635)
// Each function runs in ONE thread; the engine runs many of them side by side (56 header strategy
// trials per base, litlen and dist trees in different warps).
#pragma once
#include "common.cuh"

namespace d4 {

// Workspace of one tree build.  The algorithm below is written against the members of a workspace type:
//   key_t / ID_BITS   heap entry = (weight << ID_BITS) | node id
//   stack_t / STACK_SHIFT   DFS stack entry = node | (depth << STACK_SHIFT)
//   idx_t / NONE / NDEPTH   node ids, "no node", number of depth slots
template <int MAXLEAF, int MAXN>
struct TreeWs {
    typedef unsigned long long key_t;
    typedef uint32_t stack_t;
    typedef uint16_t idx_t;
    static constexpr int ID_BITS = 16, STACK_SHIFT = 16, NDEPTH = MAXLEAF + 4;
    static constexpr unsigned NONE = 0xFFFF;
    union {                                // the DFS stack is only used once the heap is empty
        key_t heap[MAXLEAF + 2];
        stack_t stack[MAXLEAF + 4];
    };
    idx_t parent[MAXN], left[MAXN], right[MAXN];
    uint8_t side[MAXN];
    idx_t value[MAXLEAF + 2];         // leaf id -> symbol index (dummies included)
    idx_t leafDepth[MAXLEAF + 2];
    idx_t first[MAXLEAF + 4];         // first leaf (DFS order) at each depth == depthMap.get(d).get(0)
};

// Compact workspace for the header code (19 symbols + 2 dummies, <= 46 nodes) when the weights add up to less than
// 1024 — true for header trials, whose weights count RLE pairs of at most 320 code lengths.  305 bytes, so a warp's
// worth of them fits in shared memory next to the engine state.
struct TreeWsCLc {
    typedef uint16_t key_t;
    typedef uint16_t stack_t;
    typedef uint8_t idx_t;
    static constexpr int ID_BITS = 6, STACK_SHIFT = 8, NDEPTH = 25;
    static constexpr unsigned NONE = 0xFF;
    union {
        key_t heap[23];
        stack_t stack[25];
    };
    idx_t parent[46], left[46], right[46];
    uint8_t side[46];
    idx_t value[23];
    idx_t leafDepth[23];
    idx_t first[25];
};

// Returns 0 on success, 1 when the tree cannot be balanced (the reference throws AssertionError there).
// Node ids: [0, nleaf) leaves in insertion order, then internal nodes.  lens[0..n) receives the code lengths (0 for
// unused symbols).
template <class WS>
D4_DEV_BIG int huff_tree_ws(const uint32_t* freq, int n, int limit, uint8_t* lens, WS& ws) {
    typedef typename WS::key_t key_t;
    typedef typename WS::stack_t stack_t;
    typedef typename WS::idx_t idx_t;
    constexpr int IDB = WS::ID_BITS, SSH = WS::STACK_SHIFT;
    constexpr unsigned IDM = (1u << IDB) - 1u, SNM = (1u << SSH) - 1u;
    int hs = 0;  // heap size
    auto W = [](key_t k) { return (unsigned long long)k >> IDB; };
    auto add = [&](key_t x) {  // PriorityQueue.offer + siftUp
        int k = hs++;
        while (k > 0) {
            int p = (k - 1) >> 1;
            key_t e = ws.heap[p];
            if (W(x) >= W(e)) break;
            ws.heap[k] = e;
            k = p;
        }
        ws.heap[k] = x;
    };
    auto poll = [&]() {  // PriorityQueue.poll + siftDown
        key_t result = ws.heap[0];
        int s = --hs;
        key_t x = ws.heap[s];
        if (s > 0) {
            int k = 0, half = s >> 1;
            while (k < half) {
                int child = 2 * k + 1;
                key_t c = ws.heap[child];
                int r = child + 1;
                if (r < s) {
                    key_t cr = ws.heap[r];
                    if (W(c) > W(cr)) { c = cr; child = r; }
                }
                if (W(x) <= W(c)) break;
                ws.heap[k] = c;
                k = child;
            }
            ws.heap[k] = x;
        }
        return result;
    };
    for (int i = 0; i < n; i++) lens[i] = 0;
    int nleaf = 0;
    for (int i = 0; i < n; i++)
        if (freq[i] > 0) {
            ws.value[nleaf] = (idx_t)i;
            add((key_t)(((unsigned long long)freq[i] << IDB) | (unsigned)nleaf));
            nleaf++;
        }
    int index = 0;
    while (hs < 2) {  // dummy leaves (HuffmanTree.java:50-58)
        if (index >= n || freq[index] == 0) {
            ws.value[nleaf] = (idx_t)index;
            add((key_t)((1ull << IDB) | (unsigned)nleaf));
            nleaf++;
        }
        index++;
    }
    int nn = nleaf;
    const int total = hs;
    for (int i = 0; i < total - 1; i++) {
        key_t l = poll(), r = poll();
        int id = nn++;
        int li = (int)(l & IDM), ri = (int)(r & IDM);
        ws.left[id] = (idx_t)li; ws.right[id] = (idx_t)ri;
        ws.parent[li] = (idx_t)id; ws.side[li] = 0;
        ws.parent[ri] = (idx_t)id; ws.side[ri] = 1;
        add((key_t)(((W(l) + W(r)) << IDB) | (unsigned)id));
    }
    const int root = (int)(poll() & IDM);
    ws.parent[root] = (idx_t)WS::NONE;
    int maxDepth = 0;
    auto traverse = [&]() {  // HuffmanTree.traverse (:134-158), left-first DFS
        for (int d = 0; d < WS::NDEPTH; d++) ws.first[d] = (idx_t)WS::NONE;
        maxDepth = 0;
        int sp = 0;
        ws.stack[sp++] = (stack_t)root;
        while (sp > 0) {
            stack_t e = ws.stack[--sp];
            int node = (int)(e & SNM), d = (int)(e >> SSH);
            if (d > maxDepth) maxDepth = d;
            if (node >= nleaf) {
                ws.stack[sp++] = (stack_t)((unsigned)ws.right[node] | ((unsigned)(d + 1) << SSH));
                ws.stack[sp++] = (stack_t)((unsigned)ws.left[node] | ((unsigned)(d + 1) << SSH));
            } else {
                if (ws.first[d] == (idx_t)WS::NONE) ws.first[d] = (idx_t)node;
                ws.leafDepth[node] = (idx_t)d;
            }
        }
    };
    traverse();
    while (maxDepth > limit) {  // HuffmanTree.java:75-127
        int leafA = ws.first[maxDepth];
        int parent1 = ws.parent[leafA];
        int leafB = (ws.side[leafA] == 0) ? ws.right[parent1] : ws.left[parent1];
        int parent2 = ws.parent[parent1];
        if (ws.side[parent1] == 0) { ws.left[parent2] = (idx_t)leafB; ws.side[leafB] = 0; }
        else                       { ws.right[parent2] = (idx_t)leafB; ws.side[leafB] = 1; }
        ws.parent[leafB] = (idx_t)parent2;
        bool moved = false;
        for (int i = maxDepth - 2; i >= 1; i--) {
            if (ws.first[i] != (idx_t)WS::NONE) {
                int leafC = ws.first[i];
                int parent3 = ws.parent[leafC];
                int sideC = ws.side[leafC];
                const int in = parent1;  // parent1 has just left the tree: its slot becomes the new internal node
                ws.left[in] = (idx_t)leafA; ws.parent[leafA] = (idx_t)in; ws.side[leafA] = 0;
                ws.right[in] = (idx_t)leafC; ws.parent[leafC] = (idx_t)in; ws.side[leafC] = 1;
                if (sideC == 0) { ws.left[parent3] = (idx_t)in; ws.side[in] = 0; }
                else            { ws.right[parent3] = (idx_t)in; ws.side[in] = 1; }
                ws.parent[in] = (idx_t)parent3;
                moved = true;
                break;
            }
        }
        if (!moved) return 1;
        traverse();
    }
    for (int l = 0; l < nleaf; l++)
        if (ws.value[l] < n) lens[ws.value[l]] = (uint8_t)ws.leafDepth[l];
    return 0;
}
template <int MAXLEAF, int MAXN>
D4_DEV int huff_tree(const uint32_t* freq, int n, int limit, uint8_t* lens, TreeWs<MAXLEAF, MAXN>& ws) {
    return huff_tree_ws(freq, n, limit, lens, ws);
}

// ---- the litlen tree on the fast path ---------------------------------------------------------------------------------
// HuffmanTree (huffman/HuffmanTree.java:36-73) with the PriorityQueue mechanics of huff_tree_ws, for the common case that
// the tree needs no depth limiting: 32-bit heap keys (weight << 10 | node id), parent links only.
//   huff_tree_fast_build : the merge loop (one thread).  freq[] is overwritten: parent[] lives in the same storage.
//                          Returns nleaf | root << 16.
//   huff_tree_fast_depths: code length of every leaf by walking up to the root; `lane` of `nlanes` takes every nlanes-th
//                          leaf (a warp does this side by side).  Returns the deepest leaf seen by this lane.
// A tree deeper than the limit is rebuilt by huff_tree_ws, whose limiter needs the full node arrays.
D4_DEV_BIG uint32_t huff_tree_fast_build(uint32_t* freq, int n, uint32_t* heap, uint16_t* value) {
    uint16_t* parent = reinterpret_cast<uint16_t*>(freq);
    int hs = 0;
    auto add = [&](uint32_t x) {
        int k = hs++;
        while (k > 0) {
            const int p = (k - 1) >> 1;
            const uint32_t e = heap[p];
            if ((x >> 10) >= (e >> 10)) break;
            heap[k] = e;
            k = p;
        }
        heap[k] = x;
    };
    auto poll = [&]() {
        const uint32_t result = heap[0];
        const int s = --hs;
        const uint32_t x = heap[s];
        if (s > 0) {
            int k = 0;
            const int half = s >> 1;
            while (k < half) {
                int child = 2 * k + 1;
                uint32_t c = heap[child];
                const int r = child + 1;
                if (r < s) {
                    const uint32_t cr = heap[r];
                    if ((c >> 10) > (cr >> 10)) { c = cr; child = r; }
                }
                if ((x >> 10) <= (c >> 10)) break;
                heap[k] = c;
                k = child;
            }
            heap[k] = x;
        }
        return result;
    };
    int nleaf = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t f = freq[i];
        if (f > 0) { value[nleaf] = (uint16_t)i; add((f << 10) | (uint32_t)nleaf); nleaf++; }
    }
    int index = 0;
    while (hs < 2) {  // dummy leaves (HuffmanTree.java:50-58)
        if (index >= n || freq[index] == 0) { value[nleaf] = (uint16_t)index; add((1u << 10) | (uint32_t)nleaf); nleaf++; }
        index++;
    }
    // from here on freq[] is dead and parent[] takes its storage
    int nn = nleaf;
    const int total = hs;
    for (int i = 0; i < total - 1; i++) {
        const uint32_t l = poll(), r = poll();
        const int id = nn++;
        parent[l & 1023u] = (uint16_t)id;
        parent[r & 1023u] = (uint16_t)id;
        add((((l >> 10) + (r >> 10)) << 10) | (uint32_t)id);
    }
    const uint32_t root = poll() & 1023u;
    return (uint32_t)nleaf | (root << 16);
}
D4_DEV int huff_tree_fast_depths(const uint16_t* parent, const uint16_t* value, int n, uint32_t nleafRoot, uint8_t* lens, int lane,
                                 int nlanes) {
    const int nleaf = (int)(nleafRoot & 0xFFFF), root = (int)(nleafRoot >> 16);
    int maxDepth = 0;
    for (int l = lane; l < nleaf; l += nlanes) {
        int d = 0;
        for (int node = l; node != root; node = parent[node]) d++;
        if (d > maxDepth) maxDepth = d;
        if (value[l] < n) lens[value[l]] = (uint8_t)d;
    }
    return maxDepth;
}
// both steps in one thread: 0 and the code lengths, or 2 when the tree is deeper than `limit`
D4_DEV int huff_tree_fast(uint32_t* freq, int n, int limit, uint8_t* lens, uint32_t* heap, uint16_t* value) {
    const uint32_t nr = huff_tree_fast_build(freq, n, heap, value);
    return huff_tree_fast_depths(reinterpret_cast<const uint16_t*>(freq), value, n, nr, lens, 0, 1) > limit ? 2 : 0;
}

// ---- the header code on the fast path -----------------------------------------------------------------------------------
// The same for the 19-symbol header code (limit 7) when the weights add up to less than 1024 (they count RLE pairs of at
// most 320 code lengths): 16-bit heap keys (weight << 6 | node id) and byte parent links — 96 bytes of workspace, small
// enough to give every thread of a CTA its own in shared memory.  The leaf -> symbol map is not stored: real leaves are the
// symbols with freq > 0 in order, dummies follow at the first indices with freq == 0 (HuffmanTree.java:41-58).
// Returns 0 and the code lengths, or 2 when the tree is deeper than `limit` (the caller runs huff_tree_ws instead).
D4_DEV_BIG int huff_tree_tiny(const uint32_t* freq, int n, int limit, uint8_t* lens, uint16_t* heap, uint8_t* parent) {
    int hs = 0;
    auto add = [&](uint32_t x) {
        int k = hs++;
        while (k > 0) {
            const int p = (k - 1) >> 1;
            const uint32_t e = heap[p];
            if ((x >> 6) >= (e >> 6)) break;
            heap[k] = (uint16_t)e;
            k = p;
        }
        heap[k] = (uint16_t)x;
    };
    auto poll = [&]() {
        const uint32_t result = heap[0];
        const int s = --hs;
        const uint32_t x = heap[s];
        if (s > 0) {
            int k = 0;
            const int half = s >> 1;
            while (k < half) {
                int child = 2 * k + 1;
                uint32_t c = heap[child];
                const int r = child + 1;
                if (r < s) {
                    const uint32_t cr = heap[r];
                    if ((c >> 6) > (cr >> 6)) { c = cr; child = r; }
                }
                if ((x >> 6) <= (c >> 6)) break;
                heap[k] = (uint16_t)c;
                k = child;
            }
            heap[k] = (uint16_t)x;
        }
        return result;
    };
    int nleaf = 0;
    for (int i = 0; i < n; i++) {
        lens[i] = 0;
        if (freq[i] > 0) { add((freq[i] << 6) | (uint32_t)nleaf); nleaf++; }
    }
    const int nreal = nleaf;
    int index = 0;
    while (hs < 2) {
        if (index >= n || freq[index] == 0) { add((1u << 6) | (uint32_t)nleaf); nleaf++; }
        index++;
    }
    int nn = nleaf;
    const int total = hs;
    for (int i = 0; i < total - 1; i++) {
        const uint32_t l = poll(), r = poll();
        const int id = nn++;
        parent[l & 63u] = (uint8_t)id;
        parent[r & 63u] = (uint8_t)id;
        add((((l >> 6) + (r >> 6)) << 6) | (uint32_t)id);
    }
    const int root = (int)(poll() & 63u);
    int maxDepth = 0, leaf = 0;
    for (int i = 0; i < n; i++) {          // real leaves, in insertion order
        if (freq[i] == 0) continue;
        int d = 0;
        for (int node = leaf; node != root; node = parent[node]) d++;
        if (d > maxDepth) maxDepth = d;
        lens[i] = (uint8_t)d;
        leaf++;
    }
    for (int i = 0; leaf < nleaf; i++) {   // dummies: they receive a code when their index is inside the alphabet (H3)
        if (i < n && freq[i] != 0) continue;
        int d = 0;
        for (int node = leaf; node != root; node = parent[node]) d++;
        if (d > maxDepth) maxDepth = d;
        if (i < n) lens[i] = (uint8_t)d;
        leaf++;
    }
    (void)nreal;
    return maxDepth > limit ? 2 : 0;
}
constexpr int TINY_WS_BYTES = 100;   // heap u16[24] at +0, parent u8[46] at +48; 25 words: conflict-free across a warp
// workspace adaptor: the fast path in `tiny` (shared memory on the device), the full algorithm in `slow` when needed
struct TreeWsTiny {
    uint16_t* heap;
    uint8_t* parent;
    TreeWsCLc* slow;
};
D4_DEV int huff_tree_ws(const uint32_t* freq, int n, int limit, uint8_t* lens, TreeWsTiny& ws) {
    const int rc = huff_tree_tiny(freq, n, limit, lens, ws.heap, ws.parent);
    if (rc != 2) return rc;
    return huff_tree_ws(freq, n, limit, lens, *ws.slow);
}

using TreeWsCL = TreeWs<21, 46>;

// ---- header model ---------------------------------------------------------------------------------
// strategy flags: bit0 ohh, bit1 use8, bit2 use7, bit3 alt8, bit4 noRep, bit5 noZRep, bit6 noZRep2,
// bit7 noRepZeros, bit8 prune (DeflateStream.java:184-198,277-316)
D4_CONST uint16_t c_trial_flags[56] = {
    0x7, 0x3, 0x5, 0x0, 0x47, 0x43, 0x45, 0x40, 0x27, 0x23, 0x25, 0x20, 0x67, 0x63, 0x65, 0x60, 0x10, 0x50, 0x30,
    0x70, 0x107, 0x103, 0x105, 0x100, 0x147, 0x143, 0x145, 0x140, 0x127, 0x123, 0x125, 0x120, 0x167, 0x163, 0x165,
    0x160, 0x110, 0x150, 0x130, 0x170, 0xa7, 0xa3, 0xa5, 0xa0, 0xe7, 0xe3, 0xe5, 0xe0, 0x1a7, 0x1a3, 0x1a5, 0x1a0,
    0x1e7, 0x1e3, 0x1e5, 0x1e0};
constexpr int FLAGS_DEFAULT = 0x7;  // rewriteHeader() defaults (DeflateBlockHuffman.java:480-482)

D4_DEV int pair_extra_bits(int sym) { return sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0; }

// getRLEPairSize (:133-163)
D4_DEV int pair_size(uint16_t p, const uint8_t* CL) {
    int s = pair_sym(p);
    return CL[s] + (pair_run(p) > 0 ? pair_extra_bits(s) : 0);
}

// removeDynHeaderTrailingZeroLenCodelens (:335-364); returns bits removed
D4_DEV int hdr_trim(Hdr& h) {
    int saved = 0;
    while (true) {
        int lastZero = -1, lastNonZero = h.ncl;
        for (int i = 0; i < h.ncl; i++) {
            if (h.CL[c_codelen_order[i]] == 0) lastZero = i; else lastNonZero = i;
        }
        if (lastZero > lastNonZero) { h.ncl = (uint8_t)lastZero; saved += 3; } else break;
    }
    return saved;
}

// Huffman.ofRLEPacked (Huffman.java:117-134) on the pair list
D4_DEV int hdr_build_code(Hdr& h, TreeWsCL& ws) {
    uint32_t freq[19];
    for (int i = 0; i < 19; i++) freq[i] = 0;
    for (int i = 0; i < h.np; i++) freq[pair_sym(h.pairs[i])]++;
    return huff_tree<21, 46>(freq, 19, 7, h.CL, ws);
}

D4_DEV int hdr_pairs_bits(const Hdr& h) {
    int b = 0;
    for (int i = 0; i < h.np; i++) b += pair_size(h.pairs[i], h.CL);
    return b;
}

// rewriteHeader (:484-577): pack (HuffmanTable.java:70-159) straight into pairs, build the header code,
// size it, trim.
D4_DEV_BIG int hdr_rewrite(const Tab& t, int flags, Hdr& h, TreeWsCL& ws) {
    const bool ohh = flags & 1, use8 = flags & 2, use7 = flags & 4, alt8 = flags & 8, noRep = flags & 16,
               noZRep = flags & 32, noZRep2 = flags & 64, noRepZeros = flags & 128;
    const int nL = t.nL, n = t.nL + t.nD;
    int np = 0;
    auto get = [&](int i) -> int { return i < nL ? t.L[i] : t.D[i - nL]; };
    auto emit = [&](int sym, int run, int val) { h.pairs[np++] = pair_pack(sym, run, val); };
    int last = get(0), runLength = 1;
    for (int i = 1; i <= n; i++) {
        if (i < n && get(i) == last) { runLength++; continue; }
        if (last == 0) {
            if (!noZRep2) {
                while (runLength >= 138) { emit(18, 138, 0); runLength -= 138; }
                if (runLength >= 11) { emit(18, runLength, 0); runLength = 0; }
            }
            if (!noZRep) {
                while (runLength >= 10) { emit(17, 10, 0); runLength -= 10; }
                if (runLength >= 3) { emit(17, runLength, 0); runLength = 0; }
            }
        }
        if (!noRep && runLength > 0 && (!noRepZeros || last != 0)) {
            emit(last, 0, last);
            runLength--;
            int j = 6;
            while (j >= 3) {
                if (ohh) {
                    if (use8 && runLength == 8) {
                        emit(16, alt8 ? 5 : 4, last); emit(16, alt8 ? 3 : 4, last);
                        runLength -= 8;
                        break;
                    }
                    if (use7 && runLength == 7) {
                        emit(16, 4, last); emit(16, 3, last);
                        runLength -= 7;
                        break;
                    }
                }
                if (runLength - j >= 0) { emit(16, j, last); runLength -= j; } else j--;
            }
        }
        while (runLength > 0) { emit(last, 0, last); runLength--; }
        if (i < n) { last = get(i); runLength = 1; }
    }
    h.np = (uint16_t)np;
    if (hdr_build_code(h, ws)) return 1;
    h.ncl = 19;
    h.bits = 5 + 5 + 4 + 19 * 3 + hdr_pairs_bits(h);
    h.bits -= hdr_trim(h);
    return 0;
}

// replaceRLERunsWithLiteralsIfSmaller (:321-332) via replaceWithLiteralsIfSmaller (:222-296): a run is
// replaced by `run` plain lengths when those cost less (prune: no more) than the run code; only when
// the repeated value has a header code.  In-place expansion from the back.
D4_DEV_BIG void hdr_replace_runs(Hdr& h, bool prune) {
    int newNp = 0, saved = 0;
    bool any = false;
    for (int i = 0; i < h.np; i++) {
        uint16_t p = h.pairs[i];
        int run = pair_run(p);
        int add = 1;
        if (run > 0) {
            int size = pair_size(p, h.CL);
            int b = h.CL[pair_val(p)];
            int tot = b * run;
            if (b >= 1 && (prune ? tot <= size : tot < size)) { add = run; saved += size - tot; any = true; }
        }
        newNp += add;
    }
    if (!any) return;
    int w = newNp;
    for (int i = h.np - 1; i >= 0; i--) {
        uint16_t p = h.pairs[i];
        int run = pair_run(p);
        bool rep = false;
        if (run > 0) {
            int size = pair_size(p, h.CL);
            int b = h.CL[pair_val(p)];
            int tot = b * run;
            rep = b >= 1 && (prune ? tot <= size : tot < size);
        }
        if (rep) {
            int v = pair_val(p);
            for (int k = 0; k < run; k++) h.pairs[--w] = pair_pack(v, 0, v);
        } else {
            h.pairs[--w] = p;
        }
    }
    h.np = (uint16_t)newNp;
    h.bits -= saved;
}

// recodeHeader (:579-629): new header code from the existing pairs; numCodelenLens is NOT reset
// (SURVEY.md H8), only trimmed further.
D4_DEV_BIG int hdr_recode(Hdr& h, TreeWsCL& ws) {
    if (hdr_build_code(h, ws)) return 1;
    hdr_trim(h);
    h.bits = 5 + 5 + 4 + h.ncl * 3 + hdr_pairs_bits(h);
    return 0;
}
// recodeHeaderToLessRLEMatches (:632-635)
D4_DEV int hdr_recode_less(Hdr& h, TreeWsCL& ws) {
    hdr_replace_runs(h, true);
    return hdr_recode(h, ws);
}
// optimiseHeader (:471-476)
D4_DEV void hdr_optimise(Hdr& h) {
    h.bits -= hdr_trim(h);
    hdr_replace_runs(h, false);
}
// optimiseBlockDynBlock (DeflateStream.java:184-198): one header strategy trial
D4_DEV int hdr_trial(const Tab& t, int flags, Hdr& h, TreeWsCL& ws) {
    if (hdr_rewrite(t, flags & 0xFF, h, ws)) return 1;
    if (flags & 0x100) { if (hdr_recode_less(h, ws)) return 1; }
    hdr_optimise(h);
    return 0;
}

// ---- size-only header trials over a run list ---------------------------------------------------------
// The candidate enumerator only needs the SIZE of a header strategy trial (optimiseBlockDynBlock,
// DeflateStream.java:184-198); the winner alone is materialised (hdr_trial).  trial_sizes() evaluates one
// rewrite strategy (the 8 flags of HuffmanTable.pack) for both values of `prune` without storing the RLE
// pairs: the pair stream of rewriteHeader is re-generated from the runs of equal code lengths each time it
// is needed (three times), in exactly the order hdr_rewrite emits it, and everything else the reference
// derives from the pairs is a sum over them:
//   A  symbol frequencies                                   -> header code CL1 (Huffman.ofRLEPacked), trimmed ncl1
//   B  under CL1: pair bits; runs replaceRLERunsWithLiteralsIfSmaller would expand with `<` (prune = false:
//      optimiseHeader) and with `<=` (prune = true: recodeHeaderToLessRLEMatches) -> size without prune, and the
//      frequencies after the `<=` expansion                  -> CL2 (recodeHeader), ncl2 trimmed on from ncl1 (H8)
//   C  under CL2: expanded runs as plain lengths, the others as pairs that optimiseHeader may still expand (`<`).
struct RunList {
    uint16_t n;
    uint8_t val[MAX_PAIRS];
    uint16_t len[MAX_PAIRS];
};

// runs of equal values over litlen lengths followed by distance lengths (they may straddle the boundary,
// HuffmanTable.java:70-159)
D4_DEV void runlist_build(const Tab& t, RunList& r) {
    const int nL = t.nL, n = t.nL + t.nD;
    int k = 0;
    int last = t.nL ? t.L[0] : t.D[0], run = 1;
    for (int i = 1; i <= n; i++) {
        const int v = i < n ? (i < nL ? t.L[i] : t.D[i - nL]) : -1;
        if (v == last) { run++; continue; }
        r.val[k] = (uint8_t)last; r.len[k] = (uint16_t)run; k++;
        last = v; run = 1;
    }
    r.n = (uint16_t)k;
}

// the pairs hdr_rewrite emits for one run, in its order: f(sym, run, repeated value, count)
template <class F>
D4_DEV void emit_run(int last, int runLength, int flags, F&& f) {
    const bool ohh = flags & 1, use8 = flags & 2, use7 = flags & 4, alt8 = flags & 8, noRep = flags & 16,
               noZRep = flags & 32, noZRep2 = flags & 64, noRepZeros = flags & 128;
    if (last == 0) {
        if (!noZRep2) {
            if (runLength >= 138) { f(18, 138, 0, runLength / 138); runLength %= 138; }
            if (runLength >= 11) { f(18, runLength, 0, 1); runLength = 0; }
        }
        if (!noZRep) {
            if (runLength >= 10) { f(17, 10, 0, runLength / 10); runLength %= 10; }
            if (runLength >= 3) { f(17, runLength, 0, 1); runLength = 0; }
        }
    }
    if (!noRep && runLength > 0 && (!noRepZeros || last != 0)) {
        f(last, 0, last, 1);
        runLength--;
        int j = 6;
        while (j >= 3) {
            if (ohh) {
                if (use8 && runLength == 8) { f(16, alt8 ? 5 : 4, last, 1); f(16, alt8 ? 3 : 4, last, 1); runLength -= 8; break; }
                if (use7 && runLength == 7) { f(16, 4, last, 1); f(16, 3, last, 1); runLength -= 7; break; }
            }
            if (runLength - j >= 0) { f(16, j, last, 1); runLength -= j; } else j--;
        }
    }
    if (runLength > 0) f(last, 0, last, runLength);
}

// removeTrailingHeaderCodes on a bare length array: trims on from ncl (never grows it)
D4_DEV int trim_ncl(const uint8_t* CL, int ncl) {
    while (true) {
        int lastZero = -1, lastNonZero = ncl;
        for (int i = 0; i < ncl; i++) { if (CL[c_codelen_order[i]] == 0) lastZero = i; else lastNonZero = i; }
        if (lastZero > lastNonZero) ncl = lastZero; else break;
    }
    return ncl;
}

// The pairs of one run as at most nine (symbol, run field, count) groups in closed form — what emit_run produces, without
// its data-dependent loops: 28 threads of a warp evaluate 28 different strategies side by side, and a loop whose trip
// count depends on the flags would serialise them.  The repeated value of every group is the run's value.
struct RunGroups { int sym[9], run[9], cnt[9]; };
D4_DEV void run_groups(int v, int n, int flags, RunGroups& g) {
    const bool ohh = flags & 1, use8 = flags & 2, use7 = flags & 4, alt8 = flags & 8, noRep = flags & 16,
               noZRep = flags & 32, noZRep2 = flags & 64, noRepZeros = flags & 128;
    const bool z18 = v == 0 && !noZRep2, z17 = v == 0 && !noZRep;
    int c = z18 ? n / 138 : 0;
    g.sym[0] = 18; g.run[0] = 138; g.cnt[0] = c; n -= 138 * c;
    c = (z18 && n >= 11) ? 1 : 0;
    g.sym[1] = 18; g.run[1] = n; g.cnt[1] = c; if (c) n = 0;
    c = z17 ? n / 10 : 0;
    g.sym[2] = 17; g.run[2] = 10; g.cnt[2] = c; n -= 10 * c;
    c = (z17 && n >= 3) ? 1 : 0;
    g.sym[3] = 17; g.run[3] = n; g.cnt[3] = c; if (c) n = 0;
    const bool rep = !noRep && n > 0 && (!noRepZeros || v != 0);
    g.sym[4] = v; g.run[4] = 0; g.cnt[4] = rep ? 1 : 0;
    int m = rep ? n - 1 : 0;                      // what the 16-runs may cover
    const int rest = rep ? 0 : n;                 // no repeat codes: everything is spelled out
    const int r6 = m % 6;
    const bool sp8 = ohh && use8 && m >= 8 && r6 == 2;           // the walk m, m-6, ... stops at exactly 8
    const bool sp7 = !sp8 && ohh && use7 && m >= 7 && r6 == 1;   // ... or at exactly 7
    const int k6 = sp8 ? (m - 8) / 6 : sp7 ? (m - 7) / 6 : m / 6;
    int rem = (sp8 || sp7) ? 0 : r6;
    g.sym[5] = 16; g.run[5] = 6; g.cnt[5] = k6;
    const bool tail16 = !(sp8 || sp7) && rem >= 3;
    g.sym[6] = 16; g.run[6] = sp8 ? (alt8 ? 5 : 4) : sp7 ? 4 : rem; g.cnt[6] = (sp8 || sp7 || tail16) ? 1 : 0;
    if (tail16) rem = 0;
    g.sym[7] = 16; g.run[7] = sp8 ? (alt8 ? 3 : 4) : 3; g.cnt[7] = (sp8 || sp7) ? 1 : 0;
    g.sym[8] = v; g.run[8] = 0; g.cnt[8] = rest + rem;
}

// sizes in bits of the trials (flags, prune = false) and (flags, prune = true); returns 1 when a tree cannot be
// balanced (the reference throws).  Three walks over the runs: (0) pair frequencies -> header code c1; (1) sizes under
// c1, what optimiseHeader / the prune step expand, frequencies after the prune expansion -> c2; (2) sizes under c2.
// The groups of a run use four symbols only - 18, 17, 16 and the run's own value - so the counters of the three run
// codes live in registers and the frequency arrays are touched once per run.  One loop nest for all three walks: the
// code is expanded once (the engine kernel is instruction-cache bound).
template <class WS>
D4_DEV int trial_sizes(const RunList& rl, int flags, int* bitsNoPrune, int* bitsPrune, WS& ws) {
    uint32_t f[19];
    uint8_t c1[19], c2[19];
    int ncl1 = 19, sizeSum = 0, saved = 0, sum2 = 0;
    int s16 = 0, s17 = 0, s18 = 0, t16 = 0, t17 = 0, t18 = 0;
    RunGroups g;
#pragma unroll 1
    for (int walk = 0; walk < 3; walk++) {
        for (int i = 0; i < 19; i++) f[i] = 0;
        uint32_t a16 = 0, a17 = 0, a18 = 0;
#pragma unroll 1
        for (int r = 0; r < rl.n; r++) {
            const int val = rl.val[r], n = rl.len[r];
            // A run too short for any run code (zeros: 17 needs 3; other values: a length plus a 16 of 3) is spelled out
            // whatever the strategy: most runs of a code-length array are like that, and the test does not depend on the
            // flags, so the 28 strategies of a warp take this shortcut together.
            const bool plain = n < 3 || (val != 0 && n < 4);
            if (plain) {
                if (walk == 0) f[val] += (uint32_t)n;
                else if (walk == 1) { sizeSum += c1[val] * n; f[val] += (uint32_t)n; }
                else sum2 += c2[val] * n;
                continue;
            }
            run_groups(val, n, flags, g);
            const int lit = g.cnt[4] + g.cnt[8];   // plain lengths of the run's value (groups 4 and 8)
            if (walk == 0) {
                a18 += (uint32_t)(g.cnt[0] + g.cnt[1]);
                a17 += (uint32_t)(g.cnt[2] + g.cnt[3]);
                a16 += (uint32_t)(g.cnt[5] + g.cnt[6] + g.cnt[7]);
                f[val] += (uint32_t)lit;
                continue;
            }
            const int b1 = c1[val], b2 = walk == 2 ? c2[val] : 0;
            uint32_t toVal = (uint32_t)lit;
            if (walk == 1) sizeSum += b1 * lit; else sum2 += b2 * lit;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k == 4) continue;
                const int size1 = k < 2 ? s18 : k < 4 ? s17 : s16;
                const int run = g.run[k], cnt = g.cnt[k];
                const int tot = b1 * run;
                const bool can = b1 >= 1;
                const bool expand = can && tot <= size1;      // the prune step turns the run into `run` plain lengths
                if (walk == 1) {
                    sizeSum += size1 * cnt;
                    if (can && tot < size1) saved += (size1 - tot) * cnt;   // optimiseHeader does so only when strictly smaller
                    toVal += expand ? (uint32_t)(run * cnt) : 0u;
                    const uint32_t keep = expand ? 0u : (uint32_t)cnt;
                    if (k < 2) a18 += keep; else if (k < 4) a17 += keep; else a16 += keep;
                } else {
                    int bits;
                    if (expand) bits = run * b2;
                    else {
                        bits = k < 2 ? t18 : k < 4 ? t17 : t16;
                        if (b2 >= 1 && b2 * run < bits) bits = b2 * run;    // expanded by optimiseHeader
                    }
                    sum2 += bits * cnt;
                }
            }
            if (walk == 1) f[val] += toVal;
        }
        if (walk == 2) break;
        f[16] += a16; f[17] += a17; f[18] += a18;
        uint8_t* c = walk == 0 ? c1 : c2;
        if (huff_tree_ws(f, 19, 7, c, ws)) return 1;
        if (walk == 0) {
            ncl1 = trim_ncl(c1, 19);
            s16 = c1[16] + 2; s17 = c1[17] + 3; s18 = c1[18] + 7;
        } else {
            *bitsNoPrune = 5 + 5 + 4 + 3 * ncl1 + sizeSum - saved;
            const int ncl2 = trim_ncl(c2, ncl1);
            sum2 = 3 * ncl2;
            t16 = c2[16] + 2; t17 = c2[17] + 3; t18 = c2[18] + 7;
        }
    }
    *bitsPrune = 5 + 5 + 4 + sum2;
    return 0;
}

}  // namespace d4
