/*
 * deft_oracle.cpp — CPU ORACLE: literal restatement of deft4j-base's optimiser path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include, link or call this file.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/deft4j-base/src/main/java/com/github/NeRdTheNed/deft4j/).
 *
 * Parity status: PINNED against the reference's nine golden pairs (tests/test_oracle_golden.py).
 *
 * Written from a reading of the reference's behaviour; the object model (block list, copy-on-write
 * symbol lists, candidate enumeration order, java.util.PriorityQueue heap mechanics) is restated so
 * that tie-breaks come out identically.  decodedVal byte arrays are represented as views into one
 * stream-wide decoded buffer (the bytes are the same; only the storage differs).
 */
#include "deft_oracle.h"

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace {

ora_stats g_stats;
// candidate trace (ora_trace_begin): {index within the optimiseBlock call, size}; {-1, incumbent} opens a call
int64_t* g_trace = nullptr;
size_t g_trace_cap = 0, g_trace_n = 0;
long g_trace_idx = 0;
void tracePut(long idx, long sz) {
    if (!g_trace) return;
    if (g_trace_n < g_trace_cap) { g_trace[2 * g_trace_n] = idx; g_trace[2 * g_trace_n + 1] = sz; }
    g_trace_n++;
}

// ---------------------------------------------------------------------------------------------
// Constants (deflate/Constants.java:9-128): the RFC 1951 tables.
// ---------------------------------------------------------------------------------------------
const int LITLEN_TBL_OFFSET = 257, LITLEN_EOB = 256, LITLEN_MAX = 285, DISTSYM_MAX = 29;
const int MAX_LEN = 258;
const int MAX_CODELEN_LENS = 19, MIN_CODELEN_LENS = 4, MIN_LITLEN_LENS = 257, MIN_DIST_LENS = 1;
const int MAX_LITLEN_LENS = 288, MAX_DIST_LENS = 32;
const int codelen_lengths_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
const int len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                          67, 83, 99, 115, 131, 163, 195, 227, 258};
const int len_ebits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                           4, 4, 4, 4, 5, 5, 5, 5, 0};
const int dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                           1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const int dist_ebits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                            9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// Constants.len2litlen (Constants.java:15-23): 284 when edgecase, else the RFC mapping.
int len2litlen(int len, bool edgecase) {
    if (edgecase) return 284;
    if (len == 258) return 285;
    for (int i = 27; i >= 0; i--)
        if (len >= len_base[i]) return 257 + i;
    return 0xffff;
}
// Constants.distance2dist (Constants.java:9-13).
int distance2dist(long distance) {
    for (int i = 29; i >= 0; i--)
        if (distance >= dist_base[i]) return i;
    return 0;
}

// Util.rev (util/Util.java:244-254): reverse `size` bits (the extra iteration with shift -1 is a no-op
// for in-range codes).
int rev(int bits, int size) {
    int r = 0;
    for (int i = 0; i < size; i++) r |= ((bits >> i) & 1) << (size - 1 - i);
    return r;
}

// ---------------------------------------------------------------------------------------------
// BitInputStream (io/BitInputStream.java:59-88): LSB-first, sticky EOF returning -1.
// ---------------------------------------------------------------------------------------------
struct BitIn {
    const uint8_t* d;
    size_t n, pos = 0;
    long accum = 0;
    int bitpos = 0;
    bool eof = false;
    long read8() {
        if (eof) return -1;
        if (pos >= n) { pos++; eof = true; return -1; }
        return d[pos++];
    }
    long readBit() {
        if (eof) return -1;
        if (bitpos == 0) {
            bitpos = 8;
            accum = read8();
            if (eof) return -1;
        }
        long r = accum & 1;
        accum >>= 1;
        bitpos--;
        return r;
    }
    long readBits(int count) {
        if (eof) return -1;
        long read = 0, amount = 0;
        while (count > 0) {
            bool aligned = (count >= 8) && (bitpos == 0);
            long v = aligned ? read8() : readBit();
            long a = aligned ? 8 : 1;
            if (eof) return -1;
            read |= v << amount;
            amount += a;
            count -= (int)a;
        }
        return read;
    }
    void readToByteAligned() { if (bitpos != 0) readBits(bitpos); }
    // bytes pulled from the InputStream (a failed read() at EOF does not consume)
    size_t consumed() const { return pos > n ? n : pos; }
};

// ---------------------------------------------------------------------------------------------
// BitOutputStream (io/BitOutputStream.java:31-69): LSB-first bit packing.
// ---------------------------------------------------------------------------------------------
struct BitOut {
    std::vector<uint8_t> buf;
    uint64_t acc = 0;
    int nacc = 0;  // bits pending (< 8 after each call)
    void writeNBits(uint64_t bits, int n) {
        for (int i = 0; i < n; i++) {  // bit at a time, as the reference does (n can reach 48)
            acc |= ((bits >> i) & 1ull) << nacc;
            if (++nacc == 8) { buf.push_back((uint8_t)acc); acc = 0; nacc = 0; }
        }
    }
    void flushToByteAligned() { if (nacc != 0) writeNBits(0, 8 - nacc); }
    void writeBytes(const uint8_t* p, size_t n) {
        if (nacc == 0) buf.insert(buf.end(), p, p + n);
        else for (size_t i = 0; i < n; i++) writeNBits(p[i], 8);
    }
};

// ---------------------------------------------------------------------------------------------
// HuffmanTable / Huffman (huffman/HuffmanTable.java:14-31, huffman/Huffman.java:35-64,97-115,170-213)
// ---------------------------------------------------------------------------------------------
struct Table {
    std::vector<int> code, codeLen;
    explicit Table(int n) : code(n, 0), codeLen(n, 0) {}
    int getSymLen(int n) const { return codeLen[n]; }           // Huffman.java:211-213
    int getSym(int n) const { return rev(code[n], codeLen[n]); }  // Huffman.java:207-209
};
using TableP = std::shared_ptr<Table>;

// Huffman.buildCodes (Huffman.java:35-64): canonical codes, ascending over the *used* lengths; no
// validity check (over/under-subscribed sets are accepted).
TableP ofCodelens(const std::vector<int>& codelens) {
    auto t = std::make_shared<Table>((int)codelens.size());
    t->codeLen = codelens;
    int nextCode = 0, lastShift = 0;
    for (int length = 1; length <= 15; length++) {
        bool used = false;
        for (int v : codelens) if (v == length) { used = true; break; }
        if (!used) continue;
        nextCode <<= length - lastShift;
        lastShift = length;
        for (size_t i = 0; i < codelens.size(); i++)
            if (codelens[i] == length) t->code[i] = nextCode++;
    }
    return t;
}

// Decoder side of Huffman (constructor Huffman.java:97-115 + readSymbol :170-197): read one bit at a
// time, after each bit look for a symbol of that length with that code (first match), give up after 15.
struct Decoder {
    // per length: list of (code, symbol) in symbol order  == codeMap + codes.indexOf
    std::vector<std::pair<int, int>> byLen[16];
    explicit Decoder(const Table& t) {
        for (size_t i = 0; i < t.codeLen.size(); i++) {
            int l = t.codeLen[i];
            if (l > 0 && l < 16) byLen[l].push_back({t.code[i], (int)i});
        }
    }
    // returns symbol or -1; *codeLen receives the length (0 on failure)
    int readSym(BitIn& is, int* codeLenOut) const {
        int code = 0, codeLen = 0;
        while (true) {
            if (codeLen == 15) { *codeLenOut = 0; return -1; }
            code <<= 1;
            code |= (int)is.readBits(1);
            codeLen++;
            for (auto& cs : byLen[codeLen])
                if (cs.first == code) { *codeLenOut = codeLen; return cs.second; }
        }
    }
};

// HuffmanTable.LIT / DIST (HuffmanTable.java:166-209): 286 / 30 entries.
TableP fixedLit() {
    static TableP t;
    if (!t) {
        t = std::make_shared<Table>(286);
        int next = 0;
        for (int i = 256; i <= 279; i++) { t->code[i] = next++; t->codeLen[i] = 7; }
        next <<= 1;
        for (int i = 0; i <= 143; i++) { t->code[i] = next++; t->codeLen[i] = 8; }
        for (int i = 280; i <= 285; i++) { t->code[i] = next++; t->codeLen[i] = 8; }
        next += 2;
        next <<= 1;
        for (int i = 144; i <= 255; i++) { t->code[i] = next++; t->codeLen[i] = 9; }
    }
    return t;
}
TableP fixedDist() {
    static TableP t;
    if (!t) {
        t = std::make_shared<Table>(30);
        for (int i = 0; i <= 29; i++) { t->code[i] = i; t->codeLen[i] = 5; }
    }
    return t;
}

// ---------------------------------------------------------------------------------------------
// HuffmanTree (huffman/HuffmanTree.java:36-128,134-192,218-221) incl. java.util.PriorityQueue
// ---------------------------------------------------------------------------------------------
struct Node {
    Node* parent = nullptr;
    int side = 0;
    int weight = 0;
    bool leaf = false;
    int value = 0;          // leaf
    Node *left = nullptr, *right = nullptr;  // internal
};

// java.util.PriorityQueue<Node> with natural ordering (weight difference): array binary heap with
// OpenJDK's siftUp / siftDown.
struct JavaPQ {
    std::vector<Node*> q;
    static int cmp(const Node* a, const Node* b) { return a->weight - b->weight; }
    void add(Node* x) {
        size_t k = q.size();
        q.push_back(x);
        while (k > 0) {
            size_t parent = (k - 1) >> 1;
            Node* e = q[parent];
            if (cmp(x, e) >= 0) break;
            q[k] = e;
            k = parent;
        }
        q[k] = x;
    }
    Node* remove() {
        Node* result = q[0];
        size_t s = q.size() - 1;
        Node* x = q[s];
        q.pop_back();
        if (s != 0) {
            size_t k = 0, half = s >> 1;
            while (k < half) {
                size_t child = 2 * k + 1;
                Node* c = q[child];
                size_t right = child + 1;
                if (right < s && cmp(c, q[right]) > 0) c = q[child = right];
                if (cmp(x, c) <= 0) break;
                q[k] = c;
                k = child;
            }
            q[k] = x;
        }
        return result;
    }
    size_t size() const { return q.size(); }
};

struct HuffmanTree {
    int numSymbols;
    std::map<int, std::vector<Node*>> depthMap;  // TreeMap: ascending depth; lists in DFS order
    int maxDepth = 0;
    std::vector<std::unique_ptr<Node>> pool;

    Node* newLeaf(int value, int weight) {
        pool.emplace_back(new Node());
        Node* n = pool.back().get();
        n->leaf = true; n->value = value; n->weight = weight;
        return n;
    }
    // InternalNode constructor (HuffmanTree.java:243-251)
    Node* newInternal(Node* l, Node* r) {
        pool.emplace_back(new Node());
        Node* n = pool.back().get();
        l->parent = n; l->side = 0; n->left = l;
        r->parent = n; r->side = 1; n->right = r;
        n->weight = l->weight + r->weight;
        return n;
    }
    void traverse(Node* node, int depth) {  // :145-158
        if (depth > maxDepth) maxDepth = depth;
        if (!node->leaf) {
            traverse(node->left, depth + 1);
            traverse(node->right, depth + 1);
        } else {
            depthMap[depth].push_back(node);
        }
    }
    void traverseRoot(Node* root) { depthMap.clear(); maxDepth = 0; traverse(root, 0); }  // :134-138

    HuffmanTree(const std::vector<int>& freq, int limit) : numSymbols((int)freq.size()) {  // :36-128
        JavaPQ queue;
        for (int i = 0; i < numSymbols; i++)
            if (freq[i] > 0) queue.add(newLeaf(i, freq[i]));
        int index = 0;
        while (queue.size() < 2) {  // :50-58 dummy leaves (index may run past numSymbols)
            if (index >= numSymbols || freq[index] == 0) queue.add(newLeaf(index, 1));
            index++;
        }
        const int n = (int)queue.size();
        for (int i = 0; i < n - 1; i++) {
            Node* left = queue.remove();
            Node* right = queue.remove();
            queue.add(newInternal(left, right));
        }
        Node* root = queue.remove();
        traverseRoot(root);
        while (maxDepth > limit) {  // :75-127 bespoke depth limiter
            Node* leafA = depthMap[maxDepth][0];
            Node* parent1 = leafA->parent;
            Node* leafB = (leafA->side == 0) ? parent1->right : parent1->left;
            Node* parent2 = parent1->parent;
            if (parent1->side == 0) { parent2->left = leafB; leafB->parent = parent2; leafB->side = 0; }
            else                    { parent2->right = leafB; leafB->parent = parent2; leafB->side = 1; }
            bool moved = false;
            for (int i = maxDepth - 2; i >= 1; i--) {
                auto it = depthMap.find(i);
                if (it != depthMap.end()) {  // (lists in the map are never empty)
                    Node* leafC = it->second[0];
                    Node* parent3 = leafC->parent;
                    int sideC = leafC->side;  // read before the constructor re-parents leafC
                    Node* in = newInternal(leafA, leafC);
                    if (sideC == 0) { parent3->left = in; in->parent = parent3; in->side = 0; }
                    else            { parent3->right = in; in->parent = parent3; in->side = 1; }
                    moved = true;
                    break;
                }
            }
            if (!moved) { fprintf(stderr, "oracle: Can't balance the tree\n"); abort(); }
            traverseRoot(root);
        }
    }

    TableP getTable() {  // :164-192
        auto table = std::make_shared<Table>(numSymbols);
        int nextCode = 0, lastShift = 0;
        for (auto& entry : depthMap) {
            int length = entry.first;
            nextCode <<= length - lastShift;
            lastShift = length;
            auto leaves = entry.second;
            std::stable_sort(leaves.begin(), leaves.end(),
                             [](const Node* a, const Node* b) { return a->value < b->value; });
            for (Node* leaf : leaves) {
                if (leaf->value < numSymbols) {
                    table->code[leaf->value] = nextCode;
                    table->codeLen[leaf->value] = length;
                }
                nextCode++;
            }
        }
        return table;
    }
};

TableP buildTree(const std::vector<int>& freq, int limit) {
    if (limit == 7) g_stats.tree_builds_small++; else g_stats.tree_builds_big++;
    HuffmanTree t(freq, limit);
    return t.getTable();
}

// Huffman.ofRLEPacked (Huffman.java:117-134): frequencies of the 19 header symbols, skipping the
// run-length operand after 16/17/18; depth limit 7.
TableP ofRLEPacked(const std::vector<int>& flat) {
    std::vector<int> lenFreq(MAX_CODELEN_LENS, 0);
    for (size_t i = 0; i < flat.size(); i++) {
        int s = flat[i];
        lenFreq[s]++;
        if (s == 16 || s == 17 || s == 18) i++;
    }
    return buildTree(lenFreq, 7);
}

// HuffmanTable.pack (HuffmanTable.java:70-159)
void pack(std::vector<int>& lengths, const std::vector<int>& codeLen, bool ohh, bool use8, bool use7,
          bool alt8, bool noRep, bool noZRep, bool noZRep2, bool noRepZeros) {
    const int n = (int)codeLen.size();
    int last = codeLen[0];
    int runLength = 1;
    for (int i = 1; i <= n; i++) {
        if (i < n && codeLen[i] == last) {
            runLength++;
        } else {
            if (last == 0) {
                if (!noZRep2) {
                    int j = 138;
                    while (j >= 11) {
                        if (runLength - j >= 0) { lengths.push_back(18); lengths.push_back(j - 11); runLength -= j; }
                        else j--;
                    }
                }
                if (!noZRep) {
                    int j = 10;
                    while (j >= 3) {
                        if (runLength - j >= 0) { lengths.push_back(17); lengths.push_back(j - 3); runLength -= j; }
                        else j--;
                    }
                }
            }
            if (!noRep && runLength > 0 && (!noRepZeros || last != 0)) {
                lengths.push_back(last);
                runLength--;
                int j = 6;
                while (j >= 3) {
                    if (ohh) {
                        if (use8 && runLength == 8) {
                            lengths.push_back(16); lengths.push_back((alt8 ? 5 : 4) - 3);
                            lengths.push_back(16); lengths.push_back((alt8 ? 3 : 4) - 3);
                            runLength -= 8;
                            break;
                        }
                        if (use7 && runLength == 7) {
                            lengths.push_back(16); lengths.push_back(4 - 3);
                            lengths.push_back(16); lengths.push_back(3 - 3);
                            runLength -= 7;
                            break;
                        }
                    }
                    if (runLength - j >= 0) { lengths.push_back(16); lengths.push_back(j - 3); runLength -= j; }
                    else j--;
                }
            }
            while (runLength > 0) { lengths.push_back(last); runLength--; }
            if (i < n) { last = codeLen[i]; runLength = 1; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LitLen (deflate/LitLen.java:29-47).  decodedVal is a view (off, n) into the stream's decoded buffer.
// ---------------------------------------------------------------------------------------------
struct LitLen {
    int32_t dist;     // 0 for literal / EOB
    int32_t litlen;   // literal value, 256 for EOB, or the match length
    bool edgecase;
    uint64_t off;     // decodedVal = out[off .. off + n)
    uint32_t nDecoded() const { return dist > 0 ? (uint32_t)litlen : (litlen < 256 ? 1u : 0u); }
};
// Header "RLE pair" (a LitLen in the reference): dist = run length (0 = plain length value),
// sym = header symbol, decodedVal = `val` repeated max(dist,1) times.
struct Pair {
    int32_t dist;
    int32_t sym;
    uint8_t val;
};

enum BType { STORED = 0, FIXED = 1, DYNAMIC = 2 };

struct Stream;

// One class for the three block kinds (DeflateBlock.java, DeflateBlockUncompressed.java,
// DeflateBlockHuffman.java); `type` selects the behaviour.
struct Block {
    Stream* st;
    BType type;
    Block* prev = nullptr;
    std::shared_ptr<Block> next;
    // decoded / stored data: view into st->out
    uint64_t dataOff = 0, dataLen = 0;
    // Huffman state
    TableP litlenDec, distDec, codeLenDec;
    std::shared_ptr<std::vector<LitLen>> litlens;
    bool didCopyLitLens = false;
    long sizeBits = 0, litlenSizeBits = 0, dynamicHeaderSizeBits = 0;
    int numLitlenLens = 0, numDistLens = 0, numCodelenLens = 0;
    std::shared_ptr<std::vector<Pair>> rlePairs;
    bool didCopyRLEPairs = false;

    Block(Stream* s, BType t, Block* p) : st(s), type(t), prev(p) {}
};
using BlockP = std::shared_ptr<Block>;

struct Stream {
    std::vector<uint8_t> out;  // all decoded bytes of the stream, in order
    BlockP first;
};

// DeflateBlockUncompressed.getSizeBits (DeflateBlockUncompressed.java:70-74) /
// DeflateBlockHuffman.getSizeBits (DeflateBlockHuffman.java:1169-1172)
long getSizeBits(const Block& b, long alignment) {
    if (b.type == STORED) {
        long c = alignment % 8;
        c = c == 0 ? 0 : 8 - c;
        return ((long)b.dataLen + 4) * 8 + c;
    }
    return b.sizeBits;
}

// DeflateBlockHuffman.getLitLenSize (DeflateBlockHuffman.java:112-131)
int getLitLenSize(const LitLen& l, const Table& lit, const Table& dist) {
    if (l.dist > 0) {
        int litlen = len2litlen(l.litlen, l.edgecase);
        long nbits = lit.getSymLen(litlen);
        nbits += len_ebits[litlen - LITLEN_TBL_OFFSET];
        int d = distance2dist(l.dist);
        nbits += dist.getSymLen(d);
        nbits += dist_ebits[d];
        return (int)nbits;
    }
    return lit.getSymLen(l.litlen);
}
// getRLEPairSize (:133-163)
int getRLEPairSize(const Pair& p, const Table& cl) {
    int s = cl.getSymLen(p.sym);
    if (p.dist > 0) {
        switch (p.sym) {
            case 16: s += 2; break;
            case 17: s += 3; break;
            case 18: s += 7; break;
            default: fprintf(stderr, "oracle: invalid RLE symbol\n"); abort();
        }
    }
    return s;
}

// DeflateBlock.copy(old,new) + DeflateBlockHuffman.copy (:1174-1205) / DeflateBlockUncompressed.copy (:83-91)
BlockP copyBlock(const Block& b) {
    auto c = std::make_shared<Block>(b.st, b.type, b.prev);
    c->next = b.next;
    c->dataOff = b.dataOff; c->dataLen = b.dataLen;
    if (b.type != STORED) {
        c->litlenDec = b.litlenDec; c->distDec = b.distDec;
        c->litlens = b.litlens;            // shared until mutated (didCopyLitLens = false)
        c->sizeBits = b.sizeBits; c->litlenSizeBits = b.litlenSizeBits;
        if (b.type == DYNAMIC) {
            c->dynamicHeaderSizeBits = b.dynamicHeaderSizeBits;
            c->codeLenDec = b.codeLenDec;
            c->numLitlenLens = b.numLitlenLens; c->numDistLens = b.numDistLens;
            c->numCodelenLens = b.numCodelenLens;
            c->rlePairs = b.rlePairs;
        }
    }
    return c;
}
void ensureDidCopyLitLens(Block& b) {  // :298-303
    if (!b.didCopyLitLens) { b.litlens = std::make_shared<std::vector<LitLen>>(*b.litlens); b.didCopyLitLens = true; }
}
void ensureDidCopyRLEPairs(Block& b) {  // :305-310
    if (!b.didCopyRLEPairs) { b.rlePairs = std::make_shared<std::vector<Pair>>(*b.rlePairs); b.didCopyRLEPairs = true; }
}

// DeflateBlock.discard (DeflateBlock.java:36-51) + subclass field clearing
void discardBlock(Block& b) {
    Block* prev = b.prev;
    BlockP next = b.next;
    if (prev && prev->next.get() == &b) prev->next = nullptr;
    if (next && next->prev == &b) next->prev = nullptr;
    b.next = nullptr;
    b.prev = nullptr;
    b.litlenDec = nullptr; b.distDec = nullptr; b.codeLenDec = nullptr;
    b.litlens = nullptr; b.rlePairs = nullptr;
}
// DeflateBlock.replace (DeflateBlock.java:92-100)
void replaceBlock(Block& b, const BlockP& replacement) {
    if (b.prev) b.prev->next = replacement;
    if (b.next) b.next->prev = replacement.get();
}
// DeflateBlock.remove (DeflateBlock.java:134-144)
void removeBlock(Block& b) {
    BlockP keepNext = b.next;
    if (b.prev) b.prev->next = b.next;
    if (b.next) b.next->prev = b.prev;
    discardBlock(b);
}
// DeflateStream.setFirstBlock (DeflateStream.java:151-157)
void setFirstBlock(Stream& s, const BlockP& nb) {
    if (s.first) replaceBlock(*s.first, nb);
    s.first = nb;
}

// replaceWithLiteralsIfSmaller for symbols (DeflateBlockHuffman.java:222-296, litLen == true branch)
long replaceSymsWithLiterals(std::vector<LitLen>& list, const Table& lit, const Table& dist, bool prune,
                             bool estimateOnly, const uint8_t* out) {
    g_stats.symbol_passes++;
    long savedTotal = 0, seenRemove = 0;
    std::vector<LitLen> result;
    if (!estimateOnly) result.reserve(list.size());
    for (const LitLen& check : list) {
        bool doReplace = false;
        int checkSize = 0, totalSize = 0;
        if (check.dist != 0) {
            checkSize = getLitLenSize(check, lit, dist);
            doReplace = true;
            for (int i = 0; i < check.litlen; i++) {
                int b = out[check.off + i];
                int bSize = lit.getSymLen(b);
                if (bSize < 1) { doReplace = false; break; }
                totalSize += bSize;
                if (prune ? totalSize > checkSize : totalSize >= checkSize) { doReplace = false; break; }
            }
        }
        if (doReplace) {
            savedTotal += checkSize - totalSize;
            seenRemove++;
            if (!estimateOnly)
                for (int i = 0; i < check.litlen; i++)
                    result.push_back(LitLen{0, out[check.off + i], false, check.off + i});
        } else if (!estimateOnly) {
            result.push_back(check);
        }
    }
    if (estimateOnly && savedTotal <= 0 && seenRemove <= 0) return -1;
    if (!estimateOnly) list.swap(result);
    return savedTotal;
}
// same, header pairs branch (decoder = codeLenDec, distDec == null)
long replacePairsWithLiterals(std::vector<Pair>& list, const Table& cl, bool prune, bool estimateOnly) {
    long savedTotal = 0, seenRemove = 0;
    std::vector<Pair> result;
    for (const Pair& check : list) {
        bool doReplace = false;
        int checkSize = 0, totalSize = 0;
        if (check.dist != 0) {
            checkSize = getRLEPairSize(check, cl);
            doReplace = true;
            for (int i = 0; i < check.dist; i++) {
                int bSize = cl.getSymLen(check.val);
                if (bSize < 1) { doReplace = false; break; }
                totalSize += bSize;
                if (prune ? totalSize > checkSize : totalSize >= checkSize) { doReplace = false; break; }
            }
        }
        if (doReplace) {
            savedTotal += checkSize - totalSize;
            seenRemove++;
            if (!estimateOnly)
                for (int i = 0; i < check.dist; i++) result.push_back(Pair{0, check.val, check.val});
        } else if (!estimateOnly) {
            result.push_back(check);
        }
    }
    if (estimateOnly && savedTotal <= 0 && seenRemove <= 0) return -1;
    if (!estimateOnly) list.swap(result);
    return savedTotal;
}

const uint8_t* outOf(const Block& b);

// replaceBackrefsWithLiteralsIfSmaller (:312-319)
void replaceBackrefs(Block& b, bool prune) {
    if (replaceSymsWithLiterals(*b.litlens, *b.litlenDec, *b.distDec, prune, true, outOf(b)) >= 0) {
        ensureDidCopyLitLens(b);
        long saved = replaceSymsWithLiterals(*b.litlens, *b.litlenDec, *b.distDec, prune, false, outOf(b));
        b.sizeBits -= saved;
        b.litlenSizeBits -= saved;
    }
}
// replaceRLERunsWithLiteralsIfSmaller (:321-332)
void replaceRLERuns(Block& b, bool prune) {
    if (b.type != DYNAMIC) return;
    if (replacePairsWithLiterals(*b.rlePairs, *b.codeLenDec, prune, true) >= 0) {
        ensureDidCopyRLEPairs(b);
        long saved = replacePairsWithLiterals(*b.rlePairs, *b.codeLenDec, prune, false);
        b.sizeBits -= saved;
        b.dynamicHeaderSizeBits -= saved;
    }
}
// removeDynHeaderTrailingZeroLenCodelens (:335-364)
long removeTrailingZeroCodelens(Block& b) {
    if (b.type != DYNAMIC) return 0;
    int lastZero = -1, lastNonZero = b.numCodelenLens;
    for (int i = 0; i < b.numCodelenLens; i++) {
        int l = b.codeLenDec->codeLen[codelen_lengths_order[i]];
        if (l == 0) lastZero = i; else lastNonZero = i;
    }
    if (lastZero > lastNonZero) {
        b.numCodelenLens = lastZero;
        return 3 + removeTrailingZeroCodelens(b);
    }
    return 0;
}
void removeTrailingHeaderCodes(Block& b) {  // :366-370
    long saved = removeTrailingZeroCodelens(b);
    b.sizeBits -= saved;
    b.dynamicHeaderSizeBits -= saved;
}

// removeDistLitLeastExpensive (:373-458)
void removeDistLitLeastExpensive(Block& b, int mode) {
    if (b.type != DYNAMIC) return;
    g_stats.symbol_passes++;
    int litSize[MAX_DIST_LENS] = {0}, litFreq[MAX_DIST_LENS] = {0};
    bool litNoAllow[MAX_DIST_LENS] = {false}, litSeen[MAX_DIST_LENS] = {false};
    const uint8_t* out = outOf(b);
    for (const LitLen& check : *b.litlens) {
        if (check.dist != 0) {
            int litlen = len2litlen(check.litlen, check.edgecase) - LITLEN_TBL_OFFSET;
            if (litNoAllow[litlen]) continue;
            litSeen[litlen] = true;
            int checkSize = getLitLenSize(check, *b.litlenDec, *b.distDec);
            int totalSize = 0;
            bool blocked = false;
            for (int i = 0; i < check.litlen; i++) {
                int bSize = b.litlenDec->getSymLen(out[check.off + i]);
                if (bSize < 1) { litNoAllow[litlen] = true; blocked = true; break; }
                totalSize += bSize;
            }
            if (blocked) continue;
            litSize[litlen] += totalSize - checkSize;
            litFreq[litlen]++;
        }
    }
    int litlenRem = -1, litlenRemSize = 0, litlenRemFreq = 0;
    for (int i = 0; i < MAX_DIST_LENS; i++) {
        if (!litNoAllow[i] && litSeen[i]) {
            bool doRem = mode == 1 ? litFreq[i] < litlenRemFreq : litSize[i] < litlenRemSize;
            if (litlenRem == -1 || doRem) { litlenRem = i; litlenRemSize = litSize[i]; litlenRemFreq = litFreq[i]; }
        }
    }
    if (litlenRem >= 0) {
        ensureDidCopyLitLens(b);
        std::vector<LitLen> result;
        result.reserve(b.litlens->size());
        for (const LitLen& check : *b.litlens) {
            if (check.dist != 0 && len2litlen(check.litlen, check.edgecase) - LITLEN_TBL_OFFSET == litlenRem) {
                for (int i = 0; i < check.litlen; i++)
                    result.push_back(LitLen{0, out[check.off + i], false, check.off + i});
            } else {
                result.push_back(check);
            }
        }
        b.litlens->swap(result);
    }
    b.sizeBits += litlenRemSize;
    b.litlenSizeBits += litlenRemSize;
}

// optimiseHeader (:471-476)
long optimiseHeader(Block& b) {
    long original = b.sizeBits;
    removeTrailingHeaderCodes(b);
    replaceRLERuns(b, false);
    return original - b.sizeBits;
}
// optimise (:460-469; DeflateBlockUncompressed.java:77-81 returns 0)
long optimiseBlockInPlace(Block& b) {
    if (b.type == STORED) return 0;
    long original = b.sizeBits;
    replaceBackrefs(b, false);
    optimiseHeader(b);
    return original - b.sizeBits;
}

// rewriteHeader (:484-577)
void rewriteHeader(Block& b, bool ohh, bool use8, bool use7, bool alt8, bool noRep, bool noZRep, bool noZRep2,
                   bool noRepZeros) {
    if (b.type != DYNAMIC) return;
    g_stats.header_rewrites++;
    b.sizeBits -= b.dynamicHeaderSizeBits;
    b.dynamicHeaderSizeBits = 0;
    b.numLitlenLens = (int)b.litlenDec->codeLen.size();
    b.numDistLens = (int)b.distDec->codeLen.size();
    b.didCopyRLEPairs = true;
    std::vector<int> combined(b.litlenDec->codeLen);  // Util.combine (HuffmanTable.java:42-46)
    combined.insert(combined.end(), b.distDec->codeLen.begin(), b.distDec->codeLen.end());
    std::vector<int> repack;
    pack(repack, combined, ohh, use8, use7, alt8, noRep, noZRep, noZRep2, noRepZeros);
    b.codeLenDec = ofRLEPacked(repack);
    b.numCodelenLens = MAX_CODELEN_LENS;
    b.dynamicHeaderSizeBits = 5 + 5 + 4 + MAX_CODELEN_LENS * 3L;
    int i = 0;
    size_t it = 0;
    const int combinedLens = b.numLitlenLens + b.numDistLens;
    b.rlePairs = std::make_shared<std::vector<Pair>>();
    b.rlePairs->reserve(combinedLens);
    uint8_t prevLast = 0;
    while (i < combinedLens) {
        int sym = repack[it++];
        int dist;
        uint8_t val;
        if (sym >= 0 && sym <= 15) {
            dist = 0; val = (uint8_t)sym; i++;
        } else {
            dist = repack[it++];
            switch (sym) {
                case 16: dist += 3; val = prevLast; break;
                case 17: dist += 3; val = 0; break;
                case 18: dist += 11; val = 0; break;
                default: fprintf(stderr, "oracle: invalid RLE symbol when encoding\n"); abort();
            }
            i += dist;
        }
        Pair p{dist, sym, val};
        b.rlePairs->push_back(p);
        b.dynamicHeaderSizeBits += getRLEPairSize(p, *b.codeLenDec);
        prevLast = val;
    }
    b.sizeBits += b.dynamicHeaderSizeBits;
    removeTrailingHeaderCodes(b);
}

// recodeHeader (:579-629)
void recodeHeader(Block& b) {
    if (b.type != DYNAMIC) return;
    b.sizeBits -= b.dynamicHeaderSizeBits;
    b.dynamicHeaderSizeBits = 0;
    std::vector<int> lengths;
    for (const Pair& p : *b.rlePairs) {
        lengths.push_back(p.sym);
        if (p.dist > 0) lengths.push_back(p.dist);
    }
    b.codeLenDec = ofRLEPacked(lengths);
    removeTrailingZeroCodelens(b);  // keeps the (possibly stale) numCodelenLens as the start (H8)
    b.dynamicHeaderSizeBits = 5 + 5 + 4 + b.numCodelenLens * 3L;
    for (const Pair& p : *b.rlePairs) b.dynamicHeaderSizeBits += getRLEPairSize(p, *b.codeLenDec);
    b.sizeBits += b.dynamicHeaderSizeBits;
}
// recodeHeaderToLessRLEMatches (:632-635)
void recodeHeaderToLessRLEMatches(Block& b) {
    replaceRLERuns(b, true);
    recodeHeader(b);
}

// recodeToHuffmanInternal (:759-770)
void recodeToHuffmanInternal(Block& b, const TableP& lit, const TableP& dist) {
    g_stats.symbol_passes++;
    b.litlenDec = lit;
    b.distDec = dist;
    b.sizeBits -= b.litlenSizeBits;
    b.litlenSizeBits = 0;
    for (const LitLen& l : *b.litlens) b.litlenSizeBits += getLitLenSize(l, *lit, *dist);
    b.sizeBits += b.litlenSizeBits;
}
// recodeToFixedHuffman (:637-653)
void recodeToFixedHuffman(Block& b) {
    if (b.type == FIXED) return;
    b.sizeBits -= b.dynamicHeaderSizeBits;
    b.type = FIXED;
    b.dynamicHeaderSizeBits = 0;
    b.codeLenDec = nullptr;
    b.numLitlenLens = b.numDistLens = b.numCodelenLens = 0;
    b.rlePairs = nullptr;
    b.didCopyRLEPairs = true;
    recodeToHuffmanInternal(b, fixedLit(), fixedDist());
}
// recodeHuffman (:670-743) with MIN_DIST_CODES = MIN_LIT_CODES = 0 (:667-668)
void recodeHuffman(Block& b) {
    g_stats.symbol_passes++;
    std::vector<int> litFreqTemp(MAX_LITLEN_LENS - 2, 0), distFreqTemp(MAX_DIST_LENS - 2, 0);
    for (const LitLen& l : *b.litlens) {
        if (l.dist > 0) {
            litFreqTemp[len2litlen(l.litlen, l.edgecase)]++;
            distFreqTemp[distance2dist(l.dist)]++;
        } else {
            litFreqTemp[l.litlen]++;
        }
    }
    int lastNonZeroLit = (int)litFreqTemp.size();
    while (lastNonZeroLit > 0 && litFreqTemp[lastNonZeroLit - 1] == 0) lastNonZeroLit--;
    int lastNonZeroDist = (int)distFreqTemp.size();
    while (lastNonZeroDist > 0 && distFreqTemp[lastNonZeroDist - 1] == 0) lastNonZeroDist--;
    const int realLastNonZeroDist = lastNonZeroDist;
    std::vector<int> litFreq(litFreqTemp.begin(), litFreqTemp.begin() + lastNonZeroLit);
    std::vector<int> distFreq(distFreqTemp.begin(), distFreqTemp.begin() + lastNonZeroDist);
    const bool handleZero = realLastNonZeroDist == 0;
    TableP newLit = buildTree(litFreq, 15);
    int nonZero = 0;
    for (int v : distFreq) if (v != 0) nonZero++;
    const bool handleOne = !handleZero && nonZero <= 1;
    TableP newDist;
    if (handleZero || handleOne) {
        newDist = std::make_shared<Table>(handleZero ? 1 : realLastNonZeroDist);
        if (handleOne) {
            newDist->codeLen[realLastNonZeroDist - 1] = 1;
            newDist->code[realLastNonZeroDist - 1] = 0;
        }
    } else {
        newDist = buildTree(distFreq, 15);
    }
    // recodeToHuffman (:745-757): never the FIXED instances here
    b.type = DYNAMIC;
    recodeToHuffmanInternal(b, newLit, newDist);
    rewriteHeader(b, true, true, true, false, false, false, false, false);
}
// recodeHuffmanLessMatches (:655-658)
void recodeHuffmanLessMatches(Block& b) {
    replaceBackrefs(b, true);
    recodeHuffman(b);
}

// DeflateBlock.asUncompressed (DeflateBlock.java:53-62)
BlockP asUncompressed(const Block& b) {
    if (b.type != STORED) {
        auto u = std::make_shared<Block>(b.st, STORED, b.prev);
        u->next = b.next;
        u->dataOff = b.dataOff; u->dataLen = b.dataLen;
        return u;
    }
    return copyBlock(b);
}

// canMerge (DeflateBlockUncompressed.java:99-109, DeflateBlockHuffman.java:1228-1230)
bool canMerge(const Block& a, const Block* append) {
    if (a.type == STORED) return append && (a.dataLen + append->dataLen <= 65535);
    return append && (append->type == FIXED || append->type == DYNAMIC);
}
// merge (DeflateBlockUncompressed.java:111-117, DeflateBlockHuffman.java:1233-1271)
BlockP mergeBlocksPair(const BlockP& self, const BlockP& append) {
    if (self->type == STORED) {
        auto m = std::make_shared<Block>(self->st, STORED, self->prev);
        m->next = append->next;
        m->dataOff = self->dataOff;
        m->dataLen = self->dataLen + append->dataLen;
        return m;
    }
    if (append->type == FIXED || append->type == DYNAMIC) {
        BlockP thisFixed = self, otherFixed = append;
        if (self->type == DYNAMIC) { thisFixed = copyBlock(*self); recodeToFixedHuffman(*thisFixed); }
        if (append->type == DYNAMIC) { otherFixed = copyBlock(*append); recodeToFixedHuffman(*otherFixed); }
        BlockP merged = copyBlock(*thisFixed);
        merged->dataOff = self->dataOff;
        merged->dataLen = self->dataLen + append->dataLen;
        ensureDidCopyLitLens(*merged);
        LitLen eob = merged->litlens->back();
        merged->litlens->pop_back();
        merged->litlens->insert(merged->litlens->end(), otherFixed->litlens->begin(), otherFixed->litlens->end());
        merged->sizeBits -= merged->litlenSizeBits;
        merged->litlenSizeBits += otherFixed->litlenSizeBits;
        merged->litlenSizeBits -= getLitLenSize(eob, *merged->litlenDec, *merged->distDec);
        merged->sizeBits += merged->litlenSizeBits;
        merged->next = append->next;
        return merged;
    }
    return nullptr;
}

const uint8_t* outOf(const Block& b) { return b.st->out.data(); }

// ---------------------------------------------------------------------------------------------
// Parsing (DeflateStream.java:72-126, DeflateBlockUncompressed.java:23-36,
// DeflateBlockHuffman.java:778-1031)
// ---------------------------------------------------------------------------------------------
bool parseStored(Block& b, BitIn& is) {
    is.readToByteAligned();
    long len = is.readBits(16) & 0xffff;
    long nlen = is.readBits(16) & 0xffff;
    if (nlen != (~len & 0xffff)) return false;
    b.dataOff = b.st->out.size();
    for (long i = 0; i < len; i++) b.st->out.push_back((uint8_t)is.readBits(8));  // 0xFF past EOF (H10)
    b.dataLen = (uint64_t)len;
    return true;
}

bool initDynamicDecoder(Block& b, BitIn& is) {  // :892-1010
    b.numLitlenLens = (int)is.readBits(5) + MIN_LITLEN_LENS;
    b.numDistLens = (int)is.readBits(5) + MIN_DIST_LENS;
    b.numCodelenLens = (int)is.readBits(4) + MIN_CODELEN_LENS;
    if (is.eof) return false;  // the Java path fails a few reads later; no state survives either way
    std::vector<int> codelenLengths(MAX_CODELEN_LENS, 0);
    for (int i = 0; i < b.numCodelenLens; i++) codelenLengths[codelen_lengths_order[i]] = (int)is.readBits(3);
    if (is.eof) return false;
    b.dynamicHeaderSizeBits = 5 + 5 + 4 + b.numCodelenLens * 3L;
    b.codeLenDec = ofCodelens(codelenLengths);
    Decoder clDec(*b.codeLenDec);
    std::vector<int> codeLengths(MAX_LITLEN_LENS + MAX_DIST_LENS, 0);
    b.rlePairs = std::make_shared<std::vector<Pair>>();
    b.didCopyRLEPairs = true;
    int i = 0;
    const int combinedLens = b.numLitlenLens + b.numDistLens;
    while (i < combinedLens) {
        int cl;
        int sym = clDec.readSym(is, &cl);
        b.dynamicHeaderSizeBits += cl;
        int dist;
        uint8_t val;
        if (sym >= 0 && sym <= 15) {
            codeLengths[i++] = sym;
            dist = 0; val = (uint8_t)sym;
        } else if (sym == 16) {
            if (i < 1) return false;
            int n = (int)is.readBits(2) + 3;
            b.dynamicHeaderSizeBits += 2;
            dist = n;
            if (i + n > combinedLens) return false;
            val = (uint8_t)codeLengths[i - 1];
            for (int k = 0; k < n; k++) codeLengths[i++] = val;
        } else if (sym == 17) {
            int n = (int)is.readBits(3) + 3;
            b.dynamicHeaderSizeBits += 3;
            dist = n;
            if (i + n > combinedLens) return false;
            val = 0; i += n;
        } else if (sym == 18) {
            int n = (int)is.readBits(7) + 11;
            b.dynamicHeaderSizeBits += 7;
            dist = n;
            if (i + n > combinedLens) return false;
            val = 0; i += n;
        } else {
            return false;
        }
        b.rlePairs->push_back(Pair{dist, sym, val});
    }
    std::vector<int> ll(codeLengths.begin(), codeLengths.begin() + b.numLitlenLens);
    std::vector<int> dl(codeLengths.begin() + b.numLitlenLens, codeLengths.begin() + b.numLitlenLens + b.numDistLens);
    b.litlenDec = ofCodelens(ll);
    b.distDec = ofCodelens(dl);
    b.sizeBits += b.dynamicHeaderSizeBits;
    return true;
}

bool decodeStream(Block& b, BitIn& is) {  // :778-890
    std::vector<uint8_t>& out = b.st->out;
    b.dataOff = out.size();
    b.litlens = std::make_shared<std::vector<LitLen>>();
    b.didCopyLitLens = true;
    Decoder litDec(*b.litlenDec), distDec(*b.distDec);
    while (true) {
        int cl;
        int litlen = litDec.readSym(is, &cl);
        if (litlen < 0 || litlen > LITLEN_MAX) return false;
        if (litlen <= 0xff) {
            b.litlens->push_back(LitLen{0, litlen, false, out.size()});
            b.sizeBits += cl; b.litlenSizeBits += cl;
            out.push_back((uint8_t)litlen);
            continue;
        }
        if (litlen == LITLEN_EOB) {
            b.sizeBits += cl; b.litlenSizeBits += cl;
            b.litlens->push_back(LitLen{0, LITLEN_EOB, false, out.size()});
            b.dataLen = out.size() - b.dataOff;
            return true;
        }
        int totalSize = cl;
        long len = len_base[litlen - LITLEN_TBL_OFFSET];
        long ebits = len_ebits[litlen - LITLEN_TBL_OFFSET];
        if (ebits != 0) { totalSize += (int)ebits; len += is.readBits((int)ebits); }
        bool edgecase = (len == MAX_LEN) && (litlen == 284);
        int dcl;
        int distsym = distDec.readSym(is, &dcl);
        totalSize += dcl;
        if (distsym < 0 || distsym > DISTSYM_MAX) return false;
        long dist = dist_base[distsym];
        ebits = dist_ebits[distsym];
        if (ebits != 0) { totalSize += (int)ebits; dist += is.readBits((int)ebits); }
        if (is.eof) return false;
        b.sizeBits += totalSize; b.litlenSizeBits += totalSize;
        // readSlice (DeflateBlock.java:147-222): LZ77 copy, walking back through earlier blocks.  A
        // distance reaching before the start of the stream dereferences a null prevBlock in the
        // reference (crash); the oracle reports a parse failure instead.
        uint64_t p = out.size();
        if ((uint64_t)dist > p) return false;
        b.litlens->push_back(LitLen{(int32_t)dist, (int32_t)len, edgecase, p});
        for (long k = 0; k < len; k++) out.push_back(out[p - dist + k]);
    }
}

bool parseStream(Stream& s, BitIn& bis) {  // DeflateStream.java:72-126
    bool bfinal;
    Block* prev = nullptr;
    BlockP prevP;
    bool first = true;
    do {
        long bits = bis.readBits(3);
        bfinal = (bits & 1) != 0;
        bits = (long)((unsigned long)bits >> 1);
        BlockP nb;
        switch ((int)bits) {
            case 0: nb = std::make_shared<Block>(&s, STORED, prev); break;
            case 1: nb = std::make_shared<Block>(&s, FIXED, prev); break;
            case 2: nb = std::make_shared<Block>(&s, DYNAMIC, prev); break;
            default: return false;
        }
        bool ok;
        if (nb->type == STORED) ok = parseStored(*nb, bis);
        else {
            nb->sizeBits = 0;
            if (nb->type == DYNAMIC) ok = initDynamicDecoder(*nb, bis);
            else { nb->litlenDec = fixedLit(); nb->distDec = fixedDist(); ok = true; }
            ok = ok && decodeStream(*nb, bis);
        }
        if (!ok) return false;
        if (prev) prev->next = nb;
        prev = nb.get();
        prevP = nb;
        if (first) { setFirstBlock(s, nb); first = false; }
    } while (!bfinal);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Writing (DeflateStream.java:128-145, DeflateBlockUncompressed.java:39-56,
// DeflateBlockHuffman.java:1033-1156)
// ---------------------------------------------------------------------------------------------
bool writeBlock(const Block& b, BitOut& os, bool finalBlock) {
    const uint8_t* out = outOf(b);
    if (b.type == STORED) {
        os.writeNBits(finalBlock ? 1 : 0, 3);
        int len = (int)b.dataLen;
        uint8_t h[4];
        h[0] = (uint8_t)len; h[1] = (uint8_t)(len >> 8); h[2] = (uint8_t)~h[0]; h[3] = (uint8_t)~h[1];
        os.flushToByteAligned();
        os.writeBytes(h, 4);
        os.writeBytes(out + b.dataOff, b.dataLen);
        return true;
    }
    os.writeNBits(((int)b.type << 1) | (finalBlock ? 1 : 0), 3);  // writeProlog :1148-1151
    if (b.type == DYNAMIC) {                                       // writeHuffCode :1033-1103
        os.writeNBits((uint64_t)(long)(b.numLitlenLens - MIN_LITLEN_LENS), 5);
        os.writeNBits((uint64_t)(long)(b.numDistLens - MIN_DIST_LENS), 5);
        os.writeNBits((uint64_t)(long)(b.numCodelenLens - MIN_CODELEN_LENS), 4);
        for (int i = 0; i < b.numCodelenLens; i++)
            os.writeNBits(b.codeLenDec->codeLen[codelen_lengths_order[i]], 3);
        for (const Pair& p : *b.rlePairs) {
            os.writeNBits(b.codeLenDec->getSym(p.sym), b.codeLenDec->getSymLen(p.sym));
            if (p.dist != 0) {
                int off, sz;
                switch (p.sym) {
                    case 16: off = 3; sz = 2; break;
                    case 17: off = 3; sz = 3; break;
                    case 18: off = 11; sz = 7; break;
                    default: return false;
                }
                os.writeNBits(p.dist - off, sz);
            }
        }
    }
    for (const LitLen& l : *b.litlens) {  // writeDefBlock :1140-1146
        if (l.dist == 0) {
            os.writeNBits(b.litlenDec->getSym(l.litlen), b.litlenDec->getSymLen(l.litlen));
        } else {  // writeBackref :1110-1130
            int litlen = len2litlen(l.litlen, l.edgecase);
            uint64_t bits = b.litlenDec->getSym(litlen);
            int nbits = b.litlenDec->getSymLen(litlen);
            uint64_t ebits = l.litlen - len_base[litlen - LITLEN_TBL_OFFSET];
            bits |= ebits << nbits;
            nbits += len_ebits[litlen - LITLEN_TBL_OFFSET];
            int d = distance2dist(l.dist);
            bits |= (uint64_t)b.distDec->getSym(d) << nbits;
            nbits += b.distDec->getSymLen(d);
            ebits = l.dist - dist_base[d];
            bits |= ebits << nbits;
            nbits += dist_ebits[d];
            os.writeNBits(bits, nbits);
        }
    }
    return true;
}

void writeStream(const Stream& s, BitOut& os) {
    for (Block* b = s.first.get(); b; b = b->next.get()) writeBlock(*b, os, b->next == nullptr);
    os.flushToByteAligned();
}

// ---------------------------------------------------------------------------------------------
// The candidate enumerator (DeflateStream.java:184-490)
// ---------------------------------------------------------------------------------------------
using Callback = std::function<void(const BlockP&)>;

BlockP optimiseBlockDynBlock(const BlockP& block, bool ohh, bool use8, bool use7, bool alt8, bool noRep,
                             bool noZRep, bool noZRep2, bool prune, bool noRepZeros) {  // :184-198
    if (block->type != DYNAMIC) return nullptr;
    BlockP o = copyBlock(*block);
    rewriteHeader(*o, ohh, use8, use7, alt8, noRep, noZRep, noZRep2, noRepZeros);
    if (prune) recodeHeaderToLessRLEMatches(*o);
    optimiseHeader(*o);
    return o;
}
BlockP recodedHuffman(const BlockP& block, bool prune) {  // :200-210
    BlockP r = copyBlock(*block);
    if (prune) recodeHuffmanLessMatches(*r); else recodeHuffman(*r);
    return r;
}
BlockP recodedHuffmanFull(BlockP block, long align) {  // :212-229
    long prevSize = getSizeBits(*block, align);
    while (true) {
        BlockP check = recodedHuffman(block, true);
        long thisSize = getSizeBits(*check, align);
        if (thisSize >= prevSize) break;
        block = check;
        prevSize = thisSize;
    }
    return block;
}
BlockP leastExpPruned(const BlockP& b) { BlockP r = copyBlock(*b); removeDistLitLeastExpensive(*r, 0); return r; }   // :231-235
BlockP leastSeenPruned(const BlockP& b) { BlockP r = copyBlock(*b); removeDistLitLeastExpensive(*r, 1); return r; }  // :237-241
BlockP optimiseBlockCopyHelper(const BlockP& b) { BlockP o = copyBlock(*b); optimiseBlockInPlace(*o); return o; }     // :248-252
BlockP optimiseBlockHelper(const BlockP& b) { optimiseBlockInPlace(*b); return b; }                                    // :254-257

// addOptimisedRecoded (:265-317).  A candidate from a non-dynamic base would be null in the reference
// (NullPointerException in the callback); every base here is dynamic by construction.
void addOptimisedRecoded(const Callback& callback, const BlockP& toOptimise, long position) {
    std::vector<BlockP> blocks;
    blocks.push_back(optimiseBlockCopyHelper(toOptimise));
    blocks.push_back(optimiseBlockHelper(recodedHuffman(toOptimise, false)));
    BlockP pruned = recodedHuffman(toOptimise, true);
    blocks.push_back(optimiseBlockCopyHelper(pruned));
    BlockP prunedFull = recodedHuffmanFull(pruned, position);
    if (prunedFull != pruned) blocks.push_back(optimiseBlockHelper(prunedFull));
    static const bool TRUE_FALSE[2] = {true, false}, FALSE_TRUE[2] = {false, true};
    for (const BlockP& block : blocks) {
        for (bool noRepZeros : FALSE_TRUE)
            for (bool prune : FALSE_TRUE)
                for (int a = 0; a < (noRepZeros ? 1 : 2); a++) {
                    bool noRep = noRepZeros ? false : FALSE_TRUE[a];
                    for (int z = 0; z < (noRepZeros ? 1 : 2); z++) {
                        bool noZRep = noRepZeros ? true : FALSE_TRUE[z];
                        for (bool noZRep2 : FALSE_TRUE)
                            for (bool ohh : TRUE_FALSE) {
                                if (ohh) {
                                    if (noRep) continue;
                                    const bool alt8 = false;  // ALT_8_ARR = {DEFAULT_8}  (:243-246,263)
                                    for (bool use8 : TRUE_FALSE)
                                        for (bool use7 : TRUE_FALSE) {
                                            if (!use8 && (alt8 || !use7)) continue;
                                            callback(optimiseBlockDynBlock(block, true, use8, use7, alt8, false, noZRep,
                                                                           noZRep2, prune, noRepZeros));
                                        }
                                } else {
                                    callback(optimiseBlockDynBlock(block, false, false, false, false, noRep, noZRep,
                                                                   noZRep2, prune, noRepZeros));
                                }
                            }
                    }
                }
    }
}

BlockP optimiseBlockNormal(const BlockP& block) {  // :319-327
    BlockP o = copyBlock(*block);
    if (optimiseBlockInPlace(*o) > 0) return o;
    return nullptr;
}
BlockP toFixedHuffman(const BlockP& block) {  // :329-337
    if (block->type != DYNAMIC) return nullptr;
    BlockP f = copyBlock(*block);
    recodeToFixedHuffman(*f);
    return f;
}

BlockP optimiseBlock(const BlockP& toOptimise, long position) {  // :343-490
    g_stats.optimise_block_calls++;
    BlockP smallest = toOptimise;
    long smallestSize = getSizeBits(*toOptimise, position);
    tracePut(-1, smallestSize);
    g_trace_idx = 0;
    Callback callback = [&](const BlockP& cand) {
        g_stats.candidates++;
        if (!cand) { fprintf(stderr, "oracle: null candidate (reference would throw)\n"); abort(); }
        long newSize = getSizeBits(*cand, position);
        tracePut(g_trace_idx++, newSize);
        if (newSize < smallestSize) { smallest = cand; smallestSize = newSize; }
    };
    BlockP optimised = optimiseBlockNormal(toOptimise);
    if (optimised) callback(optimised);
    if (toOptimise->type != STORED) {
        BlockP stored = asUncompressed(*toOptimise);
        if (stored->dataLen <= 65535) callback(stored);
    }
    BlockP toOptimiseHuffman, optimisedHuffman;
    const bool isOrigDyn = toOptimise->type == DYNAMIC, isOrigFixed = toOptimise->type == FIXED;
    if (isOrigDyn) {
        toOptimiseHuffman = toOptimise;
        optimisedHuffman = optimised;
    } else if (isOrigFixed) {
        toOptimiseHuffman = copyBlock(*toOptimise);
        recodeHuffman(*toOptimiseHuffman);
        optimisedHuffman = optimiseBlockNormal(toOptimiseHuffman);
    }
    if (toOptimiseHuffman) {
        auto runOptimisations = [&](const BlockP& k) {  // :400-442
            BlockP post = copyBlock(*k);
            recodeHeader(*post);
            callback(post);
            BlockP postOpt = optimiseBlockNormal(post);
            if (postOpt) callback(postOpt);
            addOptimisedRecoded(callback, post, position);
            BlockP prune = copyBlock(*k);
            recodeHeaderToLessRLEMatches(*prune);
            callback(prune);
            BlockP pruneOpt = optimiseBlockNormal(prune);
            if (pruneOpt) callback(pruneOpt);
            addOptimisedRecoded(callback, prune, position);
            addOptimisedRecoded(callback, leastExpPruned(k), position);
            addOptimisedRecoded(callback, leastSeenPruned(k), position);
        };
        auto runMulti = [&](const BlockP& e) {  // :443-463
            callback(e);
            runOptimisations(e);
            BlockP huffRec = recodedHuffman(e, false);
            callback(huffRec);
            runOptimisations(huffRec);
            BlockP pruned = recodedHuffman(e, true);
            callback(pruned);
            runOptimisations(pruned);
            BlockP prunedFull = recodedHuffmanFull(pruned, position);
            if (prunedFull != pruned) {
                callback(prunedFull);
                runOptimisations(prunedFull);
            }
        };
        runMulti(toOptimiseHuffman);
        if (optimisedHuffman) runMulti(optimisedHuffman);
        if (!isOrigFixed) {
            BlockP fixed = toFixedHuffman(toOptimiseHuffman);
            if (fixed) { optimiseBlockInPlace(*fixed); callback(fixed); }
        }
        runMulti(leastExpPruned(toOptimiseHuffman));
        runMulti(leastSeenPruned(toOptimiseHuffman));
    }
    return smallest;
}

// DeflateStream.mergeBlocks (:568-650)
long mergeBlocksPhase(Stream& s) {
    long pos = 0, saved = 0;
    bool first = true;
    BlockP currentBlock = s.first;
    while (currentBlock) {
        bool finishPass = true, didRemove = false;
        BlockP nextBlock = currentBlock->next;
        if (first && !nextBlock) {
            pos += getSizeBits(*currentBlock, pos + 3) + 3;
        } else if (currentBlock->dataLen > 0) {
            pos += 3;
            if (nextBlock && canMerge(*currentBlock, nextBlock.get())) {
                BlockP merged = optimiseBlock(mergeBlocksPair(currentBlock, nextBlock), pos);
                long currentSizeNoMerge = getSizeBits(*currentBlock, pos);
                long nextSizeNoMerge = getSizeBits(*nextBlock, pos + currentSizeNoMerge + 3);
                long currentSaved = (currentSizeNoMerge + 3 + nextSizeNoMerge) - getSizeBits(*merged, pos);
                if (merged != currentBlock && currentSaved > 0) {
                    finishPass = false;
                    saved += currentSaved;
                    replaceBlock(*currentBlock, merged);
                    discardBlock(*currentBlock);
                    BlockP next = nextBlock->next;
                    merged->next = next;
                    if (next) next->prev = merged.get();
                    currentBlock = merged;
                    if (first) setFirstBlock(s, merged);
                }
            }
            pos += getSizeBits(*currentBlock, pos);
        } else {
            long currentSaved = getSizeBits(*currentBlock, pos + 3) + 3;
            saved += currentSaved;
            if (first) setFirstBlock(s, currentBlock->next);
            removeBlock(*currentBlock);
            didRemove = true;
        }
        if (finishPass) {
            currentBlock = currentBlock->next;  // null after remove() (discard clears it): H6
            if (first && !didRemove) first = false;
        }
    }
    return saved;
}

// DeflateStream.optimise(boolean) (:496-566)
long optimiseStream(Stream& s, bool mergeBlocks) {
    long pos = 0, saved = 0;
    bool first = true;
    BlockP currentBlock = s.first;
    while (currentBlock) {
        bool finishPass = true, didRemove = false;
        if (currentBlock->dataLen > 0 || (first && !currentBlock->next)) {
            pos += 3;
            BlockP optimisedBlock = optimiseBlock(currentBlock, pos);
            long currentSaved = getSizeBits(*currentBlock, pos) - getSizeBits(*optimisedBlock, pos);
            if (optimisedBlock != currentBlock && currentSaved > 0) {
                finishPass = false;
                saved += currentSaved;
                replaceBlock(*currentBlock, optimisedBlock);
                discardBlock(*currentBlock);
                currentBlock = optimisedBlock;
                if (first) setFirstBlock(s, optimisedBlock);
            }
            pos += getSizeBits(*currentBlock, pos);
        } else {
            long currentSaved = getSizeBits(*currentBlock, pos + 3) + 3;
            saved += currentSaved;
            if (first) setFirstBlock(s, currentBlock->next);
            removeBlock(*currentBlock);
            didRemove = true;
        }
        if (finishPass) {
            currentBlock = currentBlock->next;
            if (first && !didRemove) first = false;
        }
    }
    return mergeBlocks ? saved + mergeBlocksPhase(s) : saved;
}

long streamSizeBits(const Stream& s) {  // DeflateStream.java:171-182
    long size = 0;
    for (Block* b = s.first.get(); b; b = b->next.get()) { size += 3; size += getSizeBits(*b, size); }
    return size;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------
struct ora_stream { Stream s; };

extern "C" {

ora_stream* ora_parse(const uint8_t* data, size_t len, size_t* consumed) {
    auto* h = new ora_stream();
    BitIn bis{data, len};
    bool ok = parseStream(h->s, bis);
    if (consumed) *consumed = bis.consumed();
    if (!ok) { ora_free(h); return nullptr; }
    return h;
}
void ora_free(ora_stream* h) {
    if (!h) return;
    // break the next-chain iteratively (deep recursion in shared_ptr destructors otherwise)
    BlockP b = h->s.first;
    h->s.first = nullptr;
    while (b) { BlockP n = b->next; b->next = nullptr; b = n; }
    delete h;
}
int64_t ora_optimise(ora_stream* h, int merge) { return optimiseStream(h->s, merge != 0); }
int64_t ora_size_bits(const ora_stream* h) { return streamSizeBits(h->s); }
size_t ora_uncompressed_len(const ora_stream* h) {
    size_t n = 0;
    for (Block* b = h->s.first.get(); b; b = b->next.get()) n += b->dataLen;
    return n;
}
void ora_uncompressed(const ora_stream* h, uint8_t* dst) {
    for (Block* b = h->s.first.get(); b; b = b->next.get()) {
        memcpy(dst, h->s.out.data() + b->dataOff, b->dataLen);
        dst += b->dataLen;
    }
}
size_t ora_write(const ora_stream* h, uint8_t* dst, size_t cap) {
    BitOut os;
    writeStream(h->s, os);
    if (dst && cap >= os.buf.size()) memcpy(dst, os.buf.data(), os.buf.size());
    return os.buf.size();
}
uint32_t ora_block_count(const ora_stream* h) {
    uint32_t n = 0;
    for (Block* b = h->s.first.get(); b; b = b->next.get()) n++;
    return n;
}
static Block* nthBlock(const ora_stream* h, uint32_t idx, long* posOut) {
    long pos = 0;
    uint32_t i = 0;
    for (Block* b = h->s.first.get(); b; b = b->next.get(), i++) {
        if (i == idx) { if (posOut) *posOut = pos; return b; }
        pos += 3;
        pos += getSizeBits(*b, pos);
    }
    return nullptr;
}
int ora_block_info_get(const ora_stream* h, uint32_t idx, ora_block_info* o) {
    long pos;
    Block* b = nthBlock(h, idx, &pos);
    if (!b) return 1;
    memset(o, 0, sizeof *o);
    o->type = b->type;
    o->position = pos;
    o->size_bits = getSizeBits(*b, pos + 3);
    o->uncompressed_len = b->dataLen;
    if (b->type != STORED) {
        o->n_symbols = (uint32_t)b->litlens->size();
        o->litlen_size_bits = b->litlenSizeBits;
        if (b->type == DYNAMIC) {
            o->n_rle_pairs = (uint32_t)b->rlePairs->size();
            o->num_litlen_lens = b->numLitlenLens; o->num_dist_lens = b->numDistLens;
            o->num_codelen_lens = b->numCodelenLens;
            o->header_size_bits = b->dynamicHeaderSizeBits;
        }
    }
    return 0;
}
uint32_t ora_block_symbols(const ora_stream* h, uint32_t idx, int32_t* dst, uint32_t cap) {
    Block* b = nthBlock(h, idx, nullptr);
    if (!b || b->type == STORED) return 0;
    uint32_t n = (uint32_t)b->litlens->size();
    for (uint32_t i = 0; i < n && i < cap; i++) {
        const LitLen& l = (*b->litlens)[i];
        dst[3 * i] = l.dist; dst[3 * i + 1] = l.litlen; dst[3 * i + 2] = l.edgecase;
    }
    return n;
}
uint32_t ora_block_rle_pairs(const ora_stream* h, uint32_t idx, int32_t* dst, uint32_t cap) {
    Block* b = nthBlock(h, idx, nullptr);
    if (!b || b->type != DYNAMIC) return 0;
    uint32_t n = (uint32_t)b->rlePairs->size();
    for (uint32_t i = 0; i < n && i < cap; i++) {
        dst[2 * i] = (*b->rlePairs)[i].dist; dst[2 * i + 1] = (*b->rlePairs)[i].sym;
    }
    return n;
}
uint32_t ora_block_codelens(const ora_stream* h, uint32_t idx, int which, int32_t* dst, uint32_t cap) {
    Block* b = nthBlock(h, idx, nullptr);
    if (!b || b->type == STORED) return 0;
    const Table* t = which == 0 ? b->litlenDec.get() : which == 1 ? b->distDec.get() : b->codeLenDec.get();
    if (!t) return 0;
    uint32_t n = (uint32_t)t->codeLen.size();
    for (uint32_t i = 0; i < n && i < cap; i++) dst[i] = t->codeLen[i];
    return n;
}

int ora_optimise_stream(const uint8_t* data, size_t len, int merge, uint8_t** out, size_t* out_len,
                        int64_t* saved_bits, size_t* consumed) {
    ora_stream* h = ora_parse(data, len, consumed);
    if (!h) return 1;
    int64_t saved = ora_optimise(h, merge);
    if (saved_bits) *saved_bits = saved;
    size_t n = ora_write(h, nullptr, 0);
    *out = (uint8_t*)malloc(n ? n : 1);
    ora_write(h, *out, n);
    *out_len = n;
    ora_free(h);
    return 0;
}
void ora_free_buf(uint8_t* p) { free(p); }

void ora_trace_begin(int64_t* buf, size_t cap_pairs) { g_trace = buf; g_trace_cap = cap_pairs; g_trace_n = 0; }
size_t ora_trace_end(void) { g_trace = nullptr; return g_trace_n; }

void ora_huffman_tree(const int32_t* freq, int n, int limit, int32_t* code_out, int32_t* len_out) {
    std::vector<int> f(freq, freq + n);
    TableP t = buildTree(f, limit);
    for (int i = 0; i < n; i++) { code_out[i] = t->code[i]; len_out[i] = t->codeLen[i]; }
}
int ora_pack_code_lengths(const int32_t* lit, int nlit, const int32_t* dist, int ndist, int flags, int32_t* dst,
                          int cap) {
    std::vector<int> combined(lit, lit + nlit);
    combined.insert(combined.end(), dist, dist + ndist);
    std::vector<int> lengths;
    pack(lengths, combined, flags & 1, flags & 2, flags & 4, flags & 8, flags & 16, flags & 32, flags & 64,
         flags & 128);
    for (size_t i = 0; i < lengths.size() && (int)i < cap; i++) dst[i] = lengths[i];
    return (int)lengths.size();
}
// One header-strategy trial (DeflateStream.optimiseBlockDynBlock, :184-198) on a bare dynamic block that has
// only its two code-length tables: rewriteHeader(flags) [+ recodeHeaderToLessRLEMatches] + optimiseHeader.
// ops: bit8 of flags = prune.  After the trial, `post_ops` (0 none, 1 recodeHeader, 2 recodeHeaderToLessRLEMatches,
// 3 optimiseHeader) is applied once more so those mutators can be pinned on a realistic header too.
int64_t ora_header_trial(const int32_t* lit, int nlit, const int32_t* dist, int ndist, int flags, int post_op,
                         int32_t* pairs_out, int32_t* np_out, int32_t* cl_out, int32_t* ncl_out) {
    Stream st;
    auto b = std::make_shared<Block>(&st, DYNAMIC, nullptr);
    b->litlenDec = ofCodelens(std::vector<int>(lit, lit + nlit));
    b->distDec = ofCodelens(std::vector<int>(dist, dist + ndist));
    b->litlens = std::make_shared<std::vector<LitLen>>();
    b->rlePairs = std::make_shared<std::vector<Pair>>();
    BlockP o = optimiseBlockDynBlock(b, flags & 1, flags & 2, flags & 4, flags & 8, flags & 16, flags & 32, flags & 64,
                                     flags & 256, flags & 128);
    if (post_op == 1) recodeHeader(*o);
    else if (post_op == 2) recodeHeaderToLessRLEMatches(*o);
    else if (post_op == 3) optimiseHeader(*o);
    int n = 0;
    for (const Pair& p : *o->rlePairs) { pairs_out[2 * n] = p.dist; pairs_out[2 * n + 1] = p.sym; n++; }
    *np_out = n;
    for (int i = 0; i < 19; i++) cl_out[i] = o->codeLenDec->codeLen[i];
    *ncl_out = o->numCodelenLens;
    return o->dynamicHeaderSizeBits;
}
void ora_get_stats(ora_stats* o) { *o = g_stats; }
void ora_reset_stats(void) { memset(&g_stats, 0, sizeof g_stats); }

}  // extern "C"
