"""CPU pins of the cost model's single-thread building blocks: the SAME source the kernels run
(deft4j_b200/csrc/huff.cuh, compiled as host code by hosttest.cu) against the oracle.

 * huff_tree        vs HuffmanTree (huffman/HuffmanTree.java:36-128,164-192) incl. PriorityQueue tie-breaks
 * hdr_trial & co.  vs optimiseBlockDynBlock (DeflateStream.java:184-198) = rewriteHeader + recodeHeaderToLessRLEMatches
                       + optimiseHeader (DeflateBlockHuffman.java:471-635), all 56 strategy combinations
"""
import random

import pytest

import hosttest_lib as H


def _rand_freq(rnd, n, style):
    if style == 0:      # geometric-ish text-like
        return [int(rnd.expovariate(1.0 / rnd.choice([1, 5, 50, 500]))) for _ in range(n)]
    if style == 1:      # sparse
        return [rnd.choice([0, 0, 0, 1, 2, 1000]) for _ in range(n)]
    if style == 2:      # many ties
        return [rnd.choice([0, 1, 1, 2, 2, 3]) for _ in range(n)]
    if style == 3:      # fibonacci-like -> deep trees, exercises the depth limiter
        f, a, b = [], 1, 1
        for _ in range(n):
            f.append(a); a, b = b, a + b
            if a > 1 << 24: a, b = 1, 1
        rnd.shuffle(f)
        return f
    return [rnd.randint(0, 3) * rnd.randint(0, 40) for _ in range(n)]


@pytest.mark.parametrize("n,limit", [(19, 7), (30, 15), (286, 15), (257, 15), (5, 7), (2, 15), (1, 15)])
def test_huff_tree_matches_oracle(oracle, n, limit):
    rnd = random.Random(n * 100 + limit)
    for it in range(300 if n < 100 else 120):
        freq = _rand_freq(rnd, n, it % 5)
        if it % 17 == 0:
            freq = [0] * n
        if it % 19 == 0:
            freq = [0] * n; freq[rnd.randrange(n)] = rnd.randint(1, 9)
        rc, lens = H.huff_tree(freq, limit)
        _, olens = oracle.huffman_tree(freq, limit)
        assert rc == 0
        assert lens == olens, (freq, lens, olens)


def _rand_tables(rnd, oracle):
    nl = rnd.choice([257, 258, 260, 270, 286, 286, 286])
    nd = rnd.choice([1, 2, 5, 20, 30, 30])
    style = rnd.randrange(5)
    _, L = oracle.huffman_tree(_rand_freq(rnd, nl, style), 15)
    _, D = oracle.huffman_tree(_rand_freq(rnd, nd, rnd.randrange(5)), 15)
    if rnd.random() < 0.2:
        D = [0] * nd
    return L, D


def test_header_trials_match_oracle(oracle):
    rnd = random.Random(7)
    flags = H.trial_flags()
    assert len(set(flags)) == 56
    for it in range(60):
        L, D = _rand_tables(rnd, oracle)
        for f in flags:
            got = H.header_trial(L, D, f)
            want = oracle.header_trial(L, D, f)
            assert got == want, (it, hex(f), L, D, got[0], want[0])


@pytest.mark.parametrize("post_op", [1, 2, 3])
def test_header_mutators_match_oracle(oracle, post_op):
    rnd = random.Random(11 + post_op)
    flags = H.trial_flags()
    for it in range(40):
        L, D = _rand_tables(rnd, oracle)
        for f in rnd.sample(flags, 12):
            assert H.header_trial(L, D, f, post_op) == oracle.header_trial(L, D, f, post_op), (it, hex(f), post_op)


def test_size_only_trials_match_full_trials():
    """trial_sizes (run list, no pair storage, both prune values at once — what the engine's trials() evaluates)
    against hdr_trial (the materialising implementation pinned on the oracle above) for every strategy."""
    rnd = random.Random(23)
    flags = H.trial_flags()
    rewrite = sorted({f & 0xFF for f in flags})
    assert len(rewrite) == 28

    def tables(kind):
        if kind == 0:      # sparse alphabets: long zero runs (18/17 paths), straddling the litlen/dist boundary
            L = [0] * rnd.randint(257, 288)
            for _ in range(rnd.randint(1, 12)):
                L[rnd.randrange(len(L))] = rnd.randint(1, 15)
            L[256] = L[256] or rnd.randint(1, 15)
            D = [0] * rnd.randint(1, 32)
            for _ in range(rnd.randint(0, 3)):
                D[rnd.randrange(len(D))] = rnd.randint(1, 15)
        elif kind == 1:    # long equal non-zero runs (16 paths incl. the 8 / 7 special cases)
            L, v = [], rnd.randint(1, 15)
            while len(L) < 286:
                L += [v] * rnd.choice([1, 2, 3, 7, 8, 9, 13, 14, 15, 20, 21, 22, 70])
                v = rnd.randint(0, 15)
            L = L[:rnd.randint(257, 288)]
            D = [rnd.choice([5, 5, 5, 0, 6])] * rnd.randint(1, 32)
        else:              # noisy
            L = [rnd.choice([0, 0, 7, 8, 8, 9, 9, 10, 12]) for _ in range(rnd.randint(257, 288))]
            D = [rnd.choice([0, 4, 5, 5, 6]) for _ in range(rnd.randint(1, 32))]
        return L, D

    checked = 0
    for it in range(90):
        L, D = tables(it % 3)
        for f in rewrite:
            rc, a, b = H.trial_sizes(L, D, f)
            want_a = H.header_trial(L, D, f)[0]
            want_b = H.header_trial(L, D, f | 0x100)[0]
            if want_a < 0 or want_b < 0:
                continue
            assert rc == 0 and (a, b) == (want_a, want_b), (it, hex(f), L, D, (a, b), (want_a, want_b))
            checked += 1
    assert checked > 2000


def test_compact_header_tree_matches_the_general_one():
    """huff_tree_ws over the 305-byte TreeWsCLc (what trial threads keep in shared memory) against the general
    workspace pinned on the oracle: 19 symbols, limit 7, weights adding up to at most 322 — including shapes that
    need the depth limiter and shapes with fewer than two used symbols (dummy leaves)."""
    rnd = random.Random(5)
    n_limited = n_fast = 0
    for it in range(6000):
        k = rnd.choice([0, 1, 2, 3, 5, 8, 12, 19])
        freq = [0] * 19
        budget = rnd.randint(k, 320) if k else 0
        for s in rnd.sample(range(19), k):
            freq[s] = 1
            budget -= 1
        while budget > 0 and k:
            s = rnd.choice([i for i in range(19) if freq[i]])
            add = min(budget, rnd.choice([1, 1, 2, 3, 8, 21, 55, 144]))
            freq[s] += add
            budget -= add
        if it % 7 == 0 and k >= 9:   # Fibonacci-like weights force depth > 7
            fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89]
            used = [i for i in range(19) if freq[i]][:len(fib)]
            for i in range(19):
                freq[i] = 0
            for i, s in enumerate(used):
                freq[s] = fib[i]
        assert sum(freq) <= 322
        a = H.huff_tree(freq, 7)
        b = H.huff_tree_compact(freq, 7)
        assert a == b, (freq, a, b)
        c = H.huff_tree_tiny(freq, 7)
        assert c[:2] == a, (freq, a, c)
        n_limited += max(a[1]) == 7
        n_fast += not c[2]
    assert n_limited > 50 and n_fast > 50
