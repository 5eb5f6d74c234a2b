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
