"""Generates tests/golden/c2_oracle_hashes.json: the CPU oracle's output (sha256, length, saved bits) on C2-style
streams too large to run the oracle inside the GPU test suite.  Deterministic inputs (workloads.py seeds); the
oracle is pinned on the reference's own golden pairs (tests/test_oracle_golden.py).

usage: python scripts/make_c2_golden.py            (about 5 minutes, one process per case)
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

# (name, target compressed bytes, seed, mergeBlocks)
CASES = [("c2_4MiB_nomerge", 4 << 20, 0xDEF7, False), ("c2_8MiB_nomerge_seed2", 8 << 20, 0xDEF8, False),
         ("c2_640KiB_merge", 640 << 10, 0xDEF7, True), ("c2_640KiB_nomerge", 640 << 10, 0xDEF7, False)]


def run(case):
    import workloads as W
    import oracle_lib
    name, target, seed, merge = case
    raw = W._c2_stream(target, seed)
    s = oracle_lib.OracleDeflateStream()
    assert s.parse(raw)
    saved = s.optimise(merge)
    out = s.asBytes()
    return name, {"target_bytes": target, "seed": seed, "merge": merge, "in_len": len(raw),
                  "in_sha256": hashlib.sha256(raw).hexdigest(), "saved_bits": saved, "out_len": len(out),
                  "out_sha256": hashlib.sha256(out).hexdigest(), "blocks_out": oracle_lib.lib().ora_block_count(s.h)}


if __name__ == "__main__":
    with mp.get_context("spawn").Pool(len(CASES)) as pool:
        res = dict(pool.map(run, CASES))
    path = os.path.join(ROOT, "tests", "golden", "c2_oracle_hashes.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))
