import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (input, golden output, mergeBlocks) — flags from the reference's runTestOpt.sh:3-11
GOLDEN_PAIRS = [
    ("deflate-store-2.txt.gz", "deflate-store-2-opt.txt.gz", True),
    ("lz-twice-twice.txt.gz", "lz-twice-twice-opt.txt.gz", True),
    ("text.png", "text-opt.png", True),
    ("asyoulik/asyoulik-zopfli.txt.gz", "asyoulik/asyoulik-zopfli-opt.txt.gz", True),
    ("asyoulik/asyoulik-gzip.txt.gz", "asyoulik/asyoulik-gzip-opt.txt.gz", True),
    ("apng/ball.png", "apng/ball-opt.png", True),
    ("284-edge-case/284.png", "284-edge-case/284-opt.png", True),
    ("nerd/nerd.png", "nerd/nerd-opt.png", False),
    ("nerd/nerd-extopt.png", "nerd/nerd-fullopt.png", False),
]

UNPAIRED_INPUTS = ["ban.txt.gz", "lz.txt.gz", "deflate-store.txt.gz", "deflate-fixed.txt.gz", "deflate-fixed.txt.zz",
                   "deflate-dynamic.txt.gz", "asyoulik/asyoulik-gzip-extopt.txt.gz",
                   "asyoulik/asyoulik-zopfli-extopt.txt.gz", "nerd/nerd-best.png"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")


def golden_path(rel):
    return os.path.join(GOLDEN, rel)


def read_golden(rel):
    with open(golden_path(rel), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib
