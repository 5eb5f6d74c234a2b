"""ctypes binding of deft4j_b200/libdeft4cu_hosttest.so — huff.cuh compiled as host code (TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "deft4j_b200", "libdeft4cu_hosttest.so")
_SRC = os.path.join(_ROOT, "deft4j_b200", "csrc")
_lib = None


def build():
    srcs = [os.path.join(_SRC, f) for f in ("hosttest.cu", "huff.cuh", "common.cuh", "enum.cuh")]
    if (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-o", _SO, os.path.join(_SRC, "hosttest.cu")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.host_huff_tree.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
        L.host_header_trial.restype = C.c_longlong
        L.host_header_trial.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def huff_tree(freq, limit):
    n = len(freq)
    f = (C.c_uint32 * n)(*freq)
    lens = (C.c_uint8 * n)()
    rc = lib().host_huff_tree(f, n, limit, lens)
    return rc, list(lens)


def header_trial(lit, dist, flags, post_op=0):
    a = (C.c_uint8 * len(lit))(*lit)
    b = (C.c_uint8 * len(dist))(*dist)
    pairs = (C.c_int32 * 700)()
    np_, ncl = C.c_int32(0), C.c_int32(0)
    cl = (C.c_int32 * 19)()
    bits = lib().host_header_trial(a, len(lit), b, len(dist), flags, post_op, pairs, C.byref(np_), cl, C.byref(ncl))
    return bits, [(pairs[2 * i], pairs[2 * i + 1]) for i in range(np_.value)], list(cl), ncl.value


def trial_flags():
    return [lib().host_trial_flags(k) for k in range(56)]


def trial_sizes(lit, dist, flags):
    """(bits without prune, bits with prune) of one rewrite strategy, by the size-only evaluator."""
    a = (C.c_uint8 * len(lit))(*lit)
    b = (C.c_uint8 * len(dist))(*dist)
    x, y = C.c_int32(0), C.c_int32(0)
    L = lib()
    L.host_trial_sizes.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    rc = L.host_trial_sizes(a, len(lit), b, len(dist), flags, C.byref(x), C.byref(y))
    return rc, x.value, y.value


def huff_tree_compact(freq, limit):
    n = len(freq)
    f = (C.c_uint32 * n)(*freq)
    lens = (C.c_uint8 * n)()
    L = lib()
    L.host_huff_tree_compact.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    rc = L.host_huff_tree_compact(f, n, limit, lens)
    return rc, list(lens)


def huff_tree_tiny(freq, limit):
    """The fast header-code tree of the trial threads (16-bit keys, byte parents, no stored leaf map); returns
    (rc, lens, fell back to the full algorithm)."""
    n = len(freq)
    f = (C.c_uint32 * n)(*freq)
    lens = (C.c_uint8 * n)()
    fb = C.c_int(0)
    L = lib()
    L.host_huff_tree_tiny.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
    rc = L.host_huff_tree_tiny(f, n, limit, lens, C.byref(fb))
    return rc, list(lens), fb.value


# ---- the native file front-ends over the CPU oracle (no GPU needed) ------------------------------------------------
_FRONT_SO = os.path.join(_ROOT, "tests", "_build", "libfront_oracle.so")
_front = None


def front_oracle_lib():
    """deft4j_b200/csrc/png_front.cpp linked against tests/front_oracle_shim.cpp (deft4cu_optimise_batch done by the
    oracle): the chunk model of the shipped front-end, checked on the CPU."""
    global _front
    if _front is None:
        srcs = [os.path.join(_SRC, "png_front.cpp"), os.path.join(_SRC, "zip_front.cpp"), os.path.join(_SRC, "gz_front.cpp"), os.path.join(_ROOT, "tests", "front_oracle_shim.cpp"),
                os.path.join(_ROOT, "oracle", "deft_oracle.cpp")]
        deps = srcs + [os.path.join(_ROOT, "include", "deft4cu.h"), os.path.join(_ROOT, "oracle", "deft_oracle.h"),
                       os.path.join(_SRC, "front_util.h")]
        if (not os.path.exists(_FRONT_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_FRONT_SO) for s in deps):
            os.makedirs(os.path.dirname(_FRONT_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", _FRONT_SO] + srcs)
        _front = C.CDLL(_FRONT_SO)
        _front.deft4cu_crc32.restype = C.c_uint32
        _front.deft4cu_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_uint64]
    return _front
