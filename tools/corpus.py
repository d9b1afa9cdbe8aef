"""ctypes wrapper of tools/libsdzcorpus.so: synthetic corpora compressed with
system zlib in the reference Deflater's wrappers.  Test / bench infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdzcorpus.so")
TEXT, BINARY, RANDOM, TINY, RUNS = range(5)
RAW, ZLIB, GZIP, GZIP_NAME, ZLIB_DICT = range(5)
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "corpus.c")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        L.sdzc_generate.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t]
        L.sdzc_generate.restype = None
        L.sdzc_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32,
                                    C.c_void_p, C.c_size_t]
        L.sdzc_compress.restype = C.c_size_t
        L.sdzc_make_batch.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                      C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def generate(kind, index, n):
    out = np.empty(max(n, 1), dtype=np.uint8)
    lib().sdzc_generate(kind, index, out.ctypes.data, n)
    return out[:n]


def compress(plain, level=6, container=ZLIB, dictionary=None, dictid=0):
    plain = np.ascontiguousarray(np.frombuffer(bytes(plain), dtype=np.uint8)) if not isinstance(plain, np.ndarray) else plain
    cap = int(plain.size * 1.01) + 4096
    out = np.empty(cap, dtype=np.uint8)
    d = None if dictionary is None else np.frombuffer(bytes(dictionary), dtype=np.uint8)
    n = lib().sdzc_compress(plain.ctypes.data if plain.size else None, plain.size, level, container,
                            None if d is None else d.ctypes.data, 0 if d is None else d.size, dictid & 0xFFFFFFFF,
                            out.ctypes.data, cap)
    if n == 0:
        raise RuntimeError("compress failed")
    return out[:n].tobytes()


def make_batch(kind, n, plain_len, level=6, container=ZLIB, first_index=0, keep_plain=False, threads=None):
    """Returns (comp_arena u8[n*stride], stride, comp_len u64[n], plain u8[n*plain_len] or None)."""
    threads = threads or min(os.cpu_count() or 1, 32)
    stride = ((int(plain_len * 1.01) + 4096 + 15) // 16) * 16
    comp = np.empty(n * stride, dtype=np.uint8)
    comp_len = np.zeros(n, dtype=np.uint64)
    plain = np.empty(n * plain_len, dtype=np.uint8) if keep_plain else None
    rc = lib().sdzc_make_batch(kind, first_index, n, plain_len, level, container, comp.ctypes.data, stride,
                               comp_len.ctypes.data, None if plain is None else plain.ctypes.data, threads)
    if rc:
        raise RuntimeError("corpus batch failed")
    return comp, stride, comp_len, plain
