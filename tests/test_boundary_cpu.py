"""CPU tests of the drop-in boundary: libsdzcuda.so loads, exports every symbol that
include/sdzcuda.h declares, refuses to work without a B200 (no CPU fallback), and the
host-side mirror validates arguments like the reference."""
import os
import re

import pytest

import sdzlib
from sdzlib import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sdzcuda.h")).read()
    declared = set(re.findall(r"\b(sdz_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sdz_ctx"}
    L = N.load()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(N.EXPORTS) <= declared
    assert b"sm_100a" in L.sdz_version()


def test_record_layout_matches_header():
    # struct sdz_result is 72 bytes: 3*u64 + 5*i32 + 3*u32 + 9*u8 + 7 pad
    assert N.ctypes_sizeof_result() == 72 if hasattr(N, "ctypes_sizeof_result") else True
    import ctypes
    assert ctypes.sizeof(N.Result) == 72
    from oracle import oracle as O
    assert ctypes.sizeof(O.Result) == 72


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(N.NativeError):
        N.Context(0)
    with pytest.raises(N.NativeError):
        sdzlib.adler32(b"abc")
    with pytest.raises(N.NativeError):
        sdzlib.inflate(b"\x78\x01\x03\x00\x00\x00\x00\x01")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sd-zlib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".ts", ".c", ".cc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sdz_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_argument_errors_like_reference():
    with pytest.raises(TypeError, match="data must be an ArrayBuffer or buffer view"):
        sdzlib.inflate(12345)
    with pytest.raises(ValueError, match="data buffer is too small"):
        sdzlib.inflate(b"x")
    with pytest.raises(TypeError, match="options.raw must be undefined or true or false"):
        sdzlib.Inflater(raw=1)
    with pytest.raises(ValueError, match="options.dictionary cannot be set when options.raw is true"):
        sdzlib.Inflater(raw=True, dictionary=b"abc")
    with pytest.raises(TypeError, match="options.dictionary must be undefined or a buffer or a buffer view"):
        sdzlib.Inflater(dictionary=3.5)
    with pytest.raises(TypeError, match="source must be a BufferSource"):
        sdzlib.adler32(None)
    assert sdzlib.Inflater().append(b"") == []
    assert sdzlib.mergeBuffers([b"ab", b"", b"c"]) == b"abc"


def test_crc32_combine_is_host_arithmetic():
    """sdz_crc32_combine needs no device: crc(A || B) from the parts, against zlib, in the reference's signed domain."""
    import zlib
    import numpy as np
    from sdzlib import large as LG
    rng = np.random.default_rng(2)
    sgn = lambda v: v - (1 << 32) if v & 0x80000000 else v
    for la, lb in ((0, 0), (1, 0), (0, 1), (5, 7), (4096, 1), (70001, 123457), (3, 1 << 20)):
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert LG.crc32_combine(sgn(zlib.crc32(a)), sgn(zlib.crc32(b)), lb) == sgn(zlib.crc32(a + b)), (la, lb)
    parts = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (10, 0, 33000, 1)]
    assert LG.combine_crcs([(sgn(zlib.crc32(p)), len(p)) for p in parts]) == sgn(zlib.crc32(b"".join(parts)))
