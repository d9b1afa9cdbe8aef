"""sdzlib - host-side mirror of the @stardazed/zlib public API for the B200 engine.

Same names, argument meaning and error behaviour as the reference's ES module
(dist/sd-zlib.d.ts:11-148) for the inflate + checksum path:

    adler32(source, seed=1)          src/adler32.ts:17
    crc32(source, seed=0)            src/crc32.ts:17
    inflate(data, dictionary=None)   src/sd-inflate.ts:189
    Inflater(raw=, dictionary=)      src/sd-inflate.ts:54   (.append / .finish)
    mergeBuffers(buffers)            src/common.ts:116
    inflateBatch(buffers, ...)       new entry point (one sdz_inflate_batch call)
    deflate_wrap_batch(...)          the containers Deflater writes (src/sd-deflate.ts:98-165) + source checksums

Everything is computed by libsdzcuda.so on a B200; there is no CPU implementation here.
(The reference host language is TypeScript on Node; this image has no JS runtime, so the
mirror is Python.  The TypeScript facade + N-API shim a maintainer would ship are in
sd-zlib_b200/ts/ and INTEGRATION.md.)
"""
from .api import (InflateError, InflateHang, InflateResult, Inflater, deflate_wrap_batch, adler32, adler32_chain, checksum_batch, crc32, crc32_chain, inflate, inflateBatch,
                  inflateLarge, inflate_batch_raw, inflate_large_raw, mergeBuffers)
from ._native import Context, NativeError, default_context

__all__ = ["adler32", "crc32", "adler32_chain", "crc32_chain", "checksum_batch", "inflate", "Inflater", "InflateResult", "inflateBatch",
           "inflateLarge", "inflate_batch_raw", "inflate_large_raw", "mergeBuffers", "Context", "NativeError", "default_context",
           "InflateError", "InflateHang", "deflate_wrap_batch"]
