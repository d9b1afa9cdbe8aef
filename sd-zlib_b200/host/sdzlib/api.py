"""Reference-shaped API over the C ABI (see package docstring)."""
import ctypes as C
import datetime
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N

MODE_SNIFF, MODE_INFLATER, MODE_RAW = 0, 1, 2
OUTPUT_BUFSIZE = 16384                      # src/zstream.ts:11
_CHECK = ("unchecked", "match", "mismatch")
_MSG = [
    "", "invalid gzip id", "unknown compression method", "invalid window size", "incorrect header check",
    "need dictionary", "invalid block type", "invalid stored block lengths", "too many length or distance symbols",
    "invalid bit length repeat", "oversubscribed dynamic bit lengths tree", "incomplete dynamic bit lengths tree",
    "oversubscribed literal/length tree", "incomplete literal/length tree", "oversubscribed distance tree",
    "incomplete distance tree", "empty distance tree with lengths", "invalid distance code",
    "invalid literal/length code",
    # SDZ_PARITY_SPEC only (zlib 1.3's texts)
    "invalid distance too far back", "invalid code -- missing end-of-block", "invalid code lengths set",
    "invalid literal/lengths set", "invalid distances set", "unknown header flags set", "header crc mismatch",
]
_THROWN = [
    "", "inflate error: bad input", "Custom dictionary is not valid for this data",
    "Custom dictionary required for this data", "inflate error: ", "inflate error: bad input data",
    "reference implementation does not terminate on this input (trailing bytes after the stream end)",
    "data buffer is too small", "Unexpected EOF during decompression", "Data integrity check failed",
    "Data size check failed", "Decompression error",
]


def _as_u8(source, what):
    """u8ArrayFromBufferSource (src/common.ts:102-114): bytes-like or any buffer view."""
    if isinstance(source, np.ndarray):
        return np.ascontiguousarray(source).view(np.uint8).reshape(-1)
    try:
        return np.frombuffer(memoryview(source).cast("B"), dtype=np.uint8)
    except TypeError:
        raise TypeError(what)


def _i32(v):
    v = int(v) & 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


# ------------------------------------------------------------------ checksums

def _checksum(fn_name, source, seed, ctx):
    view = _as_u8(source, "source must be a BufferSource")
    ctx = ctx or N.default_context()
    out = C.c_int32()
    fn = getattr(ctx.lib, fn_name)
    total = int(view.size)
    if total < (1 << 32):
        ctx.check(fn(ctx.h, view.ctypes.data if total else None, total, _i32(seed), 0, C.byref(out)))
        return out.value
    raise ValueError("buffers of 4 GiB or more are undefined in the reference (SURVEY Q13)")


def adler32(source, seed=1, ctx=None):
    """adler32(source, seed = 1) -> signed int32, bit-exact with src/adler32.ts (incl. Q1)."""
    return _checksum("sdz_adler32", source, seed, ctx)


def crc32(source, seed=0, ctx=None):
    """crc32(source, seed = 0) -> signed int32, bit-exact with src/crc32.ts."""
    return _checksum("sdz_crc32", source, seed, ctx)


def _chain(fn_name, source, seg_lens, seed, ctx, device_ptr=None):
    ctx = ctx or N.default_context()
    lens = np.ascontiguousarray(seg_lens, dtype=np.uint64)
    vals = np.zeros(len(lens), dtype=np.int32)
    last = C.c_int32()
    if device_ptr is None:
        view = _as_u8(source, "source must be a BufferSource")
        ptr, on_dev = view.ctypes.data, 0
    else:
        ptr, on_dev = device_ptr, 1
    ctx.check(getattr(ctx.lib, fn_name)(ctx.h, ptr, lens.ctypes.data, len(lens), _i32(seed), on_dev,
                                        vals.ctypes.data, C.byref(last)))
    return vals


def adler32_chain(source, seg_lens, seed=1, ctx=None, device_ptr=None):
    """values[i] of `s = adler32(segment_i, s)` over consecutive segments, one device pass."""
    return _chain("sdz_adler32_chain", source, seg_lens, seed, ctx, device_ptr)


def crc32_chain(source, seg_lens, seed=0, ctx=None, device_ptr=None):
    return _chain("sdz_crc32_chain", source, seg_lens, seed, ctx, device_ptr)


def checksum_batch(buffers, kinds, seeds=None, ctx=None):
    """[adler32(b, seed) if kind == 'adler32' else crc32(b, seed) for b in buffers] in one launch
    (what Deflater needs for its inputs, src/sd-deflate.ts:185-190)."""
    ctx = ctx or N.default_context()
    views = [_as_u8(b, "source must be a BufferSource") for b in buffers]
    n = len(views)
    ptrs = (C.c_void_p * max(n, 1))(*[v.ctypes.data if v.size else None for v in views])
    lens = np.array([v.size for v in views], dtype=np.uint64)
    kind = np.array([0 if k in (0, "adler32") else 1 for k in kinds], dtype=np.uint8)
    sd = None if seeds is None else np.array([_i32(s) for s in seeds], dtype=np.int32)
    out = np.zeros(max(n, 1), dtype=np.int32)
    ctx.check(ctx.lib.sdz_checksum_batch(ctx.h, ptrs, lens.ctypes.data, kind.ctypes.data,
                                         None if sd is None else sd.ctypes.data, n, out.ctypes.data))
    return [int(x) for x in out[:n]]


def deflate_wrap_batch(payloads, sources, format="deflate", file_names=None, dictionaries=None, mtime=None, ctx=None):
    """What `Deflater` puts around its raw deflate output (src/sd-deflate.ts:98-165), for a batch: the source checksums
    are computed on the device in one launch, headers / trailers are the reference's.  The compressor itself is not
    part of this package (out of scope) - `payloads` come from the caller's deflate."""
    import time as _time
    ctx = ctx or N.default_context()
    fmt = {"raw": 0, "deflate": 1, "gzip": 2}
    if format not in fmt:
        raise ValueError("container must be one of `raw`, `deflate`, `gzip`")       # RangeError in JS
    n = len(payloads)
    pv = [_as_u8(p, "data must be an ArrayBuffer or buffer view") for p in payloads]
    sv = [_as_u8(p, "data must be an ArrayBuffer or buffer view") for p in sources]
    ins = (N.WrapIn * max(n, 1))()
    t = int(_time.time()) if mtime is None else int(mtime)
    names = []
    for i in range(n):
        ins[i].payload = pv[i].ctypes.data if pv[i].size else None
        ins[i].payload_len = int(pv[i].size)
        ins[i].source = sv[i].ctypes.data if sv[i].size else None
        ins[i].source_len = int(sv[i].size)
        ins[i].format = fmt[format]
        ins[i].mtime = t & 0xFFFFFFFF
        name = (file_names[i] if file_names else None) or ""
        if name and format == "gzip":               # src/sd-deflate.ts:126-131: code points above 0xff become "_"
            names.append(bytes(ord(c) if ord(c) <= 0xff else 95 for c in name))
            ins[i].file_name = names[-1]
        d = dictionaries[i] if dictionaries else None
        if d is not None:
            if format != "deflate":
                raise TypeError("Can only provide a dictionary for `deflate` containers.")
            ins[i].dict_adler = adler32(d, ctx=ctx)
    sizes = np.zeros(max(n, 1), dtype=np.uint64)
    ctx.check(ctx.lib.sdz_deflate_wrap_sizes(ins, n, sizes.ctypes.data))
    off = np.zeros(max(n, 1), dtype=np.uint64)
    off[1:] = np.cumsum(sizes[:-1])
    arena = np.empty(max(int(sizes[:n].sum()), 1), dtype=np.uint8)
    ctx.check(ctx.lib.sdz_deflate_wrap_batch(ctx.h, ins, n, arena.ctypes.data, off.ctypes.data, None))
    return [arena[int(off[i]):int(off[i]) + int(sizes[i])].tobytes() for i in range(n)]


def mergeBuffers(buffers: Sequence[bytes]) -> bytes:
    """src/common.ts:116-126"""
    return b"".join(bytes(b) for b in buffers)


# ------------------------------------------------------------------ inflate

@dataclass
class InflateResult:
    """src/sd-inflate.ts:39-52"""
    success: bool
    complete: bool
    checksum: str
    fileSize: str
    fileName: str
    modDate: Optional[datetime.datetime]


def _record_to_result(r, data_view) -> InflateResult:
    name = bytes(data_view[r.name_off:r.name_off + r.name_len]).decode("latin-1") if r.name_len else ""
    # `new Date(mtime * 1000)` with a signed int32 mtime
    mod = None if r.mtime == 0 else datetime.datetime.fromtimestamp(r.mtime, tz=datetime.timezone.utc)
    return InflateResult(bool(r.success), bool(r.complete), _CHECK[r.checksum_state], _CHECK[r.size_state], name, mod)


class InflateError(ValueError):
    """Error thrown by Inflater.append() / inflate(); .thrown / .msg_id are the record's codes."""

    def __init__(self, text, thrown, msg_id):
        super().__init__(text)
        self.thrown, self.msg_id = thrown, msg_id


class InflateHang(RuntimeError):
    """The reference's append() never returns on this input; the drop-in raises instead of spinning."""

    def __init__(self, text, thrown, msg_id):
        super().__init__(text)
        self.thrown, self.msg_id = thrown, msg_id


def _raise_thrown(thrown, msg_id):
    text = _THROWN[thrown] + (_MSG[msg_id] if thrown == 4 else "")
    raise (InflateError if thrown != 6 else InflateHang)(text, thrown, msg_id if thrown == 4 else 0)


PARITY_REFERENCE, PARITY_SPEC = 0, 1


def inflate_batch_raw(views, dictionaries=None, modes=None, caps=None, ctx=None, flags=PARITY_REFERENCE):
    """One sdz_inflate_batch call.  Returns (out_arena: np.uint8[], out_off, records).
    flags = PARITY_SPEC: RFC 1950 / 1951 / 1952 behaviour as zlib 1.3 implements it instead of the reference's
    (no Q1 / Q2 / Q3 / Q4 / Q5 / Q6 / Q9 / Q10 / Q13 / Q14 / Q15 emulation)."""
    ctx = ctx or N.default_context()
    n = len(views)
    ins = (N.In * max(n, 1))()
    keep = []
    for i, v in enumerate(views):
        ins[i].data = v.ctypes.data if v.size else None
        ins[i].len = int(v.size)
        d = None if dictionaries is None else dictionaries[i]
        if d is not None:
            dv = _as_u8(d, "options.dictionary must be undefined or a buffer or a buffer view")
            keep.append(dv)
            # a zero-length dictionary is still "a dictionary was supplied"
            ins[i].dict = dv.ctypes.data if dv.size else C.cast(C.create_string_buffer(1), C.c_void_p).value
            ins[i].dict_len = int(dv.size)
        ins[i].mode = MODE_SNIFF if modes is None else modes[i]
    if caps is None:
        sizes = np.zeros(max(n, 1), dtype=np.uint64)
        ctx.check(ctx.lib.sdz_inflate_sizes(ctx.h, ins, n, sizes.ctypes.data, flags))
        caps = sizes[:n]
    caps = np.ascontiguousarray(caps, dtype=np.uint64)
    slot = (caps + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        off[1:] = np.cumsum(slot[:-1])
    total = int(slot.sum()) if n else 0
    arena = np.empty(max(total, 1), dtype=np.uint8)
    res = (N.Result * max(n, 1))()
    ctx.check(ctx.lib.sdz_inflate_batch(ctx.h, ins, n, arena.ctypes.data, off.ctypes.data, slot.ctypes.data, res, flags),
              allow=(N.SDZ_E_OUT_CAP,))
    return arena, off, res


def inflateBatch(buffers, dictionaries=None, raw=None, ctx=None, parity="reference"):
    """inflateBatch(buffers[]) - the one entry point added to the reference API.

    Each buffer is decoded exactly like `inflate(buffer, dictionary)` (container sniffing
    included) unless raw[i] is given, in which case it behaves like
    `new Inflater({raw: raw[i], dictionary}).append(buffer); finish()`.  Errors are recorded
    per stream instead of thrown: returns a list of dicts {data, result, error}.
    """
    views = [_as_u8(b, "data must be an ArrayBuffer or buffer view") for b in buffers]
    modes = None
    if raw is not None:
        modes = [MODE_SNIFF if r is None else (MODE_RAW if r else MODE_INFLATER) for r in raw]
    if parity not in ("reference", "spec"):
        raise ValueError("parity must be `reference` or `spec`")
    arena, off, res = inflate_batch_raw(views, dictionaries, modes, None, ctx, PARITY_SPEC if parity == "spec" else PARITY_REFERENCE)
    out = []
    for i, v in enumerate(views):
        r = res[i]
        err = None
        sniff = modes is None or modes[i] == MODE_SNIFF
        thrown = r.thrown_inflate if sniff else r.thrown_append
        if thrown == 12:
            err = "options.dictionary cannot be set when options.raw is true"
        elif thrown:
            err = _THROWN[thrown] + (_MSG[r.msg_id] if thrown == 4 else "")
        data = bytes(arena[int(off[i]):int(off[i]) + int(r.out_len)]) if (r.out_len and not r.thrown_append) else b""
        out.append({"data": data, "result": _record_to_result(r, v), "error": err, "record": r})
    return out


def inflate_large_raw(view, mode=MODE_SNIFF, cap=None, ctx=None):
    """One sdz_inflate_large call (a single big stream spread over the whole GPU).
    Returns (out: np.uint8[], record)."""
    ctx = ctx or N.default_context()
    res = N.Result()
    if cap is None:
        rc = ctx.check(ctx.lib.sdz_inflate_large(ctx.h, view.ctypes.data, int(view.size), mode, 0, None, 0, C.byref(res)),
                       allow=(N.SDZ_E_OUT_CAP,))
        if rc == N.SDZ_OK:                      # empty output (or a stream that failed before producing anything)
            return np.empty(0, dtype=np.uint8), res
        cap = int(res.out_len)
    out = np.empty(max(int(cap), 1), dtype=np.uint8)
    ctx.check(ctx.lib.sdz_inflate_large(ctx.h, view.ctypes.data, int(view.size), mode, 0, out.ctypes.data, int(cap), C.byref(res)),
              allow=(N.SDZ_E_OUT_CAP,))
    return out, res


def inflateLarge(data, raw=None, ctx=None):
    """`new Inflater({raw}).append(data); finish()` for ONE large stream, decoded block-parallel.
    Returns {data, result, error, record} like an element of inflateBatch()."""
    view = _as_u8(data, "data must be an ArrayBuffer or buffer view")
    mode = MODE_SNIFF if raw is None else (MODE_RAW if raw else MODE_INFLATER)
    out, r = inflate_large_raw(view, mode, None, ctx)
    thrown = r.thrown_inflate if mode == MODE_SNIFF else r.thrown_append
    err = None
    if thrown:
        err = _THROWN[thrown] + (_MSG[r.msg_id] if thrown == 4 else "")
    body = bytes(out[:int(r.out_len)]) if (r.out_len and not r.thrown_append) else b""
    return {"data": body, "result": _record_to_result(r, view), "error": err, "record": r}


def inflate(data, dictionary=None, ctx=None) -> bytes:
    """inflate(data, dictionary?) - src/sd-inflate.ts:189-228 (auto-detects the container)."""
    view = _as_u8(data, "data must be an ArrayBuffer or buffer view")
    if view.size < 2:
        raise ValueError("data buffer is too small")
    arena, off, res = inflate_batch_raw([view], [dictionary], [MODE_SNIFF], None, ctx)
    r = res[0]
    if r.thrown_inflate == 12:
        raise ValueError("options.dictionary cannot be set when options.raw is true")   # RangeError in JS
    if r.thrown_inflate:
        _raise_thrown(r.thrown_inflate, r.msg_id)
    return bytes(arena[:int(r.out_len)])


class Inflater:
    """class Inflater - src/sd-inflate.ts:54-180, over a device session (sdz_inflater_*).

    The compressed bytes received so far and the decoded bytes stay in HBM; between append() calls the session keeps
    where the reference stopped (an sdz_resume record), so every call decodes only its own input.  The reference's
    behaviour at chunk boundaries is reproduced, defects included: see include/sdzcuda.h.
    """

    def __init__(self, raw=None, dictionary=None, ctx=None):
        if raw is not None and raw is not True and raw is not False:
            raise TypeError("options.raw must be undefined or true or false")
        self._raw = bool(raw)
        self._dict = None
        if dictionary is not None:
            if self._raw:
                raise ValueError("options.dictionary cannot be set when options.raw is true")   # RangeError
            self._dict = _as_u8(dictionary, "options.dictionary must be undefined or a buffer or a buffer view")
        self._ctx = ctx
        self._h = None
        self._rec = None

    def _session(self):
        if self._h is None:
            self._ctx = self._ctx or N.default_context()
            h = C.c_void_p()
            d = self._dict
            dptr = None
            if d is not None:                      # a zero-length dictionary is still "a dictionary was supplied"
                dptr = d.ctypes.data if d.size else C.cast(C.create_string_buffer(1), C.c_void_p).value
            self._ctx.check(self._ctx.lib.sdz_inflater_create(self._ctx.h, int(self._raw), dptr, 0 if d is None else int(d.size), C.byref(h)))
            self._h = h
        return self._h

    def append(self, data) -> List[bytes]:
        chunk = _as_u8(data, "data must be an ArrayBuffer or buffer view")
        if chunk.size == 0:
            return []
        h = self._session()
        lib = self._ctx.lib
        nb = C.c_uint64()
        rec = N.Result()
        self._ctx.check(lib.sdz_inflater_append(h, chunk.ctypes.data, int(chunk.size), C.byref(nb), C.byref(rec)))
        self._rec = rec
        if rec.thrown_append:
            _raise_thrown(rec.thrown_append, rec.msg_id)
        n = int(nb.value)
        if n == 0:
            return []
        out = np.empty(n, dtype=np.uint8)
        self._ctx.check(lib.sdz_inflater_read(h, out.ctypes.data, n))
        new = out.tobytes()
        return [new[i:i + OUTPUT_BUFSIZE] for i in range(0, n, OUTPUT_BUFSIZE)]

    def finish(self) -> InflateResult:
        if self._rec is None:
            return InflateResult(False, False, "unchecked", "unchecked", "", None)
        rec = N.Result()
        self._ctx.check(self._ctx.lib.sdz_inflater_finish(self._h, C.byref(rec)))
        name = np.zeros(max(1, int(rec.name_len)), dtype=np.uint8)
        if rec.name_len:
            self._ctx.check(self._ctx.lib.sdz_inflater_input(self._h, int(rec.name_off), int(rec.name_len), name.ctypes.data))
        r = InflateResult(bool(rec.success), bool(rec.complete), _CHECK[rec.checksum_state], _CHECK[rec.size_state],
                          bytes(name[:int(rec.name_len)]).decode("latin-1") if rec.name_len else "",
                          None if rec.mtime == 0 else datetime.datetime.fromtimestamp(rec.mtime, tz=datetime.timezone.utc))
        self.record = rec
        return r

    def close(self):
        if self._h is not None:
            self._ctx.lib.sdz_inflater_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
