"""ctypes binding of oracle/_build/libsdzoracle.so (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY.  Function names mirror the reference API they restate:
adler32 / crc32 (src/adler32.ts:17, src/crc32.ts:17), Inflater (src/sd-inflate.ts:54),
inflate_oneshot == new Inflater + append + finish (+ inflate() throw mapping, :189-228).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdzoracle.so")

MSG_TEXT = [
    "", "invalid gzip id", "unknown compression method", "invalid window size", "incorrect header check",
    "need dictionary", "invalid block type", "invalid stored block lengths", "too many length or distance symbols",
    "invalid bit length repeat", "oversubscribed dynamic bit lengths tree", "incomplete dynamic bit lengths tree",
    "oversubscribed literal/length tree", "incomplete literal/length tree", "oversubscribed distance tree",
    "incomplete distance tree", "empty distance tree with lengths", "invalid distance code",
    "invalid literal/length code",
]
THROWN_TEXT = [
    "", "inflate error: bad input", "Custom dictionary is not valid for this data",
    "Custom dictionary required for this data", "inflate error: ", "inflate error: bad input data",
    "<reference does not terminate>", "data buffer is too small", "Unexpected EOF during decompression",
    "Data integrity check failed", "Data size check failed", "Decompression error",
]
(THROW_NONE, THROW_BAD_INPUT, THROW_DICT_INVALID, THROW_DICT_REQUIRED, THROW_INFLATE_ERROR, THROW_BAD_INPUT_DATA,
 THROW_HANG, THROW_TOO_SMALL, THROW_UNEXPECTED_EOF, THROW_INTEGRITY, THROW_SIZE_CHECK, THROW_DECOMPRESSION) = range(12)
MODE_SNIFF, MODE_INFLATER, MODE_RAW = 0, 1, 2
CHECK_TEXT = ["unchecked", "match", "mismatch"]


class Result(C.Structure):
    """struct sdz_result (include/sdz_codes.h)."""
    _fields_ = [
        ("out_off", C.c_uint64), ("out_len", C.c_uint64), ("total_in", C.c_uint64),
        ("zstatus", C.c_int32), ("stored_checksum", C.c_int32), ("running_checksum", C.c_int32),
        ("stored_isize", C.c_int32), ("mtime", C.c_int32),
        ("name_off", C.c_uint32), ("name_len", C.c_uint32), ("n_blocks", C.c_uint32),
        ("msg_id", C.c_uint8), ("thrown_append", C.c_uint8), ("thrown_inflate", C.c_uint8), ("container", C.c_uint8),
        ("complete", C.c_uint8), ("checksum_state", C.c_uint8), ("size_state", C.c_uint8), ("success", C.c_uint8),
        ("have_running", C.c_uint8), ("reserved", C.c_uint8 * 7),
    ]

    # fields that define the observable record (out_off / n_blocks / zstatus are diagnostics)
    OBSERVABLE = ("out_len", "stored_checksum", "running_checksum", "stored_isize", "mtime", "name_len",
                  "msg_id", "thrown_append", "thrown_inflate", "container", "complete", "checksum_state",
                  "size_state", "success", "have_running")

    def as_dict(self, keys=None):
        keys = keys or [f[0] for f in self._fields_ if f[0] != "reserved"]
        return {k: getattr(self, k) for k in keys}

    def observable(self):
        d = self.as_dict(self.OBSERVABLE)
        if d["thrown_append"]:
            # nothing but the exception is observable once append() throws
            return {"thrown_append": d["thrown_append"], "thrown_inflate": d["thrown_inflate"],
                    "msg_id": d["msg_id"] if d["thrown_append"] == THROW_INFLATE_ERROR else 0}
        d["msg_id"] = 0
        return d


class _Chunks(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t),
                ("chunk_len", C.POINTER(C.c_uint32)), ("n_chunks", C.c_size_t), ("chunk_cap", C.c_size_t)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("sdz_oracle.c", "sdz_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.sdzo_adler32.restype = C.c_int32
        L.sdzo_adler32.argtypes = [C.c_void_p, C.c_uint64, C.c_int32]
        L.sdzo_crc32.restype = C.c_int32
        L.sdzo_crc32.argtypes = [C.c_void_p, C.c_uint64, C.c_int32]
        L.sdzo_inflater_new.restype = C.c_void_p
        L.sdzo_inflater_new.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.sdzo_inflater_free.argtypes = [C.c_void_p]
        L.sdzo_chunks_free.argtypes = [C.POINTER(_Chunks)]
        L.sdzo_append.restype = C.c_int
        L.sdzo_append.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_Chunks)]
        L.sdzo_finish.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.sdzo_file_name.restype = C.c_size_t
        L.sdzo_file_name.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8))]
        L.sdzo_inflate_oneshot.restype = C.c_int
        L.sdzo_inflate_oneshot.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                                           C.c_void_p, C.c_size_t, C.POINTER(Result)]
        L.sdzo_inflate_batch_mt.restype = C.c_int
        L.sdzo_inflate_batch_mt.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + [C.c_void_p] * 4 + [C.c_int]
        L.sdzo_fixed_tables.argtypes = [C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int),
                                        C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _buf(b):
    b = bytes(b) if not isinstance(b, (bytes, bytearray)) else b
    return (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b if len(b) else b"\0"), len(b)


def adler32(data, seed=1):
    buf, n = _buf(data)
    return lib().sdzo_adler32(buf, n, C.c_int32(seed & 0xFFFFFFFF if seed >= 0 else seed).value)


def crc32(data, seed=0):
    buf, n = _buf(data)
    return lib().sdzo_crc32(buf, n, C.c_int32(seed & 0xFFFFFFFF if seed >= 0 else seed).value)


class OracleThrow(Exception):
    def __init__(self, thrown, msg_id):
        self.thrown, self.msg_id = thrown, msg_id
        text = THROWN_TEXT[thrown] + (MSG_TEXT[msg_id] if thrown == THROW_INFLATE_ERROR else "")
        super().__init__(text)


class Inflater:
    """Oracle twin of the reference class Inflater (src/sd-inflate.ts:54-180)."""

    def __init__(self, raw=False, dictionary=None):
        if raw and dictionary is not None:
            raise ValueError("options.dictionary cannot be set when options.raw is true")
        if dictionary is not None:
            dbuf, dn = _buf(dictionary)
            self._h = lib().sdzo_inflater_new(int(bool(raw)), dbuf, dn)
        else:
            self._h = lib().sdzo_inflater_new(int(bool(raw)), None, 0)

    def append(self, data):
        buf, n = _buf(data)
        ch = _Chunks()
        thrown = lib().sdzo_append(self._h, buf, n, C.byref(ch))
        try:
            if thrown:
                r = Result()
                lib().sdzo_finish(self._h, C.byref(r))
                raise OracleThrow(thrown, r.msg_id)
            out, off = [], 0
            raw = C.string_at(ch.data, ch.len) if ch.len else b""
            for i in range(ch.n_chunks):
                out.append(raw[off:off + ch.chunk_len[i]])
                off += ch.chunk_len[i]
            return out
        finally:
            lib().sdzo_chunks_free(C.byref(ch))

    def finish(self):
        r = Result()
        lib().sdzo_finish(self._h, C.byref(r))
        p = C.POINTER(C.c_uint8)()
        n = lib().sdzo_file_name(self._h, C.byref(p))
        r.file_name = C.string_at(p, n).decode("latin-1") if n else ""
        return r

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sdzo_inflater_free(self._h)
            self._h = None


def inflate_oneshot(data, dictionary=None, mode=MODE_SNIFF, out_cap=None):
    """new Inflater(opts).append(data) + finish() (+ inflate() throw mapping for MODE_SNIFF).
    Returns (bytes, Result)."""
    buf, n = _buf(data)
    cap = out_cap if out_cap is not None else max(1 << 16, n * 8)
    r = Result()
    dbuf, dn = (None, 0) if dictionary is None else _buf(dictionary)
    while True:
        out = (C.c_uint8 * cap)()
        rc = lib().sdzo_inflate_oneshot(buf, n, dbuf, dn, mode, out, cap, C.byref(r))
        if rc == 0:
            return bytes(out[:r.out_len]) if r.out_len else b"", r
        cap = r.out_len


def inflate_oneshot_np(data, out, mode=MODE_SNIFF):
    """Same call on numpy uint8 arrays without any Python-side copy (bench.py's CPU baseline leg for one
    large stream).  Returns the Result; `out` must be large enough."""
    r = Result()
    rc = lib().sdzo_inflate_oneshot(data.ctypes.data_as(C.POINTER(C.c_uint8)), data.size, None, 0, mode,
                                    out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size, C.byref(r))
    if rc != 0:
        raise ValueError("output buffer too small: need %d" % r.out_len)
    return r


def inflate_batch_mt(in_arena, in_off, in_len, out_off, out_cap, n_threads, modes=None, out_arena=None):
    """Threaded batch of one-shots over numpy arrays (CPU baseline).  Returns (out_arena, results)."""
    import numpy as np
    n = len(in_off)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
    if out_arena is None:
        out_arena = np.empty(int((out_off + out_cap).max()) if n else 1, dtype=np.uint8)
    res = (Result * n)()
    mp = None if modes is None else np.ascontiguousarray(modes, dtype=np.uint8).ctypes.data
    rc = lib().sdzo_inflate_batch_mt(in_arena.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, mp, n,
                                     out_arena.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                     C.addressof(res), n_threads)
    if rc:
        raise RuntimeError("oracle batch: output capacity too small")
    return out_arena, res


def table_usage(lit_lens, dist_lens):
    """(lit entries, dist entries, status_lit, status_dist) of huft_build with the MANY limit lifted"""
    lens = bytes(lit_lens) + bytes(dist_lens)
    buf, _ = _buf(lens)
    a, b = C.c_int(), C.c_int()
    st = (C.c_int * 2)()
    lib().sdzo_table_usage(buf, len(lit_lens), len(dist_lens), C.byref(a), C.byref(b), st)
    return a.value, b.value, st[0], st[1]


def fixed_tables():
    tl, td = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    ntl, ntd = C.c_int(), C.c_int()
    lib().sdzo_fixed_tables(C.byref(tl), C.byref(ntl), C.byref(td), C.byref(ntd))
    return [tl[i] for i in range(ntl.value * 3)], [td[i] for i in range(ntd.value * 3)]
