"""ctypes binding of libsdzcuda.so (C ABI in include/sdzcuda.h).

There is no fallback: if the shared library or a B200 is missing, importing works but the
first call raises.  Nothing here imports the CPU oracle."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDZ_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "csrc", "libsdzcuda.so"))

SDZ_OK, SDZ_E_NO_DEVICE, SDZ_E_CUDA, SDZ_E_ARG, SDZ_E_NOMEM, SDZ_E_OUT_CAP, SDZ_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
_ERR_NAME = {-1: "no sm_100a CUDA device", -2: "CUDA error", -3: "bad argument", -4: "out of memory",
             -5: "output slot too small", -6: "unsupported"}

EXPORTS = [
    "sdz_ctx_create", "sdz_ctx_create_multi", "sdz_ctx_device_count", "sdz_last_partition", "sdz_host_alloc_near", "sdz_ctx_numa_node", "sdz_ctx_destroy", "sdz_last_error", "sdz_version", "sdz_launch_count", "sdz_last_timing", "sdz_last_phase_timing", "sdz_last_fast_stats", "sdz_debug_table_totals",
    "sdz_host_alloc", "sdz_host_free", "sdz_device_alloc", "sdz_device_free", "sdz_memcpy_h2d", "sdz_memcpy_d2h",
    "sdz_adler32", "sdz_crc32", "sdz_adler32_chain", "sdz_crc32_chain", "sdz_checksum_batch",
    "sdz_deflate_wrap_sizes", "sdz_deflate_wrap_batch",
    "sdz_inflater_create", "sdz_inflater_append", "sdz_inflater_read", "sdz_inflater_finish", "sdz_inflater_input", "sdz_inflater_destroy",
    "sdz_inflate_batch", "sdz_inflate_sizes", "sdz_inflate_batch_device", "sdz_inflate_large", "sdz_sync",
    "sdz_large_open", "sdz_large_close", "sdz_large_index", "sdz_large_plan", "sdz_large_range", "sdz_large_decode",
    "sdz_large_windows", "sdz_large_resolve", "sdz_large_finish", "sdz_large_is_gzip", "sdz_crc32_combine", "sdz_adler32_combine",
]


class Result(C.Structure):
    """struct sdz_result (include/sdz_codes.h)"""
    _fields_ = [
        ("out_off", C.c_uint64), ("out_len", C.c_uint64), ("total_in", C.c_uint64),
        ("zstatus", C.c_int32), ("stored_checksum", C.c_int32), ("running_checksum", C.c_int32),
        ("stored_isize", C.c_int32), ("mtime", C.c_int32),
        ("name_off", C.c_uint32), ("name_len", C.c_uint32), ("n_blocks", C.c_uint32),
        ("msg_id", C.c_uint8), ("thrown_append", C.c_uint8), ("thrown_inflate", C.c_uint8), ("container", C.c_uint8),
        ("complete", C.c_uint8), ("checksum_state", C.c_uint8), ("size_state", C.c_uint8), ("success", C.c_uint8),
        ("have_running", C.c_uint8), ("reserved", C.c_uint8 * 7),
    ]
    OBSERVABLE = ("out_len", "stored_checksum", "running_checksum", "stored_isize", "mtime", "name_len",
                  "msg_id", "thrown_append", "thrown_inflate", "container", "complete", "checksum_state",
                  "size_state", "success", "have_running")

    def observable(self):
        d = {k: getattr(self, k) for k in self.OBSERVABLE}
        if d["thrown_append"]:
            return {"thrown_append": d["thrown_append"], "thrown_inflate": d["thrown_inflate"],
                    "msg_id": d["msg_id"] if d["thrown_append"] == 4 else 0}
        d["msg_id"] = 0
        return d


class In(C.Structure):
    """struct sdz_in"""
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("dict", C.c_void_p), ("dict_len", C.c_uint32),
                ("mode", C.c_uint8), ("reserved", C.c_uint8 * 3)]


class WrapIn(C.Structure):
    """struct sdz_wrap_in"""
    _fields_ = [("payload", C.c_void_p), ("payload_len", C.c_uint64), ("source", C.c_void_p), ("source_len", C.c_uint64),
                ("file_name", C.c_char_p), ("mtime", C.c_uint32), ("dict_adler", C.c_int32), ("format", C.c_uint8),
                ("reserved", C.c_uint8 * 7)]


class BatchDev(C.Structure):
    """struct sdz_batch_dev"""
    _fields_ = [(k, C.c_void_p) for k in ("d_in", "d_in_off", "d_in_len", "d_mode", "d_dict", "d_dict_off",
                                          "d_dict_len", "d_dict_adler", "d_out", "d_out_off", "d_out_cap",
                                          "d_results")] + [("n", C.c_uint64)]


class NativeError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load():
    """dlopen libsdzcuda.so and declare prototypes (no CUDA call is made)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError("libsdzcuda.so is not built (%s); run `python __graft_entry__.py build`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32
        L.sdz_ctx_create.argtypes = [C.c_int, u32, C.POINTER(vp)]
        L.sdz_ctx_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, u32, C.POINTER(vp)]
        L.sdz_ctx_device_count.argtypes = [vp]
        L.sdz_last_partition.argtypes = [vp, vp, C.c_int]
        L.sdz_host_alloc_near.argtypes = [vp, C.c_size_t]
        L.sdz_host_alloc_near.restype = vp
        L.sdz_ctx_numa_node.argtypes = [vp]
        L.sdz_ctx_destroy.argtypes = [vp]
        L.sdz_ctx_destroy.restype = None
        L.sdz_last_error.argtypes = [vp]
        L.sdz_last_error.restype = C.c_char_p
        L.sdz_version.restype = C.c_char_p
        L.sdz_launch_count.argtypes = [vp]
        L.sdz_launch_count.restype = u64
        L.sdz_last_timing.argtypes = [vp, C.POINTER(C.c_float * 3)]
        L.sdz_last_phase_timing.argtypes = [vp, C.POINTER(C.c_float * 5)]
        L.sdz_last_fast_stats.argtypes = [vp, C.POINTER(C.c_uint64 * 2)]
        L.sdz_debug_table_totals.argtypes = [vp, vp, vp, vp, u64, C.c_int, vp]
        L.sdz_host_alloc.argtypes = [C.c_size_t]
        L.sdz_host_alloc.restype = vp
        L.sdz_host_free.argtypes = [vp]
        L.sdz_host_free.restype = None
        L.sdz_device_alloc.argtypes = [vp, C.c_size_t]
        L.sdz_device_alloc.restype = vp
        L.sdz_device_free.argtypes = [vp, vp]
        L.sdz_device_free.restype = None
        L.sdz_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
        L.sdz_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
        L.sdz_sync.argtypes = [vp]
        for f in (L.sdz_adler32, L.sdz_crc32):
            f.argtypes = [vp, vp, u64, i32, C.c_int, C.POINTER(i32)]
        for f in (L.sdz_adler32_chain, L.sdz_crc32_chain):
            f.argtypes = [vp, vp, vp, u64, i32, C.c_int, vp, C.POINTER(i32)]
        L.sdz_checksum_batch.argtypes = [vp, vp, vp, vp, vp, u64, vp]
        L.sdz_deflate_wrap_sizes.argtypes = [vp, u64, vp]
        L.sdz_deflate_wrap_batch.argtypes = [vp, vp, u64, vp, vp, vp]
        L.sdz_inflater_create.argtypes = [vp, C.c_int, vp, u32, C.POINTER(vp)]
        L.sdz_inflater_append.argtypes = [vp, vp, u64, C.POINTER(u64), C.POINTER(Result)]
        L.sdz_inflater_read.argtypes = [vp, vp, u64]
        L.sdz_inflater_finish.argtypes = [vp, C.POINTER(Result)]
        L.sdz_inflater_input.argtypes = [vp, u64, u64, vp]
        L.sdz_inflater_destroy.argtypes = [vp]
        L.sdz_inflater_destroy.restype = None
        L.sdz_inflate_batch.argtypes = [vp, vp, u64, vp, vp, vp, vp, u32]
        L.sdz_inflate_sizes.argtypes = [vp, vp, u64, vp, u32]
        L.sdz_inflate_batch_device.argtypes = [vp, C.POINTER(BatchDev), u32, C.c_int]
        L.sdz_inflate_large.argtypes = [vp, vp, u64, C.c_uint8, C.c_int, vp, u64, C.POINTER(Result)]
        L.sdz_large_open.argtypes = [vp, vp, u64, C.c_uint8, C.c_int, C.POINTER(vp)]
        L.sdz_large_close.argtypes = [vp]
        L.sdz_large_close.restype = None
        L.sdz_large_index.argtypes = [vp, u32, u32, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
        L.sdz_large_plan.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.sdz_large_range.argtypes = [vp, u32, u32, C.POINTER(u64), C.POINTER(u64)]
        L.sdz_large_decode.argtypes = [vp, u32, u32, vp]
        L.sdz_large_windows.argtypes = [vp]
        L.sdz_large_resolve.argtypes = [vp]
        L.sdz_large_finish.argtypes = [vp, i32, C.POINTER(Result)]
        L.sdz_large_is_gzip.argtypes = [vp]
        L.sdz_crc32_combine.argtypes = [i32, i32, u64]
        L.sdz_crc32_combine.restype = i32
        L.sdz_adler32_combine.argtypes = [i32, i32, u64]
        L.sdz_adler32_combine.restype = i32
        _lib = L
        return L


class Context:
    """sdz_ctx wrapper: one per (process, device)."""

    def __init__(self, device=0):
        """device: one index, or a list of indices (multi-device context: batches are partitioned per stream)."""
        self.lib = load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            rc = self.lib.sdz_ctx_create_multi(devs, len(device), 0, C.byref(h))
        else:
            rc = self.lib.sdz_ctx_create(device, 0, C.byref(h))
        if rc != SDZ_OK:
            raise NativeError("sdz_ctx_create(device=%s) failed: %s" % (device, _ERR_NAME.get(rc, rc)))
        self.h = h
        self.device = device

    def device_count(self):
        return int(self.lib.sdz_ctx_device_count(self.h))

    def last_partition(self):
        n = self.device_count() + 1
        cut = (C.c_uint64 * n)()
        self.check(self.lib.sdz_last_partition(self.h, cut, n))
        return [int(x) for x in cut]

    def check(self, rc, allow=()):
        if rc != SDZ_OK and rc not in allow:
            msg = self.lib.sdz_last_error(self.h).decode() if rc == SDZ_E_CUDA else ""
            raise NativeError("libsdzcuda: %s %s" % (_ERR_NAME.get(rc, rc), msg))
        return rc

    def launch_count(self):
        return int(self.lib.sdz_launch_count(self.h))

    def last_timing(self):
        ms = (C.c_float * 3)()
        self.check(self.lib.sdz_last_timing(self.h, C.byref(ms)))
        return [float(x) for x in ms]

    def last_phase_timing(self):
        ms = (C.c_float * 5)()
        self.check(self.lib.sdz_last_phase_timing(self.h, C.byref(ms)))
        return [float(x) for x in ms]

    def last_fast_stats(self):
        v = (C.c_uint64 * 2)()
        self.check(self.lib.sdz_last_fast_stats(self.h, C.byref(v)))
        return int(v[0]), int(v[1])

    def close(self):
        if getattr(self, "h", None):
            self.lib.sdz_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("SDZ_DEVICE", "0")))
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx
