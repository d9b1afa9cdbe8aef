"""One large stream spread over several parts (GPUs): host-side driver of the sdz_large_* phases.

The reference decodes a stream strictly sequentially through one 32 KiB window (src/sd-inflate.ts:87-153);
nothing in it corresponds to this file.  What is reproduced is its RESULT: the bytes and the finish() record
(src/sd-inflate.ts:159-179).

Protocol for rank r of n (every rank holds the whole compressed stream):

    index(r, n)                      candidate blocks + resume points found in slice r of the compressed bits
    all-gather of those records      (the only variable-size exchange; a few hundred KB for a 1 GiB stream)
    plan(all records)                identical on every rank: the chain of real blocks, cut into pieces
    range(r, n)                      the slice [lo, hi) of the OUTPUT that rank r produces
    decode(r, n, out)                pieces -> marker symbols, windows composed inside the slice
    recv 32 KiB from r - 1           the final window before `lo`   (rank 0: nothing to wait for)
    windows()                        the last 32 KiB of the slice become final
    send 32 KiB to r + 1
    resolve()                        every remaining marker of the slice
    crc32(slice), all-gather (crc, length), combine in rank order -> finish() record on every rank

`Backend` is what a rank does locally (CudaBackend binds libsdzcuda.so; tests substitute a CPU stand-in for the
message flow), `Comm` is how ranks talk (TorchComm = torch.distributed, NCCL on GPUs / gloo on CPUs; LoopComm runs all
"ranks" in one process, which is also how several parts are exercised on a single GPU).
"""
import ctypes as C

import numpy as np

from . import _native as N

WIN = 32768
MODE_SNIFF, MODE_INFLATER, MODE_RAW = 0, 1, 2

BLOCK_DT = np.dtype([("bit", "<u8"), ("end_bit", "<u8"), ("out_len", "<u8"), ("last", "u1"), ("btype", "u1"), ("ok", "u1"),
                     ("reserved", "u1", (5,))])
CKPT_DT = np.dtype([("block_bit", "<u8"), ("bit", "<u8"), ("pos", "<u4"), ("reserved", "<u4")])
assert BLOCK_DT.itemsize == 32 and CKPT_DT.itemsize == 24


class NeedsSequentialDecoder(Exception):
    """The stream is not one the block-parallel path handles (stored blocks, preset dictionary, truncated or damaged
    data, trailing bytes): decode it with inflateBatch / inflate instead."""


def crc32_combine(crc_a, crc_b, len_b):
    """crc32(A || B) from the parts (signed int32 in and out, like the reference's values)."""
    return int(N.load().sdz_crc32_combine(int(crc_a), int(crc_b), int(len_b)))


def adler32_combine(adler_a, adler_b, len_b):
    """adler32(A || B) from adler32(A), adler32(B) (both with the STANDARD arithmetic, B seeded with 1) and len(B)."""
    return int(N.load().sdz_adler32_combine(int(adler_a), int(adler_b), int(len_b)))


OUTBUF = 16384                              # src/zstream.ts:11


class CudaBackend:
    """The phases of one rank, on one device, through the C ABI.  `stream_ptr` is a DEVICE pointer to the whole
    compressed stream (16-byte aligned, readable for 1 KiB past its end)."""

    def __init__(self, ctx, stream_ptr, length, mode=MODE_SNIFF):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        rc = self.lib.sdz_large_open(ctx.h, stream_ptr, int(length), mode, 1, C.byref(h))
        self._check(rc)
        self.h = h

    def _check(self, rc):
        if rc == N.SDZ_E_UNSUPPORTED:
            raise NeedsSequentialDecoder()
        self.ctx.check(rc)

    def index(self, part, n_parts):
        bp, cp = C.c_void_p(), C.c_void_p()
        nb, nc = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sdz_large_index(self.h, part, n_parts, C.byref(bp), C.byref(nb), C.byref(cp), C.byref(nc)))
        blocks = np.empty(nb.value, dtype=BLOCK_DT)
        ckpts = np.empty(nc.value, dtype=CKPT_DT)
        if nb.value:
            C.memmove(blocks.ctypes.data, bp.value, blocks.nbytes)
        if nc.value:
            C.memmove(ckpts.ctypes.data, cp.value, ckpts.nbytes)
        return blocks, ckpts

    def plan(self, blocks, ckpts):
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DT)
        ckpts = np.ascontiguousarray(ckpts, dtype=CKPT_DT)
        total, pieces = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sdz_large_plan(self.h, blocks.ctypes.data, blocks.size, ckpts.ctypes.data, ckpts.size,
                                            C.byref(total), C.byref(pieces)))
        return total.value, pieces.value

    def range(self, part, n_parts):
        lo, hi = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sdz_large_range(self.h, part, n_parts, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def decode(self, part, n_parts, out_ptr):
        self._check(self.lib.sdz_large_decode(self.h, part, n_parts, out_ptr))

    def windows(self):
        self._check(self.lib.sdz_large_windows(self.h))

    def resolve(self):
        self._check(self.lib.sdz_large_resolve(self.h))

    def is_gzip(self):
        return bool(self.lib.sdz_large_is_gzip(self.h))

    def crc32(self, ptr, n):
        out = C.c_int32()
        self.ctx.check(self.lib.sdz_crc32(self.ctx.h, ptr, int(n), 0, 1, C.byref(out)))
        return out.value

    def adler32_std(self, ptr, n):
        """STANDARD adler32 (seed 1) of n bytes in HBM.  The reference's function differs from it only for a single
        call whose length is a non-zero multiple of 5552, so such a slice is checksummed as two calls."""
        if n == 0:
            return 1
        segs = np.array([n - 1, 1] if n % 5552 == 0 else [n], dtype=np.uint64)
        vals = np.zeros(len(segs), dtype=np.int32)
        last = C.c_int32()
        self.ctx.check(self.lib.sdz_adler32_chain(self.ctx.h, ptr, segs.ctypes.data, len(segs), 1, 1, vals.ctypes.data, C.byref(last)))
        return last.value

    def adler32_ref(self, data, seed):
        """the reference's adler32(chunk, seed) of a small host buffer (one call: Q1 applies)"""
        buf = np.ascontiguousarray(data, dtype=np.uint8)
        out = C.c_int32()
        self.ctx.check(self.lib.sdz_adler32(self.ctx.h, buf.ctypes.data if buf.size else None, int(buf.size), int(seed), 0, C.byref(out)))
        return out.value

    def read(self, ptr, n):
        buf = np.empty(max(n, 1), dtype=np.uint8)
        if n:
            self.ctx.check(self.lib.sdz_memcpy_d2h(self.ctx.h, buf.ctypes.data, ptr, int(n)))
        return buf[:n]

    def finish(self, running):
        r = N.Result()
        self._check(self.lib.sdz_large_finish(self.h, int(running), C.byref(r)))
        return r

    def close(self):
        if getattr(self, "h", None):
            self.lib.sdz_large_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_index(parts):
    """parts: [(blocks, ckpts)] of every rank -> the concatenated index plan() takes."""
    blocks = np.concatenate([np.asarray(b, dtype=BLOCK_DT) for b, _ in parts]) if parts else np.empty(0, BLOCK_DT)
    ckpts = np.concatenate([np.asarray(c, dtype=CKPT_DT) for _, c in parts]) if parts else np.empty(0, CKPT_DT)
    return blocks, ckpts


def combine_crcs(parts):
    """parts: [(crc32 of the slice, slice length)] in rank order -> crc32 of the whole output (0 for no output)."""
    crc = 0
    for c, n in parts:
        if n:
            crc = crc32_combine(crc, c, n)
    return crc


def run_rank(backend, comm, rank, world, alloc):
    """The protocol above for one rank.  `alloc(nbytes)` returns (object to keep alive, device/host pointer) for
    the rank's output slice with WIN bytes of headroom BEFORE and 64 bytes after it.
    Returns (keepalive, pointer to the slice, lo, hi, finish() record)."""
    mine = backend.index(rank, world)
    everyone = comm.allgather_index(mine)
    total, _ = backend.plan(*merge_index(everyone))
    lo, hi = backend.range(rank, world)
    keep, base = alloc(WIN + (hi - lo) + 64)
    out_ptr = base + WIN
    backend.decode(rank, world, out_ptr)
    if rank > 0:
        comm.recv_window(rank - 1, keep, base)              # lands in the headroom: the 32 KiB before `lo`
    backend.windows()
    if rank + 1 < world:
        # the last 32 KiB before `hi`; a slice shorter than the window forwards part of what it received
        comm.send_window(rank + 1, keep, out_ptr + (hi - lo) - WIN)
    backend.resolve()
    if backend.is_gzip():
        crc = backend.crc32(out_ptr, hi - lo) if hi > lo else 0
        parts = comm.allgather_crc((crc, hi - lo))
        rec = backend.finish(combine_crcs(parts))
    else:
        rec = backend.finish(running_adler32(backend, comm, out_ptr, lo, hi, total))
    return keep, out_ptr, lo, hi, rec


def running_adler32(backend, comm, out_ptr, lo, hi, total):
    """Inflater.checksum of a zlib / raw stream whose output is sharded over the ranks.  append() chains
    adler32 over its <= 16 KiB chunks (src/sd-inflate.ts:133-149); every chunk but the last is exactly 16,384 bytes, where
    the reference's adler32 IS the standard one, so all bytes before the final chunk combine associatively: each rank
    sends (standard adler32 of its part, length), joined with adler32_combine in rank order.  Only the final chunk
    (total mod 16,384 bytes, or 16,384) can hit the reference's defect (a call of 5552 or 11104 bytes, Q1): its bytes
    (<= 16 KiB, possibly from two ranks) are gathered and checksummed as ONE reference call seeded with the rest."""
    if total == 0:
        return 0                                            # no chunk was ever produced
    last_len = total % OUTBUF or OUTBUF
    tail_lo = total - last_len
    pre_n = max(0, min(hi, tail_lo) - lo)                   # my bytes before the final chunk
    mine = backend.adler32_std(out_ptr, pre_n)
    seed = 1
    for a, n in comm.allgather_crc((mine, pre_n)):
        if n:
            seed = adler32_combine(seed, a, n)
    t_lo, t_hi = max(lo, tail_lo), hi                       # my part of the final chunk
    piece = backend.read(out_ptr + (t_lo - lo), t_hi - t_lo) if t_hi > t_lo else np.empty(0, dtype=np.uint8)
    chunk = np.concatenate(comm.allgather_bytes(piece))
    assert chunk.size == last_len
    return backend.adler32_ref(chunk, seed)


class TorchComm:
    """torch.distributed plumbing: NCCL with CUDA tensors, gloo with CPU tensors."""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device, self.group = torch, dist, device, group

    def _allgather_bytes(self, payload):
        torch, dist = self.torch, self.dist
        world = dist.get_world_size(self.group)
        n = torch.tensor([payload.size], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        cap = max(max(sizes), 1)
        mine = torch.zeros(cap, dtype=torch.uint8, device=self.device)
        if payload.size:
            mine[:payload.size] = torch.from_numpy(payload).to(self.device)
        bufs = [torch.empty(cap, dtype=torch.uint8, device=self.device) for _ in range(world)]
        dist.all_gather(bufs, mine, group=self.group)
        return [b[:s].cpu().numpy() for b, s in zip(bufs, sizes)]

    def allgather_bytes(self, payload):
        return self._allgather_bytes(np.ascontiguousarray(payload, dtype=np.uint8))

    def allgather_index(self, mine):
        blocks, ckpts = mine
        head = np.array([blocks.size, ckpts.size], dtype=np.uint64)
        payload = np.frombuffer(head.tobytes() + blocks.tobytes() + ckpts.tobytes(), dtype=np.uint8)      # one exchange
        out = []
        for g in self._allgather_bytes(payload):
            raw = g.tobytes()
            nb, nc = (int(v) for v in np.frombuffer(raw[:16], dtype=np.uint64))
            b_end = 16 + nb * BLOCK_DT.itemsize
            out.append((np.frombuffer(raw[16:b_end], dtype=BLOCK_DT), np.frombuffer(raw[b_end:b_end + nc * CKPT_DT.itemsize], dtype=CKPT_DT)))
        return out

    def allgather_crc(self, mine):
        arr = np.array([mine[0], mine[1]], dtype=np.int64)
        got = self._allgather_bytes(np.frombuffer(arr.tobytes(), dtype=np.uint8))
        return [tuple(int(v) for v in np.frombuffer(g.tobytes(), dtype=np.int64)) for g in got]

    def _window(self, keep, ptr):
        # `keep` is the torch uint8 tensor that owns the rank's buffer; ptr is an address inside it
        off = ptr - keep.data_ptr()
        return keep[off:off + WIN]

    def recv_window(self, src, keep, ptr):
        w = self._window(keep, ptr)
        self.dist.recv(w, src=src, group=self.group)
        if w.is_cuda:
            # NCCL only enqueues the receive; the next phase runs on the library's own stream
            self.torch.cuda.current_stream(w.device).synchronize()

    def send_window(self, dst, keep, ptr):
        w = self._window(keep, ptr)
        self.dist.send(w, dst=dst, group=self.group)
        if w.is_cuda:
            self.torch.cuda.current_stream(w.device).synchronize()


def torch_alloc(device):
    import torch

    def alloc(nbytes):
        t = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        return t, t.data_ptr()
    return alloc


def inflate_large_parts(view, n_parts, mode=MODE_SNIFF, ctx=None):
    """All `n_parts` ranks of the protocol on ONE GPU, one after the other (the windows travel through the host).
    Exercises every multi-part code path without a second device.  Returns (bytes as np.uint8, finish() record)."""
    ctx = ctx or N.default_context()
    lib = ctx.lib
    n = int(view.size)
    d_in = lib.sdz_device_alloc(ctx.h, n + 1024)
    pad = np.zeros(n + 1024, dtype=np.uint8)
    pad[:n] = view
    ctx.check(lib.sdz_memcpy_h2d(ctx.h, d_in, pad.ctypes.data, pad.size))
    ranks, bufs = [], []
    try:
        ranks = [CudaBackend(ctx, d_in, n, mode) for _ in range(n_parts)]
        gz = ranks[0].is_gzip()
        merged = merge_index([r.index(p, n_parts) for p, r in enumerate(ranks)])
        total = 0
        for r in ranks:
            total, _ = r.plan(*merged)
        out = np.empty(total, dtype=np.uint8)
        window = np.zeros(WIN, dtype=np.uint8)
        crcs = []
        for p, r in enumerate(ranks):
            lo, hi = r.range(p, n_parts)
            base = lib.sdz_device_alloc(ctx.h, WIN + (hi - lo) + 64)
            bufs.append(base)
            r.decode(p, n_parts, base + WIN)
            if p > 0:
                ctx.check(lib.sdz_memcpy_h2d(ctx.h, base, window.ctypes.data, WIN))
            r.windows()
            ctx.check(lib.sdz_memcpy_d2h(ctx.h, window.ctypes.data, base + (hi - lo), WIN))    # [hi - WIN, hi) of this slice
            r.resolve()
            if hi > lo:
                ctx.check(lib.sdz_memcpy_d2h(ctx.h, out[lo:hi].ctypes.data, base + WIN, hi - lo))
            if gz:
                crcs.append((r.crc32(base + WIN, hi - lo) if hi > lo else 0, hi - lo))
        if gz:
            rec = ranks[0].finish(combine_crcs(crcs))
        else:
            # the same exchange as running_adler32(), all ranks played by this process
            class _Loop:
                def __init__(self):
                    self.sums, self.pieces = [], []
            loop = _Loop()
            last_len = (total % OUTBUF or OUTBUF) if total else 0
            tail_lo = total - last_len
            for p, r in enumerate(ranks):
                lo, hi = r.range(p, n_parts)
                pre_n = max(0, min(hi, tail_lo) - lo)
                loop.sums.append((r.adler32_std(bufs[p] + WIN, pre_n), pre_n))
                t_lo = max(lo, tail_lo)
                loop.pieces.append(r.read(bufs[p] + WIN + (t_lo - lo), hi - t_lo) if hi > t_lo else np.empty(0, dtype=np.uint8))
            seed = 1
            for a, n_ in loop.sums:
                if n_:
                    seed = adler32_combine(seed, a, n_)
            rec = ranks[0].finish(ranks[0].adler32_ref(np.concatenate(loop.pieces), seed) if total else 0)
        return out, rec
    finally:
        for r in ranks:
            r.close()
        for b in bufs:
            lib.sdz_device_free(ctx.h, b)
        lib.sdz_device_free(ctx.h, d_in)
