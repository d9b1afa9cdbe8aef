"""CPU test of the N > 1 path (gloo, world size 2): the per-stream sharding used by bench.py and
the record gather (the only exchange between ranks - payload bytes never cross ranks).  The
decode itself is stood in for by the CPU oracle here; on GPUs each rank calls libsdzcuda.so."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(n_streams, rank, world):
    """contiguous per-rank slice of the stream list (bench.py: every rank owns its streams end to end)"""
    lo = n_streams * rank // world
    hi = n_streams * (rank + 1) // world
    return lo, hi


def _worker(rank, world, port, n_streams, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tools import corpus as K
    lo, hi = shard(n_streams, rank, world)
    comp, stride, clen, plain = K.make_batch(K.TEXT, hi - lo, 4096, 6, K.ZLIB, first_index=lo, keep_plain=True, threads=1)
    rec_size = C.sizeof(O.Result)
    mine = np.zeros((hi - lo) * rec_size, dtype=np.uint8)
    for i in range(hi - lo):
        out, r = O.inflate_oneshot(comp[i * stride:i * stride + int(clen[i])].tobytes())
        assert out == plain[i * 4096:(i + 1) * 4096].tobytes()
        mine[i * rec_size:(i + 1) * rec_size] = np.frombuffer(bytes(r), dtype=np.uint8)
    # K8: gather fixed-size records (ranks may own different counts: pad to the maximum)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    mx = int(max(c.item() for c in cnts))
    padded = torch.zeros(mx * rec_size, dtype=torch.uint8)
    padded[:mine.size] = torch.from_numpy(mine)
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    allrec = np.concatenate([p.numpy()[:int(c.item()) * rec_size] for p, c in zip(parts, cnts)])
    # timing convention of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out_q.put((allrec.tobytes(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_record_gather():
    from oracle import oracle as O
    world, n_streams = 2, 9                       # odd on purpose: ranks own 4 and 5 streams
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    raw, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec_size = C.sizeof(O.Result)
    recs = (O.Result * n_streams).from_buffer_copy(raw)
    assert len(raw) == n_streams * rec_size and tmax == 2.0
    assert all(r.success and r.out_len == 4096 and r.checksum_state == 1 for r in recs)
    # shards tile the stream list exactly once
    covered = []
    for r in range(world):
        lo, hi = shard(n_streams, r, world)
        covered += list(range(lo, hi))
    assert covered == list(range(n_streams))


# ---------------------------------------------------------------- one large stream over several ranks

class _StandInBackend:
    """CPU stand-in for sdzlib.large.CudaBackend: same methods, numpy instead of kernels.  Every rank 'finds' a few
    records in its slice of the compressed bits; decode() refuses to produce the right bytes unless the protocol
    delivered the full index, and windows() checks that the 32 KiB before the slice arrived from the left."""

    def __init__(self, plain, rank, world, gz=True):
        self.plain, self.rank, self.world, self.gz = plain, rank, world, gz

    def is_gzip(self):
        return self.gz

    # zlib / raw streams: the three local operations of sdzlib.large.running_adler32()
    def adler32_std(self, ptr, n):
        import zlib
        buf = np.empty(max(n, 1), dtype=np.uint8)
        C.memmove(buf.ctypes.data, ptr, n)
        v = zlib.adler32(buf[:n].tobytes())
        return v - (1 << 32) if v & 0x80000000 else v

    def adler32_ref(self, data, seed):
        from oracle import oracle as O
        return O.adler32(np.ascontiguousarray(data, dtype=np.uint8).tobytes(), seed)

    def read(self, ptr, n):
        buf = np.empty(max(n, 1), dtype=np.uint8)
        C.memmove(buf.ctypes.data, ptr, n)
        return buf[:n]

    def index(self, part, n_parts):
        from sdzlib import large as LG
        b = np.zeros(3 + part, dtype=LG.BLOCK_DT)
        b["bit"] = 1000 * part + np.arange(b.size)
        b["ok"] = 1
        c = np.zeros(2 * part + 1, dtype=LG.CKPT_DT)
        c["block_bit"] = 1000 * part
        c["pos"] = np.arange(c.size)
        return b, c

    def plan(self, blocks, ckpts):
        exp_b = sorted(1000 * p + i for p in range(self.world) for i in range(3 + p))
        assert sorted(int(v) for v in blocks["bit"]) == exp_b
        assert ckpts.size == sum(2 * p + 1 for p in range(self.world))
        return self.plain.size, 0

    def range(self, part, n_parts):
        n = self.plain.size
        return n * part // n_parts, n * (part + 1) // n_parts

    def decode(self, part, n_parts, out_ptr):
        self.lo, self.hi = self.range(part, n_parts)
        self.out_ptr = out_ptr
        seg = np.ascontiguousarray(self.plain[self.lo:self.hi])
        C.memmove(out_ptr, seg.ctypes.data, seg.size)

    def windows(self):
        if self.rank > 0:
            from sdzlib import large as LG
            got = np.empty(LG.WIN, dtype=np.uint8)
            C.memmove(got.ctypes.data, self.out_ptr - LG.WIN, LG.WIN)
            w0 = self.lo - LG.WIN
            exp = self.plain[max(w0, 0):self.lo]
            assert np.array_equal(got[LG.WIN - exp.size:], exp), "left window did not arrive intact"

    def resolve(self):
        pass

    def crc32(self, ptr, n):
        import zlib
        buf = np.empty(n, dtype=np.uint8)
        C.memmove(buf.ctypes.data, ptr, n)
        v = zlib.crc32(buf.tobytes())
        return v - (1 << 32) if v & 0x80000000 else v

    def finish(self, running):
        return running


def _large_worker(rank, world, port, n_bytes, out_q, gz=True):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdzlib import large as LG
    plain = np.random.default_rng(5).integers(0, 256, n_bytes, dtype=np.uint8)
    comm = LG.TorchComm(torch.device("cpu"))
    keep, ptr, lo, hi, rec = LG.run_rank(_StandInBackend(plain, rank, world, gz), comm, rank, world, LG.torch_alloc(torch.device("cpu")))
    got = np.empty(hi - lo, dtype=np.uint8)
    C.memmove(got.ctypes.data, ptr, hi - lo)
    assert np.array_equal(got, plain[lo:hi])
    out_q.put((rank, lo, hi, int(rec)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_bytes", [(2, 300000), (3, 70000)])      # 70000 / 3 < 32 KiB: windows span two slices
def test_large_stream_rank_protocol(world, n_bytes):
    import zlib
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_large_worker, args=(r, world, port, n_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plain = np.random.default_rng(5).integers(0, 256, n_bytes, dtype=np.uint8)
    exp = zlib.crc32(plain.tobytes())
    exp = exp - (1 << 32) if exp & 0x80000000 else exp
    assert [g[1] for g in got] == [n_bytes * r // world for r in range(world)]
    assert got[-1][2] == n_bytes
    assert all(g[3] == exp for g in got)                                  # every rank assembled the whole-stream CRC


@pytest.mark.parametrize("world,n_bytes", [(2, 16384 * 9 + 5552), (3, 16384 + 11104), (3, 16384 * 4)])
def test_large_stream_adler_over_ranks(world, n_bytes):
    """zlib / raw streams: every rank ends up with the reference's chunk-chained Adler-32 (Q1 on a final chunk of 5552 or
    11104 bytes), joined from per-rank (standard adler32, length) pairs + the gathered final chunk.  (3, 27488): the
    final chunk straddles two ranks."""
    from oracle import oracle as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_large_worker, args=(r, world, port, n_bytes, q, False)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plain = np.random.default_rng(5).integers(0, 256, n_bytes, dtype=np.uint8).tobytes()
    exp = 1
    for o in range(0, n_bytes, 16384):                              # what append() computes (src/sd-inflate.ts:133-149)
        exp = O.adler32(plain[o:o + 16384], exp)
    assert all(g[3] == exp for g in got)
    if n_bytes % 16384 in (5552, 11104):
        import zlib
        std = zlib.adler32(plain)
        assert exp != (std - (1 << 32) if std & 0x80000000 else std)      # the defect is really in play
