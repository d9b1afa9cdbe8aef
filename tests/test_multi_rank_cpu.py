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
