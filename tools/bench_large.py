#!/usr/bin/env python
"""BASELINE config 5: ONE large gzip stream, decoded block-parallel (two-pass: index, then marker decode with
window resolution) on 1 GPU or spread over N GPUs.

    python tools/bench_large.py --mib 1024 --steps 3 --warmup 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_large.py --gpus 2 --mib 1024

Prints one JSON line on rank 0: decompressed GB/s with the stream resident in HBM (`value`; N > 1: every rank holds
the compressed stream, the output stays sharded over the GPUs) and, at N = 1, through the host-pointer call
(`e2e`), next to the sequential CPU decode of a prefix of the same plaintext by the oracle port (test
infrastructure, timed only as the baseline; a single stream cannot use more than one core in the reference
either).  The plaintext is --mib MiB of the synthetic text corpus, compressed here by zlib level 6 as ONE gzip
member (parallel chunk compression cannot be used: it would add flush points).  Scaling is STRONG: the stream is
the same at every N.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))

from tools import corpus as K  # noqa: E402
import sdzlib  # noqa: E402
from sdzlib import _native as N  # noqa: E402
from sdzlib import large as LG  # noqa: E402


def make_stream(mib, level):
    cache = os.environ.get("SDZ_STREAM_CACHE")              # reuse the stream between invocations on the same box
    path = "%s.%d.%d" % (cache, mib, level) if cache else None
    if path and os.path.exists(path + ".crc"):
        return open(path, "rb").read(), int(open(path + ".crc").read())
    stream, crc = _make_stream(mib, level)
    if path and int(os.environ.get("RANK", "0")) == 0:
        open(path, "wb").write(stream)
        open(path + ".crc", "w").write(str(crc))
    return stream, crc


def _make_stream(mib, level):
    co = zlib.compressobj(level, zlib.DEFLATED, 31)
    parts, crc = [], 0
    for i in range(mib * 16):
        p = K.generate(K.TEXT, 500000 + i, 65536).tobytes()
        crc = zlib.crc32(p, crc)
        parts.append(co.compress(p))
    parts.append(co.flush())
    return b"".join(parts), crc


def cpu_baseline(a):
    from oracle import oracle as O
    cm = min(a.cpu_mib, a.mib)
    co = zlib.compressobj(a.level, zlib.DEFLATED, 31)
    sp = b"".join(K.generate(K.TEXT, 500000 + i, 65536).tobytes() for i in range(cm * 16))
    ss = np.frombuffer(co.compress(sp) + co.flush(), dtype=np.uint8)
    ob = np.empty(len(sp) + 64, dtype=np.uint8)
    t = time.perf_counter()
    orec = O.inflate_oneshot_np(ss, ob)
    dt = time.perf_counter() - t
    assert orec.success and orec.out_len == len(sp)
    return {"value": round(len(sp) / dt / 1e9, 3), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "first %d MiB of the same plaintext as one gzip stream, oracle C port" % cm}


def single_gpu(a, stream, crc, n_out, ctx=None):
    ctx = ctx or sdzlib.default_context()
    lib = ctx.lib
    h_in = lib.sdz_host_alloc(len(stream) + 1024)
    h_out = lib.sdz_host_alloc(n_out + 64)
    C.memmove(h_in, stream, len(stream))
    res = N.Result()
    times = []
    for it in range(a.warmup + a.steps):
        t = time.perf_counter()
        rc = lib.sdz_inflate_large(ctx.h, h_in, len(stream), 0, 0, h_out, n_out, C.byref(res))
        dt = time.perf_counter() - t
        assert rc == 0, rc
        if it >= a.warmup:
            times.append(dt)
    assert res.success and res.out_len == n_out, res.observable()
    got = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint8)), shape=(n_out,))
    assert zlib.crc32(got) == crc, "decoded bytes differ from the plaintext"
    d_in = lib.sdz_device_alloc(ctx.h, len(stream) + 1024)
    d_out = lib.sdz_device_alloc(ctx.h, n_out + 64)
    ctx.check(lib.sdz_memcpy_h2d(ctx.h, d_in, h_in, len(stream) + 1024))
    dtimes = []
    launches0 = ctx.launch_count()
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            launches0 = ctx.launch_count()
        t = time.perf_counter()
        rc = lib.sdz_inflate_large(ctx.h, d_in, len(stream), 0, 1, d_out, n_out, C.byref(res))
        dt = time.perf_counter() - t
        assert rc == 0, rc
        if it >= a.warmup:
            dtimes.append(dt)
    assert res.success and res.out_len == n_out
    lib.sdz_device_free(ctx.h, d_in); lib.sdz_device_free(ctx.h, d_out)
    lib.sdz_host_free(h_in); lib.sdz_host_free(h_out)
    return {"value": round(n_out / min(dtimes) / 1e9, 2), "ms_per_step": round(min(dtimes) * 1e3, 2),
            "e2e": {"value": round(n_out / min(times) / 1e9, 2), "unit": "GB/s", "ms_per_step": round(min(times) * 1e3, 2),
                    "h2d_bytes_per_step": len(stream), "d2h_bytes_per_step": n_out},
            "blocks": int(res.n_blocks), "gpu_launches": (ctx.launch_count() - launches0) // a.steps}


def multi_gpu(a, stream, crc, n_out, ctx=None):
    """every rank calls this inside an initialised NCCL process group; rank 0 gets the result dict"""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    ctx = ctx or sdzlib.default_context(local)
    comm = LG.TorchComm(dev)
    d_in = torch.zeros(len(stream) + 1024, dtype=torch.uint8, device=dev)
    d_in[:len(stream)] = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8)).to(dev)
    alloc = LG.torch_alloc(dev)
    times = []
    rec = keep = None
    launches0 = 0
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            launches0 = ctx.launch_count()
        keep = None
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        be = LG.CudaBackend(ctx, d_in.data_ptr(), len(stream), 0)
        keep, ptr, lo, hi, rec = LG.run_rank(be, comm, rank, world, alloc)
        be.close()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)                 # a step ends when the slowest rank is done
        if it >= a.warmup:
            times.append(float(dt.item()))
    assert rec.success and rec.out_len == n_out, rec.observable()
    # the slices tile the output and hold the plaintext: CRC of each slice, combined on rank 0, equals zlib's
    off = ptr - keep.data_ptr()
    mine = zlib.crc32(keep[off:off + (hi - lo)].cpu().numpy().tobytes())
    parts = comm.allgather_crc((mine, hi - lo))
    joined = LG.combine_crcs([(c - (1 << 32) if c & 0x80000000 else c, n) for c, n in parts]) & 0xFFFFFFFF
    assert joined == crc, "decoded bytes differ from the plaintext"
    launches = torch.tensor([ctx.launch_count() - launches0], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    out = {"value": round(n_out / min(times) / 1e9, 2), "ms_per_step": round(min(times) * 1e3, 2), "e2e": None,
           "blocks": int(rec.n_blocks), "gpu_launches": int(launches.item()) // a.steps,
           "slices": [n for _, n in parts]}
    dist.barrier()
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--cpu-mib", type=int, default=32, help="prefix of the plaintext the CPU baseline decodes")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus != world:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (see the module docstring)" % (a.gpus, a.gpus))
    t0 = time.time()
    stream, crc = make_stream(a.mib, a.level)
    gen_s = time.time() - t0
    n_out = a.mib << 20
    if world == 1:
        r = single_gpu(a, stream, crc, n_out)
    else:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        r = multi_gpu(a, stream, crc, n_out)
        dist.destroy_process_group()
    if r is None:
        return
    line = {
        "metric": "single-stream inflate decompressed GB/s", "value": r["value"], "unit": "GB/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "one %d MiB synthetic-text gzip stream, zlib level %d" % (a.mib, a.level),
                   "compressed_bytes": len(stream), "blocks": r["blocks"], "corpus_gen_s": round(gen_s, 1),
                   "timing": "best of %d steps, host clock around the whole call (all phases, host syncs included)" % a.steps},
        "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
        "cpu_baseline": cpu_baseline(a),
    }
    if "slices" in r:
        line["config"]["output_slices"] = r["slices"]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
