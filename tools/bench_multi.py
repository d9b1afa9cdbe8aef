#!/usr/bin/env python
"""ONE batch through ONE sdz_inflate_batch call on a multi-device context (sdz_ctx_create_multi): the library partitions the
batch per stream over every visible B200 (contiguous ranges balanced on compressed bytes), one host thread and one
copy / compute pipeline per device, records in caller order.  This is the form north_star describes for inflateBatch() on an
8-GPU box (single process, no NCCL).  Prints one JSON line: end-to-end GB/s with HOST buffers (pinned, one arena), the
partition, per-device shares, and the same batch on one device for comparison.

    python tools/bench_multi.py [--streams-per-gpu 65536] [--devices 0,1,2,3,4,5,6,7]
"""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))
import numpy as np
import torch
import bench as B
from sdzlib import _native as N
from oracle import oracle as O


def run(ctx, ins, n, h_out, o_off, o_cap, hres, steps):
    times = []
    for it in range(1 + steps):
        t0 = time.perf_counter()
        ctx.check(ctx.lib.sdz_inflate_batch(ctx.h, ins, n, h_out, o_off.ctypes.data, o_cap.ctypes.data, hres, 0))
        if it:
            times.append(time.perf_counter() - t0)
    return 1000.0 * min(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams-per-gpu", type=int, default=65536)
    ap.add_argument("--devices", default="")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    devs = [int(x) for x in a.devices.split(",")] if a.devices else list(range(torch.cuda.device_count()))
    nd = len(devs)
    n1 = a.streams_per_gpu
    cores = os.cpu_count() or 1
    comp, stride, clen, gen_s = B.make_corpus(max(1, n1 // 8), 0, cores)
    arena, off, ln = B.pack(comp, stride, clen, 8 * nd if n1 >= 8 else nd)
    n = len(ln)
    lib = N.load()
    h_in = lib.sdz_host_alloc(arena.size)
    out_bytes = n * B.STREAM_BYTES
    h_out = lib.sdz_host_alloc(out_bytes)
    C.memmove(h_in, arena.ctypes.data, arena.size)
    ins = (N.In * n)()
    for i in range(n):
        ins[i].data = h_in + int(off[i]); ins[i].len = int(ln[i]); ins[i].mode = 0
    o_off = np.arange(n, dtype=np.uint64) * np.uint64(B.STREAM_BYTES)
    o_cap = np.full(n, B.STREAM_BYTES, dtype=np.uint64)
    hres = (N.Result * n)()
    mctx = N.Context(devs)
    ms_multi = run(mctx, ins, n, h_out, o_off, o_cap, hres, a.steps)
    cut = mctx.last_partition()
    ok = all(hres[i].success and hres[i].out_len == B.STREAM_BYTES for i in range(0, n, 53))
    for i in (0, n // 2, n - 1):
        exp, er = O.inflate_oneshot(arena[int(off[i]):int(off[i]) + int(ln[i])].tobytes())
        got = bytes((C.c_uint8 * B.STREAM_BYTES).from_address(h_out + i * B.STREAM_BYTES))
        ok = ok and got == exp and er.observable() == hres[i].observable()
    mctx.close()
    # the same batch on one device (one pipeline)
    sctx = N.Context(devs[0])
    ms_one = run(sctx, ins, n, h_out, o_off, o_cap, hres, 1) if nd > 1 else ms_multi
    print(json.dumps({
        "metric": "inflateBatch end to end, ONE call on a multi-device context (host buffers)", "unit": "GB/s",
        "value": round(out_bytes / (ms_multi / 1000.0) / 1e9, 2), "ms": round(ms_multi, 2), "n_gpus": nd, "devices": devs,
        "streams": n, "out_bytes": out_bytes, "compressed_bytes": int(ln.astype(np.uint64).sum()),
        "partition": cut, "shares_compressed_bytes": [int(ln[cut[d]:cut[d + 1]].astype(np.uint64).sum()) for d in range(nd)],
        "one_device_same_batch": {"GB/s": round(out_bytes / (ms_one / 1000.0) / 1e9, 2), "ms": round(ms_one, 2)},
        "parity_ok": bool(ok), "timing": "best of %d calls, host clock around sdz_inflate_batch" % a.steps,
        "numa_node_first_device": int(lib.sdz_ctx_numa_node(sctx.h)), "host_cores": cores}))


if __name__ == "__main__":
    main()
