#!/usr/bin/env python
"""BASELINE config 5: ONE large gzip stream through sdz_inflate_large (block-parallel two-pass decode).

    python tools/bench_large.py --mib 1024 --steps 3 --warmup 1

Prints one JSON line: decompressed GB/s with host buffers (e2e) and the device phases, next to the
sequential CPU decode of the same stream by the oracle port (test infrastructure, timed only as the
baseline).  The plaintext is --mib MiB of the synthetic text corpus, compressed here by zlib level 6
as ONE gzip member (chunks compressed by parallel workers cannot be used: that would add flush points).
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
from sdzlib import api as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--cpu-mib", type=int, default=64, help="prefix of the plaintext the CPU baseline decodes")
    a = ap.parse_args()

    t0 = time.time()
    co = zlib.compressobj(a.level, zlib.DEFLATED, 31)
    parts = []
    crc = 0
    for i in range(a.mib * 16):
        p = K.generate(K.TEXT, 500000 + i, 65536).tobytes()
        crc = zlib.crc32(p, crc)
        parts.append(co.compress(p))
    parts.append(co.flush())
    stream = b"".join(parts)
    gen_s = time.time() - t0
    n_out = a.mib << 20

    ctx = sdzlib.default_context()
    lib = ctx.lib
    # pinned host buffers, as a production caller would hold them
    h_in = lib.sdz_host_alloc(len(stream) + 1024)
    h_out = lib.sdz_host_alloc(n_out + 64)
    C.memmove(h_in, stream, len(stream))
    res = N.Result()
    times, dev = [], []
    for it in range(a.warmup + a.steps):
        t = time.perf_counter()
        rc = lib.sdz_inflate_large(ctx.h, h_in, len(stream), 0, 0, h_out, n_out, C.byref(res))
        dt = time.perf_counter() - t
        assert rc == 0, rc
        if it >= a.warmup:
            times.append(dt)
            dev.append(ctx.last_timing())
    assert res.success and res.out_len == n_out, res.observable()
    got = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_uint8)), shape=(n_out,))
    assert zlib.crc32(got) == crc, "decoded bytes differ from the plaintext"

    # device-resident variant: input and output stay in HBM
    d_in = lib.sdz_device_alloc(ctx.h, len(stream) + 1024)
    d_out = lib.sdz_device_alloc(ctx.h, n_out + 64)
    ctx.check(lib.sdz_memcpy_h2d(ctx.h, d_in, h_in, len(stream) + 1024))
    dtimes = []
    for it in range(a.warmup + a.steps):
        t = time.perf_counter()
        rc = lib.sdz_inflate_large(ctx.h, d_in, len(stream), 0, 1, d_out, n_out, C.byref(res))
        dt = time.perf_counter() - t
        assert rc == 0, rc
        if it >= a.warmup:
            dtimes.append(dt)
    assert res.success and res.out_len == n_out

    # CPU baseline: the oracle port decoding a gzip stream of the first --cpu-mib MiB, one thread
    # (a single stream cannot use more than one core in the reference either)
    from oracle import oracle as O
    cm = min(a.cpu_mib, a.mib)
    sample = zlib.compressobj(a.level, zlib.DEFLATED, 31)
    sp = b"".join(K.generate(K.TEXT, 500000 + i, 65536).tobytes() for i in range(cm * 16))
    ss = sample.compress(sp) + sample.flush()
    ss_np = np.frombuffer(ss, dtype=np.uint8)
    ob = np.empty(len(sp) + 64, dtype=np.uint8)
    t = time.perf_counter()
    orec = O.inflate_oneshot_np(ss_np, ob)
    cpu_dt = time.perf_counter() - t
    assert orec.success and orec.out_len == len(sp)

    best = min(times)
    k = times.index(best)
    line = {
        "metric": "single-stream inflate decompressed GB/s", "unit": "GB/s", "n_gpus": 1,
        "value": round(n_out / min(dtimes) / 1e9, 2), "e2e": {"value": round(n_out / best / 1e9, 2), "unit": "GB/s",
                                                                 "h2d_bytes_per_step": len(stream), "d2h_bytes_per_step": n_out},
        "ms_per_step": round(min(dtimes) * 1e3, 2), "e2e_ms_per_step": round(best * 1e3, 2),
        "device_ms": {"index+decode": round(dev[k][0], 2), "windows+markers": round(dev[k][1], 2)},
        "config": {"workload": "one %d MiB synthetic-text gzip stream, zlib level %d" % (a.mib, a.level),
                   "compressed_bytes": len(stream), "blocks": int(res.n_blocks), "corpus_gen_s": round(gen_s, 1)},
        "cpu_baseline": {"value": round(len(sp) / cpu_dt / 1e9, 3), "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": "first %d MiB of the same plaintext as one gzip stream, oracle C port" % cm},
        "steps": a.steps, "warmup": a.warmup, "dtype": "u8", "data": "synthetic",
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
