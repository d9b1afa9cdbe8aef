#!/usr/bin/env python
"""Times the general decoder (SDZ_FAST=0) on batches of one corpus kind, device arm (used to find out what the hand-over run
of the mixed batch spends its time on).  usage: tools/probe_stored.py [n_streams]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))
os.environ["SDZ_FAST"] = os.environ.get("SDZ_FAST", "0")
import numpy as np
import torch
from tools import corpus as K
from sdzlib import _native as N
import bench as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = N.Context(0)
for kind, plen, level, name in ((K.RANDOM, 40000, 6, "random40k L6"), (K.RANDOM, 10000, 6, "random10k L6"), (K.RANDOM, 40000, 0, "random40k L0"),
                                (K.TEXT, 40000, 6, "text40k L6"), (K.TINY, 150, 6, "tiny150 L6")):
    comp, stride, clen, _ = K.make_batch(kind, 256, plen, level, K.ZLIB, first_index=90000, threads=8)
    reps = n // 256
    ln = np.tile(clen.astype(np.uint32), reps)
    al = (ln.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(al[:-1])
    arena = np.zeros(int(al.sum()) + 1024, dtype=np.uint8)
    for i in range(n):
        j = i % 256
        arena[int(off[i]):int(off[i]) + int(ln[i])] = comp[j * stride:j * stride + int(clen[j])]
    db = B.DeviceBatch(torch, N, arena, off, ln, cap=np.full(n, (plen + 15) & ~15, dtype=np.uint32))
    best = 1e9
    for _ in range(4):
        ctx.check(ctx.lib.sdz_inflate_batch_device(ctx.h, C.byref(db.b), 0, 1))
        t = ctx.last_timing()
        best = min(best, t[0])
    recs = db.records(N)
    ok = all(recs[i].success and recs[i].out_len == plen for i in range(0, n, 97))
    print("%-14s %6d streams  decode kernels %.3f ms  ok=%s" % (name, n, best, ok), flush=True)
    del db
