#!/usr/bin/env python
"""Which classes of the cfg4 mixed batch does phase A hand to the general decoder?  One small batch per (kind, container),
fast-path statistics of each (GPU box only; diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))
import numpy as np
from tools import corpus as K
from sdzlib import api as A, _native as N
ctx = N.default_context()
kinds = [("text64k", K.TEXT, 65536), ("binary64k", K.BINARY, 65536), ("tiny150", K.TINY, 150), ("random40k", K.RANDOM, 40000),
         ("runs64k", K.RUNS, 65536), ("text20k", K.TEXT, 20000)]
conts = [("gzip", K.GZIP), ("zlib", K.ZLIB), ("raw", K.RAW)]
for kn, kind, plen in kinds:
    for cn, cont in conts:
        for level in (1, 6, 9):
            comp, stride, clen, _ = K.make_batch(kind, 256, plen, level, cont, first_index=5000)
            views = [comp[i * stride:i * stride + int(clen[i])] for i in range(256)]
            caps = np.full(256, (plen + 15) & ~15, dtype=np.uint64)
            A.inflate_batch_raw(views, None, [2 if cont == K.RAW else 0] * 256, caps, ctx=ctx)
            done, handed = ctx.last_fast_stats()
            print("%-10s %-5s L%d  fast %3d  handed %3d  (mean %d compressed bytes)" % (kn, cn, level, done, handed, int(clen.mean())), flush=True)
