#!/bin/bash
# compute-sanitizer over smoke() and a 256-stream mixed batch: memcheck, racecheck, synccheck, initcheck.
# Logs land in gpurun_out/<tag>_sanitize_*.log; the last lines of each carry the error summary.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/sdz_sanitize_case.py <<'PY'
import sys, os, random, zlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "sd-zlib_b200", "host"))
import numpy as np
import __graft_entry__ as G
G.smoke()
from tools import corpus as K
from oracle import oracle as O
from sdzlib import api as A
import sdzlib
dic = bytes(K.generate(K.TEXT, 4242, 470)); dictid = O.adler32(dic)
rnd = random.Random(44)
streams, dicts, modes = [], [], []
kinds = [(K.TEXT, 65536), (K.BINARY, 65536), (K.TINY, 0), (K.RANDOM, 0), (K.RUNS, 65536), (K.TEXT, 20000)]
for i in range(256):
    kind, n = kinds[i % len(kinds)]
    if kind == K.TINY: n = 1 + rnd.randrange(200)
    elif kind == K.RANDOM: n = 1 + rnd.randrange(49151)
    plain = K.generate(kind, 7000 + i, n)
    cont = (K.GZIP, K.RAW, K.ZLIB, K.GZIP_NAME, K.ZLIB_DICT)[i % 5]
    if cont == K.ZLIB_DICT:
        streams.append(K.compress(plain, (1, 6, 9)[i % 3], cont, dic, dictid)); dicts.append(dic); modes.append(1)
    else:
        streams.append(K.compress(plain, (1, 6, 9)[i % 3], cont)); dicts.append(None); modes.append(2 if cont == K.RAW and i % 2 else 0)
views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
arena, off, res = A.inflate_batch_raw(views, dicts, modes, None)
bad = 0
for i, s in enumerate(streams):
    eb, er = O.inflate_oneshot(bytes(s), dictionary=dicts[i], mode=modes[i])
    if er.observable() != res[i].observable() or (not er.thrown_append and bytes(arena[int(off[i]):int(off[i]) + int(res[i].out_len)]) != eb):
        bad += 1
# a streaming session and one large stream ride along
s = bytes(K.compress(K.generate(K.TEXT, 5, 90000), 6, K.GZIP_NAME))
inf = sdzlib.Inflater(); got = b"".join(b"".join(inf.append(s[i:i + 7001])) for i in range(0, len(s), 7001)); r = inf.finish()
bad += 0 if (got == zlib.decompress(s, 31) and r.success) else 1
big = zlib.compress(b"".join(K.generate(K.TEXT, 900 + i, 65536).tobytes() for i in range(48)), 6)
out = sdzlib.inflateLarge(big)
bad += 0 if out["data"] == zlib.decompress(big) else 1
print("sanitize case: %d streams, bad=%d" % (len(streams), bad))
sys.exit(1 if bad else 0)
PY
for tool in memcheck synccheck initcheck racecheck; do
    timeout 1500 $CS --tool $tool --print-limit 20 python /tmp/sdz_sanitize_case.py > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1
    echo "$tool rc=$?"; tail -4 gpurun_out/${TAG}_sanitize_${tool}.log
done
