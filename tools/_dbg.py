import sys, zlib, random
sys.path.insert(0,'.'); sys.path.insert(0,'sd-zlib_b200/host')
import numpy as np
from oracle import oracle as O
from tools import corpus as K
from sdzlib import api as A
dic = bytes(K.generate(K.TEXT, 4242, 470)); dictid = O.adler32(dic)
rnd = random.Random(44)
streams, dicts, modes, plains = [], [], [], []
kinds = [(K.TEXT, 65536), (K.BINARY, 65536), (K.TINY, 0), (K.RANDOM, 0), (K.RUNS, 65536), (K.TEXT, 20000)]
for i in range(1500):
    kind, n = kinds[i % len(kinds)]
    if kind == K.TINY: n = 1 + rnd.randrange(200)
    elif kind == K.RANDOM: n = 1 + rnd.randrange(49151)
    plain = K.generate(kind, 7000 + i, n)
    level = (1, 6, 9)[i % 3]
    cont = (K.GZIP, K.RAW, K.ZLIB, K.GZIP_NAME, K.ZLIB_DICT)[i % 5]
    plains.append(plain.tobytes())
    if cont == K.ZLIB_DICT:
        streams.append(K.compress(plain, level, cont, dic, dictid)); dicts.append(dic); modes.append(O.MODE_INFLATER)
    else:
        streams.append(K.compress(plain, level, cont)); dicts.append(None)
        modes.append(O.MODE_RAW if cont == K.RAW and i % 2 else O.MODE_SNIFF)
views = [np.frombuffer(s, dtype=np.uint8) for s in streams]
for trial in range(2):
    arena, off, res = A.inflate_batch_raw(views, dicts, modes, None)
    nbad = 0
    for i in range(1500):
        r = res[i]
        got = bytes(arena[int(off[i]):int(off[i]) + int(r.out_len)])
        if got != plains[i] and modes[i] != O.MODE_RAW and r.out_len == len(plains[i]):
            diffs = [k for k in range(len(got)) if got[k] != plains[i][k]]
            nbad += 1
            if nbad <= 6:
                k = diffs[0]
                print(trial, i, "ndiff", len(diffs), "first", k, "last", diffs[-1], "plain", plains[i][max(0,k-6):k+10].hex(), "got", got[max(0,k-6):k+10].hex(), "off%16", int(off[i]) % 16)
    print("trial", trial, "bad", nbad)
