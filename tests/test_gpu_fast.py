"""GPU tests of the two-phase fast path (fast_kernels.cuh): the same batches through the fast path (phase A tokens ->
phase B bytes -> hand-over of everything irregular) and through the general decoder alone must give identical bytes
and records, both identical to the oracle's; clean text must really be finished by the fast path."""
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)
from tools import corpus as K  # noqa: E402

from sdzlib import _native as N  # noqa: E402
from sdzlib import api as A  # noqa: E402


def _ctx(fast):
    old = os.environ.get("SDZ_FAST")
    os.environ["SDZ_FAST"] = "1" if fast else "0"
    try:
        return N.Context(0)
    finally:
        if old is None:
            del os.environ["SDZ_FAST"]
        else:
            os.environ["SDZ_FAST"] = old


@pytest.fixture(scope="module")
def both():
    f, g = _ctx(True), _ctx(False)
    yield f, g
    f.close()
    g.close()


def _run(ctx, streams, modes=None):
    views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
    arena, off, res = A.inflate_batch_raw(views, None, modes, None, ctx)
    return [(bytes(arena[int(off[i]):int(off[i]) + int(res[i].out_len)]), res[i]) for i in range(len(views))]


def _mixed(n_each=6):
    streams = []
    for seed in range(n_each):
        for kind, n in ((K.TEXT, 65536), (K.TEXT, 3000 + 977 * seed), (K.BINARY, 30000), (K.RANDOM, 20000), (K.TINY, 100), (K.RUNS, 50000)):
            plain = K.generate(kind, 100 + seed, n)
            for cont in (K.ZLIB, K.GZIP_NAME, K.RAW):
                for level in (1, 6, 9):
                    streams.append(K.compress(plain, level, cont))
    return streams


def test_fast_equals_general_equals_oracle(both):
    f, g = both
    streams = _mixed()
    # damaged copies: truncated, trailing byte, flipped bit -> all of these must be handed over and come out identical
    extra = []
    for i, s in enumerate(streams[:60]):
        extra.append(s[:len(s) - 1 - (i % 7)])
        extra.append(s + b"\x00")
        b = bytearray(s)
        b[len(b) // 2] ^= 1 << (i % 8)
        extra.append(bytes(b))
    streams += extra
    rf, rg = _run(f, streams), _run(g, streams)
    bad = []
    for i, s in enumerate(streams):
        exp_bytes, exp = O.inflate_oneshot(bytes(s))
        (fb, fr), (gb, gr) = rf[i], rg[i]
        if fr.observable() != gr.observable() or (not fr.thrown_append and fb != gb):
            bad.append(("fast!=general", i, len(s), fr.observable(), gr.observable()))
        elif exp.thrown_append != 6 and (fr.observable() != exp.observable() or (not exp.thrown_append and fb != exp_bytes)):
            bad.append(("fast!=oracle", i, len(s), fr.observable(), exp.observable()))
    assert not bad, "%d of %d differ: %s" % (len(bad), len(streams), bad[:6])


def test_clean_text_stays_on_the_fast_path(both):
    f, _ = both
    streams = [K.compress(K.generate(K.TEXT, 500 + i, 65536), 6, K.ZLIB) for i in range(256)]
    got = _run(f, streams)
    done, handed = f.last_fast_stats()
    assert (done, handed) == (256, 0)
    for i, s in enumerate(streams):
        assert got[i][0] == zlib.decompress(s) and got[i][1].success == 1 and got[i][1].checksum_state == 1


def test_handover_classes(both):
    f, _ = both
    text = K.generate(K.TEXT, 9, 20000)
    stored = K.compress(K.generate(K.RANDOM, 9, 20000), 6, K.ZLIB)          # stored blocks
    trunc = K.compress(text, 6, K.ZLIB)[:-2]                                  # trailer incomplete
    trail = K.compress(text, 6, K.GZIP_NAME) + b"x"                           # byte after the end (Q4)
    early = bytes.fromhex("030200")                                           # D1: distance before the start (Q6)
    ok = K.compress(text, 9, K.GZIP_NAME)
    got = _run(f, [stored, trunc, trail, early, ok])
    done, handed = f.last_fast_stats()
    assert (done, handed) == (1, 4)
    for (b, r), s in zip(got, [stored, trunc, trail, early, ok]):
        eb, er = O.inflate_oneshot(s)
        assert r.observable() == er.observable()
        if not er.thrown_append:
            assert b == eb


def test_ragged_batch_fast(both):
    """stream sizes from a few bytes to 1 MiB in one batch: lanes of a warp finish at very different times"""
    f, g = both
    streams = []
    for i in range(96):
        n = [5, 70, 900, 4096, 65536, 300000, 1 << 20][i % 7] + i
        kind = [K.TEXT, K.BINARY, K.RUNS][i % 3]
        streams.append(K.compress(K.generate(kind, 40 + i, n), [1, 6, 9][i % 3], [K.ZLIB, K.GZIP_NAME, K.RAW][(i // 3) % 3]))
    rf, rg = _run(f, streams), _run(g, streams)
    for i, s in enumerate(streams):
        assert rf[i][1].observable() == rg[i][1].observable(), i
        assert rf[i][0] == rg[i][0], i
        if rf[i][1].container != 0:
            assert rf[i][1].success == 1, i


def _run_dict(ctx, streams, dicts):
    views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
    dv = [None if d is None else np.frombuffer(bytes(d), dtype=np.uint8) for d in dicts]
    arena, off, res = A.inflate_batch_raw(views, dv, [O.MODE_INFLATER] * len(views), None, ctx)
    return [(bytes(arena[int(off[i]):int(off[i]) + int(res[i].out_len)]), res[i]) for i in range(len(views))]


def test_preset_dictionary_streams_on_the_fast_path(both):
    """zlib streams with FDICT whose dictionary the caller supplied are finished by the two-phase path (matches that start
    in the dictionary, and matches that run from the dictionary into the output); a missing, wrong or >= 32 KiB dictionary
    goes to the general decoder.  Fast path == general decoder == oracle everywhere."""
    f, g = both
    streams, dicts = [], []
    for k, dl in enumerate((1, 2, 31, 470, 5552, 20000, 32767)):
        dic = bytes(K.generate(K.TEXT, 3000 + k, dl))
        did = O.adler32(dic)
        for j in range(4):
            # the plaintext starts with a piece of the dictionary's tail and goes on inside the dictionary's text: the first
            # matches reach into the dictionary, some of them across its end
            head = dic[-min(dl, 40 + 13 * j):] + dic[:min(dl, 300)]
            plain = head + bytes(K.generate(K.TEXT, 3100 + 10 * k + j, 3000 + 9000 * j))
            streams.append(K.compress(plain, (1, 6, 9, 6)[j], K.ZLIB_DICT, dic, did)); dicts.append(dic)
    n_fast = len(streams)
    dic = bytes(K.generate(K.TEXT, 3999, 470))
    s0 = K.compress(dic[-100:] + bytes(K.generate(K.TEXT, 4000, 5000)), 6, K.ZLIB_DICT, dic, O.adler32(dic))
    big = bytes(K.generate(K.TEXT, 4001, 32768))
    streams += [s0, s0, K.compress(big[-500:] + bytes(K.generate(K.TEXT, 4002, 5000)), 6, K.ZLIB_DICT, big, O.adler32(big))]
    dicts += [None, dic[:-1] + b"?", big]                                       # required / invalid / 32 KiB (Q14)
    rf, rg = _run_dict(f, streams, dicts), _run_dict(g, streams, dicts)
    done, handed = f.last_fast_stats()
    assert (done, handed) == (n_fast, 3)
    for i, s in enumerate(streams):
        eb, er = O.inflate_oneshot(bytes(s), dictionary=dicts[i], mode=O.MODE_INFLATER)
        assert rf[i][1].observable() == rg[i][1].observable() == er.observable(), i
        if not er.thrown_append:
            assert rf[i][0] == rg[i][0] == eb, i
    assert all(rf[i][1].success == 1 for i in range(n_fast))


def test_raw_streams_the_reference_leaves_incomplete_stay_on_the_fast_path(both):
    """the reference looks a code up only when its table's index width is available, so about one raw stream
    in 100 - 250 ends with the final end-of-block code undecodable (Z_OK, complete = false, all data delivered).  Phase A
    writes that record itself instead of handing the stream to the general decoder at the very end of its decode."""
    f, g = both
    streams = []
    for i in range(3000):
        kind = (K.TEXT, K.BINARY, K.RUNS)[i % 3]
        streams.append(K.compress(K.generate(kind, 20000 + i, 200 + (i * 37) % 900), (1, 6, 9)[i % 3], K.RAW))
    for i in range(48):
        streams.append(K.compress(K.generate((K.TEXT, K.BINARY)[i % 2], 30000 + i, 65536), (1, 6, 9)[i % 3], K.RAW))
    modes = [O.MODE_RAW] * len(streams)
    rf = _run(f, streams, modes)
    done, handed = f.last_fast_stats()
    rg = _run(g, streams, modes)
    n_incomplete = 0
    for i, s in enumerate(streams):
        eb, er = O.inflate_oneshot(bytes(s), mode=O.MODE_RAW)
        assert rf[i][1].observable() == rg[i][1].observable() == er.observable(), i
        assert rf[i][0] == rg[i][0] == eb, i
        n_incomplete += 0 if er.complete else 1
    assert n_incomplete >= 10                              # the case is really in the batch ...
    assert handed <= n_incomplete // 4, (done, handed)     # ... and (nearly) none of it needs the general decoder
