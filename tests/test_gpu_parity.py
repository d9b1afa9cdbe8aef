"""GPU parity tests (run with -m gpu on a B200): every result of the CUDA path, taken
through the C ABI, is compared bit for bit with the CPU oracle on the same inputs, with the
reference's fixtures, and through size-independent properties at benchmark sizes."""
import gzip
import os
import random
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)
from tools import corpus as K  # noqa: E402

import sdzlib  # noqa: E402
from sdzlib import api as A  # noqa: E402


def raw_deflate(data, level=6, wbits=-15, zdict=None):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits) if zdict is None else zlib.compressobj(level, zlib.DEFLATED, wbits, 8, 0, zdict)
    return co.compress(data) + co.flush()


def run_batch(streams, dicts=None, modes=None, slack=64):
    """GPU result for each stream: (bytes, record)."""
    views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
    arena, off, res = A.inflate_batch_raw(views, dicts, modes, None)
    out = []
    for i in range(len(views)):
        r = res[i]
        out.append((bytes(arena[int(off[i]):int(off[i]) + int(r.out_len)]), r))
    return out


def check_against_oracle(streams, dicts=None, modes=None, allow_hang=True):
    got = run_batch(streams, dicts, modes)
    bad = []
    for i, s in enumerate(streams):
        d = None if dicts is None else dicts[i]
        m = O.MODE_SNIFF if modes is None else modes[i]
        exp_bytes, exp = O.inflate_oneshot(bytes(s), dictionary=d, mode=m)
        g_bytes, g = got[i]
        go, eo = g.observable(), exp.observable()
        if go != eo:
            diff = {k: (go.get(k), eo.get(k)) for k in set(go) | set(eo) if go.get(k) != eo.get(k)}
            bad.append((i, m, len(s), bytes(s)[:20].hex(), diff))
        elif not exp.thrown_append and g_bytes != exp_bytes:
            bad.append((i, m, len(s), "bytes differ", len(g_bytes), len(exp_bytes)))
    assert not bad, "%d of %d streams differ (got, expected): %s" % (len(bad), len(streams), bad[:12])
    return got


# ---------------------------------------------------------------- reference fixtures (Appendix B)

def test_fixtures_through_public_api(fx):
    txt = fx("paradiselost.txt")
    assert sdzlib.inflate(fx("paradiselost.deflate")) == txt
    assert sdzlib.inflate(fx("paradiselost.gz")) == txt
    assert sdzlib.inflate(fx("simple.deflate")) == fx("simple.txt")
    assert sdzlib.inflate(fx("simple.gz")) == fx("simple.txt")
    assert sdzlib.inflate(fx("simple.raw")) == fx("simple.txt")
    v = sdzlib.inflate(fx("vertices.deflate"))
    assert len(v) == 43440 and zlib.crc32(v) == 0xbf55c371


def test_fixture_records(fx):
    names = ["paradiselost.deflate", "paradiselost.gz", "simple.deflate", "simple.gz", "simple.raw", "vertices.deflate"]
    got = check_against_oracle([fx(n) for n in names])
    r = got[1][1]
    assert (r.success, r.complete, r.checksum_state, r.size_state) == (1, 1, 1, 1)
    assert r.mtime == 1530824734 and r.stored_checksum == -499006831
    assert fx("paradiselost.gz")[r.name_off:r.name_off + r.name_len] == b"paradiselost.txt"
    assert got[0][1].stored_checksum == -1949153550 and got[0][1].n_blocks == 7


def test_inflater_class(fx):
    inf = sdzlib.Inflater()
    b1 = inf.append(fx("paradiselost.part1.deflate"))
    b2 = inf.append(fx("paradiselost.part2.deflate"))
    assert sdzlib.mergeBuffers(b1 + b2) == fx("paradiselost.txt")
    assert all(len(c) <= 16384 for c in b1 + b2)
    res = inf.finish()
    assert (res.success, res.complete, res.checksum, res.fileSize, res.fileName, res.modDate) == (True, True, "match", "unchecked", "", None)
    inf = sdzlib.Inflater()
    out = sdzlib.mergeBuffers(inf.append(fx("simple.gz")))
    res = inf.finish()
    assert out == fx("simple.txt") and res.success and res.fileName == "simple.txt"
    assert res.modDate.timestamp() == 1576725008 and res.checksum == "match" and res.fileSize == "match"
    inf = sdzlib.Inflater(raw=True)
    assert sdzlib.mergeBuffers(inf.append(fx("simple.raw"))) == fx("simple.txt")


def test_inflate_batch_entry_point(fx):
    bufs = [fx("simple.deflate"), fx("simple.gz"), b"\x78\x01\x07", fx("simple.raw"), fx("paradiselost.gz")[:5000]]
    out = sdzlib.inflateBatch(bufs)
    assert out[0]["data"] == fx("simple.txt") and out[0]["error"] is None and out[0]["result"].checksum == "match"
    assert out[1]["result"].fileName == "simple.txt"
    assert out[2]["error"] == "inflate error: invalid block type" and out[2]["data"] == b""
    assert out[3]["data"] == fx("simple.txt")
    assert out[4]["error"] == "Unexpected EOF during decompression" and not out[4]["result"].complete
    assert fx("paradiselost.txt").startswith(out[4]["data"]) and len(out[4]["data"]) > 1000


# ---------------------------------------------------------------- checksums

def test_checksums_vs_oracle_and_zlib():
    rnd = random.Random(5)
    for n in [0, 1, 2, 3, 15, 16, 17, 31, 511, 512, 513, 4095, 5551, 5552, 5553, 11104, 16384, 16656, 65536, 100003,
              262144, 262145, 1 << 20, (1 << 20) + 5552 * 3 - (1 << 20) % 5552]:
        b = bytes(rnd.getrandbits(8) for _ in range(n)) if n < 70000 else os.urandom(n)
        for seed in (None, 0, 1, -1, 0x7fffffff, -2147483648, 0xfff1fff1 - (1 << 32)):
            a_exp = O.adler32(b) if seed is None else O.adler32(b, seed)
            c_exp = O.crc32(b) if seed is None else O.crc32(b, seed)
            a_got = sdzlib.adler32(b) if seed is None else sdzlib.adler32(b, seed)
            c_got = sdzlib.crc32(b) if seed is None else sdzlib.crc32(b, seed)
            assert a_got == a_exp, ("adler", n, seed, a_got, a_exp)
            assert c_got == c_exp, ("crc", n, seed, c_got, c_exp)
        if n % 5552 or n == 0:
            assert sdzlib.adler32(b) & 0xFFFFFFFF == zlib.adler32(b)
        assert sdzlib.crc32(b) & 0xFFFFFFFF == zlib.crc32(b)


def test_checksum_unaligned_views():
    base = np.frombuffer(os.urandom(300000), dtype=np.uint8)
    for start in (1, 2, 3, 5, 7, 13, 15):
        for n in (0, 1, 17, 5552, 70001, 262144 + 19):
            v = base[start:start + n]
            assert sdzlib.adler32(v) == O.adler32(v.tobytes())
            assert sdzlib.crc32(v) == O.crc32(v.tobytes())


def test_seed_chaining_one_pass():
    data = os.urandom(3_000_000)
    lens = [5552 * 4, 1, 0, 700000, 5552, 333333, 11104 * 8, 262144, 5552 * 30]
    lens.append(len(data) - sum(lens))
    a_vals = sdzlib.adler32_chain(data, lens)
    c_vals = sdzlib.crc32_chain(data, lens)
    a, c, off = 1, 0, 0
    for i, n in enumerate(lens):
        a = O.adler32(data[off:off + n], a)
        c = O.crc32(data[off:off + n], c)
        off += n
        assert int(a_vals[i]) == a, ("adler chain", i, n)
        assert int(c_vals[i]) == c, ("crc chain", i, n)
    assert int(c_vals[-1]) & 0xFFFFFFFF == zlib.crc32(data)


# ---------------------------------------------------------------- generated corpora vs oracle

@pytest.mark.parametrize("level", [1, 6, 9])
def test_corpora_all_containers(level):
    streams, plains = [], []
    for kind, n in ((K.TEXT, 65536), (K.BINARY, 65536), (K.RUNS, 65536), (K.RANDOM, 40000), (K.TINY, 150), (K.TEXT, 300000)):
        for idx in range(3):
            plain = K.generate(kind, 100 * level + idx, n)
            for cont in (K.RAW, K.ZLIB, K.GZIP, K.GZIP_NAME):
                streams.append(K.compress(plain, level, cont))
                plains.append(plain.tobytes())
    got = check_against_oracle(streams)
    n_complete = 0
    for (b, r), p in zip(got, plains):
        assert b == p
        n_complete += r.complete
    assert n_complete >= len(streams) * 3 // 4          # some raw streams are incomplete by Q15


def test_preset_dictionary_streams():
    dic = bytes(K.generate(K.TEXT, 999, 470))
    dictid = O.adler32(dic)
    streams, dicts = [], []
    for idx in range(6):
        plain = K.generate(K.TEXT, 500 + idx, 20000 + 7777 * idx)
        streams.append(K.compress(plain, 6, K.ZLIB_DICT, dic, dictid)); dicts.append(dic)
    streams.append(streams[0]); dicts.append(None)                # dictionary required
    streams.append(streams[0]); dicts.append(dic + b"!")          # wrong dictionary
    streams.append(K.compress(K.generate(K.TEXT, 1, 5000), 6, K.ZLIB)); dicts.append(dic)   # not needed: ignored
    big = bytes(K.generate(K.TEXT, 77, 40000))                    # > 32 KiB dictionary: last 32767 bytes are used (Q14)
    streams.append(K.compress(K.generate(K.TEXT, 78, 30000), 6, K.ZLIB_DICT, big, O.adler32(big))); dicts.append(big)
    q1 = bytes(K.generate(K.TEXT, 79, 5552))                      # dictionary length hits Q1
    streams.append(K.compress(K.generate(K.TEXT, 80, 9000), 6, K.ZLIB_DICT, q1, O.adler32(q1))); dicts.append(q1)
    streams.append(K.compress(K.generate(K.TEXT, 80, 9000), 6, K.ZLIB_DICT, q1, zlib.adler32(q1))); dicts.append(q1)
    got = check_against_oracle(streams, dicts, [O.MODE_INFLATER] * len(streams))
    assert got[0][1].success and got[6][1].thrown_append == O.THROW_DICT_REQUIRED
    assert got[7][1].thrown_append == O.THROW_DICT_INVALID and got[8][1].success


# ---------------------------------------------------------------- Appendix A behaviours

def test_divergence_vectors_and_quirks():
    streams = [
        bytes.fromhex("030200"),                                   # D1 / Q6 zero fill
        bytes.fromhex("05c08100000000009056fe2b0000"),             # D2 / Q9 no EOB
        bytes.fromhex("0dc081080000000020d6fd252e02"),             # D3 / Q9 empty distance tree
        zlib.compress(b"hello world") + b"\0",                     # Q4 trailing byte
        zlib.compress(b""),                                        # Q8
        raw_deflate(b"hello hello hello"),                         # Q15
        raw_deflate(b"a"), raw_deflate(b"ab"), raw_deflate(b"abc"), raw_deflate(b"abcdefgh"),
        b"\x78\x02\x03\x00", b"\x79\x9c\x03\x00", b"\x88\x1c\x03\x00", b"\x1f\x8c\x08\x00", b"\x78\x01\x07",
        b"\x78\x01\x01\x05\x00\x00\x00", b"x", b"", b"\x1f", b"\x78", b"\x1f\x8b", b"\x1f\x8b\x08",
    ]
    g = bytearray(gzip.compress(b"hello", mtime=1))                # Q5 FEXTRA
    g[3] |= 4
    g[10:10] = b"\x02\x00ab"
    streams.append(bytes(g))
    g2 = bytearray(gzip.compress(b"hello comment", mtime=77))      # FCOMMENT + FHCRC are skipped
    g2[3] |= 16 | 2
    g2[10:10] = b"a comment\x00\x12\x34"
    streams.append(bytes(g2))
    check_against_oracle(streams)
    check_against_oracle(streams, None, [O.MODE_INFLATER] * len(streams))
    check_against_oracle(streams, None, [O.MODE_RAW] * len(streams))


def test_q1_final_chunk_lengths():
    """A valid zlib stream whose last 16 KiB chunk is 5552 or 11104 bytes long is reported as a
    checksum mismatch by the reference (Q1)."""
    streams = []
    for total in (5552, 11104, 16384 + 5552, 32768 + 11104, 16384 * 3, 5551, 16384 + 5553):
        streams.append(K.compress(K.generate(K.TEXT, total, total), 6, K.ZLIB))
    got = check_against_oracle(streams)
    assert [r.checksum_state for _, r in got] == [2, 2, 2, 2, 1, 1, 1]


def test_q2_stored_blocks():
    streams = []
    for n in (1, 100, 16383, 32768, 49151, 49152, 49153, 65535, 65536, 70000, 100000, 131072):
        streams.append(zlib.compress(os.urandom(n), 6))
        streams.append(zlib.compress(os.urandom(n), 0))
    # text, then incompressible, then text: stored blocks in the middle of a stream
    for k in range(4):
        plain = K.generate(K.TEXT, k, 30000 + 5000 * k).tobytes() + os.urandom(20000 + 3000 * k) + K.generate(K.TEXT, k + 9, 40000).tobytes()
        streams.append(zlib.compress(plain, 6))
    check_against_oracle(streams)


def test_truncation_at_every_byte(fx):
    """Input exhaustion at every position of small streams (incl. mid-header, mid-trailer)."""
    bases = [fx("simple.deflate"), fx("simple.gz"), fx("simple.raw"),
             K.compress(K.generate(K.TEXT, 3, 3000), 6, K.ZLIB), K.compress(K.generate(K.BINARY, 3, 2000), 9, K.GZIP_NAME),
             K.compress(K.generate(K.RANDOM, 3, 600), 6, K.ZLIB), K.compress(K.generate(K.RUNS, 4, 5000), 6, K.RAW)]
    streams = []
    for b in bases:
        step = 1 if len(b) < 400 else 7
        streams += [b[:k] for k in range(0, len(b), step)]
    check_against_oracle(streams)
    check_against_oracle(streams, None, [O.MODE_INFLATER] * len(streams))


def test_truncation_large(fx):
    b = fx("paradiselost.deflate")
    rnd = random.Random(9)
    cuts = sorted(rnd.sample(range(2, len(b)), 60)) + [len(b) - 1, len(b) - 2, len(b) - 3, len(b) - 4, len(b) - 5]
    check_against_oracle([b[:k] for k in cuts])
    g = fx("paradiselost.gz")
    check_against_oracle([g[:len(g) - k] for k in range(1, 12)])


def test_corrupted_streams():
    """Random single-byte corruptions: the first event (error, stall or end) must agree."""
    rnd = random.Random(21)
    base = [K.compress(K.generate(K.TEXT, 40, 20000), 6, K.ZLIB), K.compress(K.generate(K.BINARY, 41, 9000), 6, K.GZIP),
            K.compress(K.generate(K.TINY, 42, 120), 6, K.ZLIB), K.compress(K.generate(K.RUNS, 43, 30000), 6, K.RAW)]
    streams = []
    for b in base:
        for _ in range(60):
            x = bytearray(b)
            pos = rnd.randrange(len(x)) if rnd.random() < 0.5 else rnd.randrange(min(len(x), 120))
            x[pos] ^= 1 << rnd.randrange(8)
            streams.append(bytes(x))
    got = run_batch(streams)
    bad = []
    for i, (s, (gb, gr)) in enumerate(zip(streams, got)):
        eb, er = O.inflate_oneshot(s)
        # exact, including the streams on which the reference never returns (SDZ_THROW_HANG)
        if gr.observable() != er.observable() or (not er.thrown_append and gb != eb):
            bad.append((i, len(s), gr.observable(), er.observable()))
    assert not bad, "%d of %d differ: %s" % (len(bad), len(streams), bad[:8])


# ---------------------------------------------------------------- benchmark-size properties

def test_batch_4096_text_streams_roundtrip():
    n, plen = 4096, 65536
    comp, stride, clen, plain = K.make_batch(K.TEXT, n, plen, 6, K.ZLIB, keep_plain=True)
    views = [comp[i * stride:i * stride + int(clen[i])] for i in range(n)]
    arena, off, res = A.inflate_batch_raw(views, None, None, np.full(n, plen, dtype=np.uint64))
    assert all(res[i].success and res[i].checksum_state == 1 and res[i].out_len == plen for i in range(n))
    assert np.array_equal(arena[:n * plen], plain)
    # checksum of checksums: every running Adler-32 equals zlib's over the plaintext
    for i in range(0, n, 97):
        assert res[i].running_checksum & 0xFFFFFFFF == zlib.adler32(plain[i * plen:(i + 1) * plen].tobytes())


def test_sizing_pass_matches_decode():
    streams = [K.compress(K.generate(K.TEXT, i, 1000 + 3571 * i), 6, K.ZLIB) for i in range(20)]
    views = [np.frombuffer(s, dtype=np.uint8) for s in streams]
    arena, off, res = A.inflate_batch_raw(views)
    assert [int(r.out_len) for r in res] == [1000 + 3571 * i for i in range(20)]


def test_pipelined_host_path_many_small_streams():
    """n >= 32768 streams makes sdz_inflate_batch cut the batch into sub-batches (pipelined on three
    CUDA streams); records and bytes must not depend on the cut.  Separate Python buffers force the
    staging path; test_batch_4096_text_streams_roundtrip covers the packed one."""
    n = 40000
    plains = [K.generate(K.TINY if i % 3 else K.TEXT, i, 20 + (i * 37) % 400).tobytes() for i in range(n)]
    streams = [zlib.compress(p, 6) if i % 2 else gzip.compress(p, 6, mtime=i + 1) for i, p in enumerate(plains)]
    views = [np.frombuffer(s, dtype=np.uint8) for s in streams]
    arena, off, res = A.inflate_batch_raw(views)
    for i in range(n):
        r = res[i]
        assert r.success and r.out_len == len(plains[i]), i
        if i % 997 == 0 or i in (19999, 20000, 39999):
            assert bytes(arena[int(off[i]):int(off[i]) + int(r.out_len)]) == plains[i]
            eb, er = O.inflate_oneshot(streams[i])
            assert r.observable() == er.observable()


def test_cfg4_mixed_container_batch():
    """BASELINE configs[3] on one GPU: gzip + raw + zlib (+ preset dictionary), levels 1/6/9, text / binary /
    tiny (fixed blocks) / incompressible <= 49,151 B (stored blocks) / runs - one batch, every finish() record
    and every byte compared with the oracle."""
    streams, dicts, modes = _cfg4_batch(1500)
    got = check_against_oracle(streams, dicts, modes)
    assert sum(1 for _, r in got if r.success) > 1300          # raw streams may be incomplete by Q15
    assert {r.container for _, r in got} == {0, 1, 2}


def _cfg4_batch(n_streams, seed0=7000):
    dic = bytes(K.generate(K.TEXT, 4242, 470))
    dictid = O.adler32(dic)
    rnd = random.Random(44)
    streams, dicts, modes = [], [], []
    kinds = [(K.TEXT, 65536), (K.BINARY, 65536), (K.TINY, 0), (K.RANDOM, 0), (K.RUNS, 65536), (K.TEXT, 20000)]
    for i in range(n_streams):
        kind, n = kinds[i % len(kinds)]
        if kind == K.TINY:
            n = 1 + rnd.randrange(200)
        elif kind == K.RANDOM:
            n = 1 + rnd.randrange(49151)
        plain = K.generate(kind, seed0 + i, n)
        level = (1, 6, 9)[i % 3]
        cont = (K.GZIP, K.RAW, K.ZLIB, K.GZIP_NAME, K.ZLIB_DICT)[i % 5]
        if cont == K.ZLIB_DICT:
            streams.append(K.compress(plain, level, cont, dic, dictid)); dicts.append(dic); modes.append(O.MODE_INFLATER)
        else:
            streams.append(K.compress(plain, level, cont)); dicts.append(None)
            modes.append(O.MODE_RAW if cont == K.RAW and i % 2 else O.MODE_SNIFF)
    return streams, dicts, modes


def test_multi_device_context_partitions_one_batch():
    """X2 / SURVEY 8e: ONE batch handed to a multi-device context is partitioned per stream (contiguous ranges balanced
    on compressed bytes), decoded by one pipeline per device, and comes back - bytes and records, in caller order -
    bit-identical to the single-device result and to the oracle.  On a one-GPU box the device list repeats device 0
    (three independent pipelines): the same partition / thread / gather code runs."""
    import torch
    from sdzlib import _native as N
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    streams, dicts, modes = _cfg4_batch(1200, seed0=9100)
    streams += [b"", b"\x78", b"\x78\x01\x07"]                     # degenerate inputs ride along
    dicts += [None] * 3
    modes += [O.MODE_SNIFF] * 3
    views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
    single = A.inflate_batch_raw(views, dicts, modes, None)
    mctx = N.Context(devices)
    assert mctx.device_count() == len(devices)
    arena, off, res = A.inflate_batch_raw(views, dicts, modes, None, ctx=mctx)
    cut = mctx.last_partition()
    assert cut[0] == 0 and cut[-1] == len(streams) and all(a <= b for a, b in zip(cut, cut[1:]))
    share = [sum(len(streams[i]) for i in range(cut[d], cut[d + 1])) for d in range(len(devices))]
    assert max(share) - min(share) <= 2 * max(len(s) for s in streams) + 2048, share      # balanced on compressed bytes
    assert list(off) == list(single[1])
    for i, s in enumerate(streams):
        r, r1 = res[i], single[2][i]
        assert r.observable() == r1.observable(), i
        n = int(r.out_len)
        if not r.thrown_append:
            assert bytes(arena[int(off[i]):int(off[i]) + n]) == bytes(single[0][int(off[i]):int(off[i]) + n]), i
        if i % 7 == 0:
            eb, er = O.inflate_oneshot(bytes(s), dictionary=dicts[i], mode=modes[i])
            assert r.observable() == er.observable(), i
            if not er.thrown_append:
                assert bytes(arena[int(off[i]):int(off[i]) + n]) == eb, i
    # a second, larger call reuses the grown buffers of every child; sizes pass goes through the partition too
    sizes = np.zeros(len(views), dtype=np.uint64)
    ins = (N.In * len(views))()
    keep = []
    for i, v in enumerate(views):
        ins[i].data = v.ctypes.data if v.size else None
        ins[i].len = int(v.size)
        ins[i].mode = modes[i]
        if dicts[i] is not None:
            dv = np.frombuffer(dicts[i], dtype=np.uint8)
            keep.append(dv)
            ins[i].dict = dv.ctypes.data; ins[i].dict_len = int(dv.size)
    mctx.check(mctx.lib.sdz_inflate_sizes(mctx.h, ins, len(views), sizes.ctypes.data, 0))
    assert [int(x) for x in sizes] == [int(r.out_len) for r in res]
    mctx.close()


def test_checksum_batch():
    rnd = random.Random(8)
    sizes = [0, 1, 15, 16, 17, 5551, 5552, 5553, 11104, 16384, 40000, 65536, 100000, 262147]
    bufs = [os.urandom(n) for n in sizes] * 3
    kinds = ["adler32" if i % 2 else "crc32" for i in range(len(bufs))]
    seeds = [rnd.choice([0, 1, -1, 0x12345678, -2147483648]) for _ in bufs]
    got = sdzlib.checksum_batch(bufs, kinds, seeds)
    for b, k, s, g in zip(bufs, kinds, seeds, got):
        exp = O.adler32(b, s) if k == "adler32" else O.crc32(b, s)
        assert g == exp, (len(b), k, s, g, exp)
    got = sdzlib.checksum_batch(bufs[:6], kinds[:6])          # reference default seeds
    for b, k, g in zip(bufs, kinds, got):
        assert g == (O.adler32(b) if k == "adler32" else O.crc32(b))


def test_mixed_kind_stress_roundtrip():
    """8,192 streams of every corpus kind and several sizes/levels in one batch, decoded into a poisoned
    arena: bytes must equal the plaintext for every stream (catches intra-warp ordering bugs)."""
    views, plains = [], []
    for kind, n, level, cnt in ((K.RUNS, 65536, 6, 2048), (K.TEXT, 30000, 1, 2048), (K.BINARY, 65536, 9, 1024),
                                (K.RUNS, 9000, 1, 1024), (K.TEXT, 131072, 6, 512), (K.TINY, 180, 6, 1536)):
        comp, stride, clen, plain = K.make_batch(kind, cnt, n, level, K.ZLIB, first_index=90000 + len(views), keep_plain=True)
        for i in range(cnt):
            views.append(comp[i * stride:i * stride + int(clen[i])])
            plains.append(plain[i * n:(i + 1) * n])
    order = list(range(len(views)))
    random.Random(3).shuffle(order)
    views = [views[i] for i in order]
    plains = [plains[i] for i in order]
    arena, off, res = A.inflate_batch_raw(views)
    for i, p in enumerate(plains):
        r = res[i]
        assert r.success and r.out_len == p.size, (i, r.observable())
        assert np.array_equal(arena[int(off[i]):int(off[i]) + p.size], p), i


# ---------------------------------------------------------------- one large stream (config 5)

def _large_plain(kind, mib, first_index):
    return np.concatenate([K.generate(kind, first_index + i, 65536) for i in range(mib * 16)])


def _check_large(stream, mode, plain=None):
    view = np.frombuffer(stream, dtype=np.uint8)
    out, r = A.inflate_large_raw(view, mode)
    exp_bytes, exp = O.inflate_oneshot(stream, mode=mode)
    go, eo = r.observable(), exp.observable()
    assert go == eo, {k: (go.get(k), eo.get(k)) for k in set(go) | set(eo) if go.get(k) != eo.get(k)}
    if not exp.thrown_append:
        got = out[:int(r.out_len)]
        assert got.size == len(exp_bytes)
        assert np.array_equal(got, np.frombuffer(exp_bytes, dtype=np.uint8))
    if plain is not None:
        assert np.array_equal(out[:int(r.out_len)], plain)
    return r


@pytest.mark.parametrize("container", ["gzip", "zlib", "raw"])
def test_large_stream_block_parallel(container):
    """One 24 MiB stream (several hundred deflate blocks) through sdz_inflate_large: identical bytes and
    record to the oracle's one-shot Inflater."""
    plain = _large_plain(K.TEXT, 24, 7000)
    if container == "gzip":
        s = gzip.compress(plain.tobytes(), 6, mtime=1234567)
        mode = O.MODE_SNIFF
    elif container == "zlib":
        s = zlib.compress(plain.tobytes(), 6)
        mode = O.MODE_INFLATER
    else:
        s = raw_deflate(plain.tobytes(), 6)
        mode = O.MODE_RAW
    ctx = sdzlib.default_context()
    before = ctx.launch_count()
    r = _check_large(s, mode, plain)
    assert r.n_blocks > 100                      # took the block-parallel path, not the sequential hand-over
    assert ctx.launch_count() - before >= 6


def test_large_stream_kinds_levels_and_flush_points():
    """Binary / run-heavy data, levels 1 and 9, sync-flush points (empty stored blocks -> sequential
    hand-over) and a multi-member-looking tail all give the oracle's record."""
    cases = []
    cases.append((zlib.compress(_large_plain(K.BINARY, 6, 100).tobytes(), 9), O.MODE_INFLATER))
    cases.append((zlib.compress(_large_plain(K.RUNS, 8, 200).tobytes(), 1), O.MODE_SNIFF))
    cases.append((gzip.compress(_large_plain(K.RANDOM, 2, 300).tobytes(), 6, mtime=0), O.MODE_SNIFF))   # stored blocks
    co = zlib.compressobj(6)
    p = _large_plain(K.TEXT, 2, 400).tobytes()
    cases.append((co.compress(p[:700000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(p[700000:]) + co.flush(), O.MODE_SNIFF))
    cases.append((gzip.compress(p, 6, mtime=5)[:-9], O.MODE_SNIFF))                                    # truncated trailer
    cases.append((zlib.compress(p, 6)[:300000], O.MODE_SNIFF))                                         # truncated body
    bad = bytearray(zlib.compress(p, 6))
    bad[len(bad) // 2] ^= 0x55
    cases.append((bytes(bad), O.MODE_SNIFF))                                                           # damaged body
    bad2 = bytearray(gzip.compress(p, 6, mtime=9))
    bad2[-6] ^= 1
    cases.append((bytes(bad2), O.MODE_SNIFF))                                                          # wrong CRC
    for s, m in cases:
        _check_large(s, m)


def test_large_stream_stored_blocks_stay_block_parallel():
    """Sync-flush points (empty stored blocks, what pigz-style writers emit between chunks) and small stored blocks of
    incompressible data: sdz_large_plan replays the reference's window bookkeeping over the chain of blocks, so these
    streams stay on the block-parallel path; a stored block the reference itself cuts short still goes to the
    sequential decoder, with the reference's (broken) record."""
    p = _large_plain(K.TEXT, 8, 500).tobytes()
    co = zlib.compressobj(6)
    s = b""
    for k in range(0, len(p), 1 << 20):
        s += co.compress(p[k:k + (1 << 20)]) + co.flush(zlib.Z_SYNC_FLUSH)
    s += co.flush()
    ctx = sdzlib.default_context()
    before = ctx.launch_count()
    r = _check_large(s, O.MODE_SNIFF, np.frombuffer(p, dtype=np.uint8))
    assert r.success and r.n_blocks > 100 and ctx.launch_count() - before >= 6
    # incompressible pieces of 1 .. 12 KiB between text: stored blocks with data
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    s, plain = b"", b""
    rng = random.Random(17)
    for k in range(40):
        t = p[k * 150000:(k + 1) * 150000]
        noise = os.urandom(rng.choice((1, 100, 1000, 4000, 12000)))
        s += co.compress(t) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(noise) + co.flush(zlib.Z_SYNC_FLUSH)
        plain += t + noise
    s += co.flush()
    before = ctx.launch_count()
    r = _check_large(s, O.MODE_SNIFF)
    exp_bytes, exp = O.inflate_oneshot(s)
    if exp.success:                                     # (every stored block survived the reference's Q2 rule)
        assert ctx.launch_count() - before >= 6
    # one large stored block in the middle: the reference loses its tail (Q2) - the oracle's record, whatever it is
    _check_large(zlib.compress(p[:1500000] + os.urandom(60000) + p[1500000:3000000], 6), O.MODE_SNIFF)


def test_large_stream_far_references_and_small_blocks():
    """Windows that span many tiny blocks (Z_FULL_FLUSH is avoided: it emits stored blocks; level-1 output
    of short period data has long chains of window references across block boundaries)."""
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, 30000, dtype=np.uint8)
    parts = []
    for i in range(400):
        b = base.copy()
        idx = rng.integers(0, b.size, 40)
        b[idx] = rng.integers(0, 256, 40, dtype=np.uint8)
        parts.append(b)
    plain = np.concatenate(parts)                                   # every byte is (transitively) a ~30 KB-far copy
    for level in (1, 6):
        _check_large(zlib.compress(plain.tobytes(), level), O.MODE_SNIFF, plain)


@pytest.mark.parametrize("n_parts,mib", [(2, 8), (3, 24), (7, 3)])
def test_large_stream_multi_part_protocol_on_one_gpu(n_parts, mib):
    """The multi-GPU protocol (index slices, merged plan, output slices, 32 KiB window relay, CRC combine) with all
    ranks played by one device: identical bytes and record to the oracle."""
    from sdzlib import large as LG
    plain = _large_plain(K.TEXT, mib, 9100 + mib)
    s = gzip.compress(plain.tobytes(), 6, mtime=77)
    out, r = LG.inflate_large_parts(np.frombuffer(s, dtype=np.uint8), n_parts)
    assert np.array_equal(out, plain)
    _, exp = O.inflate_oneshot(s)
    assert r.observable() == exp.observable()


def test_large_stream_slices_shorter_than_the_window():
    """100 KB of output over 6 parts: every slice is shorter than 32 KiB, so windows span several slices."""
    from sdzlib import large as LG
    rng = np.random.default_rng(4)
    base = rng.integers(97, 123, 9000, dtype=np.uint8)
    plain = np.concatenate([np.roll(base, i * 17) for i in range(11)])[:100000]
    s = gzip.compress(plain.tobytes(), 6, mtime=1)
    out, r = LG.inflate_large_parts(np.frombuffer(s, dtype=np.uint8), 6)
    assert np.array_equal(out, plain)
    assert r.success and r.checksum_state == 1 and r.size_state == 1


@pytest.mark.parametrize("n_parts", [2, 3, 5])
def test_large_stream_zlib_and_raw_over_several_parts(n_parts):
    """zlib / raw streams on the multi-part protocol: the running Adler-32 is joined from per-part (adler32, length)
    with adler32_combine, and the final <= 16 KiB chunk is one reference call - so a stream whose output ends with a
    5552- or 11104-byte chunk is reported "mismatch" exactly like the reference reports it."""
    from sdzlib import large as LG
    for total in (6 << 20, (6 << 20) + 5552, (5 << 20) + 11104, (3 << 20) + 16384 * 3 + 1, 16384 + 11104):
        plain = _large_plain(K.TEXT, (total >> 20) + 1, 9300)[:total]
        for s in (zlib.compress(plain.tobytes(), 6), raw_deflate(plain.tobytes(), 6)):
            out, r = LG.inflate_large_parts(np.frombuffer(s, dtype=np.uint8), n_parts)
            assert np.array_equal(out, plain)
            _, exp = O.inflate_oneshot(s)
            assert r.observable() == exp.observable(), (total, n_parts)
            if s[0] == 0x78:
                assert r.checksum_state == (2 if total % 16384 in (5552, 11104) else 1)


def _append_events(make, parts, thrown_of):
    """[(chunk lengths, bytes) per append() ..., ('throw', thrown, msg_id)] - the sequence ends at the first throw"""
    inf = make()
    ev = []
    for p in parts:
        try:
            chunks = inf.append(p)
            ev.append(([len(x) for x in chunks], b"".join(bytes(x) for x in chunks)))
        except Exception as e:                      # noqa: BLE001 - both twins raise their own classes
            ev.append(("throw",) + thrown_of(e))
            return ev, None
    return ev, inf.finish()


def _compare_streaming(s, parts, raw=False, dictionary=None):
    """device Inflater vs the oracle twin of class Inflater on the same append() sequence: chunk shapes and bytes of every
    call, what the reference throws (and at which call), and the finish() record.  Returns True when the reference threw."""
    state = {"unchecked": 0, "match": 1, "mismatch": 2}
    oev, orr = _append_events(lambda: O.Inflater(raw=raw, dictionary=dictionary), parts,
                              lambda e: (e.thrown, e.msg_id if e.thrown == 4 else 0))
    gev, grr = _append_events(lambda: sdzlib.Inflater(raw=raw, dictionary=dictionary), parts, lambda e: (e.thrown, e.msg_id))
    assert len(gev) == len(oev), ([len(p) for p in parts], gev[-1][:1], oev[-1][:1])
    for k, (ge, oe) in enumerate(zip(gev, oev)):
        assert ge[0] == oe[0], ("append #%d" % k, [len(p) for p in parts], ge[:1] if ge[0] != "throw" else ge, oe[:1] if oe[0] != "throw" else oe)
        assert ge == oe, ("append #%d: bytes / codes differ" % k, [len(p) for p in parts])
    if orr is None:
        return True
    assert (grr.success, grr.complete, state[grr.checksum], state[grr.fileSize]) == \
           (bool(orr.success), bool(orr.complete), orr.checksum_state, orr.size_state), [len(p) for p in parts]
    assert grr.fileName == orr.file_name
    return False


def test_inflater_append_split_points():
    """Inflater.append() over 2-4 chunks cut at random byte positions, six kinds of streams, ALL 240 splits
    compared with the oracle twin of class Inflater - including the splits where the reference breaks: a boundary inside a
    dynamic block header makes the next append() throw "inflate error: " (Q3), a boundary inside stored data ends the
    stored block there and the rest is parsed as block headers (Q2), and whatever follows from that."""
    import random
    rnd = random.Random(7)
    cases = []
    for kind, n, cont in ((K.TEXT, 40000, K.ZLIB), (K.TEXT, 70000, K.GZIP_NAME), (K.BINARY, 50000, K.RAW), (K.RUNS, 60000, K.ZLIB),
                          (K.RANDOM, 30000, K.GZIP), (K.TINY, 150, K.ZLIB)):
        cases.append((cont, bytes(K.compress(K.generate(kind, 99, n), 6, cont))))
    threw = total = 0
    for cont, s in cases:
        for _ in range(40):
            cuts = sorted(rnd.randrange(1, len(s)) for _ in range(rnd.choice((1, 1, 2, 3))))
            parts = [s[a:b] for a, b in zip([0] + cuts, cuts + [len(s)])]
            threw += _compare_streaming(s, parts, raw=cont == K.RAW)
            total += 1
    assert total == 240 and 10 <= threw <= 120, (total, threw)


def test_inflater_boundary_inside_every_header_state():
    """cuts at EVERY byte of the first 400 bytes of a dynamic-block stream (container header, TYPE, TABLE, BTREE, DTREE, first
    symbols) and at every byte around the second block header and the trailer: the Q3 window is hit exactly where the
    reference has it"""
    s = bytes(K.compress(K.generate(K.TEXT, 5, 90000), 6, K.GZIP_NAME))
    threw = 0
    for c in list(range(1, 400)) + list(range(len(s) - 24, len(s))):
        threw += _compare_streaming(s, [s[:c], s[c:]])
    assert threw > 20
    # and three-way cuts with a tiny middle part (the reference's input frontier within a few bytes of both ends)
    rnd = random.Random(21)
    for _ in range(60):
        a = rnd.randrange(1, len(s) - 40)
        b = a + rnd.randrange(1, 12)
        _compare_streaming(s, [s[:a], s[a:b], s[b:]])


def test_inflater_many_small_appends():
    """carried state over hundreds of append() calls: 1-byte appends of a small stream, 1-300 byte appends of a 70 KB
    one, a preset-dictionary stream in pieces, stored blocks cut at their edges"""
    rnd = random.Random(3)
    tiny = bytes(K.compress(K.generate(K.TINY, 1, 120), 6, K.ZLIB))
    _compare_streaming(tiny, [tiny[i:i + 1] for i in range(len(tiny))])
    fixed = zlib.compress(b"hello hello hello hello, streaming world" * 3, 9)
    _compare_streaming(fixed, [fixed[i:i + 1] for i in range(len(fixed))])
    for kind, cont, n in ((K.TEXT, K.ZLIB, 70000), (K.BINARY, K.GZIP, 40000), (K.RUNS, K.RAW, 70000)):
        s = bytes(K.compress(K.generate(kind, 17, n), 6, cont))
        parts, o = [], 0
        while o < len(s):
            m = rnd.randrange(1, 300)
            parts.append(s[o:o + m])
            o += m
        _compare_streaming(s, parts, raw=cont == K.RAW)
    dic = bytes(K.generate(K.TEXT, 4242, 470))
    s = bytes(K.compress(K.generate(K.TEXT, 8, 50000), 6, K.ZLIB_DICT, dic, O.adler32(dic)))
    for cuts in ((3,), (5, 6), (1, 2, 3, 4, 5, 6, 7), (2000, 9000)):
        parts = [s[a:b] for a, b in zip((0,) + cuts, cuts + (len(s),))]
        _compare_streaming(s, parts, dictionary=dic)
    # stored blocks: boundaries at the LEN / NLEN words, inside the data, exactly at a block's end
    plain = os.urandom(20000)
    co = zlib.compressobj(0, zlib.DEFLATED, 15)
    st = co.compress(plain) + co.flush()
    for c in (2, 3, 4, 5, 6, 7, 8, 100, 16383 + 7, 16383 + 8, len(st) - 5, len(st) - 4, len(st) - 1):
        _compare_streaming(st, [st[:c], st[c:]])
    # bytes after the end of the stream arriving in a later append(): the reference spins (Q4)
    z = zlib.compress(b"abc" * 100)
    assert _compare_streaming(z + b"x", [z, b"x"])


def test_deflate_wrap_batch_matches_reference_containers():
    """the containers `Deflater` writes (src/sd-deflate.ts:98-165) around raw deflate payloads, with the source
    checksums computed on the device in one launch.  Byte-identical to the reference-format streams of the corpus tool
    (which reproduces the reference's own fixture, tests/test_corpus_tool.py), and the result inflates back through the
    drop-in API with checksum / fileSize "match" - including a source whose length is a multiple of 5552 (Q1: the
    reference's Deflater and Inflater agree with each other, not with zlib)."""
    plains = [K.generate(K.TEXT, 300 + i, n).tobytes() for i, n in enumerate((1, 100, 5552, 11104, 40000, 65536, 70001))]
    raws = [K.compress(p, 6, K.RAW) for p in plains]
    z = sdzlib.deflate_wrap_batch(raws, plains, "deflate")
    g = sdzlib.deflate_wrap_batch(raws, plains, "gzip", file_names=["stream.bin"] * len(plains), mtime=0x5d211b5e)
    g0 = sdzlib.deflate_wrap_batch(raws, plains, "gzip", mtime=0x5d211b5e)
    r = sdzlib.deflate_wrap_batch(raws, plains, "raw")
    for i, p in enumerate(plains):
        assert r[i] == raws[i]
        assert g[i] == K.compress(p, 6, K.GZIP_NAME) and g0[i] == K.compress(p, 6, K.GZIP)
        if len(p) % 5552:
            assert z[i] == K.compress(p, 6, K.ZLIB)
        else:                                                   # Q1: the trailer holds the REFERENCE's adler32 of the source
            assert z[i][:-4] == K.compress(p, 6, K.ZLIB)[:-4]
            assert int.from_bytes(z[i][-4:], "big", signed=True) == O.adler32(p) != zlib.adler32(p)
    for streams in (z, g):
        for i, o in enumerate(sdzlib.inflateBatch(streams)):
            assert o["data"] == plains[i] and o["error"] is None
            # a zlib stream of 5552 / 11104 bytes is "mismatch" in the reference even from its own Deflater? No: the Inflater
            # checksums the same <= 16 KiB chunk with the same function, so both sides carry the same Q1 value
            assert o["result"].checksum == "match", (i, len(plains[i]))
    assert all(o["result"].fileName == "stream.bin" and o["result"].fileSize == "match" for o in sdzlib.inflateBatch(g))
    dic = bytes(K.generate(K.TEXT, 4242, 470))
    p = plains[4]
    zd = sdzlib.deflate_wrap_batch([raw_deflate(p, 6, -15, dic)], [p], "deflate", dictionaries=[dic])[0]
    assert zd[:2] == b"\x78\x20" and int.from_bytes(zd[2:6], "big", signed=True) == O.adler32(dic)
    assert sdzlib.inflate(zd, dic) == p
