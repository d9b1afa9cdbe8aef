"""SDZ_PARITY_SPEC (include/sdzcuda.h): the RFC-strict mode of sdz_inflate_batch, checked against system zlib 1.3 -
the implementation the RFCs were written from - on the corpora of the reference-parity tests plus the streams the
reference mishandles (Q1, Q2, Q4, Q5, Q6, Q9, Q10, Q13, Q14).  The checker here is zlib, not the oracle:
the oracle restates the reference, quirks included."""
import gzip
import os
import random
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tools import corpus as K  # noqa: E402

import sdzlib  # noqa: E402
from sdzlib import api as A  # noqa: E402

MODE_SNIFF, MODE_INFLATER, MODE_RAW = 0, 1, 2


def zlib_expect(s, mode, zdict=None):
    """(bytes, status, text): status 'ok' | 'eof' (input ended early) | 'error' (+ zlib's message) | 'need_dict' | 'bad_dict'"""
    if mode == MODE_SNIFF:
        ident = len(s) >= 2 and ((s[0] == 0x78 and ((s[0] << 8) + s[1]) % 31 == 0) or (s[0] == 0x1f and s[1] == 0x8b))
        mode = MODE_INFLATER if ident else MODE_RAW
    wbits = -15 if mode == MODE_RAW else 47
    try:
        d = zlib.decompressobj(wbits) if zdict is None or mode == MODE_RAW else zlib.decompressobj(15, zdict)
        out = d.decompress(bytes(s))
    except zlib.error as e:
        t = str(e)
        if t.startswith("Error 2 "):
            return b"", "need_dict", ""
        if "setting zdict" in t:
            return b"", "bad_dict", ""
        return b"", "error", t.split(": ", 1)[1] if ": " in t else t
    return out, ("ok" if d.eof else "eof"), ""


def run_spec(streams, dicts=None, modes=None):
    views = [np.frombuffer(bytes(s), dtype=np.uint8) for s in streams]
    arena, off, res = A.inflate_batch_raw(views, dicts, modes, None, None, A.PARITY_SPEC)
    return [(bytes(arena[int(off[i]):int(off[i]) + int(res[i].out_len)]), res[i]) for i in range(len(views))]


def check_spec(streams, dicts=None, modes=None, texts=True):
    got = run_spec(streams, dicts, modes)
    bad = []
    for i, s in enumerate(streams):
        m = MODE_SNIFF if modes is None else modes[i]
        zd = None if dicts is None else dicts[i]
        if len(s) == 0 or (m == MODE_SNIFF and len(s) < 2):
            continue                                           # the API rejects these before any decoder sees them
        eb, st, text = zlib_expect(s, m, zd)
        gb, r = got[i]
        if st == "ok":
            good = r.success == 1 and r.complete == 1 and r.thrown_inflate == 0 and gb == eb
        elif st == "eof":
            good = r.success == 0 and r.complete == 0 and r.thrown_inflate == 8 and gb == eb      # SDZ_THROW_UNEXPECTED_EOF
        elif st == "need_dict":
            good = r.thrown_append == 3
        elif st == "bad_dict":
            good = r.thrown_append == 2
        else:
            if text in ("incorrect data check", "incorrect length check"):
                good = r.success == 0 and r.thrown_inflate == (9 if text == "incorrect data check" else 10)
            else:
                good = r.thrown_append == 4 and (not texts or A._MSG[r.msg_id] == text)
        if not good:
            bad.append((i, m, len(s), bytes(s)[:12].hex(), st, text, dict(success=r.success, complete=r.complete, ta=r.thrown_append,
                        ti=r.thrown_inflate, msg=A._MSG[r.msg_id], n=len(gb), n_exp=len(eb), same=gb == eb)))
    assert not bad, "%d of %d streams differ from zlib: %s" % (len(bad), len(streams), bad[:8])
    return got


def raw_deflate(data, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    return co.compress(data) + co.flush()


def test_spec_valid_corpora_equal_zlib():
    streams = []
    for kind in (K.TEXT, K.BINARY, K.TINY, K.RANDOM, K.RUNS):
        for level in (1, 6, 9):
            for fmt in (K.ZLIB, K.GZIP, K.GZIP_NAME, K.RAW):
                n = 300 if kind == K.TINY else 70000
                streams.append(K.compress(K.generate(kind, 100 + level, n), level, fmt))
    got = check_spec(streams)
    assert all(r.success for _, r in got)


def test_spec_streams_the_reference_mishandles():
    streams, dicts = [], []

    def add(s, d=None):
        streams.append(bytes(s)); dicts.append(d)

    # Q2: stored blocks whose data crosses the reference's 49,151-byte point
    for n in (49151, 49152, 65535, 65536, 70000, 131072, 300000):
        add(zlib.compress(os.urandom(n), 0))
        add(zlib.compress(os.urandom(n), 6))
    # Q1: last 16 KiB chunk of 5552 / 11104 bytes
    for total in (5552, 11104, 16384 + 5552, 32768 + 11104):
        add(K.compress(K.generate(K.TEXT, total, total), 6, K.ZLIB))
    # Q4: bytes after the end of the stream
    add(zlib.compress(b"hello world") + b"\0")
    add(gzip.compress(b"hello world", mtime=3) + b"trailing bytes")
    # Q5: FEXTRA; FCOMMENT; FHCRC with the right and a wrong header CRC
    g = bytearray(gzip.compress(b"hello extra field", mtime=1))
    g[3] |= 4
    g[10:10] = b"\x05\x00abcde"
    add(g)
    g = bytearray(gzip.compress(b"hello comment", mtime=77))
    g[3] |= 16
    g[10:10] = b"a comment\x00"
    add(g)
    g = bytearray(gzip.compress(b"hello header crc", mtime=78))
    g[3] |= 2
    g[10:10] = (zlib.crc32(bytes(g[:10])) & 0xffff).to_bytes(2, "little")
    add(g)
    g2 = bytearray(g)
    g2[10] ^= 1
    add(g2)
    g3 = bytearray(gzip.compress(b"reserved flag", mtime=5))
    g3[3] |= 0x20
    add(g3)
    # Q8 / Q13: empty streams (stored checksum 1 / CRC 0, ISIZE 0 are checked, not skipped)
    add(zlib.compress(b""))
    add(gzip.compress(b"", mtime=0))
    # Q6: a distance before the start of the output; Q9: no end-of-block code, empty distance tree
    add(bytes.fromhex("030200"))
    add(bytes.fromhex("05c08100000000009056fe2b0000"))
    add(bytes.fromhex("0dc081080000000020d6fd252e02"))
    # Q14: dictionaries of 32,767 / 32,768 / 40,000 bytes whose first bytes are referenced; Q1 on the dictionary id
    for dl in (32767, 32768, 40000, 5552, 11104):
        d = K.generate(K.TEXT, 900 + dl, dl).tobytes()
        plain = d[-32768:][:600] + K.generate(K.TEXT, 77, 5000).tobytes() + d[-30000:][:300]
        co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, 0, d)
        add(co.compress(plain) + co.flush(), d)
    s_dict = streams[-1]
    add(s_dict, None)                                          # dictionary required
    add(s_dict, b"not the dictionary")                         # wrong dictionary
    # wrong trailers
    z = bytearray(zlib.compress(b"adler mismatch" * 10)); z[-1] ^= 1; add(z)
    z = bytearray(gzip.compress(b"crc mismatch" * 10, mtime=1)); z[-5] ^= 1; add(z)
    z = bytearray(gzip.compress(b"size mismatch" * 10, mtime=1)); z[-1] ^= 1; add(z)
    # header variants
    for h in (b"\x78\x02\x03\x00", b"\x79\x9c\x03\x00", b"\x88\x1c\x03\x00", b"\x1f\x8c\x08\x00", b"\x78\x01\x07", b"\x1f\x8b\x07\x00",
              b"\x1f\x8b", b"\x1f\x8b\x08", b"\x78\x9c", b"\x78"):
        add(h)
    modes = [MODE_INFLATER] * len(streams)
    got = check_spec(streams, dicts, modes)
    assert got[0][1].success and len(got[0][0]) == 49151 and got[12][1].success and len(got[12][0]) == 300000
    # the same streams through inflate()'s sniffing rule
    check_spec(streams, dicts, None)


def test_spec_crafted_code_sets():
    """code sets zlib accepts and the reference rejects (arena beyond MANY = 1400, Q10) or the other way round (incomplete
    sets longer than one bit, Q9); a lone one-bit code-length code (D4 / Q11)"""
    import test_crafted as TC
    streams = [s for s in TC.lone_code_length_code_vectors()]
    streams += [s for s, _ in TC.literal_literal_match_vectors()]
    streams += TC.two_block_tail_vectors()[::7]
    check_spec(streams)
    check_spec(streams, None, [MODE_RAW] * len(streams))


def test_spec_truncation_and_corruption():
    rnd = random.Random(33)
    base = [K.compress(K.generate(K.TEXT, 40, 20000), 6, K.ZLIB), K.compress(K.generate(K.BINARY, 41, 9000), 6, K.GZIP),
            K.compress(K.generate(K.TINY, 42, 120), 6, K.ZLIB), K.compress(K.generate(K.RUNS, 43, 30000), 6, K.RAW),
            zlib.compress(os.urandom(3000), 0)]
    streams = []
    for b in base:
        for _ in range(80):
            x = bytearray(b)
            pos = rnd.randrange(len(x)) if rnd.random() < 0.5 else rnd.randrange(min(len(x), 120))
            x[pos] ^= 1 << rnd.randrange(8)
            streams.append(bytes(x))
        step = 1 if len(b) < 400 else max(1, len(b) // 150)
        streams += [b[:k] for k in range(2, len(b), step)]
    check_spec(streams)
    check_spec(streams, None, [MODE_INFLATER] * len(streams))


def test_spec_through_public_api():
    big_stored = zlib.compress(os.urandom(200000), 0)
    out = sdzlib.inflateBatch([big_stored, zlib.compress(b"x" * 5552)], parity="spec")
    assert out[0]["error"] is None and len(out[0]["data"]) == 200000 and out[0]["result"].checksum == "match"
    assert out[1]["result"].checksum == "match"                                    # Q1 is the reference's, not the RFC's
    ref = sdzlib.inflateBatch([big_stored, zlib.compress(b"x" * 5552)])
    assert ref[1]["result"].checksum == "mismatch"
