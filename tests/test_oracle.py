"""CPU tests: the oracle (C restatement of the reference) against the reference's own fixtures
(SURVEY Appendix B), against system zlib 1.3 on valid streams, and on the quirk vectors of
SURVEY Appendix A.  These pin the checker that the GPU parity tests rely on."""
import gzip
import os
import random
import zlib

import pytest

from oracle import oracle as O

DICT_470 = None


def raw_deflate(data, level=6, wbits=-15, zdict=None):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits) if zdict is None else zlib.compressobj(level, zlib.DEFLATED, wbits, 8, 0, zdict)
    return co.compress(data) + co.flush()


# ---------------------------------------------------------------- Appendix B fixtures

def test_paradiselost_deflate(fx):
    out, r = O.inflate_oneshot(fx("paradiselost.deflate"))
    assert out == fx("paradiselost.txt")
    assert (r.success, r.complete, r.checksum_state, r.size_state) == (1, 1, O.CHECK_TEXT.index("match"), 0)
    assert r.stored_checksum == -1949153550 and r.running_checksum == -1949153550
    assert r.container == 1 and r.mtime == 0 and r.name_len == 0 and r.n_blocks == 7


def test_paradiselost_gz(fx):
    out, r = O.inflate_oneshot(fx("paradiselost.gz"))
    assert out == fx("paradiselost.txt")
    assert (r.success, r.complete, r.checksum_state, r.size_state) == (1, 1, 1, 1)
    assert r.stored_checksum == -499006831 and r.stored_isize == 471162 and r.mtime == 1530824734
    assert fx("paradiselost.gz")[r.name_off:r.name_off + r.name_len] == b"paradiselost.txt"


def test_simple_all_containers(fx):
    txt = fx("simple.txt")
    out, r = O.inflate_oneshot(fx("simple.deflate"))
    assert out == txt and r.success and r.container == 1 and r.stored_checksum == -1612443532 and r.n_blocks == 1
    out, r = O.inflate_oneshot(fx("simple.gz"))
    assert out == txt and r.success and r.container == 2 and r.stored_checksum == 1488305224
    assert r.mtime == 1576725008 and r.size_state == 1
    out, r = O.inflate_oneshot(fx("simple.raw"))
    assert out == txt and r.success and r.container == 0 and r.checksum_state == 0 and r.size_state == 0


def test_vertices(fx):
    out, r = O.inflate_oneshot(fx("vertices.deflate"))
    assert len(out) == 43440 and r.success and r.checksum_state == 1 and r.n_blocks == 2
    assert r.stored_checksum == 1386812979
    assert zlib.crc32(out) == 0xbf55c371


def test_two_part_streaming_and_chunk_shapes(fx):
    inf = O.Inflater()
    b1 = inf.append(fx("paradiselost.part1.deflate"))
    b2 = inf.append(fx("paradiselost.part2.deflate"))
    assert b"".join(b1 + b2) == fx("paradiselost.txt")
    r = inf.finish()
    assert r.success and r.checksum_state == 1 and r.size_state == 0 and r.file_name == ""
    # single-buffer chunk shape: 28 x 16384 + 12410 (Appendix B)
    inf = O.Inflater()
    chunks = inf.append(fx("paradiselost.deflate"))
    assert [len(c) for c in chunks] == [16384] * 28 + [12410]


def test_gzip_inflater_record(fx):
    inf = O.Inflater()
    out = b"".join(inf.append(fx("simple.gz")))
    r = inf.finish()
    assert out == fx("simple.txt") and r.success and r.file_name == "simple.txt" and r.mtime == 1576725008


# ---------------------------------------------------------------- checksums

def test_checksums_golden(fx):
    t = fx("paradiselost.txt")
    assert O.adler32(t) == -1949153550 and O.crc32(t) == -499006831
    s = fx("simple.txt")
    assert O.adler32(s) == -1612443532 and O.crc32(s) == 1488305224
    d = (b"a" * 470)
    assert O.adler32(b"") == 1 and O.crc32(b"") == 0


def test_checksums_vs_zlib_and_seed_chaining():
    rnd = random.Random(7)
    for n in [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 255, 256, 4095, 5551, 5553, 16384, 65536, 100003]:
        b = bytes(rnd.getrandbits(8) for _ in range(n))
        assert O.adler32(b) & 0xFFFFFFFF == zlib.adler32(b)
        assert O.crc32(b) & 0xFFFFFFFF == zlib.crc32(b)
        cut = n // 3
        assert O.adler32(b[cut:], O.adler32(b[:cut])) & 0xFFFFFFFF == zlib.adler32(b) or cut % 5552 == 0
        assert O.crc32(b[cut:], O.crc32(b[:cut])) & 0xFFFFFFFF == zlib.crc32(b)
        # unsigned seed is accepted the same as the signed one
        assert O.crc32(b[cut:], O.crc32(b[:cut]) & 0xFFFFFFFF) == O.crc32(b)


def test_adler32_q1_defect():
    """Q1: lengths that are non-zero multiples of 5552 skip the final reduction of sum2."""
    rnd = random.Random(11)
    for n in [5552, 11104, 16656, 5552 * 7]:
        b = bytes(rnd.getrandbits(8) for _ in range(n))
        ref = O.adler32(b) & 0xFFFFFFFF
        assert ref != zlib.adler32(b)
        # model of the defect: low half is the true a, high half is the unreduced sum truncated to 16 bits
        a, s2 = 1, 0
        for k in range(0, n, 5552):
            for x in b[k:k + 5552]:
                a += x
                s2 += a
            a %= 65521
            s2 += 65521
        assert ref == (a | ((s2 & 0xFFFF) << 16))


# ---------------------------------------------------------------- vs zlib on generated valid streams

@pytest.mark.parametrize("level", [1, 6, 9])
def test_generated_vs_zlib(level):
    rnd = random.Random(level)
    words = [bytes(rnd.choice(b"etaoinshrdlu") for _ in range(rnd.randint(2, 9))) for _ in range(500)]
    for n in [0, 1, 10, 300, 5000, 70000, 200000]:
        data = b" ".join(rnd.choice(words) for _ in range(n // 5 + 1))[:n]
        z = zlib.compress(data, level)
        out, r = O.inflate_oneshot(z)
        if n == 0:
            assert r.checksum_state == 2      # Q8
            continue
        assert out == data and r.success and r.checksum_state == 1
        g = gzip.compress(data, level, mtime=12345)
        out, r = O.inflate_oneshot(g)
        assert out == data and r.success and r.size_state == 1 and r.mtime == 12345
        out, r = O.inflate_oneshot(raw_deflate(data, level), mode=O.MODE_RAW)
        assert out == data      # completeness of raw streams depends on Q15


def test_preset_dictionary():
    dic = b" ".join([b"the", b"and", b"of", b"to", b"heaven", b"earth", b"light"] * 12)
    data = b"of heaven and earth the light to the and of " * 50
    z = raw_deflate(data, 6, 15, dic)
    assert z[1] & 0x20
    out, r = O.inflate_oneshot(z, dictionary=dic)
    assert out == data and r.success
    out, r = O.inflate_oneshot(z)
    assert r.thrown_append == O.THROW_DICT_REQUIRED
    out, r = O.inflate_oneshot(z, dictionary=dic + b"x")
    assert r.thrown_append == O.THROW_DICT_INVALID


# ---------------------------------------------------------------- Appendix A vectors

def test_divergence_vectors():
    out, r = O.inflate_oneshot(bytes.fromhex("030200"))                       # D1 / Q6
    assert out == b"\0\0\0" and r.success and r.checksum_state == 0
    out, r = O.inflate_oneshot(bytes.fromhex("05c08100000000009056fe2b0000"))  # D2 / Q9
    assert set(out) == {ord("a")} and not r.complete and r.thrown_inflate == O.THROW_UNEXPECTED_EOF
    out, r = O.inflate_oneshot(bytes.fromhex("0dc081080000000020d6fd252e02"))  # D3 / Q9
    assert r.thrown_append == O.THROW_INFLATE_ERROR and O.MSG_TEXT[r.msg_id] == "empty distance tree with lengths"


def test_q2_stored_blocks():
    rnd = random.Random(3)
    b = bytes(rnd.getrandbits(8) for _ in range(49151))
    out, r = O.inflate_oneshot(zlib.compress(b, 6))
    assert out == b and r.success and r.n_blocks == 3
    b = bytes(rnd.getrandbits(8) for _ in range(49152))
    out, r = O.inflate_oneshot(zlib.compress(b, 6))
    assert not r.success


def test_q4_q5_q8_q15():
    out, r = O.inflate_oneshot(zlib.compress(b"hello world") + b"\0")          # Q4
    assert r.thrown_append == O.THROW_HANG
    g = bytearray(gzip.compress(b"hello", mtime=1))                             # Q5
    g[3] |= 4
    g[10:10] = b"\x02\x00ab"
    out, r = O.inflate_oneshot(bytes(g))
    assert out == b"" and not r.complete and r.thrown_inflate == O.THROW_UNEXPECTED_EOF and r.container == 2
    out, r = O.inflate_oneshot(zlib.compress(b""))                             # Q8
    assert r.complete and r.checksum_state == 2 and r.thrown_inflate == O.THROW_INTEGRITY
    out, r = O.inflate_oneshot(raw_deflate(b"hello hello hello"))             # Q15: 0 padding bits after EOB
    assert out == b"hello hello hello" and not r.complete


def test_error_messages():
    cases = [
        (b"\x78\x02" + b"\x03\x00", "incorrect header check"),
        (b"\x79\x9c" + b"\x03\x00", "unknown compression method"),
        (b"\x88\x1c" + b"\x03\x00", "invalid window size"),
        (b"\x1f\x8c\x08\x00", "invalid gzip id"),
        (b"\x78\x01\x07", "invalid block type"),
        (b"\x78\x01\x01\x05\x00\x00\x00", "invalid stored block lengths"),
    ]
    for data, msg in cases:
        out, r = O.inflate_oneshot(data, mode=O.MODE_INFLATER)
        assert r.thrown_append == O.THROW_INFLATE_ERROR and O.MSG_TEXT[r.msg_id] == msg, (data, r.observable())
    out, r = O.inflate_oneshot(b"x")
    assert r.thrown_inflate == O.THROW_TOO_SMALL


def test_q15_raw_stream_final_end_of_block_needs_lookahead():
    """the reference looks a code up only when its table's index width is in the bit buffer, and nothing
    follows the end-of-block code of a raw stream.  A stream whose last code is shorter than that width therefore never
    completes: every data byte is delivered, z stays Z_OK, finish() says complete = false and inflate() throws
    "Unexpected EOF" - while zlib decodes the same bytes to the end.  (The GPU fast path writes this record itself:
    tests/test_gpu_fast.py::test_raw_streams_the_reference_leaves_incomplete_stay_on_the_fast_path.)"""
    import zlib
    from tools import corpus as K
    seen = 0
    for i in range(600):
        plain = K.generate(K.TEXT if i % 2 else K.BINARY, 20000 + i, 200 + (i * 37) % 900)
        s = K.compress(plain, (1, 6, 9)[i % 3], K.RAW)
        out, r = O.inflate_oneshot(bytes(s), mode=O.MODE_RAW)
        assert zlib.decompress(bytes(s), -15) == bytes(plain)
        if r.complete:
            assert r.zstatus == 1 and r.success == 1 and out == bytes(plain)
            continue
        seen += 1
        assert out == bytes(plain) and r.out_len == len(plain)          # all data delivered ...
        assert r.zstatus == 0 and r.total_in == len(s) and r.thrown_append == 0 and r.success == 0
        assert r.thrown_inflate == O.THROW_UNEXPECTED_EOF                  # ... but inflate() throws
    assert seen >= 3
