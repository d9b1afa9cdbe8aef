"""the corpus tool (tools/corpus.c = system zlib 1.3 inside the reference Deflater's wrappers,
src/sd-deflate.ts:98-165) stands in for "the reference's own deflate".  The reference publishes exactly two things
about its compressor's output: the fixture test/paradiselost.deflate (level 6) and the per-level sizes of the same
text in test/perf.html:63-69.  Both are pinned here."""
import zlib

import numpy as np
import pytest

from tools import corpus as K

# test/perf.html:63-69 - zlib-wrapped size of deflate(paradiselost.txt) per level
PERF_HTML_SIZES = {9: 193162, 6: 193730, 5: 197239, 4: 203828, 3: 207545, 2: 216830, 1: 226188}


def test_level6_reproduces_the_reference_fixture_byte_for_byte(fx):
    text = fx("paradiselost.txt")
    assert K.compress(text, 6, K.ZLIB) == fx("paradiselost.deflate")


@pytest.mark.parametrize("level", sorted(PERF_HTML_SIZES))
def test_level_sizes_of_perf_html(fx, level):
    text = fx("paradiselost.txt")
    z = K.compress(text, level, K.ZLIB)
    assert len(z) == PERF_HTML_SIZES[level]
    assert z[:2] == b"\x78\x01"                                   # src/sd-deflate.ts:98-115: always 78 01
    assert zlib.decompress(z) == text


def test_gzip_wrapper_is_the_reference_writers(fx):
    """src/sd-deflate.ts:117-152: ID 1f 8b, CM 8, FLG = FNAME or 0, XFL 0, OS 0xff; the payload is the level-6 fixture's."""
    text = fx("paradiselost.txt")
    g = K.compress(text, 6, K.GZIP)
    assert g[:4] == b"\x1f\x8b\x08\x00" and g[8] == 0 and g[9] == 0xFF
    assert g[10:-8] == fx("paradiselost.deflate")[2:-4]
    assert int.from_bytes(g[-8:-4], "little") == zlib.crc32(text) and int.from_bytes(g[-4:], "little") == len(text)
    gn = K.compress(text, 6, K.GZIP_NAME)
    assert gn[3] == 0x08 and gn[10:].index(0) > 0


def test_generators_are_deterministic_and_sized():
    for kind in (K.TEXT, K.BINARY, K.RANDOM, K.RUNS):
        a, b = K.generate(kind, 5, 65536), K.generate(kind, 5, 65536)
        assert a.size == 65536 and np.array_equal(a, b)
        assert not np.array_equal(a, K.generate(kind, 6, 65536))
