"""Hand-made deflate blocks for crafted parity vectors (test infrastructure).

zlib never emits the code sets these tests need (table arenas of exactly 1399 / 1400 / 1401 entries, a lone 1-bit
code-length code, two-symbol blocks that end a few bytes before the end of the input), so the blocks are written bit
by bit here.  RFC 1951 section 3.2: bits are packed LSB first, Huffman codes MSB first."""

BORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
             8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


class BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def bits(self, value, count):
        """`count` bits of `value`, least significant first"""
        self.acc |= (value & ((1 << count) - 1)) << self.n
        self.n += count
        while self.n >= 8:
            self.out.append(self.acc & 0xff)
            self.acc >>= 8
            self.n -= 8

    def code(self, code, length):
        """a Huffman code: most significant bit first"""
        for i in range(length - 1, -1, -1):
            self.bits((code >> i) & 1, 1)

    def align(self):
        if self.n:
            self.bits(0, 8 - self.n)

    def bit_length(self):
        return len(self.out) * 8 + self.n

    def done(self):
        self.align()
        return bytes(self.out)


def canonical_codes(lens):
    """symbol -> (code, length) for a list of code lengths (RFC 1951 3.2.2); codes are assigned even when the set is
    over- or under-subscribed (the decoder's acceptance is what the tests are about)"""
    max_len = max(lens) if lens else 0
    bl_count = [0] * (max_len + 2)
    for l in lens:
        if l:
            bl_count[l] += 1
    code, next_code = 0, [0] * (max_len + 2)
    for b in range(1, max_len + 1):
        code = (code + bl_count[b - 1]) << 1
        next_code[b] = code
    out = {}
    for s, l in enumerate(lens):
        if l:
            out[s] = (next_code[l], l)
            next_code[l] += 1
    return out


def length_symbol(length):
    for i in range(28, -1, -1):
        if length >= LEN_BASE[i]:
            return 257 + i, length - LEN_BASE[i], LEN_EXTRA[i]
    raise ValueError(length)


def dist_symbol(dist):
    for i in range(29, -1, -1):
        if dist >= DIST_BASE[i]:
            return i, dist - DIST_BASE[i], DIST_EXTRA[i]
    raise ValueError(dist)


def write_dynamic_block(w, lit_lens, dist_lens, symbols, final=True, cl_lens=None, rle=True, cl_bit_noise=None):
    """One dynamic block.  lit_lens: HLIT + 257 code lengths, dist_lens: HDIST + 1 code lengths.
    symbols: list of ints (literal bytes), ('m', length, distance) tuples and 256 for end of block (not added here).
    cl_lens: code-length-code lengths (19, by symbol); default: every length that occurs gets a 4- or 5-bit code...
    rle: use symbol 18 / 17 for runs of zeros.  cl_bit_noise: for a lone 1-bit code-length code, a callable giving the
    bit actually written for each code-length symbol (both patterns decode to the symbol in the reference)."""
    nl, nd = len(lit_lens), len(dist_lens)
    assert 257 <= nl <= 286 and 1 <= nd <= 30
    seq = list(lit_lens) + list(dist_lens)
    # code-length symbols
    cls = []
    i = 0
    while i < len(seq):
        if rle and seq[i] == 0:
            j = i
            while j < len(seq) and seq[j] == 0:
                j += 1
            run = j - i
            while run >= 11:
                r = min(run, 138)
                cls.append((18, r - 11, 7))
                run -= r
            if run >= 3:
                cls.append((17, run - 3, 3))
                run = 0
            cls.extend([(0, 0, 0)] * run)
            i = j
        else:
            cls.append((seq[i], 0, 0))
            i += 1
    if cl_lens is None:
        used = sorted({c[0] for c in cls})
        # a complete code over the used symbols: lengths from a balanced tree
        cl_lens = [0] * 19
        n = len(used)
        if n == 1:
            used = used + [(used[0] + 1) % 19]
            n = 2
        depth = 1
        while (1 << depth) < n:
            depth += 1
        short = (1 << depth) - n            # symbols that get depth - 1 bits
        for k, s in enumerate(used):
            cl_lens[s] = depth - 1 if k < short else depth
        assert max(cl_lens) <= 7
    cl_codes = canonical_codes(cl_lens)
    hclen = 19
    while hclen > 4 and cl_lens[BORDER[hclen - 1]] == 0:
        hclen -= 1
    w.bits(1 if final else 0, 1)
    w.bits(2, 2)
    w.bits(nl - 257, 5)
    w.bits(nd - 1, 5)
    w.bits(hclen - 4, 4)
    for k in range(hclen):
        w.bits(cl_lens[BORDER[k]], 3)
    for n_sym, (s, extra, xbits) in enumerate(cls):
        code, length = cl_codes[s]
        if cl_bit_noise is not None:
            assert length == 1
            w.bits(cl_bit_noise(n_sym), 1)
        else:
            w.code(code, length)
        if xbits:
            w.bits(extra, xbits)
    lit_codes, dist_codes = canonical_codes(list(lit_lens)), canonical_codes(list(dist_lens))
    for s in symbols:
        if isinstance(s, tuple):
            _, length, dist = s
            ls, lx, lxb = length_symbol(length)
            w.code(*lit_codes[ls])
            if lxb:
                w.bits(lx, lxb)
            ds, dx, dxb = dist_symbol(dist)
            w.code(*dist_codes[ds])
            if dxb:
                w.bits(dx, dxb)
        else:
            w.code(*lit_codes[s])
    w.code(*lit_codes[256])


def write_stored_block(w, data, final=True):
    w.bits(1 if final else 0, 1)
    w.bits(0, 2)
    w.align()
    w.bits(len(data), 16)
    w.bits(len(data) ^ 0xffff, 16)
    for b in data:
        w.bits(b, 8)


def kraft_complete_lens(n_symbols, max_len, rng, must_have=()):
    """a random COMPLETE code over some of n_symbols symbols (Kraft sum exactly 1, lengths <= max_len): repeatedly
    split a leaf.  must_have: symbols that must get a code."""
    leaves = [1, 1]
    target = rng.randint(max(2, len(must_have)), n_symbols)
    while len(leaves) < target:
        cand = [i for i, l in enumerate(leaves) if l < max_len]
        if not cand:
            break
        i = rng.choice(cand)
        l = leaves.pop(i)
        leaves += [l + 1, l + 1]
    lens = [0] * n_symbols
    syms = list(must_have)
    rest = [s for s in range(n_symbols) if s not in must_have]
    rng.shuffle(rest)
    syms += rest[:len(leaves) - len(syms)]
    rng.shuffle(leaves)
    for s, l in zip(syms, leaves):
        lens[s] = l
    return lens


def zlib_wrap(payload, plain):
    import zlib
    return b"\x78\x01" + payload + zlib.adler32(plain).to_bytes(4, "big")


def gzip_wrap(payload, plain):
    import zlib
    return (b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff" + payload + (zlib.crc32(plain) & 0xffffffff).to_bytes(4, "little")
            + (len(plain) & 0xffffffff).to_bytes(4, "little"))
