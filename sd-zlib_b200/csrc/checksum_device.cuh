// checksum_device.cuh - device-side Adler-32 / CRC-32 building blocks shared by the
// per-stream record finaliser and the large-buffer checksum kernels.
//
// Reference semantics reproduced here:
//   computeAdler32   src/adler32.ts:34-105  (NMAX = 5552 blocking; the missing sum2
//                                            reduction when len is a non-zero multiple of
//                                            5552 - SURVEY Q1 - is modelled exactly)
//   computeCRC32     src/crc32.ts:48-106    (IEEE reflected CRC-32, seed = previous value)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdz {

constexpr uint32_t ADLER_BASE = 65521u;
constexpr uint32_t ADLER_NMAX = 5552u;
constexpr uint32_t CRC_POLY = 0xedb88320u;

// x^(2^n) mod P for n = 0..31 (reflected), filled by the host at context creation
__device__ __constant__ uint32_t c_x2n[32];
// CRC byte table (slice-by-1) in global memory; kernels copy it to shared memory
__device__ uint32_t g_crc_tab[4][256];

// a(x) * b(x) mod P, reflected representation (same recurrence as zlib's multmodp)
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    #pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : (b >> 1);
    }
    return p;
}

// x^(n * 2^k) mod P
__device__ __forceinline__ uint32_t gf2_x2n(uint64_t n, uint32_t k)
{
    uint32_t p = 0x80000000u;
    while (n) {
        if (n & 1) p = gf2_mulmod(c_x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

// advance a CRC register over `nbytes` zero bytes (the shift operator of crc32_combine)
__device__ __forceinline__ uint32_t crc_shift(uint32_t reg, uint64_t nbytes)
{
    if (nbytes == 0 || reg == 0) return reg;
    return gf2_mulmod(gf2_x2n(nbytes, 3), reg);
}

// Running Adler-32 state across the units (<= 5552 bytes) of ONE adler32() call.
struct AdlerCall {
    uint32_t a;        // low sum as the reference holds it entering the next unit
    uint64_t b;        // high sum reduced mod BASE (valid when a tail exists)
    uint32_t sq;       // unreduced high sum mod 2^32 (what `sum2 << 16` keeps when no tail exists)
    uint32_t b0;       // seed high half
    bool any, all_full;
    __device__ __forceinline__ void begin(uint32_t seed)
    {
        a = seed & 0xffffu; b0 = (seed >> 16) & 0xffffu; b = b0 % ADLER_BASE; sq = b0; any = false; all_full = true;
    }
    // S = sum of the unit's bytes, W = sum of (L - i) * d_i over the unit, L = unit length
    __device__ __forceinline__ void unit(uint32_t S, uint32_t W, uint32_t L)
    {
        any = true;
        if (L != ADLER_NMAX) all_full = false;
        sq += ADLER_NMAX * a + W + ADLER_BASE;                          // src/adler32.ts:47-68
        b = (b + (uint64_t)L * a + W) % ADLER_BASE;
        a = (uint32_t)(((uint64_t)a + S) % ADLER_BASE);
    }
    __device__ __forceinline__ uint32_t end(bool standard = false) const
    {
        if (!any) return a | (b0 << 16);                                // empty buffer: seed recombined (:104)
        uint32_t hi = (all_full && !standard) ? (sq & 0xffffu) : (uint32_t)b;   // Q1: no `sum2 %= BASE` without a tail
        return a | (hi << 16);
    }
};

// warp-cooperative (S, T = sum of i * d_i) over `L` bytes at p; every lane returns the totals
__device__ __forceinline__ void adler_unit_sums(const uint8_t* p, uint32_t L, uint32_t lane, uint32_t& S, uint32_t& T)
{
    uint32_t s = 0, t = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        uint32_t nv = L >> 4;
        for (uint32_t v = lane; v < nv; v += 32) {
            uint4 q = __ldg(p4 + v);
            uint32_t s16 = __dp4a(q.x, 0x01010101u, 0u);
            s16 = __dp4a(q.y, 0x01010101u, s16);
            s16 = __dp4a(q.z, 0x01010101u, s16);
            s16 = __dp4a(q.w, 0x01010101u, s16);
            uint32_t j16 = __dp4a(q.x, 0x03020100u, 0u);
            j16 = __dp4a(q.y, 0x07060504u, j16);
            j16 = __dp4a(q.z, 0x0b0a0908u, j16);
            j16 = __dp4a(q.w, 0x0f0e0d0cu, j16);
            s += s16;
            t += (v << 4) * s16 + j16;
        }
        for (uint32_t i = (nv << 4) + lane; i < L; i += 32) { uint32_t d = p[i]; s += d; t += i * d; }
    } else {
        for (uint32_t i = lane; i < L; i += 32) { uint32_t d = p[i]; s += d; t += i * d; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    S = s; T = t;
}

// Adler-32 of one adler32(buf, seed) call evaluated by a warp (all lanes return the value)
// (standard = true: RFC 1950, without the reference's Q1 defect)
__device__ __forceinline__ uint32_t adler_call_warp(const uint8_t* p, uint64_t n, uint32_t seed, uint32_t lane, bool standard = false)
{
    AdlerCall st;
    st.begin(seed);
    for (uint64_t off = 0; off < n; off += ADLER_NMAX) {
        uint32_t L = (uint32_t)min((uint64_t)ADLER_NMAX, n - off);
        uint32_t S, T;
        adler_unit_sums(p + off, L, lane, S, T);
        uint32_t W = (uint32_t)((uint64_t)L * S - T);
        st.unit(S, W, L);
    }
    return st.end(standard);
}

// raw CRC register update over bytes, slice-by-1 with a shared-memory table
__device__ __forceinline__ uint32_t crc_bytes(const uint32_t* tab0, uint32_t c, const uint8_t* p, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++) c = tab0[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c;
}

// standard crc32(buf, seed) by a warp: 32 contiguous slices, joined with the GF(2) shift
// operator (crc32_combine).  tab = 4 x 256 slicing tables in shared memory.
__device__ __forceinline__ uint32_t crc_call_warp(const uint32_t* tab, const uint8_t* p, uint64_t n, uint32_t seed, uint32_t lane)
{
    uint64_t slice = ((n + 31) / 32 + 3) & ~3ull;
    uint64_t lo = min(n, (uint64_t)lane * slice), hi = min(n, lo + slice);
    uint32_t c = 0;                                     // zero-init remainder G(slice)
    const uint8_t* q = p + lo;
    uint64_t len = hi - lo;
    // head bytes up to 4-byte alignment, then slicing-by-4 words, then tail bytes
    while (len && (reinterpret_cast<uintptr_t>(q) & 3)) { c = tab[(c ^ *q++) & 0xffu] ^ (c >> 8); len--; }
    const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
    uint64_t nw = len >> 2;
    for (uint64_t i = 0; i < nw; i++) {
        c ^= q4[i];
        c = tab[768 + (c & 0xffu)] ^ tab[512 + ((c >> 8) & 0xffu)] ^ tab[256 + ((c >> 16) & 0xffu)] ^ tab[c >> 24];
    }
    q += nw << 2; len &= 3;
    while (len--) c = tab[(c ^ *q++) & 0xffu] ^ (c >> 8);
    uint32_t x = crc_shift(c, n - hi);
    if (lane == 0) x ^= crc_shift(~seed, n);            // F(c0, s) = shift(c0, |s|) ^ G(s)
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, o);
    return ~x;
}

}  // namespace sdz
