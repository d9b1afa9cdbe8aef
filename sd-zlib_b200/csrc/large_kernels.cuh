// large_kernels.cuh - single large stream (BASELINE config 5): two-pass decode.
//
//   pass 1a  prefilter_headers +    every bit position of the deflate payload is tested, in parallel,
//            verify_headers         for "a dynamic block header that the reference would accept starts
//                                   here" (src/infblocks.ts:334-523 + the tree rules of
//                                   src/inftree.ts:313-379) -> candidate block starts
//   pass 1b  inflate_kernel<4,false,TM_INDEX>: count-only walk of every candidate block: its extent (end
//            bit, output bytes) and a resume point every 16 KiB of output; the host then follows the chain
//            of blocks from the first one (false candidates are never reached), cuts the blocks into
//            pieces at the resume points and lays the pieces out in the output
//   pass 2a  inflate_kernel<4,true,TM_MARK>: every piece is decoded WITHOUT its 32 KiB window into
//            16-bit symbols: a byte, or 256 + index into the window before the piece
//   pass 2b-d window propagation     (see below) and marker resolution
//
// Nothing here exists in the reference (it streams through one 32 KiB window, strictly sequentially).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdz {

// 32 bits of the stream starting at absolute bit position p (LSB first).  The buffer must be readable
// for 12 bytes past the last byte (the device arenas have SDZ_IN_PAD bytes of slack).
__device__ __forceinline__ uint32_t bits32_at(const uint32_t* src32, uint64_t p)
{
    const uint64_t w = p >> 5;
    const uint32_t lo = __ldg(src32 + w), hi = __ldg(src32 + w + 1);
    return __funnelshift_r(lo, hi, (uint32_t)(p & 31));
}

// reference acceptance of one code-length set given its per-length counts (src/inftree.ts:131-178,:298):
// 0 ok, 1 oversubscribed, 2 incomplete, 3 empty
// reference acceptance of one code-length set (src/inftree.ts:131-178,:298) from its Kraft sum `kr` (in units of
// 2^-15, exact for lengths <= 15) and its longest code: 0 ok, 1 oversubscribed, 2 incomplete, 3 empty.
// (huft_build's level-by-level "y < 0" test fires iff the total exceeds 1; "y != 0 && g != 1" is "not complete,
// unless all codes have length 1".)
__device__ __forceinline__ int classify_kraft(uint32_t kr, uint32_t max_len)
{
    if (kr == 0u) return 3;
    if (kr > 32768u) return 1;
    return (kr != 32768u && max_len != 1u) ? 2 : 0;
}

__device__ __forceinline__ uint64_t bits64_at(const uint32_t* src32, uint64_t p)
{
    const uint64_t w = p >> 5;
    const uint32_t a = __ldg(src32 + w), b = __ldg(src32 + w + 1), c = __ldg(src32 + w + 2);
    const uint32_t sh = (uint32_t)(p & 31);
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}

// Does the reference accept a dynamic block header at bit p?  (src/infblocks.ts:334-523 with the tree rules of
// src/inftree.ts:131-178,:298-379.)  Registers only and no data-dependent branches inside a step, so the 32 candidates of
// a warp stay converged: the code-length code is decoded canonically (per-length symbol masks + find-nth-set-bit)
// instead of through a per-thread table, and the two big trees are judged by their Kraft sums, so no per-length
// counters are kept.  A false candidate is random data: its lengths oversubscribe a tree within a few dozen symbols
// and the walk stops there.
__device__ bool is_dynamic_header(const uint32_t* src32, uint64_t p, uint64_t total_bits)
{
    if (p + 17 + 12 > total_bits) return false;
    const uint32_t h = bits32_at(src32, p);
    if (((h >> 1) & 3u) != 2u) return false;                              // BTYPE = dynamic
    const uint32_t hlit = (h >> 3) & 31u, hdist = (h >> 8) & 31u, hclen = (h >> 13) & 15u;
    if (hlit > 29u || hdist > 29u) return false;                          // src/infblocks.ts:355
    const uint32_t ncl = 4u + hclen;
    const int nl = 257 + (int)hlit, nd = 1 + (int)hdist;
    uint64_t q = p + 17;
    if (q + 3ull * ncl > total_bits) return false;
    // the 3-bit lengths in transmission order (fields 0..9 in x0, 10..18 in x1) -> symbol order (y0: symbols 0..9, y1: 10..18)
    const uint64_t v = bits64_at(src32, q) & ((1ull << (3u * ncl)) - 1ull);
    q += 3ull * ncl;
    const uint32_t x0 = (uint32_t)v & 0x3fffffffu, x1 = (uint32_t)(v >> 30);
    uint32_t y0 = 0, y1 = 0;
    {
        constexpr int border[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
        #pragma unroll
        for (int i = 0; i < 19; i++) {
            const uint32_t f = ((i < 10 ? x0 : x1) >> (3 * (i % 10))) & 7u;
            const int sy = border[i];
            if (sy < 10) y0 |= f << (3 * sy); else y1 |= f << (3 * (sy - 10));
        }
    }
    // per-length symbol masks (bit 3 s of m0 / bit 3 (s - 10) of m1), counts and first canonical codes
    constexpr uint32_t M = 0x09249249u;
    const uint32_t a0 = y0 & M, a1 = (y0 >> 1) & M, a2 = (y0 >> 2) & M, c0 = y1 & M, c1 = (y1 >> 1) & M, c2 = (y1 >> 2) & M;
    uint32_t m0[8], m1[8], cnt[8], first[8];
    #pragma unroll
    for (int L = 1; L <= 7; L++) {
        m0[L] = ((L & 1) ? a0 : (a0 ^ M)) & ((L & 2) ? a1 : (a1 ^ M)) & ((L & 4) ? a2 : (a2 ^ M));
        m1[L] = ((L & 1) ? c0 : (c0 ^ M)) & ((L & 2) ? c1 : (c1 ^ M)) & ((L & 4) ? c2 : (c2 ^ M));
        cnt[L] = (uint32_t)(__popc(m0[L]) + __popc(m1[L]));
    }
    uint32_t kraft = 0, n_codes = 0, code = 0;
    #pragma unroll
    for (int L = 1; L <= 7; L++) {
        kraft += cnt[L] << (7 - L);
        n_codes += cnt[L];
        first[L] = code;
        code = (code + cnt[L]) << 1;
    }
    const bool single = n_codes == 1u && cnt[1] == 1u;                    // one code of length 1: both bit values decode to it (Q11)
    if (kraft != 128u && !single) return false;
    // the nl + nd code lengths (run-length coded)
    const int total = nl + nd;
    int index = 0;
    uint32_t prev = 0, krL = 0, krD = 0, maxL = 0, maxD = 0;
    while (index < total) {
        if (q + 14 > total_bits) return false;
        const uint32_t w = bits32_at(src32, q);
        const uint32_t x = __brev(w) >> 25;                               // the next 7 bits, first bit most significant
        uint32_t tb = 1, j = 0, s0 = m0[1], s1 = m1[1];
        bool found = single;
        #pragma unroll
        for (int L = 1; L <= 7; L++) {
            const uint32_t d = (x >> (7 - L)) - first[L];
            const bool hit = !found && d < cnt[L];
            if (hit) { tb = L; j = d; s0 = m0[L]; s1 = m1[L]; }
            found = found || hit;
        }
        if (!found) return false;                                         // (not reachable for a complete code)
        const uint32_t in0 = (uint32_t)__popc(s0);
        const uint32_t bitpos = j < in0 ? __fns(s0, 0, (int)j + 1) : __fns(s1, 0, (int)(j - in0) + 1);
        const uint32_t c = bitpos / 3u + (j < in0 ? 0u : 10u);            // the symbol: 0..18
        uint32_t rep = 1, val = c;
        if (c >= 16u) {
            const uint32_t xb = c == 18u ? 7u : c - 14u;
            rep = (c == 18u ? 11u : 3u) + ((w >> tb) & ((1u << xb) - 1u));
            tb += xb;
            if ((c == 16u && index < 1) || index + (int)rep > total) return false;
            val = c == 16u ? prev : 0u;
        }
        q += tb;
        prev = val;
        const uint32_t toL = index < nl ? min(rep, (uint32_t)(nl - index)) : 0u, toD = rep - toL;
        const uint32_t kr = val ? (32768u >> val) : 0u;
        krL += toL * kr; krD += toD * kr;
        if (toL && val > maxL) maxL = val;
        if (toD && val > maxD) maxD = val;
        index += (int)rep;
        if (krL > 32768u || krD > 32768u) return false;                   // oversubscribed: src/inftree.ts:146-166
    }
    // literal/length and distance trees (inflate_trees_dynamic); the MANY arena limit is ignored here: a block rejected
    // only by that rule is still a block boundary for the index
    if (classify_kraft(krL, maxL) != 0) return false;
    const int sd = classify_kraft(krD, maxD);
    if (sd == 1 || sd == 2) return false;
    if (sd == 3 && nl > 257) return false;
    return true;
}

// Necessary condition, in registers only: the 4 + HCLEN code-length-code lengths form a complete code
// (Kraft sum exactly 1) or the one-code special case.  The 3-bit fields are counted per value with bit planes and
// population counts instead of nineteen extract-shift-add steps.
__device__ __forceinline__ bool precode_plausible(const uint32_t* src32, uint64_t p, uint32_t h)
{
    const uint32_t ncl = 4u + ((h >> 13) & 15u);
    const uint64_t v = bits64_at(src32, p + 17) & ((1ull << (3u * ncl)) - 1ull);
    constexpr uint64_t M = 0x1249249249249249ull;                      // bit 0 of every 3-bit field
    const uint64_t b0 = v & M, b1 = (v >> 1) & M, b2 = (v >> 2) & M;
    const uint64_t n0 = b0 ^ M, n1 = b1 ^ M, n2 = b2 ^ M;
    uint32_t k = 0;
    k += (uint32_t)__popcll(b0 & n1 & n2) << 6;                         // length 1: 2^-1 of 128
    k += (uint32_t)__popcll(n0 & b1 & n2) << 5;
    k += (uint32_t)__popcll(b0 & b1 & n2) << 4;
    k += (uint32_t)__popcll(n0 & n1 & b2) << 3;
    k += (uint32_t)__popcll(b0 & n1 & b2) << 2;
    k += (uint32_t)__popcll(n0 & b1 & b2) << 1;
    k += (uint32_t)__popcll(b0 & b1 & b2);
    return k == 128u || k == 64u;
}

// pass 1a, stage 1.  Every lane tests 32 consecutive bit positions at once with shifted copies of a 64-bit window:
// BTYPE = 10b, HLIT <= 29, HDIST <= 29 (22 % pass).  The positions that pass are queued per warp so that the Kraft test
// runs on full warps; its survivors (0.14 %) go to surv[] for stage 2.
constexpr int PF_TILE = 4096;                                          // positions per warp and tile (4 rounds of 1024)
__global__ void __launch_bounds__(256) prefilter_headers(const uint8_t* src, uint64_t first_bit, uint64_t end_bit, uint64_t total_bits,
                                                         uint64_t* surv, unsigned long long* n_surv, unsigned long long cap)
{
    __shared__ uint16_t queue_all[8][1024 + 32];
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint16_t* q = queue_all[warp];
    const uint64_t n_pos = end_bit - first_bit;                      // positions [first_bit, end_bit) are tested
    const uint64_t warps_total = (uint64_t)gridDim.x * 8u;
    auto kraft = [&](uint64_t p) {
        const uint32_t h = bits32_at(src32, p);
        if (precode_plausible(src32, p, h)) {
            const unsigned long long slot = atomicAdd(n_surv, 1ull);
            if (slot < cap) surv[slot] = p;
        }
    };
    for (uint64_t tile = (uint64_t)blockIdx.x * 8u + warp; tile * PF_TILE < n_pos; tile += warps_total) {
        const uint64_t base = first_bit + tile * PF_TILE;
        uint32_t qn = 0;
        for (uint32_t r = 0; r < PF_TILE / 1024; r++) {
            const uint32_t rel0 = r * 1024u + lane * 32u;
            const uint64_t p0 = base + rel0;
            uint32_t m = 0;
            if (p0 < end_bit && p0 + 17 + 12 <= total_bits) {
                const uint64_t w = bits64_at(src32, p0);
                const uint64_t t = ~(w >> 1) & (w >> 2)                                       // BTYPE: low bit 0, high bit 1
                                   & ~((w >> 4) & (w >> 5) & (w >> 6) & (w >> 7))             // HLIT  < 30
                                   & ~((w >> 9) & (w >> 10) & (w >> 11) & (w >> 12));         // HDIST < 30
                m = (uint32_t)t;
                // positions of this lane beyond the scan range or too close to the end of the stream
                const uint64_t lim = end_bit < total_bits - 28 ? end_bit : total_bits - 28;   // p + 29 <= total_bits
                if (p0 + 32 > lim) m = lim > p0 ? m & (uint32_t)((1ull << (lim - p0)) - 1ull) : 0u;
            }
            // warp-wide exclusive prefix of the per-lane counts
            const uint32_t c = (uint32_t)__popc(m);
            uint32_t inc = c;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (uint32_t)d) inc += up;
            }
            uint32_t at = qn + inc - c;
            while (m) {
                const uint32_t i = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                q[at++] = (uint16_t)(rel0 + i);
            }
            qn += __shfl_sync(0xffffffffu, inc, 31);
            __syncwarp();
            while (qn >= 32u) {
                qn -= 32u;
                kraft(base + q[qn + lane]);
            }
            __syncwarp();
        }
        if (lane < qn) kraft(base + q[lane]);
        __syncwarp();
    }
}

// pass 1a, stage 2: the full header check for the survivors
__global__ void __launch_bounds__(128) verify_headers(const uint8_t* src, uint64_t total_bits, const uint64_t* surv,
                                                      const unsigned long long* n_surv, unsigned long long cap, uint64_t* cand,
                                                      unsigned long long* n_cand, unsigned long long max_cand)
{
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    const unsigned long long n = *n_surv < cap ? *n_surv : cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t p = surv[i];
        if (is_dynamic_header(src32, p, total_bits)) {
            const unsigned long long slot = atomicAdd(n_cand, 1ull);
            if (slot < max_cand) cand[slot] = p;
        }
    }
}

// ---- window propagation.  A block's marker 256 + i stands for "byte i of the 32 KiB before this block".
// Making every window final strictly block after block would be a chain of nb dependent steps; instead
// the blocks are grouped into segments of `bps` consecutive blocks:
//   pass 2b  propagate_in_segment   all segments in parallel, blocks of a segment in order: the LAST 32 KiB
//                                   of every block are rewritten so that their markers refer to the window
//                                   before the SEGMENT (composition of the per-block gather maps)
//   pass 2c  propagate_segments     one CTA, segments in order: the last 32 KiB of every segment become
//                                   final bytes
//   pass 2d  resolve_markers        everything else, fully parallel (at most two gathers per symbol)
// blk_off[b] = absolute output offset of block b (blk_off[nb] = total).  All offsets are computed modulo
// 2^64, so "start - 32768 + i" is fine for blocks near the start of the stream (the decoder never emits a
// marker for a position before the stream: those are literal zeros).
constexpr uint32_t WIN = 32768u;

template <typename F>
__device__ __forceinline__ void for_tail(uint64_t lo, uint64_t end, F&& f)
{
    // 8 independent elements per thread and round, so the dependent gathers of a round overlap
    for (uint64_t x0 = lo + threadIdx.x; x0 < end; x0 += 8ull * blockDim.x) {
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint64_t x = x0 + (uint64_t)k * blockDim.x;
            if (x < end) f(x, k, 0);
        }
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint64_t x = x0 + (uint64_t)k * blockDim.x;
            if (x < end) f(x, k, 1);
        }
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint64_t x = x0 + (uint64_t)k * blockDim.x;
            if (x < end) f(x, k, 2);
        }
    }
}

__global__ void __launch_bounds__(1024) propagate_in_segment(uint16_t* sym, const uint64_t* blk_off, uint64_t nb, uint64_t bps)
{
    const uint64_t b0 = (uint64_t)blockIdx.x * bps, b1 = b0 + bps < nb ? b0 + bps : nb;
    if (b0 >= nb) return;
    const uint64_t seg_start = blk_off[b0];
    for (uint64_t b = b0; b < b1; b++) {
        const uint64_t start = blk_off[b], end = blk_off[b + 1];
        const uint64_t lo = end - start > WIN ? end - WIN : start;
        uint32_t v[8];
        for_tail(lo, end, [&](uint64_t x, int k, int stage) {
            if (stage == 0) v[k] = sym[x];
            else if (stage == 1) {
                if (v[k] >= 256u) {
                    const uint64_t p = start - WIN + (v[k] - 256u);       // absolute position the marker stands for
                    v[k] = p < seg_start ? 256u + (uint32_t)(p - (seg_start - WIN)) : (uint32_t)sym[p] | 0x10000u;
                } else v[k] = 0xffffffffu;                               // literal: nothing to write
            } else if (v[k] != 0xffffffffu) sym[x] = (uint16_t)v[k];
        });
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) propagate_segments(const uint16_t* sym, uint8_t* out, const uint64_t* blk_off, uint64_t nb,
                                                           uint64_t bps)
{
    for (uint64_t b0 = 0; b0 < nb; b0 += bps) {
        const uint64_t b1 = b0 + bps < nb ? b0 + bps : nb;
        const uint64_t seg_start = blk_off[b0], seg_end = blk_off[b1];
        const uint64_t lo = seg_end - seg_start > WIN ? seg_end - WIN : seg_start;
        uint32_t v[8];
        for_tail(lo, seg_end, [&](uint64_t x, int k, int stage) {
            if (stage == 0) v[k] = sym[x];
            else if (stage == 1) { if (v[k] >= 256u) v[k] = out[seg_start - WIN + (v[k] - 256u)]; }
            else out[x] = (uint8_t)v[k];
        });
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) resolve_markers(const uint16_t* sym, uint8_t* out, const uint64_t* blk_off, uint64_t nb, uint64_t bps)
{
    for (uint64_t b = blockIdx.y; b < nb; b += gridDim.y) {
        const uint64_t b0 = b / bps * bps, b1 = b0 + bps < nb ? b0 + bps : nb;
        const uint64_t seg_start = blk_off[b0], seg_end = blk_off[b1];
        const uint64_t start = blk_off[b], end = blk_off[b + 1];
        const uint64_t fin = seg_end - seg_start > WIN ? seg_end - WIN : seg_start;    // [fin, seg_end) is final already (pass 2c)
        const uint64_t hi = end < fin ? end : fin;
        const uint64_t tail = end - start > WIN ? end - WIN : start;                   // [tail, end) was rewritten by pass 2b
        for (uint64_t x = start + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < hi; x += (uint64_t)gridDim.x * blockDim.x) {
            uint32_t s = sym[x];
            if (s >= 256u) {
                if (x >= tail) s = out[seg_start - WIN + (s - 256u)];
                else {
                    const uint64_t p = start - WIN + (s - 256u);
                    if (p < seg_start) s = out[p];
                    else {
                        s = sym[p];
                        if (s >= 256u) s = out[seg_start - WIN + (s - 256u)];
                    }
                }
            }
            out[x] = (uint8_t)s;
        }
    }
}

}  // namespace sdz
