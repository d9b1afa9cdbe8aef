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

__device__ __constant__ uint8_t c_border_l[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// reference acceptance of one code-length set given its per-length counts (src/inftree.ts:131-178,:298):
// 0 ok, 1 oversubscribed, 2 incomplete, 3 empty
__device__ __forceinline__ int classify_counts(const uint16_t* cnt, int max_len)
{
    int j = 1;
    while (j <= max_len && cnt[j] == 0) j++;
    if (j > max_len) return 3;
    int g = max_len;
    while (cnt[g] == 0) g--;
    int y = 1 << j;
    for (; j < g; j++, y <<= 1) { y -= cnt[j]; if (y < 0) return 1; }
    y -= cnt[g];
    if (y < 0) return 1;
    return (y != 0 && g != 1) ? 2 : 0;
}

// Does the reference accept a dynamic block header at bit p?  (src/infblocks.ts:334-523 with the tree rules of
// src/inftree.ts:131-178,:298-379.)  lut: 128 bytes of scratch private to the thread.
__device__ bool is_dynamic_header(const uint32_t* src32, uint64_t p, uint64_t total_bits, uint8_t* lut)
{
    if (p + 17 + 12 > total_bits) return false;
    const uint32_t h = bits32_at(src32, p);
    if (((h >> 1) & 3u) != 2u) return false;                              // BTYPE = dynamic
    const uint32_t hlit = (h >> 3) & 31u, hdist = (h >> 8) & 31u, hclen = (h >> 13) & 15u;
    if (hlit > 29u || hdist > 29u) return false;                          // src/infblocks.ts:355
    const int ncl = 4 + (int)hclen, nl = 257 + (int)hlit, nd = 1 + (int)hdist;
    uint64_t q = p + 17;
    if (q + 3ull * ncl > total_bits) return false;
    // code-length-code lengths and their tree (inflate_trees_bits)
    uint8_t cl[19];
    #pragma unroll
    for (int i = 0; i < 19; i++) cl[i] = 0;
    uint16_t cnt[16];
    #pragma unroll
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    {
        uint32_t w = bits32_at(src32, q);
        int have = 32;
        for (int i = 0; i < ncl; i++) {
            if (have < 3) { w = bits32_at(src32, q); have = 32; }
            const uint32_t v = w & 7u;
            w >>= 3; have -= 3; q += 3;
            cl[c_border_l[i]] = (uint8_t)v;
            cnt[v]++;
        }
    }
    {
        const int st = classify_counts(cnt, 7);
        if (st != 0) return false;
    }
    int g = 7;
    while (cnt[g] == 0) g--;
    const int l = g;                                                  // root width = min(7, g)
    // 2^l-entry LUT: symbol | length << 5
    {
        uint32_t code = 0;
        for (int k = 1; k <= g; k++) {
            for (int s = 0; s < 19; s++) {
                if (cl[s] != k) continue;
                const uint32_t rev = __brev(code) >> (32 - k);
                for (uint32_t t = rev; t < (1u << l); t += (1u << k)) lut[t] = (uint8_t)(s | (k << 5));
                code++;
            }
            code <<= 1;
        }
        if (g == 1 && cnt[1] == 1) lut[1] = lut[0];
    }
    // the nl + nd code lengths (run-length coded); only their per-length counts are needed
    uint16_t cL[16], cD[16];
    #pragma unroll
    for (int i = 0; i < 16; i++) { cL[i] = 0; cD[i] = 0; }
    const int total = nl + nd;
    int index = 0;
    uint32_t prev = 0;
    bool ok = true;
    while (index < total) {
        if (q + 14 > total_bits) { ok = false; break; }
        const uint32_t w = bits32_at(src32, q);
        const uint32_t e = lut[w & ((1u << l) - 1u)];
        const int tb = (int)(e >> 5), c = (int)(e & 31u);
        if (c < 16) {
            q += tb;
            if (index < nl) cL[c]++; else cD[c]++;
            prev = (uint32_t)c;
            index++;
        } else {
            const int xb = c == 18 ? 7 : c - 14;
            int rep = (c == 18 ? 11 : 3) + (int)((w >> tb) & ((1u << xb) - 1u));
            q += tb + xb;
            if (index + rep > total || (c == 16 && index < 1)) { ok = false; break; }
            const uint32_t v = c == 16 ? prev : 0u;
            prev = v;
            while (rep--) { if (index < nl) cL[v]++; else cD[v]++; index++; }
        }
    }
    if (!ok) return false;
    // literal/length and distance trees (inflate_trees_dynamic); the MANY arena limit is ignored here:
    // a block rejected only by that rule is still a block boundary for the index
    {
        const int sl = classify_counts(cL, 15);
        if (sl != 0) return false;
        const int sd = classify_counts(cD, 15);
        if (sd == 1 || sd == 2) return false;
        if (sd == 3 && nl > 257) return false;
    }
    return true;
}

__device__ __forceinline__ uint64_t bits64_at(const uint32_t* src32, uint64_t p)
{
    const uint64_t w = p >> 5;
    const uint32_t a = __ldg(src32 + w), b = __ldg(src32 + w + 1), c = __ldg(src32 + w + 2);
    const uint32_t sh = (uint32_t)(p & 31);
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}

// Necessary condition, in registers only: the 4 + HCLEN code-length-code lengths form a complete code
// (Kraft sum exactly 1) or the one-code special case.
__device__ __forceinline__ bool precode_plausible(const uint32_t* src32, uint64_t p, uint32_t h)
{
    const uint32_t ncl = 4u + ((h >> 13) & 15u);
    const uint64_t v = bits64_at(src32, p + 17) & ((1ull << (3u * ncl)) - 1ull);
    uint32_t x0 = (uint32_t)v & 0x3fffffffu, x1 = (uint32_t)(v >> 30);
    uint32_t k = 0;
    #pragma unroll
    for (int i = 0; i < 10; i++) { k += (128u >> (x0 & 7u)) & 0x7fu; x0 >>= 3; }
    #pragma unroll
    for (int i = 0; i < 9; i++) { k += (128u >> (x1 & 7u)) & 0x7fu; x1 >>= 3; }
    return k == 128u || k == 64u;
}

// pass 1a, stage 1: every bit position is tested for BTYPE / HLIT / HDIST (22 % pass); the positions that
// pass are queued per warp so that the Kraft test runs on full warps; its survivors (a fraction of a
// per cent) go to surv[] for stage 2.
constexpr int PF_TILE = 4096;
__global__ void __launch_bounds__(256) prefilter_headers(const uint8_t* src, uint64_t first_bit, uint64_t end_bit, uint64_t total_bits,
                                                         uint64_t* surv, unsigned long long* n_surv, unsigned long long cap)
{
    __shared__ uint32_t queue_all[8][64];
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* q = queue_all[warp];
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
        for (uint32_t r = 0; r < PF_TILE / 32; r++) {
            const uint64_t p = base + r * 32u + lane;
            bool ok = p < end_bit && p + 17 + 12 <= total_bits;
            if (ok) {
                const uint32_t h = bits32_at(src32, p);
                ok = ((h >> 1) & 3u) == 2u && ((h >> 3) & 31u) <= 29u && ((h >> 8) & 31u) <= 29u;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (ok) q[qn + __popc(m & ((1u << lane) - 1u))] = r * 32u + lane;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32u) {
                qn -= 32u;
                kraft(base + q[qn + lane]);
                __syncwarp();
            }
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
    __shared__ uint8_t lut_all[128 * 128];
    uint8_t* lut = lut_all + threadIdx.x * 128;
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    const unsigned long long n = *n_surv < cap ? *n_surv : cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t p = surv[i];
        if (is_dynamic_header(src32, p, total_bits, lut)) {
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
