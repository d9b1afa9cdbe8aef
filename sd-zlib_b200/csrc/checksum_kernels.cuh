// checksum_kernels.cuh - kernels built on checksum_device.cuh
//   finalize_streams_kernel : per-stream running checksum of the inflated bytes and the
//                             finish() record (src/sd-inflate.ts:133-149, :159-179, :214-225)
//   adler_units_kernel / adler_fold_kernel : large-buffer adler32 with seed chaining
//   crc_tasks_kernel / crc_fold_kernel     : large-buffer crc32 with seed chaining
#pragma once
#include "checksum_device.cuh"
#include "../../include/sdz_codes.h"

namespace sdz {

// ------------------------------------------------------------------ per-stream records
// One warp per stream (dynamic scheduling).  The reference checksums every <= 16 KiB chunk
// append() emits, chaining the seed (src/sd-inflate.ts:137-146); with single-buffer input
// all chunks but the last are exactly 16384 bytes, so chunk k = bytes [16384 k, 16384 (k+1)).
// SDZ_PARITY_SPEC: the record as zlib 1.3 would settle it (RFC 1950 / 1952): standard Adler-32 over the whole output (no Q1,
// so the 16 KiB chunking is invisible), a stored checksum / ISIZE of zero is still checked, both compared as unsigned
// 32-bit values (no Q13), and an empty stream has the checksum of nothing (1 / 0) rather than none.
__device__ __forceinline__ void finalize_spec(const uint32_t* tab, const uint8_t* out, sdz_result* R, uint32_t lane)
{
    if (R->thrown_inflate != 0) return;
    const uint32_t thrown = R->thrown_append;
    const uint64_t len = R->out_len;
    const uint32_t stored = (uint32_t)R->stored_checksum, isize = (uint32_t)R->stored_isize;
    const bool complete = R->complete != 0;
    const int container = R->container;
    uint32_t running = container == SDZ_GZIP ? 0u : 1u;
    if (!thrown && len > 0) {
        const uint8_t* p = out + R->out_off;
        if (container == SDZ_GZIP) running = crc_call_warp(tab, p, len, 0u, lane);
        else running = adler_call_warp(p, len, 1u, lane, true);
    }
    if (lane == 0) {
        const bool have = !thrown;
        int cks = SDZ_UNCHECKED, fsz = SDZ_UNCHECKED;
        if (complete && container != SDZ_RAW) cks = stored == running ? SDZ_MATCH : SDZ_MISMATCH;
        if (complete && container == SDZ_GZIP) fsz = isize == (uint32_t)len ? SDZ_MATCH : SDZ_MISMATCH;
        const int success = complete && !thrown && cks != SDZ_MISMATCH && fsz != SDZ_MISMATCH;
        R->running_checksum = have ? (int32_t)running : 0;
        R->have_running = have ? 1 : 0;
        R->checksum_state = (uint8_t)cks;
        R->size_state = (uint8_t)fsz;
        R->success = (uint8_t)success;
        int ti = SDZ_THROW_NONE;
        if (thrown) ti = (int)thrown;
        else if (!complete) ti = SDZ_THROW_UNEXPECTED_EOF;
        else if (cks == SDZ_MISMATCH) ti = SDZ_THROW_INTEGRITY;
        else if (fsz == SDZ_MISMATCH) ti = SDZ_THROW_SIZE_CHECK;
        R->thrown_inflate = (uint8_t)ti;
    }
}

__global__ void __launch_bounds__(256) finalize_streams_kernel(const uint8_t* out, sdz_result* res,
                                                               unsigned long long n, unsigned long long* counter, uint32_t spec)
{
    __shared__ uint32_t tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (&g_crc_tab[0][0])[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n) break;
        sdz_result* R = res + idx;
        if (spec) { finalize_spec(tab, out, R, lane); continue; }
        if (R->thrown_inflate != 0) continue;            // inflate() rejected the call before decoding
        const uint32_t thrown = R->thrown_append;
        const uint64_t len = R->out_len;
        const int32_t stored = R->stored_checksum, isize = R->stored_isize;
        const bool complete = R->complete != 0;
        uint32_t running = 0;
        bool have = false;
        if (!thrown && len > 0) {
            const uint8_t* p = out + R->out_off;
            have = true;
            if (R->container == SDZ_GZIP) {
                running = crc_call_warp(tab, p, len, 0u, lane);
            } else {
                running = 1u;
                for (uint64_t off = 0; off < len; off += 16384) {
                    uint64_t cl = min((uint64_t)16384, len - off);
                    running = adler_call_warp(p + off, cl, running, lane);
                }
            }
        }
        if (lane == 0) {
            int cks = stored == 0 ? SDZ_UNCHECKED : ((have && stored == (int32_t)running) ? SDZ_MATCH : SDZ_MISMATCH);
            int fsz = isize == 0 ? SDZ_UNCHECKED : (((int64_t)isize == (int64_t)len) ? SDZ_MATCH : SDZ_MISMATCH);
            int success = complete && cks != SDZ_MISMATCH && fsz != SDZ_MISMATCH;
            R->running_checksum = have ? (int32_t)running : 0;
            R->have_running = have ? 1 : 0;
            R->checksum_state = (uint8_t)cks;
            R->size_state = (uint8_t)fsz;
            R->success = (uint8_t)success;
            int ti = SDZ_THROW_NONE;
            if (thrown) ti = (int)thrown;
            else if (!success) {
                if (!complete) ti = SDZ_THROW_UNEXPECTED_EOF;
                else if (cks == SDZ_MISMATCH) ti = SDZ_THROW_INTEGRITY;
                else if (fsz == SDZ_MISMATCH) ti = SDZ_THROW_SIZE_CHECK;
                else ti = SDZ_THROW_DECOMPRESSION;
            }
            R->thrown_inflate = (uint8_t)ti;
        }
    }
}

// ------------------------------------------------------------------ batched per-buffer checksums
// one warp per buffer (dynamic scheduling): adler32(buf, seed) or crc32(buf, seed), one reference call each
__global__ void __launch_bounds__(256) checksum_batch_kernel(const uint8_t* base, const uint64_t* off, const uint64_t* len,
                                                             const uint8_t* kind, const int32_t* seeds, unsigned long long n,
                                                             int32_t* out, unsigned long long* counter)
{
    __shared__ uint32_t tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (&g_crc_tab[0][0])[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n) break;
        const uint8_t* p = base + off[idx];
        const bool crc = kind[idx] != 0;
        const uint32_t seed = seeds ? (uint32_t)seeds[idx] : (crc ? 0u : 1u);
        const uint32_t v = crc ? crc_call_warp(tab, p, len[idx], seed, lane) : adler_call_warp(p, len[idx], seed, lane);
        if (lane == 0) out[idx] = (int32_t)v;
    }
}

// ------------------------------------------------------------------ large-buffer adler32
// Segment s = bytes [seg_off[s], seg_off[s] + seg_len[s]) of the buffer; it is cut into
// units of 5552 bytes (the reference's NMAX blocks); unit_base[s] = first unit of segment s.
struct SegTable {
    const uint64_t* seg_off;
    const uint64_t* seg_len;
    const uint64_t* unit_base;     // n_seg + 1 entries
    uint32_t n_seg;
};

__device__ __forceinline__ uint32_t find_segment(const uint64_t* base, uint32_t n_seg, uint64_t u)
{
    uint32_t lo = 0, hi = n_seg;                          // base[lo] <= u < base[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (base[mid] <= u) lo = mid; else hi = mid;
    }
    return lo;
}

// one warp per unit: partial[u] = (S, W)
__global__ void __launch_bounds__(256) adler_units_kernel(const uint8_t* p, SegTable T, uint64_t n_units, uint2* partial)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t u = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_units; u += warps) {
        uint32_t s = find_segment(T.unit_base, T.n_seg, u);
        uint64_t k = u - T.unit_base[s];
        uint64_t start = k * ADLER_NMAX;
        uint32_t L = (uint32_t)min((uint64_t)ADLER_NMAX, T.seg_len[s] - start);
        uint32_t S, Tt;
        adler_unit_sums(p + T.seg_off[s] + start, L, lane, S, Tt);
        if (lane == 0) partial[u] = make_uint2(S, (uint32_t)((uint64_t)L * S - Tt));
    }
}

// Per-segment, seed-independent part (one CTA per segment):
//   pref[u]   reduced exclusive prefix of the unit byte sums inside the segment
//   seg[s]    { S = sum of bytes mod BASE, C = sum_k (L_k * pref_k + W_k) mod BASE,
//               Wq = sum_k W_k mod 2^32, Pq = sum_k pref_k mod 2^32 }
// so that adler32(seg, seed) = ((a0 + S) % BASE) | (((b0 + a0 * len + C) % BASE) << 16)
// (the usual adler32_combine), and the Q1 value can be rebuilt from Wq, Pq and pref[].
__global__ void __launch_bounds__(1024) adler_seg_kernel(SegTable T, const uint2* partial, uint32_t* pref, uint4* seg)
{
    __shared__ uint64_t sh[1024];
    const uint32_t t = threadIdx.x, NT = blockDim.x, s = blockIdx.x;
    const uint64_t u0 = T.unit_base[s], K = T.unit_base[s + 1] - u0;
    const uint64_t slen = T.seg_len[s];
    const uint64_t per = (K + NT - 1) / NT;
    const uint64_t lo = min(K, per * t), hi = min(K, lo + per);
    uint64_t mysum = 0;
    for (uint64_t k = lo; k < hi; k++) mysum += partial[u0 + k].x;
    sh[t] = mysum % ADLER_BASE;
    __syncthreads();
    for (uint32_t o = 1; o < NT; o <<= 1) {
        uint64_t v = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] = (sh[t] + v) % ADLER_BASE;
        __syncthreads();
    }
    uint32_t P = t ? (uint32_t)sh[t - 1] : 0u;
    const uint32_t total = (uint32_t)sh[NT - 1];
    __syncthreads();
    uint64_t c = 0;
    uint32_t wq = 0, pq = 0;
    for (uint64_t k = lo; k < hi; k++) {
        uint2 pw = partial[u0 + k];
        uint32_t L = (uint32_t)min((uint64_t)ADLER_NMAX, slen - k * ADLER_NMAX);
        pref[u0 + k] = P;
        c = (c + (uint64_t)L * P + pw.y) % ADLER_BASE;
        wq += pw.y; pq += P;
        P = (uint32_t)(((uint64_t)P + pw.x) % ADLER_BASE);
    }
    // block reduce (c mod BASE, wq and pq wrapping)
    sh[t] = c; __syncthreads();
    for (uint32_t o = NT >> 1; o > 0; o >>= 1) { if (t < o) sh[t] = (sh[t] + sh[t + o]) % ADLER_BASE; __syncthreads(); }
    const uint32_t C = (uint32_t)sh[0];
    __syncthreads();
    sh[t] = (uint64_t)wq | ((uint64_t)pq << 32); __syncthreads();
    for (uint32_t o = NT >> 1; o > 0; o >>= 1) {
        if (t < o) {
            uint64_t x = sh[t], y = sh[t + o];
            sh[t] = (uint64_t)((uint32_t)x + (uint32_t)y) | ((uint64_t)((uint32_t)(x >> 32) + (uint32_t)(y >> 32)) << 32);
        }
        __syncthreads();
    }
    if (t == 0) seg[s] = make_uint4(total, C, (uint32_t)sh[0], (uint32_t)(sh[0] >> 32));
}

// single CTA: the seed chain over segments.  Ordinary segments are one combine; a segment
// whose length is a non-zero multiple of 5552 needs sum_k floor((a0 + pref_k) / BASE), which
// depends on the incoming seed and is counted cooperatively.
__global__ void __launch_bounds__(1024) adler_chain_kernel(SegTable T, const uint32_t* pref, const uint4* seg,
                                                           uint32_t seed, int32_t* values)
{
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t sh_seed;
    const uint32_t t = threadIdx.x, NT = blockDim.x;
    if (t == 0) sh_seed = seed;
    __syncthreads();
    for (uint32_t s = 0; s < T.n_seg; s++) {
        const uint64_t u0 = T.unit_base[s], K = T.unit_base[s + 1] - u0;
        const uint64_t slen = T.seg_len[s];
        const uint32_t sd = sh_seed;
        const uint32_t a0 = sd & 0xffffu, b0 = (sd >> 16) & 0xffffu;
        const bool quirk = K > 0 && (slen % ADLER_NMAX) == 0;
        uint32_t wraps = 0;
        if (quirk) {
            uint32_t cnt = 0;
            for (uint64_t k = 1 + t; k < K; k += NT) cnt += (a0 + pref[u0 + k]) / ADLER_BASE;   // unit 0 sees a0 unreduced
            sh[t] = cnt;
            __syncthreads();
            for (uint32_t o = NT >> 1; o > 0; o >>= 1) { if (t < o) sh[t] += sh[t + o]; __syncthreads(); }
            wraps = sh[0];
        }
        if (t == 0) {
            uint32_t v;
            if (K == 0) v = a0 | (b0 << 16);
            else {
                uint4 g = seg[s];
                uint32_t lo16 = (uint32_t)(((uint64_t)a0 + g.x) % ADLER_BASE);
                uint32_t hi16;
                if (quirk) {
                    // sum2 = b0 + sum_k (5552 * A_k + W_k + BASE), A_k = a0 + pref_k - BASE * floor(.)  (mod 2^32)
                    uint32_t Kq = (uint32_t)K;
                    uint32_t sum2 = b0 + ADLER_NMAX * (Kq * a0 + g.w - ADLER_BASE * wraps) + g.z + ADLER_BASE * Kq;
                    hi16 = sum2 & 0xffffu;
                } else {
                    hi16 = (uint32_t)(((uint64_t)b0 + (uint64_t)a0 * (slen % ADLER_BASE) + g.y) % ADLER_BASE);
                }
                v = lo16 | (hi16 << 16);
            }
            values[s] = (int32_t)v;
            sh_seed = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ large-buffer crc32
// Strided Horner: a warp walks its task in rows of 512 bytes; lane l owns bytes
// [16 l, 16 l + 16) of every row as four independent 32-bit accumulators, each advanced by
// acc = (acc * x^4096 mod P) ^ word.  The multiply is four byte-indexed table lookups; the
// tables are replicated per lane in shared memory (address = byte * 128 + lane * 4) so every
// lookup is bank-conflict free.  Loads are fully coalesced 16-byte vectors.
constexpr uint32_t CRC_ROW = 512;
constexpr uint32_t CRC_TASK = 256 * 1024;

struct CrcTaskTable {
    const uint64_t* seg_off;
    const uint64_t* seg_len;
    const uint64_t* task_base;     // n_seg + 1 entries; tasks of CRC_TASK bytes per segment
    uint32_t n_seg;
};

// tables for multiply-by-x^4096, global copy: [4][256]
__device__ uint32_t g_crc_stride_tab[4][256];

__device__ __forceinline__ uint32_t crc_stride_step(const uint8_t* tl, uint32_t acc, uint32_t w)
{
    // tl points at this lane's column of table 0; table j is 32 KiB further
    uint32_t i0 = (acc & 0xffu) << 7, i1 = ((acc >> 8) & 0xffu) << 7, i2 = ((acc >> 16) & 0xffu) << 7, i3 = (acc >> 24) << 7;
    uint32_t t0 = *reinterpret_cast<const uint32_t*>(tl + i0);
    uint32_t t1 = *reinterpret_cast<const uint32_t*>(tl + 32768 + i1);
    uint32_t t2 = *reinterpret_cast<const uint32_t*>(tl + 65536 + i2);
    uint32_t t3 = *reinterpret_cast<const uint32_t*>(tl + 98304 + i3);
    return (t0 ^ t1 ^ t2) ^ (t3 ^ w);
}

// dynamic shared memory: 4 x 32 KiB replicated stride tables + 4 KiB standard tables
__global__ void __launch_bounds__(1024, 1) crc_tasks_kernel(const uint8_t* p, CrcTaskTable T, uint64_t n_tasks,
                                                            uint32_t* partial, unsigned long long* counter)
{
    extern __shared__ __align__(16) unsigned char crc_smem[];
    uint32_t* rep = reinterpret_cast<uint32_t*>(crc_smem);              // [4][256][32]
    uint32_t* std_tab = rep + 4 * 256 * 32;                             // [4][256]
    for (uint32_t i = threadIdx.x; i < 4 * 256 * 32; i += blockDim.x) rep[i] = (&g_crc_stride_tab[0][0])[i >> 5];
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) std_tab[i] = (&g_crc_tab[0][0])[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint8_t* tl = crc_smem + lane * 4;

    for (;;) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(counter, 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        uint32_t s = find_segment(T.task_base, T.n_seg, task);
        uint64_t k = task - T.task_base[s];
        uint64_t t_lo = k * CRC_TASK;
        uint64_t t_len = min((uint64_t)CRC_TASK, T.seg_len[s] - t_lo);
        const uint8_t* q = p + T.seg_off[s] + t_lo;

        // head up to 16-byte alignment, whole rows, tail
        uint64_t head = min(t_len, (uint64_t)((16 - (reinterpret_cast<uintptr_t>(q) & 15)) & 15));
        uint64_t rows = (t_len - head) / CRC_ROW;
        uint64_t tail = t_len - head - rows * CRC_ROW;

        uint32_t x = 0;
        if (lane == 0 && head) x = crc_shift(crc_bytes(std_tab, 0u, q, head), t_len - head);

        const uint4* q4 = reinterpret_cast<const uint4*>(q + head) + lane;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        uint64_t r = 0;
        for (; r + 2 <= rows; r += 2) {
            uint4 w0 = __ldg(q4 + r * 32);
            uint4 w1 = __ldg(q4 + (r + 1) * 32);
            a0 = crc_stride_step(tl, a0, w0.x); a1 = crc_stride_step(tl, a1, w0.y);
            a2 = crc_stride_step(tl, a2, w0.z); a3 = crc_stride_step(tl, a3, w0.w);
            a0 = crc_stride_step(tl, a0, w1.x); a1 = crc_stride_step(tl, a1, w1.y);
            a2 = crc_stride_step(tl, a2, w1.z); a3 = crc_stride_step(tl, a3, w1.w);
        }
        for (; r < rows; r++) {
            uint4 w0 = __ldg(q4 + r * 32);
            a0 = crc_stride_step(tl, a0, w0.x); a1 = crc_stride_step(tl, a1, w0.y);
            a2 = crc_stride_step(tl, a2, w0.z); a3 = crc_stride_step(tl, a3, w0.w);
        }
        if (rows) {
            // the four accumulators are four consecutive words: fold with the standard tables
            uint32_t c = 0, accs[4] = { a0, a1, a2, a3 };
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                c ^= accs[j];
                c = std_tab[768 + (c & 0xffu)] ^ std_tab[512 + ((c >> 8) & 0xffu)] ^ std_tab[256 + ((c >> 16) & 0xffu)] ^ std_tab[c >> 24];
            }
            x ^= crc_shift(c, 16ull * (31 - lane) + tail);
        }
        if (tail) {
            const uint8_t* tq = q + head + rows * CRC_ROW;
            uint64_t lo = min(tail, (uint64_t)(16u * lane)), hi = min(tail, lo + (uint64_t)16);
            if (hi > lo) x ^= crc_shift(crc_bytes(std_tab, 0u, tq + lo, hi - lo), tail - hi);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) partial[task] = x;                 // G(task): zero-init remainder
    }
}

// one CTA per segment, seed-independent: gseg[s] = { XOR_t shift(G_t, bytes after task t), x^(8 len) mod P }
__global__ void __launch_bounds__(256) crc_seg_kernel(CrcTaskTable T, const uint32_t* partial, uint2* gseg)
{
    __shared__ uint32_t sh[256];
    const uint32_t t = threadIdx.x, NT = blockDim.x, s = blockIdx.x;
    const uint64_t t0 = T.task_base[s], K = T.task_base[s + 1] - t0;
    const uint64_t slen = T.seg_len[s];
    uint32_t x = 0;
    for (uint64_t k = t; k < K; k += NT) {
        uint64_t end = min(slen, (k + 1) * CRC_TASK);
        x ^= crc_shift(partial[t0 + k], slen - end);
    }
    sh[t] = x;
    __syncthreads();
    for (uint32_t o = NT >> 1; o > 0; o >>= 1) { if (t < o) sh[t] ^= sh[t + o]; __syncthreads(); }
    if (t == 0) gseg[s] = make_uint2(sh[0], gf2_x2n(slen, 3));
}

// the seed chain: reg = reg * x^(8 len) ^ G(seg) on the raw register (value = ~reg)
__global__ void crc_chain_kernel(const uint2* gseg, uint32_t n_seg, uint32_t seed, int32_t* values)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t reg = ~seed;
    for (uint32_t s = 0; s < n_seg; s++) {
        uint2 g = gseg[s];
        reg = gf2_mulmod(g.y, reg) ^ g.x;
        values[s] = (int32_t)~reg;
    }
}

}  // namespace sdz
