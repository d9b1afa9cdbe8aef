// fast_kernels.cuh - two-phase batched inflate for sm_100a (the common case of inflate_kernel.cuh).
//
// inflate_kernel.cuh decodes a stream with a sub-warp group whose lanes all replicate the Huffman decoder
// state and whose copies are done a few bytes per lane: ~25 warp instructions per symbol, issue-bound
// (profiles/r01e_full_summary.md).  Here the two halves of the reference's hot loop
// (inflate_fast, src/infcodes.ts:62-301) are split:
//
//   phase A  huff_tokens_kernel   ONE LANE per stream: 32 streams per warp run the symbol loop in lockstep,
//            straight-line and predicated, and do nothing but Huffman decoding (src/infcodes.ts:96-158,
//            tables of src/inftree.ts:95-299 as shared-memory LUTs per lane).  Every symbol becomes one
//            32-bit token (literal byte, or length + distance) in HBM; four tokens leave a lane as one
//            16-byte store.  No window, no copies.
//   phase B  lz_resolve2_kernel   ONE WARP per stream, lane = TOKEN (src/infcodes.ts:160-207 is what it replaces): 32
//            tokens per step, a warp scan of their lengths gives the output offsets; the bytes are assembled in a
//            1 KiB shared-memory ring per warp and leave it as coalesced 16-byte vector stores (see the comment at
//            the kernel).  lz_resolve_kernel is its first, byte-centric version (lane = output byte; SDZ_B2=0).
//
// Phase A decodes optimistically and validates at the end of the stream; it only finishes streams whose record it
// can write with certainty: container header without FEXTRA (a preset dictionary only when the FDICT header names the
// dictionary the caller supplied and it is shorter than the window: SURVEY Q14), fixed and dynamic blocks whose trees the
// reference accepts (complete codes, table arena within MANY), no distance reaching before the start of
// the window (Q6), the whole trailer present and nothing after it (Q4), and - raw streams - the reference's lookahead
// rule replayed on the final end-of-block lookup and, when that one stalls, on the last data symbol (Q15: the stream is
// then complete = false with all data delivered).
// EVERYTHING else (stored blocks with their Q2 rule, errors, truncation, missing / wrong / 32 KiB dictionaries, slots that
// are too small ...) is handed, whole stream, to the general decoder inflate_kernel<4> through a device-side list,
// so records stay bit-exact with the reference in every case.  There is no CPU path.
#pragma once
#include "inflate_kernel.cuh"

namespace sdz {

constexpr uint32_t TOK_LIT = 0x80000000u;      // literal: TOK_LIT | byte.  match: (dist - 1) << 9 | len.  0: no-op
constexpr uint32_t NTOK_HANDED_OVER = 0xffffffffu;
// LUT root widths (entry formats of make_lut()).  Phase A is latency-bound with ONE warp per scheduler when the
// literal/length root is 9 bits wide (1,484 bytes of tables per lane -> 4 warps per SM: profiles/r02a_summary.md, every
// issued instruction waits 3 more cycles).  An 8-bit root lets 6 warps share an SM, a 7-bit root 7 (8 with a 4-chunk
// input ring).  Codes longer than the root go through the per-group general decoder (canon_long), so a narrower root
// trades residency for more of those (synthetic text: 0.5 % of the symbols at 8 or 9 bits, 0.9 % at 7; English text
// 1.6 % at 9).  Measured on the 65,536-stream batch (gpurun_out/variants_r02b.log): phase A 17.6 ms at 9 bits,
// 14.9 ms at 8, 11.2 ms at 7.
#ifndef SDZ_FA_RL
#define SDZ_FA_RL 7
#endif
#ifndef SDZ_FA_RING
#define SDZ_FA_RING 8
#endif
#ifndef SDZ_CACHE_HINTS
#define SDZ_CACHE_HINTS 1              // L2 policies: compressed input and tokens evict_first (read / written once), output evict_last (the window)
#endif
// -DSDZ_CHECKED: every global / shared address the two fast-path kernels form is checked against the region it must stay in
// (trap on violation -> the call fails with a CUDA error).  compute-sanitizer is closed on the GPU pool
// (profiles/r02j_compute_sanitizer_closed.log); the GPU test suite is run once per change against this build instead.
#ifdef SDZ_CHECKED
#define SDZ_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define SDZ_CHECK(cond) do { } while (0)
#endif
#ifndef SDZ_FA_DEFER
#define SDZ_FA_DEFER 4u               // the general single-symbol decoder runs every n-th group of four symbols (power of two)
#endif
#if SDZ_FA_RL == 9
#define SDZ_FA_RL_MASK 511
#elif SDZ_FA_RL == 8
#define SDZ_FA_RL_MASK 255
#else
#define SDZ_FA_RL_MASK 127
#endif
// commit groups that may still be in flight when a group of symbols ends: a chunk is requested when its slot is free,
// i.e. RING - 1 chunks (4 RING - 3 words) ahead of the reader, and a group of four symbols reads at most six words
#if SDZ_FA_RING == 8
#define SDZ_FA_RING_WAIT 3
#define SDZ_FA_RING_WMASK 31
#elif SDZ_FA_RING == 4
#define SDZ_FA_RING_WAIT 1
#define SDZ_FA_RING_WMASK 15
#else
#error "SDZ_FA_RING: 4 or 8 chunks of 16 bytes"
#endif
#define SDZ_STR2(x) #x
#define SDZ_STR(x) SDZ_STR2(x)
constexpr int FA_RL = SDZ_FA_RL, FA_RD = 7;
static_assert(FA_RL >= 7 && FA_RL <= 9, "literal/length root width");
// while a dynamic header is parsed the lane's table area holds the block's code lengths (320 bytes from the start of
// lut_l, running into lut_d when the root is 7 bits wide), the code-length-code lengths and their 7-bit LUT (tail of lut_d)
constexpr int FA_HDR_CL = (1 << FA_RL) * 2 + 96, FA_HDR_BLUT = (1 << FA_RL) * 2 + 128;

// per-lane (= per resident stream) decode tables
struct alignas(4) LaneSmem {
    uint16_t lut_l[1 << FA_RL];
    uint16_t lut_d[1 << FA_RD];
    uint16_t cnt_l[16];                // [1..15] codes per length; [0] = longest code after the build
    uint16_t cnt_d[16];
    uint16_t start[4];                 // canonical-walk state after the root bits (canon_long)
    uint16_t long_l[32];               // first literal/length symbols whose code is longer than the root
    uint16_t sorted_d[32];             // distance symbols in canonical order
    uint16_t pad_[2];                  // an odd number of words: the stride spreads the 32 lanes' tables over the banks
};
static_assert(sizeof(LaneSmem) == 460 + (2 << FA_RL) && (sizeof(LaneSmem) / 4) % 2 == 1, "LaneSmem layout");
// compressed input of the symbol loop: per lane a ring of FA_RING_CHUNKS x 16 bytes in shared memory, filled by 16-byte
// cp.async (LDGSTS) several groups of symbols ahead of the reader.  A plain `nw = input[wp]` costs the whole warp one
// global-memory round trip per ITERATION: the lanes take their next word in different iterations, but the register
// scoreboard is per warp, so the consumer of iteration i waits for the load some other lane issued in iteration i - 1
// (first version: 26 ms for the 65,536-stream batch, 12 cycles per issued instruction).  LDGSTS is tracked by
// commit/wait groups instead, so the copies stay in flight across iterations.
// Layout: lane l owns FA_RING_CHUNKS * 16 contiguous bytes at l * FA_RING_STRIDE (word w of the stream sits at
// (w mod 32) * 4: one AND + one multiply-add per read); the 16 bytes of padding per lane keep the stride a multiple
// of 16 (LDGSTS.128 alignment) and spread lanes over the banks (4-way conflict when all lanes read the same slot).
constexpr int FA_RING_CHUNKS = SDZ_FA_RING;
constexpr int FA_RING_STRIDE = FA_RING_CHUNKS * 16 + 16;
constexpr int FA_RING_BYTES = 32 * FA_RING_STRIDE;
constexpr int FA_WARP_SMEM = 32 * (int)sizeof(LaneSmem) + SCRATCH_U16 * 2 + FA_RING_BYTES;      // tables | build scratch | input ring

struct FastParams {
    InflateParams I;                   // the batch
    uint32_t* tokens;
    const uint64_t* tok_off;           // first token of every stream (multiple of 4); tok_off[n] = arena size
    uint32_t* ntok;                    // tokens written per stream (multiple of 4), or NTOK_HANDED_OVER
    uint32_t* fb_list;                 // streams handed to the general decoder
    unsigned long long* fb_count;
    unsigned long long* counter_a;     // work counters of THIS launch: the launch covers streams [first, first + count)
    unsigned long long* counter_b;
    uint32_t first, count;
    uint16_t* sorted_l;                // 288 u16 per lane of the phase-A grid
    const uint32_t* order;             // work item k of the batch is stream order[k]: longest compressed streams first
    uint32_t dict_ok;                  // phase B resolves sources inside a preset dictionary (lz_resolve2_kernel only)
};

// token capacity of a stream: a token needs at least one output byte, and streams with fewer than four
// input bits per symbol are left to the general decoder (extremely repetitive data)
__device__ __host__ __forceinline__ uint64_t token_cap(uint32_t in_len, uint32_t out_cap)
{
    // Slots, not symbols: a lane that meets a rare symbol leaves the rest of its group of four empty (and, while it waits
    // for the general decoder, whole groups).  Fixed-Huffman blocks (every literal is longer than the 7-bit root) and
    // run-heavy data need up to four slots per symbol; they are short streams, so the allowance is additive.
    uint64_t c = 2ull * in_len + 4096;
    if (c > 8ull * in_len) c = 8ull * in_len;
    if (c > 4ull * out_cap) c = 4ull * out_cap;
    return (c + 32 + 3) & ~3ull;
}

// exclusive scan of the token capacities (one block; n is a few 10^4)
// The kernel also orders the streams for phase A: a lane decodes its stream alone, so a launch lasts as long as its
// slowest lane, and lanes that finish early take the next stream of the launch's range.  Work items are therefore handed
// out longest compressed stream first (the decode time of a Huffman-coded stream follows its compressed size; 64 classes
// of 1 KiB, a counting sort): expensive streams start together at the beginning of a launch and cheap ones fill the gaps
// behind them.  Streams that did not shrink (compressed >= decompressed: stored blocks, which phase A hands over at
// their first block header) go last.
__device__ __forceinline__ uint32_t cost_class(uint32_t in_len, uint32_t out_cap)
{
    if (in_len >= out_cap) return 63u;
    const uint32_t c = 62u - min(in_len >> 10, 62u);
    return c;                                            // class 0 = the longest streams
}

__global__ void __launch_bounds__(1024) token_offsets_kernel(const uint32_t* in_len, const uint32_t* out_cap, unsigned long long n,
                                                             uint64_t* tok_off, uint32_t* order)
{
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry_s;
    __shared__ uint32_t hist[64], cursor[64];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned long long i = threadIdx.x; i < n; i += 1024) atomicAdd(&hist[cost_class(in_len[i], out_cap ? out_cap[i] : 0xffffffffu)], 1u);
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t acc = 0; for (int k = 0; k < 64; k++) { cursor[k] = acc; acc += hist[k]; } }
    __syncthreads();
    for (unsigned long long i = threadIdx.x; i < n; i += 1024)
        order[atomicAdd(&cursor[cost_class(in_len[i], out_cap ? out_cap[i] : 0xffffffffu)], 1u)] = (uint32_t)i;
    __syncthreads();
    for (unsigned long long base = 0; base < n; base += 1024) {
        const unsigned long long i = base + threadIdx.x;
        const uint64_t v = i < n ? token_cap(in_len[i], out_cap ? out_cap[i] : 0xffffffffu) : 0;
        uint64_t incl = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint64_t s = wsum[lane];
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= (uint32_t)d) s += t; }
            wsum[lane] = s;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t before = carry + (w ? wsum[w - 1] : 0) + incl - v;
        if (i < n) tok_off[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) tok_off[n] = carry_s;
}

// ---------------------------------------------------------------------- phase A

__device__ __forceinline__ uint32_t dict_len_of(const InflateParams& I, unsigned long long i)
{
    return (I.mode[i] & 0x80) && I.dict_len ? I.dict_len[i] : 0u;
}
__device__ __forceinline__ int32_t dict_adler_of(const InflateParams& I, unsigned long long i)
{
    return (I.mode[i] & 0x80) && I.dict_adler ? I.dict_adler[i] : 0;
}

// container header of a stream the fast path may finish (src/inflate.ts:142-401).  false: anything else
// (errors, truncation, FDICT, FEXTRA ...) - the general decoder writes that record.
struct Container {
    uint32_t hp;                       // first byte of the deflate data
    int32_t mtime;
    uint32_t name_off, name_len;
    uint32_t dict_used;                // bytes of preset dictionary that sit before the output (FDICT accepted), else 0
    int method;
    bool raw, is_gzip;
};
// `dict_len` / `dict_adler`: the dictionary the caller supplied (mode_raw & 0x80), as inflateSetDictionary would see it
// (src/inflate.ts:465-491): a zlib stream with FDICT is finished here when the caller's dictionary is the one the stream
// names (DICTID == the reference's adler32 of it) and is shorter than the window - then the reference's window simply
// starts with those bytes.  32 KiB and more (the 32,767 rule), a missing or a wrong dictionary: general decoder.
constexpr uint32_t FA_MAX_DICT = 32767;
__device__ __forceinline__ bool parse_container_clean(const uint8_t* src, uint32_t in_len, uint8_t mode_raw, uint32_t dict_len, int32_t dict_adler,
                                                      bool dict_ok, Container& C)
{
    const int mode = mode_raw & 0x7f;
    const bool has_dict = (mode_raw & 0x80) != 0;
    C.hp = 0; C.mtime = 0; C.name_off = 0; C.name_len = 0; C.method = 0; C.is_gzip = false; C.dict_used = 0;
    C.raw = mode == SDZ_MODE_RAW;
    if (in_len == 0) return false;
    if (mode == SDZ_MODE_SNIFF) {                                       // inflate(): src/sd-inflate.ts:194-207
        if (in_len < 2) return false;
        const uint32_t b0 = src[0], b1 = src[1];
        const bool ident = (b0 == 0x78 && (((b0 << 8) + b1) % 31) == 0) || (b0 == 0x1f && b1 == 0x8b);
        C.raw = !ident;
    }
    if (C.raw) return !has_dict;                                        // RangeError otherwise (src/sd-inflate.ts:69-71)
    uint32_t hp = 0, b;
    if (src[hp] == 0x1f) {
        hp++;
        if (hp >= in_len) return false;
        if (src[hp++] != 0x8b) return false;
        C.is_gzip = true;
    }
    if (hp >= in_len) return false;
    C.method = src[hp++];
    if ((C.method & 0xf) != 8 || (C.method >> 4) + 8 > 15) return false;
    if (hp >= in_len) return false;
    b = src[hp++];
    if (C.is_gzip) {
        const uint32_t gflags = b;
        if (hp + 6 > in_len) return false;                              // MTIME, XFL, OS
        C.mtime = (int32_t)((uint32_t)src[hp] | ((uint32_t)src[hp + 1] << 8) | ((uint32_t)src[hp + 2] << 16) | ((uint32_t)src[hp + 3] << 24));
        hp += 6;
        if (gflags & 4) return false;                                   // FEXTRA
        if (gflags & 8) {
            C.name_off = hp;
            for (;;) { if (hp >= in_len) return false; b = src[hp++]; if (b == 0) break; C.name_len++; }
        }
        if (gflags & 16)
            for (;;) { if (hp >= in_len) return false; b = src[hp++]; if (b == 0) break; }
        if (gflags & 2) { if (hp + 2 > in_len) return false; hp += 2; }
    } else {
        if ((((uint32_t)C.method << 8) + b) % 31 != 0) return false;
        if (b & 0x20) {                                                 // FDICT: preset dictionary (src/inflate.ts:196-230)
            if (!has_dict || !dict_ok || dict_len > FA_MAX_DICT || hp + 4 > in_len) return false;
            const int32_t dictid = (int32_t)(((uint32_t)src[hp] << 24) | ((uint32_t)src[hp + 1] << 16) | ((uint32_t)src[hp + 2] << 8) | (uint32_t)src[hp + 3]);
            if (dictid != dict_adler) return false;
            hp += 4;
            C.dict_used = dict_len;
        }
    }
    C.hp = hp;
    return true;
}

// ---- table construction for phase A: the warp builds one lane's tables.  Same results as build_tables<32>() of
// inflate_kernel.cuh (same acceptance rules, same LUT / sorted-symbol / start formats); two things are done
// differently because 32 builds per warp run back to back whenever the lanes of a warp reach a block boundary
// together (20 % of phase A's time in the first profile):
//   * the counting sort of the symbols by code length is done 32 symbols at a time (match_any ranks the lanes with
//     equal lengths) instead of serially by lane 0;
//   * the reference's table-arena limit (MANY = 1400 entries) needs the exact sub-table walk
//     (ref_table_total; huft_build's multi-level splitting has no simple closed-form bound) only when a code is longer
//     than its root: 2^lbits + 2^dbits <= 576 entries otherwise.

template <int KIND, int R>
__device__ __forceinline__ void make_lut_warp(uint32_t* aux, const uint8_t* lens, int n, int nzero, const uint16_t* cnt, uint16_t* start,
                                              uint16_t* sorted, uint16_t* lut, uint32_t lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    if (lane == 0) {
        uint32_t off = 0, code = 0;
        for (int k = 1; k <= 15; k++) {
            aux[k] = off;
            aux[16 + k] = code;
            off += cnt[k];
            code = (code + cnt[k]) << 1;
        }
        // canonical-walk state after R bits (canon_long)
        uint32_t first = 0, index = 0;
        for (int k = 1; k <= R && k <= 15; k++) { index += cnt[k]; first = (first + cnt[k]) << 1; }
        start[0] = (uint16_t)first; start[1] = (uint16_t)index;
    }
    uint32_t* lut32 = reinterpret_cast<uint32_t*>(lut);
    for (uint32_t i = lane; i < (1u << R) / 2; i += 32) lut32[i] = E_INVALID | (E_INVALID << 16);
    __syncwarp();
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int base = 0; base < n; base += 32) {
        const int sidx = base + (int)lane;
        const uint32_t k = sidx < n ? lens[sidx] : 0u;
        const uint32_t same = __match_any_sync(FULL, k);
        if (k) sorted[aux[k] + __popc(same & lt_mask)] = (uint16_t)(sidx | (k << 12));
        __syncwarp();
        if (k && (same & lt_mask) == 0u) aux[k] += __popc(same);        // (after the loop: one past the last symbol of this length)
        __syncwarp();
    }
    const int ncodes = n - nzero;
    for (int k = (int)lane; k < ncodes; k += 32) {
        const uint32_t e = sorted[k];
        const uint32_t sym = e & 0xfff, len = e >> 12;
        const uint32_t idx = (uint32_t)k - (aux[len] - cnt[len]);
        const uint32_t code = aux[16 + len] + idx;
        const uint32_t rev = __brev(code) >> (32 - len);
        if (len > (uint32_t)R) { lut[rev & ((1u << R) - 1u)] = (uint16_t)E_LONG; continue; }
        uint32_t entry;
        if (KIND == 0) {
            if (sym < 256) entry = sym;
            else if (sym == 256) entry = 0x100;
            else {
                const uint32_t i = sym - 257;
                if (i > 28) continue;                                   // 286/287: invalid (fixed block only)
                const uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                const uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                entry = 0x800 | (xb << 8) | (base - 3);
            }
        } else {
            if (sym > 29) continue;                                     // 30/31: invalid (fixed block only)
            const uint32_t xb = sym < 4 ? 0 : (sym >> 1) - 1;
            const uint32_t m = sym < 4 ? sym : 2 + (sym & 1);
            entry = (xb << 8) | m;
        }
        entry |= len << 12;
        for (uint32_t j = rev; j < (1u << R); j += (1u << len)) lut[j] = (uint16_t)entry;
    }
    __syncwarp();
}

__device__ __noinline__ TreeInfo build_tables_warp(LaneSmem* S, uint16_t* wscr, int nl, int nd, bool fixed, uint32_t lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    TreeInfo T;
    T.msg = SDZ_MSG_NONE; T.lbits = T.dbits = T.g_l = T.g_d = 0;
    int pad_l = 0, pad_d = 0, nz_l = 0, nz_d = 0;
    const uint8_t* lens = reinterpret_cast<const uint8_t*>(wscr + SORTED_L + SORTED_D);
    uint32_t* aux = reinterpret_cast<uint32_t*>(wscr + SORTED_L + SORTED_D + 160);
    const int st_l = classify<32>(lens, nl, 9, S->cnt_l, aux, &T.lbits, &T.g_l, &pad_l, &nz_l, (int)lane, FULL);
    if (!fixed && st_l != 0) { T.msg = st_l == 1 ? SDZ_MSG_OVERSUB_LITLEN_TREE : SDZ_MSG_INCOMPLETE_LITLEN_TREE; return T; }
    const int st_d = classify<32>(lens + nl, nd, fixed ? 5 : 6, S->cnt_d, aux, &T.dbits, &T.g_d, &pad_d, &nz_d, (int)lane, FULL);
    if (!fixed) {
        if (st_d == 1) { T.msg = SDZ_MSG_OVERSUB_DIST_TREE; return T; }
        if (st_d == 2) { T.msg = SDZ_MSG_INCOMPLETE_DIST_TREE; return T; }
        if (st_d == 3 && nl > 257) { T.msg = SDZ_MSG_EMPTY_DIST_TREE; return T; }
        // MANY = 1400 entries for both tables
        if (T.g_l > T.lbits || T.g_d > T.dbits) {
            const int used = ref_table_total<32>(S->cnt_l, T.g_l, pad_l, T.lbits, (int)lane, FULL);
            if (used > 1400) { T.msg = SDZ_MSG_OVERSUB_LITLEN_TREE; return T; }
            if (st_d != 3 && used + ref_table_total<32>(S->cnt_d, T.g_d, pad_d, T.dbits, (int)lane, FULL) > 1400) {
                T.msg = SDZ_MSG_OVERSUB_DIST_TREE; return T;
            }
        }
    }
    make_lut_warp<0, FA_RL>(reinterpret_cast<uint32_t*>(S->lut_d), lens, nl, nz_l, S->cnt_l, S->start, wscr, S->lut_l, lane);
    {
        const int i0 = (int)S->start[1], nc = nl - nz_l;
        S->long_l[lane] = i0 + (int)lane < nc ? wscr[i0 + lane] : (uint16_t)0;
        __syncwarp();
    }
    make_lut_warp<1, FA_RD>(aux, lens + nl, nd, nz_d, S->cnt_d, S->start + 2, wscr + SORTED_L, S->lut_d, lane);
    if (fixed) { T.lbits = 9; T.dbits = 5; }
    return T;
}

enum : int { LS_CODES = 0, LS_DONE = 1, LS_FETCH = 2, LS_BLOCK = 3, LS_BUILD = 4, LS_FINISH = 5, LS_HANDOVER = 6 };

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s)
{
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));              // shift amounts >= 32 give 0
    return r;
}
__device__ __forceinline__ uint32_t bfe_u32(uint32_t v, uint32_t pos, uint32_t len)
{
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
// 16-byte asynchronous copy global -> shared (LDGSTS.128), predicated; L1 is bypassed (.cg): every byte is read once
__device__ __forceinline__ void cp_async16_if(uint32_t dst_smem, const void* src_gmem, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst_smem),
                 "l"(src_gmem), "r"((uint32_t)pred)
                 : "memory");
}
// the same with an L2 eviction policy (createpolicy): the compressed input is read exactly once
__device__ __forceinline__ void cp_async16_if(uint32_t dst_smem, const void* src_gmem, bool pred, uint64_t policy)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %3;\n\t}" ::"r"(dst_smem),
                 "l"(src_gmem), "r"((uint32_t)pred), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void stg_v4_hint(void* p, uint4 v, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void stg_v4_hint_if(void* p, uint4 v, uint64_t policy, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %6, 0;\n\t@q st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\n\t}" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy), "r"((uint32_t)pred)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint16_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
    return r;
}

// One lane per stream.  Persistent warps: a lane that finishes its stream takes the next one from the
// atomic counter; block headers, table builds and stream changes are serviced between lockstep runs
// (headers by the lanes that need one, each for itself; table builds by the whole warp, one lane's
// tables at a time).
__global__ void __launch_bounds__(32) huff_tokens_kernel(FastParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x;
    constexpr unsigned FULL = 0xffffffffu;
    LaneSmem* const lanes = reinterpret_cast<LaneSmem*>(smem_raw);
    LaneSmem* const L = lanes + lane;
    uint16_t* const wscr = reinterpret_cast<uint16_t*>(smem_raw + 32 * sizeof(LaneSmem));     // sorted_l | sorted_d | lens | aux
    uint16_t* const my_sorted = P.sorted_l + ((size_t)blockIdx.x * 32 + lane) * SORTED_L;
    const uint32_t ring_l = smem_addr(smem_raw + 32 * sizeof(LaneSmem) + SCRATCH_U16 * 2) + lane * (uint32_t)FA_RING_STRIDE;
#if SDZ_CACHE_HINTS
    const uint64_t pol_in = l2_policy_evict_first();
#endif

    // ---- lane state
    int state = LS_FETCH;
    uint32_t idx = 0;
    const uint32_t* wbase = nullptr;
    uint64_t bb = 0;
    int bc = 0;
    uint32_t nw = 0, wp = 0, lim_wp = 0, in_len = 0;
    uint32_t pos = 0, cap = 0, ntok = 0, cap_tok = 0, n_blocks = 0, dbias = 0;
    uint32_t* tokp = nullptr;
    int last = 0, eob_len = 0, lbits = 0, dbits = 0, g_l = 0, g_d = 0;
    uint32_t ck_bit = 0;                                                // (mod 2^32) a symbol boundary of the current block: its first symbol, or the last data symbol the general decoder took
    int nl = 0, nd = 0;
    bool fixed = false;
    uint32_t ci = 0;                                                    // input ring: next 16-byte chunk to request
    bool fresh = false;                                                 // the ring has to be (re)filled at the reader's position
    bool pending = false;                                               // the next symbol has to go through the general decoder
    uint32_t it = 0;

    auto refill = [&]() {
        if (bc <= 32) {
            bb |= (uint64_t)nw << bc;
            bc += 32;
            wp++;
            nw = wp < lim_wp ? __ldg(wbase + wp) : 0u;
        }
    };
    auto drop = [&](int n) { bb >>= n; bc -= n; };

    for (;;) {
        // ================================================================ service (divergent code is fine here)
        // ---- new streams
        while (state == LS_FETCH || state == LS_HANDOVER || state == LS_FINISH) {
            pending = false;                                            // (a lane can leave its stream while it waits for the general decoder)
            if (state == LS_FINISH) {
                // final block complete: trailer (src/inflate.ts:409-463) and the record - when this really is the
                // plain complete stream
                const uint8_t* src = reinterpret_cast<const uint8_t*>(wbase);
                const uint64_t consumed = (uint64_t)wp * 32 - (uint64_t)bc, total_bits = (uint64_t)in_len * 8;
                bool ok = consumed <= total_bits, stalled_eob = false;
                Container C;
                ok = ok && parse_container_clean(src, in_len, P.I.mode[idx], dict_len_of(P.I, idx), dict_adler_of(P.I, idx), P.dict_ok != 0u, C);
                uint32_t tp = (uint32_t)((consumed + 7) >> 3);
                int32_t stored = 0, isize = 0;
                if (ok) {
                    if (C.raw) {
                        // The reference looks up a symbol only when its table's index width is available, and
                        // nothing follows the final end-of-block code of a raw stream: replay that one lookup with the
                        // reference's rule (slow_lookup: root width, then the width of the sub-table a longer code sits
                        // in).  When it succeeds every earlier symbol had at least as many bits behind it (its own code plus
                        // >= lbits more; distance roots are at most 6 bits wide, hence lbits >= 6).
                        const uint64_t eob_start = consumed - (uint64_t)eob_len;
                        const uint8_t* pb = src + (eob_start >> 3);                  // (SDZ_IN_PAD bytes are readable past the stream)
                        uint64_t v = 0;
                        for (int k = 0; k < 8; k++) v |= (uint64_t)pb[k] << (8 * k);
                        const uint64_t A = total_bits - eob_start;
                        const uint32_t r = slow_lookup(L->cnt_l, my_sorted, lbits, g_l, (uint32_t)(v >> (eob_start & 7)), (int)(A < 64 ? A : 64));
                        ok = lbits >= 6 && (r >> 28) == (uint32_t)R_OK && (r & 0xffffu) == 256u && ((r >> 16) & 0xffu) == (uint32_t)eob_len;
                        if (!ok && lbits >= 6 && (r >> 28) == (uint32_t)R_STALL) {
                            // The reference cannot look the final end-of-block code up: it returns with everything before it
                            // decoded and waits for input that never comes (Z_OK, complete = false - about one raw stream in
                            // 250).  That record is written here when the LAST DATA SYMBOL passes the same rule at each of its
                            // lookups (its extra bits are its own; every symbol before it has at least lbits >= dbits bits
                            // behind its codes, and a sub-table is never wider than its root): walk from the checkpoint to the
                            // end-of-block code to find that symbol.  Anything else: general decoder.
                            uint64_t q = consumed - (uint64_t)((uint32_t)consumed - ck_bit), qs = 0;
                            uint32_t pn = 0, pxb = 0, pdn = 0, pdx = 0;
                            bool pmatch = false, walked = false, bad = q > eob_start;
                            for (int it2 = 0; it2 < 8192 && !bad && q < eob_start; it2++) {
                                const uint8_t* pq = src + (q >> 3);
                                uint64_t u = 0;
                                for (int k = 0; k < 8; k++) u |= (uint64_t)pq[k] << (8 * k);
                                u >>= (q & 7);
                                uint32_t e2 = L->lut_l[(uint32_t)u & ((1u << FA_RL) - 1u)], n2, x2 = 0;
                                bool m2 = false;
                                if (e2 == E_LONG) {
                                    const uint32_t r2 = canon_long(L->cnt_l, my_sorted, FA_RL, g_l, L->start, (uint32_t)u, L->long_l, 32);
                                    const uint32_t sy = r2 & 0xffffu;
                                    n2 = r2 >> 16;
                                    if (r2 == 0u || sy == 256u || sy > 285u) { bad = true; break; }
                                    if (sy > 256u) { const uint32_t i2 = sy - 257u; m2 = true; x2 = i2 < 8 ? 0 : (i2 == 28 ? 0 : (i2 >> 2) - 1); }
                                } else if (e2 < 0x1000u) { bad = true; break; }
                                else { n2 = e2 >> 12; x2 = (e2 >> 8) & 7u; m2 = (e2 & 0x800u) != 0u; }
                                uint32_t dn2 = 0, dx2 = 0;
                                if (m2) {
                                    const uint32_t u2 = (uint32_t)(u >> (n2 + x2));
                                    const uint32_t de2 = L->lut_d[u2 & ((1u << FA_RD) - 1u)];
                                    if (de2 < 0x1000u) {
                                        const uint32_t r2 = (de2 == E_LONG && g_d > FA_RD) ? canon_long(L->cnt_d, L->sorted_d, FA_RD, g_d, L->start + 2, u2) : 0u;
                                        const uint32_t ds = r2 & 0xffffu;
                                        if (r2 == 0u || ds > 29u) { bad = true; break; }
                                        dn2 = r2 >> 16; dx2 = ds < 4 ? 0u : (ds >> 1) - 1u;
                                    } else { dn2 = de2 >> 12; dx2 = (de2 >> 8) & 15u; }
                                }
                                qs = q; pn = n2; pxb = x2; pdn = dn2; pdx = dx2; pmatch = m2; walked = true;
                                q += n2 + x2 + dn2 + dx2;
                            }
                            if (!bad && walked && q == eob_start) {
                                const uint8_t* pq = src + (qs >> 3);
                                uint64_t u = 0;
                                for (int k = 0; k < 8; k++) u |= (uint64_t)pq[k] << (8 * k);
                                u >>= (qs & 7);
                                const uint64_t AP = total_bits - qs;
                                const uint32_t r1 = slow_lookup(L->cnt_l, my_sorted, lbits, g_l, (uint32_t)u, (int)(AP < 64 ? AP : 64));
                                bool pass = (r1 >> 28) == (uint32_t)R_OK && ((r1 >> 16) & 0xffu) == pn;
                                if (pass && pmatch) {
                                    const uint64_t AD = AP - pn - pxb;
                                    const uint32_t r2 = slow_lookup(L->cnt_d, L->sorted_d, dbits, g_d, (uint32_t)(u >> (pn + pxb)), (int)(AD < 64 ? AD : 64));
                                    pass = g_d > 0 && (r2 >> 28) == (uint32_t)R_OK && ((r2 >> 16) & 0xffu) == pdn && AD >= (uint64_t)(pdn + pdx);
                                }
                                if (pass) { ok = true; stalled_eob = true; }
                            }
                        }
                    } else {
                        const uint32_t nbytes = C.is_gzip ? 8u : 4u;
                        ok = tp + nbytes <= in_len;
                        if (ok) {
                            if (C.is_gzip) {
                                stored = (int32_t)((uint32_t)src[tp] | ((uint32_t)src[tp + 1] << 8) | ((uint32_t)src[tp + 2] << 16) | ((uint32_t)src[tp + 3] << 24));
                                isize = (int32_t)((uint32_t)src[tp + 4] | ((uint32_t)src[tp + 5] << 8) | ((uint32_t)src[tp + 6] << 16) | ((uint32_t)src[tp + 7] << 24));
                            } else {
                                stored = (int32_t)(((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) | ((uint32_t)src[tp + 2] << 8) | (uint32_t)src[tp + 3]);
                            }
                            tp += nbytes;
                        }
                    }
                    ok = ok && tp == in_len;                            // bytes after the end: append() spins
                }
                if (ok) {
                    sdz_result R;
                    R.out_off = P.I.out_off ? P.I.out_off[idx] : 0;
                    R.out_len = pos - dbias;
                    R.total_in = tp;
                    R.zstatus = stalled_eob ? SDZ_Z_OK : SDZ_Z_STREAM_END;
                    R.stored_checksum = stored;
                    R.running_checksum = 0;
                    R.stored_isize = isize;
                    R.mtime = C.mtime;
                    R.name_off = C.name_len ? C.name_off : 0;
                    R.name_len = C.name_len;
                    R.n_blocks = n_blocks;
                    R.msg_id = SDZ_MSG_NONE;
                    R.thrown_append = SDZ_THROW_NONE;
                    R.thrown_inflate = 0;
                    R.container = (uint8_t)(C.is_gzip ? SDZ_GZIP : (C.method == 0 ? SDZ_RAW : SDZ_ZLIB));
                    R.complete = stalled_eob ? 0 : 1;
                    R.checksum_state = R.size_state = R.success = R.have_running = 0;
                    for (int k = 0; k < 7; k++) R.reserved[k] = 0;
                    P.I.res[idx] = R;
                    P.ntok[idx] = ntok;
                    state = LS_FETCH;
                } else state = LS_HANDOVER;
            }
            if (state == LS_HANDOVER) {
                const unsigned long long slot = atomicAdd(P.fb_count, 1ull);
                P.fb_list[slot] = idx;
                P.ntok[idx] = NTOK_HANDED_OVER;
                state = LS_FETCH;
            }
            // LS_FETCH
            unsigned long long i = atomicAdd(P.counter_a, 1ull);
            if (i < P.count) i = P.order[P.first + i];
            else {
                // out of streams: the lane idles through the lockstep loop on zeroed tables (every lookup is a "literal"
                // of zero bits) with a full bit buffer, so it never loads and never consumes
                state = LS_DONE; lim_wp = 0xfffffff0u; wp = 0; nw = 0; bb = 0; bc = 64;
                uint32_t* z = reinterpret_cast<uint32_t*>(L->lut_l);
                for (int k = 0; k < ((1 << FA_RL) + (1 << FA_RD)) / 2; k++) z[k] = 0u;
                break;
            }
            idx = (uint32_t)i;
            in_len = P.I.in_len[i];
            const uint8_t* src = P.I.in + P.I.in_off[i];
            wbase = reinterpret_cast<const uint32_t*>(src);
            Container C;
            if (!parse_container_clean(src, in_len, P.I.mode[i], dict_len_of(P.I, i), dict_adler_of(P.I, i), P.dict_ok != 0u, C)) { state = LS_HANDOVER; continue; }
            lim_wp = (in_len + 3) / 4;
            // `pos` and `cap` count from the first byte of the dictionary: a distance is valid while it stays inside
            // dictionary + output, which is the one compare the symbol loop makes (dist - 1 >= pos: SURVEY Q6)
            dbias = C.dict_used;
            pos = dbias;
            cap = P.I.out_cap ? P.I.out_cap[i] : 0xffffffffu;
            cap = cap > 0xffffffffu - dbias ? 0xffffffffu : cap + dbias;
            ntok = 0;
            cap_tok = (uint32_t)(P.tok_off[i + 1] - P.tok_off[i]);
            tokp = P.tokens + P.tok_off[i];
            n_blocks = 0; last = 0; eob_len = 0; pending = false;
            // seek(hp)
            wp = C.hp >> 2; bb = 0; bc = 0;
            nw = wp < lim_wp ? __ldg(wbase + wp) : 0u;
            refill();
            drop((int)(C.hp & 3) * 8);
            state = LS_BLOCK;
        }
        if (__all_sync(FULL, state == LS_DONE)) break;

        // ---- block headers (src/infblocks.ts:159-523), every lane for its own stream
        if (state == LS_BLOCK) {
            refill();
            const uint32_t t3 = (uint32_t)bb & 7u;
            drop(3);
            last = (int)(t3 & 1u);
            n_blocks++;
            const uint32_t type = t3 >> 1;
            uint8_t* lens = reinterpret_cast<uint8_t*>(L->lut_l);
            state = LS_BUILD;
            if (type == 0 || type == 3) state = LS_HANDOVER;            // stored / invalid: the general decoder's
            else if (type == 1) {
                for (int k = 0; k < 320; k++) lens[k] = (uint8_t)(k < 144 ? 8 : (k < 256 ? 9 : (k < 280 ? 7 : (k < 288 ? 8 : 5))));
                nl = 288; nd = 30; fixed = true;
            } else {
                fixed = false;
                refill();
                const uint32_t t = (uint32_t)bb & 0x3fffu;
                if ((t & 0x1f) > 29 || ((t >> 5) & 0x1f) > 29) state = LS_HANDOVER;
                else {
                    drop(14);
                    nl = 257 + (int)(t & 0x1f); nd = 1 + (int)((t >> 5) & 0x1f);
                    const int ncl = 4 + (int)(t >> 10), total = nl + nd;
                    uint8_t* cl = lens + FA_HDR_CL;
                    uint8_t* blut = lens + FA_HDR_BLUT;
                    uint16_t* cnt = L->cnt_l;
                    for (int i = 0; i < 19; i++) cl[i] = 0;
                    for (int i = 0; i < 16; i++) cnt[i] = 0;
                    for (int i = 0; i < ncl; i++) { refill(); cl[c_border[i]] = (uint8_t)((uint32_t)bb & 7u); drop(3); }
                    for (int i = 0; i < 19; i++) cnt[cl[i]]++;
                    // code-length code: only complete sets here (src/inftree.ts:313-331; the reference also accepts a
                    // lone 1-bit code, SURVEY Q11: the general decoder's business)
                    int y = 1, g = 0;
                    bool over = false;
                    for (int k = 1; k <= 7; k++) { y = 2 * y - (int)cnt[k]; if (cnt[k]) g = k; if (y < 0) over = true; }
                    if (cnt[0] == 19 || over || y != 0) state = LS_HANDOVER;
                    else {
                        const int l = g;                                // = min(7, longest)
                        uint32_t code = 0;
                        for (int k = 1; k <= g; k++) {
                            for (int s = 0; s < 19; s++) {
                                if (cl[s] != k) continue;
                                const uint32_t rev = __brev(code) >> (32 - k);
                                for (uint32_t q = rev; q < (1u << l); q += (1u << k)) blut[q] = (uint8_t)(s | (k << 5));
                                code++;
                            }
                            code <<= 1;
                        }
                        int index = 0;
                        uint32_t prev = 0;
                        while (index < total) {
                            refill();
                            const uint32_t e = blut[(uint32_t)bb & ((1u << l) - 1u)];
                            const int tb = (int)(e >> 5), c = (int)(e & 31);
                            if (c < 16) {
                                drop(tb);
                                lens[index++] = (uint8_t)c;
                                prev = (uint32_t)c;
                            } else {
                                const int xi = c == 18 ? 7 : c - 14;
                                int j = c == 18 ? 11 : 3;
                                drop(tb);
                                j += (int)((uint32_t)bb & ((1u << xi) - 1u));
                                drop(xi);
                                if (index + j > total || (c == 16 && index < 1)) { state = LS_HANDOVER; break; }
                                const uint8_t v = c == 16 ? (uint8_t)prev : (uint8_t)0;
                                prev = v;
                                for (int q = 0; q < j; q++) lens[index + q] = v;
                                index += j;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- table builds: the warp builds one lane's tables at a time (src/inftree.ts:95-379)
        {
            unsigned todo = __ballot_sync(FULL, state == LS_BUILD);
            while (todo) {
                const int who = __ffs(todo) - 1;
                todo &= todo - 1;
                LaneSmem* const T = lanes + who;
                const int t_nl = __shfl_sync(FULL, nl, who), t_nd = __shfl_sync(FULL, nd, who);
                const bool t_fixed = __shfl_sync(FULL, (int)fixed, who) != 0;
                {
                    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(T->lut_l);
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(wscr + SORTED_L + SORTED_D);
                    for (uint32_t i = lane; i < 80; i += 32) d32[i] = s32[i];
                }
                __syncwarp();
                const TreeInfo TI = build_tables_warp(T, wscr, t_nl, t_nd, t_fixed, lane);
                __syncwarp();
                if (TI.msg == SDZ_MSG_NONE) {
                    uint16_t* gs = P.sorted_l + ((size_t)blockIdx.x * 32 + who) * SORTED_L;
                    for (uint32_t i = lane; i < (uint32_t)SORTED_L; i += 32) gs[i] = wscr[i];
                    T->sorted_d[lane] = wscr[SORTED_L + lane];
                    // end-of-block entries join the "rare" class of the symbol loop (one compare): 0x100 | code length
                    for (uint32_t i = lane; i < (1u << FA_RL); i += 32) {
                        const uint32_t e = T->lut_l[i];
                        if (e >= 0x1000u && (e & 0xfffu) == 0x100u) T->lut_l[i] = (uint16_t)(0x100u | (e >> 12));
                    }
                }
                __syncwarp();
                if ((int)lane == who) {
                    if (TI.msg != SDZ_MSG_NONE) state = LS_HANDOVER;    // the general decoder reproduces the message
                    else { lbits = TI.lbits; dbits = TI.dbits; g_l = TI.g_l; g_d = TI.g_d; state = LS_CODES; fresh = true; pending = false; }
                }
            }
        }
        __syncwarp();
        if (__any_sync(FULL, state >= LS_FETCH)) continue;             // hand-overs found by the header / build steps

        // ================================================================ lockstep symbol loop
        // During one run every lane is either decoding (`live`) or out of streams; the run ends for the whole warp as
        // soon as one lane meets an end-of-block code or anything the general decoder has to look at.
        //
        // Symbols are decoded in groups of four.  The four symbols are ONE straight-line PTX block without a single
        // branch or vote: the per-symbol warp votes of the first version sat on the dependent chain (table read ->
        // compare -> VOTE -> BRA -> next table read: about half of all issue slots were spent waiting on them with one
        // warp per scheduler).  A lane that meets anything rare - a code longer than the table root, end of block, an
        // invalid code, a bit buffer that does not cover a distance code with its extra bits - switches itself off for
        // the rest of the group (predicate `pa`: nothing is consumed, its token slots stay no-ops) and remembers the
        // slot; after the group ONE vote sends those lanes through the general single-symbol decoder below, whose token
        // goes into the remembered slot.  Checks that can wait (slot full, distance before the start of the output,
        // input overrun, token arena full) are accumulated and looked at once per group; phase A writes no output
        // bytes, so a late hand-over costs nothing.
        {
            const bool live = state == LS_CODES;
            const uint32_t a_l = smem_addr(L->lut_l), a_d = smem_addr(L->lut_d);
            int ev = 0;                                                 // 1: end of block, 2: hand the stream over
            // lanes that come from a block header: the ring takes over at word wp + 1 (nw already holds word wp)
            {
                const bool fill = live && fresh;
                if (fill) ck_bit = wp * 32u - (uint32_t)bc;             // the block's first symbol
                if (fill) ci = (wp + 1u) >> 2;
                #pragma unroll
                for (int k = 0; k < FA_RING_CHUNKS; k++) {
#if SDZ_CACHE_HINTS
                    cp_async16_if(ring_l + ((ci & (FA_RING_CHUNKS - 1u)) << 4), reinterpret_cast<const uint8_t*>(wbase) + (size_t)ci * 16u, fill, pol_in);
#else
                    cp_async16_if(ring_l + ((ci & (FA_RING_CHUNKS - 1u)) << 4), reinterpret_cast<const uint8_t*>(wbase) + (size_t)ci * 16u, fill);
#endif
                    ci += fill ? 1u : 0u;
                }
                cp_async_commit();
                cp_async_wait_all();
                fresh = false;
            }
            for (;;) {
                uint32_t tk0, tk1, tk2, tk3;
                uint32_t early = 0u;                                    // a distance reached before the start of the output
                // su: the slot in which the lane switched itself off.  A lane that is waiting for the general decoder
                // (`pending`) sits out whole groups: its slots stay no-ops until its symbol lands in slot 0 of the group
                // after which the decoder ran.
                uint32_t act = live && !pending ? 1u : 0u, su = pending ? 0u : 4u;
#define SDZ_FA_SYMBOL(U, TK)                                                                                              \
                    /* top-up to at least 32 valid bits */                                                                \
                    "setp.lt.s32 pt, %1, 32;\n\t"                                                                         \
                    "mov.b32 pw, 0;\n\t"                                                                                  \
                    "@pt shl.b32 pw, 1, %1;\n\t"                                                                          \
                    "mad.wide.u32 %0, %3, pw, %0;\n\t"            /* bits above bc are zero: add == or */                 \
                    "@pt add.s32 %1, %1, 32;\n\t"                                                                         \
                    "@pt add.u32 %2, %2, 1;\n\t"                                                                          \
                    "and.b32 t, %2, " SDZ_STR(SDZ_FA_RING_WMASK) ";\n\t"                                                                      \
                    "mad.lo.u32 ra, t, 4, %12;\n\t"                                                                       \
                    "@pt ld.shared.u32 %3, [ra];\n\t"                                                                     \
                    /* literal/length lookup and fields */                                                                \
                    "cvt.u32.u64 lo, %0;\n\t"                                                                             \
                    "and.b32 t, lo, " SDZ_STR(SDZ_FA_RL_MASK) ";\n\t"                                                                       \
                    "mad.lo.u32 ra, t, 2, %13;\n\t"                                                                       \
                    "ld.shared.u16 e, [ra];\n\t"                                                                          \
                    "shr.u32 n, e, 12;\n\t"                                                                               \
                    "bfe.u32 xb, e, 8, 3;\n\t"                    /* 0 for a literal */                                   \
                    "add.u32 c1, n, xb;\n\t"                                                                              \
                    "shr.u32 t, lo, n;\n\t"                                                                               \
                    "shl.b32 b, 0xffffffff, xb;\n\t"                                                                      \
                    "lop3.b32 t, t, b, 0, 0x30;\n\t"              /* t & ~b: the extra bits */                            \
                    "and.b32 b, e, 255;\n\t"                                                                              \
                    "add.u32 lenf, b, t;\n\t"                     /* length - 3 (literal: the byte) */                    \
                    "and.b32 t, e, 2048;\n\t"                                                                             \
                    "setp.ne.u32 pm, t, 0;\n\t"                                                                           \
                    /* distance: the root lookup needs 7 of the >= 12 bits that are left */                               \
                    "shr.u64 w, %0, c1;\n\t"                                                                              \
                    "cvt.u32.u64 lo2, w;\n\t"                                                                             \
                    "and.b32 t, lo2, 127;\n\t"                                                                            \
                    "mad.lo.u32 ra, t, 2, %14;\n\t"                                                                       \
                    "ld.shared.u16 de, [ra];\n\t"                                                                         \
                    "shr.u32 dn, de, 12;\n\t"                                                                             \
                    "bfe.u32 dx, de, 8, 4;\n\t"                                                                           \
                    "add.u32 c2, dn, dx;\n\t"                                                                             \
                    "sub.s32 rem, %1, c1;\n\t"                                                                            \
                    /* rare: no root entry for either code, or the buffer does not cover the distance */                 \
                    "setp.lt.s32 p1, rem, c2;\n\t"                                                                        \
                    "setp.lt.u32 p2, de, 4096;\n\t"                                                                       \
                    "or.pred p1, p1, p2;\n\t"                                                                             \
                    "and.pred p1, p1, pm;\n\t"                                                                            \
                    "setp.lt.u32 p2, e, 4096;\n\t"                                                                        \
                    "or.pred p1, p1, p2;\n\t"                                                                             \
                    "and.pred p2, pa, p1;\n\t"                                                                            \
                    "@p2 mov.b32 %7, " U ";\n\t"                                                                          \
                    "and.pred pa, pa, !p1;\n\t"                                                                           \
                    "shr.u32 t, lo2, dn;\n\t"                                                                             \
                    "shl.b32 m, 0xffffffff, dx;\n\t"                                                                      \
                    "lop3.b32 t, t, m, 0, 0x30;\n\t"                                                                      \
                    "and.b32 m, de, 3;\n\t"                                                                               \
                    "shl.b32 m, m, dx;\n\t"                                                                               \
                    "add.u32 dm1, m, t;\n\t"                      /* distance - 1 */                                      \
                    /* consumption (nothing when the lane is off), token */                                               \
                    "selp.u32 c2, c2, 0, pm;\n\t"                                                                         \
                    "add.u32 c, c1, c2;\n\t"                      /* <= 48 */                                             \
                    "selp.u32 c, c, 0, pa;\n\t"                                                                           \
                    "shr.u64 %0, %0, c;\n\t"                                                                              \
                    "sub.s32 %1, %1, c;\n\t"                                                                              \
                    "add.u32 len, lenf, 3;\n\t"                                                                           \
                    "selp.u32 len, len, 1, pm;\n\t"                                                                       \
                    "and.pred p2, pa, pm;\n\t"                                                                            \
                    "setp.ge.and.u32 p2, dm1, %4, p2;\n\t"        /* the general decoder's */                  \
                    "@p2 mov.b32 %5, 1;\n\t"                                                                              \
                    "@pa add.u32 %4, %4, len;\n\t"                                                                        \
                    "mad.lo.u32 t, dm1, 512, len;\n\t"                                                                    \
                    "or.b32 tl, lenf, 0x80000000;\n\t"                                                                    \
                    "selp.u32 t, t, tl, pm;\n\t"                                                                          \
                    "selp.u32 " TK ", t, 0, pa;\n\t"
                asm volatile("{\n\t"
                    ".reg .pred pt, pm, p1, p2, pa;\n\t"
                    ".reg .b32 pw, ra, t, lo, e, n, xb, c1, b, lenf, lo2, de, dn, dx, c2, rem, m, dm1, c, len, tl;\n\t"
                    ".reg .b64 w;\n\t"
                    "setp.ne.u32 pa, %6, 0;\n\t"
                    SDZ_FA_SYMBOL("0", "%8")
                    SDZ_FA_SYMBOL("1", "%9")
                    SDZ_FA_SYMBOL("2", "%10")
                    SDZ_FA_SYMBOL("3", "%11")
                    "}"
                    : "+l"(bb), "+r"(bc), "+r"(wp), "+r"(nw), "+r"(pos), "+r"(early), "+r"(act), "+r"(su),
                      "=r"(tk0), "=r"(tk1), "=r"(tk2), "=r"(tk3)
                    : "r"(ring_l), "r"(a_l), "r"(a_d));
#undef SDZ_FA_SYMBOL
                // ---- lanes that switched themselves off: one symbol through the general decoder.  The decoder's ~130
                // instructions are paid by the whole warp, and with 128 symbols per group some lane needs it in two groups
                // out of three (profiles/r02c_summary.md: a third of all issued instructions, 1.5 lanes active).  So it
                // runs every SDZ_FA_DEFER-th group only, or as soon as a quarter of the warp is waiting; waiting lanes sit
                // out (a lane meets a rare symbol every ~28 groups, so the wait costs it a few per cent).
                pending = su < 4u;
                const unsigned pend = __ballot_sync(FULL, pending);
                it++;
                if ((pend != 0u) & (((it & (SDZ_FA_DEFER - 1u)) == 0u) | (__popc(pend) >= 8))) {       // (one branch, not three)
                    if (pending) {
                        pending = false;
                        uint32_t tslow = 0u;
                        if (bc < 32) {
                            bb |= (uint64_t)nw << bc;
                            bc += 32; wp++;
                            nw = lds_u32(ring_l + ((wp & (4u * FA_RING_CHUNKS - 1u)) << 2));
                        }
                        const uint32_t sym_bit = wp * 32u - (uint32_t)bc;
                        const uint32_t lo = (uint32_t)bb;
                        uint32_t e = L->lut_l[lo & ((1u << FA_RL) - 1u)];
                        if (e == E_LONG) {
                            const uint32_t r = canon_long(L->cnt_l, my_sorted, FA_RL, g_l, L->start, lo, L->long_l, 32);
                            const uint32_t sym = r & 0xffffu;
                            if (r != 0u && sym < 256u) e = ((r >> 16) << 12) | sym;
                            else if (r != 0u && sym == 256u) e = 0x100u | (r >> 16);
                            else if (r != 0u && sym - 257u <= 28u) {
                                const uint32_t i = sym - 257u;
                                const uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                                const uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                                e = ((r >> 16) << 12) | 0x800u | (xb << 8) | (base - 3u);
                            } else e = E_INVALID;
                        }
                        if (e < 0x1000u) {
                            if ((e & 0xf00u) == 0x100u) {               // end of block (LUT fix-up after the build: 0x100 | code length)
                                const int n = (int)(e & 0xffu);
                                bb >>= n; bc -= n; eob_len = n;
                                ev = 1;
                            } else ev = 2;                              // invalid literal/length code
                        } else {
                            ck_bit = sym_bit;                           // (a data symbol: where LS_FINISH may start its look at the stream's last symbols)
                            const uint32_t n = e >> 12, xb = (e >> 8) & 7u;
                            const bool ismatch = (e & 0x800u) != 0u;
                            const uint32_t lenf = (e & 0xffu) + ((lo >> n) & ((1u << xb) - 1u));
                            bb >>= (n + xb); bc -= (int)(n + xb);
                            if (!ismatch) { tslow = TOK_LIT | lenf; pos += 1u; }
                            else {
                                if (bc < 32) {
                                    bb |= (uint64_t)nw << bc;
                                    bc += 32; wp++;
                                    nw = lds_u32(ring_l + ((wp & (4u * FA_RING_CHUNKS - 1u)) << 2));
                                }
                                const uint32_t lo2 = (uint32_t)bb;
                                uint32_t de = L->lut_d[lo2 & ((1u << FA_RD) - 1u)];
                                if (de < 0x1000u) {
                                    const uint32_t r = (de == E_LONG && g_d > FA_RD) ? canon_long(L->cnt_d, L->sorted_d, FA_RD, g_d, L->start + 2, lo2) : 0u;
                                    const uint32_t ds = r & 0xffffu;
                                    if (r == 0u || ds > 29u) { ev = 2; de = 0x1000u; }         // invalid distance code
                                    else de = ((r >> 16) << 12) | ((ds < 4 ? 0u : (ds >> 1) - 1u) << 8) | (ds < 4 ? ds : 2u + (ds & 1u));
                                }
                                const uint32_t dn = de >> 12, dx = (de >> 8) & 15u;
                                const uint32_t dm1 = ((de & 3u) << dx) + ((lo2 >> dn) & ((1u << dx) - 1u));
                                bb >>= (dn + dx); bc -= (int)(dn + dx);
                                if (dm1 >= pos) early = 1u;
                                pos += lenf + 3u;
                                tslow = dm1 * 512u + lenf + 3u;
                            }
                        }
                        tk0 = su == 0u ? tslow : tk0; tk1 = su == 1u ? tslow : tk1;
                        tk2 = su == 2u ? tslow : tk2; tk3 = su == 3u ? tslow : tk3;
                    }
                }
                // input ring: request the chunks whose slots have been read completely (the reader holds word wp in a
                // register: chunk ci - FA_RING_CHUNKS is free once wp has reached its last word).  A group of four symbols
                // takes at most six words, so two requests keep up; a chunk is requested >= 4.5 groups before its first
                // word is read, and wait_group 3 has completed it by then.
                #pragma unroll
                for (int k = 0; k < 2; k++) {
                    const bool need = live && 4u * ci <= wp + (4u * FA_RING_CHUNKS - 3u);
#if SDZ_CACHE_HINTS
                    cp_async16_if(ring_l + ((ci & (FA_RING_CHUNKS - 1u)) << 4), reinterpret_cast<const uint8_t*>(wbase) + (size_t)ci * 16u, need, pol_in);
#else
                    cp_async16_if(ring_l + ((ci & (FA_RING_CHUNKS - 1u)) << 4), reinterpret_cast<const uint8_t*>(wbase) + (size_t)ci * 16u, need);
#endif
                    ci += need ? 1u : 0u;
                }
                cp_async_commit();
                asm volatile("cp.async.wait_group " SDZ_STR(SDZ_FA_RING_WAIT) ";" ::: "memory");
                // (no branches: a taken or not-taken branch costs a latency-bound warp ~20 cycles, and these were three per group)
                {
                    const bool bad = live & ((early != 0u) | (pos > cap) | (wp > lim_wp + 1u) | (ntok + 4u > cap_tok));
                    const bool put = live & !bad & ((tk0 | tk1 | tk2 | tk3) != 0u);
                    SDZ_CHECK(!put || (ntok + 4u <= cap_tok && ((reinterpret_cast<uintptr_t>(tokp + ntok) & 15u) == 0u) && wp <= lim_wp + 1u));
#if SDZ_CACHE_HINTS
                    stg_v4_hint_if(tokp + ntok, make_uint4(tk0, tk1, tk2, tk3), pol_in, put);     // written once, read once by phase B
#else
                    if (put) *reinterpret_cast<uint4*>(tokp + ntok) = make_uint4(tk0, tk1, tk2, tk3);
#endif
                    ntok += put ? 4u : 0u;
                    ev = bad ? 2 : ev;
                }
                if (__any_sync(FULL, ev != 0)) break;
            }
            if (ev == 2) state = LS_HANDOVER;
            else if (ev == 1) state = last ? LS_FINISH : LS_BLOCK;
        }
    }
}

// ---------------------------------------------------------------------- phase B

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));      // read once: evict first
    return v;
}

// One warp per stream.  32 tokens per batch; their bytes are produced in rows of 32 (lane = output byte).
__global__ void __launch_bounds__(256) lz_resolve_kernel(FastParams P)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t squeeze[8][32];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t le_mask = 0xffffffffu >> (31u - lane);
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(P.counter_b, 1ull);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= P.count) break;
        idx = P.order[P.first + idx];
        const uint32_t nt = P.ntok[idx];
        if (nt == NTOK_HANDED_OVER || nt == 0u) continue;
        const uint32_t* tk = P.tokens + P.tok_off[idx];
        uint8_t* const out = P.I.out + P.I.out_off[idx];
        uint32_t ab = lane;                                            // my byte of the current row, as an offset into the stream's output
        uint32_t tnext = lane < nt ? ld_stream_u32(tk + lane) : 0u;    // (the next batch's tokens are requested one batch ahead)
        for (uint32_t base = 0; base < nt; base += 32) {
            uint32_t t = tnext;
            tnext = base + 32u + lane < nt ? ld_stream_u32(tk + base + 32u + lane) : 0u;
            {
                // no-op tokens (slots of a lane that had switched itself off in phase A) are squeezed out - the row loop
                // indexes tokens by counting starts: scatter through shared memory to the rank among the real tokens
                const uint32_t nz = __ballot_sync(FULL, t != 0u);
                if (nz != 0xffffffffu) {
                    __syncwarp();
                    if (t != 0u) squeeze[threadIdx.x >> 5][__popc(nz & (le_mask >> 1))] = t;
                    __syncwarp();
                    t = lane < (uint32_t)__popc(nz) ? squeeze[threadIdx.x >> 5][lane] : 0u;
                }
            }
            const bool lit = (int32_t)t < 0;
            const uint32_t len = lit ? 1u : (t & 511u);
            uint32_t incl = len;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); incl += lane >= (uint32_t)d ? v : 0u; }
            const uint32_t N = __shfl_sync(FULL, incl, 31);            // bytes of this batch
            // what a byte needs to know about its token: a literal keeps its (negative) word, a match becomes its distance
            const uint32_t tinfo = lit ? t : (t >> 9) + 1u;
            uint32_t first = len ? incl - len : 0xffffffffu;           // row-relative position of my token's first byte
            const uint32_t ab_end = ab - lane + N;                     // end of the batch
            uint32_t row0 = ab - lane;                                 // first byte of the row (uniform)
            uint32_t nxt = 0xffffffffu;                                // (tokens that started in earlier rows) - 1
            for (uint32_t R = 0; R < N; R += 32) {
                const uint32_t m = __reduce_or_sync(FULL, shl_clamp(1u, first));       // bit b: a token starts at byte b of this row
                const uint32_t ti = nxt + __popc(m & le_mask);          // no start at or before my byte: the token that began in an earlier row
                const uint32_t tt = __shfl_sync(FULL, tinfo, ti);
                nxt += __popc(m);
                const bool valid = ab < ab_end;
                const bool tmatch = (int32_t)tt >= 0;
                const uint32_t so = ab - tt;                            // source byte (phase A guarantees dist <= position)
                const bool inrow = valid && tmatch && so >= row0;
                uint32_t v = tt & 0xffu;
                if (valid && tmatch && so < row0) v = out[so];
                if (__any_sync(FULL, inrow)) {
                    // sources inside this row: pointer jumping over the lanes (depth <= 31: five doublings)
                    uint32_t q = inrow ? so - row0 : lane;
                    for (int r = 0; r < 6; r++) {
                        const uint32_t vv = __shfl_sync(FULL, v, q), qq = __shfl_sync(FULL, q, q);
                        const bool moved = __any_sync(FULL, qq != q);
                        if (q != lane) { v = vv; q = qq; }
                        if (!moved) break;
                    }
                }
                if (valid) out[ab] = (uint8_t)v;
                __syncwarp();
                first -= 32u; ab += 32u; row0 += 32u;
            }
            ab = ab_end + lane;
        }
    }
}

// ---------------------------------------------------------------------- phase B, token-centric (default)
//
// lz_resolve_kernel above is byte-centric: lane = output byte, one byte gather per lane and row, 39 warp instructions
// and one exposed global-memory round trip per 32 output bytes (profiles/r02c_summary.md: 7.15 G warp instructions per
// 65,536-stream batch, long_scoreboard 16.6 of 25 stall cycles per issued instruction, window reads missing L2 because
// 9,472 resident streams x 32 KiB of live window do not fit).  Here lane = TOKEN:
//
//   * the stream's write frontier lives in a 1 KiB shared-memory ring per warp; tokens write their bytes into the ring
//     (literals one byte, matches up to 16 bytes read as three aligned 8-byte words from HBM/L2 - or from the ring when
//     the source is less than 480 bytes back - shifted into place and stored byte-wise), all 32 tokens of a batch at once:
//     one global round trip per ~140 output bytes, and the loads of 32 matches in flight together;
//   * complete 16-byte vectors leave the ring as coalesced 16-byte vector stores (STG.128), the only global stores of
//     the kernel apart from the unaligned head / tail of a stream;
//   * tokens that depend on bytes of their own batch, are longer than 16 bytes or overlap themselves (dist < len) are
//     "hard": after the parallel step they are replayed one at a time, in order, by the whole warp (lane = byte of the
//     match, lane-strided replicate src[i mod dist] for overlapping copies).
// Positions are kept in g-coordinates: g = stream position + (address of the stream's first byte & 15), so that
// multiples of 16 are 16-byte aligned global addresses whatever the caller's out_off is.
constexpr uint32_t B2_RING = 1024;     // bytes per warp (ring index = g & 1023), + 16 bytes of spill behind it
constexpr uint32_t B2_HIST = 480;      // sources at most this far behind the frontier are read from the ring
constexpr uint32_t B2_SUB = 512;       // bytes produced per parallel step (ring size - history - alignment slack)
constexpr int B2_WARPS = 4;
#ifndef SDZ_B2_MINBLOCKS
#define SDZ_B2_MINBLOCKS 12            // blocks per SM the register allocation aims for (12: 40 registers: eight blocks fit next to phase A)
#endif

__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// bytes K .. K + 3 of a piece: word >> 0 / 8 / 16 / 24 to [addr + K ..], byte b only when n > K + b (predicated, no branches)
#define SDZ_STS4(K)                                                                                                           \
    __device__ __forceinline__ void sts4_##K(uint32_t addr, uint32_t w, uint32_t n)                                           \
    {                                                                                                                         \
        asm volatile("{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 a, b, c;\n\t"                                               \
                     "setp.gt.u32 p0, %2, " #K " + 0;\n\tsetp.gt.u32 p1, %2, " #K " + 1;\n\t"                                   \
                     "setp.gt.u32 p2, %2, " #K " + 2;\n\tsetp.gt.u32 p3, %2, " #K " + 3;\n\t"                                   \
                     "shr.u32 a, %1, 8;\n\tshr.u32 b, %1, 16;\n\tshr.u32 c, %1, 24;\n\t"                                        \
                     "@p0 st.shared.u8 [%0 + " #K " + 0], %1;\n\t@p1 st.shared.u8 [%0 + " #K " + 1], a;\n\t"                     \
                     "@p2 st.shared.u8 [%0 + " #K " + 2], b;\n\t@p3 st.shared.u8 [%0 + " #K " + 3], c;\n\t}" ::"r"(addr),       \
                     "r"(w), "r"(n)                                                                                           \
                     : "memory");                                                                                             \
    }
SDZ_STS4(0)
SDZ_STS4(4)
SDZ_STS4(8)
SDZ_STS4(12)
#undef SDZ_STS4
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr)
{
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint2 ldg_v2(const uint8_t* p)
{
    uint2 r;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

// DICT: the batch has a dictionary arena (some streams may carry a preset dictionary); batches without one run the
// kernel without the three places that look at it.
template <bool DICT>
__global__ void __launch_bounds__(32 * B2_WARPS, SDZ_B2_MINBLOCKS) lz_resolve2_kernel(FastParams P)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) uint8_t ring_all[B2_WARPS][B2_RING + 16];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ring = smem_addr(ring_all[threadIdx.x >> 5]);
#if SDZ_CACHE_HINTS
    const uint64_t pol_out = l2_policy_evict_last();
#endif
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(P.counter_b, 1ull);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= P.count) break;
        idx = P.order[P.first + idx];
        const uint32_t nt = P.ntok[idx];
        if (nt == NTOK_HANDED_OVER || nt == 0u) continue;
        const uint32_t* tk = P.tokens + P.tok_off[idx];
        uint8_t* const out = P.I.out + P.I.out_off[idx];
        const uint32_t gb = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);   // g-coordinate of the stream's first byte
        uint8_t* const outg = out - gb;                                 // g-coordinate 0: a 16-byte aligned address
#ifdef SDZ_CHECKED
        const uint32_t chk_cap = P.I.out_cap ? P.I.out_cap[idx] : 0xfffffff0u;
#endif
        // preset dictionary: its last byte sits at stream position -1 (phase A only lets a distance reach back that far
        // when the stream's FDICT header named the caller's dictionary)
        const uint32_t dlen = DICT ? dict_len_of(P.I, idx) : 0u;
        const uint8_t* const dend = dlen ? P.I.dict + P.I.dict_off[idx] + dlen : nullptr;
        uint32_t pos = gb;                                              // write frontier (uniform)
        uint32_t flushed = 0;                                           // bytes below are in global memory; multiple of 16
        __syncwarp();                                                   // the previous stream's ring reads are over
        uint32_t tnext = lane < nt ? ld_stream_u32(tk + lane) : 0u;
        for (uint32_t base = 0; base < nt; base += 32) {
            const uint32_t t = tnext;
            tnext = base + 32u + lane < nt ? ld_stream_u32(tk + base + 32u + lane) : 0u;
            const bool lit = (int32_t)t < 0;
            const uint32_t len = lit ? 1u : (t & 511u);                 // 0: no-op slot
            uint32_t incl = len;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); incl += lane >= (uint32_t)d ? v : 0u; }
            const uint32_t N = __shfl_sync(FULL, incl, 31);
            const uint32_t excl = incl - len;
            const uint32_t dist = (t >> 9) + 1u;                        // (matches only)
            uint32_t done = 0;
            while (done < N) {                                          // one pass unless the batch is longer than B2_SUB bytes
                // ---- the tokens of this step: the longest prefix of the remaining ones that fits B2_SUB bytes
                const bool mine = len != 0u && excl >= done && incl - done <= B2_SUB;
                const uint32_t step_end = N - done <= B2_SUB ? N : __reduce_max_sync(FULL, mine ? incl : done);
                const uint32_t dst = pos + (excl - done);               // g-coordinate of my token's first byte
                const uint32_t src = dst - dist;
                const uint32_t rd = ring + (dst & (B2_RING - 1u));
                const bool match = mine && !lit;
                // simple: the whole source lies before this step's bytes and the copy is one 16-byte piece
                const uint32_t o = src & 7u, a = src - o;
                bool simple = match && len <= 16u && src + len <= pos && a >= gb;
                if (DICT && dlen) simple = simple && dist <= dst - gb;  // (dist > dst - gb: the source starts in the dictionary; `src` has wrapped)
                const bool hard = match && !simple;
                if (mine && lit) sts_u8(rd, t);
                // ---- parallel step: every simple match
                {
                    const bool from_ring = src + B2_HIST >= pos;
                    const bool third = o + len > 16u;
                    uint2 w0 = make_uint2(0u, 0u), w1 = w0, w2 = w0;
                    if (simple) {
                        SDZ_CHECK(src >= gb && src + len <= pos && dst + len <= gb + chk_cap && (from_ring || a + (third ? 24u : 16u) <= flushed));
                        SDZ_CHECK(!from_ring || pos - a <= B2_HIST + 8u);
                        if (from_ring) {
                            w0 = lds_v2(ring + (a & (B2_RING - 1u)));
                            w1 = lds_v2(ring + ((a + 8u) & (B2_RING - 1u)));
                            if (third) w2 = lds_v2(ring + ((a + 16u) & (B2_RING - 1u)));
                        } else {
                            w0 = ldg_v2(outg + a);
                            w1 = ldg_v2(outg + a + 8u);
                            if (third) w2 = ldg_v2(outg + a + 16u);
                        }
                    }
                    const uint32_t sh = (o & 3u) * 8u;
                    const uint32_t y0 = __funnelshift_r(w0.x, w0.y, sh), y1 = __funnelshift_r(w0.y, w1.x, sh),
                                   y2 = __funnelshift_r(w1.x, w1.y, sh), y3 = __funnelshift_r(w1.y, w2.x, sh),
                                   y4 = __funnelshift_r(w2.x, w2.y, sh);
                    const bool hi = (o & 4u) != 0u;
                    const uint32_t x0 = hi ? y1 : y0, x1 = hi ? y2 : y1, x2 = hi ? y3 : y2, x3 = hi ? y4 : y3;
                    const uint32_t sl = simple ? len : 0u;
                    const uint32_t lmax = __reduce_max_sync(FULL, sl);
                    if (lmax > 0u) sts4_0(rd, x0, sl);
                    if (lmax > 4u) sts4_4(rd, x1, sl);
                    if (lmax > 8u) sts4_8(rd, x2, sl);
                    if (lmax > 12u) sts4_12(rd, x3, sl);
                    // a simple match that ran over the end of the ring wrote into the 16 spill bytes: bring them round
                    const unsigned sp = __ballot_sync(FULL, simple && (dst & (B2_RING - 1u)) + len > B2_RING);
                    if (sp) {
                        const uint32_t cnt = __shfl_sync(FULL, (dst & (B2_RING - 1u)) + len - B2_RING, __ffs(sp) - 1);
                        __syncwarp();
                        if (lane < cnt) sts_u8(ring + lane, lds_u8(ring + B2_RING + lane));
                    }
                }
                __syncwarp();
                // ---- hard matches, in order, the warp on one at a time (lane = byte of the match)
                unsigned hm = __ballot_sync(FULL, hard);
                while (hm) {
                    const int k = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const uint32_t h_dst = __shfl_sync(FULL, dst, k), h_len = __shfl_sync(FULL, len, k), h_dist = __shfl_sync(FULL, dist, k);
                    const uint32_t h_src = h_dst - h_dist;
                    if (DICT && h_dist > h_dst - gb) {
                        // the source starts in the preset dictionary (and may run on into the stream's first bytes)
                        for (uint32_t c = 0; c < h_len; c += 32u) {
                            const uint32_t i = c + lane;
                            if (i < h_len) {
                                const uint32_t q = h_src + (h_dist < 32u ? i % h_dist : i);
                                const int32_t back = (int32_t)(gb - q);   // > 0: that many bytes before the stream's first byte
                                SDZ_CHECK((back > 0 ? (uint32_t)back <= dlen : q < h_dst + i) && h_dst + h_len <= gb + chk_cap &&
                                          (back > 0 || q < flushed || h_dst + i - q < B2_RING - 16u));
                                const uint32_t v = back > 0 ? (uint32_t)dend[-back] : (q >= flushed ? lds_u8(ring + (q & (B2_RING - 1u))) : (uint32_t)outg[q]);
                                sts_u8(ring + ((h_dst + i) & (B2_RING - 1u)), v);
                            }
                            __syncwarp();
                        }
                        continue;
                    }
                    for (uint32_t c = 0; c < h_len; c += 32u) {
                        const uint32_t i = c + lane;
                        if (i < h_len) {
                            // dist >= 32: bytes at or after h_dst were written by earlier rounds of this loop;
                            // dist < 32: replicate the dist bytes before the match
                            const uint32_t q = h_src + (h_dist < 32u ? i % h_dist : i);
                            SDZ_CHECK(h_src >= gb && q < h_dst + i && h_dst + h_len <= gb + chk_cap && (q < flushed || h_dst + i - q < B2_RING - 16u));
                            const uint32_t v = q >= flushed ? lds_u8(ring + (q & (B2_RING - 1u))) : (uint32_t)outg[q];
                            sts_u8(ring + ((h_dst + i) & (B2_RING - 1u)), v);
                        }
                        __syncwarp();
                    }
                }
                // ---- complete 16-byte vectors leave the ring
                pos += step_end - done;
                done = step_end;
                const uint32_t upto = pos & ~15u;
                for (uint32_t g = flushed + lane * 16u; g < upto; g += 512u) {
                    const uint4 v = lds_v4(ring + (g & (B2_RING - 1u)));
                    SDZ_CHECK(g + 16u <= pos && pos <= gb + chk_cap && pos - flushed <= B2_SUB + 16u);
#if SDZ_CACHE_HINTS
                    if (g >= gb) stg_v4_hint(outg + g, v, pol_out);    // the next 32 KiB of this stream read from here
#else
                    if (g >= gb) *reinterpret_cast<uint4*>(outg + g) = v;
#endif
                    else {                                              // the vector that holds the stream's first byte
                        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                        #pragma unroll
                        for (uint32_t b = 0; b < 16u; b++)
                            if (g + b >= gb) outg[g + b] = (uint8_t)(w[b >> 2] >> ((b & 3u) * 8u));
                    }
                }
                flushed = upto;
                __syncwarp();
            }
        }
        // tail of the stream: the bytes of the last, incomplete vector
        if (flushed + lane < pos && flushed + lane >= gb) outg[flushed + lane] = (uint8_t)lds_u8(ring + ((flushed + lane) & (B2_RING - 1u)));
    }
}

}  // namespace sdz
