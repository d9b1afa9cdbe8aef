// inflate_kernel.cuh - batched inflate for sm_100a: one sub-warp group of G lanes per stream.
//
// Replaces, per stream, the reference's whole L2 stack for a single-buffer append():
//   container header/trailer state machine      src/inflate.ts:132-473
//   block-header parser (stored/fixed/dynamic)  src/infblocks.ts:123-628
//   Huffman table construction                  src/inftree.ts:95-379
//   symbol decode + LZ77 copy                   src/infcodes.ts:62-676
// It is NOT a port of that state machine: the stream is decoded linearly into its output
// slot in HBM (no 32 KiB ring, no 16 KiB output buffer).  What the reference's ring and
// chunking make observable (Q2, Q3, Q6, Q15) is reproduced by a tiny
// scalar model of its window pointers (RingModel) and by its lookahead rule in the stream
// tail, so that records stay bit-exact wherever the reference terminates.
//
// Data layout / staging:
//   * compressed input is staged per group into a 2 x 128 B shared-memory ring by TMA bulk
//     copies (cp.async.bulk + mbarrier); the decoder keeps a 64-bit bit buffer and one
//     prefetched word in registers, so the ring read is off the decode critical path;
//   * per-group decode tables live in shared memory: a 2^RL-entry u16 literal/length LUT and
//     a 2^RD-entry u16 distance LUT with base/extra pre-baked; codes longer than the root
//     fall back to a canonical (count/sorted-symbol) decode;
//   * every lane of a group holds the same decoder state (no shuffles on the critical path);
//     lanes only differ in which bytes of a match / stored block they move;
//   * the groups of a warp run the symbol loop in lockstep and one iteration is straight-line
//     predicated code (Decoder::step_flat): literal groups execute the match path with its side
//     effects switched off, so the warp never diverges between literals and matches.
// Knobs that were measured and rejected (piecewise long matches, a third staging slot, G = 2 / 8,
// an if-form top-up, FMA-pipe address arithmetic) are recorded in DESIGN.md section 6, not kept here.
#pragma once
#include "sdz_device.cuh"
#include "../../include/sdz_codes.h"

namespace sdz {

#ifndef SDZ_LIT_RUN
#define SDZ_LIT_RUN 2                  // plain literals folded in front of every lockstep symbol (branchy loop: 0 -> 90, 1 -> 99, 2 -> 97 GB/s;
                                       // straight-line loop, where a fold is ~13 predicated instructions: 1 -> 127.0, 2 -> 128.8 GB/s)
#endif
#ifndef SDZ_CAPMARGIN
#define SDZ_CAPMARGIN 18               // step_flat(): output room below which step_general() takes over (fold + literal or a deferred
                                       // match of <= 16 bytes always fit, so the fast path needs no room test: 124.7 -> 127.0 GB/s)
#endif
#ifndef SDZ_MARK_WIDE
#define SDZ_MARK_WIDE 1                // marker mode: 8 bytes (4 symbols) per lane in the deferred copy path
#endif
#ifndef SDZ_TWOSLOT
#define SDZ_TWOSLOT 1                  // far matches of 17..32 bytes take both staging slots (one half each) instead of a synchronous copy
#endif
#ifndef SDZ_LONG_SMEM
#define SDZ_LONG_SMEM 32               // literal/length symbols with codes longer than the root kept in shared memory (0 or 32):
                                       // 8 blocks x (16 x 1,760 B + 1 KiB) = exactly the 228 KiB of one SM (+1.5 %)
#endif
constexpr int STAGE_LONG_WORDS = 32;   // staging slots + long-code symbols share 128 bytes per stream (GroupSmem::stage_long)
constexpr int RL = 9;                  // literal/length LUT root bits
constexpr int RD = 7;                  // distance LUT root bits
constexpr int CH = 128;                // bytes per TMA bulk copy
constexpr int NBUF = 2;                // chunks in the per-group input ring
constexpr int CHW = CH / 4;
constexpr int WSIZE = 32768;           // reference window (src/inflate.ts:98)
constexpr int OUTBUF = 16384;          // reference OUTPUT_BUFSIZE (src/zstream.ts:11)

constexpr uint32_t E_LONG = 0x0ffe;    // root entry: code longer than the root -> canonical decode
constexpr uint32_t E_INVALID = 0x0fff; // root entry: no code has this prefix

struct alignas(16) GroupSmem {
    uint32_t ring[NBUF * CHW];         // 256 B input staging (TMA bulk copies)
    uint64_t mbar[NBUF];
    uint16_t lut_l[1 << RL];
    uint16_t lut_d[1 << RD];
    uint16_t cnt_l[16];                // [1..15] codes per length (unpadded)
    uint16_t cnt_d[16];
    uint16_t start[4];                 // canonical-walk state after the root bits: first_l, index_l, first_d, index_d
    uint16_t pad_[4];
    // staging slots of the pending (deferred) matches, then - in what is left of the 128 bytes - the first symbols (canonical
    // order) whose literal/length code is longer than RL bits.  Byte mode: 2 slots x 8 words + 32 symbols; marker mode
    // (two-byte symbols, 8 bytes per lane): 2 slots x 12 words + 16 symbols; see Decoder::SLOTW / LONG_OFF / LONG_N.
    uint32_t stage_long[STAGE_LONG_WORDS];
};
constexpr int MAX_G_DEFERRED = 4;      // groups wider than this use the synchronous copy only (stage[] holds 8 words)

// per-group scratch in global memory: symbols ordered by (code length, symbol) - read by the
// table build and by the rare canonical (slow) decode - and the block's code-length array
constexpr int SORTED_L = 288, SORTED_D = 32;
constexpr int SCRATCH_U16 = SORTED_L + SORTED_D + 160 + 64;  // + 320 bytes of code lengths + 32 words of build scratch

struct InflateParams {
    const uint8_t* in;
    const uint64_t* in_off;
    const uint32_t* in_len;
    const uint8_t* mode;               // enum sdz_mode | 0x80 when a dictionary was supplied
    const uint8_t* dict;
    const uint64_t* dict_off;
    const uint32_t* dict_len;
    const int32_t* dict_adler;
    uint8_t* out;                      // nullptr: sizing pass
    const uint64_t* out_off;
    const uint32_t* out_cap;
    sdz_result* res;
    uint16_t* scratch;                 // SCRATCH_U16 entries per group of the grid
    // block-task modes (large single stream, sdz_inflate_large): every work item is (a piece of) ONE
    // deflate block of stream 0 whose header starts at bit task_bit[i].
    //   TM_INDEX  count-only walk of the whole block; every `ckpt_step` output bytes a resume point
    //             (bit position, output position) is appended to ckpt[]
    //   TM_MARK   decode from the block header (task_resume[i] == 0) or from a resume point, for exactly
    //             task_limit[i] output symbols, into out16 + task_out[i]
    const uint64_t* task_bit;
    const uint64_t* task_resume;       // TM_MARK: bit position to continue at once the block's tables are built (0: none)
    const uint64_t* task_out;          // TM_MARK: absolute output offset of the piece
    const uint32_t* task_limit;        // TM_MARK: output symbols of the piece
    uint16_t* out16;                   // TM_MARK output: byte value, or 256 + index into the 32 KiB window before the piece
    struct Ckpt* ckpt;                 // TM_INDEX
    unsigned long long* ckpt_count;
    unsigned long long ckpt_cap;
    uint32_t ckpt_step;                // power of two
    unsigned long long n;
    unsigned long long* counter;       // dynamic stream scheduler
    // hand-over from the two-phase fast path (fast_kernels.cuh): work item i is stream list[i], and the number of
    // work items is only known on the device
    const uint32_t* list;
    const unsigned long long* n_dev;
    // streaming sessions (sdz_inflater_*, src/sd-inflate.ts:87-153): where every stream stopped at the end of the input it
    // had (resume_out), and where to pick it up now that the input has grown (resume_in).  nullptr: plain one-shot decode.
    const sdz_resume* resume_in;
    sdz_resume* resume_out;
    uint32_t spec;                     // SDZ_PARITY_SPEC: zlib 1.3's behaviour instead of the reference's (general decoder only)
};

// resume point inside a block (TM_INDEX): the symbol at bit `bit` produces output byte `pos` of task `task`
struct Ckpt {
    uint64_t bit;
    uint32_t task;
    uint32_t pos;
};

enum { TM_NONE = 0, TM_MARK = 1, TM_INDEX = 2 };

// how a decode step ended
enum : int { R_OK = 0, R_EOB = 1, R_STALL = 2, R_ERROR = 3, R_OUTFULL = 4 };

// where input ran out (only the classes the record depends on)
enum : int { ST_NONE = 0, ST_OTHER = 1 /* a block header state the reference can re-enter */, ST_DYNHDR = 2 /* BTREE/DTREE: not
              resumable, SURVEY Q3 */, ST_STORED = 3 /* inside stored data: `left` is lost on re-entry (src/infblocks.ts:134) */,
              ST_CODES = 4 /* inside a symbol: InfCodes.proc is resumable at any bit */ };

__device__ __constant__ uint8_t c_border[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// Scalar model of InfBlocks' window pointers and ZStream.avail_out (src/infblocks.ts:61-121,
// :154, :289-313; src/infcodes.ts:547-573).  Only byte COUNTS flow through it.
struct RingModel {
    int q, r, ao;                      // write, read, avail_out
    __host__ __device__ __forceinline__ void init(int dict_used) { q = r = dict_used; ao = OUTBUF; }
    __host__ __device__ __forceinline__ int room() const { return q < r ? r - q - 1 : WSIZE - q; }
    __host__ __device__ __forceinline__ void flush()
    {
        int n = (r <= q ? q : WSIZE) - r;
        n = n < ao ? n : ao; ao -= n; r += n;
        if (r == WSIZE) {
            r = 0;
            if (q == WSIZE) q = 0;
            n = q - r < ao ? q - r : ao; ao -= n; r += n;
        }
    }
    // the "no room" dance; returns how many times proc() went back to append()
    __host__ __device__ __forceinline__ int make_room()
    {
        int returns = 0;
        for (;;) {
            if (room()) break;
            if (q == WSIZE && r != 0) q = 0;
            if (room()) break;
            flush();
            if (q == WSIZE && r != 0) q = 0;
            if (room()) break;
            flush(); ao = OUTBUF; returns++;      // return to append(): fresh 16 KiB buffer
        }
        return returns;
    }
    __host__ __device__ __forceinline__ void write(uint32_t n)
    {
        while (n) {
            int m = room();
            if (!m) { make_room(); m = room(); }
            uint32_t t = n < (uint32_t)m ? n : (uint32_t)m;
            q += (int)t; n -= t;
        }
    }
    // WASH / DRY: everything must leave the window before the block ends (src/infcodes.ts:626-639)
    __host__ __device__ __forceinline__ void wash()
    {
        flush();
        while (r != q) { flush(); ao = OUTBUF; flush(); }
    }
};


// ---------------------------------------------------------------------- reference table geometry
// (pure functions of the code-length counts; kept out of line so that the decoder state
// stays in registers)

// Width of the sub-table huft_build creates for the codes that share the first `w` bits
// `prefix` (MSB-first) - src/inftree.ts:217-239.  cnt[] unpadded, `pad` dummy codes at g.
// Returns 0 when no code has that prefix.
__device__ __noinline__ int ref_subtable_width(const uint16_t* cnt, int g, int pad, int l, int w, uint32_t prefix)
{
    uint32_t fc = 0;
    for (int k = 1; k <= g; k++) {
        uint32_t ck = cnt[k] + (k == g ? (uint32_t)pad : 0u);
        if (k > w) {
            uint32_t lo = prefix << (k - w), hi = lo + (1u << (k - w));
            uint32_t a0 = max(lo, fc), a1 = min(hi, fc + ck);
            if (a0 < a1) {
                int a = (int)(ck - (a0 - fc)) - 1;
                int z = min(g - w, l);
                int j = k - w;
                int f = 1 << j;
                if (f > a + 1) {
                    f -= a + 1;
                    int xp = k;
                    if (j < z) {
                        while (++j < z) {
                            f <<= 1;
                            ++xp;
                            int cx = (int)(cnt[xp] + (xp == g ? (uint32_t)pad : 0u));
                            if (f <= cx) break;
                            f -= cx;
                        }
                    }
                }
                return j;
            }
        }
        fc = (fc + ck) << 1;
    }
    return 0;
}

// Number of (exop, bits, base) entries huft_build allocates for this code set
// (src/inftree.ts:217-246): the root table plus one sub-table per distinct l-bit (and, for
// very long codes, 2l-bit) prefix.  Evaluated by the group's lanes in parallel.
template <int G>
__device__ __noinline__ int ref_table_total(const uint16_t* cnt, int g, int pad, int l, int glane, unsigned gmask)
{
    int total = 1 << l;
    for (int w = l; w < g; w += l) {
        uint32_t fc = 0, pmin = 1u << w;
        for (int k = 1; k <= g; k++) {
            uint32_t ck = cnt[k] + (k == g ? (uint32_t)pad : 0u);
            if (k > w && ck) { pmin = fc >> (k - w); break; }
            fc = (fc + ck) << 1;
        }
        int part = 0;
        for (uint32_t P = pmin + (uint32_t)glane; P < (1u << w); P += G) {
            int j = ref_subtable_width(cnt, g, pad, l, w, P);
            if (j) part += 1 << j;
        }
        #pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o, G);
        total += part;
        if (total > 4 * 1400) break;
    }
    return total;
}

// Decode one code the slow way (canonical counts), applying the reference's lookahead rule
// (SURVEY Q15): a lookup happens only when the table's index width is available.
// bits = next 32 bits of the stream (LSB first), A = bits left in the input (capped).
// Returns status << 28 | index width needed << 24 | code_length << 16 | symbol; status = R_OK / R_STALL / R_ERROR.
__device__ __noinline__ uint32_t slow_lookup(const uint16_t* cnt, const uint16_t* sorted, int l, int g, uint32_t bits, int A, bool spec = false)
{
    if (A < l && !spec) return (uint32_t)R_STALL << 28;
    int ncodes = 0, y = 1;
    for (int k = 1; k <= g; k++) { ncodes += (int)cnt[k]; y <<= 1; y -= (int)cnt[k]; }
    const int pad = y;                                  // unused codes of length g
    if (g == 1 && ncodes == 1) {                        // the one incomplete set the reference accepts
        if (bits & 1) return (uint32_t)R_ERROR << 28;   // exop 192: invalid code
        return (1u << 24) | (1u << 16) | (sorted[0] & 0xfffu);
    }
    int code = 0, first = 0, index = 0, klen = 0;
    bool found = false;
    uint32_t sym = 0;
    int lim = min(A, 15);
    for (int len = 1; len <= lim; len++) {
        code |= (int)((bits >> (len - 1)) & 1);
        int count = len <= g ? (int)cnt[len] : 0;
        if (code - count < first) {
            sym = sorted[index + (code - first)] & 0xfffu;
            klen = len; found = true;
            break;
        }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    const uint32_t ok = ((uint32_t)klen << 16) | sym;
    if (spec) {
        // zlib decodes a symbol as soon as its own bits are in the buffer (no root-width lookahead)
        if (found) return ok | ((uint32_t)klen << 24);
        return (uint32_t)(A >= g ? R_ERROR : R_STALL) << 28;
    }
    if (found && klen <= l) return ok | ((uint32_t)l << 24);
    if (!found && A >= g) return (uint32_t)R_ERROR << 28;   // every bit of the longest code is there: no such code
    // walk the reference's table levels
    int w = l;
    for (int level = 0; level < 4; level++) {
        if (A - w < 1) return (uint32_t)R_STALL << 28;
        uint32_t prefix = __brev(bits) >> (32 - w);
        int j = ref_subtable_width(cnt, g, pad, l, w, prefix);
        if (A - w < j) return (uint32_t)R_STALL << 28;
        if (found && klen <= w + j) return ok | ((uint32_t)(w + j) << 24);      // bits 24..27: index width of the last table level
        if (j == 0 || (!found && w + j >= 15)) return (uint32_t)R_ERROR << 28;
        w += l;
    }
    return (uint32_t)R_ERROR << 28;
}

// Long code (more than R root bits) away from the stream tail: canonical decode that starts at
// length R + 1.  `start` = state of the canonical walk after R bits, precomputed per block:
// { first code value (<< 1), symbol index }.
// Returns code_length << 16 | symbol, or 0 if no code matches (invalid).
// `near` (optional): shared-memory copy of sorted[start[1] .. start[1] + n_near).
__device__ __forceinline__ uint32_t canon_long(const uint16_t* cnt, const uint16_t* sorted, int R, int g, const uint16_t* start, uint32_t bits,
                                               const uint16_t* near = nullptr, int n_near = 0)
{
    int first = (int)start[0], index = (int)start[1];
    const int index0 = index;
    int code = (int)((__brev(bits) >> (32 - R)) << 1);
    for (int len = R + 1; len <= g; len++) {
        code |= (int)((bits >> (len - 1)) & 1u);
        const int count = (int)cnt[len];
        if (code - count < first) {
            const int at = index + (code - first);
            const uint32_t sym = (near != nullptr && at - index0 < n_near) ? near[at - index0] : sorted[at];
            return ((uint32_t)len << 16) | (sym & 0xfffu);
        }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return 0u;
}

// First touch of input chunk `c` (absolute index in the stream): wait for its TMA copy, then
// reuse the slot of the chunk that was just finished for the next outstanding chunk.
// Slot = c % NBUF; `phasebits` holds the mbarrier parity to wait for next on every slot.
// Returns issued_abs (low 16.. bits are not packed: plain value); phasebits updated by reference
// would force it to memory, so both come back packed: (issued_abs << 4) | phasebits.
__device__ __noinline__ uint64_t chunk_cross(GroupSmem* S, const uint8_t* gsrc, uint32_t c, uint32_t chunk0, uint32_t issued_abs,
                                             uint32_t total_chunks, uint32_t phasebits, unsigned gmask, int glane)
{
    const uint32_t slot = c % NBUF;
    mbar_wait(&S->mbar[slot], (phasebits >> slot) & 1u);
    phasebits ^= 1u << slot;
    if (c > chunk0 && issued_abs < total_chunks) {
        __syncwarp(gmask);                              // every lane is done with the slot being recycled
        if (glane == 0) {
            const uint32_t ns = issued_abs % NBUF;
            mbar_arrive_expect_tx(&S->mbar[ns], CH);
            bulk_copy_g2s(&S->ring[ns * CHW], gsrc + (size_t)issued_abs * CH, CH, &S->mbar[ns]);
        }
        issued_abs++;
    }
    return ((uint64_t)issued_abs << 4) | phasebits;
}

// ---------------------------------------------------------------------- table construction
// (out of line: runs once per block; works only on the group's shared memory)

struct TreeInfo {
    int lbits, dbits, g_l, g_d;        // reference root widths / longest codes (src/inftree.ts:146-165)
    int msg;
};

// Reference acceptance test for one code-length set (src/inftree.ts:131-178,:298) and the
// per-length counts.  Returns 0 ok, 1 oversubscribed, 2 incomplete, 3 empty.
template <int G>
__device__ __forceinline__ int classify(const uint8_t* lens, int n, int want_bits, uint16_t* cnt, uint32_t* gcount, int* l_out,
                                        int* g_out, int* pad_out, int* nzero_out, int glane, unsigned gmask)
{
    for (int i = glane; i < 16; i += G) gcount[i] = 0;
    __syncwarp(gmask);
    for (int i = glane; i < n; i += G) atomicAdd(&gcount[lens[i]], 1u);
    __syncwarp(gmask);
    for (int i = glane; i < 16; i += G) cnt[i] = (uint16_t)gcount[i];
    __syncwarp(gmask);
    *nzero_out = (int)cnt[0];
    *pad_out = 0;
    if ((int)cnt[0] == n) { *l_out = 0; *g_out = 0; return 3; }
    int j = 1;
    while (j <= 15 && cnt[j] == 0) j++;
    int g = 15;
    while (g > 0 && cnt[g] == 0) g--;
    int l = want_bits;
    if (l < j) l = j;
    if (l > g) l = g;
    *l_out = l; *g_out = g;
    int y = 1 << j;
    for (; j < g; j++, y <<= 1) {
        y -= (int)cnt[j];
        if (y < 0) return 1;
    }
    y -= (int)cnt[g];
    if (y < 0) return 1;
    *pad_out = y;
    return (y != 0 && g != 1) ? 2 : 0;
}

// sorted symbols + first codes (aux[16..31]) + end offsets (aux[0..15]), then the root LUT.
// KIND 0 = literal/length (R = RL), 1 = distance (R = RD).
template <int G, int KIND, int R>
__device__ __forceinline__ void make_lut(uint32_t* aux, const uint8_t* lens, int n, int nzero, const uint16_t* cnt, uint16_t* start,
                                         uint16_t* sorted, uint16_t* lut, int glane, unsigned gmask)
{
    if (glane == 0) {
        uint32_t off = 0, code = 0;
        for (int k = 1; k <= 15; k++) {
            aux[k] = off;
            aux[16 + k] = code;
            off += cnt[k];
            code = (code + cnt[k]) << 1;
        }
        for (int s = 0; s < n; s++) {
            uint32_t k = lens[s];
            if (k) { uint32_t o = aux[k]; sorted[o] = (uint16_t)(s | (k << 12)); aux[k] = o + 1; }
        }
    }
    uint32_t* lut32 = reinterpret_cast<uint32_t*>(lut);
    for (int i = glane; i < (1 << R) / 2; i += G) lut32[i] = E_INVALID | (E_INVALID << 16);
    __syncwarp(gmask);
    const int ncodes = n - nzero;
    if (glane == 0) {
        // canonical-walk state after R bits (canon_long)
        uint32_t first = 0, index = 0;
        for (int k = 1; k <= R && k <= 15; k++) { index += cnt[k]; first = (first + cnt[k]) << 1; }
        start[0] = (uint16_t)first; start[1] = (uint16_t)index;
    }
    for (int k = glane; k < ncodes; k += G) {
        uint32_t e = sorted[k];
        uint32_t sym = e & 0xfff, len = e >> 12;
        uint32_t idx = (uint32_t)k - (aux[len] - cnt[len]);             // aux[len] = one past the last of this length
        uint32_t code = aux[16 + len] + idx;
        uint32_t rev = __brev(code) >> (32 - len);
        if (len > (uint32_t)R) { lut[rev & ((1u << R) - 1u)] = (uint16_t)E_LONG; continue; }
        uint32_t entry;
        if (KIND == 0) {
            if (sym < 256) entry = sym;
            else if (sym == 256) entry = 0x100;
            else {
                uint32_t i = sym - 257;
                if (i > 28) continue;                                   // 286/287: invalid (fixed block only)
                uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                entry = 0x800 | (xb << 8) | (base - 3);
            }
        } else {
            if (sym > 29) continue;                                     // 30/31: invalid (fixed block only)
            uint32_t xb = sym < 4 ? 0 : (sym >> 1) - 1;
            uint32_t m = sym < 4 ? sym : 2 + (sym & 1);
            entry = (xb << 8) | m;
        }
        entry |= len << 12;
        for (uint32_t j = rev; j < (1u << R); j += (1u << len)) lut[j] = (uint16_t)entry;
    }
    __syncwarp(gmask);
}

// lit/len + distance tables for lens[0..nl) and lens[nl..nl+nd) with the reference's checks
// and messages (inflate_trees_dynamic, src/inftree.ts:333-379).  fixed: no checks.
template <int G, class SM>
__device__ __noinline__ TreeInfo build_tables(SM* S, uint16_t* gsorted, int nl, int nd, bool fixed, int glane, unsigned gmask,
                                              uint16_t* long_l, int long_n, bool spec = false)
{
    TreeInfo T;
    T.msg = SDZ_MSG_NONE; T.lbits = T.dbits = T.g_l = T.g_d = 0;
    int pad_l = 0, pad_d = 0, used = 0, nz_l = 0, nz_d = 0;
    const uint8_t* lens = reinterpret_cast<const uint8_t*>(gsorted + SORTED_L + SORTED_D);
    uint32_t* aux = reinterpret_cast<uint32_t*>(gsorted + SORTED_L + SORTED_D + 160);
    int st = classify<G>(lens, nl, 9, S->cnt_l, aux, &T.lbits, &T.g_l, &pad_l, &nz_l, glane, gmask);
    if (!fixed && spec) {
        // zlib 1.3 (inflate.c, state LENLENS -> CODELENS -> tables): the end-of-block code must exist; sets may be
        // incomplete only when their longest code is one bit; no distance code at all is fine until one is used; no arena
        // limit
        if (lens[256] == 0) { T.msg = SDZ_MSG_MISSING_EOB; return T; }
        if (st == 1 || st == 2) { T.msg = SDZ_MSG_BAD_LITLEN_SET; return T; }
        st = classify<G>(lens + nl, nd, 6, S->cnt_d, aux, &T.dbits, &T.g_d, &pad_d, &nz_d, glane, gmask);
        if (st == 1 || st == 2) { T.msg = SDZ_MSG_BAD_DIST_SET; return T; }
    } else
    if (!fixed) {
        // the lit/len and distance tables share an arena of MANY = 1400 entries; running out of
        // it is reported as DATA_ERROR, i.e. with the "oversubscribed" text
        if (st == 1) { T.msg = SDZ_MSG_OVERSUB_LITLEN_TREE; return T; }
        if (st != 3) used = ref_table_total<G>(S->cnt_l, T.g_l, pad_l, T.lbits, glane, gmask);
        if (used > 1400) { T.msg = SDZ_MSG_OVERSUB_LITLEN_TREE; return T; }
        if (st == 2 || st == 3) { T.msg = SDZ_MSG_INCOMPLETE_LITLEN_TREE; return T; }
    }
    if (!(spec && !fixed)) st = classify<G>(lens + nl, nd, fixed ? 5 : 6, S->cnt_d, aux, &T.dbits, &T.g_d, &pad_d, &nz_d, glane, gmask);
    if (!fixed && !spec) {
        if (st == 1) { T.msg = SDZ_MSG_OVERSUB_DIST_TREE; return T; }
        if (st != 3 && used + ref_table_total<G>(S->cnt_d, T.g_d, pad_d, T.dbits, glane, gmask) > 1400) {
            T.msg = SDZ_MSG_OVERSUB_DIST_TREE; return T;
        }
        if (st == 2) { T.msg = SDZ_MSG_INCOMPLETE_DIST_TREE; return T; }
        if (st == 3 && nl > 257) { T.msg = SDZ_MSG_EMPTY_DIST_TREE; return T; }
    }
    // (the 288-symbol counting sort is a serial read-modify-write chain on aux[]: it runs on the not yet built
    // distance LUT's shared memory instead of the global scratch - ~30 instead of ~500 cycles per symbol)
    make_lut<G, 0, RL>(reinterpret_cast<uint32_t*>(S->lut_d), lens, nl, nz_l, S->cnt_l, S->start, gsorted, S->lut_l, glane, gmask);
#if SDZ_LONG_SMEM > 0
    {
        const int i0 = (int)S->start[1], nc = nl - nz_l;
        for (int j = glane; j < long_n; j += G) long_l[j] = i0 + j < nc ? gsorted[i0 + j] : (uint16_t)0;
        __syncwarp(gmask);
    }
#endif
    make_lut<G, 1, RD>(aux, lens + nl, nd, nz_d, S->cnt_d, S->start + 2, gsorted + SORTED_L, S->lut_d, glane, gmask);
    if (fixed) { T.lbits = 9; T.dbits = 5; }
    return T;
}

// SDZ_PARITY_SPEC: zlib 1.3 words the tree errors differently (inflate.c: "invalid code lengths set" etc.)
__device__ __forceinline__ int spec_msg(int m)
{
    if (m == SDZ_MSG_OVERSUB_BITS_TREE || m == SDZ_MSG_INCOMPLETE_BITS_TREE) return SDZ_MSG_BAD_CODE_LENGTHS_SET;
    if (m == SDZ_MSG_OVERSUB_LITLEN_TREE || m == SDZ_MSG_INCOMPLETE_LITLEN_TREE) return SDZ_MSG_BAD_LITLEN_SET;
    if (m == SDZ_MSG_OVERSUB_DIST_TREE || m == SDZ_MSG_INCOMPLETE_DIST_TREE) return SDZ_MSG_BAD_DIST_SET;
    return m;
}

// code-length-code LUT for the dynamic header (inflate_trees_bits, src/inftree.ts:313-331).
// cl[19] are the code-length-code lengths; blut[128] receives sym | len << 5.
// Returns bb (index width, >= 1) or -msg on error.
__device__ __noinline__ int build_bits_lut(const uint8_t* cl, uint8_t* blut, uint16_t* cnt, int glane, unsigned gmask, bool spec = false)
{
    if (glane == 0) {
        for (int i = 0; i < 16; i++) cnt[i] = 0;
        for (int i = 0; i < 19; i++) cnt[cl[i]]++;
    }
    __syncwarp(gmask);
    if (cnt[0] == 19) return -SDZ_MSG_INCOMPLETE_BITS_TREE;
    int j = 1;
    while (cnt[j] == 0) j++;
    int g = 7;
    while (cnt[g] == 0) g--;
    int l = 7;
    if (l > g) l = g;
    int y = 1 << j;
    for (; j < g; j++, y <<= 1) { y -= (int)cnt[j]; if (y < 0) return -SDZ_MSG_OVERSUB_BITS_TREE; }
    y -= (int)cnt[g];
    if (y < 0) return -SDZ_MSG_OVERSUB_BITS_TREE;
    if (y != 0 && (g != 1 || spec)) return -SDZ_MSG_INCOMPLETE_BITS_TREE;      // zlib: the code-length code must be complete
    if (glane == 0) {
        uint32_t code = 0;
        for (int k = 1; k <= g; k++) {
            for (int s = 0; s < 19; s++) {
                if (cl[s] != k) continue;
                uint32_t rev = __brev(code) >> (32 - k);
                for (uint32_t q = rev; q < (1u << l); q += (1u << k)) blut[q] = (uint8_t)(s | (k << 5));
                code++;
            }
            code <<= 1;
        }
        if (g == 1 && cnt[1] == 1) blut[1] = blut[0];      // a lone 1-bit code answers both patterns (Q11)
    }
    __syncwarp(gmask);
    return l;
}

// source starts before the output: preset dictionary tail, else the reference's
// zero-initialised window (src/infcodes.ts:174-193)
template <int G>
__device__ __noinline__ void copy_before_start_impl(uint8_t* o, uint32_t p0, uint32_t len, uint32_t dist,
                                                    const uint8_t* dt, int Dn, int gl)
{
    for (uint32_t i = gl; i < len; i += G) {
        uint32_t k = dist >= len ? i : i % dist;
        int64_t s = (int64_t)p0 - (int64_t)dist + (int64_t)k;
        uint8_t v = 0;
        if (s >= 0) v = o[s];
        else if (s >= -(int64_t)Dn) v = dt[(int64_t)Dn + s];
        o[p0 + i] = v;
    }
}

enum : int { PH_FETCH = 0, PH_BLOCK = 1, PH_CODES = 2, PH_EXIT = 3 };

template <int G, bool STORE, int TM = TM_NONE>
struct Decoder {
    static constexpr bool MARK = TM == TM_MARK;
    GroupSmem* S;
    uint16_t* gsorted;                 // global scratch: sorted symbols of the current block
    unsigned gmask;
    int glane;

    // ---- bit reader
    uint64_t bb;
    int bc;
    uint32_t nw;                       // prefetched word `wp`
    uint32_t wp, end_wp;
    uint32_t in_len;
    const uint8_t* gsrc;
    uint32_t chunk0, total_chunks, issued_abs, waited_abs, phasebits;

    // ---- output
    uint8_t* out;
    uint16_t* out16;                   // marker mode
    uint64_t abs_start;                // marker mode: absolute stream offset of the piece being decoded
    uint64_t resume_bit;               // TM_MARK
    uint32_t limit;                    // TM_MARK: stop after this many symbols
    uint32_t next_ck;                  // TM_INDEX: output position of the next resume point
    uint32_t pos, cap;
    const uint8_t* dict_tail;
    int D;

    // deferred match copy: bytes loaded for the previous short match, stored when the next match
    // arrives, so that the L2 round trip of a copy overlaps the decode of the following symbols
    // two matches can be pending: `o_` the older one (its copies are complete after wait_group 1),
    // `n_` the newer one.  len | soff << 8 is packed in *_meta (0 = empty); the older one's words
    // live in staging slot `ptog`, the newer one's in slot `ptog ^ 1`.
    uint32_t o_dst, o_meta, n_dst, n_meta, ptog;
    __device__ __forceinline__ bool pending_any() const
    {
        return (o_meta | n_meta) != 0u;
    }
    static constexpr uint32_t NSLOT = 2;               // staging slots = deferred matches in flight per stream
    static __device__ __forceinline__ uint32_t slot_after(uint32_t p) { return p ^ 1u; }

    RingModel ring;
    // Model of the reference's input frontier near the end of the input (src/infcodes.ts:339, :96-100, :287-291): whether
    // a symbol - in particular an end-of-block code - is decoded by inflate_fast() or by the one-symbol slow path depends
    // on how many input BYTES the reference has already loaded into its bit buffer (n >= 10 at the START check / at the
    // end of every fast symbol), and the two paths end a block differently (block_end()).  ref_F = bytes loaded so far.
    uint32_t ref_F, ref_Fentry, ring_done;
    uint64_t blk_sym0_bit;             // bit position of the first symbol of the current block
    bool ref_on, ref_burst, ref_forced_slow, eob_emu, eob_fast;
    int msg;
    int stall_kind;
    int lbits, dbits, g_l, g_d;
    int eob_len;                       // code length of the end-of-block symbol just decoded

    // ---- stream / block bookkeeping (phase machine)
    int phase;
    unsigned long long idx;
    uint64_t out_off;
    uint32_t start_pos, n_blocks, name_off, name_len;
    int32_t mtime;
    int last, method;
    bool raw, is_gzip;
    // streaming sessions: bit position of the current block's header and of the symbol being decoded; the first symbol
    // after a resume point continues in the reference's one-symbol path whatever the input (it re-enters InfCodes.proc
    // past its START check, src/infcodes.ts:339-357); ref_floor = input bytes the reference had loaded when it stopped
    uint64_t blk_hdr_bit, last_sym_bit;
    uint32_t ref_floor;
    bool resume_first, hdr_counted;
    bool spec;                         // SDZ_PARITY_SPEC (include/sdzcuda.h)

    // ------------------------------------------------------------------ input staging
    __device__ __forceinline__ uint32_t load_word(uint32_t w)
    {
        const uint32_t c = w / CHW;
        if (c >= waited_abs) {
            uint64_t r = chunk_cross(S, gsrc, c, chunk0, issued_abs, total_chunks, phasebits, gmask, glane);
            phasebits = (uint32_t)r & 15u;
            issued_abs = (uint32_t)(r >> 4);
            waited_abs = c + 1;
        }
        return S->ring[w % (NBUF * CHW)];
    }

    __device__ __forceinline__ void drain()
    {
        for (uint32_t c = waited_abs; c < issued_abs; c++) {
            const uint32_t slot = c % NBUF;
            mbar_wait(&S->mbar[slot], (phasebits >> slot) & 1u);
            phasebits ^= 1u << slot;
        }
        waited_abs = issued_abs;
    }

    // (re)position the reader on a byte boundary of the stream
    __device__ __forceinline__ void seek(uint32_t byte_pos)
    {
        drain();
        __syncwarp(gmask);
        uint32_t w = byte_pos >> 2;
        chunk0 = w / CHW;
        uint32_t n0 = total_chunks > chunk0 ? total_chunks - chunk0 : 0;
        if (n0 > (uint32_t)NBUF) n0 = NBUF;
        if (glane == 0) {
            for (uint32_t i = 0; i < n0; i++) {
                const uint32_t slot = (chunk0 + i) % NBUF;
                mbar_arrive_expect_tx(&S->mbar[slot], CH);
                bulk_copy_g2s(&S->ring[slot * CHW], gsrc + (size_t)(chunk0 + i) * CH, CH, &S->mbar[slot]);
            }
        }
        waited_abs = chunk0;
        issued_abs = chunk0 + n0;
        wp = w; bb = 0; bc = 0;
        nw = wp < end_wp ? load_word(wp) : 0u;
        refill();
        uint32_t skip = (byte_pos & 3) * 8;
        if (skip) { uint32_t k = min(skip, (uint32_t)bc); bb >>= k; bc -= (int)k; }
        refill();
    }

    // Bits past the end of the stream that ride along in the last word are never counted in
    // `bc`; every consumer either checks `bc` / avail_bits() or is far from the end.
    __device__ __forceinline__ void refill()
    {
        if (bc <= 32 && wp < end_wp) {
            bb |= (uint64_t)nw << bc;
            const uint32_t rem = in_len - wp * 4u;
            bc += rem >= 4u ? 32 : (int)(rem * 8u);
            wp++;
            nw = wp < end_wp ? load_word(wp) : 0u;
        }
    }

    // total bits between the read position and the end of the input (capped)
    __device__ __forceinline__ int avail_bits() const
    {
        uint64_t unloaded = wp < end_wp ? (uint64_t)in_len * 8 - (uint64_t)wp * 32 : 0;
        uint64_t a = (uint64_t)bc + unloaded;
        return a > 4096 ? 4096 : (int)a;
    }
    __device__ __forceinline__ uint64_t bit_pos() const
    {
        uint64_t loaded = wp < end_wp ? (uint64_t)wp * 32 : (uint64_t)in_len * 8;
        return loaded - (uint64_t)bc;
    }
    __device__ __forceinline__ uint32_t byte_pos() const { return (uint32_t)(bit_pos() >> 3); }
    __device__ __forceinline__ bool ensure(int n) { refill(); return bc >= n; }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; bc -= n; }

    // ------------------------------------------------------------------ output
    // Deferred match copy.  A short non-overlapping match (<= 16 bytes, the common case) only
    // ISSUES asynchronous copies (cp.async / LDGSTS): lane j stages the aligned words that cover
    // its DB = 16 / G source bytes [DB j, DB j + DB) into its private slice of stage[].  The bytes are moved to
    // their destination when the group's next match (or the end of the stream) arrives, so the
    // L2 / DRAM round trip of the window read overlaps the decode of the following symbols instead
    // of stalling the lockstep warp.  Each lane reads back only what it staged itself: no
    // cross-lane synchronisation is needed at commit time.
    __device__ __forceinline__ void store_lit(uint32_t v)
    {
        if (STORE) {
            if (glane == 0) { if (MARK) out16[pos] = (uint16_t)v; else out[pos] = (uint8_t)v; }
        }
    }

    // marker mode: the block is decoded without its 32 KiB window.  A source position before the block
    // start becomes the symbol 256 + (index into the window that ends at the block start); symbols are
    // copied like bytes otherwise (they may themselves be markers).  Before the start of the whole
    // stream the reference's window holds zeros.
    __device__ __forceinline__ void copy_match_marked(uint32_t len, uint32_t dist)
    {
        __syncwarp(gmask);
        uint16_t* dst = out16 + pos;
        for (uint32_t i = glane; i < len; i += G) {
            const uint32_t k = dist >= len ? i : i % dist;
            const int64_t s = (int64_t)pos + (int64_t)k - (int64_t)dist;          // relative to the block start
            uint16_t v;
            if (s >= 0) v = out16[s];
            else if ((int64_t)abs_start + s < 0) v = 0;
            else v = (uint16_t)(256 + 32768 + s);
            dst[i] = v;
        }
    }

    // bytes of a deferred match each lane moves, and the aligned words it stages for them
    // marker mode moves two-byte symbols: a lane takes 8 bytes (4 symbols), so that copies of up to 16 symbols are deferred
    static constexpr int WIDE = (MARK && G == 4 && SDZ_MARK_WIDE) ? 2 : 1;
    static constexpr int DB = (G <= MAX_G_DEFERRED ? 16 / G : 4) * WIDE;
    static constexpr int DW = DB / 4 + 1;
    static constexpr int SLOTW = (G <= MAX_G_DEFERRED ? G : 4) * DW;            // words per staging slot
    static constexpr uint32_t DEFER_MAX = 16u * WIDE;                            // bytes a deferred copy can move
    static constexpr int LONG_OFF = (int)NSLOT * SLOTW;
    static constexpr int LONG_N = SDZ_LONG_SMEM > 0 ? (STAGE_LONG_WORDS - LONG_OFF) * 2 : 0;
    static_assert(LONG_OFF <= STAGE_LONG_WORDS, "staging slots do not fit");
    __device__ __forceinline__ uint16_t* long_l() const { return reinterpret_cast<uint16_t*>(S->stage_long + LONG_OFF); }

    // move one pending match from its staging slot to its destination (straight-line, predicated:
    // nothing happens when meta == 0)
    __device__ __forceinline__ void commit_slot(uint32_t dst_off, uint32_t meta, uint32_t slot)
    {
        const uint32_t plen = meta & 0xffu, soff = meta >> 8;
        const uint32_t jb = (uint32_t)DB * (uint32_t)glane;
        const uint32_t* st = &S->stage_long[SLOTW * slot + DW * glane];
        uint8_t* dst = out + dst_off + jb;
        const uint32_t nb = plen > jb ? plen - jb : 0u;
        uint32_t w[DW];
        #pragma unroll
        for (int k = 0; k < DW; k++) w[k] = st[k];
        #pragma unroll
        for (int q = 0; q < DB / 4; q++) {
            const uint32_t v = __funnelshift_r(w[q], w[q + 1], soff * 8u);
            if (MARK) {                                  // two-byte symbols: everything is 2-byte aligned
                st_u16_if(reinterpret_cast<uint16_t*>(dst + 4 * q), v & 0xffffu, nb > 4u * q);
                st_u16_if(reinterpret_cast<uint16_t*>(dst + 4 * q + 2), v >> 16, nb > 4u * q + 2);
            } else {
                st_u8_if(dst + 4 * q, v, nb > 4u * q);
                st_u8_if(dst + 4 * q + 1, v >> 8, nb > 4u * q + 1);
                st_u8_if(dst + 4 * q + 2, v >> 16, nb > 4u * q + 2);
                st_u8_if(dst + 4 * q + 3, v >> 24, nb > 4u * q + 3);
            }
        }
    }

    // complete both pending matches (end of stream, or a copy that may read their bytes)
    __device__ __forceinline__ void flush_pending()
    {
        if (STORE && G <= MAX_G_DEFERRED) {
            cp_async_wait_all();
            commit_slot(o_dst, o_meta, ptog);
            commit_slot(n_dst, n_meta, ptog ^ 1u);
        }
        o_meta = 0; n_meta = 0;
    }

    // Memory ordering: the lockstep loop executes a full-mask __syncwarp() at the top of every
    // iteration, so stores of earlier iterations (literals, committed matches) are ordered before
    // the reads issued here.  Only bytes committed in THIS iteration need an extra group sync.
    // `lit_now`: number of literals lane 0 stored earlier in THIS lockstep iteration (at pos - lit_now .. pos - 1).
    __device__ __forceinline__ int copy_match(uint32_t len, uint32_t dist, uint32_t lit_now)
    {
        if constexpr (TM == TM_NONE) {
            if (spec && dist > pos + (uint32_t)D) { msg = SDZ_MSG_DIST_TOO_FAR; return R_ERROR; }      // (the reference copies zeros)
        }
        if (len > cap - pos) return R_OUTFULL;
        if (MARK && dist > pos) {
            // reaches before the piece: those symbols become markers (synchronous; group sync inside)
            if (pending_any()) flush_pending();
            copy_match_marked(len, dist);
        } else if (STORE) {
            // Marker mode: a symbol is two bytes, and a copy that stays inside the piece is an ordinary copy of
            // 2 len bytes at distance 2 dist in the symbol buffer - everything below runs in BYTE units.
            constexpr uint32_t E = MARK ? 2u : 1u;
            const uint32_t bpos = pos * E, blen = len * E, bdist = dist * E;
            const bool simple = G <= MAX_G_DEFERRED && bdist >= blen && blen <= DEFER_MAX && bdist <= bpos;
            // the new source must not overlap bytes that are still pending (the older pending match has
            // the lower destination); copies on the synchronous path read arbitrary earlier bytes
            const uint32_t first_pending = o_meta ? o_dst : n_dst;
            const bool any_pending = (o_meta | n_meta) != 0;
            // (a copy on the synchronous path whose source lies entirely below the pending destinations can
            // overtake them: its own destination is disjoint from theirs)
            const bool hazard = any_pending && (bdist > bpos || bpos - bdist + blen > first_pending);
            if (hazard) flush_pending();
            // bytes written in this iteration (just-committed matches, the folded literals at pos - lit_now .. pos - 1)
            // are only ordered before the reads below by a group sync; a deferred match (dist >= len) reads one of
            // the folded literals iff dist < len + lit_now
            if (hazard || !simple || bdist < blen + lit_now * E) __syncwarp(gmask);
            uint8_t* dst = out + bpos;
#if SDZ_TWOSLOT
            const bool two = G <= MAX_G_DEFERRED && !simple && bdist >= blen && blen <= 2u * DEFER_MAX && bdist <= bpos;
#else
            const bool two = false;
#endif
            if (two) {
                // 17..32 bytes, not overlapping: both staging slots are emptied and take one half of the match each,
                // so this copy has no load-to-store round trip either (the warp's other groups are waiting here)
                if (any_pending) { flush_pending(); __syncwarp(gmask); }
                const uint8_t* src = dst - bdist;
                const uint32_t so = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
                const uint32_t jb = (uint32_t)DB * (uint32_t)glane;
                const uint8_t* w0 = src - so + jb;
                #pragma unroll
                for (int k = 0; k < DW; k++) cp_async4_if(&S->stage_long[SLOTW * ptog + DW * glane + k], w0 + 4 * k, jb + 4u * k < DEFER_MAX + so);
                cp_async_commit();
                #pragma unroll
                for (int k = 0; k < DW; k++) cp_async4_if(&S->stage_long[SLOTW * slot_after(ptog) + DW * glane + k], w0 + DEFER_MAX + 4 * k, jb + 4u * k < blen - DEFER_MAX + so);
                cp_async_commit();
                o_dst = bpos; o_meta = DEFER_MAX | (so << 8);
                n_dst = bpos + DEFER_MAX; n_meta = (blen - DEFER_MAX) | (so << 8);
            } else if (simple) {
                // the older pending match was issued two matches ago: wait for it (only), store it, and
                // reuse its staging slot for this match
                cp_async_wait_but_one();
                commit_slot(o_dst, o_meta, ptog);
                const uint8_t* src = dst - bdist;
                const uint32_t so = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
                const uint32_t jb = (uint32_t)DB * (uint32_t)glane;
                const uint8_t* w0 = src - so + jb;
                #pragma unroll
                for (int k = 0; k < DW; k++) cp_async4_if(&S->stage_long[SLOTW * ptog + DW * glane + k], w0 + 4 * k, jb + 4u * k < blen + so);
                cp_async_commit();
                o_dst = n_dst; o_meta = n_meta;
                n_dst = bpos; n_meta = blen | (so << 8);
                ptog = slot_after(ptog);
            } else if (bdist <= bpos) {
                const uint8_t* src = dst - bdist;
                if (bdist >= blen) {
                    // eight loads in flight per lane: one memory round trip for matches up to 8 G bytes
                    for (uint32_t i = glane; i < blen; i += 8 * G) {
                        uint8_t v[8];
                        #pragma unroll
                        for (int k = 0; k < 8; k++) v[k] = i + k * G < blen ? src[i + k * G] : (uint8_t)0;
                        #pragma unroll
                        for (int k = 0; k < 8; k++) if (i + k * G < blen) dst[i + k * G] = v[k];
                    }
                } else if (bdist == 1) {
                    uint8_t v = src[0];
                    for (uint32_t i = glane; i < blen; i += G) dst[i] = v;
                } else {                             // lane-strided replicate of the period
                    for (uint32_t i = glane; i < blen; i += G) dst[i] = src[i % bdist];
                }
            } else {
                copy_before_start_impl<G>(out, pos, len, dist, dict_tail, D, glane);     // byte mode only (bdist > bpos)
            }
        }
        pos += len;
        return R_OK;
    }

    // dynamic block header: HLIT/HDIST/HCLEN, code-length code, RLE-coded lengths
    // (src/infblocks.ts:334-523).  Uses lut_l as scratch for the 7-bit code-length LUT.
    __device__ __forceinline__ int dynamic_header(int* nl_out, int* nd_out)
    {
        if (!ensure(14)) { stall_kind = ST_OTHER; return R_STALL; }
        if (TM == TM_NONE) ref_need(bit_pos(), 14u);
        uint32_t t = peek(14);
        if ((t & 0x1f) > 29 || ((t >> 5) & 0x1f) > 29) { msg = SDZ_MSG_TOO_MANY_SYMS; return R_ERROR; }
        drop(14);
        int nl = 257 + (int)(t & 0x1f), nd = 1 + (int)((t >> 5) & 0x1f), ncl = 4 + (int)(t >> 10);
        int total = nl + nd;
        uint8_t* lens = reinterpret_cast<uint8_t*>(gsorted + SORTED_L + SORTED_D);
        uint8_t* cl = reinterpret_cast<uint8_t*>(S->lut_d);   // 19 code-length-code lengths (scratch)
        __syncwarp(gmask);
        for (int i = glane; i < 19; i += G) cl[i] = 0;
        __syncwarp(gmask);
        for (int i = 0; i < ncl; i++) {
            if (!ensure(3)) { stall_kind = ST_DYNHDR; return R_STALL; }
            if (TM == TM_NONE) ref_need(bit_pos(), 3u);
            if (glane == 0) cl[c_border[i]] = (uint8_t)peek(3);
            drop(3);
        }
        __syncwarp(gmask);
        uint8_t* blut = reinterpret_cast<uint8_t*>(S->lut_l);
        int bb_bits = build_bits_lut(cl, blut, S->cnt_l, glane, gmask, spec);
        if (bb_bits < 0) { msg = spec ? spec_msg(-bb_bits) : -bb_bits; return R_ERROR; }
        int index = 0;
        uint32_t prev = 0;
        while (index < total) {
            if (!ensure(bb_bits)) { stall_kind = ST_DYNHDR; return R_STALL; }
            if (TM == TM_NONE) ref_need(bit_pos(), (uint32_t)bb_bits);
            uint32_t e = blut[peek(bb_bits)];
            int tbits = (int)(e >> 5), c = (int)(e & 31);
            if (c < 16) {
                drop(tbits);
                if (glane == 0) lens[index] = (uint8_t)c;
                prev = (uint32_t)c;
                index++;
            } else {
                int i = c == 18 ? 7 : c - 14;
                int j = c == 18 ? 11 : 3;
                if (bc < tbits + i) { stall_kind = ST_DYNHDR; return R_STALL; }     // ensure() above left >= 33 bits unless the input ends
                if (TM == TM_NONE) ref_need(bit_pos(), (uint32_t)(tbits + i));
                drop(tbits);
                j += (int)peek(i);
                drop(i);
                if (index + j > total || (c == 16 && index < 1)) { msg = SDZ_MSG_BAD_REPEAT; return R_ERROR; }
                uint8_t v = c == 16 ? (uint8_t)prev : (uint8_t)0;
                prev = v;
                if (glane == 0) for (int q = 0; q < j; q++) lens[index + q] = v;
                index += j;
            }
        }
        __syncwarp(gmask);
        *nl_out = nl; *nd_out = nd;
        return R_OK;
    }

    // ------------------------------------------------------------------ one symbol
    // `tail` (fewer than five input words left) and root entries marked long/invalid go through
    // slow_lookup(), which also enforces the reference's lookahead rule; everything else is
    // one shared-memory LUT read per code.
    static __device__ __forceinline__ uint32_t ceil8(uint64_t bits) { return (uint32_t)((bits + 7) >> 3); }
    // the reference loads input until `need` bits from bit position p are in its buffer
    __device__ __forceinline__ void ref_need(uint64_t p, uint32_t need) { ref_F = max(ref_F, ceil8(p + need)); }
    // a symbol starts at bit p0 within the last 16 bytes of the input: does inflate_fast() decode it?
    __device__ __noinline__ bool ref_symbol_begin(uint64_t p0)
    {
        if (!ref_on) {
            ring.write(pos - ring_done); ring_done = pos;
            const bool fresh = p0 == blk_sym0_bit;                       // first symbol of its block: the START check is pending
            ref_burst = !fresh && ring.room() >= 258;                   // (n >= 10 held up to here)
            if (!fresh) ref_F = max(ceil8(p0), ref_floor);
            ref_Fentry = 0; ref_forced_slow = false; ref_on = true;
        }
        bool fast = ref_burst;
        if (!fast && !ref_forced_slow && ring.room() >= 258 && in_len >= ref_F + 10u) { fast = true; ref_burst = true; ref_Fentry = ref_F; }
        ref_forced_slow = false;
        return fast;
    }
    // whole unused bytes go back to the input when inflate_fast() returns (src/infcodes.ts:287-291)
    __device__ __forceinline__ void ref_give_back(uint64_t p_after)
    {
        const uint32_t k = (uint32_t)((uint64_t)ref_F * 8 - p_after);
        ref_F -= min(k >> 3, ref_F - ref_Fentry);
        ref_burst = false;
    }
    __device__ __noinline__ void ref_symbol_end(bool fast, uint32_t n_out)
    {
        ring.write(n_out); ring_done = pos;
        if (fast && !(ring.room() >= 258 && in_len >= ref_F + 10u)) { ref_give_back(bit_pos()); ref_forced_slow = true; }
    }

    __device__ __forceinline__ int step_general(uint32_t lit_now)
    {
        refill();
        const bool tail = wp + 5 > end_wp;
        const uint64_t p0 = bit_pos();
        const bool emu = TM == TM_NONE && (uint64_t)in_len - (p0 >> 3) <= 16u;
        bool fastsym = emu ? ref_symbol_begin(p0) : false;
        bool rf = false;
        if constexpr (TM == TM_NONE) {
            last_sym_bit = p0;
            rf = resume_first;
            resume_first = false;
            if (rf) { fastsym = false; ref_burst = false; }
        }
        uint32_t need1 = 20u;
        uint32_t e = S->lut_l[(uint32_t)bb & ((1u << RL) - 1u)];
        uint32_t n = e >> 12, p = e & 0xfff;
        if (tail || n == 0) {
            uint32_t r = slow_lookup(S->cnt_l, gsorted, lbits, g_l, (uint32_t)bb, avail_bits(), spec);
            uint32_t st = r >> 28;
            if (st) { if (st == (uint32_t)R_ERROR) msg = SDZ_MSG_BAD_LITLEN_CODE; return (int)st; }
            n = (r >> 16) & 0xff;
            need1 = (r >> 24) & 15u;
            uint32_t sym = r & 0xffff;
            if (sym <= 256) p = sym;
            else {
                uint32_t i = sym - 257;
                if (i > 28) { msg = SDZ_MSG_BAD_LITLEN_CODE; return R_ERROR; }
                uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                p = 0x800 | (xb << 8) | (base - 3);
            }
        }
        if (emu) ref_need(p0, fastsym ? 20u : need1);
        bb >>= n; bc -= (int)n;
        if (p < 256) {
            if (pos >= cap) return R_OUTFULL;
            store_lit(p);
            pos++;
            if (emu) ref_symbol_end(fastsym, 1u);
            return R_OK;
        }
        if (p == 256) {
            eob_len = (int)n; eob_emu = emu || rf; eob_fast = fastsym;
            if (emu) {
                // the block's codes object hands whole unused bytes back: all of them after inflate_fast(), at most one in
                // WASH (src/infcodes.ts:620-624)
                if (fastsym) ref_give_back(bit_pos());
                else if ((uint64_t)ref_F * 8 - bit_pos() > 7) ref_F--;
                ref_burst = false; ref_forced_slow = false;
            }
            return R_EOB;
        }
        uint32_t xb = (p >> 8) & 7;
        if (tail && avail_bits() < (int)xb) return R_STALL;
        if (emu && !fastsym) ref_need(p0 + n, xb);
        uint32_t len = 3 + (p & 0xff) + ((uint32_t)bb & ((1u << xb) - 1u));
        bb >>= xb; bc -= (int)xb;
        refill();
        uint32_t need2 = 15u;
        uint32_t de = S->lut_d[(uint32_t)bb & ((1u << RD) - 1u)];
        uint32_t dn = de >> 12;
        if (tail || dn == 0) {
            if (g_d == 0) { msg = SDZ_MSG_BAD_DIST_CODE; return R_ERROR; }
            uint32_t r = slow_lookup(S->cnt_d, gsorted + SORTED_L, dbits, g_d, (uint32_t)bb, avail_bits(), spec);
            uint32_t st = r >> 28;
            if (st) { if (st == (uint32_t)R_ERROR) msg = SDZ_MSG_BAD_DIST_CODE; return (int)st; }
            dn = (r >> 16) & 0xff;
            need2 = (r >> 24) & 15u;
            uint32_t ds = r & 0xffff;
            if (ds > 29) { msg = SDZ_MSG_BAD_DIST_CODE; return R_ERROR; }
            de = ((ds < 4 ? 0u : (ds >> 1) - 1u) << 8) | (ds < 4 ? ds : 2u + (ds & 1u));
        }
        if (emu) ref_need(p0 + n + xb, fastsym ? 15u : need2);
        bb >>= dn; bc -= (int)dn;
        uint32_t dx = (de >> 8) & 15;
        if (tail && avail_bits() < (int)dx) return R_STALL;
        if (emu) ref_need(p0 + n + xb + dn, dx);
        uint32_t dist = 1 + ((de & 3) << dx) + ((uint32_t)bb & ((1u << dx) - 1u));
        bb >>= dx; bc -= (int)dx;
        const int rc = copy_match(len, dist, lit_now);
        if (emu && rc == R_OK) ref_symbol_end(fastsym, len);
        return rc;
    }

    // branch-free top-up used by the fast path (at least five whole input words remain):
    // predicated instructions only, except for the rare hop into the next 128-byte chunk
    __device__ __forceinline__ void refill_fast()
    {
        const bool take = bc <= 32;
        bb |= take ? ((uint64_t)nw << bc) : 0ull;
        bc += take ? 32 : 0;
        wp += take ? 1u : 0u;
        // chunk(wp) had been waited for, so wp < waited_abs * CHW held before the increment
        if (take && wp == waited_abs * CHW) {                             // rare: first word of the next 128-byte chunk
            uint64_t r = chunk_cross(S, gsrc, wp / CHW, chunk0, issued_abs, total_chunks, phasebits, gmask, glane);
            phasebits = (uint32_t)r & 15u;
            issued_abs = (uint32_t)(r >> 4);
            waited_abs = wp / CHW + 1;
        }
        const uint32_t w = S->ring[wp % (NBUF * CHW)];                    // always readable: chunk(wp) has been waited for
        nw = take ? w : nw;
    }

    // One lockstep iteration as straight-line code.  The warp executes the match path in practically every
    // iteration (one of its eight groups almost always has a match), so literal groups run it too, with
    // every side effect predicated off: no divergence between literal and match groups, no reconvergence
    // points, one branch for everything rare (long / invalid code, end of block), one for copies that are
    // not plain deferred copies.  The stream tail and the last 260 bytes of the output slot go through
    // step_general(), which has all the checks.
    __device__ __forceinline__ int step_flat()
    {
        constexpr uint32_t LMASK = (1u << RL) - 1u, DMASK = (1u << RD) - 1u;
        // (seven words: every symbol that starts within the last 16 bytes of the input goes through step_general(), which
        //  models the reference's input frontier there; one call site for step_general(): it is inlined once)
        bool general = wp + 7 > end_wp || cap - pos < (uint32_t)SDZ_CAPMARGIN;
        if constexpr (TM == TM_NONE) general = general || resume_first;
        bool fold = false;
        uint32_t nfold = 0;                                // literals stored by this iteration's folds
        uint32_t e = 0;
        if (!general) {
        refill_fast();
        e = S->lut_l[(uint32_t)bb & LMASK];
        // leading plain literal (marker mode: never the last symbol of the piece)
        fold = SDZ_LIT_RUN > 0 && e >= 0x1000u && (e & 0xf00u) == 0u && (!MARK || pos + 1u < limit);
        {
            const uint32_t n0 = fold ? e >> 12 : 0u;
            if (STORE) { if (MARK) st_u16_if(out16 + pos, e & 0xffu, fold && glane == 0); else st_u8_if(out + pos, e & 0xffu, fold && glane == 0); }
            pos += fold ? 1u : 0u;
            nfold = fold ? 1u : 0u;
            bb >>= n0; bc -= (int)n0;
            e = S->lut_l[(uint32_t)bb & LMASK];          // the same entry again when nothing was folded
        }
#if SDZ_LIT_RUN >= 2
        if constexpr (TM == TM_NONE) {
            // a second leading literal, when the bit buffer still covers it and the longest code + extra bits (9 + 20).
            // Batch kernel only: the block-task passes of the large-stream path are bound by the latency of one task, where
            // the longer dependent chain costs more than the saved iterations (1 GiB stream: 35.4 ms without, 35.9 ms with)
            const bool fold2 = fold && bc >= 29 && e >= 0x1000u && (e & 0xf00u) == 0u && (!MARK || pos + 1u < limit);
            const uint32_t n0 = fold2 ? e >> 12 : 0u;
            if (STORE) { if (MARK) st_u16_if(out16 + pos, e & 0xffu, fold2 && glane == 0); else st_u8_if(out + pos, e & 0xffu, fold2 && glane == 0); }
            pos += fold2 ? 1u : 0u;
            nfold += fold2 ? 1u : 0u;
            bb >>= n0; bc -= (int)n0;
            e = S->lut_l[(uint32_t)bb & LMASK];
        }
#endif
        if ((e >> 12) == 0u || (e & 0xfffu) == 0x100u) {    // rare: code longer than the root, invalid code, end of block
            bool ok = false;
            if (e == E_LONG) {
#if SDZ_LONG_SMEM > 0
                const uint32_t r = canon_long(S->cnt_l, gsorted, RL, g_l, S->start, (uint32_t)bb, long_l(), LONG_N);
#else
                const uint32_t r = canon_long(S->cnt_l, gsorted, RL, g_l, S->start, (uint32_t)bb);
#endif
                const uint32_t sym = r & 0xffffu;
                if (r != 0u && sym <= 256u) { e = ((r >> 16) << 12) | sym; ok = true; }
                else if (r != 0u && sym - 257u <= 28u) {
                    const uint32_t i = sym - 257u;
                    const uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                    const uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                    e = ((r >> 16) << 12) | 0x800u | (xb << 8) | (base - 3u);
                    ok = true;
                }
            } else if ((e >> 12) != 0u) ok = true;          // end of block with a root-table code
            if (!ok) general = true;
            else if ((e & 0xfffu) == 0x100u) { const uint32_t n = e >> 12; bb >>= n; bc -= (int)n; eob_len = (int)n; eob_emu = false; return R_EOB; }
        }
        }
        if (general) return step_general(nfold);
        const uint32_t n = e >> 12, p = e & 0xfffu;
        const bool ismatch = p >= 256u;
        if (STORE) { if (MARK) st_u16_if(out16 + pos, p, !ismatch && glane == 0); else st_u8_if(out + pos, p, !ismatch && glane == 0); }
        // a literal is a "match" of length 1 without extra bits: one consume and one `pos +=` serve both
        const uint32_t xb = ismatch ? (p >> 8) & 7u : 0u;
        const uint32_t len = ismatch ? 3u + (p & 0xffu) + (((uint32_t)bb >> n) & ((1u << xb) - 1u)) : 1u;
        bb >>= (n + xb); bc -= (int)(n + xb);
        refill_fast();                                     // harmless for literal groups: at least five whole words remain
        uint32_t de = S->lut_d[(uint32_t)bb & DMASK];
        uint32_t dn = de >> 12;
        if (ismatch && dn == 0u) {                         // rare: distance code longer than the root, or invalid
            const uint32_t r = de == E_LONG ? canon_long(S->cnt_d, gsorted + SORTED_L, RD, g_d, S->start + 2, (uint32_t)bb) : 0u;
            const uint32_t ds = r & 0xffffu;
            if (r == 0u || ds > 29u) { msg = SDZ_MSG_BAD_DIST_CODE; return R_ERROR; }     // far from the tail: no stall possible
            dn = r >> 16;
            de = ((ds < 4 ? 0u : (ds >> 1) - 1u) << 8) | (ds < 4 ? ds : 2u + (ds & 1u));
        }
        const uint32_t dx = (de >> 8) & 15u;
        const uint32_t dist = 1u + ((de & 3u) << dx) + (((uint32_t)bb >> dn) & ((1u << dx) - 1u));
        const uint32_t dcons = ismatch ? dn + dx : 0u;
        bb >>= dcons; bc -= (int)dcons;
        if (!STORE) {
            if constexpr (TM == TM_NONE) {
                if (spec && ismatch && dist > pos + (uint32_t)D) { msg = SDZ_MSG_DIST_TOO_FAR; return R_ERROR; }
            }
            pos += len; return R_OK;
        }
        // ---- copy (byte units; a marker symbol is two bytes)
        constexpr uint32_t E = MARK ? 2u : 1u;
        const uint32_t bpos = pos * E, blen = len * E, bdist = dist * E;
        bool cm = ismatch;                                 // a copy that reads earlier output
        if (MARK && G <= MAX_G_DEFERRED) {
            // marker mode: a short match whose source lies entirely before the piece reads nothing - its symbols are
            // markers (256 + index into the 32 KiB window before the piece; zeros before the start of the stream, Q6)
            // computed and stored here; pending copies are not disturbed (early in a piece this is the common match)
            constexpr uint32_t SPL = (uint32_t)DB / 2u;                         // symbols per lane
            const bool pre = ismatch && dist >= pos + len && len <= (uint32_t)G * SPL;
            #pragma unroll
            for (uint32_t q = 0; q < SPL; q++) {
                const uint32_t i = SPL * (uint32_t)glane + q;
                const uint32_t before = dist - pos - i;                         // 1 .. 32768 source symbols before the piece
                const uint32_t v = (uint64_t)before > abs_start ? 0u : 256u + 32768u - before;
                st_u16_if(out16 + pos + i, v, pre && i < len);
            }
            cm = ismatch && !pre;
        }
        // (room for a deferred match, <= 16 bytes, is guaranteed by the margin test at the top when SDZ_CAPMARGIN >= 18)
        // (a deferred copy issues its cp.async without a group sync, so its source must not contain a literal folded in
        //  this iteration: those sit at pos - nfold .. pos - 1, i.e. inside the source iff dist < len + nfold)
        const bool simple = G <= MAX_G_DEFERRED && cm && bdist >= blen + nfold * E && blen <= DEFER_MAX && bdist <= bpos &&
                            (SDZ_CAPMARGIN >= 18 || len <= cap - pos);
        const uint32_t first_pending = o_meta ? o_dst : n_dst;
        const bool hazard = (o_meta | n_meta) != 0u && (bdist > bpos || bpos - bdist + blen > first_pending);
        if (cm && (!simple || hazard)) return copy_match(len, dist, nfold);     // long / overlapping / marker / early source, pending bytes
        // plain deferred copy, predicated on `simple` (nothing happens for a literal)
        if constexpr (G <= MAX_G_DEFERRED) {
        cp_async_wait_but_one();
        commit_slot(o_dst, simple ? o_meta : 0u, ptog);
        {
            const uint8_t* src = out + bpos - bdist;
            const uint32_t so = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
            const uint32_t jb = (uint32_t)DB * (uint32_t)glane;
            const uint8_t* w0 = src - so + jb;
            #pragma unroll
            for (int k = 0; k < DW; k++) cp_async4_if(&S->stage_long[SLOTW * ptog + DW * glane + k], w0 + 4 * k, simple && jb + 4u * k < blen + so);
            cp_async_commit();
            o_dst = simple ? n_dst : o_dst; o_meta = simple ? n_meta : o_meta;
            n_dst = simple ? bpos : n_dst; n_meta = simple ? (blen | (so << 8)) : n_meta;
            ptog = simple ? slot_after(ptog) : ptog;
        }
        }
        pos += len;
        return R_OK;
    }

    // stored block body (src/infblocks.ts:278-333) with the Q2 truncation
    __device__ __forceinline__ int stored_block(uint32_t left)
    {
        uint32_t start = byte_pos();
        uint32_t n_in = in_len - start;
        uint32_t copied = 0;
        int r = R_OK;
        while (left) {
            if (n_in == 0) { stall_kind = ST_STORED; r = R_STALL; break; }
            if (ring.room() == 0) {
                int returns = ring.make_room();
                // `left` is a local of proc(): lost on return.  Block tasks of the large-stream path start with
                // an empty window model, so they take the whole block; sdz_large_plan replays the real window over the chain
                // of blocks and sends streams in which the reference would have lost `left` to the sequential decoder.
                if (returns && !spec && TM == TM_NONE) { left = 0; break; }
            }
            uint32_t t = min(min(left, n_in), (uint32_t)ring.room());
            ring.q += (int)t;
            copied += t; n_in -= t; left -= t;
        }
        if (copied > cap - pos) return R_OUTFULL;
        if (STORE) {
            const uint8_t* src = gsrc + start;
            // (eight independent loads in flight per lane: the plain loop waited for every byte before it stored it)
            uint32_t i = glane;
            if (MARK) {
                uint16_t* dst = out16 + pos;
                for (; i + 7u * G < copied; i += 8u * G) {
                    uint8_t v[8];
                    #pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = src[i + k * G];
                    #pragma unroll
                    for (int k = 0; k < 8; k++) dst[i + k * G] = v[k];
                }
                for (; i < copied; i += G) dst[i] = src[i];
            } else {
                uint8_t* dst = out + pos;
                for (; i + 7u * G < copied; i += 8u * G) {
                    uint8_t v[8];
                    #pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = src[i + k * G];
                    #pragma unroll
                    for (int k = 0; k < 8; k++) dst[i + k * G] = v[k];
                }
                for (; i < copied; i += G) dst[i] = src[i];
            }
        }
        pos += copied;
        seek(start + copied);
        return r;
    }

    // ------------------------------------------------------------------ phase machine
    // PH_FETCH: take the next stream, parse its container header (src/inflate.ts:142-401,
    // byte by byte from global memory) and position the bit reader on the first block.
    // block-task mode: work item i = the deflate block of stream 0 that starts at bit task_bit[i]
    __device__ __forceinline__ void fetch_task(const InflateParams& P, unsigned long long i)
    {
        in_len = P.in_len[0];
        gsrc = P.in + P.in_off[0];
        end_wp = (in_len + 3) / 4;
        total_chunks = (in_len + CH - 1) / CH;
        out_off = 0;
        abs_start = MARK ? P.task_out[i] : 0;
        out16 = MARK ? P.out16 + abs_start : nullptr;
        out = reinterpret_cast<uint8_t*>(out16);                    // byte view of the symbols (copy_match)
        resume_bit = MARK ? P.task_resume[i] : 0;
        limit = MARK ? P.task_limit[i] : 0xffffffffu;
        next_ck = TM == TM_INDEX ? P.ckpt_step : 0xffffffffu;
        pos = 0; cap = 0xffffffffu;
        msg = SDZ_MSG_NONE; stall_kind = ST_NONE;
        D = 0; dict_tail = nullptr;
        lbits = dbits = g_l = g_d = 0; eob_len = 0;
        ring.init(0);
        ref_F = ref_Fentry = ring_done = 0; blk_sym0_bit = 0; ref_on = ref_burst = ref_forced_slow = eob_emu = eob_fast = false;
        o_dst = o_meta = n_dst = n_meta = 0;
        is_gzip = false; method = 0; n_blocks = 0; mtime = 0; name_off = 0; name_len = 0; last = 0; raw = true;
        spec = false;
        const uint64_t sb = P.task_bit[i];
        if ((sb >> 3) >= (uint64_t)in_len) { finish_task(P, R_STALL); return; }
        seek((uint32_t)(sb >> 3));
        const int skip = (int)(sb & 7);
        if (bc < skip) { finish_task(P, R_STALL); return; }
        drop(skip);
        phase = PH_BLOCK;
    }

    // record of one block task: out_len = symbols produced, total_in = bit position after the block,
    // zstatus = R_EOB (block complete) / R_STALL / R_ERROR, n_blocks = BFINAL, container = BTYPE
    __device__ __forceinline__ void finish_task(const InflateParams& P, int r)
    {
        flush_pending();
        const uint64_t endbit = bit_pos();
        drain();
        __syncwarp(gmask);
        if (glane == 0) {
            sdz_result R;
            memset(&R, 0, sizeof R);
            R.out_len = pos;
            R.total_in = endbit;
            R.zstatus = r;
            R.n_blocks = (uint32_t)last;
            R.msg_id = (uint8_t)msg;
            R.container = (uint8_t)method;
            P.res[idx] = R;
        }
        phase = PH_FETCH;
    }

    __device__ __forceinline__ void fetch(const InflateParams& P)
    {
        unsigned long long i = 0;
        if (glane == 0) i = atomicAdd(P.counter, 1ull);
        i = __shfl_sync(gmask, i, 0, G);
        if (i >= (P.n_dev ? *P.n_dev : P.n)) { phase = PH_EXIT; return; }
        if (P.list) i = P.list[i];
        idx = i;
        if (TM != TM_NONE) { fetch_task(P, i); return; }
        in_len = P.in_len[i];
        const uint8_t mode_raw = P.mode[i];
        const int mode = mode_raw & 0x7f;
        const bool has_dict = (mode_raw & 0x80) != 0;
        const uint8_t* src = P.in + P.in_off[i];
        out_off = P.out_off ? P.out_off[i] : 0;
        gsrc = src;
        end_wp = (in_len + 3) / 4;
        total_chunks = (in_len + CH - 1) / CH;
        out = STORE ? P.out + out_off : nullptr;
        pos = 0;
        cap = (STORE && P.out_cap) ? P.out_cap[i] : 0xffffffffu;
        msg = SDZ_MSG_NONE;
        stall_kind = ST_NONE;
        D = 0; dict_tail = nullptr;
        lbits = dbits = g_l = g_d = 0; eob_len = 0;
        ring.init(0);
        ref_F = ref_Fentry = ring_done = 0; blk_sym0_bit = 0; ref_on = ref_burst = ref_forced_slow = eob_emu = eob_fast = false;
        o_dst = o_meta = n_dst = n_meta = 0;
        is_gzip = false; method = 0; n_blocks = 0; mtime = 0; name_off = 0; name_len = 0; last = 0;
        blk_hdr_bit = last_sym_bit = 0; ref_floor = 0; resume_first = false; hdr_counted = false; resume_bit = 0;
        spec = TM == TM_NONE && P.spec != 0;

        int thrown = SDZ_THROW_NONE, thrown_inflate = 0, zstatus = SDZ_Z_OK;
        bool decode = true;
        raw = mode == SDZ_MODE_RAW;
        if (P.resume_in && P.resume_in[i].kind >= SDZ_RESUME_AT_BLOCK && P.resume_in[i].kind <= SDZ_RESUME_AT_TRAILER) {
            // a streaming session past its container header: put the decoder where the reference stopped
            const sdz_resume z = P.resume_in[i];
            is_gzip = (z.flags & 1) != 0; raw = (z.flags & 2) != 0; last = (z.flags & 4) ? 1 : 0;
            method = z.method; mtime = z.mtime; name_off = z.name_off; name_len = z.name_len;
            n_blocks = z.n_blocks; pos = z.pos;
            D = (int)z.dict_used;
            if (D) dict_tail = P.dict + P.dict_off[i] + (P.dict_len[i] - (uint32_t)D);
            ring.q = ring.r = z.ring_q; ring.ao = OUTBUF;               // append() returned with the window flushed
            ring_done = pos;
            ref_F = ref_floor = z.prev_len;
            seek((uint32_t)(z.block_bit >> 3));
            drop((int)(z.block_bit & 7));
            if (z.kind == SDZ_RESUME_AT_TRAILER) { finish_stream(P, R_EOB); return; }
            if (z.kind == SDZ_RESUME_IN_CODES) { resume_bit = z.sym_bit; resume_first = true; }
            phase = PH_BLOCK;
            return;
        }
        if (mode == SDZ_MODE_SNIFF) {                                   // inflate(): src/sd-inflate.ts:194-207
            if (in_len < 2) { thrown_inflate = SDZ_THROW_TOO_SMALL; decode = false; }
            else {
                uint32_t b0 = src[0], b1 = src[1];
                bool ident = (b0 == 0x78 && (((b0 << 8) + b1) % 31) == 0) || (b0 == 0x1f && b1 == 0x8b);
                raw = !ident;
            }
        }
        if (decode && raw && has_dict) { thrown_inflate = SDZ_THROW__COUNT; decode = false; }  // RangeError (src/sd-inflate.ts:69-71)
        if (in_len == 0) decode = false;                                // append() of an empty chunk returns []

        uint32_t hp = 0, end_byte = 0;
        if (decode && !raw && spec) {
            // SDZ_PARITY_SPEC: the container header as zlib 1.3 reads it (inflate.c, states HEAD .. DICTID)
            int err = 0;
            bool ok = false;
            do {
                if (in_len < 2) { hp = in_len; break; }
                const uint32_t h0 = src[0], h1 = src[1];
                hp = 2;
                if (h0 == 0x1f && h1 == 0x8b) {
                    is_gzip = true;
                    if (in_len < 4) { hp = in_len; break; }
                    method = (int)src[2];
                    const uint32_t gflags = src[3];
                    hp = 4;
                    if (method != 8) { err = SDZ_MSG_BAD_METHOD; break; }
                    if (gflags & 0xe0) { err = SDZ_MSG_BAD_GZIP_FLAGS; break; }
                    if (in_len < 10) { hp = in_len; break; }
                    mtime = (int32_t)((uint32_t)src[4] | ((uint32_t)src[5] << 8) | ((uint32_t)src[6] << 16) | ((uint32_t)src[7] << 24));
                    hp = 10;
                    bool trunc = false;
                    if (gflags & 4) {                                   // RFC 1952: XLEN (little-endian) + that many bytes
                        if (hp + 2 > in_len) { hp = in_len; break; }
                        const uint32_t xlen = (uint32_t)src[hp] | ((uint32_t)src[hp + 1] << 8);
                        hp += 2;
                        if (hp + xlen > in_len) { hp = in_len; break; }
                        hp += xlen;
                    }
                    if (gflags & 8) {
                        name_off = hp;
                        for (;;) { if (hp >= in_len) { trunc = true; break; } const uint32_t b = src[hp++]; if (b == 0) break; name_len++; }
                        if (trunc) break;
                    }
                    if (gflags & 16) {
                        for (;;) { if (hp >= in_len) { trunc = true; break; } const uint32_t b = src[hp++]; if (b == 0) break; }
                        if (trunc) break;
                    }
                    if (gflags & 2) {                                   // FHCRC: low half of the CRC-32 of the header so far
                        if (hp + 2 > in_len) { hp = in_len; break; }
                        uint32_t c = 0xffffffffu;
                        for (uint32_t k = 0; k < hp; k++) {
                            c ^= src[k];
                            for (int t = 0; t < 8; t++) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
                        }
                        c = ~c;
                        const uint32_t want = (uint32_t)src[hp] | ((uint32_t)src[hp + 1] << 8);
                        hp += 2;
                        if ((c & 0xffffu) != want) { err = SDZ_MSG_BAD_HEADER_CRC; break; }
                    }
                } else {
                    if (((h0 << 8) + h1) % 31 != 0) { err = SDZ_MSG_BAD_HEADER_CHECK; break; }
                    method = (int)h0;
                    if ((h0 & 0xf) != 8) { err = SDZ_MSG_BAD_METHOD; break; }
                    if ((h0 >> 4) + 8 > 15) { err = SDZ_MSG_BAD_WINDOW; break; }
                    if (h1 & 0x20) {
                        if (hp + 4 > in_len) { hp = in_len; break; }
                        int32_t dictid = (int32_t)(((uint32_t)src[hp] << 24) | ((uint32_t)src[hp + 1] << 16) |
                                                   ((uint32_t)src[hp + 2] << 8) | src[hp + 3]);
                        hp += 4;
                        if (!has_dict) { thrown = SDZ_THROW_DICT_REQUIRED; zstatus = SDZ_Z_NEED_DICT; break; }
                        if (P.dict_adler[i] != dictid) { thrown = SDZ_THROW_DICT_INVALID; zstatus = SDZ_Z_NEED_DICT; break; }
                        uint32_t dl = P.dict_len[i];
                        uint32_t used = dl >= (uint32_t)WSIZE ? (uint32_t)WSIZE : dl;      // (the reference keeps 32,767: SURVEY Q14)
                        D = (int)used;
                        dict_tail = P.dict + P.dict_off[i] + (dl - used);
                        ring.init((int)min(used, (uint32_t)WSIZE - 1u));
                    }
                }
                ok = true;
            } while (0);
            if (err) { msg = err; thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_DATA_ERROR; }
            if (!ok) { decode = false; end_byte = thrown ? hp : in_len; }
        } else
        if (decode && !raw) {
            int err = 0;
            bool ok = false;
            do {
                uint32_t b;
                if (hp >= in_len) break;
                if (src[hp] == 0x1f) {
                    hp++;
                    if (hp >= in_len) break;
                    b = src[hp++];
                    if (b != 0x8b) { err = SDZ_MSG_BAD_GZIP_ID; break; }
                    is_gzip = true;
                }
                if (hp >= in_len) break;
                method = src[hp++];
                if ((method & 0xf) != 8) { err = SDZ_MSG_BAD_METHOD; break; }
                if ((method >> 4) + 8 > 15) { err = SDZ_MSG_BAD_WINDOW; break; }
                if (hp >= in_len) break;
                b = src[hp++];
                if (is_gzip) {
                    const uint32_t gflags = b;
                    bool trunc = false;
                    for (int k = 0; k < 4; k++) {
                        if (hp >= in_len) { trunc = true; break; }
                        mtime = (int32_t)(((uint32_t)mtime >> 8) | ((uint32_t)src[hp++] << 24));
                    }
                    if (trunc) break;
                    if (hp + 2 > in_len) { hp = in_len; break; }        // XFL, OS
                    hp += 2;
                    if (gflags & 4) { hp = in_len; break; }             // FEXTRA never leaves EXTRA0
                    if (gflags & 8) {
                        name_off = hp;
                        for (;;) { if (hp >= in_len) { trunc = true; break; } b = src[hp++]; if (b == 0) break; name_len++; }
                        if (trunc) break;
                    }
                    if (gflags & 16) {
                        for (;;) { if (hp >= in_len) { trunc = true; break; } b = src[hp++]; if (b == 0) break; }
                        if (trunc) break;
                    }
                    if (gflags & 2) { if (hp + 2 > in_len) { hp = in_len; break; } hp += 2; }
                } else {
                    if ((((uint32_t)method << 8) + b) % 31 != 0) { err = SDZ_MSG_BAD_HEADER_CHECK; break; }
                    if (b & 0x20) {
                        if (hp + 4 > in_len) { hp = in_len; break; }
                        int32_t dictid = (int32_t)(((uint32_t)src[hp] << 24) | ((uint32_t)src[hp + 1] << 16) |
                                                   ((uint32_t)src[hp + 2] << 8) | src[hp + 3]);
                        hp += 4;
                        // NEED_DICT -> inflateSetDictionary (src/sd-inflate.ts:116-126, src/inflate.ts:475-503)
                        if (!has_dict) { thrown = SDZ_THROW_DICT_REQUIRED; zstatus = SDZ_Z_NEED_DICT; break; }
                        if (P.dict_adler[i] != dictid) { thrown = SDZ_THROW_DICT_INVALID; zstatus = SDZ_Z_NEED_DICT; break; }
                        uint32_t dl = P.dict_len[i];
                        uint32_t used = dl >= (uint32_t)WSIZE ? (uint32_t)WSIZE - 1 : dl;
                        D = (int)used;
                        dict_tail = P.dict + P.dict_off[i] + (dl - used);
                        ring.init((int)used);
                    }
                }
                ok = true;
            } while (0);
            if (err) { msg = err; thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_DATA_ERROR; }
            if (!ok) { decode = false; end_byte = thrown ? hp : in_len; }
        }
        if (!decode) {
            write_record(P, thrown, thrown_inflate, zstatus, false, 0, 0, end_byte);
            if (P.resume_out && glane == 0) {                           // the container header is parsed again from byte 0
                sdz_resume z;
                memset(&z, 0, sizeof z);
                z.kind = (uint8_t)(thrown || thrown_inflate ? SDZ_RESUME_FAILED : SDZ_RESUME_START);
                z.prev_len = in_len;
                P.resume_out[idx] = z;
            }
            return;                                                     // stays in PH_FETCH
        }
        seek(hp);
        phase = PH_BLOCK;
    }

    __device__ __forceinline__ void write_record(const InflateParams& P, int thrown, int thrown_inflate, int zstatus, bool done,
                                                 int32_t stored, int32_t isize, uint32_t end_byte)
    {
        if (glane == 0) {
            sdz_result R;
            R.out_off = out_off;
            R.out_len = (thrown && STORE) ? 0 : pos;    // the sizing pass always reports the decoded size
            R.total_in = end_byte;
            R.zstatus = zstatus;
            R.stored_checksum = stored;
            R.running_checksum = 0;
            R.stored_isize = isize;
            R.mtime = mtime;
            R.name_off = name_len ? name_off : 0;
            R.name_len = name_len;
            R.n_blocks = n_blocks;
            R.msg_id = (uint8_t)msg;
            R.thrown_append = (uint8_t)thrown;
            R.thrown_inflate = (uint8_t)thrown_inflate;
            R.container = (uint8_t)(is_gzip ? SDZ_GZIP : (method == 0 ? SDZ_RAW : SDZ_ZLIB));
            R.complete = done ? 1 : 0;
            R.checksum_state = R.size_state = R.success = R.have_running = 0;
            for (int k = 0; k < 7; k++) R.reserved[k] = 0;
            P.res[idx] = R;
        }
    }

    // The stream stops here: r = R_EOB (final block complete), R_STALL, R_ERROR or R_OUTFULL.
    __device__ __forceinline__ void finish_stream(const InflateParams& P, int r)
    {
        if (TM != TM_NONE) { finish_task(P, r); return; }
        int thrown = SDZ_THROW_NONE, zstatus = SDZ_Z_OK;
        bool done = false;
        int32_t stored = 0, isize = 0;
        uint32_t end_byte = byte_pos();
        int rk = SDZ_RESUME_FAILED;                                     // streaming sessions: where the next append() continues
        uint64_t rk_bit = 0;
        if (r == R_ERROR) { thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_DATA_ERROR; }
        else if (r == R_OUTFULL) { zstatus = SDZ_Z_BUF_ERROR; rk = -1; }
        else if (r == R_STALL) {
            // input exhausted.  proc() flushes; if that fills the 16 KiB buffer append() calls
            // again, and BTREE/DTREE cannot be re-entered -> STREAM_ERROR is thrown.
            if (stall_kind == ST_DYNHDR) {
                ring.flush();
                if (ring.ao == 0 && !spec) { thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_STREAM_ERROR; msg = SDZ_MSG_NONE; }
                else rk = SDZ_RESUME_BROKEN_Q3;
            } else if (stall_kind == ST_STORED) {
                // `left` is gone when proc() is entered again: the stored block ends where the input ended
                rk = last ? SDZ_RESUME_AT_TRAILER : SDZ_RESUME_AT_BLOCK;
                rk_bit = bit_pos();
            } else if (stall_kind == ST_CODES) {
                rk = SDZ_RESUME_IN_CODES;
            } else {
                rk = SDZ_RESUME_AT_BLOCK;
                rk_bit = blk_hdr_bit;
                if (hdr_counted) n_blocks--;                            // the header is parsed (and counted) again
            }
        } else {
            // final block done: unused whole bytes go back, partial bits are dropped (src/inflate.ts:409-421);
            // the trailer is read byte by byte (src/inflate.ts:423-463)
            uint32_t tp = (uint32_t)((bit_pos() + 7) >> 3);
            if (raw) {
                done = true;
            } else {
                const int nbytes = is_gzip ? 8 : 4;
                int k = 0;
                for (; k < nbytes && tp < in_len; k++) {
                    uint32_t b = gsrc[tp++];
                    if (is_gzip) {
                        if (k < 4) stored = (int32_t)(((uint32_t)stored >> 8) | (b << 24));
                        else isize = (int32_t)(((uint32_t)isize >> 8) | (b << 24));
                    } else {
                        stored = (int32_t)(((uint32_t)stored << 8) | b);
                    }
                }
                done = k == nbytes;
            }
            if (done) {
                zstatus = SDZ_Z_STREAM_END;
                if (tp < in_len && !spec) thrown = SDZ_THROW_HANG;   // bytes after the end: append() spins
            }
            rk = done ? SDZ_RESUME_DONE : SDZ_RESUME_AT_TRAILER;
            rk_bit = bit_pos();
            end_byte = tp;
        }
        flush_pending();
        drain();
        __syncwarp(gmask);
        write_record(P, thrown, 0, zstatus, done, stored, isize, end_byte);
        if constexpr (TM == TM_NONE) {
            if (P.resume_out && rk >= 0 && glane == 0) {
                // append() returns: everything that fits leaves the window, and append() calls again for as long as the
                // 16 KiB buffer comes back full (src/sd-inflate.ts:101-150), so the window is empty afterwards
                for (;;) { ring.flush(); if (ring.ao != 0) break; ring.ao = OUTBUF; }
                sdz_resume z;
                memset(&z, 0, sizeof z);
                z.kind = (uint8_t)(thrown ? SDZ_RESUME_FAILED : rk);
                z.block_bit = rk == SDZ_RESUME_IN_CODES ? blk_hdr_bit : rk_bit;
                z.sym_bit = rk == SDZ_RESUME_IN_CODES ? last_sym_bit : 0;
                z.pos = pos;
                z.ring_q = ring.q;
                z.n_blocks = rk == SDZ_RESUME_IN_CODES ? n_blocks - 1u : n_blocks;
                z.mtime = mtime; z.name_off = name_off; z.name_len = name_len;
                z.prev_len = in_len;
                z.dict_used = (uint16_t)D;
                z.flags = (uint8_t)((is_gzip ? 1 : 0) | (raw ? 2 : 0) | (last ? 4 : 0));
                z.method = (uint8_t)method;
                P.resume_out[idx] = z;
            }
        }
        phase = PH_FETCH;
    }

    // PH_BLOCK: one block header.  Stored blocks are completed here; for fixed / dynamic blocks
    // the tables are built and the group joins the lockstep symbol loop (PH_CODES).
    __device__ __forceinline__ void block_begin(const InflateParams& P)
    {
        if constexpr (TM == TM_NONE) { blk_hdr_bit = bit_pos(); hdr_counted = false; }
        if (!ensure(3)) { stall_kind = ST_OTHER; finish_stream(P, R_STALL); return; }
        if (TM == TM_NONE) ref_need(bit_pos(), 3u);
        uint32_t t = peek(3);
        drop(3);
        last = (int)(t & 1);
        n_blocks++;
        hdr_counted = true;
        start_pos = pos;
        ring_done = pos;
        const uint32_t type = t >> 1;
        if (TM != TM_NONE) method = (int)type;                           // block-task records carry BTYPE
        int r;
        if (type == 0) {
            drop(bc & 7);
            if (!ensure(32)) { stall_kind = ST_OTHER; finish_stream(P, R_STALL); return; }
            uint32_t v = (uint32_t)bb;
            if ((((~v) >> 16) & 0xffff) != (v & 0xffff)) { msg = SDZ_MSG_BAD_STORED_LEN; finish_stream(P, R_ERROR); return; }
            drop(32);
            r = stored_block(v & 0xffff);
            if (r != R_OK) { finish_stream(P, r); return; }
            if (TM != TM_NONE) { finish_task(P, R_EOB); return; }
            if (last) { ring.wash(); finish_stream(P, R_EOB); }
            return;                                                     // next block: stays in PH_BLOCK
        }
        if (type == 3) { msg = SDZ_MSG_BAD_BLOCK_TYPE; finish_stream(P, R_ERROR); return; }
        int nl = 288, nd = 30;
        if (type == 1) {
            __syncwarp(gmask);
            uint8_t* lens = reinterpret_cast<uint8_t*>(gsorted + SORTED_L + SORTED_D);
            for (int k = glane; k < 320; k += G) {
                uint8_t v = k < 144 ? 8 : (k < 256 ? 9 : (k < 280 ? 7 : (k < 288 ? 8 : 5)));
                lens[k] = v;
            }
            __syncwarp(gmask);
        } else {
            r = dynamic_header(&nl, &nd);
            if (r != R_OK) { finish_stream(P, r); return; }
        }
        TreeInfo T = build_tables<G>(S, gsorted, nl, nd, type == 1, glane, gmask, long_l(), LONG_N, spec);
        if (T.msg) { msg = T.msg; finish_stream(P, R_ERROR); return; }
        lbits = T.lbits; dbits = T.dbits; g_l = T.g_l; g_d = T.g_d;
        if (MARK && resume_bit) {                                       // continue in the middle of the block
            seek((uint32_t)(resume_bit >> 3));
            const int skip = (int)(resume_bit & 7);
            if (bc < skip) { finish_task(P, R_STALL); return; }
            drop(skip);
        }
        blk_sym0_bit = bit_pos();
        if (TM == TM_NONE && resume_bit) {                              // streaming session: the symbol the last append() stopped at
            seek((uint32_t)(resume_bit >> 3));
            drop((int)(resume_bit & 7));
            resume_bit = 0;
        }
        ref_burst = false; ref_forced_slow = false;                    // a new codes object: the START check comes first
        phase = PH_CODES;
    }

    // TM_INDEX: called at a symbol boundary of the lockstep loop
    __device__ __forceinline__ void checkpoint(const InflateParams& P)
    {
        if (pos >= next_ck) {
            if (glane == 0) {
                const unsigned long long slot = atomicAdd(P.ckpt_count, 1ull);
                if (slot < P.ckpt_cap) { Ckpt c; c.bit = bit_pos(); c.task = (uint32_t)idx; c.pos = pos; P.ckpt[slot] = c; }
            }
            next_ck = (pos | (P.ckpt_step - 1u)) + 1u;
        }
    }

    // a step() returned something other than R_OK
    __device__ __forceinline__ void block_end(const InflateParams& P, int r)
    {
        if (TM != TM_NONE) { finish_task(P, r); return; }
        ring.write(pos - ring_done); ring_done = pos;
        if (r != R_EOB) { if (r == R_STALL) stall_kind = ST_CODES; finish_stream(P, r); return; }
        // End of block.  When inflate_fast() decodes the EOB its STREAM_END status leaks through
        // WASH's early return (src/infcodes.ts:264,:357,:627-638 -> src/infblocks.ts:560-564), so
        // the block completes after ONE flush attempt; only a slow-path EOB (fewer than 258 bytes
        // of window room or fewer than 10 unread input bytes, src/infcodes.ts:339) washes the
        // window completely, returning to append() as often as needed.  Far from the end of the input only the
        // window room decides; within the last 16 bytes step_general() has followed the reference's input frontier
        // symbol by symbol (ref_symbol_begin).
        {
            const bool fast_eob = eob_emu ? eob_fast : ring.room() >= 258;
            if (fast_eob) ring.flush(); else ring.wash();
        }
        if (last) { ring.wash(); finish_stream(P, R_EOB); return; }         // DRY (src/infblocks.ts:579-594)
        phase = PH_BLOCK;
    }
};

#ifndef SDZ_MINBLOCKS
#define SDZ_MINBLOCKS 4                // 128 registers: eight 64-thread blocks per SM
#endif

// One sub-warp group of G lanes per stream; the 32 / G groups of a warp run the symbol loop in
// lockstep (they re-converge at the ballot after every symbol), so one instruction stream
// serves 32 / G streams.  Block headers, table builds and stream changes are serviced between
// lockstep runs while the other groups of the warp wait.
template <int G, bool STORE, int TM = TM_NONE>
__global__ void __launch_bounds__(128, SDZ_MINBLOCKS) inflate_kernel(InflateParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gid = threadIdx.x / G;
    GroupSmem* S = reinterpret_cast<GroupSmem*>(smem_raw) + gid;

    Decoder<G, STORE, TM> d;
    d.S = S;
    d.gsorted = P.scratch + ((size_t)blockIdx.x * (blockDim.x / G) + gid) * SCRATCH_U16;
    d.glane = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    d.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - d.glane));
    d.issued_abs = 0; d.waited_abs = 0; d.chunk0 = 0; d.phasebits = 0;
    d.ptog = 0; d.o_dst = d.o_meta = d.n_dst = d.n_meta = 0;
    d.phase = PH_FETCH;
    if (d.glane == 0) {
        for (int i = 0; i < NBUF; i++) mbar_init(&S->mbar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    for (;;) {
        if (d.phase == PH_FETCH) d.fetch(P);
        else if (d.phase == PH_BLOCK) d.block_begin(P);
        __syncwarp();
        if (__all_sync(0xffffffffu, d.phase == PH_EXIT)) break;
        // lockstep symbol loop: runs until some group needs a new block or a new stream
        for (;;) {
            __syncwarp();                            // re-converge + order the stores of earlier iterations
            if (d.phase == PH_CODES) {
                if (TM == TM_INDEX) d.checkpoint(P);
                int r = d.step_flat();
                if (TM == TM_MARK && r == R_OK && d.pos >= d.limit) r = R_EOB;     // the piece is complete
                if (r != R_OK) d.block_end(P, r);
            }
            if (__ballot_sync(0xffffffffu, d.phase != PH_CODES) != 0u) {
                if (__any_sync(0xffffffffu, d.phase == PH_FETCH || d.phase == PH_BLOCK)) break;
                if (__all_sync(0xffffffffu, d.phase == PH_EXIT)) break;
            }
        }
    }
}

}  // namespace sdz
