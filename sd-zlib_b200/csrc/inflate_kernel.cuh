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
//   * compressed input is staged per group into a 4 x 128 B shared-memory ring by TMA bulk
//     copies (cp.async.bulk + mbarrier); the decoder keeps a 64-bit bit buffer and one
//     prefetched word in registers, so the ring read is off the decode critical path;
//   * per-group decode tables live in shared memory: a 2^RL-entry u16 literal/length LUT and
//     a 2^RD-entry u16 distance LUT with base/extra pre-baked; codes longer than the root
//     fall back to a canonical (count/sorted-symbol) decode;
//   * every lane of a group holds the same decoder state (no shuffles on the critical path);
//     lanes only differ in which bytes of a match / stored block they move.
#pragma once
#include "sdz_device.cuh"
#include "../../include/sdz_codes.h"

namespace sdz {

constexpr int RL = 10;                 // literal/length LUT root bits
constexpr int RD = 8;                  // distance LUT root bits
constexpr int CH = 128;                // bytes per TMA bulk copy
constexpr int NBUF = 4;                // chunks in the per-group input ring
constexpr int CHW = CH / 4;
constexpr int WSIZE = 32768;           // reference window (src/inflate.ts:98)
constexpr int OUTBUF = 16384;          // reference OUTPUT_BUFSIZE (src/zstream.ts:11)

constexpr uint32_t E_LONG = 0x0ffe;    // root entry: code longer than the root -> canonical decode
constexpr uint32_t E_INVALID = 0x0fff; // root entry: no code has this prefix

struct alignas(16) GroupSmem {
    uint32_t ring[NBUF * CHW];         // 512 B input staging
    uint64_t mbar[NBUF];
    uint16_t lut_l[1 << RL];
    uint16_t lut_d[1 << RD];
    uint16_t sorted_l[288];            // symbols ordered by (code length, symbol)
    uint16_t sorted_d[32];
    uint32_t cnt_l[16];                // codes per length (unpadded)
    uint32_t cnt_d[16];
    uint32_t aux[32];                  // build scratch: [0..15] offsets, [16..31] first codes
    uint8_t lens[320];                 // code lengths of the current block
};

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
    unsigned long long n;
    unsigned long long* counter;       // dynamic stream scheduler
};

// how a decode step ended
enum : int { R_OK = 0, R_EOB = 1, R_STALL = 2, R_ERROR = 3, R_OUTFULL = 4 };

// where input ran out (only the classes the record depends on)
enum : int { ST_NONE = 0, ST_OTHER = 1, ST_DYNHDR = 2 /* BTREE/DTREE: not resumable, SURVEY Q3 */ };

__device__ __constant__ uint8_t c_border[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// Scalar model of InfBlocks' window pointers and ZStream.avail_out (src/infblocks.ts:61-121,
// :154, :289-313; src/infcodes.ts:547-573).  Only byte COUNTS flow through it.
struct RingModel {
    int q, r, ao;                      // write, read, avail_out
    __device__ __forceinline__ void init(int dict_used) { q = r = dict_used; ao = OUTBUF; }
    __device__ __forceinline__ int room() const { return q < r ? r - q - 1 : WSIZE - q; }
    __device__ void flush()
    {
        int n = (r <= q ? q : WSIZE) - r;
        n = min(n, ao); ao -= n; r += n;
        if (r == WSIZE) {
            r = 0;
            if (q == WSIZE) q = 0;
            n = min(q - r, ao); ao -= n; r += n;
        }
    }
    // the "no room" dance; returns how many times proc() went back to append()
    __device__ int make_room()
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
    __device__ void write(uint32_t n)
    {
        while (n) {
            int m = room();
            if (!m) { make_room(); m = room(); }
            uint32_t t = min(n, (uint32_t)m);
            q += (int)t; n -= t;
        }
    }
    // WASH / DRY: everything must leave the window before the block ends (src/infcodes.ts:626-639)
    __device__ void wash()
    {
        flush();
        while (r != q) { flush(); ao = OUTBUF; flush(); }
    }
};

template <int G, bool STORE>
struct Decoder {
    GroupSmem* S;
    unsigned gmask;
    int glane;

    // ---- bit reader
    uint64_t bb;
    int bc;
    uint32_t nw;                       // prefetched word `wp`
    uint32_t wp, end_wp, tail_mask;
    uint32_t in_len;
    const uint8_t* gsrc;
    uint32_t chunk0, total_chunks, issued, waited, base_seq;

    // ---- output
    uint8_t* out;
    uint32_t pos, cap;
    const uint8_t* dict_tail;
    int D;

    RingModel ring;
    int msg;
    int stall_kind;

    // reference root widths of the current block's trees (src/inftree.ts:146-165)
    int lbits, dbits, g_l, g_d;
    int last_klen;                     // length of the code slow_code() decoded last
    int eob_len;                       // code length of the end-of-block symbol just decoded

    // ------------------------------------------------------------------ input staging
    __device__ __forceinline__ void issue_chunk(uint32_t rel)
    {
        uint32_t seq = base_seq + rel;
        uint32_t slot = seq % NBUF;
        if (glane == 0) {
            mbar_arrive_expect_tx(&S->mbar[slot], CH);
            bulk_copy_g2s(&S->ring[slot * CHW], gsrc + (size_t)(chunk0 + rel) * CH, CH, &S->mbar[slot]);
        }
    }

    __device__ __forceinline__ uint32_t load_word(uint32_t w)
    {
        if (w >= end_wp) return 0;
        uint32_t rel = w / CHW - chunk0;
        uint32_t seq = base_seq + rel;
        if (rel >= waited) {
            mbar_wait(&S->mbar[seq % NBUF], (seq / NBUF) & 1);
            waited = rel + 1;
            if (rel >= 1 && chunk0 + issued < total_chunks) {
                __syncwarp(gmask);      // every lane is done with the slot being recycled
                issue_chunk(issued);
                issued++;
            }
        }
        uint32_t v = S->ring[(seq % NBUF) * CHW + (w % CHW)];
        if (w + 1 == end_wp) v &= tail_mask;
        return v;
    }

    __device__ void drain()
    {
        for (uint32_t rel = waited; rel < issued; rel++) {
            uint32_t seq = base_seq + rel;
            mbar_wait(&S->mbar[seq % NBUF], (seq / NBUF) & 1);
        }
        waited = issued;
    }

    // (re)position the reader on a byte boundary of the stream
    __device__ void seek(uint32_t byte_pos)
    {
        drain();
        __syncwarp(gmask);
        base_seq += issued;
        issued = 0; waited = 0;
        uint32_t w = byte_pos >> 2;
        chunk0 = w / CHW;
        uint32_t avail_chunks = total_chunks > chunk0 ? total_chunks - chunk0 : 0;
        uint32_t n0 = avail_chunks < (uint32_t)NBUF ? avail_chunks : (uint32_t)NBUF;
        for (uint32_t i = 0; i < n0; i++) issue_chunk(i);
        issued = n0;
        wp = w; bb = 0; bc = 0;
        nw = load_word(wp);
        refill();
        uint32_t skip = (byte_pos & 3) * 8;
        if (skip) { uint32_t s = min(skip, (uint32_t)bc); bb >>= s; bc -= (int)s; }
        refill();
    }

    __device__ __forceinline__ void refill()
    {
        if (bc <= 32 && wp < end_wp) {
            bb |= (uint64_t)nw << bc;
            bc += (wp + 1 < end_wp) ? 32 : (int)((in_len - wp * 4u) * 8u);
            wp++;
            nw = load_word(wp);
        }
    }

    // total bits between the read position and the end of the input
    __device__ __forceinline__ uint32_t avail_bits() const
    {
        uint64_t unloaded = wp < end_wp ? (uint64_t)in_len * 8 - (uint64_t)wp * 32 : 0;
        uint64_t a = (uint64_t)bc + unloaded;
        return a > 4096 ? 4096u : (uint32_t)a;
    }
    __device__ __forceinline__ uint32_t byte_pos() const
    {
        // bits loaded so far are whole bytes; bc of them are still unread
        uint64_t loaded = wp < end_wp ? (uint64_t)wp * 32 : (uint64_t)in_len * 8;
        return (uint32_t)((loaded - (uint64_t)bc) >> 3);
    }
    __device__ __forceinline__ uint64_t bit_pos() const
    {
        uint64_t loaded = wp < end_wp ? (uint64_t)wp * 32 : (uint64_t)in_len * 8;
        return loaded - (uint64_t)bc;
    }
    __device__ __forceinline__ bool ensure(int n) { refill(); return bc >= n; }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; bc -= n; }

    // ------------------------------------------------------------------ output
    __device__ __forceinline__ int put_literal(uint32_t v)
    {
        if (pos >= cap) return R_OUTFULL;
        if (STORE) { if (glane == 0) out[pos] = (uint8_t)v; }
        pos++;
        return R_OK;
    }

    __device__ __forceinline__ int copy_match(uint32_t len, uint32_t dist)
    {
        if (len > cap - pos) return R_OUTFULL;
        if (STORE) {
            __syncwarp(gmask);                       // earlier stores of this group are visible
            uint8_t* dst = out + pos;
            if (dist <= pos) {
                const uint8_t* src = dst - dist;
                if (dist >= len) {
                    for (uint32_t i = glane; i < len; i += G) dst[i] = src[i];
                } else if (dist == 1) {
                    uint8_t v = src[0];
                    for (uint32_t i = glane; i < len; i += G) dst[i] = v;
                } else {                             // lane-strided replicate of the period
                    for (uint32_t i = glane; i < len; i += G) dst[i] = src[i % dist];
                }
            } else {
                // source starts before the output: preset dictionary tail, else the
                // reference's zero-initialised window (src/infcodes.ts:174-193)
                for (uint32_t i = glane; i < len; i += G) {
                    uint32_t k = dist >= len ? i : i % dist;
                    int64_t s = (int64_t)pos - (int64_t)dist + (int64_t)k;
                    uint8_t v = 0;
                    if (s >= 0) v = out[s];
                    else if (s >= -(int64_t)D) v = dict_tail[(int64_t)D + s];
                    dst[i] = v;
                }
            }
        }
        pos += len;
        return R_OK;
    }

    // ------------------------------------------------------------------ reference table geometry
    // Width of the sub-table huft_build creates for the codes that share the first `w` bits
    // `prefix` (MSB-first) - src/inftree.ts:217-239.  cnt[] unpadded, `pad` dummy codes at g.
    __device__ int ref_subtable_width(const uint32_t* cnt, int g, int pad, int l, int w, uint32_t prefix) const
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
    // (src/inftree.ts:217-246): the root table plus one sub-table per distinct l-bit
    // (and, for very long codes, 2l-bit) prefix.  Evaluated by the group's lanes in parallel.
    __device__ int ref_table_total(const uint32_t* cnt, int g, int pad, int l) const
    {
        int total = 1 << l;
        for (int w = l; w < g; w += l) {
            // first code longer than w bits -> first prefix that owns a sub-table at this level
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

    // Decode one code the slow way (canonical counts), applying the reference's lookahead
    // rule: a lookup happens only when the table's index width is available.
    __device__ __noinline__ int slow_code(const uint32_t* cnt, const uint16_t* sorted, int l, int g, uint32_t* sym_out)
    {
        refill();
        int A = (int)avail_bits();
        if (A < l) return R_STALL;
        int ncodes = 0, kraft = 0;
        for (int k = 1; k <= g; k++) { ncodes += (int)cnt[k]; }
        kraft = 0;
        {
            int y = 1;
            for (int k = 1; k <= g; k++) { y <<= 1; y -= (int)cnt[k]; }
            kraft = y;                                  // unused codes of length g
        }
        if (g == 1 && ncodes == 1) {                    // the one incomplete set the reference accepts
            if (bb & 1) return R_ERROR;                 // exop 192: invalid code
            drop(1);
            last_klen = 1;
            *sym_out = sorted[0] & 0xfffu;
            return R_OK;
        }
        int code = 0, first = 0, index = 0, klen = 0;
        bool found = false;
        uint32_t sym = 0;
        int lim = min(A, 15);
        for (int len = 1; len <= lim; len++) {
            code |= (int)((bb >> (len - 1)) & 1);
            int count = len <= g ? (int)cnt[len] : 0;
            if (code - count < first) {
                sym = sorted[index + (code - first)] & 0xfffu;
                klen = len; found = true;
                break;
            }
            index += count; first += count; first <<= 1; code <<= 1;
        }
        last_klen = klen;
        if (found && klen <= l) { drop(klen); *sym_out = sym; return R_OK; }
        if (!found && A >= g) return R_ERROR;           // every bit of the longest code is there: no such code
        // walk the reference's table levels
        int w = l;
        for (int level = 0; level < 4; level++) {
            if (A - w < 1) return R_STALL;
            uint32_t prefix = __brev((uint32_t)bb) >> (32 - w);
            int j = ref_subtable_width(cnt, g, kraft, l, w, prefix);
            if (A - w < j) return R_STALL;
            if (found && klen <= w + j) { drop(klen); *sym_out = sym; return R_OK; }
            if (j == 0 || (!found && w + j >= 15)) return R_ERROR;
            w += l;
        }
        return R_ERROR;
    }

    // distance code + extra bits after a length has been consumed
    __device__ __noinline__ int slow_dist(uint32_t* dist_out)
    {
        uint32_t ds;
        if (g_d == 0) return R_ERROR;                   // no distance codes at all (nl == 257 blocks)
        int r = slow_code(S->cnt_d, S->sorted_d, dbits, g_d, &ds);
        if (r != R_OK) { if (r == R_ERROR) msg = SDZ_MSG_BAD_DIST_CODE; return r; }
        if (ds > 29) { msg = SDZ_MSG_BAD_DIST_CODE; return R_ERROR; }
        int xb = ds < 4 ? 0 : (int)(ds >> 1) - 1;
        uint32_t base = ds < 4 ? ds + 1 : 1 + ((2 + (ds & 1)) << xb);
        refill();
        if ((int)avail_bits() < xb) return R_STALL;
        *dist_out = base + peek(xb);
        drop(xb);
        return R_OK;
    }

    // one full symbol (literal, match or EOB) the slow way
    __device__ __noinline__ int slow_symbol()
    {
        uint32_t sym;
        int r = slow_code(S->cnt_l, S->sorted_l, lbits, g_l, &sym);
        if (r != R_OK) { if (r == R_ERROR) msg = SDZ_MSG_BAD_LITLEN_CODE; return r; }
        if (sym < 256) return put_literal(sym);
        if (sym == 256) { eob_len = last_klen; return R_EOB; }
        uint32_t i = sym - 257;
        if (i > 28) { msg = SDZ_MSG_BAD_LITLEN_CODE; return R_ERROR; }
        int xb = i < 8 ? 0 : (i == 28 ? 0 : (int)(i >> 2) - 1);
        uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
        refill();
        if ((int)avail_bits() < xb) return R_STALL;
        uint32_t len = base + peek(xb);
        drop(xb);
        uint32_t dist;
        r = slow_dist(&dist);
        if (r != R_OK) return r;
        return copy_match(len, dist);
    }

    // ------------------------------------------------------------------ table construction
    // Reference acceptance test for one code-length set (src/inftree.ts:131-178,:298) and the
    // canonical structures (counts, sorted symbols).  Returns 0 ok, 1 oversubscribed,
    // 2 incomplete, 3 empty.  *l_out = reference root width, *g_out = longest code.
    __device__ int classify(const uint8_t* lens, int n, int want_bits, uint32_t* cnt, int* l_out, int* g_out, int* pad_out)
    {
        for (int i = glane; i < 16; i += G) cnt[i] = 0;
        __syncwarp(gmask);
        for (int i = glane; i < n; i += G) atomicAdd(&cnt[lens[i]], 1u);
        __syncwarp(gmask);
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

    // sorted symbols + first codes (aux[16..31]) + offsets (aux[0..15])
    __device__ void canonical(const uint8_t* lens, int n, const uint32_t* cnt, uint16_t* sorted)
    {
        if (glane == 0) {
            uint32_t off = 0, code = 0;
            for (int k = 1; k <= 15; k++) {
                S->aux[k] = off;
                S->aux[16 + k] = code;
                off += cnt[k];
                code = (code + cnt[k]) << 1;
            }
            for (int s = 0; s < n; s++) {
                uint32_t k = lens[s];
                if (k) { uint32_t o = S->aux[k]; sorted[o] = (uint16_t)(s | (k << 12)); S->aux[k] = o + 1; }
            }
        }
        __syncwarp(gmask);
    }

    // fill one root LUT from the sorted list; KIND 0 = literal/length, 1 = distance
    template <int KIND, int R>
    __device__ void fill_lut(uint16_t* lut, const uint16_t* sorted, const uint32_t* cnt, int ncodes)
    {
        uint32_t* lut32 = reinterpret_cast<uint32_t*>(lut);
        for (int i = glane; i < (1 << R) / 2; i += G) lut32[i] = E_INVALID | (E_INVALID << 16);
        __syncwarp(gmask);
        for (int k = glane; k < ncodes; k += G) {
            uint32_t e = sorted[k];
            uint32_t sym = e & 0xfff, len = e >> 12;
            // aux[len] now points one past the last symbol of this length
            uint32_t idx = (uint32_t)k - (S->aux[len] - cnt[len]);
            uint32_t code = S->aux[16 + len] + idx;
            uint32_t rev = __brev(code) >> (32 - len);
            if (len > (uint32_t)R) { lut[rev & ((1u << R) - 1u)] = (uint16_t)E_LONG; continue; }
            uint32_t entry;
            if (KIND == 0) {
                if (sym < 256) entry = sym;
                else if (sym == 256) entry = 0x100;
                else {
                    uint32_t i = sym - 257;
                    if (i > 28) continue;                               // 286/287: invalid (fixed block only)
                    uint32_t xb = i < 8 ? 0 : (i == 28 ? 0 : (i >> 2) - 1);
                    uint32_t base = i < 8 ? 3 + i : (i == 28 ? 258 : 3 + ((4 + (i & 3)) << xb));
                    entry = 0x800 | (xb << 8) | (base - 3);
                }
            } else {
                if (sym > 29) continue;                                 // 30/31: invalid (fixed block only)
                uint32_t xb = sym < 4 ? 0 : (sym >> 1) - 1;
                uint32_t m = sym < 4 ? sym : 2 + (sym & 1);
                entry = (xb << 8) | m;
            }
            entry |= len << 12;
            for (uint32_t j = rev; j < (1u << R); j += (1u << len)) lut[j] = (uint16_t)entry;
        }
        __syncwarp(gmask);
    }

    // lit/len + distance tables for lens[0..nl) and lens[nl..nl+nd); reference checks and
    // messages of inflate_trees_dynamic (src/inftree.ts:333-379).  fixed: skip the checks.
    __device__ int build_tables(int nl, int nd, bool fixed)
    {
        int pad_l = 0, pad_d = 0;
        int st = classify(S->lens, nl, 9, S->cnt_l, &lbits, &g_l, &pad_l);
        int used = 0;
        if (!fixed) {
            // the lit/len and distance tables share an arena of MANY = 1400 entries; running out
            // of it is reported as DATA_ERROR, i.e. with the "oversubscribed" text
            if (st == 1) { msg = SDZ_MSG_OVERSUB_LITLEN_TREE; return R_ERROR; }
            if (st != 3) used = ref_table_total(S->cnt_l, g_l, pad_l, lbits);
            if (used > 1400) { msg = SDZ_MSG_OVERSUB_LITLEN_TREE; return R_ERROR; }
            if (st == 2 || st == 3) { msg = SDZ_MSG_INCOMPLETE_LITLEN_TREE; return R_ERROR; }
        }
        st = classify(S->lens + nl, nd, fixed ? 5 : 6, S->cnt_d, &dbits, &g_d, &pad_d);
        if (!fixed) {
            if (st == 1) { msg = SDZ_MSG_OVERSUB_DIST_TREE; return R_ERROR; }
            if (st != 3 && used + ref_table_total(S->cnt_d, g_d, pad_d, dbits) > 1400) { msg = SDZ_MSG_OVERSUB_DIST_TREE; return R_ERROR; }
            if (st == 2) { msg = SDZ_MSG_INCOMPLETE_DIST_TREE; return R_ERROR; }
            if (st == 3 && nl > 257) { msg = SDZ_MSG_EMPTY_DIST_TREE; return R_ERROR; }
        }
        int ncl = nl - (int)S->cnt_l[0], ncd = nd - (int)S->cnt_d[0];
        canonical(S->lens, nl, S->cnt_l, S->sorted_l);
        fill_lut<0, RL>(S->lut_l, S->sorted_l, S->cnt_l, ncl);
        canonical(S->lens + nl, nd, S->cnt_d, S->sorted_d);
        fill_lut<1, RD>(S->lut_d, S->sorted_d, S->cnt_d, ncd);
        return R_OK;
    }

    // dynamic block header: HLIT/HDIST/HCLEN, code-length code, RLE-coded lengths
    // (src/infblocks.ts:334-523).  Uses lut_l as scratch for the 7-bit code-length LUT.
    __device__ int dynamic_header(int* nl_out, int* nd_out)
    {
        if (!ensure(14)) { stall_kind = ST_OTHER; return R_STALL; }
        uint32_t t = peek(14);
        if ((t & 0x1f) > 29 || ((t >> 5) & 0x1f) > 29) { msg = SDZ_MSG_TOO_MANY_SYMS; return R_ERROR; }
        drop(14);
        int nl = 257 + (int)(t & 0x1f), nd = 1 + (int)((t >> 5) & 0x1f), ncl = 4 + (int)(t >> 10);
        int total = nl + nd;
        uint8_t* cl = S->lens + 300;                    // 19 code-length-code lengths (tail of lens[])
        __syncwarp(gmask);
        for (int i = glane; i < 19; i += G) cl[i] = 0;
        __syncwarp(gmask);
        for (int i = 0; i < ncl; i++) {
            if (!ensure(3)) { stall_kind = ST_DYNHDR; return R_STALL; }
            if (glane == 0) cl[c_border[i]] = (uint8_t)peek(3);
            drop(3);
        }
        __syncwarp(gmask);
        // code-length tree: inflate_trees_bits (src/inftree.ts:313-331), requested root 7
        int bb_bits, g_b;
        uint32_t* cnt = S->aux;                         // counts for the 19-symbol set
        int st;
        {
            for (int i = glane; i < 16; i += G) cnt[i] = 0;
            __syncwarp(gmask);
            if (glane == 0) for (int i = 0; i < 19; i++) cnt[cl[i]]++;
            __syncwarp(gmask);
            if (cnt[0] == 19) { msg = SDZ_MSG_INCOMPLETE_BITS_TREE; return R_ERROR; }
            int j = 1;
            while (cnt[j] == 0) j++;
            int g = 7;
            while (cnt[g] == 0) g--;
            int l = 7;
            if (l > g) l = g;
            bb_bits = l; g_b = g;
            int y = 1 << j;
            st = 0;
            for (; j < g; j++, y <<= 1) { y -= (int)cnt[j]; if (y < 0) { st = 1; break; } }
            if (!st) { y -= (int)cnt[g]; if (y < 0) st = 1; else if (y != 0 && g != 1) st = 2; }
            if (st == 1) { msg = SDZ_MSG_OVERSUB_BITS_TREE; return R_ERROR; }
            if (st == 2) { msg = SDZ_MSG_INCOMPLETE_BITS_TREE; return R_ERROR; }
        }
        // 2^bb_bits-entry LUT: sym | len << 5.  A lone 1-bit code answers both patterns (Q11).
        uint8_t* blut = reinterpret_cast<uint8_t*>(S->lut_l);
        if (glane == 0) {
            uint32_t code = 0;
            for (int k = 1; k <= g_b; k++) {
                for (int s = 0; s < 19; s++) {
                    if (cl[s] != k) continue;
                    uint32_t rev = __brev(code) >> (32 - k);
                    for (uint32_t j = rev; j < (1u << bb_bits); j += (1u << k)) blut[j] = (uint8_t)(s | (k << 5));
                    code++;
                }
                code <<= 1;
            }
            if (g_b == 1 && cnt[1] == 1) blut[1] = blut[0];
        }
        __syncwarp(gmask);
        int index = 0;
        uint32_t prev = 0;
        while (index < total) {
            if (!ensure(bb_bits)) { stall_kind = ST_DYNHDR; return R_STALL; }
            uint32_t e = blut[peek(bb_bits)];
            int tbits = (int)(e >> 5), c = (int)(e & 31);
            if (c < 16) {
                drop(tbits);
                if (glane == 0) S->lens[index] = (uint8_t)c;
                prev = (uint32_t)c;
                index++;
            } else {
                int i = c == 18 ? 7 : c - 14;
                int j = c == 18 ? 11 : 3;
                if (!ensure(tbits + i)) { stall_kind = ST_DYNHDR; return R_STALL; }
                drop(tbits);
                j += (int)peek(i);
                drop(i);
                if (index + j > total || (c == 16 && index < 1)) { msg = SDZ_MSG_BAD_REPEAT; return R_ERROR; }
                uint8_t v = c == 16 ? (uint8_t)prev : (uint8_t)0;
                prev = v;
                if (glane == 0) for (int q = 0; q < j; q++) S->lens[index + q] = v;
                index += j;
            }
        }
        __syncwarp(gmask);
        *nl_out = nl; *nd_out = nd;
        return R_OK;
    }

    // ------------------------------------------------------------------ symbol loop
    __device__ int decode_codes()
    {
        for (;;) {
            refill();
            uint32_t e = S->lut_l[(uint32_t)bb & ((1u << RL) - 1u)];
            uint32_t n = e >> 12;
            if (wp + 3 > end_wp || n == 0) {           // stream tail or long/invalid code
                int r = slow_symbol();
                if (r == R_OK) continue;
                return r;
            }
            bb >>= n; bc -= (int)n;
            uint32_t p = e & 0xfff;
            if (p < 256) {
                if (pos >= cap) return R_OUTFULL;
                if (STORE) { if (glane == 0) out[pos] = (uint8_t)p; }
                pos++;
                continue;
            }
            if (p == 256) { eob_len = (int)n; return R_EOB; }
            uint32_t xb = (p >> 8) & 7;
            uint32_t len = 3 + (p & 0xff) + ((uint32_t)bb & ((1u << xb) - 1u));
            bb >>= xb; bc -= (int)xb;
            refill();
            uint32_t de = S->lut_d[(uint32_t)bb & ((1u << RD) - 1u)];
            uint32_t dn = de >> 12;
            uint32_t dist;
            if (dn == 0) {
                int r = slow_dist(&dist);
                if (r != R_OK) return r;
            } else {
                bb >>= dn; bc -= (int)dn;
                uint32_t dx = (de >> 8) & 15;
                dist = 1 + ((de & 3) << dx) + ((uint32_t)bb & ((1u << dx) - 1u));
                bb >>= dx; bc -= (int)dx;
            }
            int r = copy_match(len, dist);
            if (r != R_OK) return r;
        }
    }

    // stored block body (src/infblocks.ts:278-333) with the Q2 truncation
    __device__ int stored_block(uint32_t left)
    {
        uint32_t start = byte_pos();
        uint32_t n_in = in_len - start;
        uint32_t copied = 0;
        int r = R_OK;
        while (left) {
            if (n_in == 0) { stall_kind = ST_OTHER; r = R_STALL; break; }
            if (ring.room() == 0) {
                int returns = ring.make_room();
                if (returns) { left = 0; break; }      // `left` is a local of proc(): lost on return
            }
            uint32_t t = min(min(left, n_in), (uint32_t)ring.room());
            ring.q += (int)t;
            copied += t; n_in -= t; left -= t;
        }
        if (copied > cap - pos) return R_OUTFULL;
        if (STORE) {
            const uint8_t* src = gsrc + start;
            uint8_t* dst = out + pos;
            for (uint32_t i = glane; i < copied; i += G) dst[i] = src[i];
        }
        pos += copied;
        seek(start + copied);
        return r;
    }

    // all deflate blocks; returns R_EOB when the final block completed
    __device__ int blocks(uint32_t* n_blocks)
    {
        for (;;) {
            if (!ensure(3)) { stall_kind = ST_OTHER; return R_STALL; }
            uint32_t t = peek(3);
            drop(3);
            int last = (int)(t & 1);
            (*n_blocks)++;
            uint32_t start_pos = pos;
            int r;
            switch (t >> 1) {
            case 0: {
                drop(bc & 7);
                if (!ensure(32)) { stall_kind = ST_OTHER; return R_STALL; }
                uint32_t v = (uint32_t)bb;
                if ((((~v) >> 16) & 0xffff) != (v & 0xffff)) { msg = SDZ_MSG_BAD_STORED_LEN; return R_ERROR; }
                drop(32);
                r = stored_block(v & 0xffff);
                if (r != R_OK) return r;
                if (last) { ring.wash(); return R_EOB; }
                continue;
            }
            case 1: {
                __syncwarp(gmask);
                for (int i = glane; i < 320; i += G) {
                    uint8_t v = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : (i < 288 ? 8 : 5)));
                    S->lens[i] = v;
                }
                __syncwarp(gmask);
                build_tables(288, 30, true);
                lbits = 9; dbits = 5;
                break;
            }
            case 2: {
                int nl, nd;
                r = dynamic_header(&nl, &nd);
                if (r != R_OK) return r;
                r = build_tables(nl, nd, false);
                if (r != R_OK) return r;
                break;
            }
            default:
                msg = SDZ_MSG_BAD_BLOCK_TYPE;
                return R_ERROR;
            }
            r = decode_codes();
            ring.write(pos - start_pos);
            if (r != R_EOB) { if (r == R_STALL) stall_kind = ST_OTHER; return r; }
            // End of block.  When inflate_fast() decodes the EOB its STREAM_END status leaks through
            // WASH's early return (src/infcodes.ts:264,:357,:627-638 -> src/infblocks.ts:560-564), so
            // the block completes after ONE flush attempt; only a slow-path EOB (fewer than 258 bytes
            // of window room or fewer than 10 unread input bytes, src/infcodes.ts:339) washes the
            // window completely, returning to append() as often as needed.
            {
                uint64_t b_before = bit_pos() - (uint64_t)eob_len;
                uint32_t kcur = 4u + ((0u - (uint32_t)b_before - 4u) & 7u);     // reference bit-buffer fill (approx.)
                uint64_t loaded = (b_before + kcur) >> 3;
                bool fast_eob = ring.room() >= 258 && (uint64_t)in_len >= loaded + 10;
                if (fast_eob) ring.flush(); else ring.wash();
            }
            if (last) { ring.wash(); return R_EOB; }                           // DRY (src/infblocks.ts:579-594)
        }
    }
};

// ---------------------------------------------------------------------- per-stream driver
template <int G, bool STORE>
__device__ void inflate_stream(Decoder<G, STORE>& d, const InflateParams& P, unsigned long long idx)
{
    sdz_result R;
    memset(&R, 0, sizeof R);
    const uint32_t in_len = P.in_len[idx];
    const uint8_t mode_raw = P.mode[idx];
    const int mode = mode_raw & 0x7f;
    const bool has_dict = (mode_raw & 0x80) != 0;
    const uint8_t* src = P.in + P.in_off[idx];
    R.out_off = P.out_off ? P.out_off[idx] : 0;

    d.gsrc = src;
    d.in_len = in_len;
    d.end_wp = (in_len + 3) / 4;
    d.tail_mask = (in_len & 3) ? ((1u << ((in_len & 3) * 8)) - 1u) : 0xffffffffu;
    d.total_chunks = (in_len + CH - 1) / CH;
    d.out = STORE ? P.out + R.out_off : nullptr;
    d.pos = 0;
    d.cap = (STORE && P.out_cap) ? P.out_cap[idx] : 0xffffffffu;
    d.msg = SDZ_MSG_NONE;
    d.stall_kind = ST_NONE;
    d.D = 0;
    d.dict_tail = nullptr;
    d.lbits = d.dbits = d.g_l = d.g_d = 0;

    int thrown = SDZ_THROW_NONE;
    int zstatus = SDZ_Z_OK;
    bool done = false;          // Inflate reached Mode.DONE
    bool is_gzip = false;
    int method = 0;
    uint32_t n_blocks = 0;
    int32_t stored = 0, isize = 0, mtime = 0;
    uint32_t name_off = 0, name_len = 0;

    bool raw = mode == SDZ_MODE_RAW;
    bool skip_all = false;
    if (mode == SDZ_MODE_SNIFF) {
        // inflate(): src/sd-inflate.ts:194-207
        if (in_len < 2) { R.thrown_inflate = SDZ_THROW_TOO_SMALL; skip_all = true; }
        else {
            uint32_t b0 = src[0], b1 = src[1];
            bool ident = (b0 == 0x78 && (((b0 << 8) + b1) % 31) == 0) || (b0 == 0x1f && b1 == 0x8b);
            raw = !ident;
        }
    }
    if (!skip_all && raw && has_dict) {                 // RangeError in the constructor (src/sd-inflate.ts:69-71)
        R.thrown_inflate = SDZ_THROW__COUNT;
        skip_all = true;
    }
    if (skip_all || in_len == 0) {
        // append() of an empty chunk returns [] (src/sd-inflate.ts:92-94); finish() on a fresh Inflater
        if (d.glane == 0) P.res[idx] = R;
        return;
    }

    d.ring.init(0);
    d.seek(0);

    int r = R_OK;
    // ---- container header (src/inflate.ts:142-401)
    if (!raw) {
        bool ok = true;                                 // false: input ran out (incomplete)
        #define GETBYTE(v) do { if (!d.ensure(8)) { ok = false; goto hdr_done; } (v) = d.peek(8); d.drop(8); } while (0)
        uint32_t b;
        if (!d.ensure(8)) { ok = false; goto hdr_done; }
        if (d.peek(8) == 0x1f) {
            d.drop(8);
            GETBYTE(b);
            if (b != 0x8b) { d.msg = SDZ_MSG_BAD_GZIP_ID; r = R_ERROR; goto hdr_done; }
            is_gzip = true;
        }
        GETBYTE(b);
        method = (int)b;
        if ((method & 0xf) != 8) { d.msg = SDZ_MSG_BAD_METHOD; r = R_ERROR; goto hdr_done; }
        if ((method >> 4) + 8 > 15) { d.msg = SDZ_MSG_BAD_WINDOW; r = R_ERROR; goto hdr_done; }
        GETBYTE(b);
        if (is_gzip) {
            uint32_t gflags = b;
            for (int i = 0; i < 4; i++) { GETBYTE(b); mtime = (int32_t)(((uint32_t)mtime >> 8) | (b << 24)); }
            GETBYTE(b);                                 // XFL
            GETBYTE(b);                                 // OS
            if (gflags & 4) {                           // FEXTRA never leaves EXTRA0
                ok = false;
                goto hdr_done;
            }
            if (gflags & 8) {
                name_off = d.byte_pos();
                for (;;) { GETBYTE(b); if (b == 0) break; name_len++; }
            }
            if (gflags & 16) { for (;;) { GETBYTE(b); if (b == 0) break; } }
            if (gflags & 2) { GETBYTE(b); GETBYTE(b); }
        } else {
            if ((((uint32_t)method << 8) + b) % 31 != 0) { d.msg = SDZ_MSG_BAD_HEADER_CHECK; r = R_ERROR; goto hdr_done; }
            if (b & 0x20) {
                int32_t dictid = 0;
                for (int i = 0; i < 4; i++) { uint32_t v; GETBYTE(v); dictid = (int32_t)(((uint32_t)dictid << 8) | v); }
                // NEED_DICT -> inflateSetDictionary (src/sd-inflate.ts:116-126, src/inflate.ts:475-503)
                if (!has_dict) { thrown = SDZ_THROW_DICT_REQUIRED; zstatus = SDZ_Z_NEED_DICT; goto finish; }
                if (P.dict_adler[idx] != dictid) { thrown = SDZ_THROW_DICT_INVALID; zstatus = SDZ_Z_NEED_DICT; goto finish; }
                uint32_t dl = P.dict_len[idx];
                uint32_t used = dl >= (uint32_t)WSIZE ? (uint32_t)WSIZE - 1 : dl;
                d.D = (int)used;
                d.dict_tail = P.dict + P.dict_off[idx] + (dl - used);
                d.ring.init((int)used);
            }
        }
    hdr_done:
        #undef GETBYTE
        if (r == R_ERROR) { thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_DATA_ERROR; goto finish; }
        if (!ok) { zstatus = SDZ_Z_OK; goto finish; }   // truncated header: incomplete, no output
    }

    // ---- deflate blocks
    r = d.blocks(&n_blocks);
    if (r == R_ERROR) { thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_DATA_ERROR; goto finish; }
    if (r == R_OUTFULL) { zstatus = SDZ_Z_BUF_ERROR; goto finish; }
    if (r == R_STALL) {
        // input exhausted.  proc() flushes; if that fills the 16 KiB buffer append() calls
        // again, and BTREE/DTREE cannot be re-entered -> STREAM_ERROR is thrown.
        if (d.stall_kind == ST_DYNHDR) {
            d.ring.flush();
            if (d.ring.ao == 0) { thrown = SDZ_THROW_INFLATE_ERROR; zstatus = SDZ_Z_STREAM_ERROR; d.msg = SDZ_MSG_NONE; }
        }
        goto finish;
    }
    // ---- final block done: unused whole bytes go back, partial bits are dropped (src/inflate.ts:409-421)
    d.drop(d.bc & 7);
    if (raw) {
        done = true;
    } else {
        int nbytes = is_gzip ? 8 : 4;
        int i = 0;
        for (; i < nbytes; i++) {
            if (!d.ensure(8)) break;
            uint32_t b = d.peek(8);
            d.drop(8);
            if (is_gzip) {
                if (i < 4) stored = (int32_t)(((uint32_t)stored >> 8) | (b << 24));
                else isize = (int32_t)(((uint32_t)isize >> 8) | (b << 24));
            } else {
                stored = (int32_t)(((uint32_t)stored << 8) | b);
            }
        }
        done = i == nbytes;
    }
    if (done) {
        zstatus = SDZ_Z_STREAM_END;
        if (d.byte_pos() < in_len) thrown = SDZ_THROW_HANG;   // bytes after the end: append() spins
    }

finish:
    d.drain();
    __syncwarp(d.gmask);
    R.out_len = (thrown && STORE) ? 0 : d.pos;      // the sizing pass always reports the decoded size
    R.total_in = d.byte_pos();
    R.zstatus = zstatus;
    R.stored_checksum = stored;
    R.stored_isize = isize;
    R.mtime = mtime;
    R.name_off = name_len ? name_off : 0;
    R.name_len = name_len;
    R.n_blocks = n_blocks;
    R.msg_id = (uint8_t)d.msg;
    R.thrown_append = (uint8_t)thrown;
    R.container = (uint8_t)(is_gzip ? SDZ_GZIP : (method == 0 ? SDZ_RAW : SDZ_ZLIB));
    R.complete = done ? 1 : 0;
    if (d.glane == 0) P.res[idx] = R;
}

template <int G, bool STORE>
__global__ void __launch_bounds__(128) inflate_kernel(InflateParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int gid = threadIdx.x / G;
    GroupSmem* S = reinterpret_cast<GroupSmem*>(smem_raw) + gid;
    (void)groups;

    Decoder<G, STORE> d;
    d.S = S;
    d.glane = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    d.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - d.glane));
    d.base_seq = 0; d.issued = 0; d.waited = 0; d.chunk0 = 0;
    if (d.glane == 0) {
        for (int i = 0; i < NBUF; i++) mbar_init(&S->mbar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    for (;;) {
        unsigned long long idx = 0;
        if (d.glane == 0) idx = atomicAdd(P.counter, 1ull);
        idx = __shfl_sync(d.gmask, idx, 0, G);
        if (idx >= P.n) break;
        inflate_stream<G, STORE>(d, P, idx);
    }
}

}  // namespace sdz
