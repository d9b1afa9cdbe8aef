/*
 * sdzcuda.h - C ABI of libsdzcuda.so: the B200 (sm_100a) engine behind the public API of
 * @stardazed/zlib for ONE path: inflate (raw / zlib / gzip, preset dictionary) and the
 * adler32 / crc32 checksums.  This is the boundary an N-API shim (sd-zlib_b200/ts/), the
 * Python mirror (sd-zlib_b200/host/sdzlib) and the tests bind to.  Plain pointers and
 * sizes only; no C++ or torch types; no exceptions; one sdz_ctx per host thread.
 *
 * What each entry point replaces in the reference (paths relative to the reference tree):
 *
 *   sdz_adler32 / sdz_adler32_chain   adler32(source, seed=1)            src/adler32.ts:17-105
 *   sdz_crc32   / sdz_crc32_chain     crc32(source, seed=0)              src/crc32.ts:17-106
 *   sdz_inflate_batch                 new Inflater(opts).append(buf) + finish() per buffer,
 *                                     and inflate(buf, dict)             src/sd-inflate.ts:54-228
 *                                     (driving src/inflate.ts:132-503, src/infblocks.ts:123-633,
 *                                      src/inftree.ts:95-392, src/infcodes.ts:62-676)
 *   sdz_inflate_sizes                 (new) sizing pass for inflateBatch()'s output arena
 *   sdz_inflate_batch_device          same as sdz_inflate_batch with every buffer already in HBM
 *
 * Results are bit-exact with the reference, including its documented quirks (SURVEY
 * Appendix A); values that the reference holds as signed int32 JS numbers (checksums,
 * ISIZE, MTIME) are int32_t here.  There is no CPU fallback: every entry point fails with
 * SDZ_E_CUDA / SDZ_E_NO_DEVICE when the device is unusable.
 */
#ifndef SDZCUDA_H
#define SDZCUDA_H

#include <stddef.h>
#include <stdint.h>
#include "sdz_codes.h"

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (negative); 0 = success */
enum sdz_err {
    SDZ_OK = 0,
    SDZ_E_NO_DEVICE = -1,     /* no CUDA device / wrong architecture                         */
    SDZ_E_CUDA = -2,          /* a CUDA call failed; sdz_last_error() has the text           */
    SDZ_E_ARG = -3,           /* bad argument (NULL, misaligned device offset, >= 4 GiB ...) */
    SDZ_E_NOMEM = -4,         /* host or device allocation failed                            */
    SDZ_E_OUT_CAP = -5,       /* output arena / slot too small for at least one stream       */
    SDZ_E_UNSUPPORTED = -6    /* flag or mode not implemented                                */
};

/* sdz_inflate_batch flags */
#define SDZ_PARITY_REFERENCE 0u   /* default: identical to @stardazed/zlib, quirks included  */
#define SDZ_PARITY_SPEC      1u   /* RFC 1950/1951/1952 as zlib 1.3 implements them: stored blocks of any size (no Q2), FEXTRA
                                     parsed (Q5), distances before the start of the output rejected (Q6), zlib's Huffman acceptance
                                     rules and texts (Q9; no MANY limit, Q10), standard Adler-32 (Q1), unsigned ISIZE (Q13),
                                     32,768-byte dictionaries (Q14), symbols decoded as soon as their own bits are there (Q15),
                                     bytes after the end ignored (Q4).  sdz_inflate_batch / _sizes / _batch_device only;
                                     streams go through the general decoder (the two-phase fast path is reference-exact only) */

typedef struct sdz_ctx sdz_ctx;

/* Context = one CUDA device + its streams, staging buffers and constant tables. */
int  sdz_ctx_create(int device, uint32_t flags, sdz_ctx** out);
/* Multi-device context (SURVEY 8b `sdz_ctx_create(devices[], ndev)`, 8e): one child context per entry of `devices`
 * (an index may repeat: several independent pipelines on one GPU, which is also how the partition logic is tested on a
 * one-GPU box).  sdz_inflate_batch / sdz_inflate_sizes on such a context partition the batch PER STREAM: device d
 * decodes the contiguous range of streams whose compressed bytes come closest to an equal share (greedy sweep over the
 * prefix sum), one host thread and one copy/compute pipeline per device, records and bytes written in caller order;
 * nothing crosses between devices.  This is what `inflateBatch(buffers[])` (the one entry point added to
 * dist/sd-zlib.d.ts:43-67's API) sits on when a box has several B200s.  Every other entry point of a multi-device
 * context runs on its first device. */
int  sdz_ctx_create_multi(const int* devices, int ndev, uint32_t flags, sdz_ctx** out);
int  sdz_ctx_device_count(sdz_ctx* ctx);
/* partition of the most recent batch of a multi-device context: device d took streams [cut[d], cut[d + 1]);
 * n_cut >= device count + 1 */
int  sdz_last_partition(sdz_ctx* ctx, uint64_t* cut, int n_cut);
void sdz_ctx_destroy(sdz_ctx* ctx);
const char* sdz_last_error(sdz_ctx* ctx);          /* text of the last failure on this ctx   */
const char* sdz_version(void);
/* number of kernel launches issued through this ctx so far (bench.py's gpu_launches) */
uint64_t sdz_launch_count(sdz_ctx* ctx);
/* device time of the most recent call's kernels in milliseconds (CUDA events on the ctx
 * stream), split by phase: [0] inflate kernel, [1] checksum kernel, [2] whole device phase */
int sdz_last_timing(sdz_ctx* ctx, float ms[3]);
/* the same for the most recent batched inflate, by kernel: [0] phase A (Huffman decode -> tokens; the earlier chunks'
 * phase B runs next to it), [2] the general decoder over the streams phase A handed over (queued right behind the last
 * phase A, running next to the last phase B; the whole batch when the fast path is off), [1] what phase B (tokens ->
 * bytes) still needs after that, [3] finalize (checksums + records), [4] all of it.  [0] + [2] + [1] + [3] = [4]. */
int sdz_last_phase_timing(sdz_ctx* ctx, float ms[5]);
/* most recent fast-path launch (the last sub-batch of a pipelined call): out[0] = streams finished by the two-phase
 * path, out[1] = streams it handed to the general decoder */
int sdz_last_fast_stats(sdz_ctx* ctx, uint64_t out[2]);
/* White-box TEST entry (not bound by the N-API shim): for n code-length sets (320 bytes each: nl[i] literal/length lengths
 * followed by nd[i] distance lengths) the acceptance class (0 ok, 1 oversubscribed, 2 incomplete, 3 empty) and the number of
 * table entries the reference's huft_build allocates (src/inftree.ts:217-246) as the kernels re-derive it in closed form -
 * out[4 i .. 4 i + 3] = { class_lit, entries_lit, class_dist, entries_dist }.  group = 4 (general decoder) or 32 (fast path). */
int sdz_debug_table_totals(sdz_ctx* ctx, const uint8_t* lens, const int32_t* nl, const int32_t* nd, uint64_t n, int group, int32_t* out);

/* Pinned host memory helpers (so that callers can hand over DMA-able buffers). */
void* sdz_host_alloc(size_t bytes);
void  sdz_host_free(void* p);
/* Pinned host memory on the NUMA node next to ctx's device (mmap + mbind + cudaHostRegister, portable across devices);
 * the library's own staging buffers are allocated this way.  sdz_ctx_numa_node: that node, or -1 when unknown. */
void* sdz_host_alloc_near(sdz_ctx* ctx, size_t bytes);
int   sdz_ctx_numa_node(sdz_ctx* ctx);
/* Device memory helpers for callers that keep data resident in HBM. */
void* sdz_device_alloc(sdz_ctx* ctx, size_t bytes);
void  sdz_device_free(sdz_ctx* ctx, void* p);
int   sdz_memcpy_h2d(sdz_ctx* ctx, void* dst, const void* src, size_t bytes);
int   sdz_memcpy_d2h(sdz_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---------------------------------------------------------------- checksums
 * `on_device` != 0: `p` is a device pointer on ctx's device (data already in HBM).
 * n must be < 4 GiB per call (the reference's crc32 is undefined beyond that).
 * The value domain is the reference's: signed int32, seed signed or unsigned bit pattern. */
int sdz_adler32(sdz_ctx* ctx, const uint8_t* p, uint64_t n, int32_t seed, int on_device, int32_t* out);
int sdz_crc32(sdz_ctx* ctx, const uint8_t* p, uint64_t n, int32_t seed, int on_device, int32_t* out);

/* Seed chaining over consecutive segments of one buffer, evaluated on the device without a
 * host round trip per segment:  v[0] = f(seg 0, seed), v[i] = f(seg i, v[i-1]).
 * Exactly what a caller gets from `s = adler32(chunk_i, s)` in a loop (including Q1 for
 * segments whose length is a non-zero multiple of 5552).  out_values may be NULL except
 * for the last element, which is always written to *out_last. */
int sdz_adler32_chain(sdz_ctx* ctx, const uint8_t* p, const uint64_t* seg_len, uint64_t n_seg,
                      int32_t seed, int on_device, int32_t* out_values, int32_t* out_last);
int sdz_crc32_chain(sdz_ctx* ctx, const uint8_t* p, const uint64_t* seg_len, uint64_t n_seg,
                    int32_t seed, int on_device, int32_t* out_values, int32_t* out_last);

/* Checksums of n independent buffers in one launch (one warp per buffer): out[i] =
 * adler32(buf_i, seed_i) when kind[i] == 0, crc32(buf_i, seed_i) when kind[i] == 1 - what
 * Deflater.append() computes over its source data (src/sd-deflate.ts:185-190) and what the gzip /
 * zlib trailers need.  seeds may be NULL (reference defaults: 1 for adler32, 0 for crc32).  Each
 * call to the reference is ONE call here, so Q1 applies per buffer. */
int sdz_checksum_batch(sdz_ctx* ctx, const uint8_t* const* bufs, const uint64_t* lens, const uint8_t* kind,
                       const int32_t* seeds, uint64_t n, int32_t* out);

/* ---------------------------------------------------------------- Deflater support
 * The compressor stays on the CPU; the device supplies what `Deflater` computes over its SOURCE data - checksum =
 * adler32(chunk, 1) for "deflate", crc32(chunk, 0) for "gzip" (src/sd-deflate.ts:185-190), one launch for the batch - and
 * writes the reference's containers around raw deflate payloads: zlib 78 01 (78 20 + DICTID when dict_adler != 0,
 * src/sd-deflate.ts:98-116), gzip 1f 8b 08 FLG MTIME 00 ff [FNAME] (:118-152), trailers Adler-32 BE / CRC-32 + ISIZE LE
 * (:154-165).  One source buffer == one Deflater.append(), so Q1 applies per buffer exactly as in the reference. */
enum sdz_wrap_format { SDZ_WRAP_RAW = 0, SDZ_WRAP_DEFLATE = 1, SDZ_WRAP_GZIP = 2 };
typedef struct sdz_wrap_in {
    const uint8_t* payload;  uint64_t payload_len;   /* raw deflate data produced by the compressor        */
    const uint8_t* source;   uint64_t source_len;    /* the data that was compressed                       */
    const char*    file_name;                        /* gzip FNAME (Latin-1, NUL-terminated) or NULL       */
    uint32_t       mtime;                            /* gzip MTIME (the reference uses Date.now() / 1000)   */
    int32_t        dict_adler;                       /* adler32(dictionary) or 0 (src/sd-deflate.ts:88)     */
    uint8_t        format;                           /* enum sdz_wrap_format                               */
    uint8_t        reserved[7];
} sdz_wrap_in;
int sdz_deflate_wrap_sizes(const sdz_wrap_in* in, uint64_t n, uint64_t* out_len);
int sdz_deflate_wrap_batch(sdz_ctx* ctx, const sdz_wrap_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off,
                           uint64_t* out_len);

/* ---------------------------------------------------------------- batched inflate */

/* one input buffer == one `new Inflater(options)` fed with a single append() */
typedef struct sdz_in {
    const uint8_t* data;     /* borrowed for the duration of the call                         */
    uint64_t len;
    const uint8_t* dict;     /* options.dictionary, or NULL                                   */
    uint32_t dict_len;
    uint8_t  mode;           /* enum sdz_mode                                                 */
    uint8_t  reserved[3];
} sdz_in;

/* Decode n independent buffers.  Stream i writes its bytes to
 * out_arena[out_off[i] .. out_off[i] + out_cap[i]) (host memory; pinned memory from
 * sdz_host_alloc avoids a staging copy).  results[i] receives the finish() record; a bad
 * stream never aborts the batch.  Returns SDZ_E_OUT_CAP if some stream needed more than
 * its slot (that stream's record has out_len = bytes that fit, complete = 0 and
 * zstatus = SDZ_Z_BUF_ERROR; all others are valid). */
int sdz_inflate_batch(sdz_ctx* ctx, const sdz_in* in, uint64_t n,
                      uint8_t* out_arena, const uint64_t* out_off, const uint64_t* out_cap,
                      sdz_result* results, uint32_t flags);

/* Sizing pass: out_len[i] = number of bytes stream i decodes to (same decode walk, no
 * stores).  Lets inflateBatch() allocate its output arena exactly. */
int sdz_inflate_sizes(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint64_t* out_len, uint32_t flags);

/* Device-resident form: everything below lives in HBM on ctx's device.
 *   d_in      compressed arena; stream i = d_in[in_off[i] .. in_off[i] + in_len[i]); every
 *             in_off[i] must be a multiple of 16 and the arena must be readable for
 *             SDZ_IN_PAD bytes past its last stream (TMA bulk staging reads whole chunks)
 *   d_dict    dictionary arena (may be NULL); dict_off/dict_len per stream (len 0 = none)
 *   dict_adler  reference adler32 (src/adler32.ts, incl. Q1) of each WHOLE dictionary
 *   d_out     output arena; stream i owns d_out[out_off[i] .. out_off[i] + out_cap[i])
 *   d_results n records
 * Host arrays: none - every array argument is a device pointer.  Asynchronous on the ctx
 * stream unless `sync` != 0. */
#define SDZ_IN_PAD 1024
typedef struct sdz_batch_dev {
    const uint8_t*  d_in;
    const uint64_t* d_in_off;
    const uint32_t* d_in_len;
    const uint8_t*  d_mode;        /* enum sdz_mode per stream                              */
    const uint8_t*  d_dict;        /* may be NULL                                           */
    const uint64_t* d_dict_off;    /* may be NULL when d_dict is NULL                       */
    const uint32_t* d_dict_len;
    const int32_t*  d_dict_adler;
    uint8_t*        d_out;         /* may be NULL for a sizing pass                         */
    const uint64_t* d_out_off;
    const uint32_t* d_out_cap;
    sdz_result*     d_results;
    uint64_t        n;
} sdz_batch_dev;
int sdz_inflate_batch_device(sdz_ctx* ctx, const sdz_batch_dev* batch, uint32_t flags, int sync);

/* ---------------------------------------------------------------- streaming: class Inflater over several append() calls
 * One sdz_inflater == one `new Inflater({raw, dictionary})` (src/sd-inflate.ts:54-80).  The state the reference carries
 * between calls (src/inflate.ts:79-95, src/infblocks.ts:40-50, src/infcodes.ts:44-60: mode, bit buffer, window, tables) is
 * kept as an sdz_resume record + the decoded bytes in HBM, so an append() decodes only what is new.  Behaviour at the
 * boundaries is the reference's, defects included: input that ends inside a dynamic block header makes the NEXT append()
 * throw "inflate error: " (BTREE / DTREE are not re-entrant), a stored block ends where its append() ended
 * (`left` is a local of proc(), Q2), bytes after the end of the stream make append() spin when they arrive in the same
 * call as the end (Q4: SDZ_THROW_HANG) and throw "inflate error: bad input data" when they arrive in a later one.
 *   sdz_inflater_append   = Inflater.append(data) (src/sd-inflate.ts:87-153): *new_bytes = bytes this call produced (the
 *                           caller cuts them into <= 16 KiB chunks, exactly the shapes the reference returns, because
 *                           every append() starts with an empty 16 KiB buffer); res->thrown_append != 0: the call throws
 *                           and its output is lost
 *   sdz_inflater_read     copies the bytes of the most recent append() to host memory
 *   sdz_inflater_finish   = Inflater.finish() (src/sd-inflate.ts:159-179): the record as of now */
typedef struct sdz_inflater sdz_inflater;
int  sdz_inflater_create(sdz_ctx* ctx, int raw, const uint8_t* dict, uint32_t dict_len, sdz_inflater** out);
int  sdz_inflater_append(sdz_inflater* s, const uint8_t* data, uint64_t len, uint64_t* new_bytes, sdz_result* res);
int  sdz_inflater_read(sdz_inflater* s, uint8_t* dst, uint64_t cap);
int  sdz_inflater_finish(sdz_inflater* s, sdz_result* res);
/* bytes [off, off + len) of the input received so far (e.g. the gzip FNAME at res->name_off) */
int  sdz_inflater_input(sdz_inflater* s, uint64_t off, uint64_t len, uint8_t* dst);
void sdz_inflater_destroy(sdz_inflater* s);

/* ONE large stream (BASELINE config 5), decoded by all SMs: pass 1 indexes the deflate block boundaries
 * (speculative header search + per-block extents, chained from the first block), pass 2 decodes every
 * block in parallel with 16-bit marker symbols for back-references into the unknown 32 KiB window,
 * propagates the windows in stream order and resolves the markers.  Same record as one
 * `new Inflater(opts).append(data); finish()` (src/sd-inflate.ts:54-180).  on_device != 0: `data` and
 * `out` are device pointers (data 16-byte aligned and readable for SDZ_IN_PAD bytes past its end, out
 * with 64 bytes of slack).  Streams the index cannot handle (preset dictionary, FEXTRA, a stored block the
 * reference cuts short, anything that is not a complete well-formed stream) are handed to the
 * ordinary one-group decoder, which is exact but slow.
 * Returns SDZ_E_OUT_CAP (res->out_len = needed size) if out_cap is too small. */
int sdz_inflate_large(sdz_ctx* ctx, const uint8_t* data, uint64_t len, uint8_t mode, int on_device,
                      uint8_t* out, uint64_t out_cap, sdz_result* res);

/* The same path as phases of a session, for a stream spread over several GPUs (one process per GPU, every
 * rank holding the whole compressed stream).  Rank r of n:
 *   sdz_large_open; sdz_large_index(r, n) -> all-gather the block / resume-point records -> sdz_large_plan on
 *   every rank (identical result); sdz_large_range(r, n) = its slice of the output; sdz_large_decode(r, n, d_out);
 *   receive the final 32 KiB before its slice from rank r-1 into d_out - 32768 (rank 0: nothing);
 *   sdz_large_windows; send the last 32 KiB of its slice to rank r+1; sdz_large_resolve; sdz_crc32 of the slice;
 *   sdz_crc32_combine in rank order; sdz_large_finish with the combined value.
 * sdz_inflate_large is exactly this sequence for r = 0, n = 1.  SDZ_E_UNSUPPORTED from any phase = the stream needs
 * the sequential decoder (sdz_inflate_batch). */
typedef struct sdz_large sdz_large;
typedef struct sdz_large_block {      /* one candidate deflate block                                          */
    uint64_t bit, end_bit;            /* bit position of its header / of the first bit after it               */
    uint64_t out_len;                 /* bytes it decodes to                                                  */
    uint8_t  last, btype, ok;         /* BFINAL, BTYPE, 1 = walked to its end-of-block code                   */
    uint8_t  reserved[5];
} sdz_large_block;
typedef struct sdz_large_ckpt {       /* resume point inside a block: the symbol at `bit` produces byte `pos` */
    uint64_t block_bit, bit;
    uint32_t pos, reserved;
} sdz_large_ckpt;
int  sdz_large_open(sdz_ctx* ctx, const uint8_t* data, uint64_t len, uint8_t mode, int on_device, sdz_large** out);
void sdz_large_close(sdz_large* L);
int  sdz_large_index(sdz_large* L, uint32_t part, uint32_t n_parts, const sdz_large_block** blocks, uint64_t* n_blocks,
                     const sdz_large_ckpt** ckpts, uint64_t* n_ckpts);      /* arrays owned by the session */
int  sdz_large_plan(sdz_large* L, const sdz_large_block* blocks, uint64_t n_blocks, const sdz_large_ckpt* ckpts, uint64_t n_ckpts,
                    uint64_t* total_out, uint64_t* n_pieces);
int  sdz_large_range(sdz_large* L, uint32_t part, uint32_t n_parts, uint64_t* off_lo, uint64_t* off_hi);
int  sdz_large_decode(sdz_large* L, uint32_t part, uint32_t n_parts, uint8_t* d_out);
int  sdz_large_windows(sdz_large* L);
int  sdz_large_resolve(sdz_large* L);
int  sdz_large_finish(sdz_large* L, int32_t running_checksum, sdz_result* res);
int  sdz_large_is_gzip(sdz_large* L);
/* crc32(A || B) from crc32(A), crc32(B), len(B) - host arithmetic, no device needed */
int32_t sdz_crc32_combine(int32_t crc_a, int32_t crc_b, uint64_t len_b);
/* adler32(A || B) from adler32(A), adler32(B) and len(B) with the STANDARD arithmetic (B seeded with 1) - what the sharded
 * running checksum of a zlib / raw stream is joined with; the reference's defect can only touch the final
 * <= 16 KiB chunk, which is checksummed as one sdz_adler32 call seeded with the combined value (sdzlib/large.py) */
int32_t sdz_adler32_combine(int32_t adler_a, int32_t adler_b, uint64_t len_b);
int sdz_sync(sdz_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SDZCUDA_H */
