/*
 * sdz_oracle.h - CPU oracle for the inflate + adler32/crc32 hot path of @stardazed/zlib.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sd-zlib_b200/ may include, link or call this.
 * It is a plain-C restatement of the reference's TypeScript state machine (citations in
 * sdz_oracle.c) and is used as the checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED for valid streams by the reference's own fixtures (test/ .deflate,
 * .gz, .raw -> .txt files) and cross-checked against system zlib 1.3;
 * the quirk behaviours are pinned by code reading only,
 * because the reference (TypeScript, no JS runtime in this image) cannot be executed here.
 */
#ifndef SDZ_ORACLE_H
#define SDZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/sdz_codes.h"

#ifdef __cplusplus
extern "C" {
#endif

/* adler32(source, seed = 1), src/adler32.ts:17-105 (including the NMAX-multiple defect) */
int32_t sdzo_adler32(const uint8_t* buf, uint64_t len, int32_t seed);
/* crc32(source, seed = 0), src/crc32.ts:17-106 */
int32_t sdzo_crc32(const uint8_t* buf, uint64_t len, int32_t seed);

/* class Inflater, src/sd-inflate.ts:54-180 */
typedef struct sdzo_inflater sdzo_inflater;

/* new Inflater({raw, dictionary}); dict may be NULL.  The dictionary bytes are copied.
 * Option-type errors (TypeError/RangeError, src/sd-inflate.ts:62-76) belong to the host
 * language binding and are not modelled here; raw && dict returns NULL (RangeError). */
sdzo_inflater* sdzo_inflater_new(int raw, const uint8_t* dict, size_t dict_len);
void sdzo_inflater_free(sdzo_inflater*);

/* Output of one append() call: the Uint8Array[] it returns. */
typedef struct sdzo_chunks {
    uint8_t* data;      /* all chunks of this call, concatenated (malloc'd, grows)  */
    size_t len, cap;
    uint32_t* chunk_len; /* length of each chunk (<= 16384)                         */
    size_t n_chunks, chunk_cap;
} sdzo_chunks;
void sdzo_chunks_free(sdzo_chunks*);

/* Inflater.append(data).  Returns enum sdz_thrown (0 = returned normally).  When it
 * throws, `out` is reset to empty (the JS caller never sees the call's output). */
int sdzo_append(sdzo_inflater*, const uint8_t* data, size_t len, sdzo_chunks* out);

/* Inflater.finish(): fills the record (container, complete, checksum/fileSize state,
 * success, stored/running values, mtime, name offsets are relative to the FIRST append). */
void sdzo_finish(sdzo_inflater*, sdz_result* res);
/* gzip FNAME as collected byte by byte (src/inflate.ts:387); returns length */
size_t sdzo_file_name(sdzo_inflater*, const uint8_t** bytes);

/* One-shot: new Inflater(opts) + append(data) + finish(), and for mode SNIFF the
 * inflate() wrapper logic on top (src/sd-inflate.ts:189-228).  `out`/`out_cap` receive
 * the concatenated output; returns 0, or -1 if out_cap is too small (res->out_len holds
 * the needed size). */
int sdzo_inflate_oneshot(const uint8_t* data, size_t len, const uint8_t* dict, size_t dict_len,
                         int mode, uint8_t* out, size_t out_cap, sdz_result* res);

/* Batch of independent one-shots on `n_threads` host threads (static partition), used as
 * the CPU baseline.  in_off/in_len index into `in`; out_off/out_cap into `out`. */
int sdzo_inflate_batch_mt(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                          const uint8_t* modes, uint64_t n, uint8_t* out, const uint64_t* out_off,
                          const uint64_t* out_cap, sdz_result* res, int n_threads);

/* White-box for the table-arena limit (MANY = 1400, src/inftree.ts:242-244, src/common.ts:41): entries huft_build
 * allocates for lens[0..nl) as a literal/length set and lens[nl..nl+nd) as a distance set, with the limit lifted. */
int sdzo_table_usage(const uint8_t* lens, int nl, int nd, int* lit_entries, int* dist_entries, int* status);

/* huft_build table exposure for white-box tests: builds the reference's fixed tables
 * (src/inftree.ts:19-63) the way zlib 1.1.3 generated them; returns entry counts. */
int sdzo_fixed_tables(const int32_t** tl, int* n_tl, const int32_t** td, int* n_td);

#ifdef __cplusplus
}
#endif
#endif
