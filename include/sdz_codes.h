/*
 * sdz_codes.h - value domains shared by the C-ABI (include/sdzcuda.h), the CUDA
 * kernels and the CPU oracle (oracle/).  Plain C, no dependencies.
 *
 * Every enumerator cites the reference construct it stands for
 * (paths relative to the @stardazed/zlib source tree).
 */
#ifndef SDZ_CODES_H
#define SDZ_CODES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ZStatus, src/common.ts:15-23 */
enum sdz_zstatus {
    SDZ_Z_OK = 0,
    SDZ_Z_STREAM_END = 1,
    SDZ_Z_NEED_DICT = 2,
    SDZ_Z_STREAM_ERROR = -2,
    SDZ_Z_DATA_ERROR = -3,
    SDZ_Z_MEM_ERROR = -4,
    SDZ_Z_BUF_ERROR = -5
};

/* z.msg values.  Index into sdz_msg_text[]. */
enum sdz_msg {
    SDZ_MSG_NONE = 0,                 /* ""                                     */
    SDZ_MSG_BAD_GZIP_ID,              /* src/inflate.ts:169                     */
    SDZ_MSG_BAD_METHOD,               /* src/inflate.ts:187                     */
    SDZ_MSG_BAD_WINDOW,               /* src/inflate.ts:192                     */
    SDZ_MSG_BAD_HEADER_CHECK,         /* src/inflate.ts:216                     */
    SDZ_MSG_NEED_DICT,                /* src/inflate.ts:274                     */
    SDZ_MSG_BAD_BLOCK_TYPE,           /* src/infblocks.ts:231                   */
    SDZ_MSG_BAD_STORED_LEN,           /* src/infblocks.ts:263                   */
    SDZ_MSG_TOO_MANY_SYMS,            /* src/infblocks.ts:357                   */
    SDZ_MSG_BAD_REPEAT,               /* src/infblocks.ts:505                   */
    SDZ_MSG_OVERSUB_BITS_TREE,        /* src/inftree.ts:325                     */
    SDZ_MSG_INCOMPLETE_BITS_TREE,     /* src/inftree.ts:327                     */
    SDZ_MSG_OVERSUB_LITLEN_TREE,      /* src/inftree.ts:350                     */
    SDZ_MSG_INCOMPLETE_LITLEN_TREE,   /* src/inftree.ts:353                     */
    SDZ_MSG_OVERSUB_DIST_TREE,        /* src/inftree.ts:365                     */
    SDZ_MSG_INCOMPLETE_DIST_TREE,     /* src/inftree.ts:368                     */
    SDZ_MSG_EMPTY_DIST_TREE,          /* src/inftree.ts:372                     */
    SDZ_MSG_BAD_DIST_CODE,            /* src/infcodes.ts:215,:500               */
    SDZ_MSG_BAD_LITLEN_CODE,          /* src/infcodes.ts:266,:417               */
    /* SDZ_PARITY_SPEC only: zlib 1.3's texts for what the reference accepts, or words differently */
    SDZ_MSG_DIST_TOO_FAR,             /* zlib inflate.c / inffast.c              */
    SDZ_MSG_MISSING_EOB,
    SDZ_MSG_BAD_CODE_LENGTHS_SET,
    SDZ_MSG_BAD_LITLEN_SET,
    SDZ_MSG_BAD_DIST_SET,
    SDZ_MSG_BAD_GZIP_FLAGS,
    SDZ_MSG_BAD_HEADER_CRC,
    SDZ_MSG__COUNT
};

/* What Inflater.append() / inflate() would throw.
 * 0 = nothing thrown. */
enum sdz_thrown {
    SDZ_THROW_NONE = 0,
    SDZ_THROW_BAD_INPUT,          /* "inflate error: bad input"                      src/sd-inflate.ts:113 */
    SDZ_THROW_DICT_INVALID,       /* "Custom dictionary is not valid for this data"  src/sd-inflate.ts:120 */
    SDZ_THROW_DICT_REQUIRED,      /* "Custom dictionary required for this data"      src/sd-inflate.ts:124 */
    SDZ_THROW_INFLATE_ERROR,      /* "inflate error: " + z.msg                       src/sd-inflate.ts:128 */
    SDZ_THROW_BAD_INPUT_DATA,     /* "inflate error: bad input data"                 src/sd-inflate.ts:131 */
    SDZ_THROW_HANG,               /* reference append() never returns; not a JS exception       */
    SDZ_THROW_TOO_SMALL,          /* "data buffer is too small"                      src/sd-inflate.ts:195 */
    SDZ_THROW_UNEXPECTED_EOF,     /* "Unexpected EOF during decompression"           src/sd-inflate.ts:216 */
    SDZ_THROW_INTEGRITY,          /* "Data integrity check failed"                   src/sd-inflate.ts:219 */
    SDZ_THROW_SIZE_CHECK,         /* "Data size check failed"                        src/sd-inflate.ts:222 */
    SDZ_THROW_DECOMPRESSION,      /* "Decompression error"                           src/sd-inflate.ts:224 */
    SDZ_THROW__COUNT
};

/* ContainerFormat, src/inflate.ts:72-76 */
enum sdz_container { SDZ_RAW = 0, SDZ_ZLIB = 1, SDZ_GZIP = 2 };

/* "match" | "mismatch" | "unchecked", src/sd-inflate.ts:45-47 */
enum sdz_check { SDZ_UNCHECKED = 0, SDZ_MATCH = 1, SDZ_MISMATCH = 2 };

/* How the container is chosen for one input buffer. */
enum sdz_mode {
    SDZ_MODE_SNIFF = 0,    /* inflate():  zlib iff 0x78 + %31, gzip iff 1F 8B, else raw (src/sd-inflate.ts:203-207) */
    SDZ_MODE_INFLATER = 1, /* new Inflater(): gzip iff first byte 0x1F, else zlib         (src/inflate.ts:142-175) */
    SDZ_MODE_RAW = 2       /* new Inflater({raw:true})                                    (src/inflate.ts:100)     */
};

static const char* const sdz_msg_text[SDZ_MSG__COUNT] = {
    "",
    "invalid gzip id",
    "unknown compression method",
    "invalid window size",
    "incorrect header check",
    "need dictionary",
    "invalid block type",
    "invalid stored block lengths",
    "too many length or distance symbols",
    "invalid bit length repeat",
    "oversubscribed dynamic bit lengths tree",
    "incomplete dynamic bit lengths tree",
    "oversubscribed literal/length tree",
    "incomplete literal/length tree",
    "oversubscribed distance tree",
    "incomplete distance tree",
    "empty distance tree with lengths",
    "invalid distance code",
    "invalid literal/length code",
    "invalid distance too far back",
    "invalid code -- missing end-of-block",
    "invalid code lengths set",
    "invalid literal/lengths set",
    "invalid distances set",
    "unknown header flags set",
    "header crc mismatch",
};

static const char* const sdz_thrown_text[SDZ_THROW__COUNT] = {
    "",
    "inflate error: bad input",
    "Custom dictionary is not valid for this data",
    "Custom dictionary required for this data",
    "inflate error: ",               /* + sdz_msg_text[msg_id] */
    "inflate error: bad input data",
    "<reference does not terminate>",
    "data buffer is too small",
    "Unexpected EOF during decompression",
    "Data integrity check failed",
    "Data size check failed",
    "Decompression error",
};

/* One record per input buffer: everything Inflater.finish() reports
 * (src/sd-inflate.ts:159-179) plus what append()/inflate() would throw. */
typedef struct sdz_result {
    uint64_t out_off;           /* where this stream's bytes start in the output arena           */
    uint64_t out_len;           /* bytes produced == z.total_out (src/infblocks.ts:81)           */
    uint64_t total_in;          /* bytes consumed == z.total_in                                  */
    int32_t  zstatus;           /* last ZStatus returned by Inflate.inflate() (src/inflate.ts:132) */
    int32_t  stored_checksum;   /* Inflate.checksum, signed (src/inflate.ts:120,:434-441)        */
    int32_t  running_checksum;  /* Inflater.checksum, signed; 0 when have_running == 0           */
    int32_t  stored_isize;      /* Inflate.fullSize, signed (src/inflate.ts:124,:461)            */
    int32_t  mtime;             /* gzip MTIME, signed (src/inflate.ts:289); 0 => modDate undefined */
    uint32_t name_off;          /* gzip FNAME bytes (Latin-1) inside the input buffer            */
    uint32_t name_len;
    uint32_t n_blocks;          /* deflate blocks whose header was parsed (diagnostic)           */
    uint8_t  msg_id;            /* enum sdz_msg: z.msg at the end                                */
    uint8_t  thrown_append;     /* enum sdz_thrown: what Inflater.append() throws (0 = returns)  */
    uint8_t  thrown_inflate;    /* enum sdz_thrown: what inflate() throws (0 = returns data)     */
    uint8_t  container;         /* enum sdz_container (src/inflate.ts:128-130)                   */
    uint8_t  complete;          /* finish().complete                                             */
    uint8_t  checksum_state;    /* enum sdz_check: finish().checksum                             */
    uint8_t  size_state;        /* enum sdz_check: finish().fileSize                             */
    uint8_t  success;           /* finish().success                                              */
    uint8_t  have_running;      /* 0 when no output chunk was ever produced          */
    uint8_t  reserved[7];
} sdz_result;

/* Where an Inflater stopped at the end of the input it had been given - the state the reference keeps between
 * append() calls (src/inflate.ts:79-95, src/infblocks.ts:40-50, src/infcodes.ts:44-60), reduced to what the device
 * decoder needs to continue: it re-parses the header of the current block (tables are rebuilt, not stored) and goes on at
 * the symbol the reference was waiting on.  Lives in device memory; written and read by inflate_kernel. */
enum sdz_resume_kind {
    SDZ_RESUME_START = 0,      /* nothing consumed yet (the container header is incomplete)                            */
    SDZ_RESUME_AT_BLOCK = 1,   /* a block header starts at block_bit (TYPE / LENS / TABLE are re-entrant)              */
    SDZ_RESUME_IN_CODES = 2,   /* inside the block whose header is at block_bit; the next symbol starts at sym_bit     */
    SDZ_RESUME_AT_TRAILER = 3, /* the final block ended at block_bit; the container trailer is incomplete              */
    SDZ_RESUME_BROKEN_Q3 = 4,  /* input ended inside BTREE / DTREE: the next append() throws               */
    SDZ_RESUME_DONE = 5,       /* stream complete: more input makes append() spin                          */
    SDZ_RESUME_FAILED = 6      /* append() threw                                                                       */
};
typedef struct sdz_resume {
    uint64_t block_bit;
    uint64_t sym_bit;
    uint32_t pos;              /* bytes produced so far (z.total_out)                                                  */
    int32_t  ring_q;           /* InfBlocks.write == read: append() returns with the window flushed                    */
    uint32_t n_blocks;
    int32_t  mtime;
    uint32_t name_off, name_len;
    uint32_t prev_len;         /* input bytes seen so far: the reference had loaded all of them into its bit buffer    */
    uint16_t dict_used;        /* bytes of the preset dictionary in the window                            */
    uint8_t  kind;             /* enum sdz_resume_kind                                                                 */
    uint8_t  flags;            /* 1: gzip, 2: raw, 4: BFINAL of the current block                                      */
    uint8_t  method;           /* zlib CMF / gzip CM byte                                                              */
    uint8_t  reserved[11];
} sdz_resume;

#ifdef __cplusplus
}
#endif
#endif /* SDZ_CODES_H */
