/*
 * sdz_oracle.c - CPU oracle: plain-C restatement of @stardazed/zlib's inflate path.
 *
 * TEST INFRASTRUCTURE ONLY (see sdz_oracle.h).  Not shipped, not linked by the product.
 *
 * The reference is a three-level streaming state machine; its observable behaviour
 * (bytes, 16 KiB chunking, finish() record, thrown errors, and the defects listed in
 * SURVEY Appendix A) depends on exactly when each level returns to its caller, so this
 * oracle keeps the same levels and the same return points:
 *
 *   Inflater.append / finish / inflate()  src/sd-inflate.ts:54-228   -> sdzo_append, sdzo_finish, sdzo_inflate_oneshot
 *   Inflate.inflate (container)           src/inflate.ts:132-473     -> container_step
 *   Inflate.inflateSetDictionary          src/inflate.ts:475-503     -> container_set_dictionary
 *   InfBlocks.proc / inflate_flush        src/infblocks.ts:61-628    -> blocks_proc, window_flush
 *   InfCodes.proc / inflate_fast          src/infcodes.ts:62-676     -> codes_proc, codes_fast
 *   huft_build + wrappers                 src/inftree.ts:95-392      -> huft_build, trees_bits, trees_dynamic
 *   computeAdler32                        src/adler32.ts:34-105      -> sdzo_adler32
 *   computeCRC32Little                    src/crc32.ts:48-106        -> sdzo_crc32
 *
 * JS number semantics that matter are reproduced explicitly: int32 wrap of `|`/`<<`,
 * `>>>` as a logical shift, out-of-range typed-array reads yielding `undefined` (-> 0
 * after `& 0xff`), exact integer doubles for the Adler sums.
 */
#include "sdz_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WSIZE 32768          /* 1 << MAX_BITS, src/inflate.ts:98-99            */
#define OUTBUF 16384         /* OUTPUT_BUFSIZE, src/zstream.ts:11               */
#define MANY 1400            /* ZLimits.MANY, src/common.ts:41                  */
#define BMAX 15              /* src/inftree.ts:84                               */

/* ------------------------------------------------------------------ checksums */

int32_t sdzo_adler32(const uint8_t* buf, uint64_t len, int32_t seed)
{
    /* src/adler32.ts:34-105.  `sum2` is a JS double holding an exact integer; uint64 is
     * exact over the same range (< 2^53). */
    const uint64_t BASE = 65521, NMAX = 5552;
    uint64_t sum2 = ((uint32_t)seed >> 16) & 0xffff;
    uint64_t adler = (uint32_t)seed & 0xffff;
    uint64_t off = 0;

    while (len >= NMAX) {
        len -= NMAX;
        for (uint64_t i = 0; i < NMAX; i++) { adler += buf[off++]; sum2 += adler; }
        adler %= BASE;
        sum2 += BASE;            /* :68 - the reference adds BASE instead of reducing (Q1) */
    }
    if (len) {                   /* :72 - reduction only happens when a tail exists        */
        while (len--) { adler += buf[off++]; sum2 += adler; }
        adler %= BASE;
        sum2 %= BASE;
    }
    /* :104  adler | (sum2 << 16)  with ToInt32 on both operands */
    uint32_t lo = (uint32_t)(adler & 0xffffffffu);
    uint32_t hi = (uint32_t)(sum2 & 0xffffffffu) << 16;
    return (int32_t)(lo | hi);
}

static uint32_t crc_tab[4][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;

static void crc_make_tables(void)
{
    /* src/crc32.ts:179-214 (little-endian half) */
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (0xedb88320u ^ (c >> 1)) : (c >> 1);
        crc_tab[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = crc_tab[0][n];
        for (int k = 1; k < 4; k++) {
            c = crc_tab[0][c & 0xff] ^ (c >> 8);
            crc_tab[k][n] = c;
        }
    }
}

int32_t sdzo_crc32(const uint8_t* buf, uint64_t len, int32_t seed)
{
    /* src/crc32.ts:48-106.  The head/tail split there only depends on the view's
     * alignment and does not change the value; buffers must be < 4 GiB (Q13). */
    pthread_once(&crc_once, crc_make_tables);
    uint32_t c = ~(uint32_t)seed;
    uint64_t pos = 0;
    while (len && ((uintptr_t)(buf + pos) & 3)) {
        c = crc_tab[0][(c ^ buf[pos++]) & 0xff] ^ (c >> 8);
        len--;
    }
    while (len >= 4) {
        uint32_t w;
        memcpy(&w, buf + pos, 4);
        c ^= w;
        c = crc_tab[3][c & 0xff] ^ crc_tab[2][(c >> 8) & 0xff] ^ crc_tab[1][(c >> 16) & 0xff] ^ crc_tab[0][c >> 24];
        pos += 4;
        len -= 4;
    }
    while (len--) c = crc_tab[0][(c ^ buf[pos++]) & 0xff] ^ (c >> 8);
    return (int32_t)~c;
}

/* ------------------------------------------------------------------ ZStream */

typedef struct {
    const uint8_t* next_in;   /* current append() chunk                       */
    long in_len;
    long next_in_index;       /* may become -1 (WASH, src/infcodes.ts:620-624) */
    long avail_in;
    int64_t total_in;
    uint8_t next_out[OUTBUF];
    long avail_out;
    long next_out_index;
    int64_t total_out;
    int msg;                  /* enum sdz_msg                                 */
} zstream;

/* z.next_in[p] & 0xff with typed-array out-of-range semantics (undefined & 0xff == 0) */
static inline uint32_t in_byte(const zstream* z, long p)
{
    if (p < 0 || p >= z->in_len) return 0;
    return z->next_in[p];
}

/* ------------------------------------------------------------------ inftree */

static const int cplens[31] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                                131, 163, 195, 227, 258, 0, 0 };
static const int cplext[31] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 112, 112 };
static const int cpdist[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const int cpdext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

/* module-level scratch of src/inftree.ts:87-93, one per oracle instance here */
typedef struct {
    int c[BMAX + 1];
    int r[3];
    int u[BMAX + 1];
    int x[BMAX + 1];
    int v[288];
    int hn;
    int many;       /* arena limit: MANY in the decoder; raised only by the white-box sdzo_table_usage() */
} tree_work;

static void init_work_area(tree_work* W)
{
    /* src/inftree.ts:301-311 */
    memset(W->c, 0, sizeof W->c);
    memset(W->u, 0, sizeof W->u);
    memset(W->x, 0, sizeof W->x);
    memset(W->r, 0, sizeof W->r);
    memset(W->v, 0, sizeof W->v);
}

/* src/inftree.ts:95-299.  b: code lengths; n codes; s simple codes; d/e base/extra lists;
 * *t result table index; *m requested/actual root bits; hp table arena (triples). */
static int huft_build(const uint8_t* b, int bindex, int n, int s, const int* d, const int* e,
                      int* t, int* m, int32_t* hp, tree_work* W)
{
    int a, f, g, h, i, j, k, l, mask, p, q, w, xp, y, z;
    int* c = W->c; int* r = W->r; int* u = W->u; int* x = W->x; int* v = W->v;

    p = 0; i = n;
    do { c[b[bindex + p]]++; p++; i--; } while (i != 0);          /* :131-137 */

    if (c[0] == n) { *t = -1; *m = 0; return SDZ_Z_OK; }          /* :139-143 */

    l = *m;                                                        /* :146-165 */
    for (j = 1; j <= BMAX; j++) if (c[j] != 0) break;
    k = j;
    if (l < j) l = j;
    for (i = BMAX; i != 0; i--) if (c[i] != 0) break;
    g = i;
    if (l > i) l = i;
    *m = l;

    for (y = 1 << j; j < i; j++, y <<= 1) {                        /* :168-178 */
        y -= c[j];
        if (y < 0) return SDZ_Z_DATA_ERROR;
    }
    y -= c[i];
    if (y < 0) return SDZ_Z_DATA_ERROR;
    c[i] += y;

    x[1] = j = 0; p = 1; xp = 2;                                   /* :181-188 */
    while (--i != 0) { x[xp] = (j += c[p]); xp++; p++; }

    i = 0; p = 0;                                                  /* :191-200 */
    do {
        j = b[bindex + p];
        if (j != 0) v[x[j]++] = i;
        p++;
    } while (++i < n);
    n = x[g];

    x[0] = i = 0; p = 0; h = -1; w = -l; u[0] = 0; q = 0; z = 0;   /* :203-209 */

    for (; k <= g; k++) {                                          /* :212-296 */
        a = c[k];
        while (a-- != 0) {
            while (k > w + l) {
                h++;
                w += l;
                z = g - w;
                z = (z > l) ? l : z;
                f = 1 << (j = k - w);
                if (f > a + 1) {
                    f -= a + 1;
                    xp = k;
                    if (j < z) {
                        while (++j < z) {
                            f <<= 1;
                            if (f <= c[++xp]) break;
                            f -= c[xp];
                        }
                    }
                }
                z = 1 << j;
                if (W->hn + z > W->many) return SDZ_Z_DATA_ERROR;  /* :242-244 */
                u[h] = q = W->hn;
                W->hn += z;
                if (h != 0) {
                    x[h] = i;
                    r[0] = j;
                    r[1] = l;
                    j = (int)((uint32_t)i >> (w - l));
                    r[2] = q - u[h - 1] - j;
                    memcpy(hp + (u[h - 1] + j) * 3, r, 3 * sizeof(int32_t));
                } else {
                    *t = q;
                }
            }

            r[1] = k - w;                                          /* :265-274 */
            if (p >= n) {
                r[0] = 128 + 64;
            } else if (v[p] < s) {
                r[0] = v[p] < 256 ? 0 : 32 + 64;
                r[2] = v[p++];
            } else {
                r[0] = e[v[p] - s] + 16 + 64;
                r[2] = d[v[p++] - s];
            }

            f = 1 << (k - w);                                      /* :277-280 */
            for (j = (int)((uint32_t)i >> w); j < z; j += f) memcpy(hp + (q + j) * 3, r, 3 * sizeof(int32_t));

            for (j = 1 << (k - 1); (i & j) != 0; j = (int)((uint32_t)j >> 1)) i ^= j;   /* :283-286 */
            i ^= j;

            mask = (1 << w) - 1;                                   /* :289-294 */
            while ((i & mask) != x[h]) {
                h--;
                w -= l;
                mask = (1 << w) - 1;
            }
        }
    }
    return (y != 0 && g != 1) ? SDZ_Z_BUF_ERROR : SDZ_Z_OK;       /* :298 */
}

static int trees_bits(const uint8_t* c, int* bb, int* tb, int32_t* hp, tree_work* W, zstream* z)
{
    /* src/inftree.ts:313-331 */
    init_work_area(W);
    W->hn = 0;
    W->many = MANY;
    int result = huft_build(c, 0, 19, 19, NULL, NULL, tb, bb, hp, W);
    if (result == SDZ_Z_DATA_ERROR) {
        z->msg = SDZ_MSG_OVERSUB_BITS_TREE;
    } else if (result == SDZ_Z_BUF_ERROR || *bb == 0) {
        z->msg = SDZ_MSG_INCOMPLETE_BITS_TREE;
        result = SDZ_Z_DATA_ERROR;
    }
    return result;
}

static int trees_dynamic(int nl, int nd, const uint8_t* c, int* bl, int* bd, int* tl, int* td,
                         int32_t* hp, tree_work* W, zstream* z)
{
    /* src/inftree.ts:333-379 */
    init_work_area(W);
    W->hn = 0;
    W->many = MANY;
    int result = huft_build(c, 0, nl, 257, cplens, cplext, tl, bl, hp, W);
    if (result != SDZ_Z_OK || *bl == 0) {
        if (result == SDZ_Z_DATA_ERROR) {
            z->msg = SDZ_MSG_OVERSUB_LITLEN_TREE;
        } else {
            z->msg = SDZ_MSG_INCOMPLETE_LITLEN_TREE;
            result = SDZ_Z_DATA_ERROR;
        }
        return result;
    }
    init_work_area(W);                  /* hn is NOT reset: lit+dist share the arena */
    result = huft_build(c, nl, nd, 0, cpdist, cpdext, td, bd, hp, W);
    if (result != SDZ_Z_OK || (*bd == 0 && nl > 257)) {
        if (result == SDZ_Z_DATA_ERROR) {
            z->msg = SDZ_MSG_OVERSUB_DIST_TREE;
        } else if (result == SDZ_Z_BUF_ERROR) {
            z->msg = SDZ_MSG_INCOMPLETE_DIST_TREE;
            result = SDZ_Z_DATA_ERROR;
        } else {
            z->msg = SDZ_MSG_EMPTY_DIST_TREE;
            result = SDZ_Z_DATA_ERROR;
        }
        return result;
    }
    return SDZ_Z_OK;
}

/* The reference ships its fixed tables pre-baked (src/inftree.ts:19-63).  They are the
 * output of zlib 1.1.3's huft_build on the RFC 1951 fixed code lengths with root bits
 * 9 / 5; regenerate them the same way instead of embedding the arrays
 * (oracle/check_fixed_tables.py compares the result with the reference's literals). */
static int32_t fixed_tl[512 * 3];
static int32_t fixed_td[32 * 3];
static pthread_once_t fixed_once = PTHREAD_ONCE_INIT;

static void build_fixed(void)
{
    static int32_t arena[MANY * 3];
    uint8_t c[288];
    tree_work W;
    int t = 0, m;
    int k;
    for (k = 0; k < 144; k++) c[k] = 8;
    for (; k < 256; k++) c[k] = 9;
    for (; k < 280; k++) c[k] = 7;
    for (; k < 288; k++) c[k] = 8;
    init_work_area(&W); W.hn = 0; W.many = MANY; m = 9;
    huft_build(c, 0, 288, 257, cplens, cplext, &t, &m, arena, &W);
    memcpy(fixed_tl, arena + t * 3, sizeof fixed_tl);
    for (k = 0; k < 30; k++) c[k] = 5;
    init_work_area(&W); W.hn = 0; m = 5;
    memset(arena, 0, sizeof arena);
    huft_build(c, 0, 30, 0, cpdist, cpdext, &t, &m, arena, &W);   /* incomplete (30 of 32): BUF_ERROR ignored */
    memcpy(fixed_td, arena + t * 3, sizeof fixed_td);
}

/* White-box: entries huft_build allocates for a literal/length set (lens[0..nl)) and a distance set
 * (lens[nl..nl+nd)) with the arena limit lifted - what src/inftree.ts:242-244 compares with MANY.
 * status[0], status[1]: huft_build's return value for each (Z_OK / Z_DATA_ERROR / Z_BUF_ERROR). */
int sdzo_table_usage(const uint8_t* lens, int nl, int nd, int* lit_entries, int* dist_entries, int* status)
{
    enum { BIG = 1 << 17 };
    int32_t* arena = (int32_t*)calloc((size_t)BIG * 3, sizeof(int32_t));
    if (!arena) return -1;
    tree_work W;
    int t = 0, m = 9;
    init_work_area(&W); W.hn = 0; W.many = BIG;
    status[0] = huft_build(lens, 0, nl, 257, cplens, cplext, &t, &m, arena, &W);
    *lit_entries = W.hn;
    init_work_area(&W); m = 6;
    status[1] = huft_build(lens, nl, nd, 0, cpdist, cpdext, &t, &m, arena, &W);
    *dist_entries = W.hn - *lit_entries;
    free(arena);
    return 0;
}

int sdzo_fixed_tables(const int32_t** tl, int* n_tl, const int32_t** td, int* n_td)
{
    pthread_once(&fixed_once, build_fixed);
    *tl = fixed_tl; *n_tl = 512; *td = fixed_td; *n_td = 32;
    return 0;
}

/* ------------------------------------------------------------------ infcodes / infblocks */

enum { C_START, C_LEN, C_LENEXT, C_DIST, C_DISTEXT, C_COPY, C_LIT, C_WASH, C_END, C_BADCODE };     /* src/infcodes.ts:21-32 */
enum { B_TYPE, B_LENS, B_STORED, B_TABLE, B_BTREE, B_DTREE, B_CODES, B_DRY, B_DONE, B_BAD };       /* src/infblocks.ts:21-32 */

typedef struct {
    int mode;
    int len;
    const int32_t* tree; int tree_index; int need;
    int lit;
    int get; int dist;
    int lbits, dbits;
    const int32_t* ltree; int ltree_index;
    const int32_t* dtree; int dtree_index;
} codes_t;

typedef struct {
    uint8_t window[WSIZE];
    int end;
    int32_t hufts[MANY * 3];
    codes_t codes;
    tree_work work;
    int mode;
    uint32_t bitk, bitb;
    int read, write, last;
    uint32_t n_blocks;
} blocks_t;

static const uint32_t inflate_mask[17] = { 0x0, 0x1, 0x3, 0x7, 0xf, 0x1f, 0x3f, 0x7f, 0xff, 0x1ff, 0x3ff, 0x7ff, 0xfff,
                                           0x1fff, 0x3fff, 0x7fff, 0xffff };

/* InfBlocks.inflate_flush, src/infblocks.ts:61-121 */
static int window_flush(blocks_t* s, zstream* z, int r)
{
    long p = z->next_out_index;
    int q = s->read;
    long n = (q <= s->write ? s->write : s->end) - q;
    if (n > z->avail_out) n = z->avail_out;
    if (n != 0 && r == SDZ_Z_BUF_ERROR) r = SDZ_Z_OK;
    z->avail_out -= n;
    z->total_out += n;
    memcpy(z->next_out + p, s->window + q, (size_t)n);
    p += n; q += (int)n;
    if (q == s->end) {
        q = 0;
        if (s->write == s->end) s->write = 0;
        n = s->write - q;
        if (n > z->avail_out) n = z->avail_out;
        if (n != 0 && r == SDZ_Z_BUF_ERROR) r = SDZ_Z_OK;
        z->avail_out -= n;
        z->total_out += n;
        memcpy(z->next_out + p, s->window + q, (size_t)n);
        p += n; q += (int)n;
    }
    z->next_out_index = p;
    s->read = q;
    return r;
}

#define ROOM(s, q) ((q) < (s)->read ? (s)->read - (q) - 1 : (s)->end - (q))

/* write locals back (the UPDATE pattern that precedes every return in the reference) */
#define SAVE() do { s->bitb = b; s->bitk = k; z->avail_in = n; z->total_in += p - z->next_in_index; \
                    z->next_in_index = p; s->write = q; } while (0)

/* inflate_fast, src/infcodes.ts:62-301 */
static int codes_fast(int bl, int bd, const int32_t* tl, int tl_index, const int32_t* td, int td_index,
                      blocks_t* s, zstream* z)
{
    long p = z->next_in_index, n = z->avail_in;
    uint32_t b = s->bitb, k = s->bitk;
    int q = s->write;
    int m = ROOM(s, q);
    uint32_t ml = inflate_mask[bl], md = inflate_mask[bd];
    int t, e, c, d, r;
    const int32_t* tp; int tix;

    do {
        while (k < 20) { n--; b |= in_byte(z, p++) << k; k += 8; }                 /* :96-100 */
        t = (int)(b & ml);
        tp = tl; tix = (tl_index + t) * 3;
        e = tp[tix];
        if (e == 0) {                                                              /* :107-114 */
            b >>= tp[tix + 1]; k -= tp[tix + 1];
            s->window[q++] = (uint8_t)tp[tix + 2];
            m--;
            continue;
        }
        for (;;) {
            b >>= tp[tix + 1]; k -= tp[tix + 1];
            if (e & 16) {                                                          /* :120-234 length */
                e &= 15;
                c = tp[tix + 2] + (int)(b & inflate_mask[e]);
                b >>= e; k -= e;
                while (k < 15) { n--; b |= in_byte(z, p++) << k; k += 8; }
                t = (int)(b & md);
                tp = td; tix = (td_index + t) * 3;
                e = tp[tix];
                for (;;) {
                    b >>= tp[tix + 1]; k -= tp[tix + 1];
                    if (e & 16) {
                        e &= 15;
                        while (k < (uint32_t)e) { n--; b |= in_byte(z, p++) << k; k += 8; }
                        d = tp[tix + 2] + (int)(b & inflate_mask[e]);
                        b >>= e; k -= e;
                        m -= c;
                        if (q >= d) {                                              /* :161-167 */
                            r = q - d;
                            s->window[q++] = s->window[r++];
                            s->window[q++] = s->window[r++];
                            c -= 2;
                        } else {                                                   /* :174-195 */
                            r = q - d;
                            do { r += s->end; } while (r < 0);
                            e = s->end - r;
                            if (c > e) {
                                c -= e;
                                do { s->window[q++] = s->window[r++]; } while (--e != 0);
                                r = 0;
                            }
                        }
                        do { s->window[q++] = s->window[r++]; } while (--c != 0);  /* :199-201 */
                        break;
                    } else if ((e & 64) == 0) {                                    /* :209-213 */
                        t += tp[tix + 2];
                        t += (int)(b & inflate_mask[e]);
                        tix = (td_index + t) * 3;
                        e = tp[tix];
                    } else {                                                       /* :214-231 */
                        z->msg = SDZ_MSG_BAD_DIST_CODE;
                        c = (int)(z->avail_in - n);
                        c = (int)(k >> 3) < c ? (int)(k >> 3) : c;
                        n += c; p -= c; k -= (uint32_t)c << 3;
                        SAVE();
                        return SDZ_Z_DATA_ERROR;
                    }
                }
                break;
            }
            if ((e & 64) == 0) {                                                   /* :236-248 */
                t += tp[tix + 2];
                t += (int)(b & inflate_mask[e]);
                tix = (tl_index + t) * 3;
                e = tp[tix];
                if (e == 0) {
                    b >>= tp[tix + 1]; k -= tp[tix + 1];
                    s->window[q++] = (uint8_t)tp[tix + 2];
                    m--;
                    break;
                }
            } else if (e & 32) {                                                   /* :249-264 */
                c = (int)(z->avail_in - n);
                c = (int)(k >> 3) < c ? (int)(k >> 3) : c;
                n += c; p -= c; k -= (uint32_t)c << 3;
                SAVE();
                return SDZ_Z_STREAM_END;
            } else {                                                               /* :265-282 */
                z->msg = SDZ_MSG_BAD_LITLEN_CODE;
                c = (int)(z->avail_in - n);
                c = (int)(k >> 3) < c ? (int)(k >> 3) : c;
                n += c; p -= c; k -= (uint32_t)c << 3;
                SAVE();
                return SDZ_Z_DATA_ERROR;
            }
        }
    } while (m >= 258 && n >= 10);

    c = (int)(z->avail_in - n);                                                    /* :287-300 */
    c = (int)(k >> 3) < c ? (int)(k >> 3) : c;
    n += c; p -= c; k -= (uint32_t)c << 3;
    SAVE();
    return SDZ_Z_OK;
}

static void codes_init(codes_t* cs, int bl, int bd, const int32_t* tl, int tl_index, const int32_t* td, int td_index)
{
    /* src/infcodes.ts:303-312 */
    cs->mode = C_START;
    cs->lbits = bl; cs->dbits = bd;
    cs->ltree = tl; cs->ltree_index = tl_index;
    cs->dtree = td; cs->dtree_index = td_index;
}

/* need `j` bits in the bit buffer or return to the caller (input exhausted) */
#define NEEDBITS(j) while (k < (uint32_t)(j)) { \
        if (n != 0) { r = SDZ_Z_OK; } else { SAVE(); return window_flush(s, z, r); } \
        n--; b |= in_byte(z, p++) << k; k += 8; }

/* the "no room in the window" dance shared by COPY / LIT / STORED
 * (src/infcodes.ts:547-573, :586-611; src/infblocks.ts:289-313) */
#define MAKE_ROOM() if (m == 0) { \
        if (q == s->end && s->read != 0) { q = 0; m = ROOM(s, q); } \
        if (m == 0) { \
            s->write = q; r = window_flush(s, z, r); q = s->write; m = ROOM(s, q); \
            if (q == s->end && s->read != 0) { q = 0; m = ROOM(s, q); } \
            if (m == 0) { SAVE(); return window_flush(s, z, r); } \
        } }

/* InfCodes.proc, src/infcodes.ts:314-676 */
static int codes_proc(blocks_t* s, zstream* z, int r)
{
    codes_t* cs = &s->codes;
    long p = z->next_in_index, n = z->avail_in;
    uint32_t b = s->bitb, k = s->bitk;
    int q = s->write;
    int m = ROOM(s, q);
    int j, tindex, e, f;

    for (;;) {
        switch (cs->mode) {
        case C_START:                                                              /* :338-366 */
            if (m >= 258 && n >= 10) {
                SAVE();
                r = codes_fast(cs->lbits, cs->dbits, cs->ltree, cs->ltree_index, cs->dtree, cs->dtree_index, s, z);
                p = z->next_in_index; n = z->avail_in; b = s->bitb; k = s->bitk; q = s->write; m = ROOM(s, q);
                if (r != SDZ_Z_OK) {
                    cs->mode = (r == SDZ_Z_STREAM_END) ? C_WASH : C_BADCODE;
                    break;
                }
            }
            cs->need = cs->lbits;
            cs->tree = cs->ltree;
            cs->tree_index = cs->ltree_index;
            cs->mode = C_LEN;
            /* fall through */
        case C_LEN:                                                                /* :367-426 */
            j = cs->need;
            NEEDBITS(j);
            tindex = (cs->tree_index + (int)(b & inflate_mask[j])) * 3;
            b >>= cs->tree[tindex + 1];
            k -= cs->tree[tindex + 1];
            e = cs->tree[tindex];
            if (e == 0) { cs->lit = cs->tree[tindex + 2]; cs->mode = C_LIT; break; }
            if (e & 16) { cs->get = e & 15; cs->len = cs->tree[tindex + 2]; cs->mode = C_LENEXT; break; }
            if ((e & 64) == 0) { cs->need = e; cs->tree_index = tindex / 3 + cs->tree[tindex + 2]; break; }
            if (e & 32) { cs->mode = C_WASH; break; }
            cs->mode = C_BADCODE;
            z->msg = SDZ_MSG_BAD_LITLEN_CODE;
            r = SDZ_Z_DATA_ERROR;
            SAVE();
            return window_flush(s, z, r);
        case C_LENEXT:                                                             /* :428-459 */
            j = cs->get;
            NEEDBITS(j);
            cs->len += (int)(b & inflate_mask[j]);
            b >>= j; k -= j;
            cs->need = cs->dbits;
            cs->tree = cs->dtree;
            cs->tree_index = cs->dtree_index;
            cs->mode = C_DIST;
            /* fall through */
        case C_DIST:                                                               /* :460-509 */
            j = cs->need;
            NEEDBITS(j);
            tindex = (cs->tree_index + (int)(b & inflate_mask[j])) * 3;
            b >>= cs->tree[tindex + 1];
            k -= cs->tree[tindex + 1];
            e = cs->tree[tindex];
            if (e & 16) { cs->get = e & 15; cs->dist = cs->tree[tindex + 2]; cs->mode = C_DISTEXT; break; }
            if ((e & 64) == 0) { cs->need = e; cs->tree_index = tindex / 3 + cs->tree[tindex + 2]; break; }
            cs->mode = C_BADCODE;
            z->msg = SDZ_MSG_BAD_DIST_CODE;
            r = SDZ_Z_DATA_ERROR;
            SAVE();
            return window_flush(s, z, r);
        case C_DISTEXT:                                                            /* :511-539 */
            j = cs->get;
            NEEDBITS(j);
            cs->dist += (int)(b & inflate_mask[j]);
            b >>= j; k -= j;
            cs->mode = C_COPY;
            /* fall through */
        case C_COPY:                                                               /* :540-584 */
            f = q - cs->dist;
            while (f < 0) f += s->end;
            while (cs->len != 0) {
                MAKE_ROOM();
                s->window[q++] = s->window[f++];
                m--;
                if (f == s->end) f = 0;
                cs->len--;
            }
            cs->mode = C_START;
            break;
        case C_LIT:                                                                /* :585-618 */
            MAKE_ROOM();
            r = SDZ_Z_OK;
            s->window[q++] = (uint8_t)cs->lit;
            m--;
            cs->mode = C_START;
            break;
        case C_WASH:                                                               /* :619-641 */
            if (k > 7) { k -= 8; n++; p--; }
            if (getenv("SDZO_TRACE")) fprintf(stderr, "  WASH q=%d read=%d ao=%ld total_out=%lld\n", q, s->read, z->avail_out, (long long)z->total_out);
            s->write = q; r = window_flush(s, z, r); q = s->write; m = ROOM(s, q);
            if (s->read != s->write) { SAVE(); return window_flush(s, z, r); }
            cs->mode = C_END;
            /* fall through */
        case C_END:                                                                /* :642-650 */
            r = SDZ_Z_STREAM_END;
            SAVE();
            return window_flush(s, z, r);
        case C_BADCODE:                                                            /* :652-662 */
            r = SDZ_Z_DATA_ERROR;
            SAVE();
            return window_flush(s, z, r);
        default:
            r = SDZ_Z_STREAM_ERROR;
            SAVE();
            return window_flush(s, z, r);
        }
    }
}

static const int border[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };   /* src/infblocks.ts:17-19 */

static void blocks_reset(blocks_t* s)
{
    /* src/infblocks.ts:52-58 (mode is deliberately NOT reset there) */
    s->bitk = 0; s->bitb = 0; s->read = 0; s->write = 0; s->last = 0;
}

/* InfBlocks.proc, src/infblocks.ts:123-628.  `left`, `table`, `index`, `blens`, `bb`,
 * `tb` are locals of the call in the reference (:134-140) and are therefore lost on
 * every return - the source of SURVEY Q2 and Q3.  They are locals here too. */
static int blocks_proc(blocks_t* s, zstream* z, int r)
{
    int t;
    long p = z->next_in_index, n = z->avail_in;
    uint32_t b = s->bitb, k = s->bitk;
    int q = s->write;
    int m = ROOM(s, q);
    int i;
    int left = 0;
    int table = 0;
    int index = 0;
    uint8_t blens[320];
    int bb = 0, tb = 0;
    memset(blens, 0, sizeof blens);

    for (;;) {
        switch (s->mode) {
        case B_TYPE:                                                               /* :160-242 */
            if (s->last) return SDZ_Z_STREAM_END;
            NEEDBITS(3);
            t = (int)(b & 7);
            s->last = t & 1;
            s->n_blocks++;
            if (getenv("SDZO_TRACE"))
                fprintf(stderr, "  block %u: type=%d last=%d total_out=%lld read=%d write=%d avail_out=%ld\n", s->n_blocks, t >> 1, t & 1,
                        (long long)z->total_out, s->read, q, z->avail_out);
            switch (t >> 1) {
            case 0:
                b >>= 3; k -= 3;
                t = (int)(k & 7);
                b >>= t; k -= t;
                s->mode = B_LENS;
                break;
            case 1: {
                const int32_t *ftl, *ftd; int ntl, ntd;
                sdzo_fixed_tables(&ftl, &ntl, &ftd, &ntd);
                codes_init(&s->codes, 9, 5, ftl, 0, ftd, 0);
                b >>= 3; k -= 3;
                s->mode = B_CODES;
                break;
            }
            case 2:
                b >>= 3; k -= 3;
                s->mode = B_TABLE;
                break;
            case 3:
                b >>= 3; k -= 3;
                s->mode = B_BAD;
                z->msg = SDZ_MSG_BAD_BLOCK_TYPE;
                r = SDZ_Z_DATA_ERROR;
                SAVE();
                return window_flush(s, z, r);
            }
            break;
        case B_LENS:                                                               /* :243-277 */
            NEEDBITS(32);
            if ((((~b) >> 16) & 0xffff) != (b & 0xffff)) {
                s->mode = B_BAD;
                z->msg = SDZ_MSG_BAD_STORED_LEN;
                r = SDZ_Z_DATA_ERROR;
                SAVE();
                return window_flush(s, z, r);
            }
            left = (int)(b & 0xffff);
            b = k = 0;
            s->mode = left != 0 ? B_STORED : (s->last != 0 ? B_DRY : B_TYPE);
            break;
        case B_STORED:                                                             /* :278-333 */
            if (n == 0) { SAVE(); return window_flush(s, z, r); }
            MAKE_ROOM();
            r = SDZ_Z_OK;
            t = left;
            if (t > n) t = (int)n;
            if (t > m) t = m;
            for (i = 0; i < t; i++) s->window[q + i] = (uint8_t)in_byte(z, p + i);
            p += t; n -= t; q += t; m -= t;
            left -= t;
            if (left != 0) break;
            s->mode = s->last != 0 ? B_DRY : B_TYPE;
            break;
        case B_TABLE:                                                              /* :334-380 */
            NEEDBITS(14);
            table = t = (int)(b & 0x3fff);
            if ((t & 0x1f) > 29 || ((t >> 5) & 0x1f) > 29) {
                s->mode = B_BAD;
                z->msg = SDZ_MSG_TOO_MANY_SYMS;
                r = SDZ_Z_DATA_ERROR;
                SAVE();
                return window_flush(s, z, r);
            }
            t = 258 + (t & 0x1f) + ((t >> 5) & 0x1f);
            for (i = 0; i < t; i++) blens[i] = 0;
            b >>= 14; k -= 14;
            index = 0;
            s->mode = B_BTREE;
            /* falls through: the BTREE label is commented out in the reference (:381) */
            while (index < 4 + (table >> 10)) {                                    /* :382-406 */
                NEEDBITS(3);
                blens[border[index++]] = (uint8_t)(b & 7);
                b >>= 3; k -= 3;
            }
            while (index < 19) blens[border[index++]] = 0;
            bb = 7;
            t = trees_bits(blens, &bb, &tb, s->hufts, &s->work, z);                /* :412-428 */
            if (t != SDZ_Z_OK) {
                r = t;
                if (r == SDZ_Z_DATA_ERROR) s->mode = B_BAD;
                SAVE();
                return window_flush(s, z, r);
            }
            index = 0;
            s->mode = B_DTREE;
            /* falls through: the DTREE label is commented out in the reference (:433) */
            for (;;) {                                                             /* :434-523 */
                int j, c;
                t = table;
                if (index >= 258 + (t & 0x1f) + ((t >> 5) & 0x1f)) break;
                t = bb;
                NEEDBITS(t);
                {
                    int ti = (tb + (int)(b & inflate_mask[t])) * 3;
                    t = s->hufts[ti + 1];
                    c = s->hufts[ti + 2];
                }
                if (c < 16) {
                    b >>= t; k -= t;
                    blens[index++] = (uint8_t)c;
                } else {
                    i = c == 18 ? 7 : c - 14;
                    j = c == 18 ? 11 : 3;
                    NEEDBITS(t + i);
                    b >>= t; k -= t;
                    j += (int)(b & inflate_mask[i]);
                    b >>= i; k -= i;
                    i = index;
                    t = table;
                    if (i + j > 258 + (t & 0x1f) + ((t >> 5) & 0x1f) || (c == 16 && i < 1)) {
                        s->mode = B_BAD;
                        z->msg = SDZ_MSG_BAD_REPEAT;
                        r = SDZ_Z_DATA_ERROR;
                        SAVE();
                        return window_flush(s, z, r);
                    }
                    c = c == 16 ? blens[i - 1] : 0;
                    do { blens[i++] = (uint8_t)c; } while (--j != 0);
                    index = i;
                }
            }
            tb = -1;
            {                                                                      /* :526-549 */
                int bl_ = 9, bd_ = 6, tl_ = 0, td_ = 0;
                t = trees_dynamic(257 + (t & 0x1f), 1 + ((t >> 5) & 0x1f), blens, &bl_, &bd_, &tl_, &td_,
                                  s->hufts, &s->work, z);
                if (t != SDZ_Z_OK) {
                    if (t == SDZ_Z_DATA_ERROR) s->mode = B_BAD;
                    r = t;
                    SAVE();
                    return window_flush(s, z, r);
                }
                codes_init(&s->codes, bl_, bd_, s->hufts, tl_, s->hufts, td_);
            }
            s->mode = B_CODES;
            /* fall through */
        case B_CODES:                                                              /* :552-578 */
            SAVE();
            r = codes_proc(s, z, r);
            if (r != SDZ_Z_STREAM_END) return window_flush(s, z, r);
            r = SDZ_Z_OK;
            p = z->next_in_index; n = z->avail_in; b = s->bitb; k = s->bitk; q = s->write; m = ROOM(s, q);
            if (s->last == 0) { s->mode = B_TYPE; break; }
            s->mode = B_DRY;
            /* fall through */
        case B_DRY:                                                                /* :579-594 */
            s->write = q; r = window_flush(s, z, r); q = s->write; m = ROOM(s, q);
            if (s->read != s->write) { SAVE(); return window_flush(s, z, r); }
            s->mode = B_DONE;
            /* fall through */
        case B_DONE:                                                               /* :595-604 */
            r = SDZ_Z_STREAM_END;
            SAVE();
            return window_flush(s, z, r);
        case B_BAD:                                                                /* :605-614 */
            r = SDZ_Z_DATA_ERROR;
            SAVE();
            return window_flush(s, z, r);
        default:                                                                   /* :616-625 (BTREE / DTREE re-entry, Q3) */
            r = SDZ_Z_STREAM_ERROR;
            SAVE();
            return window_flush(s, z, r);
        }
    }
}

/* ------------------------------------------------------------------ Inflate (container) */

enum {                                                                             /* src/inflate.ts:16-62 */
    M_DETECT, M_ID2, M_METHOD, M_FLAG, M_DICT4, M_DICT3, M_DICT2, M_DICT1, M_DICT0,
    M_MTIME0, M_MTIME1, M_MTIME2, M_MTIME3, M_XFLAGS, M_OS, M_EXTRA0, M_EXTRA1, M_EXTRA, M_NAME, M_COMMENT,
    M_HCRC0, M_HCRC1, M_BLOCKS, M_CHKSUM0, M_CHKSUM1, M_CHKSUM2, M_CHKSUM3, M_ISIZE0, M_ISIZE1, M_ISIZE2, M_ISIZE3,
    M_DONE, M_BAD
};
enum { G_FTEXT = 1, G_FHCRC = 2, G_FEXTRA = 4, G_FNAME = 8, G_FCOMMENT = 16 };

typedef struct {
    int mode;
    int is_gzip;
    int method, gflags;
    uint8_t* name; size_t name_len, name_cap;
    int64_t name_off;          /* absolute input offset of the first FNAME byte, -1 if none */
    int32_t mtime;
    uint32_t xlen;
    int32_t dict_checksum, full_checksum, inflated_size;
    int wbits;
    blocks_t blocks;
} container_t;

#define NEXTBYTE(var) do { if (z->avail_in == 0) return r; r = f; z->avail_in--; z->total_in++; \
                           (var) = in_byte(z, z->next_in_index++); } while (0)

static int after_os(const container_t* c)
{
    if (c->gflags & G_FEXTRA) return M_EXTRA0;
    if (c->gflags & G_FNAME) return M_NAME;
    if (c->gflags & G_FCOMMENT) return M_COMMENT;
    if (c->gflags & G_FHCRC) return M_HCRC0;
    return M_BLOCKS;
}

/* Inflate.inflate, src/inflate.ts:132-473 */
static int container_step(container_t* c, zstream* z)
{
    uint32_t b;
    const int f = SDZ_Z_OK;
    int r = SDZ_Z_BUF_ERROR;
    for (;;) {
        switch (c->mode) {
        case M_DETECT:                                                             /* :142-156 */
            if (z->avail_in == 0) return r;
            b = in_byte(z, z->next_in_index);
            if (b != 0x1f) { c->mode = M_METHOD; break; }
            c->mode = M_ID2;
            r = f;
            z->avail_in--; z->total_in++; z->next_in_index++;
            /* fall through */
        case M_ID2:                                                                /* :158-174 */
            NEXTBYTE(b);
            if (b != 0x8b) { c->mode = M_BAD; z->msg = SDZ_MSG_BAD_GZIP_ID; break; }
            c->is_gzip = 1;
            c->mode = M_METHOD;
            /* fall through */
        case M_METHOD:                                                             /* :176-196 */
            NEXTBYTE(b);
            c->method = (int)b;
            if ((c->method & 0xf) != 8) { c->mode = M_BAD; z->msg = SDZ_MSG_BAD_METHOD; break; }
            if ((c->method >> 4) + 8 > c->wbits) { c->mode = M_BAD; z->msg = SDZ_MSG_BAD_WINDOW; break; }
            c->mode = M_FLAG;
            /* fall through */
        case M_FLAG:                                                               /* :198-225 */
            NEXTBYTE(b);
            if (c->is_gzip) { c->gflags = (int)b; c->mode = M_MTIME0; break; }
            if ((((c->method << 8) + (int)b) % 31) != 0) { c->mode = M_BAD; z->msg = SDZ_MSG_BAD_HEADER_CHECK; break; }
            if ((b & 0x20) == 0) { c->mode = M_BLOCKS; break; }
            c->mode = M_DICT4;
            /* fall through */
        case M_DICT4:                                                              /* :227-237 */
            NEXTBYTE(b);
            c->dict_checksum = (int32_t)((b << 24) & 0xff000000u);
            c->mode = M_DICT3;
            /* fall through */
        case M_DICT3:
            NEXTBYTE(b);
            c->dict_checksum = (int32_t)((uint32_t)c->dict_checksum | ((b << 16) & 0xff0000u));
            c->mode = M_DICT2;
            /* fall through */
        case M_DICT2:
            NEXTBYTE(b);
            c->dict_checksum = (int32_t)((uint32_t)c->dict_checksum | ((b << 8) & 0xff00u));
            c->mode = M_DICT1;
            /* fall through */
        case M_DICT1:                                                              /* :260-270 */
            NEXTBYTE(b);
            c->dict_checksum = (int32_t)((uint32_t)c->dict_checksum | b);
            c->mode = M_DICT0;
            return SDZ_Z_NEED_DICT;
        case M_DICT0:                                                              /* :272-275 */
            c->mode = M_BAD;
            z->msg = SDZ_MSG_NEED_DICT;
            return SDZ_Z_STREAM_ERROR;
        case M_MTIME0: case M_MTIME1: case M_MTIME2: case M_MTIME3:                /* :277-295 */
            NEXTBYTE(b);
            c->mtime = (int32_t)(((uint32_t)c->mtime >> 8) | (b << 24));
            if (c->mode != M_MTIME3) { c->mode++; break; }
            c->mode = M_XFLAGS;
            /* fall through */
        case M_XFLAGS: case M_OS: case M_HCRC0: case M_HCRC1:                      /* :297-331 */
            if (z->avail_in == 0) return r;
            r = f;
            z->avail_in--; z->total_in++; z->next_in_index++;
            if (c->mode == M_OS) c->mode = after_os(c);
            else c->mode++;                     /* XFLAGS->OS, HCRC0->HCRC1, HCRC1->BLOCKS */
            break;
        case M_EXTRA0: case M_EXTRA1:                                              /* :333-347 */
            NEXTBYTE(b);
            c->xlen = (c->xlen >> 8) | (b << 24);
            if (c->mode == M_EXTRA0) break;     /* mode is never advanced: Q5 */
            c->xlen = c->xlen >> 16;
            /* fall through */
        case M_EXTRA:                                                              /* :349-373 (unreachable, kept for fidelity) */
            if (z->avail_in == 0) return r;
            r = f;
            z->avail_in--; z->total_in++; z->next_in_index++;
            c->xlen--;
            if (c->xlen == 0) {
                if (c->gflags & G_FNAME) c->mode = M_NAME;
                else if (c->gflags & G_FCOMMENT) c->mode = M_COMMENT;
                else if (c->gflags & G_FHCRC) c->mode = M_HCRC0;
                else c->mode = M_BLOCKS;
            }
            break;
        case M_NAME: case M_COMMENT:                                               /* :375-401 */
            NEXTBYTE(b);
            if (b != 0) {
                if (c->mode == M_NAME) {
                    if (c->name_len == c->name_cap) {
                        c->name_cap = c->name_cap ? c->name_cap * 2 : 64;
                        c->name = (uint8_t*)realloc(c->name, c->name_cap);
                    }
                    if (c->name_len == 0) c->name_off = z->total_in - 1;
                    c->name[c->name_len++] = (uint8_t)b;
                }
            } else {
                if (c->mode != M_COMMENT && (c->gflags & G_FCOMMENT)) c->mode = M_COMMENT;
                else if (c->gflags & G_FHCRC) c->mode = M_HCRC0;
                else c->mode = M_BLOCKS;
            }
            break;
        case M_BLOCKS:                                                             /* :403-421 */
            r = blocks_proc(&c->blocks, z, r);
            if (r == SDZ_Z_DATA_ERROR) { c->mode = M_BAD; break; }
            if (r != SDZ_Z_STREAM_END) return r;
            r = f;
            blocks_reset(&c->blocks);
            if (c->method == 0) { c->mode = M_DONE; break; }
            c->mode = M_CHKSUM0;
            /* fall through */
        case M_CHKSUM0: case M_CHKSUM1: case M_CHKSUM2: case M_CHKSUM3:            /* :423-448 */
            NEXTBYTE(b);
            if (c->is_gzip) c->full_checksum = (int32_t)(((uint32_t)c->full_checksum >> 8) | (b << 24));
            else c->full_checksum = (int32_t)(((uint32_t)c->full_checksum << 8) | b);
            c->mode++;
            if (c->mode == M_ISIZE0 && !c->is_gzip) c->mode = M_DONE;
            break;
        case M_ISIZE0: case M_ISIZE1: case M_ISIZE2: case M_ISIZE3:                /* :450-463 */
            NEXTBYTE(b);
            c->inflated_size = (int32_t)(((uint32_t)c->inflated_size >> 8) | (b << 24));
            c->mode++;
            break;
        case M_DONE:
            return SDZ_Z_STREAM_END;
        case M_BAD:
            return SDZ_Z_DATA_ERROR;
        default:
            return SDZ_Z_STREAM_ERROR;
        }
    }
}

/* Inflate.inflateSetDictionary + InfBlocks.set_dictionary, src/inflate.ts:475-503, src/infblocks.ts:630-633 */
static int container_set_dictionary(container_t* c, const uint8_t* dict, size_t dict_len)
{
    if (c->mode != M_DICT0) return SDZ_Z_STREAM_ERROR;
    size_t index = 0, length = dict_len;
    if (length >= ((size_t)1 << c->wbits)) {
        length = ((size_t)1 << c->wbits) - 1;          /* keeps 32767 bytes, not 32768 (Q14) */
        index = dict_len - length;
    }
    if (sdzo_adler32(dict, dict_len, 1) != c->dict_checksum) return SDZ_Z_DATA_ERROR;
    memcpy(c->blocks.window, dict + index, length);
    c->blocks.read = c->blocks.write = (int)length;
    c->mode = M_BLOCKS;
    return SDZ_Z_OK;
}

/* ------------------------------------------------------------------ Inflater (public wrapper) */

struct sdzo_inflater {
    container_t inf;
    zstream z;
    uint8_t* dict; size_t dict_len; int have_dict;
    int32_t checksum; int have_checksum;         /* `number | undefined`, src/sd-inflate.ts:58 */
    int last_status;
};

sdzo_inflater* sdzo_inflater_new(int raw, const uint8_t* dict, size_t dict_len)
{
    if (raw && dict) return NULL;                                                  /* RangeError, src/sd-inflate.ts:69-71 */
    sdzo_inflater* I = (sdzo_inflater*)calloc(1, sizeof *I);
    if (!I) return NULL;
    I->inf.wbits = 15;
    I->inf.blocks.end = WSIZE;
    I->inf.blocks.mode = B_TYPE;
    I->inf.mode = raw ? M_BLOCKS : M_DETECT;                                       /* src/inflate.ts:97-101 */
    I->inf.name_off = -1;
    I->z.avail_out = OUTBUF;
    if (dict) {
        I->dict = (uint8_t*)malloc(dict_len ? dict_len : 1);
        memcpy(I->dict, dict, dict_len);
        I->dict_len = dict_len;
        I->have_dict = 1;
    }
    return I;
}

void sdzo_inflater_free(sdzo_inflater* I)
{
    if (!I) return;
    free(I->dict);
    free(I->inf.name);
    free(I);
}

void sdzo_chunks_free(sdzo_chunks* c)
{
    free(c->data); free(c->chunk_len);
    memset(c, 0, sizeof *c);
}

static void chunks_push(sdzo_chunks* c, const uint8_t* p, size_t n)
{
    if (c->len + n > c->cap) {
        size_t nc = c->cap ? c->cap * 2 : 65536;
        while (nc < c->len + n) nc *= 2;
        c->data = (uint8_t*)realloc(c->data, nc);
        c->cap = nc;
    }
    memcpy(c->data + c->len, p, n);
    c->len += n;
    if (c->n_chunks == c->chunk_cap) {
        c->chunk_cap = c->chunk_cap ? c->chunk_cap * 2 : 16;
        c->chunk_len = (uint32_t*)realloc(c->chunk_len, c->chunk_cap * sizeof(uint32_t));
    }
    c->chunk_len[c->n_chunks++] = (uint32_t)n;
}

/* everything that decides what the next loop trip of append() does; if a trip leaves it
 * unchanged and the loop condition still holds, the reference spins forever (Q4) */
typedef struct {
    long avail_in, next_in_index; int64_t total_out;
    int cmode, bmode, kmode; uint32_t bitk; int read, write, len; int nomore;
} progress_t;

static progress_t snapshot(const sdzo_inflater* I, int nomore)
{
    progress_t s;
    memset(&s, 0, sizeof s);
    s.avail_in = I->z.avail_in; s.next_in_index = I->z.next_in_index; s.total_out = I->z.total_out;
    s.cmode = I->inf.mode; s.bmode = I->inf.blocks.mode; s.kmode = I->inf.blocks.codes.mode;
    s.bitk = I->inf.blocks.bitk; s.read = I->inf.blocks.read; s.write = I->inf.blocks.write;
    s.len = I->inf.blocks.codes.len; s.nomore = nomore;
    return s;
}

/* Inflater.append, src/sd-inflate.ts:87-153 */
int sdzo_append(sdzo_inflater* I, const uint8_t* data, size_t len, sdzo_chunks* out)
{
    zstream* z = &I->z;
    out->len = 0; out->n_chunks = 0;
    if (len == 0) return SDZ_THROW_NONE;                                           /* :92-94 */

    int nomoreinput = 0;
    z->next_in = data; z->in_len = (long)len; z->avail_in = (long)len; z->next_in_index = 0;   /* z.append, src/zstream.ts:46-50 */

    do {
        z->next_out_index = 0;
        z->avail_out = OUTBUF;
        if (z->avail_in == 0 && !nomoreinput) { z->next_in_index = 0; nomoreinput = 1; }       /* :105-108 */

        progress_t before = snapshot(I, nomoreinput);
        int err = container_step(&I->inf, z);
        I->last_status = err;
        if (getenv("SDZO_TRACE"))
            fprintf(stderr, "call: err=%d total_out=%lld out_idx=%ld avail_in=%ld bmode=%d cmode=%d read=%d write=%d nblocks=%u\n", err,
                    (long long)z->total_out, z->next_out_index, z->avail_in, I->inf.blocks.mode, I->inf.blocks.codes.mode,
                    I->inf.blocks.read, I->inf.blocks.write, I->inf.blocks.n_blocks);

        if (nomoreinput && err == SDZ_Z_BUF_ERROR) {                               /* :111-115 */
            if (z->avail_in != 0) { out->len = 0; out->n_chunks = 0; return SDZ_THROW_BAD_INPUT; }
        } else if (err == SDZ_Z_NEED_DICT) {                                       /* :116-126 */
            if (I->have_dict) {
                if (container_set_dictionary(&I->inf, I->dict, I->dict_len) != SDZ_Z_OK) {
                    out->len = 0; out->n_chunks = 0;
                    return SDZ_THROW_DICT_INVALID;
                }
            } else {
                out->len = 0; out->n_chunks = 0;
                return SDZ_THROW_DICT_REQUIRED;
            }
        } else if (err != SDZ_Z_OK && err != SDZ_Z_STREAM_END) {                   /* :127-129 */
            out->len = 0; out->n_chunks = 0;
            return SDZ_THROW_INFLATE_ERROR;
        }
        if ((nomoreinput || err == SDZ_Z_STREAM_END) && z->avail_in == (long)len) {  /* :130-132 */
            out->len = 0; out->n_chunks = 0;
            return SDZ_THROW_BAD_INPUT_DATA;
        }
        if (z->next_out_index) {                                                   /* :133-149 */
            int use_crc = I->inf.is_gzip;
            if (!I->have_checksum) { I->checksum = use_crc ? 0 : 1; I->have_checksum = 1; }
            if (use_crc) I->checksum = sdzo_crc32(z->next_out, (uint64_t)z->next_out_index, I->checksum);
            else I->checksum = sdzo_adler32(z->next_out, (uint64_t)z->next_out_index, I->checksum);
            chunks_push(out, z->next_out, (size_t)z->next_out_index);
        }
        if (z->avail_in > 0 || z->avail_out == 0) {
            progress_t after = snapshot(I, nomoreinput);
            if (memcmp(&before, &after, sizeof before) == 0 && z->next_out_index == 0) {
                out->len = 0; out->n_chunks = 0;
                return SDZ_THROW_HANG;
            }
        }
    } while (z->avail_in > 0 || z->avail_out == 0);                                /* :150 */
    return SDZ_THROW_NONE;
}

/* Inflater.finish, src/sd-inflate.ts:159-179 */
void sdzo_finish(sdzo_inflater* I, sdz_result* res)
{
    container_t* c = &I->inf;
    int32_t stored = c->full_checksum;
    int32_t stored_size = c->inflated_size;
    /* isComplete, src/inflate.ts:103-107 */
    int blocks_complete = (c->blocks.mode == 0 || c->blocks.mode == 8) && c->blocks.bitb == 0 && c->blocks.bitk == 0;
    int complete = c->mode == M_DONE && blocks_complete;

    int cks = stored == 0 ? SDZ_UNCHECKED
              : ((I->have_checksum && stored == I->checksum) ? SDZ_MATCH : SDZ_MISMATCH);      /* :164 (undefined never === number, Q8) */
    /* :165  storedSize === z.total_out compares an int32 with an exact double */
    int fsz = stored_size == 0 ? SDZ_UNCHECKED : ((int64_t)stored_size == I->z.total_out ? SDZ_MATCH : SDZ_MISMATCH);

    res->out_len = (uint64_t)I->z.total_out;
    res->total_in = (uint64_t)I->z.total_in;
    res->zstatus = I->last_status;
    res->stored_checksum = stored;
    res->running_checksum = I->have_checksum ? I->checksum : 0;
    res->have_running = (uint8_t)I->have_checksum;
    res->stored_isize = stored_size;
    res->mtime = c->mtime;
    res->name_off = c->name_off < 0 ? 0 : (uint32_t)c->name_off;
    res->name_len = (uint32_t)c->name_len;
    res->n_blocks = c->blocks.n_blocks;
    res->msg_id = (uint8_t)I->z.msg;
    /* containerFormat, src/inflate.ts:128-130 */
    res->container = (uint8_t)(c->is_gzip ? SDZ_GZIP : (c->method == 0 ? SDZ_RAW : SDZ_ZLIB));
    res->complete = (uint8_t)complete;
    res->checksum_state = (uint8_t)cks;
    res->size_state = (uint8_t)fsz;
    res->success = (uint8_t)(complete && cks != SDZ_MISMATCH && fsz != SDZ_MISMATCH);
}

size_t sdzo_file_name(sdzo_inflater* I, const uint8_t** bytes)
{
    *bytes = I->inf.name;
    return I->inf.name_len;
}

int sdzo_inflate_oneshot(const uint8_t* data, size_t len, const uint8_t* dict, size_t dict_len,
                         int mode, uint8_t* out, size_t out_cap, sdz_result* res)
{
    memset(res, 0, sizeof *res);
    int raw = (mode == SDZ_MODE_RAW);
    if (mode == SDZ_MODE_SNIFF) {                                                  /* src/sd-inflate.ts:189-207 */
        if (len < 2) { res->thrown_inflate = SDZ_THROW_TOO_SMALL; res->thrown_append = SDZ_THROW_NONE; return 0; }
        int ident = (data[0] == 0x78 && (((data[0] << 8) + data[1]) % 31) == 0) || (data[0] == 0x1f && data[1] == 0x8b);
        raw = !ident;
    }
    /* inflate() passes {dictionary, raw}; the constructor throws RangeError for raw+dictionary
     * only when a dictionary was actually given (src/sd-inflate.ts:67-71) */
    if (raw && dict) {
        res->thrown_append = SDZ_THROW_NONE;
        res->thrown_inflate = SDZ_THROW__COUNT;   /* RangeError: not an Appendix-D data error; host binding raises it */
        return 0;
    }
    sdzo_inflater* I = sdzo_inflater_new(raw, dict, dict_len);
    sdzo_chunks ch; memset(&ch, 0, sizeof ch);
    int thrown = sdzo_append(I, data, len, &ch);
    sdzo_finish(I, res);
    res->thrown_append = (uint8_t)thrown;
    int rc = 0;
    if (thrown) {
        res->out_len = 0;
        res->thrown_inflate = (uint8_t)thrown;           /* inflate() lets append()'s exception propagate */
    } else {
        res->out_len = ch.len;
        if (ch.len > out_cap) rc = -1;
        else if (ch.len) memcpy(out, ch.data, ch.len);
        if (!res->success) {                                                       /* src/sd-inflate.ts:214-225 */
            if (!res->complete) res->thrown_inflate = SDZ_THROW_UNEXPECTED_EOF;
            else if (res->checksum_state == SDZ_MISMATCH) res->thrown_inflate = SDZ_THROW_INTEGRITY;
            else if (res->size_state == SDZ_MISMATCH) res->thrown_inflate = SDZ_THROW_SIZE_CHECK;
            else res->thrown_inflate = SDZ_THROW_DECOMPRESSION;
        }
    }
    sdzo_chunks_free(&ch);
    sdzo_inflater_free(I);
    return rc;
}

/* ------------------------------------------------------------------ threaded batch (CPU baseline) */

typedef struct {
    const uint8_t* in; const uint64_t* in_off; const uint64_t* in_len; const uint8_t* modes;
    uint8_t* out; const uint64_t* out_off; const uint64_t* out_cap; sdz_result* res;
    uint64_t lo, hi; int rc;
} batch_job;

static void* batch_worker(void* arg)
{
    batch_job* j = (batch_job*)arg;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        int rc = sdzo_inflate_oneshot(j->in + j->in_off[i], (size_t)j->in_len[i], NULL, 0,
                                      j->modes ? j->modes[i] : SDZ_MODE_SNIFF,
                                      j->out + j->out_off[i], (size_t)j->out_cap[i], &j->res[i]);
        j->res[i].out_off = j->out_off[i];
        if (rc) j->rc = rc;
    }
    return NULL;
}

int sdzo_inflate_batch_mt(const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                          const uint8_t* modes, uint64_t n, uint8_t* out, const uint64_t* out_off,
                          const uint64_t* out_cap, sdz_result* res, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n && n > 0) n_threads = (int)n;
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof *th);
    batch_job* jobs = (batch_job*)calloc((size_t)n_threads, sizeof *jobs);
    pthread_once(&fixed_once, build_fixed);
    pthread_once(&crc_once, crc_make_tables);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (batch_job){ in, in_off, in_len, modes, out, out_off, out_cap, res,
                               n * (uint64_t)t / (uint64_t)n_threads, n * (uint64_t)(t + 1) / (uint64_t)n_threads, 0 };
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}
