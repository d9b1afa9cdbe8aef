/*
 * corpus.c - deterministic synthetic corpora for tests and bench.py and
 * their compression with system zlib, wrapped the way the reference's Deflater wraps its
 * output (src/sd-deflate.ts:98-165): zlib header 78 01 (or 78 20 + DICTID), gzip header with
 * OS = 0xff, FNAME optional.  Test/bench infrastructure only - not part of the product.
 *
 * System zlib 1.3 (memLevel 8, windowBits 15, default strategy) reproduces the reference
 * deflate's output sizes at every level on the reference's own text fixture and its level-6
 * payload byte for byte, which is why it stands in for "the reference's
 * own deflate" here (the reference itself cannot run: no JS runtime in this image).
 *
 * PRNG: splitmix64; stream i of a batch uses seed 0x5D211B00 + i.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define SEED0 0x5D211B00ull
#define VOCAB 4096

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd(rng_t* r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rnd01(rng_t* r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rnd_below(rng_t* r, uint32_t n) { return (uint32_t)((rnd(r) >> 32) * (uint64_t)n >> 32); }

static char vocab[VOCAB][12];
static uint8_t vocab_len[VOCAB];
static double zipf_cdf[VOCAB];
static pthread_once_t vocab_once = PTHREAD_ONCE_INIT;

static void make_vocab(void)
{
    /* letters by (approximate) English frequency, per mille */
    static const char letters[] = "etaoinshrdlcumwfgypbvkjxqz";
    static const int freq[] = { 127, 91, 82, 75, 70, 67, 63, 61, 60, 43, 40, 28, 28, 24, 24, 22, 20, 20, 19, 15, 10, 8, 2, 2, 1, 1 };
    int cum[26], tot = 0;
    for (int i = 0; i < 26; i++) { tot += freq[i]; cum[i] = tot; }
    rng_t r = { SEED0 - 1 };
    for (int w = 0; w < VOCAB; w++) {
        int len = 2 + (int)rnd_below(&r, 9);           /* 2..10 */
        for (int j = 0; j < len; j++) {
            int x = (int)rnd_below(&r, (uint32_t)tot), k = 0;
            while (cum[k] <= x) k++;
            vocab[w][j] = letters[k];
        }
        vocab_len[w] = (uint8_t)len;
    }
    double h = 0;
    for (int w = 0; w < VOCAB; w++) h += 1.0 / (w + 1);
    double c = 0;
    for (int w = 0; w < VOCAB; w++) { c += 1.0 / (w + 1) / h; zipf_cdf[w] = c; }
    zipf_cdf[VOCAB - 1] = 1.0;
}

/* kind 0: English-like text */
static void gen_text(uint64_t seed, uint8_t* out, size_t n)
{
    pthread_once(&vocab_once, make_vocab);
    rng_t r = { seed };
    size_t p = 0, col = 0;
    int until_punct = 6 + (int)rnd_below(&r, 9);
    while (p < n) {
        double u = rnd01(&r);
        int lo = 0, hi = VOCAB - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (zipf_cdf[mid] < u) lo = mid + 1; else hi = mid; }
        int len = vocab_len[lo];
        for (int j = 0; j < len && p < n; j++) out[p++] = (uint8_t)vocab[lo][j];
        col += (size_t)len;
        if (--until_punct == 0) {
            if (p < n) out[p++] = (rnd(&r) & 1) ? ',' : '.';
            col++;
            until_punct = 6 + (int)rnd_below(&r, 9);
        }
        if (col >= 70) { if (p < n) out[p++] = '\n'; col = 0; }
        else { if (p < n) out[p++] = ' '; col++; }
    }
}

/* kind 1: little-endian float32 triples from a quantised random walk (models vertex data) */
static void gen_binary(uint64_t seed, uint8_t* out, size_t n)
{
    rng_t r = { seed };
    float v[3] = { 0, 0, 0 };
    size_t p = 0;
    while (p < n) {
        for (int k = 0; k < 3; k++) {
            double u1 = rnd01(&r) + 1e-12, u2 = rnd01(&r);
            double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2) * 0.01;
            v[k] += (float)(floor(g * 1024.0 + 0.5) / 1024.0);
            uint8_t b[4];
            memcpy(b, &v[k], 4);
            for (int j = 0; j < 4 && p < n; j++) out[p++] = b[j];
        }
    }
}

/* kind 2: uniform random bytes (forces stored blocks) */
static void gen_random(uint64_t seed, uint8_t* out, size_t n)
{
    rng_t r = { seed };
    size_t p = 0;
    while (p < n) { uint64_t x = rnd(&r); for (int j = 0; j < 8 && p < n; j++) { out[p++] = (uint8_t)x; x >>= 8; } }
}

/* kind 3: short ASCII (fixed blocks) - the caller picks n in 1..200 */
static void gen_tiny(uint64_t seed, uint8_t* out, size_t n)
{
    rng_t r = { seed };
    for (size_t p = 0; p < n; p++) out[p] = (uint8_t)(32 + rnd_below(&r, 95));
}

/* kind 4: runs of one byte and of short motifs (period 2..7): dist < len and dist == 1 matches */
static void gen_runs(uint64_t seed, uint8_t* out, size_t n)
{
    rng_t r = { seed };
    size_t p = 0;
    while (p < n) {
        double u = rnd01(&r) + 1e-12;
        size_t run = 1 + (size_t)(-log(u) * 200.0);
        if (run > 4096) run = 4096;
        int period = (rnd(&r) & 1) ? 1 : 2 + (int)rnd_below(&r, 6);
        uint8_t motif[8];
        for (int j = 0; j < period; j++) motif[j] = (uint8_t)rnd(&r);
        for (size_t j = 0; j < run && p < n; j++) out[p++] = motif[j % (size_t)period];
    }
}

void sdzc_generate(int kind, uint64_t index, uint8_t* out, size_t n)
{
    uint64_t seed = SEED0 + index;
    switch (kind) {
    case 0: gen_text(seed, out, n); break;
    case 1: gen_binary(seed, out, n); break;
    case 2: gen_random(seed, out, n); break;
    case 3: gen_tiny(seed, out, n); break;
    default: gen_runs(seed, out, n); break;
    }
}

/* container: 0 raw, 1 zlib (78 01), 2 gzip (no name), 3 gzip with FNAME "stream.bin",
 * 4 zlib with preset dictionary (78 20 + DICTID; dict/dict_len/dictid given by the caller) */
size_t sdzc_compress(const uint8_t* plain, size_t n, int level, int container, const uint8_t* dict, size_t dict_len,
                     uint32_t dictid, uint8_t* out, size_t cap)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
    if (container == 4 && dict) deflateSetDictionary(&zs, dict, (uInt)dict_len);
    size_t p = 0;
    if (container == 1) { if (cap < 2) return 0; out[p++] = 0x78; out[p++] = 0x01; }
    else if (container == 4) {
        if (cap < 6) return 0;
        out[p++] = 0x78; out[p++] = 0x20;
        out[p++] = (uint8_t)(dictid >> 24); out[p++] = (uint8_t)(dictid >> 16); out[p++] = (uint8_t)(dictid >> 8); out[p++] = (uint8_t)dictid;
    } else if (container == 2 || container == 3) {
        static const uint8_t hdr[10] = { 0x1f, 0x8b, 8, 0, 0x5e, 0x1b, 0x21, 0x5d, 0, 0xff };   /* MTIME 0x5d211b5e */
        if (cap < 32) return 0;
        memcpy(out, hdr, 10); p = 10;
        if (container == 3) { out[3] = 8; memcpy(out + p, "stream.bin", 11); p += 11; }
    }
    zs.next_in = (Bytef*)plain; zs.avail_in = (uInt)n;
    zs.next_out = out + p; zs.avail_out = (uInt)(cap - p);
    int rc = deflate(&zs, Z_FINISH);
    size_t produced = (cap - p) - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return 0;
    p += produced;
    if (container == 1 || container == 4) {
        if (cap - p < 4) return 0;
        uint32_t a = (uint32_t)adler32(adler32(0, NULL, 0), plain, (uInt)n);
        out[p++] = (uint8_t)(a >> 24); out[p++] = (uint8_t)(a >> 16); out[p++] = (uint8_t)(a >> 8); out[p++] = (uint8_t)a;
    } else if (container == 2 || container == 3) {
        if (cap - p < 8) return 0;
        uint32_t c = (uint32_t)crc32(crc32(0, NULL, 0), plain, (uInt)n);
        for (int j = 0; j < 4; j++) out[p++] = (uint8_t)(c >> (8 * j));
        for (int j = 0; j < 4; j++) out[p++] = (uint8_t)((uint32_t)n >> (8 * j));
    }
    return p;
}

typedef struct {
    int kind, level, container;
    uint64_t first, lo, hi;
    uint32_t plain_len;
    uint8_t* comp; uint64_t stride; uint64_t* comp_len;
    uint8_t* plain;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    uint8_t* tmp = j->plain ? NULL : (uint8_t*)malloc(j->plain_len ? j->plain_len : 1);
    for (uint64_t i = j->lo; i < j->hi; i++) {
        uint8_t* pl = j->plain ? j->plain + i * (uint64_t)j->plain_len : tmp;
        sdzc_generate(j->kind, j->first + i, pl, j->plain_len);
        j->comp_len[i] = sdzc_compress(pl, j->plain_len, j->level, j->container, NULL, 0, 0, j->comp + i * j->stride, j->stride);
    }
    free(tmp);
    return NULL;
}

/* n streams of plain_len bytes each -> comp[i * stride ..]; comp_len[i] = 0 on overflow */
int sdzc_make_batch(int kind, uint64_t first_index, uint64_t n, uint32_t plain_len, int level, int container,
                    uint8_t* comp, uint64_t stride, uint64_t* comp_len, uint8_t* plain, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    pthread_once(&vocab_once, make_vocab);
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof *th);
    job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof *jobs);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job_t){ kind, level, container, first_index, n * (uint64_t)t / (uint64_t)n_threads,
                           n * (uint64_t)(t + 1) / (uint64_t)n_threads, plain_len, comp, stride, comp_len, plain };
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    for (uint64_t i = 0; i < n; i++) if (comp_len[i] == 0) return -1;
    return 0;
}
