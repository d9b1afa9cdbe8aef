/*
 * sdz_napi.c - thin N-API shim between the TypeScript facade (sd-zlib-b200.ts) and the C ABI of
 * libsdzcuda.so (include/sdzcuda.h).  It contains no decoding logic and no fallback: every
 * export forwards one call.
 *
 * NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: the image has no Node.js / node_api.h
 * (SURVEY F3).  Build where Node >= 18 exists:
 *     gcc -shared -fPIC -I$(node -p "require('node:path').dirname(process.execPath)+'/../include/node'") \
 *         -I../../include sdz_napi.c -L../csrc -lsdzcuda -o sdz_napi.node
 *
 * Exports (all synchronous, like the reference API):
 *   adler32(buf: Uint8Array, seed: number): number          -> sdz_adler32      (src/adler32.ts:17)
 *   crc32(buf: Uint8Array, seed: number): number            -> sdz_crc32        (src/crc32.ts:17)
 *   inflateSizes(bufs, dicts, modes): Float64Array          -> sdz_inflate_sizes
 *   inflateBatch(bufs, dicts, modes, out: Uint8Array, outOff: BigUint64Array, outCap: BigUint64Array,
 *                records: Uint8Array /* n * 72 bytes, struct sdz_result * /): number
 *                                                           -> sdz_inflate_batch (src/sd-inflate.ts:54-228)
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "sdzcuda.h"

static sdz_ctx* g_ctx;

/* SDZ_DEVICES=0,1,2,...: a multi-device context - sdz_inflate_batch partitions every batch per stream over the listed
 * B200s (include/sdzcuda.h, sdz_ctx_create_multi); one entry or SDZ_DEVICE=n: a single device */
static sdz_ctx* ctx_or_throw(napi_env env)
{
    if (!g_ctx) {
        int devices[64], ndev = 0, rc;
        const char* list = getenv("SDZ_DEVICES");
        if (list) {
            while (*list && ndev < 64) {
                devices[ndev++] = atoi(list);
                while (*list && *list != ',') list++;
                if (*list == ',') list++;
            }
        }
        if (ndev > 1) rc = sdz_ctx_create_multi(devices, ndev, 0, &g_ctx);
        else rc = sdz_ctx_create(ndev == 1 ? devices[0] : (getenv("SDZ_DEVICE") ? atoi(getenv("SDZ_DEVICE")) : 0), 0, &g_ctx);
        if (rc != SDZ_OK) { napi_throw_error(env, NULL, "sdzcuda: no usable B200 (sm_100a) device"); return NULL; }
    }
    return g_ctx;
}

static int u8_arg(napi_env env, napi_value v, uint8_t** p, size_t* n)
{
    bool is_ta = false;
    napi_is_typedarray(env, v, &is_ta);
    if (!is_ta) return 0;
    napi_typedarray_type t;
    napi_value ab;
    size_t off;
    void* data;
    if (napi_get_typedarray_info(env, v, &t, n, &data, &ab, &off) != napi_ok) return 0;
    *p = (uint8_t*)data;
    return 1;
}

static napi_value checksum(napi_env env, napi_callback_info info, int is_crc)
{
    size_t argc = 2;
    napi_value argv[2], out;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    uint8_t* p; size_t n; int32_t seed = 0, v = 0;
    if (!u8_arg(env, argv[0], &p, &n)) { napi_throw_type_error(env, NULL, "source must be a BufferSource"); return NULL; }
    napi_get_value_int32(env, argv[1], &seed);              /* ToInt32: signed or unsigned seeds alike */
    sdz_ctx* ctx = ctx_or_throw(env);
    if (!ctx) return NULL;
    int rc = is_crc ? sdz_crc32(ctx, p, n, seed, 0, &v) : sdz_adler32(ctx, p, n, seed, 0, &v);
    if (rc != SDZ_OK) { napi_throw_error(env, NULL, sdz_last_error(ctx)); return NULL; }
    napi_create_int32(env, v, &out);
    return out;
}
static napi_value js_adler32(napi_env env, napi_callback_info info) { return checksum(env, info, 0); }
static napi_value js_crc32(napi_env env, napi_callback_info info) { return checksum(env, info, 1); }

/* bufs: Uint8Array[], dicts: (Uint8Array | undefined)[], modes: Uint8Array -> sdz_in[] */
static sdz_in* gather_inputs(napi_env env, napi_value bufs, napi_value dicts, napi_value modes, uint32_t* n_out)
{
    uint32_t n = 0;
    napi_get_array_length(env, bufs, &n);
    uint8_t* mode_p; size_t mode_n;
    if (!u8_arg(env, modes, &mode_p, &mode_n) || mode_n < n) return NULL;
    sdz_in* in = (sdz_in*)calloc(n ? n : 1, sizeof *in);
    for (uint32_t i = 0; i < n; i++) {
        napi_value b, d;
        napi_get_element(env, bufs, i, &b);
        uint8_t* p; size_t len;
        if (!u8_arg(env, b, &p, &len)) { free(in); return NULL; }
        in[i].data = p; in[i].len = len; in[i].mode = mode_p[i];
        napi_get_element(env, dicts, i, &d);
        uint8_t* dp; size_t dl;
        if (u8_arg(env, d, &dp, &dl)) { in[i].dict = dl ? dp : (const uint8_t*)""; in[i].dict_len = (uint32_t)dl; }
    }
    *n_out = n;
    return in;
}

static napi_value js_inflate_sizes(napi_env env, napi_callback_info info)
{
    size_t argc = 3;
    napi_value argv[3];
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    uint32_t n;
    sdz_in* in = gather_inputs(env, argv[0], argv[1], argv[2], &n);
    if (!in) { napi_throw_type_error(env, NULL, "data must be an ArrayBuffer or buffer view"); return NULL; }
    sdz_ctx* ctx = ctx_or_throw(env);
    if (!ctx) { free(in); return NULL; }
    uint64_t* sizes = (uint64_t*)calloc(n ? n : 1, sizeof *sizes);
    int rc = sdz_inflate_sizes(ctx, in, n, sizes, SDZ_PARITY_REFERENCE);
    free(in);
    if (rc != SDZ_OK) { free(sizes); napi_throw_error(env, NULL, sdz_last_error(ctx)); return NULL; }
    napi_value ab, out;
    double* dst;
    napi_create_arraybuffer(env, n * sizeof(double), (void**)&dst, &ab);
    for (uint32_t i = 0; i < n; i++) dst[i] = (double)sizes[i];
    free(sizes);
    napi_create_typedarray(env, napi_float64_array, n, ab, 0, &out);
    return out;
}

static napi_value js_inflate_batch(napi_env env, napi_callback_info info)
{
    size_t argc = 7;
    napi_value argv[7], out;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    uint32_t n;
    sdz_in* in = gather_inputs(env, argv[0], argv[1], argv[2], &n);
    if (!in) { napi_throw_type_error(env, NULL, "data must be an ArrayBuffer or buffer view"); return NULL; }
    uint8_t *arena, *off, *cap, *rec;
    size_t arena_n, off_n, cap_n, rec_n;
    if (!u8_arg(env, argv[3], &arena, &arena_n) || !u8_arg(env, argv[4], &off, &off_n) ||
        !u8_arg(env, argv[5], &cap, &cap_n) || !u8_arg(env, argv[6], &rec, &rec_n) || rec_n < n * sizeof(sdz_result)) {
        free(in);
        napi_throw_type_error(env, NULL, "bad output arguments");
        return NULL;
    }
    sdz_ctx* ctx = ctx_or_throw(env);
    if (!ctx) { free(in); return NULL; }
    int rc = sdz_inflate_batch(ctx, in, n, arena, (const uint64_t*)off, (const uint64_t*)cap, (sdz_result*)rec, SDZ_PARITY_REFERENCE);
    free(in);
    if (rc != SDZ_OK && rc != SDZ_E_OUT_CAP) { napi_throw_error(env, NULL, sdz_last_error(ctx)); return NULL; }
    napi_create_int32(env, rc, &out);
    return out;
}

/* ---- class Inflater over several append() calls: one sdz_inflater per JS object (an external wrapped by the facade)
 *   inflaterNew(raw: boolean, dict: Uint8Array | undefined): External
 *   inflaterAppend(h, chunk: Uint8Array, record: Uint8Array /* 72 bytes * /): Uint8Array   (the bytes this call produced)
 *   inflaterFinish(h, record: Uint8Array): Uint8Array                                      (the gzip FNAME bytes)        */
static void inflater_gc(napi_env env, void* data, void* hint) { (void)env; (void)hint; sdz_inflater_destroy((sdz_inflater*)data); }

static napi_value js_inflater_new(napi_env env, napi_callback_info info)
{
    size_t argc = 2;
    napi_value argv[2], out;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    bool raw = false;
    napi_get_value_bool(env, argv[0], &raw);
    uint8_t* dp = NULL; size_t dl = 0;
    const int has_dict = u8_arg(env, argv[1], &dp, &dl);
    sdz_ctx* ctx = ctx_or_throw(env);
    if (!ctx) return NULL;
    sdz_inflater* h = NULL;
    int rc = sdz_inflater_create(ctx, raw, has_dict ? (dl ? dp : (const uint8_t*)"") : NULL, (uint32_t)dl, &h);
    if (rc != SDZ_OK) { napi_throw_error(env, NULL, sdz_last_error(ctx)); return NULL; }
    napi_create_external(env, h, inflater_gc, NULL, &out);
    return out;
}

static napi_value js_inflater_append(napi_env env, napi_callback_info info)
{
    size_t argc = 3;
    napi_value argv[3], ab, out;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    sdz_inflater* h = NULL;
    napi_get_value_external(env, argv[0], (void**)&h);
    uint8_t *p, *rec; size_t n, rec_n;
    if (!u8_arg(env, argv[1], &p, &n)) { napi_throw_type_error(env, NULL, "data must be an ArrayBuffer or buffer view"); return NULL; }
    if (!u8_arg(env, argv[2], &rec, &rec_n) || rec_n < sizeof(sdz_result)) { napi_throw_type_error(env, NULL, "bad record buffer"); return NULL; }
    uint64_t fresh = 0;
    int rc = sdz_inflater_append(h, p, n, &fresh, (sdz_result*)rec);
    if (rc != SDZ_OK) { napi_throw_error(env, NULL, sdz_last_error(g_ctx)); return NULL; }
    void* dst;
    napi_create_arraybuffer(env, (size_t)fresh, &dst, &ab);
    if (fresh && sdz_inflater_read(h, (uint8_t*)dst, fresh) != SDZ_OK) { napi_throw_error(env, NULL, sdz_last_error(g_ctx)); return NULL; }
    napi_create_typedarray(env, napi_uint8_array, (size_t)fresh, ab, 0, &out);
    return out;
}

static napi_value js_inflater_finish(napi_env env, napi_callback_info info)
{
    size_t argc = 2;
    napi_value argv[2], ab, out;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    sdz_inflater* h = NULL;
    napi_get_value_external(env, argv[0], (void**)&h);
    uint8_t* rec; size_t rec_n;
    if (!u8_arg(env, argv[1], &rec, &rec_n) || rec_n < sizeof(sdz_result)) { napi_throw_type_error(env, NULL, "bad record buffer"); return NULL; }
    sdz_result* r = (sdz_result*)rec;
    sdz_inflater_finish(h, r);
    void* dst;
    napi_create_arraybuffer(env, r->name_len, &dst, &ab);
    if (r->name_len) sdz_inflater_input(h, r->name_off, r->name_len, (uint8_t*)dst);
    napi_create_typedarray(env, napi_uint8_array, r->name_len, ab, 0, &out);
    return out;
}

static napi_value init(napi_env env, napi_value exports)
{
    napi_property_descriptor props[] = {
        { "adler32", NULL, js_adler32, NULL, NULL, NULL, napi_default, NULL },
        { "crc32", NULL, js_crc32, NULL, NULL, NULL, napi_default, NULL },
        { "inflateSizes", NULL, js_inflate_sizes, NULL, NULL, NULL, napi_default, NULL },
        { "inflateBatch", NULL, js_inflate_batch, NULL, NULL, NULL, napi_default, NULL },
        { "inflaterNew", NULL, js_inflater_new, NULL, NULL, NULL, napi_default, NULL },
        { "inflaterAppend", NULL, js_inflater_append, NULL, NULL, NULL, napi_default, NULL },
        { "inflaterFinish", NULL, js_inflater_finish, NULL, NULL, NULL, napi_default, NULL },
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, init)
