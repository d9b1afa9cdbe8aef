// sdzcuda.cu - host side of libsdzcuda.so (C ABI declared in include/sdzcuda.h).
// Owns the CUDA context objects, stages batches into HBM, launches the sm_100a kernels and
// returns the reference-identical records.  There is no CPU decode path in this file.
#include "../../include/sdzcuda.h"

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "checksum_kernels.cuh"
#include "inflate_kernel.cuh"
#include "fast_kernels.cuh"
#include "large_kernels.cuh"

// (after the kernels: unistd.h defines R_OK as a macro)
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct sdz_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;        // compute stream (kernels; also the timing events)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined host path
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    std::vector<cudaEvent_t> pipe_ev;    // 2 per sub-batch: input landed, kernels done
    std::string err;
    uint64_t launches = 0;
    float last_ms[3] = { 0, 0, 0 };
    // grow-only scratch
    DevBuf d_in, d_out, d_meta, d_res, d_part, d_misc, d_sym, d_task;
    // buffers of the last closed large-stream session (symbols, task arrays, descriptors): the next session takes them
    // over instead of paying cudaMalloc / cudaFree of gigabytes per call
    DevBuf large_cache[3];
    bool large_cache_full = false;
    void* h_stage = nullptr;           // pinned
    size_t h_stage_cap = 0;
    void* h_out = nullptr;             // pinned landing zone of the decoded bytes when the caller's arena is pageable: the
    size_t h_out_cap = 0;              // device -> host copy stays asynchronous, host threads move the bytes on
    void* h_res = nullptr;             // pinned landing zone of the result records (a copy into the caller's pageable
    size_t h_res_cap = 0;              // array would block the host thread and serialise the pipeline)
    unsigned long long* d_counter = nullptr;   // [0..3] lane 0 / checksums / large-stream path, [4 + 2 l ..] lane l
    // compute lanes of the pipelined host path: sub-batch kernels of different lanes may run concurrently,
    // so each lane has its own stream, work counters and table scratch (lane 0 = stream / d_misc)
    static constexpr int N_LANES = 3;
    cudaStream_t lane_stream[N_LANES] = { nullptr, nullptr, nullptr };
    DevBuf lane_scratch[N_LANES];
    int cur_lane = 0;
    int group = 4;                     // lanes per stream
    int block_threads = 64;
    // two-phase fast path (fast_kernels.cuh): per compute lane the token arena, the per-stream arrays
    // (counters | tok_off | ntok | hand-over list) and the sorted-symbol scratch of phase A
    bool fast = true;                  // SDZ_FAST=0: every stream goes through the general decoder
    int b_blocks_per_sm = 8;           // byte-centric phase B (SDZ_B2=0): 256-thread blocks per SM (SDZ_B_BLOCKS)
    bool b_tokenwise = true;           // token-centric phase B (lz_resolve2_kernel)
    int b2_blocks_per_sm = 12;         // its 128-thread blocks per SM (SDZ_B2_BLOCKS)
    static constexpr int MAX_FAST_CHUNKS = 15;
    int fast_chunks = 0;               // chunks of the phase A / phase B pipeline (SDZ_FAST_CHUNKS; 0 = one per wave of phase A)
    cudaStream_t fast_sb[N_LANES] = { nullptr, nullptr, nullptr };     // phase B streams
    cudaEvent_t fast_ev[N_LANES][MAX_FAST_CHUNKS + 1] = {};
    DevBuf fast_tok[N_LANES], fast_meta[N_LANES], fast_sorted[N_LANES];
    cudaEvent_t ev_fast[4] = { nullptr, nullptr, nullptr, nullptr };   // before A, after A, after B (joined), after the hand-over run (which overlaps B)
    bool fast_timed = false;
    const unsigned long long* last_fb_count = nullptr;   // device counter of the most recent fast-path launch
    uint64_t last_fast_n = 0;
    bool h2d_first = false;            // SDZ_H2D_FIRST=1: the pipelined host path sends every input before it fetches any output
    bool poison = false;               // SDZ_POISON=1 (tests): fill the device output arena with 0xA5 before every decode
    // multi-device context (sdz_ctx_create_multi): one child context per entry of the device list; this object then only
    // partitions batches (inflate_multi) and owns no streams of its own.  numa_node: host memory node next to the device
    // (-1 unknown), used for the pinned staging buffers and sdz_host_alloc_near.
    std::vector<sdz_ctx*> peers;
    std::vector<uint64_t> last_cut;    // partition of the most recent batch: peer d decoded streams [last_cut[d], last_cut[d + 1])
    int numa_node = -1;
};

namespace {

// entry points other than the batched host path run on the first device of a multi-device context
#define ENTER(ctx)                                                                                 \
    do {                                                                                           \
        if (!(ctx)->peers.empty()) (ctx) = (ctx)->peers[0];                                        \
        CK(cudaSetDevice((ctx)->device));                                                          \
    } while (0)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return SDZ_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

int grow(sdz_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return SDZ_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = (bytes + (1u << 20)) & ~((size_t(1) << 20) - 1);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); b.p = nullptr; return SDZ_E_NOMEM; }
    b.cap = want;
    return SDZ_OK;
}

// ---- pinned host memory next to a device.  cudaMallocHost places pages wherever the calling thread happens to run; on a
// two-socket box that puts half of the ranks' staging buffers across the socket interconnect from their GPU.  These
// allocations are mmap'ed, bound to the device's NUMA node with mbind(MPOL_PREFERRED) (raw syscall: no libnuma in the
// image) and then registered with CUDA (portable: every device may DMA from them).
std::mutex g_pin_mu;
std::map<void*, size_t> g_pin_regions;

int device_numa_node(int device)
{
    char id[32] = { 0 };
    if (cudaDeviceGetPCIBusId(id, sizeof id, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = id; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

void* pinned_alloc(size_t bytes, int node)
{
    if (bytes == 0) bytes = 1;
    const size_t page = 4096, len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
#ifdef SYS_mbind
    if (node >= 0 && node < 64) {
        unsigned long mask = 1ul << node;
        syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u);       // best effort
    }
#endif
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, len);
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pin_regions[p] = len;
    return p;
}

void pinned_free(void* p)
{
    if (!p) return;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        auto it = g_pin_regions.find(p);
        if (it != g_pin_regions.end()) { len = it->second; g_pin_regions.erase(it); }
    }
    if (len) { cudaHostUnregister(p); munmap(p, len); }
    else cudaFreeHost(p);
}

// run the calling thread on the cores of a NUMA node (best effort; the copy threads of a device sit next to its memory)
void bind_thread_to_node(int node)
{
    if (node < 0) return;
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return;
    char buf[1024] = { 0 };
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int any = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); any = 1; }
    }
    if (any) sched_setaffinity(0, sizeof set, &set);
}

int grow_stage(sdz_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->h_stage_cap) return SDZ_OK;
    if (ctx->h_stage) { pinned_free(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_stage_cap = 0; }
    size_t want = (bytes + (1u << 20)) & ~((size_t(1) << 20) - 1);
    ctx->h_stage = pinned_alloc(want, ctx->numa_node);
    if (!ctx->h_stage) { ctx->err = "pinned host allocation failed (staging)"; return SDZ_E_NOMEM; }
    ctx->h_stage_cap = want;
    return SDZ_OK;
}

int grow_res(sdz_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->h_res_cap) return SDZ_OK;
    if (ctx->h_res) { pinned_free(ctx->h_res); ctx->h_res = nullptr; ctx->h_res_cap = 0; }
    size_t want = (bytes + (1u << 20)) & ~((size_t(1) << 20) - 1);
    ctx->h_res = pinned_alloc(want, ctx->numa_node);
    if (!ctx->h_res) { ctx->err = "pinned host allocation failed (records)"; return SDZ_E_NOMEM; }
    ctx->h_res_cap = want;
    return SDZ_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- GF(2) helpers on the host (table generation only)
uint32_t h_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ sdz::CRC_POLY : (b >> 1);
    }
    return p;
}

int upload_tables(sdz_ctx* ctx)
{
    static uint32_t tab[4][256], stride[4][256], x2n[32];
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (sdz::CRC_POLY ^ (c >> 1)) : (c >> 1);
        tab[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = tab[0][n];
        for (int k = 1; k < 4; k++) { c = tab[0][c & 0xff] ^ (c >> 8); tab[k][n] = c; }
    }
    uint32_t p = 1u << 30;
    x2n[0] = p;
    for (int n = 1; n < 32; n++) x2n[n] = p = h_mulmod(p, p);
    // multiply-by-x^(8*512) per register byte: advance the register over 512 zero bytes
    for (int j = 0; j < 4; j++)
        for (uint32_t b = 0; b < 256; b++) {
            uint32_t c = b << (8 * j);
            for (uint32_t i = 0; i < sdz::CRC_ROW; i++) c = tab[0][c & 0xff] ^ (c >> 8);
            stride[j][b] = c;
        }
    CK(cudaMemcpyToSymbol(sdz::g_crc_tab, tab, sizeof tab));
    CK(cudaMemcpyToSymbol(sdz::g_crc_stride_tab, stride, sizeof stride));
    CK(cudaMemcpyToSymbol(sdz::c_x2n, x2n, sizeof x2n));
    return SDZ_OK;
}

template <int G, bool STORE, int TM = sdz::TM_NONE>
int launch_inflate_t(sdz_ctx* ctx, const sdz::InflateParams& P)
{
    const int threads = ctx->block_threads;
    const int groups = threads / G;
    const size_t smem = (size_t)groups * sizeof(sdz::GroupSmem);
    auto kern = sdz::inflate_kernel<G, STORE, TM>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long want = (P.n + groups - 1) / groups;
    unsigned long long grid = std::min<unsigned long long>(want, (unsigned long long)ctx->sm_count * per_sm);
    if (grid == 0) return SDZ_OK;
    const int lane = ctx->cur_lane;
    DevBuf& scratch = lane ? ctx->lane_scratch[lane] : ctx->d_misc;
    cudaStream_t st = ctx->lane_stream[lane];
    {
        int rc = grow(ctx, scratch, (size_t)grid * groups * sdz::SCRATCH_U16 * sizeof(uint16_t));
        if (rc) return rc;
    }
    sdz::InflateParams Q = P;
    Q.scratch = (uint16_t*)scratch.p;
    Q.counter = lane ? ctx->d_counter + 4 + 2 * lane : ctx->d_counter;
    CK(cudaMemsetAsync(Q.counter, 0, sizeof(unsigned long long), st));
    kern<<<(unsigned)grid, threads, smem, st>>>(Q);
    ctx->launches++;
    CK(cudaGetLastError());
    return SDZ_OK;
}

template <bool STORE>
int launch_inflate(sdz_ctx* ctx, const sdz::InflateParams& P)
{
    switch (ctx->group) {
    case 2: return launch_inflate_t<2, STORE>(ctx, P);
    case 8: return launch_inflate_t<8, STORE>(ctx, P);
    case 32: return launch_inflate_t<32, STORE>(ctx, P);
    default: return launch_inflate_t<4, STORE>(ctx, P);
    }
}

int launch_finalize(sdz_ctx* ctx, const uint8_t* d_out, sdz_result* d_res, uint64_t n, bool spec = false)
{
    if (n == 0) return SDZ_OK;
    const int lane = ctx->cur_lane;
    unsigned long long* counter = lane ? ctx->d_counter + 4 + 2 * lane + 1 : ctx->d_counter + 1;
    cudaStream_t st = ctx->lane_stream[lane];
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    unsigned long long warps = n;
    unsigned grid = (unsigned)std::min<unsigned long long>((warps + 7) / 8, (unsigned long long)ctx->sm_count * 8);
    sdz::finalize_streams_kernel<<<grid, 256, 0, st>>>(d_out, d_res, n, counter, spec ? 1u : 0u);
    ctx->launches++;
    CK(cudaGetLastError());
    return SDZ_OK;
}

// Two-phase fast path: token offsets -> phase A (Huffman -> tokens) -> phase B (tokens -> bytes) -> the general
// decoder for the streams phase A handed over.  `tok_total`: size of the token arena when the caller knows it
// (host path: computed from the host copies of in_len / out_cap), 0 = read it back from the device.
// Grow every device buffer a sub-batch of `n` streams / `tok_total` token slots will need on compute lane `lane` BEFORE the
// pipeline starts: cudaMalloc / cudaFree synchronise the device, and a first call that grows its buffers between
// sub-batches stalls every stream in flight (round 1: 730 ms at sub-batch 4 of the first call).
int reserve_lane(sdz_ctx* ctx, int lane, uint64_t n, uint64_t tok_total, bool fast)
{
    int rc;
    {
        const int threads = ctx->block_threads, groups = threads / ctx->group;
        // (upper bound of the persistent grid of inflate_kernel: 16 blocks per SM is more than its shared memory allows)
        const uint64_t grid = std::min<uint64_t>((n + groups - 1) / groups, (uint64_t)ctx->sm_count * 16);
        DevBuf& scratch = lane ? ctx->lane_scratch[lane] : ctx->d_misc;
        if ((rc = grow(ctx, scratch, (size_t)grid * groups * sdz::SCRATCH_U16 * sizeof(uint16_t)))) return rc;
    }
    if (!fast) return SDZ_OK;
    const size_t off_list = 256 + (n + 1) * 8 + (n * 4 + 7) / 8 * 8;
    if ((rc = grow(ctx, ctx->fast_meta[lane], off_list + align_up(n * 4, 8) + n * 4))) return rc;
    if ((rc = grow(ctx, ctx->fast_tok[lane], tok_total * 4 + 64))) return rc;
    const uint64_t grid_a = std::min<uint64_t>((n + 31) / 32, (uint64_t)ctx->sm_count * 8);
    if ((rc = grow(ctx, ctx->fast_sorted[lane], (size_t)grid_a * 32 * sdz::SORTED_L * sizeof(uint16_t)))) return rc;
    return SDZ_OK;
}

int launch_fast(sdz_ctx* ctx, const sdz::InflateParams& P, uint64_t tok_total, bool timed)
{
    const uint64_t n = P.n;
    if (n == 0) return SDZ_OK;
    if (n >= 0xffffffffull) return SDZ_E_ARG;
    const int lane = ctx->cur_lane;
    cudaStream_t st = ctx->lane_stream[lane];
    int rc;
    const size_t off_tokoff = 256, off_ntok = off_tokoff + (n + 1) * 8, off_list = off_ntok + align_up(n * 4, 8), off_order = off_list + align_up(n * 4, 8);
    if ((rc = grow(ctx, ctx->fast_meta[lane], off_order + n * 4))) return rc;
    uint8_t* fm = (uint8_t*)ctx->fast_meta[lane].p;
    unsigned long long* counters = (unsigned long long*)fm;
    uint64_t* tok_off = (uint64_t*)(fm + off_tokoff);
    CK(cudaMemsetAsync(counters, 0, 256, st));
    sdz::token_offsets_kernel<<<1, 1024, 0, st>>>(P.in_len, P.out_cap, n, tok_off, (uint32_t*)(fm + off_order));
    ctx->launches++;
    if (!tok_total) {
        CK(cudaMemcpyAsync(&tok_total, tok_off + n, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if ((rc = grow(ctx, ctx->fast_tok[lane], tok_total * 4 + 64))) return rc;

    auto kern = sdz::huff_tokens_kernel;
    const size_t smem = sdz::FA_WARP_SMEM;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) per_sm = 1;
    const unsigned grid_a = (unsigned)std::min<uint64_t>((n + 31) / 32, (uint64_t)ctx->sm_count * per_sm);
    if ((rc = grow(ctx, ctx->fast_sorted[lane], (size_t)grid_a * 32 * sdz::SORTED_L * sizeof(uint16_t)))) return rc;

    sdz::FastParams F;
    F.I = P;
    F.tokens = (uint32_t*)ctx->fast_tok[lane].p;
    F.tok_off = tok_off;
    F.ntok = (uint32_t*)(fm + off_ntok);
    F.fb_list = (uint32_t*)(fm + off_list);
    F.fb_count = counters;
    F.sorted_l = (uint16_t*)ctx->fast_sorted[lane].p;
    F.order = (const uint32_t*)(fm + off_order);
    F.dict_ok = ctx->b_tokenwise && P.dict ? 1u : 0u;
    // The batch is cut into chunks: phase A of chunk c + 1 (latency-bound, one warp per scheduler, all of the shared
    // memory) runs next to phase B of chunk c (issue-bound, no shared memory) on a second stream, so the two kernels
    // fill each other's idle issue slots on the same SMs.
    // A chunk is at most ONE wave of phase A (every lane of the grid gets one stream): phase A's time per launch is
    // the time of its slowest lane, so a chunk that gives some lanes a second stream takes twice as long, and a chunk
    // much smaller than a wave leaves lanes idle for the whole launch (measured: 4 chunks 25.4 ms, 8 chunks 43.7 ms).
    const uint64_t wave = (uint64_t)grid_a * 32;
    unsigned n_chunks = (unsigned)std::min<uint64_t>((n + wave - 1) / wave, sdz_ctx::MAX_FAST_CHUNKS);
    if (ctx->fast_chunks > 0) n_chunks = (unsigned)std::min<uint64_t>((uint64_t)ctx->fast_chunks, std::max<uint64_t>(1, n / 2048));
    if (n_chunks < 1) n_chunks = 1;
    cudaStream_t sb = ctx->fast_sb[lane];
    if (timed) CK(cudaEventRecord(ctx->ev_fast[0], st));
    for (unsigned c = 0; c < n_chunks; c++) {
        const uint64_t lo = n * c / n_chunks, hi = n * (c + 1) / n_chunks;
        if (lo == hi) continue;
        F.first = (uint32_t)lo; F.count = (uint32_t)(hi - lo);
        F.counter_a = counters + 1 + 2 * c; F.counter_b = counters + 2 + 2 * c;
        const unsigned ga = (unsigned)std::min<uint64_t>((hi - lo + 31) / 32, grid_a);
        kern<<<ga, 32, smem, st>>>(F);
        CK(cudaEventRecord(ctx->fast_ev[lane][c], st));
        CK(cudaStreamWaitEvent(sb, ctx->fast_ev[lane][c], 0));
        if (c == n_chunks - 1) {
            // hand-over run: the general decoder over the list (its length is only known on the device).  The list is
            // complete when the last phase A is, and the streams on it are none of phase B's, so the run is queued right
            // behind phase A, AHEAD of the last phase B: its blocks take their share of the SMs first (an empty list costs
            // nothing, its blocks leave at once) and phase B fills the rest - a handful of handed-over streams are decoded
            // at the latency of ONE group (~ 0.4 ms per KiB of output), and that tail used to start after phase B.
            if (timed) CK(cudaEventRecord(ctx->ev_fast[1], st));
            sdz::InflateParams Q = P;
            Q.list = F.fb_list;
            Q.n_dev = F.fb_count;
            rc = launch_inflate<true>(ctx, Q);
            if (rc) return rc;
            if (timed) CK(cudaEventRecord(ctx->ev_fast[3], st));
        }
        if (ctx->b_tokenwise) {
            const unsigned gb = (unsigned)std::min<uint64_t>((hi - lo + sdz::B2_WARPS - 1) / sdz::B2_WARPS, (uint64_t)ctx->sm_count * ctx->b2_blocks_per_sm);
            if (P.dict) sdz::lz_resolve2_kernel<true><<<gb, 32 * sdz::B2_WARPS, 0, sb>>>(F);
            else sdz::lz_resolve2_kernel<false><<<gb, 32 * sdz::B2_WARPS, 0, sb>>>(F);
        } else {
            const unsigned gb = (unsigned)std::min<uint64_t>((hi - lo + 7) / 8, (uint64_t)ctx->sm_count * ctx->b_blocks_per_sm);
            sdz::lz_resolve_kernel<<<gb, 256, 0, sb>>>(F);
        }
        ctx->launches += 2;
    }
    CK(cudaEventRecord(ctx->fast_ev[lane][sdz_ctx::MAX_FAST_CHUNKS], sb));
    CK(cudaGetLastError());
    CK(cudaStreamWaitEvent(st, ctx->fast_ev[lane][sdz_ctx::MAX_FAST_CHUNKS], 0));
    if (timed) CK(cudaEventRecord(ctx->ev_fast[2], st));
    ctx->fast_timed = timed;
    ctx->last_fb_count = F.fb_count;
    ctx->last_fast_n = n;
    return SDZ_OK;
}

int run_batch_device(sdz_ctx* ctx, const sdz_batch_dev* b, bool sizes_only, bool first = true, bool last = true, uint64_t tok_total = 0,
                     bool spec = false)
{
    sdz::InflateParams P;
    memset(&P, 0, sizeof P);
    P.in = b->d_in; P.in_off = b->d_in_off; P.in_len = b->d_in_len; P.mode = b->d_mode;
    P.dict = b->d_dict; P.dict_off = b->d_dict_off; P.dict_len = b->d_dict_len; P.dict_adler = b->d_dict_adler;
    P.out = sizes_only ? nullptr : b->d_out; P.out_off = b->d_out_off; P.out_cap = b->d_out_cap;
    P.res = b->d_results; P.n = b->n; P.counter = ctx->d_counter; P.scratch = nullptr;
    P.spec = spec ? 1u : 0u;                             // (the two-phase fast path is reference-exact only)
    cudaStream_t st = ctx->lane_stream[ctx->cur_lane];
    if (first) CK(cudaEventRecord(ctx->ev[0], st));
    ctx->fast_timed = false;
    int rc = sizes_only ? launch_inflate<false>(ctx, P) : (ctx->fast && !spec ? launch_fast(ctx, P, tok_total, first && last) : launch_inflate<true>(ctx, P));
    if (rc) return rc;
    if (last) CK(cudaEventRecord(ctx->ev[1], st));
    if (!sizes_only) {
        rc = launch_finalize(ctx, b->d_out, b->d_results, b->n, spec);
        if (rc) return rc;
    }
    if (last) CK(cudaEventRecord(ctx->ev[2], st));
    return SDZ_OK;
}

void parallel_copy(const std::vector<std::pair<uint8_t*, std::pair<const uint8_t*, size_t>>>& jobs)
{
    size_t total = 0;
    for (auto& j : jobs) total += j.second.second;
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    if (total < (8u << 20) || nt == 1) {
        for (auto& j : jobs) memcpy(j.first, j.second.first, j.second.second);
        return;
    }
    std::atomic<size_t> next{ 0 };
    std::vector<std::thread> th;
    const size_t step = std::max<size_t>(1, std::min<size_t>(256, jobs.size() / (nt * 8)));
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t lo = next.fetch_add(step);
                if (lo >= jobs.size()) break;
                size_t hi = std::min(jobs.size(), lo + step);
                for (size_t i = lo; i < hi; i++) memcpy(jobs[i].first, jobs[i].second.first, jobs[i].second.second);
            }
        });
    for (auto& t : th) t.join();
}

}  // namespace

// ---- white-box test entry: the reference-table geometry the kernels re-derive in closed form
namespace sdz {
template <int G>
__global__ void __launch_bounds__(128) debug_table_totals_kernel(const uint8_t* lens, const int32_t* nl, const int32_t* nd, unsigned long long n,
                                                                  int32_t* out)
{
    __shared__ uint8_t s_lens[128 / G][320];
    __shared__ uint16_t s_cnt[128 / G][16];
    __shared__ uint32_t s_aux[128 / G][16];
    const int gid = threadIdx.x / G, glane = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - glane));
    const unsigned long long groups = (unsigned long long)gridDim.x * (128 / G);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (128 / G) + gid; i < (n + groups - 1) / groups * groups; i += groups) {
        const bool on = i < n;                           // (all groups of a warp run the same number of rounds)
        const int a = on ? nl[i] : 257, b = on ? nd[i] : 1;
        for (int k = glane; k < 320; k += G) s_lens[gid][k] = on ? lens[i * 320 + k] : (uint8_t)(k < 2 || k == 257 ? 1 : 0);
        __syncwarp(gmask);
        int res[4];
        for (int t = 0; t < 2; t++) {
            int l = 0, g = 0, pad = 0, nz = 0;
            const int st = classify<G>(s_lens[gid] + (t ? a : 0), t ? b : a, t ? 6 : 9, s_cnt[gid], s_aux[gid], &l, &g, &pad, &nz, glane, gmask);
            res[2 * t] = st;
            res[2 * t + 1] = (st == 0 || st == 2) ? ref_table_total<G>(s_cnt[gid], g, pad, l, glane, gmask) : 0;
            __syncwarp(gmask);
        }
        if (on && glane == 0) for (int k = 0; k < 4; k++) out[i * 4 + k] = res[k];
    }
}
}  // namespace sdz

// ============================================================================ C ABI

extern "C" {

const char* sdz_version(void) { return "sdzcuda 0.1 (sm_100a)"; }

int sdz_ctx_create(int device, uint32_t flags, sdz_ctx** out)
{
    (void)flags;
    if (!out) return SDZ_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SDZ_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return SDZ_E_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SDZ_E_NO_DEVICE;
    if (prop.major != 10) return SDZ_E_NO_DEVICE;          // the kernels are built for sm_100a only
    sdz_ctx* ctx = new sdz_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->numa_node = getenv("SDZ_NO_NUMA") ? -1 : device_numa_node(device);
    auto fail = [&](int rc) { sdz_ctx_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(SDZ_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(SDZ_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) != cudaSuccess) return fail(SDZ_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) != cudaSuccess) return fail(SDZ_E_CUDA);
    for (auto& e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return fail(SDZ_E_CUDA);
    for (auto& e : ctx->ev_fast)
        if (cudaEventCreate(&e) != cudaSuccess) return fail(SDZ_E_CUDA);
    if (const char* f = getenv("SDZ_FAST")) ctx->fast = atoi(f) != 0;
    if (const char* f = getenv("SDZ_B_BLOCKS")) { int v = atoi(f); if (v >= 1 && v <= 8) ctx->b_blocks_per_sm = v; }
    if (const char* f = getenv("SDZ_B2")) ctx->b_tokenwise = atoi(f) != 0;
    if (const char* f = getenv("SDZ_B2_BLOCKS")) { int v = atoi(f); if (v >= 1 && v <= 16) ctx->b2_blocks_per_sm = v; }
    if (const char* f = getenv("SDZ_FAST_CHUNKS")) { int v = atoi(f); if (v >= 0 && v <= sdz_ctx::MAX_FAST_CHUNKS) ctx->fast_chunks = v; }
    for (int l = 0; l < sdz_ctx::N_LANES; l++) {
        if (cudaStreamCreateWithFlags(&ctx->fast_sb[l], cudaStreamNonBlocking) != cudaSuccess) return fail(SDZ_E_CUDA);
        for (auto& e : ctx->fast_ev[l])
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(SDZ_E_CUDA);
    }
    if (cudaMalloc(&ctx->d_counter, (4 + 2 * sdz_ctx::N_LANES) * sizeof(unsigned long long)) != cudaSuccess) return fail(SDZ_E_NOMEM);
    ctx->lane_stream[0] = ctx->stream;
    for (int l = 1; l < sdz_ctx::N_LANES; l++)
        if (cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking) != cudaSuccess) return fail(SDZ_E_CUDA);
    if (const char* g = getenv("SDZ_GROUP")) { int v = atoi(g); if (v == 2 || v == 4 || v == 8 || v == 32) ctx->group = v; }
    if (const char* t = getenv("SDZ_BLOCK")) { int v = atoi(t); if (v == 32 || v == 64 || v == 128) ctx->block_threads = v; }
    if (ctx->block_threads < ctx->group) ctx->block_threads = ctx->group;
    if (const char* z = getenv("SDZ_POISON")) ctx->poison = atoi(z) != 0;
    if (const char* z = getenv("SDZ_H2D_FIRST")) ctx->h2d_first = atoi(z) != 0;
    int rc = upload_tables(ctx);
    if (rc) return fail(rc);
    *out = ctx;
    return SDZ_OK;
}

int sdz_ctx_create_multi(const int* devices, int ndev, uint32_t flags, sdz_ctx** out)
{
    if (!out) return SDZ_E_ARG;
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return SDZ_E_ARG;
    sdz_ctx* m = new sdz_ctx();
    m->device = devices[0];
    for (int d = 0; d < ndev; d++) {
        sdz_ctx* c = nullptr;
        const int rc = sdz_ctx_create(devices[d], flags, &c);
        if (rc) { sdz_ctx_destroy(m); return rc; }
        m->peers.push_back(c);
    }
    m->sm_count = m->peers[0]->sm_count;
    *out = m;
    return SDZ_OK;
}

int sdz_ctx_device_count(sdz_ctx* ctx) { return !ctx ? 0 : (ctx->peers.empty() ? 1 : (int)ctx->peers.size()); }

int sdz_last_partition(sdz_ctx* ctx, uint64_t* cut, int n_cut)
{
    if (!ctx || !cut) return SDZ_E_ARG;
    if (ctx->peers.empty() || ctx->last_cut.empty() || n_cut < (int)ctx->last_cut.size()) return SDZ_E_ARG;
    for (size_t i = 0; i < ctx->last_cut.size(); i++) cut[i] = ctx->last_cut[i];
    return SDZ_OK;
}

void sdz_ctx_destroy(sdz_ctx* ctx)
{
    if (!ctx) return;
    if (!ctx->peers.empty()) {
        for (sdz_ctx* c : ctx->peers) sdz_ctx_destroy(c);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : { &ctx->d_in, &ctx->d_out, &ctx->d_meta, &ctx->d_res, &ctx->d_part, &ctx->d_misc, &ctx->d_sym, &ctx->d_task,
                       &ctx->large_cache[0], &ctx->large_cache[1], &ctx->large_cache[2] })
        if (b->p) cudaFree(b->p);
    for (int l = 1; l < sdz_ctx::N_LANES; l++) {
        if (ctx->lane_stream[l]) { cudaStreamSynchronize(ctx->lane_stream[l]); cudaStreamDestroy(ctx->lane_stream[l]); }
        if (ctx->lane_scratch[l].p) cudaFree(ctx->lane_scratch[l].p);
    }
    for (int l = 0; l < sdz_ctx::N_LANES; l++)
        for (DevBuf* b : { &ctx->fast_tok[l], &ctx->fast_meta[l], &ctx->fast_sorted[l] })
            if (b->p) cudaFree(b->p);
    for (auto& e : ctx->ev_fast)
        if (e) cudaEventDestroy(e);
    for (int l = 0; l < sdz_ctx::N_LANES; l++) {
        if (ctx->fast_sb[l]) { cudaStreamSynchronize(ctx->fast_sb[l]); cudaStreamDestroy(ctx->fast_sb[l]); }
        for (auto& e : ctx->fast_ev[l])
            if (e) cudaEventDestroy(e);
    }
    pinned_free(ctx->h_stage);
    pinned_free(ctx->h_res);
    pinned_free(ctx->h_out);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->pipe_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sdz_last_error(sdz_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
uint64_t sdz_launch_count(sdz_ctx* ctx)
{
    if (!ctx) return 0;
    uint64_t v = ctx->launches;
    for (sdz_ctx* c : ctx->peers) v += c->launches;
    return v;
}

int sdz_last_timing(sdz_ctx* ctx, float ms[3])
{
    if (!ctx) return SDZ_E_ARG;
    ENTER(ctx);
    CK(cudaEventSynchronize(ctx->ev[2]));
    CK(cudaEventElapsedTime(&ms[0], ctx->ev[0], ctx->ev[1]));
    CK(cudaEventElapsedTime(&ms[1], ctx->ev[1], ctx->ev[2]));
    CK(cudaEventElapsedTime(&ms[2], ctx->ev[0], ctx->ev[2]));
    return SDZ_OK;
}

int sdz_last_phase_timing(sdz_ctx* ctx, float ms[5])
{
    if (!ctx || !ms) return SDZ_E_ARG;
    ENTER(ctx);
    CK(cudaEventSynchronize(ctx->ev[2]));
    for (int i = 0; i < 5; i++) ms[i] = 0.f;
    if (ctx->fast_timed) {
        CK(cudaEventElapsedTime(&ms[0], ctx->ev_fast[0], ctx->ev_fast[1]));
        CK(cudaEventElapsedTime(&ms[2], ctx->ev_fast[1], ctx->ev_fast[3]));      // the hand-over run starts right behind phase A ...
        CK(cudaEventElapsedTime(&ms[1], ctx->ev_fast[3], ctx->ev_fast[2]));      // ... and this is what phase B still needs after it
    } else {
        CK(cudaEventElapsedTime(&ms[2], ctx->ev[0], ctx->ev[1]));
    }
    CK(cudaEventElapsedTime(&ms[3], ctx->ev[1], ctx->ev[2]));
    CK(cudaEventElapsedTime(&ms[4], ctx->ev[0], ctx->ev[2]));
    return SDZ_OK;
}

int sdz_debug_table_totals(sdz_ctx* ctx, const uint8_t* lens, const int32_t* nl, const int32_t* nd, uint64_t n, int group, int32_t* out)
{
    if (!ctx || !lens || !nl || !nd || !out || (group != 4 && group != 32)) return SDZ_E_ARG;
    if (n == 0) return SDZ_OK;
    ENTER(ctx);
    int rc;
    if ((rc = grow(ctx, ctx->d_in, n * 320 + 16))) return rc;
    if ((rc = grow(ctx, ctx->d_meta, n * 8 + n * 16 + 16))) return rc;
    int32_t* d_nl = (int32_t*)ctx->d_meta.p;
    int32_t* d_nd = d_nl + n;
    int32_t* d_out = d_nd + n;
    CK(cudaMemcpyAsync(ctx->d_in.p, lens, n * 320, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_nl, nl, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_nd, nd, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)std::min<uint64_t>((n * group + 127) / 128, (uint64_t)ctx->sm_count * 4);
    if (group == 4) sdz::debug_table_totals_kernel<4><<<grid, 128, 0, ctx->stream>>>((const uint8_t*)ctx->d_in.p, d_nl, d_nd, n, d_out);
    else sdz::debug_table_totals_kernel<32><<<grid, 128, 0, ctx->stream>>>((const uint8_t*)ctx->d_in.p, d_nl, d_nd, n, d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

int sdz_last_fast_stats(sdz_ctx* ctx, uint64_t out[2])
{
    if (!ctx || !out) return SDZ_E_ARG;
    out[0] = out[1] = 0;
    if (!ctx->last_fb_count) return SDZ_OK;
    ENTER(ctx);
    for (int l = 0; l < sdz_ctx::N_LANES; l++) CK(cudaStreamSynchronize(ctx->lane_stream[l]));
    unsigned long long fb = 0;
    CK(cudaMemcpy(&fb, ctx->last_fb_count, sizeof fb, cudaMemcpyDeviceToHost));
    out[0] = ctx->last_fast_n - fb;
    out[1] = fb;
    return SDZ_OK;
}

void* sdz_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void sdz_host_free(void* p) { pinned_free(p); }

void* sdz_host_alloc_near(sdz_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    if (!ctx->peers.empty()) ctx = ctx->peers[0];
    cudaSetDevice(ctx->device);
    return pinned_alloc(bytes, ctx->numa_node);
}
int sdz_ctx_numa_node(sdz_ctx* ctx)
{
    if (!ctx) return -1;
    return ctx->peers.empty() ? ctx->numa_node : ctx->peers[0]->numa_node;
}

void* sdz_device_alloc(sdz_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    void* p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void sdz_device_free(sdz_ctx* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }

int sdz_memcpy_h2d(sdz_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return SDZ_E_ARG;
    ENTER(ctx);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}
int sdz_memcpy_d2h(sdz_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return SDZ_E_ARG;
    ENTER(ctx);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}
int sdz_sync(sdz_ctx* ctx)
{
    if (!ctx) return SDZ_E_ARG;
    ENTER(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

// ---------------------------------------------------------------------------- checksums

static int checksum_chain(sdz_ctx* ctx, bool crc, const uint8_t* p, const uint64_t* seg_len, uint64_t n_seg,
                          int32_t seed, int on_device, int32_t* out_values, int32_t* out_last)
{
    if (!ctx || !out_last || (n_seg && !seg_len)) return SDZ_E_ARG;
    ENTER(ctx);
    if (n_seg == 0) { *out_last = seed; return SDZ_OK; }
    if (n_seg > 0xffffffffull) return SDZ_E_ARG;
    uint64_t total = 0;
    for (uint64_t s = 0; s < n_seg; s++) {
        if (seg_len[s] >= (1ull << 32)) return SDZ_E_ARG;            // reference undefined beyond 4 GiB
        total += seg_len[s];
    }
    if (total && !p) return SDZ_E_ARG;
    const uint8_t* d_p = p;
    if (!on_device) {
        int rc = grow(ctx, ctx->d_in, total + 16);
        if (rc) return rc;
        if (total) CK(cudaMemcpyAsync(ctx->d_in.p, p, total, cudaMemcpyHostToDevice, ctx->stream));
        d_p = (const uint8_t*)ctx->d_in.p;
    }
    const uint64_t gran = crc ? sdz::CRC_TASK : sdz::ADLER_NMAX;
    // host-side segment tables: off[n], len[n], base[n+1]
    std::vector<uint64_t> tbl(3 * n_seg + 1);
    uint64_t off = 0, base = 0;
    for (uint64_t s = 0; s < n_seg; s++) {
        tbl[s] = off; tbl[n_seg + s] = seg_len[s]; tbl[2 * n_seg + s] = base;
        off += seg_len[s];
        base += (seg_len[s] + gran - 1) / gran;
    }
    tbl[3 * n_seg] = base;
    const uint64_t n_items = base;
    size_t meta_bytes = tbl.size() * 8;
    int rc = grow(ctx, ctx->d_meta, meta_bytes + n_seg * 16 + n_seg * 4 + 64);
    if (rc) return rc;
    uint8_t* dm = (uint8_t*)ctx->d_meta.p;
    CK(cudaMemcpyAsync(dm, tbl.data(), meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d_seginfo = dm + align_up(meta_bytes, 16);
    int32_t* d_values = (int32_t*)(d_seginfo + n_seg * 16);
    rc = grow(ctx, ctx->d_part, (n_items + 1) * 12);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (crc) {
        sdz::CrcTaskTable T{ (const uint64_t*)dm, (const uint64_t*)dm + n_seg, (const uint64_t*)dm + 2 * n_seg, (uint32_t)n_seg };
        uint32_t* d_partial = (uint32_t*)ctx->d_part.p;
        if (n_items) {
            const size_t smem = 4 * 256 * 32 * 4 + 4096;
            CK(cudaFuncSetAttribute(sdz::crc_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaMemsetAsync(ctx->d_counter + 2, 0, sizeof(unsigned long long), ctx->stream));
            unsigned grid = (unsigned)std::min<uint64_t>((n_items + 31) / 32, (uint64_t)ctx->sm_count);
            sdz::crc_tasks_kernel<<<grid, 1024, smem, ctx->stream>>>(d_p, T, n_items, d_partial, ctx->d_counter + 2);
            ctx->launches++;
        }
        sdz::crc_seg_kernel<<<(unsigned)n_seg, 256, 0, ctx->stream>>>(T, d_partial, (uint2*)d_seginfo);
        sdz::crc_chain_kernel<<<1, 32, 0, ctx->stream>>>((const uint2*)d_seginfo, (uint32_t)n_seg, (uint32_t)seed, d_values);
        ctx->launches += 2;
    } else {
        sdz::SegTable T{ (const uint64_t*)dm, (const uint64_t*)dm + n_seg, (const uint64_t*)dm + 2 * n_seg, (uint32_t)n_seg };
        uint2* d_partial = (uint2*)ctx->d_part.p;
        uint32_t* d_pref = (uint32_t*)((uint8_t*)ctx->d_part.p + (n_items + 1) * 8);
        if (n_items) {
            unsigned grid = (unsigned)std::min<uint64_t>((n_items + 7) / 8, (uint64_t)ctx->sm_count * 16);
            sdz::adler_units_kernel<<<grid, 256, 0, ctx->stream>>>(d_p, T, n_items, d_partial);
            ctx->launches++;
        }
        sdz::adler_seg_kernel<<<(unsigned)n_seg, 1024, 0, ctx->stream>>>(T, d_partial, d_pref, (uint4*)d_seginfo);
        sdz::adler_chain_kernel<<<1, 1024, 0, ctx->stream>>>(T, d_pref, (const uint4*)d_seginfo, (uint32_t)seed, d_values);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    std::vector<int32_t> vals(n_seg);
    CK(cudaMemcpyAsync(vals.data(), d_values, n_seg * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (out_values) memcpy(out_values, vals.data(), n_seg * 4);
    *out_last = vals[n_seg - 1];
    return SDZ_OK;
}

int sdz_adler32_chain(sdz_ctx* ctx, const uint8_t* p, const uint64_t* seg_len, uint64_t n_seg, int32_t seed,
                      int on_device, int32_t* out_values, int32_t* out_last)
{
    return checksum_chain(ctx, false, p, seg_len, n_seg, seed, on_device, out_values, out_last);
}
int sdz_crc32_chain(sdz_ctx* ctx, const uint8_t* p, const uint64_t* seg_len, uint64_t n_seg, int32_t seed,
                    int on_device, int32_t* out_values, int32_t* out_last)
{
    return checksum_chain(ctx, true, p, seg_len, n_seg, seed, on_device, out_values, out_last);
}
int sdz_adler32(sdz_ctx* ctx, const uint8_t* p, uint64_t n, int32_t seed, int on_device, int32_t* out)
{
    return checksum_chain(ctx, false, p, &n, 1, seed, on_device, nullptr, out);
}
int sdz_crc32(sdz_ctx* ctx, const uint8_t* p, uint64_t n, int32_t seed, int on_device, int32_t* out)
{
    return checksum_chain(ctx, true, p, &n, 1, seed, on_device, nullptr, out);
}

int sdz_checksum_batch(sdz_ctx* ctx, const uint8_t* const* bufs, const uint64_t* lens, const uint8_t* kind,
                       const int32_t* seeds, uint64_t n, int32_t* out)
{
    if (!ctx || (n && (!bufs || !lens || !kind || !out))) return SDZ_E_ARG;
    if (n == 0) return SDZ_OK;
    ENTER(ctx);
    std::vector<uint64_t> off(n);
    size_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (lens[i] >= (1ull << 32) || (lens[i] && !bufs[i])) return SDZ_E_ARG;
        off[i] = total;
        total += align_up(lens[i], 16);
    }
    const size_t meta = align_up(n * (8 + 8 + 4 + 4 + 1), 16);
    int rc;
    if ((rc = grow_stage(ctx, total + 16 + meta))) return rc;
    if ((rc = grow(ctx, ctx->d_in, total + 16))) return rc;
    if ((rc = grow(ctx, ctx->d_meta, meta))) return rc;
    uint8_t* hs = (uint8_t*)ctx->h_stage;
    {
        std::vector<std::pair<uint8_t*, std::pair<const uint8_t*, size_t>>> jobs;
        for (uint64_t i = 0; i < n; i++) if (lens[i]) jobs.push_back({ hs + off[i], { bufs[i], (size_t)lens[i] } });
        parallel_copy(jobs);
    }
    uint8_t* hm = hs + align_up(total + 16, 16);
    uint64_t* m_off = (uint64_t*)hm;
    uint64_t* m_len = m_off + n;
    int32_t* m_seed = (int32_t*)(m_len + n);
    int32_t* m_out = m_seed + n;
    uint8_t* m_kind = (uint8_t*)(m_out + n);
    memcpy(m_off, off.data(), n * 8);
    memcpy(m_len, lens, n * 8);
    for (uint64_t i = 0; i < n; i++) m_seed[i] = seeds ? seeds[i] : (kind[i] ? 0 : 1);
    memcpy(m_kind, kind, n);
    if (total) CK(cudaMemcpyAsync(ctx->d_in.p, hs, total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_meta.p, hm, meta, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* dm = (uint8_t*)ctx->d_meta.p;
    const uint64_t* d_off = (const uint64_t*)dm;
    const uint64_t* d_len = d_off + n;
    const int32_t* d_seed = (const int32_t*)(d_len + n);
    int32_t* d_out = (int32_t*)(d_seed + n);
    const uint8_t* d_kind = (const uint8_t*)(d_out + n);
    CK(cudaMemsetAsync(ctx->d_counter + 3, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 8);
    sdz::checksum_batch_kernel<<<grid, 256, 0, ctx->stream>>>((const uint8_t*)ctx->d_in.p, d_off, d_len, d_kind, d_seed, n, d_out,
                                                                ctx->d_counter + 3);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CK(cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

// ---------------------------------------------------------------------------- batched inflate

int sdz_inflate_batch_device(sdz_ctx* ctx, const sdz_batch_dev* batch, uint32_t flags, int sync)
{
    if (!ctx || !batch) return SDZ_E_ARG;
    ENTER(ctx);
    int rc = run_batch_device(ctx, batch, batch->d_out == nullptr, true, true, 0, (flags & SDZ_PARITY_SPEC) != 0);
    if (rc) return rc;
    if (sync) CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

static int inflate_host(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off,
                        const uint64_t* out_cap, sdz_result* results, uint64_t* out_len, uint32_t flags, bool sizes_only);

// Partition of one batch over the devices of a multi-device context ("partition streams across GPUs by greedy
// balance on compressed bytes").  Streams keep the caller's order: device d takes the contiguous range
// [cut[d], cut[d + 1]) whose compressed bytes (+ a constant per stream for the per-stream work: header, table builds,
// record) come closest to an equal share - a greedy sweep over the prefix sum.  Contiguous ranges keep the records in
// caller order and each device's output slots dense, so every device still returns its bytes with few large copies.
static void partition_streams(const sdz_in* in, uint64_t n, unsigned ndev, std::vector<uint64_t>& cut)
{
    constexpr uint64_t PER_STREAM = 512;
    cut.assign(ndev + 1, n);
    cut[0] = 0;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) total += in[i].len + PER_STREAM;
    uint64_t acc = 0, i = 0;
    for (unsigned d = 1; d < ndev; d++) {
        const uint64_t target = total / ndev * d + (total % ndev) * d / ndev;
        while (i < n && acc + (in[i].len + PER_STREAM) / 2 <= target) { acc += in[i].len + PER_STREAM; i++; }
        cut[d] = i;
    }
}

// One host thread per device; each runs the ordinary pipelined host path (staging, H2D, kernels, D2H) on its share.
// No collective and no device-to-device traffic: the streams are independent, records land in the caller's array.
static int inflate_multi(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off,
                         const uint64_t* out_cap, sdz_result* results, uint64_t* out_len, uint32_t flags, bool sizes_only)
{
    const unsigned ndev = (unsigned)ctx->peers.size();
    partition_streams(in, n, ndev, ctx->last_cut);
    const std::vector<uint64_t>& cut = ctx->last_cut;
    std::vector<int> rcs(ndev, SDZ_OK);
    std::vector<std::thread> th;
    for (unsigned d = 0; d < ndev; d++) {
        const uint64_t lo = cut[d], hi = cut[d + 1];
        if (lo == hi) continue;
        th.emplace_back([&, d, lo, hi] {
            sdz_ctx* c = ctx->peers[d];
            bind_thread_to_node(c->numa_node);
            rcs[d] = inflate_host(c, in + lo, hi - lo, out_arena, out_off ? out_off + lo : nullptr, out_cap ? out_cap + lo : nullptr,
                                  results ? results + lo : nullptr, out_len ? out_len + lo : nullptr, flags, sizes_only);
        });
    }
    for (auto& t : th) t.join();
    int ret = SDZ_OK;
    for (unsigned d = 0; d < ndev; d++) {
        if (rcs[d] == SDZ_OK) continue;
        if (rcs[d] != SDZ_E_OUT_CAP || ret == SDZ_OK) {       // a hard error wins over "some slot was too small"
            if (rcs[d] != SDZ_E_OUT_CAP) ctx->err = "device " + std::to_string(ctx->peers[d]->device) + ": " + ctx->peers[d]->err;
            if (ret == SDZ_OK || ret == SDZ_E_OUT_CAP) ret = rcs[d];
        }
    }
    return ret;
}

// shared host path: stage inputs, run, fetch records (and bytes unless sizes_only)
static int inflate_host(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off,
                        const uint64_t* out_cap, sdz_result* results, uint64_t* out_len, uint32_t flags, bool sizes_only)
{
    if (!ctx || (!in && n)) return SDZ_E_ARG;
    const bool spec = (flags & SDZ_PARITY_SPEC) != 0;
    if (!sizes_only && (!results || (n && (!out_arena || !out_off || !out_cap)))) return SDZ_E_ARG;
    if (n == 0) return SDZ_OK;
    if (!ctx->peers.empty()) return inflate_multi(ctx, in, n, out_arena, out_off, out_cap, results, out_len, flags, sizes_only);
    ENTER(ctx);

    // ---- layout of the compressed arena, the dictionary arena and the per-stream arrays
    std::vector<uint64_t> in_off(n), dict_off(n), d_out_off(n);
    std::vector<uint32_t> in_len(n), dict_len(n), d_out_cap(n);
    std::vector<uint8_t> mode(n);
    size_t in_bytes = 0, dict_bytes = 0;
    uint64_t out_lo = ~0ull, out_hi = 0;
    bool dense = true;
    for (uint64_t i = 0; i < n; i++) {
        if (in[i].len >= (1ull << 32) - 64 || (in[i].len && !in[i].data)) return SDZ_E_ARG;
        if (in[i].mode > SDZ_MODE_RAW) return SDZ_E_ARG;
        in_off[i] = in_bytes;
        in_len[i] = (uint32_t)in[i].len;
        in_bytes += align_up(in[i].len, 16);
        dict_off[i] = dict_bytes;
        dict_len[i] = in[i].dict ? in[i].dict_len : 0;
        dict_bytes += align_up(dict_len[i], 16);
        mode[i] = (uint8_t)(in[i].mode | (in[i].dict ? 0x80 : 0));
        if (!sizes_only) {
            if (out_cap[i] >= (1ull << 32)) return SDZ_E_ARG;
            out_lo = std::min(out_lo, out_off[i]);
            out_hi = std::max(out_hi, out_off[i] + out_cap[i]);
            if (i && out_off[i] != out_off[i - 1] + out_cap[i - 1]) dense = false;
        }
    }
    if (sizes_only) { out_lo = 0; out_hi = 0; }
    for (uint64_t i = 0; i < n && !sizes_only; i++) { d_out_off[i] = out_off[i] - out_lo; d_out_cap[i] = (uint32_t)out_cap[i]; }

    const size_t in_total = in_bytes + SDZ_IN_PAD;
    // meta block: in_off, dict_off, out_off (u64) | in_len, dict_len, out_cap, dict_adler (u32) | mode (u8)
    const size_t meta_bytes = align_up(n * (8 * 3 + 4 * 4 + 1), 16);
    int rc;
    if ((rc = grow_stage(ctx, in_total + dict_bytes + 16 + meta_bytes))) return rc;
    if ((rc = grow(ctx, ctx->d_in, in_total + dict_bytes + 16))) return rc;
    if ((rc = grow(ctx, ctx->d_meta, meta_bytes))) return rc;
    if ((rc = grow(ctx, ctx->d_res, n * sizeof(sdz_result)))) return rc;
    if (!sizes_only && (rc = grow(ctx, ctx->d_out, (out_hi - out_lo) + 64))) return rc;

    uint8_t* hs = (uint8_t*)ctx->h_stage;
    uint8_t* hm = hs + align_up(in_total + dict_bytes, 16);
    uint64_t* m_in_off = (uint64_t*)hm;
    uint64_t* m_dict_off = m_in_off + n;
    uint64_t* m_out_off = m_dict_off + n;
    uint32_t* m_in_len = (uint32_t*)(m_out_off + n);
    uint32_t* m_dict_len = m_in_len + n;
    uint32_t* m_out_cap = m_dict_len + n;
    int32_t* m_dict_adler = (int32_t*)(m_out_cap + n);
    uint8_t* m_mode = (uint8_t*)(m_dict_adler + n);
    memcpy(m_in_off, in_off.data(), n * 8);
    memcpy(m_dict_off, dict_off.data(), n * 8);
    memcpy(m_out_off, d_out_off.data(), n * 8);
    memcpy(m_in_len, in_len.data(), n * 4);
    memcpy(m_dict_len, dict_len.data(), n * 4);
    memcpy(m_out_cap, d_out_cap.data(), n * 4);
    memset(m_dict_adler, 0, n * 4);
    memcpy(m_mode, mode.data(), n);
    memset(hs + in_bytes, 0, SDZ_IN_PAD);

    uint8_t* d_in = (uint8_t*)ctx->d_in.p;
    // dictionaries first (small): their reference adler32 (incl. Q1) is evaluated on the device, all of them in ONE launch
    // (one warp per dictionary, checksum_batch_kernel) and one round trip
    if (dict_bytes) {
        std::vector<uint64_t> k_off, k_len;
        std::vector<uint64_t> k_idx;
        for (uint64_t i = 0; i < n; i++) {
            if (dict_len[i]) memcpy(hs + in_total + dict_off[i], in[i].dict, dict_len[i]);
            if (!(mode[i] & 0x80)) continue;
            if (spec) {                                  // RFC 1950 Adler-32 (no Q1)
                uint32_t a = 1, b2 = 0;
                for (uint64_t k = 0; k < dict_len[i]; k++) { a = (a + in[i].dict[k]) % 65521u; b2 = (b2 + a) % 65521u; }
                m_dict_adler[i] = (int32_t)((b2 << 16) | a);
                continue;
            }
            k_idx.push_back(i); k_off.push_back(dict_off[i]); k_len.push_back(dict_len[i]);
        }
        CK(cudaMemcpyAsync(d_in + in_total, hs + in_total, dict_bytes, cudaMemcpyHostToDevice, ctx->stream));
        const uint64_t k = k_idx.size();
        if (k) {
            // device scratch: off[k] | len[k] | out[k] (i32) | kind[k] (u8, all zero = adler32)
            const size_t need = k * 8 * 2 + align_up(k * 4, 8) + align_up(k, 8);
            if ((rc = grow(ctx, ctx->d_part, need + 16))) return rc;
            uint8_t* dp = (uint8_t*)ctx->d_part.p;
            uint64_t* dk_off = (uint64_t*)dp;
            uint64_t* dk_len = dk_off + k;
            int32_t* dk_out = (int32_t*)(dk_len + k);
            uint8_t* dk_kind = (uint8_t*)dk_out + align_up(k * 4, 8);
            CK(cudaMemcpyAsync(dk_off, k_off.data(), k * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(dk_len, k_len.data(), k * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemsetAsync(dk_kind, 0, k, ctx->stream));
            CK(cudaMemsetAsync(ctx->d_counter + 3, 0, sizeof(unsigned long long), ctx->stream));
            const unsigned grid = (unsigned)std::min<uint64_t>((k + 7) / 8, (uint64_t)ctx->sm_count * 8);
            sdz::checksum_batch_kernel<<<grid, 256, 0, ctx->stream>>>(d_in + in_total, dk_off, dk_len, dk_kind, nullptr, k, dk_out, ctx->d_counter + 3);
            ctx->launches++;
            CK(cudaGetLastError());
            std::vector<int32_t> vals(k);
            CK(cudaMemcpyAsync(vals.data(), dk_out, k * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (uint64_t j = 0; j < k; j++) m_dict_adler[k_idx[j]] = vals[j];
        }
    } else {
        for (uint64_t i = 0; i < n; i++) if (mode[i] & 0x80) m_dict_adler[i] = 1;   // adler32 of an empty dictionary
    }

    uint8_t* dm = (uint8_t*)ctx->d_meta.p;
    sdz_batch_dev b;
    b.d_in = d_in;
    b.d_in_off = (const uint64_t*)dm;
    b.d_dict_off = b.d_in_off + n;
    b.d_out_off = b.d_dict_off + n;
    b.d_in_len = (const uint32_t*)(b.d_out_off + n);
    b.d_dict_len = b.d_in_len + n;
    b.d_out_cap = b.d_dict_len + n;
    b.d_dict_adler = (const int32_t*)(b.d_out_cap + n);
    b.d_mode = (const uint8_t*)(b.d_dict_adler + n);
    b.d_dict = d_in + in_total;
    b.d_out = sizes_only ? nullptr : (uint8_t*)ctx->d_out.p;
    b.d_results = (sdz_result*)ctx->d_res.p;
    b.n = n;

    if ((rc = grow_res(ctx, n * sizeof(sdz_result)))) return rc;
    sdz_result* rdst = (sdz_result*)ctx->h_res;

    // ---- pipeline: the batch is cut into sub-batches that still fill the GPU; staging (host
    // threads) -> H2D (s_h2d) -> kernels (stream) -> D2H (s_d2h) of consecutive sub-batches overlap
    // Sub-batches are small (the device -> host copy of the first one starts early and the copy engine never
    // waits for a big kernel); kernels of consecutive sub-batches go to different compute lanes, so together
    // they still fill the GPU.
    // The first sub-batches are smaller still (1/4, 1/4, 1/2 of the regular size): a stream takes the same ~5 ms to decode
    // whether the GPU is full or not, so the sooner the first kernel starts and ends, the sooner the copy engine - the
    // bottleneck of the whole call - has something to send back.
    const uint64_t sub = getenv("SDZ_SUBBATCH") ? (uint64_t)atoll(getenv("SDZ_SUBBATCH")) : 4096ull;
    uint64_t K = n / (sub ? sub : 4096ull);
    if (K < 1) K = 1;
    if (K > 32) K = 32;
    if (!sizes_only && !dense) K = 1;
    std::vector<uint64_t> cut;                           // sub-batch c = streams [cut[c], cut[c + 1])
    cut.push_back(0);
    if (K >= 4 && !getenv("SDZ_EVEN_SUBBATCHES")) {
        const uint64_t reg = n / K;
        const uint64_t ramp[3] = { reg / 4, reg / 4, reg / 2 };
        for (int i = 0; i < 3; i++) cut.push_back(cut.back() + ramp[i]);
        const uint64_t rest = n - cut.back(), kr = K - 1;
        const uint64_t base = cut.back();
        for (uint64_t c = 1; c <= kr; c++) cut.push_back(base + rest * c / kr);
    } else {
        for (uint64_t c = 1; c <= K; c++) cut.push_back(n * c / K);
    }
    K = cut.size() - 1;
    const int n_lanes = K > 1 ? sdz_ctx::N_LANES : 1;
    while (ctx->pipe_ev.size() < 2 * K) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    // zero-copy fast path: the caller's buffers already sit packed (16-byte aligned spacing, the
    // layout of the device arena) in ONE pinned host region -> DMA straight from it, no staging
    bool direct = true;
    const uint8_t* ubase = nullptr;
    for (uint64_t i = 0; i < n && direct; i++) {
        if (!in_len[i]) continue;
        if (!ubase) ubase = in[i].data - in_off[i];
        if (in[i].data != ubase + in_off[i]) direct = false;
    }
    if (direct && ubase) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ubase) != cudaSuccess || attr.type != cudaMemoryTypeHost) { direct = false; cudaGetLastError(); }
    } else direct = false;
    const size_t data_end = in_off[n - 1] + in_len[n - 1];
    // pageable output arena (what an N-API caller hands over): a device -> host copy straight into it would be staged by
    // the driver, synchronously and at a fraction of the link rate.  The bytes land in a pinned zone instead, sub-batch by
    // sub-batch, and host threads move each sub-batch on while the next one is in flight.
    bool out_staged = false;
    if (!sizes_only && dense && K > 1) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out_arena + out_lo) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered) { out_staged = true; cudaGetLastError(); }
        if (out_staged) {
            if (ctx->h_out_cap < out_hi - out_lo) {
                pinned_free(ctx->h_out);
                ctx->h_out_cap = 0;
                ctx->h_out = pinned_alloc(out_hi - out_lo, ctx->numa_node);
                if (ctx->h_out) ctx->h_out_cap = out_hi - out_lo;
                else out_staged = false;                                  // no pinned memory left: the plain copy still works
            }
        }
    }
    // every buffer the sub-batches will need, before anything is in flight
    const bool use_fast = ctx->fast && !spec;
    {
        uint64_t n_max[sdz_ctx::N_LANES] = { 0 }, tok_max[sdz_ctx::N_LANES] = { 0 };
        for (uint64_t c = 0; c < K; c++) {
            const int l = (int)(c % n_lanes);
            uint64_t tt = 0;
            if (!sizes_only && use_fast)
                for (uint64_t i = cut[c]; i < cut[c + 1]; i++) tt += sdz::token_cap(in_len[i], d_out_cap[i]);
            n_max[l] = std::max(n_max[l], cut[c + 1] - cut[c]);
            tok_max[l] = std::max(tok_max[l], tt);
        }
        for (int l = 0; l < n_lanes; l++)
            if ((rc = reserve_lane(ctx, l, n_max[l], tok_max[l], !sizes_only && use_fast))) return rc;
    }

    if (ctx->poison && !sizes_only) {
        // stale bytes of an earlier call must never be able to stand in for bytes a kernel failed to write
        CK(cudaMemsetAsync(ctx->d_out.p, 0xA5, (out_hi - out_lo) + 64, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));              // dictionaries / poison are in place before any lane starts
    // SDZ_TRACE_PIPE=1: device timeline of every sub-batch on stderr
    static const bool trace_pipe = getenv("SDZ_TRACE_PIPE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace_pipe) {
        tev.resize(3 * K + 1);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[3 * K], ctx->s_h2d);
    }
    const auto t_host0 = std::chrono::steady_clock::now();
    // (everything that enqueues work: on an error nothing may stay in flight - the copies read the caller's buffers and
    //  write into them, and the next call may free the staging areas)
    auto pipeline = [&]() -> int {
        CK(cudaMemcpyAsync(ctx->d_meta.p, hm, meta_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
        // records and bytes of sub-batch c go home once its kernels are done
        auto enqueue_d2h = [&](uint64_t c) -> int {
            const uint64_t lo = cut[c], hi = cut[c + 1];
            if (lo == hi) return SDZ_OK;
            CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->pipe_ev[2 * c + 1], 0));
            CK(cudaMemcpyAsync(rdst + lo, (sdz_result*)ctx->d_res.p + lo, (hi - lo) * sizeof(sdz_result), cudaMemcpyDeviceToHost, ctx->s_d2h));
            if (!sizes_only && dense) {
                const uint64_t o_lo = d_out_off[lo], o_hi = d_out_off[hi - 1] + d_out_cap[hi - 1];
                uint8_t* land = out_staged ? (uint8_t*)ctx->h_out : out_arena + out_lo;
                CK(cudaMemcpyAsync(land + o_lo, (uint8_t*)ctx->d_out.p + o_lo, o_hi - o_lo, cudaMemcpyDeviceToHost, ctx->s_d2h));
                if (out_staged) CK(cudaEventRecord(ctx->pipe_ev[2 * c], ctx->s_d2h));      // (the input-landed event of c has been waited on)
            }
            if (trace_pipe) cudaEventRecord(tev[3 * c + 2], ctx->s_d2h);
            return SDZ_OK;
        };
        for (uint64_t c = 0; c < K; c++) {
            const uint64_t lo = cut[c], hi = cut[c + 1];
            if (lo == hi) continue;
            if (!direct) {
                std::vector<std::pair<uint8_t*, std::pair<const uint8_t*, size_t>>> jobs;
                jobs.reserve(hi - lo);
                for (uint64_t i = lo; i < hi; i++)
                    if (in_len[i]) jobs.push_back({ hs + in_off[i], { in[i].data, in_len[i] } });
                parallel_copy(jobs);
            }
            const size_t c_lo = in_off[lo], c_hi = (hi < n ? (size_t)in_off[hi] : (direct ? data_end : in_total));
            if (c_hi > c_lo)
                CK(cudaMemcpyAsync(d_in + c_lo, (direct ? ubase : hs) + c_lo, c_hi - c_lo, cudaMemcpyHostToDevice, ctx->s_h2d));
            CK(cudaEventRecord(ctx->pipe_ev[2 * c], ctx->s_h2d));
            if (trace_pipe) cudaEventRecord(tev[3 * c], ctx->s_h2d);
            ctx->cur_lane = (int)(c % n_lanes);
            cudaStream_t lane_st = ctx->lane_stream[ctx->cur_lane];
            CK(cudaStreamWaitEvent(lane_st, ctx->pipe_ev[2 * c], 0));
            sdz_batch_dev bc = b;
            bc.d_in_off += lo; bc.d_dict_off += lo; bc.d_out_off += lo; bc.d_in_len += lo; bc.d_dict_len += lo;
            bc.d_out_cap += lo; bc.d_dict_adler += lo; bc.d_mode += lo; bc.d_results += lo;
            bc.n = hi - lo;
            uint64_t tok_total = 0;                          // token arena of this sub-batch (the device computes the same sum)
            if (!sizes_only && use_fast)
                for (uint64_t i = lo; i < hi; i++) tok_total += sdz::token_cap(in_len[i], d_out_cap[i]);
            rc = run_batch_device(ctx, &bc, sizes_only, c == 0, c == K - 1, tok_total, spec);
            ctx->cur_lane = 0;
            if (rc) return rc;
            CK(cudaEventRecord(ctx->pipe_ev[2 * c + 1], lane_st));
            if (trace_pipe) cudaEventRecord(tev[3 * c + 1], lane_st);
            if (!ctx->h2d_first) { if ((rc = enqueue_d2h(c))) return rc; }
        }
        if (ctx->h2d_first) {
            // SDZ_H2D_FIRST=1: no device -> host copy starts before the last input has landed.  On hosts whose memory system
            // sustains less with both directions active than with one (profiles/r02n_8gpu_summary.md: 8 GPUs on one socket,
            // ~100 GB/s mixed against 168 GB/s device -> host alone) the two phases are faster one after the other.
            CK(cudaEventRecord(ctx->ev[3], ctx->s_h2d));
            CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev[3], 0));
            for (uint64_t c = 0; c < K; c++)
                if ((rc = enqueue_d2h(c))) return rc;
        }
        if (out_staged) {
            for (uint64_t c = 0; c < K; c++) {
                const uint64_t lo = cut[c], hi = cut[c + 1];
                if (lo == hi) continue;
                CK(cudaEventSynchronize(ctx->pipe_ev[2 * c]));
                const uint64_t o_lo = d_out_off[lo], o_hi = d_out_off[hi - 1] + d_out_cap[hi - 1];
                std::vector<std::pair<uint8_t*, std::pair<const uint8_t*, size_t>>> jobs;
                const size_t piece = 1u << 20;
                for (uint64_t o = o_lo; o < o_hi; o += piece)
                    jobs.push_back({ out_arena + out_lo + o, { (const uint8_t*)ctx->h_out + o, (size_t)std::min<uint64_t>(piece, o_hi - o) } });
                parallel_copy(jobs);
            }
        }
        return SDZ_OK;
    };
    if ((rc = pipeline())) {
        cudaStreamSynchronize(ctx->s_h2d);
        for (int l = 0; l < sdz_ctx::N_LANES; l++) { cudaStreamSynchronize(ctx->lane_stream[l]); cudaStreamSynchronize(ctx->fast_sb[l]); }
        cudaStreamSynchronize(ctx->s_d2h);
        cudaGetLastError();
        return rc;
    }
    const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    CK(cudaStreamSynchronize(ctx->s_d2h));
    for (int l = 0; l < n_lanes; l++) CK(cudaStreamSynchronize(ctx->lane_stream[l]));
    if (trace_pipe) {
        fprintf(stderr, "[sdz pipe] %llu sub-batches, host enqueue %.2f ms\n", (unsigned long long)K, host_ms);
        for (uint64_t c = 0; c < K; c++) {
            float a = 0, b2 = 0, d = 0;
            cudaEventElapsedTime(&a, tev[3 * K], tev[3 * c]);
            cudaEventElapsedTime(&b2, tev[3 * K], tev[3 * c + 1]);
            cudaEventElapsedTime(&d, tev[3 * K], tev[3 * c + 2]);
            fprintf(stderr, "[sdz pipe] %2llu: input landed %7.2f  kernels done %7.2f  output landed %7.2f ms\n", (unsigned long long)c, a, b2, d);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    if (!sizes_only && !dense) {
        for (uint64_t i = 0; i < n; i++)
            if (rdst[i].out_len)
                CK(cudaMemcpyAsync(out_arena + out_off[i], (uint8_t*)ctx->d_out.p + d_out_off[i], rdst[i].out_len,
                                   cudaMemcpyDeviceToHost, ctx->s_d2h));
        CK(cudaStreamSynchronize(ctx->s_d2h));
    }
    int ret = SDZ_OK;
    for (uint64_t i = 0; i < n; i++) {
        if (!sizes_only) {
            rdst[i].out_off = out_off[i];
            if (rdst[i].zstatus == SDZ_Z_BUF_ERROR) ret = SDZ_E_OUT_CAP;    // the kernel ran out of slot
        }
        if (out_len) out_len[i] = rdst[i].out_len;
    }
    if (results) memcpy(results, rdst, n * sizeof(sdz_result));
    return ret;
}

int sdz_inflate_batch(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off,
                      const uint64_t* out_cap, sdz_result* results, uint32_t flags)
{
    return inflate_host(ctx, in, n, out_arena, out_off, out_cap, results, nullptr, flags, false);
}

int sdz_inflate_sizes(sdz_ctx* ctx, const sdz_in* in, uint64_t n, uint64_t* out_len, uint32_t flags)
{
    if (!out_len && n) return SDZ_E_ARG;
    return inflate_host(ctx, in, n, nullptr, nullptr, nullptr, nullptr, out_len, flags, true);
}

}  // extern "C"

// ---------------------------------------------------------------------------- streaming sessions (class Inflater)
//
// One sdz_inflater = one `new Inflater(options)` fed with several append() calls (src/sd-inflate.ts:54-180).  The
// compressed bytes received so far and the bytes decoded so far both stay in HBM; between calls the session keeps an
// sdz_resume record - where the reference stopped - so that an append() costs the decode of the NEW input only (plus
// rebuilding the tables of the block it stopped in), not a re-decode of everything (round 1).

struct sdz_inflater {
    sdz_ctx* ctx = nullptr;
    bool raw = false, has_dict = false;
    DevBuf d_in, d_out, d_meta, d_dict;
    uint64_t in_len = 0;               // compressed bytes received
    uint64_t emitted = 0;              // decoded bytes handed to the caller so far
    uint64_t last_new = 0;             // bytes produced by the most recent append()
    uint32_t dict_len = 0;
    int32_t dict_adler = 1;
    int cur = 0;                       // which of the two resume slots is current
    int kind = SDZ_RESUME_START;
    int32_t running = 0;               // Inflater.checksum: chained over the chunks append() emitted
    bool have_running = false;
    sdz_result rec;                    // record of the most recent append()
    uint8_t* h_meta = nullptr;         // pinned mirror of the meta block
};

namespace {
// meta block layout (device and pinned host copy)
constexpr size_t SM_IN_OFF = 0, SM_DICT_OFF = 8, SM_OUT_OFF = 16, SM_IN_LEN = 24, SM_DICT_LEN = 28, SM_OUT_CAP = 32,
                 SM_DICT_ADLER = 36, SM_MODE = 40, SM_RES = 64, SM_RESUME = SM_RES + ((sizeof(sdz_result) + 15) / 16) * 16,
                 SM_BYTES = SM_RESUME + 2 * sizeof(sdz_resume);

int grow_keep(sdz_ctx* ctx, DevBuf& b, size_t bytes, size_t keep)
{
    if (bytes <= b.cap) return SDZ_OK;
    size_t want = std::max(bytes, b.cap * 2);
    want = (want + (1u << 20)) & ~((size_t(1) << 20) - 1);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return SDZ_E_NOMEM; }
    if (b.p && keep) {
        e = cudaMemcpyAsync(p, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("cudaMemcpy: ") + cudaGetErrorString(e); cudaFree(p); return SDZ_E_CUDA; }
    }
    if (b.p) cudaFree(b.p);
    b.p = p; b.cap = want;
    return SDZ_OK;
}
}  // namespace

extern "C" {

int sdz_inflater_create(sdz_ctx* ctx, int raw, const uint8_t* dict, uint32_t dict_len, sdz_inflater** out)
{
    if (!ctx || !out || (dict_len && !dict)) return SDZ_E_ARG;
    *out = nullptr;
    if (raw && dict) return SDZ_E_ARG;                       // RangeError in the reference (src/sd-inflate.ts:69-71)
    ENTER(ctx);
    sdz_inflater* s = new sdz_inflater();
    s->ctx = ctx; s->raw = raw != 0; s->has_dict = dict != nullptr; s->dict_len = dict_len;
    memset(&s->rec, 0, sizeof s->rec);
    auto fail = [&](int rc) { sdz_inflater_destroy(s); return rc; };
    int rc;
    if ((rc = grow(ctx, s->d_meta, SM_BYTES))) return fail(rc);
    s->h_meta = (uint8_t*)pinned_alloc(SM_BYTES, ctx->numa_node);
    if (!s->h_meta) return fail(SDZ_E_NOMEM);
    memset(s->h_meta, 0, SM_BYTES);
    if (dict_len) {
        if ((rc = grow(ctx, s->d_dict, dict_len + 16))) return fail(rc);
        if (cudaMemcpyAsync(s->d_dict.p, dict, dict_len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return fail(SDZ_E_CUDA);
        uint64_t dl = dict_len;
        if ((rc = checksum_chain(ctx, false, (const uint8_t*)s->d_dict.p, &dl, 1, 1, 1, nullptr, &s->dict_adler))) return fail(rc);
    } else if ((rc = grow(ctx, s->d_dict, 16))) return fail(rc);
    if (cudaMemsetAsync(s->d_meta.p, 0, SM_BYTES, ctx->stream) != cudaSuccess) return fail(SDZ_E_CUDA);
    *out = s;
    return SDZ_OK;
}

void sdz_inflater_destroy(sdz_inflater* s)
{
    if (!s) return;
    if (s->ctx) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
    }
    for (DevBuf* b : { &s->d_in, &s->d_out, &s->d_meta, &s->d_dict })
        if (b->p) cudaFree(b->p);
    pinned_free(s->h_meta);
    delete s;
}

int sdz_inflater_append(sdz_inflater* s, const uint8_t* data, uint64_t len, uint64_t* new_bytes, sdz_result* res)
{
    if (!s || !new_bytes || !res || (len && !data)) return SDZ_E_ARG;
    sdz_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    *new_bytes = 0;
    s->last_new = 0;
    if (len == 0) { *res = s->rec; res->thrown_append = 0; return SDZ_OK; }          // append() of an empty chunk returns []
    if (s->kind == SDZ_RESUME_FAILED) { *res = s->rec; return SDZ_OK; }              // the engine stays in its error state
    if (s->kind == SDZ_RESUME_DONE || s->kind == SDZ_RESUME_BROKEN_Q3) {
        if (s->kind == SDZ_RESUME_DONE) {
            // DONE returns STREAM_END without consuming a byte of the new chunk: "inflate error: bad input data"
            // (src/sd-inflate.ts:130-132).  (Bytes after the end inside the SAME chunk make append() spin instead)
            s->rec.thrown_append = SDZ_THROW_BAD_INPUT_DATA;
        } else {
            // proc() is entered in BTREE / DTREE with its locals gone: `default:` -> STREAM_ERROR (src/infblocks.ts:616-625)
            s->rec.thrown_append = SDZ_THROW_INFLATE_ERROR; s->rec.zstatus = SDZ_Z_STREAM_ERROR; s->rec.msg_id = SDZ_MSG_NONE;
        }
        s->rec.out_len = 0;
        s->kind = SDZ_RESUME_FAILED;
        *res = s->rec;
        return SDZ_OK;
    }
    if (s->in_len + len >= (1ull << 32) - 64 - SDZ_IN_PAD) return SDZ_E_ARG;
    int rc;
    if ((rc = grow_keep(ctx, s->d_in, s->in_len + len + SDZ_IN_PAD + 16, s->in_len))) return rc;
    CK(cudaMemcpyAsync((uint8_t*)s->d_in.p + s->in_len, data, len, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync((uint8_t*)s->d_in.p + s->in_len + len, 0, SDZ_IN_PAD, ctx->stream));
    s->in_len += len;

    uint8_t* hm = s->h_meta;
    uint8_t* dm = (uint8_t*)s->d_meta.p;
    const uint64_t have = s->emitted;                       // == resume.pos
    uint64_t cap = have + std::max<uint64_t>(1u << 20, 16 * len);
    for (;;) {
        if (cap >= (1ull << 32) - 64) cap = (1ull << 32) - 65;
        if ((rc = grow_keep(ctx, s->d_out, cap + 64, have))) return rc;
        *(uint64_t*)(hm + SM_IN_OFF) = 0; *(uint64_t*)(hm + SM_DICT_OFF) = 0; *(uint64_t*)(hm + SM_OUT_OFF) = 0;
        *(uint32_t*)(hm + SM_IN_LEN) = (uint32_t)s->in_len; *(uint32_t*)(hm + SM_DICT_LEN) = s->dict_len;
        *(uint32_t*)(hm + SM_OUT_CAP) = (uint32_t)cap; *(int32_t*)(hm + SM_DICT_ADLER) = s->dict_adler;
        hm[SM_MODE] = (uint8_t)((s->raw ? SDZ_MODE_RAW : SDZ_MODE_INFLATER) | (s->has_dict ? 0x80 : 0));
        CK(cudaMemcpyAsync(dm, hm, SM_RES, cudaMemcpyHostToDevice, ctx->stream));     // (records and resume slots stay as they are)
        sdz::InflateParams P;
        memset(&P, 0, sizeof P);
        P.in = (const uint8_t*)s->d_in.p; P.in_off = (const uint64_t*)(dm + SM_IN_OFF); P.in_len = (const uint32_t*)(dm + SM_IN_LEN);
        P.mode = dm + SM_MODE;
        P.dict = (const uint8_t*)s->d_dict.p; P.dict_off = (const uint64_t*)(dm + SM_DICT_OFF);
        P.dict_len = (const uint32_t*)(dm + SM_DICT_LEN); P.dict_adler = (const int32_t*)(dm + SM_DICT_ADLER);
        P.out = (uint8_t*)s->d_out.p; P.out_off = (const uint64_t*)(dm + SM_OUT_OFF); P.out_cap = (const uint32_t*)(dm + SM_OUT_CAP);
        P.res = (sdz_result*)(dm + SM_RES); P.n = 1; P.counter = ctx->d_counter;
        P.resume_in = (const sdz_resume*)(dm + SM_RESUME) + s->cur;
        P.resume_out = (sdz_resume*)(dm + SM_RESUME) + (s->cur ^ 1);
        ctx->cur_lane = 0;
        if ((rc = launch_inflate_t<4, true>(ctx, P))) return rc;
        CK(cudaMemcpyAsync(hm + SM_RES, dm + SM_RES, SM_BYTES - SM_RES, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const sdz_result* r = (const sdz_result*)(hm + SM_RES);
        if (r->zstatus != SDZ_Z_BUF_ERROR) break;
        if (cap >= (1ull << 32) - 65) return SDZ_E_OUT_CAP;
        cap = have + (cap - have) * 4;                      // the slot was too small: same input again, larger slot
    }
    sdz_result r = *(const sdz_result*)(hm + SM_RES);
    const sdz_resume z = ((const sdz_resume*)(hm + SM_RESUME))[s->cur ^ 1];
    s->cur ^= 1;
    s->kind = z.kind;
    // the record as finish() would see it now: checksum of the chunks emitted so far, chained over the <= 16 KiB chunks
    // THIS call emits (src/sd-inflate.ts:133-149; Q1 fires on a final chunk of 5552 or 11104 bytes)
    if (!r.thrown_append) {
        const uint64_t total = z.pos;
        const uint64_t fresh = total - have;
        if (fresh) {
            std::vector<uint64_t> segs;
            if (r.container == SDZ_GZIP) segs.push_back(fresh);                       // CRC-32 chains exactly over any chunking
            else for (uint64_t o = 0; o < fresh; o += 16384) segs.push_back(std::min<uint64_t>(16384, fresh - o));
            int32_t v = 0;
            rc = checksum_chain(ctx, r.container == SDZ_GZIP, (const uint8_t*)s->d_out.p + have, segs.data(), segs.size(),
                                s->have_running ? s->running : (r.container == SDZ_GZIP ? 0 : 1), 1, nullptr, &v);
            if (rc) return rc;
            s->running = v; s->have_running = true;
        }
        s->emitted = total;
        s->last_new = fresh;
        *new_bytes = fresh;
        r.out_len = total;
    } else r.out_len = 0;
    {
        const int32_t stored = r.stored_checksum, isize = r.stored_isize;
        const int cks = stored == 0 ? SDZ_UNCHECKED : ((s->have_running && stored == s->running) ? SDZ_MATCH : SDZ_MISMATCH);
        const int fsz = isize == 0 ? SDZ_UNCHECKED : (((int64_t)isize == (int64_t)s->emitted) ? SDZ_MATCH : SDZ_MISMATCH);
        r.running_checksum = s->have_running ? s->running : 0;
        r.have_running = s->have_running ? 1 : 0;
        r.checksum_state = (uint8_t)cks; r.size_state = (uint8_t)fsz;
        r.success = (uint8_t)(r.complete && cks != SDZ_MISMATCH && fsz != SDZ_MISMATCH);
    }
    s->rec = r;
    *res = r;
    return SDZ_OK;
}

int sdz_inflater_read(sdz_inflater* s, uint8_t* dst, uint64_t cap)
{
    if (!s || (s->last_new && !dst) || cap < s->last_new) return SDZ_E_ARG;
    sdz_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!s->last_new) return SDZ_OK;
    CK(cudaMemcpyAsync(dst, (const uint8_t*)s->d_out.p + (s->emitted - s->last_new), s->last_new, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

int sdz_inflater_input(sdz_inflater* s, uint64_t off, uint64_t len, uint8_t* dst)
{
    if (!s || (len && !dst) || off + len > s->in_len) return SDZ_E_ARG;
    sdz_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!len) return SDZ_OK;
    CK(cudaMemcpyAsync(dst, (const uint8_t*)s->d_in.p + off, len, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SDZ_OK;
}

int sdz_inflater_finish(sdz_inflater* s, sdz_result* res)
{
    if (!s || !res) return SDZ_E_ARG;
    *res = s->rec;
    res->out_len = s->emitted;
    return SDZ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------- Deflater wrappers
//
// The compressor itself stays on the CPU (out of scope); what the device contributes to `Deflater` is the checksum of
// the SOURCE data (src/sd-deflate.ts:185-190) - one launch for the whole batch - and with it the container around each
// raw deflate payload: zlib header 78 01 (78 20 + DICTID with a preset dictionary, src/sd-deflate.ts:98-116), gzip
// header with MTIME, XFL = 0, OS = 0xff and optional FNAME (:118-152), trailer Adler-32 big-endian / CRC-32 + ISIZE
// little-endian (:154-165).

extern "C" {

static uint64_t wrap_size(const sdz_wrap_in& w)
{
    if (w.format == SDZ_WRAP_RAW) return w.payload_len;
    if (w.format == SDZ_WRAP_DEFLATE) return 2 + (w.dict_adler ? 4 : 0) + w.payload_len + 4;
    const size_t nl = w.file_name && w.file_name[0] ? strlen(w.file_name) + 1 : 0;
    return 10 + nl + w.payload_len + 8;
}

int sdz_deflate_wrap_sizes(const sdz_wrap_in* in, uint64_t n, uint64_t* out_len)
{
    if ((n && !in) || (n && !out_len)) return SDZ_E_ARG;
    for (uint64_t i = 0; i < n; i++) {
        if (in[i].format > SDZ_WRAP_GZIP || (in[i].payload_len && !in[i].payload) || (in[i].source_len && !in[i].source)) return SDZ_E_ARG;
        out_len[i] = wrap_size(in[i]);
    }
    return SDZ_OK;
}

int sdz_deflate_wrap_batch(sdz_ctx* ctx, const sdz_wrap_in* in, uint64_t n, uint8_t* out_arena, const uint64_t* out_off, uint64_t* out_len)
{
    if (!ctx || (n && (!in || !out_arena || !out_off))) return SDZ_E_ARG;
    if (n == 0) return SDZ_OK;
    std::vector<const uint8_t*> bufs(n);
    std::vector<uint64_t> lens(n);
    std::vector<uint8_t> kind(n);
    std::vector<int32_t> sums(n);
    for (uint64_t i = 0; i < n; i++) {
        if (in[i].format > SDZ_WRAP_GZIP || (in[i].payload_len && !in[i].payload) || (in[i].source_len && !in[i].source)) return SDZ_E_ARG;
        bufs[i] = in[i].source; lens[i] = in[i].format == SDZ_WRAP_RAW ? 0 : in[i].source_len;
        kind[i] = in[i].format == SDZ_WRAP_GZIP ? 1 : 0;
    }
    // one append() per source: checksum = adler32(chunk, 1) / crc32(chunk, 0), the reference's own functions (Q1 included)
    int rc = sdz_checksum_batch(ctx, bufs.data(), lens.data(), kind.data(), nullptr, n, sums.data());
    if (rc) return rc;
    for (uint64_t i = 0; i < n; i++) {
        const sdz_wrap_in& w = in[i];
        uint8_t* o = out_arena + out_off[i];
        size_t p = 0;
        if (w.format == SDZ_WRAP_DEFLATE) {
            o[p++] = 0x78; o[p++] = w.dict_adler ? 0x20 : 0x01;
            if (w.dict_adler) { const uint32_t d = (uint32_t)w.dict_adler; o[p++] = (uint8_t)(d >> 24); o[p++] = (uint8_t)(d >> 16); o[p++] = (uint8_t)(d >> 8); o[p++] = (uint8_t)d; }
        } else if (w.format == SDZ_WRAP_GZIP) {
            const bool named = w.file_name && w.file_name[0];
            o[p++] = 0x1f; o[p++] = 0x8b; o[p++] = 8; o[p++] = named ? 0x08 : 0x00;
            for (int j = 0; j < 4; j++) o[p++] = (uint8_t)(w.mtime >> (8 * j));
            o[p++] = 0; o[p++] = 0xff;
            if (named) { const size_t nl = strlen(w.file_name) + 1; memcpy(o + p, w.file_name, nl); p += nl; }
        }
        if (w.payload_len) memcpy(o + p, w.payload, w.payload_len);
        p += w.payload_len;
        const uint32_t c = (uint32_t)sums[i];
        if (w.format == SDZ_WRAP_DEFLATE) { o[p++] = (uint8_t)(c >> 24); o[p++] = (uint8_t)(c >> 16); o[p++] = (uint8_t)(c >> 8); o[p++] = (uint8_t)c; }
        else if (w.format == SDZ_WRAP_GZIP) {
            for (int j = 0; j < 4; j++) o[p++] = (uint8_t)(c >> (8 * j));
            for (int j = 0; j < 4; j++) o[p++] = (uint8_t)((uint32_t)w.source_len >> (8 * j));
        }
        if (out_len) out_len[i] = p;
    }
    return SDZ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------- one large stream
//
// The work is organised as phases of a session (sdz_large_*), so that the SAME code serves one GPU
// (sdz_inflate_large drives all phases for part 0 of 1) and a stream spread over several GPUs, one
// process each (every rank drives the phases for its part and exchanges three small things with the
// others: the block index, the 32 KiB window at its left edge, and its partial CRC).

namespace {

struct LargeTasks {
    const uint64_t* bit = nullptr;      // block header position of every task
    const uint64_t* resume = nullptr;   // TM_MARK
    const uint64_t* out = nullptr;      // TM_MARK
    const uint32_t* limit = nullptr;    // TM_MARK
    uint16_t* sym = nullptr;            // TM_MARK
    sdz::Ckpt* ckpt = nullptr;          // TM_INDEX
    unsigned long long* ckpt_count = nullptr;
    unsigned long long ckpt_cap = 0;
    uint32_t ckpt_step = 0;
};

// block-task launch: n tasks over stream 0 (descriptors d_zero_off / d_len0)
template <int G, bool STORE, int TM>
int launch_tasks(sdz_ctx* ctx, const uint8_t* d_src, const LargeTasks& T, sdz_result* d_res, uint64_t n, const uint64_t* d_zero_off,
                 const uint32_t* d_len0)
{
    sdz::InflateParams P;
    memset(&P, 0, sizeof P);
    P.in = d_src; P.in_off = d_zero_off; P.in_len = d_len0;
    P.res = d_res; P.n = n; P.counter = ctx->d_counter;
    P.task_bit = T.bit; P.task_resume = T.resume; P.task_out = T.out; P.task_limit = T.limit; P.out16 = T.sym;
    P.ckpt = T.ckpt; P.ckpt_count = T.ckpt_count; P.ckpt_cap = T.ckpt_cap; P.ckpt_step = T.ckpt_step;
    return launch_inflate_t<G, STORE, TM>(ctx, P);
}

// output bytes per piece (power of two): 16 KiB keeps one wave of pieces in flight for streams of a few hundred MB;
// larger streams have pieces to spare and save table rebuilds and window steps with 32 KiB (measured at 1 GiB: 8 KiB
// 80.6 ms, 16 KiB 74.3 ms, 32 KiB 70.5 ms).  SDZ_LARGE_STEP overrides.
uint32_t large_ckpt_step(uint64_t compressed_len)
{
    static const uint32_t forced = [] {
        const char* e = getenv("SDZ_LARGE_STEP");
        if (!e) return 0u;
        uint32_t v = (uint32_t)atoi(e), p2 = 1024;
        while (p2 < v && p2 < (1u << 24)) p2 <<= 1;
        return p2;
    }();
    if (forced) return forced;
    return compressed_len >= (160ull << 20) ? 32768u : 16384u;
}

}  // namespace

struct sdz_large {
    sdz_ctx* ctx = nullptr;
    const uint8_t* data = nullptr;      // as given by the caller (host or device)
    uint64_t len = 0;
    uint8_t mode = 0;
    int on_device = 0;
    const uint8_t* d_src = nullptr;     // the stream in HBM
    // container header (src/inflate.ts:142-401)
    bool raw = false, is_gzip = false;
    int method = 0;
    int32_t mtime = 0;
    uint32_t name_off = 0, name_len = 0;
    uint64_t first_bit = 0, total_bits = 0;
    const uint64_t* d_zero_off = nullptr;
    const uint32_t* d_len0 = nullptr;
    // index of the part this session was asked for (handed to the caller)
    std::vector<sdz_large_block> blocks;
    std::vector<sdz_large_ckpt> ckpts;
    // plan: the chain of real blocks cut into pieces
    std::vector<uint64_t> t_bit, t_resume, t_off;
    std::vector<uint32_t> t_limit;
    uint64_t total_out = 0, end_bit = 0, n_blocks = 0;
    bool planned = false;
    int32_t stored = 0, isize = 0;
    uint64_t total_in = 0;
    // the part being decoded
    uint64_t p_lo = 0, p_hi = 0, bps = 1;
    uint8_t* d_out = nullptr;           // device address of output byte off_lo()
    DevBuf d_sym, d_tasks, d_desc;
    uint64_t* d_to = nullptr;           // absolute output offsets of the part's pieces (+ end)
    uint64_t off_lo() const { return t_off[p_lo]; }
    uint64_t off_hi() const { return t_off[p_hi]; }
    // trace
    bool trace = false;
    std::chrono::steady_clock::time_point t_last;
    void lap(const char* what)
    {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[sdz_large] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    }
};

namespace {

// extent + resume points of the blocks whose headers start at `starts` (count-only walk of ONE block each)
int large_extents(sdz_large* L, const std::vector<uint64_t>& starts, std::vector<sdz_result>& recs, std::vector<sdz::Ckpt>& cks)
{
    sdz_ctx* ctx = L->ctx;
    const uint64_t n = starts.size();
    recs.clear(); cks.clear();
    if (!n) return SDZ_OK;
    int r2;
    unsigned long long* d_count = ctx->d_counter + 3;
    if ((r2 = grow(ctx, ctx->d_task, (n + 16) * sizeof(uint64_t)))) return r2;
    if ((r2 = grow(ctx, ctx->d_res, n * sizeof(sdz_result)))) return r2;
    CK(cudaMemcpyAsync(ctx->d_task.p, starts.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long cap = n == 1 ? 65536 : L->len / 64 + 65536, got = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((r2 = grow(ctx, ctx->d_part, cap * sizeof(sdz::Ckpt)))) return r2;
        CK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
        LargeTasks T;
        T.bit = (const uint64_t*)ctx->d_task.p;
        T.ckpt = (sdz::Ckpt*)ctx->d_part.p; T.ckpt_count = d_count; T.ckpt_cap = cap; T.ckpt_step = large_ckpt_step(L->len);
        r2 = launch_tasks<4, false, sdz::TM_INDEX>(ctx, L->d_src, T, (sdz_result*)ctx->d_res.p, n, L->d_zero_off, L->d_len0);
        if (r2) return r2;
        CK(cudaMemcpyAsync(&got, d_count, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (got <= cap) break;
        cap = got + 1024;                                                   // extremely compressible data: once more with room
    }
    recs.resize(n);
    cks.resize(got);
    CK(cudaMemcpyAsync(recs.data(), ctx->d_res.p, n * sizeof(sdz_result), cudaMemcpyDeviceToHost, ctx->stream));
    if (got) CK(cudaMemcpyAsync(cks.data(), ctx->d_part.p, got * sizeof(sdz::Ckpt), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // order by (task, pos): the kernel appends in completion order.  Counting sort on the task, then the handful of
    // resume points of each block by position.
    if (got) {
        std::vector<uint32_t> first(n + 1, 0);
        for (const sdz::Ckpt& c : cks) first[c.task + 1]++;
        for (uint64_t i = 0; i < n; i++) first[i + 1] += first[i];
        std::vector<sdz::Ckpt> sorted(got);
        std::vector<uint32_t> fill(first.begin(), first.end() - 1);
        for (const sdz::Ckpt& c : cks) sorted[fill[c.task]++] = c;
        for (uint64_t i = 0; i < n; i++)
            std::sort(sorted.begin() + first[i], sorted.begin() + first[i + 1], [](const sdz::Ckpt& a, const sdz::Ckpt& b) { return a.pos < b.pos; });
        cks.swap(sorted);
    }
    return SDZ_OK;
}

void large_records(const std::vector<uint64_t>& starts, const std::vector<sdz_result>& recs, const std::vector<sdz::Ckpt>& cks,
                   std::vector<sdz_large_block>& blocks, std::vector<sdz_large_ckpt>& ckpts)
{
    for (size_t i = 0; i < starts.size(); i++) {
        sdz_large_block b;
        memset(&b, 0, sizeof b);
        b.bit = starts[i]; b.end_bit = recs[i].total_in; b.out_len = recs[i].out_len;
        b.last = (uint8_t)(recs[i].n_blocks != 0); b.btype = recs[i].container; b.ok = (uint8_t)(recs[i].zstatus == sdz::R_EOB);
        blocks.push_back(b);
    }
    for (const sdz::Ckpt& c : cks) {
        sdz_large_ckpt k;
        k.block_bit = starts[c.task]; k.bit = c.bit; k.pos = c.pos; k.reserved = 0;
        ckpts.push_back(k);
    }
}

}  // namespace

extern "C" int sdz_large_open(sdz_ctx* ctx, const uint8_t* data, uint64_t len, uint8_t mode, int on_device, sdz_large** out)
{
    if (!ctx || !out || (len && !data) || mode > SDZ_MODE_RAW) return SDZ_E_ARG;
    if (len >= (1ull << 32) - 64) return SDZ_E_ARG;
    *out = nullptr;
    ENTER(ctx);
    // ---- container header on the host (a few bytes; src/inflate.ts:142-401).  Anything but a plain, complete
    // header goes to the sequential decoder, which knows every corner of the reference's state machine.
    uint8_t head[1024];
    const size_t hn = (size_t)std::min<uint64_t>(len, sizeof head);
    if (on_device) CK(cudaMemcpy(head, data, hn, cudaMemcpyDeviceToHost)); else if (hn) memcpy(head, data, hn);
    bool raw = mode == SDZ_MODE_RAW, is_gzip = false;
    int method = 0;
    int32_t mtime = 0;
    uint32_t name_off = 0, name_len = 0;
    size_t hp = 0;
    auto seq = [&]() { ctx->err = "stream needs the sequential decoder"; return SDZ_E_UNSUPPORTED; };
    if (mode == SDZ_MODE_SNIFF) {
        if (len < 2) return seq();
        const bool ident = (head[0] == 0x78 && (((head[0] << 8) + head[1]) % 31) == 0) || (head[0] == 0x1f && head[1] == 0x8b);
        raw = !ident;
    }
    if (!raw) {
        if (hn < 2) return seq();
        if (head[0] == 0x1f) {
            if (head[1] != 0x8b || hn < 10 || (head[2] & 0xf) != 8 || (head[2] >> 4) + 8 > 15) return seq();
            is_gzip = true; method = head[2];
            const uint8_t fl = head[3];
            for (int i = 0; i < 4; i++) mtime = (int32_t)(((uint32_t)mtime >> 8) | ((uint32_t)head[4 + i] << 24));
            hp = 10;
            if (fl & 4) return seq();                                       // FEXTRA
            if (fl & 8) { name_off = (uint32_t)hp; while (hp < hn && head[hp]) { hp++; name_len++; } if (hp >= hn) return seq(); hp++; }
            if (fl & 16) { while (hp < hn && head[hp]) hp++; if (hp >= hn) return seq(); hp++; }
            if (fl & 2) hp += 2;
            if (hp >= hn) return seq();
        } else {
            method = head[0];
            if ((method & 0xf) != 8 || (method >> 4) + 8 > 15 || ((method << 8) + head[1]) % 31 != 0 || (head[1] & 0x20)) return seq();
            hp = 2;
        }
    }
    if (len < hp + 1) return seq();

    sdz_large* L = new sdz_large();
    L->ctx = ctx; L->data = data; L->len = len; L->mode = mode; L->on_device = on_device;
    L->raw = raw; L->is_gzip = is_gzip; L->method = method; L->mtime = mtime; L->name_off = name_off; L->name_len = name_len;
    L->first_bit = (uint64_t)hp * 8; L->total_bits = len * 8;
    L->trace = getenv("SDZ_TRACE_LARGE") != nullptr;
    L->t_last = std::chrono::steady_clock::now();
    if (ctx->large_cache_full) {
        L->d_sym = ctx->large_cache[0]; L->d_tasks = ctx->large_cache[1]; L->d_desc = ctx->large_cache[2];
        for (auto& b : ctx->large_cache) b = DevBuf();
        ctx->large_cache_full = false;
    }
    auto fail = [&](int rc) { sdz_large_close(L); return rc; };
    // ---- the stream in HBM
    int rc;
    L->d_src = data;
    if (!on_device) {
        if ((rc = grow(ctx, ctx->d_in, len + SDZ_IN_PAD))) return fail(rc);
        if (cudaMemcpyAsync(ctx->d_in.p, data, len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync((uint8_t*)ctx->d_in.p + len, 0, SDZ_IN_PAD, ctx->stream) != cudaSuccess) { ctx->err = "staging the stream failed"; return fail(SDZ_E_CUDA); }
        L->d_src = (const uint8_t*)ctx->d_in.p;
    } else if (reinterpret_cast<uintptr_t>(data) & 15) return fail(SDZ_E_ARG);
    // stream-0 descriptors for the task launches: in_off[0] = 0, in_len[0] = len
    // (owned by the session: the context's scratch is reused by every other call, e.g. the checksum of a slice)
    if ((rc = grow(ctx, L->d_desc, 64))) return fail(rc);
    {
        uint64_t zero_off = 0;
        uint32_t len32 = (uint32_t)len;
        if (cudaMemcpyAsync(L->d_desc.p, &zero_off, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync((uint8_t*)L->d_desc.p + 8, &len32, 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "staging the stream failed"; return fail(SDZ_E_CUDA); }
    }
    L->d_zero_off = (const uint64_t*)L->d_desc.p;
    L->d_len0 = (const uint32_t*)((uint8_t*)L->d_desc.p + 8);
    L->lap("header + input staging");
    *out = L;
    return SDZ_OK;
}

extern "C" void sdz_large_close(sdz_large* L)
{
    if (!L) return;
    cudaSetDevice(L->ctx->device);
    cudaStreamSynchronize(L->ctx->stream);
    sdz_ctx* ctx = L->ctx;
    if (!ctx->large_cache_full) {
        ctx->large_cache[0] = L->d_sym; ctx->large_cache[1] = L->d_tasks; ctx->large_cache[2] = L->d_desc;
        ctx->large_cache_full = true;
    } else {
        if (L->d_sym.p) cudaFree(L->d_sym.p);
        if (L->d_tasks.p) cudaFree(L->d_tasks.p);
        if (L->d_desc.p) cudaFree(L->d_desc.p);
    }
    delete L;
}

// pass 1 for part `part` of `n_parts`: candidate headers in that slice of the payload bits (+ the first block of the
// stream in part 0), their extents and resume points
extern "C" int sdz_large_index(sdz_large* L, uint32_t part, uint32_t n_parts, const sdz_large_block** blocks, uint64_t* n_blocks,
                               const sdz_large_ckpt** ckpts, uint64_t* n_ckpts)
{
    if (!L || !n_parts || part >= n_parts || !blocks || !n_blocks || !ckpts || !n_ckpts) return SDZ_E_ARG;
    sdz_ctx* ctx = L->ctx;
    ENTER(ctx);
    L->blocks.clear(); L->ckpts.clear();
    const uint64_t span_all = L->total_bits - L->first_bit;
    uint64_t lo = L->first_bit + span_all / n_parts * part, hi = part + 1 == n_parts ? L->total_bits : L->first_bit + span_all / n_parts * (part + 1);
    const uint64_t span = hi - lo;
    const uint64_t max_cand = span / 4096 + 4096, max_surv = span / 32 + 4096;
    int rc;
    if ((rc = grow(ctx, ctx->d_task, (max_cand + 16) * sizeof(uint64_t)))) return rc;
    if ((rc = grow(ctx, ctx->d_part, max_surv * sizeof(uint64_t)))) return rc;
    uint64_t* d_cand = (uint64_t*)ctx->d_task.p;
    unsigned long long* d_ncand = ctx->d_counter + 2;                        // [2] candidates, [3] survivors / resume points
    CK(cudaMemsetAsync(d_ncand, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (span) {
        const uint64_t tiles = (span + sdz::PF_TILE - 1) / sdz::PF_TILE;
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((tiles + 7) / 8, (uint64_t)ctx->sm_count * 8));
        sdz::prefilter_headers<<<grid, 256, 0, ctx->stream>>>(L->d_src, lo, hi, L->total_bits, (uint64_t*)ctx->d_part.p, d_ncand + 1, max_surv);
        L->lap("1a prefilter");
        sdz::verify_headers<<<ctx->sm_count * 16, 128, 0, ctx->stream>>>(L->d_src, L->total_bits, (const uint64_t*)ctx->d_part.p, d_ncand + 1, max_surv,
                                                                     d_cand, d_ncand, max_cand);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    unsigned long long counts[2] = { 0, 0 };
    CK(cudaMemcpyAsync(counts, d_ncand, sizeof counts, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (counts[0] > max_cand || counts[1] > max_surv) { ctx->err = "stream needs the sequential decoder"; return SDZ_E_UNSUPPORTED; }
    std::vector<uint64_t> cand(counts[0]);
    if (counts[0]) CK(cudaMemcpy(cand.data(), d_cand, counts[0] * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (part == 0) cand.push_back(L->first_bit);                            // the first block, whatever its type
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    L->lap("1a header search");
    std::vector<sdz_result> recs;
    std::vector<sdz::Ckpt> cks;
    if ((rc = large_extents(L, cand, recs, cks))) return rc;
    large_records(cand, recs, cks, L->blocks, L->ckpts);
    L->lap("1b block extents");
    if (L->trace) fprintf(stderr, "[sdz_large] part %u/%u: %llu survivors, %llu candidates, %llu resume points\n", part, n_parts, counts[1],
                          (unsigned long long)cand.size(), (unsigned long long)cks.size());
    *blocks = L->blocks.data(); *n_blocks = L->blocks.size();
    *ckpts = L->ckpts.data(); *n_ckpts = L->ckpts.size();
    return SDZ_OK;
}

// A stored block at bit `cur`?  Reads its header (3 bits, padding to the byte boundary, LEN, NLEN) from the stream and
// fills `blk`; false when the block is of another type (or the stream ends inside the header: the kernel's business).
static bool large_stored_header(sdz_large* L, uint64_t cur, sdz_large_block& blk)
{
    uint8_t h[6] = { 0 };
    const uint64_t b0 = cur >> 3;
    const uint64_t have = L->len > b0 ? std::min<uint64_t>(6, L->len - b0) : 0;
    if (have < 5) return false;
    if (L->on_device) { if (cudaMemcpy(h, L->data + b0, have, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; } }
    else memcpy(h, L->data + b0, have);
    const uint32_t t3 = (((uint32_t)h[0] | ((uint32_t)h[1] << 8)) >> (cur & 7)) & 7u;
    if ((t3 >> 1) != 0) return false;
    const uint64_t lb = (cur + 3 + 7) >> 3;                                // LEN starts on the next byte boundary
    if (lb + 4 > L->len || lb - b0 + 4 > have) return false;
    const uint8_t* q = h + (lb - b0);
    const uint32_t len = (uint32_t)q[0] | ((uint32_t)q[1] << 8), nlen = (uint32_t)q[2] | ((uint32_t)q[3] << 8);
    memset(&blk, 0, sizeof blk);
    blk.bit = cur;
    blk.btype = 0;
    blk.last = (uint8_t)(t3 & 1u);
    blk.ok = (uint8_t)(((~nlen) & 0xffffu) == len && lb + 4 + len <= L->len);      // else: damaged / truncated -> exact path
    blk.out_len = len;
    blk.end_bit = (lb + 4 + len) * 8;
    return true;
}

// The chain of real blocks from the first one (false candidates are never reached), cut into pieces at the resume
// points.  `blocks` / `ckpts`: the index of ALL parts, in any order.
extern "C" int sdz_large_plan(sdz_large* L, const sdz_large_block* blocks, uint64_t n_blocks, const sdz_large_ckpt* ckpts, uint64_t n_ckpts,
                              uint64_t* total_out, uint64_t* n_pieces)
{
    if (!L || (n_blocks && !blocks) || (n_ckpts && !ckpts)) return SDZ_E_ARG;
    sdz_ctx* ctx = L->ctx;
    ENTER(ctx);
    auto seq = [&]() { ctx->err = "stream needs the sequential decoder"; return SDZ_E_UNSUPPORTED; };
    std::vector<sdz_large_block> B(blocks, blocks + n_blocks);
    std::vector<sdz_large_ckpt> C(ckpts, ckpts + n_ckpts);
    // (the parts concatenated in rank order are already sorted)
    auto b_less = [](const sdz_large_block& a, const sdz_large_block& b) { return a.bit < b.bit; };
    auto c_less = [](const sdz_large_ckpt& a, const sdz_large_ckpt& b) { return a.block_bit != b.block_bit ? a.block_bit < b.block_bit : a.pos < b.pos; };
    if (!std::is_sorted(B.begin(), B.end(), b_less)) std::sort(B.begin(), B.end(), b_less);
    if (!std::is_sorted(C.begin(), C.end(), c_less)) std::sort(C.begin(), C.end(), c_less);
    L->t_bit.clear(); L->t_resume.clear(); L->t_off.clear(); L->t_limit.clear();
    uint64_t cur = L->first_bit, total = 0, nb = 0, n_single = 0, n_stored = 0;
    bool finished = false;
    int rc;
    // The reference's window / output-buffer bookkeeping, replayed over the chain (byte counts only, as in the kernels):
    // it decides whether a stored block is copied whole or loses its remaining length.  Blocks that end within
    // the last 16 bytes of the input end through the reference's input-frontier rules (step_general), which the plan
    // does not model: a non-empty stored block after one of those goes to the sequential decoder.
    sdz::RingModel ring;
    ring.init(0);
    bool ring_exact = true;
    // blocks the header search does not index are measured one kernel launch at a time: a long chain of them
    // (Z_FIXED, tiny flushes) is the sequential decoder's
    constexpr uint64_t MAX_SINGLE = 4096;
    for (uint64_t guard = 0; guard < (1ull << 26); guard++) {
        auto it = std::lower_bound(B.begin(), B.end(), cur, [](const sdz_large_block& a, uint64_t v) { return a.bit < v; });
        sdz_large_block blk;
        std::vector<sdz_large_ckpt> own;
        const sdz_large_ckpt* c = nullptr;
        size_t nc = 0;
        if (it != B.end() && it->bit == cur) {
            blk = *it;
            auto lo = std::lower_bound(C.begin(), C.end(), cur, [](const sdz_large_ckpt& a, uint64_t v) { return a.block_bit < v; });
            auto hi = std::upper_bound(lo, C.end(), cur, [](uint64_t v, const sdz_large_ckpt& a) { return v < a.block_bit; });
            c = C.data() + (lo - C.begin());
            nc = (size_t)(hi - lo);
        } else if (large_stored_header(L, cur, blk)) {
            // a stored block: its extent is in its header (src/infblocks.ts:238-271)
            n_stored++;
        } else {
            // a block the header search does not look for (fixed): measure it on its own
            if (n_single >= MAX_SINGLE) return seq();
            std::vector<uint64_t> one{ cur };
            std::vector<sdz_result> r1;
            std::vector<sdz::Ckpt> k1;
            std::vector<sdz_large_block> b1;
            n_single++;
            if ((rc = large_extents(L, one, r1, k1))) return rc;
            large_records(one, r1, k1, b1, own);
            blk = b1[0];
            c = own.data();
            nc = own.size();
        }
        if (!blk.ok) return seq();                                          // truncated or damaged: exact sequential path
        if (blk.out_len >= (1ull << 32)) return seq();
        if (blk.btype == 0) {
            // a stored block's copy depends on where the reference's 16 KiB output chunks fall
            uint32_t left = (uint32_t)blk.out_len;
            if (left && !ring_exact) return seq();
            while (left) {
                if (ring.room() == 0 && ring.make_room()) return seq();     // `left` is lost: what follows is not a block boundary
                const uint32_t t = std::min<uint32_t>(left, (uint32_t)ring.room());
                ring.q += (int)t;
                left -= t;
            }
        } else {
            ring.write((uint32_t)blk.out_len);                              // block_end() of the kernels
            if (ring.room() >= 258) ring.flush(); else ring.wash();
            if (blk.end_bit + 16 * 8 >= L->len * 8) ring_exact = false;
        }
        uint64_t from = 0, resume = 0;
        for (size_t k = 0; k <= nc; k++) {
            const uint64_t to = k < nc ? c[k].pos : blk.out_len;
            if (to > from) {
                L->t_bit.push_back(cur); L->t_resume.push_back(resume); L->t_off.push_back(total + from);
                L->t_limit.push_back((uint32_t)(to - from));
            }
            if (k < nc) { from = to; resume = c[k].bit; }
        }
        nb++;
        total += blk.out_len;
        cur = blk.end_bit;
        if (blk.last) { finished = true; break; }
    }
    if (!finished) return seq();
    L->t_off.push_back(total);
    L->total_out = total; L->end_bit = cur; L->n_blocks = nb;
    // ---- trailer (src/inflate.ts:423-463)
    const uint64_t tp = (L->end_bit + 7) >> 3;
    uint8_t tail[8] = { 0 };
    const int want = L->raw ? 0 : (L->is_gzip ? 8 : 4);
    const int have = (int)std::min<uint64_t>((uint64_t)want, L->len - std::min(L->len, tp));
    if (have) { if (L->on_device) CK(cudaMemcpy(tail, L->data + tp, have, cudaMemcpyDeviceToHost)); else memcpy(tail, L->data + tp, have); }
    if (have < want || tp + want < L->len) return seq();                    // truncated trailer / trailing bytes: exact path
    L->stored = 0; L->isize = 0;
    for (int i = 0; i < want; i++) {
        const uint32_t b = tail[i];
        if (L->is_gzip) { if (i < 4) L->stored = (int32_t)(((uint32_t)L->stored >> 8) | (b << 24)); else L->isize = (int32_t)(((uint32_t)L->isize >> 8) | (b << 24)); }
        else L->stored = (int32_t)(((uint32_t)L->stored << 8) | b);
    }
    L->total_in = tp + want;
    L->planned = true;
    L->lap("chain walk");
    if (L->trace) fprintf(stderr, "[sdz_large] %llu candidates, %llu blocks (%llu measured singly), %llu pieces, %llu -> %llu bytes\n",
                          (unsigned long long)n_blocks, (unsigned long long)nb, (unsigned long long)(n_single + n_stored), (unsigned long long)L->t_bit.size(),
                          (unsigned long long)L->len, (unsigned long long)total);
    if (total_out) *total_out = total;
    if (n_pieces) *n_pieces = L->t_bit.size();
    return SDZ_OK;
}

// output bytes [*off_lo, *off_hi) belong to part `part` of `n_parts` (cut at piece boundaries, balanced on bytes)
extern "C" int sdz_large_range(sdz_large* L, uint32_t part, uint32_t n_parts, uint64_t* off_lo, uint64_t* off_hi)
{
    if (!L || !L->planned || !n_parts || part >= n_parts || !off_lo || !off_hi) return SDZ_E_ARG;
    auto cut = [&](uint32_t k) -> uint64_t {
        if (k == 0) return 0;
        if (k >= n_parts) return L->t_bit.size();
        const uint64_t target = L->total_out / n_parts * k;
        return (uint64_t)(std::lower_bound(L->t_off.begin(), L->t_off.end() - 1, target) - L->t_off.begin());
    };
    *off_lo = L->t_off[cut(part)];
    *off_hi = L->t_off[cut(part + 1)];
    return SDZ_OK;
}

// pass 2a + 2b for this part: its pieces into 16-bit symbols, then the windows inside the part are composed.
// d_out: device address of output byte off_lo; 32 KiB BELOW it must be addressable when part > 0 (the window that
// arrives from the part before) and 64 bytes above off_hi.
extern "C" int sdz_large_decode(sdz_large* L, uint32_t part, uint32_t n_parts, uint8_t* d_out)
{
    if (!L || !L->planned || !n_parts || part >= n_parts) return SDZ_E_ARG;
    sdz_ctx* ctx = L->ctx;
    ENTER(ctx);
    uint64_t lo, hi;
    sdz_large_range(L, part, n_parts, &lo, &hi);
    L->p_lo = (uint64_t)(std::lower_bound(L->t_off.begin(), L->t_off.end() - 1, lo) - L->t_off.begin());
    L->p_hi = hi == L->total_out ? L->t_bit.size() : (uint64_t)(std::lower_bound(L->t_off.begin(), L->t_off.end() - 1, hi) - L->t_off.begin());
    L->d_out = d_out;
    const uint64_t nt = L->p_hi - L->p_lo;
    if (!nt) { CK(cudaEventRecord(ctx->ev[1], ctx->stream)); return SDZ_OK; }
    if (!d_out) return SDZ_E_ARG;
    int rc;
    if ((rc = grow(ctx, L->d_sym, (hi - lo + 64) * 2))) return rc;
    if ((rc = grow(ctx, L->d_tasks, (nt * 3 + 1) * 8 + nt * 4 + 64))) return rc;
    if ((rc = grow(ctx, ctx->d_res, nt * sizeof(sdz_result)))) return rc;
    uint64_t* d_tb = (uint64_t*)L->d_tasks.p;
    uint64_t* d_tr = d_tb + nt;
    L->d_to = d_tr + nt;
    uint32_t* d_tl = (uint32_t*)(L->d_to + nt + 1);
    CK(cudaMemcpyAsync(d_tb, L->t_bit.data() + L->p_lo, nt * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tr, L->t_resume.data() + L->p_lo, nt * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(L->d_to, L->t_off.data() + L->p_lo, (nt + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tl, L->t_limit.data() + L->p_lo, nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    // the kernels address symbols and bytes by ABSOLUTE output offset: hand them the address offset 0 would have
    uint16_t* sym0 = (uint16_t*)L->d_sym.p - lo;
    LargeTasks T;
    T.bit = d_tb; T.resume = d_tr; T.out = L->d_to; T.limit = d_tl; T.sym = sym0;
    rc = launch_tasks<4, true, sdz::TM_MARK>(ctx, L->d_src, T, (sdz_result*)ctx->d_res.p, nt, L->d_zero_off, L->d_len0);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    L->lap("2a marker decode");
    // ~sqrt(pieces of the WHOLE stream) per segment: pass 2b walks the pieces of a segment in order, pass 2c walks
    // the segments of ALL parts in order (rank after rank), so this balances the two serial chains
    // ... and never more segments than SMs: pass 2b runs one 1,024-thread CTA per segment, one CTA per SM
    L->bps = 1;
    while (L->bps * L->bps < L->t_bit.size()) L->bps++;
    L->bps = std::max<uint64_t>(L->bps, (nt + ctx->sm_count - 1) / ctx->sm_count);
    const unsigned nseg = (unsigned)((nt + L->bps - 1) / L->bps);
    sdz::propagate_in_segment<<<nseg, 1024, 0, ctx->stream>>>(sym0, L->d_to, nt, L->bps);
    ctx->launches++;
    CK(cudaGetLastError());
    // the marker pass must have reproduced the extents
    std::vector<sdz_result> chk(nt);
    CK(cudaMemcpyAsync(chk.data(), ctx->d_res.p, nt * sizeof(sdz_result), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t t = 0; t < nt; t++) {
        const uint64_t g = L->p_lo + t;
        bool bad = chk[t].zstatus != sdz::R_EOB || chk[t].out_len != L->t_limit[g];
        // a piece that is followed by another piece of the same block must end exactly on that piece's resume point
        if (g + 1 < L->t_bit.size() && L->t_bit[g + 1] == L->t_bit[g] && chk[t].total_in != L->t_resume[g + 1]) bad = true;
        if (bad) {
            if (L->trace)
                fprintf(stderr, "[sdz_large] piece %llu (part %u/%u, block bit %llu, resume %llu, limit %u): status %d, out_len %llu, end bit %llu, next resume %llu\n",
                        (unsigned long long)g, part, n_parts, (unsigned long long)L->t_bit[g], (unsigned long long)L->t_resume[g], L->t_limit[g],
                        chk[t].zstatus, (unsigned long long)chk[t].out_len, (unsigned long long)chk[t].total_in,
                        (unsigned long long)(g + 1 < L->t_bit.size() ? L->t_resume[g + 1] : 0));
            ctx->err = "stream needs the sequential decoder";
            return SDZ_E_UNSUPPORTED;
        }
    }
    L->lap("2b windows inside segments");
    return SDZ_OK;
}

// pass 2c: with the final 32 KiB before off_lo in place (at d_out - 32768; not needed for the part that starts the
// stream), the last 32 KiB of every segment of this part become final - in particular the last 32 KiB of the part,
// which is the window the next part is waiting for.
extern "C" int sdz_large_windows(sdz_large* L)
{
    if (!L || !L->planned) return SDZ_E_ARG;
    sdz_ctx* ctx = L->ctx;
    ENTER(ctx);
    const uint64_t nt = L->p_hi - L->p_lo;
    if (!nt) return SDZ_OK;
    const uint64_t lo = L->off_lo();
    sdz::propagate_segments<<<1, 1024, 0, ctx->stream>>>((const uint16_t*)L->d_sym.p - lo, L->d_out - lo, L->d_to, nt, L->bps);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    L->lap("2c windows across segments");
    return SDZ_OK;
}

// pass 2d: every remaining symbol of the part
extern "C" int sdz_large_resolve(sdz_large* L)
{
    if (!L || !L->planned) return SDZ_E_ARG;
    sdz_ctx* ctx = L->ctx;
    ENTER(ctx);
    const uint64_t nt = L->p_hi - L->p_lo;
    if (nt) {
        const uint64_t lo = L->off_lo();
        const unsigned gy = (unsigned)std::min<uint64_t>(nt, 65535);
        sdz::resolve_markers<<<dim3(4, gy), 256, 0, ctx->stream>>>((const uint16_t*)L->d_sym.p - lo, L->d_out - lo, L->d_to, nt, L->bps);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    L->lap("2d markers");
    return SDZ_OK;
}

// finish() record of the whole stream (src/sd-inflate.ts:159-179) and inflate()'s throw mapping (:214-225), given the
// running checksum of ALL output bytes (crc32 for gzip, the chunked adler32 otherwise)
extern "C" int sdz_large_finish(sdz_large* L, int32_t running, sdz_result* res)
{
    if (!L || !L->planned || !res) return SDZ_E_ARG;
    memset(res, 0, sizeof *res);
    const bool have_running = L->total_out > 0;
    res->out_off = 0;
    res->out_len = L->total_out;
    res->total_in = L->total_in;
    res->zstatus = SDZ_Z_STREAM_END;
    res->stored_checksum = L->stored;
    res->running_checksum = have_running ? running : 0;
    res->have_running = have_running;
    res->stored_isize = L->isize;
    res->mtime = L->mtime;
    res->name_off = L->name_len ? L->name_off : 0;
    res->name_len = L->name_len;
    res->n_blocks = (uint32_t)L->n_blocks;
    res->container = (uint8_t)(L->is_gzip ? SDZ_GZIP : (L->method == 0 ? SDZ_RAW : SDZ_ZLIB));
    res->complete = 1;
    const int cks = L->stored == 0 ? SDZ_UNCHECKED : ((have_running && L->stored == running) ? SDZ_MATCH : SDZ_MISMATCH);
    const int fsz = L->isize == 0 ? SDZ_UNCHECKED : (((int64_t)L->isize == (int64_t)L->total_out) ? SDZ_MATCH : SDZ_MISMATCH);
    res->checksum_state = (uint8_t)cks;
    res->size_state = (uint8_t)fsz;
    res->success = (uint8_t)(cks != SDZ_MISMATCH && fsz != SDZ_MISMATCH);
    res->thrown_inflate = (uint8_t)(res->success ? SDZ_THROW_NONE : (cks == SDZ_MISMATCH ? SDZ_THROW_INTEGRITY : SDZ_THROW_SIZE_CHECK));
    return SDZ_OK;
}

extern "C" int sdz_large_is_gzip(sdz_large* L) { return L && L->is_gzip; }

// crc32(A || B) from crc32(A), crc32(B) and len(B): multiplication by x^(8 len2) modulo the CRC polynomial
// (reflected), the same GF(2) arithmetic the device kernels use for their partials.  Pure host code.
// adler32(A || B) from adler32(A), adler32(B) (standard arithmetic, B seeded with 1) and len(B): the sums are linear in
// the bytes, so B's contribution is shifted by len(B) * (a of A) and the two initial 1s are taken out once
extern "C" int32_t sdz_adler32_combine(int32_t adler1, int32_t adler2, uint64_t len2)
{
    const uint32_t BASE = 65521u;
    const uint32_t a1 = (uint32_t)adler1, a2 = (uint32_t)adler2;
    const uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = a1 & 0xffffu;
    uint32_t sum2 = (uint32_t)(((uint64_t)rem * sum1) % BASE);
    sum1 += (a2 & 0xffffu) + BASE - 1u;
    sum2 += (a1 >> 16) + (a2 >> 16) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return (int32_t)(sum1 | (sum2 << 16));
}

extern "C" int32_t sdz_crc32_combine(int32_t crc1, int32_t crc2, uint64_t len2)
{
    // x^(8 * len2) by square-and-multiply over the bits of len2 (x2n[k] = x^(2^k), k >= 3 for bytes)
    uint32_t p = 1u << 31;                                                  // the polynomial "1"
    uint32_t sq = 1u << 30;                                                 // x^1
    for (int i = 0; i < 3; i++) sq = h_mulmod(sq, sq);                      // x^8
    for (uint64_t n = len2; n; n >>= 1) {
        if (n & 1) p = h_mulmod(p, sq);
        sq = h_mulmod(sq, sq);
    }
    return (int32_t)(h_mulmod(p, (uint32_t)crc1) ^ (uint32_t)crc2);
}

extern "C" int sdz_inflate_large(sdz_ctx* ctx, const uint8_t* data, uint64_t len, uint8_t mode, int on_device,
                                 uint8_t* out, uint64_t out_cap, sdz_result* res)
{
    if (!ctx || !res || (len && !data) || mode > SDZ_MODE_RAW) return SDZ_E_ARG;
    if (len >= (1ull << 32) - 64) return SDZ_E_ARG;
    ENTER(ctx);
    memset(res, 0, sizeof *res);
    sdz_large* L = nullptr;
    auto fallback = [&]() -> int {
        // exact but sequential: the ordinary decoder, one group for the whole stream
        if (L) { sdz_large_close(L); L = nullptr; }
        if (on_device) { ctx->err = "sdz_inflate_large: stream needs the sequential decoder; pass host pointers"; return SDZ_E_UNSUPPORTED; }
        sdz_in in1;
        memset(&in1, 0, sizeof in1);
        in1.data = data; in1.len = len; in1.mode = mode;
        uint64_t need = 0;
        int rc = inflate_host(ctx, &in1, 1, nullptr, nullptr, nullptr, nullptr, &need, 0, true);
        if (rc) return rc;
        if (need > out_cap) { res->out_len = need; return SDZ_E_OUT_CAP; }
        // (an empty / rejected / immediately failing stream needs no output: the record is still the reference's)
        uint8_t none[16];
        uint64_t off0 = 0, cap0 = need;
        return inflate_host(ctx, &in1, 1, need ? out : none, &off0, &cap0, res, nullptr, 0, false);
    };
    auto done = [&](int rc) { if (L) sdz_large_close(L); return rc; };
#define LARGE_STEP(call)                                          \
    do {                                                          \
        int rc_ = (call);                                         \
        if (rc_ == SDZ_E_UNSUPPORTED) return fallback();          \
        if (rc_) return done(rc_);                                \
    } while (0)
    LARGE_STEP(sdz_large_open(ctx, data, len, mode, on_device, &L));
    const sdz_large_block* blocks = nullptr;
    const sdz_large_ckpt* ckpts = nullptr;
    uint64_t nb = 0, nc = 0, total_out = 0, nt = 0;
    LARGE_STEP(sdz_large_index(L, 0, 1, &blocks, &nb, &ckpts, &nc));
    LARGE_STEP(sdz_large_plan(L, blocks, nb, ckpts, nc, &total_out, &nt));
    if (total_out > out_cap) { res->out_len = total_out; return done(SDZ_E_OUT_CAP); }
    uint8_t* d_o = out;
    if (!on_device) {
        int rc = grow(ctx, ctx->d_out, total_out + 64);
        if (rc) return done(rc);
        d_o = (uint8_t*)ctx->d_out.p;
    }
    LARGE_STEP(sdz_large_decode(L, 0, 1, d_o));
    LARGE_STEP(sdz_large_windows(L));
    LARGE_STEP(sdz_large_resolve(L));
#undef LARGE_STEP
    float ms_keep[3];
    cudaEventElapsedTime(&ms_keep[0], ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ms_keep[1], ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ms_keep[2], ctx->ev[0], ctx->ev[2]);

    // ---- running checksum as append() computes it over its 16 KiB chunks (src/sd-inflate.ts:133-149)
    int32_t running = 0;
    if (total_out > 0) {
        std::vector<uint64_t> segs;
        uint64_t rem = total_out;
        const uint64_t last = rem % 16384 ? rem % 16384 : 16384;
        rem -= last;
        while (rem >= (1ull << 30)) { segs.push_back(1ull << 30); rem -= 1ull << 30; }        // 2^30 is not a multiple of 5552
        if (rem) {
            if ((rem / 16384) % 347 == 0) { segs.push_back(rem - 16384); segs.push_back(16384); }  // never merge into a multiple of 5552 (Q1)
            else segs.push_back(rem);
        }
        segs.push_back(last);
        segs.erase(std::remove(segs.begin(), segs.end(), 0ull), segs.end());
        int rc = checksum_chain(ctx, L->is_gzip, d_o, segs.data(), segs.size(), L->is_gzip ? 0 : 1, 1, nullptr, &running);
        if (rc) return done(rc);
    }
    L->lap("checksum");
    ctx->last_ms[0] = ms_keep[0]; ctx->last_ms[1] = ms_keep[1]; ctx->last_ms[2] = ms_keep[2];
    if (!on_device && total_out) {
        if (cudaMemcpyAsync(out, d_o, total_out, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = cudaGetErrorString(cudaGetLastError()); return done(SDZ_E_CUDA); }
    }
    L->lap("output copy");
    int rc = sdz_large_finish(L, running, res);
    return done(rc);
}
