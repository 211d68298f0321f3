// api.cu — C ABI of libitrails_b200.so (see include/itrails_b200.h).
//
// Owns the device buffers (alignment blocks, padded model tables, Viterbi
// backpointers, posterior matrix) behind an opaque context and launches the kernels
// in hmm_kernels.cuh / model_kernels.cuh on one stream.  There is no CPU fallback:
// if CUDA is unavailable every entry point fails with ITR_ERR_CUDA.
#include "../../include/itrails_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "hmm_kernels.cuh"
#include "lockstep.h"
#include "model_build.h"
#include "csv_writer.h"
#include "model_plan.h"

using namespace itr;

static thread_local std::string g_create_error;

struct itr_ctx {
    int device = 0;
    // `stream` carries block uploads and model installs; every recursion has its own
    // stream (forward log-likelihood, Viterbi, posterior forward + combine, posterior
    // backward) so that independent recursions overlap on the device.
    cudaStream_t stream = nullptr, s_ll = nullptr, s_vit = nullptr, s_post = nullptr, stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ready = nullptr;
    cudaEvent_t ev_post_compute = nullptr;   // kernels of a posterior that is being downloaded are done
    bool post_download = false;
    bool async = false;
    double *pend_total = nullptr, *pend_per_block = nullptr;   // host outputs of a deferred itr_loglik
    bool pend_ll = false;
    int pend_sets = 0;                 // shape of the pending request (snapshot at call time)
    int64_t pend_blocks = 0;
    cudaDeviceProp prop{};
    std::string err;
    int64_t launches = 0, lockstep_launches = 0;

    // blocks
    int64_t n_blocks = 0, n_cols = 0, n_chunks = 0, max_T = 0;
    uint16_t *d_sym = nullptr;
    int64_t *d_off = nullptr;
    int32_t *d_order = nullptr;
    int64_t *d_chunk_off = nullptr;
    int32_t *d_chunk_blk = nullptr;
    bool comp_done = false;                   // the Viterbi sweep of the current call wrote d_comp itself
    unsigned int *d_queue = nullptr;
    unsigned long long *d_tile_ticket = nullptr;     // group tickets of posterior_tiles_mma_kernel
    float *d_LAfT = nullptr;                         // float(log a) transposed, 32 x 32 (viterbi_check32_kernel)
    size_t cap_sym = 0, cap_off = 0, cap_order = 0, cap_chunk_off = 0, cap_chunk_blk = 0;
    std::vector<int64_t> h_off;
    std::vector<int32_t> h_order;
    // ITR_POST_TRACE=1 (debug): per-block timeline of the posterior download, printed at the next sync
    std::vector<cudaEvent_t> tr_k, tr_c;
    std::vector<int32_t> tr_b;
    cudaEvent_t tr_0 = nullptr;
    std::vector<int64_t> h_tile_off;
    std::vector<cudaStream_t> grp_streams;     // 2 per posterior group
    std::vector<cudaEvent_t> grp_events;       // 2 per posterior group
    // itr_posterior_stream: pass 2 runs in `stream_ranges` contiguous block ranges; range r is
    // complete at range_events[r] and ends at column range_col_end[r]
    int stream_ranges = 0;
    bool ranges_valid = false;                 // range_events / range_col_end describe the posterior on the device
    std::vector<cudaEvent_t> range_events;
    std::vector<int64_t> range_col_end;

    // model
    int n_sets = 0, K = 0, KP = 0;
    double *d_A = nullptr, *d_PI = nullptr, *d_Et = nullptr;   // padded
    double *d_braw = nullptr;                                   // n_sets x K x 256 staging
    uint16_t *d_digits = nullptr;
    size_t cap_A = 0, cap_PI = 0, cap_Et = 0, cap_braw = 0;

    // run compression of the forward log-likelihood (hmm_kernels.cuh)
    int32_t *d_rep = nullptr, *d_sP = nullptr;
    unsigned long long *d_hist = nullptr;
    uint8_t *d_isrun = nullptr;
    long long *d_runinfo = nullptr;
    double *d_P = nullptr, *d_Pb = nullptr, *d_ebar = nullptr, *d_ck_a = nullptr, *d_ck_b = nullptr;
    size_t cap_P = 0, cap_Pb = 0, cap_sP = 0, cap_ebar = 0, cap_ck_a = 0, cap_ck_b = 0;
    int64_t *d_tile_off = nullptr, *d_part_tile = nullptr;   // d_part_tile[b]: id of block b's partial last tile, or -1
    int32_t *d_tile_blk = nullptr;
    unsigned long long *d_tile_info = nullptr;               // one word per tile (tile_table_kernel)
    size_t cap_tile_off = 0, cap_tile_blk = 0, cap_part_tile = 0, cap_tile_info = 0;
    int64_t n_part_tiles = 0;
    int64_t n_tiles = 0;
    bool runs_valid = false, use_runs = false;

    // log-likelihood
    double *d_ll = nullptr;
    size_t cap_ll = 0;
    double *h_ll = nullptr;      // pinned
    size_t cap_hll = 0;

    // Viterbi
    double *d_LA = nullptr, *d_LEt = nullptr, *d_OM0 = nullptr, *d_tmp = nullptr;
    size_t cap_LA = 0, cap_LEt = 0, cap_OM0 = 0, cap_tmp = 0;
    uint8_t *d_bp = nullptr, *d_comp = nullptr, *d_chunk_end = nullptr, *d_path = nullptr;
    size_t cap_bp = 0, cap_comp = 0, cap_chunk_end = 0, cap_path = 0;
    int32_t *d_final = nullptr;
    size_t cap_final = 0;
    bool have_path = false;

    // posterior
    double *d_post = nullptr, *d_beta = nullptr;
    unsigned int *d_ls_ll = nullptr, *d_ls_post = nullptr;     // scratch of the lock-step launches (lockstep.h)
    size_t cap_ls_ll = 0, cap_ls_post = 0;
    size_t cap_post = 0, cap_beta = 0;
    bool have_post = false;

    // timing
    cudaEvent_t ev0[ITR_PH_COUNT]{}, ev1[ITR_PH_COUNT]{};
    bool ev_valid[ITR_PH_COUNT]{};

    ModelBuilder *builder = nullptr;
};

// ---------------------------------------------------------------------------------
static int fail(itr_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

// The recursions, the posterior length groups and the copy stream are ~20 concurrent CUDA
// streams; with the default of 8 hardware work queues they alias and serialise behind
// each other (measured: log-likelihood 7 -> 17 ms inside a step).  The variable is read
// when the CUDA context is created, so it is set when the library is loaded; a value
// chosen by the user wins.
namespace {
struct ConnectionsEnv {
    ConnectionsEnv() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_env;
}  // namespace

#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? ITR_ERR_NOMEM : ITR_ERR_CUDA, \
                        "%s failed: %s", #call, cudaGetErrorString(e_));              \
    } while (0)

template <typename T>
static cudaError_t ensure(T *&p, size_t &cap, size_t n) {
    if (n <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

static void phase_begin(itr_ctx *c, int ph, cudaStream_t st = nullptr) { cudaEventRecord(c->ev0[ph], st ? st : c->stream); }
static void phase_end(itr_ctx *c, int ph, cudaStream_t st = nullptr) {
    cudaEventRecord(c->ev1[ph], st ? st : c->stream);
    c->ev_valid[ph] = true;
}

// pad a batch of row-major matrices [nb][rows][cols] into [nb][rows_p][cols_p]
__global__ void pad_kernel(const double *__restrict__ src, double *__restrict__ dst, int nb, int rows,
                           int cols, int rows_p, int cols_p, double pad) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)nb * rows_p * cols_p;
    if (idx >= total) return;
    const int c = idx % cols_p;
    const int r = (idx / cols_p) % rows_p;
    const size_t b = idx / ((size_t)cols_p * rows_p);
    dst[idx] = (r < rows && c < cols) ? src[(b * rows + r) * cols + c] : pad;
}

static inline unsigned blocks_for(size_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------
extern "C" int itr_version(void) { return 1000; }

extern "C" const char *itr_last_error(const itr_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int itr_create(int device, itr_ctx **out) {
    itr_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, ITR_ERR_ARG, "itr_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, ITR_ERR_CUDA, "itr_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, ITR_ERR_ARG, "itr_create: device %d out of range [0,%d)", device, n);
    ctx = new (std::nothrow) itr_ctx();
    if (!ctx) return fail(nullptr, ITR_ERR_NOMEM, "itr_create: out of host memory");
    ctx->device = device;
    auto bail = [&](cudaError_t err, const char *what) {
        fail(nullptr, ITR_ERR_CUDA, "itr_create: %s: %s", what, cudaGetErrorString(err));
        delete ctx;
        return (int)ITR_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (ctx->prop.major < 10) {
        fail(nullptr, ITR_ERR_UNSUPPORTED, "itr_create: device %d is sm_%d%d; this library is built for sm_100a only",
             device, ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return ITR_ERR_UNSUPPORTED;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (cudaStream_t *st : {&ctx->s_ll, &ctx->s_vit, &ctx->s_post, &ctx->stream2})
        if ((e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_post_compute, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (int i = 0; i < ITR_PH_COUNT; ++i) {
        if ((e = cudaEventCreate(&ctx->ev0[i])) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&ctx->ev1[i])) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    if ((e = cudaMalloc((void **)&ctx->d_queue, 64 * sizeof(unsigned int))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_tile_ticket, sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_LAfT, 32 * 32 * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
    // symbol digit table: read_data.py:6-24 ordering
    {
        std::vector<uint16_t> dig(NSYM);
        int next = 256;
        for (int a = 0; a < 5; ++a)
            for (int b = 0; b < 5; ++b)
                for (int c = 0; c < 5; ++c)
                    for (int d = 0; d < 5; ++d) {
                        const uint16_t packed = (uint16_t)(a | (b << 3) | (c << 6) | (d << 9));
                        if (a < 4 && b < 4 && c < 4 && d < 4) dig[64 * a + 16 * b + 4 * c + d] = packed;
                        else dig[next++] = packed;
                    }
        if ((e = cudaMalloc((void **)&ctx->d_digits, NSYM * sizeof(uint16_t))) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = cudaMemcpy(ctx->d_digits, dig.data(), NSYM * sizeof(uint16_t), cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy");
    }
    *out = ctx;
    return ITR_OK;
}

extern "C" void itr_destroy(itr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (cudaStream_t st : {ctx->stream, ctx->s_ll, ctx->s_vit, ctx->s_post, ctx->stream2})
        if (st) cudaStreamSynchronize(st);
    delete ctx->builder;
    if (ctx->h_ll) cudaFreeHost(ctx->h_ll);
    for (cudaStream_t st : ctx->grp_streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : ctx->grp_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->range_events) cudaEventDestroy(ev);
    void *ptrs[] = {ctx->d_sym, ctx->d_off, ctx->d_order, ctx->d_chunk_off, ctx->d_chunk_blk, ctx->d_queue,
                    ctx->d_A, ctx->d_PI, ctx->d_Et, ctx->d_braw, ctx->d_digits, ctx->d_ll, ctx->d_LA,
                    ctx->d_LEt, ctx->d_OM0, ctx->d_tmp, ctx->d_bp, ctx->d_comp, ctx->d_chunk_end,
                    ctx->d_path, ctx->d_final, ctx->d_post, ctx->d_beta, ctx->d_rep, ctx->d_sP, ctx->d_hist,
                    ctx->d_isrun, ctx->d_runinfo, ctx->d_P, ctx->d_ebar, ctx->d_Pb, ctx->d_ck_a, ctx->d_ck_b,
                    ctx->d_tile_off, ctx->d_tile_blk, ctx->d_part_tile, ctx->d_tile_info, ctx->d_ls_ll, ctx->d_ls_post,
                    ctx->d_tile_ticket, ctx->d_LAfT};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < ITR_PH_COUNT; ++i) {
        if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_post_compute) cudaEventDestroy(ctx->ev_post_compute);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    for (cudaStream_t st : {ctx->s_ll, ctx->s_vit, ctx->s_post, ctx->stream2})
        if (st) cudaStreamDestroy(st);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int itr_device_info(itr_ctx *ctx, char *name, int cap, int *sm_count, int *cc_major, int *cc_minor) {
    if (!ctx) return ITR_ERR_ARG;
    if (name && cap > 0) {
        strncpy(name, ctx->prop.name, cap - 1);
        name[cap - 1] = 0;
    }
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    return ITR_OK;
}

extern "C" double itr_phase_ms(itr_ctx *ctx, int phase) {
    if (!ctx || phase < 0 || phase >= ITR_PH_COUNT || !ctx->ev_valid[phase]) return -1.0;
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->ev1[phase]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0[phase], ctx->ev1[phase]) != cudaSuccess) return -1.0;
    return (double)ms;
}

extern "C" int64_t itr_launch_count(const itr_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t itr_lockstep_launch_count(const itr_ctx *ctx) { return ctx ? ctx->lockstep_launches : 0; }
extern "C" int64_t itr_total_columns(const itr_ctx *ctx) { return ctx ? ctx->n_cols : 0; }
extern "C" int64_t itr_num_blocks(const itr_ctx *ctx) { return ctx ? ctx->n_blocks : 0; }

// ---------------------------------------------------------------------------------
// blocks
// ---------------------------------------------------------------------------------
// Data or model are about to change: drain every recursion stream first.
static int quiesce(itr_ctx *ctx);
static int prepare_runs(itr_ctx *ctx, cudaStream_t st);

static void dump_trace(itr_ctx *ctx) {
    if (!ctx->tr_0) return;
    cudaStreamSynchronize(ctx->s_post);
    for (size_t i = 0; i < ctx->tr_b.size(); ++i) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->tr_0, ctx->tr_k[i]);
        cudaEventElapsedTime(&b, ctx->tr_0, ctx->tr_c[i]);
        fprintf(stderr, "post-trace block %d len %lld pass2_done %.3f copy_done %.3f\n", ctx->tr_b[i],
                (long long)(ctx->h_off[ctx->tr_b[i] + 1] - ctx->h_off[ctx->tr_b[i]]), a, b);
        cudaEventDestroy(ctx->tr_k[i]);
        cudaEventDestroy(ctx->tr_c[i]);
    }
    for (int ph : {(int)ITR_PH_LOGLIK, (int)ITR_PH_VITERBI_FWD, (int)ITR_PH_POST_FWD, (int)ITR_PH_POST_TOTAL}) {
        float a = -1, b = -1;
        if (ctx->ev_valid[ph] && cudaEventSynchronize(ctx->ev1[ph]) == cudaSuccess) {
            if (cudaEventElapsedTime(&a, ctx->tr_0, ctx->ev0[ph]) != cudaSuccess) a = -1;
            if (cudaEventElapsedTime(&b, ctx->tr_0, ctx->ev1[ph]) != cudaSuccess) b = -1;
        }
        cudaGetLastError();
        fprintf(stderr, "post-trace phase %d begin %.3f end %.3f\n", ph, a, b);
    }
    cudaEventDestroy(ctx->tr_0);
    ctx->tr_0 = nullptr;
    ctx->tr_k.clear();
    ctx->tr_c.clear();
    ctx->tr_b.clear();
}

static int install_blocks(itr_ctx *ctx, const uint16_t *sym, const int64_t *off, int64_t n_blocks) {
    CK(cudaSetDevice(ctx->device));
    int qrc = quiesce(ctx);
    if (qrc) return qrc;
    const int64_t n_cols = off[n_blocks];
    std::vector<int32_t> order(n_blocks);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return off[x + 1] - off[x] > off[y + 1] - off[y];
    });
    std::vector<int64_t> chunk_off(n_blocks + 1, 0);
    int64_t max_T = 0;
    for (int64_t b = 0; b < n_blocks; ++b) {
        const int64_t T = off[b + 1] - off[b];
        max_T = std::max(max_T, T);
        chunk_off[b + 1] = chunk_off[b] + (T + VCHUNK - 1) / VCHUNK;
    }
    std::vector<int64_t> tile_off(n_blocks + 1, 0);
    for (int64_t b = 0; b < n_blocks; ++b) tile_off[b + 1] = tile_off[b] + (off[b + 1] - off[b] + PTILE - 1) / PTILE;
    const int64_t n_tiles = tile_off[n_blocks];
    std::vector<int64_t> part_tile(n_blocks, -1);
    int64_t n_part = 0;
    for (int64_t b = 0; b < n_blocks; ++b)
        if ((off[b + 1] - off[b]) % PTILE) { part_tile[b] = tile_off[b + 1] - 1; ++n_part; }
    const int64_t n_chunks = chunk_off[n_blocks];
    std::vector<int32_t> chunk_blk(n_chunks);
    for (int64_t b = 0; b < n_blocks; ++b)
        std::fill(chunk_blk.begin() + chunk_off[b], chunk_blk.begin() + chunk_off[b + 1], (int32_t)b);

    ctx->n_blocks = ctx->n_cols = ctx->n_chunks = 0;
    ctx->have_path = ctx->have_post = ctx->ranges_valid = false;
    // device buffers are kept across loads when they are large enough
    // (+64 columns of slack so tile prefetches past the end stay in bounds)
    CK(ensure(ctx->d_sym, ctx->cap_sym, (size_t)(n_cols + 64)));
    CK(ensure(ctx->d_off, ctx->cap_off, (size_t)(n_blocks + 1)));
    CK(ensure(ctx->d_order, ctx->cap_order, (size_t)n_blocks));
    CK(ensure(ctx->d_chunk_off, ctx->cap_chunk_off, (size_t)(n_blocks + 1)));
    CK(ensure(ctx->d_chunk_blk, ctx->cap_chunk_blk, (size_t)std::max<int64_t>(n_chunks, 1)));
    CK(ensure(ctx->d_tile_off, ctx->cap_tile_off, (size_t)(n_blocks + 1)));
    CK(ensure(ctx->d_tile_blk, ctx->cap_tile_blk, (size_t)std::max<int64_t>(n_tiles, 1)));
    CK(ensure(ctx->d_part_tile, ctx->cap_part_tile, (size_t)n_blocks));
    CK(cudaMemcpyAsync(ctx->d_part_tile, part_tile.data(), (size_t)n_blocks * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tile_off, tile_off.data(), (size_t)(n_blocks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_sym + n_cols, 0, 64 * sizeof(uint16_t), ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sym, sym, (size_t)n_cols * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_off, off, (size_t)(n_blocks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    // per-tile tables are built on the device (7.8 M tiles at 250 Mb: nothing to upload)
    CK(ensure(ctx->d_tile_info, ctx->cap_tile_info, (size_t)std::max<int64_t>(n_tiles, 1)));
    tile_table_kernel<<<(unsigned)std::min<int64_t>(n_blocks, 4 * ctx->prop.multiProcessorCount), 256, 0, ctx->stream>>>(
        ctx->d_off, ctx->d_tile_off, (int)n_blocks, ctx->d_tile_blk, ctx->d_tile_info);
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->d_order, order.data(), (size_t)n_blocks * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_chunk_off, chunk_off.data(), (size_t)(n_blocks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_chunk_blk, chunk_blk.data(), (size_t)n_chunks * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    // symbol histogram (model independent): decides whether the run-compressed forward
    // sweep is worth using — are most columns covered by a handful of symbols?
    if (!ctx->d_hist) {
        CK(cudaMalloc((void **)&ctx->d_rep, NSYM * sizeof(int32_t)));
        CK(cudaMalloc((void **)&ctx->d_hist, (NSYM + 1) * sizeof(unsigned long long)));
        CK(cudaMalloc((void **)&ctx->d_isrun, 640));
        CK(cudaMalloc((void **)&ctx->d_runinfo, 2 * sizeof(long long)));
    }
    CK(cudaMemsetAsync(ctx->d_hist, 0, (NSYM + 1) * sizeof(unsigned long long), ctx->stream));
    symbol_hist_kernel<<<std::min<unsigned>(blocks_for((size_t)n_cols, 256), 4u * ctx->prop.multiProcessorCount), 256, 0, ctx->stream>>>(
        ctx->d_sym, n_cols, ctx->d_hist);
    ctx->launches += 1;
    std::vector<unsigned long long> hist(NSYM + 1);
    CK(cudaMemcpyAsync(hist.data(), ctx->d_hist, (NSYM + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (hist[NSYM]) {                    // validated on the device; name the first offender
        for (int64_t i = 0; i < n_cols; ++i)
            if (sym[i] >= NSYM) return fail(ctx, ITR_ERR_ARG, "symbol %u at column %lld is outside 0..624", sym[i], (long long)i);
    }
    hist.pop_back();
    CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
    std::partial_sort(hist.begin(), hist.begin() + 4, hist.end(), std::greater<unsigned long long>());
    ctx->use_runs = 2 * (hist[0] + hist[1] + hist[2] + hist[3]) > (unsigned long long)n_cols;
    ctx->runs_valid = false;
    ctx->n_tiles = n_tiles;
    ctx->n_part_tiles = n_part;
    ctx->h_tile_off = tile_off;
    ctx->h_off.assign(off, off + n_blocks + 1);
    ctx->h_order = order;
    ctx->n_blocks = n_blocks;
    ctx->n_cols = n_cols;
    if (ctx->K > 0 && ctx->K <= 32 && ctx->use_runs) {
        int prc = prepare_runs(ctx, ctx->stream);
        if (prc) return prc;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
    }
    ctx->n_blocks = n_blocks;
    ctx->n_cols = n_cols;
    ctx->n_chunks = n_chunks;
    ctx->max_T = max_T;
    return ITR_OK;
}

static int check_offsets(itr_ctx *ctx, const int64_t *off, int64_t n_blocks) {
    if (!off) return fail(ctx, ITR_ERR_ARG, "block_offsets is NULL");
    if (n_blocks <= 0) return fail(ctx, ITR_ERR_ARG, "n_blocks must be positive (got %lld)", (long long)n_blocks);
    if (n_blocks > 0x7fffffff / 2) return fail(ctx, ITR_ERR_ARG, "too many blocks (%lld)", (long long)n_blocks);
    if (off[0] != 0) return fail(ctx, ITR_ERR_ARG, "block_offsets[0] must be 0");
    for (int64_t b = 0; b < n_blocks; ++b)
        if (off[b + 1] <= off[b])
            return fail(ctx, ITR_ERR_ARG, "block %lld is empty or offsets are not increasing", (long long)b);
    return ITR_OK;
}

extern "C" int itr_load_blocks(itr_ctx *ctx, const uint16_t *sym, const int64_t *off, int64_t n_blocks) {
    if (!ctx) return ITR_ERR_ARG;
    if (!sym) return fail(ctx, ITR_ERR_ARG, "sym is NULL");
    int rc = check_offsets(ctx, off, n_blocks);
    if (rc) return rc;
    return install_blocks(ctx, sym, off, n_blocks);      // symbols are range-checked on the device
}

extern "C" int itr_load_blocks_i64(itr_ctx *ctx, const int64_t *sym, const int64_t *off, int64_t n_blocks) {
    if (!ctx) return ITR_ERR_ARG;
    if (!sym) return fail(ctx, ITR_ERR_ARG, "sym is NULL");
    int rc = check_offsets(ctx, off, n_blocks);
    if (rc) return rc;
    const int64_t n = off[n_blocks];
    std::vector<uint16_t> tmp;
    try { tmp.resize(n); } catch (...) { return fail(ctx, ITR_ERR_NOMEM, "out of host memory"); }
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = sym[i];
        if (v < 0 || v >= NSYM) return fail(ctx, ITR_ERR_ARG, "symbol %lld at column %lld is outside 0..624", (long long)v, (long long)i);
        tmp[i] = (uint16_t)v;
    }
    return install_blocks(ctx, tmp.data(), off, n_blocks);
}

// ---------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------
static int padded_K(int K) { return ((K + 31) / 32) * 32; }

static void finish_loglik(itr_ctx *ctx);
static int quiesce(itr_ctx *ctx) {
    for (cudaStream_t st : {ctx->s_ll, ctx->s_vit, ctx->s_post, ctx->stream2}) CK(cudaStreamSynchronize(st));
    finish_loglik(ctx);
    return ITR_OK;
}
// (the posterior group streams always join s_post before a call returns control)

// Installs a model whose raw arrays are already on the device.
int install_model_device(itr_ctx *ctx, int n_sets, int K, const double *d_a, const double *d_b, const double *d_pi) {
    const int KP = padded_K(K);
    int qrc = quiesce(ctx);
    if (qrc) return qrc;
    CK(ensure(ctx->d_A, ctx->cap_A, (size_t)n_sets * KP * KP));
    CK(ensure(ctx->d_PI, ctx->cap_PI, (size_t)n_sets * KP));
    CK(ensure(ctx->d_Et, ctx->cap_Et, (size_t)n_sets * NSYM * KP));
    phase_begin(ctx, ITR_PH_EMIT_TABLE);
    pad_kernel<<<blocks_for((size_t)n_sets * KP * KP, 256), 256, 0, ctx->stream>>>(d_a, ctx->d_A, n_sets, K, K, KP, KP, 0.0);
    pad_kernel<<<blocks_for((size_t)n_sets * KP, 256), 256, 0, ctx->stream>>>(d_pi, ctx->d_PI, n_sets, 1, K, 1, KP, 0.0);
    emission_table_kernel<<<blocks_for((size_t)n_sets * NSYM * KP, 256), 256, 0, ctx->stream>>>(d_b, ctx->d_digits, ctx->d_Et, K, KP, n_sets);
    phase_end(ctx, ITR_PH_EMIT_TABLE);
    ctx->launches += 3;
    CK(cudaGetLastError());
    ctx->n_sets = n_sets;
    ctx->K = K;
    ctx->KP = KP;
    ctx->have_path = ctx->have_post = ctx->ranges_valid = false;
    ctx->runs_valid = false;
    if (ctx->n_blocks > 0 && K <= 32 && ctx->use_runs) {
        int prc = prepare_runs(ctx, ctx->stream);
        if (prc) return prc;
    }
    CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
    return ITR_OK;
}

extern "C" int itr_set_model(itr_ctx *ctx, int n_sets, int K, const double *a, const double *b, const double *pi) {
    if (!ctx) return ITR_ERR_ARG;
    if (!a || !b || !pi) return fail(ctx, ITR_ERR_ARG, "itr_set_model: a, b and pi must be non-NULL");
    if (n_sets < 1) return fail(ctx, ITR_ERR_ARG, "itr_set_model: n_sets must be >= 1");
    if (K < 1 || K > ITR_MAX_STATES) return fail(ctx, ITR_ERR_ARG, "itr_set_model: K=%d outside 1..%d", K, ITR_MAX_STATES);
    CK(cudaSetDevice(ctx->device));
    const size_t na = (size_t)n_sets * K * K, nb = (size_t)n_sets * K * 256, np = (size_t)n_sets * K;
    CK(ensure(ctx->d_braw, ctx->cap_braw, na + nb + np));
    double *d_a = ctx->d_braw, *d_b = d_a + na, *d_pi = d_b + nb;
    CK(cudaMemcpyAsync(d_a, a, na * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_b, b, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_pi, pi, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc = install_model_device(ctx, n_sets, K, d_a, d_b, d_pi);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return ITR_OK;
}

extern "C" int itr_num_states(int n_int_AB, int n_int_ABC) {
    if (n_int_AB < 1 || n_int_ABC < 1) return -1;
    return n_int_AB * n_int_ABC + 3 * n_int_ABC + 3 * (n_int_ABC * (n_int_ABC - 1) / 2);
}

extern "C" int itr_plan_info(int n_int_AB, int n_int_ABC, int32_t *K, int32_t *n_mats, int64_t *n_ops,
                             int64_t *n_keys, int32_t *hidden) {
    if (n_int_AB < 1 || n_int_ABC < 1) return ITR_ERR_ARG;
    try {
        itr::ModelPlan plan;
        plan.build(n_int_AB, n_int_ABC);
        if (K) *K = plan.K;
        if (n_mats) *n_mats = plan.n_mats;
        if (n_ops) *n_ops = (int64_t)plan.ops.size();
        if (n_keys) *n_keys = plan.n_keys_max;
        if (hidden)
            for (int k = 0; k < plan.K; ++k) {
                hidden[3 * k] = plan.hidden[k].topo;
                hidden[3 * k + 1] = plan.hidden[k].i;
                hidden[3 * k + 2] = plan.hidden[k].j;
            }
    } catch (...) {
        return ITR_ERR_ARG;
    }
    return ITR_OK;
}

extern "C" int itr_build_model(itr_ctx *ctx, int n_sets, const double *params, int n_int_AB, int n_int_ABC,
                               const double *cut_AB, const double *cut_ABC, double *a, double *b, double *pi,
                               int32_t *hidden) {
    if (!ctx) return ITR_ERR_ARG;
    if (!params) return fail(ctx, ITR_ERR_ARG, "itr_build_model: params is NULL");
    if (n_sets < 1) return fail(ctx, ITR_ERR_ARG, "itr_build_model: n_sets must be >= 1");
    if (n_int_AB < 1 || n_int_ABC < 1) return fail(ctx, ITR_ERR_ARG, "itr_build_model: n_int_AB and n_int_ABC must be >= 1");
    const int K = itr_num_states(n_int_AB, n_int_ABC);
    if (K > ITR_MAX_STATES) return fail(ctx, ITR_ERR_UNSUPPORTED, "itr_build_model: K=%d exceeds %d", K, ITR_MAX_STATES);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->builder) ctx->builder = new (std::nothrow) ModelBuilder();
    if (!ctx->builder) return fail(ctx, ITR_ERR_NOMEM, "out of host memory");
    std::string msg;
    phase_begin(ctx, ITR_PH_MODEL);
    const double *d_a = nullptr, *d_b = nullptr, *d_pi = nullptr;
    int64_t launched = 0;
    int rc = ctx->builder->build(ctx->stream, n_sets, params, n_int_AB, n_int_ABC, cut_AB, cut_ABC, &d_a, &d_b, &d_pi,
                                 hidden, &launched, msg);
    phase_end(ctx, ITR_PH_MODEL);
    ctx->launches += launched;
    if (rc) return fail(ctx, rc, "itr_build_model: %s", msg.c_str());
    if (a) CK(cudaMemcpyAsync(a, d_a, (size_t)n_sets * K * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (b) CK(cudaMemcpyAsync(b, d_b, (size_t)n_sets * K * 256 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (pi) CK(cudaMemcpyAsync(pi, d_pi, (size_t)n_sets * K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    rc = install_model_device(ctx, n_sets, K, d_a, d_b, d_pi);
    if (rc) return rc;
    // Deferred mode without host outputs: the build is only enqueued; the recursions that
    // use the model wait for ev_ready on the device, itr_viterbi (which does not) overlaps it.
    if (!(ctx->async && !a && !b && !pi)) CK(cudaStreamSynchronize(ctx->stream));
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// launch geometry: one warp per chain; spread chains over SMs before stacking warps.
// ---------------------------------------------------------------------------------
struct Geometry {
    int warps, grid;
};
static Geometry geometry(const itr_ctx *ctx, int64_t n_chains, int max_warps_per_sm) {
    const int sms = ctx->prop.multiProcessorCount;
    int w = (int)std::min<int64_t>(8, std::max<int64_t>(1, (n_chains + sms - 1) / sms));
    const int64_t want = (n_chains + w - 1) / w;
    const int64_t cap = (int64_t)sms * std::max(1, max_warps_per_sm / w);
    return {w, (int)std::max<int64_t>(1, std::min(want, cap))};
}

// Work-queue counters are zeroed by a one-thread kernel, not cudaMemsetAsync: a memset may
// be executed by a copy engine, where it queues behind a posterior download in progress
// (measured: the log-likelihood of a step started only after the 2 GB transfer, +5.7 ms).
__global__ void zero_u32_kernel(unsigned int *p) { *p = 0u; }
__global__ void zero_u64_kernel(unsigned long long *p) { *p = 0ull; }
static inline void reset_queue(unsigned int *q, cudaStream_t st) { zero_u32_kernel<<<1, 1, 0, st>>>(q); }

static ChainSet chain_set(const itr_ctx *ctx, int n_sets, int slot = 0) {
    return ChainSet{ctx->d_sym, ctx->d_off, ctx->d_order, (int32_t)ctx->n_blocks, n_sets, ctx->d_queue + slot};
}

// K <= 32: columns in registers, KT = K rounded up to 4.  K > 32: NS = ceil(K/32)
// states per lane, columns streamed.
#ifdef ITR_FAST_BUILD   /* experiments only: K in 25..32 and 65..96 */
#define ITR_DISPATCH_K(K, REG, GEN)     \
    do {                                \
        if ((K) <= 28) REG(28);         \
        else if ((K) <= 32) REG(32);    \
        else GEN(3);                    \
    } while (0)
#else
#define ITR_DISPATCH_K(K, REG, GEN)     \
    do {                                \
        if ((K) <= 32) {                \
            switch (((K) + 3) / 4) {    \
                case 1: REG(4); break;  \
                case 2: REG(8); break;  \
                case 3: REG(12); break; \
                case 4: REG(16); break; \
                case 5: REG(20); break; \
                case 6: REG(24); break; \
                case 7: REG(28); break; \
                default: REG(32); break;\
            }                           \
        } else {                        \
            switch (((K) + 31) / 32) {  \
                case 2: GEN(2); break;  \
                case 3: GEN(3); break;  \
                case 4: GEN(4); break;  \
                case 5: GEN(5); break;  \
                case 6: GEN(6); break;  \
                case 7: GEN(7); break;  \
                default: GEN(8); break; \
            }                           \
        }                               \
    } while (0)
#endif

#define ITR_SWITCH_KT(K, M)             \
    switch (((K) + 3) / 4) {            \
        case 1: M(4); break;            \
        case 2: M(8); break;            \
        case 3: M(12); break;           \
        case 4: M(16); break;           \
        case 5: M(20); break;           \
        case 6: M(24); break;           \
        case 7: M(28); break;           \
        default: M(32); break;          \
    }

// 32 < K <= 96 with at least two chains per SM: the lock-step sweeps of lockstep.cu
// (ITR_LOCKSTEP=0|1 forces the choice: experiments, tests)
static bool use_lockstep(const itr_ctx *ctx, int64_t n_chains) {
    if (!lockstep_supports(ctx->K)) return false;
    const char *force = getenv("ITR_LOCKSTEP");
    if (force) return force[0] == '1';
    return n_chains >= 2 * (int64_t)ctx->prop.multiProcessorCount;
}

template <int MODE>
static void launch_forward(itr_ctx *ctx, int n_sets, double *d_ll, double *d_alpha, cudaStream_t st, int slot,
                           int first = 0, int count = -1) {
    const int K = ctx->K, KP = ctx->KP;
    ChainSet cs = chain_set(ctx, n_sets, slot);
    if (count >= 0) {           // a contiguous range of the longest-first order
        cs.order += first;
        cs.n_blocks = count;
    }
    reset_queue(cs.queue, st);
    if (MODE == 0 && use_lockstep(ctx, (int64_t)n_sets * cs.n_blocks)) {   // many chains: eight per CTA on the FP64 tensor cores
        if (ensure(ctx->d_ls_ll, ctx->cap_ls_ll, lockstep_scratch_words((int64_t)n_sets * cs.n_blocks) * 2) != cudaSuccess) return;
        (void)launch_lockstep_loglik(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, KP, d_ll, ctx->d_ls_ll, ctx->prop.multiProcessorCount, st);
        ctx->launches += 2;
        ctx->lockstep_launches += 1;
        return;
    }
    if (K > 32 && K <= 96) {        // one CTA of ceil(K/32) warps per chain, columns of a in registers
        const int grid = (int)std::min<int64_t>((int64_t)n_sets * cs.n_blocks, (int64_t)ctx->prop.multiProcessorCount * 8);
        if (K <= 64) sweep_mw_kernel<2, 0, MODE><<<grid, 64, 0, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, d_ll, d_alpha);
        else sweep_mw_kernel<3, 0, MODE><<<grid, 96, 0, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, d_ll, d_alpha);
        ctx->launches += 1;
        return;
    }
    const Geometry g = geometry(ctx, (int64_t)n_sets * cs.n_blocks, 16);
    const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
#define FWD_REG(KT) \
    forward_kernel<KT, 1, true, MODE><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, d_ll, d_alpha)
#define FWD_GEN(NS) \
    forward_kernel<4, NS, false, MODE><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, d_ll, d_alpha)
    ITR_DISPATCH_K(K, FWD_REG, FWD_GEN);
#undef FWD_REG
#undef FWD_GEN
    ctx->launches += 1;
}

static void launch_backward(itr_ctx *ctx, cudaStream_t st, int slot = 1, int first = 0, int count = -1) {
    const int K = ctx->K, KP = ctx->KP;
    ChainSet cs = chain_set(ctx, 1, slot);
    if (count >= 0) {
        cs.order += first;
        cs.n_blocks = count;
    }
    reset_queue(cs.queue, st);
    if (K > 32 && K <= 96) {
        const int grid = (int)std::min<int64_t>(cs.n_blocks, (int64_t)ctx->prop.multiProcessorCount * 8);
        if (K <= 64) sweep_mw_kernel<2, 1, 1><<<grid, 64, 0, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, nullptr, ctx->d_beta);
        else sweep_mw_kernel<3, 1, 1><<<grid, 96, 0, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, K, nullptr, ctx->d_beta);
        ctx->launches += 1;
        return;
    }
    const Geometry g = geometry(ctx, cs.n_blocks, 16);
    const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
#define BWD_REG(KT) backward_kernel<KT, 1, true><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_A, ctx->d_Et, K, ctx->d_beta)
#define BWD_GEN(NS) backward_kernel<4, NS, false><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_A, ctx->d_Et, K, ctx->d_beta)
    ITR_DISPATCH_K(K, BWD_REG, BWD_GEN);
#undef BWD_REG
#undef BWD_GEN
    ctx->launches += 1;
}

template <int COLS>
static cudaError_t launch_combine_t(itr_ctx *ctx, cudaStream_t st) {
    const size_t sh = (size_t)COLS * ctx->K * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(posterior_combine_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    posterior_combine_kernel<COLS><<<blocks_for((size_t)ctx->n_cols, COLS), COLS, sh, st>>>(ctx->d_post, ctx->d_beta, ctx->K, ctx->n_cols);
    ctx->launches += 1;
    return cudaSuccess;
}
template <int COLS>
static cudaError_t launch_combine_blocks_t(itr_ctx *ctx, cudaStream_t st, int first, int count, int64_t max_T) {
    const size_t sh = (size_t)COLS * ctx->K * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(posterior_combine_blocks_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
    posterior_combine_blocks_kernel<COLS><<<dim3(blocks_for((size_t)max_T, COLS), count), COLS, sh, st>>>(
        ctx->d_post, ctx->d_beta, ctx->K, ctx->d_off, ctx->d_order + first);
    ctx->launches += 1;
    return cudaSuccess;
}
static cudaError_t launch_combine_blocks(itr_ctx *ctx, cudaStream_t st, int first, int count, int64_t max_T) {
    if (ctx->K <= 40) return launch_combine_blocks_t<256>(ctx, st, first, count, max_T);
    if (ctx->K <= 160) return launch_combine_blocks_t<64>(ctx, st, first, count, max_T);
    return launch_combine_blocks_t<32>(ctx, st, first, count, max_T);
}
static cudaError_t launch_combine(itr_ctx *ctx, cudaStream_t st) {
    if (ctx->K <= 40) return launch_combine_t<256>(ctx, st);
    if (ctx->K <= 160) return launch_combine_t<64>(ctx, st);
    return launch_combine_t<32>(ctx, st);
}

static void launch_viterbi_forward(itr_ctx *ctx, cudaStream_t st) {
    const int K = ctx->K, KP = ctx->KP;
    const ChainSet cs = chain_set(ctx, 1, 2);
    reset_queue(cs.queue, st);
    ctx->launches += 1;
    // Few chains: four warps per chain (latency); many chains: one warp per chain (throughput).
    const int sms = ctx->prop.multiProcessorCount;
    const char *vmode = getenv("ITR_VITERBI");          // experiments / tests: "stream", "check", "4warp" (and ITR_VITERBI_1WARP)
    // (speculation pays when backpointers are stable, i.e. on alignments dominated by a few
    // symbols — the same test that enables run compression; else every window mispredicts)
    // (up to ~3 chains per SM the decoupled sweep, one CTA per chain pulled longest-first from
    // the queue, beats one warp per chain: 133 against ~360 cycles per column)
    // Chains per SM decide the CTA shape: one 16-warp CTA per SM up to one chain per SM (the
    // fastest column: 133 cycles), two 8-warp CTAs per SM (screened verifiers) up to 7
    // (hundreds of chains: a GPU's share of a chromosome split over several GPUs); with
    // thousands of chains the check-first sweep below wins (measured: tools/time_vit_modes.py;
    // ITR_VSTREAM_MAX: chains per SM up to which this sweep is used).
    static const char *smax = getenv("ITR_VSTREAM_MAX");      // experiments
    const double per_sm = (double)ctx->n_blocks / sms;
    const bool want_stream = vmode ? !strncmp(vmode, "stream", 6) : (ctx->use_runs && per_sm <= (smax ? atof(smax) : 7.0));
    if (K <= 32 && want_stream && ctx->max_T < 0x7fffffff) {
        int nw = per_sm <= 1.0 ? 16 : 8;
        if (vmode && !strcmp(vmode, "stream16")) nw = 16;
        if (vmode && !strcmp(vmode, "stream8")) nw = 8;
        static const char *sper = getenv("ITR_VSTREAM_PER");     // experiments: CTAs per SM
        const int per = sper ? atoi(sper) : nw == 16 ? 1 : 2;
        const int grid = (int)std::min<int64_t>(ctx->n_blocks, (int64_t)sms * per);
#define VSTR2(KT, NWP)                                                                                            \
    do {                                                                                                         \
        const size_t shs = StreamCfg<NWP>::SMEM;                                                                  \
        cudaFuncSetAttribute(viterbi_stream_kernel<KT, NWP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shs); \
        viterbi_stream_kernel<KT, NWP><<<grid, 32 * NWP, shs, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K,      \
                                                                    ctx->d_bp, ctx->d_final);                    \
    } while (0)
#define VSTR(KT)                      \
    do {                              \
        if (nw == 16) VSTR2(KT, 16);  \
        else VSTR2(KT, 8);            \
    } while (0)
        switch ((K + 3) / 4) {
            case 1: VSTR(4); break;
            case 2: VSTR(8); break;
            case 3: VSTR(12); break;
            case 4: VSTR(16); break;
            case 5: VSTR(20); break;
            case 6: VSTR(24); break;
            case 7: VSTR(28); break;
            default: VSTR(32); break;
        }
#undef VSTR
#undef VSTR2
        return;
    }
    if (K <= 32 && (vmode ? !strcmp(vmode, "4warp") : ctx->n_blocks <= (int64_t)4 * sms) && !getenv("ITR_VITERBI_1WARP")) {
        const int grid = (int)std::min<int64_t>(ctx->n_blocks, (int64_t)4 * sms);
        switch ((K + 7) / 8) {
            case 1: viterbi_forward4_kernel<2><<<grid, 128, 0, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final); break;
            case 2: viterbi_forward4_kernel<4><<<grid, 128, 0, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final); break;
            case 3: viterbi_forward4_kernel<6><<<grid, 128, 0, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final); break;
            default: viterbi_forward4_kernel<8><<<grid, 128, 0, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final); break;
        }
        return;
    }
    const Geometry g = geometry(ctx, ctx->n_blocks, 16);
    const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
    if (K <= 32 && (vmode ? !strncmp(vmode, "check", 5) : ctx->use_runs)) {
        // many chains, stable backpointers: check the cached pointer, exact scan only on a miss.
        // One chain per warp, CTAs of four warps; at most `wps` warps per SM: more resident
        // chains than that only stretch every chain's column latency (they share one FP64
        // pipe) while the makespan is set by the LONGEST chain — the rest waits in the queue.
        static const char *wenv = getenv("ITR_VCHK_WPS");          // experiments
        const int wps = wenv ? std::max(4, atoi(wenv)) : 12;
        Geometry g;
        g.warps = 4;
        g.grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n_blocks + 3) / 4, (int64_t)sms * (wps / 4)));
        const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
        if (vmode && !strcmp(vmode, "check64")) {        // the all-FP64 check (experiments, tests)
#define VCHK(KT) viterbi_check_kernel<KT><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final)
            ITR_SWITCH_KT(K, VCHK)
#undef VCHK
            return;
        }
        // default: the same check behind an FP32 screen
        static const char *w32 = getenv("ITR_VCHK32_WPS");          // experiments
        const int wps32 = w32 ? std::max(4, atoi(w32)) : 12;
        g.grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n_blocks + 3) / 4, (int64_t)sms * (wps32 / 4)));
        const size_t sh32 = (size_t)g.warps * 3 * KP * sizeof(double);
        la_float_transposed_kernel<<<1, 1024, 0, st>>>(ctx->d_LA, ctx->d_LAfT);
        ctx->launches += 1;
#define VCHK32(KT) viterbi_check32_kernel<KT><<<g.grid, g.warps * 32, sh32, st>>>(cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final, ctx->d_chunk_off, ctx->d_comp, ctx->d_LAfT)
        ITR_SWITCH_KT(K, VCHK32)
#undef VCHK32
        ctx->comp_done = true;          // this sweep writes the traceback's chunk composites itself
        return;
    }
#define VIT_REG(KT)                                                        \
    viterbi_forward_kernel<KT, 1, true><<<g.grid, g.warps * 32, sh, st>>>( \
        cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final)
#define VIT_GEN(NS)                                                         \
    viterbi_forward_kernel<4, NS, false><<<g.grid, g.warps * 32, sh, st>>>( \
        cs, ctx->d_LA, ctx->d_LEt, ctx->d_OM0, K, ctx->d_bp, ctx->d_final)
    ITR_DISPATCH_K(K, VIT_REG, VIT_GEN);
#undef VIT_REG
#undef VIT_GEN
}

// Tables of the run-compressed forward sweep for the resident blocks + model (three
// small launches after a model change; nothing is read back).
static int prepare_runs(itr_ctx *ctx, cudaStream_t st) {
    if (ctx->runs_valid) return ITR_OK;
    const int KP = ctx->KP;
    CK(ensure(ctx->d_P, ctx->cap_P, (size_t)ctx->n_sets * RUN_TABLE * KP * KP));
    CK(ensure(ctx->d_sP, ctx->cap_sP, (size_t)ctx->n_sets * RUN_TABLE));
    CK(ensure(ctx->d_ebar, ctx->cap_ebar, (size_t)ctx->n_sets * KP));
    symbol_class_kernel<<<1, 640, 0, st>>>(ctx->d_Et, ctx->K, KP, ctx->n_sets, 1e-12, ctx->d_rep);
    pick_run_class_kernel<<<1, 640, 0, st>>>(ctx->d_hist, ctx->d_rep, ctx->d_isrun, ctx->d_runinfo);
    run_power_kernel<<<ctx->n_sets, dim3(32, 32), 0, st>>>(ctx->d_A, ctx->d_Et, ctx->d_runinfo, KP, 0, ctx->d_P, ctx->d_sP, ctx->d_ebar);
    // backward powers (diag(e) a)^(2^k) for parameter set 0 (the posterior decodes set 0)
    CK(ensure(ctx->d_Pb, ctx->cap_Pb, (size_t)RUN_TABLE * KP * KP));
    run_power_kernel<<<1, dim3(32, 32), 0, st>>>(ctx->d_A, ctx->d_Et, ctx->d_runinfo, KP, 1, ctx->d_Pb, nullptr, nullptr);
    ctx->launches += 4;
    CK(cudaGetLastError());
    ctx->runs_valid = true;
    return ITR_OK;
}

static void launch_forward_runs(itr_ctx *ctx, double *d_ll, cudaStream_t st) {
    const int K = ctx->K, KP = ctx->KP;
    const ChainSet cs = chain_set(ctx, ctx->n_sets, 0);
    const Geometry g = geometry(ctx, (int64_t)ctx->n_sets * ctx->n_blocks, 12);
    const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
    reset_queue(cs.queue, st);
#define RUN_K(KT)                                                                                        \
    forward_runs_kernel<KT><<<g.grid, g.warps * 32, sh, st>>>(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, ctx->d_P, \
                                                              ctx->d_sP, ctx->d_ebar, ctx->d_isrun, K, d_ll)
    switch ((K + 3) / 4) {
        case 1: RUN_K(4); break;
        case 2: RUN_K(8); break;
        case 3: RUN_K(12); break;
        case 4: RUN_K(16); break;
        case 5: RUN_K(20); break;
        case 6: RUN_K(24); break;
        case 7: RUN_K(28); break;
        default: RUN_K(32); break;
    }
#undef RUN_K
    ctx->launches += 1;
}

static int need_ready(itr_ctx *ctx, const char *who) {
    if (ctx->n_blocks == 0) return fail(ctx, ITR_ERR_STATE, "%s: no blocks loaded (call itr_load_blocks first)", who);
    if (ctx->K == 0) return fail(ctx, ITR_ERR_STATE, "%s: no model installed (call itr_set_model or itr_build_model first)", who);
    return ITR_OK;
}

// Sum the per-block log-likelihoods of a finished itr_loglik into the caller's
// buffers: the reference accumulates block results in block order (optimizer.py:112-113).
static void finish_loglik(itr_ctx *ctx) {
    if (!ctx->pend_ll) return;
    // shape as it was when the request was made: the blocks or the model may have been
    // replaced since (both drain the streams first, then land here)
    const int n_sets = ctx->pend_sets;
    const int64_t n_blocks = ctx->pend_blocks;
    const size_t n = (size_t)n_sets * n_blocks;
    if (ctx->pend_per_block) memcpy(ctx->pend_per_block, ctx->h_ll, n * sizeof(double));
    if (ctx->pend_total) {
        for (int s = 0; s < n_sets; ++s) {
            double acc = 0.0;
            const double *p = ctx->h_ll + (size_t)s * n_blocks;
            for (int64_t b = 0; b < n_blocks; ++b) acc += p[b];
            ctx->pend_total[s] = acc;
        }
    }
    ctx->pend_ll = false;
    ctx->pend_total = ctx->pend_per_block = nullptr;
}

// ---------------------------------------------------------------------------------
// asynchronous mode
// ---------------------------------------------------------------------------------
extern "C" int itr_set_async(itr_ctx *ctx, int on) {
    if (!ctx) return ITR_ERR_ARG;
    if (!on) {
        int rc = itr_sync(ctx);
        if (rc) return rc;
    }
    ctx->async = on != 0;
    return ITR_OK;
}

extern "C" int itr_sync(itr_ctx *ctx) {
    if (!ctx) return ITR_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t st : {ctx->stream, ctx->s_ll, ctx->s_vit, ctx->s_post, ctx->stream2}) CK(cudaStreamSynchronize(st));
    finish_loglik(ctx);
    dump_trace(ctx);
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// forward log-likelihood
// ---------------------------------------------------------------------------------
extern "C" int itr_loglik(itr_ctx *ctx, double *total, double *per_block) {
    if (!ctx) return ITR_ERR_ARG;
    int rc = need_ready(ctx, "itr_loglik");
    if (rc) return rc;
    if (!total && !per_block) return fail(ctx, ITR_ERR_ARG, "itr_loglik: total and per_block are both NULL");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->s_ll;
    if (ctx->pend_ll) {            // a previous deferred call still owns the staging buffer
        CK(cudaStreamSynchronize(st));
        finish_loglik(ctx);
    }
    const size_t n = (size_t)ctx->n_sets * ctx->n_blocks;
    CK(ensure(ctx->d_ll, ctx->cap_ll, n));
    if (n > ctx->cap_hll) {
        if (ctx->h_ll) cudaFreeHost(ctx->h_ll);
        ctx->h_ll = nullptr;
        ctx->cap_hll = 0;
        CK(cudaMallocHost((void **)&ctx->h_ll, n * sizeof(double)));
        ctx->cap_hll = n;
    }
    CK(cudaStreamWaitEvent(st, ctx->ev_ready, 0));
    // (No wait for a posterior download here, unlike itr_viterbi: this sweep is one warp per
    // block and does not crowd out the posterior's kernels — and measured, this stream's
    // wait on ev_post_compute was only released when the whole download had finished.)
    const bool runs = ctx->K <= 32 && ctx->use_runs && ctx->runs_valid && !getenv("ITR_NO_RUNS");   // (variable: experiments, tests)
    phase_begin(ctx, ITR_PH_LOGLIK, st);
    if (runs) launch_forward_runs(ctx, ctx->d_ll, st);
    else launch_forward<0>(ctx, ctx->n_sets, ctx->d_ll, nullptr, st, 0);
    phase_end(ctx, ITR_PH_LOGLIK, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_ll, ctx->d_ll, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    ctx->pend_total = total;
    ctx->pend_per_block = per_block;
    ctx->pend_sets = ctx->n_sets;
    ctx->pend_blocks = ctx->n_blocks;
    ctx->pend_ll = true;
    if (!ctx->async) {
        CK(cudaStreamSynchronize(st));
        finish_loglik(ctx);
    }
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// Viterbi
// ---------------------------------------------------------------------------------
template <typename F>
static cudaError_t chunk_smem(F kernel, size_t per_warp, int *warps, size_t *bytes) {
    int w = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(96 * 1024) / per_warp));
    *warps = w;
    *bytes = per_warp * w;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*bytes);
}

extern "C" int itr_viterbi(itr_ctx *ctx, const double *log_a, const double *log_E, const double *omega0,
                           uint8_t *path) {
    if (!ctx) return ITR_ERR_ARG;
    int rc = need_ready(ctx, "itr_viterbi");
    if (rc) return rc;
    if (!log_a || !log_E || !omega0) return fail(ctx, ITR_ERR_ARG, "itr_viterbi: log_a, log_E and omega0 must be non-NULL");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->s_vit;
    const int K = ctx->K, KP = ctx->KP;
    const int64_t nb = ctx->n_blocks;
    const double ninf = -std::numeric_limits<double>::infinity();
    CK(cudaStreamSynchronize(st));      // buffers below may still be in use by a deferred call
    CK(ensure(ctx->d_LA, ctx->cap_LA, (size_t)KP * KP));
    CK(ensure(ctx->d_LEt, ctx->cap_LEt, (size_t)NSYM * KP));
    CK(ensure(ctx->d_OM0, ctx->cap_OM0, (size_t)nb * KP));
    const size_t n_la = (size_t)K * K, n_le = (size_t)K * NSYM, n_om = (size_t)nb * K;
    CK(ensure(ctx->d_tmp, ctx->cap_tmp, n_la + n_le + n_om));
    CK(ensure(ctx->d_bp, ctx->cap_bp, (size_t)(ctx->n_cols + 1) * KP));
    CK(ensure(ctx->d_comp, ctx->cap_comp, (size_t)(ctx->n_chunks + 1) * KP));
    CK(ensure(ctx->d_chunk_end, ctx->cap_chunk_end, (size_t)ctx->n_chunks + 1));
    CK(ensure(ctx->d_path, ctx->cap_path, (size_t)ctx->n_cols));
    CK(ensure(ctx->d_final, ctx->cap_final, (size_t)nb));
    double *t_la = ctx->d_tmp, *t_le = t_la + n_la, *t_om = t_le + n_le;
    // No wait for ev_ready: this recursion reads only the resident blocks (installed
    // synchronously) and the caller's log tables, nothing of the device model — so it
    // overlaps a model build that is still in flight (itr_build_model in the same step).
    // A posterior on its way to the host is PCIe bound (2 GB); its kernels go first, so
    // that the download starts early, and this recursion runs under the transfer.
    if (ctx->post_download) CK(cudaStreamWaitEvent(st, ctx->ev_post_compute, 0));
    // With thousands of chains the sweep fills every SM for tens of milliseconds (one
    // register-limited CTA each), and a model build enqueued just before it — a chain of
    // small kernels — would crawl behind it with the log-likelihood and the posterior
    // waiting for the model (measured at config 4: build 1 -> 57 ms inside a step).  There the
    // sweep starts after the build; with fewer chains than SMs (config 2) it overlaps it as before.
    if (ctx->n_blocks >= (int64_t)ctx->prop.multiProcessorCount) CK(cudaStreamWaitEvent(st, ctx->ev_ready, 0));
    CK(cudaMemcpyAsync(t_la, log_a, n_la * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t_le, log_E, n_le * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t_om, omega0, n_om * sizeof(double), cudaMemcpyHostToDevice, st));
    pad_kernel<<<blocks_for((size_t)KP * KP, 256), 256, 0, st>>>(t_la, ctx->d_LA, 1, K, K, KP, KP, ninf);
    transpose_table_kernel<<<blocks_for((size_t)NSYM * KP, 256), 256, 0, st>>>(t_le, ctx->d_LEt, K, KP, 0.0);
    pad_kernel<<<blocks_for((size_t)nb * KP, 256), 256, 0, st>>>(t_om, ctx->d_OM0, (int)nb, 1, K, 1, KP, ninf);
    ctx->launches += 3;
    phase_begin(ctx, ITR_PH_VITERBI_FWD, st);
    ctx->comp_done = false;
    launch_viterbi_forward(ctx, st);
    phase_end(ctx, ITR_PH_VITERBI_FWD, st);
    phase_begin(ctx, ITR_PH_VITERBI_TRACE, st);
    {
        int w1 = 1, w2 = 1;
        size_t sh1 = 0, sh2 = 0;
        CK(chunk_smem(viterbi_compose_kernel, (size_t)VCHUNK * KP, &w1, &sh1));
        CK(chunk_smem(viterbi_traceback_kernel, (size_t)VCHUNK * KP + VCHUNK, &w2, &sh2));
        if (!ctx->comp_done)
            viterbi_compose_kernel<<<blocks_for((size_t)ctx->n_chunks, w1), w1 * 32, sh1, st>>>(
                ctx->d_off, ctx->d_chunk_off, ctx->d_chunk_blk, ctx->d_bp, KP, K, ctx->n_chunks, ctx->d_comp);
        viterbi_boundary_kernel<<<blocks_for((size_t)nb, VB_WARPS), 32 * VB_WARPS, 0, st>>>(ctx->d_off, ctx->d_chunk_off, ctx->d_comp,
                                                                            ctx->d_final, KP, (int)nb, ctx->d_chunk_end);
        viterbi_traceback_kernel<<<blocks_for((size_t)ctx->n_chunks, w2), w2 * 32, sh2, st>>>(
            ctx->d_off, ctx->d_chunk_off, ctx->d_chunk_blk, ctx->d_bp, ctx->d_chunk_end, KP, ctx->n_chunks, ctx->d_path);
    }
    phase_end(ctx, ITR_PH_VITERBI_TRACE, st);
    ctx->launches += ctx->comp_done ? 2 : 3;
    CK(cudaGetLastError());
    ctx->have_path = true;
    if (path) CK(cudaMemcpyAsync(path, ctx->d_path, (size_t)ctx->n_cols, cudaMemcpyDeviceToHost, st));
    if (!ctx->async) CK(cudaStreamSynchronize(st));
    return ITR_OK;
}

extern "C" int itr_viterbi_fetch(itr_ctx *ctx, uint8_t *path) {
    if (!ctx) return ITR_ERR_ARG;
    if (!path) return fail(ctx, ITR_ERR_ARG, "itr_viterbi_fetch: path is NULL");
    if (!ctx->have_path) return fail(ctx, ITR_ERR_STATE, "itr_viterbi_fetch: no Viterbi result on the device");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(path, ctx->d_path, (size_t)ctx->n_cols, cudaMemcpyDeviceToHost, ctx->s_vit));
    CK(cudaStreamSynchronize(ctx->s_vit));
    return ITR_OK;
}

extern "C" int itr_viterbi_fetch_range(itr_ctx *ctx, int64_t col0, int64_t n_cols, uint8_t *path) {
    if (!ctx) return ITR_ERR_ARG;
    if (!path && n_cols > 0) return fail(ctx, ITR_ERR_ARG, "itr_viterbi_fetch_range: path is NULL");
    if (!ctx->have_path) return fail(ctx, ITR_ERR_STATE, "itr_viterbi_fetch_range: no Viterbi result on the device");
    if (col0 < 0 || n_cols < 0 || col0 + n_cols > ctx->n_cols)
        return fail(ctx, ITR_ERR_ARG, "itr_viterbi_fetch_range: column range outside the loaded alignment");
    if (n_cols == 0) return ITR_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(path, ctx->d_path + col0, (size_t)n_cols, cudaMemcpyDeviceToHost, ctx->s_vit));
    CK(cudaStreamSynchronize(ctx->s_vit));
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// two-pass posterior launches (K <= 32, run-dominated alignments)
// ---------------------------------------------------------------------------------
// Pass 1, one direction (dir 0: forward checkpoints, 1: backward) over the chains of `cs`.
static void launch_checkpoint_sweep(itr_ctx *ctx, int dir, const ChainSet &cs, cudaStream_t st) {
    const int K = ctx->K, KP = ctx->KP;
    // many chains: the 128-register build at two CTAs per SM (throughput); else the
    // spill-free build (latency)
    static const char *force = getenv("ITR_SWEEP_REGS");           // experiments: "128" | "192"
    const bool dense = force ? !strcmp(force, "128") : cs.n_blocks > 8 * ctx->prop.multiProcessorCount;
    const Geometry g = geometry(ctx, cs.n_blocks, dense ? 16 : 12);
    const size_t sh = (size_t)g.warps * 2 * KP * sizeof(double);
    reset_queue(cs.queue, st);
#define CKS2(KT, DIR, REGS)                                                                                          \
    checkpoint_sweep_kernel<KT, DIR, REGS><<<g.grid, g.warps * 32, sh, st>>>(                                        \
        cs, ctx->d_A, ctx->d_PI, ctx->d_Et, DIR ? ctx->d_Pb : ctx->d_P, ctx->d_ebar, ctx->d_isrun, ctx->d_tile_off, K, \
        DIR ? ctx->d_ck_b : ctx->d_ck_a)
#define CKS(KT)                                                      \
    do {                                                             \
        if (dir) { if (dense) CKS2(KT, 1, 128); else CKS2(KT, 1, 192); } \
        else     { if (dense) CKS2(KT, 0, 128); else CKS2(KT, 0, 192); } \
    } while (0)
    ITR_SWITCH_KT(K, CKS)
#undef CKS
#undef CKS2
    ctx->launches += 1;
}

// Pass 2 over the blocks [b0, b1) (tile ids follow the input block order, so this is a
// contiguous range of tiles and of result rows).  K <= 28: full 32-column tiles on the FP64
// tensor cores, eight tiles per warp in lock step (posterior_tiles_mma_kernel); the
// partial last tile of each block — and everything when K > 28 or ITR_POST_TILES=fma —
// through the one-warp-per-tile FMA kernel.
static cudaError_t launch_post_tiles(itr_ctx *ctx, cudaStream_t st, int64_t b0, int64_t b1, bool dynamic_ok = false) {
    const int64_t t0 = ctx->h_tile_off[b0], t1 = ctx->h_tile_off[b1];
    if (t1 <= t0) return cudaSuccess;
    const int K = ctx->K, KP = ctx->KP;
    const int sms = ctx->prop.multiProcessorCount;
    const int wt = 4;
    const size_t sht = (size_t)wt * (2 * KP + PTILE * (KP + 1) + PTILE) * sizeof(double);
    static const char *mode = getenv("ITR_POST_TILES");          // experiments / tests: "fma" | "mma"
    const bool use_mma = K <= 28 && !(mode && !strcmp(mode, "fma"));
    cudaError_t e = cudaSuccess;
    if (use_mma) {
        const size_t shm = (size_t)MMA_WARPS * 4 * (PTILE * K + 4) * sizeof(double);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024 - 1024) / (shm + 1024)));
        const int64_t groups = (t1 - t0 + 3) / 4;
        const unsigned gm = (unsigned)std::min<int64_t>((groups + MMA_WARPS - 1) / MMA_WARPS, (int64_t)sms * per_sm);
        // more groups than warps: tickets instead of a static split (see the kernel); the launches
        // that take this path are ordered on one stream, so one counter, zeroed in-stream, serves
        static const char *stat = getenv("ITR_POST_STATIC");        // experiments
        unsigned long long *tk = (dynamic_ok && !stat && groups > (int64_t)gm * MMA_WARPS) ? ctx->d_tile_ticket : nullptr;
        if (tk) {
            zero_u64_kernel<<<1, 1, 0, st>>>(tk);
            ctx->launches += 1;
        }
#define PTM(KT)                                                                                                          \
    do {                                                                                                                 \
        if constexpr (KT <= 28) {                                                                                        \
            e = cudaFuncSetAttribute(posterior_tiles_mma_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm); \
            if (e == cudaSuccess)                                                                                        \
                posterior_tiles_mma_kernel<KT><<<gm, 32 * MMA_WARPS, shm, st>>>(ctx->d_sym, ctx->d_tile_info, t0, t1, \
                                                                         ctx->d_A, ctx->d_PI, ctx->d_Et, ctx->d_ck_a, ctx->d_ck_b, K,   \
                                                                         ctx->d_post, tk);                               \
        }                                                                                                                \
    } while (0)
        ITR_SWITCH_KT(K, PTM)
#undef PTM
        ctx->launches += 1;
        if (e != cudaSuccess) return e;
    }
    // FMA kernel: every tile (list == nullptr), or the partial last tiles of the blocks
    const int64_t *list = use_mma ? ctx->d_part_tile : nullptr;
    const int64_t l0 = use_mma ? b0 : t0, l1 = use_mma ? b1 : t1;
    if (use_mma) {
        bool any = false;
        for (int64_t b = b0; b < b1 && !any; ++b) any = (ctx->h_off[b + 1] - ctx->h_off[b]) % PTILE != 0;
        if (!any) return cudaSuccess;
    }
    const unsigned gt = (unsigned)std::min<int64_t>((l1 - l0 + wt - 1) / wt, (int64_t)sms * 6);
#define PT2(KT)                                                                                                      \
    do {                                                                                                             \
        e = cudaFuncSetAttribute(posterior_tiles_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sht); \
        if (e == cudaSuccess)                                                                                        \
            posterior_tiles_kernel<KT><<<gt, wt * 32, sht, st>>>(ctx->d_sym, ctx->d_off, ctx->d_tile_off, ctx->d_tile_blk, l0, l1, list, \
                                                                 ctx->d_A, ctx->d_PI, ctx->d_Et, ctx->d_ck_a, ctx->d_ck_b, K,       \
                                                                 ctx->d_post);                                       \
    } while (0)
    ITR_SWITCH_KT(K, PT2)
#undef PT2
    ctx->launches += 1;
    return e;
}

// ---------------------------------------------------------------------------------
// posterior
// ---------------------------------------------------------------------------------
extern "C" int itr_posterior(itr_ctx *ctx, double *post) {
    if (!ctx) return ITR_ERR_ARG;
    int rc = need_ready(ctx, "itr_posterior");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->s_post;
    const size_t n = (size_t)ctx->n_cols * ctx->K;
    CK(cudaStreamSynchronize(st));
    ctx->ranges_valid = false;
    CK(ensure(ctx->d_post, ctx->cap_post, n));
    if (use_lockstep(ctx, ctx->n_blocks)) {
        // many blocks at 32 < K <= 96: four blocks per CTA walked from both ends on the FP64
        // tensor cores, the two directions meeting in the result matrix (lockstep.cu)
        CK(cudaStreamWaitEvent(st, ctx->ev_ready, 0));
        phase_begin(ctx, ITR_PH_POST_TOTAL, st);
        phase_begin(ctx, ITR_PH_POST_COMBINE, st);
        const ChainSet cs = chain_set(ctx, 1, 3);
        reset_queue(cs.queue, st);
        CK(ensure(ctx->d_ls_post, ctx->cap_ls_post, lockstep_scratch_words(ctx->n_blocks)));
        CK(launch_lockstep_posterior(cs, ctx->d_A, ctx->d_PI, ctx->d_Et, ctx->K, ctx->KP, ctx->d_post, ctx->d_ls_post,
                                     ctx->prop.multiProcessorCount, st));
        ctx->launches += 2;
        ctx->lockstep_launches += 1;
        phase_end(ctx, ITR_PH_POST_COMBINE, st);
        phase_end(ctx, ITR_PH_POST_TOTAL, st);
        ctx->have_post = true;
        ctx->post_download = false;
        if (post) CK(cudaMemcpyAsync(post, ctx->d_post, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (!ctx->async) CK(cudaStreamSynchronize(st));
        return ITR_OK;
    }
    if (!(ctx->K <= 32 && ctx->use_runs && ctx->runs_valid && !getenv("ITR_NO_RUNS"))) CK(ensure(ctx->d_beta, ctx->cap_beta, n));
    // forward (alpha -> d_post) on the posterior stream, backward (beta -> d_beta) on
    // the second stream, concurrently; then the combine on the posterior stream.
    CK(cudaStreamWaitEvent(st, ctx->ev_ready, 0));
    if (ctx->K <= 32 && ctx->use_runs && ctx->runs_valid && !getenv("ITR_NO_RUNS")) {
        // two-pass posterior: checkpoint sweeps (forward || backward), then one warp per tile
        const int K = ctx->K, KP = ctx->KP;
        CK(ensure(ctx->d_ck_a, ctx->cap_ck_a, (size_t)(ctx->n_tiles + 1) * KP));
        CK(ensure(ctx->d_ck_b, ctx->cap_ck_b, (size_t)(ctx->n_tiles + 1) * KP));
        phase_begin(ctx, ITR_PH_POST_TOTAL, st);
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        // Device destination: one pair of sweeps over all blocks, then one pass-2 launch.
        // Host destination: blocks are independent, so they run in groups of similar
        // length (contiguous ranges of the longest-first order), shortest group first,
        // each group's sweeps on its own pair of streams; as soon as a group's
        // checkpoints exist its blocks go through pass 2 and are downloaded on a copy
        // stream — the 2 GB PCIe transfer starts after the SHORTEST blocks' sweeps and
        // hides the long blocks' latency-bound sweeps.
        const int nb = (int)ctx->n_blocks;
        // (with hundreds of chains the sweeps fill the GPU by themselves: one group, one pass-2 launch)
        const int n_groups = (getenv("ITR_POST_ONE_GROUP") || ctx->stream_ranges > 0 || (!post && nb > 256)) ? 1 : std::min(8, std::max(1, nb / 2));
        while ((int)ctx->grp_streams.size() < 2 * 8 + 1) {
            cudaStream_t s2 = nullptr;
            cudaEvent_t e2 = nullptr;
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            ctx->grp_streams.push_back(s2);
            CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            ctx->grp_events.push_back(e2);
        }
        const bool trace = post && getenv("ITR_POST_TRACE");      // debug: per-block timeline on stderr
        if (trace) {
            dump_trace(ctx);
            cudaEventCreate(&ctx->tr_0);
            cudaEventRecord(ctx->tr_0, st);
        }
        cudaStream_t scopy = ctx->grp_streams[16];
        cudaEvent_t ecopy = ctx->grp_events[16];
        for (int gi = n_groups - 1; gi >= 0; --gi) {
            const int first = (int)((int64_t)nb * gi / n_groups), last = (int)((int64_t)nb * (gi + 1) / n_groups);
            if (last <= first) continue;
            const bool whole = (n_groups == 1);
            cudaStream_t sf = whole ? st : ctx->grp_streams[2 * gi], sb = whole ? ctx->stream2 : ctx->grp_streams[2 * gi + 1];
            ChainSet cf = chain_set(ctx, 1, whole ? 3 : 24 + 2 * gi), cb = chain_set(ctx, 1, whole ? 1 : 25 + 2 * gi);
            cf.order += first; cf.n_blocks = last - first;
            cb.order += first; cb.n_blocks = last - first;
            if (!whole) {
                CK(cudaStreamWaitEvent(sf, ctx->ev_fork, 0));
                CK(cudaStreamWaitEvent(sb, ctx->ev_fork, 0));
            }
            if (gi == 0) phase_begin(ctx, ITR_PH_POST_BWD, sb);
            launch_checkpoint_sweep(ctx, 1, cb, sb);
            if (gi == 0) phase_end(ctx, ITR_PH_POST_BWD, sb);
            if (gi == 0) phase_begin(ctx, ITR_PH_POST_FWD, sf);
            launch_checkpoint_sweep(ctx, 0, cf, sf);
            if (gi == 0) phase_end(ctx, ITR_PH_POST_FWD, sf);
            CK(cudaEventRecord(whole ? ctx->ev_join : ctx->grp_events[2 * gi + 1], sb));
            CK(cudaStreamWaitEvent(sf, whole ? ctx->ev_join : ctx->grp_events[2 * gi + 1], 0));
        }
        // second loop: a download into pageable memory blocks the host, so every sweep is enqueued first
        for (int gi = n_groups - 1; gi >= 0; --gi) {
            const int first = (int)((int64_t)nb * gi / n_groups), last = (int)((int64_t)nb * (gi + 1) / n_groups);
            if (last <= first) continue;
            const bool whole = (n_groups == 1);
            cudaStream_t sf = whole ? st : ctx->grp_streams[2 * gi];
            if (whole) {
                phase_begin(ctx, ITR_PH_POST_COMBINE, st);
                // Pass 2 over contiguous ranges of blocks (input order = contiguous rows of the
                // result): one range normally; several when the result is streamed to the host
                // (itr_posterior_stream), each followed by an event the copy stream waits for.
                // (a deferred call that keeps the result on the device may be drained by
                // itr_posterior_stream later: give it ranges too)
                const int want_ranges = ctx->stream_ranges > 0 ? ctx->stream_ranges : (ctx->async && !post ? 16 : 0);
                const int n_ranges = std::max(1, std::min<int>(want_ranges, nb));
                ctx->range_col_end.clear();
                for (int r = 0; r < n_ranges; ++r) {
                    const int64_t b0 = (int64_t)nb * r / n_ranges, b1 = (int64_t)nb * (r + 1) / n_ranges;
                    if (b1 <= b0) continue;
                    CK(launch_post_tiles(ctx, st, b0, b1, true));
                    if (want_ranges > 0) {
                        while (ctx->range_events.size() <= ctx->range_col_end.size()) {
                            cudaEvent_t e = nullptr;
                            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                            ctx->range_events.push_back(e);
                        }
                        CK(cudaEventRecord(ctx->range_events[ctx->range_col_end.size()], st));
                        ctx->range_col_end.push_back(ctx->h_off[b1]);
                    }
                }
                phase_end(ctx, ITR_PH_POST_COMBINE, st);
                ctx->ranges_valid = want_ranges > 0;
                if (post) {
                    CK(cudaEventRecord(ecopy, st));
                    CK(cudaStreamWaitEvent(scopy, ecopy, 0));
                    CK(cudaMemcpyAsync(post, ctx->d_post, n * sizeof(double), cudaMemcpyDeviceToHost, scopy));
                }
            } else {
                // pass 2 and download block by block, in the group's forward stream
                if (gi == 0) phase_begin(ctx, ITR_PH_POST_COMBINE, sf);
                for (int q = last - 1; q >= first; --q) {
                    const int32_t blk = ctx->h_order[q];
                    CK(launch_post_tiles(ctx, sf, blk, blk + 1));
                    if (!post) continue;
                    CK(cudaEventRecord(ctx->grp_events[2 * gi], sf));
                    if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, sf); ctx->tr_k.push_back(e); }
                    CK(cudaStreamWaitEvent(scopy, ctx->grp_events[2 * gi], 0));
                    const size_t o = (size_t)ctx->h_off[blk] * K, len = (size_t)(ctx->h_off[blk + 1] - ctx->h_off[blk]) * K;
                    CK(cudaMemcpyAsync(post + o, ctx->d_post + o, len * sizeof(double), cudaMemcpyDeviceToHost, scopy));
                    if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, scopy); ctx->tr_c.push_back(e); ctx->tr_b.push_back(blk); }
                }
                if (gi == 0) phase_end(ctx, ITR_PH_POST_COMBINE, sf);
                if (post) {                        // stream2 collects "kernels done" of every group
                    CK(cudaEventRecord(ctx->grp_events[2 * gi + 1], sf));
                    CK(cudaStreamWaitEvent(ctx->stream2, ctx->grp_events[2 * gi + 1], 0));
                }
                if (!post) {                       // join the group into the copy stream's place: st waits below
                    CK(cudaEventRecord(ctx->grp_events[2 * gi], sf));
                    CK(cudaStreamWaitEvent(st, ctx->grp_events[2 * gi], 0));
                }
            }
        }
        CK(cudaGetLastError());
        ctx->have_post = true;
        ctx->post_download = false;
        if (post && n_groups > 1) {
            CK(cudaEventRecord(ctx->ev_post_compute, ctx->stream2));
            ctx->post_download = true;
        }
        if (post) {                       // the posterior stream also waits for the downloads
            CK(cudaEventRecord(ecopy, scopy));
            CK(cudaStreamWaitEvent(st, ecopy, 0));
        }
        phase_end(ctx, ITR_PH_POST_TOTAL, st);
        if (trace && !ctx->async) dump_trace(ctx);
        if (!ctx->async) CK(cudaStreamSynchronize(st));
        return ITR_OK;
    }
    const int n_groups = (post && ctx->n_blocks >= 16 && !getenv("ITR_POST_ONE_GROUP")) ? 8 : 1;
    if (n_groups > 1) {
        // Blocks are independent: run them in groups of similar length (contiguous ranges
        // of the longest-first order), each group on its own pair of streams, and start
        // the download of a group as soon as it is done — the PCIe transfer of the short
        // blocks overlaps the recursions of the long ones.
        while ((int)ctx->grp_streams.size() < 2 * n_groups) {
            cudaStream_t s2 = nullptr;
            cudaEvent_t e2 = nullptr;
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            ctx->grp_streams.push_back(s2);
            CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            ctx->grp_events.push_back(e2);
        }
        phase_begin(ctx, ITR_PH_POST_TOTAL, st);
        CK(cudaEventRecord(ctx->ev_fork, st));
        const int nb = (int)ctx->n_blocks;
        for (int g = n_groups - 1; g >= 0; --g) {          // shortest group first
            const int first = (int)((int64_t)nb * g / n_groups), last = (int)((int64_t)nb * (g + 1) / n_groups);
            const int count = last - first;
            if (count <= 0) continue;
            cudaStream_t sa = ctx->grp_streams[2 * g], sb = ctx->grp_streams[2 * g + 1];
            cudaEvent_t ea = ctx->grp_events[2 * g], eb = ctx->grp_events[2 * g + 1];
            CK(cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
            CK(cudaStreamWaitEvent(sb, ctx->ev_fork, 0));
            if (g == 0) phase_begin(ctx, ITR_PH_POST_BWD, sb);
            launch_backward(ctx, sb, 8 + 2 * g, first, count);
            if (g == 0) phase_end(ctx, ITR_PH_POST_BWD, sb);
            CK(cudaEventRecord(eb, sb));
            if (g == 0) phase_begin(ctx, ITR_PH_POST_FWD, sa);
            launch_forward<1>(ctx, 1, nullptr, ctx->d_post, sa, 9 + 2 * g, first, count);
            if (g == 0) phase_end(ctx, ITR_PH_POST_FWD, sa);
            CK(cudaStreamWaitEvent(sa, eb, 0));
            const int32_t longest = ctx->h_order[first];
            if (g == 0) phase_begin(ctx, ITR_PH_POST_COMBINE, sa);
            CK(launch_combine_blocks(ctx, sa, first, count, ctx->h_off[longest + 1] - ctx->h_off[longest]));
            if (g == 0) phase_end(ctx, ITR_PH_POST_COMBINE, sa);
        }
        // downloads, shortest group first (issued after every kernel is enqueued, so that a
        // pageable destination — whose copies block the host — does not delay launches)
        for (int g = n_groups - 1; g >= 0; --g) {
            const int first = (int)((int64_t)nb * g / n_groups), last = (int)((int64_t)nb * (g + 1) / n_groups);
            if (last <= first) continue;
            cudaStream_t sa = ctx->grp_streams[2 * g];
            cudaEvent_t ea = ctx->grp_events[2 * g];
            for (int q = first; q < last; ++q) {
                const int32_t blk = ctx->h_order[q];
                const size_t o = (size_t)ctx->h_off[blk] * ctx->K, len = (size_t)(ctx->h_off[blk + 1] - ctx->h_off[blk]) * ctx->K;
                CK(cudaMemcpyAsync(post + o, ctx->d_post + o, len * sizeof(double), cudaMemcpyDeviceToHost, sa));
            }
            CK(cudaEventRecord(ea, sa));
            CK(cudaStreamWaitEvent(st, ea, 0));
        }
        phase_end(ctx, ITR_PH_POST_TOTAL, st);
        CK(cudaGetLastError());
        ctx->have_post = true;
        if (!ctx->async) CK(cudaStreamSynchronize(st));
        return ITR_OK;
    }
    phase_begin(ctx, ITR_PH_POST_TOTAL, st);
    CK(cudaEventRecord(ctx->ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    phase_begin(ctx, ITR_PH_POST_BWD, ctx->stream2);
    launch_backward(ctx, ctx->stream2);
    phase_end(ctx, ITR_PH_POST_BWD, ctx->stream2);
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
    phase_begin(ctx, ITR_PH_POST_FWD, st);
    launch_forward<1>(ctx, 1, nullptr, ctx->d_post, st, 3);
    phase_end(ctx, ITR_PH_POST_FWD, st);
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    phase_begin(ctx, ITR_PH_POST_COMBINE, st);
    CK(launch_combine(ctx, st));
    phase_end(ctx, ITR_PH_POST_COMBINE, st);
    phase_end(ctx, ITR_PH_POST_TOTAL, st);
    CK(cudaGetLastError());
    ctx->have_post = true;
    if (post) CK(cudaMemcpyAsync(post, ctx->d_post, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!ctx->async) CK(cudaStreamSynchronize(st));
    return ITR_OK;
}

extern "C" int itr_posterior_fetch(itr_ctx *ctx, double *post) {
    if (!ctx) return ITR_ERR_ARG;
    if (!post) return fail(ctx, ITR_ERR_ARG, "itr_posterior_fetch: post is NULL");
    if (!ctx->have_post) return fail(ctx, ITR_ERR_STATE, "itr_posterior_fetch: no posterior on the device");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(post, ctx->d_post, (size_t)ctx->n_cols * ctx->K * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_post));
    CK(cudaStreamSynchronize(ctx->s_post));
    return ITR_OK;
}

extern "C" int itr_posterior_fetch_range(itr_ctx *ctx, int64_t col0, int64_t n_cols, double *post) {
    if (!ctx) return ITR_ERR_ARG;
    if (!post && n_cols > 0) return fail(ctx, ITR_ERR_ARG, "itr_posterior_fetch_range: post is NULL");
    if (!ctx->have_post) return fail(ctx, ITR_ERR_STATE, "itr_posterior_fetch_range: no posterior on the device");
    if (col0 < 0 || n_cols < 0 || col0 + n_cols > ctx->n_cols)
        return fail(ctx, ITR_ERR_ARG, "itr_posterior_fetch_range: column range outside the loaded alignment");
    if (n_cols == 0) return ITR_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(post, ctx->d_post + (size_t)col0 * ctx->K, (size_t)n_cols * ctx->K * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->s_post));
    CK(cudaStreamSynchronize(ctx->s_post));
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// posterior streamed to the host through a bounded ring (chromosome-scale results)
// ---------------------------------------------------------------------------------
extern "C" int itr_posterior_stream(itr_ctx *ctx, double *ring, int64_t slot_cols, int n_slots,
                                    itr_rows_sink sink, void *user) {
    if (!ctx) return ITR_ERR_ARG;
    if (!ring || slot_cols < 1 || n_slots < 1)
        return fail(ctx, ITR_ERR_ARG, "itr_posterior_stream: ring must be non-NULL with slot_cols >= 1 and n_slots >= 1");
    if (n_slots > 64) n_slots = 64;
    int rc = need_ready(ctx, "itr_posterior_stream");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    // Pass 2 in ranges of ~1/16 of the alignment (at least one block each), so that the
    // first rows leave the device while later ranges are still being computed.
    // A deferred itr_posterior(ctx, NULL) that is still in flight (or just finished) is the
    // computation that gets drained: the caller can put the posterior's kernels first, enqueue
    // the other recursions behind them, and then start the download loop.
    if (!(ctx->async && ctx->have_post && ctx->ranges_valid)) {
        const bool was_async = ctx->async;
        ctx->async = true;
        ctx->stream_ranges = 16;
        rc = itr_posterior(ctx, nullptr);
        ctx->stream_ranges = 0;
        ctx->async = was_async;
        if (rc) return rc;
    }
    const int K = ctx->K;
    const int64_t n_cols = ctx->n_cols;
    while ((int)ctx->grp_streams.size() < 2 * 8 + 1) {
        cudaStream_t s2 = nullptr;
        cudaEvent_t e2 = nullptr;
        CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        ctx->grp_streams.push_back(s2);
        CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        ctx->grp_events.push_back(e2);
    }
    cudaStream_t scopy = ctx->grp_streams[16];
    std::vector<cudaEvent_t> slot_ev(n_slots, nullptr);
    for (int q = 0; q < n_slots; ++q) CK(cudaEventCreateWithFlags(&slot_ev[q], cudaEventDisableTiming));
    struct Piece { int64_t col0, n; };
    std::vector<Piece> inflight(n_slots, Piece{0, 0});
    auto cleanup = [&]() {
        cudaStreamSynchronize(scopy);
        for (cudaEvent_t e : slot_ev)
            if (e) cudaEventDestroy(e);
    };
    // ranges computed by the two-pass path carry events; any other path: the whole result
    // is complete when the posterior stream is (one wait)
    const bool ranged = ctx->ranges_valid && !ctx->range_col_end.empty() && ctx->range_col_end.back() == n_cols;
    if (!ranged) {
        CK(cudaEventRecord(ctx->ev_join, ctx->s_post));
        CK(cudaStreamWaitEvent(scopy, ctx->ev_join, 0));
    }
    size_t next_range = 0;
    int64_t piece = 0;
    int status = ITR_OK;
    auto retire = [&](int slot) -> int {          // the copy into `slot` is complete: hand it to the sink
        if (inflight[slot].n == 0) return 0;
        if (cudaEventSynchronize(slot_ev[slot]) != cudaSuccess) return fail(ctx, ITR_ERR_CUDA, "itr_posterior_stream: download failed");
        const Piece pc = inflight[slot];
        inflight[slot].n = 0;
        if (sink && sink(user, pc.col0, pc.n, ring + (size_t)slot * slot_cols * K) != 0)
            return fail(ctx, ITR_ERR_IO, "itr_posterior_stream: the sink asked to stop at column %lld", (long long)pc.col0);
        return 0;
    };
    for (int64_t c0 = 0; c0 < n_cols && status == ITR_OK; ++piece) {
        // a piece never crosses a range boundary, so it waits for exactly one range
        int64_t lim = n_cols;
        if (ranged) {
            while (ctx->range_col_end[next_range] <= c0) ++next_range;
            lim = ctx->range_col_end[next_range];
        }
        const int64_t n = std::min(slot_cols, lim - c0);
        const int slot = (int)(piece % n_slots);
        if ((status = retire(slot)) != ITR_OK) break;
        if (ranged && cudaStreamWaitEvent(scopy, ctx->range_events[next_range], 0) != cudaSuccess) {
            status = fail(ctx, ITR_ERR_CUDA, "itr_posterior_stream: cudaStreamWaitEvent failed");
            break;
        }
        if (cudaMemcpyAsync(ring + (size_t)slot * slot_cols * K, ctx->d_post + (size_t)c0 * K, (size_t)n * K * sizeof(double),
                            cudaMemcpyDeviceToHost, scopy) != cudaSuccess ||
            cudaEventRecord(slot_ev[slot], scopy) != cudaSuccess) {
            status = fail(ctx, ITR_ERR_CUDA, "itr_posterior_stream: download failed");
            break;
        }
        inflight[slot] = Piece{c0, n};
        c0 += n;
    }
    for (int q = 0; q < n_slots && status == ITR_OK; ++q) status = retire((int)((piece + q) % n_slots));
    cleanup();
    if (status != ITR_OK) return status;
    CK(cudaStreamSynchronize(ctx->s_post));
    return ITR_OK;
}

// ---------------------------------------------------------------------------------
// posterior CSV straight from the device result (workflow_posterior.py:697-716)
// ---------------------------------------------------------------------------------
extern "C" int itr_posterior_write_csv(itr_ctx *ctx, const char *path, const int64_t *positions, int n_threads) {
    return itr_posterior_write_csv_ex(ctx, path, positions, nullptr, 1, nullptr, n_threads);
}

extern "C" int itr_posterior_write_csv_ex(itr_ctx *ctx, const char *path, const int64_t *positions,
                                          const int64_t *block_ids, int write_header, int64_t *block_bytes,
                                          int n_threads) {
    if (!ctx) return ITR_ERR_ARG;
    if (!path) return fail(ctx, ITR_ERR_ARG, "itr_posterior_write_csv: path is NULL");
    if (!ctx->have_post) return fail(ctx, ITR_ERR_STATE, "itr_posterior_write_csv: no posterior on the device");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_post));
    const int K = ctx->K;
    const int64_t nb = ctx->n_blocks;
    int64_t max_len = 0;
    for (int64_t i = 0; i < nb; ++i) max_len = std::max(max_len, ctx->h_off[i + 1] - ctx->h_off[i]);
    itr::PosteriorCsv w;
    std::string err;
    if (!w.open(path, K, n_threads, err, write_header != 0)) return fail(ctx, ITR_ERR_IO, "itr_posterior_write_csv: %s", err.c_str());
    double *h[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = ITR_OK;
    auto cleanup = [&]() {
        for (int q = 0; q < 2; ++q) {
            if (h[q]) cudaFreeHost(h[q]);
            if (ev[q]) cudaEventDestroy(ev[q]);
        }
    };
    for (int q = 0; q < 2 && rc == ITR_OK; ++q) {
        if (cudaMallocHost(&h[q], (size_t)max_len * K * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(ctx, ITR_ERR_NOMEM, "itr_posterior_write_csv: cannot page-lock %zu bytes", (size_t)max_len * K * sizeof(double));
        } else if (cudaEventCreateWithFlags(&ev[q], cudaEventDisableTiming) != cudaSuccess) {
            rc = fail(ctx, ITR_ERR_CUDA, "itr_posterior_write_csv: cudaEventCreate failed");
        }
    }
    auto fetch = [&](int64_t i) -> cudaError_t {
        const int64_t c0 = ctx->h_off[i], n = ctx->h_off[i + 1] - c0;
        cudaError_t e = cudaMemcpyAsync(h[i & 1], ctx->d_post + (size_t)c0 * K, (size_t)n * K * sizeof(double),
                                        cudaMemcpyDeviceToHost, ctx->stream2);
        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], ctx->stream2);
        return e;
    };
    if (rc == ITR_OK && nb > 0 && fetch(0) != cudaSuccess) rc = fail(ctx, ITR_ERR_CUDA, "itr_posterior_write_csv: download failed");
    for (int64_t i = 0; i < nb && rc == ITR_OK; ++i) {
        if (cudaEventSynchronize(ev[i & 1]) != cudaSuccess) { rc = fail(ctx, ITR_ERR_CUDA, "itr_posterior_write_csv: download failed"); break; }
        if (i + 1 < nb && fetch(i + 1) != cudaSuccess) { rc = fail(ctx, ITR_ERR_CUDA, "itr_posterior_write_csv: download failed"); break; }
        const int64_t c0 = ctx->h_off[i], n = ctx->h_off[i + 1] - c0;
        const int64_t before = w.bytes_written();
        if (!w.write_block(block_ids ? block_ids[i] : i, positions ? positions + c0 : nullptr, h[i & 1], n, err))
            rc = fail(ctx, ITR_ERR_IO, "itr_posterior_write_csv: %s", err.c_str());
        if (block_bytes) block_bytes[i] = w.bytes_written() - before;
    }
    cudaStreamSynchronize(ctx->stream2);
    cleanup();
    if (!w.close() && rc == ITR_OK) rc = fail(ctx, ITR_ERR_IO, "itr_posterior_write_csv: close failed");
    return rc;
}
