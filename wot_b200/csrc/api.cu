// C ABI of libwot_b200.so (see include/wot_b200.h): context management and the host-buffer entry
// points that bracket the device path with the copies a reference-side caller needs.
#include <stdarg.h>

#include <condition_variable>
#include <mutex>

#include "common.cuh"

namespace wotb {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return WOTB_OK;
    release();
    // grow geometrically: consecutive day-pairs differ in size and must not each pay a cudaMalloc
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&ptr, want);
    }
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return WOTB_ERR_NOMEM;
    }
    cap = want;
    return WOTB_OK;
}

void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

int PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return WOTB_OK;
    release();
    cudaError_t e = cudaHostAlloc(&ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_error("cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return WOTB_ERR_NOMEM;
    }
    cap = bytes;
    return WOTB_OK;
}

void PinnedBuf::release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
}

// cost.cu
int cost_median(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, const double *, double *);
int cost_matrix(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, const double *, double, void *,
                int64_t, int);
int cost_to_f32(wotb_ctx *, const double *, int64_t, int64_t, int64_t, float *, int64_t);
int coupling(wotb_ctx *, const float *, int64_t, int64_t, int64_t, const double *, const double *, double, double,
             void *, int64_t, int, double *, cudaStream_t);
int scale_into(wotb_ctx *, const double *, int64_t, int, const double *, double *);
int median_window_cap(int64_t, int64_t, int64_t *);
int median_window_rows(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, const double *, int64_t, int64_t,
                       unsigned long long *, int64_t, unsigned int *, unsigned long long *);
int median_window_finish(wotb_ctx *, int64_t, int64_t, const unsigned long long *, int64_t, unsigned long long, double *, int *);
int coupling_apply(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *, const double *,
                   double, double, int, const double *, int, double *);
int coupling_sample(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *, const double *,
                    double, double, const double *, const long long *, const double *, int64_t, long long *);
// solver.cu (online_pass.cuh)
struct OnlineSolve;
int online_open(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                const wotb_params *, int, int, double *, double *, OnlineSolve **);
int online_step(OnlineSolve *, int, double *);
int online_state(OnlineSolve *, wotb_info *, int *);
void online_rows(OnlineSolve *, int64_t *, int64_t *);
void online_close(OnlineSolve *);
int64_t online_peer_bytes_of(OnlineSolve *, int);
int online_done_flag(OnlineSolve *, int *);
int online_attach(OnlineSolve *, int, void *const *);
int sinkhorn_online(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                    const wotb_params *, double *, double *, double *, wotb_info *);
int online_rowsums(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                   const double *, int, int, double *, double *);
int bench_mufu(wotb_ctx *, double *);
int coupling_online(wotb_ctx *, const double *, int64_t, const double *, int64_t, int, double, const double *,
                    const double *, double, double, void *, int64_t, int, double *, cudaStream_t);

// Compute slots: with several contexts in flight (wot_b200.pipeline), at most `limit` host-buffer calls are in
// their solve phase at a time; a call gives its slot back before the coupling travels to the host, so a third
// context can push its coupling over PCIe while two others keep the SMs busy.  limit 0 = no limit.
struct ComputeSlots {
    std::mutex m;
    std::condition_variable cv;
    int limit = 0, used = 0;
    void acquire() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return limit <= 0 || used < limit; });
        ++used;
    }
    void release() {
        {
            std::lock_guard<std::mutex> lk(m);
            --used;
        }
        cv.notify_one();
    }
};
static ComputeSlots g_slots;

struct SlotGuard {
    bool held = false;
    SlotGuard() {
        g_slots.acquire();
        held = true;
    }
    void release() {
        if (held) g_slots.release();
        held = false;
    }
    ~SlotGuard() { release(); }
};

static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// Row-chunked coupling materialisation + D2H: chunk k is computed on the main stream while chunk
// k-1 travels on the copy stream.  `produce(r0, rows, dev_out, rowsum)` enqueues the kernel.
template <typename Produce>
static int stream_coupling_to_host(wotb_ctx *ctx, int64_t I, int64_t J, void *tmap_host, int dtype, double *rowsum_dev,
                                   Produce produce) {
    NvtxRange range("wotb:coupling -> host");
    const size_t esz = dtype == WOTB_F32 ? 4 : 8;
    const size_t row_bytes = (size_t)J * esz;
    int64_t rows = (int64_t)((size_t)(96u << 20) / row_bytes);
    if (rows < 1) rows = 1;
    if (rows > I) rows = I;
    const size_t chunk_bytes = (size_t)rows * row_bytes;
    WOTB_TRY(ctx->hOut.reserve(2 * chunk_bytes));
    const bool pinned = is_pinned(tmap_host);
    if (!pinned) WOTB_TRY(ctx->hPin.reserve(2 * chunk_bytes));
    cudaEvent_t made[2], moved[2];
    for (int k = 0; k < 2; ++k) {
        WOTB_CUDA(cudaEventCreateWithFlags(&made[k], cudaEventDisableTiming));
        WOTB_CUDA(cudaEventCreateWithFlags(&moved[k], cudaEventDisableTiming));
    }
    int rc = WOTB_OK;
    int64_t n_chunks = cdiv(I, rows);
    for (int64_t k = 0; k <= n_chunks && rc == WOTB_OK; ++k) {
        const int slot = (int)(k & 1);
        if (k < n_chunks) {
            const int64_t r0 = k * rows, nr = (r0 + rows <= I) ? rows : I - r0;
            char *dev = ctx->hOut.as<char>() + slot * chunk_bytes;
            if (k >= 2) cudaStreamWaitEvent(ctx->stream, moved[slot], 0);
            rc = produce(r0, nr, (void *)dev, rowsum_dev ? rowsum_dev + r0 : nullptr);
            cudaEventRecord(made[slot], ctx->stream);
            cudaStreamWaitEvent(ctx->copy_stream, made[slot], 0);
            char *dst = pinned ? (char *)tmap_host + (size_t)r0 * row_bytes : ctx->hPin.as<char>() + slot * chunk_bytes;
            cudaMemcpyAsync(dst, dev, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream);
            cudaEventRecord(moved[slot], ctx->copy_stream);
        }
        if (!pinned && k >= 1) {  // drain the previous chunk from the bounce buffer into pageable memory
            const int ps = (int)((k - 1) & 1);
            const int64_t r0 = (k - 1) * rows, nr = (r0 + rows <= I) ? rows : I - r0;
            cudaEventSynchronize(moved[ps]);
            memcpy((char *)tmap_host + (size_t)r0 * row_bytes, ctx->hPin.as<char>() + ps * chunk_bytes,
                   (size_t)nr * row_bytes);
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 2; ++k) {
        cudaEventDestroy(made[k]);
        cudaEventDestroy(moved[k]);
    }
    if (rc == WOTB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        set_error("coupling copy failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return WOTB_ERR_CUDA;
    }
    return rc;
}

struct HostVecs {
    double *G, *f, *g, *rowsum;
};

static int stage_vectors(wotb_ctx *ctx, int64_t I, int64_t J, size_t extra_bytes, HostVecs *hv, char **extra) {
    const size_t nI = (size_t)round_up(I, 32) * 8, nJ = (size_t)round_up(J, 32) * 8;
    WOTB_TRY(ctx->hX.reserve(3 * nI + nJ + extra_bytes + 256));
    char *base = ctx->hX.as<char>();
    hv->G = (double *)base;
    hv->f = (double *)(base + nI);
    hv->rowsum = (double *)(base + 2 * nI);
    hv->g = (double *)(base + 3 * nI);
    if (extra) *extra = base + 3 * nI + nJ;
    return WOTB_OK;
}

// Growth loop, optimal_transport.py:10-33 + ot_model.py:319.  `solve(G, f, g, rowsum, info)` runs one
// full cold-start solve; the next growth iteration takes the coupling's row sums as G.
template <typename Solve>
static int growth_loop(wotb_ctx *ctx, int64_t I, const HostVecs &hv, int growth_iters, double *learned_host,
                       wotb_info *infos, Solve solve) {
    WOTB_REQUIRE(growth_iters >= 1, "growth_iters must be >= 1");
    for (int it = 0; it < growth_iters; ++it) {
        char tag[48];
        snprintf(tag, sizeof(tag), "wotb:solve growth_iter %d", it);
        NvtxRange range(tag);
        if (it > 0) WOTB_CUDA(cudaMemcpyAsync(hv.G, hv.rowsum, (size_t)I * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (learned_host)
            WOTB_CUDA(cudaMemcpyAsync(learned_host + (size_t)it * I, hv.G, (size_t)I * 8, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        WOTB_TRY(solve(hv.G, hv.f, hv.g, hv.rowsum, &infos[it]));
    }
    if (learned_host)
        WOTB_CUDA(cudaMemcpyAsync(learned_host + (size_t)growth_iters * I, hv.rowsum, (size_t)I * 8,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    return WOTB_OK;
}

static int finish_host_outputs(wotb_ctx *ctx, int64_t I, int64_t J, const HostVecs &hv, double *f_host, double *g_host) {
    if (f_host) WOTB_CUDA(cudaMemcpyAsync(f_host, hv.f, (size_t)I * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (g_host) WOTB_CUDA(cudaMemcpyAsync(g_host, hv.g, (size_t)J * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}


// ---- coupling applied to populations / sampled without materialising it ------------------
// Stages the description of a finished solve (coordinates, scale, potentials) on the device; x0s / x1s are the
// coordinates multiplied by `scale` (the arithmetic of ot_model.py:245-247).
struct StagedCoupling {
    double *x0s, *x1s, *f, *g;
    char *extra;
};

static int stage_coupling(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int d,
                          const double *scale_host, const double *f_host, const double *g_host, size_t extra_bytes,
                          StagedCoupling *sc) {
    cudaStream_t st = ctx->stream;
    const size_t nx0 = (size_t)round_up(I * d, 32) * 8, nx1 = (size_t)round_up(J * d, 32) * 8, nsc = (size_t)round_up(d, 32) * 8;
    const size_t nI = (size_t)round_up(I, 32) * 8, nJ = (size_t)round_up(J, 32) * 8;
    WOTB_TRY(ctx->hX.reserve(2 * (nx0 + nx1) + nsc + nI + nJ + extra_bytes + 256));
    char *b = ctx->hX.as<char>();
    double *x0 = (double *)b, *x1 = (double *)(b + nx0), *scl = (double *)(b + nx0 + nx1);
    sc->x0s = (double *)(b + nx0 + nx1 + nsc), sc->x1s = (double *)(b + 2 * nx0 + nx1 + nsc);
    char *v = b + 2 * (nx0 + nx1) + nsc;
    sc->f = (double *)v, sc->g = (double *)(v + nI), sc->extra = v + nI + nJ;
    WOTB_CUDA(cudaMemcpyAsync(x0, x0_host, (size_t)I * d * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(x1, x1_host, (size_t)J * d * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(sc->f, f_host, (size_t)I * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(sc->g, g_host, (size_t)J * 8, cudaMemcpyHostToDevice, st));
    if (scale_host) WOTB_CUDA(cudaMemcpyAsync(scl, scale_host, (size_t)d * 8, cudaMemcpyHostToDevice, st));
    WOTB_TRY(scale_into(ctx, x0, I, d, scale_host ? scl : nullptr, sc->x0s));
    WOTB_TRY(scale_into(ctx, x1, J, d, scale_host ? scl : nullptr, sc->x1s));
    return WOTB_OK;
}

static int coupling_apply_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int d,
                               const double *scale_host, double median, const double *f_host, const double *g_host,
                               double eps_final, double out_scale, int forward, const double *p_host, int n_pop,
                               double *out_host) {
    WOTB_REQUIRE(ctx && x0_host && x1_host && f_host && g_host && p_host && out_host, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1 && n_pop >= 1 && median > 0 && eps_final > 0 && out_scale > 0, "bad arguments");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n_in = forward ? I : J, n_out = forward ? J : I;
    const size_t np_in = (size_t)round_up((int64_t)n_pop * n_in, 32) * 8, np_out = (size_t)round_up((int64_t)n_pop * n_out, 32) * 8;
    StagedCoupling sc;
    WOTB_TRY(stage_coupling(ctx, x0_host, I, x1_host, J, d, scale_host, f_host, g_host, np_in + np_out, &sc));
    double *p = (double *)sc.extra, *out = (double *)(sc.extra + np_in);
    WOTB_CUDA(cudaMemcpyAsync(p, p_host, (size_t)n_pop * n_in * 8, cudaMemcpyHostToDevice, st));
    WOTB_TRY(coupling_apply(ctx, sc.x0s, I, sc.x1s, J, d, median, sc.f, sc.g, eps_final, out_scale, forward, p, n_pop, out));
    WOTB_CUDA(cudaMemcpyAsync(out_host, out, (size_t)n_pop * n_out * 8, cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    return WOTB_OK;
}

static int coupling_sample_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int d,
                                const double *scale_host, double median, const double *f_host, const double *g_host,
                                double eps_final, double out_scale, const double *w_host, const int64_t *rows_host,
                                const double *targets_host, int64_t n_samples, int64_t *cols_host) {
    WOTB_REQUIRE(ctx && x0_host && x1_host && f_host && g_host && w_host && rows_host && targets_host && cols_host, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1 && n_samples >= 0 && median > 0 && eps_final > 0 && out_scale > 0, "bad arguments");
    for (int64_t s = 0; s < n_samples; ++s) WOTB_REQUIRE(rows_host[s] >= 0 && rows_host[s] < I, "row index out of range");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nJ = (size_t)round_up(J, 32) * 8, nS = (size_t)round_up(n_samples + 1, 32) * 8;
    StagedCoupling sc;
    WOTB_TRY(stage_coupling(ctx, x0_host, I, x1_host, J, d, scale_host, f_host, g_host, nJ + 3 * nS, &sc));
    double *w = (double *)sc.extra, *targets = (double *)(sc.extra + nJ + nS);
    long long *rows = (long long *)(sc.extra + nJ), *cols = (long long *)(sc.extra + nJ + 2 * nS);
    WOTB_CUDA(cudaMemcpyAsync(w, w_host, (size_t)J * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(rows, rows_host, (size_t)n_samples * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(targets, targets_host, (size_t)n_samples * 8, cudaMemcpyHostToDevice, st));
    WOTB_TRY(coupling_sample(ctx, sc.x0s, I, sc.x1s, J, d, median, sc.f, sc.g, eps_final, out_scale, w, rows, targets, n_samples, cols));
    WOTB_CUDA(cudaMemcpyAsync(cols_host, cols, (size_t)n_samples * 8, cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    return WOTB_OK;
}

}  // namespace wotb

using namespace wotb;

extern "C" {

const char *wotb_version(void) { return "wot_b200 0.1 (sm_100a)"; }
const char *wotb_last_error(void) { return g_error; }

int wotb_create(int device, void *cuda_stream, wotb_ctx **out) {
    WOTB_REQUIRE(out != nullptr, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); wot_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return WOTB_ERR_CUDA;
    }
    WOTB_REQUIRE(device >= 0 && device < n, "device index out of range");
    WOTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    WOTB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libwot_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return WOTB_ERR_CUDA;
    }
    wotb_ctx *ctx = new wotb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        WOTB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    WOTB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    WOTB_CUDA(cudaEventCreate(&ctx->ev0));
    WOTB_CUDA(cudaEventCreate(&ctx->ev1));
    *out = ctx;
    return WOTB_OK;
}

void wotb_set_compute_slots(int32_t n) {
    {
        std::lock_guard<std::mutex> lk(g_slots.m);
        g_slots.limit = n > 0 ? n : 0;
    }
    g_slots.cv.notify_all();
}

void wotb_set_pdl(int32_t on) { wotb::set_pdl(on != 0); }

int wotb_set_sm_limit(wotb_ctx *ctx, int32_t n) {
    WOTB_REQUIRE(ctx != nullptr, "ctx is NULL");
    cudaDeviceProp prop;
    WOTB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->sm_count = (n > 0 && n < prop.multiProcessorCount) ? n : prop.multiProcessorCount;
    return WOTB_OK;
}

void wotb_release_workspace(wotb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    for (DevBuf *b : {&ctx->K, &ctx->vec, &ctx->part, &ctx->ctrl, &ctx->select, &ctx->onl, &ctx->hC, &ctx->hX,
                      &ctx->hOut, &ctx->hTmp})
        b->release();
    ctx->hPin.release();
}

void wotb_destroy(wotb_ctx *ctx) {
    if (!ctx) return;
    wotb_release_workspace(ctx);
    ctx->status.release();
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int wotb_sync(wotb_ctx *ctx) {
    WOTB_REQUIRE(ctx != nullptr, "ctx is NULL");
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}

size_t wotb_workspace_bytes(const wotb_ctx *ctx) {
    if (!ctx) return 0;
    size_t tot = 0;
    for (const DevBuf *b : {&ctx->K, &ctx->vec, &ctx->part, &ctx->ctrl, &ctx->select, &ctx->onl, &ctx->hC, &ctx->hX,
                            &ctx->hOut, &ctx->hTmp})
        tot += b->cap;
    return tot;
}

int wotb_cost_median_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                         const double *scale, double *median_host) {
    return cost_median(ctx, x0, I, x1, J, d, scale, median_host);
}

int wotb_cost_median_window_cap(int64_t I, int64_t J, int64_t *cap) {
    WOTB_REQUIRE(cap != nullptr && I >= 1 && J >= 1, "bad argument");
    return median_window_cap(I, J, cap);
}

int wotb_cost_median_window_rows_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                                     const double *scale, int64_t row_lo, int64_t row_hi, uint64_t *keys, int64_t cap,
                                     uint32_t *count, uint64_t *below) {
    return median_window_rows(ctx, x0, I, x1, J, d, scale, row_lo, row_hi, (unsigned long long *)keys, cap, count,
                              (unsigned long long *)below);
}

int wotb_cost_median_window_finish_dev(wotb_ctx *ctx, int64_t I, int64_t J, const uint64_t *keys, int64_t count,
                                       uint64_t below, double *median_host, int32_t *ok) {
    int k = 0;
    const int rc = median_window_finish(ctx, I, J, (const unsigned long long *)keys, count, below, median_host, &k);
    if (ok) *ok = k;
    return rc;
}

int wotb_cost_matrix_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                         const double *scale, double median, void *C, int64_t ldc, int32_t dtype) {
    WOTB_TRY(cost_matrix(ctx, x0, I, x1, J, d, scale, median, C, ldc, dtype));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}

int wotb_cost_to_f32_dev(wotb_ctx *ctx, const double *src, int64_t ld_src, int64_t I, int64_t J, float *dst,
                         int64_t ld_dst) {
    WOTB_TRY(cost_to_f32(ctx, src, ld_src, I, J, dst, ld_dst));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}

int wotb_sinkhorn_stored_dev(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *G,
                             const wotb_params *params, double *f, double *g, double *rowsum, wotb_info *info) {
    return sinkhorn_stored(ctx, C, ldc, I, J, G, params, f, g, rowsum, info);
}

int wotb_sinkhorn_online_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                             double median, const double *G, const wotb_params *params, double *f, double *g,
                             double *rowsum, wotb_info *info) {
    return sinkhorn_online(ctx, x0, I, x1, J, d, median, G, params, f, g, rowsum, info);
}

int wotb_coupling_dev(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *f,
                      const double *g, double eps, double out_scale, void *out, int64_t ldo, int32_t dtype,
                      double *rowsum) {
    WOTB_TRY(coupling(ctx, C, ldc, I, J, f, g, eps, out_scale, out, ldo, dtype, rowsum, ctx->stream));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}

int wotb_coupling_online_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                             double median, const double *f, const double *g, double eps, double out_scale,
                             void *out, int64_t ldo, int32_t dtype, double *rowsum) {
    WOTB_TRY(coupling_online(ctx, x0, I, x1, J, d, median, f, g, eps, out_scale, out, ldo, dtype, rowsum, ctx->stream));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    return WOTB_OK;
}

int wotb_transport_map_from_cost_host(wotb_ctx *ctx, const double *C_host, int64_t I, int64_t J,
                                      const double *G_host, const wotb_params *params, int32_t growth_iters,
                                      void *tmap_host, int32_t out_dtype, double *learned_growth_host,
                                      double *f_host, double *g_host, wotb_info *infos) {
    WOTB_REQUIRE(ctx && C_host && G_host && params && infos, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1, "empty cost matrix");
    WOTB_REQUIRE(params->kernel == WOTB_KERNEL_STORED, "a caller-supplied cost matrix needs the stored kernel");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    SlotGuard slot;
    const int64_t ld = round_up(J, 32);
    WOTB_TRY(ctx->hC.reserve((size_t)I * ld * 4));
    float *C = ctx->hC.as<float>();
    // upload float64 rows in chunks and round them once to fp32 on the device
    int64_t rows = (int64_t)((size_t)(64u << 20) / ((size_t)J * 8));
    rows = rows < 1 ? 1 : (rows > I ? I : rows);
    WOTB_TRY(ctx->hTmp.reserve((size_t)rows * J * 8));
    for (int64_t r0 = 0; r0 < I; r0 += rows) {
        const int64_t nr = r0 + rows <= I ? rows : I - r0;
        WOTB_CUDA(cudaMemcpyAsync(ctx->hTmp.ptr, C_host + (size_t)r0 * J, (size_t)nr * J * 8, cudaMemcpyHostToDevice,
                                  ctx->stream));
        WOTB_TRY(cost_to_f32(ctx, ctx->hTmp.as<double>(), J, nr, J, C + (size_t)r0 * ld, ld));
    }
    HostVecs hv;
    WOTB_TRY(stage_vectors(ctx, I, J, 0, &hv, nullptr));
    WOTB_CUDA(cudaMemcpyAsync(hv.G, G_host, (size_t)I * 8, cudaMemcpyHostToDevice, ctx->stream));
    WOTB_TRY(growth_loop(ctx, I, hv, growth_iters, learned_growth_host, infos,
                         [&](double *G, double *f, double *g, double *rs, wotb_info *info) {
                             return sinkhorn_stored(ctx, C, ld, I, J, G, params, f, g, rs, info);
                         }));
    const wotb_info &last = infos[growth_iters - 1];
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    slot.release();
    if (tmap_host) {
        WOTB_TRY(stream_coupling_to_host(ctx, I, J, tmap_host, out_dtype, nullptr,
                                         [&](int64_t r0, int64_t nr, void *dev, double *rs) {
                                             return coupling(ctx, C + (size_t)r0 * ld, ld, nr, J, hv.f + r0, hv.g,
                                                             last.eps_final, last.out_scale, dev, J, out_dtype, rs,
                                                             ctx->stream);
                                         }));
    }
    return finish_host_outputs(ctx, I, J, hv, f_host, g_host);
}

int wotb_transport_map_from_coords_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host,
                                        int64_t J, int32_t d, const double *scale_host, const double *G_host,
                                        const wotb_params *params, int32_t growth_iters, void *tmap_host,
                                        int32_t out_dtype, double *learned_growth_host, double *f_host,
                                        double *g_host, double *median_out, wotb_info *infos) {
    WOTB_REQUIRE(ctx && x0_host && x1_host && G_host && params && infos, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1, "empty input");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    SlotGuard slot;
    HostVecs hv;
    char *extra = nullptr;
    const size_t nx0 = (size_t)round_up(I * d, 32) * 8, nx1 = (size_t)round_up(J * d, 32) * 8;
    const size_t nsc = (size_t)round_up(d, 32) * 8;
    WOTB_TRY(stage_vectors(ctx, I, J, 2 * (nx0 + nx1) + nsc, &hv, &extra));
    double *x0 = (double *)extra, *x1 = (double *)(extra + nx0), *sc = (double *)(extra + nx0 + nx1);
    WOTB_CUDA(cudaMemcpyAsync(x0, x0_host, (size_t)I * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(x1, x1_host, (size_t)J * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(hv.G, G_host, (size_t)I * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (scale_host) WOTB_CUDA(cudaMemcpyAsync(sc, scale_host, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream));
    double median = 0.0;
    {
        NvtxRange range("wotb:cost median (exact select)");
        WOTB_TRY(cost_median(ctx, x0, I, x1, J, d, scale_host ? sc : nullptr, &median));
    }
    if (median_out) *median_out = median;
    if (params->kernel == WOTB_KERNEL_STORED) {
        const int64_t ld = round_up(J, 32);
        WOTB_TRY(ctx->hC.reserve((size_t)I * ld * 4));
        float *C = ctx->hC.as<float>();
        WOTB_TRY(cost_matrix(ctx, x0, I, x1, J, d, scale_host ? sc : nullptr, median, C, ld, WOTB_F32));
        WOTB_TRY(growth_loop(ctx, I, hv, growth_iters, learned_growth_host, infos,
                             [&](double *G, double *f, double *g, double *rs, wotb_info *info) {
                                 return sinkhorn_stored(ctx, C, ld, I, J, G, params, f, g, rs, info);
                             }));
        const wotb_info &last = infos[growth_iters - 1];
        WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
        slot.release();
        if (tmap_host) {
            WOTB_TRY(stream_coupling_to_host(ctx, I, J, tmap_host, out_dtype, nullptr,
                                             [&](int64_t r0, int64_t nr, void *dev, double *rs) {
                                                 return coupling(ctx, C + (size_t)r0 * ld, ld, nr, J, hv.f + r0, hv.g,
                                                                 last.eps_final, last.out_scale, dev, J, out_dtype, rs,
                                                                 ctx->stream);
                                             }));
        }
    } else {
        // online kernel: scaled coordinates are kept, C and K never exist
        double *xs0 = (double *)(extra + nx0 + nx1 + nsc), *xs1 = (double *)(extra + 2 * nx0 + nx1 + nsc);
        WOTB_TRY(scale_into(ctx, x0, I, d, scale_host ? sc : nullptr, xs0));
        WOTB_TRY(scale_into(ctx, x1, J, d, scale_host ? sc : nullptr, xs1));
        WOTB_TRY(growth_loop(ctx, I, hv, growth_iters, learned_growth_host, infos,
                             [&](double *G, double *f, double *g, double *rs, wotb_info *info) {
                                 return sinkhorn_online(ctx, xs0, I, xs1, J, d, median, G, params, f, g, rs, info);
                             }));
        const wotb_info &last = infos[growth_iters - 1];
        WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
        slot.release();
        if (tmap_host) {
            WOTB_TRY(stream_coupling_to_host(ctx, I, J, tmap_host, out_dtype, nullptr,
                                             [&](int64_t r0, int64_t nr, void *dev, double *rs) {
                                                 return coupling_online(ctx, xs0 + (size_t)r0 * d, nr, xs1, J, d, median,
                                                                        hv.f + r0, hv.g, last.eps_final, last.out_scale,
                                                                        dev, J, out_dtype, rs, ctx->stream);
                                             }));
        }
    }
    return finish_host_outputs(ctx, I, J, hv, f_host, g_host);
}

int wotb_default_cost_matrix_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J,
                                  int32_t d, const double *scale_host, double *C_host, double *median_out) {
    WOTB_REQUIRE(ctx && x0_host && x1_host && C_host, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1, "empty input");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    const size_t nx0 = (size_t)round_up(I * d, 32) * 8, nx1 = (size_t)round_up(J * d, 32) * 8;
    WOTB_TRY(ctx->hX.reserve(nx0 + nx1 + (size_t)round_up(d, 32) * 8));
    char *base = ctx->hX.as<char>();
    double *x0 = (double *)base, *x1 = (double *)(base + nx0), *sc = (double *)(base + nx0 + nx1);
    WOTB_CUDA(cudaMemcpyAsync(x0, x0_host, (size_t)I * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(x1, x1_host, (size_t)J * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (scale_host) WOTB_CUDA(cudaMemcpyAsync(sc, scale_host, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream));
    double median = 0.0;
    WOTB_TRY(cost_median(ctx, x0, I, x1, J, d, scale_host ? sc : nullptr, &median));
    if (median_out) *median_out = median;
    // row chunks of float64 cost through the coupling streamer (same double-buffered D2H)
    const double *scp = scale_host ? sc : nullptr;
    return stream_coupling_to_host(ctx, I, J, C_host, WOTB_F64, nullptr,
                                   [&](int64_t r0, int64_t nr, void *dev, double *) {
                                       return cost_matrix(ctx, x0 + (size_t)r0 * d, nr, x1, J, d, scp, median, dev, J,
                                                          WOTB_F64);
                                   });
}

int wotb_online_open(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d, double median,
                     const double *G, const wotb_params *params, int32_t shard, int32_t n_shards, double *f, double *g,
                     void **solve) {
    OnlineSolve *S = nullptr;
    WOTB_TRY(online_open(ctx, x0, I, x1, J, d, median, G, params, shard, n_shards, f, g, &S));
    *solve = S;
    return WOTB_OK;
}

int wotb_online_step(void *solve, int32_t op, double *exchange) { return online_step((OnlineSolve *)solve, op, exchange); }

int wotb_online_state(void *solve, wotb_info *info, int32_t *done) {
    int d = 0;
    const int rc = online_state((OnlineSolve *)solve, info, &d);
    if (done) *done = d;
    return rc;
}

int wotb_online_done(void *solve, int32_t *done) {
    int d = 0;
    const int rc = online_done_flag((OnlineSolve *)solve, &d);
    if (done) *done = d;
    return rc;
}

int wotb_online_rows(void *solve, int64_t *row_lo, int64_t *row_hi) {
    WOTB_REQUIRE(solve && row_lo && row_hi, "NULL argument");
    online_rows((OnlineSolve *)solve, row_lo, row_hi);
    return WOTB_OK;
}

void wotb_online_close(void *solve) { online_close((OnlineSolve *)solve); }

// ---- peer-memory exchange of the row-sharded solve -------------------------------------------------------------
int wotb_peer_alloc(wotb_ctx *ctx, int64_t bytes, void **ptr, void *ipc_handle_64) {
    WOTB_REQUIRE(ctx && ptr && bytes > 0, "bad argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    void *p = nullptr;
    WOTB_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess && ipc_handle_64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(ipc_handle_64, &h, sizeof(h));
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        WOTB_CUDA(e);
    }
    *ptr = p;
    return WOTB_OK;
}

int wotb_peer_open(wotb_ctx *ctx, const void *ipc_handle_64, void **ptr) {
    WOTB_REQUIRE(ctx && ipc_handle_64 && ptr, "NULL argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, sizeof(h));
    WOTB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return WOTB_OK;
}

int wotb_peer_close(wotb_ctx *ctx, void *ptr) {
    WOTB_REQUIRE(ctx && ptr, "NULL argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    WOTB_CUDA(cudaIpcCloseMemHandle(ptr));
    return WOTB_OK;
}

int wotb_peer_free(wotb_ctx *ctx, void *ptr) {
    WOTB_REQUIRE(ctx && ptr, "NULL argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    WOTB_CUDA(cudaFree(ptr));
    return WOTB_OK;
}

int wotb_online_peer_bytes(void *solve, int32_t world, int64_t *bytes) {
    WOTB_REQUIRE(solve && bytes && world >= 1, "bad argument");
    *bytes = online_peer_bytes_of((OnlineSolve *)solve, world);
    return WOTB_OK;
}

int wotb_online_attach_peers(void *solve, int32_t world, void *const *bufs) {
    return online_attach((OnlineSolve *)solve, world, bufs);
}

int wotb_bench_matvec_dev(wotb_ctx *ctx, int64_t I, int64_t J, int32_t reps, double *ms_row, double *ms_col,
                          double *ms_fused) {
    return bench_matvec(ctx, I, J, reps, ms_row, ms_col, ms_fused);
}

int wotb_online_rowsums_dev(wotb_ctx *ctx, const double *x_out, int64_t n_out, const double *x_in, int64_t n_in, int32_t d,
                            double scale, const double *off_out, const double *off_in, int32_t impl, int32_t reps,
                            double *sums, double *ms_per_pass) {
    return online_rowsums(ctx, x_out, n_out, x_in, n_in, d, scale, off_out, off_in, impl, reps, sums, ms_per_pass);
}

int wotb_bench_mufu_dev(wotb_ctx *ctx, double *ex2_per_s) { return bench_mufu(ctx, ex2_per_s); }

int wotb_coupling_apply_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int32_t d,
                             const double *scale_host, double median, const double *f_host, const double *g_host,
                             double eps_final, double out_scale, int32_t forward, const double *p_host, int32_t n_pop,
                             double *out_host) {
    return coupling_apply_host(ctx, x0_host, I, x1_host, J, d, scale_host, median, f_host, g_host, eps_final, out_scale,
                               forward, p_host, n_pop, out_host);
}

int wotb_coupling_sample_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int32_t d,
                              const double *scale_host, double median, const double *f_host, const double *g_host,
                              double eps_final, double out_scale, const double *w_host, const int64_t *rows_host,
                              const double *targets_host, int64_t n_samples, int64_t *cols_host) {
    return coupling_sample_host(ctx, x0_host, I, x1_host, J, d, scale_host, median, f_host, g_host, eps_final, out_scale,
                                w_host, rows_host, targets_host, n_samples, cols_host);
}

int wotb_pinned_alloc(size_t bytes, void **out) {
    WOTB_REQUIRE(out != nullptr, "out is NULL");
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error("cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return WOTB_ERR_NOMEM;
    }
    return WOTB_OK;
}

void wotb_pinned_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
