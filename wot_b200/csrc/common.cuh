// Internal declarations shared by the translation units of libwot_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/wot_b200.h"

namespace wotb {

void set_error(const char *fmt, ...);

#define WOTB_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            ::wotb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err__)); \
            return err__ == cudaErrorMemoryAllocation ? WOTB_ERR_NOMEM : WOTB_ERR_CUDA;            \
        }                                                                                          \
    } while (0)

#define WOTB_TRY(expr)                  \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != WOTB_OK) return rc__; \
    } while (0)

#define WOTB_REQUIRE(cond, msg)                             \
    do {                                                    \
        if (!(cond)) {                                      \
            ::wotb::set_error("invalid argument: %s", msg); \
            return WOTB_ERR_INVALID;                        \
        }                                                   \
    } while (0)

// A grow-only device buffer: solving 39 day-pairs back to back must not pay cudaMalloc per pair.
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T *as() const {
        return static_cast<T *>(ptr);
    }
};

struct PinnedBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T *as() const {
        return static_cast<T *>(ptr);
    }
};

// NVTX range for the stages of a transport map (median, cost, solve k, coupling): visible in Nsight Systems / ncu
// --nvtx, free when no tool is attached (the reference has logger.info lines only).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t cdiv(int64_t x, int64_t m) { return (x + m - 1) / m; }

}  // namespace wotb

struct SolveCtrl;

struct wotb_ctx {
    int device = 0;
    int sm_count = wotb::kSMs;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // D2H of coupling chunks overlaps the next chunk's kernel
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // solver workspaces
    wotb::DevBuf K;        // stored Gibbs kernel, fp32 [I, ld]
    wotb::DevBuf vec;      // all O(I+J) vectors of a solve, carved by offset
    wotb::DevBuf part;     // column-sum partials
    wotb::DevBuf ctrl;     // SolveCtrl + counters
    wotb::DevBuf select;   // median radix-select histograms/state
    wotb::DevBuf onl;      // online-kernel staging of coordinates (fp32 / split layouts)
    wotb::PinnedBuf status;  // ring of SolveCtrl snapshots the host polls
    // host-API staging
    wotb::DevBuf hC;       // fp32 cost [I, ld]
    wotb::DevBuf hX;       // coordinates / G / scale / f / g / growth rows
    wotb::DevBuf hOut;     // coupling chunks (2 buffers)
    wotb::DevBuf hTmp;     // fp64 chunk staging for cost upload
    wotb::PinnedBuf hPin;  // pinned bounce buffers for pageable host memory
};

namespace wotb {
// solver.cu
int sinkhorn_stored(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *G,
                    const wotb_params *prm, double *f, double *g, double *rowsum, wotb_info *info);
int bench_matvec(wotb_ctx *ctx, int64_t I, int64_t J, int reps, double *ms_row, double *ms_col, double *ms_fused);
void set_pdl(bool on);
}  // namespace wotb
