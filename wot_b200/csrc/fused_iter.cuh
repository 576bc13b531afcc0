// Fused Sinkhorn iteration for the stored kernel: K is read from HBM ONCE per iteration.
//
// The reference does two matvecs per iteration (optimal_transport.py:133-134): K.(b dy) by rows, then
// K^T.(a dx) by columns -- two sweeps over K.  The column sweep only needs a_i of the rows it has
// already seen, and a_i only needs the sum of row i.  So one persistent CTA per SM streams its share of
// rows through shared memory with TMA bulk copies (cp.async.bulk + mbarrier ring): while a row is
// resident the CTA (1) reduces it against w to get s_i, (2) a dedicated warp turns s_i into a_i in
// float64, (3) the same row, still in shared memory, is accumulated into per-thread column partials
// with weight a_i/I.  After a grid-wide barrier (cooperative launch, all CTAs co-resident) every CTA
// reduces the partials of its slice of columns in a fixed order, applies the b update, and the last
// one closes the iteration; a second barrier publishes the new state and the CTAs start the next
// iteration of the batch without returning to the host: one launch per batch of iterations.
//
// Warp roles (1024 threads): warps 0..29 compute (each thread owns CPT float4 column groups for the
// whole kernel: its slice of w and its column accumulators live in registers), warp 30 does the
// float64 scalar math per row, warp 31 lane 0 is the TMA producer.
#pragma once

#include "solver_state.cuh"

namespace wotb {

constexpr int kFuseThreads = 1024;
constexpr int kFuseComputeWarps = 30;
constexpr int kFuseCompute = kFuseComputeWarps * 32;
constexpr int kFuseMaxStages = 8;
constexpr int kFuseTailBytes = 4096;             // barriers + reduction slots after the row stages
constexpr int kFuseSmemMax = 232448;             // 227 KB opt-in limit per CTA on sm_100
constexpr int kFuseMaxCpt = 6;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Grid-wide barrier between the co-resident CTAs of the cooperative launch.  `count` is zeroed by
// k_build at the head of every launch sequence; `target` is this CTA's running arrival target.
__device__ __forceinline__ void grid_barrier(unsigned int *count, unsigned int &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(count, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(count) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

template <int CPT>
__global__ void __launch_bounds__(kFuseThreads, 1)
    k_fused(const float *__restrict__ K, long long ld, SolveVecs V, SolveCtrl *ctrl, float *__restrict__ part,
            int n_stages, int lag, int max_iters, int area_bytes) {
    // Every CTA sees the same control state here (stream order) and after each closing grid barrier,
    // so all of them leave together.
    if (!iteration_active(ctrl)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n4 = (int)(ld >> 2);
    const uint32_t row_bytes = (uint32_t)(ld * 4);
    unsigned char *tail = smem_raw + area_bytes;  // >= n_stages rows and >= the 8 KB phase-B scratch
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty = full + kFuseMaxStages;
    uint64_t *sready = empty + kFuseMaxStages;
    uint64_t *zready = sready + kFuseMaxStages;
    double *red = reinterpret_cast<double *>(zready + kFuseMaxStages);  // [stage][32]
    float *zs = reinterpret_cast<float *>(red + kFuseMaxStages * 32);    // [stage]
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFuseComputeWarps);
            mbar_init(&sready[s], kFuseComputeWarps);
            mbar_init(&zready[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (blockIdx.x == 0) ctrl->need_build = 0;  // K is current from here on
    }
    __syncthreads();
    const int I = ctrl->I, J = ctrl->J;
    const int G = gridDim.x;
    const int r0 = (int)((long long)I * blockIdx.x / G);
    const int r1 = (int)((long long)I * (blockIdx.x + 1) / G);
    const int nr = r1 - r0;
    // column slice this CTA finishes in phase B
    const int cw = (J + G - 1) / G;
    const int j0 = blockIdx.x * cw;
    const int j1 = min(J, j0 + cw);
    const int cw_pad = (cw + 31) & ~31;
    const int n_slices = min(kFuseThreads / cw_pad, 32);
    unsigned int bar_target = 0;
    unsigned int kbase = 0;  // rows this CTA has pushed through the ring in earlier iterations
    volatile SolveCtrl *vc = ctrl;

    for (int it = 0; it < max_iters; ++it) {
        if (it > 0 && (vc->done || vc->stop != 0 || vc->batch_done >= vc->batch_iters)) break;
        const int cur = vc->cur;
        const bool first_of_batch = vc->batch_done == 0;
        // =========================== phase A: one sweep over this CTA's rows ====================
        if (wid == kFuseComputeWarps + 1) {
            // ---------------- TMA producer ------------------------------------------------------
            if (lane == 0) {
                for (int k = 0; k < nr; ++k) {
                    const unsigned int kk = kbase + k;
                    const int s = kk % n_stages;
                    if (kk >= (unsigned)n_stages) mbar_wait(&empty[s], ((kk / n_stages) + 1) & 1);
                    mbar_expect_tx(&full[s], row_bytes);
                    bulk_g2s(smem_raw + (size_t)s * row_bytes, K + (long long)(r0 + k) * ld, row_bytes, &full[s]);
                }
            }
        } else if (wid == kFuseComputeWarps) {
            // ---------------- float64 row math: a_i = (p_i / s_i)^alpha1 exp(-u_i/(lambda1+eps)) -
            const double alpha1 = ctrl->alpha1;
            const double dx = 1.0 / (double)I;
            double *a_out = V.a[cur ^ 1];
            double amax = 0.0;
            double lp_next = nr > 0 ? V.lp[r0] : 0.0, lu_next = nr > 0 ? V.lu[r0] : 0.0;
            for (int k = 0; k < nr; ++k) {
                const unsigned int kk = kbase + k;
                const int s = kk % n_stages, row = r0 + k;
                const double lp = lp_next, lu = lu_next;
                if (k + 1 < nr) {  // next row's constants travel while this row is reduced
                    lp_next = V.lp[row + 1];
                    lu_next = V.lu[row + 1];
                }
                mbar_wait(&sready[s], (kk / n_stages) & 1);
                double part_sum = lane < kFuseComputeWarps ? red[s * 32 + lane] : 0.0;
                part_sum = warp_sum(part_sum);
                if (lane == 0) {
                    const double a = scaling_update(lp, part_sum, alpha1, lu);
                    a_out[row] = a;
                    if (first_of_batch) V.sfirst[row] = part_sum;
                    zs[s] = (float)(a * dx);
                    amax = fmax(amax, fabs(a));
                    mbar_arrive(&zready[s]);
                }
            }
            if (lane == 0) atomic_max_nonneg(&ctrl->maxabs, amax);
        } else {
            // ---------------- compute warps -----------------------------------------------------
            float4 w_reg[CPT], acc[CPT];
            const float4 *w4 = reinterpret_cast<const float4 *>(V.w);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int q = tid + c * kFuseCompute;
                w_reg[c] = q < n4 ? __ldcg(w4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);  // written by peers: L2
                acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int k = 0; k < nr + lag; ++k) {
                if (k < nr) {
                    const unsigned int kk = kbase + k;
                    const int s = kk % n_stages;
                    mbar_wait(&full[s], (kk / n_stages) & 1);
                    const float4 *rowp = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * row_bytes);
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int q = tid + c * kFuseCompute;
                        if (q < n4) {
                            const float4 v = rowp[q];
                            s0 = fmaf(v.x, w_reg[c].x, s0);
                            s1 = fmaf(v.y, w_reg[c].y, s1);
                            s2 = fmaf(v.z, w_reg[c].z, s2);
                            s3 = fmaf(v.w, w_reg[c].w, s3);
                        }
                    }
                    double ps = (double)((s0 + s1) + (s2 + s3));
                    ps = warp_sum(ps);
                    if (lane == 0) {
                        red[s * 32 + wid] = ps;
                        mbar_arrive(&sready[s]);
                    }
                }
                if (k >= lag) {
                    const unsigned int kk = kbase + (k - lag);
                    const int s = kk % n_stages;
                    mbar_wait(&zready[s], (kk / n_stages) & 1);
                    const float z = zs[s];
                    const float4 *rowp = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * row_bytes);
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int q = tid + c * kFuseCompute;
                        if (q < n4) {
                            const float4 v = rowp[q];
                            acc[c].x = fmaf(v.x, z, acc[c].x);
                            acc[c].y = fmaf(v.y, z, acc[c].y);
                            acc[c].z = fmaf(v.z, z, acc[c].z);
                            acc[c].w = fmaf(v.w, z, acc[c].w);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
            }
            float4 *dst = reinterpret_cast<float4 *>(part + (long long)blockIdx.x * ld);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int q = tid + c * kFuseCompute;
                if (q < n4) dst[q] = acc[c];
            }
        }
        kbase += nr;
        grid_barrier(&ctrl->grid_bar, bar_target);  // every CTA's column partials are in L2

        // =========================== phase B: b update for this CTA's column slice ==============
        // t_j = sum over CTAs of the partials in a fixed order (slices interleave the CTAs; the slice
        // sums are added in slice order), then b_j = (q / t_j)^alpha2 exp(-v_j/(lambda2+eps)), :134.
        double *scratch = reinterpret_cast<double *>(smem_raw);  // the row stages are idle now
        {
            const int col = tid % cw_pad, sl = tid / cw_pad;
            if (sl < n_slices && j0 + col < j1) {
                double t = 0.0;
                const float *src = part + j0 + col;
#pragma unroll 4
                for (int c = sl; c < G; c += n_slices) t += (double)__ldcg(src + (long long)c * ld);
                scratch[sl * cw_pad + col] = t;
            }
        }
        __syncthreads();
        double bmax = 0.0;
        if (tid < cw && j0 + tid < j1) {
            const int j = j0 + tid;
            double t = 0.0;
            for (int sl = 0; sl < n_slices; ++sl) t += scratch[sl * cw_pad + tid];
            const double b = scaling_update(ctrl->lq, t, ctrl->alpha2, V.lv[j]);
            V.b[cur ^ 1][j] = b;
            V.t[j] = t;
            V.w[j] = (float)(b * (1.0 / (double)J));
            bmax = fabs(b);
        }
        if (tid < cw_pad) {
            bmax = warp_max(bmax);
            if (lane == 0) atomic_max_nonneg(&ctrl->maxabs, bmax);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int ticket = atomicAdd(&ctrl->col_tiles_done, 1u);
            if (ticket == (unsigned)G - 1) {
                __threadfence();
                ctrl->col_tiles_done = 0;
                close_iteration(ctrl);
            }
        }
        grid_barrier(&ctrl->grid_bar, bar_target);  // the closed iteration's state is visible to all
    }
}

struct FusePlan {
    bool ok = false;
    int cpt = 0, stages = 0, lag = 0, grid = 0, area = 0;
    size_t smem = 0;
};

inline FusePlan plan_fused(const wotb_ctx *ctx, int64_t I, int64_t ld) {
    FusePlan p;
    const int64_t n4 = ld / 4;
    p.cpt = (int)cdiv(n4, kFuseCompute);
    const size_t row_bytes = (size_t)ld * 4;
    int stages = (int)((kFuseSmemMax - kFuseTailBytes) / row_bytes);
    if (stages > kFuseMaxStages) stages = kFuseMaxStages;
    p.stages = stages;
    p.lag = stages >= 3 ? 1 : 0;
    p.grid = (int)(I < ctx->sm_count ? I : ctx->sm_count);
    p.area = (int)((size_t)stages * row_bytes > 8192 ? (size_t)stages * row_bytes : 8192);
    p.smem = (size_t)p.area + kFuseTailBytes;
    p.ok = p.cpt >= 1 && p.cpt <= kFuseMaxCpt && stages >= 2;
    return p;
}

template <int CPT>
int launch_fused_t(const FusePlan &p, cudaStream_t st, const float *K, long long ld, SolveVecs V, SolveCtrl *ctrl,
                   float *part, int max_iters) {
    static bool configured = false;
    if (!configured) {
        WOTB_CUDA(cudaFuncSetAttribute(k_fused<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmemMax));
        configured = true;
    }
    int n_stages = p.stages, lag = p.lag, iters = max_iters, area = p.area;
    void *args[] = {(void *)&K, (void *)&ld, (void *)&V, (void *)&ctrl, (void *)&part, (void *)&n_stages, (void *)&lag,
                    (void *)&iters, (void *)&area};
    WOTB_CUDA(cudaLaunchCooperativeKernel((const void *)k_fused<CPT>, dim3(p.grid), dim3(kFuseThreads), args, p.smem, st));
    return WOTB_OK;
}

inline int launch_fused(const FusePlan &p, cudaStream_t st, const float *K, long long ld, const SolveVecs &V,
                        SolveCtrl *ctrl, float *part, int max_iters) {
    switch (p.cpt) {
        case 1: return launch_fused_t<1>(p, st, K, ld, V, ctrl, part, max_iters);
        case 2: return launch_fused_t<2>(p, st, K, ld, V, ctrl, part, max_iters);
        case 3: return launch_fused_t<3>(p, st, K, ld, V, ctrl, part, max_iters);
        case 4: return launch_fused_t<4>(p, st, K, ld, V, ctrl, part, max_iters);
        case 5: return launch_fused_t<5>(p, st, K, ld, V, ctrl, part, max_iters);
        case 6: return launch_fused_t<6>(p, st, K, ld, V, ctrl, part, max_iters);
    }
    set_error("fused iteration: unsupported column count");
    return WOTB_ERR_INVALID;
}

}  // namespace wotb
