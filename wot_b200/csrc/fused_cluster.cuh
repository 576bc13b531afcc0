// Fused Sinkhorn iteration for WIDE stored kernels (J > 23k): a thread-block cluster shares every row.
//
// k_fused (fused_iter.cuh) reads K once per iteration because a whole row is resident in one CTA's shared memory
// between its reduction (s_i) and its accumulation into the column sums (weight a_i) -- which caps a row at what
// one CTA can hold in registers and shared memory (23,040 columns).  Wider matrices fell back to two sweeps over K
// (k_row + k_col): 50k x 50k ran at 0.76 of the HBM peak all-in (profiles/r1v, r2f).  Here a cluster of CL = 2, 4
// or 8 CTAs owns a set of rows and every CTA of the cluster a contiguous segment of the columns:
//   * each CTA streams ITS segment of the cluster's rows through its own TMA ring and reduces it against its slice
//     of w -> a partial row sum;
//   * the partials meet through distributed shared memory: warp 30 of every CTA stores its partial into the slot
//     [stage][own rank] of EVERY CTA of the cluster (st.shared::cluster) and arrives on that CTA's `xready`
//     mbarrier (mbarrier.arrive.release.cluster); when CL arrivals are in, every CTA adds the CL partials in rank
//     order -- the same bits everywhere -- and computes a_i in float64 redundantly;
//   * the segment, still in shared memory, is accumulated into the CTA's column partials with weight a_i / I.
// Phase B (column partial reduction over the clusters in a fixed order, b update, closing the iteration) and the two
// grid barriers per iteration are those of k_fused: one cooperative launch per batch of iterations, K read ONCE.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include "fused_iter.cuh"

namespace wotb {

constexpr int kFuseMaxCluster = 8;
// warp roles: 0..28 compute, 29 receiver (float64 row math), 30 sender (partial row sums into the cluster), 31 TMA producer
constexpr int kFclComputeWarps = 29;
constexpr int kFclCompute = kFclComputeWarps * 32;
// The exchange has its own ring of slots, deeper than the row stages: a peer may send row m while this CTA's receiver
// is still at row m - 2 lag - 2 (the sender does not wait for the receiver), so slots are reused every kFclSlots rows
// with kFclSlots >= 2 lag + 2.
constexpr int kFclSlots = 16;
constexpr int kFclBatch = 4;   // rows the receiver finishes side by side

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// bounded waits: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_trap(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void grid_barrier_trap(unsigned int *count, unsigned int &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(count, 1u);
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int *>(count) < target) {
            if (++spins > (1u << 28)) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

template <int CPT>
__global__ void __launch_bounds__(kFuseThreads, 1)
    k_fused_cl(const float *__restrict__ K, long long ld, SolveVecs V, SolveCtrl *ctrl, float *__restrict__ part,
               int n_stages, int lag, int max_iters, int area_bytes, int seg4) {
    // every CTA sees the same control state here and after each closing grid barrier: all of them leave together
    if (!iteration_active(ctrl)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int CL = (int)cluster_nctarank(), cr = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL, NC = gridDim.x / CL;
    const int n4 = (int)(ld >> 2);
    const int q0 = cr * seg4;                       // first float4 column group of this CTA's segment
    const int my4 = max(0, min(n4, q0 + seg4) - q0);
    const uint32_t seg_stride = (uint32_t)seg4 * 16u, seg_bytes = (uint32_t)my4 * 16u;
    unsigned char *tail = smem_raw + area_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty = full + kFuseMaxStages;
    uint64_t *sready = empty + kFuseMaxStages;
    uint64_t *zready = sready + kFuseMaxStages;
    uint64_t *xready = zready + kFuseMaxStages;
    double *red = reinterpret_cast<double *>(xready + kFclSlots);        // [stage][32]
    double *xsum = red + kFuseMaxStages * 32;                            // [slot][kFuseMaxCluster]
    float *zs = reinterpret_cast<float *>(xsum + kFclSlots * kFuseMaxCluster);  // [stage]
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFclComputeWarps);
            mbar_init(&sready[s], kFclComputeWarps);
            mbar_init(&zready[s], 1);
        }
        for (int x = 0; x < kFclSlots; ++x) mbar_init(&xready[x], (uint32_t)CL);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (blockIdx.x == 0) ctrl->need_build = 0;  // K is current from here on
    }
    __syncthreads();
    cluster_sync_all();  // every CTA's barriers exist before a peer arrives on them
    const int I = ctrl->I, J = ctrl->J;
    const int G = gridDim.x;
    const int r0 = (int)((long long)I * cid / NC);
    const int r1 = (int)((long long)I * (cid + 1) / NC);
    const int nr = r1 - r0;
    // column slice this CTA finishes in phase B
    const int cw = (J + G - 1) / G;
    const int j0 = blockIdx.x * cw;
    const int j1 = min(J, j0 + cw);
    const int cw_pad = (cw + 31) & ~31;
    const int n_slices = min(kFuseThreads / cw_pad, 32);
    unsigned int bar_target = 0;
    unsigned int kbase = 0;
    volatile SolveCtrl *vc = ctrl;

    for (int it = 0; it < max_iters; ++it) {
        if (it > 0 && (vc->done || vc->stop != 0 || vc->batch_done >= vc->batch_iters)) break;
        const int cur = vc->cur;
        const bool first_of_batch = vc->batch_done == 0;
        // =========================== phase A: one sweep over this cluster's rows =================
        if (wid == kFclComputeWarps + 2) {
            // ---------------- TMA producer: this CTA's column segment of every row ---------------
            if (lane == 0 && my4 > 0) {
                for (int k = 0; k < nr; ++k) {
                    const unsigned int kk = kbase + k;
                    const int s = kk % n_stages;
                    if (kk >= (unsigned)n_stages) mbar_wait_trap(&empty[s], ((kk / n_stages) + 1) & 1);
                    mbar_expect_tx(&full[s], seg_bytes);
                    bulk_g2s(smem_raw + (size_t)s * seg_stride, K + (long long)(r0 + k) * ld + (long long)q0 * 4, seg_bytes,
                             &full[s]);
                }
            }
        } else if (wid == kFclComputeWarps + 1) {
            // ---------------- sender: this CTA's partial row sum into every CTA of the cluster ----
            // (its own warp, so that the exchange of row k + 1 does not wait for the float64 math of row k)
            for (int k = 0; k < nr; ++k) {
                const unsigned int kk = kbase + k;
                const int s = kk % n_stages;
                const uint32_t par = (kk / n_stages) & 1;
                mbar_wait_trap(&sready[s], par);
                double part_sum = lane < kFclComputeWarps ? red[s * 32 + lane] : 0.0;
                part_sum = warp_sum(part_sum);
                if (lane < CL) {
                    // lane l hands the partial to CTA l of the cluster; the own CTA's copy goes through plain shared-memory
                    // forms (same barrier, same release), which is also what compute-sanitizer's racecheck can follow
                    const int x = (int)(kk % kFclSlots);
                    if (lane == cr) {
                        xsum[x * kFuseMaxCluster + cr] = part_sum;
                        asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&xready[x]))
                                     : "memory");
                    } else {
                        const uint32_t slot = mapa_shared(smem_u32(&xsum[x * kFuseMaxCluster + cr]), (uint32_t)lane);
                        const uint32_t bar = mapa_shared(smem_u32(&xready[x]), (uint32_t)lane);
                        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(slot), "d"(part_sum) : "memory");
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
                    }
                }
                __syncwarp();
            }
        } else if (wid == kFclComputeWarps) {
            // ---------------- receiver: float64 row math once the partial row sums have met -------
            // kFclBatch rows at a time, one lane each: the log / exp chain of a row (~1 us) is longer than a row's
            // share of the HBM stream when a cluster is 4 or 8 CTAs wide, so rows are finished side by side
            const double alpha1 = ctrl->alpha1;
            const double dx = 1.0 / (double)I;
            double *a_out = V.a[cur ^ 1];
            double amax = 0.0;
            const int batch = min(kFclBatch, lag);  // a row's a_i may wait for batch - 1 later rows: covered by the lag
            for (int k0 = 0; k0 < nr; k0 += batch) {
                const int k = k0 + lane;
                if (lane < batch && k < nr) {
                    const unsigned int kk = kbase + k;
                    const int s = kk % n_stages, row = r0 + k;
                    const double lp = V.lp[row], lu = V.lu[row];
                    const int x = (int)(kk % kFclSlots);
                    mbar_wait_cluster(&xready[x], (kk / kFclSlots) & 1);
                    double total = 0.0;
                    for (int c = 0; c < CL; ++c) total += xsum[x * kFuseMaxCluster + c];  // rank order: same bits in every CTA
                    const double a = scaling_update(lp, total, alpha1, lu);
                    if (cr == 0) {
                        a_out[row] = a;
                        if (first_of_batch) V.sfirst[row] = total;
                        amax = fmax(amax, fabs(a));
                    }
                    zs[s] = (float)(a * dx);
                    mbar_arrive(&zready[s]);
                }
                __syncwarp();
            }
            amax = warp_max(amax);
            if (lane == 0 && cr == 0) atomic_max_nonneg(&ctrl->maxabs, amax);
        } else {
            // ---------------- compute warps --------------------------------------------------------
            float4 w_reg[CPT], acc[CPT];
            const float4 *w4 = reinterpret_cast<const float4 *>(V.w) + q0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int q = tid + c * kFclCompute;
                w_reg[c] = q < my4 ? __ldcg(w4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int k = 0; k < nr + lag; ++k) {
                if (k < nr) {
                    const unsigned int kk = kbase + k;
                    const int s = kk % n_stages;
                    if (my4 > 0) mbar_wait_trap(&full[s], (kk / n_stages) & 1);
                    const float4 *rowp = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * seg_stride);
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int q = tid + c * kFclCompute;
                        if (q < my4) {
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
                    mbar_wait_trap(&zready[s], (kk / n_stages) & 1);
                    const float z = zs[s];
                    const float4 *rowp = reinterpret_cast<const float4 *>(smem_raw + (size_t)s * seg_stride);
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int q = tid + c * kFclCompute;
                        if (q < my4) {
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
            float4 *dst = reinterpret_cast<float4 *>(part + (long long)cid * ld) + q0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int q = tid + c * kFclCompute;
                if (q < my4) dst[q] = acc[c];
            }
        }
        kbase += nr;
        grid_barrier_trap(&ctrl->grid_bar, bar_target);  // every cluster's column partials are in L2

        // =========================== phase B: b update for this CTA's column slice ==============
        double *scratch = reinterpret_cast<double *>(smem_raw);  // the row stages are idle now
        {
            const int col = tid % cw_pad, sl = tid / cw_pad;
            if (sl < n_slices && j0 + col < j1) {
                double t = 0.0;
                const float *src = part + j0 + col;
#pragma unroll 4
                for (int c = sl; c < NC; c += n_slices) t += (double)__ldcg(src + (long long)c * ld);
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
        grid_barrier_trap(&ctrl->grid_bar, bar_target);  // the closed iteration's state is visible to all
    }
    cluster_sync_all();  // nobody leaves while a peer may still address its shared memory
}

struct FuseClusterPlan {
    bool ok = false;
    int cpt = 0, stages = 0, lag = 0, cl = 0, n_clusters = 0, area = 0, seg4 = 0;
    size_t smem = 0;
};

template <int CPT>
inline cudaError_t fused_cl_config(cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attrs, int cl, int grid, size_t smem,
                                   cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_fused_cl<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmemMax);
    if (e != cudaSuccess) return e;
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(grid), cfg->blockDim = dim3(kFuseThreads), cfg->dynamicSmemBytes = smem, cfg->stream = st;
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = cl, attrs[0].val.clusterDim.y = 1, attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg->attrs = attrs, cfg->numAttrs = 2;
    return cudaSuccess;
}

template <int CPT>
inline int fused_cl_max_clusters(int cl, size_t smem) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    if (fused_cl_config<CPT>(&cfg, attrs, cl, cl, smem, nullptr) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cfg.numAttrs = 1;  // the occupancy query takes the cluster shape only
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_fused_cl<CPT>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// smallest cluster whose column segments fit one CTA's registers (CPT <= 6) and leave >= 3 ring stages
inline FuseClusterPlan plan_fused_cluster(const wotb_ctx *ctx, int64_t I, int64_t ld) {
    FuseClusterPlan p;
    const int64_t n4 = ld / 4;
    int cl_min = 2;
    if (const char *e = getenv("WOTB_FUSE_CLUSTER")) {  // measurement knob: smallest cluster size that is tried (0: none)
        cl_min = atoi(e);
        if (cl_min <= 0) return p;
    }
    for (int cl = 2; cl <= kFuseMaxCluster; cl *= 2) {
        if (cl < cl_min) continue;
        const int64_t seg4 = cdiv(n4, cl);
        const int cpt = (int)cdiv(seg4, kFclCompute);
        const size_t seg_bytes = (size_t)seg4 * 16;
        int stages = (int)((kFuseSmemMax - kFuseTailBytes) / seg_bytes);
        if (stages > kFuseMaxStages) stages = kFuseMaxStages;
        if (cpt < 1 || cpt > kFuseMaxCpt || stages < 3) continue;
        // The SMALLEST cluster that fits wins.  Measured on B200 (profiles/r2y_fused_cluster.txt): wider clusters buy a
        // deeper ring (8 stages, lag 6) but halve the segment, and the per-row handshakes of the 29 compute warps then
        // dominate: 50k x 50k 1.96 ms per iteration with 4 CTAs x 4 stages against 3.42 ms with 8 CTAs x 8 stages.
        p.cl = cl, p.cpt = cpt, p.stages = stages, p.seg4 = (int)seg4;
        // the cluster exchange adds a round trip between a row's reduction and its a_i: as many rows of slack as the ring
        // allows (stages >= lag + 2), within what the exchange slots allow (kFclSlots >= 2 lag + 2)
        p.lag = stages - 2 < (kFclSlots - 2) / 2 ? stages - 2 : (kFclSlots - 2) / 2;
        if (const char *e = getenv("WOTB_FUSE_LAG")) {  // measurement knob
            const int l = atoi(e);
            if (l >= 1 && l <= stages - 2 && 2 * l + 2 <= kFclSlots) p.lag = l;
        }
        p.area = (int)((size_t)stages * seg_bytes > 8192 ? (size_t)stages * seg_bytes : 8192);
        p.smem = (size_t)p.area + kFuseTailBytes;
        int nc = 0;
        switch (cpt) {
            case 1: nc = fused_cl_max_clusters<1>(cl, p.smem); break;
            case 2: nc = fused_cl_max_clusters<2>(cl, p.smem); break;
            case 3: nc = fused_cl_max_clusters<3>(cl, p.smem); break;
            case 4: nc = fused_cl_max_clusters<4>(cl, p.smem); break;
            case 5: nc = fused_cl_max_clusters<5>(cl, p.smem); break;
            case 6: nc = fused_cl_max_clusters<6>(cl, p.smem); break;
        }
        if (nc > ctx->sm_count / cl) nc = ctx->sm_count / cl;
        if (nc > I) nc = (int)I;
        if (nc < 1) continue;
        p.n_clusters = nc;
        p.ok = true;
        if (getenv("WOTB_FUSE_DEBUG"))
            fprintf(stderr, "[fused cluster] ld %lld: cluster %d x %d clusters, %d float4 per thread, %d stages, lag %d\n",
                    (long long)ld, cl, nc, cpt, stages, p.lag);
        return p;
    }
    return p;
}

template <int CPT>
int launch_fused_cl_t(const FuseClusterPlan &p, cudaStream_t st, const float *K, long long ld, SolveVecs V, SolveCtrl *ctrl,
                      float *part, int max_iters) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    WOTB_CUDA(fused_cl_config<CPT>(&cfg, attrs, p.cl, p.n_clusters * p.cl, p.smem, st));
    WOTB_CUDA(cudaLaunchKernelEx(&cfg, k_fused_cl<CPT>, K, ld, V, ctrl, part, p.stages, p.lag, max_iters, p.area, p.seg4));
    return WOTB_OK;
}

inline int launch_fused_cluster(const FuseClusterPlan &p, cudaStream_t st, const float *K, long long ld, const SolveVecs &V,
                                SolveCtrl *ctrl, float *part, int max_iters) {
    switch (p.cpt) {
        case 1: return launch_fused_cl_t<1>(p, st, K, ld, V, ctrl, part, max_iters);
        case 2: return launch_fused_cl_t<2>(p, st, K, ld, V, ctrl, part, max_iters);
        case 3: return launch_fused_cl_t<3>(p, st, K, ld, V, ctrl, part, max_iters);
        case 4: return launch_fused_cl_t<4>(p, st, K, ld, V, ctrl, part, max_iters);
        case 5: return launch_fused_cl_t<5>(p, st, K, ld, V, ctrl, part, max_iters);
        case 6: return launch_fused_cl_t<6>(p, st, K, ld, V, ctrl, part, max_iters);
    }
    set_error("cluster-fused iteration: unsupported column count");
    return WOTB_ERR_INVALID;
}

}  // namespace wotb
