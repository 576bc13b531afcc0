// Stored-kernel unbalanced Sinkhorn for sm_100a: K = exp((u - C + v)/eps) lives in HBM as fp32 and
// every iteration streams it twice (K.w by rows, K^T.z by columns) with 128-bit loads.
//
// Replaces the `solver(**params)` call at /root/reference/wot/ot/optimal_transport.py:30, i.e.
// optimal_transport_duality_gap (:67-164) and transport_stablev2 (:167-236).
//
// Numerics: potentials, scalings and every convergence quantity are float64.  K and the two
// matvec operand vectors are fp32; products are accumulated in fp32 in chains of <= 8 and then
// promoted to float64, so the only fp32 error is the 6e-8 rounding of K itself (a fixed relative
// perturbation of the Gibbs kernel, SURVEY.md 7.5: 1e-6..1e-5 of the 1e-4 coupling budget).
#include <math.h>

#include "solver_state.cuh"

namespace wotb {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    // K is read once per half-iteration: do not allocate in L1; keep the default L2 policy so that
    // matrices below ~100 MB stay L2-resident between the row and the column pass.
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// Deterministic block-wide sum (fixed tree); result valid in every thread.
template <int THREADS>
__device__ double block_sum(double x, double *smem /* >= 33 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        double y = lane < THREADS / 32 ? smem[lane] : 0.0;
        y = warp_sum(y);
        if (lane == 0) smem[32] = y;
    }
    __syncthreads();
    return smem[32];
}

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long *addr, double x) {
    // non-negative doubles order like their bit patterns; NaN is ignored (x == x fails)
    if (x == x) atomicMax(addr, (unsigned long long)__double_as_longlong(x));
}

// a = (p / s)^alpha * exp(-u/(lambda+eps))  (optimal_transport.py:133-134) evaluated as
// exp(alpha (log p - log s) + lu): one log and one exp instead of a divide, a pow and a multiply, which
// is what keeps the per-row scalar work of the fused kernel off the critical path.
__device__ __forceinline__ double scaling_update(double log_mass, double s, double alpha, double log_damp) {
    return exp(fma(alpha, log_mass - log(s), log_damp));
}

__device__ __forceinline__ bool iteration_active(const SolveCtrl *c) {
    return !c->done && c->stop == 0 && c->batch_done < c->batch_iters;
}

__device__ __forceinline__ bool gap_rows_wanted(const SolveCtrl *c) {
    return !c->done && c->solver == WOTB_SOLVER_DUALITY_GAP && c->stage == WOTB_N_STAGES - 1 &&
           c->batch_done >= c->batch_iters;
}

// ------------------------------------------------------------------------------------------------
// K build: K_ij = exp((u_i + v_j - C_ij)/eps), float64 argument and exp, rounded once to fp32.
// (optimal_transport.py:124, :140, :214, :226, :197).  In the final duality-gap stage the same
// pass also reduces sum_ij exp(-C_ij/eps) -- all that primal/dual need of `_K` (:121).
// ------------------------------------------------------------------------------------------------
constexpr int kBuildThreads = 256;

// exp2(y) for a float64 exponent, rounded to fp32: the integer part is split off exactly in float64, the
// fraction (|r| <= 0.5, exact in fp32 to 3e-8) goes through MUFU.EX2 (relative error <= 2^-22), and the
// result is scaled by 2^n.  Six FP64 instructions per entry instead of the ~45 of exp(double): K builds
// become HBM-bound.  Underflow flushes to zero (entries below 1e-38 are 26 orders below the parity floor).
__device__ __forceinline__ float exp_to_f32(double y) {
    y = fmin(fmax(y, -300.0), 300.0);
    const double n = rint(y);
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"((float)(y - n)));
    return scalbnf(p, (int)n);
}

__global__ void __launch_bounds__(kBuildThreads) k_build(const float *__restrict__ C, long long ldc,
                                                         float *__restrict__ K, long long ld, SolveVecs V,
                                                         SolveCtrl *ctrl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->grid_bar = 0;  // head of every launch sequence
    if (ctrl->done || !ctrl->need_build) return;
    __shared__ double red[33];
    const int I = ctrl->I, J = ctrl->J;
    const double inv_eps_log2e = 1.4426950408889634 / ctrl->eps;
    const bool want_k0 = ctrl->solver == WOTB_SOLVER_DUALITY_GAP && ctrl->stage == WOTB_N_STAGES - 1;
    const int n4 = (int)(ld >> 2);
    double k0 = 0.0;
    for (int i = blockIdx.x; i < I; i += gridDim.x) {
        const double ui = V.u[i];
        const float4 *Crow = reinterpret_cast<const float4 *>(C + (long long)i * ldc);
        float4 *Krow = reinterpret_cast<float4 *>(K + (long long)i * ld);
        for (int j4 = threadIdx.x; j4 < n4; j4 += kBuildThreads) {
            const int j = j4 << 2;
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < J) {
                const float4 c = ldg_stream(Crow + j4);
                const float cc[4] = {c.x, c.y, c.z, c.w};
                float kk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (j + e < J) {
                        const double cij = (double)cc[e];
                        kk[e] = exp_to_f32((ui - cij + __ldg(V.v + j + e)) * inv_eps_log2e);
                        if (want_k0) k0 += (double)exp_to_f32(-cij * inv_eps_log2e);
                    }
                }
                out = make_float4(kk[0], kk[1], kk[2], kk[3]);
            }
            Krow[j4] = out;
        }
    }
    if (want_k0) {
        const double tot = block_sum<kBuildThreads>(k0, red);
        if (threadIdx.x == 0) V.sumK0_part[blockIdx.x] = tot;
    }
}

// ------------------------------------------------------------------------------------------------
// Row pass: s_i = sum_j K_ij w_j,  a_i = (p_i / s_i)^alpha1 * exp(-u_i/(lambda1+eps))
// (optimal_transport.py:133).  One warp per row, 128-bit coalesced loads, warp-shuffle reduction.
// mode 0: Sinkhorn half-step; mode 1: only s (row sums for the duality gap); mode 2: coupling row
// sums a_i s_i * scale after the solve.
// ------------------------------------------------------------------------------------------------
constexpr int kRowThreads = 256;
constexpr int kRowUnroll = 8;

__global__ void __launch_bounds__(kRowThreads) k_row(const float *__restrict__ K, long long ld, SolveVecs V,
                                                     SolveCtrl *ctrl, int mode, double *rowsum_out) {
    if (mode == 0) {
        if (!iteration_active(ctrl)) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->need_build = 0;  // K is current from here on
    } else if (mode == 1) {
        if (!gap_rows_wanted(ctrl)) return;
    }
    const int I = ctrl->I;
    const int cur = ctrl->cur;
    const bool first_of_batch = ctrl->batch_done == 0;
    const double alpha1 = ctrl->alpha1;
    const double dx = 1.0 / (double)I;
    const double out_scale = ctrl->out_scale * (double)ctrl->J;
    double *a_out = V.a[cur ^ 1];
    const double *a_cur = V.a[cur];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kRowThreads + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * kRowThreads) >> 5;
    const int n4 = (int)(ld >> 2);
    const float4 *w4 = reinterpret_cast<const float4 *>(V.w);
    double amax = 0.0;
    for (int row = warp; row < I; row += n_warps) {
        const float4 *Kr = reinterpret_cast<const float4 *>(K + (long long)row * ld);
        double acc = 0.0;
        for (int c0 = lane; c0 < n4; c0 += 32 * kRowUnroll) {
            float4 kv[kRowUnroll], wv[kRowUnroll];
#pragma unroll
            for (int q = 0; q < kRowUnroll; ++q) {
                const int c = c0 + 32 * q;
                kv[q] = c < n4 ? ldg_stream(Kr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < kRowUnroll; ++q) {
                const int c = c0 + 32 * q;
                wv[q] = c < n4 ? __ldg(w4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < kRowUnroll; ++q) {
                s0 = fmaf(kv[q].x, wv[q].x, s0);
                s1 = fmaf(kv[q].y, wv[q].y, s1);
                s2 = fmaf(kv[q].z, wv[q].z, s2);
                s3 = fmaf(kv[q].w, wv[q].w, s3);
            }
            acc += (double)((s0 + s1) + (s2 + s3));
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (mode == 0) {
                const double a = scaling_update(V.lp[row], acc, alpha1, V.lu[row]);
                a_out[row] = a;
                if (first_of_batch) V.sfirst[row] = acc;
                V.z[row] = (float)(a * dx);
                amax = fmax(amax, fabs(a));
            } else if (mode == 1) {
                V.s[row] = acc;
            } else {
                rowsum_out[row] = a_cur[row] * acc * out_scale;
            }
        }
    }
    if (mode == 0 && lane == 0) atomic_max_nonneg(&ctrl->maxabs, amax);
}

// ------------------------------------------------------------------------------------------------
// Column pass: t_j = sum_i K_ij z_i,  b_j = (q / t_j)^alpha2 * exp(-v_j/(lambda2+eps))
// (optimal_transport.py:134) without a transposed access: a CTA owns 1024 columns x one block of
// rows, each lane keeps 4 column accumulators in registers while the warp streams rows with
// 128-bit loads.  Row-block partials go to HBM; the last CTA of a column tile reduces them in a
// fixed order (deterministic, no floating-point atomics) and applies the b update; the last
// column tile of the grid closes the iteration (tau test :137, max_iter test :143).
// ------------------------------------------------------------------------------------------------
constexpr int kColThreads = 256;
constexpr int kColTile = kColThreads * 4;  // columns per CTA
constexpr int kColUnroll = 8;

__device__ void close_iteration(SolveCtrl *ctrl) {
    // single thread, after every b_j of this iteration has been written
    const unsigned long long m = atomicExch(&ctrl->maxabs, 0ull);
    ctrl->cur ^= 1;
    ctrl->iter += 1;
    ctrl->batch_done += 1;
    int stop = 0;
    if (ctrl->tau_check && __longlong_as_double((long long)m) > ctrl->tau) stop |= 1;
    if (ctrl->solver == WOTB_SOLVER_DUALITY_GAP && (double)ctrl->iter >= ctrl->max_iter) stop |= 2;
    ctrl->stop = stop;
}

__global__ void __launch_bounds__(kColThreads) k_col(const float *__restrict__ K, long long ld, SolveVecs V,
                                                     SolveCtrl *ctrl, int rows_per_block) {
    if (!iteration_active(ctrl)) return;
    __shared__ int is_last;
    const int I = ctrl->I, J = ctrl->J;
    const int cur = ctrl->cur;
    const int col = blockIdx.x * kColTile + threadIdx.x * 4;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(I, r0 + rows_per_block);
    const float *__restrict__ z = V.z;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (col < ld) {
        const float *Kc = K + col;
        for (int r = r0; r < r1; r += kColUnroll) {
            float4 kv[kColUnroll];
            float zv[kColUnroll];
#pragma unroll
            for (int q = 0; q < kColUnroll; ++q) {
                const bool ok = r + q < r1;
                kv[q] = ok ? ldg_stream(reinterpret_cast<const float4 *>(Kc + (long long)(r + q) * ld))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
                zv[q] = ok ? __ldg(z + r + q) : 0.f;
            }
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < kColUnroll; ++q) {
                s0 = fmaf(kv[q].x, zv[q], s0);
                s1 = fmaf(kv[q].y, zv[q], s1);
                s2 = fmaf(kv[q].z, zv[q], s2);
                s3 = fmaf(kv[q].w, zv[q], s3);
            }
            acc[0] += (double)s0;
            acc[1] += (double)s1;
            acc[2] += (double)s2;
            acc[3] += (double)s3;
        }
        double2 *dst = reinterpret_cast<double2 *>(V.colpart + (long long)blockIdx.y * V.ldp + col);
        dst[0] = make_double2(acc[0], acc[1]);
        dst[1] = make_double2(acc[2], acc[3]);
    }
    // ---- last CTA of this column tile reduces the row-block partials --------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&V.tile_counters[blockIdx.x], 1u);
        is_last = ticket == gridDim.y - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double alpha2 = ctrl->alpha2;
    const double lq = ctrl->lq;
    const double dy = 1.0 / (double)J;
    double *b_out = V.b[cur ^ 1];
    double bmax = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = blockIdx.x * kColTile + e * kColThreads + threadIdx.x;
        if (j < J) {
            double t = 0.0;
            for (int rb = 0; rb < (int)gridDim.y; ++rb) t += __ldcg(V.colpart + (long long)rb * V.ldp + j);
            const double b = scaling_update(lq, t, alpha2, V.lv[j]);
            b_out[j] = b;
            V.t[j] = t;
            V.w[j] = (float)(b * dy);
            bmax = fmax(bmax, fabs(b));
        }
    }
    bmax = warp_max(bmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(&ctrl->maxabs, bmax);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        V.tile_counters[blockIdx.x] = 0;
        const unsigned int ticket = atomicAdd(&ctrl->col_tiles_done, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            ctrl->col_tiles_done = 0;
            close_iteration(ctrl);
        }
    }
}

}  // namespace wotb
#include "fused_iter.cuh"
#include "fused_cluster.cuh"
namespace wotb {

// ------------------------------------------------------------------------------------------------
// State machine: one CTA, O(I + J) float64 work.
// ------------------------------------------------------------------------------------------------
constexpr int kCheckThreads = 1024;
// k_check runs as ONE thread-block cluster of kCheckCtas CTAs: its O(I + J) float64 log/exp work was 38 us per
// batch on one SM (profiles/r1e: 5 % of a solve); the CTAs split every vector loop (element e belongs to the
// same cluster thread throughout, so no cross-thread dependences arise), reductions go through distributed
// shared memory in rank order (deterministic, identical in every CTA, so control flow stays uniform), and the
// control block is written by thread 0 of rank 0 only.
constexpr int kCheckCtas = 8;
constexpr int kCheckStride = kCheckThreads * kCheckCtas;

struct CheckCluster {
    unsigned rank;
    int tid;     // thread index within the cluster
    bool lead;   // the one thread that writes SolveCtrl
    double *red;   // [33] block reduction scratch
    double *part;  // [2][8] this CTA's partial sums, read by every CTA of the cluster
    unsigned n_red;
};

// cluster_sync_all: fused_cluster.cuh

__device__ __forceinline__ double ld_dsmem(const double *local, unsigned rank) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(local), remote;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}

// Cluster-wide sums of N values; every thread of every CTA gets the same result (rank order).
template <int N>
__device__ void cluster_sum(CheckCluster &cc, double (&x)[N]) {
    static_assert(N <= 8, "at most 8 sums per reduction");
    double *mine = cc.part + (cc.n_red & 1u) * 8;
    cc.n_red += 1;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double v = block_sum<kCheckThreads>(x[k], cc.red);
        if (threadIdx.x == 0) mine[k] = v;
    }
    cluster_sync_all();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double tot = 0.0;
        for (unsigned r = 0; r < (unsigned)kCheckCtas; ++r) tot += ld_dsmem(mine + k, r);
        x[k] = tot;
    }
    // no trailing barrier: the next reduction uses the other half of `part`, and a CTA can only get two
    // reductions ahead after every CTA has arrived at the one in between, i.e. finished reading this half
}

__device__ void set_eps(SolveCtrl *c, double eps) {
    c->eps = eps;
    c->alpha1 = c->lambda1 / (c->lambda1 + eps);  // :122-123
    c->alpha2 = c->lambda2 / (c->lambda2 + eps);
    c->inv_l1e = 1.0 / (c->lambda1 + eps);
    c->inv_l2e = 1.0 / (c->lambda2 + eps);
    c->c1 = 1.4426950408889634 / eps;
    c->c2 = c->c1 * c->inv_median;
}

// u += eps log a, v += eps log b, a = b = 1 (:118-119, :138-141, :212-216, :221-228) and refresh
// everything derived from (u, v, a, b).  eps_abs is the epsilon of the absorption, eps_next the
// epsilon the following iterations run at.
__device__ void absorb(const CheckCluster &cc, const SolveVecs &V, const SolveCtrl *c, int cur, double eps_abs,
                       double eps_next) {
    const int I = c->I, J = c->J;
    const double i1 = 1.0 / (c->lambda1 + eps_next), i2 = 1.0 / (c->lambda2 + eps_next);
    const float dxf = (float)(1.0 / (double)I), dyf = (float)(1.0 / (double)J);
    const double c1 = 1.4426950408889634 / eps_next, c2 = c1 * c->inv_median;
    const double l2dx = -log2((double)I), l2dy = -log2((double)J);
    double *a = V.a[cur], *b = V.b[cur];
    for (int i = cc.tid; i < I; i += kCheckStride) {
        const double u = V.u[i] + eps_abs * log(a[i]);
        V.u[i] = u;
        a[i] = 1.0;
        V.lu[i] = -u * i1;
        if (V.online) {
            const double ps = c1 * u - c2 * V.nx[i];
            V.Ps[i] = ps;
            V.Pd[i] = (ps + l2dx);
        } else {
            V.z[i] = dxf;
        }
    }
    for (int j = cc.tid; j < J; j += kCheckStride) {
        const double v = V.v[j] + eps_abs * log(b[j]);
        V.v[j] = v;
        b[j] = 1.0;
        V.lv[j] = -v * i2;
        if (V.online) {
            const double qs = c1 * v - c2 * V.ny[j];
            V.Qs[j] = qs;
            V.Qd[j] = (qs + l2dy);
        } else {
            V.w[j] = dyf;
        }
    }
}

// `eps` is passed in: the caller may have changed c->eps in this launch, and other CTAs must not re-read it
__device__ void finish(const CheckCluster &cc, const SolveVecs &V, SolveCtrl *c, int cur, int status, double eps) {
    const int I = c->I, J = c->J;
    for (int i = cc.tid; i < I; i += kCheckStride) V.f[i] = V.u[i] + eps * log(V.a[cur][i]);
    for (int j = cc.tid; j < J; j += kCheckStride) V.g[j] = V.v[j] + eps * log(V.b[cur][j]);
    if (cc.lead) {
        c->done = 1;
        c->status = status;
        c->eps_final = eps;
        c->out_scale = status == WOTB_STATUS_MAX_ITER ? 1.0 : 1.0 / (double)J;  // :145 vs :164
    }
}

__device__ void publish(const CheckCluster &cc, SolveCtrl *c, volatile int *host_done) {
    cluster_sync_all();  // every CTA's vector writes are ordered before the flag; no CTA exits while its shared
                         // memory may still be read by a peer
    if (cc.lead) {
        c->seq += 1;
        if (c->done) {
            *host_done = 1;
            __threadfence_system();
        }
    }
}

__global__ void __launch_bounds__(kCheckThreads) k_check(SolveVecs V, SolveCtrl *ctrl, volatile int *host_done) {
    __shared__ double red[33];
    __shared__ double part[16];
    CheckCluster cc;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cc.rank));
    cc.tid = (int)cc.rank * kCheckThreads + (int)threadIdx.x;
    cc.lead = cc.tid == 0;
    cc.red = red, cc.part = part, cc.n_red = 0;
    SolveCtrl *c = ctrl;
    if (c->done) {
        publish(cc, c, host_done);
        return;
    }
    const int I = c->I, J = c->J;
    const bool complete = c->batch_done >= c->batch_iters;
    const int stop = c->stop;
    if (!complete && stop == 0) {  // the batch continues in the next replay of the sequence
        publish(cc, c, host_done);
        return;
    }
    const bool dg = c->solver == WOTB_SOLVER_DUALITY_GAP;
    const int stage = c->stage;
    const bool final_dg = dg && stage == WOTB_N_STAGES - 1;
    const int cur = c->cur;
    const double eps = c->eps;
    cluster_sync_all();

    // 0. Lazy duality gap (final stage).  primal/dual of the state at the previous batch end need the row
    //    sums K (b dy) of THAT state, which is exactly what the first iteration of the batch that followed
    //    it computed on its way (V.sfirst).  So the check of batch n is evaluated here, one batch late, from
    //    the snapshot (f, g, column sums, a) taken then: no extra sweep over K per check.  If it converged,
    //    the snapshot is the answer and the speculative batch is dropped (<= batch_size iterations per solve).
    if (final_dg && c->snap_valid && (complete || (stop & 2))) {
        const double l1 = c->lambda1, l2 = c->lambda2, qm = c->q;
        const double dx = 1.0 / (double)I, dy = 1.0 / (double)J;
        double kl1 = 0.0, kl2 = 0.0, fr = 0.0, gc = 0.0, sr = 0.0, c1 = 0.0, c2 = 0.0;
        for (int i = cc.tid; i < I; i += kCheckStride) {
            const double r = V.as[i] * V.sfirst[i] * (double)J, p = V.p[i];
            const double f = V.fs[i];
            const double x = r * dy;
            V.r[i] = r;
            kl1 += dx * (x * log(x / p) - x + p);
            fr += f * r;
            sr += r;
            c1 += (p * dx) * (exp(-f / l1) - 1.0);
        }
        for (int j = cc.tid; j < J; j += kCheckStride) {
            const double cj = V.cs[j];
            const double g = V.gs[j];
            const double y = cj * dx;
            kl2 += dy * (y * log(y / qm) - y + qm);
            gc += g * cj;
            c2 += (qm * dy) * (exp(-g / l2) - 1.0);
        }
        double k0 = 0.0;
        for (int k = cc.tid; k < V.n_sumK0_part; k += kCheckStride) k0 += V.sumK0_part[k];
        double sums[8] = {kl1, kl2, fr, gc, sr, c1, c2, k0};
        cluster_sum(cc, sums);
        kl1 = sums[0], kl2 = sums[1], fr = sums[2], gc = sums[3], sr = sums[4], c1 = sums[5], c2 = sums[6], k0 = sums[7];
        const double ij = (double)I * (double)J;
        const double pri = l1 * kl1 + l2 * kl2 + (fr + gc - eps * sr + eps * k0) / ij;
        const double dua = -l1 * c1 - l2 * c2 - eps * (sr - k0) / ij;
        const double lazy_gap = (pri - dua) / fabs(pri);
        cluster_sync_all();
        if (cc.lead) {
            c->primal = pri;
            c->dual = dua;
            c->sumK0 = k0;
            c->gap = lazy_gap;
            c->batches[stage] += 1;
            c->snap_valid = 0;
        }
        if (!(lazy_gap > c->tolerance)) {  // converged (or NaN, :129): return the snapshot
            for (int i = cc.tid; i < I; i += kCheckStride) {
                V.f[i] = V.fs[i];
                if (V.rowsum) V.rowsum[i] = V.r[i] * dy;
            }
            for (int j = cc.tid; j < J; j += kCheckStride) V.g[j] = V.gs[j];
            if (cc.lead) {
                c->done = 1;
                c->status = lazy_gap != lazy_gap ? WOTB_STATUS_NAN : WOTB_STATUS_CONVERGED;
                c->eps_final = eps;
                c->out_scale = dy;
                c->iter = c->snap_iter;
                c->rowsum_ready = V.rowsum != nullptr;
            }
            publish(cc, c, host_done);
            return;
        }
        cluster_sync_all();
    }
    // 1. column sums of R = a K b at this batch end, before any absorption (R is invariant under it; a and
    //    b are not): they go into the next snapshot
    if (final_dg && complete) {
        const double *b = V.b[cur];
        for (int j = cc.tid; j < J; j += kCheckStride) V.cs[j] = b[j] * V.t[j] * (double)I;
    }
    // 2. stabilisation (:137-141, :211-216)
    if (stop & 1) {
        absorb(cc, V, c, cur, eps, eps);
        cluster_sync_all();
        if (cc.lead) {
            c->tau_count += 1;
            c->need_build = 1;
        }
    }
    // 3. max_iter exit (:143-145)
    if (stop & 2) {
        cluster_sync_all();
        finish(cc, V, c, cur, WOTB_STATUS_MAX_ITER, eps);
        publish(cc, c, host_done);
        return;
    }
    if (!complete) {  // resume the remaining iterations of this batch after the rebuild
        cluster_sync_all();
        if (cc.lead) c->stop = 0;
        publish(cc, c, host_done);
        return;
    }
    cluster_sync_all();

    if (!dg) {
        // ---------------- transport_stablev2 schedule (:204-232) ------------------------------
        const int phase = c->phase;
        if (phase == 1) {
            finish(cc, V, c, cur, WOTB_STATUS_CONVERGED, eps);
            publish(cc, c, host_done);
            return;
        }
        const int since = c->since + c->batch_iters;
        const int sdone = c->scaling_done + c->batch_iters;
        const bool adjust = c->warm && since == c->inner_iter_max;  // :218
        double eps_next = eps;
        if (adjust) {
            const int level = stage + 1;
            eps_next = (c->epsilon0 - c->epsilon) * exp(-(double)level) + c->epsilon;  // get_reg, :184-185
            absorb(cc, V, c, cur, eps, eps_next);
        }
        cluster_sync_all();
        const bool to_extra = sdone >= c->scaling_iter;
        if (to_extra && c->extra_iter <= 0) {
            if (cc.lead && adjust) set_eps(c, eps_next);
            cluster_sync_all();
            finish(cc, V, c, cur, WOTB_STATUS_CONVERGED, adjust ? eps_next : eps);
            publish(cc, c, host_done);
            return;
        }
        if (cc.lead) {
            c->stop = 0;
            c->batch_done = 0;
            c->scaling_done = sdone;
            c->since = adjust ? 0 : since;
            if (adjust) {
                c->stage = stage + 1;
                set_eps(c, eps_next);
                c->need_build = 1;
            }
            if (to_extra) {
                c->phase = 1;
                c->tau_check = 0;
                c->batch_iters = c->extra_iter;
            } else {
                const int left = c->scaling_iter - sdone;
                c->batch_iters = c->warm ? min(c->inner_iter_max - c->since, left) : left;
            }
        }
        publish(cc, c, host_done);
        return;
    }

    // ---------------- optimal_transport_duality_gap convergence checks (:148-160) -------------
    double gap;
    if (!final_dg) {
        // max(||_a - old_a e^{u/eps}|| / (1 + ||_a||), same for b), _a = a e^{u/eps}
        const double *a = V.a[cur], *ap = V.a[cur ^ 1], *b = V.b[cur], *bp = V.b[cur ^ 1];
        const double inv_eps = 1.0 / eps;
        double na = 0.0, da = 0.0, nb = 0.0, db = 0.0;
        for (int i = cc.tid; i < I; i += kCheckStride) {
            const double e = exp(V.u[i] * inv_eps);
            const double full = __dmul_rn(a[i], e);
            const double diff = __dsub_rn(full, __dmul_rn(ap[i], e));
            na += full * full;
            da += diff * diff;
        }
        for (int j = cc.tid; j < J; j += kCheckStride) {
            const double e = exp(V.v[j] * inv_eps);
            const double full = __dmul_rn(b[j], e);
            const double diff = __dsub_rn(full, __dmul_rn(bp[j], e));
            nb += full * full;
            db += diff * diff;
        }
        double sums[4] = {na, da, nb, db};
        cluster_sum(cc, sums);
        na = sums[0], da = sums[1], nb = sums[2], db = sums[3];
        const double ga = sqrt(da) / (1.0 + sqrt(na));
        const double gb = sqrt(db) / (1.0 + sqrt(nb));
        gap = gb > ga ? gb : ga;  // Python max(ga, gb)
    } else {
        // final stage: snapshot this batch end (f and g are invariant under absorption, so taking them
        // after step 2 is the same state); its gap is evaluated by the next check (step 0)
        const double *a = V.a[cur], *b = V.b[cur];
        for (int i = cc.tid; i < I; i += kCheckStride) {
            V.fs[i] = V.u[i] + eps * log(a[i]);
            V.as[i] = a[i];
        }
        for (int j = cc.tid; j < J; j += kCheckStride) V.gs[j] = V.v[j] + eps * log(b[j]);
        cluster_sync_all();
        if (cc.lead) {
            c->snap_valid = 1;
            c->snap_iter = c->iter;
            c->stop = 0;
            c->batch_done = 0;
        }
        publish(cc, c, host_done);
        return;
    }
    const double threshold = 1e-6;      // :127 (warm stages; the final stage is handled in step 0)
    const bool again = gap > threshold;  // NaN leaves the while loop, :129
    cluster_sync_all();
    if (cc.lead) {
        c->gap = gap;
        c->batches[stage] += 1;
        c->stop = 0;
        c->batch_done = 0;
    }
    if (again) {
        publish(cc, c, host_done);
        return;
    }
    // next epsilon stage: absorb at the old epsilon (:118-119), then shrink (:120)
    const double eps_next = c->eps_sched[stage + 1];
    absorb(cc, V, c, cur, eps, eps_next);
    cluster_sync_all();
    if (cc.lead) {
        c->stage = stage + 1;
        set_eps(c, eps_next);
        c->need_build = 1;
        c->snap_valid = 0;
        c->batch_iters = (stage + 1 == WOTB_N_STAGES - 1) ? c->batch_size : 5;  // :130
    }
    publish(cc, c, host_done);
}

// Vector initialisation: u = v = 0, a = b = 1, q = mean(G) (:107-111, :192-196).
__global__ void __launch_bounds__(kCheckThreads) k_init(SolveVecs V, SolveCtrl *ctrl, long long ldw) {
    __shared__ double red[33];
    const int I = ctrl->I, J = ctrl->J;
    const float dxf = (float)(1.0 / (double)I), dyf = (float)(1.0 / (double)J);
    double sum = 0.0;
    for (int i = threadIdx.x; i < I; i += kCheckThreads) {
        sum += V.p[i];
        V.u[i] = 0.0;
        V.a[0][i] = 1.0;
        V.a[1][i] = 1.0;
        V.lu[i] = 0.0;
        V.lp[i] = log(V.p[i]);
        V.s[i] = 0.0;
        if (V.online) {
            const double ps = -ctrl->c2 * V.nx[i];
            V.Ps[i] = ps;
            V.Pd[i] = (ps - log2((double)I));
        } else {
            V.z[i] = dxf;
        }
    }
    if (V.online) {
        for (long long i = I + threadIdx.x; i < V.n_pad_i; i += kCheckThreads) V.Ps[i] = V.Pd[i] = -INFINITY;
        for (long long j = threadIdx.x; j < V.n_pad_j; j += kCheckThreads) {
            if (j < J) {
                const double qs = -ctrl->c2 * V.ny[j];
                V.v[j] = 0.0;
                V.b[0][j] = 1.0;
                V.b[1][j] = 1.0;
                V.lv[j] = 0.0;
                V.t[j] = 0.0;
                V.Qs[j] = qs;
                V.Qd[j] = (qs - log2((double)J));
            } else {
                V.Qs[j] = V.Qd[j] = -INFINITY;
            }
        }
    } else {
    for (int j = threadIdx.x; j < (int)ldw; j += kCheckThreads) {
        if (j < J) {
            V.v[j] = 0.0;
            V.b[0][j] = 1.0;
            V.b[1][j] = 1.0;
            V.lv[j] = 0.0;
            V.t[j] = 0.0;
        }
        V.w[j] = j < J ? dyf : 0.f;
    }
    }
    sum = block_sum<kCheckThreads>(sum, red);
    if (threadIdx.x == 0) {
        ctrl->q = sum / (double)I;
        ctrl->lq = log(sum / (double)I);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct VecLayout {
    SolveVecs V;
    size_t bytes;
};

static size_t take(size_t &off, size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) / 256 * 256;
    return at;
}

int carve_vectors(wotb_ctx *ctx, int64_t I, int64_t J, int64_t ldw, int n_row_blocks, int n_col_tiles,
                  int n_k0_part, const double *G, double *f, double *g, SolveVecs *out) {
    size_t off = 0;
    const size_t dI = (size_t)I * 8, dJ = (size_t)J * 8;
    const size_t o_u = take(off, dI), o_v = take(off, dJ);
    const size_t o_a0 = take(off, dI), o_a1 = take(off, dI), o_b0 = take(off, dJ), o_b1 = take(off, dJ);
    const size_t o_eu = take(off, dI), o_ev = take(off, dJ), o_s = take(off, dI), o_t = take(off, dJ);
    const size_t o_lp = take(off, dI);
    const size_t o_sf = take(off, dI), o_fs = take(off, dI), o_gs = take(off, dJ), o_cs = take(off, dJ), o_as = take(off, dI);
    const size_t o_r = take(off, dI), o_c = take(off, dJ);
    const size_t o_w = take(off, (size_t)ldw * 4), o_z = take(off, (size_t)I * 4);
    const size_t o_tc = take(off, (size_t)n_col_tiles * 4 + 64);
    const size_t o_k0 = take(off, (size_t)n_k0_part * 8 + 64);
    WOTB_TRY(ctx->vec.reserve(off));
    const int64_t ldp = round_up(ldw, 4);
    {
        const size_t unfused = (size_t)n_row_blocks * ldp * 8, fused = (size_t)ctx->sm_count * ldw * 4;
        WOTB_TRY(ctx->part.reserve(unfused > fused ? unfused : fused));
    }
    char *base = ctx->vec.as<char>();
    SolveVecs V;
    V.p = G;
    V.u = (double *)(base + o_u);
    V.v = (double *)(base + o_v);
    V.a[0] = (double *)(base + o_a0);
    V.a[1] = (double *)(base + o_a1);
    V.b[0] = (double *)(base + o_b0);
    V.b[1] = (double *)(base + o_b1);
    V.lu = (double *)(base + o_eu);
    V.lv = (double *)(base + o_ev);
    V.lp = (double *)(base + o_lp);
    V.sfirst = (double *)(base + o_sf);
    V.fs = (double *)(base + o_fs);
    V.gs = (double *)(base + o_gs);
    V.cs = (double *)(base + o_cs);
    V.as = (double *)(base + o_as);
    V.rowsum = nullptr;
    V.s = (double *)(base + o_s);
    V.t = (double *)(base + o_t);
    V.r = (double *)(base + o_r);
    V.c = (double *)(base + o_c);
    V.f = f;
    V.g = g;
    V.w = (float *)(base + o_w);
    V.z = (float *)(base + o_z);
    V.colpart = ctx->part.as<double>();
    V.tile_counters = (unsigned int *)(base + o_tc);
    V.sumK0_part = (double *)(base + o_k0);
    V.n_sumK0_part = n_k0_part;
    V.ldp = ldp;
    V.online = 0;
    V.nx = V.ny = nullptr;
    V.Ps = V.Qs = V.Pd = V.Qd = nullptr;
    V.n_pad_i = V.n_pad_j = 0;
    V.tcXB = V.tcYB = nullptr;
    V.tc_kseg = 0;
    V.tc_nseg = 3;
    V.peer = nullptr;
    WOTB_CUDA(cudaMemsetAsync(V.tile_counters, 0, (size_t)n_col_tiles * 4 + 64, ctx->stream));
    WOTB_CUDA(cudaMemsetAsync(V.sumK0_part, 0, (size_t)n_k0_part * 8 + 64, ctx->stream));
    *out = V;
    return WOTB_OK;
}

int init_ctrl(const wotb_params *prm, int64_t I, int64_t J, SolveCtrl *h, double median) {
    WOTB_REQUIRE(prm != nullptr, "params is NULL");
    WOTB_REQUIRE(I >= 1 && J >= 1 && I < (1ll << 31) && J < (1ll << 31), "I, J must be in [1, 2^31)");
    WOTB_REQUIRE(prm->solver == WOTB_SOLVER_DUALITY_GAP || prm->solver == WOTB_SOLVER_FIXED_ITERS, "unknown solver");
    WOTB_REQUIRE(prm->epsilon > 0 && prm->epsilon0 > 0, "epsilon and epsilon0 must be positive");
    memset(h, 0, sizeof(*h));
    h->I = (int)I;
    h->J = (int)J;
    h->solver = prm->solver;
    h->batch_size = prm->batch_size;
    h->lambda1 = prm->lambda1;
    h->lambda2 = prm->lambda2;
    h->tolerance = prm->tolerance;
    h->epsilon = prm->epsilon;
    h->epsilon0 = prm->epsilon0;
    h->max_iter = prm->max_iter;
    h->scaling_iter = prm->scaling_iter;
    h->extra_iter = prm->extra_iter;
    h->inner_iter_max = prm->inner_iter_max;
    h->gap = INFINITY;
    h->primal = h->dual = NAN;
    h->need_build = 1;
    h->out_scale = 1.0 / (double)J;
    h->log2_I = ::log2((double)I);
    h->log2_J = ::log2((double)J);
    const bool tau_none = prm->tau != prm->tau;
    double eps;
    if (prm->solver == WOTB_SOLVER_DUALITY_GAP) {
        WOTB_REQUIRE(prm->batch_size >= 1, "batch_size must be >= 1");
        WOTB_REQUIRE(!tau_none, "tau must be a number for the duality_gap solver");
        // same recurrence as the reference so the epsilon sequence matches to the last bit
        const double shrink = ::exp(-::log(prm->epsilon) / (double)(WOTB_N_STAGES - 1));  // :102
        double e = prm->epsilon0 * shrink;                                               // :113
        for (int s = 0; s < WOTB_N_STAGES; ++s) {
            e = e / shrink;  // :120
            h->eps_sched[s] = e;
        }
        eps = h->eps_sched[0];
        h->tau = prm->tau;
        h->tau_check = 1;
        h->batch_iters = 5;
    } else {
        WOTB_REQUIRE(prm->scaling_iter >= 0 && prm->extra_iter >= 0 && prm->scaling_iter + prm->extra_iter >= 1,
                     "scaling_iter + extra_iter must be >= 1");
        h->warm = !tau_none;                                 // :181
        WOTB_REQUIRE(!h->warm || prm->inner_iter_max >= 1, "inner_iter_max must be >= 1");
        eps = h->warm ? prm->epsilon0 : prm->epsilon;        // :187
        h->tau = tau_none ? INFINITY : prm->tau;
        h->tau_check = 1;
        if (prm->scaling_iter > 0) {
            h->phase = 0;
            h->batch_iters = h->warm ? (prm->inner_iter_max < prm->scaling_iter ? prm->inner_iter_max : prm->scaling_iter)
                                     : prm->scaling_iter;
        } else {
            h->phase = 1;
            h->tau_check = 0;
            h->batch_iters = prm->extra_iter;
        }
    }
    h->eps = eps;
    h->alpha1 = h->lambda1 / (h->lambda1 + eps);
    h->alpha2 = h->lambda2 / (h->lambda2 + eps);
    h->inv_l1e = 1.0 / (h->lambda1 + eps);
    h->inv_l2e = 1.0 / (h->lambda2 + eps);
    h->inv_median = 1.0 / median;
    h->c1 = 1.4426950408889634 / eps;
    h->c2 = h->c1 * h->inv_median;
    return WOTB_OK;
}

void fill_info(const SolveCtrl &h, wotb_info *info) {
    info->iters = h.iter;
    for (int s = 0; s < WOTB_N_STAGES; ++s) info->batches[s] = h.batches[s];
    info->tau_absorptions = h.tau_count;
    info->status = h.status;
    info->gap = h.gap;
    info->primal = h.primal;
    info->dual = h.dual;
    info->eps_final = h.eps_final;
    info->out_scale = h.out_scale;
}

void launch_init(wotb_ctx *ctx, const SolveVecs &V, SolveCtrl *d_ctrl, int64_t ldw) {
    k_init<<<1, kCheckThreads, 0, ctx->stream>>>(V, d_ctrl, ldw);
}

void launch_check(wotb_ctx *ctx, const SolveVecs &V, SolveCtrl *d_ctrl, volatile int *host_done) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(kCheckCtas), cfg.blockDim = dim3(kCheckThreads), cfg.stream = ctx->stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kCheckCtas, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_check, V, d_ctrl, host_done);
}

// Replays `sequence` (one batch worth of launches ending in the check kernel) until the device
// state machine reports done.  The host never waits on an individual batch: it keeps `depth`
// sequences in flight and only looks at a flag the check kernel raises in mapped pinned memory.
template <typename Sequence>
int pump(wotb_ctx *ctx, bool use_graph, int launches_per_seq, int matvecs_per_seq, Sequence sequence,
         wotb_info *info) {
    WOTB_TRY(ctx->status.reserve(256));
    volatile int *host_done = ctx->status.as<int>();
    constexpr int kDepth = 3;
    cudaEvent_t ev[kDepth];
    for (int k = 0; k < kDepth; ++k) WOTB_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (use_graph) {
        WOTB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        sequence();
        WOTB_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
        WOTB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    }
    int64_t enq = 0;
    int rc = WOTB_OK;
    for (;;) {
        if (use_graph) {
            cudaError_t e = cudaGraphLaunch(exec, ctx->stream);
            if (e != cudaSuccess) {
                set_error("cudaGraphLaunch: %s", cudaGetErrorString(e));
                rc = WOTB_ERR_CUDA;
                break;
            }
        } else {
            sequence();
        }
        cudaEventRecord(ev[enq % kDepth], ctx->stream);
        ++enq;
        if (enq >= kDepth) {
            cudaError_t e = cudaEventSynchronize(ev[enq % kDepth]);  // the oldest sequence in flight
            if (e != cudaSuccess) {
                set_error("solver sequence failed: %s", cudaGetErrorString(e));
                rc = WOTB_ERR_CUDA;
                break;
            }
            if (*host_done) break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == WOTB_OK && e != cudaSuccess) {
        set_error("solver stream failed: %s", cudaGetErrorString(e));
        rc = WOTB_ERR_CUDA;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (int k = 0; k < kDepth; ++k) cudaEventDestroy(ev[k]);
    info->launches += enq * launches_per_seq;
    info->matvec_launches += enq * matvecs_per_seq;
    return rc;
}

static int col_row_blocks(const wotb_ctx *ctx, int64_t I, int n_col_tiles, int *rows_per_block) {
    // enough CTAs for ~4 per SM, at least 32 rows each
    const int target = ctx->sm_count * 4;
    int nrb = (int)cdiv(target, n_col_tiles);
    int rpb = (int)cdiv(I, nrb);
    if (rpb < 32) rpb = 32;
    rpb = (int)round_up(rpb, kColUnroll);
    nrb = (int)cdiv(I, rpb);
    *rows_per_block = rpb;
    return nrb;
}

int sinkhorn_stored(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *G,
                    const wotb_params *prm, double *f, double *g, double *rowsum, wotb_info *info) {
    WOTB_REQUIRE(ctx && C && G && f && g && info, "NULL argument");
    WOTB_REQUIRE(ldc >= J && ldc % 4 == 0, "ldc must be >= J and a multiple of 4");
    WOTB_REQUIRE(((uintptr_t)C & 15) == 0, "C must be 16-byte aligned");
    memset(info, 0, sizeof(*info));
    SolveCtrl h;
    WOTB_TRY(init_ctrl(prm, I, J, &h, 1.0));
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    const int64_t ld = round_up(J, 32);
    WOTB_TRY(ctx->K.reserve((size_t)I * ld * 4));
    float *K = ctx->K.as<float>();
    const int n_col_tiles = (int)cdiv(ld, kColTile);
    int rows_per_block = 0;
    const int n_row_blocks = col_row_blocks(ctx, I, n_col_tiles, &rows_per_block);
    const int build_grid = (int)(I < ctx->sm_count * 8 ? I : ctx->sm_count * 8);
    SolveVecs V;
    WOTB_TRY(carve_vectors(ctx, I, J, ld, n_row_blocks, n_col_tiles, build_grid, G, f, g, &V));
    V.rowsum = rowsum;
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    SolveCtrl *d_ctrl = ctx->ctrl.as<SolveCtrl>();
    WOTB_TRY(ctx->status.reserve(256));
    *ctx->status.as<int>() = 0;

    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, V, d_ctrl, ld);

    const int row_grid = (int)(cdiv(I, kRowThreads / 32) < ctx->sm_count * 8 ? cdiv(I, kRowThreads / 32)
                                                                                : ctx->sm_count * 8);
    const dim3 col_grid(n_col_tiles, n_row_blocks);
    const int slots = h.solver == WOTB_SOLVER_DUALITY_GAP ? 5 : 10;
    volatile int *host_done = ctx->status.as<int>();
    const FusePlan plan = (prm->reserved & 1) ? FusePlan() : plan_fused(ctx, I, ld);
    // rows too wide for one CTA (J > 23k): a thread-block cluster shares every row (fused_cluster.cuh)
    FuseClusterPlan cplan = ((prm->reserved & 1) || plan.ok) ? FuseClusterPlan() : plan_fused_cluster(ctx, I, ld);
    float *part_f = ctx->part.as<float>();
    if (plan.ok || cplan.ok) {  // sets the dynamic shared memory attribute outside of any stream capture
        SolveCtrl idle = h;
        idle.done = 1;
        WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &idle, sizeof(idle), cudaMemcpyHostToDevice, st));
        if (plan.ok) {
            WOTB_TRY(launch_fused(plan, st, K, ld, V, d_ctrl, part_f, 1));
        } else if (launch_fused_cluster(cplan, st, K, ld, V, d_ctrl, part_f, 1) != WOTB_OK ||
                   cudaStreamSynchronize(st) != cudaSuccess) {
            cudaGetLastError();  // the cluster + cooperative launch is not available here: two sweeps per iteration
            cplan.ok = false;
        }
        WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
        launch_init(ctx, V, d_ctrl, ld);
    }
    const bool one_launch = plan.ok || cplan.ok;
    auto sequence = [&]() {
        k_build<<<build_grid, kBuildThreads, 0, st>>>(C, ldc, K, ld, V, d_ctrl);
        if (plan.ok) {
            launch_fused(plan, st, K, ld, V, d_ctrl, part_f, slots);  // one cooperative launch per batch
        } else if (cplan.ok) {
            launch_fused_cluster(cplan, st, K, ld, V, d_ctrl, part_f, slots);
        } else {
            for (int s = 0; s < slots; ++s) {
                k_row<<<row_grid, kRowThreads, 0, st>>>(K, ld, V, d_ctrl, 0, nullptr);
                k_col<<<col_grid, kColThreads, 0, st>>>(K, ld, V, d_ctrl, rows_per_block);
            }
        }
        launch_check(ctx, V, d_ctrl, host_done);
    };
    const int per_seq = 2 + (one_launch ? 1 : 2 * slots);
    info->launches = 1;
    int rc = pump(ctx, prm->use_graph != 0, per_seq, one_launch ? 1 : 2 * slots, sequence, info);
    if (rc != WOTB_OK) return rc;

    WOTB_CUDA(cudaMemcpyAsync(&h, d_ctrl, sizeof(h), cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    if (rowsum && !h.rowsum_ready) {
        if (h.need_build) {  // an absorption was the last thing that happened: bring K up to date first
            SolveCtrl tmp = h;
            tmp.done = 0;
            WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &tmp, sizeof(tmp), cudaMemcpyHostToDevice, st));
            k_build<<<build_grid, kBuildThreads, 0, st>>>(C, ldc, K, ld, V, d_ctrl);
            info->launches += 1;
        }
        k_row<<<row_grid, kRowThreads, 0, st>>>(K, ld, V, d_ctrl, 2, rowsum);
        info->launches += 1;
    }
    WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    WOTB_CUDA(cudaGetLastError());
    float ms = 0.f;
    WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    fill_info(h, info);
    info->gpu_ms = ms;
    if (h.status == WOTB_STATUS_NAN) {
        set_error("Overflow encountered in duality gap computation, please report this incident");
        return WOTB_ERR_NAN_GAP;
    }
    return WOTB_OK;
}

__global__ void k_fill(float *x, long long n, float v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = v;
}

// Average device time of the two matvec kernels on an I x J kernel matrix (bench.py roofline leg).
int bench_matvec(wotb_ctx *ctx, int64_t I, int64_t J, int reps, double *ms_row, double *ms_col, double *ms_fused) {
    WOTB_REQUIRE(ctx && ms_row && ms_col && I >= 1 && J >= 1 && reps >= 1, "bad argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    wotb_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.epsilon = 0.05, prm.lambda1 = 1, prm.lambda2 = 50, prm.epsilon0 = 1, prm.tau = INFINITY, prm.tolerance = 1e-8;
    prm.max_iter = INFINITY, prm.batch_size = 5, prm.solver = WOTB_SOLVER_DUALITY_GAP;
    SolveCtrl h;
    WOTB_TRY(init_ctrl(&prm, I, J, &h, 1.0));
    h.batch_iters = 1 << 30;
    h.need_build = 0;
    const int64_t ld = round_up(J, 32);
    WOTB_TRY(ctx->K.reserve((size_t)I * ld * 4));
    float *K = ctx->K.as<float>();
    const int n_col_tiles = (int)cdiv(ld, kColTile);
    int rows_per_block = 0;
    const int n_row_blocks = col_row_blocks(ctx, I, n_col_tiles, &rows_per_block);
    WOTB_TRY(ctx->hX.reserve((size_t)(2 * I + J) * 8 + 1024));
    double *G = ctx->hX.as<double>(), *f = G + round_up(I, 32), *g = f + round_up(I, 32);
    SolveVecs V;
    WOTB_TRY(carve_vectors(ctx, I, J, ld, n_row_blocks, n_col_tiles, 1, G, f, g, &V));
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    SolveCtrl *d_ctrl = ctx->ctrl.as<SolveCtrl>();
    std::vector<double> ones((size_t)I, 1.0);
    WOTB_CUDA(cudaMemcpyAsync(G, ones.data(), (size_t)I * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, V, d_ctrl, ld);
    k_fill<<<ctx->sm_count * 8, 256, 0, st>>>(K, (long long)I * ld, 1e-3f);
    const int row_grid = (int)(cdiv(I, kRowThreads / 32) < ctx->sm_count * 8 ? cdiv(I, kRowThreads / 32)
                                                                                : ctx->sm_count * 8);
    const dim3 col_grid(n_col_tiles, n_row_blocks);
    for (int w = 0; w < 3; ++w) {
        k_row<<<row_grid, kRowThreads, 0, st>>>(K, ld, V, d_ctrl, 0, nullptr);
        k_col<<<col_grid, kColThreads, 0, st>>>(K, ld, V, d_ctrl, rows_per_block);
    }
    float ms = 0.f;
    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    for (int r = 0; r < reps; ++r) k_row<<<row_grid, kRowThreads, 0, st>>>(K, ld, V, d_ctrl, 0, nullptr);
    WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *ms_row = ms / reps;
    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    for (int r = 0; r < reps; ++r) k_col<<<col_grid, kColThreads, 0, st>>>(K, ld, V, d_ctrl, rows_per_block);
    WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *ms_col = ms / reps;
    WOTB_CUDA(cudaGetLastError());
    if (ms_fused) {
        *ms_fused = -1.0;
        const FusePlan plan = plan_fused(ctx, I, ld);
        const FuseClusterPlan cplan = plan.ok ? FuseClusterPlan() : plan_fused_cluster(ctx, I, ld);
        if (plan.ok || cplan.ok) {
            float *part_f = ctx->part.as<float>();
            const int per_launch = 5;
            SolveCtrl zero_bar = h;
            auto reset = [&]() { return cudaMemcpyAsync(d_ctrl, &zero_bar, sizeof(zero_bar), cudaMemcpyHostToDevice, st); };
            auto run = [&]() {
                return plan.ok ? launch_fused(plan, st, K, ld, V, d_ctrl, part_f, per_launch)
                               : launch_fused_cluster(cplan, st, K, ld, V, d_ctrl, part_f, per_launch);
            };
            WOTB_CUDA(reset());
            WOTB_TRY(run());
            const int launches = (reps + per_launch - 1) / per_launch;
            WOTB_CUDA(reset());
            WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
            for (int r = 0; r < launches; ++r) {
                run();
                k_fill<<<1, 32, 0, st>>>((float *)&d_ctrl->grid_bar, 1, 0.f);  // what k_build does in a solve
            }
            WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
            WOTB_CUDA(cudaStreamSynchronize(st));
            WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            reps = launches * per_launch;
            *ms_fused = ms / reps;
            WOTB_CUDA(cudaGetLastError());
        }
    }
    return WOTB_OK;
}

}  // namespace wotb

#include "online_solve.cuh"

namespace wotb {
int sinkhorn_online(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                    const double *G, const wotb_params *prm, double *f, double *g, double *rowsum, wotb_info *info) {
    return sinkhorn_online_impl(ctx, x0, I, x1, J, d, median, G, prm, f, g, rowsum, info);
}
void online_rows(OnlineSolve *S, int64_t *lo, int64_t *hi) {
    *lo = S->P.row_lo;
    *hi = S->P.row_hi;
}
void online_close(OnlineSolve *S) {
    if (S && S->d_peer) {
        cudaSetDevice(S->ctx->device);
        cudaFree(S->d_peer);
    }
    delete S;
}
int online_done_flag(OnlineSolve *S, int *done) { return online_done(S, done); }
int64_t online_peer_bytes_of(OnlineSolve *S, int world) { return (int64_t)online_peer_bytes(S->I, S->J, world); }
int online_attach(OnlineSolve *S, int world, void *const *bufs) { return online_attach_peers(S, world, bufs); }
}  // namespace wotb
