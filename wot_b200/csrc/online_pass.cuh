// Online-kernel unbalanced Sinkhorn for sm_100a: C and K are never materialised.
//
// A half-step of optimal_transport.py:133-134 needs  s_i = sum_j K_ij b_j / J  with
// K_ij = exp((u_i + v_j - C_ij)/eps) and C_ij = |x_i - y_j|^2 / median (ot_model.py:249-252).
// In base 2, with c1 = log2(e)/eps and c2 = c1/median,
//     K_ij b_j / J = exp2( P_i + Qd_j + <X_i, Y_j> ),
//     P_i  = c1 u_i - c2 |x_i|^2,   Qd_j = c1 v_j - c2 |y_j|^2 + log2(b_j / J),   X = sqrt(2 c2) x, Y = sqrt(2 c2) y,
// so a pass is a tiled "GEMM" whose epilogue is exp2 and a row reduction: 128 x 128 tiles of the cross
// term are accumulated from shared-memory-staged coordinates (k-major, cp.async double-buffered on the
// streamed side), every entry costs d FFMA + 2 FADD + 1 MUFU.EX2, and nothing of size I x J touches HBM.
// The column half-step is the same kernel with the roles of X and Y swapped.
//
// The offsets are float64 quantities rounded once to fp32; the cross term accumulates in fp32.  The
// exponent is a difference of terms of magnitude ~c2 |x|^2 (tens to hundreds), so its absolute error is
// ~1e-5 at the default eps = 0.05 and grows like 1/eps: the Python layer selects this
// kernel only for eps >= 0.02 unless forced.
//
// Control flow is the device state machine of solver.cu (k_check): the same convergence checks,
// absorptions and epsilon stages; "building K" becomes rescaling the coordinates for the new epsilon.
#pragma once

#include "solver_state.cuh"

namespace wotb {

constexpr int kOnTile = 128;     // tile edge (rows and columns)
constexpr int kOnThreads = 256;  // 16 x 16 threads, 8 x 8 entries each
constexpr int kOnChunk = 32;     // coordinate dimensions per shared-memory chunk

struct OnlineSide {
    const float *T;   // k-major scaled coordinates: T[k * ld + index], ld % 128 == 0, zero padded
    long long ld;
    const double *off; // exponent offsets of this side (padded with -inf); rounded to fp32 on load
    int n;            // valid entries
};

struct OnlineArgs {
    OnlineSide out;   // the side that is reduced TO (one result per entry)
    OnlineSide in;    // the side that is reduced OVER
    int dp;           // padded dimension, multiple of 4
    int nseg;         // segments of the reduced-over side (grid.y)
    int seg_tiles;    // tiles per segment
    double *part;     // [nseg][out.ld] partial sums
    unsigned int *counters;  // one per out tile
    int out_tile0;    // first out tile of this launch (row-sharded solves launch a slice of the tiles)
    int in_tile0;     // first in tile that is reduced over
    int in_ntiles;    // number of in tiles reduced over
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- operand layout of the tcgen05 kernel (online_tc.cuh), needed here because the finishing code keeps the
// offset slots of the B-role rows in step with Pd / Qd -------------------------------------------------------
constexpr float kTcPad = -60000.f;  // offset of padded in rows (finite in fp16): exp2 underflows to exactly 0
// half-precision element k (0 <= k < 3*kseg: the three segments) of row r; kc = 16-byte chunks per row = 3*kseg/8
__host__ __device__ __forceinline__ long long tc_index(long long r, int k, int kc) {
    return ((r >> 3) * kc + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7);
}

constexpr float kTcLoScale = 2048.f;              // 2^11
constexpr float kTcLoInv = 1.f / 2048.f;

// x = hi + 2^-11 lo' with hi, lo' in fp16 (round to nearest).  |x| must stay below the fp16 range (6.5e4).
__device__ __forceinline__ void tc_split(double x, __half &hi, __half &lo) {
    hi = __double2half(x);
    lo = __double2half((x - (double)__half2float(hi)) * 2048.0);
}

// ---- precise operand mode (6 segments, online_tc.cuh): three 11-bit limbs per coordinate, the leading one on a
// fixed-point grid 2^-q so that the large cancelling part of the exponent accumulates EXACTLY in the fp32 TMEM
// accumulator (measured on B200, profiles/r2b_tc_accum_probe.txt: operands that share a quantum give bit-exact
// sums).  `TcGeo` holds what the grid choice needs, reduced once per solve from the coordinates.
struct TcGeo {
    double max_abs;   // max |x_ik| over both sides and all dimensions
    double max_n2x;   // max |x_i|^2
    double max_n2y;   // max |y_j|^2
};

// fractional bits of the leading limb for coordinates scaled by sc: the limb must fit fp16's 11 bits, and every
// partial sum of the leading segment (|<X,Y>| + |integer offsets|) must stay below 2^24 quanta of 2^-2q
__host__ __device__ __forceinline__ int tc_grid_q(double sc, const TcGeo &g) {
    int ib = 0;
    frexp(sc * g.max_abs * 1.0000001 + 1e-300, &ib);  // sc * max_abs < 2^ib
    if (ib < 0) ib = 0;
    const int q1 = 10 - ib;
    const double n2 = g.max_n2x > g.max_n2y ? g.max_n2x : g.max_n2y;
    int eb = 0;
    frexp(2.0 * sc * sc * n2 + 1024.0, &eb);          // bound < 2^eb
    const int q2 = (23 - eb) / 2;
    int q = q1 < q2 ? q1 : q2;
    if (q > 8) q = 8;                                  // h * 2^-(q+7) must stay a multiple of 2^-24 (fp16 subnormal step)
    if (q < 0) q = 0;
    return q;
}

// The fp32 accumulator of tcgen05.mma rounds TOWARD ZERO once per instruction: after the two exact instructions of
// the leading segment the ten that follow each cost the running exponent D half an ulp on average, i.e.
// D_computed = D (1 - kappa) with kappa = 3.1e-7, measured on B200 independent of the operand magnitudes
// (profiles/r2d_tc_accum_bias_by_magnitude.txt: mean error of log2 = 3.1e-7 |D| in every binade of |D|).  An error
// proportional to D is removed by presenting D (1 + kappa) to the tensor core: coordinates are scaled by
// sqrt(1 + kappa) and offsets by (1 + kappa) before they are split.  What remains is zero-mean.
constexpr double kTcTruncComp = 3.1e-7;

// offset = o1 + o2 + 2^-11 o3' (+ residual): o1 an integer (exact in the leading segment), o2 the fraction to fp16
// precision, o3' the next 11 bits
__device__ __forceinline__ double tc_split3(double value, __half &o1, __half &o2, __half &o3) {
    o1 = __double2half(rint(value));
    double r = value - (double)__half2float(o1);
    o2 = __double2half(r);
    r -= (double)__half2float(o2);
    o3 = __double2half(r * 2048.0);
    return r - (double)__half2float(o3) * (1.0 / 2048.0);
}

// offset `value` of in-side row `idx` into its slots (b1 | . | b2') of the 3-segment layout or (b1 | b2 | b3' | ...)
// of the 6-segment layout, see online_tc.cuh
__device__ __forceinline__ void tc_store_in_offset(__half *opB, long long idx, double value, int kseg, int nseg) {
    const int kc = nseg * (kseg >> 3);
    if (nseg == 3) {
        value = fmax(value, (double)kTcPad);
        __half b1, b2;
        tc_split(value, b1, b2);
        opB[tc_index(idx, kseg - 2, kc)] = b1;
        opB[tc_index(idx, 3 * kseg - 2, kc)] = b2;
    } else {
        value = fmax(value * (1.0 + kTcTruncComp), (double)kTcPad);
        __half b1, b2, b3;
        tc_split3(value, b1, b2, b3);
        opB[tc_index(idx, kseg - 2, kc)] = b1;
        opB[tc_index(idx, 2 * kseg - 2, kc)] = b2;
        opB[tc_index(idx, 3 * kseg - 2, kc)] = b3;
    }
}

// ---- row-sharded solves over peer memory (PeerX, solver_state.cuh): the finishing code of a pass stores its results
// into every rank's exchange buffer (own included) while the rest of the pass is still running; the stores cross
// NVLink as they are issued, and the kernel boundary in front of k_peer_barrier makes them visible system-wide.
__device__ __forceinline__ void peer_store_row(const PeerX &X, int o, double a, double s) {
    const int p = (int)(X.seq & 1ull);
    for (int w = 0; w < X.world; ++w) {
        reinterpret_cast<double *>(X.buf[w] + X.off_a[p])[o] = a;
        reinterpret_cast<double *>(X.buf[w] + X.off_s[p])[o] = s;
    }
}
__device__ __forceinline__ void peer_store_col_partial(const PeerX &X, int o, double s) {
    const int p = (int)(X.seq & 1ull);
    for (int w = 0; w < X.world; ++w) reinterpret_cast<double *>(X.buf[w] + X.off_t[p])[(long long)X.rank * X.ld_t + o] = s;
}

constexpr double kLog2e = 1.4426950408889634074;

// What a finished out entry does with its reduced sum s (shared by the SIMT and the tcgen05 pass kernels).
// Returns |a| or |b| for the tau test of a half-step, 0 otherwise.
template <bool COLPASS>
__device__ __forceinline__ double online_apply(int mode, int o, double s, const SolveVecs &V, SolveCtrl *ctrl,
                                               double *rowsum_out) {
    const int J = ctrl->J;
    const int cur = ctrl->cur;
    double vmax = 0.0;
    if (mode == 0) {
        // a = exp(z) with z = alpha (log mass - log s) + log damp (scaling_update); log2(a) is z log2(e), taken from z
        // instead of from the rounded a: log s -> z -> {a, offset} is two transcendentals deep instead of four (this
        // code is the tail of every pass for the CTA that draws the last ticket, profiles/r3d)
        if (!COLPASS) {
            const double z = fma(ctrl->alpha1, V.lp[o] - log(s), V.lu[o]);
            const double a = exp(z);
            V.a[cur ^ 1][o] = a;
            V.s[o] = s;
            if (ctrl->batch_done == 0) V.sfirst[o] = s;
            V.Pd[o] = (ctrl->c1 * V.u[o] - ctrl->c2 * V.nx[o] + z * kLog2e - ctrl->log2_I);
            if (V.tcXB) tc_store_in_offset(V.tcXB, o, V.Pd[o], V.tc_kseg, V.tc_nseg);
            if (V.peer) peer_store_row(*V.peer, o, a, s);
            vmax = fabs(a);
        } else {
            const double z = fma(ctrl->alpha2, ctrl->lq - log(s), V.lv[o]);
            const double b = exp(z);
            V.b[cur ^ 1][o] = b;
            V.t[o] = s;
            V.Qd[o] = (ctrl->c1 * V.v[o] - ctrl->c2 * V.ny[o] + z * kLog2e - ctrl->log2_J);
            if (V.tcYB) tc_store_in_offset(V.tcYB, o, V.Qd[o], V.tc_kseg, V.tc_nseg);
            vmax = fabs(b);
        }
    } else if (mode == 1) {
        V.s[o] = s;
    } else if (mode == 2) {
        rowsum_out[o] = V.a[cur][o] * s * (ctrl->out_scale * (double)J);
    } else if (mode == 3) {
        V.sumK0_part[o] = s;
    } else if (COLPASS && V.peer) {
        peer_store_col_partial(*V.peer, o, s);
    } else {
        rowsum_out[o] = s;
    }
    return vmax;
}

// mode 0: Sinkhorn half-step (row pass updates a, column pass updates b and closes the iteration)
// mode 1: row sums only (duality-gap check)       mode 2: coupling row sums after the solve
// mode 3: sum_ij exp(-C_ij/eps) row partials (final-stage `_K`, optimal_transport.py:121)
// mode 4: half-step partial sums only, written to rowsum_out (row-sharded solves reduce them across GPUs)
template <bool COLPASS>
__global__ void __launch_bounds__(kOnThreads, 2)
    k_online_pass(OnlineArgs A, SolveVecs V, SolveCtrl *ctrl, int mode, double *rowsum_out) {
    if (mode == 0 || mode == 4) {
        if (!iteration_active(ctrl)) return;
    } else if (mode == 1) {
        if (!gap_rows_wanted(ctrl)) return;
    } else if (mode == 3) {
        if (ctrl->done || !ctrl->need_build || ctrl->solver != WOTB_SOLVER_DUALITY_GAP ||
            ctrl->stage != WOTB_N_STAGES - 1)
            return;
    }
    extern __shared__ __align__(16) float smem_f[];
    float *xs = smem_f;                                   // [kOnChunk][kOnTile] out-side tile (per chunk)
    float *ys = xs + kOnChunk * kOnTile;                  // [2][kOnChunk][kOnTile] in-side tiles
    __shared__ int is_last;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int out_tile = blockIdx.x + A.out_tile0;
    const int o0 = out_tile * kOnTile;                    // first out entry of this CTA
    const int t_begin = A.in_tile0 + blockIdx.y * A.seg_tiles;
    const int t_end = min(A.in_tile0 + A.in_ntiles, t_begin + A.seg_tiles);
    const bool single_chunk = A.dp <= kOnChunk;

    float poff[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) poff[r] = (float)A.out.off[o0 + ty * 8 + r];
    double racc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) racc[r] = 0.0;

    auto load_x_chunk = [&](int k0) {
        const int kc = min(kOnChunk, A.dp - k0);
        for (int e = tid; e < kc * (kOnTile / 4); e += kOnThreads) {
            const int k = e / (kOnTile / 4), c4 = e % (kOnTile / 4);
            *reinterpret_cast<float4 *>(xs + k * kOnTile + c4 * 4) =
                *reinterpret_cast<const float4 *>(A.out.T + (long long)(k0 + k) * A.out.ld + o0 + c4 * 4);
        }
    };
    auto load_y_chunk_async = [&](int buf, int tile, int k0) {
        const int kc = min(kOnChunk, A.dp - k0);
        float *dst = ys + buf * kOnChunk * kOnTile;
        const long long c0 = (long long)tile * kOnTile;
        for (int e = tid; e < kc * (kOnTile / 4); e += kOnThreads) {
            const int k = e / (kOnTile / 4), c4 = e % (kOnTile / 4);
            cp_async16(dst + k * kOnTile + c4 * 4, A.in.T + (long long)(k0 + k) * A.in.ld + c0 + c4 * 4);
        }
        cp_async_commit();
    };

    if (single_chunk) load_x_chunk(0);
    const int n_chunks = (A.dp + kOnChunk - 1) / kOnChunk;
    // flattened (tile, chunk) pipeline over the in side
    const int n_steps = (t_end - t_begin) * n_chunks;
    if (n_steps > 0) load_y_chunk_async(0, t_begin, 0);
    float acc[8][8];
    for (int step = 0; step < n_steps; ++step) {
        const int tile = t_begin + step / n_chunks, chunk = step % n_chunks;
        const int k0 = chunk * kOnChunk;
        const int kc = min(kOnChunk, A.dp - k0);
        if (step + 1 < n_steps) {
            const int nt = t_begin + (step + 1) / n_chunks, nc = (step + 1) % n_chunks;
            load_y_chunk_async((step + 1) & 1, nt, nc * kOnChunk);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        if (!single_chunk) {
            __syncthreads();  // previous chunk's readers are done with xs
            load_x_chunk(k0);
        }
        __syncthreads();
        if (chunk == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = poff[r];
        }
        const float *yb = ys + (step & 1) * kOnChunk * kOnTile;
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 xa = *reinterpret_cast<const float4 *>(xs + k * kOnTile + ty * 8);
            const float4 xb = *reinterpret_cast<const float4 *>(xs + k * kOnTile + ty * 8 + 4);
            const float4 ya = *reinterpret_cast<const float4 *>(yb + k * kOnTile + tx * 8);
            const float4 yc = *reinterpret_cast<const float4 *>(yb + k * kOnTile + tx * 8 + 4);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yc.x, yc.y, yc.z, yc.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(xv[r], yv[c], acc[r][c]);
        }
        if (chunk == n_chunks - 1) {
            // epilogue of this tile: exp2 and the row reduction
            float qoff[8];
            const double *qsrc = A.in.off + (long long)tile * kOnTile + tx * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c) qoff[c] = (float)__ldg(qsrc + c);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) sum += ex2_approx(acc[r][c] + qoff[c]);
                racc[r] += (double)sum;
            }
        }
        __syncthreads();  // ys[(step)&1] may be overwritten by the prefetch issued next step
    }
    // reduce over the 16 threads (tx) that share each out entry: lanes of a half-warp
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) racc[r] += __shfl_xor_sync(0xffffffffu, racc[r], o);
    }
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) A.part[(long long)blockIdx.y * A.out.ld + o0 + ty * 8 + r] = racc[r];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int ticket = atomicAdd(&A.counters[out_tile], 1u);
        is_last = ticket == gridDim.y - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last CTA of this out tile: sum the segments in order and apply the update ---------------
    double vmax = 0.0;
    if (tid < kOnTile) {
        const int o = o0 + tid;
        if (o < A.out.n) {
            double s = 0.0;
            for (int sg = 0; sg < A.nseg; ++sg) s += __ldcg(A.part + (long long)sg * A.out.ld + o);
            vmax = online_apply<COLPASS>(mode, o, s, V, ctrl, rowsum_out);
        }
    }
    if (mode == 0) {
        vmax = warp_max(vmax);
        if ((tid & 31) == 0 && tid < kOnTile) atomic_max_nonneg(&ctrl->maxabs, vmax);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        A.counters[out_tile] = 0;
        if (mode == 0 && COLPASS) {
            const unsigned int ticket = atomicAdd(&ctrl->col_tiles_done, 1u);
            if (ticket == gridDim.x - 1) {
                __threadfence();
                ctrl->col_tiles_done = 0;
                close_iteration(ctrl);
            }
        }
    }
}

// Coordinates for the current epsilon: T[k][i] = (float)(sqrt(2 c2) x_ik), k-major, zero padded.
// The online analogue of rebuilding K (optimal_transport.py:124,:140): runs when need_build is set.
__global__ void k_online_scale(const double *__restrict__ x, int n, int d, float *__restrict__ T, long long ld, int dp,
                               SolveCtrl *ctrl, int clear_flag) {
    if (ctrl->done || !ctrl->need_build) return;
    const double sc = sqrt(2.0 * ctrl->c2);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)dp * ld) {
        const int k = (int)(idx / ld);
        const long long i = idx % ld;
        T[idx] = (k < d && i < n) ? (float)(sc * x[i * d + k]) : 0.f;
    }
    (void)clear_flag;
}

// S0 offsets (-c2 |x|^2) for the final stage and the need_build handshake.
__global__ void k_online_s0_offsets(SolveVecs V, SolveCtrl *ctrl, double *p0, double *q0) {
    if (ctrl->done || !ctrl->need_build) return;
    const int I = ctrl->I, J = ctrl->J;
    const double c2 = ctrl->c2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < V.n_pad_i) p0[idx] = idx < I ? (-c2 * V.nx[idx]) : -INFINITY;
    if (idx < V.n_pad_j) q0[idx] = idx < J ? (-c2 * V.ny[idx]) : -INFINITY;
}

__global__ void k_online_built(SolveCtrl *ctrl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctrl->grid_bar = 0;  // head of every launch sequence: arrival counter of the persistent batch kernel
        if (!ctrl->done) ctrl->need_build = 0;
    }
}

// raw squared norms in float64 (constant for the solve)
__global__ void k_sqnorms(const double *__restrict__ x, int n, int d, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < d; ++k) s = fma(x[(long long)i * d + k], x[(long long)i * d + k], s);
    out[i] = s;
}

}  // namespace wotb
