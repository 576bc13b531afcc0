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

// What a finished out entry does with its reduced sum s (shared by the SIMT and the tcgen05 pass kernels).
// Returns |a| or |b| for the tau test of a half-step, 0 otherwise.
template <bool COLPASS>
__device__ __forceinline__ double online_apply(int mode, int o, double s, const SolveVecs &V, SolveCtrl *ctrl,
                                               double *rowsum_out) {
    const int I = ctrl->I, J = ctrl->J;
    const int cur = ctrl->cur;
    double vmax = 0.0;
    if (mode == 0) {
        if (!COLPASS) {
            const double a = scaling_update(V.lp[o], s, ctrl->alpha1, V.lu[o]);
            V.a[cur ^ 1][o] = a;
            V.s[o] = s;
            if (ctrl->batch_done == 0) V.sfirst[o] = s;
            V.Pd[o] = (ctrl->c1 * V.u[o] - ctrl->c2 * V.nx[o] + log2(a) - log2((double)I));
            vmax = fabs(a);
        } else {
            const double b = scaling_update(ctrl->lq, s, ctrl->alpha2, V.lv[o]);
            V.b[cur ^ 1][o] = b;
            V.t[o] = s;
            V.Qd[o] = (ctrl->c1 * V.v[o] - ctrl->c2 * V.ny[o] + log2(b) - log2((double)J));
            vmax = fabs(b);
        }
    } else if (mode == 1) {
        V.s[o] = s;
    } else if (mode == 2) {
        rowsum_out[o] = V.a[cur][o] * s * (ctrl->out_scale * (double)J);
    } else if (mode == 3) {
        V.sumK0_part[o] = s;
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
    if (threadIdx.x == 0 && blockIdx.x == 0 && !ctrl->done) ctrl->need_build = 0;
}

// raw squared norms in float64 (constant for the solve)
__global__ void k_sqnorms(const double *__restrict__ x, int n, int d, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < d; ++k) s = fma(x[(long long)i * d + k], x[(long long)i * d + k], s);
    out[i] = s;
}

int sinkhorn_online_impl(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                         const double *G, const wotb_params *prm, double *f, double *g, double *rowsum,
                         wotb_info *info) {
    WOTB_REQUIRE(ctx && x0 && x1 && G && f && g && info, "NULL argument");
    WOTB_REQUIRE(d >= 1 && median > 0, "d must be >= 1 and the median positive");
    memset(info, 0, sizeof(*info));
    SolveCtrl h;
    WOTB_TRY(init_ctrl(prm, I, J, &h, median));
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    const int64_t ldi = round_up(I, kOnTile), ldj = round_up(J, kOnTile);
    const int dp = (int)round_up(d, 4);
    const int tiles_i = (int)(ldi / kOnTile), tiles_j = (int)(ldj / kOnTile);
    // segments: about two CTAs per SM in flight
    auto segs = [&](int out_tiles, int in_tiles, int *seg_tiles) {
        int nseg = (int)cdiv((int64_t)ctx->sm_count * 2, out_tiles);
        if (nseg > in_tiles) nseg = in_tiles;
        if (nseg < 1) nseg = 1;
        *seg_tiles = (int)cdiv(in_tiles, nseg);
        return (int)cdiv(in_tiles, *seg_tiles);
    };
    int seg_tiles_row = 0, seg_tiles_col = 0;
    const int nseg_row = segs(tiles_i, tiles_j, &seg_tiles_row);
    const int nseg_col = segs(tiles_j, tiles_i, &seg_tiles_col);

    // workspace: scaled coordinates (k-major fp32), norms, offsets, partials
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    const size_t o_xt = take((size_t)dp * ldi * 4), o_yt = take((size_t)dp * ldj * 4);
    const size_t o_nx = take((size_t)I * 8), o_ny = take((size_t)J * 8);
    const size_t o_ps = take((size_t)ldi * 8), o_qs = take((size_t)ldj * 8);
    const size_t o_pd = take((size_t)ldi * 8), o_qd = take((size_t)ldj * 8);
    const size_t o_p0 = take((size_t)ldi * 8), o_q0 = take((size_t)ldj * 8);
    const size_t o_part = take((size_t)(nseg_row > nseg_col ? nseg_row : nseg_col) * (ldi > ldj ? ldi : ldj) * 8);
    const size_t o_cnt = take((size_t)(tiles_i + tiles_j) * 4 + 64);
    WOTB_TRY(ctx->onl.reserve(off));
    char *ob = ctx->onl.as<char>();
    float *XT = (float *)(ob + o_xt), *YT = (float *)(ob + o_yt);
    double *nx = (double *)(ob + o_nx), *ny = (double *)(ob + o_ny);
    double *P0 = (double *)(ob + o_p0), *Q0 = (double *)(ob + o_q0);
    double *part = (double *)(ob + o_part);
    unsigned int *cnt_i = (unsigned int *)(ob + o_cnt), *cnt_j = cnt_i + tiles_i;
    WOTB_CUDA(cudaMemsetAsync(cnt_i, 0, (size_t)(tiles_i + tiles_j) * 4, st));

    SolveVecs V;
    WOTB_TRY(carve_vectors(ctx, I, J, round_up(J, 32), 1, 1, (int)I, G, f, g, &V));
    V.online = 1;
    V.nx = nx;
    V.ny = ny;
    V.Ps = (double *)(ob + o_ps);
    V.Qs = (double *)(ob + o_qs);
    V.Pd = (double *)(ob + o_pd);
    V.Qd = (double *)(ob + o_qd);
    V.n_pad_i = ldi;
    V.n_pad_j = ldj;
    V.rowsum = rowsum;
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    SolveCtrl *d_ctrl = ctx->ctrl.as<SolveCtrl>();
    WOTB_TRY(ctx->status.reserve(256));
    *ctx->status.as<int>() = 0;

    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    k_sqnorms<<<(unsigned)cdiv(I, 256), 256, 0, st>>>(x0, (int)I, d, nx);
    k_sqnorms<<<(unsigned)cdiv(J, 256), 256, 0, st>>>(x1, (int)J, d, ny);
    WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, V, d_ctrl, round_up(J, 32));

    const size_t smem = (size_t)3 * kOnChunk * kOnTile * 4;
    static bool configured = false;
    if (!configured) {
        WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    OnlineArgs row;  // reduce over j, one result per i
    row.out = {XT, ldi, V.Ps, (int)I};
    row.in = {YT, ldj, V.Qd, (int)J};
    row.dp = dp, row.nseg = nseg_row, row.seg_tiles = seg_tiles_row, row.part = part, row.counters = cnt_i;
    row.out_tile0 = 0, row.in_tile0 = 0, row.in_ntiles = tiles_j;
    OnlineArgs col;  // reduce over i, one result per j
    col.out = {YT, ldj, V.Qs, (int)J};
    col.in = {XT, ldi, V.Pd, (int)I};
    col.dp = dp, col.nseg = nseg_col, col.seg_tiles = seg_tiles_col, col.part = part, col.counters = cnt_j;
    col.out_tile0 = 0, col.in_tile0 = 0, col.in_ntiles = tiles_i;
    OnlineArgs s0 = row;
    s0.out.off = P0;
    s0.in.off = Q0;
    const dim3 grid_row(tiles_i, nseg_row), grid_col(tiles_j, nseg_col);
    const int slots = h.solver == WOTB_SOLVER_DUALITY_GAP ? 5 : 10;
    volatile int *host_done = ctx->status.as<int>();
    const unsigned pad_blocks = (unsigned)cdiv(ldi > ldj ? ldi : ldj, 256);
    auto sequence = [&]() {
        k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldi, 256), 256, 0, st>>>(x0, (int)I, d, XT, ldi, dp, d_ctrl, 0);
        k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldj, 256), 256, 0, st>>>(x1, (int)J, d, YT, ldj, dp, d_ctrl, 0);
        if (h.solver == WOTB_SOLVER_DUALITY_GAP) {
            k_online_s0_offsets<<<pad_blocks, 256, 0, st>>>(V, d_ctrl, P0, Q0);
            k_online_pass<false><<<grid_row, kOnThreads, smem, st>>>(s0, V, d_ctrl, 3, nullptr);
        }
        k_online_built<<<1, 32, 0, st>>>(d_ctrl);
        for (int s = 0; s < slots; ++s) {
            k_online_pass<false><<<grid_row, kOnThreads, smem, st>>>(row, V, d_ctrl, 0, nullptr);
            k_online_pass<true><<<grid_col, kOnThreads, smem, st>>>(col, V, d_ctrl, 0, nullptr);
        }
        launch_check(ctx, V, d_ctrl, host_done);
    };
    const int per_seq = 4 + 2 * slots + (h.solver == WOTB_SOLVER_DUALITY_GAP ? 2 : 0);
    info->launches = 3;
    int rc = pump(ctx, prm->use_graph != 0, per_seq, 2 * slots, sequence, info);
    if (rc != WOTB_OK) return rc;

    WOTB_CUDA(cudaMemcpyAsync(&h, d_ctrl, sizeof(h), cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    if (rowsum && !h.rowsum_ready) {
        // offsets and scaled coordinates are consistent with (u, v, a, b) even after a trailing absorption
        // (absorb() refreshes them; the coordinate scale only depends on eps)
        k_online_pass<false><<<grid_row, kOnThreads, smem, st>>>(row, V, d_ctrl, 2, rowsum);
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

// =================================================================================================
// Row-sharded online solve (BASELINE.json configs[3]: one 100k x 100k pair on 2/4/8 GPUs).
//
// Every rank holds all coordinates (O((I+J) d), a few MB) and the full O(I+J) solver state, replicated;
// it computes the row half-step for its slice of row tiles and the partial column sums over the same
// slice.  Two exchanges per iteration, both a SUM all-reduce of one float64 vector (the caller runs them
// with NCCL on the context's stream): the a-slices (zeros outside the slice, so the sum is an exact
// all-gather) and the partial column sums.  Because the state is replicated, the convergence checks and
// the whole state machine (k_check) run unchanged and identically on every rank: no further collective.
// =================================================================================================
__global__ void k_export_slice(const double *__restrict__ src, double *__restrict__ dst, int n, int lo, int hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (i >= lo && i < hi) ? src[i] : 0.0;
}

// a-slice into dst[0:I], row-sum slice into dst[I:2I] (the row sums feed the lazy duality-gap check)
__global__ void k_export_a_slice(SolveVecs V, SolveCtrl *ctrl, double *__restrict__ dst, int lo, int hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = ctrl->I;
    const double *a = V.a[ctrl->cur ^ 1];
    if (i < I) {
        const bool mine = i >= lo && i < hi;
        dst[i] = mine ? a[i] : 0.0;
        dst[I + i] = mine ? V.s[i] : 0.0;
    }
}

__global__ void k_import_a(SolveVecs V, SolveCtrl *ctrl, const double *__restrict__ src) {
    if (!iteration_active(ctrl)) return;
    const int I = ctrl->I;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double vmax = 0.0;
    if (i < I) {
        const double a = src[i];
        V.a[ctrl->cur ^ 1][i] = a;
        if (ctrl->batch_done == 0) V.sfirst[i] = src[I + i];
        V.Pd[i] = (ctrl->c1 * V.u[i] - ctrl->c2 * V.nx[i] + log2(a) - log2((double)I));
        vmax = fabs(a);
    }
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(&ctrl->maxabs, vmax);
}

__global__ void k_import_vec(SolveCtrl *ctrl, const double *__restrict__ src, double *__restrict__ dst, int n,
                             int what) {
    // what 1: row sums for the gap check, 3: S0 row partials
    if (what == 1 && !gap_rows_wanted(ctrl)) return;
    if (what == 3 && (ctrl->done || !ctrl->need_build || ctrl->stage != WOTB_N_STAGES - 1)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void k_online_col_finish(SolveVecs V, SolveCtrl *ctrl, const double *__restrict__ t_all) {
    if (!iteration_active(ctrl)) return;
    const int J = ctrl->J;
    const int cur = ctrl->cur;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double vmax = 0.0;
    if (j < J) {
        const double t = t_all[j];
        const double b = scaling_update(ctrl->lq, t, ctrl->alpha2, V.lv[j]);
        V.b[cur ^ 1][j] = b;
        V.t[j] = t;
        V.Qd[j] = (ctrl->c1 * V.v[j] - ctrl->c2 * V.ny[j] + log2(b) - log2((double)J));
        vmax = fabs(b);
    }
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(&ctrl->maxabs, vmax);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&ctrl->col_tiles_done, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            ctrl->col_tiles_done = 0;
            close_iteration(ctrl);
        }
    }
}

struct OnlineSolve {
    wotb_ctx *ctx;
    int64_t I, J;
    int d, dp;
    int64_t ldi, ldj;
    const double *x0, *x1;
    float *XT, *YT;
    double *P0, *Q0;
    SolveVecs V;
    SolveCtrl *d_ctrl;
    SolveCtrl h;
    OnlineArgs row, col, s0;
    dim3 grid_row, grid_col;
    int row_lo, row_hi;  // rows of this shard
    size_t smem;
    int64_t launches;
};

int online_open(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                const double *G, const wotb_params *prm, int shard, int n_shards, double *f, double *g,
                OnlineSolve **out) {
    WOTB_REQUIRE(ctx && x0 && x1 && G && f && g && out, "NULL argument");
    WOTB_REQUIRE(d >= 1 && median > 0, "d must be >= 1 and the median positive");
    WOTB_REQUIRE(n_shards >= 1 && shard >= 0 && shard < n_shards, "bad shard index");
    OnlineSolve *S = new OnlineSolve();
    S->ctx = ctx, S->I = I, S->J = J, S->d = d, S->x0 = x0, S->x1 = x1, S->launches = 0;
    int rc = init_ctrl(prm, I, J, &S->h, median);
    if (rc != WOTB_OK) {
        delete S;
        return rc;
    }
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t ldi = round_up(I, kOnTile), ldj = round_up(J, kOnTile);
    const int dp = (int)round_up(d, 4);
    S->ldi = ldi, S->ldj = ldj, S->dp = dp;
    const int tiles_i = (int)(ldi / kOnTile), tiles_j = (int)(ldj / kOnTile);
    // contiguous slices of row tiles, as even as possible
    const int t_lo = (int)((int64_t)tiles_i * shard / n_shards), t_hi = (int)((int64_t)tiles_i * (shard + 1) / n_shards);
    const int my_tiles = t_hi - t_lo;
    S->row_lo = t_lo * kOnTile;
    S->row_hi = (int)(t_hi * (int64_t)kOnTile < I ? t_hi * (int64_t)kOnTile : I);
    auto segs = [&](int out_tiles, int in_tiles, int *seg_tiles) {
        if (out_tiles < 1) out_tiles = 1;
        if (in_tiles < 1) in_tiles = 1;
        int nseg = (int)cdiv((int64_t)ctx->sm_count * 2, out_tiles);
        if (nseg > in_tiles) nseg = in_tiles;
        if (nseg < 1) nseg = 1;
        *seg_tiles = (int)cdiv(in_tiles, nseg);
        return (int)cdiv(in_tiles, *seg_tiles);
    };
    int seg_tiles_row = 0, seg_tiles_col = 0;
    const int nseg_row = segs(my_tiles, tiles_j, &seg_tiles_row);
    const int nseg_col = segs(tiles_j, my_tiles, &seg_tiles_col);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    const size_t o_xt = take((size_t)dp * ldi * 4), o_yt = take((size_t)dp * ldj * 4);
    const size_t o_nx = take((size_t)I * 8), o_ny = take((size_t)J * 8);
    const size_t o_ps = take((size_t)ldi * 8), o_qs = take((size_t)ldj * 8);
    const size_t o_pd = take((size_t)ldi * 8), o_qd = take((size_t)ldj * 8);
    const size_t o_p0 = take((size_t)ldi * 8), o_q0 = take((size_t)ldj * 8);
    const size_t o_part = take((size_t)(nseg_row > nseg_col ? nseg_row : nseg_col) * (ldi > ldj ? ldi : ldj) * 8);
    const size_t o_cnt = take((size_t)(tiles_i + tiles_j) * 4 + 64);
    WOTB_TRY(ctx->onl.reserve(off));
    char *ob = ctx->onl.as<char>();
    S->XT = (float *)(ob + o_xt), S->YT = (float *)(ob + o_yt);
    double *nx = (double *)(ob + o_nx), *ny = (double *)(ob + o_ny);
    S->P0 = (double *)(ob + o_p0), S->Q0 = (double *)(ob + o_q0);
    double *part = (double *)(ob + o_part);
    unsigned int *cnt_i = (unsigned int *)(ob + o_cnt), *cnt_j = cnt_i + tiles_i;
    WOTB_CUDA(cudaMemsetAsync(cnt_i, 0, (size_t)(tiles_i + tiles_j) * 4, st));
    WOTB_TRY(carve_vectors(ctx, I, J, round_up(J, 32), 1, 1, (int)I, G, f, g, &S->V));
    SolveVecs &V = S->V;
    V.online = 1;
    V.nx = nx, V.ny = ny;
    V.Ps = (double *)(ob + o_ps), V.Qs = (double *)(ob + o_qs);
    V.Pd = (double *)(ob + o_pd), V.Qd = (double *)(ob + o_qd);
    V.n_pad_i = ldi, V.n_pad_j = ldj;
    V.rowsum = V.r;  // row sums of a snapshot finish land here (V.r is free once the solve is done)
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    S->d_ctrl = ctx->ctrl.as<SolveCtrl>();
    WOTB_TRY(ctx->status.reserve(256));
    *ctx->status.as<int>() = 0;
    k_sqnorms<<<(unsigned)cdiv(I, 256), 256, 0, st>>>(x0, (int)I, d, nx);
    k_sqnorms<<<(unsigned)cdiv(J, 256), 256, 0, st>>>(x1, (int)J, d, ny);
    WOTB_CUDA(cudaMemcpyAsync(S->d_ctrl, &S->h, sizeof(S->h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, V, S->d_ctrl, round_up(J, 32));
    S->smem = (size_t)3 * kOnChunk * kOnTile * 4;
    WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->smem));
    WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->smem));
    S->row.out = {S->XT, ldi, V.Ps, (int)I};
    S->row.in = {S->YT, ldj, V.Qd, (int)J};
    S->row.dp = dp, S->row.nseg = nseg_row, S->row.seg_tiles = seg_tiles_row, S->row.part = part;
    S->row.counters = cnt_i, S->row.out_tile0 = t_lo, S->row.in_tile0 = 0, S->row.in_ntiles = tiles_j;
    S->col.out = {S->YT, ldj, V.Qs, (int)J};
    S->col.in = {S->XT, ldi, V.Pd, (int)I};
    S->col.dp = dp, S->col.nseg = nseg_col, S->col.seg_tiles = seg_tiles_col, S->col.part = part;
    S->col.counters = cnt_j, S->col.out_tile0 = 0, S->col.in_tile0 = t_lo, S->col.in_ntiles = my_tiles;
    S->s0 = S->row;
    S->s0.out.off = S->P0;
    S->s0.in.off = S->Q0;
    S->grid_row = dim3(my_tiles > 0 ? my_tiles : 1, nseg_row);
    S->grid_col = dim3(tiles_j, nseg_col);
    S->launches = 3;
    WOTB_CUDA(cudaGetLastError());
    *out = S;
    return WOTB_OK;
}

enum OnlineOp {
    kOpBeginA = 0,      // rescale coordinates if eps changed; final stage: S0 row partials of the slice -> exch[I]
    kOpBeginB = 1,      // take the reduced S0 partials; mark the kernel as current
    kOpRow = 2,         // row half-step on the slice; a slice -> exch[I]
    kOpColPartial = 3,  // take the gathered a; partial column sums over the slice -> exch[J]
    kOpColFinish = 4,   // take the reduced column sums; b update; close the iteration
    kOpGapRows = 5,     // final stage: row sums of the slice for the duality gap -> exch[I]
    kOpCheck = 6,       // take the gathered row sums; run the state machine
    kOpFinalRows = 7    // coupling row sums of the slice -> exch[I]
};

int online_step(OnlineSolve *S, int op, double *exch) {
    WOTB_REQUIRE(S != nullptr, "solve handle is NULL");
    wotb_ctx *ctx = S->ctx;
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const SolveVecs &V = S->V;
    SolveCtrl *c = S->d_ctrl;
    const int I = (int)S->I, J = (int)S->J;
    const unsigned bi = (unsigned)cdiv(I, 256), bj = (unsigned)cdiv(J, 256);
    const bool have_rows = S->row_hi > S->row_lo;
    const bool dg = S->h.solver == WOTB_SOLVER_DUALITY_GAP;
    switch (op) {
        case kOpBeginA:
            k_online_scale<<<(unsigned)cdiv((int64_t)S->dp * S->ldi, 256), 256, 0, st>>>(S->x0, I, S->d, S->XT, S->ldi,
                                                                                         S->dp, c, 0);
            k_online_scale<<<(unsigned)cdiv((int64_t)S->dp * S->ldj, 256), 256, 0, st>>>(S->x1, J, S->d, S->YT, S->ldj,
                                                                                         S->dp, c, 0);
            S->launches += 2;
            if (dg) {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                k_online_s0_offsets<<<(unsigned)cdiv(S->ldi > S->ldj ? S->ldi : S->ldj, 256), 256, 0, st>>>(V, c, S->P0,
                                                                                                            S->Q0);
                if (have_rows) k_online_pass<false><<<S->grid_row, kOnThreads, S->smem, st>>>(S->s0, V, c, 3, nullptr);
                k_export_slice<<<bi, 256, 0, st>>>(V.sumK0_part, exch, I, S->row_lo, S->row_hi);
                S->launches += 3;
            }
            break;
        case kOpBeginB:
            if (dg) {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                k_import_vec<<<bi, 256, 0, st>>>(c, exch, V.sumK0_part, I, 3);
            }
            k_online_built<<<1, 32, 0, st>>>(c);
            S->launches += 2;
            break;
        case kOpRow:
            WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
            if (have_rows) k_online_pass<false><<<S->grid_row, kOnThreads, S->smem, st>>>(S->row, V, c, 0, nullptr);
            k_export_a_slice<<<bi, 256, 0, st>>>(V, c, exch, S->row_lo, S->row_hi);
            S->launches += 2;
            break;
        case kOpColPartial:
            WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
            k_import_a<<<bi, 256, 0, st>>>(V, c, exch);
            WOTB_CUDA(cudaMemsetAsync(exch, 0, (size_t)J * 8, st));
            if (have_rows) k_online_pass<true><<<S->grid_col, kOnThreads, S->smem, st>>>(S->col, V, c, 4, exch);
            S->launches += 2;
            break;
        case kOpColFinish:
            WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
            k_online_col_finish<<<bj, 256, 0, st>>>(V, c, exch);
            S->launches += 1;
            break;
        case kOpGapRows:  // kept for ABI stability: the row sums of the gap now ride on kOpRow (lazy check)
            break;
        case kOpCheck:
            launch_check(ctx, V, c, ctx->status.as<int>());
            S->launches += 1;
            break;
        case kOpFinalRows:
            WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
            if (S->h.rowsum_ready) {  // converged from a snapshot: its row sums were written by the check
                k_export_slice<<<bi, 256, 0, st>>>(V.rowsum, exch, I, S->row_lo, S->row_hi);
            } else {
                WOTB_CUDA(cudaMemsetAsync(exch, 0, (size_t)I * 8, st));
                if (have_rows) k_online_pass<false><<<S->grid_row, kOnThreads, S->smem, st>>>(S->row, V, c, 2, exch);
            }
            S->launches += 1;
            break;
        default:
            WOTB_REQUIRE(false, "unknown online step");
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

int online_state(OnlineSolve *S, wotb_info *info, int *done) {
    WOTB_REQUIRE(S && info && done, "NULL argument");
    WOTB_CUDA(cudaMemcpyAsync(&S->h, S->d_ctrl, sizeof(S->h), cudaMemcpyDeviceToHost, S->ctx->stream));
    WOTB_CUDA(cudaStreamSynchronize(S->ctx->stream));
    fill_info(S->h, info);
    info->launches = S->launches;
    *done = S->h.done;
    if (S->h.done && S->h.status == WOTB_STATUS_NAN) {
        set_error("Overflow encountered in duality gap computation, please report this incident");
        return WOTB_ERR_NAN_GAP;
    }
    return WOTB_OK;
}

}  // namespace wotb
