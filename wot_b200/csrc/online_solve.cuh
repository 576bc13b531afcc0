// Host drivers of the online-kernel solves: the single-GPU solve (sinkhorn_online_impl) and the stepping
// interface of the row-sharded solve (online_open / online_step / online_state).  Both run the same device
// state machine as the stored-kernel solver (k_check in solver.cu); the passes are the tcgen05 kernel of
// online_tc.cuh when the coordinates fit its K budget (d <= 46) and the SIMT kernel of online_pass.cuh
// otherwise (or when params->reserved bit1 asks for it).
#pragma once

#include "online_pass.cuh"
#include "online_tc.cuh"

namespace wotb {

// Everything a solve needs to launch its passes, for either kernel.  `lo_blk / n_blk` select the slice of X rows
// (in units of 256) this process computes: the whole matrix for a single-GPU solve, one shard otherwise.
struct OnlinePasses {
    wotb_ctx *ctx = nullptr;
    bool tc = false;
    int64_t I = 0, J = 0, ldi = 0, ldj = 0;  // ld: lengths padded to 256
    int d = 0, dp = 0;
    const double *x0 = nullptr, *x1 = nullptr;
    double *nx = nullptr, *ny = nullptr, *P0 = nullptr, *Q0 = nullptr, *part = nullptr;
    int row_lo = 0, row_hi = 0;  // X rows of this slice
    // SIMT kernel
    float *XT = nullptr, *YT = nullptr;
    OnlineArgs row, col;
    dim3 grid_row, grid_col;
    size_t smem = 0;
    // tcgen05 kernel
    TcPlan plan;
    __half *XA = nullptr, *XB = nullptr, *YA = nullptr, *YB = nullptr;
    double *resid_x = nullptr, *resid_y = nullptr;
    TcGeo *geo = nullptr;
    TcArgs trow, tcol;
    int tgrid_row = 0, tgrid_col = 0;
    bool have_rows = true;
    bool batch_ok = false;  // the whole batch of iterations runs as one persistent cooperative launch (k_online_batch)

    // precise: 6-segment operands (exact accumulation of the large cancelling terms), for small final epsilon
    int setup(wotb_ctx *c, const double *x0_, int64_t I_, const double *x1_, int64_t J_, int d_, bool want_tc, bool precise,
              int shard, int n_shards, SolveVecs *V, cudaStream_t st, bool want_batch = false) {
        ctx = c, x0 = x0_, x1 = x1_, I = I_, J = J_, d = d_;
        tc = want_tc && tc_supported(d);
        ldi = round_up(I, kTcOut), ldj = round_up(J, kTcOut);
        dp = (int)round_up(d, 4);
        // slice of X rows, in 256-row blocks, as even as possible
        const int blocks_i = (int)(ldi / kTcOut), blocks_j = (int)(ldj / kTcOut);
        const int b_lo = (int)((int64_t)blocks_i * shard / n_shards), b_hi = (int)((int64_t)blocks_i * (shard + 1) / n_shards);
        const int my_blocks = b_hi - b_lo;
        have_rows = my_blocks > 0;
        row_lo = (int)((int64_t)b_lo * kTcOut < I ? (int64_t)b_lo * kTcOut : I);
        row_hi = (int)((int64_t)b_hi * kTcOut < I ? (int64_t)b_hi * kTcOut : I);
        const int tiles_j = (int)cdiv(J, kOnTile), tiles_jp = (int)(ldj / kOnTile);
        const int my_tiles = my_blocks * (kTcOut / kOnTile), t_lo = b_lo * (kTcOut / kOnTile);
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t at = off;
            off += (bytes + 255) / 256 * 256;
            return at;
        };
        const size_t o_nx = take((size_t)I * 8), o_ny = take((size_t)J * 8);
        const size_t o_ps = take((size_t)ldi * 8), o_qs = take((size_t)ldj * 8);
        const size_t o_pd = take((size_t)ldi * 8), o_qd = take((size_t)ldj * 8);
        const size_t o_p0 = take((size_t)ldi * 8), o_q0 = take((size_t)ldj * 8);
        const size_t o_cnt = take((size_t)(ldi + ldj) / 32 * 4 + 64);  // tcgen05 pass: one counter per 32 rows (tc_publish)
        size_t o_xt = 0, o_yt = 0, o_part = 0, o_xa = 0, o_xb = 0, o_ya = 0, o_yb = 0, o_rx = 0, o_ry = 0, o_geo = 0;
        int nseg_row = 1, nseg_col = 1, seg_tiles_row = 1, seg_tiles_col = 1, slots_row = 1, slots_col = 1;
        if (tc) {
            plan = tc_plan(d, 8, precise ? 6 : 3);
            const size_t row_bytes = plan.row_bytes();
            o_xa = take(ldi * row_bytes), o_xb = take(ldi * row_bytes), o_ya = take(ldj * row_bytes), o_yb = take(ldj * row_bytes);
            o_rx = take(ldi * 8), o_ry = take(ldj * 8), o_geo = take(sizeof(TcGeo));
            tgrid_row = tc_grid(ctx->sm_count, my_blocks > 0 ? my_blocks : 1, tiles_j, &slots_row);
            tgrid_col = tc_grid(ctx->sm_count, blocks_j, my_tiles > 0 ? my_tiles : 1, &slots_col);
            o_part = take((size_t)(slots_row > slots_col ? slots_row : slots_col) * (ldi > ldj ? ldi : ldj) * 8);
        } else {
            // segments: about two CTAs per SM in flight
            auto segs = [&](int out_tiles, int in_tiles, int *seg_tiles) {
                if (out_tiles < 1) out_tiles = 1;
                if (in_tiles < 1) in_tiles = 1;
                int nseg = (int)cdiv((int64_t)ctx->sm_count * 2, out_tiles);
                if (nseg > in_tiles) nseg = in_tiles;
                if (nseg < 1) nseg = 1;
                *seg_tiles = (int)cdiv(in_tiles, nseg);
                return (int)cdiv(in_tiles, *seg_tiles);
            };
            nseg_row = segs(my_tiles, tiles_jp, &seg_tiles_row);
            nseg_col = segs(tiles_jp, my_tiles, &seg_tiles_col);
            o_xt = take((size_t)dp * ldi * 4), o_yt = take((size_t)dp * ldj * 4);
            o_part = take((size_t)(nseg_row > nseg_col ? nseg_row : nseg_col) * (ldi > ldj ? ldi : ldj) * 8);
        }
        WOTB_TRY(ctx->onl.reserve(off));
        char *ob = ctx->onl.as<char>();
        nx = (double *)(ob + o_nx), ny = (double *)(ob + o_ny);
        P0 = (double *)(ob + o_p0), Q0 = (double *)(ob + o_q0);
        part = (double *)(ob + o_part);
        unsigned int *cnt_i = (unsigned int *)(ob + o_cnt), *cnt_j = cnt_i + ldi / 32;
        WOTB_CUDA(cudaMemsetAsync(cnt_i, 0, (size_t)(ldi + ldj) / 32 * 4, st));
        V->online = 1;
        V->nx = nx, V->ny = ny;
        V->Ps = (double *)(ob + o_ps), V->Qs = (double *)(ob + o_qs);
        V->Pd = (double *)(ob + o_pd), V->Qd = (double *)(ob + o_qd);
        V->n_pad_i = ldi, V->n_pad_j = ldj;
        if (tc) {
            WOTB_TRY(tc_configure(plan));
            XA = (__half *)(ob + o_xa), XB = (__half *)(ob + o_xb), YA = (__half *)(ob + o_ya), YB = (__half *)(ob + o_yb);
            resid_x = (double *)(ob + o_rx), resid_y = (double *)(ob + o_ry);
            V->tcXB = XB, V->tcYB = YB, V->tc_kseg = plan.kseg, V->tc_nseg = plan.nseg;
            geo = (TcGeo *)(ob + o_geo);
            WOTB_CUDA(cudaMemsetAsync(geo, 0, sizeof(TcGeo), st));
            k_tc_geo<<<(unsigned)cdiv(I, 256), 256, 0, st>>>(x0, (int)I, d, geo, 0);
            k_tc_geo<<<(unsigned)cdiv(J, 256), 256, 0, st>>>(x1, (int)J, d, geo, 1);
            trow.opA = XA, trow.opB = YB, trow.resid = resid_x, trow.out_n = (int)I, trow.out_ld = ldi;
            trow.n_blocks = my_blocks, trow.out_blk0 = b_lo, trow.in_tile0 = 0, trow.in_ntiles = tiles_j;
            trow.n_stages = plan.n_stages, trow.part = part, trow.counters = cnt_i, trow.dbg = 0, trow.prof = nullptr;
            tcol.opA = YA, tcol.opB = XB, tcol.resid = resid_y, tcol.out_n = (int)J, tcol.out_ld = ldj;
            tcol.n_blocks = blocks_j, tcol.out_blk0 = 0, tcol.in_tile0 = t_lo, tcol.in_ntiles = my_tiles;
            tcol.n_stages = plan.n_stages, tcol.part = part, tcol.counters = cnt_j, tcol.dbg = 0, tcol.prof = nullptr;
            // Optional: one persistent cooperative launch per batch (k_online_batch) when the pair is not sharded and
            // both passes fill a one-wave grid.  OFF by default: measured on B200 (profiles/r2k) it is 4 % SLOWER than
            // pass-per-launch with programmatic dependent launch (99.9 vs 96.0 us per iteration at 12.5k x 12.4k) --
            // the grid barrier + pipeline refill cost what the relaunch costs, and two solves on separate streams can
            // no longer fill each other's gaps.  WOTB_BATCH=1 or params->reserved bit4 turn it on.
            const char *nb = getenv("WOTB_BATCH");
            batch_ok = n_shards == 1 && tgrid_row == ctx->sm_count && tgrid_col == ctx->sm_count &&
                       (want_batch || (nb && nb[0] == '1')) && tc_batch_configure(plan) == WOTB_OK &&
                       tc_batch_fits(plan, ctx->sm_count, ctx->sm_count);
        } else {
            XT = (float *)(ob + o_xt), YT = (float *)(ob + o_yt);
            smem = (size_t)3 * kOnChunk * kOnTile * 4;
            WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            row.out = {XT, ldi, V->Ps, (int)I};
            row.in = {YT, ldj, V->Qd, (int)J};
            row.dp = dp, row.nseg = nseg_row, row.seg_tiles = seg_tiles_row, row.part = part, row.counters = cnt_i;
            row.out_tile0 = t_lo, row.in_tile0 = 0, row.in_ntiles = tiles_jp;
            col.out = {YT, ldj, V->Qs, (int)J};
            col.in = {XT, ldi, V->Pd, (int)I};
            col.dp = dp, col.nseg = nseg_col, col.seg_tiles = seg_tiles_col, col.part = part, col.counters = cnt_j;
            col.out_tile0 = 0, col.in_tile0 = t_lo, col.in_ntiles = my_tiles;
            grid_row = dim3(my_tiles > 0 ? my_tiles : 1, nseg_row);
            grid_col = dim3(tiles_jp, nseg_col);
        }
        k_sqnorms<<<(unsigned)cdiv(I, 256), 256, 0, st>>>(x0, (int)I, d, nx);
        k_sqnorms<<<(unsigned)cdiv(J, 256), 256, 0, st>>>(x1, (int)J, d, ny);
        return WOTB_OK;
    }

    // coordinates for the current epsilon: the online analogue of rebuilding K (gated on need_build on the device)
    int pack(cudaStream_t st, SolveCtrl *c) const {
        if (tc) {
            const int cps = plan.kseg / 8;
            k_tc_pack<<<(unsigned)cdiv(ldi * cps, 256), 256, 0, st>>>(x0, (int)I, d, ldi, plan.kseg, plan.nseg, XA, XB, c, 0.0, geo);
            k_tc_pack<<<(unsigned)cdiv(ldj * cps, 256), 256, 0, st>>>(x1, (int)J, d, ldj, plan.kseg, plan.nseg, YA, YB, c, 0.0, geo);
        } else {
            k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldi, 256), 256, 0, st>>>(x0, (int)I, d, XT, ldi, dp, c, 0);
            k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldj, 256), 256, 0, st>>>(x1, (int)J, d, YT, ldj, dp, c, 0);
        }
        return 2;
    }
    // final duality-gap stage: row partials of sum_ij exp(-C_ij / eps) over this slice (mode 3)
    int s0_pass(cudaStream_t st, const SolveVecs &V, SolveCtrl *c) const {
        const unsigned nb = (unsigned)cdiv(ldi > ldj ? ldi : ldj, 256);
        k_online_s0_offsets<<<nb, 256, 0, st>>>(V, c, P0, Q0);
        if (tc) {
            k_tc_slots<<<nb, 256, 0, st>>>(P0, (int)I, XA, resid_x, Q0, (int)J, YB, plan.kseg, plan.nseg, c, 3);
            if (have_rows) tc_launch<false>(plan, tgrid_row, st, trow, V, c, 3, nullptr);
            return 3;
        }
        OnlineArgs s0 = row;
        s0.out.off = P0;
        s0.in.off = Q0;
        if (have_rows) k_online_pass<false><<<grid_row, kOnThreads, smem, st>>>(s0, V, c, 3, nullptr);
        return 2;
    }
    // offsets of the current state into the operand slots (tcgen05 only).  gate 1: when need_build is set (after an
    // absorption or an epsilon change, which rewrite Ps/Qs/Pd/Qd); gate -1: unconditionally
    int refresh_slots(cudaStream_t st, const SolveVecs &V, SolveCtrl *c, int gate) const {
        if (!tc) return 0;
        const unsigned nb = (unsigned)cdiv(ldi > ldj ? ldi : ldj, 256);
        k_tc_slots<<<nb, 256, 0, st>>>(V.Ps, (int)I, XA, resid_x, V.Qd, (int)J, YB, plan.kseg, plan.nseg, c, gate);
        k_tc_slots<<<nb, 256, 0, st>>>(V.Qs, (int)J, YA, resid_y, V.Pd, (int)I, XB, plan.kseg, plan.nseg, c, gate);
        return 2;
    }
    // `n_iters` Sinkhorn iterations (row half-step, column half-step each); returns the number of launches
    int iterations(cudaStream_t st, const SolveVecs &V, SolveCtrl *c, int n_iters) const {
        if (batch_ok) {
            tc_batch_launch(plan, ctx->sm_count, st, trow, tcol, V, c, n_iters);
            return 1;
        }
        for (int s = 0; s < n_iters; ++s) {
            row_pass(st, V, c, 0, nullptr);
            col_pass(st, V, c, 0, nullptr);
        }
        return 2 * n_iters;
    }
    void row_pass(cudaStream_t st, const SolveVecs &V, SolveCtrl *c, int mode, double *out) const {
        if (!have_rows) return;
        if (tc)
            tc_launch<false>(plan, tgrid_row, st, trow, V, c, mode, out);
        else
            k_online_pass<false><<<grid_row, kOnThreads, smem, st>>>(row, V, c, mode, out);
    }
    void col_pass(cudaStream_t st, const SolveVecs &V, SolveCtrl *c, int mode, double *out) const {
        if (!have_rows) return;
        if (tc)
            tc_launch<true>(plan, tgrid_col, st, tcol, V, c, mode, out);
        else
            k_online_pass<true><<<grid_col, kOnThreads, smem, st>>>(col, V, c, mode, out);
    }
};

// Operand mode of the tcgen05 pass.  The default fp16 hi/lo operands accumulate the exponent with an error of about
// 5e-7 / eps that is mostly a truncation BIAS of the fp32 accumulator at the magnitude of the cancelling terms
// (profiles/r2a_online_pass_accuracy_vs_eps.txt: -4.4e-6 at eps 0.01, -1.1e-5 at 0.005); slowly converging settings
// are sensitive to such a systematic perturbation of K and end a few batches away from the reference.  Below a
// final epsilon of 0.02 (or when params->reserved bit2 asks for it; bit3 forbids it) the precise 6-segment
// operands are used: error ~1e-6 whatever epsilon, at about twice the tensor-core work.
inline bool online_precise(const wotb_params *prm) {
    if (prm->reserved & 8) return false;
    if (prm->reserved & 4) return true;
    const double eps_final = prm->solver == WOTB_SOLVER_DUALITY_GAP ? prm->epsilon * prm->epsilon0 : prm->epsilon;
    return eps_final < 0.02;
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

    SolveVecs V;
    WOTB_TRY(carve_vectors(ctx, I, J, round_up(J, 32), 1, 1, (int)I, G, f, g, &V));
    V.rowsum = rowsum;
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    SolveCtrl *d_ctrl = ctx->ctrl.as<SolveCtrl>();
    WOTB_TRY(ctx->status.reserve(256));
    *ctx->status.as<int>() = 0;

    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    OnlinePasses P;
    WOTB_TRY(P.setup(ctx, x0, I, x1, J, d, !(prm->reserved & 2), online_precise(prm), 0, 1, &V, st, (prm->reserved & 16) != 0));
    WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, V, d_ctrl, round_up(J, 32));

    const bool dg = h.solver == WOTB_SOLVER_DUALITY_GAP;
    const int slots = dg ? 5 : 10;
    volatile int *host_done = ctx->status.as<int>();
    int per_seq = 0;
    auto sequence = [&]() {
        int n = P.pack(st, d_ctrl);
        if (dg) n += P.s0_pass(st, V, d_ctrl);
        n += P.refresh_slots(st, V, d_ctrl, 1);
        k_online_built<<<1, 32, 0, st>>>(d_ctrl);
        n += P.iterations(st, V, d_ctrl, slots);
        launch_check(ctx, V, d_ctrl, host_done);
        per_seq = n + 2;
    };
    info->launches = 3;
    {  // count the launches of one sequence without running it twice: the lambda sets per_seq when it runs
        int rc = pump(ctx, prm->use_graph != 0, 0, 2 * slots, sequence, info);
        if (rc != WOTB_OK) return rc;
        info->launches += (info->matvec_launches / (2 * slots)) * per_seq;
    }

    WOTB_CUDA(cudaMemcpyAsync(&h, d_ctrl, sizeof(h), cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    if (rowsum && !h.rowsum_ready) {
        // after a trailing absorption the offsets (Ps, Qd) are current (absorb() refreshes them) but the operand
        // slots of the tcgen05 kernel are not: bring them up to date first
        info->launches += P.refresh_slots(st, V, d_ctrl, -1) + 1;
        P.row_pass(st, V, d_ctrl, 2, rowsum);
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
// slice.  ONE exchange per iteration, a SUM all-reduce of one float64 vector of 2 I + J entries (the caller
// runs it with NCCL on the context's stream): the a-slices and their row sums (zeros outside the slice, so the
// sum is an exact all-gather) followed by the partial column sums -- the column pass over a rank's own rows
// needs only that rank's a, so it does not have to wait for the gather.  Because the state is replicated, the convergence checks and
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
        if (V.tcXB) tc_store_in_offset(V.tcXB, i, V.Pd[i], V.tc_kseg, V.tc_nseg);
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
        if (V.tcYB) tc_store_in_offset(V.tcYB, j, V.Qd[j], V.tc_kseg, V.tc_nseg);
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

// =================================================================================================
// The same solve exchanging over PEER MEMORY instead of NCCL (PeerX in solver_state.cuh).
//
// Per iteration: the row half-step's finishing code stores (a_i, s_i) of the rank's own rows into every rank's
// exchange buffer, the column pass's finishing code stores the rank's partial column sums likewise -- the transfer
// rides on the passes, 256 rows at a time, instead of following them; then ONE warp exchanges flags (k_peer_barrier)
// and k_peer_finish takes the gathered a, adds the `world` partial sums of every column in rank order and applies the b
// update.  Three launches per iteration, none of them a library collective; no export / import staging, no memset.
// The S0 partials of the final stage and the coupling row sums at the end are gathers of row slices through the
// same buffers (k_peer_push_slice / k_peer_import_vec).
// =================================================================================================
__device__ __forceinline__ bool peer_gate(const SolveCtrl *c, int gate) {
    // 0: a Sinkhorn iteration is due; 3: the S0 pass of the final duality-gap stage is due; -1: always
    if (gate == 0) return iteration_active(c);
    if (gate == 3)
        return !(c->done || !c->need_build || c->solver != WOTB_SOLVER_DUALITY_GAP || c->stage != WOTB_N_STAGES - 1);
    return true;
}

__device__ __forceinline__ unsigned long long peer_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One warp: lane w tells rank w that everything this rank stored for exchange `seq` has landed (the stores were issued
// by kernels that completed before this one started), then waits for rank w's flag.  Bounded: a peer that never
// arrives (crashed process, diverged control flow) traps this context after kPeerTimeoutNs instead of hanging the GPU.
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
__global__ void k_peer_barrier(PeerX *X, const SolveCtrl *ctrl, int gate) {
    if (!peer_gate(ctrl, gate)) return;
    const int w = threadIdx.x;
    if (w >= X->world) return;
    const unsigned long long seq = X->seq, val = seq + 1;
    const int p = (int)(seq & 1ull);
    __threadfence_system();
    unsigned long long *remote = reinterpret_cast<unsigned long long *>(X->buf[w] + X->off_flags) + p * kMaxPeers + X->rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(val) : "memory");
    const unsigned long long *mine =
        reinterpret_cast<const unsigned long long *>(X->buf[X->rank] + X->off_flags) + p * kMaxPeers + w;
    const unsigned long long t0 = peer_timer_ns();
    unsigned long long got;
    unsigned int spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
        if (got < val && (++spins & 1023u) == 0 && peer_timer_ns() - t0 > kPeerTimeoutNs) __trap();
    } while (got < val);
    __threadfence_system();
}

// own slice [lo, hi) of a row vector into region a of every rank's buffer
__global__ void k_peer_push_slice(const double *__restrict__ src, int lo, int hi, PeerX *X, const SolveCtrl *ctrl, int gate) {
    if (!peer_gate(ctrl, gate)) return;
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int p = (int)(X->seq & 1ull);
    const double v = src[i];
    for (int w = 0; w < X->world; ++w) reinterpret_cast<double *>(X->buf[w] + X->off_a[p])[i] = v;
}

// gathered row vector out of the own buffer (after k_peer_barrier); the last block closes the exchange
__global__ void k_peer_import_vec(PeerX *X, const SolveCtrl *ctrl, double *__restrict__ dst, int n, int gate) {
    if (!peer_gate(ctrl, gate)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = (int)(X->seq & 1ull);
    if (i < n) dst[i] = reinterpret_cast<const double *>(X->buf[X->rank] + X->off_a[p])[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&X->ticket, 1u) == gridDim.x - 1) {
            X->ticket = 0;
            __threadfence();
            X->seq += 1;
        }
    }
}

// after k_peer_barrier: the gathered a (k_import_a) and the reduced column sums with the b update
// (k_online_col_finish) in one launch over max(I, J) entries
__global__ void k_peer_finish(SolveVecs V, SolveCtrl *ctrl, PeerX *X) {
    if (!iteration_active(ctrl)) return;
    const int I = ctrl->I, J = ctrl->J, cur = ctrl->cur;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = (int)(X->seq & 1ull);
    const unsigned char *mine = X->buf[X->rank];
    double vmax = 0.0;
    if (idx < I) {
        const double a = reinterpret_cast<const double *>(mine + X->off_a[p])[idx];
        const double s = reinterpret_cast<const double *>(mine + X->off_s[p])[idx];
        V.a[cur ^ 1][idx] = a;
        V.s[idx] = s;
        if (ctrl->batch_done == 0) V.sfirst[idx] = s;
        V.Pd[idx] = (ctrl->c1 * V.u[idx] - ctrl->c2 * V.nx[idx] + log2(a) - log2((double)I));
        if (V.tcXB) tc_store_in_offset(V.tcXB, idx, V.Pd[idx], V.tc_kseg, V.tc_nseg);
        vmax = fabs(a);
    }
    if (idx < J) {
        const double *tp = reinterpret_cast<const double *>(mine + X->off_t[p]) + idx;
        double t = 0.0;
        for (int w = 0; w < X->world; ++w) t += tp[(long long)w * X->ld_t];  // rank order: the same bits on every rank
        const double b = scaling_update(ctrl->lq, t, ctrl->alpha2, V.lv[idx]);
        V.b[cur ^ 1][idx] = b;
        V.t[idx] = t;
        V.Qd[idx] = (ctrl->c1 * V.v[idx] - ctrl->c2 * V.ny[idx] + log2(b) - log2((double)J));
        if (V.tcYB) tc_store_in_offset(V.tcYB, idx, V.Qd[idx], V.tc_kseg, V.tc_nseg);
        vmax = fmax(vmax, fabs(b));
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
            X->seq += 1;
            close_iteration(ctrl);
        }
    }
}

struct OnlineSolve {
    wotb_ctx *ctx;
    int64_t I, J;
    SolveVecs V;
    SolveCtrl *d_ctrl;
    SolveCtrl h;
    OnlinePasses P;
    int64_t launches;
    int shard = 0;
    bool peer_host_sync = false;  // WOTB_PEER_HOST_SYNC=1: see peer_barrier()
    PeerX *d_peer = nullptr;  // device copy of the peer table when the solve exchanges over peer memory (cudaMalloc)
    int world = 1;
};

inline size_t peer_align(size_t b) { return (b + 255) / 256 * 256; }

// bytes of one rank's exchange buffer for an I x J solve on `world` ranks
size_t online_peer_bytes(int64_t I, int64_t J, int world) {
    const size_t ld_t = peer_align((size_t)J * 8) / 8;
    return 256 + 2 * (2 * peer_align((size_t)I * 8) + peer_align((size_t)world * ld_t * 8));
}

// bufs[w]: rank w's exchange buffer as mapped into THIS process (own buffer at [shard]).  Zeroes the own flags; the
// caller must put a barrier over all ranks between this call and the first step.
int online_attach_peers(OnlineSolve *S, int world, void *const *bufs) {
    WOTB_REQUIRE(S && bufs, "NULL argument");
    WOTB_REQUIRE(world >= 1 && world <= kMaxPeers, "peer exchange supports 1..8 ranks");
    wotb_ctx *ctx = S->ctx;
    WOTB_CUDA(cudaSetDevice(ctx->device));
    PeerX h;
    memset(&h, 0, sizeof(h));
    h.rank = S->shard, h.world = world;
    for (int w = 0; w < world; ++w) {
        WOTB_REQUIRE(bufs[w] != nullptr, "peer buffer is NULL");
        h.buf[w] = static_cast<unsigned char *>(bufs[w]);
    }
    const size_t dI = peer_align((size_t)S->I * 8);
    h.ld_t = (long long)(peer_align((size_t)S->J * 8) / 8);
    const size_t dT = peer_align((size_t)world * h.ld_t * 8);
    size_t off = 256;
    h.off_flags = 0;
    for (int p = 0; p < 2; ++p) {
        h.off_a[p] = (long long)off, off += dI;
        h.off_s[p] = (long long)off, off += dI;
        h.off_t[p] = (long long)off, off += dT;
    }
    if (!S->d_peer) WOTB_CUDA(cudaMalloc(&S->d_peer, sizeof(PeerX)));
    cudaStream_t st = ctx->stream;
    WOTB_CUDA(cudaMemcpyAsync(S->d_peer, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemsetAsync(h.buf[h.rank], 0, 256, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    S->V.peer = S->d_peer;
    S->world = world;
    const char *hs = getenv("WOTB_PEER_HOST_SYNC");
    S->peer_host_sync = hs && hs[0] == '1';
    return WOTB_OK;
}

int online_open(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                const double *G, const wotb_params *prm, int shard, int n_shards, double *f, double *g,
                OnlineSolve **out) {
    WOTB_REQUIRE(ctx && x0 && x1 && G && f && g && out, "NULL argument");
    WOTB_REQUIRE(d >= 1 && median > 0, "d must be >= 1 and the median positive");
    WOTB_REQUIRE(n_shards >= 1 && shard >= 0 && shard < n_shards, "bad shard index");
    OnlineSolve *S = new OnlineSolve();
    S->ctx = ctx, S->I = I, S->J = J, S->launches = 0, S->shard = shard;
    int rc = init_ctrl(prm, I, J, &S->h, median);
    if (rc == WOTB_OK && cudaSetDevice(ctx->device) != cudaSuccess) rc = WOTB_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    if (rc == WOTB_OK) rc = carve_vectors(ctx, I, J, round_up(J, 32), 1, 1, (int)I, G, f, g, &S->V);
    if (rc == WOTB_OK)
        rc = S->P.setup(ctx, x0, I, x1, J, d, !(prm->reserved & 2), online_precise(prm), shard, n_shards, &S->V, st);
    if (rc == WOTB_OK) rc = ctx->ctrl.reserve(sizeof(SolveCtrl));
    if (rc == WOTB_OK) rc = ctx->status.reserve(256);
    if (rc != WOTB_OK) {
        delete S;
        return rc;
    }
    S->V.rowsum = S->V.r;  // row sums of a snapshot finish land here (V.r is free once the solve is done)
    S->d_ctrl = ctx->ctrl.as<SolveCtrl>();
    *ctx->status.as<int>() = 0;
    WOTB_CUDA(cudaMemcpyAsync(S->d_ctrl, &S->h, sizeof(S->h), cudaMemcpyHostToDevice, st));
    launch_init(ctx, S->V, S->d_ctrl, round_up(J, 32));
    S->launches = 3;
    WOTB_CUDA(cudaGetLastError());
    *out = S;
    return WOTB_OK;
}

// The flag exchange.  Ranks that are THREADS of one process on ONE GPU (the single-GPU test of the protocol) can share
// a hardware work queue: a kernel of rank B may then sit behind rank A's next kernel, which waits for A's flag kernel,
// which waits for B -- a false dependency the hardware cannot resolve.  With WOTB_PEER_HOST_SYNC=1 the host waits for
// the flag kernel before it enqueues anything behind it, so a queue never holds a blocked kernel.  One process per
// GPU (the product configuration) has no such coupling and leaves it off.
inline int peer_barrier(OnlineSolve *S, cudaStream_t st, int gate) {
    k_peer_barrier<<<1, 32, 0, st>>>(S->d_peer, S->d_ctrl, gate);
    if (S->peer_host_sync) WOTB_CUDA(cudaStreamSynchronize(st));
    return WOTB_OK;
}

enum OnlineOp {
    kOpBeginA = 0,      // rescale coordinates if eps changed; final stage: S0 row partials of the slice -> exch[I]
    kOpBeginB = 1,      // take the reduced S0 partials; mark the kernel as current
    kOpRow = 2,         // row half-step on the slice; a slice and its row sums -> exch[0 : 2I]
    kOpColPartial = 3,  // partial column sums over the slice (own rows' a only) -> exch[2I : 2I + J]
    kOpColFinish = 4,   // after ONE all-reduce of exch[0 : 2I + J]: take the gathered a and the reduced column
                        // sums; b update; close the iteration
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
    const OnlinePasses &P = S->P;
    SolveCtrl *c = S->d_ctrl;
    const int I = (int)S->I, J = (int)S->J;
    const unsigned bi = (unsigned)cdiv(I, 256), bj = (unsigned)cdiv(J, 256);
    const bool dg = S->h.solver == WOTB_SOLVER_DUALITY_GAP;
    PeerX *X = S->d_peer && S->V.peer ? S->d_peer : nullptr;
    const int n_own = P.row_hi - P.row_lo;
    const unsigned bown = (unsigned)cdiv(n_own > 0 ? n_own : 1, 256);
    switch (op) {
        case kOpBeginA:
            S->launches += P.pack(st, c);
            if (dg) {
                WOTB_REQUIRE(exch != nullptr || X, "exchange buffer is NULL");
                S->launches += P.s0_pass(st, V, c);
                if (X) {
                    if (n_own > 0) k_peer_push_slice<<<bown, 256, 0, st>>>(V.sumK0_part, P.row_lo, P.row_hi, X, c, 3);
                } else {
                    k_export_slice<<<bi, 256, 0, st>>>(V.sumK0_part, exch, I, P.row_lo, P.row_hi);
                }
                S->launches += 1;
            }
            S->launches += P.refresh_slots(st, V, c, 1);
            break;
        case kOpBeginB:
            if (dg && X) {
                WOTB_TRY(peer_barrier(S, st, 3));
                k_peer_import_vec<<<bi, 256, 0, st>>>(X, c, V.sumK0_part, I, 3);
                S->launches += 1;
            } else if (dg) {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                k_import_vec<<<bi, 256, 0, st>>>(c, exch, V.sumK0_part, I, 3);
            }
            k_online_built<<<1, 32, 0, st>>>(c);
            S->launches += 2;
            break;
        case kOpRow:
            P.row_pass(st, V, c, 0, nullptr);  // peer mode: the finishing code stores (a, s) into every rank's buffer
            S->launches += 1;
            if (!X) {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                k_export_a_slice<<<bi, 256, 0, st>>>(V, c, exch, P.row_lo, P.row_hi);
                S->launches += 1;
            }
            break;
        case kOpColPartial:
            if (X) {
                P.col_pass(st, V, c, 4, nullptr);  // partial column sums go straight into every rank's buffer
            } else {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                WOTB_CUDA(cudaMemsetAsync(exch + 2 * (size_t)I, 0, (size_t)J * 8, st));
                P.col_pass(st, V, c, 4, exch + 2 * (size_t)I);
            }
            S->launches += 1;
            break;
        case kOpColFinish:
            if (X) {
                WOTB_TRY(peer_barrier(S, st, 0));
                k_peer_finish<<<bi > bj ? bi : bj, 256, 0, st>>>(V, c, X);
            } else {
                WOTB_REQUIRE(exch != nullptr, "exchange buffer is NULL");
                k_import_a<<<bi, 256, 0, st>>>(V, c, exch);
                k_online_col_finish<<<bj, 256, 0, st>>>(V, c, exch + 2 * (size_t)I);
            }
            S->launches += 2;
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
                k_export_slice<<<bi, 256, 0, st>>>(V.rowsum, exch, I, P.row_lo, P.row_hi);
            } else {
                WOTB_CUDA(cudaMemsetAsync(exch, 0, (size_t)I * 8, st));
                S->launches += P.refresh_slots(st, V, c, -1);
                P.row_pass(st, V, c, 2, exch);
            }
            S->launches += 1;
            if (X) {  // gather the slices through the peer buffers: exch[0:I] is complete on return
                if (n_own > 0) k_peer_push_slice<<<bown, 256, 0, st>>>(exch, P.row_lo, P.row_hi, X, c, -1);
                WOTB_TRY(peer_barrier(S, st, -1));
                k_peer_import_vec<<<bi, 256, 0, st>>>(X, c, exch, I, -1);
                S->launches += 3;
            }
            break;
        default:
            WOTB_REQUIRE(false, "unknown online step");
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// Non-blocking: has the device state machine raised its done flag (mapped page-locked memory, written by k_check)?
// Lets the caller keep several batches in flight instead of synchronising on every one (online_state does).
int online_done(OnlineSolve *S, int *done) {
    WOTB_REQUIRE(S && done, "NULL argument");
    *done = *reinterpret_cast<volatile int *>(S->ctx->status.as<int>()) != 0;
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
