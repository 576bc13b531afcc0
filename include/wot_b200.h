/*
 * wot_b200.h -- C ABI of the B200-native Waddington-OT transport-map hot path.
 *
 * The reference (broadinstitute/wot) is pure Python and has no FFI; its seams for this path are
 * Python call sites.  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference checkout); INTEGRATION.md shows the ctypes binding a wot
 * maintainer would add at those call sites.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types cross this boundary.
 *   - "_dev" entry points take DEVICE pointers on the context's device and enqueue on the
 *     context's stream; they return after the work has completed unless stated otherwise.
 *   - "_host" entry points take HOST pointers, copy in, compute on the GPU, copy out.
 *   - matrices are row-major; "ld" arguments are leading dimensions in ELEMENTS.
 *   - every function returns WOTB_OK (0) or a WOTB_ERR_* code; wotb_last_error() gives the text.
 *   - a context is not thread-safe (the reference path is single-threaded, ot_model.py:182-199).
 *   - there is NO CPU fallback: without a CUDA device wotb_create fails.
 */
#ifndef WOT_B200_H
#define WOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WOTB_N_STAGES 6 /* epsilon_scalings + 1, optimal_transport.py:101,116 */

enum wotb_error {
    WOTB_OK = 0,
    WOTB_ERR_INVALID = 1,  /* bad argument */
    WOTB_ERR_CUDA = 2,     /* CUDA runtime error */
    WOTB_ERR_NOMEM = 3,    /* device or pinned allocation failed */
    WOTB_ERR_NAN_GAP = 4   /* optimal_transport.py:162-163: NaN duality gap -> RuntimeError */
};

enum wotb_solver {
    WOTB_SOLVER_DUALITY_GAP = 0, /* optimal_transport_duality_gap, optimal_transport.py:67-164 */
    WOTB_SOLVER_FIXED_ITERS = 1  /* transport_stablev2,            optimal_transport.py:167-236 */
};

enum wotb_kernel {
    WOTB_KERNEL_STORED = 0, /* K = exp((u-C+v)/eps) kept in HBM as fp32; matvecs are HBM-bound   */
    WOTB_KERNEL_ONLINE = 1  /* K recomputed tile by tile from coordinates (tcgen05 exponent + MUFU.EX2 epilogue
                               for d <= 46, SIMT FP32 otherwise); MUFU-bound                       */
};

enum wotb_status {
    WOTB_STATUS_CONVERGED = 0, /* returned R / J                          (optimal_transport.py:164) */
    WOTB_STATUS_MAX_ITER = 1,  /* returned a*K*b, NOT divided by J        (optimal_transport.py:143-145) */
    WOTB_STATUS_NAN = 2        /* final gap is NaN                        (optimal_transport.py:162-163) */
};

enum wotb_dtype { WOTB_F32 = 0, WOTB_F64 = 1 };

/* The keys of OTModel.ot_config that reach the solvers (ot_model.py:85-87, splatted at :318). */
typedef struct wotb_params {
    double epsilon;
    double lambda1;
    double lambda2;
    double epsilon0;
    double tau;       /* stabilisation threshold; NaN means Python None (fixed_iters: no warm start) */
    double tolerance;
    double max_iter;  /* the reference passes 1e7 as a float */
    int32_t batch_size;
    int32_t scaling_iter;
    int32_t extra_iter;
    int32_t inner_iter_max;
    int32_t solver; /* enum wotb_solver */
    int32_t kernel; /* enum wotb_kernel */
    int32_t use_graph; /* 1: replay the per-batch launch sequence as a CUDA graph */
    int32_t reserved;  /* flags; bit0 = 1 disables the fused one-sweep iteration kernel (two matvec kernels instead);
                          bit1 = 1 runs the online kernel on the SIMT FP32 pass instead of the tcgen05 pass;
                          bit2 / bit3 force / forbid the precise 6-segment operands of the tcgen05 pass (default: precise
                          below a final epsilon of 0.02); bit4 = 1 runs every batch of online iterations as one
                          persistent cooperative launch */
} wotb_params;

/* What the reference keeps as locals of the solver; returned for the parity criteria. */
typedef struct wotb_info {
    int64_t iters;                   /* current_iter                                   */
    int32_t batches[WOTB_N_STAGES];  /* convergence checks per epsilon stage           */
    int32_t tau_absorptions;         /* times optimal_transport.py:137-141 fired       */
    int32_t status;                  /* enum wotb_status                               */
    double gap;                      /* last duality_gap value                         */
    double primal;
    double dual;
    double eps_final;                /* epsilon_i at return                            */
    double out_scale;                /* 1/J, or 1 on the max_iter exit                 */
    double gpu_ms;                   /* device time of the solve (CUDA events)         */
    int64_t launches;                /* kernels launched by this call                  */
    int64_t matvec_launches;         /* of which K.w / K^T.z matvec kernels            */
} wotb_info;

typedef struct wotb_ctx wotb_ctx;

const char *wotb_version(void);
const char *wotb_last_error(void);

/* One context per (process, device).  cuda_stream may be NULL (context-owned stream) or a
 * cudaStream_t created by the caller (e.g. torch.cuda.current_stream().cuda_stream). */
int wotb_create(int device, void *cuda_stream, wotb_ctx **out);
void wotb_destroy(wotb_ctx *ctx);
int wotb_sync(wotb_ctx *ctx);
/* Bytes of device memory currently held by the context's workspaces. */
size_t wotb_workspace_bytes(const wotb_ctx *ctx);
void wotb_release_workspace(wotb_ctx *ctx);
/* Process-wide: at most n host-buffer calls (wotb_transport_map_from_*_host) are in their solve phase at a
 * time; a call gives its slot back before its coupling is copied to the host.  Used when several contexts
 * are driven from several threads (wot_b200/pipeline.py): n + 1 contexts with n slots keep n solves on the
 * SMs while one coupling crosses PCIe.  n <= 0 (default): no limit.  Not in the reference (its loop over
 * day-pairs, ot_model.py:182-199, is serial); the day-pairs are independent. */
void wotb_set_compute_slots(int32_t n);
/* Process-wide: launch the online pass kernels with programmatic dependent launch (default on: the next
 * half-iteration's CTAs set themselves up while the current kernel drains; +3 % with one solve at a time).
 * wot_b200/pipeline.py turns it off while several solves share the GPU (it costs 2 % there). */
void wotb_set_pdl(int32_t on);

/* Per context: size every persistent / one-wave grid of this context's kernels for n SMs instead of all of them
 * (n <= 0: all).  Two contexts with complementary limits split the GPU between an HBM-bound stored-K solve and a
 * MUFU-bound online-K solve of two different day-pairs (wot_b200/pipeline.py, mixed mode). */
int wotb_set_sm_limit(wotb_ctx *ctx, int32_t n);

/* ---- local PCA: replaces compute_pca, wot/ot/util.py:240-255 -------------------
 * m1 [n1, genes], m2 [n2, genes] float64 row-major (the two days' expression rows, util.py:241-244).
 * Computes what sklearn.decomposition.PCA(k, random_state=58951).fit(x.T) computes with its randomized
 * solver (the one svd_solver='auto' picks when max(shape) > 500 and k < 0.8 min(shape)): x = vstack - gene
 * means (:245-246), per-cell centring, randomized range finder with `size` = k + n_oversamples (10) columns
 * and n_iter power iterations (7 if k < 0.1 min(shape) else 4), small SVD.  q0 [min(genes, n1+n2) rows... see
 * below, size] is the Gaussian test matrix numpy.random.RandomState(58951).normal(size=(short side, size))
 * made by the caller so that the result equals scikit-learn's up to sign and roundoff; "short side" is genes
 * when genes < n1 + n2, else n1 + n2.  Outputs: comp [n1 + n2, k] = pca.components_.T (util.py:250),
 * singular_values [k] (ot_model.py:301), gene_means [genes] (util.py:245; may be NULL), cell_means [n1 + n2] =
 * sklearn's pca.mean_, the per-cell mean over genes of the gene-centred matrix (may be NULL), gpu_ms (may be NULL).
 * Everything is float64 and reduced in a fixed order (deterministic). */
int wotb_pca_host(wotb_ctx *ctx, const double *m1_host, int64_t n1, const double *m2_host, int64_t n2, int64_t genes,
                  int32_t k, const double *q0_host, int32_t size, int32_t n_iter, double *comp_host,
                  double *singular_values_host, double *gene_means_host, double *cell_means_host, double *gpu_ms);

/* ---- cost: replaces OTModel.compute_default_cost_matrix, ot_model.py:242-253 ------------------
 * x0 [I,d], x1 [J,d] float64 row-major; scale [d] = singular values (the diagonal of `eigenvals`,
 * ot_model.py:301) or NULL.  Distances are sum_k (x0_ik s_k - x1_jk s_k)^2 in float64 with the
 * operation order of scipy cdist('sqeuclidean') (ot_model.py:249-251). */

/* exact np.median over all I*J distances (ot_model.py:252), which are recomputed and never stored.  From 6e7
 * distances on: a random sample brackets the median, ONE pass counts the distances below the bracket and gathers the
 * ones inside, a radix select on the gathered values gives the exact order statistics (the device verifies that the
 * bracket holds them).  Otherwise, or when that check fails: a 64-bit radix select in three passes over the distances.
 * *median_host receives the value. */
int wotb_cost_median_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                         const double *scale, double *median_host);
/* The same exact median with its one pass over the distances split over ROW SHARDS (row-sharded solves: every rank
 * holds all coordinates, so every rank draws the same sample and the same window): rank r runs ..._rows_dev on its
 * rows -> keys of the distances inside the window (device buffer of *cap entries from ..._window_cap; 0 = problem too
 * small, use wotb_cost_median_dev), their count and the count of distances below the window; the caller sums `below`
 * and concatenates the keys across ranks (NCCL all-reduce / all-gather) and every rank calls ..._finish_dev on the union.
 * *ok == 0: the window missed the middle ranks (ties, overflow): fall back to wotb_cost_median_dev. */
int wotb_cost_median_window_cap(int64_t I, int64_t J, int64_t *cap);
int wotb_cost_median_window_rows_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                                     const double *scale, int64_t row_lo, int64_t row_hi, uint64_t *keys, int64_t cap,
                                     uint32_t *count, uint64_t *below);
int wotb_cost_median_window_finish_dev(wotb_ctx *ctx, int64_t I, int64_t J, const uint64_t *keys, int64_t count,
                                       uint64_t below, double *median_host, int32_t *ok);
/* C[i*ldc + j] = dist_ij / median, rounded once to dtype (WOTB_F32 for the solver, WOTB_F64 for
 * callers of compute_default_cost_matrix). median == 1.0 gives raw distances. */
int wotb_cost_matrix_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                         const double *scale, double median, void *C, int64_t ldc, int32_t dtype);
/* float64 [I,J] (ld_src) -> float32 [I,J] (ld_dst, padded columns zeroed); for caller-supplied costs. */
int wotb_cost_to_f32_dev(wotb_ctx *ctx, const double *src, int64_t ld_src, int64_t I, int64_t J, float *dst,
                         int64_t ld_dst);

/* ---- solvers: replace the `solver(**params)` callable, optimal_transport.py:30 -----------------
 * C [I, ldc] float32 device (ldc % 4 == 0, 16-byte aligned), G [I] float64 device.
 * Outputs (device, float64): f [I] = u + eps log a, g [J] = v + eps log b, rowsum [I] = row sums
 * of the returned coupling (what optimal_transport.py:27 and ot_model.py:319 take from tmap).
 * The coupling itself is tmap_ij = exp((f_i + g_j - C_ij) / info->eps_final) * info->out_scale;
 * materialise it with wotb_coupling_dev.  Both reference solvers are selected by params->solver. */
int wotb_sinkhorn_stored_dev(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *G,
                             const wotb_params *params, double *f, double *g, double *rowsum, wotb_info *info);

/* Online variant: never materialises C or K.  x0 [I,d], x1 [J,d] float64 device coordinates
 * (already multiplied by `scale` if any), median from wotb_cost_median_dev. */
int wotb_sinkhorn_online_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                             double median, const double *G, const wotb_params *params, double *f, double *g,
                             double *rowsum, wotb_info *info);

/* tmap[i*ldo + j] = exp((f_i + g_j - C_ij)/eps) * out_scale      (optimal_transport.py:153,164).
 * rowsum may be NULL. out dtype WOTB_F32 or WOTB_F64. `out` may be device memory or pinned/mapped
 * host memory reachable from the device. */
int wotb_coupling_dev(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *f,
                      const double *g, double eps, double out_scale, void *out, int64_t ldo, int32_t dtype,
                      double *rowsum);
int wotb_coupling_online_dev(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d,
                             double median, const double *f, const double *g, double eps, double out_scale,
                             void *out, int64_t ldo, int32_t dtype, double *rowsum);

/* ---- host-buffer entry points: what a reference-side binding calls -----------------------------
 * wotb_transport_map_from_cost_host replaces wot.ot.compute_transport_matrix(solver, C=..., G=...,
 * growth_iters=..., **ot_config) (optimal_transport.py:10-33, called at ot_model.py:318).
 *   C_host [I,J] float64, G_host [I] float64
 *   tmap_host [I,J] (dtype out_dtype) or NULL; learned_growth_host [(growth_iters+1), I]: rows
 *   0..growth_iters-1 are the G used by each growth iteration (optimal_transport.py:24-29), the last
 *   row is tmap.sum(axis=1) (ot_model.py:319); f_host/g_host may be NULL; infos [growth_iters]. */
int wotb_transport_map_from_cost_host(wotb_ctx *ctx, const double *C_host, int64_t I, int64_t J,
                                      const double *G_host, const wotb_params *params, int32_t growth_iters,
                                      void *tmap_host, int32_t out_dtype, double *learned_growth_host,
                                      double *f_host, double *g_host, wotb_info *infos);

/* wotb_transport_map_from_coords_host replaces ot_model.py:307-319: default cost (scaled
 * coordinates, squared Euclidean, / median) followed by the growth loop.  scale_host may be NULL.
 * median_out may be NULL.  params->kernel selects stored-K or online-K. */
int wotb_transport_map_from_coords_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host,
                                        int64_t J, int32_t d, const double *scale_host, const double *G_host,
                                        const wotb_params *params, int32_t growth_iters, void *tmap_host,
                                        int32_t out_dtype, double *learned_growth_host, double *f_host,
                                        double *g_host, double *median_out, wotb_info *infos);

/* replaces OTModel.compute_default_cost_matrix for callers that want the matrix itself (float64). */
int wotb_default_cost_matrix_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J,
                                  int32_t d, const double *scale_host, double *C_host, double *median_out);

/* ---- row-sharded online solve: one huge day-pair across GPUs (BASELINE.json configs[3]) ----------
 * One process per GPU; every rank calls the same sequence.  x0, x1, G are the FULL arrays on every rank
 * (device, float64); rank `shard` of `n_shards` computes a contiguous slice of 128-row tiles.  The solver
 * state is replicated; one float64 vector per iteration is summed across ranks by the CALLER (NCCL
 * all-reduce on the context's stream) between the steps, through `exchange` (device, 2 I + J doubles):
 *
 *   per batch:      step BEGIN_A, all-reduce exchange[0:I], step BEGIN_B
 *   per iteration:  step ROW, step COL_PARTIAL, all-reduce exchange[0 : 2I + J], step COL_FINISH
 *                   (one exchange: gathered a | their row sums | partial column sums; the column pass over a
 *                   rank's own rows needs only that rank's a)
 *   per batch end:  step GAP_ROWS, all-reduce exchange[0:I], step CHECK, then wotb_online_state() (syncs)
 *   when done:      step FINAL_ROWS, all-reduce exchange[0:I]  -> row sums of the coupling
 *
 * Steps whose work is not due (batch finished early, tau stop, solver done) are no-ops on the device, so
 * every rank issues the same collectives in the same order.  f [I], g [J] receive the potentials. */
enum wotb_online_op {
    WOTB_OP_BEGIN_A = 0, WOTB_OP_BEGIN_B = 1, WOTB_OP_ROW = 2, WOTB_OP_COL_PARTIAL = 3, WOTB_OP_COL_FINISH = 4,
    WOTB_OP_GAP_ROWS = 5, WOTB_OP_CHECK = 6, WOTB_OP_FINAL_ROWS = 7
};
int wotb_online_open(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int32_t d, double median,
                     const double *G, const wotb_params *params, int32_t shard, int32_t n_shards, double *f, double *g,
                     void **solve);
int wotb_online_step(void *solve, int32_t op, double *exchange);
int wotb_online_state(void *solve, wotb_info *info, int32_t *done);   /* synchronises the context's stream */
int wotb_online_done(void *solve, int32_t *done);                      /* non-blocking: the done flag k_check raises in
                                                                          mapped page-locked memory (keep batches in flight) */
int wotb_online_rows(void *solve, int64_t *row_lo, int64_t *row_hi);
void wotb_online_close(void *solve);

/* ---- the same solve exchanging over PEER MEMORY (NVLink / NVSwitch) instead of the caller's all-reduce ---------
 * The per-iteration exchange of the row-sharded solve (the reference has no counterpart: optimal_transport.py:133-134
 * run on one host) moves into the kernels: the finishing code of the row half-step stores (a_i, row sum) of the
 * rank's rows into EVERY rank's exchange buffer while the pass is still running, the column pass does the same with
 * its partial column sums, one warp exchanges flags and one kernel adds the partial sums in rank order (identical bits
 * on every rank) and applies the b update.  No collective library call, no `exchange` traffic between the steps:
 *
 *   once:           wotb_online_open; wotb_online_peer_bytes; wotb_peer_alloc (own buffer + 64-byte IPC handle);
 *                   exchange the handles (any host transport); wotb_peer_open for every other rank;
 *                   wotb_online_attach_peers; a host barrier over all ranks
 *   per batch:      step BEGIN_A, step BEGIN_B            (no all-reduce)
 *   per iteration:  step ROW, step COL_PARTIAL, step COL_FINISH   (no all-reduce)
 *   per batch end:  step CHECK, wotb_online_state
 *   when done:      step FINAL_ROWS  -> exchange[0:I] holds the coupling's row sums on every rank
 *
 * A rank whose peers never arrive traps after 20 s instead of hanging.  Up to 8 ranks.  Ranks inside one process may
 * pass plain device pointers to wotb_online_attach_peers instead of IPC mappings. */
int wotb_peer_alloc(wotb_ctx *ctx, int64_t bytes, void **ptr, void *ipc_handle_64 /* may be NULL */);
int wotb_peer_open(wotb_ctx *ctx, const void *ipc_handle_64, void **ptr);
int wotb_peer_close(wotb_ctx *ctx, void *ptr);   /* a mapping obtained from wotb_peer_open */
int wotb_peer_free(wotb_ctx *ctx, void *ptr);    /* a buffer obtained from wotb_peer_alloc */
int wotb_online_peer_bytes(void *solve, int32_t world, int64_t *bytes);
int wotb_online_attach_peers(void *solve, int32_t world, void *const *bufs /* [world], own buffer at [shard] */);

/* ---- a coupling applied to populations without materialising it ------------------
 * Replaces the products of TransportMapModel.push_forward / pull_back, wot/tmap/transport_map_model.py:290
 * (p @ tmap.X) and :356 (tmap.X @ p.T), where p stacks ALL populations (np.vstack, :285, :351): the coupling of a
 * finished solve is fully described by the local-PCA coordinates, the median, the potentials f, g, eps_final and
 * out_scale (wotb_info), so
 *   forward = 1:  out[k, j] = sum_i p[k, i] tmap[i, j]     p [n_pop, I] -> out [n_pop, J]
 *   forward = 0:  out[k, i] = sum_j tmap[i, j] p[k, j]     p [n_pop, J] -> out [n_pop, I]
 * is computed tile by tile in float64: one exponential per coupling entry, then one FP64 FMA per population (8
 * populations per sweep over the coupling); partial sums are added in a fixed order (deterministic).  Populations
 * may have any sign.  Host pointers, float64; scale_host as in wotb_transport_map_from_coords_host. */
int wotb_coupling_apply_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int32_t d,
                             const double *scale_host, double median, const double *f_host, const double *g_host,
                             double eps_final, double out_scale, int32_t forward, const double *p_host, int32_t n_pop,
                             double *out_host);

/* ---- sampling cell pairs from a coupling ------------------------------------------
 * Replaces the draw of interpolate_with_ot, wot/ot/util.py:140-146 (np.random.choice over the flattened
 * p = tmap / colsum^(1 - frac)): for sample s the caller gives the row rows[s] it fell into and the mass
 * targets[s] that remains inside that row (both found on the row masses sum_j tmap[i, j] w[j], which
 * wotb_coupling_apply_host(forward = 0, p = w) returns); cols[s] receives the first column j whose running sum
 * sum_{j' <= j} tmap[rows[s], j'] w[j'] exceeds targets[s] (float64, column order, like the flattened cumsum). */
int wotb_coupling_sample_host(wotb_ctx *ctx, const double *x0_host, int64_t I, const double *x1_host, int64_t J, int32_t d,
                              const double *scale_host, double median, const double *f_host, const double *g_host,
                              double eps_final, double out_scale, const double *w_host, const int64_t *rows_host,
                              const double *targets_host, int64_t n_samples, int64_t *cols_host);

/* Measurement hook for bench.py: average device time (ms, CUDA events on the context's stream) of one
 * row-pass and one column-pass launch of the stored-K matvec kernels on an I x J kernel matrix. */
int wotb_bench_matvec_dev(wotb_ctx *ctx, int64_t I, int64_t J, int32_t reps, double *ms_row, double *ms_col,
                          double *ms_fused /* one fused iteration (K read once), -1 if the shape is unsupported */);

/* Kernel-level hook for tests and bench.py: one online-kernel pass (the K.(b dy) half of optimal_transport.py:133
 * without materialising K),  sums[i] = sum_j exp2(off_out[i] + off_in[j] + scale^2 <x_out_i, x_in_j>),  all device
 * pointers, float64.  impl 0: SIMT FP32 kernel, impl 1 / 2: tcgen05 kernel with 8 / 16 epilogue warps (fp16 split operands,
 * cross term + offsets accumulated in TMEM; d <= 46).
 * ms_per_pass (may be NULL) = average device time of `reps` launches after one warm-up (CUDA events on the
 * context's stream). */
int wotb_online_rowsums_dev(wotb_ctx *ctx, const double *x_out, int64_t n_out, const double *x_in, int64_t n_in, int32_t d,
                            double scale, const double *off_out, const double *off_in, int32_t impl, int32_t reps,
                            double *sums, double *ms_per_pass);

/* Measured MUFU.EX2 throughput of the device (ex2 per second, all SMs): the roofline denominator of the online
 * kernels. */
int wotb_bench_mufu_dev(wotb_ctx *ctx, double *ex2_per_s);

/* Page-locked host memory for coupling outputs (cudaHostAlloc): a coupling written into it leaves the
 * device at PCIe speed; pageable destinations are served through an internal bounce buffer. */
int wotb_pinned_alloc(size_t bytes, void **out);
void wotb_pinned_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* WOT_B200_H */
