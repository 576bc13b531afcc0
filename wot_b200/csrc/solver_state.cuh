// Device-resident state machine of one Sinkhorn solve.
//
// The reference drives its loop from Python (optimal_transport.py:116-164 / :204-232): every batch
// it decides on the host whether to continue, absorb, change epsilon or stop.  Here that decision
// lives on the GPU.  The host only replays one fixed launch sequence
//     [build K if needed] (row matvec, column matvec) x B  [row sums for the gap]  [check]
// and every kernel reads this block to decide whether it has work; the check kernel advances the
// machine.  No host synchronisation sits between iterations, batches or epsilon stages.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

struct SolveCtrl {
    // ---- constant for the solve ---------------------------------------------------------------
    int I, J;
    int solver;      // wotb_solver
    int batch_size;  // final-stage iterations between duality-gap checks
    int warm;        // fixed_iters: tau is not None
    int scaling_iter, extra_iter, inner_iter_max;
    double lambda1, lambda2, tau, tolerance, epsilon, epsilon0;
    double eps_sched[WOTB_N_STAGES];  // epsilon_i of every stage, same recurrence as :113,:120
    double max_iter;
    double q;   // np.average(G), :108
    double lq;  // log q
    // ---- advanced by the check kernel ---------------------------------------------------------
    int stage;  // duality_gap: epsilon stage 0..5; fixed_iters: epsilon level
    double eps, alpha1, alpha2, inv_l1e, inv_l2e;
    int cur;          // which of a[2] / b[2] holds the current scalings
    long long iter;   // current_iter
    int batch_iters;  // iterations the current batch runs before the next check
    int batch_done;   // iterations of the current batch already executed
    int stop;         // bit0: tau exceeded (:137), bit1: max_iter reached (:143)
    int tau_check;    // 0 in the extra_iter phase of fixed_iters (:230-232 has no stabilisation)
    int need_build;   // K must be rebuilt from (u, v, C, eps) before the next iteration
    int done;
    int status;  // wotb_status
    int batches[WOTB_N_STAGES];
    int tau_count;
    double gap, primal, dual, sumK0;
    int phase;  // fixed_iters: 0 scaling loop, 1 extra loop
    int since;  // iterations_since_epsilon_adjusted
    int scaling_done;
    // ---- scratch written by the matvec kernels ------------------------------------------------
    unsigned long long maxabs;  // bit pattern of max(|a|, |b|) over the current iteration
    unsigned int col_tiles_done;
    unsigned int grid_bar;  // arrival counter of the fused kernel's grid barrier (zeroed by k_build)
    unsigned long long seq;  // check kernels executed so far
    double eps_final, out_scale;
    // ---- lazy duality-gap check (see k_check) --------------------------------------------------
    int snap_valid;       // a final-stage state is waiting for its row sums
    int rowsum_ready;     // V.rowsum holds the coupling row sums of the returned state
    long long snap_iter;  // current_iter of the snapshot
    // ---- online kernel only --------------------------------------------------------------------
    double inv_median;  // 1 / np.median(raw squared distances), ot_model.py:252
    double c1, c2;      // log2(e)/eps and log2(e)/(eps*median): exponents are formed in base 2
    double log2_I, log2_J;  // constants of the offsets Pd / Qd
};

// ---- row-sharded solves exchanging over peer memory (NVLink / NVSwitch), see online_solve.cuh ------------------
// Every rank owns one exchange buffer of identical layout that all peers have mapped (cudaIpc or plain pointers
// inside one process).  The pass kernels' finishing code STORES a rank's results straight into every peer's
// buffer while the pass is still running (a-slices and their row sums from the row half-step, partial column sums
// from the column half-step), a one-warp kernel exchanges flags, and the finishing kernel adds the partial sums in
// rank order -- the same bits on every rank, so the replicated state machine stays in lockstep.  Two copies of every
// region, selected by the parity of `seq` (exchanges completed so far): a rank can be at most one exchange ahead of a
// peer, so what it writes never lands in a copy the peer still reads.
constexpr int kMaxPeers = 8;
struct PeerX {
    int rank, world;
    unsigned long long seq;       // exchanges completed; identical on every rank at every exchange
    unsigned int ticket;          // last-block detection of the importing kernels
    unsigned char *buf[kMaxPeers];  // exchange buffers of all ranks (own included), byte pointers
    long long off_flags;          // unsigned long long [2][kMaxPeers]: flag of rank w = seq + 1 once w's data has landed
    long long off_a[2], off_s[2];  // double [I] each: gathered a / gathered row sums (also used for other row vectors)
    long long off_t[2];           // double [world][ld_t]: every rank's partial column sums
    long long ld_t;
};

// Pointers into the per-solve vector workspace (all device memory, fixed for the solve).
struct SolveVecs {
    const double *p;  // G, row masses
    double *u, *v;    // absorbed dual potentials
    double *a[2], *b[2];
    double *lu, *lv;  // -u/(lambda1+eps), -v/(lambda2+eps): log of the damping factors, constant between absorptions
    double *lp;       // log p (row masses), constant for the solve
    double *s, *t;    // K (b dy) and K^T (a dx) of the last matvecs
    double *r, *c;    // row / column sums of R = a K b at the last gap check
    double *f, *g;    // outputs
    double *sfirst;   // K (b dy) of the first iteration of the current batch: the row sums of the snapshot
    double *fs, *gs, *cs, *as;  // snapshot of the state at the last final-stage batch end (f, g, column sums, a)
    double *rowsum;   // coupling row sums written by the device when it finishes from a snapshot (may be NULL)
    float *w, *z;     // fp32 copies of b*dy and a*dx fed to the matvecs (w padded to ld with zeros)
    double *colpart;  // [n_row_blocks, ldp] column partial sums
    unsigned int *tile_counters;
    double *sumK0_part;
    int n_sumK0_part;
    long long ldp;
    // ---- online kernel only: exponent offsets, log2 domain (see online.cu) ----------------------
    int online;
    const double *nx, *ny;  // raw squared norms of the coordinates
    double *Ps, *Qs;         // c1 u_i - c2 |x_i|^2 and c1 v_j - c2 |y_j|^2   (change on absorption / new eps)
    double *Pd, *Qd;         // Ps + log2(a_i / I) and Qs + log2(b_j / J)      (change every half-step)
    long long n_pad_i, n_pad_j;  // padded lengths; padding holds -inf so padded entries add exp2(-inf) = 0
    // ---- tcgen05 online kernel only: the B-role operand rows whose offset slots follow Pd / Qd (online_tc.cuh) ----
    __half *tcXB, *tcYB;
    int tc_kseg;
    int tc_nseg;  // 3: fp16 hi/lo split (default); 6: precise mode, grid-aligned leading limb + two more limbs
    // ---- row-sharded solve over peer memory only (NULL otherwise) ---------------------------------------------------
    PeerX *peer;
};
