// Local PCA on the GPU: the producer of the hot path's inputs.
//
// Replaces compute_pca, /root/reference/wot/ot/util.py:240-255:
//     x = vstack(m1, m2) - gene means;  pca = sklearn PCA(k, random_state=58951).fit(x.T);  comp = pca.components_.T
// for the case in which scikit-learn's svd_solver='auto' picks the randomized solver (always, at atlas shapes:
// max(shape) > 500 and k < 0.8 min(shape)).  scikit-learn is not vendored in the reference; its published
// algorithm (randomized_svd / randomized_range_finder, Halko et al. 2011) is restated in oracle/pca_oracle.py and
// pinned there against the installed scikit-learn.  The same arithmetic, float64 throughout:
//
//     A = x - per-cell mean over genes            (PCA.fit centres every feature of x.T)          [N cells, G genes]
//     M = A if G < N else A^T                      ("transpose='auto'")
//     Q = Q0                                       (numpy RandomState(58951).normal, made by the caller)
//     n_iter times:  Q = orth(M Q);  Q = orth(M^T Q)
//     Q = orth(M Q);  B = Q^T M;  B = Uhat diag(s) Vt;  components = (Q Uhat)[:, :k] or Vt[:k]^T
//
// orth() is two rounds of Cholesky-QR (Gram matrix, 40 x 40 Cholesky and triangular inverse in one CTA, rows
// times R^-1): scikit-learn normalises with a pivoted LU between power iterations, which spans the same space.
// The small SVD goes through the s x s Gram matrix of B (cyclic Jacobi on the host, 40 x 40).  The two big
// products stream A once each (N G 8 bytes, 300 MB at atlas shapes) with 4 x 4 register tiles on the FP64 pipe;
// everything reduces in a fixed order, so results are deterministic.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace wotb {

constexpr int kPcaMaxS = 64;   // k + n_oversamples <= 64
constexpr int kPcaKc = 32;     // reduction chunk staged in shared memory
constexpr int kPcaRows = 64;   // output rows (cells or genes) per CTA
constexpr int kPcaSlab = 512;  // cells per CTA of the A^T Y product
constexpr int kPcaGramRows = 128;  // rows per CTA of the Gram product (enough CTAs to fill the GPU for the tall side)

// ---- centring -----------------------------------------------------------------------------------------------------
// column sums of X[n, g] over a slab of rows -> part[slab][g]
__global__ void k_pca_colsum(const double *__restrict__ X, long long N, int G, int rows_per_slab, double *__restrict__ part) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const long long r0 = (long long)blockIdx.y * rows_per_slab;
    const long long r1 = r0 + rows_per_slab < N ? r0 + rows_per_slab : N;
    double s = 0.0;
    for (long long r = r0; r < r1; ++r) s += X[r * G + g];
    part[(long long)blockIdx.y * G + g] = s;
}

// out[c] = scale * sum over slabs of part[slab][c], slabs added in order
__global__ void k_pca_reduce(const double *__restrict__ part, int n_slabs, long long n, double scale, double *__restrict__ out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int k = 0; k < n_slabs; ++k) s += part[(long long)k * n + c];
    out[c] = s * scale;
}

// one warp per cell: m = mean over genes of (X - mu);  X <- X - mu - m      (util.py:245-246 + PCA.fit's centring)
__global__ void k_pca_center(double *__restrict__ X, long long N, int G, const double *__restrict__ mu,
                             double *__restrict__ cell_means) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    double *x = X + row * G;
    double s = 0.0;
    for (int g = lane; g < G; g += 32) s += x[g] - mu[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double m = s / (double)G;
    if (lane == 0) cell_means[row] = m;  // sklearn's PCA.mean_ (one entry per feature of x.T, i.e. per cell)
    for (int g = lane; g < G; g += 32) x[g] = x[g] - mu[g] - m;
}

// ---- Y[N, sp] = A[N, G] Q[G, sp] -----------------------------------------------------------------------------------
// CTA: 64 cells x all sp columns, threads (sp / 4, 16), thread tile 4 cells x 4 columns.
__global__ void __launch_bounds__(256) k_pca_aq(const double *__restrict__ A, long long N, int G, const double *__restrict__ Q,
                                                int sp, double *__restrict__ Y) {
    __shared__ double As[kPcaRows][kPcaKc + 1];            // [cell][k], padded: conflict-free stores and loads
    __shared__ __align__(16) double Qs[kPcaKc][kPcaMaxS];  // [k][column]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * blockDim.x + tx, nthr = blockDim.x * blockDim.y;
    const long long row0 = (long long)blockIdx.x * kPcaRows;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int k0 = 0; k0 < G; k0 += kPcaKc) {
        for (int e = tid; e < kPcaRows * kPcaKc; e += nthr) {  // coalesced along genes
            const int r = e / kPcaKc, k = e % kPcaKc;
            const long long row = row0 + r;
            As[r][k] = (row < N && k0 + k < G) ? A[row * G + k0 + k] : 0.0;
        }
        for (int e = tid; e < kPcaKc * sp; e += nthr) {
            const int k = e / sp, c = e % sp;
            Qs[k][c] = (k0 + k < G) ? Q[(long long)(k0 + k) * sp + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kPcaKc; ++k) {
            const double2 q01 = *reinterpret_cast<const double2 *>(&Qs[k][tx * 4]);
            const double2 q23 = *reinterpret_cast<const double2 *>(&Qs[k][tx * 4 + 2]);
            const double a[4] = {As[ty * 4][k], As[ty * 4 + 1][k], As[ty * 4 + 2][k], As[ty * 4 + 3][k]};
            const double q[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], q[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long row = row0 + ty * 4 + r;
        if (row < N) {
#pragma unroll
            for (int c = 0; c < 4; ++c) Y[row * sp + tx * 4 + c] = acc[r][c];
        }
    }
}

// ---- Zpart[slab][G, sp] = A[slab rows, G]^T Y[slab rows, sp] -------------------------------------------------------
// CTA: 64 genes x all sp columns over one slab of cells, thread tile 4 genes x 4 columns.
__global__ void __launch_bounds__(256) k_pca_aty(const double *__restrict__ A, long long N, int G, const double *__restrict__ Y,
                                                 int sp, double *__restrict__ Zpart) {
    __shared__ __align__(16) double As[kPcaKc][kPcaRows];  // [cell][gene]
    __shared__ __align__(16) double Ys[kPcaKc][kPcaMaxS];  // [cell][column]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * blockDim.x + tx, nthr = blockDim.x * blockDim.y;
    const int g0 = blockIdx.x * kPcaRows;
    const long long r_lo = (long long)blockIdx.y * kPcaSlab;
    const long long r_hi = r_lo + kPcaSlab < N ? r_lo + kPcaSlab : N;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (long long r0 = r_lo; r0 < r_hi; r0 += kPcaKc) {
        for (int e = tid; e < kPcaKc * kPcaRows; e += nthr) {  // coalesced along genes
            const int r = e / kPcaRows, g = e % kPcaRows;
            As[r][g] = (r0 + r < r_hi && g0 + g < G) ? A[(r0 + r) * G + g0 + g] : 0.0;
        }
        for (int e = tid; e < kPcaKc * sp; e += nthr) {
            const int r = e / sp, c = e % sp;
            Ys[r][c] = (r0 + r < r_hi) ? Y[(r0 + r) * sp + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kPcaKc; ++k) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&As[k][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2 *>(&As[k][ty * 4 + 2]);
            const double2 q01 = *reinterpret_cast<const double2 *>(&Ys[k][tx * 4]);
            const double2 q23 = *reinterpret_cast<const double2 *>(&Ys[k][tx * 4 + 2]);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y}, q[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], q[c], acc[r][c]);
        }
        __syncthreads();
    }
    double *out = Zpart + (long long)blockIdx.y * G * sp;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int g = g0 + ty * 4 + r;
        if (g < G) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out[(long long)g * sp + tx * 4 + c] = acc[r][c];
        }
    }
}

// ---- Gram matrix of a tall matrix: Gpart[slab][sp, sp] = T[slab]^T T[slab] ----------------------------------------
__global__ void __launch_bounds__(256) k_pca_gram(const double *__restrict__ T, long long rows, int sp, int rows_per_slab,
                                                  double *__restrict__ Gpart) {
    __shared__ __align__(16) double Ts[kPcaKc][kPcaMaxS];
    const int tx = threadIdx.x, ty = threadIdx.y;  // (sp / 4, sp / 4): thread tile 4 x 4 of the Gram matrix
    const int tid = ty * blockDim.x + tx, nthr = blockDim.x * blockDim.y;
    const long long r_lo = (long long)blockIdx.x * rows_per_slab;
    const long long r_hi = r_lo + rows_per_slab < rows ? r_lo + rows_per_slab : rows;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (long long r0 = r_lo; r0 < r_hi; r0 += kPcaKc) {
        for (int e = tid; e < kPcaKc * sp; e += nthr) {
            const int r = e / sp, c = e % sp;
            Ts[r][c] = (r0 + r < r_hi) ? T[(r0 + r) * sp + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kPcaKc; ++k) {
            double a[4], q[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = Ts[k][ty * 4 + r], q[r] = Ts[k][tx * 4 + r];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], q[c], acc[r][c]);
        }
        __syncthreads();
    }
    double *out = Gpart + (long long)blockIdx.x * sp * sp;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[(ty * 4 + r) * sp + tx * 4 + c] = acc[r][c];
}

// ---- one CTA: Gram (s x s, leading dimension sp) = R^T R;  W = R^-1 (upper triangular), padded with zeros ----------
// 1024 threads as a 32 x 32 grid, thread (ty, tx) owns the 2 x 2 block of rows {ty, ty + 32} x columns {tx, tx + 32}
// in the rank-1 updates; the triangular inverse runs one warp per column, dot products by warp shuffle.
__global__ void __launch_bounds__(1024) k_pca_chol_inv(const double *__restrict__ Gram, int s, int sp, double *__restrict__ W,
                                                       int *__restrict__ fail) {
    __shared__ double L[kPcaMaxS][kPcaMaxS + 1];  // lower Cholesky factor, Gram = L L^T, R = L^T
    const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
    for (int e = t; e < kPcaMaxS * kPcaMaxS; e += 1024) {
        const int r = e / kPcaMaxS, c = e % kPcaMaxS;
        L[r][c] = (r < s && c < s) ? Gram[r * sp + c] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < s; ++j) {  // right-looking Cholesky, column j
        if (t == 0) {
            const double d = L[j][j];
            if (!(d > 0.0)) *fail = 1;
            L[j][j] = sqrt(d > 0.0 ? d : 1.0);
        }
        __syncthreads();
        if (t > j && t < s) L[t][j] /= L[j][j];
        __syncthreads();
#pragma unroll
        for (int dr = 0; dr < 2; ++dr)
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int r = ty + 32 * dr, c = tx + 32 * dc;
                if (r > j && r < s && c > j && c <= r) L[r][c] -= L[r][j] * L[c][j];
            }
        __syncthreads();
    }
    // W = R^-1 = (L^-1)^T.  Warp w solves L x = e_col for col = w, w + 32 by forward substitution: x_i needs the dot
    // product of row i of L with the x found so far, lanes hold x_lane and x_{lane + 32}.
    for (int e = t; e < sp * sp; e += 1024) W[e] = 0.0;
    __syncthreads();
    for (int col = ty; col < s; col += 32) {
        double x0 = 0.0, x1 = 0.0;  // x[tx], x[tx + 32]
        for (int i = col; i < s; ++i) {
            double part = 0.0;
            if (tx >= col && tx < i) part += L[i][tx] * x0;
            if (tx + 32 >= col && tx + 32 < i) part += L[i][tx + 32] * x1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            const double xi = ((i == col ? 1.0 : 0.0) - part) / L[i][i];
            if (tx == (i & 31)) {
                if (i < 32) x0 = xi; else x1 = xi;
            }
            if (tx == 0) W[col * sp + i] = xi;
        }
    }
}

// ---- Out[rows, n_out] = T[rows, sp] W  (W [sp, ldw] row-major; Out may alias T when ld_out == sp) -----------------
// CTA: 16 rows staged in shared memory; thread (r, g) of 16 x 16 computes outputs g, g + 16, ... of row r.
__global__ void __launch_bounds__(256) k_pca_apply(const double *__restrict__ T, long long rows, int sp, const double *__restrict__ W,
                                                   int ldw, int n_out, int ld_out, double *__restrict__ Out) {
    __shared__ double Ws[kPcaMaxS * kPcaMaxS];
    __shared__ double Ts[16][kPcaMaxS + 1];
    const long long row0 = (long long)blockIdx.x * 16;
    for (int e = threadIdx.x; e < sp * ldw; e += 256) Ws[e] = W[e];
    for (int e = threadIdx.x; e < 16 * sp; e += 256) {
        const int r = e / sp, c = e % sp;
        Ts[r][c] = row0 + r < rows ? T[(row0 + r) * sp + c] : 0.0;
    }
    __syncthreads();  // every input of this CTA's rows is in shared memory: writing Out over T is safe from here on
    const int r = threadIdx.x >> 4, g = threadIdx.x & 15;
    if (row0 + r >= rows) return;
    for (int o = g; o < n_out; o += 16) {
        double v = 0.0;
        for (int c = 0; c < sp; ++c) v = fma(Ts[r][c], Ws[c * ldw + o], v);
        Out[(row0 + r) * ld_out + o] = v;
    }
}

// ---- host: symmetric eigendecomposition of an s x s matrix, cyclic Jacobi -------------------------------------------
static void jacobi_eigh(std::vector<double> &a, int s, std::vector<double> &vec, std::vector<double> &val) {
    vec.assign((size_t)s * s, 0.0);
    for (int i = 0; i < s; ++i) vec[(size_t)i * s + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < s; ++p) {
            diag += a[(size_t)p * s + p] * a[(size_t)p * s + p];
            for (int q = p + 1; q < s; ++q) off += a[(size_t)p * s + q] * a[(size_t)p * s + q];
        }
        if (off <= 1e-34 * diag) break;
        for (int p = 0; p < s - 1; ++p) {
            for (int q = p + 1; q < s; ++q) {
                const double apq = a[(size_t)p * s + q];
                if (apq == 0.0) continue;
                const double theta = (a[(size_t)q * s + q] - a[(size_t)p * s + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < s; ++k) {  // columns p, q
                    const double akp = a[(size_t)k * s + p], akq = a[(size_t)k * s + q];
                    a[(size_t)k * s + p] = c * akp - sn * akq;
                    a[(size_t)k * s + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < s; ++k) {  // rows p, q
                    const double apk = a[(size_t)p * s + k], aqk = a[(size_t)q * s + k];
                    a[(size_t)p * s + k] = c * apk - sn * aqk;
                    a[(size_t)q * s + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < s; ++k) {
                    const double vkp = vec[(size_t)k * s + p], vkq = vec[(size_t)k * s + q];
                    vec[(size_t)k * s + p] = c * vkp - sn * vkq;
                    vec[(size_t)k * s + q] = sn * vkp + c * vkq;
                }
            }
        }
    }
    val.resize(s);
    for (int i = 0; i < s; ++i) val[i] = a[(size_t)i * s + i];
}

// ---- driver --------------------------------------------------------------------------------------------------------
struct PcaWork {
    double *A, *mu, *Q, *Y, *Z, *part, *gram, *W;
    int *fail;
    long long N;
    int G, s, sp;
    cudaStream_t st;
};

static void pca_gram(const PcaWork &w, const double *T, long long rows) {
    const int rps = kPcaGramRows;
    const int slabs = (int)cdiv(rows, rps);
    k_pca_gram<<<slabs, dim3(w.sp / 4, w.sp / 4), 0, w.st>>>(T, rows, w.sp, rps, w.part);
    k_pca_reduce<<<(unsigned)cdiv((long long)w.sp * w.sp, 256), 256, 0, w.st>>>(w.part, slabs, (long long)w.sp * w.sp, 1.0, w.gram);
}

// two rounds of Cholesky-QR, in place
static void pca_orth(const PcaWork &w, double *T, long long rows) {
    for (int round = 0; round < 2; ++round) {
        pca_gram(w, T, rows);
        k_pca_chol_inv<<<1, 1024, 0, w.st>>>(w.gram, w.s, w.sp, w.W, w.fail);
        k_pca_apply<<<(unsigned)cdiv(rows, 16), 256, 0, w.st>>>(T, rows, w.sp, w.W, w.sp, w.sp, w.sp, T);
    }
}

static void pca_aq(const PcaWork &w, const double *Qin, double *Yout) {  // [N, sp] = A [G, sp]
    k_pca_aq<<<(unsigned)cdiv(w.N, kPcaRows), dim3(w.sp / 4, 16), 0, w.st>>>(w.A, w.N, w.G, Qin, w.sp, Yout);
}

static void pca_aty(const PcaWork &w, const double *Yin, double *Zout) {  // [G, sp] = A^T [N, sp]
    const int slabs = (int)cdiv(w.N, kPcaSlab);
    k_pca_aty<<<dim3((unsigned)cdiv(w.G, kPcaRows), slabs), dim3(w.sp / 4, 16), 0, w.st>>>(w.A, w.N, w.G, Yin, w.sp, w.part);
    k_pca_reduce<<<(unsigned)cdiv((long long)w.G * w.sp, 256), 256, 0, w.st>>>(w.part, slabs, (long long)w.G * w.sp, 1.0, Zout);
}

int pca_host(wotb_ctx *ctx, const double *m1, int64_t n1, const double *m2, int64_t n2, int64_t genes, int k,
             const double *q0, int size, int n_iter, double *comp_host, double *sv_host, double *gene_means_host,
             double *cell_means_host, double *gpu_ms) {
    WOTB_REQUIRE(ctx && m1 && m2 && q0 && comp_host && sv_host, "NULL argument");
    const long long N = n1 + n2;
    const int G = (int)genes;
    WOTB_REQUIRE(n1 >= 1 && n2 >= 1 && genes >= 1 && genes < (1 << 30), "empty input");
    WOTB_REQUIRE(k >= 1 && size >= k && size <= kPcaMaxS && size <= N && size <= G, "need k <= size <= min(64, cells, genes)");
    WOTB_REQUIRE(n_iter >= 0, "n_iter must be >= 0");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool transpose = G < N;  // randomized_svd(transpose='auto'): work on the matrix with more rows than columns
    const int sp = (int)round_up(size, 4);
    const long long small = transpose ? G : N, tall = transpose ? N : G;
    const int slabs_aty = (int)cdiv(N, kPcaSlab), slabs_gram = (int)cdiv(tall, kPcaGramRows), slabs_col = (int)cdiv(N, 256);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    const size_t o_A = take((size_t)N * G * 8), o_mu = take((size_t)G * 8), o_cm = take((size_t)N * 8);
    const size_t o_Y = take((size_t)N * sp * 8), o_Z = take((size_t)G * sp * 8);
    size_t part_bytes = (size_t)slabs_aty * G * sp * 8;
    part_bytes = std::max(part_bytes, (size_t)std::max(slabs_gram, (int)cdiv(std::max(N, (long long)G), kPcaGramRows)) * sp * sp * 8);
    part_bytes = std::max(part_bytes, (size_t)slabs_col * G * 8);
    const size_t o_part = take(part_bytes), o_gram = take((size_t)sp * sp * 8), o_W = take((size_t)kPcaMaxS * kPcaMaxS * 8);
    const size_t o_fail = take(256);
    WOTB_TRY(ctx->hC.reserve(off));  // the PCA runs before the cost: the fp32 cost buffer is free
    char *base = ctx->hC.as<char>();
    PcaWork w;
    w.A = (double *)(base + o_A), w.mu = (double *)(base + o_mu), w.Y = (double *)(base + o_Y), w.Z = (double *)(base + o_Z);
    w.part = (double *)(base + o_part), w.gram = (double *)(base + o_gram), w.W = (double *)(base + o_W);
    w.fail = (int *)(base + o_fail);
    w.Q = nullptr, w.N = N, w.G = G, w.s = size, w.sp = sp, w.st = st;

    WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
    WOTB_CUDA(cudaMemcpyAsync(w.A, m1, (size_t)n1 * G * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemcpyAsync(w.A + (size_t)n1 * G, m2, (size_t)n2 * G * 8, cudaMemcpyHostToDevice, st));
    WOTB_CUDA(cudaMemsetAsync(w.fail, 0, 4, st));
    // centring: gene means over cells (util.py:245), then every cell's mean over genes (PCA.fit on x.T)
    k_pca_colsum<<<dim3((unsigned)cdiv(G, 128), slabs_col), 128, 0, st>>>(w.A, N, G, 256, w.part);
    k_pca_reduce<<<(unsigned)cdiv(G, 256), 256, 0, st>>>(w.part, slabs_col, G, 1.0 / (double)N, w.mu);
    double *cell_means = (double *)(base + o_cm);
    k_pca_center<<<(unsigned)cdiv(N, 8), 256, 0, st>>>(w.A, N, G, w.mu, cell_means);
    // Q0 [small, size] -> padded [small, sp] in the buffer of the short side
    double *Qs = transpose ? w.Z : w.Y, *Qt = transpose ? w.Y : w.Z;  // short-side / tall-side iterates
    WOTB_CUDA(cudaMemsetAsync(Qs, 0, (size_t)small * sp * 8, st));
    WOTB_CUDA(cudaMemcpy2DAsync(Qs, (size_t)sp * 8, q0, (size_t)size * 8, (size_t)size * 8, (size_t)small, cudaMemcpyHostToDevice, st));
    auto to_tall = [&]() { transpose ? pca_aq(w, Qs, Qt) : pca_aty(w, Qs, Qt); };   // Qt = M Qs
    auto to_short = [&]() { transpose ? pca_aty(w, Qt, Qs) : pca_aq(w, Qt, Qs); };  // Qs = M^T Qt
    for (int it = 0; it < n_iter; ++it) {
        to_tall();
        pca_orth(w, Qt, tall);
        to_short();
        pca_orth(w, Qs, small);
    }
    to_tall();
    pca_orth(w, Qt, tall);  // Q of the range finder, [tall, sp]
    to_short();             // B^T = M^T Q, [small, sp]
    pca_gram(w, Qs, small); // B B^T
    std::vector<double> gram((size_t)sp * sp), vec, val;
    int fail = 0;
    WOTB_CUDA(cudaMemcpyAsync(gram.data(), w.gram, gram.size() * 8, cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaMemcpyAsync(&fail, w.fail, 4, cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    if (fail) {
        set_error("local PCA: a Gram matrix of the range finder is not positive definite (rank-deficient input?)");
        return WOTB_ERR_INVALID;
    }
    std::vector<double> g2((size_t)size * size);
    for (int r = 0; r < size; ++r)
        for (int c = 0; c < size; ++c) g2[(size_t)r * size + c] = 0.5 * (gram[(size_t)r * sp + c] + gram[(size_t)c * sp + r]);
    jacobi_eigh(g2, size, vec, val);
    std::vector<int> order(size);
    for (int i = 0; i < size; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return val[x] > val[y]; });
    // W [sp, k]: transpose case components = Q Uhat[:, :k]; otherwise components = B^T Uhat[:, :k] / s
    std::vector<double> Wh((size_t)sp * k, 0.0);
    for (int j = 0; j < k; ++j) {
        const double sv = sqrt(val[order[j]] > 0 ? val[order[j]] : 0.0);
        sv_host[j] = sv;
        for (int r = 0; r < size; ++r) Wh[(size_t)r * k + j] = vec[(size_t)r * size + order[j]] * (transpose ? 1.0 : (sv > 0 ? 1.0 / sv : 0.0));
    }
    WOTB_CUDA(cudaMemcpyAsync(w.W, Wh.data(), Wh.size() * 8, cudaMemcpyHostToDevice, st));
    const double *src = transpose ? Qt : Qs;  // [N, sp] either way
    double *comp_dev = w.part;                // [N, k]
    k_pca_apply<<<(unsigned)cdiv(N, 16), 256, 0, st>>>(src, N, sp, w.W, k, k, k, comp_dev);
    WOTB_CUDA(cudaMemcpyAsync(comp_host, comp_dev, (size_t)N * k * 8, cudaMemcpyDeviceToHost, st));
    if (gene_means_host) WOTB_CUDA(cudaMemcpyAsync(gene_means_host, w.mu, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
    if (cell_means_host) WOTB_CUDA(cudaMemcpyAsync(cell_means_host, cell_means, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    WOTB_CUDA(cudaStreamSynchronize(st));
    WOTB_CUDA(cudaGetLastError());
    if (gpu_ms) {
        float ms = 0.f;
        WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *gpu_ms = ms;
    }
    return WOTB_OK;
}

}  // namespace wotb

extern "C" int wotb_pca_host(wotb_ctx *ctx, const double *m1_host, int64_t n1, const double *m2_host, int64_t n2, int64_t genes,
                             int32_t k, const double *q0_host, int32_t size, int32_t n_iter, double *comp_host,
                             double *singular_values_host, double *gene_means_host, double *cell_means_host,
                             double *gpu_ms) {
    return wotb::pca_host(ctx, m1_host, n1, m2_host, n2, genes, k, q0_host, size, n_iter, comp_host, singular_values_host,
                          gene_means_host, cell_means_host, gpu_ms);
}
