// Default cost of Waddington-OT on sm_100a: squared Euclidean distances between (singular-value
// scaled) local-PCA coordinates, divided by the median of all I*J distances.
//
// Replaces OTModel.compute_default_cost_matrix, /root/reference/wot/ot/ot_model.py:242-253:
//   :245-247  a.dot(eigenvals), b.dot(eigenvals)       -> k_scale_coords
//   :249-251  pairwise_distances(metric='sqeuclidean') -> dist_tiles<> (float64, direct differences
//             accumulated in dimension order like scipy's cdist, not the |x|^2+|y|^2-2xy expansion)
//   :252      cost_matrix / np.median(cost_matrix)     -> exact 64-bit radix select that recomputes
//             the distances each pass (6 digit passes), so the I x J float64 matrix the reference
//             sorts through is never stored; then one pass writes dist/median rounded to fp32.
// Also here: the coupling materialisation exp((f_i + g_j - C_ij)/eps)/J (optimal_transport.py:153,164).
#include <math.h>

#include <stddef.h>
#include <stdlib.h>

#include "common.cuh"

namespace wotb {

constexpr int kTile = 64;      // distances per CTA tile edge
constexpr int kTileThreads = 256;
constexpr int kChunk = 32;     // coordinate dimensions staged per shared-memory chunk
constexpr int kTilePad = kTile + 2;

__global__ void k_scale_coords(const double *__restrict__ x, const double *__restrict__ scale, double *__restrict__ out,
                               long long n, int d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n * d) out[idx] = scale ? __dmul_rn(x[idx], scale[idx % d]) : x[idx];
}

// Walks 64x64 tiles of the distance matrix; each thread owns a 4x4 micro-tile and hands every
// finished distance to `sink(i, j, dist)`.  Accumulation order over dimensions is ascending with
// separate subtract / multiply / add roundings (scipy cdist arithmetic), so every pass of the
// median select and the final cost pass see bit-identical values.
template <typename Sink>
__device__ __forceinline__ void dist_tiles(const double *__restrict__ x0, long long I, const double *__restrict__ x1,
                                           long long J, int d, Sink &sink) {
    __shared__ double xs[kChunk][kTilePad];
    __shared__ double ys[kChunk][kTilePad];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long tiles_i = (I + kTile - 1) / kTile, tiles_j = (J + kTile - 1) / kTile;
    for (long long t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
        const long long i0 = (t / tiles_j) * kTile, j0 = (t % tiles_j) * kTile;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
        for (int k0 = 0; k0 < d; k0 += kChunk) {
            const int kc = min(kChunk, d - k0);
            __syncthreads();
            for (int e = threadIdx.x; e < kTile * kChunk; e += kTileThreads) {
                const int k = e % kChunk, r = e / kChunk;
                const long long gi = i0 + r, gj = j0 + r;
                xs[k][r] = (k < kc && gi < I) ? x0[gi * d + k0 + k] : 0.0;
                ys[k][r] = (k < kc && gj < J) ? x1[gj * d + k0 + k] : 0.0;
            }
            __syncthreads();
            for (int k = 0; k < kc; ++k) {
                double xv[4], yv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) xv[r] = xs[k][ty * 4 + r];
#pragma unroll
                for (int c = 0; c < 4; ++c) yv[c] = ys[k][tx * 4 + c];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double df = __dsub_rn(xv[r], yv[c]);
                        acc[r][c] = __dadd_rn(acc[r][c], __dmul_rn(df, df));
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) sink.row(i0 + ty * 4 + r, j0 + tx * 4, acc[r], I, J);
    }
}

// ---- cost matrix writer --------------------------------------------------------------------------
template <typename T>
struct CostSink {
    T *C;
    long long ldc;
    double median;
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
        if (i >= I) return;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (j + c < J) C[i * ldc + j + c] = (T)__ddiv_rn(v[c], median);
    }
};

template <typename T>
__global__ void __launch_bounds__(kTileThreads) k_cost(const double *x0, long long I, const double *x1, long long J,
                                                       int d, T *C, long long ldc, double median) {
    CostSink<T> sink{C, ldc, median};
    dist_tiles(x0, I, x1, J, d, sink);
}

// zero the padding columns [J, ldc) so vectorised readers may touch them
__global__ void k_zero_pad(float *C, long long ldc, long long I, long long J) {
    const long long pad = ldc - J;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pad > 0 && idx < I * pad) C[(idx / pad) * ldc + J + idx % pad] = 0.f;
}

__global__ void k_cost_to_f32(const double *__restrict__ src, long long ld_src, long long I, long long J,
                              float *__restrict__ dst, long long ld_dst) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= I * ld_dst) return;
    const long long i = idx / ld_dst, j = idx % ld_dst;
    dst[idx] = j < J ? (float)src[i * ld_src + j] : 0.f;
}

// ---- exact median: MSB radix select over the float64 bit patterns --------------------------------
// Non-negative doubles order like their bit patterns.  Digits, most significant first: 9 bits, then
// five of 11 bits.  Each pass histograms the current digit of the values whose higher bits match
// the prefix found so far; k_select_pick then walks the 2048 bins to the one holding the rank.
constexpr int kSelBins = 2048;

struct SelectState {
    unsigned long long hist[kSelBins];
    unsigned long long prefix;     // higher bits of the answer found so far
    unsigned long long rank;       // rank still to find inside the prefix bucket
    unsigned long long n_less;     // elements strictly below the prefix bucket
    unsigned long long n_bucket;   // elements inside the prefix bucket
    unsigned long long min_above;  // bit pattern of the smallest value above the selected one
    int pass;
    int use_flat;                  // set by k_select_pick after the second digit when the bucket fits the gather buffer
};

__host__ __device__ inline int sel_shift(int pass) { return pass == 0 ? 55 : 55 - 11 * pass; }
__host__ __device__ inline int sel_bits(int pass) { return pass == 0 ? 9 : 11; }

struct SelectSink {
    unsigned int *hist;  // shared, kSelBins
    unsigned long long prefix;
    int pass;
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
        const int shift = sel_shift(pass), bits = sel_bits(pass);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[c]);
            const bool in = i < I && j + c < J && (pass == 0 || (key >> (shift + bits)) == prefix);
            const unsigned int digit = in ? (unsigned int)((key >> shift) & ((1u << bits) - 1u)) : 0xffffffffu;
            // warp-aggregated shared atomics: early passes put nearly every value in a few bins
            const unsigned int peers = __match_any_sync(0xffffffffu, digit);
            if (in && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
        }
    }
};

__global__ void __launch_bounds__(kTileThreads) k_select_hist(const double *x0, long long I, const double *x1,
                                                              long long J, int d, SelectState *st) {
    __shared__ unsigned int hist[kSelBins];
    if (st->use_flat) return;  // the remaining digits are found on the gathered bucket (k_flat_hist)
    for (int b = threadIdx.x; b < kSelBins; b += kTileThreads) hist[b] = 0;
    __syncthreads();
    SelectSink sink{hist, st->prefix, st->pass};
    dist_tiles(x0, I, x1, J, d, sink);
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += kTileThreads)
        if (hist[b]) atomicAdd(&st->hist[b], (unsigned long long)hist[b]);
}

// One CTA of 1024 threads: the 2048 bins are scanned with a block-wide prefix sum instead of by one thread (a pick
// used to cost ~10 us; a median runs 6 - 18 of them).
constexpr int kPickThreads = 1024;

__global__ void __launch_bounds__(kPickThreads) k_select_pick(SelectState *st, unsigned int flat_cap) {
    __shared__ unsigned long long warp_tot[kPickThreads / 32];
    __shared__ unsigned long long s_below, s_n;
    __shared__ int s_digit;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int bits = sel_bits(st->pass), n_bins = 1 << bits;
    const unsigned long long rank = st->rank;
    // thread t owns bins 2t, 2t + 1 (n_bins <= 2048)
    const unsigned long long h0 = 2 * tid < n_bins ? st->hist[2 * tid] : 0ull;
    const unsigned long long h1 = 2 * tid + 1 < n_bins ? st->hist[2 * tid + 1] : 0ull;
    unsigned long long incl = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[wid] = incl;
    if (tid == 0) s_digit = -1;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    const unsigned long long before = base + incl - (h0 + h1);  // values in the bins below bin 2t
    if (rank >= before && rank < before + h0) {
        s_digit = 2 * tid, s_below = before, s_n = h0;
    } else if (rank >= before + h0 && rank < before + h0 + h1) {
        s_digit = 2 * tid + 1, s_below = before + h0, s_n = h1;
    }
    __syncthreads();
    if (tid == 0) {
        int digit = s_digit;
        unsigned long long below = s_below;
        if (digit < 0) {  // nothing selected (an idle pass of a sequence that has switched paths): keep the state
            digit = 0;
            below = 0;
            for (int w = 0; w < kPickThreads / 32; ++w) below += warp_tot[w];
        } else {
            st->n_bucket = s_n;
        }
        st->prefix = (st->prefix << bits) | (unsigned long long)digit;
        st->rank = rank - below;
        st->n_less += below;
        st->pass += 1;
        if (st->pass == 2 && flat_cap > 0 && st->n_bucket <= (unsigned long long)flat_cap) st->use_flat = 1;
    }
    for (int b = tid; b < kSelBins; b += kPickThreads) st->hist[b] = 0;
}

struct MinAboveSink {
    unsigned long long key1;
    unsigned long long best;
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[c]);
            if (i < I && j + c < J && key > key1 && key < best) best = key;
        }
    }
};

__global__ void __launch_bounds__(kTileThreads) k_select_min_above(const double *x0, long long I, const double *x1,
                                                                   long long J, int d, SelectState *st) {
    MinAboveSink sink{st->prefix, ~0ull};
    dist_tiles(x0, I, x1, J, d, sink);
    unsigned long long best = sink.best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&st->min_above, best);
}

// ---- shortcut: once the prefix bucket is small, gather it and finish the select on the flat buffer ----------
// After the 9- and 11-bit passes the bucket that holds the median is a sliver of one binade (~0.2 % of the I*J
// values at atlas shapes): one more pass over the tiles appends its members to a buffer, and the remaining four
// digits (and the upper middle element for even counts) are found there, instead of four more FP64 passes over
// all I*J distances.  Same values, same order statistics: the result is still the exact np.median.
struct CollectSink {
    unsigned long long *buf;
    unsigned int *count;
    unsigned long long prefix;
    unsigned int cap;
    int shift;  // bits below the prefix
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[c]);
            const bool in = i < I && j + c < J && (key >> shift) == prefix;
            const unsigned int mask = __ballot_sync(0xffffffffu, in);
            if (mask) {
                const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(count, (unsigned int)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                const unsigned int at = base + (unsigned int)__popc(mask & ((1u << lane) - 1u));
                if (in && at < cap) buf[at] = key;
            }
        }
    }
};

__global__ void __launch_bounds__(kTileThreads) k_select_collect(const double *x0, long long I, const double *x1,
                                                                 long long J, int d, const SelectState *st,
                                                                 unsigned long long *buf, unsigned int *count,
                                                                 unsigned int cap) {
    if (!st->use_flat) return;
    CollectSink sink{buf, count, st->prefix, cap, sel_shift(st->pass - 1)};
    dist_tiles(x0, I, x1, J, d, sink);
}

__global__ void k_flat_hist(const unsigned long long *__restrict__ buf, const unsigned int *__restrict__ count,
                            SelectState *st) {
    __shared__ unsigned int hist[kSelBins];
    if (!st->use_flat) return;
    const unsigned int n = *count;
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const int pass = st->pass, shift = sel_shift(pass), bits = sel_bits(pass);
    const unsigned long long prefix = st->prefix;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const unsigned long long key = buf[e];
        const bool in = pass == 0 || (key >> (shift + bits)) == prefix;   // pass 0 has no prefix (and no 64-bit shift)
        if (in) atomicAdd(&hist[(unsigned int)((key >> shift) & ((1u << bits) - 1u))], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
        if (hist[b]) atomicAdd(&st->hist[b], (unsigned long long)hist[b]);
}

__global__ void k_flat_min_above(const unsigned long long *__restrict__ buf, const unsigned int *__restrict__ count,
                                 SelectState *st) {
    if (!st->use_flat) return;
    const unsigned int n = *count;
    const unsigned long long key1 = st->prefix;
    unsigned long long best = ~0ull;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const unsigned long long key = buf[e];
        if (key > key1 && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&st->min_above, best);
}

constexpr unsigned int kSelCollectCap = 1u << 22;  // 4 M keys = 32 MB

static int tile_grid(const wotb_ctx *ctx, int64_t I, int64_t J) {
    const int64_t tiles = cdiv(I, kTile) * cdiv(J, kTile);
    const int64_t cap = (int64_t)ctx->sm_count * 4;
    return (int)(tiles < cap ? tiles : cap);
}

// coordinates multiplied by `scale` (or copied) into the context's staging buffer
int scaled_coords(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, const double *scale,
                  const double **xs0, const double **xs1) {
    if (!scale) {
        *xs0 = x0;
        *xs1 = x1;
        return WOTB_OK;
    }
    WOTB_TRY(ctx->onl.reserve((size_t)(I + J) * d * 8 + 512));
    double *a = ctx->onl.as<double>();
    double *b = a + round_up(I * d, 32);
    k_scale_coords<<<(unsigned)cdiv(I * d, 256), 256, 0, ctx->stream>>>(x0, scale, a, I, d);
    k_scale_coords<<<(unsigned)cdiv(J * d, 256), 256, 0, ctx->stream>>>(x1, scale, b, J, d);
    WOTB_CUDA(cudaGetLastError());
    *xs0 = a;
    *xs1 = b;
    return WOTB_OK;
}

int scale_into(wotb_ctx *ctx, const double *x, int64_t n, int d, const double *scale, double *out) {
    k_scale_coords<<<(unsigned)cdiv(n * d, 256), 256, 0, ctx->stream>>>(x, scale, out, n, d);
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// ---- one-pass median: a sampled window instead of two digit passes -------------------------------------------------
// The radix select above needs three FP64 passes over all I*J distances (two digits + the gather), 9 ms of a 125 ms
// atlas step and 0.57 s at 100k x 100k.  Two of them only LOCATE the median; a random sample does that as well:
// n_s distances of uniformly drawn pairs are selected for the ranks n_s (1/2 -+ delta), delta = 3 / sqrt(n_s) (six
// standard deviations of a sample quantile), which brackets the true median with probability 1 - 2e-9.  ONE pass
// over the tiles then counts the distances below the window and gathers the ones inside it (0.1 - 0.6 % of them);
// the exact order statistics are found on that buffer with the same flat radix select as before.  Every value is
// still produced by the scipy-ordered float64 arithmetic of dist_tiles, the counts are exact, and the device checks
// that the wanted ranks really fall inside the gathered window: if not (heavy ties, a bucket larger than the
// buffer, the one-in-a-billion sample) the three-pass select runs instead.  The result is the exact np.median.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_median_sample(const double *__restrict__ x0, long long I, const double *__restrict__ x1, long long J, int d,
                                unsigned long long *__restrict__ keys, unsigned int n_s) {
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_s) return;
    const unsigned long long h = splitmix64(0x5EEDull + e), h2 = splitmix64(h);
    const long long i = (long long)__umul64hi(h, (unsigned long long)I), j = (long long)__umul64hi(h2, (unsigned long long)J);
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
        const double df = __dsub_rn(x0[i * d + k], x1[j * d + k]);
        acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    keys[e] = (unsigned long long)__double_as_longlong(acc);
}

struct WindowSink {
    unsigned long long *buf;
    unsigned int *count;
    unsigned long long lo, hi;
    unsigned int cap;
    unsigned long long below;  // per thread
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[c]);
            const bool valid = i < I && j + c < J;
            below += (valid && key < lo) ? 1ull : 0ull;
            const bool in = valid && key >= lo && key <= hi;
            const unsigned int mask = __ballot_sync(0xffffffffu, in);
            if (mask) {
                const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(count, (unsigned int)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                const unsigned int at = base + (unsigned int)__popc(mask & ((1u << lane) - 1u));
                if (in && at < cap) buf[at] = key;
            }
        }
    }
};

// st_lo / st_hi: finished flat selects on the sample (their `prefix` is the full key of the window's ends)
__global__ void __launch_bounds__(kTileThreads) k_median_window(const double *x0, long long I, const double *x1, long long J,
                                                                int d, const SelectState *st_lo, const SelectState *st_hi,
                                                                unsigned long long *buf, unsigned int *count, unsigned int cap,
                                                                unsigned long long *below_total) {
    WindowSink sink{buf, count, st_lo->prefix, st_hi->prefix, cap, 0ull};
    dist_tiles(x0, I, x1, J, d, sink);
    unsigned long long b = sink.below;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(below_total, b);
}

// Does the window hold the wanted order statistics?  Then the main select continues on the gathered buffer.
__global__ void k_median_window_check(SelectState *st, const unsigned int *count, unsigned int cap,
                                      const unsigned long long *below_total, unsigned long long n, int *fallback) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long below = *below_total, c = *count;
    const unsigned long long target = (n - 1) / 2, upper = n / 2;  // lower and upper middle ranks
    const bool ok = c <= (unsigned long long)cap && below <= target && upper < below + c;
    *fallback = ok ? 0 : 1;
    st->prefix = 0, st->pass = 0, st->n_bucket = 0, st->min_above = ~0ull;
    for (int b = 0; b < kSelBins; ++b) st->hist[b] = 0;
    if (ok) {
        st->use_flat = 1;
        st->n_less = below;
        st->rank = target - below;
    } else {  // the state the three-pass select starts from
        st->use_flat = 0;
        st->n_less = 0;
        st->rank = target;
    }
}

// median_window_plan: sample size, window half-width and gather capacity for an I x J problem (0 = use cost_median).
static unsigned int median_window_plan(unsigned long long n, unsigned int *n_s_out) {
    unsigned int n_s = 1u << 20;
    while ((unsigned long long)n_s * 256ull < n && n_s < (1u << 24)) n_s <<= 1;
    if (n < 60000000ull) n_s = 0;
    *n_s_out = n_s;
    if (!n_s) return 0;
    const double delta = 3.0 / sqrt((double)n_s);
    unsigned long long cap64 = (unsigned long long)(2.5 * 2.0 * delta * (double)n) + 65536ull;
    if (cap64 < kSelCollectCap) cap64 = kSelCollectCap;
    return cap64 <= (1ull << 27) ? (unsigned int)cap64 : 0u;
}

int cost_median(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, const double *scale,
                double *median_host) {
    WOTB_REQUIRE(ctx && x0 && x1 && median_host, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1, "I, J, d must be >= 1");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    const double *a, *b;
    WOTB_TRY(scaled_coords(ctx, x0, I, x1, J, d, scale, &a, &b));
    WOTB_TRY(ctx->select.reserve(3 * sizeof(SelectState)));  // the select itself + the two ends of the sampled window
    SelectState *st = ctx->select.as<SelectState>();
    const unsigned long long n = (unsigned long long)I * (unsigned long long)J;
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.rank = (n - 1) / 2;  // lower middle element; np.median averages it with the upper one for even n
    init.min_above = ~0ull;
    WOTB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    const int grid = tile_grid(ctx, I, J);
    // One launch sequence, no host decision inside it: after the second digit k_select_pick sets use_flat on the
    // device when the bucket fits the gather buffer; from then on the tile-pass kernels return at once and the
    // flat kernels do the work (or the other way round).  The scalar tail of SelectState comes back through
    // pinned memory.
    const bool shortcut = getenv("WOTB_NO_MEDIAN_SHORTCUT") == nullptr;
    WOTB_TRY(ctx->status.reserve(256));
    struct Tail {
        unsigned long long prefix, rank, n_less, n_bucket, min_above;
        int pass, use_flat;
    };
    static_assert(offsetof(SelectState, use_flat) - offsetof(SelectState, prefix) == offsetof(Tail, use_flat), "SelectState tail layout");
    Tail *pin = reinterpret_cast<Tail *>(ctx->status.as<char>() + 64);
    auto read_tail = [&](Tail *out) -> int {
        WOTB_CUDA(cudaMemcpyAsync(pin, &st->prefix, sizeof(Tail), cudaMemcpyDeviceToHost, ctx->stream));
        WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = *pin;
        return WOTB_OK;
    };
    // ---- sampled window (one pass over the distances) -------------------------------------------------------------
    // measured on B200 (profiles/r2q_median.txt): 39 M distances 2.2 ms (window) vs 1.6 ms (three passes), 155 M 3.3 vs
    // 4.6, 400 M 5.4 vs 10.5, 2.5 G 23 vs 134, 10 G 79 vs 526 -- the sample select has a fixed cost of ~24 small launches,
    // so the window is used from 6e7 distances on (median_window_plan)
    unsigned int n_s = 0;
    unsigned long long cap64 = median_window_plan(n, &n_s);
    const double delta = n_s ? 3.0 / sqrt((double)n_s) : 0.0;
    if (cap64 == 0) n_s = 0;
    if (cap64 < kSelCollectCap) cap64 = kSelCollectCap;
    const bool windowed = shortcut && n_s > 0 && cap64 <= (1ull << 27) && getenv("WOTB_NO_MEDIAN_WINDOW") == nullptr;
    const unsigned int cap = windowed ? (unsigned int)cap64 : kSelCollectCap;
    // grow-only buffers that never have to be re-allocated mid-run for pairs of similar size (cudaFree is a
    // device-wide synchronisation, which would stall the other streams of a pipeline)
    WOTB_TRY(ctx->part.reserve((size_t)cap * 8 + (size_t)n_s * 8 + 512));
    unsigned long long *buf = ctx->part.as<unsigned long long>() + 32;
    unsigned int *count = ctx->part.as<unsigned int>();
    unsigned int *count_s = count + 1;
    unsigned long long *below_total = ctx->part.as<unsigned long long>() + 2;
    int *fallback_dev = ctx->part.as<int>() + 8;
    unsigned long long *sample = buf + cap;
    WOTB_CUDA(cudaMemsetAsync(count, 0, 64, ctx->stream));
    const unsigned int flat_cap = shortcut ? kSelCollectCap : 0u;
    bool done_by_window = false;
    if (windowed) {
        SelectState *st_q[2] = {st + 1, st + 2};
        k_median_sample<<<(unsigned)cdiv(n_s, 256), 256, 0, ctx->stream>>>(a, I, b, J, d, sample, n_s);
        WOTB_CUDA(cudaMemcpyAsync(count_s, &n_s, 4, cudaMemcpyHostToDevice, ctx->stream));
        const double q[2] = {0.5 - delta, 0.5 + delta};
        for (int w = 0; w < 2; ++w) {
            SelectState sq;
            memset(&sq, 0, sizeof(sq));
            double r = q[w] * (double)n_s;
            r = r < 0 ? 0 : (r > (double)(n_s - 1) ? (double)(n_s - 1) : r);
            sq.rank = (unsigned long long)r;
            sq.min_above = ~0ull;
            sq.use_flat = 1;
            WOTB_CUDA(cudaMemcpyAsync(st_q[w], &sq, sizeof(sq), cudaMemcpyHostToDevice, ctx->stream));
            for (int pass = 0; pass < 6; ++pass) {
                k_flat_hist<<<592, 256, 0, ctx->stream>>>(sample, count_s, st_q[w]);
                k_select_pick<<<1, kPickThreads, 0, ctx->stream>>>(st_q[w], 0u);
            }
        }
        k_median_window<<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, st_q[0], st_q[1], buf, count, cap, below_total);
        k_median_window_check<<<1, 32, 0, ctx->stream>>>(st, count, cap, below_total, n, fallback_dev);
        for (int pass = 0; pass < 6; ++pass) {      // no-ops on the device (use_flat == 0) when the check failed
            k_flat_hist<<<592, 256, 0, ctx->stream>>>(buf, count, st);
            k_select_pick<<<1, kPickThreads, 0, ctx->stream>>>(st, 0u);
        }
        if (n % 2 == 0) k_flat_min_above<<<592, 256, 0, ctx->stream>>>(buf, count, st);
        int *fb_pin = reinterpret_cast<int *>(ctx->status.as<char>() + 192);
        WOTB_CUDA(cudaMemcpyAsync(fb_pin, fallback_dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
        Tail t;
        WOTB_TRY(read_tail(&t));
        done_by_window = *fb_pin == 0;
        if (!done_by_window) {  // start the three-pass select from scratch
            WOTB_CUDA(cudaMemsetAsync(count, 0, 4, ctx->stream));
            WOTB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    if (!done_by_window)
    for (int pass = 0; pass < 6; ++pass) {
        if (pass == 2) k_select_collect<<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, st, buf, count, kSelCollectCap);
        if (pass >= 2) k_flat_hist<<<592, 256, 0, ctx->stream>>>(buf, count, st);
        k_select_hist<<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, st);
        k_select_pick<<<1, kPickThreads, 0, ctx->stream>>>(st, flat_cap);
    }
    if (n % 2 == 0 && !done_by_window) k_flat_min_above<<<592, 256, 0, ctx->stream>>>(buf, count, st);
    Tail fin;
    WOTB_TRY(read_tail(&fin));
    double lo, hi;
    memcpy(&lo, &fin.prefix, 8);
    hi = lo;
    if (n % 2 == 0 && fin.n_less + fin.n_bucket <= n / 2) {
        // the upper middle element is the smallest value strictly above `lo`: the gathered bucket usually has it
        if (!(fin.use_flat && fin.min_above != ~0ull)) {  // `lo` is the largest member of its bucket (or no bucket)
            k_select_min_above<<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, st);
            WOTB_TRY(read_tail(&fin));
        }
        memcpy(&hi, &fin.min_above, 8);
    }
    WOTB_CUDA(cudaGetLastError());
    *median_host = (lo + hi) / 2.0;
    return WOTB_OK;
}

// ---- the one-pass median, split over row shards (row-sharded solves: every rank holds all coordinates) ------------------
int median_window_cap(int64_t I, int64_t J, int64_t *cap) {
    unsigned int n_s = 0;
    *cap = (int64_t)median_window_plan((unsigned long long)I * (unsigned long long)J, &n_s);
    return WOTB_OK;
}

// Rows [row_lo, row_hi) of the window pass: distances below the window are counted into *below, the ones inside are
// appended to keys[0 : cap) (*count may exceed cap: overflow, detected by median_window_finish).  The sample and
// therefore the window are the same on every rank.
int median_window_rows(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, const double *scale,
                       int64_t row_lo, int64_t row_hi, unsigned long long *keys, int64_t cap, unsigned int *count,
                       unsigned long long *below) {
    WOTB_REQUIRE(ctx && x0 && x1 && keys && count && below, "NULL argument");
    WOTB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= I, "bad row range");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    const unsigned long long n = (unsigned long long)I * (unsigned long long)J;
    unsigned int n_s = 0;
    const unsigned int want_cap = median_window_plan(n, &n_s);
    WOTB_REQUIRE(want_cap > 0 && (unsigned long long)cap >= want_cap, "problem too small for the windowed median, or cap too small");
    const double *a, *b;
    WOTB_TRY(scaled_coords(ctx, x0, I, x1, J, d, scale, &a, &b));
    WOTB_TRY(ctx->select.reserve(3 * sizeof(SelectState)));
    SelectState *st = ctx->select.as<SelectState>();
    SelectState *st_q[2] = {st + 1, st + 2};
    WOTB_TRY(ctx->part.reserve((size_t)n_s * 8 + 512));
    unsigned int *count_s = ctx->part.as<unsigned int>();
    unsigned long long *sample = ctx->part.as<unsigned long long>() + 32;
    k_median_sample<<<(unsigned)cdiv(n_s, 256), 256, 0, ctx->stream>>>(a, I, b, J, d, sample, n_s);
    WOTB_CUDA(cudaMemcpyAsync(count_s, &n_s, 4, cudaMemcpyHostToDevice, ctx->stream));
    const double delta = 3.0 / sqrt((double)n_s);
    const double q[2] = {0.5 - delta, 0.5 + delta};
    for (int w = 0; w < 2; ++w) {
        SelectState sq;
        memset(&sq, 0, sizeof(sq));
        double r = q[w] * (double)n_s;
        r = r < 0 ? 0 : (r > (double)(n_s - 1) ? (double)(n_s - 1) : r);
        sq.rank = (unsigned long long)r;
        sq.min_above = ~0ull;
        sq.use_flat = 1;
        WOTB_CUDA(cudaMemcpyAsync(st_q[w], &sq, sizeof(sq), cudaMemcpyHostToDevice, ctx->stream));
        for (int pass = 0; pass < 6; ++pass) {
            k_flat_hist<<<592, 256, 0, ctx->stream>>>(sample, count_s, st_q[w]);
            k_select_pick<<<1, kPickThreads, 0, ctx->stream>>>(st_q[w], 0u);
        }
    }
    WOTB_CUDA(cudaMemsetAsync(count, 0, 4, ctx->stream));
    WOTB_CUDA(cudaMemsetAsync(below, 0, 8, ctx->stream));
    const int64_t rows = row_hi - row_lo;
    if (rows > 0) {
        const int grid = tile_grid(ctx, rows, J);
        k_median_window<<<grid, kTileThreads, 0, ctx->stream>>>(a + (size_t)row_lo * d, rows, b, J, d, st_q[0], st_q[1], keys,
                                                                count, (unsigned int)cap, below);
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// keys[0 : count): the union of every shard's window, below: the sum of their counts.  *ok = 0 when the window does not
// hold the middle ranks (the caller then runs cost_median).
int median_window_finish(wotb_ctx *ctx, int64_t I, int64_t J, const unsigned long long *keys, int64_t count_total,
                         unsigned long long below_total, double *median_host, int *ok) {
    WOTB_REQUIRE(ctx && keys && median_host && ok, "NULL argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    const unsigned long long n = (unsigned long long)I * (unsigned long long)J;
    WOTB_REQUIRE(count_total >= 0 && count_total < (1ll << 31), "too many keys");
    WOTB_TRY(ctx->select.reserve(3 * sizeof(SelectState)));
    SelectState *st = ctx->select.as<SelectState>();
    WOTB_TRY(ctx->part.reserve(512));
    unsigned int *count = ctx->part.as<unsigned int>();
    unsigned long long *below = ctx->part.as<unsigned long long>() + 2;
    int *fallback_dev = ctx->part.as<int>() + 8;
    const unsigned int c32 = (unsigned int)count_total;
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.min_above = ~0ull;
    WOTB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(count, &c32, 4, cudaMemcpyHostToDevice, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(below, &below_total, 8, cudaMemcpyHostToDevice, ctx->stream));
    k_median_window_check<<<1, 32, 0, ctx->stream>>>(st, count, c32, below, n, fallback_dev);
    for (int pass = 0; pass < 6; ++pass) {
        k_flat_hist<<<592, 256, 0, ctx->stream>>>(keys, count, st);
        k_select_pick<<<1, kPickThreads, 0, ctx->stream>>>(st, 0u);
    }
    if (n % 2 == 0) k_flat_min_above<<<592, 256, 0, ctx->stream>>>(keys, count, st);
    WOTB_TRY(ctx->status.reserve(256));
    struct Tail {
        unsigned long long prefix, rank, n_less, n_bucket, min_above;
        int pass, use_flat;
    };
    Tail *pin = reinterpret_cast<Tail *>(ctx->status.as<char>() + 64);
    int *fb_pin = reinterpret_cast<int *>(ctx->status.as<char>() + 192);
    WOTB_CUDA(cudaMemcpyAsync(pin, &st->prefix, sizeof(Tail), cudaMemcpyDeviceToHost, ctx->stream));
    WOTB_CUDA(cudaMemcpyAsync(fb_pin, fallback_dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WOTB_CUDA(cudaStreamSynchronize(ctx->stream));
    WOTB_CUDA(cudaGetLastError());
    const Tail fin = *pin;
    *ok = *fb_pin == 0;
    if (!*ok) return WOTB_OK;
    double lo, hi;
    memcpy(&lo, &fin.prefix, 8);
    hi = lo;
    if (n % 2 == 0 && fin.n_less + fin.n_bucket <= n / 2) {
        if (fin.min_above == ~0ull) {  // the upper middle value lies outside the window: let the caller fall back
            *ok = 0;
            return WOTB_OK;
        }
        memcpy(&hi, &fin.min_above, 8);
    }
    *median_host = (lo + hi) / 2.0;
    return WOTB_OK;
}

int cost_matrix(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, const double *scale,
                double median, void *C, int64_t ldc, int dtype) {
    WOTB_REQUIRE(ctx && x0 && x1 && C, "NULL argument");
    WOTB_REQUIRE(I >= 1 && J >= 1 && d >= 1 && ldc >= J, "bad shape");
    WOTB_REQUIRE(dtype == WOTB_F32 || dtype == WOTB_F64, "dtype must be WOTB_F32 or WOTB_F64");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    const double *a, *b;
    WOTB_TRY(scaled_coords(ctx, x0, I, x1, J, d, scale, &a, &b));
    const int grid = tile_grid(ctx, I, J);
    if (dtype == WOTB_F32) {
        k_cost<float><<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, (float *)C, ldc, median);
        if (ldc > J) k_zero_pad<<<(unsigned)cdiv(I * (ldc - J), 256), 256, 0, ctx->stream>>>((float *)C, ldc, I, J);
    } else {
        k_cost<double><<<grid, kTileThreads, 0, ctx->stream>>>(a, I, b, J, d, (double *)C, ldc, median);
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

int cost_to_f32(wotb_ctx *ctx, const double *src, int64_t ld_src, int64_t I, int64_t J, float *dst, int64_t ld_dst) {
    WOTB_REQUIRE(ctx && src && dst && ld_src >= J && ld_dst >= J, "bad argument");
    k_cost_to_f32<<<(unsigned)cdiv(I * ld_dst, 256), 256, 0, ctx->stream>>>(src, ld_src, I, J, dst, ld_dst);
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// ---- coupling materialisation ------------------------------------------------------------------
// tmap_ij = exp((f_i + g_j - C_ij)/eps) * scale, float64 argument and exp.  One CTA per row at a
// time so the row sum (optimal_transport.py:27, ot_model.py:319) is a fixed-order block reduction.
constexpr int kCoupThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kCoupThreads) k_coupling(const float *__restrict__ C, long long ldc, long long I,
                                                           long long J, const double *__restrict__ f,
                                                           const double *__restrict__ g, double inv_eps, double scale,
                                                           T *__restrict__ out, long long ldo,
                                                           double *__restrict__ rowsum) {
    __shared__ double red[kCoupThreads / 32];
    for (long long i = blockIdx.x; i < I; i += gridDim.x) {
        const double fi = f[i];
        double sum = 0.0;
        for (long long j = threadIdx.x; j < J; j += kCoupThreads) {
            const double v = exp((fi + g[j] - (double)C[i * ldc + j]) * inv_eps) * scale;
            out[i * ldo + j] = (T)v;
            sum += v;
        }
        if (rowsum) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int w = 0; w < kCoupThreads / 32; ++w) tot += red[w];
                rowsum[i] = tot;
            }
        }
    }
}

int coupling(wotb_ctx *ctx, const float *C, int64_t ldc, int64_t I, int64_t J, const double *f, const double *g,
             double eps, double out_scale, void *out, int64_t ldo, int dtype, double *rowsum, cudaStream_t stream) {
    WOTB_REQUIRE(ctx && C && f && g && out, "NULL argument");
    WOTB_REQUIRE(ldc >= J && ldo >= J && eps > 0, "bad shape");
    WOTB_REQUIRE(dtype == WOTB_F32 || dtype == WOTB_F64, "dtype must be WOTB_F32 or WOTB_F64");
    const int grid = (int)(I < ctx->sm_count * 8 ? I : ctx->sm_count * 8);
    if (dtype == WOTB_F32)
        k_coupling<float><<<grid, kCoupThreads, 0, stream>>>(C, ldc, I, J, f, g, 1.0 / eps, out_scale, (float *)out, ldo,
                                                              rowsum);
    else
        k_coupling<double><<<grid, kCoupThreads, 0, stream>>>(C, ldc, I, J, f, g, 1.0 / eps, out_scale, (double *)out,
                                                               ldo, rowsum);
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// Coupling straight from coordinates (online kernel: C is never stored); float64 distances, argument
// and exp, like the stored path.
template <typename T>
struct CouplingSink {
    T *out;
    long long ldo;
    const double *f, *g;
    double inv_median, inv_eps, scale;
    __device__ __forceinline__ void row(long long i, long long j, const double (&v)[4], long long I, long long J) {
        if (i >= I) return;
        const double fi = f[i];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (j + c < J) out[i * ldo + j + c] = (T)(exp((fi + g[j + c] - v[c] * inv_median) * inv_eps) * scale);
    }
};

template <typename T>
__global__ void __launch_bounds__(kTileThreads) k_coupling_online(const double *x0, long long I, const double *x1,
                                                                  long long J, int d, CouplingSink<T> sink) {
    dist_tiles(x0, I, x1, J, d, sink);
}

int coupling_online(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                    const double *f, const double *g, double eps, double out_scale, void *out, int64_t ldo, int dtype,
                    double *rowsum, cudaStream_t stream) {
    WOTB_REQUIRE(ctx && x0 && x1 && f && g && out, "NULL argument");
    WOTB_REQUIRE(ldo >= J && eps > 0 && median > 0, "bad shape");
    WOTB_REQUIRE(rowsum == nullptr, "row sums of the online coupling come from wotb_sinkhorn_online_dev");
    const int grid = tile_grid(ctx, I, J);
    if (dtype == WOTB_F32) {
        CouplingSink<float> sink{(float *)out, ldo, f, g, 1.0 / median, 1.0 / eps, out_scale};
        k_coupling_online<float><<<grid, kTileThreads, 0, stream>>>(x0, I, x1, J, d, sink);
    } else if (dtype == WOTB_F64) {
        CouplingSink<double> sink{(double *)out, ldo, f, g, 1.0 / median, 1.0 / eps, out_scale};
        k_coupling_online<double><<<grid, kTileThreads, 0, stream>>>(x0, I, x1, J, d, sink);
    } else {
        WOTB_REQUIRE(false, "dtype must be WOTB_F32 or WOTB_F64");
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// ---- a coupling applied to several populations at once, never materialised --------------------------
// Reference: TransportMapModel.push_forward / pull_back, wot/tmap/transport_map_model.py:290 (p @ tmap.X) and :356
// (tmap.X @ p.T), where p stacks ALL populations (np.vstack, :285 / :351).  tmap_ij = exp((f_i + g_j - C_ij)/eps) * scale is
// evaluated tile by tile in float64 exactly as k_coupling_online does -- ONE exponential per entry -- and every entry
// is then used NP times: out[k, o] += tmap[.,.] * p[k, .] (FP64 FMAs, which are cheap next to the distance and the
// exponential).  A CTA owns one 64-wide tile of the OUT side and a contiguous range of tiles of the side that is
// summed over; the partial results of the ranges are added in a fixed order by k_apply_reduce: deterministic.
constexpr int kApplyNP = 8;  // populations per sweep over the coupling

template <bool FORWARD>
__global__ void __launch_bounds__(kTileThreads)
    k_apply_multi(const double *__restrict__ x0, long long I, const double *__restrict__ x1, long long J, int d,
                  const double *__restrict__ f, const double *__restrict__ g, double inv_median, double inv_eps, double scale,
                  const double *__restrict__ P, long long ldp, int np, double *__restrict__ part, long long n_out) {
    __shared__ double xs[kChunk][kTilePad];
    __shared__ double ys[kChunk][kTilePad];
    __shared__ double pw[kApplyNP][kTile];
    __shared__ double red[16][kTile];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long n_in = FORWARD ? I : J;
    const long long tiles_in = (n_in + kTile - 1) / kTile;
    const long long o0 = (long long)blockIdx.x * kTile;  // first out entry of this CTA
    const long long t_lo = tiles_in * blockIdx.y / gridDim.y, t_hi = tiles_in * (blockIdx.y + 1) / gridDim.y;
    double acc_out[kApplyNP][4];
#pragma unroll
    for (int k = 0; k < kApplyNP; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc_out[k][c] = 0.0;
    for (long long t = t_lo; t < t_hi; ++t) {
        const long long in0 = t * kTile;
        const long long i0 = FORWARD ? in0 : o0, j0 = FORWARD ? o0 : in0;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
        for (int k0 = 0; k0 < d; k0 += kChunk) {
            const int kc = min(kChunk, d - k0);
            __syncthreads();
            for (int e = threadIdx.x; e < kTile * kChunk; e += kTileThreads) {
                const int k = e % kChunk, r = e / kChunk;
                const long long gi = i0 + r, gj = j0 + r;
                xs[k][r] = (k < kc && gi < I) ? x0[gi * d + k0 + k] : 0.0;
                ys[k][r] = (k < kc && gj < J) ? x1[gj * d + k0 + k] : 0.0;
            }
            if (k0 == 0) {
                for (int e = threadIdx.x; e < kApplyNP * kTile; e += kTileThreads) {
                    const int k = e / kTile, r = e % kTile;
                    pw[k][r] = (k < np && in0 + r < n_in) ? P[(long long)k * ldp + in0 + r] : 0.0;
                }
            }
            __syncthreads();
            for (int k = 0; k < kc; ++k) {
                double xv[4], yv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) xv[r] = xs[k][ty * 4 + r];
#pragma unroll
                for (int c = 0; c < 4; ++c) yv[c] = ys[k][tx * 4 + c];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double df = __dsub_rn(xv[r], yv[c]);
                        acc[r][c] = __dadd_rn(acc[r][c], __dmul_rn(df, df));
                    }
            }
        }
        // entries of the coupling (zero outside the matrix), then NP fused multiply-adds each
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const long long i = i0 + ty * 4 + r;
            const double fi = i < I ? f[i] : 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long j = j0 + tx * 4 + c;
                const double e = (i < I && j < J) ? exp((fi + g[j] - acc[r][c] * inv_median) * inv_eps) * scale : 0.0;
#pragma unroll
                for (int k = 0; k < kApplyNP; ++k) {
                    if (FORWARD)
                        acc_out[k][c] = fma(e, pw[k][ty * 4 + r], acc_out[k][c]);   // out = column j, weight of row i
                    else
                        acc_out[k][r] = fma(e, pw[k][tx * 4 + c], acc_out[k][r]);   // out = row i, weight of column j
                }
            }
        }
    }
    // reduce over the 16 threads that share an out entry (fixed order), one population at a time
    for (int k = 0; k < np; ++k) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (FORWARD)
                red[ty][tx * 4 + c] = acc_out[k][c];
            else
                red[tx][ty * 4 + c] = acc_out[k][c];
        }
        __syncthreads();
        if (threadIdx.x < kTile && o0 + threadIdx.x < n_out) {
            double sum = 0.0;
            for (int q = 0; q < 16; ++q) sum += red[q][threadIdx.x];
            part[((long long)blockIdx.y * np + k) * n_out + o0 + threadIdx.x] = sum;
        }
    }
}

__global__ void k_apply_reduce(const double *__restrict__ part, int n_split, int np, long long n_out, double *__restrict__ out,
                               long long ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)np * n_out) return;
    const long long k = idx / n_out, o = idx % n_out;
    double sum = 0.0;
    for (int sp = 0; sp < n_split; ++sp) sum += part[((long long)sp * np + k) * n_out + o];
    out[k * ldo + o] = sum;
}

// out[k, :] for k < n_pop (device pointers; P [n_pop, n_in] row-major, out [n_pop, n_out]); x0, x1 already scaled
int coupling_apply(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                   const double *f, const double *g, double eps, double out_scale, int forward, const double *P, int n_pop,
                   double *out) {
    const int64_t n_in = forward ? I : J, n_out = forward ? J : I;
    const int tiles_out = (int)cdiv(n_out, kTile), tiles_in = (int)cdiv(n_in, kTile);
    int n_split = (int)cdiv((int64_t)ctx->sm_count * 3, tiles_out);
    if (n_split > tiles_in) n_split = tiles_in;
    if (n_split < 1) n_split = 1;
    WOTB_TRY(ctx->part.reserve((size_t)n_split * kApplyNP * n_out * 8));
    double *part = ctx->part.as<double>();
    const dim3 grid(tiles_out, n_split);
    for (int k0 = 0; k0 < n_pop; k0 += kApplyNP) {
        const int np = n_pop - k0 < kApplyNP ? n_pop - k0 : kApplyNP;
        if (forward)
            k_apply_multi<true><<<grid, kTileThreads, 0, ctx->stream>>>(x0, I, x1, J, d, f, g, 1.0 / median, 1.0 / eps, out_scale,
                                                                        P + (size_t)k0 * n_in, n_in, np, part, n_out);
        else
            k_apply_multi<false><<<grid, kTileThreads, 0, ctx->stream>>>(x0, I, x1, J, d, f, g, 1.0 / median, 1.0 / eps, out_scale,
                                                                         P + (size_t)k0 * n_in, n_in, np, part, n_out);
        k_apply_reduce<<<(unsigned)cdiv((int64_t)np * n_out, 256), 256, 0, ctx->stream>>>(part, n_split, np, n_out,
                                                                                          out + (size_t)k0 * n_out, n_out);
    }
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

// ---- sampling cell pairs from a coupling --------------------------------------------------------------
// Reference: interpolate_with_ot, wot/ot/util.py:140-147: p = tmap / colsum^(1 - frac), flattened, normalised;
// np.random.choice(I*J, p=p, size) = searchsorted(cumsum(p), uniform samples).  The flattened cumulative sum is walked
// hierarchically: the caller finds the ROW of every sample on the row masses (pull-back of the column weights,
// coupling_apply) and hands over (row, remaining mass t); one CTA per sample evaluates that row of the weighted
// coupling in float64 in column order and returns the first column whose running sum exceeds t.
constexpr int kSampleThreads = 256;

__global__ void __launch_bounds__(kSampleThreads)
    k_coupling_sample(const double *__restrict__ x0, long long I, const double *__restrict__ x1, long long J, int d,
                      const double *__restrict__ f, const double *__restrict__ g, double inv_median, double inv_eps, double scale,
                      const double *__restrict__ w, const long long *__restrict__ rows, const double *__restrict__ targets,
                      long long *__restrict__ cols) {
    __shared__ double xi[256];
    __shared__ double seg[kSampleThreads + 1];
    __shared__ int pick;
    const long long i = rows[blockIdx.x];
    const double t = targets[blockIdx.x];
    for (int k = threadIdx.x; k < d; k += kSampleThreads) xi[k] = x0[i * d + k];
    __syncthreads();
    const double fi = f[i];
    const long long per = (J + kSampleThreads - 1) / kSampleThreads;
    const long long j_lo = per * threadIdx.x, j_hi = min(J, j_lo + per);
    auto entry = [&](long long j) {
        double dist = 0.0;
        for (int k = 0; k < d; ++k) {
            const double df = __dsub_rn(xi[k], x1[j * d + k]);
            dist = __dadd_rn(dist, __dmul_rn(df, df));
        }
        return exp((fi + g[j] - dist * inv_median) * inv_eps) * scale * w[j];
    };
    double mine = 0.0;
    for (long long j = j_lo; j < j_hi; ++j) mine += entry(j);
    seg[threadIdx.x + 1] = mine;
    if (threadIdx.x == 0) seg[0] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive prefix of the 256 segment masses, in order; the segment that holds t
        int p = kSampleThreads - 1;
        double run = 0.0;
        for (int q = 0; q < kSampleThreads; ++q) {
            const double m = seg[q + 1];
            seg[q] = run;
            if (run + m > t) {
                p = q;
                break;
            }
            run += m;
        }
        pick = p;
    }
    __syncthreads();
    if (threadIdx.x == pick) {
        double run = seg[pick];
        long long last = j_lo < J ? j_lo : J - 1, found = -1;
        for (long long j = j_lo; j < j_hi; ++j) {
            const double e = entry(j);
            if (e > 0.0) last = j;
            run += e;
            if (run > t) {
                found = j;
                break;
            }
        }
        cols[blockIdx.x] = found >= 0 ? found : last;  // rounding at the very end of a row: its last positive entry
    }
}

int coupling_sample(wotb_ctx *ctx, const double *x0, int64_t I, const double *x1, int64_t J, int d, double median,
                    const double *f, const double *g, double eps, double out_scale, const double *w, const long long *rows,
                    const double *targets, int64_t n_samples, long long *cols) {
    WOTB_REQUIRE(d <= 256, "at most 256 coordinates");
    if (n_samples > 0)
        k_coupling_sample<<<(unsigned)n_samples, kSampleThreads, 0, ctx->stream>>>(x0, I, x1, J, d, f, g, 1.0 / median, 1.0 / eps,
                                                                                   out_scale, w, rows, targets, cols);
    WOTB_CUDA(cudaGetLastError());
    return WOTB_OK;
}

}  // namespace wotb
