// Online-kernel Sinkhorn pass on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// The half-step of optimal_transport.py:133-134 in the online form (see online_pass.cuh) is
//     s_i = sum_j exp2( P_i + Q_j + <X_i, Y_j> )
// with X, Y the coordinates scaled by sqrt(2 log2(e) / (eps median)).  The SIMT kernel spends d FFMA per
// entry on the cross term, which makes it FP32-bound at 1/3 of what the MUFU pipe could do.  Here the
// WHOLE exponent comes out of the tensor cores: every point is a row of 2*kseg fp32 values
//     [ hi(x_0..x_{d-1}), 0.., s_A, s_B | lo(x_0..x_{d-1}), 0.., s_A', s_B' ]        kseg = round_up(d + 2, 8)
// where hi/lo is the 2-term TF32 split (x = hi + lo, both exactly representable in TF32) and the two spare
// K slots carry the offsets:  "out" rows (A operand) hold (1, a1 | 0, a2), "in" rows (B operand) hold
// (b1, 1 | b2, 0) with P_i = a1 + a2 + resid_i and Q_j = b1 + b2 (TF32 pairs, 22 bits).  Three K segments
// (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, the 3xTF32 scheme; the dropped lo.lo term is 2^-22 relative) give
//     D_ij = <X_i, Y_j> + a1 + a2 + b1 + b2
// in the fp32 TMEM accumulator, so the epilogue is ONE MUFU.EX2 and one FADD per entry; resid_i (the part
// of P_i below 22 bits) multiplies the finished row sum in float64.  Tolerance gate (BASELINE.json
// north_star: tensor cores only for the cross term "if the stated tolerance holds"): the exponent error
// is ~1e-5 absolute at eps = 0.05, the same as the fp32 SIMT kernel; tests/test_gpu_parity.py holds both
// to the 1e-4 coupling criterion.
//
// CTA = 256 out rows (two 128-row A blocks resident in shared memory) x one segment of the in side, which
// streams through a ring of 128-row B tiles: one cp.async.bulk per tile, because the operand arrays are
// kept in HBM in exactly the canonical K-major no-swizzle UMMA layout (8-row groups of 16-byte chunks), so
// a tile is a contiguous 32 KB block.  Per B tile the MMA thread issues 2 x 3 x kseg/8 tcgen05.mma
// (M = 128, N = 128, K = 8) into two of four 128-column TMEM accumulators; eight epilogue warps (one
// warpgroup per row block, thread = row) drain them with tcgen05.ld, exp2 and an in-thread sum: no
// shuffles, no shared memory.  L2 traffic is 1 byte per entry; nothing of size I x J exists anywhere.
//
// Warp roles (384 threads): warp 0 lane 0 TMA producer, warp 1 TMEM allocation + MMA issue (lane 0),
// warps 4..11 epilogue.  Pipelines: full/empty per B stage (TMA <-> MMA), acc_full/acc_empty per TMEM
// buffer (MMA <-> epilogue).
#pragma once

#include "online_pass.cuh"

namespace wotb {

constexpr int kTcM = 128;                 // rows per accumulator (UMMA M)
constexpr int kTcRowBlocks = 2;           // A blocks per CTA
constexpr int kTcOut = kTcM * kTcRowBlocks;
constexpr int kTcN = 128;                 // in-side rows per B tile (UMMA N)
constexpr int kTcMaxStages = 4;
constexpr int kTcThreads = 384;
constexpr int kTcEpiWarp0 = 4;            // first epilogue warp (multiple of 4: warp % 4 selects the TMEM lane quadrant)
constexpr int kTcMaxKseg = 40;            // d <= 38
constexpr int kTcTmemCols = 512;
constexpr float kTcPad = -65536.f;        // offset of padded in rows: exp2 underflows to exactly 0
constexpr int kTcSmemLimit = 232448;

// element k (0 <= k < 2*kseg: hi segment then lo segment) of row r; kc = 16-byte chunks per row = kseg / 2
__host__ __device__ __forceinline__ long long tc_index(long long r, int k, int kc) {
    return ((r >> 3) * kc + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
    return __uint_as_float(y);
}

// ---- operand preparation ------------------------------------------------------------------------
// Coordinates (scaled for the current epsilon, optionally centred) into both operand roles of one side.
// The online analogue of rebuilding K (optimal_transport.py:124,:140): runs when need_build is set.
__global__ void k_tc_pack(const double *__restrict__ x, int n, int d, long long rows_pad, int kseg,
                          float *__restrict__ opA, float *__restrict__ opB, const SolveCtrl *ctrl, double scale) {
    if (ctrl && (ctrl->done || !ctrl->need_build)) return;
    const double sc = ctrl ? sqrt(2.0 * ctrl->c2) : scale;
    const int cpr = kseg >> 2;  // 16-byte chunks per segment
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * cpr) return;
    const int r_lo = (int)(idx & 7);
    const int c = (int)((idx >> 3) % cpr);
    const long long r = (idx / (8 * cpr)) * 8 + r_lo;
    float ha[4], la[4], hb[4], lb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = 4 * c + e;
        double val = 0.0;
        if (r < n && k < d) val = sc * x[r * d + k];
        const float h = tf32_rn((float)val);
        const float l = tf32_rn((float)(val - (double)h));
        ha[e] = hb[e] = h;
        la[e] = lb[e] = l;
        if (k == kseg - 2) {  // A: the 1 that picks up b1, b2;  B: b1 (set by k_tc_slots; padding rows stay at kTcPad)
            ha[e] = 1.f, la[e] = 0.f;
            hb[e] = r < n ? 0.f : kTcPad, lb[e] = 0.f;
        } else if (k == kseg - 1) {  // A: a1, a2 (set by k_tc_slots);  B: the 1 that picks up a1, a2
            ha[e] = 0.f, la[e] = 0.f;
            hb[e] = 1.f, lb[e] = 0.f;
        }
    }
    const int kc = kseg >> 1;
    *reinterpret_cast<float4 *>(opA + tc_index(r, 4 * c, kc)) = make_float4(ha[0], ha[1], ha[2], ha[3]);
    *reinterpret_cast<float4 *>(opA + tc_index(r, kseg + 4 * c, kc)) = make_float4(la[0], la[1], la[2], la[3]);
    *reinterpret_cast<float4 *>(opB + tc_index(r, 4 * c, kc)) = make_float4(hb[0], hb[1], hb[2], hb[3]);
    *reinterpret_cast<float4 *>(opB + tc_index(r, kseg + 4 * c, kc)) = make_float4(lb[0], lb[1], lb[2], lb[3]);
}

// Exponent offsets into the spare K slots: the out side's static offsets into its A-role rows (+ the
// float64 residual), the in side's current offsets into its B-role rows.  Runs before every pass.
__global__ void k_tc_slots(const double *__restrict__ off_out, int n_out, float *__restrict__ opA_out,
                           double *__restrict__ resid, const double *__restrict__ off_in, int n_in,
                           float *__restrict__ opB_in, int kseg, const SolveCtrl *ctrl) {
    if (ctrl && ctrl->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int kc = kseg >> 1;
    if (i < n_out) {
        const double al = fmax(off_out[i], -60000.0);
        const float a1 = tf32_rn((float)al);
        const float a2 = tf32_rn((float)(al - (double)a1));
        resid[i] = al - (double)a1 - (double)a2;
        opA_out[tc_index(i, kseg - 1, kc)] = a1;
        opA_out[tc_index(i, 2 * kseg - 1, kc)] = a2;
    }
    if (i < n_in) {
        const double be = fmax(off_in[i], -60000.0);
        const float b1 = tf32_rn((float)be);
        const float b2 = tf32_rn((float)(be - (double)b1));
        opB_in[tc_index(i, kseg - 2, kc)] = b1;
        opB_in[tc_index(i, 2 * kseg - 2, kc)] = b2;
    }
}

// ---- tcgen05 / TMEM primitives --------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a protocol bug must not hang the GPU
    } while (!ok);
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = distance between the two K chunks of one MMA,
// SBO = distance between consecutive 8-row groups (both in bytes, encoded >> 4); version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t desc = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    desc |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    desc |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    desc |= 1ull << 46;
    return desc;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(
            tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

#define WOTB_TMEM_LD32(v, taddr)                                                                                      \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31}, [%32];"                                                                                        \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),     \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),    \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                  \
        : "r"(taddr))
// the registers are in-out operands so that no use of them can be scheduled above the wait
#define WOTB_TMEM_WAIT32(v)                                                                                           \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                     \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),    \
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),           \
                   "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),         \
                   "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),         \
                   "+r"(v[29]), "+r"(v[30]), "+r"(v[31])                                                              \
                 :                                                                                                    \
                 : "memory")

__device__ __forceinline__ float tc_exp2_sum32(const uint32_t (&v)[32]) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        s0 += ex2_approx(__uint_as_float(v[e]));
        s1 += ex2_approx(__uint_as_float(v[e + 1]));
        s2 += ex2_approx(__uint_as_float(v[e + 2]));
        s3 += ex2_approx(__uint_as_float(v[e + 3]));
    }
    return (s0 + s1) + (s2 + s3);
}

struct TcArgs {
    const float *opA;     // out side, A role (UMMA layout, rows padded to kTcOut)
    const float *opB;     // in side, B role (rows padded to kTcOut)
    const double *resid;  // out side: float64 residual of the static offsets
    int out_n;            // valid out entries
    long long out_ld;     // stride of the partial-sum rows (>= padded out rows)
    int kseg;             // K elements per TF32 segment (multiple of 8, >= d + 2)
    int n_stages;         // B ring depth
    int nseg;             // segments of the in side (grid.y)
    int seg_tiles;        // B tiles per segment
    double *part;         // [nseg][out_ld] partial sums
    unsigned int *counters;  // one per out block
    int out_blk0;         // first out block (256 rows) of this launch
    int in_tile0;         // first in tile (128 rows) that is reduced over
    int in_ntiles;        // number of in tiles reduced over
};

// modes as in k_online_pass: 0 half-step, 1 row sums for the gap, 2 coupling row sums, 3 S0 partials,
// 4 partial sums only (row-sharded solves)
template <bool COLPASS>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_online_tc(TcArgs A, SolveVecs V, SolveCtrl *ctrl, int mode, double *rowsum_out) {
    if (mode == 0 || mode == 4) {
        if (!iteration_active(ctrl)) return;
    } else if (mode == 1) {
        if (!gap_rows_wanted(ctrl)) return;
    } else if (mode == 3) {
        if (ctrl->done || !ctrl->need_build || ctrl->solver != WOTB_SOLVER_DUALITY_GAP ||
            ctrl->stage != WOTB_N_STAGES - 1)
            return;
    }
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ int is_last;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int kseg = A.kseg;
    const uint32_t row_bytes = (uint32_t)kseg * 8u;  // 2 * kseg floats
    const uint32_t a_bytes = kTcM * row_bytes, b_bytes = kTcN * row_bytes;
    const int S = A.n_stages;
    unsigned char *sA = tc_smem;
    unsigned char *sB = sA + kTcRowBlocks * a_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + (size_t)S * b_bytes);
    uint64_t *empty = full + kTcMaxStages;
    uint64_t *acc_full = empty + kTcMaxStages;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *a_full = acc_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 1);

    const int out_blk = blockIdx.x + A.out_blk0;
    const long long o0 = (long long)out_blk * kTcOut;
    const int t_begin = A.in_tile0 + blockIdx.y * A.seg_tiles;
    const int t_end = min(A.in_tile0 + A.in_ntiles, t_begin + A.seg_tiles);
    const int n_tiles = max(t_end - t_begin, 0);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);
        }
        mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTcTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    double acc = 0.0;  // epilogue threads: the row sum of this CTA's segment
    if (wid == 0) {
        if (lane == 0 && n_tiles > 0) {
            // ===== TMA producer: the two A blocks once, then the B ring =====
            mbar_expect_tx(a_full, kTcRowBlocks * a_bytes);
            const unsigned char *srcA = reinterpret_cast<const unsigned char *>(A.opA) + (size_t)o0 * row_bytes;
            for (int rb = 0; rb < kTcRowBlocks; ++rb)
                bulk_g2s(sA + rb * a_bytes, srcA + (size_t)rb * a_bytes, a_bytes, a_full);
            const unsigned char *srcB = reinterpret_cast<const unsigned char *>(A.opB);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S, n = t / S;
                mbar_wait_bounded(&empty[s], (uint32_t)((n & 1) ^ 1));
                mbar_expect_tx(&full[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, srcB + (size_t)(t_begin + t) * b_bytes, b_bytes, &full[s]);
            }
        }
    } else if (wid == 1) {
        if (lane == 0 && n_tiles > 0) {
            // ===== MMA issuer =====
            // instruction descriptor: D fp32, A/B TF32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) |
                                   ((uint32_t)(kTcM >> 4) << 24);
            const uint32_t lbo = 128u, sbo = (uint32_t)kseg * 64u;  // 16-byte chunks per row = kseg/2, 128 B each
            const uint32_t lo_off = (uint32_t)(kseg >> 2) * 128u;   // byte offset of the lo segment inside an 8-row group
            const int ksteps = kseg >> 3;
            const uint64_t descA0 = umma_desc(smem_u32(sA), lbo, sbo);
            const uint64_t descB0 = umma_desc(smem_u32(sB), lbo, sbo);
            mbar_wait_bounded(a_full, 0);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S, n = t / S, buf = t & 1;
                mbar_wait_bounded(&acc_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
                mbar_wait_bounded(&full[s], (uint32_t)(n & 1));
                tc_fence_after();
                const uint64_t descB = descB0 + (uint64_t)(((uint32_t)s * b_bytes) >> 4);
#pragma unroll
                for (int rb = 0; rb < kTcRowBlocks; ++rb) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcRowBlocks * kTcN + rb * kTcN);
                    const uint64_t descA = descA0 + (uint64_t)(((uint32_t)rb * a_bytes) >> 4);
                    uint32_t accum = 0;
#pragma unroll
                    for (int seg = 0; seg < 3; ++seg) {  // A_hi.B_hi, A_lo.B_hi, A_hi.B_lo
                        const uint32_t offA = seg == 1 ? lo_off : 0u, offB = seg == 2 ? lo_off : 0u;
                        for (int j = 0; j < ksteps; ++j) {
                            umma_tf32(d_tmem, descA + (uint64_t)((offA + (uint32_t)j * 256u) >> 4),
                                      descB + (uint64_t)((offB + (uint32_t)j * 256u) >> 4), idesc, accum);
                            accum = 1;
                        }
                    }
                }
                umma_commit(&empty[s]);        // the stage is free once these MMAs have read it
                umma_commit(&acc_full[buf]);   // both accumulators of this buffer are complete
            }
        }
    } else if (wid >= kTcEpiWarp0) {
        // ===== epilogue: thread = one out row, exp2 and sum over the tile's 128 columns =====
        const int ew = wid - kTcEpiWarp0;
        const int rb = ew >> 2, q = ew & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t va[32], vb[32];
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            mbar_wait_bounded(&acc_full[buf], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * kTcRowBlocks * kTcN + rb * kTcN);
            float tile_sum;
            WOTB_TMEM_LD32(va, taddr);
            WOTB_TMEM_WAIT32(va);
            WOTB_TMEM_LD32(vb, taddr + 32);
            tile_sum = tc_exp2_sum32(va);
            WOTB_TMEM_WAIT32(vb);
            WOTB_TMEM_LD32(va, taddr + 64);
            tile_sum += tc_exp2_sum32(vb);
            WOTB_TMEM_WAIT32(va);
            WOTB_TMEM_LD32(vb, taddr + 96);
            tile_sum += tc_exp2_sum32(va);
            WOTB_TMEM_WAIT32(vb);
            // every column of this buffer is in registers: hand the accumulators back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            tile_sum += tc_exp2_sum32(vb);
            acc += (double)tile_sum;
        }
        const long long row = o0 + rb * kTcM + q * 32 + lane;
        A.part[(long long)blockIdx.y * A.out_ld + row] = acc;
    }
    tc_fence_before();
    __threadfence();
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcTmemCols)
                     : "memory");
    }
    if (tid == 0) {
        const unsigned int ticket = atomicAdd(&A.counters[out_blk], 1u);
        is_last = ticket == gridDim.y - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last CTA of this out block: sum the segments in order and apply the update -----------------
    double vmax = 0.0;
    if (wid >= kTcEpiWarp0) {
        const long long o = o0 + (tid - kTcEpiWarp0 * 32);
        if (o < A.out_n) {
            double s = 0.0;
            for (int sg = 0; sg < A.nseg; ++sg) s += __ldcg(A.part + (long long)sg * A.out_ld + o);
            s *= exp2(A.resid[o]);
            vmax = online_apply<COLPASS>(mode, (int)o, s, V, ctrl, rowsum_out);
        }
        if (mode == 0) {
            vmax = warp_max(vmax);
            if (lane == 0) atomic_max_nonneg(&ctrl->maxabs, vmax);
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        A.counters[out_blk] = 0;
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

// ---- host side -----------------------------------------------------------------------------------
inline int tc_kseg(int d) { return (int)round_up(d + 2, 8); }
inline bool tc_supported(int d) { return tc_kseg(d) <= kTcMaxKseg; }

struct TcPlan {
    int kseg = 0, n_stages = 0;
    size_t smem = 0;
};

inline TcPlan tc_plan(int d) {
    TcPlan p;
    p.kseg = tc_kseg(d);
    const size_t row_bytes = (size_t)p.kseg * 8;
    const size_t a = (size_t)kTcOut * row_bytes, b = (size_t)kTcN * row_bytes, tail = 256;
    int s = (int)((kTcSmemLimit - a - tail) / b);
    if (s > kTcMaxStages) s = kTcMaxStages;
    p.n_stages = s;
    p.smem = a + (size_t)s * b + tail;
    return p;
}

// Segments of the in side: minimise (waves of CTAs) x (tiles per CTA + fixed per-CTA cost).
inline int tc_segments(int sm_count, int out_blocks, int in_tiles, int *seg_tiles) {
    if (out_blocks < 1) out_blocks = 1;
    if (in_tiles < 1) in_tiles = 1;
    int best = 1;
    double best_cost = 1e300;
    for (int s = 1; s <= in_tiles && s <= 64; ++s) {
        const int per = (int)cdiv(in_tiles, s);
        const int real = (int)cdiv(in_tiles, per);
        if (real != s) continue;
        const int64_t waves = cdiv((int64_t)out_blocks * s, sm_count);
        const double cost = (double)waves * (per + 4.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = s;
        }
    }
    *seg_tiles = (int)cdiv(in_tiles, best);
    return best;
}

inline int tc_configure(const TcPlan &plan) {
    static size_t configured = 0;
    if (configured < plan.smem) {
        WOTB_CUDA(cudaFuncSetAttribute(k_online_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        WOTB_CUDA(cudaFuncSetAttribute(k_online_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        configured = plan.smem;
    }
    return WOTB_OK;
}

}  // namespace wotb

namespace wotb {

// Test / measurement entry: sums[i] = sum_j exp2(off_out[i] + off_in[j] + scale^2 <x_out_i, x_in_j>) with either
// pass kernel (impl 0: SIMT FP32, impl 1: tcgen05), `reps` timed launches after one warm-up.
__global__ void k_pad_offsets(const double *__restrict__ src, int n, double *__restrict__ dst, long long n_pad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) dst[i] = i < n ? src[i] : -INFINITY;
}

int online_rowsums(wotb_ctx *ctx, const double *x_out, int64_t n_out, const double *x_in, int64_t n_in, int d, double scale,
                   const double *off_out, const double *off_in, int impl, int reps, double *sums, double *ms_per_pass) {
    WOTB_REQUIRE(ctx && x_out && x_in && off_out && off_in && sums, "NULL argument");
    WOTB_REQUIRE(n_out >= 1 && n_in >= 1 && d >= 1 && reps >= 1, "bad sizes");
    WOTB_REQUIRE(impl == 0 || (impl == 1 && tc_supported(d)), "impl must be 0 (SIMT) or 1 (tcgen05, d <= 38)");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    wotb_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.epsilon = 0.05, prm.lambda1 = 1, prm.lambda2 = 50, prm.epsilon0 = 1, prm.tau = INFINITY, prm.tolerance = 1e-8;
    prm.max_iter = INFINITY, prm.batch_size = 5, prm.solver = WOTB_SOLVER_DUALITY_GAP;
    SolveCtrl h;
    WOTB_TRY(init_ctrl(&prm, n_out, n_in, &h, 1.0));
    h.batch_iters = 1 << 30;
    h.c2 = 0.5 * scale * scale;  // k_online_scale multiplies by sqrt(2 c2)
    WOTB_TRY(ctx->ctrl.reserve(sizeof(SolveCtrl)));
    SolveCtrl *d_ctrl = ctx->ctrl.as<SolveCtrl>();
    WOTB_CUDA(cudaMemcpyAsync(d_ctrl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    SolveVecs V;
    memset(&V, 0, sizeof(V));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    float ms = 0.f;
    if (impl == 1) {
        const TcPlan plan = tc_plan(d);
        WOTB_TRY(tc_configure(plan));
        const int64_t po = round_up(n_out, kTcOut), pi = round_up(n_in, kTcOut);
        const size_t row_bytes = (size_t)plan.kseg * 8;
        const int out_blocks = (int)(po / kTcOut), in_tiles = (int)cdiv(n_in, kTcN);
        int seg_tiles = 0;
        const int nseg = tc_segments(ctx->sm_count, out_blocks, in_tiles, &seg_tiles);
        const size_t o_ao = take(po * row_bytes), o_bo = take(po * row_bytes), o_ai = take(pi * row_bytes),
                     o_bi = take(pi * row_bytes), o_res = take(po * 8), o_part = take((size_t)nseg * po * 8),
                     o_cnt = take((size_t)out_blocks * 4 + 64);
        WOTB_TRY(ctx->onl.reserve(off));
        char *ob = ctx->onl.as<char>();
        float *Ao = (float *)(ob + o_ao), *Bo = (float *)(ob + o_bo), *Ai = (float *)(ob + o_ai), *Bi = (float *)(ob + o_bi);
        double *resid = (double *)(ob + o_res), *part = (double *)(ob + o_part);
        unsigned int *cnt = (unsigned int *)(ob + o_cnt);
        WOTB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)out_blocks * 4, st));
        const int cpr = plan.kseg / 4;
        k_tc_pack<<<(unsigned)cdiv(po * cpr, 256), 256, 0, st>>>(x_out, (int)n_out, d, po, plan.kseg, Ao, Bo, nullptr, scale);
        k_tc_pack<<<(unsigned)cdiv(pi * cpr, 256), 256, 0, st>>>(x_in, (int)n_in, d, pi, plan.kseg, Ai, Bi, nullptr, scale);
        k_tc_slots<<<(unsigned)cdiv(n_out > n_in ? n_out : n_in, 256), 256, 0, st>>>(off_out, (int)n_out, Ao, resid, off_in,
                                                                                     (int)n_in, Bi, plan.kseg, nullptr);
        TcArgs A;
        A.opA = Ao, A.opB = Bi, A.resid = resid, A.out_n = (int)n_out, A.out_ld = po, A.kseg = plan.kseg;
        A.n_stages = plan.n_stages, A.nseg = nseg, A.seg_tiles = seg_tiles, A.part = part, A.counters = cnt;
        A.out_blk0 = 0, A.in_tile0 = 0, A.in_ntiles = in_tiles;
        const dim3 grid(out_blocks, nseg);
        k_online_tc<false><<<grid, kTcThreads, plan.smem, st>>>(A, V, d_ctrl, 4, sums);
        WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
        for (int r = 0; r < reps; ++r) k_online_tc<false><<<grid, kTcThreads, plan.smem, st>>>(A, V, d_ctrl, 4, sums);
        WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    } else {
        const int64_t ldo = round_up(n_out, kOnTile), ldi = round_up(n_in, kOnTile);
        const int dp = (int)round_up(d, 4);
        const int tiles_o = (int)(ldo / kOnTile), tiles_i = (int)(ldi / kOnTile);
        int nseg = (int)cdiv((int64_t)ctx->sm_count * 2, tiles_o);
        if (nseg > tiles_i) nseg = tiles_i;
        if (nseg < 1) nseg = 1;
        const int seg_tiles = (int)cdiv(tiles_i, nseg);
        nseg = (int)cdiv(tiles_i, seg_tiles);
        const size_t o_xt = take((size_t)dp * ldo * 4), o_yt = take((size_t)dp * ldi * 4), o_po = take(ldo * 8),
                     o_pi = take(ldi * 8), o_part = take((size_t)nseg * ldo * 8), o_cnt = take((size_t)tiles_o * 4 + 64);
        WOTB_TRY(ctx->onl.reserve(off));
        char *ob = ctx->onl.as<char>();
        float *XT = (float *)(ob + o_xt), *YT = (float *)(ob + o_yt);
        double *po = (double *)(ob + o_po), *pi = (double *)(ob + o_pi), *part = (double *)(ob + o_part);
        unsigned int *cnt = (unsigned int *)(ob + o_cnt);
        WOTB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)tiles_o * 4, st));
        k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldo, 256), 256, 0, st>>>(x_out, (int)n_out, d, XT, ldo, dp, d_ctrl, 0);
        k_online_scale<<<(unsigned)cdiv((int64_t)dp * ldi, 256), 256, 0, st>>>(x_in, (int)n_in, d, YT, ldi, dp, d_ctrl, 0);
        k_pad_offsets<<<(unsigned)cdiv(ldo, 256), 256, 0, st>>>(off_out, (int)n_out, po, ldo);
        k_pad_offsets<<<(unsigned)cdiv(ldi, 256), 256, 0, st>>>(off_in, (int)n_in, pi, ldi);
        const size_t smem = (size_t)3 * kOnChunk * kOnTile * 4;
        WOTB_CUDA(cudaFuncSetAttribute(k_online_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OnlineArgs A;
        A.out = {XT, ldo, po, (int)n_out};
        A.in = {YT, ldi, pi, (int)n_in};
        A.dp = dp, A.nseg = nseg, A.seg_tiles = seg_tiles, A.part = part, A.counters = cnt;
        A.out_tile0 = 0, A.in_tile0 = 0, A.in_ntiles = tiles_i;
        const dim3 grid(tiles_o, nseg);
        k_online_pass<false><<<grid, kOnThreads, smem, st>>>(A, V, d_ctrl, 4, sums);
        WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
        for (int r = 0; r < reps; ++r) k_online_pass<false><<<grid, kOnThreads, smem, st>>>(A, V, d_ctrl, 4, sums);
        WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
    }
    WOTB_CUDA(cudaStreamSynchronize(st));
    WOTB_CUDA(cudaGetLastError());
    WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_per_pass) *ms_per_pass = ms / reps;
    return WOTB_OK;
}

}  // namespace wotb
