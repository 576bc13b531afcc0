// Online-kernel Sinkhorn pass on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// The half-step of optimal_transport.py:133-134 in the online form (see online_pass.cuh) is
//     s_i = sum_j exp2( P_i + Q_j + <X_i, Y_j> )
// with X, Y the coordinates scaled by sqrt(2 log2(e) / (eps median)).  The SIMT kernel spends d FFMA per
// entry on the cross term, which makes it FP32-bound at 1/5 of what the MUFU pipe could do.  Here the
// WHOLE exponent comes out of the tensor cores.  Every point is a row of 3*kseg fp16 values,
// kseg = round_up(d + 2, 16), built from the 2-term split x = hi + 2^-11 lo' (hi = fp16(x),
// lo' = fp16((x - hi) 2^11): 22 bits, products exact in the fp32 accumulator):
//     A role ("out" rows):  [ hi | lo' | hs ]      B role ("in" rows):  [ hi | hs | lo' ]      hs = 2^-11 hi
// so one plain K = 3*kseg GEMM gives hi.hi + 2^-11 (lo'.hi + hi.lo'); the dropped lo.lo term is 2^-22
// relative.  (Measured on B200: a tcgen05.mma costs ~64 + N/2 clocks whatever the operand type, so the fp16
// K = 16 instruction does the work of two TF32 K = 8 ones; with 3xTF32 the pass was tensor-bound at 0.48 of
// the MUFU peak.)  The two spare K slots of every segment carry the offsets: A rows hold (1, a1 | 0, a2' |
// 2^-11, 0), B rows hold (b1, 1 | 0, 2^-11 | b2', 0), P_i = a1 + 2^-11 a2' + resid_i, Q_j = b1 + 2^-11 b2',
// so
//     D_ij = <X_i, Y_j> + P_i - resid_i + Q_j
// lands in the fp32 TMEM accumulator and the epilogue is ONE MUFU.EX2 and one FADD per entry; resid_i (the
// part of P_i below 22 bits) multiplies the finished row sum in float64.  Tolerance gate (BASELINE.json
// north_star: tensor cores only for the cross term "if the stated tolerance holds"): the exponent error is
// ~1e-5 absolute at eps = 0.05, the same as the fp32 SIMT kernel; tests/test_gpu_parity.py holds both to
// the 1e-4 coupling criterion.
//
// CTA (640 threads) = 256 out rows (two 128-row A blocks, double buffered in shared memory) x a contiguous range of
// (out block, in tile) work units dealt "stream-K" style over a one-wave grid.  The in side streams through a ring
// of 128-row B tiles: one cp.async.bulk per tile, because the operand arrays are kept in HBM in exactly the
// canonical K-major no-swizzle UMMA layout (8-row groups of 16-byte chunks), so a tile is a contiguous 24 KB block.
// Per B tile each of the two MMA warps issues 3*kseg/16 tcgen05.mma (kind::f16, M = 128, N = 128, K = 16) into its
// row block's half of a double-buffered 2 x 256-column TMEM accumulator (all 512 columns); 16 epilogue warps
// (8 per row block: thread = row x 64 columns, EW = 8; or 8 warps, thread = row x 128 columns, EW = 4) drain
// them with tcgen05.ld 32x32b.x32, exp2 and an in-thread sum: no shuffles, no shared memory in the tile loop.
// L2 traffic is 0.75 byte per entry; nothing of size I x J exists anywhere.
//
// Warp roles: warp 0 lane 0 TMA producer, warp 1 TMEM allocation, warps 1..2 MMA issue (one per row block, one
// elected lane), warp 3 idle, warps 4.. epilogue.  Pipelines: full/empty per B stage (TMA <-> MMA), a_full/a_empty
// per A buffer, acc_full/acc_empty per (TMEM buffer, row block) (MMA <-> epilogue).
#pragma once

#include <cuda_fp16.h>
#include <stdlib.h>

#include "online_pass.cuh"

namespace wotb {

constexpr int kTcM = 128;                 // rows per accumulator (UMMA M)
constexpr int kTcRowBlocks = 2;           // A blocks per CTA
constexpr int kTcOut = kTcM * kTcRowBlocks;  // row padding of every operand array
constexpr int kTcN = 128;                 // in-side rows per B tile (UMMA N)
constexpr int kTcAccCols = kTcRowBlocks * kTcN;  // TMEM columns per accumulator buffer (two buffers = all 512)
constexpr int kTcTail = 4096;             // barriers + reduction scratch after the operand stages
constexpr int kTcMaxStages = 6;
constexpr int kTcEpiWarp0 = 4;            // first epilogue warp (multiple of 4: warp % 4 selects the TMEM lane quadrant)
constexpr int kTcMaxKseg = 48;            // d <= 46
constexpr int kTcTmemCols = 512;
constexpr int kTcSmemLimit = 232448;

// ---- operand preparation ------------------------------------------------------------------------
// Coordinates (scaled for the current epsilon) into both operand roles of one side.
// The online analogue of rebuilding K (optimal_transport.py:124,:140): runs when need_build is set.
//
// nseg == 3 (default): x = hi + 2^-11 lo' with hi = fp16(x):   A [ hi | lo' | hs ]   B [ hi | hs | lo' ],  hs = 2^-11 hi.
// nseg == 6 (precise): x = h + m + l, h = rint(x 2^q) 2^-q on the grid of tc_grid_q, m' = fp16((x - h) 2^(q+1)),
//   l' = fp16((x - h - m) 2^(q+13)), all three limbs of 11 bits:
//       A [ h | m' | h 2^-(q+1) | m'           | l' 2^-6     | h 2^-(q+7) ]
//       B [ h | h 2^-(q+1) | m' | m' 2^-(2q+2) | h 2^-(q+7) | l' 2^-6     ]
//   = h.h + m.h + h.m + m.m + l.h + h.l; the dropped m.l, l.l terms are below 2^-(2q+14) |x| per dimension.  The
//   leading segment (with the integer parts of both offsets in its two spare slots) is a sum of multiples of
//   2^-2q below 2^24 quanta: exact in the fp32 accumulator, so the cancellation of the large terms c2 |x|^2,
//   c2 |y|^2, 2 c2 <x, y> costs nothing, and the remaining segments add small numbers to a small number.
__global__ void k_tc_pack(const double *__restrict__ x, int n, int d, long long rows_pad, int kseg, int nseg,
                          __half *__restrict__ opA, __half *__restrict__ opB, const SolveCtrl *ctrl, double scale,
                          const TcGeo *__restrict__ geo) {
    if (ctrl && (ctrl->done || !ctrl->need_build)) return;
    const double sc = ctrl ? sqrt(2.0 * ctrl->c2) : scale;
    const int cps = kseg >> 3;  // 16-byte chunks per segment
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * cps) return;
    const int r_lo = (int)(idx & 7);
    const int c = (int)((idx >> 3) % cps);
    const long long r = (idx / (8 * cps)) * 8 + r_lo;
    const __half zero = __float2half(0.f), one = __float2half(1.f), tiny = __float2half(kTcLoInv);
    const int kc = nseg * cps;
    if (nseg == 3) {
        __align__(16) __half hi[8], lo[8], hs[8];
        __align__(16) __half bhi[8], blo[8], bhs[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * c + e;
            double val = 0.0;
            if (r < n && k < d) val = sc * x[r * d + k];
            tc_split(val, hi[e], lo[e]);
            hs[e] = __float2half(__half2float(hi[e]) * kTcLoInv);
            bhi[e] = hi[e], blo[e] = lo[e], bhs[e] = hs[e];
            if (k == kseg - 2) {
                // A: (1 | 0 | 2^-11) picks up b1 and b2';  B: (b1 | 0 | b2') set by k_tc_slots, padding rows stay at kTcPad
                hi[e] = one, lo[e] = zero, hs[e] = tiny;
                bhi[e] = r < n ? zero : __float2half(kTcPad), bhs[e] = zero, blo[e] = zero;
            } else if (k == kseg - 1) {
                // A: (a1 | a2' | 0) set by k_tc_slots;  B: (1 | 2^-11 | 0) picks up a1 and a2'
                hi[e] = zero, lo[e] = zero, hs[e] = zero;
                bhi[e] = one, bhs[e] = tiny, blo[e] = zero;
            }
        }
        *reinterpret_cast<uint4 *>(opA + tc_index(r, 8 * c, kc)) = *reinterpret_cast<const uint4 *>(hi);
        *reinterpret_cast<uint4 *>(opA + tc_index(r, kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(lo);
        *reinterpret_cast<uint4 *>(opA + tc_index(r, 2 * kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(hs);
        *reinterpret_cast<uint4 *>(opB + tc_index(r, 8 * c, kc)) = *reinterpret_cast<const uint4 *>(bhi);
        *reinterpret_cast<uint4 *>(opB + tc_index(r, kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(bhs);
        *reinterpret_cast<uint4 *>(opB + tc_index(r, 2 * kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(blo);
        return;
    }
    const int q = tc_grid_q(sc, *geo);
    const double up = ldexp(1.0, q), dn = ldexp(1.0, -q);
    const double scc = sc * sqrt(1.0 + kTcTruncComp);  // truncation compensation, see kTcTruncComp
    __align__(16) __half sa[6][8], sb[6][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = 8 * c + e;
        double val = 0.0;
        if (r < n && k < d) val = scc * x[r * d + k];
        const double h = rint(val * up) * dn;
        const double r1 = val - h;
        const __half mh = __double2half(ldexp(r1, q + 1));
        const double m = ldexp((double)__half2float(mh), -(q + 1));
        const __half lh = __double2half(ldexp(r1 - m, q + 13));
        const __half hh = __double2half(h);                              // exact: |h| 2^q < 2^10
        const __half h1 = __double2half(ldexp(h, -(q + 1)));             // exact (multiples of 2^-(2q+1) >= 2^-24)
        const __half h7 = __double2half(ldexp(h, -(q + 7)));             // exact for q <= 8
        const __half m2 = __double2half(ldexp((double)__half2float(mh), -(2 * q + 2)));
        const __half l6 = __double2half(ldexp((double)__half2float(lh), -6));
        sa[0][e] = hh, sb[0][e] = hh;
        sa[1][e] = mh, sb[1][e] = h1;
        sa[2][e] = h1, sb[2][e] = mh;
        sa[3][e] = mh, sb[3][e] = m2;
        sa[4][e] = l6, sb[4][e] = h7;
        sa[5][e] = h7, sb[5][e] = l6;
        if (k == kseg - 2) {
            // in-side offsets b1 (integer) | b2 | b3': A holds the multipliers, B the values (k_tc_slots / finishing code)
#pragma unroll
            for (int sg = 0; sg < 6; ++sg) sa[sg][e] = zero, sb[sg][e] = zero;
            sa[0][e] = one, sa[1][e] = one, sa[2][e] = tiny;
            if (r >= n) sb[0][e] = __float2half(kTcPad);
        } else if (k == kseg - 1) {
            // out-side offsets a1 | a2 | a3': A holds the values, B the multipliers
#pragma unroll
            for (int sg = 0; sg < 6; ++sg) sa[sg][e] = zero, sb[sg][e] = zero;
            sb[0][e] = one, sb[1][e] = one, sb[2][e] = tiny;
        }
    }
#pragma unroll
    for (int sg = 0; sg < 6; ++sg) {
        *reinterpret_cast<uint4 *>(opA + tc_index(r, sg * kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(sa[sg]);
        *reinterpret_cast<uint4 *>(opB + tc_index(r, sg * kseg + 8 * c, kc)) = *reinterpret_cast<const uint4 *>(sb[sg]);
    }
}

// Exponent offsets into the spare K slots: the out side's static offsets into its A-role rows (+ the
// float64 residual, stored as the factor 2^residual), the in side's current offsets into its B-role rows.  During the iterations the in-side
// slots are kept current by the finishing code (tc_store_in_offset); this kernel runs when everything changed.
__global__ void k_tc_slots(const double *__restrict__ off_out, int n_out, __half *__restrict__ opA_out,
                           double *__restrict__ resid, const double *__restrict__ off_in, int n_in,
                           __half *__restrict__ opB_in, int kseg, int nseg, const SolveCtrl *ctrl, int gate) {
    // gate 0: unless the solve is done; 1: only when need_build is set (the offsets were rewritten by an absorption,
    // an epsilon change or the initialisation); 3: only for the S0 pass of the final stage; -1: always
    if (ctrl && gate >= 0) {
        if (ctrl->done) return;
        if (gate == 1 && !ctrl->need_build) return;
        if (gate == 3 && (!ctrl->need_build || ctrl->solver != WOTB_SOLVER_DUALITY_GAP || ctrl->stage != WOTB_N_STAGES - 1))
            return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int kc = nseg * (kseg >> 3);
    if (i < n_out) {
        const double al = fmax(off_out[i], (double)kTcPad);
        if (nseg == 3) {
            __half a1, a2;
            tc_split(al, a1, a2);
            resid[i] = exp2(al - (double)__half2float(a1) - (double)__half2float(a2) * (double)kTcLoInv);
            opA_out[tc_index(i, kseg - 1, kc)] = a1;
            opA_out[tc_index(i, 2 * kseg - 1, kc)] = a2;
        } else {
            __half a1, a2, a3;
            // the residual is applied to the finished sum in float64 and is not subject to the accumulator's truncation
            resid[i] = exp2(tc_split3(fmax(al * (1.0 + kTcTruncComp), (double)kTcPad), a1, a2, a3) / (1.0 + kTcTruncComp));
            opA_out[tc_index(i, kseg - 1, kc)] = a1;
            opA_out[tc_index(i, 2 * kseg - 1, kc)] = a2;
            opA_out[tc_index(i, 3 * kseg - 1, kc)] = a3;
        }
    }
    if (i < n_in) tc_store_in_offset(opB_in, i, off_in[i], kseg, nseg);
}

// max |x_ik|, max |x_i|^2 over one side into the solve's TcGeo (non-negative doubles order like their bit patterns, so
// the atomic maximum is exact and order independent)
__global__ void k_tc_geo(const double *__restrict__ x, int n, int d, TcGeo *geo, int side) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ma = 0.0, n2 = 0.0;
    if (i < n) {
        for (int k = 0; k < d; ++k) {
            const double v = x[(long long)i * d + k];
            ma = fmax(ma, fabs(v));
            n2 = fma(v, v, n2);
        }
    }
    ma = warp_max(ma);
    n2 = warp_max(n2);
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&geo->max_abs), ma);
        atomic_max_nonneg(reinterpret_cast<unsigned long long *>(side == 0 ? &geo->max_n2x : &geo->max_n2y), n2);
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

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// warp-uniform wait: lane 0 polls, the warp reconverges behind it
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity) {
    mbar_wait_bounded(bar, parity);
    __syncwarp();
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(
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

// exp2 on the FMA pipe: round-to-nearest split x = n + r (magic-number add), degree-5 minimax polynomial for 2^r on
// [-0.5, 0.5] (max relative error 2.4e-7 in fp32 Horner form, the same as MUFU.EX2), 2^n inserted into the exponent
// field with one integer multiply-add.  The MUFU pipe (16 ex2 per clock and SM) is the roof of this kernel while
// the FMA pipes idle, so a quarter of every 32-column chunk is evaluated this way (the FlashAttention-4 trick).
__device__ __forceinline__ float exp2_fma(float x) {
    x = fminf(fmaxf(x, -126.f), 126.f);
    const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float r = x - (t - 12582912.f);
    float p = 0.0013276470126584172f;
    p = fmaf(p, r, 0.009675540961325169f);
    p = fmaf(p, r, 0.05550713464617729f);
    p = fmaf(p, r, 0.24022120237350464f);
    p = fmaf(p, r, 0.6931469440460205f);
    p = fmaf(p, r, 1.0000001192092896f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// ---- packed fp32x2 arithmetic (sm_100a: FADD2 / FFMA2 take one issue slot for two lanes of work) -----------------
// The epilogue is bound by instruction issue together with the MUFU pipe (ncu, profiles/r1e: XU 59 %, issue 57 %,
// 5.7 thread instructions per entry with scalar code), so everything that is not a MUFU.EX2 is done two entries at
// a time: the running sums (FADD2) and the FMA-pipe exp2 (range reduction and Horner steps as FADD2 / FFMA2).
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float x) { return pack2(x, x); }

// exp2 of two entries on the FMA pipe: the arithmetic of exp2_fma, two lanes per instruction.
// 2 FMNMX + 3 FADD2/FFMA2 (split) + 5 FFMA2 (Horner) + 2 integer ops = 12 issue slots per pair (scalar: 26).
__device__ __forceinline__ f32x2 exp2_fma2(float x0, float x1) {
    x0 = fmaxf(x0, -126.f);  // padding rows carry a huge negative offset; exponents never approach +127 (the MUFU
    x1 = fmaxf(x1, -126.f);  // entries of the same row would overflow first)
    const f32x2 x = pack2(x0, x1);
    const f32x2 t = fadd2(x, splat2(12582912.f));
    const f32x2 n = fadd2(t, splat2(-12582912.f));
    const f32x2 r = ffma2(n, splat2(-1.f), x);
    f32x2 p = ffma2(splat2(0.0013276470126584172f), r, splat2(0.009675540961325169f));
    p = ffma2(p, r, splat2(0.05550713464617729f));
    p = ffma2(p, r, splat2(0.24022120237350464f));
    p = ffma2(p, r, splat2(0.6931469440460205f));
    p = ffma2(p, r, splat2(1.0000001192092896f));
    float p0, p1, t0, t1;
    unpack2(p, p0, p1);
    unpack2(t, t0, t1);
    return pack2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)),
                 __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}

// Of every 16 consecutive entries, kTcFmaPairs pairs are evaluated on the FMA pipe and the rest by MUFU.EX2 (the
// FlashAttention-4 trick).  With packed arithmetic a pair costs 12 issue slots and a MUFU entry 1.5; the MUFU pipe
// takes 8 clocks per warp instruction.  Measured on B200 (profiles/r1f_pairs.txt, 12486 x 12405): 0 pairs 54.9 us, 2 pairs
// 49.3 us, 3 pairs 52.4 us, 4 pairs 55.5 us, 5 pairs 60.2 us per pass -> 2 pairs of 16 (a quarter of the entries).
#ifndef WOTB_TC_FMA_PAIRS
#define WOTB_TC_FMA_PAIRS 2
#endif
constexpr int kTcFmaPairs = WOTB_TC_FMA_PAIRS;

// true when the pair (e, e + 1), e even, of a 16-entry group goes to the FMA pipe: the pairs are spread out so that
// MUFU and FMA work interleave in program order
__host__ __device__ constexpr bool tc_pair_on_fma(int pair) {
    return kTcFmaPairs >= 8 ? true
         : kTcFmaPairs <= 0 ? false
         : ((pair + 1) * kTcFmaPairs / 8) != (pair * kTcFmaPairs / 8);
}

// measurement only (PROF build, dbg bit2): the 32 values as they come out of TMEM, summed with 16 FADD2 -- what the
// epilogue costs when the exponentials are free.  Measured on B200 (profiles/r3c_tc_ld_only.txt): 1081 clocks per
// 256 x 128 unit, bound by the MMA warps' issue chain; TMEM is read at >= 121 B/clk and is NOT what limits the pass.
__device__ __forceinline__ float tc_plain_sum32(const uint32_t (&v)[32]) {
    f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        s0 = fadd2(s0, pack2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
        s1 = fadd2(s1, pack2(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])));
    }
    float lo, hi;
    unpack2(fadd2(s0, s1), lo, hi);
    return lo + hi;
}

__device__ __forceinline__ float tc_exp2_sum32(const uint32_t (&v)[32]) {
    f32x2 s0 = 0ull, s1 = 0ull;  // (0.f, 0.f)
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        const float a0 = __uint_as_float(v[e]), a1 = __uint_as_float(v[e + 1]);
        const float b0 = __uint_as_float(v[e + 2]), b1 = __uint_as_float(v[e + 3]);
        const f32x2 ea = tc_pair_on_fma((e >> 1) & 7) ? exp2_fma2(a0, a1) : pack2(ex2_approx(a0), ex2_approx(a1));
        const f32x2 eb = tc_pair_on_fma(((e >> 1) + 1) & 7) ? exp2_fma2(b0, b1) : pack2(ex2_approx(b0), ex2_approx(b1));
        s0 = fadd2(s0, ea);
        s1 = fadd2(s1, eb);
    }
    float lo, hi;
    unpack2(fadd2(s0, s1), lo, hi);
    return lo + hi;
}

struct TcArgs {
    const __half *opA;    // out side, A role (UMMA layout, rows padded to kTcOut)
    const __half *opB;    // in side, B role (rows padded to kTcOut)
    const double *resid;  // out side: 2^(float64 residual of the static offsets), the factor of the finished sum
    int out_n;            // valid out entries
    long long out_ld;     // stride of the partial-sum slots (>= padded out rows)
    int n_blocks;         // out blocks (256 rows) of this launch
    int out_blk0;         // first out block of this launch
    int in_tile0;         // first in tile (128 rows) that is reduced over
    int in_ntiles;        // number of in tiles reduced over
    int n_stages;         // B ring depth
    double *part;         // [slots][out_ld] partial sums, one slot per (CTA that contributes to an out block, column half)
    unsigned int *counters;  // kTcRowBlocks * 4 per out block: one per group of 32 rows (tc_publish)
    int dbg;              // measurement only: bit0 skip the MMAs, bit1 skip the epilogue work, bit2 TMEM loads + plain sums
                          // (no exponentials), bit3 write cycle counters
    long long *prof;      // [CTA][epilogue warp][4]: cycles total, waiting for accumulators, waiting for tcgen05.ld, tiles
};

// The CTA that owns work unit g when T units are dealt to G CTAs in contiguous ranges [c T / G, (c + 1) T / G).
__host__ __device__ __forceinline__ int tc_cta_of_unit(long long g, long long T, int G) {
    return (int)(((g + 1) * G + T - 1) / T - 1);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- shared-memory barriers by 32-bit shared address: the loops below keep every address in a register instead of
// re-deriving it from a generic pointer each time (profiles/r2a: ~100 of the 282 instructions of the epilogue loop
// were address arithmetic, clock reads of the disabled profiler and barrier bookkeeping) ------------------------------
__device__ __forceinline__ void tcb_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tcb_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tcb_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a protocol bug must not hang the GPU
    } while (!ok);
}
// non-blocking phase test (test_wait never suspends the thread): issued early, consumed later, so that the ~200 clocks a
// barrier query takes on this part are spent under arithmetic instead of in front of it
__device__ __forceinline__ uint32_t tcb_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}
// makes a value opaque to the compiler: it stays in its register instead of being re-derived (S2UR SR_CgaCtaId / ULEA
// chains in front of every barrier access, profiles/r2p)
__device__ __forceinline__ uint32_t tc_keep(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ void tcb_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tcb_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- publish / finish, one WARP at a time ------------------------------------------------------------------------
// A warp owns 32 rows of an out block (x all or half of the columns of every tile).  When its CTA's share of the
// block is complete it stores its 32 partial sums into its own slot (CTA within the block, column half), fences,
// and takes a ticket on the counter of (block, row group).  The warp that draws the last ticket adds the slots --
// per CTA the two column halves first, then the CTAs in order, so the result does not depend on timing -- and
// applies the update to its 32 rows.  No CTA-wide barrier: the other warps of the CTA are already in the next block's
// tiles, and the eight row groups of a block are finished by (up to) eight different warps in parallel.
// (Round 1/2 did this per CTA behind three named barriers: 5.4k clocks per CTA and pass, profiles/r2p.)
template <bool COLPASS, int EW>
__device__ __forceinline__ void tc_publish(const TcArgs &A, const SolveVecs &V, SolveCtrl *ctrl, int mode, double *rowsum_out,
                                           int b, int nt, long long T, int G, int rb, int q, int h, int lane, double acc,
                                           bool async_fence) {
    constexpr int H = EW / 4;  // column halves = contributions per CTA and row
    const int blk = A.out_blk0 + b;
    const long long row = (long long)blk * kTcOut + rb * kTcM + q * 32 + lane;
    const int c_first = tc_cta_of_unit((long long)b * nt, T, G);
    const int c_last = tc_cta_of_unit((long long)(b + 1) * nt - 1, T, G);
    const int n_cta = c_last - c_first + 1;
    A.part[((long long)((int)blockIdx.x - c_first) * H + h) * A.out_ld + row] = acc;
    __threadfence();
    __syncwarp();
    unsigned int *counter = A.counters + ((long long)blk * (kTcRowBlocks * 4) + rb * 4 + q);
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(counter, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != (unsigned int)(n_cta * H - 1)) return;
    __threadfence();
    double vmax = 0.0;
    if (row < A.out_n) {
        double sum = 0.0;
        for (int sl = 0; sl < n_cta; ++sl) {
            const double *p = A.part + (long long)sl * H * A.out_ld + row;
            double t = __ldcg(p);
            if (H == 2) t += __ldcg(p + A.out_ld);
            sum += t;
        }
        sum *= A.resid[row];  // 2^residual, evaluated when the slots were written: one transcendental less in the tail
        vmax = online_apply<COLPASS>(mode, (int)row, sum, V, ctrl, rowsum_out);
        // persistent batch kernel: the offset slots just written are read by other CTAs' TMA loads after the grid barrier
        if (async_fence) asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    if (mode == 0) {
        vmax = warp_max(vmax);
        if (lane == 0) atomic_max_nonneg(&ctrl->maxabs, vmax);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        *counter = 0;
        if (mode == 0 && COLPASS) {
            const unsigned int done = atomicAdd(&ctrl->col_tiles_done, 1u);
            if (done == (unsigned int)(A.n_blocks * (kTcRowBlocks * 4) - 1)) {
                __threadfence();
                ctrl->col_tiles_done = 0;
                close_iteration(ctrl);
            }
        }
    }
}

// One pass = (out blocks) x (in tiles) work units of 256 x 128 entries, dealt to the CTAs of a one-wave
// grid in contiguous ranges (block-major), "stream-K" style: every SM gets the same number of units whatever
// the shape, and the per-CTA fixed cost (TMEM allocation, barrier setup, pipeline fill) is paid once.  A CTA's
// range may cross out-block boundaries; the A blocks are double buffered (3-segment operands) so the switch costs
// nothing.  Every epilogue warp writes the partial sums of its 32 rows into its own slot and the warp that arrives
// last at a (block, 32-row group) adds the slots in a fixed order and applies the update (tc_publish), so the result
// does not depend on timing.
//
// modes as in k_online_pass: 0 half-step, 1 row sums for the gap, 2 coupling row sums, 3 S0 partials,
// 4 partial sums only (row-sharded solves).  EW = epilogue warps per row block (4: thread = row x 128 columns,
// 8: thread = row x 64 columns).  NSEG = K segments per operand row (3, or 6 in precise mode: twice the MMAs per
// tile, A single buffered).  PROF: measurement build with cycle counters and the dbg switches of TcArgs.
template <bool COLPASS, int KSEG, int EW, int NSEG, bool PROF>
__global__ void __launch_bounds__(128 + kTcRowBlocks * EW * 32, 1)
    k_online_tc(TcArgs A, SolveVecs V, SolveCtrl *ctrl, int mode, double *rowsum_out) {
    constexpr int RB = kTcRowBlocks, NT = kTcN;
    constexpr int NCH = 4 / (EW / 4);  // 32-column chunks per tile and warp
    constexpr int kseg = KSEG;
    constexpr int NABUF = NSEG == 3 ? 2 : 1;
    constexpr uint32_t row_bytes = (uint32_t)kseg * NSEG * 2u;
    constexpr uint32_t a_bytes = kTcM * row_bytes, b_bytes = NT * row_bytes;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int S = A.n_stages;
    // shared memory map (32-bit shared addresses): barriers + reduction scratch (kTcTail bytes, at fixed offsets so
    // that a barrier address is the base plus a constant) | A blocks | B ring
    const uint32_t smem0 = smem_u32(tc_smem);
    const uint32_t bars = smem0;
    const uint32_t sA = smem0 + kTcTail, sB = sA + NABUF * RB * a_bytes;
    const uint32_t b_full = bars, b_empty = bars + 8 * kTcMaxStages;
    const uint32_t b_accf = b_empty + 8 * kTcMaxStages;  // [buffer][row block]
    const uint32_t b_acce = b_accf + 8 * 2 * RB;
    const uint32_t b_af = b_acce + 8 * 2 * RB, b_ae = b_af + 16;
    unsigned char *bars_p = tc_smem + (bars - smem0);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars_p + 8 * (2 * kTcMaxStages + 4 * RB + 4));

    const int nt = A.in_ntiles;
    const long long T = (long long)A.n_blocks * nt;
    const int G = gridDim.x;
    const long long g0 = (long long)blockIdx.x * T / G, g1 = (long long)(blockIdx.x + 1) * T / G;
    const int b_first = (int)(g0 / nt), t_first = (int)(g0 - (long long)b_first * nt);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            tcb_init(b_full + 8 * s, 1);
            tcb_init(b_empty + 8 * s, RB);  // one commit per MMA warp
        }
        for (int b = 0; b < 2 * RB; ++b) {
            tcb_init(b_accf + 8 * b, 1);
            tcb_init(b_acce + 8 * b, EW);  // the warps that drain it
        }
        for (int b = 0; b < 2; ++b) {
            tcb_init(b_af + 8 * b, 1);
            tcb_init(b_ae + 8 * b, RB);
        }
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

    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) touched only this CTA's own
    // resources and may have run while the previous kernel in the stream was still draining; from here on the
    // state written by that kernel (SolveCtrl, offsets in the operand slots, scalings) is read.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    bool active = true;
    if (mode == 0 || mode == 4) {
        active = iteration_active(ctrl);
    } else if (mode == 1) {
        active = gap_rows_wanted(ctrl);
    } else if (mode == 3) {
        active = !(ctrl->done || !ctrl->need_build || ctrl->solver != WOTB_SOLVER_DUALITY_GAP ||
                   ctrl->stage != WOTB_N_STAGES - 1);
    }

    if (!active) {
        // nothing to do (batch finished, solver done, ...): fall through to the TMEM release
    } else if (wid == 0) {
        if (lane == 0) {
            // ===== TMA producer: per out-block segment the two A blocks, then the B ring =====
            const unsigned char *srcA = reinterpret_cast<const unsigned char *>(A.opA);
            const unsigned char *srcB = reinterpret_cast<const unsigned char *>(A.opB);
            int s = 0, seg = 0, b = b_first, t = t_first;
            uint32_t empty_par = 1;
            for (long long g = g0; g < g1; ++seg, ++b, t = 0) {
                const int abuf = NABUF == 2 ? (seg & 1) : 0;
                const uint32_t a_par = (uint32_t)(NABUF == 2 ? (seg >> 1) & 1 : seg & 1);
                tcb_wait(b_ae + 8 * abuf, a_par ^ 1u);
                tcb_expect_tx(b_af + 8 * abuf, RB * a_bytes);
                const unsigned char *blockA = srcA + (size_t)(A.out_blk0 + b) * (RB * a_bytes);
                for (int rb = 0; rb < RB; ++rb)
                    tcb_bulk_g2s(sA + (abuf * RB + rb) * a_bytes, blockA + (size_t)rb * a_bytes, a_bytes, b_af + 8 * abuf);
                const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
                for (int k = 0; k < n_in_seg; ++k) {
                    tcb_wait(b_empty + 8 * s, empty_par);
                    tcb_expect_tx(b_full + 8 * s, b_bytes);
                    tcb_bulk_g2s(sB + (uint32_t)s * b_bytes, srcB + (size_t)(A.in_tile0 + t + k) * b_bytes, b_bytes,
                                 b_full + 8 * s);
                    if (++s == S) {
                        s = 0;
                        empty_par ^= 1u;
                    }
                }
                g += n_in_seg;
            }
            // every operand of this CTA is on its way: the next kernel in the stream may be scheduled onto SMs as
            // they free up (it waits at its own griddepcontrol.wait until this grid has completed)
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
    } else if (wid <= RB) {
        // ===== MMA issuers: warp 1 + rb feeds the accumulators of row block rb.  The whole warp runs the loop (uniform
        // control flow and descriptors), one elected lane issues.  Measured: the issuing warp, not the tensor pipe, is
        // what can bound this kernel -- an MMA costs ~48 clocks back to back, but a barrier wait ~150-250 and a commit
        // ~200, and with per-lane descriptor arithmetic ~100 per instruction -- hence one issuing warp per row block,
        // so that each spends one accumulator wait, its MMAs and two commits per 256 x 128 unit =====
        // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N, M
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        const bool skip_mma = PROF && (A.dbg & 1) != 0;
        constexpr uint32_t lbo = 128u, sbo = (uint32_t)kseg * NSEG * 16u;  // NSEG*kseg/8 chunks of 16 B per row, 128 B per 8 rows
        constexpr int ksteps = NSEG * kseg / 16;                           // K = 16 halves = two chunks per instruction
        const int rb = wid - 1;
        const uint64_t descA0 = umma_desc(sA, lbo, sbo);
        const uint64_t descB0 = umma_desc(sB, lbo, sbo);
        int s = 0, seg = 0, t = t_first;
        uint32_t full_par = 0, u = 0;
        const bool mprof = PROF && (A.dbg & 8) != 0;
        long long m_t0 = PROF ? clock64() : 0, m_full = 0, m_acc = 0, m_issue = 0, m_x = 0;
#define TC_M0() if (PROF && mprof) m_x = clock64()
#define TC_M1(dst) if (PROF && mprof) dst += clock64() - m_x
        for (long long g = g0; g < g1; ++seg, t = 0) {
            const int abuf = NABUF == 2 ? (seg & 1) : 0;
            const uint32_t a_par = (uint32_t)(NABUF == 2 ? (seg >> 1) & 1 : seg & 1);
            tcb_wait(b_af + 8 * abuf, a_par);
            __syncwarp();
            const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
            const uint64_t descA = descA0 + (uint64_t)(((uint32_t)(abuf * RB + rb) * a_bytes) >> 4);
            for (int k = 0; k < n_in_seg; ++k, ++u) {
                const uint32_t buf = u & 1u, par = (u >> 1) & 1u;
                // the B tile first (it landed long ago: the query's latency is spent while the epilogue still holds the
                // accumulator), then the accumulator: the MMAs go out as soon as the epilogue lets go of it
                TC_M0();
                tcb_wait(b_full + 8 * s, full_par);
                TC_M1(m_full);
                TC_M0();
                tcb_wait(b_acce + 8 * (buf * RB + rb), par ^ 1u);
                __syncwarp();
                TC_M1(m_acc);
                tc_fence_after();
                TC_M0();
                const uint64_t descB = descB0 + (uint64_t)(((uint32_t)s * b_bytes) >> 4);
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcAccCols + rb * NT);
                if (elect_one()) {
                    if (!skip_mma) {
#pragma unroll
                        for (int j = 0; j < ksteps; ++j)
                            umma_f16(d_tmem, descA + (uint64_t)(j * 16), descB + (uint64_t)(j * 16), idesc, j > 0 ? 1u : 0u);
                    }
                    tcb_commit(b_accf + 8 * (buf * RB + rb));
                    tcb_commit(b_empty + 8 * s);                              // the stage is free once both warps' MMAs have read it
                    if (k == n_in_seg - 1) tcb_commit(b_ae + 8 * abuf);       // and so are the A blocks of this segment
                }
                __syncwarp();
                TC_M1(m_issue);
                if (++s == S) {
                    s = 0;
                    full_par ^= 1u;
                }
            }
            g += n_in_seg;
        }
        if (PROF && mprof && lane == 0 && rb == 0) {
            long long *dst = A.prof + ((long long)gridDim.x * 16 + blockIdx.x) * 8;
            dst[0] = clock64() - m_t0, dst[1] = m_full, dst[2] = m_acc, dst[3] = m_issue, dst[4] = (long long)u;
        }
#undef TC_M0
#undef TC_M1
    } else if (wid >= kTcEpiWarp0) {
        // ===== epilogue: thread = one out row x (128 / (EW / 4)) columns of every tile; exp2 + in-thread sum =====
        const int ew = wid - kTcEpiWarp0;
        const int rb = ew / EW, within = ew % EW;
        const int q = within & 3, h = within >> 2;
        const bool skip_exp = PROF && (A.dbg & 2) != 0;
        const bool ld_only = PROF && (A.dbg & 4) != 0;
        const bool prof = PROF && (A.dbg & 8) != 0;
        long long c_acc = 0, c_ld = 0, c_t0 = PROF ? clock64() : 0, c_x = 0, c_pro = 0, c_fin = 0;
#define TC_T0() if (PROF && prof) c_x = clock64()
#define TC_T1(dst) if (PROF && prof) dst += clock64() - c_x
        uint32_t va[32], vb[32];
        const uint32_t n_units = (uint32_t)(g1 - g0);
        // loop-carried addresses: accumulator buffer `cur` of this row block (TMEM columns and the two barriers);
        // the other buffer is reached by XOR with the precomputed differences
        const uint32_t t_buf0 = tc_keep(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rb * NT + h * (NCH * 32)));
        const uint32_t accf0 = tc_keep(b_accf + 8 * rb), acce0 = tc_keep(b_acce + 8 * rb);
        constexpr uint32_t kBufBar = 8 * RB;  // barrier distance between buffer 0 and buffer 1
        if (!skip_exp) {
            tcb_wait(accf0, 0);
            tc_fence_after();
            WOTB_TMEM_LD32(va, t_buf0);
            if (PROF && prof) c_pro = clock64() - c_t0;
        }
        int seg = 0, b = b_first, t = t_first;
        uint32_t u = 0;
        for (long long g = g0; g < g1; ++seg, ++b, t = 0) {
            const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
            double acc = 0.0;
            for (int k0 = 0; k0 < n_in_seg; k0 += 8) {
            // fp32 across 8 tiles (<= 1024 positive terms), then float64: the FP64 pipe stays out of the tile loop
            float facc = 0.f;
            const int k1 = min(k0 + 8, n_in_seg);
            for (int k = k0; k < k1; ++k, ++u) {
                const uint32_t buf = u & 1u;
                const uint32_t taddr = t_buf0 + buf * kTcAccCols;
                if (PROF && skip_exp) {
                    tcb_wait(accf0 + buf * kBufBar, (u >> 1) & 1u);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tcb_arrive(acce0 + buf * kBufBar);
                    continue;
                }
                float tile_sum = 0.f;
                // is the NEXT unit's accumulator complete?  (the MMA warp runs a unit ahead, so normally yes.)  Asked here,
                // answered while this unit's first columns are evaluated
                // (asked unconditionally -- after the last unit the answer is simply not used -- so that the query sits in
                // the same basic block as the arithmetic and can be scheduled in front of it)
                const uint32_t next_ok = tcb_test(accf0 + (buf ^ 1u) * kBufBar, ((u + 1) >> 1) & 1u);
#pragma unroll
                for (int c = 0; c < NCH; c += 2) {
                    TC_T0();
                    WOTB_TMEM_WAIT32(va);  // chunk c (issued one step earlier)
                    TC_T1(c_ld);
                    WOTB_TMEM_LD32(vb, taddr + (c + 1) * 32);
                    tile_sum += (PROF && ld_only) ? tc_plain_sum32(va) : tc_exp2_sum32(va);
                    TC_T0();
                    WOTB_TMEM_WAIT32(vb);
                    TC_T1(c_ld);
                    if (c + 2 < NCH) {
                        WOTB_TMEM_LD32(va, taddr + (c + 2) * 32);
                    } else {
                        // every column of this accumulator is in registers: hand it back to the MMA warp, and
                        // fetch the first columns of the next unit while the last 32 of this one are evaluated
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tcb_arrive(acce0 + buf * kBufBar);
                        if (u + 1 < n_units) {
                            const uint32_t nb = buf ^ 1u;
                            TC_T0();
                            if (!next_ok) tcb_wait(accf0 + nb * kBufBar, ((u + 1) >> 1) & 1u);
                            TC_T1(c_acc);
                            tc_fence_after();
                            WOTB_TMEM_LD32(va, t_buf0 + nb * kTcAccCols);
                        }
                    }
                    tile_sum += (PROF && ld_only) ? tc_plain_sum32(vb) : tc_exp2_sum32(vb);
                }
                facc += tile_sum;
            }
            acc += (double)facc;
            }
            g += n_in_seg;
            // ---- this CTA's share of out block b is complete: publish it, and finish the rows if it was the last ----
            TC_T0();
            tc_publish<COLPASS, EW>(A, V, ctrl, mode, rowsum_out, b, nt, T, G, rb, q, h, lane, acc, false);
            TC_T1(c_fin);
        }
        if (PROF && prof && lane == 0) {
            long long *dst = A.prof + ((long long)blockIdx.x * (RB * EW) + ew) * 8;
            dst[0] = clock64() - c_t0, dst[1] = c_acc, dst[2] = c_ld, dst[3] = (long long)n_units;
            dst[4] = c_pro, dst[5] = c_fin;
        }
#undef TC_T0
#undef TC_T1
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcTmemCols)
                     : "memory");
    }
}

// =====================================================================================================================
// A whole batch of Sinkhorn iterations in ONE persistent cooperative launch: `n_iters` x (row half-step, column
// half-step), the CTAs meeting at a grid barrier between half-steps.  A pass launched on its own pays ~10 us that are
// not tile work (measured: 45.2 us at 32.4 units per CTA, 100.4 us at 83.8 units -> 1.07 us per unit + 10.4 us):
// kernel drain and launch, barrier initialisation, TMEM allocation, pipeline fill.  Here barriers and TMEM are set up
// once per batch, the ring / accumulator / A-buffer pipelines run through all passes with their phases carried along,
// and between two passes only the finish chain (last partial -> ticket -> float64 update of the block), the grid
// barrier and the first B tile's trip from L2 remain.  Everything else -- stream-K dealing, operand layout, epilogue
// arithmetic, slot-order reductions, the device state machine -- is that of k_online_tc (mode 0).
// =====================================================================================================================
__device__ __forceinline__ void tc_grid_barrier(unsigned int *count, unsigned int target) {
    // called by ONE thread after a CTA-wide barrier; bounded so that a protocol bug traps instead of hanging the GPU
    __threadfence();
    atomicAdd(count, 1u);
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int *>(count) < target) {
        if (++spins > (1u << 28)) __trap();
    }
    __threadfence();
}

template <int KSEG, int EW, int NSEG>
__global__ void __launch_bounds__(128 + kTcRowBlocks * EW * 32, 1)
    k_online_batch(TcArgs Arow, TcArgs Acol, SolveVecs V, SolveCtrl *ctrl, int n_iters) {
    constexpr int RB = kTcRowBlocks, NT = kTcN;
    constexpr int NCH = 4 / (EW / 4);
    constexpr int kseg = KSEG;
    constexpr int NABUF = NSEG == 3 ? 2 : 1;
    constexpr uint32_t row_bytes = (uint32_t)kseg * NSEG * 2u;
    constexpr uint32_t a_bytes = kTcM * row_bytes, b_bytes = NT * row_bytes;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int S = Arow.n_stages;
    const uint32_t smem0 = smem_u32(tc_smem);
    const uint32_t bars = smem0;
    const uint32_t sA = smem0 + kTcTail, sB = sA + NABUF * RB * a_bytes;
    const uint32_t b_full = bars, b_empty = bars + 8 * kTcMaxStages;
    const uint32_t b_accf = b_empty + 8 * kTcMaxStages;
    const uint32_t b_acce = b_accf + 8 * 2 * RB;
    const uint32_t b_af = b_acce + 8 * 2 * RB, b_ae = b_af + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tc_smem + 8 * (2 * kTcMaxStages + 4 * RB + 4));
    const int G = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            tcb_init(b_full + 8 * s, 1);
            tcb_init(b_empty + 8 * s, RB);
        }
        for (int b = 0; b < 2 * RB; ++b) {
            tcb_init(b_accf + 8 * b, 1);
            tcb_init(b_acce + 8 * b, EW);
        }
        for (int b = 0; b < 2; ++b) {
            tcb_init(b_af + 8 * b, 1);
            tcb_init(b_ae + 8 * b, RB);
        }
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

    // pipeline positions, carried through all passes of the launch (every role keeps only the ones it uses)
    int ring = 0;              // B stage the role touches next
    uint32_t ring_par = 0;     // producer: parity of `empty` it waits for (starts at 1); MMA: parity of `full`
    uint32_t a_seg = 0;        // out-block segments started so far (A buffer and its parity)
    uint32_t u = 0;            // work units so far (accumulator buffer and its parity)
    if (wid == 0) ring_par = 1;

    for (int pass = 0; pass < 2 * n_iters; ++pass) {
        // every CTA sees the same control state here: it changes only when a column pass closes an iteration, before
        // the grid barrier.  A finished batch, a tau stop or max_iter end the launch for all of them together.
        if (!iteration_active(ctrl)) break;
        const bool col = (pass & 1) != 0;
        const TcArgs &A = col ? Acol : Arow;
        const int nt = A.in_ntiles;
        const long long T = (long long)A.n_blocks * nt;
        const long long g0 = (long long)blockIdx.x * T / G, g1 = (long long)(blockIdx.x + 1) * T / G;
        const int b_first = (int)(g0 / nt), t_first = (int)(g0 - (long long)b_first * nt);

        if (wid == 0) {
            if (lane == 0) {
                // ===== TMA producer =====  (the offset slots of the B rows were written by other CTAs' generic stores in
                // the previous pass: ordered by their fence.proxy.async + gpu-scope fence, the grid barrier and this
                // thread's own fence inside it)
                const unsigned char *srcA = reinterpret_cast<const unsigned char *>(A.opA);
                const unsigned char *srcB = reinterpret_cast<const unsigned char *>(A.opB);
                int b = b_first, t = t_first;
                for (long long g = g0; g < g1; ++a_seg, ++b, t = 0) {
                    const int abuf = NABUF == 2 ? (int)(a_seg & 1u) : 0;
                    const uint32_t a_par = NABUF == 2 ? (a_seg >> 1) & 1u : a_seg & 1u;
                    tcb_wait(b_ae + 8 * abuf, a_par ^ 1u);
                    tcb_expect_tx(b_af + 8 * abuf, RB * a_bytes);
                    const unsigned char *blockA = srcA + (size_t)(A.out_blk0 + b) * (RB * a_bytes);
                    for (int rb = 0; rb < RB; ++rb)
                        tcb_bulk_g2s(sA + (abuf * RB + rb) * a_bytes, blockA + (size_t)rb * a_bytes, a_bytes, b_af + 8 * abuf);
                    const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
                    for (int k = 0; k < n_in_seg; ++k) {
                        tcb_wait(b_empty + 8 * ring, ring_par);
                        tcb_expect_tx(b_full + 8 * ring, b_bytes);
                        tcb_bulk_g2s(sB + (uint32_t)ring * b_bytes, srcB + (size_t)(A.in_tile0 + t + k) * b_bytes, b_bytes,
                                     b_full + 8 * ring);
                        if (++ring == S) {
                            ring = 0;
                            ring_par ^= 1u;
                        }
                    }
                    g += n_in_seg;
                }
            }
        } else if (wid <= RB) {
            // ===== MMA issuers, one warp per row block =====
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
            constexpr uint32_t lbo = 128u, sbo = (uint32_t)kseg * NSEG * 16u;
            constexpr int ksteps = NSEG * kseg / 16;
            const int rb = wid - 1;
            const uint64_t descA0 = umma_desc(sA, lbo, sbo);
            const uint64_t descB0 = umma_desc(sB, lbo, sbo);
            int t = t_first;
            for (long long g = g0; g < g1; ++a_seg, t = 0) {
                const int abuf = NABUF == 2 ? (int)(a_seg & 1u) : 0;
                const uint32_t a_par = NABUF == 2 ? (a_seg >> 1) & 1u : a_seg & 1u;
                tcb_wait(b_af + 8 * abuf, a_par);
                __syncwarp();
                const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
                const uint64_t descA = descA0 + (uint64_t)(((uint32_t)(abuf * RB + rb) * a_bytes) >> 4);
                for (int k = 0; k < n_in_seg; ++k, ++u) {
                    const uint32_t buf = u & 1u, par = (u >> 1) & 1u;
                    tcb_wait(b_acce + 8 * (buf * RB + rb), par ^ 1u);
                    __syncwarp();
                    tcb_wait(b_full + 8 * ring, ring_par);
                    __syncwarp();
                    tc_fence_after();
                    const uint64_t descB = descB0 + (uint64_t)(((uint32_t)ring * b_bytes) >> 4);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcAccCols + rb * NT);
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < ksteps; ++j)
                            umma_f16(d_tmem, descA + (uint64_t)(j * 16), descB + (uint64_t)(j * 16), idesc, j > 0 ? 1u : 0u);
                        tcb_commit(b_accf + 8 * (buf * RB + rb));
                        tcb_commit(b_empty + 8 * ring);
                        if (k == n_in_seg - 1) tcb_commit(b_ae + 8 * abuf);
                    }
                    __syncwarp();
                    if (++ring == S) {
                        ring = 0;
                        ring_par ^= 1u;
                    }
                }
                g += n_in_seg;
            }
        } else if (wid >= kTcEpiWarp0) {
            // ===== epilogue =====
            const int ew = wid - kTcEpiWarp0;
            const int rb = ew / EW, within = ew % EW;
            const int q = within & 3, h = within >> 2;
            uint32_t va[32], vb[32];
            const uint32_t u_end = u + (uint32_t)(g1 - g0);
            const uint32_t t_buf0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rb * NT + h * (NCH * 32));
            const uint32_t accf0 = b_accf + 8 * rb, acce0 = b_acce + 8 * rb;
            constexpr uint32_t kBufBar = 8 * RB;
            if (u < u_end) {
                tcb_wait(accf0 + (u & 1u) * kBufBar, (u >> 1) & 1u);
                tc_fence_after();
                WOTB_TMEM_LD32(va, t_buf0 + (u & 1u) * kTcAccCols);
            }
            int b = b_first, t = t_first;
            for (long long g = g0; g < g1; ++b, t = 0) {
                const int n_in_seg = (int)min((long long)(nt - t), g1 - g);
                double acc = 0.0;
                for (int k0 = 0; k0 < n_in_seg; k0 += 8) {
                    float facc = 0.f;
                    const int k1 = min(k0 + 8, n_in_seg);
                    for (int k = k0; k < k1; ++k, ++u) {
                        const uint32_t buf = u & 1u;
                        const uint32_t taddr = t_buf0 + buf * kTcAccCols;
                        float tile_sum = 0.f;
#pragma unroll
                        for (int c = 0; c < NCH; c += 2) {
                            WOTB_TMEM_WAIT32(va);
                            WOTB_TMEM_LD32(vb, taddr + (c + 1) * 32);
                            tile_sum += tc_exp2_sum32(va);
                            WOTB_TMEM_WAIT32(vb);
                            if (c + 2 < NCH) {
                                WOTB_TMEM_LD32(va, taddr + (c + 2) * 32);
                            } else {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) tcb_arrive(acce0 + buf * kBufBar);
                                if (u + 1 < u_end) {
                                    const uint32_t nb = buf ^ 1u;
                                    tcb_wait(accf0 + nb * kBufBar, ((u + 1) >> 1) & 1u);
                                    tc_fence_after();
                                    WOTB_TMEM_LD32(va, t_buf0 + nb * kTcAccCols);
                                }
                            }
                            tile_sum += tc_exp2_sum32(vb);
                        }
                        facc += tile_sum;
                    }
                    acc += (double)facc;
                }
                g += n_in_seg;
                // ---- publish this CTA's share of out block b; the last warp to arrive at its rows finishes them ----
                if (col)
                    tc_publish<true, EW>(A, V, ctrl, 0, nullptr, b, nt, T, G, rb, q, h, lane, acc, true);
                else
                    tc_publish<false, EW>(A, V, ctrl, 0, nullptr, b, nt, T, G, rb, q, h, lane, acc, true);
            }
        }
        // ---- end of the half-step: every partial is published, every block finished, the iteration possibly closed ----
        __syncthreads();
        // one thread arrives, spins and fences (the gpu-scope fence drops this SM's L1 lines, so the state written by
        // other CTAs is re-read from L2 by every thread behind the CTA barrier)
        if (tid == 0) tc_grid_barrier(&ctrl->grid_bar, (unsigned int)(pass + 1) * (unsigned int)G);
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcTmemCols)
                     : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------
inline int tc_kseg(int d) { return (int)round_up(d + 2, 16); }
inline bool tc_supported(int d) { return tc_kseg(d) <= kTcMaxKseg; }

struct TcPlan {
    int kseg = 0, n_stages = 0, ew = 4, nseg = 3;
    bool prof = false;  // measurement build (cycle counters, dbg switches); kseg 32 only
    size_t smem = 0;
    int threads() const { return 128 + kTcRowBlocks * ew * 32; }
    size_t row_bytes() const { return (size_t)kseg * nseg * 2; }
};

// nseg 3: the default fp16 hi/lo operands; nseg 6: precise mode (always 16 epilogue warps)
inline TcPlan tc_plan(int d, int ew = 4, int nseg = 3, bool prof = false) {
    TcPlan p;
    p.kseg = tc_kseg(d);
    p.nseg = nseg;
    p.ew = nseg == 6 ? 8 : ew;
    p.prof = prof && p.kseg == 32;
    const size_t row_bytes = p.row_bytes();
    const size_t a = (nseg == 3 ? 2 : 1) * (size_t)kTcOut * row_bytes, b = (size_t)kTcN * row_bytes;  // A double buffered (3 segments)
    int s = (int)((kTcSmemLimit - a - kTcTail) / b);
    if (s > kTcMaxStages) s = kTcMaxStages;
    p.n_stages = s;
    p.smem = a + (size_t)s * b + kTcTail;
    return p;
}

// One-wave grid and the number of partial-sum slots an out block can receive.
inline int tc_grid(int sm_count, int n_blocks, int in_tiles, int *max_slots) {
    const long long T = (long long)n_blocks * in_tiles;
    const int G = (int)(T < sm_count ? T : sm_count);
    *max_slots = 2 * ((int)cdiv(G, n_blocks) + 1);  // x 2 column halves (16 epilogue warps)
    return G;
}

// the instantiations that exist: (kseg, epilogue warps per row block, segments, profiling build)
#define WOTB_TC_VARIANTS(X) \
    X(16, 4, 3, false) X(32, 4, 3, false) X(48, 4, 3, false) X(16, 8, 3, false) X(32, 8, 3, false) X(48, 8, 3, false) \
    X(16, 8, 6, false) X(32, 8, 6, false) X(48, 8, 6, false) X(32, 4, 3, true) X(32, 8, 3, true) X(32, 8, 6, true)

inline int tc_configure(const TcPlan &plan) {
    cudaError_t e = cudaErrorInvalidValue;
    const int bytes = (int)plan.smem;
#define WOTB_TC_CFG(K, E, N, P)                                                                                          \
    if (plan.kseg == K && plan.ew == E && plan.nseg == N && plan.prof == P) {                                            \
        e = cudaFuncSetAttribute(k_online_tc<false, K, E, N, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);    \
        if (e == cudaSuccess)                                                                                            \
            e = cudaFuncSetAttribute(k_online_tc<true, K, E, N, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    }
    WOTB_TC_VARIANTS(WOTB_TC_CFG)
#undef WOTB_TC_CFG
    WOTB_CUDA(e);
    return WOTB_OK;
}

// Programmatic dependent launch between consecutive passes (and whatever precedes them in the stream): the next
// kernel's CTAs are scheduled as SMs free up and set themselves up while the current one drains.  Measured on B200
// (profiles/r1m): one solve at a time 48.0 -> 45.6 us per 12.5k x 12.4k pass, 6.41 -> 6.62 tmaps/s; with two solves
// in flight on separate streams it LOSES (7.31 -> 7.14): SMs freed by a draining kernel are better used by the
// other stream's CTAs than by waiting ones, so wot_b200.pipeline turns it off (wotb_set_pdl).  WOTB_NO_PDL=1 in
// the environment sets the initial state to off.
inline int &tc_pdl_flag() {
    static int use = -1;
    return use;
}
inline bool tc_use_pdl() {
    int &use = tc_pdl_flag();
    if (use < 0) {
        const char *e = getenv("WOTB_NO_PDL");
        use = (e && e[0] == '1') ? 0 : 1;
    }
    return use == 1;
}

template <bool COLPASS, int K, int E, int N, bool P>
inline void tc_launch_one(const TcPlan &plan, int grid, cudaStream_t st, const TcArgs &A, const SolveVecs &V, SolveCtrl *ctrl,
                          int mode, double *rowsum_out) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(128 + kTcRowBlocks * E * 32), cfg.dynamicSmemBytes = plan.smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr, cfg.numAttrs = tc_use_pdl() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_online_tc<COLPASS, K, E, N, P>, A, V, ctrl, mode, rowsum_out);
}

// ---- the persistent batch kernel: 16 epilogue warps, either operand form ----
#define WOTB_TC_BATCH_VARIANTS(X) X(16, 3) X(32, 3) X(48, 3) X(16, 6) X(32, 6) X(48, 6)

inline int tc_batch_configure(const TcPlan &plan) {
    if (plan.ew != 8 || plan.prof) return WOTB_ERR_INVALID;
    cudaError_t e = cudaErrorInvalidValue;
#define WOTB_TC_BCFG(K, N)                                                                                                   \
    if (plan.kseg == K && plan.nseg == N)                                                                                    \
        e = cudaFuncSetAttribute(k_online_batch<K, 8, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    WOTB_TC_BATCH_VARIANTS(WOTB_TC_BCFG)
#undef WOTB_TC_BCFG
    WOTB_CUDA(e);
    return WOTB_OK;
}

// can `grid` CTAs of the batch kernel be co-resident (cooperative launch)?
inline bool tc_batch_fits(const TcPlan &plan, int grid, int sm_count) {
    int per_sm = 0;
    cudaError_t e = cudaErrorInvalidValue;
#define WOTB_TC_BOCC(K, N)                                                                                                  \
    if (plan.kseg == K && plan.nseg == N)                                                                                   \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_online_batch<K, 8, N>, plan.threads(), plan.smem);
    WOTB_TC_BATCH_VARIANTS(WOTB_TC_BOCC)
#undef WOTB_TC_BOCC
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return per_sm * sm_count >= grid;
}

inline void tc_batch_launch(const TcPlan &plan, int grid, cudaStream_t st, const TcArgs &Arow, const TcArgs &Acol,
                            const SolveVecs &V, SolveCtrl *ctrl, int n_iters) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(plan.threads()), cfg.dynamicSmemBytes = plan.smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
#define WOTB_TC_BRUN(K, N)                                                                      \
    if (plan.kseg == K && plan.nseg == N) {                                                     \
        cudaLaunchKernelEx(&cfg, k_online_batch<K, 8, N>, Arow, Acol, V, ctrl, n_iters);        \
        return;                                                                                 \
    }
    WOTB_TC_BATCH_VARIANTS(WOTB_TC_BRUN)
#undef WOTB_TC_BRUN
}

template <bool COLPASS>
inline void tc_launch(const TcPlan &plan, int grid, cudaStream_t st, const TcArgs &A, const SolveVecs &V, SolveCtrl *ctrl,
                      int mode, double *rowsum_out) {
#define WOTB_TC_CASE(K, E, N, P)                                                             \
    if (plan.kseg == K && plan.ew == E && plan.nseg == N && plan.prof == P) {                \
        tc_launch_one<COLPASS, K, E, N, P>(plan, grid, st, A, V, ctrl, mode, rowsum_out);    \
        return;                                                                              \
    }
    WOTB_TC_VARIANTS(WOTB_TC_CASE)
#undef WOTB_TC_CASE
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
    WOTB_REQUIRE(n_out >= 1 && n_in >= 1 && d >= 1 && reps >= 0, "bad sizes");
    const int dbg = impl >> 4;
    impl &= 15;
    WOTB_REQUIRE(impl == 0 || ((impl >= 1 && impl <= 3) && tc_supported(d)),
                 "impl: 0 SIMT FP32, 1 tcgen05 with 8 epilogue warps, 2 tcgen05 with 16, 3 tcgen05 precise mode (d <= 46)");
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
    int grid_used = 0;
    if (impl >= 1) {
        const TcPlan plan = tc_plan(d, impl == 1 ? 4 : 8, impl == 3 ? 6 : 3, dbg != 0);
        WOTB_REQUIRE(dbg == 0 || plan.prof, "the measurement switches exist for kseg = 32 (15 <= d <= 30) only");
        WOTB_TRY(tc_configure(plan));
        const int64_t po = round_up(n_out, kTcOut), pi = round_up(n_in, kTcOut);
        const size_t row_bytes = plan.row_bytes();
        const int out_blocks = (int)cdiv(n_out, kTcOut), in_tiles = (int)cdiv(n_in, kTcN);
        int max_slots = 0;
        const int grid = tc_grid(ctx->sm_count, out_blocks, in_tiles, &max_slots);
        grid_used = grid;
        const size_t o_ao = take(po * row_bytes), o_bo = take(po * row_bytes), o_ai = take(pi * row_bytes),
                     o_bi = take(pi * row_bytes), o_res = take(po * 8), o_part = take((size_t)max_slots * po * 8),
                     o_cnt = take((size_t)out_blocks * 32 + 64), o_geo = take(sizeof(TcGeo));
        WOTB_TRY(ctx->onl.reserve(off));
        char *ob = ctx->onl.as<char>();
        __half *Ao = (__half *)(ob + o_ao), *Bo = (__half *)(ob + o_bo), *Ai = (__half *)(ob + o_ai), *Bi = (__half *)(ob + o_bi);
        double *resid = (double *)(ob + o_res), *part = (double *)(ob + o_part);
        unsigned int *cnt = (unsigned int *)(ob + o_cnt);
        TcGeo *geo = (TcGeo *)(ob + o_geo);
        WOTB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)out_blocks * 32, st));
        WOTB_CUDA(cudaMemsetAsync(geo, 0, sizeof(TcGeo), st));
        k_tc_geo<<<(unsigned)cdiv(n_out, 256), 256, 0, st>>>(x_out, (int)n_out, d, geo, 0);
        k_tc_geo<<<(unsigned)cdiv(n_in, 256), 256, 0, st>>>(x_in, (int)n_in, d, geo, 1);
        const int cpr = plan.kseg / 8;
        k_tc_pack<<<(unsigned)cdiv(po * cpr, 256), 256, 0, st>>>(x_out, (int)n_out, d, po, plan.kseg, plan.nseg, Ao, Bo, nullptr,
                                                                 scale, geo);
        k_tc_pack<<<(unsigned)cdiv(pi * cpr, 256), 256, 0, st>>>(x_in, (int)n_in, d, pi, plan.kseg, plan.nseg, Ai, Bi, nullptr,
                                                                 scale, geo);
        k_tc_slots<<<(unsigned)cdiv(n_out > n_in ? n_out : n_in, 256), 256, 0, st>>>(off_out, (int)n_out, Ao, resid, off_in,
                                                                                     (int)n_in, Bi, plan.kseg, plan.nseg, nullptr, 0);
        TcArgs A;
        A.opA = Ao, A.opB = Bi, A.resid = resid, A.out_n = (int)n_out, A.out_ld = po;
        A.n_blocks = out_blocks, A.out_blk0 = 0, A.in_tile0 = 0, A.in_ntiles = in_tiles;
        A.n_stages = plan.n_stages, A.part = part, A.counters = cnt, A.dbg = dbg;
        A.prof = nullptr;
        if (dbg & 8) {
            WOTB_TRY(ctx->hTmp.reserve((size_t)ctx->sm_count * 17 * 8 * 8));
            WOTB_CUDA(cudaMemsetAsync(ctx->hTmp.ptr, 0, (size_t)ctx->sm_count * 17 * 8 * 8, st));
            A.prof = ctx->hTmp.as<long long>();
        }
        tc_launch<false>(plan, grid, st, A, V, d_ctrl, 4, sums);
        WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
        for (int r = 0; r < reps; ++r) tc_launch<false>(plan, grid, st, A, V, d_ctrl, 4, sums);
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
    if (impl >= 1 && (dbg & 8)) {  // cycle counters of the last launch: mean over the epilogue warps
        const size_t n = (size_t)grid_used * 16 * 8, nm = (size_t)grid_used * 8;
        std::vector<long long> hp(n + nm);
        WOTB_CUDA(cudaMemcpy(hp.data(), ctx->hTmp.ptr, (n + nm) * 8, cudaMemcpyDeviceToHost));
        double tot = 0, wa = 0, wl = 0, tiles = 0, pro = 0, fin = 0, tmax = 0;
        size_t warps = 0;
        for (size_t w = 0; w + 7 < n; w += 8) {
            if (hp[w + 3] <= 0) continue;
            tot += hp[w], wa += hp[w + 1], wl += hp[w + 2], tiles += hp[w + 3], pro += hp[w + 4], fin += hp[w + 5];
            if ((double)hp[w] > tmax) tmax = (double)hp[w];
            ++warps;
        }
        if (getenv("WOTB_TC_PROF_CTAS")) {  // per-CTA view: mean total / publish+finish cycles of its epilogue warps, units
            fprintf(stderr, "[tc prof] per CTA: total fin units\n");
            for (int c = 0; c < grid_used; ++c) {
                double t = 0, f = 0, u = 0;
                int nw = 0;
                for (int w = 0; w < 16; ++w) {
                    const size_t at = ((size_t)c * 16 + w) * 8;
                    if (hp[at + 3] <= 0) continue;
                    t += hp[at], f += hp[at + 5], u += hp[at + 3], ++nw;
                }
                if (nw) fprintf(stderr, "%d %.0f %.0f %.1f\n", c, t / nw, f / nw, u / nw);
            }
        }
        if (warps)
            fprintf(stderr,
                    "[tc prof] warps %zu units/warp %.1f | cycles per warp: total %.0f (max %.0f) = prologue %.0f + publish/finish "
                    "%.0f + loop; per unit in loop %.0f, of which waiting for accumulators %.0f, for tcgen05.ld %.0f\n",
                    warps, tiles / warps, tot / warps, tmax, pro / warps, fin / warps, (tot - pro - fin) / tiles, wa / tiles,
                    wl / tiles);
        double mt = 0, mf = 0, ma = 0, mi = 0, mu = 0;
        for (size_t c = 0; c < nm; c += 8) mt += hp[n + c], mf += hp[n + c + 1], ma += hp[n + c + 2], mi += hp[n + c + 3], mu += hp[n + c + 4];
        if (mu > 0)
            fprintf(stderr, "[tc prof] MMA warp, cycles per unit: total %.0f = waiting for B tiles %.0f + waiting for free accumulators %.0f + "
                            "issue and commit %.0f + rest\n", mt / mu, mf / mu, ma / mu, mi / mu);
    }
    return WOTB_OK;
}

// MUFU.EX2 peak of this GPU, measured: the roofline denominator of the online kernels (MEASURED_PEAKS.json has
// HBM and tensor figures only).  16 independent ex2 chains per thread, 32 warps per SM, nothing else in the loop.
__global__ void __launch_bounds__(1024) k_mufu_peak(float *out, int iters) {
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = -1.f - 0.001f * (float)(threadIdx.x + e);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = ex2_approx(v[e]);
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) s += v[e];
    if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}

void set_pdl(bool on) {
    static const bool always = []() {
        const char *e = getenv("WOTB_PDL_ALWAYS");  // measurement knob: ignore the pipeline's request to turn PDL off
        return e && e[0] == '1';
    }();
    tc_pdl_flag() = (on || always) ? 1 : 0;
}

int bench_mufu(wotb_ctx *ctx, double *ex2_per_s) {
    WOTB_REQUIRE(ctx && ex2_per_s, "NULL argument");
    WOTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WOTB_TRY(ctx->hTmp.reserve(256));
    const int iters = 4096, blocks = ctx->sm_count * 2;
    k_mufu_peak<<<blocks, 1024, 0, st>>>(ctx->hTmp.as<float>(), 64);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        WOTB_CUDA(cudaEventRecord(ctx->ev0, st));
        k_mufu_peak<<<blocks, 1024, 0, st>>>(ctx->hTmp.as<float>(), iters);
        WOTB_CUDA(cudaEventRecord(ctx->ev1, st));
        WOTB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        WOTB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double rate = (double)blocks * 1024.0 * 16.0 * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    WOTB_CUDA(cudaGetLastError());
    *ex2_per_s = best;
    return WOTB_OK;
}

}  // namespace wotb
