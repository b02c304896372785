// k1_reduce.cuh - K1: fused point-to-plane residual / weight / 6-DoF Jacobian / normal-equation
// reduction for sm_100a.
//
// Replaces, per source slot and per ICP iteration (reference file:line):
//   pointBodyToGlobal (FP64 math, float32 store)         DCReg/include/utils.hpp:630-636
//   residual + LOAM weight + gate                         DCReg/src/icp_test_runner.cpp:1774-1803
//   stream compaction of flagged slots                    icp_test_runner.cpp:1816-1840 (skipped: gated in place)
//   computePointToPlaneJacobian + row fill                DCReg/include/math_utils.hpp:102-121, icp_test_runner.cpp:1863-1907
//   H = A^T A, g = A^T b                                  icp_test_runner.cpp:1910-1919
//   SymmetricHessianComputer (21 + 6 accumulators)        DCReg/include/hessian_computer.h:62-123
//
// Roofline: HBM.  Algorithmic bytes per slot = 32 (float4 point + float4 plane); 48 with the
// FP64 plane variant.  All products and sums are FP64 (precision contract: pose 1e-6, Schur
// eigenvalues 1e-8 relative), so the FP64 pipe (64 FMA/clk/SM) is the second bound: the per-slot
// FP64 work is cut by accumulating the outer products in the WORLD frame,
//     J_r = [ (p x R^T n)^T , (R^T n)^T ] = [ (Rp x n)^T , n^T ] * blkdiag(R, R),
// so the 27 sums are taken over u = [Rp x n ; n] (Rp is a by-product of the point transform) and
// the constant 6x6 congruence with blkdiag(R,R) is applied once, in the final reduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k2_solve.cuh"

namespace k1 {

using k2::kAcc;

struct Pose {            // R row-major, t
    double R[9];
    double t[3];
};

struct Acc {
    double h[21];        // upper triangle of sum a a^T, a = w * u (world frame)
    double g[6];         // sum a * b
    double sr2;          // sum r^2 over effective slots          (rmse, icp_test_runner.cpp:1803,1858)
    double sb2;          // sum b^2                                (objective, icp_test_runner.cpp:1919)
    int neff;            // effective correspondences
    int npt;             // correspondence_pt_count (set by the caller of accumulate_slot)
};

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
    for (int i = 0; i < 21; ++i) a.h[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) a.g[i] = 0.0;
    a.sr2 = 0.0; a.sb2 = 0.0; a.neff = 0; a.npt = 0;
}

// fast full-precision reciprocal for s in (0.1, 1]: MUFU.RCP64H seed + 2 Newton steps (error < 1 ulp)
__device__ __forceinline__ double rcp_newton(double s) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double e = fma(-s, y, 1.0);
    y = fma(y, e, y);
    e = fma(-s, y, 1.0);
    return fma(y, e, y);
}

// One slot.  (px,py,pz) body-frame point, (nx,ny,nz,d) plane in the world frame (FP64 values;
// the float4 variants convert before the call).  Weight rule: icp_test_runner.cpp:1776-1785.
//
// The reference stores coeff = (s n, s r) as float32 and rebuilds the normal as coeff/s
// (:1786-1790, 1889, 1906); the Jacobian row is (s + r ds_dr) [ -n^T R [p]x , n^T R ] (:1898).
// With u' = fl32(s n) this is (w/s) [ (Rp x u')^T , u'^T ] blkdiag(R,R), w = s + r ds_dr, so the 1/s
// never has to be applied to the three components: w/s = 1 without the weight derivative and
// 2 - 1/s with it (r ds_dr = -0.9|r| = s - 1 on 0 < s < 1).
__device__ __forceinline__ void accumulate_slot(Acc& a, const Pose& P, double px, double py, double pz,
                                                double nx, double ny, double nz, double d, bool use_wd,
                                                bool has_plane) {
    // q = fl32(R p + t)  (utils.hpp:630-636: FP64 math, float32 store)
    const double wx = P.R[0] * px + P.R[1] * py + P.R[2] * pz;   // Rp (world-rotated, no translation)
    const double wy = P.R[3] * px + P.R[4] * py + P.R[5] * pz;
    const double wz = P.R[6] * px + P.R[7] * py + P.R[8] * pz;
    const double qx = (double)(float)(wx + P.t[0]);
    const double qy = (double)(float)(wy + P.t[1]);
    const double qz = (double)(float)(wz + P.t[2]);
    const double r = nx * qx + ny * qy + nz * qz + d;            // icp_test_runner.cpp:1774
    const double s = 1.0 - 0.9 * fabs(r);                        // :1776 (the max(0, .) is implied by the gate)
    if (!(has_plane && s > 0.1)) return;                         // :1785
    const double ux = (double)(float)(s * nx);
    const double uy = (double)(float)(s * ny);
    const double uz = (double)(float)(s * nz);
    const double b = -(double)(float)(s * r);
    double v[6];
    v[0] = wy * uz - wz * uy;                                    // Rp x u'
    v[1] = wz * ux - wx * uz;
    v[2] = wx * uy - wy * ux;
    v[3] = ux; v[4] = uy; v[5] = uz;
    if (use_wd && s < 1.0) {                                     // :1780-1783, 1898
        const double k = 2.0 - rcp_newton(s);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] *= k;
    }
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) { a.h[k] = fma(v[i], v[j], a.h[k]); ++k; }
        a.g[i] = fma(v[i], b, a.g[i]);
    }
    a.sr2 = fma(r, r, a.sr2);
    a.sb2 = fma(b, b, a.sb2);
    a.neff += 1;
}

// (double)(float)x without the two F2F conversions: round-to-nearest-even to 24 significant bits directly on
// the FP64 bit pattern (5 integer instructions on the ALU pipe instead of 16 cycles of the 16-lane XU pipe,
// which profiles showed to be the busiest unit of this kernel).  Identical to the float round trip for
// x = 0 and for 2^-126 <= |x| < 2^128, i.e. whenever the float32 result is a normal number or zero.
__device__ __forceinline__ double round_f32(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    const unsigned lsb = ((unsigned)b >> 29) & 1u;
    b += 0x0FFFFFFFull + lsb;                      // 64-bit add: the carry runs into the exponent when needed
    b &= 0xFFFFFFFFE0000000ull;
    return __longlong_as_double((long long)b);
}

// float -> double on the integer ALU, 5 instructions.  Exact for every normal float; +-0 and float denormals
// (|x| < 1.2e-38) come out as +-2^-126-sized values instead of 0 (an absolute perturbation of 1e-38 m on a
// coordinate - far below one FP64 ulp of any coordinate that is not itself ~1e-22 m), and Inf/NaN map to ~1e38-
// sized finite values which - like NaN in the reference - fail the weight gate and drop the slot.
// Why not F2F: on B200 the 64-bit conversions, DFMA/DMUL/DADD and DMMA all issue through one shared pipe whose
// issue cost is additive (measured per warp instruction and sub-partition: F2F.F64.F32 6, F2F.F32.F64 9, DFMA 2,
// DMMA 16.4 cycles; tools/microbench*.cu), and that pipe is the binding resource of K1.
__device__ __forceinline__ double f32_to_f64(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned hi = (((u >> 3) & 0x0FFFFFFFu) + 0x38000000u) | (u & 0x80000000u);   // re-bias exponent by +896
    return __hiloint2double((int)hi, (int)(u << 29));
}

// Branch-free variant for U independent slots handled by one thread (straight-line code, so ptxas
// interleaves the U dependency chains and hides the XU / FP64 latencies).  Invalid slots (no plane, or
// weight gate failed) contribute exact zeros: their weight s is forced to 0, which zeroes u', b and the row.
template <int U, bool kUseWd>
__device__ __forceinline__ void accumulate_slots(Acc& a, const Pose& P, const double (&px)[U], const double (&py)[U],
                                                 const double (&pz)[U], const double (&nx)[U], const double (&ny)[U],
                                                 const double (&nz)[U], const double (&d)[U], const bool (&has)[U]) {
    double wx[U], wy[U], wz[U], r[U], s[U], b[U], v[U][6];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        wx[u] = P.R[0] * px[u] + P.R[1] * py[u] + P.R[2] * pz[u];
        wy[u] = P.R[3] * px[u] + P.R[4] * py[u] + P.R[5] * pz[u];
        wz[u] = P.R[6] * px[u] + P.R[7] * py[u] + P.R[8] * pz[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double qx = round_f32(wx[u] + P.t[0]);             // utils.hpp:630-636 (float32 store)
        const double qy = round_f32(wy[u] + P.t[1]);
        const double qz = round_f32(wz[u] + P.t[2]);
        const double rr = nx[u] * qx + ny[u] * qy + nz[u] * qz + d[u];   // icp_test_runner.cpp:1774
        const double ss = 1.0 - 0.9 * fabs(rr);                  // :1776
        valid[u] = has[u] && (ss > 0.1);                         // :1785
        s[u] = valid[u] ? ss : 0.0;
        r[u] = valid[u] ? rr : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double ux = round_f32(s[u] * nx[u]);               // coeff.x/y/z  (:1787-1789)
        const double uy = round_f32(s[u] * ny[u]);
        const double uz = round_f32(s[u] * nz[u]);
        b[u] = -round_f32(s[u] * r[u]);                          // -coeff.intensity (:1790, 1906)
        v[u][0] = wy[u] * uz - wz[u] * uy;                       // Rp x u'
        v[u][1] = wz[u] * ux - wx[u] * uz;
        v[u][2] = wx[u] * uy - wy[u] * ux;
        v[u][3] = ux; v[u][4] = uy; v[u][5] = uz;
        if (kUseWd) {                                            // :1780-1783, 1898: row scale w/s = 2 - 1/s on 0 < s < 1
            const double sw = (valid[u] && s[u] < 1.0) ? s[u] : 1.0;
            const double k = 2.0 - rcp_newton(sw);
#pragma unroll
            for (int i = 0; i < 6; ++i) v[u][i] *= k;
        }
        a.neff += valid[u] ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) { a.h[k] = fma(v[u][i], v[u][j], a.h[k]); ++k; }
            a.g[i] = fma(v[u][i], b[u], a.g[i]);
        }
        a.sr2 = fma(r[u], r[u], a.sr2);
        a.sb2 = fma(b[u], b[u], a.sb2);
    }
}

__device__ __forceinline__ double shfl_down_d(double v, int off) {
    return __shfl_down_sync(0xffffffffu, v, off);
}

// Warp tree -> shared staging (one row of kAcc per warp) -> block partial in global memory.
// smem must hold (blockDim.x/32) * kAcc doubles.  All threads of the block must call this.
__device__ __forceinline__ void block_reduce_store(const Acc& a, double* smem, double* block_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double vals[k2::kAccUsed];
#pragma unroll
    for (int i = 0; i < 21; ++i) vals[i] = a.h[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) vals[21 + i] = a.g[i];
    vals[k2::kAccSumR2] = a.sr2;
    vals[k2::kAccNeff] = (double)a.neff;
    vals[k2::kAccNpt] = (double)a.npt;
    vals[k2::kAccSumB2] = a.sb2;
#pragma unroll
    for (int i = 0; i < k2::kAccUsed; ++i) {
        double v = vals[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += shfl_down_d(v, off);
        if (lane == 0) smem[warp * kAcc + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < k2::kAccUsed) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += smem[w * kAcc + threadIdx.x];
        block_out[threadIdx.x] = s;
    }
}

// Congruence with Q = blkdiag(R,R): H_body = Q^T H_world Q, g_body = Q^T g_world.
// in/out: 27 packed values (21 upper + 6 rhs).  Single thread.
__device__ inline void world_to_body(double* v27, const double* R) {
    double H[36], g[6], T[36], Hb[36], gb[6];
    k2::unpack_H(v27, H, g);
    // T = H * Q  (columns in blocks: T[:, 0:3] = H[:, 0:3] R, T[:, 3:6] = H[:, 3:6] R)
    for (int i = 0; i < 6; ++i)
        for (int blk = 0; blk < 2; ++blk)
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += H[i * 6 + blk * 3 + k] * R[k * 3 + j];
                T[i * 6 + blk * 3 + j] = s;
            }
    // Hb = Q^T * T
    for (int blk = 0; blk < 2; ++blk)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * T[(blk * 3 + k) * 6 + j];
                Hb[(blk * 3 + i) * 6 + j] = s;
            }
    for (int blk = 0; blk < 2; ++blk)
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * g[blk * 3 + k];
            gb[blk * 3 + i] = s;
        }
    int k = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) v27[k++] = 0.5 * (Hb[i * 6 + j] + Hb[j * 6 + i]);
    for (int i = 0; i < 6; ++i) v27[21 + i] = gb[i];
}

// Last-block final reduce: sums the per-block partials in block order (deterministic), applies
// the world->body congruence, writes acc_out[kAcc].  Called by every thread of the LAST block.
__device__ __forceinline__ void final_reduce(const double* partials, int nblocks, const double* R,
                                             double* smem, double* acc_out) {
    // thread i < kAccUsed sums column i over blocks with 4 independent chains
    if (threadIdx.x < k2::kAccUsed) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int b = 0;
        for (; b + 3 < nblocks; b += 4) {
            s0 += __ldcg(&partials[(b + 0) * kAcc + threadIdx.x]);
            s1 += __ldcg(&partials[(b + 1) * kAcc + threadIdx.x]);
            s2 += __ldcg(&partials[(b + 2) * kAcc + threadIdx.x]);
            s3 += __ldcg(&partials[(b + 3) * kAcc + threadIdx.x]);
        }
        for (; b < nblocks; ++b) s0 += __ldcg(&partials[b * kAcc + threadIdx.x]);
        smem[threadIdx.x] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v[kAcc];
        for (int i = 0; i < k2::kAccUsed; ++i) v[i] = smem[i];
        v[kAcc - 1] = 0.0;
        world_to_body(v, R);
        for (int i = 0; i < kAcc; ++i) acc_out[i] = v[i];
    }
}

}  // namespace k1
