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

// One slot.  (px,py,pz) body-frame point, (nx,ny,nz,d) plane in the world frame (FP64 values;
// the float4 variants convert before the call).  Weight rule: icp_test_runner.cpp:1776-1785.
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
    const double s = fmax(0.0, 1.0 - 0.9 * fabs(r));             // :1776
    if (!(has_plane && s > 0.1)) return;                         // :1785
    double ds = 0.0;
    if (use_wd && s < 1.0) ds = (r > 0.0) ? -0.9 : 0.9;          // :1780-1783 (s > 0 holds here)
    // coeff = (s n, s r) is stored as float32 and the normal rebuilt as coeff/s (:1786-1790, 1889, 1906)
    const double inv_s = 1.0 / s;
    const double ux = (double)(float)(s * nx) * inv_s;
    const double uy = (double)(float)(s * ny) * inv_s;
    const double uz = (double)(float)(s * nz) * inv_s;
    const double b = -(double)(float)(s * r);
    const double w = s + r * ds;                                 // :1898
    double v[6];
    v[0] = w * (wy * uz - wz * uy);                              // w * (Rp x n)
    v[1] = w * (wz * ux - wx * uz);
    v[2] = w * (wx * uy - wy * ux);
    v[3] = w * ux; v[4] = w * uy; v[5] = w * uz;
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
