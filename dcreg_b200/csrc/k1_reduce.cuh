// k1_reduce.cuh - per-slot math and the Gram-matrix reduction shared by the two producers of the normal
// equations: the streaming kernel K1 (k1_stream.cuh: frozen planes) and the fused ICP iteration kernel
// (dcreg_b200.cu: planes come straight out of the correspondence stage).
//
// Replaces, per source slot and per ICP iteration (reference file:line):
//   pointBodyToGlobal (FP64 math, float32 store)         DCReg/include/utils.hpp:630-636
//   residual + LOAM weight + gate                         DCReg/src/icp_test_runner.cpp:1774-1803
//   stream compaction of flagged slots                    icp_test_runner.cpp:1816-1840 (skipped: gated in place)
//   computePointToPlaneJacobian + row fill                DCReg/include/math_utils.hpp:102-121, icp_test_runner.cpp:1863-1907
//   H = A^T A, g = A^T b                                  icp_test_runner.cpp:1910-1919
//   SymmetricHessianComputer (21 + 6 accumulators)        DCReg/include/hessian_computer.h:62-123
//
// Formulation.  The reference's Jacobian row is (s + r ds_dr) [ -n^T R [p]x , n^T R ] with the normal rebuilt
// from the float32-stored coeff = (s n, s r) as n = coeff/s (:1786-1790, 1889, 1898, 1906).  With u' = fl32(s n)
// and w = s + r ds_dr this row equals
//       (w/s) [ (Rp x u')^T , u'^T ] blkdiag(R, R),
// so per slot only the 8 world-frame components c = [k (Rp x u'), k u', b, r] are formed (k = w/s = 1 without
// the weight derivative, 2 - 1/s with it; b = -fl32(s r)), the sums are the 8x8 Gram matrix C = sum c c^T
//       H_world = C[0:6,0:6],  g_world = C[0:6,6],  sum b^2 = C[6,6],  sum r^2 = C[7,7],
// and the constant congruence with blkdiag(R, R) is applied once, in the last block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k2_solve.cuh"

namespace k1 {

using k2::kAcc;

constexpr int kGramPart = 66;            // per-block partial: 64 Gram entries + N_eff + N_corr_pt
constexpr int kTRow = 36;                // padded row stride (doubles) of a warp's DMMA transpose buffer

struct Pose {            // R row-major, t
    double R[9];
    double t[3];
};

// fast full-precision reciprocal for s in (0.1, 1]: MUFU.RCP64H seed + 2 Newton steps (error < 1 ulp)
__device__ __forceinline__ double rcp_newton(double s) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double e = fma(-s, y, 1.0);
    y = fma(y, e, y);
    e = fma(-s, y, 1.0);
    return fma(y, e, y);
}

// (double)(float)x without the two F2F conversions: round-to-nearest-even to 24 significant bits directly on
// the FP64 bit pattern (5 integer instructions).  Identical to the float round trip for x = 0 and for
// 2^-126 <= |x| < 2^128, i.e. whenever the float32 result is a normal number or zero.
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

// Per-slot front: residual, weight, gate, float32 round trips, Jacobian row in the world frame.
// Output c[8] = [k (Rp x u'), k u', b, r], all zeros for an invalid slot (no plane or gated out).
// slope / gate: dcreg_icp_params::weight_slope / weight_gate (0.9 / 0.1 in the reference, icp_test_runner.cpp:1776, 1785)
template <bool kUseWd>
__device__ __forceinline__ void slot_front(const Pose& P, double px, double py, double pz, double nx, double ny,
                                           double nz, double d, bool has, double (&c)[8], int& neff,
                                           double slope = 0.9, double gate = 0.1) {
    const double wx = fma(P.R[2], pz, fma(P.R[1], py, P.R[0] * px));     // Rp (no translation)
    const double wy = fma(P.R[5], pz, fma(P.R[4], py, P.R[3] * px));
    const double wz = fma(P.R[8], pz, fma(P.R[7], py, P.R[6] * px));
    const double qx = round_f32(wx + P.t[0]);                     // utils.hpp:630-636 (float32 store)
    const double qy = round_f32(wy + P.t[1]);
    const double qz = round_f32(wz + P.t[2]);
    const double rr = fma(nx, qx, fma(ny, qy, fma(nz, qz, d)));   // icp_test_runner.cpp:1774
    const double ss = 1.0 - slope * fabs(rr);                     // :1776 (max(0, .) is implied by the gate >= 0)
    const bool valid = has && (ss > gate);                        // :1785
    const double s = valid ? ss : 0.0;
    const double r = valid ? rr : 0.0;
    double ux = round_f32(s * nx);                                // coeff.x/y/z (:1787-1789)
    double uy = round_f32(s * ny);
    double uz = round_f32(s * nz);
    c[6] = -round_f32(s * r);                                     // -coeff.intensity (:1790, 1906)
    c[7] = r;
    if (kUseWd) {                                                 // :1780-1783, 1898: row scale w/s = 2 - 1/s on 0 < s < 1
        const double sw = valid ? ss : 1.0;                       // (s == 1 gives k = 1: no derivative, as in the reference)
        const double k = 2.0 - rcp_newton(sw);
        ux *= k; uy *= k; uz *= k;
    }
    c[0] = wy * uz - wz * uy;                                     // Rp x (k u')
    c[1] = wz * ux - wx * uz;
    c[2] = wx * uy - wy * ux;
    c[3] = ux; c[4] = uy; c[5] = uz;
    neff += valid ? 1 : 0;
}

// ---- Gram accumulation with the FP64 tensor-core instruction (used by the ICP iteration kernel) -----------------
// One mma.sync.m8n8k4 (SASS: DMMA) adds c c^T for 4 slots: A = c (8 components x 4 slots), B = A^T; each lane owns
// only two entries of C (flat index 2*lane, 2*lane + 1), so no per-thread block of 29 accumulators is needed next to
// the register-hungry k-NN / plane-fit code.  The per-slot components are transposed into the fragment layout
// through a 2.3 KB per-warp shared buffer (8 STS.64 + 8 LDS.64 per 32 slots, conflict-free with the padded stride).
// (The streaming kernel uses plain DFMA chains instead: there the FP64 issue slots are the bottleneck and
// 29 DFMA x 2 cycles beat 8 DMMA x 16.4 cycles.)  All 32 lanes must call this convergently.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void gram_accumulate_dmma(double* tb, int lane, const double (&c)[8], double& c0, double& c1,
                                                     double& e0, double& e1) {
    const int rd_off = (lane >> 2) * kTRow + (lane & 3);   // fragment element: component lane/4 of slot lane%4
#pragma unroll
    for (int j = 0; j < 8; ++j) tb[j * kTRow + lane] = c[j];
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; g += 2) {
        const double f0 = tb[rd_off + 4 * g], f1 = tb[rd_off + 4 * g + 4];
        dmma884(c0, c1, f0, f0);
        dmma884(e0, e1, f1, f1);                            // two accumulator pairs: halves the dependent chain
    }
    __syncwarp();
}

// ---- block / grid reduction of the Gram fragments ---------------------------------------------------------------
struct GramSmem {
    double red[8][kGramPart];
    double fin[kGramPart + 6];
    bool is_last;
};

// Every thread of a 256-thread block calls this with its two Gram entries (flat index 2*lane + {0,1} of its warp's
// 8x8 Gram) and its counters.  Block partial -> global; the last block to arrive (atomic ticket) sums the partials
// in a fixed order with its 8 warps in parallel, applies the world->body congruence with 42 threads and writes
// acc_out[kAcc] = 21 upper-triangular entries (hessian_computer.h order) + 6 rhs + {sum r^2, N_eff, N_pt, sum b^2}.
// Deterministic for a given grid size.
__device__ __forceinline__ void finish_block(double c0, double c1, int neff, int npt, GramSmem& gs, double* partials,
                                             unsigned int* counter, const double* R, double* acc_out) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        neff += __shfl_down_sync(0xffffffffu, neff, off);
        npt += __shfl_down_sync(0xffffffffu, npt, off);
    }
    gs.red[warp][2 * lane] = c0;
    gs.red[warp][2 * lane + 1] = c1;
    if (lane == 0) { gs.red[warp][64] = (double)neff; gs.red[warp][65] = (double)npt; }
    __syncthreads();
    if (tid < kGramPart) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += gs.red[w][tid];
        partials[(size_t)blockIdx.x * kGramPart + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int tk = atomicAdd(counter, 1u);
        gs.is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!gs.is_last) return;
    __threadfence();
    {
        const int nb = (int)gridDim.x;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;                     // elements lane, lane + 32, lane + 64 (< kGramPart)
        int b = warp;
        for (; b + 24 < nb; b += 32) {                           // warp w sums blocks w, w+8, ...; 4 blocks per trip
            double t0[4], t1[4], t2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* row = partials + (size_t)(b + u * 8) * kGramPart;
                t0[u] = __ldcg(row + lane);
                t1[u] = __ldcg(row + 32 + lane);
                t2[u] = (lane < kGramPart - 64) ? __ldcg(row + 64 + lane) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s0 += t0[u]; s1 += t1[u]; s2 += t2[u]; }
        }
        for (; b < nb; b += 8) {
            const double* row = partials + (size_t)b * kGramPart;
            s0 += __ldcg(row + lane);
            s1 += __ldcg(row + 32 + lane);
            if (lane < kGramPart - 64) s2 += __ldcg(row + 64 + lane);
        }
        __syncthreads();                                          // red[][] is free again
        gs.red[warp][lane] = s0;
        gs.red[warp][32 + lane] = s1;
        if (lane < kGramPart - 64) gs.red[warp][64 + lane] = s2;
    }
    __syncthreads();
    double* fin = gs.fin;
    if (tid < kGramPart) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += gs.red[w][tid];
        fin[tid] = s;
    }
    __syncthreads();
    // Gram (world frame) -> H_body = Q^T H Q, g_body = Q^T g with Q = blkdiag(R, R): one thread per output entry
    double* outv = &gs.red[0][0];
    if (tid < 36) {
        const int i = tid / 6, j = tid % 6;
        if (j >= i) {
            const int bi = (i / 3) * 3, bj = (j / 3) * 3, ii = i % 3, jj = j % 3;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const double h = 0.5 * (fin[(bi + k) * 8 + bj + l] + fin[(bj + l) * 8 + bi + k]);
                    acc = fma(R[k * 3 + ii] * h, R[l * 3 + jj], acc);
                }
            outv[i * 6 - (i * (i - 1)) / 2 + (j - i)] = acc;      // packed upper-triangular index, row-major
        }
    } else if (tid < 42) {
        const int i = tid - 36, bi = (i / 3) * 3, ii = i % 3;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = fma(R[k * 3 + ii], 0.5 * (fin[(bi + k) * 8 + 6] + fin[6 * 8 + bi + k]), acc);
        outv[21 + i] = acc;
    } else if (tid == 42) {
        outv[k2::kAccSumR2] = fin[7 * 8 + 7];
        outv[k2::kAccNeff] = fin[64];
        outv[k2::kAccNpt] = fin[65];
        outv[k2::kAccSumB2] = fin[6 * 8 + 6];
        outv[kAcc - 1] = 0.0;
    }
    __syncthreads();
    if (tid < kAcc) acc_out[tid] = outv[tid];
    if (tid == 0) *counter = 0u;
}

}  // namespace k1
