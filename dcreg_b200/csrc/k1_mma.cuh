// k1_mma.cuh - K1 on sm_100a: per-slot "front" in FP64 vector math, the 29 sums as an 8-component Gram
// matrix accumulated with the FP64 tensor-core instruction mma.sync.m8n8k4 (SASS: DMMA).
//
// Why (measured on B200, see profiles/ and DESIGN.md §K1):
//   * The straightforward kernel keeps 21 + 6 + 2 FP64 accumulators (58 registers) per thread; with the
//     front's temporaries that is 130-150 registers => 8-9 warps per SM.  Dependent FP64 issue on B200 needs
//     ~16+ independent FP64 instructions in flight per SM sub-partition to fill the 16-lane FP64 pipe
//     (tools/microbench.cu), so that kernel was latency-bound at ~130 us for 10 M slots no matter how the loads
//     were staged (LDG, or TMA bulk copies into an mbarrier ring - both tried, see k1_stream.cuh history).
//   * H = A^T A IS the one dense contraction on this path (the reference calls Eigen's GEMM for it,
//     icp_test_runner.cpp:1913-1915).  With the 8 components c = [k v0..v5, b, r] per slot,
//         C = sum_slots c c^T  (8 x 8)  holds  H (6x6 block), g = C[0:6,6], sum b^2 = C[6,6], sum r^2 = C[7,7],
//     and one DMMA consumes 4 slots: A = c (8 x 4 slots), B = A^T.  Each lane owns just TWO accumulators.
//     The per-slot components are transposed into the DMMA fragment layout through a 2.3 KB per-warp shared
//     buffer (8 STS.64 + 8 LDS.64 per 32 slots, conflict-free).
//   * DMMA shares the FP64 pipe with DFMA (16.4 cycles per DMMA per sub-partition, tools/microbench2.cu):
//     8 DMMA + ~35 front ops per 32 slots ~ 200 pipe cycles => ~54 us floor for 10 M slots, against 49 us of HBM
//     time for the 320 MB at the measured 6.48 TB/s.  No tcgen05: that unit has no FP64 kind.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k1_reduce.cuh"

namespace k1m {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kRow = 36;                 // padded row stride (doubles) of the per-warp transpose buffer
constexpr int kPart = 66;                // per-block partial: 64 Gram entries + N_eff + N_with_plane

struct Args {
    const float4* src;
    const void* plane;
    long long n;
    k1::Pose pose;
    double* partials;            // [grid][kPart]
    unsigned int* counter;
    double* acc;                 // [k2::kAcc] final, body frame
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Per-slot front: residual, weight, gate, float32 round trips, Jacobian row in the world frame.
// Output c[8] = [k (Rp x u'), k u', b, r] with zeros for an invalid slot.  Reference lines as in k1_reduce.cuh.
template <bool kUseWd>
__device__ __forceinline__ void slot_front(const k1::Pose& P, double px, double py, double pz, double nx, double ny,
                                           double nz, double d, bool has, double (&c)[8], int& neff) {
    const double wx = fma(P.R[2], pz, fma(P.R[1], py, P.R[0] * px));
    const double wy = fma(P.R[5], pz, fma(P.R[4], py, P.R[3] * px));
    const double wz = fma(P.R[8], pz, fma(P.R[7], py, P.R[6] * px));
    const double qx = k1::round_f32(wx + P.t[0]);                 // utils.hpp:630-636 (float32 store)
    const double qy = k1::round_f32(wy + P.t[1]);
    const double qz = k1::round_f32(wz + P.t[2]);
    const double rr = fma(nx, qx, ny * qy) + fma(nz, qz, d);      // icp_test_runner.cpp:1774
    const double ss = 1.0 - 0.9 * fabs(rr);                       // :1776
    const bool valid = has && (ss > 0.1);                         // :1785
    const double s = valid ? ss : 0.0;
    const double r = valid ? rr : 0.0;
    double ux = k1::round_f32(s * nx);                            // coeff.x/y/z (:1787-1789)
    double uy = k1::round_f32(s * ny);
    double uz = k1::round_f32(s * nz);
    c[6] = -k1::round_f32(s * r);                                 // -coeff.intensity (:1790, 1906)
    c[7] = r;
    if (kUseWd) {                                                 // :1780-1783, 1898: row scale w/s = 2 - 1/s on 0 < s < 1
        const double sw = (valid && ss < 1.0) ? ss : 1.0;
        const double k = 2.0 - k1::rcp_newton(sw);
        ux *= k; uy *= k; uz *= k;
    }
    c[0] = wy * uz - wz * uy;                                     // Rp x (k u')
    c[1] = wz * ux - wx * uz;
    c[2] = wx * uy - wy * ux;
    c[3] = ux; c[4] = uy; c[5] = uz;
    neff += valid ? 1 : 0;
}

// 16-byte asynchronous global -> shared copy (LDGSTS), L1 bypassed: the data is streamed exactly once
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename PlaneT>
__device__ __forceinline__ void plane_to_f64(const PlaneT& v, double& nx, double& ny, double& nz, double& d, bool& has);
template <>
__device__ __forceinline__ void plane_to_f64<float4>(const float4& v, double& nx, double& ny, double& nz, double& d,
                                                     bool& has) {
    has = (v.x != 0.0f) || (v.y != 0.0f) || (v.z != 0.0f);
    nx = k1::f32_to_f64(v.x); ny = k1::f32_to_f64(v.y); nz = k1::f32_to_f64(v.z); d = k1::f32_to_f64(v.w);
}
template <>
__device__ __forceinline__ void plane_to_f64<double4>(const double4& v, double& nx, double& ny, double& nz, double& d,
                                                      bool& has) {
    has = (v.x != 0.0) || (v.y != 0.0) || (v.z != 0.0);
    nx = v.x; ny = v.y; nz = v.z; d = v.w;
}

constexpr int kDepth = 4;        // chunks in flight per warp (each chunk: 32 slots = 1 KB (1.5 KB with FP64 planes))

template <typename PlaneT>
struct Smem {
    float4 rs[kWarpsPerBlock][kDepth][32];     // lane-private ring slots: every lane copies and reads its own slot,
    PlaneT rp[kWarpsPerBlock][kDepth][32];     // so the ring needs no barrier at all, only cp.async.wait_group
    double tbuf[kWarpsPerBlock][8 * kRow];     // per-warp transpose buffers for the DMMA fragments
    double red[kWarpsPerBlock][kPart];
    bool is_last;
};

template <typename PlaneT, bool kUseWd, int kMinBlocks, bool kDmma>
__global__ void __launch_bounds__(kThreads, kMinBlocks) reduce_mma_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<PlaneT>& sm = *reinterpret_cast<Smem<PlaneT>*>(smem_raw);
    auto& red = sm.red;
    bool& is_last = sm.is_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PlaneT* gplane = reinterpret_cast<const PlaneT*>(a.plane);
    double* tb = sm.tbuf[warp];
    const int rd_off = (lane >> 2) * kRow + (lane & 3);   // fragment element (row = lane/4, slot = lane%4)

    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;      // two accumulator pairs: halves the dependent DMMA chain
    double vh[21], vg[6], vr2 = 0.0, vb2 = 0.0;          // vector-accumulate variant (kDmma == false)
#pragma unroll
    for (int i = 0; i < 21; ++i) vh[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) vg[i] = 0.0;
    int neff = 0, npt = 0;
    // chunk = 32 consecutive slots (one per lane); warp w takes chunks w, w + W, w + 2W, ...
    const long long nchunks = (a.n + 31) >> 5;
    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    const long long w0 = (long long)blockIdx.x * kWarpsPerBlock + warp;
    const int my = (w0 < nchunks) ? (int)((nchunks - w0 + wstride - 1) / wstride) : 0;
    // only the globally last chunk can be partial; it is the last chunk of exactly one warp
    const bool owns_last = (my > 0) && (w0 + (long long)(my - 1) * wstride == nchunks - 1);
    const int klast = owns_last ? my - 1 : -1;
    const int last_cnt = (int)(a.n - ((nchunks - 1) << 5));
    const size_t step = (size_t)wstride * 32;
    const float4* gs = a.src + ((size_t)w0 << 5) + lane;          // next element to fetch (this lane)
    const PlaneT* gp = gplane + ((size_t)w0 << 5) + lane;
    float4* ring_s = &sm.rs[warp][0][lane];                       // + 32 per ring slot
    PlaneT* ring_p = &sm.rp[warp][0][lane];

#pragma unroll
    for (int j = 0; j < kDepth; ++j) {
        if (j < my && (j != klast || lane < last_cnt)) {
            cp_async16(ring_s + j * 32, gs);
            cp_async16(ring_p + j * 32, gp);
            if (sizeof(PlaneT) == 32) cp_async16(reinterpret_cast<char*>(ring_p + j * 32) + 16, reinterpret_cast<const char*>(gp) + 16);
        }
        gs += step; gp += step;
        cp_async_commit();                                        // always: uniform group count
    }
    int slot = 0;
    for (int k = 0; k < my; ++k) {
        cp_async_wait<kDepth - 1>();                              // chunk k has landed (this lane's own copies)
        const bool lv = (k != klast) || (lane < last_cnt);
        const float4 p = ring_s[slot * 32];
        const PlaneT pl = ring_p[slot * 32];
        double nx, ny, nz, d;
        bool has;
        plane_to_f64<PlaneT>(pl, nx, ny, nz, d, has);
        const double px = k1::f32_to_f64(p.x), py = k1::f32_to_f64(p.y), pz = k1::f32_to_f64(p.z);
        has = has && lv;
        // values are in registers: refill this ring slot with chunk k + kDepth
        const int kn = k + kDepth;
        if (kn < my && (kn != klast || lane < last_cnt)) {
            cp_async16(ring_s + slot * 32, gs);
            cp_async16(ring_p + slot * 32, gp);
            if (sizeof(PlaneT) == 32) cp_async16(reinterpret_cast<char*>(ring_p + slot * 32) + 16, reinterpret_cast<const char*>(gp) + 16);
        }
        gs += step; gp += step;
        cp_async_commit();
        slot = (slot + 1 == kDepth) ? 0 : slot + 1;
        npt += has ? 1 : 0;
        double c[8];
        slot_front<kUseWd>(a.pose, px, py, pz, nx, ny, nz, d, has, c, neff);
        if (kDmma) {
#pragma unroll
            for (int j = 0; j < 8; ++j) tb[j * kRow + lane] = c[j];
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; g += 2) {
                const double f0 = tb[rd_off + 4 * g], f1 = tb[rd_off + 4 * g + 4];
                dmma884(c0, c1, f0, f0);
                dmma884(e0, e1, f1, f1);
            }
            __syncwarp();
        } else {
            int q = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = i; j < 6; ++j) { vh[q] = fma(c[i], c[j], vh[q]); ++q; }
                vg[i] = fma(c[i], c[6], vg[i]);
            }
            vr2 = fma(c[7], c[7], vr2);
            vb2 = fma(c[6], c[6], vb2);
        }
    }
    cp_async_wait<0>();
    c0 += e0; c1 += e1;
    if (!kDmma) {
        // scatter this lane's 29 sums into the 8x8 Gram layout, one warp-reduced entry at a time
        int q = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = i; j < 8; ++j) {
                double v;
                if (i < 6 && j < 6) v = vh[q++];
                else if (i < 6 && j == 6) v = vg[i];
                else if (i == 6 && j == 6) v = vb2;
                else if (i == 7 && j == 7) v = vr2;
                else continue;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                // entry (i,j) lives in lane (i*8+j)/2, element (i*8+j)&1; mirror to (j,i)
                const int f1 = i * 8 + j, f2 = j * 8 + i;
                if (lane == (f1 >> 1)) { if (f1 & 1) c1 = v; else c0 = v; }
                if (f2 != f1 && lane == (f2 >> 1)) { if (f2 & 1) c1 = v; else c0 = v; }
            }
        }
    }

    // ---- block reduce: Gram fragments (flat index 2*lane + {0,1}) and the two counters ----
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        neff += __shfl_down_sync(0xffffffffu, neff, off);
        npt += __shfl_down_sync(0xffffffffu, npt, off);
    }
    red[warp][2 * lane] = c0;
    red[warp][2 * lane + 1] = c1;
    if (lane == 0) { red[warp][64] = (double)neff; red[warp][65] = (double)npt; }
    __syncthreads();
    if (tid < kPart) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) s += red[w][tid];
        a.partials[(size_t)blockIdx.x * kPart + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int tk = atomicAdd(a.counter, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    // ---- last block: deterministic sum over blocks.  Warp w sums blocks w, w+8, ... (8 loads in flight per lane),
    // then the 8 warp sums are added in warp order: a fixed summation tree for a given grid size. ----
    __threadfence();
    {
        const int nb = (int)gridDim.x;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;                     // elements lane, lane + 32, lane + 64 (< kPart)
        int b = warp;
        for (; b + 3 * kWarpsPerBlock < nb; b += 4 * kWarpsPerBlock) {
            double t0[4], t1[4], t2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* row = a.partials + (size_t)(b + u * kWarpsPerBlock) * kPart;
                t0[u] = __ldcg(row + lane);
                t1[u] = __ldcg(row + 32 + lane);
                t2[u] = (lane < kPart - 64) ? __ldcg(row + 64 + lane) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s0 += t0[u]; s1 += t1[u]; s2 += t2[u]; }
        }
        for (; b < nb; b += kWarpsPerBlock) {
            const double* row = a.partials + (size_t)b * kPart;
            s0 += __ldcg(row + lane);
            s1 += __ldcg(row + 32 + lane);
            if (lane < kPart - 64) s2 += __ldcg(row + 64 + lane);
        }
        __syncthreads();                                          // red[][] is free again
        red[warp][lane] = s0;
        red[warp][32 + lane] = s1;
        if (lane < kPart - 64) red[warp][64 + lane] = s2;
    }
    __syncthreads();
    double* fin = &sm.tbuf[0][0];
    if (tid < kPart) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) s += red[w][tid];
        fin[tid] = s;
    }
    __syncthreads();
    // Gram (world frame) -> H_body = Q^T H Q, g_body = Q^T g with Q = blkdiag(R, R): one thread per output entry
    // (36 + 6 threads, 9 / 3 FMAs each) instead of a ~300-FMA serial chain on one thread.
    double* outv = &sm.red[0][0];                                 // 27 + stats, reuse
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
                    acc = fma(a.pose.R[k * 3 + ii] * h, a.pose.R[l * 3 + jj], acc);
                }
            // packed upper-triangular index of (i, j), row-major (hessian_computer.h:62-123 order)
            const int idx = i * 6 - (i * (i - 1)) / 2 + (j - i);
            outv[idx] = acc;
        }
    } else if (tid < 42) {
        const int i = tid - 36, bi = (i / 3) * 3, ii = i % 3;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = fma(a.pose.R[k * 3 + ii], 0.5 * (fin[(bi + k) * 8 + 6] + fin[6 * 8 + bi + k]), acc);
        outv[21 + i] = acc;
    } else if (tid == 42) {
        outv[k2::kAccSumR2] = fin[7 * 8 + 7];
        outv[k2::kAccNeff] = fin[64];
        outv[k2::kAccNpt] = fin[65];
        outv[k2::kAccSumB2] = fin[6 * 8 + 6];
        outv[k2::kAcc - 1] = 0.0;
    }
    __syncthreads();
    if (tid < k2::kAcc) a.acc[tid] = outv[tid];
    if (tid == 0) *a.counter = 0u;
}

}  // namespace k1m
