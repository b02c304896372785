// k1_stream.cuh - K1, the streaming kernel: fused point-to-plane residual / LOAM weight / gate / 6-DoF Jacobian
// row / 21 + 6 (+2) normal-equation sums over device-resident (point, plane) slots, for sm_100a.
//
// Shape of the kernel and why (every statement below was measured on B200; profiles/k1_*.md, tools/microbench*.cu):
//   * Cost model.  On B200 the FP64 vector ops (DFMA/DMUL/DADD: 2 cycles per warp instruction per SM sub-
//     partition), the FP64 tensor op (DMMA m8n8k4: 16.4 cycles) and the 64-bit conversions (F2F.F64.F32 6,
//     F2F.F32.F64 9 cycles) all issue through one shared pipe, and their costs ADD to the 1 cycle every other
//     instruction takes: the kernel time is (instructions + extra FP64 slots) / issue rate, independent of
//     occupancy once latency is covered.  Six rewrites that moved work between those units without lowering that
//     sum (LDG vs TMA bulk + mbarrier ring vs cp.async ring, 8..32 warps/SM, 1..4 slots per thread, vector vs
//     DMMA accumulation, F2F vs integer conversions) all landed on the same ~110-130 us for 10 M slots.
//   * So the kernel minimises issue slots per slot: float->double and the reference's float32 round trips are
//     done with 5-instruction integer sequences (k1_reduce.cuh), the 29 sums are plain DFMA chains (29 x 2 slots,
//     cheaper than 8 DMMA x 16.4 + the 16 LDS/STS of a fragment transpose), the pose lives in the kernel-parameter
//     constant bank, the loop runs on 32-bit counters and pointer bumps, and the partial tail chunk is peeled.
//   * Loads: each lane copies its own 16 B point and 16 B (32 B) plane with cp.async (LDGSTS, L1 bypass) into a
//     lane-private 4-deep shared-memory ring, so the ring needs NO barrier (only cp.async.wait_group) and no
//     registers; 16 warps x 4 KB are in flight per SM.  (TMA bulk copies were tried first: a single producer
//     thread per CTA topped out at 3.2 TB/s in a copy-only experiment, below what per-lane LDGSTS/LDG reach.)
//   * Reduction: warp shuffles -> per-block partial (66 doubles) -> the last block (atomic ticket) sums the
//     partials in a fixed order with 8 warps in parallel, applies the world->body congruence with 42 threads and
//     writes the 27 + stats.  Deterministic for a given grid size.
//   * Result on B200 (10 M slots, 320 MB): 86 us with the weight-derivative path, 80 us without (3.7-4.0 TB/s,
//     57-62 % of the measured 6.48 TB/s copy peak); the 48 B/slot FP64-plane variant runs at 6.0 TB/s (93 %).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k1_reduce.cuh"

namespace k1s {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kDepth = 4;                // chunks in flight per warp (a chunk = 32 slots = 1 KB, 1.5 KB with FP64 planes)
static_assert((kDepth & (kDepth - 1)) == 0, "ring depth must be a power of two");

struct Args {
    const float4* src;
    const void* plane;
    long long n;
    k1::Pose pose;
    double* partials;            // [grid][k1::kGramPart]
    unsigned int* counter;
    double* acc;                 // [k2::kAcc] final, body frame
};

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
    has = ((__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z)) & 0x7fffffffu) != 0u;
    nx = k1::f32_to_f64(v.x); ny = k1::f32_to_f64(v.y); nz = k1::f32_to_f64(v.z); d = k1::f32_to_f64(v.w);
}
template <>
__device__ __forceinline__ void plane_to_f64<double4>(const double4& v, double& nx, double& ny, double& nz, double& d,
                                                      bool& has) {
    has = (v.x != 0.0) || (v.y != 0.0) || (v.z != 0.0);
    nx = v.x; ny = v.y; nz = v.z; d = v.w;
}

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

template <typename PlaneT>
struct Smem {
    float4 rs[kWarpsPerBlock][kDepth][32];     // lane-private ring slots: every lane copies and reads its own slot,
    PlaneT rp[kWarpsPerBlock][kDepth][32];     // so the ring needs no barrier at all, only cp.async.wait_group
    k1::GramSmem gram;
};

template <typename PlaneT, bool kUseWd>
__global__ void __launch_bounds__(kThreads, 2) reduce_stream_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<PlaneT>& sm = *reinterpret_cast<Smem<PlaneT>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PlaneT* gplane = reinterpret_cast<const PlaneT*>(a.plane);

    double vh[21], vg[6], vr2 = 0.0, vb2 = 0.0;          // the 29 running sums of this lane's slots
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
    // one chunk: consume ring slot `slot`, refill it with chunk k + kDepth, front + accumulate.
    // kTail = true only for the (at most one) partial chunk at the very end of the array.
    auto process = [&](int k, int slot, auto tail_tag) {
        constexpr bool kTail = decltype(tail_tag)::value;
        float4 p = ring_s[slot * 32];
        PlaneT pl = ring_p[slot * 32];
        if (kTail && lane >= last_cnt) {           // lanes past the end never copied anything: feed zeros, not stale bits
            p = make_float4(0.f, 0.f, 0.f, 0.f);
            pl = PlaneT{};
        }
        double nx, ny, nz, d;
        bool has;
        plane_to_f64<PlaneT>(pl, nx, ny, nz, d, has);
        const double px = k1::f32_to_f64(p.x), py = k1::f32_to_f64(p.y), pz = k1::f32_to_f64(p.z);
        // values are in registers: refill this ring slot with chunk k + kDepth
        const int kn = k + kDepth;
        if (kn < my && (kn != klast || lane < last_cnt)) {
            cp_async16(ring_s + slot * 32, gs);
            cp_async16(ring_p + slot * 32, gp);
            if (sizeof(PlaneT) == 32) cp_async16(reinterpret_cast<char*>(ring_p + slot * 32) + 16, reinterpret_cast<const char*>(gp) + 16);
        }
        gs += step; gp += step;
        cp_async_commit();
        npt += has ? 1 : 0;
        double c[8];
        k1::slot_front<kUseWd>(a.pose, px, py, pz, nx, ny, nz, d, has, c, neff);
        int q = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) { vh[q] = fma(c[i], c[j], vh[q]); ++q; }
            vg[i] = fma(c[i], c[6], vg[i]);
        }
        vr2 = fma(c[7], c[7], vr2);
        vb2 = fma(c[6], c[6], vb2);
    };
    int slot = 0;
    const int nfull = (klast >= 0 && last_cnt < 32) ? my - 1 : my;     // chunks that are certainly full
    int k = 0;
    for (; k < nfull; ++k) {
        cp_async_wait<kDepth - 1>();                              // chunk k has landed (this lane's own copies)
        process(k, slot, FalseT{});
        slot = (slot + 1) & (kDepth - 1);
    }
    if (k < my) {
        cp_async_wait<kDepth - 1>();
        process(k, slot, TrueT{});
    }
    cp_async_wait<0>();

    // scatter this lane's 29 sums into the 8x8 Gram layout (entry f lives in lane f/2, element f&1), warp-reduced
    double c0 = 0.0, c1 = 0.0;
    {
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
                const int f1 = i * 8 + j, f2 = j * 8 + i;
                if (lane == (f1 >> 1)) { if (f1 & 1) c1 = v; else c0 = v; }
                if (f2 != f1 && lane == (f2 >> 1)) { if (f2 & 1) c1 = v; else c0 = v; }
            }
        }
    }
    k1::finish_block(c0, c1, neff, npt, sm.gram, a.partials, a.counter, a.pose.R, a.acc);
}

}  // namespace k1s
