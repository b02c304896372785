// corr.cuh - device correspondence stage: hash-grid exact 5-NN within radius + 5x3 plane fit.
//
// Replaces (reference file:line):
//   ICPContext::setTargetCloud kd-tree build           DCReg/include/utils.hpp:393-424
//   kdtree.nearestKSearch(q, 5) + 5th-NN radius gate   DCReg/src/icp_test_runner.cpp:1720-1726
//   5x3 colPivHouseholderQr plane fit + gates          icp_test_runner.cpp:1727-1773
//
// The reference's accept rule is "the 5th nearest neighbour is closer than the search radius".
// With cubic cells of edge = radius, every point closer than the radius to q lies in the 27 cells
// around q's cell, so an exact 5-NN over those cells reproduces the accept set of the kd-tree
// (SURVEY.md §7 step 6).  Distances are float32 sums of float32 squared differences, as in FLANN's
// L2_Simple functor that PCL's KdTreeFLANN uses; ties are broken by original point index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "small_la.cuh"

namespace corr {

constexpr unsigned long long kEmptyKey = ~0ull;

struct Grid {
    unsigned long long* keys;   // capacity entries, kEmptyKey = free
    int* cell_start;            // capacity
    int* cell_count;            // capacity
    float4* pts;                // n target points grouped by cell: (x, y, z, bit-cast original index)
    unsigned int mask;          // capacity - 1 (capacity is a power of two)
    int n;
    double inv_cell;            // 1 / cell edge
};

__host__ __device__ __forceinline__ int cell_coord(float v, double inv_cell) {
    return (int)floor((double)v * inv_cell);
}

__host__ __device__ __forceinline__ unsigned long long pack_key(int ix, int iy, int iz) {
    const unsigned long long B = 1ull << 20;
    return ((unsigned long long)(ix + (long long)B) << 42) | ((unsigned long long)(iy + (long long)B) << 21) |
           (unsigned long long)(iz + (long long)B);
}

__host__ __device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned int)k;
}

// ---- build -----------------------------------------------------------------------------------
__global__ void grid_insert_kernel(const float4* __restrict__ tgt, int n, Grid g, int* __restrict__ pt_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = tgt[i];
    const unsigned long long key = pack_key(cell_coord(p.x, g.inv_cell), cell_coord(p.y, g.inv_cell),
                                            cell_coord(p.z, g.inv_cell));
    unsigned int slot = hash_key(key) & g.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&g.keys[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) break;
        slot = (slot + 1) & g.mask;
    }
    pt_slot[i] = (int)slot;
    atomicAdd(&g.cell_count[slot], 1);
}

// exclusive scan of int array, three phases (tile sums, scan of tile sums, tile rescan)
constexpr int kScanTile = 2048;   // 256 threads x 8
__global__ void scan_tile_sums_kernel(const int* __restrict__ in, int n, int* __restrict__ tile_sums) {
    __shared__ int sh[256];
    const int base = blockIdx.x * kScanTile;
    int s = 0;
    for (int k = 0; k < 8; ++k) {
        const int idx = base + threadIdx.x * 8 + k;
        if (idx < n) s += in[idx];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sh[0];
}
__global__ void scan_tile_offsets_kernel(int* tile_sums, int ntiles) {   // single block, 1024 threads
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < ntiles ? tile_sums[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (idx < ntiles) tile_sums[idx] = carry + sh[threadIdx.x] - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
}
__global__ void scan_tile_apply_kernel(const int* __restrict__ in, int n, const int* __restrict__ tile_offsets,
                                       int* __restrict__ out) {
    __shared__ int sh[256];
    const int base = blockIdx.x * kScanTile;
    int loc[8];
    int s = 0;
    for (int k = 0; k < 8; ++k) {
        const int idx = base + threadIdx.x * 8 + k;
        loc[k] = idx < n ? in[idx] : 0;
        s += loc[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int run = tile_offsets[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < 8; ++k) {
        const int idx = base + threadIdx.x * 8 + k;
        if (idx < n) out[idx] = run;
        run += loc[k];
    }
}

__global__ void grid_scatter_kernel(const float4* __restrict__ tgt, int n, Grid g, const int* __restrict__ pt_slot,
                                    int* __restrict__ fill) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = pt_slot[i];
    const int pos = g.cell_start[slot] + atomicAdd(&fill[slot], 1);
    float4 p = tgt[i];
    p.w = __int_as_float(i);
    g.pts[pos] = p;
}

// deterministic order inside every cell: sort by original index (insertion sort, cells are small)
__global__ void grid_sort_cells_kernel(Grid g, unsigned int capacity) {
    const unsigned int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= capacity) return;
    const int cnt = g.cell_count[slot];
    if (cnt < 2) return;
    float4* p = g.pts + g.cell_start[slot];
    for (int i = 1; i < cnt; ++i) {
        const float4 v = p[i];
        const int key = __float_as_int(v.w);
        int j = i - 1;
        while (j >= 0 && __float_as_int(p[j].w) > key) { p[j + 1] = p[j]; --j; }
        p[j + 1] = v;
    }
}

// ---- query -----------------------------------------------------------------------------------
struct Knn5 {
    float d2[5];
    int pos[5];     // position in g.pts
    int idx[5];     // original index (tie-break)
};

__device__ __forceinline__ void knn_init(Knn5& k) {
#pragma unroll
    for (int i = 0; i < 5; ++i) { k.d2[i] = 3.0e38f; k.pos[i] = -1; k.idx[i] = 0x7fffffff; }
}

__device__ __forceinline__ void knn_insert(Knn5& k, float d2, int pos, int idx) {
    if (d2 > k.d2[4] || (d2 == k.d2[4] && idx > k.idx[4])) return;
    k.d2[4] = d2; k.pos[4] = pos; k.idx[4] = idx;
#pragma unroll
    for (int i = 4; i > 0; --i) {
        const bool sw = (k.d2[i] < k.d2[i - 1]) || (k.d2[i] == k.d2[i - 1] && k.idx[i] < k.idx[i - 1]);
        if (sw) {
            const float td = k.d2[i]; k.d2[i] = k.d2[i - 1]; k.d2[i - 1] = td;
            const int tp = k.pos[i]; k.pos[i] = k.pos[i - 1]; k.pos[i - 1] = tp;
            const int ti = k.idx[i]; k.idx[i] = k.idx[i - 1]; k.idx[i - 1] = ti;
        }
    }
}

__device__ __forceinline__ void knn_search(const Grid& g, float qx, float qy, float qz, Knn5& k) {
    const int cx = cell_coord(qx, g.inv_cell), cy = cell_coord(qy, g.inv_cell), cz = cell_coord(qz, g.inv_cell);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const unsigned long long key = pack_key(cx + dx, cy + dy, cz + dz);
                unsigned int slot = hash_key(key) & g.mask;
                int start = 0, cnt = 0;
                while (true) {
                    const unsigned long long kk = __ldg(&g.keys[slot]);
                    if (kk == key) { start = __ldg(&g.cell_start[slot]); cnt = __ldg(&g.cell_count[slot]); break; }
                    if (kk == kEmptyKey) break;
                    slot = (slot + 1) & g.mask;
                }
                for (int j = 0; j < cnt; ++j) {
                    const float4 p = __ldg(&g.pts[start + j]);
                    // FLANN L2_Simple: float diff, float accumulate, x then y then z (no FMA contraction)
                    const float ex = __fsub_rn(qx, p.x), ey = __fsub_rn(qy, p.y), ez = __fsub_rn(qz, p.z);
                    float d2 = __fmul_rn(ex, ex);
                    d2 = __fadd_rn(d2, __fmul_rn(ey, ey));
                    d2 = __fadd_rn(d2, __fmul_rn(ez, ez));
                    knn_insert(k, d2, start + j, __float_as_int(p.w));
                }
            }
}

// Plane through the 5 neighbours: least squares of [nb] x = -1, n = x/|x|, d = 1/|x|, gates
// |x| >= min_norm and max_j (n.nb_j + d)^2 < thickness^2 (icp_test_runner.cpp:1727-1773).
// Returns true and (n, d) when a valid plane exists.
__device__ __forceinline__ bool fit_plane(const Grid& g, const Knn5& k, double min_norm, double thickness,
                                          double& nx, double& ny, double& nz, double& d) {
    double A[15], A0[15], b[5], x[3];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = __ldg(&g.pts[k.pos[j]]);
        A0[j * 3 + 0] = A[j * 3 + 0] = (double)p.x;
        A0[j * 3 + 1] = A[j * 3 + 1] = (double)p.y;
        A0[j * 3 + 2] = A[j * 3 + 2] = (double)p.z;
        b[j] = -1.0;
    }
    dla::colpiv_qr_solve<5, 3>(A, b, x);
    const double ps = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (!(ps >= min_norm)) return false;                 // :1752 (also rejects NaN)
    nx = x[0] / ps; ny = x[1] / ps; nz = x[2] / ps; d = 1.0 / ps;
    double worst = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double e = nx * A0[j * 3 + 0] + ny * A0[j * 3 + 1] + nz * A0[j * 3 + 2] + d;
        e *= e;
        worst = fmax(worst, e);
    }
    return worst < thickness * thickness;                // :1772-1773
}

}  // namespace corr
