// corr.cuh - device correspondence stage: uniform-grid exact 5-NN within radius + 5x3 plane fit.
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
//
// Layout in HBM: target points grouped by cell (float4: x, y, z, bit-cast original index).
//   dense mode : cells of the target's bounding box in x-fastest linear order + cell_start[ncells + 1];
//                the 3 x-adjacent cells of a row are ONE contiguous point range, so a query scans 9 ranges.
//   hash mode  : open-addressing table keyed by the packed cell coordinates (fallback when the bounding box
//                has more than kMaxDenseCells cells); a query probes 27 cells.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "small_la.cuh"

namespace corr {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr long long kMaxDenseCells = 1ll << 27;

struct Grid {
    float4* pts;                // n target points grouped by cell
    int n;
    int dense;                  // 1: dense mode, 0: hash mode
    int rings;                  // ceil(search radius / cell edge): cells per direction a query must look at
    double inv_cell;            // 1 / cell edge
    // dense mode
    int ox, oy, oz;             // cell coordinates of the bounding box's minimum corner
    int nx, ny, nz;
    int* cell_start;            // [nx*ny*nz + 1]
    // hash mode
    unsigned long long* keys;   // capacity entries, kEmptyKey = free
    int* hstart;                // capacity
    int* hcount;                // capacity
    unsigned int mask;          // capacity - 1 (capacity is a power of two)
};

__host__ __device__ __forceinline__ int cell_coord(float v, double inv_cell) {
    return (int)floor((double)v * inv_cell);
}

__host__ __device__ __forceinline__ unsigned long long pack_key(int ix, int iy, int iz) {
    const long long B = 1ll << 20;
    return ((unsigned long long)(ix + B) << 42) | ((unsigned long long)(iy + B) << 21) | (unsigned long long)(iz + B);
}

__host__ __device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned int)k;
}

__device__ __forceinline__ int dense_index(const Grid& g, int cx, int cy, int cz) {
    return ((cz - g.oz) * g.ny + (cy - g.oy)) * g.nx + (cx - g.ox);
}

// ---- build -----------------------------------------------------------------------------------
// bounds[0..2] = min cell coords, bounds[3..5] = max cell coords (initialised to +-2^30 by the host)
__global__ void grid_bounds_kernel(const float4* __restrict__ pts, int n, double inv_cell, int* __restrict__ bounds) {
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const int c[3] = {cell_coord(p.x, inv_cell), cell_coord(p.y, inv_cell), cell_coord(p.z, inv_cell)};
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], c[k]); hi[k] = max(hi[k], c[k]); }
    }
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], lo[k]); atomicMax(&bounds[3 + k], hi[k]); }
    }
}

// cell id of a point: dense linear index (clamped into the box when `clamp`), or hash slot (insert mode)
__global__ void grid_count_dense_kernel(const float4* __restrict__ pts, int n, Grid g, int* __restrict__ pt_cell,
                                        int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const int c = dense_index(g, cell_coord(p.x, g.inv_cell), cell_coord(p.y, g.inv_cell), cell_coord(p.z, g.inv_cell));
    pt_cell[i] = c;
    atomicAdd(&counts[c], 1);
}

__global__ void grid_insert_hash_kernel(const float4* __restrict__ pts, int n, Grid g, int* __restrict__ pt_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const unsigned long long key = pack_key(cell_coord(p.x, g.inv_cell), cell_coord(p.y, g.inv_cell),
                                            cell_coord(p.z, g.inv_cell));
    unsigned int slot = hash_key(key) & g.mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&g.keys[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) break;
        slot = (slot + 1) & g.mask;
    }
    pt_slot[i] = (int)slot;
    atomicAdd(&g.hcount[slot], 1);
}

// exclusive scan of an int array, three phases (tile sums, scan of tile sums, tile rescan)
constexpr int kScanTile = 2048;   // 256 threads x 8
__global__ void scan_tile_sums_kernel(const int* __restrict__ in, int n, int* __restrict__ tile_sums) {
    __shared__ int sh[256];
    const long long base = (long long)blockIdx.x * kScanTile;
    int s = 0;
    for (int k = 0; k < 8; ++k) {
        const long long idx = base + threadIdx.x * 8 + k;
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
    const long long base = (long long)blockIdx.x * kScanTile;
    int loc[8];
    int s = 0;
    for (int k = 0; k < 8; ++k) {
        const long long idx = base + threadIdx.x * 8 + k;
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
        const long long idx = base + threadIdx.x * 8 + k;
        if (idx < n) out[idx] = run;
        run += loc[k];
    }
}

// scatter points into their cell's range; `start` is the exclusive scan of the per-cell counts
__global__ void grid_scatter_kernel(const float4* __restrict__ pts, int n, const int* __restrict__ pt_cell,
                                    const int* __restrict__ start, int* __restrict__ fill, float4* __restrict__ out,
                                    int keep_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = pt_cell[i];
    const int pos = start[c] + atomicAdd(&fill[c], 1);
    float4 p = pts[i];
    if (!keep_w) p.w = __int_as_float(i);
    out[pos] = p;
}

// deterministic order inside every cell: sort by original index (insertion sort, cells are small)
__global__ void grid_sort_cells_kernel(float4* pts, const int* __restrict__ start, const int* __restrict__ count,
                                       const int* __restrict__ start_next, long long ncells) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int s = start[c];
    const int cnt = count ? count[c] : (start_next[c] - s);
    if (cnt < 2) return;
    float4* p = pts + s;
    for (int i = 1; i < cnt; ++i) {
        const float4 v = p[i];
        const int key = __float_as_int(v.w);
        int j = i - 1;
        while (j >= 0 && __float_as_int(p[j].w) > key) { p[j + 1] = p[j]; --j; }
        p[j + 1] = v;
    }
}

// cell of a (transformed) source point for the spatial sort of the source cloud, clamped into the target box
__global__ void source_cell_kernel(const float4* __restrict__ src, int n, Grid g, const double* __restrict__ T,
                                   int* __restrict__ pt_cell, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = src[i];
    const double px = p.x, py = p.y, pz = p.z;
    const float qx = (float)(T[0] * px + T[1] * py + T[2] * pz + T[3]);
    const float qy = (float)(T[4] * px + T[5] * py + T[6] * pz + T[7]);
    const float qz = (float)(T[8] * px + T[9] * py + T[10] * pz + T[11]);
    const int cx = min(max(cell_coord(qx, g.inv_cell) - g.ox, 0), g.nx - 1);
    const int cy = min(max(cell_coord(qy, g.inv_cell) - g.oy, 0), g.ny - 1);
    const int cz = min(max(cell_coord(qz, g.inv_cell) - g.oz, 0), g.nz - 1);
    const int c = (cz * g.ny + cy) * g.nx + cx;
    pt_cell[i] = c;
    atomicAdd(&counts[c], 1);
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

// FLANN L2_Simple: float differences, float accumulation, x then y then z (no FMA contraction)
__device__ __forceinline__ float dist2(float qx, float qy, float qz, const float4& p) {
    const float ex = __fsub_rn(qx, p.x), ey = __fsub_rn(qy, p.y), ez = __fsub_rn(qz, p.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

__device__ __forceinline__ void knn_scan_range(const float4* __restrict__ pts, int s, int e, float qx, float qy,
                                               float qz, Knn5& k) {
    int j = s;
    for (; j + 1 < e; j += 2) {                       // two candidates per trip: both loads in flight
        const float4 p0 = __ldg(&pts[j]), p1 = __ldg(&pts[j + 1]);
        const float a0 = dist2(qx, qy, qz, p0), a1 = dist2(qx, qy, qz, p1);
        const int i0 = __float_as_int(p0.w), i1 = __float_as_int(p1.w);
        if (a0 < k.d2[4] || (a0 == k.d2[4] && i0 < k.idx[4])) knn_insert(k, a0, j, i0);
        if (a1 < k.d2[4] || (a1 == k.d2[4] && i1 < k.idx[4])) knn_insert(k, a1, j + 1, i1);
    }
    if (j < e) {
        const float4 p0 = __ldg(&pts[j]);
        const float a0 = dist2(qx, qy, qz, p0);
        const int i0 = __float_as_int(p0.w);
        if (a0 < k.d2[4] || (a0 == k.d2[4] && i0 < k.idx[4])) knn_insert(k, a0, j, i0);
    }
}

__device__ __forceinline__ void knn_search(const Grid& g, float qx, float qy, float qz, Knn5& k) {
    const int cx = cell_coord(qx, g.inv_cell), cy = cell_coord(qy, g.inv_cell), cz = cell_coord(qz, g.inv_cell);
    if (g.dense) {
        // K = ceil(radius / cell) rings of cells cover the search radius (K = 1 when cell = radius; a finer grid,
        // cell = radius / 2, K = 2, scans ~3x fewer candidates because most cells are pruned by their box distance).
        // Rows (dy, dz) are visited in rings of growing max(|dy|, |dz|); inside a row the own column first, then
        // outwards.  A row / cell is skipped only when its box distance exceeds the current 5th-best distance
        // (strictly, with a 1e-5 relative margin for the float arithmetic of the bound), so the search stays exact.
        // Loops are deliberately NOT unrolled: one copy of the scan loop keeps the kernel inside the instruction
        // cache (the fully unrolled version stalled on instruction fetch, profiles/icp_iteration_r1).
        const int K = g.rings;
        const float cell = (float)(1.0 / g.inv_cell);
        const int lx = cx - g.ox, ly = cy - g.oy, lz = cz - g.oz;
        if (lx + K < 0 || lx - K >= g.nx) return;
        // position inside the own cell, in [0, cell) up to float rounding of cx * cell: every gap below is shrunk
        // by an absolute eps that covers that rounding, so a bound can only be too small (never prunes a hit)
        const float eps = 2e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + cell);
        const float fx = qx - (float)cx * cell, fy = qy - (float)cy * cell, fz = qz - (float)cz * cell;
#pragma unroll 1
        for (int ring = 0; ring <= K; ++ring) {
            // a whole ring is at least (ring - 1) * cell + (distance to the own cell's nearest face) away
            if (ring > 1) {
                const float m = fmaxf(fminf(fminf(fy, cell - fy), fminf(fz, cell - fz)) + (float)(ring - 1) * cell - eps, 0.0f);
                if (m * m * 0.99999f > k.d2[4]) break;
            }
#pragma unroll 1
            for (int dz = -ring; dz <= ring; ++dz) {
                const int zz = lz + dz;
                if (zz < 0 || zz >= g.nz) continue;
                const float gz = dz == 0 ? 0.0f : fmaxf((dz < 0 ? fz + (float)(-dz - 1) * cell : (cell - fz) + (float)(dz - 1) * cell) - eps, 0.0f);
                const int stepy = (dz == -ring || dz == ring) ? 1 : 2 * ring;     // only the ring's boundary rows
#pragma unroll 1
                for (int dy = -ring; dy <= ring; dy += (stepy > 0 ? stepy : 1)) {
                    const int yy = ly + dy;
                    if (yy < 0 || yy >= g.ny) continue;
                    const float gy = dy == 0 ? 0.0f : fmaxf((dy < 0 ? fy + (float)(-dy - 1) * cell : (cell - fy) + (float)(dy - 1) * cell) - eps, 0.0f);
                    const float row_lb = (gy * gy + gz * gz) * 0.99999f;
                    if (row_lb > k.d2[4]) continue;
                    const int* rowp = g.cell_start + (size_t)(zz * g.ny + yy) * g.nx;
                    // own column, then +-1, +-2, ... : stop a side once its gap bound exceeds the 5th-best distance
                    {
                        const int x0 = min(max(lx, 0), g.nx), x1 = min(max(lx + 1, 0), g.nx);
                        knn_scan_range(g.pts, __ldg(rowp + x0), __ldg(rowp + x1), qx, qy, qz, k);
                    }
#pragma unroll 1
                    for (int dx = 1; dx <= K; ++dx) {
                        const float gl = fmaxf(fx + (float)(dx - 1) * cell - eps, 0.0f);
                        const float gr = fmaxf((cell - fx) + (float)(dx - 1) * cell - eps, 0.0f);
                        const bool left = (row_lb + gl * gl * 0.99999f) <= k.d2[4];
                        if (left) {
                            const int x0 = min(max(lx - dx, 0), g.nx), x1 = min(max(lx - dx + 1, 0), g.nx);
                            knn_scan_range(g.pts, __ldg(rowp + x0), __ldg(rowp + x1), qx, qy, qz, k);
                        }
                        const bool right = (row_lb + gr * gr * 0.99999f) <= k.d2[4];
                        if (right) {
                            const int x0 = min(max(lx + dx, 0), g.nx), x1 = min(max(lx + dx + 1, 0), g.nx);
                            knn_scan_range(g.pts, __ldg(rowp + x0), __ldg(rowp + x1), qx, qy, qz, k);
                        }
                        if (!left && !right) break;
                    }
                }
            }
        }
    } else {
        const int K = g.rings;
        for (int dz = -K; dz <= K; ++dz)
            for (int dy = -K; dy <= K; ++dy)
                for (int dx = -K; dx <= K; ++dx) {
                    const unsigned long long key = pack_key(cx + dx, cy + dy, cz + dz);
                    unsigned int slot = hash_key(key) & g.mask;
                    int start = 0, cnt = 0;
                    while (true) {
                        const unsigned long long kk = __ldg(&g.keys[slot]);
                        if (kk == key) { start = __ldg(&g.hstart[slot]); cnt = __ldg(&g.hcount[slot]); break; }
                        if (kk == kEmptyKey) break;
                        slot = (slot + 1) & g.mask;
                    }
                    knn_scan_range(g.pts, start, start + cnt, qx, qy, qz, k);
                }
    }
}

// ---- cooperative exact 5-NN (dense grid): 8 lanes per query --------------------------------------------------------
// The fused one-thread-per-query search is latency- and divergence-bound at ICP sizes (100 k queries = 5 warps per SM
// sub-partition, ~8 k dependent instructions each, most of them sorted-list insertions with 2 of 32 lanes active).
// Here a query is searched by 8 adjacent lanes:
//   * bound.  Only points that can be among the five nearest need sorting.  The accept rule needs the five nearest
//     only if the 5th is inside the search radius, so B = fl_up(radius^2) is always a valid bound; and the previous
//     iteration's five neighbours of the same source point are five distinct target points, so the largest of their
//     distances to the new query bounds the new 5th distance from above (inclusive - ties are kept and resolved by
//     the index rule).  The seeds only provide the bound; the scan itself finds every point again.
//   * rows.  The (2K+1)^2 cell rows of the neighbourhood (own row first) are set up by the lanes in parallel:
//     one contiguous point range per row, end cells and whole rows dropped when their box distance exceeds B.
//   * scan.  The lanes walk the concatenation of the ranges with stride 8 (coalesced 128 B loads, same trip count)
//     and append every candidate with d2 <= B to a 32-entry shared-memory buffer.
//   * select.  Rank of an entry = number of entries before it in (d2, index) order; ranks 0..4 are the answer.
//     If more than 32 candidates survived (no seeds yet: first iteration), B is lowered to the 5th smallest of the
//     32 buffered ones - again five distinct real points - and the scan is repeated; if that cannot lower B (32+
//     points tied at the bound) one lane runs the sequential exact search.
//   * skip.  Every scan also yields a lower bound lb6 on the distance of everything OUTSIDE the five (the 6th
//     buffered candidate, else the bound itself) and remembers where the query was (q_scan).  Later iterations move the
//     query by delta = |q - q_scan|; while  max_i |q - nb_i| + delta < lb6  (with margins that dwarf the float32
//     evaluation error of a squared distance) no outside point can have overtaken a neighbour, so the five are only
//     re-ranked by their new distances and no cell is touched.  Near convergence almost every query takes this path.
// Record per query (3 int4): {pos0..pos3}, {pos4, bits(d2 of the 5th), bits(lb6), flags}, {bits(q_scan.xyz), 0};
// pos = position in g.pts, -1 = none; flags bit 0 = the ordered list equals the previous iteration's (the plane fitted
// to it can be reused: same five rows in the same order give the same QR bit for bit).
constexpr int kKnnRec = 3;
constexpr float kKnnLook = 1.21f;                    // squared-distance look-ahead factor beyond the bound (any value >= 1 is exact)
constexpr int kKnnLanes = 8;
constexpr int kKnnCap = 32;
constexpr int kKnnMaxRows = 81;                     // (2 * 4 + 1)^2: launch_iteration caps the ring count at 4
constexpr int kKnnBlock = 256;
constexpr int kKnnGroups = kKnnBlock / kKnnLanes;

struct KnnArgs {
    const float4* src;        // queries before the pose (sorted source)
    long long n;
    Grid grid;
    const double* pose_R;     // device pointers into the loop state: R[9] row-major, t[3]
    const double* pose_t;
    const int* done;          // loop-state flag: skip when set
    int4* nn;                 // [kKnnRec n] in: previous record (when use_seeds), out: new record
    int use_seeds;
    float r2_up;              // radius^2 rounded up to float
};

struct KnnGroupSmem {
    int rs[kKnnMaxRows], re[kKnnMaxRows];           // point range of every row (empty when pruned)
    float d2[kKnnCap];
    int pos[kKnnCap], idx[kKnnCap];
    int cnt;
    int out[7];                                     // pos0..pos4, bits(d2 of the 5th), bits(d2 of the 6th candidate)
};

__device__ __forceinline__ bool knn_less(float d, int i, float d_ref, int i_ref) {
    return d < d_ref || (d == d_ref && i < i_ref);
}

// The four groups of a warp run every loop in lock-step (trip counts = the maximum over the groups, work predicated
// per group): a warp that lets its groups drift apart executes each group's instructions separately (measured: 9 of
// 32 lanes active on average), which costs more than the idle trips.
template <int kRings>                                             // 1: cell edge = search radius; 0: 1..4 rings at run time
__global__ void __launch_bounds__(kKnnBlock) knn5_kernel(const __grid_constant__ KnnArgs a) {
    __shared__ KnnGroupSmem sm[kKnnGroups];
    if (*a.done) return;
    const Grid& g = a.grid;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x / kKnnLanes, sub = threadIdx.x & (kKnnLanes - 1);
    const int gshift = lane & ~(kKnnLanes - 1);                    // first lane of my group inside the warp
    const unsigned below = (1u << sub) - 1u;                       // group-relative mask of the lanes before me
    const long long qi = (long long)blockIdx.x * kKnnGroups + grp;
    const bool active = qi < a.n;
    KnnGroupSmem& S = sm[grp];
    if (sub < 7) S.out[sub] = sub < 5 ? -1 : __float_as_int(3.0e38f);
    __syncwarp();
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float4 p4 = __ldg(&a.src[qi]);
        const double px = p4.x, py = p4.y, pz = p4.z;
        // q = fl32(R p + t)  (utils.hpp:630-636), same expression as the fused kernel
        qx = (float)(a.pose_R[0] * px + a.pose_R[1] * py + a.pose_R[2] * pz + a.pose_t[0]);
        qy = (float)(a.pose_R[3] * px + a.pose_R[4] * py + a.pose_R[5] * pz + a.pose_t[1]);
        qz = (float)(a.pose_R[6] * px + a.pose_R[7] * py + a.pose_R[8] * pz + a.pose_t[2]);
    }
    // ---- bound from the search radius and from the previous neighbours; skip test
    float B = a.r2_up;
    int4 s0 = make_int4(-1, -1, -1, -1), s1 = make_int4(-1, 0, 0, 0);
    bool skip = false;
    if (a.use_seeds) {
        bool have = false;
        float ds = 0.0f;
        int mine = -1, mine_idx = 0x7fffffff;
        float4 qs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            s0 = a.nn[kKnnRec * qi]; s1 = a.nn[kKnnRec * qi + 1];
            have = s1.x >= 0;                                      // group-uniform
            if (have) {
                const int4 s2 = a.nn[kKnnRec * qi + 2];
                qs = make_float4(__int_as_float(s2.x), __int_as_float(s2.y), __int_as_float(s2.z), 0.f);
                if (sub < 5) {
                    mine = sub == 0 ? s0.x : (sub == 1 ? s0.y : (sub == 2 ? s0.z : (sub == 3 ? s0.w : s1.x)));
                    const float4 p = __ldg(&g.pts[mine]);
                    ds = dist2(qx, qy, qz, p);
                    mine_idx = __float_as_int(p.w);
                }
            }
        }
        const float dmine = ds;
#pragma unroll
        for (int off = 1; off < kKnnLanes; off <<= 1) ds = fmaxf(ds, __shfl_xor_sync(full, ds, off));
        if (have) {
            // scan a little beyond the five (10 % in distance): that is what finds the 6th candidate / a gap to it
            B = fminf(B, ds * kKnnLook);
            // nothing outside the five was closer than sqrt(lb6) to q_scan; it is now at least sqrt(lb6) - delta away
            const float ex = qx - qs.x, ey = qy - qs.y, ez = qz - qs.z;
            const float delta = sqrtf(ex * ex + ey * ey + ez * ez);
            const float lb6 = sqrtf(__int_as_float(s1.z));
            skip = (sqrtf(ds) + delta) * 1.00002f + 1e-7f < lb6 * 0.99998f && ds < a.r2_up;
        }
        // re-rank the five by their new distances (index rule on ties); done by every group, used when skip
        int rank = 0;
#pragma unroll
        for (int o = 1; o < 5; ++o) {
            const int from = gshift + (sub + o) % 5;
            const float od = __shfl_sync(full, dmine, from);
            const int oi = __shfl_sync(full, mine_idx, from);
            rank += knn_less(od, oi, dmine, mine_idx) ? 1 : 0;
        }
        if (skip && sub < 5) {
            S.out[rank] = mine;
            if (rank == 4) S.out[5] = __float_as_int(dmine);
        }
    }
    // ---- rows: one point range each, own row first; only the non-empty ones are kept (compacted per group)
    const int K = kRings > 0 ? kRings : g.rings, W = 2 * K + 1, nrows = W * W, centre = (nrows - 1) >> 1;
    int nr = 0;                                                    // rows of my group (group-uniform)
    {
        const float cell = (float)(1.0 / g.inv_cell);
        const int cx = cell_coord(qx, g.inv_cell), cy = cell_coord(qy, g.inv_cell), cz = cell_coord(qz, g.inv_cell);
        const int lx = cx - g.ox, ly = cy - g.oy, lz = cz - g.oz;
        // position inside the own cell; gaps are shrunk by an absolute eps that covers the float rounding of
        // cx * cell, so a bound can only be too small (it never prunes a hit)
        const float eps = 2e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + cell);
        const float fx = qx - (float)cx * cell, fy = qy - (float)cy * cell, fz = qz - (float)cz * cell;
#pragma unroll 1
        for (int r0 = 0; r0 < nrows; r0 += kKnnLanes) {
            const int r = r0 + sub;
            int s = 0, e = 0;
            if (active && !skip && r < nrows) {
                const int rr = r == 0 ? centre : (r == centre ? 0 : r);
                const int rz = rr / W;
                const int dz = rz - K, dy = rr - rz * W - K;
                const int zz = lz + dz, yy = ly + dy;
                if (zz >= 0 && zz < g.nz && yy >= 0 && yy < g.ny) {
                    const float gz = dz == 0 ? 0.0f : fmaxf((dz < 0 ? fz + (float)(-dz - 1) * cell : (cell - fz) + (float)(dz - 1) * cell) - eps, 0.0f);
                    const float gy = dy == 0 ? 0.0f : fmaxf((dy < 0 ? fy + (float)(-dy - 1) * cell : (cell - fy) + (float)(dy - 1) * cell) - eps, 0.0f);
                    const float row_lb = (gy * gy + gz * gz) * 0.99999f;
                    if (row_lb <= B) {
                        int xa = lx - K, xb = lx + K;              // drop end cells whose box distance exceeds the bound
#pragma unroll 1
                        for (; xa < lx; ++xa) {
                            const float gl = fmaxf(fx + (float)(lx - xa - 1) * cell - eps, 0.0f);
                            if (row_lb + gl * gl * 0.99999f <= B) break;
                        }
#pragma unroll 1
                        for (; xb > lx; --xb) {
                            const float gr = fmaxf((cell - fx) + (float)(xb - lx - 1) * cell - eps, 0.0f);
                            if (row_lb + gr * gr * 0.99999f <= B) break;
                        }
                        xa = max(xa, 0); xb = min(xb, g.nx - 1);
                        if (xa <= xb) {
                            const int* rowp = g.cell_start + (size_t)(zz * g.ny + yy) * g.nx;
                            s = __ldg(rowp + xa); e = __ldg(rowp + xb + 1);
                        }
                    }
                }
            }
            const unsigned bits = (__ballot_sync(full, e > s) >> gshift) & 0xffu;
            if (e > s) { const int at = nr + __popc(bits & below); S.rs[at] = s; S.re[at] = e; }
            nr += __popc(bits);
        }
    }
    __syncwarp();
    bool todo = active && !skip;                                   // my group still needs a (further) pass
    bool slow = false;
    float lb6 = 0.0f;
#pragma unroll 1
    while (__any_sync(full, todo)) {
        int cnt = 0;                                               // survivors of my group (group-uniform)
        int nrm = todo ? nr : 0;
        nrm = max(nrm, __shfl_xor_sync(full, nrm, 8));
        nrm = max(nrm, __shfl_xor_sync(full, nrm, 16));
#pragma unroll 1
        for (int r = 0; r < nrm; ++r) {
            int j = 0, e = 0;
            if (todo && r < nr) { j = S.rs[r] + sub; e = S.re[r]; }
            while (__any_sync(full, j < e)) {
                bool hit = false;
                float d = 0.0f;
                int pi = 0;
                if (j < e) {
                    const float4 p = __ldg(&g.pts[j]);
                    d = dist2(qx, qy, qz, p);
                    pi = __float_as_int(p.w);
                    hit = d <= B;
                }
                const unsigned bits = (__ballot_sync(full, hit) >> gshift) & 0xffu;
                if (hit) {
                    const int slot = cnt + __popc(bits & below);
                    if (slot < kKnnCap) { S.d2[slot] = d; S.pos[slot] = j; S.idx[slot] = pi; }
                }
                cnt += __popc(bits);
                j += kKnnLanes;
            }
        }
        __syncwarp();
        // ---- ranks of the buffered entries
        const int m = todo ? min(cnt, kKnnCap) : 0;
        int mm = m;
        mm = max(mm, __shfl_xor_sync(full, mm, 8));
        mm = max(mm, __shfl_xor_sync(full, mm, 16));
#pragma unroll 1
        for (int e0 = 0; e0 < mm; e0 += kKnnLanes) {
            const int e = e0 + sub;
            const bool mine = e < m;
            const float de = mine ? S.d2[e] : 0.0f;
            const int ie = mine ? S.idx[e] : 0;
            int rank = 0;
#pragma unroll 1
            for (int f = 0; f < mm; ++f)
                if (f < m) rank += knn_less(S.d2[f], S.idx[f], de, ie) ? 1 : 0;
            if (mine && rank < 5) S.out[rank] = S.pos[e];
            if (mine && rank == 4) S.out[5] = __float_as_int(de);
            if (mine && rank == 5) S.out[6] = __float_as_int(de);
        }
        __syncwarp();
        if (todo) {
            if (cnt <= kKnnCap) {
                todo = false;
                // everything outside the five is at least this far (squared): the 6th candidate, else the bound
                lb6 = cnt >= 6 ? __int_as_float(S.out[6]) : B;
            } else {
                const float nb = __int_as_float(S.out[5]) * kKnnLook;   // 5th smallest of 32 real points bounds the answer
                if (nb < B) B = nb;
                else { slow = true; todo = false; }
            }
        }
        __syncwarp();
    }
    if (slow && sub == 0) {                                        // 32+ candidates tied at the bound: sequential exact search
        Knn5 k;
        knn_init(k);
        knn_search(g, qx, qy, qz, k);
#pragma unroll
        for (int i = 0; i < 5; ++i) S.out[i] = k.pos[i];
        S.out[5] = __float_as_int(k.d2[4]);
        lb6 = 0.0f;                                                // no gap known: the next iteration scans again
    }
    if (active && sub == 0) {
        const int4 o0 = make_int4(S.out[0], S.out[1], S.out[2], S.out[3]);
        const bool same = a.use_seeds && o0.x == s0.x && o0.y == s0.y && o0.z == s0.z && o0.w == s0.w && S.out[4] == s1.x;
        a.nn[kKnnRec * qi] = o0;
        if (skip) {                                                // q_scan and lb6 stay those of the last real scan
            a.nn[kKnnRec * qi + 1] = make_int4(S.out[4], S.out[5], s1.z, same ? 1 : 0);
        } else {
            a.nn[kKnnRec * qi + 1] = make_int4(S.out[4], S.out[5], __float_as_int(lb6), same ? 1 : 0);
            a.nn[kKnnRec * qi + 2] = make_int4(__float_as_int(qx), __float_as_int(qy), __float_as_int(qz), 0);
        }
    }
}

// ---- exact 1-NN (post-run point-to-point metrics, DCReg/include/utils.hpp:538-589) ---------------------------
// Nearest target point of q in a dense grid: rings of cells of growing Chebyshev radius around q's cell, clipped to
// the grid box, until the best distance found is no larger than the distance to the next ring.  FLANN-style float32
// squared distances.  Returns the squared distance (3e38 if the grid is empty).
__device__ __forceinline__ float nn1_search(const Grid& g, float qx, float qy, float qz) {
    const float cell = (float)(1.0 / g.inv_cell);
    const int lx = cell_coord(qx, g.inv_cell) - g.ox, ly = cell_coord(qy, g.inv_cell) - g.oy, lz = cell_coord(qz, g.inv_cell) - g.oz;
    // first ring that can touch the box, last ring that still does
    const int ox = lx < 0 ? -lx : (lx >= g.nx ? lx - g.nx + 1 : 0);
    const int oy = ly < 0 ? -ly : (ly >= g.ny ? ly - g.ny + 1 : 0);
    const int oz = lz < 0 ? -lz : (lz >= g.nz ? lz - g.nz + 1 : 0);
    const int r0 = max(ox, max(oy, oz));
    const int r1 = max(max(lx, g.nx - 1 - lx), max(max(ly, g.ny - 1 - ly), max(lz, g.nz - 1 - lz)));
    float best = 3.0e38f;
#pragma unroll 1
    for (int r = r0; r <= r1; ++r) {
        if (r > 0) {
            const float lb = (float)(r - 1) * cell * 0.99999f;      // every point of ring r is at least this far
            if (lb * lb > best) break;
        }
        const int z0 = max(-r, -lz), z1 = min(r, g.nz - 1 - lz);
        const int y0 = max(-r, -ly), y1 = min(r, g.ny - 1 - ly);
#pragma unroll 1
        for (int dz = z0; dz <= z1; ++dz) {
#pragma unroll 1
            for (int dy = y0; dy <= y1; ++dy) {
                const int* rowp = g.cell_start + (size_t)((lz + dz) * g.ny + (ly + dy)) * g.nx;
                const bool shell = (dz == -r) || (dz == r) || (dy == -r) || (dy == r);
                if (shell) {                                         // whole x-span of the ring, one contiguous range
                    const int xa = max(lx - r, 0), xb = min(lx + r, g.nx - 1);
                    if (xa > xb) continue;
                    const int s = __ldg(rowp + xa), e = __ldg(rowp + xb + 1);
                    for (int j = s; j < e; ++j) best = fminf(best, dist2(qx, qy, qz, __ldg(&g.pts[j])));
                } else {                                             // interior row: only the two end cells dx = +-r
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        const int xx = lx + sgn * r;
                        if (xx < 0 || xx >= g.nx) continue;
                        const int s = __ldg(rowp + xx), e = __ldg(rowp + xx + 1);
                        for (int j = s; j < e; ++j) best = fminf(best, dist2(qx, qy, qz, __ldg(&g.pts[j])));
                    }
                }
            }
        }
    }
    return best;
}

// per-block partial sums: [0] sum of distances, [1] sum of squared distances below the threshold, [2] count below it.
// When T != nullptr the query is fl32(T p) (pcl::transformPointCloud: FP64 math, float32 store), else p itself.
__global__ void nn1_metrics_kernel(const float4* __restrict__ q, long long n, const double* __restrict__ T, Grid g,
                                   double threshold, double* __restrict__ partials) {
    __shared__ double sh[3][8];
    double sd = 0.0, ssq = 0.0, cnt = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = __ldg(&q[i]);
        float x = p.x, y = p.y, z = p.z;
        if (T) {
            const double px = p.x, py = p.y, pz = p.z;
            x = (float)(T[0] * px + T[1] * py + T[2] * pz + T[3]);
            y = (float)(T[4] * px + T[5] * py + T[6] * pz + T[7]);
            z = (float)(T[8] * px + T[9] * py + T[10] * pz + T[11]);
        }
        const float d2 = nn1_search(g, x, y, z);
        const double dist = sqrt((double)d2);
        sd += dist;
        if (dist < threshold) { ssq += (double)d2; cnt += 1.0; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        sd += __shfl_down_sync(0xffffffffu, sd, off);
        ssq += __shfl_down_sync(0xffffffffu, ssq, off);
        cnt += __shfl_down_sync(0xffffffffu, cnt, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = sd; sh[1][warp] = ssq; sh[2][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[threadIdx.x][w];
        partials[blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// transform + float32 store of a cloud (aligned copy for the backward Chamfer pass)
__global__ void transform_points_kernel(const float4* __restrict__ in, long long n, const double* __restrict__ T,
                                        float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    const double px = p.x, py = p.y, pz = p.z;
    out[i] = make_float4((float)(T[0] * px + T[1] * py + T[2] * pz + T[3]), (float)(T[4] * px + T[5] * py + T[6] * pz + T[7]),
                         (float)(T[8] * px + T[9] * py + T[10] * pz + T[11]), p.w);
}

// Plane through the 5 neighbours: least squares of [nb] x = -1, n = x/|x|, d = 1/|x|, gates
// |x| >= min_norm and max_j (n.nb_j + d)^2 < thickness^2 (icp_test_runner.cpp:1727-1773).
// Returns true and (n, d) when a valid plane exists.
__device__ __forceinline__ bool fit_plane(const Grid& g, const Knn5& k, double min_norm, double thickness,
                                          double& nx, double& ny, double& nz, double& d) {
    double A[15], b[5], x[3];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = __ldg(&g.pts[k.pos[j]]);
        A[j * 3 + 0] = (double)p.x;
        A[j * 3 + 1] = (double)p.y;
        A[j * 3 + 2] = (double)p.z;
        b[j] = -1.0;
    }
    dla::colpiv_qr_solve<5, 3>(A, b, x);
    const double ps = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (!(ps >= min_norm)) return false;                 // :1752 (also rejects NaN)
    nx = x[0] / ps; ny = x[1] / ps; nz = x[2] / ps; d = 1.0 / ps;
    double worst = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 p = __ldg(&g.pts[k.pos[j]]);         // re-read (L1 hit) instead of holding 15 more doubles
        double e = nx * (double)p.x + ny * (double)p.y + nz * (double)p.z + d;
        e *= e;
        worst = fmax(worst, e);
    }
    return worst < thickness * thickness;                // :1772-1773
}

}  // namespace corr
