// loop_plan.hpp - how the source slots of a run are cut into blocks of the iteration kernel (host logic, no CUDA).
//
// One block = one tile of `tile` <= block_threads consecutive source slots per pass (the OpenMP loop body of
// icp_test_runner.cpp:1714-1863, one slot per thread).  Every phase of the kernel is a latency chain per tile, so:
//   * a single run keeps all its tiles resident at once (<= 3 blocks per SM; more slots: the blocks loop over tiles);
//   * a single run of a SMALL cloud (fewer 256-slot tiles than SMs) is cut into >= 2 tiles per SM at 32-slot
//     granularity instead of leaving most SMs idle - measured 47.3 -> 36.1 us per iteration on the shipped 7 562-point
//     cloud (tools/tile_sweep.py); no effect once there is a tile per SM;
//   * batched trials (grid y = trial) are throughput-bound: full tiles, at most 64 blocks per trial.
// tests/test_host_la.py::test_loop_tile_plan checks the invariants on the CPU.
#pragma once
#include <algorithm>

namespace loop_plan {

struct Tiles {
    int tile;            // source slots per block and pass, 32 <= tile <= block_threads
    long long grid_x;    // blocks per trial
};

// tile_override: 0 = rule above; otherwise a requested tile size for single runs (measurement switch DCREG_TILE),
// ignored when it is out of range or would not fit the resident blocks.
inline Tiles plan_tiles(long long slots, int trials, int sm_count, int block_threads, int tile_override = 0) {
    Tiles t{block_threads, std::max<long long>(1, (slots + block_threads - 1) / block_threads)};
    if (trials != 1) {
        t.grid_x = std::min<long long>(t.grid_x, 64);
        return t;
    }
    const long long cap = (long long)sm_count * 3;
    if (t.grid_x > cap) { t.grid_x = cap; return t; }
    int tile = block_threads;
    if (t.grid_x < sm_count) tile = (int)std::max<long long>(32, (slots / (2LL * sm_count) + 31) / 32 * 32);
    if (tile_override) tile = tile_override;
    if (tile < 32 || tile > block_threads || (slots + tile - 1) / tile > cap) tile = block_threads;
    t.tile = tile;
    t.grid_x = std::max<long long>(1, (slots + tile - 1) / tile);
    return t;
}

}  // namespace loop_plan
