// peer_reduce.cuh - the multi-GPU exchange step of the path, inside the reducing kernel.
//
// the path shards by contiguous source-point blocks and has exactly ONE exchange per ICP iteration -
// the sum over ranks of the 27 + 5 accumulators (the OpenMP `reduction(+: ...)` of icp_test_runner.cpp:1714 and
// SymmetricHessianComputer::join, hessian_computer.h:103-108).  240 B per rank: pure latency.  A separate
// ncclAllReduce kernel behind the reduction costs ~25 us per iteration (round 1: 65.9 -> 40.5 us for 10 M slots on
// 8 GPUs, i.e. 1.6x), so the exchange lives in the LAST BLOCK of the reducing kernel instead:
//
//   every rank owns a 4.3 KB mailbox in its own HBM, mapped into every peer's address space (cudaIpc handles
//   exchanged once in dcreg_comm_init; NVLink 5 / NVSwitch P2P stores);
//   epoch e (same on all ranks: every rank runs the same sequence of reductions):
//     post : warp q of the last block stores this rank's 32 packed totals into rank q's mailbox slot
//            data[e & 1][my_rank][0..31], fences (system scope) and releases flag[my_rank] = e there;
//     wait : warp q spins (acquire, system scope) on ITS OWN mailbox's flag[q] until it reaches e, then reads
//            data[e & 1][q][lane];
//     sum  : in rank order 0..N-1 - the same order on every rank, so all ranks hold bit-identical sums and the
//            solve that follows (K2, redundantly on every rank) yields bit-identical poses: no broadcast needed.
//   Two data slots (epoch parity) suffice: a rank can only post epoch e + 1 after it has received every peer's
//   epoch e, and a peer posts epoch e only after it finished reading epoch e - 1.
// No rank waits before it has posted, so the exchange cannot deadlock; a peer that never posts (crashed process)
// trips a 4 s globaltimer timeout, which raises `error` in the mailbox instead of hanging the GPU.
// NCCL (dcreg_b200.cu: nccl_allreduce_acc) remains as the fallback when peer mapping is unavailable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace peer {

constexpr int kMaxRanks = 8;     // one warp of the 256-thread last block per rank
constexpr int kVals = 32;        // k1s::kPk packed totals

struct Mailbox {
    double data[2][kMaxRanks][kVals];
    unsigned int flag[kMaxRanks];        // written by peer r: epoch of its latest complete contribution
    unsigned int epoch;                  // this rank's own epoch counter (local)
    unsigned int error;                  // != 0: a wait timed out
    unsigned int pad[6];
};

struct View {
    int nranks;                          // <= 1: no exchange
    int rank;
    Mailbox* box[kMaxRanks];             // box[r]: rank r's mailbox in THIS process's address space (box[rank] is local)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// All 256 threads of the (single) last block call this.  fin[32] (shared memory): in = this rank's packed totals,
// out = the sum over ranks in rank order.  red: shared scratch [kMaxRanks][32].
__device__ __forceinline__ void all_reduce32(const View& pv, double* fin, double (*red)[kVals]) {
    if (pv.nranks <= 1) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = (int)(blockDim.x >> 5);
    Mailbox* mine = pv.box[pv.rank];
    const unsigned int e = mine->epoch + 1u;            // read by everyone before thread 0 advances it below
    const double v = fin[lane];
    for (int q = warp; q < pv.nranks; q += nwarps) {    // post
        if (q == pv.rank) { red[q][lane] = v; continue; }
        Mailbox* dst = pv.box[q];
        st_relaxed_sys(&dst->data[e & 1u][pv.rank][lane], v);
        __threadfence_system();
        __syncwarp();
        if (lane == 0) st_release_sys(&dst->flag[pv.rank], e);
    }
    for (int q = warp; q < pv.nranks; q += nwarps) {    // wait
        if (q == pv.rank) continue;
        if (lane == 0) {
            unsigned long long t0 = 0;
            while ((int)(ld_acquire_sys(&mine->flag[q]) - e) < 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) { mine->error = 1u + (unsigned)q; break; }
            }
        }
        __syncwarp();
        red[q][lane] = ld_relaxed_sys(&mine->data[e & 1u][q][lane]);
    }
    __syncthreads();
    if (tid < kVals) {
        double s = 0.0;
        for (int r = 0; r < pv.nranks; ++r) s += red[r][tid];
        fin[tid] = s;
    }
    if (tid == 0) mine->epoch = e;
    __syncthreads();
}

}  // namespace peer
