// peer_allreduce.cuh -- one-shot SUM all-reduce of the vertex gradient over NVLink / NVSwitch peer memory.
//
// The one collective of the path: views are sharded over the GPUs of one box, the mesh is
// replicated, and grad_V float64[V,3] (603 KB at C4) is summed once per step.  For a message this small a ring /
// tree collective is all latency; here, in ONE kernel launch, every rank
//   1 PUSHES its gradient into slot [rank] of an IPC-exported staging area of EVERY rank (posted NVLink stores),
//     and the last block to finish publishes an epoch flag in every peer's memory (one system fence per block),
//   2 waits until all peers' flags for this epoch have landed in its own memory,
//   3 adds the slots of its own staging area in rank order (local reads; every rank gets the same bits,
//     independent of arrival order).
// Staging slots and flags are double-buffered by epoch parity: a rank may run ahead by one all-reduce, and it
// cannot start epoch e+2 (which reuses the slots of epoch e) before every peer has signalled epoch e+1, i.e. has
// finished reading epoch e.  The grid is kept small enough to be co-resident (blocks spin on the flags).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drt {

constexpr int kMaxPeers = 16;
// A peer that never arrives must not hang the GPU, and it must not be summed either: the wait is bounded in WALL-CLOCK
// time (%globaltimer, nanoseconds; PeerView::timeout_ns, default 120 s, DRT_PEER_TIMEOUT_S) and expiry is FATAL -- the
// error word is set and the kernel traps, so the stale staging slots are never added into grad_V and every later
// CUDA call of the process fails loudly (drt_comm_status reports it when the context is still alive).

struct PeerView {
    double* stage[kMaxPeers];    // staging areas of all ranks: [2 parities][world slots][stride] doubles
    unsigned* flags[kMaxPeers];  // flag blocks of all ranks: [2][kMaxPeers] epochs, written by the peers
    unsigned* arrive;            // local: blocks of this launch that finished stage 1
    unsigned* error;             // local: set when a wait timed out
    int64_t stride;              // doubles per slot
    unsigned long long timeout_ns;  // wall-clock bound of the wait for the peers
    int rank, world;
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerView pv, double* __restrict__ data, int64_t n, unsigned epoch)
{
    const int par = epoch & 1u;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t slot = ((int64_t)par * pv.world + pv.rank) * pv.stride;   // my slot in everybody's staging area
    // 1: push my contribution to every rank (myself included)
    for (int64_t i = tid; i < n; i += nth) {
        const double v = data[i];
        for (int r = 0; r < pv.world; ++r) pv.stage[r][slot + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                      // this block's stores (observed through the barrier) before the count
        if (atomicAdd(pv.arrive, 1u) == gridDim.x - 1) {
            *pv.arrive = 0;
            __threadfence_system();
            for (int r = 0; r < pv.world; ++r) st_sys(pv.flags[r] + par * kMaxPeers + pv.rank, epoch);
        }
    }
    // 2: wait for every rank's flag of this epoch in MY memory
    if ((int)threadIdx.x < pv.world) {
        const unsigned* f = pv.flags[pv.rank] + par * kMaxPeers + threadIdx.x;
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_sys(f) != epoch) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > pv.timeout_ns) {
                atomicExch(pv.error, 1u);
                __threadfence_system();
                __trap();  // fatal: never fall through to the sum with a peer's slot missing
            }
            __nanosleep(32);
        }
    }
    __syncthreads();
    // 3: add the slots of my own staging area in rank order
    const double* mine = pv.stage[pv.rank] + (int64_t)par * pv.world * pv.stride;
    for (int64_t i = tid; i < n; i += nth) {
        double s = 0.0;
        for (int r = 0; r < pv.world; ++r) s += __ldcv(mine + r * pv.stride + i);
        data[i] = s;
    }
}

}  // namespace drt
