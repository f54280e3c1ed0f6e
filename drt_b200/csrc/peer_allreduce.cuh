// peer_allreduce.cuh -- one-shot SUM all-reduce of the vertex gradient over NVLink / NVSwitch peer memory.
//
// The one collective of the path: views are sharded over the GPUs of one box, the mesh is
// replicated, and grad_V float64[V,3] (603 KB at C4) is summed once per step.  For a message this small a ring /
// tree collective is all latency (NCCL: 0.05 ms at 2 GPUs, 0.12 ms at 8 for 603 KB); here every rank
//   1 copies its gradient into an IPC-exported staging buffer and publishes an epoch flag in EVERY peer's memory,
//   2 waits until all peers' flags for this epoch have landed in its own memory,
//   3 reads all staging buffers straight over NVLink and adds them in rank order (so every rank gets the same
//     bits, independent of arrival order),
// in ONE kernel launch.  Staging buffers and flags are double-buffered by epoch parity: a rank may run ahead by one
// all-reduce, and it cannot start epoch e+2 (which reuses the buffers of epoch e) before every peer has signalled
// epoch e+1, i.e. has finished reading epoch e.  The grid is kept small enough to be co-resident (blocks spin).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drt {

constexpr int kMaxPeers = 16;
constexpr unsigned kSpinLimit = 1u << 26;  // ~ seconds; a peer that never arrives sets the error flag instead of hanging the GPU

struct PeerView {
    double* buf[kMaxPeers];      // staging buffers of all ranks (buf[r] + parity * stride)
    unsigned* flags[kMaxPeers];  // flag blocks of all ranks: [2][kMaxPeers] epochs, written by the peers
    unsigned* arrive;            // local: blocks of this launch that finished stage 1
    unsigned* error;             // local: set when a wait timed out
    int64_t stride;              // doubles per parity half
    int rank, world;
};

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
    double* mine = pv.buf[pv.rank] + par * pv.stride;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    // 1: stage my contribution, then the LAST block of this launch publishes the epoch to every rank (myself included)
    for (int64_t i = tid; i < n; i += nth) mine[i] = data[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(pv.arrive, 1u) == gridDim.x - 1) {
            *pv.arrive = 0;
            __threadfence_system();
            for (int r = 0; r < pv.world; ++r) st_sys(pv.flags[r] + par * kMaxPeers + pv.rank, epoch);
        }
    }
    // 2: wait for every rank's flag of this epoch in MY memory
    if ((int)threadIdx.x < pv.world) {
        const unsigned* f = pv.flags[pv.rank] + par * kMaxPeers + threadIdx.x;
        unsigned spins = 0;
        while (ld_sys(f) != epoch) {
            if (++spins > kSpinLimit) { atomicExch(pv.error, 1u); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    __threadfence_system();
    // 3: sum the staging buffers in rank order, reading the peers' memory directly
    for (int64_t i = tid; i < n; i += nth) {
        double s = 0.0;
        for (int r = 0; r < pv.world; ++r) s += __ldcv(pv.buf[r] + par * pv.stride + i);
        data[i] = s;
    }
}

}  // namespace drt
