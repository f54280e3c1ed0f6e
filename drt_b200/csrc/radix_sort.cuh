// radix_sort.cuh -- LSD radix sort of unique 64-bit keys (sm_100a), written for the LBVH build.
//
// Keys are (39-bit Morton code << 25) | triangle index, so they are unique, carry their own payload
// and arrive sorted by their low 25 bits; only the passes that touch Morton bits are run (bits 24..63,
// five 8-bit passes).  Each pass is two launches:
//   hist    : per-tile digit histograms, digit-major [256][tiles]
//   scatter : every block first derives its global bases from the raw table (thread = digit: total of
//             the digit over all tiles, block-wide exclusive scan over digits, plus the digit's counts
//             in the tiles before this one -- <= 256 x tiles reads, L2 resident), then computes STABLE
//             ranks from warp match_any + per-warp digit counters in shared memory and scatters
// A tile is 256 threads x 8 keys; stability needs the (warp, round, lane) order to equal the key order,
// hence the warp-blocked item assignment below.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

namespace drt {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per block
constexpr int kSortWarps = kSortThreads / 32;

__device__ __forceinline__ int sort_item_index(int tile, int warp, int round, int lane)
{
    return tile * kSortTile + warp * (32 * kSortItems) + round * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint64_t* __restrict__ keys, int n, int shift,
                                                                 int tiles, unsigned* __restrict__ table)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int i = sort_item_index(blockIdx.x, warp, r, lane);
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                                    int n, int shift, int tiles,
                                                                    const unsigned* __restrict__ table)
{
    __shared__ unsigned cnt[kSortWarps][256];  // per-warp digit counts, then exclusive bases across warps
    __shared__ unsigned gbase[256];            // global base of every digit for THIS tile
    __shared__ unsigned wsum[kSortWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int j = threadIdx.x; j < kSortWarps * 256; j += kSortThreads) (&cnt[0][0])[j] = 0;
    {   // thread = digit d: total over tiles and the part that belongs to earlier tiles
        const unsigned* row = table + threadIdx.x * tiles;
        unsigned total = 0, before = 0;
        for (int t = 0; t < tiles; ++t) {
            unsigned c = row[t];
            before += t < (int)blockIdx.x ? c : 0u;
            total += c;
        }
        unsigned x = total;  // block-wide exclusive scan of `total` over the 256 digits
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, sft);
            if (lane >= sft) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        unsigned off = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) off += w < warp ? wsum[w] : 0u;
        gbase[threadIdx.x] = off + x - total + before;
    }
    __syncthreads();
    uint64_t key[kSortItems];
    unsigned rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int i = sort_item_index(blockIdx.x, warp, r, lane);
        bool ok = i < n;
        key[r] = ok ? in[i] : ~0ull;
        unsigned d = ok ? ((unsigned)(key[r] >> shift) & 255u) : 256u;  // 256 = padding lanes, grouped apart
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned before = __popc(peers & lt);
        unsigned old = 0;
        if (ok && before == 0) {  // lowest lane of the group owns the counter update (warp-private row: no atomics)
            old = cnt[warp][d];
            cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    {  // exclusive prefix over warps for each digit (thread = digit), plus the tile's global base
        unsigned d = threadIdx.x, run = gbase[d];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            unsigned c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        int i = sort_item_index(blockIdx.x, warp, r, lane);
        if (i < n) out[cnt[warp][(unsigned)(key[r] >> shift) & 255u] + rank[r]] = key[r];
    }
}

inline int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }

// Sorts n keys on bits [first_bit, 64); buf holds 2n keys (keys in the first half on entry), table holds
// 256 * sort_tiles(n) counters.  Returns the half that holds the sorted keys.  5 passes x 2 launches.
inline uint64_t* sort_keys_u64(uint64_t* buf, int n, int first_bit, unsigned* table, cudaStream_t st, std::atomic<unsigned long long>* launches)
{
    uint64_t* a = buf;
    uint64_t* b = buf + n;
    const int tiles = sort_tiles(n);
    for (int shift = first_bit & ~7; shift < 64; shift += 8) {
        sort_hist_kernel<<<tiles, kSortThreads, 0, st>>>(a, n, shift, tiles, table);
        sort_scatter_kernel<<<tiles, kSortThreads, 0, st>>>(a, b, n, shift, tiles, table);
        if (launches) *launches += 2;
        uint64_t* t = a; a = b; b = t;
    }
    return a;
}

}  // namespace drt
