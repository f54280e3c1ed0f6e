// radix_sort.cuh -- LSD radix sort of unique 64-bit keys (sm_100a), written for the LBVH build.
//
// Keys are (39-bit Morton code << 25) | triangle index, so they are unique, carry their own payload
// and arrive sorted by their low 25 bits; only the passes that touch Morton bits are run (bits 24..63,
// five 8-bit passes).  Round 2: ONE launch per pass.  The per-tile digit histogram of a pass, digit-major [256][tiles], is
// accumulated by whoever WRITES the keys that pass will read -- the Morton kernel for the first pass, the scatter of pass p
// for pass p+1 (a key's destination tile is known when it is scattered) -- with one global atomicAdd per key into a table
// that one memset cleared for all passes; integer counts, so the sort is as deterministic as before.  (The two-launch form,
// a histogram kernel in front of every scatter, cost 5 more launches = 35 of the 166 us of a 50 k-triangle rebuild.)
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

// table_next != nullptr: also count, per destination tile, the digit at next_shift of every key scattered (the next pass's table)
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                                    int n, int shift, int tiles,
                                                                    const unsigned* __restrict__ table, unsigned* __restrict__ table_next,
                                                                    int next_shift)
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
        if (i < n) {
            const unsigned pos = cnt[warp][(unsigned)(key[r] >> shift) & 255u] + rank[r];
            out[pos] = key[r];
            if (table_next) atomicAdd(&table_next[((unsigned)(key[r] >> next_shift) & 255u) * tiles + pos / kSortTile], 1u);
        }
    }
}

inline int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }
inline int sort_passes(int first_bit) { return (64 - (first_bit & ~7)) / 8; }
// counters the sort needs: one [256][tiles] table per pass
inline size_t sort_table_words(int n, int first_bit) { return (size_t)sort_passes(first_bit) * 256 * (size_t)sort_tiles(n); }

// Sorts n keys on bits [first_bit, 64); buf holds 2n keys (keys in the first half on entry), table holds sort_table_words()
// counters.  Returns the half that holds the sorted keys.  first_table_ready: the caller has cleared ALL tables and accumulated
// the first pass's histogram itself (the Morton kernel does); otherwise a histogram kernel does it here.  One launch per pass.
inline uint64_t* sort_keys_u64(uint64_t* buf, int n, int first_bit, unsigned* table, bool first_table_ready, cudaStream_t st,
                               std::atomic<unsigned long long>* launches)
{
    uint64_t* a = buf;
    uint64_t* b = buf + n;
    const int tiles = sort_tiles(n);
    const size_t words = 256 * (size_t)tiles;
    const int shift0 = first_bit & ~7;
    if (!first_table_ready) {
        cudaMemsetAsync(table, 0, sort_table_words(n, first_bit) * sizeof(unsigned), st);
        sort_hist_kernel<<<tiles, kSortThreads, 0, st>>>(a, n, shift0, tiles, table);
        if (launches) *launches += 1;
    }
    int pass = 0;
    for (int shift = shift0; shift < 64; shift += 8, ++pass) {
        const bool last = shift + 8 >= 64;
        sort_scatter_kernel<<<tiles, kSortThreads, 0, st>>>(a, b, n, shift, tiles, table + pass * words, last ? nullptr : table + (pass + 1) * words,
                                                            shift + 8);
        if (launches) *launches += 1;
        uint64_t* t = a; a = b; b = t;
    }
    return a;
}

}  // namespace drt
