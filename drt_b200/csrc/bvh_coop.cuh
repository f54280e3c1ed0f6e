// bvh_coop.cuh -- the whole LBVH (re)build of bvh.cuh as ONE cooperative launch.
//
// The 19-launch pipeline (capi.cu: build_tree) is bound by launch / dependency latency, not by work: 50 k triangles are
// 7 MB of traffic, yet the chain costs 0.17 ms -- 2 % of a one-GPU step of the 72-view benchmark, 12 % of an 8-GPU step
// (every rank rebuilds the replicated tree) and a third of a one-view step.  Here the same stage bodies run inside one
// kernel, separated by grid-wide barriers; every stage is a grid-stride loop, so any co-resident grid works.
//   cast/reset | bounds | morton | 5 x (digit histogram | stable scatter) | topology | fit | grid + emit      (15 barriers)
// Results are bit-identical to the multi-launch build (same arithmetic, same stable sort).
// MEASURED on B200 (50 k triangles): 0.208 ms against 0.170 ms for the 19 launches -- ncu shows the kernel waiting at its
// barriers (45 barrier-stall cycles per issue): a cooperative grid.sync costs ~5 us with 148 blocks, and everything that
// crosses a barrier has to be read through L2.  The launches themselves were never the cost; the serial chains inside fit
// (~30 dependent atomic levels) and the scatter prologues are.  So this is an OPTION (DRT_COOP_BUILD=1, parity-tested), not
// the default.
#pragma once
#include <cooperative_groups.h>

#include "bvh.cuh"

namespace drt {

struct BuildArgs {
    const int32_t* F;
    float* V32;          // [nV,3] the handle's float32 vertices (written when V64 is given)
    const double* V64;   // optional: cast source (DiffRender.py:311,379)
    int nV, n;           // vertices, triangles
    uint64_t* keys;      // [2n] double buffer; sorted keys end up in keys + n (five passes)
    unsigned* table;     // [256 * tiles]
    int2* children;
    int* parent;
    float4* blo;
    float4* bhi;
    int* flags;
    unsigned* scene;
    node_quad* nodes;
    uint4* nodes4;       // wide nodes (DRT_BVH4), may be null
    double2* tris;
    int refit;           // 1: keep keys / topology, run fit + emit only
};

// Inside ONE kernel the key buffers, the digit table and the boxes are rewritten by other SMs between barriers, and L1 is
// not coherent: everything that is read after having been written (or read) earlier in the kernel goes through L2 (__ldcg).
__device__ __forceinline__ int delta_cg(const uint64_t* keys, int n, uint64_t ki, int j)
{
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(ki ^ __ldcg(keys + j)));
}

__device__ __forceinline__ void coop_hist(const uint64_t* __restrict__ keys, int n, int shift, int tiles, unsigned* __restrict__ table,
                                          unsigned* h /* shared[256] */)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        h[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            int i = sort_item_index(tile, warp, r, lane);
            if (i < n) atomicAdd(&h[(unsigned)(__ldcg(keys + i) >> shift) & 255u], 1u);
        }
        __syncthreads();
        table[threadIdx.x * tiles + tile] = h[threadIdx.x];
        __syncthreads();
    }
}

struct ScatterSmem {
    unsigned cnt[kSortWarps][256];
    unsigned gbase[256];
    unsigned wsum[kSortWarps];
};

// one tile of sort_scatter_kernel (radix_sort.cuh), same ranks, same order
__device__ __forceinline__ void coop_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int n, int shift, int tiles,
                                             const unsigned* __restrict__ table, ScatterSmem& sm)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int j = threadIdx.x; j < kSortWarps * 256; j += kSortThreads) (&sm.cnt[0][0])[j] = 0;
        {
            const unsigned* row = table + threadIdx.x * tiles;
            unsigned total = 0, before = 0;
            for (int t = 0; t < tiles; ++t) {
                unsigned c = __ldcg(row + t);
                before += t < tile ? c : 0u;
                total += c;
            }
            unsigned x = total;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                unsigned y = __shfl_up_sync(0xffffffffu, x, sft);
                if (lane >= sft) x += y;
            }
            if (lane == 31) sm.wsum[warp] = x;
            __syncthreads();
            unsigned off = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) off += w < warp ? sm.wsum[w] : 0u;
            sm.gbase[threadIdx.x] = off + x - total + before;
        }
        __syncthreads();
        uint64_t key[kSortItems];
        unsigned rank[kSortItems];
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            int i = sort_item_index(tile, warp, r, lane);
            bool ok = i < n;
            key[r] = ok ? __ldcg(in + i) : ~0ull;
            unsigned d = ok ? ((unsigned)(key[r] >> shift) & 255u) : 256u;
            unsigned peers = __match_any_sync(0xffffffffu, d);
            unsigned before = __popc(peers & lt);
            unsigned old = 0;
            if (ok && before == 0) {
                old = sm.cnt[warp][d];
                sm.cnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            rank[r] = old + before;
            __syncwarp();
        }
        __syncthreads();
        {
            unsigned d = threadIdx.x, run = sm.gbase[d];
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                unsigned c = sm.cnt[w][d];
                sm.cnt[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            int i = sort_item_index(tile, warp, r, lane);
            if (i < n) out[sm.cnt[warp][(unsigned)(key[r] >> shift) & 255u] + rank[r]] = key[r];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) lbvh_build_kernel(BuildArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ ScatterSmem sm;
    const int n = a.n;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const float* V = a.V32;

    // ---- cast + reset ---------------------------------------------------------------------------------------------
    if (a.V64)
        for (int i = tid; i < 3 * a.nV; i += nth) a.V32[i] = __double2float_rn(a.V64[i]);
    if (n > 1)
        for (int i = tid; i < n - 1; i += nth) a.flags[i] = 0;
    if (!a.refit && tid < 8 && tid != 6) a.scene[tid] = tid < 3 ? 0xffffffffu : 0u;  // [6] = bad-index count: kept
    grid.sync();

    uint64_t* sorted = a.keys + n;  // five passes: the result lands in the second half
    if (!a.refit) {
        // ---- centroid bounds (centroid_bounds_kernel) ------------------------------------------------------------
        for (int base = blockIdx.x * blockDim.x; base < n; base += nth) {
            const int f = base + threadIdx.x;
            float c[3] = {INFINITY, INFINITY, INFINITY}, C[3] = {-INFINITY, -INFINITY, -INFINITY};
            float amax = 0.f;
            if (f < n) {
                float lo[3], hi[3];
                tri_box(a.F, V, f, lo, hi);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    c[k] = C[k] = 0.5f * lo[k] + 0.5f * hi[k];
                    amax = fmaxf(amax, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
                }
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    c[k] = fminf(c[k], __shfl_xor_sync(0xffffffffu, c[k], s));
                    C[k] = fmaxf(C[k], __shfl_xor_sync(0xffffffffu, C[k], s));
                }
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    atomicMin(&a.scene[k], enc_f32(c[k]));
                    atomicMax(&a.scene[3 + k], enc_f32(C[k]));
                }
                atomicMax(&a.scene[7], __float_as_uint(amax));
            }
        }
        grid.sync();
        // ---- Morton keys (morton_kernel) --------------------------------------------------------------------------
        {
            float mn[3], ext[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                mn[k] = dec_f32(__ldcg(a.scene + k));
                ext[k] = dec_f32(__ldcg(a.scene + 3 + k)) - mn[k];
            }
            for (int f = tid; f < n; f += nth) {
                float lo[3], hi[3];
                tri_box(a.F, V, f, lo, hi);
                uint32_t q[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float c = 0.5f * lo[k] + 0.5f * hi[k];
                    float u = ext[k] > 0.f ? (c - mn[k]) / ext[k] : 0.f;
                    q[k] = (uint32_t)fminf(fmaxf(u * 8192.f, 0.f), 8191.f);
                }
                uint64_t morton = (spread13(q[0]) << 2) | (spread13(q[1]) << 1) | spread13(q[2]);
                a.keys[f] = (morton << kIndexBits) | (uint64_t)(uint32_t)f;
            }
        }
        grid.sync();
        // ---- LSD radix sort over the Morton bits (radix_sort.cuh) -------------------------------------------------
        {
            const int tiles = (n + kSortTile - 1) / kSortTile;
            uint64_t* src = a.keys;
            uint64_t* dst = a.keys + n;
            for (int shift = kIndexBits & ~7; shift < 64; shift += 8) {
                coop_hist(src, n, shift, tiles, a.table, sm.gbase);
                grid.sync();
                coop_scatter(src, dst, n, shift, tiles, a.table, sm);
                grid.sync();
                uint64_t* t = src; src = dst; dst = t;
            }
            sorted = src;
        }
        // ---- topology (topology_kernel) ---------------------------------------------------------------------------
        for (int i = tid; i < n - 1; i += nth) {
            const uint64_t ki = __ldcg(sorted + i);
            int d = (delta_cg(sorted, n, ki, i + 1) - delta_cg(sorted, n, ki, i - 1)) >= 0 ? 1 : -1;
            int dmin = delta_cg(sorted, n, ki, i - d);
            int lmax = 2;
            while (delta_cg(sorted, n, ki, i + lmax * d) > dmin) lmax <<= 1;
            int l = 0;
            for (int t = lmax >> 1; t >= 1; t >>= 1)
                if (delta_cg(sorted, n, ki, i + (l + t) * d) > dmin) l += t;
            int j = i + l * d;
            int dnode = delta_cg(sorted, n, ki, j);
            int s = 0;
            for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
                if (delta_cg(sorted, n, ki, i + (s + t) * d) > dnode) s += t;
                if (t == 1) break;
            }
            int gamma = i + s * d + min(d, 0);
            int left = (min(i, j) == gamma) ? (n - 1 + gamma) : gamma;
            int right = (max(i, j) == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
            a.children[i] = make_int2(left, right);
            a.parent[left] = i;
            a.parent[right] = i;
            if (i == 0) a.parent[0] = -1;
        }
        grid.sync();
    }

    // ---- bottom-up fit (fit_kernel) -----------------------------------------------------------------------------------
    for (int k = tid; k < n; k += nth) {
        float lo[3], hi[3];
        tri_box(a.F, V, key_tri(__ldcg(sorted + k)), lo, hi);
        const int me = n - 1 + k;
        a.blo[me] = make_float4(lo[0], lo[1], lo[2], 0.f);
        a.bhi[me] = make_float4(hi[0], hi[1], hi[2], 0.f);
        if (n == 1) break;
        int cur = a.parent[me];
        while (cur >= 0) {
            cuda::atomic_ref<int, cuda::thread_scope_device> arrived(a.flags[cur]);
            if (arrived.fetch_add(1, cuda::memory_order_acq_rel) == 0) break;
            int2 ch = a.children[cur];
            float4 l0 = __ldcg(&a.blo[ch.x]), h0 = __ldcg(&a.bhi[ch.x]);
            float4 l1 = __ldcg(&a.blo[ch.y]), h1 = __ldcg(&a.bhi[ch.y]);
            a.blo[cur] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
            a.bhi[cur] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
            cur = a.parent[cur];
        }
    }
    grid.sync();

    // ---- quantisation grid (grid_kernel: every thread derives the same values from the root box) + emit --------------
#if DRT_QNODE
    float g0[3], st[3];
    {
        const float4 l = __ldcg(&a.blo[0]), h = __ldcg(&a.bhi[0]);
        float ext[3] = {__fadd_ru(h.x, -l.x), __fadd_ru(h.y, -l.y), __fadd_ru(h.z, -l.z)};
        const float lo[3] = {l.x, l.y, l.z};
        float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
        if (!(emax > 7.888609052210118e-31f)) emax = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float e = fmaxf(ext[k], emax * 9.5367431640625e-07f);
            st[k] = __fdiv_ru(e, kGridSteps);
            g0[k] = __fadd_rd(lo[k], -__fmul_ru(7.f, st[k]));
        }
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a.scene[8 + k] = __float_as_uint(g0[k]);
                a.scene[11 + k] = __float_as_uint(st[k]);
            }
        }
    }
    const float ix = __fdiv_rn(1.f, st[0]), iy = __fdiv_rn(1.f, st[1]), iz = __fdiv_rn(1.f, st[2]);
    if (n == 1) {
        if (tid == 0) {
            float4 l = __ldcg(&a.blo[0]), h = __ldcg(&a.bhi[0]);
            const unsigned x = qpair(l.x, h.x, g0[0], ix), y = qpair(l.y, h.y, g0[1], iy), z = qpair(l.z, h.z, g0[2], iz);
            a.nodes[0] = make_uint4(x, x, y, y);
            a.nodes[1] = make_uint4(z, z, (unsigned)~0, (unsigned)~0);
        }
    } else {
        for (int i = tid; i < n - 1; i += nth) {
            int2 ch = a.children[i];
            float4 l0 = __ldcg(&a.blo[ch.x]), h0 = __ldcg(&a.bhi[ch.x]), l1 = __ldcg(&a.blo[ch.y]), h1 = __ldcg(&a.bhi[ch.y]);
            int c0 = ch.x >= n - 1 ? ~(ch.x - (n - 1)) : ch.x;
            int c1 = ch.y >= n - 1 ? ~(ch.y - (n - 1)) : ch.y;
            uint4* node = a.nodes + (size_t)i * kNodeQuads;
            node[0] = make_uint4(qpair(l0.x, h0.x, g0[0], ix), qpair(l1.x, h1.x, g0[0], ix), qpair(l0.y, h0.y, g0[1], iy), qpair(l1.y, h1.y, g0[1], iy));
            node[1] = make_uint4(qpair(l0.z, h0.z, g0[2], iz), qpair(l1.z, h1.z, g0[2], iz), (unsigned)c0, (unsigned)c1);
        }
    }
    if (a.nodes4) {
        const float inv_s[3] = {ix, iy, iz};
        for (int i = tid; i < (n > 1 ? n - 1 : 1); i += nth) emit_wide_node(i, n, a.children, a.blo, a.bhi, g0, inv_s, a.nodes4);
    }
#endif
    for (int k = tid; k < n; k += nth) {
        int f = key_tri(__ldcg(sorted + k));
        const float* pa = &V[3 * (size_t)a.F[3 * f]];
        const float* pb = &V[3 * (size_t)a.F[3 * f + 1]];
        const float* pc = &V[3 * (size_t)a.F[3 * f + 2]];
        d3 va = mk3((double)pa[0], (double)pa[1], (double)pa[2]);
        d3 e1 = mk3((double)pb[0], (double)pb[1], (double)pb[2]) - va;
        d3 e2 = mk3((double)pc[0], (double)pc[1], (double)pc[2]) - va;
        double2* t = a.tris + (size_t)k * kTriD2;
        t[0] = make_double2(va.x, va.y);
        t[1] = make_double2(va.z, e1.x);
        t[2] = make_double2(e1.y, e1.z);
        t[3] = make_double2(e2.x, e2.y);
        t[4] = make_double2(e2.z, __longlong_as_double((long long)f));
    }
}

}  // namespace drt
