// bvh.cuh -- device LBVH: data layout, build and refit (sm_100a).
//
// Replaces the closed OptiX Prime builder behind optix_mesh::update (reference
// optix_extend.cpp:61-67: setTriangles + update(RTP_MODEL_HINT_ASYNC)).
//
// Pipeline (all on the caller's stream, no host sync):
//   1 bounds   : per-triangle AABB centroid -> scene centroid bounds + largest |coordinate|
//                (warp shuffle + ordered-uint atomics)
//   2 morton   : key = (39-bit Morton code of the normalised centroid, 13 bits/axis) << 25 | triangle id
//   3 sort     : hand-written LSD radix sort of the unique 64-bit keys (radix_sort.cuh)
//   4 topology : Karras 2012 -- one thread per internal node finds its key range and split
//   5 fit      : bottom-up AABB union with atomic arrival counters
//   6 emit     : traversal layout (below)
// Refit = steps 5 and 6 only (topology kept).
// (A 4-wide collapse of this tree was built and measured on B200: +40 % executed instructions and
//  35 % slower than the binary layout on the benchmark meshes -- see DESIGN.md -- so it is not kept.)
#pragma once
#include <cuda/atomic>

#include "common.cuh"
#include "radix_sort.cuh"

namespace drt {

// ---- traversal layout ---------------------------------------------------------------------------
// Binary node, 64 B = 4 x float4, holding BOTH children's boxes plane-major so that one FFMA per
// plane evaluates a slab:
//   n0 = (c0.lo.x, c0.hi.x, c1.lo.x, c1.hi.x)   n1 = same for y   n2 = same for z
//   n3 = (link0, link1, -, -) int bits: >= 0 node index, < 0: ~slot of a triangle record (1 per leaf)
// Every box is inflated by pmax * 2^-17 (pmax = largest |coordinate| of the mesh), which absorbs the
// rounding of the FMA-form slab test for ray origins up to 64 * pmax away (trace.cuh: node_step).
// Triangle record, 80 B = 5 x double2, float64 so that the exact hit test needs no conversions:
//   a.x a.y | a.z e1.x | e1.y e1.z | e2.x e2.y | e2.z id   (a = vertex 0, e1 = v1-v0, e2 = v2-v0, all
//   computed in float64 from the float32 vertices, i.e. exactly what the query-stage test uses)
//
// Default layout (DRT_QNODE = 1): the same node QUANTISED to 32 B = 2 x uint4.  The traversal kernels are bound
// by L1 data-pipe wavefronts of the divergent node fetches (ncu: l1tex__data_pipe_lsu_wavefronts 58-73 % of
// peak, every other unit lower), and a 64-byte node costs four LDG.128 per lane per step; 32 bytes cost two
// and halve the BVH footprint that L1/L2 have to hold.  Planes are 16-bit positions on a per-axis grid
// spanning the root box, g0 + q * s with s = extent / 65520, stored OUTWARD (floor - 3 / ceil + 3 steps):
//   q0 = (c0.x.lo | c0.x.hi << 16, c1.x.lo | c1.x.hi << 16, c0.y.., c1.y..)   q1 = (c0.z.., c1.z.., link0, link1)
// The slab test stays one FFMA per plane, t = q * (s/d) + (g0 - o)/d (trace.cuh: node_step); the three
// steps of margin cover the rounding of the quantisation (< 0.02 step) and of that FMA form (< 0.51 step for
// origins within 64 extents of the grid, farther origins carry an additive slack), so a box test never
// rejects a box that contains a true hit and hit ids stay bit-identical to the brute-force oracle.
// scene[8..10] = g0, scene[11..13] = s (float bits), written by grid_kernel after the fit.
#ifndef DRT_QNODE
#define DRT_QNODE 1
#endif
#if DRT_QNODE
constexpr int kNodeQuads = 2;
typedef uint4 node_quad;
#else
constexpr int kNodeQuads = 4;
typedef float4 node_quad;
#endif
constexpr int kTriD2 = 5;
constexpr int kSceneWords = 16;
constexpr int kEmptyLink = 0x7fffffff;  // link of an empty wide-node slot (its box can never be entered)
constexpr float kGridSteps = 65520.f;  // steps across the root box; 7 spare steps on the low side, 8 on the high side

// DRT_BVH4: a 4-WIDE view of the same tree for the persistent query kernels.  Wide node i (64 B = 4 x uint4) holds the
// GRANDCHILDREN of binary node i: slots 0,1 = the children of its left child, slots 2,3 = those of its right child (a child that
// is a leaf fills one slot of its pair, the other slot is an empty box that no ray can enter), planes quantised exactly like
// the binary layout:   X = (x of slot 0..3 as lo | hi << 16)   Y   Z   L = (link of slot 0..3).
// One wide step = one dependent fetch where the binary walk needs two; the query kernels are bound by the latency of those
// fetches (ncu, r02a: long-scoreboard stalls, 59 % issue slots busy, 17 of 32 lanes), not by their instruction count.  The
// push order (nearer pair first, nearer member first) is the order the binary walk takes, so the visit order is unchanged.
// Links are binary node numbers, so an entry point found by the beam pass on the binary nodes is valid here too.
#ifndef DRT_BVH4
#define DRT_BVH4 0  // measured on B200 at C4: forward 5.14 ms wide vs 5.18 ms binary (the wide step executes as many instructions as the two binary steps it replaces) -- kept as a tested option
#endif

struct BvhView {
    const node_quad* nodes;
    const uint4* nodes4;   // [nNodes][4] wide nodes (DRT_BVH4), else nullptr
    const double2* tris;
    const int32_t* F;      // [nF,3] original faces
    const unsigned* scene; // scene[7] = float bits of pmax
    int nTris;
};

__device__ __forceinline__ unsigned enc_f32(float f)
{
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e)
{
    unsigned b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(b);
}

__device__ __forceinline__ void tri_box(const int32_t* __restrict__ F, const float* __restrict__ V, int f, float lo[3],
                                        float hi[3])
{
    int i0 = F[3 * f], i1 = F[3 * f + 1], i2 = F[3 * f + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = V[3 * (size_t)i0 + k], b = V[3 * (size_t)i1 + k], c = V[3 * (size_t)i2 + k];
        lo[k] = fminf(a, fminf(b, c));
        hi[k] = fmaxf(a, fmaxf(b, c));
    }
}

// V64 -> V32, the cast of DiffRender.py:311,379 (round to nearest even)
__global__ void cast_vertices_kernel(const double* __restrict__ V64, float* __restrict__ V32, int n3)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) V32[i] = __double2float_rn(V64[i]);
}

// scene[0..2] = enc(min centroid), scene[3..5] = enc(max centroid), scene[7] = bits of max |coordinate|;
// caller presets to 0xffffffff / 0 / 0
__global__ void centroid_bounds_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, int nF,
                                       unsigned* __restrict__ scene)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {INFINITY, INFINITY, INFINITY}, C[3] = {-INFINITY, -INFINITY, -INFINITY};
    float amax = 0.f;
    if (f < nF) {
        float lo[3], hi[3];
        tri_box(F, V, f, lo, hi);
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
            atomicMin(&scene[k], enc_f32(c[k]));
            atomicMax(&scene[3 + k], enc_f32(C[k]));
        }
        atomicMax(&scene[7], __float_as_uint(amax));  // non-negative floats order like their bit patterns
    }
}

constexpr int kIndexBits = 25;              // up to 33 554 432 triangles
constexpr int kMortonBitsPerAxis = 13;      // 8192 cells per axis

__device__ __forceinline__ uint64_t spread13(uint32_t v)
{
    uint64_t x = v & 0x1fffu;
    x = (x | x << 16) & 0x0000ff0000ffull;   // keep generous masks: only 13 input bits are live
    x = (x | x << 8) & 0x00f00f00f00full;
    x = (x | x << 4) & 0x0c30c30c30c3ull;
    x = (x | x << 2) & 0x249249249249ull;
    return x;
}

// hist != nullptr: also accumulate the first sort pass's per-tile digit histogram (radix_sort.cuh), table cleared by the caller
__global__ void morton_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, int nF,
                              const unsigned* __restrict__ scene, uint64_t* __restrict__ keys, unsigned* __restrict__ hist, int hist_shift,
                              int hist_tiles)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    float lo[3], hi[3];
    tri_box(F, V, f, lo, hi);
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float mn = dec_f32(scene[k]), mx = dec_f32(scene[3 + k]);
        float c = 0.5f * lo[k] + 0.5f * hi[k];
        float ext = mx - mn;
        float u = ext > 0.f ? (c - mn) / ext : 0.f;
        float s = fminf(fmaxf(u * 8192.f, 0.f), 8191.f);
        q[k] = (uint32_t)s;
    }
    uint64_t morton = (spread13(q[0]) << 2) | (spread13(q[1]) << 1) | spread13(q[2]);
    const uint64_t key = (morton << kIndexBits) | (uint64_t)(uint32_t)f;
    keys[f] = key;
    if (hist) atomicAdd(&hist[((unsigned)(key >> hist_shift) & 255u) * hist_tiles + f / kSortTile], 1u);
}

__device__ __forceinline__ int key_tri(uint64_t key) { return (int)(key & ((1ull << kIndexBits) - 1)); }

// Karras delta: length of the common prefix of the (unique) keys i and j, -1 out of range
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, uint64_t ki, int j)
{
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(ki ^ keys[j]));
}

// node numbering during the build: internal i -> i (0..n-2), leaf k -> n-1+k
__global__ void topology_kernel(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                                int* __restrict__ parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint64_t ki = keys[i];
    int d = (delta(keys, n, i, ki, i + 1) - delta(keys, n, i, ki, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, ki, i - d);
    int lmax = 2;
    while (delta(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, ki, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int left = (min(i, j) == gamma) ? (n - 1 + gamma) : gamma;
    int right = (max(i, j) == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// bottom-up fit; box arrays indexed by build numbering; flags[n-1] zeroed by the caller
__global__ void fit_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, const uint64_t* __restrict__ keys,
                           int n, const int2* __restrict__ children, const int* __restrict__ parent,
                           float4* __restrict__ blo, float4* __restrict__ bhi, int* __restrict__ flags)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float lo[3], hi[3];
    tri_box(F, V, key_tri(keys[k]), lo, hi);
    int me = n - 1 + k;
    blo[me] = make_float4(lo[0], lo[1], lo[2], 0.f);
    bhi[me] = make_float4(hi[0], hi[1], hi[2], 0.f);
    if (n == 1) return;
    int cur = parent[me];
    while (cur >= 0) {
        // one acq_rel RMW at device scope: releases this thread's box store, acquires the sibling's
        cuda::atomic_ref<int, cuda::thread_scope_device> arrived(flags[cur]);
        if (arrived.fetch_add(1, cuda::memory_order_acq_rel) == 0) return;  // first arrival: the sibling finishes this node
        int2 ch = children[cur];
        // volatile-style reads through L2: the sibling's stores were fenced before its atomic
        float4 l0 = __ldcg(&blo[ch.x]), h0 = __ldcg(&bhi[ch.x]);
        float4 l1 = __ldcg(&blo[ch.y]), h1 = __ldcg(&bhi[ch.y]);
        blo[cur] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        bhi[cur] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        cur = parent[cur];
    }
}

__device__ __forceinline__ float4 planes(float lo0, float hi0, float lo1, float hi1, float infl)
{
    return make_float4(__fadd_rd(lo0, -infl), __fadd_ru(hi0, infl), __fadd_rd(lo1, -infl), __fadd_ru(hi1, infl));
}

// quantisation grid of the node planes from the root box (blo[0]/bhi[0]: node 0 is the root, or the only leaf)
__global__ void grid_kernel(const float4* __restrict__ blo, const float4* __restrict__ bhi, unsigned* __restrict__ scene)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float4 l = blo[0], h = bhi[0];
    float ext[3] = {__fadd_ru(h.x, -l.x), __fadd_ru(h.y, -l.y), __fadd_ru(h.z, -l.z)};
    const float lo[3] = {l.x, l.y, l.z};
    float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    if (!(emax > 7.888609052210118e-31f)) emax = 1.f;  // a single point (or an extent below 2^-100: 1/s must stay finite)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float e = fmaxf(ext[k], emax * 9.5367431640625e-07f);  // flat axis: keep a positive step (2^-20 of the largest extent)
        const float s = __fdiv_ru(e, kGridSteps);
        scene[8 + k] = __float_as_uint(__fadd_rd(lo[k], -__fmul_ru(7.f, s)));
        scene[11 + k] = __float_as_uint(s);
    }
}

#if DRT_QNODE
__device__ __forceinline__ unsigned qpair(float lo, float hi, float g0, float inv_s)
{
    // outward: floor - 3 / ceil + 3 grid steps (the products are < 65536, their rounding error < 0.02 step)
    const int ql = max(0, (int)floorf((lo - g0) * inv_s) - 3);
    const int qh = min(65535, (int)ceilf((hi - g0) * inv_s) + 3);
    return (unsigned)ql | ((unsigned)qh << 16);
}

__global__ void emit_nodes_kernel(int n, const int2* __restrict__ children, const float4* __restrict__ blo,
                                  const float4* __restrict__ bhi, const unsigned* __restrict__ scene,
                                  uint4* __restrict__ nodes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float gx = __uint_as_float(scene[8]), gy = __uint_as_float(scene[9]), gz = __uint_as_float(scene[10]);
    const float ix = __fdiv_rn(1.f, __uint_as_float(scene[11])), iy = __fdiv_rn(1.f, __uint_as_float(scene[12])),
                iz = __fdiv_rn(1.f, __uint_as_float(scene[13]));
    if (n == 1) {
        if (i == 0) {  // single triangle: both slots are the one leaf (a duplicate test is harmless)
            float4 l = blo[0], h = bhi[0];
            const unsigned x = qpair(l.x, h.x, gx, ix), y = qpair(l.y, h.y, gy, iy), z = qpair(l.z, h.z, gz, iz);
            nodes[0] = make_uint4(x, x, y, y);
            nodes[1] = make_uint4(z, z, (unsigned)~0, (unsigned)~0);
        }
        return;
    }
    if (i >= n - 1) return;
    int2 ch = children[i];
    float4 l0 = blo[ch.x], h0 = bhi[ch.x], l1 = blo[ch.y], h1 = bhi[ch.y];
    int c0 = ch.x >= n - 1 ? ~(ch.x - (n - 1)) : ch.x;
    int c1 = ch.y >= n - 1 ? ~(ch.y - (n - 1)) : ch.y;
    uint4* node = nodes + (size_t)i * kNodeQuads;
    node[0] = make_uint4(qpair(l0.x, h0.x, gx, ix), qpair(l1.x, h1.x, gx, ix), qpair(l0.y, h0.y, gy, iy), qpair(l1.y, h1.y, gy, iy));
    node[1] = make_uint4(qpair(l0.z, h0.z, gz, iz), qpair(l1.z, h1.z, gz, iz), (unsigned)c0, (unsigned)c1);
}
// wide node i from binary node i (see BvhView::nodes4); boxes are read through L2 (also called from the cooperative build)
__device__ __forceinline__ void emit_wide_node(int i, int n, const int2* __restrict__ children, const float4* blo, const float4* bhi,
                                               const float g0[3], const float inv_s[3], uint4* __restrict__ nodes4)
{
    unsigned X[4], Y[4], Z[4], L[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { X[s] = Y[s] = Z[s] = 0x0000ffffu; L[s] = (unsigned)kEmptyLink; }  // lo = 65535, hi = 0: never entered
    auto put = [&](int s, int c) {  // c in build numbering: internal < n-1 <= leaf
        const float4 l = __ldcg(&blo[c]), h = __ldcg(&bhi[c]);
        X[s] = qpair(l.x, h.x, g0[0], inv_s[0]);
        Y[s] = qpair(l.y, h.y, g0[1], inv_s[1]);
        Z[s] = qpair(l.z, h.z, g0[2], inv_s[2]);
        L[s] = (unsigned)(c >= n - 1 ? ~(c - (n - 1)) : c);
    };
    if (n == 1) {
        put(0, 0);  // the only leaf (build number n-1 = 0)
    } else {
        const int2 ch = children[i];
        const int side[2] = {ch.x, ch.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c = side[k];
            if (c >= n - 1) put(2 * k, c);
            else {
                const int2 g = children[c];
                put(2 * k, g.x);
                put(2 * k + 1, g.y);
            }
        }
    }
    uint4* node = nodes4 + (size_t)i * 4;
    node[0] = make_uint4(X[0], X[1], X[2], X[3]);
    node[1] = make_uint4(Y[0], Y[1], Y[2], Y[3]);
    node[2] = make_uint4(Z[0], Z[1], Z[2], Z[3]);
    node[3] = make_uint4(L[0], L[1], L[2], L[3]);
}

__global__ void emit_nodes4_kernel(int n, const int2* __restrict__ children, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                   const unsigned* __restrict__ scene, uint4* __restrict__ nodes4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n > 1 ? n - 1 : 1)) return;
    const float g0[3] = {__uint_as_float(scene[8]), __uint_as_float(scene[9]), __uint_as_float(scene[10])};
    const float inv_s[3] = {__fdiv_rn(1.f, __uint_as_float(scene[11])), __fdiv_rn(1.f, __uint_as_float(scene[12])),
                            __fdiv_rn(1.f, __uint_as_float(scene[13]))};
    emit_wide_node(i, n, children, blo, bhi, g0, inv_s, nodes4);
}
#else
__global__ void emit_nodes_kernel(int n, const int2* __restrict__ children, const float4* __restrict__ blo,
                                  const float4* __restrict__ bhi, const unsigned* __restrict__ scene,
                                  float4* __restrict__ nodes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float infl = __uint_as_float(scene[7]) * 7.62939453125e-06f;  // pmax * 2^-17
    if (n == 1) {
        if (i == 0) {  // single triangle: child 0 = the leaf, child 1 = an empty box that is never hit
            float4 l = blo[0], h = bhi[0];
            nodes[0] = planes(l.x, h.x, INFINITY, -INFINITY, infl);
            nodes[1] = planes(l.y, h.y, INFINITY, -INFINITY, infl);
            nodes[2] = planes(l.z, h.z, INFINITY, -INFINITY, infl);
            nodes[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.f, 0.f);
        }
        return;
    }
    if (i >= n - 1) return;
    int2 ch = children[i];
    float4 l0 = blo[ch.x], h0 = bhi[ch.x], l1 = blo[ch.y], h1 = bhi[ch.y];
    int c0 = ch.x >= n - 1 ? ~(ch.x - (n - 1)) : ch.x;
    int c1 = ch.y >= n - 1 ? ~(ch.y - (n - 1)) : ch.y;
    float4* node = nodes + (size_t)i * kNodeQuads;
    node[0] = planes(l0.x, h0.x, l1.x, h1.x, infl);
    node[1] = planes(l0.y, h0.y, l1.y, h1.y, infl);
    node[2] = planes(l0.z, h0.z, l1.z, h1.z, infl);
    node[3] = make_float4(__int_as_float(c0), __int_as_float(c1), 0.f, 0.f);
}

#endif  // DRT_QNODE

// Round 2: grid + node + triangle emission as ONE launch (three launches of the 50 k-triangle rebuild's chain less): every thread
// derives the quantisation grid from the root box itself (same arithmetic as grid_kernel, so the same values), thread 0 publishes
// it in scene[8..13] for the traversal kernels.
#if DRT_QNODE
__global__ void emit_all_kernel(int n, const int32_t* __restrict__ F, const float* __restrict__ V, const uint64_t* __restrict__ keys,
                                const int2* __restrict__ children, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                unsigned* __restrict__ scene, uint4* __restrict__ nodes, uint4* __restrict__ nodes4, double2* __restrict__ tris)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float g0[3], st[3];
    {
        const float4 l = blo[0], h = bhi[0];
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
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                scene[8 + k] = __float_as_uint(g0[k]);
                scene[11 + k] = __float_as_uint(st[k]);
            }
        }
    }
    const float inv_s[3] = {__fdiv_rn(1.f, st[0]), __fdiv_rn(1.f, st[1]), __fdiv_rn(1.f, st[2])};
    if (n == 1) {
        if (i == 0) {  // single triangle: both slots are the one leaf (a duplicate test is harmless)
            float4 l = blo[0], h = bhi[0];
            const unsigned x = qpair(l.x, h.x, g0[0], inv_s[0]), y = qpair(l.y, h.y, g0[1], inv_s[1]), z = qpair(l.z, h.z, g0[2], inv_s[2]);
            nodes[0] = make_uint4(x, x, y, y);
            nodes[1] = make_uint4(z, z, (unsigned)~0, (unsigned)~0);
            if (nodes4) emit_wide_node(0, n, children, blo, bhi, g0, inv_s, nodes4);
        }
    } else if (i < n - 1) {
        int2 ch = children[i];
        float4 l0 = blo[ch.x], h0 = bhi[ch.x], l1 = blo[ch.y], h1 = bhi[ch.y];
        int c0 = ch.x >= n - 1 ? ~(ch.x - (n - 1)) : ch.x;
        int c1 = ch.y >= n - 1 ? ~(ch.y - (n - 1)) : ch.y;
        uint4* node = nodes + (size_t)i * kNodeQuads;
        node[0] = make_uint4(qpair(l0.x, h0.x, g0[0], inv_s[0]), qpair(l1.x, h1.x, g0[0], inv_s[0]), qpair(l0.y, h0.y, g0[1], inv_s[1]),
                             qpair(l1.y, h1.y, g0[1], inv_s[1]));
        node[1] = make_uint4(qpair(l0.z, h0.z, g0[2], inv_s[2]), qpair(l1.z, h1.z, g0[2], inv_s[2]), (unsigned)c0, (unsigned)c1);
        if (nodes4) emit_wide_node(i, n, children, blo, bhi, g0, inv_s, nodes4);
    }
    if (i < n) {
        int f = key_tri(keys[i]);
        const float* pa = &V[3 * (size_t)F[3 * f]];
        const float* pb = &V[3 * (size_t)F[3 * f + 1]];
        const float* pc = &V[3 * (size_t)F[3 * f + 2]];
        d3 a = mk3((double)pa[0], (double)pa[1], (double)pa[2]);
        d3 e1 = mk3((double)pb[0], (double)pb[1], (double)pb[2]) - a;
        d3 e2 = mk3((double)pc[0], (double)pc[1], (double)pc[2]) - a;
        double2* t = tris + (size_t)i * kTriD2;
        t[0] = make_double2(a.x, a.y);
        t[1] = make_double2(a.z, e1.x);
        t[2] = make_double2(e1.y, e1.z);
        t[3] = make_double2(e2.x, e2.y);
        t[4] = make_double2(e2.z, __longlong_as_double((long long)f));
    }
}
#endif

__global__ void emit_tris_kernel(const int32_t* __restrict__ F, const float* __restrict__ V,
                                 const uint64_t* __restrict__ keys, int n, double2* __restrict__ tris)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int f = key_tri(keys[k]);
    const float* pa = &V[3 * (size_t)F[3 * f]];
    const float* pb = &V[3 * (size_t)F[3 * f + 1]];
    const float* pc = &V[3 * (size_t)F[3 * f + 2]];
    d3 a = mk3((double)pa[0], (double)pa[1], (double)pa[2]);
    d3 e1 = mk3((double)pb[0], (double)pb[1], (double)pb[2]) - a;
    d3 e2 = mk3((double)pc[0], (double)pc[1], (double)pc[2]) - a;
    double2* t = tris + (size_t)k * kTriD2;
    t[0] = make_double2(a.x, a.y);
    t[1] = make_double2(a.z, e1.x);
    t[2] = make_double2(e1.y, e1.z);
    t[3] = make_double2(e2.x, e2.y);
    t[4] = make_double2(e2.z, __longlong_as_double((long long)f));
}

// faces must index inside [0,nV): checked on the device, result read lazily by drt_bvh_info
__global__ void validate_faces_kernel(const int32_t* __restrict__ F, int n3, int nV, int* __restrict__ bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3 && (F[i] < 0 || F[i] >= nV)) atomicAdd(bad, 1);
}

}  // namespace drt
