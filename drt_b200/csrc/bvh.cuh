// bvh.cuh -- device LBVH: data layout, build and refit (sm_100a).
//
// Replaces the closed OptiX Prime builder behind optix_mesh::update (reference
// optix_extend.cpp:61-67: setTriangles + update(RTP_MODEL_HINT_ASYNC)).
//
// Pipeline (all on the caller's stream, no host sync):
//   1 bounds   : per-triangle AABB centroid -> scene centroid bounds (warp shuffle + ordered-uint atomics)
//   2 morton   : 63-bit Morton code (21 bits/axis) of the normalised centroid
//   3 sort     : LSD radix sort of (code, triangle id) pairs          (radix_sort.cuh)
//   4 topology : Karras 2012 -- one thread per internal node finds its key range and split
//   5 fit      : bottom-up AABB union with atomic arrival counters
//   6 emit     : traversal layout -- 64-B nodes holding BOTH children's boxes, 48-B triangle records
//                in Morton order (float32 vertices + original id)
// Refit = steps 5-6 only (topology kept).
#pragma once
#include "common.cuh"

namespace drt {

// ---- traversal layout ---------------------------------------------------------------------------
// node i = 4 x float4:
//   q0 = (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x)
//   q1 = (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y)
//   q2 = (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z)
//   q3 = (child0, child1, -, -) as int bits; child >= 0: internal node index, child < 0: ~slot of a
//        triangle record (1 triangle per leaf)
// triangle record s = 3 x float4:
//   r0 = (v0.x, v0.y, v0.z, v1.x)  r1 = (v1.y, v1.z, v2.x, v2.y)  r2 = (v2.z, id bits, -, -)
constexpr int kNodeQuads = 4;
constexpr int kTriQuads = 3;

struct BvhView {
    const float4* nodes;
    const float4* tris;
    const int32_t* F;  // [nF,3] original faces
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

// scene[0..2] = enc(min centroid), scene[3..5] = enc(max centroid); caller presets to 0xffffffff / 0
__global__ void centroid_bounds_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, int nF,
                                       unsigned* __restrict__ scene)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {INFINITY, INFINITY, INFINITY}, C[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (f < nF) {
        float lo[3], hi[3];
        tri_box(F, V, f, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = C[k] = 0.5f * lo[k] + 0.5f * hi[k];
    }
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
    }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, int nF,
                              const unsigned* __restrict__ scene, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals)
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
        float s = fminf(fmaxf(u * 2097152.f, 0.f), 2097151.f);
        q[k] = (uint32_t)s;
    }
    keys[f] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[f] = (uint32_t)f;
}

// Karras delta: length of the common prefix of keys i and j (index tie-break), -1 out of range
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, uint64_t ki, int j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t kj = keys[j];
    if (ki == kj) return 64 + __clz(i ^ j);
    return __clzll((long long)(ki ^ kj));
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
__global__ void fit_kernel(const int32_t* __restrict__ F, const float* __restrict__ V, const uint32_t* __restrict__ vals,
                           int n, const int2* __restrict__ children, const int* __restrict__ parent,
                           float4* __restrict__ blo, float4* __restrict__ bhi, int* __restrict__ flags)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float lo[3], hi[3];
    tri_box(F, V, (int)vals[k], lo, hi);
    int me = n - 1 + k;
    blo[me] = make_float4(lo[0], lo[1], lo[2], 0.f);
    bhi[me] = make_float4(hi[0], hi[1], hi[2], 0.f);
    if (n == 1) return;
    int cur = parent[me];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1) == 0) return;  // first arrival: the sibling will finish this node
        __threadfence();
        int2 ch = children[cur];
        // volatile-style reads through L2: the sibling's stores were fenced before its atomic
        float4 l0 = __ldcg(&blo[ch.x]), h0 = __ldcg(&bhi[ch.x]);
        float4 l1 = __ldcg(&blo[ch.y]), h1 = __ldcg(&bhi[ch.y]);
        blo[cur] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        bhi[cur] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        cur = parent[cur];
    }
}

__global__ void emit_nodes_kernel(int n, const int2* __restrict__ children, const float4* __restrict__ blo,
                                  const float4* __restrict__ bhi, float4* __restrict__ nodes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (i == 0) {  // single triangle: child 0 = the leaf, child 1 = an empty box that is never hit
            float4 l = blo[0], h = bhi[0];
            nodes[0] = make_float4(l.x, l.y, l.z, h.x);
            nodes[1] = make_float4(h.y, h.z, INFINITY, INFINITY);
            nodes[2] = make_float4(INFINITY, -INFINITY, -INFINITY, -INFINITY);
            nodes[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.f, 0.f);
        }
        return;
    }
    if (i >= n - 1) return;
    int2 ch = children[i];
    float4 l0 = blo[ch.x], h0 = bhi[ch.x], l1 = blo[ch.y], h1 = bhi[ch.y];
    int c0 = ch.x >= n - 1 ? ~(ch.x - (n - 1)) : ch.x;
    int c1 = ch.y >= n - 1 ? ~(ch.y - (n - 1)) : ch.y;
    nodes[4 * (size_t)i + 0] = make_float4(l0.x, l0.y, l0.z, h0.x);
    nodes[4 * (size_t)i + 1] = make_float4(h0.y, h0.z, l1.x, l1.y);
    nodes[4 * (size_t)i + 2] = make_float4(l1.z, h1.x, h1.y, h1.z);
    nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(c0), __int_as_float(c1), 0.f, 0.f);
}

__global__ void emit_tris_kernel(const int32_t* __restrict__ F, const float* __restrict__ V,
                                 const uint32_t* __restrict__ vals, int n, float4* __restrict__ tris)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int f = (int)vals[k];
    const float* a = &V[3 * (size_t)F[3 * f]];
    const float* b = &V[3 * (size_t)F[3 * f + 1]];
    const float* c = &V[3 * (size_t)F[3 * f + 2]];
    tris[3 * (size_t)k + 0] = make_float4(a[0], a[1], a[2], b[0]);
    tris[3 * (size_t)k + 1] = make_float4(b[1], b[2], c[0], c[1]);
    tris[3 * (size_t)k + 2] = make_float4(c[2], __int_as_float(f), 0.f, 0.f);
}

// faces must index inside [0,nV): checked on the device, result read lazily by drt_bvh_info
__global__ void validate_faces_kernel(const int32_t* __restrict__ F, int n3, int nV, int* __restrict__ bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3 && (F[i] < 0 || F[i] >= nV)) atomicAdd(bad, 1);
}

}  // namespace drt
