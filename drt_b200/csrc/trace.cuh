// trace.cuh -- BVH traversal, fused forward tracer, backward kernel (sm_100a).
#pragma once
#include <cooperative_groups.h>

#include <climits>

#include "bvh.cuh"

namespace drt {

constexpr int kStackDepth = 96;  // >= 63 Morton bits + index tie-break levels of a Karras tree

struct QRay {        // query ray = float32 cast of the chain's float64 ray (DiffRender.py:387-388)
    float ox, oy, oz, dx, dy, dz;
};

__device__ __forceinline__ QRay cast_ray(d3 o, d3 d)
{
    return QRay{__double2float_rn(o.x), __double2float_rn(o.y), __double2float_rn(o.z),
                __double2float_rn(d.x), __double2float_rn(d.y), __double2float_rn(d.z)};
}

// Closed-box slab test with near/far planes picked by the direction sign.  A zero direction
// component yields +-inf (origin strictly inside / outside the slab) or NaN (origin exactly on a
// slab plane); fmaxf/fminf drop the NaN, i.e. "no constraint", which is right for a ray running
// inside a face plane.  (plane - o) * inv carries <= 3 half-ulps of relative error, the 2^-20
// slack below makes the test conservative, so the exact float64 triangle test decides every hit.
__device__ __forceinline__ bool slab(float lx, float ly, float lz, float hx, float hy, float hz, const QRay& r,
                                     float ix, float iy, float iz, bool nx, bool ny, bool nz, float tmax, float& tnear)
{
    float t0 = fmaxf(fmaxf(fmaxf(0.f, ((nx ? hx : lx) - r.ox) * ix), ((ny ? hy : ly) - r.oy) * iy),
                     ((nz ? hz : lz) - r.oz) * iz);
    float t1 = fminf(fminf(fminf(tmax, ((nx ? lx : hx) - r.ox) * ix), ((ny ? ly : hy) - r.oy) * iy),
                     ((nz ? lz : hz) - r.oz) * iz);
    tnear = t0;
    return t0 <= t1 * 1.00000095367431640625f;
}

// Exact closest hit (ANY = false) or first hit found (ANY = true; only hit/no-hit is meaningful,
// which is all the reference uses of its third query: DiffRender.py:426-427).
// Returns triangle id (-1 on miss) and the float64 distance along the float32 ray.
template <bool ANY, bool PROF = false>
__device__ __forceinline__ void traverse(const BvhView& B, const QRay& r, double& t_best, int& id_best,
                                         unsigned long long* prof = nullptr)
{
    t_best = INFINITY;
    id_best = -1;
    if (B.nTris <= 0) return;
    const float ix = __fdiv_rn(1.f, r.dx), iy = __fdiv_rn(1.f, r.dy), iz = __fdiv_rn(1.f, r.dz);
    const bool nx = signbit(ix), ny = signbit(iy), nz = signbit(iz);
    const d3 o = mk3((double)r.ox, (double)r.oy, (double)r.oz);
    const d3 d = mk3((double)r.dx, (double)r.dy, (double)r.dz);
    float tmax = INFINITY;
    int stack[kStackDepth];
    int sp = 0;
    int node = 0;
    for (;;) {
        while (node >= 0) {
            if (PROF) {  // utilisation probe: warp-iterations vs lane-iterations of the internal-node loop
                unsigned m = __activemask();
                if ((threadIdx.x & 31) == __ffs(m) - 1) { atomicAdd(prof, 1ull); atomicAdd(prof + 1, (unsigned long long)__popc(m)); }
            }
            const float4* p = B.nodes + (size_t)node * kNodeQuads;
            float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
            float ta, tb;
            bool ha = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, r, ix, iy, iz, nx, ny, nz, tmax, ta);
            bool hb = slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, r, ix, iy, iz, nx, ny, nz, tmax, tb);
            int ca = __float_as_int(q3.x), cb = __float_as_int(q3.y);
            if (ha && hb) {
                bool a_first = ta <= tb;
                node = a_first ? ca : cb;
                if (sp < kStackDepth) stack[sp++] = a_first ? cb : ca;
            } else if (ha) {
                node = ca;
            } else if (hb) {
                node = cb;
            } else {
                if (sp == 0) return;
                node = stack[--sp];
            }
        }
        {
            if (PROF) {
                unsigned m = __activemask();
                if ((threadIdx.x & 31) == __ffs(m) - 1) { atomicAdd(prof + 2, 1ull); atomicAdd(prof + 3, (unsigned long long)__popc(m)); }
            }
            const float4* p = B.tris + (size_t)(~node) * kTriQuads;
            float4 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
            double t;
            if (query_tri(o, d, mk3((double)r0.x, (double)r0.y, (double)r0.z), mk3((double)r0.w, (double)r1.x, (double)r1.y),
                          mk3((double)r1.z, (double)r1.w, (double)r2.x), t)) {
                int id = __float_as_int(r2.y);
                if (t < t_best || (t == t_best && id < id_best)) {
                    t_best = t;
                    id_best = id;
                    tmax = __double2float_ru(t);
                }
                if (ANY) return;
            }
        }
        if (sp == 0) return;
        node = stack[--sp];
    }
}

// ---------------------------------------------------------------------------------------------
// optix_mesh::intersect replacement (optix_extend.cpp:29-57)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) closest_hit_kernel(BvhView B, const float* __restrict__ ray6, int64_t N,
                                                          float* __restrict__ T, int32_t* __restrict__ ID,
                                                          int64_t strideT, int64_t strideID)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float2* p = reinterpret_cast<const float2*>(ray6 + 6 * i);
        float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        QRay r{a.x, a.y, b.x, b.y, c.x, c.y};
        double t;
        int id;
        traverse<false>(B, r, t, id);
        T[i * strideT] = id >= 0 ? __double2float_rn(t) : -1.f;
        ID[i * strideID] = id;
    }
}

__device__ __forceinline__ void load_tri64(const BvhView& B, const double* __restrict__ V64, int id, d3& a0, d3& a1, d3& a2)
{
    int i0 = __ldg(&B.F[3 * (size_t)id]), i1 = __ldg(&B.F[3 * (size_t)id + 1]), i2 = __ldg(&B.F[3 * (size_t)id + 2]);
    a0 = ld3(V64 + 3 * (size_t)i0);
    a1 = ld3(V64 + 3 * (size_t)i1);
    a2 = ld3(V64 + 3 * (size_t)i2);
}

__device__ __forceinline__ void write_invalid(double* __restrict__ out_ori, double* __restrict__ out_dir,
                                              uint8_t* __restrict__ mask3, int64_t i)
{
    st3(out_ori + 3 * i, mk3(0, 0, 0));
    st3(out_dir + 3 * i, mk3(0, 0, 0));
    mask3[3 * i] = 0; mask3[3 * i + 1] = 0; mask3[3 * i + 2] = 0;
}

// ---------------------------------------------------------------------------------------------
// Scene.render_transparent replacement, one launch (DiffRender.py:420-432).  v1: one thread per
// ray walks the whole path Q1 -> refract -> Q2 -> refract -> Q3.
// ---------------------------------------------------------------------------------------------
template <bool PROF>
__global__ void __launch_bounds__(128) trace_fwd_kernel(BvhView B, const double* __restrict__ V64,
                                                        const double* __restrict__ origin, const double* __restrict__ dir,
                                                        int64_t N, double ext_ior, double int_ior,
                                                        double* __restrict__ out_ori, double* __restrict__ out_dir,
                                                        uint8_t* __restrict__ mask3, int4* __restrict__ rec,
                                                        int* __restrict__ rec_count, uint8_t* __restrict__ hit1,
                                                        unsigned long long* __restrict__ prof)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        d3 o = ld3(origin + 3 * i), d = ld3(dir + 3 * i);
        d3 oo = mk3(0, 0, 0), od = mk3(0, 0, 0);
        int id1, id2 = -1, id3;
        double t;
        bool valid = false;
        traverse<false, PROF>(B, cast_ray(o, d), t, id1, prof);
        if (id1 >= 0) {
            HitRec h;
            d3 a0, a1, a2, o1, d1;
            load_tri64(B, V64, id1, a0, a1, a2);
            hit_forward(h, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
            if (!h.tir) {
                traverse<false, PROF>(B, cast_ray(o1, d1), t, id2, prof + 4);
                if (id2 >= 0) {
                    d3 o2, d2;
                    load_tri64(B, V64, id2, a0, a1, a2);
                    hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
                    if (!h.tir) {
                        traverse<true, PROF>(B, cast_ray(o2, d2), t, id3, prof + 8);
                        if (id3 < 0) {
                            valid = true;
                            oo = o2;
                            od = d2;
                        }
                    }
                }
            }
        }
        st3(out_ori + 3 * i, oo);
        st3(out_dir + 3 * i, od);
        uint8_t m = valid ? 1 : 0;
        mask3[3 * i] = m; mask3[3 * i + 1] = m; mask3[3 * i + 2] = m;
        if (rec && valid) rec[atomicAdd(rec_count, 1)] = make_int4((int)i, id1, id2, 0);
        if (hit1) hit1[i] = id1 >= 0 ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Backward: replay cached hit records, analytic Jacobian, scatter-add into grad_V (optim.py:210;
// the reference's two index_put_(accumulate=True) of `vertices[faces]`, DiffRender.py:495-496).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter3(double* __restrict__ gV, int v, d3 g)
{
    atomicAdd(gV + 3 * (size_t)v, g.x);
    atomicAdd(gV + 3 * (size_t)v + 1, g.y);
    atomicAdd(gV + 3 * (size_t)v + 2, g.z);
}

__global__ void __launch_bounds__(128) trace_bwd_kernel(BvhView B, const double* __restrict__ V64,
                                                        const double* __restrict__ origin, const double* __restrict__ dir,
                                                        double ext_ior, double int_ior, const int4* __restrict__ rec,
                                                        const int* __restrict__ rec_count,
                                                        const double* __restrict__ g_ori, const double* __restrict__ g_dir,
                                                        double* __restrict__ gV)
{
    const int n = __ldg(rec_count);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 rc = __ldg(rec + k);
        const int64_t i = rc.x;
        const int id1 = rc.y, id2 = rc.z;
        d3 o = ld3(origin + 3 * i), d = ld3(dir + 3 * i);
        HitRec h1, h2;
        d3 a0, a1, a2, o1, d1, o2, d2;
        load_tri64(B, V64, id1, a0, a1, a2);
        hit_forward(h1, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
        load_tri64(B, V64, id2, a0, a1, a2);
        hit_forward(h2, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
        d3 go2 = g_ori ? ld3(g_ori + 3 * i) : mk3(0, 0, 0);
        d3 gd2 = ld3(g_dir + 3 * i);
        d3 ga[3] = {mk3(0, 0, 0), mk3(0, 0, 0), mk3(0, 0, 0)}, go1, gd1, go0, gd0;
        hit_backward(h2, go2, gd2, ga, go1, gd1);
        const int32_t* f2 = B.F + 3 * (size_t)id2;
        scatter3(gV, f2[0], ga[0]); scatter3(gV, f2[1], ga[1]); scatter3(gV, f2[2], ga[2]);
        ga[0] = ga[1] = ga[2] = mk3(0, 0, 0);
        hit_backward(h1, go1, gd1, ga, go0, gd0);
        const int32_t* f1 = B.F + 3 * (size_t)id1;
        scatter3(gV, f1[0], ga[0]); scatter3(gV, f1[1], ga[1]); scatter3(gV, f1[2], ga[2]);
    }
}

// ---------------------------------------------------------------------------------------------
// Loss_calculator.ray_loss consumer (optim.py:96-106) as one pass: g_out_dir and the loss value.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ray_loss_grad_kernel(const double* __restrict__ out_ori,
                                                            const double* __restrict__ out_dir,
                                                            const uint8_t* __restrict__ mask3,
                                                            const double* __restrict__ screen,
                                                            const uint8_t* __restrict__ valid, int64_t N,
                                                            double* __restrict__ g_dir, double* __restrict__ loss_sum)
{
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        d3 g = mk3(0, 0, 0);
        if (mask3[3 * i] && (!valid || valid[i])) {
            d3 tg = ld3(screen + 3 * i) - ld3(out_ori + 3 * i);
            tg = divs(tg, __dsqrt_rn(dot(tg, tg)));
            d3 df = ld3(out_dir + 3 * i) - tg;
            acc += dot(df, df);
            g = df * 2.0;
        }
        st3(g_dir + 3 * i, g);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        if (loss_sum && s != 0.0) atomicAdd(loss_sum, s);
    }
}

}  // namespace drt
