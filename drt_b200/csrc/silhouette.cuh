// silhouette.cuh -- the silhouette-edge side of the plugin's second consumer, fused:
//   silhouette_classify_kernel : Scene.silhouette_edge      DiffRender.py:445-457 (+ edge_face_norm :150-163)
//   silhouette_sample_kernel   : Scene.primary_visibility   DiffRender.py:459-479  +  primary_edge_sample.forward :189-241
//                                (projection of both edge ends, midpoint sample, image-space edge normal, the two probe
//                                 rays one pixel either side of the edge through the BVH, f = cover_upper - cover_lower,
//                                 the in-image filter) -- one launch instead of ~40 PyTorch ops + one query launch
//   silhouette_backward_kernel : primary_edge_sample.backward :243-267 chained through the projection (autograd of :465-472)
//                                into grad_V
// float64 throughout like the reference (captured_data.py:9); products are rounded separately (no FMA contraction) and
// summed left to right, so pixel positions agree with the PyTorch evaluation to a few ulp.
#pragma once
#include "trace.cuh"

namespace drt {

__device__ __forceinline__ d3 unit_face_normal(const double* __restrict__ V, const int32_t* __restrict__ tri)
{
    const d3 a = ld3(V + 3 * (size_t)tri[0]), b = ld3(V + 3 * (size_t)tri[1]), c = ld3(V + 3 * (size_t)tri[2]);
    const d3 n = cross(b - a, c - a);
    return divs(n, __dsqrt_rn(dot(n, n)));
}

// flags[e] = 1 iff the two faces on edge e face opposite ways as seen from `origin`
__global__ void __launch_bounds__(256) silhouette_classify_kernel(const double* __restrict__ V, const int32_t* __restrict__ e2f, int64_t nE,
                                                                   const double* __restrict__ origin, uint8_t* __restrict__ flags)
{
    const d3 o = ld3(origin);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t* f1 = e2f + 6 * e;
        const int32_t* f2 = f1 + 3;
        const double d1 = dot(unit_face_normal(V, f1), o - ld3(V + 3 * (size_t)f1[0]));
        const double d2 = dot(unit_face_normal(V, f2), o - ld3(V + 3 * (size_t)f2[0]));
        flags[e] = (d1 > 0.0) != (d2 > 0.0) ? 1 : 0;
    }
}

struct Camera {  // DEVICE pointers to the camera_M tuple of captured_data.py:112-118 (row-major float64), read through L1
    const double* __restrict__ R;   // [4,4] world -> camera
    const double* __restrict__ K;   // [3,3] intrinsics
    const double* __restrict__ Ri;  // [4,4] camera -> world
    const double* __restrict__ Ki;  // [3,3]
};

// pixel position of a world point: K @ (R @ [v,1])[:3], then x/z, y/z (DiffRender.py:465-472); p = K @ camera point
__device__ __forceinline__ void project(const Camera& cam, d3 v, d3& p)
{
    d3 c;
    c.x = addr(addr(addr(mulr(__ldg(cam.R + 0), v.x), mulr(__ldg(cam.R + 1), v.y)), mulr(__ldg(cam.R + 2), v.z)), __ldg(cam.R + 3));
    c.y = addr(addr(addr(mulr(__ldg(cam.R + 4), v.x), mulr(__ldg(cam.R + 5), v.y)), mulr(__ldg(cam.R + 6), v.z)), __ldg(cam.R + 7));
    c.z = addr(addr(addr(mulr(__ldg(cam.R + 8), v.x), mulr(__ldg(cam.R + 9), v.y)), mulr(__ldg(cam.R + 10), v.z)), __ldg(cam.R + 11));
    p.x = addr(addr(mulr(__ldg(cam.K + 0), c.x), mulr(__ldg(cam.K + 1), c.y)), mulr(__ldg(cam.K + 2), c.z));
    p.y = addr(addr(mulr(__ldg(cam.K + 3), c.x), mulr(__ldg(cam.K + 4), c.y)), mulr(__ldg(cam.K + 5), c.z));
    p.z = addr(addr(mulr(__ldg(cam.K + 6), c.x), mulr(__ldg(cam.K + 7), c.y)), mulr(__ldg(cam.K + 8), c.z));
}

// world-space direction of the ray through pixel (x, y): R_inv @ [K_inv @ [x,y,1], 1] - origin, NOT normalised (:213-222)
__device__ __forceinline__ d3 probe_direction(const Camera& cam, double x, double y, d3 o)
{
    d3 c, w;
    c.x = addr(addr(mulr(__ldg(cam.Ki + 0), x), mulr(__ldg(cam.Ki + 1), y)), __ldg(cam.Ki + 2));
    c.y = addr(addr(mulr(__ldg(cam.Ki + 3), x), mulr(__ldg(cam.Ki + 4), y)), __ldg(cam.Ki + 5));
    c.z = addr(addr(mulr(__ldg(cam.Ki + 6), x), mulr(__ldg(cam.Ki + 7), y)), __ldg(cam.Ki + 8));
    w.x = addr(addr(addr(mulr(__ldg(cam.Ri + 0), c.x), mulr(__ldg(cam.Ri + 1), c.y)), mulr(__ldg(cam.Ri + 2), c.z)), __ldg(cam.Ri + 3));
    w.y = addr(addr(addr(mulr(__ldg(cam.Ri + 4), c.x), mulr(__ldg(cam.Ri + 5), c.y)), mulr(__ldg(cam.Ri + 6), c.z)), __ldg(cam.Ri + 7));
    w.z = addr(addr(addr(mulr(__ldg(cam.Ri + 8), c.x), mulr(__ldg(cam.Ri + 9), c.y)), mulr(__ldg(cam.Ri + 10), c.z)), __ldg(cam.Ri + 11));
    return w - o;
}

// Two threads per silhouette edge (one per probe ray).  Per edge k: index_xy[k] = the midpoint sample truncated to a pixel,
// f[k] = cover(upper probe) - cover(lower probe) in {-1, 0, 1}, keep[k] = |f| > 1e-5 and the pixel inside the image
// (DiffRender.py:236, :476).
__global__ void __launch_bounds__(128) silhouette_sample_kernel(BvhView B, const double* __restrict__ V, const int64_t* __restrict__ edges,
                                                                int64_t k, Camera cam, const double* __restrict__ origin, int resx, int resy,
                                                                int64_t* __restrict__ index_xy, double* __restrict__ f_out,
                                                                uint8_t* __restrict__ keep)
{
    const d3 o = ld3(origin);
    const int64_t n2 = 2 * k;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x; base < n2; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = base + threadIdx.x;
        const bool act = t < n2;
        const int64_t e = t >> 1;
        const int upper = (int)(t & 1) == 0;
        double mx = 0.0, my = 0.0;
        int hit = 0;
        if (act) {
            d3 pa, pb;
            project(cam, ld3(V + 3 * (size_t)edges[2 * e]), pa);
            project(cam, ld3(V + 3 * (size_t)edges[2 * e + 1]), pb);
            const double ax = __ddiv_rn(pa.x, pa.z), ay = __ddiv_rn(pa.y, pa.z), bx = __ddiv_rn(pb.x, pb.z), by = __ddiv_rn(pb.y, pb.z);
            mx = __ddiv_rn(addr(ax, bx), 2.0);
            my = __ddiv_rn(addr(ay, by), 2.0);
            const double nx = subr(ay, by), ny = subr(bx, ax);                      // image-space edge normal (:204-206)
            const double len = __dsqrt_rn(addr(mulr(nx, nx), mulr(ny, ny)));
            const double ux = __ddiv_rn(nx, len), uy = __ddiv_rn(ny, len);
            const double px = upper ? addr(mx, ux) : subr(mx, ux), py = upper ? addr(my, uy) : subr(my, uy);  // eps = 1 pixel
            const d3 dir = probe_direction(cam, px, py, o);
            double tt;
            int id;
            traverse<true>(B, cast_ray(o, dir), tt, id);                           // hit / no hit is all that is used (:225-227)
            hit = id >= 0 ? 1 : 0;
        }
        const int other = __shfl_xor_sync(0xffffffffu, hit, 1);
        if (act && upper) {
            const double f = (double)(hit - other);
            const int64_t ix = (int64_t)mx, iy = (int64_t)my;                       // .to(torch.long): truncation
            index_xy[2 * e] = ix;
            index_xy[2 * e + 1] = iy;
            f_out[e] = f;
            keep[e] = (fabs(f) > 1e-5 && ix < resx - 1 && iy < resy - 1 && ix >= 0 && iy >= 0) ? 1 : 0;
        }
    }
}

// grad_V += d(sum_j g_out[j] * output_j) / d vertices for the kept samples j (edge slot kept_idx[j]):
//   d output / d E_pos[endpoint][coord] = -N[coord] * f      (the reference's hand-written backward, :243-249, 262-266)
//   E_pos = (p.x / p.z, p.y / p.z),  p = K c,  c = (R [v,1])[:3]  with c.z detached when detach_depth (:468-469)
__global__ void __launch_bounds__(128) silhouette_backward_kernel(const double* __restrict__ V, const int64_t* __restrict__ edges, Camera cam,
                                                                  int detach_depth, const double* __restrict__ f_in,
                                                                  const int64_t* __restrict__ kept_idx, const float* __restrict__ g_out, int64_t m,
                                                                  double* __restrict__ gV)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = kept_idx[j];
        const int64_t va = edges[2 * e], vb = edges[2 * e + 1];
        d3 pa, pb;
        project(cam, ld3(V + 3 * (size_t)va), pa);
        project(cam, ld3(V + 3 * (size_t)vb), pb);
        const double ax = pa.x / pa.z, ay = pa.y / pa.z, bx = pb.x / pb.z, by = pb.y / pb.z;
        const double s = -f_in[e] * (double)g_out[j];
        const double gx = (ay - by) * s, gy = (bx - ax) * s;  // gradient w.r.t. the pixel position of EITHER end
#pragma unroll
        for (int end = 0; end < 2; ++end) {
            const d3 p = end ? pb : pa;
            const d3 gp = mk3(gx / p.z, gy / p.z, -(gx * p.x + gy * p.y) / (p.z * p.z));
            d3 gc = mk3(__ldg(cam.K + 0) * gp.x + __ldg(cam.K + 3) * gp.y + __ldg(cam.K + 6) * gp.z, __ldg(cam.K + 1) * gp.x + __ldg(cam.K + 4) * gp.y + __ldg(cam.K + 7) * gp.z,
                        __ldg(cam.K + 2) * gp.x + __ldg(cam.K + 5) * gp.y + __ldg(cam.K + 8) * gp.z);
            if (detach_depth) gc.z = 0.0;
            double* g = gV + 3 * (size_t)(end ? vb : va);
            atomicAdd(g, __ldg(cam.R + 0) * gc.x + __ldg(cam.R + 4) * gc.y + __ldg(cam.R + 8) * gc.z);
            atomicAdd(g + 1, __ldg(cam.R + 1) * gc.x + __ldg(cam.R + 5) * gc.y + __ldg(cam.R + 9) * gc.z);
            atomicAdd(g + 2, __ldg(cam.R + 2) * gc.x + __ldg(cam.R + 6) * gc.y + __ldg(cam.R + 10) * gc.z);
        }
    }
}

}  // namespace drt
