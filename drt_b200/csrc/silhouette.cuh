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

// ------------------------------------------------------------------------------------------------------------------------
// The two loss terms around the ray loss in one optim.py iteration, FUSED like drt_ray_loss_step fuses the ray loss: value and
// vertex gradient from one launch, no intermediate tensors, no host synchronisation (the drop-in methods above have to return
// data-dependent-size tensors and therefore sync twice per view; a whole iteration measured 5.9 ms with them, of which the
// ray path is 0.6 ms).
//
//   silhouette_loss_kernel : Loss_calculator.vh_loss (optim.py:67-80), up to 8 views per launch (blockIdx.y = view): silhouette_edge + primary_visibility +
//       primary_edge_sample + `(mask[index] - output).abs().sum()` + its backward.  One thread per edge of the mesh: classify;
//       a silhouette edge projects its ends, traces the two probe rays, and if the sample is kept adds |mask[y,x] - 0.5| to the
//       loss and  d loss/d output = -sign(mask[y,x] - 0.5)  times the reference's dE_pos = -N f, chained through the
//       projection, to grad_V.
//   dihedral_loss_kernel   : Loss_calculator.sm_loss (optim.py:82-89): sum over edges of -log(1 + n1.n2) with the analytic
//       gradient through both unit face normals (DiffRender.py:150-163, 440-443).
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add3(double* __restrict__ g, d3 v)
{
    atomicAdd(g, v.x);
    atomicAdd(g + 1, v.y);
    atomicAdd(g + 2, v.z);
}

constexpr int kMaxSilhouetteViews = 8;  // optim.py:72 uses 8 views per iteration
struct SilhouetteViews {  // per-view device pointers, passed by value; blockIdx.y selects the view
    Camera cam[kMaxSilhouetteViews];
    const double* origin[kMaxSilhouetteViews];
    const double* mask[kMaxSilhouetteViews];
};

__global__ void __launch_bounds__(128) silhouette_loss_kernel(BvhView B, const double* __restrict__ V, const int64_t* __restrict__ edges,
                                                              const int32_t* __restrict__ e2f, int64_t nE, SilhouetteViews views,
                                                              int resx, int resy, int detach_depth, double* __restrict__ loss_sum,
                                                              double* __restrict__ gV, int* __restrict__ n_samples)
{
    const Camera cam = views.cam[blockIdx.y];
    const double* __restrict__ mask = views.mask[blockIdx.y];
    const d3 o = ld3(views.origin[blockIdx.y]);
    double acc = 0.0;
    int kept = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t* f1 = e2f + 6 * e;
        const int32_t* f2 = f1 + 3;
        const double d1 = dot(unit_face_normal(V, f1), o - ld3(V + 3 * (size_t)f1[0]));
        const double d2 = dot(unit_face_normal(V, f2), o - ld3(V + 3 * (size_t)f2[0]));
        if ((d1 > 0.0) == (d2 > 0.0)) continue;                                   // not a silhouette edge (DiffRender.py:456)
        const int64_t va = edges[2 * e], vb = edges[2 * e + 1];
        d3 pa, pb;
        project(cam, ld3(V + 3 * (size_t)va), pa);
        project(cam, ld3(V + 3 * (size_t)vb), pb);
        const double ax = __ddiv_rn(pa.x, pa.z), ay = __ddiv_rn(pa.y, pa.z), bx = __ddiv_rn(pb.x, pb.z), by = __ddiv_rn(pb.y, pb.z);
        const double mx = __ddiv_rn(addr(ax, bx), 2.0), my = __ddiv_rn(addr(ay, by), 2.0);
        const double nx = subr(ay, by), ny = subr(bx, ax);
        const double len = __dsqrt_rn(addr(mulr(nx, nx), mulr(ny, ny)));
        const double ux = __ddiv_rn(nx, len), uy = __ddiv_rn(ny, len);
        int cover[2];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double px = side ? subr(mx, ux) : addr(mx, ux), py = side ? subr(my, uy) : addr(my, uy);
            double tt;
            int id;
            traverse<true>(B, cast_ray(o, probe_direction(cam, px, py, o)), tt, id);
            cover[side] = id >= 0 ? 1 : 0;
        }
        const double f = (double)(cover[0] - cover[1]);
        const int64_t ix = (int64_t)mx, iy = (int64_t)my;
        if (!(fabs(f) > 1e-5 && ix < resx - 1 && iy < resy - 1 && ix >= 0 && iy >= 0)) continue;
        const double diff = mask[iy * (int64_t)resx + ix] - 0.5;                  // output = 0.5 (DiffRender.py:240), optim.py:79
        acc += fabs(diff);
        ++kept;
        if (gV) {
            const double g_out = diff > 0.0 ? -1.0 : (diff < 0.0 ? 1.0 : 0.0);    // d |mask - output| / d output
            const double s = -f * g_out;
            const double gx = nx * s, gy = ny * s;                                // w.r.t. the pixel position of either end
#pragma unroll
            for (int end = 0; end < 2; ++end) {
                const d3 p = end ? pb : pa;
                const d3 gp = mk3(gx / p.z, gy / p.z, -(gx * p.x + gy * p.y) / (p.z * p.z));
                d3 gc = mk3(__ldg(cam.K + 0) * gp.x + __ldg(cam.K + 3) * gp.y + __ldg(cam.K + 6) * gp.z,
                            __ldg(cam.K + 1) * gp.x + __ldg(cam.K + 4) * gp.y + __ldg(cam.K + 7) * gp.z,
                            __ldg(cam.K + 2) * gp.x + __ldg(cam.K + 5) * gp.y + __ldg(cam.K + 8) * gp.z);
                if (detach_depth) gc.z = 0.0;
                add3(gV + 3 * (size_t)(end ? vb : va),
                     mk3(__ldg(cam.R + 0) * gc.x + __ldg(cam.R + 4) * gc.y + __ldg(cam.R + 8) * gc.z,
                         __ldg(cam.R + 1) * gc.x + __ldg(cam.R + 5) * gc.y + __ldg(cam.R + 9) * gc.z,
                         __ldg(cam.R + 2) * gc.x + __ldg(cam.R + 6) * gc.y + __ldg(cam.R + 10) * gc.z));
            }
        }
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, sft);
        kept += __shfl_xor_sync(0xffffffffu, kept, sft);
    }
    if ((threadIdx.x & 31) == 0) {
        if (acc != 0.0) atomicAdd(loss_sum, acc);
        if (n_samples && kept) atomicAdd(n_samples, kept);
    }
}

// unit normal of triangle (a, b, c) with what its reverse pass needs
struct FaceN {
    d3 e1, e2, n;
    double L;
};
__device__ __forceinline__ FaceN face_normal(const double* __restrict__ V, const int32_t* __restrict__ tri)
{
    FaceN r;
    const d3 a = ld3(V + 3 * (size_t)tri[0]);
    r.e1 = ld3(V + 3 * (size_t)tri[1]) - a;
    r.e2 = ld3(V + 3 * (size_t)tri[2]) - a;
    const d3 N = cross(r.e1, r.e2);
    r.L = __dsqrt_rn(dot(N, N));
    r.n = divs(N, r.L);
    return r;
}
// gradient g_n w.r.t. the unit normal -> the three vertices (same chain as common.cuh:hit_backward's normal part)
__device__ __forceinline__ void face_normal_backward(const FaceN& fc, d3 g_n, const int32_t* __restrict__ tri, double* __restrict__ gV)
{
    const d3 g_N = (g_n - fc.n * dot(fc.n, g_n)) * __ddiv_rn(1.0, fc.L);
    const d3 g_e1 = cross(fc.e2, g_N), g_e2 = cross(g_N, fc.e1);
    add3(gV + 3 * (size_t)tri[1], g_e1);
    add3(gV + 3 * (size_t)tri[2], g_e2);
    add3(gV + 3 * (size_t)tri[0], -(g_e1 + g_e2));
}

__global__ void __launch_bounds__(256) dihedral_loss_kernel(const double* __restrict__ V, const int32_t* __restrict__ e2f, int64_t nE,
                                                            double* __restrict__ loss_sum, double* __restrict__ gV)
{
    double acc = 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nE; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t* f1 = e2f + 6 * e;
        const int32_t* f2 = f1 + 3;
        const FaceN a = face_normal(V, f1), b = face_normal(V, f2);
        const double c = dot(a.n, b.n);                                            // DiffRender.py:440-443
        acc += -log(1.0 + c);                                                      // optim.py:86-87
        if (gV) {
            const double g_c = -1.0 / (1.0 + c);
            face_normal_backward(a, b.n * g_c, f1, gV);
            face_normal_backward(b, a.n * g_c, f2, gV);
        }
    }
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(loss_sum, acc);
}

}  // namespace drt
