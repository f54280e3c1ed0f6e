// plane.cuh -- optional background-plane step after the two refractions (sm_100a).
//
// NOT part of the reference's loss (optim.py:96-106 compares the exit DIRECTION with the direction towards a
// measured screen point; there is no plane intersection anywhere in DiffRender.py).  BASELINE.json's north-star lists
// "background-plane intersection" as the last step of the path, so it is provided as an option on top of
// Scene.render_transparent's outputs: where does the exit ray (out_ori, out_dir) of a valid path meet the plane
// {x : (x - p0).n = 0}?      s = ((p0 - o).n) / (d.n),   x = o + s d
// and its reverse     g_o = g_x - n (g_x.d)/(d.n),   g_d = s g_o.
// Rays that are not valid paths (mask = 0), run parallel to the plane or meet it behind their origin (s <= 0) get
// x = 0, front = 0 and no gradient.
#pragma once
#include "common.cuh"

namespace drt {

struct Plane {
    double px, py, pz, nx, ny, nz;
};

__device__ __forceinline__ bool plane_param(const Plane& P, d3 o, d3 d, double& s, double& dn)
{
    const d3 n = mk3(P.nx, P.ny, P.nz);
    dn = dot(d, n);
    if (dn == 0.0) return false;
    s = __ddiv_rn(dot(mk3(P.px, P.py, P.pz) - o, n), dn);
    return s > 0.0;
}

__global__ void __launch_bounds__(256) plane_hit_kernel(const double* __restrict__ out_ori, const double* __restrict__ out_dir,
                                                        const uint8_t* __restrict__ mask3, int64_t N, Plane P,
                                                        double* __restrict__ pts, uint8_t* __restrict__ front)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        d3 x = mk3(0, 0, 0);
        bool ok = false;
        if (mask3[3 * i]) {
            const d3 o = ld3(out_ori + 3 * i), d = ld3(out_dir + 3 * i);
            double s, dn;
            ok = plane_param(P, o, d, s, dn);
            if (ok) x = o + d * s;
        }
        st3(pts + 3 * i, x);
        if (front) front[i] = ok ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) plane_hit_bwd_kernel(const double* __restrict__ out_ori, const double* __restrict__ out_dir,
                                                            const uint8_t* __restrict__ mask3, int64_t N, Plane P,
                                                            const double* __restrict__ g_pts, double* __restrict__ g_ori,
                                                            double* __restrict__ g_dir)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        d3 go = mk3(0, 0, 0), gd = mk3(0, 0, 0);
        if (mask3[3 * i]) {
            const d3 o = ld3(out_ori + 3 * i), d = ld3(out_dir + 3 * i);
            double s, dn;
            if (plane_param(P, o, d, s, dn)) {
                const d3 g = ld3(g_pts + 3 * i);
                go = g - mk3(P.nx, P.ny, P.nz) * __ddiv_rn(dot(g, d), dn);
                gd = go * s;
            }
        }
        st3(g_ori + 3 * i, go);
        st3(g_dir + 3 * i, gd);
    }
}

}  // namespace drt
