// common.cuh -- shared device helpers of libdrt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drt {

// ------------------------------------------------------------------------------------------------
// float64 3-vectors with explicitly ROUNDED (never FMA-contracted) arithmetic.
// The differentiable chain and the query-stage triangle test must round exactly like the CPU
// restatement they are checked against (IEEE double, no contraction), so that hit ids, TIR
// decisions and the float32 cast of the secondary-ray origins agree bit for bit.
// ------------------------------------------------------------------------------------------------
struct d3 {
    double x, y, z;
};

__device__ __forceinline__ double mulr(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double addr(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double subr(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ d3 mk3(double x, double y, double z) { return d3{x, y, z}; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return d3{addr(a.x, b.x), addr(a.y, b.y), addr(a.z, b.z)}; }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return d3{subr(a.x, b.x), subr(a.y, b.y), subr(a.z, b.z)}; }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return d3{mulr(a.x, s), mulr(a.y, s), mulr(a.z, s)}; }
__device__ __forceinline__ d3 operator-(d3 a) { return d3{-a.x, -a.y, -a.z}; }
// reference dot (DiffRender.py:23-29): ((x*x + y*y) + z*z)
__device__ __forceinline__ double dot(d3 a, d3 b)
{
    return addr(addr(mulr(a.x, b.x), mulr(a.y, b.y)), mulr(a.z, b.z));
}
__device__ __forceinline__ d3 cross(d3 a, d3 b)
{
    return d3{subr(mulr(a.y, b.z), mulr(a.z, b.y)), subr(mulr(a.z, b.x), mulr(a.x, b.z)),
              subr(mulr(a.x, b.y), mulr(a.y, b.x))};
}
__device__ __forceinline__ d3 divs(d3 a, double s) { return d3{__ddiv_rn(a.x, s), __ddiv_rn(a.y, s), __ddiv_rn(a.z, s)}; }
__device__ __forceinline__ d3 ld3(const double* __restrict__ p) { return d3{p[0], p[1], p[2]}; }
__device__ __forceinline__ void st3(double* __restrict__ p, d3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// ------------------------------------------------------------------------------------------------
// Query-stage triangle test: float64 Moller-Trumbore on float32-rounded data, closed triangle,
// accept (float)t > 0 (DiffRender.py:391).  Decides exactly like oracle/drt_oracle.c:query_tri:
// the accepted path evaluates the identical expression tree (e1 = b-a and e2 = c-a are precomputed
// in float64 at build time with the same rounding), and the early rejections below only fire when
// the reference expression is PROVABLY rejected too:
//   u = fl(U*fl(1/det)) with U = tvec.pvec.  If U and det have opposite signs (|U| > 1e-150, so the
//   product cannot underflow to -0) then u < 0; if |U| > |det|(1+2^-40) then |u| > 1 because the two
//   roundings lose at most 2^-51 relatively.  Same for v, u+v (no cancellation when the signs agree)
//   and the sign of t.  det == 0 gives u = +-inf or NaN in the reference: rejected.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool opposite(double x, double det) { return ((x < 0.0) != (det < 0.0)) && fabs(x) > 1e-150; }

__device__ __forceinline__ bool query_tri(d3 o, d3 d, d3 a, d3 e1, d3 e2, double& t_out)
{
    d3 pvec = cross(d, e2);
    double det = dot(e1, pvec);
    if (det == 0.0) return false;
    d3 tvec = o - a;
    double U = dot(tvec, pvec);
    const double lim = mulr(fabs(det), 1.0000000000009094947017729282379150390625);  // |det| (1 + 2^-40)
    if (opposite(U, det) || fabs(U) > lim) return false;
    d3 qvec = cross(tvec, e1);
    double V = dot(d, qvec);
    if (opposite(V, det) || addr(fabs(U), fabs(V)) > lim) return false;
    double T = dot(e2, qvec);
    if (opposite(T, det)) return false;
    double inv = __ddiv_rn(1.0, det);
    double u = mulr(U, inv);
    if (!(u >= 0.0 && u <= 1.0)) return false;
    double v = mulr(V, inv);
    if (!(v >= 0.0 && addr(u, v) <= 1.0)) return false;
    double t = mulr(T, inv);
    if (!(__double2float_rn(t) > 0.0f)) return false;
    t_out = t;
    return true;
}

// ------------------------------------------------------------------------------------------------
// One surface interaction of the differentiable chain, forward.
//   JIT_Dintersect   DiffRender.py:64-121  (t by Moller-Trumbore, flat face normal :103-104)
//   refract_ray      DiffRender.py:503-535 (orientation, eta swap, TIR via FrDielectric :54-56,
//                    tan-law Refract :37-47, origin advance by t and +1e-5*wt :528-532)
// Everything needed by the reverse pass is kept in the record.
// ------------------------------------------------------------------------------------------------
struct HitRec {
    d3 d;            // incoming direction
    d3 a0, e1, e2;   // triangle
    d3 N, n, np;     // N = e1 x e2, n = shading normal (N/L; smooth mode: the interpolated unit normal), np = oriented normal
    d3 wt, x;        // refracted unit direction, hit point
    double L, t, D;  // |N| (smooth mode: |interpolated normal| before normalisation), distance, d.N
    double sgn, eta, c, cT, A, nw;
    double u, v;     // barycentric coordinates of the hit (smooth mode only; detached like DiffRender.py:108-109)
    bool cT_grad;    // clamp(min=0) in Refract passes gradient
    bool tir;
};

// SMOOTH = false is the reference's live behaviour (flat face normal, DiffRender.py:103-104).
// SMOOTH = true is the OPTIONAL non-parity mode the reference keeps commented out (DiffRender.py:107-114): the shading normal
// is n = normalize((1-u-v) n0 + u n1 + v n2) over the vertex normals vn[0..2] of the hit triangle, with u, v detached.
// Everything else -- t, orientation, TIR, tan-law Refract, the 1e-5 offset -- is the same code.
template <bool SMOOTH>
__device__ __forceinline__ void hit_forward_t(HitRec& h, d3 o, d3 d, d3 a0, d3 a1, d3 a2, const d3* vn, double ext_ior,
                                              double int_ior, d3& o2, d3& d2)
{
    h.d = d; h.a0 = a0;
    h.e1 = a1 - a0; h.e2 = a2 - a0;
    d3 pvec = cross(d, h.e2);
    double det = dot(h.e1, pvec);
    double inv_det = __ddiv_rn(1.0, det);
    d3 tvec = o - a0;
    d3 qvec = cross(tvec, h.e1);
    h.t = mulr(dot(h.e2, qvec), inv_det);
    h.N = cross(h.e1, h.e2);
    if (SMOOTH) {
        h.u = mulr(dot(tvec, pvec), inv_det);  // DiffRender.py:84
        h.v = mulr(dot(d, qvec), inv_det);     // DiffRender.py:87
        const double w0 = subr(subr(1.0, h.u), h.v);
        const d3 nI = (vn[0] * w0 + vn[1] * h.u) + vn[2] * h.v;  // DiffRender.py:113
        h.L = __dsqrt_rn(dot(nI, nI));
        h.n = divs(nI, h.L);                    // DiffRender.py:114
    } else {
        h.L = __dsqrt_rn(dot(h.N, h.N));
        h.n = divs(h.N, h.L);
    }
    h.D = dot(d, h.N);
    d3 wo = -d;
    double c0 = dot(wo, h.n);
    double cc = c0 < -1.0 ? -1.0 : (c0 > 1.0 ? 1.0 : c0);
    bool entering = cc > 0.0;
    double etaI = entering ? ext_ior : int_ior, etaT = entering ? int_ior : ext_ior;
    h.sgn = entering ? 1.0 : -1.0;
    h.np = entering ? h.n : -h.n;
    double cf = entering ? cc : -cc;
    // FrDielectric: only the TIR flag is live (DiffRender.py:526)
    double s = subr(1.0, mulr(cf, cf));
    s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
    double sinI = __dsqrt_rn(s);
    double sinT = __ddiv_rn(mulr(sinI, etaI), etaT);
    h.tir = sinT >= 1.0;
    // Refract (tan-law: cosThetaT is built from sin2ThetaI, DiffRender.py:42)
    h.eta = __ddiv_rn(etaI, etaT);
    h.c = dot(h.np, wo);
    double s2 = subr(1.0, mulr(h.c, h.c));
    h.cT_grad = s2 >= 0.0;
    if (s2 < 0.0) s2 = 0.0;
    double s2c = s2 > 1.0 ? 1.0 : s2;
    h.cT = __dsqrt_rn(subr(1.0, s2c));
    h.A = subr(mulr(h.eta, h.c), h.cT);
    d3 w = (-wo) * h.eta + h.np * h.A;
    h.nw = __dsqrt_rn(dot(w, w));
    h.wt = divs(w, h.nw);
    h.x = o + d * h.t;
    o2 = h.x + h.wt * 1e-5;
    d2 = h.wt;
}

__device__ __forceinline__ void hit_forward(HitRec& h, d3 o, d3 d, d3 a0, d3 a1, d3 a2, double ext_ior,
                                            double int_ior, d3& o2, d3& d2)
{
    hit_forward_t<false>(h, o, d, a0, a1, a2, nullptr, ext_ior, int_ior, o2, d2);
}

// Reverse of hit_forward (analytic Jacobian).  (go2, gd2): gradient w.r.t. the
// outgoing ray.  Adds the gradients of a0,a1,a2 to ga[0..2]; returns gradient w.r.t. the incoming
// ray in (go, gd).  SMOOTH: the shading normal does not depend on the triangle's vertices but on its three vertex
// normals -- their gradients are added to gn[0..2]; the vertices still move the hit point through t.
template <bool SMOOTH>
__device__ __forceinline__ void hit_backward_t(const HitRec& h, d3 go2, d3 gd2, d3 ga[3], d3* gn, d3& go, d3& gd)
{
    d3 g_wt = gd2 + go2 * 1e-5;
    d3 g_x = go2;
    d3 g_o = g_x;
    double g_t = dot(g_x, h.d);
    d3 g_d = g_x * h.t;
    d3 g_w = (g_wt - h.wt * dot(h.wt, g_wt)) * __ddiv_rn(1.0, h.nw);
    g_d = g_d + g_w * h.eta;
    double g_A = dot(g_w, h.np);
    d3 g_np = g_w * h.A;
    double dcT = h.cT_grad ? __ddiv_rn(h.c, h.cT) : 0.0;
    double g_c = mulr(g_A, subr(h.eta, dcT));
    g_d = g_d - h.np * g_c;
    g_np = g_np - h.d * g_c;
    d3 g_n = g_np * h.sgn;
    d3 g_N = (g_n - h.n * dot(h.n, g_n)) * __ddiv_rn(1.0, h.L);  // SMOOTH: gradient of the un-normalised interpolated normal
    if (SMOOTH) {
        gn[0] = gn[0] + g_N * subr(subr(1.0, h.u), h.v);
        gn[1] = gn[1] + g_N * h.u;
        gn[2] = gn[2] + g_N * h.v;
        g_N = mk3(0, 0, 0);  // N = e1 x e2 only enters through t from here on
    }
    double k = __ddiv_rn(g_t, h.D);
    d3 g_a0 = h.N * k;
    g_o = g_o - h.N * k;
    g_N = g_N + (h.a0 - h.x) * k;
    g_d = g_d - h.N * mulr(k, h.t);
    d3 g_e1 = cross(h.e2, g_N);
    d3 g_e2 = cross(g_N, h.e1);
    ga[1] = ga[1] + g_e1;
    ga[2] = ga[2] + g_e2;
    ga[0] = ga[0] + ((g_a0 - g_e1) - g_e2);
    go = g_o; gd = g_d;
}

__device__ __forceinline__ void hit_backward(const HitRec& h, d3 go2, d3 gd2, d3 ga[3], d3& go, d3& gd)
{
    hit_backward_t<false>(h, go2, gd2, ga, nullptr, go, gd);
}

}  // namespace drt
