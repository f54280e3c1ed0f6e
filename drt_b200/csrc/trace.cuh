// trace.cuh -- BVH traversal, fused forward tracer, backward kernel (sm_100a).
#pragma once
#include <cooperative_groups.h>

#include <climits>

#include "bvh.cuh"

namespace drt {

constexpr int kStackDepth = 96 + 8;
// leaves a lane may queue before its triangle tests run (see "Leaves are DEFERRED" below)
#ifndef DRT_DEFER
#define DRT_DEFER 3
#endif
constexpr int kDefer = DRT_DEFER;
static_assert(kDefer >= 1 && kDefer <= 4, "the shared-memory leaf queue has 4 slots");  // binary Karras depth <= 63 key bits + 32 index bits, + deferred leaves
constexpr int kDone = INT_MIN;

struct QRay {        // query ray = float32 cast of the chain's float64 ray (DiffRender.py:387-388)
    float ox, oy, oz, dx, dy, dz;
};

__device__ __forceinline__ QRay cast_ray(d3 o, d3 d)
{
    return QRay{__double2float_rn(o.x), __double2float_rn(o.y), __double2float_rn(o.z),
                __double2float_rn(d.x), __double2float_rn(d.y), __double2float_rn(d.z)};
}

// ---------------------------------------------------------------------------------------------
// Slab test of the child boxes of a node, FMA form:  t = plane * (1/d) - o * (1/d).
//
// Conservativeness (the float64 triangle test, never a box test, must decide every hit):
//   with inv = fl(1/d), c = fl(o*inv) the FMA returns (plane - o)/d * (1+e1)(1+e3) - (o/d)(1+e1)(1+e3) e2,
//   |e| <= 2^-24.  The last term equals moving the plane by |o| 2^-24: covered by the build-time box
//   inflation of pmax 2^-17 for every |o| <= 64 pmax (2x margin); rays starting farther away carry
//   the same bound as an additive slack E (ray_setup).  The relative terms are covered by the
//   factor 1 + 2^-20 on tfar.  A direction component with |d| < 2^-100 (including 0) is given
//   inv = +-2^100: finite, so no inf - inf, and the same error analysis holds -- the ray then
//   passes every box whose (inflated) slab contains its origin on that axis, which is the closed-box
//   answer for an axis-parallel ray.  Empty child slots (lo = +inf, hi = -inf) never pass.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Traversal stacks.  max depth in use on the benchmark meshes: 11 (tools/stepsim), worst case a Karras tree can have: kStackDepth.
//  LocalStack : per-lane array in local memory (one-ray-per-thread kernels, the beam walk).
//  SharedStack: the first kSmemStack entries and the deferred-leaf queue live in shared memory as [slot][thread] -- every
//               push / pop of a warp is ONE conflict-free wavefront whatever the lanes' stack pointers are, while a per-lane
//               local array costs one wavefront per distinct stack pointer in the warp and shares L1 with the node fetches
//               (tools/microbench/pipes.cu on B200: divergent push/pop 3.1x faster); deeper entries spill to local memory.
// ---------------------------------------------------------------------------------------------
#ifndef DRT_SMEM_STACK
#define DRT_SMEM_STACK 0  // measured on B200 at C4: 8 slots in shared memory 5.78 ms forward, 12 slots 5.78, local memory 5.30
#endif
constexpr int kSmemStack = DRT_SMEM_STACK;
constexpr int kQueryBlock = 128;  // threads per block of every persistent query kernel

struct LocalStack {
    int a[kStackDepth];
    __device__ __forceinline__ void push(int& sp, int v) { a[sp++] = v; }
    __device__ __forceinline__ int pop_or(int& sp, int empty) { return sp ? a[--sp] : empty; }
    __device__ __forceinline__ void reset(int& sp) { sp = 0; }
    __device__ __forceinline__ void leaf_put(int j, int v) { a[kStackDepth - 1 - j] = v; }
    __device__ __forceinline__ int leaf_get(int j) const { return a[kStackDepth - 1 - j]; }
};

#if DRT_SMEM_STACK > 0
constexpr int kSmemStackInts = (kSmemStack + 4) * kQueryBlock;  // per block: stack slots + up to 4 queued leaves
// Hot path: push = one STS at sm + sp * 512, pop = one LDS; `sp` counts the entries held in shared memory.  When the kSmemStack
// slots are full they are moved to local memory as a block (deep[0] = number of entries parked there) and brought back when
// the shared part runs empty -- out of line, never taken on the benchmark meshes at kSmemStack >= 12, a few times per
// thousand rays at 8.
struct SharedStack {
    unsigned sm;  // shared-window address of this thread's slot 0 (slot j is at sm + j * 4 * kQueryBlock)
    int deep[kStackDepth + 1];
    static __device__ __forceinline__ void sts(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
    static __device__ __forceinline__ int lds(unsigned addr)
    {
        int v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        return v;
    }
    static __device__ __noinline__ void spill(unsigned sm, int* deep)  // all kSmemStack slots -> local memory
    {
        const int n = deep[0];
        for (int j = 0; j < kSmemStack; ++j) deep[1 + n + j] = lds(sm + (unsigned)j * (4u * kQueryBlock));
        deep[0] = n + kSmemStack;
    }
    static __device__ __noinline__ int refill(unsigned sm, int* deep)  // -> entries brought back (0: nothing parked)
    {
        int n = deep[0];
        if (n == 0) return 0;
        n -= kSmemStack;
        for (int j = 0; j < kSmemStack; ++j) sts(sm + (unsigned)j * (4u * kQueryBlock), deep[1 + n + j]);
        deep[0] = n;
        return kSmemStack;
    }
    __device__ __forceinline__ void push(int& sp, int v)
    {
        if (sp == kSmemStack) { spill(sm, deep); sp = 0; }
        sts(sm + (unsigned)sp * (4u * kQueryBlock), v);
        ++sp;
    }
    __device__ __forceinline__ int pop_or(int& sp, int empty)
    {
        if (sp == 0) {
            sp = refill(sm, deep);
            if (sp == 0) return empty;
        }
        --sp;
        return lds(sm + (unsigned)sp * (4u * kQueryBlock));
    }
    __device__ __forceinline__ void reset(int& sp) { sp = 0; deep[0] = 0; }
    __device__ __forceinline__ void leaf_put(int j, int v) { sts(sm + (unsigned)(kSmemStack + j) * (4u * kQueryBlock), v); }
    __device__ __forceinline__ int leaf_get(int j) const { return lds(sm + (unsigned)(kSmemStack + j) * (4u * kQueryBlock)); }
};
// the address is made opaque (volatile asm) so that it lives in ONE register instead of being rematerialised from
// %tid / the shared window base in every node step (5 instructions per push or pop)
#define DRT_QUERY_STACK(name)                                                                              \
    __shared__ int name##_block[kSmemStackInts];                                                           \
    SharedStack name;                                                                                      \
    asm volatile("mov.u32 %0, %1;" : "=r"(name.sm) : "r"((unsigned)__cvta_generic_to_shared(name##_block + threadIdx.x))); \
    name.deep[0] = 0
typedef SharedStack QueryStack;
#else
#define DRT_QUERY_STACK(name) LocalStack name
typedef LocalStack QueryStack;
#endif

// DRT_FOLD_E: the slack of the box test (relative 2^-20 on tfar + E) folded into the far-plane addends;
// DRT_FAR_SEL: far-plane selectors precomputed; DRT_LDG256: one 256-bit node load (LDG.E.256, sm_100a) instead of two 128-bit ones.
#ifndef DRT_FOLD_E
#define DRT_FOLD_E 0
#endif
#ifndef DRT_FAR_SEL
#define DRT_FAR_SEL 0
#endif
#ifndef DRT_LDG256
#define DRT_LDG256 0
#endif

struct RayQ {
    QRay r;
    float ix, iy, iz;  // 1/d            (quantised nodes: A = s/d, the grid step over the direction)
    float cx, cy, cz;  // o/d            (quantised nodes: C' = (g0 - o)/d - 2^23 A, ADDED by the plane FFMA)
    float E;           // additive slack, 0 for origins within 64*pmax (quantised: within 64 extents of the grid)
#if DRT_QNODE
    unsigned sx, sy, sz;  // PRMT selector of the NEAR plane of a (lo | hi << 16) pair on each axis
#if DRT_FOLD_E
    float fcx, fcy, fcz;  // addends of the FAR planes: C' + slack (see ray_setup), so that a box test is just near <= far
#endif
#if DRT_FAR_SEL
    unsigned fsx, fsy, fsz;  // selectors of the FAR planes, kept in registers (else s ^ 0x22 in every node step)
#endif
#endif
};

#ifndef DRT_PREFETCH_PUSHED
#define DRT_PREFETCH_PUSHED 0
#endif

#if DRT_QNODE
// |d| < 2^-80 (including 0) is traced as +-2^-80: finite everywhere below (A 2^23 <= s 2^103), and such a ray moves
// less than 2^-73 extents along that axis over any distance that matters -- far inside the plane margin.
__device__ __forceinline__ float safe_inv(float d)
{
    return fabsf(d) < 8.271806125530277e-25f ? copysignf(1.2089258196146292e24f, d) : __fdiv_rn(1.f, d);
}

// Quantised planes p = g0 + q s:  t = (p - o)/d = q A + C with A = s/d, C = (g0 - o)/d.  q reaches the FFMA
// without a conversion: PRMT glues the 16-bit q under the exponent of 2^23 (float bits 0x4B00qqqq = 2^23 + q,
// exact), and the constant is folded into the addend, C' = C - 2^23 A.  The same PRMT picks the NEAR or the FAR
// plane of the (lo | hi << 16) pair by the sign of A, so a node step costs 12 PRMT + 12 FFMA + 12 FMNMX -- the
// instruction count of the float layout (12 FFMA + 24 FMNMX) with half the loads.
// Error budget in grid steps, for origins within 64 extents of the grid: rounding of C' <= 0.5 (the same shift
// for every plane of an axis), FMA-form terms |q A| 2^-24 + |C| 2^-23 <= 0.51, quantisation rounding < 0.02
// (bvh.cuh: qpair) -- against 3 steps of outward margin on every stored plane.  Relative terms: the factor
// 1 + 2^-20 on tfar.  Farther origins carry E = 2^-21 max(|C| + 2^23 |A|) >= the sum of both planes' errors.
// Hence a box test never rejects a box that contains a true hit; the float64 triangle test decides every hit.
__device__ __forceinline__ RayQ ray_setup(const BvhView& B, const QRay& r)
{
    RayQ q;
    q.r = r;
    const float inx = safe_inv(r.dx), iny = safe_inv(r.dy), inz = safe_inv(r.dz);
    const float sx = __uint_as_float(__ldg(B.scene + 11)), sy = __uint_as_float(__ldg(B.scene + 12)), sz = __uint_as_float(__ldg(B.scene + 13));
    const float wx = __uint_as_float(__ldg(B.scene + 8)) - r.ox, wy = __uint_as_float(__ldg(B.scene + 9)) - r.oy,
                wz = __uint_as_float(__ldg(B.scene + 10)) - r.oz;
    q.ix = sx * inx; q.iy = sy * iny; q.iz = sz * inz;
    const float cx = wx * inx, cy = wy * iny, cz = wz * inz;
    const float mx = 8388608.f * q.ix, my = 8388608.f * q.iy, mz = 8388608.f * q.iz;
    q.cx = cx - mx; q.cy = cy - my; q.cz = cz - mz;
    q.sx = inx >= 0.f ? 0x7410u : 0x7432u;
    q.sy = iny >= 0.f ? 0x7410u : 0x7432u;
    q.sz = inz >= 0.f ? 0x7410u : 0x7432u;
    q.E = 0.f;
    const float k = 64.f * 65520.f;
    if (!(fabsf(wx) <= k * sx && fabsf(wy) <= k * sy && fabsf(wz) <= k * sz))
        q.E = 4.76837158203125e-07f * fmaxf(fabsf(cx) + fabsf(mx), fmaxf(fabsf(cy) + fabsf(my), fabsf(cz) + fabsf(mz)));
#if DRT_FOLD_E
    // Far planes carry their slack in the addend: per axis 2^-21 (|C| + 65536 |A|) >= 4 x the rounding of any plane distance
    // on that axis (|t| <= |C| + 65536 |A| inside the grid; the FMA rounds the RESULT once, 2^-24 |t|, on the near and on the
    // far side), plus the far-origin slack E.  The sum is rounded UP, so folding only ever widens a box; the rounding of the
    // folded addend itself (<= 0.5 grid step, upward) stays inside the 3 steps of outward margin of the stored planes.
    q.fcx = __fadd_ru(q.cx, fmaf(4.76837158203125e-07f, fabsf(cx) + 65536.f * fabsf(q.ix), q.E));
    q.fcy = __fadd_ru(q.cy, fmaf(4.76837158203125e-07f, fabsf(cy) + 65536.f * fabsf(q.iy), q.E));
    q.fcz = __fadd_ru(q.cz, fmaf(4.76837158203125e-07f, fabsf(cz) + 65536.f * fabsf(q.iz), q.E));
#endif
#if DRT_FAR_SEL
    asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(q.fsx) : "r"(q.sx));  // opaque: not rematerialised inside the loop
    asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(q.fsy) : "r"(q.sy));
    asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(q.fsz) : "r"(q.sz));
#endif
    return q;
}

// closest-hit bound as the box test uses it: with the folded slack the test is near <= min(far planes, tmax), so the
// bound itself carries the slack (a tie at exactly t_best in another box must still pass)
__device__ __forceinline__ float tmax_of(const RayQ& q, double t)
{
#if DRT_QNODE && DRT_FOLD_E
    return fmaf(__double2float_ru(t), 1.00000095367431640625f, q.E);
#else
    return __double2float_ru(t);
#endif
}

// raw prmt.b32: __byte_perm would first mask the selector with 0x7777 (one more ALU op per axis and step)
__device__ __forceinline__ float qplane(unsigned w, unsigned sel)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(sel));
    return __uint_as_float(r);
}

// two FP32 FMAs in one issue slot (FFMA2, sm_100a): same IEEE fma per half, so results are bit-identical to two FFMAs.
// On B200 FFMA issues every clock, PRMT / FMNMX3 / LOP3 / IMAD every other clock (tools/microbench/pipes.cu): a node step
// is bound by issue slots and the ALU pipe, and pairing the 12 plane FMAs frees 6 issue slots per step.
#ifndef DRT_FFMA2
#define DRT_FFMA2 1
#endif
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
#if DRT_FFMA2
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// one binary node: test both children, continue with the nearer hit, push the other.
// DRT_INLINE_LEAF: a hit child that is a LEAF goes straight into the lane's leaf queue when there is room, instead of being
// returned as the next "node" and costing a loop iteration of its own -- ncu (r02a, exit query) shows 15 % of the warp
// instructions in those leaf-push iterations running with 3-4 of 32 lanes while the other lanes wait.
#ifndef DRT_INLINE_LEAF
#define DRT_INLINE_LEAF 0  // measured on B200 at C4: forward 5.49 ms with it, 5.18 without (the extra predicated work in EVERY node step costs more than the leaf iterations it removes)
#endif
// DRT_PREFETCH_TRI: when a leaf is queued, pull its 80-byte triangle record towards L1 -- every queued leaf IS tested a few
// node steps later, and ncu (r02a) shows the first float64 instructions of the triangle test waiting on those loads
// (7.5 % of the exit query's stall samples).
#ifndef DRT_PREFETCH_TRI
#define DRT_PREFETCH_TRI 0
#endif
__device__ __forceinline__ void prefetch_tri(const BvhView& B, int leaf)
{
#if DRT_PREFETCH_TRI
    const char* p = reinterpret_cast<const char*>(B.tris + (size_t)(~leaf) * kTriD2);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 64));
#else
    (void)B; (void)leaf;
#endif
}
template <class S>
__device__ __forceinline__ int node_step(const BvhView& B, const RayQ& q, float tmax, int node, S& stack, int& sp, int& nd)
{
    const uint4* p = B.nodes + (size_t)node * kNodeQuads;
#if DRT_LDG256
    uint4 a, b;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
#else
    const uint4 a = __ldg(p), b = __ldg(p + 1);
#endif
#if DRT_FAR_SEL
    const unsigned fx = q.fsx, fy = q.fsy, fz = q.fsz;
#else
    const unsigned fx = q.sx ^ 0x22u, fy = q.sy ^ 0x22u, fz = q.sz ^ 0x22u;  // selectors of the FAR planes
#endif
    const float2 Axy = make_float2(q.ix, q.iy), Cxy = make_float2(q.cx, q.cy), Azz = make_float2(q.iz, q.iz);
#if DRT_FOLD_E
    const float2 Fxy = make_float2(q.fcx, q.fcy), Czz = make_float2(q.cz, q.fcz);
#else
    const float2 Fxy = Cxy, Czz = make_float2(q.cz, q.cz);
#endif
    const float2 n0 = fma2(make_float2(qplane(a.x, q.sx), qplane(a.z, q.sy)), Axy, Cxy);
    const float2 f0 = fma2(make_float2(qplane(a.x, fx), qplane(a.z, fy)), Axy, Fxy);
    const float2 n1 = fma2(make_float2(qplane(a.y, q.sx), qplane(a.w, q.sy)), Axy, Cxy);
    const float2 f1 = fma2(make_float2(qplane(a.y, fx), qplane(a.w, fy)), Axy, Fxy);
    const float2 z0 = fma2(make_float2(qplane(b.x, q.sz), qplane(b.x, fz)), Azz, Czz);  // (near, far) of child 0 on z
    const float2 z1 = fma2(make_float2(qplane(b.y, q.sz), qplane(b.y, fz)), Azz, Czz);
    const float N0 = fmaxf(fmaxf(n0.x, n0.y), fmaxf(z0.x, 0.f));
    const float F0 = fminf(fminf(f0.x, f0.y), fminf(z0.y, tmax));
    const float N1 = fmaxf(fmaxf(n1.x, n1.y), fmaxf(z1.x, 0.f));
    const float F1 = fminf(fminf(f1.x, f1.y), fminf(z1.y, tmax));
#if DRT_FOLD_E
    bool h0 = N0 <= F0, h1 = N1 <= F1;
#else
    bool h0 = N0 <= fmaf(F0, 1.00000095367431640625f, q.E);
    bool h1 = N1 <= fmaf(F1, 1.00000095367431640625f, q.E);
#endif
    const int c0 = (int)b.z, c1 = (int)b.w;
#if DRT_INLINE_LEAF
    if (h0 && c0 < 0 && nd < kDefer) { stack.leaf_put(nd, c0); ++nd; h0 = false; prefetch_tri(B, c0); }
    if (h1 && c1 < 0 && nd < kDefer) { stack.leaf_put(nd, c1); ++nd; h1 = false; prefetch_tri(B, c1); }
#endif
    if (h0 && h1) {
        const bool first0 = N0 <= N1;
        const int later = first0 ? c1 : c0;
        stack.push(sp, later);
#if DRT_PREFETCH_PUSHED
        // experiment switch (off): pull the postponed child's node towards L1 now, so that popping it later is not an L2 round trip
        if (later >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(B.nodes + (size_t)later * kNodeQuads));
#endif
        return first0 ? c0 : c1;
    }
    if (h0) return c0;
    if (h1) return c1;
    return stack.pop_or(sp, kDone);
}
#if DRT_BVH4
// one WIDE node (BvhView::nodes4): test the four grandchild boxes, continue with the nearest hit of the nearer pair, push the
// others so that they pop in the order the binary walk would visit them (nearer pair first, nearer member first)
template <class S>
__device__ __forceinline__ int node_step4(const BvhView& B, const RayQ& q, float tmax, int node, S& stack, int& sp, int& nd)
{
    const uint4* p = B.nodes4 + (size_t)node * 4;
    const uint4 X = __ldg(p), Y = __ldg(p + 1), Z = __ldg(p + 2), L = __ldg(p + 3);
    const unsigned fx = q.sx ^ 0x22u, fy = q.sy ^ 0x22u, fz = q.sz ^ 0x22u;
    const float2 Axy = make_float2(q.ix, q.iy), Cxy = make_float2(q.cx, q.cy), Azz = make_float2(q.iz, q.iz), Czz = make_float2(q.cz, q.cz);
    float d[4];
    bool h[4];
    const unsigned xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const float2 n = fma2(make_float2(qplane(xs[s], q.sx), qplane(ys[s], q.sy)), Axy, Cxy);
        const float2 f = fma2(make_float2(qplane(xs[s], fx), qplane(ys[s], fy)), Axy, Cxy);
        const float2 z = fma2(make_float2(qplane(zs[s], q.sz), qplane(zs[s], fz)), Azz, Czz);
        const float N = fmaxf(fmaxf(n.x, n.y), fmaxf(z.x, 0.f));
        const float F = fminf(fminf(f.x, f.y), fminf(z.y, tmax));
        h[s] = N <= fmaf(F, 1.00000095367431640625f, q.E);
        d[s] = h[s] ? N : INFINITY;
    }
#if DRT_INLINE_LEAF
    {
        const int ls[4] = {(int)L.x, (int)L.y, (int)L.z, (int)L.w};
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (h[s] && ls[s] < 0 && nd < kDefer) { stack.leaf_put(nd, ls[s]); ++nd; h[s] = false; d[s] = INFINITY; prefetch_tri(B, ls[s]); }
    }
#endif
    const bool s01 = d[1] < d[0], s23 = d[3] < d[2];
    const float m01 = fminf(d[0], d[1]), m23 = fminf(d[2], d[3]);
    const int near01 = s01 ? (int)L.y : (int)L.x, far01 = s01 ? (int)L.x : (int)L.y;
    const int near23 = s23 ? (int)L.w : (int)L.z, far23 = s23 ? (int)L.z : (int)L.w;
    const bool both01 = h[0] && h[1], any01 = h[0] || h[1], both23 = h[2] && h[3], any23 = h[2] || h[3];
    const bool swap = m23 < m01;  // the right pair is entered first
    const int nearP = swap ? near23 : near01, farP = swap ? far23 : far01, nearQ = swap ? near01 : near23, farQ = swap ? far01 : far23;
    const bool bothP = swap ? both23 : both01, anyP = swap ? any23 : any01, bothQ = swap ? both01 : both23, anyQ = swap ? any01 : any23;
    if (bothQ) stack.push(sp, farQ);
    if (anyQ) stack.push(sp, nearQ);
    if (bothP) stack.push(sp, farP);
    if (anyP) return nearP;
    return stack.pop_or(sp, kDone);
}
#endif

#else
__device__ __forceinline__ float safe_inv(float d)
{
    return fabsf(d) < 7.888609052210118e-31f ? copysignf(1.2676506002282294e30f, d) : __fdiv_rn(1.f, d);
}

__device__ __forceinline__ RayQ ray_setup(const BvhView& B, const QRay& r)
{
    RayQ q;
    q.r = r;
    q.ix = safe_inv(r.dx); q.iy = safe_inv(r.dy); q.iz = safe_inv(r.dz);
    q.cx = r.ox * q.ix; q.cy = r.oy * q.iy; q.cz = r.oz * q.iz;
    const float pmax = __uint_as_float(__ldg(B.scene + 7));
    const float far = fmaxf(fabsf(r.ox), fmaxf(fabsf(r.oy), fabsf(r.oz)));
    q.E = 0.f;
    if (!(far <= 64.f * pmax))
        q.E = 2.384185791015625e-07f * fmaxf(fabsf(q.cx), fmaxf(fabsf(q.cy), fabsf(q.cz)));  // 2^-22 |o/d|
    return q;
}

// one binary node: test both children, continue with the nearer hit, push the other
template <class S>
__device__ __forceinline__ int node_step(const BvhView& B, const RayQ& q, float tmax, int node, S& stack, int& sp, int& nd)
{
    (void)nd;
    const float4* p = B.nodes + (size_t)node * kNodeQuads;
    const float4 nx = __ldg(p), ny = __ldg(p + 1), nz = __ldg(p + 2);
    const float4 nl = __ldg(p + 3);
    // child 0: planes .x/.y, child 1: planes .z/.w
    const float ax0 = fmaf(nx.x, q.ix, -q.cx), bx0 = fmaf(nx.y, q.ix, -q.cx), ax1 = fmaf(nx.z, q.ix, -q.cx), bx1 = fmaf(nx.w, q.ix, -q.cx);
    const float ay0 = fmaf(ny.x, q.iy, -q.cy), by0 = fmaf(ny.y, q.iy, -q.cy), ay1 = fmaf(ny.z, q.iy, -q.cy), by1 = fmaf(ny.w, q.iy, -q.cy);
    const float az0 = fmaf(nz.x, q.iz, -q.cz), bz0 = fmaf(nz.y, q.iz, -q.cz), az1 = fmaf(nz.z, q.iz, -q.cz), bz1 = fmaf(nz.w, q.iz, -q.cz);
    const float n0 = fmaxf(fmaxf(fminf(ax0, bx0), fminf(ay0, by0)), fmaxf(fminf(az0, bz0), 0.f));
    const float f0 = fminf(fminf(fmaxf(ax0, bx0), fmaxf(ay0, by0)), fminf(fmaxf(az0, bz0), tmax));
    const float n1 = fmaxf(fmaxf(fminf(ax1, bx1), fminf(ay1, by1)), fmaxf(fminf(az1, bz1), 0.f));
    const float f1 = fminf(fminf(fmaxf(ax1, bx1), fmaxf(ay1, by1)), fminf(fmaxf(az1, bz1), tmax));
    const bool h0 = n0 <= fmaf(f0, 1.00000095367431640625f, q.E);
    const bool h1 = n1 <= fmaf(f1, 1.00000095367431640625f, q.E);
    const int c0 = __float_as_int(nl.x), c1 = __float_as_int(nl.y);
    if (h0 && h1) {
        const bool first0 = n0 <= n1;
        stack.push(sp, first0 ? c1 : c0);
        return first0 ? c0 : c1;
    }
    if (h0) return c0;
    if (h1) return c1;
    return stack.pop_or(sp, kDone);
}
#endif

#if DRT_QNODE
// ---------------------------------------------------------------------------------------------
// BEAM = the 32 primary rays of one pixel tile, which share their origin (a pinhole view has ONE camera centre,
// captured_data.py:38): {o + t d : t >= 0, d in [dmin, dmax]} with per-axis direction intervals over the tile's float32 query
// rays.  One conservative test of the beam against a box answers "can ANY ray of the tile hit it":
//   exists t >= 0 with, on every axis,  t dmax >= lo - o  and  t dmin <= hi - o          (each linear in t)
// which for an axis where every ray points the same way is the ray's own slab test with the NEAR plane divided by the
// largest |d| and the FAR plane by the smallest -- the same one-FFMA-per-plane form on the quantised planes (node_step), so
// the same rounding budget applies (3 grid steps of outward margin on every stored plane against <= 1.03 steps of rounding;
// fl(1/d) is monotonic, so no ray's own bound is ever tighter than the beam's).  An axis whose interval contains 0 (or
// |d| < 2^-80) is dropped (no constraint): looser, still conservative.  Box tests only CULL: every hit is still decided by the
// float64 triangle test of the per-ray traversal, so hit ids stay bit-identical.
// ---------------------------------------------------------------------------------------------
struct BeamQ {
    float an[3], cn[3], af[3], cf[3];  // near / far plane: t = q' a + c
    unsigned sn[3];                    // PRMT selector of the near plane per axis (far = sn ^ 0x22)
    float E;
};

__device__ __forceinline__ BeamQ beam_setup(const BvhView& B, float ox, float oy, float oz, const float dmin[3], const float dmax[3])
{
    BeamQ q;
    const float o[3] = {ox, oy, oz};
    const float tiny = 8.271806125530277e-25f;  // 2^-80
    float emax = 0.f;
    bool far_origin = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float s = __uint_as_float(__ldg(B.scene + 11 + k));
        const float w = __uint_as_float(__ldg(B.scene + 8 + k)) - o[k];
        if (!(fabsf(w) <= 64.f * 65520.f * s)) far_origin = true;
        float in, if_;
        if (dmin[k] > tiny) { in = __fdiv_rn(1.f, dmax[k]); if_ = __fdiv_rn(1.f, dmin[k]); q.sn[k] = 0x7410u; }
        else if (dmax[k] < -tiny) { in = __fdiv_rn(1.f, dmin[k]); if_ = __fdiv_rn(1.f, dmax[k]); q.sn[k] = 0x7432u; }
        else { q.an[k] = 0.f; q.cn[k] = -INFINITY; q.af[k] = 0.f; q.cf[k] = INFINITY; q.sn[k] = 0x7410u; continue; }
        q.an[k] = s * in;
        q.af[k] = s * if_;
        const float c_n = w * in, c_f = w * if_;
        const float m_n = 8388608.f * q.an[k], m_f = 8388608.f * q.af[k];
        q.cn[k] = c_n - m_n;
        q.cf[k] = c_f - m_f;
        emax = fmaxf(emax, fmaxf(fabsf(c_n) + fabsf(m_n), fabsf(c_f) + fabsf(m_f)));
    }
    q.E = far_origin ? 4.76837158203125e-07f * emax : 0.f;
    return q;
}

// Walks the tree with the beam.  Returns false when the beam touches no leaf box (every ray of the tile misses the mesh);
// otherwise (or when still undecided after max_steps node steps) true and `entry` = the node below which every possible hit
// of the tile lies: the first node at which the beam enters BOTH children (all siblings passed on the way down were missed by the whole beam), so the per-ray traversals may
// start there instead of at the root.
__device__ __forceinline__ bool beam_walk(const BvhView& B, const BeamQ& q, int* stack, int& entry, int max_steps)
{
    int sp = 0, node = 0;
    bool forked = false;
    entry = 0;
    const unsigned fx = q.sn[0] ^ 0x22u, fy = q.sn[1] ^ 0x22u, fz = q.sn[2] ^ 0x22u;
    for (int step = 0;; ++step) {
        if (step >= max_steps) {  // undecided: keep the tile (entry = the fork, or the end of the single-child chain)
            if (!forked) entry = node;
            return true;
        }
        const uint4* p = B.nodes + (size_t)node * kNodeQuads;
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        const float n0 = fmaxf(fmaxf(fmaf(qplane(a.x, q.sn[0]), q.an[0], q.cn[0]), fmaf(qplane(a.z, q.sn[1]), q.an[1], q.cn[1])),
                               fmaxf(fmaf(qplane(b.x, q.sn[2]), q.an[2], q.cn[2]), 0.f));
        const float f0 = fminf(fminf(fmaf(qplane(a.x, fx), q.af[0], q.cf[0]), fmaf(qplane(a.z, fy), q.af[1], q.cf[1])),
                               fmaf(qplane(b.x, fz), q.af[2], q.cf[2]));
        const float n1 = fmaxf(fmaxf(fmaf(qplane(a.y, q.sn[0]), q.an[0], q.cn[0]), fmaf(qplane(a.w, q.sn[1]), q.an[1], q.cn[1])),
                               fmaxf(fmaf(qplane(b.y, q.sn[2]), q.an[2], q.cn[2]), 0.f));
        const float f1 = fminf(fminf(fmaf(qplane(a.y, fx), q.af[0], q.cf[0]), fmaf(qplane(a.w, fy), q.af[1], q.cf[1])),
                               fmaf(qplane(b.y, fz), q.af[2], q.cf[2]));
        const bool h0 = n0 <= fmaf(f0, 1.00000095367431640625f, q.E);
        const bool h1 = n1 <= fmaf(f1, 1.00000095367431640625f, q.E);
        const int c0 = (int)b.z, c1 = (int)b.w;
        if ((h0 && c0 < 0) || (h1 && c1 < 0)) {  // a leaf box is touched: the tile may hit
            if (!forked) entry = node;
            return true;
        }
        if (h0 && h1) {
            if (!forked) { forked = true; entry = node; }
            const bool first0 = n0 <= n1;
            stack[sp++] = first0 ? c1 : c0;
            node = first0 ? c0 : c1;
        } else if (h0) node = c0;
        else if (h1) node = c1;
        else if (sp) node = stack[--sp];
        else return false;
    }
}

__device__ __forceinline__ float warp_min_f32(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));  // CREDUX.MIN.F32 (sm_100a)
    return r;
}
__device__ __forceinline__ float warp_max_f32(float v)
{
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
#endif  // DRT_QNODE

// exact test of the one triangle of a leaf; updates the closest hit (ties -> lowest id)
__device__ __forceinline__ bool leaf_step(const BvhView& B, const RayQ& q, int leaf, double& t_best, int& id_best, float& tmax)
{
    const QRay& r = q.r;
    const double2* p = B.tris + (size_t)(~leaf) * kTriD2;
    const double2 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    double t;
    if (!query_tri(mk3((double)r.ox, (double)r.oy, (double)r.oz), mk3((double)r.dx, (double)r.dy, (double)r.dz),
                   mk3(w0.x, w0.y, w1.x), mk3(w1.y, w2.x, w2.y), mk3(w3.x, w3.y, w4.x), t))
        return false;
    const int id = (int)__double_as_longlong(w4.y);
    if (t < t_best || (t == t_best && id < id_best)) {
        t_best = t;
        id_best = id;
        tmax = tmax_of(q, t);
    }
    return true;
}

// Leaves are DEFERRED: a lane that reaches a leaf parks it in a small per-lane queue (kept at the
// top end of its stack array) and keeps walking internal nodes; the exact triangle tests run in
// batches once kDefer leaves are queued or the walk is over.  Measured motivation on B200 with test-at-once
// traversal: a warp entered the (long, float64) leaf phase every ~1.5 node steps with ~2 of 32
// lanes active -- every lane parked at a leaf stalls until the slowest lane reaches one.  Deferral
// costs some closest-hit pruning (tmax is updated a few nodes later), never correctness: every leaf
// whose box passed the test is still tested, and ties still resolve to the lowest id.
// Depth: with every lane filling its own queue (walk) 2 was best (1: -9 %, 4: -1 %, 8: -6 % at C4); with the
// warp vote (walk_vote) 3 is (C4 forward 6.43 ms at 3 vs 6.55 at 2 and 6.47 at 4, vote 8).

// DRT_FULLQ_WALK = 1: a lane whose leaf queue is full keeps walking internal nodes and blocks only when the next LEAF arrives
// (one test on the hot path); 0: it blocks as soon as the queue is full, so its triangles are tested -- and tmax shrinks --
// a few node steps earlier.
#ifndef DRT_FULLQ_WALK
#define DRT_FULLQ_WALK 1  // measured at C4 (vote every 3 steps): 5.16 ms forward with 1, 5.19 with 0
#endif
__device__ __forceinline__ bool can_step(int node, int nd)
{
#if DRT_FULLQ_WALK
    return node >= 0 || (node != kDone && nd < kDefer);
#else
    return node != kDone && nd < kDefer;
#endif
}

#ifndef DRT_LEAF_FALL
#define DRT_LEAF_FALL 0  // measured on B200 at C4: forward 5.22 ms with it, 4.98 without
#endif
// one node step or one leaf push (the caller checked can_step)
template <bool WIDE, class S>
__device__ __forceinline__ void advance(const BvhView& B, const RayQ& q, float tmax, int& node, S& stack, int& sp, int& nd)
{
#if DRT_LEAF_FALL
    // a leaf is queued and the popped node is stepped in the SAME iteration: the queueing is a short divergent prefix instead
    // of a whole loop iteration in which the lane skips the node step the rest of the warp executes
    if (node < 0) {
        stack.leaf_put(nd, node);
        prefetch_tri(B, node);
        ++nd;
        node = stack.pop_or(sp, kDone);
    }
    if (node >= 0) {
#if DRT_QNODE && DRT_BVH4
        node = WIDE ? node_step4(B, q, tmax, node, stack, sp, nd) : node_step(B, q, tmax, node, stack, sp, nd);
#else
        node = node_step(B, q, tmax, node, stack, sp, nd);
#endif
    }
#else
    if (node >= 0) {
#if DRT_QNODE && DRT_BVH4
        node = WIDE ? node_step4(B, q, tmax, node, stack, sp, nd) : node_step(B, q, tmax, node, stack, sp, nd);
#else
        node = node_step(B, q, tmax, node, stack, sp, nd);
#endif
    } else {
        stack.leaf_put(nd, node);
        prefetch_tri(B, node);
        ++nd;
        node = stack.pop_or(sp, kDone);
    }
#endif
}

// walk until the stack is exhausted or the lane is blocked by its full leaf queue
template <bool WIDE = false, class S>
__device__ __forceinline__ void walk(const BvhView& B, const RayQ& q, float tmax, int& node, S& stack, int& sp, int& nd)
{
    while (can_step(node, nd)) advance<WIDE>(B, q, tmax, node, stack, sp, nd);
}

// Same walk with a WARP VOTE on when to stop: lanes step together, kVoteEvery node steps (or leaf pushes) per vote, and the
// warp leaves the loop as soon as `vote` lanes are blocked -- leaf queue full, or walk finished with leaves still queued --
// instead of waiting until every lane is (walk() above is the vote = 32 case).  The blocked lanes then get their triangle
// tests while the others still have a short queue; with the per-lane loop a lane that has filled its queue idles until the
// SLOWEST lane of the warp has filled its own.  All 32 lanes must call this.
// Votes cost ~12 of the ~70 instructions of an iteration: measured at C4, one vote per 1 / 2 / 3 / 4 steps: 5.56 / 5.29 / 5.19 / 5.21 ms forward.
// Re-measured on the final kernels of round 2 (beam culling, prepared tile beams, four lanes; C4 step in ms), steps per vote x vote
// threshold: 3x4 5.53 | 4x4 5.48 | 6x2 5.36 | 6x1 5.32 | 8x3 5.37 | 8x2 5.32 | 8x1 5.27 | 9x1 5.25 | 10x1 5.27 | 12x1 5.35 | 16x2 5.49
// -- the warp walks in bursts of 8 node steps and leaves the loop as soon as ONE lane is blocked at a vote; C3 7.48 -> 7.12,
// 9 views 1.054 -> 1.005, C5 (32 views) 14.54 -> 13.99.
#ifndef DRT_VOTE_EVERY
#define DRT_VOTE_EVERY 8
#endif
constexpr int kVoteEvery = DRT_VOTE_EVERY;
#ifndef DRT_VOTE_EVERY4
#define DRT_VOTE_EVERY4 2  // wide steps per vote
#endif
constexpr int kVoteEvery4 = DRT_VOTE_EVERY4;

template <bool WIDE = false, class S>
__device__ __forceinline__ void walk_vote(const BvhView& B, const RayQ& q, float tmax, int& node, S& stack, int& sp, int& nd, int vote)
{
    const unsigned FULL = 0xffffffffu;
    for (;;) {
#pragma unroll
        for (int u = 0; u < (WIDE ? kVoteEvery4 : kVoteEvery); ++u)
            if (can_step(node, nd)) advance<WIDE>(B, q, tmax, node, stack, sp, nd);
        const bool cont = can_step(node, nd);
        if (!__any_sync(FULL, cont) || __popc(__ballot_sync(FULL, nd > 0 && !cont)) >= vote) break;
    }
}

// test the queued leaves; returns true as soon as ANY is satisfied
template <bool ANY, class S>
__device__ __forceinline__ bool drain(const BvhView& B, const RayQ& q, S& stack, int& nd, double& t_best, int& id_best, float& tmax)
{
    while (nd > 0) {
        --nd;
        bool hit = leaf_step(B, q, stack.leaf_get(nd), t_best, id_best, tmax);
        if (ANY && hit) { nd = 0; return true; }
    }
    return false;
}

// DRT_COOP_DRAIN = 1: the queued leaves of a WARP are tested cooperatively.  drain() runs as many rounds as the longest
// queue of the warp, with fewer lanes in each (ncu, r02: the float64 triangle tests are 20 % of the exit query and run at
// 11 -> 3 active lanes).  Here every queued (ray, leaf) pair of the warp is one task; the tasks are numbered level by level
// (first leaf of every lane, then second, ...) through a per-warp table in shared memory, lane w tests task w with the ray
// fetched from its owner by shuffles, and the owner folds the results of its own tasks back into (t_best, id_best) -- the same
// lexicographic minimum over the same float64 tests, so hit ids and distances are unchanged; a warp with 11 + 7 + 3 queued
// leaves does ONE round of 21 lanes instead of three rounds of 11, 7 and 3.  All 32 lanes must call it (converged).
#ifndef DRT_COOP_DRAIN
#define DRT_COOP_DRAIN 0  // measured on B200 at C4: forward 5.93 ms with it, 4.98 without -- drains are frequent and short (most lanes hold one leaf), the table, the six shuffles and three warp barriers cost more than the rounds they save
#endif
struct CoopScratch {
    double t[32];
    int task[32];
    int id[32];
};

template <bool ANY, class S>
__device__ __forceinline__ bool drain_coop(const BvhView& B, const RayQ& q, S& stack, int& nd, double& t_best, int& id_best, float& tmax)
{
    const unsigned FULL = 0xffffffffu;
    unsigned m[kDefer];
#pragma unroll
    for (int j = 0; j < kDefer; ++j) m[j] = __ballot_sync(FULL, nd > j);
    if (m[0] == 0u) return false;
    const bool own = kDefer < 2 || m[1] == 0u;  // at most one leaf per lane: every lane tests its own, no table
    __shared__ CoopScratch scratch_all[kQueryBlock / 32];
    CoopScratch& sc = scratch_all[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    int k[kDefer], leaf[kDefer];  // task index of this lane's j-th leaf (-1: none)
    int total = 0;
#pragma unroll
    for (int j = 0; j < kDefer; ++j) {
        k[j] = nd > j ? total + __popc(m[j] & lt) : -1;
        leaf[j] = nd > j ? stack.leaf_get(j) : -1;
        total += __popc(m[j]);
    }
    bool found = false;
    double tb = t_best;
    int ib = id_best;
    for (int r = 0; r < (own ? 1 : total); r += 32) {
        bool work = nd > 0;
        int tri = ~leaf[0];
        QRay ray = q.r;
        if (!own) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kDefer; ++j)
                if (k[j] >= r && k[j] < r + 32) sc.task[k[j] - r] = ((~leaf[j]) << 5) | (int)lane;
            __syncwarp();
            work = (int)lane < total - r;
            const int tk = work ? sc.task[lane] : (int)lane;
            const int owner = tk & 31;
            tri = tk >> 5;
            ray.ox = __shfl_sync(FULL, q.r.ox, owner); ray.oy = __shfl_sync(FULL, q.r.oy, owner); ray.oz = __shfl_sync(FULL, q.r.oz, owner);
            ray.dx = __shfl_sync(FULL, q.r.dx, owner); ray.dy = __shfl_sync(FULL, q.r.dy, owner); ray.dz = __shfl_sync(FULL, q.r.dz, owner);
        }
        double t = 0.0;
        int id = -1;
        if (work) {
            const double2* p = B.tris + (size_t)tri * kTriD2;
            const double2 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
            if (query_tri(mk3((double)ray.ox, (double)ray.oy, (double)ray.oz), mk3((double)ray.dx, (double)ray.dy, (double)ray.dz),
                          mk3(w0.x, w0.y, w1.x), mk3(w1.y, w2.x, w2.y), mk3(w3.x, w3.y, w4.x), t))
                id = (int)__double_as_longlong(w4.y);
        }
        if (own) {
            if (id >= 0) {
                found = true;
                if (t < tb || (t == tb && id < ib)) { tb = t; ib = id; }
            }
        } else {
            if (work) { sc.t[lane] = t; sc.id[lane] = id; }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kDefer; ++j)
                if (k[j] >= r && k[j] < r + 32) {
                    const int id_j = sc.id[k[j] - r];
                    if (id_j >= 0) {
                        const double t_j = sc.t[k[j] - r];
                        found = true;
                        if (t_j < tb || (t_j == tb && id_j < ib)) { tb = t_j; ib = id_j; }
                    }
                }
        }
    }
    nd = 0;
    if (ib != id_best || tb != t_best) {
        t_best = tb;
        id_best = ib;
        tmax = tmax_of(q, tb);
    }
    return ANY && found;
}

// Exact closest hit (ANY = false) or first hit found (ANY = true; only hit/no-hit is meaningful,
// which is all the reference uses of its third query: DiffRender.py:426-427).
// Returns triangle id (-1 on miss) and the float64 distance along the float32 ray.
template <bool ANY>
__device__ __forceinline__ void traverse(const BvhView& B, const QRay& r, double& t_best, int& id_best)
{
    t_best = INFINITY;
    id_best = -1;
    if (B.nTris <= 0) return;
    const RayQ q = ray_setup(B, r);
    float tmax = INFINITY;
    LocalStack stack;
    int sp = 0, nd = 0;
    int node = 0;
    for (;;) {
        walk(B, q, tmax, node, stack, sp, nd);
        if (drain<ANY>(B, q, stack, nd, t_best, id_best, tmax)) return;
        if (node == kDone) return;
    }
}

// Work item -> ray for whole scanline-ordered images (captured_data.py:26-31): 32 consecutive work items are a
// 4 x 8 (or 8 x 4) pixel TILE instead of a 32 x 1 strip when the image size is known (img_w = 0: identity).  A tile
// is either inside or outside the silhouette far more often than a strip and its rays share more node fetches:
// C4 forward 7.43 ms with strips, 6.27 with 8 x 4, 6.17 with 4 x 8; the lists of the later stages inherit the order.
struct TileMap {
    int img_w, img_hw;  // image width and pixels per image
    int tw_log2;        // log2 of the tile width: 3 = 8 x 4 pixels, 2 = 4 x 8
    __device__ __forceinline__ int ray_of(int item) const
    {
        if (!img_w) return item;
        const int v = item / img_hw, r = item - v * img_hw;
        const int t = r >> 5, w = r & 31, tpr = img_w >> tw_log2;
        const int ty = t / tpr, tx = t - ty * tpr;
        return v * img_hw + (ty * (32 >> tw_log2) + (w >> tw_log2)) * img_w + (tx << tw_log2) + (w & ((1 << tw_log2) - 1));
    }
};

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

// the three vertex normals of triangle id (optional smooth-normal mode, common.cuh: hit_forward_t<true>)
__device__ __forceinline__ void load_vn(const BvhView& B, const double* __restrict__ VN, int id, d3 vn[3])
{
    int i0 = __ldg(&B.F[3 * (size_t)id]), i1 = __ldg(&B.F[3 * (size_t)id + 1]), i2 = __ldg(&B.F[3 * (size_t)id + 2]);
    vn[0] = ld3(VN + 3 * (size_t)i0);
    vn[1] = ld3(VN + 3 * (size_t)i1);
    vn[2] = ld3(VN + 3 * (size_t)i2);
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
__global__ void __launch_bounds__(128) trace_fwd_kernel(BvhView B, const double* __restrict__ V64,
                                                        const double* __restrict__ origin, const double* __restrict__ dir,
                                                        int64_t N, double ext_ior, double int_ior,
                                                        double* __restrict__ out_ori, double* __restrict__ out_dir,
                                                        uint8_t* __restrict__ mask3, int4* __restrict__ rec,
                                                        int* __restrict__ rec_count, uint8_t* __restrict__ hit1, TileMap tiles)
{
    for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < N; item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = tiles.ray_of((int)item);
        d3 o = ld3(origin + 3 * i), d = ld3(dir + 3 * i);
        d3 oo = mk3(0, 0, 0), od = mk3(0, 0, 0);
        int id1, id2 = -1, id3;
        double t;
        bool valid = false;
        traverse<false>(B, cast_ray(o, d), t, id1);
        if (id1 >= 0) {
            HitRec h;
            d3 a0, a1, a2, o1, d1;
            load_tri64(B, V64, id1, a0, a1, a2);
            hit_forward(h, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
            if (!h.tir) {
                traverse<false>(B, cast_ray(o1, d1), t, id2);
                if (id2 >= 0) {
                    d3 o2, d2;
                    load_tri64(B, V64, id2, a0, a1, a2);
                    hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
                    if (!h.tir) {
                        traverse<true>(B, cast_ray(o2, d2), t, id3);
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

// Scatter the vertex gradients of one hit per lane with the atomics of equal-triangle RUNS merged first:
// records are in scanline-ish order, so neighbouring lanes often hit the same triangle (it is ~3 px wide);
// a segmented shuffle reduction over runs of equal id leaves one lane per run to issue the 9 RED.ADD.F64.
// All 32 lanes must call this (inactive lanes pass id = -1 and are never merged with anything).
__device__ __forceinline__ void scatter_runs(double* __restrict__ gV, const int32_t* __restrict__ F, int id, d3 ga[3])
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const int prev = __shfl_up_sync(FULL, id, 1);
    const bool head = lane == 0 || prev != id || id < 0;
    const unsigned heads = __ballot_sync(FULL, head);
    const int run = __popc(heads & (0xffffffffu >> (31 - lane)));  // index of the run this lane belongs to
    double v[9] = {ga[0].x, ga[0].y, ga[0].z, ga[1].x, ga[1].y, ga[1].z, ga[2].x, ga[2].y, ga[2].z};
#pragma unroll
    for (int delta = 1; delta < 32; delta <<= 1) {
        const int other_run = __shfl_down_sync(FULL, run, delta);
        const bool take = lane + delta < 32 && other_run == run;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const double o = __shfl_down_sync(FULL, v[j], delta);
            if (take) v[j] += o;
        }
    }
    if (head && id >= 0) {
        const int32_t* f = F + 3 * (size_t)id;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* p = gV + 3 * (size_t)f[c];
            atomicAdd(p, v[3 * c]); atomicAdd(p + 1, v[3 * c + 1]); atomicAdd(p + 2, v[3 * c + 2]);
        }
    }
}

// MERGE = true merges the atomics of equal-triangle runs first (scatter_runs): measured 1.28 -> 0.85 ms at C3
// (4 625 vertices take 95 M float64 atomics: contention-bound) but 0.56 -> 0.62 ms at C4 (25 126 vertices), so
// the host picks it by rays per vertex.
// Launch bound: 4 blocks per SM (128 registers, a few spills) beats 3 (154 registers): the kernel is latency-bound
// on its gathers and atomics, so residency wins -- ls_loss_bwd_kernel measured 0.65 ms at 3, 0.57 ms at 4, 0.65 at 2 (C4).
#ifndef DRT_BWD_MINB
#define DRT_BWD_MINB 4
#endif
// SMOOTH (optional smooth-normal mode): VN = vertex normals [nV,3], gVN receives their gradient.
template <bool MERGE, bool SMOOTH = false>
__global__ void __launch_bounds__(128, DRT_BWD_MINB) trace_bwd_kernel(BvhView B, const double* __restrict__ V64,
                                                        const double* __restrict__ origin, const double* __restrict__ dir,
                                                        double ext_ior, double int_ior, const int4* __restrict__ rec,
                                                        const int* __restrict__ rec_count,
                                                        const double* __restrict__ g_ori, const double* __restrict__ g_dir,
                                                        double* __restrict__ gV, const double* __restrict__ VN = nullptr,
                                                        double* __restrict__ gVN = nullptr)
{
    const int n = __ldg(rec_count);
    const int stride = gridDim.x * blockDim.x;
    // warp-uniform trip count: every lane of a warp runs the same iterations (the run merge is warp-wide)
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
        const int k = base + (threadIdx.x & 31);
        const bool active = k < n;
        int id1 = -1, id2 = -1;
        d3 z = mk3(0, 0, 0);
        d3 g1[3] = {z, z, z}, g2[3] = {z, z, z};
        if (active) {
            const int4 rc = __ldg(rec + k);
            const int64_t i = rc.x;
            id1 = rc.y; id2 = rc.z;
            const d3 o = ld3(origin + 3 * i), d = ld3(dir + 3 * i);
            d3 a0, a1, a2, o1, d1, o2, d2, go1, gd1, go0, gd0;
            d3 vn[3], gn[3];
            // Only ONE hit record is live at a time (a record is ~60 doubles): hit 1 is evaluated once for its
            // outgoing ray, hit 2 is evaluated and reversed, then hit 1 is re-evaluated and reversed.
            {
                HitRec h;
                load_tri64(B, V64, id1, a0, a1, a2);
                if (SMOOTH) load_vn(B, VN, id1, vn);
                hit_forward_t<SMOOTH>(h, o, d, a0, a1, a2, vn, ext_ior, int_ior, o1, d1);
            }
            {
                HitRec h;
                load_tri64(B, V64, id2, a0, a1, a2);
                if (SMOOTH) { load_vn(B, VN, id2, vn); gn[0] = gn[1] = gn[2] = z; }
                hit_forward_t<SMOOTH>(h, o1, d1, a0, a1, a2, vn, ext_ior, int_ior, o2, d2);
                d3 go2 = g_ori ? ld3(g_ori + 3 * i) : z;
                d3 gd2 = ld3(g_dir + 3 * i);
                hit_backward_t<SMOOTH>(h, go2, gd2, g2, gn, go1, gd1);
                if (SMOOTH) {
                    const int32_t* f = B.F + 3 * (size_t)id2;
                    scatter3(gVN, f[0], gn[0]); scatter3(gVN, f[1], gn[1]); scatter3(gVN, f[2], gn[2]);
                }
            }
            {
                HitRec h;
                load_tri64(B, V64, id1, a0, a1, a2);
                if (SMOOTH) { load_vn(B, VN, id1, vn); gn[0] = gn[1] = gn[2] = z; }
                hit_forward_t<SMOOTH>(h, o, d, a0, a1, a2, vn, ext_ior, int_ior, o1, d1);
                hit_backward_t<SMOOTH>(h, go1, gd1, g1, gn, go0, gd0);
                if (SMOOTH) {
                    const int32_t* f = B.F + 3 * (size_t)id1;
                    scatter3(gVN, f[0], gn[0]); scatter3(gVN, f[1], gn[1]); scatter3(gVN, f[2], gn[2]);
                }
            }
        }
        if (MERGE) {
            scatter_runs(gV, B.F, id2, g2);
            scatter_runs(gV, B.F, id1, g1);
        } else if (active) {
            const int32_t* f2 = B.F + 3 * (size_t)id2;
            scatter3(gV, f2[0], g2[0]); scatter3(gV, f2[1], g2[1]); scatter3(gV, f2[2], g2[2]);
            const int32_t* f1 = B.F + 3 * (size_t)id1;
            scatter3(gV, f1[0], g1[0]); scatter3(gV, f1[1], g1[1]); scatter3(gV, f1[2], g1[2]);
        }
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

// Same consumer, but driven by the compact records of the valid paths (7 % of the rays at C4) instead of a
// pass over all N rays: g_out_dir rows of other rays are NOT touched (the backward kernel never reads them).
__global__ void __launch_bounds__(256) ray_loss_rec_kernel(const double* __restrict__ out_ori,
                                                           const double* __restrict__ out_dir,
                                                           const double* __restrict__ screen,
                                                           const uint8_t* __restrict__ valid, const int4* __restrict__ rec,
                                                           const int* __restrict__ rec_count, double* __restrict__ g_dir,
                                                           double* __restrict__ loss_sum)
{
    const int n = __ldg(rec_count);
    double acc = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int64_t i = __ldg(rec + k).x;
        d3 g = mk3(0, 0, 0);
        if (!valid || valid[i]) {
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
