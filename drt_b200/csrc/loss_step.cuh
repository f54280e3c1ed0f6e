// loss_step.cuh -- the fused ray-loss step: forward wavefront + loss + vertex gradient with NO dense
// per-ray output.
//
// Replaces, for one batch of rays, the whole of Loss_calculator.ray_loss + loss.backward()
// (reference optim.py:91-108, 210):
//     out_ori, out_dir, mask = scene.render_transparent(origin, ray_dir)          DiffRender.py:420-432
//     target = normalize(screen_pixel - out_ori.detach())                         optim.py:99-101
//     loss   = sum over (valid & mask) of || out_dir - target ||^2                optim.py:103-106
//     vertices.grad += d loss / d vertices                                        optim.py:210
//
// Same five query / refraction stages as wavefront.cuh (same traversal, same float64 chain, so hit
// ids and exit rays are bit-identical to drt_trace_fwd), but
//   * rays that miss write nothing (the dense path zero-fills 51 B per missed ray);
//   * refracted rays are parked in a compact component-major scratch indexed by LIST SLOT
//     (coalesced 8-byte columns) instead of being scattered into out_ori/out_dir;
//   * the survivors of the occlusion query are a list of slots; one last kernel walks that list,
//     re-derives the exit ray, looks the screen target up, adds the loss term and runs the analytic
//     backward straight away -- d loss/d out_dir never exists in memory;
//   * the ray origin may be shared by `rays_per_origin` consecutive rays (a pinhole view has ONE
//     origin: captured_data.py:38 `ray_origin.T.expand_as(ray_dir)`), and the screen targets may be
//     sparse (sorted ray index + point; captured_data.py:104 `valid = screen_pixel[:,0] != 0`).
#pragma once
#include "wavefront.cuh"

namespace drt {

struct RaySrc {
    const double* __restrict__ origin;  // [ceil(N / rpo), 3]
    const double* __restrict__ dir;     // [N, 3]
    int rpo;                            // rays per origin row (1: one row per ray)
    __device__ __forceinline__ d3 o(int i) const { return ld3(origin + 3 * (int64_t)(rpo > 1 ? i / rpo : i)); }
    __device__ __forceinline__ d3 d(int i) const { return ld3(dir + 3 * (int64_t)i); }
};

// screen targets: dense (screen[N,3] + optional valid[N]) or sparse (sorted idx[n] + xyz[n,3]).
// Sparse lookups go through a bucket table built per call (tgt_bucket_kernel): bucket[b] = position of the
// first target with ray index >= b * 2^kTgtShift, so a lookup is two table reads and a binary search over
// at most 2^kTgtShift neighbouring entries (a plain search over millions of targets is ~22 dependent L2
// round trips per path: measured +0.36 ms per 49.8 M-ray step).
constexpr int kTgtShift = 6;

struct TargetSrc {
    const double* __restrict__ screen;
    const uint8_t* __restrict__ valid;
    const int32_t* __restrict__ idx;
    const double* __restrict__ xyz;
    const int* __restrict__ bucket;  // [(N >> kTgtShift) + 2]
    int n_tgt;
    int sparse;
    // slot of ray i's target (sparse: position in idx/xyz; dense: the ray index itself), -1 when the ray has none
    __device__ __forceinline__ int find(int i) const
    {
        if (sparse) {
            int lo = __ldg(bucket + (i >> kTgtShift)), hi = __ldg(bucket + (i >> kTgtShift) + 1);  // lower_bound in [lo, hi)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(idx + mid) < i) lo = mid + 1; else hi = mid;
            }
            return (lo < n_tgt && __ldg(idx + lo) == i) ? lo : -1;
        }
        return (valid && !valid[i]) ? -1 : i;
    }
    __device__ __forceinline__ d3 point(int slot) const { return ld3((sparse ? xyz : screen) + 3 * (int64_t)slot); }
    __device__ __forceinline__ bool get(int i, d3& s) const
    {
        if (sparse) {
            int lo = __ldg(bucket + (i >> kTgtShift)), hi = __ldg(bucket + (i >> kTgtShift) + 1);  // lower_bound in [lo, hi)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(idx + mid) < i) lo = mid + 1; else hi = mid;
            }
            if (lo >= n_tgt || __ldg(idx + lo) != i) return false;
            s = ld3(xyz + 3 * (int64_t)lo);
            return true;
        }
        if (valid && !valid[i]) return false;
        s = ld3(screen + 3 * (int64_t)i);
        return true;
    }
};

__global__ void __launch_bounds__(256) tgt_bucket_kernel(const int32_t* __restrict__ idx, int n_tgt, int n_buckets,
                                                         int* __restrict__ bucket)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const int64_t key = (int64_t)b << kTgtShift;
    int lo = 0, hi = n_tgt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(idx + mid) < key) lo = mid + 1; else hi = mid;
    }
    bucket[b] = lo;
}

// parked rays, component-major: component c of slot k at park[c * cap + k]
struct Park {
    double* __restrict__ p;
    int64_t cap;
    __device__ __forceinline__ void load(int k, d3& o, d3& d) const
    {
        o = mk3(p[k], p[cap + k], p[2 * cap + k]);
        d = mk3(p[3 * cap + k], p[4 * cap + k], p[5 * cap + k]);
    }
    __device__ __forceinline__ void store(int k, d3 o, d3 d) const
    {
        p[k] = o.x; p[cap + k] = o.y; p[2 * cap + k] = o.z;
        p[3 * cap + k] = d.x; p[4 * cap + k] = d.y; p[5 * cap + k] = d.z;
    }
};

// DRT_FUSE_R = 1: the refraction passes R1 / R2 run at the RETIRE of the entry / exit query (all lanes of a warp retire together,
// so the float64 code runs converged) instead of as two dense kernels over the list: two launches and two stage boundaries less
// per step.  Out of line, so that the traversal loop keeps its registers.
#ifndef DRT_FUSE_R
#define DRT_FUSE_R 0
#endif

struct RefractCtx {
    const double* __restrict__ V64;
    const int32_t* __restrict__ F;
    double ext_ior, int_ior;
};

// entry hit of ray (o, d) on triangle id: refracted ray -> park slot; false on total internal reflection
__device__ __noinline__ bool refract_entry(const RefractCtx c, const double* __restrict__ origin3, const double* __restrict__ dir3, int id,
                                           Park park, int slot)
{
    const int32_t* f = c.F + 3 * (size_t)id;
    HitRec h;
    d3 o1, d1;
    hit_forward(h, ld3(origin3), ld3(dir3), ld3(c.V64 + 3 * (size_t)f[0]), ld3(c.V64 + 3 * (size_t)f[1]), ld3(c.V64 + 3 * (size_t)f[2]),
                c.ext_ior, c.int_ior, o1, d1);
    if (h.tir) return false;
    park.store(slot, o1, d1);
    return true;
}

// exit hit of the parked ray of `slot` on triangle id: exit ray parked in place; false on total internal reflection
__device__ __noinline__ bool refract_exit(const RefractCtx c, int id, Park park, int slot)
{
    const int32_t* f = c.F + 3 * (size_t)id;
    HitRec h;
    d3 o1, d1, o2, d2;
    park.load(slot, o1, d1);
    hit_forward(h, o1, d1, ld3(c.V64 + 3 * (size_t)f[0]), ld3(c.V64 + 3 * (size_t)f[1]), ld3(c.V64 + 3 * (size_t)f[2]), c.ext_ior, c.int_ior, o2,
                d2);
    if (h.tir) return false;
    park.store(slot, o2, d2);
    return true;
}

// ---- Q1: entry query over all rays; only hits leave a trace ---------------------------------------
// Work item -> ray through TileMap (trace.cuh): with the image size known a warp's batch is an 8 x 4 pixel tile
// (the warp scheduling model, tools/warp_sim, gives -22 % warp-wide steps for Q1 and -10 % for Q2/Q3).
struct LossEntryJob {
    static constexpr bool kBulkMiss = false;
    static constexpr bool kMissWrites = false;  // a missed ray leaves no trace
    __device__ __forceinline__ bool bulk_miss(int, unsigned) { return false; }
    __device__ __forceinline__ void finish(unsigned) {}
    RaySrc rays;
    int4* __restrict__ L;
    int* __restrict__ countL;
    TileMap tiles;
    int base;  // first ray of this launch's share of the batch (drt_ray_loss_step may split a batch over two streams)
#if DRT_FUSE_R
    RefractCtx rc;
    Park park;
#endif
    __device__ __forceinline__ int ray_of(int item) const { return tiles.ray_of(item + base); }
    __device__ __forceinline__ bool load(int item, d3& o, d3& d) const
    {
        load_ray(ray_of(item), o, d);
        return true;
    }
    __device__ __forceinline__ void load_ray(int i, d3& o, d3& d) const  // by ray index
    {
        o = rays.o(i);
        d = rays.d(i);
    }
    __device__ __forceinline__ bool same_origin_row(int i, int j) const { return rays.rpo > 1 && i / rays.rpo == j / rays.rpo; }
    __device__ __forceinline__ d3 origin_of(int i) const { return rays.o(i); }
    __device__ __forceinline__ d3 dir_of(int i) const { return rays.d(i); }
    __device__ __forceinline__ void retire(int item, int id, double) const
    {
        int slot = warp_append<>(countL, id >= 0);
        if (slot >= 0) {
            const int i = ray_of(item);
#if DRT_FUSE_R
            const int64_t row = rays.rpo > 1 ? i / rays.rpo : i;
            const bool ok = refract_entry(rc, rays.origin + 3 * row, rays.dir + 3 * (int64_t)i, id, park, slot);
            L[slot] = make_int4(i, id, -1, ok ? 0 : 1);
#else
            L[slot] = make_int4(i, id, -1, 0);
#endif
        }
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q1_kernel(BvhView B, LossEntryJob job, int N, unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    persistent_query<false>(B, job, N, work, policy, stack);
}

#if DRT_QNODE
__global__ void __launch_bounds__(128, 8) ls_beam_kernel(BvhView B, LossEntryJob job, int N, unsigned long long* work, int tpb,
                                                         int max_steps, int2* __restrict__ tiles, int* __restrict__ n_tiles,
                                                         TileBeams prepared)
{
    beam_pass(B, job, N, work, tpb, max_steps, tiles, n_tiles, prepared, job.base);
}

// drt_tile_beams: the per-tile direction intervals of a ray batch, once per view set (TileBeams, wavefront.cuh).  One thread per
// tile runs the same scan the beam pass would run in every step; thread 0 writes the signature.
__global__ void __launch_bounds__(128) tile_beams_kernel(LossEntryJob job, int N, float4* __restrict__ out, int64_t stride, int sig_hw_tw)
{
    const int64_t n_tiles = ((int64_t)N + 31) >> 5;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
        const TileBeam b = tile_scan(job, (int)(t << 5), N);
        const unsigned fl = (b.has_rays ? 1u : 0u) | (b.shared_origin ? 2u : 0u);
        out[1 + t] = make_float4(b.dmn[0], b.dmn[1], b.dmn[2], __uint_as_float(fl));
        out[stride + 1 + t] = make_float4(b.dmx[0], b.dmx[1], b.dmx[2], b.ox);
        out[2 * stride + 1 + t] = make_float4(b.oy, b.oz, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        out[0] = make_float4(__uint_as_float(kBeamMagic), __int_as_float(N), __int_as_float(job.tiles.img_w), __int_as_float(sig_hw_tw));
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q1_tiles_kernel(BvhView B, LossEntryJob job, int N, const int2* __restrict__ tiles,
                                                                const int* __restrict__ n_tiles, unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    entry_query_tiles(B, job, N, tiles, n_tiles, work, policy, stack);
}
#endif

// ---- R1: refraction at the entry hit, dense over L, refracted ray parked at its slot ---------------
__global__ void __launch_bounds__(128) ls_r1_kernel(BvhView B, const double* __restrict__ V64, RaySrc rays, double ext_ior,
                                                    double int_ior, int4* __restrict__ L, const int* __restrict__ countL,
                                                    Park park)
{
    const int n = *countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 e = L[k];
        HitRec h;
        d3 a0, a1, a2, o1, d1;
        load_tri64(B, V64, e.y, a0, a1, a2);
        hit_forward(h, rays.o(e.x), rays.d(e.x), a0, a1, a2, ext_ior, int_ior, o1, d1);
        if (h.tir) L[k].w = 1;  // dead
        else park.store(k, o1, d1);
    }
}

// ---- Q2: exit query over L -------------------------------------------------------------------------
struct LossExitJob {
    static constexpr bool kBulkMiss = false;
    __device__ __forceinline__ bool bulk_miss(int, unsigned) { return false; }
    __device__ __forceinline__ void finish(unsigned) {}
    Park park;
    int4* __restrict__ L;
#if DRT_FUSE_R
    RefractCtx rc;
    TargetSrc tgt;
    int2* __restrict__ M;
    int* __restrict__ countM;
#endif
    __device__ __forceinline__ bool load(int k, d3& o, d3& d) const
    {
        if (L[k].w) return false;
        park.load(k, o, d);
        return true;
    }
    __device__ __forceinline__ void retire(int k, int id, double) const
    {
        L[k].z = id;
#if DRT_FUSE_R
        const bool alive = id >= 0 && !L[k].w && refract_exit(rc, id, park, k);
        int slot = warp_append<>(countM, alive);
        if (slot >= 0) M[slot] = make_int2(k, tgt.find(L[k].x));
#endif
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q2_kernel(BvhView B, LossExitJob job, const int* __restrict__ countL,
                                                          unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    persistent_query<false>(B, job, *countL, work, policy, stack);
}

// ---- R2: refraction at the exit hit; exit ray parked in place, surviving SLOTS appended to M ---------
// M entry = (slot in L, slot of the ray's screen target or -1): the target search (a bucket lookup + <= 6 dependent reads) is
// done HERE, in a dense 32-lane kernel with occupancy to spare, instead of at the head of the register-bound loss/backward
// kernel, where ncu (r02a) attributed a quarter of its stall samples to that dependent chain.
__global__ void __launch_bounds__(128) ls_r2_kernel(BvhView B, const double* __restrict__ V64, double ext_ior, double int_ior,
                                                    const int4* __restrict__ L, const int* __restrict__ countL, Park park, TargetSrc tgt,
                                                    int2* __restrict__ M, int* __restrict__ countM)
{
    const int n = *countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 e = L[k];
        bool alive = false;
        if (!e.w && e.z >= 0) {
            HitRec h;
            d3 a0, a1, a2, o1, d1, o2, d2;
            park.load(k, o1, d1);
            load_tri64(B, V64, e.z, a0, a1, a2);
            hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
            alive = !h.tir;
            if (alive) park.store(k, o2, d2);
        }
        int slot = warp_append<>(countM, alive);
        if (slot >= 0) M[slot] = make_int2(k, tgt.find(e.x));
    }
}

// ---- Q3: occlusion query over M; unoccluded slots appended to S ---------------------------------------
struct LossOcclusionJob {
    static constexpr bool kBulkMiss = false;
    __device__ __forceinline__ bool bulk_miss(int, unsigned) { return false; }
    __device__ __forceinline__ void finish(unsigned) {}
    Park park;
    const int2* __restrict__ M;
    const int4* __restrict__ L;
    int4* __restrict__ S;  // valid paths: (ray, tri1, tri2, target slot) -- everything the loss/backward kernel needs, one load
    int* __restrict__ countS;
    __device__ __forceinline__ bool load(int m, d3& o, d3& d) const
    {
        park.load(M[m].x, o, d);
        return true;
    }
    __device__ __forceinline__ void retire(int m, int id, double) const
    {
        int slot = warp_append<>(countS, id < 0);
        if (slot >= 0) {
            const int2 e = M[m];
            const int4 l = L[e.x];
            S[slot] = make_int4(l.x, l.y, l.z, e.y);
        }
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q3_kernel(BvhView B, LossOcclusionJob job, const int* __restrict__ countM,
                                                          unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    persistent_query<true>(B, job, *countM, work, policy, stack);
}

// ---- the whole forward path of a ray in ONE thread (small batches) --------------------------------------
// One view per iteration is how the reference is used (optim.py:95), i.e. 10^5 .. 10^6 rays per call: the seven stage
// launches above then cost more in launch ramps and per-stage tails (a stage cannot end before its longest ray) than the
// divergence they remove.  Here a thread runs entry query -> refraction -> exit query -> refraction -> occlusion query for
// its ray (same traversal, same float64 chain: identical hit ids and exit rays) and appends the valid path to S, which
// ls_loss_bwd_kernel consumes as usual: 2 launches per step instead of 8.  countL / countM only feed the stage statistics.
#ifndef DRT_DIRECT_MINB
#define DRT_DIRECT_MINB 1
#endif
__global__ void __launch_bounds__(128, DRT_DIRECT_MINB) ls_direct_kernel(BvhView B, const double* __restrict__ V64, RaySrc rays, TileMap tiles, int N,
                                                        double ext_ior, double int_ior, TargetSrc tgt, int4* __restrict__ S,
                                                        int* __restrict__ countL, int* __restrict__ countM, int* __restrict__ countS)
{
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < N; item += gridDim.x * blockDim.x) {
        const int i = tiles.ray_of(item);
        const d3 o = rays.o(i), d = rays.d(i);
        int id1, id2 = -1, id3 = -1;
        double t;
        bool alive = false;
        traverse<false>(B, cast_ray(o, d), t, id1);
        if (id1 >= 0) {
            HitRec h;
            d3 a0, a1, a2, o1, d1, o2, d2;
            load_tri64(B, V64, id1, a0, a1, a2);
            hit_forward(h, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
            if (!h.tir) {
                traverse<false>(B, cast_ray(o1, d1), t, id2);
                if (id2 >= 0) {
                    load_tri64(B, V64, id2, a0, a1, a2);
                    hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
                    if (!h.tir) {
                        alive = true;
                        traverse<true>(B, cast_ray(o2, d2), t, id3);
                    }
                }
            }
        }
        warp_append<>(countL, id1 >= 0);
        warp_append<>(countM, alive);
        const int slot = warp_append<>(countS, alive && id3 < 0);
        if (slot >= 0) S[slot] = make_int4(i, id1, id2, tgt.find(i));
    }
}

// ---- loss + backward over the valid paths -------------------------------------------------------------
// One thread per valid path: re-evaluates the two hits in float64 (bit-identical to R1/R2), forms
//   target = normalize(screen - out_ori),  diff = out_dir - target,  loss += |diff|^2,  g_out_dir = 2 diff
// (optim.py:99-106; out_ori is detached, optim.py:100, so g_out_ori = 0) and runs the analytic reverse
// of the chain (common.cuh:hit_backward) into grad_V.  GRAD = false: loss value only.
template <bool GRAD, bool MERGE>
__global__ void __launch_bounds__(128, DRT_BWD_MINB) ls_loss_bwd_kernel(BvhView B, const double* __restrict__ V64, RaySrc rays,
                                                             double ext_ior, double int_ior,
                                                             const int4* __restrict__ S, const int* __restrict__ countS,
                                                             TargetSrc tgt, double* __restrict__ loss_sum,
                                                             double* __restrict__ gV)
{
    const int n = __ldg(countS);
    const int stride = gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
        const int s = base + (threadIdx.x & 31);
        int id1 = -1, id2 = -1;
        d3 z = mk3(0, 0, 0);
        d3 g1[3] = {z, z, z}, g2[3] = {z, z, z};
        if (s < n) {
            const int4 e = __ldg(S + s);  // (ray, tri1, tri2, target slot)
            const int i = e.x;
            if (e.w >= 0) {
                const d3 sp = tgt.point(e.w);
                const d3 o = rays.o(i), d = rays.d(i);
                d3 a0, a1, a2, o1, d1, o2, d2, go1, gd1, go0, gd0;
                {
                    HitRec h;
                    load_tri64(B, V64, e.y, a0, a1, a2);
                    hit_forward(h, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
                }
                {
                    HitRec h;
                    load_tri64(B, V64, e.z, a0, a1, a2);
                    hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
                    d3 tg = sp - o2;
                    tg = divs(tg, __dsqrt_rn(dot(tg, tg)));
                    const d3 df = d2 - tg;
                    acc += dot(df, df);
                    if (GRAD) hit_backward(h, z, df * 2.0, g2, go1, gd1);
                }
                if (GRAD) {
                    HitRec h;
                    load_tri64(B, V64, e.y, a0, a1, a2);
                    hit_forward(h, o, d, a0, a1, a2, ext_ior, int_ior, o1, d1);
                    hit_backward(h, go1, gd1, g1, go0, gd0);
                    id1 = e.y; id2 = e.z;
                }
            }
        }
        if (GRAD) {
            if (MERGE) {
                scatter_runs(gV, B.F, id2, g2);
                scatter_runs(gV, B.F, id1, g1);
            } else if (id1 >= 0) {
                const int32_t* f2 = B.F + 3 * (size_t)id2;
                scatter3(gV, f2[0], g2[0]); scatter3(gV, f2[1], g2[1]); scatter3(gV, f2[2], g2[2]);
                const int32_t* f1 = B.F + 3 * (size_t)id1;
                scatter3(gV, f1[0], g1[0]); scatter3(gV, f1[1], g1[1]); scatter3(gV, f1[2], g1[2]);
            }
        }
    }
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if ((threadIdx.x & 31) == 0 && loss_sum && acc != 0.0) atomicAdd(loss_sum, acc);
}

// ---- captured_data.generate_ray (captured_data.py:23-40) on the device --------------------------------
// pixel (x, y, 1) -> K^-1 -> camera-to-world rotation + translation -> direction from the camera
// centre, normalised.  Same expression order as the reference's two matrix products (row . column,
// accumulated left to right) and its `ray_dir / ray_dir.norm(dim=1)`; results agree with the torch
// evaluation to a few ulp (matmul summation order is the library's).  origin3 receives the ONE camera
// centre of the view (the reference returns it expanded to [N,3]).
__global__ void __launch_bounds__(256) generate_rays_kernel(int resy, int resx, const double* __restrict__ Kinv,
                                                            const double* __restrict__ Rinv, double* __restrict__ origin3,
                                                            double* __restrict__ dir)
{
    const int64_t n = (int64_t)resy * resx;
    double K[9], R[12];
#pragma unroll
    for (int j = 0; j < 9; ++j) K[j] = Kinv[j];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) R[4 * r + c] = Rinv[4 * r + c];
    if (blockIdx.x == 0 && threadIdx.x < 3) origin3[threadIdx.x] = R[4 * threadIdx.x + 3];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = (double)(i % resx), py = (double)(i / resx);
        d3 c, w;
        c.x = addr(addr(mulr(K[0], px), mulr(K[1], py)), K[2]);
        c.y = addr(addr(mulr(K[3], px), mulr(K[4], py)), K[5]);
        c.z = addr(addr(mulr(K[6], px), mulr(K[7], py)), K[8]);
        w.x = addr(addr(addr(mulr(R[0], c.x), mulr(R[1], c.y)), mulr(R[2], c.z)), R[3]);
        w.y = addr(addr(addr(mulr(R[4], c.x), mulr(R[5], c.y)), mulr(R[6], c.z)), R[7]);
        w.z = addr(addr(addr(mulr(R[8], c.x), mulr(R[9], c.y)), mulr(R[10], c.z)), R[11]);
        d3 v = w - mk3(R[3], R[7], R[11]);
        v = divs(v, __dsqrt_rn(dot(v, v)));
        st3(dir + 3 * i, v);
    }
}

}  // namespace drt
