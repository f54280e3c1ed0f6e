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

// parked rays, component-major: component c of slot k at park[c * cap + k] (cap is a multiple of 32, so every
// 32-slot batch of every column starts 128-byte aligned)
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

// the float32 QUERY copy of a parked ray (= cast_ray of the float64 one: DiffRender.py:387-388), six
// component columns of 32-slot batches = six 128-byte runs per batch -- what the bulk-copy engine stages
struct QPark {
    float* __restrict__ p;
    int64_t cap;
    __device__ __forceinline__ void store(int k, const QRay& r) const
    {
        p[k] = r.ox; p[cap + k] = r.oy; p[2 * cap + k] = r.oz;
        p[3 * cap + k] = r.dx; p[4 * cap + k] = r.dy; p[5 * cap + k] = r.dz;
    }
    __device__ __forceinline__ void store_dead(int k) const { p[k] = __int_as_float(0x7fc00000); }  // NaN origin = nothing to trace
    __device__ __forceinline__ QRay load(int k) const
    {
        return QRay{p[k], p[cap + k], p[2 * cap + k], p[3 * cap + k], p[4 * cap + k], p[5 * cap + k]};
    }
};

// ---------------------------------------------------------------------------------------------------
// Staged persistent query: the rays of a warp's NEXT 32-ray batch are copied global -> shared by the
// bulk-copy engine (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) while the warp still traverses the
// current one, so refilling a finished lane is a shared-memory read instead of a DRAM/L2 round trip that
// stalls all 32 lanes.  That is what makes refilling before the whole warp is done (thresh < 32) pay: with
// refills straight from global memory every threshold below 32 measured 15-20 % SLOWER on B200, the warp
// scheduling model (tools/warp_sim) says 12-27 % fewer warp-wide steps if the refill is cheap.
// Two buffers per warp: batch n in use, batch n+1 in flight / landed.
// Job: kStageBytes, stage_issue(base, buf, mbar) [one lane], fetch(buf or nullptr, j, item, QRay&) -> bool,
//      retire(item, id).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int BYTES>
struct WarpStage {
    alignas(128) unsigned char buf[2][BYTES];
    alignas(8) unsigned long long mbar[2];
};

// policy bits: thresh | vote << 8 | staged << 16
__host__ __device__ constexpr int make_policy3(int thresh, int vote, int staged) { return thresh | (vote << 8) | (staged << 16); }

template <bool ANY, class Job>
__device__ __forceinline__ void staged_query(const BvhView& B, Job& job, int total, unsigned long long* work, int policy,
                                             WarpStage<Job::kStageBytes>* stg)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int thresh = policy & 0xff, vote = (policy >> 8) & 0xff;
    const bool staging = (policy >> 16) & 1;
    int pbase0 = 0, pbase1 = 0;   // first ray of the batch assigned to buffer 0 / 1
    unsigned phases = 0;          // bit b = parity the next completion of buffer b will have
    int cur = 0;
    bool started = false, more = total > 0;
    int batch_base = 0, batch_next = 0, batch_end = 0;
    bool batch_staged = false;

    // claim a batch for buffer b and start its copy (full batches only; the last, partial one is read directly)
    auto claim = [&](int b) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(work, 32ull);
        base = __shfl_sync(FULL, base, 0);
        const int bs = base >= (unsigned long long)total ? total : (int)base;
        if (b) pbase1 = bs; else pbase0 = bs;
        if (staging && lane == 0 && total - bs >= 32) job.stage_issue(bs, stg->buf[b], &stg->mbar[b]);
    };
    if (staging) {
        if (lane == 0) {
            mbar_init(&stg->mbar[0], 1);
            mbar_init(&stg->mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    if (more) { claim(0); claim(1); }

    int item = -1;
    RayQ q;
    float tmax = 0.f;
    double t_best = 0.0;
    int id_best = -1, node = kDone, sp = 0, nd = 0;
    int stack[kStackDepth];

    for (;;) {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // a batch may run out mid-way: second pass opens the next one
            unsigned idle = __ballot_sync(FULL, item < 0);
            if (!idle || !more) break;
            if (batch_next >= batch_end) {
                if (started) { claim(cur); cur ^= 1; }  // every ray of the finished batch sits in a lane: its buffer is free
                started = true;
                const int base = cur ? pbase1 : pbase0;
                if (base >= total) { more = false; break; }
                const int cnt = min(32, total - base);
                batch_staged = staging && cnt == 32;
                if (batch_staged) {
                    mbar_wait(&stg->mbar[cur], (phases >> cur) & 1u);
                    phases ^= 1u << cur;
                }
                batch_base = batch_next = base;
                batch_end = base + cnt;
            }
            if (item < 0) {
                int cand = batch_next + __popc(idle & lt_mask);
                if (cand < batch_end) {
                    item = cand;
                    t_best = INFINITY; id_best = -1; sp = 0; nd = 0; tmax = INFINITY; node = kDone;
                    QRay r;
                    if (job.fetch(batch_staged ? stg->buf[cur] : nullptr, cand - batch_base, cand, r) && B.nTris > 0) {
                        q = ray_setup(B, r);
                        node = 0;
                    }
                }
            }
            batch_next = min(batch_next + __popc(idle), batch_end);
        }
        const unsigned live = __ballot_sync(FULL, item >= 0);
        if (!live) break;
        const int need = min(thresh, __popc(live));
        for (;;) {
            if (vote) walk_vote(B, q, tmax, node, stack, sp, nd, vote);
            else walk(B, q, tmax, node, stack, sp, nd);
            if (drain<ANY>(B, q.r, stack, nd, t_best, id_best, tmax)) { node = kDone; sp = 0; }
            unsigned fin = __ballot_sync(FULL, item >= 0 && node == kDone);
            if (__popc(fin) >= need) break;
        }
        if (item >= 0 && node == kDone) {
            job.retire(item, id_best);
            item = -1;
        }
    }
}

// ---- Q1: entry query over all rays; only hits leave a trace ---------------------------------------
// PER_RAY_ORIGIN = false: the origin row is shared by rpo consecutive rays (read straight from L1);
// true: one origin row per ray, staged next to the direction.
template <bool PER_RAY_ORIGIN>
struct LossEntryJob {
    static constexpr int kStageBytes = PER_RAY_ORIGIN ? 1536 : 768;
    RaySrc rays;
    int4* __restrict__ L;
    int* __restrict__ countL;
    __device__ __forceinline__ void stage_issue(int base, unsigned char* buf, unsigned long long* mbar) const
    {
        mbar_expect_tx(mbar, kStageBytes);
        bulk_load(buf, rays.dir + 3 * (int64_t)base, 768, mbar);
        if (PER_RAY_ORIGIN) bulk_load(buf + 768, rays.origin + 3 * (int64_t)base, 768, mbar);
    }
    __device__ __forceinline__ bool fetch(const unsigned char* buf, int j, int i, QRay& r) const
    {
        d3 o, d;
        if (buf) {
            d = ld3(reinterpret_cast<const double*>(buf) + 3 * j);
            o = PER_RAY_ORIGIN ? ld3(reinterpret_cast<const double*>(buf + 768) + 3 * j) : rays.o(i);
        } else {
            o = rays.o(i);
            d = rays.d(i);
        }
        r = cast_ray(o, d);
        return true;
    }
    __device__ __forceinline__ void retire(int i, int id) const
    {
        int slot = warp_append<>(countL, id >= 0);
        if (slot >= 0) L[slot] = make_int4(i, id, -1, 0);
    }
};

template <int MINB, bool PER_RAY_ORIGIN>
__global__ void __launch_bounds__(128, MINB) ls_q1_kernel(BvhView B, LossEntryJob<PER_RAY_ORIGIN> job, int N,
                                                          unsigned long long* work, int policy)
{
    __shared__ WarpStage<LossEntryJob<PER_RAY_ORIGIN>::kStageBytes> stg[4];
    staged_query<false>(B, job, N, work, policy, &stg[threadIdx.x >> 5]);
}

// ---- R1: refraction at the entry hit, dense over L: float64 ray parked for R2, float32 copy for Q2 ----
__global__ void __launch_bounds__(128) ls_r1_kernel(BvhView B, const double* __restrict__ V64, RaySrc rays, double ext_ior,
                                                    double int_ior, int4* __restrict__ L, const int* __restrict__ countL,
                                                    Park park, QPark qpark)
{
    const int n = *countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 e = L[k];
        HitRec h;
        d3 a0, a1, a2, o1, d1;
        load_tri64(B, V64, e.y, a0, a1, a2);
        hit_forward(h, rays.o(e.x), rays.d(e.x), a0, a1, a2, ext_ior, int_ior, o1, d1);
        if (h.tir) {
            L[k].w = 1;  // dead
            qpark.store_dead(k);
        } else {
            park.store(k, o1, d1);
            qpark.store(k, cast_ray(o1, d1));
        }
    }
}

// query over a float32 ray list (Q2 over L slots, Q3 over M slots)
struct ListJobBase {
    static constexpr int kStageBytes = 768;
    QPark qp;
    __device__ __forceinline__ void stage_issue(int base, unsigned char* buf, unsigned long long* mbar) const
    {
        mbar_expect_tx(mbar, 768);
#pragma unroll
        for (int c = 0; c < 6; ++c) bulk_load(buf + 128 * c, qp.p + c * qp.cap + base, 128, mbar);
    }
    __device__ __forceinline__ bool fetch(const unsigned char* buf, int j, int k, QRay& r) const
    {
        if (buf) {
            const float* f = reinterpret_cast<const float*>(buf);
            r = QRay{f[j], f[32 + j], f[64 + j], f[96 + j], f[128 + j], f[160 + j]};
        } else {
            r = qp.load(k);
        }
        return r.ox == r.ox;  // NaN origin: dead slot
    }
};

// ---- Q2: exit query over L -------------------------------------------------------------------------
struct LossExitJob : ListJobBase {
    int4* __restrict__ L;
    __device__ __forceinline__ void retire(int k, int id) const { L[k].z = id; }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q2_kernel(BvhView B, LossExitJob job, const int* __restrict__ countL,
                                                          unsigned long long* work, int policy)
{
    __shared__ WarpStage<768> stg[4];
    staged_query<false>(B, job, *countL, work, policy, &stg[threadIdx.x >> 5]);
}

// ---- R2: refraction at the exit hit; survivors appended to M (slot of L) with the float32 exit ray --------
__global__ void __launch_bounds__(128) ls_r2_kernel(BvhView B, const double* __restrict__ V64, double ext_ior, double int_ior,
                                                    const int4* __restrict__ L, const int* __restrict__ countL, Park park,
                                                    int* __restrict__ M, int* __restrict__ countM, QPark qpark2)
{
    const int n = *countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 e = L[k];
        bool alive = false;
        d3 o2, d2;
        if (!e.w && e.z >= 0) {
            HitRec h;
            d3 a0, a1, a2, o1, d1;
            park.load(k, o1, d1);
            load_tri64(B, V64, e.z, a0, a1, a2);
            hit_forward(h, o1, d1, a0, a1, a2, ext_ior, int_ior, o2, d2);
            alive = !h.tir;
        }
        int slot = warp_append<>(countM, alive);
        if (slot >= 0) {
            M[slot] = k;
            qpark2.store(slot, cast_ray(o2, d2));
        }
    }
}

// ---- Q3: occlusion query over M; unoccluded slots of L appended to S -----------------------------------
struct LossOcclusionJob : ListJobBase {
    const int* __restrict__ M;
    int* __restrict__ S;
    int* __restrict__ countS;
    __device__ __forceinline__ void retire(int m, int id) const
    {
        int slot = warp_append<>(countS, id < 0);
        if (slot >= 0) S[slot] = M[m];
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) ls_q3_kernel(BvhView B, LossOcclusionJob job, const int* __restrict__ countM,
                                                          unsigned long long* work, int policy)
{
    __shared__ WarpStage<768> stg[4];
    staged_query<true>(B, job, *countM, work, policy, &stg[threadIdx.x >> 5]);
}

// ---- loss + backward over the valid paths -------------------------------------------------------------
// One thread per valid path: re-evaluates the two hits in float64 (bit-identical to R1/R2), forms
//   target = normalize(screen - out_ori),  diff = out_dir - target,  loss += |diff|^2,  g_out_dir = 2 diff
// (optim.py:99-106; out_ori is detached, optim.py:100, so g_out_ori = 0) and runs the analytic reverse
// of the chain (common.cuh:hit_backward) into grad_V.  GRAD = false: loss value only.
template <bool GRAD, bool MERGE>
__global__ void __launch_bounds__(128, 3) ls_loss_bwd_kernel(BvhView B, const double* __restrict__ V64, RaySrc rays,
                                                             double ext_ior, double int_ior, const int4* __restrict__ L,
                                                             const int* __restrict__ S, const int* __restrict__ countS,
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
            const int4 e = L[S[s]];
            const int i = e.x;
            d3 sp;
            if (tgt.get(i, sp)) {
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
