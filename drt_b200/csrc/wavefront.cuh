// wavefront.cuh -- the production forward tracer: a stage-synchronous wavefront whose three ray
// queries run as PERSISTENT warps with dynamic work fetch and ballot compaction.
//
// Replaces Scene.render_transparent (reference DiffRender.py:420-432; trace2 :537-546).
//
//   Q1   entry query over all N rays            -> misses retire (zeros), hits compacted into list L
//   R1   refraction 1, float64, dense over L    -> refracted ray parked in out_ori/out_dir
//   Q2   exit query over L                      -> L[k].z = tri2
//   R2   refraction 2, float64, dense over L    -> exit ray parked; survivors compacted into list M
//   Q3   occlusion query (any hit) over M       -> mask = 1 + backward record, or zeros
//
// Why not one megakernel with one thread per path (trace_fwd_kernel in trace.cuh, kept as the
// simple variant): measured on B200 it ran its box tests with ~10 of 32 lanes active -- lanes whose
// query ended early, lanes without a hit and lanes parked at a leaf all idle while the warp's
// longest traversal finishes.  Here every query kernel is pure traversal: a lane that finishes
// retires its result with a couple of stores and is refilled from a global work counter (one
// atomic per warp per batch), and the float64 refraction math runs dense over compacted lists with
// its own register budget.
#pragma once
#include "trace.cuh"

namespace drt {

#ifndef DRT_FETCH_BATCH
#define DRT_FETCH_BATCH 32
#endif
constexpr int kFetchBatch = DRT_FETCH_BATCH;

// the persistent query kernels walk the 4-wide view of the tree when it is built (bvh.cuh: DRT_BVH4)
#if DRT_QNODE && DRT_BVH4
constexpr bool kWideQueries = true;
#else
constexpr bool kWideQueries = false;
#endif

template <typename Dummy = void>
__device__ __forceinline__ int warp_append(int* __restrict__ counter, bool pred)
{
    // ballot compaction: one atomic per warp, slot = base + rank among the lanes that append
    namespace cg = cooperative_groups;
    int slot = -1;
    if (pred) {
        cg::coalesced_group g = cg::coalesced_threads();
        int base = 0;
        if (g.thread_rank() == 0) base = atomicAdd(counter, (int)g.size());
        slot = g.shfl(base, 0) + (int)g.thread_rank();
    }
    return slot;
}

// Persistent query driver.  Job: bool load(int item, d3& o, d3& d)  (false = nothing to trace),
//                                void retire(int item, int id, double t).
// Lanes step their traversals together; once `thresh` of the live lanes have finished, those retire
// and are refilled from the global work counter (thresh = 32: refill only when the whole warp is done).
// `policy` = thresh | vote << 8; vote = 0: every lane walks until its own leaf queue is full (walk),
// vote = 1..31: the warp drains as soon as that many lanes are blocked (walk_vote).
__host__ __device__ constexpr int make_policy(int thresh, int vote) { return thresh | (vote << 8); }

template <bool ANY, class Job>
__device__ __forceinline__ void persistent_query(const BvhView& B, Job& job, int total, unsigned long long* work, int policy, QueryStack& stack)
{
    const int thresh = policy & 0xff, vote = policy >> 8;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int batch_next = 0, batch_end = 0;
    bool more = total > 0;
    int item = -1;
    RayQ q;
    float tmax = 0.f;
    double t_best = 0.0;
    int id_best = -1, node = kDone, sp = 0, nd = 0;

    for (;;) {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // a batch may run out mid-way: second pass opens the next one
            unsigned idle = __ballot_sync(FULL, item < 0);
            if (!idle || !more) break;
            if (batch_next >= batch_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(work, (unsigned long long)kFetchBatch);
                base = __shfl_sync(FULL, base, 0);
                if (base >= (unsigned long long)total) { more = false; break; }
                batch_next = (int)base;
                batch_end = min((int)base + kFetchBatch, total);
            }
            if (item < 0) {
                int cand = batch_next + __popc(idle & lt_mask);
                if (cand < batch_end) {
                    item = cand;
                    d3 o, d;
                    t_best = INFINITY; id_best = -1; sp = 0; nd = 0; tmax = INFINITY; node = kDone;
                    if (job.load(item, o, d) && B.nTris > 0) {
                        q = ray_setup(B, cast_ray(o, d));
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
            if (vote) walk_vote<kWideQueries>(B, q, tmax, node, stack, sp, nd, vote);
            else walk<kWideQueries>(B, q, tmax, node, stack, sp, nd);
            if (DRT_COOP_DRAIN ? drain_coop<ANY>(B, q, stack, nd, t_best, id_best, tmax) : drain<ANY>(B, q, stack, nd, t_best, id_best, tmax)) { node = kDone; stack.reset(sp); }
            unsigned fin = __ballot_sync(FULL, item >= 0 && node == kDone);
            if (__popc(fin) >= need) break;
        }
        // a whole warp of consecutive rays that all missed retires with bulk stores (TMA engine), see EntryJob
        if (Job::kBulkMiss) {
            const bool mine = item >= 0 && node == kDone && id_best < 0;
            const int first = __shfl_sync(FULL, item, 0);
            if (__all_sync(FULL, mine && item == first + (int)lane) && (first & 31) == 0 && job.bulk_miss(first, lane)) item = -1;
        }
        if (item >= 0 && node == kDone) {
            job.retire(item, id_best, t_best);
            item = -1;
        }
    }
    job.finish(lane);
}

#if DRT_QNODE
// Entry query with BEAM CULLING (trace.cuh: BeamQ).  The primary rays of a view share one origin and arrive as pixel tiles,
// and most of them miss (88 % at C4) after ~10 node steps each.  Two launches:
//  beam_pass: a warp takes `tpb` tiles (32 rays each) per work fetch;
//   A  lane t reads the 32 rays of tile t: direction intervals, and whether all of them start at the same point (checked
//      per tile on the data itself -- no promise from the caller; tiles that fail it are traced ray by ray from the root);
//   B  every lane walks the tree with ITS tile's beam: no leaf box touched -> the 32 rays of the tile retire as misses for
//      the price of one traversal; otherwise the tile is appended to a list together with its entry point, the first node
//      where the beam forks.  A beam that is still undecided after `max_steps` node steps is kept (the lanes of a warp wait
//      for its slowest beam, and the undecided ones hug the silhouette);
//  entry_query_tiles: persistent warps fetch ONE listed tile at a time (a tile is the unit of dynamic balancing: a fused
//      variant that traced the surviving tiles of its own 1024-ray batch in place measured slower than no culling at all --
//      batches whose 32 tiles all survive are 30x longer than empty ones and the kernel ends in their tail) and trace its
//      rays exactly as before (walk_vote / drain), starting at the entry point.
// Hit ids are unchanged (the beam only removes box tests that every ray of the tile would fail).
// The direction intervals and the common origin of ONE tile (32 work items), i.e. everything the beam test needs to know about
// the tile's rays.  They depend on the rays only -- not on the mesh -- and DRT's view sets are fixed for a whole optimisation
// (captured_data.py:94-108: loaded once, optim.py:95 cycles through them), so a caller may compute them ONCE per view set
// (drt_tile_beams) and hand them to every later step instead of having the beam pass re-read all ray directions (1.19 GB per
// step at C4 for 75 MB of intervals).
struct TileBeam {
    float dmn[3], dmx[3];
    float ox, oy, oz;
    bool has_rays, shared_origin;
};

// Prepared tile beams, structure of arrays: three float4 planes of `stride` = n_tiles + 1 entries.
//   plane 0: (dmin.x, dmin.y, dmin.z, flags)   flags bit 0 = has rays, bit 1 = all rays start at one point and none is NaN
//   plane 1: (dmax.x, dmax.y, dmax.z, o.x)     plane 2: (o.y, o.z, -, -)
// Entry 0 of plane 0 is a SIGNATURE (magic, N, image width, pixels per image << 3 | log2 tile width): the beam pass uses the
// prepared intervals only when it matches its own tile map, and scans the rays itself otherwise.
constexpr unsigned kBeamMagic = 0x4D414542u;  // "BEAM"
struct TileBeams {
    const float4* __restrict__ p;  // nullptr: none prepared
    int64_t stride;
    int sig_n, sig_w, sig_hw_tw;  // what the signature must say
    __device__ __forceinline__ bool usable() const
    {
        if (!p) return false;
        const float4 h = __ldg(p);
        return __float_as_uint(h.x) == kBeamMagic && __float_as_int(h.y) == sig_n && __float_as_int(h.z) == sig_w && __float_as_int(h.w) == sig_hw_tw;
    }
    __device__ __forceinline__ TileBeam load(int64_t tile) const
    {
        const float4 a = __ldg(p + 1 + tile), b = __ldg(p + stride + 1 + tile), c = __ldg(p + 2 * stride + 1 + tile);
        const unsigned fl = __float_as_uint(a.w);
        return TileBeam{{a.x, a.y, a.z}, {b.x, b.y, b.z}, b.w, c.x, c.y, (fl & 1u) != 0, (fl & 2u) != 0};
    }
};

// scans the 32 rays of the tile whose first work item is `first` (read in turn: neighbouring lanes read neighbouring tiles,
// every sector a lane touches is used up by its next loads)
template <class Job>
__device__ __forceinline__ TileBeam tile_scan(const Job& job, int first, int total)
{
    TileBeam t{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}, 0.f, 0.f, 0.f, false, true};
    // the 32 work items of a tile map to rays r0 + row * img_w + col (TileMap): one index computation per tile, not per ray
    const int r0 = job.ray_of(first);
    const int tw_log2 = job.tiles.tw_log2, tw_mask = (1 << tw_log2) - 1, row_stride = job.tiles.img_w;
    // rays that provably read the same origin row (one row per view, captured_data.py:38) need it once per tile
    const int r_last = row_stride ? r0 + (31 >> tw_log2) * row_stride + tw_mask : r0 + 31;
    const bool one_row = first + 31 < total && job.same_origin_row(r0, r_last);
    d3 o_tile = mk3(0, 0, 0);
    if (one_row) o_tile = job.origin_of(r0);
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        d3 o = o_tile, d;
        if (first + j < total) {
            const int i = row_stride ? r0 + (j >> tw_log2) * row_stride + (j & tw_mask) : r0 + j;
            if (one_row) d = job.dir_of(i);
            else job.load_ray(i, o, d);
            const QRay r = cast_ray(o, d);
            if (!t.has_rays) { t.ox = r.ox; t.oy = r.oy; t.oz = r.oz; t.has_rays = true; }
            t.shared_origin = t.shared_origin && r.ox == t.ox && r.oy == t.oy && r.oz == t.oz;
            t.dmn[0] = fminf(t.dmn[0], r.dx); t.dmn[1] = fminf(t.dmn[1], r.dy); t.dmn[2] = fminf(t.dmn[2], r.dz);
            t.dmx[0] = fmaxf(t.dmx[0], r.dx); t.dmx[1] = fmaxf(t.dmx[1], r.dy); t.dmx[2] = fmaxf(t.dmx[2], r.dz);
            if (!(r.dx == r.dx && r.dy == r.dy && r.dz == r.dz)) t.shared_origin = false;  // a NaN direction: no beam
        }
    }
    return t;
}

template <class Job>
__device__ __forceinline__ void beam_pass(const BvhView& B, Job& job, int total, unsigned long long* work, int tpb, int max_steps,
                                          int2* __restrict__ tiles, int* __restrict__ n_tiles, TileBeams prepared = TileBeams{nullptr, 0, 0, 0, 0},
                                          int base_item = 0)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    int stack[kStackDepth];
    const bool use_prepared = prepared.usable();
    for (;;) {
        unsigned long long base64 = 0;
        if (lane == 0) base64 = atomicAdd(work, (unsigned long long)(32 * tpb));
        base64 = __shfl_sync(FULL, base64, 0);
        if (base64 >= (unsigned long long)total) break;
        const int base = (int)base64;
        // ---- A: direction intervals and the common origin of the lane's own tile: prepared by the caller, or scanned here ----
        TileBeam tb{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0.f, 0.f, 0.f, false, true};
        if ((int)lane < tpb && base + 32 * (int)lane < total) {
            const int first = base + 32 * (int)lane;
            if (use_prepared) tb = prepared.load(((int64_t)first + base_item) >> 5);
            else tb = tile_scan(job, first, total);
        }
        // ---- B: one beam per lane ------------------------------------------------------------------------------
        bool keep = false;
        int entry = 0;
        if (tb.has_rays) {
            keep = true;  // different origins inside the tile: no beam, every ray from the root
            if (tb.shared_origin && B.nTris > 0) {
                const BeamQ bq = beam_setup(B, tb.ox, tb.oy, tb.oz, tb.dmn, tb.dmx);
                keep = beam_walk(B, bq, stack, entry, max_steps);
            }
        }
        const int slot = warp_append<>(n_tiles, keep);
        if (slot >= 0) tiles[slot] = make_int2(base + 32 * (int)lane, entry);
        if (Job::kMissWrites) {  // culled tiles: their rays retire as misses (dense outputs are zero-filled)
            unsigned culled = __ballot_sync(FULL, tb.has_rays && !keep);
            while (culled) {
                const int t = __ffs(culled) - 1;
                culled &= culled - 1;
                const int item = base + 32 * t + (int)lane;
                if (item < total) job.retire(item, -1, INFINITY);
            }
        }
    }
}

template <class Job>
__device__ __forceinline__ void entry_query_tiles(const BvhView& B, Job& job, int total, const int2* __restrict__ tiles,
                                                  const int* __restrict__ n_tiles, unsigned long long* work, int policy, QueryStack& stack)
{
    const int vote = policy >> 8;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const int n = *n_tiles;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(work, 1ull);
        k = __shfl_sync(FULL, k, 0);
        if (k >= (unsigned long long)n) break;
        const int2 tl = __ldg(tiles + k);
        const int item = tl.x + (int)lane;
        d3 o, d;
        RayQ q;
        float tmax = INFINITY;
        double t_best = INFINITY;
        int id_best = -1, node = kDone, sp = 0, nd = 0;
        const bool act = item < total && job.load(item, o, d);
        if (act && B.nTris > 0) {
            q = ray_setup(B, cast_ray(o, d));
            node = tl.y;
        } else {
            q = ray_setup(B, QRay{0.f, 0.f, 0.f, 1.f, 1.f, 1.f});
        }
        for (;;) {
            if (vote) walk_vote<kWideQueries>(B, q, tmax, node, stack, sp, nd, vote);
            else walk<kWideQueries>(B, q, tmax, node, stack, sp, nd);
            if (DRT_COOP_DRAIN) drain_coop<false>(B, q, stack, nd, t_best, id_best, tmax);
            else drain<false>(B, q, stack, nd, t_best, id_best, tmax);
            if (!__any_sync(FULL, node != kDone)) break;
        }
        if (act) job.retire(item, id_best, t_best);
    }
    job.finish(lane);
}
#endif  // DRT_QNODE


// ---- Q1 ------------------------------------------------------------------------------------------
// 32 x 24 B of zeros, the source of the bulk zero-fill; one per block, written once, read by the async proxy
struct ZeroTile {
    alignas(128) unsigned char z[768];
};

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes)
{
    // cp.async.bulk shared::cta -> global (UBLKCP): the copy is done by the TMA engine, no LSU wavefronts
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}

struct EntryJob {
    static constexpr bool kBulkMiss = true;
    static constexpr bool kMissWrites = true;  // a missed ray has outputs (zeros)
    const ZeroTile* zeros;  // shared memory; nullptr disables the bulk path (unaligned outputs)
    bool issued;
    const double* __restrict__ origin;
    const double* __restrict__ dir;
    double* __restrict__ out_ori;
    double* __restrict__ out_dir;
    uint8_t* __restrict__ mask3;
    uint8_t* __restrict__ hit1;
    int4* __restrict__ L;
    int* __restrict__ countL;
    TileMap tiles;  // work item -> ray (8 x 4 pixel tiles when the image size is known)
    __device__ __forceinline__ int ray_of(int item) const { return tiles.ray_of(item); }
    __device__ __forceinline__ bool load(int item, d3& o, d3& d) const
    {
        load_ray(tiles.ray_of(item), o, d);
        return true;
    }
    __device__ __forceinline__ void load_ray(int i, d3& o, d3& d) const  // by ray index
    {
        o = ld3(origin + 3 * (int64_t)i);
        d = ld3(dir + 3 * (int64_t)i);
    }
    __device__ __forceinline__ bool same_origin_row(int, int) const { return false; }  // one origin row per ray: compare the data
    __device__ __forceinline__ d3 origin_of(int i) const { return ld3(origin + 3 * (int64_t)i); }
    __device__ __forceinline__ d3 dir_of(int i) const { return ld3(dir + 3 * (int64_t)i); }
    __device__ __forceinline__ void retire(int item, int id, double) const
    {
        const int i = tiles.ray_of(item);
        if (hit1) hit1[i] = id >= 0 ? 1 : 0;
        if (id < 0) write_invalid(out_ori, out_dir, mask3, i);
        int slot = warp_append<>(countL, id >= 0);
        if (slot >= 0) L[slot] = make_int4(i, id, -1, 0);
    }
    // rays first..first+31 all missed: zero their output rows with three bulk copies issued by one lane
    // (87 % of the primary rays of the benchmark views miss; per-lane this would be 10 strided stores each)
    __device__ __forceinline__ bool bulk_miss(int first, unsigned lane)
    {
        if (!zeros || tiles.img_w) return false;  // a tile's rows are four separate runs: plain stores
        if (lane == 0) {
            bulk_store(out_ori + 3 * (int64_t)first, zeros->z, 768);
            bulk_store(out_dir + 3 * (int64_t)first, zeros->z, 768);
            bulk_store(mask3 + 3 * (int64_t)first, zeros->z, 96);
            if (hit1) bulk_store(hit1 + first, zeros->z, 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            issued = true;
        }
        return true;
    }
    __device__ __forceinline__ void finish(unsigned lane)
    {
        // all bulk stores of this lane must have completed before the CTA may exit
        if (lane == 0 && issued) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) wf_q1_kernel(BvhView B, EntryJob job, int N, unsigned long long* work, int policy)
{
    __shared__ ZeroTile zt;
    for (int j = threadIdx.x; j < 768 / 4; j += blockDim.x) reinterpret_cast<unsigned*>(zt.z)[j] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (job.zeros) job.zeros = &zt;
    job.issued = false;
    DRT_QUERY_STACK(stack);
    persistent_query<false>(B, job, N, work, policy, stack);
}

#if DRT_QNODE
// the same entry query with beam culling (beam_pass / entry_query_tiles above): culled tiles are zero-filled by retire()
__global__ void __launch_bounds__(128, 8) wf_beam_kernel(BvhView B, EntryJob job, int N, unsigned long long* work, int tpb, int max_steps,
                                                         int2* __restrict__ tiles, int* __restrict__ n_tiles)
{
    job.zeros = nullptr;
    job.issued = false;
    beam_pass(B, job, N, work, tpb, max_steps, tiles, n_tiles);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) wf_q1_tiles_kernel(BvhView B, EntryJob job, int N, const int2* __restrict__ tiles,
                                                                const int* __restrict__ n_tiles, unsigned long long* work, int policy)
{
    job.zeros = nullptr;
    job.issued = false;
    DRT_QUERY_STACK(stack);
    entry_query_tiles(B, job, N, tiles, n_tiles, work, policy, stack);
}
#endif

// ---- R1: refraction at the entry hit, dense over L ------------------------------------------------
__device__ __forceinline__ void r1_body(const BvhView& B, const double* __restrict__ V64,
                                        const double* __restrict__ origin, const double* __restrict__ dir,
                                        double ext_ior, double int_ior, double* __restrict__ out_ori,
                                        double* __restrict__ out_dir, uint8_t* __restrict__ mask3,
                                        int4* __restrict__ L, const int* countL, const double* __restrict__ VN = nullptr)
{
    const int n = *(volatile const int*)countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        int4 e = L[k];
        const int64_t i = e.x;
        HitRec h;
        d3 a0, a1, a2, o1, d1;
        load_tri64(B, V64, e.y, a0, a1, a2);
        if (VN) {  // optional smooth-normal mode (uniform branch)
            d3 vn[3];
            load_vn(B, VN, e.y, vn);
            hit_forward_t<true>(h, ld3(origin + 3 * i), ld3(dir + 3 * i), a0, a1, a2, vn, ext_ior, int_ior, o1, d1);
        } else
        hit_forward(h, ld3(origin + 3 * i), ld3(dir + 3 * i), a0, a1, a2, ext_ior, int_ior, o1, d1);
        if (h.tir) {
            write_invalid(out_ori, out_dir, mask3, i);
            L[k].w = 1;  // dead
        } else {
            st3(out_ori + 3 * i, o1);  // parked: read back by Q2 and R2
            st3(out_dir + 3 * i, d1);
        }
    }
}

__global__ void __launch_bounds__(128) wf_r1_kernel(BvhView B, const double* __restrict__ V64,
                                                    const double* __restrict__ origin, const double* __restrict__ dir,
                                                    double ext_ior, double int_ior, double* __restrict__ out_ori,
                                                    double* __restrict__ out_dir, uint8_t* __restrict__ mask3,
                                                    int4* __restrict__ L, const int* __restrict__ countL, const double* __restrict__ VN)
{
    r1_body(B, V64, origin, dir, ext_ior, int_ior, out_ori, out_dir, mask3, L, countL, VN);
}

// ---- Q2 ------------------------------------------------------------------------------------------
struct ExitJob {
    static constexpr bool kBulkMiss = false;
    __device__ __forceinline__ bool bulk_miss(int, unsigned) { return false; }
    __device__ __forceinline__ void finish(unsigned) {}
    const double* __restrict__ out_ori;
    const double* __restrict__ out_dir;
    int4* __restrict__ L;
    __device__ __forceinline__ bool load(int k, d3& o, d3& d) const
    {
        int4 e = L[k];
        if (e.w) return false;
        o = ld3(out_ori + 3 * (int64_t)e.x);
        d = ld3(out_dir + 3 * (int64_t)e.x);
        return true;
    }
    __device__ __forceinline__ void retire(int k, int id, double) const { L[k].z = id; }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) wf_q2_kernel(BvhView B, ExitJob job, const int* __restrict__ countL,
                                                    unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    persistent_query<false>(B, job, *countL, work, policy, stack);
}

// ---- R2: refraction at the exit hit, dense over L; survivors -> M ---------------------------------
__device__ __forceinline__ void r2_body(const BvhView& B, const double* __restrict__ V64, double ext_ior,
                                        double int_ior, double* __restrict__ out_ori,
                                        double* __restrict__ out_dir, uint8_t* __restrict__ mask3,
                                        const int4* L, const int* countL, int4* __restrict__ M, int* __restrict__ countM,
                                        const double* __restrict__ VN = nullptr)
{
    const int n = *(volatile const int*)countL;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int4 e = L[k];
        const int64_t i = e.x;
        bool alive = false;
        if (!e.w) {  // dead entries were zeroed by R1
            if (e.z >= 0) {
                HitRec h;
                d3 a0, a1, a2, o2, d2;
                load_tri64(B, V64, e.z, a0, a1, a2);
                if (VN) {
                    d3 vn[3];
                    load_vn(B, VN, e.z, vn);
                    hit_forward_t<true>(h, ld3(out_ori + 3 * i), ld3(out_dir + 3 * i), a0, a1, a2, vn, ext_ior, int_ior, o2, d2);
                } else
                hit_forward(h, ld3(out_ori + 3 * i), ld3(out_dir + 3 * i), a0, a1, a2, ext_ior, int_ior, o2, d2);
                alive = !h.tir;
                if (alive) {
                    st3(out_ori + 3 * i, o2);  // the exit ray, at its final place unless Q3 finds an occluder
                    st3(out_dir + 3 * i, d2);
                }
            }
            if (!alive) write_invalid(out_ori, out_dir, mask3, i);
        }
        int slot = warp_append<>(countM, alive);
        if (slot >= 0) M[slot] = e;
    }
}

__global__ void __launch_bounds__(128) wf_r2_kernel(BvhView B, const double* __restrict__ V64, double ext_ior,
                                                    double int_ior, double* __restrict__ out_ori,
                                                    double* __restrict__ out_dir, uint8_t* __restrict__ mask3,
                                                    const int4* __restrict__ L, const int* __restrict__ countL,
                                                    int4* __restrict__ M, int* __restrict__ countM, const double* __restrict__ VN)
{
    r2_body(B, V64, ext_ior, int_ior, out_ori, out_dir, mask3, L, countL, M, countM, VN);
}

// ---- Q3 ------------------------------------------------------------------------------------------
struct OcclusionJob {
    static constexpr bool kBulkMiss = false;
    __device__ __forceinline__ bool bulk_miss(int, unsigned) { return false; }
    __device__ __forceinline__ void finish(unsigned) {}
    double* __restrict__ out_ori;
    double* __restrict__ out_dir;
    uint8_t* __restrict__ mask3;
    const int4* __restrict__ M;
    int4* __restrict__ rec;
    int* __restrict__ rec_count;
    __device__ __forceinline__ bool load(int k, d3& o, d3& d) const
    {
        const int64_t i = M[k].x;
        o = ld3(out_ori + 3 * i);
        d = ld3(out_dir + 3 * i);
        return true;
    }
    __device__ __forceinline__ void retire(int k, int id, double) const
    {
        const int4 e = M[k];
        const int64_t i = e.x;
        const bool valid = id < 0;
        if (valid) { mask3[3 * i] = 1; mask3[3 * i + 1] = 1; mask3[3 * i + 2] = 1; }
        else write_invalid(out_ori, out_dir, mask3, i);
        if (rec) {
            int slot = warp_append<>(rec_count, valid);
            if (slot >= 0) rec[slot] = e;
        }
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) wf_q3_kernel(BvhView B, OcclusionJob job, const int* __restrict__ countM,
                                                    unsigned long long* work, int policy)
{
    DRT_QUERY_STACK(stack);
    persistent_query<true>(B, job, *countM, work, policy, stack);
}

// ---- all five stages in ONE cooperative launch ------------------------------------------------------
// Same stage bodies, separated by grid-wide barriers instead of kernel boundaries: one launch per
// render_transparent call, no inter-kernel gaps, the list counters never leave the device.
struct FwdArgs {
    BvhView B;
    const double* V64;
    const double* origin;
    const double* dir;
    int N;
    double ext_ior, int_ior;
    double* out_ori;
    double* out_dir;
    uint8_t* mask3;
    uint8_t* hit1;
    int4* L;
    int4* M;
    int4* rec;
    int* rec_count;
    unsigned long long* ctl;  // [0..2] work counters of Q1,Q2,Q3; [3] = {countL, countM}
    int policy[3];  // make_policy(thresh, vote) of Q1, Q2, Q3
    int bulk;
    TileMap tiles;
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) wf_fused_kernel(FwdArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ ZeroTile zt;
    for (int j = threadIdx.x; j < 768 / 4; j += blockDim.x) reinterpret_cast<unsigned*>(zt.z)[j] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    int* countL = reinterpret_cast<int*>(a.ctl + 3);
    int* countM = countL + 1;
    DRT_QUERY_STACK(stack);
    {
        EntryJob j{a.bulk ? &zt : nullptr, false, a.origin, a.dir, a.out_ori, a.out_dir, a.mask3, a.hit1, a.L, countL, a.tiles};
        persistent_query<false>(a.B, j, a.N, a.ctl + 0, a.policy[0], stack);
    }
    grid.sync();
    r1_body(a.B, a.V64, a.origin, a.dir, a.ext_ior, a.int_ior, a.out_ori, a.out_dir, a.mask3, a.L, countL);
    grid.sync();
    {
        ExitJob j{a.out_ori, a.out_dir, a.L};
        persistent_query<false>(a.B, j, *(volatile int*)countL, a.ctl + 1, a.policy[1], stack);
    }
    grid.sync();
    r2_body(a.B, a.V64, a.ext_ior, a.int_ior, a.out_ori, a.out_dir, a.mask3, a.L, countL, a.M, countM);
    grid.sync();
    {
        OcclusionJob j{a.out_ori, a.out_dir, a.mask3, a.M, a.rec, a.rec_count};
        persistent_query<true>(a.B, j, *(volatile int*)countM, a.ctl + 2, a.policy[2], stack);
    }
}

}  // namespace drt
