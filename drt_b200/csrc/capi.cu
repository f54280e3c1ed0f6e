// capi.cu -- extern "C" boundary of libdrt_b200.so (see include/drt_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <memory>

#include "../../include/drt_b200.h"
#include "bvh_coop.cuh"
#include "plane.cuh"
#include "loss_step.cuh"
#include "peer_allreduce.cuh"
#include "silhouette.cuh"

using namespace drt;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};  // kernels of this library launched so far (bench.py's gpu_launches)

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(DRT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int ensure(T*& p, size_t& cap, size_t need)
{
    if (need <= cap) return DRT_OK;
    if (p) CU(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t n = need + need / 4 + 64;
    CU(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
    return DRT_OK;
}

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

constexpr int kWorkSlots = 256;
constexpr int64_t kSimpleMaxRays = 1 << 20;
// rays per vertex above which the backward kernels merge equal-triangle runs in the warp before their atomics
// (C3, 10 760 rays per vertex: 1.28 -> 0.85 ms; C4, 1 980: 0.57 -> 0.55 ms with tile-ordered records)
constexpr int64_t kMergeRaysPerVertex = 1500;

// tuning knobs (read once): DRT_FWD_KERNEL = auto (default: by batch size) | wavefront | simple ; DRT_FWD_THRESH = 1..32
struct Tuning {
    bool simple_fwd = false;
    bool force_wavefront = false;
    bool one_launch = false;  // DRT_ONE_LAUNCH=1: the cooperative single-launch variant (measured 1.5 % slower: spills)
    bool prefer_l1 = false;  // DRT_PREFER_L1=1 forces cudaSharedmemCarveoutMaxL1: measured 22 % SLOWER (the 1 KB/block reserve then caps residency at 4 blocks/SM)
    int bwd_merge = -1;  // DRT_BWD_MERGE = 0 | 1 forces the run-merged backward scatter off / on (default: by rays per vertex)
    int tile_w_log2 = 0;  // DRT_TILE_SHAPE = 4x8 | 8x4 | 16x2 forces a shape (default: 4x8 for the fused step, 8x4 with dense outputs)
    int r_grid = 8;       // DRT_R_GRID: blocks per SM of the dense refraction kernels' grids
    bool tile = true;  // DRT_TILE=0: keep scanline batches in drt_ray_loss_step even when the image size is known (A/B switch)
    bool bulk = true;  // DRT_BULK_ZERO=0 disables the TMA bulk zero-fill of missed rays (A/B switch)
    bool coop_build = false;  // DRT_COOP_BUILD=1: the LBVH build as ONE cooperative launch (bvh_coop.cuh) -- measured SLOWER on B200 (0.208 vs 0.170 ms at 50 k triangles: 15 grid-wide barriers at ~5 us each and L2-only reads cost more than 18 launch boundaries), kept as a tested option
    bool beam = true;  // DRT_BEAM=0: entry query without beam culling of whole pixel tiles (A/B switch)
    int lanes_big = 4;  // lanes of batches above lanes_max_rays (DRT_LANES_BIG; 1 = none)
    int lanes = 2;  // DRT_LANES=1: drt_ray_loss_step on the caller's stream only (no split); =2..4 forces that many lanes at any size
    bool lanes_forced = false;
    int64_t lanes_max_rays = 16 << 20;  // default: two lanes for batches up to 16 M rays
    int64_t direct_max_rays = 1250000;  // DRT_DIRECT_MAX / drt_tuning_set("direct_max_rays"): measured on B200 -- C2 (262 k rays) step 0.342 -> 0.261 ms, one 960x720 view (691 k) 0.43 -> 0.365 (C4 mesh) / 0.395 -> 0.364 (C3), one 960x1280 view on mouse_vh (1.23 M) 0.547 -> 0.499, two 960x720 views (1.38 M) 0.510 -> 0.555, three (2.1 M) 0.59 -> 0.73: batches up to this many rays run the one-thread-per-path forward (ls_direct_kernel)
    int beam_steps = 1 << 30;  // DRT_BEAM_STEPS: node steps after which an undecided beam is kept
    int beam_tpb = 0;  // DRT_BEAM_TPB = 1..32 forces the tiles a warp takes per work fetch (default: by batch size)
    int thresh = 32;
    int pol[3];  // make_policy(thresh, vote) of the three query stages: DRT_FWD_THRESH / DRT_VOTE, per stage DRT_THRESH_Q1.. / DRT_VOTE_Q1..
    int minb = 8;
    Tuning()
    {
        const char* k = getenv("DRT_FWD_KERNEL");
        if (k && !strcmp(k, "simple")) simple_fwd = true;
        if (k && !strcmp(k, "wavefront")) force_wavefront = true;

        const char* t = getenv("DRT_FWD_THRESH");
        if (t && atoi(t) >= 1 && atoi(t) <= 32) thresh = atoi(t);
        const char* vt = getenv("DRT_VOTE");
        // round 1 (one vote per node step): 0: 6.68, 4: 6.37, 8: 6.43, 12: 6.56 ms forward at C4; with one vote per 8 steps (trace.cuh:
        // DRT_VOTE_EVERY) the best threshold is 1 -- leave the walk as soon as any lane is blocked when the vote comes
        const int vote = (vt && atoi(vt) >= 0 && atoi(vt) <= 31) ? atoi(vt) : 1;
        for (int q = 0; q < 3; ++q) {
            char name[32];
            int th = thresh, vo = vote;
            snprintf(name, sizeof name, "DRT_THRESH_Q%d", q + 1);
            const char* a = getenv(name);
            if (a && atoi(a) >= 1 && atoi(a) <= 32) th = atoi(a);
            snprintf(name, sizeof name, "DRT_VOTE_Q%d", q + 1);
            const char* c = getenv(name);
            if (c && atoi(c) >= 0 && atoi(c) <= 31) vo = atoi(c);
            pol[q] = make_policy(th, vo);
        }
        const char* ts = getenv("DRT_TILE_SHAPE");
        if (ts && !strcmp(ts, "8x4")) tile_w_log2 = 3;
        if (ts && !strcmp(ts, "4x8")) tile_w_log2 = 2;
        if (ts && !strcmp(ts, "16x2")) tile_w_log2 = 4;
        const char* rg = getenv("DRT_R_GRID");
        if (rg && atoi(rg) >= 1 && atoi(rg) <= 64) r_grid = atoi(rg);
        const char* tl = getenv("DRT_TILE");
        if (tl && !strcmp(tl, "0")) tile = false;
        const char* ol = getenv("DRT_ONE_LAUNCH");
        if (ol && !strcmp(ol, "1")) one_launch = true;
        const char* pl = getenv("DRT_PREFER_L1");
        if (pl && !strcmp(pl, "1")) prefer_l1 = true;
        const char* bm = getenv("DRT_BWD_MERGE");
        if (bm && (!strcmp(bm, "0") || !strcmp(bm, "1"))) bwd_merge = atoi(bm);
        const char* cb = getenv("DRT_COOP_BUILD");
        if (cb && !strcmp(cb, "1")) coop_build = true;
        const char* bm2 = getenv("DRT_BEAM");
        if (bm2 && !strcmp(bm2, "0")) beam = false;
        const char* ln = getenv("DRT_LANES");
        if (ln && atoi(ln) >= 1 && atoi(ln) <= 8) { lanes = atoi(ln); lanes_forced = lanes >= 2; }
        const char* dm = getenv("DRT_DIRECT_MAX");
        if (dm && atoll(dm) >= 0) direct_max_rays = atoll(dm);
        const char* lb = getenv("DRT_LANES_BIG");
        if (lb && atoi(lb) >= 1 && atoi(lb) <= 8) lanes_big = atoi(lb);
        const char* bs = getenv("DRT_BEAM_STEPS");
        if (bs && atoi(bs) >= 1) beam_steps = atoi(bs);
        const char* tp = getenv("DRT_BEAM_TPB");
        if (tp && atoi(tp) >= 1 && atoi(tp) <= 32) beam_tpb = atoi(tp);
        const char* z = getenv("DRT_BULK_ZERO");
        if (z && !strcmp(z, "0")) bulk = false;
        const char* m = getenv("DRT_Q_MINB");
        if (m) minb = atoi(m);
        if (minb != 4 && minb != 6 && minb != 7 && minb != 8 && minb != 10) minb = 8;
    }
};
Tuning& tuning_mut()
{
    static Tuning t;
    return t;
}
const Tuning& tuning() { return tuning_mut(); }

}  // namespace

struct drt_bvh {
    int device = 0;
    int sm_count = 148;
    int nF = 0, nV = 0;
    bool built = false;
    int64_t builds = 0, refits = 0;
    // geometry copies
    int32_t* F = nullptr;   size_t capF = 0;     // [nF*3]
    float* V32 = nullptr;   size_t capV = 0;     // [nV*3]
    // build scratch
    uint64_t* keys = nullptr;  size_t capK = 0;  // 2*nF (double buffer)
    unsigned* sort_table = nullptr; size_t capSt = 0; // 256 x tiles digit counters of the radix sort
    int2* children = nullptr;  size_t capCh = 0; // nF-1
    int* parent = nullptr;     size_t capP = 0;  // 2nF-1
    float4* blo = nullptr;     size_t capBl = 0; // 2nF-1
    float4* bhi = nullptr;     size_t capBh = 0;
    int* flags = nullptr;      size_t capFl = 0; // nF-1
    unsigned* scene = nullptr;                   // 6 encoded floats + 1 int (bad index count) + pad
    int4* listA = nullptr;     size_t capLA = 0; // wavefront list L (ray, tri1, tri2, dead) of the rays that hit
    int4* listB = nullptr;     size_t capLB = 0; // wavefront list M: entries of L that survive both refractions
    double* park = nullptr;    size_t capPk = 0; // loss step: parked rays, 6 component columns of capPk/6 slots
    int2* listM = nullptr;     size_t capLM = 0; // loss step: (slot of L, target slot) of the rays that survive both refractions
    int4* listS = nullptr;     size_t capLS = 0; // loss step: (ray, tri1, tri2, target slot) of the valid paths; before Q1: the beam pass's tile list
    int* tbucket = nullptr;    size_t capTb = 0; // loss step: bucket table of the sparse screen targets
    unsigned long long* work = nullptr;          // ring of work counters of the persistent tracer
    cudaStream_t lane_stream[8] = {};            // internal streams of the lanes of drt_ray_loss_step
    cudaEvent_t lane_event[17] = {};             // fork, forward done x8, done x8
    int last_lanes = 0;                          // lanes of the latest drt_ray_loss_step (0: the one-thread-per-path route)
    unsigned long long* last_ctl[8] = {};        // control blocks (one per lane) of the latest drt_ray_loss_step (drt_bvh_last_counts)
    int64_t last_tiles = 0;                      // its number of 32-ray tiles (0: no beam pass)
    int work_slot = 0;
    int build_blocks_per_sm = 0;                 // co-resident blocks of lbvh_build_kernel (0: no cooperative launch)
    int fused_blocks_per_sm = 0;                 // co-resident blocks of wf_fused_kernel<8> (0: no cooperative launch)
    int fused6_blocks_per_sm = 0;                // same for the 80-register variant
    int img_w = 0, img_h = 0;                    // drt_bvh_set_image_size: rays of drt_trace_fwd are whole scanline images
    int bwd_blocks[3] = {8, 3, 3};               // co-resident blocks per SM of ls_loss_bwd_kernel<loss only | grad | grad+merge>
    uint64_t* sorted_keys = nullptr;             // the half of `keys` that holds the sorted (Morton, id) keys
    // traversal data
    node_quad* nodes = nullptr; size_t capN = 0;
    uint4* nodes4 = nullptr;   size_t capN4 = 0;  // 4-wide view of the same tree (DRT_BVH4)
    double2* tris = nullptr;   size_t capT = 0;

    BvhView view() const { return BvhView{nodes, nodes4, tris, F, scene, nF}; }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        int cur;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// device that owns a device pointer (the handle-less entry points launch there, not on the caller's current device)
cudaError_t device_of(const void* p, int* dev)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) return e;
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) { *dev = a.device; return cudaSuccess; }
    return cudaGetDevice(dev);
}

// 32-pixel tiles are possible when the N rays are whole images whose sides a tile shape divides
TileMap tile_map(int img_w, int img_h, int64_t N, bool dense_outputs)
{
    const bool whole = tuning().tile && img_w > 0 && img_h > 0 && (int64_t)img_w * img_h <= N && N % ((int64_t)img_w * img_h) == 0;
    // preferred shape first, then the other one.  Fused step: 4 x 8 pixels (C4 forward 6.17 ms; 8 x 4: 6.27, 16 x 2: 6.59, 32 x 1
    // strips: 7.43).  Dense outputs: 8 x 4 (6.67 vs 6.90 ms: the rows a warp zero-fills for its misses are 192-byte runs instead of 96)
    const int first = tuning().tile_w_log2 ? tuning().tile_w_log2 : (dense_outputs ? 3 : 2);
    const int shapes[2] = {first, first == 2 ? 3 : 2};
    for (int lg : shapes)
        if (whole && img_w % (1 << lg) == 0 && img_h % (32 >> lg) == 0) return TileMap{img_w, img_w * img_h, lg};
    return TileMap{0, 0, 3};
}

// tiles (32 rays) a warp of the beam-culling entry query takes per work fetch: up to 32 (one beam per lane), fewer for
// small batches so that every warp still gets several fetches (the tail of a persistent kernel is one fetch long)
int beam_tiles_per_fetch(int64_t N, int warps)
{
    if (tuning().beam_tpb) return tuning().beam_tpb;
    const int64_t tiles = (N + 31) / 32;
    int tpb = 32;
    while (tpb > 4 && tiles / tpb < (int64_t)warps * 4) tpb >>= 1;
    return tpb;
}

constexpr int kMaxLanes = 8;

struct LaneCounts {
    const int* p[kMaxLanes];
};
__global__ void sum_counts_kernel(LaneCounts c, int n, int32_t* out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < n; ++k) s += *c.p[k];
        *out = s;
    }
}

// clamp + count out-of-range indices so that no later kernel can fault on a bad face list
__global__ void copy_faces_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n3, int nV, int* __restrict__ bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    int v = in[i];
    if (v < 0 || v >= nV) { atomicAdd(bad, 1); v = min(max(v, 0), nV - 1); }
    out[i] = v;
}

__global__ void init_scene_kernel(unsigned* scene, bool reset_bad)
{
    if (threadIdx.x < 3) scene[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6 || threadIdx.x == 7) scene[threadIdx.x] = 0u;
    else if (threadIdx.x == 6 && reset_bad) scene[6] = 0u;
}

// the whole (re)build, or the refit, as ONE cooperative launch (bvh_coop.cuh); V64 != nullptr: cast the vertices first
int coop_build(drt_bvh* b, const double* V64, int refit, cudaStream_t st)
{
    const int n = b->nF;
    BuildArgs a{b->F, b->V32, V64, b->nV, n, b->keys, b->sort_table, b->children, b->parent, b->blo, b->bhi, b->flags, b->scene, b->nodes, b->nodes4, b->tris, refit};
    const int64_t work = std::max<int64_t>(n, V64 ? 3 * (int64_t)b->nV : 0);
    // a small grid keeps the grid-wide barriers short: one block per SM at most
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_for(work, kSortThreads), std::min(b->sm_count, b->sm_count * b->build_blocks_per_sm)));
    void* args[] = {&a};
    CU(cudaLaunchCooperativeKernel((void*)lbvh_build_kernel, dim3(grid), dim3(kSortThreads), args, 0, st));
    ++g_launches;
    b->sorted_keys = b->keys + n;
    return DRT_OK;
}

int fit_and_emit(drt_bvh* b, cudaStream_t st, const double* pending_V64 = nullptr)
{
    const int n = b->nF;
    if (n > 0 && DRT_QNODE && tuning().coop_build && b->build_blocks_per_sm > 0 && b->sorted_keys) return coop_build(b, pending_V64, 1, st);
    if (pending_V64) {
        cast_vertices_kernel<<<blocks_for(3 * (int64_t)b->nV, 256), 256, 0, st>>>(pending_V64, b->V32, 3 * b->nV);
        ++g_launches;
    }
    if (n <= 0) return DRT_OK;
    if (n > 1) CU(cudaMemsetAsync(b->flags, 0, sizeof(int) * (size_t)(n - 1), st));
    fit_kernel<<<blocks_for(n, 256), 256, 0, st>>>(b->F, b->V32, b->sorted_keys, n, b->children, b->parent, b->blo, b->bhi,
                                                    b->flags); ++g_launches;
#if DRT_QNODE
    emit_all_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, b->F, b->V32, b->sorted_keys, b->children, b->blo, b->bhi, b->scene, b->nodes,
                                                        DRT_BVH4 ? b->nodes4 : nullptr, b->tris);
    ++g_launches;
#else
    grid_kernel<<<1, 32, 0, st>>>(b->blo, b->bhi, b->scene); ++g_launches;
    emit_nodes_kernel<<<blocks_for(n > 1 ? n - 1 : 1, 256), 256, 0, st>>>(n, b->children, b->blo, b->bhi, b->scene, b->nodes); ++g_launches;
#if DRT_QNODE && DRT_BVH4
    emit_nodes4_kernel<<<blocks_for(n > 1 ? n - 1 : 1, 256), 256, 0, st>>>(n, b->children, b->blo, b->bhi, b->scene, b->nodes4); ++g_launches;
#endif
    emit_tris_kernel<<<blocks_for(n, 256), 256, 0, st>>>(b->F, b->V32, b->sorted_keys, n, b->tris); ++g_launches;
#endif
    CU(cudaGetLastError());
    return DRT_OK;
}

// pending_V64: float64 vertices whose float32 cast (into b->V32) is still to be done -- folded into the cooperative kernel
int build_tree(drt_bvh* b, cudaStream_t st, const double* pending_V64 = nullptr)
{
    const int n = b->nF;
    b->built = true;
    b->builds++;
    if (n <= 0) {
        if (pending_V64 && b->nV > 0) {
            cast_vertices_kernel<<<blocks_for(3 * (int64_t)b->nV, 256), 256, 0, st>>>(pending_V64, b->V32, 3 * b->nV);
            ++g_launches;
        }
        return DRT_OK;
    }
    int rc;
    if ((rc = ensure(b->keys, b->capK, 2 * (size_t)n))) return rc;
    if ((rc = ensure(b->sort_table, b->capSt, sort_table_words(n, kIndexBits)))) return rc;
    if ((rc = ensure(b->children, b->capCh, (size_t)n))) return rc;
    if ((rc = ensure(b->parent, b->capP, 2 * (size_t)n))) return rc;
    if ((rc = ensure(b->blo, b->capBl, 2 * (size_t)n))) return rc;
    if ((rc = ensure(b->bhi, b->capBh, 2 * (size_t)n))) return rc;
    if ((rc = ensure(b->flags, b->capFl, (size_t)n))) return rc;
    if ((rc = ensure(b->nodes, b->capN, (size_t)kNodeQuads * (size_t)(n > 1 ? n - 1 : 1)))) return rc;
    if ((rc = ensure(b->tris, b->capT, (size_t)kTriD2 * (size_t)n))) return rc;
#if DRT_QNODE && DRT_BVH4
    if ((rc = ensure(b->nodes4, b->capN4, 4 * (size_t)(n > 1 ? n - 1 : 1)))) return rc;
#endif
    if (DRT_QNODE && tuning().coop_build && b->build_blocks_per_sm > 0) return coop_build(b, pending_V64, 0, st);
    if (pending_V64) {
        cast_vertices_kernel<<<blocks_for(3 * (int64_t)b->nV, 256), 256, 0, st>>>(pending_V64, b->V32, 3 * b->nV);
        ++g_launches;
    }

    init_scene_kernel<<<1, 32, 0, st>>>(b->scene, false); ++g_launches;
    CU(cudaMemsetAsync(b->sort_table, 0, sort_table_words(n, kIndexBits) * sizeof(unsigned), st));  // the digit tables of all sort passes
    centroid_bounds_kernel<<<blocks_for(n, 256), 256, 0, st>>>(b->F, b->V32, n, b->scene); ++g_launches;
    morton_kernel<<<blocks_for(n, 256), 256, 0, st>>>(b->F, b->V32, n, b->scene, b->keys, b->sort_table, kIndexBits & ~7, sort_tiles(n)); ++g_launches;
    // unique keys (Morton << 25 | id): keys-only LSD radix sort over the Morton bits, one launch per pass
    b->sorted_keys = sort_keys_u64(b->keys, n, kIndexBits, b->sort_table, true, st, &g_launches);
    if (n > 1) {
        topology_kernel<<<blocks_for(n - 1, 256), 256, 0, st>>>(b->sorted_keys, n, b->children, b->parent);
        ++g_launches;
    }
    CU(cudaGetLastError());
    return fit_and_emit(b, st);  // (the multi-launch path: vertices were cast above)
}

// float32 vertices are copied now; float64 ones are cast by whoever builds next (build_tree / fit_and_emit: pending_V64)
int set_vertices(drt_bvh* b, const float* V32, const double* V64, int nV, cudaStream_t st)
{
    int rc;
    if ((rc = ensure(b->V32, b->capV, 3 * (size_t)(nV > 0 ? nV : 1)))) return rc;
    b->nV = nV;
    if (nV <= 0) return DRT_OK;
    if (V32) CU(cudaMemcpyAsync(b->V32, V32, sizeof(float) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, st));
    (void)V64;
    CU(cudaGetLastError());
    return DRT_OK;
}

int build_common(drt_bvh* b, const int32_t* F, int nF, const float* V32, const double* V64, int nV, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_bvh_build: null handle");
    if (nF < 0 || nV < 0) return fail(DRT_ERR_INVALID, "drt_bvh_build: negative size (nF=%d, nV=%d)", nF, nV);
    if (nF > (1 << kIndexBits)) return fail(DRT_ERR_INVALID, "drt_bvh_build: %d triangles exceed the supported %d", nF, 1 << kIndexBits);
    if ((nF > 0 && !F) || (nV > 0 && !V32 && !V64)) return fail(DRT_ERR_INVALID, "drt_bvh_build: null F or V");
    if (nF > 0 && nV == 0) return fail(DRT_ERR_INVALID, "drt_bvh_build: faces without vertices");
    DeviceGuard g(b->device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", b->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = set_vertices(b, V32, V64, nV, st))) return rc;
    if ((rc = ensure(b->F, b->capF, 3 * (size_t)(nF > 0 ? nF : 1)))) return rc;
    b->nF = nF;
    init_scene_kernel<<<1, 32, 0, st>>>(b->scene, true); ++g_launches;
    if (nF > 0) {
        copy_faces_kernel<<<blocks_for(3 * (int64_t)nF, 256), 256, 0, st>>>(F, b->F, 3 * nF, nV, (int*)(b->scene + 6));
        ++g_launches;
    }
    CU(cudaGetLastError());
    return build_tree(b, st, V32 ? nullptr : V64);
}

}  // namespace

extern "C" {

int drt_version(void) { return 1002; }

int drt_tuning_set(const char* key, long long value)
{
    if (!key || value < 0) return fail(DRT_ERR_INVALID, "drt_tuning_set: null key or negative value");
    if (!strcmp(key, "direct_max_rays")) { tuning_mut().direct_max_rays = value; return DRT_OK; }
    return fail(DRT_ERR_INVALID, "drt_tuning_set: unknown key");
}

long long drt_tuning_get(const char* key)
{
    if (key && !strcmp(key, "direct_max_rays")) return tuning().direct_max_rays;
    return -1;
}

unsigned long long drt_kernel_launches(void) { return g_launches; }

const char* drt_last_error(void) { return g_err; }

int drt_bvh_create(int device, drt_bvh** out)
{
    if (!out) return fail(DRT_ERR_INVALID, "drt_bvh_create: out is null");
    *out = nullptr;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(DRT_ERR_INVALID, "drt_bvh_create: device %d out of range (%d visible)", device, count);
    DeviceGuard g(device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    std::unique_ptr<drt_bvh, int (*)(drt_bvh*)> guard(new drt_bvh(), drt_bvh_destroy);  // freed on every early return
    drt_bvh* b = guard.get();
    b->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    b->sm_count = prop.multiProcessorCount;
    CU(cudaMalloc(&b->scene, kSceneWords * sizeof(unsigned)));
    CU(cudaMemset(b->scene, 0, kSceneWords * sizeof(unsigned)));
    CU(cudaMalloc(&b->work, kWorkSlots * sizeof(unsigned long long)));
    if (tuning().prefer_l1) {
        // traversal kernels use < 1 KB of shared memory: ask for the largest L1 split explicitly
        const int carve = cudaSharedmemCarveoutMaxL1;
        CU(cudaFuncSetAttribute(wf_q1_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(wf_q2_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(wf_q3_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(wf_q1_kernel<7>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(wf_q2_kernel<7>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(wf_q3_kernel<7>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(closest_hit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CU(cudaFuncSetAttribute(trace_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    {
        int coop = 0;
        CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        if (coop) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->build_blocks_per_sm, lbvh_build_kernel, kSortThreads, 0));
        if (coop) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->fused_blocks_per_sm, wf_fused_kernel<8>, 128, 0));
        if (coop) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->fused6_blocks_per_sm, wf_fused_kernel<6>, 128, 0));
    }

    for (int k = 0; k < kMaxLanes; ++k) CU(cudaStreamCreateWithFlags(&b->lane_stream[k], cudaStreamNonBlocking));
    for (int k = 0; k < 1 + 2 * kMaxLanes; ++k) CU(cudaEventCreateWithFlags(&b->lane_event[k], cudaEventDisableTiming));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->bwd_blocks[0], ls_loss_bwd_kernel<false, false>, 128, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->bwd_blocks[1], ls_loss_bwd_kernel<true, false>, 128, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b->bwd_blocks[2], ls_loss_bwd_kernel<true, true>, 128, 0));
    *out = guard.release();
    return DRT_OK;
}

int drt_bvh_destroy(drt_bvh* b)
{
    if (!b) return DRT_OK;
    DeviceGuard g(b->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {b->listA, b->listB, b->park, b->listM, b->listS, b->tbucket, b->work, b->F, b->V32, b->keys, b->sort_table, b->children, b->parent, b->blo, b->bhi, b->flags, b->scene, b->nodes, b->nodes4, b->tris};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (cudaStream_t q : b->lane_stream)
        if (q) cudaStreamDestroy(q);
    for (cudaEvent_t e : b->lane_event)
        if (e) cudaEventDestroy(e);
    delete b;
    return DRT_OK;
}

int drt_bvh_build(drt_bvh* b, const int32_t* F, int32_t nF, const float* V32, int32_t nV, void* stream)
{
    if (nV > 0 && !V32) return fail(DRT_ERR_INVALID, "drt_bvh_build: V32 is null");
    return build_common(b, F, nF, V32, nullptr, nV, stream);
}

int drt_bvh_build_f64(drt_bvh* b, const int32_t* F, int32_t nF, const double* V64, int32_t nV, void* stream)
{
    if (nV > 0 && !V64) return fail(DRT_ERR_INVALID, "drt_bvh_build_f64: V64 is null");
    return build_common(b, F, nF, nullptr, V64, nV, stream);
}

int drt_bvh_update_vert(drt_bvh* b, const float* V32, const double* V64, int32_t nV, int refit, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_bvh_update_vert: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_bvh_update_vert: update_mesh has not been called (no faces)");
    if ((V32 == nullptr) == (V64 == nullptr)) return fail(DRT_ERR_INVALID, "drt_bvh_update_vert: exactly one of V32/V64 must be given");
    if (nV != b->nV) return fail(DRT_ERR_INVALID, "drt_bvh_update_vert: vertex count %d != %d of the current faces", nV, b->nV);
    DeviceGuard g(b->device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", b->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = set_vertices(b, V32, V64, nV, st))) return rc;
    if (refit) { b->refits++; return fit_and_emit(b, st, V64); }
    return build_tree(b, st, V64);
}

int drt_bvh_set_image_size(drt_bvh* b, int32_t image_w, int32_t image_h)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_bvh_set_image_size: null handle");
    if (image_w < 0 || image_h < 0) return fail(DRT_ERR_INVALID, "drt_bvh_set_image_size: negative size");
    b->img_w = image_w;
    b->img_h = image_h;
    return DRT_OK;
}

int drt_bvh_info(const drt_bvh* b, int64_t info[8])
{
    if (!b || !info) return fail(DRT_ERR_INVALID, "drt_bvh_info: null argument");
    info[0] = b->nF; info[1] = b->nV; info[2] = b->nF > 1 ? b->nF - 1 : (b->nF == 1 ? 1 : 0);
    info[3] = b->built ? 1 : 0;
    info[4] = info[2] * (int64_t)(kNodeQuads * sizeof(node_quad));
    info[5] = (int64_t)b->nF * (int64_t)(kTriD2 * sizeof(double2));
    info[6] = b->builds; info[7] = b->refits;
    return DRT_OK;
}

int drt_bvh_bad_indices(const drt_bvh* b, void* stream, int* out)
{
    if (!b || !out) return fail(DRT_ERR_INVALID, "drt_bvh_bad_indices: null argument");
    DeviceGuard g(b->device);
    CU(cudaMemcpyAsync(out, b->scene + 6, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return DRT_OK;
}

int drt_bvh_last_counts(const drt_bvh* b, void* stream, int64_t out[6])
{
    if (!b || !out) return fail(DRT_ERR_INVALID, "drt_bvh_last_counts: null argument");
    for (int k = 0; k < 6; ++k) out[k] = 0;
    if (!b->last_ctl[0]) return DRT_OK;
    DeviceGuard g(b->device);
    unsigned long long h[kMaxLanes][8] = {};
    for (int k = 0; k < kMaxLanes; ++k)
        if (b->last_ctl[k]) CU(cudaMemcpyAsync(h[k], b->last_ctl[k], sizeof h[k], cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    for (int k = 0; k < kMaxLanes; ++k) {
        const int* c = reinterpret_cast<const int*>(h[k]);
        out[0] += c[6]; out[1] += c[7]; out[2] += c[8];   // entry hits, survivors of both refractions, valid paths
        out[4] += b->last_tiles ? c[10] : 0;              // tiles kept by the beam pass
    }
    out[3] = b->last_tiles;
    out[5] = b->last_lanes;
    return DRT_OK;
}

int drt_closest_hit(const drt_bvh* b, const float* ray6, int64_t N, float* T, int32_t* ID, int64_t strideT, int64_t strideID, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_closest_hit: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_closest_hit: no mesh has been set (update_mesh first)");
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_closest_hit: N < 0");
    if (N == 0) return DRT_OK;
    if (!ray6 || !T || !ID) return fail(DRT_ERR_INVALID, "drt_closest_hit: null buffer");
    if (strideT < 1 || strideID < 1) return fail(DRT_ERR_INVALID, "drt_closest_hit: strides must be >= 1");
    if (((uintptr_t)ray6 & 7u) != 0) return fail(DRT_ERR_INVALID, "drt_closest_hit: ray6 must be 8-byte aligned");
    DeviceGuard g(b->device);
    int grid = (int)std::min<int64_t>(blocks_for(N, 256), (int64_t)b->sm_count * 32);
    closest_hit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b->view(), ray6, N, T, ID, strideT, strideID); ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

static int trace_fwd_impl(drt_bvh* b, const double* V64, const double* VN, const double* origin, const double* dir, int64_t N, double ext_ior,
                          double int_ior, double* out_ori, double* out_dir, uint8_t* mask3, int32_t* rec, int32_t* rec_count,
                          uint8_t* hit1, void* stream);

int drt_trace_fwd(drt_bvh* b, const double* V64, const double* origin, const double* dir, int64_t N, double ext_ior,
                  double int_ior, double* out_ori, double* out_dir, uint8_t* mask3, int32_t* rec, int32_t* rec_count,
                  uint8_t* hit1, void* stream)
{
    return trace_fwd_impl(b, V64, nullptr, origin, dir, N, ext_ior, int_ior, out_ori, out_dir, mask3, rec, rec_count, hit1, stream);
}

int drt_trace_fwd_smooth(drt_bvh* b, const double* V64, const double* VN64, const double* origin, const double* dir, int64_t N,
                         double ext_ior, double int_ior, double* out_ori, double* out_dir, uint8_t* mask3, int32_t* rec,
                         int32_t* rec_count, void* stream)
{
    if (b && b->nF > 0 && N > 0 && !VN64) return fail(DRT_ERR_INVALID, "drt_trace_fwd_smooth: VN64 is null");
    return trace_fwd_impl(b, V64, VN64, origin, dir, N, ext_ior, int_ior, out_ori, out_dir, mask3, rec, rec_count, nullptr, stream);
}

// VN != nullptr: optional smooth-normal mode (vertex normals float64 [nV,3]); always the five-launch wavefront
static int trace_fwd_impl(drt_bvh* b, const double* V64, const double* VN, const double* origin, const double* dir, int64_t N, double ext_ior,
                          double int_ior, double* out_ori, double* out_dir, uint8_t* mask3, int32_t* rec, int32_t* rec_count,
                          uint8_t* hit1, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_trace_fwd: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_trace_fwd: no mesh has been set (update_mesh first)");
    if (N < 0 || N > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_trace_fwd: N must be in [0, 2^31)");
    if ((rec == nullptr) != (rec_count == nullptr)) return fail(DRT_ERR_INVALID, "drt_trace_fwd: rec/rec_count must both be given or both be null");
    if (rec && ((uintptr_t)rec & 15u)) return fail(DRT_ERR_INVALID, "drt_trace_fwd: rec must be 16-byte aligned");
    DeviceGuard g(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (rec_count) CU(cudaMemsetAsync(rec_count, 0, sizeof(int32_t), st));
    if (N == 0) return DRT_OK;
    if (!origin || !dir || !out_ori || !out_dir || !mask3) return fail(DRT_ERR_INVALID, "drt_trace_fwd: null buffer");
    if (b->nF > 0 && !V64) return fail(DRT_ERR_INVALID, "drt_trace_fwd: V64 is null");
    // small batches: the five-launch wavefront has ~0.1 ms of fixed cost, the one-launch megakernel wins below ~1 M rays
    if (!VN && (tuning().simple_fwd || (!tuning().force_wavefront && N <= kSimpleMaxRays))) {
        int grid = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * 64);
        trace_fwd_kernel<<<grid, 128, 0, st>>>(b->view(), V64, origin, dir, N, ext_ior, int_ior, out_ori, out_dir, mask3,
                                               (int4*)rec, rec_count, hit1, tile_map(b->img_w, b->img_h, N, true));
    } else {
        // production path: wavefront of persistent query kernels (wavefront.cuh)
        int rc;
        if ((rc = ensure(b->listA, b->capLA, (size_t)N))) return rc;
        if ((rc = ensure(b->listB, b->capLB, (size_t)N))) return rc;
        // per-launch control block: 3 work counters (Q1,Q2,Q3) + {countL, countM} + - + {tiles kept} + tile-list work counter
        unsigned long long* ctl = b->work + (size_t)(b->work_slot++ % (kWorkSlots / 8)) * 8;
        CU(cudaMemsetAsync(ctl, 0, 8 * sizeof(unsigned long long), st));
        int* countL = (int*)(ctl + 3);
        int* countM = countL + 1;
        const int* pol = tuning().pol;
        // bulk zero-fill needs 16-byte aligned output rows for every multiple-of-32 ray index
        const bool bulk_ok = tuning().bulk && !(((uintptr_t)out_ori | (uintptr_t)out_dir | (uintptr_t)mask3 | (uintptr_t)hit1) & 15u);
        const int dgrid = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * 8);
        EntryJob j1{bulk_ok ? reinterpret_cast<const ZeroTile*>(1) : nullptr, false, origin, dir, out_ori, out_dir, mask3, hit1, b->listA, countL,
                    tile_map(b->img_w, b->img_h, N, true)};
        const int minb = tuning().minb;
        const int pg = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * minb);
        if (!VN && tuning().one_launch && b->fused_blocks_per_sm > 0) {
            // the whole wavefront as ONE cooperative launch (grid-wide barriers between the stages)
            FwdArgs fa{b->view(), V64, origin, dir, (int)N, ext_ior, int_ior, out_ori, out_dir, mask3, hit1, b->listA, b->listB,
                       (int4*)rec, rec_count, ctl, {pol[0], pol[1], pol[2]}, bulk_ok ? 1 : 0, tile_map(b->img_w, b->img_h, N, true)};
            void* kargs[] = {&fa};
            const bool six = minb == 6 && b->fused6_blocks_per_sm > 0;
            const int per_sm = six ? b->fused6_blocks_per_sm : b->fused_blocks_per_sm;
            const int cg_grid = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * per_sm);
            CU(cudaLaunchCooperativeKernel(six ? (void*)wf_fused_kernel<6> : (void*)wf_fused_kernel<8>, dim3(cg_grid), dim3(128),
                                           kargs, 0, st));
            ++g_launches;
            CU(cudaGetLastError());
            return DRT_OK;
        }
#define DRT_LAUNCH_Q(KERNEL, ...)                                                           \
    do {                                                                                    \
        if (minb == 10) KERNEL<10><<<pg, 128, 0, st>>>(__VA_ARGS__);                        \
        else if (minb == 8) KERNEL<8><<<pg, 128, 0, st>>>(__VA_ARGS__);                     \
        else if (minb == 7) KERNEL<7><<<pg, 128, 0, st>>>(__VA_ARGS__);                     \
        else if (minb == 6) KERNEL<6><<<pg, 128, 0, st>>>(__VA_ARGS__);                     \
        else KERNEL<4><<<pg, 128, 0, st>>>(__VA_ARGS__);                                    \
    } while (0)
#if DRT_QNODE
        if (tuning().beam && (pol[0] & 0xff) == 32) {
            // beam pass over all 32-ray tiles (culled tiles are zero-filled), then the per-ray entry query over the survivors;
            // the tile list lives in listB, which is free until R2
            int2* tiles = reinterpret_cast<int2*>(b->listB);
            int* n_tiles = (int*)(ctl + 5);
            wf_beam_kernel<<<pg, 128, 0, st>>>(b->view(), j1, (int)N, ctl + 0, beam_tiles_per_fetch(N, pg * 4), tuning().beam_steps, tiles, n_tiles);
            ++g_launches;
            DRT_LAUNCH_Q(wf_q1_tiles_kernel, b->view(), j1, (int)N, tiles, n_tiles, ctl + 6, pol[0]);
        } else
#endif
        DRT_LAUNCH_Q(wf_q1_kernel, b->view(), j1, (int)N, ctl + 0, pol[0]);
        wf_r1_kernel<<<dgrid, 128, 0, st>>>(b->view(), V64, origin, dir, ext_ior, int_ior, out_ori, out_dir, mask3, b->listA, countL, VN);
        ExitJob j2{out_ori, out_dir, b->listA};
        DRT_LAUNCH_Q(wf_q2_kernel, b->view(), j2, countL, ctl + 1, pol[1]);
        wf_r2_kernel<<<dgrid, 128, 0, st>>>(b->view(), V64, ext_ior, int_ior, out_ori, out_dir, mask3, b->listA, countL, b->listB, countM, VN);
        OcclusionJob j3{out_ori, out_dir, mask3, b->listB, (int4*)rec, rec_count};
        DRT_LAUNCH_Q(wf_q3_kernel, b->view(), j3, countM, ctl + 2, pol[2]);
#undef DRT_LAUNCH_Q
        g_launches += 4;
    }
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_trace_bwd(const drt_bvh* b, const double* V64, const double* origin, const double* dir, int64_t N, double ext_ior,
                  double int_ior, const int32_t* rec, const int32_t* rec_count, const double* g_out_ori,
                  const double* g_out_dir, double* grad_V, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_trace_bwd: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_trace_bwd: no mesh has been set (update_mesh first)");
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_trace_bwd: N < 0");
    if (N == 0 || b->nF == 0) return DRT_OK;
    if (!V64 || !origin || !dir || !rec || !rec_count || !g_out_dir || !grad_V) return fail(DRT_ERR_INVALID, "drt_trace_bwd: null buffer");
    DeviceGuard g(b->device);
    int grid = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * DRT_BWD_MINB);
    // many rays per vertex = heavy contention on the float64 atomics: merge equal-triangle runs in the warp first
    const bool merge = tuning().bwd_merge == 1 || (tuning().bwd_merge < 0 && N / std::max(b->nV, 1) > kMergeRaysPerVertex);
    if (merge)
        trace_bwd_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(b->view(), V64, origin, dir, ext_ior, int_ior, (const int4*)rec,
                                                                    rec_count, g_out_ori, g_out_dir, grad_V);
    else
        trace_bwd_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(b->view(), V64, origin, dir, ext_ior, int_ior, (const int4*)rec,
                                                                     rec_count, g_out_ori, g_out_dir, grad_V);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_trace_bwd_smooth(const drt_bvh* b, const double* V64, const double* VN64, const double* origin, const double* dir, int64_t N,
                         double ext_ior, double int_ior, const int32_t* rec, const int32_t* rec_count, const double* g_out_ori,
                         const double* g_out_dir, double* grad_V, double* grad_VN, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_trace_bwd_smooth: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_trace_bwd_smooth: no mesh has been set (update_mesh first)");
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_trace_bwd_smooth: N < 0");
    if (N == 0 || b->nF == 0) return DRT_OK;
    if (!V64 || !VN64 || !origin || !dir || !rec || !rec_count || !g_out_dir || !grad_V || !grad_VN)
        return fail(DRT_ERR_INVALID, "drt_trace_bwd_smooth: null buffer");
    DeviceGuard g(b->device);
    int grid = (int)std::min<int64_t>(blocks_for(N, 128), (int64_t)b->sm_count * DRT_BWD_MINB);
    trace_bwd_kernel<false, true><<<grid, 128, 0, (cudaStream_t)stream>>>(b->view(), V64, origin, dir, ext_ior, int_ior, (const int4*)rec,
                                                                       rec_count, g_out_ori, g_out_dir, grad_V, VN64, grad_VN);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_plane_hit(const double* out_ori, const double* out_dir, const uint8_t* mask3, int64_t N, const double plane[6], double* pts,
                  uint8_t* front, void* stream)
{
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_plane_hit: N < 0");
    if (N == 0) return DRT_OK;
    if (!out_ori || !out_dir || !mask3 || !plane || !pts) return fail(DRT_ERR_INVALID, "drt_plane_hit: null buffer");
    const Plane P{plane[0], plane[1], plane[2], plane[3], plane[4], plane[5]};
    plane_hit_kernel<<<(int)std::min<int64_t>(blocks_for(N, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(out_ori, out_dir, mask3, N, P, pts, front);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_plane_hit_bwd(const double* out_ori, const double* out_dir, const uint8_t* mask3, int64_t N, const double plane[6],
                      const double* g_pts, double* g_out_ori, double* g_out_dir, void* stream)
{
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_plane_hit_bwd: N < 0");
    if (N == 0) return DRT_OK;
    if (!out_ori || !out_dir || !mask3 || !plane || !g_pts || !g_out_ori || !g_out_dir) return fail(DRT_ERR_INVALID, "drt_plane_hit_bwd: null buffer");
    const Plane P{plane[0], plane[1], plane[2], plane[3], plane[4], plane[5]};
    plane_hit_bwd_kernel<<<(int)std::min<int64_t>(blocks_for(N, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(out_ori, out_dir, mask3, N, P, g_pts,
                                                                                                          g_out_ori, g_out_dir);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_ray_loss_grad(const double* out_ori, const double* out_dir, const uint8_t* mask3, const double* screen,
                      const uint8_t* valid, int64_t N, double* g_out_dir, double* loss_sum, void* stream)
{
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_ray_loss_grad: N < 0");
    if (N == 0) return DRT_OK;
    if (!out_ori || !out_dir || !mask3 || !screen || !g_out_dir) return fail(DRT_ERR_INVALID, "drt_ray_loss_grad: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(out_ori, &dev));  // launch on the device that owns the buffers, whatever the caller's current device is
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (int)std::min<int64_t>(blocks_for(N, 256), (int64_t)sms * 16);
    ray_loss_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out_ori, out_dir, mask3, screen, valid, N, g_out_dir, loss_sum); ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_ray_loss_grad_rec(const double* out_ori, const double* out_dir, const double* screen, const uint8_t* valid,
                          const int32_t* rec, const int32_t* rec_count, int64_t N, double* g_out_dir, double* loss_sum, void* stream)
{
    if (N < 0) return fail(DRT_ERR_INVALID, "drt_ray_loss_grad_rec: N < 0");
    if (N == 0) return DRT_OK;
    if (!out_ori || !out_dir || !screen || !rec || !rec_count || !g_out_dir) return fail(DRT_ERR_INVALID, "drt_ray_loss_grad_rec: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(out_ori, &dev));
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (int)std::min<int64_t>(blocks_for(N, 256), (int64_t)sms * 8);
    ray_loss_rec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out_ori, out_dir, screen, valid, (const int4*)rec, rec_count, g_out_dir, loss_sum);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

// signature word of prepared tile beams: pixels per image and the tile shape (wavefront.cuh: TileBeams)
static int beam_sig(const TileMap& tm) { return (int)(((unsigned)tm.img_hw << 3) | (unsigned)tm.tw_log2); }

int64_t drt_tile_beams_floats(int64_t N) { return N < 0 ? 0 : 12 * ((N + 31) / 32 + 1); }

int drt_tile_beams(const double* origin, int64_t rays_per_origin, const double* dir, int64_t N, int32_t image_w, int32_t image_h,
                   float* beams, void* stream)
{
    if (N < 0 || N > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_tile_beams: N must be in [0, 2^31)");
    if (rays_per_origin < 1 || rays_per_origin > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_tile_beams: rays_per_origin must be >= 1");
    if (image_w < 0 || image_h < 0) return fail(DRT_ERR_INVALID, "drt_tile_beams: negative image size");
    if (!beams || (N > 0 && (!origin || !dir))) return fail(DRT_ERR_INVALID, "drt_tile_beams: null buffer");
    if (((uintptr_t)beams & 15u) != 0) return fail(DRT_ERR_INVALID, "drt_tile_beams: beams must be 16-byte aligned");
    int dev;
    CU(device_of(beams, &dev));
    DeviceGuard g(dev);
    const TileMap tm = tile_map(image_w, image_h, N, false);
    if ((int64_t)tm.img_hw >= (1 << 28)) return fail(DRT_ERR_INVALID, "drt_tile_beams: image too large");
#if DRT_FUSE_R
    const LossEntryJob job{RaySrc{origin, dir, (int)rays_per_origin}, nullptr, nullptr, tm, 0, RefractCtx{nullptr, nullptr, 0.0, 0.0}, Park{nullptr, 0}};
#else
    const LossEntryJob job{RaySrc{origin, dir, (int)rays_per_origin}, nullptr, nullptr, tm, 0};
#endif
    const int64_t n_tiles = (N + 31) / 32;
    tile_beams_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>(blocks_for(n_tiles, 128), 148 * 16)), 128, 0, (cudaStream_t)stream>>>(
        job, (int)N, reinterpret_cast<float4*>(beams), n_tiles + 1, beam_sig(tm));
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_ray_loss_step(drt_bvh* b, const double* V64, const double* origin, int64_t rays_per_origin, const double* dir,
                      int64_t N, double ext_ior, double int_ior, int target_mode, const double* screen, const uint8_t* valid,
                      const int32_t* tgt_idx, const double* tgt_xyz, int64_t n_tgt, int32_t image_w, int32_t image_h,
                      double* loss_sum, double* grad_V, int32_t* n_paths, void* ev_after_fwd, void* stream)
{
    return drt_ray_loss_step_beams(b, V64, origin, rays_per_origin, dir, N, ext_ior, int_ior, target_mode, screen, valid, tgt_idx, tgt_xyz,
                                   n_tgt, image_w, image_h, nullptr, loss_sum, grad_V, n_paths, ev_after_fwd, stream);
}

int drt_ray_loss_step_beams(drt_bvh* b, const double* V64, const double* origin, int64_t rays_per_origin, const double* dir,
                            int64_t N, double ext_ior, double int_ior, int target_mode, const double* screen, const uint8_t* valid,
                            const int32_t* tgt_idx, const double* tgt_xyz, int64_t n_tgt, int32_t image_w, int32_t image_h,
                            const float* tile_beams, double* loss_sum, double* grad_V, int32_t* n_paths, void* ev_after_fwd, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_ray_loss_step: no mesh has been set (update_mesh first)");
    if (N < 0 || N > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: N must be in [0, 2^31)");
    if (rays_per_origin < 1 || rays_per_origin > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: rays_per_origin must be >= 1");
    if (target_mode != 0 && target_mode != 1) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: target_mode must be 0 (dense) or 1 (sparse)");
    if (n_tgt < 0 || n_tgt > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: n_tgt must be in [0, 2^31)");
    if (image_w < 0 || image_h < 0) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: negative image size");
    if (((uintptr_t)tile_beams & 15u) != 0) return fail(DRT_ERR_INVALID, "drt_ray_loss_step_beams: tile_beams must be 16-byte aligned");
    DeviceGuard g(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_paths) CU(cudaMemsetAsync(n_paths, 0, sizeof(int32_t), st));
    if (N == 0 || b->nF == 0) {
        if (ev_after_fwd) CU(cudaEventRecord((cudaEvent_t)ev_after_fwd, st));
        return DRT_OK;
    }
    if (!V64 || !origin || !dir || !loss_sum) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: null buffer");
    if (target_mode == 0 && !screen) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: dense targets need screen[N,3]");
    if (target_mode == 1 && n_tgt > 0 && (!tgt_idx || !tgt_xyz)) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: sparse targets need tgt_idx and tgt_xyz");
    int rc;
    if ((rc = ensure(b->listA, b->capLA, (size_t)N))) return rc;
    if ((rc = ensure(b->park, b->capPk, 6 * (size_t)N))) return rc;
    if ((rc = ensure(b->listM, b->capLM, (size_t)N))) return rc;
    if ((rc = ensure(b->listS, b->capLS, (size_t)N))) return rc;  // also holds the int2 list of surviving tiles (N/16 ints) before Q3
    const int* pol = tuning().pol;
    const int minb = tuning().minb;
    const RaySrc rays{origin, dir, (int)rays_per_origin};
    const int n_buckets = (int)(N >> kTgtShift) + 2;
    if (target_mode == 1) {
        if ((rc = ensure(b->tbucket, b->capTb, (size_t)n_buckets))) return rc;
        tgt_bucket_kernel<<<blocks_for(n_buckets, 256), 256, 0, st>>>(tgt_idx, (int)n_tgt, n_buckets, b->tbucket);
        ++g_launches;
    }
    const TargetSrc tgt{screen, valid, tgt_idx, tgt_xyz, b->tbucket, (int)n_tgt, target_mode};
    const bool merge = tuning().bwd_merge == 1 || (tuning().bwd_merge < 0 && N / std::max(b->nV, 1) > kMergeRaysPerVertex);

    // Two LANES: the batch is split in two halves (at an image boundary when the image size is known) that run the same seven
    // stages on two internal streams with their own halves of the scratch.  Every stage is a persistent kernel that ends in a
    // tail of a few long rays while most SMs idle; with two lanes the other half's kernels fill those tails (the block scheduler
    // starts them as the blocks of the draining kernel exit).  Results do not depend on it (the halves only share the atomically
    // accumulated loss and gradient).  DRT_LANES=1 switches it off.
    const int64_t img = (image_w > 0 && image_h > 0) ? (int64_t)image_w * image_h : 0;
    // measured on B200 (C4): 9 views (6.2 M rays) 1.32 -> 1.27 ms with two lanes, 72 views (49.8 M rays) 5.77 -> 5.85 ms: the
    // tails only matter when a stage lasts a few hundred microseconds, so large batches stay on one lane
    int n_lanes = 1;
    int64_t cut[kMaxLanes + 1];  // lane k = rays [cut[k], cut[k+1])
    cut[0] = 0;
    for (int k = 1; k <= kMaxLanes; ++k) cut[k] = N;
    // round 2, final kernels: above 16 M rays FOUR lanes win as well -- not for the tails (2 % of a stage there) but because the
    // latency-bound loss/backward kernel of one lane (occupancy 25 %) overlaps the query kernels of the others: C4 72 views
    // 5.62 -> 5.52 ms, C3 7.70 -> 7.47, C5 (32 views) 14.77 -> 14.55; 3 / 4 / 6 / 8 lanes at C4: 5.51 / 5.52 / 5.54 / 5.58
    const int want_lanes = tuning().lanes_forced ? tuning().lanes : (N <= tuning().lanes_max_rays ? tuning().lanes : tuning().lanes_big);
    if (want_lanes >= 2 && b->lane_stream[0] && N >= (1 << 17) && N > tuning().direct_max_rays) {
        // split at image boundaries when there are enough images, else at 32-ray batches (= pixel tiles: the item -> ray map
        // below is the one of the WHOLE batch, so a lane may start in the middle of an image)
        const int64_t unit = (img > 0 && N % img == 0 && N / img >= want_lanes) ? img : 32;
        if (unit > 0) {
            const int64_t units = N / unit;
            n_lanes = (int)std::min<int64_t>(want_lanes, std::max<int64_t>(1, units));
            for (int k = 1; k < n_lanes; ++k) cut[k] = (units * k / n_lanes) * unit;
            for (int k = n_lanes; k <= kMaxLanes; ++k) cut[k] = N;
            for (int k = 0; k < n_lanes; ++k)
                if (cut[k + 1] <= cut[k]) { n_lanes = 1; cut[1] = N; break; }  // degenerate split: one lane
        }
    }
    cudaStream_t ls[kMaxLanes];
    for (int k = 0; k < kMaxLanes; ++k) ls[k] = st;
    if (n_lanes > 1) {
        CU(cudaEventRecord(b->lane_event[0], st));  // fork: the lanes start after everything enqueued on the caller's stream
        for (int k = 0; k < n_lanes; ++k) {
            ls[k] = b->lane_stream[k];
            CU(cudaStreamWaitEvent(ls[k], b->lane_event[0], 0));
        }
    }
    struct Lane {
        int64_t base, n;
        unsigned long long* ctl;
        int *countL, *countM, *countS;
        int4* L;
        int2* M;
        int4* S;
        Park park;
        int pg, dgrid;
    } lane[kMaxLanes];
    const int64_t capL = (int64_t)b->capLA, capM = (int64_t)b->capLM, capS = (int64_t)b->capLS, capP = (int64_t)(b->capPk / 6);
    for (int k = 0; k < n_lanes; ++k) {
        Lane& l = lane[k];
        l.base = cut[k];
        l.n = cut[k + 1] - cut[k];
        // control block: work counters of Q1,Q2,Q3 + {countL, countM} + {countS, -} + {tiles kept, -} + tile-list work counter
        l.ctl = b->work + (size_t)(b->work_slot++ % (kWorkSlots / 8)) * 8;
        CU(cudaMemsetAsync(l.ctl, 0, 8 * sizeof(unsigned long long), ls[k]));
        l.countL = (int*)(l.ctl + 3);
        l.countM = l.countL + 1;
        l.countS = (int*)(l.ctl + 4);
        // each lane owns the part of the scratch that corresponds to its share of the rays (capacities are >= N entries)
        l.L = b->listA + (int64_t)((__int128)capL * cut[k] / N);
        l.M = b->listM + (int64_t)((__int128)capM * cut[k] / N);
        l.S = b->listS + (int64_t)((__int128)capS * cut[k] / N);
        l.park = Park{b->park + (int64_t)((__int128)capP * cut[k] / N), capP};
        l.pg = (int)std::min<int64_t>(blocks_for(l.n, 128), (int64_t)b->sm_count * minb);
        l.dgrid = (int)std::min<int64_t>(blocks_for(l.n, 128), (int64_t)b->sm_count * tuning().r_grid);
    }
    for (int k = 0; k < kMaxLanes; ++k) b->last_ctl[k] = k < n_lanes ? lane[k].ctl : nullptr;
    b->last_tiles = 0;
#define DRT_LAUNCH_Q(KERNEL, PG, ST, ...)                                                   \
    do {                                                                                    \
        if (minb == 10) KERNEL<10><<<PG, 128, 0, ST>>>(__VA_ARGS__);                        \
        else if (minb == 8) KERNEL<8><<<PG, 128, 0, ST>>>(__VA_ARGS__);                     \
        else if (minb == 7) KERNEL<7><<<PG, 128, 0, ST>>>(__VA_ARGS__);                     \
        else if (minb == 6) KERNEL<6><<<PG, 128, 0, ST>>>(__VA_ARGS__);                     \
        else KERNEL<4><<<PG, 128, 0, ST>>>(__VA_ARGS__);                                    \
    } while (0)
    const bool beam = DRT_QNODE && tuning().beam && (pol[0] & 0xff) == 32;
    // small batches: the whole forward path of a ray in one thread (ls_direct_kernel), then the usual loss/backward kernel
    const bool direct = n_lanes == 1 && N <= tuning().direct_max_rays;
    b->last_lanes = direct ? 0 : n_lanes;
    if (direct) {
        Lane& l = lane[0];
        const int dg = (int)std::min<int64_t>(blocks_for(l.n, 128), (int64_t)b->sm_count * 16);
        ls_direct_kernel<<<dg, 128, 0, st>>>(b->view(), V64, rays, tile_map(image_w, image_h, N, false), (int)N, ext_ior, int_ior, tgt, l.S,
                                             l.countL, l.countM, l.countS);
        g_launches -= 4;  // direct + loss/backward = 2 launches instead of the 6 counted below
    }
    // stage by stage, lane by lane: the launches of the two lanes alternate so that neither stream waits for the host
    for (int k = 0; k < n_lanes && !direct; ++k) {  // Q1: entry query
        Lane& l = lane[k];
        // whole images of image_w x image_h pixels that a 32-pixel tile shape divides: a warp's batch becomes a pixel tile
#if DRT_FUSE_R
        LossEntryJob j1{rays, l.L, l.countL, tile_map(image_w, image_h, N, false), (int)l.base, RefractCtx{V64, b->F, ext_ior, int_ior}, l.park};
#else
        LossEntryJob j1{rays, l.L, l.countL, tile_map(image_w, image_h, N, false), (int)l.base};
#endif
        if (l.base % 32 != 0) return fail(DRT_ERR_INVALID, "drt_ray_loss_step: internal lane split is not tile aligned");
#if DRT_QNODE
        if (beam) {
            // beam pass over all tiles -> list of the surviving tiles (in the lane's part of listS, free until Q3) -> per-ray
            // entry query over the list
            int2* tiles = reinterpret_cast<int2*>(l.S);
            int* n_tiles = (int*)(l.ctl + 5);
            b->last_tiles += (l.n + 31) / 32;
            // one beam per LANE needs 32 tiles per fetch: small batches get a smaller grid (>= 2 fetches per warp), not fewer tiles
            const int tpb = tuning().beam_tpb ? tuning().beam_tpb : 32;
            const int bg = (int)std::max<int64_t>(1, std::min<int64_t>(l.pg, ((l.n + 31) / 32 + (int64_t)tpb * 8 - 1) / ((int64_t)tpb * 8)));
            // intervals prepared by drt_tile_beams for exactly this batch (the kernel checks the signature: N, image, tile shape)
            const TileBeams prepared{reinterpret_cast<const float4*>(tile_beams), (N + 31) / 32 + 1, (int)N, j1.tiles.img_w, beam_sig(j1.tiles)};
            ls_beam_kernel<<<bg, 128, 0, ls[k]>>>(b->view(), j1, (int)l.n, l.ctl + 0, tpb, tuning().beam_steps, tiles, n_tiles, prepared);
            ++g_launches;
            DRT_LAUNCH_Q(ls_q1_tiles_kernel, l.pg, ls[k], b->view(), j1, (int)l.n, tiles, n_tiles, l.ctl + 6, pol[0]);
        } else
#endif
        DRT_LAUNCH_Q(ls_q1_kernel, l.pg, ls[k], b->view(), j1, (int)l.n, l.ctl + 0, pol[0]);
    }
    for (int k = 0; k < n_lanes && !direct; ++k) {  // R1 + Q2: refraction at the entry hit, exit query
        Lane& l = lane[k];
#if DRT_FUSE_R
        LossExitJob j2{l.park, l.L, RefractCtx{V64, b->F, ext_ior, int_ior}, tgt, l.M, l.countM};
#else
        ls_r1_kernel<<<l.dgrid, 128, 0, ls[k]>>>(b->view(), V64, rays, ext_ior, int_ior, l.L, l.countL, l.park);
        LossExitJob j2{l.park, l.L};
#endif
        DRT_LAUNCH_Q(ls_q2_kernel, l.pg, ls[k], b->view(), j2, l.countL, l.ctl + 1, pol[1]);
    }
    for (int k = 0; k < n_lanes && !direct; ++k) {  // R2 + Q3: refraction at the exit hit (+ target lookup), occlusion query
        Lane& l = lane[k];
#if !DRT_FUSE_R
        ls_r2_kernel<<<l.dgrid, 128, 0, ls[k]>>>(b->view(), V64, ext_ior, int_ior, l.L, l.countL, l.park, tgt, l.M, l.countM);
#endif
        LossOcclusionJob j3{l.park, l.M, l.L, l.S, l.countS};
        DRT_LAUNCH_Q(ls_q3_kernel, l.pg, ls[k], b->view(), j3, l.countM, l.ctl + 2, pol[2]);
    }
#undef DRT_LAUNCH_Q
    if (ev_after_fwd) {
        for (int k = 0; k < n_lanes && n_lanes > 1; ++k) {  // the marker fires when ALL lanes have finished their forward part
            CU(cudaEventRecord(b->lane_event[1 + k], ls[k]));
            CU(cudaStreamWaitEvent(st, b->lane_event[1 + k], 0));
        }
        CU(cudaEventRecord((cudaEvent_t)ev_after_fwd, st));
    }
    for (int k = 0; k < n_lanes; ++k) {  // loss + backward over the valid paths
        Lane& l = lane[k];
        // grid = what is co-resident (the kernel is register-bound at 3-4 blocks per SM): a larger grid-stride grid only adds a ragged last wave
        const int bgrid = (int)std::min<int64_t>(blocks_for(l.n, 128), (int64_t)b->sm_count * std::max(1, grad_V ? (merge ? b->bwd_blocks[2] : b->bwd_blocks[1]) : b->bwd_blocks[0]));
        if (!grad_V)
            ls_loss_bwd_kernel<false, false><<<bgrid, 128, 0, ls[k]>>>(b->view(), V64, rays, ext_ior, int_ior, l.S, l.countS, tgt, loss_sum, nullptr);
        else if (merge)
            ls_loss_bwd_kernel<true, true><<<bgrid, 128, 0, ls[k]>>>(b->view(), V64, rays, ext_ior, int_ior, l.S, l.countS, tgt, loss_sum, grad_V);
        else
            ls_loss_bwd_kernel<true, false><<<bgrid, 128, 0, ls[k]>>>(b->view(), V64, rays, ext_ior, int_ior, l.S, l.countS, tgt, loss_sum, grad_V);
    }
    g_launches += 6 * n_lanes;
    CU(cudaGetLastError());
    for (int k = 0; k < n_lanes && n_lanes > 1; ++k) {  // join: the caller's stream continues after all lanes
        CU(cudaEventRecord(b->lane_event[1 + kMaxLanes + k], ls[k]));
        CU(cudaStreamWaitEvent(st, b->lane_event[1 + kMaxLanes + k], 0));
    }
    if (n_paths) {
        LaneCounts lc{};
        for (int k = 0; k < n_lanes; ++k) lc.p[k] = lane[k].countS;
        sum_counts_kernel<<<1, 32, 0, st>>>(lc, n_lanes, n_paths);
        ++g_launches;
    }
    return DRT_OK;
}

int drt_generate_rays(int32_t resy, int32_t resx, const double* K_inverse, const double* R_inverse, double* origin3, double* dir,
                      void* stream)
{
    if (resy < 0 || resx < 0 || (int64_t)resy * resx > 2147483647LL) return fail(DRT_ERR_INVALID, "drt_generate_rays: bad resolution %d x %d", resy, resx);
    if (!K_inverse || !R_inverse || !origin3 || (!dir && (int64_t)resy * resx > 0)) return fail(DRT_ERR_INVALID, "drt_generate_rays: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(origin3, &dev));
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n = (int64_t)resy * resx;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_for(n, 256), (int64_t)sms * 16));
    generate_rays_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(resy, resx, K_inverse, R_inverse, origin3, dir); ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

// ---- silhouette-edge sampling -----------------------------------------------------------------------
int drt_silhouette_classify(const double* V64, const int32_t* e2f, int64_t nE, const double* origin3, uint8_t* flags, void* stream)
{
    if (nE < 0) return fail(DRT_ERR_INVALID, "drt_silhouette_classify: nE < 0");
    if (nE == 0) return DRT_OK;
    if (!V64 || !e2f || !origin3 || !flags) return fail(DRT_ERR_INVALID, "drt_silhouette_classify: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(flags, &dev));
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<int64_t>(blocks_for(nE, 256), (int64_t)sms * 8);
    silhouette_classify_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(V64, e2f, nE, origin3, flags); ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_silhouette_sample(const drt_bvh* b, const double* V64, const int64_t* edges, int64_t k, const double* R, const double* K,
                          const double* R_inverse, const double* K_inverse, const double* origin3, int32_t resx, int32_t resy,
                          int64_t* index_xy, double* f, uint8_t* keep, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_silhouette_sample: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_silhouette_sample: no mesh has been set (update_mesh first)");
    if (k < 0) return fail(DRT_ERR_INVALID, "drt_silhouette_sample: k < 0");
    if (k == 0) return DRT_OK;
    if (!V64 || !edges || !R || !K || !R_inverse || !K_inverse || !origin3 || !index_xy || !f || !keep)
        return fail(DRT_ERR_INVALID, "drt_silhouette_sample: null buffer");
    DeviceGuard g(b->device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", b->device);
    const int grid = (int)std::min<int64_t>(blocks_for(2 * k, 128), (int64_t)b->sm_count * 8);
    silhouette_sample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(b->view(), V64, edges, k, Camera{R, K, R_inverse, K_inverse}, origin3, resx,
                                                                     resy, index_xy, f, keep);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_silhouette_backward(const double* V64, const int64_t* edges, const double* R, const double* K, int detach_depth, const double* f,
                            const int64_t* kept_idx, const float* g_output, int64_t m, double* grad_V, void* stream)
{
    if (m < 0) return fail(DRT_ERR_INVALID, "drt_silhouette_backward: m < 0");
    if (m == 0) return DRT_OK;
    if (!V64 || !edges || !R || !K || !f || !kept_idx || !g_output || !grad_V) return fail(DRT_ERR_INVALID, "drt_silhouette_backward: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(grad_V, &dev));
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<int64_t>(blocks_for(m, 128), (int64_t)sms * 8);
    silhouette_backward_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(V64, edges, Camera{R, K, nullptr, nullptr}, detach_depth, f, kept_idx,
                                                                       g_output, m, grad_V);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_silhouette_loss(const drt_bvh* b, const double* V64, const int64_t* edges, const int32_t* e2f, int64_t nE, int32_t n_views,
                        const double* const* R, const double* const* K, const double* const* R_inverse, const double* const* K_inverse,
                        const double* const* origin3, const double* const* mask, int32_t resx, int32_t resy, int detach_depth,
                        double* loss_sum, double* grad_V, int32_t* n_samples, void* stream)
{
    if (!b) return fail(DRT_ERR_INVALID, "drt_silhouette_loss: null handle");
    if (!b->built) return fail(DRT_ERR_STATE, "drt_silhouette_loss: no mesh has been set (update_mesh first)");
    if (nE < 0 || resx <= 0 || resy <= 0 || n_views < 0) return fail(DRT_ERR_INVALID, "drt_silhouette_loss: bad size");
    if (nE == 0 || n_views == 0) return DRT_OK;
    if (!V64 || !edges || !e2f || !R || !K || !R_inverse || !K_inverse || !origin3 || !mask || !loss_sum)
        return fail(DRT_ERR_INVALID, "drt_silhouette_loss: null buffer");
    DeviceGuard g(b->device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", b->device);
    const int gx = (int)std::min<int64_t>(blocks_for(nE, 128), (int64_t)b->sm_count * 8);
    for (int v0 = 0; v0 < n_views; v0 += kMaxSilhouetteViews) {  // up to 8 views per launch
        const int nv = std::min<int>(kMaxSilhouetteViews, n_views - v0);
        SilhouetteViews sv{};
        for (int k = 0; k < nv; ++k) {
            if (!R[v0 + k] || !K[v0 + k] || !R_inverse[v0 + k] || !K_inverse[v0 + k] || !origin3[v0 + k] || !mask[v0 + k])
                return fail(DRT_ERR_INVALID, "drt_silhouette_loss: null per-view pointer (view %d)", v0 + k);
            sv.cam[k] = Camera{R[v0 + k], K[v0 + k], R_inverse[v0 + k], K_inverse[v0 + k]};
            sv.origin[k] = origin3[v0 + k];
            sv.mask[k] = mask[v0 + k];
        }
        silhouette_loss_kernel<<<dim3(gx, nv), 128, 0, (cudaStream_t)stream>>>(b->view(), V64, edges, e2f, nE, sv, resx, resy, detach_depth, loss_sum,
                                                                               grad_V, n_samples);
        ++g_launches;
    }
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_dihedral_loss(const double* V64, const int32_t* e2f, int64_t nE, double* loss_sum, double* grad_V, void* stream)
{
    if (nE < 0) return fail(DRT_ERR_INVALID, "drt_dihedral_loss: nE < 0");
    if (nE == 0) return DRT_OK;
    if (!V64 || !e2f || !loss_sum) return fail(DRT_ERR_INVALID, "drt_dihedral_loss: null buffer");
    int dev = 0, sms = 148;
    CU(device_of(loss_sum, &dev));
    DeviceGuard g(dev);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<int64_t>(blocks_for(nE, 256), (int64_t)sms * 8);
    dihedral_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(V64, e2f, nE, loss_sum, grad_V);
    ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

// ---- peer-memory all-reduce of grad_V ------------------------------------------------------
struct drt_comm {
    int device = 0, rank = 0, world = 1;
    int64_t stride = 0;             // doubles per staging slot
    unsigned char* region = nullptr; // local IPC-exported region: [2 parities][world slots][stride doubles][2*kMaxPeers flags][arrive][error]
    void* peer_region[drt::kMaxPeers] = {};
    bool connected = false;
    unsigned epoch = 0;
    int sm_count = 148;
    size_t flags_off = 0;
    unsigned long long timeout_ns = 120ull * 1000000000ull;  // DRT_PEER_TIMEOUT_S
};

namespace {
size_t comm_bytes(int64_t stride, int world) { return (size_t)stride * 2 * world * sizeof(double) + (2 * drt::kMaxPeers + 2) * sizeof(unsigned); }
}

int drt_comm_create(int device, int rank, int world, int64_t max_doubles, drt_comm** out)
{
    if (!out) return fail(DRT_ERR_INVALID, "drt_comm_create: out is null");
    *out = nullptr;
    if (world < 1 || world > drt::kMaxPeers || rank < 0 || rank >= world) return fail(DRT_ERR_INVALID, "drt_comm_create: rank %d / world %d out of range (max %d ranks)", rank, world, drt::kMaxPeers);
    if (max_doubles < 1) return fail(DRT_ERR_INVALID, "drt_comm_create: max_doubles < 1");
    DeviceGuard g(device);
    if (!g.ok) return fail(DRT_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    std::unique_ptr<drt_comm, int (*)(drt_comm*)> guard(new drt_comm(), drt_comm_destroy);  // freed on every early return
    drt_comm* c = guard.get();
    c->device = device; c->rank = rank; c->world = world;
    if (const char* to = getenv("DRT_PEER_TIMEOUT_S"))
        if (atof(to) > 0) c->timeout_ns = (unsigned long long)(atof(to) * 1e9);
    c->stride = (max_doubles + 31) / 32 * 32;
    c->flags_off = (size_t)c->stride * 2 * world * sizeof(double);
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaMalloc(&c->region, comm_bytes(c->stride, world)));
    CU(cudaMemset(c->region, 0, comm_bytes(c->stride, world)));
    CU(cudaDeviceSynchronize());
    c->peer_region[rank] = c->region;
    *out = guard.release();
    return DRT_OK;
}

int drt_comm_handle(const drt_comm* c, unsigned char handle[64])
{
    if (!c || !handle) return fail(DRT_ERR_INVALID, "drt_comm_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    DeviceGuard g(c->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->region));
    memcpy(handle, &h, 64);
    return DRT_OK;
}

int drt_comm_connect(drt_comm* c, const unsigned char* handles)
{
    if (!c || !handles) return fail(DRT_ERR_INVALID, "drt_comm_connect: null argument");
    DeviceGuard g(c->device);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        CU(cudaIpcOpenMemHandle(&c->peer_region[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->connected = true;
    return DRT_OK;
}

int drt_comm_allreduce_sum_f64(drt_comm* c, double* data, int64_t n, void* stream)
{
    if (!c) return fail(DRT_ERR_INVALID, "drt_comm_allreduce_sum_f64: null handle");
    if (n < 0 || n > c->stride) return fail(DRT_ERR_INVALID, "drt_comm_allreduce_sum_f64: n = %lld exceeds the %lld doubles the communicator was created for", (long long)n, (long long)c->stride);
    if (n == 0 || c->world == 1) return DRT_OK;
    if (!data) return fail(DRT_ERR_INVALID, "drt_comm_allreduce_sum_f64: null buffer");
    if (!c->connected) return fail(DRT_ERR_STATE, "drt_comm_allreduce_sum_f64: drt_comm_connect has not been called");
    DeviceGuard g(c->device);
    PeerView pv;
    for (int r = 0; r < c->world; ++r) {
        pv.stage[r] = reinterpret_cast<double*>(c->peer_region[r]);
        pv.flags[r] = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(c->peer_region[r]) + c->flags_off);
    }
    unsigned* local = reinterpret_cast<unsigned*>(c->region + c->flags_off);
    pv.arrive = local + 2 * kMaxPeers;
    pv.error = local + 2 * kMaxPeers + 1;
    pv.stride = c->stride;
    pv.timeout_ns = c->timeout_ns;
    pv.rank = c->rank;
    pv.world = c->world;
    // co-resident by construction: at most one block per SM (the blocks spin on the peers' flags)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_for(n, 512), std::min(c->sm_count, 64)));
    peer_allreduce_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(pv, data, n, ++c->epoch); ++g_launches;
    CU(cudaGetLastError());
    return DRT_OK;
}

int drt_comm_status(const drt_comm* c, void* stream, int* timed_out)
{
    if (!c || !timed_out) return fail(DRT_ERR_INVALID, "drt_comm_status: null argument");
    DeviceGuard g(c->device);
    unsigned e = 0;
    CU(cudaMemcpyAsync(&e, c->region + c->flags_off + (2 * kMaxPeers + 1) * sizeof(unsigned), sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    *timed_out = (int)e;
    return DRT_OK;
}

int drt_comm_destroy(drt_comm* c)
{
    if (!c) return DRT_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer_region[r]) cudaIpcCloseMemHandle(c->peer_region[r]);
    if (c->region) cudaFree(c->region);
    delete c;
    return DRT_OK;
}

}  // extern "C"
