/*
 * drt_b200.h -- C ABI of the B200-native refraction tracer (libdrt_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of lvjiahui/DRT: the ray-query plugin
 * `optix_mesh` (reference optix_extend.cpp:6-83, OptiX Prime closest-hit) and the per-ray
 * two-bounce refraction chain + its autograd backward (reference DiffRender.py:420-432, 492-546;
 * optim.py:210).  Plain pointers and sizes only; no torch, no C++ types.  All pointers are DEVICE
 * pointers on the device the handle was created for, unless stated otherwise; `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Every call is asynchronous with
 * respect to the host and ordered on `stream`.
 *
 * Return value: 0 on success, non-zero error code otherwise; drt_last_error() gives the text of
 * the last error raised on the calling thread.  No exceptions cross this boundary (the reference
 * plugin aborts through C assert / C++ exceptions: optix_extend.cpp:17-18,25,30-31).
 */
#ifndef DRT_B200_H
#define DRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DRT_API __attribute__((visibility("default")))
#else
#define DRT_API
#endif

#define DRT_OK 0
#define DRT_ERR_INVALID 1 /* bad argument (null pointer, negative size, index out of range flag) */
#define DRT_ERR_CUDA 2    /* a CUDA runtime call failed; see drt_last_error() */
#define DRT_ERR_STATE 3   /* query before any build (reference: assert(builded), optix_extend.cpp:30) */

/*
 * A handle is SINGLE-STREAM and SINGLE-THREAD, like the reference object it replaces (one optix_mesh, called from the
 * Python main thread, query->execute(0) blocking: optix_extend.cpp:45): drt_trace_fwd, drt_ray_loss_step and
 * drt_silhouette_* use wavefront lists, parked rays and work counters that live IN the handle, so two calls on the same
 * handle must be ordered on one stream (or by events).  Use one handle per stream for concurrent queries.
 */
typedef struct drt_bvh drt_bvh;

/* Replaces optix_mesh::optix_mesh(unsigned cuda_device) -- optix_extend.cpp:8-12. */
DRT_API int drt_bvh_create(int device, drt_bvh** out);
DRT_API int drt_bvh_destroy(drt_bvh* bvh);

/*
 * Replaces optix_mesh::update_mesh(F, V) -- optix_extend.cpp:14-21,61-67: set triangles and
 * (re)build the acceleration structure from scratch (LBVH: Morton codes + radix sort + Karras
 * hierarchy + bottom-up fit).  F int32[nF,3], V32 float32[nV,3], contiguous.  The handle keeps its
 * OWN device copies, so the caller's buffers may be freed after the call is enqueued (the
 * reference instead retains the tensors, optix_extend.cpp:15-16,70-71).
 * The build is enqueued on `stream`; later queries on the same stream see it (the reference's
 * update(RTP_MODEL_HINT_ASYNC) + finish() pair, optix_extend.cpp:32,66).
 */
DRT_API int drt_bvh_build(drt_bvh* bvh, const int32_t* F, int32_t nF, const float* V32, int32_t nV, void* stream);

/* Same, taking the float64 vertices of Scene.vertices and doing the float32 cast of
 * DiffRender.py:311,379 on the device (round-to-nearest-even, identical to tensor.to(float32)). */
DRT_API int drt_bvh_build_f64(drt_bvh* bvh, const int32_t* F, int32_t nF, const double* V64, int32_t nV, void* stream);

/*
 * Replaces optix_mesh::update_vert(V) -- optix_extend.cpp:23-27: new vertex positions, previous
 * faces.  `refit` = 0 rebuilds from scratch like the reference does every iteration
 * (DiffRender.py:379-380); `refit` = 1 keeps the tree topology and only refits boxes (legal
 * because topology changes only at update_mesh, optim.py:195).  Exactly one of V32 / V64 non-NULL.
 */
DRT_API int drt_bvh_update_vert(drt_bvh* bvh, const float* V32, const double* V64, int32_t nV, int refit, void* stream);

/*
 * The reference keeps the render resolution in module globals (DiffRender.py:16-17 `resy`, `resx`, assigned
 * by optim.py:179-180); this is their counterpart on the handle.  It is a HINT: when the N rays of a later
 * drt_trace_fwd call are whole images of image_w x image_h pixels in scanline order (N % (w*h) == 0,
 * w % 4 == 0 and h % 8 == 0, or w % 8 == 0 and h % 4 == 0) the entry query walks 32-pixel tiles per warp
 * instead of 32 x 1 strips.
 * Results do not depend on it.  (0, 0) clears the hint.  Host call, no device work.
 */
DRT_API int drt_bvh_set_image_size(drt_bvh* bvh, int32_t image_w, int32_t image_h);

/* Number of face indices that were outside [0,nV) at the last drt_bvh_build (they are clamped so
 * that no kernel faults; the reference would read out of bounds).  Synchronises `stream`. */
DRT_API int drt_bvh_bad_indices(const drt_bvh* bvh, void* stream, int* out);

/* info[0]=nF, [1]=nV, [2]=number of BVH nodes, [3]=built flag, [4]=node bytes, [5]=triangle bytes,
 * [6]=builds so far, [7]=refits so far.  Host call, no device sync. */
DRT_API int drt_bvh_info(const drt_bvh* bvh, int64_t info[8]);

/* Stage counters of the latest drt_ray_loss_step on this handle (synchronises `stream`; diagnostics for bench.py, no
 * reference counterpart -- the reference's equivalents are the lengths of the Ray sets after each Ray.select,
 * DiffRender.py:538-544): out = {entry hits, rays alive after both refractions, valid paths, 32-ray tiles seen by the
 * beam pass, tiles it kept, lanes (internal streams) the call ran on -- 0 for the one-thread-per-path route}. */
DRT_API int drt_bvh_last_counts(const drt_bvh* bvh, void* stream, int64_t out[6]);

/*
 * Replaces optix_mesh::intersect(Ray) -- optix_extend.cpp:29-57 (RTP_QUERY_TYPE_CLOSEST,
 * RTP_BUFFER_FORMAT_RAY_ORIGIN_DIRECTION in, RTP_BUFFER_FORMAT_HIT_T_TRIID out).
 * ray6 float32[N,6] = (origin, direction), direction need NOT be unit length (silhouette rays,
 * DiffRender.py:213-224); tmin = 0, tmax = inf, no culling.
 * T[i*strideT] = hit distance in units of |direction| (>0), or -1 on a miss; ID[i*strideID] =
 * triangle index or -1.  Strides are in ELEMENTS; strideT = strideID = 2 with ID = (int32*)T + 1
 * reproduces the reference's interleaved {float t; int id} hit buffer (optix_extend.cpp:41-56).
 * Exact closest hit: ties in t resolve to the lowest triangle index.
 */
DRT_API int drt_closest_hit(const drt_bvh* bvh, const float* ray6, int64_t N, float* T, int32_t* ID, int64_t strideT,
                    int64_t strideID, void* stream);

/*
 * Replaces Scene.render_transparent -- DiffRender.py:420-432 (trace2 :537-546, Dintersect
 * :492-501, JIT_Dintersect :64-121, refract_ray :503-535, Refract :35-49, FrDielectric :51-61) in
 * ONE launch: entry hit -> refract -> exit hit -> refract -> occlusion query, per ray.
 *   V64            float64[nV,3] current vertices (differentiable re-intersection uses these;
 *                  the queries use the float32 copy the BVH was built from, as the reference does)
 *   origin, dir    float64[N,3] (cast to float32 for the queries: DiffRender.py:387-388)
 *   ext_ior/int_ior  DiffRender.py:21 (1.00029) / optim.py:178
 *   out_ori,out_dir  float64[N,3]; zeros where the ray is not a valid two-bounce path
 *   mask3          uint8[N,3] (torch.bool layout), all three columns equal (DiffRender.py:423,431)
 *   rec, rec_count hit records for drt_trace_bwd, COMPACT: one int32[4] = (ray index, triangle of
 *                  hit 1, triangle of hit 2, 0) per VALID path, in completion order; rec must hold
 *                  4*N int32 (16-byte aligned), rec_count is one device int32 that the call resets
 *                  and the kernel increments.  Both may be NULL (inference only).
 *   hit1           optional uint8[N]: 1 where the primary ray hits anything = Scene.render_mask
 *                  (DiffRender.py:434-438) as a by-product.  May be NULL.
 */
DRT_API int drt_trace_fwd(drt_bvh* bvh, const double* V64, const double* origin, const double* dir, int64_t N,
                  double ext_ior, double int_ior, double* out_ori, double* out_dir, uint8_t* mask3,
                  int32_t* rec, int32_t* rec_count, uint8_t* hit1, void* stream);

/*
 * Replaces loss.backward() through the autograd graph of the chain above (optim.py:210): replays
 * the compact hit records written by drt_trace_fwd (one thread per VALID path), evaluates the analytic Jacobian of (out_ori, out_dir) w.r.t. the six
 * hit-triangle vertices in float64 and scatter-adds into grad_V.
 *   g_out_ori      float64[N,3] upstream gradient of out_ori, or NULL (= zeros; optim.py:100)
 *   g_out_dir      float64[N,3] upstream gradient of out_dir
 *   grad_V         float64[nV,3], ACCUMULATED into (caller zeroes) -- index_put_(accumulate=True)
 */
DRT_API int drt_trace_bwd(const drt_bvh* bvh, const double* V64, const double* origin, const double* dir, int64_t N,
                  double ext_ior, double int_ior, const int32_t* rec, const int32_t* rec_count,
                  const double* g_out_ori, const double* g_out_dir, double* grad_V, void* stream);

/*
 * Fused consumer (reference optim.py:96-106, Loss_calculator.ray_loss): upstream gradient of
 *   loss = sum_{valid & mask} || out_dir - normalize(screen - out_ori.detach()) ||^2
 * written as g_out_dir = 2*(out_dir - target) where valid&mask else 0, and the loss value
 * accumulated into loss_sum[0] (float64, caller zeroes).  valid uint8[N] may be NULL (= all true).
 */
DRT_API int drt_ray_loss_grad(const double* out_ori, const double* out_dir, const uint8_t* mask3, const double* screen,
                      const uint8_t* valid, int64_t N, double* g_out_dir, double* loss_sum, void* stream);

/*
 * The same consumer driven by the compact hit records of drt_trace_fwd (one thread per VALID path instead of
 * a pass over all N rays).  Only the g_out_dir rows of recorded rays are written -- exactly the rows
 * drt_trace_bwd reads; N is the ray count the records were produced for (grid sizing only).
 */
DRT_API int drt_ray_loss_grad_rec(const double* out_ori, const double* out_dir, const double* screen, const uint8_t* valid,
                          const int32_t* rec, const int32_t* rec_count, int64_t N, double* g_out_dir, double* loss_sum,
                          void* stream);

/*
 * The whole ray iteration on one batch of rays in one call, with NO dense per-ray output
 * (replaces Loss_calculator.ray_loss + loss.backward(), optim.py:91-108,210, i.e.
 * Scene.render_transparent DiffRender.py:420-432 -> target = normalize(screen - out_ori.detach())
 * -> sum over valid&mask of |out_dir - target|^2 -> d loss / d vertices):
 * six launches -- Q1, refract, Q2, refract, Q3 (hit ids and exit rays bit-identical to
 * drt_trace_fwd) and one kernel over the valid paths that adds the loss terms and scatter-adds
 * the analytic vertex gradient.  Rays that miss write nothing; d loss/d out_dir is never stored.
 *   origin          float64[ceil(N/rays_per_origin),3]: ray i starts at row i / rays_per_origin.
 *                   rays_per_origin = 1 is the reference layout (one row per ray); a pinhole view
 *                   has ONE origin (captured_data.py:38 expands it), so pass its row once with
 *                   rays_per_origin = rays per view (several views per call: one row per view).
 *   dir             float64[N,3]
 *   target_mode 0   dense targets as the reference holds them: screen float64[N,3] and valid
 *                   uint8[N] (NULL = all true) -- captured_data.py:101-104
 *   target_mode 1   sparse targets: tgt_idx int32[n_tgt] = ray indices with a measured screen
 *                   point, strictly ascending, tgt_xyz float64[n_tgt,3]; every other ray is
 *                   `valid = False` (captured_data.py:104: valid = screen_pixel[:,0] != 0)
 *   image_w/h       optional hint (0, 0 = none): the N rays are whole images of image_w x image_h pixels in
 *                   scanline order (captured_data.py:26-31).  The entry query then walks 4 x 8 (or 8 x 4) pixel
 *                   tiles per warp instead of 32 x 1 strips (needs a tile shape that divides the image and
 *                   N % (image_w * image_h) == 0, otherwise the hint is ignored).  Results do not depend on it.
 *   loss_sum        float64[1], ACCUMULATED into (caller zeroes)
 *   grad_V          float64[nV,3], ACCUMULATED into; NULL = loss value only
 *   n_paths         optional int32[1]: number of valid two-bounce paths of this batch (before the
 *                   `valid` filter) = mask.sum() of render_transparent
 *   ev_after_fwd    optional cudaEvent_t recorded on `stream` between the forward wavefront and the
 *                   loss/backward kernel (phase timing for bench.py); NULL = none
 * Scratch (72 B per ray of capacity) lives in the handle and is reused by later calls.
 */
DRT_API int drt_ray_loss_step(drt_bvh* bvh, const double* V64, const double* origin, int64_t rays_per_origin,
                      const double* dir, int64_t N, double ext_ior, double int_ior, int target_mode, const double* screen,
                      const uint8_t* valid, const int32_t* tgt_idx, const double* tgt_xyz, int64_t n_tgt, int32_t image_w,
                      int32_t image_h, double* loss_sum, double* grad_V, int32_t* n_paths, void* ev_after_fwd, void* stream);

/*
 * Per-tile direction intervals of a ray batch, prepared ONCE per view set.  DRT's view sets are fixed for a whole optimisation
 * (captured_data.py:94-108 loads them once, optim.py:95 cycles through them), and what the beam-culling pass of the entry query
 * needs to know about the 32 rays of a pixel tile -- their common origin and the per-axis interval of their float32 directions
 * (DiffRender.py:387-388 casts) -- depends on the rays only, never on the mesh.  drt_tile_beams computes it for a batch laid out
 * exactly as drt_ray_loss_step will receive it (same origin / rays_per_origin / dir / N / image_w / image_h) into
 * beams float32[drt_tile_beams_floats(N)] (device memory, 16-byte aligned; 1.5 B per ray); drt_ray_loss_step_beams is
 * drt_ray_loss_step with that buffer: its beam pass then reads 75 MB of intervals instead of re-reading 1.19 GB of ray directions
 * per step (C4).  tile_beams = NULL is drt_ray_loss_step.  The buffer carries a signature (N, image size, tile shape); a buffer
 * that does not match the call is ignored and the rays are scanned as usual.  It is the caller's promise that origin and dir still
 * hold the values the buffer was prepared from.  Results are identical either way (same intervals, same culling).
 */
DRT_API int64_t drt_tile_beams_floats(int64_t N);
DRT_API int drt_tile_beams(const double* origin, int64_t rays_per_origin, const double* dir, int64_t N, int32_t image_w, int32_t image_h,
                   float* beams, void* stream);
DRT_API int drt_ray_loss_step_beams(drt_bvh* bvh, const double* V64, const double* origin, int64_t rays_per_origin,
                      const double* dir, int64_t N, double ext_ior, double int_ior, int target_mode, const double* screen,
                      const uint8_t* valid, const int32_t* tgt_idx, const double* tgt_xyz, int64_t n_tgt, int32_t image_w,
                      int32_t image_h, const float* tile_beams, double* loss_sum, double* grad_V, int32_t* n_paths,
                      void* ev_after_fwd, void* stream);

/*
 * Replaces captured_data.generate_ray -- captured_data.py:23-40 (the pinhole sets build their rays
 * with it, captured_data.py:149): pixel (x, y, 1), x fastest, through K_inverse float64[3,3] and
 * R_inverse float64[4,4] (camera-to-world, row-major) -> dir float64[resy*resx,3] normalised and
 * origin3 float64[3] = the camera centre (the reference returns it expanded to [N,3]; pass it to
 * drt_ray_loss_step with rays_per_origin = resy*resx).  All pointers are device pointers.
 */
DRT_API int drt_generate_rays(int32_t resy, int32_t resx, const double* K_inverse, const double* R_inverse, double* origin3,
                      double* dir, void* stream);

/*
 * Silhouette-edge sampling, the second consumer of the plugin (optim.py:67-80 calls it for 8 views per
 * iteration).  All pointers are device pointers; float64 like the reference.
 *
 * drt_silhouette_classify replaces Scene.silhouette_edge -- DiffRender.py:445-457 (+ edge_face_norm :150-163):
 *   e2f int32[nE,2,3] = the vertex triples of the two faces on every edge (Scene.E2F, DiffRender.py:352-355), origin3 = the
 *   camera centre; flags[e] = 1 iff the two faces face opposite ways as seen from the origin (the reference's
 *   logical_xor(dot1 > 0, dot2 > 0)); the caller compacts Edges[flags] (order-preserving, as the reference's mask index does).
 *
 * drt_silhouette_sample replaces Scene.primary_visibility + primary_edge_sample.forward -- DiffRender.py:459-479, 189-241:
 *   for each of the k silhouette edges (edges int64[k,2] vertex ids): both ends projected with R float64[4,4] and
 *   K float64[3,3] (row-major), the midpoint sample, the image-space edge normal, the two probe rays one pixel either side
 *   (through K_inverse / R_inverse, un-normalised directions, float32 query like Scene.optix_intersect :386-392),
 *   index_xy int64[k,2] = the sample truncated to a pixel, f float64[k] = cover(upper) - cover(lower),
 *   keep uint8[k] = |f| > 1e-5 and 0 <= x < resx-1 and 0 <= y < resy-1 (:236, :476).  The handle is only read.
 *
 * drt_silhouette_backward replaces primary_edge_sample.backward (:243-267) chained through the projection (:465-472):
 *   kept_idx int64[m] = the edge slots whose samples were kept, g_output float32[m] = d loss / d output; accumulates
 *   d loss / d vertices into grad_V float64[nV,3] (caller zeroes).  detach_depth as in primary_visibility (:468-469).
 */
DRT_API int drt_silhouette_classify(const double* V64, const int32_t* e2f, int64_t nE, const double* origin3, uint8_t* flags,
                            void* stream);
DRT_API int drt_silhouette_sample(const drt_bvh* bvh, const double* V64, const int64_t* edges, int64_t k, const double* R,
                          const double* K, const double* R_inverse, const double* K_inverse, const double* origin3,
                          int32_t resx, int32_t resy, int64_t* index_xy, double* f, uint8_t* keep, void* stream);
DRT_API int drt_silhouette_backward(const double* V64, const int64_t* edges, const double* R, const double* K, int detach_depth,
                            const double* f, const int64_t* kept_idx, const float* g_output, int64_t m, double* grad_V,
                            void* stream);

/*
 * The two loss terms around the ray loss, fused like drt_ray_loss_step (value + vertex gradient from one launch, no host sync):
 *
 * drt_silhouette_loss replaces Loss_calculator.vh_loss -- optim.py:67-80 (per view: silhouette_edge + primary_visibility +
 *   primary_edge_sample + `(mask[index[:,1], index[:,0]] - output).abs().sum()`, summed over the views of the iteration, 8 in
 *   optim.py:72) and its backward: edges int64[nE,2] / e2f int32[nE,2,3] = Scene.Edges / Scene.E2F; R, K, R_inverse, K_inverse,
 *   origin3, mask = HOST arrays of n_views DEVICE pointers (camera_M of every view, its camera centre, its silhouette image
 *   float64[resy*resx]); adds the loss to *loss_sum, d loss / d vertices to grad_V (nullable) and the number of kept samples
 *   to *n_samples (nullable).  Up to 8 views run in one launch.
 * drt_dihedral_loss replaces Loss_calculator.sm_loss -- optim.py:82-89 (`-log(1 + dihedral_angle).sum()`, DiffRender.py:440-443)
 *   and its backward through both unit face normals of every edge.
 */
DRT_API int drt_silhouette_loss(const drt_bvh* bvh, const double* V64, const int64_t* edges, const int32_t* e2f, int64_t nE,
                        int32_t n_views, const double* const* R, const double* const* K, const double* const* R_inverse,
                        const double* const* K_inverse, const double* const* origin3, const double* const* mask,
                        int32_t resx, int32_t resy, int detach_depth, double* loss_sum, double* grad_V, int32_t* n_samples,
                        void* stream);
DRT_API int drt_dihedral_loss(const double* V64, const int32_t* e2f, int64_t nE, double* loss_sum, double* grad_V, void* stream);

/*
 * The one collective of the path -- SURVEY.md 8(e): views are sharded over the GPUs of one box, the mesh and
 * BVH are replicated, grad_V float64[nV,3] is summed once per step (the reference itself is single-GPU,
 * optix_extend.cpp:10, so there is no reference interface to mirror).  One-shot all-reduce over NVLink /
 * NVSwitch peer memory in ONE kernel launch: every rank pushes its gradient into its slot of an IPC-exported
 * staging area of every rank, publishes an epoch flag in every peer's memory, waits for the peers' flags and
 * adds the slots of its own area in rank order (all ranks get the same bits).  One process per GPU:
 *   drt_comm_create   allocates the local staging region for up to max_doubles values
 *   drt_comm_handle   -> 64-byte cudaIpcMemHandle_t of the region; exchange them with any host-side all-gather
 *   drt_comm_connect  handles = world x 64 bytes in rank order; opens the peers' regions
 *   drt_comm_allreduce_sum_f64   data float64[n] (device), in place, asynchronous on `stream`; every rank must
 *                     call it the same number of times with the same n
 *   drt_comm_status   synchronises `stream`; timed_out = 1 if a wait for a peer ever ran into the spin limit
 */
typedef struct drt_comm drt_comm;
DRT_API int drt_comm_create(int device, int rank, int world, int64_t max_doubles, drt_comm** out);
DRT_API int drt_comm_handle(const drt_comm* comm, unsigned char handle[64]);
DRT_API int drt_comm_connect(drt_comm* comm, const unsigned char* handles);
DRT_API int drt_comm_allreduce_sum_f64(drt_comm* comm, double* data, int64_t n, void* stream);
DRT_API int drt_comm_status(const drt_comm* comm, void* stream, int* timed_out);
DRT_API int drt_comm_destroy(drt_comm* comm);

/* Text of the last error on this thread ("" if none).  Never NULL. */
DRT_API const char* drt_last_error(void);

/*
 * OPTIONAL, NON-PARITY modes that BASELINE.json's north-star names and the reference does not run.
 *
 * Smooth-normal mode: the shading normal of a hit is n = normalize((1-u-v) n0 + u n1 + v n2) over the vertex normals of
 * the hit triangle with u, v detached -- the interpolation the reference keeps COMMENTED OUT in JIT_Dintersect
 * (DiffRender.py:107-114; the live code uses the flat face normal, :103-104).  VN64 = float64 [nV,3] vertex normals
 * (Scene.init_VN, DiffRender.py:319-336, on the caller's side).  Everything else is drt_trace_fwd / drt_trace_bwd:
 * same queries, same records; the backward additionally returns grad_VN float64 [nV,3] (accumulated, caller zeroes),
 * the gradient w.r.t. the vertex normals, while grad_V keeps the part through the hit distances.
 *
 * Background plane: where the exit ray of a valid path meets the plane through plane[0..2] with normal plane[3..5]
 * (host pointers): pts float64 [N,3] (zeros for rays without a valid path or without an intersection in front of them),
 * front uint8 [N] (nullable) = 1 where pts is a real intersection; drt_plane_hit_bwd maps g_pts to the gradients
 * w.r.t. out_ori / out_dir (written, not accumulated).  The reference's loss has no such step (optim.py:96-106).
 */
DRT_API int drt_trace_fwd_smooth(drt_bvh* bvh, const double* V64, const double* VN64, const double* origin, const double* dir, int64_t N,
                                 double ext_ior, double int_ior, double* out_ori, double* out_dir, uint8_t* mask3, int32_t* rec,
                                 int32_t* rec_count, void* stream);
DRT_API int drt_trace_bwd_smooth(const drt_bvh* bvh, const double* V64, const double* VN64, const double* origin, const double* dir,
                                 int64_t N, double ext_ior, double int_ior, const int32_t* rec, const int32_t* rec_count,
                                 const double* g_out_ori, const double* g_out_dir, double* grad_V, double* grad_VN, void* stream);
DRT_API int drt_plane_hit(const double* out_ori, const double* out_dir, const uint8_t* mask3, int64_t N, const double plane[6],
                          double* pts, uint8_t* front, void* stream);
DRT_API int drt_plane_hit_bwd(const double* out_ori, const double* out_dir, const uint8_t* mask3, int64_t N, const double plane[6],
                              const double* g_pts, double* g_out_ori, double* g_out_dir, void* stream);

/*
 * Runtime switch of a scheduling choice (results never depend on one).  No counterpart in the reference.
 *   "direct_max_rays"  drt_ray_loss_step batches of up to this many rays run the whole forward path of a ray in one
 *                      thread (2 launches per step) instead of the staged wavefront (8 launches): the reference renders ONE
 *                      view per iteration (optim.py:95), i.e. 10^5 .. 10^6 rays per call, where the stage launches and their
 *                      tails cost more than the divergence they remove.  Default 1 250 000 (env DRT_DIRECT_MAX); 0 = never.
 * Returns DRT_ERR_INVALID for an unknown key or a negative value.  Not thread-safe against concurrent calls of the library.
 */
DRT_API int drt_tuning_set(const char* key, long long value);
DRT_API long long drt_tuning_get(const char* key);

/* Number of kernels this library has launched in this process (monotonic; for bench accounting). */
DRT_API unsigned long long drt_kernel_launches(void);

/* Library/ABI version: major*1000 + minor. */
DRT_API int drt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DRT_B200_H */
