/*
 * oracle/drt_oracle.c -- CPU restatement of the DRT refraction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under drt_b200/ may import, link or call this file.
 * It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs as the checker and the reported CPU baseline -- never as the product path.
 *
 * What it restates (all citations are into /root/reference/):
 *   - optix_mesh::intersect            optix_extend.cpp:29-57   (closest hit, tmin=0, no culling;
 *                                      T f32 / ID i32, miss => T<0).  The arithmetic of that call
 *                                      lives in NVIDIA OptiX SDK 6.5.0 liboptix_prime (closed,
 *                                      absent; pinned only by config.py:3-4) -- its contract is
 *                                      geometric, so it is restated here as an exact closest hit:
 *                                      fp64 Moller-Trumbore on the fp32-rounded ray and fp32-rounded
 *                                      vertices, accept (float)t > 0 (DiffRender.py:391), ties go to
 *                                      the lowest triangle id.  PARITY UNPINNED for this stage (no
 *                                      reference test fixes Prime's edge/self-hit behaviour).
 *   - JIT_Dintersect                   DiffRender.py:64-121
 *   - Scene.refract_ray                DiffRender.py:503-535
 *   - FrDielectric (TIR flag only)     DiffRender.py:51-61
 *   - Refract (tan-law, bug-compatible)DiffRender.py:35-49
 *   - Scene.trace2 / render_transparent DiffRender.py:537-546, 420-432
 *   - the autograd graph of the above  (optim.py:210) as an analytic replay
 *
 * The differentiable chain is pinned against the reference's own functions by
 * oracle/make_golden.py (run in the build container, imports /root/reference/DiffRender.py
 * unmodified) -> tests/golden (npz files), checked by tests/test_oracle_golden.py.
 *
 * Also here: the CANONICAL LBVH (30-bit Morton of AABB centroids, split at the highest differing
 * bit, 1 triangle per leaf, near-child-first stack traversal with closest-hit pruning) whose
 * node/triangle visit counters define the algorithmic bytes per ray of the roofline
 * (DESIGN.md "Roofline denominator").
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: the CUDA triangle test uses __dmul_rn/__dadd_rn so that both sides
 * round identically and hit ids can be compared bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct { double x, y, z; } d3;

static inline d3 mk(double x, double y, double z) { d3 r = {x, y, z}; return r; }
static inline d3 sub3(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline d3 add3(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline d3 mul3(d3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline d3 neg3(d3 a) { return mk(-a.x, -a.y, -a.z); }
/* DiffRender.py:23-29 -- ((x*x + y*y) + z*z), left to right */
static inline double dot3(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline d3 cross3(d3 a, d3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline d3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
static inline d3 ld3f(const float* p) { return mk((double)p[0], (double)p[1], (double)p[2]); }
static inline void st3(double* p, d3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* ------------------------------------------------------------------------------------------
 * Query-stage triangle test (stand-in for the closed OptiX Prime query, optix_extend.cpp:33-45).
 * Inputs are the fp32-rounded ray and fp32-rounded vertices widened to double.
 * Returns 1 and *t when the ray crosses the closed triangle (u>=0, v>=0, u+v<=1) at (float)t > 0.
 * The CUDA kernel evaluates the identical expression tree (drt_b200/csrc/tri_test.cuh).
 * ------------------------------------------------------------------------------------------ */
static inline int query_tri(d3 o, d3 d, d3 a, d3 b, d3 c, double* t_out)
{
    d3 e1 = sub3(b, a), e2 = sub3(c, a);
    d3 pvec = cross3(d, e2);
    double det = dot3(e1, pvec);
    double inv = 1.0 / det;
    d3 tvec = sub3(o, a);
    double u = dot3(tvec, pvec) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return 0;
    d3 qvec = cross3(tvec, e1);
    double v = dot3(d, qvec) * inv;
    if (!(v >= 0.0 && (u + v) <= 1.0)) return 0;
    double t = dot3(e2, qvec) * inv;
    if (!((float)t > 0.0f)) return 0;
    *t_out = t;
    return 1;
}

/* ---------------------------------- canonical LBVH ---------------------------------------- */
typedef struct {
    float lo[3], hi[3];
    int32_t left, right; /* internal: child node indices; leaf: left = -1, right = triangle id */
} orc_node;

typedef struct orc_bvh {
    int32_t n_tris, n_verts, n_nodes, root;
    float* V;    /* fp32 vertices [nV,3] (copy) */
    int32_t* F;  /* faces [nF,3] (copy) */
    orc_node* nodes;
} orc_bvh;

static inline uint32_t expand10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

static int cmp_u64(const void* a, const void* b)
{
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

static int32_t build_range(orc_bvh* B, const uint64_t* keys, int32_t lo, int32_t hi, int32_t* next)
{
    int32_t me = (*next)++;
    orc_node* nd = &B->nodes[me];
    if (lo == hi) {
        int32_t tri = (int32_t)(keys[lo] & 0xffffffffu);
        nd->left = -1; nd->right = tri;
        for (int k = 0; k < 3; ++k) { nd->lo[k] = INFINITY; nd->hi[k] = -INFINITY; }
        for (int c = 0; c < 3; ++c) {
            const float* p = &B->V[3 * (size_t)B->F[3 * (size_t)tri + c]];
            for (int k = 0; k < 3; ++k) {
                if (p[k] < nd->lo[k]) nd->lo[k] = p[k];
                if (p[k] > nd->hi[k]) nd->hi[k] = p[k];
            }
        }
        return me;
    }
    /* split at the highest differing bit of the (unique) 64-bit keys: last index whose key has
       that bit clear (Karras 2012 sec. 3, restated top-down) */
    uint64_t diff = keys[lo] ^ keys[hi];
    int bit = 63 - __builtin_clzll(diff);
    uint64_t m = 1ull << bit;
    int32_t a = lo, b = hi; /* keys[a] bit clear, keys[b] bit set */
    while (b - a > 1) {
        int32_t mid = a + (b - a) / 2;
        if (keys[mid] & m) b = mid; else a = mid;
    }
    int32_t l = build_range(B, keys, lo, a, next);
    int32_t r = build_range(B, keys, b, hi, next);
    nd = &B->nodes[me];
    nd->left = l; nd->right = r;
    for (int k = 0; k < 3; ++k) {
        nd->lo[k] = fminf(B->nodes[l].lo[k], B->nodes[r].lo[k]);
        nd->hi[k] = fmaxf(B->nodes[l].hi[k], B->nodes[r].hi[k]);
    }
    return me;
}

ORC_API orc_bvh* orc_bvh_build(const float* V32, int32_t nV, const int32_t* F, int32_t nF)
{
    orc_bvh* B = (orc_bvh*)calloc(1, sizeof(orc_bvh));
    B->n_tris = nF; B->n_verts = nV;
    B->V = (float*)malloc(sizeof(float) * 3 * (size_t)(nV > 0 ? nV : 1));
    B->F = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(nF > 0 ? nF : 1));
    memcpy(B->V, V32, sizeof(float) * 3 * (size_t)nV);
    memcpy(B->F, F, sizeof(int32_t) * 3 * (size_t)nF);
    B->root = -1;
    if (nF <= 0) return B;
    /* centroid bounds */
    double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double* cen = (double*)malloc(sizeof(double) * 3 * (size_t)nF);
    for (int32_t f = 0; f < nF; ++f) {
        for (int k = 0; k < 3; ++k) {
            float a = V32[3 * (size_t)F[3 * f + 0] + k], b = V32[3 * (size_t)F[3 * f + 1] + k],
                  c = V32[3 * (size_t)F[3 * f + 2] + k];
            float lo = fminf(a, fminf(b, c)), hi = fmaxf(a, fmaxf(b, c));
            double m = 0.5 * ((double)lo + (double)hi);
            cen[3 * (size_t)f + k] = m;
            if (m < clo[k]) clo[k] = m;
            if (m > chi[k]) chi[k] = m;
        }
    }
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nF);
    for (int32_t f = 0; f < nF; ++f) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            double ext = chi[k] - clo[k];
            double u = ext > 0 ? (cen[3 * (size_t)f + k] - clo[k]) / ext : 0.0;
            double s = u * 1024.0;
            if (s < 0) s = 0;
            if (s > 1023.0) s = 1023.0;
            q[k] = (uint32_t)s;
        }
        uint32_t morton = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
        keys[f] = ((uint64_t)morton << 32) | (uint32_t)f;
    }
    qsort(keys, (size_t)nF, sizeof(uint64_t), cmp_u64);
    B->nodes = (orc_node*)malloc(sizeof(orc_node) * (size_t)(2 * nF - 1));
    int32_t next = 0;
    B->root = build_range(B, keys, 0, nF - 1, &next);
    B->n_nodes = next;
    free(keys); free(cen);
    return B;
}

ORC_API void orc_bvh_free(orc_bvh* B)
{
    if (!B) return;
    free(B->V); free(B->F); free(B->nodes); free(B);
}

ORC_API int32_t orc_bvh_num_nodes(const orc_bvh* B) { return B->n_nodes; }

/* Slab test in double on the float box, closed-box semantics.  Near/far planes are chosen by the
   sign of the direction; a zero direction component gives +-inf (origin strictly inside/outside
   the slab) or NaN (origin exactly on a slab plane), and fmax/fmin drop the NaN = "no constraint",
   which is the correct answer for a ray travelling inside a face plane.  The 1e-9 slack makes the
   exact triangle test, never the box test, decide every hit. */
static inline int box_test(const orc_node* nd, d3 o, d3 inv, double tbest, double* tnear)
{
    double t0 = 0.0, t1 = tbest;
    const double oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    for (int k = 0; k < 3; ++k) {
        int neg = signbit(ii[k]);
        double pn = (double)(neg ? nd->hi[k] : nd->lo[k]), pf = (double)(neg ? nd->lo[k] : nd->hi[k]);
        t0 = fmax(t0, (pn - oo[k]) * ii[k]);
        t1 = fmin(t1, (pf - oo[k]) * ii[k]);
    }
    *tnear = t0;
    return t0 * (1.0 - 1e-9) <= t1 * (1.0 + 1e-9);
}

typedef struct { int64_t nodes, tris; } orc_counters;

static void closest_bvh_one(const orc_bvh* B, d3 o, d3 d, double* t_out, int32_t* id_out, orc_counters* cnt)
{
    double best = INFINITY; int32_t best_id = -1;
    if (B->root >= 0) {
        d3 inv = mk(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
        int32_t stack[128]; int sp = 0;
        double tn;
        cnt->nodes++;
        if (box_test(&B->nodes[B->root], o, inv, best, &tn)) stack[sp++] = B->root;
        while (sp > 0) {
            const orc_node* nd = &B->nodes[stack[--sp]];
            if (nd->left < 0) {
                int32_t tri = nd->right;
                const int32_t* f = &B->F[3 * (size_t)tri];
                double t;
                cnt->tris++;
                if (query_tri(o, d, ld3f(&B->V[3 * (size_t)f[0]]), ld3f(&B->V[3 * (size_t)f[1]]),
                              ld3f(&B->V[3 * (size_t)f[2]]), &t)) {
                    if (t < best || (t == best && tri < best_id)) { best = t; best_id = tri; }
                }
                continue;
            }
            double tl, tr;
            cnt->nodes += 2;
            int hl = box_test(&B->nodes[nd->left], o, inv, best, &tl);
            int hr = box_test(&B->nodes[nd->right], o, inv, best, &tr);
            if (hl && hr) {
                if (tl <= tr) { stack[sp++] = nd->right; stack[sp++] = nd->left; }
                else          { stack[sp++] = nd->left;  stack[sp++] = nd->right; }
            } else if (hl) stack[sp++] = nd->left;
            else if (hr) stack[sp++] = nd->right;
        }
    }
    *t_out = best; *id_out = best_id;
}

static void closest_brute_one(const float* V, const int32_t* F, int32_t nF, d3 o, d3 d, double* t_out, int32_t* id_out)
{
    double best = INFINITY; int32_t best_id = -1;
    for (int32_t tri = 0; tri < nF; ++tri) {
        const int32_t* f = &F[3 * (size_t)tri];
        double t;
        if (query_tri(o, d, ld3f(&V[3 * (size_t)f[0]]), ld3f(&V[3 * (size_t)f[1]]), ld3f(&V[3 * (size_t)f[2]]), &t)) {
            if (t < best) { best = t; best_id = tri; } /* ascending ids: ties keep the lowest */
        }
    }
    *t_out = best; *id_out = best_id;
}

/* optix_mesh::intersect restated (optix_extend.cpp:29-57): ray6 f32[N,6] -> T f32[N], ID i32[N];
   miss => T=-1, ID=-1.  B==NULL or use_bvh==0 -> brute force over (V32,F). */
ORC_API void orc_closest_hit(const orc_bvh* B, int use_bvh, const float* ray6, int64_t N, float* T, int32_t* ID,
                             int64_t* nodes_visited, int64_t* tris_tested)
{
    int64_t tn = 0, tt = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : tn, tt)
    for (int64_t i = 0; i < N; ++i) {
        d3 o = ld3f(&ray6[6 * i]), d = ld3f(&ray6[6 * i + 3]);
        double t; int32_t id; orc_counters c = {0, 0};
        if (use_bvh) closest_bvh_one(B, o, d, &t, &id, &c);
        else closest_brute_one(B->V, B->F, B->n_tris, o, d, &t, &id);
        T[i] = id >= 0 ? (float)t : -1.0f;
        ID[i] = id;
        tn += c.nodes; tt += c.tris;
    }
    if (nodes_visited) *nodes_visited = tn;
    if (tris_tested) *tris_tested = tt;
}

/* ------------------------------ differentiable chain, forward ----------------------------- */
typedef struct {
    d3 o, d;          /* incoming ray */
    d3 a0, e1, e2;    /* triangle */
    d3 N, n, np;      /* N = e1 x e2, n = N/L, np = oriented normal n' */
    double L, t, D;   /* |N|, hit distance, D = d.N */
    double sgn;       /* +1 entering (n' = n), -1 exiting (n' = -n) */
    double eta, c, cT, A, nw;
    int cT_grad;      /* 1 if d cT/dc = c/cT (clamp open), 0 if the clamp(min=0) cut the gradient */
    d3 w, wt, x;
    int tir;
} hit_rec;

/* One surface interaction: JIT_Dintersect (DiffRender.py:64-121) followed by refract_ray
   (DiffRender.py:503-535).  Returns the new ray in (o2,d2). */
static void hit_forward(hit_rec* h, d3 o, d3 d, d3 a0, d3 a1, d3 a2, double ext_ior, double int_ior, d3* o2, d3* d2)
{
    h->o = o; h->d = d; h->a0 = a0;
    /* DiffRender.py:71-91 */
    h->e1 = sub3(a1, a0); h->e2 = sub3(a2, a0);
    d3 pvec = cross3(d, h->e2);
    double det = dot3(h->e1, pvec);
    double inv_det = 1.0 / det;
    d3 tvec = sub3(o, a0);
    d3 qvec = cross3(tvec, h->e1);
    h->t = dot3(h->e2, qvec) * inv_det;
    /* DiffRender.py:103-104 flat face normal */
    h->N = cross3(h->e1, h->e2);
    h->L = sqrt((h->N.x * h->N.x + h->N.y * h->N.y) + h->N.z * h->N.z);
    h->n = mk(h->N.x / h->L, h->N.y / h->L, h->N.z / h->L);
    h->D = dot3(d, h->N);
    /* DiffRender.py:508-519 */
    d3 wo = neg3(d);
    double c0 = dot3(wo, h->n);
    double cc = c0 < -1.0 ? -1.0 : (c0 > 1.0 ? 1.0 : c0);
    int entering = cc > 0.0;
    double etaI = entering ? ext_ior : int_ior, etaT = entering ? int_ior : ext_ior;
    h->sgn = entering ? 1.0 : -1.0;
    h->np = entering ? h->n : neg3(h->n);
    double cf = entering ? cc : -cc;
    /* FrDielectric, DiffRender.py:54-56 -- only the TIR flag is live (:526) */
    double s = 1.0 - cf * cf; s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
    double sinI = sqrt(s);
    double sinT = sinI * etaI / etaT;
    h->tir = sinT >= 1.0;
    /* Refract, DiffRender.py:37-47 (tan-law: cosThetaT uses sin2ThetaI) */
    h->eta = etaI / etaT;
    h->c = dot3(h->np, wo);
    double s2 = 1.0 - h->c * h->c;
    h->cT_grad = s2 >= 0.0; /* clamp(min=0) passes gradient where input >= min */
    if (s2 < 0.0) s2 = 0.0;
    double s2c = s2 > 1.0 ? 1.0 : s2;
    h->cT = sqrt(1.0 - s2c);
    h->A = h->eta * h->c - h->cT;
    h->w = add3(mul3(neg3(wo), h->eta), mul3(h->np, h->A));
    h->nw = sqrt((h->w.x * h->w.x + h->w.y * h->w.y) + h->w.z * h->w.z);
    h->wt = mk(h->w.x / h->nw, h->w.y / h->nw, h->w.z / h->nw);
    /* DiffRender.py:528-532 */
    h->x = add3(o, mul3(d, h->t));
    *o2 = add3(h->x, mul3(h->wt, 1e-5));
    *d2 = h->wt;
}

/* Analytic reverse of hit_forward.  (go2,gd2) = grad wrt the outgoing ray;
   accumulates grads of a0,a1,a2 into ga[3] and returns grads wrt the incoming ray in (go,gd). */
static void hit_backward(const hit_rec* h, d3 go2, d3 gd2, d3 ga[3], d3* go, d3* gd)
{
    d3 g_wt = add3(gd2, mul3(go2, 1e-5));
    d3 g_x = go2;
    d3 g_o = g_x;
    double g_t = dot3(g_x, h->d);
    d3 g_d = mul3(g_x, h->t);
    /* wt = w/|w| */
    d3 g_w = mul3(sub3(g_wt, mul3(h->wt, dot3(h->wt, g_wt))), 1.0 / h->nw);
    /* w = eta*d + A*n' */
    g_d = add3(g_d, mul3(g_w, h->eta));
    double g_A = dot3(g_w, h->np);
    d3 g_np = mul3(g_w, h->A);
    /* A = eta*c - cT(c);  cT = sqrt(1 - clamp(1-c^2, 0, 1)) */
    double dcT = h->cT_grad ? h->c / h->cT : 0.0;
    double g_c = g_A * (h->eta - dcT);
    /* c = n'.wo = -(n'.d) */
    g_d = sub3(g_d, mul3(h->np, g_c));
    g_np = sub3(g_np, mul3(h->d, g_c));
    d3 g_n = mul3(g_np, h->sgn);
    /* n = N/L */
    d3 g_N = mul3(sub3(g_n, mul3(h->n, dot3(h->n, g_n))), 1.0 / h->L);
    /* t = ((a0-o).N)/(d.N) */
    double k = g_t / h->D;
    d3 g_a0 = mul3(h->N, k);
    g_o = sub3(g_o, mul3(h->N, k));
    g_N = add3(g_N, mul3(sub3(h->a0, h->x), k));
    g_d = sub3(g_d, mul3(h->N, k * h->t));
    /* N = e1 x e2 */
    d3 g_e1 = cross3(h->e2, g_N);
    d3 g_e2 = cross3(g_N, h->e1);
    ga[1] = add3(ga[1], g_e1);
    ga[2] = add3(ga[2], g_e2);
    ga[0] = add3(ga[0], sub3(sub3(g_a0, g_e1), g_e2));
    *go = g_o; *gd = g_d;
}

static inline void query_ray(const orc_bvh* B, int use_bvh, d3 o, d3 d, double* t, int32_t* id, orc_counters* c)
{
    /* Scene.optix_intersect, DiffRender.py:386-392: the query sees the fp32 cast of the ray */
    d3 of = mk((double)(float)o.x, (double)(float)o.y, (double)(float)o.z);
    d3 df = mk((double)(float)d.x, (double)(float)d.y, (double)(float)d.z);
    if (use_bvh) closest_bvh_one(B, of, df, t, id, c);
    else closest_brute_one(B->V, B->F, B->n_tris, of, df, t, id);
}

/*
 * Scene.render_transparent (DiffRender.py:420-432) for N rays.
 *   B        query structure built from the fp32 cast of the vertices (DiffRender.py:311,379)
 *   V64      fp64 vertices used by the differentiable re-intersection (DiffRender.py:495-496)
 *   tri1/2   optional hit records (triangle ids of hit 1 / hit 2, -1 where the path died)
 *   stage    optional uint8[N]: 0 Q1 miss, 1 TIR1, 2 Q2 miss, 3 TIR2, 4 Q3 hit (rejected), 5 valid
 *   counters optional int64[6]: nodes,tris for Q1, Q2, Q3 (canonical BVH only)
 */
ORC_API void orc_trace_fwd(const orc_bvh* B, int use_bvh, const double* V64, const double* o_in, const double* d_in,
                           int64_t N, double ext_ior, double int_ior, double* out_ori, double* out_dir,
                           uint8_t* mask3, int32_t* tri1, int32_t* tri2, uint8_t* stage, int64_t* counters)
{
    int64_t c0n = 0, c0t = 0, c1n = 0, c1t = 0, c2n = 0, c2t = 0;
    const int32_t* F = B->F;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c0n, c0t, c1n, c1t, c2n, c2t)
    for (int64_t i = 0; i < N; ++i) {
        d3 o = ld3(&o_in[3 * i]), d = ld3(&d_in[3 * i]);
        d3 zero = mk(0, 0, 0);
        int32_t id1 = -1, id2 = -1, id3 = -1; double t; int st = 0; int valid = 0;
        d3 oo = zero, od = zero;
        orc_counters c = {0, 0};
        query_ray(B, use_bvh, o, d, &t, &id1, &c); c0n += c.nodes; c0t += c.tris;
        if (id1 >= 0) {
            hit_rec h1, h2; d3 o1, d1, o2, d2;
            const int32_t* f = &F[3 * (size_t)id1];
            hit_forward(&h1, o, d, ld3(&V64[3 * (size_t)f[0]]), ld3(&V64[3 * (size_t)f[1]]), ld3(&V64[3 * (size_t)f[2]]),
                        ext_ior, int_ior, &o1, &d1);
            st = 1;
            if (!h1.tir) {
                c.nodes = c.tris = 0;
                query_ray(B, use_bvh, o1, d1, &t, &id2, &c); c1n += c.nodes; c1t += c.tris;
                st = 2;
                if (id2 >= 0) {
                    f = &F[3 * (size_t)id2];
                    hit_forward(&h2, o1, d1, ld3(&V64[3 * (size_t)f[0]]), ld3(&V64[3 * (size_t)f[1]]),
                                ld3(&V64[3 * (size_t)f[2]]), ext_ior, int_ior, &o2, &d2);
                    st = 3;
                    if (!h2.tir) {
                        c.nodes = c.tris = 0;
                        query_ray(B, use_bvh, o2, d2, &t, &id3, &c); c2n += c.nodes; c2t += c.tris;
                        st = 4;
                        if (id3 < 0) { st = 5; valid = 1; oo = o2; od = d2; }
                    }
                }
            }
        }
        st3(&out_ori[3 * i], oo); st3(&out_dir[3 * i], od);
        if (mask3) mask3[3 * i] = mask3[3 * i + 1] = mask3[3 * i + 2] = (uint8_t)valid;
        if (tri1) tri1[i] = valid ? id1 : -1;
        if (tri2) tri2[i] = valid ? id2 : -1;
        if (stage) stage[i] = (uint8_t)st;
    }
    if (counters) { counters[0] = c0n; counters[1] = c0t; counters[2] = c1n; counters[3] = c1t; counters[4] = c2n; counters[5] = c2t; }
}

/*
 * Forward of the differentiable part only, with the hit ids GIVEN (Tier-B parity):
 * ids come from whatever intersector is under test.  tri1/tri2 < 0 => ray invalid (zeros).
 * No TIR / Q3 logic: the caller's records already encode validity; tir flags are returned.
 */
ORC_API void orc_chain_fwd(const double* V64, const int32_t* F, const double* o_in, const double* d_in, int64_t N,
                           double ext_ior, double int_ior, const int32_t* tri1, const int32_t* tri2,
                           double* out_ori, double* out_dir, uint8_t* tir_flags)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        d3 oo = mk(0, 0, 0), od = mk(0, 0, 0); uint8_t tf = 0;
        if (tri1[i] >= 0 && tri2[i] >= 0) {
            hit_rec h1, h2; d3 o1, d1, o2, d2;
            const int32_t* f = &F[3 * (size_t)tri1[i]];
            hit_forward(&h1, ld3(&o_in[3 * i]), ld3(&d_in[3 * i]), ld3(&V64[3 * (size_t)f[0]]), ld3(&V64[3 * (size_t)f[1]]),
                        ld3(&V64[3 * (size_t)f[2]]), ext_ior, int_ior, &o1, &d1);
            f = &F[3 * (size_t)tri2[i]];
            hit_forward(&h2, o1, d1, ld3(&V64[3 * (size_t)f[0]]), ld3(&V64[3 * (size_t)f[1]]), ld3(&V64[3 * (size_t)f[2]]),
                        ext_ior, int_ior, &o2, &d2);
            oo = o2; od = d2; tf = (uint8_t)(h1.tir | (h2.tir << 1));
        }
        st3(&out_ori[3 * i], oo); st3(&out_dir[3 * i], od);
        if (tir_flags) tir_flags[i] = tf;
    }
}

/*
 * Reverse pass: d(sum(out_ori*g_ori + out_dir*g_dir))/dV, the same quantity loss.backward()
 * (optim.py:210) leaves in vertices.grad.  grad_V [nV,3] is ACCUMULATED into (caller zeroes).
 * g_ori may be NULL (optim.py:100 detaches out_ori).
 */
ORC_API void orc_trace_bwd(const double* V64, const int32_t* F, int32_t nV, const double* o_in, const double* d_in, int64_t N,
                           double ext_ior, double int_ior, const int32_t* tri1, const int32_t* tri2,
                           const double* g_ori, const double* g_dir, double* grad_V)
{
    int nth = 1;
#ifdef _OPENMP
    nth = omp_get_max_threads();
#endif
    double* priv = (double*)calloc((size_t)nth * 3 * (size_t)nV, sizeof(double));
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double* g = priv + (size_t)tid * 3 * (size_t)nV;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            if (tri1[i] < 0 || tri2[i] < 0) continue;
            hit_rec h1, h2; d3 o1, d1, o2, d2;
            const int32_t* f1 = &F[3 * (size_t)tri1[i]];
            const int32_t* f2 = &F[3 * (size_t)tri2[i]];
            hit_forward(&h1, ld3(&o_in[3 * i]), ld3(&d_in[3 * i]), ld3(&V64[3 * (size_t)f1[0]]), ld3(&V64[3 * (size_t)f1[1]]),
                        ld3(&V64[3 * (size_t)f1[2]]), ext_ior, int_ior, &o1, &d1);
            hit_forward(&h2, o1, d1, ld3(&V64[3 * (size_t)f2[0]]), ld3(&V64[3 * (size_t)f2[1]]), ld3(&V64[3 * (size_t)f2[2]]),
                        ext_ior, int_ior, &o2, &d2);
            d3 go2 = g_ori ? ld3(&g_ori[3 * i]) : mk(0, 0, 0);
            d3 gd2 = ld3(&g_dir[3 * i]);
            d3 ga[3] = {mk(0, 0, 0), mk(0, 0, 0), mk(0, 0, 0)}, go1, gd1, go0, gd0;
            hit_backward(&h2, go2, gd2, ga, &go1, &gd1);
            for (int c = 0; c < 3; ++c) {
                double* p = &g[3 * (size_t)f2[c]];
                p[0] += ga[c].x; p[1] += ga[c].y; p[2] += ga[c].z;
                ga[c] = mk(0, 0, 0);
            }
            hit_backward(&h1, go1, gd1, ga, &go0, &gd0);
            for (int c = 0; c < 3; ++c) {
                double* p = &g[3 * (size_t)f1[c]];
                p[0] += ga[c].x; p[1] += ga[c].y; p[2] += ga[c].z;
            }
        }
    }
    for (int t = 0; t < nth; ++t)
        for (size_t j = 0; j < 3 * (size_t)nV; ++j) grad_V[j] += priv[(size_t)t * 3 * (size_t)nV + j];
    free(priv);
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
