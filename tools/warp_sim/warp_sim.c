/*
 * warp_sim.c -- CPU model of the persistent query kernels' WARP SCHEDULING (developer tool, not product,
 * not the oracle): how many warp-wide node steps / leaf tests a batch policy issues for a given ray list.
 *
 * It builds its own plain LBVH (30-bit Morton, split at the highest differing bit, 1 triangle per leaf),
 * runs the same per-lane state machine as drt_b200/csrc/trace.cuh (near-first binary traversal, deferred
 * leaves, closest-hit pruning) for 32 lanes in lockstep, and charges a warp-wide cost per SIMT iteration:
 *     C_NODE  if any lane executes a node step, C_PUSH if any lane queues a leaf,
 *     C_LEAF  per drain iteration in which any lane tests a triangle.
 * Policies differ in when a warp leaves the walk loop, when it drains and when finished lanes are refilled.
 * Only RELATIVE numbers between policies are meaningful (no memory latency, no issue-slot model).
 *
 * usage: warp_sim mesh.bin rays.bin   (see tools/warp_sim/make_inputs.py for the formats)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } box;
typedef struct { box b[2]; int c[2]; } node;   /* c >= 0 node, < 0: ~triangle slot */
typedef struct { double a[3], e1[3], e2[3]; int id; } tri;

static int nV, nF, nN;
static float* V; static int* F;
static node* N; static tri* T; static unsigned char* depth_of;  /* depth of every binary node (root = 0) */
static uint64_t* keys;

static uint32_t spread10(uint32_t v) { v &= 1023; v = (v | v << 16) & 0x30000ff; v = (v | v << 8) & 0x300f00f; v = (v | v << 4) & 0x30c30c3; v = (v | v << 2) & 0x9249249; return v; }
static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }

static void tri_box(int f, box* b)
{
    for (int k = 0; k < 3; ++k) {
        float x = V[3 * F[3 * f] + k], y = V[3 * F[3 * f + 1] + k], z = V[3 * F[3 * f + 2] + k];
        b->lo[k] = fminf(x, fminf(y, z)); b->hi[k] = fmaxf(x, fmaxf(y, z));
    }
}
static void merge(box* o, const box* a, const box* b) { for (int k = 0; k < 3; ++k) { o->lo[k] = fminf(a->lo[k], b->lo[k]); o->hi[k] = fmaxf(a->hi[k], b->hi[k]); } }

/* recursive split of sorted key range [l, r]; returns link, fills box */
static int cur_depth = 0;
static int build(int l, int r, box* out)
{
    if (l == r) { tri_box((int)(keys[l] & 0xffffffffu), out); return ~l; }
    uint64_t x = keys[l] ^ keys[r];
    int bit = 63 - __builtin_clzll(x);
    int lo = l, hi = r;  /* last index with that bit clear */
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if ((keys[mid] >> bit) & 1) hi = mid - 1; else lo = mid; }
    int me = nN++;
    depth_of[me] = (unsigned char)(cur_depth > 255 ? 255 : cur_depth);
    box b0, b1;
    ++cur_depth;
    int c0 = build(l, lo, &b0), c1 = build(lo + 1, r, &b1);
    --cur_depth;
    N[me].b[0] = b0; N[me].b[1] = b1; N[me].c[0] = c0; N[me].c[1] = c1;
    merge(out, &b0, &b1);
    return me;
}

static void build_bvh(void)
{
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    float* cen = malloc(sizeof(float) * 3 * nF);
    for (int f = 0; f < nF; ++f) { box b; tri_box(f, &b); for (int k = 0; k < 3; ++k) { float c = 0.5f * b.lo[k] + 0.5f * b.hi[k]; cen[3 * f + k] = c; mn[k] = fminf(mn[k], c); mx[k] = fmaxf(mx[k], c); } }
    keys = malloc(sizeof(uint64_t) * nF);
    for (int f = 0; f < nF; ++f) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) { float e = mx[k] - mn[k]; float u = e > 0 ? (cen[3 * f + k] - mn[k]) / e : 0; float s = fminf(fmaxf(u * 1024.f, 0.f), 1023.f); q[k] = (uint32_t)s; }
        uint64_t m = ((uint64_t)spread10(q[0]) << 2) | ((uint64_t)spread10(q[1]) << 1) | spread10(q[2]);
        keys[f] = (m << 32) | (uint32_t)f;
    }
    qsort(keys, nF, sizeof(uint64_t), cmp_u64);
    N = malloc(sizeof(node) * (nF > 1 ? nF - 1 : 1)); nN = 0;
    depth_of = calloc(nF > 1 ? nF - 1 : 1, 1);
    T = malloc(sizeof(tri) * nF);
    for (int k = 0; k < nF; ++k) {
        int f = (int)(keys[k] & 0xffffffffu);
        for (int c = 0; c < 3; ++c) { T[k].a[c] = V[3 * F[3 * f] + c]; T[k].e1[c] = (double)V[3 * F[3 * f + 1] + c] - T[k].a[c]; T[k].e2[c] = (double)V[3 * F[3 * f + 2] + c] - T[k].a[c]; }
        T[k].id = f;
    }
    box root; build(0, nF - 1, &root);
    free(cen);
}

/* ---- optional 4-wide collapse of the binary tree (SIM_BVH4=1): every node absorbs its internal children ------------ */
typedef struct { box b[4]; int c[4]; int n; } node4;
static node4* N4; static int* map4;  /* binary node -> wide node index */
static int use4 = 0;
static int hot_depth = 10;  /* nodes at depth <= hot_depth (2047 nodes, 64 KB) are taken as L1 hits; SIM_HOT_DEPTH */

static int collapse(int n)
{
    int me = map4[n];
    node4* w = &N4[me];
    w->n = 0;
    for (int i = 0; i < 2; ++i) {
        int c = N[n].c[i];
        if (c >= 0) {  /* internal child: take ITS two children */
            for (int j = 0; j < 2; ++j) { w->b[w->n] = N[c].b[j]; w->c[w->n] = N[c].c[j]; ++w->n; }
        } else { w->b[w->n] = N[n].b[i]; w->c[w->n] = c; ++w->n; }
    }
    for (int k = 0; k < w->n; ++k)
        if (w->c[k] >= 0) { int child = w->c[k]; w->c[k] = map4[child]; collapse(child); }
    return me;
}

static void build_bvh4(void)
{
    N4 = malloc(sizeof(node4) * (nN > 0 ? nN : 1));
    map4 = malloc(sizeof(int) * (nN > 0 ? nN : 1));
    for (int i = 0; i < nN; ++i) map4[i] = i;   /* sparse reuse of binary indices: only visited ones matter */
    if (nN > 0) collapse(0);
}

/* ---- per-lane traversal state ---------------------------------------------------------------- */
#define STACK 128
#define MAXDEFER 16
typedef struct {
    int item;            /* ray index or -1 */
    float o[3], d[3], inv[3];
    float tmax; double tbest; int idbest;
    int node;            /* current link; DONE when finished walking */
    int sp, nd;
    int stack[STACK], q[MAXDEFER];
} lane;
#define DONE INT32_MIN

static int node_step4(lane* L)
{
    const node4* n = &N4[L->node];
    float tn[4]; int id[4], m = 0;
    for (int c = 0; c < n->n; ++c) {
        float a = 0.f, b = L->tmax;
        for (int k = 0; k < 3; ++k) {
            float t0 = (n->b[c].lo[k] - L->o[k]) * L->inv[k], t1 = (n->b[c].hi[k] - L->o[k]) * L->inv[k];
            a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
        }
        if (a <= b * 1.000001f) { int j = m++; while (j > 0 && tn[j - 1] > a) { tn[j] = tn[j - 1]; id[j] = id[j - 1]; --j; } tn[j] = a; id[j] = n->c[c]; }
    }
    if (!m) return L->sp ? L->stack[--L->sp] : DONE;
    for (int j = m - 1; j >= 1; --j) L->stack[L->sp++] = id[j];   /* farthest first, nearest popped first */
    return id[0];
}

static int node_step(lane* L)
{
    if (use4) return node_step4(L);
    const node* n = &N[L->node];
    float tn[2], tf[2]; int h[2];
    for (int c = 0; c < 2; ++c) {
        float a = 0.f, b = L->tmax;
        for (int k = 0; k < 3; ++k) {
            float t0 = (n->b[c].lo[k] - L->o[k]) * L->inv[k], t1 = (n->b[c].hi[k] - L->o[k]) * L->inv[k];
            a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
        }
        tn[c] = a; tf[c] = b; h[c] = a <= b * 1.000001f;
    }
    if (h[0] && h[1]) { int f0 = tn[0] <= tn[1]; L->stack[L->sp++] = f0 ? n->c[1] : n->c[0]; return f0 ? n->c[0] : n->c[1]; }
    if (h[0]) return n->c[0];
    if (h[1]) return n->c[1];
    return L->sp ? L->stack[--L->sp] : DONE;
}

static int leaf_test(lane* L, int leaf, int any)
{
    const tri* t = &T[~leaf];
    double o[3] = {L->o[0], L->o[1], L->o[2]}, d[3] = {L->d[0], L->d[1], L->d[2]};
    double p[3] = {d[1] * t->e2[2] - d[2] * t->e2[1], d[2] * t->e2[0] - d[0] * t->e2[2], d[0] * t->e2[1] - d[1] * t->e2[0]};
    double det = t->e1[0] * p[0] + t->e1[1] * p[1] + t->e1[2] * p[2];
    if (det == 0) return 0;
    double inv = 1.0 / det, tv[3] = {o[0] - t->a[0], o[1] - t->a[1], o[2] - t->a[2]};
    double u = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) * inv;
    if (u < 0 || u > 1) return 0;
    double q[3] = {tv[1] * t->e1[2] - tv[2] * t->e1[1], tv[2] * t->e1[0] - tv[0] * t->e1[2], tv[0] * t->e1[1] - tv[1] * t->e1[0]};
    double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
    if (v < 0 || u + v > 1) return 0;
    double tt = (t->e2[0] * q[0] + t->e2[1] * q[1] + t->e2[2] * q[2]) * inv;
    if (!((float)tt > 0.f)) return 0;
    if (tt < L->tbest || (tt == L->tbest && t->id < L->idbest)) { L->tbest = tt; L->idbest = t->id; L->tmax = (float)tt * 1.0000002f; }
    (void)any;
    return 1;
}

/* ---- policy ------------------------------------------------------------------------------------ */
typedef struct {
    int defer;        /* leaf queue capacity per lane */
    int refill;       /* refill when this many of the live lanes have finished (32 = whole warp) */
    int vote_drain;   /* 0: leave the walk loop when EVERY lane is blocked/done (the kernel today);
                         k>0: leave it as soon as k lanes are blocked (queue full, or done with queued leaves) */
    int any;          /* any-hit query */
    double c_node, c_push, c_leaf, c_refill;
} policy;

typedef struct { double cost, node_iters, node_lane_steps, leaf_iters, leaf_lane_tests, refills, node_lines, leaf_lines, deep_iters; long rays, hits; } stats;

static const float* RO; static const float* RD; /* ray arrays */

static void lane_load(lane* L, int item)
{
    L->item = item;
    for (int k = 0; k < 3; ++k) { L->o[k] = RO[3 * item + k]; L->d[k] = RD[3 * item + k]; L->inv[k] = fabsf(L->d[k]) < 1e-30f ? copysignf(1e30f, L->d[k]) : 1.f / L->d[k]; }
    L->tmax = INFINITY; L->tbest = INFINITY; L->idbest = -1; L->node = 0; L->sp = 0; L->nd = 0;
}

/* one persistent warp over rays [first, last) */
static void run_warp(int first, int last, const policy* P, stats* S, int* out_id, double* out_t)
{
    lane W[32];
    for (int l = 0; l < 32; ++l) W[l].item = -1;
    int next = first;
    for (;;) {
        /* refill idle lanes in lane order from consecutive rays */
        int filled = 0;
        for (int l = 0; l < 32 && next < last; ++l) if (W[l].item < 0) { lane_load(&W[l], next++); ++filled; }
        if (filled) { S->cost += P->c_refill; S->refills += 1; }
        int live = 0;
        for (int l = 0; l < 32; ++l) live += W[l].item >= 0;
        if (!live) break;
        int need = P->refill < live ? P->refill : live;
        if (next >= last) need = live;  /* nothing left to fetch: run to completion */
        for (;;) {
            /* ---- walk phase ---- */
            for (;;) {
                int any_node = 0, any_push = 0, blocked = 0, can = 0;
                int seen[32], nseen = 0;   /* distinct nodes fetched by the warp in this iteration ~ L1 wavefronts per load instruction */
                int deep = 0;              /* does any lane fetch a node below the L1-hot top of the tree this iteration? */
                for (int l = 0; l < 32; ++l) {
                    lane* L = &W[l];
                    if (L->item < 0 || L->node == DONE || L->nd >= P->defer) continue;
                    ++can;
                    if (L->node >= 0) {
                        int dup = 0;
                        for (int k = 0; k < nseen; ++k) dup |= seen[k] == L->node;
                        if (!dup) seen[nseen++] = L->node;
                        if (!use4 && depth_of[L->node] > hot_depth) deep = 1;
                        L->node = node_step(L); any_node = 1; S->node_lane_steps += 1;
                    }
                    else { L->q[L->nd++] = L->node; L->node = L->sp ? L->stack[--L->sp] : DONE; any_push = 1; }
                }
                if (!can) break;
                S->node_lines += nseen;
                S->deep_iters += deep;
                S->cost += any_node * P->c_node + any_push * P->c_push;
                S->node_iters += any_node;
                if (P->vote_drain > 0) {
                    for (int l = 0; l < 32; ++l) { lane* L = &W[l]; if (L->item >= 0 && L->nd > 0 && (L->nd >= P->defer || L->node == DONE)) ++blocked; }
                    if (blocked >= P->vote_drain) break;
                }
            }
            /* ---- drain phase: every lane tests its queued leaves ---- */
            for (;;) {
                int any_leaf = 0;
                for (int l = 0; l < 32; ++l) {
                    lane* L = &W[l];
                    if (L->item < 0 || L->nd == 0) continue;
                    int hit = leaf_test(L, L->q[--L->nd], P->any);
                    any_leaf = 1; S->leaf_lane_tests += 1;
                    if (P->any && hit) { L->nd = 0; L->node = DONE; L->sp = 0; }
                }
                if (!any_leaf) break;
                S->cost += P->c_leaf; S->leaf_iters += 1;
            }
            int fin = 0;
            for (int l = 0; l < 32; ++l) fin += W[l].item >= 0 && W[l].node == DONE && W[l].nd == 0;
            if (fin >= need) break;
        }
        for (int l = 0; l < 32; ++l) {
            lane* L = &W[l];
            if (L->item >= 0 && L->node == DONE && L->nd == 0) {
                if (out_id) { out_id[L->item] = L->idbest; out_t[L->item] = L->tbest; }
                S->rays += 1; S->hits += L->idbest >= 0; L->item = -1;
            }
        }
    }
}

static stats run(int n, const policy* P, int warps, int* out_id, double* out_t)
{
    stats S; memset(&S, 0, sizeof S);
    /* every simulated warp owns a contiguous chunk made of 32-ray batches */
    int per = ((n + warps - 1) / warps + 31) / 32 * 32;
    for (int w = 0; w < warps; ++w) { int a = w * per, b = a + per < n ? a + per : n; if (a < b) run_warp(a, b, P, &S, out_id, out_t); }
    return S;
}

static void* slurp(const char* path, size_t* bytes)
{
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); *bytes = ftell(f); fseek(f, 0, SEEK_SET);
    void* p = malloc(*bytes); if (fread(p, 1, *bytes, f) != *bytes) exit(1); fclose(f); return p;
}

static void report(const char* name, const stats* S, const stats* base)
{
    printf("%-44s cost/ray %8.1f (x%.3f)  node iters/ray %6.2f (deep %4.2f) lanes/iter %5.2f nodes/iter %5.2f (per ray %6.2f)  leaf iters/ray %5.2f lanes/iter %5.2f\n", name,
           S->cost / S->rays, base ? S->cost / base->cost : 1.0, S->node_iters / S->rays, S->deep_iters / S->rays, S->node_lane_steps / S->node_iters,
           S->node_lines / S->node_iters, S->node_lines / S->rays,
           S->leaf_iters / S->rays, S->leaf_lane_tests / (S->leaf_iters > 0 ? S->leaf_iters : 1));
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s mesh.bin rays.bin [closest|any] [hits.bin]\n", argv[0]); return 1; }
    size_t nb; int* m = slurp(argv[1], &nb);
    nV = m[0]; nF = m[1]; V = (float*)(m + 2); F = (int*)(V + 3 * nV);
    build_bvh();
    if (getenv("SIM_BVH4")) { use4 = 1; build_bvh4(); }
    if (getenv("SIM_HOT_DEPTH")) hot_depth = atoi(getenv("SIM_HOT_DEPTH"));
    int* r = slurp(argv[2], &nb);
    int n = r[0]; RO = (float*)(r + 2); RD = RO + 3 * (size_t)n;
    int any = argc > 3 && !strcmp(argv[3], "any");
    int warps = n / 32 / 8; if (warps < 1) warps = 1; if (warps > 4736) warps = 4736;   /* >= 8 batches per warp */
    printf("%d verts, %d tris, %d nodes, %d rays, %d simulated warps, %s\n", nV, nF, nN, n, warps, any ? "any-hit" : "closest-hit");
    /* cost weights in issue slots: node step, leaf push, leaf test, refill round (override: SIM_CN / SIM_CP / SIM_CL / SIM_CR) */
    double CN = 45, CP = 6, CL = 90, CR = 60;
    if (getenv("SIM_CN")) CN = atof(getenv("SIM_CN"));
    if (getenv("SIM_CP")) CP = atof(getenv("SIM_CP"));
    if (getenv("SIM_CL")) CL = atof(getenv("SIM_CL"));
    if (getenv("SIM_CR")) CR = atof(getenv("SIM_CR"));
    policy base = {2, 32, 0, any, CN, CP, CL, CR};
    int* hid = malloc(sizeof(int) * n); double* ht = malloc(sizeof(double) * n);
    stats S0 = run(n, &base, warps, hid, ht);
    if (argc > 4) { FILE* f = fopen(argv[4], "wb"); fwrite(hid, sizeof(int), n, f); fwrite(ht, sizeof(double), n, f); fclose(f); }
    if (getenv("SIM_HITS_ONLY")) return 0;
    if (getenv("SIM_BASE_ONLY")) { printf("hit fraction %.3f, node steps per ray %.1f\n", (double)S0.hits / S0.rays, S0.node_lane_steps / S0.rays); report("kernel today: defer 2, refill 32, sync drain", &S0, NULL); return 0; }
    printf("hit fraction %.3f, node steps per ray %.1f, leaf tests per ray %.2f\n", (double)S0.hits / S0.rays, S0.node_lane_steps / S0.rays, S0.leaf_lane_tests / S0.rays);
    report("kernel today: defer 2, refill 32, sync drain", &S0, NULL);
    int defers[] = {1, 4, 8};
    for (int i = 0; i < 3; ++i) { policy p = base; p.defer = defers[i]; stats S = run(n, &p, warps, NULL, NULL); char nm[96]; snprintf(nm, 96, "defer %d, refill 32, sync drain", p.defer); report(nm, &S, &S0); }
    int refills[] = {24, 16, 8};
    for (int i = 0; i < 3; ++i) { policy p = base; p.refill = refills[i]; stats S = run(n, &p, warps, NULL, NULL); char nm[96]; snprintf(nm, 96, "defer 2, refill %d, sync drain", p.refill); report(nm, &S, &S0); }
    int dcap[] = {2, 4, 8}, votes[] = {8, 16, 24}, rf[] = {32, 24, 16, 8};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 4; ++c) {
        policy p = base; p.defer = dcap[a]; p.vote_drain = votes[b]; p.refill = rf[c];
        stats S = run(n, &p, warps, NULL, NULL); char nm[96]; snprintf(nm, 96, "defer %d, refill %d, vote drain at %d", p.defer, p.refill, p.vote_drain); report(nm, &S, &S0);
    }
    return 0;
}
