/* stepsim.c -- CPU count of traversal NODE STEPS (one internal-node fetch + both child box tests, the unit of work of
 * csrc/trace.cuh:node_step) for the three queries of the path, with and without culling stacked nodes at pop time by
 * their stored entry distance.  Test infrastructure: includes the oracle's C source for its LBVH and chain.
 *   gcc -O2 -fopenmp -ffp-contract=off -o stepsim stepsim.c -lm ; ./stepsim rays.bin   (written by run.py)           */
#include <stdio.h>
#include "../../oracle/drt_oracle.c"

typedef struct { long steps, wasted, popcull, tris, rays, maxsp; } cnt_t;

static void trav(const orc_bvh* B, d3 o, d3 d, int any, int popcull, double* t_out, int32_t* id_out, cnt_t* c)
{
    double best = INFINITY; int32_t best_id = -1;
    d3 inv = mk(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
    int32_t stack[128]; double stn[128]; int sp = 0;
    int32_t node = B->root;
    c->rays++;
    if (B->nodes[node].left < 0) { *t_out = best; *id_out = -1; return; }
    for (;;) {
        const orc_node* nd = &B->nodes[node];
        double tl, tr;
        c->steps++;
        int hl = box_test(&B->nodes[nd->left], o, inv, best, &tl);
        int hr = box_test(&B->nodes[nd->right], o, inv, best, &tr);
        int32_t cand[2]; double ct[2]; int nc = 0;
        if (hl && hr) {
            if (tl <= tr) { cand[0] = nd->left; ct[0] = tl; cand[1] = nd->right; ct[1] = tr; }
            else { cand[0] = nd->right; ct[0] = tr; cand[1] = nd->left; ct[1] = tl; }
            nc = 2;
        } else if (hl) { cand[0] = nd->left; ct[0] = tl; nc = 1; }
        else if (hr) { cand[0] = nd->right; ct[0] = tr; nc = 1; }
        else c->wasted++;
        int32_t next = -1;
        for (int k = 0; k < nc; ++k) {
            const orc_node* ch = &B->nodes[cand[k]];
            if (ch->left < 0) {  /* leaf: exact test now */
                double t; int32_t tri = ch->right;
                const int32_t* f = &B->F[3 * (size_t)tri];
                c->tris++;
                if (query_tri(o, d, ld3f(&B->V[3 * (size_t)f[0]]), ld3f(&B->V[3 * (size_t)f[1]]), ld3f(&B->V[3 * (size_t)f[2]]), &t)) {
                    if (t < best || (t == best && tri < best_id)) { best = t; best_id = tri; }
                    if (any) { *t_out = best; *id_out = best_id; return; }
                }
            } else if (next < 0) next = cand[k];
            else { stn[sp] = ct[k]; stack[sp++] = cand[k]; if (sp > c->maxsp) c->maxsp = sp; }
        }
        while (next < 0) {
            if (!sp) { *t_out = best; *id_out = best_id; return; }
            --sp;
            if (popcull && stn[sp] > best) { c->popcull++; continue; }
            next = stack[sp];
        }
        node = next;
    }
}

int main(int argc, char** argv)
{
    FILE* f = fopen(argc > 1 ? argv[1] : "rays.bin", "rb");
    int32_t hdr[3];
    if (!f || fread(hdr, 4, 3, f) != 3) return 1;
    int nV = hdr[0], nF = hdr[1], N = hdr[2];
    float* V32 = malloc(12 * (size_t)nV); double* V64 = malloc(24 * (size_t)nV); int32_t* F = malloc(12 * (size_t)nF);
    double* o = malloc(24 * (size_t)N); double* d = malloc(24 * (size_t)N);
    if (fread(V64, 24, nV, f) != (size_t)nV || fread(F, 12, nF, f) != (size_t)nF || fread(o, 24, N, f) != (size_t)N || fread(d, 24, N, f) != (size_t)N) return 2;
    for (int i = 0; i < 3 * nV; ++i) V32[i] = (float)V64[i];
    orc_bvh* B = orc_bvh_build(V32, nV, F, nF);
    for (int pc = 0; pc < 2; ++pc) {
        cnt_t c[3] = {{0}};
        for (int i = 0; i < N; ++i) {
            d3 oo = ld3(o + 3 * i), dd = ld3(d + 3 * i);
            d3 of = mk((float)oo.x, (float)oo.y, (float)oo.z), df = mk((float)dd.x, (float)dd.y, (float)dd.z);
            double t; int32_t id1, id2, id3;
            trav(B, of, df, 0, pc, &t, &id1, &c[0]);
            if (id1 < 0) continue;
            hit_rec h; d3 o1, d1, o2, d2;
            const int32_t* ff = &F[3 * (size_t)id1];
            hit_forward(&h, oo, dd, ld3(&V64[3 * (size_t)ff[0]]), ld3(&V64[3 * (size_t)ff[1]]), ld3(&V64[3 * (size_t)ff[2]]), 1.00029, 1.4723, &o1, &d1);
            if (h.tir) continue;
            of = mk((float)o1.x, (float)o1.y, (float)o1.z); df = mk((float)d1.x, (float)d1.y, (float)d1.z);
            trav(B, of, df, 0, pc, &t, &id2, &c[1]);
            if (id2 < 0) continue;
            ff = &F[3 * (size_t)id2];
            hit_forward(&h, o1, d1, ld3(&V64[3 * (size_t)ff[0]]), ld3(&V64[3 * (size_t)ff[1]]), ld3(&V64[3 * (size_t)ff[2]]), 1.00029, 1.4723, &o2, &d2);
            if (h.tir) continue;
            of = mk((float)o2.x, (float)o2.y, (float)o2.z); df = mk((float)d2.x, (float)d2.y, (float)d2.z);
            trav(B, of, df, 1, pc, &t, &id3, &c[2]);
        }
        for (int q = 0; q < 3; ++q)
            printf("popcull=%d Q%d: rays %ld  steps/ray %.2f  wasted(no child hit)/ray %.2f  pops culled/ray %.2f  tris/ray %.2f  max stack %ld\n", pc, q + 1,
                   c[q].rays, (double)c[q].steps / c[q].rays, (double)c[q].wasted / c[q].rays, (double)c[q].popcull / c[q].rays,
                   (double)c[q].tris / c[q].rays, c[q].maxsp);
    }
    return 0;
}
