/* TEST INFRASTRUCTURE ONLY -- float64 CPU restatement of the reference's collision-check hot path.
 *
 * Nothing under motionplanner_b200/ may link or load this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do (as the checker and as the reported CPU baseline).
 *
 * What it restates (reference file:line, all relative to /root/reference):
 *   - collision_check():   src/lowLevelPlannerManeuverAutomaton.py:95-125 and the twin
 *                          src/maneuverAutomatonPlannerStandalone.py:196-226
 *                          (per occupancy: time-index guards, affine_transform, space[t].contains(pgon),
 *                           early return on the first violation)
 *   - free_space():        src/maneuverAutomatonPlannerStandalone.py:151-182  (space_t = road - U obstacles_t);
 *                          evaluated in its decomposed form: contains(P) <=> no boundary segment of road_t meets
 *                          int(P), an interior point of P is inside road_t, and no obstacle interior meets int(P)
 *   - collisionChecker():  auxiliary/collisionChecker.py:6-35 (road.contains(car) and not obstacle.intersects(car))
 *                          -> same loop with mode bit 0 set (closed-set "intersects" for the obstacles)
 *   - shapely affine_transform arithmetic: xp = a*x + b*y + xoff, left to right, float64
 *   - GEOS robust predicates: orientation signs are decided exactly (float64 filter + exact expansion arithmetic),
 *     so every decision equals the decision exact rational arithmetic makes on the same float64 vertices
 *     (pinned against oracle/exact.py in tests/test_oracle_c.py).
 *
 * PARITY UNPINNED against shapely/GEOS itself (not installable here; the reference has no golden vectors for this
 * path) -- see oracle/exact.py header and DESIGN.md.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_VMAX 32

/* ------------------------------------------------------------------------------------------------------------------ */
/* exact orientation sign                                                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */

static inline void two_sum(double a, double b, double *s, double *e) {
    double x = a + b;
    double bv = x - a;
    double av = x - bv;
    *e = (a - av) + (b - bv);
    *s = x;
}

static inline void two_prod(double a, double b, double *p, double *e) {
    double x = a * b;
    *e = fma(a, b, -x);
    *p = x;
}

/* add the double b to the non-overlapping expansion e[0..n) (components in increasing magnitude) */
static int grow_expansion(int n, const double *e, double b, double *h) {
    double q = b;
    for (int i = 0; i < n; i++) {
        double s, r;
        two_sum(q, e[i], &s, &r);
        h[i] = r;
        q = s;
    }
    h[n] = q;
    return n + 1;
}

static int orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    /* (ax-cx)(by-cy) - (ay-cy)(bx-cx) = ax*by - ax*cy - cx*by - ay*bx + ay*cx + cy*bx   (cx*cy cancels) */
    double t[12];
    two_prod(ax, by, &t[0], &t[1]);
    two_prod(-ax, cy, &t[2], &t[3]);
    two_prod(-cx, by, &t[4], &t[5]);
    two_prod(-ay, bx, &t[6], &t[7]);
    two_prod(ay, cx, &t[8], &t[9]);
    two_prod(cy, bx, &t[10], &t[11]);
    double e[16], h[16];
    int n = 0;
    for (int i = 0; i < 12; i++) {
        n = grow_expansion(n, e, t[i], h);
        memcpy(e, h, sizeof(double) * (size_t)n);
    }
    for (int i = n - 1; i >= 0; i--) {
        if (e[i] > 0.0) return 1;
        if (e[i] < 0.0) return -1;
    }
    return 0;
}

static long long g_exact_calls = 0;

/* sign of |b-a, c-a| (c left of a->b: +1), exact */
int orc_orient2d(double ax, double ay, double bx, double by, double cx, double cy) {
    double detleft = (ax - cx) * (by - cy);
    double detright = (ay - cy) * (bx - cx);
    double det = detleft - detright;
    double errbound = 3.3306690738754716e-16 * (fabs(detleft) + fabs(detright));
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
#pragma omp atomic
    g_exact_calls++;
    return orient_exact(ax, ay, bx, by, cx, cy);
}

long long orc_exact_calls(void) { return g_exact_calls; }

/* ------------------------------------------------------------------------------------------------------------------ */
/* polygon helpers                                                                                                     */
/* ------------------------------------------------------------------------------------------------------------------ */

typedef struct {
    int n;
    double x[ORC_VMAX], y[ORC_VMAX];
} poly_t;

static double shoelace2(const poly_t *p) {
    double s = 0.0;
    for (int i = 0; i < p->n; i++) {
        int k = (i + 1 == p->n) ? 0 : i + 1;
        s += (p->x[i] - p->x[0]) * (p->y[k] - p->y[0]) - (p->x[k] - p->x[0]) * (p->y[i] - p->y[0]);
    }
    return s;
}

static void make_ccw(poly_t *p) {
    if (shoelace2(p) < 0.0) {
        for (int i = 0, k = p->n - 1; i < k; i++, k--) {
            double tx = p->x[i], ty = p->y[i];
            p->x[i] = p->x[k]; p->y[i] = p->y[k];
            p->x[k] = tx; p->y[k] = ty;
        }
    }
}

/* an edge of CCW convex A has all of B on its outer side (strict: strictly outside) */
static int separated_by_edge_of(const poly_t *A, const poly_t *B, int strict) {
    for (int i = 0; i < A->n; i++) {
        int k = (i + 1 == A->n) ? 0 : i + 1;
        if (A->x[i] == A->x[k] && A->y[i] == A->y[k]) continue;
        int all_out = 1;
        for (int q = 0; q < B->n; q++) {
            int o = orc_orient2d(A->x[i], A->y[i], A->x[k], A->y[k], B->x[q], B->y[q]);
            if (strict ? (o >= 0) : (o > 0)) { all_out = 0; break; }
        }
        if (all_out) return 1;
    }
    return 0;
}

/* approximate signed separation distance (metres): > 0 separated, only used for the near-tangent statistic */
static double approx_sep(const poly_t *A, const poly_t *B) {
    double best = -INFINITY;
    for (int pass = 0; pass < 2; pass++) {
        const poly_t *P = pass ? B : A, *Q = pass ? A : B;
        for (int i = 0; i < P->n; i++) {
            int k = (i + 1 == P->n) ? 0 : i + 1;
            double ex = P->x[k] - P->x[i], ey = P->y[k] - P->y[i];
            double len = sqrt(ex * ex + ey * ey);
            if (len == 0.0) continue;
            double mn = INFINITY;
            for (int q = 0; q < Q->n; q++) {
                /* outward normal of a CCW edge is (ey, -ex) */
                double d = (ey * (Q->x[q] - P->x[i]) - ex * (Q->y[q] - P->y[i])) / len;
                if (d < mn) mn = d;
            }
            if (mn > best) best = mn;
        }
    }
    return best;
}

/* mode strict: closed polygons share a point; else: open interiors share a point */
static int convex_collide(const poly_t *A, const poly_t *B, int strict) {
    if (separated_by_edge_of(A, B, strict)) return 0;
    if (separated_by_edge_of(B, A, strict)) return 0;
    return 1;
}

int orc_convex_collide(const double *a, int na, const double *b, int nb, int strict) {
    poly_t A, B;
    A.n = na; B.n = nb;
    for (int i = 0; i < na; i++) { A.x[i] = a[2 * i]; A.y[i] = a[2 * i + 1]; }
    for (int i = 0; i < nb; i++) { B.x[i] = b[2 * i]; B.y[i] = b[2 * i + 1]; }
    make_ccw(&A); make_ccw(&B);
    return convex_collide(&A, &B, strict);
}

/* closed segment s0-s1 has a point in the open interior of CCW convex P */
static int segment_meets_interior(const poly_t *P, double s0x, double s0y, double s1x, double s1y) {
    for (int i = 0; i < P->n; i++) {
        int k = (i + 1 == P->n) ? 0 : i + 1;
        if (P->x[i] == P->x[k] && P->y[i] == P->y[k]) continue;
        if (orc_orient2d(P->x[i], P->y[i], P->x[k], P->y[k], s0x, s0y) <= 0 &&
            orc_orient2d(P->x[i], P->y[i], P->x[k], P->y[k], s1x, s1y) <= 0)
            return 0;
    }
    if (s0x == s1x && s0y == s1y) return 1;
    int pos = 0, neg = 0;
    for (int q = 0; q < P->n; q++) {
        int o = orc_orient2d(s0x, s0y, s1x, s1y, P->x[q], P->y[q]);
        if (o > 0) pos = 1;
        if (o < 0) neg = 1;
    }
    return pos && neg;
}

int orc_segment_meets_interior(const double *p, int np, const double *seg) {
    poly_t P;
    P.n = np;
    for (int i = 0; i < np; i++) { P.x[i] = p[2 * i]; P.y[i] = p[2 * i + 1]; }
    make_ccw(&P);
    return segment_meets_interior(&P, seg[0], seg[1], seg[2], seg[3]);
}

/* even-odd crossing parity of the +x ray from (px,py) over the segment soup; returns 1 inside, 0 outside, 2 on
 * a boundary segment */
static int point_in_soup(double px, double py, const double *segs, int nseg) {
    int inside = 0;
    for (int i = 0; i < nseg; i++) {
        double ax = segs[4 * i], ay = segs[4 * i + 1], bx = segs[4 * i + 2], by = segs[4 * i + 3];
        int o = orc_orient2d(ax, ay, bx, by, px, py);
        if (o == 0 && fmin(ax, bx) <= px && px <= fmax(ax, bx) && fmin(ay, by) <= py && py <= fmax(ay, by))
            return 2;
        if ((ay > py) != (by > py)) {
            if ((by > ay && o > 0) || (by < ay && o < 0)) inside = !inside;
        }
    }
    return inside;
}

int orc_point_in_soup(double px, double py, const double *segs, int nseg) { return point_in_soup(px, py, segs, nseg); }

/* ------------------------------------------------------------------------------------------------------------------ */
/* batched hot path -- argument layout identical to include/mpc.h so the parity tests feed both the same arrays        */
/* ------------------------------------------------------------------------------------------------------------------ */

typedef struct {
    double x, y, c, s;
    int32_t scene, t0, step_limit, cand_set, cand_off, cand_cnt, reserved0, reserved1;
} orc_query_t;

typedef struct {
    long long pairs_decided;      /* occupancy x (obstacle | segment) pairs actually evaluated (early return honoured) */
    long long pairs_nominal;      /* candidates x occupancies-in-range x (obstacles + segments) */
    long long near_tangent;       /* evaluated pairs with |separation| < 1e-9 m */
    long long exact_calls;        /* orientation tests that needed exact arithmetic */
} orc_stats_t;

typedef struct {
    /* library */
    const double *lib_verts;   /* [P][J][Vl][2] */
    const int32_t *lib_nv;     /* [P][J]  0 = unused */
    const int32_t *lib_tidx;   /* [P][J] */
    int P, J, Vl;
    const int32_t *set_off;    /* [nsets+1] candidate sets (velocity buckets) */
    const int32_t *set_idx;
    int nsets;
    /* scenes */
    int nscenes;
    const int32_t *scene_step_off; /* [nscenes+1] */
    const int32_t *step_obst_off;  /* [nsteps+1] */
    const double *obst;            /* [nobst][Vo][2] */
    const int32_t *obst_nv;        /* [nobst] */
    int Vo;
    const int32_t *step_seg_begin; /* [nsteps]  -1 = no region constraint at this step */
    const int32_t *step_seg_end;   /* [nsteps] */
    const double *segs;            /* [nseg][4] */
} orc_problem_t;

#define MODE_STRICT 1      /* obstacles: closed-set intersects (validator) instead of interior overlap (planner) */
#define MODE_NO_EARLY_EXIT 4

static int check_candidate(const orc_problem_t *pb, const orc_query_t *q, int p, int mode, orc_stats_t *st) {
    int collide = 0;
    int s0 = pb->scene_step_off[q->scene];
    int nsteps = pb->scene_step_off[q->scene + 1] - s0;
    /* nominal work of this candidate, independent of the early return */
    for (int j = 0; j < pb->J; j++) {
        if (pb->lib_nv[(size_t)p * pb->J + j] <= 0) continue;
        int t = q->t0 + pb->lib_tidx[(size_t)p * pb->J + j];
        if (t > q->step_limit || t < 0 || t >= nsteps) continue;
        int step = s0 + t;
        st->pairs_nominal += pb->step_obst_off[step + 1] - pb->step_obst_off[step];
        if (pb->step_seg_begin[step] >= 0) st->pairs_nominal += pb->step_seg_end[step] - pb->step_seg_begin[step];
    }
    for (int j = 0; j < pb->J; j++) {
        int nv = pb->lib_nv[(size_t)p * pb->J + j];
        if (nv <= 0) continue;
        int t = q->t0 + pb->lib_tidx[(size_t)p * pb->J + j];
        if (t > q->step_limit) continue;           /* ind + time <= param['steps']   (reference :120) */
        if (t < 0 || t >= nsteps) continue;        /* ind + time < len(space)        (reference :122) */
        int step = s0 + t;
        /* affine_transform(o['space'], [cos, -sin, sin, cos, x, y])   (reference :121) */
        poly_t A;
        A.n = nv;
        const double *lv = pb->lib_verts + ((size_t)p * pb->J + j) * pb->Vl * 2;
        double a = q->c, b = -q->s, d = q->s, e = q->c;
        for (int k = 0; k < nv; k++) {
            A.x[k] = a * lv[2 * k] + b * lv[2 * k + 1] + q->x;
            A.y[k] = d * lv[2 * k] + e * lv[2 * k + 1] + q->y;
        }
        make_ccw(&A);
        int o0 = pb->step_obst_off[step], o1 = pb->step_obst_off[step + 1];
        int g0 = pb->step_seg_begin[step], g1 = pb->step_seg_end[step];
        int nseg = (g0 >= 0) ? (g1 - g0) : 0;
        int bad = 0;
        /* road / free-space boundary */
        if (g0 >= 0) {
            for (int g = g0; g < g1 && (!bad || (mode & MODE_NO_EARLY_EXIT)); g++) {
                const double *sg = pb->segs + 4 * (size_t)g;
                st->pairs_decided++;
                if (segment_meets_interior(&A, sg[0], sg[1], sg[2], sg[3])) bad = 1;
            }
            if (!bad) {
                double cx = 0.0, cy = 0.0;
                for (int k = 0; k < nv; k++) { cx += A.x[k]; cy += A.y[k]; }
                cx /= nv; cy /= nv;
                if (point_in_soup(cx, cy, pb->segs + 4 * (size_t)g0, nseg) != 1) bad = 1;
            }
        }
        /* obstacles */
        double ax0 = A.x[0], ax1 = A.x[0], ay0 = A.y[0], ay1 = A.y[0];
        for (int k = 1; k < nv; k++) {
            ax0 = fmin(ax0, A.x[k]); ax1 = fmax(ax1, A.x[k]);
            ay0 = fmin(ay0, A.y[k]); ay1 = fmax(ay1, A.y[k]);
        }
        for (int o = o0; o < o1 && (!bad || (mode & MODE_NO_EARLY_EXIT)); o++) {
            int nb = pb->obst_nv[o];
            if (nb <= 0) continue;
            poly_t B;
            B.n = nb;
            const double *ov = pb->obst + (size_t)o * pb->Vo * 2;
            for (int k = 0; k < nb; k++) { B.x[k] = ov[2 * k]; B.y[k] = ov[2 * k + 1]; }
            make_ccw(&B);
            st->pairs_decided++;
            /* envelope rejection first (GEOS short-circuits on disjoint envelopes before relate); the comparison
             * is exact: a float64 difference > 1e-6 proves the envelopes are strictly disjoint */
            double bx0 = B.x[0], bx1 = B.x[0], by0 = B.y[0], by1 = B.y[0];
            for (int k = 1; k < nb; k++) {
                bx0 = fmin(bx0, B.x[k]); bx1 = fmax(bx1, B.x[k]);
                by0 = fmin(by0, B.y[k]); by1 = fmax(by1, B.y[k]);
            }
            double gap = fmax(fmax(bx0 - ax1, ax0 - bx1), fmax(by0 - ay1, ay0 - by1));
            if (gap > 1e-6) continue;
            if (convex_collide(&A, &B, mode & MODE_STRICT)) bad = 1;
            if (fabs(approx_sep(&A, &B)) < 1e-9) st->near_tangent++;
        }
        if (bad) {
            collide = 1;
            if (!(mode & MODE_NO_EARLY_EXIT)) break;   /* reference returns False at the first violation */
        }
    }
    return collide;
}

/* out[i] = 1 if candidate i (flat order over queries) collides.  Returns 0. */
int orc_check_batch(const orc_problem_t *pb, const orc_query_t *queries, int B, const int32_t *cand_idx, int mode,
                    uint8_t *out, orc_stats_t *stats, int nthreads) {
    /* flat candidate offsets */
    long long *off = (long long *)malloc(sizeof(long long) * (size_t)(B + 1));
    off[0] = 0;
    for (int q = 0; q < B; q++) off[q + 1] = off[q] + queries[q].cand_cnt;
    orc_stats_t total;
    memset(&total, 0, sizeof(total));
    long long ex0 = g_exact_calls;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        orc_stats_t st;
        memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic, 1)
        for (int q = 0; q < B; q++) {
            const orc_query_t *qq = &queries[q];
            const int32_t *list = (qq->cand_set >= 0) ? pb->set_idx + pb->set_off[qq->cand_set] + qq->cand_off
                                                      : cand_idx + qq->cand_off;
            for (int i = 0; i < qq->cand_cnt; i++) out[off[q] + i] = (uint8_t)check_candidate(pb, qq, list[i], mode, &st);
        }
#pragma omp critical
        {
            total.pairs_decided += st.pairs_decided;
            total.pairs_nominal += st.pairs_nominal;
            total.near_tangent += st.near_tangent;
        }
    }
    total.exact_calls = g_exact_calls - ex0;
    if (stats) *stats = total;
    free(off);
    return 0;
}

/* same as orc_check_batch but parallel over the candidates of each query (for single huge queries) */
int orc_check_batch_inner(const orc_problem_t *pb, const orc_query_t *queries, int B, const int32_t *cand_idx,
                          int mode, uint8_t *out, orc_stats_t *stats, int nthreads) {
    orc_stats_t total;
    memset(&total, 0, sizeof(total));
    long long ex0 = g_exact_calls;
    long long base = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int q = 0; q < B; q++) {
        const orc_query_t *qq = &queries[q];
        const int32_t *list = (qq->cand_set >= 0) ? pb->set_idx + pb->set_off[qq->cand_set] + qq->cand_off
                                                  : cand_idx + qq->cand_off;
#pragma omp parallel
        {
            orc_stats_t st;
            memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic, 16)
            for (int i = 0; i < qq->cand_cnt; i++) out[base + i] = (uint8_t)check_candidate(pb, qq, list[i], mode, &st);
#pragma omp critical
            {
                total.pairs_decided += st.pairs_decided;
                total.pairs_nominal += st.pairs_nominal;
                total.near_tangent += st.near_tangent;
            }
        }
        base += qq->cand_cnt;
    }
    total.exact_calls = g_exact_calls - ex0;
    if (stats) *stats = total;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* primitive-selection step: cost and goal test of every child of an expansion                       */
/* ------------------------------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t time_start, time_end; /* goal['time_start'] <= i <= goal['time_end']   (reference :136)                     */
    int32_t seg_begin, seg_end;   /* boundary segments of goal['space'] in goal_segs; seg_begin < 0: space is None      */
} orc_goal_t;

/* numpy's pairwise summation of a contiguous float64 run (numpy/_core/src/umath/loops_utils.h.src, n <= 128), the
 * routine np.sum runs over `(ref_traj[0:2, index] - x[0:2, index])**2` (reference :210): eight running sums, combined
 * as a tree, then the tail in order */
static double numpy_pairwise_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

#define ORC_MAX_STATES 64

/* For every candidate (flat order over the queries): the cost expand_node gives the child (reference :199-210, without
 * the +1000 fixed-prefix penalty of :212, which is search bookkeeping) and the index goal_check returns for it
 * (reference :127-142), -1 when no goal set is reached.
 *   states [P][L][4]  primitive.x transposed (x, y, v, phi per step, rear-axle frame), nstates[P] = primitive.x.shape[1]
 *   ref_xy [2][M]     ref_traj[0:2]
 *   parent [B][2]     node.cost, node.x[3, -1]; the query's (x, y, c, s) are node.x[0:2, -1], cos, sin; t0 = ind
 * The elements are summed in the order numpy visits them: the fancy-indexed operands are Fortran-ordered, i.e.
 * x0, y0, x1, y1, ... */
int orc_expand(const double *states, const int32_t *nstates, int P, int L, const double *ref_xy, int M,
               const orc_goal_t *goals, int ngoals, const double *goal_segs, double b, const int32_t *set_off,
               const int32_t *set_idx, const orc_query_t *queries, int B, const int32_t *cand_idx, const double *parent,
               double *out_cost, int32_t *out_goal) {
    long long base = 0;
    if (L > ORC_MAX_STATES) return -1;
    for (int q = 0; q < B; q++) {
        const orc_query_t *qq = &queries[q];
        const int32_t *list = (qq->cand_set >= 0) ? set_idx + set_off[qq->cand_set] + qq->cand_off : cand_idx + qq->cand_off;
        const double pcost = parent[2 * q], pphi = parent[2 * q + 1];
        const int ind = qq->t0;
        for (int i = 0; i < qq->cand_cnt; i++) {
            const int p = list[i];
            if (p < 0 || p >= P) return -2;
            const int n = nstates[p];
            const double *st = states + (size_t)p * L * 4;
            double X[ORC_MAX_STATES], Y[ORC_MAX_STATES], a[2 * ORC_MAX_STATES];
            /* x_ = T @ primitive.x + [x, y, 0, phi]   (reference :199).  numpy hands the 4x4 @ 4xL product to its BLAS,
             * whose kernels accumulate with fused multiply-adds along k: c*x first, then fma(-s, y, .); the zero
             * entries of T add nothing.  (Checked against numpy in tests/test_expand_cpu.py; the float tolerance for
             * a BLAS that rounds differently is 1e-12 relative, stated there.) */
            for (int k = 0; k < n; k++) {
                X[k] = fma(-qq->s, st[4 * k + 1], qq->c * st[4 * k]) + qq->x;
                Y[k] = fma(qq->c, st[4 * k + 1], qq->s * st[4 * k]) + qq->y;
            }
            int K = n < M - ind ? n : M - ind;     /* index = range(ind, min(x.shape[1], ref_traj.shape[1])) (:204-207) */
            if (K < 0) K = 0;
            for (int k = 0; k < K; k++) {
                double dx = ref_xy[ind + k] - X[k], dy = ref_xy[(size_t)M + ind + k] - Y[k];
                a[2 * k] = dx * dx;
                a[2 * k + 1] = dy * dy;
            }
            out_cost[base + i] = pcost + numpy_pairwise_sum(a, 2 * K);
            int hit = -1;
            for (int k = 0; k < n && hit < 0; k++) {
                const int step = ind + k;
                for (int g = 0; g < ngoals; g++) {
                    if (!(goals[g].time_start <= step && step <= goals[g].time_end)) continue;
                    const double th = st[4 * k + 3] + pphi;
                    const double px = X[k] + cos(th) * b, py = Y[k] + sin(th) * b;
                    if (goals[g].seg_begin < 0 ||
                        point_in_soup(px, py, goal_segs + 4 * (size_t)goals[g].seg_begin, goals[g].seg_end - goals[g].seg_begin) == 1) {
                        hit = step;
                        break;
                    }
                }
            }
            out_goal[base + i] = hit;
        }
        base += qq->cand_cnt;
    }
    return 0;
}
