// mpc_expand.cuh -- the rest of the primitive-selection step: cost and goal test of every child of an expansion
// (SURVEY.md 8f N1).  The reference builds each child with expand_node (src/lowLevelPlannerManeuverAutomaton.py:192-214:
// one 4x4 @ 4x11 product, a concatenate and an np.sum per child) and tests the goal with shapely when the child is popped
// (goal_check, :127-142).  Here one thread per child does both right next to the collision bits, in float64, in the
// reference's operation order, so that the best-first search orders its queue exactly as the reference does:
//   * x_ = T @ primitive.x + offset: the BLAS behind numpy accumulates along k with fused multiply-adds
//     (c*x, then fma(-s, y, .)), the zero entries of T add nothing;
//   * np.sum over the Fortran-ordered operands visits x0, y0, x1, y1, ... with numpy's pairwise scheme (eight running
//     sums, a fixed tree, the tail in order);
//   * goal['space'].contains(Point) is the even-odd parity of the exact orientation signs; a point on the boundary is
//     not contained.
// The file is compiled with --fmad=false: nothing below is contracted behind our back.
#pragma once
#include <cstdint>

#include "mpc_exact.cuh"

namespace mpc {

constexpr int EXPAND_MAX_STATES = 64;     // columns of primitive.x (11 in the bundled automaton, 11 in AROC libraries)

struct EQuery {          // 64 B, one per expanded parent
    double x, y, c, s;   // node.x[0:2, -1], cos / sin of node.x[3, -1] as the host computed them
    double cost, phi;    // node.cost, node.x[3, -1]
    int t0;              // ind: index of the parent's last state
    int cand_base;       // first candidate in the list named by cand_src
    int cand_cnt;
    int cand_src;        // 0: staged candidate sets (caller's order), 1: explicit list of this call
};

struct EGoal {
    int time_start, time_end, seg_begin, seg_end;      // seg_begin < 0: goal['space'] is None
};

struct EParams {
    const EQuery *q;
    const int *out_off;              // [B+1] first child of every parent in the flat output
    int B, total;
    const int *sets, *explicit_cand;
    const double *states;            // [P][L][4]  primitive.x transposed: x, y, v, phi per step
    const int *nstates;              // [P]
    int L;
    const double *ref_xy;            // [2][M]
    int M;
    const EGoal *goals;
    int ngoals;
    const double *goal_segs;         // [n][4]
    double b;                        // param['b']: rear axle -> vehicle centre
    double *cost;
    int *goal_step;
};

// 1 inside, 0 outside, 2 on a boundary segment: crossings of the +x ray, every sign exact
__device__ __forceinline__ int point_in_soup_exact(double px, double py, const double *segs, int nseg) {
    int inside = 0;
    unsigned ties = 0;
    for (int i = 0; i < nseg; i++) {
        const double ax = segs[4 * i], ay = segs[4 * i + 1], bx = segs[4 * i + 2], by = segs[4 * i + 3];
        const int o = orient2d(ax, ay, bx, by, px, py, ties);
        if (o == 0 && fmin(ax, bx) <= px && px <= fmax(ax, bx) && fmin(ay, by) <= py && py <= fmax(ay, by)) return 2;
        if ((ay > py) != (by > py))
            if ((by > ay && o > 0) || (by < ay && o < 0)) inside ^= 1;
    }
    return inside;
}

__global__ void __launch_bounds__(128) k_expand(const EParams e) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= e.total) return;
    int lo = 0, hi = e.B;                       // parent of child t: largest q with out_off[q] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (e.out_off[mid] <= t) lo = mid; else hi = mid;
    }
    const EQuery Q = e.q[lo];
    const int i = t - e.out_off[lo];
    const int p = (Q.cand_src ? e.explicit_cand : e.sets)[Q.cand_base + i];
    const int n = min(e.nstates[p], EXPAND_MAX_STATES);
    const double *st = e.states + (size_t)p * e.L * 4;
    const int ind = Q.t0;
    int K = min(n, e.M - ind);
    if (K < 0) K = 0;
    const int nel = 2 * K;
    // numpy's pairwise sum, fed element by element in visiting order
    double r[8], tail = 0.0, small = 0.0;
    const int body = nel - (nel % 8);
    int hit = -1;
    for (int k = 0; k < n; k++) {
        const double sx = st[4 * k], sy = st[4 * k + 1];
        const double X = __dadd_rn(__fma_rn(-Q.s, sy, __dmul_rn(Q.c, sx)), Q.x);
        const double Y = __dadd_rn(__fma_rn(Q.c, sy, __dmul_rn(Q.s, sx)), Q.y);
        if (k < K) {
            const double dx = __dsub_rn(e.ref_xy[ind + k], X), dy = __dsub_rn(e.ref_xy[(size_t)e.M + ind + k], Y);
            const double a0 = __dmul_rn(dx, dx), a1 = __dmul_rn(dy, dy);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int m = 2 * k + h;
                const double a = h ? a1 : a0;
                if (nel < 8) small = __dadd_rn(small, a);
                else if (m < 8) r[m] = a;
                else if (m < body) r[m & 7] = __dadd_rn(r[m & 7], a);
                else {
                    if (m == body)
                        tail = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                    tail = __dadd_rn(tail, a);
                }
            }
        }
        if (hit < 0) {
            const int step = ind + k;
            for (int g = 0; g < e.ngoals; g++) {
                const EGoal G = e.goals[g];
                if (!(G.time_start <= step && step <= G.time_end)) continue;
                bool in = G.seg_begin < 0;
                if (!in) {
                    const double th = __dadd_rn(st[4 * k + 3], Q.phi);
                    const double px = __dadd_rn(X, __dmul_rn(cos(th), e.b)), py = __dadd_rn(Y, __dmul_rn(sin(th), e.b));
                    in = point_in_soup_exact(px, py, e.goal_segs + 4 * (size_t)G.seg_begin, G.seg_end - G.seg_begin) == 1;
                }
                if (in) { hit = step; break; }
            }
        }
    }
    double sum;
    if (nel < 8) sum = small;
    else if (nel == body)       // no tail: the tree was never folded inside the loop
        sum = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    else sum = tail;
    e.cost[t] = __dadd_rn(Q.cost, sum);
    e.goal_step[t] = hit;
}

}  // namespace mpc
