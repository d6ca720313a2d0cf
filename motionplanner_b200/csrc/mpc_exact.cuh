// mpc_exact.cuh -- float64 tie-breaking pass: exact orientation signs on the device.
//
// The float32 separating-axis kernel (mpc_kernels.cu) decides a pair only when the separation is outside a
// conservative error band; the pairs inside the band are re-decided by k_refine with the predicate below.  Everything below is evaluated on the SAME float64
// vertices the reference hands to GEOS (shapely affine_transform arithmetic, src/lowLevelPlannerManeuverAutomaton.py:121),
// and every orientation sign is exact: a forward-error filter first, then non-overlapping expansion arithmetic
// (two-sum / FMA two-product) when the filter cannot decide.  So the final decision equals the decision exact rational
// arithmetic makes (reference behaviour: GEOS robust predicates behind shapely contains / intersects,
// auxiliary/collisionChecker.py:21,32).
//
// The file is compiled with --fmad=false: a*b+c below really is two roundings, like Python float arithmetic.
#pragma once
#include <cstdint>

namespace mpc {

__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    double bv = __dsub_rn(s, a);
    double av = __dsub_rn(s, bv);
    e = __dadd_rn(__dsub_rn(a, av), __dsub_rn(b, bv));
}

__device__ __forceinline__ void two_prod(double a, double b, double &p, double &e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}

// sign of ax*by - ax*cy - cx*by - ay*bx + ay*cx + cy*bx, all six products and the sum carried exactly
__device__ __noinline__ int orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    double t[12];
    two_prod(ax, by, t[0], t[1]);
    two_prod(-ax, cy, t[2], t[3]);
    two_prod(-cx, by, t[4], t[5]);
    two_prod(-ay, bx, t[6], t[7]);
    two_prod(ay, cx, t[8], t[9]);
    two_prod(cy, bx, t[10], t[11]);
    double e[13];
    int n = 0;
    for (int i = 0; i < 12; i++) {          // grow-expansion: e stays non-overlapping, increasing magnitude
        double q = t[i];
        for (int k = 0; k < n; k++) {
            double s, r;
            two_sum(q, e[k], s, r);
            e[k] = r;
            q = s;
        }
        e[n++] = q;
    }
    for (int i = n - 1; i >= 0; i--) {
        if (e[i] > 0.0) return 1;
        if (e[i] < 0.0) return -1;
    }
    return 0;
}

// exact sign of |b-a, c-a| (+1: c left of a->b); *ties counts the calls that needed expansion arithmetic
__device__ __forceinline__ int orient2d(double ax, double ay, double bx, double by, double cx, double cy,
                                        unsigned &ties) {
    double detleft = __dmul_rn(__dsub_rn(ax, cx), __dsub_rn(by, cy));
    double detright = __dmul_rn(__dsub_rn(ay, cy), __dsub_rn(bx, cx));
    double det = __dsub_rn(detleft, detright);
    double errbound = __dmul_rn(3.3306690738754716e-16, __dadd_rn(fabs(detleft), fabs(detright)));
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    ties++;
    return orient_exact(ax, ay, bx, by, cx, cy);
}

}  // namespace mpc
