// mpc_kernels.cu -- libmpcheck.so: batched motion-primitive collision check for one B200 (sm_100a).
//
// Hot path replaced (reference, /root/reference):
//   collision_check()   src/lowLevelPlannerManeuverAutomaton.py:95-125, src/maneuverAutomatonPlannerStandalone.py:196-226
//   free_space()        src/maneuverAutomatonPlannerStandalone.py:151-182   (decomposed: road boundary + obstacles)
//   collisionChecker()  auxiliary/collisionChecker.py:6-35
// The reference evaluates one shapely contains()/intersects() per occupancy polygon in a Python loop; here a whole
// batch of (parent pose, candidate primitive) pairs is decided in one launch.  See DESIGN.md for the data layout,
// the error-band argument and the roofline; include/mpc.h for the ABI.
//
// Kernel structure (B200-first, nothing here is derived from reference code -- it has no kernel):
//   k_stage_obst / k_stage_segs   float64 global-frame inputs -> scene-local float32 SoA records (vertices, unit edge
//                                 normals, bounding circle), CCW-normalised float64 copy for the tie-break pass.
//   k_check<VA,VB>                one CTA per (query, occupancy slot, chunk of candidate groups).  The obstacle and
//                                 boundary-segment records of that time step are brought into shared memory with one
//                                 TMA bulk copy per field (cp.async.bulk + mbarrier).  One warp works on 32 candidate
//                                 primitives at a time, lane = primitive: the lane holds its transformed occupancy
//                                 polygon (vertices + edge lines) in registers and streams the step's obstacles from
//                                 shared memory as warp-wide broadcasts.  Axis 0 is the centre line with bounding
//                                 radii; pairs it cannot separate run the full separating-axis test.  A tile ends as
//                                 soon as __all_sync says every lane has collided; __ballot_sync assembles the 32
//                                 result bits that are OR-ed into the packed mask.
//   k_refine                      float64 / exact re-decision of the pairs inside the float32 error band.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpc.h"
#include "mpc_exact.cuh"

namespace mpc {

constexpr float BIG = 1e30f;
constexpr float FAR = 1e18f;
constexpr int GPC_MAX = 32;       // candidate groups (of 32) per CTA
constexpr unsigned FULL = 0xffffffffu;

enum { KIND_OBST = 0, KIND_SEG = 1, KIND_PARITY = 2 };

struct QDev {
    double x, y, c, s;            // float64 pose, global frame (tie-break pass)
    float lx, ly, lc, ls;         // float32 pose relative to the scene origin
    int step0, nsteps, t0, step_limit;
    int cand_base, cand_cnt, mask_off, chunks;
};

struct Counters {
    unsigned long long pairs_nominal, refined, exact_ties, near_tangent, parity_refined, cull_tests, axis_tests;
    unsigned wl_count, pad;
};

struct KParams {
    // library
    const float4 *lib_v, *lib_n, *lib_c;
    const double *lib_v64;
    const int *lib_nv, *slot_t;
    int PJ, J, lib_vmax;
    // scenes
    const float4 *ob_f;
    const double *ob_v64;
    const int *ob_nv;
    int ob_stride, ob_vmax;
    const int *step_obst_off, *step_seg_begin, *step_seg_end;
    const float4 *sg_f;
    const double *sg_64;
    int sg_stride;
    const float *maxabs_scene;
    // queries
    const QDev *q;
    const int *cta_off, *cand;
    int B;
    unsigned *mask;
    int4 *wl;
    int wl_cap;
    Counters *cnt;
    float maxabs_query;
    unsigned mode;
    int gpc, cap_o, cap_s;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (global -> shared::cta)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                             unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__host__ __device__ constexpr int ob_fields(int VB) { return VB / 2 + (VB * 3 + 3) / 4 + 1; }

// ---------------------------------------------------------------------------------------------------------------------
// staging kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_idx(const int *a, int n, int v) {   // largest i with a[i] <= v, a sorted
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void atomic_max_pos_float(float *addr, float v) {
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
}

template <int VB>
__global__ void k_stage_obst(int nobst, double *__restrict__ ob64, const int *__restrict__ ob_nv,
                             const int *__restrict__ step_obst_off, int nsteps, const int *__restrict__ scene_step_off,
                             int nscenes, const double *__restrict__ origins, float4 *__restrict__ ob_f, int stride,
                             float *maxabs) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nobst) return;
    constexpr int F = ob_fields(VB);
    int nv = ob_nv[o];
    float out[F * 4];
    double *v = ob64 + (size_t)o * VB * 2;
    bool ok = nv >= 3 && nv <= VB;
    double x[VB], y[VB];
    if (ok) {
        for (int k = 0; k < VB; k++) { x[k] = v[2 * (k < nv ? k : nv - 1)]; y[k] = v[2 * (k < nv ? k : nv - 1) + 1]; }
        double area = 0.0;
        for (int k = 0; k < nv; k++) {
            int k1 = (k + 1 == nv) ? 0 : k + 1;
            area += (x[k] - x[0]) * (y[k1] - y[0]) - (x[k1] - x[0]) * (y[k] - y[0]);
        }
        if (area == 0.0) ok = false;
        if (area < 0.0) {   // reverse to counter-clockwise, keep the float64 copy in the same order
            for (int i = 0, k = nv - 1; i < k; i++, k--) {
                double tx = x[i], ty = y[i];
                x[i] = x[k]; y[i] = y[k]; x[k] = tx; y[k] = ty;
            }
            for (int k = nv; k < VB; k++) { x[k] = x[nv - 1]; y[k] = y[nv - 1]; }
        }
    }
    if (!ok) {
        // never near, never separating-axis-positive for anyone: a point far away
        for (int k = 0; k < VB; k++) { out[2 * k] = FAR; out[2 * k + 1] = FAR; }
        for (int k = 0; k < VB; k++) { out[VB * 2 + 3 * k] = 0.f; out[VB * 2 + 3 * k + 1] = 0.f; out[VB * 2 + 3 * k + 2] = -BIG; }
        for (int k = VB * 5; k < F * 4 - 4; k++) out[k] = 0.f;
        out[F * 4 - 4] = FAR; out[F * 4 - 3] = FAR; out[F * 4 - 2] = 0.f; out[F * 4 - 1] = 0.f;
    } else {
        int step = upper_bound_idx(step_obst_off, nsteps + 1, o);
        int sc = upper_bound_idx(scene_step_off, nscenes + 1, step);
        double ox = origins[2 * sc], oy = origins[2 * sc + 1];
        for (int k = 0; k < VB; k++) { v[2 * k] = x[k]; v[2 * k + 1] = y[k]; }
        double cx = 0.0, cy = 0.0, m = 0.0;
        for (int k = 0; k < VB; k++) {
            double lx = x[k] - ox, ly = y[k] - oy;
            out[2 * k] = (float)lx; out[2 * k + 1] = (float)ly;
            m = fmax(m, fmax(fabs(lx), fabs(ly)));
            if (k < nv) { cx += lx; cy += ly; }
        }
        cx /= nv; cy /= nv;
        double r = 0.0;
        for (int k = 0; k < nv; k++) {
            double dx = x[k] - ox - cx, dy = y[k] - oy - cy;
            r = fmax(r, sqrt(dx * dx + dy * dy));
        }
        for (int k = 0; k < VB; k++) {
            float nx = 0.f, ny = 0.f, c = BIG;       // padded / degenerate edge: never a separating axis
            if (k < nv) {
                int k1 = (k + 1 == nv) ? 0 : k + 1;
                double ex = x[k1] - x[k], ey = y[k1] - y[k];
                double len = sqrt(ex * ex + ey * ey);
                if (len > 0.0) {
                    double dnx = ey / len, dny = -ex / len;            // outward normal of a CCW edge
                    nx = (float)dnx; ny = (float)dny;
                    c = (float)(dnx * (x[k] - ox) + dny * (y[k] - oy));
                }
            }
            out[VB * 2 + 3 * k] = nx; out[VB * 2 + 3 * k + 1] = ny; out[VB * 2 + 3 * k + 2] = c;
        }
        for (int k = VB * 5; k < F * 4 - 4; k++) out[k] = 0.f;
        out[F * 4 - 4] = (float)cx; out[F * 4 - 3] = (float)cy;
        out[F * 4 - 2] = __double2float_ru(r * (1.0 + 1e-6) + m * 2.4e-7);
        out[F * 4 - 1] = (float)nv;
        atomic_max_pos_float(maxabs, __double2float_ru(m));
    }
#pragma unroll
    for (int f = 0; f < F; f++)
        ob_f[(size_t)f * stride + o] = make_float4(out[4 * f], out[4 * f + 1], out[4 * f + 2], out[4 * f + 3]);
}

__global__ void k_stage_segs(int nseg, const double *__restrict__ sg64, const int *__restrict__ seg_scene,
                             const double *__restrict__ origins, float4 *__restrict__ sg_f, int stride,
                             float *maxabs) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    int sc = seg_scene[s];
    double ax = sg64[4 * s], ay = sg64[4 * s + 1], bx = sg64[4 * s + 2], by = sg64[4 * s + 3];
    double ex = bx - ax, ey = by - ay;
    double len = sqrt(ex * ex + ey * ey);
    if (sc < 0 || !(len > 0.0)) {      // unused or degenerate: far away, never near, never straddling
        sg_f[s] = make_float4(FAR, FAR, FAR, FAR);
        sg_f[(size_t)stride + s] = make_float4(0.f, 0.f, BIG, -1.f);
        return;
    }
    double ox = origins[2 * sc], oy = origins[2 * sc + 1];
    double lax = ax - ox, lay = ay - oy, lbx = bx - ox, lby = by - oy;
    double nx = ey / len, ny = -ex / len;      // right-hand normal of a->b
    sg_f[s] = make_float4((float)lax, (float)lay, (float)lbx, (float)lby);
    sg_f[(size_t)stride + s] = make_float4((float)nx, (float)ny, (float)(nx * lax + ny * lay),
                                           __double2float_ru(0.5 * len * (1.0 + 1e-6)));
    double m = fmax(fmax(fabs(lax), fabs(lay)), fmax(fabs(lbx), fabs(lby)));
    atomic_max_pos_float(maxabs, __double2float_ru(m));
}

// ---------------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void push_work(const KParams &p, int q, int ci, int j, int kind, int obj) {
    unsigned idx = atomicAdd(&p.cnt->wl_count, 1u);
    if (idx < (unsigned)p.wl_cap) p.wl[idx] = make_int4(q, ci, j | (kind << 16), obj);
}

// full separating-axis test of the lane's polygon against obstacle record `o` of the shared-memory tile.
// returns the signed separation (max over all edge axes of the minimum vertex distance); > 0 apart.
template <int VA, int VB>
__device__ __forceinline__ float sat_obstacle(const float (&ax)[VA], const float (&ay)[VA], const float (&anx)[VA],
                                              const float (&any_)[VA], const float (&ac)[VA], const float4 *s_ob,
                                              int cap_o, int o) {
    float bx[VB], by[VB];
#pragma unroll
    for (int h = 0; h < VB / 2; h++) {
        float4 v = s_ob[h * cap_o + o];
        bx[2 * h] = v.x; by[2 * h] = v.y; bx[2 * h + 1] = v.z; by[2 * h + 1] = v.w;
    }
    float sep = -BIG;
#pragma unroll
    for (int e = 0; e < VA; e++) {
        float mn = BIG;
#pragma unroll
        for (int k = 0; k < VB; k++) mn = fminf(mn, fmaf(anx[e], bx[k], fmaf(any_[e], by[k], -ac[e])));
        sep = fmaxf(sep, mn);
    }
    float ln[((VB * 3 + 3) / 4) * 4];
#pragma unroll
    for (int h = 0; h < (VB * 3 + 3) / 4; h++) {
        float4 v = s_ob[(VB / 2 + h) * cap_o + o];
        ln[4 * h] = v.x; ln[4 * h + 1] = v.y; ln[4 * h + 2] = v.z; ln[4 * h + 3] = v.w;
    }
#pragma unroll
    for (int e = 0; e < VB; e++) {
        float nx = ln[3 * e], ny = ln[3 * e + 1], c = ln[3 * e + 2];
        float mn = BIG;
#pragma unroll
        for (int k = 0; k < VA; k++) mn = fminf(mn, fmaf(nx, ax[k], fmaf(ny, ay[k], -c)));
        sep = fmaxf(sep, mn);
    }
    return sep;
}

template <int VA, int VB, bool COUNT>
__global__ void __launch_bounds__(128, 6) k_check(const KParams p) {
    constexpr int F = ob_fields(VB);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *s_coll = reinterpret_cast<unsigned *>(smem_raw + 16);
    unsigned *s_par = s_coll + GPC_MAX;
    unsigned *s_punc = s_par + GPC_MAX;
    float4 *s_ob = reinterpret_cast<float4 *>(smem_raw + 16 + 3 * GPC_MAX * 4);
    float4 *s_sg = s_ob + F * p.cap_o;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    // which (query, slot, chunk) is this CTA
    int q = upper_bound_idx(p.cta_off, p.B + 1, (int)blockIdx.x);
    const QDev *Q = p.q + q;
    int local = blockIdx.x - p.cta_off[q];
    int chunks = Q->chunks;
    int j = local / chunks, chunk = local - j * chunks;
    int t = Q->t0 + p.slot_t[j];
    // reference guards: ind + time <= param['steps'] and ind + time < len(space)  (lowLevelPlannerManeuverAutomaton.py:120,122)
    if (t > Q->step_limit || t < 0 || t >= Q->nsteps) return;
    int step = Q->step0 + t;
    const int o0 = p.step_obst_off[step], nobst = p.step_obst_off[step + 1] - o0;
    const int g0 = p.step_seg_begin[step];
    const bool has_region = g0 >= 0;
    const int nseg = has_region ? p.step_seg_end[step] - g0 : 0;
    const int cand_cnt = Q->cand_cnt, cand_base = Q->cand_base;
    const int gbeg = chunk * p.gpc;
    const int ng = min(p.gpc, (cand_cnt + 31) / 32 - gbeg);
    const float lx = Q->lx, ly = Q->ly, lc = Q->lc, ls = Q->ls;
    // float32 error band: every float32 signed distance below is within tau of its float64 value
    const float tau = fmaxf(fmaxf(*p.maxabs_scene, p.maxabs_query) * (1.0f / 131072.0f), 1e-7f);
    const bool noexit = p.mode & MPC_MODE_NO_EARLY_EXIT, nocull = p.mode & MPC_MODE_NO_CULL;

    if (tid == 0) mbar_init(bar, 1);
    for (int i = tid; i < 3 * GPC_MAX; i += blockDim.x) s_coll[i] = 0;
    __syncthreads();

    const int nph_o = (nobst + p.cap_o - 1) / p.cap_o, nph_s = (nseg + p.cap_s - 1) / p.cap_s;
    const int nphase = max(1, max(nph_o, nph_s));
    unsigned long long n_cull = 0, n_axis = 0;

    for (int ph = 0; ph < nphase; ph++) {
        const int oc0 = ph * p.cap_o, ocn = max(0, min(nobst - oc0, p.cap_o));
        const int sc0 = ph * p.cap_s, scn = max(0, min(nseg - sc0, p.cap_s));
        const bool loading = (ocn | scn) != 0;
        if (tid == 0 && loading) {
            mbar_expect_tx(bar, (unsigned)(ocn * 16 * F + scn * 32));
            if (ocn)
                for (int f = 0; f < F; f++)
                    tma_bulk_g2s(s_ob + f * p.cap_o, p.ob_f + (size_t)f * p.ob_stride + o0 + oc0, ocn * 16, bar);
            if (scn)
                for (int f = 0; f < 2; f++)
                    tma_bulk_g2s(s_sg + f * p.cap_s, p.sg_f + (size_t)f * p.sg_stride + g0 + sc0, scn * 16, bar);
        }
        bool waited = false;
        for (int g = warp; g < ng; g += nwarps) {
            // ---- lane = one candidate primitive: load and transform its occupancy polygon -------------------------
            const int ci = (gbeg + g) * 32 + lane;
            bool active = ci < cand_cnt;
            int pj = 0;
            if (active) pj = p.cand[cand_base + ci] * p.J + j;
            float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) cc = __ldg(p.lib_c + pj);
            const int nv = (int)cc.w;
            active = active && nv > 0;
            float ax[VA], ay[VA], anx[VA], any_[VA], ac[VA];
#pragma unroll
            for (int h = 0; h < VA / 2; h++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
                if (active) { v = __ldg(p.lib_v + (size_t)h * p.PJ + pj); n = __ldg(p.lib_n + (size_t)h * p.PJ + pj); }
                ax[2 * h] = fmaf(lc, v.x, fmaf(-ls, v.y, lx));
                ay[2 * h] = fmaf(ls, v.x, fmaf(lc, v.y, ly));
                ax[2 * h + 1] = fmaf(lc, v.z, fmaf(-ls, v.w, lx));
                ay[2 * h + 1] = fmaf(ls, v.z, fmaf(lc, v.w, ly));
                anx[2 * h] = fmaf(lc, n.x, -ls * n.y);
                any_[2 * h] = fmaf(ls, n.x, lc * n.y);
                anx[2 * h + 1] = fmaf(lc, n.z, -ls * n.w);
                any_[2 * h + 1] = fmaf(ls, n.z, lc * n.w);
            }
#pragma unroll
            for (int e = 0; e < VA; e++) ac[e] = (e < nv) ? fmaf(anx[e], ax[e], any_[e] * ay[e]) : BIG;
            const float acx = active ? fmaf(lc, cc.x, fmaf(-ls, cc.y, lx)) : -FAR;
            const float acy = active ? fmaf(ls, cc.x, fmaf(lc, cc.y, ly)) : -FAR;
            const float arr = cc.z + 2.0f * tau;

            bool collide = (s_coll[g] >> lane) & 1u;
            bool par = false, punc = false;

            if (ph == 0) {   // statistic: nominal pair count of this tile
                unsigned w = __reduce_add_sync(FULL, active ? (unsigned)(nobst + nseg) : 0u);
                if (lane == 0 && w) atomicAdd(&p.cnt->pairs_nominal, (unsigned long long)w);
            }
            if (!waited && loading) { mbar_wait(bar, ph & 1); waited = true; }

            // ---- obstacles of this time step --------------------------------------------------------------------
            const float4 *s_circ = s_ob + (F - 1) * p.cap_o;
            for (int ob = 0; ob < ocn; ob += 32) {
                const int cnt = min(32, ocn - ob);
                // axis 0: centre line with bounding radii.  One warp-broadcast LDS.128 per obstacle; the sign bit of
                // e = |dc|^2 - (ra + rb + 2 tau)^2 is shifted into the hit word (bit 31-k <-> obstacle ob+k).
                unsigned hits = 0;
                if (!nocull) {
                    if (cnt == 32) {
#pragma unroll
                        for (int k = 0; k < 32; k++) {
                            float4 c = s_circ[ob + k];
                            float dx = c.x - acx, dy = c.y - acy, rr = c.z + arr;
                            float e = fmaf(dx, dx, fmaf(dy, dy, -(rr * rr)));
                            hits = __funnelshift_l(__float_as_uint(e), hits, 1);
                        }
                    } else {
                        for (int k = 0; k < cnt; k++) {
                            float4 c = s_circ[ob + k];
                            float dx = c.x - acx, dy = c.y - acy, rr = c.z + arr;
                            float e = fmaf(dx, dx, fmaf(dy, dy, -(rr * rr)));
                            hits = __funnelshift_l(__float_as_uint(e), hits, 1);
                        }
                        hits <<= 32 - cnt;
                    }
                } else {
                    hits = ~((cnt == 32) ? 0u : (FULL >> cnt));
                }
                if (!active || (collide && !noexit)) hits = 0;
                if (COUNT && active) n_cull += cnt;
                while (hits) {
                    int k = __clz(hits);
                    hits &= ~(0x80000000u >> k);
                    float sep = sat_obstacle<VA, VB>(ax, ay, anx, any_, ac, s_ob, p.cap_o, ob + k);
                    if (COUNT) n_axis++;
                    if (sep < -tau) {
                        collide = true;
                        if (!noexit) hits = 0;
                    } else if (!(sep > tau)) {
                        push_work(p, q, ci, j, KIND_OBST, o0 + oc0 + ob + k);
                    }
                }
                if (!noexit && __all_sync(FULL, collide || !active)) break;   // warp-wide early exit
            }

            // ---- boundary segments of the free space at this time step -----------------------------------------
            for (int sb = 0; sb < scn; sb += 32) {
                const int cnt = min(32, scn - sb);
                unsigned hits = 0;
#pragma unroll 2
                for (int k = 0; k < cnt; k++) {
                    float4 sa = s_sg[sb + k], sl = s_sg[p.cap_s + sb + k];
                    float dl = fmaf(sl.x, acx, fmaf(sl.y, acy, -sl.z));       // centre to the segment's line
                    // even-odd crossing parity of the +x ray from the polygon centre
                    bool ua = sa.y > acy, ub = sa.w > acy;
                    if (ua != ub) {
                        par ^= ub ? (dl < 0.f) : (dl > 0.f);
                        punc |= fabsf(dl) <= tau;
                    }
                    punc |= (fabsf(sa.y - acy) <= tau) | (fabsf(sa.w - acy) <= tau);
                    float tt = fmaf(-sl.y, acx - 0.5f * (sa.x + sa.z), sl.x * (acy - 0.5f * (sa.y + sa.w)));
                    bool near = nocull ? (sl.w >= 0.f) : ((fabsf(dl) <= arr) & (fabsf(tt) <= sl.w + arr));
                    if (near) hits |= 1u << k;
                }
                if (!active || (collide && !noexit)) hits = 0;
                if (COUNT && active) n_cull += cnt;
                while (hits) {
                    int k = __ffs(hits) - 1;
                    hits &= hits - 1;
                    float4 sa = s_sg[sb + k], sl = s_sg[p.cap_s + sb + k];
                    float sep = -BIG, mn = BIG, mx = -BIG;
#pragma unroll
                    for (int e = 0; e < VA; e++) {
                        float d0 = fmaf(anx[e], sa.x, fmaf(any_[e], sa.y, -ac[e]));
                        float d1 = fmaf(anx[e], sa.z, fmaf(any_[e], sa.w, -ac[e]));
                        sep = fmaxf(sep, fminf(d0, d1));
                        float d = fmaf(sl.x, ax[e], fmaf(sl.y, ay[e], -sl.z));
                        mn = fminf(mn, d);
                        mx = fmaxf(mx, d);
                    }
                    sep = fmaxf(sep, fmaxf(mn, -mx));
                    if (COUNT) n_axis++;
                    if (sep < -tau) {
                        collide = true;
                        if (!noexit) hits = 0;
                    } else if (!(sep > tau)) {
                        push_work(p, q, ci, j, KIND_SEG, g0 + sc0 + sb + k);
                    }
                }
                // no warp-wide exit here: lanes that are still free need the parity of every segment
                if (!noexit && __all_sync(FULL, collide || !active)) break;
            }

            // ---- fold this phase into the group state ----------------------------------------------------------
            unsigned bc = __ballot_sync(FULL, collide), bp = __ballot_sync(FULL, par), bu = __ballot_sync(FULL, punc);
            if (nphase > 1) {
                if (lane == 0) { s_coll[g] |= bc; s_par[g] ^= bp; s_punc[g] |= bu; }
                __syncwarp();
                bc = s_coll[g]; bp = s_par[g]; bu = s_punc[g];
            }
            if (ph == nphase - 1) {
                bool coll = (bc >> lane) & 1u;
                if (has_region && active && !coll) {
                    if ((bu >> lane) & 1u) push_work(p, q, ci, j, KIND_PARITY, 0);
                    else if (!((bp >> lane) & 1u)) coll = true;        // centre outside the region
                }
                unsigned word = __ballot_sync(FULL, coll && active);
                if (lane == 0 && word) atomicOr(p.mask + Q->mask_off + gbeg + g, word);
            }
        }
        if (nphase > 1) {
            if (!waited && loading) mbar_wait(bar, ph & 1);
            __syncthreads();          // everyone is done with the tile before the next bulk copy overwrites it
        }
    }
    if (COUNT) {
        n_cull = __reduce_add_sync(FULL, (unsigned)n_cull) ;
        unsigned a = __reduce_add_sync(FULL, (unsigned)n_axis);
        if (lane == 0) {
            atomicAdd(&p.cnt->cull_tests, n_cull);
            atomicAdd(&p.cnt->axis_tests, (unsigned long long)a);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// float64 / exact refinement of the pairs inside the float32 band
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_refine(const KParams p) {
    const unsigned n = min(p.cnt->wl_count, (unsigned)p.wl_cap);
    unsigned refined = 0, ties = 0, near_t = 0, par_ref = 0;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int4 e = p.wl[idx];
        const int q = e.x, ci = e.y, j = e.z & 0xffff, kind = e.z >> 16, obj = e.w;
        const QDev *Q = p.q + q;
        const int pj = p.cand[Q->cand_base + ci] * p.J + j;
        Poly64 A;
        A.n = p.lib_nv[pj];
        const double *lv = p.lib_v64 + (size_t)pj * p.lib_vmax * 2;
        // shapely affine_transform(o['space'], [cos, -sin, sin, cos, x, y]): xp = a*x + b*y + xoff, left to right
        const double a = Q->c, b = -Q->s, d = Q->s, ee = Q->c;
        for (int k = 0; k < A.n; k++) {
            A.x[k] = __dadd_rn(__dadd_rn(__dmul_rn(a, lv[2 * k]), __dmul_rn(b, lv[2 * k + 1])), Q->x);
            A.y[k] = __dadd_rn(__dadd_rn(__dmul_rn(d, lv[2 * k]), __dmul_rn(ee, lv[2 * k + 1])), Q->y);
        }
        bool collide = false;
        if (kind == KIND_OBST) {
            Poly64 Bp;
            Bp.n = p.ob_nv[obj];
            const double *bv = p.ob_v64 + (size_t)obj * p.ob_vmax * 2;
            for (int k = 0; k < Bp.n; k++) { Bp.x[k] = bv[2 * k]; Bp.y[k] = bv[2 * k + 1]; }
            collide = convex_collide(A, Bp, p.mode & MPC_MODE_VALIDATOR, ties);
            if (fabs(approx_sep(A, Bp)) < 1e-9) near_t++;
            refined++;
        } else if (kind == KIND_SEG) {
            const double *sg = p.sg_64 + 4 * (size_t)obj;
            collide = segment_meets_interior(A, sg[0], sg[1], sg[2], sg[3], ties);
            refined++;
        } else {
            double cx = 0.0, cy = 0.0;
            for (int k = 0; k < A.n; k++) { cx += A.x[k]; cy += A.y[k]; }
            cx /= A.n; cy /= A.n;
            const int step = Q->step0 + Q->t0 + p.slot_t[j];
            const int g0 = p.step_seg_begin[step], g1 = p.step_seg_end[step];
            bool inside = false, on_edge = false;
            for (int g = g0; g < g1; g++) {
                const double *sg = p.sg_64 + 4 * (size_t)g;
                const double sax = sg[0], say = sg[1], sbx = sg[2], sby = sg[3];
                int o = orient2d(sax, say, sbx, sby, cx, cy, ties);
                if (o == 0 && fmin(sax, sbx) <= cx && cx <= fmax(sax, sbx) && fmin(say, sby) <= cy && cy <= fmax(say, sby))
                    on_edge = true;
                if ((say > cy) != (sby > cy)) {
                    if ((sby > say && o > 0) || (sby < say && o < 0)) inside = !inside;
                }
            }
            collide = on_edge || !inside;
            par_ref++;
        }
        if (collide) atomicOr(p.mask + Q->mask_off + (ci >> 5), 1u << (ci & 31));
    }
    refined = __reduce_add_sync(FULL, refined);
    ties = __reduce_add_sync(FULL, ties);
    near_t = __reduce_add_sync(FULL, near_t);
    par_ref = __reduce_add_sync(FULL, par_ref);
    if ((threadIdx.x & 31) == 0) {
        if (refined) atomicAdd(&p.cnt->refined, (unsigned long long)refined);
        if (ties) atomicAdd(&p.cnt->exact_ties, (unsigned long long)ties);
        if (near_t) atomicAdd(&p.cnt->near_tangent, (unsigned long long)near_t);
        if (par_ref) atomicAdd(&p.cnt->parity_refined, (unsigned long long)par_ref);
    }
}

__global__ void k_fill_u32(unsigned *p, size_t n, unsigned v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace mpc

// =====================================================================================================================
// host side
// =====================================================================================================================
using namespace mpc;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct mpc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    cudaDeviceProp prop;
    // library
    bool have_lib = false;
    int P = 0, J = 0, lib_vmax = 0, VA = 4;
    double lib_radius = 0.0;
    DevBuf lib_v, lib_n, lib_c, lib_v64, lib_nv, slot_t, sets;
    std::vector<int> set_off;
    int nsets = 0, set_total = 0;
    bool arena_valid = false;
    // scenes
    bool have_scenes = false;
    int nscenes = 0, nsteps = 0, nobst = 0, nseg = 0, ob_vmax = 4, VB = 4, max_obst_step = 0, max_seg_step = 0;
    std::vector<int> scene_step_off;
    std::vector<double> origins;
    DevBuf ob_in, ob_f, ob_nv, step_obst_off, step_seg_begin, step_seg_end, scene_step_off_d, origins_d, sg_64, sg_f,
        seg_scene, maxabs;
    // queries
    DevBuf q_d, cta_off_d, cand_d, mask_d, wl, cnt, flush;
    int wl_cap = 0;
    void *h_pin = nullptr;
    size_t h_pin_cap = 0;
    Counters *h_cnt = nullptr;
    // prepared batch
    bool prepared = false;
    int B = 0, ctas = 0, mask_words = 0, gpc = 1, cap_o = 32, cap_s = 32, threads = 128;
    size_t smem = 0;
    unsigned mode = 0;
    float maxabs_query = 0.f;
};

static int fail(mpc_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, MPC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static cudaError_t ensure(DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max<size_t>(bytes + bytes / 4, 256);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

static cudaError_t ensure_pinned(mpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_pin_cap) return cudaSuccess;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->h_pin_cap = 0;
    size_t want = std::max<size_t>(bytes * 2, 1 << 20);
    cudaError_t e = cudaMallocHost(&ctx->h_pin, want);
    if (e == cudaSuccess) ctx->h_pin_cap = want;
    return e;
}

extern "C" const char *mpc_last_error(const mpc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int mpc_create(int device, mpc_ctx **out) {
    if (!out) return MPC_E_ARG;
    *out = nullptr;
    mpc_ctx *ctx = new mpc_ctx();
    ctx->device = device;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        fail(ctx, MPC_E_CUDA, "no usable CUDA device %d (count=%d, %s); this library has no CPU fallback", device, n,
             cudaGetErrorString(e));
        *out = ctx;      // returned so the caller can read the message; still a failure
        return MPC_E_CUDA;
    }
    *out = ctx;
    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major < 9) return fail(ctx, MPC_E_CUDA, "device sm_%d%d lacks TMA bulk copies; built for sm_100a", ctx->prop.major, ctx->prop.minor);
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) CU(cudaEventCreate(&ev));
    CU(cudaMallocHost((void **)&ctx->h_cnt, sizeof(Counters)));
    CU(ensure(ctx->cnt, sizeof(Counters)));
    CU(ensure(ctx->maxabs, sizeof(float)));
    CU(cudaMemsetAsync(ctx->maxabs.p, 0, sizeof(float), ctx->stream));
    ctx->wl_cap = 1 << 18;
    CU(ensure(ctx->wl, sizeof(int4) * (size_t)ctx->wl_cap));
    return MPC_OK;
}

extern "C" void mpc_destroy(mpc_ctx *ctx) {
    if (!ctx) return;
    if (ctx->stream) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    DevBuf *bufs[] = {&ctx->lib_v, &ctx->lib_n, &ctx->lib_c, &ctx->lib_v64, &ctx->lib_nv, &ctx->slot_t, &ctx->sets,
                      &ctx->ob_in, &ctx->ob_f, &ctx->ob_nv, &ctx->step_obst_off, &ctx->step_seg_begin,
                      &ctx->step_seg_end, &ctx->scene_step_off_d, &ctx->origins_d, &ctx->sg_64, &ctx->sg_f,
                      &ctx->seg_scene, &ctx->maxabs, &ctx->q_d, &ctx->cta_off_d, &ctx->cand_d, &ctx->mask_d, &ctx->wl,
                      &ctx->cnt, &ctx->flush};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int mpc_device_info(mpc_ctx *ctx, mpc_devinfo *out) {
    if (!ctx || !out) return MPC_E_ARG;
    if (!ctx->stream) return fail(ctx, MPC_E_STATE, "context has no device");
    memset(out, 0, sizeof(*out));
    strncpy(out->name, ctx->prop.name, sizeof(out->name) - 1);
    out->sm_count = ctx->prop.multiProcessorCount;
    out->cc_major = ctx->prop.major;
    out->cc_minor = ctx->prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    out->clock_khz = khz;
    out->global_mem = (int64_t)ctx->prop.totalGlobalMem;
    out->l2_bytes = ctx->prop.l2CacheSize;
    out->abi_version = MPC_ABI_VERSION;
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mpc_upload_library(mpc_ctx *ctx, const double *verts, const int32_t *nverts, const int32_t *tidx, int P,
                                  int J, int Vmax, const int32_t *set_off, const int32_t *set_idx, int nsets) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!verts || !nverts || !tidx || P <= 0 || J <= 0 || Vmax < 3) return fail(ctx, MPC_E_ARG, "bad library arguments");
    if (J > 0xffff) return fail(ctx, MPC_E_LIMIT, "J=%d occupancy slots exceed 65535", J);
    CU(cudaSetDevice(ctx->device));
    const size_t PJ = (size_t)P * J;
    int vmax_used = 0;
    for (size_t i = 0; i < PJ; i++) {
        if (nverts[i] < 0 || nverts[i] > Vmax) return fail(ctx, MPC_E_ARG, "nverts[%zu]=%d outside 0..%d", i, nverts[i], Vmax);
        if (nverts[i] > 0 && nverts[i] < 3) return fail(ctx, MPC_E_ARG, "nverts[%zu]=%d: a polygon needs 3 vertices", i, nverts[i]);
        vmax_used = std::max(vmax_used, (int)nverts[i]);
    }
    if (vmax_used > MPC_MAX_LIB_VERTS) return fail(ctx, MPC_E_LIMIT, "occupancy polygon with %d vertices (max %d)", vmax_used, MPC_MAX_LIB_VERTS);
    const int VA = vmax_used <= 4 ? 4 : (vmax_used <= 8 ? 8 : 16);
    // every primitive that uses slot j must map it to the same time index
    std::vector<int> slot_t(J, 0), slot_set(J, 0);
    for (int p = 0; p < P; p++)
        for (int j = 0; j < J; j++) {
            if (nverts[(size_t)p * J + j] == 0) continue;
            int t = tidx[(size_t)p * J + j];
            if (!slot_set[j]) { slot_t[j] = t; slot_set[j] = 1; }
            else if (slot_t[j] != t)
                return fail(ctx, MPC_E_ARG, "slot %d has time index %d for primitive %d but %d elsewhere: re-slot the library", j, t, p, slot_t[j]);
        }
    std::vector<float4> hv((size_t)(VA / 2) * PJ), hn((size_t)(VA / 2) * PJ), hc(PJ);
    std::vector<double> h64(PJ * (size_t)VA * 2, 0.0);
    std::vector<int> hnv(PJ);
    double radius = 0.0;
    for (size_t i = 0; i < PJ; i++) {
        int nv = nverts[i];
        double x[MPC_MAX_LIB_VERTS], y[MPC_MAX_LIB_VERTS];
        if (nv > 0) {
            const double *src = verts + i * (size_t)Vmax * 2;
            for (int k = 0; k < nv; k++) { x[k] = src[2 * k]; y[k] = src[2 * k + 1]; }
            double area = 0.0;
            for (int k = 0; k < nv; k++) {
                int k1 = (k + 1 == nv) ? 0 : k + 1;
                area += (x[k] - x[0]) * (y[k1] - y[0]) - (x[k1] - x[0]) * (y[k] - y[0]);
            }
            if (area == 0.0) nv = 0;       // degenerate occupancy: nothing to contain
            if (area < 0.0) { std::reverse(x, x + nv); std::reverse(y, y + nv); }
        }
        hnv[i] = nv;
        if (nv == 0) {
            for (int h = 0; h < VA / 2; h++) { hv[(size_t)h * PJ + i] = make_float4(0, 0, 0, 0); hn[(size_t)h * PJ + i] = make_float4(0, 0, 0, 0); }
            hc[i] = make_float4(0, 0, 0, 0);
            continue;
        }
        for (int k = nv; k < VA; k++) { x[k] = x[nv - 1]; y[k] = y[nv - 1]; }
        float nx[MPC_MAX_LIB_VERTS], ny[MPC_MAX_LIB_VERTS];
        double cx = 0, cy = 0;
        for (int k = 0; k < nv; k++) { cx += x[k]; cy += y[k]; }
        cx /= nv; cy /= nv;
        double r = 0;
        for (int k = 0; k < VA; k++) {
            nx[k] = ny[k] = 0.f;
            if (k < nv) {
                int k1 = (k + 1 == nv) ? 0 : k + 1;
                double ex = x[k1] - x[k], ey = y[k1] - y[k], len = std::sqrt(ex * ex + ey * ey);
                if (len > 0) { nx[k] = (float)(ey / len); ny[k] = (float)(-ex / len); }
                r = std::max(r, std::hypot(x[k] - cx, y[k] - cy));
                radius = std::max(radius, std::hypot(x[k], y[k]));
            }
            h64[(i * VA + k) * 2] = x[k];
            h64[(i * VA + k) * 2 + 1] = y[k];
        }
        for (int h = 0; h < VA / 2; h++) {
            hv[(size_t)h * PJ + i] = make_float4((float)x[2 * h], (float)y[2 * h], (float)x[2 * h + 1], (float)y[2 * h + 1]);
            hn[(size_t)h * PJ + i] = make_float4(nx[2 * h], ny[2 * h], nx[2 * h + 1], ny[2 * h + 1]);
        }
        double rr = r * (1.0 + 1e-6) + std::hypot(cx, cy) * 2.4e-7 + 1e-7;
        hc[i] = make_float4((float)cx, (float)cy, std::nextafter((float)rr, INFINITY), (float)nv);
    }
    // note: a zero-length edge keeps normal (0,0); k_check turns its line into "never separating" only for e >= nv,
    // so give such edges the same treatment by dropping them here
    CU(ensure(ctx->lib_v, hv.size() * sizeof(float4)));
    CU(ensure(ctx->lib_n, hn.size() * sizeof(float4)));
    CU(ensure(ctx->lib_c, hc.size() * sizeof(float4)));
    CU(ensure(ctx->lib_v64, h64.size() * sizeof(double)));
    CU(ensure(ctx->lib_nv, hnv.size() * sizeof(int)));
    CU(ensure(ctx->slot_t, slot_t.size() * sizeof(int)));
    CU(cudaMemcpyAsync(ctx->lib_v.p, hv.data(), hv.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lib_n.p, hn.data(), hn.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lib_c.p, hc.data(), hc.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lib_v64.p, h64.data(), h64.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lib_nv.p, hnv.data(), hnv.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->slot_t.p, slot_t.data(), slot_t.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ctx->set_off.assign(1, 0);
    ctx->nsets = 0;
    ctx->set_total = 0;
    if (nsets > 0) {
        if (!set_off || !set_idx) return fail(ctx, MPC_E_ARG, "candidate sets missing");
        for (int s = 0; s < nsets; s++)
            if (set_off[s + 1] < set_off[s]) return fail(ctx, MPC_E_ARG, "set_off not monotone");
        int total = set_off[nsets];
        for (int i = 0; i < total; i++)
            if (set_idx[i] < 0 || set_idx[i] >= P) return fail(ctx, MPC_E_ARG, "set_idx[%d]=%d outside the library", i, set_idx[i]);
        ctx->set_off.assign(set_off, set_off + nsets + 1);
        ctx->nsets = nsets;
        ctx->set_total = total;
        CU(ensure(ctx->sets, std::max<size_t>(total, 1) * sizeof(int)));
        if (total) CU(cudaMemcpyAsync(ctx->sets.p, set_idx, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->P = P; ctx->J = J; ctx->lib_vmax = VA; ctx->VA = VA; ctx->lib_radius = radius;
    ctx->have_lib = true;
    ctx->prepared = false;
    ctx->arena_valid = false;
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mpc_set_scenes(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                              const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                              const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (nscenes <= 0 || !scene_step_off || !origins || !step_obst_off || !step_seg_begin || !step_seg_end)
        return fail(ctx, MPC_E_ARG, "bad scene arguments");
    CU(cudaSetDevice(ctx->device));
    const int nsteps = scene_step_off[nscenes];
    if (scene_step_off[0] != 0 || nsteps < 0) return fail(ctx, MPC_E_ARG, "scene_step_off must start at 0");
    for (int s = 0; s < nscenes; s++)
        if (scene_step_off[s + 1] < scene_step_off[s]) return fail(ctx, MPC_E_ARG, "scene_step_off not monotone");
    const int nobst = step_obst_off[nsteps];
    if (step_obst_off[0] != 0) return fail(ctx, MPC_E_ARG, "step_obst_off must start at 0");
    if (nobst > 0 && (!obst || !obst_nv)) return fail(ctx, MPC_E_ARG, "obstacle arrays missing");
    if (nobst > 0 && (Vo < 3 || Vo > MPC_MAX_OBST_VERTS)) return fail(ctx, MPC_E_LIMIT, "Vo=%d outside 3..%d", Vo, MPC_MAX_OBST_VERTS);
    if (nseg > 0 && !segs) return fail(ctx, MPC_E_ARG, "segment array missing");
    int max_o = 0, max_s = 0;
    std::vector<int> seg_scene(std::max(nseg, 1), -1);
    for (int s = 0; s < nscenes; s++)
        for (int k = scene_step_off[s]; k < scene_step_off[s + 1]; k++) {
            int no = step_obst_off[k + 1] - step_obst_off[k];
            if (no < 0) return fail(ctx, MPC_E_ARG, "step_obst_off not monotone at step %d", k);
            max_o = std::max(max_o, no);
            int b = step_seg_begin[k], e = step_seg_end[k];
            if (b < 0) continue;
            if (e < b || e > nseg) return fail(ctx, MPC_E_ARG, "segment range of step %d outside 0..%d", k, nseg);
            max_s = std::max(max_s, e - b);
            // a range shared by all steps of a scene is marked once
            if (e > b && seg_scene[b] == s && seg_scene[e - 1] == s) continue;
            for (int g = b; g < e; g++) {
                if (seg_scene[g] >= 0 && seg_scene[g] != s) return fail(ctx, MPC_E_ARG, "segment %d is shared by scenes %d and %d", g, seg_scene[g], s);
                seg_scene[g] = s;
            }
        }
    int vmax_used = 3;
    for (int o = 0; o < nobst; o++) {
        if (obst_nv[o] > Vo || obst_nv[o] < 0) return fail(ctx, MPC_E_ARG, "obst_nv[%d]=%d outside 0..%d", o, obst_nv[o], Vo);
        vmax_used = std::max(vmax_used, (int)obst_nv[o]);
    }
    const int VB = vmax_used <= 4 ? 4 : 8;
    const int F = ob_fields(VB);
    const int ob_stride = std::max(nobst, 1), sg_stride = std::max(nseg, 1);
    CU(ensure(ctx->scene_step_off_d, (size_t)(nscenes + 1) * sizeof(int)));
    CU(ensure(ctx->origins_d, (size_t)nscenes * 2 * sizeof(double)));
    CU(ensure(ctx->step_obst_off, (size_t)(nsteps + 1) * sizeof(int)));
    CU(ensure(ctx->step_seg_begin, (size_t)std::max(nsteps, 1) * sizeof(int)));
    CU(ensure(ctx->step_seg_end, (size_t)std::max(nsteps, 1) * sizeof(int)));
    CU(ensure(ctx->ob_in, (size_t)ob_stride * VB * 2 * sizeof(double)));
    CU(ensure(ctx->ob_nv, (size_t)ob_stride * sizeof(int)));
    CU(ensure(ctx->ob_f, (size_t)ob_stride * F * sizeof(float4)));
    CU(ensure(ctx->sg_64, (size_t)sg_stride * 4 * sizeof(double)));
    CU(ensure(ctx->sg_f, (size_t)sg_stride * 2 * sizeof(float4)));
    CU(ensure(ctx->seg_scene, (size_t)sg_stride * sizeof(int)));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->scene_step_off_d.p, scene_step_off, (size_t)(nscenes + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->origins_d.p, origins, (size_t)nscenes * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->step_obst_off.p, step_obst_off, (size_t)(nsteps + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (nsteps) {
        CU(cudaMemcpyAsync(ctx->step_seg_begin.p, step_seg_begin, (size_t)nsteps * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->step_seg_end.p, step_seg_end, (size_t)nsteps * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemsetAsync(ctx->maxabs.p, 0, sizeof(float), st));
    if (nobst) {
        if (Vo == VB) {
            CU(cudaMemcpyAsync(ctx->ob_in.p, obst, (size_t)nobst * VB * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {   // caller's vertex stride differs from the kernel's: repack rows
            CU(cudaMemcpy2DAsync(ctx->ob_in.p, (size_t)VB * 2 * sizeof(double), obst, (size_t)Vo * 2 * sizeof(double),
                                 (size_t)std::min(Vo, VB) * 2 * sizeof(double), nobst, cudaMemcpyHostToDevice, st));
        }
        CU(cudaMemcpyAsync(ctx->ob_nv.p, obst_nv, (size_t)nobst * sizeof(int), cudaMemcpyHostToDevice, st));
        int blocks = (nobst + 127) / 128;
        if (VB == 4)
            k_stage_obst<4><<<blocks, 128, 0, st>>>(nobst, (double *)ctx->ob_in.p, (const int *)ctx->ob_nv.p, (const int *)ctx->step_obst_off.p, nsteps,
                                                   (const int *)ctx->scene_step_off_d.p, nscenes, (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, ob_stride, (float *)ctx->maxabs.p);
        else
            k_stage_obst<8><<<blocks, 128, 0, st>>>(nobst, (double *)ctx->ob_in.p, (const int *)ctx->ob_nv.p, (const int *)ctx->step_obst_off.p, nsteps,
                                                   (const int *)ctx->scene_step_off_d.p, nscenes, (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, ob_stride, (float *)ctx->maxabs.p);
        CU(cudaGetLastError());
    }
    if (nseg) {
        CU(cudaMemcpyAsync(ctx->sg_64.p, segs, (size_t)nseg * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->seg_scene.p, seg_scene.data(), (size_t)nseg * sizeof(int), cudaMemcpyHostToDevice, st));
        k_stage_segs<<<(nseg + 127) / 128, 128, 0, st>>>(nseg, (const double *)ctx->sg_64.p, (const int *)ctx->seg_scene.p, (const double *)ctx->origins_d.p,
                                                       (float4 *)ctx->sg_f.p, sg_stride, (float *)ctx->maxabs.p);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(st));      // host buffers (incl. the local seg_scene) may be reused after return
    ctx->nscenes = nscenes; ctx->nsteps = nsteps; ctx->nobst = nobst; ctx->nseg = nseg;
    ctx->ob_vmax = VB; ctx->VB = VB; ctx->max_obst_step = max_o; ctx->max_seg_step = max_s;
    ctx->scene_step_off.assign(scene_step_off, scene_step_off + nscenes + 1);
    ctx->origins.assign(origins, origins + 2 * (size_t)nscenes);
    ctx->have_scenes = true;
    ctx->prepared = false;
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
static KParams make_params(mpc_ctx *ctx) {
    KParams p;
    memset(&p, 0, sizeof p);
    p.lib_v = (const float4 *)ctx->lib_v.p; p.lib_n = (const float4 *)ctx->lib_n.p; p.lib_c = (const float4 *)ctx->lib_c.p;
    p.lib_v64 = (const double *)ctx->lib_v64.p; p.lib_nv = (const int *)ctx->lib_nv.p; p.slot_t = (const int *)ctx->slot_t.p;
    p.PJ = ctx->P * ctx->J; p.J = ctx->J; p.lib_vmax = ctx->lib_vmax;
    p.ob_f = (const float4 *)ctx->ob_f.p; p.ob_v64 = (const double *)ctx->ob_in.p; p.ob_nv = (const int *)ctx->ob_nv.p;
    p.ob_stride = std::max(ctx->nobst, 1); p.ob_vmax = ctx->ob_vmax;
    p.step_obst_off = (const int *)ctx->step_obst_off.p; p.step_seg_begin = (const int *)ctx->step_seg_begin.p;
    p.step_seg_end = (const int *)ctx->step_seg_end.p;
    p.sg_f = (const float4 *)ctx->sg_f.p; p.sg_64 = (const double *)ctx->sg_64.p; p.sg_stride = std::max(ctx->nseg, 1);
    p.maxabs_scene = (const float *)ctx->maxabs.p;
    p.q = (const QDev *)ctx->q_d.p; p.cta_off = (const int *)ctx->cta_off_d.p; p.cand = (const int *)ctx->cand_d.p;
    p.B = ctx->B; p.mask = (unsigned *)ctx->mask_d.p; p.wl = (int4 *)ctx->wl.p; p.wl_cap = ctx->wl_cap;
    p.cnt = (Counters *)ctx->cnt.p; p.maxabs_query = ctx->maxabs_query; p.mode = ctx->mode;
    p.gpc = ctx->gpc; p.cap_o = ctx->cap_o; p.cap_s = ctx->cap_s;
    return p;
}

template <int VA, int VB>
static cudaError_t launch_check_vb(mpc_ctx *ctx, const KParams &p) {
    auto kern = (ctx->mode & MPC_MODE_COUNT_WORK) ? k_check<VA, VB, true> : k_check<VA, VB, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem);
    if (e != cudaSuccess) return e;
    kern<<<ctx->ctas, ctx->threads, ctx->smem, ctx->stream>>>(p);
    return cudaGetLastError();
}

static cudaError_t launch_check(mpc_ctx *ctx, const KParams &p) {
    if (ctx->ctas == 0) return cudaSuccess;
    const int va = ctx->VA, vb = ctx->VB;
    if (va == 4 && vb == 4) return launch_check_vb<4, 4>(ctx, p);
    if (va == 4 && vb == 8) return launch_check_vb<4, 8>(ctx, p);
    if (va == 8 && vb == 4) return launch_check_vb<8, 4>(ctx, p);
    if (va == 8 && vb == 8) return launch_check_vb<8, 8>(ctx, p);
    if (va == 16 && vb == 4) return launch_check_vb<16, 4>(ctx, p);
    return launch_check_vb<16, 8>(ctx, p);
}

static int launch_all(mpc_ctx *ctx, cudaEvent_t e0, cudaEvent_t e1, cudaEvent_t e2) {
    KParams p = make_params(ctx);
    cudaStream_t st = ctx->stream;
    if (e0) CU(cudaEventRecord(e0, st));
    CU(cudaMemsetAsync(ctx->cnt.p, 0, sizeof(Counters), st));
    CU(cudaMemsetAsync(ctx->mask_d.p, 0, (size_t)std::max(ctx->mask_words, 1) * sizeof(unsigned), st));
    if (e1) CU(cudaEventRecord(e1, st));
    CU(launch_check(ctx, p));
    if (e2) CU(cudaEventRecord(e2, st));
    k_refine<<<std::max(1, ctx->prop.multiProcessorCount), 128, 0, st>>>(p);
    CU(cudaGetLastError());
    return MPC_OK;
}

extern "C" int mpc_prepare_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                                 unsigned mode) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->have_lib || !ctx->have_scenes) return fail(ctx, MPC_E_STATE, "upload the library and the scenes before querying");
    if (B < 0 || (B > 0 && !queries)) return fail(ctx, MPC_E_ARG, "bad query arguments");
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    // grid shape: groups of 32 candidates, gpc groups per CTA; aim for a few CTAs per SM
    long long total_groups = 0;
    for (int q = 0; q < B; q++) {
        const mpc_query_desc &d = queries[q];
        if (d.scene < 0 || d.scene >= ctx->nscenes) return fail(ctx, MPC_E_ARG, "query %d: scene %d outside 0..%d", q, d.scene, ctx->nscenes - 1);
        if (d.cand_cnt < 0 || d.cand_off < 0) return fail(ctx, MPC_E_ARG, "query %d: negative candidate range", q);
        if (d.cand_set >= 0) {
            if (d.cand_set >= ctx->nsets) return fail(ctx, MPC_E_ARG, "query %d: candidate set %d not staged", q, d.cand_set);
            int len = ctx->set_off[d.cand_set + 1] - ctx->set_off[d.cand_set];
            if (d.cand_off + d.cand_cnt > len) return fail(ctx, MPC_E_ARG, "query %d: range exceeds candidate set %d (%d)", q, d.cand_set, len);
        } else {
            if (d.cand_off + d.cand_cnt > ncand || (d.cand_cnt > 0 && !cand_idx)) return fail(ctx, MPC_E_ARG, "query %d: range exceeds cand_idx (%d)", q, ncand);
        }
        total_groups += ((long long)d.cand_cnt + 31) / 32 * ctx->J;
    }
    for (int i = 0; i < ncand; i++)
        if (cand_idx[i] < 0 || cand_idx[i] >= ctx->P) return fail(ctx, MPC_E_ARG, "cand_idx[%d]=%d outside the library", i, cand_idx[i]);
    const int sms = ctx->prop.multiProcessorCount;
    ctx->threads = 128;
    const int nwarps = ctx->threads / 32;
    int gpc = (int)((total_groups + (long long)sms * 6 - 1) / ((long long)sms * 6));
    gpc = std::max(1, std::min(gpc, GPC_MAX));
    gpc = std::min(GPC_MAX, (gpc + nwarps - 1) / nwarps * nwarps);   // whole rounds of the CTA's warps
    ctx->gpc = gpc;
    const int F = ob_fields(ctx->VB);
    ctx->cap_o = std::max(32, std::min((ctx->max_obst_step + 31) / 32 * 32, ctx->VB == 4 ? 320 : 160));
    ctx->cap_s = std::max(32, std::min((ctx->max_seg_step + 31) / 32 * 32, 512));
    ctx->smem = 16 + 3 * GPC_MAX * 4 + (size_t)ctx->cap_o * F * 16 + (size_t)ctx->cap_s * 32;
    // pack: QDev[B], cta_off[B+1], explicit candidates
    size_t bytes_q = sizeof(QDev) * (size_t)std::max(B, 1), bytes_c = sizeof(int) * (size_t)(B + 1), bytes_i = sizeof(int) * (size_t)ncand;
    CU(ensure_pinned(ctx, bytes_q + bytes_c + bytes_i + 64));
    QDev *hq = (QDev *)ctx->h_pin;
    int *hcta = (int *)((char *)ctx->h_pin + bytes_q);
    int *hidx = (int *)((char *)ctx->h_pin + bytes_q + bytes_c);
    long long ctas = 0, words = 0;
    double maxq = 0.0;
    for (int q = 0; q < B; q++) {
        const mpc_query_desc &d = queries[q];
        QDev &o = hq[q];
        const double ox = ctx->origins[2 * d.scene], oy = ctx->origins[2 * d.scene + 1];
        o.x = d.x; o.y = d.y; o.c = d.c; o.s = d.s;
        o.lx = (float)(d.x - ox); o.ly = (float)(d.y - oy); o.lc = (float)d.c; o.ls = (float)d.s;
        o.step0 = ctx->scene_step_off[d.scene];
        o.nsteps = ctx->scene_step_off[d.scene + 1] - o.step0;
        o.t0 = d.t0; o.step_limit = d.step_limit;
        o.cand_base = d.cand_set >= 0 ? ctx->set_off[d.cand_set] + d.cand_off : ctx->set_total + d.cand_off;
        o.cand_cnt = d.cand_cnt;
        o.mask_off = (int)words;
        int groups = (d.cand_cnt + 31) / 32;
        o.chunks = std::max(1, (groups + gpc - 1) / gpc);
        hcta[q] = (int)ctas;
        ctas += groups ? (long long)o.chunks * ctx->J : 0;
        words += groups;
        maxq = std::max(maxq, std::max(std::fabs(d.x - ox), std::fabs(d.y - oy)) + ctx->lib_radius);
        if (!(std::fabs(d.c) <= 1.0000001 && std::fabs(d.s) <= 1.0000001)) return fail(ctx, MPC_E_ARG, "query %d: (c, s) is not a rotation", q);
    }
    hcta[B] = (int)ctas;
    if (ctas > 0x7fffffffLL || words > 0x7fffffffLL) return fail(ctx, MPC_E_LIMIT, "batch too large");
    if (ncand) memcpy(hidx, cand_idx, bytes_i);
    CU(ensure(ctx->q_d, bytes_q));
    CU(ensure(ctx->cta_off_d, bytes_c));
    // candidate arena = staged sets followed by this call's explicit lists
    size_t arena = sizeof(int) * (size_t)std::max(ctx->set_total + ncand, 1);
    void *arena_before = ctx->cand_d.p;
    CU(ensure(ctx->cand_d, arena));
    if ((ctx->cand_d.p != arena_before || !ctx->arena_valid) && ctx->set_total)
        CU(cudaMemcpyAsync(ctx->cand_d.p, ctx->sets.p, sizeof(int) * (size_t)ctx->set_total, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->arena_valid = true;
    CU(ensure(ctx->mask_d, sizeof(unsigned) * (size_t)std::max<long long>(words, 1)));
    CU(cudaMemcpyAsync(ctx->q_d.p, hq, bytes_q, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->cta_off_d.p, hcta, bytes_c, cudaMemcpyHostToDevice, ctx->stream));
    if (ncand) CU(cudaMemcpyAsync((int *)ctx->cand_d.p + ctx->set_total, hidx, bytes_i, cudaMemcpyHostToDevice, ctx->stream));
    ctx->B = B; ctx->ctas = (int)ctas; ctx->mask_words = (int)words; ctx->mode = mode;
    ctx->maxabs_query = std::nextafter((float)maxq, INFINITY);
    ctx->prepared = true;
    return MPC_OK;
}

static void fill_stats(mpc_ctx *ctx, mpc_stats *st, float ms, int kernels, int overflow) {
    if (!st) return;
    const Counters &c = *ctx->h_cnt;
    st->pairs_nominal = (int64_t)c.pairs_nominal;
    st->refined_fp64 = (int64_t)c.refined;
    st->exact_ties = (int64_t)c.exact_ties;
    st->near_tangent = (int64_t)c.near_tangent;
    st->parity_refined = (int64_t)c.parity_refined;
    st->cull_tests = (int64_t)c.cull_tests;
    st->axis_tests = (int64_t)c.axis_tests;
    st->worklist_overflow = overflow;
    st->ms_device = ms;
    st->ctas = ctx->ctas;
    st->kernels = kernels;
    float m = 0.f;
    cudaMemcpyAsync(&m, ctx->maxabs.p, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    st->tau = std::max(std::max(m, ctx->maxabs_query) / 131072.0f, 1e-7f);
}

// run the prepared batch once, growing the refinement list if it overflowed; leaves counters in h_cnt
static int run_once(mpc_ctx *ctx, float *ms, int *kernels, int *overflow) {
    *overflow = 0;
    *kernels = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        int rc = launch_all(ctx, ctx->ev[0], nullptr, nullptr);
        if (rc) return rc;
        *kernels += 2;
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_cnt, ctx->cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
        if (ctx->h_cnt->wl_count <= (unsigned)ctx->wl_cap) return MPC_OK;
        *overflow = 1;
        ctx->wl_cap = (int)std::min<unsigned long long>((unsigned long long)ctx->h_cnt->wl_count * 5 / 4 + 1024, 0x7fffffffULL / sizeof(int4));
        CU(ensure(ctx->wl, sizeof(int4) * (size_t)ctx->wl_cap));
    }
    return fail(ctx, MPC_E_LIMIT, "refinement list still overflows at %d entries", ctx->wl_cap);
}

extern "C" int mpc_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                         unsigned mode, uint32_t *out_mask, int mask_words, mpc_stats *stats) {
    int rc = mpc_prepare_query(ctx, B, queries, cand_idx, ncand, mode);
    if (rc) return rc;
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words > 0 && !out_mask) return fail(ctx, MPC_E_ARG, "out_mask missing");
    float ms = 0.f;
    int kernels = 0, overflow = 0;
    rc = run_once(ctx, &ms, &kernels, &overflow);
    if (rc) return rc;
    if (mask_words) {
        CU(cudaMemcpyAsync(out_mask, ctx->mask_d.p, sizeof(unsigned) * (size_t)mask_words, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

extern "C" int mpc_run_prepared(mpc_ctx *ctx, int iters, int flush_l2, float *ms_total, float *ms_main, mpc_stats *stats) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    CU(cudaSetDevice(ctx->device));
    float ms = 0.f;
    int kernels = 0, overflow = 0;
    int rc = run_once(ctx, &ms, &kernels, &overflow);     // sizes the refinement list, warms nothing else up
    if (rc) return rc;
    const size_t flush_bytes = 256u << 20;
    if (flush_l2) CU(ensure(ctx->flush, flush_bytes));
    for (int it = 0; it < iters; it++) {
        if (flush_l2) {
            k_fill_u32<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>((unsigned *)ctx->flush.p, flush_bytes / 4, (unsigned)it);
            CU(cudaGetLastError());
        }
        rc = launch_all(ctx, ctx->ev[0], ctx->ev[1], ctx->ev[2]);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[3]));
        CU(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
        if (ms_total) ms_total[it] = a;
        if (ms_main) ms_main[it] = b;
        ms = a;
        kernels += 2;
    }
    CU(cudaMemcpyAsync(ctx->h_cnt, ctx->cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

extern "C" int mpc_fetch_mask(mpc_ctx *ctx, uint32_t *out_mask, int mask_words) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words) {
        CU(cudaMemcpyAsync(out_mask, ctx->mask_d.p, sizeof(unsigned) * (size_t)mask_words, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return MPC_OK;
}
