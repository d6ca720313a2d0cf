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
//   k_stage_obst / k_stage_segs   float64 global-frame inputs -> scene-local float32 SoA records (vertex pairs, unit
//                                 edge lines, quad-packed bounding circles, one box per block of 32 segments) in
//                                 32-aligned slot ranges per time step; CCW-normalised float64 copy for the tie-break.
//   k_check<VA,VB,COUNT,NOCULL>   one CTA (8 warps) per (query, occupancy slot, chunk of candidate groups).  The
//                                 obstacle and boundary-segment records of that time step are brought into shared
//                                 memory with one TMA bulk copy per field (cp.async.bulk + mbarrier).  A warp works on
//                                 32 candidate primitives at a time, lane = primitive: it transforms the occupancy
//                                 polygons, parks them in shared memory and streams the step's obstacles as
//                                 warp-broadcast loads.  Axis 0 (centre line, bounding radii) runs on packed
//                                 FADD2/FFMA2, two obstacles per instruction; the pairs it cannot separate are queued
//                                 per warp and the full separating-axis test runs on 32 queued pairs at once.  A tile
//                                 ends as soon as __all_sync says every lane has collided; __ballot_sync assembles the
//                                 32 result bits that are OR-ed into the packed mask.
//   k_refine                      float64 / exact re-decision of the pairs inside the float32 error band, one warp per
//                                 pair (lanes over the orientation tests).
// mpc_query replays the sequence (H2D, clear, k_check, k_refine, D2H) as one instantiated CUDA graph per batch shape.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
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
    int pos_base, pad0, pad1, pad2;   // >= 0: candidates are a spatially sorted set, cand_pos[pos_base + i] = original slot
};

struct Counters {
    unsigned long long pairs_nominal, refined, exact_ties, near_tangent, parity_refined, cull_tests, axis_tests;
    unsigned wl_count, pad;
    unsigned long long box_skipped, pad1[7];
};

struct KParams {
    // library
    const float4 *lib_v, *lib_n, *lib_c;       // [field][P*J]: gathered through the candidate index
    const float4 *slib_v, *slib_n, *slib_c;    // [field][J][set_total]: copy in sorted-set order, read coalesced
    int set_total;
    const double *lib_v64;
    const int *lib_nv, *slot_t;
    int PJ, J, lib_vmax;
    // scenes
    const float4 *ob_f;          // [F-1][ob_stride] vertex and edge-line fields, slot layout (steps 32-aligned)
    const float4 *ob_box;        // [ob_stride/32] bounding box of the circles of every block of 32 slots
    const float *ob_c;           // [3][ob_stride]   bounding circle x | y | r, slot layout
    const double *ob_v64;
    const int *ob_nv;
    int ob_stride, ob_vmax;
    const int *step_obst_off, *step_obst_cnt, *step_seg_begin, *step_seg_end;
    const float4 *sg_f, *sg_box;
    const double *sg_64;
    int sg_stride;
    const float *maxabs_scene;
    // queries
    const QDev *q;
    const int *cta_off, *cand, *cand_pos;
    int B, nwords;
    unsigned *mask;              // result bits, caller's candidate order
    unsigned *mask_sorted;       // result bits of spatially sorted candidate sets, un-permuted by k_refine
    int4 *wl;
    int wl_cap;
    Counters *cnt;
    const float *maxabs_query;   // batch header in the query pack (device), so that a captured graph stays valid
    float lib_rmax;
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

// packed float32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 -- two lanes of work per issued instruction)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__host__ __device__ constexpr int ob_fields(int VB) { return VB / 2 + (VB * 3 + 3) / 4; }   // float4 fields (circle apart)

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

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

__device__ __forceinline__ unsigned morton16_dev(unsigned x, unsigned y) {
    auto spread = [](unsigned v) {
        v &= 0xffff;
        v = (v | (v << 8)) & 0x00ff00ff;
        v = (v | (v << 4)) & 0x0f0f0f0f;
        v = (v | (v << 2)) & 0x33333333;
        v = (v | (v << 1)) & 0x55555555;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}

// One CTA per time step: order the step's obstacles along a Morton curve (keys from the vertex mean relative to the
// scene origin, 6 cm cells over +-2 km) with a bitonic sort in shared memory, so that a block of 32 consecutive slots
// is spatially compact and k_check can skip it by its bounding box.  order[slot] = source obstacle (or -1: padding).
// Steps with more than SORT_MAX obstacles keep their input order (the boxes are then just less selective).
constexpr int SORT_MAX = 2048;
template <int VB>
__global__ void __launch_bounds__(256) k_order_obst(const double *__restrict__ src64, const int *__restrict__ src_nv,
                                                    const int *__restrict__ src_off, const int *__restrict__ slot_off,
                                                    const int *__restrict__ scene_step_off, int nscenes,
                                                    const double *__restrict__ origins, int *__restrict__ order) {
    __shared__ unsigned s_key[SORT_MAX];
    __shared__ int s_idx[SORT_MAX];
    const int step = blockIdx.x;
    const int o0 = src_off[step], cnt = src_off[step + 1] - o0, s0 = slot_off[step], nslot = slot_off[step + 1] - s0;
    if (cnt > SORT_MAX) {
        for (int i = threadIdx.x; i < nslot; i += blockDim.x) order[s0 + i] = (i < cnt) ? o0 + i : -1;
        return;
    }
    int n2 = 32;
    while (n2 < cnt) n2 <<= 1;
    const int sc = upper_bound_idx(scene_step_off, nscenes + 1, step);
    const double ox = origins[2 * sc], oy = origins[2 * sc + 1];
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned key = 0xffffffffu;
        if (i < cnt) {
            const int nv = src_nv[o0 + i];
            const double *v = src64 + (size_t)(o0 + i) * VB * 2;
            double cx = 0.0, cy = 0.0;
            for (int k = 0; k < nv && k < VB; k++) { cx += v[2 * k]; cy += v[2 * k + 1]; }
            if (nv > 0) { cx = cx / nv - ox; cy = cy / nv - oy; }
            const double qx = fmin(fmax((cx + 2048.0) * 16.0, 0.0), 65535.0), qy = fmin(fmax((cy + 2048.0) * 16.0, 0.0), 65535.0);
            key = morton16_dev((unsigned)qx, (unsigned)qy);
            if (key == 0xffffffffu) key = 0xfffffffeu;
        }
        s_key[i] = key;
        s_idx[i] = (i < cnt) ? o0 + i : -1;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ jj;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned a = s_key[i], b = s_key[l];
                    // ties are broken by the source index so that the order is deterministic
                    const bool gt = a > b || (a == b && s_idx[i] > s_idx[l]);
                    if (gt == up) {
                        s_key[i] = b; s_key[l] = a;
                        const int t = s_idx[i]; s_idx[i] = s_idx[l]; s_idx[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < nslot; i += blockDim.x) order[s0 + i] = (i < cnt) ? s_idx[i] : -1;
}

// one thread per obstacle *slot*: the obstacles of a step occupy a 32-aligned slot range [slot_off[k], slot_off[k+1]),
// the tail is padding that can never be near anything.  Writes the float32 SoA records, the bounding circle arrays and
// the CCW-normalised float64 copy (slot layout) for the tie-break pass.
template <int VB>
__global__ void k_stage_obst(int nslots, const double *__restrict__ src64, const int *__restrict__ src_nv,
                             const int *__restrict__ order, const int *__restrict__ slot_off, int nsteps,
                             const int *__restrict__ scene_step_off, int nscenes, const double *__restrict__ origins,
                             float4 *__restrict__ ob_f, float *__restrict__ ob_c, float4 *__restrict__ ob_box,
                             double *__restrict__ ob64, int *__restrict__ ob_nv, int stride, float *maxabs) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;     // nslots is a multiple of 32: whole warps are in range
    if (slot >= nslots) return;
    constexpr int F = ob_fields(VB);
    const int step = upper_bound_idx(slot_off, nsteps + 1, slot);
    const int o = order[slot];
    int nv = (o >= 0) ? src_nv[o] : 0;
    float out[F * 4];
    float ccx = FAR, ccy = FAR, ccr = 0.f;
    double *v = ob64 + (size_t)slot * VB * 2;
    bool ok = nv >= 3 && nv <= VB;
    double x[VB], y[VB];
    if (ok) {
        const double *sv = src64 + (size_t)o * VB * 2;
        for (int k = 0; k < VB; k++) { x[k] = sv[2 * (k < nv ? k : nv - 1)]; y[k] = sv[2 * (k < nv ? k : nv - 1) + 1]; }
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
        // never near, and "always separated" for whoever tests it anyway: a point far away
        for (int k = 0; k < VB * 2; k++) out[k] = FAR;
        for (int k = 0; k < VB; k++) { out[VB * 2 + 3 * k] = 0.f; out[VB * 2 + 3 * k + 1] = 0.f; out[VB * 2 + 3 * k + 2] = -BIG; }
        for (int k = VB * 5; k < F * 4; k++) out[k] = 0.f;
        for (int k = 0; k < VB * 2; k++) v[k] = 0.0;
        nv = 0;
    } else {
        int sc = upper_bound_idx(scene_step_off, nscenes + 1, step);
        double ox = origins[2 * sc], oy = origins[2 * sc + 1];
        for (int k = 0; k < VB; k++) { v[2 * k] = x[k]; v[2 * k + 1] = y[k]; }
        double cx = 0.0, cy = 0.0, m = 0.0;
        for (int k = 0; k < VB; k++) {
            double lx = x[k] - ox, ly = y[k] - oy;
            out[4 * (k >> 1) + (k & 1)] = (float)lx; out[4 * (k >> 1) + 2 + (k & 1)] = (float)ly;   // (x0,x1,y0,y1)
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
        for (int k = VB * 5; k < F * 4; k++) out[k] = 0.f;
        ccx = (float)cx; ccy = (float)cy;
        ccr = __double2float_ru(r * (1.0 + 1e-6) + m * 2.4e-7);
        atomic_max_pos_float(maxabs, __double2float_ru(m));
    }
    ob_nv[slot] = nv;
#pragma unroll
    for (int f = 0; f < F; f++)
        ob_f[(size_t)f * stride + slot] = make_float4(out[4 * f], out[4 * f + 1], out[4 * f + 2], out[4 * f + 3]);
    ob_c[slot] = ccx;
    ob_c[(size_t)stride + slot] = ccy;
    ob_c[(size_t)2 * stride + slot] = ccr;
    // bounding box of the block's circles (of the float32 values the main kernel will see)
    const bool real = nv > 0;
    const float x0 = warp_min(real ? ccx - ccr : BIG), x1 = warp_max(real ? ccx + ccr : -BIG);
    const float y0 = warp_min(real ? ccy - ccr : BIG), y1 = warp_max(real ? ccy + ccr : -BIG);
    if ((threadIdx.x & 31) == 0)
        ob_box[slot >> 5] = make_float4(__fadd_rd(x0, -1e-6f * fabsf(x0)), __fadd_rd(y0, -1e-6f * fabsf(y0)),
                                        __fadd_ru(x1, 1e-6f * fabsf(x1)), __fadd_ru(y1, 1e-6f * fabsf(y1)));
}

// one thread per segment slot (slots are laid out in 32-aligned, spatially sorted ranges by the host); a warp = one
// block of 32 slots and also reduces the block's bounding box (of the float32 data the main kernel will see)
__global__ void k_stage_segs(int nslots, const double *__restrict__ sg64, const int *__restrict__ seg_scene,
                             const double *__restrict__ origins, float4 *__restrict__ sg_f, int stride,
                             float4 *__restrict__ sg_box, float *maxabs) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    float4 rec = make_float4(FAR, FAR, FAR, FAR), line = make_float4(0.f, 0.f, BIG, -1.f);
    if (s < nslots) {
        int sc = seg_scene[s];
        double ax = sg64[4 * s], ay = sg64[4 * s + 1], bx = sg64[4 * s + 2], by = sg64[4 * s + 3];
        double ex = bx - ax, ey = by - ay;
        double len = sqrt(ex * ex + ey * ey);
        if (sc >= 0 && len > 0.0) {       // else: padding or degenerate: far away, never near, never straddling
            double ox = origins[2 * sc], oy = origins[2 * sc + 1];
            double lax = ax - ox, lay = ay - oy, lbx = bx - ox, lby = by - oy;
            double nx = ey / len, ny = -ex / len;      // right-hand normal of a->b
            rec = make_float4((float)lax, (float)lay, (float)lbx, (float)lby);
            line = make_float4((float)nx, (float)ny, (float)(nx * lax + ny * lay), __double2float_ru(0.5 * len * (1.0 + 1e-6)));
            double m = fmax(fmax(fabs(lax), fabs(lay)), fmax(fabs(lbx), fabs(lby)));
            atomic_max_pos_float(maxabs, __double2float_ru(m));
            valid = true;
        }
        sg_f[s] = rec;
        sg_f[(size_t)stride + s] = line;
    }
    float x0 = valid ? fminf(rec.x, rec.z) : BIG, x1 = valid ? fmaxf(rec.x, rec.z) : -BIG;
    float y0 = valid ? fminf(rec.y, rec.w) : BIG, y1 = valid ? fmaxf(rec.y, rec.w) : -BIG;
    x0 = warp_min(x0); y0 = warp_min(y0); x1 = warp_max(x1); y1 = warp_max(y1);
    if ((threadIdx.x & 31) == 0 && s < nslots) sg_box[s >> 5] = make_float4(x0, y0, x1, y1);
}

// ---------------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void push_work(const KParams &p, int q, int ci, int j, int kind, int obj) {
    unsigned idx = atomicAdd(&p.cnt->wl_count, 1u);
    if (idx < (unsigned)p.wl_cap) p.wl[idx] = make_int4(q, ci, j | (kind << 16), obj);
}

// A polygon record in shared memory is SoA by float4 field: V/2 fields of vertex pairs (x0,x1,y0,y1), then
// ceil(3V/4) fields holding the packed edge lines (nx, ny, c) -- the same layout for obstacle records (staged by
// k_stage_obst) and for the transformed occupancy polygons the warps write themselves.  Vertex pairs stay packed in
// 64-bit registers so that one FFMA2 projects two vertices on an edge normal.
__device__ __forceinline__ float lo32(f32x2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int V>
struct PolyRegs {
    f32x2 X[V / 2], Y[V / 2];
    float ln[((V * 3 + 3) / 4) * 4];
    __device__ __forceinline__ void load(const float4 *rec, int stride, int i) {
#pragma unroll
        for (int h = 0; h < V / 2; h++) {
            float4 v = rec[h * stride + i];
            X[h] = pack2(v.x, v.y);
            Y[h] = pack2(v.z, v.w);
        }
#pragma unroll
        for (int h = 0; h < (V * 3 + 3) / 4; h++) {
            float4 v = rec[(V / 2 + h) * stride + i];
            ln[4 * h] = v.x; ln[4 * h + 1] = v.y; ln[4 * h + 2] = v.z; ln[4 * h + 3] = v.w;
        }
    }
    __device__ __forceinline__ float x(int k) const { return (k & 1) ? hi32(X[k >> 1]) : lo32(X[k >> 1]); }
    __device__ __forceinline__ float y(int k) const { return (k & 1) ? hi32(Y[k >> 1]) : lo32(Y[k >> 1]); }
};

// three-input min / max (sm_100: FMNMX3)
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// minimum over the vertices of Q of the signed distance to the line (nx, ny, c): two vertices per FFMA2, one FMNMX3
// per vertex pair
template <int VQ>
__device__ __forceinline__ float min_dist(const PolyRegs<VQ> &Q, float nx, float ny, float c) {
    const f32x2 nx2 = pack2(nx, nx), ny2 = pack2(ny, ny), nc2 = pack2(-c, -c);
    float mn = BIG;
#pragma unroll
    for (int h = 0; h < VQ / 2; h++) {
        const f32x2 d = fma2(nx2, Q.X[h], fma2(ny2, Q.Y[h], nc2));
        mn = (h == 0) ? fminf(lo32(d), hi32(d)) : min3(mn, lo32(d), hi32(d));
    }
    return mn;
}

// signed separation of two convex polygons: max over all edge axes of the minimum vertex distance (> 0 apart)
template <int VA, int VB>
__device__ __forceinline__ float sat_polys(const PolyRegs<VA> &A, const PolyRegs<VB> &B) {
    float sep = -BIG;
#pragma unroll
    for (int e = 0; e < VA; e += 2)
        sep = max3(sep, min_dist<VB>(B, A.ln[3 * e], A.ln[3 * e + 1], A.ln[3 * e + 2]),
                   min_dist<VB>(B, A.ln[3 * e + 3], A.ln[3 * e + 4], A.ln[3 * e + 5]));
#pragma unroll
    for (int e = 0; e < VB; e += 2)
        sep = max3(sep, min_dist<VA>(A, B.ln[3 * e], B.ln[3 * e + 1], B.ln[3 * e + 2]),
                   min_dist<VA>(A, B.ln[3 * e + 3], B.ln[3 * e + 4], B.ln[3 * e + 5]));
    return sep;
}

// signed separation of a convex polygon and a segment (a 2-gon): edges of A against the two end points, the
// segment's line against the vertices of A on either side
template <int VA>
__device__ __forceinline__ float sat_segment(const PolyRegs<VA> &A, const float4 sa, const float4 sl) {
    float sep = -BIG, mn = BIG, mx = -BIG;
#pragma unroll
    for (int e = 0; e < VA; e++) {
        const float nx = A.ln[3 * e], ny = A.ln[3 * e + 1], c = A.ln[3 * e + 2];
        const float d0 = fmaf(nx, sa.x, fmaf(ny, sa.y, -c)), d1 = fmaf(nx, sa.z, fmaf(ny, sa.w, -c));
        sep = fmaxf(sep, fminf(d0, d1));
        const float d = fmaf(sl.x, A.x(e), fmaf(sl.y, A.y(e), -sl.z));
        mn = fminf(mn, d);
        mx = fmaxf(mx, d);
    }
    return fmaxf(sep, fmaxf(mn, -mx));
}

constexpr int K_THREADS = 256;      // 8 warps share one staged tile
constexpr int QCAP = 64;            // per-warp queue of (lane, record) pairs that passed axis 0

template <int VA, int VB, bool COUNT, bool NOCULL>
__global__ void __launch_bounds__(K_THREADS, 4) k_check(const KParams p) {
    constexpr int F = ob_fields(VB), FA = ob_fields(VA), NW = K_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *s_coll = reinterpret_cast<unsigned *>(smem_raw + 16);
    unsigned *s_par = s_coll + GPC_MAX;
    unsigned *s_punc = s_par + GPC_MAX;
    unsigned *s_q = s_punc + GPC_MAX;                                          // [NW][QCAP]
    float4 *s_A = reinterpret_cast<float4 *>(s_q + NW * QCAP);                 // [NW][FA][32] occupancy polygons
    float4 *s_ob = s_A + NW * FA * 32;                                         // [F][cap_o] obstacle records
    float *s_c = reinterpret_cast<float *>(s_ob + F * p.cap_o);                // circle x | y | -(r + rcull)^2
    float4 *s_sg = reinterpret_cast<float4 *>(s_c + 3 * p.cap_o);              // [2][cap_s] segments

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // which (query, slot, chunk) is this CTA
    // (largest q with cta_off[q] <= blockIdx.x; every warp finds it with warp-wide loads and votes instead of a
    // chain of dependent loads: a 32-way search per round)
    int q = 0;
    {
        int lo = 0, hi = p.B + 1;                 // answer in [lo, hi)
        while (hi - lo > 1) {
            const int span = hi - lo, stride = (span + 31) >> 5;
            const int idx = lo + lane * stride;
            const bool le = idx < hi && __ldg(p.cta_off + idx) <= (int)blockIdx.x;
            const int k = __popc(__ballot_sync(FULL, le));       // cta_off is sorted: the first k probes are <=
            const int nlo = lo + (k - 1) * stride;
            hi = min(hi, nlo + stride);
            lo = nlo;
        }
        q = lo;
    }
    const QDev *Q = p.q + q;
    int local = blockIdx.x - p.cta_off[q];
    int chunks = Q->chunks;
    int j = local / chunks, chunk = local - j * chunks;
    int t = Q->t0 + p.slot_t[j];
    // reference guards: ind + time <= param['steps'] and ind + time < len(space)  (lowLevelPlannerManeuverAutomaton.py:120,122)
    if (t > Q->step_limit || t < 0 || t >= Q->nsteps) return;
    int step = Q->step0 + t;
    // obstacles of the step: a 32-aligned slot range (padding slots are far away), `nobst` real ones
    const int o0 = p.step_obst_off[step], nslot = p.step_obst_off[step + 1] - o0, nobst = p.step_obst_cnt[step];
    const int g0 = p.step_seg_begin[step];
    const bool has_region = g0 >= 0;
    const int nseg = has_region ? p.step_seg_end[step] - g0 : 0;
    const int cand_cnt = Q->cand_cnt, cand_base = Q->cand_base;
    const int gbeg = chunk * p.gpc;
    const int ng = min(p.gpc, (cand_cnt + 31) / 32 - gbeg);
    const float lx = Q->lx, ly = Q->ly, lc = Q->lc, ls = Q->ls;
    // float32 error band: every float32 signed distance below is within tau of its float64 value
    const float tau = fmaxf(fmaxf(*p.maxabs_scene, *p.maxabs_query) * (1.0f / 131072.0f), 1e-7f);
    const bool noexit = p.mode & MPC_MODE_NO_EARLY_EXIT;

    if (tid == 0) mbar_init(bar, 1);
    for (int i = tid; i < 3 * GPC_MAX; i += K_THREADS) s_coll[i] = 0;
    __syncthreads();

    const int nph_o = (nslot + p.cap_o - 1) / p.cap_o, nph_s = (nseg + p.cap_s - 1) / p.cap_s;
    const int nphase = max(1, max(nph_o, nph_s));
    // cull radius shared by the whole CTA: largest library bounding radius + band (conservative for smaller polygons)
    const float rcull = p.lib_rmax + 2.0f * tau;
    unsigned n_cull = 0, n_axis = 0, n_skip = 0;
    const int pos_base = Q->pos_base;
    float4 *sA = s_A + warp * FA * 32;
    unsigned *sq = s_q + warp * QCAP;

    for (int ph = 0; ph < nphase; ph++) {
        const int oc0 = ph * p.cap_o, ocn = max(0, min(nslot - oc0, p.cap_o));       // multiple of 32
        const int sc0 = ph * p.cap_s, scn = max(0, min(nseg - sc0, p.cap_s));
        const bool loading = (ocn | scn) != 0;
        if (tid == 0 && loading) {
            mbar_expect_tx(bar, (unsigned)(ocn * (16 * F + 12) + scn * 32));
            if (ocn) {
                for (int f = 0; f < F; f++)
                    tma_bulk_g2s(s_ob + f * p.cap_o, p.ob_f + (size_t)f * p.ob_stride + o0 + oc0, ocn * 16, bar);
                for (int f = 0; f < 3; f++)
                    tma_bulk_g2s(s_c + f * p.cap_o, p.ob_c + (size_t)f * p.ob_stride + o0 + oc0, ocn * 4, bar);
            }
            if (scn)
                for (int f = 0; f < 2; f++)
                    tma_bulk_g2s(s_sg + f * p.cap_s, p.sg_f + (size_t)f * p.sg_stride + g0 + sc0, scn * 16, bar);
        }
        if (loading) {
            mbar_wait(bar, ph & 1);
            // third circle array: r  ->  -(r + rcull)^2, the constant term of the centre-line test
            for (int i = tid; i < ocn; i += K_THREADS) {
                float r = s_c[2 * p.cap_o + i] + rcull;
                s_c[2 * p.cap_o + i] = -(r * r);
            }
            __syncthreads();
        }
        // candidate index of the warp's first group; the next group's is fetched
        // while the current one is processed, which takes one link out of the dependent-load chain of the set-up
        int nxt_prim = 0;
        {
            const int ci0 = (gbeg + warp) * 32 + lane;
            if (pos_base < 0 && warp < ng && ci0 < cand_cnt) nxt_prim = __ldg(p.cand + cand_base + ci0);
        }
        for (int g = warp; g < ng; g += NW) {
            // ---- lane = one candidate primitive: transform its occupancy polygon, park it in shared memory ----------
            const int ci = (gbeg + g) * 32 + lane;
            bool active = ci < cand_cnt;
            // library record of the lane's candidate: sorted sets have their own copy of the library in set order
            // (consecutive lanes read consecutive 16-byte records); other candidates are gathered through their index
            const bool direct = pos_base >= 0;
            const float4 *Lv = direct ? p.slib_v : p.lib_v, *Ln = direct ? p.slib_n : p.lib_n;
            const size_t fstride = direct ? (size_t)p.J * p.set_total : (size_t)p.PJ;
            size_t pj = 0;
            if (active) pj = direct ? (size_t)j * p.set_total + pos_base + ci : (size_t)nxt_prim * p.J + j;
            if (!direct) {
                const int cin = ci + NW * 32;
                if (g + NW < ng && cin < cand_cnt) nxt_prim = __ldg(p.cand + cand_base + cin);
            }
            float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) cc = __ldg((direct ? p.slib_c : p.lib_c) + pj);
            const int nv = (int)cc.w;
            active = active && nv > 0;
            {
                float ax[VA], ay[VA], ln[((VA * 3 + 3) / 4) * 4];
#pragma unroll
                for (int i = 0; i < ((VA * 3 + 3) / 4) * 4; i++) ln[i] = 0.f;
#pragma unroll
                for (int h = 0; h < VA / 2; h++) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
                    if (active) { v = __ldg(Lv + h * fstride + pj); n = __ldg(Ln + h * fstride + pj); }
                    ax[2 * h] = fmaf(lc, v.x, fmaf(-ls, v.y, lx));
                    ay[2 * h] = fmaf(ls, v.x, fmaf(lc, v.y, ly));
                    ax[2 * h + 1] = fmaf(lc, v.z, fmaf(-ls, v.w, lx));
                    ay[2 * h + 1] = fmaf(ls, v.z, fmaf(lc, v.w, ly));
                    ln[6 * h] = fmaf(lc, n.x, -ls * n.y);
                    ln[6 * h + 1] = fmaf(ls, n.x, lc * n.y);
                    ln[6 * h + 3] = fmaf(lc, n.z, -ls * n.w);
                    ln[6 * h + 4] = fmaf(ls, n.z, lc * n.w);
                }
#pragma unroll
                for (int e = 0; e < VA; e++)       // line offset; an edge beyond nv never separates
                    ln[3 * e + 2] = (e < nv) ? fmaf(ln[3 * e], ax[e], ln[3 * e + 1] * ay[e]) : BIG;
                __syncwarp();                       // previous group's readers are done with sA
#pragma unroll
                for (int h = 0; h < VA / 2; h++) sA[h * 32 + lane] = make_float4(ax[2 * h], ax[2 * h + 1], ay[2 * h], ay[2 * h + 1]);
#pragma unroll
                for (int h = 0; h < (VA * 3 + 3) / 4; h++)
                    sA[(VA / 2 + h) * 32 + lane] = make_float4(ln[4 * h], ln[4 * h + 1], ln[4 * h + 2], ln[4 * h + 3]);
                __syncwarp();
            }
            const float acx = active ? fmaf(lc, cc.x, fmaf(-ls, cc.y, lx)) : -FAR;
            const float acy = active ? fmaf(ls, cc.x, fmaf(lc, cc.y, ly)) : -FAR;
            const float arr = cc.z + 2.0f * tau;

            bool collide = (s_coll[g] >> lane) & 1u;
            bool par = false, punc = false;
            unsigned qn = 0;                        // entries in this warp's queue (warp-uniform)
            // bounding box of the warp's 32 polygon centres: decides which 32-blocks of the spatially sorted obstacle
            // slots / boundary segments can matter for anyone in the warp
            const bool cull_o = !NOCULL && nslot > 32, cull_s = !NOCULL && nseg > 32;
            float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f, wr = 0.f;
            if (cull_o || cull_s) {
                wx0 = warp_min(active ? acx : BIG); wx1 = warp_max(active ? acx : -BIG);
                wy0 = warp_min(active ? acy : BIG); wy1 = warp_max(active ? acy : -BIG);
                wr = warp_max(active ? arr : 0.f);
            }

            if (ph == 0) {   // statistic: nominal pair count of this tile
                unsigned w = __reduce_add_sync(FULL, active ? (unsigned)(nobst + nseg) : 0u);
                if (lane == 0 && w) atomicAdd(&p.cnt->pairs_nominal, (unsigned long long)w);
            }

            // queue entry = (owner lane << 16) | record index in the tile.  Draining evaluates up to 32 queued pairs at
            // once, one per lane, whichever lane they belong to: the full separating-axis test runs with all lanes busy
            // although only ~1 % of the pairs reach it.
            auto push_hits = [&](unsigned &hits, int base, auto &&drain) {
                while (__any_sync(FULL, hits != 0)) {
                    if (qn >= 32) drain();
                    const bool has = hits != 0;
                    const unsigned bal = __ballot_sync(FULL, has);
                    if (has) {
                        const int k = __clz(hits);
                        hits &= ~(0x80000000u >> k);
                        sq[qn + __popc(bal & lt_mask)] = ((unsigned)lane << 16) | (unsigned)(base + k);
                    }
                    qn += __popc(bal);
                    __syncwarp();
                    if (collide && !noexit) hits = 0;
                }
            };
            auto pop_batch = [&](unsigned &ent) -> bool {       // take min(32, qn) entries, compact the rest
                const int n = min(32u, qn);
                const bool have = lane < n;
                ent = have ? sq[lane] : 0u;
                unsigned rest = (lane + 32 < (int)qn) ? sq[lane + 32] : 0u;
                __syncwarp();
                if (lane + 32 < (int)qn) sq[lane] = rest;
                __syncwarp();
                qn -= n;
                return have;
            };
            auto drain_obst = [&]() {
                unsigned ent;
                const bool have = pop_batch(ent);
                const int src = ent >> 16, slot = ent & 0xffff;
                float sep = BIG;
                if (have) {
                    PolyRegs<VA> A;
                    PolyRegs<VB> B;
                    A.load(sA, 32, src);
                    B.load(s_ob, p.cap_o, slot);
                    sep = sat_polys<VA, VB>(A, B);
                    if (COUNT) n_axis++;
                }
                const bool hit = have && sep < -tau;
                if (have && !hit && !(sep > tau)) push_work(p, q, (gbeg + g) * 32 + src, j, KIND_OBST, o0 + oc0 + slot);
                const unsigned m = __reduce_or_sync(FULL, hit ? (1u << src) : 0u);
                collide |= (m >> lane) & 1u;
            };
            auto drain_seg = [&]() {
                unsigned ent;
                const bool have = pop_batch(ent);
                const int src = ent >> 16, k = ent & 0xffff;
                float sep = BIG;
                if (have) {
                    PolyRegs<VA> A;
                    A.load(sA, 32, src);
                    sep = sat_segment<VA>(A, s_sg[k], s_sg[p.cap_s + k]);
                    if (COUNT) n_axis++;
                }
                const bool hit = have && sep < -tau;
                if (have && !hit && !(sep > tau)) push_work(p, q, (gbeg + g) * 32 + src, j, KIND_SEG, g0 + sc0 + k);
                const unsigned m = __reduce_or_sync(FULL, hit ? (1u << src) : 0u);
                collide |= (m >> lane) & 1u;
            };

            // ---- obstacles of this time step --------------------------------------------------------------------
            if (NOCULL) {
                // full evaluation: every lane runs the complete test against every obstacle of the step
                PolyRegs<VA> A;
                A.load(sA, 32, lane);
                for (int k = 0; k < min(ocn, nobst - oc0); k++) {
                    PolyRegs<VB> B;
                    B.load(s_ob, p.cap_o, k);
                    const float sep = sat_polys<VA, VB>(A, B);
                    if (active) {
                        if (COUNT) n_axis++;
                        if (sep < -tau) collide = true;
                        else if (!(sep > tau)) push_work(p, q, ci, j, KIND_OBST, o0 + oc0 + k);
                    }
                    if (!noexit && (k & 31) == 31 && __all_sync(FULL, collide || !active)) break;
                }
            } else {
                const float4 *s_cx4 = reinterpret_cast<const float4 *>(s_c);
                const float4 *s_cy4 = reinterpret_cast<const float4 *>(s_c + p.cap_o);
                const float4 *s_cr4 = reinterpret_cast<const float4 *>(s_c + 2 * p.cap_o);
                const f32x2 nax = pack2(-acx, -acx), nay = pack2(-acy, -acy);
                // which blocks of 32 slots can matter: lane b tests block b (one load latency for the whole tile).  A
                // block whose circles (box already inflated by their radii) cannot come within rcull of any centre of
                // this warp has every pair separated on the box axis.
                unsigned live_o = FULL;
                if (cull_o) {
                    const int nblk = ocn >> 5;               // <= 10 (cap_o <= 320)
                    bool live = false;
                    if (lane < nblk) {
                        const float4 bb = __ldg(p.ob_box + ((o0 + oc0) >> 5) + lane);
                        live = !(bb.x > wx1 + rcull || bb.z < wx0 - rcull || bb.y > wy1 + rcull || bb.w < wy0 - rcull);
                    }
                    live_o = __ballot_sync(FULL, live);
                }
                for (int ob = 0; ob < ocn; ob += 32) {
                    if (!((live_o >> (ob >> 5)) & 1u)) {
                        if (COUNT && active) n_skip += min(32, nobst - oc0 - ob);
                        continue;
                    }
                    // axis 0: centre line with bounding radii.  Three warp-broadcast LDS.128 bring four obstacles; packed
                    // FADD2/FFMA2 evaluate e = |dc|^2 - (rb + rcull)^2 for two of them per instruction and the sign bit
                    // of e is shifted into the hit word (bit 31-k <-> slot ob+k).
                    unsigned hits = 0;
#pragma unroll
                    for (int qd = 0; qd < 8; qd++) {
                        const float4 X = s_cx4[(ob >> 2) + qd], Y = s_cy4[(ob >> 2) + qd], R = s_cr4[(ob >> 2) + qd];
                        const f32x2 dx01 = add2(pack2(X.x, X.y), nax), dy01 = add2(pack2(Y.x, Y.y), nay);
                        const f32x2 dx23 = add2(pack2(X.z, X.w), nax), dy23 = add2(pack2(Y.z, Y.w), nay);
                        const f32x2 e01 = fma2(dx01, dx01, fma2(dy01, dy01, pack2(R.x, R.y)));
                        const f32x2 e23 = fma2(dx23, dx23, fma2(dy23, dy23, pack2(R.z, R.w)));
                        hits = __funnelshift_l((unsigned)e01, hits, 1);
                        hits = __funnelshift_l((unsigned)(e01 >> 32), hits, 1);
                        hits = __funnelshift_l((unsigned)e23, hits, 1);
                        hits = __funnelshift_l((unsigned)(e23 >> 32), hits, 1);
                    }
                    if (!active || (collide && !noexit)) hits = 0;
                    if (COUNT && active) n_cull += min(32, nobst - oc0 - ob);
                    push_hits(hits, ob, drain_obst);
                    if (qn >= 32) {
                        drain_obst();
                        if (!noexit && __all_sync(FULL, collide || !active)) { qn = 0; break; }   // warp-wide early exit
                    }
                }
                while (qn) drain_obst();
            }

            // ---- boundary segments of the free space at this time step -----------------------------------------
            if (scn && (noexit || !__all_sync(FULL, collide || !active))) {
                const bool blockcull = cull_s;        // a single block cannot be skipped: no box work
                const float4 *boxes = p.sg_box + ((g0 + sc0) >> 5);
                // the even-odd ray may run along +x, -x, +y or -y; take the direction whose ray corridor meets the
                // fewest blocks of the whole range (a road aligned with the ray would otherwise keep every block alive)
                int dir = 0;
                if (blockcull) {
                    const float4 *all = p.sg_box + (g0 >> 5);
                    const int nblk = (nseg + 31) >> 5;
                    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                    for (int b = lane; b < ((nblk + 31) & ~31); b += 32) {
                        float4 bb = (b < nblk) ? __ldg(all + b) : make_float4(BIG, BIG, -BIG, -BIG);
                        const bool iny = bb.w >= wy0 && bb.y <= wy1, inx = bb.z >= wx0 && bb.x <= wx1;
                        c0 += __popc(__ballot_sync(FULL, iny && bb.z >= wx0));      // +x
                        c1 += __popc(__ballot_sync(FULL, iny && bb.x <= wx1));      // -x
                        c2 += __popc(__ballot_sync(FULL, inx && bb.w >= wy0));      // +y
                        c3 += __popc(__ballot_sync(FULL, inx && bb.y <= wy1));      // -y
                    }
                    int best = c0;
                    if (c1 < best) { best = c1; dir = 1; }
                    if (c2 < best) { best = c2; dir = 2; }
                    if (c3 < best) { best = c3; dir = 3; }
                }
                const bool ray_y = dir >= 2, flip = (dir == 1) || (dir == 2);
                const float pc = ray_y ? acx : acy;          // coordinate the ray keeps constant
                // live blocks of this phase (<= 16: cap_s <= 512), lane b tests block b
                unsigned live_s = FULL;
                if (blockcull) {
                    bool live = false;
                    if (lane < ((scn + 31) >> 5)) {
                        const float4 bb = __ldg(boxes + lane);
                        // a segment can meet a polygon only if the boxes overlap; it can be crossed by the ray of a
                        // centre only if it reaches that centre's constant coordinate and is not entirely behind it
                        const bool sat_rel = bb.x <= wx1 + wr && bb.z >= wx0 - wr && bb.y <= wy1 + wr && bb.w >= wy0 - wr;
                        const bool iny = bb.w >= wy0 && bb.y <= wy1, inx = bb.z >= wx0 && bb.x <= wx1;
                        const bool par_rel = dir == 0 ? (iny && bb.z >= wx0) : dir == 1 ? (iny && bb.x <= wx1)
                                           : dir == 2 ? (inx && bb.w >= wy0) : (inx && bb.y <= wy1);
                        live = sat_rel || par_rel;
                    }
                    live_s = __ballot_sync(FULL, live);
                }
                for (int sb = 0; sb < scn; sb += 32) {
                    const int cnt = min(32, scn - sb);
                    if (!((live_s >> (sb >> 5)) & 1u)) continue;
                    unsigned hits = 0;
#pragma unroll 2
                    for (int k = 0; k < cnt; k++) {
                        float4 sa = s_sg[sb + k], sl = s_sg[p.cap_s + sb + k];
                        float dl = fmaf(sl.x, acx, fmaf(sl.y, acy, -sl.z));       // centre to the segment's line
                        // even-odd crossing parity of the ray from the polygon centre.  The two comparisons are exact
                        // on the float32 data and consistent between segments that share a vertex; the side test is
                        // trusted only outside the error band
                        const float ea = ray_y ? sa.x : sa.y, eb = ray_y ? sa.z : sa.w;
                        bool ua = ea > pc, ub = eb > pc;
                        if (ua != ub) {
                            par ^= (ub != flip) ? (dl < 0.f) : (dl > 0.f);
                            punc |= fabsf(dl) <= tau;
                        }
                        float tt = fmaf(-sl.y, acx - 0.5f * (sa.x + sa.z), sl.x * (acy - 0.5f * (sa.y + sa.w)));
                        bool near = NOCULL ? (sl.w >= 0.f) : ((fabsf(dl) <= arr) & (fabsf(tt) <= sl.w + arr));
                        if (near) hits |= 0x80000000u >> k;
                    }
                    if (!active || (collide && !noexit)) hits = 0;
                    if (COUNT && active) n_cull += cnt;
                    push_hits(hits, sb, drain_seg);
                    if (qn >= 32) drain_seg();
                }
                while (qn) drain_seg();
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
                if (pos_base < 0) {
                    unsigned word = __ballot_sync(FULL, coll && active);
                    if (lane == 0 && word) atomicOr(p.mask + Q->mask_off + gbeg + g, word);
                } else {        // spatially sorted candidate set: bits in sorted order, k_refine moves them home
                    unsigned word = __ballot_sync(FULL, coll && active);
                    if (lane == 0 && word) atomicOr(p.mask_sorted + Q->mask_off + gbeg + g, word);
                }
            }
        }
        if (nphase > 1) __syncthreads();      // everyone is done with the tile before the next bulk copy overwrites it
    }
    if (COUNT) {
        n_cull = __reduce_add_sync(FULL, n_cull);
        n_axis = __reduce_add_sync(FULL, n_axis);
        n_skip = __reduce_add_sync(FULL, n_skip);
        if (lane == 0) {
            atomicAdd(&p.cnt->cull_tests, (unsigned long long)n_cull);
            atomicAdd(&p.cnt->axis_tests, (unsigned long long)n_axis);
            atomicAdd(&p.cnt->box_skipped, (unsigned long long)n_skip);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// float64 / exact refinement of the pairs inside the float32 band
// ---------------------------------------------------------------------------------------------------------------------
// One WARP per work-list item: the orientation tests of an item (up to nA*nB of them per direction) are independent,
// so the lanes split them and votes put the answer together.  A single thread needs ~20 k cycles per polygon pair
// (dependent float64 chains, local-memory polygon arrays); a warp needs ~1-2 k, which is what a planner that waits on
// every query cares about.
__global__ void __launch_bounds__(128) k_refine(const KParams p) {
    __shared__ double s_pts[4][2 * (MPC_MAX_LIB_VERTS + MPC_MAX_OBST_VERTS)];
    const unsigned n = min(p.cnt->wl_count, (unsigned)p.wl_cap);
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    double *Ax = s_pts[wib], *Ay = Ax + MPC_MAX_LIB_VERTS, *Bx = Ay + MPC_MAX_LIB_VERTS, *By = Bx + MPC_MAX_OBST_VERTS;
    unsigned refined = 0, ties = 0, near_t = 0, par_ref = 0;
    const bool strict = p.mode & MPC_MODE_VALIDATOR;
    // phase A: result bits of spatially sorted candidate sets go home to the caller's candidate order: one thread per
    // mask word (its query found by bisection on the mask offsets), only set bits cost an atomic
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < p.nwords; w += gridDim.x * blockDim.x) {
        unsigned bits = p.mask_sorted[w];
        if (!bits) continue;
        int lo = 0, hi = p.B;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.q[mid].mask_off <= w) lo = mid; else hi = mid;
        }
        while (lo + 1 < p.B && p.q[lo + 1].mask_off <= w) lo++;          // queries without candidates share an offset
        const QDev *Q = p.q + lo;
        const int base = Q->pos_base + (w - Q->mask_off) * 32;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int oi = __ldg(p.cand_pos + base + b);
            atomicOr(p.mask + Q->mask_off + (oi >> 5), 1u << (oi & 31));
        }
    }
    for (unsigned item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n; item += warps) {
        const int4 e = p.wl[item];
        const int q = e.x, ci = e.y, j = e.z & 0xffff, kind = e.z >> 16, obj = e.w;
        const QDev *Q = p.q + q;
        const int pj = p.cand[Q->cand_base + ci] * p.J + j;
        const int nA = p.lib_nv[pj];
        __syncwarp();
        if ((int)lane < nA) {
            const double *lv = p.lib_v64 + ((size_t)pj * p.lib_vmax + lane) * 2;
            // shapely affine_transform(o['space'], [cos, -sin, sin, cos, x, y]): xp = a*x + b*y + xoff, left to right
            Ax[lane] = __dadd_rn(__dadd_rn(__dmul_rn(Q->c, lv[0]), __dmul_rn(-Q->s, lv[1])), Q->x);
            Ay[lane] = __dadd_rn(__dadd_rn(__dmul_rn(Q->s, lv[0]), __dmul_rn(Q->c, lv[1])), Q->y);
        }
        bool collide = false;
        if (kind == KIND_OBST) {
            const int nB = p.ob_nv[obj];
            if ((int)lane < nB) {
                const double *bv = p.ob_v64 + ((size_t)obj * p.ob_vmax + lane) * 2;
                Bx[lane] = bv[0]; By[lane] = bv[1];
            }
            __syncwarp();
            // an edge separates iff none of the other polygon's vertices is on its inner side (strict: or on it)
            unsigned badA = 0, badB = 0;       // bit e set: edge e of A (of B) does not separate
            for (int c = lane; c < ((nA * nB + 31) & ~31); c += 32) {
                bool bad = false, bad2 = false;
                int ea = 0, eb = 0;
                if (c < nA * nB) {
                    ea = c / nB;
                    const int k = c - ea * nB, e1 = (ea + 1 == nA) ? 0 : ea + 1;
                    if (Ax[ea] == Ax[e1] && Ay[ea] == Ay[e1]) bad = true;
                    else {
                        const int o = orient2d(Ax[ea], Ay[ea], Ax[e1], Ay[e1], Bx[k], By[k], ties);
                        bad = strict ? (o >= 0) : (o > 0);
                    }
                    eb = c / nA;
                    const int k2 = c - eb * nA, f1 = (eb + 1 == nB) ? 0 : eb + 1;
                    if (Bx[eb] == Bx[f1] && By[eb] == By[f1]) bad2 = true;
                    else {
                        const int o = orient2d(Bx[eb], By[eb], Bx[f1], By[f1], Ax[k2], Ay[k2], ties);
                        bad2 = strict ? (o >= 0) : (o > 0);
                    }
                }
                badA |= __reduce_or_sync(FULL, bad ? (1u << ea) : 0u);
                badB |= __reduce_or_sync(FULL, bad2 ? (1u << eb) : 0u);
            }
            const bool sepA = (~badA & ((1u << nA) - 1u)) != 0, sepB = (~badB & ((1u << nB) - 1u)) != 0;
            collide = !(sepA || sepB);
            // near-tangent statistic: |separation| < 1e-9 m, one edge per lane
            double mine = -1e300;
            if ((int)lane < nA + nB) {
                const bool fromA = (int)lane < nA;
                const int ee = fromA ? lane : lane - nA, nn = fromA ? nA : nB, e1 = (ee + 1 == nn) ? 0 : ee + 1;
                const double *Px = fromA ? Ax : Bx, *Py = fromA ? Ay : By, *Qx = fromA ? Bx : Ax, *Qy = fromA ? By : Ay;
                const int nq = fromA ? nB : nA;
                const double ex = Px[e1] - Px[ee], ey = Py[e1] - Py[ee], len = sqrt(ex * ex + ey * ey);
                if (len > 0.0) {
                    double mn = 1e300;
                    for (int k = 0; k < nq; k++) mn = fmin(mn, ey * (Qx[k] - Px[ee]) - ex * (Qy[k] - Py[ee]));
                    mine = mn / len;
                }
            }
            for (int o = 16; o; o >>= 1) mine = fmax(mine, __shfl_xor_sync(FULL, mine, o));
            if (lane == 0) { if (fabs(mine) < 1e-9) near_t++; refined++; }
        } else if (kind == KIND_SEG) {
            __syncwarp();
            const double *sg = p.sg_64 + 4 * (size_t)obj;
            const double s0x = sg[0], s0y = sg[1], s1x = sg[2], s1y = sg[3];
            // lanes 0..nA-1: edge i of the polygon against both end points; lanes 16..16+nA-1: vertex against the line
            bool excl = false, pos = false, neg = false;
            if ((int)lane < nA) {
                const int i = lane, k = (i + 1 == nA) ? 0 : i + 1;
                if (!(Ax[i] == Ax[k] && Ay[i] == Ay[k]))
                    excl = orient2d(Ax[i], Ay[i], Ax[k], Ay[k], s0x, s0y, ties) <= 0 &&
                           orient2d(Ax[i], Ay[i], Ax[k], Ay[k], s1x, s1y, ties) <= 0;
            } else if (lane >= 16 && (int)lane - 16 < nA) {
                const int o = orient2d(s0x, s0y, s1x, s1y, Ax[lane - 16], Ay[lane - 16], ties);
                pos = o > 0; neg = o < 0;
            }
            const bool any_excl = __any_sync(FULL, excl), any_pos = __any_sync(FULL, pos), any_neg = __any_sync(FULL, neg);
            const bool point = (s0x == s1x && s0y == s1y);
            collide = !any_excl && (point || (any_pos && any_neg));
            if (lane == 0) refined++;
        } else {
            __syncwarp();
            double cx = 0.0, cy = 0.0;
            for (int k = 0; k < nA; k++) { cx += Ax[k]; cy += Ay[k]; }
            cx /= nA; cy /= nA;
            const int step = Q->step0 + Q->t0 + p.slot_t[j];
            const int g0 = p.step_seg_begin[step], g1 = p.step_seg_end[step];
            bool inside = false, on_edge = false;
            for (int g = g0 + (int)lane; g < g1; g += 32) {
                const double *sg = p.sg_64 + 4 * (size_t)g;
                const double sax = sg[0], say = sg[1], sbx = sg[2], sby = sg[3];
                if (sax == sbx && say == sby) continue;
                const int o = orient2d(sax, say, sbx, sby, cx, cy, ties);
                if (o == 0 && fmin(sax, sbx) <= cx && cx <= fmax(sax, sbx) && fmin(say, sby) <= cy && cy <= fmax(say, sby))
                    on_edge = true;
                if ((say > cy) != (sby > cy)) {
                    if ((sby > say && o > 0) || (sby < say && o < 0)) inside = !inside;
                }
            }
            const unsigned odd = __popc(__ballot_sync(FULL, inside)) & 1u;
            collide = __any_sync(FULL, on_edge) || !odd;
            if (lane == 0) par_ref++;
        }
        if (lane == 0 && collide) {
            const int oi = Q->pos_base >= 0 ? p.cand_pos[Q->pos_base + ci] : ci;
            atomicOr(p.mask + Q->mask_off + (oi >> 5), 1u << (oi & 31));
        }
    }
    refined = __reduce_add_sync(FULL, refined);
    ties = __reduce_add_sync(FULL, ties);
    near_t = __reduce_add_sync(FULL, near_t);
    par_ref = __reduce_add_sync(FULL, par_ref);
    if (lane == 0) {
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
    float lib_rmax = 0.f;
    DevBuf lib_v, lib_n, lib_c, lib_v64, lib_nv, slot_t, sets, slib_v, slib_n, slib_c;
    std::vector<int> set_off;
    int nsets = 0, set_total = 0;
    bool arena_valid = false;
    // scenes
    bool have_scenes = false;
    int nscenes = 0, nsteps = 0, nobst = 0, nseg = 0, ob_vmax = 4, VB = 4, max_obst_step = 0, max_seg_step = 0;
    std::vector<int> scene_step_off;
    std::vector<double> origins;
    DevBuf ob_src, ob_src_nv, ob_src_off, ob_c, ob_box, ob_order, step_obst_cnt, sets_pos;
    int ob_stride = 32;
    DevBuf ob_in, ob_f, ob_nv, step_obst_off, step_seg_begin, step_seg_end, scene_step_off_d, origins_d, sg_64, sg_f,
        sg_box, seg_scene, maxabs;
    int sg_stride = 32;
    float maxabs_scene = 0.f;
    DevBuf meta;                  // one allocation behind the small per-scene arrays (views below), one H2D copy
    void *h_meta = nullptr;
    size_t h_meta_cap = 0;
    long long meta_sig = 0;
    // queries
    DevBuf qpack_d, cand_d, res_d, wl, flush;
    size_t qpack_cta_off = 0, qpack_hdr_off = 0, up_bytes = 0, up_cand_bytes = 0, up_cand_src = 0;
    unsigned long long generation = 0;          // bumped whenever a pointer a captured graph holds may have changed
    std::map<std::vector<long long>, cudaGraphExec_t> graphs;
    void *h_res = nullptr;
    size_t h_res_cap = 0;
    std::map<int, size_t> smem_attr;
    int wl_cap = 0;
    void *h_pin = nullptr;
    size_t h_pin_cap = 0;
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

static void invalidate_graphs(mpc_ctx *ctx);

// grow-only device buffer; a reallocation moves pointers that captured graphs hold, so it drops them
static cudaError_t ensure(mpc_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return cudaSuccess;
    invalidate_graphs(ctx);
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 18);   // generous floor: reallocations stall the stream
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

static void invalidate_graphs(mpc_ctx *ctx) {
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->generation++;
}

static cudaError_t ensure_pinned(mpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_pin_cap) return cudaSuccess;
    invalidate_graphs(ctx);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
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
    CU(ensure(ctx, ctx->res_d, sizeof(Counters) + 4096));
    CU(ensure(ctx, ctx->maxabs, sizeof(float)));
    CU(cudaMemsetAsync(ctx->maxabs.p, 0, sizeof(float), ctx->stream));
    ctx->wl_cap = 1 << 18;
    if (const char *e = getenv("MPC_WL_CAP")) ctx->wl_cap = std::max(1, atoi(e));      // test hook: force the overflow path
    CU(ensure(ctx, ctx->wl, sizeof(int4) * (size_t)ctx->wl_cap));
    return MPC_OK;
}

extern "C" void mpc_destroy(mpc_ctx *ctx) {
    if (!ctx) return;
    if (ctx->stream) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    DevBuf *bufs[] = {&ctx->lib_v, &ctx->lib_n, &ctx->lib_c, &ctx->lib_v64, &ctx->lib_nv, &ctx->slot_t, &ctx->sets,
                      &ctx->slib_v, &ctx->slib_n, &ctx->slib_c,
                      &ctx->ob_src, &ctx->ob_src_nv, &ctx->ob_c, &ctx->ob_box, &ctx->ob_order,
                      &ctx->sets_pos,
                      &ctx->ob_in, &ctx->ob_f, &ctx->ob_nv, &ctx->meta, &ctx->sg_f,
                      &ctx->sg_box, &ctx->maxabs, &ctx->qpack_d, &ctx->cand_d, &ctx->res_d, &ctx->wl,
                      &ctx->flush};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
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

static inline unsigned morton16(unsigned x, unsigned y) {
    auto spread = [](unsigned v) {
        v &= 0xffff;
        v = (v | (v << 8)) & 0x00ff00ff;
        v = (v | (v << 4)) & 0x0f0f0f0f;
        v = (v | (v << 2)) & 0x33333333;
        v = (v | (v << 1)) & 0x55555555;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}

extern "C" int mpc_alloc_pinned(mpc_ctx *ctx, uint64_t bytes, void **out) {
    if (!ctx || !ctx->stream || !out) return MPC_E_ARG;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMallocHost(out, (size_t)std::max<uint64_t>(bytes, 16)));
    return MPC_OK;
}

extern "C" int mpc_free_pinned(mpc_ctx *ctx, void *ptr) {
    if (!ctx || !ctx->stream) return MPC_E_ARG;
    if (ptr) CU(cudaFreeHost(ptr));
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
    float rmax = 0.f;
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
        rmax = std::max(rmax, hc[i].z);
    }
    // note: a zero-length edge keeps normal (0,0); k_check turns its line into "never separating" only for e >= nv,
    // so give such edges the same treatment by dropping them here
    CU(ensure(ctx, ctx->lib_v, hv.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_n, hn.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_c, hc.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_v64, h64.size() * sizeof(double)));
    CU(ensure(ctx, ctx->lib_nv, hnv.size() * sizeof(int)));
    CU(ensure(ctx, ctx->slot_t, slot_t.size() * sizeof(int)));
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
        // every set is kept twice: in the caller's order (sub-ranges of a set) and sorted along a Morton curve of the
        // primitives' last occupancy centre (whole-set queries): neighbouring lanes then hold neighbouring polygons at
        // every time step, which is what makes the per-warp block culling of k_check selective.  pos[] maps a sorted
        // entry back to its slot in the caller's order (that is where its result bit goes).
        std::vector<int> both((size_t)2 * std::max(total, 1)), pos((size_t)std::max(total, 1));
        for (int i = 0; i < total; i++) both[i] = set_idx[i];
        for (int sidx = 0; sidx < nsets; sidx++) {
            const int b = set_off[sidx], e = set_off[sidx + 1], n = e - b;
            std::vector<std::pair<float, float>> c(n);
            float x0 = 1e30f, x1 = -1e30f, y0 = 1e30f, y1 = -1e30f;
            for (int i = 0; i < n; i++) {
                const int pr = set_idx[b + i];
                float cx = 0.f, cy = 0.f;
                for (int j = J - 1; j >= 0; j--)
                    if (hnv[(size_t)pr * J + j] > 0) { cx = hc[(size_t)pr * J + j].x; cy = hc[(size_t)pr * J + j].y; break; }
                c[i] = {cx, cy};
                x0 = std::min(x0, cx); x1 = std::max(x1, cx); y0 = std::min(y0, cy); y1 = std::max(y1, cy);
            }
            const float sx = x1 > x0 ? 65535.f / (x1 - x0) : 0.f, sy = y1 > y0 ? 65535.f / (y1 - y0) : 0.f;
            std::vector<std::pair<unsigned, int>> order(n);
            for (int i = 0; i < n; i++)
                order[i] = {morton16((unsigned)((c[i].first - x0) * sx), (unsigned)((c[i].second - y0) * sy)), i};
            std::sort(order.begin(), order.end());
            for (int i = 0; i < n; i++) {
                both[(size_t)total + b + i] = set_idx[b + order[i].second];
                pos[b + i] = order[i].second;
            }
        }
        // the library once more, in sorted-set order and slot-major, for coalesced reads by k_check
        const int H = VA / 2;
        std::vector<float4> sv((size_t)H * J * std::max(total, 1)), sn(sv.size()), scv((size_t)J * std::max(total, 1));
        for (int a = 0; a < total; a++) {
            const size_t pr = (size_t)both[(size_t)total + a];
            for (int j = 0; j < J; j++) {
                scv[(size_t)j * total + a] = hc[pr * J + j];
                for (int h = 0; h < H; h++) {
                    sv[((size_t)h * J + j) * total + a] = hv[(size_t)h * PJ + pr * J + j];
                    sn[((size_t)h * J + j) * total + a] = hn[(size_t)h * PJ + pr * J + j];
                }
            }
        }
        CU(ensure(ctx, ctx->slib_v, sv.size() * sizeof(float4)));
        CU(ensure(ctx, ctx->slib_n, sn.size() * sizeof(float4)));
        CU(ensure(ctx, ctx->slib_c, scv.size() * sizeof(float4)));
        CU(ensure(ctx, ctx->sets, both.size() * sizeof(int)));
        CU(ensure(ctx, ctx->sets_pos, pos.size() * sizeof(int)));
        if (total) {
            CU(cudaMemcpyAsync(ctx->slib_v.p, sv.data(), sv.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->slib_n.p, sn.data(), sn.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->slib_c.p, scv.data(), scv.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->sets.p, both.data(), (size_t)2 * total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->sets_pos.p, pos.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->P = P; ctx->J = J; ctx->lib_vmax = VA; ctx->VA = VA; ctx->lib_radius = radius; ctx->lib_rmax = rmax;
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

    // ---- boundary segments: every distinct [begin, end) range becomes a 32-aligned, spatially sorted (Morton order of
    // the midpoints) block range on the device, so that k_check can skip whole blocks by their bounding box
    int max_o = 0, max_s = 0;
    std::map<std::pair<int, int>, std::pair<int, int>> ranges;        // (begin, end) -> (device begin, scene)
    std::vector<int> dev_begin(std::max(nsteps, 1), -1), dev_end(std::max(nsteps, 1), -1);
    std::vector<double> dsegs;
    std::vector<int> dscene;
    for (int s = 0; s < nscenes; s++)
        for (int k = scene_step_off[s]; k < scene_step_off[s + 1]; k++) {
            int no = step_obst_off[k + 1] - step_obst_off[k];
            if (no < 0) return fail(ctx, MPC_E_ARG, "step_obst_off not monotone at step %d", k);
            max_o = std::max(max_o, no);
            int b = step_seg_begin[k], e = step_seg_end[k];
            if (b < 0) continue;
            if (e < b || e > nseg) return fail(ctx, MPC_E_ARG, "segment range of step %d outside 0..%d", k, nseg);
            max_s = std::max(max_s, e - b);
            auto key = std::make_pair(b, e);
            auto it = ranges.find(key);
            if (it != ranges.end() && it->second.second != s) it = ranges.end();   // same range, other origin: copy
            if (it == ranges.end()) {
                const int n = e - b, start = (int)dscene.size();
                std::vector<std::pair<unsigned, int>> order(n);
                double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
                for (int g = b; g < e; g++) {
                    double mx = 0.5 * (segs[4 * g] + segs[4 * g + 2]), my = 0.5 * (segs[4 * g + 1] + segs[4 * g + 3]);
                    x0 = std::min(x0, mx); x1 = std::max(x1, mx); y0 = std::min(y0, my); y1 = std::max(y1, my);
                }
                const double sx = (x1 > x0) ? 65535.0 / (x1 - x0) : 0.0, sy = (y1 > y0) ? 65535.0 / (y1 - y0) : 0.0;
                for (int g = b; g < e; g++) {
                    double mx = 0.5 * (segs[4 * g] + segs[4 * g + 2]), my = 0.5 * (segs[4 * g + 1] + segs[4 * g + 3]);
                    order[g - b] = {morton16((unsigned)((mx - x0) * sx), (unsigned)((my - y0) * sy)), g};
                }
                std::sort(order.begin(), order.end());
                const int padded = (n + 31) / 32 * 32;
                dsegs.resize((size_t)(start + padded) * 4, 0.0);
                dscene.resize((size_t)(start + padded), -1);
                for (int i = 0; i < n; i++) {
                    const int g = order[i].second;
                    for (int c = 0; c < 4; c++) dsegs[(size_t)(start + i) * 4 + c] = segs[4 * g + c];
                    dscene[start + i] = s;
                }
                it = ranges.insert_or_assign(key, std::make_pair(start, s)).first;
            }
            dev_begin[k] = it->second.first;
            dev_end[k] = it->second.first + (e - b);
        }
    const int nslots = (int)dscene.size();
    int vmax_used = 3;
    for (int o = 0; o < nobst; o++) {
        if (obst_nv[o] > Vo || obst_nv[o] < 0) return fail(ctx, MPC_E_ARG, "obst_nv[%d]=%d outside 0..%d", o, obst_nv[o], Vo);
        vmax_used = std::max(vmax_used, (int)obst_nv[o]);
    }
    const int VB = vmax_used <= 4 ? 4 : 8;
    const int F = ob_fields(VB);
    // obstacle slots: the obstacles of step k live in the 32-aligned range [slot_off[k], slot_off[k+1])
    std::vector<int> slot_off(nsteps + 1, 0), step_cnt(std::max(nsteps, 1), 0);
    for (int k = 0; k < nsteps; k++) {
        step_cnt[k] = step_obst_off[k + 1] - step_obst_off[k];
        slot_off[k + 1] = slot_off[k] + (step_cnt[k] + 31) / 32 * 32;
    }
    const int noslots = slot_off[nsteps];
    const int ob_stride = std::max(noslots, 32), sg_stride = std::max(nslots, 32);
    CU(ensure(ctx, ctx->ob_src, (size_t)std::max(nobst, 1) * VB * 2 * sizeof(double)));
    CU(ensure(ctx, ctx->ob_src_nv, (size_t)std::max(nobst, 1) * sizeof(int)));
    CU(ensure(ctx, ctx->ob_in, (size_t)ob_stride * VB * 2 * sizeof(double)));
    CU(ensure(ctx, ctx->ob_nv, (size_t)ob_stride * sizeof(int)));
    CU(ensure(ctx, ctx->ob_f, (size_t)ob_stride * F * sizeof(float4)));
    CU(ensure(ctx, ctx->ob_c, (size_t)ob_stride * 3 * sizeof(float)));
    CU(ensure(ctx, ctx->ob_box, (size_t)(ob_stride / 32 + 1) * sizeof(float4)));
    CU(ensure(ctx, ctx->ob_order, (size_t)ob_stride * sizeof(int)));
    CU(ensure(ctx, ctx->sg_f, (size_t)sg_stride * 2 * sizeof(float4)));
    CU(ensure(ctx, ctx->sg_box, (size_t)(sg_stride / 32 + 1) * sizeof(float4)));
    cudaStream_t st = ctx->stream;
    // the small per-scene arrays travel as one packed block (one driver call instead of nine)
    {
        const size_t n_org = (size_t)nscenes * 2 * sizeof(double), n_sg = (size_t)sg_stride * 4 * sizeof(double);
        const size_t n_sso = (size_t)(nscenes + 1) * 4, n_st1 = (size_t)(nsteps + 1) * 4, n_st = (size_t)std::max(nsteps, 1) * 4;
        const size_t n_ss = (size_t)sg_stride * 4;
        size_t off = 0;
        auto take = [&](size_t n) { size_t o = off; off += (n + 15) / 16 * 16; return o; };
        const size_t o_org = take(n_org), o_sg = take(n_sg), o_sso = take(n_sso), o_slot = take(n_st1), o_src = take(n_st1),
                     o_cnt = take(n_st), o_sb = take(n_st), o_se = take(n_st), o_ss = take(n_ss);
        if (off > ctx->h_meta_cap) {
            if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
            ctx->h_meta = nullptr;
            ctx->h_meta_cap = 0;
            CU(cudaMallocHost(&ctx->h_meta, off * 2));
            ctx->h_meta_cap = off * 2;
        }
        CU(ensure(ctx, ctx->meta, off));
        char *h = (char *)ctx->h_meta, *d = (char *)ctx->meta.p;
        memcpy(h + o_org, origins, n_org);
        if (nslots) memcpy(h + o_sg, dsegs.data(), (size_t)nslots * 4 * sizeof(double));
        memcpy(h + o_sso, scene_step_off, n_sso);
        memcpy(h + o_slot, slot_off.data(), n_st1);
        memcpy(h + o_src, step_obst_off, n_st1);
        if (nsteps) {
            memcpy(h + o_cnt, step_cnt.data(), (size_t)nsteps * 4);
            memcpy(h + o_sb, dev_begin.data(), (size_t)nsteps * 4);
            memcpy(h + o_se, dev_end.data(), (size_t)nsteps * 4);
        }
        if (nslots) memcpy(h + o_ss, dscene.data(), (size_t)nslots * 4);
        CU(cudaMemcpyAsync(d, h, off, cudaMemcpyHostToDevice, st));
        ctx->origins_d.p = d + o_org; ctx->sg_64.p = d + o_sg; ctx->scene_step_off_d.p = d + o_sso;
        ctx->step_obst_off.p = d + o_slot; ctx->ob_src_off.p = d + o_src; ctx->step_obst_cnt.p = d + o_cnt;
        ctx->step_seg_begin.p = d + o_sb; ctx->step_seg_end.p = d + o_se; ctx->seg_scene.p = d + o_ss;
        // captured graphs hold these addresses: a different layout is a different graph
        ctx->meta_sig = (long long)(o_sg * 1315423911ull ^ o_sso * 2654435761ull ^ o_slot * 40503ull ^ o_ss * 97ull ^ off);
    }
    CU(cudaMemsetAsync(ctx->maxabs.p, 0, sizeof(float), st));
    if (nobst) {
        if (Vo == VB) {
            CU(cudaMemcpyAsync(ctx->ob_src.p, obst, (size_t)nobst * VB * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {   // caller's vertex stride differs from the kernel's: repack rows
            CU(cudaMemcpy2DAsync(ctx->ob_src.p, (size_t)VB * 2 * sizeof(double), obst, (size_t)Vo * 2 * sizeof(double),
                                 (size_t)std::min(Vo, VB) * 2 * sizeof(double), nobst, cudaMemcpyHostToDevice, st));
        }
        CU(cudaMemcpyAsync(ctx->ob_src_nv.p, obst_nv, (size_t)nobst * sizeof(int), cudaMemcpyHostToDevice, st));
        int blocks = (noslots + 127) / 128;
        if (VB == 4) {
            k_order_obst<4><<<nsteps, 256, 0, st>>>((const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_src_off.p,
                                                    (const int *)ctx->step_obst_off.p, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                    (const double *)ctx->origins_d.p, (int *)ctx->ob_order.p);
            k_stage_obst<4><<<blocks, 128, 0, st>>>(noslots, (const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_order.p,
                                                   (const int *)ctx->step_obst_off.p, nsteps, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                   (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, (float *)ctx->ob_c.p, (float4 *)ctx->ob_box.p,
                                                   (double *)ctx->ob_in.p, (int *)ctx->ob_nv.p, ob_stride, (float *)ctx->maxabs.p);
        } else {
            k_order_obst<8><<<nsteps, 256, 0, st>>>((const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_src_off.p,
                                                    (const int *)ctx->step_obst_off.p, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                    (const double *)ctx->origins_d.p, (int *)ctx->ob_order.p);
            k_stage_obst<8><<<blocks, 128, 0, st>>>(noslots, (const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_order.p,
                                                   (const int *)ctx->step_obst_off.p, nsteps, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                   (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, (float *)ctx->ob_c.p, (float4 *)ctx->ob_box.p,
                                                   (double *)ctx->ob_in.p, (int *)ctx->ob_nv.p, ob_stride, (float *)ctx->maxabs.p);
        }
        CU(cudaGetLastError());
    }
    if (nslots) {
        k_stage_segs<<<(nslots + 127) / 128, 128, 0, st>>>(nslots, (const double *)ctx->sg_64.p, (const int *)ctx->seg_scene.p, (const double *)ctx->origins_d.p,
                                                         (float4 *)ctx->sg_f.p, sg_stride, (float4 *)ctx->sg_box.p, (float *)ctx->maxabs.p);
        CU(cudaGetLastError());
    }
    float m = 0.f;
    CU(cudaMemcpyAsync(&m, ctx->maxabs.p, sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));      // host buffers (incl. the local staging vectors) may be reused after return
    ctx->maxabs_scene = m;
    ctx->nscenes = nscenes; ctx->nsteps = nsteps; ctx->nobst = nobst; ctx->nseg = nslots; ctx->sg_stride = sg_stride;
    ctx->ob_stride = ob_stride;
    ctx->ob_vmax = VB; ctx->VB = VB; ctx->max_obst_step = max_o; ctx->max_seg_step = max_s;
    ctx->scene_step_off.assign(scene_step_off, scene_step_off + nscenes + 1);
    ctx->origins.assign(origins, origins + 2 * (size_t)nscenes);
    ctx->have_scenes = true;
    ctx->prepared = false;
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// result block on the device: [Counters (64 B)] [mask words]; one memset clears it, one copy brings it back
static_assert(sizeof(Counters) == 128, "Counters is the 128-byte head of the result block");

static KParams make_params(mpc_ctx *ctx) {
    KParams p;
    memset(&p, 0, sizeof p);
    p.lib_v = (const float4 *)ctx->lib_v.p; p.lib_n = (const float4 *)ctx->lib_n.p; p.lib_c = (const float4 *)ctx->lib_c.p;
    p.slib_v = (const float4 *)ctx->slib_v.p; p.slib_n = (const float4 *)ctx->slib_n.p; p.slib_c = (const float4 *)ctx->slib_c.p;
    p.set_total = ctx->set_total;
    p.lib_v64 = (const double *)ctx->lib_v64.p; p.lib_nv = (const int *)ctx->lib_nv.p; p.slot_t = (const int *)ctx->slot_t.p;
    p.PJ = ctx->P * ctx->J; p.J = ctx->J; p.lib_vmax = ctx->lib_vmax;
    p.ob_f = (const float4 *)ctx->ob_f.p; p.ob_c = (const float *)ctx->ob_c.p; p.ob_box = (const float4 *)ctx->ob_box.p;
    p.ob_v64 = (const double *)ctx->ob_in.p; p.ob_nv = (const int *)ctx->ob_nv.p;
    p.ob_stride = ctx->ob_stride; p.ob_vmax = ctx->ob_vmax;
    p.step_obst_off = (const int *)ctx->step_obst_off.p; p.step_obst_cnt = (const int *)ctx->step_obst_cnt.p;
    p.step_seg_begin = (const int *)ctx->step_seg_begin.p;
    p.step_seg_end = (const int *)ctx->step_seg_end.p;
    p.sg_f = (const float4 *)ctx->sg_f.p; p.sg_box = (const float4 *)ctx->sg_box.p; p.sg_64 = (const double *)ctx->sg_64.p;
    p.sg_stride = ctx->sg_stride;
    p.maxabs_scene = (const float *)ctx->maxabs.p;
    char *qp = (char *)ctx->qpack_d.p;
    p.q = (const QDev *)qp; p.cta_off = (const int *)(qp + ctx->qpack_cta_off); p.cand = (const int *)ctx->cand_d.p;
    p.cand_pos = (const int *)ctx->sets_pos.p;
    p.B = ctx->B;
    p.cnt = (Counters *)ctx->res_d.p; p.mask = (unsigned *)((char *)ctx->res_d.p + sizeof(Counters));
    p.mask_sorted = p.mask + ctx->mask_words;
    p.nwords = ctx->mask_words;
    p.wl = (int4 *)ctx->wl.p; p.wl_cap = ctx->wl_cap;
    p.maxabs_query = (const float *)(qp + ctx->qpack_hdr_off); p.lib_rmax = ctx->lib_rmax; p.mode = ctx->mode;
    p.gpc = ctx->gpc; p.cap_o = ctx->cap_o; p.cap_s = ctx->cap_s;
    return p;
}

template <int VA, int VB>
static cudaError_t launch_check_vb(mpc_ctx *ctx, const KParams &p) {
    const bool count = ctx->mode & MPC_MODE_COUNT_WORK, nocull = ctx->mode & MPC_MODE_NO_CULL;
    auto kern = nocull ? (count ? k_check<VA, VB, true, true> : k_check<VA, VB, false, true>)
                       : (count ? k_check<VA, VB, true, false> : k_check<VA, VB, false, false>);
    const int key = VA * 1000 + VB * 10 + (count ? 1 : 0) + (nocull ? 2 : 0);
    if (ctx->smem > 48 * 1024 && ctx->smem_attr[key] < ctx->smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem);
        if (e != cudaSuccess) return e;
        ctx->smem_attr[key] = ctx->smem;
    }
    kern<<<ctx->ctas, K_THREADS, ctx->smem, ctx->stream>>>(p);
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
    CU(cudaMemsetAsync(ctx->res_d.p, 0, sizeof(Counters) + (size_t)2 * ctx->mask_words * sizeof(unsigned), st));
    if (e1) CU(cudaEventRecord(e1, st));
    CU(launch_check(ctx, p));
    if (e2) CU(cudaEventRecord(e2, st));
    k_refine<<<std::max(1, ctx->prop.multiProcessorCount) * 8, 128, 0, st>>>(p);     // one warp per item, 32 warps/SM
    CU(cudaGetLastError());
    return MPC_OK;
}

static int issue_uploads(mpc_ctx *ctx);
static cudaError_t ensure_result_host(mpc_ctx *ctx, size_t bytes);

static int prepare_host(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                        unsigned mode, bool upload) {
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
    ctx->threads = K_THREADS;
    const int nwarps = K_THREADS / 32;
    // groups (of 32 candidates) per CTA: enough CTAs for ~2 waves of 4 CTAs per SM, whole rounds of the CTA's warps
    int gpc = (int)((total_groups + (long long)sms * 8 - 1) / ((long long)sms * 8));
    gpc = std::max(1, std::min(gpc, GPC_MAX));
    if (gpc > nwarps / 2) gpc = std::min(GPC_MAX, (gpc + nwarps - 1) / nwarps * nwarps);
    ctx->gpc = gpc;
    const int F = ob_fields(ctx->VB), FA = ob_fields(ctx->VA);
    ctx->cap_o = std::max(32, std::min((ctx->max_obst_step + 31) / 32 * 32, ctx->VB == 4 ? 320 : 160));
    ctx->cap_s = std::max(32, std::min((ctx->max_seg_step + 31) / 32 * 32, 512));
    ctx->smem = 16 + 3 * GPC_MAX * 4 + (size_t)nwarps * QCAP * 4 + (size_t)nwarps * FA * 32 * 16 +
                (size_t)ctx->cap_o * (F * 16 + 12) + (size_t)ctx->cap_s * 32;
    // pack: QDev[B], cta_off[B+1] (one copy), explicit candidates (second copy, only when present)
    size_t bytes_q = sizeof(QDev) * (size_t)std::max(B, 1), bytes_c = (sizeof(int) * (size_t)(B + 1) + 15) / 16 * 16;
    const size_t bytes_h = 16, bytes_i = sizeof(int) * (size_t)ncand;
    CU(ensure_pinned(ctx, bytes_q + bytes_c + bytes_h + bytes_i + 64));
    QDev *hq = (QDev *)ctx->h_pin;
    int *hcta = (int *)((char *)ctx->h_pin + bytes_q);
    float *hhdr = (float *)((char *)ctx->h_pin + bytes_q + bytes_c);
    int *hidx = (int *)((char *)ctx->h_pin + bytes_q + bytes_c + bytes_h);
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
        o.pos_base = -1;
        o.pad0 = o.pad1 = o.pad2 = 0;
        if (d.cand_set < 0) {
            o.cand_base = 2 * ctx->set_total + d.cand_off;
        } else if (d.cand_off == 0 && d.cand_cnt == ctx->set_off[d.cand_set + 1] - ctx->set_off[d.cand_set] && d.cand_cnt > 512) {
            o.cand_base = ctx->set_total + ctx->set_off[d.cand_set];        // whole set: spatially sorted copy
            o.pos_base = ctx->set_off[d.cand_set];
        } else {
            o.cand_base = ctx->set_off[d.cand_set] + d.cand_off;
        }
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
    ctx->maxabs_query = std::nextafter((float)maxq, INFINITY);
    hhdr[0] = ctx->maxabs_query;
    hhdr[1] = hhdr[2] = hhdr[3] = 0.f;
    CU(ensure(ctx, ctx->qpack_d, bytes_q + bytes_c + bytes_h));
    ctx->qpack_cta_off = bytes_q;
    ctx->qpack_hdr_off = bytes_q + bytes_c;
    ctx->up_bytes = bytes_q + bytes_c + bytes_h;
    ctx->up_cand_bytes = bytes_i;
    ctx->up_cand_src = bytes_q + bytes_c + bytes_h;
    // candidate arena = staged sets followed by this call's explicit lists
    size_t arena = sizeof(int) * (size_t)std::max(2 * ctx->set_total + ncand, 1);
    void *arena_before = ctx->cand_d.p;
    CU(ensure(ctx, ctx->cand_d, arena));
    if ((ctx->cand_d.p != arena_before || !ctx->arena_valid) && ctx->set_total)
        CU(cudaMemcpyAsync(ctx->cand_d.p, ctx->sets.p, sizeof(int) * (size_t)2 * ctx->set_total, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->arena_valid = true;
    CU(ensure(ctx, ctx->res_d, sizeof(Counters) + sizeof(unsigned) * (size_t)2 * std::max<long long>(words, 1)));
    CU(ensure_result_host(ctx, sizeof(Counters) + sizeof(unsigned) * (size_t)std::max<long long>(words, 1)));
    ctx->B = B; ctx->ctas = (int)ctas; ctx->mask_words = (int)words; ctx->mode = mode;
    ctx->prepared = true;
    if (upload) return issue_uploads(ctx);
    return MPC_OK;
}

// host -> device copies of the packed batch (pinned staging buffer -> query pack, explicit candidate lists)
static int issue_uploads(mpc_ctx *ctx) {
    CU(cudaMemcpyAsync(ctx->qpack_d.p, ctx->h_pin, ctx->up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->up_cand_bytes)
        CU(cudaMemcpyAsync((int *)ctx->cand_d.p + 2 * ctx->set_total, (const char *)ctx->h_pin + ctx->up_cand_src, ctx->up_cand_bytes,
                           cudaMemcpyHostToDevice, ctx->stream));
    return MPC_OK;
}

extern "C" int mpc_prepare_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                                 unsigned mode) {
    return prepare_host(ctx, B, queries, cand_idx, ncand, mode, true);
}

static void fill_stats(mpc_ctx *ctx, mpc_stats *st, float ms, int kernels, int overflow) {
    if (!st) return;
    const Counters &c = *(const Counters *)ctx->h_res;
    st->pairs_nominal = (int64_t)c.pairs_nominal;
    st->refined_fp64 = (int64_t)c.refined;
    st->exact_ties = (int64_t)c.exact_ties;
    st->near_tangent = (int64_t)c.near_tangent;
    st->parity_refined = (int64_t)c.parity_refined;
    st->cull_tests = (int64_t)c.cull_tests;
    st->axis_tests = (int64_t)c.axis_tests;
    st->box_skipped = (int64_t)c.box_skipped;
    st->worklist_overflow = overflow;
    st->ms_device = ms;
    st->ctas = ctx->ctas;
    st->kernels = kernels;
    st->tau = std::max(std::max(ctx->maxabs_scene, ctx->maxabs_query) / 131072.0f, 1e-7f);
}

static cudaError_t ensure_result_host(mpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_res_cap) return cudaSuccess;
    invalidate_graphs(ctx);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    ctx->h_res = nullptr;
    ctx->h_res_cap = 0;
    size_t want = std::max<size_t>(bytes * 2, 4096);
    cudaError_t e = cudaMallocHost(&ctx->h_res, want);
    if (e == cudaSuccess) ctx->h_res_cap = want;
    return e;
}

// run the prepared batch once and bring the result block (counters + mask) to pinned host memory with a single
// synchronisation; grows the refinement list and re-runs if it overflowed
static int run_once(mpc_ctx *ctx, float *ms, int *kernels, int *overflow, bool timed) {
    *overflow = 0;
    *kernels = 0;
    const size_t bytes = sizeof(Counters) + (size_t)ctx->mask_words * sizeof(unsigned);
    CU(ensure_result_host(ctx, bytes));
    for (int attempt = 0; attempt < 3; attempt++) {
        int rc = launch_all(ctx, timed ? ctx->ev[0] : nullptr, nullptr, nullptr);
        if (rc) return rc;
        *kernels += 2;
        if (timed) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_res, ctx->res_d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (timed) CU(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
        const Counters *c = (const Counters *)ctx->h_res;
        if (c->wl_count <= (unsigned)ctx->wl_cap) return MPC_OK;
        *overflow = 1;
        ctx->wl_cap = (int)std::min<unsigned long long>((unsigned long long)c->wl_count * 5 / 4 + 1024, 0x7fffffffULL / sizeof(int4));
        CU(ensure(ctx, ctx->wl, sizeof(int4) * (size_t)ctx->wl_cap));
    }
    return fail(ctx, MPC_E_LIMIT, "refinement list still overflows at %d entries", ctx->wl_cap);
}

// the whole query as one CUDA graph: H2D of the batch, clear, k_check, k_refine, D2H of the result block.  A planner
// issues hundreds of small queries of a handful of shapes; replaying an instantiated graph replaces six driver calls
// by one.  Graphs are keyed by everything that is baked into their nodes and dropped whenever a buffer moves.
static int run_graph(mpc_ctx *ctx, float *ms, bool timed) {
    const size_t bytes = sizeof(Counters) + (size_t)ctx->mask_words * sizeof(unsigned);
    std::vector<long long> key = {(long long)ctx->B, (long long)ctx->ctas, (long long)ctx->mask_words, (long long)ctx->mode,
                                  (long long)ctx->gpc, (long long)ctx->cap_o, (long long)ctx->cap_s, (long long)ctx->smem,
                                  (long long)ctx->up_bytes, (long long)ctx->up_cand_bytes, (long long)ctx->wl_cap,
                                  (long long)ctx->VA, (long long)ctx->VB, (long long)ctx->ob_stride, (long long)ctx->sg_stride,
                                  (long long)ctx->ob_vmax, (long long)ctx->P, (long long)ctx->J, (long long)ctx->lib_vmax,
                                  (long long)__builtin_bit_cast(unsigned, ctx->lib_rmax), (long long)ctx->set_total, ctx->meta_sig};
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        if (ctx->graphs.size() > 256) invalidate_graphs(ctx);
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int rc = issue_uploads(ctx);
        if (!rc) rc = launch_all(ctx, nullptr, nullptr, nullptr);
        cudaError_t e1 = cudaSuccess;
        if (!rc) e1 = cudaMemcpyAsync(ctx->h_res, ctx->res_d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return fail(ctx, MPC_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        }
        cudaGraphExec_t exec = nullptr;
        cudaError_t e3 = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e3 != cudaSuccess) return fail(ctx, MPC_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e3));
        it = ctx->graphs.emplace(key, exec).first;
    }
    if (timed) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaGraphLaunch(it->second, ctx->stream));
    if (timed) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (timed) CU(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return MPC_OK;
}

extern "C" int mpc_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                         unsigned mode, uint32_t *out_mask, int mask_words, mpc_stats *stats) {
    int rc = prepare_host(ctx, B, queries, cand_idx, ncand, mode, false);
    if (rc) return rc;
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words > 0 && !out_mask) return fail(ctx, MPC_E_ARG, "out_mask missing");
    float ms = 0.f;
    int kernels = 2, overflow = 0;
    if (ctx->smem > 48 * 1024) {          // large tiles need the opt-in attribute, which cannot be set during capture
        rc = issue_uploads(ctx);
        if (!rc) rc = run_once(ctx, &ms, &kernels, &overflow, stats != nullptr);
    } else {
        rc = run_graph(ctx, &ms, stats != nullptr);
        if (!rc && ((const Counters *)ctx->h_res)->wl_count > (unsigned)ctx->wl_cap) {
            // refinement list overflowed: the plain path grows it and re-runs (and drops the graphs)
            rc = issue_uploads(ctx);
            if (!rc) rc = run_once(ctx, &ms, &kernels, &overflow, stats != nullptr);
            overflow = 1;
        }
    }
    if (rc) return rc;
    if (mask_words) memcpy(out_mask, (const char *)ctx->h_res + sizeof(Counters), sizeof(unsigned) * (size_t)mask_words);
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

extern "C" int mpc_run_prepared(mpc_ctx *ctx, int iters, int flush_l2, float *ms_total, float *ms_main, mpc_stats *stats) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    CU(cudaSetDevice(ctx->device));
    float ms = 0.f;
    int kernels = 0, overflow = 0;
    int rc = run_once(ctx, &ms, &kernels, &overflow, true);     // sizes the refinement list
    if (rc) return rc;
    const size_t flush_bytes = 256u << 20;
    if (flush_l2) CU(ensure(ctx, ctx->flush, flush_bytes));
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
    const size_t bytes = sizeof(Counters) + (size_t)ctx->mask_words * sizeof(unsigned);
    CU(cudaMemcpyAsync(ctx->h_res, ctx->res_d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

extern "C" int mpc_fetch_mask(mpc_ctx *ctx, uint32_t *out_mask, int mask_words) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words) {
        CU(cudaMemcpyAsync(out_mask, (const char *)ctx->res_d.p + sizeof(Counters), sizeof(unsigned) * (size_t)mask_words, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return MPC_OK;
}
