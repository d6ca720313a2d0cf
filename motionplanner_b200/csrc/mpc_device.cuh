// mpc_device.cuh -- device code of libmpcheck.so: staging kernels, k_check (float32 separating-axis pass with
// TMA-staged tiles, block-box culling, packed FFMA2 centre-line axis, per-warp SAT queue) and k_refine (float64 /
// exact tie-break).  Included by mpc_kernels.cu (host side / C ABI).  See the header of mpc_kernels.cu and DESIGN.md.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mpc.h"
#include "mpc_exact.cuh"

// index invariants of the kernels, compiled in by `make debug` (-DMPC_DEBUG_ASSERTS; libmpcheck_dbg.so, selected with
// MPC_LIB_PATH): compute-sanitizer is not available on the measurement pool, the parity suites run against this build instead
#ifdef MPC_DEBUG_ASSERTS
#include <cassert>
#define MPC_ASSERT(c) assert(c)
#else
#define MPC_ASSERT(c) ((void)0)
#endif

namespace mpc {

constexpr float BIG = 1e30f;
constexpr float FAR = 1e18f;
constexpr int GPC_MAX = 128;      // candidate groups (of 32) per CTA
static_assert(GPC_MAX <= 128, "the compaction scan gives every lane of one warp four words");
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
    unsigned wl_count;
    float tau;                   // the float32 error band the launch used (k_check block 0 reports it)
    unsigned long long box_skipped, done_skipped;
    unsigned done_ctas, pad0;    // CTAs of k_check that have finished (single-kernel queries: the last one refines and publishes)
    unsigned long long pad1[5];
};

struct KParams {
    // library
    const float4 *lib_v, *lib_n, *lib_c;       // [field][P*J]: gathered through the candidate index
    const float4 *slib_v, *slib_n, *slib_c;    // [field][J][set_total]: copy in sorted-set order, read coalesced
    int set_total;
    const double *lib_v64;
    const int *lib_nv, *slot_t, *slot_perm;      // slot_perm[l]: time slot launched at level l of the slot-major grid
    int PJ, J, lib_vmax;
    // scenes
    const float4 *ob_f;          // [F-1][ob_stride] vertex and edge-line fields, slot layout (steps 32-aligned)
    const float4 *ob_box;        // [ob_stride/8] bounding box of the circles of every sub-block of 8 slots
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
    int c_total;                 // CTAs per time slot: the grid is (c_total, time slots), slot-major
    int tail;                    // k_check<..., WS = true>: the last CTA to finish runs the refinement pass and copies the result
                                 // block to host_res (page-locked, device-visible), then raises host_flag to host_seq
    unsigned *host_res;
    volatile unsigned *host_flag;
    unsigned host_seq;
    int res_words;               // 32-bit words of the result block the host wants (counters + caller-order mask)
    int wsplit;                  // k_check<..., WS = true>: warps per candidate group (2, 4 or 8)
    int compact;                 // renumber the undecided candidates densely (needs: every library slot occupied, big CTAs)
    int cmin;                    // ... in CTAs of at least this many groups (the pre-pass has to pay for itself)
    int cross_exit;              // grid larger than one resident wave: later time slots may use the bits of earlier ones
    const QDev *q;
    const int2 *cta_tab;         // [c_total] (query, chunk) of every CTA of one time slot, built by the host with the batch
    const int *cand, *cand_pos, *cand_inv;      // cand_pos: sorted -> caller's position in the set, cand_inv: back
    int B, nwords;
    int wmax, any_sorted;        // most result words of one query; some query uses a spatially sorted candidate set
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

constexpr int REFINE_SCRATCH = 2 * (MPC_MAX_LIB_VERTS + MPC_MAX_OBST_VERTS);      // doubles of scratch per warp in refine_pass
__device__ void refine_pass_smem(const KParams *ps, unsigned warp0, unsigned warps, double *pts);

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

// programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while its
// predecessor in the stream is still draining; everything it does before pdl_wait() must not depend on that kernel
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// position of the n-th (0-based) set bit of x; x must have more than n bits set
__device__ __forceinline__ int nth_set_bit(unsigned x, unsigned n) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const unsigned c = __popc(x & ((1u << w) - 1u));
        if (n >= c) { n -= c; x >>= w; pos += w; }
    }
    return pos;
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
                                                    const double *__restrict__ origins, int *__restrict__ order, int vsrc) {
    __shared__ unsigned s_key[SORT_MAX];
    __shared__ int s_idx[SORT_MAX];
    const int step = blockIdx.x;
    const int o0 = src_off[step], cnt = src_off[step + 1] - o0, s0 = slot_off[step], nslot = slot_off[step + 1] - s0;
    // a single block of 32 slots is never culled by boxes: its order does not matter
    if (cnt > SORT_MAX || cnt <= 32) {
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
            const double *v = src64 + (size_t)(o0 + i) * vsrc * 2;   // caller's vertex stride
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
                             double *__restrict__ ob64, int *__restrict__ ob_nv, int stride, float *maxabs, int vsrc) {
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
        const double *sv = src64 + (size_t)o * vsrc * 2;
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
        // the separating-axis test needs convex pieces (a reflex edge would "separate" what it must not): every turn
        // to the left, one revolution in all.  Offenders are counted, mpc_set_scenes fails with MPC_E_ARG.
        if (ok) {
            double ux[VB], uy[VB], turn = 0.0;        // the ring without repeated vertices (zero-length edges turn nowhere)
            int m = 0;
            for (int k = 0; k < nv; k++) {
                if (m > 0 && x[k] == ux[m - 1] && y[k] == uy[m - 1]) continue;
                ux[m] = x[k]; uy[m] = y[k]; m++;
            }
            while (m > 1 && ux[m - 1] == ux[0] && uy[m - 1] == uy[0]) m--;
            bool reflex = false;
            for (int k = 0; k < m; k++) {
                const int k1 = (k + 1 == m) ? 0 : k + 1, k2 = (k1 + 1 == m) ? 0 : k1 + 1;
                const double ex = ux[k1] - ux[k], ey = uy[k1] - uy[k], fx = ux[k2] - ux[k1], fy = uy[k2] - uy[k1];
                const double cr = ex * fy - ey * fx;
                if (cr < -1e-12 * (fabs(ex * fy) + fabs(ey * fx)) - 1e-300) reflex = true;
                turn += atan2(cr, ex * fx + ey * fy);
            }
            if (m >= 3 && (reflex || fabs(turn - 6.283185307179586) > 1e-6)) atomicAdd(reinterpret_cast<int *>(maxabs) + 1, 1);
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
    // bounding box of the circles of every sub-block of 8 slots (of the float32 values the main kernel will see)
    const bool real = nv > 0;
    float x0 = real ? ccx - ccr : BIG, x1 = real ? ccx + ccr : -BIG, y0 = real ? ccy - ccr : BIG, y1 = real ? ccy + ccr : -BIG;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        x0 = fminf(x0, __shfl_xor_sync(FULL, x0, d)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, d));
        y0 = fminf(y0, __shfl_xor_sync(FULL, y0, d)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, d));
    }
    if ((threadIdx.x & 7) == 0)
        ob_box[slot >> 3] = make_float4(__fadd_rd(x0, -1e-6f * fabsf(x0)), __fadd_rd(y0, -1e-6f * fabsf(y0)),
                                        __fadd_ru(x1, 1e-6f * fabsf(x1)), __fadd_ru(y1, 1e-6f * fabsf(y1)));
}

// Large scene sets: one CTA per distinct segment range sorts it along a Morton curve of the segment midpoints (the same
// order the host computes for small sets, mpc_set_scenes pass 2) and writes the 32-aligned, padded slot range: float64
// segments for the tie-break pass and the owning scene per slot.  Ranges longer than SEG_SORT_MAX stay on the host.
constexpr int SEG_SORT_MAX = 4096;
__global__ void __launch_bounds__(256) k_order_segs(const int4 *__restrict__ jobs, const double *__restrict__ raw,
                                                    double *__restrict__ sg64, int *__restrict__ seg_scene) {
    __shared__ unsigned s_key[SEG_SORT_MAX];
    __shared__ unsigned short s_idx[SEG_SORT_MAX];
    __shared__ double s_red[4][8];
    const int4 J = jobs[blockIdx.x];                       // (begin, end, scene, first slot)
    const int n = J.y - J.x, padded = (n + 31) / 32 * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double *g = raw + 4 * (size_t)(J.x + i);
        const double mx = 0.5 * (g[0] + g[2]), my = 0.5 * (g[1] + g[3]);
        x0 = fmin(x0, mx); x1 = fmax(x1, mx); y0 = fmin(y0, my); y1 = fmax(y1, my);
    }
    for (int o = 16; o; o >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(FULL, x0, o)); x1 = fmax(x1, __shfl_xor_sync(FULL, x1, o));
        y0 = fmin(y0, __shfl_xor_sync(FULL, y0, o)); y1 = fmax(y1, __shfl_xor_sync(FULL, y1, o));
    }
    if (lane == 0) { s_red[0][warp] = x0; s_red[1][warp] = x1; s_red[2][warp] = y0; s_red[3][warp] = y1; }
    __syncthreads();
    for (int w = 0; w < 8; w++) {
        x0 = fmin(x0, s_red[0][w]); x1 = fmax(x1, s_red[1][w]); y0 = fmin(y0, s_red[2][w]); y1 = fmax(y1, s_red[3][w]);
    }
    const double sx = (x1 > x0) ? 65535.0 / (x1 - x0) : 0.0, sy = (y1 > y0) ? 65535.0 / (y1 - y0) : 0.0;
    int n2 = 32;
    while (n2 < n) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned key = 0xffffffffu;
        if (i < n) {
            const double *g = raw + 4 * (size_t)(J.x + i);
            const double mx = 0.5 * (g[0] + g[2]), my = 0.5 * (g[1] + g[3]);
            key = morton16_dev((unsigned)((mx - x0) * sx), (unsigned)((my - y0) * sy));
        }
        s_key[i] = key;
        s_idx[i] = (unsigned short)i;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ jj;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned a = s_key[i], b = s_key[l];
                    const bool gt = a > b || (a == b && s_idx[i] > s_idx[l]);      // ties by source index: deterministic
                    if (gt == up) {
                        s_key[i] = b; s_key[l] = a;
                        const unsigned short t = s_idx[i]; s_idx[i] = s_idx[l]; s_idx[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        double *dst = sg64 + 4 * (size_t)(J.w + i);
        if (i < n) {
            const double *g = raw + 4 * (size_t)(J.x + s_idx[i]);
            dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2]; dst[3] = g[3];
            seg_scene[J.w + i] = J.z;
        } else {
            dst[0] = dst[1] = dst[2] = dst[3] = 0.0;
            seg_scene[J.w + i] = -1;
        }
    }
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
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {            // one box per sub-block of 8 slots
        x0 = fminf(x0, __shfl_xor_sync(FULL, x0, d)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, d));
        y0 = fminf(y0, __shfl_xor_sync(FULL, y0, d)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, d));
    }
    if ((threadIdx.x & 7) == 0 && s < nslots) sg_box[s >> 3] = make_float4(x0, y0, x1, y1);
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

// (A scalar-FFMA form of the same test was measured in k_check_full: the register-only instruction mix runs 15 % faster
// with scalar FFMA -- 60.1 vs 52.4 TFLOP/s, mpc_measure_fp32_peak -- because FFMA2 issues every 2.18 clk against 1.05 clk
// for FFMA while min / max take 2 clk and do not overlap with either; inside the kernel the extra 32 issue slots per
// test cost more than that: 2.42e11 checks/s scalar vs 2.73e11 packed.)
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
#ifndef K_MIN_CTAS
#define K_MIN_CTAS 4
#endif
constexpr int QCAP = 64;            // per-warp queue of (lane, record) pairs that passed axis 0

// SEGCULL: a step may have more than one block of 32 boundary segments, compile the sub-block culling of segments in.
// The kernel is sensitive to its shape: with that code present C4 (4 segments per step) runs 5 % slower; the same
// switch for obstacles changes nothing, and funnelling the three inlined copies of each drain into one call site
// (smaller code, more control flow per round) costs 8-15 %.
// WS: "warps over obstacles".  A CTA with fewer candidate groups than warps (planner-sized queries: one to four groups
// per CTA) gives every group to WO = p.wsplit warps; each takes every WO-th block of 32 obstacle slots / boundary
// segments of the tile and the partial results (collision bits, crossing parity) are merged in shared memory.  Without it
// seven of the eight warps of such a CTA idle while one walks the whole tile (2362 segments on the ZAM_ACC map).
template <int VA, int VB, bool COUNT, bool SEGCULL, bool WS>
__global__ void __launch_bounds__(K_THREADS, K_MIN_CTAS) k_check(const __grid_constant__ KParams p) {
    constexpr int F = ob_fields(VB), FA = ob_fields(VA), NW = K_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *s_coll = reinterpret_cast<unsigned *>(smem_raw + 16);
    unsigned *s_par = s_coll + GPC_MAX;
    unsigned *s_punc = s_par + GPC_MAX;
    unsigned *s_act = s_punc + GPC_MAX;                                        // WS: candidates of a group that hold an occupancy
    unsigned *s_q = s_act + GPC_MAX;                                           // [NW][QCAP]
    float4 *s_A = reinterpret_cast<float4 *>(s_q + NW * QCAP);                 // [NW][FA][32] occupancy polygons
    float4 *s_ob = s_A + NW * FA * 32;                                         // [F][cap_o] obstacle records
    float *s_c = reinterpret_cast<float *>(s_ob + F * p.cap_o);                // circle x | y | -(r + rcull)^2
    float4 *s_sg = reinterpret_cast<float4 *>(s_c + 3 * p.cap_o);              // [2][cap_s] segments

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int WO = WS ? p.wsplit : 1;                 // warps per group
    const int wg = WS ? warp / WO : warp, wo = WS ? warp - wg * WO : 0, GW = WS ? NW / WO : NW;
    pdl_trigger();                    // the next kernel of the step (main launch / k_refine) may start its own prologue

    // which (slot, query, chunk) is this CTA.  The grid is SLOT-major: blockIdx.y counts the entries of the slot order,
    // blockIdx.x the (query, chunk) pairs of the batch; CTAs start in index order, so by the time the later slots of a
    // batch run, the earlier ones have published their collision bits and whole groups can return early (see `done`
    // below).  (query, chunk) come from a table the host packs with the batch.
    const int2 ent = __ldg(p.cta_tab + blockIdx.x);
    const int q = ent.x;
    const QDev *Q = p.q + q;
    const int chunk = ent.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {     // for the statistics: the host never reads the scene's extent itself
        pdl_wait();                        // (the result block is cleared by the kernel before this one)
        p.cnt->tau = fmaxf(fmaxf(*p.maxabs_scene, *p.maxabs_query) * (1.0f / 131072.0f), 1e-7f);
    }
    const int cand_cnt = Q->cand_cnt, cand_base = Q->cand_base;
    const int gbeg = chunk * p.gpc;
    const int ng = min(p.gpc, (cand_cnt + 31) / 32 - gbeg);
    const float lx = Q->lx, ly = Q->ly, lc = Q->lc, ls = Q->ls;
    const int q_t0 = Q->t0, q_limit = Q->step_limit, q_nsteps = Q->nsteps, q_step0 = Q->step0;
    // float32 error band: every float32 signed distance below is within tau of its float64 value
    const float tau = fmaxf(fmaxf(*p.maxabs_scene, *p.maxabs_query) * (1.0f / 131072.0f), 1e-7f);
    const bool noexit = p.mode & MPC_MODE_NO_EARLY_EXIT;
    // cull radius shared by the whole CTA: largest library bounding radius + band (conservative for smaller polygons)
    const float rcull = p.lib_rmax + 2.0f * tau;
    unsigned n_cull = 0, n_axis = 0, n_skip = 0, n_done = 0;
    const unsigned *maskp = (Q->pos_base < 0 ? p.mask : p.mask_sorted) + Q->mask_off;
    const int pos_base = Q->pos_base;
    float4 *sA = s_A + warp * FA * 32;
    unsigned *sq = s_q + warp * QCAP;
    unsigned nom = 0;                             // nominal pairs of this warp's groups (one atomic at the end)
    unsigned *s_live = s_par, *s_pref = s_punc, *s_gstart = s_coll;       // s_gstart[g]: word holding rank 32 g
    unsigned *s_tot = reinterpret_cast<unsigned *>(smem_raw + 8);          // (the mbarrier uses bytes 0..7)

    if (tid == 0) mbar_init(bar, 1);
    unsigned tile_no = 0;             // tiles fetched so far (CTA-uniform): parity of the mbarrier phase to wait for
    bool waited = false;              // pdl_wait() done (CTA-uniform)

    // (A variant in which a scout CTA walked several entries of the slot order in sequence, its candidates' collision
    // words kept in shared memory in between, was built and measured: C4 batch 0.1137 / 0.1125 / 0.1109 / 0.1151 / 0.1399
    // ms at 1 / 2 / 3 / 4 / 8 slots per CTA, mixed regime 1.5 % slower at 2-3, and the bookkeeping cost every launch
    // 3.6 % -- removed again)
    auto run_slot = [&]() {
    const int j = __ldg(p.slot_perm + blockIdx.y);
    const int t = q_t0 + __ldg(p.slot_t + j);
    for (int i = tid; i < 4 * GPC_MAX; i += K_THREADS) s_coll[i] = 0;
    __syncthreads();
    // reference guards: ind + time <= param['steps'] and ind + time < len(space)  (lowLevelPlannerManeuverAutomaton.py:120,122)
    if (t > q_limit || t < 0 || t >= q_nsteps) return;
    const int step = q_step0 + t;
    // obstacles of the step: a 32-aligned slot range (padding slots are far away), `nobst` real ones
    const int o0 = p.step_obst_off[step], nslot = p.step_obst_off[step + 1] - o0, nobst = p.step_obst_cnt[step];
    const int g0 = p.step_seg_begin[step];
    const bool has_region = g0 >= 0;
    const int nseg = has_region ? p.step_seg_end[step] - g0 : 0;
    // Everything above reads what the host staged before the step.  What follows reads the result words earlier launches
    // of the step publish (early return across time slots) or adds to the result block: a launch that looks at those
    // words waits for its predecessor here, one that does not waits only when its first tile has landed.
    const bool xexit = p.cross_exit && !noexit;
    if (xexit && !waited) { pdl_wait(); waited = true; }

    const int nphase = (nslot <= p.cap_o && nseg <= p.cap_s) ? 1 : max((nslot + p.cap_o - 1) / p.cap_o, (nseg + p.cap_s - 1) / p.cap_s);

    // ---- early return of the whole CTA, before any tile is fetched.  In a batch big enough that later time slots start
    // after earlier ones have finished, the result words already hold the collision bits of those slots: if none of the
    // CTA's candidates is still undecided there is nothing left to do -- the reference returns at the first violation
    // too (collision_check :124, collisionChecker.py:22,33).  Only the nominal pair statistic is kept exact.
    // Compaction: when every library slot is occupied (so "candidate" = "occupancy present") and the step fits one tile,
    // the undecided candidates of the CTA are renumbered densely -- groups are formed from the live ones only, instead
    // of running every group that still has one straggler.  s_par / s_punc (multi-phase state, unused here) hold the
    // live words and their running counts.  The pre-pass pays from ~32 groups on.
    const bool cmode = !WS && xexit && p.compact && nphase == 1 && ng >= p.cmin;
    int total_live = 0;
    if (cmode) {
        if (warp == 0) {
            unsigned cnt[4], run = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {          // lane owns words 4*lane .. 4*lane+3 (GPC_MAX = 128)
                const int i = 4 * lane + r;
                unsigned live = 0;
                if (i < ng) {
                    const int nvalid = min(32, cand_cnt - (gbeg + i) * 32);
                    const unsigned valid = nvalid >= 32 ? FULL : ((1u << nvalid) - 1u);
                    live = valid & ~__ldcg(maskp + gbeg + i);
                    s_live[i] = live;
                }
                cnt[r] = run;
                run += __popc(live);
            }
            unsigned incl = run;                    // inclusive scan of the lanes' sums
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned excl = incl - run;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int i = 4 * lane + r;
                if (i < ng) {
                    const unsigned a = excl + cnt[r], n = (r < 3 ? excl + cnt[r + 1] : incl) - a;     // ranks [a, a + n)
                    s_pref[i] = a;
                    // at most one multiple of 32 falls into a word's rank range: the word where that group starts
                    const unsigned gfirst = (a + 31u) >> 5;
                    if (gfirst * 32u < a + n) s_gstart[gfirst] = (unsigned)i;
                }
            }
            if (lane == 31) *s_tot = incl;
        }
        __syncthreads();
        total_live = (int)*s_tot;
        if (tid == 0) {                              // the statistic needs no library access here
            const unsigned long long w = (unsigned long long)min(ng * 32, cand_cnt - gbeg * 32) * (unsigned)(nobst + nseg);
            atomicAdd(&p.cnt->pairs_nominal, w);
            if (COUNT) atomicAdd(&p.cnt->done_skipped, (unsigned long long)(min(ng * 32, cand_cnt - gbeg * 32) - total_live) * (unsigned)(nobst + nseg));
        }
        if (total_live == 0) return;                 // nothing undecided: the reference has returned already
    } else if (xexit) {
        bool live = false;
        for (int g = warp; g < ng; g += NW) {
            const int ci = (gbeg + g) * 32 + lane;
            if (ci < cand_cnt && !((__ldcg(maskp + gbeg + g) >> lane) & 1u)) live = true;
        }
        if (!__syncthreads_or(live)) {
            unsigned skipped = 0;
            for (int g = warp; g < ng; g += NW) {
                const int ci = (gbeg + g) * 32 + lane;
                if (ci < cand_cnt) {
                    const size_t pj = pos_base >= 0 ? (size_t)j * p.set_total + pos_base + ci
                                                    : (size_t)__ldg(p.cand + cand_base + ci) * p.J + j;
                    if (__ldg((pos_base >= 0 ? p.slib_c : p.lib_c) + pj).w > 0.f) skipped += (unsigned)(nobst + nseg);
                }
            }
            nom += skipped;
            if (COUNT) n_done += skipped;
            return;
        }
    }
    const int ngl = cmode ? (total_live + 31) / 32 : ng;      // groups this CTA runs

    for (int ph = 0; ph < nphase; ph++) {
        const int oc0 = ph * p.cap_o, ocn = max(0, min(nslot - oc0, p.cap_o));       // multiple of 32
        const int sc0 = ph * p.cap_s, scn = max(0, min(nseg - sc0, p.cap_s));
        const bool loading = (ocn | scn) != 0;
        if (tid == 0 && loading) {
            MPC_ASSERT(ocn <= p.cap_o && scn <= p.cap_s && (ocn & 31) == 0 && o0 + oc0 + ocn <= p.ob_stride && (!scn || g0 + sc0 + scn <= p.sg_stride));
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
            mbar_wait(bar, tile_no & 1);
            tile_no++;
            // third circle array: r  ->  -(r + rcull)^2, the constant term of the centre-line test
            for (int i = tid; i < ocn; i += K_THREADS) {
                float r = s_c[2 * p.cap_o + i] + rcull;
                s_c[2 * p.cap_o + i] = -(r * r);
            }
            __syncthreads();
        }
        if (!waited) { pdl_wait(); waited = true; }
        // candidate index of the warp's first group; the next group's is fetched
        // while the current one is processed, which takes one link out of the dependent-load chain of the set-up
        int nxt_prim = 0;
        {
            const int ci0 = (gbeg + wg) * 32 + lane;
            if (!cmode && pos_base < 0 && wg < ng && ci0 < cand_cnt) nxt_prim = __ldg(p.cand + cand_base + ci0);
        }
        for (int g = wg; g < ngl; g += GW) {
            // ---- lane = one candidate primitive: transform its occupancy polygon, park it in shared memory ----------
            int ci = (gbeg + g) * 32 + lane;
            bool active = ci < cand_cnt;
            if (cmode) {
                // the lane's rank among the CTA's undecided candidates -> word (bisection on the running counts) -> bit
                const int rk = g * 32 + lane;
                active = rk < total_live;
                int lo = (int)s_gstart[g], hi = (g + 1 < ngl) ? (int)s_gstart[g + 1] + 1 : ng;    // usually 1-2 words
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if ((int)s_pref[mid] <= rk) lo = mid; else hi = mid;
                }
                MPC_ASSERT(!active || (lo >= 0 && lo < ng && rk >= (int)s_pref[lo] && rk - (int)s_pref[lo] < __popc(s_live[lo])));
                ci = active ? (gbeg + lo) * 32 + nth_set_bit(s_live[lo], (unsigned)(rk - (int)s_pref[lo])) : 0;
                MPC_ASSERT(!active || ci < cand_cnt);
            }
            unsigned done = 0;
            // library record of the lane's candidate: sorted sets have their own copy of the library in set order
            // (consecutive lanes read consecutive 16-byte records); other candidates are gathered through their index
            const bool direct = pos_base >= 0;
            const float4 *Lv = direct ? p.slib_v : p.lib_v, *Ln = direct ? p.slib_n : p.lib_n;
            const size_t fstride = direct ? (size_t)p.J * p.set_total : (size_t)p.PJ;
            size_t pj = 0;
            if (cmode && !direct && active) nxt_prim = __ldg(p.cand + cand_base + ci);
            if (active) pj = direct ? (size_t)j * p.set_total + pos_base + ci : (size_t)nxt_prim * p.J + j;
            if (!direct && !cmode) {
                const int cin = ci + GW * 32;
                if (g + GW < ng && cin < cand_cnt) nxt_prim = __ldg(p.cand + cand_base + cin);
            }
            float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) cc = __ldg((direct ? p.slib_c : p.lib_c) + pj);
            const int nv = (int)cc.w;
            active = active && nv > 0;
            if (xexit && !cmode) {
                // the same test per group: collision bits other time slots have published (L2, not L1).
                // Every candidate of the group collides at another time step already: the reference returns at the
                // first violation too (collision_check :124, collisionChecker.py:22,33).
                done = __ldcg(maskp + gbeg + g);
                if (__all_sync(FULL, !active || ((done >> lane) & 1u))) {
                    if (active && ph == 0) {
                        nom += (unsigned)(nobst + nseg);
                        if (COUNT) n_done += (unsigned)(nobst + nseg);
                    }
                    continue;
                }
            }
            if (!cmode && active && ph == 0 && wo == 0) nom += (unsigned)(nobst + nseg);
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

            bool collide = cmode ? false : (((s_coll[g] | done) >> lane) & 1u);
            bool par = false, punc = false;
            unsigned qn = 0;                        // entries in this warp's queue (warp-uniform)
            // bounding box of the warp's 32 polygon centres: decides which 32-blocks of the spatially sorted obstacle
            // slots / boundary segments can matter for anyone in the warp
            const bool cull_o = nslot > 32, cull_s = SEGCULL && nseg > 32;
            float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f, wr = 0.f;
            if (cull_o || cull_s) {
                wx0 = warp_min(active ? acx : BIG); wx1 = warp_max(active ? acx : -BIG);
                wy0 = warp_min(active ? acy : BIG); wy1 = warp_max(active ? acy : -BIG);
                wr = warp_max(active ? arr : 0.f);
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
                        MPC_ASSERT(qn + __popc(bal) <= (unsigned)QCAP && base + k < 0x10000);
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
                MPC_ASSERT(!have || (src < 32 && slot < p.cap_o));
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
                const int ci_src = __shfl_sync(FULL, ci, src);
                if (have && !hit && !(sep > tau)) push_work(p, q, ci_src, j, KIND_OBST, o0 + oc0 + slot);
                const unsigned m = __reduce_or_sync(FULL, hit ? (1u << src) : 0u);
                collide |= (m >> lane) & 1u;
            };
            auto drain_seg = [&]() {
                unsigned ent;
                const bool have = pop_batch(ent);
                const int src = ent >> 16, k = ent & 0xffff;
                MPC_ASSERT(!have || (src < 32 && k < p.cap_s));
                float sep = BIG;
                if (have) {
                    PolyRegs<VA> A;
                    A.load(sA, 32, src);
                    sep = sat_segment<VA>(A, s_sg[k], s_sg[p.cap_s + k]);
                    if (COUNT) n_axis++;
                }
                const bool hit = have && sep < -tau;
                const int ci_src = __shfl_sync(FULL, ci, src);
                if (have && !hit && !(sep > tau)) push_work(p, q, ci_src, j, KIND_SEG, g0 + sc0 + k);
                const unsigned m = __reduce_or_sync(FULL, hit ? (1u << src) : 0u);
                collide |= (m >> lane) & 1u;
            };

            // ---- obstacles of this time step --------------------------------------------------------------------
            {
                const float4 *s_cx4 = reinterpret_cast<const float4 *>(s_c);
                const float4 *s_cy4 = reinterpret_cast<const float4 *>(s_c + p.cap_o);
                const float4 *s_cr4 = reinterpret_cast<const float4 *>(s_c + 2 * p.cap_o);
                const f32x2 nax = pack2(-acx, -acx), nay = pack2(-acy, -acy);
                // which sub-blocks of 8 slots can matter: lane b tests sub-blocks b and b + 32 (one load latency for the
                // whole tile).  A sub-block whose circles (box already inflated by their radii) cannot come within
                // rcull of any centre of this warp has every pair separated on the box axis.
                unsigned long long live_o = ~0ull;
                if (cull_o) {
                    const int nsb = ocn >> 3;                // <= 40 (cap_o <= 320)
                    const float4 *bx = p.ob_box + ((o0 + oc0) >> 3);
                    bool l0 = false, l1 = false;
                    if (lane < nsb) {
                        const float4 bb = __ldg(bx + lane);
                        l0 = !(bb.x > wx1 + rcull || bb.z < wx0 - rcull || bb.y > wy1 + rcull || bb.w < wy0 - rcull);
                    }
                    if (lane + 32 < nsb) {
                        const float4 bb = __ldg(bx + lane + 32);
                        l1 = !(bb.x > wx1 + rcull || bb.z < wx0 - rcull || bb.y > wy1 + rcull || bb.w < wy0 - rcull);
                    }
                    live_o = (unsigned long long)__ballot_sync(FULL, l0) | ((unsigned long long)__ballot_sync(FULL, l1) << 32);
                }
                for (int ob = wo * 32; ob < ocn; ob += 32 * WO) {
                    const unsigned m4 = (unsigned)(live_o >> (ob >> 3)) & 0xFu;
                    if (COUNT && active) {
#pragma unroll
                        for (int sub = 0; sub < 4; sub++) {
                            const int real = max(0, min(8, nobst - oc0 - ob - 8 * sub));
                            if ((m4 >> sub) & 1u) n_cull += real; else n_skip += real;
                        }
                    }
                    if (!m4) continue;
                    // axis 0: centre line with bounding radii.  Three warp-broadcast LDS.128 bring four obstacles; packed
                    // FADD2/FFMA2 evaluate e = |dc|^2 - (rb + rcull)^2 for two of them per instruction and the sign bit
                    // of e is shifted into the hit word (bit 31-k <-> slot ob+k).
                    unsigned hits = 0;
#pragma unroll
                    for (int sub = 0; sub < 4; sub++) {
                        if (!((m4 >> sub) & 1u)) { hits <<= 8; continue; }
#pragma unroll
                        for (int qd = 2 * sub; qd < 2 * sub + 2; qd++) {
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
                    }
                    if (!active || (collide && !noexit)) hits = 0;
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
                const float4 *boxes = p.sg_box + ((g0 + sc0) >> 3);      // one box per sub-block of 8 segment slots
                // the even-odd ray may run along +x, -x, +y or -y; take the direction whose ray corridor meets the
                // fewest sub-blocks of the whole range (a road aligned with the ray would otherwise keep all of them)
                int dir = 0;
                if (blockcull) {
                    const float4 *all = p.sg_box + (g0 >> 3);
                    const int nsb = (nseg + 7) >> 3;
                    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                    for (int b = lane; b < ((nsb + 31) & ~31); b += 32) {
                        float4 bb = (b < nsb) ? __ldg(all + b) : make_float4(BIG, BIG, -BIG, -BIG);
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
                // live sub-blocks of this phase (<= 64: cap_s <= 512), lane b tests sub-blocks b and b + 32
                unsigned long long live_s = ~0ull;
                if (blockcull) {
                    const int nsb = (scn + 7) >> 3;
                    auto test = [&](int b) -> bool {
                        if (b >= nsb) return false;
                        const float4 bb = __ldg(boxes + b);
                        // a segment can meet a polygon only if the boxes overlap; it can be crossed by the ray of a
                        // centre only if it reaches that centre's constant coordinate and is not entirely behind it
                        const bool sat_rel = bb.x <= wx1 + wr && bb.z >= wx0 - wr && bb.y <= wy1 + wr && bb.w >= wy0 - wr;
                        const bool iny = bb.w >= wy0 && bb.y <= wy1, inx = bb.z >= wx0 && bb.x <= wx1;
                        const bool par_rel = dir == 0 ? (iny && bb.z >= wx0) : dir == 1 ? (iny && bb.x <= wx1)
                                           : dir == 2 ? (inx && bb.w >= wy0) : (inx && bb.y <= wy1);
                        return sat_rel || par_rel;
                    };
                    const bool l0 = test(lane), l1 = test(lane + 32);
                    live_s = (unsigned long long)__ballot_sync(FULL, l0) | ((unsigned long long)__ballot_sync(FULL, l1) << 32);
                }
                // one segment against the lane's polygon centre: crossing parity of the centre's ray + "near" test
                auto seg_one = [&](int idx, int k, unsigned &hits) {
                    float4 sa = s_sg[idx], sl = s_sg[p.cap_s + idx];
                    float dl = fmaf(sl.x, acx, fmaf(sl.y, acy, -sl.z));       // centre to the segment's line
                    // even-odd crossing parity of the ray from the polygon centre.  The two comparisons are exact on
                    // the float32 data and consistent between segments that share a vertex; the side test is trusted
                    // only outside the error band
                    const float ea = ray_y ? sa.x : sa.y, eb = ray_y ? sa.z : sa.w;
                    bool ua = ea > pc, ub = eb > pc;
                    if (ua != ub) {
                        par ^= (ub != flip) ? (dl < 0.f) : (dl > 0.f);
                        punc |= fabsf(dl) <= tau;
                    }
                    float tt = fmaf(-sl.y, acx - 0.5f * (sa.x + sa.z), sl.x * (acy - 0.5f * (sa.y + sa.w)));
                    bool near = (fabsf(dl) <= arr) & (fabsf(tt) <= sl.w + arr);
                    if (near) hits |= 0x80000000u >> k;
                };
                for (int sb = wo * 32; sb < scn; sb += 32 * WO) {
                    const int cnt = min(32, scn - sb);
                    unsigned hits = 0;
                    if (!blockcull) {
#pragma unroll 2
                        for (int k = 0; k < cnt; k++) seg_one(sb + k, k, hits);
                        if (COUNT && active) n_cull += cnt;
                    } else {
                        const unsigned m4 = (unsigned)(live_s >> (sb >> 3)) & 0xFu;
                        if (COUNT && active) {
                            for (int sub = 0; sub < 4; sub++) {
                                const int real = max(0, min(8, cnt - 8 * sub));
                                if ((m4 >> sub) & 1u) n_cull += real; else n_skip += real;
                            }
                        }
                        if (!m4) continue;
                        for (int sub = 0; sub < 4 && 8 * sub < cnt; sub++) {
                            if (!((m4 >> sub) & 1u)) continue;
                            const int kend = min(8 * sub + 8, cnt);
#pragma unroll 2
                            for (int k = 8 * sub; k < kend; k++) seg_one(sb + k, k, hits);
                        }
                    }
                    if (!active || (collide && !noexit)) hits = 0;
                    push_hits(hits, sb, drain_seg);
                    if (qn >= 32) drain_seg();
                }
                while (qn) drain_seg();
            }

            // ---- fold this phase into the group state ----------------------------------------------------------
            unsigned bc = __ballot_sync(FULL, collide), bp = __ballot_sync(FULL, par), bu = __ballot_sync(FULL, punc);
            if (WS) {
                // several warps share the group: merge the partial results; published after the last phase (below)
                if (lane == 0) {
                    if (bc) atomicOr(s_coll + g, bc);
                    if (bp) atomicXor(s_par + g, bp);
                    if (bu) atomicOr(s_punc + g, bu);
                }
                const unsigned ba = __ballot_sync(FULL, active);
                if (lane == 0 && wo == 0 && ph == 0) s_act[g] = ba;
                continue;
            }
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
                // spatially sorted candidate sets: bits in sorted order, k_refine moves them home
                unsigned *out = (pos_base < 0 ? p.mask : p.mask_sorted) + Q->mask_off;
                if (cmode) {    // compacted group: every lane has its own word
                    if (coll && active) atomicOr(out + (ci >> 5), 1u << (ci & 31));
                } else {
                    unsigned word = __ballot_sync(FULL, coll && active);
                    if (lane == 0 && word) atomicOr(out + gbeg + g, word);
                }
            }
        }
        if (nphase > 1 || WS) __syncthreads();      // everyone is done with the tile before the next bulk copy overwrites it
    }
    if (WS) {
        // the merged state of every group (the barrier above closed the last phase): one warp per group publishes
        unsigned *out = (pos_base < 0 ? p.mask : p.mask_sorted) + Q->mask_off;
        for (int g = warp; g < ng; g += NW) {
            const unsigned bc = s_coll[g], bp = s_par[g], bu = s_punc[g], ba = s_act[g];
            const int ci = (gbeg + g) * 32 + lane;
            const bool active = (ba >> lane) & 1u;
            bool coll = (bc >> lane) & 1u;
            if (has_region && active && !coll) {
                if ((bu >> lane) & 1u) push_work(p, q, ci, j, KIND_PARITY, 0);
                else if (!((bp >> lane) & 1u)) coll = true;        // centre outside the region
            }
            const unsigned word = __ballot_sync(FULL, coll && active);
            if (lane == 0 && word) atomicOr(out + gbeg + g, word);
        }
    }
    };      // run_slot

    run_slot();
    nom = __reduce_add_sync(FULL, nom);
    if (lane == 0 && nom) atomicAdd(&p.cnt->pairs_nominal, (unsigned long long)nom);
    if (COUNT) {
        n_cull = __reduce_add_sync(FULL, n_cull);
        n_axis = __reduce_add_sync(FULL, n_axis);
        n_skip = __reduce_add_sync(FULL, n_skip);
        n_done = __reduce_add_sync(FULL, n_done);
        if (lane == 0) {
            atomicAdd(&p.cnt->done_skipped, (unsigned long long)n_done);
            atomicAdd(&p.cnt->cull_tests, (unsigned long long)n_cull);
            atomicAdd(&p.cnt->axis_tests, (unsigned long long)n_axis);
            atomicAdd(&p.cnt->box_skipped, (unsigned long long)n_skip);
        }
    }
    if (WS && p.tail) {
        // Single-kernel query: the CTA that finishes last runs the float64 / exact pass over the work list (a handful
        // of items for a planner-sized query) and writes the result block straight into page-locked host memory, then
        // raises a flag the host is polling -- no k_refine launch, no copy node, no stream synchronisation to wake up
        // from.
        unsigned *s_is_last = reinterpret_cast<unsigned *>(smem_raw + 12);      // (bytes 0..7 mbarrier, 8..11 s_tot)
        __threadfence();
        __syncthreads();
        if (tid == 0) *s_is_last = atomicAdd(&p.cnt->done_ctas, 1u) == gridDim.x * gridDim.y - 1u;
        __syncthreads();
        if (*s_is_last) {
            __threadfence();
            KParams *ps = reinterpret_cast<KParams *>(s_ob);                 // (tile and polygon areas are free now)
            for (int i = tid; i < (int)(sizeof(KParams) / 4); i += K_THREADS)
                reinterpret_cast<unsigned *>(ps)[i] = reinterpret_cast<const unsigned *>(&p)[i];
            __syncthreads();
            refine_pass_smem(ps, (unsigned)warp, (unsigned)NW, reinterpret_cast<double *>(s_A) + warp * REFINE_SCRATCH);
            __threadfence();
            __syncthreads();
            const unsigned *src = reinterpret_cast<const unsigned *>(p.cnt);
            for (int i = tid; i < p.res_words; i += K_THREADS) p.host_res[i] = __ldcg(src + i);
            __threadfence_system();
            __syncthreads();
            if (tid == 0) *p.host_flag = p.host_seq;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// full evaluation (MPC_MODE_NO_CULL): every (occupancy, obstacle) and (occupancy, segment) pair of the batch runs the
// complete separating-axis test -- no bounding-circle axis, no boxes, no queue, no early return across time slots.
// This is the arithmetic the path is made of, executed for every nominal pair: the regime in which algorithmic and
// executed work coincide and the FP32 roofline is the bound.
// Same grid, tile staging and result conventions as k_check.  A lane holds NP = 2 occupancy polygons (of two candidate
// groups) in registers, so every warp-broadcast obstacle record read from shared memory (5 LDS.128 for a rectangle) feeds
// two tests: with one polygon per lane the shared-memory data pipe, not the FMA pipe, was the binding unit (ncu capture
// r01c: l1tex data-pipe wavefronts 78 %).
// ---------------------------------------------------------------------------------------------------------------------
template <int VA, int VB, bool COUNT>
__global__ void __launch_bounds__(K_THREADS, 2) k_check_full(const KParams p) {
    constexpr int F = ob_fields(VB), FA = ob_fields(VA), NW = K_THREADS / 32, NP = (VA <= 4) ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *s_coll = reinterpret_cast<unsigned *>(smem_raw + 16);            // same carve-up as k_check (one size)
    unsigned *s_par = s_coll + GPC_MAX;
    unsigned *s_punc = s_par + GPC_MAX;
    float4 *s_ob = reinterpret_cast<float4 *>(s_punc + 2 * GPC_MAX + NW * QCAP) + NW * FA * 32;
    float4 *s_sg = reinterpret_cast<float4 *>(reinterpret_cast<float *>(s_ob + F * p.cap_o) + 3 * p.cap_o);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int j = __ldg(p.slot_perm + blockIdx.y);               // grid: (CTAs of one time slot, time slots), slot-major
    const int2 ent = __ldg(p.cta_tab + blockIdx.x);
    const int q = ent.x, chunk = ent.y;
    const QDev *Q = p.q + q;
    const int t = Q->t0 + p.slot_t[j];
    const float tau = fmaxf(fmaxf(*p.maxabs_scene, *p.maxabs_query) * (1.0f / 131072.0f), 1e-7f);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        pdl_wait();
        p.cnt->tau = tau;
    }
    if (t > Q->step_limit || t < 0 || t >= Q->nsteps) return;      // reference guards (:120,122)
    const int step = Q->step0 + t;
    const int o0 = p.step_obst_off[step], nslot = p.step_obst_off[step + 1] - o0, nobst = p.step_obst_cnt[step];
    const int g0 = p.step_seg_begin[step];
    const bool has_region = g0 >= 0;
    const int nseg = has_region ? p.step_seg_end[step] - g0 : 0;
    const int cand_cnt = Q->cand_cnt, cand_base = Q->cand_base, pos_base = Q->pos_base;
    const int gbeg = chunk * p.gpc;
    const int ng = min(p.gpc, (cand_cnt + 31) / 32 - gbeg);
    const float lx = Q->lx, ly = Q->ly, lc = Q->lc, ls = Q->ls;
    const bool noexit = p.mode & MPC_MODE_NO_EARLY_EXIT;
    if (tid == 0) mbar_init(bar, 1);
    for (int i = tid; i < 3 * GPC_MAX; i += K_THREADS) s_coll[i] = 0;
    __syncthreads();
    const int nph_o = (nslot + p.cap_o - 1) / p.cap_o, nph_s = (nseg + p.cap_s - 1) / p.cap_s;
    const int nphase = max(1, max(nph_o, nph_s));
    const bool direct = pos_base >= 0;
    const float4 *Lv = direct ? p.slib_v : p.lib_v, *Ln = direct ? p.slib_n : p.lib_n, *Lc = direct ? p.slib_c : p.lib_c;
    const size_t fstride = direct ? (size_t)p.J * p.set_total : (size_t)p.PJ;
    unsigned *out = (direct ? p.mask_sorted : p.mask) + Q->mask_off;
    unsigned nom = 0, n_axis = 0;

    for (int ph = 0; ph < nphase; ph++) {
        const int oc0 = ph * p.cap_o, ocn = max(0, min(nslot - oc0, p.cap_o));
        const int sc0 = ph * p.cap_s, scn = max(0, min(nseg - sc0, p.cap_s));
        const bool loading = (ocn | scn) != 0;
        if (tid == 0 && loading) {
            mbar_expect_tx(bar, (unsigned)(ocn * 16 * F + scn * 32));
            for (int f = 0; f < F && ocn; f++)
                tma_bulk_g2s(s_ob + f * p.cap_o, p.ob_f + (size_t)f * p.ob_stride + o0 + oc0, ocn * 16, bar);
            for (int f = 0; f < 2 && scn; f++)
                tma_bulk_g2s(s_sg + f * p.cap_s, p.sg_f + (size_t)f * p.sg_stride + g0 + sc0, scn * 16, bar);
        }
        if (loading) mbar_wait(bar, ph & 1);
        if (ph == 0) pdl_wait();
        const int kobst = min(ocn, nobst - oc0);                      // real obstacles of this phase
        for (int gp = warp * NP; gp < ng; gp += NW * NP) {
            PolyRegs<VA> A[NP];
            float acx[NP], acy[NP];
            int ci[NP];
            bool active[NP], collide[NP], par[NP], punc[NP];
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const int g = gp + u;
                ci[u] = (gbeg + g) * 32 + lane;
                active[u] = g < ng && ci[u] < cand_cnt;
                size_t pj = 0;
                if (active[u]) pj = direct ? (size_t)j * p.set_total + pos_base + ci[u] : (size_t)__ldg(p.cand + cand_base + ci[u]) * p.J + j;
                float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active[u]) cc = __ldg(Lc + pj);
                const int nv = (int)cc.w;
                active[u] = active[u] && nv > 0;
                if (active[u] && ph == 0) nom += (unsigned)(nobst + nseg);
                float ax[VA], ay[VA];
#pragma unroll
                for (int i = 0; i < ((VA * 3 + 3) / 4) * 4; i++) A[u].ln[i] = 0.f;
#pragma unroll
                for (int h = 0; h < VA / 2; h++) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
                    if (active[u]) { v = __ldg(Lv + h * fstride + pj); n = __ldg(Ln + h * fstride + pj); }
                    ax[2 * h] = fmaf(lc, v.x, fmaf(-ls, v.y, lx));
                    ay[2 * h] = fmaf(ls, v.x, fmaf(lc, v.y, ly));
                    ax[2 * h + 1] = fmaf(lc, v.z, fmaf(-ls, v.w, lx));
                    ay[2 * h + 1] = fmaf(ls, v.z, fmaf(lc, v.w, ly));
                    A[u].ln[6 * h] = fmaf(lc, n.x, -ls * n.y);
                    A[u].ln[6 * h + 1] = fmaf(ls, n.x, lc * n.y);
                    A[u].ln[6 * h + 3] = fmaf(lc, n.z, -ls * n.w);
                    A[u].ln[6 * h + 4] = fmaf(ls, n.z, lc * n.w);
                    A[u].X[h] = pack2(ax[2 * h], ax[2 * h + 1]);
                    A[u].Y[h] = pack2(ay[2 * h], ay[2 * h + 1]);
                }
#pragma unroll
                for (int e = 0; e < VA; e++)       // line offset; an edge beyond nv never separates
                    A[u].ln[3 * e + 2] = (e < nv) ? fmaf(A[u].ln[3 * e], ax[e], A[u].ln[3 * e + 1] * ay[e]) : BIG;
                acx[u] = active[u] ? fmaf(lc, cc.x, fmaf(-ls, cc.y, lx)) : -FAR;
                acy[u] = active[u] ? fmaf(ls, cc.x, fmaf(lc, cc.y, ly)) : -FAR;
                collide[u] = g < ng ? ((s_coll[g] >> lane) & 1u) : false;
                par[u] = punc[u] = false;
            }
            // ---- obstacles: one broadcast record, NP tests
            for (int k = 0; k < kobst; k++) {
                PolyRegs<VB> B;
                B.load(s_ob, p.cap_o, k);
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const float sep = sat_polys<VA, VB>(A[u], B);
                    if (active[u]) {
                        if (COUNT) n_axis++;
                        if (sep < -tau) collide[u] = true;
                        else if (!(sep > tau)) push_work(p, q, ci[u], j, KIND_OBST, o0 + oc0 + k);
                    }
                }
                if (!noexit && (k & 31) == 31) {
                    bool all_done = true;
#pragma unroll
                    for (int u = 0; u < NP; u++) all_done = all_done && (collide[u] || !active[u]);
                    if (__all_sync(FULL, all_done)) break;
                }
            }
            // ---- boundary segments: even-odd parity of the centre (ray along +x) + the segment against the polygon
            for (int k = 0; k < scn; k++) {
                const float4 sa = s_sg[k], sl = s_sg[p.cap_s + k];
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const float dl = fmaf(sl.x, acx[u], fmaf(sl.y, acy[u], -sl.z));
                    const bool ua = sa.y > acy[u], ub = sa.w > acy[u];
                    if (ua != ub) {
                        par[u] ^= ub ? (dl < 0.f) : (dl > 0.f);
                        punc[u] |= fabsf(dl) <= tau;
                    }
                    if (sl.w >= 0.f && active[u] && (noexit || !collide[u])) {
                        const float sep = sat_segment<VA>(A[u], sa, sl);
                        if (COUNT) n_axis++;
                        if (sep < -tau) collide[u] = true;
                        else if (!(sep > tau)) push_work(p, q, ci[u], j, KIND_SEG, g0 + sc0 + k);
                    }
                }
            }
            // ---- fold the phase into the group state, publish after the last one
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const int g = gp + u;
                if (g >= ng) continue;                                   // warp-uniform
                unsigned bc = __ballot_sync(FULL, collide[u]), bp = __ballot_sync(FULL, par[u]), bu = __ballot_sync(FULL, punc[u]);
                if (nphase > 1) {
                    if (lane == 0) { s_coll[g] |= bc; s_par[g] ^= bp; s_punc[g] |= bu; }
                    __syncwarp();
                    bc = s_coll[g]; bp = s_par[g]; bu = s_punc[g];
                }
                if (ph == nphase - 1) {
                    bool coll = (bc >> lane) & 1u;
                    if (has_region && active[u] && !coll) {
                        if ((bu >> lane) & 1u) push_work(p, q, ci[u], j, KIND_PARITY, 0);
                        else if (!((bp >> lane) & 1u)) coll = true;        // centre outside the region
                    }
                    const unsigned word = __ballot_sync(FULL, coll && active[u]);
                    if (lane == 0 && word) atomicOr(out + gbeg + g, word);
                }
            }
        }
        if (nphase > 1) __syncthreads();
    }
    nom = __reduce_add_sync(FULL, nom);
    if (lane == 0 && nom) atomicAdd(&p.cnt->pairs_nominal, (unsigned long long)nom);
    if (COUNT) {
        n_axis = __reduce_add_sync(FULL, n_axis);
        if (lane == 0) atomicAdd(&p.cnt->axis_tests, (unsigned long long)n_axis);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// float64 / exact refinement of the pairs inside the float32 band
// ---------------------------------------------------------------------------------------------------------------------
// One WARP per work-list item: the orientation tests of an item (up to nA*nB of them per direction) are independent,
// so the lanes split them and votes put the answer together.  A single thread needs ~20 k cycles per polygon pair
// (dependent float64 chains, local-memory polygon arrays); a warp needs ~1-2 k, which is what a planner that waits on
// every query cares about.
// `warp0` / `warps`: this warp's index among all warps that share the pass and their number; `pts`: scratch of
// REFINE_SCRATCH doubles per warp in shared memory.  Called by k_refine (whole grid) and by the last CTA of a
// single-kernel query (k_check<..., WS>).
__device__ __forceinline__ void refine_pass(const KParams &p, const unsigned warp0, const unsigned warps, double *pts) {
    const unsigned n = min(__ldcg(&p.cnt->wl_count), (unsigned)p.wl_cap);
    const unsigned lane = threadIdx.x & 31;
    double *Ax = pts, *Ay = Ax + MPC_MAX_LIB_VERTS, *Bx = Ay + MPC_MAX_LIB_VERTS, *By = Bx + MPC_MAX_OBST_VERTS;
    unsigned refined = 0, ties = 0, near_t = 0, par_ref = 0;
    const bool strict = p.mode & MPC_MODE_VALIDATOR;
    // phase A: result bits of spatially sorted candidate sets go home to the caller's candidate order.  One warp per
    // destination mask word: lane = candidate, which looks up its position in the sorted set and fetches its bit (a
    // gather: no atomics, no serial chain per word); one OR per word merges with the bits phase B adds below.
    // (query, word) pairs are enumerated arithmetically -- B x wmax slots, the ones beyond a query's own words are empty --
    // instead of searching the query of every word; batches without sorted sets skip the phase
    const unsigned long long slots = p.any_sorted ? (unsigned long long)p.B * (unsigned)p.wmax : 0ull;
    for (unsigned long long idx = warp0; idx < slots; idx += warps) {
        const int qi = (int)(idx / (unsigned)p.wmax), wq = (int)(idx - (unsigned long long)qi * (unsigned)p.wmax);
        const QDev *Q = p.q + qi;
        if (Q->pos_base < 0 || wq * 32 >= Q->cand_cnt) continue;          // caller's order already / no such word
        const unsigned w = (unsigned)(Q->mask_off + wq);
        const int ci = wq * 32 + (int)lane;
        bool bit = false;
        if (ci < Q->cand_cnt) {
            const int sp = __ldg(p.cand_inv + Q->pos_base + ci);          // position of candidate ci in the sorted set
            bit = (__ldcg(p.mask_sorted + Q->mask_off + (sp >> 5)) >> (sp & 31)) & 1u;
        }
        const unsigned word = __ballot_sync(FULL, bit);
        if (lane == 0 && word) atomicOr(p.mask + w, word);
    }
    for (unsigned item = warp0; item < n; item += warps) {
        const int4 e = __ldcg(p.wl + item);
        const int q = e.x, ci = e.y, j = e.z & 0xffff, kind = e.z >> 16, obj = e.w;
        const QDev *Q = p.q + q;
        const int pj = p.cand[Q->cand_base + ci] * p.J + j;
        const int nA = p.lib_nv[pj];
        MPC_ASSERT(q >= 0 && q < p.B && ci >= 0 && ci < Q->cand_cnt && j < p.J && nA >= 3 && nA <= MPC_MAX_LIB_VERTS);
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
            MPC_ASSERT(obj >= 0 && obj < p.ob_stride && nB >= 3 && nB <= MPC_MAX_OBST_VERTS);
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
            // near-tangent statistic: |separation| < 1e-9 m with separation = max over edges of (min signed distance of
            // the other polygon's vertices).  One edge per lane, compared in squared form (no sqrt, no division):
            //   s_e < +eps  <=>  mn_e < 0  or  mn_e^2 < eps^2 |e|^2 ;   s_e > -eps  <=>  mn_e > 0  or  mn_e^2 < eps^2 |e|^2
            bool below = true, above = false;      // all edges below +eps / some edge above -eps
            if ((int)lane < nA + nB) {
                const bool fromA = (int)lane < nA;
                const int ee = fromA ? lane : lane - nA, nn = fromA ? nA : nB, e1 = (ee + 1 == nn) ? 0 : ee + 1;
                const double *Px = fromA ? Ax : Bx, *Py = fromA ? Ay : By, *Qx = fromA ? Bx : Ax, *Qy = fromA ? By : Ay;
                const int nq = fromA ? nB : nA;
                const double ex = Px[e1] - Px[ee], ey = Py[e1] - Py[ee], len2 = ex * ex + ey * ey;
                if (len2 > 0.0) {
                    double mn = 1e300;
                    for (int k = 0; k < nq; k++) mn = fmin(mn, ey * (Qx[k] - Px[ee]) - ex * (Qy[k] - Py[ee]));
                    const bool tiny = mn * mn < 1e-18 * len2;
                    below = mn < 0.0 || tiny;
                    above = mn > 0.0 || tiny;
                }
            }
            const bool near = __all_sync(FULL, below) && __any_sync(FULL, above);
            if (lane == 0) { if (near) near_t++; refined++; }
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

// out-of-line copy for the tail of k_check (its 64-register budget stays untouched).  The parameter block is handed over
// in shared memory: a reference to the kernel parameters themselves would make the compiler copy all 1.3 KB of them
// to every thread's stack.
__device__ __noinline__ void refine_pass_smem(const KParams *ps, const unsigned warp0, const unsigned warps, double *pts) {
    refine_pass(*ps, warp0, warps, pts);
}

__global__ void __launch_bounds__(128, 8) k_refine(const KParams p) {
    __shared__ double s_pts[4][REFINE_SCRATCH];
    pdl_wait();                  // everything here reads what k_check wrote
    refine_pass(p, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, (gridDim.x * blockDim.x) >> 5, s_pts[threadIdx.x >> 5]);
}

// clears the result block (counters + both mask arrays) at the head of a step; a kernel rather than a memset node so that
// the first k_check launch can be a programmatic dependent of it
__global__ void k_clear(uint4 *p, int n16) {
    pdl_trigger();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// register-only arithmetic chains for mpc_measure_fp32_peak.  One round = NF scalar FFMA + NF2 packed FFMA2 + N3
// three-input min / max + N2 two-input min / max on 8 independent accumulators; 8 rounds per loop iteration.
template <int NF, int NF2, int N3, int N2>
__global__ void __launch_bounds__(256) k_peak(float *out, int iters, float b, float c, long long *cycles) {
    const long long t0 = clock64();
    float a[8], mn = BIG, mx = -BIG;
    f32x2 a2[8];
    const f32x2 b2 = pack2(b, b + 1e-7f), c2 = pack2(c, c - 1e-7f);
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = (float)(threadIdx.x + i); a2[i] = pack2((float)(threadIdx.x + i), (float)i); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < NF; i++) a[i & 7] = fmaf(a[i & 7], b, c);
#pragma unroll
            for (int i = 0; i < NF2; i++) a2[i & 7] = fma2(a2[i & 7], b2, c2);
            const float *src = NF2 ? reinterpret_cast<const float *>(a2) : a;      // (registers: the casts fold away)
#pragma unroll
            for (int i = 0; i < N3; i++) {
                if (i & 1) mx = max3(mx, NF2 ? lo32(a2[(2 * i) & 7]) : a[(2 * i) & 7], NF2 ? hi32(a2[(2 * i + 1) & 7]) : a[(2 * i + 1) & 7]);
                else mn = min3(mn, NF2 ? lo32(a2[(2 * i) & 7]) : a[(2 * i) & 7], NF2 ? hi32(a2[(2 * i + 1) & 7]) : a[(2 * i + 1) & 7]);
            }
#pragma unroll
            for (int i = 0; i < N2; i++) {
                if (i & 1) mx = fmaxf(mx, NF2 ? hi32(a2[(i + 3) & 7]) : a[(i + 3) & 7]);
                else mn = fminf(mn, NF2 ? lo32(a2[(i + 5) & 7]) : a[(i + 5) & 7]);
            }
            (void)src;
        }
    }
    float keep = mn + mx;
#pragma unroll
    for (int i = 0; i < 8; i++) keep += a[i] + lo32(a2[i]) + hi32(a2[i]);
    if (keep == 12345.678f) out[0] = keep;              // never true in practice: keeps the chains alive
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void k_fill_u32(unsigned *p, size_t n, unsigned v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace mpc
