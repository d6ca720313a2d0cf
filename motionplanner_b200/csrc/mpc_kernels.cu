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
//   k_check<VA,VB,COUNT,SEGCULL,WS>   one CTA (8 warps) per (query, occupancy slot, chunk of candidate groups).  The
//                                 obstacle and boundary-segment records of that time step are brought into shared
//                                 memory with one TMA bulk copy per field (cp.async.bulk + mbarrier).  A warp works on
//                                 32 candidate primitives at a time, lane = primitive: it transforms the occupancy
//                                 polygons, parks them in shared memory and streams the step's obstacles as
//                                 warp-broadcast loads.  Axis 0 (centre line, bounding radii) runs on packed
//                                 FADD2/FFMA2, two obstacles per instruction; the pairs it cannot separate are queued
//                                 per warp and the full separating-axis test runs on 32 queued pairs at once.  A tile
//                                 ends as soon as __all_sync says every lane has collided; __ballot_sync assembles the
//                                 32 result bits that are OR-ed into the packed mask.  WS (planner-sized batches):
//                                 several warps per candidate group, each with a share of the tile; the last CTA to
//                                 finish refines the work list and hands the result block to the host.
//   k_check_full<VA,VB,COUNT>     MPC_MODE_NO_CULL: every nominal pair runs the complete test, two polygons per lane.
//   k_refine                      float64 / exact re-decision of the pairs inside the float32 error band, one warp per
//                                 pair (lanes over the orientation tests).
// A step = k_clear -> k_check (scout launch + main launch for big batches) -> k_refine, chained as programmatic
// dependent launches; mpc_query replays (H2D, the chain, D2H) as one instantiated CUDA graph per batch shape,
// planner-sized batches as (H2D, one k_check<..., WS> launch) with the host polling a page-locked flag;
// mpc_submit / mpc_wait enqueue and collect a cycle (scene staging + query) without blocking in between.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mpc.h"
#include "mpc_device.cuh"
#include "mpc_expand.cuh"

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
    bool have_lib = false, lib_full = false;      // lib_full: every (primitive, slot) holds an occupancy
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
    std::vector<const void *> known_pinned;       // host addresses cudaPointerGetAttributes reported as page-locked
    std::vector<double> origins;
    DevBuf ob_src, ob_src_nv, ob_src_off, ob_c, ob_box, ob_order, step_obst_cnt, sets_pos;
    int ob_stride = 32;
    DevBuf ob_in, ob_f, ob_nv, step_obst_off, step_seg_begin, step_seg_end, scene_step_off_d, origins_d, sg_64, sg_f,
        sg_box, seg_scene, maxabs;
    int sg_stride = 32;
    DevBuf sg_raw, sg_jobs;       // large scene sets: caller's segments as they are + the range table, sorted on the device
    DevBuf meta;                  // one allocation behind the small per-scene arrays (views below), one H2D copy
    void *h_meta = nullptr;
    size_t h_meta_cap = 0;
    long long meta_sig = 0;
    size_t meta_flag_off = 0;     // where in h_meta the staging kernels' report lands (count of non-convex pieces)
    bool pending = false, pending_scenes = false, pending_timed = false;     // mpc_submit without its mpc_wait yet
    // queries
    DevBuf qpack_d, cand_d, res_d, wl, flush;
    void *res_p = nullptr;        // result block of the prepared batch: inside qpack_d (fused_clear) or res_d
    bool fused_clear = false;     // small result blocks are cleared by the zeros the query upload carries
    size_t qpack_cta_off = 0, qpack_cta2_off = 0, qpack_hdr_off = 0, up_bytes = 0, up_cand_bytes = 0, up_cand_src = 0;
    int wmax = 1, any_sorted = 0; // most result words of one query of the batch; some query uses a sorted candidate set
    int wsplit = 1;               // warps per candidate group in k_check (few groups per CTA)
    bool tail = false;            // prepared batch runs as ONE kernel: refinement + result hand-over in the last CTA of k_check
    volatile unsigned *h_flag = nullptr;      // page-locked word the tail raises
    int scouts = 0, scout_gpc = 1, scout_c_total = 0;      // leading time slots launched on their own (see launch_check)
    unsigned long long generation = 0;          // bumped whenever a pointer a captured graph holds may have changed
    std::map<std::vector<long long>, cudaGraphExec_t> graphs;
    std::map<std::vector<long long>, cudaGraphExec_t> scene_graphs;      // staging of small scenes from pinned arrays
    void *h_res = nullptr;
    size_t h_res_cap = 0;
    int wl_cap = 0;
    void *h_pin = nullptr;
    size_t h_pin_cap = 0;
    // prepared batch
    bool prepared = false;
    int B = 0, ctas = 0, mask_words = 0, gpc = 1, cap_o = 32, cap_s = 32, threads = 128;
    size_t smem = 0;
    unsigned mode = 0;
    float maxabs_query = 0.f;
    // primitive-selection step (mpc_expand): primitive trajectories, reference trajectory + goal sets
    bool have_traj = false, have_guidance = false;
    int tr_P = 0, tr_L = 0, g_M = 0, g_ngoals = 0, g_nseg = 0;
    double g_b = 0.0;
    DevBuf tr_states, tr_n, g_ref, g_goals, g_segs, e_pack, e_out;
    void *h_exp = nullptr;
    size_t h_exp_cap = 0;
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
static cudaError_t raise_smem_limits(const cudaDeviceProp &prop);

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
    for (auto &kv : ctx->scene_graphs) cudaGraphExecDestroy(kv.second);
    ctx->scene_graphs.clear();
    ctx->generation++;
}

static cudaError_t ensure_pinned(mpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_pin_cap) return cudaSuccess;
    invalidate_graphs(ctx);
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
    CU(raise_smem_limits(ctx->prop));     // once, here: the attribute cannot be changed while a graph is being captured
    for (auto &ev : ctx->ev) CU(cudaEventCreate(&ev));
    CU(ensure(ctx, ctx->res_d, sizeof(Counters) + 4096));
    CU(ensure(ctx, ctx->maxabs, 2 * sizeof(float)));       // [largest |coordinate| of the scenes, count of non-convex obstacle pieces]
    CU(cudaMemsetAsync(ctx->maxabs.p, 0, 2 * sizeof(float), ctx->stream));
    CU(cudaMallocHost((void **)&ctx->h_flag, 64));
    *ctx->h_flag = 0u;
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
                      &ctx->flush, &ctx->sg_raw, &ctx->sg_jobs, &ctx->tr_states, &ctx->tr_n, &ctx->g_ref, &ctx->g_goals, &ctx->g_segs, &ctx->e_pack,
                      &ctx->e_out};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    for (auto &kv : ctx->scene_graphs) cudaGraphExecDestroy(kv.second);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
    if (ctx->h_exp) cudaFreeHost(ctx->h_exp);
    if (ctx->h_flag) cudaFreeHost((void *)ctx->h_flag);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int mpc_device_info(mpc_ctx *ctx, mpc_devinfo *out) {
    if (!ctx || !out) return MPC_E_ARG;
    if (!ctx->stream) return fail(ctx, MPC_E_STATE, "context has no device");
    memset(out, 0, sizeof(*out));
    memcpy(out->name, ctx->prop.name, sizeof(out->name) - 1);     // device names are short; the field stays NUL-terminated
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
    if (ptr) {
        ctx->known_pinned.erase(std::remove(ctx->known_pinned.begin(), ctx->known_pinned.end(), (const void *)ptr), ctx->known_pinned.end());
        CU(cudaFreeHost(ptr));
    }
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
            // repeated consecutive vertices (zero-length edges) are dropped: such an edge has no normal, its axis would
            // report separation 0 and send every pair no other axis separates by more than tau to the tie-break pass
            int m = 0;
            for (int k = 0; k < nv; k++) {
                const double vx = src[2 * k], vy = src[2 * k + 1];
                if (m > 0 && x[m - 1] == vx && y[m - 1] == vy) continue;
                x[m] = vx; y[m] = vy; m++;
            }
            while (m > 1 && x[m - 1] == x[0] && y[m - 1] == y[0]) m--;
            nv = m >= 3 ? m : 0;
            double area = 0.0;
            for (int k = 0; k < nv; k++) {
                int k1 = (k + 1 == nv) ? 0 : k + 1;
                area += (x[k] - x[0]) * (y[k1] - y[0]) - (x[k1] - x[0]) * (y[k] - y[0]);
            }
            if (area == 0.0) nv = 0;       // degenerate occupancy: nothing to contain
            if (area < 0.0) { std::reverse(x, x + nv); std::reverse(y, y + nv); }
            // the separating-axis test is only correct for convex polygons: a reflex edge "separates" what it should
            // not, an unsafe false negative.  Every turn must be to the left and the turns must add up to one revolution
            // (a self-wrapping ring turns left all the way, twice).
            double turn = 0.0;
            for (int k = 0; k < nv; k++) {
                const int k1 = (k + 1 == nv) ? 0 : k + 1, k2 = (k1 + 1 == nv) ? 0 : k1 + 1;
                const double ex = x[k1] - x[k], ey = y[k1] - y[k], fx = x[k2] - x[k1], fy = y[k2] - y[k1];
                const double cr = ex * fy - ey * fx, dt = ex * fx + ey * fy;
                if (cr < -1e-12 * (std::fabs(ex * fy) + std::fabs(ey * fx)) - 1e-300)
                    return fail(ctx, MPC_E_ARG, "occupancy %zu (primitive %zu, slot %zu) is not convex at vertex %d: decompose it into convex pieces", i, i / J, i % J, k1);
                turn += std::atan2(cr, dt);
            }
            if (nv > 0 && std::fabs(turn - 6.283185307179586) > 1e-6)
                return fail(ctx, MPC_E_ARG, "occupancy %zu (primitive %zu, slot %zu) is not a simple convex ring (total turning %.3f rad)", i, i / J, i % J, turn);
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
    CU(ensure(ctx, ctx->lib_v, hv.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_n, hn.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_c, hc.size() * sizeof(float4)));
    CU(ensure(ctx, ctx->lib_v64, h64.size() * sizeof(double)));
    CU(ensure(ctx, ctx->lib_nv, hnv.size() * sizeof(int)));
    // Order in which the time slots are launched (slot-major grid).  Any order gives the same bits; what matters is how
    // soon candidates get decided, because decided candidates are skipped at all other slots.  Violations become more
    // likely with time (a primitive that has left the road stays off it, traffic ahead is reached later), so the
    // latest slots go first, interleaved with the earliest: latest, earliest, second latest, ...  Measured on the C4
    // batch: forward 0.185 ms, reverse 0.149 ms, interleaved 0.145 ms, middle-out 0.166 ms.  MPC_SLOT_ORDER=0/1/3 select
    // forward / reverse / middle-out for experiments.
    {
        std::vector<int> perm(J);
        const char *e = getenv("MPC_SLOT_ORDER");
        const int kind = e ? atoi(e) : 4;
        for (int l = 0; l < J; l++) perm[l] = l;
        if (kind == 1) std::reverse(perm.begin(), perm.end());
        if (kind == 2)
            for (int l = 0; l < J; l++) perm[l] = (l & 1) ? l / 2 : J - 1 - l / 2;
        if (kind == 3)
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return std::abs(2 * a - J) < std::abs(2 * b - J); });
        if (kind == 4) {        // latest first, then golden-ratio steps through the horizon: every prefix samples it evenly
            std::vector<char> used(J, 0);
            int n = 0;
            double x = 0.0;
            for (int it = 0; it < 64 * J && n < J; it++, x = std::fmod(x + 0.6180339887498949, 1.0)) {
                const int k = J - 1 - std::min(J - 1, (int)(x * J));
                if (!used[k]) { used[k] = 1; perm[n++] = k; }
            }
            for (int k = J - 1; k >= 0; k--)
                if (!used[k]) perm[n++] = k;
        }
        slot_t.insert(slot_t.end(), perm.begin(), perm.end());
    }
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
        std::vector<int> both((size_t)2 * std::max(total, 1)), pos((size_t)2 * std::max(total, 1));     // pos: [sorted -> caller's position | inverse]
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
                pos[(size_t)total + b + order[i].second] = i;
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
            CU(cudaMemcpyAsync(ctx->sets_pos.p, pos.data(), (size_t)2 * total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->P = P; ctx->J = J; ctx->lib_vmax = VA; ctx->VA = VA; ctx->lib_radius = radius; ctx->lib_rmax = rmax;
    ctx->lib_full = std::all_of(hnv.begin(), hnv.end(), [](int n) { return n > 0; });
    invalidate_graphs(ctx);      // captured launches bake in what depends on the library (compaction flag, slot order, radii)
    ctx->have_lib = true;
    ctx->prepared = false;
    ctx->arena_valid = false;
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
static int set_scenes_impl(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                           const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                           const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg,
                           bool sync);
static int check_scene_report(mpc_ctx *ctx);

extern "C" int mpc_set_scenes(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                              const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                              const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (ctx->pending) return fail(ctx, MPC_E_STATE, "a submitted batch is waiting for mpc_wait");
    const int rc = set_scenes_impl(ctx, nscenes, scene_step_off, origins, step_obst_off, obst, obst_nv, Vo, step_seg_begin,
                                   step_seg_end, segs, nseg, true);
    // the obstacle copy is started before the host-side passes: never hand the caller's buffers back with it in flight
    if (rc != MPC_OK) cudaStreamSynchronize(ctx->stream);
    return rc;
}

static int set_scenes_impl(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                           const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                           const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg,
                           bool sync) {
    if (nscenes <= 0 || !scene_step_off || !origins || !step_obst_off || !step_seg_begin || !step_seg_end)
        return fail(ctx, MPC_E_ARG, "bad scene arguments");
    CU(cudaSetDevice(ctx->device));
    const bool dbg_t = getenv("MPC_DEBUG_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (dbg_t) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[mpc set_scenes] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; } };
    const int nsteps = scene_step_off[nscenes];
    if (scene_step_off[0] != 0 || nsteps < 0) return fail(ctx, MPC_E_ARG, "scene_step_off must start at 0");
    for (int s = 0; s < nscenes; s++)
        if (scene_step_off[s + 1] < scene_step_off[s]) return fail(ctx, MPC_E_ARG, "scene_step_off not monotone");
    const int nobst = step_obst_off[nsteps];
    if (step_obst_off[0] != 0) return fail(ctx, MPC_E_ARG, "step_obst_off must start at 0");
    if (nobst > 0 && (!obst || !obst_nv)) return fail(ctx, MPC_E_ARG, "obstacle arrays missing");
    if (nobst > 0 && (Vo < 3 || Vo > MPC_MAX_OBST_VERTS)) return fail(ctx, MPC_E_LIMIT, "Vo=%d outside 3..%d", Vo, MPC_MAX_OBST_VERTS);
    if (nseg > 0 && !segs) return fail(ctx, MPC_E_ARG, "segment array missing");
    cudaStream_t st = ctx->stream;
    // large segment sets are sorted on the device
    const bool seg_dev_wanted = nseg >= 8192;
    // A small scene in page-locked arrays (a planner re-staging its prediction every cycle) is staged by replaying a
    // captured graph: one driver call instead of eight.  Everything the nodes hold is in the key below.
    // (the answer is remembered per address: the query costs 1-3 us, twice per call.  A stale "page-locked" for memory
    // that was freed and came back pageable only makes the replayed copy node slower, not wrong)
    auto pinned = [ctx](const void *ptr) {
        for (const void *known : ctx->known_pinned)
            if (known == ptr) return true;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
        const bool is = a.type == cudaMemoryTypeHost;
        if (is) {
            if (ctx->known_pinned.size() >= 64) ctx->known_pinned.clear();
            ctx->known_pinned.push_back(ptr);
        }
        return is;
    };
    static const bool no_scene_graph = getenv("MPC_NO_SCENE_GRAPH") != nullptr;
    const bool use_graph = !no_scene_graph && !seg_dev_wanted && (size_t)nobst * Vo * 2 * sizeof(double) <= ((size_t)4 << 20) &&
                           (nobst == 0 || (pinned(obst) && pinned(obst_nv)));
    // the bulk of the bytes first: the caller's obstacle rows travel as they are (one contiguous DMA, read by the staging
    // kernels with stride Vo) while the host validates and sorts the rest
    CU(ensure(ctx, ctx->ob_src, (size_t)std::max(nobst, 1) * Vo * 2 * sizeof(double)));
    CU(ensure(ctx, ctx->ob_src_nv, (size_t)std::max(nobst, 1) * sizeof(int)));
    auto copy_obstacles = [&]() -> int {
        if (nobst) {
            CU(cudaMemcpyAsync(ctx->ob_src.p, obst, (size_t)nobst * Vo * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(ctx->ob_src_nv.p, obst_nv, (size_t)nobst * sizeof(int), cudaMemcpyHostToDevice, st));
        }
        return MPC_OK;
    };
    if (!use_graph) {
        const int rc = copy_obstacles();
        if (rc) return rc;
    }
    if (seg_dev_wanted) {
        CU(ensure(ctx, ctx->sg_raw, (size_t)nseg * 4 * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->sg_raw.p, segs, (size_t)nseg * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    lap("start obstacle copy");

    // ---- boundary segments: every distinct [begin, end) range becomes a 32-aligned, spatially sorted (Morton order of
    // the midpoints) block range on the device, so that k_check can skip whole blocks by their bounding box
    int max_o = 0, max_s = 0;
    struct SegJob { int b, e, scene, start; };
    std::vector<SegJob> jobs;                                         // one per distinct range (pass 2 sorts them)
    std::map<std::pair<int, int>, int> ranges;                        // (begin, end) -> job
    std::vector<int> dev_begin(std::max(nsteps, 1), -1), dev_end(std::max(nsteps, 1), -1);
    int nslots_acc = 0;
    for (int s = 0; s < nscenes; s++) {
        int last_b = -2, last_e = -2, last_job = -1;                  // consecutive steps usually share their range
        for (int k = scene_step_off[s]; k < scene_step_off[s + 1]; k++) {
            int no = step_obst_off[k + 1] - step_obst_off[k];
            if (no < 0) return fail(ctx, MPC_E_ARG, "step_obst_off not monotone at step %d", k);
            max_o = std::max(max_o, no);
            int b = step_seg_begin[k], e = step_seg_end[k];
            if (b < 0) continue;
            if (e < b || e > nseg) return fail(ctx, MPC_E_ARG, "segment range of step %d outside 0..%d", k, nseg);
            max_s = std::max(max_s, e - b);
            int job = -1;
            if (b == last_b && e == last_e) job = last_job;
            else {
                auto it = ranges.find({b, e});
                if (it != ranges.end() && jobs[it->second].scene == s) job = it->second;
                else {                                                // new range (or same range under another origin)
                    job = (int)jobs.size();
                    jobs.push_back({b, e, s, nslots_acc});
                    nslots_acc += (e - b + 31) / 32 * 32;
                    ranges[{b, e}] = job;
                }
            }
            last_b = b; last_e = e; last_job = job;
            dev_begin[k] = jobs[job].start;
            dev_end[k] = jobs[job].start + (e - b);
        }
    }
    lap("pass 1 (ranges)");
    const int nslots = nslots_acc;
    int vmax_used = 3, vmin_used = 0;
#pragma omp parallel for reduction(max : vmax_used) reduction(min : vmin_used) if (nobst > 65536)
    for (int o = 0; o < nobst; o++) {
        vmax_used = std::max(vmax_used, (int)obst_nv[o]);
        vmin_used = std::min(vmin_used, (int)obst_nv[o]);
    }
    if (vmax_used > Vo || vmin_used < 0) return fail(ctx, MPC_E_ARG, "obst_nv outside 0..%d (min %d, max %d)", Vo, vmin_used, vmax_used);
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
    CU(ensure(ctx, ctx->ob_in, (size_t)ob_stride * VB * 2 * sizeof(double)));
    CU(ensure(ctx, ctx->ob_nv, (size_t)ob_stride * sizeof(int)));
    CU(ensure(ctx, ctx->ob_f, (size_t)ob_stride * F * sizeof(float4)));
    CU(ensure(ctx, ctx->ob_c, (size_t)ob_stride * 3 * sizeof(float)));
    CU(ensure(ctx, ctx->ob_box, (size_t)(ob_stride / 8 + 1) * sizeof(float4)));
    CU(ensure(ctx, ctx->ob_order, (size_t)ob_stride * sizeof(int)));
    CU(ensure(ctx, ctx->sg_f, (size_t)sg_stride * 2 * sizeof(float4)));
    CU(ensure(ctx, ctx->sg_box, (size_t)(sg_stride / 8 + 1) * sizeof(float4)));
    lap("validate + ensure");
    // the small per-scene arrays travel as one packed block (one driver call instead of nine)
    size_t meta_bytes = 0, m_org = 0, m_sg = 0, m_sso = 0, m_ss = 0, m_flag = 0;
    bool use_seg_dev = false;
    {
        const size_t n_org = (size_t)nscenes * 2 * sizeof(double), n_sg = (size_t)sg_stride * 4 * sizeof(double);
        const size_t n_sso = (size_t)(nscenes + 1) * 4, n_st1 = (size_t)(nsteps + 1) * 4, n_st = (size_t)std::max(nsteps, 1) * 4;
        const size_t n_ss = (size_t)sg_stride * 4;
        size_t off = 0;
        auto take = [&](size_t n) { size_t o = off; off += (n + 15) / 16 * 16; return o; };
        const size_t o_org = take(n_org), o_sg = take(n_sg), o_sso = take(n_sso), o_slot = take(n_st1), o_src = take(n_st1),
                     o_cnt = take(n_st), o_sb = take(n_st), o_se = take(n_st), o_ss = take(n_ss);
        const size_t o_flag = take(16);                    // host side only: what the staging kernels report back
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
        // pass 2: sort every distinct range along a Morton curve of the segment midpoints -- on the host straight into the
        // block for small sets, by k_order_segs for large ones (2000 scenes: 1.2 ms of host time otherwise)
        bool seg_dev = seg_dev_wanted;
        for (const SegJob &J : jobs) seg_dev = seg_dev && (J.e - J.b) <= SEG_SORT_MAX;
        if (!seg_dev) {
            double *dsegs = (double *)(h + o_sg);
            int *dscene = (int *)(h + o_ss);
            const int njobs = (int)jobs.size();
#pragma omp parallel for schedule(dynamic, 4) if (njobs > 8)
            for (int jb = 0; jb < njobs; jb++) {
                const SegJob &J = jobs[jb];
                const int n = J.e - J.b, padded = (n + 31) / 32 * 32;
                std::vector<std::pair<unsigned, int>> order(n);
                double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
                for (int g = J.b; g < J.e; g++) {
                    double mx = 0.5 * (segs[4 * g] + segs[4 * g + 2]), my = 0.5 * (segs[4 * g + 1] + segs[4 * g + 3]);
                    x0 = std::min(x0, mx); x1 = std::max(x1, mx); y0 = std::min(y0, my); y1 = std::max(y1, my);
                }
                const double sx = (x1 > x0) ? 65535.0 / (x1 - x0) : 0.0, sy = (y1 > y0) ? 65535.0 / (y1 - y0) : 0.0;
                for (int g = J.b; g < J.e; g++) {
                    double mx = 0.5 * (segs[4 * g] + segs[4 * g + 2]), my = 0.5 * (segs[4 * g + 1] + segs[4 * g + 3]);
                    order[g - J.b] = {morton16((unsigned)((mx - x0) * sx), (unsigned)((my - y0) * sy)), g};
                }
                std::sort(order.begin(), order.end());
                for (int i = 0; i < padded; i++) {
                    double *dst = dsegs + (size_t)(J.start + i) * 4;
                    if (i < n) {
                        const double *src = segs + 4 * (size_t)order[i].second;
                        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                        dscene[J.start + i] = J.scene;
                    } else {
                        dst[0] = dst[1] = dst[2] = dst[3] = 0.0;
                        dscene[J.start + i] = -1;
                    }
                }
            }
        }
        memcpy(h + o_sso, scene_step_off, n_sso);
        memcpy(h + o_slot, slot_off.data(), n_st1);
        memcpy(h + o_src, step_obst_off, n_st1);
        if (nsteps) {
            memcpy(h + o_cnt, step_cnt.data(), (size_t)nsteps * 4);
            memcpy(h + o_sb, dev_begin.data(), (size_t)nsteps * 4);
            memcpy(h + o_se, dev_end.data(), (size_t)nsteps * 4);
        }
        lap("pack meta (pass 2)");
        ctx->origins_d.p = d + o_org; ctx->sg_64.p = d + o_sg; ctx->scene_step_off_d.p = d + o_sso;
        ctx->step_obst_off.p = d + o_slot; ctx->ob_src_off.p = d + o_src; ctx->step_obst_cnt.p = d + o_cnt;
        ctx->step_seg_begin.p = d + o_sb; ctx->step_seg_end.p = d + o_se; ctx->seg_scene.p = d + o_ss;
        // captured graphs hold these addresses: a different layout is a different graph
        ctx->meta_sig = (long long)(o_sg * 1315423911ull ^ o_sso * 2654435761ull ^ o_slot * 40503ull ^ o_ss * 97ull ^ off);
        meta_bytes = off; m_org = o_org; m_sg = o_sg; m_sso = o_sso; m_ss = o_ss; m_flag = o_flag;
        use_seg_dev = seg_dev;
    }
    // everything that goes to the stream after the host-side passes
    auto enqueue_rest = [&]() -> int {
        char *h = (char *)ctx->h_meta, *d = (char *)ctx->meta.p;
        if (!use_seg_dev) {
            CU(cudaMemcpyAsync(d, h, meta_bytes, cudaMemcpyHostToDevice, st));
        } else {
            // everything but the two segment arrays, which k_order_segs writes from the raw copy
            CU(cudaMemcpyAsync(d + m_org, h + m_org, m_sg - m_org, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d + m_sso, h + m_sso, m_ss - m_sso, cudaMemcpyHostToDevice, st));
            const int njobs = (int)jobs.size();
            CU(ensure(ctx, ctx->sg_jobs, sizeof(int4) * (size_t)std::max(njobs, 1)));
            int4 *hj = (int4 *)(h + m_sg);                     // the unused segment area of the pinned block as staging
            for (int jb = 0; jb < njobs; jb++) hj[jb] = make_int4(jobs[jb].b, jobs[jb].e, jobs[jb].scene, jobs[jb].start);
            CU(cudaMemcpyAsync(ctx->sg_jobs.p, hj, sizeof(int4) * (size_t)njobs, cudaMemcpyHostToDevice, st));
            if (njobs) {
                k_order_segs<<<njobs, 256, 0, st>>>((const int4 *)ctx->sg_jobs.p, (const double *)ctx->sg_raw.p,
                                                   (double *)(d + m_sg), (int *)(d + m_ss));
                CU(cudaGetLastError());
            }
        }
        CU(cudaMemsetAsync(ctx->maxabs.p, 0, 2 * sizeof(float), st));
        if (nobst) {
            int blocks = (noslots + 127) / 128;
            if (VB == 4) {
                k_order_obst<4><<<nsteps, 256, 0, st>>>((const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_src_off.p,
                                                        (const int *)ctx->step_obst_off.p, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                        (const double *)ctx->origins_d.p, (int *)ctx->ob_order.p, Vo);
                k_stage_obst<4><<<blocks, 128, 0, st>>>(noslots, (const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_order.p,
                                                       (const int *)ctx->step_obst_off.p, nsteps, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                       (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, (float *)ctx->ob_c.p, (float4 *)ctx->ob_box.p,
                                                       (double *)ctx->ob_in.p, (int *)ctx->ob_nv.p, ob_stride, (float *)ctx->maxabs.p, Vo);
            } else {
                k_order_obst<8><<<nsteps, 256, 0, st>>>((const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_src_off.p,
                                                        (const int *)ctx->step_obst_off.p, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                        (const double *)ctx->origins_d.p, (int *)ctx->ob_order.p, Vo);
                k_stage_obst<8><<<blocks, 128, 0, st>>>(noslots, (const double *)ctx->ob_src.p, (const int *)ctx->ob_src_nv.p, (const int *)ctx->ob_order.p,
                                                       (const int *)ctx->step_obst_off.p, nsteps, (const int *)ctx->scene_step_off_d.p, nscenes,
                                                       (const double *)ctx->origins_d.p, (float4 *)ctx->ob_f.p, (float *)ctx->ob_c.p, (float4 *)ctx->ob_box.p,
                                                       (double *)ctx->ob_in.p, (int *)ctx->ob_nv.p, ob_stride, (float *)ctx->maxabs.p, Vo);
            }
            CU(cudaGetLastError());
        }
        if (nslots) {
            k_stage_segs<<<(nslots + 127) / 128, 128, 0, st>>>(nslots, (const double *)ctx->sg_64.p, (const int *)ctx->seg_scene.p, (const double *)ctx->origins_d.p,
                                                             (float4 *)ctx->sg_f.p, sg_stride, (float4 *)ctx->sg_box.p, (float *)ctx->maxabs.p);
            CU(cudaGetLastError());
        }
        CU(cudaMemcpyAsync(h + m_flag, ctx->maxabs.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
        return MPC_OK;
    };
    if (!use_graph) {
        const int rc = enqueue_rest();
        if (rc) return rc;
    } else {
        const std::vector<long long> key = {(long long)(size_t)obst, (long long)(size_t)obst_nv, nobst, Vo, VB, nsteps, nscenes, noslots, nslots,
                                            (long long)meta_bytes, ctx->meta_sig, ob_stride, sg_stride, (long long)ctx->generation,
                                            (long long)(size_t)ctx->h_meta};
        auto it = ctx->scene_graphs.find(key);
        if (it == ctx->scene_graphs.end()) {
            if (ctx->scene_graphs.size() > 64) {
                for (auto &kv : ctx->scene_graphs) cudaGraphExecDestroy(kv.second);
                ctx->scene_graphs.clear();
            }
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = copy_obstacles();
            if (!rc) rc = enqueue_rest();
            const cudaError_t e2 = cudaStreamEndCapture(st, &graph);
            if (rc || e2 != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                return rc ? rc : fail(ctx, MPC_E_CUDA, "scene graph capture failed: %s", cudaGetErrorString(e2));
            }
            cudaGraphExec_t exec = nullptr;
            const cudaError_t e3 = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e3 != cudaSuccess) return fail(ctx, MPC_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e3));
            it = ctx->scene_graphs.emplace(key, exec).first;
        }
        CU(cudaGraphLaunch(it->second, st));
    }
    lap("enqueue copies+kernels");
    ctx->meta_flag_off = m_flag;
    ctx->nscenes = nscenes; ctx->nsteps = nsteps; ctx->nobst = nobst; ctx->nseg = nslots; ctx->sg_stride = sg_stride;
    ctx->ob_stride = ob_stride;
    ctx->ob_vmax = VB; ctx->VB = VB; ctx->max_obst_step = max_o; ctx->max_seg_step = max_s;
    ctx->scene_step_off.assign(scene_step_off, scene_step_off + nscenes + 1);
    ctx->origins.assign(origins, origins + 2 * (size_t)nscenes);
    ctx->have_scenes = true;
    ctx->prepared = false;
    if (!sync) return MPC_OK;       // mpc_submit: the caller's buffers stay untouched until mpc_wait
    CU(cudaStreamSynchronize(st));
    lap("device sync");      // host buffers (incl. the local staging vectors) may be reused after return
    return check_scene_report(ctx);
}

// what the staging kernels report back (after the stream has drained)
static int check_scene_report(mpc_ctx *ctx) {
    const int bad = ((const int *)((const char *)ctx->h_meta + ctx->meta_flag_off))[1];
    if (bad > 0) {
        ctx->have_scenes = false;
        return fail(ctx, MPC_E_ARG, "%d obstacle piece(s) are not convex: decompose them before staging (the separating-axis test "
                                    "would report false separations)", bad);
    }
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
    p.lib_v64 = (const double *)ctx->lib_v64.p; p.lib_nv = (const int *)ctx->lib_nv.p; p.slot_t = (const int *)ctx->slot_t.p; p.slot_perm = p.slot_t + ctx->J;
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
    // early return across time slots needs earlier slots to have finished: only grids beyond one resident wave
    // compaction of the undecided candidates costs a serial pre-pass per CTA: it pays from ~32 groups per CTA on
    // (C4, ms per batch of 4 / 8 / 16 / 64 queries: 0.073 / 0.085 / 0.111 / 0.219 with, 0.066 / 0.088 / 0.123 / 0.307 without)
    p.compact = (ctx->lib_full && ctx->gpc >= 32) ? 1 : 0;
    if (getenv("MPC_NO_COMPACT")) p.compact = 0;                                                // tuning hook
    p.cmin = 32;
    if (const char *e = getenv("MPC_CMODE_MIN")) { p.cmin = std::max(1, atoi(e)); if (ctx->lib_full) p.compact = 1; }      // tuning hook
    p.c_total = ctx->J > 0 ? ctx->ctas / ctx->J : 0;
    p.wsplit = ctx->wsplit;
    p.tail = 0; p.host_res = (unsigned *)ctx->h_res; p.host_flag = ctx->h_flag; p.host_seq = 1u;
    p.res_words = (int)((sizeof(Counters) + sizeof(unsigned) * (size_t)ctx->mask_words) / 4);
    p.cross_exit = ctx->ctas > ctx->prop.multiProcessorCount * K_MIN_CTAS;
    p.q = (const QDev *)qp; p.cta_tab = (const int2 *)(qp + ctx->qpack_cta_off); p.cand = (const int *)ctx->cand_d.p;
    p.cand_pos = (const int *)ctx->sets_pos.p;
    p.cand_inv = p.cand_pos + ctx->set_total;
    p.B = ctx->B;
    p.wmax = ctx->wmax; p.any_sorted = ctx->any_sorted;
    p.cnt = (Counters *)ctx->res_p; p.mask = (unsigned *)((char *)ctx->res_p + sizeof(Counters));
    p.mask_sorted = p.mask + ctx->mask_words;
    p.nwords = ctx->mask_words;
    p.wl = (int4 *)ctx->wl.p; p.wl_cap = ctx->wl_cap;
    p.maxabs_query = (const float *)(qp + ctx->qpack_hdr_off); p.lib_rmax = ctx->lib_rmax; p.mode = ctx->mode;
    p.gpc = ctx->gpc; p.cap_o = ctx->cap_o; p.cap_s = ctx->cap_s;
    return p;
}

// Kernels of one step are chained as programmatic dependents (griddepcontrol): a launch may begin -- CTA set-up, table
// and query loads, tile fetch -- while the previous kernel of the step drains, and waits (pdl_wait) before it touches
// the result block.  MPC_NO_PDL=1 launches them as plain stream-ordered kernels (A/B hook).
static const bool g_use_pdl = getenv("MPC_NO_PDL") == nullptr;

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t st, bool dependent, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (dependent && g_use_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int VA, int VB, bool SEGCULL>
static cudaError_t launch_check_vb(mpc_ctx *ctx, const KParams &p, dim3 grid, bool dependent) {
    const bool count = ctx->mode & MPC_MODE_COUNT_WORK;
    if (ctx->mode & MPC_MODE_NO_CULL)          // full evaluation: its own kernel, two occupancy polygons per lane
        return launch_pdl(count ? k_check_full<VA, VB, true> : k_check_full<VA, VB, false>, grid, K_THREADS, ctx->smem, ctx->stream, dependent, p);
    if (p.wsplit > 1 && !count)                // few groups per CTA: several warps per group (planner-sized queries)
        return launch_pdl(k_check<VA, VB, false, SEGCULL, true>, grid, K_THREADS, ctx->smem, ctx->stream, dependent, p);
    return launch_pdl(count ? k_check<VA, VB, true, SEGCULL, false> : k_check<VA, VB, false, SEGCULL, false>, grid, K_THREADS, ctx->smem, ctx->stream, dependent, p);
}

template <int VA, int VB>
static cudaError_t launch_check_seg(mpc_ctx *ctx, const KParams &p, dim3 grid, bool dependent) {
    // steps with more than one block of 32 boundary segments get the instantiation with the segment sub-block culling
    // compiled in (the kernel is sensitive to its code size: C4, 4 segments, runs 5 % slower with it)
    return ctx->max_seg_step > 32 ? launch_check_vb<VA, VB, true>(ctx, p, grid, dependent) : launch_check_vb<VA, VB, false>(ctx, p, grid, dependent);
}

static cudaError_t launch_check_grid(mpc_ctx *ctx, const KParams &p, dim3 grid, bool dependent) {
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    const int va = ctx->VA, vb = ctx->VB;
    if (va == 4 && vb == 4) return launch_check_seg<4, 4>(ctx, p, grid, dependent);
    if (va == 4 && vb == 8) return launch_check_seg<4, 8>(ctx, p, grid, dependent);
    if (va == 8 && vb == 4) return launch_check_seg<8, 4>(ctx, p, grid, dependent);
    if (va == 8 && vb == 8) return launch_check_seg<8, 8>(ctx, p, grid, dependent);
    if (va == 16 && vb == 4) return launch_check_seg<16, 4>(ctx, p, grid, dependent);
    return launch_check_seg<16, 8>(ctx, p, grid, dependent);
}

// `dependent`: the kernel right before this one in the stream is a kernel of the same step (k_clear)
static cudaError_t launch_check(mpc_ctx *ctx, const KParams &p, bool dependent) {
    if (ctx->ctas == 0) return cudaSuccess;
    const int scouts = std::min(ctx->scouts, ctx->J - 1);
    if (scouts <= 0) return launch_check_grid(ctx, p, dim3((unsigned)p.c_total, (unsigned)ctx->J), dependent);
    // one launch per scout slot (finer chunks, nothing to look up yet / only the scouts before it), then the rest
    const char *qp = (const char *)ctx->qpack_d.p;
    for (int l = 0; l < scouts; l++) {
        KParams s = p;
        s.slot_perm = p.slot_perm + l;
        s.cta_tab = (const int2 *)(qp + ctx->qpack_cta2_off);
        s.gpc = ctx->scout_gpc;
        s.c_total = ctx->scout_c_total;
        s.wsplit = 1;
        s.cross_exit = l > 0;
        s.compact = (s.compact && s.gpc >= p.cmin) ? 1 : 0;
        cudaError_t e = launch_check_grid(ctx, s, dim3((unsigned)ctx->scout_c_total, 1u), dependent || l > 0);
        if (e != cudaSuccess) return e;
    }
    KParams m = p;
    m.slot_perm = p.slot_perm + scouts;
    m.cross_exit = 1;
    return launch_check_grid(ctx, m, dim3((unsigned)p.c_total, (unsigned)(ctx->J - scouts)), true);
}

// every instantiation of k_check may use the whole opt-in shared memory of the device (tiles of many-vertex polygons
// or long boundaries exceed the 48 KB default)
template <int VA, int VB>
static cudaError_t raise_smem_va_vb(int bytes) {
    cudaError_t e;
#define MPC_RAISE(...) \
    if ((e = cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    MPC_RAISE(k_check<VA, VB, false, false, false>)
    MPC_RAISE(k_check<VA, VB, false, true, false>)
    MPC_RAISE(k_check<VA, VB, true, false, false>)
    MPC_RAISE(k_check<VA, VB, true, true, false>)
    MPC_RAISE(k_check<VA, VB, false, false, true>)
    MPC_RAISE(k_check<VA, VB, false, true, true>)
    MPC_RAISE(k_check_full<VA, VB, false>)
    MPC_RAISE(k_check_full<VA, VB, true>)
#undef MPC_RAISE
    return cudaSuccess;
}

static cudaError_t raise_smem_limits(const cudaDeviceProp &prop) {
    const int bytes = (int)prop.sharedMemPerBlockOptin;
    cudaError_t e;
    if ((e = raise_smem_va_vb<4, 4>(bytes)) != cudaSuccess) return e;
    if ((e = raise_smem_va_vb<4, 8>(bytes)) != cudaSuccess) return e;
    if ((e = raise_smem_va_vb<8, 4>(bytes)) != cudaSuccess) return e;
    if ((e = raise_smem_va_vb<8, 8>(bytes)) != cudaSuccess) return e;
    if ((e = raise_smem_va_vb<16, 4>(bytes)) != cudaSuccess) return e;
    return raise_smem_va_vb<16, 8>(bytes);
}

static int launch_all(mpc_ctx *ctx, cudaEvent_t e0, cudaEvent_t e1, cudaEvent_t e2, bool clear = true, bool tail = false) {
    KParams p = make_params(ctx);
    p.tail = tail ? 1 : 0;
    cudaStream_t st = ctx->stream;
    if (e0) CU(cudaEventRecord(e0, st));
    if (clear) {
        const int n16 = (int)((sizeof(Counters) + (size_t)2 * ctx->mask_words * sizeof(unsigned) + 15) / 16);
        k_clear<<<std::min(64, (n16 + 255) / 256), 256, 0, st>>>((uint4 *)ctx->res_p, n16);
        CU(cudaGetLastError());
    }
    // event records between the kernels would break the programmatic chain: the instrumented path (mpc_run_prepared)
    // times the whole step with e0 / e3 and k_check alone in a second pass (see there)
    if (e1) CU(cudaEventRecord(e1, st));
    CU(launch_check(ctx, p, clear && !e1));
    if (e2) CU(cudaEventRecord(e2, st));
    if (tail) return MPC_OK;          // the last CTA of k_check has refined and handed the result over
    // one warp per work-list item / result word; a planner-sized batch has a handful of either: a small grid starts faster
    const int rgrid = std::max(1, ctx->prop.multiProcessorCount) * (ctx->fused_clear ? 1 : 8);
    CU(launch_pdl(k_refine, dim3((unsigned)rgrid), 128, 0, st, !e2 && ctx->ctas > 0, p));
    return MPC_OK;
}

static int issue_uploads(mpc_ctx *ctx);
static cudaError_t ensure_result_host(mpc_ctx *ctx, size_t bytes);

static int prepare_host(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                        unsigned mode, bool upload) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (ctx->pending) return fail(ctx, MPC_E_STATE, "a submitted batch is waiting for mpc_wait");
    if (!ctx->have_lib || !ctx->have_scenes) return fail(ctx, MPC_E_STATE, "upload the library and the scenes before querying");
    if (B < 0 || (B > 0 && !queries)) return fail(ctx, MPC_E_ARG, "bad query arguments");
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    // grid shape: groups of 32 candidates, gpc groups per CTA; aim for a few CTAs per SM
    long long total_groups = 0, total_words = 0;
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
        total_words += ((long long)d.cand_cnt + 31) / 32;
    }
    for (int i = 0; i < ncand; i++)
        if (cand_idx[i] < 0 || cand_idx[i] >= ctx->P) return fail(ctx, MPC_E_ARG, "cand_idx[%d]=%d outside the library", i, cand_idx[i]);
    const int sms = ctx->prop.multiProcessorCount;
    ctx->threads = K_THREADS;
    const int nwarps = K_THREADS / 32;
    // groups (of 32 candidates) per CTA: enough CTAs for ~2 waves of 4 CTAs per SM (2-4 waves measure the same); big batches get big CTAs (every CTA
    // pays for its tile: measured 0.78 / 0.49 / 0.40 / 0.33 / 0.30 / 0.29 ms per C4 batch at 4 / 8 / 16 / 32 / 64 / 128
    // groups).  The count is then evened out over the chunks of the largest query (48 + 48 + 32 is a tail, 43 + 43 + 42
    // is not).
    int waves = (mode & MPC_MODE_NO_CULL) ? 6 : 2;      // full evaluation: long uniform CTAs, more of them balance better
    if (const char *e = getenv("MPC_WAVES")) waves = std::max(1, atoi(e));                       // tuning hook
    const long long target = (long long)sms * K_MIN_CTAS * waves;
    int gpc = (int)((total_groups + target - 1) / target);
    // with compaction two chunks per query beat one (0.184 vs 0.189 ms): less work in flight per time slot, so the
    // collision bits of earlier slots arrive sooner
    const int gpc_cap = ctx->lib_full ? GPC_MAX / 2 : GPC_MAX;
    gpc = std::max(1, std::min(gpc, gpc_cap));
    if (gpc > nwarps / 2 && gpc <= 2 * nwarps) gpc = (gpc + nwarps - 1) / nwarps * nwarps;     // small: whole rounds of the warps
    {
        int gmax = 1;
        for (int q = 0; q < B; q++) gmax = std::max(gmax, (queries[q].cand_cnt + 31) / 32);
        if (gpc > 2 * nwarps && gpc < gmax) {
            const int chunks = (gmax + gpc - 1) / gpc;
            gpc = std::min(gpc_cap, (gmax + chunks - 1) / chunks);
        }
    }
    // Scout launches.  In a slot-major grid the CTAs of the first resident wave start together and cannot use each
    // other's collision bits; when one time slot fills only a fraction of that wave, the wave spans several slots and
    // every candidate is evaluated at all of them.  The first slot of the launch order is therefore launched on its
    // own, cut into enough CTAs to fill the machine (C4: slot 49 alone decides 75 % of the candidates); the rest of
    // the grid starts with those bits visible, and since most of its candidates are decided already it runs best as
    // one CTA per (query, slot).  C4 batch: 0.129 ms -> 0.112 ms (scout) -> 0.104 ms (scout + 128 groups per CTA);
    // two scouts 0.124 ms (every launch boundary drains the machine).
    const long long resident = (long long)sms * K_MIN_CTAS;
    auto ctas_per_slot = [&](int g) {
        long long c = 0;
        for (int q = 0; q < B; q++) {
            const int groups = (queries[q].cand_cnt + 31) / 32;
            c += groups ? std::max(1, (groups + g - 1) / g) : 0;
        }
        return c;
    };
    int scouts = 0, sg = nwarps;               // sg: groups per CTA of the scout launch, one slot ~ one resident wave
    {
        long long groups_slot = 0;
        for (int q = 0; q < B; q++) groups_slot += (queries[q].cand_cnt + 31) / 32;
        sg = (int)((groups_slot + resident - 1) / std::max<long long>(resident, 1));
        sg = std::max(nwarps, std::min((sg + nwarps - 1) / nwarps * nwarps, gpc));
        if (const char *e = getenv("MPC_SCOUT_GPC")) sg = std::max(1, std::min(atoi(e), GPC_MAX));             // tuning hook
        const long long c = ctas_per_slot(gpc);
        // worth it when a slot fills less than half the wave, the grid is at least two waves, and the scout itself puts
        // a CTA on most SMs (C4, ms per batch of 4 / 8 / 16 / 32 / 64 queries: 0.045 / 0.073 / 0.082 / 0.097 / 0.129
        // without, 0.052 / 0.046 / 0.064 / 0.077 / 0.104 with)
        if (!(mode & (MPC_MODE_NO_EARLY_EXIT | MPC_MODE_NO_CULL)) && ctx->J >= 8 && c > 0 && 2 * c <= resident && c * ctx->J >= 2 * resident &&
            ctas_per_slot(sg) * 8 >= resident)
            scouts = 1;
        if (const char *e = getenv("MPC_SCOUTS")) scouts = std::max(0, std::min(atoi(e), ctx->J - 1));      // tuning hook
        if (scouts && gpc == gpc_cap && gpc < GPC_MAX) {
            int gmax = 1;
            for (int q = 0; q < B; q++) gmax = std::max(gmax, (queries[q].cand_cnt + 31) / 32);
            const int chunks = (gmax + GPC_MAX - 1) / GPC_MAX;
            gpc = std::max(gpc, std::min(GPC_MAX, (gmax + chunks - 1) / chunks));
        }
    }
    if (const char *e = getenv("MPC_GPC")) gpc = std::max(1, std::min(atoi(e), GPC_MAX));      // tuning hook
    ctx->gpc = gpc;
    const int F = ob_fields(ctx->VB), FA = ob_fields(ctx->VA);
    ctx->cap_o = std::max(32, std::min((ctx->max_obst_step + 31) / 32 * 32, ctx->VB == 4 ? 320 : 160));
    ctx->cap_s = std::max(32, std::min((ctx->max_seg_step + 31) / 32 * 32, 512));
    ctx->smem = 16 + 4 * GPC_MAX * 4 + (size_t)nwarps * QCAP * 4 + (size_t)nwarps * FA * 32 * 16 +
                (size_t)ctx->cap_o * (F * 16 + 12) + (size_t)ctx->cap_s * 32;
    if (ctx->smem > ctx->prop.sharedMemPerBlockOptin)
        return fail(ctx, MPC_E_LIMIT, "tile of %zu bytes exceeds the device's shared memory", ctx->smem);
    // pack: QDev[B] | CTA table of the main grid | CTA table of the scout launch | header | (zeros: fused clear) |
    // explicit candidates (second copy, only when present).  A table entry is (query, chunk) of one CTA of a time slot.
    const long long c_main = ctas_per_slot(gpc), c_scout = ctas_per_slot(sg);
    if (c_main > 0x3fffffffLL || c_scout > 0x3fffffffLL) return fail(ctx, MPC_E_LIMIT, "batch too large");
    size_t bytes_q = sizeof(QDev) * (size_t)std::max(B, 1), bytes_c1 = (sizeof(int2) * (size_t)std::max<long long>(c_main, 1) + 15) / 16 * 16,
           bytes_c2 = (sizeof(int2) * (size_t)std::max<long long>(c_scout, 1) + 15) / 16 * 16, bytes_c = bytes_c1 + bytes_c2;
    const size_t bytes_h = 16, bytes_i = sizeof(int) * (size_t)ncand;
    // a small result block (counters + two mask arrays) sits behind the query pack in the same device buffer and is
    // cleared by zeros appended to the upload: one copy node instead of copy + clear in the replayed graph
    const size_t up0 = bytes_q + bytes_c + bytes_h;
    const size_t clear_bytes = sizeof(Counters) + sizeof(unsigned) * (size_t)2 * (size_t)total_words;
    const bool fused = clear_bytes <= 8192;
    const size_t zoff = (up0 + 255) & ~(size_t)255, zbytes = fused ? (clear_bytes + 15) / 16 * 16 : 0;
    const size_t cand_src = fused ? zoff + zbytes : up0;
    CU(ensure_pinned(ctx, cand_src + bytes_i + 64));
    if (fused) memset((char *)ctx->h_pin + up0, 0, zoff + zbytes - up0);
    QDev *hq = (QDev *)ctx->h_pin;
    int2 *htab = (int2 *)((char *)ctx->h_pin + bytes_q), *htab2 = (int2 *)((char *)ctx->h_pin + bytes_q + bytes_c1);
    float *hhdr = (float *)((char *)ctx->h_pin + bytes_q + bytes_c);
    int *hidx = (int *)((char *)ctx->h_pin + cand_src);
    long long ctas = 0, words = 0, c2 = 0;
    double maxq = 0.0;
    int wmax = 1, any_sorted = 0;
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
        const int groups = (d.cand_cnt + 31) / 32;
        wmax = std::max(wmax, groups);
        any_sorted |= o.pos_base >= 0;
        o.chunks = std::max(1, (groups + gpc - 1) / gpc);
        if (groups) {                              // per time slot: the grid is slot-major, ctas * J blocks in all
            for (int c = 0; c < o.chunks; c++) htab[ctas + c] = make_int2(q, c);
            ctas += o.chunks;
            const int ch2 = std::max(1, (groups + sg - 1) / sg);
            for (int c = 0; c < ch2; c++) htab2[c2 + c] = make_int2(q, c);
            c2 += ch2;
        }
        words += groups;
        maxq = std::max(maxq, std::max(std::fabs(d.x - ox), std::fabs(d.y - oy)) + ctx->lib_radius);
        if (!(std::fabs(d.c) <= 1.0000001 && std::fabs(d.s) <= 1.0000001)) return fail(ctx, MPC_E_ARG, "query %d: (c, s) is not a rotation", q);
    }
    if (ctas != c_main || c2 != c_scout) return fail(ctx, MPC_E_STATE, "internal: CTA table size");
    if (ctas == 0) scouts = 0;
    ctx->scouts = scouts; ctx->scout_gpc = sg; ctx->scout_c_total = (int)c2;
    ctas *= ctx->J;
    if (ctas > 0x7fffffffLL || words > 0x7fffffffLL) return fail(ctx, MPC_E_LIMIT, "batch too large");
    // few groups per CTA and a grid that fits the machine at once (no early return across slots to coordinate): several
    // warps per group, each with a share of the tile
    ctx->wsplit = 1;
    if (gpc <= 4 && scouts == 0 && ctas <= (long long)sms * K_MIN_CTAS && !(mode & (MPC_MODE_NO_CULL | MPC_MODE_COUNT_WORK)))
        ctx->wsplit = gpc == 1 ? 8 : (gpc == 2 ? 4 : 2);
    if (getenv("MPC_NO_WSPLIT")) ctx->wsplit = 1;                                              // tuning hook
    // ... and such a batch with a small result block runs as ONE kernel (see the tail of k_check)
    static const bool no_tail = getenv("MPC_NO_TAIL") != nullptr;
    // (not on maps whose boundary needs several tiles per step: their work lists run to hundreds of items, too many for
    // the eight warps of one CTA -- ZAM_ACC, 2362 segments: 62 us single-kernel vs 57 us with k_refine on all SMs)
    ctx->tail = ctx->wsplit > 1 && fused && !no_tail && ctx->max_seg_step <= 512 && ctx->max_obst_step <= 320;
    if (ncand) memcpy(hidx, cand_idx, bytes_i);
    ctx->maxabs_query = std::nextafter((float)maxq, INFINITY);
    hhdr[0] = ctx->maxabs_query;
    hhdr[1] = hhdr[2] = hhdr[3] = 0.f;
    if (words != total_words) return fail(ctx, MPC_E_STATE, "internal: mask word count");
    CU(ensure(ctx, ctx->qpack_d, fused ? zoff + zbytes : up0));
    ctx->qpack_cta_off = bytes_q;
    ctx->qpack_cta2_off = bytes_q + bytes_c1;
    ctx->qpack_hdr_off = bytes_q + bytes_c;
    ctx->up_bytes = fused ? zoff + zbytes : up0;
    ctx->up_cand_bytes = bytes_i;
    ctx->up_cand_src = cand_src;
    ctx->fused_clear = fused;
    // candidate arena = staged sets followed by this call's explicit lists
    size_t arena = sizeof(int) * (size_t)std::max(2 * ctx->set_total + ncand, 1);
    void *arena_before = ctx->cand_d.p;
    CU(ensure(ctx, ctx->cand_d, arena));
    if ((ctx->cand_d.p != arena_before || !ctx->arena_valid) && ctx->set_total)
        CU(cudaMemcpyAsync(ctx->cand_d.p, ctx->sets.p, sizeof(int) * (size_t)2 * ctx->set_total, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->arena_valid = true;
    CU(ensure(ctx, ctx->res_d, sizeof(Counters) + sizeof(unsigned) * (size_t)2 * std::max<long long>(words, 1)));
    CU(ensure_result_host(ctx, sizeof(Counters) + sizeof(unsigned) * (size_t)std::max<long long>(words, 1)));
    ctx->res_p = fused ? (void *)((char *)ctx->qpack_d.p + zoff) : ctx->res_d.p;
    ctx->B = B; ctx->ctas = (int)ctas; ctx->mask_words = (int)words; ctx->mode = mode;
    ctx->wmax = wmax; ctx->any_sorted = any_sorted;
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
    st->done_skipped = (int64_t)c.done_skipped;
    st->worklist_overflow = overflow;
    st->ms_device = ms;
    st->ctas = ctx->ctas;
    st->kernels = kernels;
    st->tau = c.tau;
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
        *kernels += 3 + ctx->scouts;
        if (timed) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_res, ctx->res_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
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
static int run_graph(mpc_ctx *ctx, float *ms, bool timed, bool sync = true) {
    const size_t bytes = sizeof(Counters) + (size_t)ctx->mask_words * sizeof(unsigned);
    const bool tail = ctx->tail && ctx->ctas > 0;
    std::vector<long long> key = {(long long)ctx->B, (long long)ctx->ctas, (long long)ctx->scouts * 1000 + ctx->scout_gpc + 10000000LL * ctx->wsplit, (long long)ctx->scout_c_total, (long long)ctx->fused_clear + 2 * (long long)tail + 4 * (long long)ctx->any_sorted + 8LL * ctx->wmax, (long long)ctx->mask_words, (long long)ctx->mode,
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
        if (!rc) rc = launch_all(ctx, nullptr, nullptr, nullptr, !ctx->fused_clear, tail);     // the upload just cleared it
        cudaError_t e1 = cudaSuccess;
        if (!rc && !tail) e1 = cudaMemcpyAsync(ctx->h_res, ctx->res_p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
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
    if (tail) *ctx->h_flag = 0u;
    if (timed) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaGraphLaunch(it->second, ctx->stream));
    if (timed) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (!sync) return MPC_OK;
    if (tail && !timed) {
        // the last CTA raises the flag right after the result block has landed in host memory: watching that word is a
        // few microseconds faster than waking up from a stream synchronisation
        for (unsigned spins = 1; *ctx->h_flag == 0u; spins++) {
            __builtin_ia32_pause();
            if ((spins & 0xfffu) == 0u) {
                const cudaError_t qe = cudaStreamQuery(ctx->stream);
                if (qe == cudaSuccess) break;                                   // drained: the flag is up (or the launch did nothing)
                if (qe != cudaErrorNotReady) return fail(ctx, MPC_E_CUDA, "query failed on the device: %s", cudaGetErrorString(qe));
            }
        }
        if (*ctx->h_flag == 0u) CU(cudaStreamSynchronize(ctx->stream));
        return MPC_OK;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (timed) CU(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return MPC_OK;
}

// after the graph of the batch has drained: grow the refinement list and re-run if it overflowed, hand out the result
static int finish_batch(mpc_ctx *ctx, uint32_t *out_mask, int mask_words, mpc_stats *stats, float ms, bool timed) {
    int kernels = (ctx->fused_clear ? 2 : 3) + ctx->scouts - ((ctx->tail && ctx->ctas > 0) ? 1 : 0), overflow = 0, rc = MPC_OK;
    if (((const Counters *)ctx->h_res)->wl_count > (unsigned)ctx->wl_cap) {
        // refinement list overflowed: the plain path grows it and re-runs (and drops the graphs)
        rc = issue_uploads(ctx);
        if (!rc) rc = run_once(ctx, &ms, &kernels, &overflow, timed);
        overflow = 1;
    }
    if (rc) return rc;
    if (mask_words) memcpy(out_mask, (const char *)ctx->h_res + sizeof(Counters), sizeof(unsigned) * (size_t)mask_words);
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

// run the batch prepare_host() packed: graph replay for the common small-tile case, plain launches otherwise
static int run_batch(mpc_ctx *ctx, uint32_t *out_mask, int mask_words, mpc_stats *stats) {
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words > 0 && !out_mask) return fail(ctx, MPC_E_ARG, "out_mask missing");
    float ms = 0.f;
    const bool timed = stats != nullptr && !(ctx->mode & MPC_MODE_NO_TIMING);
    const int rc = run_graph(ctx, &ms, timed);
    if (rc) return rc;
    return finish_batch(ctx, out_mask, mask_words, stats, ms, timed);
}

// ---------------------------------------------------------------------------------------------------------------------
// asynchronous cycle: mpc_submit stages the scenes (optional) and enqueues the query batch, mpc_wait collects it.  One
// host thread can keep several contexts (lanes) of a GPU busy this way: while one lane's k_check runs, the next lane's
// prediction travels host -> device -- without a host thread per lane.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mpc_submit(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                          const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                          const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg,
                          int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand, unsigned mode) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (ctx->pending) return fail(ctx, MPC_E_STATE, "a submitted batch is waiting for mpc_wait");
    static const bool dbg_t = getenv("MPC_DEBUG_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = MPC_OK;
    ctx->pending_scenes = false;
    if (nscenes > 0) {
        rc = set_scenes_impl(ctx, nscenes, scene_step_off, origins, step_obst_off, obst, obst_nv, Vo, step_seg_begin, step_seg_end,
                             segs, nseg, false);
        if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
        ctx->pending_scenes = true;
    }
    const auto t1 = std::chrono::steady_clock::now();
    rc = prepare_host(ctx, B, queries, cand_idx, ncand, mode | MPC_MODE_NO_TIMING, false);
    const auto t2 = std::chrono::steady_clock::now();
    float ms = 0.f;
    if (!rc) rc = run_graph(ctx, &ms, false, false);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    if (dbg_t) {
        const auto t3 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mpc submit] scenes %.1f us, pack %.1f us, graph launch %.1f us\n",
                std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t1).count(),
                std::chrono::duration<double, std::micro>(t3 - t2).count());
    }
    ctx->pending = true;
    return MPC_OK;
}

extern "C" int mpc_wait(mpc_ctx *ctx, uint32_t *out_mask, int mask_words, mpc_stats *stats) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->pending) return fail(ctx, MPC_E_STATE, "nothing submitted");
    CU(cudaSetDevice(ctx->device));
    ctx->pending = false;
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->pending_scenes) {
        ctx->pending_scenes = false;
        const int rc = check_scene_report(ctx);
        if (rc) return rc;
    }
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words > 0 && !out_mask) return fail(ctx, MPC_E_ARG, "out_mask missing");
    return finish_batch(ctx, out_mask, mask_words, stats, 0.f, false);
}

extern "C" int mpc_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                         unsigned mode, uint32_t *out_mask, int mask_words, mpc_stats *stats) {
    static const bool dbg_t = getenv("MPC_DEBUG_TIMING") != nullptr;
    if (ctx && ctx->pending) return fail(ctx, MPC_E_STATE, "a submitted batch is waiting for mpc_wait");
    const auto t0 = std::chrono::steady_clock::now();
    int rc = prepare_host(ctx, B, queries, cand_idx, ncand, mode, false);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = run_batch(ctx, out_mask, mask_words, stats);
    if (dbg_t) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mpc query] pack %.1f us, launch + wait + copy back %.1f us\n",
                std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t1).count());
    }
    return rc;
}

extern "C" int mpc_run_prepared(mpc_ctx *ctx, int iters, int flush_l2, float *ms_total, float *ms_main, mpc_stats *stats) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    CU(cudaSetDevice(ctx->device));
    float ms = 0.f;
    int kernels = 0, overflow = 0;
    int rc = run_once(ctx, &ms, &kernels, &overflow, true);     // sizes the refinement list
    if (rc) return rc;
    kernels = 0;                                                // stats->kernels: launches of the timed pass only
    const size_t flush_bytes = 256u << 20;
    if (flush_l2) CU(ensure(ctx, ctx->flush, flush_bytes));
    // pass 1 times the step as it runs in production (clear, k_check launches, k_refine chained as programmatic
    // dependents, no event between them); pass 2 repeats it with events around the k_check launches alone, which makes
    // those launches plain stream-ordered ones (ms_main: the kernel the roofline is stated for)
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !ms_main) break;
        for (int it = 0; it < iters; it++) {
            if (flush_l2) {
                k_fill_u32<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>((unsigned *)ctx->flush.p, flush_bytes / 4, (unsigned)it);
                CU(cudaGetLastError());
            }
            rc = pass == 0 ? launch_all(ctx, ctx->ev[0], nullptr, nullptr) : launch_all(ctx, ctx->ev[0], ctx->ev[1], ctx->ev[2]);
            if (rc) return rc;
            CU(cudaEventRecord(ctx->ev[3], ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float a = 0.f, b = 0.f;
            CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[3]));
            if (pass == 0) {
                if (ms_total) ms_total[it] = a;
                ms = a;
                kernels += 3 + ctx->scouts;
            } else {
                CU(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
                ms_main[it] = b;
            }
        }
    }
    const size_t bytes = sizeof(Counters) + (size_t)ctx->mask_words * sizeof(unsigned);
    CU(cudaMemcpyAsync(ctx->h_res, ctx->res_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    fill_stats(ctx, stats, ms, kernels, overflow);
    return MPC_OK;
}

extern "C" int mpc_fetch_mask(mpc_ctx *ctx, uint32_t *out_mask, int mask_words) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->prepared) return fail(ctx, MPC_E_STATE, "no prepared batch");
    if (mask_words != ctx->mask_words) return fail(ctx, MPC_E_ARG, "mask_words=%d but the batch needs %d", mask_words, ctx->mask_words);
    if (mask_words) {
        CU(cudaMemcpyAsync(out_mask, (const char *)ctx->res_p + sizeof(Counters), sizeof(unsigned) * (size_t)mask_words, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return MPC_OK;
}

extern "C" int mpc_measure_fp32_peak(mpc_ctx *ctx, double out[8]) {
    if (!ctx || !ctx->stream || !out) return MPC_E_STATE;
    CU(cudaSetDevice(ctx->device));
    CU(ensure(ctx, ctx->flush, 1 << 20));
    float *sink = (float *)ctx->flush.p;
    long long *cyc = (long long *)ctx->flush.p + 64;
    const int sms = ctx->prop.multiProcessorCount, grid = sms * 8, iters = 2048;
    // per round: {FFMA, FFMA2, FMNMX3, FMNMX}; flop counted per round (multiply-add = 2, min / max = 1 per input pair)
    struct Kind { void (*k)(float *, int, float, float, long long *); double flop_round; };
    const Kind kinds[7] = {
        {k_peak<8, 0, 0, 0>, 8 * 2.0},                       // 0 scalar FFMA
        {k_peak<0, 8, 0, 0>, 8 * 4.0},                       // 1 packed FFMA2
        {k_peak<0, 8, 3, 2>, 8 * 4.0 + 3 * 2 + 2},           // 2 full test as k_check runs it: 8 FFMA2 : 3 FMNMX3 : 2 FMNMX (160 flop per 8 rounds)
        {k_peak<16, 0, 3, 2>, 16 * 2.0 + 3 * 2 + 2},         // 3 the same test with scalar FFMA
        {k_peak<0, 0, 8, 0>, 8 * 2.0},                       // 4 FMNMX3 alone
        {k_peak<0, 0, 0, 8>, 8 * 1.0},                       // 5 FMNMX alone
        {k_peak<8, 0, 0, 4>, 8 * 2.0 + 4},                   // 6 FFMA with half as many FMNMX (do the pipes overlap?)
    };
    for (int kind = 0; kind < 7; kind++) {
        float best = 1e30f;
        long long cycles = 0;
        for (int rep = 0; rep < 5; rep++) {                     // rep 0 warms up
            CU(cudaEventRecord(ctx->ev[0], ctx->stream));
            kinds[kind].k<<<grid, 256, 0, ctx->stream>>>(sink, iters, 1.0000001f, 1e-9f, cyc);
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            if (rep > 0 && ms < best) {
                best = ms;
                CU(cudaMemcpy(&cycles, cyc, sizeof cycles, cudaMemcpyDeviceToHost));
            }
        }
        out[kind] = (double)grid * 256 * iters * 8 * kinds[kind].flop_round / (best * 1e-3) / 1e12;
        if (kind == 0) out[7] = (double)cycles / (best * 1e-3) / 1e6;    // one CTA's cycles over the launch time: a lower bound of the clock
    }
    return MPC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// primitive-selection step: collision bits + child cost + goal test in one call
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mpc_upload_trajectories(mpc_ctx *ctx, const double *states, const int32_t *nstates, int P, int L) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!states || !nstates || P <= 0 || L <= 0) return fail(ctx, MPC_E_ARG, "bad trajectory arguments");
    if (L > EXPAND_MAX_STATES) return fail(ctx, MPC_E_LIMIT, "L=%d exceeds %d states per primitive", L, EXPAND_MAX_STATES);
    for (int p = 0; p < P; p++)
        if (nstates[p] < 0 || nstates[p] > L) return fail(ctx, MPC_E_ARG, "nstates[%d]=%d outside 0..%d", p, nstates[p], L);
    CU(cudaSetDevice(ctx->device));
    ctx->have_traj = false;
    CU(ensure(ctx, ctx->tr_states, sizeof(double) * 4 * (size_t)P * L));
    CU(ensure(ctx, ctx->tr_n, sizeof(int) * (size_t)P));
    CU(cudaMemcpyAsync(ctx->tr_states.p, states, sizeof(double) * 4 * (size_t)P * L, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tr_n.p, nstates, sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->tr_P = P; ctx->tr_L = L;
    ctx->have_traj = true;
    return MPC_OK;
}

extern "C" int mpc_set_guidance(mpc_ctx *ctx, const double *ref_xy, int M, const mpc_goal *goals, int ngoals,
                                const double *goal_segs, int nseg, double b) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (M < 0 || ngoals < 0 || nseg < 0 || (M > 0 && !ref_xy) || (ngoals > 0 && !goals) || (nseg > 0 && !goal_segs))
        return fail(ctx, MPC_E_ARG, "bad guidance arguments");
    for (int g = 0; g < ngoals; g++)
        if (goals[g].seg_begin >= 0 && (goals[g].seg_end < goals[g].seg_begin || goals[g].seg_end > nseg))
            return fail(ctx, MPC_E_ARG, "goal %d: segment range %d..%d outside 0..%d", g, goals[g].seg_begin, goals[g].seg_end, nseg);
    CU(cudaSetDevice(ctx->device));
    ctx->have_guidance = false;
    static_assert(sizeof(mpc_goal) == sizeof(EGoal), "goal record layout");
    CU(ensure(ctx, ctx->g_ref, sizeof(double) * 2 * (size_t)std::max(M, 1)));
    CU(ensure(ctx, ctx->g_goals, sizeof(EGoal) * (size_t)std::max(ngoals, 1)));
    CU(ensure(ctx, ctx->g_segs, sizeof(double) * 4 * (size_t)std::max(nseg, 1)));
    if (M) CU(cudaMemcpyAsync(ctx->g_ref.p, ref_xy, sizeof(double) * 2 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
    if (ngoals) CU(cudaMemcpyAsync(ctx->g_goals.p, goals, sizeof(EGoal) * (size_t)ngoals, cudaMemcpyHostToDevice, ctx->stream));
    if (nseg) CU(cudaMemcpyAsync(ctx->g_segs.p, goal_segs, sizeof(double) * 4 * (size_t)nseg, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->g_M = M; ctx->g_ngoals = ngoals; ctx->g_nseg = nseg; ctx->g_b = b;
    ctx->have_guidance = true;
    return MPC_OK;
}

extern "C" int mpc_expand(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                          const double *parent, unsigned mode, uint32_t *out_mask, int mask_words, double *out_cost,
                          int32_t *out_goal_step, mpc_stats *stats) {
    if (!ctx || !ctx->stream) return MPC_E_STATE;
    if (!ctx->have_traj || !ctx->have_guidance) return fail(ctx, MPC_E_STATE, "upload the trajectories and the guidance before expanding");
    int rc = prepare_host(ctx, B, queries, cand_idx, ncand, mode, false);      // validates the batch, packs the collision part
    if (rc) return rc;
    if (ctx->tr_P != ctx->P) return fail(ctx, MPC_E_STATE, "trajectories for %d primitives, library has %d", ctx->tr_P, ctx->P);
    long long total = 0;
    for (int q = 0; q < B; q++) total += queries[q].cand_cnt;
    if (total > 0x7fffffffLL) return fail(ctx, MPC_E_LIMIT, "batch too large");
    if (total > 0 && (!parent || !out_cost || !out_goal_step)) return fail(ctx, MPC_E_ARG, "parent / out_cost / out_goal_step missing");
    if (total > 0) {
        // pack: EQuery[B] | out_off[B+1] | explicit candidates; results: cost[total] | goal_step[total]
        const size_t b_q = sizeof(EQuery) * (size_t)B, b_o = (sizeof(int) * (size_t)(B + 1) + 15) / 16 * 16, b_c = sizeof(int) * (size_t)ncand;
        const size_t b_in = b_q + b_o + b_c, b_cost = sizeof(double) * (size_t)total, b_out = b_cost + sizeof(int) * (size_t)total;
        if (b_in + b_out > ctx->h_exp_cap) {
            if (ctx->h_exp) cudaFreeHost(ctx->h_exp);
    if (ctx->h_flag) cudaFreeHost((void *)ctx->h_flag);
            ctx->h_exp = nullptr;
            ctx->h_exp_cap = 0;
            CU(cudaMallocHost(&ctx->h_exp, (b_in + b_out) * 2));
            ctx->h_exp_cap = (b_in + b_out) * 2;
        }
        CU(ensure(ctx, ctx->e_pack, b_in));
        CU(ensure(ctx, ctx->e_out, b_out));
        EQuery *hq = (EQuery *)ctx->h_exp;
        int *hoff = (int *)((char *)ctx->h_exp + b_q);
        int acc = 0;
        for (int q = 0; q < B; q++) {
            const mpc_query_desc &d = queries[q];
            EQuery &o = hq[q];
            o.x = d.x; o.y = d.y; o.c = d.c; o.s = d.s;
            o.cost = parent[2 * q]; o.phi = parent[2 * q + 1];
            o.t0 = d.t0;
            o.cand_src = d.cand_set < 0;
            o.cand_base = d.cand_set < 0 ? d.cand_off : ctx->set_off[d.cand_set] + d.cand_off;
            o.cand_cnt = d.cand_cnt;
            hoff[q] = acc;
            acc += d.cand_cnt;
        }
        hoff[B] = acc;
        if (ncand) memcpy((char *)ctx->h_exp + b_q + b_o, cand_idx, b_c);
        cudaStream_t st = ctx->stream;
        CU(cudaMemcpyAsync(ctx->e_pack.p, ctx->h_exp, b_in, cudaMemcpyHostToDevice, st));
        EParams e;
        e.q = (const EQuery *)ctx->e_pack.p;
        e.out_off = (const int *)((const char *)ctx->e_pack.p + b_q);
        e.B = B; e.total = (int)total;
        e.sets = (const int *)ctx->sets.p;
        e.explicit_cand = (const int *)((const char *)ctx->e_pack.p + b_q + b_o);
        e.states = (const double *)ctx->tr_states.p; e.nstates = (const int *)ctx->tr_n.p; e.L = ctx->tr_L;
        e.ref_xy = (const double *)ctx->g_ref.p; e.M = ctx->g_M;
        e.goals = (const EGoal *)ctx->g_goals.p; e.ngoals = ctx->g_ngoals;
        e.goal_segs = (const double *)ctx->g_segs.p; e.b = ctx->g_b;
        e.cost = (double *)ctx->e_out.p;
        e.goal_step = (int *)((char *)ctx->e_out.p + b_cost);
        k_expand<<<(int)((total + 127) / 128), 128, 0, st>>>(e);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync((char *)ctx->h_exp + b_in, ctx->e_out.p, b_out, cudaMemcpyDeviceToHost, st));
        rc = run_batch(ctx, out_mask, mask_words, stats);          // same stream; its synchronisation covers the copy above
        if (rc) return rc;
        memcpy(out_cost, (const char *)ctx->h_exp + b_in, b_cost);
        memcpy(out_goal_step, (const char *)ctx->h_exp + b_in + b_cost, sizeof(int) * (size_t)total);
        if (stats) stats->kernels += 1;
        return MPC_OK;
    }
    return run_batch(ctx, out_mask, mask_words, stats);
}
