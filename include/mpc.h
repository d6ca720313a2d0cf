/* mpc.h -- C ABI of libmpcheck.so: batched motion-primitive collision check on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of KochdumperNiklas/MotionPlanner that this repository
 * accelerates.  The reference is pure Python and has no FFI for this path; each entry point below names the
 * reference interface (file:line under /root/reference) whose work it replaces, and INTEGRATION.md shows the ctypes
 * stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MPC_E_* code; nothing throws across the ABI;
 *     mpc_last_error(ctx) gives a human-readable message for the last failure on that context.
 *   - the caller owns all host buffers (plain pointers + sizes, little-endian, C order); the library owns all device
 *     memory.  Host buffers may be reused as soon as a call returns -- with one exception: the buffers handed to
 *     mpc_submit stay in use until the matching mpc_wait has returned.
 *   - one context <=> one device + one CUDA stream.  A context is not thread-safe; contexts are independent (one per
 *     GPU, one process per GPU in the multi-GPU harness) and share no mutable state, so several contexts on the SAME
 *     device may be driven from different host threads at once, or from ONE thread through the asynchronous pair
 *     mpc_submit / mpc_wait: the copies and staging kernels of one overlap the check kernels of another
 *     (motionplanner_b200/lanes.py).  Calls do not hold the Python GIL (ctypes releases it).
 *   - polygons must be CONVEX (occupancy polygons of the library, obstacle pieces of a scene): the separating-axis
 *     test is only correct for convex sets -- a reflex edge would "separate" what it must not.  Non-convex input is
 *     rejected with MPC_E_ARG (library: at upload; obstacle pieces: by the staging kernel, reported by
 *     mpc_set_scenes / mpc_wait); repeated vertices are tolerated.  Decompose first (checker.convex_pieces).
 *   - all geometry comes in as float64 in the caller's global frame, exactly the numbers the reference hands to
 *     shapely.  Re-centring to a scene-local frame and the float32 cast happen inside, on the device.
 *   - decisions are exact with respect to those float64 vertices: a float32 separating-axis pass with a conservative
 *     error band, a float64 pass for the pairs inside the band, and exact expansion arithmetic for float64 ties.
 *   - there is no CPU fallback: without a CUDA device mpc_create fails with MPC_E_CUDA.
 */
#ifndef MPC_H_
#define MPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_ABI_VERSION 3      /* 3: mpc_submit / mpc_wait, mpc_measure_fp32_peak, convexity is checked */

/* error codes */
#define MPC_OK 0
#define MPC_E_CUDA (-1)     /* CUDA runtime error (message in mpc_last_error) */
#define MPC_E_ARG (-2)      /* invalid argument */
#define MPC_E_STATE (-3)    /* call order violated (e.g. query before library / scenes) */
#define MPC_E_LIMIT (-4)    /* a size limit of this build was exceeded (vertex count, ...) */

/* query mode bits */
#define MPC_MODE_PLANNER 0u       /* obstacles: open interiors must not overlap  (space.contains, gap >= 0 is free) */
#define MPC_MODE_VALIDATOR 1u     /* obstacles: closed sets must not intersect   (obstacle.intersects, gap > 0 is free) */
#define MPC_MODE_NO_CULL 2u       /* skip the bounding-circle axis: every pair runs the full separating-axis test */
#define MPC_MODE_NO_EARLY_EXIT 4u /* keep evaluating a tile after its primitives are all known to collide */
#define MPC_MODE_COUNT_WORK 8u    /* instrumented launch: fill mpc_stats.axis_tests / cull_tests (slower) */
#define MPC_MODE_NO_TIMING 16u    /* no CUDA events around the kernels: mpc_stats.ms_device stays 0 (saves a few us per call) */

#define MPC_MAX_LIB_VERTS 16      /* vertices per primitive occupancy polygon (reference: 4, AROC export: <= 12) */
#define MPC_MAX_OBST_VERTS 8      /* vertices per convex obstacle piece (reference scenarios: rectangles, triangles) */

typedef struct mpc_ctx mpc_ctx;

/* One planning query: a parent pose and a set of candidate primitives appended at that pose.
 * Reference: one call of collision_check(node, primitive, space, param, timepoint) per candidate,
 * src/lowLevelPlannerManeuverAutomaton.py:95-125 (twin: src/maneuverAutomatonPlannerStandalone.py:196-226). */
typedef struct {
    double x, y;          /* node.x[0:2, ind]: rear-axle position at the start of the candidate primitive (:112)     */
    double c, s;          /* np.cos(x[3]), np.sin(x[3]) computed by the caller in float64 (:121)                      */
    int32_t scene;        /* which staged scene (free space / obstacle prediction) this query runs against          */
    int32_t t0;           /* "ind": time index of the parent state (:99)                                             */
    int32_t step_limit;   /* param['steps']: occupancies with ind+time > step_limit are skipped (:120)               */
    int32_t cand_set;     /* >= 0: candidate set staged with the library (velocity bucket, ManeuverAutomaton.py:29-32);
                             -1: explicit list in the cand_idx argument of mpc_query                                  */
    int32_t cand_off;     /* offset into that set / into cand_idx                                                    */
    int32_t cand_cnt;     /* number of candidates                                                                    */
    int32_t reserved0, reserved1;
} mpc_query_desc;

typedef struct {
    int64_t pairs_nominal;    /* candidates x occupancies in range x (obstacles + boundary segments) of their step */
    int64_t refined_fp64;     /* pairs inside the float32 error band, re-decided in float64                        */
    int64_t exact_ties;       /* orientation tests that needed exact expansion arithmetic                          */
    int64_t near_tangent;     /* refined pairs with |separation| < 1e-9 m (the stated epsilon policy)              */
    int64_t parity_refined;   /* occupancies whose inside/outside parity was re-decided in float64                 */
    int64_t cull_tests;       /* MPC_MODE_COUNT_WORK: bounding-circle axis evaluations                             */
    int64_t axis_tests;       /* MPC_MODE_COUNT_WORK: full separating-axis evaluations (pairs that passed the cull)*/
    int64_t worklist_overflow;/* 1 if the refinement list had to be grown and the query was re-run                 */
    int64_t box_skipped;      /* MPC_MODE_COUNT_WORK: pairs separated wholesale by a block bounding box            */
    int64_t done_skipped;     /* MPC_MODE_COUNT_WORK: pairs never looked at because every candidate of their group
                                 already collided at another time step (the reference's early return)              */
    float tau;                /* float32 uncertainty band used (metres)                                            */
    float ms_device;          /* device time of the kernels of this call (CUDA events on the context's stream)     */
    int32_t ctas;             /* thread blocks of the main kernel                                                  */
    int32_t kernels;          /* kernel launches issued by this call                                               */
} mpc_stats;

typedef struct {
    char name[64];
    int32_t sm_count, cc_major, cc_minor, clock_khz;
    int64_t global_mem;
    int32_t l2_bytes;
    int32_t abi_version;
} mpc_devinfo;

/* -- lifetime ------------------------------------------------------------------------------------------------------ */
int mpc_create(int device, mpc_ctx **out);
void mpc_destroy(mpc_ctx *ctx);
const char *mpc_last_error(const mpc_ctx *ctx);
int mpc_device_info(mpc_ctx *ctx, mpc_devinfo *out);

/* -- pinned host memory for the caller's input arrays (optional) -------------------------------------------------------
 * Arrays handed to mpc_set_scenes / mpc_query may live anywhere; if they live in memory obtained here the host->device
 * copies are true asynchronous DMA transfers.  The reference has no counterpart (it never leaves the host). */
int mpc_alloc_pinned(mpc_ctx *ctx, uint64_t bytes, void **out);
int mpc_free_pinned(mpc_ctx *ctx, void *ptr);

/* -- primitive library, staged once ----------------------------------------------------------------------------------
 * Replaces the per-call walk over MA.primitives[i].occ (maneuverAutomaton/ManeuverAutomaton.py:35-44; occupancy
 * polygons built by maneuverAutomaton/createManeuverAutomaton.py:76-86 or loadAROCautomaton.py:110-137).
 *   verts  [P][J][Vmax][2]  occupancy polygons in the primitive's own (rear-axle) frame, float64, any orientation
 *   nverts [P][J]           vertex count, 0 = this primitive has no occupancy in slot j
 *   tidx   [P][J]           int(o['time']/dt) resp. int(o['starttime']/dt), computed by the caller exactly as
 *                           src/lowLevelPlannerManeuverAutomaton.py:116-119 does.  Must be identical for all
 *                           primitives that use slot j (the Python side re-slots irregular libraries).
 *   set_off/set_idx         CSR of candidate sets kept on the device (the velocity buckets vel_dic,
 *                           ManeuverAutomaton.py:14-27); nsets may be 0. */
int mpc_upload_library(mpc_ctx *ctx, const double *verts, const int32_t *nverts, const int32_t *tidx, int P, int J,
                       int Vmax, const int32_t *set_off, const int32_t *set_idx, int nsets);

/* -- scenes: obstacle occupancies and free-space boundary per time step ------------------------------------------------
 * Replaces free_space() (src/maneuverAutomatonPlannerStandalone.py:151-182: road minus obstacle occupancies per step),
 * the space_all[t] list handed to lowLevelPlannerManeuverAutomaton (:19) and the per-step obstacle loop of
 * collisionChecker (auxiliary/collisionChecker.py:25-33).
 *   scene_step_off [nscenes+1]  steps of scene s are scene_step_off[s] .. scene_step_off[s+1]-1 (= len(space))
 *   origins        [nscenes][2] float64 point near the action; float32 work happens relative to it
 *   step_obst_off  [nsteps+1]   obstacles of step k are obst[step_obst_off[k] .. step_obst_off[k+1])
 *   obst           [nobst][Vo][2], obst_nv [nobst]   convex pieces (rectangles; triangles of ShapeGroups; pieces of
 *                               non-convex polygons), float64 global frame, any orientation, nv in 3..Vo
 *   step_seg_begin/end [nsteps] boundary segments of the region the occupancy must stay inside at step k:
 *                               segs[begin..end).  begin = -1: no region constraint.  begin == end >= 0: empty
 *                               region (nothing is contained, shapely: empty.contains(x) is False).
 *   segs           [nseg][4]    (ax, ay, bx, by) every ring of the region (exterior and holes of every part),
 *                               any orientation. */
int mpc_set_scenes(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
                   const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
                   const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg);

/* -- the check ---------------------------------------------------------------------------------------------------------
 * out_mask: for query q, words [mask_off[q], mask_off[q] + ceil(cand_cnt/32)) with mask_off the running sum of
 * ceil(cand_cnt/32) in query order; bit i%32 of word i/32 is 1 iff candidate i COLLIDES (collision_check would
 * return False because of geometry).  mask_words must equal the total.  stats may be NULL. */
int mpc_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand, unsigned mode,
              uint32_t *out_mask, int mask_words, mpc_stats *stats);

/* -- primitive-selection step: collision bits + child cost + goal test ------------------------------
 * The reference creates every child of an expansion with expand_node (src/lowLevelPlannerManeuverAutomaton.py:192-214:
 * x_ = T @ primitive.x + [x, y, 0, phi], cost = node.cost + np.sum((ref_traj[0:2, index] - x[0:2, index])**2)) and
 * tests popped children with goal_check (:127-142: vehicle centre inside goal['space'] within its time window).
 * mpc_expand returns both for all candidates of a batch together with the collision bits of mpc_query, in float64 and
 * in the reference's operation order (numpy's BLAS fused multiply-add chain, numpy's pairwise summation over
 * x0, y0, x1, y1, ...), so a search that orders its queue by these costs pops the same nodes as the reference.  The
 * +1000 fixed-prefix penalty (:212) stays with the caller: it is search bookkeeping.
 *
 *   mpc_upload_trajectories  states [P][L][4] = primitive.x transposed (x, y, v, phi per step, rear-axle frame),
 *                            nstates[P] = primitive.x.shape[1] (<= L <= 64); once per automaton
 *   mpc_set_guidance         ref_xy [2][M] = ref_traj[0:2]; goals = param['goal'] (time window + boundary segments of
 *                            goal['space'], seg_begin < 0 for None); b = param['b']; once per planning problem
 *   mpc_expand               queries as for mpc_query with t0 = ind (index of the parent's last state) and (x, y, c, s)
 *                            = node.x[0:2, -1], cos / sin of node.x[3, -1]; parent [B][2] = node.cost, node.x[3, -1].
 *                            out_cost / out_goal_step [sum cand_cnt] in candidate order; goal step -1 = not reached. */
typedef struct {
    int32_t time_start, time_end;   /* goal['time_start'] <= i <= goal['time_end'] */
    int32_t seg_begin, seg_end;     /* rows of goal_segs; seg_begin < 0: goal['space'] is None (always contained) */
} mpc_goal;

int mpc_upload_trajectories(mpc_ctx *ctx, const double *states, const int32_t *nstates, int P, int L);
int mpc_set_guidance(mpc_ctx *ctx, const double *ref_xy, int M, const mpc_goal *goals, int ngoals,
                     const double *goal_segs, int nseg, double b);
int mpc_expand(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
               const double *parent, unsigned mode, uint32_t *out_mask, int mask_words, double *out_cost,
               int32_t *out_goal_step, mpc_stats *stats);

/* -- asynchronous planning cycle --------------------------------------------------------------------------------------
 * One planning cycle of the reference is "new prediction in, collision bits of all candidates out"
 * (src/lowLevelPlannerManeuverAutomaton.py:19-61 per replanning step, mainCARLA.py:252-415).  mpc_submit = mpc_set_scenes
 * (skipped with nscenes = 0: the staged scenes are kept) followed by mpc_query, WITHOUT waiting for the device: it
 * returns as soon as copies and kernels are enqueued on the context's stream.  mpc_wait blocks until that batch is done
 * and hands out mask and statistics exactly as mpc_query would.  Between the two calls every host buffer passed to
 * mpc_submit must stay valid and unchanged, and no other call may be made on this context (MPC_E_STATE).  One host
 * thread can drive several contexts of one GPU this way (submit on lane k+1 while lane k computes), which is how
 * independent planning cycles overlap their transfers with each other's kernels without a host thread per context. */
int mpc_submit(mpc_ctx *ctx, int nscenes, const int32_t *scene_step_off, const double *origins,
               const int32_t *step_obst_off, const double *obst, const int32_t *obst_nv, int Vo,
               const int32_t *step_seg_begin, const int32_t *step_seg_end, const double *segs, int nseg,
               int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand, unsigned mode);
int mpc_wait(mpc_ctx *ctx, uint32_t *out_mask, int mask_words, mpc_stats *stats);

/* -- measurement helpers (bench.py) ---------------------------------------------------------------------------------
 * mpc_prepare_query stages a query batch (H2D) without running it; mpc_run_prepared launches the kernels of the
 * staged batch `iters` times on the context's stream, optionally evicting L2 (writes a buffer larger than L2)
 * before each iteration, and reports per-iteration device times measured with CUDA events around the kernels only:
 * ms_total[i] = all kernels of the query, ms_main[i] = the separating-axis kernel alone.  mpc_fetch_mask copies the
 * result mask of the last run to the host. */
int mpc_prepare_query(mpc_ctx *ctx, int B, const mpc_query_desc *queries, const int32_t *cand_idx, int ncand,
                      unsigned mode);
int mpc_run_prepared(mpc_ctx *ctx, int iters, int flush_l2, float *ms_total, float *ms_main, mpc_stats *stats);
int mpc_fetch_mask(mpc_ctx *ctx, uint32_t *out_mask, int mask_words);

/* Measured float32 arithmetic peaks of the device (the roofline denominators of bench.py; nothing of the reference
 * corresponds to this).  Register-only instruction chains on every SM, reported in TFLOP/s (multiply-add = 2 flop,
 * min / max = 1 flop per compared pair):
 *   out[0] scalar FFMA        out[1] packed FFMA2 (fma.rn.f32x2)
 *   out[2] the instruction mix of the full rectangle test as k_check runs it (8 FFMA2 : 3 FMNMX3 : 2 FMNMX, 160 flop per pair)
 *   out[3] the same arithmetic with scalar FFMA (16 FFMA : 3 FMNMX3 : 2 FMNMX)
 *   out[4] FMNMX3 alone   out[5] FMNMX alone   out[6] 8 FFMA : 4 FMNMX   out[7] SM clock seen (MHz, lower bound) */
int mpc_measure_fp32_peak(mpc_ctx *ctx, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* MPC_H_ */
