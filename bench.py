#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched motion-primitive collision check (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--queries Q]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3] "synthetic dense traffic: 4096 primitives x 256 obstacles x 50
timesteps per query, 1 GPU" (seeds 0/1/2).  A *step* is one pass of the hot path over one batch of Q
such queries (Q ego poses with jitter against the same predicted scene, all 4096 candidates each).  One *check* is one
(primitive occupancy, obstacle, time step) pair decided, so a step is Q x 4096 x 256 x 50 checks; the 4 road-boundary
segments per (primitive, step) are decided too but not counted.

  value   device-resident throughput: scene, library and queries already in HBM, CUDA-event time of every kernel of
          the step (memsets + k_check + k_refine), L2 flushed before every timed step.
  e2e     same metric through the public call a user makes (CollisionEngine.set_scenes + .query, i.e. mpc_set_scenes +
          mpc_query of include/mpc.h) with HOST numpy buffers: per step the float64 obstacle prediction and the poses
          go host->device and the packed result mask comes back, all inside the timed region.
  N > 1   one process per GPU, replicated library, independent queries per rank (weak scaling), no collective on the
          hot path; times are max over ranks.

--impl reference times the reference's CPU algorithm for the same path: the float64 oracle port (oracle/oracle_c.c,
OpenMP over all host cores) because shapely/commonroad are not installable here.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, O_LANES, SLOTS, T = 4096, 8, 32, 50
O = O_LANES * SLOTS
CHECKS_PER_QUERY = P * O * T
FLOP_PER_CHECK = 304            # rectangle x rectangle SAT, a = 8 axes, a*(4a+6)
SAT_FLOP_EXECUTED = 160         # this kernel's full 4x4 test: 64 FMA (2 flop) + 24 min + 8 max
CULL_FLOP_EXECUTED = 6          # centre-line axis per pair: 2 add + 2 fma (packed FADD2/FFMA2: two pairs per instruction)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # no FP32 entry in MEASURED_PEAKS.json: derived from its sm_max_mhz
NCU_DRAM_BYTES_PER_LAUNCH = 4896512    # ncu capture n (profiles/): dram__bytes_read.sum of the two k_check launches (scout 0.46 MB + main 4.44 MB) of one 64-query step; write 0
NCU_WARP_INST_PER_LAUNCH = 54931122   # same capture: smsp__inst_executed.sum of both launches (scout 10.99 M + main 43.94 M; a count, not a time)
ISSUE_PEAK = 148 * 4 * 1.965e9         # warp instructions per second: 4 schedulers per SM, one issue per clock
METRIC = "primitive x obstacle x timestep collision checks/sec"
WORKLOAD = ("C4 dense traffic: 4096 primitives x 256 obstacles x 50 timesteps per query (BASELINE.json configs[3])")


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons during the timed region (NVML, ~20 ms period)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def build_problem(nq, rank):
    from motionplanner_b200 import workloads as W
    lib = W.synthetic_library(P, T, seed=0)
    scene = W.dense_traffic_scene(T, O_LANES, SLOTS, seed=1)
    queries, origins = W.dense_traffic_queries(scene, P, nq, seed=2 + 1000 * rank)
    return lib, scene, origins, queries


def cpu_baseline(lib, scene, queries, budget_s=8.0):
    """oracle port (reference semantics: float64, exact signs, early return per primitive) on all host cores"""
    from oracle import oracle as orc
    pb = orc.Problem(lib, scene)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    pb.check(queries[:1], None, 0, nthreads=cores, inner_parallel=True)
    one = time.perf_counter() - t0
    n = int(max(1, min(len(queries) * 50, budget_s / max(one, 1e-3))))
    qs = np.concatenate([queries] * (n // len(queries) + 1))[:n]
    t0 = time.perf_counter()
    _, st = pb.check(qs, None, 0, nthreads=cores, inner_parallel=True)
    dt = time.perf_counter() - t0
    return {"value": n * CHECKS_PER_QUERY / dt, "unit": "checks/s", "cores": cores, "kind": "port",
            "sample": "%d C4 queries (4096x256x50 each) through oracle/oracle_c.c, OpenMP over candidates, reference "
                      "early return per primitive; %.2f s" % (n, dt),
            "pairs_decided_frac": st["pairs_decided"] / max(st["pairs_nominal"], 1)}


def run_reference(args, rank, world):
    """--impl reference: the CPU algorithm of the path (oracle port) on this box's host cores, rank 0 only"""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    lib, scene, origins, queries = build_problem(args.ref_queries, 0)
    pb = orc.Problem(lib, scene)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        pb.check(queries, None, 0, nthreads=cores, inner_parallel=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, st = pb.check(queries, None, 0, nthreads=cores, inner_parallel=True)
    dt = time.perf_counter() - t0
    checks = args.steps * len(queries) * CHECKS_PER_QUERY
    v = checks / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "checks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_per_step": args.queries, "seeds": [0, 1, 2],
                       "sample_queries_per_step": len(queries)},
            "cpu_baseline": {"value": v, "unit": "checks/s", "cores": cores, "kind": "port",
                             "sample": "each step = %d of the %d C4 queries of a step through oracle/oracle_c.c (float64, "
                                       "exact signs, reference early return), OpenMP %d threads"
                                       % (len(queries), args.queries, cores)},
            "e2e": {"value": v, "unit": "checks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "shapely/commonroad-io are not installable offline, so the reference's CPU path is timed through "
                    "its float64 restatement (kind=port)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=64,
                    help="queries (ego poses) per step; default = the 64 jittered poses SURVEY.md 8(d) defines for C4")
    ap.add_argument("--ref-queries", type=int, default=2, help="queries per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--lanes", type=int, default=4,
                    help="contexts per GPU for the end-to-end figure (1 = one context, strictly serial steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # NCCL and friends may print to stdout: keep fd 1 for the single JSON line, send everything else to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the collision check has no CPU fallback (use --impl reference for "
                         "the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from motionplanner_b200._cabi import MODE_NO_TIMING
    from motionplanner_b200.engine import (MODE_COUNT_WORK, MODE_NO_CULL, MODE_NO_EARLY_EXIT, MODE_PLANNER,
                                           CollisionEngine)
    eng = CollisionEngine(local)
    info = eng.device_info()
    lib, scene, origins, queries = build_problem(args.queries, rank)
    nq = len(queries)
    eng.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])

    # the step's inputs live in pinned host memory (what a sensor/prediction pipeline would hand over)
    for k in ("obst", "obst_nv", "segs", "step_obst_off", "step_seg_begin", "step_seg_end", "scene_step_off"):
        scene[k] = eng.pinned_like(scene[k])
    origins = eng.pinned_like(origins)
    queries = eng.pinned_like(queries)

    def set_scene():
        eng.set_scenes(scene["scene_step_off"], origins, scene["step_obst_off"], scene["obst"], scene["obst_nv"],
                       scene["step_seg_begin"], scene["step_seg_end"], scene["segs"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    set_scene()
    checks_step = nq * CHECKS_PER_QUERY

    # ---- device-resident throughput (value) -------------------------------------------------------------------------
    eng.prepare(queries, None, MODE_PLANNER)
    eng.run_prepared(args.warmup, flush_l2=True)
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ms_total, ms_main, st = eng.run_prepared(args.steps, flush_l2=True)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.result()
    dev_s = float(ms_total.sum()) * 1e-3
    main_ms = float(ms_main.mean())
    collide = eng.fetch()

    # warm-L2 variant (what a planner sees between expansions: library and scene stay in the 126 MB L2)
    ms_warm, ms_main_warm, _ = eng.run_prepared(max(10, args.steps // 4), flush_l2=False)

    # single-query latency on the device (B=1)
    eng.prepare(queries[:1], None, MODE_PLANNER)
    eng.run_prepared(3, flush_l2=False)
    ms_one, _, _ = eng.run_prepared(50, flush_l2=False)

    # instrumented launches: executed work, and the full-evaluation variant (no culling axis, no early exit)
    eng.prepare(queries, None, MODE_PLANNER | MODE_COUNT_WORK)
    _, _, st_cnt = eng.run_prepared(1, flush_l2=False)
    eng.prepare(queries[:max(1, nq // 8)], None, MODE_NO_CULL | MODE_NO_EARLY_EXIT)
    eng.run_prepared(2, flush_l2=True)
    ms_full, ms_full_main, st_full = eng.run_prepared(5, flush_l2=True)
    full_checks = max(1, nq // 8) * CHECKS_PER_QUERY

    # ---- end to end through the public API with host buffers (e2e) ----------------------------------------------------
    h2d = scene["obst"].nbytes + scene["obst_nv"].nbytes + scene["segs"].nbytes + scene["step_obst_off"].nbytes + \
        2 * scene["step_seg_begin"].nbytes + queries.nbytes + origins.nbytes + scene["scene_step_off"].nbytes
    d2h = ((P + 31) // 32) * 4 * nq + 72
    for _ in range(args.warmup):
        set_scene()
        eng.query_mask(queries, None, MODE_PLANNER | MODE_NO_TIMING)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        set_scene()
        mask, _ = eng.query_mask(queries, None, MODE_PLANNER | MODE_NO_TIMING)
    barrier()
    e2e_serial_s = time.perf_counter() - t0
    # the same steps dealt to `--lanes` contexts on this GPU, one host thread each (motionplanner_b200/lanes.py): every
    # step still copies its scene and queries host -> device and its result bits back inside the timed region; the
    # copies and staging kernels of one lane overlap k_check of the other
    e2e_s, e2e_lanes = e2e_serial_s, 1
    # every lane is a host thread that waits on its stream: no more lanes than this rank's share of the host cores
    args.lanes = max(1, min(args.lanes, len(os.sched_getaffinity(0)) // world))
    if args.lanes > 1:
        from motionplanner_b200.lanes import CollisionLanes
        pool = CollisionLanes(local, lanes=args.lanes)
        pool.upload_library(lib["verts"], lib["nv"], lib["tidx"], lib["set_off"], lib["set_idx"])
        cycle = (dict(scene, origins=origins), queries)
        pool.run_cycles([cycle] * (args.lanes * args.warmup), MODE_PLANNER | MODE_NO_TIMING)
        barrier()
        t0 = time.perf_counter()
        res = pool.run_cycles([cycle] * args.steps, MODE_PLANNER | MODE_NO_TIMING)
        barrier()
        e2e_lanes_s = time.perf_counter() - t0
        if not all(np.array_equal(r[0], mask) for r in res[-args.lanes:]):
            raise SystemExit("bench.py: lanes and single context disagree")
        pool.close()
        if e2e_lanes_s < e2e_serial_s:
            e2e_s, e2e_lanes = e2e_lanes_s, args.lanes
    # planning-step latency: one query (B=1) against an already staged scene, host doubles in -> bitmask out
    lat = []
    q1 = queries[:1]
    q1_ptr, q1_words = q1.ctypes.data_as(ctypes.c_void_p), (P + 31) // 32
    for _ in range(200):
        t1 = time.perf_counter()
        eng.query_fast(q1, MODE_PLANNER | MODE_NO_TIMING, nw=q1_words, qptr=q1_ptr)    # the call the planner loop makes
        lat.append(time.perf_counter() - t1)
    lat_p50 = float(np.median(lat[10:]))

    # ---- reduce over ranks (max time), gather the result bitmasks -----------------------------------------------------
    times = torch.tensor([dev_s, e2e_s, wall, e2e_serial_s], dtype=torch.float64, device="cuda")
    nfree = torch.tensor([float(len(collide) - int(collide.sum()))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        masks = [torch.empty(len(mask), dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(mask.view(np.int32)).cuda(), masks, dst=0)      # only result bitmasks travel
        dist.all_reduce(nfree)
    dev_s, e2e_s, wall, e2e_serial_s = [float(x) for x in times.cpu()]

    if rank == 0:
        value = world * args.steps * checks_step / dev_s
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        algo_tflops = checks_step * FLOP_PER_CHECK / (main_ms * 1e-3) / 1e12
        exec_flop = st_cnt["cull_tests"] * CULL_FLOP_EXECUTED + st_cnt["axis_tests"] * SAT_FLOP_EXECUTED
        exec_tflops = exec_flop / (main_ms * 1e-3) / 1e12
        # algorithmic HBM bytes of one launch: library tile (80 B per occupancy: vertices, edge normals, circle), the
        # scene (96 B per obstacle record), masks; shared by the Q queries of the step
        algo_bytes = P * T * 80 + T * O * 96 + nq * ((P + 31) // 32) * 4 + nq * 64
        full_main = float(ms_full_main.mean())
        line = {
            "metric": METRIC, "value": value, "unit": "checks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_per_step": nq,
                       "queries_what": "one step = one batch of ego poses (seed 2 jitter) against one staged scene",
                       "checks_per_step": checks_step, "seeds": [0, 1, 2], "l2": "value: flushed before every timed step "
                       "(256 MiB fill, outside the event bracket); e2e: wall clock over whole cycles, no flush (the "
                       "library stays L2-resident between cycles as in a planner's steady state; warm_l2 gives the "
                       "device time without the flush, ~5 % below the flushed one)", "early_exit": "reference semantics (return at the first violation): warp-wide per tile, and whole "
                       "groups / CTAs whose candidates already collide at another time step skip their tile; the "
                       "no-exit, no-cull rate is roofline.full_evaluation", "parallelism": "replicated library, "
                       "queries sharded per GPU, no collective on the hot path"},
            "e2e": {"value": world * args.steps * checks_step / e2e_s, "unit": "checks/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "lanes": e2e_lanes,
                    "what": "CollisionEngine.set_scenes + query_mask (mpc_set_scenes + mpc_query) from pinned host "
                            "numpy buffers, every step" + ("" if e2e_lanes == 1 else "; steps dealt round-robin to %d "
                            "contexts on the GPU, one host thread each (CollisionLanes.run_cycles), so one step's "
                            "copies, staging kernels and launch boundaries overlap another's k_check; no L2 flush "
                            "between steps here, which is why it can exceed `value` (single stream, L2 flushed)"
                            % e2e_lanes),
                    "single_context": {"value": world * args.steps * checks_step / e2e_serial_s,
                                       "ms_per_step": 1e3 * e2e_serial_s / args.steps}},
            "gpu_launches": int(args.steps * (1 + st["kernels"] // (args.steps + 1))),
            "gpu_launches_what": "per timed step: k_fill_u32 (L2 flush, outside the event bracket), k_check (scout launch "
                                 "of the first time slot + the rest of the grid when the batch qualifies), k_refine",
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "kernel": "k_check<4,4,0,0,0>", "achieved": exec_tflops, "peak": FP32_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": exec_tflops / FP32_PEAK_TFLOPS,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if nq == 64 else None,
                "what": "executed FP32 flop of the timed launch (an instrumented launch of the same batch counts "
                        "centre-line axis tests x 6 flop and full separating-axis tests x 160 flop) / CUDA-event "
                        "duration of k_check (scout launch + main launch). The kernel is NOT FP32-bound and is not meant to "
                        "be: 93 % of the nominal pairs are never looked at (their candidate already collides at another "
                        "time step -- the reference's early return), sub-block bounding boxes separate most of the "
                        "rest, the centre-line axis all but 0.04 %. ncu (profiles/r01n_k_check_ncu_full.csv, both "
                        "launches): instruction issue 54-57 % of the active cycles (see issue_slots), ALU pipe 33-36 %, "
                        "FMA pipe 14-15 %; HBM traffic is below the algorithmic bytes because skipped tiles are never "
                        "fetched. The arithmetic rate of the full test is full_evaluation (no culling, no exits): "
                        "55-57 % of the FP32 peak.",
                "peak_source": "derived: 148 SM x 128 lanes x 2 x sm_max_mhz 1965 (MEASURED_PEAKS.json has no FP32 "
                               "figure)",
                "kernel_ms": main_ms,
                "executed": {"cull_axis_tests": st_cnt["cull_tests"], "full_sat_tests": st_cnt["axis_tests"],
                             "pairs_separated_by_block_box": st_cnt["box_skipped"],
                             "pairs_skipped_after_collision_at_another_step": st_cnt["done_skipped"],
                             "pairs_nominal": st_cnt["pairs_nominal"], "flop": exec_flop,
                             "pairs_looked_at_frac": (st_cnt["cull_tests"] + st_cnt["box_skipped"]) / max(st_cnt["pairs_nominal"], 1),
                             "note": "instrumented launch of the same batch; which tiles are skipped depends on "
                                     "scheduling, the decisions do not"},
                "algorithmic_survey": {"flop_per_check": FLOP_PER_CHECK, "achieved": algo_tflops,
                                       "frac": algo_tflops / FP32_PEAK_TFLOPS,
                                       "note": "SURVEY.md 8(d) definition: 304 flop x checks / time. Above 1 by "
                                               "construction: early return, sub-block boxes and the centre-line axis (6 flop) "
                                               "decide 99.96 % of the pairs, so the algorithmic count is not executed."},
                "full_evaluation": {"checks_per_s": full_checks / (full_main * 1e-3),
                                    "achieved": full_checks * SAT_FLOP_EXECUTED / (full_main * 1e-3) / 1e12,
                                    "frac": full_checks * SAT_FLOP_EXECUTED / (full_main * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                                    "frac_survey_304": full_checks * FLOP_PER_CHECK / (full_main * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                                    "note": "MPC_MODE_NO_CULL|NO_EARLY_EXIT: every pair runs the complete 4x4 "
                                            "separating-axis test (160 executed flop: 64 FMA + 32 min/max); here "
                                            "algorithmic = executed"},
                "issue_slots": {"achieved": (NCU_WARP_INST_PER_LAUNCH / (main_ms * 1e-3) / 1e9) if nq == 64 else None,
                                "peak": ISSUE_PEAK / 1e9, "unit": "G warp-instructions/s",
                                "frac": (NCU_WARP_INST_PER_LAUNCH / (main_ms * 1e-3) / ISSUE_PEAK) if nq == 64 else None,
                                "note": "the binding resource per ncu: executed warp instructions of one launch (count "
                                        "from the committed capture) / live CUDA-event duration, against 148 SM x 4 "
                                        "schedulers x sm_max_mhz"},
                "ncu": {"source": "profiles/r01n_k_check_ncu_full.csv", "launches": "scout (512 CTAs, 24.3 us under ncu) "
                        "+ main (3136 CTAs, 79.4 us)", "smsp__issue_active_pct": [56.8, 53.8],
                        "l1tex__data_pipe_lsu_wavefronts_pct": [22.4, 30.5], "sm__inst_executed_pipe_fma_pct": [15.2, 13.8],
                        "sm__inst_executed_pipe_alu_pct": [35.5, 33.2], "sm__warps_active_pct": [38.0, 33.9],
                        "smsp__inst_executed": NCU_WARP_INST_PER_LAUNCH,
                        "dram__bytes_read": NCU_DRAM_BYTES_PER_LAUNCH, "dram__bytes_write": 0,
                        "note": "which tiles are skipped depends on scheduling: counts vary by a few percent between launches"},
                "hbm": {"algorithmic_bytes": algo_bytes, "achieved_gbs": algo_bytes / (main_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "frac": algo_bytes / (main_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "measured" if peaks else "fallback"}},
            "warm_l2": {"value": world * checks_step / (float(ms_warm.mean()) * 1e-3), "ms_per_step": float(ms_warm.mean())},
            "latency": {"planning_step_p50_ms": 1e3 * lat_p50, "device_ms_single_query": float(np.median(ms_one)),
                        "what": "one query (4096 candidates x 50 steps x 256 obstacles), host doubles in -> bitmask out"},
            "refinement": {"refined_fp64": st["refined_fp64"], "exact_ties": st["exact_ties"],
                           "near_tangent_lt_1e-9": st["near_tangent"], "parity_refined": st["parity_refined"],
                           "tau_m": st["tau"]},
            "result": {"free_candidates": float(nfree.item()), "candidates": world * nq * P},
            "wall_s_timed_region": wall, "device": info["name"], "sm_count": info["sm_count"],
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            line["cpu_baseline"] = cpu_baseline(lib, scene, queries)
        else:
            line["cpu_baseline"] = None
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
