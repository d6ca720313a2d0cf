#!/usr/bin/env python
"""BASELINE.json configs[4] / SURVEY.md 8(d) C5: a batch of 2000 scenario-shaped planning queries, sharded over the
GPUs of one box (one process per GPU, replicated library, contiguous blocks, only result bits gathered).

Each query = one search expansion on a bundled scenario drawn with replacement (seed 3): the scenario's road
boundary and predicted obstacle occupancies, the ego's initial pose, all primitives of its velocity bucket
(<= 195 candidates x 11 occupancies).  `--full-library` tests all 13 702 primitives per query instead.

    python tools/bench_c5.py [--queries 2000] [--full-library] [--check]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_c5.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import load_golden  # noqa: E402
from motionplanner_b200.checker import SceneSet  # noqa: E402
from motionplanner_b200.engine import CollisionEngine, make_queries, unpack_mask  # noqa: E402
from motionplanner_b200.maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton  # noqa: E402
from motionplanner_b200.sharding import block_bounds  # noqa: E402


def build(MA, golden, nq, seed, full_library, lo, hi):
    """scenes + queries of the block [lo, hi) of the global batch (every rank draws the same global sequence)"""
    rng = np.random.default_rng(seed)
    names = sorted(golden)
    draw = rng.integers(0, len(names), nq)
    lay = MA.device_layout(0.1)
    scenes = SceneSet()
    q = make_queries(hi - lo)
    cand = None
    for k, gi in enumerate(draw[lo:hi]):
        g = golden[names[gi]]
        n = min(g['nsteps'], 31)                               # 3 s horizon like the CARLA loop (mainCARLA.py:57-59)
        off = g['step_obst_off'][:n + 1]
        segs = g['road_safe'] if len(g['road_safe']) else g['road']
        scenes.add_scene_explicit(n, off, g['obst'][:off[-1]], g['obst_nv'][:off[-1]], segs, origin=g['x0'][0:2])
        x0 = g['x0']
        bucket = int(MA._bucket(x0[2]))
        q[k] = (x0[0] - np.cos(x0[3]) * 1.15, x0[1] - np.sin(x0[3]) * 1.15, np.cos(x0[3]), np.sin(x0[3]), k, 0,
                g['t_end'], bucket, 0, len(MA.vel_dic[bucket]), 0, 0)
    if full_library:
        cand = np.arange(len(MA.primitives), dtype=np.int32)
        q['cand_set'], q['cand_off'], q['cand_cnt'] = -1, 0, len(cand)
    return lay, scenes, q, cand


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--queries', type=int, default=2000)
    ap.add_argument('--full-library', action='store_true')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--check', action='store_true', help='compare rank 0 block with the CPU oracle')
    ap.add_argument('--pageable', action='store_true', help='keep the host arrays in pageable memory')
    ap.add_argument('--lanes', type=int, default=3, help='contexts per GPU for the second end-to-end figure: the rank\'s '
                    'block is cut into this many sub-blocks, staged and checked concurrently (CollisionLanes)')
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    import torch
    import torch.distributed as dist
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    MA = load_maneuver_automaton()
    golden = load_golden()
    b = block_bounds(args.queries, world)
    lay, scenes, q, cand = build(MA, golden, args.queries, 3, args.full_library, int(b[rank]), int(b[rank + 1]))
    eng = CollisionEngine(local)
    eng.upload_library(lay['verts'], lay['nv'], lay['tidx'], lay['set_off'], lay['set_idx'])
    arrays = scenes.arrays()
    if not args.pageable:           # per-cycle inputs in pinned host memory: the copies are real DMA
        arrays = {k: eng.pinned_like(v) for k, v in arrays.items()}
        q = eng.pinned_like(q)

    def stage():
        eng.set_scenes(arrays['scene_step_off'], arrays['origins'], arrays['step_obst_off'], arrays['obst'],
                       arrays['obst_nv'], arrays['step_seg_begin'], arrays['step_seg_end'], arrays['segs'])

    stage()
    eng.prepare(q, cand, 0)
    eng.run_prepared(3, flush_l2=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms_total, ms_main, st = eng.run_prepared(args.iters, flush_l2=True)
    bits = eng.fetch()
    # end to end: scenes + queries from host buffers every iteration
    for _ in range(2):
        stage()
        eng.query_mask(q, cand, 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_stage = 0.0
    for _ in range(args.iters):
        ts = time.perf_counter()
        stage()
        t_stage += time.perf_counter() - ts
        mask, _ = eng.query_mask(q, cand, 0)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / args.iters
    t_stage /= args.iters
    # the same block cut into sub-blocks, one per lane: copies / host passes of one overlap the kernels of another
    e2e_lanes = float('nan')
    if args.lanes > 1 and not args.pageable:
        from motionplanner_b200.lanes import CollisionLanes
        pool = CollisionLanes(local, lanes=args.lanes)
        pool.upload_library(lay['verts'], lay['nv'], lay['tidx'], lay['set_off'], lay['set_idx'])
        sub = block_bounds(int(b[rank + 1] - b[rank]), args.lanes) + int(b[rank])
        cycles = []
        for k in range(args.lanes):
            _, sc_k, q_k, _ = build(MA, golden, args.queries, 3, args.full_library, int(sub[k]), int(sub[k + 1]))
            cycles.append(({n: pool.pinned_like(v) for n, v in sc_k.arrays().items()}, pool.pinned_like(q_k)))
        pool.run_cycles(cycles * 2, 0, cand)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pool.run_cycles(cycles * args.iters, 0, cand)
        torch.cuda.synchronize()
        e2e_lanes = (time.perf_counter() - t0) / args.iters
        got = np.concatenate([unpack_mask(m, c[1]) for (m, _), c in zip(res[-args.lanes:], cycles)])
        if not np.array_equal(got, bits):
            raise SystemExit('bench_c5: sub-blocks through lanes disagree with the single batch')
        pool.close()
    times = torch.tensor([float(np.mean(ms_total)) * 1e-3, e2e, e2e_lanes], dtype=torch.float64, device='cuda')
    tot = torch.tensor([float(st['pairs_nominal']), float(bits.sum()), float(len(bits)), float(st['refined_fp64']),
                        float(st['near_tangent'])], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(np.packbits(bits, bitorder='little'), gathered, dst=0)       # result bitmasks only
    disagreements = None
    if args.check and rank == 0:
        from oracle import oracle as orc
        lib = dict(verts=lay['verts'], nv=lay['nv'], tidx=lay['tidx'], set_off=lay['set_off'], set_idx=lay['set_idx'])
        want, _ = orc.Problem(lib, arrays).check(q, cand, orc.MODE_NO_EARLY_EXIT)
        disagreements = int((want != bits).sum())
    if rank == 0:
        dev_s, e2e_s, e2e_lanes_s = [float(v) for v in times.cpu()]
        pairs, coll, ncand, refined, near = [float(v) for v in tot.cpu()]
        line = {"workload": "C5: %d scenario-shaped queries (seed 3), %s" % (args.queries, "full library 13702" if args.full_library else "velocity bucket <=195 candidates"),
                "n_gpus": world, "queries_per_s_device": args.queries / dev_s, "pair_checks_per_s_device": pairs / dev_s,
                "queries_per_s_e2e": args.queries / e2e_s, "ms_device": 1e3 * dev_s, "ms_e2e": 1e3 * e2e_s, "ms_stage_scenes_rank0": 1e3 * t_stage,
                "lanes": args.lanes, "ms_e2e_lanes": 1e3 * e2e_lanes_s, "queries_per_s_e2e_lanes": args.queries / e2e_lanes_s,
                "pairs_nominal": pairs, "candidates": ncand, "colliding": coll, "refined_fp64": refined,
                "near_tangent_lt_1e-9": near, "disagreements_vs_oracle_rank0": disagreements,
                "h2d_bytes": int(sum(v.nbytes for v in arrays.values()) + q.nbytes)}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
