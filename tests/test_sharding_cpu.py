"""N > 1 host path on CPU: two gloo ranks shard a query batch, each decides its block (oracle standing in for the
per-GPU engine), rank 0 gathers the bitmasks; the result must equal the single-process answer."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from helpers import OracleBackend
    from motionplanner_b200 import workloads as W
    from motionplanner_b200.sharding import run_sharded
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lib, sc, org, q = W.adversarial_problem(42, nqueries=7)
    be = OracleBackend()
    be.upload_library(lib['verts'], lib['nv'], lib['tidx'], lib['set_off'], lib['set_idx'])
    be.set_scenes(sc['scene_step_off'], org, sc['step_obst_off'], sc['obst'], sc['obst_nv'], sc['step_seg_begin'],
                  sc['step_seg_end'], sc['segs'])
    bits, stats = run_sharded(be, q, None, 0)
    if rank == 0:
        ret['bits'] = bits
        ret['nstats'] = len(stats)
        ret['calls'] = len(be.log[0][0])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    from helpers import OracleBackend
    from motionplanner_b200 import workloads as W
    from motionplanner_b200.sharding import block_bounds, shard_queries
    assert block_bounds(7, 2).tolist() == [0, 4, 7] and block_bounds(3, 4).tolist() == [0, 1, 2, 3, 3]
    lib, sc, org, q = W.adversarial_problem(42, nqueries=7)
    assert len(shard_queries(q, 1, 2)) == 3
    be = OracleBackend()
    be.upload_library(lib['verts'], lib['nv'], lib['tidx'], lib['set_off'], lib['set_idx'])
    be.set_scenes(sc['scene_step_off'], org, sc['step_obst_off'], sc['obst'], sc['obst_nv'], sc['step_seg_begin'],
                  sc['step_seg_end'], sc['segs'])
    want, _ = be.query(q, None, 0)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29631, ret), nprocs=2, join=True)
    assert ret['nstats'] == 2 and ret['calls'] == 4            # rank 0 decided 4 of the 7 queries
    assert np.array_equal(ret['bits'], want)
