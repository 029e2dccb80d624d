"""N > 1 path on CPU: world_size-2 gloo. Each rank accumulates its shard of residual blocks (the oracle
stands in for the CUDA accumulate kernel, which has no CPU build), all-reduces the 55-double accumulator
and runs the identical LM logic; the result must equal the single-process oracle solve."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, K: int, seed: int, out_q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from macvo_b200 import sharded_pgo as sp
    from oracle import pgo as opgo
    from tests.golden import cases
    c = cases.pgo_inputs(K, seed)
    g = cases.pgo_graph(c)
    lo, hi = sp.shard_bounds(K, world, rank)
    shard = opgo.GraphData(pos_Tw=g.pos_Tw[lo:hi], kp2_uv=g.kp2_uv[lo:hi], kp2_disp=g.kp2_disp[lo:hi],
                           uv_cov=g.uv_cov[lo:hi], disp_cov=g.disp_cov[lo:hi], fx=g.fx, fy=g.fy, cx=g.cx, cy=g.cy,
                           baseline=g.baseline, init_pose=g.init_pose)

    def allreduce(acc):
        t = torch.from_numpy(np.ascontiguousarray(acc))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    pose, stats = sp.lm_solve_sharded(lambda p: opgo.accumulate_packed(shard, p), allreduce, g.init_pose)
    out_q.put((rank, pose, stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("K,seed", [(512, 6), (200, 9)])
def test_sharded_lm_two_ranks_equals_single_process(K, seed):
    from oracle import pgo as opgo
    from tests.golden import cases
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    trace = opgo.LMTrace()
    ref = opgo.lm_solve(cases.pgo_graph(cases.pgo_inputs(K, seed)), trace=trace)
    np.testing.assert_array_equal(results[0][1], results[1][1])          # every rank holds the same bits
    np.testing.assert_allclose(results[0][1], ref, rtol=1e-8, atol=1e-10)
    assert results[0][2]["steps"] == trace.steps and results[0][2]["evaluations"] == trace.evaluations
    assert results[0][2]["collectives"] == trace.steps + trace.evaluations


def test_shard_bounds_partition():
    from macvo_b200 import sharded_pgo as sp
    for K in (0, 1, 7, 4096):
        for world in (1, 2, 3, 4, 8):
            b = [sp.shard_bounds(K, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == K and all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_packed_accumulator_is_additive_over_shards():
    from oracle import pgo as opgo
    from tests.golden import cases
    g = cases.pgo_graph(cases.pgo_inputs(200, 9))
    pose = opgo.se3_exp(np.array([0.02, -0.01, 0.03, 0.004, -0.003, 0.002]))
    full = opgo.accumulate_packed(g, pose)
    parts = np.zeros(55)
    for lo, hi in ((0, 77), (77, 200)):
        sh = opgo.GraphData(pos_Tw=g.pos_Tw[lo:hi], kp2_uv=g.kp2_uv[lo:hi], kp2_disp=g.kp2_disp[lo:hi],
                            uv_cov=g.uv_cov[lo:hi], disp_cov=g.disp_cov[lo:hi], fx=g.fx, fy=g.fy, cx=g.cx, cy=g.cy,
                            baseline=g.baseline, init_pose=g.init_pose)
        parts += opgo.accumulate_packed(sh, pose)
    np.testing.assert_allclose(parts, full, rtol=1e-11, atol=1e-9)
