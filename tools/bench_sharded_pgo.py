"""Crossover table for the sharded two-frame LM solve (run under torchrun with N ranks, one per GPU):
single-GPU persistent kernel vs fused peer-memory all-reduce (FusedShardedPGO) vs NCCL-per-evaluation baseline, for K residual
blocks in {512 .. 65536}. Device times (CUDA events, max over ranks); rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded_pgo.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import ops, sharded_pgo as sp  # noqa: E402
from tests.golden import cases  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
fused = sp.FusedShardedPGO() if world > 1 else None


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3        # us


out = {"world": world, "rows": []}
for K in (512, 2048, 4096, 16384, 65536):
    c = cases.pgo_inputs(K, 6)
    f64 = lambda t: t.double().to(dev)
    Kt = c["K"].double()
    intr = (Kt[0, 0].item(), Kt[1, 1].item(), Kt[0, 2].item(), Kt[1, 2].item(), float(torch.tensor([c["baseline"]]).double()))
    full = [f64(c[k]) for k in ("pos_Tw", "kp2_uv", "kp2_disp", "uv_cov", "disp_cov")]
    init = f64(c["init_pose"])
    row = {"K": K}
    pose1, st1 = ops.pgo_solve(*full, intr, init)
    row["single_gpu_us"] = timed(lambda: ops.pgo_solve(*full, intr, init))
    row["steps"], row["evaluations"] = int(st1[0].item()), int(st1[1].item())
    if fused is not None:
        posef, _ = fused.solve(*full, intr, init)
        row["fused_peer_us"] = timed(lambda: fused.solve(*full, intr, init))
        row["nccl_baseline_us"] = timed(lambda: sp.solve_on_gpus(*full, intr, init), reps=2)
        row["fused_vs_single_max_abs"] = float((posef - pose1).abs().max().item())
        row["exchanges_per_solve"] = row["steps"] + row["evaluations"]
    out["rows"].append(row)
if fused is not None:
    fused.close()
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
