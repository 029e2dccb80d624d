"""Two-frame PGO with the residual blocks SHARDED across ranks (BASELINE config 4: 1280x720, 4096
keypoints, 4 x B200).

Every rank owns a contiguous shard of the K residual blocks and reduces it to the 55-double packed
accumulator  [A (21) | b (6) | G (21) | h (6) | loss]  with one kernel launch (`ops.pgo_accumulate`);
ONE all-reduce(SUM) of those 440 bytes per evaluation (NCCL over NVLink on GPUs, gloo in the CPU tests)
gives every rank the same totals, and every rank then runs the identical tiny 6x6 solve / trust-region /
accept-reject logic below (a restatement of LM_analytic.step, Module/Optimization/PyposeOptimizers.py:160-194,
TrustRegion and StopOnPlateau of pypose 0.6.8) redundantly — no broadcast of the pose is needed.

At 440 B per collective the exchange is pure latency (~10-20 us per evaluation, <= ~30 evaluations per
frame); for K <= 4096 the single-GPU persistent kernel (`ops.pgo_solve`, ~0.8 ms at K = 4096) is faster —
the crossover is reported in DESIGN.md. This path exists because the north star asks for it and it is the
building block for pose graphs that do not fit one launch.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

NACC = 55
_IU = np.triu_indices(6)


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def _quat_rot(q, p):
    v, w = q[:3], q[3]
    uv = 2 * np.cross(v, p)
    return p + w * uv + np.cross(v, uv)


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def se3_retract(pose: np.ndarray, step: np.ndarray) -> np.ndarray:
    """pose <- Exp(step[:6]) * pose  (pypose left retraction; [t, q_xyzw] layout)."""
    tau, phi = step[:3], step[3:6]
    t2 = float(phi @ phi)
    th = np.sqrt(t2)
    eps = np.finfo(np.float64).eps
    if th > eps:
        c1, c2 = (1 - np.cos(th)) / t2, (th - np.sin(th)) / (t2 * th)
        imag, real = np.sin(0.5 * th) / th, np.cos(0.5 * th)
    else:
        c1, c2 = 0.5 - t2 / 24 + t2 * t2 / 720, 1.0 / 6 - t2 / 120 + t2 * t2 / 5040
        imag, real = 0.5 - t2 / 48 + t2 * t2 / 3840, 1 - t2 / 8 + t2 * t2 / 384
    K = _skew(phi)
    t = (np.eye(3) + c1 * K + c2 * (K @ K)) @ tau
    q = np.concatenate([phi * imag, [real]])
    return np.concatenate([_quat_rot(q, pose[:3]) + t, _quat_mul(q, pose[3:])])


def shard_bounds(k: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous, deterministic partition of K residual blocks"""
    return (k * rank) // world, (k * (rank + 1)) // world


def lm_solve_sharded(accumulate: Callable[[np.ndarray], np.ndarray], allreduce: Callable[[np.ndarray], np.ndarray],
                     init_pose: np.ndarray, max_steps: int = 10, patience: int = 2, decreasing: float = 1e-5,
                     radius: float = 1e3, reject: int = 16, diag_min: float = 1e-6, diag_max: float = 1e32):
    """accumulate(pose) -> this rank's packed accumulator (55,), allreduce(x) -> elementwise sum over ranks.
    Returns (pose (7,), stats dict). Identical control flow on every rank."""
    TR_MIN, TR_MAX, HIGH, LOW, UP, DOWN, FACTOR = 1e-3, 1e5, 0.5, 1e-3, 2.0, 0.5, 0.5
    pose = np.asarray(init_pose, dtype=np.float64).copy()
    damping, down = 1.0 / radius, 0.5
    loss = None
    steps = patience_count = evals = collectives = 0
    while True:
        tot = allreduce(accumulate(pose)); collectives += 1
        A = np.zeros((6, 6)); G = np.zeros((6, 6))
        A[_IU] = tot[:21]; A = A + A.T - np.diag(np.diag(A))
        G[_IU] = tot[27:48]; G = G + G.T - np.diag(np.diag(G))
        b, h = tot[21:27], tot[48:54]
        if loss is None:
            loss = float(tot[54])
        last, reject_count = loss, 0
        A[np.diag_indices(6)] = np.clip(np.diag(A), diag_min, diag_max)
        while last <= loss:
            A[np.diag_indices(6)] = np.diag(A) + np.diag(A) * damping
            D = np.linalg.solve(A, b)
            trial = se3_retract(pose, D)
            loss = float(allreduce(accumulate(trial))[54]); collectives += 1; evals += 1
            with np.errstate(divide="ignore", invalid="ignore"):
                quality = (last - loss) / -(2.0 * (D @ h) + D @ G @ D)
            rad = 1.0 / damping
            if quality > HIGH:
                rad, down = rad * UP, DOWN
            elif quality > LOW:
                down = DOWN
            else:
                rad, down = rad * down, down * FACTOR
            down = max(TR_MIN, min(down, TR_MAX))
            rad = max(TR_MIN, min(rad, TR_MAX))
            damping = 1.0 / rad
            if last < loss and reject_count < reject:
                pose = se3_retract(trial, -D)
                loss, reject_count = last, reject_count + 1
            else:
                pose = trial
                break
        steps += 1
        cont = steps < max_steps
        patience_count = patience_count + 1 if (last - loss) < decreasing else 0
        if patience_count >= patience or reject_count >= reject:
            cont = False
        if not cont:
            break
    return pose, {"steps": steps, "evaluations": evals, "collectives": collectives, "loss": loss}


class FusedShardedPGO:
    """The product path for BASELINE config 4: residual blocks sharded across the ranks of a `torch.distributed` group
    (one process per GPU, one node), the whole Levenberg-Marquardt loop ONE persistent launch per rank with the
    all-reduce of the 55-double accumulator fused into it over NVLink peer memory (`ops.pgo_solve_sharded`,
    csrc/pgo.cu::exchange_ranks). torch.distributed is used once, at construction, to exchange the 64-byte CUDA IPC
    handles of the exchange buffers; the solves themselves issue no collective call."""

    def __init__(self, group=None):
        import torch.distributed as dist
        from . import ops
        self.ops, self.group = ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.exchange = ops.PeerExchange(self.world, self.rank)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.exchange.handle, group=group)
        self.exchange.connect(handles)
        dist.barrier(group)

    def solve(self, pos_Tw, kp2_uv, kp2_disp, uv_cov, disp_cov, intr, init_pose, cluster: int = 0):
        """every rank passes the FULL (K, .) CUDA float64 arrays (e.g. after a broadcast of the packed observation buffer)
        and works on its contiguous shard; returns (pose (7,) float64 CUDA, stats) — the same bits on every rank."""
        lo, hi = shard_bounds(pos_Tw.shape[0], self.world, self.rank)
        shard = [t[lo:hi].contiguous() for t in (pos_Tw, kp2_uv, kp2_disp, uv_cov, disp_cov)]
        return self.ops.pgo_solve_sharded(shard, intr, init_pose, self.exchange, cluster=cluster)

    # ---- one stream, N GPUs (BASELINE config 4): rank `src` owns the frontend and the observation buffer ----------------
    def solve_packed(self, obs, intr, pose_io, stats, min_k: int, src: int = 0):
        """Collective over the group. Rank `src` passes its filled `ops.ObservationBuffers` + the initial pose in `pose_io`;
        the other ranks pass their own (same capacity) buffers as receive space. Two NCCL broadcasts ship the five LM input
        arrays (80 B per keypoint slot) and [survivor count | initial pose]; then every rank solves its capacity-shard with the
        fused peer-memory all-reduce; `pose_io` holds the same optimised pose on every rank afterwards."""
        import torch
        import torch.distributed as dist
        c = obs.capacity
        if getattr(self, "_meta", None) is None:
            self._meta = torch.zeros(8, dtype=torch.float64, device=obs.packed.device)
            self._n = torch.zeros(1, dtype=torch.int32, device=obs.packed.device)
        meta = self._meta
        if self.rank == src:
            meta[0:1] = obs.n_obs
            meta[1:8] = pose_io
        dist.broadcast(obs.packed[:10 * c], src=src, group=self.group)
        dist.broadcast(meta, src=src, group=self.group)
        if self.rank != src:
            pose_io.copy_(meta[1:8])
        self._n.copy_(meta[0:1])
        lo, hi = shard_bounds(c, self.world, self.rank)
        shard = [obs.section("pos_Tw")[lo:hi], obs.section("pixel2_uv")[lo:hi], obs.section("pixel2_disp")[lo:hi],
                 obs.section("pixel2_uv_cov")[lo:hi], obs.section("pixel2_disp_cov")[lo:hi]]
        self.ops.pgo_solve_sharded(shard, intr, None, self.exchange, k_total=self._n, k_offset=lo, min_k=min_k,
                                   pose_io=pose_io, stats=stats)
        return pose_io

    def close(self) -> None:
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(self.group)          # no rank unmaps / frees a buffer a peer's kernel may still write
        self.exchange.close()


def solve_on_gpus(pos_Tw, kp2_uv, kp2_disp, uv_cov, disp_cov, intr, init_pose, group=None):
    """NCCL baseline of `FusedShardedPGO` (host-driven LM loop, one `dist.all_reduce` + one device->host read per evaluation):
    every rank passes the FULL (K, .) CUDA float64 tensors (or its own copy),
    works on its shard and all-reduces over NCCL. Returns (pose tensor (7,) float64 on the device, stats)."""
    import torch
    import torch.distributed as dist
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(pos_Tw.shape[0], world, rank)
    shard = [t[lo:hi].contiguous() for t in (pos_Tw, kp2_uv, kp2_disp, uv_cov, disp_cov)]
    dev = pos_Tw.device

    def accumulate(pose_np):
        return ops.pgo_accumulate(*shard, intr, torch.from_numpy(pose_np).to(dev))

    def allreduce(acc):
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)      # 440 bytes over NVLink
        return acc.cpu().numpy()

    pose, stats = lm_solve_sharded(accumulate, allreduce, init_pose.detach().cpu().numpy().reshape(7))
    return torch.from_numpy(pose).to(dev), stats
