"""Per-frame driver of the hot path, for benchmarking / smoke tests without the MAC-VO tree.

It is a restatement of the data flow of `MACVO.run_pair` (Odometry/MACVO.py:173-311) restricted to what the
two-frame pose graph consumes — frontend -> keypoint selection -> in-bound filter -> per-keypoint gathers ->
observation covariances -> sanity filter -> point registration -> two-frame PGO — with the map / factor-graph
bookkeeping (Module/Map, CPU) left out: inside MAC-VO that part is unchanged and it is the CALLER of the
plugins, here the optimiser input is assembled directly.

The plugin objects are whatever implements the interfaces: the B200 plugins (`plugins.py`) on the GPU, or
the CPU oracle stand-ins (`oracle/pipeline_cpu.py`) for the CPU baseline — same driver, same order of
calls, same RNG consumption (`select_point` is called for keypoints, then for mapping points).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from .plugins import PGOInput


def quat_rotate(q: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """R(q) p for q = [x,y,z,w] (7-vector pose layout of pypose SE3)."""
    v, w = q[..., :3], q[..., 3:4]
    uv = 2 * torch.linalg.cross(v.expand_as(p), p, dim=-1)
    return p + w * uv + torch.linalg.cross(v.expand_as(p), uv, dim=-1)


def se3_act(pose: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    return quat_rotate(pose[3:7].to(p), p) + pose[:3].to(p)


@dataclass
class FrameResult:
    num_kp: int
    num_obs: int
    pose_init: torch.Tensor
    optimizer_output: object
    kp0_uv: torch.Tensor | None = None
    kp1_uv: torch.Tensor | None = None
    map_points: int = 0
    extras: dict = field(default_factory=dict)


class TwoFrameOdometry:
    """args mirror `Odometry.args` of the MAC-VO YAML: num_point, edgewidth, match_cov_default, mapping."""

    def __init__(self, frontend, kp_selector, cov_model, optimizer, num_point: int = 200, edgewidth: int = 32,
                 match_cov_default: float = 0.25, mapping: bool = True, map_selector=None, min_num_point: int = 10,
                 keep_debug: bool = False):
        self.frontend, self.kp_selector, self.cov_model, self.optimizer = frontend, kp_selector, cov_model, optimizer
        self.map_selector = map_selector
        self.num_point, self.edgewidth, self.match_cov_default = num_point, edgewidth, match_cov_default
        self.mapping, self.min_num_point, self.keep_debug = mapping, min_num_point, keep_debug
        self.prev = None            # (frame, depth output)
        self.poses: list[torch.Tensor] = []
        self._pending = False

    # --- Odometry/MACVO.py:158-171 ---------------------------------------------------------------
    def initialize(self, frame0) -> None:
        depth0 = self.frontend.estimate_depth(frame0)
        self.prev = (frame0, depth0)
        self.poses = [torch.tensor([0., 0., 0., 0., 0., 0., 1.])]

    def _write_back(self) -> None:
        """Optimizer.write_map (Odometry/MACVO.py:187): blocks on the previous frame's result."""
        if self._pending:
            res = self.optimizer.get_result()
            self.poses[-1] = res.motion.reshape(-1)[:7].detach().double().cpu().float()
            self._pending = False

    # --- Odometry/MACVO.py:173-311 ---------------------------------------------------------------
    def run_pair(self, frame1) -> FrameResult:
        frame0, depth0 = self.prev
        fe = self.frontend
        depth1, match01 = fe.estimate_pair(frame0, frame1)
        self._write_back()
        prev_pose = self.poses[-1]
        est_pose = prev_pose.clone()                                   # StaticMotionModel.predict (MotionModel.py:133-137)

        kp0_uv = self.kp_selector.select_point(frame0, self.num_point, depth0, depth1, match01)
        kp1_uv = kp0_uv + fe.retrieve_pixels(kp0_uv, match01.flow).T
        ew = self.edgewidth
        inb = ((kp1_uv[..., 0] < frame1.width - ew) & (kp1_uv[..., 0] > ew)
               & (kp1_uv[..., 1] < frame1.height - ew) & (kp1_uv[..., 1] > ew))
        kp0_uv, kp1_uv = kp0_uv[inb], kp1_uv[inb]
        num_kp = kp0_uv.size(0)

        kp0_d = fe.retrieve_pixels(kp0_uv, depth0.depth).squeeze(0)
        kp0_sigma_dd = fe.retrieve_pixels(kp0_uv, depth0.cov).squeeze(0)
        kp1_disparity = fe.retrieve_pixels(kp1_uv, depth1.disparity)
        kp1_sigma_disparity = fe.retrieve_pixels(kp1_uv, depth1.disparity_uncertainty)
        kp1_sigma_dd = fe.retrieve_pixels(kp1_uv, depth1.cov).squeeze(0)

        dev = kp0_uv.device
        kp0_sigma_uv = torch.ones((num_kp, 3), device=dev) * self.match_cov_default
        kp0_sigma_uv[..., 2] = 0.
        kp1_sigma_uv = fe.retrieve_pixels(kp0_uv, match01.cov).T.contiguous()

        K = frame0.frame_K.to(dev)
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        pos0_Tc = torch.stack([kp0_d, (kp0_uv[:, 0] - cx) / fx * kp0_d, (kp0_uv[:, 1] - cy) / fy * kp0_d], dim=-1)  # NED
        pos0_cov = self.cov_model.estimate(frame0, kp0_uv, depth0, kp0_sigma_dd, kp0_sigma_uv)
        pos1_cov = self.cov_model.estimate(frame1, kp1_uv, depth1, kp1_sigma_dd, kp1_sigma_uv)

        # CovarianceSanityFilter (Module/OutlierFilter.py:91-100)
        bad = (pos0_cov.isnan().any(dim=(-1, -2)) | pos0_cov.isinf().any(dim=(-1, -2))
               | pos1_cov.isnan().any(dim=(-1, -2)) | pos1_cov.isinf().any(dim=(-1, -2)))
        num_obs = int((~bad).sum())                                    # `bad` lives on the host like the covariances: no sync
        keep = (~bad).to(dev)
        pos_Tw = se3_act(prev_pose.to(dev), pos0_Tc.float())[keep]

        self.poses.append(est_pose)
        out = None
        if num_obs >= self.min_num_point:
            inp = PGOInput(pos_Tw=pos_Tw, kp2_uv=kp1_uv[keep].float(), kp2_disp=kp1_disparity.T[keep],
                           uv_cov=kp1_sigma_uv[keep], disp_cov=kp1_sigma_disparity.T[keep], K=frame1.frame_K,
                           baseline=frame1.frame_baseline, init_pose=est_pose)
            self.optimizer.start_optimize(inp)
            self._pending = True
            out = self.optimizer.get_result() if hasattr(self.optimizer, "optimize_res") else None

        n_map = 0
        if self.mapping and self.map_selector is not None:            # Odometry/MACVO.py:314-337
            map0_uv = self.map_selector.select_point(frame0, 2000, depth0, depth1, match01)
            n_map = map0_uv.size(0)
            if n_map:
                map0_sigma_dd = fe.retrieve_pixels(map0_uv, depth0.cov).squeeze(0)
                map0_sigma_uv = torch.ones((n_map, 3), device=map0_uv.device) * self.match_cov_default
                map0_sigma_uv[..., 2] = 0.
                self.cov_model.estimate(frame0, map0_uv, depth0, map0_sigma_dd, map0_sigma_uv)

        self.prev = (frame1, depth1)
        res = FrameResult(num_kp=num_kp, num_obs=num_obs, pose_init=est_pose, optimizer_output=out, map_points=n_map)
        if self.keep_debug:
            res.kp0_uv, res.kp1_uv = kp0_uv, kp1_uv
            res.extras = {"depth1": depth1, "match01": match01, "pos0_cov": pos0_cov, "pos1_cov": pos1_cov,
                          "pos_Tw": pos_Tw, "keep": keep}
        return res

    def finish(self) -> torch.Tensor:
        """Synchronise on the last optimisation and return all poses (F, 7)."""
        self._write_back()
        return torch.stack(self.poses)


class FusedTwoFrameOdometry:
    """Same per-frame data flow as `TwoFrameOdometry`, with everything between the frontend and the optimiser result kept
    on the device: observation building, CovarianceSanityFilter and MatchObs packing are two launches
    (`ops.observe_pack`, csrc/observe.cu), the LM kernel reads the survivor count on the device
    (`ops.pgo_solve_counted`), the optimised pose of frame t is consumed by frame t+1 without visiting the host, and a
    frame's observations / mapping points / pose travel to pinned host memory in asynchronous copies.

    Host synchronisations per frame: ONE — the two candidate counts that `torch.randperm` needs on the CPU default
    generator (kept for bit-exact keypoints; drawn in MAC-VO's order: keypoints, then mapping points). `TwoFrameOdometry`
    with the plugin-API calls has >= 7 (candidate counts x2, boolean indexing x2, covariance `.cpu()` x3).

    Requires the B200 plugins (uses their device buffers); keypoints and poses equal `TwoFrameOdometry`'s
    (tests/test_gpu_pipeline.py::test_fused_tail_matches_plugin_path)."""

    def __init__(self, frontend, kp_selector, cov_model, optimizer, num_point: int = 200, edgewidth: int = 32,
                 match_cov_default: float = 0.25, mapping: bool = True, map_selector=None, min_num_point: int = 10,
                 num_map_point: int = 2000, keep_debug: bool = False, solver=None):
        from . import ops
        self.ops = ops
        # solver(obs, intr5, pose_io, stats, min_k): the LM solve on the packed observation buffer; default = one persistent
        # launch on this GPU. bench.py --config sharded plugs in the multi-GPU solve (broadcast + sharded_pgo.FusedShardedPGO)
        self.solver = solver
        self.frontend, self.kp_selector, self.cov_model, self.optimizer = frontend, kp_selector, cov_model, optimizer
        self.map_selector = map_selector
        self.num_point, self.edgewidth, self.match_cov_default = num_point, edgewidth, match_cov_default
        self.mapping, self.min_num_point, self.num_map_point = mapping and map_selector is not None, min_num_point, num_map_point
        self.keep_debug = keep_debug
        self.device = kp_selector.device
        cc = cov_model.config
        self.cov_args = dict(kernel_size=cc.kernel_size, min_flow_cov=cc.min_flow_cov, min_depth_cov=cc.min_depth_cov)
        self.cluster = int(getattr(optimizer, "context", {}).get("cluster", 0)) if hasattr(optimizer, "context") else 0
        self.obs = [ops.ObservationBuffers(num_point, self.device) for _ in range(2)]      # double buffered
        self.stats = [torch.zeros((8,), dtype=torch.float64, device=self.device) for _ in range(2)]
        if self.mapping:
            self.map_cov = [torch.empty((num_map_point, 3, 3), dtype=torch.float64, device=self.device) for _ in range(2)]
            self.map_cov_host = [torch.empty((num_map_point, 3, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
            self.map_pt_host = [torch.empty((num_map_point, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.pose_dev: list[torch.Tensor] = []          # (7,) float64 per frame, on the device
        self.pose_host = torch.zeros((2, 7), dtype=torch.float64).pin_memory()
        self.pose_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.n_map = [0, 0]
        self.frame_no = 0
        self.prev = None
        self.last: FrameResult | None = None
        self._prefetched = None         # (frame, depth, match) of a frontend launched ahead by run_pair(..., next_frame=)
        self._tail_stream = None
        self._tail_done = None

    def initialize(self, frame0) -> None:
        depth0 = self.frontend.estimate_depth(frame0)
        self.prev = (frame0, depth0)
        self.pose_dev = [torch.tensor([0., 0., 0., 0., 0., 0., 1.], dtype=torch.float64, device=self.device)]

    @staticmethod
    def _intr(frame) -> tuple[float, float, float, float]:
        K = frame.frame_K if hasattr(frame, "frame_K") else frame.K[0]
        return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])

    @torch.inference_mode()
    def run_pair(self, frame1, next_frame=None) -> FrameResult:
        """One frame. With `next_frame` (the frame the NEXT call will pass) the frames are software-pipelined: its frontend —
        which depends on images only, not on this frame's pose — is enqueued before the host waits for this frame's candidate
        counts, and this frame's tail (sampling gathers, observation building, LM, mapping covariances, result copies) runs on
        a second stream next to it, so neither the tail nor the host round trip of the one synchronisation leaves the GPU idle.
        Results are identical (same kernels, same `randperm` order)."""
        ops = self.ops
        frame0, depth0 = self.prev
        slot = self.frame_no & 1
        if self._prefetched is not None and self._prefetched[0] is frame1:
            depth1, match01 = self._prefetched[1:]
        else:
            depth1, match01 = self.frontend.estimate_pair(frame0, frame1)
        self._prefetched = None
        main = torch.cuda.current_stream()
        if self._tail_done is not None:            # the previous frame's tail read the candidate lists the selectors now overwrite
            main.wait_event(self._tail_done)
        # selection kernels for keypoints AND mapping points first, then a single synchronisation for both counts
        reqs = [(self.kp_selector.enqueue_candidates(match01), self.num_point)]
        if self.mapping:
            reqs.append((self.map_selector.enqueue_candidates(depth0), self.num_map_point))
        counts = ops.request_candidate_counts(reqs)
        if next_frame is None:
            counts.synchronize()
            return self._tail(frame0, frame1, depth0, depth1, match01, reqs, slot)
        self._prefetched = (next_frame, *self.frontend.estimate_pair(frame1, next_frame))
        counts.synchronize()
        if self._tail_stream is None:
            self._tail_stream = torch.cuda.Stream(self.device)
        tail = self._tail_stream
        tail.wait_event(counts)
        for t in (match01.flow, match01.cov, depth0.depth, depth1.depth, depth1.disparity, depth1.disparity_uncertainty):
            t.record_stream(tail)              # allocated on the main stream, consumed on the tail stream
        with torch.cuda.stream(tail):
            res = self._tail(frame0, frame1, depth0, depth1, match01, reqs, slot)
            self._tail_done = torch.cuda.Event()
            self._tail_done.record(tail)
        return res

    def _tail(self, frame0, frame1, depth0, depth1, match01, reqs, slot) -> FrameResult:
        """everything after the candidate counts reached the host, on the current stream"""
        ops = self.ops
        picks = ops.sample_from_counts(reqs)
        kp0_uv = picks[0]
        obs, stats = self.obs[slot], self.stats[slot]
        next_pose = torch.empty((7,), dtype=torch.float64, device=self.device)
        i0, i1 = self._intr(frame0), self._intr(frame1)
        ops.observe_pack(obs, kp0_uv, match01.flow, match01.cov, depth0.depth, depth1.depth, depth1.disparity,
                         depth1.disparity_uncertainty, self.edgewidth, i0, i1, self.pose_dev[-1], next_pose,
                         match_cov_default=self.match_cov_default, **self.cov_args)
        bl = float(torch.as_tensor(frame1.frame_baseline, dtype=torch.float32).double().reshape(-1)[0])
        stats.zero_()
        if self.solver is not None:
            self.solver(obs, (*i1, bl), next_pose, stats, self.min_num_point)
        else:
            ops.pgo_solve_counted(obs, (*i1, bl), next_pose, stats, min_k=self.min_num_point, cluster=self.cluster)
        self.pose_dev.append(next_pose)
        n_map = 0
        if self.mapping:
            map0_uv = picks[1]
            n_map = map0_uv.size(0)
            if n_map:   # constant quantisation covariance for manually selected pixels, clamped like any flow_cov
                sig = max(float(torch.tensor(self.match_cov_default, dtype=torch.float32)),
                          float(torch.tensor(self.cov_args["min_flow_cov"], dtype=torch.float32) ** 2))
                _, pt, _ = ops.match_covariance(map0_uv, depth0.depth, None, *i0, kernel_size=self.cov_args["kernel_size"],
                                                min_flow_cov=self.cov_args["min_flow_cov"],
                                                min_depth_cov=self.cov_args["min_depth_cov"], match_cov_default=sig,
                                                want_point=True, out_cov=self.map_cov[slot][:n_map])
                self.map_cov_host[slot][:n_map].copy_(self.map_cov[slot][:n_map], non_blocking=True)
                self.map_pt_host[slot][:n_map].copy_(pt, non_blocking=True)
        self.n_map[slot] = n_map
        # one asynchronous copy ships the frame's MatchObs columns; the pose follows; nothing waits here
        obs.download_async()
        self.pose_host[slot].copy_(next_pose, non_blocking=True)
        self.pose_ready[slot].record()
        self.prev = (frame1, depth1)
        self.frame_no += 1
        res = FrameResult(num_kp=-1, num_obs=-1, pose_init=None, optimizer_output=None, map_points=n_map)
        if self.keep_debug:
            res.kp0_uv = kp0_uv
            res.extras = {"depth1": depth1, "match01": match01, "slot": slot}
        self.last = res
        return res

    def latest_pose(self) -> torch.Tensor:
        """Optimised pose of the newest frame on the HOST (float64 (7,)); waits for that frame's pose copy only."""
        slot = (self.frame_no - 1) & 1
        self.pose_ready[slot].synchronize()
        return self.pose_host[slot].clone()

    def observations(self) -> dict:
        """The newest frame's packed observations from pinned host memory (waits for its copy)."""
        slot = (self.frame_no - 1) & 1
        obs = self.obs[slot]
        obs.ready.synchronize()
        hdr = obs.section("header", host=True)
        n = int(hdr[0])
        out = {k: obs.section(k, host=True)[:n].clone() for k in
               ("pos_Tw", "pixel2_uv", "pixel2_disp", "pixel2_uv_cov", "pixel2_disp_cov", "obs1_covTc", "obs2_covTc",
                "pixel1_uv", "pixel1_d")}
        out.update(num_obs=n, num_kp=int(hdr[1]), num_selected=int(hdr[2]), status=int(hdr[3]))
        if self.mapping:
            self.pose_ready[slot].synchronize()
            m = self.n_map[slot]
            out.update(map_cov=self.map_cov_host[slot][:m].clone(), map_pos_Tc=self.map_pt_host[slot][:m].clone())
        return out

    def finish(self) -> torch.Tensor:
        """All poses (F, 7) float32 like `TwoFrameOdometry.finish` (the map stores fp32 poses)."""
        if self._tail_done is not None:
            self._tail_done.synchronize()
        torch.cuda.current_stream().synchronize()
        return torch.stack([p.cpu().float() for p in self.pose_dev])
