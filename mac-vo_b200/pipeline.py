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
