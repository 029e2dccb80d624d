"""TEST INFRASTRUCTURE — CPU stand-ins for the five plugins, built from the oracle functions.

They let `macvo_b200.pipeline.TwoFrameOdometry` run the reference's CPU arithmetic end to end
(`bench.py --impl reference`, the `cpu_baseline` leg, and the end-to-end parity tests). The dense
network layers are the SAME torch code as the product's (`flowformer_cov.py`, validated against the
reference network by tests/golden/net_*.pt) executed on the CPU with the oracle's correlation volume
(torch.bmm) and window lookup (grid_sample) injected — i.e. exactly what the reference runs on a CPU.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import covariance as ocov
from . import frontend as ofe
from . import keypoint as okp
from . import pgo as opgo


class _Out(SimpleNamespace):
    pass


class CpuFrontend:
    def __init__(self, state_dict, enc_dtype=torch.float32, dec_dtype=torch.float32, decoder_depth: int = 12,
                 enforce_positive_disparity: bool = False):
        from macvo_b200.flowformer_cov import FlowFormerCovNet
        self.net = FlowFormerCovNet(state_dict, "cpu", enc_dtype, dec_dtype, decoder_depth,
                                    corr_fn=ofe.corr_volume, lookup_fn=ofe.window_lookup)
        self.epd = enforce_positive_disparity

    @property
    def provide_cov(self):
        return True, True

    def _post(self, flow, cov, frame):
        d = ofe.dense_postproc(flow.float(), cov.float(), frame.frame_baseline, frame.fx, self.epd)
        depth = _Out(depth=d["depth"], cov=d["depth_cov"], disparity=d["disparity"],
                     disparity_uncertainty=d["disparity_uncertainty"], mask=d["depth_mask"])
        match = _Out(flow=d["flow"], cov=d["flow_cov"], mask=None)
        return depth, match

    @torch.inference_mode()
    def estimate_depth(self, frame):
        flow, cov = self.net.inference(frame.imageL, frame.imageR)
        return self._post(torch.cat([flow, flow]), torch.cat([cov, cov]), frame)[0]

    @torch.inference_mode()
    def estimate_pair(self, f1, f2):
        flow, cov = self.net.inference(torch.cat([f2.imageL, f1.imageL]), torch.cat([f2.imageR, f2.imageL]))
        return self._post(flow, cov, f2)

    @staticmethod
    def retrieve_pixels(pixel_uv, scalar_map, interpolate=False):
        return ofe.retrieve_pixels(pixel_uv, scalar_map)


class CpuSelector:
    def __init__(self, kernel_size=7, mask_width=32, max_match_cov=100.0):
        self.k, self.mw, self.mc = kernel_size, mask_width, max_match_cov

    def select_point(self, frame, numPoint, depth0, depth1, match):
        return okp.cov_aware_select_nodepth(match.cov, numPoint, self.k, self.mw, self.mc, match.mask)


class CpuMapSelector:
    def __init__(self, max_depth=5.0, max_depth_cov=0.005, mask_width=32):
        self.md, self.mdc, self.mw = max_depth, max_depth_cov, mask_width

    def select_point(self, frame, numPoint, depth0, depth1, match):
        return okp.mapping_select(depth0.depth, depth0.cov, numPoint, self.md, self.mdc, self.mw)


class CpuCovariance:
    def __init__(self, kernel_size=31, min_flow_cov=0.25, min_depth_cov=0.05, match_cov_default=0.25):
        self.a = (kernel_size, min_flow_cov, min_depth_cov, match_cov_default)

    def estimate(self, frame, kp, depth_est, depth_cov, flow_cov):
        return ocov.match_covariance(kp, depth_est.depth, flow_cov, frame.fx, frame.fy, frame.cx, frame.cy, *self.a)


class CpuPGO:
    def __init__(self):
        self.optimize_res = None
        self.trace = None

    def start_optimize(self, inp):
        K = inp.K.double().numpy().reshape(3, 3)
        g = opgo.GraphData(
            pos_Tw=inp.pos_Tw.double().numpy(), kp2_uv=inp.kp2_uv.double().numpy(),
            kp2_disp=inp.kp2_disp.double().numpy().reshape(-1), uv_cov=inp.uv_cov.double().numpy(),
            disp_cov=inp.disp_cov.double().numpy().reshape(-1), fx=float(K[0, 0]), fy=float(K[1, 1]),
            cx=float(K[0, 2]), cy=float(K[1, 2]),
            baseline=float(torch.as_tensor(inp.baseline, dtype=torch.float32).double().reshape(-1)[0]),
            init_pose=inp.init_pose.double().numpy().reshape(7))
        self.trace = opgo.LMTrace()
        pose = opgo.lm_solve(g, trace=self.trace)
        self.optimize_res = _Out(motion=torch.tensor(np.asarray(pose)).reshape(1, 7))

    def get_result(self):
        return self.optimize_res
