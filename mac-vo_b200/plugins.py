"""B200 plugin classes for MAC-VO's `Module` interfaces (the drop-in boundary).

    B200_FlowFormerCovFrontend       IFrontend          replaces CUDAGraph_FlowFormerCovFrontend (Frontend.py:264-353)
    B200_CovAwareSelector_NoDepth    IKeypointSelector  replaces CovAwareSelector_NoDepth (KeypointSelector.py:349-407)
    B200_CovAwareSelector            IKeypointSelector  replaces CovAwareSelector (KeypointSelector.py:250-347)
    B200_MappingPointSelector        IKeypointSelector  replaces MappingPointSelector (KeypointSelector.py:78-100)
    B200_MatchCovariance             ICovariance2to3    replaces MatchCovariance (Covariance/Project2to3.py:114-182)
    B200_CovarianceSanityFilter      IObservationFilter replaces CovarianceSanityFilter (OutlierFilter.py:91-100)
    B200_MotionInterpolate           IMapProcessor      replaces MotionInterpolate (MapProcessor.py:52-79)
    B200_TwoFrame_PGO                IOptimizer         replaces TwoFrame_PGO (Optimization/TwoFramePGO/Optimizer.py:23-108)

Same constructor signature (`__init__(config: SimpleNamespace)`), same `is_valid_config` contract
(unknown keys are an error), same argument meaning and return types as the classes they replace; class
names are new because the registry forbids duplicates. With MAC-VO importable they subclass MAC-VO's own
interfaces (see `interfaces.py`); YAML selects them with `type: B200_...` — see INTEGRATION.md.

All compute goes through `ops` (ctypes -> libmacvo_b200.so). No CPU fallback: constructing a plugin with a
non-CUDA device raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from types import SimpleNamespace

import torch

from . import interfaces as _local
from . import ops
from .flowformer_cov import FlowFormerCovNet, synthetic_state_dict

_REF = _local.reference_available()
if _REF:   # subclass MAC-VO's own interfaces so that importing this module registers the plugins there
    from DataLoader import StereoData  # type: ignore
    from Module.Frontend.Frontend import IFrontend  # type: ignore
    from Module.Frontend.StereoDepth import IStereoDepth  # type: ignore
    from Module.Frontend.Matching import IMatcher  # type: ignore
    from Module.KeypointSelector import IKeypointSelector  # type: ignore
    from Module.Covariance.Project2to3 import ICovariance2to3  # type: ignore
    from Module.OutlierFilter import IObservationFilter  # type: ignore
    from Module.MapProcessor import IMapProcessor  # type: ignore
    from Module.Optimization.TwoFramePGO.Optimizer import TwoFrame_PGO as _PGOBase  # type: ignore
    from Module.Optimization.TwoFramePGO.Graphs import GraphOutput as _RefGraphOutput  # type: ignore
else:
    StereoData = _local.StereoData
    IFrontend, IStereoDepth, IMatcher = _local.IFrontend, _local.IStereoDepth, _local.IMatcher
    IKeypointSelector, ICovariance2to3 = _local.IKeypointSelector, _local.ICovariance2to3
    IObservationFilter, IMapProcessor = _local.IObservationFilter, _local.IMapProcessor
    _PGOBase = _local.IOptimizer

_DTYPES = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def _require_cuda(device: str, who: str) -> torch.device:
    if "cuda" not in str(device):
        raise ValueError(f"{who}: the B200 plugins only run on a CUDA device (got device={device!r}); "
                         "use the reference classes for CPU")
    return torch.device(device)


# ================================================================================================
# Frontend
# ================================================================================================
class B200_FlowFormerCovFrontend(IFrontend):
    """FlowFormerCov stereo + flow frontend: correlation volume and window lookup on sm_100a kernels,
    dense post-processing + keypoint scoring fused into one pass, the whole `estimate_pair` replayed as a
    CUDA graph (like CUDAGraph_FlowFormerCovFrontend, Frontend.py:301-353).

    config: weight (checkpoint path, or "synthetic[:seed]" for the deterministic stand-in), device,
    enc_dtype / dec_dtype in {fp32, fp16, bf16}, decoder_depth, enforce_positive_disparity,
    cuda_graph (bool), score_kernel_size (odd int; NMS window of the fused keypoint scoring, 7 in the
    reference configs)."""

    def __init__(self, config: SimpleNamespace):
        super().__init__(config)
        self.device = _require_cuda(config.device, "B200_FlowFormerCovFrontend")
        w = config.weight
        if isinstance(w, str) and w.startswith("synthetic"):
            sd = synthetic_state_dict(int(w.split(":")[1]) if ":" in w else 0)
        else:
            sd = torch.load(w, map_location="cpu", weights_only=True)
        enc, dec = _DTYPES[config.enc_dtype], _DTYPES[config.dec_dtype]
        # MACVO_Fast (enc fp16 / dec bf16, Config/Experiment/MACVO/MACVO_Fast.yaml:8-9) exists because half-precision tensor
        # cores are the fast path on the GPUs MAC-VO targets. On B200 the TF32 pipeline of this class (own kernels + TF32
        # cuDNN / cuBLAS) is both faster than a half-precision torch-op network and ~4x closer to exact arithmetic than the
        # reference's own fp16 / bf16 run (flow 9e-4 vs 3.3e-3 of its scale, tests/test_gpu_pipeline.py::test_fast_config_*),
        # so half-precision configs are served by it unless `half_precision: native` asks for the literal dtypes.
        self.half_precision = getattr(config, "half_precision", "tf32")
        if self.half_precision == "tf32":
            enc = dec = torch.float32
        self.net = FlowFormerCovNet(sd, self.device, enc, dec, decoder_depth=config.decoder_depth)
        # the reference frontend enables TF32 tensor cores for the dense layers (Frontend.py:275-277)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        if os.environ.get("MACVO_B200_CUDNN_BENCHMARK") == "1":      # experiment switch: cuDNN autotuning of the conv algorithms
            torch.backends.cudnn.benchmark = True
        torch.set_float32_matmul_precision("medium")
        self._graph = None
        self._static: dict = {}
        self._score: ops.ScoreBuffers | None = None
        self.dedup_shared_image = os.environ.get("MACVO_B200_DEDUP", "1") != "0"

    @property
    def provide_cov(self) -> tuple[bool, bool]:
        return True, True

    # ---- (a7) + fused (a8) scoring ---------------------------------------------------------------
    def _postprocess(self, est_flow, est_cov, bl_fx: float):
        H, W = est_flow.shape[-2:]
        if self._score is None or (self._score.h, self._score.w) != (H, W):
            self._score = ops.ScoreBuffers(H, W, self.device, int(getattr(self.config, "score_kernel_size", 7)))
        return ops.dense_postproc(est_flow, est_cov, bl_fx, self.config.enforce_positive_disparity, score=self._score)

    def _run(self, input_A, input_B, bl_fx: float):
        # input_A = [t2.L, t1.L], input_B = [t2.R, t2.L]: B[1] is A[0], which the feature encoder then sees once
        est_flow, est_cov = self.net.inference(input_A, input_B, shared=(0, 1) if self.dedup_shared_image else None)
        return self._postprocess(est_flow.float(), est_cov.float(), bl_fx)

    def _outputs(self, d: dict, clone: bool):
        c = (lambda t: t.clone() if t is not None else None) if clone else (lambda t: t)
        depth = IStereoDepth.Output(depth=c(d["depth"]), cov=c(d["depth_cov"]), disparity=c(d["disparity"]),
                                    disparity_uncertainty=c(d["disparity_uncertainty"]), mask=c(d["depth_mask"]))
        match = IMatcher.Output(flow=c(d["flow"]), cov=c(d["flow_cov"]), mask=None)
        # let the selector plugin reuse the fused scores instead of re-reading the covariance map
        self._score.generation += 1
        # (the token is keyed by buffer generation + map address; the covariance map must not be edited in place between
        # estimate_pair and select_point — inference tensors carry no version counter that could detect it)
        match._b200_score = (self._score, self._score.generation, match.cov.data_ptr())  # type: ignore[attr-defined]
        return depth, match

    @torch.inference_mode()
    def estimate_depth(self, frame: StereoData):
        A = frame.imageL.to(self.device, non_blocking=True)
        B = frame.imageR.to(self.device, non_blocking=True)
        est_flow, est_cov = self.net.inference(A, B)
        est_flow, est_cov = est_flow.float(), est_cov.float()
        # B = 1 call: reuse the pair kernel by duplicating the slot (slot 1 outputs are ignored)
        d = ops.dense_postproc(torch.cat([est_flow, est_flow]), torch.cat([est_cov, est_cov]),
                               frame.frame_baseline * frame.fx, self.config.enforce_positive_disparity, score=None)
        return IStereoDepth.Output(depth=d["depth"], cov=d["depth_cov"], disparity=d["disparity"],
                                   disparity_uncertainty=est_cov[0:1, :1], mask=d["depth_mask"])

    @torch.inference_mode()
    def estimate_pair(self, frame_t1: StereoData, frame_t2: StereoData):
        """-> (IStereoDepth.Output of t2, IMatcher.Output t1 -> t2); batches [t2.L, t1.L] vs [t2.R, t2.L]
        exactly like the reference (Frontend.py:284-285)."""
        bl_fx = frame_t2.frame_baseline * frame_t2.fx
        if self._graph is not None:
            # steady state: the four images go straight into the graph's static input buffers (no host-side concatenation;
            # from pinned host memory these are asynchronous copies that overlap the previous frame's tail)
            st = self._static
            shape = (2,) + tuple(frame_t2.imageL.shape[1:])
            assert shape == st["shape"], f"Input shape mismatch for CUDAGraph replay: {shape} != {st['shape']}"
            assert bl_fx == st["bl_fx"], "camera baseline * fx changed since the CUDA graph was captured"
            # each image crosses PCIe once: t1.L is last call's t2.L and already sits in A[0] (device->device move), and
            # B[1] = t2.L is copied from A[0] after the upload
            if st.get("last_t2L") is frame_t1.imageL:
                st["A"][1:2].copy_(st["A"][0:1])
            else:
                st["A"][1:2].copy_(frame_t1.imageL, non_blocking=True)
            st["A"][0:1].copy_(frame_t2.imageL, non_blocking=True)
            st["B"][0:1].copy_(frame_t2.imageR, non_blocking=True)
            st["B"][1:2].copy_(st["A"][0:1])
            st["last_t2L"] = frame_t2.imageL
            self._graph.replay()
            ops.LAUNCHES[0] += st["launches"]
            return self._outputs(st["out"], clone=True)
        input_A = torch.cat([frame_t2.imageL, frame_t1.imageL], dim=0)
        input_B = torch.cat([frame_t2.imageR, frame_t2.imageL], dim=0)
        if not getattr(self.config, "cuda_graph", True):
            out = self._run(input_A.to(self.device, non_blocking=True), input_B.to(self.device, non_blocking=True), bl_fx)
            return self._outputs(out, clone=False)
        # first call: warm up (cuDNN autotune, workspace growth) on a side stream, then capture one frame
        sA = torch.empty(input_A.shape, dtype=torch.float32, device=self.device)
        sB = torch.empty_like(sA)
        sA.copy_(input_A)
        sB.copy_(input_B)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._run(sA, sB, bl_fx)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES[0]
        with torch.cuda.graph(graph):
            out = self._run(sA, sB, bl_fx)
        self._graph = graph
        self._static = {"A": sA, "B": sB, "out": out, "shape": tuple(input_A.shape), "bl_fx": bl_fx,
                        "launches": ops.LAUNCHES[0] - n0,     # macvo_b200 kernels inside one replay
                        "last_t2L": frame_t2.imageL}
        graph.replay()                             # (the reference returns its warm-up result here)
        return self._outputs(out, clone=True)

    @staticmethod
    def retrieve_pixels(pixel_uv: torch.Tensor, scalar_map: torch.Tensor | None, interpolate: bool = False):
        """(a9) gather kernel; same contract as IFrontend.retrieve_pixels (Frontend.py:104-118)."""
        if scalar_map is None:
            return None
        if interpolate:
            raise NotImplementedError("Not implemented yet")
        if scalar_map.is_cuda and scalar_map.dtype == torch.float32 and pixel_uv.dtype in (torch.int64, torch.float32):
            return ops.retrieve_pixels(pixel_uv.to(scalar_map.device), scalar_map)
        return scalar_map[0, ..., pixel_uv[..., 1].long(), pixel_uv[..., 0].long()]   # e.g. the CPU colour image

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        spec = {
            "weight": lambda s: isinstance(s, str),
            "device": lambda s: isinstance(s, str) and "cuda" in s,
            "dec_dtype": lambda b: b in ("fp32", "fp16", "bf16"),
            "enc_dtype": lambda b: b in ("fp32", "fp16", "bf16"),
            "enforce_positive_disparity": lambda b: isinstance(b, bool),
            "decoder_depth": lambda v: isinstance(v, int),
        }
        # optional keys (the reference class has none of them): cuda_graph defaults to True, score_kernel_size to 7,
        # half_precision to "tf32" (how fp16 / bf16 enc_dtype / dec_dtype are served, see __init__)
        optional = {"cuda_graph": lambda b: isinstance(b, bool),
                    "half_precision": lambda v: v in ("tf32", "native"),
                    "score_kernel_size": lambda k: isinstance(k, int) and k % 2 == 1 and 1 <= k <= 15}
        if config is not None:
            spec.update({k: v for k, v in optional.items() if hasattr(config, k)})
        cls._enforce_config_spec(config, spec)


# ================================================================================================
# Keypoint selectors
# ================================================================================================
class B200_CovAwareSelector_NoDepth(IKeypointSelector):
    """Bit-exact replacement of CovAwareSelector_NoDepth.select_point (KeypointSelector.py:362-407).
    One host round trip (the candidate count, needed for the CPU `torch.randperm`) instead of two."""

    def __init__(self, config: SimpleNamespace):
        super().__init__(config)
        self.device = _require_cuda(config.device, "B200_CovAwareSelector_NoDepth")
        self._score: ops.ScoreBuffers | None = None
        self._cand: ops.CandidateList | None = None

    @torch.inference_mode()
    def enqueue_candidates(self, match_est) -> "ops.CandidateList":
        """Kernels only (median threshold, flags, ordered compaction) — no host synchronisation."""
        if match_est is None or match_est.cov is None:
            raise ValueError("B200_CovAwareSelector_NoDepth needs match_est.cov (the reference falls back to a grid "
                             "selector here; compose it with GridSelector in the YAML if that is wanted)")
        cov = match_est.cov
        H, W = cov.shape[-2:]
        token = getattr(match_est, "_b200_score", None)
        score = None
        if token is not None:                              # scores fused into the frontend's post-processing pass
            sc, gen, ptr = token
            if (sc.generation == gen and ptr == cov.data_ptr()
                    and sc.ksize == self.config.kernel_size and (sc.h, sc.w) == (H, W)):
                score = sc
        if score is None:                                  # foreign frontend / modified map: score it ourselves
            if self._score is None or (self._score.h, self._score.w) != (H, W):
                self._score = ops.ScoreBuffers(H, W, self.device, self.config.kernel_size)
            score = self._score
            ops.score_only(cov.to(self.device), score)
        if self._cand is None or self._cand.idx.numel() != H * W:
            self._cand = ops.CandidateList(H, W, self.device)
        ops.select_candidates(score, self.config.mask_width, self.config.max_match_cov, match_est.mask, self._cand)
        return self._cand

    @torch.inference_mode()
    def select_point(self, frame, numPoint: int, depth0_est, depth1_est, match_est) -> torch.Tensor:
        return ops.sample_candidates(self.enqueue_candidates(match_est), numPoint)

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        cls._enforce_config_spec(config, {
            "device": lambda dev: isinstance(dev, str) and "cuda" in dev,
            "mask_width": lambda m: isinstance(m, int) and m >= 0,
            "kernel_size": lambda k: isinstance(k, int) and k > 0 and (k % 2 == 1) and k <= 15,
            "max_match_cov": lambda c: isinstance(c, (int, float)) and c > 0.,
        })


class B200_CovAwareSelector(IKeypointSelector):
    """Bit-exact replacement of CovAwareSelector.select_point (KeypointSelector.py:260-334): the depth-aware
    selector of Paper_Reproduce.yaml (quality = (depth_cov0 + depth_cov1) * flow quality, depth and depth-cov gates)."""

    def __init__(self, config: SimpleNamespace):
        super().__init__(config)
        self.device = _require_cuda(config.device, "B200_CovAwareSelector")
        self._score: ops.ScoreBuffers | None = None
        self._cand: ops.CandidateList | None = None

    @torch.inference_mode()
    def select_point(self, frame, numPoint: int, depth0_est, depth1_est, match_est) -> torch.Tensor:
        assert depth0_est.cov is not None
        assert depth1_est.cov is not None
        if match_est is None or match_est.cov is None:
            raise ValueError("B200_CovAwareSelector needs match_est.cov")
        if self.config.max_depth == "auto":
            self.config.max_depth = frame.fx * frame.frame_baseline
        H, W = match_est.cov.shape[-2:]
        if self._score is None or (self._score.h, self._score.w) != (H, W):
            self._score = ops.ScoreBuffers(H, W, self.device, self.config.kernel_size)
            self._cand = ops.CandidateList(H, W, self.device)
        dev = self.device
        ops.score_depth_aware(match_est.cov.to(dev), depth0_est.cov.to(dev), depth1_est.cov.to(dev), self._score)
        ops.select_candidates_depth(self._score, depth0_est.depth.to(dev), depth1_est.depth.to(dev), depth0_est.cov.to(dev),
                                    self.config.mask_width, self.config.max_depth, self.config.max_depth_cov,
                                    self.config.max_match_cov, depth0_est.mask, match_est.mask, self._cand)
        return ops.sample_candidates(self._cand, numPoint)

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        cls._enforce_config_spec(config, {
            "device": lambda dev: isinstance(dev, str) and "cuda" in dev,
            "mask_width": lambda m: isinstance(m, int) and m >= 0,
            "max_depth": lambda dist: (dist == "auto") or (isinstance(dist, (int, float)) and dist > 0.),
            "kernel_size": lambda k: isinstance(k, int) and k > 0 and (k % 2 == 1) and k <= 15,
            "max_depth_cov": lambda c: isinstance(c, (int, float)) and c > 0.,
            "max_match_cov": lambda c: isinstance(c, (int, float)) and c > 0.,
        })


class B200_MappingPointSelector(IKeypointSelector):
    """Bit-exact replacement of MappingPointSelector.select_point (KeypointSelector.py:87-100)."""

    def __init__(self, config: SimpleNamespace):
        super().__init__(config)
        self._cand: ops.CandidateList | None = None

    @torch.inference_mode()
    def enqueue_candidates(self, depth0_est) -> "ops.CandidateList":
        assert depth0_est.cov is not None
        H, W = depth0_est.depth.shape[-2:]
        if self._cand is None or self._cand.idx.numel() != H * W or self._cand.idx.device != depth0_est.depth.device:
            self._cand = ops.CandidateList(H, W, depth0_est.depth.device)
        ops.select_mapping_candidates(depth0_est.depth, depth0_est.cov, self.config.mask_width, self.config.max_depth,
                                      self.config.max_depth_cov, self._cand)
        return self._cand

    @torch.inference_mode()
    def select_point(self, frame, numPoint: int, depth0_est, depth1_est, match_est) -> torch.Tensor:
        return ops.sample_candidates(self.enqueue_candidates(depth0_est), numPoint)

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        cls._enforce_config_spec(config, {
            "max_depth": lambda v: isinstance(v, float),
            "max_depth_cov": lambda v: isinstance(v, float),
            "mask_width": lambda v: isinstance(v, int),
        })


# ================================================================================================
# Covariance model
# ================================================================================================
class B200_MatchCovariance(ICovariance2to3):
    """Replacement of MatchCovariance.estimate (Project2to3.py:124-182). Returns a CPU float64 (K,3,3)
    tensor like the reference (its `create_3x3_matrix` assembles the result on the CPU, and
    Odometry/MACVO.py multiplies it with CPU rotations) — filled by ONE device->host copy instead of nine."""

    def __init__(self, config: SimpleNamespace):
        super().__init__(config)
        self.device = _require_cuda(config.device, "B200_MatchCovariance")
        self.last_status: torch.Tensor | None = None
        self._status_host: torch.Tensor | None = None

    def estimate_device(self, frame, kp, depth_est, depth_cov, flow_cov, want_point: bool = False):
        """Device-resident variant: (cov (K,3,3) fp64 CUDA, point (K,3) fp32 CUDA or None).

        flow_cov may be ANY (K,3) view — Odometry/MACVO.py:231-232 passes `retrieve_pixels(kp, match.cov).T`, a
        transposed view of a (3,K) tensor — the kernel addresses it through its strides, so the reference's in-place
        clamp (Project2to3.py:130-133) lands in the caller's storage (that tensor later becomes `pixel2_uv_cov`).
        A tensor on another device / of another dtype is staged and copied back after the clamp."""
        fc = flow_cov
        staged = None
        if fc is not None and not (fc.is_cuda and fc.device == self.device and fc.dtype == torch.float32):
            staged = fc.to(device=self.device, dtype=torch.float32).contiguous()
            fc = staged
        cov, pt, status = ops.match_covariance(
            kp.to(self.device), depth_est.depth, fc, frame.fx, frame.fy, frame.cx, frame.cy,
            kernel_size=self.config.kernel_size, min_flow_cov=self.config.min_flow_cov,
            min_depth_cov=self.config.min_depth_cov, match_cov_default=self.config.match_cov_default,
            want_point=want_point, depth_cov=depth_cov.to(self.device) if (fc is None and depth_cov is not None) else None)
        if staged is not None:
            flow_cov.copy_(staged)
        self.last_status = status
        return cov, pt

    @torch.inference_mode()
    def estimate(self, frame, kp, depth_est, depth_cov, flow_cov) -> torch.Tensor:
        cov, _ = self.estimate_device(frame, kp, depth_est, depth_cov, flow_cov)
        if self._status_host is None:
            self._status_host = torch.zeros((1,), dtype=self.last_status.dtype).pin_memory()
        self._status_host.copy_(self.last_status, non_blocking=True)      # rides in front of the blocking copy below
        out = cov.cpu()                                                     # ONE synchronising device->host copy
        if int(self._status_host[0]) != 0:
            raise IndexError("MatchCovariance: a keypoint's depth patch leaves the image")
        return out

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        cls._enforce_config_spec(config, {
            "device": lambda dev: isinstance(dev, str) and "cuda" in dev,
            "kernel_size": lambda k: isinstance(k, int) and k % 2 == 1 and 1 <= k <= 31,
            "match_cov_default": lambda c: isinstance(c, (int, float)) and c > 0,
            "min_depth_cov": lambda c: isinstance(c, (int, float)) and c > 0,
            "min_flow_cov": lambda c: isinstance(c, (int, float)) and c > 0,
        })


# ================================================================================================
# Observation filter
# ================================================================================================
class B200_CovarianceSanityFilter(IObservationFilter):
    """Replacement of CovarianceSanityFilter (Module/OutlierFilter.py:91-100): drops observations whose 3x3 covariances
    hold a NaN / Inf. Device-resident covariances (a caller that keeps MatchObs on the GPU, e.g.
    `pipeline.FusedTwoFrameOdometry`, where the same test runs inside `observe_kernel`) go through
    `macvo_cov_sanity_filter`; the CPU float64 tensors `Odometry/MACVO.py:246-270` builds are tested where they live —
    uploading 2 x K x 72 bytes to test them would cost more than the test."""

    @property
    def required_keys(self) -> set:
        return {"obs1_covTc", "obs2_covTc"}

    @torch.inference_mode()
    def filter(self, values, device: torch.device) -> torch.Tensor:
        c1, c2 = values.data["obs1_covTc"], values.data["obs2_covTc"]
        c1 = c1.tensor if hasattr(c1, "tensor") and not isinstance(c1, torch.Tensor) else c1
        c2 = c2.tensor if hasattr(c2, "tensor") and not isinstance(c2, torch.Tensor) else c2
        if c1.is_cuda and c2.is_cuda and c1.dtype == torch.float64 and c2.dtype == torch.float64:
            return ops.cov_sanity_filter(c1, c2).to(device)
        bad = c1.isnan().any(dim=(-1, -2)) | c1.isinf().any(dim=(-1, -2)) | c2.isnan().any(dim=(-1, -2)) | c2.isinf().any(dim=(-1, -2))
        return (~bad).to(device)

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        return


# ================================================================================================
# Map post-processing
# ================================================================================================
class B200_MotionInterpolate(IMapProcessor):
    """Replacement of MotionInterpolate.elaborate_map (Module/MapProcessor.py:52-79), run once at `MACVO.terminate()`:
    relative motions, se3-linear interpolation of the `need_interp` ones, re-integration with quaternion renormalisation —
    one kernel launch (csrc/motion_interp.cu) instead of a Python loop of F pypose compositions. config: {device}."""

    def __init__(self, config: SimpleNamespace | None):
        super().__init__(config)
        self.device = _require_cuda(getattr(config, "device", "cuda"), "B200_MotionInterpolate")

    @torch.inference_mode()
    def elaborate_map(self, frames):
        pose_store, flag_store = frames.data["pose"], frames.data["need_interp"]
        poses = pose_store.tensor if hasattr(pose_store, "tensor") and not isinstance(pose_store, torch.Tensor) else pose_store
        flags = flag_store.tensor if hasattr(flag_store, "tensor") and not isinstance(flag_store, torch.Tensor) else flag_store
        n = poses.shape[0]
        bad = flags[1:].bool().clone()
        bad[:2] = False
        bad[-2:] = False
        interp_idx = torch.nonzero(bad).flatten()
        if n >= 2:
            dev_poses = poses.to(device=self.device, dtype=torch.float32).contiguous()
            ops.motion_interpolate_(dev_poses, flags.to(self.device))
            frames.data["pose"][1:] = dev_poses[1:].to(poses.device)
        return frames, interp_idx

    @classmethod
    def is_valid_config(cls, config: SimpleNamespace | None) -> None:
        cls._enforce_config_spec(config, {"device": lambda d: isinstance(d, str) and "cuda" in d})


# ================================================================================================
# Two-frame pose-graph optimisation
# ================================================================================================
@dataclass
class PGOInput:
    """What Analytic_ReprojDisp_TwoFramePGO reads out of GraphInput (Graphs.py:76-134), as plain tensors."""
    pos_Tw: torch.Tensor        # (K,3)  NED world points
    kp2_uv: torch.Tensor        # (K,2)
    kp2_disp: torch.Tensor      # (K,) or (K,1)
    uv_cov: torch.Tensor        # (K,3)  sigma_uu, sigma_vv, sigma_uv
    disp_cov: torch.Tensor      # (K,) or (K,1)
    K: torch.Tensor             # (3,3)
    baseline: float
    init_pose: torch.Tensor     # (7,) [t, q_xyzw]
    frame_idx: torch.Tensor | None = None
    from_idx: torch.Tensor | None = None
    # graph type "icp" (Graphs.py:33-73) additionally reads:
    kp2_d: torch.Tensor | None = None       # (K,) or (K,1) pixel2_d
    obs_cov: torch.Tensor | None = None     # (K,3,3) float64 obs2_covTc
    pts_cov: torch.Tensor | None = None     # (K,3,3) float64 cov_Tw


@dataclass
class PGOOutput:
    motion: torch.Tensor        # (1,7) float64 on the device (synchronise by reading it)
    frame_idx: torch.Tensor | None
    from_idx: torch.Tensor | None
    stats: torch.Tensor | None = None


def solve_two_frame_pgo(inp: PGOInput, device, cluster: int = 0, graph_type: str = "disp") -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 map values -> float64 (the reference's `.to(dtype=torch.double)`, Optimizer.py:85) -> one launch."""
    f64 = lambda t: t.detach().to(device=device, dtype=torch.float64, non_blocking=True).contiguous()
    K = inp.K.detach().double().cpu()
    intr = (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]),
            float(torch.as_tensor(inp.baseline, dtype=torch.float32).double().reshape(-1)[0]))
    if graph_type == "disp":
        return ops.pgo_solve(f64(inp.pos_Tw), f64(inp.kp2_uv), f64(inp.kp2_disp).reshape(-1), f64(inp.uv_cov),
                             f64(inp.disp_cov).reshape(-1), intr, f64(inp.init_pose).reshape(7), cluster=cluster)
    if graph_type == "reproj":
        return ops.pgo_solve_graph("reproj", f64(inp.pos_Tw), intr, f64(inp.init_pose).reshape(7), kp2_uv=f64(inp.kp2_uv),
                                   uv_cov=f64(inp.uv_cov), cluster=cluster)
    # icp: points_Tc = pixel2point_NED(pixel2_uv, pixel2_d, K) is a registered fp32 buffer of the reference graph (Graphs.py:49-51)
    uv = inp.kp2_uv.detach().to(device=device, dtype=torch.float32)
    d = inp.kp2_d.detach().to(device=device, dtype=torch.float32).reshape(-1)
    Kf = inp.K.detach().to(device=device, dtype=torch.float32)
    pc = torch.stack([d, (uv[:, 0] - Kf[0, 2]) / Kf[0, 0] * d, (uv[:, 1] - Kf[1, 2]) / Kf[1, 1] * d], dim=-1)
    return ops.pgo_solve_graph("icp", f64(inp.pos_Tw), intr, f64(inp.init_pose).reshape(7), pc_obs=pc.double().contiguous(),
                               obs_cov=f64(inp.obs_cov), pts_cov=f64(inp.pts_cov), cluster=cluster)


class B200_TwoFrame_PGO(_PGOBase):
    """Replacement of TwoFrame_PGO (graph types "disp" / "reproj" / "icp", analytic Jacobians): the whole Levenberg-Marquardt loop is
    one persistent kernel launch; `start_optimize` returns immediately and `write_map` synchronises on the
    result, which preserves the frontend / optimiser overlap MAC-VO gets from its spawned CPU process."""

    @staticmethod
    def init_context(config) -> dict:
        gt = getattr(config, "graph_type", "disp")
        if gt not in ("disp", "reproj", "icp") or getattr(config, "autodiff", False):
            raise ValueError("B200_TwoFrame_PGO implements graph_type 'disp' / 'reproj' / 'icp' with the analytic Jacobians "
                             "(autodiff: false)")
        return {"device": _require_cuda(config.device, "B200_TwoFrame_PGO"), "cluster": int(getattr(config, "cluster", 0)),
                "graph_type": gt}

    @staticmethod
    def _optimize(context: dict, graph_data):
        if isinstance(graph_data, PGOInput):
            inp = graph_data
        else:   # MAC-VO's GraphInput (Graphs.py:11-22)
            obs, pts = graph_data.observations.data, graph_data.points.data
            inp = PGOInput(pos_Tw=pts["pos_Tw"], kp2_uv=obs["pixel2_uv"], kp2_disp=obs["pixel2_disp"],
                           uv_cov=obs["pixel2_uv_cov"], disp_cov=obs["pixel2_disp_cov"], K=graph_data.images_intrinsic,
                           baseline=graph_data.baseline, init_pose=torch.as_tensor(graph_data.init_motion).reshape(-1)[:7],
                           frame_idx=graph_data.frame_idx, from_idx=graph_data.from_idx)
            if context.get("graph_type") == "icp":
                inp.kp2_d, inp.obs_cov, inp.pts_cov = obs["pixel2_d"], obs["obs2_covTc"], pts["cov_Tw"]
        pose, stats = solve_two_frame_pgo(inp, context["device"], context["cluster"], context.get("graph_type", "disp"))
        if _REF and not isinstance(graph_data, PGOInput):
            return context, _RefGraphOutput(motion=pose.reshape(1, 7), frame_idx=inp.frame_idx, from_idx=inp.from_idx)
        return context, PGOOutput(motion=pose.reshape(1, 7), frame_idx=inp.frame_idx, from_idx=inp.from_idx, stats=stats)

    if not _REF:
        @classmethod
        def is_valid_config(cls, config: SimpleNamespace | None) -> None:
            cls._enforce_config_spec(config, {
                "graph_type": lambda s: s in {"icp", "reproj", "disp"},
                "device": lambda v: isinstance(v, str) and "cuda" in v,
                "vectorize": lambda b: isinstance(b, bool),
                "parallel": lambda b: b is False,
                "autodiff": lambda b: isinstance(b, bool),
            })


PLUGINS = {
    "frontend": B200_FlowFormerCovFrontend,
    "keypoint": B200_CovAwareSelector_NoDepth,
    "keypoint_depth": B200_CovAwareSelector,
    "mappoint": B200_MappingPointSelector,
    "cov": B200_MatchCovariance,
    "outlier": B200_CovarianceSanityFilter,
    "postprocess": B200_MotionInterpolate,
    "optimizer": B200_TwoFrame_PGO,
}
