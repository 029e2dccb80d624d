"""BASELINE configs[3] shape on ONE GPU: 1280x720 stereo, 4096 keypoints, whole hot path (FusedTwoFrameOdometry).
Prints ms / frame, peak memory and — with --check — the frontend's parity against the CPU oracle network at this size
(N = 14 400 tokens, 1.66 GB correlation volume per estimate_pair; the CPU run takes a few minutes).

    python tools/run_720p.py [--frames 6] [--check] [--strict]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--strict", action="store_true", help="allow_tf32 = False for the parity check")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "run_720p.json"))
    a = ap.parse_args()
    from types import SimpleNamespace as NS
    from macvo_b200 import plugins, synthetic
    from macvo_b200.pipeline import FusedTwoFrameOdometry
    H, W, KP = 720, 1280, 4096
    dev = "cuda:0"
    frames = synthetic.make_sequence(4, H, W, pin=True)
    fe = plugins.B200_FlowFormerCovFrontend(NS(weight="synthetic:0", device=dev, enc_dtype="fp32", dec_dtype="fp32",
                                               decoder_depth=12, enforce_positive_disparity=False, cuda_graph=True))
    sel = plugins.B200_CovAwareSelector_NoDepth(NS(device=dev, kernel_size=3,   # 3x3 NMS: > 4096 candidates with the stand-in network
                                                    mask_width=32, max_match_cov=100.0))
    msel = plugins.B200_MappingPointSelector(NS(max_depth=5.0, max_depth_cov=0.005, mask_width=32))
    cov = plugins.B200_MatchCovariance(NS(device=dev, kernel_size=31, match_cov_default=0.25, min_depth_cov=0.05, min_flow_cov=0.25))
    pgo = plugins.B200_TwoFrame_PGO(NS(graph_type="disp", device=dev, vectorize=True, parallel=False, autodiff=False))
    odo = FusedTwoFrameOdometry(fe, sel, cov, pgo, num_point=KP, map_selector=msel)
    torch.manual_seed(5)
    odo.initialize(frames[0])
    seq = [frames[1 + (i % 3)] for i in range(a.frames + 3)]
    for f in seq[:3]:
        odo.run_pair(f)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for f in seq[3:]:
        odo.run_pair(f)
        odo.latest_pose()
    e.record()
    torch.cuda.synchronize()
    obs = odo.observations()
    res = {"shape": [H, W], "keypoints": KP, "ms_per_frame": s.elapsed_time(e) / a.frames, "frames": a.frames,
           "num_obs": obs["num_obs"], "num_kp_inbound": obs["num_kp"], "status": obs["status"],
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           "pose_finite": bool(torch.isfinite(odo.latest_pose()).all())}
    print(json.dumps(res))
    if a.check:
        from macvo_b200.flowformer_cov import FlowFormerCovNet, synthetic_state_dict
        from oracle import frontend as ofe
        if a.strict:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            torch.set_float32_matmul_precision("highest")
        A = torch.cat([frames[2].imageL, frames[1].imageL])
        B = torch.cat([frames[2].imageR, frames[2].imageL])
        net = FlowFormerCovNet(synthetic_state_dict(0), dev)
        gf, gc = net.inference(A.to(dev), B.to(dev))
        torch.cuda.synchronize()
        torch.set_float32_matmul_precision("highest")
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        t0 = time.time()
        cpu = FlowFormerCovNet(synthetic_state_dict(0), "cpu", corr_fn=ofe.corr_volume, lookup_fn=ofe.window_lookup)
        cf, cc = cpu.inference(A, B)
        res["cpu_seconds"] = time.time() - t0
        res["mode"] = "strict fp32" if a.strict else "tf32"
        res["flow_rel_vs_cpu_fp32"] = ((gf.cpu() - cf).abs().max() / cf.abs().mean()).item()
        res["cov_rel_vs_cpu_fp32"] = ((gc.cpu() - cc).abs() / cc.abs()).max().item()
        print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
