"""Per-stage error ladder of the FlowFormerCov frontend at 640x480 / decoder_depth 12 (BASELINE configs[1] shape):
GPU network (strict fp32, and the TF32 / fp16-GMA mode bench.py times) against the float64 ground truth of
tests/golden/net_cfgA.pt, next to the reference's own fp32 distance from that truth (the noise floor).

    python tools/parity_ladder.py [--out gpurun_out/parity_ladder.json]

Error measure per stage: max |x - truth| / mean |truth|  (relative to the stage's scale); covariance: max relative.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from tests.golden import cases  # noqa: E402

STAGES = ("context", "feats", "corr_rows", "cost_memory")


def set_mode(mode: str) -> None:
    tf32 = mode != "strict"
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.set_float32_matmul_precision("medium" if tf32 else "highest")


def ladder(mode: str, golden: dict, device: str = "cuda") -> dict:
    from macvo_b200.flowformer_cov import FlowFormerCovNet, synthetic_state_dict
    set_mode(mode)
    A, B = cases.cfgA_inputs()
    net = FlowFormerCovNet(synthetic_state_dict(0), device, decoder_depth=golden["decoder_depth"])
    net.taps = {}
    flow, cov = net.inference(A.to(device), B.to(device))
    torch.cuda.synchronize()
    T, S = golden["truth"], cases.cfgA_sample
    rel = lambda x, t: ((x.double().cpu() - t).abs().max() / t.abs().mean()).item()
    out = {}
    for name in STAGES:
        out[name] = rel(S(name, net.taps[name][0]), T[name])
    for name in ("flow_iter", "cov_iter"):
        out[name] = [rel(S(name, x), t) for x, t in zip(net.taps[name], T[name])]
    fs, cs = S("flow", flow).double().cpu(), S("cov", cov).double().cpu()
    out["flow"] = rel(fs, T["flow"])
    out["flow_abs_max"] = (fs - T["flow"]).abs().max().item()
    out["cov_rel_max"] = ((cs - T["cov"]).abs() / T["cov"].abs()).max().item()
    out["cov_rel_p99"] = ((cs - T["cov"]).abs() / T["cov"].abs()).flatten().quantile(0.99).item()
    r = golden["ref32"]
    out["flow_vs_ref32"] = ((fs - r["flow"].double()).abs().max() / T["flow"].abs().mean()).item()
    out["cov_rel_vs_ref32"] = ((cs - r["cov"].double()).abs() / r["cov"].double().abs()).max().item()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "parity_ladder.json"))
    a = ap.parse_args()
    g = torch.load(os.path.join(REPO, "tests", "golden", "net_cfgA.pt"), weights_only=False)
    T, r = g["truth"], g["ref32"]
    floor = {"flow": ((r["flow"].double() - T["flow"]).abs().max() / T["flow"].abs().mean()).item(),
             "cov_rel_max": ((r["cov"].double() - T["cov"]).abs() / T["cov"].abs()).max().item()}
    res = {"floor_ref32_vs_truth": floor, "floor_full_tensor": g["floor"]}
    for mode in ("strict", "tf32"):
        res[mode] = ladder(mode, g)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(f"{'stage':<14}{'strict fp32':>14}{'tf32 mode':>14}")
    for name in STAGES:
        print(f"{name:<14}{res['strict'][name]:>14.3e}{res['tf32'][name]:>14.3e}")
    for i in range(len(res["strict"]["flow_iter"])):
        print(f"flow iter {i + 1:<4}{res['strict']['flow_iter'][i]:>14.3e}{res['tf32']['flow_iter'][i]:>14.3e}"
              f"   cov iter {res['strict']['cov_iter'][i]:>10.3e}{res['tf32']['cov_iter'][i]:>12.3e}")
    for name in ("flow", "flow_abs_max", "cov_rel_max", "cov_rel_p99", "flow_vs_ref32", "cov_rel_vs_ref32"):
        print(f"{name:<14}{res['strict'][name]:>14.3e}{res['tf32'][name]:>14.3e}")
    print("reference fp32 vs truth (noise floor):", floor)


if __name__ == "__main__":
    main()
