"""Kernel-level timings on the GPU box (CUDA events, warm-up, L2 flush between iterations).
Writes gpurun_out/kernels.json. Not part of the bench contract; used to steer optimisation."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import build, ops  # noqa: E402
from tests.golden import cases  # noqa: E402

build.build(verbose=False)
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return {"median_us": ts[len(ts) // 2], "min_us": ts[0]}


res = {}
peak = 6566.7
for (B, H1, W1) in [(2, 60, 80), (2, 80, 80), (2, 90, 160)]:
    N = H1 * W1
    f1, f2 = cases.corr_inputs(B, H1, W1)
    d1, d2 = f1.to(dev), f2.to(dev)
    algo_bytes = B * (4 * N * N + 8 * N * 256)
    for mode, name in ((ops.CORR_SIMT, "simt"), (ops.CORR_TC_3XF16, "tc3"), (ops.CORR_TC_1XF16, "tc1")):
        if mode == ops.CORR_SIMT and N > 6400:
            continue
        r = timeit(lambda: ops.corr_build(d1, d2, mode=mode), iters=10)
        r["GBps"] = algo_bytes / r["median_us"] / 1e3
        r["frac_of_hbm"] = r["GBps"] / peak
        res[f"corr_{name}_B{B}_N{N}"] = r
        print(f"corr_{name}_B{B}_N{N}", r, flush=True)
    del d1, d2
B, H1, W1 = 2, 60, 80
cm, co = cases.lookup_inputs(B, H1, W1)
dcm, dco = cm.to(dev), co.to(dev)
res["lookup_B2_60x80"] = timeit(lambda: ops.corr_lookup(dcm, dco))
print("lookup", res["lookup_B2_60x80"], flush=True)
flow, cov = cases.selector_inputs(480, 640, "plain")
dfl, dcv = flow.to(dev), cov.to(dev)
score = ops.ScoreBuffers(480, 640, dev, 7)
cand = ops.CandidateList(480, 640, dev)
res["dense_score_480x640"] = timeit(lambda: ops.dense_postproc(dfl, dcv, 80.0, False, score=score))
def sel():
    ops.dense_postproc(dfl, dcv, 80.0, False, score=score)
    ops.select_candidates(score, 32, 100.0, None, cand)
    return ops.sample_candidates(cand, 512)
res["select_total_480x640"] = timeit(sel)
print("dense", res["dense_score_480x640"], "select", res["select_total_480x640"], flush=True)
for K in (200, 512, 2048, 4096):
    c = cases.pgo_inputs(K, 6)
    f64 = lambda t: t.double().to(dev)
    args = (f64(c["pos_Tw"]), f64(c["kp2_uv"]), f64(c["kp2_disp"]), f64(c["uv_cov"]), f64(c["disp_cov"]),
            (320.0, 320.0, 320.0, 240.0, 0.25), f64(c["init_pose"]))
    for cl in (1, 8):
        res[f"pgo_K{K}_cluster{cl}"] = timeit(lambda: ops.pgo_solve(*args, cluster=cl))
        print(f"pgo_K{K}_cluster{cl}", res[f"pgo_K{K}_cluster{cl}"], flush=True)
    kp, depth, fc = cases.cov_inputs(480, 640, K, "float_cov")
    dkp, dd, dfc = kp.to(dev), depth.to(dev), fc.to(dev)
    res[f"cov_K{K}"] = timeit(lambda: ops.match_covariance(dkp, dd, dfc, 320., 320., 320., 240.))
    print(f"cov_K{K}", res[f"cov_K{K}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/kernels.json", "w"), indent=1)
