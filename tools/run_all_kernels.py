"""Launch every macvo_b200 kernel once or twice at the 640x480 workload sizes (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import build, ops
from tests.golden import cases
build.build(verbose=False)
dev = "cuda"
reps = 2
cm, co = cases.lookup_inputs(2, 60, 80)
dcm, dco = cm.to(dev), co.to(dev)
flow, cov = cases.selector_inputs(480, 640, "plain")
dfl, dcv = flow.to(dev), cov.to(dev)
score = ops.ScoreBuffers(480, 640, dev, 7)
cand, mcand = ops.CandidateList(480, 640, dev), ops.CandidateList(480, 640, dev)
kp, depth, fc = cases.cov_inputs(480, 640, 512, "float_cov")
dkp, dd, dfc = kp.to(dev), depth.to(dev), fc.to(dev)
c = cases.pgo_inputs(512, 6)
f64 = lambda t: t.double().to(dev)
pargs = (f64(c["pos_Tw"]), f64(c["kp2_uv"]), f64(c["kp2_disp"]), f64(c["uv_cov"]), f64(c["disp_cov"]),
         (320.0, 320.0, 320.0, 240.0, 0.25), f64(c["init_pose"]))
f1, f2 = cases.corr_inputs(2, 60, 80)
d1, d2 = f1.to(dev), f2.to(dev)
for _ in range(reps):
    ops.corr_build(d1, d2, mode=ops.CORR_TC_3XF16)
    ops.corr_lookup(dcm, dco)
    d = ops.dense_postproc(dfl, dcv, 80.0, False, score=score)
    ops.select_candidates(score, 32, 100.0, None, cand)
    torch.manual_seed(5)
    kps = ops.sample_candidates(cand, 512)
    ops.select_mapping_candidates(d["depth"], d["depth_cov"], 32, 5.0, 0.005, mcand)
    ops.sample_candidates(mcand, 2000)
    ops.retrieve_pixels(kps, d["depth"])
    ops.match_covariance(dkp, dd, dfc, 320., 320., 320., 240.)
    ops.pgo_solve(*pargs, cluster=2)
    ops.pgo_solve(*pargs, cluster=8)
torch.cuda.synchronize()
print("done")
