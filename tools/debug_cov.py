import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace as NS
from macvo_b200 import synthetic, plugins as P
from macvo_b200.flowformer_cov import synthetic_state_dict
from oracle import pipeline_cpu as pc, covariance as ocov, frontend as ofe, keypoint as okp
DEV = "cuda"
H, W = 192, 256
frames = synthetic.make_sequence(3, H, W)
fe = pc.CpuFrontend(synthetic_state_dict(0), decoder_depth=4)
depth0 = fe.estimate_depth(frames[0])
depth1, match = fe.estimate_pair(frames[0], frames[1])
torch.manual_seed(5)
kp0 = okp.cov_aware_select_nodepth(match.cov, 64)
kp1 = kp0 + ofe.retrieve_pixels(kp0, match.flow).T
inb = ofe.filter_points_in_range(kp1, (32, W - 32), (32, H - 32))
kp0, kp1 = kp0[inb], kp1[inb]
fc = ofe.retrieve_pixels(kp0, match.cov).T.contiguous()
ref = ocov.match_covariance(kp1, depth1.depth, fc.clone(), frames[1].fx, frames[1].fy, frames[1].cx, frames[1].cy)
covm = P.B200_MatchCovariance(NS(device=DEV, kernel_size=31, match_cov_default=0.25, min_depth_cov=0.05, min_flow_cov=0.25))
d1 = NS(depth=depth1.depth.to(DEV))
gpu = covm.estimate(frames[1], kp1.to(DEV), d1, None, fc.clone().to(DEV))
rel = ((gpu - ref).abs() / ref.abs().amax(dim=(1, 2), keepdim=True)).amax(dim=(1, 2))
i = int(rel.argmax()); print("worst", i, rel[i].item(), "kp", kp1[i], "fc", fc[i])
print("gpu", gpu[i]); print("ref", ref[i])
# fp64 recomputation
u, v = kp1[i].double(); ul, vl = int(kp1[i, 0]), int(kp1[i, 1])
suu, svv, suv = [max(float(fc[i, 0]), 0.0625), max(float(fc[i, 1]), 0.0625), float(fc[i, 2])]
xs = np.arange(-15, 16, dtype=np.float64)
inv = np.linalg.inv(np.array([[suu, suv], [suv, svv]]))
X, Y = np.meshgrid(xs, xs, indexing="ij")
z = np.exp(-0.5 * (X * X * inv[0, 0] + 2 * X * Y * inv[0, 1] + Y * Y * inv[1, 1])); wgt = z / z.sum()
patch = depth1.depth[0, 0, vl - 15:vl + 16, ul - 15:ul + 16].double().numpy()
wavg = (wgt * patch).sum(); wvar = (wgt * (patch - wavg) ** 2).sum()
print("fp64 wavg", wavg, "wvar", wvar, "patch min/max", patch.min(), patch.max(), "suu svv", suu, svv)
