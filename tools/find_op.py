"""which aten op (with input shapes) launches a given kernel inside the decoder? usage: find_op.py <kernel substring>"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from macvo_b200 import plugins as P, synthetic
from torch.profiler import profile, ProfilerActivity
pat = sys.argv[1:] or ["gemv"]
dev = "cuda"
frames = synthetic.make_sequence(3, 480, 640, pin=True)
fe = P.B200_FlowFormerCovFrontend(NS(weight="synthetic:0", device=dev, enc_dtype="fp32", dec_dtype="fp32", decoder_depth=2,
                                     enforce_positive_disparity=False, cuda_graph=False))
for _ in range(2):
    fe.estimate_pair(frames[0], frames[1])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    fe.estimate_pair(frames[1], frames[2])
    torch.cuda.synchronize()
seen = {}
for e in prof.events():
    for k in getattr(e, "kernels", []):
        if any(p in k.name for p in pat):
            key = (e.name, str(e.input_shapes), k.name[:60])
            d = seen.setdefault(key, [0, 0.0])
            d[0] += 1; d[1] += k.duration
for (n, s, k), (c, us) in sorted(seen.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:8.3f} ms {c:4d}x {n:28s} {s[:120]}  <- {k}")
