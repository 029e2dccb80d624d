"""decoder_token_kernel standalone at the 640x480 shape (P = 9600): CUDA-event timing (L2 flushed / warm) and a target for ncu."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import build, ops
from macvo_b200.flowformer_cov import synthetic_state_dict
build.build(verbose=False)
dev = "cuda"
sd = {k: v.to(dev) for k, v in synthetic_state_dict(0).items() if k.startswith("memory_decoder.")}
b, h, w = 2, 60, 80
P = b * h * w
g = torch.Generator().manual_seed(1)
cf = torch.randn(P, 81, generator=g).to(dev)
coords = (torch.rand(b, 2, h, w, generator=g) * 80).to(dev)
key, value = torch.randn(P, 8, 64, generator=g).to(dev), torch.randn(P, 8, 64, generator=g).to(dev)
blob = ops.decoder_token_blob(sd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.decoder_token(cf, coords, key, value, blob)
res = {}
for name, do_flush in (("cold_l2", True), ("warm_l2", False)):
    ts = []
    for _ in range(10):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.decoder_token(cf, coords, key, value, blob); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); res[name + "_us"] = ts[5]
print(json.dumps(res))
