"""Which copies are left in a frame? torch.profiler with shapes + stacks over one eager estimate_pair: aten::copy_ by shape / site."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from macvo_b200 import plugins as P, synthetic
frames = synthetic.make_sequence(4, 480, 640, pin=True)
fe = P.B200_FlowFormerCovFrontend(NS(weight="synthetic:0", device="cuda", enc_dtype="fp32", dec_dtype="fp32", decoder_depth=12,
                                     enforce_positive_disparity=False, cuda_graph=False))
for _ in range(2):
    fe.estimate_pair(frames[0], frames[1])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    fe.estimate_pair(frames[1], frames[2])
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True, group_by_stack_n=4)
rows = [e for e in ka if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::gelu", "aten::mul", "aten::cat", "aten::fill_", "aten::zero_")]
rows.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"elementwise / copy ops: {tot / 1e3:.3f} ms")
for e in rows[:40]:
    stack = [s for s in e.stack if "flowformer_cov" in s or "plugins" in s or "ops.py" in s]
    print(f"{e.device_time_total / 1e3:8.3f} ms {e.count:4d}x {e.key:12s} {str(e.input_shapes)[:70]:70s} {' <- '.join(x.split('/')[-1][:60] for x in stack[:2])}")
