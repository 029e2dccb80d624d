"""corr op = split pre-pass + main kernel: time the whole op and the main kernel alone (MACVO_B200_CORR_DEBUG=16 skips the
pre-pass and reuses the operands of the previous call). usage: python tools/time_corr_parts.py [debug-bits]"""
import os, sys
os.environ["MACVO_B200_CORR_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "0"
import torch
sys.path.insert(0, ".")
from macvo_b200 import ops
g = torch.Generator().manual_seed(2)
f1 = (torch.randn(2, 256, 60, 80, generator=g) * 0.5).cuda(); f2 = (torch.randn(2, 256, 60, 80, generator=g) * 0.5).cuda()
if len(sys.argv) > 2:
    f1, f2 = f1.contiguous(memory_format=torch.channels_last), f2.contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode, name in ((ops.CORR_TC_3XF16, "tc3"), (ops.CORR_TC_1XF16, "tc1")):
    for _ in range(3): ops.corr_build(f1, f2, mode=mode)
    ts = []
    for _ in range(12):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.corr_build(f1, f2, mode=mode); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print(f"debug={os.environ['MACVO_B200_CORR_DEBUG']} {name}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us")
