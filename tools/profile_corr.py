"""Tiny driver for ncu: runs the corr build (tc3, tc1) a few times at 640x480 size with channels_last features."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import build, ops
from tests.golden import cases
build.build(verbose=False)
f1, f2 = cases.corr_inputs(2, 60, 80)
d1 = f1.cuda().contiguous(memory_format=torch.channels_last)
d2 = f2.cuda().contiguous(memory_format=torch.channels_last)
modes = [ops.CORR_TC_3XF16, ops.CORR_TC_1XF16] if len(sys.argv) < 2 else [int(sys.argv[1])]
for mode in modes:
    for _ in range(3):
        out = ops.corr_build(d1, d2, mode=mode)
    torch.cuda.synchronize()
print("done")
