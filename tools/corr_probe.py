"""Correlation-build timing probe (CUDA events, L2 flushed): modes tc3 / tf32 at B=2, D=256, N=4800 and 14400, channels_last
features. MACVO_B200_CORR_DEBUG ablations (read once per process): 1 = no global stores, 2 = no MMAs, 3 = neither."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macvo_b200 import build, ops  # noqa: E402

build.build(verbose=False)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"dbg": os.environ.get("MACVO_B200_CORR_DEBUG", "0")}
for (H1, W1) in (((60, 80),) if "--small" in sys.argv else ((60, 80), (90, 160))):
    n = H1 * W1
    g = torch.Generator().manual_seed(2)
    f1 = (torch.randn(2, 256, H1, W1, generator=g) * 0.5).to(dev).contiguous(memory_format=torch.channels_last)
    f2 = (torch.randn(2, 256, H1, W1, generator=g) * 0.5).to(dev).contiguous(memory_format=torch.channels_last)
    for mode, name in ((ops.CORR_TC_3XF16, "tc3"), (ops.CORR_TC_TF32, "tf32"), (ops.CORR_TC_1XF16, "tc1")):
        for _ in range(3):
            ops.corr_build(f1, f2, mode=mode)
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            c = ops.corr_build(f1, f2, mode=mode)
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
            del c
        ts.sort()
        algo = 2 * (4 * n * n + 8 * n * 256)
        out[f"{name}_N{n}"] = {"median_us": ts[5], "min_us": ts[0], "GBps": algo / ts[5] / 1e3, "frac": algo / ts[5] / 1e3 / 6566.7}
print(json.dumps(out))
