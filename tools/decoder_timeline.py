"""In-graph timeline of the tensor-core decoder kernels (globaltimer stamps by block 0 of every conv / GRU launch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
import torch
from macvo_b200 import plugins as P, synthetic, ops

dev = "cuda"
frames = synthetic.make_sequence(3, 480, 640, pin=True)
fe = P.B200_FlowFormerCovFrontend(NS(weight="synthetic:0", device=dev, enc_dtype="fp32", dec_dtype="fp32", decoder_depth=12,
                                     enforce_positive_disparity=False, cuda_graph=False))
net = fe.net
A = torch.cat([frames[2].imageL, frames[1].imageL]).to(dev)
B = torch.cat([frames[2].imageR, frames[2].imageL]).to(dev)
CAP = 1024
tl = torch.zeros(1 + 3 * CAP, dtype=torch.int64, device=dev)
lib = ops.load_library()
with torch.inference_mode():
    i1, i2 = ((2 * A) - 1.0), ((2 * B) - 1.0)
    ctx = net.svt(i1, "context_encoder")
    feats = net._conv(net.svt(torch.cat([i1, i2]), "memory_encoder.feat_encoder"), "memory_encoder.channel_convertor")
    cv = net.corr_fn(feats[:2], feats[2:]).to(feats.dtype)
    cm, cmaps = net.cost_perceiver(cv, ctx)
    ctx, cmaps = ctx.float(), cmaps.float()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            net.memory_decoder(cm, ctx, cmaps)
        torch.cuda.synchronize()
        lib.macvo_tc_set_timeline(tl.data_ptr(), CAP)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            net.memory_decoder(cm, ctx, cmaps)
        lib.macvo_tc_set_timeline(None, 0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    tl.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
print("decoder replay ms:", a.elapsed_time(b))
t = tl.cpu()
n = min(int(t[0]), CAP)
ev = sorted([(int(t[2 + 3 * i]), int(t[3 + 3 * i]), int(t[1 + 3 * i])) for i in range(n)])
t0 = ev[0][0]
print(f"{n} events; columns: start us | duration us | gap since the previous kernel's end | kernel")
prev_end = t0


def name(k):
    if k < 100:
        return f"gru stage {k & 1} pass {k >> 1}"
    k -= 100
    return f"conv taps {k // 1000 % 10} kblocks {k // 10000 % 10} n_cta {k // 100000}"


# third iteration only (steady state)
per_iter = n // 12
for st, en, k in ev[2 * per_iter: 3 * per_iter + 2]:
    print(f"{(st - t0) / 1e3:9.2f} {(en - st) / 1e3:7.2f} {(st - prev_end) / 1e3:7.2f}  {name(k)}")
    prev_end = en
