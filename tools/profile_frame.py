"""Where does a frame go? torch.profiler over a few eager estimate_pair calls + pipeline stages (CUDA events)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from macvo_b200 import build, plugins as P, synthetic, ops
from macvo_b200.pipeline import TwoFrameOdometry
build.build(verbose=False)
dev = "cuda"
cfg = dict(enc_dtype=sys.argv[1] if len(sys.argv) > 1 else "fp32", dec_dtype=sys.argv[2] if len(sys.argv) > 2 else "fp32")
frames = synthetic.make_sequence(4, 480, 640, pin=True)
fe = P.B200_FlowFormerCovFrontend(NS(weight="synthetic:0", device=dev, enc_dtype=cfg["enc_dtype"], dec_dtype=cfg["dec_dtype"],
                                     decoder_depth=12, enforce_positive_disparity=False, cuda_graph=False))
for _ in range(3):
    fe.estimate_pair(frames[0], frames[1])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fe.estimate_pair(frames[1], frames[2])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
# section timing of the network
net = fe.net
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
A = torch.cat([frames[2].imageL, frames[1].imageL]).to(dev); B = torch.cat([frames[2].imageR, frames[2].imageL]).to(dev)
with torch.inference_mode():
    for rep in range(2):
        t = [ev()]
        i1 = ((2 * A) - 1.0).to(net.enc_dtype); i2 = ((2 * B) - 1.0).to(net.enc_dtype)
        ctx = net.svt(i1, "context_encoder"); t.append(ev())
        feats = net.svt(torch.cat([i1, i2]), "memory_encoder.feat_encoder"); feats = net._conv(feats, "memory_encoder.channel_convertor"); t.append(ev())
        cv = net.corr_fn(feats[:2], feats[2:]).to(feats.dtype); t.append(ev())
        cost_maps = cv.view(2 * 4800, 1, 60, 80)
        tok = net.patch_embed(cost_maps); t.append(ev())
        cm, cmaps = net.cost_perceiver(cv, ctx); t.append(ev())
        out = net.memory_decoder(cm, ctx.float(), cmaps.float()); t.append(ev())
        torch.cuda.synchronize()
    names = ["context svt", "feat svt+conv", "corr", "patch_embed(alone)", "cost_perceiver(total incl patch_embed)", "decoder x12"]
    for n, a, b in zip(names, t[:-1], t[1:]):
        print(f"{n:45s} {a.elapsed_time(b):8.3f} ms")
    def table(title, fn):
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            fn()
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in pr.key_averages()]
        rows.sort(key=lambda r: -r[1])
        print(f"--- {title} kernels: total {sum(r[1] for r in rows) / 1e3:.2f} ms, {sum(r[2] for r in rows)} launches")
        for k, us, n in rows[:32]:
            print(f"{us / 1e3:8.3f} ms {n:5d}x  {k[:110]}")
    table("cost_perceiver", lambda: net.cost_perceiver(cv, ctx))
    table("svt x2 (context + features)", lambda: (net.svt(i1, "context_encoder"), net.svt(torch.cat([i1, i2]), "memory_encoder.feat_encoder")))
    # decoder alone under the profiler: kernel-level table
    with profile(activities=[ProfilerActivity.CUDA]) as prof2:
        net.memory_decoder(cm, ctx.float(), cmaps.float())
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof2.key_averages()]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"--- decoder kernels: total {tot / 1e3:.2f} ms, {sum(r[2] for r in rows)} launches")
    for k, us, n in rows[:40]:
        print(f"{us / 1e3:8.3f} ms {n:5d}x  {k[:110]}")
