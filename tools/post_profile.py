"""cProfile + torch.profiler of the part of run_pair AFTER the frontend (selector, gathers, covariance, PGO, mapping)"""
import os, sys, time, cProfile, pstats, io, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, copy
from macvo_b200 import synthetic
dev = "cuda:0"
odo = bench.build_gpu_pipeline(bench.CONFIGS["performant"], dev)
frames = []
for f in synthetic.make_sequence(8, 480, 640):
    fd = copy.copy(f); fd.imageL, fd.imageR = f.imageL.to(dev), f.imageR.to(dev); frames.append(fd)
odo.initialize(frames[0])
for f in frames[1:5]:
    odo.run_pair(f)
torch.cuda.synchronize()
fe = odo.frontend
orig = fe.estimate_pair
pr = cProfile.Profile()
def wrapped(a, b):
    pr.disable()
    r = orig(a, b)
    torch.cuda.synchronize()
    pr.enable()
    return r
fe.estimate_pair = wrapped
for f in frames[5:] + frames[1:5]:
    pr.enable(); odo.run_pair(f); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
