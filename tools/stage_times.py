"""per-frame split of the hot path: frontend (graph replay + fused post-processing) vs everything after it.
Wall clock with a device synchronisation after each part. usage: python tools/stage_times.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from macvo_b200 import synthetic
dev = "cuda:0"
odo = bench.build_gpu_pipeline(bench.CONFIGS["performant"], dev)
frames = synthetic.make_sequence(8, 480, 640)
import copy
dframes = []
for f in frames:
    fd = copy.copy(f)
    fd.imageL, fd.imageR = f.imageL.to(dev), f.imageR.to(dev)
    dframes.append(fd)
frames = dframes
odo.initialize(frames[0])
for f in frames[1:4]:
    odo.run_pair(f)
torch.cuda.synchronize()
fe = odo.frontend
orig = fe.estimate_pair
acc = {"frontend": 0.0, "total": 0.0}
def timed(a, b):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(a, b)
    torch.cuda.synchronize(); acc["frontend"] += time.perf_counter() - t0
    return r
fe.estimate_pair = timed
n = 0
for rep in range(3):
    for f in frames[4:] + frames[1:4]:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        odo.run_pair(f)
        torch.cuda.synchronize(); acc["total"] += time.perf_counter() - t0
        n += 1
print(f"frames {n}: frontend {acc['frontend'] / n * 1e3:.3f} ms, whole run_pair {acc['total'] / n * 1e3:.3f} ms, "
      f"after-frontend {(acc['total'] - acc['frontend']) / n * 1e3:.3f} ms")
