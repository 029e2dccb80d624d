import os, sys, subprocess, json
code = r'''
import os, sys, torch
sys.path.insert(0, ".")
from macvo_b200 import ops
from tests.golden import cases
f1, f2 = cases.corr_inputs(2, 60, 80)
d1, d2 = f1.cuda(), f2.cuda()
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
for mode, name in ((1, "tc3"), (2, "tc1")):
    for _ in range(3): ops.corr_build(d1, d2, mode=mode)
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.corr_build(d1, d2, mode=mode); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); print(os.environ.get("MACVO_B200_CORR_DEBUG", "0"), name, round(ts[5], 1), "us")
'''
for dbg in ("0", "1", "2", "3", "5", "7"):
    env = dict(os.environ, MACVO_B200_CORR_DEBUG=dbg)
    print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
