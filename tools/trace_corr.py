import os, sys, ctypes as C
os.environ["MACVO_B200_CORR_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "8"
import torch
sys.path.insert(0, ".")
from macvo_b200 import ops
from tests.golden import cases
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f1, f2 = cases.corr_inputs(2, 60, 80)
d1, d2 = f1.cuda(), f2.cuda()
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.corr_build(d1, d2, mode=mode)
flush.zero_(); torch.cuda.synchronize()
ops.corr_build(d1, d2, mode=mode); torch.cuda.synchronize()
lib = ops.load_library()
buf = (C.c_ulonglong * (2*4*512*2))()
lib.macvo_corr_debug_trace.restype = C.c_int
n = lib.macvo_corr_debug_trace(buf, len(buf))
ev = []
for cta in range(2):
    for role in range(4):
        for i in range(512):
            o = ((cta*4+role)*512 + i)*2
            if buf[o]: ev.append((buf[o], cta, role, buf[o+1]))
ev.sort(); t0 = ev[0][0]
names = {0: "PROD", 1: "MMA ", 2: "EPI "}
for t, cta, role, tag in ev[:int(sys.argv[3]) if len(sys.argv) > 3 else 260]:
    print(f"{(t-t0)/1000:9.2f} us  cta{cta} {names[role]} {tag}")
print("last event", (ev[-1][0]-t0)/1000, "us; events", len(ev))
