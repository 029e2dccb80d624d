import torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = True; torch.set_float32_matmul_precision("medium")
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
def explicit(q, k, v):
    a = (torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)).softmax(-1)
    return torch.matmul(a, v)
shapes = {
 "perceiver input_layer": ((9600, 8, 8, 16), (9600, 8, 80, 16)),
 "perceiver vert local": ((1728, 8, 49, 16), (1728, 8, 49, 16)),
 "perceiver vert global": ((16, 8, 4800, 16), (16, 8, 300, 16)),
 "svt s0 local B4": ((1656, 4, 49, 32), (1656, 4, 49, 32)),
 "svt s0 global B4": ((4, 4, 19200, 32), (4, 4, 300, 32)),
 "svt s1 local B4": ((432, 8, 49, 32), (432, 8, 49, 32)),
 "svt s1 global B4": ((4, 8, 4800, 32), (4, 8, 300, 32)),
}
for name, (qs, ks) in shapes.items():
    q = torch.randn(qs, device=dev); k = torch.randn(ks, device=dev); v = torch.randn(ks, device=dev)
    r = {"default": t(lambda: F.scaled_dot_product_attention(q, k, v)), "explicit": t(lambda: explicit(q, k, v))}
    try:
        with sdpa_kernel(SDPBackend.MATH):
            r["math"] = t(lambda: F.scaled_dot_product_attention(q, k, v))
    except Exception as ex: r["math"] = str(ex)[:30]
    qb, kb, vb = q.bfloat16(), k.bfloat16(), v.bfloat16()
    r["bf16 sdpa(+casts)"] = t(lambda: F.scaled_dot_product_attention(q.bfloat16(), k.bfloat16(), v.bfloat16()).float())
    r["bf16 sdpa(no casts)"] = t(lambda: F.scaled_dot_product_attention(qb, kb, vb))
    print(f"{name:26s}", {k_: (round(v_, 1) if isinstance(v_, float) else v_) for k_, v_ in r.items()}, flush=True)
# layer norm on the big token tensor and elementwise
x = torch.randn(9600, 80, 128, device=dev); w = torch.ones(128, device=dev); b = torch.zeros(128, device=dev)
print("LN 9600x80x128", round(t(lambda: F.layer_norm(x, (128,), w, b)), 1), "us; relu", round(t(lambda: F.relu(x)), 1), "; add", round(t(lambda: x + x), 1))
lin = torch.randn(128, 128, device=dev)
print("linear 768000x128x128", round(t(lambda: F.linear(x, lin, b)), 1))
cm = torch.randn(9600, 1, 64, 80, device=dev)
w1 = torch.randn(16, 1, 6, 6, device=dev).contiguous(memory_format=torch.channels_last); b1 = torch.zeros(16, device=dev)
w2 = torch.randn(32, 16, 6, 6, device=dev).contiguous(memory_format=torch.channels_last); b2 = torch.zeros(32, device=dev)
w3 = torch.randn(64, 32, 6, 6, device=dev).contiguous(memory_format=torch.channels_last); b3 = torch.zeros(64, device=dev)
torch.backends.cudnn.allow_tf32 = True
y1 = F.conv2d(cm, w1, b1, stride=2, padding=2); y2 = F.conv2d(F.relu(y1), w2, b2, stride=2, padding=2)
print("conv1", round(t(lambda: F.conv2d(cm, w1, b1, stride=2, padding=2)), 1), "conv2", round(t(lambda: F.conv2d(y1, w2, b2, stride=2, padding=2)), 1),
      "conv3", round(t(lambda: F.conv2d(y2, w3, b3, stride=2, padding=2)), 1), y1.is_contiguous(memory_format=torch.channels_last))
torch.backends.cudnn.benchmark = True
print("benchmark=True conv1", round(t(lambda: F.conv2d(cm, w1, b1, stride=2, padding=2)), 1), "conv2", round(t(lambda: F.conv2d(y1, w2, b2, stride=2, padding=2)), 1),
      "conv3", round(t(lambda: F.conv2d(y2, w3, b3, stride=2, padding=2)), 1))
