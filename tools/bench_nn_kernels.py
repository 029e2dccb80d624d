"""time csrc/nn_kernels.cu against the torch ops they replace at the 640x480 shapes (gpurun)"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from macvo_b200 import ops
dev = "cuda:0"
torch.set_float32_matmul_precision("medium")

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

x = torch.randn(9600 * 80, 128, device=dev); w = torch.randn(128, device=dev); b = torch.randn(128, device=dev)
print("LN 768000x128  torch %.1f us  native %.1f us" % (t(lambda: F.layer_norm(x, (128,), w, b)), t(lambda: ops.layer_norm(x, w, b))))
x2 = torch.randn(9600 * 8, 128, device=dev)
print("LN 76800x128   torch %.1f us  native %.1f us" % (t(lambda: F.layer_norm(x2, (128,), w, b)), t(lambda: ops.layer_norm(x2, w, b))))
maps = torch.randn(9600, 1, 60, 80, device=dev); wt = torch.randn(16, 1, 6, 6, device=dev); bb = torch.randn(16, device=dev)
def torch_conv():
    return F.relu(F.conv2d(F.pad(maps, (0, 0, 0, 4)), wt, bb, stride=2, padding=2))
print("conv1 9600x60x80  torch %.1f us  native fp32 %.1f us  native tf32 %.1f us" % (t(torch_conv), t(lambda: ops.patch_embed_conv1(maps, wt, bb, False)), t(lambda: ops.patch_embed_conv1(maps, wt, bb, True))))
def sd(q, k, v, heads):
    B, J, C = k.shape; d = C // heads
    qh = q.reshape(q.shape[0], -1, heads, d).permute(0, 2, 1, 3).expand(B, -1, -1, -1)
    kh, vh = (z.reshape(B, J, heads, d).permute(0, 2, 1, 3) for z in (k, v))
    return F.scaled_dot_product_attention(qh, kh, vh).permute(0, 2, 1, 3).reshape(B, -1, C)
for name, (B, nq, nk, heads, d, bc) in {"input_layer": (9600, 8, 80, 8, 16, True), "latent": (9600, 8, 8, 8, 16, False),
        "decoder_cross": (9600, 1, 8, 8, 8, False), "vert_local": (1728, 49, 49, 8, 16, False),
        "vert_global": (16, 4800, 300, 8, 16, False), "svt_s0_local": (4 * 18 * 23, 49, 49, 4, 32, False),
        "svt_s0_global": (4, 19200, 300, 4, 32, False), "svt_s1_global": (4, 4800, 300, 8, 32, False)}.items():
    q = torch.randn(1 if bc else B, nq, heads * d, device=dev); k = torch.randn(B, nk, heads * d, device=dev); v = torch.randn_like(k)
    print("attn %-14s torch-sdpa %.1f us  native fp32 %.1f us  native tf32 %.1f us" % (
        name, t(lambda: sd(q, k, v, heads)), t(lambda: ops.small_attention(q, k, v, heads, False)),
        t(lambda: ops.small_attention(q, k, v, heads, True))))

tok = torch.randn(9600, 80, 128, device=dev); q8 = torch.randn(8, 128, device=dev); wk = torch.randn(128, 128, device=dev) * 0.1; wv = torch.randn(128, 128, device=dev) * 0.1; bv = torch.randn(128, device=dev)
print("latent_pool 9600x80x128: %.1f us" % t(lambda: ops.latent_pool(tok, q8, wk, wv, bv)))
