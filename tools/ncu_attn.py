"""one launch each of latent_pool / attn_tc (vert_global shape) for an ncu --set full capture"""
import sys, torch
sys.path.insert(0, ".")
from macvo_b200 import ops
dev = "cuda:0"
tok = torch.randn(9600, 80, 128, device=dev); q8 = torch.randn(8, 128, device=dev)
wk = torch.randn(128, 128, device=dev) * 0.1; wv = torch.randn(128, 128, device=dev) * 0.1; bv = torch.randn(128, device=dev)
q = torch.randn(16, 4800, 128, device=dev); k = torch.randn(16, 300, 128, device=dev); v = torch.randn_like(k)
for _ in range(2):
    ops.latent_pool(tok, q8, wk, wv, bv)
    ops.small_attention(q, k, v, 8, True)
torch.cuda.synchronize()
