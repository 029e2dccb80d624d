"""TEST INFRASTRUCTURE — CPU oracle for MAC-VO's 2D->3D observation covariance model.

Restates (torch CPU, fp32 like the reference, result cast to fp64 at the end):

* `match_covariance` <- MatchCovariance.estimate       Module/Covariance/Project2to3.py:124-182
                        gaussain_full_kernels          Utility/Math.py:44-63
                        Covariance_2to3_full           Module/Covariance/Project2to3.py:377-424
                        create_3x3_matrix (CPU fp32)   Module/Covariance/Project2to3.py:426-434
* `pixel2point_ned`  <- pixel2point_NED                Utility/Point.py:15-17 (+ pypose.pixel2point)

Reference quirks kept on purpose:
  - `flow_cov[:, :2]` of the CALLER's tensor is clamped in place to >= min_flow_cov**2;
  - the Gaussian kernel's first axis (weighted with sigma_uu) is paired with image ROWS of the
    depth patch (the `permute(0, 2, 1)` at Project2to3.py:158);
  - the `depth_cov` argument is ignored whenever `flow_cov` is given.

PINNED by tests/golden/covariance_*.pt (generated from the reference class itself).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


def gaussian_full_kernels(cov_2x2: Tensor, kernel_size: int) -> Tensor:
    N = cov_2x2.size(0)
    det_cov = cov_2x2.det()
    inv_cov = cov_2x2.pinverse().float()
    half = (kernel_size - 1) / 2.0
    x = torch.linspace(-half, half, kernel_size)
    idx = torch.stack(torch.meshgrid(x, x, indexing="ij"), dim=-1).unsqueeze(0).repeat(N, 1, 1, 1)
    z = torch.einsum("bxyi,bij,bxyj->bxy", idx, -0.5 * inv_cov, idx).exp()
    kernel = z / (2 * torch.pi * torch.sqrt(det_cov)).view(N, 1, 1)
    return kernel / kernel.sum(dim=[-1, -2], keepdim=True)


def match_covariance(kp: Tensor, depth_map: Tensor, flow_cov: Tensor | None, fx: float, fy: float, cx: float,
                     cy: float, kernel_size: int = 31, min_flow_cov: float = 0.25, min_depth_cov: float = 0.05,
                     match_cov_default: float = 0.25, depth_cov: Tensor | None = None) -> Tensor:
    """kp (K,2) int64 or fp32 [u,v]; depth_map (1,1,H,W); flow_cov (K,3) or None -> (K,3,3) float64."""
    n = kp.size(0)
    half = kernel_size // 2
    kp_long = kp.clone().long()
    has_flow_cov = flow_cov is not None
    if has_flow_cov:
        flow_cov[..., :2].clamp_(min=min_flow_cov ** 2)          # in place on the caller's tensor
    else:
        flow_cov = torch.ones((n, 3), dtype=torch.float) * match_cov_default
        flow_cov[..., 2] = 0.0
    var_u, var_v, var_uv = flow_cov[..., 0], flow_cov[..., 1], flow_cov[..., 2]
    kp_u, kp_v = kp[..., 0], kp[..., 1]

    off = torch.arange(-half, half + 1, dtype=torch.long)
    uu, vv = torch.meshgrid(off, off, indexing="ij")
    all_u = kp_long[:, 0].unsqueeze(-1) + uu.reshape(1, -1)
    all_v = kp_long[:, 1].unsqueeze(-1) + vv.reshape(1, -1)

    cov2 = torch.empty((n, 2, 2))
    cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 0], cov2[:, 1, 1] = var_u, var_uv, var_uv, var_v
    filt = gaussian_full_kernels(cov2, kernel_size)

    patches = depth_map[..., all_v, all_u].view(n, kernel_size, kernel_size).permute(0, 2, 1)
    wavg = (filt * patches).sum(dim=[1, 2])
    if has_flow_cov or depth_cov is None:
        wvar = torch.sum(filt * (patches - wavg.unsqueeze(1).unsqueeze(1)).square(), dim=[1, 2])
    else:
        wvar = depth_cov
    wvar = wvar.clamp(min=min_depth_cov)

    u, v, d = kp_u, kp_v, wavg
    s_xx = (((u - cx).square() * wvar) + (d.square() * var_u) + (var_u * wvar)) / (fx ** 2)
    s_yy = (((v - cy).square() * wvar) + (d.square() * var_v) + (var_v * wvar)) / (fy ** 2)
    s_zz = wvar
    s_xy = (((u - cx) * (v - cy) * wvar) + (d.square() + wvar) * var_uv) / (fx * fy)
    s_xz = (wvar * (u - cx)) / fx
    s_yz = (wvar * (v - cy)) / fy
    mat = torch.empty((n, 3, 3))
    rows = [[s_zz, s_xz, s_yz], [s_xz, s_xx, s_xy], [s_yz, s_xy, s_yy]]
    for i in range(3):
        for j in range(3):
            mat[..., i, j] = rows[i][j]
    return mat.double()


def pixel2point_ned(pixels: Tensor, depths: Tensor, K: Tensor) -> Tensor:
    """[(u-cx)/fx d, (v-cy)/fy d, d] rolled to NED [d, x, y]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    x = (pixels[..., 0] - cx) / fx * depths
    y = (pixels[..., 1] - cy) / fy * depths
    return torch.stack([x, y, depths], dim=-1).roll(shifts=1, dims=-1)
