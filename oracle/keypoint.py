"""TEST INFRASTRUCTURE — CPU oracle for MAC-VO's keypoint selectors (bit-exact target).

Restates, with the same torch-CPU primitives the reference uses (so that `torch.median`'s
lower-median rule, `torch.nonzero`'s row-major order and the CPU `torch.randperm` stream are
identical):

* `cov_aware_select_nodepth` <- CovAwareSelector_NoDepth.select_point  Module/KeypointSelector.py:362-407
* `mapping_select`           <- MappingPointSelector.select_point      Module/KeypointSelector.py:87-100
* `cov_aware_select`         <- CovAwareSelector.select_point          Module/KeypointSelector.py:260-334
* `candidate_mask_nodepth`   — the deterministic part of the first (everything before randperm)

PINNED by tests/golden/selector_*.pt (generated from the reference classes themselves).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


def _border_mask(like: Tensor, mask_width: int) -> Tensor:
    border = torch.zeros_like(like, dtype=torch.bool)
    # NB: with mask_width == 0 the reference slice `0:-0` is empty -> nothing is selectable
    border[..., mask_width:-mask_width, mask_width:-mask_width] = True
    return border


def candidate_mask_nodepth(match_cov: Tensor, kernel_size: int, mask_width: int, max_match_cov: float,
                           match_mask: Tensor | None = None) -> tuple[Tensor, float]:
    """(1,3,H,W) fp32 -> (bool mask (1,1,H,W), threshold as the fp32-rounded python float)."""
    quality = (match_cov[:, 0] + match_cov[:, 1] - 2 * match_cov[:, 2]).unsqueeze(1)
    eroded = -torch.nn.functional.max_pool2d(-quality, kernel_size=kernel_size, stride=1, padding=kernel_size // 2)
    nms = torch.logical_and(quality == eroded, ~quality.isnan())
    border = _border_mask(nms, mask_width)
    thresh = min(max_match_cov, quality[nms].median().item() * 1.5)
    mask = nms & border & (quality < thresh)
    if match_mask is not None:
        mask = mask & match_mask
    return mask, thresh


def _sample(mask: Tensor, num_point: int) -> Tensor:
    selected = torch.nonzero(mask, as_tuple=False)
    perm = torch.randperm(selected.size(0))[:num_point]
    return selected[perm][..., 2:].roll(shifts=1, dims=1)


def cov_aware_select_nodepth(match_cov: Tensor, num_point: int, kernel_size: int = 7, mask_width: int = 32,
                             max_match_cov: float = 100.0, match_mask: Tensor | None = None) -> Tensor:
    """-> int64 (K, 2) in (u, v) order; consumes one `torch.randperm` from the CPU default generator."""
    mask, _ = candidate_mask_nodepth(match_cov, kernel_size, mask_width, max_match_cov, match_mask)
    return _sample(mask, num_point)


def cov_aware_select(match_cov: Tensor, depth0: Tensor, depth0_cov: Tensor, depth1: Tensor, depth1_cov: Tensor,
                     num_point: int, kernel_size: int = 7, mask_width: int = 32, max_depth: float = 80.0,
                     max_depth_cov: float = 250.0, max_match_cov: float = 100.0, depth0_mask: Tensor | None = None,
                     match_mask: Tensor | None = None) -> Tensor:
    """CovAwareSelector.select_point (Module/KeypointSelector.py:260-334), the depth-aware variant."""
    quality = depth0_cov + depth1_cov
    fq = (match_cov[:, 0] + match_cov[:, 1] - 2 * match_cov[:, 2]).unsqueeze(1)
    quality = quality * fq
    eroded = -torch.nn.functional.max_pool2d(-quality, kernel_size=kernel_size, stride=1, padding=kernel_size // 2)
    nms = torch.logical_and(quality == eroded, ~quality.isnan())
    border = _border_mask(nms, mask_width)
    depth_mask = (depth0 < max_depth) & (depth1 < max_depth)
    thr0 = min(max_depth_cov, depth0_cov[nms].nanmedian().item() * 1.5)
    thrf = min(max_match_cov, fq[nms].nanmedian().item() * 1.5)
    mask = nms & border & depth_mask & (depth0_cov < thr0) & (fq < thrf)
    if depth0_mask is not None:
        mask = mask & depth0_mask
    if match_mask is not None:
        mask = mask & match_mask
    return _sample(mask, num_point)


def mapping_select(depth: Tensor, depth_cov: Tensor, num_point: int, max_depth: float = 5.0,
                   max_depth_cov: float = 0.005, mask_width: int = 32) -> Tensor:
    cand = (depth < max_depth) & (depth_cov < max_depth_cov) & _border_mask(depth, mask_width)
    return _sample(cand, num_point)
