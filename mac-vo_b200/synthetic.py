"""Seeded synthetic stereo sequences of TartanAir shape: no dataset / checkpoint exists
offline, so benchmarks and end-to-end tests run on these (and say so in their `data` field)."""
from __future__ import annotations

import torch

from .interfaces import StereoData


def camera(H: int, W: int) -> tuple[torch.Tensor, float]:
    """TartanAir: fx = fy = cx = 320, cy = 240, baseline 0.25 at 640x480; scaled for other sizes."""
    f = 320.0 * W / 640.0
    K = torch.tensor([[[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]]])
    return K, 0.25


def make_sequence(n_frames: int, H: int = 480, W: int = 640, seed: int = 1000, pin: bool = False) -> list[StereoData]:
    """imageL_t = smoothed noise; imageR_t = roll(imageL_t, -8 px); imageL_{t+1} = roll(imageL_t, (+2, +3))."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, H + 16, W + 16, generator=g)
    base = torch.nn.functional.avg_pool2d(base, 9, stride=1, padding=4)
    base = (base - base.min()) / (base.max() - base.min())
    K, bl = camera(H, W)
    frames = []
    for t in range(n_frames):
        cur = torch.roll(base, shifts=(2 * t, 3 * t), dims=(2, 3))
        left = cur[..., 8:8 + H, 8:8 + W].contiguous()
        right = torch.roll(cur, shifts=-8, dims=3)[..., 8:8 + H, 8:8 + W].contiguous()
        if pin:
            left, right = left.pin_memory(), right.pin_memory()
        frames.append(StereoData(T_BS=None, K=K.clone(), baseline=torch.tensor([bl]), time_ns=[t * 100_000_000],
                                 height=H, width=W, imageL=left, imageR=right))
    return frames
