"""TEST INFRASTRUCTURE — numpy float64 restatement of `MotionInterpolate.elaborate_map`
(Module/MapProcessor.py:52-79) and `interpolate_pose` (Utility/Math.py:96-122): the trajectory post-process MAC-VO runs
at `terminate()` — frames flagged `need_interp` (lost track / skipped) get the se3-linear interpolation of the
neighbouring relative motions, then the trajectory is re-integrated with quaternion renormalisation.

pypose semantics used (pypose 0.6.8 is not installable here -> PARITY UNPINNED beyond the shim, see oracle/pypose_shim):
SE3 Log (phi = v * 2 atan(|v| / w) / |v|, tau = Jl^-1(phi) t), `cumops` = inclusive left fold.
Pinned by tests/test_oracle_golden.py::test_motion_interpolate_* against the REFERENCE class executed on the shim.
"""
from __future__ import annotations

import numpy as np

from . import pgo as opgo


def so3_log(q: np.ndarray) -> np.ndarray:
    v, w = q[:3], q[3]
    n = np.linalg.norm(v)
    if n > opgo.EPS64:
        return v * (2.0 * np.arctan(n / w) / n)
    return v * (2.0 / w - 2.0 * n * n / (3.0 * w ** 3))


def so3_jl_inv(phi: np.ndarray) -> np.ndarray:
    theta = np.linalg.norm(phi)
    t2 = theta * theta
    if theta > opgo.EPS64:
        coef = 1.0 / t2 - (1 + np.cos(theta)) / (2 * theta * np.sin(theta))
    else:
        coef = 1.0 / 12 + t2 / 720 + t2 * t2 / 30240
    K = opgo.skew(phi)
    return np.eye(3) - 0.5 * K + coef * (K @ K)


def se3_log(x: np.ndarray) -> np.ndarray:
    phi = so3_log(x[3:7])
    return np.concatenate([so3_jl_inv(phi) @ x[:3], phi])


def normalize_quat(x: np.ndarray) -> np.ndarray:
    y = x.copy()
    y[3:] = y[3:] / np.linalg.norm(y[3:])
    return y


def motion_interpolate(poses: np.ndarray, need_interp: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """poses (F,7) float32 [t, q_xyzw], need_interp (F,) bool -> (new poses (F,7) float32, interpolated motion indices)."""
    P = poses.astype(np.float64)
    F = P.shape[0]
    if F < 2:
        return poses.copy(), np.zeros(0, dtype=np.int64)
    motions = np.stack([opgo.se3_mul(opgo.se3_inv(P[i]), P[i + 1]) for i in range(F - 1)])
    bad = need_interp[1:].astype(bool).copy()
    bad[:2] = False
    bad[-2:] = False
    interp_idx = np.nonzero(bad)[0]
    good_idx = np.nonzero(~bad)[0]
    for i in interp_idx:
        # interpolate_pose: segment [last good before i, first good after i] in the list of good motions
        e = int(np.searchsorted(good_idx, i, side="left"))
        g0, g1 = good_idx[e - 1], good_idx[e]
        t = (i - g0) / (g1 - g0)
        lie = se3_log(opgo.se3_mul(motions[g1], opgo.se3_inv(motions[g0])))
        motions[i] = opgo.se3_mul(opgo.se3_exp(t * lie), motions[g0])
    # pp.cumops(motions, 0, lambda a, b: NormalizeQuat(a) @ NormalizeQuat(b))
    cum = [motions[0]]
    for i in range(1, F - 1):
        cum.append(opgo.se3_mul(normalize_quat(cum[-1]), normalize_quat(motions[i])))
    out = poses.copy()
    for i in range(F - 1):
        out[i + 1] = opgo.se3_mul(P[0], cum[i]).astype(np.float32)
    return out, interp_idx
