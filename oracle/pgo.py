"""TEST INFRASTRUCTURE — fp64 CPU oracle (numpy) for MAC-VO's two-frame pose-graph optimisation.

Restates, in plain numpy float64:

* residual / covariance / analytic Jacobian of `Analytic_ReprojDisp_TwoFramePGO`
      Module/Optimization/TwoFramePGO/Graphs.py:76-148 (buffers, forward), :201-230 (build_jacobian)
* the Levenberg-Marquardt step of `LM_analytic`
      Module/Optimization/PyposeOptimizers.py:160-194 (control flow kept verbatim: FastTriggs
      re-weighting, A = J^T W J, diag clamp, cumulative damping, accept/reject on the UNWEIGHTED
      Huber loss, <= 16 rejects)
* the driver loop of `TwoFrame_PGO._optimize`
      Module/Optimization/TwoFramePGO/Optimizer.py:82-102 (Huber(0.1), PINV, TrustRegion(radius=1e3),
      StopOnPlateau(steps=10, patience=2, decreasing=1e-5), weight = block_diag(pinv(cov_i)))
* `pypose` 0.6.8 pieces those call (un-vendored third-party dependency, absent here): SE3
  Exp / Inv / Act / product / left retraction, Huber, FastTriggs, PINV, TrustRegion, StopOnPlateau.

PARITY STATUS: the control flow and graph arithmetic are PINNED against the reference's own
`LM_analytic` + `Analytic_ReprojDisp_TwoFramePGO` executed on top of `oracle/pypose_shim`
(tests/golden/pgo_*.pt, tests/test_oracle_golden.py). The pypose internals themselves are PARITY
UNPINNED (no reference test or fixture pins them; they are restated from the published behaviour).
Self-checks: analytic vs finite-difference Jacobian, recovery of a known synthetic pose.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EPS64 = np.finfo(np.float64).eps


# ------------------------------------------------------------------------------------------
# SE3 (pypose layout [tx,ty,tz,qx,qy,qz,qw])
# ------------------------------------------------------------------------------------------
def skew(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    O = np.zeros_like(v[..., 0])
    return np.stack([np.stack([O, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], O, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], O], -1)], -2)


def quat_matrix(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_rot(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    """R(q) p = p + 2 w (v x p) + 2 v x (v x p);  p (..., 3)."""
    v, w = q[:3], q[3]
    uv = 2 * np.cross(np.broadcast_to(v, p.shape), p)
    return p + w * uv + np.cross(np.broadcast_to(v, p.shape), uv)


def so3_exp(phi: np.ndarray) -> np.ndarray:
    theta = np.linalg.norm(phi)
    t2 = theta * theta
    if theta > EPS64:
        imag, real = np.sin(0.5 * theta) / theta, np.cos(0.5 * theta)
    else:
        imag, real = 0.5 - t2 / 48 + t2 * t2 / 3840, 1 - t2 / 8 + t2 * t2 / 384
    return np.concatenate([phi * imag, [real]])


def so3_jl(phi: np.ndarray) -> np.ndarray:
    theta = np.linalg.norm(phi)
    t2 = theta * theta
    if theta > EPS64:
        c1, c2 = (1 - np.cos(theta)) / t2, (theta - np.sin(theta)) / (t2 * theta)
    else:
        c1, c2 = 0.5 - t2 / 24 + t2 * t2 / 720, 1.0 / 6 - t2 / 120 + t2 * t2 / 5040
    K = skew(phi)
    return np.eye(3) + c1 * K + c2 * (K @ K)


def se3_exp(xi: np.ndarray) -> np.ndarray:
    return np.concatenate([so3_jl(xi[3:6]) @ xi[:3], so3_exp(xi[3:6])])


def se3_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.concatenate([quat_rot(a[3:], b[:3]) + a[:3], quat_mul(a[3:], b[3:])])


def se3_inv(a: np.ndarray) -> np.ndarray:
    qi = a[3:] * np.array([-1.0, -1.0, -1.0, 1.0])
    return np.concatenate([-quat_rot(qi, a[:3]), qi])


def se3_act(a: np.ndarray, p: np.ndarray) -> np.ndarray:
    return quat_rot(a[3:], p) + a[:3]


def retract(pose: np.ndarray, step: np.ndarray) -> np.ndarray:
    """pypose `LieTensor.add_`: pose <- Exp(step[:6]) * pose (7th entry of the step ignored)."""
    return se3_mul(se3_exp(step[:6]), pose)


# ------------------------------------------------------------------------------------------
# factor graph: reprojection + disparity residual (Graphs.py:121-148, 201-230)
# ------------------------------------------------------------------------------------------
@dataclass
class GraphData:
    """All values are the fp32 numbers the reference stores, promoted to fp64 (`.to(torch.double)`)."""
    pos_Tw: np.ndarray      # (K,3) NED world points
    kp2_uv: np.ndarray      # (K,2)
    kp2_disp: np.ndarray    # (K,)
    uv_cov: np.ndarray      # (K,3)  sigma_uu, sigma_vv, sigma_uv
    disp_cov: np.ndarray    # (K,)
    fx: float
    fy: float
    cx: float
    cy: float
    baseline: float
    init_pose: np.ndarray   # (7,)
    # graph type (TwoFramePGO/Optimizer.py:51-68): "disp" = reprojection + disparity (Graphs.py:121-148, MACVO_Performant /
    # _Fast), "reproj" = reprojection only (:76-118), "icp" = 3-D point alignment (:33-73, Paper_Reproduce.yaml)
    graph_type: str = "disp"
    pc_obs: np.ndarray | None = None     # icp: (K,3) observed points in the camera frame = pixel2point_NED(pixel2_uv, pixel2_d)
    obs_cov: np.ndarray | None = None    # icp: (K,3,3) obs2_covTc
    pts_cov: np.ndarray | None = None    # icp: (K,3,3) cov_Tw of the map points

    def cov_blocks(self, pose: np.ndarray | None = None) -> np.ndarray:
        """`covariance_array()` of the graph; pose dependent for icp (R Sigma_obs R^T + Sigma_pts, Graphs.py:61-66)"""
        K = self.pos_Tw.shape[0]
        if self.graph_type == "icp":
            R = quat_matrix((self.init_pose if pose is None else pose)[3:])
            return R @ self.obs_cov @ R.T + self.pts_cov
        c = np.zeros((K, 3, 3))
        c[:, 0, 0], c[:, 1, 1] = self.uv_cov[:, 0], self.uv_cov[:, 1]
        c[:, 0, 1] = c[:, 1, 0] = self.uv_cov[:, 2]
        # reproj: the block is 2x2 (Graphs.py:95-101); carried as a 3x3 whose third row / column never meets a residual
        c[:, 2, 2] = self.disp_cov if self.graph_type == "disp" else 1.0
        return c


def residual(g: GraphData, pose: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """-> R (K,3), pos_Tc (K,3).  r = [fx y/x + cx - u, fy z/x + cy - v, fx bl / x - disp] (NED: x fwd)."""
    if g.graph_type == "icp":                       # frame_pose.Act(points_Tc) - points_Tw  (Graphs.py:56-58)
        pw = se3_act(pose, g.pc_obs)
        return pw - g.pos_Tw, pw
    pc = se3_act(se3_inv(pose), g.pos_Tw)
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    third = (1.0 / x) * (g.fx * g.baseline) - g.kp2_disp if g.graph_type == "disp" else np.zeros_like(x)
    r = np.stack([g.fx * y / x + g.cx - g.kp2_uv[:, 0],
                  g.fy * z / x + g.cy - g.kp2_uv[:, 1],
                  third], axis=-1)
    return r, pc


def jacobian(g: GraphData, pose: np.ndarray, pc: np.ndarray) -> np.ndarray:
    """(K,3,7); column 7 is identically 0 (pypose's 7-wide SE3 parameter)."""
    K = pc.shape[0]
    if g.graph_type == "icp":                       # J = [I | -[T p_c]x]  (Graphs.py:151-167); pc holds T p_c here
        J = np.zeros((K, 3, 7))
        J[:, :, :3] = np.eye(3)
        J[:, :, 3:6] = -skew(pc)
        return J
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    x2 = x ** 2
    Jh = np.zeros((K, 2, 3))
    Jh[:, 0, 0], Jh[:, 0, 1] = -g.fx * y / x2, g.fx / x
    Jh[:, 1, 0], Jh[:, 1, 2] = -g.fy * z / x2, g.fy / x
    RT = quat_matrix(pose[3:]).T
    Jp = np.zeros((K, 3, 7))
    Jp[:, :, :3] = -RT
    Jp[:, :, 3:6] = RT @ skew(g.pos_Tw)
    Jr = Jh @ Jp
    Jd = (-(g.baseline * g.fx) / x2).reshape(-1, 1, 1) * Jp[:, 0:1, :]
    if g.graph_type == "reproj":
        Jd = np.zeros_like(Jd)
    return np.concatenate([Jr, Jd], axis=1)


def huber(x: np.ndarray, delta: float) -> np.ndarray:
    """pypose Huber on SQUARED norms."""
    s = np.sqrt(x)
    return np.where(s < delta, x, 2 * delta * s - delta * delta)


def huber_grad_sqrt(x: np.ndarray, delta: float) -> np.ndarray:
    """FastTriggs scale sqrt(rho'(x)): 1 inside, sqrt(delta / sqrt(x)) outside."""
    s = np.sqrt(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.where(s < delta, 1.0, delta / s))


def robust_loss(g: GraphData, pose: np.ndarray, delta: float) -> float:
    r, _ = residual(g, pose)
    return float(huber((r ** 2).sum(-1), delta).sum())


def normal_equations(g: GraphData, pose: np.ndarray, W: np.ndarray, delta: float):
    """-> A (7,7) = Js^T W Js, JTW-weighted rhs b (7,) = -Js^T W Rs, Js (K,3,7), Rs (K,3)."""
    r, pc = residual(g, pose)
    J = jacobian(g, pose, pc)
    s = huber_grad_sqrt((r ** 2).sum(-1, keepdims=True), delta)      # (K,1)
    Rs, Js = s * r, s[:, :, None] * J
    JTW = np.einsum("kai,kab->kib", Js, W)                           # (K,7,3)
    A = np.einsum("kib,kbj->ij", JTW, Js)
    b = -np.einsum("kib,kb->i", JTW, Rs)
    return A, b, Js, Rs


@dataclass
class LMTrace:
    losses: list = field(default_factory=list)
    rejects: list = field(default_factory=list)
    steps: int = 0
    evaluations: int = 0


def lm_solve(g: GraphData, max_steps: int = 10, patience: int = 2, decreasing: float = 1e-5, delta: float = 0.1,
             radius: float = 1e3, reject: int = 16, diag_min: float = 1e-6, diag_max: float = 1e32,
             trace: LMTrace | None = None) -> np.ndarray:
    """Returns the optimised pose (7,) fp64."""
    pose = np.asarray(g.init_pose, dtype=np.float64).copy()
    weights = lambda p: (np.stack([np.linalg.pinv(c, rcond=1e-15) for c in g.cov_blocks(p)]) if g.pos_Tw.shape[0]
                         else np.zeros((0, 3, 3)))
    W = weights(pose)
    # TrustRegion(radius) defaults: high .5, low 1e-3, up 2, down .5, factor .5, clamp [1e-3, 1e5]
    tr = {"damping": 1.0 / radius, "down": 0.5}
    TR_MIN, TR_MAX, HIGH, LOW, UP, DOWN, FACTOR = 1e-3, 1e5, 0.5, 1e-3, 2.0, 0.5, 0.5
    loss = None
    steps, patience_count = 0, 0
    while True:
        if g.graph_type == "icp":       # the driver recomputes `weight` from covariance_array() before every step (Optimizer.py:96-100)
            W = weights(pose)
        A, b, Js, Rs = normal_equations(g, pose, W, delta)
        if loss is None:
            loss = robust_loss(g, pose, delta)
        last = loss
        reject_count = 0
        d = np.clip(np.diag(A).copy(), diag_min, diag_max)
        A[np.diag_indices(7)] = d
        while last <= loss:
            A[np.diag_indices(7)] = np.diag(A) + np.diag(A) * tr["damping"]
            D = np.linalg.pinv(A, rcond=7 * EPS64) @ b
            pose = retract(pose, D)
            loss = robust_loss(g, pose, delta)
            if trace is not None:
                trace.evaluations += 1
            # TrustRegion.update
            JD = np.einsum("kaj,j->ka", Js, D).reshape(-1)
            with np.errstate(divide="ignore", invalid="ignore"):
                quality = (last - loss) / -(JD @ (2 * Rs.reshape(-1) + JD))
            rad = 1.0 / tr["damping"]
            if quality > HIGH:
                rad, tr["down"] = rad * UP, DOWN
            elif quality > LOW:
                tr["down"] = DOWN
            else:
                rad, tr["down"] = rad * tr["down"], tr["down"] * FACTOR
            tr["down"] = max(TR_MIN, min(tr["down"], TR_MAX))
            rad = max(TR_MIN, min(rad, TR_MAX))
            tr["damping"] = 1.0 / rad
            if last < loss and reject_count < reject:
                pose = retract(pose, -D)
                loss, reject_count = last, reject_count + 1
            else:
                break
        # StopOnPlateau.step
        steps += 1
        if trace is not None:
            trace.losses.append(loss)
            trace.rejects.append(reject_count)
            trace.steps = steps
        cont = steps < max_steps
        patience_count = patience_count + 1 if (last - loss) < decreasing else 0
        if patience_count >= patience:
            cont = False
        if reject_count >= reject:
            cont = False
        if not cont:
            break
    return pose


def accumulate_packed(g: GraphData, pose: np.ndarray, delta: float = 0.1, W: np.ndarray | None = None) -> np.ndarray:
    """The 55-entry packed accumulator of the sharded (multi-GPU) path, for a shard `g` of residual blocks:
    [A = Js^T W Js upper 6x6 (21) | b = -Js^T W Rs (6) | G = Js^T Js upper (21) | h = Js^T Rs (6) | robust loss (1)]."""
    if W is None:
        W = np.stack([np.linalg.pinv(c, rcond=1e-15) for c in g.cov_blocks(pose)]) if g.pos_Tw.shape[0] else np.zeros((0, 3, 3))
    A, b, Js, Rs = normal_equations(g, pose, W, delta)
    iu = np.triu_indices(6)
    G = np.einsum("kai,kaj->ij", Js, Js)
    h = np.einsum("kai,ka->i", Js, Rs)
    return np.concatenate([A[:6, :6][iu], b[:6], G[:6, :6][iu], h[:6], [robust_loss(g, pose, delta)]])
