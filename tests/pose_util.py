"""Shared helpers of the pose tests: seeded synthetic 3-D <-> 2-D problems in the reference's conventions."""
import numpy as np


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_pose_problem(n=300, seed=0, w=1280, h=720, noise_px=0.5, outlier_frac=0.1, pose_noise=(0.05, 0.02)):
    """n world points seen by a camera T_wc = [t, q(x,y,z,w)]: pixels (noise + gross outliers), unit bearing vectors
    (Frame::computeKeypoint: normalised K^-1 [u v 1]), an initial pose guess perturbed by pose_noise (m, rad)."""
    rng = np.random.default_rng(seed)
    f = min(w / 2 / np.tan(np.deg2rad(45 * w / h) / 2), h / 2 / np.tan(np.deg2rad(45) / 2))
    K = np.array([f, f, w / 2, h / 2])
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(0.05, 0.4)
    q = np.concatenate([np.sin(ang / 2) * ax, [np.cos(ang / 2)]])
    t = rng.normal(0, 0.5, 3)
    R = quat_to_R(q)
    uv = np.stack([rng.uniform(30, w - 30, n), rng.uniform(30, h - 30, n)], 1)
    z = rng.uniform(2, 8, n)
    Xc = np.stack([(uv[:, 0] - K[2]) / K[0] * z, (uv[:, 1] - K[3]) / K[1] * z, z], 1)
    X = Xc @ R.T + t
    obs = uv + rng.normal(0, noise_px, uv.shape)
    out = rng.random(n) < outlier_frac
    obs[out] += rng.normal(0, 25, (int(out.sum()), 2))
    bv = np.stack([(obs[:, 0] - K[2]) / K[0], (obs[:, 1] - K[3]) / K[1], np.ones(n)], 1)
    bv /= np.linalg.norm(bv, axis=1, keepdims=True)
    dq_ax = rng.normal(size=3); dq_ax /= np.linalg.norm(dq_ax)
    da = pose_noise[1]
    dq = np.concatenate([np.sin(da / 2) * dq_ax, [np.cos(da / 2)]])
    x1, y1, z1, w1 = dq; x2, y2, z2, w2 = q
    q0 = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                   w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    pose0 = np.concatenate([t + rng.normal(0, pose_noise[0], 3), q0 / np.linalg.norm(q0)])
    return dict(K=K, pose_true=np.concatenate([t, q]), pose0=pose0, X=np.ascontiguousarray(X), uv=np.ascontiguousarray(obs),
                bv=np.ascontiguousarray(bv), outlier_true=out)
