"""Seeded synthetic inputs shared by the oracle, the GPU tests and bench.py.

* texture : 3 octaves of uniform noise, Gaussian-blurred (sigma 1.5 / 3 / 6, weights proportional to sigma),
            contrast-normalised to mean 128 / std 60 and clipped -- about 15 k FAST-9 (thr 20, NMS) corners per
            1280x720 view, so 1000 features/frame is always reachable.
* frames  : a pinhole camera on a smooth seeded SE(3) trajectory (about 3 px/frame lateral motion, +-0.2 deg/frame
            rotation) looking at the texture plane z = 4; bilinear sampling; RGBA with R = G = B, A = 255.
* descriptors / BA problems for the matcher and bundle-adjustment stages.
Pure numpy (+ scipy.ndimage for the blur); nothing here is on the product path.
"""
import numpy as np

_TEX_CACHE = {}


def make_texture(size=2048, seed=1234):
    key = (size, seed)
    if key not in _TEX_CACHE:
        from scipy.ndimage import gaussian_filter
        rng = np.random.default_rng(seed)
        acc = np.zeros((size, size), np.float32)
        for sigma in (1.5, 3.0, 6.0):
            n = rng.random((size, size), dtype=np.float32)
            b = gaussian_filter(n, sigma, mode="wrap")
            b = (b - b.mean()) / b.std()
            acc += sigma * b
        acc = (acc - acc.mean()) / acc.std() * 60.0 + 128.0
        _TEX_CACHE[key] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return _TEX_CACHE[key]


def intrinsics(w, h):
    """AlvaAR's JS shim: f = min(w/2 / tan(fovH/2), h/2 / tan(fovV/2)), fovV = 45 deg, fovH = 45 deg * aspect
    (reference: src/system.js:101-123)."""
    fov_v = np.deg2rad(45.0)
    fov_h = fov_v * (w / h)
    f = min(w / 2 / np.tan(fov_h / 2), h / 2 / np.tan(fov_v / 2))
    return float(f), float(f), w / 2.0, h / 2.0


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def trajectory(nframes, w, h, seed=99, plane_z=4.0):
    """Camera-to-world poses (R_wc, t_wc): smooth random walk, ~3 px/frame lateral, +-0.2 deg/frame."""
    rng = np.random.default_rng(seed)
    fx = intrinsics(w, h)[0]
    step = 3.0 * plane_z / fx
    poses = []
    ang = np.zeros(3)
    pos = np.zeros(3)
    dirn = rng.uniform(0, 2 * np.pi)
    for k in range(nframes):
        poses.append((_rot(*ang), pos.copy()))
        dirn += rng.normal(0, 0.05)
        pos = pos + np.array([np.cos(dirn) * step, np.sin(dirn) * step, rng.normal(0, 0.002)])
        ang = np.clip(ang + np.deg2rad(rng.uniform(-0.2, 0.2, 3)), -0.15, 0.15)
    return poses


def render_gray(tex, pose, w, h, plane_z=4.0, tex_scale=None):
    """Bilinear render of the texture plane z = plane_z seen from pose (R_wc, t_wc)."""
    fx, fy, cx, cy = intrinsics(w, h)
    if tex_scale is None:
        tex_scale = fx / plane_z          # texture pixels per world unit: ~1 texel per image pixel
    R, t = pose
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    rays = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1) @ R.T
    s = (plane_z - t[2]) / rays[..., 2]
    X = t[0] + s * rays[..., 0]
    Y = t[1] + s * rays[..., 1]
    S = tex.shape[0]
    tu = (X * tex_scale + S / 2.0) % S
    tv = (Y * tex_scale + S / 2.0) % S
    x0 = np.floor(tu).astype(np.int64)
    y0 = np.floor(tv).astype(np.int64)
    ax = (tu - x0).astype(np.float32)
    ay = (tv - y0).astype(np.float32)
    x1 = (x0 + 1) % S
    y1 = (y0 + 1) % S
    T = tex.astype(np.float32)
    val = (T[y0, x0] * (1 - ax) * (1 - ay) + T[y0, x1] * ax * (1 - ay) + T[y1, x0] * (1 - ax) * ay + T[y1, x1] * ax * ay)
    return np.clip(np.rint(val), 0, 255).astype(np.uint8)


def gray_to_rgba(g):
    out = np.empty(g.shape + (4,), np.uint8)
    out[..., 0] = g
    out[..., 1] = g
    out[..., 2] = g
    out[..., 3] = 255
    return out


def make_frames(nframes, w, h, seed=99, rgba=True):
    tex = make_texture()
    poses = trajectory(nframes, w, h, seed)
    frames = [render_gray(tex, p, w, h) for p in poses]
    g = np.stack(frames)
    return (gray_to_rgba(g) if rgba else g), poses


def crop(w, h, ox=300, oy=200):
    """A plain crop of the texture (cheap test input with the same statistics)."""
    return np.ascontiguousarray(make_texture()[oy:oy + h, ox:ox + w])


def random_rgba(w, h, n=1, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 4), dtype=np.uint8)


def make_descriptors(nq=1000, nt=10000, seed=7, planted=0.3, flip=0.08):
    """Random 256-bit strings with `planted` fraction of true matches at Hamming ~ Binomial(256, flip)."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    npl = int(nq * planted)
    idx = rng.choice(nt, npl, replace=False)
    noise = np.packbits(rng.random((npl, 256)) < flip, axis=1)
    q[:npl] = t[idx] ^ noise
    return q, t
