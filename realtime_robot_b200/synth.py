"""Deterministic synthetic clouds (SURVEY.md 8d recipe): counter-hash uniforms, area-uniform surface samples at
rho = 4700 pts/m^2 (the repo clouds' density), sigma = 1 mm isotropic noise, base seed 20170427.

Harness-side input generation only (numpy); no registration arithmetic lives here.
"""
import numpy as np

BASE_SEED = 20170427
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, start: int, count: int) -> np.ndarray:
    """u(seed, i) = (splitmix64(seed * GOLD + i) >> 40) * 2^-24 for i in [start, start + count)."""
    with np.errstate(over="ignore"):
        i = np.arange(start, start + count, dtype=np.uint64)
        z = _mix64(np.uint64(seed) * _GOLD + i)
    return ((z >> np.uint64(40)).astype(np.float64) * (2.0 ** -24))


def gaussian(seed: int, start: int, count: int) -> np.ndarray:
    u1 = uniform(seed, 2 * start, 2 * count)[0::2]
    u2 = uniform(seed, 2 * start, 2 * count)[1::2]
    return np.sqrt(-2.0 * np.log(np.maximum(u1, 2.0 ** -24))) * np.cos(2.0 * np.pi * u2)


def _box_faces(lo, hi):
    """Six rectangles (origin, edge u, edge v) of an axis-aligned box."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    faces = []
    for a in range(3):
        b, c = (a + 1) % 3, (a + 2) % 3
        eu = np.zeros(3); eu[b] = d[b]
        ev = np.zeros(3); ev[c] = d[c]
        o0 = lo.copy()
        o1 = lo.copy(); o1[a] = hi[a]
        faces += [(o0, eu, ev), (o1, eu, ev)]
    return faces


def sample_rects(rects, n: int, seed: int, noise: float = 0.001) -> np.ndarray:
    """n points, uniform by area over the rectangles, + isotropic Gaussian noise.  Returns (n, 4) float32 xyz1."""
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in rects])
    cum = np.cumsum(areas) / areas.sum()
    pick = np.searchsorted(cum, uniform(seed, 0, n), side="right").clip(0, len(rects) - 1)
    a = uniform(seed + 1, 0, n)[:, None]
    b = uniform(seed + 2, 0, n)[:, None]
    O = np.stack([r[0] for r in rects])[pick]
    U = np.stack([r[1] for r in rects])[pick]
    V = np.stack([r[2] for r in rects])[pick]
    p = O + a * U + b * V
    if noise > 0:
        p = p + noise * np.stack([gaussian(seed + 3 + k, 0, n) for k in range(3)], axis=1)
    out = np.ones((n, 4), dtype=np.float32)
    out[:, :3] = p.astype(np.float32)
    return out


def chair_rects(origin=(0.0, 0.0, 0.0)):
    """Seat 0.45 x 0.45 x 0.05 at z = 0.45, back 0.45 x 0.05 x 0.45, four 0.04 x 0.04 x 0.45 legs."""
    o = np.asarray(origin, float)
    boxes = [((0, 0, 0.45), (0.45, 0.45, 0.50)), ((0, 0.40, 0.50), (0.45, 0.45, 0.95))]
    for x in (0.0, 0.41):
        for y in (0.0, 0.41):
            boxes.append(((x, y, 0.0), (x + 0.04, y + 0.04, 0.45)))
    rects = []
    for lo, hi in boxes:
        rects += _box_faces(np.asarray(lo) + o, np.asarray(hi) + o)
    return rects


def room_rects(size=(8.0, 8.0, 3.0), n_boxes=12, seed=BASE_SEED, chair_at=None):
    """Floor + 4 walls + random furniture boxes (+ optionally the chair primitive somewhere on the floor)."""
    sx, sy, sz = size
    rects = [(np.zeros(3), np.array([sx, 0, 0.]), np.array([0, sy, 0.])),
             (np.zeros(3), np.array([sx, 0, 0.]), np.array([0, 0, sz])),
             (np.array([0, sy, 0.]), np.array([sx, 0, 0.]), np.array([0, 0, sz])),
             (np.zeros(3), np.array([0, sy, 0.]), np.array([0, 0, sz])),
             (np.array([sx, 0, 0.]), np.array([0, sy, 0.]), np.array([0, 0, sz]))]
    u = uniform(seed + 100, 0, 6 * n_boxes).reshape(n_boxes, 6)
    for k in range(n_boxes):
        w, d, h = 0.3 + 1.2 * u[k, 0], 0.3 + 1.2 * u[k, 1], 0.3 + 1.0 * u[k, 2]
        x, y = u[k, 3] * (sx - w), u[k, 4] * (sy - d)
        rects += _box_faces((x, y, 0.0), (x + w, y + d, h))
    if chair_at is not None:
        rects += chair_rects(chair_at)
    return rects


def rigid(rx_deg=0.0, ry_deg=0.0, rz_deg=0.0, t=(0.0, 0.0, 0.0), about=(0.0, 0.0, 0.0)) -> np.ndarray:
    """4x4 float64 rigid transform: rotate (Z * Y * X) about a point, then translate."""
    rx, ry, rz = np.deg2rad([rx_deg, ry_deg, rz_deg])
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    c = np.asarray(about, float)
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = c - R @ c + np.asarray(t, float)
    return M


def apply(M, xyz1) -> np.ndarray:
    out = np.ones((len(xyz1), 4), dtype=np.float32)
    out[:, :3] = (np.asarray(xyz1[:, :3], np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    return out


def icp_config(n_model=100_000, n_scene=1_000_000, seed=BASE_SEED):
    """BASELINE.json configs[2]: chair primitive (n_model pts) vs a room containing that chair (n_scene pts), with the
    SURVEY 8d ground-truth offset: Rz-dominant 5 deg about (0.2, 0.3, 0.93) + t = (0.02, -0.015, 0.01).
    Returns (model_xyz1 already displaced by the inverse pose, scene_xyz1, ground-truth model->scene 4x4)."""
    chair_pos = (3.0, 2.5, 0.0)
    model_in_scene = sample_rects(chair_rects(chair_pos), n_model, seed)
    # room area ~ n_scene / 4700 m^2 : scale the footprint to reach the target point count at the repo density
    target_area = n_scene / 4700.0
    side = max(4.0, (-12.0 + np.sqrt(144.0 + 4.0 * (target_area * 0.8))) / 2.0)
    scene = sample_rects(room_rects((side, side, 3.0), n_boxes=max(4, int(side)), seed=seed, chair_at=chair_pos), n_scene, seed + 7)
    gt = rigid(0.5, -0.4, 5.0, (0.02, -0.015, 0.01), about=np.asarray(chair_pos) + np.array([0.2, 0.3, 0.93]))
    model = apply(np.linalg.inv(gt), model_in_scene)
    return model, scene, gt
