"""Synthetic textured-room scenes.

The reference's datasets (Stanford2D-3D-S, OmniScenes; `data_utils.py:16-135`) are not available
offline, so every config in BASELINE.json runs on a procedurally textured box room:

* points: sampled uniformly by area on the six faces of an axis-aligned box, coloured by a fixed
  non-periodic texture and quantised to uint8/255 (the reference only ever sees uint8/255 colours:
  `data_utils.py:33`, `color_utils.py:60-61`);
* panorama: rendered analytically per pixel (ray/box intersection) from a ground-truth pose using
  the inverse of the projection convention of `utils.py:16-61` + `F.grid_sample(align_corners=False)`
  (`utils.py:86`), same texture, same quantisation; top/bottom H/16 rows are exact black (the
  Stanford black caps, which exercise the zero mask of `omniloc.py:198`).

Everything is float64 numpy and seeded (seed 2 = the reference's seed, `localize.py:95-98`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

ROOM_DEFAULT = (8.0, 6.0, 3.0)


def rot_zyx(yaw: float, pitch: float, roll: float) -> np.ndarray:
    """R = Rz(yaw) @ Ry(pitch) @ Rx(roll)  (convention of `utils.py:425-453`)."""
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    rz = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cr, -sr], [0.0, sr, cr]])
    return rz @ ry @ rx


def texture(p: np.ndarray) -> np.ndarray:
    """Procedural colour field tex(x, y, z) -> (…,3) in [0.06, 0.94]; never exact black."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    cell = np.floor(x / 0.55) + np.floor(y / 0.45) + np.floor(z / 0.6)
    checker = np.mod(cell, 2.0)
    blob = np.sin(0.37 * x * x - 0.21 * y * z)
    r = 0.50 + 0.22 * np.sin(1.7 * x + 0.9 * y + 2.3 * z) + 0.14 * np.sin(4.3 * y - 1.1 * z) + 0.10 * (checker - 0.5) + 0.06 * blob
    g = 0.48 + 0.20 * np.sin(2.9 * y - 1.3 * z + 0.7 * x) + 0.15 * np.cos(3.7 * x + 0.4 * z) - 0.12 * (checker - 0.5) + 0.07 * blob
    b = 0.52 + 0.21 * np.sin(1.1 * z + 2.1 * x - 1.9 * y) + 0.13 * np.cos(5.1 * z + 0.8 * y) + 0.08 * (checker - 0.5) - 0.08 * blob
    return np.clip(np.stack([r, g, b], axis=-1), 0.06, 0.94)


def quantise(c: np.ndarray) -> np.ndarray:
    """round(255 c) as uint8."""
    return np.clip(np.rint(c * 255.0), 0, 255).astype(np.uint8)


def sample_room_points(n: int, room=ROOM_DEFAULT, seed: int = 2, origin=(0.0, 0.0, 0.0)):
    """N points uniform by area on the six faces.  Returns xyz float32 (N,3), rgb uint8 (N,3)."""
    rng = np.random.default_rng(seed)
    lx, ly, lz = room
    areas = np.array([ly * lz, ly * lz, lx * lz, lx * lz, lx * ly, lx * ly])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    a = rng.random(n)
    b = rng.random(n)
    p = np.empty((n, 3))
    for f in range(6):
        m = face == f
        axis, side = f // 2, f % 2
        dims = [lx, ly, lz]
        others = [k for k in range(3) if k != axis]
        p[m, axis] = side * dims[axis]
        p[m, others[0]] = a[m] * dims[others[0]]
        p[m, others[1]] = b[m] * dims[others[1]]
    rgb8 = quantise(texture(p))
    p += np.asarray(origin)[None, :]
    return p.astype(np.float32), rgb8


def pixel_rays(h: int, w: int, rows=None) -> np.ndarray:
    """Camera-frame unit directions of pixel centres (inverse of `utils.py:44-59` + grid_sample
    align_corners=False): u=2(col+.5)/W-1, v=2(row+.5)/H-1, phi=pi(1-u), theta=pi(v+1)/2."""
    rows = np.arange(h) if rows is None else rows
    u = 2.0 * (np.arange(w) + 0.5) / w - 1.0
    v = 2.0 * (rows + 0.5) / h - 1.0
    phi = np.pi * (1.0 - u)
    theta = np.pi * (v + 1.0) / 2.0
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    d = np.stack([st * np.cos(phi - np.pi)[None, :], st * np.sin(phi - np.pi)[None, :], np.broadcast_to(ct, (len(rows), w))], axis=-1)
    return d


def render_panorama(pose, h: int, w: int, room=ROOM_DEFAULT, origin=(0.0, 0.0, 0.0), black_caps: bool = True) -> np.ndarray:
    """Analytic equirectangular render of the textured room from pose=(tx,ty,tz,yaw,pitch,roll).
    Returns uint8 (H,W,3)."""
    t = np.asarray(pose[:3], dtype=np.float64) - np.asarray(origin)
    R = rot_zyx(*[float(a) for a in pose[3:6]])
    lo = np.zeros(3)
    hi = np.asarray(room, dtype=np.float64)
    out = np.zeros((h, w, 3), dtype=np.uint8)
    cap = h // 16 if black_caps else 0
    chunk = max(1, (1 << 21) // w)
    for r0 in range(cap, h - cap, chunk):
        rows = np.arange(r0, min(h - cap, r0 + chunk))
        dw = pixel_rays(h, w, rows) @ R  # world direction = R^T q  (row-vector form)
        with np.errstate(divide="ignore", invalid="ignore"):
            t_lo = (lo - t) / dw
            t_hi = (hi - t) / dw
        far = np.where(dw > 0, t_hi, t_lo)
        far = np.where(dw == 0, np.inf, far)
        dist = far.min(axis=-1)
        hit = t + dist[..., None] * dw
        hit = np.clip(hit, lo, hi)
        out[rows] = quantise(texture(hit))
    return out


def rgb_from_u8(rgb8: np.ndarray) -> np.ndarray:
    """Point colours as the reference reads them: float64 /255 then .float() (`data_utils.py:33`,
    `localize.py:160`)."""
    return (rgb8.astype(np.float64) / 255.0).astype(np.float32)


def img_from_u8(img8: np.ndarray) -> np.ndarray:
    """Panorama as the reference reads it: uint8 -> float32, then float32 division by 255
    (`localize.py:169`)."""
    return img8.astype(np.float32) / np.float32(255.0)


@dataclass
class Scene:
    xyz: np.ndarray      # (N,3) float32
    rgb8: np.ndarray     # (N,3) uint8
    img8: np.ndarray     # (H,W,3) uint8
    gt_pose: np.ndarray  # (6,) float64: tx,ty,tz,yaw,pitch,roll
    room: tuple

    @property
    def rgb(self) -> np.ndarray:
        return rgb_from_u8(self.rgb8)

    @property
    def img(self) -> np.ndarray:
        return img_from_u8(self.img8)


def random_gt_pose(room=ROOM_DEFAULT, seed: int = 2, yaw_only: bool = False, tilt: float = 0.1) -> np.ndarray:
    rng = np.random.default_rng(seed + 7919)
    frac = 0.2 + 0.6 * rng.random(3)
    t = frac * np.asarray(room)
    yaw = rng.random() * 2 * np.pi
    pitch, roll = (0.0, 0.0) if yaw_only else tuple((rng.random(2) * 2 - 1) * tilt)
    return np.array([t[0], t[1], t[2], yaw, pitch, roll])


def make_scene(n: int, h: int, w: int, room=ROOM_DEFAULT, seed: int = 2, gt_pose=None, yaw_only: bool = False,
               black_caps: bool = True) -> Scene:
    xyz, rgb8 = sample_room_points(n, room, seed)
    if gt_pose is None:
        gt_pose = random_gt_pose(room, seed, yaw_only)
    img8 = render_panorama(gt_pose, h, w, room, black_caps=black_caps)
    return Scene(xyz, rgb8, img8, np.asarray(gt_pose, dtype=np.float64), tuple(room))


def perturb_panorama(img8: np.ndarray, seed: int, gamma: float = 1.0, const: float = 1.0, wb=(1.0, 1.0, 1.0),
                     retexture_frac: float = 0.0) -> np.ndarray:
    """Colour / scene perturbations for the multi-query config (the synthetic illumination change of
    `localize.py:384-393`: constant division, gamma, white balance) plus re-textured patches."""
    rng = np.random.default_rng(seed)
    x = img8.astype(np.float64) / 255.0
    x = x / const
    x = np.power(x, gamma)
    x = x * np.asarray(wb)[None, None, :]
    h, w, _ = x.shape
    n_patch = int(retexture_frac * 64)
    for _ in range(n_patch):
        ph, pw = h // 8, w // 16
        r0 = rng.integers(h // 16, h - h // 16 - ph)
        c0 = rng.integers(0, w - pw)
        x[r0:r0 + ph, c0:c0 + pw] = 0.2 + 0.6 * rng.random(3)[None, None, :]
    black = (img8.sum(axis=-1) == 0)
    out = quantise(np.clip(x, 0.0, 1.0))
    out[black] = 0
    return out


def pose_grid(room=ROOM_DEFAULT, n_xyz=(5, 5, 3), n_yaw: int = 24, seed: int = 2) -> np.ndarray:
    """Explicit (P,6) start-pose grid: translations on a regular lattice inside the 10-90 % box,
    yaw uniformly spaced, pitch=roll=0.  Pose index = i_trans * n_yaw + j_rot (`utils.py:501-505`)."""
    axes = [np.linspace(0.1 * room[k], 0.9 * room[k], n_xyz[k]) if n_xyz[k] > 1 else np.array([0.5 * room[k]]) for k in range(3)]
    tx, ty, tz = np.meshgrid(*axes, indexing="ij")
    trans = np.stack([tx.ravel(), ty.ravel(), tz.ravel()], axis=-1)
    yaw = np.arange(n_yaw) * 2 * np.pi / n_yaw
    rot = np.stack([yaw, np.zeros_like(yaw), np.zeros_like(yaw)], axis=-1)
    poses = np.concatenate([np.repeat(trans, n_yaw, axis=0), np.tile(rot, (len(trans), 1))], axis=-1)
    return poses.astype(np.float32)
