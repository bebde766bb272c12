"""Drop-in for the hot-path functions of the reference's `utils.py`: `trim_input_loss` (:462-507),
`quantile` (:208-229), `rot_from_ypr` (:425-453).  CUDA tensors only."""
from __future__ import annotations

from typing import Tuple

import torch

from . import engine


def grid_poses(trans: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
    """(T,3) x (R,3) -> (T*R, 6) with pose index i*R + j  (loop order of utils.py:484-485)."""
    T, Rn = trans.shape[0], rot.shape[0]
    return torch.cat([trans.to(torch.float32).repeat_interleave(Rn, 0), rot.to(torch.float32).repeat(T, 1)], dim=1)


def score_grid(img, xyz, rgb, trans, rot, q: float = None) -> torch.Tensor:
    """loss_table (T,R) of utils.py:481-499 from one kernel launch (structured-grid scoring, pcl_grid.cu).
    The clamp-box quantile plays no part in scoring: q=None shares whatever packed cloud is cached."""
    cloud = engine.get_cloud(xyz, rgb, q)
    image = engine.get_image(img)
    loss, _ = engine.score_grid(cloud, image, trans, rot)
    return loss.reshape(trans.shape[0], rot.shape[0])


def trim_input_loss(img: torch.Tensor, xyz: torch.Tensor, rgb: torch.Tensor, trans: torch.Tensor, rot: torch.Tensor,
                    num_input: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same contract as the reference: the `num_input` grid poses with the smallest sampling loss,
    ascending; returns (trimmed_trans (k,3), trimmed_rot (k,3)) on img.device."""
    loss_table = score_grid(img, xyz, rgb, trans, rot)
    num_input = min(num_input, loss_table.numel())
    min_inds = engine.topk(loss_table.flatten(), num_input)
    return trans[min_inds // len(rot)], rot[min_inds % len(rot)]


def quantile(x: torch.Tensor, q: float):
    """Order statistics x_sorted[int(N q)], x_sorted[int(N (1-q))] (utils.py:208-229)."""
    with torch.no_grad():
        srt = torch.sort(x).values
        return srt[int(len(x) * q)], srt[int(len(x) * (1 - q))]


def rot_from_ypr(ypr_array: torch.Tensor) -> torch.Tensor:
    """R = Rz(yaw)·Ry(pitch)·Rx(roll) (utils.py:425-453)."""
    from .omniloc import _rotation_from_angles
    yaw, pitch, roll = ypr_array
    return _rotation_from_angles(yaw.reshape(1), pitch.reshape(1), roll.reshape(1), ypr_array.device)


def generate_rot_points(init_dict=None, device="cpu") -> torch.Tensor:
    """Rotation start grid (utils.py:321-360): yaw-only, or the num_yaw x num_pitch x num_roll Euler lattice
    with triples that produce the same rotation removed (24 of 64 survive for 4x4x4).  The reference
    de-duplicates through a python `set` (order depends on PYTHONHASHSEED); here duplicates are
    dropped keeping the FIRST occurrence in lattice order, which fixes the pose index i*R+j."""
    import numpy as np
    if init_dict["yaw_only"]:
        rot_arr = torch.zeros(init_dict["num_yaw"], 3, device=device)
        rot_arr[:, 0] = torch.arange(init_dict["num_yaw"], dtype=torch.float, device=device) * 2 * np.pi / init_dict["num_yaw"]
        return rot_arr
    ny, npi, nr = init_dict["num_yaw"], init_dict["num_pitch"], init_dict["num_roll"]
    y, p, r = torch.meshgrid(torch.arange(ny).float() / ny, torch.arange(npi).float() / npi, torch.arange(nr).float() / nr, indexing="ij")
    rot_arr = torch.stack([y.reshape(-1), p.reshape(-1), r.reshape(-1)], dim=1)
    for k, name in enumerate(("yaw", "pitch", "roll")):
        lo, hi = init_dict.get("min_" + name, 0.0), init_dict.get("max_" + name, 2 * np.pi)
        rot_arr[:, k] = rot_arr[:, k] * (hi - lo) + lo
    seen, keep = set(), []
    for i, ypr in enumerate(rot_arr):
        key = tuple(np.round(rot_from_ypr(ypr).numpy() + 0.0, 3).reshape(-1).tolist())
        key = tuple(0.0 if v == 0 else v for v in key)     # -0.0 == 0.0
        if key not in seen:
            seen.add(key)
            keep.append(i)
    return rot_arr[keep].to(device)


def trim_input_hist_secondary(img: torch.Tensor, xyz: torch.Tensor, rgb: torch.Tensor, trans: torch.Tensor, rot: torch.Tensor,
                              num_input: int, num_split_h: int, num_split_w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same contract as the reference (utils.py:510-588): the `num_input` candidates whose rendered colour
    histograms intersect the query's best, in descending order of intersection."""
    cloud = engine.get_cloud(xyz, rgb)           # the clamp-box quantile plays no part in the re-rank
    poses = torch.cat([trans.reshape(-1, 3), rot.reshape(-1, 3)], dim=1).to(torch.float32)
    scores = engine.hist_rerank(cloud, img, poses, num_split_h, num_split_w)
    order = engine.topk(-scores, min(num_input, scores.numel()))
    return trans[order], rot[order]


def adaptive_trans_num(xyz: torch.Tensor, max_trans_num: int, xy_only: bool = False):
    """Number of translation start points per axis from the 10-90 % extents (utils.py:282-318)."""
    from math import ceil
    ext = (torch.quantile(xyz, dim=0, q=0.90) - torch.quantile(xyz, dim=0, q=0.10)).tolist()
    if xy_only:
        return ceil((ext[0] * max_trans_num / ext[1]) ** (1 / 2)), ceil((ext[1] * max_trans_num / ext[0]) ** (1 / 2))
    n = [ceil((ext[a] ** 2 * max_trans_num / (ext[b] * ext[c])) ** (1 / 3)) for a, b, c in ((0, 1, 2), (1, 0, 2), (2, 0, 1))]
    return tuple(v - 1 if v % 2 == 0 else v for v in n)


def _quantile_any(x: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """torch.quantile refuses inputs above 16M elements; the sort-based equivalent (linear interpolation)."""
    if x.numel() <= (1 << 24):
        return torch.quantile(x, q)
    srt = torch.sort(x).values
    pos = q.to(torch.float64) * (x.numel() - 1)
    lo = pos.floor().long()
    hi = torch.clamp(lo + 1, max=x.numel() - 1)
    frac = (pos - lo).to(x.dtype)
    return srt[lo] + (srt[hi] - srt[lo]) * frac


def generate_trans_points(xyz: torch.Tensor, init_dict=None, device="cpu") -> torch.Tensor:
    """Translation start grid (utils.py:363-422): quantile (default) / uniform / manual lattices; pose order =
    meshgrid 'ij' order flattened, as the reference."""
    mode = init_dict["trans_init_mode"]

    def axis_points(k, n):
        ar = torch.arange(n, device=device)
        if mode == "uniform":
            return (ar + 1) / (n + 1) * (xyz[:, k].max() - xyz[:, k].min()) + xyz[:, k].min()
        if mode == "manual":
            lo, hi = init_dict[("x_min", "y_min", "z_min")[k]], init_dict[("x_max", "y_max", "z_max")[k]]
            return ar / (n - 1) * (hi - lo) + lo
        split = (ar + 1) / (n + 1) if 1 / (n + 1) > 0.1 else torch.linspace(0.1, 0.9, n, device=device)
        return _quantile_any(xyz[:, k], split.to(xyz.dtype))

    if init_dict["xy_only"]:
        if init_dict["dataset"] not in ("Stanford2D-3D-S", "OmniScenes"):
            raise NotImplementedError("Other datasets not supported")
        nx, ny = adaptive_trans_num(xyz, init_dict["num_trans"], xy_only=True)
        gx, gy = torch.meshgrid(axis_points(0, nx), axis_points(1, ny), indexing="ij")
        trans = torch.zeros(nx * ny, 3, device=device)
        trans[:, 0], trans[:, 1] = gx.reshape(-1), gy.reshape(-1)
        trans[:, 2] = init_dict["z_prior"] if init_dict["z_prior"] is not None else xyz[:, 2].mean()
        return trans
    nx, ny, nz = adaptive_trans_num(xyz, init_dict["num_trans"], xy_only=False)
    gx, gy, gz = torch.meshgrid(axis_points(0, nx), axis_points(1, ny), axis_points(2, nz), indexing="ij")
    return torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)


def make_input(img: torch.Tensor, xyz: torch.Tensor, rgb: torch.Tensor, num_input: int, init_dict=None, criterion: str = "histogram",
               num_intermediate=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Start-pose selection (utils.py:591-629): rotation grid x translation grid -> loss scoring (top
    `num_intermediate`) -> colour-histogram re-rank (top `num_input`).  Like the reference, only
    criterion == 'loss_histogram' is implemented (anything else left `input_trans` unbound there)."""
    rot = generate_rot_points(init_dict, device=img.device)
    trans = generate_trans_points(xyz, init_dict, device=img.device)
    if criterion != "loss_histogram":
        raise ValueError("only criterion='loss_histogram' is supported (as in the reference, utils.py:625)")
    if init_dict.get("sample_rate_for_init") is not None:
        raise NotImplementedError("sample_rate_for_init is broken in the reference (utils.py:618-620 subsamples xyz but not rgb)")
    trimmed_trans, trimmed_rot = trim_input_loss(img, xyz, rgb, trans, rot, num_intermediate)
    return trim_input_hist_secondary(img, xyz, rgb, trimmed_trans, trimmed_rot, num_input, init_dict["num_split_h"], init_dict["num_split_w"])


def out_of_room(xyz: torch.Tensor, trans: torch.Tensor, out_quantile: float = 0.05) -> bool:
    """True if `trans` (3,1) lies outside the out_quantile box of the cloud (utils.py:232-254)."""
    with torch.no_grad():
        for k in range(3):
            lo, hi = quantile(xyz[:, k], out_quantile)
            if not (lo < trans[k][0] < hi):
                return True
    return False


def make_pano(xyz: torch.Tensor, rgb: torch.Tensor, resolution=(200, 400), return_torch: bool = False):
    """Painter's-algorithm render of camera-frame points to an equirectangular image (utils.py:134-205), used for
    the result PNGs.  Deterministic version of the reference's nine `index_put_` calls: nearest point wins, the
    centre write beats the 3x3 dilation writes."""
    import numpy as np
    H, W = resolution
    with torch.no_grad():
        q = xyz.detach().to(torch.float32).cpu()
        col = rgb.detach().to(torch.float32).cpu()
        dist = torch.linalg.vector_norm(q, dim=-1)
        order = torch.argsort(dist, descending=True, stable=True)
        q, col = q[order], col[order]
        theta = torch.atan2(torch.linalg.vector_norm(q[:, :2], dim=-1), q[:, 2] + 1e-6)
        phi = torch.atan2(q[:, 1], q[:, 0] + 1e-6) + np.pi
        u, v = 2 * (1.0 - phi / (2 * np.pi)) - 1, 2 * (theta / np.pi) - 1
        x = (((u + 1.0) / 2.0) * (W - 1)).long()
        y = (((v + 1.0) / 2.0) * (H - 1)).long()
        image = torch.zeros(H * W, 3)
        yp, ym, xp, xm = (y + 1).clamp(max=H - 1), (y - 1).clamp(min=0), (x + 1).clamp(max=W - 1), (x - 1).clamp(min=0)
        rank = torch.arange(len(q))
        best = torch.full((H * W,), -1, dtype=torch.long)
        for call, (yy, xx) in enumerate(((y, xm), (y, xp), (ym, xm), (ym, x), (ym, xp), (yp, xm), (yp, x), (yp, xp), (y, x))):
            best.scatter_reduce_(0, yy * W + xx, call * len(q) + rank, reduce="amax")   # later call, nearer point wins
        hit = best >= 0
        image[hit] = col[best[hit] % len(q)]
        image = image.reshape(H, W, 3) * 255
        return image if return_torch else image.numpy().astype(np.uint8)
