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


def score_grid(img, xyz, rgb, trans, rot, q: float = 0.05) -> torch.Tensor:
    """loss_table (T,R) of utils.py:481-499 from one kernel launch."""
    cloud = engine.get_cloud(xyz, rgb, q)
    image = engine.get_image(img)
    loss, _ = engine.score(cloud, image, grid_poses(trans, rot))
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
