"""Query provider for the drivers.  The reference reads Stanford2D-3D-S / OmniScenes files from ./data
(`data_utils.py:16-182`, `localize.py:103-120`, `:326-335`); those datasets are not available offline, so the
drivers run on seeded synthetic textured rooms (`piccolo_b200/synth.py`) presented through the same record
shape: room cloud (xyz, rgb in [0,1]), original-resolution uint8 panorama, ground-truth translation (3,1) and
rotation (3,3).  Config keys (all optional): synthetic, synthetic_rooms, synthetic_queries, synthetic_points,
synthetic_height.

The dataset FILE READERS (`read_stanford`, `read_omniscenes`, `obtain_gt_*`, `obtain_align_matrix`) are outside the
scope of this build (drivers run "on the synthetic dataset provider"), and the provider never pretends
otherwise: synthetic records carry `synthetic://` names (in the CSV, the log and the result PNG paths), a run that
finds dataset files under ./data refuses to start unless the config says `synthetic: True`, and config keys that
select dataset files or need code that is not there (`area`, `room_name`, `scene_number`, `split_name`,
`gravity_aligned: False`, `visualize: True`) raise instead of being ignored."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator

import numpy as np

from . import synth


@dataclass
class Query:
    area_num: int
    img_name: str          # Stanford: camera_<id>_<roomtype>_<roomno>_frame_equirectangular_domain_rgb.png
    filename: str
    room_type: str
    pcd_name: str
    xyz_np: np.ndarray
    rgb_np: np.ndarray
    orig_img: np.ndarray   # (H,W,3) uint8 RGB
    gt_trans: np.ndarray   # (3,1)
    gt_rot: np.ndarray     # (3,3)


DATA_DIRS = {"Stanford2D-3D-S": "./data/stanford", "OmniScenes": "./data/omniscenes"}
# key -> the value(s) under which the key changes nothing (reference defaults)
_DATASET_ONLY_KEYS = {"area": (None,), "room_name": (None,), "scene_number": (None,), "split_name": (None, "extreme"),
                      "gravity_aligned": (None, True), "visualize": (None, False)}


class DatasetUnavailable(RuntimeError):
    pass


def check_config(cfg, dataset: str) -> None:
    """Raise on config keys this build would otherwise silently ignore, and on dataset files it cannot read."""
    bad = [f"{k}: {getattr(cfg, k)!r}" for k, ok in _DATASET_ONLY_KEYS.items() if getattr(cfg, k, None) not in ok]
    if bad:
        raise DatasetUnavailable("config keys that select dataset files or need code outside this build's scope "
                                 "(data_utils.py file readers, gravity alignment, GIF frames): " + ", ".join(bad) +
                                 " — remove them (or set them to the reference defaults) to run on synthetic rooms")
    import os
    d = DATA_DIRS.get(dataset)
    if d and os.path.isdir(d) and not bool(getattr(cfg, "synthetic", False)):
        raise DatasetUnavailable(f"{d} exists, but the dataset file readers (data_utils.py:16-182) are not part of this build; "
                                 "set `synthetic: True` in the config to run on synthetic rooms instead")


def queries(cfg, dataset: str) -> Iterator[Query]:
    """The drivers' query source: validates the config, then yields synthetic records (marked as such)."""
    check_config(cfg, dataset)
    return synthetic_queries(cfg, dataset)


def synthetic_queries(cfg, dataset: str) -> Iterator[Query]:
    n_rooms = getattr(cfg, "synthetic_rooms", 1)
    n_queries = getattr(cfg, "synthetic_queries", 2)
    n_points = getattr(cfg, "synthetic_points", 200_000)
    height = getattr(cfg, "synthetic_height", 512 if dataset == "Stanford2D-3D-S" else 1024)
    sample_rate = getattr(cfg, "sample_rate", 1)
    yaw_only = bool(getattr(cfg, "yaw_only", False))
    rng = np.random.default_rng(2)
    for room in range(n_rooms):
        room_dims = (8.0 + 1.5 * room, 6.0 + 1.0 * room, 3.0)
        xyz, rgb8 = synth.sample_room_points(n_points, room_dims, seed=2 + room)
        if sample_rate > 1:                                   # random subsample like data_utils.py:36-41
            idx = rng.permutation(len(xyz))[: int(len(xyz) / sample_rate)]
            xyz, rgb8 = xyz[idx], rgb8[idx]
        rgb = rgb8.astype(np.float64) / 255.0
        for q in range(n_queries):
            gt = synth.random_gt_pose(room_dims, seed=3 + 17 * room + q, yaw_only=yaw_only)
            z_prior = getattr(cfg, "z_prior", None)
            if yaw_only and z_prior is not None:
                gt[2] = float(z_prior)
            img8 = synth.render_panorama(gt, height, 2 * height, room_dims)
            room_type, room_no = "synthroom", str(room + 1)
            if dataset == "Stanford2D-3D-S":
                img_name = f"camera_{q:04d}_{room_type}_{room_no}_frame_equirectangular_domain_rgb.png"
                filename = f"synthetic://stanford/pano/area_1/{img_name}"
                pcd_name = f"synthetic://stanford/pcd_not_aligned/area_1/{room_type}_{room_no}.txt"
            else:
                img_name = f"{room_type}_{room_no}/{q:06d}.jpg"
                filename = f"synthetic://omniscenes/synthetic_pano/{img_name}"
                pcd_name = f"synthetic://omniscenes/pcd/{room_type}_{room_no}.txt"
            yield Query(1, img_name, filename, room_type, pcd_name, xyz.astype(np.float64), rgb, img8,
                        gt[:3].reshape(3, 1).copy(), synth.rot_zyx(*gt[3:]))
