"""Dataset drivers with the reference's contract (`localize.py:76-297` Stanford, `:300-536` OmniScenes): same
config keys and defaults, same CSV header/rows, TensorBoard tags, result PNGs and stdout, with the pose search
running on the CUDA path.  Queries come from `piccolo_b200/datasets.py` (synthetic rooms; the datasets are not
available offline)."""
from __future__ import annotations

import csv
import os
import random
import time
from collections import defaultdict

import cv2
import numpy as np
import torch

from . import datasets
from .color_utils import color_match, color_mod
from .omniloc import omniloc_all, omniloc_batch
from .utils import make_input, make_pano, out_of_room


def get_init_dict(cfg):
    """Initialisation settings with the reference's defaults (localize.py:18-73)."""
    two_pi = 2 * np.pi
    spec = [("xy_only", True), ("num_trans", 50), ("yaw_only", True), ("num_yaw", 4), ("num_pitch", 0), ("num_roll", 0),
            ("max_yaw", two_pi), ("min_yaw", 0), ("max_pitch", two_pi), ("min_pitch", 0), ("max_roll", two_pi), ("min_roll", 0),
            ("z_prior", None), ("sample_rate_for_init", None), ("trans_init_mode", "quantile"),
            ("x_max", None), ("x_min", None), ("y_max", None), ("y_min", None), ("z_max", None), ("z_min", None),
            ("num_split_h", 2), ("num_split_w", 4)]
    init = {k: getattr(cfg, k, d) for k, d in spec}
    init["dataset"] = cfg.dataset
    return init


def _flat(a) -> str:
    return str(np.asarray(a).flatten())[1:-1].replace("\n", "")


def write_summaries(writer, scalar_summaries, step):
    """utils.py:455-459: the mean of everything accumulated so far (the reference's reset is a no-op)."""
    for k, v in scalar_summaries.items():
        writer.add_scalar(k, np.array(v).mean().item(), step)


SPECS = {
    "Stanford2D-3D-S": dict(csv="stanford_results.csv", area_column=True, t_thr=0.2, r_thr=float(np.rad2deg(0.2))),   # localize.py:250
    "OmniScenes": dict(csv="omniscenes_results.csv", area_column=False, t_thr=0.1, r_thr=5.0),                        # localize.py:513
}


def _localize(cfg, writer, log_dir: str, dataset: str):
    spec = SPECS[dataset]
    out_q = getattr(cfg, "out_of_room_quantile", 0.05)
    eval_full = getattr(cfg, "eval_full", False)
    parallel = getattr(cfg, "parallel", False)
    scalar_summaries = defaultdict(list)
    torch.manual_seed(2); torch.cuda.manual_seed(2); np.random.seed(2); random.seed(2)      # localize.py:95-98
    if not torch.cuda.is_available():
        raise RuntimeError("piccolo_b200 drivers need a CUDA device (there is no CPU path)")
    device = torch.device("cuda:0")

    well_posed = total_img = 0
    accuracy = 0.0
    failed, skipped = [], []
    summary = open(os.path.join(log_dir, spec["csv"]), "w", encoding="utf-8", newline="")
    out = csv.writer(summary)
    header = ["pano_name", "gt_trans", "gt_rot", "skipped?", "OmniLoc_trans", "OmniLoc_rot", "t_error (m)", "r_error (degrees)", "time (s)"]
    out.writerow((["area_num"] if spec["area_column"] else []) + header)

    if dataset == "Stanford2D-3D-S":
        init_ds = (getattr(cfg, "init_downsample_h", 1), getattr(cfg, "init_downsample_w", 1))
    else:   # localize.py:348-349: "match resolution with stanford"
        init_ds = (max(getattr(cfg, "init_downsample_h", 1) // 2, 1), max(getattr(cfg, "init_downsample_w", 1) // 2, 1))
    main_ds = (getattr(cfg, "main_downsample_h", 1), getattr(cfg, "main_downsample_w", 1))

    past_pcd = ""
    query_source = datasets.queries(cfg, dataset)         # raises on dataset-selecting keys / dataset files it cannot honour
    print(f"[piccolo_b200] {dataset}: SYNTHETIC rooms (the dataset file readers are outside this build's scope); records are named synthetic://...", flush=True)
    for trial, q in enumerate(query_source):
        if past_pcd != q.pcd_name:
            xyz = torch.from_numpy(q.xyz_np).float().to(device)
            rgb = torch.from_numpy(q.rgb_np).float().to(device)
            raw_rgb = rgb.clone().detach()
            past_pcd = q.pcd_name
        orig_img = q.orig_img
        if dataset == "OmniScenes":
            orig_img = cv2.resize(orig_img, (2048, 1024))                                   # localize.py:381
            if getattr(cfg, "synth_const", None) is not None:
                orig_img = orig_img // cfg.synth_const
            if getattr(cfg, "synth_gamma", None) is not None:
                orig_img = (((orig_img / 255.) ** cfg.synth_gamma) * 255).astype(np.uint8)
            if getattr(cfg, "synth_wb", None):
                for c, key in enumerate(("synth_r", "synth_g", "synth_b")):
                    # clamped at 255: the reference casts first and clamps afterwards (localize.py:389-393), so its uint8 cast
                    # WRAPS values above 255 and the clamp is a no-op; the clamp it evidently intends is what is done here
                    orig_img[..., c] = np.minimum(((orig_img[..., c] / 255.) * getattr(cfg, key)) * 255, 255).astype(np.uint8)
        raw_img = torch.from_numpy(orig_img).float().to(device) / 255.
        num_bins = getattr(cfg, "num_bins", 256)
        if dataset == "OmniScenes":
            if getattr(cfg, "match_color", False):
                orig_img = (255 * color_match(raw_img, rgb).cpu().numpy()).astype(np.uint8)
            if getattr(cfg, "sharpen_color", False):
                new_img, rgb = color_mod(raw_img, raw_rgb.clone(), num_bins)
                orig_img = (255 * new_img.cpu().numpy()).astype(np.uint8)
        img = cv2.resize(orig_img, (orig_img.shape[1] // init_ds[1], orig_img.shape[0] // init_ds[0]))
        img = (torch.from_numpy(img).float() / 255.).to(device)
        if dataset == "Stanford2D-3D-S" and getattr(cfg, "sharpen_color", False):
            img, rgb = color_mod(img, raw_rgb.clone(), num_bins)                             # localize.py:173-179

        gt_trans, gt_rot = torch.from_numpy(q.gt_trans).float(), torch.from_numpy(q.gt_rot).float()
        if out_of_room(xyz.cpu(), gt_trans, out_q) and not (eval_full and dataset == "Stanford2D-3D-S"):
            print("corrupted file : {}, gt_trans is out of the room\n".format(q.filename))
            skipped.append(q.filename)
            writer.add_text("skipped rooms", q.filename)
            out.writerow(([q.area_num] if spec["area_column"] else []) + [q.img_name, _flat(gt_trans.numpy()), _flat(gt_rot.numpy()), 1])
            continue

        num_input = getattr(cfg, "num_input", 6)
        num_intermediate = getattr(cfg, "num_intermediate", 20)
        criterion = getattr(cfg, "criterion", "histogram")
        start = time.time()
        input_trans, input_rot = make_input(img, xyz, rgb, num_input, get_init_dict(cfg), criterion, num_intermediate)
        img = cv2.resize(orig_img, (orig_img.shape[1] // main_ds[1], orig_img.shape[0] // main_ds[0]))
        img = (torch.from_numpy(img).float() / 255.).to(device)
        if parallel:
            result = [omniloc_batch(img, xyz, rgb, input_trans, input_rot, cfg, scalar_summaries)]
        else:   # the per-candidate `omniloc` loop of localize.py:219-220 as one batch (identical trajectories)
            result = omniloc_all(img, xyz, rgb, input_trans, input_rot, cfg, scalar_summaries)
        time_spent = time.time() - start

        losses = np.array([float(r[2]) for r in result])
        min_ind = int(np.nanargmin(losses)) if not np.all(np.isnan(losses)) else 0
        t, r = result[min_ind][0], result[min_ind][1]
        gt_t, gt_r = gt_trans.numpy(), gt_rot.numpy()
        print("\n" + q.img_name)
        print("min_index : {}".format(min_ind))
        print("min loss : {}".format(result[min_ind][2]))
        t_error = np.linalg.norm(gt_t - np.array(t))
        print("translation error : {}".format(t_error))
        tr = np.trace(np.matmul(np.transpose(r.numpy()), gt_r))
        tr = -2 - tr if tr < -1 else (6 - tr if tr > 3 else tr)                              # localize.py:242-247
        r_error = np.rad2deg(np.abs(np.arccos((tr - 1) / 2)))
        print("rotation error : {}\n".format(r_error))
        if t_error < spec["t_thr"] and r_error < spec["r_thr"]:
            well_posed += 1
        else:
            failed.append(q.filename)
            writer.add_text("failed rooms", q.filename)
        total_img += 1
        accuracy = well_posed / total_img
        scalar_summaries["current_accuracy"] += [accuracy]
        print("current accuracy : {} ({}/{})\n".format(accuracy, well_posed, total_img))
        out.writerow(([q.area_num] if spec["area_column"] else []) +
                     [q.img_name, _flat(gt_t), _flat(gt_r), 0, _flat(t.numpy()), _flat(r.numpy()), t_error, r_error, time_spent])

        # result PNG: query on top, cloud rendered from the found pose below (localize.py:266-279)
        new_xyz = torch.transpose(torch.matmul(r, torch.transpose(xyz.cpu(), 0, 1) - t), 0, 1)
        render = cv2.cvtColor(make_pano(new_xyz, raw_rgb.cpu(), resolution=(img.shape[0] // 2, img.shape[1] // 2)), cv2.COLOR_RGB2BGR)
        top = cv2.cvtColor(cv2.resize((raw_img.cpu().numpy() * 255).astype(np.uint8), (render.shape[1], render.shape[0])), cv2.COLOR_RGB2BGR)
        sub = "area_{}".format(q.area_num) if spec["area_column"] else os.path.dirname(q.img_name)
        save_dir = os.path.join(log_dir, "results", sub)
        os.makedirs(save_dir, exist_ok=True)
        cv2.imwrite(os.path.join(save_dir, os.path.basename(q.img_name).rsplit(".", 1)[0] + ".png"), cv2.vconcat([top, render]))
        write_summaries(writer, scalar_summaries, trial)

    summary.close()
    writer.add_scalar("final accuracy", accuracy)
    print(f"Final Accuracy : {accuracy}")
    print("failed {} rooms : {}\n".format(len(failed), failed))
    print("skipped {} rooms : {}".format(len(skipped), skipped))
    return accuracy


def localize_stanford(cfg, writer, log_dir: str):
    return _localize(cfg, writer, log_dir, "Stanford2D-3D-S")


def localize_omniscenes(cfg, writer, log_dir: str):
    return _localize(cfg, writer, log_dir, "OmniScenes")
