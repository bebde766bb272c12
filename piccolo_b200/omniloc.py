"""Drop-in for the sampling-loss call surface of the reference's `omniloc.py`.

Same names, signatures, tensor shapes and return conventions as
`SamplingLoss` (omniloc.py:160-202), `BatchSamplingLoss` (:299-356), `omniloc` (:11-102),
`omniloc_batch` (:205-296) and `sampling_loss` (:105-157); the torch op chain behind them is
replaced by the CUDA kernels of libpiccolo_b200.so.  CUDA tensors only — there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import cos, sin

from . import engine


class _SamplingLossFn(torch.autograd.Function):
    """loss_b = sampling loss of pose b; one fused launch computes the loss AND its analytic gradient,
    so backward only scales the saved gradient."""

    @staticmethod
    def forward(ctx, poses, cloud, image):
        loss, count, grad = engine.loss_fwd_bwd(cloud, image, poses)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(count)
        return loss, count

    @staticmethod
    def backward(ctx, g_loss, _g_count):
        (grad,) = ctx.saved_tensors
        return g_loss.reshape(-1, 1) * grad, None, None


def sampling_loss_poses(cloud: engine.Cloud, image: engine.Image, poses: torch.Tensor):
    """Differentiable per-pose loss for poses (B,6).  Returns (loss (B,), count (B,))."""
    return _SamplingLossFn.apply(poses, cloud, image)


class SamplingLoss(nn.Module):
    """Same contract as the reference module: ctor (xyz, rgb, img, device, cfg);
    forward(translation (3,1), yaw (1,), pitch (1,), roll (1,)) -> 0-dim loss, differentiable w.r.t. the
    four pose tensors."""

    def __init__(self, xyz: torch.Tensor, rgb: torch.Tensor, img: torch.Tensor, device: torch.device, cfg):
        super().__init__()
        self.xyz, self.rgb, self.img, self.cfg = xyz, rgb, img, cfg
        q = getattr(cfg, "out_of_room_quantile", 0.05)
        self.cloud = engine.get_cloud(xyz, rgb, q)
        self.image = engine.get_image(img)

    def forward(self, translation, yaw, pitch, roll):
        pose = torch.cat([translation.reshape(3), yaw.reshape(1), pitch.reshape(1), roll.reshape(1)]).reshape(1, 6)
        loss, _ = sampling_loss_poses(self.cloud, self.image, pose)
        return loss[0]


class BatchSamplingLoss(nn.Module):
    """forward(translation (B,3,1), yaw (B,1), pitch (B,1), roll (B,1)) -> (Σ_b loss_b, loss_list (B,))."""

    def __init__(self, xyz: torch.Tensor, rgb: torch.Tensor, img: torch.Tensor, device: torch.device, cfg):
        super().__init__()
        self.xyz, self.rgb, self.img, self.cfg = xyz, rgb, img, cfg
        self.num_input = cfg.num_input
        q = getattr(cfg, "out_of_room_quantile", 0.05)
        self.cloud = engine.get_cloud(xyz, rgb, q)
        self.image = engine.get_image(img)

    def forward(self, translation, yaw, pitch, roll):
        B = translation.shape[0]
        pose = torch.cat([translation.reshape(B, 3), yaw.reshape(B, 1), pitch.reshape(B, 1), roll.reshape(B, 1)], dim=1)
        loss_list, _ = sampling_loss_poses(self.cloud, self.image, pose)
        return loss_list.sum(), loss_list


def _rotation_from_angles(yaw, pitch, roll, device):
    """R = Rz(yaw)·Ry(pitch)·Rx(roll) of the final angles (what omniloc.py:71-87 rebuilds), in closed form:
        [ cy·cp   cy·sp·sr − sy·cr   cy·sp·cr + sy·sr ]
        [ sy·cp   sy·sp·sr + cy·cr   sy·sp·cr − cy·sr ]
        [ −sp     cp·sr              cp·cr            ]"""
    y, p, r = [a.reshape(()).to(device=device, dtype=torch.float32) for a in (yaw, pitch, roll)]
    cy, sy, cp, sp, cr, sr = cos(y), sin(y), cos(p), sin(p), cos(r), sin(r)
    rows = [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr]]
    return torch.stack([torch.stack(row) for row in rows])


def _cfg_train(cfg):
    return (getattr(cfg, "lr", 0.1), getattr(cfg, "num_iter", 100), getattr(cfg, "patience", 5),
            getattr(cfg, "factor", 0.9), getattr(cfg, "out_of_room_quantile", 0.05))


def refine_candidates(img, xyz, rgb, input_trans, input_rot, cfg, batch_semantics: bool):
    """All candidates refined concurrently, one fused launch per iteration.  Candidate trajectories are
    independent, so with batch_semantics=False this equals running `omniloc` per candidate.
    Returns dict(pose (B,6), param (B,6), loss (B,)) on the device."""
    lr, num_iter, patience, factor, q = _cfg_train(cfg)
    cloud = engine.get_cloud(xyz, rgb, q)
    image = engine.get_image(img)
    poses0 = torch.cat([input_trans.reshape(-1, 3), input_rot.reshape(-1, 3)], dim=1).to(torch.float32)
    ref = engine.Refiner(poses0.shape[0], lr=lr, factor=factor, patience=patience, batch_semantics=batch_semantics)
    ref.reset(poses0).run(cloud, image, num_iter)
    return ref.read()


def omniloc(img, xyz, rgb, input_trans, input_rot, starting_point, cfg, scalar_summaries):
    """Sequential refinement of candidate `starting_point` (omniloc.py:11-102).
    Returns [translation (3,1) cpu, R (3,3) cpu, loss 0-dim cpu]."""
    out = refine_candidates(img, xyz, rgb, input_trans[starting_point:starting_point + 1],
                            input_rot[starting_point:starting_point + 1], cfg, batch_semantics=False)
    pose = out["pose"][0]
    R = _rotation_from_angles(pose[3:4], pose[4:5], pose[5:6], pose.device)
    return [pose[:3].reshape(3, 1).cpu(), R.cpu(), out["loss"][0].cpu()]


def omniloc_all(img, xyz, rgb, input_trans, input_rot, cfg, scalar_summaries=None):
    """Extension: the `for i in range(num_input): omniloc(...)` loop of localize.py:219-220 as ONE batch.
    Returns the list of [translation, R, loss] the loop would have produced."""
    out = refine_candidates(img, xyz, rgb, input_trans, input_rot, cfg, batch_semantics=False)
    res = []
    for b in range(out["pose"].shape[0]):
        pose = out["pose"][b]
        R = _rotation_from_angles(pose[3:4], pose[4:5], pose[5:6], pose.device)
        res.append([pose[:3].reshape(3, 1).cpu(), R.cpu(), out["loss"][b].cpu()])
    return res


def omniloc_batch(img, xyz, rgb, input_trans, input_rot, cfg, scalar_summaries):
    """All candidates jointly with the reference's batch semantics, arg-min inside (omniloc.py:205-296).
    Returns [translation (3,1) cpu, R (3,3) cpu, loss cpu] of the best candidate."""
    assert cfg.num_input > 1
    out = refine_candidates(img, xyz, rgb, input_trans, input_rot, cfg, batch_semantics=True)
    min_idx = out["loss"].argmin().item()
    pose = out["pose"][min_idx]
    R = _rotation_from_angles(pose[3:4], pose[4:5], pose[5:6], pose.device)
    return [pose[:3].reshape(3, 1).cpu(), R.cpu(), out["loss"][min_idx].cpu()]


def sampling_loss(img, xyz, rgb, input_trans, input_rot, starting_point, cfg, return_list=True):
    """Forward-only loss of one candidate (omniloc.py:105-157)."""
    q = getattr(cfg, "out_of_room_quantile", 0.05)
    cloud = engine.get_cloud(xyz, rgb, q)
    image = engine.get_image(img)
    pose = torch.cat([input_trans[starting_point].reshape(3), input_rot[starting_point].reshape(3)]).reshape(1, 6).to(torch.float32)
    loss, _ = engine.score(cloud, image, pose)
    if not return_list:
        return loss[0].cpu()
    R = _rotation_from_angles(pose[0, 3:4], pose[0, 4:5], pose[0, 5:6], pose.device)
    return [pose[0, :3].reshape(3, 1).cpu(), R.cpu(), loss[0].cpu()]
