"""Per-query colour preprocessing of the reference (`color_utils.py:7-65` color_mod, `:146-234` color_match),
restated with numpy + cv2.  Host-side, once per query, not on the sampling-loss path;
both keep the reference's side effect that outputs are re-quantised through uint8."""
from __future__ import annotations

import cv2
import numpy as np
import torch


def _lit_mask(flat: np.ndarray) -> np.ndarray:
    """pixels whose truncated 8-bit channels do not sum to zero (`(img*255).long().sum(-1) > 0`)."""
    return (flat * np.float32(255.0)).astype(np.int64).sum(-1) > 0


def _to_ycc(unit_rgb: np.ndarray) -> np.ndarray:
    u8 = (unit_rgb * np.float32(255.0)).astype(np.uint8).reshape(1, -1, 3)
    return cv2.cvtColor(u8, cv2.COLOR_RGB2YCR_CB).reshape(-1, 3).astype(np.float32) / np.float32(255.0)


def _to_rgb(unit_ycc: np.ndarray) -> np.ndarray:
    u8 = (unit_ycc * np.float32(255.0)).astype(np.uint8).reshape(1, -1, 3)
    return cv2.cvtColor(u8, cv2.COLOR_YCR_CB2RGB).reshape(-1, 3).astype(np.float32) / np.float32(255.0)


def color_mod(img: torch.Tensor, rgb: torch.Tensor, num_bins: int):
    """Joint histogram equalisation of the luma of panorama and cloud (YCrCb), `sharpen_color` of the configs.
    Returns (img (H,W,3), rgb (N,3)) float32 on img.device."""
    device = img.device
    H, W, _ = img.shape
    flat = img.detach().cpu().numpy().astype(np.float32).reshape(-1, 3).copy()
    lit = _lit_mask(flat)
    ycc_img, ycc_pts = _to_ycc(flat[lit]), _to_ycc(rgb.detach().cpu().numpy().astype(np.float32))
    scale = np.float32(num_bins - 1)
    bin_img, bin_pts = (ycc_img[:, 0] * scale).astype(np.int64), (ycc_pts[:, 0] * scale).astype(np.int64)
    hist = (np.bincount(bin_img, minlength=num_bins) + np.bincount(bin_pts, minlength=num_bins)).astype(np.float32)
    cdf = np.cumsum(hist / hist.sum(), dtype=np.float32)
    ycc_img[:, 0] = cdf[bin_img]
    ycc_pts[:, 0] = cdf[bin_pts]
    flat[lit] = _to_rgb(ycc_img)
    return torch.from_numpy(flat.reshape(H, W, 3)).to(device), torch.from_numpy(_to_rgb(ycc_pts)).to(rgb.device)


def _match_channel(source: np.ndarray, template: np.ndarray, weight: np.ndarray) -> np.ndarray:
    """CDF matching of one channel (`_match_cumulative_cdf` + `_interp`), quirks kept: the source histogram is
    indexed by truncated level `(source*255).int()`, the result is looked up by unique-value rank."""
    _, inverse = np.unique(source, return_inverse=True)
    tmp_values, tmp_counts = np.unique(template, return_counts=True)
    levels = (source * np.float32(255.0)).astype(np.int32)
    src_counts = np.bincount(levels, weights=weight.astype(np.float64)).astype(np.float32)
    src_q = np.cumsum(src_counts, dtype=np.float32)
    src_q = src_q / src_q[-1]
    tmp_q = (np.cumsum(tmp_counts) / np.float32(len(template))).astype(np.float32)
    # periodic extension with period 360 (sentinels far outside [0,1]), then piecewise-linear interpolation
    order = np.argsort(tmp_q, kind="stable")
    xp, fp = tmp_q[order], tmp_values[order]
    xp = np.concatenate([xp[-1:] - np.float32(360), xp, xp[:1] + np.float32(360)]).astype(np.float32)
    fp = np.concatenate([fp[-1:], fp, fp[:1]]).astype(np.float32)
    big = len(xp) - (src_q[:, None] < xp[None, :]).sum(1)
    small = big - 1
    out = ((src_q - xp[small]) * fp[big] + (xp[big] - src_q) * fp[small]) / (xp[big] - xp[small])
    return out.astype(np.float32)[inverse].reshape(source.shape)


def color_match(img: torch.Tensor, rgb: torch.Tensor) -> torch.Tensor:
    """Match the panorama's per-channel colour distribution (rows weighted by sin(latitude)) to the cloud's,
    `match_color` of the configs.  Returns img (H,W,3) float32 on img.device."""
    device = img.device
    H, W, _ = img.shape
    rows = np.repeat(np.arange(H, dtype=np.float32), W)
    weight = np.sin(rows / np.float32(H) * np.float32(np.pi)).astype(np.float32)
    flat = img.detach().cpu().numpy().astype(np.float32).reshape(-1, 3).copy()
    lit = _lit_mask(flat)
    pts = rgb.detach().cpu().numpy().astype(np.float32)
    src = flat[lit]
    matched = np.stack([_match_channel(src[:, c], pts[:, c], weight[lit]) for c in range(3)], axis=1)
    flat[lit] = matched
    return torch.from_numpy(flat.reshape(H, W, 3)).to(device)
