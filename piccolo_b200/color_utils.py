"""Per-query colour preprocessing of the reference (`color_utils.py:7-65` color_mod, `:146-234` color_match),
SURVEY §8f next #4.  CUDA tensors are processed on the device (pcl_color.cu: histogram and rewrite passes; the
<= 256-entry table arithmetic stays here and is shared with the numpy + cv2 restatement used for CPU tensors, which is
pinned to the reference's golden outputs); both keep the reference's side effect that outputs are re-quantised
through uint8."""
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


def _equalise_cdf(hist_img: np.ndarray, hist_pts: np.ndarray) -> np.ndarray:
    """cumulative distribution of the joint luma histogram (color_utils.py:38-47): `.float()` each, add, normalise,
    cumsum — all in fp32.  Shared by the CPU and the CUDA path."""
    hist = hist_img.astype(np.float32) + hist_pts.astype(np.float32)
    return np.cumsum(hist / hist.sum(), dtype=np.float32)


def _color_mod_cuda(img: torch.Tensor, rgb: torch.Tensor, num_bins: int):
    """CUDA path (pcl_color.cu): luma-bin histograms on the device, the num_bins-entry cumulative sum here, one
    rewrite pass over pixels and points on the device."""
    from . import _lib
    from .engine import _f32c, _stream
    lib = _lib.load()
    H, W, _ = img.shape
    im, pts = _f32c(img), _f32c(rgb)
    hist = torch.empty(2 * num_bins, dtype=torch.int64, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.pcl_color_mod_stats(im.data_ptr(), H, W, pts.data_ptr(), pts.shape[0], num_bins, hist.data_ptr(), _stream(img.device)))
    h = hist.cpu().numpy()                                          # synchronises
    cdf = torch.from_numpy(_equalise_cdf(h[:num_bins], h[num_bins:])).to(img.device)
    out_img, out_rgb = torch.empty_like(im), torch.empty_like(pts)
    with torch.cuda.device(img.device):
        _lib.check(lib.pcl_color_mod_apply(im.data_ptr(), H, W, pts.data_ptr(), pts.shape[0], num_bins, cdf.data_ptr(), out_img.data_ptr(),
                                           out_rgb.data_ptr(), _stream(img.device)))
    return out_img, out_rgb


def color_mod(img: torch.Tensor, rgb: torch.Tensor, num_bins: int):
    """Joint histogram equalisation of the luma of panorama and cloud (YCrCb), `sharpen_color` of the configs.
    Returns (img (H,W,3), rgb (N,3)) float32 on img.device.  CUDA tensors are processed on the device."""
    device = img.device
    H, W, _ = img.shape
    if img.is_cuda and rgb.is_cuda and 2 <= num_bins <= 4096:
        return _color_mod_cuda(img.detach(), rgb.detach(), int(num_bins))
    flat = img.detach().cpu().numpy().astype(np.float32).reshape(-1, 3).copy()
    lit = _lit_mask(flat)
    ycc_img, ycc_pts = _to_ycc(flat[lit]), _to_ycc(rgb.detach().cpu().numpy().astype(np.float32))
    scale = np.float32(num_bins - 1)
    bin_img, bin_pts = (ycc_img[:, 0] * scale).astype(np.int64), (ycc_pts[:, 0] * scale).astype(np.int64)
    cdf = _equalise_cdf(np.bincount(bin_img, minlength=num_bins), np.bincount(bin_pts, minlength=num_bins))
    ycc_img[:, 0] = cdf[bin_img]
    ycc_pts[:, 0] = cdf[bin_pts]
    flat[lit] = _to_rgb(ycc_img)
    return torch.from_numpy(flat.reshape(H, W, 3)).to(device), torch.from_numpy(_to_rgb(ycc_pts)).to(rgb.device)


def _interp_levels(src_counts: np.ndarray, tmp_values: np.ndarray, tmp_counts: np.ndarray, n_template: int) -> np.ndarray:
    """The <= 256-entry core of `_match_cumulative_cdf` + `_interp` (color_utils.py:159-199): src_counts[level] is the
    weighted histogram of the source by truncated level, (tmp_values, tmp_counts) the sorted unique template values
    and their counts.  Returns the interpolated value per source LEVEL.  Shared by the CPU and the CUDA path."""
    src_q = np.cumsum(src_counts.astype(np.float32), dtype=np.float32)
    src_q = src_q / src_q[-1]
    tmp_q = (np.cumsum(tmp_counts) / np.float32(n_template)).astype(np.float32)
    # periodic extension with period 360 (sentinels far outside [0,1]), then piecewise-linear interpolation
    order = np.argsort(tmp_q, kind="stable")
    xp, fp = tmp_q[order], tmp_values[order]
    xp = np.concatenate([xp[-1:] - np.float32(360), xp, xp[:1] + np.float32(360)]).astype(np.float32)
    fp = np.concatenate([fp[-1:], fp, fp[:1]]).astype(np.float32)
    big = len(xp) - (src_q[:, None] < xp[None, :]).sum(1)
    small = big - 1
    out = ((src_q - xp[small]) * fp[big] + (xp[big] - src_q) * fp[small]) / (xp[big] - xp[small])
    return out.astype(np.float32)


def _match_channel(source: np.ndarray, template: np.ndarray, weight: np.ndarray) -> np.ndarray:
    """CDF matching of one channel (`_match_cumulative_cdf` + `_interp`), quirks kept: the source histogram is
    indexed by truncated level `(source*255).int()`, the result is looked up by unique-value rank."""
    _, inverse = np.unique(source, return_inverse=True)
    tmp_values, tmp_counts = np.unique(template, return_counts=True)
    levels = (source * np.float32(255.0)).astype(np.int32)
    src_counts = np.bincount(levels, weights=weight.astype(np.float64)).astype(np.float32)
    return _interp_levels(src_counts, tmp_values, tmp_counts, len(template))[inverse].reshape(source.shape)


def _row_weight(H: int) -> np.ndarray:
    """sin(row / H * pi) in fp32 (color_utils.py:216-217)"""
    return np.sin(np.arange(H, dtype=np.float32) / np.float32(H) * np.float32(np.pi)).astype(np.float32)


def _color_match_cuda(img: torch.Tensor, rgb: torch.Tensor):
    """CUDA path (pcl_color.cu): three 256-bin histograms per side on the device, the <= 256-entry interpolation
    here, one rewrite pass on the device.  Returns None when an input is not exactly uint8/255 data."""
    import ctypes
    from . import _lib
    from .engine import _f32c, _stream
    lib = _lib.load()
    H, W, _ = img.shape
    im, pts = _f32c(img), _f32c(rgb)
    stats = torch.empty(768 * 24 + 16, dtype=torch.uint8, device=img.device)
    roww = torch.from_numpy(_row_weight(H)).to(img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.pcl_color_stats(im.data_ptr(), H, W, roww.data_ptr(), pts.data_ptr(), pts.shape[0], stats.data_ptr(), _stream(img.device)))
    raw = stats.cpu().numpy()                                       # 18 KB, synchronises
    whist = raw[: 768 * 8].view(np.float64).reshape(3, 256)
    level_cnt = raw[768 * 8: 768 * 12].view(np.uint32).reshape(3, 256)
    value_cnt = raw[768 * 12: 768 * 16].view(np.uint32).reshape(3, 256)
    cloud_cnt = raw[768 * 16: 768 * 24].view(np.uint64).reshape(3, 256)
    if raw[768 * 24: 768 * 24 + 4].view(np.int32)[0] != 0:
        return None
    lut = np.zeros((3, 256), np.float32)
    unit = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)      # the value of id k, as `uint8.float() / 255.`
    for c in range(3):
        if level_cnt[c].sum() == 0:
            continue                                                                   # no lit pixel: nothing is rewritten
        n_levels = int(np.nonzero(level_cnt[c])[0].max()) + 1                          # bincount length = max level + 1
        present = np.nonzero(cloud_cnt[c])[0]
        out = _interp_levels(whist[c, :n_levels].astype(np.float32), unit[present], cloud_cnt[c, present].astype(np.int64), int(pts.shape[0]))
        ks = np.nonzero(value_cnt[c])[0]                                               # sorted unique source values -> rank
        lut[c, ks] = out[np.minimum(np.arange(len(ks)), len(out) - 1)]                 # looked up by unique-value RANK (the reference's quirk)
    out_img = torch.empty_like(im)
    lut_d = torch.from_numpy(lut).to(img.device)
    with torch.cuda.device(img.device):
        _lib.check(lib.pcl_color_apply(im.data_ptr(), H, W, lut_d.data_ptr(), out_img.data_ptr(), _stream(img.device)))
    return out_img


def requantize(img: torch.Tensor) -> torch.Tensor:
    """The uint8 round trip the drivers apply after colour preprocessing, `(255 * img).astype(np.uint8)` followed by
    `torch.from_numpy(..).float() / 255.` (localize.py:404, :413-414), without leaving the device.  The division is
    taken from a table computed on the host: torch's CUDA division by a scalar multiplies by the reciprocal, which is
    not the correctly rounded k/255 the reference (and the uint8 texel formats) see."""
    table = (torch.arange(256, dtype=torch.float32) / 255.).to(img.device)
    return table[(img * 255).to(torch.uint8).long()]


def color_match(img: torch.Tensor, rgb: torch.Tensor) -> torch.Tensor:
    """Match the panorama's per-channel colour distribution (rows weighted by sin(latitude)) to the cloud's,
    `match_color` of the configs.  Returns img (H,W,3) float32 on img.device."""
    device = img.device
    H, W, _ = img.shape
    if img.is_cuda and rgb.is_cuda:
        done = _color_match_cuda(img.detach(), rgb.detach())
        if done is not None:
            return done
    weight = np.repeat(_row_weight(H), W)
    flat = img.detach().cpu().numpy().astype(np.float32).reshape(-1, 3).copy()
    lit = _lit_mask(flat)
    pts = rgb.detach().cpu().numpy().astype(np.float32)
    src = flat[lit]
    matched = np.stack([_match_channel(src[:, c], pts[:, c], weight[lit]) for c in range(3)], axis=1)
    flat[lit] = matched
    return torch.from_numpy(flat.reshape(H, W, 3)).to(device)
