"""Per-query colour preprocessing of the reference (`color_utils.py:7-65` color_mod, `:146-234` color_match),
SURVEY §8f next #4, on the device (pcl_color.cu: histogram and rewrite passes); only the <= 256-entry table arithmetic
between the two passes runs here on the host.  There is no CPU path: CPU tensors and panoramas that are not uint8/255 data
raise.  (The numpy + cv2 restatement that pins these steps to the reference's golden outputs is test infrastructure and lives
with the other CPU restatements, outside this package.)  Both steps keep the reference's side effect that outputs are re-quantised through uint8."""
from __future__ import annotations

import numpy as np
import torch


def _equalise_cdf(hist_img: np.ndarray, hist_pts: np.ndarray) -> np.ndarray:
    """cumulative distribution of the joint luma histogram (color_utils.py:38-47): `.float()` each, add, normalise,
    cumsum — all in fp32."""
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



def _require_cuda_pair(img: torch.Tensor, rgb: torch.Tensor, what: str):
    from ._lib import PiccoloError
    if not (isinstance(img, torch.Tensor) and isinstance(rgb, torch.Tensor) and img.is_cuda and rgb.is_cuda):
        raise PiccoloError(f"{what} needs CUDA tensors: piccolo_b200 has no CPU path")


def color_mod(img: torch.Tensor, rgb: torch.Tensor, num_bins: int):
    """Joint histogram equalisation of the luma of panorama and cloud (YCrCb), `sharpen_color` of the configs.
    Returns (img (H,W,3), rgb (N,3)) float32 on img.device."""
    from ._lib import PiccoloError
    _require_cuda_pair(img, rgb, "color_mod")
    if not 2 <= int(num_bins) <= 4096:
        raise PiccoloError(f"color_mod: num_bins {num_bins} outside [2, 4096]")
    return _color_mod_cuda(img.detach(), rgb.detach(), int(num_bins))


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


def _row_weight(H: int) -> np.ndarray:
    """sin(row / H * pi) in fp32 (color_utils.py:216-217)"""
    return np.sin(np.arange(H, dtype=np.float32) / np.float32(H) * np.float32(np.pi)).astype(np.float32)


def _color_match_cuda(img: torch.Tensor, rgb: torch.Tensor):
    """CUDA path (pcl_color.cu): three 256-bin histograms per side on the device, the <= 256-entry interpolation
    here, one rewrite pass on the device.  Returns None when an input is not exactly uint8/255 data."""
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
    `match_color` of the configs.  Returns img (H,W,3) float32 on img.device.  Both inputs must be uint8/255 data (what
    the drivers produce, localize.py:381-414): the device path works on 256-level histograms."""
    from ._lib import PiccoloError
    _require_cuda_pair(img, rgb, "color_match")
    done = _color_match_cuda(img.detach(), rgb.detach())
    if done is None:
        raise PiccoloError("color_match: panorama or cloud colours are not exactly uint8/255 data")
    return done
