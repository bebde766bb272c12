"""CPU oracle for the PICCOLO sampling-loss pose search.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.  The product path
(`piccolo_b200/`) never imports it and has no CPU fallback.

Two independent restatements of the reference algorithm live here:

* the **numpy restatement** (`loss_and_grad_np`, `refine_np`, …): written from the mathematics in
  SURVEY.md §8a, runs in float64 or float32, carries its own bilinear sampler and the *analytic*
  6-DoF gradient.  It does not call torch at all.
* the **ATen-chain restatement** (`sampling_loss_torch`, `refine_torch`, `score_grid_torch`): the
  same algorithm expressed with the third-party primitives the reference itself calls — `torch.atan2`,
  `F.grid_sample(bilinear, zeros, align_corners=False)`, `torch.optim.Adam`, `ReduceLROnPlateau`
  (reference pins torch==1.7.0, `requirements.txt:1`; semantics unchanged in 2.11) — differentiated
  by autograd.  This is what `bench.py` times as the host-CPU baseline ("port").

Parity pin: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the *reference itself*, imported unmodified from /root/reference in the build
container by `tests/golden/make_golden.py`; the vectors are committed under `tests/golden/` and
checked by `tests/test_oracle_golden.py`.

Reference lines followed (all into /root/reference):
  rotation order Rz·Ry·Rx ............ utils.py:425-453, omniloc.py:172-188
  rigid transform q = R (p - t) ...... omniloc.py:190-191, :332-340
  equirect projection ................ utils.py:16-61
  clip ±0.99 + bilinear sample ....... utils.py:64-103
  zero mask + L2 residual + mean ..... omniloc.py:198-200, :347-355
  quantile box ....................... utils.py:208-229
  refinement loop order .............. omniloc.py:44-58 (sequential), :249-269 (batched quirk)
  grid scoring + top-K ............... utils.py:462-507
"""
from __future__ import annotations

import math

import numpy as np

PI = math.pi


# --------------------------------------------------------------------------------------
# numpy restatement (no torch)
# --------------------------------------------------------------------------------------
def rot_and_derivs_np(ypr, dtype=np.float64):
    """R = Rz(yaw)·Ry(pitch)·Rx(roll) and dR/dyaw, dR/dpitch, dR/droll  (utils.py:425-453)."""
    yaw, pitch, roll = [dtype(a) for a in ypr]
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    z, o = dtype(0), dtype(1)
    rz = np.array([[cy, -sy, z], [sy, cy, z], [z, z, o]], dtype=dtype)
    ry = np.array([[cp, z, sp], [z, o, z], [-sp, z, cp]], dtype=dtype)
    rx = np.array([[o, z, z], [z, cr, -sr], [z, sr, cr]], dtype=dtype)
    drz = np.array([[-sy, -cy, z], [cy, -sy, z], [z, z, z]], dtype=dtype)
    dry = np.array([[-sp, z, cp], [z, z, z], [-cp, z, -sp]], dtype=dtype)
    drx = np.array([[z, z, z], [z, -sr, -cr], [z, cr, -sr]], dtype=dtype)
    R = (rz @ ry) @ rx
    return R, (drz @ ry) @ rx, (rz @ dry) @ rx, (rz @ ry) @ drx


def project_np(q, dtype=np.float64):
    """cloud2idx (utils.py:44-59): (N,3) camera-frame points -> u (image x), v (image y) in [-1,1]."""
    eps = dtype(1e-6)
    rho = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1])
    theta = np.arctan2(rho, q[:, 2] + eps)
    phi = np.arctan2(q[:, 1], q[:, 0] + eps) + dtype(PI)
    u = dtype(2) * (dtype(1) - phi / dtype(2 * PI)) - dtype(1)
    v = dtype(2) * (theta / dtype(PI)) - dtype(1)
    return u, v, rho


def _texel(img, iy, ix):
    """Zero-padded texel fetch (grid_sample padding_mode='zeros')."""
    H, W, _ = img.shape
    ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    val = img[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
    return np.where(ok[:, None], val, img.dtype.type(0))


def bilinear_np(img, u, v, dtype=np.float64):
    """clip(±0.99) then bilinear, align_corners=False, zeros padding (utils.py:96-98).
    Returns sample (N,3), ds/dix (N,3), ds/diy (N,3), and the inclusive clip pass-through flags."""
    H, W, _ = img.shape
    c = dtype(0.99)
    uc, vc = np.clip(u, -c, c), np.clip(v, -c, c)
    pass_u, pass_v = (u >= -c) & (u <= c), (v >= -c) & (v <= c)
    ix = ((uc + dtype(1)) * dtype(W) - dtype(1)) / dtype(2)
    iy = ((vc + dtype(1)) * dtype(H) - dtype(1)) / dtype(2)
    x0f, y0f = np.floor(ix), np.floor(iy)
    x0, y0 = x0f.astype(np.int64), y0f.astype(np.int64)
    x1f, y1f = x0f + dtype(1), y0f + dtype(1)
    nw, ne = _texel(img, y0, x0), _texel(img, y0, x0 + 1)
    sw, se = _texel(img, y0 + 1, x0), _texel(img, y0 + 1, x0 + 1)
    wx1, wx0 = (ix - x0f)[:, None], (x1f - ix)[:, None]
    wy1, wy0 = (iy - y0f)[:, None], (y1f - iy)[:, None]
    s = nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1)
    ds_dix = (ne - nw) * wy0 + (se - sw) * wy1
    ds_diy = (sw - nw) * wx0 + (se - ne) * wx1
    return s, ds_dix, ds_diy, pass_u, pass_v


def loss_and_grad_np(xyz, rgb, img, pose, dtype=np.float64, want_grad=True):
    """Sampling loss of ONE pose=(tx,ty,tz,yaw,pitch,roll) and its analytic gradient.

    Follows omniloc.py:171-202; backward is SURVEY.md §8a.  Returns (loss, count, grad(6) or None).
    loss is NaN when no point survives the zero mask (as the reference's empty mean)."""
    xyz = np.asarray(xyz, dtype=dtype)
    rgb = np.asarray(rgb, dtype=dtype)
    img = np.asarray(img, dtype=dtype)
    H, W, _ = img.shape
    pose = np.asarray(pose, dtype=dtype)
    t = pose[:3]
    R, dRy, dRp, dRr = rot_and_derivs_np(pose[3:6], dtype)
    d = xyz - t[None, :]
    q = d @ R.T
    u, v, rho = project_np(q, dtype)
    s, ds_dix, ds_diy, pass_u, pass_v = bilinear_np(img, u, v, dtype)
    m = ~np.all(s == 0, axis=1)
    diff = s - rgb
    e = np.sqrt(np.sum(diff * diff, axis=1))
    M = int(m.sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = dtype(np.sum(e[m], dtype=dtype)) / dtype(M)
    if not want_grad:
        return loss, M, None
    with np.errstate(invalid="ignore", divide="ignore"):
        inv_e = np.where(e > 0, dtype(1) / e, dtype(0))
    g_s = diff * (inv_e * m)[:, None] / dtype(max(M, 1))
    g_ix = dtype(W) / dtype(2) * np.sum(g_s * ds_dix, axis=1) * pass_u
    g_iy = dtype(H) / dtype(2) * np.sum(g_s * ds_diy, axis=1) * pass_v
    g_phi = -g_ix / dtype(PI)
    g_theta = dtype(2) * g_iy / dtype(PI)
    eps = dtype(1e-6)
    qx, qy, qz = q[:, 0], q[:, 1], q[:, 2]
    xp, zp = qx + eps, qz + eps
    with np.errstate(invalid="ignore", divide="ignore"):
        den_phi = xp * xp + qy * qy
        den_th = rho * rho + zp * zp
        inv_rho = np.where(rho > 0, dtype(1) / rho, dtype(0))
        a_phi = g_phi / den_phi
        a_th = g_theta / den_th
    g_qx = a_phi * (-qy) + a_th * zp * qx * inv_rho
    g_qy = a_phi * xp + a_th * zp * qy * inv_rho
    g_qz = a_th * (-rho)
    g_q = np.stack([g_qx, g_qy, g_qz], axis=1)
    g_q = np.where(m[:, None], g_q, dtype(0))
    a = g_q.sum(axis=0)
    G = g_q.T @ d
    grad = np.empty(6, dtype=dtype)
    grad[:3] = -(R.T @ a)
    grad[3] = np.sum(dRy * G)
    grad[4] = np.sum(dRp * G)
    grad[5] = np.sum(dRr * G)
    if M == 0:
        grad[:] = np.nan
    return loss, M, grad


def score_poses_np(xyz, rgb, img, poses, dtype=np.float64):
    """Forward loss of every pose in (P,6).  Returns loss (P,), count (P,)."""
    out = np.empty(len(poses), dtype=dtype)
    cnt = np.empty(len(poses), dtype=np.int64)
    for i, p in enumerate(poses):
        out[i], cnt[i], _ = loss_and_grad_np(xyz, rgb, img, p, dtype, want_grad=False)
    return out, cnt


def topk_ascending(loss, k):
    """Indices of the k smallest losses, ties -> lower index, NaN last (utils.py:501-502 uses an
    unstable argsort; the stable order is the deterministic representative)."""
    loss = np.asarray(loss)
    key = np.where(np.isnan(loss), np.inf, loss)
    return np.argsort(key, kind="stable")[: min(k, len(loss))]


def quantile_box_np(xyz, q):
    """quantile() of utils.py:208-229 for the three axes: order statistics int(N q), int(N (1-q)).
    Returns (lo(3), hi(3)) float32."""
    xyz = np.asarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    i_lo, i_hi = int(n * q), int(n * (1 - q))
    srt = np.sort(xyz, axis=0)
    return srt[i_lo].copy(), srt[i_hi].copy()


class AdamPlateauNP:
    """torch.optim.Adam (defaults β=(0.9,0.999), eps=1e-8) on float32 parameters plus
    ReduceLROnPlateau(mode='min', threshold=1e-4 rel, cooldown=0, min_lr=0, eps=1e-8), restated.
    Scalars (bias corrections, lr, best) are python doubles exactly as torch keeps them
    (called at omniloc.py:33,37,49-50)."""

    def __init__(self, n_param, lr, patience, factor):
        self.m = np.zeros(n_param, dtype=np.float32)
        self.v = np.zeros(n_param, dtype=np.float32)
        self.step_count = 0
        self.lr = float(lr)
        self.patience, self.factor = int(patience), float(factor)
        self.best, self.bad = math.inf, 0

    def adam(self, p, g):
        b1, b2, eps = 0.9, 0.999, 1e-8
        g = g.astype(np.float32)
        self.step_count += 1
        self.m = (self.m + (g - self.m) * np.float32(1 - b1)).astype(np.float32)
        self.v = (self.v * np.float32(b2) + g * g * np.float32(1 - b2)).astype(np.float32)
        bc1 = 1 - b1 ** self.step_count
        bc2 = 1 - b2 ** self.step_count
        step_size = self.lr / bc1
        denom = (np.sqrt(self.v) / np.float32(math.sqrt(bc2)) + np.float32(eps)).astype(np.float32)
        return (p - np.float32(step_size) * (self.m / denom)).astype(np.float32)

    def plateau(self, loss):
        cur = float(loss)
        if cur < self.best * (1 - 1e-4):
            self.best, self.bad = cur, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            new_lr = max(self.lr * self.factor, 0.0)
            if self.lr - new_lr > 1e-8:
                self.lr = new_lr
            self.bad = 0


def refine_np(xyz, rgb, img, poses0, lr=0.1, num_iter=100, patience=5, factor=0.9, q=0.05,
              batch_semantics=False, dtype=np.float32, return_history=False):
    """Refinement loop for B candidates (independent trajectories).

    sequential semantics (omniloc.py:44-58): forward+backward at the clamped parameter.
    batch semantics (omniloc.py:249-269): the forward of iteration k+1 is evaluated at the
    translation as it was AFTER the Adam step of iteration k but BEFORE the clamp (the `torch.cat`
    copy at :260 precedes the clamp at :265-269), while Adam keeps stepping the clamped parameter.

    Returns dict(pose (B,6) float32 — what the reference would return as t/angles,
                 loss (B,) — loss of the LAST forward, param (B,6) — the clamped Adam parameter)."""
    poses0 = np.asarray(poses0, dtype=np.float32).reshape(-1, 6)
    B = poses0.shape[0]
    lo, hi = quantile_box_np(xyz, q)
    param = poses0.copy()
    evalp = poses0.copy()
    opts = [AdamPlateauNP(6, lr, patience, factor) for _ in range(B)]
    last = np.full(B, np.nan, dtype=np.float32)
    hist = []
    for _ in range(num_iter):
        for b in range(B):
            loss, _, g = loss_and_grad_np(xyz, rgb, img, evalp[b], dtype)
            last[b] = np.float32(loss)
            new = opts[b].adam(param[b], np.asarray(g, dtype=np.float32))
            opts[b].plateau(np.float32(loss))
            unclamped = new.copy()
            new[:3] = np.minimum(np.maximum(new[:3], lo), hi)
            param[b] = new
            evalp[b] = unclamped if batch_semantics else new
        if return_history:
            hist.append((last.copy(), evalp.copy()))
    out = {"pose": evalp.copy() if batch_semantics else param.copy(), "loss": last, "param": param,
           "lr": np.array([o.lr for o in opts])}
    if return_history:
        out["history"] = hist
    return out


# --------------------------------------------------------------------------------------
# ATen-chain restatement (torch; the host-CPU baseline that bench.py times)
# --------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def rot_zyx_torch(yaw, pitch, roll):
    """(…,) angle tensors -> (…,3,3) rotation Rz·Ry·Rx, differentiable."""
    torch = _torch()
    cy, sy, cp, sp, cr, sr = torch.cos(yaw), torch.sin(yaw), torch.cos(pitch), torch.sin(pitch), torch.cos(roll), torch.sin(roll)
    r00 = cy * cp
    r01 = cy * sp * sr - sy * cr
    r02 = cy * sp * cr + sy * sr
    r10 = sy * cp
    r11 = sy * sp * sr + cy * cr
    r12 = sy * sp * cr - cy * sr
    r20 = -sp
    r21 = cp * sr
    r22 = cp * cr
    return torch.stack([torch.stack([r00, r01, r02], -1), torch.stack([r10, r11, r12], -1), torch.stack([r20, r21, r22], -1)], -2)


def sampling_loss_torch(xyz, rgb, img, pose):
    """Batched sampling loss with autograd.  pose (B,6) -> per-candidate loss (B,), count (B,).
    xyz (N,3), rgb (N,3), img (H,W,3) share dtype/device with pose."""
    torch = _torch()
    import torch.nn.functional as F
    B = pose.shape[0]
    R = rot_zyx_torch(pose[:, 3], pose[:, 4], pose[:, 5])            # (B,3,3)
    d = xyz[None, :, :] - pose[:, None, :3]                           # (B,N,3)
    q = torch.einsum("bij,bnj->bni", R, d)
    rho = torch.linalg.vector_norm(q[..., :2], dim=-1)
    theta = torch.atan2(rho, q[..., 2] + 1e-6)
    phi = torch.atan2(q[..., 1], q[..., 0] + 1e-6) + PI
    u = 2 * (1.0 - phi / (2 * PI)) - 1
    v = 2 * (theta / PI) - 1
    grid = torch.stack([u, v], dim=-1).reshape(B, -1, 1, 2).clamp(-0.99, 0.99)
    chw = img.permute(2, 0, 1)[None].expand(B, -1, -1, -1)
    s = F.grid_sample(chw, grid, mode="bilinear", padding_mode="zeros", align_corners=False)  # (B,3,N,1)
    s = s[..., 0].transpose(1, 2)                                     # (B,N,3)
    m = (s == 0).sum(-1) != 3
    e = torch.linalg.vector_norm(s - rgb[None], dim=-1) * m
    cnt = m.sum(-1)
    return e.sum(-1) / cnt, cnt


def score_grid_torch(img, xyz, rgb, trans, rot, num_input, pose_chunk=1):
    """trim_input_loss (utils.py:462-507): forward loss over the T×R grid, ascending top-k.
    Returns (trans_k, rot_k, loss_table (T,R))."""
    torch = _torch()
    T, Rn = trans.shape[0], rot.shape[0]
    poses = torch.cat([trans.repeat_interleave(Rn, 0), rot.repeat(T, 1)], dim=1)
    table = torch.empty(T * Rn, dtype=img.dtype)
    with torch.no_grad():
        for i in range(0, T * Rn, pose_chunk):
            table[i:i + pose_chunk] = sampling_loss_torch(xyz, rgb, img, poses[i:i + pose_chunk])[0]
    idx = torch.from_numpy(topk_ascending(table.numpy(), num_input))
    return trans[idx // Rn], rot[idx % Rn], table.reshape(T, Rn)


def quantile_box_torch(xyz, q):
    torch = _torch()
    n = xyz.shape[0]
    srt = torch.sort(xyz, dim=0).values
    return srt[int(n * q)].clone(), srt[int(n * (1 - q))].clone()


def refine_torch(xyz, rgb, img, poses0, lr=0.1, num_iter=100, patience=5, factor=0.9, q=0.05, batch_semantics=False):
    """Refinement with the same third-party optimiser objects the reference instantiates
    (torch.optim.Adam + ReduceLROnPlateau per candidate; omniloc.py:33-37, :235-237).
    poses0 (B,6).  Returns dict(pose (B,6), loss (B,), param (B,6))."""
    torch = _torch()
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    B = poses0.shape[0]
    leaves = [poses0[b].detach().clone().requires_grad_() for b in range(B)]
    opts = [torch.optim.Adam([leaves[b]], lr=lr) for b in range(B)]
    scheds = [ReduceLROnPlateau(opts[b], mode="min", patience=patience, factor=factor) for b in range(B)]
    lo, hi = quantile_box_torch(xyz, q)
    evalp = torch.stack([l.detach().clone() for l in leaves])
    last = None
    for _ in range(num_iter):
        for o in opts:
            o.zero_grad()
        if batch_semantics:
            # evaluate at the un-clamped copy; route its gradient to the leaves (grad of cat is identity)
            probe = evalp.clone().requires_grad_()
            losses, _ = sampling_loss_torch(xyz, rgb, img, probe)
            losses.sum().backward()
            for b in range(B):
                leaves[b].grad = probe.grad[b].clone()
        else:
            losses, _ = sampling_loss_torch(xyz, rgb, img, torch.stack(leaves))
            losses.sum().backward()
        last = losses.detach().clone()
        for b in range(B):
            opts[b].step()
            scheds[b].step(last[b])
        evalp = torch.stack([l.detach().clone() for l in leaves])
        with torch.no_grad():
            for b in range(B):
                leaves[b][:3] = torch.minimum(torch.maximum(leaves[b][:3], lo), hi)
    param = torch.stack([l.detach().clone() for l in leaves])
    return {"pose": evalp if batch_semantics else param, "loss": last, "param": param}


# --------------------------------------------------------------------------------------
# histogram re-rank: make_pano + histogram + trim_input_hist_secondary
# --------------------------------------------------------------------------------------
def make_pano_np(xyz_cam, rgb, H, W):
    """make_pano (utils.py:134-205), CPU semantics: painter's algorithm by nine index_put_ calls in a fixed
    order (neighbours first, centre last), far points first inside a call, last write wins.
    xyz_cam (N,3) float32 camera-frame points.  Returns float32 (H,W,3) = rgb*255 (0 where nothing landed)."""
    xyz_cam = np.asarray(xyz_cam, dtype=np.float32)
    rgb = np.asarray(rgb, dtype=np.float32)
    dist = np.sqrt(np.sum(xyz_cam * xyz_cam, axis=1, dtype=np.float32))
    order = np.argsort(dist, kind="stable")[::-1]
    q, col = xyz_cam[order], rgb[order]
    u, v, _ = project_np(q, np.float32)
    cx = ((u + np.float32(1.0)) / np.float32(2.0)) * np.float32(W - 1)
    cy = ((v + np.float32(1.0)) / np.float32(2.0)) * np.float32(H - 1)
    x, y = cx.astype(np.int64), cy.astype(np.int64)
    img = np.zeros((H, W, 3), dtype=np.float32)
    yp, ym = np.minimum(y + 1, H - 1), np.maximum(y - 1, 0)
    xp, xm = np.minimum(x + 1, W - 1), np.maximum(x - 1, 0)
    for yy, xx in ((y, xm), (y, xp), (ym, xm), (ym, x), (ym, xp), (yp, xm), (yp, x), (yp, xp), (y, x)):
        img[yy, xx] = col            # numpy fancy assignment: later elements win, like CPU index_put_
    return img * np.float32(255.0)


def _hist512(vals_long):
    """histogram() of color_utils.py:68-119 with channels [8,8,8] (bin size ceil(255/8)=32), normalised."""
    b = vals_long // 32
    idx = b[:, 0] + 8 * b[:, 1] + 64 * b[:, 2]
    h = np.bincount(idx, minlength=512).astype(np.float32)
    return h / h.sum()


def hist_rerank_scores_np(img, xyz, rgb, poses, num_split_h, num_split_w):
    """hist_intersect of trim_input_hist_secondary (utils.py:531-579) for K poses, INCLUDING its quirks:
    only the middle row blocks are compared, the split table is not reset between candidates, and an empty
    block `break`s the inner loop leaving the remaining columns stale."""
    img255 = np.asarray(img, dtype=np.float32) * np.float32(255.0)
    H, W, _ = img255.shape
    img_mask = ~np.all(img255 == 0, axis=2)
    bh, bw = H // num_split_h, W // num_split_w
    split = np.zeros(num_split_h * num_split_w, dtype=np.float32)
    out = np.zeros(len(poses), dtype=np.float32)
    xyz = np.asarray(xyz, dtype=np.float32)
    for i, pose in enumerate(np.asarray(poses, dtype=np.float32)):
        R = rot_and_derivs_np(pose[3:6], np.float32)[0]
        q = ((xyz - pose[None, :3]) @ R.T).astype(np.float32)
        proj = make_pano_np(q, rgb, H, W)
        proj_mask = ~np.all(proj == 0, axis=2)
        for h in range(1, num_split_h - 1):
            for w in range(num_split_w):
                block = np.zeros((H, W), dtype=bool)
                block[h * bh:(h + 1) * bh, w * bw:(w + 1) * bw] = True
                fm, fim = proj_mask & img_mask & block, img_mask & block
                if fm.sum() == 0 or fim.sum() == 0:
                    split[h * num_split_w + w] = 0.0
                    break
                ph = _hist512(proj[fm].astype(np.int64))
                ih = _hist512(img255[fim].astype(np.int64))
                split[h * num_split_w + w] = np.minimum(ih, ph).sum()
        split[np.isnan(split)] = 0.0
        out[i] = split.sum() / (num_split_h * num_split_w)
    return out


def hist_rerank_select(scores, num_input):
    """`argsort()[-num_input:]` flipped (utils.py:583-584): descending intersection."""
    order = np.argsort(np.asarray(scores), kind="stable")[-num_input:]
    return order[::-1]
