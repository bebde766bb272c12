"""Thin torch-facing wrappers over the C ABI: device handles and the three compute calls.

torch is plumbing only (device memory, streams); all arithmetic of the hot path runs in the
hand-written CUDA kernels of libpiccolo_b200.so.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import torch

from . import _lib

IMAGE_AUTO, IMAGE_U8Q, IMAGE_F32, IMAGE_U8P, IMAGE_TEX, IMAGE_F16D = 0, 1, 2, 3, 4, 5
IMAGE_FORMATS = {"auto": IMAGE_AUTO, "u8q": IMAGE_U8Q, "f32": IMAGE_F32, "u8p": IMAGE_U8P, "tex": IMAGE_TEX, "f16d": IMAGE_F16D}
CLOUD_KEEP_ORDER, CLOUD_MORTON = 0, 1


def _require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PiccoloError(f"{name} must be a CUDA tensor: piccolo_b200 has no CPU path")
    return t


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _stream(dev) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class Cloud:
    """Device-resident coloured point cloud in kernel layout (SoA, Morton order) + its clamp box.
    Replaces the per-room `xyz.to(device)`, `rgb.to(device)` (localize.py:163-164) and the per-iteration
    `quantile()` calls (omniloc.py:53-55)."""

    def __init__(self, xyz: torch.Tensor, rgb: torch.Tensor, out_of_room_quantile: float = 0.05, order: int = CLOUD_MORTON):
        lib = _lib.load()
        _require_cuda(xyz, "xyz"); _require_cuda(rgb, "rgb")
        if xyz.dim() != 2 or xyz.shape[1] != 3 or rgb.shape != xyz.shape:
            raise _lib.PiccoloError(f"xyz/rgb must both be (N,3); got {tuple(xyz.shape)} and {tuple(rgb.shape)}")
        self.device = xyz.device
        xyz_c, rgb_c = _f32c(xyz), _f32c(rgb)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.pcl_cloud_create(xyz_c.data_ptr(), rgb_c.data_ptr(), xyz_c.shape[0], float(out_of_room_quantile),
                                            int(order), _stream(self.device), ctypes.byref(h)))
        self._h = h
        self.n = int(xyz_c.shape[0])
        self.quantile = float(out_of_room_quantile)
        self._box = None

    def _bounds(self):
        if self._box is None:                 # lazy: the first read blocks on the creation stream
            buf = (ctypes.c_float * 6)()
            _lib.check(_lib.load().pcl_cloud_bounds(self._h, buf))
            self._box = (torch.tensor(list(buf[:3]), dtype=torch.float32), torch.tensor(list(buf[3:]), dtype=torch.float32))
        return self._box

    @property
    def box_lo(self) -> torch.Tensor:
        return self._bounds()[0]

    @property
    def box_hi(self) -> torch.Tensor:
        return self._bounds()[1]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().pcl_cloud_destroy(h)
            except Exception:
                pass


class Image:
    """Device-resident panorama texel table.  Replaces `img.to(device)` (localize.py:170, :213)."""

    def __init__(self, img: torch.Tensor, fmt="auto"):
        lib = _lib.load()
        _require_cuda(img, "img")
        if img.dim() != 3 or img.shape[2] != 3:
            raise _lib.PiccoloError(f"img must be (H,W,3); got {tuple(img.shape)}")
        self.device = img.device
        img_c = _f32c(img)
        self.H, self.W = int(img_c.shape[0]), int(img_c.shape[1])
        h = ctypes.c_void_p()
        code = IMAGE_FORMATS[fmt] if isinstance(fmt, str) else int(fmt)
        with torch.cuda.device(self.device):
            _lib.check(lib.pcl_image_create(img_c.data_ptr(), self.H, self.W, code, _stream(self.device), ctypes.byref(h)))
        self._h = h
        self.format = int(lib.pcl_image_format(h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().pcl_image_destroy(h)
            except Exception:
                pass


def _poses(poses: torch.Tensor) -> torch.Tensor:
    _require_cuda(poses, "poses")
    p = _f32c(poses)
    if p.dim() != 2 or p.shape[1] != 6:
        raise _lib.PiccoloError(f"poses must be (P,6) = (tx,ty,tz,yaw,pitch,roll); got {tuple(p.shape)}")
    return p


def score(cloud: Cloud, image: Image, poses: torch.Tensor):
    """Forward-only loss of P poses.  Returns (loss (P,), count (P,)) on the device."""
    lib = _lib.load()
    p = _poses(poses)
    P = p.shape[0]
    loss = torch.empty(P, dtype=torch.float32, device=p.device)
    count = torch.empty(P, dtype=torch.float32, device=p.device)
    if P == 0:
        return loss, count
    with torch.cuda.device(p.device):
        _lib.check(lib.pcl_score(cloud._h, image._h, p.data_ptr(), P, loss.data_ptr(), count.data_ptr(), _stream(p.device)))
    return loss, count


def score_grid(cloud: Cloud, image: Image, trans: torch.Tensor, rot: torch.Tensor):
    """Forward-only loss of the T x R start grid (translation i, rotation j) -> flat index i*R+j, the loop of
    `trim_input_loss` (utils.py:484-499).  Rotations related by an in-plane turn about the camera z axis share most
    of the per-point work (pcl_grid.cu).  Returns (loss (T*R,), count (T*R,)) on the device."""
    lib = _lib.load()
    _require_cuda(trans, "trans")
    _require_cuda(rot, "rot")
    t, r = _f32c(trans), _f32c(rot)
    if t.dim() != 2 or t.shape[1] != 3 or r.dim() != 2 or r.shape[1] != 3:
        raise _lib.PiccoloError(f"trans must be (T,3) and rot (R,3) = (yaw,pitch,roll); got {tuple(t.shape)}, {tuple(r.shape)}")
    P = t.shape[0] * r.shape[0]
    loss = torch.empty(P, dtype=torch.float32, device=t.device)
    count = torch.empty(P, dtype=torch.float32, device=t.device)
    if P == 0:
        return loss, count
    with torch.cuda.device(t.device):
        _lib.check(lib.pcl_score_grid(cloud._h, image._h, t.data_ptr(), t.shape[0], r.data_ptr(), r.shape[0], loss.data_ptr(), count.data_ptr(),
                                      _stream(t.device)))
    return loss, count


def loss_fwd_bwd(cloud: Cloud, image: Image, poses: torch.Tensor):
    """Loss and analytic 6-DoF gradient of B poses in one launch.  Returns (loss (B,), count (B,), grad (B,6))."""
    lib = _lib.load()
    p = _poses(poses)
    B = p.shape[0]
    loss = torch.empty(B, dtype=torch.float32, device=p.device)
    count = torch.empty(B, dtype=torch.float32, device=p.device)
    grad = torch.empty(B, 6, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(lib.pcl_loss_fwd_bwd(cloud._h, image._h, p.data_ptr(), B, loss.data_ptr(), count.data_ptr(), grad.data_ptr(), _stream(p.device)))
    return loss, count, grad


def topk(loss: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k smallest losses (ascending, ties -> lower index, NaN last), int64 on the device."""
    lib = _lib.load()
    _require_cuda(loss, "loss")
    l = _f32c(loss).reshape(-1)
    k = min(int(k), l.numel())
    idx = torch.empty(k, dtype=torch.int64, device=l.device)
    if k == 0:
        return idx
    with torch.cuda.device(l.device):
        _lib.check(lib.pcl_topk(l.data_ptr(), l.numel(), k, idx.data_ptr(), _stream(l.device)))
    return idx


def hist_rerank(cloud: Cloud, img: torch.Tensor, poses: torch.Tensor, num_split_h: int = 4, num_split_w: int = 4) -> torch.Tensor:
    """`hist_intersect` of trim_input_hist_secondary (utils.py:531-579) for K poses: (K,) float32, larger = better."""
    lib = _lib.load()
    _require_cuda(img, "img")
    p = _poses(poses)
    img_c = _f32c(img)
    out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(lib.pcl_hist_rerank(cloud._h, img_c.data_ptr(), int(img_c.shape[0]), int(img_c.shape[1]), p.data_ptr(), p.shape[0],
                                       int(num_split_h), int(num_split_w), out.data_ptr(), _stream(p.device)))
    return out


def hist_rerank_blocks(cloud: Cloud, img: torch.Tensor, poses: torch.Tensor, num_split_h: int = 4, num_split_w: int = 4):
    """Stage 1 of the re-rank, independent per candidate (pcl_hist_rerank_blocks): returns (rows (K, 2*nblk), ngt (nblk,))."""
    lib = _lib.load()
    _require_cuda(img, "img")
    p = _poses(poses)
    img_c = _f32c(img)
    nblk = max(int(num_split_h) - 2, 0) * int(num_split_w)
    rows = torch.zeros(p.shape[0], 2 * nblk, dtype=torch.float32, device=p.device)
    ngt = torch.zeros(nblk, dtype=torch.float32, device=p.device)
    if p.shape[0] > 0:
        with torch.cuda.device(p.device):
            _lib.check(lib.pcl_hist_rerank_blocks(cloud._h, img_c.data_ptr(), int(img_c.shape[0]), int(img_c.shape[1]), p.data_ptr(), p.shape[0],
                                                  int(num_split_h), int(num_split_w), rows.data_ptr(), ngt.data_ptr(), _stream(p.device)))
    return rows, ngt


def hist_rerank_finish(rows: torch.Tensor, ngt: torch.Tensor, num_split_h: int = 4, num_split_w: int = 4) -> torch.Tensor:
    """Stage 2 over ALL candidates in order (pcl_hist_rerank_finish): (K,) hist_intersect."""
    lib = _lib.load()
    _require_cuda(rows, "rows")
    r, g = _f32c(rows), _f32c(ngt)
    out = torch.empty(r.shape[0], dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device):
        _lib.check(lib.pcl_hist_rerank_finish(r.data_ptr(), g.data_ptr(), r.shape[0], int(num_split_h), int(num_split_w), out.data_ptr(), _stream(r.device)))
    return out


class Refiner:
    """Fused refinement of B candidates: every iteration is ONE kernel launch doing loss, backward,
    reduction, Adam, ReduceLROnPlateau and the translation clamp (omniloc.py:44-58, :249-269)."""

    def __init__(self, B: int, lr: float = 0.1, factor: float = 0.9, patience: int = 5, batch_semantics: bool = False):
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.pcl_refine_create(int(B), float(lr), float(factor), int(patience), int(bool(batch_semantics)), ctypes.byref(h)))
        self._h = h
        self.B = int(B)

    def reset(self, poses: torch.Tensor):
        p = _poses(poses)
        if p.shape[0] != self.B:
            raise _lib.PiccoloError(f"expected {self.B} start poses, got {p.shape[0]}")
        self.device = p.device
        with torch.cuda.device(p.device):
            _lib.check(_lib.load().pcl_refine_reset(self._h, p.data_ptr(), _stream(p.device)))
        return self

    def run(self, cloud: Cloud, image: Image, num_iter: int, comm: "PeerComm" = None):
        """comm: a connected PeerComm -> the POINTS of the cloud are sharded over its ranks (every rank passes the same
        cloud, image and start poses; the per-CTA partial sums travel as peer stores over NVLink inside the kernel)."""
        with torch.cuda.device(self.device):
            if comm is not None and comm.size > 1:
                _lib.check(_lib.load().pcl_refine_run_sharded(self._h, cloud._h, image._h, int(num_iter), comm._h, _stream(self.device)))
            else:
                _lib.check(_lib.load().pcl_refine_run(self._h, cloud._h, image._h, int(num_iter), _stream(self.device)))
        return self

    def debug_stats(self):
        """Option RF_DEBUG: (ctas + 1, 4) uint64 numpy array of the last persistent run: compute CTAs {cycles in phases, cycles
        waiting for poses, phases with prefetched poses, 0}, last row = service CTA {waiting, working}."""
        import numpy as np
        buf = np.zeros((512, 4), dtype=np.uint64)
        with torch.cuda.device(self.device):
            n = _lib.load().pcl_refine_debug_stats(self._h, buf.ctypes.data, 512, _stream(self.device))
        if n < 0:
            _lib.check(n)
        return buf[:n]

    def debug_timeline(self):
        """Option RF_DEBUG: wall-clock stamps (ns, uint64 numpy array) of the end of every iteration of the last persistent run."""
        import numpy as np
        buf = np.zeros(4096, dtype=np.uint64)
        with torch.cuda.device(self.device):
            n = _lib.load().pcl_refine_debug_timeline(self._h, buf.ctypes.data, 4096, _stream(self.device))
        if n < 0:
            _lib.check(n)
        return buf[:n]

    def read(self):
        """Returns dict(pose (B,6), param (B,6), loss (B,), lr (B,) float64) on the device."""
        dev = self.device
        pose = torch.empty(self.B, 6, dtype=torch.float32, device=dev)
        param = torch.empty(self.B, 6, dtype=torch.float32, device=dev)
        loss = torch.empty(self.B, dtype=torch.float32, device=dev)
        lr = torch.empty(self.B, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pcl_refine_read(self._h, pose.data_ptr(), param.data_ptr(), loss.data_ptr(), lr.data_ptr(), _stream(dev)))
        return {"pose": pose, "param": param, "loss": loss, "lr": lr}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().pcl_refine_destroy(h)
            except Exception:
                pass


class PeerComm:
    """Peer-memory communicator of the ranks of one box (include/piccolo_b200.h: pcl_comm_*): every rank owns a
    device window that the others map through CUDA IPC; kernels then exchange data by stores and system-scope
    atomics over NVLink.  `exchange` is any callable that all-gathers one bytes object per rank in rank order
    (piccolo_b200.dist.peer_comm() passes torch.distributed.all_gather_object)."""

    def __init__(self, rank: int, size: int, exchange, device=None, window_bytes: int = 0):
        lib = _lib.load()
        self.rank, self.size = int(rank), int(size)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.pcl_comm_create(self.rank, self.size, int(window_bytes), ctypes.byref(h)))
            self._h = h
            if self.size > 1:
                mine = ctypes.create_string_buffer(64)
                _lib.check(lib.pcl_comm_handle(h, mine))
                handles = exchange(mine.raw)
                if len(handles) != self.size or any(len(x) != 64 for x in handles):
                    raise _lib.PiccoloError("peer handle exchange returned a malformed list")
                _lib.check(lib.pcl_comm_connect(h, ctypes.create_string_buffer(b"".join(handles), 64 * self.size)))

    def barrier(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pcl_comm_barrier(self._h, _stream(self.device)))

    def all_gather(self, local: torch.Tensor) -> torch.Tensor:
        """(n,) float32 on the device -> (size, n): row k is rank k's vector (n <= 16384, equal on all ranks)."""
        _require_cuda(local, "local")
        x = _f32c(local).reshape(-1)
        out = torch.empty(self.size, x.numel(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pcl_comm_allgather_f32(self._h, x.data_ptr(), x.numel(), out.data_ptr(), _stream(x.device)))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().pcl_comm_destroy(h)
            except Exception:
                pass


# ------------------------------------------------------------------------------------------------
# handle cache: the reference API passes raw tensors on every call (and builds a new loss module per
# candidate, omniloc.py:42); re-packing the cloud each time would dominate small queries.
# ------------------------------------------------------------------------------------------------
_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_SIZE = 8


def _key(*tensors, extra=()):
    """Identity of a packed handle = storage address, shape, dtype, device and torch's in-place version counter of its
    source tensors.  Every write torch knows about (in-place ops, `copy_`, slicing assignments) bumps the counter and
    misses the cache; a write made behind torch's back (a foreign CUDA kernel or memcpy on the raw pointer) does not —
    call `clear_cache()` after such writes."""
    return tuple((t.data_ptr(), tuple(t.shape), t._version, str(t.device), t.dtype) for t in tensors) + tuple(extra)


def _cached(key, tensors, build):
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit[0]
    obj = build()
    _CACHE[key] = (obj, tensors)        # hold the tensors: their storage cannot be recycled while cached
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return obj


def get_cloud(xyz: torch.Tensor, rgb: torch.Tensor, q: float = None) -> Cloud:
    """The packed cloud of (xyz, rgb), built once per tensor pair.  `q` (out_of_room_quantile) only matters to the
    refinement's clamp box: callers that do not care (scoring, re-rank) pass None and share whatever cloud is cached, so
    a config with another quantile does not pack (Morton sort + three radix sorts) the same cloud twice per query."""
    key = _key(xyz, rgb, extra=("cloud",))
    hit = _CACHE.get(key)
    if hit is not None and (q is None or abs(hit[0].quantile - float(q)) < 1e-12):
        _CACHE.move_to_end(key)
        return hit[0]
    _CACHE.pop(key, None)
    return _cached(key, (xyz, rgb), lambda: Cloud(xyz, rgb, 0.05 if q is None else q))


def get_image(img: torch.Tensor, fmt="auto") -> Image:
    return _cached(_key(img, extra=("image", fmt)), (img,), lambda: Image(img, fmt))


def clear_cache():
    _CACHE.clear()
