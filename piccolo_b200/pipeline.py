"""One localisation query = grid scoring -> top-K -> refinement -> arg-min  (localize.py:207-233 without
file I/O).  Shared by bench.py, main.py/localize.py and the multi-GPU drivers."""
from __future__ import annotations

from collections import namedtuple

import torch

from . import engine
from .utils import grid_poses

TrainCfg = namedtuple("TrainCfg", ["num_input", "num_intermediate", "lr", "num_iter", "patience", "factor", "out_of_room_quantile", "parallel"])

# configs/stanford.ini and configs/stanford_parallel.ini of the reference ([Train]/[Initialization] values)
STANFORD = TrainCfg(6, 50, 0.1, 100, 5, 0.8, 0.05, False)
STANFORD_PARALLEL = STANFORD._replace(parallel=True)


class StartGrid:
    """The start grid as the reference builds it: T translations x R rotations, pose index i*R+j
    (utils.py:484-485).  Scored by the structured-grid kernel; `localize_query*` also accept a plain (P,6) tensor."""

    def __init__(self, trans: torch.Tensor, rot: torch.Tensor):
        self.trans = trans.to(torch.float32).contiguous()
        self.rot = rot.to(torch.float32).contiguous()

    def __len__(self):
        return self.trans.shape[0] * self.rot.shape[0]

    def to(self, device, non_blocking=False):
        return StartGrid(self.trans.to(device, non_blocking=non_blocking), self.rot.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return StartGrid(self.trans.pin_memory(), self.rot.pin_memory())

    def rows(self, lo: int, hi: int) -> "StartGrid":
        """translations [lo, hi) x all rotations = flat indices [lo*R, hi*R)"""
        return StartGrid(self.trans[lo:hi], self.rot)

    def index_select(self, dim: int, idx: torch.Tensor) -> torch.Tensor:
        R = self.rot.shape[0]
        return torch.cat([self.trans.index_select(0, idx // R), self.rot.index_select(0, idx % R)], dim=1)

    def poses(self) -> torch.Tensor:
        return grid_poses(self.trans, self.rot)


def _score(cloud, image, grid):
    if isinstance(grid, StartGrid):
        return engine.score_grid(cloud, image, grid.trans, grid.rot)[0]
    return engine.score(cloud, image, grid)[0]


def query_evals(n_points: int, n_grid: int, cfg) -> int:
    """pose·point evaluations of one query: forward-only grid scoring + num_iter fused fwd+bwd iterations."""
    return n_points * (n_grid + cfg.num_iter * cfg.num_input)


def localize_query(cloud: engine.Cloud, image: engine.Image, grid: torch.Tensor, cfg, timers=None, img: torch.Tensor = None,
                   num_split=(4, 4)):
    """grid: (P,6) start poses on the device, or a StartGrid (translations x rotations).  Returns dict(pose (6,), loss, index, candidates (B,6), losses (B,)).

    Candidate selection as `make_input` (utils.py:624-627): the `num_intermediate` grid poses with the smallest
    sampling loss, re-ranked by colour-histogram intersection down to `num_input`.  The re-rank needs the raw
    panorama `img` (H,W,3); without it the top-`num_input` by loss are refined directly."""
    ev = timers if timers is not None else {}

    def mark(name):
        if name in ev:
            ev[name].record()

    mark("score0")
    loss = _score(cloud, image, grid)
    mark("score1")
    if img is not None:
        idx = engine.topk(loss, cfg.num_intermediate)
        mid = grid.index_select(0, idx)
        scores = engine.hist_rerank(cloud, img, mid, num_split[0], num_split[1])
        keep = engine.topk(-scores, cfg.num_input)
        idx = idx.index_select(0, keep)
    else:
        idx = engine.topk(loss, cfg.num_input)
    mark("rerank1")
    starts = grid.index_select(0, idx)
    ref = engine.Refiner(starts.shape[0], cfg.lr, cfg.factor, cfg.patience, bool(cfg.parallel)).reset(starts)
    mark("refine0")
    ref.run(cloud, image, cfg.num_iter)
    mark("refine1")
    out = ref.read()
    best = out["loss"].argmin()
    b1 = best.reshape(1)                  # index_select, not out[...][best]: indexing with a 0-dim device tensor reads it back (host sync)
    return {"pose": out["pose"].index_select(0, b1)[0], "loss": out["loss"].index_select(0, b1)[0], "index": best, "candidates": out["pose"], "losses": out["loss"],
            "start_index": idx}


def localize_query_host(xyz_h: torch.Tensor, rgb_h: torch.Tensor, img_h: torch.Tensor, grid_h, cfg, device):
    """End-to-end call with HOST (pinned) buffers: upload, pack, score, refine, and read the answer back.
    `rgb_h` / `img_h` may be float32 in [0,1] (the reference's tensors) or uint8 (its data before `/ 255.`).
    Returns (pose (6,) cpu, loss cpu float).  The panorama travels on the upload stream while the cloud is packed
    (Morton sort, clamp box) on the current one; the two join before scoring."""
    main, side = torch.cuda.current_stream(device), _side_stream(device)
    xyz = xyz_h.to(device, non_blocking=True)
    rgb = _unit_from_host(rgb_h, device)                 # float32, or uint8 (expanded to k/255 on the device)
    grid = grid_h.to(device, non_blocking=True)          # (P,6) tensor or StartGrid
    copied = main.record_event()
    cloud = engine.Cloud(xyz, rgb, cfg.out_of_room_quantile)      # enqueued first: Image() below waits on the host for one word
    side.wait_event(copied)                              # copies share the link: the panorama queues behind the cloud's arrays
    with torch.cuda.stream(side):
        img = _unit_from_host(img_h, device)
        img.record_stream(main)
        image = engine.Image(img)
    main.wait_stream(side)
    out = localize_query(cloud, image, grid, cfg, img=img)
    res = torch.cat([out["pose"], out["loss"].reshape(1)]).cpu()
    return res[:6], float(res[6])


_SIDE_STREAMS = {}
_U8_TABLES = {}


def _unit_from_host(t_h: torch.Tensor, device) -> torch.Tensor:
    """Host buffer -> float32 device tensor.  float32 buffers are uploaded as they are; uint8 buffers (what the reference's
    loaders hold before their `/ 255.`: cv2 images, integer colour columns) are uploaded as bytes — a quarter of the PCIe
    traffic — and expanded on the device through a 256-entry table of the correctly rounded k/255 (torch's CUDA division by a
    scalar multiplies by the reciprocal and would not give the values the reference computes on the host)."""
    if t_h.dtype != torch.uint8:
        return t_h.to(device, non_blocking=True)
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _U8_TABLES:
        _U8_TABLES[key] = (torch.arange(256, dtype=torch.float32) / 255.).to(device)
    return _U8_TABLES[key][t_h.to(device, non_blocking=True).long()]



def _side_stream(device) -> "torch.cuda.Stream":
    """ONE upload stream per device for the life of the process: the caching allocators (torch's and the library's
    stream-ordered pool) keep their free lists per stream, so a fresh stream per call would re-allocate every buffer."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def localize_stream(queries, cfg, device, num_split=(4, 4)):
    """A stream of queries with HOST (pinned) buffers: `queries` yields (xyz_h, rgb_h, img_h, grid_h); yields
    (pose (6,) cpu, loss float) per query, in order.  The upload and packing (Morton sort, clamp box, texel tables) of
    query i+1 run on a side stream while query i is scored and refined on the current stream, so in steady state the
    host->device copies cost nothing (they are ~10 % of a C2 query otherwise).  Every query is still uploaded in full."""
    main = torch.cuda.current_stream(device)
    side = _side_stream(device)

    def stage(q):
        xyz_h, rgb_h, img_h, grid_h = q
        with torch.cuda.stream(side):
            xyz = xyz_h.to(device, non_blocking=True)
            rgb = _unit_from_host(rgb_h, device)
            cloud = engine.Cloud(xyz, rgb, cfg.out_of_room_quantile)
            img = _unit_from_host(img_h, device)
            grid = grid_h.to(device, non_blocking=True)
            image = engine.Image(img)
            ready = torch.cuda.Event()
            ready.record(side)
        return cloud, image, img, grid, ready

    def launch(st):
        cloud, image, img, grid, ready = st
        main.wait_event(ready)
        for t in (img, grid.trans, grid.rot) if isinstance(grid, StartGrid) else (img, grid):
            t.record_stream(main)                         # allocated on the side stream, read on this one
        out = localize_query(cloud, image, grid, cfg, img=img, num_split=num_split)
        return torch.cat([out["pose"], out["loss"].reshape(1)])

    it = iter(queries)
    try:
        cur = stage(next(it))
    except StopIteration:
        return
    while cur is not None:
        res_dev = launch(cur)                             # asynchronous: the GPU works on this query ...
        try:
            nxt = stage(next(it))                         # ... while the next one is uploaded and packed
        except StopIteration:
            nxt = None
        res = res_dev.cpu()                               # the step's device->host read
        del cur                                           # handles go back in stream order (the library waits for their last use)
        yield res[:6], float(res[6])
        cur = nxt


def localize_query_sharded(cloud: engine.Cloud, image: engine.Image, grid: torch.Tensor, cfg, img: torch.Tensor = None, num_split=(4, 4),
                           refine: str = "points", timers=None):
    """ONE query sharded over the ranks of the default process group (config C3).  Cloud and panorama are replicated on
    every GPU.  Every phase shards over its own natural unit:
      scoring     contiguous slices of the pose grid per rank; the per-pose losses are all-gathered (NCCL, KB-sized) and
                  every rank runs the same deterministic top-K
      re-rank     contiguous slices of the K intermediate candidates per rank (render + block histograms), rows all-gathered,
                  the reference's sequential candidate loop replayed by every rank
      refinement  refine="points": every rank refines ALL candidates over its share of the POINTS; the per-CTA partial sums
                  travel as peer stores over NVLink inside the persistent kernel (engine.PeerComm), every rank steps the
                  same optimiser on the same sums, so no gather is needed and B < #GPUs does not idle GPUs;
                  refine="candidates": candidates dealt round-robin, one all-gather of the (loss, pose) rows before the arg-min.
    Returns the same dict as `localize_query` on every rank."""
    from . import dist as pdist
    ev = timers if timers is not None else {}

    def mark(name):
        if name in ev:
            ev[name].record()

    mark("score0")
    if isinstance(grid, StartGrid):      # shard the translations: rank slices are whole rows of the loss table
        R = grid.rot.shape[0]
        loss = pdist.score_sharded(lambda tr: engine.score_grid(cloud, image, tr, grid.rot)[0], grid.trans, width=R)
    else:
        loss = pdist.score_sharded(lambda p: engine.score(cloud, image, p)[0], grid)
    mark("score1")
    if img is not None:
        idx = engine.topk(loss, cfg.num_intermediate)
        mid = grid.index_select(0, idx)
        scores = pdist.rerank_sharded(lambda p: engine.hist_rerank_blocks(cloud, img, p, num_split[0], num_split[1]),
                                      lambda rows, ngt: engine.hist_rerank_finish(rows, ngt, num_split[0], num_split[1]), mid)
        keep = engine.topk(-scores, cfg.num_input)
        idx = idx.index_select(0, keep)
    else:
        idx = engine.topk(loss, cfg.num_input)
    mark("rerank1")
    starts = grid.index_select(0, idx)
    rank, ws = pdist.world()
    if refine == "points" and ws > 1 and starts.shape[0] <= 16:
        ref = engine.Refiner(starts.shape[0], cfg.lr, cfg.factor, cfg.patience, bool(cfg.parallel)).reset(starts)
        mark("refine0")
        out = ref.run(cloud, image, cfg.num_iter, comm=pdist.peer_comm()).read()
        mark("refine1")
        table = torch.cat([out["loss"].reshape(-1, 1), out["pose"]], dim=1)       # identical on every rank
    else:
        def refine_fn(s):
            r = engine.Refiner(s.shape[0], cfg.lr, cfg.factor, cfg.patience, bool(cfg.parallel)).reset(s)
            o = r.run(cloud, image, cfg.num_iter).read()
            return torch.cat([o["loss"].reshape(-1, 1), o["pose"]], dim=1)
        mark("refine0")
        table = pdist.refine_sharded(refine_fn, starts)
        mark("refine1")
    k, pose, best = pdist.argmin_candidate(table)
    return {"pose": pose, "loss": best, "index": k, "candidates": table[:, 1:], "losses": table[:, 0], "start_index": idx, "grid_loss": loss}
