"""Config C5 on N GPUs (torchrun, one rank per GPU): N points x P poses with the POSES sharded over the ranks (cloud and
panorama replicated), forward-only scoring and fused forward+backward; the per-pose results are all-gathered (NCCL) so that
every rank holds the full vectors, as the query path needs them.  Time = CUDA events around the sharded call incl. the
all-gather, max over ranks.  Writes a markdown table on rank 0.
    torchrun --nproc-per-node N scripts/sweep_c5_multi.py out.md [sizes_M ...]"""
import json, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import dist as pdist, engine, synth

rank, ws, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=dev)
out_path = sys.argv[1]
sizes = [int(float(a) * 1e6) for a in sys.argv[2:]] or [1_000_000, 10_000_000, 20_000_000]
peak = 6548.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    best = None
    for _ in range(3):
        if ws > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms) if best is None else min(best, float(ms))
    return best


def fwd_bwd_sharded(cloud, image, poses):
    lo, hi = pdist.shard_bounds(poses.shape[0], rank, ws)
    l, c, g = engine.loss_fwd_bwd(cloud, image, poses[lo:hi])
    return pdist._all_gather_rows(torch.cat([l.reshape(-1, 1), g], dim=1), poses.shape[0]) if ws > 1 else g


lines = [f"# C5 sweep on {ws} B200 (poses sharded, cloud replicated): whole-job pose·point evaluations/s; `frac` = per-GPU share of the HBM roofline "
         f"(24 B per evaluation, {peak:.0f} GB/s measured per GPU)\n",
         "Best of 3 x (CUDA-event average incl. the NCCL all-gather of the results, max over ranks); 1024x2048 panorama, Morton-ordered cloud.\n",
         "| N points | P poses | scoring G pp/s | frac per GPU | fwd+bwd B=P G pp/s | frac per GPU |", "|---|---|---|---|---|---|"]
for N in sizes:
    sc = synth.make_scene(N, 1024, 2048, seed=3)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    del xyz, rgb
    rng = np.random.default_rng(0)
    for P in (64, 1024, 8192):
        poses = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.4, 3)]) for _ in range(P)]).astype(np.float32)).to(dev)
        iters = max(1, min(10, int(5e10 * ws / (N * P))))
        ms = timed(lambda: pdist.score_sharded(lambda p: engine.score(cloud, image, p)[0], poses), iters)
        g_f = N * P / ms / 1e6
        ms_b = timed(lambda: fwd_bwd_sharded(cloud, image, poses), iters)
        g_b = N * P / ms_b / 1e6
        lines.append(f"| {N/1e6:.0f} M | {P} | {g_f:.1f} | {24*g_f/peak/ws:.2f} | {g_b:.1f} | {24*g_b/peak/ws:.2f} |")
        if rank == 0:
            print(lines[-1], flush=True)
    del cloud, image
    torch.cuda.empty_cache()
if rank == 0:
    open(out_path, "w").write("\n".join(lines) + "\n")
if ws > 1:
    dist.barrier()
    dist.destroy_process_group()
