"""Config C4: Q query panoramas (colour / scene perturbed) against one cloud, queries sharded over the ranks
(launch with torchrun), omniscenes.ini settings (xy-only translation lattice at a z prior, 8 yaws, K=50, B=6,
100 iterations, omniloc_batch semantics when --parallel).  Every rank localises its queries with zero
communication; one NCCL all-gather of the result rows at the end.  Reports seconds per query and accuracy.
    python -m torch.distributed.run --nproc-per-node N scripts/run_multi_query.py [n_points] [n_queries]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, pipeline, synth
from piccolo_b200.color_utils import color_match, requantize
from piccolo_b200.dist import shard_bounds
from piccolo_b200.utils import generate_rot_points, generate_trans_points, grid_poses

rank, local, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if ws > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 64
MATCH = os.environ.get("PCL_MATCH_COLOR", "1") == "1"
H, W = 1024, 2048                                                  # localize.py:381 forces 2048x1024
room = (8.0, 6.0, 3.0)
xyz_np, rgb8 = synth.sample_room_points(N, room, seed=2)
xyz, rgb = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(synth.rgb_from_u8(rgb8)).to(dev)
cloud = engine.Cloud(xyz, rgb)
init = {"yaw_only": True, "num_yaw": 8, "xy_only": True, "num_trans": 150, "z_prior": 1.5, "trans_init_mode": "quantile", "dataset": "OmniScenes"}
grid = pipeline.StartGrid(generate_trans_points(xyz, init, device=dev), generate_rot_points(init, device=dev))
cfg = pipeline.STANFORD_PARALLEL
lo, hi = shard_bounds(Q, rank, ws)
# queries of this rank: seeded GT poses (yaw-only, z at the prior), perturbed panoramas (gamma, white balance, re-textured patches)
queries = []
for q in range(lo, hi):
    gt = synth.random_gt_pose(room, seed=100 + q, yaw_only=True); gt[2] = 1.5
    img8 = synth.perturb_panorama(synth.render_panorama(gt, H, W, room), seed=q, gamma=1.0 + 0.05 * (q % 5), wb=(1.0, 0.96 + 0.01 * (q % 7), 1.02), retexture_frac=0.1)
    queries.append((gt, torch.from_numpy(synth.img_from_u8(img8)).pin_memory()))
def run_all():
    rows = []
    for gt, img_h in queries:
        img = img_h.to(dev, non_blocking=True)
        if MATCH:                                                    # match_color=True of omniscenes.ini (localize.py:402-404):
            img = requantize(color_match(img, rgb))                  # CDF matching on the device + the drivers' uint8 round trip
        out = pipeline.localize_query(cloud, engine.Image(img), grid, cfg, img=img)
        rows.append(torch.cat([out["pose"], out["loss"].reshape(1)]))
    return torch.stack(rows) if rows else torch.zeros((0, 7), device=dev)
run_all()                                                            # warm-up
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
t0 = time.perf_counter()
rows = run_all()
if ws > 1:                                                           # the only exchange: result rows
    pad = torch.zeros(((Q + ws - 1) // ws, 7), device=dev); pad[: rows.shape[0]] = rows
    allrows = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(allrows, pad)
torch.cuda.synchronize()
if ws > 1:
    dist.barrier()
dt = time.perf_counter() - t0
errs = []
for (gt, _), r in zip(queries, rows.cpu().numpy()):
    errs.append((np.linalg.norm(r[:3] - gt[:3]), abs(((r[3] - gt[3] + np.pi) % (2 * np.pi)) - np.pi) * 180 / np.pi))
ok = sum(1 for t, a in errs if t < 0.1 and a < 5.0)
stat = torch.tensor([float(ok), float(len(errs))], device=dev)
if ws > 1:
    dist.all_reduce(stat)
if rank == 0:
    evals = Q * pipeline.query_evals(N, len(grid), cfg)
    print(f"C4: ranks={ws} N={N} queries={Q} grid={len(grid)} poses: {dt:.3f} s total, {dt/Q*1e3:.1f} ms/query wall ({dt/max(1,hi-lo)*1e3:.1f} ms per query per GPU), "
          f"{evals/dt/1e9:.1f} G pp/s aggregate; match_color={MATCH}; localised {int(stat[0])}/{int(stat[1])} (t<0.1 m, r<5 deg)")
if ws > 1:
    dist.destroy_process_group()
