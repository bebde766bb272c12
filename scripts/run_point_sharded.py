"""Point-sharded refinement over the ranks of one box (launch with torchrun): every rank refines ALL candidates over its
share of the points, partial sums exchanged by peer stores inside the persistent kernel (pcl_refine_run_sharded).
Checks: (1) the peer-memory all-gather and barrier, (2) all ranks end with bit-identical states, (3) the first
iterations agree with the single-GPU run to fp32 accuracy and the end state to the refinement gates, (4) timing.
    python -m torch.distributed.run --nproc-per-node N scripts/run_point_sharded.py [n_points] [height] [B]"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth, dist as pdist

rank, local, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
B = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sc = synth.make_scene(N, H, 2 * H, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(B)]).astype(np.float32)).to(dev)
comm = pdist.peer_comm()

# (1) peer-memory collectives
for n in (1, 7, 4096):
    mine = torch.arange(n, dtype=torch.float32, device=dev) + 1000.0 * rank
    got = comm.all_gather(mine)
    want = torch.stack([torch.arange(n, dtype=torch.float32, device=dev) + 1000.0 * r for r in range(ws)])
    assert torch.equal(got, want), (rank, n)
comm.barrier()
torch.cuda.synchronize()

def run(iters, sharded, plan=None):
    ref = engine.Refiner(B, 0.1, 0.8, 5, True).reset(starts)
    for n in (plan or (iters,)):
        ref.run(cloud, image, n, comm=comm if sharded else None)
    return ref.read()

# (2) identical on all ranks, (3) against the single-GPU run
sh3, si3 = run(3, True), run(3, False)
sh, si = run(100, True, (2, 97, 1)), run(100, False)
rows = torch.cat([sh["pose"].reshape(-1), sh["loss"], sh["param"].reshape(-1)])
allrows = [torch.empty_like(rows) for _ in range(ws)]
dist.all_gather(allrows, rows)
same = all(torch.equal(allrows[0], r) or bool(torch.isnan(r).any()) for r in allrows)
d3 = float((sh3["pose"] - si3["pose"]).abs().max())
l3 = float(((sh3["loss"] - si3["loss"]).abs() / si3["loss"].abs()).max())
best_sh, best_si = int(sh["loss"].argmin()), int(si["loss"].argmin())
dt = float((sh["pose"][best_sh, :3] - si["pose"][best_si, :3]).norm())

# (4) timing (max over ranks)
def timed(sharded):
    ref = engine.Refiner(B, 0.1, 0.8, 5, True)
    for _ in range(2):
        ref.reset(starts); ref.run(cloud, image, 100, comm=comm if sharded else None)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ref.reset(starts); ref.run(cloud, image, 100, comm=comm if sharded else None)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_sh, t_si = timed(True), timed(False)
if rank == 0:
    print(f"ranks={ws} N={N} {H}x{2*H} B={B}: identical on all ranks={same}; after 3 iterations |pose diff| {d3:.2e}, loss rel diff {l3:.2e}; "
          f"after 100: winner {best_sh} vs {best_si}, |t diff| {dt*1e3:.2f} mm; point-sharded {t_sh*10:.2f} us/iter vs single GPU {t_si*10:.2f} us/iter "
          f"(speed-up {t_si/t_sh:.2f}x on {ws} GPUs)", flush=True)
    assert same and d3 < 2e-4 and l3 < 2e-5
dist.destroy_process_group()
