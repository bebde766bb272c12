"""Small run that touches every kernel of the library (all texel formats, scoring, fwd+bwd, refinement with PDL,
re-rank, top-k, ragged cloud sizes) — target for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(7001, 64, 128, seed=2)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
rng = np.random.default_rng(0)
poses = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.2, 3)]) for _ in range(70)]).astype(np.float32)).to(dev)
LIGHT = os.environ.get("PCL_SANITIZE_LIGHT") == "1"     # initcheck is ~50x slower than the other tools: one order, two formats
for order in ((1,) if LIGHT else (0, 1)):
    cloud = engine.Cloud(xyz, rgb, 0.05, order)
    for fmt in (("auto", "u8q") if LIGHT else ("auto", "f16d", "u8q", "u8p", "tex", "f32")):
        image = engine.Image(img, fmt)
        loss, cnt = engine.score(cloud, image, poses)
        from piccolo_b200.utils import generate_rot_points
        for rot in (generate_rot_points({"yaw_only": True, "num_yaw": 8}), generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4}),
                    poses[:40, 3:].cpu()):                              # one group | 6 groups of 4 | R > 32: expanded + generic kernel
            gl, gc = engine.score_grid(cloud, image, poses[:7, :3].contiguous(), rot.to(dev))
        l2, c2, g = engine.loss_fwd_bwd(cloud, image, poses[:40])
        l3, c3, g3 = engine.loss_fwd_bwd(cloud, image, poses[:6])
        out = engine.Refiner(6, 0.1, 0.8, 5, True).reset(poses[:6]).run(cloud, image, 4).read()
    # fused refinement: persistent cooperative kernel (resident points, service CTA, split-phase hand-over) and the
    # per-iteration fallback, B = 1 .. 16, both loop semantics, split runs; B = 20 takes the generic kernel
    from piccolo_b200 import _lib
    image = engine.Image(img)
    for persist in (1, 0):
        _lib.set_option("PERSIST", persist)
        for B, batch in ((1, False), (2, True), (5, False), (6, True), (11, True), (16, False), (20, True)):
            r = engine.Refiner(B, 0.1, 0.8, 5, batch).reset(poses[:B])
            r.run(cloud, image, 3); r.run(cloud, image, 2)
            out = r.read()
    _lib.set_option("PERSIST", -1)
    for res, small in ((0, 1), (1, 2), (1, 0)):            # streamed points / resident points x compact / fp16-basis table
        _lib.set_option("RF_RES", res); _lib.set_option("SMALL_TABLE", small)
        out = engine.Refiner(6, 0.1, 0.8, 5, True).reset(poses[:6]).run(cloud, image, 3).read()
    _lib.set_option("RF_RES", -1); _lib.set_option("SMALL_TABLE", -1)
    idx = engine.topk(loss, 10)
    s = engine.hist_rerank(cloud, img, poses[:12], 4, 4)
from piccolo_b200.color_utils import color_match, color_mod
cm = color_match(img, rgb)
mi, mr = color_mod(img, rgb, 256)
# a stream of queries: handles built on a side stream, used on the current one, freed in stream order
from piccolo_b200 import pipeline
grid = pipeline.StartGrid(poses[:5, :3].cpu(), generate_rot_points({"yaw_only": True, "num_yaw": 8})).pin_memory()
q = tuple(torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)) + (grid,)
cfg = pipeline.STANFORD_PARALLEL._replace(num_iter=3, num_intermediate=12, num_input=3)
res = list(pipeline.localize_stream((q for _ in range(3)), cfg, dev))
torch.cuda.synchronize()
print("sanitize target ok", float(loss.min()), idx[:3].tolist(), float(s.max()))
