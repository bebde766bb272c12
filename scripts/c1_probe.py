"""C1 query start selection: ours vs the reference's two runs (tests/golden/query_c1.npz, variants.npz c1_*_t4)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import synth
from piccolo_b200.localize import get_init_dict
from piccolo_b200.omniloc import omniloc_all
from piccolo_b200.parse_utils import parse_ini
from piccolo_b200.utils import make_input
g = np.load("tests/golden/query_c1.npz"); v = np.load("tests/golden/variants.npz")
cfg = parse_ini("configs/stanford.ini")
sc = synth.make_scene(200_000, 512, 1024, seed=3)
dev = torch.device("cuda:0")
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
in_t, in_r = make_input(img, xyz, rgb, cfg.num_input, get_init_dict(cfg), cfg.criterion, cfg.num_intermediate)
ours = np.concatenate([in_t.cpu().numpy(), in_r.cpu().numpy()], 1).astype(np.float64)
ref = np.concatenate([g["input_trans"], g["input_rot"]], 1).astype(np.float64)
ref4 = np.concatenate([v["c1_input_trans_t4"], v["c1_input_rot_t4"]], 1).astype(np.float64)
print("ours\n", np.round(ours, 4)); print("reference\n", np.round(ref, 4))
print("max |ours - ref| row by row:", np.abs(ours - ref).max(1), " ref vs ref(4 threads):", np.abs(ref - ref4).max())
res = omniloc_all(img, xyz, rgb, in_t, in_r, cfg)
print("ours final losses", [round(float(r[2]), 5) for r in res]); print("ref  final losses", np.round(g["final_loss"], 5) if "final_loss" in g else None, np.round(v["c1_final_loss_t4"], 5))
print("ours final t", [np.round(r[0].numpy().reshape(3), 4).tolist() for r in res]); print("ref final t", np.round(g["final_t"], 4).tolist()); print("ref4 final t", np.round(v["c1_final_t_t4"], 4).tolist())
