"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference (pure Python, `omniloc.py`, `utils.py`) is imported from /root/reference behind three
import-only stub modules (`torch_scatter`, `matplotlib.pyplot`, `open3d` are imported by the
reference but never used on this path — SURVEY.md §8c).  Every fixture stores its INPUTS (uint8
colours/panorama, float32 xyz, poses) and the reference's OUTPUTS, so the tests never need the
reference or the synthetic generator to be bit-reproducible on another machine.
"""
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PICCOLO_REFERENCE", "/root/reference")


def import_reference():
    ts = types.ModuleType("torch_scatter"); ts.scatter_min = None
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
    o3d = types.ModuleType("open3d")
    sys.modules.setdefault("torch_scatter", ts)
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("open3d", o3d)
    sys.path.insert(0, REF)
    import omniloc as ref_omniloc
    import utils as ref_utils
    return ref_omniloc, ref_utils


def ref_loss_grad(ref_omniloc, xyz, rgb, img, pose, dtype):
    cfg = namedtuple("C", ["num_input"])(1)
    xyz_t, rgb_t, img_t = [torch.from_numpy(a).to(dtype) for a in (xyz, rgb, img)]
    mod = ref_omniloc.SamplingLoss(xyz_t, rgb_t, img_t, torch.device("cpu"), cfg)
    mod.tensor_0 = mod.tensor_0.to(dtype); mod.tensor_1 = mod.tensor_1.to(dtype)
    p = torch.tensor(pose, dtype=dtype)
    t = p[:3].reshape(3, 1).clone().requires_grad_()
    yaw, pitch, roll = [p[i:i + 1].clone().requires_grad_() for i in (3, 4, 5)]
    loss = mod(t, yaw, pitch, roll)
    if torch.isnan(loss):
        return float("nan"), np.full(6, np.nan)
    loss.backward()
    g = np.concatenate([t.grad.reshape(3).numpy(), yaw.grad.numpy(), pitch.grad.numpy(), roll.grad.numpy()])
    return float(loss), g.astype(np.float64)


def make_score_lattice(ref_omniloc, ref_utils):
    """fixture 2d: trim_input_loss over translations x the 24 distinct rotations of the 4x4x4 Euler lattice (the
    stanford.ini start grid), on the loss_small scene: the multi-group case of the structured-grid kernel."""
    from piccolo_b200 import synth, utils as pu
    small = np.load(os.path.join(HERE, "loss_small.npz"))
    xyz, rgb, img = small["xyz"], synth.rgb_from_u8(small["rgb8"]), synth.img_from_u8(small["img8"])
    rot = pu.generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4})        # deterministic order
    rng = np.random.default_rng(21)
    lo, hi = xyz.min(0), xyz.max(0)
    trans = torch.from_numpy((lo + (hi - lo) * rng.uniform(0.2, 0.8, (5, 3))).astype(np.float32))
    xyz_t, rgb_t, img_t = torch.from_numpy(xyz), torch.from_numpy(rgb), torch.from_numpy(img)
    tt, rr = ref_utils.trim_input_loss(img_t, xyz_t, rgb_t, trans, rot, 10)
    grid = torch.cat([trans.repeat_interleave(len(rot), 0), rot.repeat(len(trans), 1)], 1).numpy()
    table = np.array([ref_loss_grad(ref_omniloc, xyz, rgb, img, g, torch.float32)[0] for g in grid])
    np.savez_compressed(os.path.join(HERE, "score_lattice.npz"), trans=trans.numpy(), rot=rot.numpy(), loss_table=table,
                        top10_trans=tt.numpy(), top10_rot=rr.numpy())
    print("score_lattice: best", table.min(), "worst", table.max(), "poses", len(table))


def make_color_small():
    """fixture: colour preprocessing (color_utils.py:7-65 color_mod, :146-234 color_match) of the unmodified reference on
    a perturbed 64x128 panorama (gamma, white balance, re-textured patches) against a 20 k-point cloud."""
    from piccolo_b200 import synth
    sys.path.insert(0, REF)
    import color_utils as ref_color
    sc = synth.make_scene(20000, 64, 128, seed=9)
    img8 = synth.perturb_panorama(sc.img8, seed=4, gamma=1.15, wb=(1.0, 0.95, 1.04), retexture_frac=0.1)
    img, rgb = torch.from_numpy(synth.img_from_u8(img8)), torch.from_numpy(synth.rgb_from_u8(sc.rgb8))
    mod_img, mod_rgb = ref_color.color_mod(img.clone(), rgb.clone(), 256)
    match_img = ref_color.color_match(img.clone(), rgb.clone())
    np.savez_compressed(os.path.join(HERE, "color_small.npz"), img8=img8, rgb8=sc.rgb8, mod_img=mod_img.numpy(), mod_rgb=mod_rgb.numpy(),
                        match_img=match_img.numpy())
    print("color_small: lit fraction", float((img8.reshape(-1, 3).sum(1) > 0).mean()), "match range", float(match_img.min()), float(match_img.max()))


def _rows_to_index(rows, table, decimals=5):
    """index of every row of `rows` in `table` (exact match after rounding)"""
    key = {tuple(np.round(r, decimals)): i for i, r in enumerate(np.asarray(table, dtype=np.float64))}
    return np.array([key[tuple(np.round(r, decimals))] for r in np.asarray(rows, dtype=np.float64)], dtype=np.int64)


def make_variants(ref_omniloc, ref_utils, with_c1=True):
    """fixture `variants`: how much the REFERENCE ITSELF varies on the discrete decisions of the path when only the
    floating-point evaluation order / precision changes (1 vs 8 intra-op threads, fp32 vs fp64).  The GPU tests gate
    "identical top-K" against this measured set instead of a hand-picked slack."""
    from piccolo_b200 import synth, utils as pu
    out = {}

    def with_threads(n, fn):
        old = torch.get_num_threads()
        torch.set_num_threads(n)
        try:
            return fn()
        finally:
            torch.set_num_threads(old)

    def with_dtype(dtype, fn):
        torch.set_default_dtype(dtype)
        try:
            return fn()
        finally:
            torch.set_default_dtype(torch.float32)

    # ---- (1) trim_input_loss on the 5 x 24 lattice grid of score_lattice ---------------------------------------
    small = np.load(os.path.join(HERE, "loss_small.npz"))
    lat = np.load(os.path.join(HERE, "score_lattice.npz"))
    xyz, rgb, img = small["xyz"], synth.rgb_from_u8(small["rgb8"]), synth.img_from_u8(small["img8"])
    trans, rot = lat["trans"], lat["rot"]
    grid = np.concatenate([np.repeat(trans, len(rot), 0), np.tile(rot, (len(trans), 1))], 1)

    def lattice(dtype):
        x, c, i = [torch.from_numpy(a).to(dtype) for a in (xyz, rgb, img)]
        tt, rr = ref_utils.trim_input_loss(i, x, c, torch.from_numpy(trans).to(dtype), torch.from_numpy(rot).to(dtype), 10)
        idx = _rows_to_index(np.concatenate([tt.numpy(), rr.numpy()], 1), grid)
        table = np.array([ref_loss_grad(ref_omniloc, xyz, rgb, img, g.astype(np.float64 if dtype == torch.float64 else np.float32), dtype)[0] for g in grid])
        return idx, table
    for tag, fn in (("t1", lambda: with_threads(1, lambda: lattice(torch.float32))), ("t8", lambda: with_threads(8, lambda: lattice(torch.float32))),
                    ("f64", lambda: with_dtype(torch.float64, lambda: lattice(torch.float64)))):
        idx, table = fn()
        out["lattice_top10_" + tag], out["lattice_table_" + tag] = idx, table
    print("variants/lattice: top-10", {t: out["lattice_top10_" + t].tolist() for t in ("t1", "t8", "f64")})

    # ---- (2) trim_input_hist_secondary on rerank_small -----------------------------------------------------------
    rr = np.load(os.path.join(HERE, "rerank_small.npz"))
    xh, ch, ih, ph = rr["xyz"], synth.rgb_from_u8(rr["rgb8"]), synth.img_from_u8(rr["img8"]), rr["poses"]

    def rerank(dtype):
        # geometry in `dtype`; colours and panorama stay fp32 (make_pano writes them into an fp32 image, utils.py:186-190)
        x, c, i = torch.from_numpy(xh).to(dtype), torch.from_numpy(ch), torch.from_numpy(ih)
        t, r = ref_utils.trim_input_hist_secondary(i, x, c, torch.from_numpy(ph[:, :3].copy()).to(dtype), torch.from_numpy(ph[:, 3:].copy()).to(dtype),
                                                   len(ph), 4, 4)
        return _rows_to_index(np.concatenate([t.numpy(), r.numpy()], 1), ph)
    out["rerank_order_t1"] = with_threads(1, lambda: rerank(torch.float32))
    out["rerank_order_t8"] = with_threads(8, lambda: rerank(torch.float32))
    out["rerank_order_f64"] = with_dtype(torch.float64, lambda: rerank(torch.float64))
    print("variants/rerank: top-6", {t: out["rerank_order_" + t][:6].tolist() for t in ("t1", "t8", "f64")})

    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
    # ---- (3) the complete C1 query: make_input + six omniloc runs -------------------------------------------------
    if with_c1:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_parse_cfg", os.path.join(REF, "parse_utils.py"))
        ref_parse = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_parse)
        import localize as ref_localize
        sc1 = synth.make_scene(200000, 512, 1024, seed=3)
        cfg1 = ref_parse.parse_ini(os.path.join(REF, "configs", "stanford.ini"))
        init1 = ref_localize.get_init_dict(cfg1)

        def c1(dtype):
            x1, c1_, i1 = [torch.from_numpy(a).to(dtype) for a in (sc1.xyz, sc1.rgb, sc1.img)]
            np.random.seed(2); torch.manual_seed(2)
            in_t, in_r = ref_utils.make_input(i1, x1, c1_, cfg1.num_input, init1, cfg1.criterion, cfg1.num_intermediate)
            st_t, st_r = in_t.clone(), in_r.clone()
            res = [ref_omniloc.omniloc(i1, x1, c1_, in_t, in_r, k, cfg1, None) for k in range(cfg1.num_input)]
            return {"input_trans": st_t.numpy().astype(np.float64), "input_rot": st_r.numpy().astype(np.float64),
                    "final_t": np.stack([r[0].detach().numpy().reshape(3) for r in res]).astype(np.float64),
                    "final_R": np.stack([r[1].detach().numpy() for r in res]).astype(np.float64),
                    "final_loss": np.array([float(r[2]) for r in res]), "best": int(np.argmin([float(r[2]) for r in res]))}
        # (no fp64 variant here: make_input mixes the cloud with fp32 images inside grid_sample / index_put_, which refuse mixed dtypes)
        for tag, fn in (("t4", lambda: with_threads(4, lambda: c1(torch.float32))), ("t2", lambda: with_threads(2, lambda: c1(torch.float32)))):
            r = fn()
            for k, v in r.items():
                out["c1_" + k + "_" + tag] = v
            print("variants/c1", tag, "best", r["best"], "t", r["final_t"][r["best"]], "losses", np.round(r["final_loss"], 4).tolist(), flush=True)
            np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)


def main():
    if "variants" in sys.argv[1:]:
        ref_omniloc, ref_utils = import_reference()
        make_variants(ref_omniloc, ref_utils, with_c1="noc1" not in sys.argv[1:])
        return
    if "color" in sys.argv[1:]:
        make_color_small()
        return
    if "lattice" in sys.argv[1:]:                 # only the newest fixture (the others take minutes)
        ref_omniloc, ref_utils = import_reference()
        make_score_lattice(ref_omniloc, ref_utils)
        return
    from piccolo_b200 import synth
    ref_omniloc, ref_utils = import_reference()
    torch.manual_seed(2)
    torch.set_num_threads(8)
    Cfg = namedtuple("Cfg", ["num_input", "lr", "num_iter", "patience", "factor", "out_of_room_quantile"])

    # ---------------- fixture 1: loss + gradient, small scene with edge cases -----------------
    sc = synth.make_scene(4096, 64, 128, seed=2)
    rng = np.random.default_rng(11)
    poses = []
    gt = sc.gt_pose
    poses.append(gt)
    for _ in range(9):
        p = gt + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.25, 3)])
        poses.append(p)
    for _ in range(4):   # arbitrary poses anywhere in the room, any orientation
        p = np.concatenate([np.asarray(sc.room) * (0.15 + 0.7 * rng.random(3)), rng.random(3) * 2 * np.pi])
        poses.append(p)
    xyz = sc.xyz.copy()
    poses.append(np.concatenate([xyz[17].astype(np.float64), [0.3, 0.0, 0.0]]))           # camera exactly on a point
    poses.append(np.array([4.0, 3.0, 2.95, 1.0, 0.0, 0.0]))                               # near the ceiling: many v at the clip
    poses = np.asarray(poses, dtype=np.float32)
    rgb, img = sc.rgb, sc.img
    out = {"xyz": xyz, "rgb8": sc.rgb8, "img8": sc.img8, "poses": poses}
    l32, g32, l64, g64 = [], [], [], []
    for p in poses:
        a, b = ref_loss_grad(ref_omniloc, xyz, rgb, img, p, torch.float32); l32.append(a); g32.append(b)
        a, b = ref_loss_grad(ref_omniloc, xyz, rgb, img, p.astype(np.float64), torch.float64); l64.append(a); g64.append(b)
    out.update(loss32=np.array(l32), grad32=np.array(g32), loss64=np.array(l64), grad64=np.array(g64))
    # batched module on the first 4 poses
    cfgb = Cfg(4, 0.1, 100, 5, 0.9, 0.05)
    bmod = ref_omniloc.BatchSamplingLoss(torch.from_numpy(xyz), torch.from_numpy(rgb), torch.from_numpy(img), torch.device("cpu"), cfgb)
    pb = torch.from_numpy(poses[:4])
    tot, lst = bmod(pb[:, :3].unsqueeze(-1), pb[:, 3:4], pb[:, 4:5], pb[:, 5:6])
    out.update(batch_total=np.float32(tot.item()), batch_list=lst.numpy())
    # all-black panorama -> NaN
    a, _ = ref_loss_grad(ref_omniloc, xyz, rgb, np.zeros_like(img), poses[0], torch.float32)
    out.update(black_loss=np.float32(a))
    # quantile box
    qs = [ref_utils.quantile(torch.from_numpy(xyz[:, k]), 0.05) for k in range(3)]
    out.update(box_lo=np.array([float(q[0]) for q in qs], dtype=np.float32), box_hi=np.array([float(q[1]) for q in qs], dtype=np.float32))
    np.savez_compressed(os.path.join(HERE, "loss_small.npz"), **out)
    print("loss_small: loss32", out["loss32"][:4], "max|g32-g64|", np.nanmax(np.abs(out["grad32"] - out["grad64"])))

    # ---------------- fixture 2: grid scoring / top-K (trim_input_loss) ------------------------
    grid = synth.pose_grid(sc.room, (3, 3, 2), 8)
    trans = torch.from_numpy(np.ascontiguousarray(grid[::8, :3]))
    rot = torch.from_numpy(np.ascontiguousarray(grid[:8, 3:]))
    xyz_t, rgb_t, img_t = torch.from_numpy(xyz), torch.from_numpy(rgb), torch.from_numpy(img)
    tt_all, rr_all = ref_utils.trim_input_loss(img_t, xyz_t, rgb_t, trans, rot, len(trans) * len(rot))
    tt, rr = ref_utils.trim_input_loss(img_t, xyz_t, rgb_t, trans, rot, 10)
    table = np.array([ref_loss_grad(ref_omniloc, xyz, rgb, img, g, torch.float32)[0] for g in grid])
    np.savez_compressed(os.path.join(HERE, "score_small.npz"), trans=trans.numpy(), rot=rot.numpy(), grid=grid,
                        loss_table=table, top10_trans=tt.numpy(), top10_rot=rr.numpy(),
                        all_trans=tt_all.numpy(), all_rot=rr_all.numpy())
    print("score_small: best", table.min(), "worst", table.max())

    # ---------------- fixture 2b: histogram re-rank (trim_input_hist_secondary) ------------------
    sch = synth.make_scene(30000, 128, 256, seed=7)
    gridh = synth.pose_grid(sch.room, (3, 3, 2), 6)
    rngh = np.random.default_rng(3)
    posesh = np.concatenate([gridh[rngh.choice(len(gridh), 20, replace=False)],
                             (sch.gt_pose + np.array([0.05, -0.04, 0.02, 0.03, 0.01, -0.01]))[None].astype(np.float32),
                             (sch.gt_pose + np.array([0.4, 0.3, -0.1, 0.3, 0.0, 0.0]))[None].astype(np.float32)]).astype(np.float32)
    xh, ch, ih = torch.from_numpy(sch.xyz), torch.from_numpy(sch.rgb), torch.from_numpy(sch.img)
    tr_all, rr_all = ref_utils.trim_input_hist_secondary(ih, xh, ch, torch.from_numpy(posesh[:, :3].copy()), torch.from_numpy(posesh[:, 3:].copy()), len(posesh), 4, 4)
    tr6, rr6 = ref_utils.trim_input_hist_secondary(ih, xh, ch, torch.from_numpy(posesh[:, :3].copy()), torch.from_numpy(posesh[:, 3:].copy()), 6, 4, 4)
    Rg = ref_utils.rot_from_ypr(torch.from_numpy(posesh[20, 3:].copy()))
    pano = ref_utils.make_pano(torch.transpose(torch.matmul(Rg, torch.transpose(xh - torch.from_numpy(posesh[20, :3].copy()), 0, 1)), 0, 1), ch,
                               resolution=(128, 256), return_torch=True)
    np.savez_compressed(os.path.join(HERE, "rerank_small.npz"), xyz=sch.xyz, rgb8=sch.rgb8, img8=sch.img8, poses=posesh,
                        all_trans=tr_all.numpy(), all_rot=rr_all.numpy(), top6_trans=tr6.numpy(), top6_rot=rr6.numpy(),
                        pano20=pano.numpy())
    print("rerank_small: best pose", tr6[0].numpy(), "gt", sch.gt_pose[:3])

    # ---------------- fixture 2c: candidate grids (generate_rot_points / generate_trans_points) ------------
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_localize_cfg", os.path.join(REF, "parse_utils.py"))
    ref_parse = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_parse)
    grids = {}
    scg = synth.make_scene(50000, 32, 64, seed=2)
    xg = torch.from_numpy(scg.xyz)
    for name in ("stanford", "omniscenes"):
        cfgg = ref_parse.parse_ini(os.path.join(REF, "configs", name + ".ini"))
        import localize as ref_localize
        init = ref_localize.get_init_dict(cfgg)
        grids[name + "_rot"] = ref_utils.generate_rot_points(init).numpy()
        grids[name + "_trans"] = ref_utils.generate_trans_points(xg, init).numpy()
    np.savez_compressed(os.path.join(HERE, "grids_small.npz"), xyz=scg.xyz, **grids)
    print("grids:", {k: v.shape for k, v in grids.items()})

    # ---------------- fixture 2d: one full C1 query through the reference (make_input + omniloc) ----------
    if os.environ.get("GOLDEN_C1", "1") == "1":
        sc1 = synth.make_scene(200000, 512, 1024, seed=3)
        x1, c1, i1 = torch.from_numpy(sc1.xyz), torch.from_numpy(sc1.rgb), torch.from_numpy(sc1.img)
        cfg1 = ref_parse.parse_ini(os.path.join(REF, "configs", "stanford.ini"))
        init1 = ref_localize.get_init_dict(cfg1)
        np.random.seed(2); torch.manual_seed(2)
        in_t, in_r = ref_utils.make_input(i1, x1, c1, cfg1.num_input, init1, cfg1.criterion, cfg1.num_intermediate)
        # omniloc optimises VIEWS of input_trans / input_rot in place (omniloc.py:15-19): keep the start poses
        start_t, start_r = in_t.clone(), in_r.clone()
        res = [ref_omniloc.omniloc(i1, x1, c1, in_t, in_r, k, cfg1, None) for k in range(cfg1.num_input)]
        in_t, in_r = start_t, start_r
        best = int(np.argmin([float(r[2]) for r in res]))
        np.savez_compressed(os.path.join(HERE, "query_c1.npz"), gt_pose=sc1.gt_pose, input_trans=in_t.numpy(), input_rot=in_r.numpy(),
                            final_t=np.stack([r[0].detach().numpy().reshape(3) for r in res]), final_R=np.stack([r[1].detach().numpy() for r in res]),
                            final_loss=np.array([float(r[2]) for r in res], dtype=np.float32), best=best)
        print("query_c1: best", best, "t", res[best][0].detach().numpy().reshape(3), "gt", sc1.gt_pose[:3], "losses", [round(float(r[2]), 4) for r in res])

    # ---------------- fixture 3: refinement trajectories (omniloc / omniloc_batch) -------------
    def run_refine(scn, starts, num_iter, tag, factor, early_iter):
        """omniloc per candidate + omniloc_batch in fp32 (the parity target), the same after `early_iter`
        iterations (trajectories still correlated: tight gate), and in fp64 (the reference's own
        fp32-vs-fp64 spread bounds how reproducible the end state is)."""
        def run(dtype, n_it):
            torch.set_default_dtype(dtype)
            try:
                xyz_t, rgb_t, img_t = [torch.from_numpy(a).to(dtype) for a in (scn.xyz, scn.rgb, scn.img)]
                st = torch.from_numpy(starts).to(dtype)
                cfg = Cfg(len(starts), 0.1, n_it, 5, factor, 0.05)
                seq = [ref_omniloc.omniloc(img_t, xyz_t, rgb_t, st[:, :3].clone(), st[:, 3:].clone(), i, cfg, None) for i in range(len(starts))]
                bat = ref_omniloc.omniloc_batch(img_t, xyz_t, rgb_t, st[:, :3].clone(), st[:, 3:].clone(), cfg, None)
            finally:
                torch.set_default_dtype(torch.float32)
            return {"seq_t": np.stack([s[0].detach().numpy().reshape(3) for s in seq]),
                    "seq_R": np.stack([s[1].detach().numpy() for s in seq]),
                    "seq_loss": np.array([float(s[2]) for s in seq]),
                    "bat_t": bat[0].detach().numpy().reshape(3), "bat_R": bat[1].detach().numpy(), "bat_loss": float(bat[2])}
        res = {"xyz": scn.xyz, "rgb8": scn.rgb8, "img8": scn.img8, "starts": starts, "gt_pose": scn.gt_pose,
               "num_iter": num_iter, "factor": factor, "early_iter": early_iter}
        full32, early32, full64 = run(torch.float32, num_iter), run(torch.float32, early_iter), run(torch.float64, num_iter)
        for k, v in full32.items():
            res[k] = np.asarray(v, dtype=np.float32)
        for k, v in early32.items():
            res["early_" + k] = np.asarray(v, dtype=np.float32)
        for k, v in full64.items():
            res["f64_" + k] = np.asarray(v, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **res)
        print(tag, "seq_loss", res["seq_loss"], "bat_loss", res["bat_loss"], "t", res["bat_t"], "gt", scn.gt_pose[:3])

    rng = np.random.default_rng(5)
    starts = np.stack([gt + np.concatenate([rng.normal(0, 0.25, 3), rng.normal(0, 0.15, 3)]) for _ in range(3)]).astype(np.float32)
    starts[2, :3] = [7.9, 0.2, 2.9]   # starts outside the 5-95 % box: exercises the clamp (and the batch quirk)
    run_refine(sc, starts, 30, "refine_small", 0.8, 8)

    sc2 = synth.make_scene(50000, 256, 512, seed=3)
    rng = np.random.default_rng(6)
    starts2 = np.stack([sc2.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.12, 3)]) for _ in range(3)]).astype(np.float32)
    run_refine(sc2, starts2, 100, "refine_medium", 0.8, 20)

    # medium loss/grad vectors
    poses2 = np.concatenate([starts2, sc2.gt_pose[None].astype(np.float32)], axis=0)
    l32 = []; g32 = []; l64 = []; g64 = []
    for p in poses2:
        a, b = ref_loss_grad(ref_omniloc, sc2.xyz, sc2.rgb, sc2.img, p, torch.float32); l32.append(a); g32.append(b)
        a, b = ref_loss_grad(ref_omniloc, sc2.xyz, sc2.rgb, sc2.img, p.astype(np.float64), torch.float64); l64.append(a); g64.append(b)
    np.savez_compressed(os.path.join(HERE, "loss_medium.npz"), poses=poses2, loss32=np.array(l32), grad32=np.array(g32),
                        loss64=np.array(l64), grad64=np.array(g64))
    print("loss_medium", l32)
    make_score_lattice(ref_omniloc, ref_utils)
    make_color_small()


if __name__ == "__main__":
    main()
