"""CPU tests of the host-side mirror of the reference interface: config parsing, colour preprocessing,
candidate grids, the torch make_pano used for result images."""
import os

import numpy as np
import pytest
import torch

from piccolo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs_parse_to_reference_values():
    from piccolo_b200.parse_utils import apply_override, parse_ini, parse_override, parse_value, save_effective_config
    cfg = parse_ini(os.path.join(ROOT, "configs", "stanford.ini"))
    assert cfg.dataset == "Stanford2D-3D-S" and cfg.num_trans == 50 and cfg.num_yaw == 4 and cfg.factor == 0.8
    assert cfg.area is None and cfg.sharpen_color is True and cfg.visualize is False and cfg.criterion == "loss_histogram"
    par = parse_ini(os.path.join(ROOT, "configs", "stanford_parallel.ini"))
    assert par.parallel is True and par.sample_rate == 6
    omni = parse_ini(os.path.join(ROOT, "configs", "omniscenes.ini"))
    assert omni.xy_only and omni.yaw_only and omni.z_prior == 1.5 and omni.num_trans == 150 and omni.match_color
    assert parse_value("1e-3") == 1e-3 and parse_value("None") is None and parse_value("a, b") == ["a", "b"]
    ov = parse_override("lr=0.05,room_name=a,b,num_iter=10")
    assert ov == {"lr": 0.05, "room_name": ["a", "b"], "num_iter": 10}
    cfg2 = apply_override(cfg, ov)
    assert cfg2.lr == 0.05 and cfg2.num_iter == 10 and cfg2.room_name == ["a", "b"] and cfg2.patience == 5
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        save_effective_config(cfg2, os.path.join(d, "config.ini"))
        back = parse_ini(os.path.join(d, "config.ini"))
        assert back.lr == 0.05 and back.num_iter == 10 and back.dataset == cfg.dataset and back.area is None


def test_colour_preprocessing_matches_reference(golden):
    """The CPU restatement (oracle/color_oracle.py: test infrastructure; the product has no CPU path) against the outputs of the
    unmodified reference.  The device path is compared with both in tests/test_gpu_parity.py."""
    from oracle.color_oracle import color_match_np as color_match, color_mod_np as color_mod
    g = golden("color_small")
    img = torch.from_numpy(synth.img_from_u8(g["img8"]))
    rgb = torch.from_numpy(synth.rgb_from_u8(g["rgb8"]))
    a_img, a_rgb = color_mod(img.clone(), rgb.clone(), 256)
    np.testing.assert_array_equal(a_img.numpy(), g["mod_img"])
    np.testing.assert_array_equal(a_rgb.numpy(), g["mod_rgb"])
    m = color_match(img.clone(), rgb.clone())
    np.testing.assert_allclose(m.numpy(), g["match_img"], atol=2e-6)
    assert ((255 * m.numpy()).astype(np.uint8) != (255 * g["match_img"]).astype(np.uint8)).mean() < 1e-3


def test_candidate_grids():
    from piccolo_b200.localize import get_init_dict
    from piccolo_b200.parse_utils import parse_ini
    from piccolo_b200.utils import adaptive_trans_num, generate_rot_points, generate_trans_points, grid_poses
    sc = synth.make_scene(50_000, 32, 64, seed=2)
    xyz = torch.from_numpy(sc.xyz)
    init = get_init_dict(parse_ini(os.path.join(ROOT, "configs", "stanford.ini")))
    rot = generate_rot_points(init)
    assert rot.shape == (24, 3)                                            # 24 of the 64 Euler triples are distinct rotations
    assert adaptive_trans_num(xyz, 50) == (5, 5, 3)                        # 8 x 6 x 3 m room
    trans = generate_trans_points(xyz, init)
    assert trans.shape == (75, 3)
    lo, hi = torch.quantile(xyz, 0.1, dim=0), torch.quantile(xyz, 0.9, dim=0)
    assert (trans >= lo - 1e-4).all() and (trans <= hi + 1e-4).all()
    poses = grid_poses(trans, rot)
    assert poses.shape == (1800, 6) and torch.equal(poses[25, :3], trans[1]) and torch.equal(poses[25, 3:], rot[1])
    omni = get_init_dict(parse_ini(os.path.join(ROOT, "configs", "omniscenes.ini")))
    t2, r2 = generate_trans_points(xyz, omni), generate_rot_points(omni)
    assert r2.shape == (8, 3) and (r2[:, 1:] == 0).all() and (t2[:, 2] == 1.5).all() and t2.shape[0] >= 150


def test_make_pano_matches_oracle_painter_order():
    from oracle import piccolo_oracle as orc
    from piccolo_b200.utils import make_pano
    sc = synth.make_scene(20_000, 32, 64, seed=4)
    pose = sc.gt_pose.astype(np.float32)
    R = orc.rot_and_derivs_np(pose[3:6], np.float32)[0]
    q = ((sc.xyz - pose[None, :3]) @ R.T).astype(np.float32)
    ours = make_pano(torch.from_numpy(q), torch.from_numpy(sc.rgb), resolution=(64, 128), return_torch=True).numpy()
    ref = orc.make_pano_np(q, sc.rgb, 64, 128)
    np.testing.assert_array_equal(ours.sum(2) > 0, ref.sum(2) > 0)
    assert (np.abs(ours - ref).max(axis=2) > 0).mean() < 5e-3             # atan2 ulps at pixel-truncation boundaries


def test_candidate_grids_match_reference(golden):
    """generate_rot_points / generate_trans_points against the reference's outputs for both shipped protocols."""
    from piccolo_b200.localize import get_init_dict
    from piccolo_b200.parse_utils import parse_ini
    from piccolo_b200.utils import generate_rot_points, generate_trans_points
    g = golden("grids_small")
    xyz = torch.from_numpy(g["xyz"])
    for name in ("stanford", "omniscenes"):
        init = get_init_dict(parse_ini(os.path.join(ROOT, "configs", name + ".ini")))
        trans = generate_trans_points(xyz, init).numpy()
        np.testing.assert_allclose(trans, g[name + "_trans"], rtol=0, atol=1e-6)      # same lattice, same order
        rot = generate_rot_points(init).numpy()
        ref = g[name + "_rot"]
        assert rot.shape == ref.shape
        # the reference's order comes out of a python set (PYTHONHASHSEED dependent): compare as sets of rotations
        key = lambda a: sorted(tuple(np.round(r, 5)) for r in a)
        assert key(rot) == key(ref)


def test_integer_ycrcb_matches_cv2():
    """The fixed-point 8-bit RGB <-> YCrCb conversion that pcl_color.cu restates in integers, against cv2.cvtColor
    (the reference's color_mod goes through it, color_utils.py:30-33, :51, :62) on 2^22 triples incl. all extremes."""
    import cv2

    def rgb2ycc(u8):
        r, g, b = [u8[..., i].astype(np.int32) for i in range(3)]
        y = (r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14
        cr = ((r - y) * 11682 + (128 << 14) + (1 << 13)) >> 14
        cb = ((b - y) * 9241 + (128 << 14) + (1 << 13)) >> 14
        return np.stack([np.clip(y, 0, 255), np.clip(cr, 0, 255), np.clip(cb, 0, 255)], -1).astype(np.uint8)

    def ycc2rgb(u8):
        y, cr, cb = [u8[..., i].astype(np.int32) for i in range(3)]
        b = y + (((cb - 128) * 29049 + (1 << 13)) >> 14)
        g = y + (((cb - 128) * -5636 + (cr - 128) * -11698 + (1 << 13)) >> 14)
        r = y + (((cr - 128) * 22987 + (1 << 13)) >> 14)
        return np.stack([np.clip(r, 0, 255), np.clip(g, 0, 255), np.clip(b, 0, 255)], -1).astype(np.uint8)

    rng = np.random.default_rng(0)
    tri = rng.integers(0, 256, (1, 1 << 22, 3), dtype=np.uint8)
    edge = np.array([[a, b, c] for a in (0, 1, 127, 128, 254, 255) for b in (0, 1, 127, 128, 254, 255) for c in (0, 1, 127, 128, 254, 255)], np.uint8)
    tri[0, :len(edge)] = edge
    np.testing.assert_array_equal(cv2.cvtColor(tri, cv2.COLOR_RGB2YCR_CB), rgb2ycc(tri))
    np.testing.assert_array_equal(cv2.cvtColor(tri, cv2.COLOR_YCR_CB2RGB), ycc2rgb(tri))


def test_start_grid_indexing_matches_flat_pose_list():
    """pipeline.StartGrid: pose index i*R+j (utils.py:484-485), row slices for sharding, selection by flat index."""
    from piccolo_b200.pipeline import StartGrid
    from piccolo_b200.utils import grid_poses
    rng = np.random.default_rng(0)
    trans, rot = torch.from_numpy(rng.normal(size=(7, 3)).astype(np.float32)), torch.from_numpy(rng.normal(size=(5, 3)).astype(np.float32))
    g = StartGrid(trans, rot)
    flat = grid_poses(trans, rot)
    assert len(g) == 35 and torch.equal(g.poses(), flat)
    idx = torch.tensor([0, 34, 12, 5, 29])
    assert torch.equal(g.index_select(0, idx), flat.index_select(0, idx))
    assert torch.equal(g.rows(2, 5).poses(), flat[2 * 5: 5 * 5])
    assert g.to("cpu").trans.data_ptr() == g.trans.data_ptr()


def test_requantize_is_the_drivers_uint8_round_trip():
    from piccolo_b200.color_utils import requantize
    x = torch.rand(64, 128, 3)
    want = torch.from_numpy((255 * x.numpy()).astype(np.uint8)).float() / 255.
    assert torch.equal(requantize(x), want)


def test_dataset_provider_never_ignores_dataset_keys(tmp_path, monkeypatch):
    """The drivers run on synthetic rooms only (the dataset file readers are out of scope): keys that select dataset
    files, or dataset files on disk, must raise instead of being silently ignored; synthetic records say so by name."""
    from types import SimpleNamespace
    import pytest
    from piccolo_b200 import datasets
    ok = SimpleNamespace(area=None, room_name=None, gravity_aligned=True, visualize=False, split_name="extreme")
    datasets.check_config(ok, "Stanford2D-3D-S")
    for key, val in (("area", 3), ("room_name", "office_1"), ("scene_number", 2), ("split_name", "change_handheld"),
                     ("gravity_aligned", False), ("visualize", True)):
        bad = SimpleNamespace(**{key: val})
        with pytest.raises(datasets.DatasetUnavailable):
            datasets.check_config(bad, "OmniScenes")
    monkeypatch.chdir(tmp_path)
    (tmp_path / "data" / "stanford").mkdir(parents=True)
    with pytest.raises(datasets.DatasetUnavailable):
        datasets.check_config(ok, "Stanford2D-3D-S")
    datasets.check_config(SimpleNamespace(synthetic=True), "Stanford2D-3D-S")
    cfg = SimpleNamespace(synthetic=True, synthetic_points=2000, synthetic_queries=1, synthetic_height=32)
    q = next(iter(datasets.queries(cfg, "Stanford2D-3D-S")))
    assert q.filename.startswith("synthetic://") and q.pcd_name.startswith("synthetic://")


def test_colour_preprocessing_has_no_cpu_path():
    from piccolo_b200 import _lib
    from piccolo_b200.color_utils import color_match, color_mod
    img, rgb = torch.zeros(4, 8, 3), torch.zeros(5, 3)
    with pytest.raises(_lib.PiccoloError):
        color_mod(img, rgb, 256)
    with pytest.raises(_lib.PiccoloError):
        color_match(img, rgb)
