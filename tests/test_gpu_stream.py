"""pipeline.localize_stream (upload + packing of the next query on a side stream) must return exactly what the
one-at-a-time host-buffer entry returns, query by query; handles created on the side stream and used on the main one are
freed safely (the library records cross-stream uses)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stream_of_queries_equals_one_at_a_time():
    from piccolo_b200 import pipeline, synth
    from piccolo_b200.utils import grid_poses
    dev = torch.device("cuda:0")
    cfg = pipeline.STANFORD_PARALLEL._replace(num_iter=20, num_intermediate=12, num_input=3)
    qs = []
    for seed in (3, 4, 5, 6):
        sc = synth.make_scene(40_000 + 1000 * seed, 128, 256, seed=seed)
        rng = np.random.default_rng(seed)
        trans = np.stack([sc.gt_pose[:3] + rng.normal(0, 0.3, 3) for _ in range(6)]).astype(np.float32)
        trans[0] = sc.gt_pose[:3]
        rot = np.zeros((8, 3), np.float32); rot[:, 0] = sc.gt_pose[3] + np.arange(8) * 2 * np.pi / 8
        grid = pipeline.StartGrid(torch.from_numpy(trans), torch.from_numpy(rot)).pin_memory()
        qs.append(tuple(torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)) + (grid,))
    one = [pipeline.localize_query_host(*q, cfg, dev) for q in qs]
    for rep in range(3):                                   # repeated: buffers are recycled between the two streams
        got = list(pipeline.localize_stream(iter(qs), cfg, dev))
        assert len(got) == len(one)
        for (p0, l0), (p1, l1) in zip(one, got):
            assert torch.equal(p0, p1) and l0 == l1
    torch.cuda.synchronize()


def test_uint8_host_buffers_give_the_float_path_bits():
    """Colours and panorama handed over as uint8 (the reference's data before its `/ 255.`) are expanded on the device to the
    correctly rounded k/255: same poses, bit for bit, as with the float32 buffers — a quarter of the upload."""
    from piccolo_b200 import pipeline, synth
    dev = torch.device("cuda:0")
    cfg = pipeline.STANFORD_PARALLEL._replace(num_iter=15, num_intermediate=12, num_input=3)
    sc = synth.make_scene(50_000, 128, 256, seed=7)
    rgb8 = np.rint(sc.rgb * 255).astype(np.uint8); img8 = np.rint(sc.img * 255).astype(np.uint8)
    assert np.array_equal(synth.rgb_from_u8(rgb8), sc.rgb) and np.array_equal(synth.img_from_u8(img8), sc.img)
    rng = np.random.default_rng(7)
    trans = np.stack([sc.gt_pose[:3] + rng.normal(0, 0.3, 3) for _ in range(6)]).astype(np.float32); trans[0] = sc.gt_pose[:3]
    rot = np.zeros((8, 3), np.float32); rot[:, 0] = sc.gt_pose[3] + np.arange(8) * 2 * np.pi / 8
    grid = pipeline.StartGrid(torch.from_numpy(trans), torch.from_numpy(rot)).pin_memory()
    f32 = tuple(torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)) + (grid,)
    u8 = (f32[0], torch.from_numpy(rgb8).pin_memory(), torch.from_numpy(img8).pin_memory(), grid)
    a = pipeline.localize_query_host(*f32, cfg, dev)
    b = pipeline.localize_query_host(*u8, cfg, dev)
    c = list(pipeline.localize_stream(iter([u8, u8]), cfg, dev))
    assert torch.equal(a[0], b[0]) and a[1] == b[1]
    for p, l in c:
        assert torch.equal(a[0], p) and a[1] == l
