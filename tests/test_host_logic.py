"""CPU-only host logic: entrypoint flags, sharding arithmetic, rollout writer schema, asset loaders,
world_size-2 gloo allreduce semantics used by the PPO update."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_main_ppo_flags_match_reference_defaults():
    from egogen_b200.main_ppo import get_args, shard
    a = get_args([])
    # defaults of motion/crowd_ppo/main_ppo.py:40-92
    assert (a.seed, a.buffer_size, a.lr, a.gamma, a.epoch) == (0, 4096, 3e-4, 0.99, 3000)
    assert (a.step_per_epoch, a.step_per_collect, a.repeat_per_collect, a.batch_size) == (20000, 1024, 1, 256)
    assert (a.training_num, a.test_num, a.vf_coef, a.ent_coef, a.gae_lambda) == (256, 20, 1.0, 0.01, 0.95)
    assert (a.max_grad_norm, a.eps_clip, a.norm_adv, a.save_interval) == (0.1, 0.1, 1, 2)
    assert not a.watch and not a.finetune and a.resume_path is None
    assert shard(256, 8, 3, "training-num") == 32
    with pytest.raises(ValueError):
        shard(250, 8, 0, "training-num")


def test_rollout_writer_schema(tmp_path):
    from egogen_b200.utils import save_rollout_results, MP_KEYS
    mp = [torch.zeros(1, 20, 67, 3), torch.zeros(1, 20, 93), torch.zeros(10), "male", torch.eye(3), torch.zeros(1, 3),
          torch.zeros(1, 20, 3), "2-frame"]
    p = save_rollout_results({"wpath": torch.zeros(2, 3), "navmesh_path": "x.ply", "scene_path": "s.ply"}, [mp, mp],
                             str(tmp_path), man_id="t")
    d = pickle.load(open(p, "rb"))
    assert set(d) == {"motion", "wpath", "navmesh_path", "scene_path"} and len(d["motion"]) == 2
    m = d["motion"][0]
    assert list(m) == MP_KEYS and m["blended_marker"].shape == (20, 67, 3) and m["smplx_params"].shape == (1, 20, 93)
    assert m["pelvis_loc"].shape == (20, 3) and m["mp_type"] == "2-frame"


def synthetic_rollout():
    """deterministic (scene, outmps) in the layout CrowdEnv.step appends (crowd_env_2f.py:156), the 4x duplicated batch
    included; shared with tests/golden/gen_rollout_golden.py, which feeds it to the reference's own writer"""
    g = torch.Generator().manual_seed(12)
    r = lambda *s: torch.randn(*s, generator=g)
    outmps = [[r(4, 20, 67, 3), r(4, 20, 93), r(10), "male", r(3, 3), r(1, 3), r(4, 20, 3), "2-frame"] for _ in range(3)]
    scene = {"wpath": r(2, 3), "navmesh_path": "scenes/room_0/navmesh_tight.ply", "scene_path": "scenes/room_0/mesh.ply",
             "obj_id": 7, "obj_transform": np.eye(4)}
    return scene, outmps


def test_rollout_writer_matches_reference_writer(tmp_path, golden_dir):
    """egogen_b200.utils.save_rollout_results vs the pickle the reference's own save_rollout_results wrote for the same
    rollout (tests/golden/gen_rollout_golden.py): same keys in the same order, same values, same file name rule."""
    import os
    from egogen_b200.utils import save_rollout_results
    scene, outmps = synthetic_rollout()
    p = save_rollout_results(scene, outmps, str(tmp_path / "out"), man_id="golden7")
    assert os.path.basename(p) == "motion_golden7.pkl"
    mine = pickle.load(open(p, "rb"))
    ref = pickle.load(open(os.path.join(golden_dir, "rollout_golden.pkl"), "rb"))

    def same(a, b):
        if isinstance(a, dict):
            return isinstance(b, dict) and list(a) == list(b) and all(same(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return type(a) is type(b) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(a, np.ndarray):
            return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        return a == b
    assert same(ref, mine)
    assert ref["motion"][0]["smplx_params"].shape == (1, 20, 93) and isinstance(ref["obj_transform"], tuple)


def test_surrogate_assets_have_real_shapes(smplx_model):
    from egogen_b200 import assets
    m = smplx_model
    assert m["v_template"].shape == (10475, 3) and m["shapedirs"].shape == (10475, 3, 20)
    assert m["posedirs"].shape == (486, 31425) and m["J_regressor"].shape == (55, 10475)
    assert m["lbs_weights"].shape == (10475, 55) and np.allclose(m["lbs_weights"].sum(1), 1, atol=1e-5)
    assert len(assets.marker_ids()) == 67 and len(assets.feet_vids()) == 502 and len(assets.feet_marker_idx()) == 6
    assert max(assets.marker_ids()) < 10475
    scene = assets.make_box_scene(0)
    sdf = assets.rasterize_scene_sdf(scene, D=32)
    assert sdf["sdf"].shape == (32, 32, 32)
    segs = assets.rings_to_segments(assets.scene_polygon(scene))
    assert segs.shape == (8, 4)


def test_reference_polygon_fixture_parses():
    """The reference's replica_room0_shapely.pkl (WKB) parses without shapely: 1 exterior + 5 holes."""
    p = "/root/reference/motion/data/replica_room0_shapely.pkl"
    if not os.path.exists(p):
        pytest.skip("reference tree not present on this box")
    from egogen_b200.assets import load_wkb_polygon
    rings = load_wkb_polygon(p)
    assert [len(r) for r in rings] == [47, 13, 10, 9, 9, 7]


_GLOO_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# gradient allreduce (sum of per-rank gradients scaled by 1/global_batch) == single-process mean gradient
torch.manual_seed(0)
X = torch.randn(64, 8); y = torch.randn(64); w = torch.zeros(8, requires_grad=True)
loss = ((X @ w - y) ** 2).mean(); loss.backward(); g_ref = w.grad.clone()
xs, ys = X.chunk(world)[rank], y.chunk(world)[rank]
w2 = torch.zeros(8, requires_grad=True)
(((xs @ w2 - ys) ** 2).sum() / 64).backward()
g = w2.grad.clone(); dist.all_reduce(g)
assert torch.allclose(g, g_ref, atol=1e-6)
# global advantage moments {sum, sumsq, count} -> mean / unbiased std of the whole minibatch
adv = torch.randn(64, dtype=torch.float64); a = adv.chunk(world)[rank]
m = torch.tensor([a.sum(), (a * a).sum(), float(a.numel())], dtype=torch.float64); dist.all_reduce(m)
mean = m[0] / m[2]; var = (m[1] - m[2] * mean * mean) / (m[2] - 1)
assert abs(mean - adv.mean()) < 1e-12 and abs(var.sqrt() - adv.std()) < 1e-12
from egogen_b200.main_ppo import shard
assert shard(256, world, rank, "training-num") * world == 256
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_gradient_and_moment_allreduce(tmp_path):
    script = tmp_path / "gloo_check.py"
    script.write_text(_GLOO_SCRIPT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_crowd_holes_oracle_known_answers():
    """f-2 oracle pieces on hand-checkable geometry (crowd_env_crowd_eval.py:742-822): a rectangle straight ahead is hit
    at its near edge, an eye inside a rectangle reads 0, and map cells under a rectangle turn unwalkable."""
    import numpy as np
    import torch
    from oracle.env import egosensing, get_map
    floor = np.array([[4, 4, 4, -4], [4, -4, -4, -4], [-4, -4, -4, 4], [-4, 4, 4, 4]], np.float64)
    j = torch.zeros(1, 2, 127, 3)
    j[:, :, 23] = torch.tensor([0.0, 0.1, 1.6]); j[:, :, 24] = torch.tensor([0.0, -0.1, 1.6])       # eyes -> eye_2d = origin
    j[:, :, 57] = torch.tensor([1.0, 0.1, 1.6]); j[:, :, 56] = torch.tensor([1.0, -0.1, 1.6])       # gaze = +x
    free = egosensing(j, floor)
    holes = torch.tensor([[[2.0, -1.0, 3.0, 1.0]]])
    blocked = egosensing(j, floor, holes=holes)
    ang = np.linspace(-np.pi / 2, np.pi / 2, 32)
    mid = int(np.argmin(np.abs(ang)))                       # the ray closest to straight ahead
    d_free = (free[0, 0, mid] + 1) / 2 * 7
    d_hit = (blocked[0, 0, mid] + 1) / 2 * 7
    assert abs(float(d_free) - 4.0 / np.cos(ang[mid])) < 1e-4
    assert abs(float(d_hit) - 2.0 / np.cos(ang[mid])) < 1e-4
    assert float(blocked[0, 0, 0]) == float(free[0, 0, 0])  # the sideways rays do not see the rectangle
    inside = egosensing(j, floor, holes=torch.tensor([[[-0.5, -0.5, 0.5, 0.5]]]))
    assert torch.all(inside == -1.0)                        # eye inside a hole: off the polygon, distance 0
    fl = np.asarray([[4, 4], [4, -4], [-4, -4], [-4, 4]], np.float32)
    tris = np.stack([fl[[0, 1, 2]], fl[[2, 3, 0]]])
    R, T = torch.eye(3)[None], torch.zeros(1, 1, 3)
    _, m0 = get_map(tris, R, T)
    _, m1 = get_map(tris, R, T, holes=torch.tensor([[[0.0, 0.0, 0.5, 0.5]]]))
    assert bool((m0 == 1).all()) and int((m1 == -1).sum()) == 25          # 5 x 5 grid points of the 16 x 16 @ 0.8 m map
