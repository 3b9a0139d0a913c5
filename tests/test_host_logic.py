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
