"""GPU parity of the config-5 ego-depth sweep against its defining oracle (oracle/ego_depth.py)."""
import pytest
import torch

from egogen_b200 import assets

pytestmark = pytest.mark.gpu


def test_ego_depth_matches_oracle():
    from egogen_b200 import ego_depth
    from oracle import ego_depth as oed
    dev = torch.device("cuda:0")
    scene = assets.make_box_scene(4, n_boxes=3)
    sdf_cpu = assets.rasterize_scene_sdf(scene, D=64)
    sdf = {k: v.to(dev) for k, v in sdf_cpu.items()}
    g = torch.Generator().manual_seed(0)
    A = 6
    eye = torch.cat([torch.rand(A, 2, generator=g) * 4 - 2, torch.full((A, 1), 1.6)], 1)
    yaw = torch.rand(A, generator=g) * 6.28
    fwd = torch.stack([yaw.cos(), yaw.sin(), torch.full((A,), -0.1)], 1)
    fwd = fwd / fwd.norm(dim=1, keepdim=True)
    right = torch.cross(fwd, torch.tensor([0.0, 0.0, 1.0]).expand(A, 3), dim=1)
    right = right / right.norm(dim=1, keepdim=True)
    up = torch.cross(right, fwd, dim=1)
    cam = torch.cat([eye, right, up, fwd], 1)
    depth, steps = ego_depth(sdf, cam.to(dev), H=32, W=32, fx=20.0, fy=20.0, return_steps=True)
    ref, ref_steps = oed.ego_depth(sdf_cpu, cam, H=32, W=32, fx=20.0, fy=20.0)
    d, r = depth.cpu(), ref
    # rays that march the same number of steps took the same path: fp32 agreement (measured: every ray, max 6.7e-6).
    # Sphere tracing is discontinuous at silhouettes, so a ray may stop one step apart; those are the ONLY exclusions,
    # they are counted (< 0.5 %) and still bounded by one hit threshold
    same = steps.cpu() == ref_steps
    assert same.float().mean() > 0.995, same.float().mean()
    assert (d - r).abs()[same].max() < 2e-5, (d - r).abs()[same].max()
    if (~same).any():
        assert (d - r).abs()[~same].max() < 5e-3, (d - r).abs()[~same].max()
    assert d.min() >= 0 and d.max() <= 7.0
    assert (d < 6.99).float().mean() > 0.2          # the scene is actually visible
    # empty batch
    e = ego_depth(sdf, cam[:0].to(dev), H=8, W=8)
    assert e.shape == (0, 8, 8)
