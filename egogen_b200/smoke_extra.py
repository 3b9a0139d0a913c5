"""Extra smoke stages (called by __graft_entry__.smoke): one vector env step on cuda:0."""
import torch

from . import assets


def run(dev):
    from .crowd_env import BoxSceneSampler, CrowdVectorEnv, default_cfg
    from .models_gamma_primitive import GAMMAPrimitiveComboGenOP, load_vposer
    from .smplx_parser import get_lbs_model
    lbs = get_lbs_model("male", dev, marker_vids=assets.marker_ids())
    genop = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": dev.index or 0})
    genop.build_model(seed=0)
    vposer, _ = load_vposer(seed=0, device=dev)
    scene = assets.make_box_scene(0)
    sdf = {k: v.to(dev) for k, v in assets.rasterize_scene_sdf(scene, D=64).items()}
    venv = CrowdVectorEnv(default_cfg(), genop.model, lbs, vposer, sdf, assets.scene_polygon(scene),
                          BoxSceneSampler(sdf, lbs, dev, seed=0), 4, dev)
    venv.reset()
    obs, rew, term, _, _ = venv.step(torch.zeros(4, 128, device=dev))
    torch.cuda.synchronize(dev)
    assert torch.isfinite(rew).all() and torch.isfinite(obs["state"]).all()
    print("smoke env step ok: reward", [round(float(r), 4) for r in rew])
    # one 64-env step (tcgen05 decode cluster + regressor, tensor-core dense layers, fused LBS + SDF) followed by a device-side
    # restart from the pool; returns the motion model's outputs for a fixed input so the caller (the smoke test) can check
    # them against its oracle
    venv64 = CrowdVectorEnv(default_cfg(), genop.model, lbs, vposer, sdf, assets.scene_polygon(scene),
                            BoxSceneSampler(sdf, lbs, dev, seed=1), 64, dev)
    venv64.reset()
    g = torch.Generator().manual_seed(0)
    X = torch.randn(2, 64, 201, generator=g) * 0.3
    z = torch.randn(64, 128, generator=g)
    betas = torch.zeros(18, 64, 10)
    Y, Yb = genop.model.sample_prior(X.to(dev), betas.to(dev), z.to(dev))
    _, rew, term, _, _ = venv64.step(z.to(dev))
    venv64.reset_masked(term)
    torch.cuda.synchronize(dev)
    assert torch.isfinite(rew).all()
    print(f"smoke 64-env step ok: restarted {int(term.sum())} envs")
    return dict(genop=genop, X=X, z=z, betas=betas, Y=Y.cpu(), Yb=Yb.cpu())
