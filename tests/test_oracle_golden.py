"""Pin the CPU oracle against golden vectors produced by the reference's own classes
(tests/golden/gen_golden.py). CPU-only."""
import os

import numpy as np
import torch

from egogen_b200.assets import fill_params_
from oracle import nets, sdf as osdf, tgm
from oracle.smplx_lbs import SMPLXParserOracle


def _t(a):
    return torch.as_tensor(np.asarray(a))


def test_calc_sdf_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "sdf_golden.npz"))
    for k in "ab":
        d = {"center": _t(g[f"center_{k}"]), "scale": _t(g[f"scale_{k}"]), "sdf": _t(g[f"grid_{k}"])}
        out = osdf.calc_sdf(_t(g[f"pts_{k}"]), d)
        assert torch.equal(out, _t(g[f"out_{k}"]))          # same ATen op => bit-identical
        out2, idx = osdf.calc_sdf_explicit(_t(g[f"pts_{k}"]), d)
        assert torch.allclose(out2, out, atol=2e-6, rtol=0)
        assert torch.equal(out2 < 0, out < 0) or ((out2 < 0) != (out < 0)).sum() <= 1
        D = torch.tensor(d["sdf"].shape)
        assert (idx >= 0).all() and (idx <= D - 1).all()


def test_predictor_regressor_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets_motion_golden.npz"))
    pred = fill_params_(nets.PredictorOracle().eval(), seed=11)
    reg = fill_params_(nets.RegressorOracle().eval(), seed=12, w_gain=0.7)
    with torch.no_grad():
        Y = pred.sample_prior(_t(g["X"]), _t(g["z"]))
        xb = reg.forward_cont(Y.reshape(-1, 201), _t(g["betas"]))
        rot = tgm.cont2rotmat(xb[:, 3:135].contiguous().view(xb.shape[0], -1, 6))
    assert torch.allclose(Y, _t(g["Y"]), atol=1e-6, rtol=1e-6)
    assert torch.allclose(xb, _t(g["xb_cont"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(rot, _t(g["rotmat"]), atol=1e-5, rtol=1e-5)


def test_policy_nets_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets_policy_golden.npz"))
    actor = fill_params_(nets.ActorOracle(), seed=21)
    critic = fill_params_(nets.CriticOracle(), seed=22)
    shared = fill_params_(nets.PolicyBaseOracle(), seed=23)
    obs = {k: _t(g[k]) for k in ("state", "egosensing", "dist", "time")}
    with torch.no_grad():
        hx = shared(obs)
        mu, logvar = actor(hx)
        val = critic(hx)
    assert torch.allclose(hx, _t(g["hx"]), atol=1e-6, rtol=1e-6)
    assert torch.allclose(mu, _t(g["mu"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(logvar, _t(g["logvar"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(val, _t(g["value"]), atol=1e-5, rtol=1e-5)


def test_policy_param_counts():
    a, c, s = nets.init_policy_nets(0)
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert (n(a), n(c), n(s)) == (5608192, 5314177, 2245632)
    p, r = nets.PredictorOracle(), nets.RegressorOracle()
    assert (n(p), n(r)) == (2455497, 398239)


def test_coordinate_extractor_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "coord_golden.npz"))
    R, T = SMPLXParserOracle.new_coordinate_from_joints(_t(g["jts"]))
    assert torch.allclose(R, _t(g["R"]), atol=1e-7) and torch.equal(T, _t(g["T"]))
    R2, T2 = SMPLXParserOracle.new_coordinate_from_joints(_t(g["jts_loco"]))
    assert torch.allclose(R2, _t(g["R_loco"]), atol=1e-7)
    # model-free invariant of the reference fixture subseq_00343.npz:
    assert torch.allclose(R2[0], torch.eye(3), atol=1e-3) and T2.abs().max() < 1e-3
