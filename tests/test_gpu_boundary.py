"""Drop-in boundary of the crowd_ppo path on the GPU: the tianshou vector-env / policy surface the
reference binds (dummy_vector_env.py:33-45,78-84, ppo_policy.py:93-179), the rollout pickle written at episode end
(crowd_env_2f.py:154-155,305-309) and the entrypoint's asset wiring (main_ppo.py:246-304, primitive_model.py:56-96)."""
import glob
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def w8(dev):
    from egogen_b200.runtime import build_world
    return build_world(dev, 8, seed=5, sdf_res=64, capture_rollout=True)


def test_vector_env_step_subset_and_attrs(dev, w8):
    """step(action, id): only the named envs advance, results follow the order of `id`, info carries env_id;
    get_env_attr / set_env_attr read and write per-env rows (tianshou BaseVectorEnv surface)."""
    venv = w8["venv"]
    venv.reset()
    g = torch.Generator().manual_seed(0)
    z = torch.randn(8, 128, generator=g).to(dev)
    before = {k: venv.buf[k].clone() for k in ("state", "seed", "R0", "T0", "steps", "dist")}
    obs, rew, term, trunc, info = venv.step(z[[5, 2]], id=[5, 2])
    assert obs["state"].shape == (2, 2, 402) and rew.shape == (2,) and [i["env_id"] for i in info] == [5, 2]
    for e in range(8):
        moved = e in (2, 5)
        assert int(venv.buf["steps"][e]) == (1 if moved else 0)
        assert torch.equal(venv.buf["state"][e], before["state"][e]) != moved
    # the subset result equals a full step of the same actions from the same state
    after = {k: venv.buf[k].clone() for k in before}
    for k, v in before.items():
        venv.buf[k].copy_(v)
    obs_f, rew_f, term_f, _, info_f = venv.step(z)
    assert [i["env_id"] for i in info_f] == list(range(8))
    assert torch.equal(obs_f["state"][5], obs["state"][0]) and torch.equal(obs_f["state"][2], obs["state"][1])
    assert torch.equal(rew_f[[5, 2]], rew) and torch.equal(after["seed"][[5, 2]], venv.buf["seed"][[5, 2]])
    assert len(venv.get_env_attr("dist")) == 8 and torch.equal(venv.get_env_attr("dist", [3])[0], venv.buf["dist"][3])
    venv.set_env_attr("goal", torch.tensor([1.0, 2.0, 0.9]), [4])
    assert torch.allclose(venv.buf["goal"][4].cpu(), torch.tensor([1.0, 2.0, 0.9]))
    assert venv.get_env_attr("finetuning") == [False] * 8


def test_policy_dist_and_process_fn(dev, w8):
    """forward(...).dist is the reference's Independent(Normal(mu, sigma), 1); process_fn (ppo_policy.py:93-140) on the
    transitions of a collect reproduces the collector's own returns / advantages / old log-probabilities."""
    from egogen_b200.ppo_policy import Batch
    venv, pol, col = w8["venv"], w8["policy"], w8["collector"]
    pol.train()
    col.reset()
    out = pol.forward(Batch(obs=venv.observation()))
    d = out.dist
    assert torch.allclose(d.log_prob(out.act), out.logp, atol=2e-4, rtol=1e-5)
    assert torch.allclose(d.mean, out.z_mu) and d.entropy().shape == (8,)
    T = 4
    batch, _ = col.collect(8 * T)
    b = col.buf
    fl = lambda x: x.transpose(0, 1).reshape(8 * T, *x.shape[2:]).contiguous()
    obs = {"state": b.state, "egosensing": b.ego, "dist": b.dist, "time": b.time}
    last = venv.observation()
    nxt = {k: torch.cat([v[1:], last[kk].reshape(1, *v.shape[1:])]) for (k, v), kk in zip(obs.items(), ("state", "egosensing", "dist", "time"))}
    tb = Batch(obs={k: fl(v) for k, v in obs.items()}, obs_next={k: fl(v) for k, v in nxt.items()}, act=fl(b.act), rew=fl(b.rew),
               terminated=fl(b.term), truncated=torch.zeros(8 * T, dtype=torch.uint8, device=dev))
    from types import SimpleNamespace
    res = pol.process_fn(tb, SimpleNamespace(E=8), np.arange(8 * T))
    assert torch.allclose(res.v_s, batch.v_s, atol=1e-5)
    assert torch.allclose(res.logp_old, batch.logp_old, atol=5e-4, rtol=1e-5)
    # obs_next of a terminated step is the restarted env's observation here, but tianshou masks v(obs_next) there
    assert torch.allclose(res.returns, batch.returns, atol=1e-4) and torch.allclose(res.adv, batch.adv, atol=1e-4)


def test_single_env_writes_rollout_pickle(dev, w8, tmp_path):
    """CrowdEnv(save_rollout=True) writes the reference's rollout pickle when the episode ends; every primitive carries
    the canonical frame it was GENERATED in (the outmps.append at :155 precedes the frame update at :247)."""
    from types import SimpleNamespace
    from egogen_b200 import SMPLXParser, assets
    from egogen_b200.crowd_env import CrowdEnv
    parser = SimpleNamespace(device=dev, bm_male=w8["lbs"])
    init_env = [w8["cfg"], w8["genop"], w8["genop"], None, w8["sampler"], parser, parser, parser, assets.feet_marker_idx(),
                assets.marker_ids(), w8["vposer"], w8["scene_sdf"]]
    env = CrowdEnv(init_env, save_rollout=True)
    env.rollout_dir = str(tmp_path / "eval_results")
    env.seed(1)
    env.reset()
    frames = []
    term = False
    while not term:
        frames.append((env._venv.buf["R0"][0].clone(), env._venv.buf["T0"][0].clone()))
        _, _, term, _, _ = env.step(np.zeros(128, dtype=np.float32))
    files = glob.glob(os.path.join(env.rollout_dir, "motion_*.pkl"))
    assert len(files) == 1 and files[0] == env.last_rollout_file
    node = pickle.load(open(files[0], "rb"))
    assert set(node) >= {"motion", "wpath", "navmesh_path"} and len(node["motion"]) == len(frames)
    assert node["wpath"].shape == (2, 3)
    for mp, (R0, T0) in zip(node["motion"], frames):
        assert list(mp) == ["blended_marker", "smplx_params", "betas", "gender", "transf_rotmat", "transf_transl", "pelvis_loc", "mp_type"]
        assert mp["blended_marker"].shape == (20, 67, 3) and mp["smplx_params"].shape == (1, 20, 93)
        assert np.allclose(mp["transf_rotmat"], R0.cpu().numpy()) and np.allclose(mp["transf_transl"].reshape(3), T0.cpu().numpy())
    assert not np.allclose(node["motion"][0]["transf_transl"], node["motion"][-1]["transf_transl"])


def test_main_ppo_cli_with_checkpoint_tree(dev, tmp_path, monkeypatch):
    """A checkpoint tree in the reference's layout on disk (results/crowd_ppo/<cfg>/checkpoints/epoch-*.ckp, a VPoser
    snapshot, a policy checkpoint) goes through the CLI: --watch loads all of it, evaluates and writes rollout pickles;
    a named directory without its checkpoint is an error, not a silent fall-back to synthetic weights."""
    from egogen_b200 import assets, main_ppo
    from egogen_b200.models_gamma_primitive import GAMMAPrimitiveCombo, VPoserEncoder
    from egogen_b200.runtime import build_policy, motion_checkpoint_dirs
    from egogen_b200.crowd_env import default_cfg
    root = tmp_path / "results" / "crowd_ppo"
    pdir, rdir = motion_checkpoint_dirs(str(root))
    os.makedirs(pdir); os.makedirs(rdir)
    combo = GAMMAPrimitiveCombo()
    assets.fill_params_(combo.predictor, seed=77); assets.fill_params_(combo.regressor, seed=78, w_gain=0.7)
    with torch.no_grad():
        combo.predictor.d_out.weight.mul_(0.02); combo.predictor.d_out.bias.mul_(0.02); combo.regressor.pnet.out_fc.weight.mul_(0.3)
    torch.save({"model_state_dict": combo.predictor.state_dict()}, os.path.join(pdir, "epoch-400.ckp"))
    torch.save({"model_state_dict": combo.regressor.state_dict()}, os.path.join(rdir, "epoch-100.ckp"))
    vdir = tmp_path / "vposer_v1_0"
    os.makedirs(vdir / "snapshots")
    vp = assets.fill_params_(VPoserEncoder(), seed=79)
    torch.save(vp.state_dict(), vdir / "snapshots" / "TR00_E096.pt")
    pol, _ = build_policy(default_cfg(), dev)
    torch.save({"model": pol.state_dict()}, tmp_path / "checkpoint_87.pth")
    logdir = tmp_path / "log"
    argv = ["--watch", "--resume-path", str(tmp_path / "checkpoint_87.pth"), "--test-num", "4", "--training-num", "4",
            "--step-per-collect", "16", "--batch-size", "4", "--sdf-res", "64", "--logdir", str(logdir),
            "--motion-results-root", str(root), "--vposer-dir", str(vdir), "--deterministic-eval"]
    captured = {}
    import egogen_b200.runtime as rt
    real = rt.build_world

    def spy(*a, **k):
        w = real(*a, **k)
        captured.setdefault("worlds", []).append(w)
        return w
    monkeypatch.setattr(rt, "build_world", spy)
    main_ppo.main(main_ppo.get_args(argv))
    w = captured["worlds"][0]
    assert w["genop"].weights != "synthetic" and w["genop"].weights["predictor"].endswith("epoch-400.ckp")
    sd = w["genop"].model.predictor.state_dict()
    assert torch.equal(sd["d_out.weight"].cpu(), combo.predictor.state_dict()["d_out.weight"])
    assert torch.equal(w["vposer"].bodyprior_enc_fc1.weight.cpu(), vp.bodyprior_enc_fc1.weight)
    assert torch.equal(w["policy"].state_dict()["actor.pnet.out_fc.weight"].cpu(), pol.state_dict()["actor.pnet.out_fc.weight"].cpu())
    pk = glob.glob(os.path.join(str(logdir), "eval_results", "motion_*.pkl"))
    assert len(pk) >= 4, "every evaluation episode writes a rollout pickle"
    node = pickle.load(open(pk[0], "rb"))
    assert len(node["motion"]) >= 1 and node["motion"][0]["smplx_params"].shape == (1, 20, 93)
    # a named root without checkpoints must raise
    with pytest.raises(FileNotFoundError):
        main_ppo.main(main_ppo.get_args(argv[:-5] + ["--motion-results-root", str(tmp_path / "nowhere"), "--deterministic-eval"]))


def test_host_boundary_collect_matches_device_collect(dev):
    """bench.py's e2e path: the collector with the reference's host-side data path (pinned host buffers, two host
    synchronisations per vector step) produces the transitions of the device-resident collector bit for bit from the same
    seeds, and counts the bytes it moved."""
    from egogen_b200.collector import Collector
    from egogen_b200.runtime import build_world
    outs = []
    for host in (False, True):
        torch.manual_seed(11); np.random.seed(11)
        w = build_world(dev, 16, seed=7, sdf_res=64)
        w["policy"].train()
        col = Collector(w["policy"], w["venv"], host_boundary=host)
        col.reset()
        torch.manual_seed(12)
        n_ep = 0
        for _ in range(4):                                   # 16 vector steps: 13-step episodes end and restart inside
            batch, st = col.collect(16 * 4)
            n_ep += int(st["n/ep"])
        outs.append((batch, n_ep, col))
    (b0, e0, _), (b1, e1, c1) = outs
    assert e0 == e1 and e0 > 0
    for k in ("state", "egosensing", "dist", "time"):
        assert torch.equal(b0.obs[k], b1.obs[k]), k
    assert torch.equal(b0.act, b1.act) and torch.equal(b0.returns, b1.returns) and torch.equal(b0.adv, b1.adv)
    per_step = 16 * (2 * 402 + 2 * 32 + 1 + 1) * 4
    assert c1.h2d_bytes == 16 * (per_step + 16 * 128 * 4)
    assert c1.d2h_bytes == (16 + 4) * per_step + 16 * (16 * 128 * 4 + 16 * 5)
