"""GPU parity: motion model, VPoser, and the vectorised CrowdEnv step / reset vs the CPU oracle."""
import numpy as np
import pytest
import torch

from egogen_b200 import assets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def world(dev, smplx_model):
    """GPU operators + the matching CPU oracle objects built from the same arrays / weights."""
    from egogen_b200.crowd_env import BoxSceneSampler, CrowdVectorEnv, default_cfg
    from egogen_b200.models_gamma_primitive import GAMMAPrimitiveComboGenOP, load_vposer
    from egogen_b200.smplx_parser import get_lbs_model
    from oracle import nets
    from oracle.env import CrowdEnvOracle
    from oracle.smplx_lbs import SMPLXParserOracle
    markers = assets.marker_ids()
    lbs = get_lbs_model("male", dev, arrays=smplx_model, marker_vids=markers)
    genop = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": 0})
    genop.build_model(seed=0)
    vposer, _ = load_vposer(seed=0, device=dev)
    scene = assets.make_box_scene(2, n_boxes=2)
    sdf_cpu = assets.rasterize_scene_sdf(scene, D=64)
    sdf = {k: v.to(dev) for k, v in sdf_cpu.items()}
    rings = assets.scene_polygon(scene)
    sampler = BoxSceneSampler(sdf, lbs, dev, seed=0)
    E = 6
    venv = CrowdVectorEnv(default_cfg(), genop.model, lbs, vposer, sdf, rings, sampler, E, dev, debug_terms=True,
                          capture_rollout=True)
    # oracle twins with identical weights
    combo = nets.ComboOracle()
    combo.predictor.load_state_dict(genop.model.predictor.state_dict())
    combo.regressor.load_state_dict(genop.model.regressor.state_dict())
    vp_o = nets.VPoserEncoderOracle()
    vp_o.load_state_dict(vposer.state_dict())
    orc = CrowdEnvOracle(SMPLXParserOracle(smplx_model, marker=markers), combo.eval(), vp_o.eval(), sdf_cpu,
                         assets.rings_to_segments(rings), markers, assets.feet_marker_idx(), assets.feet_vids())
    return dict(venv=venv, orc=orc, genop=genop, vposer=vposer, combo=combo, vp_o=vp_o, sampler=sampler, E=E, lbs=lbs)


def _assert_sample_prior_close(world, X, betas, z, Y, Yb):
    """GPU sample_prior against the oracle nets evaluated in FLOAT64 (the fp32 oracle is itself ~1e-4 off in the axis-angle
    columns of ill-conditioned joints, so it cannot referee between two fp32-class implementations).
    * markers Y and the body-parameter columns the regressor emits directly (transl, hand PCA): tight absolute bounds;
    * axis-angle columns: 1e-4 wherever the 6-D -> rotation conversion is well conditioned. Gram-Schmidt divides by the norm
      of the (orthogonalised) 6-D columns, and the synthetic regressor emits some joints with norms down to 2e-3, where an
      input error of 5e-7 becomes 2.5e-4; those joints - counted and required to be few - get 1e-5 x the amplification."""
    import copy
    c64 = copy.deepcopy(world["combo"]).double().eval()
    b = X.shape[1]
    b18 = betas.double().unsqueeze(0).repeat(18, 1, 1)
    with torch.no_grad():
        Yo = c64.predictor.sample_prior(X.double(), z.double())
        xb = c64.regressor.forward_cont(Yo.reshape(18 * b, 201), b18.reshape(18 * b, 10))
        Ybo = c64.regressor.cont2aa(xb).view(18, b, 93)
    Y, Yb = Y.cpu().double(), Yb.cpu().double()
    assert Y.shape == (18, b, 201) and Yb.shape == (18, b, 93)
    assert torch.allclose(Y, Yo, atol=2e-5, rtol=1e-4), (Y - Yo).abs().max()
    assert torch.allclose(Yb[..., :3], Ybo[..., :3], atol=5e-6, rtol=1e-5), (Yb[..., :3] - Ybo[..., :3]).abs().max()
    assert torch.allclose(Yb[..., 69:], Ybo[..., 69:], atol=5e-6, rtol=1e-5), (Yb[..., 69:] - Ybo[..., 69:]).abs().max()
    x6 = xb[:, 3:135].reshape(-1, 22, 3, 2)
    a1, a2 = x6[..., 0], x6[..., 1]
    b1 = a1 / a1.norm(dim=-1, keepdim=True)
    a2o = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    amp = (1.0 / torch.minimum(a1.norm(dim=-1), a2o.norm(dim=-1))).clamp_min(1.0).view(18, b, 22, 1)
    tol = torch.maximum(torch.full_like(amp, 1e-4), 1e-5 * amp)
    err = (Yb[..., 3:69] - Ybo[..., 3:69]).abs().view(18, b, 22, 3)
    assert (amp > 10.0).double().mean().item() < 0.1, "the test model must keep most joints well conditioned"
    bad = err > tol + 1e-4 * Ybo[..., 3:69].abs().view(18, b, 22, 3)
    assert not bad.any(), (err.max(), int(bad.sum()))


def test_sample_prior_matches_oracle(dev, world):
    g = torch.Generator().manual_seed(1)
    b = 5
    X = torch.randn(2, b, 201, generator=g) * 0.3
    z = torch.randn(b, 128, generator=g)
    betas = torch.randn(b, 10, generator=g) * 0.5
    Y, Yb = world["genop"].model.sample_prior(X.to(dev), betas.unsqueeze(0).repeat(18, 1, 1).to(dev), z.to(dev))
    _assert_sample_prior_close(world, X, betas, z, Y, Yb)


def test_fused_motion_kernels_match_layerwise(dev, world):
    """The fused decode / regressor kernels against the layer-by-layer path (same weights, same inputs), at a
    batch that spans several row tiles and is not a multiple of the tile sizes."""
    g = torch.Generator().manual_seed(3)
    b = 37
    X = (torch.randn(2, b, 201, generator=g) * 0.3).to(dev)
    z = torch.randn(b, 128, generator=g).to(dev)
    betas = (torch.randn(b, 10, generator=g) * 0.5).to(dev)
    m = world["genop"].model
    m.set_fused(True)
    Y1, Yb1 = m.sample_prior(X, betas, z)
    m.set_fused(False)
    Y0, Yb0 = m.sample_prior(X, betas, z)
    m.set_fused(True)
    assert torch.allclose(Y1, Y0, atol=2e-5, rtol=1e-4)
    assert torch.allclose(Yb1, Yb0, atol=2e-4, rtol=1e-3)


def test_vposer_matches_oracle(dev, world):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(40, 63, generator=g) * 0.4
    loc = world["vposer"].encode(x.to(dev)).loc
    with torch.no_grad():
        ref = world["vp_o"].encode_loc(x)
    assert torch.allclose(loc.cpu(), ref, atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("b", [96, 70, 256])
def test_sample_prior_large_batch_matches_oracle(dev, world, b):
    """The tcgen05 decode (one 16-CTA cluster per 128 rows) and regressor (128 marker frames per CTA): 70 and 96 exercise
    ragged last tiles, 256 is the bench shape."""
    g = torch.Generator().manual_seed(100 + b)
    X = torch.randn(2, b, 201, generator=g) * 0.3
    z = torch.randn(b, 128, generator=g)
    betas = torch.randn(b, 10, generator=g) * 0.5
    Y, Yb = world["genop"].model.sample_prior(X.to(dev), betas.unsqueeze(0).repeat(18, 1, 1).to(dev), z.to(dev))
    _assert_sample_prior_close(world, X, betas, z, Y, Yb)
    # and twice in a row (bit-identical: fixed MMA order, no atomics)
    Y2, _ = world["genop"].model.sample_prior(X.to(dev), betas.unsqueeze(0).repeat(18, 1, 1).to(dev), z.to(dev))
    assert torch.equal(Y, Y2)


def _sync_oracle(orc, venv):
    b = venv.buf
    orc.set_state(state=b["state"].cpu(), seed=b["seed"].cpu(), R0=b["R0"].cpu(), T0=b["T0"].cpu().view(-1, 1, 3),
                  betas=b["betas"].cpu(), dist=b["dist"].cpu(), steps=b["steps"].cpu().to(torch.int64),
                  goal=b["goal"].cpu())


def test_reset_matches_oracle(dev, world):
    venv, orc, E = world["venv"], world["orc"], world["E"]
    s = world["sampler"].next_body(E)
    accept = venv.reset_from(torch.arange(E), s["world_params"], s["goals"], s["betas"])
    ref = orc.reset_from(s["world_params"].cpu(), s["goals"].cpu(), s["betas"].cpu())
    assert torch.equal(accept.cpu().bool(), ref["accept"])
    assert accept.sum() > 0
    m = ref["accept"]
    b = venv.buf
    assert torch.allclose(b["R0"].cpu()[m], ref["R0"][m], atol=1e-5)
    assert torch.allclose(b["T0"].cpu()[m], ref["T0"][m, 0], atol=1e-5)
    assert torch.allclose(b["seed"].cpu()[m], ref["seed"][m], atol=2e-5)
    assert torch.allclose(b["state"].cpu()[m], ref["state"][m], atol=2e-5)
    assert torch.allclose(b["dist"].cpu()[m], ref["dist"][m], atol=1e-5)
    assert torch.allclose(b["ego"].cpu()[m], ref["egosensing"][m], atol=1e-3), (b["ego"].cpu()[m] - ref["egosensing"][m]).abs().max()


def test_step_matches_oracle(dev, world):
    """4 consecutive vector steps; the oracle is re-seeded from the GPU state before each step so every
    step is an operator-level comparison on identical inputs (chaotic end-to-end)."""
    venv, orc, E = world["venv"], world["orc"], world["E"]
    venv.sampler.seed(3)
    venv.reset()
    g = torch.Generator().manual_seed(4)
    for it in range(4):
        _sync_oracle(orc, venv)
        z = torch.randn(E, 128, generator=g)
        obs, rew, term, _, _ = venv.step(z.to(dev))
        ref = orc.step(z)
        b = venv.buf
        assert torch.allclose(b["out_params"].cpu(), ref["params"], atol=3e-4, rtol=1e-3)
        assert torch.allclose(b["out_markers"].cpu(), ref["marker_b"], atol=1e-4)
        assert torch.allclose(b["reward_terms"].cpu(), ref["terms"], atol=2e-4), (b["reward_terms"].cpu() - ref["terms"]).abs().max(0)
        assert torch.allclose(rew.cpu(), ref["reward"], atol=5e-4)
        assert torch.equal(term.cpu().bool(), ref["terminated"])
        assert torch.allclose(b["R0"].cpu(), ref["R0"], atol=1e-4)
        assert torch.allclose(b["T0"].cpu(), ref["T0"][:, 0], atol=1e-4)
        assert torch.allclose(b["seed"].cpu(), ref["seed"], atol=3e-4, rtol=1e-3)
        assert torch.allclose(obs["state"].cpu(), ref["state"], atol=2e-4)
        d_ego = (obs["egosensing"].cpu() - ref["egosensing"]).abs()
        assert torch.allclose(obs["egosensing"].cpu(), ref["egosensing"], atol=2e-3), (it, float(d_ego.max()), int((d_ego > 2e-3).sum()), d_ego.flatten().topk(4).values)
        assert torch.allclose(obs["dist"].cpu()[:, 0], ref["dist"], atol=1e-4)
        assert torch.allclose(obs["time"].cpu()[:, 0], ref["time"], atol=1e-6)
        assert int(b["steps"][0]) == it + 1


def test_single_env_gym_surface(dev, world, smplx_model):
    """CrowdEnv keeps the reference's 12-element init_env constructor and 5-tuple step."""
    from types import SimpleNamespace
    from egogen_b200 import SMPLXParser
    from egogen_b200.crowd_env import CrowdEnv, default_cfg
    parser = SMPLXParser({"n_batch": 20, "device": dev, "marker_placement": "ssm2_67",
                          "smplx_models": {"male": smplx_model, "female": smplx_model}})
    sampler = world["sampler"]
    scene = assets.make_box_scene(2, n_boxes=2)
    sampler.scene_rings = assets.scene_polygon(scene)
    genop = world["genop"]
    init_env = [default_cfg(), genop, genop, None, sampler, parser, parser, parser, assets.feet_marker_idx(),
                assets.marker_ids(), world["vposer"], world["venv"].scene_sdf]
    env = CrowdEnv(init_env, save_rollout=True)
    env.seed(0)
    obs, info = env.reset()
    assert obs["state"].shape == (2, 402) and obs["egosensing"].shape == (2, 32)
    assert obs["dist"].shape == (1,) and obs["time"].shape == (1,) and info == {}
    n = 0
    term = False
    while not term:
        obs, rew, term, trunc, info = env.step(np.zeros(128, dtype=np.float32))
        assert isinstance(rew, float) and isinstance(term, bool) and trunc is False
        n += 1
    assert 1 <= n <= 13 and len(env.outmps) == n


def test_box_scene_env_matches_oracle(dev, world, smplx_model):
    """f-1: the random_box_obstacle_new env (crowd_env_2f_box.py): 2-D walkability-map penetration, always-terminating."""
    from egogen_b200.crowd_env import BoxSceneSampler, CrowdVectorEnv, default_cfg_box
    from oracle.env import CrowdEnvOracle, get_map
    from oracle.smplx_lbs import SMPLXParserOracle
    scene = assets.make_box_scene(7, n_boxes=3)
    sdf_cpu = assets.rasterize_scene_sdf(scene, D=64)
    sdf = {k: v.to(dev) for k, v in sdf_cpu.items()}
    rings, tris = assets.scene_polygon(scene), assets.scene_navmesh_triangles(scene)
    E = 8
    sampler = BoxSceneSampler(sdf, world["lbs"], dev, seed=5)
    venv = CrowdVectorEnv(default_cfg_box(), world["genop"].model, world["lbs"], world["vposer"], sdf, rings, sampler, E, dev,
                          debug_terms=True, capture_rollout=True, box_mode=True, navmesh_tris=tris)
    markers = assets.marker_ids()
    orc = CrowdEnvOracle(SMPLXParserOracle(smplx_model, marker=markers), world["combo"].eval(), world["vp_o"].eval(), sdf_cpu,
                         assets.rings_to_segments(rings), markers, assets.feet_marker_idx(), assets.feet_vids(), max_depth=11,
                         box_mode=True, navmesh_tris=tris, weight_look=0.1)
    # reset parity incl. a candidate standing inside a box (must be rejected by both)
    s = sampler.next_body(E)
    bx = scene["boxes"][0]
    s["world_params"][0, :, 0] = float(bx[0] + bx[3]) / 2; s["world_params"][0, :, 1] = float(bx[1] + bx[4]) / 2
    acc = venv.reset_from(torch.arange(E), s["world_params"], s["goals"], s["betas"])
    ref = orc.reset_from(s["world_params"].cpu(), s["goals"].cpu(), s["betas"].cpu())
    assert torch.equal(acc.cpu().bool(), ref["accept"]) and not bool(acc[0]) and acc.sum() > 0
    venv.reset()
    g = torch.Generator().manual_seed(8)
    for it in range(3):
        _sync_oracle(orc, venv)
        z = torch.randn(E, 128, generator=g)
        obs, rew, term, _, _ = venv.step(z.to(dev))
        r = orc.step(z)
        b = venv.buf
        assert torch.allclose(b["reward_terms"].cpu(), r["terms"], atol=2e-4), (b["reward_terms"].cpu() - r["terms"]).abs().max(0)
        assert torch.allclose(rew.cpu(), r["reward"], atol=5e-4)
        assert torch.equal(term.cpu().bool(), r["terminated"])
        assert torch.allclose(obs["state"].cpu(), r["state"], atol=2e-4)
    # the walkability map itself: every grid cell classification matches get_map for random frames
    R = torch.eye(3).repeat(4, 1, 1); T = torch.rand(4, 1, 3) * 4 - 2
    _, lmap = get_map(tris, R, T)
    assert (lmap == -1).any() and (lmap == 1).any()


def test_crowd_scene_env_matches_oracle(dev, world, smplx_model):
    """f-2: multi-agent crowd dynamics (dummy_vector_env.py:29-128 + crowd_env_crowd_eval.py): the other agents' marker
    boxes are holes of the floor polygon for the walkability map and the ego rays, boxes are refreshed before every
    agent's step (Gauss-Seidel order of the synchronous DummyVectorEnv workers), no penetration termination."""
    from egogen_b200.crowd_env import BoxSceneSampler, CrowdSceneVectorEnv, default_cfg_box
    from oracle.env import CrowdEnvOracle, egosensing
    from oracle.smplx_lbs import SMPLXParserOracle
    S, A = 3, 4
    E = S * A
    scene = assets.make_box_scene(3, n_boxes=0)
    sdf_cpu = assets.rasterize_scene_sdf(scene, D=32)
    sdf = {k: v.to(dev) for k, v in sdf_cpu.items()}
    venv = CrowdSceneVectorEnv(default_cfg_box(), world["genop"].model, world["lbs"], world["vposer"], sdf, S, dev, n_agents=A,
                               sequential=True, debug_terms=True)
    markers = assets.marker_ids()
    floor = [np.asarray(CrowdSceneVectorEnv.FLOOR, np.float64)]
    fl = np.asarray(CrowdSceneVectorEnv.FLOOR, np.float32)
    tris = np.stack([fl[[0, 1, 2]], fl[[2, 3, 0]]])
    orcs = []
    for a in range(A):
        o = CrowdEnvOracle(SMPLXParserOracle(smplx_model, marker=markers), world["combo"].eval(), world["vp_o"].eval(), sdf_cpu,
                           assets.rings_to_segments(floor), markers, assets.feet_marker_idx(), assets.feet_vids(), max_depth=11,
                           box_mode=True, navmesh_tris=tris, weight_look=0.1)
        o.crowd = True
        orcs.append(o)
    # start data: agents of a scene on a 0.7 m circle, agent 1 almost on top of agent 0 (overlapping boxes)
    sampler = BoxSceneSampler(sdf, world["lbs"], dev, seed=11)
    s = sampler.next_body(E)
    wp, goals, betas = s["world_params"].clone(), s["goals"].clone(), s["betas"].clone()
    for a in range(A):
        for sc in range(S):
            e = a * S + sc
            ang = 2 * np.pi * a / A + 0.3 * sc
            pos = np.array([0.7 * np.cos(ang), 0.7 * np.sin(ang)]) if a != 1 else np.array([0.7 + 0.25, 0.05 * sc])
            d = wp[e, 1, :2] - wp[e, 0, :2]
            wp[e, 0, :2] = torch.tensor(pos, dtype=torch.float32)
            wp[e, 1, :2] = wp[e, 0, :2] + d
            goals[e, :2] = torch.tensor(-3.0 * pos / np.linalg.norm(pos), dtype=torch.float32)
    sl = lambda a: slice(a * S, (a + 1) * S)

    def holes_for(bbox_list, a):
        return torch.stack([bbox_list[o] for o in range(A) if o != a], dim=1)       # [S,A-1,4]

    # ---- reset: boxes first, then the observation against the distributed holes -------------------------------
    venv.reset_from(torch.arange(E), wp, goals, betas)
    bb = []
    for a in range(A):
        orcs[a].holes = None
        bb.append(orcs[a].reset_from(wp[sl(a)].cpu(), goals[sl(a)].cpu(), betas[sl(a)].cpu())["bbox"])
    assert torch.allclose(venv.bbox.cpu(), torch.cat(bb), atol=1e-5)
    ego_changed = 0.0
    for a in range(A):
        orcs[a].holes = holes_for(bb, a)
        assert torch.allclose(venv.holes[sl(a)].cpu(), orcs[a].holes, atol=1e-5)
        ref = orcs[a].reset_from(wp[sl(a)].cpu(), goals[sl(a)].cpu(), betas[sl(a)].cpu())
        assert bool(ref["accept"].all())
        assert torch.allclose(venv.buf["state"][sl(a)].cpu(), ref["state"], atol=2e-5)
        assert torch.allclose(venv.buf["ego"][sl(a)].cpu(), ref["egosensing"], atol=1e-3)
        orcs[a].holes = None
        ego_changed += (orcs[a].reset_from(wp[sl(a)].cpu(), goals[sl(a)].cpu(), betas[sl(a)].cpu())["egosensing"]
                        - ref["egosensing"]).abs().sum().item()
    assert ego_changed > 1.0, "the other agents must be visible to the ego rays in this set-up"
    # ---- steps in the reference's update order ------------------------------------------------------------------
    g = torch.Generator().manual_seed(21)
    pene_seen = False
    for it in range(2):
        b = venv.buf
        for a in range(A):
            orcs[a].set_state(state=b["state"][sl(a)].cpu(), seed=b["seed"][sl(a)].cpu(), R0=b["R0"][sl(a)].cpu(),
                              T0=b["T0"][sl(a)].cpu().view(-1, 1, 3), betas=b["betas"][sl(a)].cpu(), dist=b["dist"][sl(a)].cpu(),
                              steps=b["steps"][sl(a)].cpu().to(torch.int64), goal=b["goal"][sl(a)].cpu())
        bb = [venv.bbox[sl(a)].cpu().clone() for a in range(A)]
        z = torch.randn(E, 128, generator=g) * 0.5
        obs, rew, term, _, _ = venv.step(z.to(dev))
        for a in range(A):
            orcs[a].holes = holes_for(bb, a)                   # update_holes_for_each_agent() before this worker's step
            r = orcs[a].step(z[sl(a)])
            bb[a] = orcs[a].bbox                                # agents > a see the new box in the same vector step
            assert torch.allclose(b["reward_terms"][sl(a)].cpu(), r["terms"], atol=2e-4)
            assert torch.allclose(rew[sl(a)].cpu(), r["reward"], atol=5e-4)
            assert torch.equal(term[sl(a)].cpu().bool(), r["terminated"])
            assert torch.allclose(obs["state"][sl(a)].cpu(), r["state"], atol=2e-4)
            assert torch.allclose(obs["egosensing"][sl(a)].cpu(), r["egosensing"], atol=1e-3)
            assert torch.allclose(venv.bbox[sl(a)].cpu(), bb[a], atol=2e-4)
            pene_seen = pene_seen or bool((r["terms"][:, 6] == 0).any())
        assert not bool(term.any()) or it > 0                   # no penetration termination in the crowd env
    assert pene_seen, "agents 0/1 overlap: the map penetration of the dynamic holes must trigger at least once"
    # Jacobi variant: one launch sequence for all agents against the previous step's boxes
    venv.sequential = False
    venv.step(torch.zeros(E, 128, device=dev))
    assert bool(torch.isfinite(venv.buf["reward"]).all())
    venv.close()


def test_sync_free_collect_restarts_finished_envs(dev):
    """The device-side restart path (eg_env_reset_masked + pre-validated start candidates): every env flagged
    `terminated` in step t starts step t+1 from a fresh episode (time observation 1, distance reset), every other env
    continues, and the lazily read statistics count exactly the terminated transitions."""
    from egogen_b200.runtime import build_world
    w = build_world(dev, 32, seed=3, sdf_res=64)
    col, venv = w["collector"], w["venv"]
    w["policy"].train()
    col.reset()
    # validated candidates are accepted by construction
    s = venv._validated_candidates(32)
    acc = venv.reset_from(torch.arange(32), s["world_params"], s["goals"], s["betas"])
    assert bool((acc != 0).all())
    total_term = 0
    for it in range(5):                       # 13-step episodes: 20 vector steps see every env finish at least once
        batch, st = col.collect(32 * 4)
        b = col.buf
        term = b.term.bool()
        total_term += int(term.sum())
        assert st["n/ep"] == int(term.sum())
        assert bool(torch.isfinite(batch.returns).all()) and bool(torch.isfinite(batch.adv).all())
        for t in range(3):
            nxt_time = b.time[t + 1]
            assert torch.allclose(nxt_time[term[t]], torch.ones_like(nxt_time[term[t]]))          # restarted: 1 - 0/max_depth
            assert bool((nxt_time[~term[t]] < b.time[t][~term[t]]).all())                        # running: time decreases
        if term[3].any():
            assert torch.allclose(venv.buf["obs_time"][term[3]], torch.ones_like(venv.buf["obs_time"][term[3]]))
            assert bool((venv.buf["steps"][term[3]] == 0).all())
    assert total_term >= 32
    assert col.collect_episode == total_term


def test_masked_reset_edge_cases(dev, world):
    """reset_masked (eg_env_restart_from_pool: copy of pre-computed initial states) and the C entry point
    eg_env_reset_masked: an all-zero mask changes nothing AND consumes no candidate, a sparse mask consumes exactly as many
    pool rows as it has set entries (k-th flagged env <- k-th unused row), an all-one mask restarts every slot in order, the
    pooled initial state is what a direct reset of the same candidate produces, a null mask is an error."""
    import ctypes as C
    from egogen_b200 import _lib
    venv, E = world["venv"], world["E"]
    venv.reset()
    before = {k: venv.buf[k].clone() for k in ("state", "seed", "R0", "T0", "steps", "goal")}
    venv.reset_masked(torch.ones(E, dtype=torch.uint8, device=dev))       # builds the pool
    venv.reset()
    before = {k: venv.buf[k].clone() for k in ("state", "seed", "R0", "T0", "steps", "goal")}
    c0 = int(venv._cursor[0].item())
    venv.reset_masked(torch.zeros(E, dtype=torch.uint8, device=dev))
    for k, v in before.items():
        assert torch.equal(venv.buf[k], v), k
    assert int(venv._cursor[0].item()) == c0                              # nothing terminated -> no candidate consumed
    sparse = torch.zeros(E, dtype=torch.uint8, device=dev); sparse[1] = 1; sparse[E - 1] = 1
    venv.reset_masked(sparse)
    assert int(venv._cursor[0].item()) == c0 + 2
    assert torch.equal(venv.buf["goal"][1], venv._vpool["goal"][c0]) and torch.equal(venv.buf["goal"][E - 1], venv._vpool["goal"][c0 + 1])
    assert torch.equal(venv.buf["state"][0], before["state"][0])          # unflagged slots untouched
    venv.step(torch.zeros(E, 128, device=dev))
    assert bool((venv.buf["steps"] == 1).all())
    venv.reset_masked(torch.ones(E, dtype=torch.uint8, device=dev))
    assert bool((venv.buf["steps"] == 0).all())
    lo = int(venv._cursor[0].item()) - E
    pooled = {k: venv._vpool[k][lo:lo + E].clone() for k in ("goal", "world_params", "betas", "state", "seed", "R0", "T0", "ego", "dist")}
    assert torch.equal(venv.buf["goal"], pooled["goal"])                # candidate e -> slot e
    after = {k: venv.buf[k].clone() for k in ("state", "seed", "R0", "T0", "ego", "dist")}
    # a direct reset of the same candidates reproduces the pooled state bit for bit
    acc = venv.reset_from(torch.arange(E), pooled["world_params"], pooled["goal"], pooled["betas"])
    assert bool((acc != 0).all())
    for k, v in after.items():
        assert torch.equal(venv.buf[k], v), k
    # C entry point: masked commit through the full reset pipeline gives the same state; null mask rejected
    venv.step(torch.zeros(E, 128, device=dev))
    mask = torch.zeros(E, dtype=torch.uint8, device=dev); mask[::2] = 1
    acc32 = torch.zeros(E, dtype=torch.int32, device=dev)
    wp, gl, be = pooled["world_params"].contiguous(), pooled["goal"].contiguous(), pooled["betas"].contiguous()
    _lib.check(_lib.lib().eg_env_reset_masked(venv._h, C.byref(venv._cbuf), _lib.ptr(mask), E, _lib.ptr(wp), _lib.ptr(gl),
                                              _lib.ptr(be), _lib.ptr(acc32), _lib.stream_ptr(dev)))
    assert torch.equal(acc32.bool(), mask.bool())
    assert torch.equal(venv.buf["state"][::2], after["state"][::2]) and bool((venv.buf["steps"][1::2] == 1).all())
    rc = _lib.lib().eg_env_reset_masked(venv._h, C.byref(venv._cbuf), None, E, _lib.ptr(wp), _lib.ptr(gl), _lib.ptr(be),
                                        _lib.ptr(acc32), _lib.stream_ptr(dev))
    assert rc < 0


def test_product_sampler_matches_reference_golden(dev, world, golden_dir):
    """The PRODUCT start-body sampler (egogen_b200/scene_sampler.py::CrowdMotionSampler.gen_init_bodies, SMPL-X joints from
    the CUDA LBS operator) against the outputs of the reference's own CrowdMotion.gen_init_body
    (exp_GAMMAPrimitive/utils/environments.py:1041-1131, tests/golden/gen_sampler_golden.py), batched over the three golden
    cases, plus the sampler-dict contract (batched_to_dicts -> sampler_dicts_to_candidates -> a reset the env accepts)."""
    import os
    from scipy.spatial.transform import Rotation
    from egogen_b200.scene_sampler import CrowdMotionSampler, batched_to_dicts, sampler_dicts_to_candidates
    g = np.load(os.path.join(golden_dir, "sampler_golden.npz"))
    seed = np.load(os.path.join(golden_dir, "locomotion_seed_00343.npz"))
    smp = CrowdMotionSampler(world["lbs"], dev, {"poses": seed["poses"], "trans": seed["trans"], "betas": seed["betas"][:10]}, seed=0)
    sf = g["start_frame"].astype(int)
    ms = (np.stack([seed["betas"][:10]] * 3), np.stack([seed["poses"][f:f + 2, 3:66] for f in sf]),
          np.stack([seed["poses"][f:f + 2, :3] for f in sf]), np.stack([seed["trans"][f:f + 2] for f in sf]))
    out = smp.gen_init_bodies(g["start"], g["target"], motion_seed=ms, yaw=g["yaw"].astype(np.float32))
    assert np.abs(out["transl"].cpu().numpy() - g["transl"]).max() < 5e-5
    assert np.abs(out["wpath"].cpu().numpy() - g["wpath"]).max() < 5e-5
    assert np.array_equal(out["body_pose"].cpu().numpy(), g["body_pose"]) and np.array_equal(out["betas"].cpu().numpy(), g["betas"])
    R = Rotation.from_rotvec(out["global_orient"].cpu().numpy().reshape(-1, 3).astype(np.float64)).as_matrix()
    R_ref = Rotation.from_rotvec(g["global_orient"].reshape(-1, 3).astype(np.float64)).as_matrix()
    assert np.abs(R - R_ref).max() < 5e-5                      # same rotations (the axis-angle branch may differ at pi)
    # random draws: frames come from the seed recording, bodies stand on the floor above their start point
    rnd = smp.gen_init_bodies(g["start"], g["target"])
    j = smp._joints(rnd["transl"], rnd["global_orient"], rnd["body_pose"], rnd["betas"])
    assert float(j[:, 0, :, 2].amin(dim=1).abs().max()) < 1e-4
    assert float((j[:, 0, 0, :2].cpu() - torch.as_tensor(g["start"][:, :2])).abs().max()) < 1e-4
    # dict contract: reference-format dicts -> reset candidates -> accepted by the env on an empty floor
    dicts = batched_to_dicts(out)
    assert set(dicts[0]) == {"gender", "motion_seed", "betas", "wpath", "scene_path", "navmesh", "navmesh_path", "floor_height"}
    wp, goals, betas = sampler_dicts_to_candidates(dicts, dev)
    assert wp.shape == (3, 2, 93) and torch.allclose(goals.cpu(), torch.as_tensor(g["wpath"][:, 1]), atol=5e-5)
    one = smp.next_body((g["start"][0], g["target"][0]), num_agents=1)
    assert one["motion_seed"]["transl"].shape == (2, 3) and one["wpath"].shape == (2, 3)


def test_env_step_at_bench_configuration(dev, world, smplx_model):
    """BASELINE config 2 shape (crowd_env_2f.py:78-317 at 256 parallel envs, 256^3 SDF, full 10 475-vertex mesh): reset and
    two vector steps of the tensor-core decode / regressor / fused LBS + SDF path against the CPU oracle, re-seeded from
    the GPU state before each step (operator-level comparison on identical inputs)."""
    from egogen_b200.crowd_env import BoxSceneSampler, CrowdVectorEnv, default_cfg
    from oracle.env import CrowdEnvOracle
    from oracle.smplx_lbs import SMPLXParserOracle
    E = 256
    markers = assets.marker_ids()
    scene = assets.make_box_scene(7, n_boxes=2)
    sdf_cpu = assets.rasterize_scene_sdf(scene, D=256)
    sdf = {k: v.to(dev) for k, v in sdf_cpu.items()}
    rings = assets.scene_polygon(scene)
    venv = CrowdVectorEnv(default_cfg(), world["genop"].model, world["lbs"], world["vposer"], sdf, rings,
                          BoxSceneSampler(sdf, world["lbs"], dev, seed=5), E, dev, debug_terms=True, capture_rollout=True)
    orc = CrowdEnvOracle(SMPLXParserOracle(smplx_model, marker=markers), world["combo"].eval(), world["vp_o"].eval(), sdf_cpu,
                         assets.rings_to_segments(rings), markers, assets.feet_marker_idx(), assets.feet_vids())
    venv.reset()
    g = torch.Generator().manual_seed(41)
    n_term = 0
    for it in range(2):
        _sync_oracle(orc, venv)
        z = torch.randn(E, 128, generator=g)
        obs, rew, term, _, _ = venv.step(z.to(dev))
        ref = orc.step(z)
        b = venv.buf
        assert torch.allclose(b["out_markers"].cpu(), ref["marker_b"], atol=1e-4)
        assert torch.allclose(b["reward_terms"].cpu(), ref["terms"], atol=2e-4), (b["reward_terms"].cpu() - ref["terms"]).abs().max(0)
        assert torch.allclose(rew.cpu(), ref["reward"], atol=5e-4)
        assert torch.equal(term.cpu().bool(), ref["terminated"])
        assert torch.allclose(obs["state"].cpu(), ref["state"], atol=2e-4)
        assert torch.allclose(obs["egosensing"].cpu(), ref["egosensing"], atol=2e-3)
        assert torch.allclose(obs["dist"].cpu()[:, 0], ref["dist"], atol=1e-4)
        n_term += int(term.sum())
    # the penetration term must have been exercised: some bodies touch the boxes / floor, most do not
    pen = venv.buf["reward_terms"][:, 6].cpu()
    assert int((pen < pen.max()).sum()) > 0 and int((pen == pen.max()).sum()) > 0
    venv.close()


def test_egosensing_operator_tight(dev):
    """eg_egosensing (the kernel the step / reset paths launch) on GIVEN joints against the oracle's restatement of
    CrowdEnv._calc_egosensing (crowd_env_2f.py:524-613): both take the same float32 joints and run the rays in float64, so
    the ray distances agree to 1e-5 of the [-1,1] output (7e-5 m... in fact ~1e-7) - EXCEPT rays that graze a polygon
    corner, where which segment is hit first flips with the last bit of the float32 world transform. Those are identified
    from the oracle itself (its answer moves by more than 1e-5 when the joints move by 2e-6) and must stay below 1 %."""
    from egogen_b200.crowd_env import calc_egosensing
    from oracle.env import egosensing
    scene = assets.make_box_scene(3, n_boxes=3)
    rings = assets.scene_polygon(scene)
    segs = assets.rings_to_segments(rings)
    n = 200
    g = torch.Generator().manual_seed(12)
    j = torch.randn(n, 2, 127, 3, generator=g) * 0.05
    j[:, :, 23, 0] += 0.03; j[:, :, 24, 0] -= 0.03               # eyes apart, eyeballs 56 / 57 ahead of them
    j[:, :, 56] = j[:, :, 24] + torch.tensor([0.0, 0.0, 0.1]); j[:, :, 57] = j[:, :, 23] + torch.tensor([0.0, 0.0, 0.1])
    ang = torch.rand(n, generator=g) * 6.2831853
    R0 = torch.zeros(n, 3, 3)
    R0[:, 0, 0] = ang.cos(); R0[:, 0, 2] = ang.sin(); R0[:, 1, 0] = ang.sin(); R0[:, 1, 2] = -ang.cos(); R0[:, 2, 1] = 1
    T0 = torch.cat([(torch.rand(n, 2, generator=g) * 2 - 1) * 3.5, torch.full((n, 1), 1.5)], 1)
    holes = torch.cat([T0[:, None, :2].roll(1, 0) - 0.3, T0[:, None, :2].roll(1, 0) + 0.3], dim=2)   # one rectangle per item

    def world(jl):
        return torch.einsum("nij,ntpj->ntpi", R0, jl) + T0[:, None, None]
    for hl in (None, holes):
        got = calc_egosensing(j.to(dev), R0.to(dev), T0.to(dev), rings, 7.0, None if hl is None else hl.to(dev)).cpu()
        ref = egosensing(world(j), segs, 7.0, hl)
        ref_p = egosensing(world(j) + 2e-6, segs, 7.0, hl)
        ref_m = egosensing(world(j) - 2e-6, segs, 7.0, hl)
        grazing = ((ref_p - ref).abs() > 1e-5) | ((ref_m - ref).abs() > 1e-5)
        assert grazing.float().mean().item() < 0.01, grazing.float().mean()
        err = (got - ref).abs()
        assert err[~grazing].max().item() <= 1e-5, err[~grazing].max()
        assert int((ref > -1).sum()) > 0 and int((ref < 1).sum()) > 0        # some eyes off the polygon, some rays hit
