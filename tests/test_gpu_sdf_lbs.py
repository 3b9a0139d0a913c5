"""GPU parity: calc_sdf and SMPL-X LBS (through the C ABI / Python mirrors) vs the CPU oracle."""
import os

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
def parser(dev, smplx_model):
    from egogen_b200 import SMPLXParser
    return SMPLXParser({"n_batch": 8, "device": dev, "marker_placement": "ssm2_67",
                        "smplx_models": {"male": smplx_model, "female": smplx_model}})


def _t(a):
    return torch.as_tensor(np.asarray(a))


def test_calc_sdf_golden_bit_exact_indices(dev, golden_dir):
    """Golden vectors from the reference's own calc_sdf (utils.py:54-84): values within 2e-6,
    `<0` masks identical, base indices identical to the explicit ATen index arithmetic."""
    from egogen_b200 import calc_sdf
    from oracle import sdf as osdf
    g = np.load(os.path.join(golden_dir, "sdf_golden.npz"))
    for k in "ab":
        d = {"center": _t(g[f"center_{k}"]), "scale": _t(g[f"scale_{k}"]), "sdf": _t(g[f"grid_{k}"])}
        pts = _t(g[f"pts_{k}"])
        val, idx = calc_sdf(pts.to(dev), {kk: v.to(dev) for kk, v in d.items()}, return_index=True)
        ref = _t(g[f"out_{k}"])
        _, ridx = osdf.calc_sdf_explicit(pts, d)
        assert torch.equal(idx.cpu(), ridx)
        assert torch.allclose(val.cpu(), ref, atol=2e-6, rtol=0)
        assert torch.equal(val.cpu() < 0, ref < 0)


def test_calc_sdf_edge_cases(dev):
    from egogen_b200 import calc_sdf
    from oracle import sdf as osdf
    scene = assets.rasterize_scene_sdf(assets.make_box_scene(3, n_boxes=2), D=32)
    sd = {k: v.to(dev) for k, v in scene.items()}
    # empty input
    out = calc_sdf(torch.zeros(2, 0, 3, device=dev), sd)
    assert out.shape == (2, 0)
    # far outside (border clamp), exactly on the faces of the cube, huge values
    pts = torch.tensor([[[100., -100., 50.], [4., 4., 7.], [-4., -4., -1.], [0., 0., 1e20], [1e-30, 0., 0.]]])
    val, idx = calc_sdf(pts.to(dev), sd, return_index=True)
    rv, ridx = osdf.calc_sdf_explicit(pts, scene)
    assert torch.equal(idx.cpu(), ridx) and torch.allclose(val.cpu(), rv, atol=1e-6)
    # large random batch: linearity of the sampler in the grid values (size-independent property)
    P = 1 << 20
    g = torch.Generator().manual_seed(5)
    pts = ((torch.rand(1, P, 3, generator=g) * 2 - 1) * 4.5).to(dev)
    v1 = calc_sdf(pts, sd)
    sd2 = dict(sd); sd2["sdf"] = sd["sdf"] * 2.0
    v2 = calc_sdf(pts, sd2)
    assert torch.allclose(v2, 2 * v1, atol=1e-6)


def _rand_inputs(n, seed, scale=0.3):
    g = torch.Generator().manual_seed(seed)
    xb = torch.randn(n, 93, generator=g) * scale
    xb[:, :3] = torch.rand(n, 3, generator=g) * 6 - 3
    return xb, torch.randn(10, generator=g)


@pytest.mark.parametrize("tcgen05", [True, False])
@pytest.mark.parametrize("n", [1, 4, 8, 37, 300])
def test_lbs_matches_oracle(dev, parser, smplx_model, n, tcgen05):
    """LBS vertices / joints within 1e-4 relative (north_star tolerance) of the smplx restatement, for both
    full-mesh mainloops (tcgen05 TF32 tiles with hi/lo-split shape rows; fp32 SIMT tiles)."""
    from oracle.smplx_lbs import SMPLXParserOracle
    parser.bm_male.set_mainloop(tcgen05)
    xb, betas = _rand_inputs(n, 100 + n)
    out = parser.forward_smplx(betas.to(dev), "male", xb.to(dev), to_numpy=False, output_type="raw")
    ref = SMPLXParserOracle(smplx_model, marker=assets.marker_ids()).forward_smplx(betas, "male", xb, "raw")
    for a, b in ((out.vertices.cpu(), ref.vertices), (out.joints.cpu(), ref.joints)):
        assert a.shape == b.shape
        rel = (a - b).norm(dim=-1).max() / b.norm(dim=-1).max()
        assert rel < 1e-4, rel                      # tolerance stated by BASELINE.json north_star
        assert (a - b).abs().max() < (5e-5 if tcgen05 else 2e-5), (a - b).abs().max()
    mk = parser.get_markers(betas.to(dev), "male", xb.to(dev), to_numpy=False)
    if tcgen05:   # markers come from the fp32 compact-set kernel, full vertices from the TF32 tiles
        assert torch.allclose(mk, out.vertices[:, parser.marker], atol=5e-5)
    else:         # same kernel, same arithmetic: compact set == gathered full set bit for bit
        assert torch.equal(mk, out.vertices[:, parser.marker])
    parser.bm_male.set_mainloop(True)
    j22 = parser.get_jts(betas.to(dev), "male", xb.to(dev), to_numpy=False)
    assert torch.equal(j22, out.joints[:, :22])


def test_lbs_zero_pose_and_betas(dev, parser, smplx_model):
    """The reference's all-zero case (samplers return betas=0): identity rotations, template + hand mean."""
    from oracle.smplx_lbs import SMPLXParserOracle
    xb = torch.zeros(4, 93)
    betas = torch.zeros(10)
    out = parser.forward_smplx(betas.to(dev), "male", xb.to(dev), to_numpy=False, output_type="raw")
    ref = SMPLXParserOracle(smplx_model).forward_smplx(betas, "male", xb, "raw")
    assert (out.vertices.cpu() - ref.vertices).abs().max() < 1e-5
    assert (out.joints.cpu() - ref.joints).abs().max() < 1e-5


def test_lbs_transl_equivariance_large_batch(dev, parser):
    """Size-independent property at the bench size (5120 bodies): translating xb translates the output."""
    n = 5120
    xb, betas = _rand_inputs(n, 7)
    xb = xb.to(dev)
    j0 = parser.get_all_jts(betas.to(dev), "male", xb, to_numpy=False)
    xb2 = xb.clone(); xb2[:, :3] += torch.tensor([1.0, -2.0, 0.5], device=dev)
    j1 = parser.get_all_jts(betas.to(dev), "male", xb2, to_numpy=False)
    assert torch.allclose(j1 - j0, torch.tensor([1.0, -2.0, 0.5], device=dev).expand_as(j0), atol=2e-6)


def test_fused_lbs_sdf_counts(dev, parser, smplx_model):
    """Fused LBS->world->SDF->count equals the unfused operator chain on the same GPU vertices
    bit-for-bit, and the oracle chain up to vertices within 1e-5 of the SDF zero level."""
    from egogen_b200 import calc_sdf, penetration_count
    from oracle import sdf as osdf
    from oracle.smplx_lbs import SMPLXParserOracle
    E, T = 3, 20
    xb, betas = _rand_inputs(E * T, 11)
    xb[:, :3] *= 0.2
    scene = assets.rasterize_scene_sdf(assets.make_box_scene(1, n_boxes=2), D=64)
    sd = {k: v.to(dev) for k, v in scene.items()}
    g = torch.Generator().manual_seed(3)
    ang = torch.rand(E, generator=g) * 6.28
    R0 = torch.zeros(E, 3, 3); R0[:, 0, 0] = ang.cos(); R0[:, 0, 1] = -ang.sin()
    R0[:, 1, 0] = ang.sin(); R0[:, 1, 1] = ang.cos(); R0[:, 2, 2] = 1
    T0 = torch.rand(E, 1, 3, generator=g) * 3 - 1.5
    T0[:, :, 2] = 0.9
    skip = torch.zeros(assets.V_SMPLX, dtype=torch.uint8)
    skip[assets.feet_vids()] = 1
    bm = parser.bm_male
    counts, joints, markers = bm.forward_sdf(xb.to(dev), betas.to(dev), T, R0.to(dev), T0.to(dev), sd, skip.to(dev))
    # unfused chain on the GPU
    out = parser.forward_smplx(betas.to(dev), "male", xb.to(dev), to_numpy=False, output_type="raw")
    vw = torch.einsum("bij,btpj->btpi", R0.to(dev), out.vertices.view(E, T, -1, 3)) + T0.to(dev)[:, None]
    sv = calc_sdf(vw.reshape(E * T, -1, 3), sd)
    c2 = penetration_count(sv, skip.to(dev))
    near = (sv.abs() < 1e-5).sum(dim=1).cpu()
    assert ((counts.cpu() - c2.cpu()).abs() <= near).all()
    assert torch.equal(joints, out.joints) and torch.allclose(markers, out.vertices[:, parser.marker], atol=5e-5)
    assert counts.sum() > 0, "test scene should produce some penetrations"
    # oracle chain
    ref = SMPLXParserOracle(smplx_model).forward_smplx(betas, "male", xb, "raw")
    rvw = torch.einsum("bij,btpj->btpi", R0, ref.vertices.view(E, T, -1, 3)) + T0[:, None]
    rs = osdf.calc_sdf(rvw.reshape(E * T, -1, 3), scene).reshape(E, T, -1)
    rc, _, _ = osdf.penetration_counts(rs, assets.feet_vids(), T)
    near_o = (rs.abs() < 1e-4).sum(dim=-1)
    assert ((counts.cpu().view(E, T) - rc).abs() <= near_o).all()


def test_fused_counts_at_bench_size(dev, parser):
    """BASELINE config-2 size (256 envs x 20 frames = 5120 bodies, 256^3 SDF, bodies standing on the floor next to a
    box): the fused tcgen05 kernel's per-body penetration counts (2-level conservative sign bits + exact sample) equal
    the unfused operator chain (materialised vertices -> calc_sdf -> count) on the same GPU vertices, body for body,
    up to vertices within 1e-5 of the zero level."""
    from egogen_b200 import calc_sdf, penetration_count
    E, T = 256, 20
    g = torch.Generator().manual_seed(17)
    xb = torch.randn(E * T, 93, generator=g) * 0.15
    xb[:, 3] += 3.14159265 / 2                                  # y-up template -> z-up
    xb[:, :3] = 0.0
    betas = torch.randn(E, 10, generator=g) * 0.5
    scene = assets.rasterize_scene_sdf(assets.make_box_scene(5, n_boxes=2), D=256, device=str(dev))
    sd = {k: v.to(dev) for k, v in scene.items()}
    ang = torch.rand(E, generator=g) * 6.28
    R0 = torch.zeros(E, 3, 3); R0[:, 0, 0] = ang.cos(); R0[:, 0, 1] = -ang.sin()
    R0[:, 1, 0] = ang.sin(); R0[:, 1, 1] = ang.cos(); R0[:, 2, 2] = 1
    T0 = torch.cat([torch.rand(E, 1, 2, generator=g) * 5 - 2.5, torch.full((E, 1, 1), 1.05)], dim=2)
    skip = torch.zeros(assets.V_SMPLX, dtype=torch.uint8)
    skip[assets.feet_vids()] = 1
    bm = parser.bm_male
    brow = betas.repeat_interleave(T, 0)
    counts, _, _ = bm.forward_sdf(xb.to(dev), brow.to(dev), T, R0.to(dev), T0.to(dev), sd, skip.to(dev))
    verts = bm.forward(xb.to(dev), brow.to(dev), want_verts=True)[0]
    vw = torch.einsum("bij,btpj->btpi", R0.to(dev), verts.view(E, T, -1, 3)) + T0.to(dev)[:, None]
    sv = calc_sdf(vw.reshape(E * T, -1, 3), sd)
    c2 = penetration_count(sv, skip.to(dev))
    near = (sv.abs() < 1e-5).sum(dim=1)
    assert bool(((counts - c2).abs() <= near).all()), (counts - c2).abs().max()
    assert int((counts > 0).sum()) > E * T // 20, "bodies intersecting the floor / boxes must be present"
    assert int((counts == 0).sum()) > 0
    # ... and an INDEPENDENT reference for the vertices: the fp32 SIMT mainloop (no tensor cores, no fp16 operands). The
    # tcgen05 vertices deviate from it by <= 5e-5 m (tolerance 1e-4 relative, test_forward_matches_oracle) and the SDF is
    # 1-Lipschitz in metres after un-scaling, so only vertices within 1e-4 of the zero level may flip.
    bm.set_mainloop(False)
    try:
        verts32 = bm.forward(xb.to(dev), brow.to(dev), want_verts=True)[0]
    finally:
        bm.set_mainloop(True)
    assert (verts32 - verts).abs().max().item() <= 1e-4
    vw32 = torch.einsum("bij,btpj->btpi", R0.to(dev), verts32.view(E, T, -1, 3)) + T0.to(dev)[:, None]
    sv32 = calc_sdf(vw32.reshape(E * T, -1, 3), sd)
    c32 = penetration_count(sv32, skip.to(dev))
    near32 = (sv32.abs() < 1e-4 * float(sd["scale"].reshape(-1)[0].abs().clamp_min(1.0))).sum(dim=1)
    assert bool(((counts - c32).abs() <= near32).all()), ((counts - c32).abs() - near32).max()
    # ... and the CPU oracle (dense LBS restatement + the reference-pinned calc_sdf) on the first 3 envs (60 bodies)
    from oracle import sdf as osdf
    from oracle.smplx_lbs import SMPLXParserOracle
    nb = 3 * T
    orc = SMPLXParserOracle(assets.make_surrogate_smplx(seed=0), marker=assets.marker_ids())
    vo = orc.forward_smplx(brow[:nb], "male", xb[:nb], "raw").vertices
    vwo = torch.einsum("bij,btpj->btpi", R0[:3], vo.view(3, T, -1, 3)) + T0[:3, None]
    svo = osdf.calc_sdf(vwo.reshape(nb, -1, 3), {k: v.cpu() for k, v in sd.items() if torch.is_tensor(v)})
    keep = (skip == 0)
    co = ((svo < 0) & keep[None]).sum(dim=1)
    nearo = (svo.abs() < 1e-4 * float(sd["scale"].reshape(-1)[0].abs().clamp_min(1.0))).sum(dim=1)
    assert bool(((counts[:nb].cpu() - co).abs() <= nearo).all()), ((counts[:nb].cpu() - co).abs() - nearo).max()


def test_markers_backward_matches_autograd(dev, parser, smplx_model):
    """f-4 building block: eg_lbs_markers_backward (pose blend shapes, kinematic chain, smplx Rodrigues, hand PCA,
    translation) against torch autograd through the LBS oracle; including the
    reference's all-zero pose (the Rodrigues 1e-8 quirk keeps the gradient finite there) and ragged betas rows."""
    from oracle.smplx_lbs import SMPLXParserOracle
    orc = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    g = torch.Generator().manual_seed(5)
    for N, scale in [(6, 0.3), (4, 0.0), (3, 1.2)]:
        xb = torch.randn(N, 93, generator=g) * scale
        betas = torch.randn(N, 10, generator=g) * 0.7
        w = torch.randn(N, 67, 3, generator=g)
        xr = xb.clone().requires_grad_(True)
        (orc.forward_smplx(betas, "male", xr, "markers") * w).sum().backward()
        ref = xr.grad
        got = parser.bm_male.markers_backward(xb.to(dev), betas.to(dev), w.to(dev)).cpu()
        assert torch.isfinite(got).all()
        err = (got - ref).abs().max().item()
        assert err <= 2e-4 * max(ref.abs().max().item(), 1.0), (N, scale, err, ref.abs().max().item())
    # one shared betas row
    xb = torch.randn(5, 93, generator=g) * 0.3
    betas = torch.randn(1, 10, generator=g)
    w = torch.randn(5, 67, 3, generator=g)
    xr = xb.clone().requires_grad_(True)
    (orc.forward_smplx(betas.repeat(5, 1), "male", xr, "markers") * w).sum().backward()
    got = parser.bm_male.markers_backward(xb.to(dev), betas.to(dev), w.to(dev)).cpu()
    assert (got - xr.grad).abs().max().item() <= 2e-4 * max(xr.grad.abs().max().item(), 1.0)


def test_lbs_stress_mesh_many_weights(dev):
    """A mesh that does NOT flatter the tensor-core kernel (assets.make_stress_smplx): 4..8 non-zero skinning weights per
    vertex, so every vertex with more than four takes the epilogue's uncached extra-weight path, and joint sets that differ
    from vertex to vertex, so the layout builder has to close tiles at the 10-joint limit. Vertices / joints against the
    oracle in both mainloops, fused penetration counts against the unfused chain, and the kernel's time at the bench size."""
    from egogen_b200 import SMPLXParser, calc_sdf, penetration_count
    from oracle.smplx_lbs import SMPLXParserOracle
    model = assets.make_stress_smplx(0)
    nnz = (model["lbs_weights"] > 0).sum(axis=1)
    assert nnz.max() == 8 and (nnz > 4).mean() > 0.5
    sp = SMPLXParser({"n_batch": 8, "device": dev, "marker_placement": "ssm2_67", "smplx_models": {"male": model, "female": model}})
    orc = SMPLXParserOracle(model, marker=assets.marker_ids())
    for n in (4, 37):
        xb, betas = _rand_inputs(n, 900 + n)
        ref = orc.forward_smplx(betas, "male", xb, "raw")
        for tcgen05 in (True, False):
            sp.bm_male.set_mainloop(tcgen05)
            out = sp.forward_smplx(betas.to(dev), "male", xb.to(dev), to_numpy=False, output_type="raw")
            for a, b in ((out.vertices.cpu(), ref.vertices), (out.joints.cpu(), ref.joints)):
                rel = (a - b).norm(dim=-1).max() / b.norm(dim=-1).max()
                assert rel < 1e-4, (n, tcgen05, rel)
                assert (a - b).abs().max() < (5e-5 if tcgen05 else 2e-5), (n, tcgen05, (a - b).abs().max())
    sp.bm_male.set_mainloop(True)
    # fused counts == unfused chain on the same vertices (64 envs x 20 frames)
    E, T = 64, 20
    g = torch.Generator().manual_seed(23)
    xb = torch.randn(E * T, 93, generator=g) * 0.15
    xb[:, 3] += 3.14159265 / 2
    xb[:, :3] = 0.0
    betas = torch.randn(E, 10, generator=g) * 0.5
    sd = {k: v.to(dev) for k, v in assets.rasterize_scene_sdf(assets.make_box_scene(5, n_boxes=2), D=128, device=str(dev)).items()}
    ang = torch.rand(E, generator=g) * 6.28
    R0 = torch.zeros(E, 3, 3); R0[:, 0, 0] = ang.cos(); R0[:, 0, 1] = -ang.sin(); R0[:, 1, 0] = ang.sin(); R0[:, 1, 1] = ang.cos(); R0[:, 2, 2] = 1
    T0 = torch.cat([torch.rand(E, 1, 2, generator=g) * 5 - 2.5, torch.full((E, 1, 1), 1.05)], dim=2)
    skip = torch.zeros(assets.V_SMPLX, dtype=torch.uint8)
    skip[assets.feet_vids()] = 1
    bm = sp.bm_male
    brow = betas.repeat_interleave(T, 0)
    counts, _, _ = bm.forward_sdf(xb.to(dev), brow.to(dev), T, R0.to(dev), T0.to(dev), sd, skip.to(dev))
    verts = bm.forward(xb.to(dev), brow.to(dev), want_verts=True)[0]
    vw = torch.einsum("bij,btpj->btpi", R0.to(dev), verts.view(E, T, -1, 3)) + T0.to(dev)[:, None]
    sv = calc_sdf(vw.reshape(E * T, -1, 3), sd)
    c2 = penetration_count(sv, skip.to(dev))
    near = (sv.abs() < 1e-5).sum(dim=1)
    assert bool(((counts - c2).abs() <= near).all()), (counts - c2).abs().max()
    assert int((counts > 0).sum()) > 0
    # time of the fused kernel at the bench size on this mesh (reported, not asserted: profiles/README.md quotes it)
    xb5 = xb.repeat(4, 1); b5 = brow.repeat(4, 1)
    R5, T5 = R0.repeat(4, 1, 1).to(dev), T0.repeat(4, 1, 1).to(dev)
    for _ in range(2):
        bm.forward_sdf(xb5.to(dev), b5.to(dev), T, R5, T5, sd, skip.to(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        bm.forward_sdf(xb5.to(dev), b5.to(dev), T, R5, T5, sd, skip.to(dev))
    e1.record(); torch.cuda.synchronize()
    print(f"stress mesh: fused LBS + SDF call for {xb5.shape[0]} bodies = {e0.elapsed_time(e1) / 5:.3f} ms (pose prep + tc + compact + finish)")
