"""GPU parity of the canonical-frame helpers of SMPLXParser (baseops.py:465-598) and of the primitive canonicalisation /
noise-augmented batches built on them (utils_canonicalize_samp.py:123-187, batch_gen_amass.py:229-263) vs the CPU oracle."""
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
def parsers(dev, smplx_model):
    from egogen_b200 import SMPLXParser
    mk = lambda p: SMPLXParser({"n_batch": 8, "device": dev, "marker_placement": p,
                                "smplx_models": {"male": smplx_model, "female": smplx_model}})
    return {"ssm2_67": mk("ssm2_67"), "cmu_41": mk("cmu_41")}


@pytest.fixture(scope="module")
def orc(smplx_model):
    from oracle.smplx_lbs import SMPLXParserOracle
    return SMPLXParserOracle(smplx_model, marker=assets.marker_ids())


def _motion(seed, T):
    g = torch.Generator().manual_seed(seed)
    transl = torch.randn(1, 3, generator=g) * 0.5 + torch.cumsum(torch.randn(T, 3, generator=g) * 0.01, 0)
    pose = torch.randn(1, 165, generator=g) * 0.25 + torch.cumsum(torch.randn(T, 165, generator=g) * 0.005, 0)
    pose[:, :3] += torch.tensor([1.2, 0.3, -0.4])
    return transl.numpy(), pose.numpy(), (torch.randn(16, generator=g) * 0.5).numpy()


def test_frame_helpers_match_oracle(dev, parsers, orc):
    p = parsers["ssm2_67"]
    g = torch.Generator().manual_seed(2)
    N = 9
    xb = torch.randn(N, 93, generator=g) * 0.4
    betas = torch.randn(1, 10, generator=g) * 0.6
    R_ref, T_ref = orc.get_new_coordinate(betas, "male", xb)
    R, T = p.get_new_coordinate(betas.numpy(), "male", xb.numpy())
    assert R.shape == (N, 3, 3) and T.shape == (N, 1, 3)
    assert np.abs(R - R_ref.numpy()).max() < 2e-5 and np.abs(T - T_ref.numpy()).max() < 2e-5
    d_ref = orc.calc_calibrate_offset("male", betas, xb[:, 6:69])
    d = p.calc_calibrate_offset(p.bm_male, betas.numpy(), xb[:, 6:69].numpy())
    assert d.shape == (N, 3) and np.abs(d - d_ref.numpy()).max() < 2e-5
    ref = orc.update_transl_glorot(R_ref, T_ref, betas, "male", xb)
    xin = xb.numpy().copy()
    out = p.update_transl_glorot(R, T, betas.numpy(), "male", xin, to_numpy=True, inplace=False)
    assert np.abs(out - ref.numpy()).max() < 5e-5 and np.array_equal(xin, xb.numpy())
    out2 = p.update_transl_glorot(R, T, betas.numpy(), "male", xin, to_numpy=True, inplace=True)
    assert out2 is xin and np.array_equal(xin, out)
    xd = xb.to(dev)
    out3 = p.update_transl_glorot(torch.as_tensor(R).to(dev), torch.as_tensor(T).to(dev), betas.to(dev), "male", xd,
                                  to_numpy=False, inplace=True)
    assert out3.data_ptr() == xd.data_ptr() and np.array_equal(xd.cpu().numpy(), out)
    # a body expressed in its own canonical frame has the identity frame
    R2, T2 = p.get_new_coordinate(betas.numpy(), "male", out)
    assert np.abs(R2 - np.eye(3)).max() < 1e-4 and np.abs(T2).max() < 1e-4


def test_canonicalize_subsequence_matches_oracle(parsers, orc):
    from egogen_b200.primitive_batches import PRIMITIVE_KEYS, canonicalize_subsequence
    transl, pose, betas = _motion(4, 200)
    assert canonicalize_subsequence(parsers, betas, transl, pose, 150, 210) is None      # recording too short
    out = canonicalize_subsequence(parsers, betas, transl, pose, 30, 90)
    assert set(out) == set(PRIMITIVE_KEYS)
    T = 20
    assert out["trans"].shape == (T, 3) and out["poses"].shape == (T, 165) and out["joints"].shape == (T, 22, 3)
    assert out["marker_ssm2_67"].shape == (T, 67, 3) and out["marker_cmu_41"].shape == (T, 41, 3)
    assert out["transf_rotmat"].shape == (3, 3) and out["transf_transl"].shape == (1, 3)
    # oracle pipeline on the same frames
    tr = torch.as_tensor(transl[30:90:3]); po = torch.as_tensor(pose[30:90:3, :66]); be = torch.as_tensor(betas[:10]).reshape(1, 10)
    xb = torch.cat([tr, po, torch.zeros(T, 24)], dim=1)
    R, Tt = orc.get_new_coordinate(be, "male", xb[:1])
    xn = orc.update_transl_glorot(R.repeat(T, 1, 1), Tt.repeat(T, 1, 1), be, "male", xb)
    assert np.abs(out["transf_rotmat"] - R[0].numpy()).max() < 2e-5
    assert np.abs(out["trans"] - xn[:, :3].numpy()).max() < 5e-5 and np.abs(out["poses"][:, :3] - xn[:, 3:6].numpy()).max() < 5e-5
    assert np.array_equal(out["poses"][:, 3:], pose[30:90:3, 3:].astype(np.float32))
    assert np.abs(out["joints"] - orc.get_jts(be, "male", xn).numpy()).max() < 2e-4
    assert np.abs(out["marker_ssm2_67"] - orc.get_markers(be, "male", xn).reshape(T, 67, 3).numpy()).max() < 2e-4
    # canonical: first pelvis at the origin, hips along +x on the floor plane
    assert np.abs(out["joints"][0, 0]).max() < 1e-4
    hip = out["joints"][0, 2] - out["joints"][0, 1]
    assert hip[0] > 0 and abs(hip[1]) < 1e-4


def test_noise_augmented_batches(dev, parsers, tmp_path):
    from egogen_b200.primitive_batches import BatchGeneratorAMASSCanonicalized, canonicalize_subsequence, save_primitive
    for i in range(5):
        transl, pose, betas = _motion(10 + i, 100)
        save_primitive(str(tmp_path / "canon" / "s" / f"subseq_{i:05d}.npz"), canonicalize_subsequence(parsers, betas, transl, pose, 0, 60))
    gen = BatchGeneratorAMASSCanonicalized(str(tmp_path / "canon"), amass_subset_name=["s"], sample_rate=1, body_repr="ssm2_67",
                                           device=dev, parser=parsers["ssm2_67"])
    gen.get_rec_list(shuffle_seed=0, to_gpu=True)
    assert gen.data_all.is_cuda and gen.data_all.shape == (5, 20, 201)
    clean = gen.next_batch(4)
    gen.index_rec = 0
    same = gen.next_batch(4, noise=0.0)                                   # zero noise: the body model reproduces the file
    assert same.shape == (20, 4, 201) and (same - clean).abs().max().item() < 1e-4
    tail = gen.next_batch(4, noise=0.0)                                   # ragged last batch
    assert tail.shape == (20, 1, 201) and not gen.has_next_rec()
    gen.index_rec = 0
    noisy = gen.next_batch(4, noise=0.1)
    assert torch.isfinite(noisy).all() and (noisy - clean).abs().max().item() > 1e-3
    # the perturbation is constant over a sequence: the pelvis-relative marker motion keeps its smoothness
    assert (noisy[1:] - noisy[:-1]).abs().max().item() < 10 * (clean[1:] - clean[:-1]).abs().max().item() + 1e-3
