"""CPU-only cross-checks of the UNPINNED oracle pieces (third-party algorithms absent from /root/reference) against
independent implementations available in this image: scipy rotations, float64 re-derivations, torch autograd."""
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from oracle import tgm
from oracle.smplx_lbs import SMPLXOracle, batch_rodrigues


def test_tgm_aa_rotmat_roundtrip_vs_scipy():
    g = torch.Generator().manual_seed(0)
    aa = torch.randn(500, 3, generator=g) * 1.2
    aa[:5] *= 1e-5                                   # Taylor branch (theta^2 <= 1e-6)
    R = tgm.angle_axis_to_rotation_matrix(aa)[:, :3, :3]
    Rs = torch.as_tensor(Rotation.from_rotvec(aa.numpy().astype(np.float64)).as_matrix(), dtype=torch.float32)
    assert torch.allclose(R[5:], Rs[5:], atol=3e-6)   # tgm divides by (theta + 1e-6): ~1e-6 relative deviation
    assert torch.allclose(R[:5], Rs[:5], atol=1e-6)
    back = tgm.rotation_matrix_to_angle_axis(torch.nn.functional.pad(R, [0, 1]))
    bs = torch.as_tensor(Rotation.from_matrix(R.numpy().astype(np.float64)).as_rotvec(), dtype=torch.float32)
    assert torch.allclose(back, bs, atol=2e-5)


def test_cont6d_to_rotmat_is_orthonormal_and_matches_gram_schmidt():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 22, 6, generator=g)
    R = tgm.cont2rotmat(x)
    eye = torch.eye(3).expand(R.shape[0], 3, 3)
    assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(R.shape[0]), atol=1e-5)
    aa = tgm.cont2aa(x)
    R2 = tgm.angle_axis_to_rotation_matrix(aa.reshape(-1, 3))[:, :3, :3]
    assert torch.allclose(R2, R, atol=2e-5)


def test_rodrigues_matches_scipy():
    g = torch.Generator().manual_seed(2)
    r = torch.randn(300, 3, generator=g)
    R = batch_rodrigues(r)
    Rs = torch.as_tensor(Rotation.from_rotvec(r.numpy().astype(np.float64)).as_matrix(), dtype=torch.float32)
    assert torch.allclose(R, Rs, atol=2e-6)


def test_lbs_oracle_float64_rederivation(smplx_model):
    """Independent float64 numpy re-derivation of SMPL-X LBS (explicit per-joint chain, dense skinning)."""
    m = smplx_model
    o = SMPLXOracle(m)
    g = torch.Generator().manual_seed(3)
    n = 3
    go, bp = torch.randn(n, 3, generator=g) * 0.4, torch.randn(n, 63, generator=g) * 0.3
    lh, rh = torch.randn(n, 12, generator=g) * 0.5, torch.randn(n, 12, generator=g) * 0.5
    betas, tr = torch.randn(n, 10, generator=g), torch.randn(n, 3, generator=g)
    out = o.forward(betas, go, bp, lh, rh, tr)
    f64 = lambda k: np.asarray(m[k], np.float64)
    for i in range(n):
        pose = np.concatenate([go[i], bp[i], np.zeros(9), lh[i].double().numpy() @ f64("hand_comp_l"),
                               rh[i].double().numpy() @ f64("hand_comp_r")]).astype(np.float64) + f64("pose_mean")
        Rm = Rotation.from_rotvec(pose.reshape(-1, 3)).as_matrix()
        shape = np.concatenate([betas[i].double().numpy(), np.zeros(10)])
        v_shaped = f64("v_template") + f64("shapedirs") @ shape
        J = f64("J_regressor") @ v_shaped
        feat = (Rm[1:] - np.eye(3)).reshape(-1)
        v_posed = v_shaped + (feat @ f64("posedirs")).reshape(-1, 3)
        G = [None] * 55
        par = m["parents"]
        for j in range(55):
            T = np.eye(4); T[:3, :3] = Rm[j]; T[:3, 3] = J[j] - (J[par[j]] if j else 0)
            G[j] = T if j == 0 else G[par[j]] @ T
        A = []
        for j in range(55):
            Aj = G[j].copy(); Aj[:3, 3] -= G[j][:3, :3] @ J[j]; A.append(Aj)
        A = np.stack(A)
        Tv = np.einsum("vj,jab->vab", f64("lbs_weights"), A)
        verts = np.einsum("vab,vb->va", Tv[:, :3, :3], v_posed) + Tv[:, :3, 3] + tr[i].double().numpy()
        assert np.abs(out.vertices[i].double().numpy() - verts).max() < 5e-6
        joints = np.stack([G[j][:3, 3] for j in range(55)]) + tr[i].double().numpy()
        assert np.abs(out.joints[i, :55].double().numpy() - joints).max() < 5e-6
        assert np.abs(out.joints[i, 55:76].double().numpy() - verts[m["extra_vids"]]).max() < 5e-6


def test_gae_matches_closed_form():
    from oracle.ppo import gae_return
    rng = np.random.default_rng(0)
    T = 7
    v, vn, r = rng.normal(size=T), rng.normal(size=T), rng.normal(size=T)
    end = np.zeros(T); end[3] = 1; end[-1] = 1
    adv = gae_return(v, vn, r, end, 0.99, 0.95)
    d = r + 0.99 * vn - v
    exp = np.zeros(T)
    for i in range(T):
        acc, w = 0.0, 1.0
        for j in range(i, T):
            acc += w * d[j]
            if end[j]:
                break
            w *= 0.99 * 0.95
        exp[i] = acc
    assert np.allclose(adv, exp)


def test_start_body_oracle_invariants():
    """f-3 oracle (oracle/sampler.py, environments.py:1041-1131) on the reference's own motion seed fixture: after
    gen_init_body the body's forward axis points from start to target, the pelvis stands over the start point, the lowest
    joint of frame 0 touches the floor and wpath is snapped to the pelvis height."""
    import os
    import numpy as np
    import torch
    from egogen_b200 import assets
    from oracle.sampler import gen_init_body
    from oracle.smplx_lbs import SMPLXParserOracle
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "locomotion_seed_00343.npz"))
    parser = SMPLXParserOracle(assets.make_surrogate_smplx(seed=0), marker=assets.marker_ids())
    start, target = np.array([1.0, -2.0, 0.0], np.float32), np.array([-2.5, 1.5, 0.0], np.float32)
    r = gen_init_body(parser, start, target, d["betas"], d["poses"][3:5, 3:66], d["poses"][3:5, :3], d["trans"][3:5], yaw=0.0)
    j = r["joints"]
    x_axis = j[0, 2] - j[0, 1]; x_axis[2] = 0
    fwd = torch.cross(torch.tensor([0.0, 0.0, 1.0]), x_axis / x_axis.norm(), dim=0)
    t = torch.as_tensor(target - start); t = t / t.norm()
    assert torch.allclose(fwd / fwd.norm(), t, atol=1e-4)                     # faces the target (yaw jitter 0)
    assert torch.allclose(j[0, 0, :2], torch.as_tensor(start[:2]), atol=1e-5)  # pelvis over the start point
    assert abs(float(j[0, :, 2].min())) < 1e-5                                 # lowest joint on the floor
    assert torch.allclose(r["wpath"][0], j[0, 0]) and float(r["wpath"][1, 2]) == float(r["wpath"][0, 2])
    R = r["global_orient_matrix"]
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-5)
