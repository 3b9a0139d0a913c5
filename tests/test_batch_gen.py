"""Host logic of the canonicalised-primitive batch generator (batch_gen_amass.py:61-429): the `.npz` schema,
the filters (frame rate, NaN), truncation, feature selection, batching order and shuffling. No body-model compute here."""
import numpy as np
import pytest
import torch

from egogen_b200.primitive_batches import (PRIMITIVE_KEYS, BatchGeneratorAMASSCanonicalized, body_feature,
                                          get_target_feature, save_primitive)


def _primitive(tag, T=30, gender="male", fps=120, nan=False):
    g = np.random.default_rng(tag)
    d = {"trans": g.normal(size=(T, 3)).astype(np.float32), "poses": g.normal(size=(T, 165)).astype(np.float32) * 0.2,
         "betas": g.normal(size=16).astype(np.float32), "gender": gender, "mocap_framerate": fps,
         "joints": g.normal(size=(T, 22, 3)).astype(np.float32), "marker_cmu_41": g.normal(size=(T, 41, 3)).astype(np.float32),
         "marker_ssm2_67": g.normal(size=(T, 67, 3)).astype(np.float32), "transf_rotmat": np.eye(3, dtype=np.float32),
         "transf_transl": np.array([[0.1, 0.2, 0.9]], dtype=np.float32)}
    d["marker_ssm2_67"][:, 0, 0] = tag          # tags that survive shuffling
    d["joints"][:, 1, 1] = tag
    if nan:
        d["poses"][3, 5] = np.nan
    return d


def build_dataset(tmp):
    """10 usable recordings in two subsets (2 female), one 60 fps and one NaN recording; shared with
    tests/golden/gen_primitive_batches_golden.py, which runs the reference's class on the same files."""
    import os
    root = os.path.join(str(tmp), "canon")
    for i in range(10):
        save_primitive(os.path.join(root, "setA" if i < 6 else "setB", f"subseq_{i:05d}.npz"),
                       _primitive(i + 1, gender="female" if i in (2, 7) else "male"))
    save_primitive(os.path.join(root, "setA", "subseq_slow.npz"), _primitive(50, fps=60))
    save_primitive(os.path.join(root, "setA", "subseq_nan.npz"), _primitive(51, nan=True))
    return root


@pytest.fixture()
def dataset(tmp_path):
    return build_dataset(tmp_path)


def test_schema_roundtrip(tmp_path):
    d = _primitive(3)
    save_primitive(str(tmp_path / "x" / "a.npz"), d)
    with np.load(str(tmp_path / "x" / "a.npz")) as f:
        assert set(f.files) == set(PRIMITIVE_KEYS)
        assert str(f["gender"]) == "male" and int(f["mocap_framerate"]) == 120
        assert f["transf_transl"].shape == (1, 3) and f["joints"].shape == (30, 22, 3)
    bad = dict(d); bad.pop("joints")
    with pytest.raises(KeyError):
        save_primitive(str(tmp_path / "x" / "b.npz"), bad)


@pytest.mark.parametrize("repr_,dim", [("ssm2_67", 201), ("cmu_41", 123), ("joints", 66), ("smpl_params", 69),
                                        ("ssm2_67_marker2tarloc", 402)])
def test_get_rec_list_filters_and_features(dataset, repr_, dim):
    gen = BatchGeneratorAMASSCanonicalized(dataset, sample_rate=1, body_repr=repr_, device="cpu")
    gen.get_rec_list(shuffle_seed=3)
    assert len(gen.rec_list) == 12 and len(gen.data_list) == 10          # 60 fps and NaN recordings dropped
    assert gen.data_all.shape == (10, 20, dim)                           # truncated to max_len = 20 frames
    assert gen.jts_all.shape == (10, 20, 22, 3) and gen.pose_all.shape == (10, 20, 66)
    assert gen.beta_all.shape == (10, 16) and sorted(gen.gender_all.tolist()).count("female") == 2
    b = gen.next_batch(4)
    assert b.shape == (20, 4, dim) and gen.index_rec == 4 and gen.has_next_rec()
    assert torch.equal(b, torch.as_tensor(gen.data_all[:4]).permute(1, 0, 2))
    gen.next_batch(4); gen.next_batch(4)
    assert not gen.has_next_rec()
    assert gen.get_all_data().shape == (20, 10, dim)


def test_bone_transform_and_bad_repr():
    d = _primitive(1, T=5)
    f = body_feature("bone_transform", d["trans"], d["poses"][:, :66], d["joints"], d["marker_cmu_41"], d["marker_ssm2_67"], None)
    assert f.shape == (5, 22, 6)
    with pytest.raises(NameError):
        body_feature("nope", d["trans"], d["poses"][:, :66], d["joints"], d["marker_cmu_41"], d["marker_ssm2_67"], None)


def test_target_feature_known_answers():
    joints = np.zeros((3, 22, 3), dtype=np.float64)
    joints[:, 0] = [[0, 0, 1.0], [1, 0, 1.0], [3, 4, 1.0]]
    mk = np.zeros((3, 2, 3)); mk[:, 1, 2] = 2.0
    vec, wpath_n, locn = get_target_feature(joints, mk, np.eye(3), np.array([[0.0, 0.0, 0.25]]))
    assert np.allclose(wpath_n[0], [0.6, 0.8]) and np.allclose(wpath_n[2], [0.0, 0.0])       # last frame: 0 / 1e-8
    assert joints[2, 0, 2] == 0.75                                       # the reference's in-place pelvis shift
    t = np.array([3, 4, 0.75])
    assert np.allclose(locn[0, 0], t / np.linalg.norm(t))
    assert np.allclose(locn[1, 1], (t - [0, 0, 2.0]) / np.linalg.norm(t - [0, 0, 2.0]))
    assert np.allclose(vec, 0.0)
    assert np.allclose(np.linalg.norm(locn, axis=-1), 1.0)


def test_subsets_and_shuffle_alignment(dataset):
    gen = BatchGeneratorAMASSCanonicalized(dataset, amass_subset_name=["setB"], sample_rate=1, body_repr="ssm2_67", device="cpu")
    gen.get_rec_list(shuffle_seed=0)
    assert len(gen.data_list) == 4 and gen.max_len == 20
    gen.reset_with_jts()                                                 # markers and joints permuted together
    d, j = gen.next_batch_with_jts(4)
    assert d.shape == (20, 4, 201) and j.shape == (20, 4, 22, 3)
    assert torch.equal(d[0, :, 0], j[0, :, 1, 1])                        # the per-recording tags still line up
    assert sorted(d[0, :, 0].tolist()) == [7.0, 8.0, 9.0, 10.0]
    # sample_rate strides the frames before truncation
    gen3 = BatchGeneratorAMASSCanonicalized(dataset, amass_subset_name=["setB"], sample_rate=3, body_repr="ssm2_67", device="cpu")
    gen3.get_rec_list(shuffle_seed=0)
    assert gen3.data_all.shape == (4, 10, 201)


def test_genderselection_and_next_sequence(dataset):
    gen = BatchGeneratorAMASSCanonicalized(dataset, amass_subset_name=["setB"], sample_rate=1, body_repr="ssm2_67", device="cpu")
    gen.get_rec_list(shuffle_seed=1)
    out = gen.next_batch_genderselection(3, "male")
    assert out is not None and len(out) == 6
    betas, feat, transl, glorot, thetas, jts = out
    assert betas.shape == (3, 30, 10) and feat.shape == (3, 30, 201) and transl.shape == (3, 30, 3)
    assert glorot.shape == (3, 30, 3) and thetas.shape == (3, 30, 63) and jts.shape == (3, 30, 66)
    assert 8.0 not in feat[:, 0, 0].tolist()                             # recording 8 (index 7) is female
    assert gen.next_batch_genderselection(3, "male") is None             # not enough left
    gen.index_rec = 0
    t_first = gen.next_batch_genderselection(2, "male", batch_first=False)[1]
    assert t_first.shape == (30, 2, 201)
    gen.index_rec = 0
    s = gen.next_sequence()
    assert set(s) == {"betas", "gender", "transl", "glorot", "poses", "body_feature", "transf_rotmat", "transf_transl", "pelvis_loc"}
    assert s["betas"].shape == (10,) and s["poses"].shape == (30, 63) and s["pelvis_loc"].shape == (30, 3) and gen.index_rec == 1


# ---- parity with the reference's own class (tests/golden/primitive_batches_golden.npz) ------------------------------
@pytest.fixture(scope="module")
def ref_golden(golden_dir):
    import os
    return np.load(os.path.join(golden_dir, "primitive_batches_golden.npz"))


def _by_tag(gen):
    return np.argsort(np.asarray(gen.jts_all)[:, 0, 1, 1])


def test_get_rec_list_matches_reference_class(dataset, ref_golden):
    """Every array the reference's get_rec_list builds (all body_repr, incl. the in-place pelvis shift that leaks into
    jts_all), record by record on the same files. cmu_41 is where the reference itself raises (:199)."""
    g = ref_golden
    for repr_ in ["ssm2_67", "joints", "smpl_params", "ssm2_67_marker2tarloc", "bone_transform"]:
        gen = BatchGeneratorAMASSCanonicalized(dataset, sample_rate=1, body_repr=repr_, device="cpu")
        gen.get_rec_list(shuffle_seed=11)                       # any order: records are matched by their tag
        o = _by_tag(gen)
        mine = gen.data_all[o]
        if repr_ == "ssm2_67_marker2tarloc":
            assert np.array_equal(mine[..., :201], g["data_ssm2_67"]) and int(g["tarloc_head_is_ssm2_67"]) == 1
            mine = mine[..., 201:]
        assert mine.shape == g[f"data_{repr_}"].shape, repr_
        assert np.array_equal(mine, g[f"data_{repr_}"]), repr_
        if repr_ == "ssm2_67":
            assert np.array_equal(gen.jts_all[o][:, 0, 1, 1], g["tags"])
            assert np.array_equal(gen.jts_all[o], g["jts_all"])
            assert np.array_equal(gen.beta_all[o], g["beta_all"]) and np.array_equal(gen.transl_all[o], g["transl_all"])
            assert gen.gender_all[o].tolist() == g["gender_all"].tolist()
            assert np.array_equal(gen.get_all_data().numpy(), gen.data_all.transpose(1, 0, 2)) and int(g["all_data_is_data_all_tmajor"]) == 1
    assert int(g["cmu_41_raises"]) == 1
    gen3 = BatchGeneratorAMASSCanonicalized(dataset, amass_subset_name=["setB"], sample_rate=3, body_repr="ssm2_67", device="cpu")
    gen3.get_rec_list(shuffle_seed=5)
    assert np.array_equal(gen3.data_all[_by_tag(gen3)], g["data_stride3"])


def test_genderselection_and_next_sequence_match_reference_class(dataset, ref_golden):
    g = ref_golden
    gen = BatchGeneratorAMASSCanonicalized(dataset, amass_subset_name=["setB"], sample_rate=1, body_repr="ssm2_67", device="cpu")
    gen.get_rec_list(shuffle_seed=1)
    gen.rec_list = sorted(gen.rec_list); gen.index_rec = 0      # the fixture was generated on the sorted file order
    sel = gen.next_batch_genderselection(3, "male")
    for k, t in zip(["betas", "feature", "transl", "glorot", "thetas", "jts"], sel):
        assert np.array_equal(t.numpy(), g[f"sel_{k}"]), k
    assert gen.index_rec == int(g["sel_index_after"])
    assert (gen.next_batch_genderselection(3, "male") is None) == bool(g["sel_second_is_none"])
    gen.index_rec = 0
    tm = gen.next_batch_genderselection(2, "male", batch_first=False)[1].numpy()
    assert np.array_equal(tm, g["sel_feature"][:2].transpose(1, 0, 2)) and int(g["sel_tmajor_is_transpose"]) == 1
    gen.index_rec = 0
    seq = gen.next_sequence()
    for k in ["betas", "transl", "glorot", "poses", "body_feature", "transf_rotmat", "transf_transl", "pelvis_loc"]:
        assert np.array_equal(np.asarray(seq[k]), g[f"seq_{k}"]), k
    assert str(seq["gender"]) == str(g["seq_gender"])


def test_target_feature_matches_reference_function(ref_golden):
    g = ref_golden
    J = g["tf_joints"].copy()
    v, w, l = get_target_feature(J, g["tf_markers"], np.eye(3), g["tf_transl"])
    assert np.array_equal(v, g["tf_vec"]) and np.array_equal(w, g["tf_wpath"]) and np.array_equal(l, g["tf_locn"])
    assert np.array_equal(J, g["tf_joints_after"]) and not np.array_equal(J, g["tf_joints"])
