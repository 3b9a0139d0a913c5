"""Golden vectors for the primitive batch generator from the REFERENCE'S OWN class (build container only).

motion/exp_GAMMAPrimitive/utils/batch_gen_amass.py is imported with sys.modules stubs for the absent third-party packages
(smplx, pytorch3d, trimesh, pyrender, human_body_prior, torchgeometry) and for utils_canonicalize_babel (which builds body
models at import time); BatchGeneratorAMASSCanonicalized.__init__ also builds SMPL-X models, so the object is created
without it and given the attributes __init__ would set. Only file parsing / batching code runs:
  get_rec_list (every body_repr), _get_target_feature, get_all_data, next_batch_genderselection, next_sequence.
torch.cuda.FloatTensor is aliased to torch.FloatTensor (no GPU here). The dataset is the deterministic synthetic one of
tests/test_batch_gen.py (`_primitive(tag)`), rebuilt by the test; records are identified by the tag stored in
joints[:, 1, 1], so the fixture does not depend on directory order.

Run:  python tests/golden/gen_primitive_batches_golden.py   ->  tests/golden/primitive_batches_golden.npz
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


for name in ["smplx", "torchgeometry", "pytorch3d", "pytorch3d.structures", "pytorch3d.transforms", "trimesh", "pyrender",
             "human_body_prior", "human_body_prior.tools", "human_body_prior.tools.model_loader", "tensorboardX",
             "matplotlib", "matplotlib.pyplot", "omegaconf"]:
    stub(name)
sys.modules["human_body_prior.tools.model_loader"].load_vposer = None
sys.modules["tensorboardX"].SummaryWriter = object
stub("exp_GAMMAPrimitive.utils.utils_canonicalize_babel", get_body_model=None, marker_ssm_67=list(range(67)))
sys.path.insert(0, os.path.join(REF, "motion"))
ref = importlib.import_module("exp_GAMMAPrimitive.utils.batch_gen_amass")
torch.cuda.FloatTensor = torch.FloatTensor

from test_batch_gen import build_dataset            # noqa: E402


def make(root, subsets, body_repr, sample_rate=1):
    g = object.__new__(ref.BatchGeneratorAMASSCanonicalized)
    g.rec_list, g.index_rec, g.data_list, g.jts_list = [], 0, [], []
    g.amass_data_path, g.amass_subset_name, g.sample_rate = root, subsets, sample_rate
    g.body_repr, g.read_to_ram, g.max_len = body_repr, True, 20
    return g


out = {}
with tempfile.TemporaryDirectory() as tmp:
    root = build_dataset(tmp)
    for body_repr in ["ssm2_67", "joints", "smpl_params", "ssm2_67_marker2tarloc", "bone_transform"]:
        g = make(root, None, body_repr)
        g.get_rec_list(shuffle_seed=3)
        order = np.argsort(g.jts_all[:, 0, 1, 1])
        d = g.data_all[order].astype(np.float32)
        out[f"data_{body_repr}"] = d[..., 201:] if body_repr == "ssm2_67_marker2tarloc" else d      # first 201 = ssm2_67
        if body_repr == "ssm2_67_marker2tarloc":
            out["tarloc_head_is_ssm2_67"] = np.array(int(np.array_equal(d[..., :201], out["data_ssm2_67"])))
        if body_repr == "ssm2_67":
            out["tags"] = g.jts_all[order, 0, 1, 1].astype(np.float32)
            out["jts_all"] = g.jts_all[order].astype(np.float32)
            out["beta_all"] = g.beta_all[order].astype(np.float32)
            out["transl_all"] = g.transl_all[order].astype(np.float32)
            out["gender_all"] = np.array([str(x) for x in g.gender_all[order]])
            out["all_data_is_data_all_tmajor"] = np.array(int(np.array_equal(g.get_all_data().numpy(), g.data_all.transpose(1, 0, 2))))
    try:
        make(root, None, "cmu_41").get_rec_list(shuffle_seed=3)
        out["cmu_41_raises"] = np.array(0)
    except ValueError:
        out["cmu_41_raises"] = np.array(1)               # the 41-marker reshape with 67*3 (:199)
    g3 = make(root, ["setB"], "ssm2_67", sample_rate=3)
    g3.get_rec_list(shuffle_seed=0)
    out["data_stride3"] = g3.data_all[np.argsort(g3.jts_all[:, 0, 1, 1])].astype(np.float32)
    # same-gender batches and next_sequence on a fixed (sorted) file order
    g = make(root, ["setB"], "ssm2_67")
    g.get_rec_list(shuffle_seed=1)
    g.rec_list = sorted(g.rec_list)
    g.index_rec = 0
    sel = g.next_batch_genderselection(3, "male")
    for k, t in zip(["betas", "feature", "transl", "glorot", "thetas", "jts"], sel):
        out[f"sel_{k}"] = t.numpy()
    out["sel_index_after"] = np.array(g.index_rec)
    out["sel_second_is_none"] = np.array(int(g.next_batch_genderselection(3, "male") is None))
    g.index_rec = 0
    tm = g.next_batch_genderselection(2, "male", batch_first=False)[1].numpy()
    out["sel_tmajor_is_transpose"] = np.array(int(np.array_equal(tm, out["sel_feature"][:2].transpose(1, 0, 2))))
    g.index_rec = 0
    seq = g.next_sequence()
    for k in ["betas", "transl", "glorot", "poses", "body_feature", "transf_rotmat", "transf_transl", "pelvis_loc"]:
        out[f"seq_{k}"] = np.asarray(seq[k])
    out["seq_gender"] = np.array(str(seq["gender"]))
    # _get_target_feature on its own
    rng = np.random.default_rng(5)
    J = rng.normal(size=(7, 22, 3)); M = rng.normal(size=(7, 67, 3)); T = np.array([[0.3, -0.1, 0.8]])
    Jc = J.copy()
    v, w, l = ref.BatchGeneratorAMASSCanonicalized._get_target_feature(None, Jc, M, np.eye(3), T)
    out.update(tf_joints=J, tf_markers=M, tf_transl=T, tf_vec=v, tf_wpath=w, tf_locn=l, tf_joints_after=Jc)

np.savez_compressed(os.path.join(HERE, "primitive_batches_golden.npz"), **out)
print("wrote primitive_batches_golden.npz", {k: np.asarray(v).shape for k, v in out.items()})
