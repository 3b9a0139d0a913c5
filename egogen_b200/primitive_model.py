"""Config -> model wiring - mirror of the reference's motion/crowd_ppo/primitive_model.py (load_model :74-96,
configure_model :56-72). Returns (cfg, genop_2frame_male, genop_2frame_female); both genders load the SAME male
regressor config, like the reference (MPVAEPolicy_samp_collision.yaml:59-60 -> MPVAECombo_samp_2frame.yml:2-3)."""
from __future__ import annotations

import os

from .crowd_env import default_cfg, load_cfg
from .models_gamma_primitive import GAMMAPrimitiveComboGenOP


def configure_model(cfg_name, gpu_index, seed, results_root="results/crowd_ppo"):
    """Checkpoints are looked up where the reference's ConfigCreator puts them:
    results/crowd_ppo/<predictor cfg>/checkpoints/epoch-400.ckp and .../<regressor cfg>/checkpoints/epoch-100.ckp."""
    op = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": gpu_index, "seed": seed})
    pdir = os.path.join(results_root, "MPVAE_samp20_2frame_rollout", "checkpoints")
    rdir = os.path.join(results_root, "MoshRegressor_v3_male", "checkpoints")
    op.build_model(load_pretrained_model=True, predictor_dir=pdir, regressor_dir=rdir, seed=seed)
    return op


def load_model(box=False, cfg_path=None, gpu_index=0, seed=0):
    if cfg_path is None:
        cand = "crowd_ppo/cfg_samp20/MPVAEPolicy_samp_collision_2.yaml" if box else \
            "crowd_ppo/cfg_samp20/MPVAEPolicy_samp_collision.yaml"
        cfg_path = cand if os.path.exists(cand) else None
    cfg = load_cfg(cfg_path) if cfg_path else default_cfg()
    male = configure_model("MPVAECombo_samp_2frame", gpu_index, seed)
    female = configure_model("MPVAECombo_samp_2frame", gpu_index, seed)
    return cfg, male, female
