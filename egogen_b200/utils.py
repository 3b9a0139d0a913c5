"""Rollout writer - mirror of the reference's motion/crowd_ppo/utils.py:10-51 (save_rollout_results):
same pickle schema ({'motion': [{blended_marker, smplx_params, betas, gender, transf_rotmat, transf_transl,
pelvis_loc, mp_type}], 'wpath', 'navmesh_path', 'scene_path'}) consumed by the reference's vis.py /
gen_egobody_depth.py. calc_sdf lives in egogen_b200/sdf.py."""
import os
import pickle
import time

from .sdf import calc_sdf  # noqa: F401  (re-export under the reference's module name)

MP_KEYS = ["blended_marker", "smplx_params", "betas", "gender", "transf_rotmat", "transf_transl", "pelvis_loc",
           "mp_type"]


def save_rollout_results(scene, outmps, outfolder, man_id=None):
    os.makedirs(outfolder, exist_ok=True)
    wpath = scene["wpath"]
    node = {"motion": [], "wpath": wpath.detach().cpu().numpy() if hasattr(wpath, "detach") else wpath,
            "navmesh_path": scene.get("navmesh_path")}
    if "obj_id" in scene:
        node["obj_id"] = scene["obj_id"]
    if "obj_transform" in scene:
        node["obj_transform"] = (scene["obj_transform"],)      # 1-tuple, as the reference writes it (utils.py:27)
    if "scene_path" in scene:
        node["scene_path"] = scene["scene_path"]
    for mp in outmps:
        mp_node = {}
        for idx, key in enumerate(MP_KEYS):
            v = mp[idx]
            if key in ("gender", "mp_type", "betas", "transf_rotmat", "transf_transl"):
                mp_node[key] = v if isinstance(v, str) else v.detach().cpu().numpy()
            elif key == "smplx_params":
                mp_node[key] = v[0:1].detach().cpu().numpy()
            else:
                mp_node[key] = v[0].detach().cpu().numpy()
        node["motion"].append(mp_node)
    name = "motion_%s.pkl" % (str(time.time()) if man_id is None else man_id)
    with open(os.path.join(outfolder, name), "wb") as f:
        pickle.dump(node, f)
    return os.path.join(outfolder, name)
