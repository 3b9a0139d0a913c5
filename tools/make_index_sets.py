"""Extract the vertex-index sets the crowd_ppo hot path depends on from the reference's
data fixtures into one compact JSON shipped with the package (the GPU box has no /root/reference).

Sources (read-only, this container only):
  motion/data/SSM2.json                     -> 67 SSM2 marker names + SMPL-X vertex ids (main_ppo.py:296-300)
  motion/data/CMU.json                      -> 41 CMU marker ids (baseops.py:330-332)
  motion/data/smplx_vert_segmentation.json  -> feet vertex ids (crowd_env_2f.py:53-59)
Run:  python tools/make_index_sets.py
"""
import json, os, sys
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")
out = {}
for name, fn in (("ssm2_67", "SSM2.json"), ("cmu_41", "CMU.json")):
    d = json.load(open(f"{REF}/motion/data/{fn}"))["markersets"][0]["indices"]
    out[name] = {"names": list(d.keys()), "ids": [int(v) for v in d.values()]}
seg = json.load(open(f"{REF}/motion/data/smplx_vert_segmentation.json"))
feet = []
for part in seg:
    if part in ("leftToeBase", "rightToeBase", "leftFoot", "rightFoot"):
        feet.extend(seg[part])
out["feet_vids"] = sorted(set(int(v) for v in feet))
out["feet_markers"] = ["RHEE", "RTOE", "RRSTBEEF", "LHEE", "LTOE", "LRSTBEEF"]  # main_ppo.py:298
dst = os.path.join(os.path.dirname(__file__), "..", "egogen_b200", "data", "index_sets.json")
json.dump(out, open(dst, "w"))
print("wrote", dst, {k: (len(v["ids"]) if isinstance(v, dict) else len(v)) for k, v in out.items()})
