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
# per-vertex body-part label (index into PARTS) - gives the surrogate SMPL-X the real id -> part coherence
PARTS = ["hips", "leftUpLeg", "rightUpLeg", "spine", "leftLeg", "rightLeg", "spine1", "leftFoot", "rightFoot", "spine2",
         "leftToeBase", "rightToeBase", "neck", "leftShoulder", "rightShoulder", "head", "leftArm", "rightArm",
         "leftForeArm", "rightForeArm", "leftHand", "rightHand", "leftHandIndex1", "rightHandIndex1", "leftEye",
         "rightEye"]
nv = 1 + max(max(v) for v in seg.values())
label = [15] * nv                      # default: head
for pi, part in enumerate(PARTS):      # later parts override earlier ones where the segmentation overlaps
    for v in seg[part]:
        label[v] = pi
out["parts"] = PARTS
# run-length encoded labels keep the JSON small
rle, prev, cnt = [], label[0], 0
for l in label:
    if l == prev:
        cnt += 1
    else:
        rle.append([prev, cnt]); prev, cnt = l, 1
rle.append([prev, cnt])
out["part_rle"] = rle
out["feet_markers"] = ["RHEE", "RTOE", "RRSTBEEF", "LHEE", "LTOE", "LRSTBEEF"]  # main_ppo.py:298
dst = os.path.join(os.path.dirname(__file__), "..", "egogen_b200", "data", "index_sets.json")
json.dump(out, open(dst, "w"))
print("wrote", dst, {k: (len(v["ids"]) if isinstance(v, dict) else len(v)) for k, v in out.items()})
