"""Golden trajectories of the REFERENCE'S OWN CrowdEnv (motion/crowd_ppo/crowd_env_2f.py) - build container only.

CrowdEnv.reset / step / _canonicalize_2frame / _get_feature / _blend_params / _calc_egosensing and the reference's
SMPLXParser (forward_smplx, get_new_coordinate, update_transl_glorot, calc_calibrate_offset), GAMMAPrimitiveCombo.sample_prior
and calc_sdf run UNMODIFIED on the CPU. What the reference imports from absent third-party packages is supplied by:
  smplx.create(...)        -> the oracle's SMPL-X restatement on the surrogate body model (oracle.smplx_lbs.SMPLXOracle)
  torchgeometry            -> oracle.tgm
  vposer.encode(x).loc     -> oracle.nets.VPoserEncoderOracle
  shapely                  -> MiniShapely below: an independent float64 clip of a segment against a polygon with holes
                              (GEOS' LineString.intersection(Polygon) / Polygon.contains for this use)
  gymnasium / trimesh / pyrender / pytorch3d -> empty stubs (rendering is off)
and `.cuda()` / torch.cuda.FloatTensor are aliased to their CPU forms. The world (surrogate SMPL-X, seeded weights, box
scene, start bodies) is oracle.harness.build_oracle_world / sample_candidates_cpu, which the test rebuilds.

Run:  python tests/golden/gen_env_golden.py   ->  tests/golden/env_golden.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

from egogen_b200 import assets                                # noqa: E402
from oracle import harness, nets, tgm as oracle_tgm           # noqa: E402
from oracle.smplx_lbs import SMPLXOracle                      # noqa: E402


# ---- MiniShapely ----------------------------------------------------------------------------------------
class Point:
    geom_type = "Point"

    def __init__(self, xy):
        self.xy = np.asarray(xy, dtype=np.float64)
        self.coords = [tuple(self.xy)]

    def distance(self, other):
        return other.distance(self) if isinstance(other, LineString) else float(np.linalg.norm(self.xy - other.xy))


class MultiPoint:
    def __init__(self, pts):
        self.geoms = [Point(p) for p in np.asarray(pts, dtype=np.float64)]


class LineString:
    geom_type = "LineString"

    def __init__(self, coords):
        self.coords = [tuple(np.asarray(c, dtype=np.float64)) for c in coords]

    def distance(self, pt):
        a, b = np.asarray(self.coords[0]), np.asarray(self.coords[1])
        d = b - a
        t = np.clip(np.dot(pt.xy - a, d) / max(np.dot(d, d), 1e-300), 0.0, 1.0)
        return float(np.linalg.norm(a + t * d - pt.xy))

    def intersection(self, poly):
        return poly.clip(self)


class MultiLineString:
    geom_type = "MultiLineString"

    def __init__(self, parts):
        self.geoms = parts


class Polygon:
    """exterior ring + holes (closed rings [n,2]); interior = inside the exterior and outside every hole."""

    is_valid = True

    def __init__(self, shell, holes=None):
        self.rings = [np.asarray(r, dtype=np.float64) for r in [shell] + list(holes or [])]
        self.exterior = types.SimpleNamespace(coords=[tuple(p) for p in self.rings[0]])

    @staticmethod
    def _in_ring(ring, p):
        x, y = p
        inside = False
        for (x0, y0), (x1, y1) in zip(ring[:-1], ring[1:]):
            if (y0 > y) != (y1 > y) and x < (x1 - x0) * (y - y0) / (y1 - y0) + x0:
                inside = not inside
        return inside

    def _inside(self, p):
        return self._in_ring(self.rings[0], p) and not any(self._in_ring(h, p) for h in self.rings[1:])

    def contains(self, pt):
        return self._inside(pt.xy)

    def clip(self, line):
        a, b = np.asarray(line.coords[0]), np.asarray(line.coords[1])
        d = b - a
        ts = [0.0, 1.0]
        for ring in self.rings:
            for p, q in zip(ring[:-1], ring[1:]):
                e = q - p
                den = d[0] * e[1] - d[1] * e[0]
                if den == 0.0:
                    continue
                w = p - a
                t = (w[0] * e[1] - w[1] * e[0]) / den
                u = (w[0] * d[1] - w[1] * d[0]) / den
                if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
                    ts.append(t)
        ts = sorted(set(ts))
        parts = []
        for t0, t1 in zip(ts[:-1], ts[1:]):
            if t1 - t0 > 1e-12 and self._inside(a + 0.5 * (t0 + t1) * d):
                if parts and abs(parts[-1][1] - t0) < 1e-15:
                    parts[-1][1] = t1                 # merge pieces that only touch a vertex
                else:
                    parts.append([t0, t1])
        segs = [LineString([a + t0 * d, a + t1 * d]) for t0, t1 in parts]
        if len(segs) == 1:
            return segs[0]
        return MultiLineString(segs)


class MultiPolygon:
    """union_all of the other agents' boxes: kept as the list of boxes - for containment and for the first hit of a ray
    the union of (possibly overlapping) holes and the set of holes are the same thing"""

    def __init__(self, geoms):
        self.geoms = list(geoms)


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Space:
    def __init__(self, *a, **k):
        pass


stub("gymnasium", Env=object)
stub("gymnasium.spaces", Dict=_Space, Box=_Space)
for n in ["trimesh", "pyrender", "pytorch3d", "tensorboardX", "matplotlib", "matplotlib.pyplot", "omegaconf"]:
    stub(n)
sys.modules["tensorboardX"].SummaryWriter = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
stub("shapely", LineString=LineString, union_all=lambda polys: MultiPolygon(polys), is_valid=lambda g: True)
stub("shapely.geometry", MultiPoint=MultiPoint, Point=Point, Polygon=Polygon, MultiPolygon=MultiPolygon, mapping=None,
     LinearRing=None, multipolygon=types.SimpleNamespace(MultiPolygon=MultiPolygon), polygon=types.SimpleNamespace(Polygon=Polygon))
stub("shapely.plotting", plot_polygon=None)
sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]
sys.modules["shapely"].plotting = sys.modules["shapely.plotting"]
sys.modules["torchgeometry"] = oracle_tgm
for n in ["pytorch3d.structures", "pytorch3d.transforms", "human_body_prior", "human_body_prior.tools"]:
    stub(n)
stub("human_body_prior.tools.model_loader", load_vposer=None)
stub("exp_GAMMAPrimitive.utils.utils_canonicalize_babel", get_body_model=None, marker_ssm_67=list(range(67)))


class _BodyModel(torch.nn.Module):
    """what smplx.create(...) returns, as far as SMPLXParser uses it"""

    def __init__(self, arrays):
        super().__init__()
        self.m = SMPLXOracle(arrays)

    def forward(self, return_verts=True, **kw):
        n = kw["body_pose"].shape[0]
        z = lambda d: torch.zeros(n, d)
        return self.m.forward(betas=kw["betas"], global_orient=kw["global_orient"], body_pose=kw["body_pose"],
                              left_hand_pose=kw.get("left_hand_pose", z(12)), right_hand_pose=kw.get("right_hand_pose", z(12)),
                              transl=kw["transl"])


MODEL = assets.make_surrogate_smplx(seed=0)
stub("smplx", create=lambda *a, **k: _BodyModel(MODEL))
torch.Tensor.cuda = lambda self, *a, **k: self
_tensor_to = torch.Tensor.to


def _to_without_cuda(self, *a, **k):                       # get_map moves its grid with .to(device='cuda')
    if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
        k.pop("device")
    a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
    return _tensor_to(self, *a, **k) if (a or k) else self


torch.Tensor.to = _to_without_cuda
torch.cuda.FloatTensor = lambda *a, device=None: torch.FloatTensor(*a)

sys.path.insert(0, os.path.join(REF, "motion"))
os.chdir(os.path.join(REF, "motion"))                       # CrowdEnv opens data/smplx_vert_segmentation.json relatively (read only)
from models import baseops as ref_baseops                   # noqa: E402
from models import models_GAMMA_primitive as ref_gamma      # noqa: E402
ref_baseops.get_body_marker_path = lambda: os.path.join(REF, "motion", "data")
ref_baseops.get_body_model_path = lambda: ""
from crowd_ppo import crowd_env_2f as ref_env               # noqa: E402
from crowd_ppo import crowd_env_2f_box as ref_env_box       # noqa: E402   (box scenes: 2-D walkability-map penetration)
from crowd_ppo import crowd_env_crowd_eval as ref_env_crowd  # noqa: E402   (multi-agent: other agents' boxes are holes)


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def build_reference_env(world, sampler, finetuning, box=False):
    """cfg values: MPVAEPolicy_samp_collision.yaml (SDF env) / MPVAEPolicy_samp_collision_2.yaml (box env)"""
    cfg = AttrDict(args=AttrDict(gpu_index=0),
                   modelconfig=AttrDict(reproj_factor=0.5, body_repr="ssm2_67_condi_marker_map", map_res=16, map_extent=0.8),
                   trainconfig=AttrDict(goal_thresh=0.1, max_depth=13, random_rotation_range=0.0, pene_thres=3),
                   lossconfig=AttrDict(weight_skate=0.3, weight_floor=0.1, weight_face_target=0.1,
                                       weight_look_target=0.1 if box else 0.3, weight_success=0.5, weight_target_dist=1.0,
                                       weight_vp=0.1, weight_pene=0.1, pene_type="body"))
    combo = ref_gamma.GAMMAPrimitiveCombo(
        {"body_repr": "ssm2_67", "h_dim": 256, "z_dim": 128, "t_his": 2, "t_pred": 18, "use_drnn_mlp": True,
         "hdims_mlp": [512, 256], "residual": True},
        {"gender": "male", "h_dim": 128, "n_blocks": 10, "n_recur": 3, "body_repr": "ssm2_67", "actfun": "relu", "use_cont": True})
    combo.predictor.load_state_dict(world["env"].combo.predictor.state_dict())
    combo.regressor.load_state_dict(world["env"].combo.regressor.state_dict())
    genop = types.SimpleNamespace(model=combo.eval())
    dev = torch.device("cpu")
    parsers = [ref_baseops.SMPLXParser({"n_batch": n, "device": dev, "marker_placement": "ssm2_67"}) for n in (4, 8, 80)]
    vp = world["env"].vposer
    vposer = types.SimpleNamespace(encode=lambda x: types.SimpleNamespace(loc=vp.encode_loc(x)))
    init_env = (cfg, genop, genop, "", sampler, parsers[0], parsers[1], parsers[2], assets.feet_marker_idx(),
                parsers[0].marker, vposer, world["sdf"])
    if box == "crowd":      # sampler is the agent's start data here (crowd_env_crowd_eval.py:50-51)
        return ref_env_crowd.CrowdEnv(init_env[:-1] + ("0", "golden"), save_rollout=False, render=False)
    if box:
        return ref_env_box.CrowdEnv(init_env[:-1], save_rollout=False, render=False)
    return ref_env.CrowdEnv(init_env, save_rollout=False, render=False, finetuning=finetuning)


class Sampler:
    def __init__(self, wp, goal, betas, poly, navmesh=None):
        self.wp, self.goal, self.betas, self.poly, self.calls, self.navmesh = wp, goal, betas, poly, 0, navmesh

    def next_body(self, **kw):
        self.calls += 1
        assert self.calls <= 2, "start body rejected by the reference's collision test"
        start = torch.cat([self.wp[0, :2], self.goal[2:]]).reshape(1, 3)
        return {"motion_seed": {"transl": self.wp[:, :3].clone(), "global_orient": self.wp[:, 3:6].clone(),
                                "body_pose": self.wp[:, 6:69].clone()},
                "gender": "male", "betas": self.betas.clone(), "wpath": torch.cat([start, self.goal.reshape(1, 3)]),
                "shapely_poly": self.poly, "navmesh": self.navmesh}


out = {}
world = harness.build_oracle_world(0, sdf_res=64)
rings = assets.scene_polygon(assets.make_box_scene(0))
N_ENVS, N_STEPS = 3, 3
wp, goals, betas = harness.sample_candidates_cpu(world, N_ENVS, seed=5)
g = torch.Generator().manual_seed(17)
Z = torch.randn(N_ENVS, N_STEPS, 128, generator=g)
for fin in (0, 1):
    for e in range(N_ENVS):
        env = build_reference_env(world, Sampler(wp[e], goals[e], betas[e], Polygon(rings[0], rings[1:])), bool(fin))
        obs, _ = env.reset()
        rec = {"state": [obs["state"]], "ego": [obs["egosensing"]], "dist": [obs["dist"].reshape(1)], "time": [obs["time"]],
               "reward": [], "term": [], "seed": [env.body_param_seed[0]], "R0": [env.R0[0]], "T0": [env.T0[0]]}
        for s in range(N_STEPS):
            obs, rew, term, trunc, _ = env.step(Z[e, s].clone())
            rec["state"].append(obs["state"]); rec["ego"].append(obs["egosensing"]); rec["dist"].append(obs["dist"])
            rec["time"].append(obs["time"]); rec["reward"].append(torch.tensor(rew)); rec["term"].append(torch.tensor(term))
            rec["seed"].append(env.body_param_seed[0]); rec["R0"].append(env.R0[0]); rec["T0"].append(env.T0[0])
            assert not trunc
        for k, v in rec.items():
            out[f"f{fin}_e{e}_{k}"] = torch.stack([torch.as_tensor(x).detach().float() for x in v]).numpy()
        if fin == 0 and e == 0:
            out["feet_vids_sorted"] = np.array(sorted(env.feet_vids))
# ---- box-scene env (crowd_env_2f_box.py): walkability map from the navmesh triangles, bbox penetration count --------
tris = assets.scene_navmesh_triangles(assets.make_box_scene(0))                      # [F,3,2]
navmesh = types.SimpleNamespace(vertices=np.concatenate([tris.reshape(-1, 2), np.zeros((tris.shape[0] * 3, 1))], axis=1),
                                faces=np.arange(tris.shape[0] * 3).reshape(-1, 3))
for e in range(N_ENVS):
    env = build_reference_env(world, Sampler(wp[e], goals[e], betas[e], Polygon(rings[0], rings[1:]), navmesh), False, box=True)
    obs, _ = env.reset()
    rec = {"state": [obs["state"]], "ego": [obs["egosensing"]], "reward": [], "term": [], "seed": [env.body_param_seed[0]],
           "R0": [env.R0[0]], "T0": [env.T0[0]]}
    for s in range(N_STEPS):
        obs, rew, term, trunc, _ = env.step(Z[e, s].clone())
        rec["state"].append(obs["state"]); rec["ego"].append(obs["egosensing"])
        rec["reward"].append(torch.tensor(rew)); rec["term"].append(torch.tensor(term))
        rec["seed"].append(env.body_param_seed[0]); rec["R0"].append(env.R0[0]); rec["T0"].append(env.T0[0])
        if term:
            break
    for k, v in rec.items():
        out[f"box_e{e}_{k}"] = torch.stack([torch.as_tensor(x).detach().float() for x in v]).numpy()
# ---- 4-agent crowd scene (crowd_env_crowd_eval.py driven in the order of dummy_vector_env.py:29-128) -----------------
# every agent's marker box is a hole of the floor polygon for the others; the boxes are redistributed before EACH
# agent's step, so agent i already sees the new boxes of agents < i
A = 4
cw, cg, cb = harness.sample_candidates_cpu(world, A, seed=9)
for a in range(A):                                            # on a 0.8 m circle, walking through the centre
    ang = 2 * np.pi * a / A + 0.2
    pos = torch.tensor([0.8 * np.cos(ang), 0.8 * np.sin(ang)], dtype=torch.float32)
    cw[a, :, :2] = pos
    cg[a, :2] = -3.0 * pos / pos.norm()
envs = [build_reference_env(world, Sampler(cw[a], cg[a], cb[a], None).next_body(), False, box="crowd") for a in range(A)]


def distribute_holes():
    for i in range(A):
        envs[i].holes = [envs[j].bbox for j in range(A) if j != i]


distribute_holes()
rec = {k: [] for k in ("state", "ego", "reward", "term", "bbox", "seed", "T0")}
first = [e.reset()[0] for e in envs]
rec["state"].append(torch.stack([o["state"] for o in first])); rec["ego"].append(torch.stack([o["egosensing"] for o in first]))
rec["bbox"].append(torch.tensor(np.array([[e.bbox[0][0], e.bbox[0][1], e.bbox[2][0], e.bbox[2][1]] for e in envs], dtype=np.float32)))
Zc = torch.randn(2, A, 128, generator=torch.Generator().manual_seed(23)) * 0.5
for s in range(2):
    st, eg, rw, tm, bb, sd, t0 = [], [], [], [], [], [], []
    for a in range(A):
        distribute_holes()
        obs, rew, term, trunc, _ = envs[a].step(Zc[s, a].clone())
        st.append(obs["state"]); eg.append(obs["egosensing"]); rw.append(torch.tensor(rew)); tm.append(torch.tensor(term))
        bb.append(torch.tensor([envs[a].bbox[0][0], envs[a].bbox[0][1], envs[a].bbox[2][0], envs[a].bbox[2][1]], dtype=torch.float32))
        sd.append(envs[a].body_param_seed[0]); t0.append(envs[a].T0[0])
    for k, v in zip(("state", "ego", "reward", "term", "bbox", "seed", "T0"), (st, eg, rw, tm, bb, sd, t0)):
        rec[k].append(torch.stack([x.float() for x in v]))
for k, v in rec.items():
    out[f"crowd_{k}"] = torch.stack(v).numpy()
out.update(crowd_wp=cw.numpy(), crowd_goals=cg.numpy(), crowd_betas=cb.numpy(), crowd_Z=Zc.numpy())
# ---- a start that walks into the box: penetration terminates the fine-tuning episode (and the box-scene episode) ----
from scipy.spatial.transform import Rotation                 # noqa: E402
bx = assets.make_box_scene(0)["boxes"][0]
cxy = np.array([(bx[0] + bx[3]) / 2, (bx[1] + bx[4]) / 2])
rad = float(np.hypot(bx[3] - bx[0], bx[4] - bx[1]) / 2 + 0.3)
pw, pg, pb = harness.sample_candidates_cpu(world, 1, seed=33)
found = None
for k in range(12):
    for yk in range(6):
        ang, yaw = 2 * np.pi * k / 12, 2 * np.pi * yk / 6
        pos = cxy + rad * np.array([np.cos(ang), np.sin(ang)])
        if np.abs(pos).max() > 3.3:
            continue
        R = np.zeros((3, 3)); c_, s_ = np.cos(yaw), np.sin(yaw)
        R[0, 0] = c_; R[0, 2] = s_; R[1, 0] = s_; R[1, 2] = -c_; R[2, 1] = 1
        cand = pw[0].clone()
        cand[:, :2] = torch.tensor(pos, dtype=torch.float32)
        cand[:, 3:6] = torch.tensor(Rotation.from_matrix(R).as_rotvec(), dtype=torch.float32)
        goal = torch.tensor([2 * cxy[0] - pos[0], 2 * cxy[1] - pos[1], float(pg[0, 2])], dtype=torch.float32)
        try:
            env = build_reference_env(world, Sampler(cand, goal, pb[0], Polygon(rings[0], rings[1:])), True)
            obs, _ = env.reset()
        except AssertionError:
            continue
        rec = {"state": [obs["state"]], "reward": [], "term": []}
        for s in range(N_STEPS):
            obs, rew, term, _, _ = env.step(Z[0, s].clone())
            rec["state"].append(obs["state"]); rec["reward"].append(torch.tensor(rew)); rec["term"].append(torch.tensor(term))
            if term:
                break
        if rec["term"][-1] and len(rec["term"]) < 13:
            found = (cand, goal, rec)
            break
    if found:
        break
assert found is not None, "no start walks into the box within 3 steps"
cand, goal, rec = found
for k, v in rec.items():
    out[f"pen_{k}"] = torch.stack([torch.as_tensor(x).detach().float() for x in v]).numpy()
out.update(pen_wp=cand.numpy(), pen_goal=goal.numpy(), pen_betas=pb[0].numpy())
env = build_reference_env(world, Sampler(cand, goal, pb[0], Polygon(rings[0], rings[1:]), navmesh), False, box=True)
try:
    obs, _ = env.reset()
    rec = {"state": [obs["state"]], "reward": [], "term": []}
    for s in range(N_STEPS):
        obs, rew, term, _, _ = env.step(Z[0, s].clone())
        rec["state"].append(obs["state"]); rec["reward"].append(torch.tensor(rew)); rec["term"].append(torch.tensor(term))
        if term:
            break
    for k, v in rec.items():
        out[f"penbox_{k}"] = torch.stack([torch.as_tensor(x).detach().float() for x in v]).numpy()
except AssertionError:
    pass                                                     # the 2-D map test rejects this start: nothing to record
# the same search for the box-scene env, whose 2-D start test is stricter: a start it accepts that ends by penetration
found = None
for k in range(12):
    for yk in range(6):
        ang, yaw = 2 * np.pi * k / 12, 2 * np.pi * yk / 6
        pos = cxy + (rad + 0.25) * np.array([np.cos(ang), np.sin(ang)])
        if np.abs(pos).max() > 3.3:
            continue
        R = np.zeros((3, 3)); c_, s_ = np.cos(yaw), np.sin(yaw)
        R[0, 0] = c_; R[0, 2] = s_; R[1, 0] = s_; R[1, 2] = -c_; R[2, 1] = 1
        cand = pw[0].clone()
        cand[:, :2] = torch.tensor(pos, dtype=torch.float32)
        cand[:, 3:6] = torch.tensor(Rotation.from_matrix(R).as_rotvec(), dtype=torch.float32)
        goal = torch.tensor([2 * cxy[0] - pos[0], 2 * cxy[1] - pos[1], float(pg[0, 2])], dtype=torch.float32)
        try:
            env = build_reference_env(world, Sampler(cand, goal, pb[0], Polygon(rings[0], rings[1:]), navmesh), False, box=True)
            obs, _ = env.reset()
        except AssertionError:
            continue
        rec = {"state": [obs["state"]], "reward": [], "term": []}
        for s in range(N_STEPS):
            obs, rew, term, _, _ = env.step(Z[0, s].clone())
            rec["state"].append(obs["state"]); rec["reward"].append(torch.tensor(rew)); rec["term"].append(torch.tensor(term))
            if term:
                break
        if rec["term"][-1]:
            found = (cand, goal, rec)
            break
    if found:
        break
if found is not None:
    cand, goal, rec = found
    for k, v in rec.items():
        out[f"penbox_{k}"] = torch.stack([torch.as_tensor(x).detach().float() for x in v]).numpy()
    out.update(penbox_wp=cand.numpy(), penbox_goal=goal.numpy())
print("penetration case: steps", len(out["pen_term"]), "term", out["pen_term"], "box:", out.get("penbox_term"))
out.update(wp=wp.numpy(), goals=goals.numpy(), betas=betas.numpy(), Z=Z.numpy())
np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **out)
print("wrote env_golden.npz", len(out), "arrays;", "rewards env0:", out["f0_e0_reward"], "ego[0,:4]:", out["f0_e0_ego"][0, 0, :4])
