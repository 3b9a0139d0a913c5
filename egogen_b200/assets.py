"""Asset loaders and deterministic synthetic surrogates for the crowd_ppo hot path.

The licensed assets the reference downloads (SMPLX_MALE.npz, VPoser weights, room0_sdf.pkl,
C-VAE / regressor checkpoints; reference README.md:53-59) are absent in this environment, so
every array the path needs can be produced two ways:

* ``load_*``  - read the real file format (same keys / layouts the reference reads);
* ``make_*``  - seeded synthetic surrogate with the REAL shapes.

Everything here is host-side numpy / torch-CPU; nothing in this file touches the GPU.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np

V_SMPLX = 10475          # vertices
J_SMPLX = 55             # skeleton joints
N_JOINTS_OUT = 127       # 55 + 21 vertex joints + 51 landmarks
N_SHAPE = 20             # 10 betas + 10 expression coefficients
N_POSE_BASIS = 486       # 54 * 9
N_FACES = 20908
N_HAND_PCA = 12
N_LANDMARKS = 51

# SMPL-X kinematic tree (smplx 0.1.28 kintree_table[0])
SMPLX_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
     20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
     21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53], dtype=np.int32)

# smplx.vertex_ids['smplx'] in VertexJointSelector order [recalled from smplx 0.1.28]:
# nose, reye, leye, rear, lear, LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel,
# l{thumb,index,middle,ring,pinky}, r{thumb,index,middle,ring,pinky}
SMPLX_EXTRA_JOINT_VIDS = np.array(
    [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
     5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022], dtype=np.int32)

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def index_sets() -> dict:
    """Marker / feet vertex-id sets extracted from the reference fixtures (tools/make_index_sets.py)."""
    with open(os.path.join(_DATA_DIR, "index_sets.json")) as f:
        return json.load(f)


def marker_ids(placement: str = "ssm2_67"):
    return list(index_sets()[placement]["ids"])


def feet_marker_idx():
    s = index_sets()
    names = s["ssm2_67"]["names"]
    return [names.index(n) for n in s["feet_markers"]]


def feet_vids():
    return list(index_sets()["feet_vids"])


# ----------------------------------------------------------------------------------------------
# SMPL-X model arrays
# ----------------------------------------------------------------------------------------------
SMPLX_KEYS = ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights",
              "hand_comp_l", "hand_comp_r", "pose_mean", "extra_vids", "faces", "lmk_faces_idx",
              "lmk_bary")


def vertex_parts():
    """Per-vertex body-part label (index into index_sets()['parts']) decoded from the run-length form."""
    rle = index_sets()["part_rle"]
    return np.concatenate([np.full(c, l, dtype=np.int32) for l, c in rle])


# body part (order of index_sets()['parts']) -> SMPL-X joint the part hangs off
_PART_JOINT = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, -1, -2, 23, 24]


def make_surrogate_smplx(seed: int = 0, nnz_per_vertex: int = 4) -> Dict[str, np.ndarray]:
    """Synthetic SMPL-X-shaped model: real V / J / basis sizes, the real kinematic
    tree and the REAL vertex-id -> body-part map (motion/data/smplx_vert_segmentation.json via
    tools/make_index_sets.py), so that - like the licensed model - consecutive vertex ids lie next to each
    other on the surface and share their skinning joints (runs of 10-500 ids per part). Geometry is a
    capsule figure: every part's vertices spiral around its bone in id order. Skinning weights have
    ``nnz_per_vertex`` non-zeros on kinematic neighbours (own joint, parent, child, grand-parent) varying
    smoothly along the bone. Blend-shape bases are dense Gaussian noise of realistic magnitude.
    Keys are SMPLX_KEYS; float32 / int32.

    posedirs is stored the way smplx keeps it at run time: [486, V*3] row-major
    (smplx reshapes the npz [V,3,486] array to (-1,486).T in SMPL.__init__).
    """
    rng = np.random.default_rng(seed)
    V, J = V_SMPLX, J_SMPLX
    parents = SMPLX_PARENTS.copy()
    children = [[c for c in range(J) if parents[c] == j] for j in range(J)]
    depth = np.zeros(J, dtype=np.int32)
    for j in range(1, J):
        depth[j] = depth[parents[j]] + 1
    # rest skeleton: a y-up stick figure
    jrest = np.zeros((J, 3), dtype=np.float64)
    dirs = rng.normal(size=(J, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[1] = [1.0, -0.3, 0.0]; dirs[2] = [-1.0, -0.3, 0.0]; dirs[3] = [0.0, 1.0, 0.0]   # hips sideways
    for j in (4, 5, 7, 8):
        dirs[j] = [0.0, -1.0, 0.02]
    dirs[10] = dirs[11] = [0.0, -0.3, 1.0]
    for j in (6, 9, 12, 15):
        dirs[j] = [0.0, 1.0, 0.0]
    dirs[13] = [0.6, 0.8, 0.0]; dirs[14] = [-0.6, 0.8, 0.0]
    dirs[16] = dirs[18] = dirs[20] = [1.0, -0.2, 0.0]
    dirs[17] = dirs[19] = dirs[21] = [-1.0, -0.2, 0.0]
    dirs[22] = [0.0, -0.3, 1.0]
    dirs[23] = [0.3, 0.2, 0.9]; dirs[24] = [-0.3, 0.2, 0.9]      # eyeball joints: left (+x) / right (-x)
    for j in range(25, 40):
        dirs[j] = [1.0, -0.1, 0.1 * ((j - 25) // 3 - 2)]
    for j in range(40, 55):
        dirs[j] = [-1.0, -0.1, 0.1 * ((j - 40) // 3 - 2)]
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    seg_len = np.full(J, 0.12)
    seg_len[[4, 5, 7, 8]] = 0.40; seg_len[[16, 17, 18, 19]] = 0.26; seg_len[[10, 11]] = 0.12
    seg_len[[6, 9]] = 0.15; seg_len[[13, 14]] = 0.14; seg_len[[20, 21]] = 0.08
    seg_len[[22, 23, 24]] = 0.08
    seg_len[25:] = 0.03
    for j in range(1, J):
        jrest[j] = jrest[parents[j]] + seg_len[j] * dirs[j]
    jrest[:, 1] -= 0.35                                          # pelvis slightly below origin like SMPL-X
    # vertex -> joint from the real part labels; the two hand-finger parts are spread over the 15 finger joints
    part = vertex_parts()
    assign = np.zeros(V, dtype=np.int64)
    rank = np.zeros(V, dtype=np.float64)                          # position of the vertex inside its part, in [0,1)
    for pi, pj in enumerate(_PART_JOINT):
        ids = np.nonzero(part == pi)[0]
        if len(ids) == 0:
            continue
        r = np.arange(len(ids)) / len(ids)
        if pj >= 0:
            assign[ids] = pj
            rank[ids] = r
        else:                                                     # fingers: 15 joints, contiguous id blocks each
            base = 25 if pj == -1 else 40
            blk = np.minimum((r * 15).astype(np.int64), 14)
            assign[ids] = base + blk
            rank[ids] = r * 15 - blk
    # geometry: spiral around the bone from the joint towards its (first) child / along its own direction
    radius = np.where(depth <= 3, 0.11, 0.05)
    radius[15] = 0.10; radius[[23, 24]] = 0.012; radius[25:] = 0.008; radius[[7, 8, 10, 11]] = 0.04
    axis = dirs.copy()
    length = np.array([seg_len[children[j][0]] if children[j] else seg_len[j] for j in range(J)])
    length[15] = 0.22
    ref = np.where(np.abs(axis[:, [1]]) < 0.9, [[0.0, 1.0, 0.0]], [[1.0, 0.0, 0.0]])
    e1 = np.cross(axis, ref); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(axis, e1)
    a, t = assign, rank
    turns = 14.0
    ang = 2 * np.pi * turns * t
    v_template = (jrest[a] + (t * length[a])[:, None] * axis[a]
                  + (radius[a] * np.cos(ang))[:, None] * e1[a] + (radius[a] * np.sin(ang))[:, None] * e2[a]
                  + rng.normal(scale=0.002, size=(V, 3)))
    # eye-surface vertices used as joints 56/57 sit in front of the eyeball joints
    v_template[SMPLX_EXTRA_JOINT_VIDS[1]] = jrest[24] + [0.0, 0.0, 0.02]
    v_template[SMPLX_EXTRA_JOINT_VIDS[2]] = jrest[23] + [0.0, 0.0, 0.02]
    shapedirs = rng.normal(scale=0.01, size=(V, 3, N_SHAPE))
    posedirs = rng.normal(scale=1e-3, size=(N_POSE_BASIS, V * 3))
    # joint regressor: each row a softmax over 32 vertices of (or near) that joint
    J_regressor = np.zeros((J, V))
    for j in range(J):
        own = np.nonzero(assign == j)[0]
        if len(own) < 32:
            near = np.argsort(np.linalg.norm(v_template - jrest[j], axis=1))[:64]
            own = np.unique(np.concatenate([own, near]))
        pick = rng.choice(own, size=32, replace=False)
        w = np.exp(rng.normal(size=32)); w /= w.sum()
        J_regressor[j, pick] = w
    # skinning: own joint, parent, child (else grand-parent), grand-parent / sibling; smooth along the bone
    par = np.maximum(parents, 0)
    child = np.array([children[j][0] if children[j] else par[par[j]] for j in range(J)])
    gpar = par[par]
    cand = np.stack([a, par[a], child[a], gpar[a]], axis=1)[:, :max(nnz_per_vertex, 1)]
    wraw = np.stack([np.full(V, 1.0), 0.6 * (1 - t), 0.6 * t, np.full(V, 0.08)], axis=1)[:, :cand.shape[1]]
    wraw = wraw * np.exp(rng.normal(scale=0.05, size=wraw.shape))
    wraw /= wraw.sum(axis=1, keepdims=True)
    lbs_weights = np.zeros((V, J))
    for k in range(cand.shape[1]):
        np.add.at(lbs_weights, (np.arange(V), cand[:, k]), wraw[:, k])
    hand_comp_l = rng.normal(scale=0.1, size=(N_HAND_PCA, 45))
    hand_comp_r = rng.normal(scale=0.1, size=(N_HAND_PCA, 45))
    pose_mean = np.zeros(J * 3)
    pose_mean[75:120] = rng.normal(scale=0.1, size=45)          # flat_hand_mean=False: hand means live here
    pose_mean[120:165] = rng.normal(scale=0.1, size=45)
    faces = rng.integers(0, V, size=(N_FACES, 3))
    lmk_faces_idx = rng.integers(0, N_FACES, size=N_LANDMARKS)
    lmk_bary = rng.dirichlet(np.ones(3), size=N_LANDMARKS)
    return dict(
        v_template=v_template.astype(np.float32), shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32), J_regressor=J_regressor.astype(np.float32),
        parents=parents.astype(np.int32), lbs_weights=lbs_weights.astype(np.float32),
        hand_comp_l=hand_comp_l.astype(np.float32), hand_comp_r=hand_comp_r.astype(np.float32),
        pose_mean=pose_mean.astype(np.float32), extra_vids=SMPLX_EXTRA_JOINT_VIDS.copy(),
        faces=faces.astype(np.int32), lmk_faces_idx=lmk_faces_idx.astype(np.int32),
        lmk_bary=lmk_bary.astype(np.float32))


def make_stress_smplx(seed: int = 0) -> Dict[str, np.ndarray]:
    """The surrogate with the skinning structure of a HARD mesh for the tensor-core LBS kernel: 5..8 non-zero weights per
    vertex (the kernel caches four slots per vertex and fetches the rest through its uncached path) drawn from the
    vertex's kinematic neighbourhood (own joint, ancestors up to three levels, children, siblings), so that neighbouring
    vertices disagree about their joint sets and 80-vertex tiles meet more than the 10 distinct joints a tile may hold
    (the layout builder has to close tiles early). Everything else is the seed-0 surrogate."""
    m = make_surrogate_smplx(seed=0)
    rng = np.random.default_rng(1000 + seed)
    V, J = V_SMPLX, J_SMPLX
    parents = SMPLX_PARENTS.copy()
    children = [[c for c in range(J) if parents[c] == j] for j in range(J)]
    own = m["lbs_weights"].argmax(axis=1)
    pools = []
    for j in range(J):
        p = {j}
        a = j
        for _ in range(3):
            a = max(int(parents[a]), 0)
            p.add(a)
        p.update(children[j])
        for c in children[j]:
            p.update(children[c])
        if parents[j] >= 0:
            p.update(children[int(parents[j])])
        pools.append(np.array(sorted(p), dtype=np.int64))
    w = np.zeros((V, J))
    for v in range(V):
        pool = pools[own[v]]
        k = int(min(rng.integers(5, 9), len(pool)))
        js = rng.choice(pool, size=k, replace=False)
        if own[v] not in js:
            js[0] = own[v]
        wr = rng.uniform(0.05, 1.0, size=k)
        wr[js == own[v]] += 1.5
        w[v, js] = wr / wr.sum()
    m = dict(m)
    m["lbs_weights"] = w.astype(np.float32)
    return m


def load_smplx_npz(path: str, num_betas: int = 10, num_expression: int = 10,
                   num_pca_comps: int = N_HAND_PCA) -> Dict[str, np.ndarray]:
    """Read a licensed ``SMPLX_{MALE,FEMALE,NEUTRAL}.npz`` into the SMPLX_KEYS layout, the way
    smplx.create(..., model_type='smplx', ext='npz', num_pca_comps=12, flat_hand_mean=False)
    prepares its buffers (reference call site: motion/models/baseops.py:291-320)."""
    d = np.load(path, allow_pickle=True, encoding="latin1")
    sd = np.asarray(d["shapedirs"], dtype=np.float64)
    # smplx: shapedirs[:, :, :num_betas] ++ expr_dirs = shapedirs[:, :, 300:300+num_expression]
    shapedirs = np.concatenate([sd[:, :, :num_betas], sd[:, :, 300:300 + num_expression]], axis=-1)
    pd_ = np.asarray(d["posedirs"], dtype=np.float64)
    posedirs = pd_.reshape(-1, pd_.shape[-1]).T                  # [486, V*3]
    parents = np.asarray(d["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    pose_mean = np.zeros(J_SMPLX * 3)
    pose_mean[75:120] = np.asarray(d["hands_meanl"], dtype=np.float64)
    pose_mean[120:165] = np.asarray(d["hands_meanr"], dtype=np.float64)
    return dict(
        v_template=np.asarray(d["v_template"], np.float32), shapedirs=shapedirs.astype(np.float32),
        posedirs=np.ascontiguousarray(posedirs).astype(np.float32),
        J_regressor=np.asarray(d["J_regressor"], np.float32), parents=parents.astype(np.int32),
        lbs_weights=np.asarray(d["weights"], np.float32),
        hand_comp_l=np.asarray(d["hands_componentsl"][:num_pca_comps], np.float32),
        hand_comp_r=np.asarray(d["hands_componentsr"][:num_pca_comps], np.float32),
        pose_mean=pose_mean.astype(np.float32), extra_vids=SMPLX_EXTRA_JOINT_VIDS.copy(),
        faces=np.asarray(d["f"], np.int32), lmk_faces_idx=np.asarray(d["lmk_faces_idx"], np.int32),
        lmk_bary=np.asarray(d["lmk_bary_coords"], np.float32))


def get_smplx_model(gender: str = "male", body_model_path: Optional[str] = None, seed: int = 0):
    """Real model if ``<body_model_path>/smplx/SMPLX_<GENDER>.npz`` exists, else the surrogate."""
    if body_model_path:
        p = os.path.join(body_model_path, "smplx", f"SMPLX_{gender.upper()}.npz")
        if os.path.exists(p):
            return load_smplx_npz(p)
    return make_surrogate_smplx(seed=seed + (0 if gender == "male" else 1))


# ----------------------------------------------------------------------------------------------
# Scene SDF grids
# ----------------------------------------------------------------------------------------------
def make_box_scene(seed: int = 0, n_boxes: int = 1, floor_half: float = 4.0):
    """Random-box scene description: an 8x8 m floor at z=0 with
    ``n_boxes`` axis-aligned boxes, side U(0.5,2) m, centre U(-2,2)^2, height U(0.5,2)."""
    rng = np.random.default_rng(seed)
    boxes = []
    for _ in range(n_boxes):
        c = rng.uniform(-2.0, 2.0, size=2)
        s = rng.uniform(0.5, 2.0, size=2)
        h = rng.uniform(0.5, 2.0)
        boxes.append([c[0] - s[0] / 2, c[1] - s[1] / 2, 0.0, c[0] + s[0] / 2, c[1] + s[1] / 2, h])
    return dict(boxes=np.asarray(boxes, np.float32).reshape(-1, 6), floor_half=float(floor_half))


def rasterize_scene_sdf(scene: dict, D: int = 256, device: str = "cpu"):
    """Rasterise a box scene to the ``{'center','scale','sdf'}`` dict calc_sdf consumes
    (reference utils.py:54-84 / main_ppo.py:302-304). Stored sign convention follows the
    reference: calc_sdf returns ``-grid`` and treats negative as penetration, so the grid holds
    MINUS the free-space distance (negative in free space, positive inside obstacles / below the
    floor / outside the walls). Axis order sdf[ix, iy, iz]; grid sample positions are the
    align_corners=False cell centres ``((2i+1)/D - 1)``.
    """
    import torch
    fh = scene["floor_half"]
    center = torch.tensor([0.0, 0.0, fh - 1.0], dtype=torch.float32)   # z range [-1, 2*fh-1]
    scale = torch.tensor([1.0 / fh], dtype=torch.float32)
    lin = ((2.0 * torch.arange(D, dtype=torch.float32, device=device) + 1.0) / D - 1.0) * fh
    x = (lin + center[0]).view(D, 1, 1)
    y = (lin + center[1]).view(1, D, 1)
    z = (lin + center[2]).view(1, 1, D)
    d = torch.minimum(z.expand(D, D, D), (fh - x.abs()).expand(D, D, D))
    d = torch.minimum(d, (fh - y.abs()).expand(D, D, D))
    for b in torch.as_tensor(scene["boxes"], device=device):
        c = (b[:3] + b[3:]) / 2
        h = (b[3:] - b[:3]) / 2
        qx = (x - c[0]).abs() - h[0]
        qy = (y - c[1]).abs() - h[1]
        qz = (z - c[2]).abs() - h[2]
        outside = torch.sqrt(qx.clamp(min=0) ** 2 + qy.clamp(min=0) ** 2 + qz.clamp(min=0) ** 2)
        inside = torch.maximum(torch.maximum(qx, qy), qz).clamp(max=0)
        d = torch.minimum(d, outside + inside)
    return {"center": center.to(device), "scale": scale.to(device), "sdf": (-d).contiguous()}


def scene_polygon(scene: dict):
    """2-D walkable polygon of a box scene for ego-sensing: exterior ring = floor square,
    holes = box footprints. Returned as a list of closed rings [n_i, 2] float64 (first = exterior),
    the same information the reference keeps in a shapely Polygon (crowd_env_2f.py:381)."""
    fh = scene["floor_half"]
    rings = [np.array([[-fh, -fh], [fh, -fh], [fh, fh], [-fh, fh], [-fh, -fh]], np.float64)]
    for b in scene["boxes"]:
        x0, y0, _, x1, y1, _ = [float(t) for t in b]
        rings.append(np.array([[x0, y0], [x0, y1], [x1, y1], [x1, y0], [x0, y0]], np.float64))
    return rings


def scene_navmesh_triangles(scene: dict) -> np.ndarray:
    """Walkable-floor triangles [F,3,2] of a box scene - the role of the reference's navmesh_tight.ply /
    per-scene navmesh (`navmesh.vertices[navmesh.faces, :2]`, batch_gen_amass.py:948-950): the floor square is cut along
    every box edge and each cell that no box footprint covers becomes two triangles."""
    fh = scene["floor_half"]
    xs, ys = {-fh, fh}, {-fh, fh}
    for b in scene["boxes"]:
        xs.update([float(np.clip(b[0], -fh, fh)), float(np.clip(b[3], -fh, fh))])
        ys.update([float(np.clip(b[1], -fh, fh)), float(np.clip(b[4], -fh, fh))])
    xs, ys = sorted(xs), sorted(ys)
    tris = []
    for x0, x1 in zip(xs[:-1], xs[1:]):
        for y0, y1 in zip(ys[:-1], ys[1:]):
            cx, cy = 0.5 * (x0 + x1), 0.5 * (y0 + y1)
            if any(b[0] < cx < b[3] and b[1] < cy < b[4] for b in scene["boxes"]):
                continue
            tris.append([[x0, y0], [x1, y0], [x1, y1]])
            tris.append([[x0, y0], [x1, y1], [x0, y1]])
    return np.asarray(tris, np.float32)


def rings_to_segments(rings):
    """Flatten closed rings to a [S,4] float64 array of boundary segments (x0,y0,x1,y1)."""
    segs = []
    for r in rings:
        r = np.asarray(r, np.float64)
        segs.append(np.concatenate([r[:-1], r[1:]], axis=1))
    return np.ascontiguousarray(np.concatenate(segs, axis=0))


def load_wkb_polygon(path: str):
    """Parse the reference's pickled shapely polygon (motion/data/replica_room0_shapely.pkl) without
    shapely: the pickle payload is a WKB byte string; returns rings as in scene_polygon()."""
    import struct
    raw = open(path, "rb").read()
    # find the WKB blob: little-endian (0x01) polygon (type 3)
    idx = raw.find(b"\x01\x03\x00\x00\x00")
    if idx < 0:
        raise ValueError("no WKB polygon found in " + path)
    off = idx + 5
    nrings = struct.unpack_from("<I", raw, off)[0]; off += 4
    rings = []
    for _ in range(nrings):
        n = struct.unpack_from("<I", raw, off)[0]; off += 4
        pts = np.frombuffer(raw, dtype="<f8", count=2 * n, offset=off).reshape(n, 2).copy()
        off += 16 * n
        rings.append(pts)
    return rings


def load_scene_sdf(path: str):
    """``data/room0_sdf.pkl`` loader (main_ppo.py:302-304): dict of numpy arrays center/scale/sdf."""
    d = np.load(path, allow_pickle=True)
    return {k: np.asarray(v, np.float32) for k, v in d.items()}


# ----------------------------------------------------------------------------------------------
# Deterministic synthetic network weights (no checkpoints are available offline)
# ----------------------------------------------------------------------------------------------
def fill_params_(module, seed: int = 0, w_gain: float = 1.0):
    """Fill every parameter / float buffer of a torch module in place from a generator keyed by
    (seed, crc32(name)) so that two structurally different implementations with the SAME state_dict
    keys and shapes (reference class, oracle class, CUDA weight pack) get identical values without
    depending on construction order. Weights ~ N(0, gain/sqrt(fan_in)), biases ~ N(0, 0.05),
    BatchNorm running_var ~ U(0.5, 1.5)."""
    import zlib
    import torch
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if not t.dtype.is_floating_point:
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            if name.endswith("running_var"):
                v = torch.rand(t.shape, generator=g) + 0.5
            elif t.dim() >= 2:
                v = torch.randn(t.shape, generator=g) * (w_gain / (t.shape[-1] ** 0.5))
            else:
                v = torch.randn(t.shape, generator=g) * 0.05
                if name.endswith("bn1.weight") or name.endswith("bn2.weight"):
                    v = v + 1.0
            t.copy_(v.to(t.dtype))
    return module
