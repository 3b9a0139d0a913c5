"""CrowdEnv.step / reset oracle (TEST INFRASTRUCTURE): PyTorch-CPU restatement of
motion/crowd_ppo/crowd_env_2f.py:78-317 (step), :320-415 (reset), :524-613 (_calc_egosensing),
:615-644 (_canonicalize_2frame), :680-727 (_get_feature), :729-739 (_blend_params), batched over E
environments with the reference's 4x batch duplication removed (only element [0] of the duplicated
batch is ever consumed: :169,174-175,185,194,202,219,229,233,235,287,296,312).

Pinned: tests/golden/gen_env_golden.py runs the reference's OWN CrowdEnv.reset / step (and its SMPLXParser,
GAMMAPrimitiveCombo, calc_sdf) on the CPU with the absent third-party packages supplied by the oracle's restatements;
tests/test_oracle_golden.py::test_env_oracle_matches_reference_crowd_env reproduces those trajectories (state 2e-7,
reward exact, ego-sensing 2e-4). shapely/GEOS itself is absent => the ray/polygon intersection follows the closed form
of SURVEY.md Appendix A6 in float64 and is cross-checked against an independent segment-vs-polygon clip in that
generator (GEOS parity unpinned for that piece).
"""
import numpy as np
import torch

from . import sdf as osdf


def rect_segments(rects) -> np.ndarray:
    """[H,4] (xmin,ymin,xmax,ymax) -> [4H,4] boundary segments in the order of the reference's bbox ring
    (crowd_env_crowd_eval.py:74-75)."""
    out = []
    for x0, y0, x1, y1 in np.asarray(rects, np.float64).reshape(-1, 4):
        c = [(x0, y0), (x1, y0), (x1, y1), (x0, y1), (x0, y0)]
        out += [[c[q][0], c[q][1], c[q + 1][0], c[q + 1][1]] for q in range(4)]
    return np.asarray(out, np.float64).reshape(-1, 4)


def egosensing(joints_w: torch.Tensor, segments: np.ndarray, ray_len: float = 7.0, holes=None) -> torch.Tensor:
    """joints_w [E,2,127,3] float32 world joints -> [E,2,32] float32 in [-1,1] (crowd_env_2f.py:524-613).
    holes [E,H,4] (optional): the other agents' rectangles cut out of the scene polygon
    (crowd_env_crowd_eval.py:796-822, Polygon(floor, holes=union of the rectangles)): an eye inside any rectangle is
    off the polygon (distance 0), and the first crossing of the union's boundary along a ray that starts outside
    every rectangle lies on an edge of one of them."""
    joint = joints_w.detach().cpu().numpy()
    E = joint.shape[0]
    out = np.zeros((E, 2, 32), np.float64)
    angle_grids = np.linspace(-np.pi / 2, np.pi / 2, 32)
    seg0 = np.asarray(segments, np.float64)
    for e in range(E):
        seg, hole_e = seg0, None
        if holes is not None:
            hole_e = np.asarray(holes[e], np.float64).reshape(-1, 4)
            seg = np.concatenate([seg0, rect_segments(hole_e)], axis=0)
        ax, ay, bx, by = seg[:, 0], seg[:, 1], seg[:, 2], seg[:, 3]
        ax0, ay0, bx0, by0 = seg0[:, 0], seg0[:, 1], seg0[:, 2], seg0[:, 3]
        j = joint[e]
        look_at = j[:, 57] - j[:, 23] + j[:, 56] - j[:, 24]
        look_at = look_at.astype(np.float64)
        look_at[:, -1] = 0.0
        look_at = look_at / np.linalg.norm(look_at, axis=-1, keepdims=True)
        eye_2d = (j[:, 23] + j[:, 24]) / 2
        eye_2d[:, -1] = 0.0
        for t in range(2):
            ex, ey = float(eye_2d[t, 0]), float(eye_2d[t, 1])
            # polygon.contains(eye): even-odd rule over exterior + hole rings
            cond = (ay0 > ey) != (by0 > ey)
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = ax0 + (ey - ay0) * (bx0 - ax0) / (by0 - ay0)
            inside = (np.count_nonzero(cond & (xi > ex)) % 2) == 1
            if hole_e is not None and inside:
                inside = not bool(((ex >= hole_e[:, 0]) & (ex <= hole_e[:, 2]) & (ey >= hole_e[:, 1]) & (ey <= hole_e[:, 3])).any())
            if not inside:
                continue                                  # end points = eye => distance 0
            l0, l1 = look_at[t, 0], look_at[t, 1]
            dx = l0 * np.cos(angle_grids) - l1 * np.sin(angle_grids)
            dy = l1 * np.cos(angle_grids) + l0 * np.sin(angle_grids)
            sx, sy = bx - ax, by - ay
            for r in range(32):
                den = dx[r] * sy - dy[r] * sx
                qx, qy = ax - ex, ay - ey
                with np.errstate(divide="ignore", invalid="ignore"):
                    tt = (qx * sy - qy * sx) / den
                    u = (qx * dy[r] - qy * dx[r]) / den
                ok = (den != 0.0) & (tt >= 0.0) & (u >= 0.0) & (u <= 1.0)
                tmin = min(ray_len, tt[ok].min()) if ok.any() else ray_len
                hx, hy = ex + tmin * dx[r], ey + tmin * dy[r]
                out[e, t, r] = np.sqrt((hx - ex) ** 2 + (hy - ey) ** 2)
    return torch.as_tensor((-1 + 2 * out / ray_len).astype(np.float32))


def get_feature(Y_l, pel, R0, T0, pt_wpath):
    """_get_feature (:680-727). Y_l [E,t,201], pel [E,t,3], R0 [E,3,3], T0 [E,1,3], pt_wpath [E,1,3].
    Returns (dist_xy, dist_xyz, fea_marker_3d_n)."""
    nb, nt = pel.shape[:2]
    Y_l = Y_l.reshape(nb, nt, -1, 3)
    pt_l = torch.einsum("bij,btj->bti", R0.permute(0, 2, 1), pt_wpath - T0)
    fea_xy = pt_l[:, :, :2] - pel[:, :, :2]
    fea_xyz = pt_l[:, :, :3] - pel[:, :, :3]
    dist_xy = torch.norm(fea_xy, dim=-1, keepdim=True).clip(min=1e-12)
    dist_xyz = torch.norm(fea_xyz, dim=-1, keepdim=True).clip(min=1e-12)
    fea_marker = pt_l[:, :, None, :] - Y_l
    dist_m_3d = torch.norm(fea_marker, dim=-1, keepdim=True).clip(min=1e-12)
    return dist_xy, dist_xyz, (fea_marker / dist_m_3d).reshape(nb, nt, -1)


def get_map(tris, R, T, res=16, extent=0.8, holes=None):
    """get_map (exp_GAMMAPrimitive/utils/batch_gen_amass.py:934-968): tris [F,3,2] = navmesh.vertices[faces, :2].
    Returns (points_local [b,res*res,3], local_map [b,res*res] with +1 walkable / -1 not, crowd_env_2f_box.py:769-770).
    holes [b,H,4] (optional) = _get_dynamic_map of crowd_env_crowd_eval.py:742-765: a grid point is walkable when the
    floor polygon contains it and no (closed) rectangle of another agent does."""
    b = R.shape[0]
    x = torch.linspace(-extent, extent, res)
    xv, yv = torch.meshgrid(x, x, indexing="ij")
    points = torch.stack([xv, yv, torch.zeros_like(xv)], dim=2).reshape(1, -1, 3).repeat(b, 1, 1)
    points_scene = torch.einsum("bij,bpj->bpi", R, points) + T
    p2 = points_scene[:, :, :2].reshape(b * res * res, 1, 2)
    tri = torch.as_tensor(tris, dtype=torch.float32)[None]            # [1,F,3,2]

    def sign(p1, p2_, p3):
        return (p1[:, :, 0] - p3[:, :, 0]) * (p2_[:, :, 1] - p3[:, :, 1]) - (p2_[:, :, 0] - p3[:, :, 0]) * (p1[:, :, 1] - p3[:, :, 1])
    d1 = sign(p2, tri[:, :, 0, :], tri[:, :, 1, :])
    d2 = sign(p2, tri[:, :, 1, :], tri[:, :, 2, :])
    d3 = sign(p2, tri[:, :, 2, :], tri[:, :, 0, :])
    has_neg = (d1 < 0) | (d2 < 0) | (d3 < 0)
    has_pos = (d1 > 0) | (d2 > 0) | (d3 > 0)
    inside = (~(has_neg & has_pos)).any(-1).reshape(b, res * res)
    if holes is not None:
        h = torch.as_tensor(holes, dtype=torch.float32).reshape(b, 1, -1, 4)
        px, py = points_scene[:, :, 0:1], points_scene[:, :, 1:2]
        in_hole = ((px >= h[..., 0]) & (py >= h[..., 1]) & (px <= h[..., 2]) & (py <= h[..., 3])).any(-1)
        inside = inside & ~in_hole
    local_map = inside.float()
    local_map[~inside] = -1
    return points, local_map


def map_penetration(marker_seed, points_local, local_map):
    """crowd_env_2f_box.py:279-292 (pene_type 'body'): marker_seed [b,t,67,3] local."""
    nb = marker_seed.shape[0]
    xy = marker_seed[:, :, :, :2]
    box_min = xy.amin(dim=[1, 2]).reshape(nb, 1, 2)
    box_max = xy.amax(dim=[1, 2]).reshape(nb, 1, 2)
    inside = ((points_local[:, :, :2] >= box_min).all(-1) & (points_local[:, :, :2] <= box_max).all(-1)).float()
    return (inside * (1 - local_map) * 0.5).sum(dim=1)


def blend_params(body_params, t_his=2):
    """_blend_params (:729-739) on [t,b,93], in place."""
    s = 6
    body_params[t_his, :, s:] = (body_params[t_his - 1, :, s:] + body_params[t_his + 1, :, s:]) / 2.0
    t = t_his + 1
    body_params[t, :, s:] = (body_params[t - 1, :, s:] + body_params[t + 1, :, s:]) / 2.0
    return body_params


class CrowdEnvOracle:
    """State: state [E,2,402], seed [E,2,93], R0 [E,3,3], T0 [E,1,3], betas [E,10], dist [E], steps [E], goal [E,3]."""

    W = dict(skate=0.3, floor=0.1, face=0.1, look=0.3, success=0.5, dist=1.0, vp=0.1)   # yaml :41-53

    def __init__(self, parser, combo, vposer, scene_sdf, segments, marker_ids, feet_marker_idx, feet_vids,
                 finetuning=False, max_depth=13, goal_thresh=0.1, reproj_factor=0.5, box_mode=False, navmesh_tris=None,
                 pene_thres=3, weight_pene_box=0.1, weight_look=None):
        self.parser, self.combo, self.vposer = parser, combo, vposer
        self.sdf, self.segments = scene_sdf, segments
        self.marker, self.feet_marker_idx, self.feet_vids = marker_ids, feet_marker_idx, feet_vids
        self.finetuning, self.max_depth, self.goal_thresh, self.rf = finetuning, max_depth, goal_thresh, reproj_factor
        self.box_mode, self.tris, self.pene_thres, self.w_pene_box = box_mode, navmesh_tris, pene_thres, weight_pene_box
        if weight_look is not None:
            self.W = dict(self.W, look=weight_look)
        # crowd dynamics (crowd_env_crowd_eval.py): holes [E,H,4] of the other agents, set by the vector env before every
        # step; crowd=True drops the penetration termination (:367) and the start-pose rejection (:391-405)
        self.holes, self.crowd, self.bbox = None, False, None

    @staticmethod
    def marker_bbox(marker_seed, R0, T0):
        """crowd_env_crowd_eval.py:345-352: xy bounding box of the 2-frame seed's markers in the world frame -> [E,4]."""
        E = marker_seed.shape[0]
        mw = torch.einsum("bij,btpj->btpi", R0, marker_seed.reshape(E, 2, -1, 3)) + T0[:, None, :, :]
        xy = mw[:, :, :, :2]
        return torch.cat([xy.amin(dim=[1, 2]), xy.amax(dim=[1, 2])], dim=1)

    def set_state(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v.clone() if torch.is_tensor(v) else v)

    def _lbs(self, xb_flat, betas_rows):
        """xb_flat [n,93]; betas_rows [n,10] (per body)."""
        bm = self.parser.bm_male
        return bm.forward(betas=betas_rows, global_orient=xb_flat[:, 3:6], body_pose=xb_flat[:, 6:69],
                          left_hand_pose=xb_flat[:, 69:81], right_hand_pose=xb_flat[:, 81:], transl=xb_flat[:, :3])

    def _update_transl_glorot(self, R, T, betas_rows, xb):
        """baseops.py:537-598 with per-row betas."""
        import torch.nn.functional as F
        from . import tgm
        n = xb.shape[0]
        z = torch.zeros
        delta_T = self.parser.bm_male.forward(betas=betas_rows, global_orient=z(n, 3), body_pose=xb[:, 6:69],
                                              left_hand_pose=z(n, 12), right_hand_pose=z(n, 12),
                                              transl=z(n, 3)).joints[:, 0, :]
        global_ori = tgm.angle_axis_to_rotation_matrix(xb[:, 3:6])[:, :3, :3]
        global_ori_new = torch.einsum("bij,bjk->bik", R.permute(0, 2, 1), global_ori)
        glorot = tgm.rotation_matrix_to_angle_axis(F.pad(global_ori_new, [0, 1])).view(-1, 3).contiguous()
        transl = torch.einsum("bij,bj->bi", R.permute(0, 2, 1), xb[:, :3] + delta_T - T[:, 0]) - delta_T
        return torch.cat([transl, glorot, xb[:, 6:]], dim=1)

    @torch.no_grad()
    def step(self, action_z):
        """action_z [E,128] -> dict(obs..., reward [E], terminated [E], terms [E,8])."""
        from .smplx_lbs import SMPLXParserOracle
        E = self.state.shape[0]
        self.steps = self.steps + 1
        t_his, nt = 2, 20
        X = self.state.permute(1, 0, 2)[:, :, :201]                      # [2,E,201]
        Xb = self.seed.permute(1, 0, 2)                                  # [2,E,93]
        betas18 = self.betas.unsqueeze(0).repeat(18, 1, 1)
        Y_gen, Yb_gen = self.combo.sample_prior(X, betas=betas18, z=action_z)
        Y = torch.cat((X, Y_gen), dim=0)
        Yb = blend_params(torch.cat((Xb, Yb_gen), dim=0).clone(), t_his)
        pred_markers = Y.reshape(nt, E, -1, 3).permute(1, 0, 2, 3)        # [E,20,67,3]
        pred_params = Yb.permute(1, 0, 2)                                 # [E,20,93]
        betas_rows = self.betas.unsqueeze(1).repeat(1, nt, 1).reshape(E * nt, 10)
        out = self._lbs(pred_params.reshape(E * nt, -1), betas_rows)
        joints_all = out.joints.reshape(E, nt, -1, 3)
        pred_joints = joints_all[:, :, :22]
        pelvis = pred_joints[:, :, 0]
        markers_proj = out.vertices[:, self.marker, :].reshape(E, nt, -1, 3)
        marker_b = self.rf * markers_proj + (1 - self.rf) * pred_markers
        # sdf penetration (:162-177)
        if self.box_mode:
            counts = torch.zeros(E, nt, dtype=torch.int64)
        else:
            verts = out.vertices.reshape(E, nt, -1, 3)
            verts_w = torch.einsum("bij,btpj->btpi", self.R0, verts) + self.T0[:, None, :, :]
            sdf_values = osdf.calc_sdf(verts_w.reshape(E * nt, -1, 3), self.sdf).reshape(E, nt, -1)
            sdf_values[:, :, self.feet_vids] = 0.0
            counts = sdf_values.lt(0.0).sum(dim=-1)
        num_inside = counts.sum(dim=1) / nt / 10
        num_inside_max = counts.max(dim=-1).values
        penetration = num_inside_max >= 40
        r_pene = torch.exp(-num_inside)
        # skate (:182-185)
        h = 1 / 40
        speed = torch.norm(marker_b[:, 2:] - marker_b[:, :-2], dim=-1) / 2.0 / h
        dist2skat = (speed[:, :, self.feet_marker_idx].amin(dim=-1) - 0.075).clamp(min=0).mean(dim=-1)
        r_skate = torch.exp(-dist2skat)
        # floor (:191-194)
        marker_w = torch.einsum("bij,btpj->btpi", self.R0, marker_b) + self.T0[:, None, :, :]
        dist2gp = torch.abs(marker_w[:, :, self.feet_marker_idx, 2].amin(dim=-1) - 0.02).mean(dim=-1)
        r_floor = torch.exp(-dist2gp)
        # vposer (:197-204)
        vp = self.vposer.encode_loc(pred_params[:, :, 6:69].reshape(E * nt, -1))
        vp_norm = torch.norm(vp.reshape(E, nt, -1), dim=-1).mean(dim=1)
        r_vp = torch.where(vp_norm > 11, torch.tensor(0.0), torch.tensor(0.05))
        # facing / looking (:206-229)
        joints_end = pred_joints[:, -1]
        x_axis = (joints_end[:, 2, :] - joints_end[:, 1, :]).clone()
        x_axis[:, -1] = 0
        x_axis = x_axis / torch.norm(x_axis, dim=-1, keepdim=True).clip(min=1e-12)
        z_axis = torch.tensor([[0.0, 0.0, 1.0]]).repeat(E, 1)
        b_ori = torch.cross(z_axis, x_axis, dim=-1)[:, :2]
        goal = self.goal.reshape(E, 1, 3)
        target_l = torch.einsum("bij,btj->bti", self.R0.permute(0, 2, 1), goal - self.T0)[:, :, :3]
        face = target_l[:, 0, :2] - pelvis[:, -1, :2]
        face = face / torch.norm(face, dim=-1, keepdim=True).clip(min=1e-12)
        r_face = (torch.einsum("bi,bi->b", face, b_ori) + 1) / 2.0
        eye_x = (joints_all[:, -1, 24] - joints_all[:, -1, 23]).clone()
        eye_x[:, -1] = 0
        eye_x = eye_x / torch.norm(eye_x, dim=-1, keepdim=True).clip(min=1e-12)
        look_at = torch.cross(z_axis, eye_x, dim=-1)[:, :2]
        r_look = (torch.einsum("bi,bi->b", face, look_at) + 1) / 2.0
        # distance / goal (:231-235)
        dist2target = torch.norm(target_l - pelvis, dim=-1).clip(min=1e-12)[:, -1]
        r_dist = self.dist - dist2target
        self.dist = dist2target
        r_goal = (self.dist < self.goal_thresh).float()
        # re-canonicalise (:238-265)
        seed = pred_params[:, -t_his:]
        R_, T_ = SMPLXParserOracle.new_coordinate_from_joints(
            self._lbs(seed[:, 0], self.betas).joints[:, :22])
        T0_new = torch.einsum("bij,btj->bti", self.R0, T_) + self.T0
        R0_new = torch.einsum("bij,bjk->bik", self.R0, R_)
        seed_new = self._update_transl_glorot(R_.repeat_interleave(t_his, 0), T_.repeat_interleave(t_his, 0),
                                              self.betas.repeat_interleave(t_his, 0),
                                              seed.reshape(E * t_his, -1)).reshape(E, t_his, -1)
        marker_seed = torch.einsum("bij,btpj->btpi", R_.permute(0, 2, 1), marker_b[:, -t_his:] - T_[..., None, :])
        pel_seed = torch.einsum("bij,btj->bti", R_.permute(0, 2, 1), pelvis[:, -t_his:] - T_)
        self.R0, self.T0 = R0_new, T0_new
        _, _, fea_marker = get_feature(marker_seed, pel_seed, self.R0, self.T0, goal)
        self.state = torch.cat([marker_seed.reshape(E, t_his, -1), fea_marker], dim=-1)
        self.seed = seed_new
        w_pene = 0.1 if self.finetuning else 1.0
        if self.box_mode:      # 2-D walkability-map penetration in the NEW frame (crowd_env_2f_box.py:279-295)
            pts_l, lmap = get_map(self.tris, self.R0, self.T0, holes=self.holes)
            num_pene = map_penetration(marker_seed, pts_l, lmap)
            penetration = num_pene > self.pene_thres
            self.bbox = self.marker_bbox(marker_seed, self.R0, self.T0)
            r_pene = torch.where(penetration, torch.tensor(0.0), torch.tensor(0.05))
            w_pene = self.w_pene_box
        W = self.W
        reward = r_skate * W["skate"] + r_floor * W["floor"] + r_face * W["face"] + r_look * W["look"] + \
            r_goal * W["success"] + r_dist * W["dist"] + r_pene * w_pene + r_vp * W["vp"]
        # ego-sensing on the new seed (:290-296)
        ja = self._lbs(self.seed.reshape(E * t_his, -1), self.betas.repeat_interleave(t_his, 0)).joints
        ja_w = torch.einsum("bij,btpj->btpi", self.R0, ja.reshape(E, t_his, -1, 3)) + self.T0[:, None, :, :]
        self.ego = egosensing(ja_w, self.segments, holes=self.holes)
        at_max = self.steps == self.max_depth
        terminated = (r_goal > 0) | at_max | (penetration if ((self.finetuning or self.box_mode) and not self.crowd)
                                              else torch.zeros(E, dtype=torch.bool))
        return dict(state=self.state, egosensing=self.ego, dist=1 / (dist2target + 1),
                    time=torch.as_tensor([1 - s / self.max_depth for s in self.steps.tolist()], dtype=torch.float32),
                    reward=reward, terminated=terminated, counts=counts, seed=self.seed, R0=self.R0, T0=self.T0,
                    terms=torch.stack([r_skate, r_floor, r_face, r_look, r_goal, r_dist, r_pene, r_vp], dim=1),
                    marker_b=marker_b, params=pred_params, Y=Y.permute(1, 0, 2))

    @torch.no_grad()
    def reset_from(self, world_params, goals, betas):
        """reset (:320-415) for explicit candidates: world_params [n,2,93], goals [n,3], betas [n,10].
        Returns dict with accept mask and the initial state of every candidate."""
        from .smplx_lbs import SMPLXParserOracle
        n, t_his = world_params.shape[0], 2
        R0, T0 = SMPLXParserOracle.new_coordinate_from_joints(self._lbs(world_params[:, 0], betas).joints[:, :22])
        seed = self._update_transl_glorot(R0.repeat_interleave(2, 0), T0.repeat_interleave(2, 0),
                                          betas.repeat_interleave(2, 0), world_params.reshape(n * 2, -1)).reshape(n, 2, -1)
        out = self._lbs(seed.reshape(n * 2, -1), betas.repeat_interleave(2, 0))
        marker_seed = out.vertices[:, self.marker, :].reshape(n, 2, -1)
        joints_all = out.joints.reshape(n, 2, -1, 3)
        pelvis = joints_all[:, :, 0]
        goal = goals.reshape(n, 1, 3)
        _, dist, fea_marker = get_feature(marker_seed, pelvis, R0, T0, goal)
        verts_w = torch.einsum("bij,btpj->btpi", R0, out.vertices.reshape(n, 2, -1, 3)) + T0[:, None, :, :]
        sdf_values = osdf.calc_sdf(verts_w.reshape(n * 2, -1, 3), self.sdf).reshape(n, 2, -1)
        sdf_values[:, :, self.feet_vids] = 0.0
        counts = sdf_values.lt(0.0).sum(dim=-1)
        accept = counts.sum(dim=1) == 0
        if self.box_mode:
            pts_l, lmap = get_map(self.tris, R0, T0, holes=self.holes)
            accept = map_penetration(marker_seed.reshape(n, 2, -1, 3), pts_l, lmap) == 0
        if self.crowd:
            accept = torch.ones(n, dtype=torch.bool)
        ja_w = torch.einsum("bij,btpj->btpi", R0, joints_all) + T0[:, None, :, :]
        ego = egosensing(ja_w, self.segments, holes=self.holes)
        state = torch.cat([marker_seed, fea_marker], dim=-1)
        return dict(accept=accept, state=state, seed=seed, R0=R0, T0=T0, dist=dist[:, 0, 0], egosensing=ego,
                    bbox=self.marker_bbox(marker_seed, R0, T0),
                    counts=counts, obs_dist=(1 / (dist + 1))[:, 0, 0])
