"""calc_sdf oracle (TEST INFRASTRUCTURE). Follows motion/crowd_ppo/utils.py:54-84.

``calc_sdf`` below is the reference's op sequence verbatim in behaviour (F.grid_sample 5-D,
mode 'bilinear' == trilinear, align_corners=False default, padding_mode='border', result negated).
``calc_sdf_explicit`` re-derives ATen's index arithmetic so the integer
corner indices - the bit-exact parity target - are observable; it is checked against
``calc_sdf`` in tests/test_oracle_sdf.py.
"""
import torch
import torch.nn.functional as F


def calc_sdf(vertices: torch.Tensor, sdf_dict: dict) -> torch.Tensor:
    """vertices [B,P,3] world space -> [B,P]; negative = penetration (utils.py:54-84)."""
    sdf_centroid = sdf_dict["center"].reshape(1, 1, 3)
    sdf_scale = sdf_dict["scale"]
    sdf_grids = sdf_dict["sdf"].squeeze().unsqueeze(0).unsqueeze(0)
    batch_size, num_vertices, _ = vertices.shape
    v = vertices.reshape(1, -1, 3)
    v = (v - sdf_centroid) * sdf_scale
    vals = F.grid_sample(sdf_grids, v[:, :, [2, 1, 0]].view(1, batch_size * num_vertices, 1, 1, 3),
                         padding_mode="border", align_corners=False).reshape(batch_size, num_vertices)
    return -vals


def calc_sdf_explicit(vertices: torch.Tensor, sdf_dict: dict):
    """Same value as calc_sdf plus the base corner indices.

    Per axis (vertex x -> grid axis 0, y -> 1, z -> 2): i = ((p+1)*D - 1)/2, clamp to [0, D-1],
    i0 = floor(i); weights (i0+1-i) and (i-i0); corners with index > D-1 are skipped (their weight
    is 0 after the clamp). Returns (values [B,P], idx int32 [B,P,3]).
    """
    grid = sdf_dict["sdf"].squeeze()
    D0, D1, D2 = grid.shape
    c = sdf_dict["center"].reshape(1, 1, 3)
    s = sdf_dict["scale"]
    p = (vertices - c) * s

    def unnorm(q, D):
        i = ((q + 1.0) * D - 1.0) / 2.0
        return torch.clamp(i, 0.0, float(D - 1))

    ix, iy, iz = unnorm(p[..., 0], D0), unnorm(p[..., 1], D1), unnorm(p[..., 2], D2)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    out = torch.zeros_like(ix)
    for dx in (0, 1):
        wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
        for dy in (0, 1):
            wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
            for dz in (0, 1):
                wz = (z0 + 1 - iz) if dz == 0 else (iz - z0)
                xi, yi, zi = (x0 + dx).long(), (y0 + dy).long(), (z0 + dz).long()
                ok = (xi <= D0 - 1) & (yi <= D1 - 1) & (zi <= D2 - 1)
                g = grid[xi.clamp(max=D0 - 1), yi.clamp(max=D1 - 1), zi.clamp(max=D2 - 1)]
                out = out + torch.where(ok, g * (wx * wy * wz), torch.zeros_like(g))
    idx = torch.stack([x0, y0, z0], dim=-1).to(torch.int32)
    return -out, idx


def penetration_counts(sdf_values: torch.Tensor, feet_vids, nt: int):
    """crowd_env_2f.py:170-177 on sdf_values [E, nt, V] (one row per env, dup removed):
    feet vertices zeroed, per-frame counts of sdf<0; returns (count_per_frame int64 [E,nt],
    num_inside [E] = total/nt/10, num_inside_max [E])."""
    s = sdf_values.clone()
    s[:, :, feet_vids] = 0.0
    cnt = s.lt(0.0).sum(dim=-1)
    num_inside = cnt.sum(dim=1) / nt / 10
    return cnt, num_inside, cnt.max(dim=-1).values
