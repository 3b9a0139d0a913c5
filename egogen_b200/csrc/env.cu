// Vectorised crowd_ppo environment transition - replaces CrowdEnv.step / CrowdEnv.reset
// (reference motion/crowd_ppo/crowd_env_2f.py:78-317, 320-415) for E environments at once with the
// reference's 4x batch duplication removed (only element [0] of each duplicated batch is ever consumed).
//
// Pipeline per step (all on one stream, no host synchronisation):
//   eg_motion_sample_prior -> env_prepare_params (seed frames + _blend_params) -> eg_lbs_forward_sdf
//   (LBS + world transform + calc_sdf + feet skip + per-frame counts, vertices never materialised)
//   -> eg_vposer_encode -> env_reward_recanon_kernel (8 reward terms, termination, new canonical frame,
//   update_transl_glorot, marker/goal features -> next state) -> eg_lbs_forward (joints of the new
//   2-frame seed) -> env_egosensing_kernel (2 x 32 fp64 rays against the scene polygon).
#include <stdlib.h>

#include <vector>

#include "geom.cuh"

namespace eg {

constexpr int NT = 20;       // frames per primitive (2 history + 18 predicted)
constexpr int NM = 67;       // markers
constexpr int NJ = EG_SMPLX_JOINTS_OUT;

struct StepArgs {
  EgEnvConfig cfg;
  EgEnvBuffers b;
  const float* Y;          // [E,20,201]
  const float* params;     // [E,20,93]
  const int32_t* counts;   // [E,20]
  const float* joints;     // [E,20,127,3]
  const float* mproj;      // [E,20,67,3]
  const float* vp_loc;     // [E,20,32]
  const float* pelvis_rest;// [E,3]
  const float* tris;       // [n_tris,3,2] navmesh (pene_mode 1)
  int n_tris;
  // crowd dynamics (crowd_env_crowd_eval.py / dummy_vector_env.py): the other agents of the scene as axis-aligned
  // hole rectangles (xmin, ymin, xmax, ymax), this agent's own rectangle as output
  const float* holes;      // [E,n_holes,4] or null
  int n_holes;
  float* bbox_out;         // [E,4] or null
  int pene_terminates;     // 0: crowd eval (:367) never terminates on penetration
};

// the union of the other agents' marker bounding boxes is cut out of the floor polygon (crowd_env_crowd_eval.py:796-822);
// a point is off the walkable polygon when any closed rectangle holds it
__device__ __forceinline__ bool in_any_hole(const float* holes, int H, float x, float y) {
  for (int k = 0; k < H; ++k) {
    const float* r = holes + 4 * k;
    if (x >= r[0] && y >= r[1] && x <= r[2] && y <= r[3]) return true;
  }
  return false;
}

__device__ __forceinline__ float norm3_clip(float x, float y, float z) {
  return fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
}

// SMPLXParser.update_transl_glorot, torch branch (baseops.py:570-596) for one body
__device__ __forceinline__ void update_transl_glorot(const float* Rn, const float* Tn, const float* delta,
                                                     const float* xb, float* out) {
  float Rg[9], Rnew[9], aa[3];
  tgm_aa_to_rotmat(xb + 3, Rg);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      Rnew[i * 3 + k] = Rn[0 * 3 + i] * Rg[0 * 3 + k] + Rn[1 * 3 + i] * Rg[1 * 3 + k] + Rn[2 * 3 + i] * Rg[2 * 3 + k];
  tgm_rotmat_to_aa(Rnew, aa);
  float d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) d[j] = xb[j] + delta[j] - Tn[j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[i] = (Rn[0 * 3 + i] * d[0] + Rn[1 * 3 + i] * d[1] + Rn[2 * 3 + i] * d[2]) - delta[i];
  out[3] = aa[0]; out[4] = aa[1]; out[5] = aa[2];
}

// seed frames + _blend_params (crowd_env_2f.py:116-120, 729-739)
__global__ void __launch_bounds__(128)
env_prepare_params_kernel(const float* __restrict__ seed, float* __restrict__ params, int E) {
  const int e = blockIdx.x, i = threadIdx.x;
  if (i >= 93) return;
  float* p = params + (int64_t)e * NT * 93;
  const float p1 = seed[((int64_t)e * 2 + 1) * 93 + i];
  p[i] = seed[((int64_t)e * 2) * 93 + i];
  p[93 + i] = p1;
  if (i >= 6) {
    const float p2 = (p1 + p[3 * 93 + i]) / 2.0f;
    p[2 * 93 + i] = p2;
    p[3 * 93 + i] = (p2 + p[4 * 93 + i]) / 2.0f;   // second blend reads the first blend's output
  }
}

// marker / goal features of _get_feature (crowd_env_2f.py:680-707) for one marker
__device__ __forceinline__ void write_state_marker(float* st, int p, const float* m, const float* goal_l) {
  const float fx = goal_l[0] - m[0], fy = goal_l[1] - m[1], fz = goal_l[2] - m[2];
  const float d = norm3_clip(fx, fy, fz);
  st[p * 3 + 0] = m[0]; st[p * 3 + 1] = m[1]; st[p * 3 + 2] = m[2];
  st[201 + p * 3 + 0] = fx / d; st[201 + p * 3 + 1] = fy / d; st[201 + p * 3 + 2] = fz / d;
}

// torch.linspace(-extent, extent, res)[i] in fp32 (symmetric evaluation like ATen's kernel)
__device__ __forceinline__ float linspace_sym(float extent, int res, int i) {
  const float step = (extent - (-extent)) / (float)(res - 1);
  return i < res / 2 ? -extent + step * (float)i : extent - step * (float)(res - 1 - i);
}

// 2-D penetration of crowd_env_2f_box.py:279-295 + get_map (batch_gen_amass.py:934-968) for one env; called by all
// threads of the CTA. ms_xy: the markers of the 2-frame seed in the (new) local frame, [n_pts][2] in shared memory.
// Returns the number of local-map cells inside the markers' bounding box that no navmesh triangle covers.
__device__ float map_penetration(const float* Rl, const float* Tl, const float* ms_xy, int n_pts, const float* tris,
                                 int n_tris, int res, float extent, float* red /* shared, >= 8 floats */,
                                 const float* holes = nullptr, int n_holes = 0) {
  const int tid = threadIdx.x, nth = blockDim.x;
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = tid; i < n_pts; i += nth) {
    mnx = fminf(mnx, ms_xy[2 * i]); mxx = fmaxf(mxx, ms_xy[2 * i]);
    mny = fminf(mny, ms_xy[2 * i + 1]); mxy = fmaxf(mxy, ms_xy[2 * i + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  __shared__ float bb[8][4];
  if ((tid & 31) == 0) { bb[tid >> 5][0] = mnx; bb[tid >> 5][1] = mny; bb[tid >> 5][2] = mxx; bb[tid >> 5][3] = mxy; }
  __syncthreads();
  mnx = mny = INFINITY; mxx = mxy = -INFINITY;
  for (int w = 0; w < (nth >> 5); ++w) {
    mnx = fminf(mnx, bb[w][0]); mny = fminf(mny, bb[w][1]); mxx = fmaxf(mxx, bb[w][2]); mxy = fmaxf(mxy, bb[w][3]);
  }
  float cnt = 0.0f;
  for (int p = tid; p < res * res; p += nth) {
    const float lx = linspace_sym(extent, res, p / res), ly = linspace_sym(extent, res, p % res);   // meshgrid 'ij'
    const bool in_box = lx >= mnx && ly >= mny && lx <= mxx && ly <= mxy;
    if (!in_box) continue;
    const float wx = (Rl[0] * lx + Rl[1] * ly + Rl[2] * 0.0f) + Tl[0];
    const float wy = (Rl[3] * lx + Rl[4] * ly + Rl[5] * 0.0f) + Tl[1];
    bool walkable = false;
    for (int f = 0; f < n_tris && !walkable; ++f) {
      const float* t = tris + f * 6;
      const float d1 = (wx - t[2]) * (t[1] - t[3]) - (t[0] - t[2]) * (wy - t[3]);
      const float d2 = (wx - t[4]) * (t[3] - t[5]) - (t[2] - t[4]) * (wy - t[5]);
      const float d3 = (wx - t[0]) * (t[5] - t[1]) - (t[4] - t[0]) * (wy - t[1]);
      const bool has_neg = d1 < 0.f || d2 < 0.f || d3 < 0.f, has_pos = d1 > 0.f || d2 > 0.f || d3 > 0.f;
      walkable = !(has_neg && has_pos);
    }
    if (walkable && holes != nullptr && in_any_hole(holes, n_holes, wx, wy)) walkable = false;
    cnt += walkable ? 0.0f : 1.0f;              // inside * (1 - map) * 0.5 with map = +1 / -1
  }
  cnt = warp_sum(cnt);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = cnt;
  __syncthreads();
  float total = 0.0f;
  for (int w = 0; w < (nth >> 5); ++w) total += red[w];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(256)
env_reward_recanon_kernel(const StepArgs a) {
  const int e = blockIdx.x, tid = threadIdx.x;
  __shared__ float mb[NT * NM * 3];
  __shared__ float s_skate[18], s_floor[NT], s_vp[NT];
  __shared__ float Rn[9], Tn[3], R0n[9], T0n[3], goal_l2[3], delta[3];
  __shared__ float sc[12], ms_xy[2 * NM * 2], red[8];
  const EgEnvConfig& c = a.cfg;
  const float* R0 = a.b.R0 + (int64_t)e * 9;
  const float* T0 = a.b.T0 + (int64_t)e * 3;
  const float* goal = a.b.goal + (int64_t)e * 3;
  const float* jall = a.joints + (int64_t)e * NT * NJ * 3;

  // blended markers: reproj_factor * reprojected + (1 - reproj_factor) * predicted (:151-152)
  const float rf = c.reproj_factor, rf1 = 1.0f - c.reproj_factor;
  for (int i = tid; i < NT * NM * 3; i += blockDim.x) {
    const float v = rf * a.mproj[(int64_t)e * NT * NM * 3 + i] + rf1 * a.Y[(int64_t)e * NT * NM * 3 + i];
    mb[i] = v;
    if (a.b.out_markers) a.b.out_markers[(int64_t)e * NT * NM * 3 + i] = v;
  }
  if (a.b.out_pelvis && tid < NT * 3) a.b.out_pelvis[(int64_t)e * NT * 3 + tid] = jall[(tid / 3) * NJ * 3 + tid % 3];
  if (a.b.out_params)
    for (int i = tid; i < NT * 93; i += blockDim.x) a.b.out_params[(int64_t)e * NT * 93 + i] = a.params[(int64_t)e * NT * 93 + i];
  __syncthreads();

  if (tid < 18) {                       // foot skating (:182-185)
    float mn = INFINITY;
    for (int k = 0; k < 6; ++k) {
      const int p = c.feet_marker_idx[k];
      const float* m2 = mb + ((tid + 2) * NM + p) * 3;
      const float* m0 = mb + (tid * NM + p) * 3;
      const float dx = m2[0] - m0[0], dy = m2[1] - m0[1], dz = m2[2] - m0[2];
      mn = fminf(mn, sqrtf(dx * dx + dy * dy + dz * dz) / 2.0f / 0.025f);
    }
    s_skate[tid] = fmaxf(mn - 0.075f, 0.0f);
  } else if (tid >= 32 && tid < 32 + NT) {   // floor contact (:191-194)
    const int t = tid - 32;
    float mn = INFINITY;
    for (int k = 0; k < 6; ++k) {
      const float* m = mb + (t * NM + c.feet_marker_idx[k]) * 3;
      mn = fminf(mn, R0[6] * m[0] + R0[7] * m[1] + R0[8] * m[2] + T0[2]);
    }
    s_floor[t] = fabsf(mn - 0.02f);
  } else if (tid >= 64 && tid < 64 + NT) {   // VPoser latent norm (:197-200)
    const int t = tid - 64;
    const float* l = a.vp_loc + ((int64_t)e * NT + t) * 32;
    float s = 0.0f;
    for (int k = 0; k < 32; ++k) s += l[k] * l[k];
    s_vp[t] = sqrtf(s);
  }
  __syncthreads();

  if (tid == 0) {
    int total = 0, mx = 0;
    if (c.pene_mode == 0)
      for (int t = 0; t < NT; ++t) { const int v = a.counts[e * NT + t]; total += v; mx = max(mx, v); }
    const float num_inside = (float)total / (float)NT / 10.0f;
    const float r_pene = expf(-num_inside);
    const bool penetration = mx >= c.pene_terminate_count;
    float s = 0.f;
    for (int t = 0; t < 18; ++t) s += s_skate[t];
    const float r_skate = expf(-(s / 18.0f));
    s = 0.f;
    for (int t = 0; t < NT; ++t) s += s_floor[t];
    const float r_floor = expf(-(s / (float)NT));
    s = 0.f;
    for (int t = 0; t < NT; ++t) s += s_vp[t];
    const float r_vp = (s / (float)NT) > 11.0f ? 0.0f : 0.05f;
    // facing / looking at the goal (:206-229)
    const float* je = jall + (int64_t)19 * NJ * 3;
    float x0 = je[2 * 3 + 0] - je[1 * 3 + 0], x1 = je[2 * 3 + 1] - je[1 * 3 + 1];
    float n = fmaxf(sqrtf(x0 * x0 + x1 * x1 + 0.0f), 1e-12f);
    x0 /= n; x1 /= n;
    const float bo0 = -x1, bo1 = x0;                   // cross((0,0,1), x)[:2]
    float tl[3];
    const float g0 = goal[0] - T0[0], g1 = goal[1] - T0[1], g2 = goal[2] - T0[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) tl[i] = R0[0 * 3 + i] * g0 + R0[1 * 3 + i] * g1 + R0[2 * 3 + i] * g2;
    float f0 = tl[0] - je[0], f1 = tl[1] - je[1];
    n = fmaxf(sqrtf(f0 * f0 + f1 * f1), 1e-12f);
    f0 /= n; f1 /= n;
    const float r_face = ((f0 * bo0 + f1 * bo1) + 1.0f) / 2.0f;
    float e0 = je[24 * 3 + 0] - je[23 * 3 + 0], e1 = je[24 * 3 + 1] - je[23 * 3 + 1];
    n = fmaxf(sqrtf(e0 * e0 + e1 * e1 + 0.0f), 1e-12f);
    e0 /= n; e1 /= n;
    const float r_look = ((f0 * (-e1) + f1 * e0) + 1.0f) / 2.0f;
    // distance to the goal (:231-235)
    const float dist2 = norm3_clip(tl[0] - je[0], tl[1] - je[1], tl[2] - je[2]);
    const float r_dist = a.b.dist[e] - dist2;
    const float r_goal = dist2 < c.goal_thresh ? 1.0f : 0.0f;
    // reward sum / termination are finalised after the re-canonicalisation (the box-scene penetration model
    // needs the seed markers in the NEW frame); stash the terms
    sc[0] = r_skate; sc[1] = r_floor; sc[2] = r_face; sc[3] = r_look; sc[4] = r_goal; sc[5] = r_dist; sc[6] = r_pene;
    sc[7] = r_vp; sc[8] = penetration ? 1.0f : 0.0f;
    const int steps = a.b.steps[e] + 1;
    a.b.dist[e] = dist2;
    a.b.steps[e] = steps;
    a.b.obs_dist[e] = 1.0f / (dist2 + 1.0f);
    a.b.obs_time[e] = (float)(1.0 - (double)steps / (double)c.max_depth);
    // new canonical frame at the second-last body (:238-248)
    const float* jb = jall + (int64_t)18 * NJ * 3;
    new_coordinate(jb, jb + 3, jb + 6, Rn, Tn);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      T0n[i] = (R0[i * 3 + 0] * Tn[0] + R0[i * 3 + 1] * Tn[1] + R0[i * 3 + 2] * Tn[2]) + T0[i];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        R0n[i * 3 + k] = R0[i * 3 + 0] * Rn[0 * 3 + k] + R0[i * 3 + 1] * Rn[1 * 3 + k] + R0[i * 3 + 2] * Rn[2 * 3 + k];
    }
    const float h0 = goal[0] - T0n[0], h1 = goal[1] - T0n[1], h2 = goal[2] - T0n[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) goal_l2[i] = R0n[0 * 3 + i] * h0 + R0n[1 * 3 + i] * h1 + R0n[2 * 3 + i] * h2;
#pragma unroll
    for (int i = 0; i < 3; ++i) delta[i] = a.pelvis_rest[(int64_t)e * 3 + i];
  }
  __syncthreads();
  // all reads of the old R0/T0 happened before this barrier
  if (tid < 9) a.b.R0[(int64_t)e * 9 + tid] = R0n[tid];
  if (tid < 3) a.b.T0[(int64_t)e * 3 + tid] = T0n[tid];

  // new 2-frame seed in the new frame (:250-257)
  const float* p18 = a.params + ((int64_t)e * NT + 18) * 93;
  float* seed = a.b.seed + (int64_t)e * 2 * 93;
  if (tid < 2) update_transl_glorot(Rn, Tn, delta, p18 + tid * 93, seed + tid * 93);
  for (int i = tid; i < 2 * 87; i += blockDim.x) seed[(i / 87) * 93 + 6 + i % 87] = p18[(i / 87) * 93 + 6 + i % 87];
  // marker seed + goal features -> next state (:259-265)
  for (int i = tid; i < 2 * NM; i += blockDim.x) {
    const int k = i / NM, p = i % NM;
    const float* m = mb + ((18 + k) * NM + p) * 3;
    const float d0 = m[0] - Tn[0], d1 = m[1] - Tn[1], d2 = m[2] - Tn[2];
    float ml[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) ml[q] = Rn[0 * 3 + q] * d0 + Rn[1 * 3 + q] * d1 + Rn[2 * 3 + q] * d2;
    write_state_marker(a.b.state + ((int64_t)e * 2 + k) * 402, p, ml, goal_l2);
    ms_xy[2 * i] = ml[0]; ms_xy[2 * i + 1] = ml[1];
  }
  __syncthreads();
  float num_pene = 0.0f;
  if (c.pene_mode == 1)
    num_pene = map_penetration(R0n, T0n, ms_xy, 2 * NM, a.tris, a.n_tris, c.map_res, c.map_extent, red,
                               a.holes ? a.holes + (int64_t)e * a.n_holes * 4 : nullptr, a.n_holes);
  if (tid == 0 && a.bbox_out != nullptr) {
    // bbox of the new 2-frame seed's markers on the world xy plane (crowd_env_crowd_eval.py:345-352)
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = 0; i < 2 * NM; ++i) {
      const float* st = a.b.state + ((int64_t)e * 2 + i / NM) * 402 + (i % NM) * 3;
      const float wx = (R0n[0] * st[0] + R0n[1] * st[1] + R0n[2] * st[2]) + T0n[0];
      const float wy = (R0n[3] * st[0] + R0n[4] * st[1] + R0n[5] * st[2]) + T0n[1];
      mnx = fminf(mnx, wx); mxx = fmaxf(mxx, wx); mny = fminf(mny, wy); mxy = fmaxf(mxy, wy);
    }
    float* bb = a.bbox_out + (int64_t)e * 4;
    bb[0] = mnx; bb[1] = mny; bb[2] = mxx; bb[3] = mxy;
  }
  if (tid == 0) {
    float r_pene = sc[6];
    bool penetration = sc[8] != 0.0f;
    if (c.pene_mode == 1) {                     // crowd_env_2f_box.py:292-295
      penetration = num_pene > c.pene_thres;
      r_pene = penetration ? 0.0f : 0.05f;
    }
    float reward = sc[0] * c.w_skate + sc[1] * c.w_floor;
    reward += sc[2] * c.w_face;
    reward += sc[3] * c.w_look;
    reward += sc[4] * c.w_success;
    reward += sc[5] * c.w_dist;
    reward += r_pene * c.w_pene;
    reward += sc[7] * c.w_vp;
    const int steps = a.b.steps[e];             // already incremented above
    const bool term = (sc[4] > 0.0f) || (steps == c.max_depth) ||
                      ((c.finetuning || c.pene_mode == 1) && penetration && a.pene_terminates);   // box env always terminates on penetration (:325)
    a.b.reward[e] = reward;
    a.b.terminated[e] = term ? 1 : 0;
    if (a.b.goal_reached) a.b.goal_reached[e] = sc[4] > 0.0f ? 1 : 0;
    if (a.b.reward_terms) {
      float* rt = a.b.reward_terms + (int64_t)e * 8;
      rt[0] = sc[0]; rt[1] = sc[1]; rt[2] = sc[2]; rt[3] = sc[3]; rt[4] = sc[4]; rt[5] = sc[5]; rt[6] = r_pene; rt[7] = sc[7];
    }
  }
}

// reset: canonical frame of the sampled 2-frame world seed (_canonicalize_2frame, :615-644)
__global__ void __launch_bounds__(32)
env_reset_canon_kernel(const float* __restrict__ wp, const float* __restrict__ joints_w,
                       const float* __restrict__ pelvis_rest, int n, float* __restrict__ R0c,
                       float* __restrict__ T0c, float* __restrict__ seedc) {
  const int i = blockIdx.x, tid = threadIdx.x;
  __shared__ float Rn[9], Tn[3];
  if (tid == 0) {
    const float* j = joints_w + (int64_t)i * 2 * NJ * 3;      // frame 0
    new_coordinate(j, j + 3, j + 6, Rn, Tn);
    for (int k = 0; k < 9; ++k) R0c[(int64_t)i * 9 + k] = Rn[k];
    for (int k = 0; k < 3; ++k) T0c[(int64_t)i * 3 + k] = Tn[k];
  }
  __syncwarp();
  if (tid < 2) update_transl_glorot(Rn, Tn, pelvis_rest + (int64_t)i * 3, wp + ((int64_t)i * 2 + tid) * 93,
                                    seedc + ((int64_t)i * 2 + tid) * 93);
  for (int q = tid; q < 2 * 87; q += 32)
    seedc[((int64_t)i * 2 + q / 87) * 93 + 6 + q % 87] = wp[((int64_t)i * 2 + q / 87) * 93 + 6 + q % 87];
}

// reset: accept the candidate if no non-feet vertex penetrates (:379-380) and commit it to its env slot
__global__ void __launch_bounds__(256)
env_reset_commit_kernel(EgEnvBuffers b, const int32_t* __restrict__ env_ids, int n,
                        const int32_t* __restrict__ counts, const float* __restrict__ R0c,
                        const float* __restrict__ T0c, const float* __restrict__ seedc,
                        const float* __restrict__ joints_c, const float* __restrict__ markers_c,
                        const float* __restrict__ goals, const float* __restrict__ betas_c,
                        int32_t* __restrict__ accept, EgEnvConfig cfg, const float* __restrict__ tris, int n_tris,
                        int check_start, float* __restrict__ bbox_out, const uint8_t* __restrict__ mask) {
  const int i = blockIdx.x, tid = threadIdx.x;
  __shared__ float ms_xy[2 * NM * 2], red[8];
  if (mask != nullptr && mask[i] == 0) {        // masked reset: this slot keeps running its episode
    if (tid == 0) accept[i] = 0;
    return;
  }
  bool ok;
  if (!check_start) {                           // crowd eval: fixed start data, no rejection (crowd_env_crowd_eval.py:391-405)
    ok = true;
  } else if (cfg.pene_mode == 1) {                     // crowd_env_2f_box.py reset: start pose must not cover unwalkable cells
    for (int q = tid; q < 2 * NM; q += blockDim.x) {
      ms_xy[2 * q] = markers_c[((int64_t)i * 2 * NM + q) * 3];
      ms_xy[2 * q + 1] = markers_c[((int64_t)i * 2 * NM + q) * 3 + 1];
    }
    __syncthreads();
    ok = map_penetration(R0c + (int64_t)i * 9, T0c + (int64_t)i * 3, ms_xy, 2 * NM, tris, n_tris, cfg.map_res,
                         cfg.map_extent, red) == 0.0f;
  } else {
    ok = (counts[i * 2] + counts[i * 2 + 1]) == 0;
  }
  if (tid == 0) accept[i] = ok ? 1 : 0;
  if (!ok) return;
  const int e = env_ids ? env_ids[i] : i;
  __shared__ float goal_l[3];
  const float* R0 = R0c + (int64_t)i * 9;
  const float* T0 = T0c + (int64_t)i * 3;
  if (tid == 0) {
    const float* goal = goals + (int64_t)i * 3;
    const float g0 = goal[0] - T0[0], g1 = goal[1] - T0[1], g2 = goal[2] - T0[2];
    for (int q = 0; q < 3; ++q) goal_l[q] = R0[0 * 3 + q] * g0 + R0[1 * 3 + q] * g1 + R0[2 * 3 + q] * g2;
    const float* pel = joints_c + (int64_t)i * 2 * NJ * 3;    // pelvis of frame 0
    const float d = norm3_clip(goal_l[0] - pel[0], goal_l[1] - pel[1], goal_l[2] - pel[2]);
    b.dist[e] = d;
    b.obs_dist[e] = 1.0f / (d + 1.0f);
    b.obs_time[e] = 1.0f;
    b.steps[e] = 0;
    for (int q = 0; q < 3; ++q) b.goal[(int64_t)e * 3 + q] = goal[q];
    for (int q = 0; q < 10; ++q) b.betas[(int64_t)e * 10 + q] = betas_c[(int64_t)i * 10 + q];
  }
  __syncthreads();
  if (tid < 9) b.R0[(int64_t)e * 9 + tid] = R0[tid];
  if (tid < 3) b.T0[(int64_t)e * 3 + tid] = T0[tid];
  if (tid == 0 && bbox_out != nullptr) {        // crowd_env_crowd_eval.py:66-75
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int q = 0; q < 2 * NM; ++q) {
      const float* m = markers_c + ((int64_t)i * 2 * NM + q) * 3;
      const float wx = (R0[0] * m[0] + R0[1] * m[1] + R0[2] * m[2]) + T0[0];
      const float wy = (R0[3] * m[0] + R0[4] * m[1] + R0[5] * m[2]) + T0[1];
      mnx = fminf(mnx, wx); mxx = fmaxf(mxx, wx); mny = fminf(mny, wy); mxy = fmaxf(mxy, wy);
    }
    float* bb = bbox_out + (int64_t)e * 4;
    bb[0] = mnx; bb[1] = mny; bb[2] = mxx; bb[3] = mxy;
  }
  for (int q = tid; q < 2 * 93; q += blockDim.x) b.seed[(int64_t)e * 2 * 93 + q] = seedc[(int64_t)i * 2 * 93 + q];
  for (int q = tid; q < 2 * NM; q += blockDim.x) {
    const int k = q / NM, p = q % NM;
    write_state_marker(b.state + ((int64_t)e * 2 + k) * 402, p, markers_c + (((int64_t)i * 2 + k) * NM + p) * 3, goal_l);
  }
}

// _calc_egosensing (:524-613) in closed form: 2 frames x 32 rays, fp64.
// joints: local-frame joints of the 2-frame seed, item i frames at joints[(i*2+t)*127*3];
// world = R0[slot] j + T0[slot] in fp32 like the reference einsum, then numpy float64 arithmetic.
__global__ void __launch_bounds__(64)
env_egosensing_kernel(const float* __restrict__ joints, const float* __restrict__ R0a,
                      const float* __restrict__ T0a, const int32_t* __restrict__ slot_ids,
                      const int32_t* __restrict__ accept, const double* __restrict__ segs, int S,
                      double ray_len, float* __restrict__ ego, const float* __restrict__ holes_all, int n_holes,
                      int frames = 2, int frame0 = 0) {
  const int i = blockIdx.x;
  if (accept && !accept[i]) return;
  const int t = threadIdx.x >> 5, ray = threadIdx.x & 31;
  const int slot = slot_ids ? slot_ids[i] : i;     // output row; R0/T0/joints are indexed by item i
  const float* R0 = R0a + (int64_t)i * 9;
  const float* T0 = T0a + (int64_t)i * 3;
  const float* j = joints + ((int64_t)i * frames + frame0 + t) * NJ * 3;   // the item's frames frame0, frame0 + 1
  float w[4][2];
  const int ids[4] = {23, 24, 56, 57};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float* p = j + ids[q] * 3;
    w[q][0] = (R0[0] * p[0] + R0[1] * p[1] + R0[2] * p[2]) + T0[0];
    w[q][1] = (R0[3] * p[0] + R0[4] * p[1] + R0[5] * p[2]) + T0[1];
  }
  // look_at = j57 - j23 + j56 - j24 (float32), then float64, z := 0, normalised
  double lx = (double)(((w[3][0] - w[0][0]) + w[2][0]) - w[1][0]);
  double ly = (double)(((w[3][1] - w[0][1]) + w[2][1]) - w[1][1]);
  const double ln = sqrt(lx * lx + ly * ly);
  lx /= ln; ly /= ln;
  const double ex = (double)((w[0][0] + w[1][0]) / 2.0f), ey = (double)((w[0][1] + w[1][1]) / 2.0f);
  const double PI_2 = 1.57079632679489661923;
  const double ang = ray == 31 ? PI_2 : -PI_2 + (double)ray * ((PI_2 - (-PI_2)) / 31.0);
  const double ca = cos(ang), sa = sin(ang);
  const double dx = lx * ca - ly * sa, dy = ly * ca + lx * sa;
  // even-odd containment of the eye + nearest boundary crossing along the ray
  int crossings = 0;
  double tmin = ray_len;
  for (int s = 0; s < S; ++s) {
    const double ax = segs[4 * s], ay = segs[4 * s + 1], bx = segs[4 * s + 2], by = segs[4 * s + 3];
    if ((ay > ey) != (by > ey)) {
      const double xi = ax + (ey - ay) * (bx - ax) / (by - ay);
      if (xi > ex) ++crossings;
    }
    const double sx = bx - ax, sy = by - ay;
    const double den = dx * sy - dy * sx;
    if (den != 0.0) {
      const double qx = ax - ex, qy = ay - ey;
      const double tt = (qx * sy - qy * sx) / den;
      const double u = (qx * dy - qy * dx) / den;
      if (tt >= 0.0 && u >= 0.0 && u <= 1.0 && tt < tmin) tmin = tt;
    }
  }
  // crowd dynamics: the other agents' rectangles are holes of the scene polygon (their union's boundary is hit
  // first on an edge of one of them as long as the eye is outside all of them; inside one => off the polygon => 0)
  bool in_hole = false;
  if (holes_all != nullptr) {
    const float* holes = holes_all + (int64_t)slot * n_holes * 4;
    for (int k = 0; k < n_holes; ++k) {
      const double x0 = (double)holes[4 * k], y0 = (double)holes[4 * k + 1], x1 = (double)holes[4 * k + 2], y1 = (double)holes[4 * k + 3];
      if (ex >= x0 && ex <= x1 && ey >= y0 && ey <= y1) in_hole = true;
      const double cxs[5] = {x0, x1, x1, x0, x0}, cys[5] = {y0, y0, y1, y1, y0};
      for (int q = 0; q < 4; ++q) {
        const double ax = cxs[q], ay = cys[q], sx = cxs[q + 1] - ax, sy = cys[q + 1] - ay;
        const double den = dx * sy - dy * sx;
        if (den != 0.0) {
          const double qx = ax - ex, qy = ay - ey;
          const double tt = (qx * sy - qy * sx) / den;
          const double u = (qx * dy - qy * dx) / den;
          if (tt >= 0.0 && u >= 0.0 && u <= 1.0 && tt < tmin) tmin = tt;
        }
      }
    }
  }
  // the hit point is reconstructed like shapely's end coordinate and its distance re-measured
  double d = 0.0;
  if ((crossings & 1) && !in_hole) {
    const double hx = ex + tmin * dx, hy = ey + tmin * dy;
    d = sqrt((hx - ex) * (hx - ex) + (hy - ey) * (hy - ey));
  }
  ego[((int64_t)slot * 2 + t) * 32 + ray] = (float)(-1.0 + 2.0 * (d / ray_len));
}

}  // namespace eg

using namespace eg;

struct EgEnv {
  int device = 0;
  EgEnvConfig cfg;
  EgLbs* lbs = nullptr;
  EgMotion* motion = nullptr;
  EgVposer* vposer = nullptr;
  // scene
  const float* grid = nullptr; int D0 = 0, D1 = 0, D2 = 0;
  const float *center = nullptr, *scale = nullptr;
  const uint8_t* skip = nullptr;
  const double* segs = nullptr; int S = 0;
  const float* tris = nullptr; int n_tris = 0;
  // crowd dynamics (eg_env_set_crowd)
  const float* holes = nullptr; int n_holes = 0; float* bbox_out = nullptr; int pene_terminates = 1;
  // workspace
  int cap = 0;
  float *Y = nullptr, *params = nullptr, *joints = nullptr, *mproj = nullptr, *vp = nullptr, *prest = nullptr;
  float *joints2 = nullptr, *R0c = nullptr, *T0c = nullptr, *seedc = nullptr;
  int32_t* counts = nullptr;
  // VPoser (needs only the blended parameters) runs beside the LBS + SDF pass on a side stream
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int overlap = 1;          // EG_ENV_OVERLAP=0: everything on the caller's stream
  int seed_lbs = 0;         // EG_ENV_SEED_LBS=1: ego-sensing joints from a second SMPL-X pass over the new seed (reference order)
};

static int env_ws(EgEnv* h, int E) {
  if (E <= h->cap) return EG_OK;
  float** fb[] = {&h->Y, &h->params, &h->joints, &h->mproj, &h->vp, &h->prest, &h->joints2, &h->R0c, &h->T0c, &h->seedc};
  for (auto p : fb) { cudaFree(*p); *p = nullptr; }
  cudaFree(h->counts); h->counts = nullptr; h->cap = 0;
  const size_t e = (size_t)E;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Y, e * NT * 201 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->params, e * NT * 93 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->joints, e * NT * NJ * 3 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->mproj, e * NT * NM * 3 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->vp, e * NT * 32 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->prest, e * 3 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->joints2, e * 2 * NJ * 3 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->R0c, e * 9 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->T0c, e * 3 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->seedc, e * 2 * 93 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->counts, e * NT * 4));
  h->cap = E;
  return EG_OK;
}

extern "C" int eg_env_create(const EgEnvConfig* cfg, EgLbs* lbs, EgMotion* motion, EgVposer* vposer,
                             int device, EgEnv** out) {
  EG_REQUIRE(cfg && lbs && motion && vposer && out, "null pointer");
  EG_REQUIRE(cfg->max_depth > 0 && cfg->ray_len > 0, "bad config");
  for (int k = 0; k < 6; ++k) EG_REQUIRE(cfg->feet_marker_idx[k] >= 0 && cfg->feet_marker_idx[k] < NM, "feet marker index");
  EgEnv* h = new EgEnv();
  h->device = device; h->cfg = *cfg; h->lbs = lbs; h->motion = motion; h->vposer = vposer;
  *out = h;
  EG_CUDA_CHECK(cudaSetDevice(device));
  EG_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  EG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  EG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  const char* e1 = getenv("EG_ENV_OVERLAP");
  h->overlap = (e1 != nullptr && e1[0] == '0') ? 0 : 1;
  const char* e2 = getenv("EG_ENV_SEED_LBS");
  h->seed_lbs = (e2 != nullptr && e2[0] == '1') ? 1 : 0;
  return EG_OK;
}

extern "C" void eg_env_destroy(EgEnv* h) {
  if (!h) return;
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaSetDevice(h->device);
  float* fb[] = {h->Y, h->params, h->joints, h->mproj, h->vp, h->prest, h->joints2, h->R0c, h->T0c, h->seedc};
  for (auto p : fb) cudaFree(p);
  cudaFree(h->counts);
  delete h;
}

extern "C" int eg_env_set_config(EgEnv* h, const EgEnvConfig* cfg) {
  EG_REQUIRE(h && cfg, "null pointer");
  h->cfg = *cfg;
  return EG_OK;
}

extern "C" int eg_env_set_scene(EgEnv* h, const float* grid, int D0, int D1, int D2, const float* center_dev,
                                const float* scale_dev, const uint8_t* skip_mask, const double* segments_dev,
                                int n_segments) {
  EG_REQUIRE(h && grid && center_dev && scale_dev && segments_dev, "null pointer");
  EG_REQUIRE(D0 > 0 && D1 > 0 && D2 > 0 && n_segments > 0, "bad sizes");
  h->grid = grid; h->D0 = D0; h->D1 = D1; h->D2 = D2; h->center = center_dev; h->scale = scale_dev;
  h->skip = skip_mask; h->segs = segments_dev; h->S = n_segments;
  return eg_sdf_prepare(grid, D0, D1, D2, nullptr);   // conservative coarse grid for the fused sign query
}

extern "C" int eg_env_set_navmesh(EgEnv* h, const float* tris_dev, int n_tris) {
  EG_REQUIRE(h && tris_dev && n_tris > 0, "bad arguments");
  h->tris = tris_dev; h->n_tris = n_tris;
  return EG_OK;
}

extern "C" int eg_env_set_crowd(EgEnv* h, const float* holes_dev, int n_holes, float* bbox_out_dev,
                                int penetration_terminates) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(n_holes >= 0 && (holes_dev != nullptr || n_holes == 0), "holes pointer / count mismatch");
  h->holes = n_holes > 0 ? holes_dev : nullptr; h->n_holes = n_holes; h->bbox_out = bbox_out_dev;
  h->pene_terminates = penetration_terminates ? 1 : 0;
  return EG_OK;
}

#define EG_TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

extern "C" int eg_env_step(EgEnv* h, const EgEnvBuffers* b, const float* z, int E, void* stream) {
  EG_REQUIRE(h && b && z, "null pointer");
  EG_REQUIRE(h->grid != nullptr, "scene not set (eg_env_set_scene)");
  EG_REQUIRE(b->state && b->seed && b->R0 && b->T0 && b->betas && b->dist && b->steps && b->goal && b->ego &&
             b->obs_dist && b->obs_time && b->reward && b->terminated, "null buffer");
  if (E <= 0) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(env_ws(h, E));
  cudaStream_t st = as_stream(stream);
  stage_mark(st, 0);
  // b: C-VAE rollout + body regressor (:109)
  EG_TRY(eg_motion_sample_prior(h->motion, b->state, 2 * 402, 402, z, b->betas, E, h->Y, h->params, stream));
  stage_mark(st, 1);
  // c: history frames + parameter blending (:116-120)
  EG_LAUNCH(env_prepare_params_kernel, E, 128, 0, st, b->seed, h->params, E);
  stage_mark(st, 2);
  // f: VPoser latent of every frame's body pose (:197-200) - independent of the body pass: side stream
  void* vp_stream = stream;
  if (h->overlap) {
    EG_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
    EG_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    vp_stream = h->side;
  }
  EG_TRY(eg_vposer_encode(h->vposer, h->params + 6, 93, E * NT, h->vp, vp_stream));
  if (h->overlap) EG_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
  if (h->cfg.pene_mode == 1) {
    // box-scene env (crowd_env_2f_box.py): no vertex-level SDF query - joints and markers only
    EG_REQUIRE(h->tris != nullptr, "navmesh not set (eg_env_set_navmesh)");
    EG_TRY(eg_lbs_forward(h->lbs, h->params, b->betas, E, E * NT, nullptr, h->joints, h->mproj, stream));
  } else {
    // d,e: SMPL-X on 20 bodies per env fused with the SDF penetration query (:133-177)
    EG_TRY(eg_lbs_forward_sdf(h->lbs, h->params, b->betas, E, E * NT, NT, b->R0, b->T0, h->grid, h->D0, h->D1,
                              h->D2, h->center, h->scale, h->skip, h->counts, h->joints, h->mproj, stream));
  }
  stage_mark(st, 3);
  if (h->overlap) EG_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
  stage_mark(st, 4);
  // i: ego-sensing of the new 2-frame seed (:290-296). The seed is frames 18 / 19 of this primitive, and re-canonicalising
  // moves the FRAME, not the body: R0' j' + T0' = R0 j + T0. The world joints the rays start from are therefore the joints
  // of frames 18 / 19 the body pass above already produced, under the frame of THIS step (before the reward kernel updates
  // R0 / T0) - no second SMPL-X pass. (The reference re-evaluates SMPL-X on the updated parameters; same points up to fp32
  // rounding of its axis-angle round trip. EG_ENV_SEED_LBS=1 keeps that order.)
  if (!h->seed_lbs)
    EG_LAUNCH(env_egosensing_kernel, E, 64, 0, st, h->joints, b->R0, b->T0, nullptr, nullptr, h->segs, h->S,
              (double)h->cfg.ray_len, b->ego, h->holes, h->n_holes, NT, NT - 2);
  EG_TRY(eg_lbs_rest_pelvis(h->lbs, b->betas, E, E, h->prest, stream));
  StepArgs a{h->cfg, *b, h->Y, h->params, h->counts, h->joints, h->mproj, h->vp, h->prest, h->tris, h->n_tris,
             h->holes, h->n_holes, h->bbox_out, h->pene_terminates};
  EG_LAUNCH(env_reward_recanon_kernel, E, 256, 0, st, a);
  stage_mark(st, 5);
  if (h->seed_lbs) {
    EG_TRY(eg_lbs_forward(h->lbs, b->seed, b->betas, E, E * 2, nullptr, h->joints2, nullptr, stream));
    stage_mark(st, 6);
    EG_LAUNCH(env_egosensing_kernel, E, 64, 0, st, h->joints2, b->R0, b->T0, nullptr, nullptr, h->segs, h->S,
              (double)h->cfg.ray_len, b->ego, h->holes, h->n_holes, 2, 0);
  }
  stage_mark(st, 7);
  return EG_OK;
}

static int env_reset_impl(EgEnv* h, const EgEnvBuffers* b, const int32_t* env_ids, const uint8_t* mask, int n,
                          const float* world_params, const float* goals, const float* betas_cand,
                          int32_t* accept, void* stream);

extern "C" int eg_env_reset(EgEnv* h, const EgEnvBuffers* b, const int32_t* env_ids, int n,
                            const float* world_params, const float* goals, const float* betas_cand,
                            int32_t* accept, void* stream) {
  EG_REQUIRE(env_ids != nullptr, "null pointer");
  return env_reset_impl(h, b, env_ids, nullptr, n, world_params, goals, betas_cand, accept, stream);
}

// standalone SMPLXParser.update_transl_glorot (baseops.py:537-598): one thread per body; rows 6..92 pass through
__global__ void __launch_bounds__(128)
update_transl_glorot_kernel(const float* __restrict__ R, const float* __restrict__ T, const float* __restrict__ delta,
                            const float* xb, int N, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float o[6];
  update_transl_glorot(R + (int64_t)i * 9, T + (int64_t)i * 3, delta + (int64_t)i * 3, xb + (int64_t)i * 93, o);
  float* y = out + (int64_t)i * 93;
  if (out != xb)
    for (int k = 6; k < 93; ++k) y[k] = xb[(int64_t)i * 93 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) y[k] = o[k];
}

extern "C" int eg_update_transl_glorot(EgLbs* lbs, const float* transf_rotmat, const float* transf_transl, const float* betas,
                                       int betas_rows, const float* xb, int N, float* delta_T, float* xb_out, void* stream) {
  EG_REQUIRE(lbs && transf_rotmat && transf_transl && betas && xb && delta_T && xb_out && N >= 0, "bad arguments");
  if (N == 0) return EG_OK;
  int rc = eg_lbs_rest_pelvis(lbs, betas, betas_rows, N, delta_T, stream);
  if (rc) return rc;
  EG_LAUNCH(update_transl_glorot_kernel, (N + 127) / 128, 128, 0, as_stream(stream), transf_rotmat, transf_transl,
            (const float*)delta_T, xb, N, xb_out);
  return EG_OK;
}

// ego-sensing alone (CrowdEnv._calc_egosensing, crowd_env_2f.py:524-613): the operator the step / reset paths launch,
// exposed so it can be checked on GIVEN joints
extern "C" int eg_egosensing(const float* joints_local, const float* R0, const float* T0, int n, const double* segments_dev,
                             int n_segments, double ray_len, const float* holes_dev, int n_holes, float* ego_out, void* stream) {
  EG_REQUIRE(joints_local && R0 && T0 && segments_dev && ego_out && n >= 0 && n_segments >= 0, "bad arguments");
  EG_REQUIRE(holes_dev != nullptr || n_holes == 0, "holes pointer missing");
  if (n == 0) return EG_OK;
  EG_LAUNCH(env_egosensing_kernel, n, 64, 0, as_stream(stream), joints_local, R0, T0, nullptr, nullptr, segments_dev, n_segments,
            ray_len, ego_out, holes_dev, n_holes);
  return EG_OK;
}

// ---- episode restart from a pool of pre-computed initial states (no host synchronisation) ------------------------
// block e: if mask[e], rank(e) = number of set entries below e, pool row = (cursor + rank) mod pool_rows, copy every
// state field of that row into slot e. The LAST block to finish (ticket counter) advances the cursor by the number of set
// entries - every block read the cursor before it took its ticket, so all of them used the old value.
struct PoolCursor { long long cursor; unsigned ticket; unsigned pad; };
__global__ void __launch_bounds__(128)
env_restart_pool_kernel(const EgEnvBuffers dst, const EgEnvBuffers pool, int pool_rows, const uint8_t* __restrict__ mask, int E,
                        PoolCursor* pc) {
  __shared__ int red[4];
  __shared__ int is_last;
  const int e = blockIdx.x, tid = threadIdx.x;
  const long long cur = *reinterpret_cast<volatile long long*>(&pc->cursor);
  auto count_below = [&](int n) {
    int c = 0;
    for (int i = tid; i < n; i += 128) c += mask[i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) red[tid >> 5] = c;
    __syncthreads();
    const int tot = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return tot;
  };
  if (mask[e] != 0) {
    const int rank = count_below(e);
    const int64_t r = (int64_t)((cur + rank) % pool_rows);
    auto cp = [&](float* d, const float* s_, int n) {
      for (int i = tid; i < n; i += 128) d[(int64_t)e * n + i] = s_[r * n + i];
    };
    cp(dst.state, pool.state, 2 * 402); cp(dst.seed, pool.seed, 2 * 93); cp(dst.R0, pool.R0, 9); cp(dst.T0, pool.T0, 3);
    cp(dst.betas, pool.betas, 10); cp(dst.goal, pool.goal, 3); cp(dst.ego, pool.ego, 64);
    if (tid == 0) {
      dst.dist[e] = pool.dist[r]; dst.steps[e] = pool.steps[r];
      dst.obs_dist[e] = pool.obs_dist[r]; dst.obs_time[e] = pool.obs_time[r];
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&pc->ticket, 1u) == (unsigned)(E - 1);
  }
  __syncthreads();
  if (is_last) {
    const int tot = count_below(E);
    if (tid == 0) { pc->cursor = cur + tot; pc->ticket = 0; __threadfence(); }
  }
}

extern "C" int eg_env_restart_from_pool(const EgEnvBuffers* dst, const EgEnvBuffers* pool, int pool_rows, const uint8_t* mask,
                                        int E, void* cursor_dev, void* stream) {
  EG_REQUIRE(dst && pool && mask && cursor_dev, "null pointer");
  EG_REQUIRE(pool_rows > 0 && E >= 0, "bad sizes");
  EG_REQUIRE(dst->state && dst->seed && dst->R0 && dst->T0 && dst->betas && dst->dist && dst->steps && dst->goal && dst->ego &&
             dst->obs_dist && dst->obs_time, "destination buffers incomplete");
  EG_REQUIRE(pool->state && pool->seed && pool->R0 && pool->T0 && pool->betas && pool->dist && pool->steps && pool->goal &&
             pool->ego && pool->obs_dist && pool->obs_time, "pool buffers incomplete");
  if (E == 0) return EG_OK;
  EG_LAUNCH(env_restart_pool_kernel, E, 128, 0, as_stream(stream), *dst, *pool, pool_rows, mask, E,
            reinterpret_cast<PoolCursor*>(cursor_dev));
  return EG_OK;
}

extern "C" int eg_env_reset_masked(EgEnv* h, const EgEnvBuffers* b, const uint8_t* mask, int E,
                                   const float* world_params, const float* goals, const float* betas_cand,
                                   int32_t* accept, void* stream) {
  EG_REQUIRE(mask != nullptr, "null pointer");
  return env_reset_impl(h, b, nullptr, mask, E, world_params, goals, betas_cand, accept, stream);
}

static int env_reset_impl(EgEnv* h, const EgEnvBuffers* b, const int32_t* env_ids, const uint8_t* mask, int n,
                          const float* world_params, const float* goals, const float* betas_cand,
                          int32_t* accept, void* stream) {
  EG_REQUIRE(h && b && world_params && goals && betas_cand && accept, "null pointer");
  EG_REQUIRE(h->grid != nullptr, "scene not set (eg_env_set_scene)");
  if (n <= 0) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(env_ws(h, n));
  cudaStream_t st = as_stream(stream);
  EG_TRY(eg_lbs_rest_pelvis(h->lbs, betas_cand, n, n, h->prest, stream));
  EG_TRY(eg_lbs_forward(h->lbs, world_params, betas_cand, n, n * 2, nullptr, h->joints, nullptr, stream));
  EG_LAUNCH(env_reset_canon_kernel, n, 32, 0, st, world_params, h->joints, h->prest, n, h->R0c, h->T0c, h->seedc);
  if (h->cfg.pene_mode == 1) {
    EG_REQUIRE(h->tris != nullptr, "navmesh not set (eg_env_set_navmesh)");
    EG_TRY(eg_lbs_forward(h->lbs, h->seedc, betas_cand, n, n * 2, nullptr, h->joints2, h->mproj, stream));
  } else {
    EG_TRY(eg_lbs_forward_sdf(h->lbs, h->seedc, betas_cand, n, n * 2, 2, h->R0c, h->T0c, h->grid, h->D0, h->D1, h->D2,
                              h->center, h->scale, h->skip, h->counts, h->joints2, h->mproj, stream));
  }
  EG_LAUNCH(env_reset_commit_kernel, n, 256, 0, st, *b, env_ids, n, h->counts, h->R0c, h->T0c, h->seedc,
            h->joints2, h->mproj, goals, betas_cand, accept, h->cfg, h->tris, h->n_tris, h->pene_terminates, h->bbox_out, mask);
  EG_LAUNCH(env_egosensing_kernel, n, 64, 0, st, h->joints2, h->R0c, h->T0c, env_ids, accept, h->segs, h->S,
            (double)h->cfg.ray_len, b->ego, h->holes, h->n_holes);
  return EG_OK;
}
