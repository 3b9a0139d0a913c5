/*
 * egogen_b200 - C ABI of the B200-native crowd_ppo hot path.
 *
 * The reference (ligengen/EgoGen @ ce2c590) is pure Python and has no FFI layer; its "operator
 * API" for this path is a set of Python call surfaces. Each entry point
 * below replaces the arithmetic behind one of those surfaces and cites it. The Python host side
 * (egogen_b200/*.py) mirrors the reference's class / function names and calls these through
 * ctypes; see INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every `const float*` / `float*` / `int32_t*` argument is a DEVICE pointer
 *     to contiguous row-major memory owned by the caller unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *     asynchronous w.r.t. the host and ordered on that stream;
 *   - return 0 on success, a negative EG_ERR_* otherwise; eg_last_error() gives the message
 *     (thread-local);
 *   - handles are bound to one device and are not thread-safe; no allocation happens inside a
 *     hot call unless a batch larger than any seen before forces the workspace to grow.
 */
#ifndef EGOGEN_B200_H
#define EGOGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EG_OK 0
#define EG_ERR_INVALID_ARG (-1)
#define EG_ERR_CUDA (-2)
#define EG_ERR_ALLOC (-3)
#define EG_ERR_STATE (-4)

#define EG_SMPLX_V 10475
#define EG_SMPLX_J 55
#define EG_SMPLX_JOINTS_OUT 127
#define EG_XB_DIM 93

int eg_version(void);
const char* eg_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches counter) */
int64_t eg_launch_count(void);
/* device-side timing of the dominant kernel (fused LBS vertex kernel): CUDA events on the launching stream.
 * eg_profile_enable(1) starts collecting; eg_profile_read synchronises, returns the summed kernel time and the
 * number of timed launches, and clears the counters. */
int eg_profile_enable(int on);
int eg_profile_read(double* total_ms, int64_t* launches, int64_t* units /* bodies processed */);
/* per-stage device time of eg_env_step (ids: 1 C-VAE decode + regressor, 2 param blend, 3 fused LBS+SDF,
 * 4 VPoser, 5 rewards + re-canonicalisation, 6 seed-joint LBS, 7 ego-sensing); ms_out[n >= 8] */
int eg_stage_profile_enable(int on);
int eg_stage_profile_read(double* ms_out, int n);

/* ------------------------------------------------------------------------------------------
 * calc_sdf  - replaces motion/crowd_ppo/utils.py:54-84 (F.grid_sample 5-D trilinear,
 * align_corners=False, padding_mode='border', negated). grid is sdf[D0,D1,D2] with vertex
 * x -> axis 0 (slowest). pts [P,3] world space. val [P] (negative = penetration).
 * base_idx (nullable) [P,3] int32 = floor of the clamped un-normalised index per axis, the
 * bit-exact parity target. center_dev [3], scale_dev [1] are device pointers (the reference keeps
 * them as CUDA tensors, main_ppo.py:302-304).
 * ------------------------------------------------------------------------------------------ */
int eg_sdf_sample(const float* grid, int D0, int D1, int D2, const float* center_dev,
                  const float* scale_dev, const float* pts, int64_t P, float* val,
                  int32_t* base_idx, void* stream);

/* Optional: register the sign structures of `grid` (kept inside the library, keyed by the pointer): a conservative
 * 8^3-cell coarse grid with its sign bits, 2^3-cell sign bits, and an exact 2-bit class per fine cell (every corner > 0 /
 * no corner > 0 / both), D0*D1*D2/4 bytes. The fused penetration count (eg_lbs_forward_sdf / eg_env_step) then answers
 * "sdf < 0 ?" without the 8-corner sample wherever the class decides it; results are identical to the full sample.
 * Call again after mutating the grid in place; eg_sdf_release drops the entry. Synchronises `stream`. */
int eg_sdf_prepare(const float* grid, int D0, int D1, int D2, void* stream);
int eg_sdf_release(const float* grid);

/* crowd_env_2f.py:170-176: per-body count of vertices with sdf < 0, skipping vertices whose
 * skip_mask byte is non-zero (feet). sdf_vals [N,V]; counts int32 [N]. */
int eg_penetration_count(const float* sdf_vals, int N, int V, const uint8_t* skip_mask,
                         int32_t* counts, void* stream);

/* Config-5 ego-depth sweep (no reference implementation): sphere-trace
 * H*W pinhole rays per camera through the same SDF grid / calc_sdf sampling. cam [A,12] =
 * (eye xyz, right xyz, up xyz, forward xyz); depth [A,H,W] metres (max_range where nothing hit);
 * steps_out (nullable) int32 [A,H,W] = sphere-trace iterations used. */
int eg_ego_depth(const float* grid, int D0, int D1, int D2, const float* center_dev,
                 const float* scale_dev, const float* cam, int A, int H, int W, float fx, float fy,
                 float max_range, int max_steps, float hit_eps, float* depth, int32_t* steps_out,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * SMPL-X linear blend skinning - replaces smplx.SMPLX.forward / smplx.lbs.lbs behind
 * SMPLXParser.forward_smplx (motion/models/baseops.py:338-398) and its selectors (:401-463).
 * ------------------------------------------------------------------------------------------ */
typedef struct EgLbs EgLbs;

typedef struct EgLbsModel {          /* all HOST pointers, float32 / int32, row-major */
  int32_t n_verts;                   /* 10475 */
  int32_t n_joints;                  /* 55 */
  int32_t n_shape;                   /* 20 = 10 betas + 10 expression */
  int32_t n_pose_basis;              /* 486 */
  int32_t n_faces;
  int32_t n_hand_pca;                /* 12 */
  int32_t n_extra;                   /* 21 vertex joints */
  int32_t n_landmarks;               /* 51 */
  const float* v_template;           /* [V,3] */
  const float* shapedirs;            /* [V,3,n_shape] */
  const float* posedirs;             /* [n_pose_basis, V*3] */
  const float* J_regressor;          /* [J,V] */
  const int32_t* parents;            /* [J], parents[0] = -1 */
  const float* lbs_weights;          /* [V,J] */
  const float* hand_comp_l;          /* [n_hand_pca,45] */
  const float* hand_comp_r;          /* [n_hand_pca,45] */
  const float* pose_mean;            /* [J*3] */
  const int32_t* extra_vids;         /* [n_extra] */
  const int32_t* faces;              /* [n_faces,3] */
  const int32_t* lmk_faces_idx;      /* [n_landmarks] */
  const float* lmk_bary;             /* [n_landmarks,3] */
} EgLbsModel;

int eg_lbs_create(const EgLbsModel* model_host, int device, EgLbs** out);
void eg_lbs_destroy(EgLbs* h);
/* marker placement (baseops.py:328-335): vertex ids whose positions eg_lbs_forward returns in
 * `markers`. Rebuilds the compact vertex set (markers + vertex joints + landmark corners). */
int eg_lbs_set_markers(EgLbs* h, const int32_t* marker_vids_host, int n_markers);
int eg_lbs_max_skin_nnz(const EgLbs* h);
/* full-mesh mainloop: 1 (default) = tcgen05/TMEM/TMA TF32 tiles, 0 = fp32 SIMT tiles (debug / comparison) */
int eg_lbs_set_mainloop(EgLbs* h, int use_tcgen05);
/* backward of the marker output of eg_lbs_forward w.r.t. the body parameters: d_xb [N,93] = (d markers / d xb)^T
 * d_markers [N,n_markers,3] (transl, global_orient, body_pose, hand PCA; betas are data). This is the SMPL-X gradient the
 * reference's regressor / combo training loss needs (bm(...).vertices[:, markers] inside calc_loss,
 * motion/models/models_GAMMA_primitive.py:616-631). */
int eg_lbs_markers_backward(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                            const float* d_markers, float* d_xb, void* stream);
/* same, and also d_rot [N,22,9] = dL/dR_j (row-major) for the global + 21 body joints, for callers whose pose comes from
 * another rotation parameterisation (the regressor's 6-D representation, baseops.py:120-130) */
int eg_lbs_markers_backward_rot(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                                const float* d_markers, float* d_xb, float* d_rot, void* stream);
/* SMPLXParser.calc_calibrate_offset (baseops.py:494-534): pelvis of the zero-transl / zero-orient body,
 * i.e. the rest position of the root joint J_0(betas). out [N,3]. */
int eg_lbs_rest_pelvis(EgLbs* h, const float* betas, int betas_rows, int N, float* out, void* stream);

/* xb [N,93] = [transl3, global_orient3, body_pose63, lhand_pca12, rhand_pca12] (baseops.py:366-374),
 * betas [betas_rows,10], betas_rows divides N and body n uses row n / (N / betas_rows); expression / jaw / eye poses are zero as in the
 * reference. Outputs (each nullable): verts [N,V,3], joints [N,127,3], markers [N,n_markers,3]. */
int eg_lbs_forward(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                   float* verts, float* joints, float* markers, void* stream);

/* Fused crowd_env_2f.py:133-177 for N = E*frames bodies (env-major): LBS, world transform with the
 * env's R0 [E,3,3] / T0 [E,3], calc_sdf, feet skip, per-body penetration count - vertices are never
 * written to HBM. counts int32 [N]; joints / markers as above (body-local frame, not world). */
int eg_lbs_forward_sdf(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                       int frames_per_env, const float* R0, const float* T0, const float* grid,
                       int D0, int D1, int D2, const float* center_dev, const float* scale_dev,
                       const uint8_t* skip_mask, int32_t* counts, float* joints, float* markers,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Motion model - replaces GAMMAPrimitiveCombo.sample_prior
 * (motion/models/models_GAMMA_primitive.py:334-360; predictor decode :83-101, regressor :222-301,
 * 6-D -> axis-angle tail :208-219). Weights stay in the caller's tensors: `weights_host` is a HOST
 * array of DEVICE pointers in this state_dict order:
 *   predictor.x_enc.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0},
 *   predictor.drnn_mlp.layers.{0,1,2}.{weight,bias}, predictor.d_rnn.{weight_ih,weight_hh,bias_ih,bias_hh},
 *   predictor.d_mlp.layers.{0,1}.{weight,bias}, predictor.d_out.{weight,bias},
 *   regressor.pnet.in_fc.{weight,bias}, regressor.pnet.layers.{0..n-1}.layers.{0,1}.{weight,bias},
 *   regressor.pnet.out_fc.{weight,bias}
 * ------------------------------------------------------------------------------------------ */
typedef struct EgMotion EgMotion;
typedef struct EgMotionDims {
  int32_t in_dim;      /* 201 marker coords */
  int32_t h_dim;       /* 256 */
  int32_t z_dim;       /* 128 */
  int32_t mlp_dim;     /* 512 */
  int32_t reg_h;       /* 128 */
  int32_t reg_blocks;  /* 10 */
  int32_t reg_recur;   /* 3 */
  int32_t body_dim;    /* 159 = 3 + 22*6 + 24 */
} EgMotionDims;

int eg_motion_create(const EgMotionDims* dims, const void* const* weights_host, int n_weights,
                     int device, EgMotion** out);
void eg_motion_destroy(EgMotion* h);
/* the fused decode / regressor kernels read transposed copies of the weights: call after changing weights in place */
int eg_motion_refresh(EgMotion* h, void* stream);
/* 1 (default): fused row-tile kernels for the 18-step decode loop and the 3x22-layer regressor; 0: layer-by-layer */
int eg_motion_set_fused(EgMotion* h, int fused);
/* X: marker history, frame t of env b at X + b*ldx_env + t*ldx_frame (201 floats); z [B,128];
 * betas [B,10]. Y [B,20,201] receives the 2 history frames + 18 predicted frames;
 * Yb [B,20,93]: frames 2..19 are written (axis-angle body params), frames 0..1 are left untouched. */
int eg_motion_sample_prior(EgMotion* h, const float* X, int ldx_env, int ldx_frame, const float* z,
                           const float* betas, int B, float* Y, float* Yb, void* stream);

/* VPoser v1.0 encoder, eval mode, `.loc` only (reference call site crowd_env_2f.py:197-200).
 * weights_host order: bodyprior_enc_bn1.{weight,bias,running_mean,running_var}, bodyprior_enc_fc1.{weight,bias},
 * bodyprior_enc_bn2.{weight,bias,running_mean,running_var}, bodyprior_enc_fc2.{weight,bias},
 * bodyprior_enc_mu.{weight,bias}.  x: row m at x + m*ldx (63 floats); loc [M,32]. */
typedef struct EgVposer EgVposer;
int eg_vposer_create(const void* const* weights_host, int n_weights, int device, EgVposer** out);
void eg_vposer_destroy(EgVposer* h);
int eg_vposer_encode(EgVposer* h, const float* x, int ldx, int M, float* loc, void* stream);

/* ------------------------------------------------------------------------------------------
 * Vectorised environment - replaces CrowdEnv.step / CrowdEnv.reset
 * (motion/crowd_ppo/crowd_env_2f.py:78-317, 320-415) for E environments per call with the
 * reference's 4x batch duplication removed. All state lives in caller-owned device buffers.
 * ------------------------------------------------------------------------------------------ */
typedef struct EgEnv EgEnv;
typedef struct EgEnvConfig {
  int32_t max_depth;             /* 13 (MPVAEPolicy_samp_collision.yaml:80) */
  int32_t finetuning;            /* crowd_env_2f.py:268-271,299-302 */
  int32_t pene_terminate_count;  /* 40 (:176) */
  int32_t feet_marker_idx[6];    /* main_ppo.py:298-299 */
  float reproj_factor;           /* 0.5 */
  float goal_thresh;             /* 0.1 */
  float w_skate, w_floor, w_face, w_look, w_success, w_dist, w_vp, w_pene;
  float ray_len;                 /* 7 */
  /* penetration model: 0 = SDF vertex count of crowd_env_2f.py:162-177 (Replica room0 env);
   * 1 = 2-D walkability map of crowd_env_2f_box.py:279-295 (random_box_obstacle_new env): cells of the local
   * map_res x map_res grid (get_map, batch_gen_amass.py:934-968) that lie inside the markers' xy bounding box and
   * outside every navmesh triangle are counted; count > pene_thres => r_pene = 0 and the episode terminates */
  int32_t pene_mode;
  int32_t map_res;               /* 16 */
  float map_extent;              /* 0.8 */
  float pene_thres;              /* 3 */
} EgEnvConfig;

typedef struct EgEnvBuffers {    /* device pointers; (n) = nullable */
  float* state;          /* [E,2,402] */
  float* seed;           /* [E,2,93]  body_param_seed */
  float* R0;             /* [E,3,3] */
  float* T0;             /* [E,3] */
  float* betas;          /* [E,10] */
  float* dist;           /* [E] distance to goal after the previous step */
  int32_t* steps;        /* [E] */
  float* goal;           /* [E,3] world-space wpath[-1] */
  float* ego;            /* [E,2,32] */
  float* obs_dist;       /* [E] 1/(dist+1) */
  float* obs_time;       /* [E] 1 - steps/max_depth */
  float* reward;         /* [E] */
  uint8_t* terminated;   /* [E] */
  uint8_t* goal_reached; /* [E] (n) */
  float* reward_terms;   /* [E,8] (n): skate, floor, face, look, goal, dist, pene, vp */
  float* out_markers;    /* [E,20,67,3] (n) blended markers of this primitive (rollout writer) */
  float* out_params;     /* [E,20,93]   (n) */
  float* out_pelvis;     /* [E,20,3]    (n) */
} EgEnvBuffers;

int eg_env_create(const EgEnvConfig* cfg, EgLbs* lbs, EgMotion* motion, EgVposer* vposer, int device,
                  EgEnv** out);
void eg_env_destroy(EgEnv* h);
int eg_env_set_config(EgEnv* h, const EgEnvConfig* cfg);
/* scene: SDF grid as in eg_sdf_sample, feet skip mask [V], 2-D polygon boundary segments
 * (x0,y0,x1,y1) float64 [n_segments,4] (exterior ring + holes of the reference's shapely polygon) */
int eg_env_set_scene(EgEnv* h, const float* grid, int D0, int D1, int D2, const float* center_dev,
                     const float* scale_dev, const uint8_t* skip_mask, const double* segments_dev,
                     int n_segments);
/* navmesh triangles (xy) for pene_mode 1: tris_dev float [n_tris,3,2] (navmesh.vertices[faces, :2]) */
int eg_env_set_navmesh(EgEnv* h, const float* tris_dev, int n_tris);
/* crowd dynamics - replaces DummyCrowdVectorEnv.update_holes_for_each_agent (motion/crowd_ppo/dummy_vector_env.py:33-39)
 * and the bbox / hole handling of crowd_env_crowd_eval.CrowdEnv (crowd_env_crowd_eval.py:66-75,345-352,796-822):
 * holes_dev float [E,n_holes,4] = (xmin,ymin,xmax,ymax) of the OTHER agents of each env's scene (read by the next
 * eg_env_step / eg_env_reset: cut out of the walkability map and hit by the ego rays), bbox_out_dev float [E,4]
 * (nullable) receives each env's own marker bounding box after the step / reset; penetration_terminates = 0 selects
 * the crowd-eval termination (goal or max_depth only, :367) and the reset without start-pose rejection (:391-405).
 * Both pointers are indexed by the env index of the CALL, so a caller stepping a contiguous slice of envs passes
 * offset pointers. n_holes = 0 / NULL restores the single-agent behaviour. */
int eg_env_set_crowd(EgEnv* h, const float* holes_dev, int n_holes, float* bbox_out_dev, int penetration_terminates);
/* one transition of all E envs with actions z [E,128] */
int eg_env_step(EgEnv* h, const EgEnvBuffers* b, const float* z, int E, void* stream);
/* try to (re)start the envs env_ids [n] from sampled world-frame 2-frame seeds world_params [n,2,93],
 * goals [n,3], betas_cand [n,10]; accept [n] = 1 where the start pose is SDF-clean (:379-380) and the
 * slot was initialised, 0 where the caller must resample. */
int eg_env_reset(EgEnv* h, const EgEnvBuffers* b, const int32_t* env_ids, int n, const float* world_params,
                 const float* goals, const float* betas_cand, int32_t* accept, void* stream);
/* same for ALL E slots at once, committed only where mask[e] != 0 (device uint8 [E], typically the `terminated` flags of
 * the step that just ran): candidate e is for slot e. Lets a collector restart finished episodes (tianshou
 * Collector.collect: finished envs are reset immediately, crowd_env_2f.py:320) without reading the mask back to the
 * host. accept [E] = 1 where a slot was restarted. */
int eg_env_reset_masked(EgEnv* h, const EgEnvBuffers* b, const uint8_t* mask, int E, const float* world_params,
                        const float* goals, const float* betas_cand, int32_t* accept, void* stream);
/* CrowdEnv._calc_egosensing (crowd_env_2f.py:524-613) on given joints: joints_local [n,2,127,3] body-frame SMPL-X joints of
 * the 2-frame seed, R0 [n,3,3] / T0 [n,3] the canonical frame; 2 x 32 rays in float64 against the polygon boundary
 * segments (and the optional per-item hole rectangles [n,n_holes,4]); ego_out [n,2,32] in [-1,1]. */
int eg_egosensing(const float* joints_local, const float* R0, const float* T0, int n, const double* segments_dev,
                  int n_segments, double ray_len, const float* holes_dev, int n_holes, float* ego_out, void* stream);
/* Restart from a POOL of pre-computed initial states (rows of `pool`, the state fields of EgEnvBuffers as written by
 * eg_env_reset for accepted candidates): slot e with mask[e] != 0 takes pool row (cursor + rank(e)) mod pool_rows, rank(e) =
 * number of set mask entries below e, and the device cursor advances by the number of set entries - only finished
 * episodes consume candidates, and nothing is read back to the host. cursor_dev: 16 zero-initialised device bytes
 * (int64 cursor, uint32 ticket, pad) owned by the caller. Same reference behaviour as above (crowd_env_2f.py:320). */
int eg_env_restart_from_pool(const EgEnvBuffers* dst, const EgEnvBuffers* pool, int pool_rows, const uint8_t* mask,
                             int E, void* cursor_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * PPO policy - replaces GAMMAPolicyBase / GAMMAActor / GAMMACritic forward
 * (motion/models/models_policy_ppo.py:287-350), GAMMAPPOPolicy.forward / _compute_returns / learn
 * (motion/crowd_ppo/ppo_policy.py:105-265), tianshou's GAE and torch's clip_grad_norm_ + AdamW
 * (main_ppo.py:134). Parameters / gradients / Adam moments are four caller-owned flat fp32 device
 * buffers laid out in ActorCritic(actor, critic, shared_net).parameters() order:
 *   actor.pnet.layers.{k}.layers.{0,1}.{weight,bias}, actor.pnet.out_fc.{weight,bias},
 *   critic.vnet.(same), shared_net.x_enc.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0},
 *   shared_net.ego_enc.(same).
 * ------------------------------------------------------------------------------------------ */
typedef struct EgPolicy EgPolicy;
typedef struct EgPolicyDims {
  int32_t in_dim;    /* 402 */
  int32_t ego_dim;   /* 32 */
  int32_t h_dim;     /* 512 */
  int32_t pe_L;      /* 32 frequencies per scalar */
  int32_t n_blocks;  /* 2 */
  int32_t z_dim;     /* 128 */
} EgPolicyDims;

int64_t eg_policy_param_count(const EgPolicyDims* dims, int64_t* n_actor_critic);
/* offset (in floats) of every parameter tensor inside the flat buffers, in ActorCritic(actor, critic,
 * shared_net).parameters() order; every tensor starts on a 16-byte boundary (padding elements are zero parameters).
 * Returns the number of tensors written, -1 when n_max is too small. */
int eg_policy_param_offsets(const EgPolicyDims* dims, int64_t* offsets, int n_max);
int eg_policy_create(const EgPolicyDims* dims, float* params_flat, float* grads_flat, int device, EgPolicy** out);
void eg_policy_destroy(EgPolicy* h);
/* obs batch: state [B,2,402], ego [B,2,32], dist [B], time [B]. out_actor [B,256] = raw [mu | logvar]
 * (nullable), value [B] (nullable), hx_out [B,1152] (nullable). */
int eg_policy_forward(EgPolicy* h, const float* state, const float* ego, const float* dist, const float* time,
                      int B, int want_actor, int want_critic, float* out_actor, float* value, float* hx_out,
                      void* stream);
/* ppo_policy.py:168-179: act = mu + sqrt(exp(clamp(logvar))) * eps (eps NULL => act = mu), logp [B] (nullable) */
int eg_gauss_sample(const float* out_actor, const float* eps, int B, int Z, float min_logvar, float max_logvar,
                    float* act, float* logp, void* stream);
/* one minibatch of GAMMAPPOPolicy.learn (:189-242): forward, loss, full backward into the flat gradient
 * buffer. adv_norm is the already normalised advantage; inv_B = 1 / (global minibatch size) so that
 * summing gradients over ranks reproduces the single-process mean. stats (device float[8], caller-zeroed):
 * 0 clip loss, 1 value loss, 2 entropy, 3 kld indicator, 4 mean(logp_old - logp_new). */
int eg_ppo_loss_backward(EgPolicy* h, const float* state, const float* ego, const float* dist, const float* time,
                         const float* act, const float* logp_old, const float* adv_norm, const float* returns,
                         int B, float inv_B, float eps_clip, float vf_coef, float ent_coef, float min_logvar,
                         float max_logvar, int zero_grads, float* stats, void* stream);
/* The same in two calls, so that a data-parallel caller can start summing the actor + critic gradients (the prefix of the
 * flat buffer, 94 % of it) over the ranks while the encoders' backward still runs: _mlp = forward, loss head and the
 * backward of the actor / critic chains; _encoders = the GRU encoders' backward (the shared_net tail). */
int eg_ppo_loss_backward_mlp(EgPolicy* h, const float* state, const float* ego, const float* dist, const float* time,
                             const float* act, const float* logp_old, const float* adv_norm, const float* returns,
                             int B, float inv_B, float eps_clip, float vf_coef, float ent_coef, float min_logvar,
                             float max_logvar, int zero_grads, float* stats, void* stream);
int eg_ppo_backward_encoders(EgPolicy* h, const float* ego, int B, void* stream);
/* out2 (device double[2]) = {sum x, sum x^2} */
int eg_moments(const float* x, int64_t n, double* out2, void* stream);
/* (adv - mean) / (std_unbiased + eps) from moments3 (device double[3]) = {sum, sumsq, count} (:192-195) */
int eg_adv_normalize(const float* adv, int n, const double* moments3, float eps, float* out, void* stream);
/* clip_grad_norm_ over the actor+critic prefix (max_grad_norm <= 0 disables) fused with one AdamW step */
int eg_clip_adamw_step(EgPolicy* h, float* exp_avg, float* exp_avg_sq, float max_grad_norm, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, void* stream);
/* Data-parallel form of  all_reduce(gradients) ; eg_clip_adamw_step  for `world` ranks of one NVLink domain, one process
 * per GPU (csrc/dp_optim.cu; same reference step, ppo_policy.py:241-247). The flat gradient and parameter vectors of every
 * rank live in peer-mapped (symmetric) memory, padded to n_pad = multiple of 4 * world floats; rank r owns elements
 * [r * n_pad / world, (r + 1) * n_pad / world).
 *   eg_dp_reduce_norm : gred [n_pad / world] = sum over ranks of the own gradient slice (grads_mc: multicast address of the
 *     gradient buffers -> in-switch reduction; NULL -> peer loads through grads_ptrs, summed in rank order), and this rank's
 *     share of the squared norm over the first n_clip elements is written to slot `rank` of EVERY rank's scratch
 *     (scratch_ptrs: `world` device pointers, each to double[world]). work: 16 zero-initialised device bytes.
 *   -- the caller runs a cross-rank barrier before (all backwards done) and after eg_dp_reduce_norm --
 *   eg_dp_adamw_gather: clip coefficient from the sum of the scratch slots, AdamW on the own slice (moment slices
 *     [n_pad / world], local), updated parameters stored to every rank (params_mc multicast, or peer stores).
 *   -- and a barrier after it, before any rank reads the parameters again.
 * grads_ptrs / params_ptrs / scratch_ptrs are HOST arrays of `world` device pointers (index = rank). */
/* eg_dp_reduce_norm restricted to the flat elements [elem_lo, elem_hi) (multiples of 4): the part of this rank's slice inside
 * the range is summed over the ranks; publish_norm != 0 also publishes the squared-norm share (use it on the call whose range
 * covers the clip prefix). */
int eg_dp_reduce_range(const void* const* grads_ptrs, const void* grads_mc, int world, int rank, int64_t n_pad,
                       int64_t n_clip, int64_t elem_lo, int64_t elem_hi, int publish_norm, float* gred,
                       const void* const* scratch_ptrs, void* work, void* stream);
int eg_dp_reduce_norm(const void* const* grads_ptrs, const void* grads_mc, int world, int rank, int64_t n_pad,
                      int64_t n_clip, float* gred, const void* const* scratch_ptrs, void* work, void* stream);
int eg_dp_adamw_gather(const void* const* params_ptrs, void* params_mc, int world, int rank, int64_t n_pad,
                       int64_t n_clip, const float* gred, const double* scratch_local, float* exp_avg_slice,
                       float* exp_avg_sq_slice, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, void* stream);
/* tianshou _gae_return on [T,E] time-major rollouts: v_next must already be critic(obs_next) unmasked;
 * terminated masks it, end_flag = terminated | truncated | last-stored-step. adv, ret [T,E]. */
int eg_gae(const float* v_s, const float* v_next, const float* rew, const uint8_t* terminated,
           const uint8_t* end_flag, int T, int E, double gamma, double gae_lambda, float* adv, float* ret,
           void* stream);

/* ------------------------------------------------------------------------------------------
 * C-VAE marker-predictor training (BASELINE config 3) - replaces GAMMAPrimitiveVAE.forward + the loss of
 * GAMMAPrimitiveVAETrainOP.calc_loss / one primitive of calc_loss_rollout
 * (motion/models/models_GAMMA_primitive.py:75-110, 400-432, 476-490) with a hand-written backward.
 * Flat parameter / gradient buffers in GAMMAPrimitiveVAE.parameters() order.
 * ------------------------------------------------------------------------------------------ */
typedef struct EgCvae EgCvae;
typedef struct EgCvaeDims { int32_t in_dim /*201*/, h_dim /*256*/, z_dim /*128*/, mlp_dim /*512*/; } EgCvaeDims;
int64_t eg_cvae_param_count(const EgCvaeDims* dims);
int eg_cvae_create(const EgCvaeDims* dims, float* params_flat, float* grads_flat, int device, EgCvae** out);
void eg_cvae_destroy(EgCvae* h);
/* X [2,B,201], Y [18,B,201] time-major, eps [B,128] the reparameterisation noise. Gradients are accumulated
 * (scaled by loss_scale); Y_rec [18,B,201] out; stats (device float[4], accumulated): loss, rec, kld. */
int eg_cvae_loss_backward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float w_rec, float w_td,
                          float w_kld, int robust_kld, float loss_scale, float* Y_rec, float* stats, void* stream);
/* The same primitive in two calls, for objectives that add a downstream loss on Y_rec before the backward
 * (GAMMAPrimitiveComboTrainOP.calc_loss_one, models_GAMMA_primitive.py:819-838): eg_cvae_forward_train keeps the activations;
 * eg_cvae_backward evaluates the losses and back-propagates. rec_in_loss = 0 reports the reconstruction term in stats[1]
 * but keeps it out of the objective (calc_loss_marker :797-815 without scheduled sampling); dY_extra [18,B,201] (nullable)
 * is added to dL/dY_rec. */
int eg_cvae_forward_train(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float* Y_rec, void* stream);
int eg_cvae_backward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float w_rec, float w_td, float w_kld,
                     int robust_kld, float loss_scale, int rec_in_loss, const float* Y_rec, const float* dY_extra,
                     float* stats, void* stream);
/* torch.optim.Adam / AdamW step on flat buffers (weight_decay is the decoupled AdamW form; 0 for Adam) */
int eg_adam_step_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

/* Body-regressor training step (GAMMARegressorTrainOP, motion/models/models_GAMMA_primitive.py:594-633 and the loop
 * :664-682): MoshRegressor forward (use_cont, relu; :222-301) on marker_ref [M,201] / betas [M,10], 6-D -> axis-angle,
 * SMPL-X markers of the regressed bodies, loss = L1(marker_ref, markers) + weight_reg_hpose * mean(hand_pca^2), and the
 * full backward into the flat gradient buffer (ZEROED first, like optimizer.zero_grad()). Flat buffers are in
 * MoshRegressor.parameters() order. xb_out [M,93] (axis-angle body parameters) may be null; stats = device float[3]:
 * loss, loss_marker, loss_hpose (overwritten). The step itself is eg_adam_step_flat. */
typedef struct EgRegTrain EgRegTrain;
typedef struct EgRegressorDims { int32_t in_dim /*201*/, h_dim /*128*/, n_blocks /*10*/, n_recur /*3*/, body_dim /*159*/; } EgRegressorDims;
int64_t eg_regressor_param_count(const EgRegressorDims* dims);
int eg_regressor_train_create(const EgRegressorDims* dims, float* params_flat, float* grads_flat, EgLbs* lbs, int device,
                              EgRegTrain** out);
void eg_regressor_train_destroy(EgRegTrain* h);
int eg_regressor_loss_backward(EgRegTrain* h, const float* marker_ref, const float* betas, int M, float weight_reg_hpose,
                               float* xb_out, float* stats, void* stream);
/* Regressor part of the combo objective (calc_loss_regressor, models_GAMMA_primitive.py:787-794): markers_in [T*B,201]
 * (t-major; the predictor's Y_rec) -> body parameters -> SMPL-X markers x_pred; loss_scale * (w_rec L1(Y_ref, x_pred) +
 * w_td L1(dt x_pred, dt Y_ref) + w_hpose mean(hand_pca^2)) and its gradient w.r.t. markers_in (d_markers_in [T*B,201],
 * overwritten; the regressor's weights are not trained by that op). want_grad = 0 only evaluates the terms (the no_grad
 * branch). stats (device float[2], ACCUMULATED): marker term, hand term. xb_out [T*B,93] may be null. */
int eg_regressor_cycle_backward(EgRegTrain* h, const float* markers_in, const float* betas, const float* Y_ref, int T, int B,
                                float w_rec, float w_td, float w_hpose, float loss_scale, int want_grad, float* d_markers_in,
                                float* xb_out, float* stats, void* stream);
/* SMPLXParser.update_transl_glorot, torch branch (baseops.py:537-598): re-express transl / global_orient of xb [N,93] in
 * the frame (transf_rotmat [N,3,3], transf_transl [N,3]) with the root-vs-pelvis offset compensation
 * (calc_calibrate_offset :494-534). delta_T [N,3] receives that offset; xb_out may alias xb (inplace=True). */
int eg_update_transl_glorot(EgLbs* lbs, const float* transf_rotmat, const float* transf_transl, const float* betas,
                            int betas_rows, const float* xb, int N, float* delta_T, float* xb_out, void* stream);
/* CanonicalCoordinateExtractor.get_new_coordinate_torch (baseops.py:214-225): joints of body b at joints + b*ld_body */
int eg_new_coordinate(const float* joints, int ld_body, int B, float* R, float* T, void* stream);
/* pts [nt,B,P,3]: inverse=0 -> R p + T, inverse=1 -> R^T (p - T)  (models_GAMMA_primitive.py:462-466) */
int eg_rigid_points(const float* R, const float* T, const float* pts, int nt, int B, int P, int inverse, float* out,
                    void* stream);

/* y[M,out] = act(x W^T + b) + residual, W [out,in] row-major (nn.Linear; baseops.py:615-641 MLP layers).
 * act: 0 none, 1 tanh, 2 relu, 3 leaky-relu(slope). */
/* generic product C[M,N] (+)= op(A) op(B) in the layouts the layers' forward / backward passes use (tests):
 * trans_a = 0: A[m*lda+k], 1: A[k*lda+m]; trans_b = 1: B[n*ldb+k] (nn.Linear weight), 0: B[k*ldb+n].
 * y = x W^T is (0,1), dX = dY W is (0,0), dW = dY^T X is (1,0). */
int eg_matmul(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, int M, int N, int K,
              float* C, int ldc, int accumulate, void* stream);
int eg_linear_forward(const float* x, int ldx, int M, const float* W, const float* b, int in_dim,
                      int out_dim, int act, float slope, const float* residual, int ldr, float* y,
                      int ldy, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOGEN_B200_H */
