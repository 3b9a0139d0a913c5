// SMPL-X linear blend skinning for sm_100a.
//
// Replaces smplx.SMPLX.forward / smplx.lbs.lbs as called from SMPLXParser.forward_smplx
// (reference motion/models/baseops.py:338-398) and the fused consumers in
// crowd_env_2f.py:133-177 (world transform -> calc_sdf -> feet skip -> per-frame counts).
//
// Data layout in HBM (built once in eg_lbs_create):
//   basis   [KPAD=512][3*n_pad]   rows 0..485 posedirs, 486..505 shapedirs^T, 506..511 zero;
//                                 columns are vertices (xyz interleaved) padded to 128-vertex tiles.
//                                 The shape blend is folded into the pose-blend contraction so the
//                                 vertex kernel is ONE [N,512]x[512,3V] product + skinning epilogue.
//   vt      [3*n_pad]             v_template
//   skin    [nnz][n_pad] idx / w  ELL form of lbs_weights (zeros dropped; exact)
//   Jt,Js   [55,3] / [55,3,20]    J_regressor folded through v_template / shapedirs
// Per call workspace: Ft [512][Npad] (features, k-major), A [N][55][12], Jp [N][55][3].
//
// Kernels: lbs_pose_prep_kernel (1 CTA / body: hand PCA, Rodrigues, features, joint regression,
// level-parallel kinematic chain) -> lbs_verts_kernel (128 vertices x 32 bodies per CTA,
// cp.async double-buffered basis tiles, skinning epilogue, optional fused world-transform + SDF
// sample + penetration count so vertices never reach HBM) -> lbs_finish_kernel (127 joints,
// markers). The same vertex kernel runs on the full mesh and on a compact gathered vertex set
// (markers + vertex joints + landmark corners) for the joints-only calls.
#include <stdlib.h>

#include <vector>
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "lbs_tc.cuh"

namespace eg {

constexpr int KPAD = 512;      // padded contraction length (486 pose + 20 shape + 6 zero)
constexpr int TILE_V = 128;    // vertices per CTA
constexpr int KC = 16;         // k-chunk per pipeline stage
constexpr int BPG = 16;        // bodies per thread (per body-group)
constexpr int NGROUPS = 2;     // body-groups per CTA
constexpr int TILE_B = BPG * NGROUPS;          // 32 bodies per CTA
constexpr int VERT_THREADS = TILE_V * NGROUPS; // 256
constexpr int MAXJ = 64;
constexpr int VERT_SMEM = sizeof(float) * 2 * KC * (TILE_V * 3 + TILE_B) + sizeof(int) * TILE_B;

struct VertexSet {
  int n = 0, n_pad = 0, nnz = 0;
  float* basis = nullptr;
  // tcgen05 path (full set only): joint-coherent vertex order, see build_tc_layout()
  __half* basisT = nullptr;  // [3][n_pad_tc][tc::KT] planar K-major fp16 copy of the basis, rows in tc order
  float4* tc_rec = nullptr;  // [n_pad_tc][3] per-vertex records
  int32_t* tc_jl = nullptr;  // [n_vt_tc][NJ_MAX]
  int32_t* tc_nj = nullptr;  // [n_vt_tc]
  uint32_t* tc_xoff = nullptr;
  float* tc_xw = nullptr;
  int n_pad_tc = 0, n_vt_tc = 0;
  float* vt = nullptr;
  int32_t* skin_idx = nullptr;
  float* skin_w = nullptr;
};

}  // namespace eg

struct EgLbs {
  int device = 0;
  int V = 0, J = 0, S = 0, P = 0, n_extra = 0, n_lmk = 0, n_markers = 0;
  int n_levels = 0;
  // prep-kernel constants
  float *Jt = nullptr, *Js = nullptr, *hand_l = nullptr, *hand_r = nullptr, *pose_mean = nullptr;
  int32_t *parents = nullptr, *level_joints = nullptr, *level_start = nullptr;
  float* lmk_bary = nullptr;
  eg::VertexSet full, compact;
  // host copies needed to (re)build the compact set
  std::vector<int32_t> h_extra, h_lmk_verts;
  std::vector<int32_t> h_sidx;          // [nnz][full.n_pad] skinning joints / weights / template of the full mesh (host)
  std::vector<float> h_sw, h_vt;
  std::vector<int32_t> h_row_of_vertex; // tc row of every mesh vertex in the full set's joint-coherent order
  // workspace
  int cap_N = 0;
  float *Ft = nullptr, *A = nullptr, *Jp = nullptr, *cout_ = nullptr;
  __half* Ftc = nullptr;      // [cap_Ntc][tc::KT] fp16 features for the tcgen05 mainloop
  float* Aw = nullptr;        // [J][cap_Ntc][12] joint-major transforms in the tensor-core epilogue's pair layout
  float4* rec_call = nullptr; // [n_pad_tc][3] per-call vertex records with the skip mask folded in
  int cap_Ntc = 0;
  int use_tc = 1;             // 1: tcgen05/TMEM/TMA mainloop for the full mesh, 0: SIMT mainloop
  unsigned int* tile_sched = nullptr;   // {ticket, finished} counters of the tcgen05 kernel's tile scheduler
  int max_clusters = 64;      // co-resident 2-CTA clusters of the tcgen05 kernel (cudaOccupancyMaxActiveClusters)
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
  CUtensorMap mapA, mapA_c;   // K-major basis of the full / compact set
  float* Al = nullptr;        // [J][cap_Ntc][12] joint-major LOCAL transforms (fused calls: Aw is world-composed)
};

namespace eg {

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// --------------------------------------------------------------------------------------------
// prep: one CTA of 64 threads per body
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
lbs_pose_prep_kernel(const float* __restrict__ xb, const float* __restrict__ betas, int betas_div,
                     int N, int Npad, int J, int S, int n_levels,
                     const float* __restrict__ hand_l, const float* __restrict__ hand_r,
                     const float* __restrict__ pose_mean, const float* __restrict__ Jt,
                     const float* __restrict__ Js, const int32_t* __restrict__ parents,
                     const int32_t* __restrict__ level_joints, const int32_t* __restrict__ level_start,
                     float* __restrict__ Ft, __half* __restrict__ Ftc, float* __restrict__ A,
                     float* __restrict__ Jp, const float* __restrict__ R0w, const float* __restrict__ T0w,
                     int frames_per_env, float* __restrict__ Aw, int AwRows, float* __restrict__ Al) {
  const int n = blockIdx.x;
  const int t = threadIdx.x;
  __shared__ float pose[MAXJ * 3];
  __shared__ float R[MAXJ][9];
  __shared__ float Jr[MAXJ][3];
  __shared__ float G[MAXJ][12];
  __shared__ float shape[32];
  const float* x = xb + (int64_t)n * EG_XB_DIM;
  const float* be = betas + (int64_t)(n / betas_div) * 10;

  // full_pose = [global_orient, body_pose(21), jaw, leye, reye, lhand(15), rhand(15)] + pose_mean
  for (int i = t; i < J * 3; i += 64) {
    float v = 0.0f;
    if (i < 66) {
      v = x[3 + i];
    } else if (i >= 75 && i < 120) {          // left hand: einsum('bi,ij->bj', pca12, comps[12,45])
      const int c = i - 75;
      for (int k = 0; k < 12; ++k) v += x[69 + k] * __ldg(hand_l + k * 45 + c);
    } else if (i >= 120 && i < 165) {
      const int c = i - 120;
      for (int k = 0; k < 12; ++k) v += x[81 + k] * __ldg(hand_r + k * 45 + c);
    }
    pose[i] = v + __ldg(pose_mean + i);
  }
  if (t < S) shape[t] = (t < 10) ? be[t] : 0.0f;   // shape components = betas ++ expression(0)
  __syncthreads();

  // Rodrigues (smplx.lbs.batch_rodrigues): angle = ||r + 1e-8||, R = I + sin K + (1-cos) K K
  if (t < J) {
    const float rx = pose[3 * t], ry = pose[3 * t + 1], rz = pose[3 * t + 2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float dx = rx / angle, dy = ry / angle, dz = rz / angle;
    const float s = sinf(angle), c = cosf(angle);
    const float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
    const float omc = 1.0f - c;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float kk = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) kk += K[a * 3 + m] * K[m * 3 + b];
        R[t][a * 3 + b] = (a == b ? 1.0f : 0.0f) + s * K[a * 3 + b] + omc * kk;
      }
    // joint regression folded through the blend shapes: J = Jt + Js . shape
    for (int cidx = 0; cidx < 3; ++cidx) {
      float v = __ldg(Jt + t * 3 + cidx);
      const float* js = Js + (t * 3 + cidx) * S;
      for (int k = 0; k < S; ++k) v += __ldg(js + k) * shape[k];
      Jr[t][cidx] = v;
    }
  }
  __syncthreads();

  // features, k-major: rows 0..485 = vec(R_1..54 - I), 486..505 = shape, rest 0
  for (int k = t; k < KPAD; k += 64) {
    float v = 0.0f;
    const int npose = (J - 1) * 9;
    if (k < npose) {
      const int j = 1 + k / 9, e = k % 9;
      v = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
    } else if (k < npose + S) {
      v = shape[k - npose];
    }
    Ft[(int64_t)k * Npad + n] = v;
  }
  // same features for the tensor-core mainloop: row-major [n][KT] fp16 (round-to-nearest), shape
  // coefficients split hi/lo and laid against the hi/hi, lo/hi, hi/lo shape rows of basisT
  if (Ftc != nullptr) {
    const int npose = (J - 1) * 9;
    for (int k = t; k < tc::KT; k += 64) {
      float v = 0.0f;
      if (k < npose) {
        const int j = 1 + k / 9, e = k % 9;
        v = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
      } else if (k < npose + 3 * S) {
        const int seg = (k - npose) / S, q = (k - npose) % S;
        const float hi = __half2float(__float2half_rn(shape[q]));
        v = seg == 1 ? shape[q] - hi : hi;
      }
      Ftc[(int64_t)n * tc::KT + k] = __float2half_rn(v);
    }
  }

  // kinematic chain by tree level: G_j = G_parent [R_j | J_j - J_parent]
  for (int lv = 0; lv < n_levels; ++lv) {
    const int beg = __ldg(level_start + lv), end = __ldg(level_start + lv + 1);
    for (int q = beg + t; q < end; q += 64) {
      const int j = __ldg(level_joints + q);
      const int p = __ldg(parents + j);
      if (p < 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          G[j][a * 4 + 0] = R[j][a * 3 + 0];
          G[j][a * 4 + 1] = R[j][a * 3 + 1];
          G[j][a * 4 + 2] = R[j][a * 3 + 2];
          G[j][a * 4 + 3] = Jr[j][a];
        }
      } else {
        const float r0 = Jr[j][0] - Jr[p][0], r1 = Jr[j][1] - Jr[p][1], r2 = Jr[j][2] - Jr[p][2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float g0 = G[p][a * 4 + 0], g1 = G[p][a * 4 + 1], g2 = G[p][a * 4 + 2], g3 = G[p][a * 4 + 3];
#pragma unroll
          for (int b = 0; b < 3; ++b)
            G[j][a * 4 + b] = g0 * R[j][0 * 3 + b] + g1 * R[j][1 * 3 + b] + g2 * R[j][2 * 3 + b];
          G[j][a * 4 + 3] = g0 * r0 + g1 * r1 + g2 * r2 + g3;
        }
      }
    }
    __syncthreads();
  }

  // posed joints and relative transforms A_j = G_j with translation minus G_j.R J_j
  if (t < J) {
    float* a_out = A + ((int64_t)n * J + t) * 12;
    float* jp = Jp + ((int64_t)n * J + t) * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float g0 = G[t][a * 4 + 0], g1 = G[t][a * 4 + 1], g2 = G[t][a * 4 + 2], g3 = G[t][a * 4 + 3];
      a_out[a * 4 + 0] = g0;
      a_out[a * 4 + 1] = g1;
      a_out[a * 4 + 2] = g2;
      a_out[a * 4 + 3] = g3 - (g0 * Jr[t][0] + g1 * Jr[t][1] + g2 * Jr[t][2]);
      jp[a] = g3;
    }
    // copy for the tensor-core epilogue (Aw): world-composed R0 (A [v;1] + transl) + T0 when R0w is given (skinning
    // weights sum to 1, so the translation part can ride inside every A_j), the local A_j otherwise. Stored in the
    // epilogue's packed-FMA order {m00,m10,m01,m11 | m02,m12,m03,m13 | m20,m21,m22,m23}.
    if (Aw != nullptr) {
      float M[3][4];
      if (R0w != nullptr) {
        const int e = n / frames_per_env;
        const float* Rw = R0w + (int64_t)e * 9;
        const float* Tw = T0w + (int64_t)e * 3;
        float col[4][3];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 3; ++a) col[b][a] = a_out[a * 4 + b];
        col[3][0] += x[0]; col[3][1] += x[1]; col[3][2] += x[2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            M[a][b] = Rw[a * 3 + 0] * col[b][0] + Rw[a * 3 + 1] * col[b][1] + Rw[a * 3 + 2] * col[b][2] +
                      (b == 3 ? Tw[a] : 0.0f);
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) M[a][b] = a_out[a * 4 + b];
      }
      float4* w_out = reinterpret_cast<float4*>(Aw + ((int64_t)t * AwRows + n) * 12);   // joint-major [J][AwRows][12]
      w_out[0] = make_float4(M[0][0], M[1][0], M[0][1], M[1][1]);
      w_out[1] = make_float4(M[0][2], M[1][2], M[0][3], M[1][3]);
      w_out[2] = make_float4(M[2][0], M[2][1], M[2][2], M[2][3]);
    }
    if (Al != nullptr) {                       // the same layout with the LOCAL transforms (compact set beside a fused call)
      float4* l_out = reinterpret_cast<float4*>(Al + ((int64_t)t * AwRows + n) * 12);
      l_out[0] = make_float4(a_out[0], a_out[4], a_out[1], a_out[5]);
      l_out[1] = make_float4(a_out[2], a_out[6], a_out[3], a_out[7]);
      l_out[2] = make_float4(a_out[8], a_out[9], a_out[10], a_out[11]);
    }
  }
}

// --------------------------------------------------------------------------------------------
// vertex kernel
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct VertArgs {
  const float* basis;     // [KPAD][3*n_pad]
  const float* vt;        // [3*n_pad]
  const int32_t* skin_idx;
  const float* skin_w;
  int n_real, n_pad, nnz, J;
  const float* Ft;        // [KPAD][Npad]
  const float* A;         // [N][J][12]
  const float* xb;        // [N][93] (transl)
  int N, Npad;
  int add_transl;
  float* out;             // [N][n_real][3] or null
  // fused SDF
  SdfGrid sdf;
  const float* R0;        // [E][9]
  const float* T0;        // [E][3]
  int frames_per_env;
  const uint8_t* skip;
  int32_t* counts;        // [N]
  // tensor-core layout (joint-coherent vertex order; see lbs_tc.cuh)
  const float4* tc_rec;   // [n_pad_tc][3]: {vt.xyz, original id}, {w0..w3}, {slot byte offsets 0..3}
  const int32_t* tc_jl;   // [n_vt][NJ_MAX] joints of each vertex tile
  const int32_t* tc_nj;   // [n_vt]
  const uint32_t* tc_xoff;  // [(nnz-4)][n_pad_tc] slot byte offsets of the skinning entries beyond the 4th
  const float* tc_xw;
  int n_pad_tc;
  int A_rows;             // tc path: A is joint-major [J][A_rows][12]
  unsigned int* tile_sched;   // tc path: {ticket counter, finished-CTA counter}; both return to 0 when a launch ends
};

// One (vertex, body) of the epilogue shared by the SIMT and tcgen05 kernels: skinning T = sum_k w_k A[n][j_k],
// v = T [v_posed; 1] (+ transl), optional store, optional world transform + calc_sdf sample; returns sdf < 0.
template <bool FUSE_SDF>
__device__ __forceinline__ bool vertex_epilogue(const VertArgs& a, int v, bool v_ok, int n, float px, float py,
                                                float pz, bool skip, float cx, float cy, float cz, float sc) {
  float T[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.0f;
  const float4* An = reinterpret_cast<const float4*>(a.A + (int64_t)n * a.J * 12);
  for (int k = 0; k < a.nnz; ++k) {
    const int j = a.skin_idx[k * a.n_pad + v];
    const float w = a.skin_w[k * a.n_pad + v];
    const float4 r0 = __ldg(An + j * 3 + 0), r1 = __ldg(An + j * 3 + 1), r2 = __ldg(An + j * 3 + 2);
    T[0] += w * r0.x; T[1] += w * r0.y; T[2] += w * r0.z; T[3] += w * r0.w;
    T[4] += w * r1.x; T[5] += w * r1.y; T[6] += w * r1.z; T[7] += w * r1.w;
    T[8] += w * r2.x; T[9] += w * r2.y; T[10] += w * r2.z; T[11] += w * r2.w;
  }
  float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
  float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
  float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
  if (a.add_transl) {
    const float* x = a.xb + (int64_t)n * EG_XB_DIM;
    ox = __fadd_rn(ox, __ldg(x)); oy = __fadd_rn(oy, __ldg(x + 1)); oz = __fadd_rn(oz, __ldg(x + 2));
  }
  if (a.out != nullptr && v_ok) {
    float* o = a.out + ((int64_t)n * a.n_real + v) * 3;
    o[0] = ox; o[1] = oy; o[2] = oz;
  }
  bool neg = false;
  if (FUSE_SDF) {
    const int e = n / a.frames_per_env;
    const float* Rw = a.R0 + (int64_t)e * 9;
    const float* Tw = a.T0 + (int64_t)e * 3;
    const float wx = __ldg(Rw + 0) * ox + __ldg(Rw + 1) * oy + __ldg(Rw + 2) * oz + __ldg(Tw + 0);
    const float wy = __ldg(Rw + 3) * ox + __ldg(Rw + 4) * oy + __ldg(Rw + 5) * oz + __ldg(Tw + 1);
    const float wz = __ldg(Rw + 6) * ox + __ldg(Rw + 7) * oy + __ldg(Rw + 8) * oz + __ldg(Tw + 2);
    if (!skip) neg = sdf_is_negative(a.sdf, cx, cy, cz, sc, wx, wy, wz);
  }
  return neg;
}

template <bool FUSE_SDF>
__global__ void __launch_bounds__(VERT_THREADS, 2)
lbs_verts_kernel(const VertArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*Ps)[KC][TILE_V * 3] = reinterpret_cast<float (*)[KC][TILE_V * 3]>(smem_raw);
  float (*Fs)[KC][TILE_B] = reinterpret_cast<float (*)[KC][TILE_B]>(smem_raw + sizeof(float) * 2 * KC * TILE_V * 3);
  int* cnt_s = reinterpret_cast<int*>(smem_raw + sizeof(float) * 2 * KC * (TILE_V * 3 + TILE_B));

  const int tid = threadIdx.x;
  const int vl = tid % TILE_V;            // vertex within tile
  const int grp = tid / TILE_V;           // body group
  const int v0 = blockIdx.x * TILE_V;
  const int b0 = blockIdx.y * TILE_B;
  const int64_t brow = (int64_t)a.n_pad * 3;

  if (FUSE_SDF && tid < TILE_B) cnt_s[tid] = 0;

  auto load_stage = [&](int st, int kbase) {
    // basis tile: KC rows x 384 floats = KC*96 float4
#pragma unroll
    for (int i = 0; i < (KC * TILE_V * 3 / 4) / VERT_THREADS; ++i) {
      const int q = tid + i * VERT_THREADS;
      const int r = q / (TILE_V * 3 / 4), c4 = q % (TILE_V * 3 / 4);
      cp_async16(&Ps[st][r][c4 * 4], a.basis + (int64_t)(kbase + r) * brow + (int64_t)v0 * 3 + c4 * 4);
    }
    // feature tile: KC rows x 32 floats = KC*8 float4
    if (tid < KC * TILE_B / 4) {
      const int r = tid / (TILE_B / 4), c4 = tid % (TILE_B / 4);
      cp_async16(&Fs[st][r][c4 * 4], a.Ft + (int64_t)(kbase + r) * a.Npad + b0 + c4 * 4);
    }
    cp_async_commit();
  };

  float acc[BPG][3];
#pragma unroll
  for (int b = 0; b < BPG; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.0f;

  load_stage(0, 0);
  constexpr int NCHUNK = KPAD / KC;
  for (int ch = 0; ch < NCHUNK; ++ch) {
    const int st = ch & 1;
    if (ch + 1 < NCHUNK) {
      load_stage(st ^ 1, (ch + 1) * KC);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float p0 = Ps[st][kk][vl * 3 + 0];
      const float p1 = Ps[st][kk][vl * 3 + 1];
      const float p2 = Ps[st][kk][vl * 3 + 2];
      const float4* f4 = reinterpret_cast<const float4*>(&Fs[st][kk][grp * BPG]);
#pragma unroll
      for (int q = 0; q < BPG / 4; ++q) {
        const float4 f = f4[q];
        acc[q * 4 + 0][0] += f.x * p0; acc[q * 4 + 0][1] += f.x * p1; acc[q * 4 + 0][2] += f.x * p2;
        acc[q * 4 + 1][0] += f.y * p0; acc[q * 4 + 1][1] += f.y * p1; acc[q * 4 + 1][2] += f.y * p2;
        acc[q * 4 + 2][0] += f.z * p0; acc[q * 4 + 2][1] += f.z * p1; acc[q * 4 + 2][2] += f.z * p2;
        acc[q * 4 + 3][0] += f.w * p0; acc[q * 4 + 3][1] += f.w * p1; acc[q * 4 + 3][2] += f.w * p2;
      }
    }
    __syncthreads();
  }

  // ---- epilogue: v_posed -> skinning -> (+transl) -> store / fused SDF ----
  const int v = v0 + vl;
  const bool v_ok = v < a.n_real;
  const float t0 = a.vt[v * 3 + 0], t1 = a.vt[v * 3 + 1], t2 = a.vt[v * 3 + 2];
  float cx = 0.f, cy = 0.f, cz = 0.f, sc = 0.f;
  bool skip = false;
  if (FUSE_SDF) {
    cx = __ldg(a.sdf.center); cy = __ldg(a.sdf.center + 1); cz = __ldg(a.sdf.center + 2);
    sc = __ldg(a.sdf.scale);
    skip = v_ok ? (a.skip != nullptr && a.skip[v] != 0) : true;
  }
#pragma unroll
  for (int b = 0; b < BPG; ++b) {
    const int n = b0 + grp * BPG + b;
    if (n >= a.N) continue;                               // warp-uniform (n depends on grp, b only)
    const bool neg = vertex_epilogue<FUSE_SDF>(a, v, v_ok, n, t0 + acc[b][0], t1 + acc[b][1], t2 + acc[b][2], skip,
                                               cx, cy, cz, sc);
    if (FUSE_SDF) {
      const unsigned m = __ballot_sync(0xffffffffu, neg);
      if ((tid & 31) == 0 && m) atomicAdd(&cnt_s[grp * BPG + b], __popc(m));
    }
  }
  if (FUSE_SDF) {
    __syncthreads();
    if (tid < TILE_B && b0 + tid < a.N && cnt_s[tid] > 0) atomicAdd(a.counts + b0 + tid, cnt_s[tid]);
  }
}

// --------------------------------------------------------------------------------------------
// tcgen05 vertex kernel (full mesh): persistent, warp-specialised; see lbs_tc.cuh for the tile plan
// --------------------------------------------------------------------------------------------
#ifndef EG_LBS_PROF
#define EG_LBS_PROF 0
#endif
#if EG_LBS_PROF
// per-CTA stall accounting of the tc kernel (debug builds only: EG_NVCC_EXTRA=-DEG_LBS_PROF=1; read with eg_lbs_prof_dump)
__device__ unsigned long long g_lbs_prof[160][48];
__device__ unsigned long long g_lbs_prof_vt[512];     // epilogue work clocks of warp 4, summed per vertex tile
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PROF_T0() const long long _pt0 = clock64()
#define PROF_ADD(slot) do { if (lane == 0) g_lbs_prof[blockIdx.x][slot] += (unsigned long long)(clock64() - _pt0); } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(slot) do {} while (0)
#endif

template <bool FUSE_SDF, bool XW>      // XW: some vertex has more than 4 skinning weights (extras read uncached)
__global__ void __cluster_dims__(tc::CLUSTER, 1, 1) __launch_bounds__(tc::THREADS, 1)
lbs_verts_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const VertArgs a, int n_vt, int n_bt) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;                 // [STAGES] operand ring
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2] accumulator sets
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* tab_full = tmem_empty + 2;       // [2] joint-transform table + vertex records
  uint64_t* tab_empty = tab_full + 2;        // [2]
  uint64_t* sched_full = tab_empty + 2;      // [NSCHED] tile-scheduler ring
  uint64_t* sched_empty = sched_full + NSCHED;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sched_empty + NSCHED);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_ptr + 1);   // [NSCHED]
  const uint32_t smem_base = smem_u32(smem);
  static_assert(CLUSTER == 1, "the ticket scheduler hands tiles to single CTAs");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = 0;
  const int n_btp = n_bt;
  const int n_tiles = n_vt * n_btp;
  // every consumer role walks the scheduler ring with its own cursor: next tile or -1
  int sc_slot = 0; uint32_t sc_phase = 0;
  auto next_tile = [&](bool whole_warp) -> int {     // whole_warp: all 32 lanes call it (else lane 0 alone)
    mbar_wait(&sched_full[sc_slot], sc_phase);
    const int t = sched_tile[sc_slot];
    if (whole_warp) __syncwarp();
    if (lane == 0) mbar_arrive(&sched_empty[sc_slot]);
    if (++sc_slot == NSCHED) { sc_slot = 0; sc_phase ^= 1u; }
    return t < n_tiles ? t : -1;
  };

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CLUSTER); }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS);
        mbar_init(&tab_full[b], 1); mbar_init(&tab_empty[b], EPI_WARPS);
      }
      for (int s = 0; s < NSCHED; ++s) { mbar_init(&sched_full[s], 1); mbar_init(&sched_empty[s], 3 + EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                     // barriers of BOTH CTAs are initialised before any remote arrive / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== operand producer: TMA ring =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = next_tile(false); tile >= 0; tile = next_tile(false)) {
        int vt, bt;
        tile_coords(tile, n_vt, n_btp, vt, bt);
        for (int ch = 0; ch < NCHUNK; ++ch) {
          { PROF_T0(); mbar_wait_relaxed(&empty_bar[stage], phase ^ 1); PROF_ADD(26); }
          unsigned char* st = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
          for (int c = 0; c < 3; ++c)       // my half of the basis tile, multicast to both CTAs of the pair
            tma_load_2d_mc(smem_base + stage * STAGE_BYTES + c * V_TILE_BYTES + rank * HALF_BYTES, &mapA, &full_bar[stage],
                           ch * BKT, c * a.n_pad_tc + vt * TV + rank * HALF_ROWS, (uint16_t)((1u << CLUSTER) - 1u));
          tma_load_2d(st + 3 * V_TILE_BYTES, &mapB, &full_bar[stage], ch * BKT, bt * TB);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===== table producer: joint-transform table + vertex records of each tile (bulk copies), decoupled from the ring =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = next_tile(false); tile >= 0; tile = next_tile(false), ++it) {
        int vt, bt;
        tile_coords(tile, n_vt, n_btp, vt, bt);
        const uint32_t tb = it % NTAB, tph = (it / NTAB) & 1u;
        { PROF_T0(); mbar_wait_relaxed(&tab_empty[tb], tph ^ 1u); PROF_ADD(27); }   // epilogue of tile it-NTAB is done with this buffer
        const int nj = __ldg(a.tc_nj + vt);
        mbar_expect_tx(&tab_full[tb], (uint32_t)nj * SLOT_BYTES + REC_BYTES);
        for (int s = 0; s < nj; ++s) {
          const int j = __ldg(a.tc_jl + vt * NJ_MAX + s);
          bulk_load(smem_base + OFF_TAB + tb * TAB_BYTES + s * SLOT_BYTES,
                    a.A + ((int64_t)j * a.A_rows + (int64_t)bt * TB) * 12, SLOT_BYTES, &tab_full[tb]);
        }
        bulk_load(smem_base + OFF_REC + tb * REC_BYTES, a.tc_rec + (int64_t)vt * TV * 3, REC_BYTES, &tab_full[tb]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A = feature tile (bodies -> TMEM lanes), B = basis tile of component c (vertices -> columns) =====
    int stage = 0; uint32_t phase = 0, it = 0;
    for (int tile = next_tile(true); tile >= 0; tile = next_tile(true), ++it) {
      const uint32_t buf = it & 1u;
      { PROF_T0(); mbar_wait(&tmem_empty[buf], ((it >> 1) & 1u) ^ 1u); PROF_ADD(24); }   // epilogue of tile it-2 has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ch = 0; ch < NCHUNK; ++ch) {
        { PROF_T0(); mbar_wait(&full_bar[stage], phase); PROF_ADD(25); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        PROF_T0();
        if (lane == 0) {
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          const uint64_t df = make_desc(sbase + 3 * V_TILE_BYTES);
          const uint64_t dbs = make_desc(sbase);            // x, y, z basis tiles = one 240-row tile -> columns c * TV + v
#pragma unroll
          for (int kk = 0; kk < BKT / UMMA_K; ++kk)       // UMMA_K = 16 fp16: advance 32 B inside the swizzle atom
            umma_f16(tmem_base + buf * ACC_COLS, df + (uint64_t)(kk * 2), dbs + (uint64_t)(kk * 2), (ch | kk) ? 1u : 0u);
          umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << CLUSTER) - 1u));   // frees the slot in BOTH CTAs' rings
          if (ch == NCHUNK - 1) umma_commit(&tmem_full[buf]);   // accumulators complete
        }
        __syncwarp();
        PROF_ADD(32);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===== tile scheduler: first tile = this CTA's index, then tickets; the entry past the last tile ends every role =====
    if (lane == 0) {
      int slot = 0; uint32_t phase = 0;
      int tile = (int)blockIdx.x;
      for (;;) {
        mbar_wait_relaxed(&sched_empty[slot], phase ^ 1u);
        sched_tile[slot] = tile;
        mbar_arrive(&sched_full[slot]);
        if (tile >= n_tiles) break;
        tile = (int)gridDim.x + (int)atomicAdd(a.tile_sched, 1u);
        if (++slot == NSCHED) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = body (TMEM lane q*32+lane), warp `sub` of the quarter walks vertices [sub*VPW, +VPW) =====
    const int q = warp & 3, sub = (warp - 4) >> 2, eidx = tid - 128;
    constexpr int NEPI = EPI_WARPS * 32;
    // SDF: coarse-cell sign bits in shared memory (a body-per-lane warp scatters its lookups over the whole grid;
    // 1 bit per cell keeps them on chip) addressed with one FMA per axis; the bit grid is dilated by one fine
    // corner (eg_sdf_prepare), which covers the rounding difference to the exact index sequence.
    float cx = 0.f, cy = 0.f, cz = 0.f, sc = 0.f, gax = 0.f, gay = 0.f, gaz = 0.f, gbx = 0.f, gby = 0.f, gbz = 0.f;
    const uint32_t mask_u32 = smem_base + OFF_MASK;
    const bool smem_mask = FUSE_SDF && a.sdf.coarse_bits != nullptr && a.sdf.cell_class != nullptr && a.sdf.n_bit_words <= MASK_WORDS;
    if (FUSE_SDF) {
      cx = __ldg(a.sdf.center); cy = __ldg(a.sdf.center + 1); cz = __ldg(a.sdf.center + 2);
      sc = __ldg(a.sdf.scale);
      // fine index = x * (s D / 2) + ((1 - c s) D - 1) / 2; the epilogue wants the 8^3-cell index, so both are / 8
      const float inv = 1.0f / (float)(1 << kCoarseShift);
      gax = sc * (float)a.sdf.D0 * 0.5f * inv; gbx = ((1.0f - cx * sc) * (float)a.sdf.D0 - 1.0f) * 0.5f * inv;
      gay = sc * (float)a.sdf.D1 * 0.5f * inv; gby = ((1.0f - cy * sc) * (float)a.sdf.D1 - 1.0f) * 0.5f * inv;
      gaz = sc * (float)a.sdf.D2 * 0.5f * inv; gbz = ((1.0f - cz * sc) * (float)a.sdf.D2 - 1.0f) * 0.5f * inv;
      if (smem_mask) {
        uint32_t* mk = reinterpret_cast<uint32_t*>(smem + OFF_MASK);
        for (int i = eidx; i < a.sdf.n_bit_words; i += NEPI) mk[i] = __ldg(a.sdf.coarse_bits + i);
      }
    }
    // level-1 cell index without F2I (a quarter-rate pipe, 12 per vertex group): s = sat((x' - 0.5) / Nx) with x' the
    // coordinate in 8^3-cell units and Nx = C - 0.51, then s * Nx + 1.5 * 2^23 rounds to the integer cell in the low
    // mantissa bits: round-to-nearest of x' - 0.5 is floor(x') except on exact cell boundaries (either neighbour; the
    // bit grid's one-corner dilation covers both), s = 1 lands on the last cell, s = 0 on the first.
    constexpr float kMagic = 12582912.0f;                           // bits 0x4B400000
    float nn0 = (float)a.sdf.C0 - 0.51f, nn1 = (float)a.sdf.C1 - 0.51f, nn2 = (float)a.sdf.C2 - 0.51f;
    float na0 = gax / nn0, nb0 = (gbx - 0.5f) / nn0, na1 = gay / nn1, nb1 = (gby - 0.5f) / nn1, na2 = gaz / nn2, nb2 = (gbz - 0.5f) / nn2;
    const uint32_t kc = 0x4B400000u * ((uint32_t)a.sdf.C1 * (uint32_t)a.sdf.C2 + (uint32_t)a.sdf.C2 + 1u);
    // keep the nine constants in registers (ptxas otherwise re-derives them from the kernel parameters in every group)
    asm volatile("" : "+f"(na0), "+f"(nb0), "+f"(nn0), "+f"(na1), "+f"(nb1), "+f"(nn1), "+f"(na2), "+f"(nb2), "+f"(nn2));
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");         // sign bits visible to every epilogue warp
    // SDF work queues of this warp (see lbs_tc.cuh): entry = {world x, y, z, body}; stage 1 is a stack growing up from
    // entry 0, stage 2 a stack growing down from entry QCAP-1; n1 + n2 <= 63 + 31 at any time
    const uint32_t qbase = smem_base + OFF_Q + (uint32_t)(warp - 4) * Q_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int n1 = 0, n2 = 0;                                             // warp-uniform
    // one batch (<= 32 entries) of stage 2: the exact trilinear sample, one atomic per negative sample
    auto drain2_batch = [&]() {
      __syncwarp();
      const int take = min(n2, 32);
      n2 -= take;
      if (lane < take) {
        const float4 e = lds128(qbase + (uint32_t)(QCAP - 1 - (n2 + lane)) * 16u);
        if (sdf_exact_negative(a.sdf.grid, a.sdf.D0, a.sdf.D1, a.sdf.D2, cx, cy, cz, sc, e.x, e.y, e.z))
          atomicAdd(a.counts + __float_as_int(e.w), 1);
      }
      __syncwarp();
    };
    // one batch of stage 1: the exact class of the sample's cell (same index sequence as the sampler, one 8-byte load):
    // every corner > 0 -> counted here, no corner > 0 -> dropped, corners of both signs -> stage 2, which is drained
    // when it holds a full batch
    auto drain1_batch = [&]() {
      __syncwarp();
      const int take = min(n1, 32);
      n1 -= take;
      bool pass = false;
      float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < take) {
        e = lds128(qbase + (uint32_t)(n1 + lane) * 16u);
        const int x0 = (int)floorf(sdf_unnormalize(__fmul_rn(__fsub_rn(e.x, cx), sc), a.sdf.D0));
        const int y0 = (int)floorf(sdf_unnormalize(__fmul_rn(__fsub_rn(e.y, cy), sc), a.sdf.D1));
        const int z0 = (int)floorf(sdf_unnormalize(__fmul_rn(__fsub_rn(e.z, cz), sc), a.sdf.D2));
        const uint32_t ci = ((uint32_t)x0 * (uint32_t)a.sdf.D1 + (uint32_t)y0) * (uint32_t)a.sdf.D2 + (uint32_t)z0;
        const uint2 cl = __ldg(a.sdf.cell_class + (ci >> 5));
        pass = ((cl.y >> (ci & 31u)) & 1u) != 0u;
        if ((cl.x >> (ci & 31u)) & 1u) atomicAdd(a.counts + __float_as_int(e.w), 1);
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (pass) sts128(qbase + (uint32_t)(QCAP - 1 - (n2 + __popc(m & lt_mask))) * 16u, e);
      n2 += __popc(m);
      while (n2 >= 32) drain2_batch();
    };
    // wait for a tile buffer; a warp that is ahead of its CTA resolves queued SDF work (partial batches too) instead of
    // spinning, so the forced drains inside the vertex walk - the main source of skew between the warps - become rare
    auto wait_busy = [&](uint64_t* bar, uint32_t parity) {
      while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if (FUSE_SDF) {
          if (n2 > 0) drain2_batch();
          else if (n1 > 0) drain1_batch();
        }
      }
    };
    uint32_t it = 0;
    for (int tile = next_tile(true); tile >= 0; tile = next_tile(true), ++it) {
      int vt, bt;
      tile_coords(tile, n_vt, n_btp, vt, bt);
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const uint32_t tb = it % NTAB, tph = (it / NTAB) & 1u;
      const uint32_t my_tab = smem_base + OFF_TAB + tb * TAB_BYTES + (uint32_t)(q * 32 + lane) * 48u;
      const uint32_t my_rec = smem_base + OFF_REC + tb * REC_BYTES + (uint32_t)(sub * VPW) * 48u;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + (uint32_t)(sub * VPW);
      const int n = bt * TB + q * 32 + lane;
      const bool n_ok = n < a.N;
      float trx = 0.f, try_ = 0.f, trz = 0.f;
      if (!FUSE_SDF && n_ok && a.add_transl) { const float* x = a.xb + (int64_t)n * EG_XB_DIM; trx = __ldg(x); try_ = __ldg(x + 1); trz = __ldg(x + 2); }
      const float n_bits = __int_as_float(n);
      int cnt = 0;
      float2 C[4][6];                                               // register cache: slot k -> transform of MY body
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 6; ++e) C[k][e] = make_float2(0.f, 0.f);
      const int vbase = vt * TV + sub * VPW;
      { PROF_T0(); wait_busy(&tab_full[tb], tph); PROF_ADD(warp - 4); }          // table + records landed
      { PROF_T0(); wait_busy(&tmem_full[buf], ph); PROF_ADD(8 + warp - 4); }     // accumulators complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      PROF_T0();
#pragma unroll 1
      for (int g = 0; g < VPW / 4; ++g) {
        float ax[4], ay[4], az[4];
        tmem_ld4(trow + (uint32_t)(g * 4), ax);
        tmem_ld4(trow + TV + (uint32_t)(g * 4), ay);
        tmem_ld4(trow + 2 * TV + (uint32_t)(g * 4), az);
        float4 R0[4], R1[4], R2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                               // warp-uniform addresses: broadcast loads
          const uint32_t ra = my_rec + (uint32_t)(g * 4 + u) * 48u;
          R0[u] = lds128(ra); R1[u] = lds128(ra + 16u); R2[u] = lds128(ra + 32u);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float ox[4], oy[4], oz[4];
        int pv[4];
        // blend T = sum_k w_k C[k] (packed FMAs) and apply it to the posed-blend vertex
        auto blend_apply = [&](int u, const float4& r0, const float4& r1) {
          float2 w2 = make_float2(r1.x, r1.x);
          float2 c0 = __fmul2_rn(w2, C[0][0]), c1 = __fmul2_rn(w2, C[0][1]), c2 = __fmul2_rn(w2, C[0][2]);
          float2 c3 = __fmul2_rn(w2, C[0][3]), z0 = __fmul2_rn(w2, C[0][4]), z1 = __fmul2_rn(w2, C[0][5]);
          w2 = make_float2(r1.y, r1.y);
          c0 = __ffma2_rn(w2, C[1][0], c0); c1 = __ffma2_rn(w2, C[1][1], c1); c2 = __ffma2_rn(w2, C[1][2], c2);
          c3 = __ffma2_rn(w2, C[1][3], c3); z0 = __ffma2_rn(w2, C[1][4], z0); z1 = __ffma2_rn(w2, C[1][5], z1);
          w2 = make_float2(r1.z, r1.z);
          c0 = __ffma2_rn(w2, C[2][0], c0); c1 = __ffma2_rn(w2, C[2][1], c1); c2 = __ffma2_rn(w2, C[2][2], c2);
          c3 = __ffma2_rn(w2, C[2][3], c3); z0 = __ffma2_rn(w2, C[2][4], z0); z1 = __ffma2_rn(w2, C[2][5], z1);
          w2 = make_float2(r1.w, r1.w);
          c0 = __ffma2_rn(w2, C[3][0], c0); c1 = __ffma2_rn(w2, C[3][1], c1); c2 = __ffma2_rn(w2, C[3][2], c2);
          c3 = __ffma2_rn(w2, C[3][3], c3); z0 = __ffma2_rn(w2, C[3][4], z0); z1 = __ffma2_rn(w2, C[3][5], z1);
          if (XW) for (int k = 4; k < a.nnz; ++k) {                 // rare: more than 4 non-zero weights (uncached)
            const int64_t xi = (int64_t)(k - 4) * a.n_pad_tc + vbase + g * 4 + u;
            const uint32_t xo = __ldg(a.tc_xoff + xi);
            const float xw = __ldg(a.tc_xw + xi);
            const float4 q0 = lds128(my_tab + xo), q1 = lds128(my_tab + xo + 16u), q2 = lds128(my_tab + xo + 32u);
            w2 = make_float2(xw, xw);
            c0 = __ffma2_rn(w2, make_float2(q0.x, q0.y), c0); c1 = __ffma2_rn(w2, make_float2(q0.z, q0.w), c1);
            c2 = __ffma2_rn(w2, make_float2(q1.x, q1.y), c2); c3 = __ffma2_rn(w2, make_float2(q1.z, q1.w), c3);
            z0 = __ffma2_rn(w2, make_float2(q2.x, q2.y), z0); z1 = __ffma2_rn(w2, make_float2(q2.z, q2.w), z1);
          }
          const float px = r0.x + ax[u], py = r0.y + ay[u], pz = r0.z + az[u];
          float2 xy = __ffma2_rn(c0, make_float2(px, px), c3);
          xy = __ffma2_rn(c1, make_float2(py, py), xy);
          xy = __ffma2_rn(c2, make_float2(pz, pz), xy);
          ox[u] = xy.x; oy[u] = xy.y;
          oz[u] = fmaf(z1.x, pz, fmaf(z0.y, py, fmaf(z0.x, px, z1.y)));
          pv[u] = __float_as_int(r0.w);                             // original vertex id, -1 = padding / skipped vertex
        };
        const uint32_t gchg = (__float_as_uint(R2[0].x) | __float_as_uint(R2[1].x) | __float_as_uint(R2[2].x) |
                               __float_as_uint(R2[3].x)) >> 24;
        if (gchg == 0) {                                            // common: the 4 vertices keep all four cached slots
#pragma unroll
          for (int u = 0; u < 4; ++u) blend_apply(u, R0[u], R1[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 r2 = R2[u];
            const uint32_t chg = __float_as_uint(r2.x) >> 24;       // warp-uniform: the record is per vertex
            const uint32_t off[4] = {__float_as_uint(r2.x) & 0xffffffu, __float_as_uint(r2.y), __float_as_uint(r2.z), __float_as_uint(r2.w)};
            if (chg) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (chg & (1u << k)) {
                  const float4 q0 = lds128(my_tab + off[k]), q1 = lds128(my_tab + off[k] + 16u), q2 = lds128(my_tab + off[k] + 32u);
                  C[k][0] = make_float2(q0.x, q0.y); C[k][1] = make_float2(q0.z, q0.w);
                  C[k][2] = make_float2(q1.x, q1.y); C[k][3] = make_float2(q1.z, q1.w);
                  C[k][4] = make_float2(q2.x, q2.y); C[k][5] = make_float2(q2.z, q2.w);
                }
              }
            }
            blend_apply(u, R0[u], R1[u]);
          }
        }
        if (FUSE_SDF) {                                  // transforms are world-composed: (ox,oy,oz) is the world point
          if (smem_mask) {
            // level 1: one FMA + one saturating convert per axis straight to the 8^3-cell index, sign bit from smem
            uint32_t m1 = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t u0 = __float_as_uint(fmaf(__saturatef(fmaf(ox[u], na0, nb0)), nn0, kMagic));
              const uint32_t u1 = __float_as_uint(fmaf(__saturatef(fmaf(oy[u], na1, nb1)), nn1, kMagic));
              const uint32_t u2 = __float_as_uint(fmaf(__saturatef(fmaf(oz[u], na2, nb2)), nn2, kMagic));
              const uint32_t ci = (u0 * (uint32_t)a.sdf.C1 + u1) * (uint32_t)a.sdf.C2 + u2 - kc;
              const uint32_t bit = (lds32(mask_u32 + (ci >> 5) * 4u) >> (ci & 31u)) & 1u;
              m1 |= ((pv[u] >= 0 && n_ok) ? bit : 0u) << u;
            }
            if (__any_sync(0xffffffffu, m1 != 0u)) {     // queue the flagged (vertex, body) pairs; resolve them 32 at a time
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const bool f = (m1 >> u) & 1u;
                const unsigned m = __ballot_sync(0xffffffffu, f);
                if (m) {
                  if (f) sts128(qbase + (uint32_t)(n1 + __popc(m & lt_mask)) * 16u, make_float4(ox[u], oy[u], oz[u], n_bits));
                  n1 += __popc(m);
                  while (n1 >= 32) drain1_batch();
                }
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (pv[u] >= 0 && sdf_coarse_value(a.sdf, cx, cy, cz, sc, ox[u], oy[u], oz[u]) <= 0.0f)
                cnt += sdf_exact_negative(a.sdf.grid, a.sdf.D0, a.sdf.D1, a.sdf.D2, cx, cy, cz, sc, ox[u], oy[u], oz[u]) ? 1 : 0;
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (n_ok && a.out != nullptr && pv[u] >= 0) {
              float* o = a.out + ((int64_t)n * a.n_real + pv[u]) * 3;
              o[0] = __fadd_rn(ox[u], trx); o[1] = __fadd_rn(oy[u], try_); o[2] = __fadd_rn(oz[u], trz);
            }
          }
        }
      }
      PROF_ADD(16 + warp - 4);
#if EG_LBS_PROF
      if (tid == 128) { atomicAdd(&g_lbs_prof_vt[vt & 511], (unsigned long long)(clock64() - _pt0)); g_lbs_prof[blockIdx.x][30] += 1; }
#endif
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&tmem_empty[buf]); mbar_arrive(&tab_empty[tb]); }   // both buffers of this tile are free
      if (FUSE_SDF && n_ok && cnt > 0) atomicAdd(a.counts + n, cnt);
    }
    if (FUSE_SDF) {
      while (n1 > 0) drain1_batch();
      while (n2 > 0) drain2_batch();
    }
  }
#if EG_LBS_PROF
  if (lane == 0 && warp >= 4) g_lbs_prof[blockIdx.x][28] = max(g_lbs_prof[blockIdx.x][28], gtimer());   // (racy max: debug only)
  if (tid == 0) g_lbs_prof[blockIdx.x][29] = gtimer();
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                     // no CTA leaves while its peer may still multicast into it
  if (tid == 96) {                        // the scheduler lane: the last CTA to finish re-arms the ticket counter
    __threadfence();
    if (atomicInc(a.tile_sched + 1, gridDim.x - 1) == gridDim.x - 1) a.tile_sched[0] = 0u;
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// per-call copy of the tc vertex records with the caller's skip mask folded into the id field (-1 = do not count)
__global__ void tc_fold_skip_kernel(const float4* __restrict__ rec, const uint8_t* __restrict__ skip, int n_pad,
                                    float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 r0 = rec[i * 3];
  const int id = __float_as_int(r0.w);
  if (id >= 0 && skip != nullptr && skip[id] != 0) r0.w = __int_as_float(-1);
  out[i * 3] = r0; out[i * 3 + 1] = rec[i * 3 + 1]; out[i * 3 + 2] = rec[i * 3 + 2];
}

// --------------------------------------------------------------------------------------------
// Backward of the marker positions w.r.t. the body parameters (the loss of the reference's
// GAMMARegressorTrainOP / ComboTrainOP differentiates through SMPL-X, models_GAMMA_primitive.py:616-631):
//   markers[n,m] = sum_k w_mk A_jk [p_m; 1] + transl,   p_m = t_m + S_m beta + P_m F(theta),   A_j from the kinematic chain
// One CTA per body re-derives the forward quantities (rotations, chain, relative transforms) in shared memory, then
//   dA_j += w g (x) [p;1],  dp = sum_k w_k R(A_jk)^T g,  dF = P^T dp  ->  dR_j (pose blend path),
//   dG from dA, the chain walked from the leaves to the root,  dR_j -> dtheta_j through smplx's Rodrigues
//   (angle = |theta + 1e-8|), hand PCA transposed, d transl = sum_m g.
// betas are data in that loss (no gradient). Accumulation is fp32 with shared-memory atomics.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lbs_markers_backward_kernel(const float* __restrict__ xb, const float* __restrict__ betas, int betas_div, int N, int J, int S,
                            int n_levels, const float* __restrict__ hand_l, const float* __restrict__ hand_r,
                            const float* __restrict__ pose_mean, const float* __restrict__ Jt, const float* __restrict__ Js,
                            const int32_t* __restrict__ parents, const int32_t* __restrict__ level_joints,
                            const int32_t* __restrict__ level_start,
                            const float* __restrict__ basis /*[KPAD][3*n_pad]*/, const float* __restrict__ vt,
                            const int32_t* __restrict__ skin_idx, const float* __restrict__ skin_w, int n_pad, int nnz,
                            int n_markers, const float* __restrict__ d_markers, float* __restrict__ d_xb,
                            float* __restrict__ d_rot /*[N,22,9] dL/dR of the global + body joints, or null*/) {
  const int n = blockIdx.x, t = threadIdx.x, NT = blockDim.x;
  __shared__ float pose[MAXJ * 3], R[MAXJ][9], Jr[MAXJ][3], G[MAXJ][12], shape[32];
  __shared__ float dA[MAXJ][12], dG[MAXJ][12], dR[MAXJ][9], dth[MAXJ * 3], dtr[3];
  __shared__ float feat[KPAD], dF[KPAD];
  extern __shared__ float dyn[];                 // dp [n_markers][3]
  const float* x = xb + (int64_t)n * EG_XB_DIM;
  const float* be = betas + (int64_t)(n / betas_div) * 10;
  const int npose = (J - 1) * 9;
  // ---- forward re-derivation (same arithmetic as lbs_pose_prep_kernel) ----
  for (int i = t; i < J * 3; i += NT) {
    float v = 0.0f;
    if (i < 66) v = x[3 + i];
    else if (i >= 75 && i < 120) { for (int k = 0; k < 12; ++k) v += x[69 + k] * __ldg(hand_l + k * 45 + (i - 75)); }
    else if (i >= 120 && i < 165) { for (int k = 0; k < 12; ++k) v += x[81 + k] * __ldg(hand_r + k * 45 + (i - 120)); }
    pose[i] = v + __ldg(pose_mean + i);
  }
  if (t < 32) shape[t] = (t < 10) ? be[t] : 0.0f;
  for (int i = t; i < MAXJ * 12; i += NT) { (&dA[0][0])[i] = 0.0f; (&dG[0][0])[i] = 0.0f; }
  for (int i = t; i < MAXJ * 9; i += NT) (&dR[0][0])[i] = 0.0f;
  if (t < 3) dtr[t] = 0.0f;
  __syncthreads();
  if (t < J) {
    const float rx = pose[3 * t], ry = pose[3 * t + 1], rz = pose[3 * t + 2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float dx = rx / angle, dy = ry / angle, dz = rz / angle;
    const float sn = sinf(angle), cs = cosf(angle), omc = 1.0f - cs;
    const float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        float kk = 0.f;
        for (int m = 0; m < 3; ++m) kk += K[a * 3 + m] * K[m * 3 + b];
        R[t][a * 3 + b] = (a == b ? 1.0f : 0.0f) + sn * K[a * 3 + b] + omc * kk;
      }
    for (int c = 0; c < 3; ++c) {
      float v = __ldg(Jt + t * 3 + c);
      for (int k = 0; k < S; ++k) v += __ldg(Js + (t * 3 + c) * S + k) * shape[k];
      Jr[t][c] = v;
    }
  }
  __syncthreads();
  for (int k = t; k < KPAD; k += NT) {
    float v = 0.0f;
    if (k < npose) { const int j = 1 + k / 9, e = k % 9; v = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f); }
    else if (k < npose + S) v = shape[k - npose];
    feat[k] = v;
  }
  for (int lv = 0; lv < n_levels; ++lv) {
    const int beg = __ldg(level_start + lv), end = __ldg(level_start + lv + 1);
    for (int q = beg + t; q < end; q += NT) {
      const int j = __ldg(level_joints + q), p = __ldg(parents + j);
      if (p < 0) {
        for (int a = 0; a < 3; ++a) { G[j][a * 4] = R[j][a * 3]; G[j][a * 4 + 1] = R[j][a * 3 + 1]; G[j][a * 4 + 2] = R[j][a * 3 + 2]; G[j][a * 4 + 3] = Jr[j][a]; }
      } else {
        const float r0 = Jr[j][0] - Jr[p][0], r1 = Jr[j][1] - Jr[p][1], r2 = Jr[j][2] - Jr[p][2];
        for (int a = 0; a < 3; ++a) {
          const float g0 = G[p][a * 4], g1 = G[p][a * 4 + 1], g2 = G[p][a * 4 + 2], g3 = G[p][a * 4 + 3];
          for (int b = 0; b < 3; ++b) G[j][a * 4 + b] = g0 * R[j][b] + g1 * R[j][3 + b] + g2 * R[j][6 + b];
          G[j][a * 4 + 3] = g0 * r0 + g1 * r1 + g2 * r2 + g3;
        }
      }
    }
    __syncthreads();
  }
  // ---- markers: dA, dp ----
  const int64_t brow = (int64_t)n_pad * 3;
  for (int m = t; m < n_markers; m += NT) {
    float p[3];
    for (int c = 0; c < 3; ++c) {
      float v = vt[m * 3 + c];
      for (int k = 0; k < npose + S; ++k) v += basis[k * brow + m * 3 + c] * feat[k];
      p[c] = v;
    }
    const float* g = d_markers + ((int64_t)n * n_markers + m) * 3;
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    for (int k = 0; k < nnz; ++k) {
      const float w = skin_w[k * n_pad + m];
      if (w == 0.0f) continue;
      const int j = skin_idx[k * n_pad + m];
      const float gv[3] = {w * g0, w * g1, w * g2};
      for (int a = 0; a < 3; ++a) {
        atomicAdd(&dA[j][a * 4 + 0], gv[a] * p[0]); atomicAdd(&dA[j][a * 4 + 1], gv[a] * p[1]);
        atomicAdd(&dA[j][a * 4 + 2], gv[a] * p[2]); atomicAdd(&dA[j][a * 4 + 3], gv[a]);
      }
      // A_j rotation part = G_j rotation part
      dp0 += G[j][0] * gv[0] + G[j][4] * gv[1] + G[j][8] * gv[2];
      dp1 += G[j][1] * gv[0] + G[j][5] * gv[1] + G[j][9] * gv[2];
      dp2 += G[j][2] * gv[0] + G[j][6] * gv[1] + G[j][10] * gv[2];
    }
    dyn[m * 3] = dp0; dyn[m * 3 + 1] = dp1; dyn[m * 3 + 2] = dp2;
    atomicAdd(&dtr[0], g0); atomicAdd(&dtr[1], g1); atomicAdd(&dtr[2], g2);
  }
  __syncthreads();
  // ---- pose blend path: dF[k] = sum_m P[k][m] . dp[m]  ->  dR_j ----
  for (int k = t; k < npose; k += NT) {
    float v = 0.0f;
    for (int m = 0; m < n_markers; ++m)
      v += basis[k * brow + m * 3] * dyn[m * 3] + basis[k * brow + m * 3 + 1] * dyn[m * 3 + 1] + basis[k * brow + m * 3 + 2] * dyn[m * 3 + 2];
    dF[k] = v;
  }
  __syncthreads();
  for (int k = t; k < npose; k += NT) dR[1 + k / 9][k % 9] = dF[k];
  // ---- dA -> dG:  A_R = G_R,  A_t = G_t - G_R J ----
  if (t < J) {
    for (int a = 0; a < 3; ++a) {
      const float dat = dA[t][a * 4 + 3];
      for (int b = 0; b < 3; ++b) dG[t][a * 4 + b] = dA[t][a * 4 + b] - dat * Jr[t][b];
      dG[t][a * 4 + 3] = dat;
    }
  }
  __syncthreads();
  // ---- chain, leaves to root:  G_j,R = G_p,R R_j,  G_j,t = G_p,R r_j + G_p,t ----
  for (int lv = n_levels - 1; lv >= 0; --lv) {
    const int beg = __ldg(level_start + lv), end = __ldg(level_start + lv + 1);
    for (int q = beg + t; q < end; q += NT) {
      const int j = __ldg(level_joints + q), p = __ldg(parents + j);
      if (p < 0) {
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) atomicAdd(&dR[j][a * 3 + b], dG[j][a * 4 + b]);
      } else {
        const float r[3] = {Jr[j][0] - Jr[p][0], Jr[j][1] - Jr[p][1], Jr[j][2] - Jr[p][2]};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) {
            // dR_j[a][b] += sum_c G_p,R[c][a] dG_j,R[c][b]
            float v = 0.0f;
            for (int c = 0; c < 3; ++c) v += G[p][c * 4 + a] * dG[j][c * 4 + b];
            atomicAdd(&dR[j][a * 3 + b], v);
            // dG_p,R[a][b] += sum_c dG_j,R[a][c] R_j[b][c] + dG_j,t[a] r[b]
            float u = dG[j][a * 4 + 3] * r[b];
            for (int c = 0; c < 3; ++c) u += dG[j][a * 4 + c] * R[j][b * 3 + c];
            atomicAdd(&dG[p][a * 4 + b], u);
          }
        for (int a = 0; a < 3; ++a) atomicAdd(&dG[p][a * 4 + 3], dG[j][a * 4 + 3]);
      }
    }
    __syncthreads();
  }
  // rotation-matrix gradient of the 22 regressed joints, for callers whose pose parameterisation is not axis-angle
  if (d_rot != nullptr)
    for (int i = t; i < 22 * 9; i += NT) d_rot[(int64_t)n * 198 + i] = dR[i / 9][i % 9];
  // ---- Rodrigues backward (smplx batch_rodrigues: angle = |theta + 1e-8|, dir = theta / angle) ----
  if (t < J) {
    const float th[3] = {pose[3 * t], pose[3 * t + 1], pose[3 * t + 2]};
    const float av[3] = {th[0] + 1e-8f, th[1] + 1e-8f, th[2] + 1e-8f};
    const float angle = sqrtf(av[0] * av[0] + av[1] * av[1] + av[2] * av[2]);
    const float d[3] = {th[0] / angle, th[1] / angle, th[2] / angle};
    const float sn = sinf(angle), cs = cosf(angle), omc = 1.0f - cs;
    const float K[9] = {0.f, -d[2], d[1], d[2], 0.f, -d[0], -d[1], d[0], 0.f};
    float K2[9], dK[9];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { float v = 0.f; for (int m = 0; m < 3; ++m) v += K[a * 3 + m] * K[m * 3 + b]; K2[a * 3 + b] = v; }
    float dLs = 0.f, dLc = 0.f;
    for (int e = 0; e < 9; ++e) { dLs += dR[t][e] * K[e]; dLc -= dR[t][e] * K2[e]; }
    // dL/dK = s dR + (1 - c) (dR K^T + K^T dR)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        float v = sn * dR[t][a * 3 + b];
        for (int m = 0; m < 3; ++m) v += omc * (dR[t][a * 3 + m] * K[b * 3 + m] + K[m * 3 + a] * dR[t][m * 3 + b]);
        dK[a * 3 + b] = v;
      }
    const float dd[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    const float dang = dLs * cs - dLc * sn;
    const float inv = 1.0f / angle, inv3 = inv * inv * inv;
    const float dth_dot = dd[0] * th[0] + dd[1] * th[1] + dd[2] * th[2];
    for (int i = 0; i < 3; ++i) dth[3 * t + i] = dang * av[i] * inv + dd[i] * inv - dth_dot * av[i] * inv3;
  }
  __syncthreads();
  // ---- d xb: transl, global_orient + body_pose, hand PCA (jaw / eyes are constants of the parser) ----
  float* out = d_xb + (int64_t)n * EG_XB_DIM;
  if (t < 3) out[t] = dtr[t];
  for (int i = t; i < 66; i += NT) out[3 + i] = dth[i];
  for (int k = t; k < 24; k += NT) {
    const float* comp = k < 12 ? hand_l + k * 45 : hand_r + (k - 12) * 45;
    const int base = k < 12 ? 75 : 120;
    float v = 0.0f;
    for (int c = 0; c < 45; ++c) v += dth[base + c] * __ldg(comp + c);
    out[69 + k] = v;
  }
}

// joints [N,127,3] = posed joints ++ vertex joints ++ landmarks (+transl); markers [N,M,3]
__global__ void __launch_bounds__(128)
lbs_finish_kernel(const float* __restrict__ Jp, const float* __restrict__ cverts,
                  const float* __restrict__ xb, const float* __restrict__ lmk_bary, int N, int J,
                  int n_markers, int n_extra, int n_lmk, int nc, float* __restrict__ joints,
                  float* __restrict__ markers, const VertArgs ex, int exact_extra) {
  const int n = blockIdx.x;
  const float* x = xb + (int64_t)n * EG_XB_DIM;
  const float tx = x[0], ty = x[1], tz = x[2];
  const float tr[3] = {tx, ty, tz};
  const float* cv = cverts + (int64_t)n * nc * 3;
  const int n_out = J + n_extra + n_lmk;
  // exact_extra: the compact set came from the tcgen05 tiles (fp16-rounded pose rows, ~5e-6 m). The vertex joints feed
  // DIFFERENCES of nearby points (ego-sensing's look-at = j57 - j23 + j56 - j24 over a few cm), so they are re-evaluated
  // here in fp32: blend over the SIMT basis of the compact set, then the 4-weight skinning of vertex_epilogue.
  __shared__ float Fsh[KPAD];
  __shared__ float exj[32 * 3];
  if (exact_extra && joints) {
    for (int k = threadIdx.x; k < KPAD; k += blockDim.x) Fsh[k] = ex.Ft[(int64_t)k * ex.Npad + n];
    __syncthreads();
    const int ne = min(n_extra, 32);
    __shared__ float part[4][32 * 3];
    {   // warp w sums k = w, w + 4, ... for every output: lanes walk 3 * ne consecutive floats of one basis row
      const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
      float acc[3] = {0.f, 0.f, 0.f};
      const float* bp = ex.basis + (int64_t)n_markers * 3;
#pragma unroll 4
      for (int k = w; k < KPAD; k += 4) {
        const float f = Fsh[k];
        const float* row = bp + (int64_t)k * 3 * ex.n_pad;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int o = lane + 32 * q;
          if (o < ne * 3) acc[q] += f * __ldg(row + o);
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) if (lane + 32 * q < ne * 3) part[w][lane + 32 * q] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < ne * 3) {
      const int i = threadIdx.x / 3, c = threadIdx.x % 3, v = n_markers + i;
      exj[threadIdx.x] = ex.vt[v * 3 + c] + (((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x < ne * 3) {
      const int i = threadIdx.x / 3, c = threadIdx.x % 3, v = n_markers + i;
      const float px = exj[i * 3 + 0], py = exj[i * 3 + 1], pz = exj[i * 3 + 2];
      const float4* An = reinterpret_cast<const float4*>(ex.A + (int64_t)n * ex.J * 12);
      float T0 = 0.f, T1 = 0.f, T2 = 0.f, T3 = 0.f;
      for (int k = 0; k < ex.nnz; ++k) {
        const int j = ex.skin_idx[k * ex.n_pad + v];
        const float w = ex.skin_w[k * ex.n_pad + v];
        const float4 r = __ldg(An + j * 3 + c);
        T0 += w * r.x; T1 += w * r.y; T2 += w * r.z; T3 += w * r.w;
      }
      joints[((int64_t)n * n_out + J + i) * 3 + c] = __fadd_rn(T0 * px + T1 * py + T2 * pz + T3, tr[c]);
    }
  }
  if (joints) {
    for (int i = threadIdx.x; i < n_out * 3; i += blockDim.x) {
      const int j = i / 3, c = i % 3;
      float v;
      if (j < J) {
        v = Jp[((int64_t)n * J + j) * 3 + c];
      } else if (j < J + n_extra) {
        if (exact_extra && j - J < 32) continue;        // written above
        v = cv[(n_markers + (j - J)) * 3 + c];
      } else {
        const int l = j - J - n_extra;
        const float* q = cv + (n_markers + n_extra + 3 * l) * 3;
        // einsum('blfi,blf->bli'): sum over the 3 face corners in order
        v = q[0 + c] * __ldg(lmk_bary + 3 * l) + q[3 + c] * __ldg(lmk_bary + 3 * l + 1) +
            q[6 + c] * __ldg(lmk_bary + 3 * l + 2);
      }
      joints[((int64_t)n * n_out + j) * 3 + c] = __fadd_rn(v, tr[c]);
    }
  }
  if (markers) {
    for (int i = threadIdx.x; i < n_markers * 3; i += blockDim.x)
      markers[((int64_t)n * n_markers) * 3 + i] = __fadd_rn(cv[i], tr[i % 3]);
  }
}

__global__ void gather_vertex_set_kernel(const float* __restrict__ src_basis, int64_t src_row,
                                         const float* __restrict__ src_vt,
                                         const int32_t* __restrict__ src_idx,
                                         const float* __restrict__ src_w, int src_npad, int nnz,
                                         const int32_t* __restrict__ vids, int n, int n_pad,
                                         float* __restrict__ basis, float* __restrict__ vt,
                                         int32_t* __restrict__ idx, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = vids[i];
  for (int k = 0; k < KPAD; ++k)
    for (int c = 0; c < 3; ++c) basis[(int64_t)k * n_pad * 3 + i * 3 + c] = src_basis[k * src_row + v * 3 + c];
  for (int c = 0; c < 3; ++c) vt[i * 3 + c] = src_vt[v * 3 + c];
  for (int k = 0; k < nnz; ++k) {
    idx[k * n_pad + i] = src_idx[k * src_npad + v];
    w[k * n_pad + i] = src_w[k * src_npad + v];
  }
}

// calc_calibrate_offset (baseops.py:494-534): pelvis of the body with zero transl / global_orient.
// The root joint's posed position is its rest position, so this is J_0(betas) - no LBS pass needed.
__global__ void rest_pelvis_kernel(const float* __restrict__ betas, int betas_div, int N, int S,
                                   const float* __restrict__ Jt, const float* __restrict__ Js,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 3) return;
  const int n = i / 3, c = i % 3;
  const float* be = betas + (int64_t)(n / betas_div) * 10;
  float v = __ldg(Jt + c);
  for (int k = 0; k < S && k < 10; ++k) v += __ldg(Js + c * S + k) * be[k];
  out[i] = v;
}

static void free_vertex_set(VertexSet& s) {
  cudaFree(s.basis); cudaFree(s.basisT); cudaFree(s.vt); cudaFree(s.skin_idx); cudaFree(s.skin_w);
  cudaFree(s.tc_rec); cudaFree(s.tc_jl); cudaFree(s.tc_nj); cudaFree(s.tc_xoff); cudaFree(s.tc_xw);
  s = VertexSet();
}

static int ensure_workspace(EgLbs* h, int N) {
  if (N <= h->cap_N) return EG_OK;
  int cap = std::max(N, 64);
  cap = (cap + 31) / 32 * 32;
  cudaFree(h->Ft); cudaFree(h->A); cudaFree(h->Jp); cudaFree(h->cout_); cudaFree(h->Ftc); cudaFree(h->Aw); cudaFree(h->Al);
  h->Ft = h->A = h->Jp = h->cout_ = h->Aw = h->Al = nullptr;
  h->Ftc = nullptr;
  h->cap_N = 0;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Ft, (size_t)KPAD * cap * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(h->Ft, 0, (size_t)KPAD * cap * sizeof(float)));
  const size_t cap128 = ((size_t)cap + tc::CLUSTER * tc::TB - 1) / (tc::CLUSTER * tc::TB) * (tc::CLUSTER * tc::TB);   // whole body-tile pairs
  EG_CUDA_CHECK(cudaMalloc((void**)&h->A, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(h->A, 0, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Aw, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(h->Aw, 0, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Al, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(h->Al, 0, cap128 * h->J * 12 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Jp, (size_t)cap * h->J * 3 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->cout_, (size_t)cap * 512 * 3 * sizeof(float)));
  h->cap_Ntc = (int)cap128;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->Ftc, (size_t)h->cap_Ntc * tc::KT * sizeof(__half)));
  EG_CUDA_CHECK(cudaMemset(h->Ftc, 0, (size_t)h->cap_Ntc * tc::KT * sizeof(__half)));
  h->cap_N = cap;
  return EG_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp16 tensor [rows][tc::KT] (K contiguous), box = [64 k][box_rows], SWIZZLE_128B
static int encode_map(EgLbs* h, CUtensorMap* map, const __half* base, uint64_t rows, uint32_t box_rows) {
  if (!h->encode_fn) return set_error(EG_ERR_STATE, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[2] = {(cuuint64_t)tc::KT, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)tc::KT * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)tc::BKT, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)h->encode_fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims,
                                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return EG_OK;
}

static int run_forward(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                       float* verts, float* joints, float* markers, bool fuse, int frames_per_env,
                       const float* R0, const float* T0, SdfGrid sdf, const uint8_t* skip,
                       int32_t* counts, cudaStream_t st) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(N >= 0, "negative N");
  if (N == 0) return EG_OK;
  EG_REQUIRE(xb && betas, "null pointer");
  EG_REQUIRE(betas_rows >= 1 && N % betas_rows == 0, "betas_rows must divide N (row = body / (N / betas_rows))");
  const int betas_div = N / betas_rows;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = ensure_workspace(h, N);
  if (rc) return rc;
  const int Npad = h->cap_N;   // row stride of Ft (fixed per workspace so tiles never read OOB)
  const bool want_tc = h->use_tc && h->full.basisT != nullptr && (verts != nullptr || fuse);
  // the compact set (joints / markers) rides the same tcgen05 kernel when the features of this call are already in its layout
  // (opt-in, EG_LBS_COMPACT_TC=1: -0.28 ms per 256-env iteration, but the fp16-rounded pose rows leave ~5e-6 m on the vertex
  // joints, which ego-sensing's look-at differences amplify past its tolerance, and re-evaluating those joints in fp32 in
  // the finish kernel costs more than the SIMT pass it replaces - DESIGN.md section 4)
  static const bool compact_tc_on = [] { const char* e = getenv("EG_LBS_COMPACT_TC"); return e && e[0] == '1'; }();
  const bool compact_tc = compact_tc_on && want_tc && (joints != nullptr || markers != nullptr) && h->compact.basisT != nullptr &&
                          h->encode_fn != nullptr;
  EG_LAUNCH(lbs_pose_prep_kernel, N, 64, 0, st, xb, betas, betas_div, N, Npad, h->J, h->S,
            h->n_levels, h->hand_l, h->hand_r, h->pose_mean, h->Jt, h->Js, h->parents,
            h->level_joints, h->level_start, h->Ft, want_tc ? h->Ftc : nullptr, h->A, h->Jp, R0, T0,
            frames_per_env, want_tc ? h->Aw : nullptr, h->cap_Ntc, compact_tc && fuse ? h->Al : nullptr);
  VertArgs a{};
  a.Ft = h->Ft; a.A = h->A; a.xb = xb; a.N = N; a.Npad = Npad; a.J = h->J;
  const int by = (N + TILE_B - 1) / TILE_B;
  if (verts != nullptr || fuse) {
    const VertexSet& s = h->full;
    a.basis = s.basis; a.vt = s.vt; a.skin_idx = s.skin_idx; a.skin_w = s.skin_w;
    a.n_real = s.n; a.n_pad = s.n_pad; a.nnz = s.nnz; a.add_transl = 1; a.out = verts;
    dim3 grid(s.n_pad / TILE_V, by);
    if (fuse) {
      a.sdf = sdf; a.R0 = R0; a.T0 = T0; a.frames_per_env = frames_per_env; a.skip = skip;
      a.counts = counts;
      EG_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)N * sizeof(int32_t), st));
    }
    if (want_tc) {
      // B-operand tensor map over the feature matrix of this call (host-side encode, no device work)
      CUtensorMap mapB;
      int rc2 = encode_map(h, &mapB, h->Ftc, (uint64_t)h->cap_Ntc, tc::TB);
      if (rc2) return rc2;
      const int n_vt = s.n_vt_tc, n_bt = (N + tc::TB - 1) / tc::TB;
      a.tc_rec = s.tc_rec; a.tc_jl = s.tc_jl; a.tc_nj = s.tc_nj; a.tc_xoff = s.tc_xoff; a.tc_xw = s.tc_xw;
      a.n_pad_tc = s.n_pad_tc; a.A_rows = h->cap_Ntc;
      if (fuse && skip != nullptr) {               // fold the caller's skip mask into a per-call copy of the records
        EG_LAUNCH(tc_fold_skip_kernel, (s.n_pad_tc + 127) / 128, 128, 0, st, s.tc_rec, skip, s.n_pad_tc, h->rec_call);
        a.tc_rec = h->rec_call;
      }
      a.A = h->Aw;      // pair-layout transforms (world-composed when fused: the epilogue goes straight to the SDF sample)
      const int n_btp = (n_bt + tc::CLUSTER - 1) / tc::CLUSTER;
      const int grid_tc = tc::CLUSTER * std::min(n_vt * n_btp, h->max_clusters);
      prof_begin(st, N);
      a.tile_sched = h->tile_sched;
      const bool xw = s.nnz > 4;
      if (fuse && xw) EG_LAUNCH((lbs_verts_tc_kernel<true, true>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA, mapB, a, n_vt, n_bt);
      else if (fuse) EG_LAUNCH((lbs_verts_tc_kernel<true, false>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA, mapB, a, n_vt, n_bt);
      else if (xw) EG_LAUNCH((lbs_verts_tc_kernel<false, true>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA, mapB, a, n_vt, n_bt);
      else EG_LAUNCH((lbs_verts_tc_kernel<false, false>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA, mapB, a, n_vt, n_bt);
      prof_end(st);
    } else if (fuse) {
      prof_begin(st, N);
      EG_LAUNCH(lbs_verts_kernel<true>, grid, VERT_THREADS, VERT_SMEM, st, a);
      prof_end(st);
    } else {
      EG_LAUNCH(lbs_verts_kernel<false>, grid, VERT_THREADS, VERT_SMEM, st, a);
    }
  }
  if (joints != nullptr || markers != nullptr) {
    const VertexSet& s = h->compact;
    EG_REQUIRE(s.n > 0, "compact vertex set missing (eg_lbs_set_markers)");
    VertArgs c = a;
    c.basis = s.basis; c.vt = s.vt; c.skin_idx = s.skin_idx; c.skin_w = s.skin_w;
    c.n_real = s.n; c.n_pad = s.n_pad; c.nnz = s.nnz; c.add_transl = 0; c.out = h->cout_;
    c.counts = nullptr;
    c.A = h->A;                     // local-frame transforms (the fused path may have switched a.A to Aw)
    dim3 grid(s.n_pad / TILE_V, by);
    if (compact_tc) {
      CUtensorMap mapB;
      int rc2 = encode_map(h, &mapB, h->Ftc, (uint64_t)h->cap_Ntc, tc::TB);
      if (rc2) return rc2;
      const int n_vt = s.n_vt_tc, n_bt = (N + tc::TB - 1) / tc::TB;
      c.tc_rec = s.tc_rec; c.tc_jl = s.tc_jl; c.tc_nj = s.tc_nj; c.tc_xoff = s.tc_xoff; c.tc_xw = s.tc_xw;
      c.n_pad_tc = s.n_pad_tc; c.A_rows = h->cap_Ntc;
      c.A = fuse ? h->Al : h->Aw;   // joint-major local transforms
      c.tile_sched = h->tile_sched;
      const int grid_tc = std::min(n_vt * n_bt, h->max_clusters);
      if (s.nnz > 4) EG_LAUNCH((lbs_verts_tc_kernel<false, true>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA_c, mapB, c, n_vt, n_bt);
      else EG_LAUNCH((lbs_verts_tc_kernel<false, false>), grid_tc, tc::THREADS, tc::SMEM_BYTES, st, h->mapA_c, mapB, c, n_vt, n_bt);
    } else
    EG_LAUNCH(lbs_verts_kernel<false>, grid, VERT_THREADS, VERT_SMEM, st, c);
    c.A = h->A; c.Ft = h->Ft;      // fp32 operands of the exact vertex-joint pass
    EG_LAUNCH(lbs_finish_kernel, N, 128, 0, st, h->Jp, h->cout_, xb, h->lmk_bary, N, h->J,
              h->n_markers, h->n_extra, h->n_lmk, s.n, joints, markers, c, compact_tc ? 1 : 0);
  }
  return EG_OK;
}

// records / joint lists / extra-weight tables of vertex set `s` for the list `items` of mesh vertex ids (the record's id
// field is the INDEX into `items`, i.e. the output slot); perm[row] = item of tc row `row` or -1
static int build_tc_records(EgLbs* h, VertexSet& s, const std::vector<int32_t>& items, std::vector<int>& perm) {
  const int n_items = (int)items.size(), nnz = h->full.nnz, npf = h->full.n_pad;
  const std::vector<int32_t>& sidx = h->h_sidx;
  const std::vector<float>& sw = h->h_sw;
  auto jw = [&](int it, int k, int& j, float& w) { j = sidx[(size_t)k * npf + items[it]]; w = sw[(size_t)k * npf + items[it]]; };
  auto cnt = [&](int it) { int c = 0; for (int k = 0; k < nnz; ++k) c += sw[(size_t)k * npf + items[it]] != 0.0f; return c; };
  std::vector<int> order(n_items);
  for (int v = 0; v < n_items; ++v) order[v] = v;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    for (int k = 0; k < nnz; ++k) {
      int jx, jy; float wx, wy;
      jw(x, k, jx, wx); jw(y, k, jy, wy);
      if (wx == 0.0f) jx = 1 << 20;
      if (wy == 0.0f) jy = 1 << 20;
      if (jx != jy) return jx < jy;
    }
    return false;
  });
  // tiles
  std::vector<std::vector<int>> tiles, tile_joints;
  {
    std::vector<int> cur, joints;
    for (int v : order) {
      std::vector<int> add;
      for (int k = 0; k < cnt(v); ++k) {
        int j; float w; jw(v, k, j, w);
        if (std::find(joints.begin(), joints.end(), j) == joints.end()) add.push_back(j);
      }
      if ((int)cur.size() == tc::TV || (int)(joints.size() + add.size()) > tc::NJ_MAX) {
        tiles.push_back(cur); tile_joints.push_back(joints);
        cur.clear(); joints.clear(); add.clear();
        for (int k = 0; k < cnt(v); ++k) { int j; float w; jw(v, k, j, w); add.push_back(j); }
      }
      if ((int)add.size() > tc::NJ_MAX) return set_error(EG_ERR_INVALID_ARG, "a vertex has more skinning joints than the tensor-core tile table holds (9)");
      joints.insert(joints.end(), add.begin(), add.end());
      cur.push_back(v);
    }
    if (!cur.empty()) { tiles.push_back(cur); tile_joints.push_back(joints); }
  }
  const int n_vt = (int)tiles.size();
  const int n_pad = n_vt * tc::TV;
  const int nx = std::max(nnz - 4, 0);
  std::vector<float> rec((size_t)n_pad * 12, 0.0f);
  std::vector<int32_t> jl((size_t)n_vt * tc::NJ_MAX, 0), nj(n_vt, 0);
  std::vector<uint32_t> xoff((size_t)std::max(nx, 1) * n_pad, 0u);
  std::vector<float> xw((size_t)std::max(nx, 1) * n_pad, 0.0f);
  perm.assign(n_pad, -1);
  for (int t = 0; t < n_vt; ++t) {
    const std::vector<int>& J = tile_joints[t];
    nj[t] = (int)J.size();
    for (size_t i = 0; i < J.size(); ++i) jl[(size_t)t * tc::NJ_MAX + i] = J[i];
    auto local = [&](int j) { return (uint32_t)(std::find(J.begin(), J.end(), j) - J.begin()); };
    int prev[4] = {J.empty() ? 0 : J[0], J.empty() ? 0 : J[0], J.empty() ? 0 : J[0], J.empty() ? 0 : J[0]};
    for (int i = 0; i < tc::TV; ++i) {
      float* r = &rec[((size_t)t * tc::TV + i) * 12];
      int slot_j[4] = {prev[0], prev[1], prev[2], prev[3]};
      float slot_w[4] = {0.f, 0.f, 0.f, 0.f};
      int id = -1;
      if (i < (int)tiles[t].size()) {
        const int v = id = tiles[t][i];
        perm[(size_t)t * tc::TV + i] = v;
        const int c = cnt(v);
        bool taken[4] = {false, false, false, false};
        std::vector<std::pair<int, float>> rest;
        for (int k = 0; k < c; ++k) {
          int j; float w; jw(v, k, j, w);
          int hit = -1;
          for (int q = 0; q < 4; ++q) if (!taken[q] && prev[q] == j) { hit = q; break; }
          if (hit >= 0) { taken[hit] = true; slot_j[hit] = j; slot_w[hit] = w; }
          else rest.push_back({j, w});
        }
        int xk = 0;
        for (auto& jwv : rest) {
          int q = 0;
          while (q < 4 && taken[q]) ++q;
          if (q < 4) { taken[q] = true; slot_j[q] = jwv.first; slot_w[q] = jwv.second; }
          else {                                   // beyond 4 entries: uncached extras
            xoff[(size_t)xk * n_pad + (size_t)t * tc::TV + i] = local(jwv.first) * (uint32_t)tc::SLOT_BYTES;
            xw[(size_t)xk * n_pad + (size_t)t * tc::TV + i] = jwv.second;
            ++xk;
          }
        }
        const float* vtp = &h->h_vt[(size_t)items[v] * 3];
        r[0] = vtp[0]; r[1] = vtp[1]; r[2] = vtp[2];
      }
      memcpy(&r[3], &id, 4);
      // bit 24+q of the first offset word: slot q differs from the previous vertex of the same epilogue warp's
      // range (every range of VPW vertices starts with all four set) -> the epilogue reloads exactly those slots
      uint32_t chg = (i % tc::VPW == 0) ? 0xFu : 0u;
      for (int q = 0; q < 4; ++q) {
        r[4 + q] = slot_w[q];
        if (slot_j[q] != prev[q]) chg |= 1u << q;
        prev[q] = slot_j[q];
      }
      for (int q = 0; q < 4; ++q) {
        uint32_t o = local(slot_j[q]) * (uint32_t)tc::SLOT_BYTES;
        if (q == 0) o |= chg << 24;
        memcpy(&r[8 + q], &o, 4);
      }
    }
  }
  s.n_pad_tc = n_pad; s.n_vt_tc = n_vt;
  int rc = dev_alloc_copy(reinterpret_cast<float**>(&s.tc_rec), rec.data(), rec.size());
  rc |= dev_alloc_copy(&s.tc_jl, jl.data(), jl.size());
  rc |= dev_alloc_copy(&s.tc_nj, nj.data(), nj.size());
  rc |= dev_alloc_copy(&s.tc_xoff, xoff.data(), xoff.size());
  rc |= dev_alloc_copy(&s.tc_xw, xw.data(), xw.size());
  return rc;
}

// Joint-coherent layout of the full mesh for the tcgen05 kernel (lbs_tc.cuh): vertices sorted by their tuple of
// skinning joints, cut into tiles of <= 80 vertices touching <= NJ_MAX distinct joints; inside a tile every vertex
// keeps as many of its (up to 4) register slots as possible on the joint the previous vertex had there, so the
// epilogue's per-body register cache is reloaded only where the joint really changes.
static int build_tc_layout(EgLbs* h, const EgLbsModel* m) {
  VertexSet& s = h->full;
  const int V = s.n, P = h->P, S = h->S;
  std::vector<int32_t> items(V);
  for (int v = 0; v < V; ++v) items[v] = v;
  std::vector<int> perm;
  int rc = build_tc_records(h, s, items, perm);
  if (rc) return rc;
  const int n_pad = s.n_pad_tc;
  h->h_row_of_vertex.assign(V, 0);
  for (int row = 0; row < n_pad; ++row) if (perm[row] >= 0) h->h_row_of_vertex[perm[row]] = row;
  // planar K-major fp16 copy of the basis in tc order: basisT[c][row][k]; shape rows in hi/hi, hi(again), lo form
  auto rnd = [](float x) { return __half2float(__float2half_rn(x)); };
  std::vector<__half> bt((size_t)3 * n_pad * tc::KT, __float2half_rn(0.0f));
  for (int c = 0; c < 3; ++c)
    for (int row = 0; row < n_pad; ++row) {
      const int v = perm[row];
      if (v < 0) continue;
      __half* dst = &bt[((size_t)c * n_pad + row) * tc::KT];
      for (int k = 0; k < P; ++k) dst[k] = __float2half_rn(m->posedirs[(size_t)k * V * 3 + (size_t)v * 3 + c]);
      for (int k = 0; k < S; ++k) {
        const float p = m->shapedirs[((size_t)v * 3 + c) * S + k];
        const float hi = rnd(p);
        dst[P + k] = __float2half_rn(hi);            // x shape_hi
        dst[P + S + k] = __float2half_rn(hi);        // x shape_lo
        dst[P + 2 * S + k] = __float2half_rn(p - hi);   // x shape_hi
      }
    }
  rc = dev_alloc_copy(&s.basisT, bt.data(), bt.size());
  {
    std::vector<float> rec((size_t)n_pad * 12);
    if (!rc && cudaMemcpy(rec.data(), s.tc_rec, rec.size() * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) rc = EG_ERR_CUDA;
    rc |= dev_alloc_copy(reinterpret_cast<float**>(&h->rec_call), rec.data(), rec.size());
  }
  if (!rc && h->encode_fn) rc |= encode_map(h, &h->mapA, s.basisT, (uint64_t)3 * n_pad, tc::HALF_ROWS);
  return rc;
}

// rows of the compact set's K-major basis are rows of the full set's (same fp16 values): dst[c][r] = src[c][src_row[r]]
__global__ void gather_basisT_kernel(const __half* __restrict__ src, int src_pad, const int32_t* __restrict__ src_row,
                                     int dst_pad, __half* __restrict__ dst) {
  const int r = blockIdx.x, c = blockIdx.y;
  const int sr = src_row[r];
  if (sr < 0) return;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + ((size_t)c * src_pad + sr) * tc::KT);
  uint4* d4 = reinterpret_cast<uint4*>(dst + ((size_t)c * dst_pad + r) * tc::KT);
  for (int i = threadIdx.x; i < tc::KT * 2 / 16; i += blockDim.x) d4[i] = s4[i];
}

// tensor-core layout of the compact set (markers + vertex joints + landmark corners): the 20-frame env pass evaluates
// it with the same tcgen05 kernel right after the full mesh instead of a second SIMT pass
static int build_tc_compact(EgLbs* h, VertexSet& s, const std::vector<int32_t>& vids) {
  if (h->full.basisT == nullptr || vids.empty()) return EG_OK;
  std::vector<int> perm;
  int rc = build_tc_records(h, s, vids, perm);
  if (rc) return rc;
  const int n_pad = s.n_pad_tc;
  std::vector<int32_t> src_row(n_pad, -1);
  for (int r = 0; r < n_pad; ++r) if (perm[r] >= 0) src_row[r] = h->h_row_of_vertex[vids[perm[r]]];
  int32_t* d_rows = nullptr;
  rc = dev_alloc_copy(&d_rows, src_row.data(), src_row.size());
  if (rc) return rc;
  EG_CUDA_CHECK(cudaMalloc((void**)&s.basisT, (size_t)3 * n_pad * tc::KT * sizeof(__half)));
  EG_CUDA_CHECK(cudaMemset(s.basisT, 0, (size_t)3 * n_pad * tc::KT * sizeof(__half)));
  EG_LAUNCH(gather_basisT_kernel, dim3(n_pad, 3), 64, 0, 0, h->full.basisT, h->full.n_pad_tc, d_rows, n_pad, s.basisT);
  EG_CUDA_CHECK(cudaDeviceSynchronize());
  cudaFree(d_rows);
  if (h->encode_fn) rc = encode_map(h, &h->mapA_c, s.basisT, (uint64_t)3 * n_pad, tc::HALF_ROWS);
  return rc;
}

static int build_compact(EgLbs* h, const int32_t* marker_vids, int n_markers) {
  std::vector<int32_t> vids(marker_vids, marker_vids + n_markers);
  vids.insert(vids.end(), h->h_extra.begin(), h->h_extra.end());
  vids.insert(vids.end(), h->h_lmk_verts.begin(), h->h_lmk_verts.end());
  const int n = (int)vids.size();
  if (n > 512) return set_error(EG_ERR_INVALID_ARG, "compact vertex set larger than 512 vertices");
  for (int v : vids)
    if (v < 0 || v >= h->V) return set_error(EG_ERR_INVALID_ARG, "vertex id out of range");
  free_vertex_set(h->compact);
  VertexSet s;
  s.n = n; s.n_pad = (n + TILE_V - 1) / TILE_V * TILE_V; s.nnz = h->full.nnz;
  EG_CUDA_CHECK(cudaMalloc((void**)&s.basis, (size_t)KPAD * s.n_pad * 3 * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(s.basis, 0, (size_t)KPAD * s.n_pad * 3 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&s.vt, (size_t)s.n_pad * 3 * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(s.vt, 0, (size_t)s.n_pad * 3 * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&s.skin_idx, (size_t)s.nnz * s.n_pad * sizeof(int32_t)));
  EG_CUDA_CHECK(cudaMemset(s.skin_idx, 0, (size_t)s.nnz * s.n_pad * sizeof(int32_t)));
  EG_CUDA_CHECK(cudaMalloc((void**)&s.skin_w, (size_t)s.nnz * s.n_pad * sizeof(float)));
  EG_CUDA_CHECK(cudaMemset(s.skin_w, 0, (size_t)s.nnz * s.n_pad * sizeof(float)));
  int32_t* d_vids = nullptr;
  int rc = dev_alloc_copy(&d_vids, vids.data(), vids.size());
  if (rc) return rc;
  EG_LAUNCH(gather_vertex_set_kernel, (n + 63) / 64, 64, 0, 0, h->full.basis,
            (int64_t)h->full.n_pad * 3, h->full.vt, h->full.skin_idx, h->full.skin_w, h->full.n_pad,
            s.nnz, d_vids, n, s.n_pad, s.basis, s.vt, s.skin_idx, s.skin_w);
  EG_CUDA_CHECK(cudaDeviceSynchronize());
  cudaFree(d_vids);
  rc = build_tc_compact(h, s, vids);
  if (rc) { free_vertex_set(s); return rc; }
  h->compact = s;
  h->n_markers = n_markers;
  return EG_OK;
}

}  // namespace eg

using namespace eg;

extern "C" int eg_lbs_create(const EgLbsModel* m, int device, EgLbs** out) {
  EG_REQUIRE(m && out, "null pointer");
  EG_REQUIRE(m->n_joints > 0 && m->n_joints <= MAXJ, "n_joints must be in (0,64]");
  EG_REQUIRE((m->n_joints - 1) * 9 == m->n_pose_basis, "n_pose_basis must be 9*(n_joints-1)");
  EG_REQUIRE(m->n_pose_basis + m->n_shape <= KPAD && m->n_shape <= 32, "basis too large");
  EG_REQUIRE(m->n_pose_basis + 3 * m->n_shape <= tc::KT, "basis too large for the tensor-core layout");
  EG_REQUIRE(m->n_hand_pca == 12 && m->n_joints == 55, "SMPL-X layout expected (55 joints, 12 hand PCA)");
  EG_CUDA_CHECK(cudaSetDevice(device));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VERT_SMEM));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VERT_SMEM));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  EG_CUDA_CHECK(cudaFuncSetAttribute(lbs_verts_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  EgLbs* h = new EgLbs();
  h->device = device;
  {
    // persistent kernel: the grid must not exceed what is co-resident (a GPC with an odd SM count strands one SM)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs / tc::CLUSTER * tc::CLUSTER); cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = tc::CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, lbs_verts_tc_kernel<true, true>, &cfg) == cudaSuccess && nc > 0)
      h->max_clusters = std::min(nc, kNumSMs / tc::CLUSTER);
    else {
      cudaGetLastError();
      h->max_clusters = tc::CLUSTER == 1 ? kNumSMs : 64;
    }
    EG_CUDA_CHECK(cudaMalloc((void**)&h->tile_sched, 2 * sizeof(unsigned int)));
    EG_CUDA_CHECK(cudaMemset(h->tile_sched, 0, 2 * sizeof(unsigned int)));
  }
  {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      h->encode_fn = fn;
  }
  const int V = h->V = m->n_verts, J = h->J = m->n_joints, S = h->S = m->n_shape, P = h->P = m->n_pose_basis;
  h->n_extra = m->n_extra; h->n_lmk = m->n_landmarks;
  h->h_extra.assign(m->extra_vids, m->extra_vids + m->n_extra);
  for (int l = 0; l < m->n_landmarks; ++l) {
    const int f = m->lmk_faces_idx[l];
    if (f < 0 || f >= m->n_faces) { delete h; return set_error(EG_ERR_INVALID_ARG, "landmark face id out of range"); }
    for (int c = 0; c < 3; ++c) h->h_lmk_verts.push_back(m->faces[3 * f + c]);
  }
  // ---- full vertex set ----
  VertexSet& s = h->full;
  s.n = V; s.n_pad = (V + TILE_V - 1) / TILE_V * TILE_V;
  const size_t row = (size_t)s.n_pad * 3;
  std::vector<float> basis((size_t)KPAD * row, 0.0f);
  for (int k = 0; k < P; ++k) memcpy(&basis[k * row], m->posedirs + (size_t)k * V * 3, (size_t)V * 3 * sizeof(float));
  for (int k = 0; k < S; ++k)
    for (size_t c = 0; c < (size_t)V * 3; ++c) basis[(size_t)(P + k) * row + c] = m->shapedirs[c * S + k];
  std::vector<float> vt(row, 0.0f);
  memcpy(vt.data(), m->v_template, (size_t)V * 3 * sizeof(float));
  int nnz = 1;
  for (int v = 0; v < V; ++v) {
    int c = 0;
    for (int j = 0; j < J; ++j) c += m->lbs_weights[(size_t)v * J + j] != 0.0f;
    nnz = std::max(nnz, c);
  }
  s.nnz = nnz;
  std::vector<int32_t> sidx((size_t)nnz * s.n_pad, 0);
  std::vector<float> sw((size_t)nnz * s.n_pad, 0.0f);
  for (int v = 0; v < V; ++v) {
    int c = 0;
    for (int j = 0; j < J; ++j) {                 // ascending joint order = dense matmul order
      const float w = m->lbs_weights[(size_t)v * J + j];
      if (w != 0.0f) { sidx[(size_t)c * s.n_pad + v] = j; sw[(size_t)c * s.n_pad + v] = w; ++c; }
    }
  }
  // ---- joint regressor folded through template / shapedirs (double accumulation on host) ----
  std::vector<float> Jt((size_t)J * 3), Js((size_t)J * 3 * S);
  for (int j = 0; j < J; ++j) {
    double t[3] = {0, 0, 0};
    std::vector<double> sd((size_t)3 * S, 0.0);
    for (int v = 0; v < V; ++v) {
      const double w = m->J_regressor[(size_t)j * V + v];
      if (w == 0.0) continue;
      for (int c = 0; c < 3; ++c) {
        t[c] += w * m->v_template[(size_t)v * 3 + c];
        for (int k = 0; k < S; ++k) sd[c * S + k] += w * m->shapedirs[((size_t)v * 3 + c) * S + k];
      }
    }
    for (int c = 0; c < 3; ++c) {
      Jt[j * 3 + c] = (float)t[c];
      for (int k = 0; k < S; ++k) Js[((size_t)j * 3 + c) * S + k] = (float)sd[c * S + k];
    }
  }
  // ---- tree levels ----
  std::vector<int> depth(J, 0);
  int maxd = 0;
  for (int j = 0; j < J; ++j) {
    const int p = m->parents[j];
    if (p >= j) { delete h; return set_error(EG_ERR_INVALID_ARG, "parents must precede children"); }
    depth[j] = p < 0 ? 0 : depth[p] + 1;
    maxd = std::max(maxd, depth[j]);
  }
  std::vector<int32_t> lj, ls;
  for (int d = 0; d <= maxd; ++d) {
    ls.push_back((int32_t)lj.size());
    for (int j = 0; j < J; ++j) if (depth[j] == d) lj.push_back(j);
  }
  ls.push_back((int32_t)lj.size());
  h->n_levels = maxd + 1;
  int rc = 0;
  h->h_sidx = sidx; h->h_sw = sw;
  h->h_vt.assign(m->v_template, m->v_template + (size_t)V * 3);
  rc |= build_tc_layout(h, m);
  rc |= dev_alloc_copy(&s.basis, basis.data(), basis.size());
  rc |= dev_alloc_copy(&s.vt, vt.data(), vt.size());
  rc |= dev_alloc_copy(&s.skin_idx, sidx.data(), sidx.size());
  rc |= dev_alloc_copy(&s.skin_w, sw.data(), sw.size());
  rc |= dev_alloc_copy(&h->Jt, Jt.data(), Jt.size());
  rc |= dev_alloc_copy(&h->Js, Js.data(), Js.size());
  rc |= dev_alloc_copy(&h->hand_l, m->hand_comp_l, (size_t)12 * 45);
  rc |= dev_alloc_copy(&h->hand_r, m->hand_comp_r, (size_t)12 * 45);
  rc |= dev_alloc_copy(&h->pose_mean, m->pose_mean, (size_t)J * 3);
  rc |= dev_alloc_copy(&h->parents, m->parents, (size_t)J);
  rc |= dev_alloc_copy(&h->level_joints, lj.data(), lj.size());
  rc |= dev_alloc_copy(&h->level_start, ls.data(), ls.size());
  rc |= dev_alloc_copy(&h->lmk_bary, m->lmk_bary, (size_t)m->n_landmarks * 3);
  if (rc) { eg_lbs_destroy(h); return EG_ERR_CUDA; }
  rc = build_compact(h, nullptr, 0);   // vertex joints + landmarks only, until markers are set
  if (rc) { eg_lbs_destroy(h); return rc; }
  *out = h;
  return EG_OK;
}

extern "C" void eg_lbs_destroy(EgLbs* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  free_vertex_set(h->full);
  free_vertex_set(h->compact);
  cudaFree(h->Jt); cudaFree(h->Js); cudaFree(h->hand_l); cudaFree(h->hand_r); cudaFree(h->pose_mean);
  cudaFree(h->parents); cudaFree(h->level_joints); cudaFree(h->level_start); cudaFree(h->lmk_bary);
  cudaFree(h->Ft); cudaFree(h->A); cudaFree(h->Jp); cudaFree(h->cout_); cudaFree(h->Ftc); cudaFree(h->Aw); cudaFree(h->rec_call);
  cudaFree(h->tile_sched); cudaFree(h->Al);
  delete h;
}

extern "C" int eg_lbs_set_markers(EgLbs* h, const int32_t* marker_vids_host, int n_markers) {
  EG_REQUIRE(h && (marker_vids_host || n_markers == 0) && n_markers >= 0, "bad arguments");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  return build_compact(h, marker_vids_host, n_markers);
}

extern "C" int eg_lbs_max_skin_nnz(const EgLbs* h) { return h ? h->full.nnz : 0; }

extern "C" int eg_lbs_forward(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                              float* verts, float* joints, float* markers, void* stream) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(markers == nullptr || h->n_markers > 0, "markers requested but none set");
  return run_forward(h, xb, betas, betas_rows, N, verts, joints, markers, false, 1, nullptr, nullptr,
                     SdfGrid{}, nullptr, nullptr, as_stream(stream));
}

extern "C" int eg_lbs_forward_sdf(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                                  int frames_per_env, const float* R0, const float* T0,
                                  const float* grid, int D0, int D1, int D2, const float* center_dev,
                                  const float* scale_dev, const uint8_t* skip_mask, int32_t* counts,
                                  float* joints, float* markers, void* stream) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(R0 && T0 && grid && center_dev && scale_dev && counts, "null pointer");
  EG_REQUIRE(frames_per_env > 0 && N % frames_per_env == 0, "N must be a multiple of frames_per_env");
  EG_REQUIRE(markers == nullptr || h->n_markers > 0, "markers requested but none set");
  SdfGrid g{grid, D0, D1, D2, center_dev, scale_dev};
  sdf_attach_coarse(g);     // sign-only early-out grid, if eg_sdf_prepare was called for this grid
  return run_forward(h, xb, betas, betas_rows, N, nullptr, joints, markers, true, frames_per_env, R0,
                     T0, g, skip_mask, counts, as_stream(stream));
}

extern "C" int eg_lbs_rest_pelvis(EgLbs* h, const float* betas, int betas_rows, int N, float* out, void* stream) {
  EG_REQUIRE(h && betas && out && N >= 0, "bad arguments");
  if (N == 0) return EG_OK;
  EG_REQUIRE(betas_rows >= 1 && N % betas_rows == 0, "betas_rows must divide N");
  EG_LAUNCH(rest_pelvis_kernel, (N * 3 + 127) / 128, 128, 0, as_stream(stream), betas, N / betas_rows, N, h->S,
            h->Jt, h->Js, out);
  return EG_OK;
}

extern "C" int eg_lbs_markers_backward(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                                       const float* d_markers, float* d_xb, void* stream) {
  return eg_lbs_markers_backward_rot(h, xb, betas, betas_rows, N, d_markers, d_xb, nullptr, stream);
}

extern "C" int eg_lbs_markers_backward_rot(EgLbs* h, const float* xb, const float* betas, int betas_rows, int N,
                                           const float* d_markers, float* d_xb, float* d_rot, void* stream) {
  EG_REQUIRE(h && xb && betas && d_markers && d_xb && N >= 0, "bad arguments");
  if (N == 0) return EG_OK;
  EG_REQUIRE(h->n_markers > 0, "no marker set (eg_lbs_set_markers)");
  EG_REQUIRE(betas_rows >= 1 && N % betas_rows == 0, "betas_rows must divide N");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  const VertexSet& s = h->compact;
  EG_LAUNCH(lbs_markers_backward_kernel, N, 128, (size_t)h->n_markers * 3 * sizeof(float), as_stream(stream), xb, betas,
            N / betas_rows, N, h->J, h->S, h->n_levels, h->hand_l, h->hand_r, h->pose_mean, h->Jt, h->Js, h->parents,
            h->level_joints, h->level_start, s.basis, s.vt, s.skin_idx, s.skin_w, s.n_pad, s.nnz, h->n_markers, d_markers,
            d_xb, d_rot);
  return EG_OK;
}

extern "C" int eg_lbs_set_mainloop(EgLbs* h, int use_tcgen05) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(!use_tcgen05 || (h->encode_fn && h->full.basisT), "tcgen05 path unavailable (no driver entry point)");
  h->use_tc = use_tcgen05 ? 1 : 0;
  return EG_OK;
}

#if EG_LBS_PROF
extern "C" int eg_lbs_prof_dump(unsigned long long* host_cta, unsigned long long* host_vt, int zero) {
  using namespace eg;
  cudaDeviceSynchronize();
  if (host_cta) cudaMemcpyFromSymbol(host_cta, g_lbs_prof, sizeof(g_lbs_prof));
  if (host_vt) cudaMemcpyFromSymbol(host_vt, g_lbs_prof_vt, sizeof(g_lbs_prof_vt));
  if (zero) {
    void* p; cudaGetSymbolAddress(&p, g_lbs_prof); cudaMemset(p, 0, sizeof(g_lbs_prof));
    cudaGetSymbolAddress(&p, g_lbs_prof_vt); cudaMemset(p, 0, sizeof(g_lbs_prof_vt));
  }
  return 0;
}
#endif
