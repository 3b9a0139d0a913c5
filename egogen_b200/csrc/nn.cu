// Dense-layer kernels (fp32 SIMT GEMM with fused bias/activation/residual epilogue, GRU gates,
// BatchNorm-eval) and the two inference-only network operators of the env step:
//   * GAMMAPrimitiveCombo.sample_prior  (reference motion/models/models_GAMMA_primitive.py:334-360,
//     predictor decode :83-101, regressor :222-301, 6-D -> axis-angle :208-219 + baseops.py:120-162)
//   * VPoser v1 encoder `.loc`         (reference call site crowd_env_2f.py:197-200)
#include <cooperative_groups.h>
#include <vector>

#include "geom.cuh"
#include <stdlib.h>

#include "nn.cuh"
#include "motion_tc.cuh"

namespace eg {

constexpr int BK = 16;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_TANH: return tanhf(v);
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LRELU: return v > 0.0f ? v : v * slope;
    default: return v;
  }
}

// fp32 SIMT GEMM, BM x BN x BKT tiles, TM x TN outputs per thread, 256 threads, double-buffered shared
// memory with register prefetch of the next k-tile (one __syncthreads per k-tile). Operands whose K dimension is
// contiguous and 16-byte aligned are fetched with 128-bit loads (VEC). launch_gemm picks the largest tile shape
// that still fills the 148 SMs.
template <int BM, int BN, int BKT, int TM, int TN, bool TA, bool TB, bool VEC, int SPLIT = 1>
__global__ void __launch_bounds__(256)
gemm_kernel(const GemmArgs g) {
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "tile / thread mismatch");
  constexpr bool VA = VEC && !TA, VB = VEC && TB;           // K-contiguous operands only
  constexpr int LA = BM * BKT / NT, LB = BN * BKT / NT;      // elements each thread loads per tile
  static_assert(LA % 4 == 0 || !VA, "vector A load needs 4 elements per thread");
  static_assert(LB % 4 == 0 || !VB, "vector B load needs 4 elements per thread");
  __shared__ __align__(16) float As[2][BKT][BM + 4];
  __shared__ __align__(16) float Bs[2][BKT][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
  float ra[LA], rb[LB];

  auto load_global = [&](int k0) {
    if (VA) {
#pragma unroll
      for (int i = 0; i < LA / 4; ++i) {
        const int idx = tid + i * NT;                         // one float4 = 4 consecutive k of one row
        const int m = idx / (BKT / 4), k = (idx % (BKT / 4)) * 4;
        const int gm = m0 + m, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < g.M && gk < g.K) v = __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)(gm / g.a_div) * g.lda + gk));
        ra[4 * i] = v.x; ra[4 * i + 1] = v.y; ra[4 * i + 2] = v.z; ra[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LA; ++i) {
        const int idx = tid + i * NT;
        int m, k;
        if (TA) { k = idx / BM; m = idx % BM; } else { m = idx / BKT; k = idx % BKT; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.0f;
        if (gm < g.M && gk < g.K)
          v = TA ? __ldg(g.A + (int64_t)gk * g.lda + gm) : __ldg(g.A + (int64_t)(gm / g.a_div) * g.lda + gk);
        ra[i] = v;
      }
    }
    if (VB) {
#pragma unroll
      for (int i = 0; i < LB / 4; ++i) {
        const int idx = tid + i * NT;
        const int n = idx / (BKT / 4), k = (idx % (BKT / 4)) * 4;
        const int gn = n0 + n, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gn < g.N && gk < g.K) v = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)gn * g.ldb + gk));
        rb[4 * i] = v.x; rb[4 * i + 1] = v.y; rb[4 * i + 2] = v.z; rb[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < LB; ++i) {
        const int idx = tid + i * NT;
        int n, k;
        if (TB) { n = idx / BKT; k = idx % BKT; } else { k = idx / BN; n = idx % BN; }
        const int gn = n0 + n, gk = k0 + k;
        float v = 0.0f;
        if (gn < g.N && gk < g.K)
          v = TB ? __ldg(g.B + (int64_t)gn * g.ldb + gk) : __ldg(g.B + (int64_t)gk * g.ldb + gn);
        rb[i] = v;
      }
    }
  };
  auto store_smem = [&](int buf) {
    if (VA) {
#pragma unroll
      for (int i = 0; i < LA / 4; ++i) {
        const int idx = tid + i * NT;
        const int m = idx / (BKT / 4), k = (idx % (BKT / 4)) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][k + q][m] = ra[4 * i + q];
      }
    } else {
#pragma unroll
      for (int i = 0; i < LA; ++i) {
        const int idx = tid + i * NT;
        int m, k;
        if (TA) { k = idx / BM; m = idx % BM; } else { m = idx / BKT; k = idx % BKT; }
        As[buf][k][m] = ra[i];
      }
    }
    if (VB) {
#pragma unroll
      for (int i = 0; i < LB / 4; ++i) {
        const int idx = tid + i * NT;
        const int n = idx / (BKT / 4), k = (idx % (BKT / 4)) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) Bs[buf][k + q][n] = rb[4 * i + q];
      }
    } else {
#pragma unroll
      for (int i = 0; i < LB; ++i) {
        const int idx = tid + i * NT;
        int n, k;
        if (TB) { n = idx / BKT; k = idx % BKT; } else { k = idx / BN; n = idx % BN; }
        Bs[buf][k][n] = rb[i];
      }
    }
  };

  const int nk_all = (g.K + BKT - 1) / BKT;
  // SPLIT == 2: the two CTAs of a (1,1,2) cluster each take half of the k-tiles; rank 0 reduces through DSMEM
  const int kt_begin = SPLIT == 1 ? 0 : (int)blockIdx.z * ((nk_all + 1) / 2);
  const int kt_end = SPLIT == 1 ? nk_all : min(nk_all, kt_begin + (nk_all + 1) / 2);
  const int nk = kt_end - kt_begin;
  if (nk > 0) { load_global(kt_begin * BKT); store_smem(0); }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_global((kt_begin + kt + 1) * BKT);
#pragma unroll
    for (int k = 0; k < BKT; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += (TM >= 4 ? 4 : TM)) {
        if (TM >= 4) {
          const float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * TM + i]);
          av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
        } else {
          const float2 a = *reinterpret_cast<const float2*>(&As[cur][k][ty * TM + i]);
          av[i] = a.x; av[i + 1] = a.y;
        }
      }
#pragma unroll
      for (int j = 0; j < TN; j += (TN >= 4 ? 4 : TN)) {
        if (TN >= 4) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * TN + j]);
          bv[j] = b.x; bv[j + 1] = b.y; bv[j + 2] = b.z; bv[j + 3] = b.w;
        } else {
          const float2 b = *reinterpret_cast<const float2*>(&Bs[cur][k][tx * TN + j]);
          bv[j] = b.x; bv[j + 1] = b.y;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += av[i] * bv[j];
    }
    if (kt + 1 < nk) store_smem(cur ^ 1);
    __syncthreads();
  }
  if (SPLIT == 2) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    static_assert(SPLIT == 1 || sizeof(As) >= sizeof(float) * BM * BN, "reduction buffer does not fit in As");
    float* red = &As[0][0][0];
    if (blockIdx.z == 1) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) red[(ty * TM + i) * BN + tx * TN + j] = acc[i][j];
    }
    cluster.sync();
    if (blockIdx.z == 0) {
      const float* remote = cluster.map_shared_rank(red, 1);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += remote[(ty * TM + i) * BN + tx * TN + j];
    }
    cluster.sync();                    // rank 1's shared memory must outlive rank 0's reads
    if (blockIdx.z != 0) return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias) v += __ldg(g.bias + n);
      if (g.beta) v += g.C[(int64_t)m * g.ldc + n];
      v = apply_act(v, g.act, g.slope);
      if (g.residual) v += g.residual[(int64_t)m * g.ldr + n];
      g.C[(int64_t)m * g.ldc + n] = v;
    }
  }
}

template <int BM, int BN, int BKT, int TM, int TN>
static int launch_gemm_cfg(const GemmArgs& g, bool TA, bool TB, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  // 128-bit loads need K-contiguous storage (A: !TA, B: TB), 16-byte aligned rows and K % 4 == 0
  auto al = [](const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0; };
  const bool vec = (g.K & 3) == 0 && (TA || al(g.A, g.lda)) && (!TB || al(g.B, g.ldb)) && (!TA || TB);
  if (!TA && TB) {
    if (vec) EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, false, true, true>), grid, 256, 0, st, g);
    else EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, false, true, false>), grid, 256, 0, st, g);
  } else if (!TA && !TB) {
    if (vec) EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, false, false, true>), grid, 256, 0, st, g);
    else EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, false, false, false>), grid, 256, 0, st, g);
  } else if (TA && !TB) {
    EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, true, false, false>), grid, 256, 0, st, g);
  } else {
    EG_LAUNCH((gemm_kernel<BM, BN, BKT, TM, TN, true, true, false>), grid, 256, 0, st, g);
  }
  return EG_OK;
}

// 64x64x32 tiles with the k range split over a 2-CTA cluster (DSMEM reduction): for M <= 256 layers whose 64x64
// tiling alone would leave half of the SMs idle. nn.Linear forward layout only (A row-major, W [out,in]).
static int launch_gemm_splitk2(const GemmArgs& g, bool TB, cudaStream_t st) {
  auto al = [](const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0; };
  const bool vec = (g.K & 3) == 0 && al(g.A, g.lda) && (!TB || al(g.B, g.ldb));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((g.N + 63) / 64, (g.M + 63) / 64, 2);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 2;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e;
  if (TB) e = vec ? cudaLaunchKernelEx(&cfg, gemm_kernel<64, 64, 32, 4, 4, false, true, true, 2>, g)
                  : cudaLaunchKernelEx(&cfg, gemm_kernel<64, 64, 32, 4, 4, false, true, false, 2>, g);
  else e = vec ? cudaLaunchKernelEx(&cfg, gemm_kernel<64, 64, 32, 4, 4, false, false, true, 2>, g)
               : cudaLaunchKernelEx(&cfg, gemm_kernel<64, 64, 32, 4, 4, false, false, false, 2>, g);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EG_CUDA_CHECK(e);
  return EG_OK;
}

int launch_gemm(const GemmArgs& g, bool TA, bool TB, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return EG_OK;
  {                                      // tcgen05 3xTF32 path (gemm_tc.cu) when the shape / layout is eligible
    const int rc = launch_gemm_tc(g, TA, TB, st);
    if (rc <= 0) return rc;
  }
  auto ctas = [&](int bm, int bn) { return (int64_t)((g.M + bm - 1) / bm) * ((g.N + bn - 1) / bn); };
  if (!TA && g.a_div == 1 && g.K >= 512 && ctas(64, 64) < kNumSMs * 3 / 4 && 2 * ctas(64, 64) >= kNumSMs / 2)
    return launch_gemm_splitk2(g, TB, st);
  // wave quantisation: a grid of 149..236 big tiles runs a nearly empty second wave; prefer the smaller tile then
  auto wave_eff = [&](int64_t c) { return (double)c / (double)(((c + kNumSMs - 1) / kNumSMs) * kNumSMs); };
  if (ctas(128, 64) >= kNumSMs && wave_eff(ctas(128, 64)) >= 0.8) return launch_gemm_cfg<128, 64, 16, 8, 4>(g, TA, TB, st);
  if (ctas(64, 64) >= kNumSMs) return launch_gemm_cfg<64, 64, 32, 4, 4>(g, TA, TB, st);
  if (ctas(32, 64) >= kNumSMs * 3 / 4) return launch_gemm_cfg<32, 64, 32, 2, 4>(g, TA, TB, st);
  return launch_gemm_cfg<32, 32, 32, 2, 2>(g, TA, TB, st);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
gru_gate_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ b_hh,
                const float* __restrict__ h_in, float* __restrict__ h_out, int M, int H, int ld_h_out,
                float* r_save, float* z_save, float* n_save, float* ghn_save) {
  eg_pdl_enter();
  const int64_t total = (int64_t)M * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / H), j = (int)(i % H);
    const float* gim = gi + (int64_t)m * 3 * H;
    float hr, hz, hn, hp;
    if (gh) {
      const float* ghm = gh + (int64_t)m * 3 * H;
      hr = ghm[j]; hz = ghm[H + j]; hn = ghm[2 * H + j];
      hp = h_in[(int64_t)m * H + j];
    } else {
      hr = __ldg(b_hh + j); hz = __ldg(b_hh + H + j); hn = __ldg(b_hh + 2 * H + j);
      hp = 0.0f;
    }
    const float r = sigmoidf_(gim[j] + hr);
    const float z = sigmoidf_(gim[H + j] + hz);
    const float n = tanhf(gim[2 * H + j] + r * hn);
    h_out[(int64_t)m * ld_h_out + j] = (1.0f - z) * n + z * hp;
    if (r_save) { r_save[i] = r; z_save[i] = z; n_save[i] = n; ghn_save[i] = hn; }
  }
}

int launch_gru_gate(cudaStream_t st, const float* gi, const float* gh, const float* b_hh,
                    const float* h_in, float* h_out, int M, int H, int ld_h_out, float* r_save,
                    float* z_save, float* n_save, float* ghn_save) {
  const int64_t total = (int64_t)M * H;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, kNumSMs * 8);
  EG_LAUNCH_PDL(gru_gate_kernel, grid, 256, 0, st, gi, gh, b_hh, h_in, h_out, M, H, ld_h_out, r_save, z_save,
            n_save, ghn_save);
  return EG_OK;
}

// eval-mode BatchNorm1d: y = (x - mean) / sqrt(var + eps) * gamma + beta
__global__ void __launch_bounds__(256)
bn_eval_kernel(const float* __restrict__ x, int ldx, int M, int D, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ mean,
               const float* __restrict__ var, float eps, float* __restrict__ y) {
  const int64_t total = (int64_t)M * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / D), j = (int)(i % D);
    const float v = x[(int64_t)m * ldx + j];
    y[i] = (v - __ldg(mean + j)) / sqrtf(__ldg(var + j) + eps) * __ldg(gamma + j) + __ldg(beta + j);
  }
}

__global__ void __launch_bounds__(256)
pad_rows_kernel(const float* __restrict__ src, int ld_src, int rows, int cols, float* __restrict__ dst, int ld_dst) {
  const int64_t n = (int64_t)rows * ld_dst;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst), c = (int)(i % ld_dst);
    dst[i] = c < cols ? src[(int64_t)r * ld_src + c] : 0.0f;
  }
}

int launch_pad_rows(cudaStream_t st, const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst) {
  const int64_t n = (int64_t)rows * ld_dst;
  if (n <= 0) return EG_OK;
  EG_LAUNCH(pad_rows_kernel, (int)std::min<int64_t>((n + 255) / 256, kNumSMs * 8), 256, 0, st, src, ld_src, rows, cols, dst, ld_dst);
  return EG_OK;
}

// tensor-core prologue of sample_prior: the two history frames into Y[B,20,D] AND a 16-byte-pitched copy Yp[B,2,DP], and
// the step-invariant GRUCell input row cin[b] = [hx (filled later) | z | y_0 | 0-pad] with pitch KP
__global__ void prologue_pack_kernel(const float* __restrict__ X, int ldx_env, int ldx_frame, const float* __restrict__ z,
                                     int B, int D, int DP, int H, int Z, int KP, float* __restrict__ Y, float* __restrict__ Yp,
                                     float* __restrict__ cin) {
  const int per = 2 * DP + (KP - H);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * per) return;
  const int b = i / per, r = i % per;
  if (r < 2 * DP) {
    const int t = r / DP, d = r % DP;
    const float v = d < D ? X[(int64_t)b * ldx_env + t * ldx_frame + d] : 0.0f;
    Yp[((int64_t)b * 2 + t) * DP + d] = v;
    if (d < D) Y[((int64_t)b * 20 + t) * D + d] = v;
  } else {
    const int c = H + (r - 2 * DP);                        // column of cin
    float v = 0.0f;
    if (c < H + Z) v = z[(int64_t)b * Z + (c - H)];
    else if (c < H + Z + D) v = X[(int64_t)b * ldx_env + ldx_frame + (c - H - Z)];
    cin[(int64_t)b * KP + c] = v;
  }
}

// copy the two history frames' markers into Y[B,20,201]
__global__ void copy_history_kernel(const float* __restrict__ X, int ldx_env, int ldx_frame, int B, int D,
                                    float* __restrict__ Y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * D) return;
  const int b = i / (2 * D), t = (i / D) % 2, d = i % D;
  Y[((int64_t)b * 20 + t) * D + d] = X[(int64_t)b * ldx_env + t * ldx_frame + d];
}

// regressor tail (models_GAMMA_primitive.py:208-219): 22 x 6-D -> rotmat (Gram-Schmidt) -> axis-angle
// (torchgeometry 0.1.2 rotation_matrix_to_angle_axis); writes Yb[b][t][93] for frames t >= t_skip.
__global__ void __launch_bounds__(128)
regressor_tail_kernel(const float* __restrict__ xb_cont, int M, int frames, int t_skip, float* __restrict__ Yb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (row, joint-or-misc slot)
  const int row = i / 32, slot = i % 32;
  if (row >= M) return;
  if ((row % frames) < t_skip) return;
  const float* x = xb_cont + (int64_t)row * 159;
  float* y = Yb + (int64_t)row * 93;
  if (slot < 22) {
    float R[9], aa[3];
    cont6d_to_rotmat(x + 3 + slot * 6, R);
    tgm_rotmat_to_aa(R, aa);
    y[3 + slot * 3 + 0] = aa[0]; y[3 + slot * 3 + 1] = aa[1]; y[3 + slot * 3 + 2] = aa[2];
  } else if (slot == 22) {
    y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
  } else if (slot == 23) {
    for (int k = 0; k < 24; ++k) y[69 + k] = x[135 + k];
  }
}

// ------------------------------------------------------------------------------------------
// Fused row-tile kernels for the two long sequential chains of the motion model. Weights are read from
// TRANSPOSED, 4-padded copies ([K][N4]) so a warp's float4 loads of one k-row are contiguous.
// ------------------------------------------------------------------------------------------
__global__ void transpose_pad_kernel(const float* __restrict__ src, int ld_src, int N, int K, float* __restrict__ dst,
                                     int ld_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // dst[k][n] = src[n][k]
  if (i >= K * ld_dst) return;
  const int k = i / ld_dst, n = i % ld_dst;
  dst[i] = n < N ? src[(int64_t)n * ld_src + k] : 0.0f;
}

// partial[ks][r][n] = sum over this thread's k-slice of x[r][k] * Wt[k][n]; threads = TN (float4 columns) x KS
template <int R>
__device__ __forceinline__ void tile_gemm_splitk(const float* __restrict__ Wt, int ldw, int K, int N,
                                                 const float* xs, int ldx, float* part, int ldp, int tid,
                                                 int nthreads, int& KS) {
  const int TN = (N + 3) >> 2;
  KS = nthreads / TN;
  if (KS > 8) KS = 8;
  if (KS < 1) KS = 1;
  if (tid < TN * KS) {
    const int n4 = tid % TN, ks = tid / TN;
    const int Kc = (K + KS - 1) / KS;
    const int k0 = ks * Kc, k1 = min(K, k0 + Kc);
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
    const float4* wp = reinterpret_cast<const float4*>(Wt + 4 * n4);
    const int ldw4 = ldw >> 2;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {                      // 8 independent 128-bit weight loads in flight per thread
      float4 wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = __ldg(wp + (int64_t)(k + u) * ldw4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float x = xs[r * ldx + k + u];
          acc[r][0] += x * wv[u].x; acc[r][1] += x * wv[u].y; acc[r][2] += x * wv[u].z; acc[r][3] += x * wv[u].w;
        }
      }
    }
    for (; k < k1; ++k) {
      const float4 w0 = __ldg(wp + (int64_t)k * ldw4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float x0 = xs[r * ldx + k];
        acc[r][0] += x0 * w0.x; acc[r][1] += x0 * w0.y; acc[r][2] += x0 * w0.z; acc[r][3] += x0 * w0.w;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      *reinterpret_cast<float4*>(part + (ks * R + r) * ldp + 4 * n4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
}

__device__ __forceinline__ void cp_async16_nn(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

struct DecodeW {            // transposed weights of the 18-step decode loop
  const float *WyT, *WhhT, *W1T, *W2T, *WoT;      // [201][768] [256][768] [256][512] [512][256] [256][204]
  const float *bhh, *b1, *b2, *bo;
};

// 18 GRUCell + MLP steps for DR rows per CTA (models_GAMMA_primitive.py:91-99). h0 = drnn_mlp(hx) and the
// step-invariant input term c = [hx,z] W_ih[:, :384]^T + b_ih are computed beforehand by the layer kernels.
constexpr int DR = 2;
constexpr int DEC_THREADS = 384;
__global__ void __launch_bounds__(DEC_THREADS)
fused_decode_kernel(DecodeW w, const float* __restrict__ c_in, const float* __restrict__ h_in, float* __restrict__ Y,
                    int B, int D, int H, int Hm) {
  extern __shared__ __align__(16) float sm[];
  const int H3 = 3 * H, D4 = (D + 3) & ~3;
  float* cs = sm;                       // [DR][H3]
  float* hs = cs + DR * H3;             // [DR][H]
  float* ys = hs + DR * H;              // [DR][D4]
  float* t1 = ys + DR * D4;             // [DR][Hm]
  float* t2 = t1 + DR * Hm;             // [DR][H]
  float* part = t2 + DR * H;            // [8][DR][H3] split-k partials (gi)
  float* part2 = part + 8 * DR * H3;    // [8][DR][H3] (gh)
  const int tid = threadIdx.x, b0 = blockIdx.x * DR;
  const int ldY = 20 * D;
  for (int i = tid; i < DR * H3; i += DEC_THREADS) { const int r = i / H3, b = min(b0 + r, B - 1); cs[i] = c_in[(int64_t)b * H3 + i % H3]; }
  for (int i = tid; i < DR * H; i += DEC_THREADS) { const int r = i / H, b = min(b0 + r, B - 1); hs[i] = h_in[(int64_t)b * H + i % H]; }
  for (int i = tid; i < DR * D4; i += DEC_THREADS) {
    const int r = i / D4, d = i % D4, b = min(b0 + r, B - 1);
    ys[i] = d < D ? Y[(int64_t)b * ldY + D + d] : 0.0f;          // history frame 1
  }
  __syncthreads();
  for (int step = 0; step < 18; ++step) {
    int KS1, KS2;
    tile_gemm_splitk<DR>(w.WyT, H3, D, H3, ys, D4, part, H3, tid, DEC_THREADS, KS1);
    tile_gemm_splitk<DR>(w.WhhT, H3, H, H3, hs, H, part2, H3, tid, DEC_THREADS, KS2);
    __syncthreads();
    // GRUCell gates (PyTorch order r,z,n)
    for (int i = tid; i < DR * H; i += DEC_THREADS) {
      const int r = i / H, j = i % H;
      float gi[3], gh[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float a = cs[r * H3 + g * H + j], bq = __ldg(w.bhh + g * H + j);
        for (int ks = 0; ks < KS1; ++ks) a += part[(ks * DR + r) * H3 + g * H + j];
        for (int ks = 0; ks < KS2; ++ks) bq += part2[(ks * DR + r) * H3 + g * H + j];
        gi[g] = a; gh[g] = bq;
      }
      const float rr = 1.0f / (1.0f + expf(-(gi[0] + gh[0])));
      const float zz = 1.0f / (1.0f + expf(-(gi[1] + gh[1])));
      const float nn = tanhf(gi[2] + rr * gh[2]);
      t2[i] = (1.0f - zz) * nn + zz * hs[i];                  // new h, staged
    }
    __syncthreads();
    for (int i = tid; i < DR * H; i += DEC_THREADS) hs[i] = t2[i];
    __syncthreads();
    int KS;
    tile_gemm_splitk<DR>(w.W1T, Hm, H, Hm, hs, H, part, Hm, tid, DEC_THREADS, KS);
    __syncthreads();
    for (int i = tid; i < DR * Hm; i += DEC_THREADS) {
      const int r = i / Hm, j = i % Hm;
      float a = __ldg(w.b1 + j);
      for (int ks = 0; ks < KS; ++ks) a += part[(ks * DR + r) * Hm + j];
      t1[i] = tanhf(a);
    }
    __syncthreads();
    tile_gemm_splitk<DR>(w.W2T, H, Hm, H, t1, Hm, part, H, tid, DEC_THREADS, KS);
    __syncthreads();
    for (int i = tid; i < DR * H; i += DEC_THREADS) {
      const int r = i / H, j = i % H;
      float a = __ldg(w.b2 + j);
      for (int ks = 0; ks < KS; ++ks) a += part[(ks * DR + r) * H + j];
      t2[i] = tanhf(a);
    }
    __syncthreads();
    tile_gemm_splitk<DR>(w.WoT, D4, H, D, t2, H, part, D4, tid, DEC_THREADS, KS);
    __syncthreads();
    for (int i = tid; i < DR * D; i += DEC_THREADS) {
      const int r = i / D, j = i % D;
      float a = __ldg(w.bo + j);
      for (int ks = 0; ks < KS; ++ks) a += part[(ks * DR + r) * D4 + j];
      a += ys[r * D4 + j];                                      // residual: y_i = d_out(..) + y_{i-1}
      ys[r * D4 + j] = a;
      if (b0 + r < B) Y[(int64_t)(b0 + r) * ldY + (2 + step) * D + j] = a;
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------------
// Weight-stationary decode (B >= 64): fused_decode_kernel re-streams all 2.66 MB of step weights from L2 for every
// pair of rows and every step (6.1 GB per 256-env call - the per-SM L2 ingest rate is its bound). Here the work is cut
// in 2-D: a CTA owns a block of 32 rows and 1/18 of every layer's OUTPUT columns, keeps that weight slice (166 KB) in
// shared memory for all 18 steps, and the 18 CTAs of a row block exchange each layer's activations through L2
// (4 hand-offs per step over a per-row-block arrival counter). L2 traffic drops to ~1 GB per call.
// Arithmetic per output element is the plain fp32 k-ascending dot product (same rounding class as the layer kernels).
namespace dws {
constexpr int RB = 32, CS = 18, THREADS = 256, KC = 128;
constexpr int NG = 48, N1 = 32, N2 = 16, NO = 16;          // padded slice widths: GRU (3 gates x 16), W1, W2, Wo
}

struct DwsArgs {
  DecodeW w;
  const float* c_in;      // [B][3H]
  const float* h_in;      // [B][H]
  float* Y;               // [B][20][D]
  float* Hx[2];           // [B][H] ping-pong hidden state
  float* T1;              // [B][Hm]
  float* T2;              // [B][H]
  float* Yp[2];           // [B][D4] ping-pong padded previous frame
  unsigned* cnt;          // [row blocks] arrival counters (zeroed before the launch)
  int B, D, H, Hm;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all 18 CTAs of the row block have published phase `target / CS`
__device__ __forceinline__ void rowblock_sync(unsigned* cnt, unsigned target) {
  __syncthreads();                       // every thread's exchange-buffer writes happen-before thread 0 ...
  if (threadIdx.x == 0) {
    __threadfence();                     // ... whose cumulative gpu-scope fence orders them before the arrival
    atomicAdd(cnt, 1u);
    unsigned spins = 0;
    while (ld_acquire_u32(cnt) < target) {
      if (++spins > (1u << 24)) __trap();          // a peer CTA never arrived: fail loudly instead of hanging the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

// acc[r][c] += sum_k in[row rg*4+r][k] * Ws[k][cl + 32 c]: the input rows stream from global through a double-buffered
// shared-memory ring (cp.async.cg = L2, never a stale L1 line), the weight slice is resident in shared memory.
template <int NC>     // columns per thread (1 or 2); second column only for lanes < ncols2
__device__ __forceinline__ void dws_gemm(const float* __restrict__ in, int ld_in, int K, int b0, int B, const float* Ws, int ldw,
                                         float* ring, int tid, float acc[4][NC], bool col2) {
  using namespace dws;
  const int rg = tid >> 5, cl = tid & 31;
  const int nchunks = (K + KC - 1) / KC;
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int k0 = c * KC, kc = min(KC, K - k0);           // multiples of 4
      float* dst = ring + (c & 1) * RB * KC;
      const int n4 = kc >> 2;
      for (int i = tid; i < RB * n4; i += THREADS) {
        const int r = i / n4, q = i - r * n4;
        const int b = min(b0 + r, B - 1);
        cp_async16_nn(dst + r * KC + q * 4, in + (int64_t)b * ld_in + k0 + q * 4);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    issue(c + 1);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    const int k0 = c * KC, kc = min(KC, K - k0);
    const float* xs = ring + (c & 1) * RB * KC + rg * 4 * KC;
    const float* wp = Ws + (int64_t)k0 * ldw + cl;
#pragma unroll 2
    for (int k = 0; k < kc; k += 4) {
      float4 x[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(xs + r * KC + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float w0 = wp[(k + kk) * ldw];
        const float w1 = (NC > 1 && col2) ? wp[(k + kk) * ldw + 32] : 0.0f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float xv = kk == 0 ? x[r].x : kk == 1 ? x[r].y : kk == 2 ? x[r].z : x[r].w;
          acc[r][0] = fmaf(xv, w0, acc[r][0]);
          if (NC > 1) acc[r][1] = fmaf(xv, w1, acc[r][1]);
        }
      }
    }
    __syncthreads();                     // ring buffer (c & 1) may be refilled by issue(c + 2)
  }
}

__global__ void __launch_bounds__(dws::THREADS, 1)
decode_ws_kernel(const DwsArgs a) {
  using namespace dws;
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, Hm = a.Hm, H3 = 3 * H, D4 = (D + 3) & ~3, B = a.B;
  float* Wy_s = sm;                          // [D4][NG]
  float* Whh_s = Wy_s + D4 * NG;             // [H][NG]
  float* W1_s = Whh_s + H * NG;              // [H][N1]
  float* W2_s = W1_s + H * N1;               // [Hm][N2]
  float* Wo_s = W2_s + Hm * N2;              // [H][NO]
  float* c_s = Wo_s + H * NO;                // [RB][NG]
  float* gin = c_s + RB * NG;                // [RB][NG] gi (all gates)
  float* ghn = gin + RB * NG;                // [RB][NG] gh (all gates)
  float* ring = ghn + RB * NG;               // [2][RB][KC]
  const int tid = threadIdx.x, cs = blockIdx.x, rb = blockIdx.y, b0 = rb * RB;
  const int rg = tid >> 5, cl = tid & 31;
  const int j0 = cs * H / CS, nj = (cs + 1) * H / CS - j0;        // my hidden columns (h, t2)
  const int m0 = cs * Hm / CS, nm = (cs + 1) * Hm / CS - m0;      // my mlp columns (t1)
  const int d0 = cs * D / CS, nd = (cs + 1) * D / CS - d0;        // my output columns (y)
  unsigned* cnt = a.cnt + rb;
  unsigned phase = 0;

  // ---- resident weight slices + the step-invariant input term ----
  for (int i = tid; i < D4 * NG; i += THREADS) {
    const int k = i / NG, c = i % NG, g = c >> 4, jj = c & 15;
    Wy_s[i] = (k < D && jj < nj) ? __ldg(a.w.WyT + (int64_t)k * H3 + g * H + j0 + jj) : 0.0f;
  }
  for (int i = tid; i < H * NG; i += THREADS) {
    const int k = i / NG, c = i % NG, g = c >> 4, jj = c & 15;
    Whh_s[i] = jj < nj ? __ldg(a.w.WhhT + (int64_t)k * H3 + g * H + j0 + jj) : 0.0f;
  }
  for (int i = tid; i < H * N1; i += THREADS) { const int k = i / N1, c = i % N1; W1_s[i] = c < nm ? __ldg(a.w.W1T + (int64_t)k * Hm + m0 + c) : 0.0f; }
  for (int i = tid; i < Hm * N2; i += THREADS) { const int k = i / N2, c = i % N2; W2_s[i] = c < nj ? __ldg(a.w.W2T + (int64_t)k * H + j0 + c) : 0.0f; }
  for (int i = tid; i < H * NO; i += THREADS) { const int k = i / NO, c = i % NO; Wo_s[i] = c < nd ? __ldg(a.w.WoT + (int64_t)k * D4 + d0 + c) : 0.0f; }
  for (int i = tid; i < RB * NG; i += THREADS) {
    const int r = i / NG, c = i % NG, g = c >> 4, jj = c & 15, b = min(b0 + r, B - 1);
    c_s[i] = jj < nj ? a.c_in[(int64_t)b * H3 + g * H + j0 + jj] : 0.0f;
  }
  // ---- phase 0: publish my slice of h0 and of the last history frame ----
  for (int i = tid; i < RB * nj; i += THREADS) {
    const int r = i / nj, jj = i % nj, b = b0 + r;
    if (b < B) a.Hx[0][(int64_t)b * H + j0 + jj] = a.h_in[(int64_t)b * H + j0 + jj];
  }
  for (int i = tid; i < RB * nd; i += THREADS) {
    const int r = i / nd, dd = i % nd, b = b0 + r;
    if (b < B) a.Yp[0][(int64_t)b * D4 + d0 + dd] = a.Y[(int64_t)b * 20 * D + D + d0 + dd];
  }
  if (cs == CS - 1)                                     // zero padding columns of both ping-pong frames
    for (int i = tid; i < RB * (D4 - D) * 2; i += THREADS) {
      const int r = (i >> 1) / (D4 - D), dd = (i >> 1) % (D4 - D), b = b0 + r;
      if (b < B) a.Yp[i & 1][(int64_t)b * D4 + D + dd] = 0.0f;
    }
  rowblock_sync(cnt, ++phase * CS);

  for (int step = 0; step < 18; ++step) {
    const int cur = step & 1, nxt = cur ^ 1;
    // ---- GRUCell: gi = c + y_prev Wy, gh = b_hh + h Whh for my 3 x nj gate columns ----
    {
      float gi[4][2], gh[4][2];
      const bool col2 = cl < 16;                        // columns cl (gates r / z) and 32 + cl (gate n)
#pragma unroll
      for (int r = 0; r < 4; ++r) { gi[r][0] = gi[r][1] = 0.0f; gh[r][0] = gh[r][1] = 0.0f; }
      dws_gemm<2>(a.Yp[cur], D4, D4, b0, B, Wy_s, NG, ring, tid, gi, col2);
      dws_gemm<2>(a.Hx[cur], H, H, b0, B, Whh_s, NG, ring, tid, gh, col2);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = rg * 4 + r;
        gin[row * NG + cl] = gi[r][0]; ghn[row * NG + cl] = gh[r][0];
        if (col2) { gin[row * NG + 32 + cl] = gi[r][1]; ghn[row * NG + 32 + cl] = gh[r][1]; }
      }
      __syncthreads();
      for (int i = tid; i < RB * nj; i += THREADS) {
        const int r = i / nj, jj = i % nj, b = b0 + r, j = j0 + jj;
        const float gir = c_s[r * NG + jj] + gin[r * NG + jj], giz = c_s[r * NG + 16 + jj] + gin[r * NG + 16 + jj];
        const float gin_ = c_s[r * NG + 32 + jj] + gin[r * NG + 32 + jj];
        const float ghr = __ldg(a.w.bhh + j) + ghn[r * NG + jj], ghz = __ldg(a.w.bhh + H + j) + ghn[r * NG + 16 + jj];
        const float ghn_ = __ldg(a.w.bhh + 2 * H + j) + ghn[r * NG + 32 + jj];
        const float rr = 1.0f / (1.0f + expf(-(gir + ghr)));
        const float zz = 1.0f / (1.0f + expf(-(giz + ghz)));
        const float nn = tanhf(gin_ + rr * ghn_);
        if (b < B) {
          const float hp = __ldcg(a.Hx[cur] + (int64_t)b * H + j);
          a.Hx[nxt][(int64_t)b * H + j] = (1.0f - zz) * nn + zz * hp;
        }
      }
    }
    rowblock_sync(cnt, ++phase * CS);
    // ---- t1 = tanh(b1 + h W1) ----
    {
      float acc[4][1];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = 0.0f;
      dws_gemm<1>(a.Hx[nxt], H, H, b0, B, W1_s, N1, ring, tid, acc, false);
      if (cl < nm) {
        const float bb = __ldg(a.w.b1 + m0 + cl);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int b = b0 + rg * 4 + r;
          if (b < B) a.T1[(int64_t)b * Hm + m0 + cl] = tanhf(bb + acc[r][0]);
        }
      }
    }
    rowblock_sync(cnt, ++phase * CS);
    // ---- t2 = tanh(b2 + t1 W2) ----
    {
      float acc[4][1];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = 0.0f;
      // lanes beyond the slice width keep the ring / barrier protocol and read W2_s[0] (one call site: no divergent barrier)
      dws_gemm<1>(a.T1, Hm, Hm, b0, B, cl < N2 ? W2_s : W2_s - cl, cl < N2 ? N2 : 0, ring, tid, acc, false);
      if (cl < nj) {
        const float bb = __ldg(a.w.b2 + j0 + cl);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int b = b0 + rg * 4 + r;
          if (b < B) a.T2[(int64_t)b * H + j0 + cl] = tanhf(bb + acc[r][0]);
        }
      }
    }
    rowblock_sync(cnt, ++phase * CS);
    // ---- y = bo + t2 Wo + y_prev ----
    {
      float acc[4][1];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = 0.0f;
      dws_gemm<1>(a.T2, H, H, b0, B, cl < NO ? Wo_s : Wo_s - cl, cl < NO ? NO : 0, ring, tid, acc, false);
      if (cl < nd) {
        const float bb = __ldg(a.w.bo + d0 + cl);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int b = b0 + rg * 4 + r;
          if (b < B) {
            const float y = bb + acc[r][0] + __ldcg(a.Yp[cur] + (int64_t)b * D4 + d0 + cl);
            a.Yp[nxt][(int64_t)b * D4 + d0 + cl] = y;
            a.Y[(int64_t)b * 20 * D + (2 + step) * D + d0 + cl] = y;
          }
        }
      }
    }
    rowblock_sync(cnt, ++phase * CS);
  }
}

struct RegW {               // transposed regressor weights
  const float *WaT, *WbT, *WcT;     // in_fc split: markers [201][128], xb [159][128], betas [10][128]
  const float* b_in;
  const float* blkT;                // n_blocks x 2 x [128][128]
  const float* blk_b;               // n_blocks x 2 x [128]
  const float *WoT, *b_out;         // [128][160], [159]
};

// ---- weight-chunk pipeline shared by the fused kernels: the whole ordered stream of weight chunks of a kernel is
// described by a table in shared memory; chunk c+1 is fetched with cp.async (all threads) while chunk c is consumed.
struct WChunk { const float* p; int rows; int ld; int pad; };

struct ChunkPipe {
  const WChunk* table; int n; float* buf; int buf_floats; int tid, nthreads; int next;
  __device__ __forceinline__ void issue(int c) {
    if (c < n) {
      const WChunk ch = table[c];
      const float4* src = reinterpret_cast<const float4*>(ch.p);
      float4* dst = reinterpret_cast<float4*>(buf + (c & 1) * buf_floats);
      const int n4 = ch.rows * ch.ld / 4;
      for (int i = tid; i < n4; i += nthreads) cp_async16_nn(dst + i, src + i);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }
  // returns the shared-memory address of chunk c (rows x ld floats); call release() when done reading it
  __device__ __forceinline__ const float* acquire(int c) {
    issue(c + 1);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    return buf + (c & 1) * buf_floats;
  }
  __device__ __forceinline__ void release() { __syncthreads(); }
};

// acc[r][0..3] += sum_k x[r][k] * W[k][4 n4 .. 4 n4+3] for a warp's RW rows, weights from a shared-memory chunk
template <int RW_>
__device__ __forceinline__ void warp_rows_gemm_smem(const float* ws, int ldw, int K, int n4, const float* xs, int ldx,
                                                    float acc[RW_][4]) {
  const float4* wp = reinterpret_cast<const float4*>(ws) + n4;
  const int ldw4 = ldw >> 2;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float4 wv = wp[k * ldw4];
#pragma unroll
    for (int r = 0; r < RW_; ++r) {
      const float x = xs[r * ldx + k];
      acc[r][0] += x * wv.x; acc[r][1] += x * wv.y; acc[r][2] += x * wv.z; acc[r][3] += x * wv.w;
    }
  }
}

// Whole MoshRegressor._forward (3 recurrences x (in_fc + 10 residual blocks + out_fc)) for RR rows per CTA,
// activations resident in shared memory, weights streamed through a double-buffered shared-memory chunk pipeline
// (models_GAMMA_primitive.py:222-259, ResNetBlock :160-175).
constexpr int RR = 36, RWARPS = 9, RW = 4;
constexpr int RCHUNK_ROWS = 64, RCHUNK_FLOATS = RCHUNK_ROWS * 160, RTABLE = 192;
__global__ void __launch_bounds__(RWARPS * 32)
fused_regressor_kernel(RegW w, const float* __restrict__ Yin, const float* __restrict__ betas, int betas_div, int M,
                       int n_blocks, int n_recur, float* __restrict__ xb_out) {
  constexpr int HR = 128, D = 201, BD = 159, BD4 = 160, LDX = 204, NTH = RWARPS * 32;
  extern __shared__ __align__(16) float sm[];
  float* xr = sm;                   // [RR][LDX] markers
  float* be = xr + RR * LDX;        // [RR][12]
  float* base = be + RR * 12;       // [RR][HR]
  float* h = base + RR * HR;        // [RR][HR]
  float* t = h + RR * HR;           // [RR][HR]
  float* xb = t + RR * HR;          // [RR][BD4]
  float* wbuf = xb + RR * BD4;      // 2 x RCHUNK_FLOATS
  WChunk* table = reinterpret_cast<WChunk*>(wbuf + 2 * RCHUNK_FLOATS);
  __shared__ int n_chunks_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * RR;
  if (tid == 0) {                   // ordered stream of weight chunks for the whole forward pass
    int n = 0;
    auto add = [&](const float* p, int K, int ld) {
      for (int k0 = 0; k0 < K; k0 += RCHUNK_ROWS) { table[n].p = p + (int64_t)k0 * ld; table[n].rows = min(RCHUNK_ROWS, K - k0); table[n].ld = ld; ++n; }
    };
    add(w.WaT, D, HR); add(w.WcT, 10, HR);
    for (int rec = 0; rec < n_recur; ++rec) {
      if (rec > 0) add(w.WbT, BD, HR);
      for (int b = 0; b < n_blocks * 2; ++b) add(w.blkT + (int64_t)b * HR * HR, HR, HR);
      add(w.WoT, HR, BD4);
    }
    n_chunks_s = n;
  }
  for (int i = tid; i < RR * LDX; i += NTH) {
    const int r = i / LDX, d = i % LDX, m = min(m0 + r, M - 1);
    xr[i] = d < D ? Yin[(int64_t)m * D + d] : 0.0f;
  }
  for (int i = tid; i < RR * 12; i += NTH) {
    const int r = i / 12, d = i % 12, m = min(m0 + r, M - 1);
    be[i] = d < 10 ? betas[(int64_t)(m / betas_div) * 10 + d] : 0.0f;
  }
  for (int i = tid; i < RR * BD4; i += NTH) xb[i] = 0.0f;
  __syncthreads();
  ChunkPipe pipe{table, n_chunks_s, wbuf, RCHUNK_FLOATS, tid, NTH, 0};
  pipe.issue(0);
  int c = 0;
  const int r0 = warp * RW;
  float acc[RW][4];
  auto zero = [&]() {
#pragma unroll
    for (int r = 0; r < RW; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
  };
  // acc += X[r0.., :K] W  with W streamed in chunks (N = 128 columns, lane owns 4)
  auto stream_gemm = [&](const float* xs, int ldx, int K) {
    for (int k0 = 0; k0 < K; k0 += RCHUNK_ROWS) {
      const float* ws = pipe.acquire(c++);
      warp_rows_gemm_smem<RW>(ws, HR, min(RCHUNK_ROWS, K - k0), lane, xs + r0 * ldx + k0, ldx, acc);
      pipe.release();
    }
  };
  // base = markers Wa^T + betas Wc^T + b_in   (constant over the recurrences)
  zero();
  stream_gemm(xr, LDX, D);
  stream_gemm(be, 12, 10);
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(w.b_in) + lane);
#pragma unroll
    for (int r = 0; r < RW; ++r)
      *reinterpret_cast<float4*>(base + (r0 + r) * HR + 4 * lane) =
          make_float4(acc[r][0] + bb.x, acc[r][1] + bb.y, acc[r][2] + bb.z, acc[r][3] + bb.w);
  }
  for (int rec = 0; rec < n_recur; ++rec) {
    zero();
    if (rec > 0) stream_gemm(xb, BD4, BD);            // h = base + xb Wb^T
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const float4 bv = *reinterpret_cast<const float4*>(base + (r0 + r) * HR + 4 * lane);
      *reinterpret_cast<float4*>(h + (r0 + r) * HR + 4 * lane) =
          make_float4(acc[r][0] + bv.x, acc[r][1] + bv.y, acc[r][2] + bv.z, acc[r][3] + bv.w);
    }
    for (int blk = 0; blk < n_blocks; ++blk) {
      const float* B0 = w.blk_b + (blk * 2) * HR;
      zero();
      stream_gemm(h, HR, HR);                          // (the barriers inside order the smem hand-offs)
      float4 bb = __ldg(reinterpret_cast<const float4*>(B0) + lane);
#pragma unroll
      for (int r = 0; r < RW; ++r)
        *reinterpret_cast<float4*>(t + (r0 + r) * HR + 4 * lane) =
            make_float4(fmaxf(acc[r][0] + bb.x, 0.f), fmaxf(acc[r][1] + bb.y, 0.f), fmaxf(acc[r][2] + bb.z, 0.f),
                        fmaxf(acc[r][3] + bb.w, 0.f));
      zero();
      stream_gemm(t, HR, HR);
      bb = __ldg(reinterpret_cast<const float4*>(B0 + HR) + lane);
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float4* hp = reinterpret_cast<float4*>(h + (r0 + r) * HR + 4 * lane);
        const float4 hv = *hp;
        *hp = make_float4(fmaxf(acc[r][0] + bb.x, 0.f) + hv.x, fmaxf(acc[r][1] + bb.y, 0.f) + hv.y,
                          fmaxf(acc[r][2] + bb.z, 0.f) + hv.z, fmaxf(acc[r][3] + bb.w, 0.f) + hv.w);
      }
    }
    // xb += h Wout^T + b_out  (159 outputs = 40 float4 groups: lanes 0..31, then lanes 0..7 again)
    float acc2[RW][4];
    zero();
#pragma unroll
    for (int r = 0; r < RW; ++r) acc2[r][0] = acc2[r][1] = acc2[r][2] = acc2[r][3] = 0.0f;
    for (int k0 = 0; k0 < HR; k0 += RCHUNK_ROWS) {
      const float* ws = pipe.acquire(c++);
      warp_rows_gemm_smem<RW>(ws, BD4, min(RCHUNK_ROWS, HR - k0), lane, h + r0 * HR + k0, HR, acc);
      if (lane < BD4 / 4 - 32)
        warp_rows_gemm_smem<RW>(ws, BD4, min(RCHUNK_ROWS, HR - k0), lane + 32, h + r0 * HR + k0, HR, acc2);
      pipe.release();
    }
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = 4 * lane + j;
        xb[(r0 + r) * BD4 + n] += acc[r][j] + __ldg(w.b_out + n);
        const int n2 = 4 * (lane + 32) + j;
        if (lane < BD4 / 4 - 32 && n2 < BD) xb[(r0 + r) * BD4 + n2] += acc2[r][j] + __ldg(w.b_out + n2);
      }
    __syncwarp();
  }
  for (int i = lane; i < RW * BD; i += 32) {
    const int r = r0 + i / BD, n = i % BD;
    if (m0 + r < M) xb_out[(int64_t)(m0 + r) * BD + n] = xb[r * BD4 + n];
  }
}

static inline int ew_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, kNumSMs * 8); }

}  // namespace eg

using namespace eg;

// ------------------------------------------------------------------------------------------
// Motion model (C-VAE predictor + marker->body regressor)
// ------------------------------------------------------------------------------------------
struct EgMotion {
  int device = 0;
  EgMotionDims d;
  std::vector<const float*> w;
  int cap_B = 0;
  float *gi = nullptr, *gh = nullptr, *h = nullptr, *hx = nullptr, *c = nullptr, *t1 = nullptr, *t2 = nullptr;
  float *rh = nullptr, *rbase = nullptr, *rt = nullptr, *xbc = nullptr;
  float *hx2 = nullptr, *yp2 = nullptr;   // decode_ws exchange buffers: [2][B][H], [2][B][D4]
  unsigned* dcnt = nullptr;               // decode_ws per-row-block arrival counters
  int decode_ws = -1;                     // -1: read EG_DECODE_WS (default on) at first use
  float* wt = nullptr;        // transposed, 4-padded weight copies for the fused kernels
  DecodeW dw{};
  RegW rw{};
  int fused = 1;
  mtc::MotionTc* tc = nullptr;            // tcgen05 decode / regressor (motion_tc.cu); nullptr or unavailable -> SIMT fused kernels
  // 16-byte-pitched operands of the decode prologue (D = 201 and H + Z + D = 585 floats are not): x_enc W_ih [3H][DP],
  // d_rnn W_ih [3H][KP], history frames [B,2,DP], GRUCell input rows [B][KP]
  int DP = 0, KP = 0;
  float *wx_pad = nullptr, *wd_pad = nullptr, *yp_pad = nullptr, *cin = nullptr;
};

namespace {
enum {  // weight table order (state_dict names in comments)
  P_XENC_WIH = 0, P_XENC_WHH, P_XENC_BIH, P_XENC_BHH,         // predictor.x_enc.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0}
  P_DRNN_W0, P_DRNN_B0, P_DRNN_W1, P_DRNN_B1, P_DRNN_W2, P_DRNN_B2,   // predictor.drnn_mlp.layers.{0,1,2}
  P_DRNN_WIH, P_DRNN_WHH, P_DRNN_BIH, P_DRNN_BHH,             // predictor.d_rnn.{weight_ih,weight_hh,bias_ih,bias_hh}
  P_DMLP_W0, P_DMLP_B0, P_DMLP_W1, P_DMLP_B1,                 // predictor.d_mlp.layers.{0,1}
  P_DOUT_W, P_DOUT_B,                                         // predictor.d_out
  R_IN_W, R_IN_B,                                             // regressor.pnet.in_fc
  R_BLOCKS                                                    // then n_blocks x {l0.w,l0.b,l1.w,l1.b}, out_fc.{w,b}
};

int motion_ws(EgMotion* h, int B) {
  if (B <= h->cap_B) return EG_OK;
  const EgMotionDims& d = h->d;
  float** bufs[] = {&h->gi, &h->gh, &h->h, &h->hx, &h->c, &h->t1, &h->t2, &h->rh, &h->rbase, &h->rt, &h->xbc, &h->hx2, &h->yp2};
  for (auto p : bufs) { cudaFree(*p); *p = nullptr; }
  cudaFree(h->dcnt); h->dcnt = nullptr;
  h->cap_B = 0;
  const size_t b = (size_t)B, M = b * 20;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->gi, b * 3 * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->gh, b * 3 * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->h, b * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->hx, b * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->c, b * 3 * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->t1, b * d.mlp_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->t2, b * d.mlp_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->rh, M * d.reg_h * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->rbase, M * d.reg_h * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->rt, M * d.reg_h * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->xbc, M * 159 * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->hx2, 2 * b * d.h_dim * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->yp2, 2 * b * (size_t)((d.in_dim + 3) & ~3) * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->dcnt, ((b + dws::RB - 1) / dws::RB) * sizeof(unsigned)));
  cudaFree(h->yp_pad); cudaFree(h->cin); h->yp_pad = h->cin = nullptr;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->yp_pad, b * 2 * h->DP * 4));
  EG_CUDA_CHECK(cudaMalloc((void**)&h->cin, b * h->KP * 4));
  h->cap_B = B;
  return EG_OK;
}
}  // namespace

static int motion_build_transposed(EgMotion* h, cudaStream_t st) {
  const EgMotionDims& d = h->d;
  const int D = d.in_dim, H = d.h_dim, Z = d.z_dim, Hm = d.mlp_dim, H3 = 3 * H, D4 = (D + 3) & ~3;
  const int Hr = d.reg_h, BD = d.body_dim, BD4 = (BD + 3) & ~3, Kr = D + BD + 10, nb = d.reg_blocks;
  const float* const* w = h->w.data();
  const size_t n_dec = (size_t)D * H3 + (size_t)H * H3 + (size_t)H * Hm + (size_t)Hm * H + (size_t)H * D4;
  const size_t n_reg = (size_t)D * Hr + (size_t)BD * Hr + (size_t)10 * Hr + (size_t)nb * 2 * Hr * Hr + (size_t)nb * 2 * Hr +
                       (size_t)Hr * BD4;
  if (!h->wt) EG_CUDA_CHECK(cudaMalloc((void**)&h->wt, (n_dec + n_reg) * sizeof(float)));
  float* p = h->wt;
  auto tr = [&](const float* src, int ld_src, int N, int K, int ld_dst) -> const float* {
    float* dst = p;
    const int total = K * ld_dst;
    transpose_pad_kernel<<<(total + 255) / 256, 256, 0, st>>>(src, ld_src, N, K, dst, ld_dst);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    p += total;
    return dst;
  };
  const int Kin = H + Z + D;
  h->dw.WyT = tr(w[P_DRNN_WIH] + H + Z, Kin, H3, D, H3);
  h->dw.WhhT = tr(w[P_DRNN_WHH], H, H3, H, H3);
  h->dw.W1T = tr(w[P_DMLP_W0], H, Hm, H, Hm);
  h->dw.W2T = tr(w[P_DMLP_W1], Hm, H, Hm, H);
  h->dw.WoT = tr(w[P_DOUT_W], H, D, H, D4);
  h->dw.bhh = w[P_DRNN_BHH]; h->dw.b1 = w[P_DMLP_B0]; h->dw.b2 = w[P_DMLP_B1]; h->dw.bo = w[P_DOUT_B];
  const float* Win = w[R_IN_W];
  h->rw.WaT = tr(Win, Kr, Hr, D, Hr);
  h->rw.WbT = tr(Win + D, Kr, Hr, BD, Hr);
  h->rw.WcT = tr(Win + D + BD, Kr, Hr, 10, Hr);
  h->rw.b_in = w[R_IN_B];
  const float* const* wb = w + R_BLOCKS;
  h->rw.blkT = p;
  for (int k = 0; k < nb; ++k) { tr(wb[k * 4], Hr, Hr, Hr, Hr); tr(wb[k * 4 + 2], Hr, Hr, Hr, Hr); }
  float* bb = p;
  for (int k = 0; k < nb; ++k) {
    EG_CUDA_CHECK(cudaMemcpyAsync(p, wb[k * 4 + 1], Hr * sizeof(float), cudaMemcpyDeviceToDevice, st)); p += Hr;
    EG_CUDA_CHECK(cudaMemcpyAsync(p, wb[k * 4 + 3], Hr * sizeof(float), cudaMemcpyDeviceToDevice, st)); p += Hr;
  }
  h->rw.blk_b = bb;
  h->rw.WoT = tr(wb[nb * 4], Hr, BD, Hr, BD4);
  h->rw.b_out = wb[nb * 4 + 1];
  EG_CUDA_CHECK(cudaGetLastError());
  // padded copies for the tensor-core prologue
  if (!h->wx_pad) EG_CUDA_CHECK(cudaMalloc((void**)&h->wx_pad, (size_t)H3 * h->DP * sizeof(float)));
  if (!h->wd_pad) EG_CUDA_CHECK(cudaMalloc((void**)&h->wd_pad, (size_t)H3 * h->KP * sizeof(float)));
  int rc = launch_pad_rows(st, w[P_XENC_WIH], D, H3, D, h->wx_pad, h->DP);
  if (rc) return rc;
  if ((rc = launch_pad_rows(st, w[P_DRNN_WIH], Kin, H3, Kin, h->wd_pad, h->KP))) return rc;
  return EG_OK;
}

static size_t decode_smem(const EgMotionDims& d) {
  const int H = d.h_dim, H3 = 3 * H, D4 = (d.in_dim + 3) & ~3, Hm = d.mlp_dim;
  return sizeof(float) * ((size_t)DR * (H3 + H + D4 + Hm + H) + (size_t)2 * 8 * DR * H3);
}
static size_t decode_ws_smem(const EgMotionDims& d) {
  const int H = d.h_dim, D4 = (d.in_dim + 3) & ~3, Hm = d.mlp_dim;
  return sizeof(float) * ((size_t)D4 * dws::NG + (size_t)H * dws::NG + (size_t)H * dws::N1 + (size_t)Hm * dws::N2 +
                          (size_t)H * dws::NO + (size_t)3 * dws::RB * dws::NG + (size_t)2 * dws::RB * dws::KC);
}
constexpr size_t kRegSmem = sizeof(float) * (RR * (204 + 12 + 128 * 3 + 160) + 2 * RCHUNK_FLOATS) + sizeof(WChunk) * RTABLE;

extern "C" int eg_motion_create(const EgMotionDims* dims, const void* const* weights_host, int n_weights,
                                int device, EgMotion** out) {
  EG_REQUIRE(dims && weights_host && out, "null pointer");
  EG_REQUIRE(dims->in_dim == 201 && dims->body_dim == 159, "marker dim 201 / cont body dim 159 expected");
  const int expect = R_BLOCKS + dims->reg_blocks * 4 + 2;
  EG_REQUIRE(n_weights == expect, "unexpected number of weight tensors");
  for (int i = 0; i < n_weights; ++i) EG_REQUIRE(weights_host[i] != nullptr, "null weight pointer");
  EgMotion* h = new EgMotion();
  h->device = device; h->d = *dims;
  h->DP = (dims->in_dim + 3) & ~3;
  h->KP = (dims->h_dim + dims->z_dim + dims->in_dim + 3) & ~3;
  h->w.assign((const float* const*)weights_host, (const float* const*)weights_host + n_weights);
  EG_CUDA_CHECK(cudaSetDevice(device));
  h->fused = (dims->reg_h == 128 && dims->mlp_dim <= 3 * dims->h_dim) ? 1 : 0;   // fused kernels' static assumptions
  if (h->fused) {
    EG_CUDA_CHECK(cudaFuncSetAttribute(fused_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem(*dims)));
    if (decode_ws_smem(*dims) <= 232448)
      EG_CUDA_CHECK(cudaFuncSetAttribute(decode_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_ws_smem(*dims)));
    EG_CUDA_CHECK(cudaFuncSetAttribute(fused_regressor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRegSmem));
    int rc = motion_build_transposed(h, nullptr);
    if (rc) { delete h; return rc; }
    rc = mtc::create(&h->tc, *dims, h->w.data(), nullptr);
    if (rc) { mtc::destroy(h->tc); delete h; return rc; }
    EG_CUDA_CHECK(cudaDeviceSynchronize());
  }
  *out = h;
  return EG_OK;
}

extern "C" int eg_motion_refresh(EgMotion* h, void* stream) {
  EG_REQUIRE(h != nullptr, "null handle");
  if (!h->fused) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = motion_build_transposed(h, as_stream(stream));
  if (rc) return rc;
  return mtc::refresh(h->tc, h->w.data(), as_stream(stream));
}

extern "C" int eg_motion_set_fused(EgMotion* h, int fused) {
  EG_REQUIRE(h != nullptr, "null handle");
  EG_REQUIRE(!fused || h->wt != nullptr, "fused kernels unavailable for these dimensions");
  h->fused = fused ? 1 : 0;
  return EG_OK;
}

extern "C" void eg_motion_destroy(EgMotion* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->wt);
  float* bufs[] = {h->gi, h->gh, h->h, h->hx, h->c, h->t1, h->t2, h->rh, h->rbase, h->rt, h->xbc, h->hx2, h->yp2,
                   h->wx_pad, h->wd_pad, h->yp_pad, h->cin};
  for (auto p : bufs) cudaFree(p);
  cudaFree(h->dcnt);
  mtc::destroy(h->tc);
  delete h;
}

extern "C" int eg_motion_sample_prior(EgMotion* hd, const float* X, int ldx_env, int ldx_frame,
                                      const float* z, const float* betas, int B, float* Y, float* Yb,
                                      void* stream) {
  EG_REQUIRE(hd && X && z && betas && Y && Yb, "null pointer");
  EG_REQUIRE(B >= 0, "negative batch");
  if (B == 0) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(hd->device));
  int rc = motion_ws(hd, B);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const EgMotionDims& d = hd->d;
  const int D = d.in_dim, H = d.h_dim, Z = d.z_dim, Hm = d.mlp_dim, H3 = 3 * H;
  const float* const* w = hd->w.data();
#define EG_TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  const int ldY = 20 * D;
  const int Kin = H + Z + D;
  const bool tc_path = hd->fused && mtc::decode_available(hd->tc) && hd->wx_pad != nullptr;
  if (tc_path) {
    // every product of the prologue on the tensor-core dense layer: operands re-pitched to 16 bytes, and the three partial
    // products of the first GRUCell input term become ONE layer over the concatenated row [hx | z | y_0] (the reference's
    // torch.cat, models_GAMMA_primitive.py:93-94)
    const int DP = hd->DP, KP = hd->KP;
    const int per = 2 * DP + (KP - H);
    EG_LAUNCH(prologue_pack_kernel, (B * per + 255) / 256, 256, 0, st, X, ldx_env, ldx_frame, z, B, D, DP, H, Z, KP, Y, hd->yp_pad, hd->cin);
    EG_TRY(linear(st, hd->yp_pad, 2 * DP, B, hd->wx_pad, DP, w[P_XENC_BIH], D, H3, hd->gi, H3));
    EG_TRY(launch_gru_gate(st, hd->gi, nullptr, w[P_XENC_BHH], nullptr, hd->h, B, H, H));
    EG_TRY(linear(st, hd->yp_pad + DP, 2 * DP, B, hd->wx_pad, DP, w[P_XENC_BIH], D, H3, hd->gi, H3));
    EG_TRY(linear(st, hd->h, H, B, w[P_XENC_WHH], H, w[P_XENC_BHH], H, H3, hd->gh, H3));
    EG_TRY(launch_gru_gate(st, hd->gi, hd->gh, nullptr, hd->h, hd->cin, B, H, KP));            // hx -> cin[:, :H]
    EG_TRY(linear(st, hd->cin, KP, B, w[P_DRNN_W0], H, w[P_DRNN_B0], H, Hm, hd->t1, Hm, ACT_TANH));
    EG_TRY(linear(st, hd->t1, Hm, B, w[P_DRNN_W1], Hm, w[P_DRNN_B1], Hm, H, hd->t2, H, ACT_TANH));
    EG_TRY(linear(st, hd->t2, H, B, w[P_DRNN_W2], H, w[P_DRNN_B2], H, H, hd->h, H, ACT_TANH));
    EG_TRY(linear(st, hd->cin, KP, B, hd->wd_pad, KP, w[P_DRNN_BIH], Kin, H3, hd->c, H3));      // gi_1 = b_ih + [hx, z, y_0] W_ih^T
    EG_TRY(mtc::decode(hd->tc, hd->c, hd->h, Y, B, st));
  } else {
  EG_LAUNCH(copy_history_kernel, (B * 2 * D + 255) / 256, 256, 0, st, X, ldx_env, ldx_frame, B, D, Y);
  // ---- x_enc GRU over the 2 history frames (h0 = 0) ----
  EG_TRY(linear(st, Y, ldY, B, w[P_XENC_WIH], D, w[P_XENC_BIH], D, H3, hd->gi, H3));
  EG_TRY(launch_gru_gate(st, hd->gi, nullptr, w[P_XENC_BHH], nullptr, hd->h, B, H, H));
  EG_TRY(linear(st, Y + D, ldY, B, w[P_XENC_WIH], D, w[P_XENC_BIH], D, H3, hd->gi, H3));
  EG_TRY(linear(st, hd->h, H, B, w[P_XENC_WHH], H, w[P_XENC_BHH], H, H3, hd->gh, H3));
  EG_TRY(launch_gru_gate(st, hd->gi, hd->gh, nullptr, hd->h, hd->hx, B, H, H));
  // ---- h_rnn = drnn_mlp(hx) ----
  EG_TRY(linear(st, hd->hx, H, B, w[P_DRNN_W0], H, w[P_DRNN_B0], H, Hm, hd->t1, Hm, ACT_TANH));
  EG_TRY(linear(st, hd->t1, Hm, B, w[P_DRNN_W1], Hm, w[P_DRNN_B1], Hm, H, hd->t2, H, ACT_TANH));
  EG_TRY(linear(st, hd->t2, H, B, w[P_DRNN_W2], H, w[P_DRNN_B2], H, H, hd->h, H, ACT_TANH));
  // ---- step-invariant part of the GRUCell input: [hx, z] W_ih[:, :H+Z]^T + b_ih ----
  EG_TRY(linear(st, hd->hx, H, B, w[P_DRNN_WIH], Kin, w[P_DRNN_BIH], H, H3, hd->c, H3));
  EG_TRY(linear(st, z, Z, B, w[P_DRNN_WIH] + H, Kin, nullptr, Z, H3, hd->c, H3, ACT_NONE, 0.f, nullptr, 0, 1));
  if (hd->decode_ws < 0) {
    const char* e = getenv("EG_DECODE_WS");
    hd->decode_ws = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (hd->fused && mtc::decode_available(hd->tc)) {
    // tcgen05 decode (motion_tc.cu): first-step input term gi_1 = c + y_0 Wy^T, then one 16-CTA cluster per 128 rows
    EG_TRY(linear(st, Y + D, ldY, B, w[P_DRNN_WIH] + H + Z, Kin, nullptr, D, H3, hd->c, H3, ACT_NONE, 0.f, nullptr, 0, 1));
    EG_TRY(mtc::decode(hd->tc, hd->c, hd->h, Y, B, st));
  } else if (hd->fused && hd->decode_ws && B >= 64 && H == 256 && Hm == 512 && D == 201) {
    // weight-stationary 2-D decode: 18 column-slice CTAs per 32-row block (see decode_ws_kernel)
    const int n_rb = (B + dws::RB - 1) / dws::RB, D4 = (D + 3) & ~3;
    EG_CUDA_CHECK(cudaMemsetAsync(hd->dcnt, 0, n_rb * sizeof(unsigned), st));
    DwsArgs da{hd->dw, hd->c, hd->h, Y, {hd->hx2, hd->hx2 + (size_t)B * H}, hd->t1, hd->t2,
               {hd->yp2, hd->yp2 + (size_t)B * D4}, hd->dcnt, B, D, H, Hm};
    EG_LAUNCH(decode_ws_kernel, dim3(dws::CS, n_rb), dws::THREADS, decode_ws_smem(d), st, da);
  } else if (hd->fused) {
    EG_LAUNCH(fused_decode_kernel, (B + DR - 1) / DR, DEC_THREADS, decode_smem(d), st, hd->dw, hd->c, hd->h, Y, B, D, H, Hm);
  } else
  for (int i = 0; i < 18; ++i) {
    const float* yp = Y + (1 + i) * D;     // previous frame (history frame 1 for i == 0)
    float* yo = Y + (2 + i) * D;
    EG_TRY(linear(st, yp, ldY, B, w[P_DRNN_WIH] + H + Z, Kin, nullptr, D, H3, hd->gi, H3, ACT_NONE, 0.f, hd->c, H3));
    EG_TRY(linear(st, hd->h, H, B, w[P_DRNN_WHH], H, w[P_DRNN_BHH], H, H3, hd->gh, H3));
    EG_TRY(launch_gru_gate(st, hd->gi, hd->gh, nullptr, hd->h, hd->h, B, H, H));
    EG_TRY(linear(st, hd->h, H, B, w[P_DMLP_W0], H, w[P_DMLP_B0], H, Hm, hd->t1, Hm, ACT_TANH));
    EG_TRY(linear(st, hd->t1, Hm, B, w[P_DMLP_W1], Hm, w[P_DMLP_B1], Hm, H, hd->t2, H, ACT_TANH));
    EG_TRY(linear(st, hd->t2, H, B, w[P_DOUT_W], H, w[P_DOUT_B], H, D, yo, ldY, ACT_NONE, 0.f, yp, ldY));
  }
  }   // !tc_path
  // ---- regressor over all B*20 marker frames (frames 0,1 are computed and discarded) ----
  const int M = B * 20, Hr = d.reg_h, BD = d.body_dim, Kr = D + BD + 10;
  if (hd->fused && mtc::regress_available(hd->tc)) {
    EG_TRY(mtc::regress(hd->tc, Y, betas, B, hd->xbc, st));
    EG_LAUNCH(regressor_tail_kernel, (M * 32 + 127) / 128, 128, 0, st, hd->xbc, M, 20, 2, Yb);
    return EG_OK;
  }
  if (hd->fused) {
    EG_LAUNCH(fused_regressor_kernel, (M + RR - 1) / RR, RWARPS * 32, kRegSmem, st, hd->rw, Y, betas, 20, M, d.reg_blocks,
              d.reg_recur, hd->xbc);
    EG_LAUNCH(regressor_tail_kernel, (M * 32 + 127) / 128, 128, 0, st, hd->xbc, M, 20, 2, Yb);
    return EG_OK;
  }
  const float* Win = w[R_IN_W];
  EG_TRY(linear(st, Y, D, M, Win, Kr, w[R_IN_B], D, Hr, hd->rbase, Hr));
  EG_TRY(linear(st, betas, 10, M, Win + D + BD, Kr, nullptr, 10, Hr, hd->rbase, Hr, ACT_NONE, 0.f, nullptr, 0, 1, 20));
  const float* const* wb = w + R_BLOCKS;
  const float* Wout = wb[d.reg_blocks * 4];
  const float* Bout = wb[d.reg_blocks * 4 + 1];
  for (int r = 0; r < d.reg_recur; ++r) {
    const float* hcur = hd->rbase;
    if (r > 0) {   // h = base + xb W_in[:, 201:360]^T
      EG_TRY(linear(st, hd->xbc, BD, M, Win + D, Kr, nullptr, BD, Hr, hd->rh, Hr, ACT_NONE, 0.f, hd->rbase, Hr));
      hcur = hd->rh;
    }
    for (int k = 0; k < d.reg_blocks; ++k) {
      EG_TRY(linear(st, hcur, Hr, M, wb[k * 4], Hr, wb[k * 4 + 1], Hr, Hr, hd->rt, Hr, ACT_RELU));
      EG_TRY(linear(st, hd->rt, Hr, M, wb[k * 4 + 2], Hr, wb[k * 4 + 3], Hr, Hr, hd->rh, Hr, ACT_RELU, 0.f, hcur, Hr));
      hcur = hd->rh;
    }
    // xb = out_fc(h) + xb   (xb starts at zero)
    EG_TRY(linear(st, hcur, Hr, M, Wout, Hr, Bout, Hr, BD, hd->xbc, BD, ACT_NONE, 0.f, r > 0 ? hd->xbc : nullptr, BD));
  }
  EG_LAUNCH(regressor_tail_kernel, (M * 32 + 127) / 128, 128, 0, st, hd->xbc, M, 20, 2, Yb);
#undef EG_TRY
  return EG_OK;
}

// ------------------------------------------------------------------------------------------
// VPoser v1 encoder (.loc)
// ------------------------------------------------------------------------------------------
struct EgVposer {
  int device = 0;
  std::vector<const float*> w;   // bn1.{weight,bias,running_mean,running_var}, fc1.{w,b}, bn2.{...}, fc2.{w,b}, mu.{w,b}
  int cap_M = 0;
  float *a = nullptr, *b = nullptr;
};

extern "C" int eg_vposer_create(const void* const* weights_host, int n_weights, int device, EgVposer** out) {
  EG_REQUIRE(weights_host && out && n_weights == 14, "expected 14 weight tensors");
  EgVposer* h = new EgVposer();
  h->device = device;
  h->w.assign((const float* const*)weights_host, (const float* const*)weights_host + n_weights);
  *out = h;
  return EG_OK;
}

extern "C" void eg_vposer_destroy(EgVposer* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->a); cudaFree(h->b);
  delete h;
}

extern "C" int eg_vposer_encode(EgVposer* h, const float* x, int ldx, int M, float* loc, void* stream) {
  EG_REQUIRE(h && x && loc && M >= 0, "bad arguments");
  if (M == 0) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  if (M > h->cap_M) {
    cudaFree(h->a); cudaFree(h->b); h->a = h->b = nullptr; h->cap_M = 0;
    EG_CUDA_CHECK(cudaMalloc((void**)&h->a, (size_t)M * 512 * 4));
    EG_CUDA_CHECK(cudaMalloc((void**)&h->b, (size_t)M * 512 * 4));
    h->cap_M = M;
  }
  cudaStream_t st = as_stream(stream);
  const float* const* w = h->w.data();
  int rc;
  EG_LAUNCH(bn_eval_kernel, ew_grid((int64_t)M * 63), 256, 0, st, x, ldx, M, 63, w[0], w[1], w[2], w[3], 1e-5f, h->a);
  if ((rc = linear(st, h->a, 63, M, w[4], 63, w[5], 63, 512, h->b, 512, ACT_LRELU, 0.2f))) return rc;
  EG_LAUNCH(bn_eval_kernel, ew_grid((int64_t)M * 512), 256, 0, st, h->b, 512, M, 512, w[6], w[7], w[8], w[9], 1e-5f, h->a);
  if ((rc = linear(st, h->a, 512, M, w[10], 512, w[11], 512, 512, h->b, 512, ACT_LRELU, 0.2f))) return rc;
  if ((rc = linear(st, h->b, 512, M, w[12], 512, w[13], 512, 32, loc, 32))) return rc;
  return EG_OK;
}

// generic product C[M,N] = op(A) op(B) in the library's three backward / forward layouts (GemmArgs in nn.cuh), exposed for
// tests: trans_a: A element (m,k) at A[k*lda+m]; trans_b: B element (k,n) at B[n*ldb+k] (nn.Linear weight)
extern "C" int eg_matmul(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, int M, int N, int K,
                         float* Cm, int ldc, int accumulate, void* stream) {
  EG_REQUIRE(A && B && Cm && M >= 0 && N >= 0 && K > 0, "bad arguments");
  GemmArgs g{A, lda, 1, B, ldb, Cm, ldc, nullptr, nullptr, 0, M, N, K, ACT_NONE, 0.0f, accumulate ? 1 : 0, 1.0f};
  return launch_gemm(g, trans_a != 0, trans_b != 0, as_stream(stream));
}

// generic dense layer exposed for tests and host-side composition
extern "C" int eg_linear_forward(const float* x, int ldx, int M, const float* W, const float* b, int in_dim,
                                 int out_dim, int act, float slope, const float* residual, int ldr, float* y,
                                 int ldy, void* stream) {
  EG_REQUIRE(x && W && y && M >= 0 && in_dim > 0 && out_dim > 0, "bad arguments");
  return linear(as_stream(stream), x, ldx, M, W, in_dim, b, in_dim, out_dim, y, ldy, act, slope, residual, ldr);
}
