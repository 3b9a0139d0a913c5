// PPO policy for the crowd_ppo path: actor-critic forward, hand-written backward, PPO loss,
// GAE, gradient clipping and AdamW - replaces
//   GAMMAPolicyBase.forward / GAMMAActor / GAMMACritic   (motion/models/models_policy_ppo.py:287-350)
//   GAMMAPPOPolicy.forward / _compute_returns / learn    (motion/crowd_ppo/ppo_policy.py:105-265)
//   tianshou 0.5.0 _gae_return / compute_episodic_return (third-party)
//   torch.optim.AdamW + clip_grad_norm_                  (main_ppo.py:134, ppo_policy.py:243-247)
//
// Parameters, gradients and both Adam moments live in four flat fp32 buffers owned by the caller in
// the order of ActorCritic(actor, critic, shared_net).parameters(): one NCCL allreduce covers the whole
// gradient, one kernel applies AdamW, and the clip-grad-norm range (actor + critic only - the
// reference's `_actor_critic` quirk) is a prefix of the buffer.
#include <stdlib.h>

#include <vector>

#include "nn.cuh"
#include "geom.cuh"

namespace eg {

struct Lin { int64_t w, b; int in, out; };   // offsets into the flat buffers

struct PolicyLayout {
  Lin a_blk[4][2], a_out, c_blk[4][2], c_out;
  int64_t x_wih, x_whh, x_bih, x_bhh, e_wih, e_whh, e_bih, e_bhh;
  int64_t n_actor_critic, n_total;
  int hx_dim;
  int n_tensors;
  int64_t tensor_off[64];      // offset of every parameter tensor, in ActorCritic(actor, critic, shared_net).parameters() order
};

// Every tensor starts on a 16-byte boundary of the flat buffers (the critic's 1-element output bias would otherwise
// leave every later matrix misaligned for TMA): the padding elements are zero parameters with zero gradients.
static PolicyLayout make_layout(const EgPolicyDims& d) {
  PolicyLayout L{};
  const int D = 2 * d.h_dim + 4 * d.pe_L;    // 1152
  L.hx_dim = D;
  int64_t off = 0;
  int nt = 0;
  auto take = [&](int64_t n) { off = (off + 3) & ~(int64_t)3; const int64_t o = off; L.tensor_off[nt++] = o; off += n; return o; };
  auto lin = [&](int in, int out) { Lin l{}; l.in = in; l.out = out; l.w = take((int64_t)in * out); l.b = take(out); return l; };
  for (int k = 0; k < d.n_blocks; ++k) { L.a_blk[k][0] = lin(D, D); L.a_blk[k][1] = lin(D, D); }
  L.a_out = lin(D, 2 * d.z_dim);
  for (int k = 0; k < d.n_blocks; ++k) { L.c_blk[k][0] = lin(D, D); L.c_blk[k][1] = lin(D, D); }
  L.c_out = lin(D, 1);
  off = (off + 3) & ~(int64_t)3;
  L.n_actor_critic = off;
  const int H = d.h_dim, H3 = 3 * d.h_dim;
  L.x_wih = take((int64_t)H3 * d.in_dim);
  L.x_whh = take((int64_t)H3 * H);
  L.x_bih = take(H3);
  L.x_bhh = take(H3);
  L.e_wih = take((int64_t)H3 * d.ego_dim);
  L.e_whh = take((int64_t)H3 * H);
  L.e_bih = take(H3);
  L.e_bhh = take(H3);
  L.n_total = (off + 3) & ~(int64_t)3;
  L.n_tensors = nt;
  return L;
}

// positional_encoding(x, L): [sin(x f0), cos(x f0), sin(x f1), ...], f_k = 2^k (models_policy_ppo.py:276-285)
__global__ void __launch_bounds__(128)
pe_kernel(const float* __restrict__ dist, const float* __restrict__ time, int B, int L, int ld, int off,
          float* __restrict__ hx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * L) return;
  const int b = i / (2 * L), r = i % (2 * L);
  const int which = r / L, k = r % L;
  const float x = which == 0 ? dist[b] : time[b];
  const float a = x * exp2f((float)k);       // exact power-of-two scaling, like x * freq in fp32
  float* o = hx + (int64_t)b * ld + off + which * 2 * L + 2 * k;
  o[0] = sinf(a);                            // full-range sinf/cosf (no fast-math): arguments reach 2^31
  o[1] = cosf(a);
}

__global__ void __launch_bounds__(256)
lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope, int64_t n,
                 float* __restrict__ dx) {
  eg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.0f ? dy[i] : dy[i] * slope;
}

__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ c) {
  eg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c[i] = a[i] + b[i];
}

// db[n] += sum_m dY[m, n]: a cluster of 4 CTAs per 32-column block (each a quarter of the rows, 8 row lanes), partial sums
// added by rank 0 in rank order through distributed shared memory (deterministic)
__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dY, int ld, int M, int N, float* __restrict__ db) {
  __shared__ float part[8][33];
  __shared__ float csum[32];
  eg_pdl_enter();
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cl;
  const int rows_per = (M + 3) / 4, r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.0f;
  if (col < N)
#pragma unroll 4
    for (int m = r0 + rl; m < r1; m += 8) s += dY[(int64_t)m * ld + col];
  part[rl][cl] = s;
  __syncthreads();
  if (rl == 0) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][cl];
    csum[cl] = t;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (blockIdx.y == 0 && rl == 0 && col < N) {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&csum[cl]);
    float t = 0.0f;
#pragma unroll
    for (uint32_t r = 0; r < 4; ++r) {
      uint32_t remote; float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
      t += v;
    }
    db[col] += t;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // remote reads done
}

// dx = dy * lrelu'(y) and, in the same pass, db[n] += sum_m dx[m, n] (the bias gradient of the layer that produced y), as
// a cluster of 4 CTAs per 32-column block: each CTA walks a quarter of the rows with 128-byte coalesced row segments (144
// CTAs for the 1152-wide layers; the first version used 72 CTAs on 64-byte segments and took 12 us), the four partial column
// sums meet in rank 0 through distributed shared memory and are added in rank order - deterministic.
__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(256)
lrelu_bwd_colsum_cl_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope, int M, int N,
                           float* __restrict__ dx, float* __restrict__ db) {
  __shared__ float part[8][33];
  __shared__ float csum[32];
  eg_pdl_enter();
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cl;
  const int rows_per = (M + 3) / 4, r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.0f;
  if (col < N)
#pragma unroll 4
    for (int m = r0 + rl; m < r1; m += 8) {
      const int64_t i = (int64_t)m * N + col;
      const float g = y[i] > 0.0f ? dy[i] : dy[i] * slope;
      dx[i] = g;
      s += g;
    }
  part[rl][cl] = s;
  __syncthreads();
  if (rl == 0) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][cl];
    csum[cl] = t;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (blockIdx.y == 0 && rl == 0 && col < N) {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&csum[cl]);
    float t = 0.0f;
#pragma unroll
    for (uint32_t r = 0; r < 4; ++r) {
      uint32_t remote; float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
      t += v;
    }
    db[col] += t;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // remote reads done
}

// y[m] = act(x[m, :] . w + b): the N = 1 layer (critic head) as one warp per row
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ x, int ldx, int M, int K, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ y, int ldy) {
  eg_pdl_enter();
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float s = 0.0f;
  for (int k = lane; k < K; k += 32) s = fmaf(x[(int64_t)m * ldx + k], __ldg(w + k), s);
  s = warp_sum(s);
  if (lane == 0) y[(int64_t)m * ldy] = s + (b ? __ldg(b) : 0.0f);
}

// GRU cell backward given dh (grad of the cell output); writes dgi, dgh [M,3H] and dh_prev = dh * z
__global__ void __launch_bounds__(256)
gru_bwd_kernel(const float* __restrict__ dh, int ld_dh, const float* __restrict__ r, const float* __restrict__ z,
               const float* __restrict__ n, const float* __restrict__ ghn, const float* __restrict__ h_prev,
               int M, int H, float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh_prev) {
  eg_pdl_enter();
  const int64_t total = (int64_t)M * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / H), j = (int)(i % H);
    const float g = dh[(int64_t)m * ld_dh + j];
    const float rr = r[i], zz = z[i], nn = n[i];
    const float hp = h_prev ? h_prev[i] : 0.0f;
    const float dn = g * (1.0f - zz);
    const float dz = g * (hp - nn);
    const float dpre_n = dn * (1.0f - nn * nn);
    const float dr = dpre_n * ghn[i];
    const float dpre_z = dz * zz * (1.0f - zz);
    const float dpre_r = dr * rr * (1.0f - rr);
    float* gi = dgi + (int64_t)m * 3 * H;
    float* gh = dgh + (int64_t)m * 3 * H;
    gi[j] = dpre_r; gi[H + j] = dpre_z; gi[2 * H + j] = dpre_n;
    gh[j] = dpre_r; gh[H + j] = dpre_z; gh[2 * H + j] = dpre_n * rr;
    if (dh_prev) dh_prev[i] = g * zz;
  }
}

// Diagonal-Gaussian head (ppo_policy.py:168-179): logvar clamp, sigma = exp(logvar)^0.5,
// act = mu + sigma * eps (or mu when deterministic), log_prob summed over the 128 latent dims.
__global__ void __launch_bounds__(128)
gauss_sample_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int ld,
                    const float* __restrict__ eps, int B, int Z, float min_lv, float max_lv,
                    float* __restrict__ act, float* __restrict__ logp) {
  const int b = blockIdx.x;
  float lp = 0.0f;
  for (int d = threadIdx.x; d < Z; d += blockDim.x) {
    const float m = mu[(int64_t)b * ld + d];
    const float lv = fminf(fmaxf(logvar[(int64_t)b * ld + d], min_lv), max_lv);
    const float sigma = sqrtf(expf(lv));
    const float a = eps ? m + sigma * eps[(int64_t)b * Z + d] : m;
    act[(int64_t)b * Z + d] = a;
    const float df = a - m;
    lp += -(df * df) / (2.0f * sigma * sigma) - logf(sigma) - 0.91893853320467274178f;
  }
  lp = warp_sum(lp);
  __shared__ float part[4];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x == 0 && logp) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    logp[b] = t;
  }
}

// PPO loss head (ppo_policy.py:191-240): per-row clipped surrogate, value MSE, entropy; writes the
// gradients w.r.t. the raw actor output [mu | logvar] and the critic output. stats (device, fp32[8]):
// 0 clip_loss, 1 vf_loss, 2 ent_loss, 3 kld (0.5 mean mu^2), 4 approx-kl mean(logp_old - logp_new).
__global__ void __launch_bounds__(128)
ppo_head_kernel(const float* __restrict__ out_a, const float* __restrict__ value, const float* __restrict__ act,
                const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ ret,
                int B, int Z, float inv_B, float eps_clip, float vf_coef, float ent_coef, float min_lv, float max_lv,
                float* __restrict__ d_out_a, float* __restrict__ d_value, float* __restrict__ stats) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* oa = out_a + (int64_t)b * 2 * Z;
  __shared__ float part[4][3];
  __shared__ float s_g;
  float lp = 0.f, ent = 0.f, mu2 = 0.f;
  for (int d = tid; d < Z; d += blockDim.x) {
    const float m = oa[d];
    const float lv = fminf(fmaxf(oa[Z + d], min_lv), max_lv);
    const float sigma = sqrtf(expf(lv));
    const float df = act[(int64_t)b * Z + d] - m;
    lp += -(df * df) / (2.0f * sigma * sigma) - logf(sigma) - 0.91893853320467274178f;
    ent += 0.5f + 0.91893853320467274178f + logf(sigma);
    mu2 += m * m;
  }
  lp = warp_sum(lp); ent = warp_sum(ent); mu2 = warp_sum(mu2);
  if ((tid & 31) == 0) { part[tid >> 5][0] = lp; part[tid >> 5][1] = ent; part[tid >> 5][2] = mu2; }
  __syncthreads();
  if (tid == 0) {
    lp = ent = mu2 = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { lp += part[w][0]; ent += part[w][1]; mu2 += part[w][2]; }
    const float ratio = expf(lp - logp_old[b]);
    const float a = adv[b];
    const float surr1 = ratio * a;
    const float rc = fminf(fmaxf(ratio, 1.0f - eps_clip), 1.0f + eps_clip);
    const float surr2 = rc * a;
    const bool inside = ratio >= 1.0f - eps_clip && ratio <= 1.0f + eps_clip;
    // d(-min(surr1,surr2))/d logp ; torch.min splits ties evenly and clamp passes grad inside the range
    float g = 0.0f;
    if (surr1 < surr2) g = -a * ratio;
    else if (surr1 == surr2) g = -a * ratio * (inside ? 1.0f : 0.5f);
    else g = inside ? -a * ratio : 0.0f;
    s_g = g * inv_B;
    const float v = value[b];
    const float dv = v - ret[b];
    d_value[b] = vf_coef * 2.0f * dv * inv_B;
    atomicAdd(stats + 0, -fminf(surr1, surr2) * inv_B);
    atomicAdd(stats + 1, dv * dv * inv_B);
    atomicAdd(stats + 2, ent * inv_B);
    atomicAdd(stats + 3, 0.5f * mu2 * inv_B / (float)Z);
    atomicAdd(stats + 4, (logp_old[b] - lp) * inv_B);
  }
  __syncthreads();
  const float g = s_g;
  for (int d = tid; d < Z; d += blockDim.x) {
    const float m = oa[d];
    const float raw = oa[Z + d];
    const float lv = fminf(fmaxf(raw, min_lv), max_lv);
    const float var = expf(lv);
    const float df = act[(int64_t)b * Z + d] - m;
    d_out_a[(int64_t)b * 2 * Z + d] = g * df / var;
    const bool pass = raw >= min_lv && raw <= max_lv;
    // d logp / d lv = 0.5 (df^2 / var - 1);  d ent / d lv = 0.5
    d_out_a[(int64_t)b * 2 * Z + Z + d] = pass ? (g * 0.5f * (df * df / var - 1.0f) - ent_coef * inv_B * 0.5f) : 0.0f;
  }
}

// sum / sum of squares of x[0..n) into out[0], out[1] (double accumulation per block, fp32 atomics avoided)
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  double s = 0.0, q = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    s += v; q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double ps[8], pq[8];
  if ((threadIdx.x & 31) == 0) { ps[threadIdx.x >> 5] = s; pq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = q = 0.0;
    for (int w = 0; w < 8; ++w) { s += ps[w]; q += pq[w]; }
    atomicAdd(out, s);
    atomicAdd(out + 1, q);
  }
}

// adv_n = (adv - mean) / (std + eps) with the unbiased std from global moments {sum, sumsq, count}
__global__ void __launch_bounds__(256)
adv_normalize_kernel(const float* __restrict__ adv, int n, const double* __restrict__ mom, float eps,
                     float* __restrict__ out) {
  const double cnt = mom[2];
  const double mean = mom[0] / cnt;
  const double var = (mom[1] - cnt * mean * mean) / (cnt - 1.0);
  const float m = (float)mean, sd = (float)sqrt(var > 0.0 ? var : 0.0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (adv[i] - m) / (sd + eps);
}

// clip_grad_norm_ over g[0..n_clip) (scale from the device-side norm) fused with AdamW over all n params
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, int64_t n_clip, const double* __restrict__ sumsq, float max_norm, float lr, float beta1,
             float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
  float clip = 1.0f;
  if (max_norm > 0.0f) {
    const float total = (float)sqrt(sumsq[1]);
    clip = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (i < n_clip) gi *= clip;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// tianshou _gae_return per env trajectory segment (float64 like the numba kernel). Layout [T,E] time-major.
__global__ void __launch_bounds__(128)
gae_kernel(const float* __restrict__ v_s, const float* __restrict__ v_next, const float* __restrict__ rew,
           const uint8_t* __restrict__ terminated, const uint8_t* __restrict__ end_flag, int T, int E, double gamma,
           double lam, float* __restrict__ adv, float* __restrict__ ret) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double gae = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const int i = t * E + e;
    const double vs = (double)v_s[i];
    // v_s_ = v_next * value_mask (~terminated); computed in float32 numpy like the reference, then promoted
    const double vn = (double)(terminated[i] ? v_next[i] * 0.0f : v_next[i]);
    const double delta = (double)rew[i] + vn * gamma - vs;
    const double disc = (1.0 - (end_flag[i] ? 1.0 : 0.0)) * (gamma * lam);
    gae = delta + disc * gae;
    adv[i] = (float)gae;
    ret[i] = (float)(gae + vs);
  }
}

static inline int ew_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, kNumSMs * 8); }

}  // namespace eg

using namespace eg;

struct EgPolicy {
  int device = 0;
  EgPolicyDims d;
  PolicyLayout L;
  float *P = nullptr, *G = nullptr;    // flat params / grads (caller-owned)
  int cap = 0;
  // saved activations
  float *gi = nullptr, *gh = nullptr;                    // [B,3H] scratch
  float *xr[2], *xz[2], *xn[2], *xg[2], *xh1 = nullptr;  // x_enc saves per step
  float *er[2], *ez[2], *en[2], *eg_[2], *eh1 = nullptr; // ego_enc saves
  float* hx = nullptr;                                   // [B,1152]
  float *a_in[5], *a_t[4], *a_u[4], *c_in[5], *c_t[4], *c_u[4];
  float *out_a = nullptr, *out_c = nullptr;
  // backward scratch
  float *d_out_a = nullptr, *d_out_c = nullptr, *dh = nullptr, *da = nullptr, *dt = nullptr, *dhx = nullptr,
        *dgi = nullptr, *dgh = nullptr, *dh1 = nullptr, *dh1b = nullptr;
  double* mom = nullptr;                                 // [4] device scalars
  std::vector<float*> owned;
  // x_enc operands with a 16-byte row pitch (in_dim = 402 floats is not): padded copies so the products run on gemm_tc
  int in_pad = 0;                                        // in_dim rounded up to 4
  float *xpad = nullptr, *wih_pad = nullptr;             // [B,2,in_pad], [3H,in_pad]
  // actor and critic chains are independent between the shared encoder and the GRU backward: the critic runs on a side
  // stream so that two latency-bound layer chains share the SMs
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int two_streams = 1;
  float *dh2 = nullptr, *da2 = nullptr, *dt2 = nullptr, *dhx2 = nullptr;   // critic-chain backward scratch
};

#define EG_TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static int policy_ws(EgPolicy* h, int B) {
  if (B <= h->cap) return EG_OK;
  for (float* p : h->owned) cudaFree(p);
  h->owned.clear(); h->cap = 0;
  const int H = h->d.h_dim, H3 = 3 * H, D = h->L.hx_dim, Z2 = 2 * h->d.z_dim;
  auto alloc = [&](float** p, size_t n) -> int {
    EG_CUDA_CHECK(cudaMalloc((void**)p, n * sizeof(float)));
    h->owned.push_back(*p);
    return EG_OK;
  };
  const size_t b = (size_t)B;
  EG_TRY(alloc(&h->gi, b * H3)); EG_TRY(alloc(&h->gh, b * H3));
  for (int s = 0; s < 2; ++s) {
    EG_TRY(alloc(&h->xr[s], b * H)); EG_TRY(alloc(&h->xz[s], b * H)); EG_TRY(alloc(&h->xn[s], b * H)); EG_TRY(alloc(&h->xg[s], b * H));
    EG_TRY(alloc(&h->er[s], b * H)); EG_TRY(alloc(&h->ez[s], b * H)); EG_TRY(alloc(&h->en[s], b * H)); EG_TRY(alloc(&h->eg_[s], b * H));
  }
  EG_TRY(alloc(&h->xh1, b * H)); EG_TRY(alloc(&h->eh1, b * H)); EG_TRY(alloc(&h->hx, b * D));
  for (int k = 0; k <= h->d.n_blocks; ++k) {
    if (k == 0) { h->a_in[0] = h->hx; h->c_in[0] = h->hx; }
    else { EG_TRY(alloc(&h->a_in[k], b * D)); EG_TRY(alloc(&h->c_in[k], b * D)); }
    if (k < h->d.n_blocks) {
      EG_TRY(alloc(&h->a_t[k], b * D)); EG_TRY(alloc(&h->a_u[k], b * D));
      EG_TRY(alloc(&h->c_t[k], b * D)); EG_TRY(alloc(&h->c_u[k], b * D));
    }
  }
  EG_TRY(alloc(&h->out_a, b * Z2)); EG_TRY(alloc(&h->out_c, b));
  EG_TRY(alloc(&h->d_out_a, b * Z2)); EG_TRY(alloc(&h->d_out_c, b));
  EG_TRY(alloc(&h->dh, b * D)); EG_TRY(alloc(&h->da, b * D)); EG_TRY(alloc(&h->dt, b * D)); EG_TRY(alloc(&h->dhx, b * D));
  EG_TRY(alloc(&h->dgi, b * H3)); EG_TRY(alloc(&h->dgh, b * H3)); EG_TRY(alloc(&h->dh1, b * H)); EG_TRY(alloc(&h->dh1b, b * H));
  EG_TRY(alloc(&h->dh2, b * D)); EG_TRY(alloc(&h->da2, b * D)); EG_TRY(alloc(&h->dt2, b * D)); EG_TRY(alloc(&h->dhx2, b * D));
  EG_TRY(alloc(&h->xpad, b * 2 * h->in_pad));
  h->cap = B;
  return EG_OK;
}

extern "C" int64_t eg_policy_param_count(const EgPolicyDims* d, int64_t* n_actor_critic) {
  if (!d) return -1;
  PolicyLayout L = make_layout(*d);
  if (n_actor_critic) *n_actor_critic = L.n_actor_critic;
  return L.n_total;
}

extern "C" int eg_policy_param_offsets(const EgPolicyDims* d, int64_t* offsets, int n_max) {
  if (!d || !offsets) return -1;
  PolicyLayout L = make_layout(*d);
  if (L.n_tensors > n_max) return -1;
  for (int i = 0; i < L.n_tensors; ++i) offsets[i] = L.tensor_off[i];
  return L.n_tensors;
}

extern "C" int eg_policy_create(const EgPolicyDims* dims, float* params_flat, float* grads_flat, int device,
                                EgPolicy** out) {
  EG_REQUIRE(dims && params_flat && out, "null pointer");
  EG_REQUIRE(dims->n_blocks >= 1 && dims->n_blocks <= 4, "n_blocks must be in [1,4]");
  EG_CUDA_CHECK(cudaSetDevice(device));
  EgPolicy* h = new EgPolicy();
  h->device = device; h->d = *dims; h->L = make_layout(*dims); h->P = params_flat; h->G = grads_flat;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->mom, 4 * sizeof(double)));
  h->in_pad = (dims->in_dim + 3) & ~3;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->wih_pad, (size_t)3 * dims->h_dim * h->in_pad * sizeof(float)));
  EG_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  EG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  EG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  const char* e = getenv("EG_POLICY_TWO_STREAMS");
  h->two_streams = (e != nullptr && e[0] == '0') ? 0 : 1;
  *out = h;
  return EG_OK;
}

extern "C" void eg_policy_destroy(EgPolicy* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (float* p : h->owned) cudaFree(p);
  cudaFree(h->mom);
  cudaFree(h->wih_pad);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

// one GRU encoder over the 2 observation frames (h0 = 0), saving gate activations for backward
static int gru2_forward(EgPolicy* h, cudaStream_t st, const float* x, int ld_env, int ld_frame, int in_dim, int B,
                        const float* Wih, int ld_wih, int64_t whh, int64_t bih, int64_t bhh, float** r, float** z, float** n, float** g,
                        float* h1, float* out, int ld_out) {
  const int H = h->d.h_dim, H3 = 3 * H;
  const float* P = h->P;
  EG_TRY(linear(st, x, ld_env, B, Wih, ld_wih, P + bih, in_dim, H3, h->gi, H3));
  EG_TRY(launch_gru_gate(st, h->gi, nullptr, P + bhh, nullptr, h1, B, H, H, r[0], z[0], n[0], g[0]));
  EG_TRY(linear(st, x + ld_frame, ld_env, B, Wih, ld_wih, P + bih, in_dim, H3, h->gi, H3));
  EG_TRY(linear(st, h1, H, B, P + whh, H, P + bhh, H, H3, h->gh, H3));
  EG_TRY(launch_gru_gate(st, h->gi, h->gh, nullptr, h1, out, B, H, ld_out, r[1], z[1], n[1], g[1]));
  return EG_OK;
}

static int mlp_block_forward(EgPolicy* h, cudaStream_t st, const Lin blk[][2], const Lin& outl, float** in, float** t,
                             float** u, float* out, int B) {
  const int D = h->L.hx_dim;
  const float* P = h->P;
  for (int k = 0; k < h->d.n_blocks; ++k) {
    EG_TRY(linear(st, in[k], D, B, P + blk[k][0].w, D, P + blk[k][0].b, D, D, t[k], D, ACT_LRELU, 0.01f));
    EG_TRY(linear(st, t[k], D, B, P + blk[k][1].w, D, P + blk[k][1].b, D, D, u[k], D, ACT_LRELU, 0.01f));
    EG_LAUNCH_PDL(add_kernel, ew_grid((int64_t)B * D), 256, 0, st, u[k], in[k], (int64_t)B * D, in[k + 1]);
  }
  if (outl.out == 1) EG_LAUNCH_PDL(gemv_rows_kernel, (B + 7) / 8, 256, 0, st, in[h->d.n_blocks], D, B, D, P + outl.w, P + outl.b, out, 1);
  else EG_TRY(linear(st, in[h->d.n_blocks], D, B, P + outl.w, D, P + outl.b, D, outl.out, out, outl.out));
  return EG_OK;
}

extern "C" int eg_policy_forward(EgPolicy* h, const float* state, const float* ego, const float* dist,
                                 const float* time, int B, int want_actor, int want_critic, float* out_actor,
                                 float* value, float* hx_out, void* stream) {
  EG_REQUIRE(h && state && ego && dist && time && B >= 0, "bad arguments");
  if (B == 0) return EG_OK;
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(policy_ws(h, B));
  cudaStream_t st = as_stream(stream);
  const EgPolicyDims& d = h->d;
  const PolicyLayout& L = h->L;
  const int H = d.h_dim, D = L.hx_dim;
  // x_enc: the 402-float frames and W_ih rows are re-pitched to 404 floats (16 B) so the products are TMA-eligible
  const int IP = h->in_pad;
  EG_TRY(launch_pad_rows(st, state, d.in_dim, B * 2, d.in_dim, h->xpad, IP));
  EG_TRY(launch_pad_rows(st, h->P + L.x_wih, d.in_dim, 3 * H, d.in_dim, h->wih_pad, IP));
  EG_TRY(gru2_forward(h, st, h->xpad, 2 * IP, IP, d.in_dim, B, h->wih_pad, IP, L.x_whh, L.x_bih, L.x_bhh, h->xr,
                      h->xz, h->xn, h->xg, h->xh1, h->hx, D));
  EG_TRY(gru2_forward(h, st, ego, 2 * d.ego_dim, d.ego_dim, d.ego_dim, B, h->P + L.e_wih, d.ego_dim, L.e_whh, L.e_bih, L.e_bhh, h->er,
                      h->ez, h->en, h->eg_, h->eh1, h->hx + H, D));
  EG_LAUNCH(pe_kernel, (B * 2 * d.pe_L + 127) / 128, 128, 0, st, dist, time, B, d.pe_L, D, 2 * H, h->hx);
  const bool fork = want_actor && want_critic && h->two_streams;
  cudaStream_t sc = fork ? h->side : st;                 // critic chain
  if (fork) {
    EG_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
    EG_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  }
  if (want_actor) {
    EG_TRY(mlp_block_forward(h, st, L.a_blk, L.a_out, h->a_in, h->a_t, h->a_u, h->out_a, B));
    if (out_actor) EG_CUDA_CHECK(cudaMemcpyAsync(out_actor, h->out_a, (size_t)B * 2 * d.z_dim * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (want_critic) {
    EG_TRY(mlp_block_forward(h, sc, L.c_blk, L.c_out, h->c_in, h->c_t, h->c_u, h->out_c, B));
    if (value) EG_CUDA_CHECK(cudaMemcpyAsync(value, h->out_c, (size_t)B * 4, cudaMemcpyDeviceToDevice, sc));
  }
  if (fork) {
    EG_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
    EG_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
  }
  if (hx_out) EG_CUDA_CHECK(cudaMemcpyAsync(hx_out, h->hx, (size_t)B * D * 4, cudaMemcpyDeviceToDevice, st));
  return EG_OK;
}

extern "C" int eg_gauss_sample(const float* out_actor, const float* eps, int B, int Z, float min_logvar,
                               float max_logvar, float* act, float* logp, void* stream) {
  EG_REQUIRE(out_actor && act && B >= 0 && Z > 0, "bad arguments");
  if (B == 0) return EG_OK;
  EG_LAUNCH(gauss_sample_kernel, B, 128, 0, as_stream(stream), out_actor, out_actor + Z, 2 * Z, eps, B, Z, min_logvar,
            max_logvar, act, logp);
  return EG_OK;
}

// dY [B,out] -> dW += dY^T X, db += colsum(dY) (unless the producer of dY already added it), dX = dY W (+ optional accumulate)
static int linear_backward(EgPolicy* h, cudaStream_t st, const float* dY, int ld_dy, const float* X, int ldx, int B,
                           const Lin& l, float* dX, int ld_dx, int dx_beta, bool bias_done = false) {
  const float* P = h->P;
  float* G = h->G;
  GemmArgs gw{dY, ld_dy, 1, X, ldx, G + l.w, l.in, nullptr, nullptr, 0, l.out, l.in, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(gw, true, false, st));
  if (!bias_done) EG_LAUNCH_PDL(colsum_kernel, dim3((l.out + 31) / 32, 4), 256, 0, st, dY, ld_dy, B, l.out, G + l.b);
  if (dX) {
    GemmArgs gx{dY, ld_dy, 1, P + l.w, l.in, dX, ld_dx, nullptr, nullptr, 0, B, l.in, l.out, ACT_NONE, 0.f, dx_beta, 1.0f};
    EG_TRY(launch_gemm(gx, false, false, st));
  }
  return EG_OK;
}

// dh / da / dt: the chain's scratch set [B, D]; the gradient w.r.t. hx is left in dh
static int mlp_block_backward(EgPolicy* h, cudaStream_t st, const Lin blk[][2], const Lin& outl, float** in, float** t,
                              float** u, const float* d_out, int B, float* dh, float* da, float* dt) {
  const int D = h->L.hx_dim;
  float* G = h->G;
  // out = in[nb] W_o^T + b_o
  EG_TRY(linear_backward(h, st, d_out, outl.out, in[h->d.n_blocks], D, B, outl, dh, D, 0));
  for (int k = h->d.n_blocks - 1; k >= 0; --k) {
    // in[k+1] = u + in[k];  u = lrelu(t W2^T + b2);  t = lrelu(in[k] W1^T + b1)   (bias gradients fused into the lrelu backward)
    EG_LAUNCH_PDL(lrelu_bwd_colsum_cl_kernel, dim3((D + 31) / 32, 4), 256, 0, st, dh, u[k], 0.01f, B, D, da, G + blk[k][1].b);
    EG_TRY(linear_backward(h, st, da, D, t[k], D, B, blk[k][1], dt, D, 0, true));
    EG_LAUNCH_PDL(lrelu_bwd_colsum_cl_kernel, dim3((D + 31) / 32, 4), 256, 0, st, dt, t[k], 0.01f, B, D, da, G + blk[k][0].b);
    // d in[k] = da1 W1 + dh (residual): accumulate straight into dh
    EG_TRY(linear_backward(h, st, da, D, in[k], D, B, blk[k][0], dh, D, 1, true));
  }
  return EG_OK;
}

static int gru2_backward(EgPolicy* h, cudaStream_t st, const float* x, int ld_env, int ld_frame, int in_dim, int B,
                         int64_t wih, int64_t whh, int64_t bih, int64_t bhh, float** r, float** z, float** n, float** g,
                         const float* h1, const float* dh2, int ld_dh2) {
  const int H = h->d.h_dim, H3 = 3 * H;
  const float* P = h->P;
  float* G = h->G;
  // step 2
  EG_LAUNCH_PDL(gru_bwd_kernel, ew_grid((int64_t)B * H), 256, 0, st, dh2, ld_dh2, r[1], z[1], n[1], g[1], h1, B, H, h->dgi,
            h->dgh, h->dh1);
  GemmArgs w1{h->dgi, H3, 1, x + ld_frame, ld_env, G + wih, in_dim, nullptr, nullptr, 0, H3, in_dim, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w1, true, false, st));
  EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgi, H3, B, H3, G + bih);
  GemmArgs w2{h->dgh, H3, 1, h1, H, G + whh, H, nullptr, nullptr, 0, H3, H, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w2, true, false, st));
  EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgh, H3, B, H3, G + bhh);
  // dh1 = dh2 * z + dgh W_hh
  GemmArgs x2{h->dgh, H3, 1, P + whh, H, h->dh1, H, nullptr, nullptr, 0, B, H, H3, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(x2, false, false, st));
  // step 1 (h0 = 0: no W_hh gradient, b_hh still receives dgh)
  EG_LAUNCH_PDL(gru_bwd_kernel, ew_grid((int64_t)B * H), 256, 0, st, h->dh1, H, r[0], z[0], n[0], g[0], nullptr, B, H,
            h->dgi, h->dgh, nullptr);
  GemmArgs w3{h->dgi, H3, 1, x, ld_env, G + wih, in_dim, nullptr, nullptr, 0, H3, in_dim, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w3, true, false, st));
  EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgi, H3, B, H3, G + bih);
  EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgh, H3, B, H3, G + bhh);
  return EG_OK;
}

// forward + loss head + backward of the actor and critic chains (everything whose gradients lie in the actor + critic
// prefix of the flat buffer, i.e. the clip-grad-norm range), leaving d loss / d hx in h->dhx
extern "C" int eg_ppo_loss_backward_mlp(EgPolicy* h, const float* state, const float* ego, const float* dist,
                                        const float* time, const float* act, const float* logp_old, const float* adv_norm,
                                        const float* returns, int B, float inv_B, float eps_clip, float vf_coef,
                                        float ent_coef, float min_logvar, float max_logvar, int zero_grads,
                                        float* stats, void* stream) {
  EG_REQUIRE(h && h->G && act && logp_old && adv_norm && returns && stats, "null pointer (was the policy created with a gradient buffer?)");
  if (B <= 0) return EG_OK;
  cudaStream_t st = as_stream(stream);
  const EgPolicyDims& d = h->d;
  const PolicyLayout& L = h->L;
  const int D = L.hx_dim;
  EG_TRY(eg_policy_forward(h, state, ego, dist, time, B, 1, 1, nullptr, nullptr, nullptr, stream));
  if (zero_grads) EG_CUDA_CHECK(cudaMemsetAsync(h->G, 0, (size_t)L.n_total * 4, st));
  EG_LAUNCH(ppo_head_kernel, B, 128, 0, st, h->out_a, h->out_c, act, logp_old, adv_norm, returns, B, d.z_dim, inv_B,
            eps_clip, vf_coef, ent_coef, min_logvar, max_logvar, h->d_out_a, h->d_out_c, stats);
  const int64_t nD = (int64_t)B * D;
  const bool fork = h->two_streams != 0;
  cudaStream_t sc = fork ? h->side : st;
  if (fork) {
    EG_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
    EG_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  }
  EG_TRY(mlp_block_backward(h, st, L.a_blk, L.a_out, h->a_in, h->a_t, h->a_u, h->d_out_a, B, h->dh, h->da, h->dt));
  EG_TRY(mlp_block_backward(h, sc, L.c_blk, L.c_out, h->c_in, h->c_t, h->c_u, h->d_out_c, B, h->dh2, h->da2, h->dt2));
  if (fork) {
    EG_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
    EG_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
  }
  EG_LAUNCH_PDL(add_kernel, ew_grid(nD), 256, 0, st, h->dh, h->dh2, nD, h->dhx);      // d hx = actor path + critic path
  return EG_OK;
}

// backward of the two GRU encoders (the shared_net tail of the flat gradient) from h->dhx; call after eg_ppo_loss_backward_mlp
// with the same state / ego / B
extern "C" int eg_ppo_backward_encoders(EgPolicy* h, const float* ego, int B, void* stream) {
  EG_REQUIRE(h && h->G && ego, "null pointer");
  if (B <= 0) return EG_OK;
  EG_REQUIRE(B <= h->cap, "eg_ppo_loss_backward_mlp must run first with the same batch");
  cudaStream_t st = as_stream(stream);
  const EgPolicyDims& d = h->d;
  const PolicyLayout& L = h->L;
  const int H = d.h_dim, D = L.hx_dim, IP = h->in_pad;
  EG_TRY(gru2_backward(h, st, h->xpad, 2 * IP, IP, d.in_dim, B, L.x_wih, L.x_whh, L.x_bih, L.x_bhh, h->xr,
                       h->xz, h->xn, h->xg, h->xh1, h->dhx, D));
  EG_TRY(gru2_backward(h, st, ego, 2 * d.ego_dim, d.ego_dim, d.ego_dim, B, L.e_wih, L.e_whh, L.e_bih, L.e_bhh, h->er,
                       h->ez, h->en, h->eg_, h->eh1, h->dhx + H, D));
  return EG_OK;
}

extern "C" int eg_ppo_loss_backward(EgPolicy* h, const float* state, const float* ego, const float* dist,
                                    const float* time, const float* act, const float* logp_old, const float* adv_norm,
                                    const float* returns, int B, float inv_B, float eps_clip, float vf_coef,
                                    float ent_coef, float min_logvar, float max_logvar, int zero_grads,
                                    float* stats, void* stream) {
  EG_TRY(eg_ppo_loss_backward_mlp(h, state, ego, dist, time, act, logp_old, adv_norm, returns, B, inv_B, eps_clip, vf_coef,
                                  ent_coef, min_logvar, max_logvar, zero_grads, stats, stream));
  return eg_ppo_backward_encoders(h, ego, B, stream);
}

extern "C" int eg_moments(const float* x, int64_t n, double* out2, void* stream) {
  EG_REQUIRE(x && out2 && n >= 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  EG_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
  if (n == 0) return EG_OK;
  EG_LAUNCH(moments_kernel, ew_grid(n), 256, 0, st, x, n, out2);
  return EG_OK;
}

extern "C" int eg_adv_normalize(const float* adv, int n, const double* moments3, float eps, float* out, void* stream) {
  EG_REQUIRE(adv && moments3 && out && n >= 0, "bad arguments");
  if (n == 0) return EG_OK;
  EG_LAUNCH(adv_normalize_kernel, (n + 255) / 256, 256, 0, as_stream(stream), adv, n, moments3, eps, out);
  return EG_OK;
}

extern "C" int eg_clip_adamw_step(EgPolicy* h, float* exp_avg, float* exp_avg_sq, float max_grad_norm, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  EG_REQUIRE(h && h->G && exp_avg && exp_avg_sq && step >= 1, "bad arguments");
  cudaStream_t st = as_stream(stream);
  const PolicyLayout& L = h->L;
  if (max_grad_norm > 0.0f) EG_TRY(eg_moments(h->G, L.n_actor_critic, h->mom, stream));
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  EG_LAUNCH(adamw_kernel, kNumSMs * 8, 256, 0, st, h->P, h->G, exp_avg, exp_avg_sq, L.n_total, L.n_actor_critic, h->mom,
            max_grad_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2s);
  return EG_OK;
}

extern "C" int eg_gae(const float* v_s, const float* v_next, const float* rew, const uint8_t* terminated,
                      const uint8_t* end_flag, int T, int E, double gamma, double gae_lambda, float* adv, float* ret,
                      void* stream) {
  EG_REQUIRE(v_s && v_next && rew && terminated && end_flag && adv && ret && T >= 0 && E >= 0, "bad arguments");
  if (T == 0 || E == 0) return EG_OK;
  EG_LAUNCH(gae_kernel, (E + 127) / 128, 128, 0, as_stream(stream), v_s, v_next, rew, terminated, end_flag, T, E, gamma,
            gae_lambda, adv, ret);
  return EG_OK;
}

// ==========================================================================================================
// C-VAE marker-predictor TRAINING (BASELINE config 3) - replaces GAMMAPrimitiveVAE.forward (encode + reparameterise
// + decode, reference motion/models/models_GAMMA_primitive.py:75-110) and GAMMAPrimitiveVAETrainOP._calc_loss_rec /
// calc_loss / one primitive of calc_loss_rollout (:400-432, :476-490) with a hand-written backward (BPTT through the
// 18-step GRUCell decoder and the 18-step encoder GRU). Parameters / gradients are flat buffers in
// GAMMAPrimitiveVAE.parameters() order: x_enc, e_rnn (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0), e_mlp.layers.{0,1},
// e_mu, e_logvar, drnn_mlp.layers.{0,1,2}, d_rnn (weight_ih, weight_hh, bias_ih, bias_hh), d_mlp.layers.{0,1}, d_out.
// ==========================================================================================================
namespace eg {

struct CvaeLayout {
  int64_t x_wih, x_whh, x_bih, x_bhh, e_wih, e_whh, e_bih, e_bhh;
  Lin e_mlp0, e_mlp1, e_mu, e_lv, dr0, dr1, dr2;
  int64_t d_wih, d_whh, d_bih, d_bhh;
  Lin d_mlp0, d_mlp1, d_out;
  int64_t n_total;
};

static CvaeLayout make_cvae_layout(const EgCvaeDims& d) {
  CvaeLayout L{};
  const int D = d.in_dim, H = d.h_dim, Z = d.z_dim, Hm = d.mlp_dim, H3 = 3 * H;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += n; return o; };
  auto lin = [&](int in, int out) { Lin l{off, off + (int64_t)in * out, in, out}; off += (int64_t)in * out + out; return l; };
  L.x_wih = take((int64_t)H3 * D); L.x_whh = take((int64_t)H3 * H); L.x_bih = take(H3); L.x_bhh = take(H3);
  L.e_wih = take((int64_t)H3 * D); L.e_whh = take((int64_t)H3 * H); L.e_bih = take(H3); L.e_bhh = take(H3);
  L.e_mlp0 = lin(2 * H, Hm); L.e_mlp1 = lin(Hm, H); L.e_mu = lin(H, Z); L.e_lv = lin(H, Z);
  L.dr0 = lin(H, Hm); L.dr1 = lin(Hm, H); L.dr2 = lin(H, H);
  L.d_wih = take((int64_t)H3 * (D + Z + H)); L.d_whh = take((int64_t)H3 * H); L.d_bih = take(H3); L.d_bhh = take(H3);
  L.d_mlp0 = lin(H, Hm); L.d_mlp1 = lin(Hm, H); L.d_out = lin(H, D);
  L.n_total = off;
  return L;
}

__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n, float* __restrict__ dx) {
  eg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = dy[i] * (1.0f - y[i] * y[i]);
}

// z = mu + eps * exp(0.5 logvar); also accumulates sum(-1 - lv + mu^2 + e^lv) into kl_sum (double)
__global__ void __launch_bounds__(256)
reparam_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps, int64_t n,
               float* __restrict__ z, double* __restrict__ kl_sum) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i];
    z[i] = m + eps[i] * expf(0.5f * l);
    s += (double)(-1.0f - l + m * m + expf(l));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(kl_sum, s);
}

// dmu = dz + c mu / n ; dlv = dz * eps * 0.5 e^{lv/2} + c * 0.5 (e^lv - 1) / n, c = w_kld * scale * d robust / d kld
__global__ void __launch_bounds__(256)
reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mu, const float* __restrict__ lv,
                   const float* __restrict__ eps, int64_t n, const double* __restrict__ kl_sum, float w_kld, int robust,
                   float scale, float* __restrict__ dmu, float* __restrict__ dlv, float* __restrict__ stats) {
  const double k = 0.5 * kl_sum[0] / (double)n;
  const float c = w_kld * scale * (robust ? (float)(k / sqrt(1.0 + k * k)) : 1.0f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i];
    dmu[i] = dz[i] + c * m / (float)n;
    dlv[i] = dz[i] * eps[i] * 0.5f * expf(0.5f * l) + c * 0.5f * (expf(l) - 1.0f) / (float)n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {
    const float kk = robust ? (float)(sqrt(1.0 + k * k) - 1.0) : (float)k;
    atomicAdd(stats + 2, kk * scale);
    atomicAdd(stats + 0, w_kld * kk * scale);
  }
}

__device__ __forceinline__ float sgnf(float x) { return (x > 0.0f) - (x < 0.0f); }

// dL/dY_rec for w_rec * mean|Y - Yrec| + w_td * mean|(Yrec[t+1]-Yrec[t]) - (Y[t+1]-Y[t])| ; Y, Yrec [T,B,D] t-major
__global__ void __launch_bounds__(256)
rec_loss_grad_kernel(const float* __restrict__ Y, const float* __restrict__ Yr, int T, int64_t BD, float w_rec,
                     float w_td, float scale, float* __restrict__ dYr, double* __restrict__ sums) {
  double s_rec = 0.0, s_td = 0.0;
  const int64_t n = (int64_t)T * BD;
  const float c_rec = w_rec * scale / (float)n, c_td = w_td * scale / (float)((int64_t)(T - 1) * BD);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / BD);
    const float e = Yr[i] - Y[i];
    float g = c_rec * sgnf(e);
    s_rec += fabsf(e);
    if (t + 1 < T) {
      const float d = (Yr[i + BD] - Yr[i]) - (Y[i + BD] - Y[i]);
      g -= c_td * sgnf(d);
      s_td += fabsf(d);
    }
    if (t > 0) {
      const float d = (Yr[i] - Yr[i - BD]) - (Y[i] - Y[i - BD]);
      g += c_td * sgnf(d);
    }
    dYr[i] = g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s_rec += __shfl_xor_sync(0xffffffffu, s_rec, o); s_td += __shfl_xor_sync(0xffffffffu, s_td, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(sums, s_rec); atomicAdd(sums + 1, s_td); }
}

__global__ void cvae_stats_kernel(const double* __restrict__ sums, int T, int64_t BD, float w_rec, float w_td, float scale,
                                  int in_loss, float* __restrict__ stats) {
  const float rec = w_rec * (float)(sums[0] / (double)((int64_t)T * BD)) + w_td * (float)(sums[1] / (double)((int64_t)(T - 1) * BD));
  atomicAdd(stats + 1, rec * scale);
  if (in_loss) atomicAdd(stats + 0, rec * scale);
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 float lr, float beta1, float beta2, float eps, float wd_decoupled, float bc1, float bc2_sqrt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * (1.0f - lr * wd_decoupled);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    p[i] = pi;
  }
}

// new frame from joints (CanonicalCoordinateExtractor, baseops.py:214-225): jts [B,J,3] -> R [B,9], T [B,3]
__global__ void new_coordinate_kernel(const float* __restrict__ jts, int ld, int B, float* __restrict__ R, float* __restrict__ T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* j = jts + (int64_t)b * ld;
  const float x0r = j[6] - j[3], x1r = j[7] - j[4];
  const float nx = sqrtf(x0r * x0r + x1r * x1r + 0.0f);
  const float x0 = x0r / nx, x1 = x1r / nx;
  float y0 = -x1, y1 = x0;
  const float ny = sqrtf(y0 * y0 + y1 * y1 + 0.0f);
  y0 /= ny; y1 /= ny;
  float* r = R + (int64_t)b * 9;
  r[0] = x0; r[1] = y0; r[2] = 0.f; r[3] = x1; r[4] = y1; r[5] = 0.f; r[6] = 0.f; r[7] = 0.f; r[8] = 1.f;
  T[b * 3] = j[0]; T[b * 3 + 1] = j[1]; T[b * 3 + 2] = j[2];
}

// pts [t,B,P,3]: inverse==0: out = R p + T ; inverse==1: out = R^T (p - T)   (einsum patterns of :462-466)
__global__ void __launch_bounds__(256)
rigid_points_kernel(const float* __restrict__ R, const float* __restrict__ T, const float* __restrict__ pts, int nt, int B,
                    int P, int inverse, float* __restrict__ out) {
  const int64_t n = (int64_t)nt * B * P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)((i / P) % B);
    const float* r = R + (int64_t)b * 9;
    const float* t = T + (int64_t)b * 3;
    const float* p = pts + i * 3;
    float* o = out + i * 3;
    if (inverse) {
      const float d0 = p[0] - t[0], d1 = p[1] - t[1], d2 = p[2] - t[2];
      for (int q = 0; q < 3; ++q) o[q] = r[0 * 3 + q] * d0 + r[1 * 3 + q] * d1 + r[2 * 3 + q] * d2;
    } else {
      for (int q = 0; q < 3; ++q) o[q] = (r[q * 3 + 0] * p[0] + r[q * 3 + 1] * p[1] + r[q * 3 + 2] * p[2]) + t[q];
    }
  }
}

}  // namespace eg

struct EgCvae {
  int device = 0;
  EgCvaeDims d;
  CvaeLayout L;
  float *P = nullptr, *G = nullptr;
  int cap = 0;
  std::vector<float*> owned;
  // saved activations
  float *gi = nullptr, *gh = nullptr;
  float *xr[2], *xz[2], *xn[2], *xg[2], *xh[2];            // x_enc steps (xh[t] = h after step t)
  float *er[18], *ez[18], *en[18], *eg_[18], *eh[18];       // e_rnn steps
  float *hcat = nullptr, *ea1 = nullptr, *ea2 = nullptr, *mu = nullptr, *lv = nullptr, *z = nullptr, *hz = nullptr;
  float *dr_a0 = nullptr, *dr_a1 = nullptr, *h0 = nullptr, *c = nullptr;
  float *dr_[18], *dz_[18], *dn_[18], *dg_[18], *dh_[18], *f1[18], *f2[18];
  // backward scratch
  float *dY = nullptr, *dy = nullptr, *dh = nullptr, *dhp = nullptr, *da = nullptr, *db = nullptr, *dgi = nullptr,
        *dgh = nullptr, *dc = nullptr, *dhz = nullptr, *dmu = nullptr, *dlv = nullptr, *dhcat = nullptr, *dhx = nullptr;
  double* sums = nullptr;       // [0] |rec| sum, [1] |td| sum, [2] kl sum
};

#undef EG_TRY
#define EG_TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static int cvae_ws(EgCvae* h, int B) {
  if (B <= h->cap) return EG_OK;
  for (float* p : h->owned) cudaFree(p);
  h->owned.clear(); h->cap = 0;
  const EgCvaeDims& d = h->d;
  const size_t b = (size_t)B, H = d.h_dim, H3 = 3 * H, D = d.in_dim, Z = d.z_dim, Hm = d.mlp_dim;
  auto A = [&](float** p, size_t n) -> int {
    EG_CUDA_CHECK(cudaMalloc((void**)p, n * sizeof(float)));
    h->owned.push_back(*p);
    return EG_OK;
  };
  EG_TRY(A(&h->gi, b * H3)); EG_TRY(A(&h->gh, b * H3));
  for (int t = 0; t < 2; ++t) { EG_TRY(A(&h->xr[t], b * H)); EG_TRY(A(&h->xz[t], b * H)); EG_TRY(A(&h->xn[t], b * H)); EG_TRY(A(&h->xg[t], b * H)); EG_TRY(A(&h->xh[t], b * H)); }
  for (int t = 0; t < 18; ++t) {
    EG_TRY(A(&h->er[t], b * H)); EG_TRY(A(&h->ez[t], b * H)); EG_TRY(A(&h->en[t], b * H)); EG_TRY(A(&h->eg_[t], b * H)); EG_TRY(A(&h->eh[t], b * H));
    EG_TRY(A(&h->dr_[t], b * H)); EG_TRY(A(&h->dz_[t], b * H)); EG_TRY(A(&h->dn_[t], b * H)); EG_TRY(A(&h->dg_[t], b * H)); EG_TRY(A(&h->dh_[t], b * H));
    EG_TRY(A(&h->f1[t], b * Hm)); EG_TRY(A(&h->f2[t], b * H));
  }
  EG_TRY(A(&h->hcat, b * 2 * H)); EG_TRY(A(&h->ea1, b * Hm)); EG_TRY(A(&h->ea2, b * H)); EG_TRY(A(&h->mu, b * Z)); EG_TRY(A(&h->lv, b * Z));
  EG_TRY(A(&h->z, b * Z)); EG_TRY(A(&h->hz, b * (H + Z))); EG_TRY(A(&h->dr_a0, b * Hm)); EG_TRY(A(&h->dr_a1, b * H)); EG_TRY(A(&h->h0, b * H));
  EG_TRY(A(&h->c, b * H3));
  EG_TRY(A(&h->dY, 18 * b * D)); EG_TRY(A(&h->dy, b * D)); EG_TRY(A(&h->dh, b * H)); EG_TRY(A(&h->dhp, b * H)); EG_TRY(A(&h->da, b * Hm));
  EG_TRY(A(&h->db, b * Hm)); EG_TRY(A(&h->dgi, b * H3)); EG_TRY(A(&h->dgh, b * H3)); EG_TRY(A(&h->dc, b * H3)); EG_TRY(A(&h->dhz, b * (H + Z)));
  EG_TRY(A(&h->dmu, b * Z)); EG_TRY(A(&h->dlv, b * Z)); EG_TRY(A(&h->dhcat, b * 2 * H)); EG_TRY(A(&h->dhx, b * H));
  h->cap = B;
  return EG_OK;
}

extern "C" int64_t eg_cvae_param_count(const EgCvaeDims* d) { return d ? make_cvae_layout(*d).n_total : -1; }

extern "C" int eg_cvae_create(const EgCvaeDims* dims, float* params_flat, float* grads_flat, int device, EgCvae** out) {
  EG_REQUIRE(dims && params_flat && grads_flat && out, "null pointer");
  EG_CUDA_CHECK(cudaSetDevice(device));
  EgCvae* h = new EgCvae();
  h->device = device; h->d = *dims; h->L = make_cvae_layout(*dims); h->P = params_flat; h->G = grads_flat;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->sums, 4 * sizeof(double)));
  *out = h;
  return EG_OK;
}

extern "C" void eg_cvae_destroy(EgCvae* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (float* p : h->owned) cudaFree(p);
  cudaFree(h->sums);
  delete h;
}

namespace {
// dY [B,out] -> dW += dY^T X, db += colsum(dY), optional dX (= or +=) dY W, for a Linear at flat offsets (w, b)
int lin_bwd(EgCvae* h, cudaStream_t st, const float* dY, int ld_dy, const float* X, int ldx, int B, int64_t w, int64_t b,
            int in, int out, int ldw, float* dX, int ld_dx, int dx_beta) {
  GemmArgs gw{dY, ld_dy, 1, X, ldx, h->G + w, ldw, nullptr, nullptr, 0, out, in, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(gw, true, false, st));
  if (b >= 0) EG_LAUNCH_PDL(colsum_kernel, dim3((out + 31) / 32, 4), 256, 0, st, dY, ld_dy, B, out, h->G + b);
  if (dX) {
    GemmArgs gx{dY, ld_dy, 1, h->P + w, ldw, dX, ld_dx, nullptr, nullptr, 0, B, in, out, ACT_NONE, 0.f, dx_beta, 1.0f};
    EG_TRY(launch_gemm(gx, false, false, st));
  }
  return EG_OK;
}
}  // namespace

// One primitive: forward (encode, reparameterise, decode), losses, full backward. X [2,B,D] and Y [18,B,D] are
// time-major like the reference tensors. Gradients are ACCUMULATED into the flat buffer scaled by loss_scale
// (1 / #primitives of a rollout, calc_loss_rollout :500). Y_rec [18,B,D] is returned for the next rollout seed.
// stats (device float[4], accumulated): 0 total loss, 1 rec loss (weighted), 2 kld term.
#define EG_CVAE_LOCALS                                                                                              \
  const EgCvaeDims& d = h->d;                                                                                      \
  const CvaeLayout& L = h->L;                                                                                      \
  const int D = d.in_dim, H = d.h_dim, Z = d.z_dim, Hm = d.mlp_dim, H3 = 3 * H, T = 18, Kin = H + Z + D;           \
  const float* P = h->P;                                                                                           \
  float* G = h->G;                                                                                                 \
  const int64_t BD = (int64_t)B * D;                                                                               \
  const int64_t nBH = (int64_t)B * H;                                                                              \
  (void)G; (void)nBH; (void)Kin; (void)Hm; (void)Z; (void)P

// forward of one primitive with every activation the backward needs kept in the handle's workspace
static int cvae_forward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float* Y_rec, cudaStream_t st) {
  EG_CVAE_LOCALS;
  // ---------------- forward: encoder ----------------
  for (int t = 0; t < 2; ++t) {   // x_enc
    EG_TRY(linear(st, X + t * BD, D, B, P + L.x_wih, D, P + L.x_bih, D, H3, h->gi, H3));
    if (t > 0) EG_TRY(linear(st, h->xh[t - 1], H, B, P + L.x_whh, H, P + L.x_bhh, H, H3, h->gh, H3));
    EG_TRY(launch_gru_gate(st, h->gi, t ? h->gh : nullptr, P + L.x_bhh, t ? h->xh[t - 1] : nullptr, h->xh[t], B, H, H,
                           h->xr[t], h->xz[t], h->xn[t], h->xg[t]));
  }
  for (int t = 0; t < T; ++t) {   // e_rnn over the 18 target frames
    EG_TRY(linear(st, Y + t * BD, D, B, P + L.e_wih, D, P + L.e_bih, D, H3, h->gi, H3));
    if (t > 0) EG_TRY(linear(st, h->eh[t - 1], H, B, P + L.e_whh, H, P + L.e_bhh, H, H3, h->gh, H3));
    EG_TRY(launch_gru_gate(st, h->gi, t ? h->gh : nullptr, P + L.e_bhh, t ? h->eh[t - 1] : nullptr, h->eh[t], B, H, H,
                           h->er[t], h->ez[t], h->en[t], h->eg_[t]));
  }
  const float* hx = h->xh[1];
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->hcat, 2 * H * 4, hx, H * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->hcat + H, 2 * H * 4, h->eh[T - 1], H * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_TRY(linear(st, h->hcat, 2 * H, B, P + L.e_mlp0.w, 2 * H, P + L.e_mlp0.b, 2 * H, Hm, h->ea1, Hm, ACT_TANH));
  EG_TRY(linear(st, h->ea1, Hm, B, P + L.e_mlp1.w, Hm, P + L.e_mlp1.b, Hm, H, h->ea2, H, ACT_TANH));
  EG_TRY(linear(st, h->ea2, H, B, P + L.e_mu.w, H, P + L.e_mu.b, H, Z, h->mu, Z));
  EG_TRY(linear(st, h->ea2, H, B, P + L.e_lv.w, H, P + L.e_lv.b, H, Z, h->lv, Z));
  EG_CUDA_CHECK(cudaMemsetAsync(h->sums, 0, 4 * sizeof(double), st));
  EG_LAUNCH(reparam_kernel, ew_grid((int64_t)B * Z), 256, 0, st, h->mu, h->lv, eps, (int64_t)B * Z, h->z, h->sums + 2);
  // ---------------- forward: decoder ----------------
  EG_TRY(linear(st, hx, H, B, P + L.dr0.w, H, P + L.dr0.b, H, Hm, h->dr_a0, Hm, ACT_TANH));
  EG_TRY(linear(st, h->dr_a0, Hm, B, P + L.dr1.w, Hm, P + L.dr1.b, Hm, H, h->dr_a1, H, ACT_TANH));
  EG_TRY(linear(st, h->dr_a1, H, B, P + L.dr2.w, H, P + L.dr2.b, H, H, h->h0, H, ACT_TANH));
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->hz, (H + Z) * 4, hx, H * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->hz + H, (H + Z) * 4, h->z, Z * 4, Z * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_TRY(linear(st, h->hz, H + Z, B, P + L.d_wih, Kin, P + L.d_bih, H + Z, H3, h->c, H3));
  for (int i = 0; i < T; ++i) {
    const float* yp = i ? Y_rec + (int64_t)(i - 1) * BD : X + BD;
    const float* hp = i ? h->dh_[i - 1] : h->h0;
    EG_TRY(linear(st, yp, D, B, P + L.d_wih + H + Z, Kin, nullptr, D, H3, h->gi, H3, ACT_NONE, 0.f, h->c, H3));
    EG_TRY(linear(st, hp, H, B, P + L.d_whh, H, P + L.d_bhh, H, H3, h->gh, H3));
    EG_TRY(launch_gru_gate(st, h->gi, h->gh, nullptr, hp, h->dh_[i], B, H, H, h->dr_[i], h->dz_[i], h->dn_[i], h->dg_[i]));
    EG_TRY(linear(st, h->dh_[i], H, B, P + L.d_mlp0.w, H, P + L.d_mlp0.b, H, Hm, h->f1[i], Hm, ACT_TANH));
    EG_TRY(linear(st, h->f1[i], Hm, B, P + L.d_mlp1.w, Hm, P + L.d_mlp1.b, Hm, H, h->f2[i], H, ACT_TANH));
    EG_TRY(linear(st, h->f2[i], H, B, P + L.d_out.w, H, P + L.d_out.b, H, D, Y_rec + (int64_t)i * BD, D, ACT_NONE, 0.f, yp, D));
  }
  return EG_OK;
}

// losses + backward of the primitive cvae_forward just ran. rec_in_loss = 0 keeps the reconstruction term out of the
// objective (it is still reported in stats[1]); dY_extra [18,B,D] (nullable) is an additional gradient w.r.t. Y_rec from
// a downstream loss (the regressor cycle loss of the combo training op).
static int cvae_backward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float w_rec, float w_td,
                         float w_kld, int robust_kld, float loss_scale, int rec_in_loss, const float* Y_rec,
                         const float* dY_extra, float* stats, cudaStream_t st) {
  EG_CVAE_LOCALS;
  const float* hx = h->xh[1];
  // ---------------- losses ----------------
  EG_LAUNCH(rec_loss_grad_kernel, ew_grid((int64_t)T * BD), 256, 0, st, Y, Y_rec, T, BD, w_rec, w_td,
            rec_in_loss ? loss_scale : 0.0f, h->dY, h->sums);
  EG_LAUNCH(cvae_stats_kernel, 1, 1, 0, st, h->sums, T, BD, w_rec, w_td, loss_scale, rec_in_loss, stats);
  if (dY_extra) EG_LAUNCH_PDL(add_kernel, ew_grid((int64_t)T * BD), 256, 0, st, (const float*)h->dY, dY_extra, (int64_t)T * BD, h->dY);
  // ---------------- backward: decoder BPTT ----------------
  EG_CUDA_CHECK(cudaMemsetAsync(h->dh, 0, nBH * 4, st));
  EG_CUDA_CHECK(cudaMemsetAsync(h->dc, 0, (size_t)B * H3 * 4, st));
  EG_CUDA_CHECK(cudaMemsetAsync(h->dy, 0, (size_t)BD * 4, st));
  for (int i = T - 1; i >= 0; --i) {
    const float* yp = i ? Y_rec + (int64_t)(i - 1) * BD : X + BD;
    const float* hp = i ? h->dh_[i - 1] : h->h0;
    // dy_i = dL/dY_rec[i] + carried gradient w.r.t. y_p of step i+1
    EG_LAUNCH_PDL(add_kernel, ew_grid(BD), 256, 0, st, h->dY + (int64_t)i * BD, h->dy, BD, h->dy);
    // y_i = d_out(f2) + y_p
    EG_TRY(lin_bwd(h, st, h->dy, D, h->f2[i], H, B, L.d_out.w, L.d_out.b, H, D, H, h->db, H, 0));
    EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid(nBH), 256, 0, st, h->db, h->f2[i], nBH, h->db);
    EG_TRY(lin_bwd(h, st, h->db, H, h->f1[i], Hm, B, L.d_mlp1.w, L.d_mlp1.b, Hm, H, Hm, h->da, Hm, 0));
    EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid((int64_t)B * Hm), 256, 0, st, h->da, h->f1[i], (int64_t)B * Hm, h->da);
    EG_TRY(lin_bwd(h, st, h->da, Hm, h->dh_[i], H, B, L.d_mlp0.w, L.d_mlp0.b, H, Hm, H, h->dh, H, 1));   // dh_i += ...
    // GRUCell backward
    EG_LAUNCH_PDL(gru_bwd_kernel, ew_grid(nBH), 256, 0, st, h->dh, H, h->dr_[i], h->dz_[i], h->dn_[i], h->dg_[i], hp, B, H,
              h->dgi, h->dgh, h->dhp);
    EG_TRY(lin_bwd(h, st, h->dgh, H3, hp, H, B, L.d_whh, L.d_bhh, H, H3, H, h->dhp, H, 1));              // dh_{i-1}
    // gi = c + y_p Wy^T : weight slice of d_rnn.weight_ih (columns H+Z..), bias lives in c
    EG_TRY(lin_bwd(h, st, h->dgi, H3, yp, D, B, L.d_wih + H + Z, -1, D, H3, Kin, h->db, D, 0));           // d y_p (GRU path)
    EG_LAUNCH_PDL(add_kernel, ew_grid((int64_t)B * H3), 256, 0, st, h->dc, h->dgi, (int64_t)B * H3, h->dc);
    // carry: dy_{i-1} = dy_i (residual) + dgi Wy
    EG_LAUNCH_PDL(add_kernel, ew_grid(BD), 256, 0, st, h->dy, h->db, BD, h->dy);
    std::swap(h->dh, h->dhp);
  }
  // h->dh now holds dL/dh0 ; c = [hx,z] W_ih[:, :H+Z]^T + b_ih
  EG_TRY(lin_bwd(h, st, h->dc, H3, h->hz, H + Z, B, L.d_wih, L.d_bih, H + Z, H3, Kin, h->dhz, H + Z, 0));
  // drnn_mlp backward: h0 = tanh(W2 tanh(W1 tanh(W0 hx)))
  EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid(nBH), 256, 0, st, h->dh, h->h0, nBH, h->dh);
  EG_TRY(lin_bwd(h, st, h->dh, H, h->dr_a1, H, B, L.dr2.w, L.dr2.b, H, H, H, h->db, H, 0));
  EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid(nBH), 256, 0, st, h->db, h->dr_a1, nBH, h->db);
  EG_TRY(lin_bwd(h, st, h->db, H, h->dr_a0, Hm, B, L.dr1.w, L.dr1.b, Hm, H, Hm, h->da, Hm, 0));
  EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid((int64_t)B * Hm), 256, 0, st, h->da, h->dr_a0, (int64_t)B * Hm, h->da);
  EG_TRY(lin_bwd(h, st, h->da, Hm, hx, H, B, L.dr0.w, L.dr0.b, H, Hm, H, h->dhx, H, 0));                 // dhx (1)
  // ---------------- backward: latent + encoder ----------------
  // dz = dhz[:, H:], dhx (2) = dhz[:, :H]; gather dz contiguous through a strided copy
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->z, Z * 4, h->dhz + H, (H + Z) * 4, Z * 4, B, cudaMemcpyDeviceToDevice, st));   // reuse z as dz
  EG_LAUNCH(reparam_bwd_kernel, ew_grid((int64_t)B * Z), 256, 0, st, h->z, h->mu, h->lv, eps, (int64_t)B * Z, h->sums + 2,
            w_kld, robust_kld, loss_scale, h->dmu, h->dlv, stats);
  EG_TRY(lin_bwd(h, st, h->dmu, Z, h->ea2, H, B, L.e_mu.w, L.e_mu.b, H, Z, H, h->db, H, 0));
  EG_TRY(lin_bwd(h, st, h->dlv, Z, h->ea2, H, B, L.e_lv.w, L.e_lv.b, H, Z, H, h->db, H, 1));
  EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid(nBH), 256, 0, st, h->db, h->ea2, nBH, h->db);
  EG_TRY(lin_bwd(h, st, h->db, H, h->ea1, Hm, B, L.e_mlp1.w, L.e_mlp1.b, Hm, H, Hm, h->da, Hm, 0));
  EG_LAUNCH_PDL(tanh_bwd_kernel, ew_grid((int64_t)B * Hm), 256, 0, st, h->da, h->ea1, (int64_t)B * Hm, h->da);
  EG_TRY(lin_bwd(h, st, h->da, Hm, h->hcat, 2 * H, B, L.e_mlp0.w, L.e_mlp0.b, 2 * H, Hm, 2 * H, h->dhcat, 2 * H, 0));
  // dhx total = drnn path + c path + encoder path
  // dhx += dhz[:, :H] + dhcat[:, :H]  (strided adds via 2-D copies into scratch then add)
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->db, H * 4, h->dhz, (H + Z) * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_LAUNCH_PDL(add_kernel, ew_grid(nBH), 256, 0, st, h->dhx, h->db, nBH, h->dhx);
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->db, H * 4, h->dhcat, 2 * H * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  EG_LAUNCH_PDL(add_kernel, ew_grid(nBH), 256, 0, st, h->dhx, h->db, nBH, h->dhx);
  // x_enc BPTT (2 steps)
  {
    float* dcur = h->dhx;
    for (int t = 1; t >= 0; --t) {
      const float* hp = t ? h->xh[t - 1] : nullptr;
      EG_LAUNCH_PDL(gru_bwd_kernel, ew_grid(nBH), 256, 0, st, dcur, H, h->xr[t], h->xz[t], h->xn[t], h->xg[t], hp, B, H, h->dgi,
                h->dgh, h->dhp);
      EG_TRY(lin_bwd(h, st, h->dgi, H3, X + t * BD, D, B, L.x_wih, L.x_bih, D, H3, D, nullptr, 0, 0));
      if (t > 0) EG_TRY(lin_bwd(h, st, h->dgh, H3, hp, H, B, L.x_whh, L.x_bhh, H, H3, H, h->dhp, H, 1));
      else EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgh, H3, B, H3, G + L.x_bhh);
      dcur = h->dhp;
    }
  }
  // e_rnn BPTT (18 steps); dh_T = dhcat[:, H:]
  EG_CUDA_CHECK(cudaMemcpy2DAsync(h->dh, H * 4, h->dhcat + H, 2 * H * 4, H * 4, B, cudaMemcpyDeviceToDevice, st));
  for (int t = T - 1; t >= 0; --t) {
    const float* hp = t ? h->eh[t - 1] : nullptr;
    EG_LAUNCH_PDL(gru_bwd_kernel, ew_grid(nBH), 256, 0, st, h->dh, H, h->er[t], h->ez[t], h->en[t], h->eg_[t], hp, B, H, h->dgi,
              h->dgh, h->dhp);
    EG_TRY(lin_bwd(h, st, h->dgi, H3, Y + t * BD, D, B, L.e_wih, L.e_bih, D, H3, D, nullptr, 0, 0));
    if (t > 0) EG_TRY(lin_bwd(h, st, h->dgh, H3, hp, H, B, L.e_whh, L.e_bhh, H, H3, H, h->dhp, H, 1));
    else EG_LAUNCH_PDL(colsum_kernel, dim3((H3 + 31) / 32, 4), 256, 0, st, h->dgh, H3, B, H3, G + L.e_bhh);
    std::swap(h->dh, h->dhp);
  }
  return EG_OK;
}

extern "C" int eg_cvae_loss_backward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float w_rec,
                                     float w_td, float w_kld, int robust_kld, float loss_scale, float* Y_rec, float* stats,
                                     void* stream) {
  EG_REQUIRE(h && X && Y && eps && Y_rec && stats && B > 0, "bad arguments");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(cvae_ws(h, B));
  cudaStream_t st = as_stream(stream);
  EG_TRY(cvae_forward(h, X, Y, eps, B, Y_rec, st));
  return cvae_backward(h, X, Y, eps, B, w_rec, w_td, w_kld, robust_kld, loss_scale, 1, Y_rec, nullptr, stats, st);
}

extern "C" int eg_cvae_forward_train(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float* Y_rec,
                                     void* stream) {
  EG_REQUIRE(h && X && Y && eps && Y_rec && B > 0, "bad arguments");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(cvae_ws(h, B));
  return cvae_forward(h, X, Y, eps, B, Y_rec, as_stream(stream));
}

extern "C" int eg_cvae_backward(EgCvae* h, const float* X, const float* Y, const float* eps, int B, float w_rec, float w_td,
                                float w_kld, int robust_kld, float loss_scale, int rec_in_loss, const float* Y_rec,
                                const float* dY_extra, float* stats, void* stream) {
  EG_REQUIRE(h && X && Y && eps && Y_rec && stats && B > 0, "bad arguments");
  EG_REQUIRE(B <= h->cap, "eg_cvae_forward_train must run first with the same batch");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  return cvae_backward(h, X, Y, eps, B, w_rec, w_td, w_kld, robust_kld, loss_scale, rec_in_loss, Y_rec, dY_extra, stats,
                       as_stream(stream));
}

extern "C" int eg_adam_step_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  EG_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad arguments");
  if (n == 0) return EG_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  EG_LAUNCH(adam_flat_kernel, kNumSMs * 8, 256, 0, as_stream(stream), params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
            eps, weight_decay, bc1, bc2s);
  return EG_OK;
}

extern "C" int eg_new_coordinate(const float* joints, int ld_body, int B, float* R, float* T, void* stream) {
  EG_REQUIRE(joints && R && T && B >= 0 && ld_body >= 9, "bad arguments");
  if (B == 0) return EG_OK;
  EG_LAUNCH(new_coordinate_kernel, (B + 127) / 128, 128, 0, as_stream(stream), joints, ld_body, B, R, T);
  return EG_OK;
}

extern "C" int eg_rigid_points(const float* R, const float* T, const float* pts, int nt, int B, int P, int inverse, float* out,
                               void* stream) {
  EG_REQUIRE(R && T && pts && out && nt >= 0 && B >= 0 && P >= 0, "bad arguments");
  if ((int64_t)nt * B * P == 0) return EG_OK;
  EG_LAUNCH(rigid_points_kernel, ew_grid((int64_t)nt * B * P), 256, 0, as_stream(stream), R, T, pts, nt, B, P, inverse, out);
  return EG_OK;
}

// ============================================================================================
// Body-regressor training: GAMMARegressorTrainOP.calc_loss + the step of its train loop
// (motion/models/models_GAMMA_primitive.py:594-633, :664-682) for MoshRegressor (:178-301, use_cont + relu):
//   xb_0 = 0;  xb_{r+1} = pnet([markers, xb_r, betas]) + xb_r  (n_recur times, ResNetBlock :160-175)
//   yb = [transl, aa(GramSchmidt(6-D)) x 22, hand PCA]           (_cont2aa :208-219)
//   loss = L1(markers, SMPL-X markers(yb, betas)) + w * mean(hand PCA^2)
// The SMPL-X gradient arrives as dL/dR of the 22 regressed joints (eg_lbs_markers_backward_rot) and enters the 6-D
// parameters through the Gram-Schmidt Jacobian: axis-angle is only an intermediate re-parameterisation of the same
// rotation (exp(log R) = R, and the Gram-Schmidt image is tangent to SO(3)), so the chain rule through it cancels.
// ============================================================================================
namespace eg {

struct RegLayout { int64_t in_w, in_b, blk, out_w, out_b, n_total; };
static RegLayout make_reg_layout(const EgRegressorDims& d) {
  RegLayout L;
  const int64_t H = d.h_dim, K = d.in_dim + d.body_dim + 10;
  int64_t off = 0;
  L.in_w = off; off += H * K;
  L.in_b = off; off += H;
  L.blk = off; off += (int64_t)d.n_blocks * 2 * (H * H + H);
  L.out_w = off; off += (int64_t)d.body_dim * H;
  L.out_b = off; off += d.body_dim;
  L.n_total = off;
  return L;
}

// xin[m] = [markers[m] (D) | xb[m] (BD) | betas[m] (10)]
__global__ void __launch_bounds__(256)
reg_concat_kernel(const float* __restrict__ mk, const float* __restrict__ xb, const float* __restrict__ betas, int M, int D,
                  int BD, float* __restrict__ xin) {
  const int K = D + BD + 10;
  const int64_t total = (int64_t)M * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / K), k = (int)(i % K);
    xin[i] = k < D ? mk[(int64_t)m * D + k] : k < D + BD ? xb[(int64_t)m * BD + (k - D)] : betas[(int64_t)m * 10 + (k - D - BD)];
  }
}

// MoshRegressor._cont2aa: same arithmetic as the inference tail (nn.cu regressor_tail_kernel)
__global__ void __launch_bounds__(128)
reg_cont2aa_kernel(const float* __restrict__ xb_cont, int M, float* __restrict__ yb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = i / 32, slot = i % 32;
  if (row >= M) return;
  const float* x = xb_cont + (int64_t)row * 159;
  float* y = yb + (int64_t)row * 93;
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

// F.l1_loss(ref, pred): sums[0] += sum |pred - ref|, d_pred = sign(pred - ref) / n
__global__ void __launch_bounds__(256)
reg_l1_kernel(const float* __restrict__ pred, const float* __restrict__ ref, int64_t n, float* __restrict__ d_pred,
              double* __restrict__ sum) {
  double s = 0.0;
  const float inv = 1.0f / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - ref[i];
    s += (double)fabsf(d);
    d_pred[i] = d > 0.0f ? inv : d < 0.0f ? -inv : 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}

// sum of the squared hand-PCA entries xb[:, 135:159]
__global__ void __launch_bounds__(256)
reg_hpose_kernel(const float* __restrict__ xb_cont, int M, double* __restrict__ sum) {
  double s = 0.0;
  const int64_t total = (int64_t)M * 24;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = xb_cont[(i / 24) * 159 + 135 + (i % 24)];
    s += (double)(v * v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}

__global__ void reg_stats_kernel(const double* __restrict__ sums, int64_t n_marker, int64_t n_hpose, float w_hpose,
                                 float* __restrict__ stats) {
  const float lm = (float)(sums[0] / (double)n_marker), lh = (float)(sums[1] / (double)n_hpose);
  stats[0] = lm + w_hpose * lh; stats[1] = lm; stats[2] = lh;
}

// d xb_cont from (d yb, dL/dR): transl and hand PCA pass through (+ the hand regulariser), the 22 rotations go through
// the Gram-Schmidt backward of cont6d_to_rotmat (geom.cuh): a' = a/|a|, u = c - (a'.c) a', b = u/|u|, d = a' x b,
// R columns (a', b, d); a = x[0,2,4], c = x[1,3,5].
__global__ void __launch_bounds__(128)
reg_gs_bwd_kernel(const float* __restrict__ xb_cont, const float* __restrict__ d_rot, const float* __restrict__ d_yb, int M,
                  float hpose_scale /* w * 2 / (M * 24) */, float* __restrict__ d_xb_cont) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = i / 32, slot = i % 32;
  if (row >= M) return;
  const float* x = xb_cont + (int64_t)row * 159;
  const float* gy = d_yb + (int64_t)row * 93;
  float* gx = d_xb_cont + (int64_t)row * 159;
  if (slot < 22) {
    const float* xs = x + 3 + slot * 6;
    const float* gR = d_rot + (int64_t)row * 198 + slot * 9;
    const float a[3] = {xs[0], xs[2], xs[4]}, c[3] = {xs[1], xs[3], xs[5]};
    const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
    const float an[3] = {a[0] / na, a[1] / na, a[2] / na};
    const float dot = an[0] * c[0] + an[1] * c[1] + an[2] * c[2];
    const float u[3] = {c[0] - dot * an[0], c[1] - dot * an[1], c[2] - dot * an[2]};
    const float nu = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    const float b[3] = {u[0] / nu, u[1] / nu, u[2] / nu};
    float ga[3] = {gR[0], gR[3], gR[6]}, gb[3] = {gR[1], gR[4], gR[7]};
    const float gd[3] = {gR[2], gR[5], gR[8]};
    // d = a' x b
    ga[0] += b[1] * gd[2] - b[2] * gd[1]; ga[1] += b[2] * gd[0] - b[0] * gd[2]; ga[2] += b[0] * gd[1] - b[1] * gd[0];
    gb[0] += gd[1] * an[2] - gd[2] * an[1]; gb[1] += gd[2] * an[0] - gd[0] * an[2]; gb[2] += gd[0] * an[1] - gd[1] * an[0];
    // b = u / |u|
    const float bgb = b[0] * gb[0] + b[1] * gb[1] + b[2] * gb[2];
    const float gu[3] = {(gb[0] - b[0] * bgb) / nu, (gb[1] - b[1] * bgb) / nu, (gb[2] - b[2] * bgb) / nu};
    // u = c - (a'.c) a'
    const float agu = an[0] * gu[0] + an[1] * gu[1] + an[2] * gu[2];
    const float gc[3] = {gu[0] - an[0] * agu, gu[1] - an[1] * agu, gu[2] - an[2] * agu};
    for (int k = 0; k < 3; ++k) ga[k] -= dot * gu[k] + agu * c[k];
    // a' = a / |a|
    const float aga = an[0] * ga[0] + an[1] * ga[1] + an[2] * ga[2];
    float* o = gx + 3 + slot * 6;
    for (int k = 0; k < 3; ++k) { o[2 * k] = (ga[k] - an[k] * aga) / na; o[2 * k + 1] = gc[k]; }
  } else if (slot == 22) {
    gx[0] = gy[0]; gx[1] = gy[1]; gx[2] = gy[2];
  } else if (slot == 23) {
    for (int k = 0; k < 24; ++k) gx[135 + k] = gy[69 + k] + hpose_scale * x[135 + k];
  }
}

// d xb[m, :] += d xin[m, D : D + BD]   (the body-parameter slice of the next recurrence's input)
__global__ void __launch_bounds__(256)
reg_add_slice_kernel(const float* __restrict__ d_xin, int M, int K, int D, int BD, float* __restrict__ d_xb) {
  const int64_t total = (int64_t)M * BD;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    d_xb[i] += d_xin[(i / BD) * K + D + (i % BD)];
}

}  // namespace eg

struct EgRegTrain {
  int device = 0;
  EgRegressorDims d;
  RegLayout L;
  float *P = nullptr, *G = nullptr;
  EgLbs* lbs = nullptr;
  int cap = 0;
  std::vector<float*> owned;
  float *xin = nullptr, *hs = nullptr, *t1 = nullptr, *t2 = nullptr, *xbc = nullptr;   // saved per recurrence
  float *yb = nullptr, *mk = nullptr, *dmk = nullptr, *dyb = nullptr, *drot = nullptr;
  float *dxbc = nullptr, *dh = nullptr, *da = nullptr, *db = nullptr, *dc = nullptr, *din = nullptr;
  double* sums = nullptr;
};

static int reg_ws(EgRegTrain* h, int M) {
  if (M <= h->cap) return EG_OK;
  for (float* p : h->owned) cudaFree(p);
  h->owned.clear(); h->cap = 0;
  const EgRegressorDims& d = h->d;
  const size_t m = (size_t)M, H = d.h_dim, K = d.in_dim + d.body_dim + 10, NB = d.n_blocks, NR = d.n_recur, BD = d.body_dim;
  auto A = [&](float** p, size_t n) -> int {
    EG_CUDA_CHECK(cudaMalloc((void**)p, n * sizeof(float)));
    h->owned.push_back(*p);
    return EG_OK;
  };
  EG_TRY(A(&h->xin, NR * m * K)); EG_TRY(A(&h->hs, NR * (NB + 1) * m * H));
  EG_TRY(A(&h->t1, NR * NB * m * H)); EG_TRY(A(&h->t2, NR * NB * m * H)); EG_TRY(A(&h->xbc, (NR + 1) * m * BD));
  EG_TRY(A(&h->yb, m * 93)); EG_TRY(A(&h->mk, m * d.in_dim)); EG_TRY(A(&h->dmk, m * d.in_dim));
  EG_TRY(A(&h->dyb, m * 93)); EG_TRY(A(&h->drot, m * 198)); EG_TRY(A(&h->dxbc, m * BD));
  EG_TRY(A(&h->dh, m * H)); EG_TRY(A(&h->da, m * H)); EG_TRY(A(&h->db, m * H)); EG_TRY(A(&h->dc, m * H));
  EG_TRY(A(&h->din, m * K));
  h->cap = M;
  return EG_OK;
}

extern "C" int64_t eg_regressor_param_count(const EgRegressorDims* d) { return d ? make_reg_layout(*d).n_total : -1; }

extern "C" int eg_regressor_train_create(const EgRegressorDims* dims, float* params_flat, float* grads_flat, EgLbs* lbs,
                                         int device, EgRegTrain** out) {
  EG_REQUIRE(dims && params_flat && grads_flat && lbs && out, "null pointer");
  EG_REQUIRE(dims->in_dim == 201 && dims->body_dim == 159, "ssm2_67 markers and the 6-D body vector (use_cont) expected");
  EG_REQUIRE(dims->h_dim > 0 && dims->n_blocks > 0 && dims->n_recur > 0, "bad dims");
  EG_CUDA_CHECK(cudaSetDevice(device));
  EgRegTrain* h = new EgRegTrain();
  h->device = device; h->d = *dims; h->L = make_reg_layout(*dims); h->P = params_flat; h->G = grads_flat; h->lbs = lbs;
  if (cudaMalloc((void**)&h->sums, 4 * sizeof(double)) != cudaSuccess) { delete h; return set_error(EG_ERR_CUDA, "cudaMalloc failed"); }
  *out = h;
  return EG_OK;
}

extern "C" void eg_regressor_train_destroy(EgRegTrain* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (float* p : h->owned) cudaFree(p);
  cudaFree(h->sums);
  delete h;
}

namespace {
int lin_bwd_pg(const float* P, float* G, cudaStream_t st, const float* dY, int ld_dy, const float* X, int ldx, int B, int64_t w,
               int64_t b, int in, int out, int ldw, float* dX, int ld_dx, int dx_beta, bool wgrad = true) {
  if (wgrad) {
    GemmArgs gw{dY, ld_dy, 1, X, ldx, G + w, ldw, nullptr, nullptr, 0, out, in, B, ACT_NONE, 0.f, 1, 1.0f};
    EG_TRY(launch_gemm(gw, true, false, st));
    EG_LAUNCH_PDL(colsum_kernel, dim3((out + 31) / 32, 4), 256, 0, st, dY, ld_dy, B, out, G + b);
  }
  if (dX) {
    GemmArgs gx{dY, ld_dy, 1, P + w, ldw, dX, ld_dx, nullptr, nullptr, 0, B, in, out, ACT_NONE, 0.f, dx_beta, 1.0f};
    EG_TRY(launch_gemm(gx, false, false, st));
  }
  return EG_OK;
}

// regressor forward over its recurrences (activations kept), 6-D -> axis-angle (h->yb) and the SMPL-X markers (h->mk)
int reg_forward(EgRegTrain* h, const float* markers, const float* betas, int M, void* stream) {
  cudaStream_t st = as_stream(stream);
  const EgRegressorDims& d = h->d;
  const RegLayout& L = h->L;
  const int D = d.in_dim, BD = d.body_dim, H = d.h_dim, NB = d.n_blocks, NR = d.n_recur, K = D + BD + 10;
  const float* P = h->P;
  const int64_t MH = (int64_t)M * H, MK = (int64_t)M * K, MB = (int64_t)M * BD;
  const int64_t blk_sz = 2 * ((int64_t)H * H + H);
  EG_CUDA_CHECK(cudaMemsetAsync(h->xbc, 0, MB * sizeof(float), st));
  for (int r = 0; r < NR; ++r) {
    float* xin = h->xin + r * MK;
    float* hr = h->hs + (int64_t)r * (NB + 1) * MH;
    const float* xb_r = h->xbc + r * MB;
    EG_LAUNCH(reg_concat_kernel, ew_grid(MK), 256, 0, st, markers, xb_r, betas, M, D, BD, xin);
    EG_TRY(linear(st, xin, K, M, P + L.in_w, K, P + L.in_b, K, H, hr, H));
    for (int b = 0; b < NB; ++b) {
      const int64_t w1 = L.blk + b * blk_sz, b1 = w1 + (int64_t)H * H, w2 = b1 + H, b2 = w2 + (int64_t)H * H;
      float* t1 = h->t1 + ((int64_t)r * NB + b) * MH;
      float* t2 = h->t2 + ((int64_t)r * NB + b) * MH;
      EG_TRY(linear(st, hr + b * MH, H, M, P + w1, H, P + b1, H, H, t1, H, ACT_RELU));
      EG_TRY(linear(st, t1, H, M, P + w2, H, P + b2, H, H, t2, H, ACT_RELU));
      EG_LAUNCH_PDL(add_kernel, ew_grid(MH), 256, 0, st, (const float*)t2, (const float*)(hr + b * MH), MH, hr + (b + 1) * MH);
    }
    EG_TRY(linear(st, hr + NB * MH, H, M, P + L.out_w, H, P + L.out_b, H, BD, h->xbc + (r + 1) * MB, BD, ACT_NONE, 0.f, xb_r, BD));
  }
  EG_LAUNCH(reg_cont2aa_kernel, (M * 32 + 127) / 128, 128, 0, st, (const float*)(h->xbc + NR * MB), M, h->yb);
  return eg_lbs_forward(h->lbs, h->yb, betas, M, M, nullptr, nullptr, h->mk, stream);
}

// backward from h->dmk (dL/d SMPL-X markers) and the hand regulariser to the parameters (wgrad) and / or to the marker
// input of the regressor (d_in [M,D], overwritten)
int reg_backward(EgRegTrain* h, const float* betas, int M, float hpose_scale, bool wgrad, float* d_in, void* stream) {
  cudaStream_t st = as_stream(stream);
  const EgRegressorDims& d = h->d;
  const RegLayout& L = h->L;
  const int D = d.in_dim, BD = d.body_dim, H = d.h_dim, NB = d.n_blocks, NR = d.n_recur, K = D + BD + 10;
  const float* P = h->P;
  float* G = h->G;
  const int64_t MH = (int64_t)M * H, MK = (int64_t)M * K, MB = (int64_t)M * BD;
  const int64_t blk_sz = 2 * ((int64_t)H * H + H);
  const float* xb_fin = h->xbc + NR * MB;
  if (wgrad) EG_CUDA_CHECK(cudaMemsetAsync(G, 0, (size_t)L.n_total * sizeof(float), st));
  if (d_in) EG_CUDA_CHECK(cudaMemsetAsync(d_in, 0, (size_t)M * D * sizeof(float), st));
  EG_TRY(eg_lbs_markers_backward_rot(h->lbs, h->yb, betas, M, M, h->dmk, h->dyb, h->drot, stream));
  EG_LAUNCH(reg_gs_bwd_kernel, (M * 32 + 127) / 128, 128, 0, st, xb_fin, (const float*)h->drot, (const float*)h->dyb, M,
            hpose_scale, h->dxbc);
  for (int r = NR - 1; r >= 0; --r) {
    const float* xin = h->xin + r * MK;
    const float* hr = h->hs + (int64_t)r * (NB + 1) * MH;
    EG_TRY(lin_bwd_pg(P, G, st, h->dxbc, BD, hr + NB * MH, H, M, L.out_w, L.out_b, H, BD, H, h->dh, H, 0, wgrad));
    for (int b = NB - 1; b >= 0; --b) {
      const int64_t w1 = L.blk + b * blk_sz, b1 = w1 + (int64_t)H * H, w2 = b1 + H, b2 = w2 + (int64_t)H * H;
      const float* t1 = h->t1 + ((int64_t)r * NB + b) * MH;
      const float* t2 = h->t2 + ((int64_t)r * NB + b) * MH;
      EG_LAUNCH_PDL(lrelu_bwd_kernel, ew_grid(MH), 256, 0, st, (const float*)h->dh, t2, 0.0f, MH, h->da);
      EG_TRY(lin_bwd_pg(P, G, st, h->da, H, t1, H, M, w2, b2, H, H, H, h->db, H, 0, wgrad));
      EG_LAUNCH_PDL(lrelu_bwd_kernel, ew_grid(MH), 256, 0, st, (const float*)h->db, t1, 0.0f, MH, h->dc);
      EG_TRY(lin_bwd_pg(P, G, st, h->dc, H, hr + b * MH, H, M, w1, b1, H, H, H, h->dh, H, 1, wgrad));   // + the skip path
    }
    const bool need_din = r > 0 || d_in != nullptr;
    EG_TRY(lin_bwd_pg(P, G, st, h->dh, H, xin, K, M, L.in_w, L.in_b, K, H, K, need_din ? h->din : nullptr, K, 0, wgrad));
    if (r > 0) EG_LAUNCH(reg_add_slice_kernel, ew_grid(MB), 256, 0, st, (const float*)h->din, M, K, D, BD, h->dxbc);
    if (d_in) EG_LAUNCH(reg_add_slice_kernel, ew_grid((int64_t)M * D), 256, 0, st, (const float*)h->din, M, K, 0, D, d_in);
  }
  return EG_OK;
}
}  // namespace

extern "C" int eg_regressor_loss_backward(EgRegTrain* h, const float* marker_ref, const float* betas, int M,
                                          float w_hpose, float* xb_out, float* stats, void* stream) {
  EG_REQUIRE(h && marker_ref && betas && stats && M > 0, "bad arguments");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  EG_TRY(reg_ws(h, M));
  cudaStream_t st = as_stream(stream);
  const int D = h->d.in_dim;
  EG_TRY(reg_forward(h, marker_ref, betas, M, stream));
  if (xb_out) EG_CUDA_CHECK(cudaMemcpyAsync(xb_out, h->yb, (size_t)M * 93 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EG_CUDA_CHECK(cudaMemsetAsync(h->sums, 0, 4 * sizeof(double), st));
  EG_LAUNCH(reg_l1_kernel, ew_grid((int64_t)M * D), 256, 0, st, (const float*)h->mk, marker_ref, (int64_t)M * D, h->dmk, h->sums);
  EG_LAUNCH(reg_hpose_kernel, ew_grid((int64_t)M * 24), 256, 0, st, (const float*)(h->xbc + (int64_t)h->d.n_recur * M * h->d.body_dim),
            M, h->sums + 1);
  EG_LAUNCH(reg_stats_kernel, 1, 1, 0, st, (const double*)h->sums, (int64_t)M * D, (int64_t)M * 24, w_hpose, stats);
  return reg_backward(h, betas, M, w_hpose * 2.0f / (float)((int64_t)M * 24), true, nullptr, stream);
}

// Regressor part of the combo objective (GAMMAPrimitiveComboTrainOP.calc_loss_regressor, models_GAMMA_primitive.py:787-794):
// markers_in [T*B,201] (t-major, the predictor's Y_rec) -> body parameters -> SMPL-X markers x_pred;
//   loss = w_rec L1(Y_ref, x_pred) + w_td L1(dt x_pred, dt Y_ref) + w_hpose mean(hand^2)   (x loss_scale)
// and its gradient w.r.t. markers_in (the regressor's weights are not trained by that op). want_grad = 0 only evaluates
// the two loss terms (the torch.no_grad() branch). stats (device float[2], ACCUMULATED): marker term, hand term.
__global__ void reg_cycle_stats_kernel(const double* __restrict__ sums, int T, int64_t BD, int64_t n_hpose, float w_rec,
                                       float w_td, float scale, float* __restrict__ stats) {
  const float rec = w_rec * (float)(sums[0] / (double)((int64_t)T * BD)) + w_td * (float)(sums[1] / (double)((int64_t)(T - 1) * BD));
  atomicAdd(stats + 0, rec * scale);
  atomicAdd(stats + 1, (float)(sums[2] / (double)n_hpose) * scale);
}

extern "C" int eg_regressor_cycle_backward(EgRegTrain* h, const float* markers_in, const float* betas, const float* Y_ref,
                                           int T, int B, float w_rec, float w_td, float w_hpose, float loss_scale,
                                           int want_grad, float* d_markers_in, float* xb_out, float* stats, void* stream) {
  EG_REQUIRE(h && markers_in && betas && Y_ref && stats && T > 1 && B > 0, "bad arguments");
  EG_REQUIRE(!want_grad || d_markers_in, "d_markers_in required when want_grad");
  EG_CUDA_CHECK(cudaSetDevice(h->device));
  const int M = T * B, D = h->d.in_dim;
  EG_TRY(reg_ws(h, M));
  cudaStream_t st = as_stream(stream);
  EG_TRY(reg_forward(h, markers_in, betas, M, stream));
  if (xb_out) EG_CUDA_CHECK(cudaMemcpyAsync(xb_out, h->yb, (size_t)M * 93 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EG_CUDA_CHECK(cudaMemsetAsync(h->sums, 0, 4 * sizeof(double), st));
  const int64_t BDm = (int64_t)B * D;
  EG_LAUNCH(rec_loss_grad_kernel, ew_grid((int64_t)T * BDm), 256, 0, st, Y_ref, (const float*)h->mk, T, BDm, w_rec, w_td,
            want_grad ? loss_scale : 0.0f, h->dmk, h->sums);
  EG_LAUNCH(reg_hpose_kernel, ew_grid((int64_t)M * 24), 256, 0, st, (const float*)(h->xbc + (int64_t)h->d.n_recur * M * h->d.body_dim),
            M, h->sums + 2);
  EG_LAUNCH(reg_cycle_stats_kernel, 1, 1, 0, st, (const double*)h->sums, T, BDm, (int64_t)M * 24, w_rec, w_td, loss_scale, stats);
  if (!want_grad) return EG_OK;
  return reg_backward(h, betas, M, w_hpose * loss_scale * 2.0f / (float)((int64_t)M * 24), false, d_markers_in, stream);
}
