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
#include <vector>

#include "nn.cuh"

namespace eg {

struct Lin { int64_t w, b; int in, out; };   // offsets into the flat buffers

struct PolicyLayout {
  Lin a_blk[4][2], a_out, c_blk[4][2], c_out;
  int64_t x_wih, x_whh, x_bih, x_bhh, e_wih, e_whh, e_bih, e_bhh;
  int64_t n_actor_critic, n_total;
  int hx_dim;
};

static PolicyLayout make_layout(const EgPolicyDims& d) {
  PolicyLayout L{};
  const int D = 2 * d.h_dim + 4 * d.pe_L;    // 1152
  L.hx_dim = D;
  int64_t off = 0;
  auto lin = [&](int in, int out) { Lin l{off, off + (int64_t)in * out, in, out}; off += (int64_t)in * out + out; return l; };
  for (int k = 0; k < d.n_blocks; ++k) { L.a_blk[k][0] = lin(D, D); L.a_blk[k][1] = lin(D, D); }
  L.a_out = lin(D, 2 * d.z_dim);
  for (int k = 0; k < d.n_blocks; ++k) { L.c_blk[k][0] = lin(D, D); L.c_blk[k][1] = lin(D, D); }
  L.c_out = lin(D, 1);
  L.n_actor_critic = off;
  const int H = d.h_dim, H3 = 3 * d.h_dim;
  L.x_wih = off; off += (int64_t)H3 * d.in_dim;
  L.x_whh = off; off += (int64_t)H3 * H;
  L.x_bih = off; off += H3;
  L.x_bhh = off; off += H3;
  L.e_wih = off; off += (int64_t)H3 * d.ego_dim;
  L.e_whh = off; off += (int64_t)H3 * H;
  L.e_bih = off; off += H3;
  L.e_bhh = off; off += H3;
  L.n_total = off;
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
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.0f ? dy[i] : dy[i] * slope;
}

__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ c) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c[i] = a[i] + b[i];
}

// db[n] += sum_m dY[m, n]   (block = 32 columns x 8 row lanes)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dY, int ld, int M, int N, float* __restrict__ db) {
  __shared__ float part[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x % 32, rl = threadIdx.x / 32;
  float s = 0.0f;
  if (col < N)
    for (int m = rl; m < M; m += 8) s += dY[(int64_t)m * ld + col];
  part[rl][threadIdx.x % 32] = s;
  __syncthreads();
  if (rl == 0 && col < N) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x];
    db[col] += t;
  }
}

// GRU cell backward given dh (grad of the cell output); writes dgi, dgh [M,3H] and dh_prev = dh * z
__global__ void __launch_bounds__(256)
gru_bwd_kernel(const float* __restrict__ dh, int ld_dh, const float* __restrict__ r, const float* __restrict__ z,
               const float* __restrict__ n, const float* __restrict__ ghn, const float* __restrict__ h_prev,
               int M, int H, float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh_prev) {
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
  h->cap = B;
  return EG_OK;
}

extern "C" int64_t eg_policy_param_count(const EgPolicyDims* d, int64_t* n_actor_critic) {
  if (!d) return -1;
  PolicyLayout L = make_layout(*d);
  if (n_actor_critic) *n_actor_critic = L.n_actor_critic;
  return L.n_total;
}

extern "C" int eg_policy_create(const EgPolicyDims* dims, float* params_flat, float* grads_flat, int device,
                                EgPolicy** out) {
  EG_REQUIRE(dims && params_flat && out, "null pointer");
  EG_REQUIRE(dims->n_blocks >= 1 && dims->n_blocks <= 4, "n_blocks must be in [1,4]");
  EG_CUDA_CHECK(cudaSetDevice(device));
  EgPolicy* h = new EgPolicy();
  h->device = device; h->d = *dims; h->L = make_layout(*dims); h->P = params_flat; h->G = grads_flat;
  EG_CUDA_CHECK(cudaMalloc((void**)&h->mom, 4 * sizeof(double)));
  *out = h;
  return EG_OK;
}

extern "C" void eg_policy_destroy(EgPolicy* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (float* p : h->owned) cudaFree(p);
  cudaFree(h->mom);
  delete h;
}

// one GRU encoder over the 2 observation frames (h0 = 0), saving gate activations for backward
static int gru2_forward(EgPolicy* h, cudaStream_t st, const float* x, int ld_env, int ld_frame, int in_dim, int B,
                        int64_t wih, int64_t whh, int64_t bih, int64_t bhh, float** r, float** z, float** n, float** g,
                        float* h1, float* out, int ld_out) {
  const int H = h->d.h_dim, H3 = 3 * H;
  const float* P = h->P;
  EG_TRY(linear(st, x, ld_env, B, P + wih, in_dim, P + bih, in_dim, H3, h->gi, H3));
  EG_TRY(launch_gru_gate(st, h->gi, nullptr, P + bhh, nullptr, h1, B, H, H, r[0], z[0], n[0], g[0]));
  EG_TRY(linear(st, x + ld_frame, ld_env, B, P + wih, in_dim, P + bih, in_dim, H3, h->gi, H3));
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
    EG_LAUNCH(add_kernel, ew_grid((int64_t)B * D), 256, 0, st, u[k], in[k], (int64_t)B * D, in[k + 1]);
  }
  EG_TRY(linear(st, in[h->d.n_blocks], D, B, P + outl.w, D, P + outl.b, D, outl.out, out, outl.out));
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
  EG_TRY(gru2_forward(h, st, state, 2 * d.in_dim, d.in_dim, d.in_dim, B, L.x_wih, L.x_whh, L.x_bih, L.x_bhh, h->xr,
                      h->xz, h->xn, h->xg, h->xh1, h->hx, D));
  EG_TRY(gru2_forward(h, st, ego, 2 * d.ego_dim, d.ego_dim, d.ego_dim, B, L.e_wih, L.e_whh, L.e_bih, L.e_bhh, h->er,
                      h->ez, h->en, h->eg_, h->eh1, h->hx + H, D));
  EG_LAUNCH(pe_kernel, (B * 2 * d.pe_L + 127) / 128, 128, 0, st, dist, time, B, d.pe_L, D, 2 * H, h->hx);
  if (want_actor) {
    EG_TRY(mlp_block_forward(h, st, L.a_blk, L.a_out, h->a_in, h->a_t, h->a_u, h->out_a, B));
    if (out_actor) EG_CUDA_CHECK(cudaMemcpyAsync(out_actor, h->out_a, (size_t)B * 2 * d.z_dim * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (want_critic) {
    EG_TRY(mlp_block_forward(h, st, L.c_blk, L.c_out, h->c_in, h->c_t, h->c_u, h->out_c, B));
    if (value) EG_CUDA_CHECK(cudaMemcpyAsync(value, h->out_c, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
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

// dY [B,out] -> dW += dY^T X, db += colsum(dY), dX = dY W (+ optional accumulate into dX)
static int linear_backward(EgPolicy* h, cudaStream_t st, const float* dY, int ld_dy, const float* X, int ldx, int B,
                           const Lin& l, float* dX, int ld_dx, int dx_beta) {
  const float* P = h->P;
  float* G = h->G;
  GemmArgs gw{dY, ld_dy, 1, X, ldx, G + l.w, l.in, nullptr, nullptr, 0, l.out, l.in, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(gw, true, false, st));
  EG_LAUNCH(colsum_kernel, (l.out + 31) / 32, 256, 0, st, dY, ld_dy, B, l.out, G + l.b);
  if (dX) {
    GemmArgs gx{dY, ld_dy, 1, P + l.w, l.in, dX, ld_dx, nullptr, nullptr, 0, B, l.in, l.out, ACT_NONE, 0.f, dx_beta, 1.0f};
    EG_TRY(launch_gemm(gx, false, false, st));
  }
  return EG_OK;
}

static int mlp_block_backward(EgPolicy* h, cudaStream_t st, const Lin blk[][2], const Lin& outl, float** in, float** t,
                              float** u, const float* d_out, int B, float* dhx, int dhx_beta) {
  const int D = h->L.hx_dim;
  const int64_t n = (int64_t)B * D;
  // out = in[nb] W_o^T + b_o
  EG_TRY(linear_backward(h, st, d_out, outl.out, in[h->d.n_blocks], D, B, outl, h->dh, D, 0));
  for (int k = h->d.n_blocks - 1; k >= 0; --k) {
    // in[k+1] = u + in[k];  u = lrelu(t W2^T + b2);  t = lrelu(in[k] W1^T + b1)
    EG_LAUNCH(lrelu_bwd_kernel, ew_grid(n), 256, 0, st, h->dh, u[k], 0.01f, n, h->da);
    EG_TRY(linear_backward(h, st, h->da, D, t[k], D, B, blk[k][1], h->dt, D, 0));
    EG_LAUNCH(lrelu_bwd_kernel, ew_grid(n), 256, 0, st, h->dt, t[k], 0.01f, n, h->da);
    // d in[k] = da1 W1 + dh (residual): accumulate straight into dh
    EG_TRY(linear_backward(h, st, h->da, D, in[k], D, B, blk[k][0], h->dh, D, 1));
  }
  if (dhx_beta) EG_LAUNCH(add_kernel, ew_grid(n), 256, 0, st, dhx, h->dh, n, dhx);
  else EG_CUDA_CHECK(cudaMemcpyAsync(dhx, h->dh, n * 4, cudaMemcpyDeviceToDevice, st));
  return EG_OK;
}

static int gru2_backward(EgPolicy* h, cudaStream_t st, const float* x, int ld_env, int ld_frame, int in_dim, int B,
                         int64_t wih, int64_t whh, int64_t bih, int64_t bhh, float** r, float** z, float** n, float** g,
                         const float* h1, const float* dh2, int ld_dh2) {
  const int H = h->d.h_dim, H3 = 3 * H;
  const float* P = h->P;
  float* G = h->G;
  // step 2
  EG_LAUNCH(gru_bwd_kernel, ew_grid((int64_t)B * H), 256, 0, st, dh2, ld_dh2, r[1], z[1], n[1], g[1], h1, B, H, h->dgi,
            h->dgh, h->dh1);
  GemmArgs w1{h->dgi, H3, 1, x + ld_frame, ld_env, G + wih, in_dim, nullptr, nullptr, 0, H3, in_dim, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w1, true, false, st));
  EG_LAUNCH(colsum_kernel, (H3 + 31) / 32, 256, 0, st, h->dgi, H3, B, H3, G + bih);
  GemmArgs w2{h->dgh, H3, 1, h1, H, G + whh, H, nullptr, nullptr, 0, H3, H, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w2, true, false, st));
  EG_LAUNCH(colsum_kernel, (H3 + 31) / 32, 256, 0, st, h->dgh, H3, B, H3, G + bhh);
  // dh1 = dh2 * z + dgh W_hh
  GemmArgs x2{h->dgh, H3, 1, P + whh, H, h->dh1, H, nullptr, nullptr, 0, B, H, H3, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(x2, false, false, st));
  // step 1 (h0 = 0: no W_hh gradient, b_hh still receives dgh)
  EG_LAUNCH(gru_bwd_kernel, ew_grid((int64_t)B * H), 256, 0, st, h->dh1, H, r[0], z[0], n[0], g[0], nullptr, B, H,
            h->dgi, h->dgh, nullptr);
  GemmArgs w3{h->dgi, H3, 1, x, ld_env, G + wih, in_dim, nullptr, nullptr, 0, H3, in_dim, B, ACT_NONE, 0.f, 1, 1.0f};
  EG_TRY(launch_gemm(w3, true, false, st));
  EG_LAUNCH(colsum_kernel, (H3 + 31) / 32, 256, 0, st, h->dgi, H3, B, H3, G + bih);
  EG_LAUNCH(colsum_kernel, (H3 + 31) / 32, 256, 0, st, h->dgh, H3, B, H3, G + bhh);
  return EG_OK;
}

extern "C" int eg_ppo_loss_backward(EgPolicy* h, const float* state, const float* ego, const float* dist,
                                    const float* time, const float* act, const float* logp_old, const float* adv_norm,
                                    const float* returns, int B, float inv_B, float eps_clip, float vf_coef,
                                    float ent_coef, float min_logvar, float max_logvar, int zero_grads,
                                    float* stats, void* stream) {
  EG_REQUIRE(h && h->G && act && logp_old && adv_norm && returns && stats, "null pointer (was the policy created with a gradient buffer?)");
  if (B <= 0) return EG_OK;
  cudaStream_t st = as_stream(stream);
  const EgPolicyDims& d = h->d;
  const PolicyLayout& L = h->L;
  const int H = d.h_dim, D = L.hx_dim;
  EG_TRY(eg_policy_forward(h, state, ego, dist, time, B, 1, 1, nullptr, nullptr, nullptr, stream));
  if (zero_grads) EG_CUDA_CHECK(cudaMemsetAsync(h->G, 0, (size_t)L.n_total * 4, st));
  EG_LAUNCH(ppo_head_kernel, B, 128, 0, st, h->out_a, h->out_c, act, logp_old, adv_norm, returns, B, d.z_dim, inv_B,
            eps_clip, vf_coef, ent_coef, min_logvar, max_logvar, h->d_out_a, h->d_out_c, stats);
  EG_TRY(mlp_block_backward(h, st, L.a_blk, L.a_out, h->a_in, h->a_t, h->a_u, h->d_out_a, B, h->dhx, 0));
  EG_TRY(mlp_block_backward(h, st, L.c_blk, L.c_out, h->c_in, h->c_t, h->c_u, h->d_out_c, B, h->dhx, 1));
  EG_TRY(gru2_backward(h, st, state, 2 * d.in_dim, d.in_dim, d.in_dim, B, L.x_wih, L.x_whh, L.x_bih, L.x_bhh, h->xr,
                       h->xz, h->xn, h->xg, h->xh1, h->dhx, D));
  EG_TRY(gru2_backward(h, st, ego, 2 * d.ego_dim, d.ego_dim, d.ego_dim, B, L.e_wih, L.e_whh, L.e_bih, L.e_bhh, h->er,
                       h->ez, h->en, h->eg_, h->eh1, h->dhx + H, D));
  return EG_OK;
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
