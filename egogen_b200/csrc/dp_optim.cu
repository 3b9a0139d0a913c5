// Data-parallel optimiser step over NVLink peer / NVSwitch multicast memory (one process per GPU).
//
// Replaces, for N > 1 ranks, the pair  all_reduce(flat_grads) ; eg_clip_adamw_step  of the PPO update (reference
// motion/crowd_ppo/ppo_policy.py:241-247: backward -> clip_grad_norm_(actor + critic) -> optim.step, one process) by a
// sharded form in which every rank owns 1/N of the flat parameter vector:
//   eg_dp_reduce_norm   rank r sums ITS slice of the gradient over all ranks - one multimem.ld_reduce per 16 bytes when the
//                       buffers have a multicast mapping (the NVSwitch adds the N copies in flight), else N peer loads summed
//                       in rank order - keeps the reduced slice locally, and publishes its share of the squared gradient norm
//                       of the clipped range into slot r of every rank's scratch;
//   eg_dp_adamw_gather  after a cross-rank barrier: total norm = sum of the N slots (rank order, so every rank computes the
//                       same bits), clip + AdamW on the own slice with the local moment slices, and the updated parameters
//                       are written to every rank (multimem.st, or N peer stores).
// Per rank and step the links carry one gradient vector in and one parameter vector out, in 1/N-sized pieces that start as
// soon as the backward has finished, and the 369 MB optimiser pass of the single-GPU path shrinks to 1/N.
// The buffers are torch symmetric-memory allocations (the caller passes the peer pointers and, when available, the
// multicast pointer); the three barriers of a step are the caller's (symmetric-memory signal pads).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace eg {

constexpr int kMaxRanks = 16;
struct PeerPtrs { void* p[kMaxRanks]; };

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// slice of rank r: elements [r * chunk, (r + 1) * chunk), chunk = n_pad / world (a multiple of 4)
template <bool MC>
__global__ void __launch_bounds__(256)
dp_reduce_norm_kernel(PeerPtrs grads, const float* grads_mc, int world, int rank, int64_t chunk, int64_t n_clip,
                      float* __restrict__ gred, PeerPtrs scratch, double* partial, unsigned* ticket, int64_t i_lo, int64_t i_hi,
                      int publish) {
  // i_lo / i_hi: the part of this rank's slice to reduce now (slice-relative, multiples of 4)
  const int64_t base = (int64_t)rank * chunk;
  double q = 0.0;
  // 4 independent 16-byte (multicast / peer) loads in flight per thread: the reduce is bound by NVLink latency x bytes in
  // flight, and a grid that is small enough to leave SMs to a concurrent backward still has to keep ~1 MB outstanding
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i0 = i_lo + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i0 < i_hi; i0 += stride * U) {
    float4 g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < i_hi) {
        if (MC) {
          g[u] = mc_ld_reduce(grads_mc + base + i);
        } else {
          for (int r = 0; r < world; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(grads.p[r]) + base + i);
            g[u].x += v.x; g[u].y += v.y; g[u].z += v.z; g[u].w += v.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < i_hi) {
        *reinterpret_cast<float4*>(gred + i) = g[u];
        const int64_t gi = base + i;
        if (gi < n_clip) q += (double)g[u].x * g[u].x;
        if (gi + 1 < n_clip) q += (double)g[u].y * g[u].y;
        if (gi + 2 < n_clip) q += (double)g[u].z * g[u].z;
        if (gi + 3 < n_clip) q += (double)g[u].w * g[u].w;
      }
    }
  }
  if (!publish) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __shared__ double pq[8];
  __shared__ int is_last;
  if ((threadIdx.x & 31) == 0) pq[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    q = 0.0;
    for (int w = 0; w < 8; ++w) q += pq[w];
    atomicAdd(partial, q);
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    // every block's contribution is in `partial`: publish it to slot `rank` of every rank's scratch, re-arm the counters
    const double tot = *reinterpret_cast<volatile double*>(partial);
    for (int r = 0; r < world; ++r) static_cast<double*>(scratch.p[r])[rank] = tot;
    *partial = 0.0;
    *ticket = 0u;
    __threadfence_system();
  }
}

template <bool MC>
__global__ void __launch_bounds__(256)
dp_adamw_gather_kernel(PeerPtrs params, float* params_mc, int world, int rank, int64_t chunk, int64_t n_clip,
                       const float* __restrict__ gred, const double* __restrict__ scratch_local, float* __restrict__ m,
                       float* __restrict__ v, float max_norm, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                       float bc2_sqrt) {
  float clip = 1.0f;
  if (max_norm > 0.0f) {
    double tot = 0.0;
    for (int r = 0; r < world; ++r) tot += scratch_local[r];
    clip = fminf(max_norm / ((float)sqrt(tot) + 1e-6f), 1.0f);
  }
  const int64_t base = (int64_t)rank * chunk;
  const float* p_local = static_cast<const float*>(params.p[rank]);
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < chunk; i += (int64_t)gridDim.x * blockDim.x * 4) {
    const float4 g4 = *reinterpret_cast<const float4*>(gred + i);
    const float4 p4 = *reinterpret_cast<const float4*>(p_local + base + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
    float gs[4] = {g4.x, g4.y, g4.z, g4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
    float ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {                  // same arithmetic as adamw_kernel (ppo.cu)
      float gi = gs[k];
      if (base + i + k < n_clip) gi *= clip;
      float pi = ps[k] * (1.0f - lr * wd);
      const float mi = beta1 * ms[k] + (1.0f - beta1) * gi;
      const float vi = beta2 * vs[k] + (1.0f - beta2) * gi * gi;
      ms[k] = mi; vs[k] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pi -= (lr / bc1) * (mi / denom);
      ps[k] = pi;
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vs[0], vs[1], vs[2], vs[3]);
    const float4 out = make_float4(ps[0], ps[1], ps[2], ps[3]);
    if (MC) {
      mc_st(params_mc + base + i, out);
    } else {
      for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(static_cast<float*>(params.p[r]) + base + i) = out;
    }
  }
}

}  // namespace eg

using namespace eg;

static int fill_peers(PeerPtrs& pp, const void* const* ptrs, int world) {
  EG_REQUIRE(ptrs != nullptr && world >= 1 && world <= kMaxRanks, "1..16 ranks with a peer pointer each");
  for (int r = 0; r < kMaxRanks; ++r) pp.p[r] = r < world ? const_cast<void*>(ptrs[r]) : nullptr;
  for (int r = 0; r < world; ++r) EG_REQUIRE(pp.p[r] != nullptr, "null peer pointer");
  return EG_OK;
}

extern "C" int eg_dp_reduce_range(const void* const* grads_ptrs, const void* grads_mc, int world, int rank, int64_t n_pad,
                                  int64_t n_clip, int64_t elem_lo, int64_t elem_hi, int publish_norm, float* gred,
                                  const void* const* scratch_ptrs, void* work, void* stream) {
  EG_REQUIRE(gred && work && rank >= 0 && rank < world, "bad arguments");
  EG_REQUIRE(n_pad > 0 && n_pad % (4 * (int64_t)world) == 0, "n_pad must be a multiple of 4 * world");
  EG_REQUIRE(elem_lo >= 0 && elem_lo <= elem_hi && elem_hi <= n_pad && elem_lo % 4 == 0 && elem_hi % 4 == 0, "element range must be 4-aligned");
  PeerPtrs g, s;
  int rc;
  if ((rc = fill_peers(g, grads_ptrs, world))) return rc;
  if ((rc = fill_peers(s, scratch_ptrs, world))) return rc;
  const int64_t chunk = n_pad / world, base = (int64_t)rank * chunk;
  const int64_t i_lo = std::min(std::max(elem_lo - base, (int64_t)0), chunk), i_hi = std::min(std::max(elem_hi - base, (int64_t)0), chunk);
  if (i_hi <= i_lo && !publish_norm) return EG_OK;             // nothing of the range lies in this rank's slice
  // EG_DP_REDUCE_BLOCKS caps the grid (default: 4 CTAs per SM; a smaller grid leaves SMs to a concurrent backward in the
  // opt-in overlapped mode, EG_DP_OVERLAP=1)
  static const int cap = getenv("EG_DP_REDUCE_BLOCKS") ? std::max(1, atoi(getenv("EG_DP_REDUCE_BLOCKS"))) : kNumSMs * 4;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((i_hi - i_lo) / 16 + 255) / 256, cap));
  double* partial = static_cast<double*>(work);
  unsigned* ticket = reinterpret_cast<unsigned*>(partial + 1);
  if (grads_mc) EG_LAUNCH(dp_reduce_norm_kernel<true>, grid, 256, 0, as_stream(stream), g, static_cast<const float*>(grads_mc), world, rank,
                          chunk, n_clip, gred, s, partial, ticket, i_lo, i_hi, publish_norm);
  else EG_LAUNCH(dp_reduce_norm_kernel<false>, grid, 256, 0, as_stream(stream), g, nullptr, world, rank, chunk, n_clip, gred, s,
                 partial, ticket, i_lo, i_hi, publish_norm);
  return EG_OK;
}

extern "C" int eg_dp_reduce_norm(const void* const* grads_ptrs, const void* grads_mc, int world, int rank, int64_t n_pad,
                                 int64_t n_clip, float* gred, const void* const* scratch_ptrs, void* work, void* stream) {
  return eg_dp_reduce_range(grads_ptrs, grads_mc, world, rank, n_pad, n_clip, 0, n_pad, 1, gred, scratch_ptrs, work, stream);
}

extern "C" int eg_dp_adamw_gather(const void* const* params_ptrs, void* params_mc, int world, int rank, int64_t n_pad,
                                  int64_t n_clip, const float* gred, const double* scratch_local, float* exp_avg_slice,
                                  float* exp_avg_sq_slice, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int step, void* stream) {
  EG_REQUIRE(gred && scratch_local && exp_avg_slice && exp_avg_sq_slice && rank >= 0 && rank < world && step >= 1, "bad arguments");
  EG_REQUIRE(n_pad > 0 && n_pad % (4 * (int64_t)world) == 0, "n_pad must be a multiple of 4 * world");
  PeerPtrs p;
  int rc;
  if ((rc = fill_peers(p, params_ptrs, world))) return rc;
  const int64_t chunk = n_pad / world;
  const int grid = (int)std::min<int64_t>((chunk / 4 + 255) / 256, kNumSMs * 4);
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  if (params_mc) EG_LAUNCH(dp_adamw_gather_kernel<true>, grid, 256, 0, as_stream(stream), p, static_cast<float*>(params_mc), world, rank,
                           chunk, n_clip, gred, scratch_local, exp_avg_slice, exp_avg_sq_slice, max_grad_norm, lr, beta1, beta2,
                           eps, weight_decay, bc1, bc2s);
  else EG_LAUNCH(dp_adamw_gather_kernel<false>, grid, 256, 0, as_stream(stream), p, nullptr, world, rank, chunk, n_clip, gred,
                 scratch_local, exp_avg_slice, exp_avg_sq_slice, max_grad_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2s);
  return EG_OK;
}
