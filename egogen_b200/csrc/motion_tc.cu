// tcgen05 / TMEM / TMA kernels for the two sequential layer chains of GAMMAPrimitiveCombo.sample_prior (sm_100a).
//
// Precision scheme (both kernels): every product A*W runs as three kind::f16 passes over FP16 hi/lo splits,
//      x = hi + lo,  hi = fp16(x),  lo = fp16(x - hi):      A*W ~= A_hi*W_hi + A_hi*W_lo + A_lo*W_hi     (fp32 accumulation, TMEM)
// Weights are multiplied by a per-matrix power of two before the split (max |w| lands in [2^13, 2^14), so the lo part of
// every weight that matters is a normal fp16 number) and the accumulator is scaled back in the epilogue - exact.
// Activations of these nets are O(1) (tanh / gate outputs, marker coordinates in metres, ReLU features), far inside
// fp16's range; their lo part is exact to 2^-25 absolute. The result carries ~2^-22 relative error per product - the class
// of the 3xTF32 dense layers (gemm_tc.cu) at half the shared-memory bytes and twice the tensor rate. A value beyond fp16's
// range turns into inf/NaN and surfaces in the outputs (no silent saturation).
//
// decode_tc_kernel - the 18 GRUCell + MLP steps (reference models_GAMMA_primitive.py:91-99) for 64 rows per 16-CTA
//   cluster (UMMA M stays 128; see dtc::ROWS). CTA j owns 1/16 of every layer's OUTPUT columns and keeps that weight slice (176 KB, hi + lo) resident in
//   shared memory for all steps; the cluster exchanges each layer's activations as fp16 hi/lo rows through L2
//   (st.global -> barrier.cluster release/acquire -> TMA loads into a 6-stage ring), so a step costs 3 hardware cluster
//   barriers instead of the 4 global arrival-counter rounds of the SIMT weight-stationary kernel (nn.cu). Step algebra:
//     gh_{t+1} = h_t Whh^T                      is produced together with t1_t = tanh(h_t W1^T + b1)      (phase A, K = 256)
//     t2_t     = tanh(t1_t W2^T + b2)                                                                   (phase B, K = 512)
//     y_t      = y_{t-1} + t2_t Wo^T + bo   and   gi_{t+1} = gi_t + t2_t (Wy Wo)^T + Wy bo              (phase C, K = 256)
//   i.e. the K = 201 product y_t Wy^T of the next step is folded through d_out (Wf = Wy Wo, built in double); the running
//   y and gi live in epilogue registers (fp32, round-to-nearest adds in the reference's order) and the GRU gates of step
//   t+1 run in the epilogue of phase C.
//
// regressor_tc_kernel - MoshRegressor._forward (3 recurrences x (in_fc + 10 residual blocks + out_fc), reference
//   models_GAMMA_primitive.py:160-175,222-259) for 128 marker frames per CTA: the activation tile lives in shared memory as
//   the UMMA A operand (rewritten in place by the epilogue in the swizzled K-major layout), the residual stream h stays in
//   registers, and the weights stream once per CTA through a 4-stage TMA ring.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "lbs_tc.cuh"
#include "motion_tc.cuh"

namespace eg {
namespace mtc {

using tc::elect_one;
using tc::make_desc;
using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_try_wait;
using tc::smem_u32;
using tc::tma_load_2d;
using tc::umma_commit;

// a wait that cannot hang the GPU: a barrier that never completes (a protocol bug) traps after ~2 s instead
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// kind::f16, fp16 A / B (K-major), fp32 accumulate, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// sc * x[0..7] -> 8 fp16 hi values and 8 fp16 lo values, packed as 4 words each. `sc` is a power of two that lifts the
// lo parts of O(1) activations out of fp16's subnormal range (|lo| <= 2^-12 |x| is subnormal for |x| < 0.25 unscaled).
__device__ __forceinline__ void split8(const float* x, float sc, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float x0 = x[2 * p] * sc, x1 = x[2 * p + 1] * sc;
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
    const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    l[p] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ============================================================================================================
// weight preparation
// ============================================================================================================
__global__ void absmax_kernel(const float* __restrict__ p, int rows, int cols, int ld, unsigned* out) {
  float m = 0.0f;
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(p[(i / cols) * ld + (i % cols)]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}
// scale = 2^(13 - floor(log2 max)) so that max * scale lies in [2^13, 2^14); inv = 1 / scale (both exact)
__global__ void scales_kernel(const unsigned* maxbits, int n, float* scale, float* inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = __uint_as_float(maxbits[i]);
  int e = 0;
  if (m > 0.0f && isfinite(m)) e = ilogbf(m);
  e = max(-100, min(100, e));
  scale[i] = ldexpf(1.0f, 13 - e);
  inv[i] = ldexpf(1.0f, e - 13);
}
// Wf[n][k] = sum_d Wy[n][d] Wo[d][k],  bf[n] = sum_d Wy[n][d] bo[d]   (double accumulation)
__global__ void fold_out_kernel(const float* __restrict__ Wy, int ldy, const float* __restrict__ Wo, int ldo,
                                const float* __restrict__ bo, int N, int D, int K, float* __restrict__ Wf, float* __restrict__ bf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (K + 1)) return;
  const int n = i / (K + 1), k = i % (K + 1);
  double acc = 0.0;
  if (k < K) {
    for (int d = 0; d < D; ++d) acc += (double)Wy[(int64_t)n * ldy + d] * (double)Wo[(int64_t)d * ldo + k];
    Wf[(int64_t)n * K + k] = (float)acc;
  } else {
    for (int d = 0; d < D; ++d) acc += (double)Wy[(int64_t)n * ldy + d] * (double)bo[d];
    bf[n] = (float)acc;
  }
}
__device__ __forceinline__ void put_split(__half* dst, int64_t hi_idx, int64_t lo_idx, float v) {
  const __half h = __float2half_rn(v);
  dst[hi_idx] = h;
  dst[lo_idx] = __float2half_rn(v - __half2float(h));
}

namespace dtc {
constexpr int S = 16;                         // column slices = CTAs per cluster
constexpr int TBR = 128;                      // UMMA M (TMEM lanes)
constexpr int ROWS = 64;                      // REAL rows per cluster: the upper 64 rows of every A tile are whatever lies behind
                                              // the stage in shared memory (their accumulator lanes are never read). A phase is
                                              // bound by the activation bytes a CTA can keep in flight (48 KB of ring against
                                              // ~0.8 us of L2 latency), so half the rows per cluster = half the bytes per phase on
                                              // twice the SMs (4 clusters = 64 CTAs at 256 envs)
constexpr int H = 256, HM = 512, D = 201;     // the only dimensions this kernel is built for
constexpr int NA = 80, NB = 16, NC = 64;      // per-CTA output columns: A = 32 t1 + 3 x 16 gh, B = 16 t2, C = 16 y + 3 x 16 gi
constexpr int YS = 13;                        // y columns owned per slice (16 x 13 = 208 >= 201)
constexpr int CH = ROWS * 128;                // one activation chunk: 64 rows x 64 fp16
constexpr int WA_CH = NA * 128, WB_CH = NB * 128, WC_CH = NC * 128;
constexpr int RING = 6;
constexpr int OFF_RING = 0;                             // the ring comes FIRST: the 128-row A descriptor of the last stage reads
                                                        // 8 KB past it, which must still be this CTA's shared memory (the weights)
constexpr int OFF_WA = OFF_RING + RING * CH;            // [4 k-chunks][hi | lo]
constexpr int OFF_WB = OFF_WA + 4 * 2 * WA_CH;          // [8][hi | lo]
constexpr int OFF_WC = OFF_WB + 8 * 2 * WB_CH;          // [4][hi | lo]
constexpr int OFF_BARS = OFF_WC + 4 * 2 * WC_CH;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
constexpr int W_BYTES = OFF_BARS - OFF_WA;
constexpr int THREADS = 320;                  // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
// TMEM: every product has TWO accumulators - the hi*hi pass and the two correction passes (hi*lo + lo*hi, 2^-11 smaller).
// The tensor core truncates when it folds an MMA into its accumulator; kept apart, the large accumulator sees K/16
// truncations instead of 3K/16 and the corrections lose nothing that matters (measured: one shared accumulator that also
// carried y over the 18 steps drifted by 6e-5).
constexpr int TM_A = 0, TM_B = 80, TM_C = 96, TM_CORR = 256, TMEM_COLS = 512;
constexpr int N_PHASES = 1 + 3 * 18;
constexpr float ACT_SCALE = 1024.0f, ACT_INV = 1.0f / 1024.0f;   // h, t1, t2 are gate / tanh outputs in [-1, 1]
static_assert(OFF_WA % 1024 == 0 && OFF_WB % 1024 == 0 && OFF_WC % 1024 == 0 && OFF_RING % 1024 == 0 && WA_CH % 1024 == 0 && WB_CH % 1024 == 0 &&
              CH % 1024 == 0, "swizzle atoms");
static_assert(OFF_RING + RING * CH + (TBR - ROWS) * 128 <= OFF_BARS, "the A descriptor's upper rows stay inside the CTA's shared memory");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
enum { SC_W1 = 0, SC_WHH, SC_W2, SC_WO, SC_WF, N_SC };
}  // namespace dtc

struct DecPack {
  const float *W1, *Whh, *W2, *Wo, *Wf;       // [512][256] [768][256] [256][512] [201][256] [768][256]
  const float* scale;                         // [N_SC]
  __half *WA, *WB, *WC;                       // [2 x 1280][256], [2 x 256][512], [2 x 1024][256]  (hi rows, then lo rows)
};
__global__ void dec_pack_kernel(const DecPack p) {
  using namespace dtc;
  constexpr int nA = S * NA * H, nB = S * NB * HM, nC = S * NC * H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nA) {
    const int row = i / H, k = i % H, j = row / NA, r = row % NA;
    float v;
    if (r < 32) v = p.W1[(int64_t)(32 * j + r) * H + k] * p.scale[SC_W1];
    else { const int g = (r - 32) >> 4, jj = (r - 32) & 15; v = p.Whh[(int64_t)(g * H + 16 * j + jj) * H + k] * p.scale[SC_WHH]; }
    put_split(p.WA, i, (int64_t)nA + i, v);
  } else if (i < nA + nB) {
    const int e = i - nA;
    put_split(p.WB, e, (int64_t)nB + e, p.W2[e] * p.scale[SC_W2]);          // slice j = rows 16 j .. 16 j + 15 of W2
  } else if (i < nA + nB + nC) {
    const int e = i - nA - nB, row = e / H, k = e % H, j = row / NC, r = row % NC;
    float v = 0.0f;
    if (r < 16) { const int d = YS * j + r; if (r < YS && d < D) v = p.Wo[(int64_t)d * H + k] * p.scale[SC_WO]; }
    else { const int g = (r - 16) >> 4, jj = (r - 16) & 15; v = p.Wf[(int64_t)(g * H + 16 * j + jj) * H + k] * p.scale[SC_WF]; }
    put_split(p.WC, e, (int64_t)nC + e, v);
  }
}

// ============================================================================================================
// decode kernel
// ============================================================================================================
struct DtcArgs {
  const float* h0;        // [B][256]
  const float* gi1;       // [B][768]
  float* Y;               // [B][20][201]
  __half *Hact, *T1act, *T2act;     // [2 x rows_pad][256 | 512 | 256] exchange buffers (hi rows, then lo rows)
  int rows_pad;
  const float *b1, *bhh, *b2, *bo, *bf;
  const float* scale;     // [N_SC]
  const float* inv;       // [N_SC]
  int B;
};

__global__ void __launch_bounds__(dtc::THREADS, 1)
decode_tc_kernel(const __grid_constant__ CUtensorMap mapWA, const __grid_constant__ CUtensorMap mapWB,
                 const __grid_constant__ CUtensorMap mapWC, const __grid_constant__ CUtensorMap mapH,
                 const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ CUtensorMap mapT2, const DtcArgs a) {
  using namespace dtc;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;                 // [RING] activation chunk landed
  uint64_t* empty_bar = bars + RING;         // [RING] MMAs that read the chunk retired
  uint64_t* w_bar = bars + 2 * RING;         // resident weight slices landed
  uint64_t* acc_bar = bars + 2 * RING + 1;   // accumulators of the phase complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * RING + 2);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x;                  // column slice = rank in the cluster
  const int rt = blockIdx.y;                 // row tile
  const int row0 = rt * ROWS;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(w_bar, 1); mbar_init(acc_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_async_smem();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== producer: resident weight slices once, then the activation chunks of every phase =====
    if (lane == 0) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int c = 0; c < 4; ++c) {
        tma_load_2d(smem + OFF_WA + c * 2 * WA_CH, &mapWA, w_bar, c * 64, j * NA);
        tma_load_2d(smem + OFF_WA + c * 2 * WA_CH + WA_CH, &mapWA, w_bar, c * 64, S * NA + j * NA);
      }
      for (int c = 0; c < 8; ++c) {
        tma_load_2d(smem + OFF_WB + c * 2 * WB_CH, &mapWB, w_bar, c * 64, j * NB);
        tma_load_2d(smem + OFF_WB + c * 2 * WB_CH + WB_CH, &mapWB, w_bar, c * 64, S * NB + j * NB);
      }
      for (int c = 0; c < 4; ++c) {
        tma_load_2d(smem + OFF_WC + c * 2 * WC_CH, &mapWC, w_bar, c * 64, j * NC);
        tma_load_2d(smem + OFF_WC + c * 2 * WC_CH + WC_CH, &mapWC, w_bar, c * 64, S * NC + j * NC);
      }
    }
    __syncwarp();
    cluster_arrive(); cluster_wait();                  // barrier 0: h_0 published by every CTA of the cluster
    uint32_t it = 0;
    for (int ph = 0; ph < N_PHASES; ++ph) {
      const int kind = ph == 0 ? 0 : (ph - 1) % 3;
      fence_async_all();                               // peers' generic global writes (acquired above) -> async-proxy reads
      if (lane == 0) {
        const CUtensorMap* map = kind == 0 ? &mapH : kind == 1 ? &mapT1 : &mapT2;
        const int nch = kind == 1 ? 8 : 4;
        for (int c = 0; c < nch; ++c)
          for (int hl = 0; hl < 2; ++hl, ++it) {
            const uint32_t s = it % RING, par = (it / RING) & 1u;
            mbar_wait(&empty_bar[s], par ^ 1u);
            mbar_expect_tx(&full_bar[s], CH);
            tma_load_2d(smem + OFF_RING + s * CH, map, &full_bar[s], c * 64, hl * a.rows_pad + row0);
          }
      }
      __syncwarp();
      cluster_arrive(); cluster_wait();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    cluster_arrive(); cluster_wait();                  // barrier 0
    tc_fence_after();
    mbar_wait(w_bar, 0);
    uint32_t it = 0;
    for (int ph = 0; ph < N_PHASES; ++ph) {
      const int kind = ph == 0 ? 0 : (ph - 1) % 3;
      const int nch = kind == 1 ? 8 : 4;
      const uint32_t n = kind == 0 ? NA : kind == 1 ? NB : NC;
      const uint32_t idesc = idesc_f16((int)n);
      const uint32_t tm = tmem_base + (kind == 0 ? TM_A : kind == 1 ? TM_B : TM_C);
      const uint32_t wbase = sbase + (kind == 0 ? OFF_WA : kind == 1 ? OFF_WB : OFF_WC), wch = n * 128u;
      for (int c = 0; c < nch; ++c) {
        const uint64_t w_hi = make_desc(wbase + c * 2 * wch), w_lo = make_desc(wbase + c * 2 * wch + wch);
        {  // hi chunk: A_hi W_hi + A_hi W_lo
          const uint32_t s = it % RING, par = (it / RING) & 1u;
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = make_desc(sbase + OFF_RING + s * CH);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma(tm, da + (uint64_t)(kk * 2), w_hi + (uint64_t)(kk * 2), idesc, (c | kk) ? 1u : 0u);
              umma(tm + TM_CORR, da + (uint64_t)(kk * 2), w_lo + (uint64_t)(kk * 2), idesc, (c | kk) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          ++it;
        }
        {  // lo chunk: A_lo W_hi
          const uint32_t s = it % RING, par = (it / RING) & 1u;
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = make_desc(sbase + OFF_RING + s * CH);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma(tm + TM_CORR, da + (uint64_t)(kk * 2), w_hi + (uint64_t)(kk * 2), idesc, 1u);
            umma_commit(&empty_bar[s]);
            if (c == nch - 1) umma_commit(acc_bar);
          }
          __syncwarp();
          ++it;
        }
      }
      tc_fence_before();
      cluster_arrive(); cluster_wait();
      tc_fence_after();
    }
  } else {
    // ===== epilogue warps: thread = row (TMEM lane q*32 + lane), `half` = which half of the phase's columns =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    if (q >= ROWS / 32) {
      // TMEM lanes 64..127 carry no rows: these warps only keep the cluster barriers of every phase company
      cluster_arrive(); cluster_wait();                // barrier 0
      for (int ph = 0; ph < N_PHASES; ++ph) { cluster_arrive(); cluster_wait(); }
    } else {
    const int row = q * 32 + lane;
    const int rg = row0 + row;                          // row of the exchange buffers (always allocated)
    const int b = min(rg, a.B - 1);                     // env this row mirrors (rows >= B duplicate the last env)
    const bool b_ok = rg < a.B;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float inv_w1 = __ldg(a.inv + SC_W1) * ACT_INV, inv_hh = __ldg(a.inv + SC_WHH) * ACT_INV, inv_w2 = __ldg(a.inv + SC_W2) * ACT_INV,
                inv_wo = __ldg(a.inv + SC_WO) * ACT_INV, inv_wf = __ldg(a.inv + SC_WF) * ACT_INV;
    const int hc0 = 16 * j + 8 * half;                  // my 8 hidden columns
    float hprev[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hprev[i] = a.h0[(int64_t)b * H + hc0 + i];
    // running fp32 sums of the two quantities that accumulate over the steps: my y columns and my gi columns
    float yrun[8], girun[3][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 8 * half + i, d = YS * j + r;
      yrun[i] = (r < YS && d < D) ? a.Y[((int64_t)b * 20 + 1) * D + d] : 0.0f;
#pragma unroll
      for (int g = 0; g < 3; ++g) girun[g][i] = a.gi1[(int64_t)b * 3 * H + g * H + hc0 + i];
    }
    {  // publish my slice of h_0
      uint4 hi, lo;
      split8(hprev, ACT_SCALE, hi, lo);
      *reinterpret_cast<uint4*>(a.Hact + (int64_t)rg * H + hc0) = hi;
      *reinterpret_cast<uint4*>(a.Hact + (int64_t)(a.rows_pad + rg) * H + hc0) = lo;
    }
    fence_async_all();
    tc_fence_before();
    cluster_arrive(); cluster_wait();                  // barrier 0

    // sum of the two accumulators of 8 columns
    auto ld8 = [&](uint32_t col, float* v) {
      float c2[8];
      tmem_ld8(trow + col, v);
      tmem_ld8(trow + TM_CORR + col, c2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += c2[i];
    };
    // GRU gates of the next step (PyTorch order r, z, n): gi = running input term, gh = region A (+ b_hh)
    auto gates = [&]() {
      float gh[3][8];
#pragma unroll
      for (int g = 0; g < 3; ++g) ld8(TM_A + 32 + 16 * g + 8 * half, gh[g]);
      float hn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = hc0 + i;
        const float ghr = fmaf(gh[0][i], inv_hh, __ldg(a.bhh + c)), ghz = fmaf(gh[1][i], inv_hh, __ldg(a.bhh + H + c)),
                    ghn = fmaf(gh[2][i], inv_hh, __ldg(a.bhh + 2 * H + c));
        const float rr = sigmoid_(girun[0][i] + ghr), zz = sigmoid_(girun[1][i] + ghz);
        const float nn = tanhf(girun[2][i] + rr * ghn);
        hn[i] = (1.0f - zz) * nn + zz * hprev[i];
        hprev[i] = hn[i];
      }
      uint4 hi, lo;
      split8(hn, ACT_SCALE, hi, lo);
      *reinterpret_cast<uint4*>(a.Hact + (int64_t)rg * H + hc0) = hi;
      *reinterpret_cast<uint4*>(a.Hact + (int64_t)(a.rows_pad + rg) * H + hc0) = lo;
    };

    for (int ph = 0; ph < N_PHASES; ++ph) {
      const int kind = ph == 0 ? 0 : (ph - 1) % 3;
      const int t = ph == 0 ? 0 : (ph - 1) / 3 + 1;
      mbar_wait(acc_bar, (uint32_t)ph & 1u);
      tc_fence_after();
      if (kind == 0) {
        if (t == 0) {
          gates();
        } else {                                          // t1 = tanh(h W1^T + b1): my 16 of the slice's 32 columns
          float v[16];
          ld8(TM_A + 16 * half, v);
          ld8(TM_A + 16 * half + 8, v + 8);
          const int c0 = 32 * j + 16 * half;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = tanhf(fmaf(v[i], inv_w1, __ldg(a.b1 + c0 + i)));
          uint4 hi, lo;
          split8(v, ACT_SCALE, hi, lo);
          *reinterpret_cast<uint4*>(a.T1act + (int64_t)rg * HM + c0) = hi;
          *reinterpret_cast<uint4*>(a.T1act + (int64_t)(a.rows_pad + rg) * HM + c0) = lo;
          split8(v + 8, ACT_SCALE, hi, lo);
          *reinterpret_cast<uint4*>(a.T1act + (int64_t)rg * HM + c0 + 8) = hi;
          *reinterpret_cast<uint4*>(a.T1act + (int64_t)(a.rows_pad + rg) * HM + c0 + 8) = lo;
        }
      } else if (kind == 1) {                             // t2 = tanh(t1 W2^T + b2): my 8 of the slice's 16 columns
        float v[8];
        ld8(TM_B + 8 * half, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = tanhf(fmaf(v[i], inv_w2, __ldg(a.b2 + hc0 + i)));
        uint4 hi, lo;
        split8(v, ACT_SCALE, hi, lo);
        *reinterpret_cast<uint4*>(a.T2act + (int64_t)rg * H + hc0) = hi;
        *reinterpret_cast<uint4*>(a.T2act + (int64_t)(a.rows_pad + rg) * H + hc0) = lo;
      } else {                                            // y_t = (t2 Wo^T + bo) + y_{t-1}; gi_{t+1} = gi_t + (t2 Wf^T + bf)
        float v[8];
        ld8(TM_C + 8 * half, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 8 * half + i, d = YS * j + r;
          if (r < YS && d < D) {
            yrun[i] = fmaf(v[i], inv_wo, __ldg(a.bo + d)) + yrun[i];
            if (b_ok) a.Y[((int64_t)b * 20 + 1 + t) * D + d] = yrun[i];
          }
        }
        if (t < 18) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            ld8(TM_C + 16 + 16 * g + 8 * half, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) girun[g][i] += fmaf(v[i], inv_wf, __ldg(a.bf + g * H + hc0 + i));
          }
          gates();
        }
      }
      fence_async_all();                                  // my global writes -> peers' TMA reads after the barrier
      tc_fence_before();
      cluster_arrive(); cluster_wait();
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(dtc::TMEM_COLS) : "memory");
  }
}

// ============================================================================================================
// regressor
// ============================================================================================================
namespace rtc {
constexpr int TBR = 128, HR = 128, D = 201, BD = 159, BDP = 160, NBETA = 10;
constexpr int FPE = 18;                         // regressed frames per env (frames 2..19 of the 20-frame primitive)
constexpr float ACT_SCALE = 256.0f, ACT_INV = 1.0f / 256.0f;   // activations (metres, ReLU features) stay far below 65504 / 256
constexpr int THREADS = 320;
constexpr int A_HALF = TBR * 128;               // one [128 x 64] fp16 chunk
constexpr int A_CHUNK = 2 * A_HALF;             // hi + lo
constexpr int NA_CH = 3;                        // activation tile: up to 192 k
constexpr int ST_BYTES = 32768, RING = 4;       // weight stage: [n rows hi | n rows lo] x 64 k, n = 128 or 32
constexpr int OFF_A = 0, OFF_RING = NA_CH * A_CHUNK, OFF_BARS = OFF_RING + RING * ST_BYTES;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
// TMEM: the tensor core truncates every time it folds an MMA into its accumulator, and 8 (K = 128) truncations at full
// magnitude per layer, 63 layers deep, showed up as 3e-4 in the axis-angle outputs. So a layer's k-steps are dealt to FOUR
// accumulators (column bases 0 / 128 / 256 / 384; three of 160 columns for out_fc): each takes the small correction passes
// (hi*lo, lo*hi) of its two k-steps first - truncated at a magnitude 2^-11 below the result - and then only two hi*hi
// MMAs; the epilogue adds the accumulators in fp32 round-to-nearest. CPU emulation of this scheme (tools/emulate_tc_regressor.py)
// lands on the error level of an fp32 SGEMM chain.
constexpr int TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
// stage program word y: flags, activation chunk, accumulators of k-steps {0,1} and {2,3}, accumulator stride, column offset
enum { OP_SMALL = 1, OP_LAST = 2, OP_WAIT_A = 4, OP_FIRST0 = 8, OP_FIRST1 = 16, OP_WIDE = 1 << 11 };
__host__ __device__ constexpr uint32_t op_word(uint32_t flags, int a_chunk, int acc0, int acc1, int col_off) {
  return flags | ((uint32_t)a_chunk << 5) | ((uint32_t)acc0 << 7) | ((uint32_t)acc1 << 9) | ((uint32_t)col_off << 12);
}
}  // namespace rtc

// one 64-wide k-chunk of one weight matrix -> prepared rows [dst_row0, +n_rows) hi and [dst_row0 + n_rows, +n_rows) lo
struct RegPackStage {
  const float* src; int ld;       // source matrix [out][in]
  int src_row0, n_valid_rows;     // rows src_row0 .. (zero rows beyond n_valid_rows)
  int k0;                         // first k of the chunk
  int mode;                       // 0: column = k (valid < kvalid); 1: base columns (markers | betas); 2: xb columns
  int kvalid;
  int dst_row0, n_rows;
  int scale_idx;
};
__global__ void reg_pack_kernel(const RegPackStage* __restrict__ stages, const float* __restrict__ scale, __half* __restrict__ dst) {
  using namespace rtc;
  const RegPackStage s = stages[blockIdx.x];
  const float sc = scale[s.scale_idx];
  for (int i = threadIdx.x; i < s.n_rows * 64; i += blockDim.x) {
    const int r = i >> 6, kk = i & 63, k = s.k0 + kk;
    int col = -1;
    if (s.mode == 0) col = k < s.kvalid ? k : -1;
    else if (s.mode == 1) col = k < D ? k : (k < D + NBETA ? D + BD + (k - D) : -1);
    else col = k < BD ? D + k : -1;
    float v = 0.0f;
    if (col >= 0 && r < s.n_valid_rows) v = s.src[(int64_t)(s.src_row0 + r) * s.ld + col] * sc;
    put_split(dst, (int64_t)(s.dst_row0 + r) * 64 + kk, (int64_t)(s.dst_row0 + s.n_rows + r) * 64 + kk, v);
  }
}

struct RtcArgs {
  const float* Y;          // [B][20][201]
  const float* betas;      // [B][10]
  float* xbc;              // [B*20][159]
  float* base_buf;         // [n_tiles*128][128] in_fc output without the xb term (constant over the recurrences)
  float* xb_buf;           // [n_tiles*128][160] running xb
  const float* bias;       // [b_in 128 | 2 nb x 128 | b_out 160 (padded with 0)]
  const float* inv;        // [1 + 2 nb + 1] inverse weight scales: in_fc, block layers, out_fc
  const uint2* prog;       // stage program
  int n_stages, M, nb, nrec;
};

__global__ void __launch_bounds__(rtc::THREADS, 1)
regressor_tc_kernel(const __grid_constant__ CUtensorMap map256, const __grid_constant__ CUtensorMap map64, const RtcArgs a) {
  using namespace rtc;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;                 // [RING]
  uint64_t* empty_bar = bars + RING;         // [RING]
  uint64_t* acc_bar = bars + 2 * RING;       // MMAs of the layer group retired
  uint64_t* a_bar = bars + 2 * RING + 1;     // activation tile rewritten for the next group (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * RING + 2);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(acc_bar, 1); mbar_init(a_bar, 8);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_async_smem();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== weight producer: the whole ordered stream of weight stages, independent of the activations =====
    if (lane == 0) {
      for (int s = 0; s < a.n_stages; ++s) {
        const uint2 op = __ldg(a.prog + s);
        const uint32_t st = (uint32_t)s % RING, par = ((uint32_t)s / RING) & 1u;
        mbar_wait(&empty_bar[st], par ^ 1u);
        if (op.y & OP_SMALL) {
          mbar_expect_tx(&full_bar[st], 64 * 128);
          tma_load_2d(smem + OFF_RING + st * ST_BYTES, &map64, &full_bar[st], 0, (int)op.x);
        } else {
          mbar_expect_tx(&full_bar[st], 256 * 128);
          tma_load_2d(smem + OFF_RING + st * ST_BYTES, &map256, &full_bar[st], 0, (int)op.x);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    uint32_t g = 0;
    for (int s = 0; s < a.n_stages; ++s) {
      const uint2 op = __ldg(a.prog + s);
      if (op.y & OP_WAIT_A) { mbar_wait(a_bar, g & 1u); tc_fence_after(); }
      const uint32_t st = (uint32_t)s % RING, par = ((uint32_t)s / RING) & 1u;
      mbar_wait(&full_bar[st], par);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t n = (op.y & OP_SMALL) ? 32u : 128u;
        const uint32_t idesc = idesc_f16((int)n);
        const uint32_t a_chunk = (op.y >> 5) & 3u, stride = (op.y & OP_WIDE) ? 160u : 128u, col_off = (op.y >> 12) & 0x1ffu;
        const uint64_t a_hi = make_desc(sbase + OFF_A + a_chunk * A_CHUNK), a_lo = make_desc(sbase + OFF_A + a_chunk * A_CHUNK + A_HALF);
        const uint64_t b_hi = make_desc(sbase + OFF_RING + st * ST_BYTES), b_lo = make_desc(sbase + OFF_RING + st * ST_BYTES + n * 128u);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint32_t acc = (op.y >> (7 + 2 * hf)) & 3u;
          const uint32_t tm = tmem_base + acc * stride + col_off;
          uint32_t accum = (op.y & (hf ? OP_FIRST1 : OP_FIRST0)) ? 0u : 1u;
#pragma unroll
          for (int kk = 2 * hf; kk < 2 * hf + 2; ++kk) {          // corrections first: truncated while the accumulator is small
            umma(tm, a_hi + (uint64_t)(kk * 2), b_lo + (uint64_t)(kk * 2), idesc, accum);
            umma(tm, a_lo + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc, 1u);
            accum = 1u;
          }
#pragma unroll
          for (int kk = 2 * hf; kk < 2 * hf + 2; ++kk) umma(tm, a_hi + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc, 1u);
        }
        umma_commit(&empty_bar[st]);
        if (op.y & OP_LAST) umma_commit(acc_bar);
      }
      __syncwarp();
      if (op.y & OP_LAST) ++g;
    }
  } else {
    // ===== epilogue warps: thread = row (TMEM lane), `half` = columns [64 half, +64) of a 128-wide layer =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int rg = blockIdx.x * TBR + row;
    const bool r_ok = rg < a.M;
    const int rr = min(rg, a.M - 1), env = rr / FPE, frame = 2 + rr % FPE;
    const float* yrow = a.Y + ((int64_t)env * 20 + frame) * D;
    const float* brow = a.betas + (int64_t)env * NBETA;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sbase + OFF_A + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    // 8 consecutive k (k8 = k / 8 inside the 192-wide tile) of my row -> swizzled hi / lo slots of the A operand
    auto store_a8 = [&](int k8, const float* x) {
      uint4 hi, lo;
      split8(x, ACT_SCALE, hi, lo);
      const uint32_t addr = a_row + (uint32_t)(k8 >> 3) * A_CHUNK + ((((uint32_t)k8 & 7u) ^ sw) << 4);
      sts_u4(addr, hi);
      sts_u4(addr + A_HALF, lo);
    };
    auto publish_a = [&]() {                      // activation tile complete -> MMA warp
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_bar);
    };
    auto src_in = [&](int k) -> float {           // regressor input column k of [markers | betas | 0]
      return k < D ? __ldg(yrow + k) : (k < D + NBETA ? __ldg(brow + (k - D)) : 0.0f);
    };
    uint32_t g = 0;
    auto wait_acc = [&]() { mbar_wait(acc_bar, g & 1u); ++g; tc_fence_after(); };
    auto ld16 = [&](uint32_t col, float* v) {     // sum of the four 128-wide accumulators of 16 columns
      float c1[16], c2[16], c3[16];
      tmem_ld16(trow + col, v);
      tmem_ld16(trow + 128 + col, c1);
      tmem_ld16(trow + 256 + col, c2);
      tmem_ld16(trow + 384 + col, c3);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (v[i] + c1[i]) + (c2[i] + c3[i]);
    };
    auto ld16w = [&](uint32_t col, float* v) {    // out_fc: three 160-wide accumulators
      float c1[16], c2[16];
      tmem_ld16(trow + col, v);
      tmem_ld16(trow + 160 + col, c1);
      tmem_ld16(trow + 320 + col, c2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (v[i] + c1[i]) + c2[i];
    };

    // group 0: input columns 0..191 (half 0: 0..95, half 1: 96..191)
    for (int k8 = 12 * half; k8 < 12 * half + 12; ++k8) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = src_in(8 * k8 + i);
      store_a8(k8, x);
    }
    publish_a();
    // group 1: input columns 192..255 into chunk 0 (half 0: 192..223, half 1: 224..255)
    wait_acc();
    for (int k8 = 4 * half; k8 < 4 * half + 4; ++k8) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = src_in(192 + 8 * k8 + i);
      store_a8(k8, x);
    }
    publish_a();

    float h[64];                                   // residual stream: my 64 columns of my row
    const int cb = 64 * half;
    float* base_row = a.base_buf + (int64_t)rg * HR + cb;
    // base = [markers | betas] W^T + b_in ; h = base
    wait_acc();
    {
      const float inv = __ldg(a.inv) * ACT_INV;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16];
        ld16(cb + 16 * c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = fmaf(v[i], inv, __ldg(a.bias + cb + 16 * c + i)); h[16 * c + i] = v[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(base_row)[4 * c + i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        store_a8((cb >> 3) + 2 * c, v);
        store_a8((cb >> 3) + 2 * c + 1, v + 8);
      }
    }
    publish_a();

    for (int rec = 0; rec < a.nrec; ++rec) {
      if (rec > 0) {                                // h = base + xb Wb^T
        wait_acc();
        const float inv = __ldg(a.inv) * ACT_INV;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[16];
          ld16(cb + 16 * c, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bv = reinterpret_cast<const float4*>(base_row)[4 * c + i];
            v[4 * i] = fmaf(v[4 * i], inv, bv.x); v[4 * i + 1] = fmaf(v[4 * i + 1], inv, bv.y);
            v[4 * i + 2] = fmaf(v[4 * i + 2], inv, bv.z); v[4 * i + 3] = fmaf(v[4 * i + 3], inv, bv.w);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) h[16 * c + i] = v[i];
          store_a8((cb >> 3) + 2 * c, v);
          store_a8((cb >> 3) + 2 * c + 1, v + 8);
        }
        publish_a();
      }
      for (int l = 0; l < 2 * a.nb; ++l) {          // residual blocks: t = relu(W1 h + b1); h = relu(W2 t + b2) + h
        wait_acc();
        const float inv = __ldg(a.inv + 1 + l) * ACT_INV;
        const float* bias = a.bias + HR + l * HR + cb;
        const bool second = (l & 1) != 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[16];
          ld16(cb + 16 * c, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = fmaxf(fmaf(v[i], inv, __ldg(bias + 16 * c + i)), 0.0f);
            if (second) { x += h[16 * c + i]; h[16 * c + i] = x; }
            v[i] = x;
          }
          store_a8((cb >> 3) + 2 * c, v);
          store_a8((cb >> 3) + 2 * c + 1, v + 8);
        }
        publish_a();
      }
      // xb = (h Wout^T + b_out) + xb: my columns [80 half, +80); the running xb stays in L2 (one private row per thread)
      wait_acc();
      {
        const float inv = __ldg(a.inv + 1 + 2 * a.nb) * ACT_INV;
        const float* bo = a.bias + HR + 2 * a.nb * HR;
        const bool last = rec == a.nrec - 1;
        float* xo = a.xbc + ((int64_t)env * 20 + frame) * BD;
        float* xrow = a.xb_buf + (int64_t)rg * BDP;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          float v[16];
          const int c0 = 80 * half + 16 * c;
          ld16w(c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], inv, __ldg(bo + c0 + i));
          if (rec > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 pv = reinterpret_cast<const float4*>(xrow + c0)[i];
              v[4 * i] += pv.x; v[4 * i + 1] += pv.y; v[4 * i + 2] += pv.z; v[4 * i + 3] += pv.w;
            }
          }
          if (last) {
            if (r_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (c0 + i < BD) xo[c0 + i] = v[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(xrow + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            store_a8(c0 >> 3, v);
            store_a8((c0 >> 3) + 1, v + 8);
          }
        }
        if (!last) {
          if (half == 1) {                          // zero columns 160..191 of the xb operand
            float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k8 = 20; k8 < 24; ++k8) store_a8(k8, z);
          }
          publish_a();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(rtc::TMEM_COLS) : "memory");
  }
}

// ============================================================================================================
// host side
// ============================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MotionTc {
  EgMotionDims d{};
  EncodeTiledFn encode = nullptr;
  bool dec_ok = false, reg_ok = false;
  // decode
  float *Wf = nullptr, *bf = nullptr;
  unsigned* maxbits = nullptr;
  float *scale = nullptr, *inv = nullptr;             // decode scales [N_SC] then regressor scales [2 nb + 2]
  __half *WA = nullptr, *WB = nullptr, *WC = nullptr;
  CUtensorMap mapWA{}, mapWB{}, mapWC{}, mapH{}, mapT1{}, mapT2{};
  __half *Hact = nullptr, *T1act = nullptr, *T2act = nullptr;
  int cap_rows = 0;
  const float *b1 = nullptr, *bhh = nullptr, *b2 = nullptr, *bo = nullptr;
  // regressor
  __half* RW = nullptr;
  int rw_rows = 0;
  RegPackStage* pack_dev = nullptr;
  int n_pack = 0;
  float* rbias = nullptr;
  uint2* prog = nullptr;
  int n_stages = 0;
  CUtensorMap map256{}, map64{};
  float *base_buf = nullptr, *xb_buf = nullptr;
  int cap_tiles = 0;
};

static int encode_f16_2d(const MotionTc* m, CUtensorMap* map, const __half* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  if (!m->encode) return set_error(EG_ERR_STATE, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
  const cuuint32_t box[2] = {64u, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return EG_OK;
}

namespace {
enum {  // EgMotion weight table order (nn.cu)
  P_XENC_WIH = 0, P_XENC_WHH, P_XENC_BIH, P_XENC_BHH,
  P_DRNN_W0, P_DRNN_B0, P_DRNN_W1, P_DRNN_B1, P_DRNN_W2, P_DRNN_B2,
  P_DRNN_WIH, P_DRNN_WHH, P_DRNN_BIH, P_DRNN_BHH,
  P_DMLP_W0, P_DMLP_B0, P_DMLP_W1, P_DMLP_B1,
  P_DOUT_W, P_DOUT_B,
  R_IN_W, R_IN_B,
  R_BLOCKS
};
}  // namespace

static int absmax(const float* p, int rows, int cols, int ld, unsigned* out, cudaStream_t st) {
  const int64_t n = (int64_t)rows * cols;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 296);
  EG_LAUNCH(absmax_kernel, grid, 256, 0, st, p, rows, cols, ld, out);
  return EG_OK;
}

static int prepare_weights(MotionTc* m, const float* const* w, cudaStream_t st) {
  const EgMotionDims& d = m->d;
  const int nb = d.reg_blocks, n_rs = 2 * nb + 2, n_sc = dtc::N_SC + n_rs;
  EG_CUDA_CHECK(cudaMemsetAsync(m->maxbits, 0, n_sc * sizeof(unsigned), st));
  int rc;
  if (m->dec_ok) {
    using namespace dtc;
    const int Kin = H + d.z_dim + D;
    const int n = 3 * H * (H + 1);
    EG_LAUNCH(fold_out_kernel, (n + 255) / 256, 256, 0, st, w[P_DRNN_WIH] + H + d.z_dim, Kin, w[P_DOUT_W], H, w[P_DOUT_B], 3 * H, D, H,
              m->Wf, m->bf);
    if ((rc = absmax(w[P_DMLP_W0], HM, H, H, m->maxbits + SC_W1, st))) return rc;
    if ((rc = absmax(w[P_DRNN_WHH], 3 * H, H, H, m->maxbits + SC_WHH, st))) return rc;
    if ((rc = absmax(w[P_DMLP_W1], H, HM, HM, m->maxbits + SC_W2, st))) return rc;
    if ((rc = absmax(w[P_DOUT_W], D, H, H, m->maxbits + SC_WO, st))) return rc;
    if ((rc = absmax(m->Wf, 3 * H, H, H, m->maxbits + SC_WF, st))) return rc;
  }
  if (m->reg_ok) {
    const float* const* wb = w + R_BLOCKS;
    const int Kr = rtc::D + rtc::BD + rtc::NBETA;
    if ((rc = absmax(w[R_IN_W], rtc::HR, Kr, Kr, m->maxbits + dtc::N_SC, st))) return rc;
    for (int l = 0; l < 2 * nb; ++l)
      if ((rc = absmax(wb[(l / 2) * 4 + (l % 2) * 2], rtc::HR, rtc::HR, rtc::HR, m->maxbits + dtc::N_SC + 1 + l, st))) return rc;
    if ((rc = absmax(wb[nb * 4], rtc::BD, rtc::HR, rtc::HR, m->maxbits + dtc::N_SC + 1 + 2 * nb, st))) return rc;
  }
  EG_LAUNCH(scales_kernel, 1, 64, 0, st, m->maxbits, n_sc, m->scale, m->inv);
  if (m->dec_ok) {
    using namespace dtc;
    DecPack p{w[P_DMLP_W0], w[P_DRNN_WHH], w[P_DMLP_W1], w[P_DOUT_W], m->Wf, m->scale, m->WA, m->WB, m->WC};
    const int n = S * NA * H + S * NB * HM + S * NC * H;
    EG_LAUNCH(dec_pack_kernel, (n + 255) / 256, 256, 0, st, p);
    m->b1 = w[P_DMLP_B0]; m->bhh = w[P_DRNN_BHH]; m->b2 = w[P_DMLP_B1]; m->bo = w[P_DOUT_B];
  }
  if (m->reg_ok) {
    using namespace rtc;
    const float* const* wb = w + R_BLOCKS;
    const int Kr = D + BD + NBETA;
    // host table of pack stages (source pointers change when the caller re-registers weights, so rebuilt every time)
    std::vector<RegPackStage> ps;
    int row = 0;
    for (int c = 0; c < 4; ++c) { ps.push_back({w[R_IN_W], Kr, 0, HR, c * 64, 1, 0, row, 128, 0}); row += 256; }
    for (int c = 0; c < 3; ++c) { ps.push_back({w[R_IN_W], Kr, 0, HR, c * 64, 2, 0, row, 128, 0}); row += 256; }
    for (int l = 0; l < 2 * nb; ++l)
      for (int c = 0; c < 2; ++c) { ps.push_back({wb[(l / 2) * 4 + (l % 2) * 2], HR, 0, HR, c * 64, 0, HR, row, 128, 1 + l}); row += 256; }
    for (int c = 0; c < 2; ++c) {
      ps.push_back({wb[nb * 4], HR, 0, 128, c * 64, 0, HR, row, 128, 1 + 2 * nb}); row += 256;
      ps.push_back({wb[nb * 4], HR, 128, BD - 128, c * 64, 0, HR, row, 32, 1 + 2 * nb}); row += 64;
    }
    if (row != m->rw_rows || (int)ps.size() != m->n_pack) return set_error(EG_ERR_STATE, "regressor pack table mismatch");
    EG_CUDA_CHECK(cudaMemcpyAsync(m->pack_dev, ps.data(), ps.size() * sizeof(RegPackStage), cudaMemcpyHostToDevice, st));
    EG_CUDA_CHECK(cudaStreamSynchronize(st));            // `ps` is pageable host memory
    EG_LAUNCH(reg_pack_kernel, (int)ps.size(), 256, 0, st, m->pack_dev, m->scale + dtc::N_SC, m->RW);
    // biases: b_in | block biases | b_out (160, last entry 0)
    EG_CUDA_CHECK(cudaMemsetAsync(m->rbias, 0, (size_t)(HR + 2 * nb * HR + BDP) * sizeof(float), st));
    EG_CUDA_CHECK(cudaMemcpyAsync(m->rbias, w[R_IN_B], HR * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int l = 0; l < 2 * nb; ++l)
      EG_CUDA_CHECK(cudaMemcpyAsync(m->rbias + HR + l * HR, wb[(l / 2) * 4 + (l % 2) * 2 + 1], HR * sizeof(float), cudaMemcpyDeviceToDevice, st));
    EG_CUDA_CHECK(cudaMemcpyAsync(m->rbias + HR + 2 * nb * HR, wb[nb * 4 + 1], BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return EG_OK;
}

int create(MotionTc** out, const EgMotionDims& d, const float* const* w, cudaStream_t st) {
  MotionTc* m = new MotionTc();
  *out = m;
  m->d = d;
  {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      m->encode = reinterpret_cast<EncodeTiledFn>(fn);
    else
      cudaGetLastError();
  }
  const char* env = getenv("EG_MOTION_TC");
  const bool enabled = !(env != nullptr && env[0] == '0') && m->encode != nullptr;
  const int nb = d.reg_blocks;
  m->dec_ok = enabled && d.h_dim == dtc::H && d.mlp_dim == dtc::HM && d.in_dim == dtc::D;
  m->reg_ok = enabled && d.reg_h == rtc::HR && d.in_dim == rtc::D && d.body_dim == rtc::BD && nb >= 1 && d.reg_recur >= 1;
  if (m->dec_ok) {
    // the decode kernel needs one 16-CTA cluster (non-portable size) to be schedulable with its shared-memory footprint
    if (cudaFuncSetAttribute(decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dtc::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(decode_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      m->dec_ok = false;
    } else {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(dtc::S, 1, 1); q.blockDim = dim3(dtc::THREADS); q.dynamicSmemBytes = dtc::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = dtc::S; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, decode_tc_kernel, &q) != cudaSuccess) { cudaGetLastError(); nc = 0; }
      if (getenv("EG_MOTION_TC_DEBUG")) fprintf(stderr, "[motion_tc] 16-CTA clusters co-resident: %d\n", nc);
      if (nc < 1) m->dec_ok = false;
    }
  }
  if (m->reg_ok) {
    if (cudaFuncSetAttribute(regressor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rtc::SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      m->reg_ok = false;
    }
  }
  if (!m->dec_ok && !m->reg_ok) return EG_OK;
  const int n_sc = dtc::N_SC + 2 * nb + 2;
  EG_CUDA_CHECK(cudaMalloc((void**)&m->maxbits, n_sc * sizeof(unsigned)));
  EG_CUDA_CHECK(cudaMalloc((void**)&m->scale, n_sc * sizeof(float)));
  EG_CUDA_CHECK(cudaMalloc((void**)&m->inv, n_sc * sizeof(float)));
  int rc;
  if (m->dec_ok) {
    using namespace dtc;
    EG_CUDA_CHECK(cudaMalloc((void**)&m->Wf, (size_t)3 * H * H * sizeof(float)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->bf, (size_t)3 * H * sizeof(float)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->WA, (size_t)2 * S * NA * H * sizeof(__half)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->WB, (size_t)2 * S * NB * HM * sizeof(__half)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->WC, (size_t)2 * S * NC * H * sizeof(__half)));
    if ((rc = encode_f16_2d(m, &m->mapWA, m->WA, 2 * S * NA, H, NA))) return rc;
    if ((rc = encode_f16_2d(m, &m->mapWB, m->WB, 2 * S * NB, HM, NB))) return rc;
    if ((rc = encode_f16_2d(m, &m->mapWC, m->WC, 2 * S * NC, H, NC))) return rc;
  }
  if (m->reg_ok) {
    using namespace rtc;
    m->rw_rows = (4 + 3 + 2 * nb * 2) * 256 + 2 * 320;
    m->n_pack = 4 + 3 + 2 * nb * 2 + 4;
    EG_CUDA_CHECK(cudaMalloc((void**)&m->RW, (size_t)m->rw_rows * 64 * sizeof(__half)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->pack_dev, (size_t)m->n_pack * sizeof(RegPackStage)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->rbias, (size_t)(HR + 2 * nb * HR + BDP) * sizeof(float)));
    if ((rc = encode_f16_2d(m, &m->map256, m->RW, m->rw_rows, 64, 256))) return rc;
    if ((rc = encode_f16_2d(m, &m->map64, m->RW, m->rw_rows, 64, 64))) return rc;
    // stage program: x = first prepared row of the stage, y = op_word(...)
    std::vector<uint2> prog;
    auto push = [&](int row0, uint32_t flags, int a_chunk, int acc0, int acc1, int col_off) {
      prog.push_back(make_uint2((uint32_t)row0, op_word(flags, a_chunk, acc0, acc1, col_off)));
    };
    const int R_XB = 4 * 256, R_BLK = R_XB + 3 * 256, R_OUT = R_BLK + 2 * nb * 2 * 256;
    const uint32_t F01 = OP_FIRST0 | OP_FIRST1;
    // base, part 1 (k 0..191: accumulators 0,1 | 2,3 | 0,1) and part 2 (k 192..255 on accumulators 2,3; activation chunk 0)
    push(0 * 256, F01 | OP_WAIT_A, 0, 0, 1, 0);
    push(1 * 256, F01, 1, 2, 3, 0);
    push(2 * 256, OP_LAST, 2, 0, 1, 0);
    push(3 * 256, OP_WAIT_A | OP_LAST, 0, 2, 3, 0);
    for (int rec = 0; rec < d.reg_recur; ++rec) {
      if (rec > 0) {
        push(R_XB + 0 * 256, F01 | OP_WAIT_A, 0, 0, 1, 0);
        push(R_XB + 1 * 256, F01, 1, 2, 3, 0);
        push(R_XB + 2 * 256, OP_LAST, 2, 0, 1, 0);
      }
      for (int l = 0; l < 2 * nb; ++l) {
        push(R_BLK + (l * 2 + 0) * 256, F01 | OP_WAIT_A, 0, 0, 1, 0);
        push(R_BLK + (l * 2 + 1) * 256, F01 | OP_LAST, 1, 2, 3, 0);
      }
      // out_fc (N = 160 = 128 + 32 columns): three 160-wide accumulators, k-steps {0,1} | {2,3} | {4..7}
      push(R_OUT + 0 * 320, OP_WIDE | F01 | OP_WAIT_A, 0, 0, 1, 0);
      push(R_OUT + 0 * 320 + 256, OP_WIDE | OP_SMALL | F01, 0, 0, 1, 128);
      push(R_OUT + 1 * 320, OP_WIDE | OP_FIRST0, 1, 2, 2, 0);
      push(R_OUT + 1 * 320 + 256, OP_WIDE | OP_SMALL | OP_FIRST0 | OP_LAST, 1, 2, 2, 128);
    }
    m->n_stages = (int)prog.size();
    EG_CUDA_CHECK(cudaMalloc((void**)&m->prog, prog.size() * sizeof(uint2)));
    EG_CUDA_CHECK(cudaMemcpy(m->prog, prog.data(), prog.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  }
  return prepare_weights(m, w, st);
}

int refresh(MotionTc* m, const float* const* w, cudaStream_t st) {
  if (!m || (!m->dec_ok && !m->reg_ok)) return EG_OK;
  return prepare_weights(m, w, st);
}

void destroy(MotionTc* m) {
  if (!m) return;
  void* bufs[] = {m->Wf, m->bf, m->maxbits, m->scale, m->inv, m->WA, m->WB, m->WC, m->Hact, m->T1act, m->T2act,
                  m->RW, m->pack_dev, m->rbias, m->prog, m->base_buf, m->xb_buf};
  for (void* p : bufs) cudaFree(p);
  delete m;
}

bool decode_available(const MotionTc* m) { return m && m->dec_ok; }
bool regress_available(const MotionTc* m) { return m && m->reg_ok; }

int decode(MotionTc* m, const float* gi1, const float* h0, float* Y, int B, cudaStream_t st) {
  using namespace dtc;
  EG_REQUIRE(m && m->dec_ok, "tensor-core decode unavailable");
  const int n_rt = (B + ROWS - 1) / ROWS, rows = n_rt * ROWS;
  if (rows > m->cap_rows) {
    EG_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(m->Hact); cudaFree(m->T1act); cudaFree(m->T2act);
    m->Hact = m->T1act = m->T2act = nullptr; m->cap_rows = 0;
    EG_CUDA_CHECK(cudaMalloc((void**)&m->Hact, (size_t)2 * rows * H * sizeof(__half)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->T1act, (size_t)2 * rows * HM * sizeof(__half)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->T2act, (size_t)2 * rows * H * sizeof(__half)));
    int rc;
    if ((rc = encode_f16_2d(m, &m->mapH, m->Hact, 2 * rows, H, ROWS))) return rc;
    if ((rc = encode_f16_2d(m, &m->mapT1, m->T1act, 2 * rows, HM, ROWS))) return rc;
    if ((rc = encode_f16_2d(m, &m->mapT2, m->T2act, 2 * rows, H, ROWS))) return rc;
    m->cap_rows = rows;
  }
  DtcArgs a{h0, gi1, Y, m->Hact, m->T1act, m->T2act, m->cap_rows, m->b1, m->bhh, m->b2, m->bo, m->bf, m->scale, m->inv, B};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(S, n_rt, 1);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_tc_kernel, m->mapWA, m->mapWB, m->mapWC, m->mapH, m->mapT1, m->mapT2, a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EG_CUDA_CHECK(e);
  return EG_OK;
}

int regress(MotionTc* m, const float* Y, const float* betas, int B, float* xbc, cudaStream_t st) {
  using namespace rtc;
  EG_REQUIRE(m && m->reg_ok, "tensor-core regressor unavailable");
  const int M = B * FPE, n_tiles = (M + TBR - 1) / TBR;
  if (n_tiles > m->cap_tiles) {
    EG_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(m->base_buf); cudaFree(m->xb_buf); m->base_buf = m->xb_buf = nullptr; m->cap_tiles = 0;
    EG_CUDA_CHECK(cudaMalloc((void**)&m->base_buf, (size_t)n_tiles * TBR * HR * sizeof(float)));
    EG_CUDA_CHECK(cudaMalloc((void**)&m->xb_buf, (size_t)n_tiles * TBR * BDP * sizeof(float)));
    m->cap_tiles = n_tiles;
  }
  RtcArgs a{Y, betas, xbc, m->base_buf, m->xb_buf, m->rbias, m->inv + dtc::N_SC, m->prog, m->n_stages, M, m->d.reg_blocks, m->d.reg_recur};
  EG_LAUNCH(regressor_tc_kernel, n_tiles, THREADS, SMEM_BYTES, st, m->map256, m->map64, a);
  return EG_OK;
}

}  // namespace mtc
}  // namespace eg
