// tcgen05 dense layer for sm_100a: C[M,N] = act(alpha * A[M,K] * W[N,K]^T + bias (+ C)) + residual at fp32 accuracy.
//
// Replaces the fp32 SIMT tiles for the nn.Linear-forward shape (activations row-major, weight [out,in]: both operands
// K-major) of the policy / C-VAE / VPoser layers (reference models_policy_ppo.py:24-39,287-350, models_GAMMA_primitive.py).
// The parity bar on this path is 1e-4 relative through chains of tens of layers, so a single TF32 pass (2^-11 per
// operand) is not enough; this kernel runs the 3xTF32 split
//      A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo,   x_hi = x with the low 13 mantissa bits cleared, x_lo = x - x_hi
// (relative error ~2^-21) on the tensor cores: TMA stages the raw fp32 tiles (SWIZZLE_128B), four transform warps
// derive the lo copy of every staged tile (element-wise, so the swizzled layout is irrelevant; the raw tile IS the hi
// operand because kind::tf32 ignores the low 13 mantissa bits - tests/test_gpu_gemm_tc.py would catch a rounding unit), one
// thread issues 3 x 4 UMMAs (128x64x8, kind::tf32) per 32-wide k block into 64 fp32 TMEM columns.
// Small-M layers (M = 256 is the whole PPO batch) would leave most SMs idle, so the k range is split over a
// thread-block cluster of up to 8 CTAs whose partial tiles are reduced through distributed shared memory, and the
// fused epilogue (alpha, bias, beta, activation, residual) runs on the reduced rows only.
#include <cooperative_groups.h>
#include <cuda.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "nn.cuh"

#ifndef EG_GEMM_TC_REWRITE_HI
#define EG_GEMM_TC_REWRITE_HI 0
#endif

namespace eg {
namespace gtc {

constexpr int BM = 128, BK = 32;                     // BK fp32 = 128 B = one swizzle row; BN (64 or 128) is a template parameter
#ifndef EG_GEMM_TC_STAGES
#define EG_GEMM_TC_STAGES 2                             // ring depth (48 KB per stage at BN = 64); see DESIGN 6b for the 3 / 4-stage A/B
#endif
#ifndef EG_GEMM_TC_TS
#define EG_GEMM_TC_TS 1                                 // 1: the A lo tile lives in TMEM (UMMA TMEM-A operand), 3-stage ring in 96 KB
#endif
constexpr int A_BYTES = BM * BK * 4;                // 16 KB
constexpr int XF_WARPS = 8;                          // transform + epilogue warps
constexpr int THREADS = 64 + XF_WARPS * 32;          // 320
constexpr int XF_THREADS = XF_WARPS * 32;
// geometry that depends on the tile width
template <int BN, bool DEEP = false>
struct Geo {
  static constexpr int B_BYTES = BN * BK * 4;                   // 8 / 16 KB
  // DEEP: 4 (BN = 64) / 3 (BN = 128) stages, opt-in per launch through EG_GEMM_TC_DEEP
  // TS (default, not DEEP): the lo copy of the A tile is written to TMEM by the transform warps and read as the UMMA's
  // TMEM A operand, so a stage holds raw A, raw B and lo B only (32 KB at BN = 64): THREE stages fit the 96 KB the 2-stage
  // hi + lo ring used - two CTAs per SM stay co-resident (the PDL chain and the actor / critic stream overlap need that,
  // DESIGN 6b) with 1.5x the operand bytes in flight.
  static constexpr bool TS = !DEEP && EG_GEMM_TC_TS;
  static constexpr int STAGES = TS ? 3 : (DEEP ? (BN == 64 ? 4 : 3) : ((BN == 64 || EG_GEMM_TC_STAGES < 3) ? EG_GEMM_TC_STAGES : 3));
  static constexpr int STAGE_BYTES = TS ? (A_BYTES + 2 * B_BYTES) : 2 * (A_BYTES + B_BYTES);   // raw A, raw B, (lo A,) lo B
  static constexpr int OFF_LO = TS ? (A_BYTES + B_BYTES) : (A_BYTES + B_BYTES);                 // lo tiles inside a stage (TS: lo B only)
  static constexpr int TMEM_COLS = TS ? 256 : BN;               // accumulator + 3 x 32 columns of A lo, power of two
  static constexpr int TM_ALO = BN;                             // first A lo column
  static constexpr int EPI_COLS = BN / (XF_WARPS / 4);          // columns per epilogue thread (two warps share a TMEM lane quarter)
  static constexpr int RED_LD = BN + 1;                         // padded row of the split-k reduction buffer
  static constexpr int OFF_BARS = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
  static_assert(BM * RED_LD * 4 <= STAGES * STAGE_BYTES, "reduction buffer aliases the operand ring");
  static_assert((A_BYTES + B_BYTES) / 16 % XF_THREADS == 0 && B_BYTES / 16 % XF_THREADS == 0, "transform work split");
  static_assert(!TS || BN + STAGES * BK <= TMEM_COLS, "TMEM budget");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B shared-memory descriptor: LBO = 1, SBO = 1024 B, version 1 (same form as lbs_tc.cuh)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major descriptor for 32-bit operands (cute make_umma_desc<Major::MN>, Layout_MN_SW128_32B_Atom - the only MN-major
// shared-memory layout tf32 accepts): the tile is stored as 32-element (128 B) chunks along M/N, each chunk a [k][32]
// box of 128 B rows written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32 B chunks swizzled over 4 rows);
// layout type 1 = SWIZZLE_128B_BASE32B, LBO = stride between chunks, SBO = stride between 4-row k atoms (512 B)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((chunk_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, M = 128, N = 64; bit 15 / 16 = A / B is MN-major
template <bool A_MN, bool B_MN, int BN>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same instruction with the A operand in TMEM (lane = row, 8 consecutive 32-bit columns = one k step); A from TMEM is
// K-major by construction, so bit 15 stays clear whatever the layout of the raw A tile
template <bool B_MN, int BN>
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate) {
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
// distributed shared memory through 32-bit shared::cluster addresses (half the registers of generic pointers)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
  float v;
  asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));       // (pure load: ordered by the cluster.sync() around the reduction)
  return v;
}
__device__ __forceinline__ float lo_tf32(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float act_apply(float x, int act, float slope) {
  switch (act) {
    case ACT_TANH: return tanhf(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_LRELU: return x > 0.0f ? x : x * slope;
    default: return x;
  }
}

struct Params {
  float* C; int ldc;
  const float* bias;
  const float* residual; int ldr;
  int M, N, K;
  int act; float slope; int beta; float alpha;
  int kb_per_split;      // 32-wide k blocks per cluster rank
};

// A_MN / B_MN: the operand is stored with its M / N index contiguous (dW = dY^T X has both, dX = dY W has B) instead of
// its k index (nn.Linear forward). MN-major tiles are staged as 32-wide chunks, one TMA box [32 k][32 mn] each.
template <bool A_MN, bool B_MN, int BN, bool DEEP>
__global__ void __launch_bounds__(THREADS, 2)          // two CTAs per SM must stay possible: <= 96 registers after allocation granularity
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Params p) {
  namespace cg = cooperative_groups;
  using G = Geo<BN, DEEP>;
  constexpr int STAGES = G::STAGES, B_BYTES = G::B_BYTES, STAGE_BYTES = G::STAGE_BYTES, TMEM_COLS = G::TMEM_COLS, EPI_COLS = G::EPI_COLS,
                RED_LD = G::RED_LD, OFF_BARS = G::OFF_BARS;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;                  // [STAGES] TMA landed
  uint64_t* ready_bar = bars + STAGES;        // [STAGES] hi / lo copies written
  uint64_t* empty_bar = bars + 2 * STAGES;    // [STAGES] MMAs retired
  uint64_t* acc_bar = bars + 3 * STAGES;      // accumulators complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int split = blockIdx.z, n_split = gridDim.z;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb0 = min(split * p.kb_per_split, kb_total);
  const int nkb = min(kb0 + p.kb_per_split, kb_total) - kb0;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], XF_WARPS); mbar_init(&empty_bar[s], 1); }
      mbar_init(acc_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  // programmatic dependent launch: the next kernel of the stream may start its own prologue (barrier init, TMEM
  // allocation, descriptor fetch) now; this kernel touches global memory only after its predecessor has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES; const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        const uint32_t st = smem_base + s * STAGE_BYTES;
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) tma_load_2d(st + c * (BK * 128), &mapA, &full_bar[s], m0 + c * 32, (kb0 + i) * BK);
        } else {
          tma_load_2d(st, &mapA, &full_bar[s], (kb0 + i) * BK, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(st + A_BYTES + c * (BK * 128), &mapB, &full_bar[s], n0 + c * 32, (kb0 + i) * BK);
        } else {
          tma_load_2d(st + A_BYTES, &mapB, &full_bar[s], (kb0 + i) * BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: hi*hi + lo*hi + hi*lo per 8-wide k step =====
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES; const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      mbar_wait(&ready_bar[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t st = smem_base + s * STAGE_BYTES;
        constexpr uint32_t CH = BK * 128;                  // bytes of one 32-wide MN chunk ([32 k][128 B])
        const uint64_t a_hi = A_MN ? make_desc_mn(st, CH) : make_desc(st);
        const uint64_t b_hi = B_MN ? make_desc_mn(st + A_BYTES, CH) : make_desc(st + A_BYTES);
        const uint32_t lo_b_addr = G::TS ? st + A_BYTES + B_BYTES : st + 2 * A_BYTES + B_BYTES;
        const uint64_t a_lo = A_MN ? make_desc_mn(st + A_BYTES + B_BYTES, CH) : make_desc(st + A_BYTES + B_BYTES);   // (not TS)
        const uint64_t b_lo = B_MN ? make_desc_mn(lo_b_addr, CH) : make_desc(lo_b_addr);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          // one 8-wide k step: K-major advances 32 B inside the swizzle atom, MN-major one 8-row group (1024 B)
          const uint64_t oa = (uint64_t)(A_MN ? kk * 64 : kk * 2), ob = (uint64_t)(B_MN ? kk * 64 : kk * 2);
          if (G::TS) umma_tf32_ts<B_MN, BN>(tmem_base, tmem_base + (uint32_t)(G::TM_ALO + s * BK + kk * 8), b_hi + ob, (i | kk) ? 1u : 0u);
          else umma_tf32<A_MN, B_MN, BN>(tmem_base, a_lo + oa, b_hi + ob, (i | kk) ? 1u : 0u);
          umma_tf32<A_MN, B_MN, BN>(tmem_base, a_hi + oa, b_lo + ob, 1u);
          umma_tf32<A_MN, B_MN, BN>(tmem_base, a_hi + oa, b_hi + ob, 1u);
        }
        umma_commit(&empty_bar[s]);
        if (i == nkb - 1) umma_commit(acc_bar);
      }
      __syncwarp();
    }
  } else {
    // ===== transform warps: split the staged tiles into hi (in place) and lo copies; then the epilogue =====
    const int xt = tid - 64;                               // 0..XF_THREADS-1
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES; const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      const uint32_t raw = smem_base + s * STAGE_BYTES;                    // A then B, contiguous
      const uint32_t lo = raw + A_BYTES + B_BYTES;
      if (G::TS) {
        // lo B: element-wise into shared memory (layout-agnostic); lo A: this thread's row (TMEM lane), 16 of the 32 k
        // columns, read through the tile's swizzle and written to the stage's TMEM columns
        float4 xb[B_BYTES / 16 / XF_THREADS];
#pragma unroll
        for (int q = 0; q < B_BYTES / 16 / XF_THREADS; ++q) xb[q] = lds128(raw + A_BYTES + (uint32_t)(xt + q * XF_THREADS) * 16u);
        const int rq = warp & 3, hf = (warp - 2) >> 2, r = rq * 32 + lane;
        float av[16];
        if (A_MN) {
          // [32 k][32 m] chunks of 128 B rows, 32 B units XOR-swizzled with (k & 3) (SWIZZLE_128B_ATOM_32B)
          const uint32_t cb = raw + (uint32_t)rq * (BK * 128) + (uint32_t)((lane & 7) << 2);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int k = 16 * hf + e;
            av[e] = lds32f(cb + (uint32_t)(k * 128) + (uint32_t)((((lane >> 3) ^ (k & 3)) & 3) << 5));
          }
        } else {
          const uint32_t rb = raw + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 x = lds128(rb + (uint32_t)(((4 * hf + c) ^ (r & 7)) << 4));
            av[4 * c + 0] = x.x; av[4 * c + 1] = x.y; av[4 * c + 2] = x.z; av[4 * c + 3] = x.w;
          }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) av[e] = lo_tf32(av[e]);
        tmem_st16(tmem_base + ((uint32_t)(rq * 32) << 16) + (uint32_t)(G::TM_ALO + s * BK + 16 * hf), av);
#pragma unroll
        for (int q = 0; q < B_BYTES / 16 / XF_THREADS; ++q) {
          float4 l;
          l.x = lo_tf32(xb[q].x); l.y = lo_tf32(xb[q].y); l.z = lo_tf32(xb[q].z); l.w = lo_tf32(xb[q].w);
          sts128(lo + (uint32_t)(xt + q * XF_THREADS) * 16u, l);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[s]);
        continue;
      }
      float4 xs[(A_BYTES + B_BYTES) / 16 / XF_THREADS];
#pragma unroll
      for (int q = 0; q < (A_BYTES + B_BYTES) / 16 / XF_THREADS; ++q) xs[q] = lds128(raw + (uint32_t)(xt + q * XF_THREADS) * 16u);
#pragma unroll
      for (int q = 0; q < (A_BYTES + B_BYTES) / 16 / XF_THREADS; ++q) {
        const uint32_t off = (uint32_t)(xt + q * XF_THREADS) * 16u;
        const float4 x = xs[q];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
#if EG_GEMM_TC_REWRITE_HI
        sts128(raw + off, h);      // not needed: kind::tf32 reads the fp32 words and ignores the low 13 mantissa bits
#endif
        sts128(lo + off, l);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready_bar[s]);
    }
    // ---- epilogue: TMEM lane quarter (warp & 3) -> row, column half ((warp - 2) >> 2) -> EPI_COLS columns per thread ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;                         // row of the tile this thread owns in TMEM
    const int c0 = half * EPI_COLS;
    float acc[EPI_COLS];
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
#pragma unroll
      for (int c = 0; c < EPI_COLS / 16; ++c) tmem_ld16(taddr + c * 16, acc + c * 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int c = 0; c < EPI_COLS; ++c) acc[c] = 0.0f;
    }
    // partial tile -> shared memory (the operand ring is dead: acc_bar implies every MMA has retired); the fused
    // epilogue then runs with consecutive threads on consecutive columns (coalesced stores)
#pragma unroll
    for (int c = 0; c < EPI_COLS; ++c) sts32(smem_base + (uint32_t)(row * RED_LD + c0 + c) * 4u, acc[c]);
    if (n_split == 1) {
      asm volatile("bar.sync 1, %0;" ::"n"(XF_THREADS) : "memory");
#pragma unroll 8
      for (int e = xt; e < BM * BN; e += XF_THREADS) {
        const int r = e / BN, c = e % BN;
        const int m = m0 + r, n = n0 + c;
        if (m < p.M && n < p.N) {
          float v = p.alpha * lds32f(smem_base + (uint32_t)(r * RED_LD + c) * 4u);
          if (p.bias) v += __ldg(p.bias + n);
          if (p.beta) v += p.C[(int64_t)m * p.ldc + n];
          v = act_apply(v, p.act, p.slope);
          if (p.residual) v += p.residual[(int64_t)m * p.ldr + n];
          p.C[(int64_t)m * p.ldc + n] = v;
        }
      }
    }
  }
  if (n_split > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                         // every rank's partial tile is in its shared memory
    if (warp >= 2) {
      const int rows_per = (BM + n_split - 1) / n_split;    // rows this rank reduces and writes
      const int r_lo = split * rows_per, r_n = max(0, min(BM, r_lo + rows_per) - r_lo);
      const int xt = tid - 64;
      // four elements per pass and the rank loop unrolled: all remote loads of a pass are in flight together (the rolled
      // form waited one distributed-shared-memory round trip per rank and element); ranks are still summed in order
      uint32_t rk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) rk[k] = mapa_u32(smem_base, (uint32_t)(k < n_split ? k : 0));
      for (int e0 = xt; e0 < r_n * BN; e0 += 4 * XF_THREADS) {
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = min(e0 + u * XF_THREADS, r_n * BN - 1);
          off[u] = ((r_lo + e / BN) * RED_LD + e % BN) * 4;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < n_split) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] += ld_dsmem_f32(rk[k] + (uint32_t)off[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * XF_THREADS;
          if (e >= r_n * BN) break;
          const int r = r_lo + e / BN, c = e % BN;
          const int m = m0 + r, n = n0 + c;
          if (m < p.M && n < p.N) {
            float w = v[u] * p.alpha;
            if (p.bias) w += __ldg(p.bias + n);
            if (p.beta) w += p.C[(int64_t)m * p.ldc + n];
            w = act_apply(w, p.act, p.slope);
            if (p.residual) w += p.residual[(int64_t)m * p.ldr + n];
            p.C[(int64_t)m * p.ldc + n] = w;
          }
        }
      }
    }
    cluster.sync();                                         // remote shared memory must outlive the reads
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
  const void* p; int rows, cols, ld, box, mn;
  bool operator==(const MapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box && mn == o.mn;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (size_t)k.ld * 0x165667B19E3779F9ull + (size_t)k.box * 2 + (size_t)k.mn;
    return h;
  }
};

static EncodeTiledFn g_encode = nullptr;
static bool g_init_done = false;
static std::mutex g_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D fp32 tensor [rows][cols] with row pitch ld (elements), box = [32 cols][box_rows], SWIZZLE_128B, OOB -> 0
static bool get_map(const float* base, int rows, int cols, int ld, int box_rows, bool mn_major, CUtensorMap* out) {
  MapKey key{base, rows, cols, ld, box_rows, mn_major ? 1 : 0};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return true; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  if (g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return true;
}

}  // namespace gtc

static int g_gemm_tc_enabled = -1;          // -1: read EG_GEMM_TC (default on) at first use
void gemm_tc_set_enabled(int on) { g_gemm_tc_enabled = on ? 1 : 0; }

template <bool A_MN, bool B_MN, int BN>
static int launch_variant(const GemmArgs& g, cudaStream_t st) {
  using namespace gtc;
  constexpr int SMEM_BYTES = Geo<BN, false>::SMEM_BYTES, SMEM_DEEP = Geo<BN, true>::SMEM_BYTES;
  static bool attr_done = false;
  static int max_clusters[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};     // indexed by cluster size
  auto kernel = gemm_tc_kernel<A_MN, B_MN, BN, false>;
  auto kernel_deep = gemm_tc_kernel<A_MN, B_MN, BN, true>;
  if (!attr_done) {
    EG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    EG_CUDA_CHECK(cudaFuncSetAttribute(kernel_deep, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DEEP));
    attr_done = true;
  }
  CUtensorMap mapA, mapB;
  // K-major operand: tensor [rows = M or N][cols = K], box [32 k][BM or BN rows]
  // MN-major operand: tensor [rows = K][cols = M or N], box [32 mn][32 k]
  const bool okA = A_MN ? get_map(g.A, g.K, g.M, g.lda, BK, true, &mapA) : get_map(g.A, g.M, g.K, g.lda, BM, false, &mapA);
  const bool okB = B_MN ? get_map(g.B, g.K, g.N, g.ldb, BK, true, &mapB) : get_map(g.B, g.N, g.K, g.ldb, BN, false, &mapB);
  if (!okA || !okB) return 1;
  const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  const int kb_total = (g.K + BK - 1) / BK;
  // split-k factor: the largest cluster size whose clusters are all co-resident (a second wave of clusters would
  // double the layer's latency) and that leaves >= 2 k blocks per rank
  int sk = 1;
  for (int c = 8; c >= 2; --c) {
    if (kb_total / c < 2) continue;
    if (max_clusters[c] < 0) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(1, 1, c); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 1; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = c;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kernel, &q) != cudaSuccess) { cudaGetLastError(); nc = kNumSMs / c / 2; }
      max_clusters[c] = nc;
    }
    if (tiles <= max_clusters[c]) { sk = c; break; }
  }
  static const bool debug = getenv("EG_GEMM_TC_DEBUG") != nullptr;
  if (debug) fprintf(stderr, "[gemm_tc<%d,%d,%d>] M=%d N=%d K=%d tiles=%d sk=%d\n", (int)A_MN, (int)B_MN, BN, g.M, g.N, g.K, tiles, sk);
  Params p{g.C, g.ldc, g.bias, g.residual, g.ldr, g.M, g.N, g.K, g.act, g.slope, g.beta, g.alpha, (kb_total + sk - 1) / sk};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, sk);
  cfg.blockDim = dim3(THREADS);
  // deeper operand ring per launch (EG_GEMM_TC_DEEP, default 0 = never): 1 = always, 2 = forward layout only,
  // 3 = launches that leave at least half the SMs free for the next kernel's programmatic-dependent prologue
  static const int deep_mode = getenv("EG_GEMM_TC_DEEP") ? atoi(getenv("EG_GEMM_TC_DEEP")) : 0;
  const bool deep = deep_mode == 1 || (deep_mode == 2 && !A_MN && !B_MN) || (deep_mode == 3 && tiles * sk <= kNumSMs / 2);
  cfg.dynamicSmemBytes = deep ? SMEM_DEEP : SMEM_BYTES;
  cfg.stream = st;
  static const bool pdl = !(getenv("EG_GEMM_TC_PDL") != nullptr && getenv("EG_GEMM_TC_PDL")[0] == '0');
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = sk;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t e = deep ? cudaLaunchKernelEx(&cfg, kernel_deep, mapA, mapB, p) : cudaLaunchKernelEx(&cfg, kernel, mapA, mapB, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  EG_CUDA_CHECK(e);
  return EG_OK;
}

// Returns EG_OK when the layer was launched on the tensor-core path, 1 when the shape / layout is not eligible
// (the caller then uses the SIMT tiles), a negative EG_ERR_* on a CUDA failure.
//   !TA &&  TB : y = x W^T            (nn.Linear forward)        A K-major,  B K-major
//   !TA && !TB : dX = dY W            (input gradient)           A K-major,  B MN-major
//    TA && !TB : dW = dY^T X          (weight gradient)          A MN-major, B MN-major
int launch_gemm_tc(const GemmArgs& g, bool TA, bool TB, cudaStream_t st) {
  using namespace gtc;
  if (g_gemm_tc_enabled < 0) {
    const char* e = getenv("EG_GEMM_TC");
    g_gemm_tc_enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (!g_gemm_tc_enabled) return 1;
  if (TA && TB) return 1;
  if (g.a_div != 1 || g.K < 64 || g.N < 32 || g.M < 1) return 1;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(g.A) || !al(g.B) || (g.lda & 3) || (g.ldb & 3)) return 1;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_init_done) {
      g_init_done = true;
      cudaDriverEntryPointQueryResult qres;
      void* fn = nullptr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
          qres == cudaDriverEntryPointSuccess)
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
  }
  if (!g_encode) return 1;
  // 128-wide tiles (EG_GEMM_TC_BN128=1) cut the operand traffic per flop but halve the CTA count. Measured on the path's
  // shapes: no faster for M = 256 (18.8 us), slower for M >= 4096 (53.7 vs 41.9 us), and using them only where the
  // 64-wide tiling spills into a second wave (dW of the 1152 x 1152 layers, 162 vs 81 tiles) cost 2 % of the whole
  // iteration in an A/B on one box (50.35 k vs 51.3 k env-steps/s) - so 64 is the default everywhere.
  static const char* wide_env = getenv("EG_GEMM_TC_BN128");
  const bool wide = wide_env != nullptr && wide_env[0] == '1' && g.N >= 512;
  if (!TA && TB) return wide ? launch_variant<false, false, 128>(g, st) : launch_variant<false, false, 64>(g, st);
  if (!TA && !TB) return wide ? launch_variant<false, true, 128>(g, st) : launch_variant<false, true, 64>(g, st);
  return wide ? launch_variant<true, true, 128>(g, st) : launch_variant<true, true, 64>(g, st);
}

}  // namespace eg
