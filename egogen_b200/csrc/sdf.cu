// SDF-grid kernels: trilinear sample (calc_sdf), penetration count, sphere-traced ego depth.
// HBM/L2-bound gather work: one thread per query point, coalesced point reads, the 3-D grid stays
// resident (256^3 fp32 = 67 MB < 126 MB L2) so the 8 corner gathers are L2 hits after first touch.
#include "common.cuh"

namespace eg {

__global__ void __launch_bounds__(256)
sdf_sample_kernel(SdfGrid g, const float* __restrict__ pts, int64_t P, float* __restrict__ val,
                  int32_t* __restrict__ base_idx) {
  const float cx = __ldg(g.center), cy = __ldg(g.center + 1), cz = __ldg(g.center + 2);
  const float s = __ldg(g.scale);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
    int ox, oy, oz;
    const float v = sdf_sample_point(g, cx, cy, cz, s, x, y, z, ox, oy, oz);
    val[i] = v;
    if (base_idx) {
      base_idx[3 * i] = ox;
      base_idx[3 * i + 1] = oy;
      base_idx[3 * i + 2] = oz;
    }
  }
}

// one CTA per body: count sdf<0 over non-skipped vertices (crowd_env_2f.py:171,174-175)
__global__ void __launch_bounds__(256)
penetration_count_kernel(const float* __restrict__ sdf, int V, const uint8_t* __restrict__ skip,
                         int32_t* __restrict__ counts) {
  const float* row = sdf + (int64_t)blockIdx.x * V;
  int c = 0;
  for (int v = threadIdx.x; v < V; v += blockDim.x)
    c += (row[v] < 0.0f && !(skip && skip[v])) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    counts[blockIdx.x] = t;
  }
}

// Sphere tracing through the calc_sdf field: t += max(d, hit_eps) until d < hit_eps, t > max_range
// or max_steps (oracle/ego_depth.py defines the same loop). One thread per ray; rays of one camera
// are adjacent so neighbouring threads walk neighbouring cells.
__global__ void __launch_bounds__(256)
ego_depth_kernel(SdfGrid g, const float* __restrict__ cam, int A, int H, int W, float fx, float fy,
                 float max_range, int max_steps, float hit_eps, float* __restrict__ depth,
                 int32_t* __restrict__ steps_out) {
  const float cx = __ldg(g.center), cy = __ldg(g.center + 1), cz = __ldg(g.center + 2);
  const float s = __ldg(g.scale);
  const int64_t total = (int64_t)A * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i / (H * W));
    const int r = (int)(i % (H * W));
    const int py = r / W, px = r % W;
    const float* c = cam + 12 * a;
    const float u = ((float)px + 0.5f - 0.5f * (float)W) / fx;
    const float v = ((float)py + 0.5f - 0.5f * (float)H) / fy;
    float dx = c[9] + u * c[3] - v * c[6];
    float dy = c[10] + u * c[4] - v * c[7];
    float dz = c[11] + u * c[5] - v * c[8];
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    float t = 0.0f;
    int it = 0;
    for (; it < max_steps; ++it) {
      int ox, oy, oz;
      const float d = sdf_sample_point(g, cx, cy, cz, s, c[0] + t * dx, c[1] + t * dy, c[2] + t * dz,
                                       ox, oy, oz);
      if (d < hit_eps) break;
      t += fmaxf(d, hit_eps);
      if (t > max_range) { t = max_range; break; }
    }
    depth[i] = fminf(t, max_range);
    if (steps_out) steps_out[i] = it;
  }
}

// coarse[c] = min over fine indices [8c, min(8c+8, D-1)]^3 of (-grid): one thread per coarse cell
__global__ void __launch_bounds__(128)
sdf_coarse_kernel(const float* __restrict__ grid, int D0, int D1, int D2, int C0, int C1, int C2,
                  float* __restrict__ coarse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C0 * C1 * C2) return;
  const int c2 = i % C2, c1 = (i / C2) % C1, c0 = i / (C1 * C2);
  const int cs = 1 << kCoarseShift;
  float mn = INFINITY;
  for (int x = c0 * cs; x <= min(c0 * cs + cs, D0 - 1); ++x)
    for (int y = c1 * cs; y <= min(c1 * cs + cs, D1 - 1); ++y)
      for (int z = c2 * cs; z <= min(c2 * cs + cs, D2 - 1); ++z)
        mn = fminf(mn, -grid[((int64_t)x * D1 + y) * D2 + z]);
  coarse[i] = mn;
}

// dilated by one fine corner per side: min over [8c-1, 8c+9]^3. The tcgen05 LBS epilogue addresses this grid with an
// FMA-rounded index that can sit one fine cell off the exact grid_sample index right at a coarse-cell boundary; the
// extra corner keeps "bit clear => the exact sample cannot be negative" true for those points too.
__global__ void __launch_bounds__(128)
sdf_coarse_dilated_kernel(const float* __restrict__ grid, int D0, int D1, int D2, int C0, int C1, int C2, int shift,
                          float* __restrict__ coarse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C0 * C1 * C2) return;
  const int c2 = i % C2, c1 = (i / C2) % C1, c0 = i / (C1 * C2);
  const int cs = 1 << shift;
  float mn = INFINITY;
  for (int x = max(c0 * cs - 1, 0); x <= min(c0 * cs + cs + 1, D0 - 1); ++x)
    for (int y = max(c1 * cs - 1, 0); y <= min(c1 * cs + cs + 1, D1 - 1); ++y)
      for (int z = max(c2 * cs - 1, 0); z <= min(c2 * cs + cs + 1, D2 - 1); ++z)
        mn = fminf(mn, -grid[((int64_t)x * D1 + y) * D2 + z]);
  coarse[i] = mn;
}

// 1 bit per coarse cell: set when the cell may hold a negative sample (coarse <= 0). 4 KB for a 256^3 grid, small
// enough to sit in shared memory next to the tcgen05 LBS kernel's operand ring.
__global__ void sdf_coarse_bits_kernel(const float* __restrict__ coarse, int n, uint32_t* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool f = i < n && coarse[i] <= 0.0f;
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && i < ((n + 31) / 32) * 32) bits[i >> 5] = m;
}

// class of every fine cell (see SdfGrid::cell_class). The sample is sum_i w_i g_i over the cell's in-range corners with
// w_i >= 0: all g_i > 0 (and not so tiny that a product could round to zero) => sample > 0; no g_i > 0 => sample <= 0.
__global__ void __launch_bounds__(256)
sdf_cell_class_kernel(const float* __restrict__ grid, int D0, int D1, int D2, int64_t n, uint2* __restrict__ cls) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool all_pos = false, mixed = false;
  if (i < n) {
    const int z = (int)(i % D2), y = (int)((i / D2) % D1), x = (int)(i / ((int64_t)D1 * D2));
    int n_pos = 0, n_big = 0, n_c = 0;
    bool bad = false;
    for (int dx = 0; dx <= 1; ++dx)
      for (int dy = 0; dy <= 1; ++dy)
        for (int dz = 0; dz <= 1; ++dz) {
          if (x + dx > D0 - 1 || y + dy > D1 - 1 || z + dz > D2 - 1) continue;
          const float g = grid[((int64_t)(x + dx) * D1 + (y + dy)) * D2 + (z + dz)];
          ++n_c;
          if (!(g == g) || fabsf(g) > 1e30f) bad = true;
          if (g > 0.0f) ++n_pos;
          if (g > 1e-20f) ++n_big;
        }
    all_pos = !bad && n_big == n_c;
    mixed = bad || (n_pos > 0 && !all_pos);
  }
  const unsigned mp = __ballot_sync(0xffffffffu, all_pos), mm = __ballot_sync(0xffffffffu, mixed);
  if ((threadIdx.x & 31) == 0 && i < (n + 31) / 32 * 32) cls[i >> 5] = make_uint2(mp, mm);
}

static inline int grid_for(int64_t n, int block) {
  int64_t b = (n + block - 1) / block;
  const int64_t cap = (int64_t)kNumSMs * 16;  // multiple of the SM count, grid-stride beyond
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace eg

#include <algorithm>
#include <map>
#include <mutex>
namespace eg {
struct CoarseEntry { float* coarse; int D0, D1, D2, C0, C1, C2; uint32_t* fine_bits = nullptr; uint2* cell_class = nullptr; };
static std::map<const float*, CoarseEntry> g_coarse;
static std::mutex g_coarse_mu;
void sdf_attach_coarse(SdfGrid& g) {
  std::lock_guard<std::mutex> lk(g_coarse_mu);
  auto it = g_coarse.find(g.grid);
  if (it == g_coarse.end() || it->second.D0 != g.D0 || it->second.D1 != g.D1 || it->second.D2 != g.D2) return;
  g.coarse = it->second.coarse; g.C0 = it->second.C0; g.C1 = it->second.C1; g.C2 = it->second.C2;
  const int n = g.C0 * g.C1 * g.C2;
  g.coarse_bits = reinterpret_cast<const uint32_t*>(it->second.coarse + (n + 31) / 32 * 32);
  g.n_bit_words = (n + 31) / 32;
  g.fine_bits = it->second.fine_bits;
  g.cell_class = it->second.cell_class;
}
}  // namespace eg

using namespace eg;

extern "C" int eg_sdf_prepare(const float* grid, int D0, int D1, int D2, void* stream) {
  EG_REQUIRE(grid && D0 > 0 && D1 > 0 && D2 > 0, "bad arguments");
  const int cs = 1 << kCoarseShift;
  CoarseEntry e{nullptr, D0, D1, D2, (D0 + cs - 1) / cs, (D1 + cs - 1) / cs, (D2 + cs - 1) / cs};
  {
    std::lock_guard<std::mutex> lk(g_coarse_mu);
    auto it = g_coarse.find(grid);
    if (it != g_coarse.end()) { cudaFree(it->second.coarse); cudaFree(it->second.fine_bits); cudaFree(it->second.cell_class); g_coarse.erase(it); }
  }
  const int n = e.C0 * e.C1 * e.C2;
  const int n32 = (n + 31) / 32 * 32;                     // floats, then one bit per cell
  EG_CUDA_CHECK(cudaMalloc((void**)&e.coarse, (size_t)(n32 + n32 / 32) * sizeof(float)));
  EG_LAUNCH(sdf_coarse_kernel, (n + 127) / 128, 128, 0, as_stream(stream), grid, D0, D1, D2, e.C0, e.C1, e.C2, e.coarse);
  {
    // sign bits come from the dilated minimum; level 1 = 8^3 cells (shared-memory resident in the LBS kernel),
    // level 2 = 2^3 cells (global, consulted only where level 1 is set)
    const int f0 = (D0 + 1) / 2, f1 = (D1 + 1) / 2, f2 = (D2 + 1) / 2;
    const int64_t nf = (int64_t)f0 * f1 * f2, nf32 = (nf + 31) / 32 * 32;
    float* dil = nullptr;
    EG_CUDA_CHECK(cudaMalloc((void**)&dil, (size_t)std::max<int64_t>(n32, nf32) * sizeof(float)));
    EG_CUDA_CHECK(cudaMalloc((void**)&e.fine_bits, (size_t)(nf32 / 32) * sizeof(uint32_t)));
    EG_LAUNCH(sdf_coarse_dilated_kernel, (n + 127) / 128, 128, 0, as_stream(stream), grid, D0, D1, D2, e.C0, e.C1, e.C2,
              kCoarseShift, dil);
    EG_LAUNCH(sdf_coarse_bits_kernel, n32 / 128 + 1, 128, 0, as_stream(stream), dil, n,
              reinterpret_cast<uint32_t*>(e.coarse + n32));
    EG_LAUNCH(sdf_coarse_dilated_kernel, (int)((nf + 127) / 128), 128, 0, as_stream(stream), grid, D0, D1, D2, f0, f1, f2, 1, dil);
    EG_LAUNCH(sdf_coarse_bits_kernel, (int)(nf32 / 128 + 1), 128, 0, as_stream(stream), dil, (int)nf, e.fine_bits);
    EG_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    cudaFree(dil);
    const int64_t nc = (int64_t)D0 * D1 * D2, nc32 = (nc + 31) / 32;
    EG_CUDA_CHECK(cudaMalloc((void**)&e.cell_class, (size_t)nc32 * sizeof(uint2)));
    EG_LAUNCH(sdf_cell_class_kernel, (int)((nc32 * 32 + 255) / 256), 256, 0, as_stream(stream), grid, D0, D1, D2, nc, e.cell_class);
  }
  EG_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  std::lock_guard<std::mutex> lk(g_coarse_mu);
  g_coarse[grid] = e;
  return EG_OK;
}

extern "C" int eg_sdf_release(const float* grid) {
  std::lock_guard<std::mutex> lk(g_coarse_mu);
  auto it = g_coarse.find(grid);
  if (it != g_coarse.end()) { cudaFree(it->second.coarse); cudaFree(it->second.fine_bits); cudaFree(it->second.cell_class); g_coarse.erase(it); }
  return EG_OK;
}

extern "C" int eg_sdf_sample(const float* grid, int D0, int D1, int D2, const float* center_dev,
                             const float* scale_dev, const float* pts, int64_t P, float* val,
                             int32_t* base_idx, void* stream) {
  EG_REQUIRE(D0 > 0 && D1 > 0 && D2 > 0 && P >= 0, "bad sizes");
  if (P == 0) return EG_OK;   // empty batch: nothing to do (torch gives NULL data pointers here)
  EG_REQUIRE(grid && center_dev && scale_dev && val && pts, "null pointer");
  SdfGrid g{grid, D0, D1, D2, center_dev, scale_dev};
  EG_LAUNCH(sdf_sample_kernel, grid_for(P, 256), 256, 0, as_stream(stream), g, pts, P, val, base_idx);
  return EG_OK;
}

extern "C" int eg_penetration_count(const float* sdf_vals, int N, int V, const uint8_t* skip_mask,
                                    int32_t* counts, void* stream) {
  EG_REQUIRE(N >= 0 && V > 0, "bad sizes");
  if (N == 0) return EG_OK;
  EG_REQUIRE(sdf_vals && counts, "null pointer");
  EG_LAUNCH(penetration_count_kernel, N, 256, 0, as_stream(stream), sdf_vals, V, skip_mask, counts);
  return EG_OK;
}

extern "C" int eg_ego_depth(const float* grid, int D0, int D1, int D2, const float* center_dev,
                            const float* scale_dev, const float* cam, int A, int H, int W, float fx,
                            float fy, float max_range, int max_steps, float hit_eps, float* depth,
                            int32_t* steps_out, void* stream) {
  EG_REQUIRE(A >= 0 && H > 0 && W > 0 && max_steps > 0, "bad sizes");
  if (A == 0) return EG_OK;       // empty batch (torch hands out NULL data pointers for empty tensors)
  EG_REQUIRE(grid && center_dev && scale_dev && cam && depth, "null pointer");
  SdfGrid g{grid, D0, D1, D2, center_dev, scale_dev};
  EG_LAUNCH(ego_depth_kernel, grid_for((int64_t)A * H * W, 256), 256, 0, as_stream(stream), g, cam, A,
            H, W, fx, fy, max_range, max_steps, hit_eps, depth, steps_out);
  return EG_OK;
}
