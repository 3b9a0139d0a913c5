// Library-wide C-ABI entry points and state.
#include "common.cuh"

namespace eg {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launch_count{0};
}  // namespace eg

extern "C" int eg_version(void) { return 100; }  // 0.1.0
extern "C" const char* eg_last_error(void) { return eg::g_last_error; }
extern "C" int64_t eg_launch_count(void) { return eg::g_launch_count.load(); }

// ---- kernel timing hooks (dominant kernel = fused LBS vertex kernel) -----------------------------
#include <vector>
namespace eg {
static bool g_prof_on = false;
static std::vector<ProfSlot> g_prof;
static size_t g_prof_used = 0;
static bool g_prof_open = false;
void prof_begin(cudaStream_t st, int64_t units) {
  if (!g_prof_on) return;
  if (g_prof_used == g_prof.size()) {
    ProfSlot s;
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    g_prof.push_back(s);
  }
  g_prof[g_prof_used].units = units;
  cudaEventRecord(g_prof[g_prof_used].a, st);
  g_prof_open = true;
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || !g_prof_open) return;
  cudaEventRecord(g_prof[g_prof_used].b, st);
  ++g_prof_used;
  g_prof_open = false;
}
}  // namespace eg

// ---- env-step stage timing: events between the stages of eg_env_step when enabled ---------------
namespace eg {
static bool g_stage_on = false;
static std::vector<cudaEvent_t> g_stage_ev;
static std::vector<int> g_stage_id;
void stage_mark(cudaStream_t st, int id) {
  if (!g_stage_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_stage_ev.push_back(e);
  g_stage_id.push_back(id);
}
}  // namespace eg

extern "C" int eg_stage_profile_enable(int on) {
  for (auto e : eg::g_stage_ev) cudaEventDestroy(e);
  eg::g_stage_ev.clear(); eg::g_stage_id.clear();
  eg::g_stage_on = on != 0;
  return EG_OK;
}

// ms_out[16]: time attributed to stage id k = sum over consecutive marks (prev -> mark with id k)
extern "C" int eg_stage_profile_read(double* ms_out, int n) {
  EG_REQUIRE(ms_out && n > 0, "bad arguments");
  for (int i = 0; i < n; ++i) ms_out[i] = 0.0;
  for (size_t i = 1; i < eg::g_stage_ev.size(); ++i) {
    const int id = eg::g_stage_id[i];
    if (id <= 0 || id >= n) continue;              // id 0 marks the start of a step
    EG_CUDA_CHECK(cudaEventSynchronize(eg::g_stage_ev[i]));
    float ms = 0.f;
    EG_CUDA_CHECK(cudaEventElapsedTime(&ms, eg::g_stage_ev[i - 1], eg::g_stage_ev[i]));
    ms_out[id] += ms;
  }
  return eg_stage_profile_enable(eg::g_stage_on ? 1 : 0);
}

extern "C" int eg_profile_enable(int on) {
  eg::g_prof_on = on != 0;
  eg::g_prof_used = 0;
  eg::g_prof_open = false;
  return EG_OK;
}

extern "C" int eg_profile_read(double* total_ms, int64_t* launches, int64_t* units) {
  EG_REQUIRE(total_ms && launches && units, "null pointer");
  double t = 0.0;
  int64_t u = 0;
  for (size_t i = 0; i < eg::g_prof_used; ++i) {
    EG_CUDA_CHECK(cudaEventSynchronize(eg::g_prof[i].b));
    float ms = 0.f;
    EG_CUDA_CHECK(cudaEventElapsedTime(&ms, eg::g_prof[i].a, eg::g_prof[i].b));
    t += ms;
    u += eg::g_prof[i].units;
  }
  *total_ms = t;
  *launches = (int64_t)eg::g_prof_used;
  *units = u;
  eg::g_prof_used = 0;
  return EG_OK;
}
