// Library-wide C-ABI entry points and state.
#include "common.cuh"

namespace eg {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launch_count{0};
}  // namespace eg

extern "C" int eg_version(void) { return 100; }  // 0.1.0
extern "C" const char* eg_last_error(void) { return eg::g_last_error; }
extern "C" int64_t eg_launch_count(void) { return eg::g_launch_count.load(); }
