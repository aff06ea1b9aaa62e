// lib.cu - error string, launch counter, version.
#include "common.cuh"

namespace acx {
static thread_local std::string g_error;
uint64_t g_launch_count = 0;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace acx

extern "C" {
const char* acx_last_error(void) { return acx::g_error.c_str(); }
int acx_version(void) { return 100; }
uint64_t acx_launch_count(void) { return acx::g_launch_count; }
void acx_reset_launch_count(void) { acx::g_launch_count = 0; }
}
