// ABI plumbing: version + thread-local error string.
#include "common.cuh"

namespace df3d {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace df3d

extern "C" int df3d_abi_version(void) { return DF3D_ABI_VERSION; }
extern "C" const char* df3d_last_error(void) { return df3d::g_err; }
