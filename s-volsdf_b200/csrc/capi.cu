// Library-wide C-ABI helpers: last-error string, version, engine query.
#include <stdarg.h>

#include "svs_common.cuh"

namespace svs {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace svs

extern "C" const char* svs_last_error(void) { return svs::g_err; }
extern "C" int svs_abi_version(void) { return SVS_ABI_VERSION; }
extern "C" int svs_has_engine(int engine) {
#ifdef SVS_WITH_TCGEN05
  return engine == SVS_ENGINE_FP32 || engine == SVS_ENGINE_BF16;
#else
  return engine == SVS_ENGINE_FP32;
#endif
}
