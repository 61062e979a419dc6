// Error plumbing and version for the C ABI (include/nbm_b200.h).
#include "common.cuh"

namespace nbm {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace nbm

extern "C" const char *nbm_last_error(void) { return nbm::g_err; }
extern "C" int nbm_version(void) { return NBM_B200_VERSION; }
