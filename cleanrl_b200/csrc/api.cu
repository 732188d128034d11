// Error plumbing + version entry points of the C-ABI (include/b200rl.h).
#include "common.cuh"
#include <cstring>

namespace b200rl {
static thread_local char g_err[512] = {0};
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace b200rl

extern "C" int b200rl_version(void) { return 100; }  // 0.1.0
extern "C" const char* b200rl_last_error(void) { return b200rl::err_buf(); }
extern "C" int b200rl_compiled_arch(void) {
#ifdef B200RL_ARCH
    return B200RL_ARCH;
#else
    return 100;
#endif
}
