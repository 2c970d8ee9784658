// Library identification and the thread-local error string of libonmf_b200.so.
#include "common.cuh"

namespace onmf {
thread_local char g_err[512] = "";
}

extern "C" int onmf_version(void) { return 100; }
extern "C" const char* onmf_last_error(void) { return onmf::g_err; }
extern "C" int onmf_built_arch(void) { return 100; }
