// Library identification and the thread-local error string of libonmf_b200.so.
#include "common.cuh"

namespace onmf {
thread_local char g_err[512] = "";
thread_local long long g_launches = 0;
thread_local int g_lars_fast = 1;
thread_local int g_lars_reserved_sms = 0;   // per host thread, like the error string: engines on different threads do not interfere
}

extern "C" int onmf_version(void) { return 100; }
extern "C" const char* onmf_last_error(void) { return onmf::g_err; }
extern "C" int onmf_built_arch(void) { return 100; }
extern "C" long long onmf_launch_count(void) { return onmf::g_launches; }

extern "C" int onmf_set_option(int key, int value) {
  switch (key) {
    case ONMF_OPT_LARS_RESERVED_SMS:
      if (value < 0 || value > 64) return onmf::fail(ONMF_E_ARG, "set_option: reserved SMs must be in [0, 64]");
      onmf::g_lars_reserved_sms = value;
      return ONMF_OK;
    case ONMF_OPT_LARS_FAST_TIER:
      onmf::g_lars_fast = value != 0;
      return ONMF_OK;
  }
  return onmf::fail(ONMF_E_ARG, "set_option: unknown key");
}

extern "C" int onmf_get_option(int key, int* value) {
  if (!value) return onmf::fail(ONMF_E_ARG, "get_option: null pointer");
  switch (key) {
    case ONMF_OPT_LARS_RESERVED_SMS:
      *value = onmf::g_lars_reserved_sms;
      return ONMF_OK;
    case ONMF_OPT_LARS_FAST_TIER:
      *value = onmf::g_lars_fast;
      return ONMF_OK;
  }
  return onmf::fail(ONMF_E_ARG, "get_option: unknown key");
}
