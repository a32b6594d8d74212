// capi.cu -- library-wide pieces of the C ABI (version, error string, launch counter).
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"

namespace tda {
char* tls_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
int64_t& launch_counter() {
  static thread_local int64_t c = 0;
  return c;
}
}  // namespace tda

extern "C" int tda_version(void) { return 100; }
extern "C" const char* tda_last_error(void) { return tda::tls_error_buffer(); }
extern "C" int64_t tda_launch_count(void) { return tda::launch_counter(); }
extern "C" void tda_launch_count_reset(void) { tda::launch_counter() = 0; }
