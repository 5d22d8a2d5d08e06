// Library-level entry points: version and thread-local error text.
#include "common.cuh"

namespace agnn {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace agnn

extern "C" int agnn_version(void) { return 100; }  // major*10000 + minor*100 + patch

extern "C" const char* agnn_last_error(void) { return agnn::error_buffer(); }
