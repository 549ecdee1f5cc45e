// Error plumbing and small utility entry points of libtikeb200.
#include <cstdarg>
#include <cstdio>

#include "../../include/tike_b200.h"
#include "common.cuh"

namespace tb {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return TB_OK;
}

}  // namespace tb

extern "C" {

const char* tb_last_error(void) { return tb::last_error_ref().c_str(); }

int tb_version(void) { return 101; }

int tb_sm_count(int* count) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess)
    e = cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess)
    return tb::set_error((int)e, "tb_sm_count: %s", cudaGetErrorString(e));
  return TB_OK;
}

}  // extern "C"
