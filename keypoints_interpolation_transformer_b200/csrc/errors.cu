#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace kit {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
bool pdl_enabled() {
  static const bool on = []() {
    const char* v = getenv("KIT_PDL");
    return v == nullptr || v[0] != '0';
  }();
  return on;
}
}  // namespace kit

extern "C" const char* kit_last_error(void) { return kit::get_error(); }
extern "C" int kit_version(void) { return KIT_ABI_VERSION; }
