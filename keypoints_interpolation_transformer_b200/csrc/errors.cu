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
// SMs the persistent kernels leave free (kit_set_sm_reserve / KIT_SM_RESERVE): under data parallelism the NCCL all-reduce
// kernels run beside backward; the compute kernels are ONE wave of persistent CTAs that fill an SM, so a collective's CTAs
// could only start at kernel boundaries and then displaced CTAs of the next kernel into a second wave (+7 % on every kernel
// with the collectives resident, SCALE_r01).  With a few SMs reserved the two never compete.
// The reservation only matters while a collective can be in flight: from backward's first bucket callback to its end.  The
// engine switches it off for the forward pass and for the part of backward before the first bucket is complete
// (set_sm_reserve_active): those kernels -- two thirds of the step -- keep the whole GPU.
static int g_sm_reserve = -1;
static bool g_sm_reserve_active = true;
int sm_reserve() {
  if (g_sm_reserve < 0) {
    const char* v = getenv("KIT_SM_RESERVE");
    g_sm_reserve = v != nullptr ? atoi(v) : 0;
    if (g_sm_reserve < 0 || g_sm_reserve > 64) g_sm_reserve = 0;
    g_sm_reserve &= ~1;   // CTA pairs
  }
  return g_sm_reserve_active ? g_sm_reserve : 0;
}
void set_sm_reserve_active(bool on) { g_sm_reserve_active = on; }
void set_sm_reserve(int n) { g_sm_reserve = (n < 0 || n > 64) ? 0 : (n & ~1); }
}  // namespace kit

extern "C" int kit_set_sm_reserve(int32_t n) {
  kit::set_sm_reserve(n);
  return KIT_OK;
}
extern "C" const char* kit_last_error(void) { return kit::get_error(); }
extern "C" int kit_version(void) { return KIT_ABI_VERSION; }
