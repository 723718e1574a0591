// tcgen05 / TMA fused transform path.  (Being brought up: until the kernels land every shape reports
// "not supported" and the generic SIMT kernels in sc_generic.cu run instead.)
#include "sc_fast.h"

namespace sc {

bool fast_plan_init(Plan*) { return true; }
void fast_plan_destroy(Plan*) {}
bool fast_can_analyze(const Plan*, bool) { return false; }
bool fast_can_synthesize(const Plan*, bool) { return false; }
bool fast_analyze(const Plan*, const float*, int64_t, float2*, bool, cudaStream_t) {
  set_error("fast path not available for this shape");
  return false;
}
bool fast_synthesize(const Plan*, const float2*, int64_t, int, const float*, float*, bool, cudaStream_t) {
  set_error("fast path not available for this shape");
  return false;
}

}  // namespace sc
