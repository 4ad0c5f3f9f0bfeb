// Entry points that are declared in include/backpack_b200.h but not implemented yet.
#include "bp_host.h"
extern "C" int bp_linear_bias_act_fwd(const void*, const void*, const void*, void*, int64_t, int32_t, int32_t,
                                      int32_t, int32_t, void*) {
  return bp::fail(BP_ERR_UNSUPPORTED, "bp_linear_bias_act_fwd: not implemented yet");
}
