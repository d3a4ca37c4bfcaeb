// NerfMLP forward on the 5th-generation tensor cores (tcgen05 + TMEM).
// Placeholder translation unit: the entry points exist so the ABI is complete; the
// kernel lands in the next commit.
#include "common.cuh"
#include "../../include/nlb200.h"

extern "C" size_t nlb_nerf_mlp_packed_bytes(void) { return 0; }

extern "C" int nlb_nerf_mlp_pack(const nlb_nerf_mlp_weights_t*, void*, void*) {
  nlb_set_error("nerf_mlp_pack: tcgen05 kernel not built in this revision");
  return NLB_EUNSUPPORTED;
}

extern "C" int nlb_nerf_mlp_forward(const float*, const float*, int, int, const void*, float*, float*, float*, float*,
                                    void*) {
  nlb_set_error("nerf_mlp_forward: tcgen05 kernel not built in this revision");
  return NLB_EUNSUPPORTED;
}
