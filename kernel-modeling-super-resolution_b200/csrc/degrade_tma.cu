// degrade_tma.cu -- placeholder until the TMA row-streaming kernel lands.
#include "common.cuh"
namespace kmsr {
bool tma_shape_ok(const DegradeArgs&, const char** why) { *why = "not built"; return false; }
int launch_degrade_tma(const DegradeArgs&, cudaStream_t) { set_error("TMA kernel not built"); return KMSR_E_UNSUPPORTED; }
}
