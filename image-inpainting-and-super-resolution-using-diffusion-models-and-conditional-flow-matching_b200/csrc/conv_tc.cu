// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands in this file).
#include "engine.h"
namespace cfm {
bool tc_conv_supported(const Engine&, const Op&) { return false; }
int tc_conv_prepare(Engine&, Op&, const std::vector<float>&, const std::vector<float>&) { return 0; }
int tc_conv_launch(Engine& e, const Op&, int, cudaStream_t) { e.err = "tcgen05 conv not built"; return CFM_ERR_INTERNAL; }
void tc_conv_release(Engine&) {}
bool gn_bf16_supported(const Engine&, const Op&) { return false; }
int gn_bf16_launch(Engine& e, const Op&, int, cudaStream_t) { e.err = "bf16 groupnorm not built"; return CFM_ERR_INTERNAL; }
}  // namespace cfm
