// conv_tc.cu -- tcgen05 implicit-GEMM convolution (placeholder until the TMA/UMMA kernel lands in this file)
#include "engine.h"
namespace bbocr {
bool conv_tc_supported(const ConvW&, const Act&, const Act&) { return false; }
void conv_tc_forward(Handle*, cudaStream_t, const ConvW&, const Act&, const Act&, Act&, int) {
    fail(BBOCR_E_UNSUPPORTED, "tcgen05 convolution not built");
}
}  // namespace bbocr
