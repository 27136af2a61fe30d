// tcgen05 + TMA implicit-GEMM convolution (placeholder until the kernel lands in this round).
#include "snnb_internal.h"

namespace snnb {

bool conv2d_umma_supported(const ConvArgs&) { return false; }

int launch_conv2d_umma(snnb_context*, const ConvArgs&) {
    set_error("launch_conv2d_umma: not built");
    return 3;
}

} // namespace snnb
