// Channels-last TMA + tcgen05 convolution (conv_nhwc.cu, conv_c3.cu).
#pragma once
#include "common.cuh"
namespace bf {
bool conv_nhwc_supported(int C, int N, int fy, int fx);
}  // namespace bf
