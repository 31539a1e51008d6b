// splice_b200 — tcgen05 (kind::tf32, 3-product split) implicit-GEMM convolution of the generator, see conv_tc.cu
#pragma once
#include "gen_dev.cuh"

namespace splice {
// 3x3 stride-1 "same" convolution (DGRAD = false: y = conv(T(x)) + bias, + BatchNorm partials / last-ticket merge when
// stats_part != null) or its data gradient (DGRAD = true: x is dy [N, w_cout, H, W], y receives dx [N, w_cin, H, W]);
// weights in nn.Conv2d layout [w_cout][w_cin][3][3]. Same argument meaning as launch_conv_tiled (generator.cu).
template <bool DGRAD>
int launch_conv_tc(const float* x, int N, int Cin, int H, int W, InTf tf, const float* Wt, int w_cout, int w_cin, const float* bias,
                   int Cout, float* y, int accumulate, float* stats_part, BnFin fin, cudaStream_t st);
}  // namespace splice
