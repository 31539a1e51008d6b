// splice_b200 — row-wise / elementwise kernels of the ViT path (see elementwise.cu)
#pragma once
#include "common.cuh"

namespace splice {
// y16[M,D] = LayerNorm(x[M,D]) * gamma + beta (eps), stats[M,2] = (mean, rstd) (optional)
int layernorm_fwd(const float* x, const float* gamma, const float* beta, bf16* y16, float* stats, int M, int D, float eps,
                  cudaStream_t stream);
// g_out = g_in + LayerNormBackward(dy; x, stats, gamma); g16 = bf16(g_out). g_out may alias g_in. g_in may be NULL (= 0).
// gamma may be NULL (= 1: the LayerNorm's gamma is folded into the weights of the GEMM that consumed it)
int layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* g_in, float* g_out,
                  bf16* g16, int M, int D, cudaStream_t stream);
// dst16 = bf16(src32), n elements (n % 4 == 0)
int cast_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t stream);
// x[s*t + 0, :] = cls[:] + pos[0, :] for s in [0, S); optionally the bf16 copy of those rows and their per-32-column
// (sum, centred sum of squares) partials [row][D/32] (LayerNorm folded into the GEMMs, gemm.h)
int write_cls_rows(float* x, const float* cls, const float* pos, int S, int t, int D, cudaStream_t stream, bf16* x16 = nullptr,
                   float2* stat_part = nullptr);
// dqkv16[s*t + r, col0 + c] += dk32[s*t + r, c]   (adds the loss gradient w.r.t. the layer-11 keys)
int add_f32_into_bf16_cols(bf16* dst, int ldd, int col0, const float* src, int lds, int rows, int cols, cudaStream_t stream);
}  // namespace splice
