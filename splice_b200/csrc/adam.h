// splice_b200 — fused multi-tensor Adam (see adam.cu)
#pragma once
#include "common.cuh"

namespace splice {
static constexpr int ADAM_MAX_TENSORS = 64;
struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    int n[ADAM_MAX_TENSORS];
};
static constexpr int ACC_MAX_SRC = 4;
struct AccTable { const float* src[ACC_MAX_SRC]; };
// dst += src[0] + ... + src[n_src-1] (fixed order): folds per-call gradient buffers into .grad
int accumulate_f32(float* dst, const AccTable& t, int n_src, size_t n, cudaStream_t stream);
int adam_step(const AdamTable& tab, int n_tensors, int max_n, float lr_over_bc1, float inv_bc2_sqrt, float b1, float b2,
              float eps, cudaStream_t stream);
}  // namespace splice
