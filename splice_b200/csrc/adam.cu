// splice_b200 — fused multi-tensor Adam for the generator's 112 parameter tensors (1 037 523 elements).
//
// Replaces torch.optim.Adam.step() created by get_optimizer (util/util.py:28-32; lr 2e-3, betas (0, 0.99),
// eps 1e-8, no weight decay / amsgrad). Update rule of torch/optim/adam.py (single-tensor path):
//   m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g g;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM-bound (29 MB per step); one launch covers up to 64 tensors through a by-value pointer table, so the
// whole optimiser step is 2 launches instead of ~10 foreach kernels.
#include "adam.h"

namespace splice {

__global__ void __launch_bounds__(256) adam_kernel(AdamTable tab, float lr_over_bc1, float inv_bc2_sqrt, float b1, float b2,
                                                   float eps) {
    const int ti = blockIdx.y;
    const int n = tab.n[ti];
    const int base = blockIdx.x * 1024;
    if (base >= n) return;
    float* __restrict__ p = tab.p[ti];
    const float* __restrict__ g = tab.g[ti];
    float* __restrict__ m = tab.m[ti];
    float* __restrict__ v = tab.v[ti];
    const float w = 1.f - b1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = base + k * 256 + threadIdx.x;
        if (i < n) {
            const float gi = g[i];
            float mi = m[i], vi = v[i];
            mi = (w < 0.5f) ? mi + w * (gi - mi) : gi - (gi - mi) * (1.f - w);   // torch.lerp
            vi = vi * b2 + (1.f - b2) * gi * gi;
            const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
            p[i] -= lr_over_bc1 * (mi / denom);
            m[i] = mi;
            v[i] = vi;
        }
    }
}

__global__ void __launch_bounds__(256) accumulate_kernel(float* __restrict__ dst, AccTable t, int n_src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        float v = dst[i];
        for (int k = 0; k < n_src; ++k) v += t.src[k][i];
        dst[i] = v;
    }
}

int accumulate_f32(float* dst, const AccTable& t, int n_src, size_t n, cudaStream_t stream) {
    const size_t blocks = (n + 255) / 256;
    accumulate_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, stream>>>(dst, t, n_src, n);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int adam_step(const AdamTable& tab, int n_tensors, int max_n, float lr_over_bc1, float inv_bc2_sqrt, float b1, float b2,
              float eps, cudaStream_t stream) {
    SPLICE_REQUIRE(n_tensors > 0 && n_tensors <= ADAM_MAX_TENSORS, "adam: n_tensors %d out of range", n_tensors);
    dim3 grid(ceil_div(max_n, 1024), n_tensors);
    adam_kernel<<<grid, 256, 0, stream>>>(tab, lr_over_bc1, inv_bc2_sqrt, b1, b2, eps);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

}  // namespace splice
