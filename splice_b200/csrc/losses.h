// splice_b200 — loss kernels of the Splice objective (see losses.cu)
#pragma once
#include "common.cuh"

namespace splice {
// Row norms of keys [t, D] (leading dim ldk) and the split-bf16 operands of the Gram GEMM:
//   a_split[i] = [hi | lo | hi], b_split[i] = [hi | hi | lo]  of  khat_i = k_i / |k_i|   (rows [t, 3D], ld 3D)
// so that a_split · b_split^T = hi·hi + lo·hi + hi·lo ~ khat khat^T to ~2^-16.
int selfsim_prep(const float* keys, int ldk, int t, int D, bf16* a_split, bf16* b_split, float* inv_norm, cudaStream_t stream);
// khat_T[d, i] = bf16(keys[i, d] * inv_norm[i]),  [D, ldt] with ldt >= t (columns >= t zero-filled)
int selfsim_transpose(const float* keys, int ldk, const float* inv_norm, int t, int D, bf16* khat_T, int ldt, cudaStream_t stream);
// E = (2 / t^2) (Sx - Sa) -> bf16 [t, lde]; c[i] = sum_j E_ij Sx_ij; row_loss[i] = sum_j (Sx - Sa)^2 / t^2
int selfsim_err(const float* Sx, const float* Sa, int lds, int t, bf16* E16, int lde, float* c, float* row_loss, cudaStream_t stream);
// dK[i,:] = coef * 2 * inv_norm[i] * (R[i,:] - c[i] * keys[i,:] * inv_norm[i])
int selfsim_grad(const float* R, int ldr, const float* keys, int ldk, const float* inv_norm, const float* c, float coef,
                 float* dK, int lddk, int t, int D, cudaStream_t stream);
// row_loss[r] = sum_c (a[r,c]-b[r,c])^2 * inv_count; if grad: grad[r,c] = coef * 2 * inv_count * (a - b)
int mse_rows(const float* a, int lda, const float* b, int ldb, int rows, int cols, float inv_count, float coef, float* grad,
             int ldg, float* row_loss, cudaStream_t stream);
// out[0] = scale * sum(partial[0..n))   (single block, deterministic)
int reduce_sum(const float* partial, int n, float scale, float* out, cudaStream_t stream);
// total[0] = sum_i w[i] * terms[i] for i < n (n <= 8), weights passed by value
int weighted_total(const float* terms, const float* w_host, int n, float* total, cudaStream_t stream);
}  // namespace splice
