// splice_b200 — the three loss terms of LossG (util/losses.py:74-105) and their gradients.
//   * key self-similarity ("structure", losses.py:74-83 + extractor.py:4-9,158-163):
//       S = K K^T / clamp(|k_i||k_j|, 1e-8),  loss = mean (S_x - S_a)^2
//     row norms by warp shuffle over coalesced float4 reads; the two t x D x t Gram products and the
//     t x t x D gradient product run on the tcgen05 GEMM (gemm.cu) with split-bf16 operands (hi + lo) so
//     the similarity matrix keeps ~fp32 accuracy; the error / gradient passes below are HBM(L2)-bound
//     row kernels.  NOTE: rows with |k_i||k_j| < 1e-8 (never the case for LayerNorm-ed activations) would be
//     clamped by the reference; here they are normalised like any other row.
//   * [CLS] appearance (losses.py:85-94) and key identity (losses.py:96-105): plain MSE rows.
#include "losses.h"

namespace splice {

__global__ void __launch_bounds__(128) selfsim_prep_kernel(const float* __restrict__ keys, int ldk, int t, int D,
                                                           bf16* __restrict__ a_split, bf16* __restrict__ b_split,
                                                           float* __restrict__ inv_norm) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= t) return;
    const float4* k4 = reinterpret_cast<const float4*>(keys + (size_t)row * ldk);
    const int n4 = D >> 2;
    float ss = 0.f;
    for (int i = lane; i < n4; i += 32) {
        const float4 v = k4[i];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-20f);
    if (lane == 0) inv_norm[row] = inv;
    uint2* a = reinterpret_cast<uint2*>(a_split + (size_t)row * 3 * D);
    uint2* b = reinterpret_cast<uint2*>(b_split + (size_t)row * 3 * D);
    for (int i = lane; i < n4; i += 32) {
        const float4 v = k4[i];
        const float x0 = v.x * inv, x1 = v.y * inv, x2 = v.z * inv, x3 = v.w * inv;
        uint2 hi, lo;
        hi.x = pack_bf16x2(x0, x1);
        hi.y = pack_bf16x2(x2, x3);
        const float2 h01 = unpack_bf16x2(hi.x), h23 = unpack_bf16x2(hi.y);
        lo.x = pack_bf16x2(x0 - h01.x, x1 - h01.y);
        lo.y = pack_bf16x2(x2 - h23.x, x3 - h23.y);
        a[i] = hi; a[n4 + i] = lo; a[2 * n4 + i] = hi;
        b[i] = hi; b[n4 + i] = hi; b[2 * n4 + i] = lo;
    }
}

__global__ void __launch_bounds__(256) selfsim_transpose_kernel(const float* __restrict__ keys, int ldk,
                                                                const float* __restrict__ inv_norm, int t, int D,
                                                                bf16* __restrict__ out, int ldt) {
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r;
        tile[r][tx] = (i < t) ? keys[(size_t)i * ldk + d0 + tx] * inv_norm[i] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx;
        if (i < ldt) out[(size_t)(d0 + r) * ldt + i] = __float2bfloat16(tile[tx][r]);  // columns >= t are zero
    }
}

__global__ void __launch_bounds__(128) selfsim_err_kernel(const float* __restrict__ Sx, const float* __restrict__ Sa, int lds,
                                                          int t, bf16* __restrict__ E16, int lde, float* __restrict__ c,
                                                          float* __restrict__ row_loss) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= t) return;
    const float inv_tt = 1.f / ((float)t * (float)t);
    const float* sx = Sx + (size_t)row * lds;
    const float* sa = Sa + (size_t)row * lds;
    bf16* e = E16 + (size_t)row * lde;
    float cs = 0.f, ls = 0.f;
    for (int j = lane; j < t; j += 32) {
        const float x = sx[j], d = x - sa[j];
        const float ev = 2.f * inv_tt * d;
        e[j] = __float2bfloat16(ev);
        cs += ev * x;
        ls += d * d;
    }
    for (int j = t + lane; j < lde; j += 32) e[j] = __float2bfloat16(0.f);  // K-padding of the gradient GEMM
    cs = warp_sum(cs);
    ls = warp_sum(ls);
    if (lane == 0) {
        c[row] = cs;
        row_loss[row] = ls * inv_tt;
    }
}

__global__ void __launch_bounds__(128) selfsim_grad_kernel(const float* __restrict__ R, int ldr, const float* __restrict__ keys,
                                                           int ldk, const float* __restrict__ inv_norm,
                                                           const float* __restrict__ c, float coef, float* __restrict__ dK,
                                                           int lddk, int t, int D) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= t) return;
    const float inv = inv_norm[row], cr = c[row];
    const float4* r4 = reinterpret_cast<const float4*>(R + (size_t)row * ldr);
    const float4* k4 = reinterpret_cast<const float4*>(keys + (size_t)row * ldk);
    float4* o4 = reinterpret_cast<float4*>(dK + (size_t)row * lddk);
    const float f = coef * 2.f * inv, g = cr * inv;
    for (int i = lane; i < (D >> 2); i += 32) {
        const float4 r = r4[i], k = k4[i];
        o4[i] = make_float4(f * (r.x - g * k.x), f * (r.y - g * k.y), f * (r.z - g * k.z), f * (r.w - g * k.w));
    }
}

__global__ void __launch_bounds__(128) mse_rows_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                       int rows, int cols, float inv_count, float coef, float* __restrict__ grad,
                                                       int ldg, float* __restrict__ row_loss) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* ar = a + (size_t)row * lda;
    const float* br = b + (size_t)row * ldb;
    float ls = 0.f;
    const float gc = coef * 2.f * inv_count;
    for (int j = lane; j < cols; j += 32) {
        const float d = ar[j] - br[j];
        ls += d * d;
        if (grad) grad[(size_t)row * ldg + j] = gc * d;
    }
    ls = warp_sum(ls);
    if (lane == 0) row_loss[row] = ls * inv_count;
}

__global__ void __launch_bounds__(256) reduce_sum_kernel(const float* __restrict__ partial, int n, float scale,
                                                         float* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s * scale;
}

struct TotalW { float w[8]; };
__global__ void weighted_total_kernel(const float* __restrict__ terms, TotalW w, int n, float* __restrict__ total) {
    float s = 0.f;
    for (int i = 0; i < n; ++i)
        if (w.w[i] != 0.f) s += w.w[i] * terms[i];
    total[0] = s;
}

int selfsim_prep(const float* keys, int ldk, int t, int D, bf16* a_split, bf16* b_split, float* inv_norm, cudaStream_t stream) {
    SPLICE_REQUIRE(D % 4 == 0 && ldk % 4 == 0, "selfsim_prep: D/ldk must be multiples of 4");
    selfsim_prep_kernel<<<ceil_div(t, 4), 128, 0, stream>>>(keys, ldk, t, D, a_split, b_split, inv_norm);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int selfsim_transpose(const float* keys, int ldk, const float* inv_norm, int t, int D, bf16* khat_T, int ldt, cudaStream_t stream) {
    SPLICE_REQUIRE(D % 32 == 0 && ldt >= t, "selfsim_transpose: D %% 32, ldt >= t");
    dim3 grid(ceil_div(ldt, 32), D / 32);
    selfsim_transpose_kernel<<<grid, 256, 0, stream>>>(keys, ldk, inv_norm, t, D, khat_T, ldt);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int selfsim_err(const float* Sx, const float* Sa, int lds, int t, bf16* E16, int lde, float* c, float* row_loss, cudaStream_t stream) {
    selfsim_err_kernel<<<ceil_div(t, 4), 128, 0, stream>>>(Sx, Sa, lds, t, E16, lde, c, row_loss);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int selfsim_grad(const float* R, int ldr, const float* keys, int ldk, const float* inv_norm, const float* c, float coef,
                 float* dK, int lddk, int t, int D, cudaStream_t stream) {
    SPLICE_REQUIRE(D % 4 == 0 && ldr % 4 == 0 && ldk % 4 == 0 && lddk % 4 == 0, "selfsim_grad: alignment");
    selfsim_grad_kernel<<<ceil_div(t, 4), 128, 0, stream>>>(R, ldr, keys, ldk, inv_norm, c, coef, dK, lddk, t, D);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int mse_rows(const float* a, int lda, const float* b, int ldb, int rows, int cols, float inv_count, float coef, float* grad,
             int ldg, float* row_loss, cudaStream_t stream) {
    mse_rows_kernel<<<ceil_div(rows, 4), 128, 0, stream>>>(a, lda, b, ldb, rows, cols, inv_count, coef, grad, ldg, row_loss);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int reduce_sum(const float* partial, int n, float scale, float* out, cudaStream_t stream) {
    reduce_sum_kernel<<<1, 256, 0, stream>>>(partial, n, scale, out);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
int weighted_total(const float* terms, const float* w_host, int n, float* total, cudaStream_t stream) {
    SPLICE_REQUIRE(n <= 8, "weighted_total: n <= 8");
    TotalW w;
    for (int i = 0; i < 8; ++i) w.w[i] = i < n ? w_host[i] : 0.f;
    weighted_total_kernel<<<1, 1, 0, stream>>>(terms, w, n, total);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

}  // namespace splice
