// splice_b200 - generator kernels shared by the two generator engines: generator.cu (the default-argument skip() network of
// the optimisation loop) and generator_x.cu (other skip() configurations: inversion.py's 6-scale, 7x7 / 5x5, reflection-padded
// variant). Concat + up-sampling gather and its adjoint, the BatchNorm backward family, sigmoid backward, weight-gradient fold, the
// epilogues of convolutions whose reduction was split over several CTAs.
// `static`: each translation unit gets its own copy (no relocatable device code in this build).
#pragma once
#include "gen_dev.cuh"

namespace splice {

static constexpr int TH = 8, TW = 32;    // output tile of the conv kernels (256 threads, one pixel each)

// block-wide sums of NV values over 256 threads; result broadcast to every thread. red: >= 8*NV floats.
template <int NV>
__device__ __forceinline__ void block_reduce_vec(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[w * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += red[j * NV + i];
        v[i] = s;
    }
}


// concat( crop(lrelu(bn(s_raw))), crop(bilinear_x2(T(u_raw))) ) -> cat raw, + statistics per channel
static __global__ void __launch_bounds__(256) cat_build_kernel(const float* __restrict__ s_raw, int Cs, int Hs, int Ws, InTf tf_s,
                                                        int offy_s, int offx_s, const float* __restrict__ u_raw, int Cu, int Hu,
                                                        int Wu, InTf tf_u, int offy_u, int offx_u, float* __restrict__ cat, int H,
                                                        int W, float* __restrict__ stats_part, BnFin fin) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    __shared__ float red[8];
    __shared__ int s_flag;
    const int tiles_x = (W + TW - 1) / TW;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int c = blockIdx.y, n = blockIdx.z, C = Cs + Cu;
    const int y = ty0 + (threadIdx.x >> 5), x = tx0 + (threadIdx.x & 31);
    const bool valid = y < H && x < W;
    float v = 0.f;
    if (valid) {
        if (c < Cs) {
            v = apply_tf(tf_s, c, s_raw[((size_t)(n * Cs + c) * Hs + y + offy_s) * Ws + x + offx_s]);
        } else {
            // nn.Upsample(scale_factor=2, mode='bilinear'), align_corners=False: src = (dst + 0.5) / 2 - 0.5, clamped at 0
            const int cu = c - Cs;
            const int Y = y + offy_u, X = x + offx_u;
            float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = min(y0 + 1, Hu - 1), x1 = min(x0 + 1, Wu - 1);
            const float ly = sy - y0, lx = sx - x0;
            const float* p = u_raw + (size_t)(n * Cu + cu) * Hu * Wu;
            const float v00 = apply_tf(tf_u, cu, p[(size_t)y0 * Wu + x0]), v01 = apply_tf(tf_u, cu, p[(size_t)y0 * Wu + x1]);
            const float v10 = apply_tf(tf_u, cu, p[(size_t)y1 * Wu + x0]), v11 = apply_tf(tf_u, cu, p[(size_t)y1 * Wu + x1]);
            v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
        }
        cat[((size_t)(n * C + c) * H + y) * W + x] = v;
    }
    float a[1] = {valid ? v : 0.f};
    block_reduce_vec<1>(a, red);
    const int th = (H - ty0 < TH) ? H - ty0 : TH, tw = (W - tx0 < TW) ? W - tx0 : TW;
    const float cnt = (float)(th * tw), mean = a[0] / cnt;
    const float d = v - mean;
    a[0] = valid ? d * d : 0.f;
    block_reduce_vec<1>(a, red);
    if (threadIdx.x == 0) {
        const size_t pb = (size_t)n * gridDim.x + blockIdx.x;
        float* o = stats_part + (pb * C + c) * 3;
        o[0] = cnt; o[1] = mean; o[2] = a[0];
    }
    bn_finish_if_last(stats_part, gridDim.x * gridDim.z, C, c, 1, c, gridDim.x * gridDim.z, fin, &s_flag);
}


// -------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------
// per-channel sums of dz and dz*yhat over (N, H, W): block = 2048 pixels of one (n, c)
static __global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ y,
                                                            const float4* __restrict__ konst, int lrelu, int C, int HW,
                                                            float* __restrict__ part) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    __shared__ float red[16];
    const int c = blockIdx.y, n = blockIdx.z;
    const float4 k = konst[c];
    const size_t base = (size_t)(n * C + c) * HW;
    float a[2] = {0.f, 0.f};
    for (int i = blockIdx.x * 2048 + threadIdx.x; i < min(HW, (int)(blockIdx.x + 1) * 2048); i += 256) {
        const float yv = y[base + i];
        float dz = dA[base + i];
        if (lrelu && !(fmaf(k.z, yv, k.w) > 0.f)) dz *= LRELU;
        a[0] += dz;
        a[1] += dz * (yv - k.x) * k.y;
    }
    block_reduce_vec<2>(a, red);
    if (threadIdx.x == 0) {
        float* o = part + (((size_t)n * gridDim.x + blockIdx.x) * C + c) * 2;
        o[0] = a[0]; o[1] = a[1];
    }
}
// dgamma += sum dz*yhat, dbeta += sum dz, (m1, m2) = sums / count
static __global__ void __launch_bounds__(32) bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, double count,
                                                             float* dgamma, float* dbeta, float2* __restrict__ m, int accumulate) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const int c = blockIdx.x, lane = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < nparts; i += 32) {
        s1 += part[((size_t)i * C + c) * 2];
        s2 += part[((size_t)i * C + c) * 2 + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
        m[c] = make_float2((float)(s1 / count), (float)(s2 / count));
    }
}

// dA <- d(raw conv output) in place: dy = a * (dz - m1 - yhat * m2), dz = dA * LeakyReLU'(z)  (large layers)
static __global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dA, const float* __restrict__ y,
                                                           const float4* __restrict__ konst, const float2* __restrict__ m, int lrelu,
                                                           int C, int HW, size_t total) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c = (i / HW) % C;
        const float4 k = konst[c];
        const float2 mm = m[c];
        const float yv = y[i];
        float dz = dA[i];
        if (lrelu && !(fmaf(k.z, yv, k.w) > 0.f)) dz *= LRELU;
        dA[i] = k.z * (dz - mm.x - (yv - k.x) * k.y * mm.y);
    }
}
// the whole BatchNorm backward of one channel in one block (small layers: N*H*W <= 8192):
// reduce, dgamma/dbeta accumulation, and the in-place dA -> dy rewrite
static __global__ void __launch_bounds__(256) bn_bwd_small_kernel(float* __restrict__ dA, const float* __restrict__ y,
                                                           const float4* __restrict__ konst, int lrelu, int N, int C, int HW,
                                                           float* dgamma, float* dbeta, int accumulate) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    __shared__ float red[16];
    const int c = blockIdx.x, count = N * HW;
    const float4 k = konst[c];
    float a[2] = {0.f, 0.f};
    for (int e = threadIdx.x; e < count; e += 256) {
        const size_t idx = ((size_t)(e / HW) * C + c) * HW + (e % HW);
        const float yv = y[idx];
        float dz = dA[idx];
        if (lrelu && !(fmaf(k.z, yv, k.w) > 0.f)) dz *= LRELU;
        a[0] += dz;
        a[1] += dz * (yv - k.x) * k.y;
    }
    block_reduce_vec<2>(a, red);
    if (threadIdx.x == 0) {
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + a[1];
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + a[0];
    }
    const float m1 = a[0] / count, m2 = a[1] / count;
    for (int e = threadIdx.x; e < count; e += 256) {
        const size_t idx = ((size_t)(e / HW) * C + c) * HW + (e % HW);
        const float yv = y[idx];
        float dz = dA[idx];
        if (lrelu && !(fmaf(k.z, yv, k.w) > 0.f)) dz *= LRELU;
        dA[idx] = k.z * (dz - m1 - (yv - k.x) * k.y * m2);
    }
}
// d(sigmoid output) -> d(pre-sigmoid) for the final 1x1 conv
static __global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                          float* __restrict__ dy, size_t total) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const float o = out[i];
        dy[i] = dout[i] * o * (1.f - o);
    }
}


// grad_w += sum over chunks, grad_b += sum over chunks (fixed order)
static __global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int nchunks, size_t nW, int Cout,
                                                           float* __restrict__ gw, float* __restrict__ gb, int accumulate) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    const size_t tot = nW + Cout;
    if (i >= tot) return;
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * tot + i];
    if (i < nW) gw[i] = (accumulate ? gw[i] : 0.f) + s;
    else gb[i - nW] = (accumulate ? gb[i - nW] : 0.f) + s;
}

// adjoint of cat_build for the skip branch: d(s activated) = d(cat)[:, :Cs] placed at the crop offset, zero elsewhere
static __global__ void __launch_bounds__(256) cat_bwd_skip_kernel(const float* __restrict__ dcat, int C, int H, int W, int Cs, int Hs, int Ws, int offy,
                                                           int offx, float* __restrict__ dS) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const size_t total = (size_t)gridDim.z * Cs * Hs * Ws;
    const int n = blockIdx.z;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < (size_t)Cs * Hs * Ws; i += (size_t)gridDim.x * 256) {
        const int c = i / ((size_t)Hs * Ws), y = (i / Ws) % Hs, x = i % Ws;
        const int cy = y - offy, cx = x - offx;
        float v = 0.f;
        if (cy >= 0 && cy < H && cx >= 0 && cx < W) v = dcat[((size_t)(n * C + c) * H + cy) * W + cx];
        dS[(size_t)n * Cs * Hs * Ws + i] = v;
    }
    (void)total;
}
// adjoint of the bilinear x2 up-sampling (+ crop): d(u activated)[n,cu,yu,xu] = sum over the <= 4x4 fine pixels that read it
static __global__ void __launch_bounds__(256) cat_bwd_up_kernel(const float* __restrict__ dcat, int C, int H, int W, int Cs, int Cu, int Hu, int Wu, int offy,
                                                         int offx, float* __restrict__ dU) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const int n = blockIdx.z;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < (size_t)Cu * Hu * Wu; i += (size_t)gridDim.x * 256) {
        const int cu = i / ((size_t)Hu * Wu), yu = (i / Wu) % Hu, xu = i % Wu;
        float acc = 0.f;
        for (int Y = 2 * yu - 1; Y <= 2 * yu + 2; ++Y) {
            if (Y < 0 || Y >= 2 * Hu) continue;
            const float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f);
            const int y0 = (int)sy, y1 = min(y0 + 1, Hu - 1);
            const float ly = sy - y0;
            const float wy = (y0 == yu ? 1.f - ly : 0.f) + (y1 == yu ? ly : 0.f);
            const int cy = Y - offy;
            if (wy == 0.f || cy < 0 || cy >= H) continue;
            for (int X = 2 * xu - 1; X <= 2 * xu + 2; ++X) {
                if (X < 0 || X >= 2 * Wu) continue;
                const float sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
                const int x0 = (int)sx, x1 = min(x0 + 1, Wu - 1);
                const float lx = sx - x0;
                const float wx = (x0 == xu ? 1.f - lx : 0.f) + (x1 == xu ? lx : 0.f);
                const int cx = X - offx;
                if (wx == 0.f || cx < 0 || cx >= W) continue;
                acc += wy * wx * dcat[((size_t)(n * C + Cs + cu) * H + cy) * W + cx];
            }
        }
        dU[(size_t)n * Cu * Hu * Wu + i] = acc;
    }
}


// split-K epilogue: y = bias + sum of partials, per-block statistics of the result and (last block of a channel) the
// BatchNorm constants. grid (ceil(N*HW / 1024), C): 256 threads x 4 pixels of one channel
static __global__ void __launch_bounds__(256) conv_finish_stats_kernel(const float* __restrict__ part, int splitK, const float* __restrict__ bias,
                                                                float* __restrict__ y, int N, int C, int HW,
                                                                float* __restrict__ stats_part, BnFin fin) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    __shared__ float red[8];
    __shared__ int s_flag;
    const int c = blockIdx.y, count = N * HW;
    const size_t total = (size_t)N * C * HW;
    const float b = bias[c];
    float v[4];
    bool ok[4];
    float a[1] = {0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = blockIdx.x * 1024 + i * 256 + threadIdx.x;
        ok[i] = e < count;
        v[i] = 0.f;
        if (ok[i]) {
            const size_t idx = ((size_t)(e / HW) * C + c) * HW + (e % HW);
            float acc = b;
            int k = 0;
            for (; k + 4 <= splitK; k += 4) {
                const float p0 = part[(size_t)k * total + idx], p1 = part[(size_t)(k + 1) * total + idx];
                const float p2 = part[(size_t)(k + 2) * total + idx], p3 = part[(size_t)(k + 3) * total + idx];
                acc += p0; acc += p1; acc += p2; acc += p3;
            }
            for (; k < splitK; ++k) acc += part[(size_t)k * total + idx];
            y[idx] = acc;
            v[i] = acc;
            a[0] += acc;
        }
    }
    block_reduce_vec<1>(a, red);
    const int nb = min(1024, count - (int)blockIdx.x * 1024);
    const float mean = a[0] / nb;
    a[0] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float d = v[i] - mean;
        a[0] += ok[i] ? d * d : 0.f;
    }
    block_reduce_vec<1>(a, red);
    if (threadIdx.x == 0) {
        float* o = stats_part + ((size_t)blockIdx.x * C + c) * 3;
        o[0] = (float)nb; o[1] = mean; o[2] = a[0];
    }
    bn_finish_if_last(stats_part, gridDim.x, C, c, 1, c, gridDim.x, fin, &s_flag);
}

// dst (=|+=) sum over the split partials
static __global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int splitK, size_t total,
                                                           float* __restrict__ dst, int accumulate) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        float v = accumulate ? dst[i] : 0.f;
        for (int k = 0; k < splitK; ++k) v += part[(size_t)k * total + i];
        dst[i] = v;
    }
}

}  // namespace splice
