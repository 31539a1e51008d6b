// splice_b200 — native generator: the reference's default-argument `skip()` U-Net, forward and backward.
//
// Replaces netG(x) (models/unet/skip.py:4-102 with models/unet/common.py:11-42,76-124: 26 zero-padded convs
// with bias, 30 BatchNorm2d that are ALWAYS in training mode, 25 LeakyReLU(0.2), 5 bilinear x2 up-samplings,
// 5 channel concats with centre crop, sigmoid) and its autograd backward (train.py:78), accumulating parameter
// gradients across the 2-3 netG calls of a step like autograd does.
//
// Fusion plan (fp32, NCHW):
//   * every conv kernel applies the PRODUCER's BatchNorm affine + LeakyReLU while loading its input tile
//     ("BN-apply folded into the consumer"), adds its bias, writes the raw output once and reduces the
//     per-channel batch statistics of that output in the same pass (per-block (count, mean, M2) partials,
//     merged deterministically with Chan's formula — no atomics, robust to large channel means);
//   * the backward never materialises d(conv output): dgrad/wgrad kernels rebuild it on the fly from
//     (d activated output, raw output, per-channel constants) = LeakyReLU' and the BatchNorm backward;
//   * concat + centre-crop + bilinear up-sampling are one gather kernel forward and one gather kernel
//     (adjoint) backward.
// Round-1 kernels are direct fp32 SIMT convolutions (bit-comparable to the reference's fp32 CPU path).
#include "generator.h"

namespace splice {

static constexpr int TH = 8, TW = 32;    // output tile of the conv kernels (256 threads, one pixel each)
static constexpr float LRELU = 0.2f;

struct InTf {               // per-channel transform applied to a raw tensor when it is consumed
    const float4* k;        // (mean, invstd, a, b): value -> a*value + b ; nullptr = identity
    int lrelu;
};

enum DyMode : int { DY_PLAIN = 0, DY_BN_LRELU = 1, DY_BN = 2, DY_SIGMOID = 3 };
struct DySrc {              // d(raw conv output) rebuilt from d(activated output)
    const float* dA;
    const float* y;         // raw conv output (DY_SIGMOID: the sigmoid output)
    const float4* k;        // (mean, invstd, a, b)
    const float2* m;        // (m1, m2) = (mean(dz), mean(dz * yhat))
    int mode;
};

__device__ __forceinline__ float apply_tf(const InTf& tf, int c, float v) {
    if (tf.k) {
        const float4 k = tf.k[c];
        v = fmaf(k.z, v, k.w);
        if (tf.lrelu && v < 0.f) v *= LRELU;
    }
    return v;
}
__device__ __forceinline__ float dy_value(const DySrc& s, int c, size_t idx) {
    const float g = s.dA[idx];
    if (s.mode == DY_PLAIN) return g;
    const float yv = s.y[idx];
    if (s.mode == DY_SIGMOID) return g * yv * (1.f - yv);
    const float4 k = s.k[c];
    const float2 m = s.m[c];
    float dz = g;
    if (s.mode == DY_BN_LRELU) {
        const float z = fmaf(k.z, yv, k.w);
        if (!(z > 0.f)) dz *= LRELU;
    }
    const float yhat = (yv - k.x) * k.y;
    return k.z * (dz - m.x - yhat * m.y);
}

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

// -------------------------------------------------------------------------------------------------
// forward convolution (+ producer BN/LeakyReLU on load, + bias, + optional sigmoid, + output statistics)
// -------------------------------------------------------------------------------------------------
template <int K, int S, int CO_T>
__global__ void __launch_bounds__(256) conv_fwd_kernel(const float* __restrict__ x, int Cin, int Hin, int Win, InTf tf,
                                                       const float* __restrict__ Wt, const float* __restrict__ bias, int Cout,
                                                       float* __restrict__ y, int Ho, int Wo, int out_sigmoid,
                                                       float* __restrict__ stats_part) {
    constexpr int CI_T = 8, PAD = (K - 1) / 2;
    constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    __shared__ float s_in[CI_T][IH][IW + 1];
    __shared__ __align__(16) float s_w[CI_T][K * K][CO_T];
    __shared__ float red[8 * CO_T];
    const int tiles_x = (Wo + TW - 1) / TW;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int co0 = blockIdx.y * CO_T, n = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int oy = ty0 + ty, ox = tx0 + tx;
    float acc[CO_T];
#pragma unroll
    for (int i = 0; i < CO_T; ++i) acc[i] = 0.f;

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
        for (int idx = threadIdx.x; idx < CI_T * IH * IW; idx += 256) {
            const int ci = idx / (IH * IW), r = (idx / IW) % IH, c = idx % IW;
            const int iy = ty0 * S - PAD + r, ix = tx0 * S - PAD + c;
            float v = 0.f;
            if (ci0 + ci < Cin && iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
                v = apply_tf(tf, ci0 + ci, x[((size_t)(n * Cin + ci0 + ci) * Hin + iy) * Win + ix]);
            s_in[ci][r][c] = v;   // zero padding lives in the post-BN/activation domain, like the reference
        }
        for (int idx = threadIdx.x; idx < CI_T * K * K * CO_T; idx += 256) {
            const int co = idx % CO_T, kk = (idx / CO_T) % (K * K), ci = idx / (CO_T * K * K);
            s_w[ci][kk][co] = (co0 + co < Cout && ci0 + ci < Cin) ? Wt[((size_t)(co0 + co) * Cin + ci0 + ci) * K * K + kk] : 0.f;
        }
        __syncthreads();
        const int cmax = (Cin - ci0 < CI_T) ? Cin - ci0 : CI_T;
        for (int ci = 0; ci < cmax; ++ci) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float v = s_in[ci][ty * S + ky][tx * S + kx];
#pragma unroll
                    for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(v, s_w[ci][ky * K + kx][co], acc[co]);
                }
        }
        __syncthreads();
    }
    const bool valid = oy < Ho && ox < Wo;
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
        if (co0 + co < Cout) {
            float v = acc[co] + bias[co0 + co];
            if (out_sigmoid) v = 1.f / (1.f + __expf(-v));
            acc[co] = v;
            if (valid) y[((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox] = v;
        }
    }
    if (stats_part) {
        // per-block (count, mean, M2) of each output channel: two block reductions, centred second pass
        float v[CO_T];
#pragma unroll
        for (int co = 0; co < CO_T; ++co) v[co] = valid ? acc[co] : 0.f;
        block_reduce_vec<CO_T>(v, red);
        const int th = (Ho - ty0 < TH) ? Ho - ty0 : TH, tw = (Wo - tx0 < TW) ? Wo - tx0 : TW;
        const float cnt = (float)(th * tw);
        float mean[CO_T];
#pragma unroll
        for (int co = 0; co < CO_T; ++co) {
            mean[co] = v[co] / cnt;
            const float d = acc[co] - mean[co];
            v[co] = valid ? d * d : 0.f;
        }
        block_reduce_vec<CO_T>(v, red);
        if (threadIdx.x < CO_T && co0 + threadIdx.x < Cout) {
            const int co = threadIdx.x;
            const size_t pb = (size_t)n * gridDim.x + blockIdx.x;
            float* o = stats_part + (pb * Cout + co0 + co) * 3;
            // mean[]/v[] are per-thread register arrays indexed by a runtime value: spill-free form below
            float mco = 0.f, vco = 0.f;
#pragma unroll
            for (int j = 0; j < CO_T; ++j)
                if (j == co) { mco = mean[j]; vco = v[j]; }
            o[0] = cnt; o[1] = mco; o[2] = vco;
        }
    }
}

// concat( crop(lrelu(bn(s_raw))), crop(bilinear_x2(T(u_raw))) ) -> cat raw, + statistics per channel
__global__ void __launch_bounds__(256) cat_build_kernel(const float* __restrict__ s_raw, int Cs, int Hs, int Ws, InTf tf_s,
                                                        int offy_s, int offx_s, const float* __restrict__ u_raw, int Cu, int Hu,
                                                        int Wu, InTf tf_u, int offy_u, int offx_u, float* __restrict__ cat, int H,
                                                        int W, float* __restrict__ stats_part) {
    __shared__ float red[8];
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
}

// merge the per-block (count, mean, M2) partials of one channel (one warp per channel, fixed order),
// produce (mean, invstd, a, b) and update the running statistics (momentum 0.1, unbiased variance)
__global__ void __launch_bounds__(32) bn_finalize_kernel(const float* __restrict__ part, int nparts, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                         float4* __restrict__ konst, float* running_mean, float* running_var,
                                                         long long* nbt, float momentum) {
    const int c = blockIdx.x, lane = threadIdx.x;
    double n = 0.0, mean = 0.0, M2 = 0.0;
    for (int i = lane; i < nparts; i += 32) {
        const float* p = part + ((size_t)i * C + c) * 3;
        const double nb = p[0], mb = p[1], Mb = p[2];
        if (nb > 0.0) {
            const double nn = n + nb, delta = mb - mean;
            mean += delta * nb / nn;
            M2 += Mb + delta * delta * n * nb / nn;
            n = nn;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                     Mb = __shfl_xor_sync(0xffffffffu, M2, o);
        const double nn = n + nb;
        if (nn > 0.0) {
            // symmetric merge so that both partners end with identical values
            const double delta = mb - mean;
            const double new_mean = (n * mean + nb * mb) / nn;
            M2 = M2 + Mb + delta * delta * n * nb / nn;
            mean = new_mean;
            n = nn;
        }
    }
    if (lane == 0) {
        const double var = M2 / n;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float a = gamma[c] * invstd;
        konst[c] = make_float4((float)mean, invstd, a, beta[c] - (float)mean * a);
        if (running_mean) {
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            const double unbiased = n > 1.0 ? M2 / (n - 1.0) : var;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            if (c == 0 && nbt) *nbt += 1;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------
// per-channel sums of dz and dz*yhat over (N, H, W): block = 2048 pixels of one (n, c)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ y,
                                                            const float4* __restrict__ konst, int lrelu, int C, int HW,
                                                            float* __restrict__ part) {
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
__global__ void __launch_bounds__(32) bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, double count,
                                                             float* dgamma, float* dbeta, float2* __restrict__ m) {
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
        if (dgamma) dgamma[c] += (float)s2;
        if (dbeta) dbeta[c] += (float)s1;
        m[c] = make_float2((float)(s1 / count), (float)(s2 / count));
    }
}

// d(transformed conv input) [N,Cin,Hin,Win] (= or +=)  from  d(conv output) rebuilt on the fly
template <int K, int S, int CI_T>
__global__ void __launch_bounds__(256) conv_dgrad_kernel(DySrc src, int Cout, int Ho, int Wo, const float* __restrict__ Wt, int Cin,
                                                         float* __restrict__ dX, int Hin, int Win, int accumulate) {
    constexpr int CO_C = 8, PAD = (K - 1) / 2;
    constexpr int DH = (TH + K - 2) / S + 2, DW = (TW + K - 2) / S + 2;
    __shared__ float s_dy[CO_C][DH][DW + 1];
    __shared__ __align__(16) float s_w[CO_C][K * K][CI_T];
    const int tiles_x = (Win + TW - 1) / TW;
    const int iy0 = (blockIdx.x / tiles_x) * TH, ix0 = (blockIdx.x % tiles_x) * TW;
    const int ci0 = blockIdx.y * CI_T, n = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int iy = iy0 + ty, ix = ix0 + tx;
    // first output row/col that any pixel of this tile can touch (floor division, arguments may be negative)
    const int oyb = (iy0 + PAD - (K - 1) + 4 * S) / S - 4, oxb = (ix0 + PAD - (K - 1) + 4 * S) / S - 4;
    float acc[CI_T];
#pragma unroll
    for (int i = 0; i < CI_T; ++i) acc[i] = 0.f;
    for (int co0 = 0; co0 < Cout; co0 += CO_C) {
        for (int idx = threadIdx.x; idx < CO_C * DH * DW; idx += 256) {
            const int co = idx / (DH * DW), r = (idx / DW) % DH, c = idx % DW;
            const int oy = oyb + r, ox = oxb + c;
            float v = 0.f;
            if (co0 + co < Cout && oy >= 0 && oy < Ho && ox >= 0 && ox < Wo)
                v = dy_value(src, co0 + co, ((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox);
            s_dy[co][r][c] = v;
        }
        for (int idx = threadIdx.x; idx < CO_C * K * K * CI_T; idx += 256) {
            const int ci = idx % CI_T, kk = (idx / CI_T) % (K * K), co = idx / (CI_T * K * K);
            s_w[co][kk][ci] = (co0 + co < Cout && ci0 + ci < Cin) ? Wt[((size_t)(co0 + co) * Cin + ci0 + ci) * K * K + kk] : 0.f;
        }
        __syncthreads();
        const int cmax = (Cout - co0 < CO_C) ? Cout - co0 : CO_C;
        for (int co = 0; co < cmax; ++co) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const int t = iy + PAD - ky + 4 * S;          // oy * S = iy + PAD - ky
                if (S > 1 && (t % S) != 0) continue;           // warp-uniform (iy is)
                const int r = t / S - 4 - oyb;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const int u = ix + PAD - kx + 4 * S;
                    const bool ok = (S == 1) || (u % S) == 0;
                    const int c = u / S - 4 - oxb;
                    const float v = ok ? s_dy[co][r][c] : 0.f;
#pragma unroll
                    for (int ci = 0; ci < CI_T; ++ci) acc[ci] = fmaf(v, s_w[co][ky * K + kx][ci], acc[ci]);
                }
            }
        }
        __syncthreads();
    }
    if (iy < Hin && ix < Win) {
#pragma unroll
        for (int ci = 0; ci < CI_T; ++ci)
            if (ci0 + ci < Cin) {
                float* o = dX + ((size_t)(n * Cin + ci0 + ci) * Hin + iy) * Win + ix;
                *o = accumulate ? *o + acc[ci] : acc[ci];
            }
    }
}

// partial weight / bias gradients over a strided subset of the spatial tiles.
// part layout: [gridDim.x][Cout*Cin*K*K + Cout]  (bias gradient partials at the end)
template <int K, int S>
constexpr int wgrad_smem_floats() {
    constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    constexpr int a = 8 * IH * (IW + 1) + 16 * TH * (TW + 1), b = 8 * 32 * (4 * K * K + 4);
    return a > b ? a : b;
}
template <int K, int S>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, int Cin, int Hin, int Win, InTf tf, DySrc src,
                                                         int Cout, int Ho, int Wo, int N, float* __restrict__ part) {
    constexpr int CO_T = 16, CI_T = 8, PAD = (K - 1) / 2, KK = K * K;
    constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K, IWP = IW + 1;
    constexpr int SX = CI_T * IH * IWP;
    extern __shared__ float smem[];   // wgrad_smem_floats<K, S>() floats (K=3, S=2 needs 52.8 KB: opt-in dynamic)
    float (*s_x)[IH][IWP] = reinterpret_cast<float (*)[IH][IWP]>(smem);
    float (*s_dy)[TH][TW + 1] = reinterpret_cast<float (*)[TH][TW + 1]>(smem + SX);
    const int co0 = (blockIdx.y / ((Cin + CI_T - 1) / CI_T)) * CO_T;
    const int ci0 = (blockIdx.y % ((Cin + CI_T - 1) / CI_T)) * CI_T;
    const int g = threadIdx.x >> 5, w = threadIdx.x & 31;   // pixel group, weight thread
    const int cos = (w / CI_T) * 4, ci = w % CI_T;          // this thread: 4 output channels x 1 input channel x KK taps
    const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
    const int ntiles = N * tiles_y * tiles_x;
    float acc[4][KK];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) acc[j][kk] = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / (tiles_y * tiles_x), tr = tile % (tiles_y * tiles_x);
        const int ty0 = (tr / tiles_x) * TH, tx0 = (tr % tiles_x) * TW;
        for (int idx = threadIdx.x; idx < CI_T * IH * IW; idx += 256) {
            const int c = idx / (IH * IW), r = (idx / IW) % IH, q = idx % IW;
            const int iy = ty0 * S - PAD + r, ix = tx0 * S - PAD + q;
            float v = 0.f;
            if (ci0 + c < Cin && iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
                v = apply_tf(tf, ci0 + c, x[((size_t)(n * Cin + ci0 + c) * Hin + iy) * Win + ix]);
            s_x[c][r][q] = v;
        }
        for (int idx = threadIdx.x; idx < CO_T * TH * TW; idx += 256) {
            const int c = idx / (TH * TW), r = (idx / TW) % TH, q = idx % TW;
            const int oy = ty0 + r, ox = tx0 + q;
            float v = 0.f;
            if (co0 + c < Cout && oy < Ho && ox < Wo) v = dy_value(src, co0 + c, ((size_t)(n * Cout + co0 + c) * Ho + oy) * Wo + ox);
            s_dy[c][r][q] = v;
        }
        __syncthreads();
        for (int p = g; p < TH * TW; p += 8) {
            const int py = p / TW, px = p % TW;
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = s_dy[cos + j][py][px];
#pragma unroll
            for (int j = 0; j < 4; ++j) bacc[j] += d[j];
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float xv = s_x[ci][py * S + ky][px * S + kx];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][ky * K + kx] = fmaf(d[j], xv, acc[j][ky * K + kx]);
                }
        }
        __syncthreads();
    }
    // reduce over the 8 pixel groups through shared memory, then one writer per weight
    float* red = smem;   // [8][32][4*KK + 4]
    constexpr int RS = 4 * KK + 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) red[(g * 32 + w) * RS + j * KK + kk] = acc[j][kk];
        red[(g * 32 + w) * RS + 4 * KK + j] = bacc[j];
    }
    __syncthreads();
    const size_t nW = (size_t)Cout * Cin * KK;
    float* out = part + (size_t)blockIdx.x * (nW + Cout);
    for (int idx = threadIdx.x; idx < 32 * RS; idx += 256) {
        const int ww = idx / RS, e = idx % RS;
        float s = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) s += red[(gg * 32 + ww) * RS + e];
        const int wcos = (ww / CI_T) * 4, wci = ww % CI_T;
        if (e < 4 * KK) {
            const int j = e / KK, kk = e % KK;
            const int co = co0 + wcos + j, c = ci0 + wci;
            if (co < Cout && c < Cin) out[((size_t)co * Cin + c) * KK + kk] = s;
        } else if (ci0 == 0 && wci == 0) {
            const int co = co0 + wcos + (e - 4 * KK);
            if (co < Cout) out[nW + co] = s;
        }
    }
}
// grad_w += sum over chunks, grad_b += sum over chunks (fixed order)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int nchunks, size_t nW, int Cout,
                                                           float* __restrict__ gw, float* __restrict__ gb) {
    const size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    const size_t tot = nW + Cout;
    if (i >= tot) return;
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * tot + i];
    if (i < nW) gw[i] += s;
    else gb[i - nW] += s;
}

// adjoint of cat_build for the skip branch: d(s activated) = d(cat)[:, :Cs] placed at the crop offset, zero elsewhere
__global__ void __launch_bounds__(256) cat_bwd_skip_kernel(DySrc src, int C, int H, int W, int Cs, int Hs, int Ws, int offy,
                                                           int offx, float* __restrict__ dS) {
    const size_t total = (size_t)gridDim.z * Cs * Hs * Ws;
    const int n = blockIdx.z;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < (size_t)Cs * Hs * Ws; i += (size_t)gridDim.x * 256) {
        const int c = i / ((size_t)Hs * Ws), y = (i / Ws) % Hs, x = i % Ws;
        const int cy = y - offy, cx = x - offx;
        float v = 0.f;
        if (cy >= 0 && cy < H && cx >= 0 && cx < W) v = dy_value(src, c, ((size_t)(n * C + c) * H + cy) * W + cx);
        dS[(size_t)n * Cs * Hs * Ws + i] = v;
    }
    (void)total;
}
// adjoint of the bilinear x2 up-sampling (+ crop): d(u activated)[n,cu,yu,xu] = sum over the <= 4x4 fine pixels that read it
__global__ void __launch_bounds__(256) cat_bwd_up_kernel(DySrc src, int C, int H, int W, int Cs, int Cu, int Hu, int Wu, int offy,
                                                         int offx, float* __restrict__ dU) {
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
                acc += wy * wx * dy_value(src, Cs + cu, ((size_t)(n * C + Cs + cu) * H + cy) * W + cx);
            }
        }
        dU[(size_t)n * Cu * Hu * Wu + i] = acc;
    }
}

// -------------------------------------------------------------------------------------------------
// host: launch helpers
// -------------------------------------------------------------------------------------------------
static int launch_conv_fwd(int K, int S, const float* x, int N, int Cin, int Hin, int Win, InTf tf, const float* Wt,
                           const float* bias, int Cout, float* y, int Ho, int Wo, int sigmoid, float* stats, cudaStream_t st) {
    const int tiles = ceil_div(Ho, TH) * ceil_div(Wo, TW);
    const bool small = Cout <= 4;
    dim3 grid(tiles, ceil_div(Cout, small ? 4 : 16), N);
#define CF(KK, SS, CT) conv_fwd_kernel<KK, SS, CT><<<grid, 256, 0, st>>>(x, Cin, Hin, Win, tf, Wt, bias, Cout, y, Ho, Wo, sigmoid, stats)
    if (K == 1 && S == 1) { if (small) CF(1, 1, 4); else CF(1, 1, 16); }
    else if (K == 3 && S == 1) { if (small) CF(3, 1, 4); else CF(3, 1, 16); }
    else if (K == 3 && S == 2) { if (small) CF(3, 2, 4); else CF(3, 2, 16); }
    else { set_error("generator: unsupported conv k=%d stride=%d", K, S); return SPLICE_ERR_UNSUPPORTED; }
#undef CF
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
static int launch_conv_dgrad(int K, int S, DySrc src, int N, int Cout, int Ho, int Wo, const float* Wt, int Cin, float* dX, int Hin,
                             int Win, int accumulate, cudaStream_t st) {
    const int tiles = ceil_div(Hin, TH) * ceil_div(Win, TW);
    const bool small = Cin <= 4;
    dim3 grid(tiles, ceil_div(Cin, small ? 4 : 16), N);
#define DG(KK, SS, CT) conv_dgrad_kernel<KK, SS, CT><<<grid, 256, 0, st>>>(src, Cout, Ho, Wo, Wt, Cin, dX, Hin, Win, accumulate)
    if (K == 1 && S == 1) { if (small) DG(1, 1, 4); else DG(1, 1, 16); }
    else if (K == 3 && S == 1) { if (small) DG(3, 1, 4); else DG(3, 1, 16); }
    else if (K == 3 && S == 2) { if (small) DG(3, 2, 4); else DG(3, 2, 16); }
    else { set_error("generator: unsupported conv k=%d stride=%d", K, S); return SPLICE_ERR_UNSUPPORTED; }
#undef DG
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

GenEngine::GenEngine() {
    const int cin[GEN_SCALES] = {3, 16, 32, 64, 128};
    const int cd[GEN_SCALES] = {16, 32, 64, 128, 128};
    const int cu[GEN_SCALES] = {16, 32, 64, 128, 128};
    for (int i = 0; i < GEN_SCALES; ++i) {
        Scale& s = sc_[i];
        const int pre = 12 * i, post = 60 + 10 * (GEN_SCALES - 1 - i);
        const int bpre = 3 * i, bpost = 15 + 3 * (GEN_SCALES - 1 - i);
        s.cdeep = (i == GEN_SCALES - 1) ? cd[i] : cu[i + 1];
        s.s = Conv{cin[i], 4, 1, 1, pre + 0, pre + 1};
        s.bs = Bn{4, pre + 2, pre + 3, bpre + 0};
        s.d1 = Conv{cin[i], cd[i], 3, 2, pre + 4, pre + 5};
        s.bd1 = Bn{cd[i], pre + 6, pre + 7, bpre + 1};
        s.d2 = Conv{cd[i], cd[i], 3, 1, pre + 8, pre + 9};
        s.bd2 = Bn{cd[i], pre + 10, pre + 11, bpre + 2};
        s.bcat = Bn{4 + s.cdeep, post + 0, post + 1, bpost + 0};
        s.c1 = Conv{4 + s.cdeep, cu[i], 3, 1, post + 2, post + 3};
        s.bc1 = Bn{cu[i], post + 4, post + 5, bpost + 1};
        s.c2 = Conv{cu[i], cu[i], 1, 1, post + 6, post + 7};
        s.bc2 = Bn{cu[i], post + 8, post + 9, bpost + 2};
    }
    final_ = Conv{16, 3, 1, 1, 110, 111};
}

GenEngine::~GenEngine() {
    for (auto& s : slots_) cudaFree(s.pool);
    cudaFree(scratch_);
}

int GenEngine::ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes_) return SPLICE_OK;
    SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(scratch_);
    scratch_ = nullptr;
    scratch_bytes_ = 0;
    SPLICE_CHECK_CUDA(cudaMalloc(&scratch_, bytes));
    scratch_bytes_ = bytes;
    return SPLICE_OK;
}

int GenEngine::configure(Slot& s, int N, int H, int W) {
    if (s.pool && s.N == N && s.H == H && s.W == W) return SPLICE_OK;
    size_t off = 0;
    std::vector<size_t> offs;
    auto plan = [&](size_t bytes) { offs.push_back(off); off += (bytes + 255) & ~(size_t)255; };
    int h = H, w = W;
    int hs[GEN_SCALES], ws[GEN_SCALES], hd[GEN_SCALES], wd[GEN_SCALES];
    for (int i = 0; i < GEN_SCALES; ++i) {
        hs[i] = h; ws[i] = w; hd[i] = (h + 1) / 2; wd[i] = (w + 1) / 2;
        h = hd[i]; w = wd[i];
    }
    SPLICE_REQUIRE(hs[GEN_SCALES - 1] >= 1 && ws[GEN_SCALES - 1] >= 1, "generator: input %dx%d too small", H, W);
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        const size_t px = (size_t)N * hs[i] * ws[i], pd = (size_t)N * hd[i] * wd[i];
        const size_t sz[6] = {px * 4, pd * c.d1.cout, pd * c.d2.cout, px * (4 + c.cdeep), px * c.c1.cout, px * c.c2.cout};
        for (int k = 0; k < 6; ++k) plan(sz[k] * 4);   // raw tensors
        for (int k = 0; k < 6; ++k) plan(sz[k] * 4);   // gradients w.r.t. the activated / normalised tensors
        const int ch[6] = {4, c.d1.cout, c.d2.cout, 4 + c.cdeep, c.c1.cout, c.c2.cout};
        for (int k = 0; k < 6; ++k) plan((size_t)ch[k] * sizeof(float4));
        for (int k = 0; k < 6; ++k) plan((size_t)ch[k] * sizeof(float2));
    }
    plan((size_t)N * 3 * H * W * 4);  // x copy
    plan((size_t)N * 3 * H * W * 4);  // out copy
    if (off > s.pool_bytes) {
        SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(s.pool);
        s.pool = nullptr;
        s.pool_bytes = 0;
        SPLICE_CHECK_CUDA(cudaMalloc(&s.pool, off));
        s.pool_bytes = off;
    }
    uint8_t* base = static_cast<uint8_t*>(s.pool);
    size_t k = 0;
    auto nx = [&]() { return base + offs[k++]; };
    for (int i = 0; i < GEN_SCALES; ++i) {
        ScaleBuf& b = s.sb[i];
        b.h = hs[i]; b.w = ws[i]; b.hd = hd[i]; b.wd = wd[i];
        b.s_raw = (float*)nx(); b.d1_raw = (float*)nx(); b.d2_raw = (float*)nx(); b.cat = (float*)nx(); b.c1_raw = (float*)nx(); b.c2_raw = (float*)nx();
        b.dA_s = (float*)nx(); b.dA_d1 = (float*)nx(); b.dA_d2 = (float*)nx(); b.dcat = (float*)nx(); b.dA_c1 = (float*)nx(); b.dA_c2 = (float*)nx();
        b.k_s = (float4*)nx(); b.k_d1 = (float4*)nx(); b.k_d2 = (float4*)nx(); b.k_cat = (float4*)nx(); b.k_c1 = (float4*)nx(); b.k_c2 = (float4*)nx();
        b.m_s = (float2*)nx(); b.m_d1 = (float2*)nx(); b.m_d2 = (float2*)nx(); b.m_cat = (float2*)nx(); b.m_c1 = (float2*)nx(); b.m_c2 = (float2*)nx();
    }
    s.x_copy = (float*)nx();
    s.out = (float*)nx();
    s.N = N; s.H = H; s.W = W;
    s.valid = false;
    return SPLICE_OK;
}

#define GRC(expr)              \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

int GenEngine::forward(const GenPointers& p, const float* x, int N, int H, int W, float* out, int slot, bool keep,
                       bool update_running, cudaStream_t st) {
    SPLICE_REQUIRE(slot >= 0 && slot < GEN_SLOTS, "generator: slot out of range");
    SPLICE_REQUIRE(x && out && N > 0 && H > 0 && W > 0, "generator: bad input");
    Slot& s = slots_[slot];
    GRC(configure(s, N, H, W));
    // scratch: statistics partials of the largest layer: N * tiles(H,W) * Cmax(132) * 3 floats; wgrad partials (backward)
    const size_t tiles0 = (size_t)ceil_div(H, TH) * ceil_div(W, TW);
    const size_t wg = (size_t)32 * (132 * 128 * 9 + 128);
    GRC(ensure_scratch((N * tiles0 * 132 * 3 + wg) * sizeof(float) + 4096));
    float* part = static_cast<float*>(scratch_);
    const float eps = 1e-5f, mom = 0.1f;

    auto bn_fin = [&](const Bn& b, int nparts, float4* k) -> int {
        bn_finalize_kernel<<<b.c, 32, 0, st>>>(part, nparts, b.c, p.param[b.pg], p.param[b.pb], eps, k,
                                               update_running ? p.running_mean[b.idx] : nullptr,
                                               update_running ? p.running_var[b.idx] : nullptr,
                                               update_running ? p.num_batches_tracked[b.idx] : nullptr, mom);
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };
    auto conv_bn = [&](const Conv& c, const Bn& b, const float* in, int hin, int win, InTf tf, float* y, int ho, int wo,
                       float4* k) -> int {
        GRC(launch_conv_fwd(c.k, c.stride, in, N, c.cin, hin, win, tf, p.param[c.pw], p.param[c.pb], c.cout, y, ho, wo, 0, part, st));
        return bn_fin(b, N * ceil_div(ho, TH) * ceil_div(wo, TW), k);
    };

    // keep a private copy of the input: the caller's tensor may be freed before backward() (wgrad of scale 0 reads it)
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(s.x_copy, x, (size_t)N * 3 * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s.x = s.x_copy;

    // down path
    const float* in = s.x;
    InTf tf_in{nullptr, 0};
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        GRC(conv_bn(c.s, c.bs, in, b.h, b.w, tf_in, b.s_raw, b.h, b.w, b.k_s));
        GRC(conv_bn(c.d1, c.bd1, in, b.h, b.w, tf_in, b.d1_raw, b.hd, b.wd, b.k_d1));
        GRC(conv_bn(c.d2, c.bd2, b.d1_raw, b.hd, b.wd, InTf{b.k_d1, 1}, b.d2_raw, b.hd, b.wd, b.k_d2));
        in = b.d2_raw;
        tf_in = InTf{b.k_d2, 1};
    }
    // up path
    for (int i = GEN_SCALES - 1; i >= 0; --i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        const float* u = (i == GEN_SCALES - 1) ? b.d2_raw : s.sb[i + 1].c2_raw;
        InTf tf_u = (i == GEN_SCALES - 1) ? InTf{b.k_d2, 1} : InTf{s.sb[i + 1].k_c2, 1};
        const int hu = b.hd, wu = b.wd;                                      // == the deeper scale's size
        const int th = min(b.h, 2 * hu), tw = min(b.w, 2 * wu);            // Concat crops to the smaller size
        SPLICE_REQUIRE(th == b.h && tw == b.w, "generator: unexpected concat geometry");
        const int oys = (b.h - th) / 2, oxs = (b.w - tw) / 2, oyu = (2 * hu - th) / 2, oxu = (2 * wu - tw) / 2;
        const int C = 4 + c.cdeep;
        dim3 grid(ceil_div(th, TH) * ceil_div(tw, TW), C, N);
        cat_build_kernel<<<grid, 256, 0, st>>>(b.s_raw, 4, b.h, b.w, InTf{b.k_s, 1}, oys, oxs, u, c.cdeep, hu, wu, tf_u, oyu, oxu,
                                               b.cat, th, tw, part);
        SPLICE_LAUNCH_CHECK();
        GRC(bn_fin(c.bcat, N * ceil_div(th, TH) * ceil_div(tw, TW), b.k_cat));
        GRC(conv_bn(c.c1, c.bc1, b.cat, th, tw, InTf{b.k_cat, 0}, b.c1_raw, th, tw, b.k_c1));
        GRC(conv_bn(c.c2, c.bc2, b.c1_raw, th, tw, InTf{b.k_c1, 1}, b.c2_raw, th, tw, b.k_c2));
    }
    GRC(launch_conv_fwd(1, 1, s.sb[0].c2_raw, N, final_.cin, H, W, InTf{s.sb[0].k_c2, 1}, p.param[final_.pw], p.param[final_.pb], 3,
                        s.out, H, W, 1, nullptr, st));
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(out, s.out, (size_t)N * 3 * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s.valid = keep;
    return SPLICE_OK;
}

int GenEngine::backward(const GenPointers& p, const float* dout, int slot, cudaStream_t st) {
    SPLICE_REQUIRE(slot >= 0 && slot < GEN_SLOTS, "generator: slot out of range");
    Slot& s = slots_[slot];
    SPLICE_REQUIRE(s.pool && s.valid, "generator backward: slot %d holds no kept forward pass", slot);
    SPLICE_REQUIRE(dout, "generator backward: null gradient");
    const int N = s.N, H = s.H, W = s.W;
    const size_t tiles0 = (size_t)ceil_div(H, TH) * ceil_div(W, TW);
    float* part = static_cast<float*>(scratch_);
    float* wpart = part + N * tiles0 * 132 * 3;

    auto bn_bwd = [&](const Bn& b, const float* dA, const float* y, const float4* k, int lrelu, int hw_h, int hw_w, float2* m) -> int {
        const int HW = hw_h * hw_w;
        dim3 grid(ceil_div(HW, 2048), b.c, N);
        bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dA, y, k, lrelu, b.c, HW, part);
        SPLICE_LAUNCH_CHECK();
        bn_bwd_finalize_kernel<<<b.c, 32, 0, st>>>(part, N * (int)grid.x, b.c, (double)N * HW, p.grad[b.pg], p.grad[b.pb], m);
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };
    auto wgrad = [&](const Conv& c, const float* in, int hin, int win, InTf tf, DySrc src, int ho, int wo) -> int {
        const int ntiles = N * ceil_div(ho, TH) * ceil_div(wo, TW);
        const int chunks = ntiles < 32 ? ntiles : 32;
        dim3 grid(chunks, ceil_div(c.cout, 16) * ceil_div(c.cin, 8));
        static bool attr = false;
        if (!attr) {
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<1, 1>() * 4));
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<3, 1>() * 4));
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<3, 2>() * 4));
            attr = true;
        }
        if (c.k == 1 && c.stride == 1)
            conv_wgrad_kernel<1, 1><<<grid, 256, wgrad_smem_floats<1, 1>() * 4, st>>>(in, c.cin, hin, win, tf, src, c.cout, ho, wo, N, wpart);
        else if (c.k == 3 && c.stride == 1)
            conv_wgrad_kernel<3, 1><<<grid, 256, wgrad_smem_floats<3, 1>() * 4, st>>>(in, c.cin, hin, win, tf, src, c.cout, ho, wo, N, wpart);
        else
            conv_wgrad_kernel<3, 2><<<grid, 256, wgrad_smem_floats<3, 2>() * 4, st>>>(in, c.cin, hin, win, tf, src, c.cout, ho, wo, N, wpart);
        SPLICE_LAUNCH_CHECK();
        const size_t nW = (size_t)c.cout * c.cin * c.k * c.k;
        wgrad_reduce_kernel<<<ceil_div((int)(nW + c.cout), 256), 256, 0, st>>>(wpart, chunks, nW, c.cout, p.grad[c.pw], p.grad[c.pb]);
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };

    // final 1x1 conv + sigmoid
    {
        DySrc src{dout, s.out, nullptr, nullptr, DY_SIGMOID};
        GRC(wgrad(final_, s.sb[0].c2_raw, H, W, InTf{s.sb[0].k_c2, 1}, src, H, W));
        GRC(launch_conv_dgrad(1, 1, src, N, 3, H, W, p.param[final_.pw], final_.cin, s.sb[0].dA_c2, H, W, 0, st));
    }
    // up path, top to bottom
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        const int h = b.h, w = b.w;
        GRC(bn_bwd(c.bc2, b.dA_c2, b.c2_raw, b.k_c2, 1, h, w, b.m_c2));
        DySrc s_c2{b.dA_c2, b.c2_raw, b.k_c2, b.m_c2, DY_BN_LRELU};
        GRC(wgrad(c.c2, b.c1_raw, h, w, InTf{b.k_c1, 1}, s_c2, h, w));
        GRC(launch_conv_dgrad(1, 1, s_c2, N, c.c2.cout, h, w, p.param[c.c2.pw], c.c2.cin, b.dA_c1, h, w, 0, st));

        GRC(bn_bwd(c.bc1, b.dA_c1, b.c1_raw, b.k_c1, 1, h, w, b.m_c1));
        DySrc s_c1{b.dA_c1, b.c1_raw, b.k_c1, b.m_c1, DY_BN_LRELU};
        GRC(wgrad(c.c1, b.cat, h, w, InTf{b.k_cat, 0}, s_c1, h, w));
        GRC(launch_conv_dgrad(3, 1, s_c1, N, c.c1.cout, h, w, p.param[c.c1.pw], c.c1.cin, b.dcat, h, w, 0, st));

        GRC(bn_bwd(c.bcat, b.dcat, b.cat, b.k_cat, 0, h, w, b.m_cat));
        DySrc s_cat{b.dcat, b.cat, b.k_cat, b.m_cat, DY_BN};
        const int C = 4 + c.cdeep, hu = b.hd, wu = b.wd;
        const int oyu = (2 * hu - h) / 2, oxu = (2 * wu - w) / 2;
        {
            dim3 grid(min(ceil_div(4 * h * w, 256), 148 * 8), 1, N);
            cat_bwd_skip_kernel<<<grid, 256, 0, st>>>(s_cat, C, h, w, 4, h, w, 0, 0, b.dA_s);
            SPLICE_LAUNCH_CHECK();
            float* dU = (i == GEN_SCALES - 1) ? b.dA_d2 : s.sb[i + 1].dA_c2;
            dim3 grid2(min(ceil_div(c.cdeep * hu * wu, 256), 148 * 8), 1, N);
            cat_bwd_up_kernel<<<grid2, 256, 0, st>>>(s_cat, C, h, w, 4, c.cdeep, hu, wu, oyu, oxu, dU);
            SPLICE_LAUNCH_CHECK();
        }
    }
    // down path, bottom to top
    for (int i = GEN_SCALES - 1; i >= 0; --i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        const float* in = (i == 0) ? s.x : s.sb[i - 1].d2_raw;
        InTf tf_in = (i == 0) ? InTf{nullptr, 0} : InTf{s.sb[i - 1].k_d2, 1};
        float* dIn = (i == 0) ? nullptr : s.sb[i - 1].dA_d2;

        GRC(bn_bwd(c.bs, b.dA_s, b.s_raw, b.k_s, 1, b.h, b.w, b.m_s));
        DySrc s_s{b.dA_s, b.s_raw, b.k_s, b.m_s, DY_BN_LRELU};
        GRC(wgrad(c.s, in, b.h, b.w, tf_in, s_s, b.h, b.w));
        if (dIn) GRC(launch_conv_dgrad(1, 1, s_s, N, 4, b.h, b.w, p.param[c.s.pw], c.s.cin, dIn, b.h, b.w, 0, st));

        GRC(bn_bwd(c.bd2, b.dA_d2, b.d2_raw, b.k_d2, 1, b.hd, b.wd, b.m_d2));
        DySrc s_d2{b.dA_d2, b.d2_raw, b.k_d2, b.m_d2, DY_BN_LRELU};
        GRC(wgrad(c.d2, b.d1_raw, b.hd, b.wd, InTf{b.k_d1, 1}, s_d2, b.hd, b.wd));
        GRC(launch_conv_dgrad(3, 1, s_d2, N, c.d2.cout, b.hd, b.wd, p.param[c.d2.pw], c.d2.cin, b.dA_d1, b.hd, b.wd, 0, st));

        GRC(bn_bwd(c.bd1, b.dA_d1, b.d1_raw, b.k_d1, 1, b.hd, b.wd, b.m_d1));
        DySrc s_d1{b.dA_d1, b.d1_raw, b.k_d1, b.m_d1, DY_BN_LRELU};
        GRC(wgrad(c.d1, in, b.h, b.w, tf_in, s_d1, b.hd, b.wd));
        if (dIn) GRC(launch_conv_dgrad(3, 2, s_d1, N, c.d1.cout, b.hd, b.wd, p.param[c.d1.pw], c.d1.cin, dIn, b.h, b.w, 1, st));
    }
    s.valid = false;
    return SPLICE_OK;
}

}  // namespace splice
