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

#include "conv_tc.h"
#include "gen_dev.cuh"
#include "gen_kernels.cuh"

namespace splice {


// -------------------------------------------------------------------------------------------------
// forward convolution (+ producer BN/LeakyReLU on load, + bias, + optional sigmoid, + output statistics)
//   one thread per output pixel (linear over N*Ho*Wo), CO_T output channels per thread, optional split over the
//   input channels (blockIdx.z) for the low-resolution layers whose grids would otherwise not fill 148 SMs.
//   Input taps come straight from global/L1 (neighbouring threads share them); weights are staged in smem.
// -------------------------------------------------------------------------------------------------
static constexpr int CONV_THREADS = 128;

template <int K, int S, int CO_T>
__global__ void __launch_bounds__(CONV_THREADS) conv_fwd_kernel(const float* __restrict__ x, int N, int Cin, int Hin, int Win, InTf tf,
                                                                const float* __restrict__ Wt, const float* __restrict__ bias,
                                                                int Cout, float* __restrict__ y, int Ho, int Wo, int out_sigmoid,
                                                                float* __restrict__ stats_part, int splitK, BnFin fin) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CI_C = 8, KK = K * K, PAD = (K - 1) / 2;
    __shared__ __align__(16) float s_w[CI_C][KK][CO_T];
    __shared__ float2 s_ab[CI_C];
    __shared__ float red[4 * CO_T];
    __shared__ int s_flag;
    const int P = N * Ho * Wo;
    const int p = blockIdx.x * CONV_THREADS + threadIdx.x;
    const bool active = p < P;
    const int n = active ? p / (Ho * Wo) : 0, rem = active ? p % (Ho * Wo) : 0;
    const int oy = rem / Wo, ox = rem % Wo;
    const int co0 = blockIdx.y * CO_T;
    const int cps = (Cin + splitK - 1) / splitK;
    const int c_begin = blockIdx.z * cps, c_end = min(Cin, c_begin + cps);
    const int iy0 = oy * S - PAD, ix0 = ox * S - PAD;
    // taps are loaded unconditionally from clamped coordinates (so the 9 loads of a channel issue back to back and the
    // next channel's taps are prefetched while this one's FMAs run); out-of-image taps are zeroed by mask afterwards
    int yoff[K], xoff[K];
    bool rok[K], cok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        rok[k] = active && iy0 + k >= 0 && iy0 + k < Hin;
        cok[k] = ix0 + k >= 0 && ix0 + k < Win;
        yoff[k] = min(max(iy0 + k, 0), Hin - 1) * Win;
        xoff[k] = min(max(ix0 + k, 0), Win - 1);
    }
    float acc[CO_T];
#pragma unroll
    for (int i = 0; i < CO_T; ++i) acc[i] = 0.f;
    const size_t plane = (size_t)Hin * Win;
    const float* xn = x + (size_t)n * Cin * plane;
    float cur[KK], nxt[KK];
    auto load_taps = [&](float (&v)[KK], int c) {
        const float* base = xn + (size_t)c * plane;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) v[ky * K + kx] = __ldg(base + yoff[ky] + xoff[kx]);
    };
    if (c_begin < c_end) load_taps(nxt, c_begin);

    for (int c0 = c_begin; c0 < c_end; c0 += CI_C) {
        const int cn = min(CI_C, c_end - c0);
        for (int idx = threadIdx.x; idx < CI_C * KK * CO_T; idx += CONV_THREADS) {
            const int co = idx % CO_T, kk = (idx / CO_T) % KK, ci = idx / (CO_T * KK);
            s_w[ci][kk][co] = (ci < cn && co0 + co < Cout) ? Wt[((size_t)(co0 + co) * Cin + c0 + ci) * KK + kk] : 0.f;
        }
        if (threadIdx.x < CI_C) {
            float2 ab = make_float2(1.f, 0.f);
            if (tf.k && threadIdx.x < cn) { const float4 k4 = tf.k[c0 + threadIdx.x]; ab = make_float2(k4.z, k4.w); }
            s_ab[threadIdx.x] = ab;
        }
        __syncthreads();
        for (int ci = 0; ci < cn; ++ci) {
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) cur[kk] = nxt[kk];
            if (c0 + ci + 1 < c_end) load_taps(nxt, c0 + ci + 1);
            const float2 ab = s_ab[ci];
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    float v = cur[ky * K + kx];
                    if (tf.k) {
                        v = fmaf(ab.x, v, ab.y);
                        if (tf.lrelu) v = v < 0.f ? v * LRELU : v;
                    }
                    v = (rok[ky] && cok[kx]) ? v : 0.f;   // zero padding lives in the post-BN/activation domain
#pragma unroll
                    for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(v, s_w[ci][ky * K + kx][co], acc[co]);
                }
            }
        }
        __syncthreads();
    }
    if (splitK > 1) {   // partial sums; conv_finish_bn_kernel adds the bias and does the statistics
        if (active) {
            float* o = y + (size_t)blockIdx.z * ((size_t)N * Cout * Ho * Wo);
#pragma unroll
            for (int co = 0; co < CO_T; ++co)
                if (co0 + co < Cout) o[((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox] = acc[co];
        }
        return;
    }
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
        if (co0 + co < Cout) {
            float v = acc[co] + bias[co0 + co];
            if (out_sigmoid) v = 1.f / (1.f + __expf(-v));
            acc[co] = v;
            if (active) y[((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox] = v;
        }
    }
    if (stats_part) {
        // per-block (count, mean, M2) of each output channel: two block reductions, centred second pass
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const float cnt = (float)min(CONV_THREADS, P - (int)blockIdx.x * CONV_THREADS);
        float mean[CO_T], m2[CO_T];
#pragma unroll
        for (int co = 0; co < CO_T; ++co) {
            const float sv = warp_sum(active ? acc[co] : 0.f);
            if (lane == 0) red[w * CO_T + co] = sv;
        }
        __syncthreads();
#pragma unroll
        for (int co = 0; co < CO_T; ++co) mean[co] = (red[co] + red[CO_T + co] + red[2 * CO_T + co] + red[3 * CO_T + co]) / cnt;
        __syncthreads();
#pragma unroll
        for (int co = 0; co < CO_T; ++co) {
            const float d = acc[co] - mean[co];
            const float sv = warp_sum(active ? d * d : 0.f);
            if (lane == 0) red[w * CO_T + co] = sv;
        }
        __syncthreads();
#pragma unroll
        for (int co = 0; co < CO_T; ++co) m2[co] = red[co] + red[CO_T + co] + red[2 * CO_T + co] + red[3 * CO_T + co];
        if (threadIdx.x == 0) {
#pragma unroll
            for (int co = 0; co < CO_T; ++co)
                if (co0 + co < Cout) {
                    float* o = stats_part + ((size_t)blockIdx.x * Cout + co0 + co) * 3;
                    o[0] = cnt; o[1] = mean[co]; o[2] = m2[co];
                }
        }
        if (fin.konst) bn_finish_if_last(stats_part, gridDim.x, Cout, co0, CO_T, blockIdx.y, gridDim.x, fin, &s_flag);
    }
}

// nn.BatchNorm2d's running statistics (momentum 0.1, unbiased variance, num_batches_tracked) from the batch statistics
// a forward pass left in its slot. A separate launch so that netG calls issued on parallel streams can apply their
// updates afterwards in call order, as the reference's sequential calls do (they never influence outputs: the
// reference never calls .eval(); kept for state_dict() fidelity). One block per BatchNorm layer.
struct RunningTable {
    const float2* bstat[GEN_BN];
    float* rmean[GEN_BN];
    float* rvar[GEN_BN];
    long long* nbt[GEN_BN];
    int C[GEN_BN];
};
__global__ void __launch_bounds__(160) update_running_kernel(RunningTable t, float momentum) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const int l = blockIdx.x, c = threadIdx.x;
    if (c < t.C[l]) {
        const float2 b = t.bstat[l][c];
        t.rmean[l][c] = (1.f - momentum) * t.rmean[l][c] + momentum * b.x;
        t.rvar[l][c] = (1.f - momentum) * t.rvar[l][c] + momentum * b.y;
    }
    if (c == 0) *t.nbt[l] += 1;
}

// d(transformed conv input) [N,Cin,Hin,Win] from dy [N,Cout,Ho,Wo]: one thread per input pixel, CI_T input channels
// per thread, optional split over the output channels (blockIdx.z) for the low-resolution layers
template <int K, int S, int CI_T>
__global__ void __launch_bounds__(CONV_THREADS) conv_dgrad_kernel(const float* __restrict__ dy, int N, int Cout, int Ho, int Wo,
                                                                  const float* __restrict__ Wt, int Cin, float* __restrict__ dX,
                                                                  int Hin, int Win, int accumulate, int splitK) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CO_C = 8, KK = K * K, PAD = (K - 1) / 2;
    __shared__ __align__(16) float s_w[CO_C][KK][CI_T];
    const int P = N * Hin * Win;
    const int p = blockIdx.x * CONV_THREADS + threadIdx.x;
    const bool active = p < P;
    const int n = active ? p / (Hin * Win) : 0, rem = active ? p % (Hin * Win) : 0;
    const int iy = rem / Win, ix = rem % Win;
    const int ci0 = blockIdx.y * CI_T;
    const int cps = (Cout + splitK - 1) / splitK;
    const int c_begin = blockIdx.z * cps, c_end = min(Cout, c_begin + cps);
    // taps: output row / column reached through tap k (clamped so that loads are unconditional) + validity
    int yoff[K], xoff[K];
    bool rok[K], cok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int ty = iy + PAD - k, tx = ix + PAD - k;
        rok[k] = active && ty >= 0 && ty % S == 0 && ty / S < Ho;
        cok[k] = tx >= 0 && tx % S == 0 && tx / S < Wo;
        yoff[k] = (rok[k] ? ty / S : 0) * Wo;
        xoff[k] = cok[k] ? tx / S : 0;
    }
    float acc[CI_T];
#pragma unroll
    for (int i = 0; i < CI_T; ++i) acc[i] = 0.f;
    const size_t plane = (size_t)Ho * Wo;
    const float* dyn = dy + (size_t)n * Cout * plane;
    float cur[KK], nxt[KK];
    auto load_taps = [&](float (&v)[KK], int c) {
        const float* base = dyn + (size_t)c * plane;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) v[ky * K + kx] = __ldg(base + yoff[ky] + xoff[kx]);
    };
    if (c_begin < c_end) load_taps(nxt, c_begin);
    for (int c0 = c_begin; c0 < c_end; c0 += CO_C) {
        const int cn = min(CO_C, c_end - c0);
        for (int idx = threadIdx.x; idx < CO_C * KK * CI_T; idx += CONV_THREADS) {
            const int ci = idx % CI_T, kk = (idx / CI_T) % KK, co = idx / (CI_T * KK);
            s_w[co][kk][ci] = (co < cn && ci0 + ci < Cin) ? Wt[((size_t)(c0 + co) * Cin + ci0 + ci) * KK + kk] : 0.f;
        }
        __syncthreads();
        for (int co = 0; co < cn; ++co) {
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) cur[kk] = nxt[kk];
            if (c0 + co + 1 < c_end) load_taps(nxt, c0 + co + 1);
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float v = (rok[ky] && cok[kx]) ? cur[ky * K + kx] : 0.f;
#pragma unroll
                    for (int ci = 0; ci < CI_T; ++ci) acc[ci] = fmaf(v, s_w[co][ky * K + kx][ci], acc[ci]);
                }
            }
        }
        __syncthreads();
    }
    if (!active) return;
    float* o = dX + (splitK > 1 ? (size_t)blockIdx.z * ((size_t)N * Cin * Hin * Win) : 0);
#pragma unroll
    for (int ci = 0; ci < CI_T; ++ci)
        if (ci0 + ci < Cin) {
            float* q = o + ((size_t)(n * Cin + ci0 + ci) * Hin + iy) * Win + ix;
            *q = (accumulate && splitK == 1) ? *q + acc[ci] : acc[ci];
        }
}

// partial weight / bias gradients over a strided subset of the spatial tiles.
// part layout: [gridDim.x][Cout*Cin*K*K + Cout]  (bias gradient partials at the end)
//   A CTA owns 16 output channels x 8 input channels (x K*K taps) and walks 8 x 32 pixel tiles: x tile (+ halo, producer
//   BatchNorm + LeakyReLU applied while staging) and dy tile in shared memory. A warp visits every 8th pixel of the tile;
//   its 32 threads are the 4 output-channel quads x 8 input channels, each with 4 x K*K accumulators: per pixel ONE
//   128-bit load of dy (the tile is stored pixel-major, channels contiguous, rows padded to 20 floats so that both the
//   transposing store and the load are conflict-free) + K*K loads of x for 4*K*K FMAs.
//   History: round 1 kept the dy tile channel-major - 4 scalar loads per pixel, all four quads in ONE bank (4-way
//   conflict): 25 shared-memory wavefronts per 36 FMAs, FMA pipe 34-38 % active. Padding the channel stride (13
//   wavefronts): 896 px backward 11.4 -> 10.6 ms. An 8-channel register block (72 accumulators, 11 wavefronts per 72
//   FMAs) measured SLOWER (11.3 ms): 122 registers halve the resident warps and the loads of a pixel are exposed.
static constexpr int DYP = 20;   // padded row of the pixel-major dy tile: 16 channels + 4
template <int K, int S>
constexpr int wgrad_smem_floats() {
    constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    constexpr int a = 8 * IH * (IW + 1) + 4 + TH * TW * DYP, b = 8 * 32 * (4 * K * K + 4);
    return a > b ? a : b;
}
template <int K, int S>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, int Cin, int Hin, int Win, InTf tf,
                                                         const float* __restrict__ dy, int Cout, int Ho, int Wo, int N,
                                                         float* __restrict__ part) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CO_T = 16, CI_T = 8, PAD = (K - 1) / 2, KK = K * K;
    constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K, IWP = IW + 1;
    constexpr int SX = CI_T * IH * IWP;
    extern __shared__ __align__(16) float smem[];   // wgrad_smem_floats<K, S>() floats (opt-in dynamic: > 48 KB for K=3, S=2)
    float (*s_x)[IH][IWP] = reinterpret_cast<float (*)[IH][IWP]>(smem);
    float* s_dy = smem + ((SX + 3) & ~3);           // [TH*TW][DYP], 16-byte aligned rows
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
        // dy tile, transposed to pixel-major: one item = (pixel, 4 consecutive channels) -> 4 coalesced global loads, 1 STS.128
        for (int idx = threadIdx.x; idx < TH * TW * (CO_T / 4); idx += 256) {
            const int pix = idx % (TH * TW), c4 = (idx / (TH * TW)) * 4;
            const int oy = ty0 + pix / TW, ox = tx0 + pix % TW;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (oy < Ho && ox < Wo) {
                const float* src = dy + ((size_t)(n * Cout + co0 + c4) * Ho + oy) * Wo + ox;
                const size_t cs = (size_t)Ho * Wo;
                if (co0 + c4 + 0 < Cout) v.x = src[0];
                if (co0 + c4 + 1 < Cout) v.y = src[cs];
                if (co0 + c4 + 2 < Cout) v.z = src[2 * cs];
                if (co0 + c4 + 3 < Cout) v.w = src[3 * cs];
            }
            *reinterpret_cast<float4*>(s_dy + pix * DYP + c4) = v;
        }
        __syncthreads();
        for (int p = g; p < TH * TW; p += 8) {
            const int py = p / TW, px = p % TW;
            const float4 d4 = *reinterpret_cast<const float4*>(s_dy + p * DYP + cos);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
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
        float sv = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) sv += red[(gg * 32 + ww) * RS + e];
        const int wcos = (ww / CI_T) * 4, wci = ww % CI_T;
        if (e < 4 * KK) {
            const int j = e / KK, kk = e % KK;
            const int co = co0 + wcos + j, c = ci0 + wci;
            if (co < Cout && c < Cin) out[((size_t)co * Cin + c) * KK + kk] = sv;
        } else if (ci0 == 0 && wci == 0) {
            const int co = co0 + wcos + (e - 4 * KK);
            if (co < Cout) out[nW + co] = sv;
        }
    }
}
// -------------------------------------------------------------------------------------------------
// Tiled stride-1 "same" convolution for the HIGH-RESOLUTION layers (K = 3 or 1), forward and data gradient.
//   The direct kernels above fetch every tap from global / L1 and keep one pixel per thread: fine for the
//   latency-bound low-resolution layers, ~10-20 % of the FP32 FMA rate once a layer has enough pixels to be
//   throughput-bound (netG at 448 / 896 px: 10 ms and 35 ms per step). Here a CTA owns a TR x 32 output tile x 16
//   output channels: the input tile (+ halo) of 8 reduction channels goes through shared memory once - producer BatchNorm +
//   LeakyReLU and the zero padding applied while staging, the next chunk's global loads issued before this chunk's
//   math - and every thread keeps a 4-pixel x 8-channel register block: 12 FMAs per shared-memory instruction
//   (128-bit input reads, 128-bit broadcast weight reads), 288 FMAs per reduction channel and thread.
//   DGRAD: the same loop over dy with the weights read transposed and flipped (d x = dy * W^T_flipped), no bias,
//   no statistics, optional accumulation.
//   grid (ceil(W/32) * ceil(H/TR), ceil(Cout/16), N), TR * 16 threads. Measured (B200, ncu, netG at 896 px): 23 % (forward) /
//   34 % (data gradient) of the FP32 FMA rate on the layers it serves, against ~12 % for the direct kernels; used only
//   where 16-row tiles fill the GPU twice over (with fewer CTAs the per-chunk staging latency is exposed and the direct
//   kernels win: 13.5 us against 103 us on the 224 px layers). netG forward + backward, 2 calls: 448 px 5.21 -> 5.02 ms,
//   896 px 19.1 -> 17.4 ms; 224 px unchanged. What still limits it: ~40 % of its instructions are the index arithmetic of
//   the staging loops (profiles/ANALYSIS_r2.md).
// -------------------------------------------------------------------------------------------------
template <int K, int TR, bool DGRAD>
__global__ void __launch_bounds__(TR * 16, TR == 8 ? 4 : 2)
conv_tiled_kernel(const float* __restrict__ x, int Cin, int H, int W, InTf tf, const float* __restrict__ Wt, int w_cout, int w_cin,
                  const float* __restrict__ bias, int Cout, float* __restrict__ y, int accumulate, float* __restrict__ stats_part,
                  BnFin fin) {
    pdl_sync();
    constexpr int NT = TR * 16, KK = K * K, PAD = (K - 1) / 2, CI_C = 8, CO_T = 16;
    constexpr int IR = TR + K - 1, IC = 32 + K - 1, ICP = 36;          // staged rows / columns, padded row stride (16-byte rows)
    constexpr int NWARP = NT / 32, WPG = NWARP / 2;                    // warps per output-channel group
    __shared__ __align__(16) float s_in[CI_C][IR][ICP];
    __shared__ __align__(16) float s_w[CI_C][KK][CO_T];
    __shared__ float red[NWARP][8];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int q = tid % (TR * 8), cg = tid / (TR * 8);
    const int r = q / 8, xq = q % 8;
    const int tiles_x = (W + 31) / 32;
    const int ty0 = (blockIdx.x / tiles_x) * TR, tx0 = (blockIdx.x % tiles_x) * 32;
    const int co0 = blockIdx.y * CO_T, n = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* xn = x + (size_t)n * Cin * plane;

    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

    // global -> registers of one chunk's input tile (raw values + validity mask: nothing waits for the loads here),
    // registers -> shared memory one chunk later (producer BatchNorm + LeakyReLU applied, padding zeroed).
    // A warp owns whole tile rows (row id = warp + 8 i: channel c = id / IR, tile row rr = id % IR, advanced without
    // divisions), lane l the column l of the 32 main columns; the K - 1 halo columns of all rows are spread over the first
    // threads. Staging costs ~8 instructions per element this way (the flat index -> (c, row, col) form cost ~25).
    constexpr int NR = CI_C * IR / NWARP;              // rows per warp and chunk (18 for K = 3, 16 for K = 1)
    constexpr int NH = (K - 1) * CI_C * IR;            // halo elements per chunk
    constexpr int NHT = (NH + NT - 1) / NT;            // ... per thread
    static_assert(CI_C * IR % NWARP == 0 && NWARP <= IR, "row ownership");
    const int lane = tid & 31, wid = tid >> 5;
    const int ix_main = tx0 - PAD + lane;
    const bool col_ok = ix_main >= 0 && ix_main < W;
    float pre[NR], preh[NHT > 0 ? NHT : 1];
    unsigned pre_ok = 0, preh_ok = 0;
    auto fetch = [&](int c0) {
        pre_ok = 0; preh_ok = 0;
        int c = 0, rr = wid;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int iy = ty0 - PAD + rr;
            const bool ok = col_ok && c0 + c < Cin && iy >= 0 && iy < H;
            pre[i] = ok ? xn[(size_t)(c0 + c) * plane + (size_t)iy * W + ix_main] : 0.f;
            pre_ok |= ok ? (1u << i) : 0u;
            rr += NWARP;
            if (rr >= IR) { rr -= IR; ++c; }
        }
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
            const int e = tid + h * NT;
            const int row = e / (K - 1 > 0 ? K - 1 : 1), side = e % (K - 1 > 0 ? K - 1 : 1);
            const int c2 = row / IR, iy = ty0 - PAD + row % IR, ix = tx0 - PAD + 32 + side;
            const bool ok = e < NH && c0 + c2 < Cin && iy >= 0 && iy < H && ix < W;
            preh[h] = ok ? xn[(size_t)(c0 + c2) * plane + (size_t)iy * W + ix] : 0.f;
            preh_ok |= ok ? (1u << h) : 0u;
        }
    };
    auto transform = [&](float v, int c) {
        if (!DGRAD && tf.k) {
            const float4 k4 = __ldg(tf.k + c);
            v = fmaf(k4.z, v, k4.w);
            if (tf.lrelu) v = v < 0.f ? v * LRELU : v;
        }
        return v;
    };
    auto stash = [&](int c0) {
        int c = 0, rr = wid;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            s_in[c][rr][lane] = ((pre_ok >> i) & 1u) ? transform(pre[i], c0 + c) : 0.f;
            rr += NWARP;
            if (rr >= IR) { rr -= IR; ++c; }
        }
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
            const int e = tid + h * NT;
            if (e < NH) {
                const int row = e / (K - 1 > 0 ? K - 1 : 1), side = e % (K - 1 > 0 ? K - 1 : 1);
                const int c2 = row / IR;
                s_in[c2][row % IR][32 + side] = ((preh_ok >> h) & 1u) ? transform(preh[h], c0 + c2) : 0.f;
            }
        }
    };

    fetch(0);
    for (int c0 = 0; c0 < Cin; c0 += CI_C) {
        stash(c0);
        for (int idx = tid; idx < CI_C * KK * CO_T; idx += NT) {
            const int co = idx % CO_T, kk = (idx / CO_T) % KK, ci = idx / (CO_T * KK);
            float w = 0.f;
            if (c0 + ci < Cin && co0 + co < Cout) {
                if (DGRAD) w = __ldg(Wt + ((size_t)(c0 + ci) * w_cin + co0 + co) * KK + (KK - 1 - kk));   // transposed, flipped
                else w = __ldg(Wt + ((size_t)(co0 + co) * w_cin + c0 + ci) * KK + kk);
            }
            s_w[ci][kk][co] = w;
        }
        __syncthreads();
        if (c0 + CI_C < Cin) fetch(c0 + CI_C);     // in flight while this chunk's FMAs run
#pragma unroll 1
        for (int ci = 0; ci < CI_C; ++ci) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const float* row = &s_in[ci][r + ky][xq * 4];
                float in[4 + K - 1];
                const float4 a = *reinterpret_cast<const float4*>(row);
                in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w;
                if (K == 3) {
                    const float2 b = *reinterpret_cast<const float2*>(row + 4);
                    in[4] = b.x; in[4 + K - 2] = b.y;
                }
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&s_w[ci][ky * K + kx][cg * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&s_w[ci][ky * K + kx][cg * 8 + 4]);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(in[p + kx], w[j], acc[p][j]);
                }
            }
        }
        __syncthreads();
    }

    const int oy = ty0 + r, ox0 = tx0 + xq * 4;
    const bool rowok = oy < H;
    if (DGRAD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = co0 + cg * 8 + j;
            if (co < Cout && rowok) {
                float* o = y + ((size_t)(n * Cout + co) * H + oy) * W + ox0;
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (ox0 + p < W) o[p] = accumulate ? o[p] + acc[p][j] : acc[p][j];
            }
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int co = co0 + cg * 8 + j;
        const float b = co < Cout ? __ldg(bias + co) : 0.f;
        float* o = y + ((size_t)(n * Cout + min(co, Cout - 1)) * H + min(oy, H - 1)) * W + ox0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc[p][j] += b;
            if (co < Cout && rowok && ox0 + p < W) o[p] = acc[p][j];
        }
    }
    if (stats_part) {
        // per-tile (count, mean, M2) of each output channel over the tile's valid pixels: two block reductions, centred
        const int w = wid;
        const float cnt = (float)(min(TR, H - ty0) * min(32, W - tx0));
        float mean[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float sv = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) sv += (rowok && ox0 + p < W) ? acc[p][j] : 0.f;
            sv = warp_sum(sv);
            if (lane == 0) red[w][j] = sv;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float sv = 0.f;
#pragma unroll
            for (int u = 0; u < WPG; ++u) sv += red[cg * WPG + u][j];
            mean[j] = sv / cnt;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float sv = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float d = acc[p][j] - mean[j];
                sv += (rowok && ox0 + p < W) ? d * d : 0.f;
            }
            sv = warp_sum(sv);
            if (lane == 0) red[w][j] = sv;
        }
        __syncthreads();
        const int part = blockIdx.z * gridDim.x + blockIdx.x, nparts = gridDim.x * gridDim.z;
        if (q == 0) {       // one thread per output-channel group
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int co = co0 + cg * 8 + j;
                if (co < Cout) {
                    float m2 = 0.f;
#pragma unroll
                    for (int u = 0; u < WPG; ++u) m2 += red[cg * WPG + u][j];
                    float* o = stats_part + ((size_t)part * Cout + co) * 3;
                    o[0] = cnt; o[1] = mean[j]; o[2] = m2;
                }
            }
        }
        __syncthreads();    // both groups' partials are written before the ticket of this tile is drawn
        if (fin.konst) bn_finish_if_last(stats_part, nparts, Cout, co0, CO_T, blockIdx.y, nparts, fin, &s_flag);
    }
}

// the tiled kernel serves stride-1 K = 1 / 3 layers with enough pixels per image to be throughput-bound
static inline bool use_tiled(int K, int S, int H, int W) {
    static int on = -1;     // SPLICE_B200_GEN_TILED=0: every layer on the direct kernels (A/B cross-check of the two conv paths)
    if (on < 0) {
        const char* v = getenv("SPLICE_B200_GEN_TILED");
        on = (v && v[0] == '0') ? 0 : 1;
    }
    return on && S == 1 && (K == 1 || K == 3) && H * W >= 2500 && W >= 24;
}
// ... and only when 16-row tiles fill the GPU twice over: with fewer CTAs the per-chunk staging latency of a CTA is exposed
// (measured: 103 us against 13.5 us of the direct kernel for the 224 px layers)
// SPLICE_B200_GEN_TC=1: the 3x3 layers that fill the GPU run on the tcgen05 (3 x TF32) implicit-GEMM kernel (conv_tc.cu)
static inline bool use_tc(int K, int S, int N, int H, int W) {
    static int on = -1;
    if (on < 0) {
        const char* v = getenv("SPLICE_B200_GEN_TC");
        on = (v && v[0] == '1') ? 1 : 0;
    }
    return on && K == 3 && S == 1 && (long long)ceil_div(W, 32) * ceil_div(H, 4) * N >= 2 * 148;
}
static inline bool tiled_fills(int N, int Cout, int H, int W) { return (long long)ceil_div(W, 32) * ceil_div(H, 16) * ceil_div(Cout, 16) * N >= 2 * 148; }
template <bool DGRAD>
static int launch_conv_tiled(int K, const float* x, int N, int Cin, int H, int W, InTf tf, const float* Wt, int w_cout, int w_cin,
                             const float* bias, int Cout, float* y, int accumulate, float* stats_part, BnFin fin, cudaStream_t st) {
    dim3 grid(ceil_div(W, 32) * ceil_div(H, 16), ceil_div(Cout, 16), N);
    if (K == 3) SPLICE_CHECK_CUDA(launch_pdl(conv_tiled_kernel<3, 16, DGRAD>, grid, dim3(256), 0, st, x, Cin, H, W, tf, Wt, w_cout, w_cin, bias, Cout, y, accumulate, stats_part, fin));
    else SPLICE_CHECK_CUDA(launch_pdl(conv_tiled_kernel<1, 16, DGRAD>, grid, dim3(256), 0, st, x, Cin, H, W, tf, Wt, w_cout, w_cin, bias, Cout, y, accumulate, stats_part, fin));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

// -------------------------------------------------------------------------------------------------
// host: launch helpers
// -------------------------------------------------------------------------------------------------
static constexpr int TARGET_BLOCKS = 296;          // two CTAs per SM on 148 SMs
static constexpr size_t SPLIT_FLOATS = 8u << 20;   // scratch for split partial sums (floats)
static constexpr size_t WGRAD_FLOATS = 8u << 20;   // scratch for weight-gradient partials (floats)

// Channel tile and split factor of a layer. A thread does CT FMAs per loaded tap, so wide channel tiles are kept
// (CT = 16 whenever the layer has >= 16 fast channels) and grids that would not fill 148 SMs are widened by splitting
// the reduction channels instead (partials are folded by conv_finish_bn_kernel / sum_partials_kernel).
static void pick_tiling(int P, int Cfast, int Cslow, size_t out_elems, int* ct, int* split) {
    const int pb = ceil_div(P, CONV_THREADS);
    const int t = Cfast <= 4 ? 4 : (Cfast <= 8 ? 8 : 16);
    int sk = 1;
    const int blocks = pb * ceil_div(Cfast, t);
    if (blocks < TARGET_BLOCKS && P <= 16384) {
        sk = ceil_div(TARGET_BLOCKS, blocks);
        const int max_sk = Cslow / 8 > 0 ? Cslow / 8 : 1;          // at least 8 reduction channels per split
        if (sk > max_sk) sk = max_sk;
        if (sk > 16) sk = 16;
        while (sk > 1 && (size_t)sk * out_elems > SPLIT_FLOATS) --sk;
    }
    *ct = t; *split = sk;
}

struct BnOut {                 // where the BatchNorm constants / running statistics of a conv's output go
    const float *gamma, *beta;
    float4* konst;
    float2* bstat;
    int* counter;              // ticket counters of this BatchNorm layer (GEN_COUNTERS_PER_BN ints)
};

static int launch_conv_fwd(int K, int S, const float* x, int N, int Cin, int Hin, int Win, InTf tf, const float* Wt,
                           const float* bias, int Cout, float* y, int Ho, int Wo, int sigmoid, const BnOut* bn, float* stats_part,
                           float* split_part, cudaStream_t st) {
    const int P = N * Ho * Wo;
    if (bn && !sigmoid && Hin == Ho && Win == Wo && use_tc(K, S, N, Ho, Wo))
        return launch_conv_tc<false>(x, N, Cin, Ho, Wo, tf, Wt, Cout, Cin, bias, Cout, y, 0, stats_part,
                                     BnFin{bn->gamma, bn->beta, bn->konst, bn->bstat, bn->counter, 1e-5f}, st);
    if (bn && !sigmoid && Hin == Ho && Win == Wo && use_tiled(K, S, Ho, Wo) && tiled_fills(N, Cout, Ho, Wo))
        return launch_conv_tiled<false>(K, x, N, Cin, Ho, Wo, tf, Wt, Cout, Cin, bias, Cout, y, 0, stats_part,
                                        BnFin{bn->gamma, bn->beta, bn->konst, bn->bstat, bn->counter, 1e-5f}, st);
    int ct, sk;
    pick_tiling(P, Cout, Cin, (size_t)N * Cout * Ho * Wo, &ct, &sk);
    if (!bn) sk = 1;   // the final conv has no BatchNorm epilogue to fold the split reduction into
    dim3 grid(ceil_div(P, CONV_THREADS), ceil_div(Cout, ct), sk);
    float* dst = sk > 1 ? split_part : y;
    float* stats = (sk > 1 || !bn) ? nullptr : stats_part;
    const float eps = 1e-5f;
    BnFin fin{nullptr, nullptr, nullptr, nullptr, nullptr, eps};
    if (bn) fin = BnFin{bn->gamma, bn->beta, bn->konst, bn->bstat, bn->counter, eps};
    BnFin fin_conv = fin;
    if (sk > 1) fin_conv.konst = nullptr;   // the split-K epilogue kernel does the statistics
#define CF(KK, SS, CT) SPLICE_CHECK_CUDA(launch_pdl(conv_fwd_kernel<KK, SS, CT>, grid, dim3(CONV_THREADS), 0, st, x, N, Cin, Hin, Win, tf, Wt, bias, Cout, dst, Ho, Wo, sigmoid, stats, sk, fin_conv))
#define CF3(KK, SS) do { if (ct == 4) CF(KK, SS, 4); else if (ct == 8) CF(KK, SS, 8); else CF(KK, SS, 16); } while (0)
    if (K == 1 && S == 1) CF3(1, 1);
    else if (K == 3 && S == 1) CF3(3, 1);
    else if (K == 3 && S == 2) CF3(3, 2);
    else { set_error("generator: unsupported conv k=%d stride=%d", K, S); return SPLICE_ERR_UNSUPPORTED; }
#undef CF3
#undef CF
    SPLICE_LAUNCH_CHECK();
    if (bn && sk > 1) {
        dim3 fgrid(ceil_div(P, 1024), Cout);
        SPLICE_CHECK_CUDA(launch_pdl(conv_finish_stats_kernel, fgrid, dim3(256), 0, st, (const float*)split_part, sk, bias, y, N, Cout, Ho * Wo, stats_part, fin));
        SPLICE_LAUNCH_CHECK();
    }
    return SPLICE_OK;
}

static int launch_conv_dgrad(int K, int S, const float* dy, int N, int Cout, int Ho, int Wo, const float* Wt, int Cin, float* dX,
                             int Hin, int Win, int accumulate, float* split_part, cudaStream_t st) {
    const int P = N * Hin * Win;
    if (Hin == Ho && Win == Wo && use_tc(K, S, N, Hin, Win))
        return launch_conv_tc<true>(dy, N, Cout, Hin, Win, InTf{nullptr, 0}, Wt, Cout, Cin, nullptr, Cin, dX, accumulate, nullptr,
                                    BnFin{nullptr, nullptr, nullptr, nullptr, nullptr, 1e-5f}, st);
    if (Hin == Ho && Win == Wo && use_tiled(K, S, Hin, Win) && tiled_fills(N, Cin, Hin, Win))
        return launch_conv_tiled<true>(K, dy, N, Cout, Hin, Win, InTf{nullptr, 0}, Wt, Cout, Cin, nullptr, Cin, dX, accumulate, nullptr,
                                       BnFin{nullptr, nullptr, nullptr, nullptr, nullptr, 1e-5f}, st);
    int ct, sk;
    const size_t total = (size_t)N * Cin * Hin * Win;
    pick_tiling(P, Cin, Cout, total, &ct, &sk);
    dim3 grid(ceil_div(P, CONV_THREADS), ceil_div(Cin, ct), sk);
    float* dst = sk > 1 ? split_part : dX;
#define DG(KK, SS, CT) SPLICE_CHECK_CUDA(launch_pdl(conv_dgrad_kernel<KK, SS, CT>, grid, dim3(CONV_THREADS), 0, st, dy, N, Cout, Ho, Wo, Wt, Cin, dst, Hin, Win, accumulate, sk))
#define DG3(KK, SS) do { if (ct == 4) DG(KK, SS, 4); else if (ct == 8) DG(KK, SS, 8); else DG(KK, SS, 16); } while (0)
    if (K == 1 && S == 1) DG3(1, 1);
    else if (K == 3 && S == 1) DG3(3, 1);
    else if (K == 3 && S == 2) DG3(3, 2);
    else { set_error("generator: unsupported conv k=%d stride=%d", K, S); return SPLICE_ERR_UNSUPPORTED; }
#undef DG3
#undef DG
    SPLICE_LAUNCH_CHECK();
    if (sk > 1) {
        const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        SPLICE_CHECK_CUDA(launch_pdl(sum_partials_kernel, dim3(blocks), dim3(256), 0, st, (const float*)split_part, sk, total, dX, accumulate));
        SPLICE_LAUNCH_CHECK();
    }
    return SPLICE_OK;
}

// test hook (capi: splice_gen_debug_conv): ONE stride-1 convolution (or its data gradient) on either kernel family, without
// producer transform or statistics, so that the tiled and the direct kernels can be checked against a reference separately
int gen_debug_conv(const float* x, int N, int Cin, int H, int W, const float* Wt, int Cout, int K, const float* bias, float* y,
                   int dgrad, int tiled, cudaStream_t st) {
    SPLICE_REQUIRE(K == 1 || K == 3, "gen_debug_conv: K must be 1 or 3");
    const InTf none{nullptr, 0};
    const BnFin nofin{nullptr, nullptr, nullptr, nullptr, nullptr, 1e-5f};
    if (tiled == 2) {   // tcgen05 implicit GEMM (3x3 only)
        SPLICE_REQUIRE(K == 3, "gen_debug_conv: the tcgen05 kernel is 3x3 only");
        if (!dgrad) return launch_conv_tc<false>(x, N, Cin, H, W, none, Wt, Cout, Cin, bias, Cout, y, 0, nullptr, nofin, st);
        return launch_conv_tc<true>(x, N, Cout, H, W, none, Wt, Cout, Cin, nullptr, Cin, y, 0, nullptr, nofin, st);
    }
    if (!dgrad) {   // y[N,Cout,H,W] = conv(x[N,Cin,H,W], Wt[Cout,Cin,K,K]) + bias
        if (tiled) return launch_conv_tiled<false>(K, x, N, Cin, H, W, none, Wt, Cout, Cin, bias, Cout, y, 0, nullptr, nofin, st);
        dim3 grid(ceil_div(N * H * W, CONV_THREADS), ceil_div(Cout, 16), 1);
        if (K == 3) conv_fwd_kernel<3, 1, 16><<<grid, CONV_THREADS, 0, st>>>(x, N, Cin, H, W, none, Wt, bias, Cout, y, H, W, 0, nullptr, 1, nofin);
        else conv_fwd_kernel<1, 1, 16><<<grid, CONV_THREADS, 0, st>>>(x, N, Cin, H, W, none, Wt, bias, Cout, y, H, W, 0, nullptr, 1, nofin);
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    }
    // y[N,Cin,H,W] = d x given x = dy[N,Cout,H,W]
    if (tiled) return launch_conv_tiled<true>(K, x, N, Cout, H, W, none, Wt, Cout, Cin, nullptr, Cin, y, 0, nullptr, nofin, st);
    dim3 grid(ceil_div(N * H * W, CONV_THREADS), ceil_div(Cin, 16), 1);
    if (K == 3) conv_dgrad_kernel<3, 1, 16><<<grid, CONV_THREADS, 0, st>>>(x, N, Cout, H, W, Wt, Cin, y, H, W, 0, 1);
    else conv_dgrad_kernel<1, 1, 16><<<grid, CONV_THREADS, 0, st>>>(x, N, Cout, H, W, Wt, Cin, y, H, W, 0, 1);
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
    for (auto& s : slots_) {
        cudaFree(s.pool);
        cudaFree(s.scratch);
        cudaFree(s.counters);
        if (s.side) cudaStreamDestroy(s.side);
        if (s.ev_fork) cudaEventDestroy(s.ev_fork);
        if (s.ev_join) cudaEventDestroy(s.ev_join);
    }
}

int GenEngine::ensure_scratch(Slot& s, size_t bytes) {
    if (!s.side) {
        SPLICE_CHECK_CUDA(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
    }
    if (!s.counters) {
        SPLICE_CHECK_CUDA(cudaMalloc(&s.counters, (size_t)GEN_BN * GEN_COUNTERS_PER_BN * sizeof(int)));
        SPLICE_CHECK_CUDA(cudaMemset(s.counters, 0, (size_t)GEN_BN * GEN_COUNTERS_PER_BN * sizeof(int)));
    }
    if (bytes <= s.scratch_bytes) return SPLICE_OK;
    SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(s.scratch);
    s.scratch = nullptr;
    s.scratch_bytes = 0;
    SPLICE_CHECK_CUDA(cudaMalloc(&s.scratch, bytes));
    s.scratch_bytes = bytes;
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
    plan((size_t)N * 3 * H * W * 4);  // d(pre-sigmoid)
    plan((size_t)N * 3 * H * W * 4);  // dout copy (stable address for the backward graph)
    plan((size_t)GEN_BN * BSTAT_STRIDE * sizeof(float2));  // batch statistics (mean, unbiased variance) per BN layer
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
    s.dfin = (float*)nx();
    s.dout_copy = (float*)nx();
    s.bstat = (float2*)nx();
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
    NvtxRange nvtx("splice_gen_forward");
    SPLICE_REQUIRE(slot >= 0 && slot < GEN_SLOTS, "generator: slot out of range");
    SPLICE_REQUIRE(x && out && N > 0 && H > 0 && W > 0, "generator: bad input");
    Slot& s = slots_[slot];
    GRC(configure(s, N, H, W));
    // scratch = [statistics partials | split partial sums | weight-gradient partials]
    const size_t stats_floats = (size_t)ceil_div(N * H * W, CONV_THREADS) * 132 * 3 + (size_t)N * ceil_div(H, TH) * ceil_div(W, TW) * 132 * 3;
    GRC(ensure_scratch(s, (stats_floats + SPLIT_FLOATS + WGRAD_FLOATS) * sizeof(float) + 4096));
    s.stats_floats = stats_floats;
    // keep a private copy of the input: the caller's tensor may be freed before backward() (wgrad of scale 0 reads it)
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(s.x_copy, x, (size_t)N * 3 * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s.x = s.x_copy;

    if (use_graphs_) {
        KeyHasher k;
        k.add((uint64_t)11).add((uint64_t)slot).add((uint64_t)N).add((uint64_t)H).add((uint64_t)W).add(s.pool).add(s.scratch);
        for (int i = 0; i < GEN_PARAMS; ++i) k.add(p.param[i]);
        GRC(graphs_.run(k.h, st, [&](cudaStream_t cs) { return forward_body(p, s, cs); }));
    } else {
        GRC(forward_body(p, s, st));
    }
    s.stats_pending = true;
    if (update_running) GRC(update_running_stats(p, slot, st));
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(out, s.out, (size_t)N * 3 * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s.valid = keep;
    return SPLICE_OK;
}

int GenEngine::update_running_stats(const GenPointers& p, int slot, cudaStream_t st) {
    SPLICE_REQUIRE(slot >= 0 && slot < GEN_SLOTS, "generator: slot out of range");
    Slot& s = slots_[slot];
    SPLICE_REQUIRE(s.pool && s.stats_pending, "generator: slot %d holds no forward pass whose batch statistics are unapplied", slot);
    RunningTable t;
    const Bn* order[GEN_BN];
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        const Bn* b[6] = {&c.bs, &c.bd1, &c.bd2, &c.bcat, &c.bc1, &c.bc2};
        for (int k = 0; k < 6; ++k) order[b[k]->idx] = b[k];
    }
    for (int l = 0; l < GEN_BN; ++l) {
        SPLICE_REQUIRE(p.running_mean[l] && p.running_var[l] && p.num_batches_tracked[l], "generator: BN buffer %d is null", l);
        t.bstat[l] = s.bstat + (size_t)l * BSTAT_STRIDE;
        t.rmean[l] = p.running_mean[l]; t.rvar[l] = p.running_var[l]; t.nbt[l] = p.num_batches_tracked[l];
        t.C[l] = order[l]->c;
    }
    SPLICE_CHECK_CUDA(launch_pdl(update_running_kernel, dim3(GEN_BN), dim3(160), 0, st, t, 0.1f));
    SPLICE_LAUNCH_CHECK();
    s.stats_pending = false;
    return SPLICE_OK;
}

int GenEngine::forward_body(const GenPointers& p, Slot& s, cudaStream_t st) {
    const int N = s.N, H = s.H, W = s.W;
    const size_t stats_floats = s.stats_floats;
    float* part = static_cast<float*>(s.scratch);
    float* split = part + stats_floats;
    const float eps = 1e-5f;

    auto bn_out = [&](const Bn& b, float4* k) {
        return BnOut{p.param[b.pg], p.param[b.pb], k, s.bstat + (size_t)b.idx * BSTAT_STRIDE,
                     s.counters + (size_t)b.idx * GEN_COUNTERS_PER_BN};
    };
    auto conv_bn_on = [&](const Conv& c, const Bn& b, const float* in, int hin, int win, InTf tf, float* y, int ho, int wo,
                          float4* k, float* stats_scratch, float* split_scratch, cudaStream_t cs) -> int {
        const BnOut bo = bn_out(b, k);
        return launch_conv_fwd(c.k, c.stride, in, N, c.cin, hin, win, tf, p.param[c.pw], p.param[c.pb], c.cout, y, ho, wo, 0, &bo,
                               stats_scratch, split_scratch, cs);
    };
    auto conv_bn = [&](const Conv& c, const Bn& b, const float* in, int hin, int win, InTf tf, float* y, int ho, int wo,
                       float4* k) -> int { return conv_bn_on(c, b, in, hin, win, tf, y, ho, wo, k, part, split, st); };

    // down path. The 1x1 skip convolutions only feed the concats of the up path: they run on the slot's side stream (a
    // parallel branch of the captured graph) with their own scratch (the weight-gradient region, idle during a forward).
    cudaStream_t ss = s.side;
    float* part_skip = split + SPLIT_FLOATS;
    float* split_skip = part_skip + (1u << 20);
    const float* in = s.x;
    InTf tf_in{nullptr, 0};
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        SPLICE_CHECK_CUDA(cudaEventRecord(s.ev_fork, st));
        SPLICE_CHECK_CUDA(cudaStreamWaitEvent(ss, s.ev_fork, 0));
        GRC(conv_bn_on(c.s, c.bs, in, b.h, b.w, tf_in, b.s_raw, b.h, b.w, b.k_s, part_skip, split_skip, ss));
        GRC(conv_bn(c.d1, c.bd1, in, b.h, b.w, tf_in, b.d1_raw, b.hd, b.wd, b.k_d1));
        GRC(conv_bn(c.d2, c.bd2, b.d1_raw, b.hd, b.wd, InTf{b.k_d1, 1}, b.d2_raw, b.hd, b.wd, b.k_d2));
        in = b.d2_raw;
        tf_in = InTf{b.k_d2, 1};
    }
    SPLICE_CHECK_CUDA(cudaEventRecord(s.ev_join, ss));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(st, s.ev_join, 0));
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
        {
            const BnOut bo = bn_out(c.bcat, b.k_cat);
            SPLICE_CHECK_CUDA(launch_pdl(cat_build_kernel, grid, dim3(256), 0, st, (const float*)b.s_raw, 4, b.h, b.w, InTf{b.k_s, 1}, oys, oxs, u, c.cdeep, hu,
                                         wu, tf_u, oyu, oxu, b.cat, th, tw, part, BnFin{bo.gamma, bo.beta, bo.konst, bo.bstat, bo.counter, eps}));
            SPLICE_LAUNCH_CHECK();
        }
        GRC(conv_bn(c.c1, c.bc1, b.cat, th, tw, InTf{b.k_cat, 0}, b.c1_raw, th, tw, b.k_c1));
        GRC(conv_bn(c.c2, c.bc2, b.c1_raw, th, tw, InTf{b.k_c1, 1}, b.c2_raw, th, tw, b.k_c2));
    }
    GRC(launch_conv_fwd(1, 1, s.sb[0].c2_raw, N, final_.cin, H, W, InTf{s.sb[0].k_c2, 1}, p.param[final_.pw], p.param[final_.pb], 3,
                        s.out, H, W, 1, nullptr, nullptr, nullptr, st));
    return SPLICE_OK;
}

int GenEngine::backward(const GenPointers& p, const float* dout, int slot, bool accumulate, cudaStream_t st) {
    NvtxRange nvtx("splice_gen_backward");
    SPLICE_REQUIRE(slot >= 0 && slot < GEN_SLOTS, "generator: slot out of range");
    Slot& s = slots_[slot];
    SPLICE_REQUIRE(s.pool && s.valid, "generator backward: slot %d holds no kept forward pass", slot);
    SPLICE_REQUIRE(dout, "generator backward: null gradient");
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(s.dout_copy, dout, (size_t)s.N * 3 * s.H * s.W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (use_graphs_) {
        KeyHasher k;
        k.add((uint64_t)12).add((uint64_t)slot).add((uint64_t)s.N).add((uint64_t)s.H).add((uint64_t)s.W).add(s.pool).add(s.scratch)
            .add((uint64_t)accumulate);
        for (int i = 0; i < GEN_PARAMS; ++i) k.add(p.param[i]).add(p.grad[i]);
        GRC(graphs_.run(k.h, st, [&](cudaStream_t cs) { return backward_body(p, s, accumulate, cs); }));
    } else {
        GRC(backward_body(p, s, accumulate, st));
    }
    s.valid = false;
    return SPLICE_OK;
}

int GenEngine::backward_body(const GenPointers& p, Slot& s, bool accumulate, cudaStream_t st) {
    const int N = s.N, H = s.H, W = s.W;
    const int acc = accumulate ? 1 : 0;
    float* part = static_cast<float*>(s.scratch);
    float* split = part + s.stats_floats;
    float* wpart = split + SPLIT_FLOATS;

    // BatchNorm(+LeakyReLU) backward of one layer: afterwards dA holds d(raw conv output)
    auto bn_bwd = [&](const Bn& b, float* dA, const float* y, const float4* k, int lrelu, int hh, int ww, float2* m) -> int {
        const int HW = hh * ww;
        if ((size_t)N * HW <= 8192) {
            SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_small_kernel, dim3(b.c), dim3(256), 0, st, dA, y, k, lrelu, N, b.c, HW, p.grad[b.pg], p.grad[b.pb], acc));
            SPLICE_LAUNCH_CHECK();
            return SPLICE_OK;
        }
        dim3 grid(ceil_div(HW, 2048), b.c, N);
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_reduce_kernel, grid, dim3(256), 0, st, (const float*)dA, y, k, lrelu, b.c, HW, part));
        SPLICE_LAUNCH_CHECK();
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3(b.c), dim3(32), 0, st, (const float*)part, N * (int)grid.x, b.c, (double)N * HW, p.grad[b.pg], p.grad[b.pb], m, acc));
        SPLICE_LAUNCH_CHECK();
        const size_t total = (size_t)N * b.c * HW;
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_apply_kernel, dim3(blocks), dim3(256), 0, st, dA, y, k, (const float2*)m, lrelu, b.c, HW, total));
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };
    // Weight gradients hang off the critical path (bn_bwd -> dgrad -> bn_bwd -> ...): each one only needs this layer's dy,
    // which is complete on `st` when wgrad() is called, so they run on the slot's side stream (a parallel branch of the
    // captured graph) and are joined at the end of the pass. The side stream is in order, so `wpart` is reused safely.
    cudaStream_t ws = s.side;
    auto wgrad = [&](const Conv& c, const float* in, int hin, int win, InTf tf, const float* dy, int ho, int wo) -> int {
        SPLICE_CHECK_CUDA(cudaEventRecord(s.ev_fork, st));
        SPLICE_CHECK_CUDA(cudaStreamWaitEvent(ws, s.ev_fork, 0));
        const int ntiles = N * ceil_div(ho, TH) * ceil_div(wo, TW);
        const size_t nW = (size_t)c.cout * c.cin * c.k * c.k;
        size_t cap = WGRAD_FLOATS / (nW + c.cout);
        if (cap > 256) cap = 256;
        if (cap < 1) cap = 1;
        const int chunks = ntiles < (int)cap ? ntiles : (int)cap;
        dim3 grid(chunks, ceil_div(c.cout, 16) * ceil_div(c.cin, 8));
        static bool attr = false;
        if (!attr) {
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<1, 1>() * 4));
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<3, 1>() * 4));
            SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_floats<3, 2>() * 4));
            attr = true;
        }
        if (c.k == 1 && c.stride == 1)
            SPLICE_CHECK_CUDA(launch_pdl(conv_wgrad_kernel<1, 1>, grid, dim3(256), wgrad_smem_floats<1, 1>() * 4, ws, in, c.cin, hin, win, tf, dy, c.cout, ho, wo, N, wpart));
        else if (c.k == 3 && c.stride == 1)
            SPLICE_CHECK_CUDA(launch_pdl(conv_wgrad_kernel<3, 1>, grid, dim3(256), wgrad_smem_floats<3, 1>() * 4, ws, in, c.cin, hin, win, tf, dy, c.cout, ho, wo, N, wpart));
        else
            SPLICE_CHECK_CUDA(launch_pdl(conv_wgrad_kernel<3, 2>, grid, dim3(256), wgrad_smem_floats<3, 2>() * 4, ws, in, c.cin, hin, win, tf, dy, c.cout, ho, wo, N, wpart));
        SPLICE_LAUNCH_CHECK();
        SPLICE_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(ceil_div((int)(nW + c.cout), 256)), dim3(256), 0, ws, (const float*)wpart, chunks, nW, c.cout, p.grad[c.pw], p.grad[c.pb], acc));
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };
    auto dgrad = [&](const Conv& c, const float* dy, int ho, int wo, float* dX, int hin, int win, int accumulate) -> int {
        return launch_conv_dgrad(c.k, c.stride, dy, N, c.cout, ho, wo, p.param[c.pw], c.cin, dX, hin, win, accumulate, split, st);
    };

    // final 1x1 conv + sigmoid
    {
        const size_t total = (size_t)N * 3 * H * W;
        SPLICE_CHECK_CUDA(launch_pdl(sigmoid_bwd_kernel, dim3((int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8)), dim3(256), 0, st,
                                     (const float*)s.dout_copy, (const float*)s.out, s.dfin, total));
        SPLICE_LAUNCH_CHECK();
        GRC(wgrad(final_, s.sb[0].c2_raw, H, W, InTf{s.sb[0].k_c2, 1}, s.dfin, H, W));
        GRC(dgrad(final_, s.dfin, H, W, s.sb[0].dA_c2, H, W, 0));
    }
    // up path, top to bottom
    for (int i = 0; i < GEN_SCALES; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = s.sb[i];
        const int h = b.h, w = b.w;
        GRC(bn_bwd(c.bc2, b.dA_c2, b.c2_raw, b.k_c2, 1, h, w, b.m_c2));
        GRC(wgrad(c.c2, b.c1_raw, h, w, InTf{b.k_c1, 1}, b.dA_c2, h, w));
        GRC(dgrad(c.c2, b.dA_c2, h, w, b.dA_c1, h, w, 0));

        GRC(bn_bwd(c.bc1, b.dA_c1, b.c1_raw, b.k_c1, 1, h, w, b.m_c1));
        GRC(wgrad(c.c1, b.cat, h, w, InTf{b.k_cat, 0}, b.dA_c1, h, w));
        GRC(dgrad(c.c1, b.dA_c1, h, w, b.dcat, h, w, 0));

        GRC(bn_bwd(c.bcat, b.dcat, b.cat, b.k_cat, 0, h, w, b.m_cat));
        const int C = 4 + c.cdeep, hu = b.hd, wu = b.wd;
        const int oyu = (2 * hu - h) / 2, oxu = (2 * wu - w) / 2;
        {
            dim3 grid(min(ceil_div(4 * h * w, 256), 148 * 8), 1, N);
            SPLICE_CHECK_CUDA(launch_pdl(cat_bwd_skip_kernel, grid, dim3(256), 0, st, (const float*)b.dcat, C, h, w, 4, h, w, 0, 0, b.dA_s));
            SPLICE_LAUNCH_CHECK();
            float* dU = (i == GEN_SCALES - 1) ? b.dA_d2 : s.sb[i + 1].dA_c2;
            dim3 grid2(min(ceil_div(c.cdeep * hu * wu, 256), 148 * 8), 1, N);
            SPLICE_CHECK_CUDA(launch_pdl(cat_bwd_up_kernel, grid2, dim3(256), 0, st, (const float*)b.dcat, C, h, w, 4, c.cdeep, hu, wu, oyu, oxu, dU));
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
        GRC(wgrad(c.s, in, b.h, b.w, tf_in, b.dA_s, b.h, b.w));
        if (dIn) GRC(dgrad(c.s, b.dA_s, b.h, b.w, dIn, b.h, b.w, 0));

        GRC(bn_bwd(c.bd2, b.dA_d2, b.d2_raw, b.k_d2, 1, b.hd, b.wd, b.m_d2));
        GRC(wgrad(c.d2, b.d1_raw, b.hd, b.wd, InTf{b.k_d1, 1}, b.dA_d2, b.hd, b.wd));
        GRC(dgrad(c.d2, b.dA_d2, b.hd, b.wd, b.dA_d1, b.hd, b.wd, 0));

        GRC(bn_bwd(c.bd1, b.dA_d1, b.d1_raw, b.k_d1, 1, b.hd, b.wd, b.m_d1));
        GRC(wgrad(c.d1, in, b.h, b.w, tf_in, b.dA_d1, b.hd, b.wd));
        if (dIn) GRC(dgrad(c.d1, b.dA_d1, b.hd, b.wd, dIn, b.h, b.w, 1));
    }
    SPLICE_CHECK_CUDA(cudaEventRecord(s.ev_join, ws));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(st, s.ev_join, 0));
    return SPLICE_OK;
}

}  // namespace splice
