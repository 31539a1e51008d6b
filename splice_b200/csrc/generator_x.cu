// splice_b200 — native generator for the OTHER skip() configurations the reference uses: inversion.py:21-25 builds
//   skip(32, 3, num_channels_down = up = [16, 32, 64, 128, 128, 128], num_channels_skip = [4] * 6,
//        filter_size_down = filter_size_up = [7, 7, 5, 5, 3, 3], downsample_mode = 'stride', pad = 'reflection')
// i.e. six scales, 7x7 / 5x5 / 3x3 filters, nn.ReflectionPad2d in front of every conv (models/unet/common.py:113-118) and a
// 32-channel noise input, against the optimisation loop's default-argument network that generator.cu serves. This engine
// takes the configuration as data (GenXConfig: scales, channel lists, per-scale filter sizes, zero / reflection padding,
// sigmoid) and covers any skip() built from strided convs, BatchNorm2d (training mode), LeakyReLU(0.2), bilinear x2
// up-sampling and 1x1 "need1x1_up" convs. Same fusion plan as generator.cu (fp32, NCHW):
//   * a conv applies its PRODUCER's BatchNorm affine + LeakyReLU while loading, adds the bias, writes the raw output once
//     and reduces that output's batch statistics in the same pass ((count, mean, M2) partials per block, merged by the
//     block that draws the last ticket: gen_dev.cuh);
//   * reflection padding is an index map inside the forward and weight-gradient kernels (no padded copy). Its adjoint -
//     border pixels receive the gradient of the virtual pixels mirrored onto them - is done in two steps: the data-gradient
//     kernel evaluates the gradient on the padded domain [-p, H-1+p] x [-p, W-1+p], reflect_fold_kernel adds the <= 9
//     mirror images of every pixel;
//   * weight gradients of a K x K filter are K independent 1 x K problems (grid.z = filter row): 4 x K accumulators per
//     thread whatever K is, the input tile staged once per filter row;
//   * concat + crop + bilinear up-sampling, BatchNorm backward, sigmoid backward and the weight-gradient fold are the
//     kernels generator.cu uses (gen_kernels.cuh).
// Direct fp32 SIMT convolutions: this path serves a batch-1, 224 px, 20 000-iteration feature inversion whose cost is the
// ViT; it is parity-first (fp32 like the reference), not a throughput showcase.
// The file compiles two ways: by nvcc into the product library, and by g++ -DSPLICE_EMU against tests/emu/cuda_emu.h so that
// the same kernel bodies and host orchestration run on the CPU-only build box (tests/test_genx_emu.py).
#include "generator_x.h"

#include <vector>

#include "gen_dev.cuh"
#include "gen_kernels.cuh"
#include "splice_b200.h"

namespace splice {

static constexpr int XCONV_THREADS = 128;
static constexpr int XDYP = 20;                       // padded row of the pixel-major dy tile: 16 channels + 4
static constexpr size_t XWGRAD_FLOATS = 8u << 20;     // scratch for weight-gradient partials (floats)
static constexpr size_t XSPLIT_FLOATS = 4u << 20;     // scratch for the partial sums of split reductions (main stream)
static constexpr size_t XSPLIT_SKIP_FLOATS = 1u << 18;   // same for the skip-branch convolutions on the side stream

// nn.ReflectionPad2d index map: -1 -> 1, n -> n - 2 (identity inside [0, n))
__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// -------------------------------------------------------------------------------------------------
// forward convolution, K x K, stride S, zero or reflection padding of (K-1)/2
//   (+ producer BN/LeakyReLU on load, + bias, + optional sigmoid, + output statistics and BatchNorm constants)
//   one thread per output pixel (linear over N*Ho*Wo), CO_T output channels per thread; taps from global / L1 one filter
//   row at a time, weights of CI_C input channels staged in shared memory. Low-resolution layers whose grids would not fill
//   148 SMs split the input channels over blockIdx.z (partials folded by conv_finish_stats_kernel).
// -------------------------------------------------------------------------------------------------
// pixels per thread: the wide stride-1 filters (5x5, 7x7) give a thread two horizontally adjacent output pixels - a filter row is
// then K + 1 loads for 2 K taps, and every shared-memory weight read feeds two FMAs per channel
template <int K, int S>
struct XPx { static constexpr int value = (S == 1 && K >= 5) ? 2 : 1; };

template <int K, int S, int CO_T>
static __global__ void __launch_bounds__(XCONV_THREADS)
convx_fwd_kernel(const float* __restrict__ x, int N, int Cin, int Hin, int Win, InTf tf, const float* __restrict__ Wt,
                 const float* __restrict__ bias, int Cout, float* __restrict__ y, int Ho, int Wo, int reflect, int out_sigmoid,
                 float* __restrict__ stats_part, int splitK, BnFin fin) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CI_C = (K >= 5 ? 4 : 8), KK = K * K, PAD = (K - 1) / 2, PX = XPx<K, S>::value, NC = K + PX - 1;
    __shared__ __align__(16) float s_w[CI_C][KK][CO_T];
    __shared__ float2 s_ab[CI_C];
    __shared__ float red[4 * (CO_T + 1)];
    __shared__ int s_flag;
    const int Wg = (Wo + PX - 1) / PX;                                 // pixel groups per output row
    const int P = N * Ho * Wg;
    const int p = blockIdx.x * XCONV_THREADS + threadIdx.x;
    const bool active = p < P;
    const int n = active ? p / (Ho * Wg) : 0, rem = active ? p % (Ho * Wg) : 0;
    const int oy = rem / Wg, ox0 = (rem % Wg) * PX;
    bool pok[PX];                                                      // which of this thread's pixels exist
#pragma unroll
    for (int j = 0; j < PX; ++j) pok[j] = active && ox0 + j < Wo;
    const int co0 = blockIdx.y * CO_T;
    const int cps = (Cin + splitK - 1) / splitK;                       // input channels of this split (blockIdx.z)
    const int c_begin = blockIdx.z * cps, c_end = min(Cin, c_begin + cps);
    const int iy0 = oy * S - PAD, ix0 = ox0 * S - PAD;
    // taps are loaded unconditionally from in-image coordinates (reflected, or clamped and masked afterwards)
    int yoff[K], xoff[NC];
    bool rok[K], cok[NC];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        int iy = iy0 + k;
        if (reflect) { rok[k] = true; iy = reflect_idx(iy, Hin); }
        else rok[k] = iy >= 0 && iy < Hin;
        yoff[k] = min(max(iy, 0), Hin - 1) * Win;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        int ix = ix0 + k;
        if (reflect) { cok[k] = true; ix = reflect_idx(ix, Win); }
        else cok[k] = ix >= 0 && ix < Win;
        xoff[k] = min(max(ix, 0), Win - 1);
    }
    float acc[PX][CO_T];
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int i = 0; i < CO_T; ++i) acc[j][i] = 0.f;
    const size_t plane = (size_t)Hin * Win;
    const float* xn = x + (size_t)n * Cin * plane;

    for (int c0 = c_begin; c0 < c_end; c0 += CI_C) {
        const int cn = min(CI_C, c_end - c0);
        for (int idx = threadIdx.x; idx < CI_C * KK * CO_T; idx += XCONV_THREADS) {
            const int co = idx % CO_T, kk = (idx / CO_T) % KK, ci = idx / (CO_T * KK);
            s_w[ci][kk][co] = (ci < cn && co0 + co < Cout) ? Wt[((size_t)(co0 + co) * Cin + c0 + ci) * KK + kk] : 0.f;
        }
        if (threadIdx.x < CI_C) {
            float2 ab = make_float2(1.f, 0.f);
            if (tf.k && (int)threadIdx.x < cn) { const float4 k4 = tf.k[c0 + threadIdx.x]; ab = make_float2(k4.z, k4.w); }
            s_ab[threadIdx.x] = ab;
        }
        __syncthreads();
        for (int ci = 0; ci < cn; ++ci) {
            const float* base = xn + (size_t)(c0 + ci) * plane;
            const float2 ab = s_ab[ci];
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                float row[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) row[k] = __ldg(base + yoff[ky] + xoff[k]);
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    float v = row[k];
                    if (tf.k) {
                        v = fmaf(ab.x, v, ab.y);
                        if (tf.lrelu) v = v < 0.f ? v * LRELU : v;
                    }
                    row[k] = (rok[ky] && cok[k]) ? v : 0.f;   // zero padding lives in the post-BN/activation domain
                }
#pragma unroll
                for (int kx = 0; kx < K; ++kx)
#pragma unroll
                    for (int co = 0; co < CO_T; ++co) {
                        const float wv = s_w[ci][ky * K + kx][co];
#pragma unroll
                        for (int j = 0; j < PX; ++j) acc[j][co] = fmaf(row[kx + j], wv, acc[j][co]);
                    }
            }
        }
        __syncthreads();
    }
    if (splitK > 1) {   // partial sums; conv_finish_stats_kernel adds the bias and does the statistics
        float* o = y + (size_t)blockIdx.z * ((size_t)N * Cout * Ho * Wo);
#pragma unroll
        for (int j = 0; j < PX; ++j)
#pragma unroll
            for (int co = 0; co < CO_T; ++co)
                if (pok[j] && co0 + co < Cout) o[((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox0 + j] = acc[j][co];
        return;
    }
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
        if (co0 + co < Cout) {
            const float b = bias[co0 + co];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float v = acc[j][co] + b;
                if (out_sigmoid) v = 1.f / (1.f + __expf(-v));
                acc[j][co] = v;
                if (pok[j]) y[((size_t)(n * Cout + co0 + co) * Ho + oy) * Wo + ox0 + j] = v;
            }
        }
    }
    if (stats_part) {
        // per-block (count, mean, M2) of each output channel: two block reductions, centred second pass
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < PX; ++j) mine += pok[j] ? 1.f : 0.f;
        {
            const float sv = warp_sum(mine);
            if (lane == 0) red[w * (CO_T + 1) + CO_T] = sv;
        }
        float mean[CO_T], m2[CO_T];
#pragma unroll
        for (int co = 0; co < CO_T; ++co) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < PX; ++j) t += pok[j] ? acc[j][co] : 0.f;
            const float sv = warp_sum(t);
            if (lane == 0) red[w * (CO_T + 1) + co] = sv;
        }
        __syncthreads();
        const float cnt = red[CO_T] + red[(CO_T + 1) + CO_T] + red[2 * (CO_T + 1) + CO_T] + red[3 * (CO_T + 1) + CO_T];
#pragma unroll
        for (int co = 0; co < CO_T; ++co)
            mean[co] = (red[co] + red[(CO_T + 1) + co] + red[2 * (CO_T + 1) + co] + red[3 * (CO_T + 1) + co]) / cnt;
        __syncthreads();
#pragma unroll
        for (int co = 0; co < CO_T; ++co) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const float d = acc[j][co] - mean[co];
                t += pok[j] ? d * d : 0.f;
            }
            const float sv = warp_sum(t);
            if (lane == 0) red[w * (CO_T + 1) + co] = sv;
        }
        __syncthreads();
#pragma unroll
        for (int co = 0; co < CO_T; ++co) m2[co] = red[co] + red[(CO_T + 1) + co] + red[2 * (CO_T + 1) + co] + red[3 * (CO_T + 1) + co];
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

// -------------------------------------------------------------------------------------------------
// data gradient on the (possibly padded) input domain: dXq[n, ci, qy, qx], q = input coordinate + ext, ext = (K-1)/2 for
// reflection padding (the virtual border is part of the domain; reflect_fold_kernel folds it back) and 0 for zero padding.
// One thread per domain pixel, CI_T input channels per thread: dXq[q] = sum_co sum_k W[co, ci, k] dy[co, (q - ext + PAD - k) / S]
// -------------------------------------------------------------------------------------------------
template <int K, int S, int CI_T>
static __global__ void __launch_bounds__(XCONV_THREADS)
convx_dgrad_kernel(const float* __restrict__ dy, int N, int Cout, int Ho, int Wo, const float* __restrict__ Wt, int Cin,
                   float* __restrict__ dXq, int Hq, int Wq, int ext, int accumulate, int splitK) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CO_C = (K >= 5 ? 4 : 8), KK = K * K, PAD = (K - 1) / 2, PX = XPx<K, S>::value, NC = K + PX - 1;
    __shared__ __align__(16) float s_w[CO_C][KK][CI_T];
    const int Wg = (Wq + PX - 1) / PX;                                 // pixel groups per domain row
    const int P = N * Hq * Wg;
    const int p = blockIdx.x * XCONV_THREADS + threadIdx.x;
    const bool active = p < P;
    const int n = active ? p / (Hq * Wg) : 0, rem = active ? p % (Hq * Wg) : 0;
    const int qy = rem / Wg, qx0 = (rem % Wg) * PX;
    const int ci0 = blockIdx.y * CI_T;
    const int cps = (Cout + splitK - 1) / splitK;                      // output channels of this split (blockIdx.z)
    const int c_begin = blockIdx.z * cps, c_end = min(Cout, c_begin + cps);
    // rows: output row reached through tap ky. columns: pixel j of the thread meets tap kx in column tx = qx0 + j - ext + PAD - kx
    // of dy; the PX pixels together touch NC = K + PX - 1 distinct columns, indexed c = K - 1 - kx + j (PX > 1 only for S = 1).
    // Indices are 0 where there is no such output (loads stay unconditional), validity is kept beside them.
    int yoff[K], xoff[NC];
    bool rok[K], cok[NC];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int ty = qy - ext + PAD - k;
        rok[k] = ty >= 0 && ty % S == 0 && ty / S < Ho;
        yoff[k] = (rok[k] ? ty / S : 0) * Wo;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int tx = qx0 - ext + PAD - (K - 1) + c;
        cok[c] = tx >= 0 && tx % S == 0 && tx / S < Wo;
        xoff[c] = cok[c] ? tx / S : 0;
    }
    float acc[PX][CI_T];
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int i = 0; i < CI_T; ++i) acc[j][i] = 0.f;
    const size_t plane = (size_t)Ho * Wo;
    const float* dyn = dy + (size_t)n * Cout * plane;
    for (int c0 = c_begin; c0 < c_end; c0 += CO_C) {
        const int cn = min(CO_C, c_end - c0);
        for (int idx = threadIdx.x; idx < CO_C * KK * CI_T; idx += XCONV_THREADS) {
            const int ci = idx % CI_T, kk = (idx / CI_T) % KK, co = idx / (CI_T * KK);
            s_w[co][kk][ci] = (co < cn && ci0 + ci < Cin) ? Wt[((size_t)(c0 + co) * Cin + ci0 + ci) * KK + kk] : 0.f;
        }
        __syncthreads();
        for (int co = 0; co < cn; ++co) {
            const float* base = dyn + (size_t)(c0 + co) * plane;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                float row[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) row[c] = __ldg(base + yoff[ky] + xoff[c]);
#pragma unroll
                for (int c = 0; c < NC; ++c) row[c] = (rok[ky] && cok[c]) ? row[c] : 0.f;
#pragma unroll
                for (int kx = 0; kx < K; ++kx)
#pragma unroll
                    for (int ci = 0; ci < CI_T; ++ci) {
                        const float wv = s_w[co][ky * K + kx][ci];
#pragma unroll
                        for (int j = 0; j < PX; ++j) acc[j][ci] = fmaf(row[K - 1 - kx + j], wv, acc[j][ci]);
                    }
            }
        }
        __syncthreads();
    }
    if (!active) return;
    float* o = dXq + (splitK > 1 ? (size_t)blockIdx.z * ((size_t)N * Cin * Hq * Wq) : 0);   // split: partials for sum_partials_kernel
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int ci = 0; ci < CI_T; ++ci)
            if (ci0 + ci < Cin && qx0 + j < Wq) {
                float* q = o + ((size_t)(n * Cin + ci0 + ci) * Hq + qy) * Wq + qx0 + j;
                *q = (accumulate && splitK == 1) ? *q + acc[j][ci] : acc[j][ci];
            }
}

// adjoint of nn.ReflectionPad2d(p): dX[y, x] (=|+=) sum of dXq over (y, x) and its mirror images in the border of width p
// (rows -y for 1 <= y <= p and 2(H-1) - y for H-1-p <= y <= H-2, likewise columns; fixed summation order)
static __global__ void __launch_bounds__(256) reflect_fold_kernel(const float* __restrict__ dXq, int H, int W, int p, size_t total,
                                                                  float* __restrict__ dX, int accumulate) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const int Hq = H + 2 * p, Wq = W + 2 * p;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const size_t nc = i / ((size_t)H * W);
        int ys[3], xs[3], ny = 1, nx = 1;
        ys[0] = y; xs[0] = x;
        if (y >= 1 && y <= p) ys[ny++] = -y;
        if (y <= H - 2 && y >= H - 1 - p) ys[ny++] = 2 * (H - 1) - y;
        if (x >= 1 && x <= p) xs[nx++] = -x;
        if (x <= W - 2 && x >= W - 1 - p) xs[nx++] = 2 * (W - 1) - x;
        float s = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nx; ++b) s += dXq[(nc * Hq + ys[a] + p) * Wq + xs[b] + p];
        dX[i] = accumulate ? dX[i] + s : s;
    }
}

// -------------------------------------------------------------------------------------------------
// partial weight / bias gradients of ONE FILTER ROW (blockIdx.z = ky) over a strided subset of the spatial tiles.
// part layout: [gridDim.x][Cout*Cin*K*K + Cout]  (bias gradient partials at the end, written by the ky = 0, first-input-tile CTAs)
//   A CTA owns 16 output channels x 8 input channels x K taps of row ky and walks 8 x 32 pixel tiles: the input rows the
//   tile's output rows meet through filter row ky (producer BatchNorm + LeakyReLU and the padding applied while staging) and
//   the dy tile (pixel-major) in shared memory. A warp visits every 8th pixel; its 32 threads are 4 output-channel quads x
//   8 input channels, each with 4 x K accumulators: one 128-bit load of dy + K loads of x for 4*K FMAs (stride 1: two adjacent
//   pixels per visit, two loads of dy + K + 1 loads of x for 8*K FMAs).
// -------------------------------------------------------------------------------------------------
template <int K, int S>
static __global__ void __launch_bounds__(256)
convx_wgrad_kernel(const float* __restrict__ x, int Cin, int Hin, int Win, InTf tf, const float* __restrict__ dy, int Cout, int Ho,
                   int Wo, int N, int reflect, float* __restrict__ part) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    constexpr int CO_T = 16, CI_T = 8, PAD = (K - 1) / 2, KK = K * K;
    constexpr int IW = (TW - 1) * S + K, IWP = IW + 1;
    constexpr int SXA = (CI_T * TH * IWP + 3) & ~3;
    constexpr int RS = 4 * K + 4;
    constexpr int PXW = (S == 1 && K >= 3) ? 2 : 1;          // pixels per visit of the accumulation loop
    constexpr int SM_A = SXA + TH * TW * XDYP, SM_B = 8 * 32 * RS;
    __shared__ __align__(16) float smem[SM_A > SM_B ? SM_A : SM_B];
    float (*s_x)[TH][IWP] = reinterpret_cast<float (*)[TH][IWP]>(smem);
    float* s_dy = smem + SXA;                             // [TH*TW][XDYP], 16-byte aligned rows
    const int ci_tiles = (Cin + CI_T - 1) / CI_T;
    const int co0 = ((int)blockIdx.y / ci_tiles) * CO_T;
    const int ci0 = ((int)blockIdx.y % ci_tiles) * CI_T;
    const int ky = blockIdx.z;
    const int ext = reflect ? PAD : 0;
    const int g = threadIdx.x >> 5, w = threadIdx.x & 31;   // pixel group, weight thread
    const int cos = (w / CI_T) * 4, ci = w % CI_T;          // this thread: 4 output channels x 1 input channel x K taps
    const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
    const int ntiles = N * tiles_y * tiles_x;
    float acc[4][K];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) acc[j][kx] = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / (tiles_y * tiles_x), tr = tile % (tiles_y * tiles_x);
        const int ty0 = (tr / tiles_x) * TH, tx0 = (tr % tiles_x) * TW;
        for (int idx = threadIdx.x; idx < CI_T * TH * IW; idx += 256) {
            const int c = idx / (TH * IW), r = (idx / IW) % TH, q = idx % IW;
            int iy = (ty0 + r) * S - PAD + ky, ix = tx0 * S - PAD + q;
            float v = 0.f;
            if (ci0 + c < Cin && iy >= -ext && iy < Hin + ext && ix >= -ext && ix < Win + ext) {
                iy = reflect_idx(iy, Hin); ix = reflect_idx(ix, Win);   // identity inside the image (always, when ext == 0)
                v = apply_tf(tf, ci0 + c, x[((size_t)(n * Cin + ci0 + c) * Hin + iy) * Win + ix]);
            }
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
            *reinterpret_cast<float4*>(s_dy + pix * XDYP + c4) = v;
        }
        __syncthreads();
        if (PXW == 2) {
            // stride 1: two horizontally adjacent pixels per visit share K - 1 of their K input columns
            for (int p2 = g; p2 < TH * TW / 2; p2 += 8) {
                const int py = p2 / (TW / 2), px = (p2 % (TW / 2)) * 2;
                const float4 da = *reinterpret_cast<const float4*>(s_dy + (py * TW + px) * XDYP + cos);
                const float4 db = *reinterpret_cast<const float4*>(s_dy + (py * TW + px + 1) * XDYP + cos);
                const float d0[4] = {da.x, da.y, da.z, da.w}, d1[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) bacc[j] += d0[j] + d1[j];
                float xr[K + 1];
#pragma unroll
                for (int c = 0; c < K + 1; ++c) xr[c] = s_x[ci][py][px + c];
#pragma unroll
                for (int kx = 0; kx < K; ++kx)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][kx] = fmaf(d1[j], xr[kx + 1], fmaf(d0[j], xr[kx], acc[j][kx]));
            }
        } else {
            for (int p = g; p < TH * TW; p += 8) {
                const int py = p / TW, px = p % TW;
                const float4 d4 = *reinterpret_cast<const float4*>(s_dy + p * XDYP + cos);
                const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) bacc[j] += d[j];
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float xv = s_x[ci][py][px * S + kx];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][kx] = fmaf(d[j], xv, acc[j][kx]);
                }
            }
        }
        __syncthreads();
    }
    // reduce over the 8 pixel groups through shared memory, then one writer per weight
    float* red = smem;   // [8][32][4*K + 4]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) red[(g * 32 + w) * RS + j * K + kx] = acc[j][kx];
        red[(g * 32 + w) * RS + 4 * K + j] = bacc[j];
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
        if (e < 4 * K) {
            const int j = e / K, kx = e % K;
            const int co = co0 + wcos + j, c = ci0 + wci;
            if (co < Cout && c < Cin) out[((size_t)co * Cin + c) * KK + ky * K + kx] = sv;
        } else if (ci0 == 0 && wci == 0 && ky == 0) {
            const int co = co0 + wcos + (e - 4 * K);
            if (co < Cout) out[nW + co] = sv;
        }
    }
}

// nn.BatchNorm2d's running statistics (momentum 0.1, unbiased variance, num_batches_tracked += 1). One block per layer.
struct RunningTableX {
    const float2* bstat[GENX_MAX_BN];
    float* rmean[GENX_MAX_BN];
    float* rvar[GENX_MAX_BN];
    long long* nbt[GENX_MAX_BN];
    int C[GENX_MAX_BN];
};
static __global__ void __launch_bounds__(GENX_MAX_CH) update_running_x_kernel(RunningTableX t, float momentum) {
    pdl_sync();   // programmatic dependent launch: scheduled under the previous kernel's tail, waits for its completion here
    const int l = blockIdx.x, c = threadIdx.x;
    if (c < t.C[l]) {
        const float2 b = t.bstat[l][c];
        t.rmean[l][c] = (1.f - momentum) * t.rmean[l][c] + momentum * b.x;
        t.rvar[l][c] = (1.f - momentum) * t.rvar[l][c] + momentum * b.y;
    }
    if (c == 0) *t.nbt[l] += 1;
}

// -------------------------------------------------------------------------------------------------
// host: launch helpers
// -------------------------------------------------------------------------------------------------
static inline int chan_tile(int c) { return c <= 4 ? 4 : (c <= 8 ? 8 : 16); }
static inline int px_per_thread(int K, int S) { return (S == 1 && K >= 5) ? 2 : 1; }   // == XPx<K, S>::value
static inline int elementwise_blocks(size_t total, int per_sm) {
    const size_t b = (total + 255) / 256, cap = (size_t)148 * per_sm;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

#define XDISPATCH_KS(M)                                                                        \
    do {                                                                                       \
        if (K == 1 && S == 1) M(1, 1);                                                         \
        else if (K == 1 && S == 2) M(1, 2);                                                    \
        else if (K == 3 && S == 1) M(3, 1);                                                    \
        else if (K == 3 && S == 2) M(3, 2);                                                    \
        else if (K == 5 && S == 1) M(5, 1);                                                    \
        else if (K == 5 && S == 2) M(5, 2);                                                    \
        else if (K == 7 && S == 1) M(7, 1);                                                    \
        else if (K == 7 && S == 2) M(7, 2);                                                    \
        else { set_error("generator: unsupported conv k=%d stride=%d", K, S); return SPLICE_ERR_UNSUPPORTED; } \
    } while (0)

// Split factor of a layer's reduction channels: grids that would not fill 148 SMs twice over (low-resolution layers: a few hundred
// pixels x 128 channels) are widened by splitting the reduction, at least 8 channels per split, partials within `cap` floats.
static constexpr int XTARGET_BLOCKS = 296;
static int pick_split(int P, int c_fast, int ct, int c_slow, size_t out_elems, size_t cap) {
    const int blocks = ceil_div(P, XCONV_THREADS) * ceil_div(c_fast, ct);
    if (blocks >= XTARGET_BLOCKS || P > 16384) return 1;
    int sk = ceil_div(XTARGET_BLOCKS, blocks);
    const int max_sk = c_slow / 8 > 0 ? c_slow / 8 : 1;
    if (sk > max_sk) sk = max_sk;
    if (sk > 16) sk = 16;
    while (sk > 1 && (size_t)sk * out_elems > cap) --sk;
    return sk;
}

// split / split_cap: scratch for the partial sums of a split reduction (floats); only layers with a BatchNorm epilogue split
static int launch_convx_fwd(int K, int S, const float* x, int N, int Cin, int Hin, int Win, InTf tf, const float* Wt, const float* bias,
                            int Cout, float* y, int Ho, int Wo, int reflect, int sigmoid, float* stats_part, BnFin fin, float* split,
                            size_t split_cap, cudaStream_t st) {
    const int ct = chan_tile(Cout);
    const size_t out_elems = (size_t)N * Cout * Ho * Wo;
    const int P = N * Ho * ceil_div(Wo, px_per_thread(K, S));          // threads: one per pixel group (XPx)
    const int sk = (fin.konst && split) ? pick_split(P, Cout, ct, Cin, out_elems, split_cap) : 1;
    dim3 grid(ceil_div(P, XCONV_THREADS), ceil_div(Cout, ct), sk);
    float* dst = sk > 1 ? split : y;
    float* stats = sk > 1 ? nullptr : stats_part;
    BnFin fin_conv = fin;
    if (sk > 1) fin_conv.konst = nullptr;   // the split epilogue kernel does the statistics
#define XF(KK, SS, CT) SPLICE_CHECK_CUDA(launch_pdl(convx_fwd_kernel<KK, SS, CT>, grid, dim3(XCONV_THREADS), 0, st, x, N, Cin, Hin, Win, tf, Wt, bias, Cout, dst, Ho, Wo, reflect, sigmoid, stats, sk, fin_conv))
#define XF3(KK, SS) do { if (ct == 4) XF(KK, SS, 4); else if (ct == 8) XF(KK, SS, 8); else XF(KK, SS, 16); } while (0)
    XDISPATCH_KS(XF3);
#undef XF3
#undef XF
    SPLICE_LAUNCH_CHECK();
    if (sk > 1) {
        dim3 fgrid(ceil_div(N * Ho * Wo, 1024), Cout);
        SPLICE_CHECK_CUDA(launch_pdl(conv_finish_stats_kernel, fgrid, dim3(256), 0, st, (const float*)split, sk, bias, y, N, Cout, Ho * Wo, stats_part, fin));
        SPLICE_LAUNCH_CHECK();
    }
    return SPLICE_OK;
}

// d(transformed conv input) [N,Cin,Hin,Win] (=|+=) from dy [N,Cout,Ho,Wo]; dpad: scratch for the padded domain (reflection)
static int launch_convx_dgrad(int K, int S, const float* dy, int N, int Cout, int Ho, int Wo, const float* Wt, int Cin, float* dX, int Hin,
                              int Win, int reflect, int accumulate, float* dpad, float* split, size_t split_cap, cudaStream_t st) {
    const int ext = (reflect && K > 1) ? (K - 1) / 2 : 0;
    const int Hq = Hin + 2 * ext, Wq = Win + 2 * ext;
    float* dst = ext ? dpad : dX;                       // where the (summed) data gradient of the domain goes
    const int acc_k = ext ? 0 : accumulate;
    const int ct = chan_tile(Cin);
    const size_t total_q = (size_t)N * Cin * Hq * Wq;
    const int P = N * Hq * ceil_div(Wq, px_per_thread(K, S));          // threads: one per pixel group (XPx)
    const int sk = pick_split(P, Cin, ct, Cout, total_q, split_cap);
    dim3 grid(ceil_div(P, XCONV_THREADS), ceil_div(Cin, ct), sk);
    float* kdst = sk > 1 ? split : dst;
#define XD(KK, SS, CT) SPLICE_CHECK_CUDA(launch_pdl(convx_dgrad_kernel<KK, SS, CT>, grid, dim3(XCONV_THREADS), 0, st, dy, N, Cout, Ho, Wo, Wt, Cin, kdst, Hq, Wq, ext, acc_k, sk))
#define XD3(KK, SS) do { if (ct == 4) XD(KK, SS, 4); else if (ct == 8) XD(KK, SS, 8); else XD(KK, SS, 16); } while (0)
    XDISPATCH_KS(XD3);
#undef XD3
#undef XD
    SPLICE_LAUNCH_CHECK();
    if (sk > 1) {
        SPLICE_CHECK_CUDA(launch_pdl(sum_partials_kernel, dim3(elementwise_blocks(total_q, 8)), dim3(256), 0, st, (const float*)split, sk, total_q, dst, acc_k));
        SPLICE_LAUNCH_CHECK();
    }
    if (ext) {
        const size_t total = (size_t)N * Cin * Hin * Win;
        SPLICE_CHECK_CUDA(launch_pdl(reflect_fold_kernel, dim3(elementwise_blocks(total, 8)), dim3(256), 0, st, (const float*)dpad, Hin, Win, ext,
                                     total, dX, accumulate));
        SPLICE_LAUNCH_CHECK();
    }
    return SPLICE_OK;
}

static int launch_convx_wgrad(int K, int S, const float* x, int N, int Cin, int Hin, int Win, InTf tf, const float* dy, int Cout, int Ho,
                              int Wo, int reflect, float* wpart, float* gw, float* gb, int accumulate, cudaStream_t st) {
    const int ntiles = N * ceil_div(Ho, TH) * ceil_div(Wo, TW);
    const size_t nW = (size_t)Cout * Cin * K * K;
    size_t cap = XWGRAD_FLOATS / (nW + Cout);
    if (cap > 256) cap = 256;
    SPLICE_REQUIRE(cap >= 1, "generator: weight tensor %d x %d x %d x %d exceeds the weight-gradient scratch", Cout, Cin, K, K);
    const int chunks = ntiles < (int)cap ? ntiles : (int)cap;
    dim3 grid(chunks, ceil_div(Cout, 16) * ceil_div(Cin, 8), K);
    const int refl = (reflect && K > 1) ? 1 : 0;
#define XW(KK, SS) SPLICE_CHECK_CUDA(launch_pdl(convx_wgrad_kernel<KK, SS>, grid, dim3(256), 0, st, x, Cin, Hin, Win, tf, dy, Cout, Ho, Wo, N, refl, wpart))
    XDISPATCH_KS(XW);
#undef XW
    SPLICE_LAUNCH_CHECK();
    SPLICE_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(ceil_div((int)(nW + Cout), 256)), dim3(256), 0, st, (const float*)wpart, chunks, nW, Cout, gw,
                                 gb, accumulate));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

// -------------------------------------------------------------------------------------------------
// engine
// -------------------------------------------------------------------------------------------------
static bool filter_ok(int k) { return k == 1 || k == 3 || k == 5 || k == 7; }

int GenXEngine::create(const GenXConfig& c, GenXEngine** out) {
    SPLICE_REQUIRE(out, "generator: null output");
    SPLICE_REQUIRE(c.n_scales >= 1 && c.n_scales <= GENX_MAX_SCALES, "generator: n_scales %d outside [1, %d]", c.n_scales, GENX_MAX_SCALES);
    SPLICE_REQUIRE(c.in_channels >= 1 && c.in_channels <= GENX_MAX_CH && c.out_channels >= 1 && c.out_channels <= 16,
                   "generator: %d input / %d output channels unsupported (<= %d / <= 16)", c.in_channels, c.out_channels, GENX_MAX_CH);
    SPLICE_REQUIRE(filter_ok(c.k_skip), "generator: filter_skip_size %d unsupported (1, 3, 5, 7)", c.k_skip);
    for (int i = 0; i < c.n_scales; ++i) {
        SPLICE_REQUIRE(filter_ok(c.k_down[i]) && filter_ok(c.k_up[i]), "generator: filter sizes %d / %d at scale %d unsupported (1, 3, 5, 7)",
                       c.k_down[i], c.k_up[i], i);
        SPLICE_REQUIRE(c.ch_skip[i] >= 1 && c.ch_down[i] >= 1 && c.ch_up[i] >= 1, "generator: scale %d needs >= 1 skip / down / up channels", i);
        const int cdeep = (i == c.n_scales - 1) ? c.ch_down[i] : c.ch_up[i + 1];
        SPLICE_REQUIRE(c.ch_skip[i] + cdeep <= GENX_MAX_CH && c.ch_down[i] <= GENX_MAX_CH && c.ch_up[i] <= GENX_MAX_CH,
                       "generator: more than %d channels at scale %d", GENX_MAX_CH, i);
    }
    GenXEngine* e = new GenXEngine();
    e->cfg_ = c;
    const int ns = e->ns_ = c.n_scales;
    for (int i = 0; i < ns; ++i) {
        Scale& s = e->sc_[i];
        const int cin = (i == 0) ? c.in_channels : c.ch_down[i - 1];
        const int cd = c.ch_down[i], cu = c.ch_up[i], cs = c.ch_skip[i];
        // netG.parameters() / BatchNorm module order: [skip conv, bn | down conv 1, bn | down conv 2, bn] of every scale on the way
        // down, then [concat bn | up conv, bn | 1x1 conv, bn] on the way back up, the final conv last
        const int pre = 12 * i, post = 12 * ns + 10 * (ns - 1 - i);
        const int bpre = 3 * i, bpost = 3 * ns + 3 * (ns - 1 - i);
        s.cskip = cs;
        s.cdeep = (i == ns - 1) ? cd : c.ch_up[i + 1];
        s.s = Conv{cin, cs, c.k_skip, 1, pre + 0, pre + 1};
        s.bs = Bn{cs, pre + 2, pre + 3, bpre + 0};
        s.d1 = Conv{cin, cd, c.k_down[i], 2, pre + 4, pre + 5};
        s.bd1 = Bn{cd, pre + 6, pre + 7, bpre + 1};
        s.d2 = Conv{cd, cd, c.k_down[i], 1, pre + 8, pre + 9};
        s.bd2 = Bn{cd, pre + 10, pre + 11, bpre + 2};
        s.bcat = Bn{cs + s.cdeep, post + 0, post + 1, bpost + 0};
        s.c1 = Conv{cs + s.cdeep, cu, c.k_up[i], 1, post + 2, post + 3};
        s.bc1 = Bn{cu, post + 4, post + 5, bpost + 1};
        s.c2 = Conv{cu, cu, 1, 1, post + 6, post + 7};
        s.bc2 = Bn{cu, post + 8, post + 9, bpost + 2};
        const int widths[4] = {cs + s.cdeep, cd, cu, cs};
        for (int k = 0; k < 4; ++k) if (widths[k] > e->max_c_) e->max_c_ = widths[k];
        if (cs > e->max_cskip_) e->max_cskip_ = cs;
    }
    e->final_ = Conv{c.ch_up[0], c.out_channels, 1, 1, 22 * ns, 22 * ns + 1};
    *out = e;
    return SPLICE_OK;
}

GenXEngine::~GenXEngine() {
    cudaFree(pool_);
    cudaFree(scratch_);
    cudaFree(counters_);
    if (side_) cudaStreamDestroy(side_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
}

int GenXEngine::bind(float* const* params, float* const* grads, float* const* running_mean, float* const* running_var,
                     long long* const* num_batches_tracked) {
    SPLICE_REQUIRE(params, "generator: null parameter table");
    for (int i = 0; i < n_params(); ++i) {
        SPLICE_REQUIRE(params[i], "generator: parameter %d is null", i);
        param_[i] = params[i];
        grad_[i] = grads ? grads[i] : nullptr;
    }
    have_grads_ = grads != nullptr;
    if (have_grads_)
        for (int i = 0; i < n_params(); ++i) SPLICE_REQUIRE(grad_[i], "generator: gradient %d is null", i);
    have_running_ = running_mean && running_var && num_batches_tracked;
    for (int i = 0; i < n_bn(); ++i) {
        rmean_[i] = have_running_ ? running_mean[i] : nullptr;
        rvar_[i] = have_running_ ? running_var[i] : nullptr;
        nbt_[i] = have_running_ ? num_batches_tracked[i] : nullptr;
        if (have_running_) SPLICE_REQUIRE(rmean_[i] && rvar_[i] && nbt_[i], "generator: BN buffer %d is null", i);
    }
    bound_ = true;
    return SPLICE_OK;
}

int GenXEngine::configure(int N, int H, int W) {
    if (!side_) {
        SPLICE_CHECK_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    }
    if (!counters_) {
        SPLICE_CHECK_CUDA(cudaMalloc(&counters_, (size_t)GENX_MAX_BN * GENX_MAX_CH * sizeof(int)));
        SPLICE_CHECK_CUDA(cudaMemset(counters_, 0, (size_t)GENX_MAX_BN * GENX_MAX_CH * sizeof(int)));
    }
    if (pool_ && N_ == N && H_ == H && W_ == W) return SPLICE_OK;
    int hs[GENX_MAX_SCALES], ws[GENX_MAX_SCALES], hd[GENX_MAX_SCALES], wd[GENX_MAX_SCALES];
    int h = H, w = W;
    size_t dpad = 0;
    for (int i = 0; i < ns_; ++i) {
        hs[i] = h; ws[i] = w; hd[i] = (h + 1) / 2; wd[i] = (w + 1) / 2;
        if (cfg_.reflect) {
            // nn.ReflectionPad2d(p) needs p < size; and the padded-domain scratch of the data gradients that fold a border back
            const int p_in = ((cfg_.k_down[i] > cfg_.k_up[i] ? cfg_.k_down[i] : cfg_.k_up[i]) - 1) / 2, p_sk = (cfg_.k_skip - 1) / 2;
            const int p_dn = (cfg_.k_down[i] - 1) / 2;
            const int pmax = p_in > p_sk ? p_in : p_sk;
            SPLICE_REQUIRE(h > pmax && w > pmax && hd[i] > p_dn && wd[i] > p_dn,
                           "generator: input %dx%d too small for reflection padding at scale %d (%dx%d)", H, W, i, h, w);
            const Scale& c = sc_[i];
            auto need = [&](int cin, int hh, int ww, int k) {
                const int p = (k - 1) / 2;
                if (p == 0) return;
                const size_t f = (size_t)N * cin * (hh + 2 * p) * (ww + 2 * p);
                if (f > dpad) dpad = f;
            };
            need(c.c1.cin, h, w, c.c1.k);
            need(c.d2.cin, hd[i], wd[i], c.d2.k);
            if (i > 0) { need(c.d1.cin, h, w, c.d1.k); need(c.s.cin, h, w, c.s.k); }
        }
        h = hd[i]; w = wd[i];
    }
    // nn.BatchNorm2d in training mode refuses a single value per channel ("Expected more than 1 value per channel when training")
    SPLICE_REQUIRE((size_t)N * hd[ns_ - 1] * wd[ns_ - 1] > 1, "generator: input %dx%d too small: one value per channel at the deepest scale "
                   "(BatchNorm in training mode needs more than 1)", H, W);
    const size_t io_in = (size_t)N * cfg_.in_channels * H * W, io_out = (size_t)N * cfg_.out_channels * H * W;

    size_t off = 0;
    std::vector<size_t> offs;
    auto plan = [&](size_t bytes) { offs.push_back(off); off += (bytes + 255) & ~(size_t)255; };
    for (int i = 0; i < ns_; ++i) {
        const Scale& c = sc_[i];
        const size_t px = (size_t)N * hs[i] * ws[i], pd = (size_t)N * hd[i] * wd[i];
        const int ch[6] = {c.cskip, c.d1.cout, c.d2.cout, c.cskip + c.cdeep, c.c1.cout, c.c2.cout};
        const size_t sz[6] = {px * ch[0], pd * ch[1], pd * ch[2], px * ch[3], px * ch[4], px * ch[5]};
        for (int k = 0; k < 6; ++k) plan(sz[k] * 4);   // raw tensors
        for (int k = 0; k < 6; ++k) plan(sz[k] * 4);   // gradients w.r.t. the activated / normalised tensors
        for (int k = 0; k < 6; ++k) plan((size_t)ch[k] * sizeof(float4));
        for (int k = 0; k < 6; ++k) plan((size_t)ch[k] * sizeof(float2));
    }
    plan(io_in * 4);    // x copy (the caller's tensor may be gone before backward(): wgrad of scale 0 reads it)
    plan(io_out * 4);   // out copy
    plan(io_out * 4);   // d(pre-sigmoid)
    plan(io_out * 4);   // dout copy (stable address for the backward graph)
    plan((size_t)GENX_MAX_BN * GENX_MAX_CH * sizeof(float2));   // batch statistics (mean, unbiased variance) per BN layer
    if (off > pool_bytes_) {
        SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(pool_);
        pool_ = nullptr;
        pool_bytes_ = 0;
        SPLICE_CHECK_CUDA(cudaMalloc(&pool_, off));
        pool_bytes_ = off;
    }
    uint8_t* base = static_cast<uint8_t*>(pool_);
    size_t k = 0;
    auto nx = [&]() { return base + offs[k++]; };
    for (int i = 0; i < ns_; ++i) {
        ScaleBuf& b = sb_[i];
        b.h = hs[i]; b.w = ws[i]; b.hd = hd[i]; b.wd = wd[i];
        b.s_raw = (float*)nx(); b.d1_raw = (float*)nx(); b.d2_raw = (float*)nx(); b.cat = (float*)nx(); b.c1_raw = (float*)nx(); b.c2_raw = (float*)nx();
        b.dA_s = (float*)nx(); b.dA_d1 = (float*)nx(); b.dA_d2 = (float*)nx(); b.dcat = (float*)nx(); b.dA_c1 = (float*)nx(); b.dA_c2 = (float*)nx();
        b.k_s = (float4*)nx(); b.k_d1 = (float4*)nx(); b.k_d2 = (float4*)nx(); b.k_cat = (float4*)nx(); b.k_c1 = (float4*)nx(); b.k_c2 = (float4*)nx();
        b.m_s = (float2*)nx(); b.m_d1 = (float2*)nx(); b.m_d2 = (float2*)nx(); b.m_cat = (float2*)nx(); b.m_c1 = (float2*)nx(); b.m_c2 = (float2*)nx();
    }
    x_copy_ = (float*)nx();
    out_ = (float*)nx();
    dfin_ = (float*)nx();
    dout_copy_ = (float*)nx();
    bstat_ = (float2*)nx();

    // scratch = [statistics partials | skip-branch statistics partials | split partial sums | skip-branch split partial sums |
    //            weight-gradient partials | padded data gradient]
    const size_t conv_blocks = (size_t)ceil_div(N * H * W, XCONV_THREADS);
    stats_floats_ = (conv_blocks + (size_t)N * ceil_div(H, TH) * ceil_div(W, TW)) * max_c_ * 3;
    skip_floats_ = conv_blocks * max_cskip_ * 3;
    dpad_floats_ = dpad;
    const size_t bytes = (stats_floats_ + skip_floats_ + XSPLIT_FLOATS + XSPLIT_SKIP_FLOATS + XWGRAD_FLOATS + dpad_floats_) * sizeof(float) + 4096;
    if (bytes > scratch_bytes_) {
        SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(scratch_);
        scratch_ = nullptr;
        scratch_bytes_ = 0;
        SPLICE_CHECK_CUDA(cudaMalloc(&scratch_, bytes));
        scratch_bytes_ = bytes;
    }
    N_ = N; H_ = H; W_ = W;
    valid_ = false;
    return SPLICE_OK;
}

#define GRC(expr)              \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

int GenXEngine::forward(const float* x, int N, int H, int W, float* out, bool keep, bool update_running_stats, cudaStream_t st) {
    NvtxRange nvtx("splice_genx_forward");
    SPLICE_REQUIRE(bound_, "generator: splice_genx_bind has not been called");
    SPLICE_REQUIRE(x && out && N > 0 && H > 0 && W > 0, "generator: bad input");
    SPLICE_REQUIRE(!update_running_stats || have_running_, "generator: running statistics requested but no BatchNorm buffers are bound");
    GRC(configure(N, H, W));
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(x_copy_, x, (size_t)N * cfg_.in_channels * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (use_graphs_) {
        KeyHasher k;
        k.add((uint64_t)21).add((uint64_t)N).add((uint64_t)H).add((uint64_t)W).add(pool_).add(scratch_);
        for (int i = 0; i < n_params(); ++i) k.add(param_[i]);
        GRC(graphs_.run(k.h, st, [&](cudaStream_t cs) { return forward_body(cs); }));
    } else {
        GRC(forward_body(st));
    }
    if (update_running_stats) GRC(update_running(st));
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(out, out_, (size_t)N * cfg_.out_channels * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    valid_ = keep;
    return SPLICE_OK;
}

int GenXEngine::update_running(cudaStream_t st) {
    RunningTableX t;
    const Bn* order[GENX_MAX_BN];
    for (int i = 0; i < ns_; ++i) {
        const Scale& c = sc_[i];
        const Bn* b[6] = {&c.bs, &c.bd1, &c.bd2, &c.bcat, &c.bc1, &c.bc2};
        for (int k = 0; k < 6; ++k) order[b[k]->idx] = b[k];
    }
    for (int l = 0; l < GENX_MAX_BN; ++l) {
        const bool on = l < n_bn();
        t.bstat[l] = on ? bstat_ + (size_t)l * GENX_MAX_CH : nullptr;
        t.rmean[l] = on ? rmean_[l] : nullptr; t.rvar[l] = on ? rvar_[l] : nullptr; t.nbt[l] = on ? nbt_[l] : nullptr;
        t.C[l] = on ? order[l]->c : 0;
    }
    SPLICE_CHECK_CUDA(launch_pdl(update_running_x_kernel, dim3(n_bn()), dim3(GENX_MAX_CH), 0, st, t, 0.1f));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int GenXEngine::forward_body(cudaStream_t st) {
    const int N = N_, H = H_, W = W_, refl = cfg_.reflect;
    float* part = static_cast<float*>(scratch_);
    float* part_skip = part + stats_floats_;
    float* split = part_skip + skip_floats_;
    float* split_skip = split + XSPLIT_FLOATS;
    const float eps = 1e-5f;

    auto bn_fin = [&](const Bn& b, float4* k) {
        return BnFin{param_[b.pg], param_[b.pb], k, bstat_ + (size_t)b.idx * GENX_MAX_CH, counters_ + (size_t)b.idx * GENX_MAX_CH, eps};
    };
    auto conv_bn = [&](const Conv& c, const Bn& b, const float* in, int hin, int win, InTf tf, float* y, int ho, int wo, float4* k,
                       float* stats, cudaStream_t cs) -> int {
        const bool on_side = cs == side_ && stats == part_skip;
        return launch_convx_fwd(c.k, c.stride, in, N, c.cin, hin, win, tf, param_[c.pw], param_[c.pb], c.cout, y, ho, wo, refl, 0, stats,
                                bn_fin(b, k), on_side ? split_skip : split, on_side ? XSPLIT_SKIP_FLOATS : XSPLIT_FLOATS, cs);
    };

    // down path. The skip convolutions only feed the concats of the up path: they run on the side stream (a parallel branch
    // of the captured graph) with their own statistics scratch.
    const float* in = x_copy_;
    InTf tf_in{nullptr, 0};
    for (int i = 0; i < ns_; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = sb_[i];
        SPLICE_CHECK_CUDA(cudaEventRecord(ev_fork_, st));
        SPLICE_CHECK_CUDA(cudaStreamWaitEvent(side_, ev_fork_, 0));
        GRC(conv_bn(c.s, c.bs, in, b.h, b.w, tf_in, b.s_raw, b.h, b.w, b.k_s, part_skip, side_));
        GRC(conv_bn(c.d1, c.bd1, in, b.h, b.w, tf_in, b.d1_raw, b.hd, b.wd, b.k_d1, part, st));
        GRC(conv_bn(c.d2, c.bd2, b.d1_raw, b.hd, b.wd, InTf{b.k_d1, 1}, b.d2_raw, b.hd, b.wd, b.k_d2, part, st));
        in = b.d2_raw;
        tf_in = InTf{b.k_d2, 1};
    }
    SPLICE_CHECK_CUDA(cudaEventRecord(ev_join_, side_));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join_, 0));
    // up path
    for (int i = ns_ - 1; i >= 0; --i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = sb_[i];
        const float* u = (i == ns_ - 1) ? b.d2_raw : sb_[i + 1].c2_raw;
        InTf tf_u = (i == ns_ - 1) ? InTf{b.k_d2, 1} : InTf{sb_[i + 1].k_c2, 1};
        const int hu = b.hd, wu = b.wd;                                      // == the deeper scale's size
        const int th = min(b.h, 2 * hu), tw = min(b.w, 2 * wu);              // Concat crops to the smaller size
        SPLICE_REQUIRE(th == b.h && tw == b.w, "generator: unexpected concat geometry");
        const int oyu = (2 * hu - th) / 2, oxu = (2 * wu - tw) / 2;
        const int C = c.cskip + c.cdeep;
        dim3 grid(ceil_div(th, TH) * ceil_div(tw, TW), C, N);
        SPLICE_CHECK_CUDA(launch_pdl(cat_build_kernel, grid, dim3(256), 0, st, (const float*)b.s_raw, c.cskip, b.h, b.w, InTf{b.k_s, 1}, 0, 0, u, c.cdeep, hu,
                                     wu, tf_u, oyu, oxu, b.cat, th, tw, part, bn_fin(c.bcat, b.k_cat)));
        SPLICE_LAUNCH_CHECK();
        GRC(conv_bn(c.c1, c.bc1, b.cat, th, tw, InTf{b.k_cat, 0}, b.c1_raw, th, tw, b.k_c1, part, st));
        GRC(conv_bn(c.c2, c.bc2, b.c1_raw, th, tw, InTf{b.k_c1, 1}, b.c2_raw, th, tw, b.k_c2, part, st));
    }
    GRC(launch_convx_fwd(1, 1, sb_[0].c2_raw, N, final_.cin, H, W, InTf{sb_[0].k_c2, 1}, param_[final_.pw], param_[final_.pb], final_.cout, out_, H, W,
                         0, cfg_.sigmoid ? 1 : 0, nullptr, BnFin{nullptr, nullptr, nullptr, nullptr, nullptr, eps}, nullptr, 0, st));
    return SPLICE_OK;
}

int GenXEngine::backward(const float* dout, bool accumulate, cudaStream_t st) {
    NvtxRange nvtx("splice_genx_backward");
    SPLICE_REQUIRE(pool_ && valid_, "generator backward: no kept forward pass");
    SPLICE_REQUIRE(dout, "generator backward: null gradient");
    SPLICE_REQUIRE(have_grads_, "generator backward: no gradient table is bound");
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(dout_copy_, dout, (size_t)N_ * cfg_.out_channels * H_ * W_ * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (use_graphs_) {
        KeyHasher k;
        k.add((uint64_t)22).add((uint64_t)N_).add((uint64_t)H_).add((uint64_t)W_).add(pool_).add(scratch_).add((uint64_t)accumulate);
        for (int i = 0; i < n_params(); ++i) k.add(param_[i]).add(grad_[i]);
        GRC(graphs_.run(k.h, st, [&](cudaStream_t cs) { return backward_body(accumulate, cs); }));
    } else {
        GRC(backward_body(accumulate, st));
    }
    valid_ = false;
    return SPLICE_OK;
}

int GenXEngine::backward_body(bool accumulate, cudaStream_t st) {
    const int N = N_, H = H_, W = W_, refl = cfg_.reflect;
    const int acc = accumulate ? 1 : 0;
    float* part = static_cast<float*>(scratch_);
    float* split = part + stats_floats_ + skip_floats_;
    float* wpart = split + XSPLIT_FLOATS + XSPLIT_SKIP_FLOATS;
    float* dpad = wpart + XWGRAD_FLOATS;

    // BatchNorm(+LeakyReLU) backward of one layer: afterwards dA holds d(raw conv output)
    auto bn_bwd = [&](const Bn& b, float* dA, const float* y, const float4* k, int lrelu, int hh, int ww, float2* m) -> int {
        const int HW = hh * ww;
        if ((size_t)N * HW <= 8192) {
            SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_small_kernel, dim3(b.c), dim3(256), 0, st, dA, y, k, lrelu, N, b.c, HW, grad_[b.pg], grad_[b.pb], acc));
            SPLICE_LAUNCH_CHECK();
            return SPLICE_OK;
        }
        dim3 grid(ceil_div(HW, 2048), b.c, N);
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_reduce_kernel, grid, dim3(256), 0, st, (const float*)dA, y, k, lrelu, b.c, HW, part));
        SPLICE_LAUNCH_CHECK();
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3(b.c), dim3(32), 0, st, (const float*)part, N * (int)grid.x, b.c, (double)N * HW, grad_[b.pg], grad_[b.pb], m, acc));
        SPLICE_LAUNCH_CHECK();
        const size_t total = (size_t)N * b.c * HW;
        SPLICE_CHECK_CUDA(launch_pdl(bn_bwd_apply_kernel, dim3(elementwise_blocks(total, 16)), dim3(256), 0, st, dA, y, k, (const float2*)m, lrelu, b.c, HW, total));
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    };
    // Weight gradients hang off the critical path (bn_bwd -> dgrad -> bn_bwd -> ...): each one only needs this layer's dy,
    // which is complete on `st` when wgrad() is called, so they run on the side stream (a parallel branch of the captured
    // graph) and are joined at the end of the pass. The side stream is in order, so `wpart` is reused safely.
    cudaStream_t ws = side_;
    auto wgrad = [&](const Conv& c, const float* in, int hin, int win, InTf tf, const float* dy, int ho, int wo) -> int {
        SPLICE_CHECK_CUDA(cudaEventRecord(ev_fork_, st));
        SPLICE_CHECK_CUDA(cudaStreamWaitEvent(ws, ev_fork_, 0));
        return launch_convx_wgrad(c.k, c.stride, in, N, c.cin, hin, win, tf, dy, c.cout, ho, wo, refl, wpart, grad_[c.pw], grad_[c.pb], acc, ws);
    };
    auto dgrad = [&](const Conv& c, const float* dy, int ho, int wo, float* dX, int hin, int win, int accumulate_dx) -> int {
        return launch_convx_dgrad(c.k, c.stride, dy, N, c.cout, ho, wo, param_[c.pw], c.cin, dX, hin, win, refl, accumulate_dx, dpad, split, XSPLIT_FLOATS,
                                  st);
    };

    // final 1x1 conv (+ sigmoid)
    {
        const size_t total = (size_t)N * cfg_.out_channels * H * W;
        const float* dfin = dout_copy_;
        if (cfg_.sigmoid) {
            SPLICE_CHECK_CUDA(launch_pdl(sigmoid_bwd_kernel, dim3(elementwise_blocks(total, 8)), dim3(256), 0, st, (const float*)dout_copy_, (const float*)out_, dfin_, total));
            SPLICE_LAUNCH_CHECK();
            dfin = dfin_;
        }
        GRC(wgrad(final_, sb_[0].c2_raw, H, W, InTf{sb_[0].k_c2, 1}, dfin, H, W));
        GRC(dgrad(final_, dfin, H, W, sb_[0].dA_c2, H, W, 0));
    }
    // up path, top to bottom
    for (int i = 0; i < ns_; ++i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = sb_[i];
        const int h = b.h, w = b.w;
        GRC(bn_bwd(c.bc2, b.dA_c2, b.c2_raw, b.k_c2, 1, h, w, b.m_c2));
        GRC(wgrad(c.c2, b.c1_raw, h, w, InTf{b.k_c1, 1}, b.dA_c2, h, w));
        GRC(dgrad(c.c2, b.dA_c2, h, w, b.dA_c1, h, w, 0));

        GRC(bn_bwd(c.bc1, b.dA_c1, b.c1_raw, b.k_c1, 1, h, w, b.m_c1));
        GRC(wgrad(c.c1, b.cat, h, w, InTf{b.k_cat, 0}, b.dA_c1, h, w));
        GRC(dgrad(c.c1, b.dA_c1, h, w, b.dcat, h, w, 0));

        GRC(bn_bwd(c.bcat, b.dcat, b.cat, b.k_cat, 0, h, w, b.m_cat));
        const int C = c.cskip + c.cdeep, hu = b.hd, wu = b.wd;
        const int oyu = (2 * hu - h) / 2, oxu = (2 * wu - w) / 2;
        {
            dim3 grid(min(ceil_div(c.cskip * h * w, 256), 148 * 8), 1, N);
            SPLICE_CHECK_CUDA(launch_pdl(cat_bwd_skip_kernel, grid, dim3(256), 0, st, (const float*)b.dcat, C, h, w, c.cskip, h, w, 0, 0, b.dA_s));
            SPLICE_LAUNCH_CHECK();
            float* dU = (i == ns_ - 1) ? b.dA_d2 : sb_[i + 1].dA_c2;
            dim3 grid2(min(ceil_div(c.cdeep * hu * wu, 256), 148 * 8), 1, N);
            SPLICE_CHECK_CUDA(launch_pdl(cat_bwd_up_kernel, grid2, dim3(256), 0, st, (const float*)b.dcat, C, h, w, c.cskip, c.cdeep, hu, wu, oyu, oxu, dU));
            SPLICE_LAUNCH_CHECK();
        }
    }
    // down path, bottom to top
    for (int i = ns_ - 1; i >= 0; --i) {
        const Scale& c = sc_[i];
        ScaleBuf& b = sb_[i];
        const float* in = (i == 0) ? x_copy_ : sb_[i - 1].d2_raw;
        InTf tf_in = (i == 0) ? InTf{nullptr, 0} : InTf{sb_[i - 1].k_d2, 1};
        float* dIn = (i == 0) ? nullptr : sb_[i - 1].dA_d2;

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
    SPLICE_CHECK_CUDA(cudaEventRecord(ev_join_, ws));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join_, 0));
    return SPLICE_OK;
}

}  // namespace splice

// ---- C-ABI (include/splice_b200.h: splice_genx_*) -----------------------------------------------------------------
using namespace splice;
extern "C" {

SPLICE_API int splice_genx_create(const SpliceGenXConfig* cfg, void** ctx) {
    SPLICE_REQUIRE(cfg && ctx, "splice_genx_create: null argument");
    static_assert(SPLICE_GENX_MAX_SCALES == GENX_MAX_SCALES, "header / engine mismatch");
    GenXConfig c{};
    c.n_scales = cfg->n_scales; c.in_channels = cfg->in_channels; c.out_channels = cfg->out_channels;
    for (int i = 0; i < GENX_MAX_SCALES; ++i) {
        c.ch_down[i] = cfg->ch_down[i]; c.ch_up[i] = cfg->ch_up[i]; c.ch_skip[i] = cfg->ch_skip[i];
        c.k_down[i] = cfg->k_down[i]; c.k_up[i] = cfg->k_up[i];
    }
    c.k_skip = cfg->k_skip; c.reflect = cfg->reflect ? 1 : 0; c.sigmoid = cfg->sigmoid ? 1 : 0;
    GenXEngine* e = nullptr;
    const int rc = GenXEngine::create(c, &e);
    if (rc) return rc;
    *ctx = e;
    return SPLICE_OK;
}
SPLICE_API int splice_genx_destroy(void* ctx) {
    delete static_cast<GenXEngine*>(ctx);
    return SPLICE_OK;
}
SPLICE_API int splice_genx_counts(void* ctx, int* n_params, int* n_bn) {
    SPLICE_REQUIRE(ctx && n_params && n_bn, "splice_genx_counts: null argument");
    *n_params = static_cast<GenXEngine*>(ctx)->n_params();
    *n_bn = static_cast<GenXEngine*>(ctx)->n_bn();
    return SPLICE_OK;
}
SPLICE_API int splice_genx_bind(void* ctx, void* const* params, void* const* grads, void* const* running_mean, void* const* running_var,
                                void* const* num_batches_tracked) {
    SPLICE_REQUIRE(ctx, "splice_genx_bind: null ctx");
    return static_cast<GenXEngine*>(ctx)->bind((float* const*)params, (float* const*)grads, (float* const*)running_mean,
                                               (float* const*)running_var, (long long* const*)num_batches_tracked);
}
SPLICE_API int splice_genx_forward(void* ctx, const void* x, int N, int H, int W, void* out, int keep, int update_running, void* stream) {
    SPLICE_REQUIRE(ctx, "splice_genx_forward: null ctx");
    return static_cast<GenXEngine*>(ctx)->forward((const float*)x, N, H, W, (float*)out, keep != 0, update_running != 0, (cudaStream_t)stream);
}
SPLICE_API int splice_genx_backward(void* ctx, const void* dout, int accumulate, void* stream) {
    SPLICE_REQUIRE(ctx, "splice_genx_backward: null ctx");
    return static_cast<GenXEngine*>(ctx)->backward((const float*)dout, accumulate != 0, (cudaStream_t)stream);
}
SPLICE_API int splice_genx_set_graphs(void* ctx, int on) {
    SPLICE_REQUIRE(ctx, "splice_genx_set_graphs: null ctx");
    static_cast<GenXEngine*>(ctx)->set_graphs(on != 0);
    return SPLICE_OK;
}

}  // extern "C"
