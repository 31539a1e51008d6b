// splice_b200 — LayerNorm forward/backward and small glue kernels of the ViT path.
// Replaces the nn.LayerNorm(eps=1e-6) calls (norm1 / norm2 of each DINO block) inside
// `self.model(input_img)` (models/extractor.py:83,91,99) and their autograd backward (train.py:78).
// HBM-bound row kernels: one warp per token row, float4 loads, warp-shuffle reductions.
#include "elementwise.h"

namespace splice {

static constexpr int LN_MAX_V4 = 8;  // D <= 1024

template <bool kStats>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float* __restrict__ stats, int M, int D, float eps) {
    pdl_sync();
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nv = D >> 7;  // float4 per lane
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[LN_MAX_V4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
        if (i < nv) {
            v[i] = xr[lane + i * 32];
            sum += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    const float mean = warp_sum(sum) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    const float rstd = rsqrtf(warp_sum(sq) / D + eps);
    if (kStats && lane == 0) {
        stats[2 * row] = mean;
        stats[2 * row + 1] = rstd;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
        if (i < nv) {
            const float4 g = __ldg(g4 + lane + i * 32), b = __ldg(b4 + lane + i * 32);
            uint2 o;
            o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            yr[lane + i * 32] = o;
        }
}

__global__ void __launch_bounds__(128) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ stats, const float* __restrict__ gamma,
                                                     const float* g_in, float* g_out, bf16* __restrict__ g16, int M, int D) {
    pdl_sync();
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nv = D >> 7;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * D);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4 dg[LN_MAX_V4], xh[LN_MAX_V4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
        if (i < nv) {
            const float4 d = dyr[lane + i * 32], xv = xr[lane + i * 32];
            const float4 g = gamma ? __ldg(g4 + lane + i * 32) : make_float4(1.f, 1.f, 1.f, 1.f);   // null: gamma folded into the weights
            dg[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            s1 += dg[i].x + dg[i].y + dg[i].z + dg[i].w;
            s2 += dg[i].x * xh[i].x + dg[i].y * xh[i].y + dg[i].z * xh[i].z + dg[i].w * xh[i].w;
        }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    const float4* gi = reinterpret_cast<const float4*>(g_in ? g_in + (size_t)row * D : nullptr);
    float4* go = reinterpret_cast<float4*>(g_out + (size_t)row * D);
    uint2* g16r = g16 ? reinterpret_cast<uint2*>(g16 + (size_t)row * D) : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
        if (i < nv) {
            float4 r;
            r.x = rstd * (dg[i].x - s1 - xh[i].x * s2);
            r.y = rstd * (dg[i].y - s1 - xh[i].y * s2);
            r.z = rstd * (dg[i].z - s1 - xh[i].z * s2);
            r.w = rstd * (dg[i].w - s1 - xh[i].w * s2);
            if (g_in) {
                const float4 a = gi[lane + i * 32];
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
            }
            go[lane + i * 32] = r;
            if (g16r) {
                uint2 o;
                o.x = pack_bf16x2(r.x, r.y);
                o.y = pack_bf16x2(r.z, r.w);
                g16r[lane + i * 32] = o;
            }
        }
}

__global__ void cast_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        dst[i] = o;
    }
}

// x16 / stat_part (optional): the bf16 copy of the row and its per-32-column (sum, centred sum of squares) partials, what
// the GEMM epilogues write for every other residual-stream row when LayerNorm is folded into the GEMMs (gemm.h)
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int t,
                                int D, bf16* __restrict__ x16, float2* __restrict__ stat_part) {
    pdl_sync();
    const int s = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int c0 = warp * 32; c0 < D; c0 += nwarp * 32) {     // D is a multiple of 128
        const int c = c0 + lane;
        const float v = cls[c] + pos[c];
        x[(size_t)s * t * D + c] = v;
        if (x16) x16[(size_t)s * t * D + c] = __float2bfloat16(v);
        if (stat_part) {
            const float sum = warp_sum(v);
            const float d = v - sum * (1.f / 32.f);
            const float m2 = warp_sum(d * d);
            if (lane == 0) stat_part[(size_t)s * t * (D / 32) + (c0 >> 5)] = make_float2(sum, m2);
        }
    }
}

__global__ void add_cols_kernel(bf16* __restrict__ dst, int ldd, int col0, const float* __restrict__ src, int lds, int rows,
                                int cols) {
    pdl_sync();
    const int c2 = cols >> 1;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * c2; i += (size_t)gridDim.x * blockDim.x) {
        const int r = i / c2, c = (i % c2) * 2;
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)r * ldd + col0 + c);
        const float2 a = unpack_bf16x2(*d);
        const float2 b = *reinterpret_cast<const float2*>(src + (size_t)r * lds + c);
        *d = pack_bf16x2(a.x + b.x, a.y + b.y);
    }
}

int layernorm_fwd(const float* x, const float* gamma, const float* beta, bf16* y16, float* stats, int M, int D, float eps,
                  cudaStream_t stream) {
    SPLICE_REQUIRE(M > 0 && D % 128 == 0 && D <= 128 * LN_MAX_V4, "layernorm: D=%d must be a multiple of 128, <= %d", D,
                   128 * LN_MAX_V4);
    if (stats)
        SPLICE_CHECK_CUDA(launch_pdl(ln_fwd_kernel<true>, dim3(ceil_div(M, 4)), dim3(128), 0, stream, x, gamma, beta, y16, stats, M, D, eps));
    else
        SPLICE_CHECK_CUDA(launch_pdl(ln_fwd_kernel<false>, dim3(ceil_div(M, 4)), dim3(128), 0, stream, x, gamma, beta, y16, (float*)nullptr, M, D, eps));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* g_in, float* g_out,
                  bf16* g16, int M, int D, cudaStream_t stream) {
    SPLICE_REQUIRE(M > 0 && D % 128 == 0 && D <= 128 * LN_MAX_V4, "layernorm_bwd: D=%d must be a multiple of 128, <= %d", D,
                   128 * LN_MAX_V4);
    SPLICE_CHECK_CUDA(launch_pdl(ln_bwd_kernel, dim3(ceil_div(M, 4)), dim3(128), 0, stream, dy, x, stats, gamma, g_in, g_out, g16, M, D));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int cast_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t stream) {
    SPLICE_REQUIRE(n % 4 == 0, "cast: n=%zu must be a multiple of 4", n);
    if (n == 0) return SPLICE_OK;
    const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 : 148 * 8);
    cast_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int write_cls_rows(float* x, const float* cls, const float* pos, int S, int t, int D, cudaStream_t stream, bf16* x16,
                   float2* stat_part) {
    SPLICE_CHECK_CUDA(launch_pdl(cls_rows_kernel, dim3(S), dim3(256), 0, stream, x, cls, pos, t, D, x16, stat_part));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int add_f32_into_bf16_cols(bf16* dst, int ldd, int col0, const float* src, int lds, int rows, int cols, cudaStream_t stream) {
    SPLICE_REQUIRE(cols % 2 == 0 && col0 % 2 == 0 && ldd % 2 == 0 && lds % 2 == 0, "add_cols: even sizes required");
    const size_t n = (size_t)rows * cols / 2;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    SPLICE_CHECK_CUDA(launch_pdl(add_cols_kernel, dim3(blocks), dim3(256), 0, stream, dst, ldd, col0, src, lds, rows, cols));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

}  // namespace splice
