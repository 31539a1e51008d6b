// splice_b200 — GEMM interface shared by the tcgen05 kernel, the SIMT cross-check kernel and capi.cu.
//   C[M,N] = epilogue( A[M,K] · B[N,K]^T )      A, B bf16 row-major (K contiguous), fp32 accumulation.
// This single shape ("TN", both operands K-major) covers every dense contraction of the frozen ViT:
// forward uses the weight W[out,in] as B; the dgrad-only backward uses the pre-transposed copy W^T.
#pragma once
#include "common.cuh"

namespace splice {

enum GemmAct : int {
    GEMM_ACT_NONE = 0,
    GEMM_ACT_GELU = 1,        // v = gelu(v); pre-activation optionally saved to aux16 (bf16)
    GEMM_ACT_GELU_GRAD = 2,   // v = v * gelu'(aux16[row,col])
};

struct GemmEpilogue {
    float* c32 = nullptr;            // optional fp32 output [*, ldc32]
    int ldc32 = 0;
    bf16* c16 = nullptr;             // optional bf16 output [*, ldc16]
    int ldc16 = 0;
    const float* bias = nullptr;     // optional [N]
    const float* residual = nullptr; // optional fp32 [*, ldr]; added before the store (may alias c32)
    int ldr = 0;
    int act = GEMM_ACT_NONE;
    bf16* aux16 = nullptr;           // GELU: pre-activation out; GELU_GRAD: pre-activation in
    int ldaux = 0;
    // patch-embed token remap: GEMM row r (patch r of a batch of sequences, rows_per_seq patches each)
    // is written to token row r + r / rows_per_seq + 1 and gets pos[(r % rows_per_seq) + 1, :] added.
    int rows_per_seq = 0;            // 0 = no remap
    const float* pos = nullptr;
    int ldpos = 0;
    // fp32 export of the column slice [slice_c0, slice_c1) (layer-11 keys for the loss kernels)
    float* slice32 = nullptr;
    int slice_c0 = 0, slice_c1 = 0, ldslice = 0;
    // B is a constant (a weight matrix no kernel of the stream writes): its first tiles may be fetched before the
    // programmatic-dependent-launch wait, i.e. while the previous kernel is still draining
    int b_const = 0;
    // ---- LayerNorm folded into the neighbouring GEMMs (no LayerNorm launch on the forward chain) ----
    // producer side (the GEMM that writes a residual-stream row, N = D): per 32-column chunk of the FINAL value the pair
    // (sum, centred sum of squares) goes to stat_part[orow * stat_nparts + col / 32]
    float2* stat_part = nullptr;
    int stat_nparts = 0;
    // consumer side (A = the bf16 copy of the RAW residual stream, B = W * gamma): the row's (mean, rstd) are merged
    // from the producer's partials and applied to the accumulator, v = rstd * (acc - mean * colsum[col]) (+ bias = the
    // folded bias b + W beta); the column-0 chunk writes (mean, rstd) to ln_stat_out[row] for the LayerNorm backward
    const float2* ln_part = nullptr;
    int ln_nparts = 0;
    const float* ln_colsum = nullptr;   // [N]: sum over k of bf16(W[n,k] * gamma[k])
    float2* ln_stat_out = nullptr;
    float ln_eps = 0.f;
};

struct GemmRowCtx { float mean, rstd; };

enum GemmImpl : int { GEMM_IMPL_TCGEN05 = 0, GEMM_IMPL_SIMT = 1, GEMM_IMPL_TCGEN05_TILE = 2 };

// Launches on `stream`. bn_hint: 0 = pick automatically, else one of 64/128/256.
int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, const GemmEpilogue& ep,
                 int impl, int bn_hint, cudaStream_t stream);

// Cached TMA descriptor of a row-major bf16 matrix [rows, cols] (leading dimension ld elements): box = box_rows rows x 64
// columns (128 bytes), 128B swizzle. Shared with the attention kernels.
int make_tmap_bf16(CUtensorMap* tm, const bf16* ptr, int rows, int cols, int ld, int box_rows);

#ifdef __CUDACC__
// Per-row context of the epilogue (computed once per tile and thread, before the accumulator is waited for): the
// LayerNorm statistics of `row` merged from the producer's per-chunk partials (Chan's formula; 32 values per part).
__device__ __forceinline__ GemmRowCtx gemm_epilogue_row(const GemmEpilogue& ep, int row) {
    GemmRowCtx rc{0.f, 1.f};
    if (ep.ln_part) {
        const float2* p = ep.ln_part + (size_t)row * ep.ln_nparts;
        float sum = 0.f;
        for (int i = 0; i < ep.ln_nparts; ++i) sum += p[i].x;
        const float inv_n = 1.f / (32.f * ep.ln_nparts);
        const float mean = sum * inv_n;
        float m2 = 0.f;
        for (int i = 0; i < ep.ln_nparts; ++i) {
            const float2 q = p[i];
            const float d = q.x * (1.f / 32.f) - mean;
            m2 += q.y + 32.f * d * d;
        }
        rc.mean = mean;
        rc.rstd = rsqrtf(m2 * inv_n + ep.ln_eps);
    }
    return rc;
}

// Shared epilogue: 32 consecutive accumulator columns [col, col+32) of GEMM row `row`.
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmEpilogue& ep, int row, int col, float (&v)[32],
                                                    const GemmRowCtx rc = GemmRowCtx{0.f, 1.f}) {
    if (ep.ln_part) {
        const float4* c4 = reinterpret_cast<const float4*>(ep.ln_colsum + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 c = __ldg(c4 + j);
            v[4 * j + 0] = rc.rstd * (v[4 * j + 0] - rc.mean * c.x); v[4 * j + 1] = rc.rstd * (v[4 * j + 1] - rc.mean * c.y);
            v[4 * j + 2] = rc.rstd * (v[4 * j + 2] - rc.mean * c.z); v[4 * j + 3] = rc.rstd * (v[4 * j + 3] - rc.mean * c.w);
        }
        if (ep.ln_stat_out && col == 0) ep.ln_stat_out[row] = make_float2(rc.mean, rc.rstd);
    }
    if (ep.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    if (ep.act == GEMM_ACT_GELU) {
        if (ep.aux16) {
            uint4* a4 = reinterpret_cast<uint4*>(ep.aux16 + (size_t)row * ep.ldaux + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                a4[j] = u;
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (ep.act == GEMM_ACT_GELU_GRAD) {
        const uint4* a4 = reinterpret_cast<const uint4*>(ep.aux16 + (size_t)row * ep.ldaux + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 u = a4[j];
            float2 p;
            p = unpack_bf16x2(u.x); v[8 * j + 0] *= gelu_erf_grad(p.x); v[8 * j + 1] *= gelu_erf_grad(p.y);
            p = unpack_bf16x2(u.y); v[8 * j + 2] *= gelu_erf_grad(p.x); v[8 * j + 3] *= gelu_erf_grad(p.y);
            p = unpack_bf16x2(u.z); v[8 * j + 4] *= gelu_erf_grad(p.x); v[8 * j + 5] *= gelu_erf_grad(p.y);
            p = unpack_bf16x2(u.w); v[8 * j + 6] *= gelu_erf_grad(p.x); v[8 * j + 7] *= gelu_erf_grad(p.y);
        }
    }
    int orow = row;
    if (ep.rows_per_seq > 0) {
        const int s = row / ep.rows_per_seq;
        const int r = row - s * ep.rows_per_seq;
        orow = row + s + 1;
        if (ep.pos) {
            const float4* p4 = reinterpret_cast<const float4*>(ep.pos + (size_t)(r + 1) * ep.ldpos + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 p = __ldg(p4 + j);
                v[4 * j + 0] += p.x; v[4 * j + 1] += p.y; v[4 * j + 2] += p.z; v[4 * j + 3] += p.w;
            }
        }
    }
    if (ep.residual) {
        const float4* r4 = reinterpret_cast<const float4*>(ep.residual + (size_t)orow * ep.ldr + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 r = r4[j];
            v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
    }
    if (ep.stat_part) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += v[j];
        const float mc = sum * (1.f / 32.f);
        float m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) m2 += (v[j] - mc) * (v[j] - mc);
        ep.stat_part[(size_t)orow * ep.stat_nparts + (col >> 5)] = make_float2(sum, m2);
    }
    if (ep.c32) {
        float4* c4 = reinterpret_cast<float4*>(ep.c32 + (size_t)orow * ep.ldc32 + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) c4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (ep.c16) {
        uint4* c4 = reinterpret_cast<uint4*>(ep.c16 + (size_t)orow * ep.ldc16 + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            c4[j] = u;
        }
    }
    if (ep.slice32 && col >= ep.slice_c0 && col < ep.slice_c1) {
        float4* s4 = reinterpret_cast<float4*>(ep.slice32 + (size_t)orow * ep.ldslice + (col - ep.slice_c0));
#pragma unroll
        for (int j = 0; j < 8; ++j) s4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}
#endif

}  // namespace splice
