// splice_b200 — fused softmax attention of the frozen DINO ViT, forward and dgrad, flash-style:
// the [H,t,t] probability matrix (29.6 MB fp32 per layer at t=785, saved x12 by the reference's autograd)
// never leaves the SM; only the per-row log-sum-exp is kept for the backward.
//
// Replaces  attn = softmax(q k^T * dh^-0.5); x = attn @ v   inside `self.model(input_img)`
// (models/extractor.py:83,91,99 -> DINO Attention.forward) and its autograd backward (train.py:78).
//
// Layout: qkv bf16 [S*t, 3D] exactly as the qkv Linear writes it (q | k | v, head h at columns h*64..h*64+63
// inside each third, extractor.py:139-151); o / do bf16 [S*t, D]; lse fp32 [S, H, t] in the log2 domain.
// Head dim is 64 for every DINO ViT (384/6, 768/12).
//
// Round-1 implementation: mma.sync.m16n8k16 bf16 (register accumulators), 64x64 tiles, cp.async
// double-buffered K/V (or Q/dO) tiles in XOR-swizzled shared memory. The backward is split into a dQ kernel
// (CTA per query tile, loops over key tiles) and a dK/dV kernel (CTA per key tile, loops over query tiles):
// deterministic, no atomics, at the price of recomputing S and dP once.
#include "attention.h"

#include <stdlib.h>
#include <string.h>

namespace splice {

static constexpr int HD = 64;       // head dim
#ifdef SPLICE_B200_CROSSCHECK   // the mma.sync kernels are cross-check code: built only into the test library (build.py)
static constexpr int TQ = 64;       // rows per CTA tile
static constexpr int TK = 64;       // columns (keys / queries) per inner tile
static constexpr int TILE_BYTES = 64 * 64 * 2;

// ---------------------------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros are written
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16) * B(16x8, bf16)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 x 64 bf16 tile, 128 B per row, 16-byte chunks XOR-swizzled by (row & 7): conflict-free ldmatrix.
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// global rows [row0, row0+64) x 64 columns starting at `src` (leading dim ld elements) -> swizzled tile.
// Rows >= rows_valid are zero-filled. 128 threads, 4 chunks each.
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* __restrict__ src, int ld, int row0,
                                                int rows_valid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 128;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = (row0 + r) < rows_valid;
        const bf16* g = src + (size_t)(ok ? (row0 + r) : 0) * ld + c * 8;
        cp_async16(tile_addr(sbase, r, c), g, ok);
    }
}

// A-operand fragments (16 rows x 64 cols, 4 k-steps) of this warp's 16 rows from a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t sbase, int warp, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(f[ks], tile_addr(sbase, warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
}

// acc[nb] (16 x 8 per n-block, 8 n-blocks = 64 columns) += A(16 x 64) * T^T where the tile T is stored
// [n][k] row-major (k contiguous): S = Q K^T with T = K, dP = dO V^T with T = V, S^T = K Q^T with T = Q.
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
            uint32_t b[4];
            ldsm_x4(b, tile_addr(sbase, nb2 * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
            mma16816(acc[2 * nb2], a[ks], b[0], b[1]);
            mma16816(acc[2 * nb2 + 1], a[ks], b[2], b[3]);
        }
    }
}
// acc (16 x 64) += P(16 x 64, given as A fragments over 4 k-steps of 16) * T where T is stored [k][n]
// row-major (n contiguous): O = P V with T = V, dQ = dS K with T = K, dV = P^T dO, dK = dS^T Q.
__device__ __forceinline__ void mma_a_tile(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
            uint32_t b[4];
            ldsm_x4_t(b, tile_addr(sbase, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, nb2 * 2 + (lane >> 4)));
            mma16816(acc[2 * nb2], a[kk], b[0], b[1]);
            mma16816(acc[2 * nb2 + 1], a[kk], b[2], b[3]);
        }
    }
}
// fp32 accumulator tile (16 x 64) -> bf16 A fragments for a following MMA (k = the 64 columns)
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4][4], const float (&acc)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        a[kk][0] = pack_bf16x2(acc[2 * kk][0], acc[2 * kk][1]);
        a[kk][1] = pack_bf16x2(acc[2 * kk][2], acc[2 * kk][3]);
        a[kk][2] = pack_bf16x2(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
        a[kk][3] = pack_bf16x2(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
    }
}
__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}
// store this warp's 16 x 64 accumulator rows (scaled) as bf16 into dst[row, col0 + 0..63]
__device__ __forceinline__ void store_acc_bf16(bf16* __restrict__ dst, int ld, int row_lo, int rows_valid, const float (&acc)[8][4],
                                               float scale_lo, float scale_hi, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int r0 = row_lo + g, r1 = row_lo + g + 8;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int c = nb * 8 + 2 * q;
        if (r0 < rows_valid)
            *reinterpret_cast<uint32_t*>(dst + (size_t)r0 * ld + c) = pack_bf16x2(acc[nb][0] * scale_lo, acc[nb][1] * scale_lo);
        if (r1 < rows_valid)
            *reinterpret_cast<uint32_t*>(dst + (size_t)r1 * ld + c) = pack_bf16x2(acc[nb][2] * scale_hi, acc[nb][3] * scale_hi);
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                       float* __restrict__ lse, int t, int D, float scale_log2) {
    __shared__ __align__(128) uint8_t smem[5 * TILE_BYTES];  // Q | K0 | K1 | V0 | V1
    pdl_sync();
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK[2] = {sQ + TILE_BYTES, sQ + 2 * TILE_BYTES};
    const uint32_t sV[2] = {sQ + 3 * TILE_BYTES, sQ + 4 * TILE_BYTES};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, qd = lane & 3;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int q0 = blockIdx.x * TQ;
    const int ld = 3 * D;
    const bf16* base = qkv + (size_t)s * t * ld;
    const bf16* gQ = base + h * HD;
    const bf16* gK = base + D + h * HD;
    const bf16* gV = base + 2 * D + h * HD;
    const int nkv = (t + TK - 1) / TK;

    load_tile_async(sQ, gQ, ld, q0, t);
    load_tile_async(sK[0], gK, ld, 0, t);
    load_tile_async(sV[0], gV, ld, 0, t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t qf[4][4];
    load_a_frags(qf, sQ, warp, lane);

    float oacc[8][4];
    zero_acc(oacc);
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
            load_tile_async(sK[buf ^ 1], gK, ld, (j + 1) * TK, t);
            load_tile_async(sV[buf ^ 1], gV, ld, (j + 1) * TK, t);
            cp_async_commit();
        }
        float sacc[8][4];
        zero_acc(sacc);
        mma_a_tileT(sacc, qf, sK[buf], lane);
        // scale into the log2 domain and mask keys beyond the sequence
        const int kvalid = t - j * TK;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int c = nb * 8 + 2 * qd;
            sacc[nb][0] = (c < kvalid) ? sacc[nb][0] * scale_log2 : -INFINITY;
            sacc[nb][1] = (c + 1 < kvalid) ? sacc[nb][1] * scale_log2 : -INFINITY;
            sacc[nb][2] = (c < kvalid) ? sacc[nb][2] * scale_log2 : -INFINITY;
            sacc[nb][3] = (c + 1 < kvalid) ? sacc[nb][3] * scale_log2 : -INFINITY;
        }
        float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            mx_lo = fmaxf(mx_lo, fmaxf(sacc[nb][0], sacc[nb][1]));
            mx_hi = fmaxf(mx_hi, fmaxf(sacc[nb][2], sacc[nb][3]));
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float corr_lo = exp2f(m_lo - mx_lo), corr_hi = exp2f(m_hi - mx_hi);  // first tile: exp2(-inf) = 0
        m_lo = mx_lo;
        m_hi = mx_hi;
        float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            sacc[nb][0] = exp2f(sacc[nb][0] - m_lo);
            sacc[nb][1] = exp2f(sacc[nb][1] - m_lo);
            sacc[nb][2] = exp2f(sacc[nb][2] - m_hi);
            sacc[nb][3] = exp2f(sacc[nb][3] - m_hi);
            sum_lo += sacc[nb][0] + sacc[nb][1];
            sum_hi += sacc[nb][2] + sacc[nb][3];
        }
        l_lo = l_lo * corr_lo + sum_lo;  // per-thread partial sums; reduced across the quad at the end
        l_hi = l_hi * corr_hi + sum_hi;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            oacc[nb][0] *= corr_lo; oacc[nb][1] *= corr_lo;
            oacc[nb][2] *= corr_hi; oacc[nb][3] *= corr_hi;
        }
        uint32_t pf[4][4];
        acc_to_a(pf, sacc);
        mma_a_tile(oacc, pf, sV[buf], lane);
        if (j + 1 < nkv) cp_async_wait<0>();
        __syncthreads();
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const int row_lo = q0 + warp * 16;
    store_acc_bf16(o + (size_t)s * t * D + h * HD, D, row_lo, t, oacc, 1.f / l_lo, 1.f / l_hi, lane);
    if (qd == 0) {
        float* L = lse + ((size_t)s * H + h) * t;
        if (row_lo + g < t) L[row_lo + g] = m_lo + log2f(l_lo);
        if (row_lo + g + 8 < t) L[row_lo + g + 8] = m_hi + log2f(l_hi);
    }
}

// ---------------------------------------------------------------------------------------------
// backward, part 1: dQ (and delta = rowsum(dO * O), which part 2 consumes)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                          const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                          float* __restrict__ delta, bf16* __restrict__ dqkv, int t, int D,
                                                          float scale, float scale_log2) {
    __shared__ __align__(128) uint8_t smem[5 * TILE_BYTES];  // Q(then dO) | K0 | K1 | V0 | V1
    pdl_sync();
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK[2] = {sQ + TILE_BYTES, sQ + 2 * TILE_BYTES};
    const uint32_t sV[2] = {sQ + 3 * TILE_BYTES, sQ + 4 * TILE_BYTES};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, qd = lane & 3;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int q0 = blockIdx.x * TQ;
    const int ld = 3 * D;
    const bf16* base = qkv + (size_t)s * t * ld;
    const bf16* gQ = base + h * HD;
    const bf16* gK = base + D + h * HD;
    const bf16* gV = base + 2 * D + h * HD;
    const bf16* gO = o + (size_t)s * t * D + h * HD;
    const bf16* gdO = dout + (size_t)s * t * D + h * HD;
    const int nkv = (t + TK - 1) / TK;

    // stage Q, O (in K1's slot) and dO (in V1's slot) to build the register-resident A operands
    load_tile_async(sQ, gQ, ld, q0, t);
    load_tile_async(sK[1], gO, D, q0, t);
    load_tile_async(sV[1], gdO, D, q0, t);
    load_tile_async(sK[0], gK, ld, 0, t);
    load_tile_async(sV[0], gV, ld, 0, t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(qf, sQ, warp, lane);
    load_a_frags(dof, sV[1], warp, lane);
    float dl_lo = 0.f, dl_hi = 0.f;
    {
        uint32_t of[4][4];
        load_a_frags(of, sK[1], warp, lane);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            float2 a, b;
            a = unpack_bf16x2(of[ks][0]); b = unpack_bf16x2(dof[ks][0]); dl_lo += a.x * b.x + a.y * b.y;
            a = unpack_bf16x2(of[ks][2]); b = unpack_bf16x2(dof[ks][2]); dl_lo += a.x * b.x + a.y * b.y;
            a = unpack_bf16x2(of[ks][1]); b = unpack_bf16x2(dof[ks][1]); dl_hi += a.x * b.x + a.y * b.y;
            a = unpack_bf16x2(of[ks][3]); b = unpack_bf16x2(dof[ks][3]); dl_hi += a.x * b.x + a.y * b.y;
        }
        dl_lo += __shfl_xor_sync(0xffffffffu, dl_lo, 1);
        dl_lo += __shfl_xor_sync(0xffffffffu, dl_lo, 2);
        dl_hi += __shfl_xor_sync(0xffffffffu, dl_hi, 1);
        dl_hi += __shfl_xor_sync(0xffffffffu, dl_hi, 2);
    }
    const int row_lo = q0 + warp * 16;
    const float* L = lse + ((size_t)s * H + h) * t;
    const float lse_lo = (row_lo + g < t) ? L[row_lo + g] : 0.f;
    const float lse_hi = (row_lo + g + 8 < t) ? L[row_lo + g + 8] : 0.f;
    if (qd == 0) {
        float* Dl = delta + ((size_t)s * H + h) * t;
        if (row_lo + g < t) Dl[row_lo + g] = dl_lo;
        if (row_lo + g + 8 < t) Dl[row_lo + g + 8] = dl_hi;
    }
    __syncthreads();  // everyone has read the O / dO staging slots before tile 1 is prefetched into them

    float dq[8][4];
    zero_acc(dq);
    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j + 1 < nkv) {
            load_tile_async(sK[buf ^ 1], gK, ld, (j + 1) * TK, t);
            load_tile_async(sV[buf ^ 1], gV, ld, (j + 1) * TK, t);
            cp_async_commit();
        }
        float sacc[8][4], dp[8][4];
        zero_acc(sacc);
        zero_acc(dp);
        mma_a_tileT(sacc, qf, sK[buf], lane);   // S  = Q K^T
        mma_a_tileT(dp, dof, sV[buf], lane);    // dP = dO V^T
        const int kvalid = t - j * TK;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int c = nb * 8 + 2 * qd;
            const float p0 = (c < kvalid) ? exp2f(sacc[nb][0] * scale_log2 - lse_lo) : 0.f;
            const float p1 = (c + 1 < kvalid) ? exp2f(sacc[nb][1] * scale_log2 - lse_lo) : 0.f;
            const float p2 = (c < kvalid) ? exp2f(sacc[nb][2] * scale_log2 - lse_hi) : 0.f;
            const float p3 = (c + 1 < kvalid) ? exp2f(sacc[nb][3] * scale_log2 - lse_hi) : 0.f;
            sacc[nb][0] = p0 * (dp[nb][0] - dl_lo);
            sacc[nb][1] = p1 * (dp[nb][1] - dl_lo);
            sacc[nb][2] = p2 * (dp[nb][2] - dl_hi);
            sacc[nb][3] = p3 * (dp[nb][3] - dl_hi);
        }
        uint32_t dsf[4][4];
        acc_to_a(dsf, sacc);
        mma_a_tile(dq, dsf, sK[buf], lane);     // dQ += dS K
        if (j + 1 < nkv) cp_async_wait<0>();
        __syncthreads();
    }
    store_acc_bf16(dqkv + (size_t)s * t * ld + h * HD, ld, row_lo, t, dq, scale, scale, lane);
}

// ---------------------------------------------------------------------------------------------
// backward, part 2: dK and dV. CTA per 64-key tile, each warp owns 16 keys; loops over query tiles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ delta,
                                                           bf16* __restrict__ dqkv, int t, int D, float scale,
                                                           float scale_log2) {
    __shared__ __align__(128) uint8_t smem[4 * TILE_BYTES];  // Q0 | Q1 | dO0 | dO1  (K, V staged in Q1 / dO1 first)
    __shared__ float s_lse[2][TK];
    __shared__ float s_dl[2][TK];
    pdl_sync();
    const uint32_t s0 = smem_u32(smem);
    const uint32_t sQ[2] = {s0, s0 + TILE_BYTES};
    const uint32_t sdO[2] = {s0 + 2 * TILE_BYTES, s0 + 3 * TILE_BYTES};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qd = lane & 3;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int k0 = blockIdx.x * TK;
    const int ld = 3 * D;
    const bf16* base = qkv + (size_t)s * t * ld;
    const bf16* gQ = base + h * HD;
    const bf16* gK = base + D + h * HD;
    const bf16* gV = base + 2 * D + h * HD;
    const bf16* gdO = dout + (size_t)s * t * D + h * HD;
    const float* L = lse + ((size_t)s * H + h) * t;
    const float* Dl = delta + ((size_t)s * H + h) * t;
    const int nq = (t + TQ - 1) / TQ;

    auto load_stats = [&](int buf, int qbase) {
        if (threadIdx.x < TK) {
            const int r = qbase + threadIdx.x;
            // +inf lse => P = exp2(-inf) = 0 for query columns beyond the sequence
            s_lse[buf][threadIdx.x] = (r < t) ? L[r] : INFINITY;
            s_dl[buf][threadIdx.x] = (r < t) ? Dl[r] : 0.f;
        }
    };

    load_tile_async(sQ[1], gK, ld, k0, t);
    load_tile_async(sdO[1], gV, ld, k0, t);
    load_tile_async(sQ[0], gQ, ld, 0, t);
    load_tile_async(sdO[0], gdO, D, 0, t);
    cp_async_commit();
    load_stats(0, 0);
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(kf, sQ[1], warp, lane);
    load_a_frags(vf, sdO[1], warp, lane);
    __syncthreads();

    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);
    for (int i = 0; i < nq; ++i) {
        const int buf = i & 1;
        if (i + 1 < nq) {
            load_tile_async(sQ[buf ^ 1], gQ, ld, (i + 1) * TQ, t);
            load_tile_async(sdO[buf ^ 1], gdO, D, (i + 1) * TQ, t);
            cp_async_commit();
            load_stats(buf ^ 1, (i + 1) * TQ);
        }
        float st[8][4], dpt[8][4];
        zero_acc(st);
        zero_acc(dpt);
        mma_a_tileT(st, kf, sQ[buf], lane);     // S^T  = K Q^T   (rows: keys, cols: queries)
        mma_a_tileT(dpt, vf, sdO[buf], lane);   // dP^T = V dO^T
        uint32_t pf[4][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int c = nb * 8 + 2 * qd;
            const float l0 = s_lse[buf][c], l1 = s_lse[buf][c + 1];
            const float d0 = s_dl[buf][c], d1 = s_dl[buf][c + 1];
            const float p0 = exp2f(st[nb][0] * scale_log2 - l0);
            const float p1 = exp2f(st[nb][1] * scale_log2 - l1);
            const float p2 = exp2f(st[nb][2] * scale_log2 - l0);
            const float p3 = exp2f(st[nb][3] * scale_log2 - l1);
            st[nb][0] = p0; st[nb][1] = p1; st[nb][2] = p2; st[nb][3] = p3;
            dpt[nb][0] = p0 * (dpt[nb][0] - d0);
            dpt[nb][1] = p1 * (dpt[nb][1] - d1);
            dpt[nb][2] = p2 * (dpt[nb][2] - d0);
            dpt[nb][3] = p3 * (dpt[nb][3] - d1);
        }
        acc_to_a(pf, st);
        mma_a_tile(dv, pf, sdO[buf], lane);     // dV += P^T dO
        acc_to_a(pf, dpt);
        mma_a_tile(dk, pf, sQ[buf], lane);      // dK += dS^T Q
        if (i + 1 < nq) cp_async_wait<0>();
        __syncthreads();
    }
    const int row_lo = k0 + warp * 16;
    store_acc_bf16(dqkv + (size_t)s * t * ld + D + h * HD, ld, row_lo, t, dk, scale, scale, lane);
    store_acc_bf16(dqkv + (size_t)s * t * ld + 2 * D + h * HD, ld, row_lo, t, dv, 1.f, 1.f, lane);
}

#endif  // SPLICE_B200_CROSSCHECK
// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int check_dims(int S, int t, int D, int H) {
    SPLICE_REQUIRE(S > 0 && t > 0, "attention: empty problem S=%d t=%d", S, t);
    SPLICE_REQUIRE(H > 0 && D == H * HD, "attention: head dim must be 64 (D=%d, H=%d)", D, H);
    return SPLICE_OK;
}

// 0 = tcgen05 kernels (attention_tc.cu, product path); SPLICE_B200_ATTN=legacy | tcfwd | tcbwd selects the mma.sync
// cross-check kernels for both / the backward / the forward direction (A/B comparison and parity tests)
static int attn_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* v = getenv("SPLICE_B200_ATTN");
        mode = 0;
        if (v && !strcmp(v, "legacy")) mode = 3;
        else if (v && !strcmp(v, "tcfwd")) mode = 2;   // legacy backward
        else if (v && !strcmp(v, "tcbwd")) mode = 1;   // legacy forward
    }
    return mode;
}

int attention_fwd(const bf16* qkv, bf16* o, float* lse, int S, int t, int D, int H, cudaStream_t stream) {
    int rc = check_dims(S, t, D, H);
    if (rc) return rc;
    if (!(attn_mode() & 1)) return attention_fwd_tc(qkv, o, lse, S, t, D, H, stream);
#ifndef SPLICE_B200_CROSSCHECK
    set_error("attention: the mma.sync cross-check kernels are not in this build (SPLICE_B200_CROSSCHECK=1 python -m splice_b200.build)");
    return SPLICE_ERR_UNSUPPORTED;
#else
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e)
    dim3 grid(ceil_div(t, TQ), H, S);
    SPLICE_CHECK_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(128), 0, stream, qkv, o, lse, t, D, scale_log2));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
#endif
}

int attention_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int S, int t,
                  int D, int H, cudaStream_t stream) {
    int rc = check_dims(S, t, D, H);
    if (rc) return rc;
    if (!(attn_mode() & 2)) return attention_bwd_tc(qkv, o, dout, lse, delta, dqkv, S, t, D, H, stream);
#ifndef SPLICE_B200_CROSSCHECK
    set_error("attention: the mma.sync cross-check kernels are not in this build (SPLICE_B200_CROSSCHECK=1 python -m splice_b200.build)");
    return SPLICE_ERR_UNSUPPORTED;
#else
    const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
    dim3 grid(ceil_div(t, TQ), H, S);
    SPLICE_CHECK_CUDA(launch_pdl(attn_bwd_dq_kernel, grid, dim3(128), 0, stream, qkv, o, dout, lse, delta, dqkv, t, D, scale, scale_log2));
    SPLICE_LAUNCH_CHECK();
    SPLICE_CHECK_CUDA(launch_pdl(attn_bwd_dkv_kernel, grid, dim3(128), 0, stream, qkv, dout, lse, delta, dqkv, t, D, scale, scale_log2));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
#endif
}

}  // namespace splice
