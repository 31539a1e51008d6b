// splice_b200 — softmax attention of the frozen DINO ViT on the 5th-generation tensor cores (tcgen05.mma, TMEM
// accumulators, TMA-staged 128B-swizzled tiles), forward and dgrad. Product path of attention.h; the mma.sync kernels
// in attention.cu remain as the cross-check implementation (SPLICE_B200_ATTN=legacy).
//
// Replaces  attn = softmax(q k^T * dh^-0.5); x = attn @ v  inside `self.model(input_img)` (models/extractor.py:83,91,99
// -> DINO Attention.forward) and its autograd backward (train.py:78). Layout as in attention.cu: qkv bf16 [S*t, 3D]
// (q | k | v, head h at columns h*64.. of each third), o / do bf16 [S*t, D], lse fp32 [S,H,t] in the log2 domain.
//
// All three kernels share one shape: 6 warps per CTA, two CTAs co-resident per SM (<= 100 KB shared memory, 256 TMEM
// columns each), so that one CTA's tensor-core phase overlaps the other's exponentials:
//   warp 0     TMA producer (one lane): Q / K / V / dO tiles of 64 head-dim columns = one 128-byte swizzle row
//   warp 1     MMA issuer (one lane) + TMEM allocation
//   warps 2-5  one thread per accumulator row (= TMEM lane): tcgen05.ld -> softmax / dS arithmetic in registers ->
//              bf16 operand tile written back to shared memory in the same 128B-swizzled K-major layout TMA produces,
//              so the next tcgen05.mma consumes it as its A operand
// Contractions and their operand layouts (dh = 64 everywhere):
//   forward    S = Q K^T      A = Q  [q][dh]  K-major      B = K [key][dh] K-major
//              O_j = P V      A = P  [q][key] K-major      B = V [key][dh] MN-major (dh contiguous: the tile as loaded)
//   dK/dV      S^T = K Q^T, dP^T = V dO^T                 (A = K / V tile, B = Q / dO tile, all K-major)
//              dV += P^T dO,  dK += dS^T Q                 A = P^T / dS^T [key][q] K-major,  B = dO / Q [q][dh] MN-major
//   dQ         S = Q K^T, dP = dO V^T, dQ += dS K          B = K [key][dh] MN-major for the last one
// The forward does not rescale an accumulator in TMEM: each key tile's P V product lands in its own TMEM buffer and
// the row threads fold it into fp32 registers with the running-max correction (o = o * 2^(m_old - m_new) + P V).
// The backward is three launches: delta = rowsum(dO * O) (row kernel), then the dQ kernel and the dK/dV kernel side by
// side on two streams (they only share read-only inputs; one CTA of each fits an SM). No atomics: deterministic.
#include "attention.h"
#include "gemm.h"

namespace splice {

static constexpr int HD = 64;
static constexpr uint32_t ROW_B = 128;                 // bytes per tile row (64 bf16)
static constexpr uint32_t T128 = 128 * ROW_B;          // 128-row tile, 16 KB
static constexpr uint32_t T64 = 64 * ROW_B;            // 64-row tile, 8 KB
static constexpr uint32_t IDESC_B_MN = 1u << 16;       // instruction descriptor: B operand is MN-major ("transposed")

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// named barrier among the 128 row threads (warps 2..5)
__device__ __forceinline__ void rows_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// thread `row` of a 128-row K-major operand tile writes 16-byte chunk `chunk` (8 bf16, chunk < 8) of its row
__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int row, int chunk, uint4 v) {
    *reinterpret_cast<uint4*>(tile + row * ROW_B + ((chunk ^ (row & 7)) << 4)) = v;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    return u;
}
// A (K-major) or B (K-major) operand: k-step ks (16 elements = 32 bytes) inside the 128-byte swizzle row
__device__ __forceinline__ uint64_t desc_k(const uint8_t* tile, int ks) { return make_sw128_kmajor_desc(smem_u32(tile)) + 2u * ks; }
// B (MN-major) operand from a [k][n = 64] tile: k-step ks = 16 rows of 128 bytes = 2048 bytes (two 8-row swizzle atoms,
// 1024 bytes apart = the descriptor's stride byte offset)
__device__ __forceinline__ uint64_t desc_mn(const uint8_t* tile, int ks) { return make_sw128_kmajor_desc(smem_u32(tile)) + 128u * ks; }

static inline uint32_t align_slack(uint32_t bytes) { return bytes + 1024; }

// =================================================================================================================
// forward: CTA = (128-query tile, head, sequence); key tiles of 128
// =================================================================================================================
namespace fwd {
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 6, S_FULL = 7, P_READY = 8, P_FREE = 9, O_FULL = 10, O_FREE = 12, NBAR = 14 };
static constexpr uint32_t OFF_Q = 0, OFF_K = T128, OFF_V = 3 * T128, OFF_P = 4 * T128, OFF_BAR = 6 * T128;
static constexpr uint32_t SMEM = OFF_BAR + 128;
}  // namespace fwd

__global__ void __launch_bounds__(192, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ o, float* __restrict__ lse, int t, int D,
                   float scale_log2) {
    using namespace fwd;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem + OFF_Q;
    uint8_t* sK = smem + OFF_K;    // 2 stages
    uint8_t* sV = smem + OFF_V;    // 1 stage
    uint8_t* sP = smem + OFF_P;    // 2 k-blocks of 64 keys
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int q0 = blockIdx.x * 128;
    const int n = (t + 127) / 128;
    const int row_base = s * t;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv);
        for (int i = 0; i < NBAR; ++i) mbar_init(&bar[i], (i == P_READY || i == O_FREE || i == O_FREE + 1) ? 4 : 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar[Q_FULL], T128);
            tma_load_2d(sQ, &tm_qkv, &bar[Q_FULL], h * HD, row_base + q0);
            auto load_k = [&](int j) {
                const int st = j & 1;
                mbar_wait(&bar[K_EMPTY + st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar[K_FULL + st], T128);
                tma_load_2d(sK + st * T128, &tm_qkv, &bar[K_FULL + st], D + h * HD, row_base + j * 128);
            };
            load_k(0);
            for (int j = 0; j < n; ++j) {
                if (j + 1 < n) load_k(j + 1);
                mbar_wait(&bar[V_EMPTY], (j & 1) ^ 1);
                mbar_arrive_expect_tx(&bar[V_FULL], T128);
                tma_load_2d(sV, &tm_qkv, &bar[V_FULL], 2 * D + h * HD, row_base + j * 128);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64) | IDESC_B_MN;
            mbar_wait(&bar[Q_FULL], 0);
            for (int j = 0; j <= n; ++j) {
                if (j < n) {     // S_j = Q K_j^T
                    mbar_wait(&bar[K_FULL + (j & 1)], (j >> 1) & 1);
                    if (j > 0) mbar_wait(&bar[P_READY], (j - 1) & 1);     // the row threads have read S_{j-1}
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, desc_k(sQ, k), desc_k(sK + (j & 1) * T128, k), idesc_qk, k > 0);
                    umma_commit(&bar[S_FULL]);
                    umma_commit(&bar[K_EMPTY + (j & 1)]);
                }
                if (j > 0) {     // O_i = P_i V_i into its own TMEM buffer
                    const int i = j - 1;
                    if (j == n) mbar_wait(&bar[P_READY], i & 1);
                    mbar_wait(&bar[V_FULL], i & 1);
                    mbar_wait(&bar[O_FREE + (i & 1)], ((i >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const int kvalid = min(128, t - i * 128);
                    const int ksteps = (kvalid + 15) >> 4;
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ss(tO + (i & 1) * 64, desc_k(sP + (k >> 2) * T128, k & 3), desc_mn(sV, k), idesc_pv, k > 0);
                    umma_commit(&bar[O_FULL + (i & 1)]);
                    umma_commit(&bar[V_EMPTY]);
                    umma_commit(&bar[P_FREE]);
                }
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        float m = -INFINITY, l = 0.f, corr_prev = 0.f;
        float oacc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) oacc[i] = 0.f;

        auto fold_o = [&](int i, float corr) {    // oacc = oacc * corr + O_i
            const int b = i & 1;
            mbar_wait(&bar[O_FULL + b], (i >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tO + b * 64 + c * 32 + lane_addr, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) oacc[c * 32 + e] = fmaf(oacc[c * 32 + e], corr, __uint_as_float(v[e]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[O_FREE + b]);
        };

        for (int j = 0; j < n; ++j) {
            const int kvalid = min(128, t - j * 128);
            mbar_wait(&bar[S_FULL], j & 1);
            tc_fence_after();
            // Both passes stream the 128 columns through two register buffers: the tcgen05.ld of chunk c+1 is in flight
            // while chunk c is processed; maxima / sums are kept in 4 independent accumulators (no 32-deep chains).
            const int nch = (kvalid + 31) >> 5;    // 32-column chunks holding valid keys (warp-uniform)
            uint32_t va[32], vb[32];
            // pass 1: row maximum
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            tmem_ld_32x32(tS + lane_addr, va);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nch) {
                    uint32_t (&cur)[32] = (c & 1) ? vb : va;
                    uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                    if (c + 1 < nch) tmem_ld_32x32(tS + (c + 1) * 32 + lane_addr, nxt);
                    if (c * 32 + 32 <= kvalid) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) mx4[e & 3] = fmaxf(mx4[e & 3], __uint_as_float(cur[e]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            mx4[e & 3] = (c * 32 + e < kvalid) ? fmaxf(mx4[e & 3], __uint_as_float(cur[e])) : mx4[e & 3];
                    }
                    if (c + 1 < nch) tmem_ld_wait();
                }
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const float m_new = fmaxf(m, mx * scale_log2);
            const float corr = ex2_approx(m - m_new);     // first tile: 2^(-inf) = 0
            // pass 2: P = 2^(S * scale - m_new) -> bf16 operand tile, row sum
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            tmem_ld_32x32(tS + lane_addr, va);
            if (j > 0) mbar_wait(&bar[P_FREE], (j - 1) & 1);   // P_{j-1} V_{j-1} has read the P tile
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nch) {
                    uint32_t (&cur)[32] = (c & 1) ? vb : va;
                    uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                    if (c + 1 < nch) tmem_ld_32x32(tS + (c + 1) * 32 + lane_addr, nxt);
                    uint8_t* blk = sP + (c >> 1) * T128;
                    const bool full = c * 32 + 32 <= kvalid;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float p[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float x = ex2_approx(fmaf(__uint_as_float(cur[q * 8 + e]), scale_log2, -m_new));
                            p[e] = (full || c * 32 + q * 8 + e < kvalid) ? x : 0.f;
                            sum4[e & 3] += p[e];
                        }
                        st_tile_chunk(blk, r, (c & 1) * 4 + q, pack8(p));
                    }
                    if (c + 1 < nch) tmem_ld_wait();
                }
            }
            const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
            l = fmaf(l, corr, sum);
            m = m_new;
            tc_fence_before();
            fence_proxy_async();     // the generic-proxy stores above are read by the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[P_READY]);
            if (j > 0) fold_o(j - 1, corr_prev);
            corr_prev = corr;
        }
        fold_o(n - 1, corr_prev);

        const int q = q0 + r;
        if (q < t) {
            const float inv = 1.f / l;
            uint4* dst = reinterpret_cast<uint4*>(o + (size_t)(row_base + q) * D + h * HD);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = oacc[c * 8 + e] * inv;
                dst[c] = pack8(v);
            }
            lse[((size_t)s * H + h) * t + q] = m + log2f(l);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

// =================================================================================================================
// backward, part 0: delta[s,h,q] = sum_d dO[q, h*64 + d] * O[q, h*64 + d]  (one warp per token row; the 8 lanes that hold
// the eight 16-byte chunks of a head reduce by shuffle). Lets the dQ and the dK/dV kernels run side by side.
// =================================================================================================================
__global__ void __launch_bounds__(128) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                                         int S, int t, int D) {
    pdl_sync();
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= S * t) return;
    const int s = row / t, q = row - s * t, H = D / HD;
    const uint4* po = reinterpret_cast<const uint4*>(o + (size_t)row * D);
    const uint4* pd = reinterpret_cast<const uint4*>(dout + (size_t)row * D);
    for (int c0 = 0; c0 < D / 8; c0 += 32) {       // D / 8 chunks of 8 bf16; D is a multiple of 64 => whole heads per pass
        const int c = c0 + lane;
        float dl = 0.f;
        if (c < D / 8) {
            const uint4 a = po[c], b = pd[c];
            float2 x, y;
            x = unpack_bf16x2(a.x); y = unpack_bf16x2(b.x); dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
            x = unpack_bf16x2(a.y); y = unpack_bf16x2(b.y); dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
            x = unpack_bf16x2(a.z); y = unpack_bf16x2(b.z); dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
            x = unpack_bf16x2(a.w); y = unpack_bf16x2(b.w); dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
        }
        dl += __shfl_xor_sync(0xffffffffu, dl, 1);
        dl += __shfl_xor_sync(0xffffffffu, dl, 2);
        dl += __shfl_xor_sync(0xffffffffu, dl, 4);
        if ((lane & 7) == 0 && c < D / 8) delta[((size_t)s * H + (c >> 3)) * t + q] = dl;
    }
}

// =================================================================================================================
// backward, part 1: dQ. CTA = (128-query tile, head, sequence); key tiles of 64
// =================================================================================================================
namespace bdq {
enum { QDO_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, S_FULL = 5, DS_READY = 6, DS_FREE = 7, DQ_FULL = 8, NBAR = 9 };
static constexpr uint32_t OFF_Q = 0, OFF_DO = T128, OFF_K = 2 * T128, OFF_V = 2 * T128 + 2 * T64, OFF_DS = 2 * T128 + 4 * T64,
                          OFF_BAR = 3 * T128 + 4 * T64;
static constexpr uint32_t SMEM = OFF_BAR + 128;
}  // namespace bdq

__global__ void __launch_bounds__(192, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                      const __grid_constant__ CUtensorMap tm_do128, const float* __restrict__ lse,
                      const float* __restrict__ delta, bf16* __restrict__ dqkv, int t, int D, float scale, float scale_log2) {
    using namespace bdq;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem + OFF_Q;
    uint8_t* sdO = smem + OFF_DO;
    uint8_t* sK = smem + OFF_K;     // 2 stages of 64 keys
    uint8_t* sV = smem + OFF_V;     // 2 stages
    uint8_t* sdS = smem + OFF_DS;   // [128 q][64 keys]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int q0 = blockIdx.x * 128;
    const int n = (t + 63) / 64;
    const int row_base = s * t;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv128);
        tma_prefetch_desc(&tm_qkv64);
        tma_prefetch_desc(&tm_do128);
        for (int i = 0; i < NBAR; ++i) mbar_init(&bar[i], (i == DS_READY) ? 4 : 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdQ = tmem_base + 128;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar[QDO_FULL], 2 * T128);
            tma_load_2d(sQ, &tm_qkv128, &bar[QDO_FULL], h * HD, row_base + q0);
            tma_load_2d(sdO, &tm_do128, &bar[QDO_FULL], h * HD, row_base + q0);
            for (int j = 0; j < n; ++j) {
                const int st = j & 1;
                mbar_wait(&bar[KV_EMPTY + st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar[KV_FULL + st], 2 * T64);
                tma_load_2d(sK + st * T64, &tm_qkv64, &bar[KV_FULL + st], D + h * HD, row_base + j * 64);
                tma_load_2d(sV + st * T64, &tm_qkv64, &bar[KV_FULL + st], 2 * D + h * HD, row_base + j * 64);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
            constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64) | IDESC_B_MN;
            mbar_wait(&bar[QDO_FULL], 0);
            for (int j = 0; j <= n; ++j) {
                if (j < n) {     // S_j = Q K_j^T, dP_j = dO V_j^T
                    mbar_wait(&bar[KV_FULL + (j & 1)], (j >> 1) & 1);
                    if (j > 0) mbar_wait(&bar[DS_READY], (j - 1) & 1);   // the row threads have read S_{j-1}, dP_{j-1}
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, desc_k(sQ, k), desc_k(sK + (j & 1) * T64, k), idesc_s, k > 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tdP, desc_k(sdO, k), desc_k(sV + (j & 1) * T64, k), idesc_s, k > 0);
                    umma_commit(&bar[S_FULL]);
                }
                if (j > 0) {     // dQ += dS_i K_i
                    const int i = j - 1;
                    if (j == n) mbar_wait(&bar[DS_READY], i & 1);
                    tc_fence_after();
                    const int kvalid = min(64, t - i * 64);
                    const int ksteps = (kvalid + 15) >> 4;
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ss(tdQ, desc_k(sdS, k), desc_mn(sK + (i & 1) * T64, k), idesc_dq, (i | k) != 0);
                    umma_commit(&bar[KV_EMPTY + (i & 1)]);
                    umma_commit(&bar[DS_FREE]);
                }
            }
            umma_commit(&bar[DQ_FULL]);
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const int q = q0 + r;
        const bool qok = q < t;
        const float dl = qok ? delta[((size_t)s * H + h) * t + q] : 0.f;
        const float lse_r = qok ? lse[((size_t)s * H + h) * t + q] : INFINITY;   // +inf => P = 0 for padded query rows

        for (int j = 0; j < n; ++j) {
            const int kvalid = min(64, t - j * 64);
            mbar_wait(&bar[S_FULL], j & 1);
            tc_fence_after();
            if (j > 0) mbar_wait(&bar[DS_FREE], (j - 1) & 1);    // dS_{j-1} K_{j-1} has read the dS tile
            {
                const int nch = (((kvalid + 15) & ~15) + 31) >> 5;   // 32-column chunks the dQ MMA will read (warp-uniform)
                uint32_t s0[32], d0[32], s1[32], d1[32];
                tmem_ld_32x32(tS + lane_addr, s0);
                tmem_ld_32x32(tdP + lane_addr, d0);
                tmem_ld_wait();
                if (nch > 1) {     // chunk 1 in flight while chunk 0 is processed
                    tmem_ld_32x32(tS + 32 + lane_addr, s1);
                    tmem_ld_32x32(tdP + 32 + lane_addr, d1);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c < nch) {
                        uint32_t (&sv)[32] = c ? s1 : s0;
                        uint32_t (&dv)[32] = c ? d1 : d0;
                        if (c == 1) tmem_ld_wait();
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            float ds[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float p = ex2_approx(fmaf(__uint_as_float(sv[qq * 8 + e]), scale_log2, -lse_r));
                                const float g = p * (__uint_as_float(dv[qq * 8 + e]) - dl);
                                ds[e] = (c * 32 + qq * 8 + e < kvalid) ? g : 0.f;
                            }
                            st_tile_chunk(sdS, r, c * 4 + qq, pack8(ds));
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[DS_READY]);
        }
        mbar_wait(&bar[DQ_FULL], 0);
        tc_fence_after();
        {
            uint4* dst = reinterpret_cast<uint4*>(dqkv + (size_t)(row_base + q) * 3 * D + h * HD);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tdQ + c * 32 + lane_addr, v);
                tmem_ld_wait();
                if (qok) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[qq * 8 + e]) * scale;
                        dst[c * 4 + qq] = pack8(f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

// =================================================================================================================
// backward, part 2: dK and dV. CTA = (128-key tile, head, sequence); query tiles of 64; accumulators stay in TMEM
// =================================================================================================================
namespace bkv {
enum { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = 3, S_FULL = 5, P_READY = 6, P_FREE = 7, ACC_FULL = 8, NBAR = 9 };
static constexpr uint32_t OFF_K = 0, OFF_V = T128, OFF_Q = 2 * T128, OFF_DO = 2 * T128 + 2 * T64, OFF_PT = 2 * T128 + 4 * T64,
                          OFF_DST = 3 * T128 + 4 * T64, OFF_BAR = 4 * T128 + 4 * T64, OFF_STAT = OFF_BAR + 128;
static constexpr uint32_t SMEM = OFF_STAT + 2 * 2 * 64 * 4;   // (lse, delta) x 2 stages x 64 queries
}  // namespace bkv

__global__ void __launch_bounds__(192, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                       const __grid_constant__ CUtensorMap tm_do64, const float* __restrict__ lse, const float* __restrict__ delta,
                       bf16* __restrict__ dqkv, int t, int D, float scale, float scale_log2) {
    using namespace bkv;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sK = smem + OFF_K;
    uint8_t* sV = smem + OFF_V;
    uint8_t* sQ = smem + OFF_Q;      // 2 stages of 64 queries
    uint8_t* sdO = smem + OFF_DO;    // 2 stages
    uint8_t* sPt = smem + OFF_PT;    // [128 keys][64 q]
    uint8_t* sdSt = smem + OFF_DST;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + NBAR);
    float* s_stat = reinterpret_cast<float*>(smem + OFF_STAT);   // [2 stages][lse 64 | delta 64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, s = blockIdx.z, H = gridDim.y;
    const int k0 = blockIdx.x * 128;
    const int n = (t + 63) / 64;
    const int row_base = s * t;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv128);
        tma_prefetch_desc(&tm_qkv64);
        tma_prefetch_desc(&tm_do64);
        for (int i = 0; i < NBAR; ++i) mbar_init(&bar[i], (i == P_READY) ? 4 : 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tSt = tmem_base, tdPt = tmem_base + 64, tdV = tmem_base + 128, tdK = tmem_base + 192;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar[KV_FULL], 2 * T128);
            tma_load_2d(sK, &tm_qkv128, &bar[KV_FULL], D + h * HD, row_base + k0);
            tma_load_2d(sV, &tm_qkv128, &bar[KV_FULL], 2 * D + h * HD, row_base + k0);
            for (int i = 0; i < n; ++i) {
                const int st = i & 1;
                mbar_wait(&bar[QDO_EMPTY + st], ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar[QDO_FULL + st], 2 * T64);
                tma_load_2d(sQ + st * T64, &tm_qkv64, &bar[QDO_FULL + st], h * HD, row_base + i * 64);
                tma_load_2d(sdO + st * T64, &tm_do64, &bar[QDO_FULL + st], h * HD, row_base + i * 64);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, 64) | IDESC_B_MN;
            mbar_wait(&bar[KV_FULL], 0);
            for (int j = 0; j <= n; ++j) {
                if (j < n) {     // S^T = K Q_j^T, dP^T = V dO_j^T
                    mbar_wait(&bar[QDO_FULL + (j & 1)], (j >> 1) & 1);
                    if (j > 0) mbar_wait(&bar[P_READY], (j - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tSt, desc_k(sK, k), desc_k(sQ + (j & 1) * T64, k), idesc_s, k > 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tdPt, desc_k(sV, k), desc_k(sdO + (j & 1) * T64, k), idesc_s, k > 0);
                    umma_commit(&bar[S_FULL]);
                }
                if (j > 0) {     // dV += P^T dO_i, dK += dS^T Q_i
                    const int i = j - 1;
                    if (j == n) mbar_wait(&bar[P_READY], i & 1);
                    tc_fence_after();
                    const int qvalid = min(64, t - i * 64);
                    const int ksteps = (qvalid + 15) >> 4;
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ss(tdV, desc_k(sPt, k), desc_mn(sdO + (i & 1) * T64, k), idesc_acc, (i | k) != 0);
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16_ss(tdK, desc_k(sdSt, k), desc_mn(sQ + (i & 1) * T64, k), idesc_acc, (i | k) != 0);
                    umma_commit(&bar[QDO_EMPTY + (i & 1)]);
                    umma_commit(&bar[P_FREE]);
                }
            }
            umma_commit(&bar[ACC_FULL]);
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;      // key row of this thread
        const int tid = threadIdx.x - 64;    // 0..127 among the row threads
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const float* L = lse + ((size_t)s * H + h) * t;
        const float* Dl = delta + ((size_t)s * H + h) * t;
        // (lse | delta) of the 64 queries of a tile: fetched one tile ahead into a register and parked in shared memory at
        // the start of the tile that uses them, so that the global-load latency never sits between two tiles
        // (+inf lse => P = 0 for query columns beyond the sequence)
        auto fetch_stats = [&](int i) -> float {
            const int qi = i * 64 + (tid & 63);
            if (tid < 64) return (qi < t) ? L[qi] : INFINITY;
            return (qi < t) ? Dl[qi] : 0.f;
        };
        float stat_next = fetch_stats(0);
        for (int i = 0; i < n; ++i) {
            s_stat[(i & 1) * 128 + tid] = stat_next;
            rows_barrier();                    // stats of tile i visible; everyone is past tile i-2's reads of this stage
            if (i + 1 < n) stat_next = fetch_stats(i + 1);
            const float* st_lse = s_stat + (i & 1) * 128;
            const float* st_dl = st_lse + 64;
            const int qvalid = min(64, t - i * 64);
            mbar_wait(&bar[S_FULL], i & 1);
            tc_fence_after();
            if (i > 0) mbar_wait(&bar[P_FREE], (i - 1) & 1);
            {
                const int nch = (((qvalid + 15) & ~15) + 31) >> 5;
                uint32_t s0[32], d0[32], s1[32], d1[32];
                tmem_ld_32x32(tSt + lane_addr, s0);
                tmem_ld_32x32(tdPt + lane_addr, d0);
                tmem_ld_wait();
                if (nch > 1) {
                    tmem_ld_32x32(tSt + 32 + lane_addr, s1);
                    tmem_ld_32x32(tdPt + 32 + lane_addr, d1);
                }
                const float4* lse4 = reinterpret_cast<const float4*>(st_lse);
                const float4* dl4 = reinterpret_cast<const float4*>(st_dl);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c < nch) {
                        uint32_t (&sv)[32] = c ? s1 : s0;
                        uint32_t (&dv)[32] = c ? d1 : d0;
                        if (c == 1) tmem_ld_wait();
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            float ls[8], dl[8], p[8], ds[8];
                            *reinterpret_cast<float4*>(ls) = lse4[c * 8 + qq * 2];
                            *reinterpret_cast<float4*>(ls + 4) = lse4[c * 8 + qq * 2 + 1];
                            *reinterpret_cast<float4*>(dl) = dl4[c * 8 + qq * 2];
                            *reinterpret_cast<float4*>(dl + 4) = dl4[c * 8 + qq * 2 + 1];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                p[e] = ex2_approx(fmaf(__uint_as_float(sv[qq * 8 + e]), scale_log2, -ls[e]));
                                ds[e] = p[e] * (__uint_as_float(dv[qq * 8 + e]) - dl[e]);
                            }
                            st_tile_chunk(sPt, r, c * 4 + qq, pack8(p));
                            st_tile_chunk(sdSt, r, c * 4 + qq, pack8(ds));
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[P_READY]);
        }
        mbar_wait(&bar[ACC_FULL], 0);
        tc_fence_after();
        const int key = k0 + r;
        const bool kok = key < t;
        uint4* dstK = reinterpret_cast<uint4*>(dqkv + (size_t)(row_base + key) * 3 * D + D + h * HD);
        uint4* dstV = reinterpret_cast<uint4*>(dqkv + (size_t)(row_base + key) * 3 * D + 2 * D + h * HD);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t kv[32], vv[32];
            tmem_ld_32x32(tdK + c * 32 + lane_addr, kv);
            tmem_ld_32x32(tdV + c * 32 + lane_addr, vv);
            tmem_ld_wait();
            if (kok) {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    float f[8], g[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        f[e] = __uint_as_float(kv[qq * 8 + e]) * scale;
                        g[e] = __uint_as_float(vv[qq * 8 + e]);
                    }
                    dstK[c * 4 + qq] = pack8(f);
                    dstV[c * 4 + qq] = pack8(g);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

// =================================================================================================================
// host
// =================================================================================================================
template <typename K>
static int set_smem(K kernel, uint32_t bytes) {
    SPLICE_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return SPLICE_OK;
}

int attention_fwd_tc(const bf16* qkv, bf16* o, float* lse, int S, int t, int D, int H, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        int rc = set_smem(attn_fwd_tc_kernel, align_slack(fwd::SMEM));
        if (rc) return rc;
        attr = true;
    }
    SPLICE_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)o & 15) == 0 && D % 8 == 0, "attention: qkv / o must be 16-byte aligned");
    CUtensorMap tm;
    int rc = make_tmap_bf16(&tm, qkv, S * t, 3 * D, 3 * D, 128);
    if (rc) return rc;
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e)
    dim3 grid(ceil_div(t, 128), H, S);
    SPLICE_CHECK_CUDA(launch_pdl(attn_fwd_tc_kernel, grid, dim3(192), align_slack(fwd::SMEM), stream, tm, o, lse, t, D, scale_log2));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int attention_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int S, int t,
                     int D, int H, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        int rc = set_smem(attn_bwd_dq_tc_kernel, align_slack(bdq::SMEM));
        if (rc) return rc;
        rc = set_smem(attn_bwd_dkv_tc_kernel, align_slack(bkv::SMEM));
        if (rc) return rc;
        attr = true;
    }
    SPLICE_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)o & 15) == 0 && ((uintptr_t)dout & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
                   "attention: operands must be 16-byte aligned");
    CUtensorMap tq128, tq64, td128, td64;
    int rc;
    if ((rc = make_tmap_bf16(&tq128, qkv, S * t, 3 * D, 3 * D, 128))) return rc;
    if ((rc = make_tmap_bf16(&tq64, qkv, S * t, 3 * D, 3 * D, 64))) return rc;
    if ((rc = make_tmap_bf16(&td128, dout, S * t, D, D, 128))) return rc;
    if ((rc = make_tmap_bf16(&td64, dout, S * t, D, D, 64))) return rc;
    const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
    SPLICE_CHECK_CUDA(launch_pdl(attn_delta_kernel, dim3(ceil_div(S * t, 4)), dim3(128), 0, stream, o, dout, delta, S, t, D));
    SPLICE_LAUNCH_CHECK();
    // dQ and dK/dV only share read-only inputs: the dK/dV kernel runs on a side stream (a parallel branch when the
    // caller is capturing a graph); one CTA of each fits an SM, so each hides the other's tensor-core / exponential phases
    static cudaStream_t side = nullptr;
    static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (!side) {
        SPLICE_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        SPLICE_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    dim3 grid(ceil_div(t, 128), H, S);
    SPLICE_CHECK_CUDA(cudaEventRecord(ev_fork, stream));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    SPLICE_CHECK_CUDA(launch_pdl(attn_bwd_dkv_tc_kernel, grid, dim3(192), align_slack(bkv::SMEM), side, tq128, tq64, td64, lse,
                                 (const float*)delta, dqkv, t, D, scale, scale_log2));
    SPLICE_LAUNCH_CHECK();
    SPLICE_CHECK_CUDA(launch_pdl(attn_bwd_dq_tc_kernel, grid, dim3(192), align_slack(bdq::SMEM), stream, tq128, tq64, td128, lse,
                                 (const float*)delta, dqkv, t, D, scale, scale_log2));
    SPLICE_LAUNCH_CHECK();
    SPLICE_CHECK_CUDA(cudaEventRecord(ev_join, side));
    SPLICE_CHECK_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    return SPLICE_OK;
}

}  // namespace splice
