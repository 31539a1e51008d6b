// splice_b200 — 3x3 stride-1 "same" convolution of the generator as an implicit GEMM on tcgen05 (kind::tf32), forward and
// data gradient. Replaces nn.Conv2d (models/unet/common.py:99-124) for the high-resolution layers that fill the GPU.
//
//   D[128 pixels, N = Cout] = A[128 pixels, K = 9 * Cp] * B[N, K]^T,   k = (ky * 3 + kx) * Cp + ci,  Cp = Cin rounded up to 4
//
// fp32 accuracy on tensor cores: every operand is split into two TF32 numbers, a = hi + lo with hi = a truncated to
// TF32 (the 19 bits the tensor core reads) and lo = a - hi (exact in fp32), and three products are accumulated in fp32
// (TMEM): hi*hi + hi*lo + lo*hi. The dropped lo*lo term and lo's own truncation are 2^-22 relative - the generator's
// 1e-4 pixel tolerance and its ill-conditioned BatchNorm backward need that (3 x BF16 would give 2^-16).
//
// The tensor work is almost free at these shapes (N = 16 ... 144): the kernel is bound by BUILDING the A operand. A CTA
// owns a 4 x 32 pixel tile; thread m is pixel m and TMEM lane m. Per k-block of 32 reduction elements a thread gathers its
// 32 im2col values (producer BatchNorm + LeakyReLU and the zero padding applied on the fly; global loads, L1-resident across
// the nine taps), splits them and writes its two 128-byte rows (A_hi, A_lo) in the 128B-swizzled K-major layout that
// tcgen05.mma reads; the CTA's threads split the weight k-block the same way (B_hi, B_lo; transposed + flipped on the fly
// for the data gradient). One thread issues 4 k-steps x 3 MMAs and commits to an mbarrier; two such CTAs per SM overlap one
// CTA's operand build with the other's MMAs. Instruction count per (pixel, ci, tap): ~10, independent of Cout - the
// SIMT tiled kernel needs ~1.1 per output channel.
// Measured (B200, netG at 896 px, SPLICE_B200_GEN_TC=1): 3-7e-6 of fp64 conv2d, every generator check green; forward
// 1.6x SLOWER than the SIMT tiled kernel (long reductions: every element a global gather + producer transform), data
// gradient 5 % faster (short reductions, wide N). Off by default; profiles/ANALYSIS_r2.md lists what it needs next
// (input tile staged in shared memory, double-buffered operand tiles, TMA im2col from NHWC copies).
//
// Epilogue (forward): tcgen05.ld of the accumulator row, bias, store (a warp = 32 consecutive pixels of a row: coalesced
// per channel), per-tile (count, mean, M2) BatchNorm partials and the last-ticket merge, exactly like the other kernels.
#include "conv_tc.h"

namespace splice {

static constexpr int TCM = 128;          // pixels per CTA = UMMA M
static constexpr int TCK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
static constexpr int TC_TR = 4, TC_TW = 32;

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: tf32 x tf32 -> fp32, both operands K-major (cute/arch/mma_sm100_desc.hpp: F16F32Format TF32 = 2)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// thread-row `row` of a K-major 128B-swizzled tile: 16-byte chunk `chunk` (4 fp32, chunk < 8)
__device__ __forceinline__ void st_row_chunk(uint8_t* tile, int row, int chunk, float4 v) {
    *reinterpret_cast<float4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

template <bool DGRAD>
__global__ void __launch_bounds__(128, 2)
conv_tc_kernel(const float* __restrict__ x, int Cin, int H, int W, InTf tf, const float* __restrict__ Wt, int w_cout, int w_cin,
               const float* __restrict__ bias, int Cout, float* __restrict__ y, int accumulate, float* __restrict__ stats_part,
               BnFin fin, int NT, int tmem_cols) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + TCM * 128;
    uint8_t* b_hi = smem + 2 * TCM * 128;
    uint8_t* b_lo = b_hi + ((NT * 128 + 1023) & ~1023);
    uint8_t* tail = b_lo + ((NT * 128 + 1023) & ~1023);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tail + 16);
    int* s_flag = reinterpret_cast<int*>(tail + 32);
    float* red = reinterpret_cast<float*>(tail + 64);          // [4 warps][NT]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (W + TC_TW - 1) / TC_TW;
    const int ty0 = (blockIdx.x / tiles_x) * TC_TR, tx0 = (blockIdx.x % tiles_x) * TC_TW;
    const int n = blockIdx.z;
    const int oy = ty0 + warp, ox = tx0 + lane;                // thread = pixel = TMEM lane
    const bool pix_ok = oy < H && ox < W;
    const size_t plane = (size_t)H * W;
    const float* xn = x + (size_t)n * Cin * plane;
    const int Ktot = ((Cin + 3) & ~3) * 9, num_kb = (Ktot + TCK - 1) / TCK;

    if (tid == 0) {
        mbar_init(mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, (uint32_t)tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t idesc = make_idesc_tf32(TCM, NT);

    // tap geometry of this pixel: offsets and validity of the 3 rows / 3 columns
    int roff[3], coff[3];
    bool rok[3], cok[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int iy = oy + k - 1, ix = ox + k - 1;
        rok[k] = pix_ok && iy >= 0 && iy < H;
        cok[k] = ix >= 0 && ix < W;
        roff[k] = min(max(iy, 0), H - 1) * W;
        coff[k] = min(max(ix, 0), W - 1);
    }

    // Reduction order: k = tap * Cp + ci with Cp = Cin rounded up to 4 (tap-major): the 4 elements of a 16-byte chunk share one
    // tap - one offset, one validity bit - and are 4 consecutive channels; (tap, ci) advance without divisions.
    // (The first version used k = ci * 9 + tap: every element decoded its own tap - ~30 instructions per element.)
    const int Cp = (Cin + 3) & ~3;
    unsigned okmask = 0;      // bit tap: the tap's source pixel lies inside the image
#pragma unroll
    for (int t = 0; t < 9; ++t) okmask |= (rok[t / 3] && cok[t % 3]) ? (1u << t) : 0u;
    int ci = 0, tap = 0;      // (first channel, tap) of this thread's next chunk
    for (int kb = 0; kb < num_kb; ++kb) {
        if (kb > 0) {         // the MMAs of the previous k-block have read the operand tiles
            mbar_wait(mma_bar, (uint32_t)((kb - 1) & 1));
            tc_fence_after();
        }
        // ---- A: this pixel's 32 im2col values of the k-block, split, swizzled rows
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (tap < 9 && ((okmask >> tap) & 1u)) {
                const int ky = tap / 3, kx = tap - ky * 3;
                const int off = (ky == 0 ? roff[0] : ky == 1 ? roff[1] : roff[2]) + (kx == 0 ? coff[0] : kx == 1 ? coff[1] : coff[2]);
                const float* src = xn + (size_t)ci * plane + off;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ci + j < Cin) {
                        float t = src[(size_t)j * plane];
                        if (!DGRAD && tf.k) {
                            const float4 k4 = __ldg(tf.k + ci + j);
                            t = fmaf(k4.z, t, k4.w);
                            if (tf.lrelu) t = t < 0.f ? t * LRELU : t;
                        }
                        v[j] = t;
                    }
            }
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = tf32_hi(v[j]);
                lo[j] = v[j] - hi[j];
            }
            st_row_chunk(a_hi, tid, c, make_float4(hi[0], hi[1], hi[2], hi[3]));
            st_row_chunk(a_lo, tid, c, make_float4(lo[0], lo[1], lo[2], lo[3]));
            ci += 4;
            if (ci >= Cp) { ci = 0; ++tap; }
        }
        // ---- B: the weight k-block [NT rows = output channels][32] in the same k order, split the same way
        for (int idx = tid; idx < NT * 8; idx += 128) {
            const int o = idx >> 3, c = idx & 7;
            const int k0 = kb * TCK + c * 4;
            const int tp = k0 / Cp, cr0 = k0 - tp * Cp;          // one tap, 4 consecutive reduction channels
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cr = cr0 + j;
                float wv = 0.f;
                if (tp < 9 && cr < Cin && o < Cout) {
                    if (DGRAD) wv = __ldg(Wt + ((size_t)cr * w_cin + o) * 9 + (8 - tp));      // transposed, flipped
                    else wv = __ldg(Wt + ((size_t)o * w_cin + cr) * 9 + tp);
                }
                hi[j] = tf32_hi(wv);
                lo[j] = wv - hi[j];
            }
            st_row_chunk(b_hi, o, c, make_float4(hi[0], hi[1], hi[2], hi[3]));
            st_row_chunk(b_lo, o, c, make_float4(lo[0], lo[1], lo[2], lo[3]));
        }
        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t dah = make_sw128_kmajor_desc(smem_u32(a_hi)), dal = make_sw128_kmajor_desc(smem_u32(a_lo));
            const uint64_t dbh = make_sw128_kmajor_desc(smem_u32(b_hi)), dbl = make_sw128_kmajor_desc(smem_u32(b_lo));
#pragma unroll
            for (int ks = 0; ks < TCK / 8; ++ks) {      // 8 tf32 = 32 bytes per MMA: +2 in the (addr >> 4) field
                umma_tf32_ss(tmem_base, dah + 2u * ks, dbh + 2u * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
                umma_tf32_ss(tmem_base, dah + 2u * ks, dbl + 2u * ks, idesc, 1u);
                umma_tf32_ss(tmem_base, dal + 2u * ks, dbh + 2u * ks, idesc, 1u);
            }
            umma_commit(mma_bar);
        }
    }
    mbar_wait(mma_bar, (uint32_t)((num_kb - 1) & 1));
    tc_fence_after();

    // ---- epilogue: accumulator row of this pixel, 16 channels at a time
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float* yp = y + (size_t)n * Cout * plane + (size_t)min(oy, H - 1) * W + min(ox, W - 1);
    if (DGRAD) {
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(lane_addr + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < Cout && pix_ok) {
                    float* o = yp + (size_t)(c0 + j) * plane;
                    const float v = __uint_as_float(r[j]);
                    *o = accumulate ? *o + v : v;
                }
        }
    } else {
        const float cnt = (float)(min(TC_TR, H - ty0) * min(TC_TW, W - tx0));
        // pass 1: bias, store, per-channel sums
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(lane_addr + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int co = c0 + j;
                float v = 0.f;
                if (co < Cout) {
                    v = __uint_as_float(r[j]) + __ldg(bias + co);
                    if (pix_ok) yp[(size_t)co * plane] = v;
                }
                if (stats_part) {
                    const float sv = warp_sum(pix_ok ? v : 0.f);
                    if (lane == 0) red[warp * NT + co] = sv;
                }
            }
        }
        if (stats_part) {
            __syncthreads();
            // pass 2: centred sum of squares (the accumulator is read again from TMEM)
            for (int c0 = 0; c0 < NT; c0 += 16) {
                uint32_t r[16];
                tmem_ld_32x16(lane_addr + (uint32_t)c0, r);
                tmem_ld_wait();
                float m2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int co = c0 + j;
                    const float mean = (red[co] + red[NT + co] + red[2 * NT + co] + red[3 * NT + co]) / cnt;
                    const float d = (co < Cout ? __uint_as_float(r[j]) + __ldg(bias + co) : 0.f) - mean;
                    m2[j] = warp_sum(pix_ok ? d * d : 0.f);
                }
                // a second scratch row per warp holds the M2 partials: [4 warps][NT] after the sums' [4][NT]
                if (lane == 0)
#pragma unroll
                    for (int j = 0; j < 16; ++j) red[(4 + warp) * NT + c0 + j] = m2[j];
            }
            __syncthreads();
            const int part = blockIdx.z * gridDim.x + blockIdx.x, nparts = gridDim.x * gridDim.z;
            for (int co = tid; co < Cout; co += 128) {
                const float mean = (red[co] + red[NT + co] + red[2 * NT + co] + red[3 * NT + co]) / cnt;
                const float m2 = red[4 * NT + co] + red[5 * NT + co] + red[6 * NT + co] + red[7 * NT + co];
                float* o = stats_part + ((size_t)part * Cout + co) * 3;
                o[0] = cnt; o[1] = mean; o[2] = m2;
            }
            __syncthreads();
            if (fin.konst) bn_finish_if_last(stats_part, nparts, Cout, 0, Cout, 0, nparts, fin, s_flag);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

template <bool DGRAD>
int launch_conv_tc(const float* x, int N, int Cin, int H, int W, InTf tf, const float* Wt, int w_cout, int w_cin, const float* bias,
                   int Cout, float* y, int accumulate, float* stats_part, BnFin fin, cudaStream_t st) {
    const int NT = (Cout + 15) & ~15;
    SPLICE_REQUIRE(NT <= 256, "conv_tc: %d output channels", Cout);
    const int tmem_cols = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
    const size_t smem = 1024 + 2 * TCM * 128 + 2 * (size_t)((NT * 128 + 1023) & ~1023) + 64 + (size_t)8 * NT * sizeof(float);
    static bool attr[2] = {false, false};
    if (!attr[DGRAD]) {
        SPLICE_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr[DGRAD] = true;
    }
    dim3 grid(ceil_div(W, TC_TW) * ceil_div(H, TC_TR), 1, N);
    SPLICE_CHECK_CUDA(launch_pdl(conv_tc_kernel<DGRAD>, grid, dim3(128), smem, st, x, Cin, H, W, tf, Wt, w_cout, w_cin, bias, Cout, y,
                                 accumulate, stats_part, fin, NT, tmem_cols));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}
template int launch_conv_tc<false>(const float*, int, int, int, int, InTf, const float*, int, int, const float*, int, float*, int, float*,
                                   BnFin, cudaStream_t);
template int launch_conv_tc<true>(const float*, int, int, int, int, InTf, const float*, int, int, const float*, int, float*, int, float*,
                                  BnFin, cudaStream_t);

}  // namespace splice
