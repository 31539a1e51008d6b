// splice_b200 — device-side pieces shared by the generator's kernels (generator.cu, conv_tc.cu)
#pragma once
#ifdef SPLICE_EMU
#include "cuda_emu.h"   // tests/emu: CPU emulation of the CUDA subset these kernels use (host-side checks without a GPU)
#else
#include "common.cuh"
#endif

namespace splice {

static constexpr float LRELU = 0.2f;

struct InTf {               // per-channel transform applied to a raw tensor when it is consumed
    const float4* k;        // (mean, invstd, a, b): value -> a*value + b ; nullptr = identity
    int lrelu;
};

__device__ __forceinline__ float apply_tf(const InTf& tf, int c, float v) {
    if (tf.k) {
        const float4 k = tf.k[c];
        v = fmaf(k.z, v, k.w);
        if (tf.lrelu && v < 0.f) v *= LRELU;
    }
    return v;
}

// -------------------------------------------------------------------------------------------------
// BatchNorm statistics without a second launch: every block that produced a (count, mean, M2) partial of a channel
// group takes a ticket; the block that draws the last ticket merges the partials of that group (fixed order: the
// result does not depend on which block happens to be last) and writes the per-channel constants. The ticket counters
// live in the slot, start at zero and are reset by the finishing block, so captured graphs can be replayed.
// -------------------------------------------------------------------------------------------------
struct BnFin {
    const float* gamma;
    const float* beta;
    float4* konst;       // (mean, invstd, a, b) per channel; nullptr = no statistics wanted
    float2* bstat;       // (mean, unbiased variance) for the running-statistics update
    int* counter;        // one ticket counter per channel group of this layer
    float eps;
};

// Merge of the (count, mean, M2) partials [nparts][C][3] of channel c by one warp, in double precision and in a fixed
// order: total count and mean first, then M2 = sum(M2_i + n_i (mean_i - mean)^2) (the pairwise update of Chan et al.
// summed over all parts; no divisions inside the loops, partials fetched eight at a time so that the L2 loads overlap).
// Lane 0 writes the constants.
__device__ __forceinline__ void bn_merge_channel(const float* part, int nparts, int C, int c, const BnFin& f) {
    const int lane = threadIdx.x & 31;
    double n = 0.0, s1 = 0.0;
    for (int i0 = lane; i0 < nparts; i0 += 32 * 8) {
        float nb[8], mb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            nb[u] = 0.f; mb[u] = 0.f;
            if (i < nparts) {
                const float* p = part + ((size_t)i * C + c) * 3;
                nb[u] = __ldcg(p); mb[u] = __ldcg(p + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { n += (double)nb[u]; s1 += (double)nb[u] * (double)mb[u]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const double mean = s1 / n;
    double M2 = 0.0;
    for (int i0 = lane; i0 < nparts; i0 += 32 * 8) {
        float nb[8], mb[8], Mb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            nb[u] = 0.f; mb[u] = 0.f; Mb[u] = 0.f;
            if (i < nparts) {
                const float* p = part + ((size_t)i * C + c) * 3;
                nb[u] = __ldcg(p); mb[u] = __ldcg(p + 1); Mb[u] = __ldcg(p + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double d = (double)mb[u] - mean;
            M2 += (double)Mb[u] + (double)nb[u] * d * d;
        }
    }
    for (int o = 16; o > 0; o >>= 1) M2 += __shfl_xor_sync(0xffffffffu, M2, o);
    if (lane == 0) {
        const double var = M2 / n;
        const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
        const float a = f.gamma[c] * invstd;
        f.konst[c] = make_float4((float)mean, invstd, a, f.beta[c] - (float)mean * a);
        if (f.bstat) f.bstat[c] = make_float2((float)mean, (float)(n > 1.0 ? M2 / (n - 1.0) : var));
    }
}

// Called by ALL threads of a block after thread 0 has written the block's partials of channels [c0, c0 + nch).
// `expected` = number of blocks contributing to this channel group. Uses one int of shared memory (s_flag).
__device__ __forceinline__ void bn_finish_if_last(const float* part, int nparts, int C, int c0, int nch, int group, int expected,
                                                  const BnFin& f, int* s_flag) {
    if (threadIdx.x == 0) {
        __threadfence();                                  // partials visible before the ticket
        const int ticket = atomicAdd(f.counter + group, 1);
        const int last = ticket == expected - 1;
        if (last) f.counter[group] = 0;                   // self-reset for the next launch / graph replay
        *s_flag = last;
    }
    __syncthreads();
    if (*s_flag) {
        __threadfence();
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int j = w; j < nch; j += nw)
            if (c0 + j < C) bn_merge_channel(part, nparts, C, c0 + j, f);
    }
}


}  // namespace splice
