// Micro-benchmarks of the tcgen05 building blocks used by attention_tc.cu (diagnostic, not part of the library):
// cycles from issuing a group of MMAs + tcgen05.commit until the mbarrier flips, for K-major / MN-major B operands and
// A from shared or tensor memory; cost of fence.proxy.async; cost of the swizzled operand-tile stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I splice_b200/csrc tools/ubench/tc_latency.cu -o /tmp/tc_latency
#include <cstdio>
#include "common.cuh"
using namespace splice;

__device__ __forceinline__ uint64_t desc_k(const uint8_t* tile, int ks) { return make_sw128_kmajor_desc(smem_u32(tile)) + 2u * ks; }
__device__ __forceinline__ uint64_t desc_mn(const uint8_t* tile, int ks) { return make_sw128_kmajor_desc(smem_u32(tile)) + 128u * ks; }

__global__ void __launch_bounds__(128) bench(long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;              // 128 x 64 bf16
    uint8_t* sB = smem + 16384;      // 128 x 64 bf16
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 4);
    for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(tptr, 256); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tptr;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        auto run = [&](int which, int nmma) -> long long {
            long long best = 1ll << 60;
            for (int rep = 0; rep < 20; ++rep) {
                const long long t0 = clock64();
                for (int k = 0; k < nmma; ++k) {
                    const int ks = k & 3;
                    if (which == 0) umma_bf16_ss(tb, desc_k(sA, ks), desc_k(sB, ks), make_idesc_bf16(128, 64), k > 0);           // SS, K-major, N=64
                    else if (which == 1) umma_bf16_ss(tb, desc_k(sA, ks), desc_mn(sB, ks), make_idesc_bf16(128, 64) | (1u << 16), k > 0);  // SS, MN-major B
                    else if (which == 2) umma_bf16_ts(tb, tb + 128 + 8 * ks, desc_k(sB, ks), make_idesc_bf16(128, 32), k > 0);  // TS, N=32
                    else if (which == 3) umma_bf16_ss(tb, desc_k(sA, ks), desc_k(sB, ks), make_idesc_bf16(128, 128), k > 0);    // SS, N=128
                    else umma_bf16_ss(tb, desc_k(sA, ks), desc_k(sB, ks), make_idesc_bf16(128, 32), k > 0);                     // SS, N=32
                }
                umma_commit(&bar[0]);
                while (!mbar_try_wait(&bar[0], ph)) {}
                ph ^= 1;
                const long long t1 = clock64();
                if (t1 - t0 < best) best = t1 - t0;
            }
            return best;
        };
        int o = 0;
        for (int which = 0; which < 5; ++which)
            for (int nmma : {1, 4, 8, 16, 64}) out[o++] = run(which, nmma);
        // 64 MMAs (128x128x16) round-robin over 1 / 2 independent accumulator tiles: dependency latency or issue cost?
        for (int nacc = 1; nacc <= 2; ++nacc) {
            long long best = 1ll << 60;
            for (int rep = 0; rep < 20; ++rep) {
                const long long t0 = clock64();
                for (int k = 0; k < 64; ++k) {
                    const uint32_t d = tb + ((nacc == 2 && (k & 1)) ? 128u : 0u);
                    umma_bf16_ss(d, desc_k(sA, (k >> 1) & 3), desc_k(sB, (k >> 1) & 3), make_idesc_bf16(128, 128), k >= nacc ? 1u : 0u);
                }
                umma_commit(&bar[0]);
                while (!mbar_try_wait(&bar[0], ph)) {}
                ph ^= 1;
                const long long t1 = clock64();
                if (t1 - t0 < best) best = t1 - t0;
            }
            out[29 + nacc] = best;
        }
        // commit with nothing outstanding
        { long long best = 1ll << 60; for (int r = 0; r < 20; ++r) { long long t0 = clock64(); umma_commit(&bar[0]); while (!mbar_try_wait(&bar[0], ph)) {} ph ^= 1; long long t1 = clock64(); if (t1 - t0 < best) best = t1 - t0; } out[o++] = best; }
        // fence.proxy.async after 4 x 16-byte shared stores
        { long long best = 1ll << 60; for (int r = 0; r < 20; ++r) { long long t0 = clock64(); for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(sA + q * 16) = make_uint4(r, q, 0, 0); fence_proxy_async(); long long t1 = clock64(); if (t1 - t0 < best) best = t1 - t0; } out[o++] = best; }
    }
    __syncthreads();
    // tcgen05.ld x32 + wait latency (warp 0), all lanes
    if (threadIdx.x < 32) {
        long long best = 1ll << 60; uint32_t v[32]; uint32_t acc = 0;
        for (int r = 0; r < 20; ++r) { long long t0 = clock64(); tmem_ld_32x32(tb, v); tmem_ld_wait(); acc += v[r & 31]; long long t1 = clock64(); if (t1 - t0 < best) best = t1 - t0; }
        if (threadIdx.x == 0) { out[27] = best; out[40] = acc; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tb, 256);
}

int main() {
    long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    bench<<<1, 128, 60000>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    const char* names[5] = {"SS K-major N=64", "SS MN-major B N=64", "TS N=32", "SS N=128", "SS N=32"};
    int o = 0;
    for (int w = 0; w < 5; ++w) { printf("%-20s issue->commit->barrier cycles for 1/4/8/16/64 MMAs:", names[w]); for (int i = 0; i < 5; ++i) printf(" %lld", h[o++]); printf("\n"); }
    printf("empty commit: %lld cycles\n", h[o++]);
    printf("4 x st.shared.v4 + fence.proxy.async: %lld cycles\n", h[o++]);
    printf("tcgen05.ld x32 + wait: %lld cycles\n", h[27]);
    printf("64 MMAs 128x128x16 into 1 / 2 alternating accumulator tiles: %lld / %lld cycles\n", h[30], h[31]);
    return 0;
}
