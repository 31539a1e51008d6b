// Legacy tensor-core path on B200: issue rate of mma.sync m16n8k8 TF32 and m16n8k16 BF16 (fp32 accumulate), all SMs busy,
// 8 warps per CTA x 2 CTAs per SM, 8 independent accumulator tiles per warp. Prints dense TFLOP/s and FMA/clk/SM.
// Decides whether a 3xTF32 / 3xBF16 split implicit-GEMM convolution on mma.sync can beat the FP32 FMA pipe (128 FMA/clk/SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mma_sync_rate.bin tools/ubench/mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0 = tf32 m16n8k8, 1 = bf16 m16n8k16, 2 = fp32 FFMA reference
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters) {
    float acc[8][4];
    for (int t = 0; t < 8; ++t)
        for (int j = 0; j < 4; ++j) acc[t][j] = threadIdx.x * 1e-9f + t;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[t][j] = fmaf(acc[t][j], 1.0000001f, 1e-7f);
            }
        }
    }
    float s = 0.f;
    for (int t = 0; t < 8; ++t)
        for (int j = 0; j < 4; ++j) s += acc[t][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double fma_per_instr_per_warp) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 2, iters = 20000;
    float* out;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    rate_kernel<MODE><<<grid, 256>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)grid * 8 * iters * 8 * (MODE == 2 ? 4 : 1);     // warp-level instructions
    const double fma = instr * fma_per_instr_per_warp;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.1f TFLOP/s dense  %7.1f FMA/clk/SM (at %d MHz nominal)\n", name, ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("mma.sync m16n8k8 tf32", 16.0 * 8 * 8);
    run<1>("mma.sync m16n8k16 bf16", 16.0 * 8 * 16);
    run<2>("FFMA fp32", 32.0);
    return 0;
}
