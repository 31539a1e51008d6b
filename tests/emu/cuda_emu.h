// splice_b200 tests - a small CPU emulation of the CUDA subset the generator kernels use (TEST INFRASTRUCTURE, never shipped:
// the product library is built by nvcc without SPLICE_EMU and has no CPU path).
//
// Why: the container that builds the library has no GPU. Compiling csrc/generator_x.cu with g++ -DSPLICE_EMU against this
// header runs the SAME kernel bodies and the SAME host orchestration (buffer planning, launch geometry, parameter tables,
// BatchNorm tickets) on the CPU, so that tests/test_genx_emu.py can compare the engine with torch's modules before a GPU
// ever sees it. What it does not model: memory-ordering / races between threads of a block (fibers run one at a time,
// switching only at barriers and shuffles), occupancy, performance.
//
// Execution model: a kernel launch = loop over the grid; a block = blockDim.x ucontext fibers on one OS thread, scheduled
// round-robin; __syncthreads() and __shfl_xor_sync() are barriers (block-wide / warp-wide) implemented by yielding until
// every live fiber of the block / warp has arrived. __shared__ variables are function-local statics (one block runs at a
// time). Streams, events, graphs and PDL are no-ops (everything is synchronous).
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <tuple>
#include <utility>
#include <vector>

#define SPLICE_OK 0
#define SPLICE_ERR_ARG -1
#define SPLICE_ERR_CUDA -2
#define SPLICE_ERR_UNSUPPORTED -3
#define SPLICE_ERR_STATE -4

// ---- language keywords --------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

// ---- runtime API subset -------------------------------------------------------------------------------
typedef int cudaError_t;
static constexpr cudaError_t cudaSuccess = 0;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaMemcpyDeviceToDevice = 3, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = static_cast<T*>(aligned_alloc(256, (bytes + 255) & ~(size_t)255));
    if (*p) memset(*p, 0xff, bytes);   // poison (NaN floats): reading something never written shows up in the comparison
    return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

// ---- the fiber scheduler ------------------------------------------------------------------------------
namespace emu {
static constexpr int MAX_THREADS = 1024, STACK_BYTES = 256 * 1024;
struct Sched {
    ucontext_t main_ctx;
    ucontext_t ctx[MAX_THREADS];
    char* stacks = nullptr;
    bool done[MAX_THREADS];
    int n = 0, cur = 0, live = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int warp_live[MAX_THREADS / 32], warp_count[MAX_THREADS / 32];
    unsigned warp_gen[MAX_THREADS / 32];
    alignas(8) unsigned char shfl[MAX_THREADS][8];
    const std::function<void()>* body = nullptr;
    long long idle_rounds = 0;
};
inline Sched& sched() { static Sched s; return s; }
}  // namespace emu

inline uint3_emu threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {
inline void yield() {
    Sched& s = sched();
    swapcontext(&s.ctx[s.cur], &s.main_ctx);
}
inline void release_if_complete(Sched& s) {
    if (s.bar_count > 0 && s.bar_count == s.live) { s.bar_count = 0; ++s.bar_gen; }
    for (int w = 0; w * 32 < s.n; ++w)
        if (s.warp_count[w] > 0 && s.warp_count[w] == s.warp_live[w]) { s.warp_count[w] = 0; ++s.warp_gen[w]; }
}
inline void trampoline() {
    Sched& s = sched();
    (*s.body)();
    s.done[s.cur] = true;
    --s.live;
    --s.warp_live[s.cur / 32];
    release_if_complete(s);   // a thread that has exited no longer takes part in barriers
}
inline void block_barrier() {
    Sched& s = sched();
    const unsigned gen = s.bar_gen;
    if (++s.bar_count == s.live) { s.bar_count = 0; ++s.bar_gen; s.idle_rounds = 0; return; }
    while (s.bar_gen == gen) yield();
}
inline void warp_barrier() {
    Sched& s = sched();
    const int w = s.cur / 32;
    const unsigned gen = s.warp_gen[w];
    if (++s.warp_count[w] == s.warp_live[w]) { s.warp_count[w] = 0; ++s.warp_gen[w]; s.idle_rounds = 0; return; }
    while (s.warp_gen[w] == gen) yield();
}
inline void run_block(int nthreads, const std::function<void()>& body) {
    Sched& s = sched();
    if (nthreads > MAX_THREADS || nthreads % 32 != 0) { fprintf(stderr, "emu: unsupported block size %d\n", nthreads); abort(); }
    if (!s.stacks) s.stacks = static_cast<char*>(malloc((size_t)MAX_THREADS * STACK_BYTES));
    s.n = nthreads; s.live = nthreads; s.bar_count = 0; s.body = &body; s.idle_rounds = 0;
    for (int w = 0; w * 32 < nthreads; ++w) { s.warp_live[w] = 32; s.warp_count[w] = 0; }
    for (int t = 0; t < nthreads; ++t) {
        s.done[t] = false;
        getcontext(&s.ctx[t]);
        s.ctx[t].uc_stack.ss_sp = s.stacks + (size_t)t * STACK_BYTES;
        s.ctx[t].uc_stack.ss_size = STACK_BYTES;
        s.ctx[t].uc_link = &s.main_ctx;
        makecontext(&s.ctx[t], (void (*)())trampoline, 0);
    }
    while (s.live > 0) {
        for (int t = 0; t < nthreads; ++t) {
            if (s.done[t]) continue;
            s.cur = t;
            threadIdx.x = (unsigned)t; threadIdx.y = 0; threadIdx.z = 0;
            swapcontext(&s.main_ctx, &s.ctx[t]);
        }
        if (++s.idle_rounds > 1000000) { fprintf(stderr, "emu: deadlock (a barrier some live thread never reaches)\n"); abort(); }
    }
}
inline void run_grid(dim3 grid, dim3 block, const std::function<void()>& body) {
    if (block.y != 1 || block.z != 1) { fprintf(stderr, "emu: 1-D blocks only\n"); abort(); }
    gridDim = grid; blockDim = block;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                run_block((int)block.x, body);
            }
}
}  // namespace emu

// ---- device intrinsics --------------------------------------------------------------------------------
static inline void __syncthreads() { emu::block_barrier(); }
static inline void __threadfence() {}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "emu shuffle: <= 8 bytes");
    emu::Sched& s = emu::sched();
    const int me = s.cur;
    memcpy(s.shfl[me], &v, sizeof(T));
    emu::warp_barrier();
    T r;
    memcpy(&r, s.shfl[(me & ~31) | ((me & 31) ^ lane_mask)], sizeof(T));
    emu::warp_barrier();
    return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
#define __expf(x) expf(x)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }

// ---- splice host helpers (common.cuh / graph.h counterparts) ------------------------------------------
namespace splice {
inline char* emu_error_buf() { static char buf[1024] = ""; return buf; }
inline void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(emu_error_buf(), 1024, fmt, ap);
    va_end(ap);
}
inline const char* get_error() { return emu_error_buf(); }
inline long long& emu_launches() { static long long n = 0; return n; }
inline void count_launch(int n = 1) { emu_launches() += n; }
inline long long launch_count_now() { return emu_launches(); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
struct NvtxRange { explicit NvtxRange(const char*) {} };

#define SPLICE_CHECK_CUDA(expr)                                                              \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::splice::set_error("%s:%d CUDA error %d in %s", __FILE__, __LINE__, (int)_e, #expr); \
            return SPLICE_ERR_CUDA;                                                          \
        }                                                                                    \
    } while (0)
#define SPLICE_REQUIRE(cond, ...)                  \
    do {                                           \
        if (!(cond)) {                             \
            ::splice::set_error(__VA_ARGS__);      \
            return SPLICE_ERR_ARG;                 \
        }                                          \
    } while (0)
#define SPLICE_LAUNCH_CHECK() do { ::splice::count_launch(); } while (0)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args&&... args) {
    std::tuple<KArgs...> a{KArgs(args)...};
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x == 0) return 9;   // cudaErrorInvalidConfiguration
    if (grid.y > 65535 || grid.z > 65535) return 9;
    emu::run_grid(grid, block, [&]() { std::apply(kernel, a); });
    return cudaSuccess;
}
static inline void pdl_sync() {}
static inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct KeyHasher {
    uint64_t h = 0xcbf29ce484222325ull;
    KeyHasher& add(uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return *this; }
    KeyHasher& add(const void* p) { return add((uint64_t)reinterpret_cast<uintptr_t>(p)); }
};
class GraphCache {   // graphs are a replay optimisation: the emulation runs the body every time
public:
    void clear() {}
    int run(uint64_t, cudaStream_t stream, const std::function<int(cudaStream_t)>& body) { return body(stream); }
};
}  // namespace splice
