// splice_b200 — extern "C" entry points (include/splice_b200.h). Thin argument marshalling only; the
// kernels live in the other translation units.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "gemm.h"
#include "splice_b200.h"

namespace splice {

static thread_local char g_err[1024] = {0};
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace splice

using namespace splice;

extern "C" {

SPLICE_API int splice_version(void) { return SPLICE_B200_VERSION; }
SPLICE_API const char* splice_last_error(void) { return get_error(); }
SPLICE_API long long splice_launch_count(void) { return g_launches.load(); }
SPLICE_API void splice_launch_count_reset(void) { g_launches.store(0); }

SPLICE_API int splice_gemm_bf16(const SpliceGemmArgs* a, void* stream) {
    SPLICE_REQUIRE(a != nullptr, "splice_gemm_bf16: null args");
    GemmEpilogue ep;
    ep.c32 = static_cast<float*>(a->c32); ep.ldc32 = a->ldc32;
    ep.c16 = static_cast<bf16*>(a->c16); ep.ldc16 = a->ldc16;
    ep.bias = static_cast<const float*>(a->bias);
    ep.residual = static_cast<const float*>(a->residual); ep.ldr = a->ldr;
    ep.act = a->act;
    ep.aux16 = static_cast<bf16*>(a->aux16); ep.ldaux = a->ldaux;
    ep.rows_per_seq = a->rows_per_seq;
    ep.pos = static_cast<const float*>(a->pos); ep.ldpos = a->ldpos;
    ep.slice32 = static_cast<float*>(a->slice32);
    ep.slice_c0 = a->slice_c0; ep.slice_c1 = a->slice_c1; ep.ldslice = a->ldslice;
    return gemm_bf16_tn(static_cast<const bf16*>(a->A), a->lda, static_cast<const bf16*>(a->B), a->ldb, a->M, a->N, a->K,
                        ep, a->impl, a->bn_hint, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
