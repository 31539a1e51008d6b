/* splice_b200 — C-ABI of the B200-native hot path of omerbt/Splice.
 *
 * The reference has no native code and no FFI: its hot path is Python calling torch / torchvision / the
 * hub DINO ViT. This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md): plain pointers, sizes and a cudaStream_t, no torch types.  Each entry point cites the
 * reference call site it replaces as `ref: file:line` (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return value 0 = ok, negative = error (SPLICE_ERR_*), message via splice_last_error();
 *   - there is no CPU fallback: unsupported shapes/configs are errors.
 */
#ifndef SPLICE_B200_H_
#define SPLICE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPLICE_API __attribute__((visibility("default")))
#else
#define SPLICE_API
#endif

#define SPLICE_B200_VERSION 100

/* ---- library ------------------------------------------------------------------------------------ */
SPLICE_API int splice_version(void);
/* last error message of the calling thread ("" if none) */
SPLICE_API const char* splice_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench.py: gpu_launches) */
SPLICE_API long long splice_launch_count(void);
SPLICE_API void splice_launch_count_reset(void);

/* ---- dense contraction -------------------------------------------------------------------------- */
/* C[M,N] = epilogue(A[M,K] · B[N,K]^T), bf16 operands (row-major, K contiguous), fp32 accumulate on tcgen05.
 * ref: the nn.Linear / patch-embed Conv2d calls inside `self.model(input_img)`, models/extractor.py:83,91,99.
 * Epilogue (all optional, applied in this order): +bias[N]; act (0 none, 1 GELU(erf) with the pre-activation
 * saved to aux16, 2 multiply by GELU'(aux16)); patch->token row remap + pos_embed add (rows_per_seq > 0);
 * +residual (fp32); store fp32 (c32) and/or bf16 (c16); fp32 export of columns [slice_c0, slice_c1). */
typedef struct SpliceGemmArgs {
    const void* A; int lda;          /* bf16 [M, lda]  */
    const void* B; int ldb;          /* bf16 [N, ldb]  */
    int M, N, K;
    void* c32; int ldc32;            /* fp32 out or NULL */
    void* c16; int ldc16;            /* bf16 out or NULL */
    const void* bias;                /* fp32 [N] or NULL */
    const void* residual; int ldr;   /* fp32 or NULL (may alias c32) */
    int act;
    void* aux16; int ldaux;          /* bf16 pre-activation (out for act=1, in for act=2) */
    int rows_per_seq;                /* 0 = no token remap */
    const void* pos; int ldpos;      /* fp32 pos_embed [1+rows_per_seq, ldpos] */
    void* slice32; int slice_c0, slice_c1, ldslice;
    int impl;                        /* 0 = tcgen05 (product path), 1 = SIMT cross-check (tests only) */
    int bn_hint;                     /* 0 = auto; 64 / 128 / 256 */
} SpliceGemmArgs;
SPLICE_API int splice_gemm_bf16(const SpliceGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPLICE_B200_H_ */
