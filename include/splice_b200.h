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
    int impl;                        /* 0 = tcgen05 persistent (product path), 1 = SIMT cross-check (tests only), 2 = tcgen05 one-tile-per-CTA */
    int bn_hint;                     /* 0 = auto; 64 / 128 / 256 */
} SpliceGemmArgs;
SPLICE_API int splice_gemm_bf16(const SpliceGemmArgs* args, void* stream);

/* ---- per-kernel entry points (parity tests call the same kernels the engine chains) --------------- */
/* y16 = LayerNorm(x) * gamma + beta, eps; stats [M,2] = (mean, rstd) or NULL.  ref: DINO Block.norm1/norm2 */
SPLICE_API int splice_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y16, void* stats, int M,
                                    int D, float eps, void* stream);
/* g_out = g_in + dLN(dy; x, stats, gamma); g16 = bf16(g_out). g_in may be NULL, g16 may be NULL */
SPLICE_API int splice_layernorm_bwd(const void* dy, const void* x, const void* stats, const void* gamma, const void* g_in,
                                    void* g_out, void* g16, int M, int D, void* stream);
/* o = softmax(q k^T / 8) v per (sequence, head); qkv bf16 [S*t, 3D]; o bf16 [S*t, D]; lse fp32 [S,H,t] (log2 domain).
 * tcgen05 / TMEM / TMA kernels (csrc/attention_tc.cu); SPLICE_B200_ATTN=legacy selects the mma.sync cross-check kernels.
 * ref: DINO Attention.forward behind models/extractor.py:83; the probs tap is extractor.py:44-45,57-61 */
SPLICE_API int splice_attention_fwd(const void* qkv, void* o, void* lse, int S, int t, int D, int H, void* stream);
SPLICE_API int splice_attention_bwd(const void* qkv, const void* o, const void* dout, const void* lse, void* delta_scratch,
                                    void* dqkv, int S, int t, int D, int H, void* stream);
/* torchvision Resize(size, max_size) output size.  ref: util/losses.py:20 */
SPLICE_API void splice_resized_hw(int h, int w, int size, int max_size, int* oh, int* ow);
/* img fp32 [3,h,w] -> antialiased resize (oh,ow) -> ImageNet normalise -> patch matrix bf16 rows row0.. (ld 3*p*p).
 * ref: LossG.global_transform util/losses.py:19-24 + DINO PatchEmbed unfold */
SPLICE_API int splice_preprocess_fwd(const void* img, int h, int w, int oh, int ow, int patch, void* patches, int row0,
                                     int normalize, void* stream);
/* out fp32 [3,oh,ow] = Normalize(Resize(img)) as a plain image.  ref: LossG.global_transform util/losses.py:19-24 */
SPLICE_API int splice_resize_normalize(const void* img, int h, int w, int oh, int ow, void* out, int normalize, void* stream);
SPLICE_API int splice_preprocess_bwd(const void* dpatch, int ldp, int row0, int h, int w, int oh, int ow, int patch,
                                     void* dimg, int normalize, void* stream);

/* ---- frozen DINO ViT engine ------------------------------------------------------------------------ */
/* ref: VitExtractor.__init__ models/extractor.py:19-29 (the hub model), and every `self.model(input_img)`.
 * `packed_weights` is ONE fp32 device buffer (the same buffer rank 0 broadcasts over NCCL), laid out as
 *   cls_token[D] pos_embed[n_pos*D] patch_embed.proj.weight[D*3*p*p] patch_embed.proj.bias[D]
 *   depth x { norm1.weight norm1.bias attn.qkv.weight[3D*D] attn.qkv.bias[3D] attn.proj.weight[D*D] attn.proj.bias
 *             norm2.weight norm2.bias mlp.fc1.weight[4D*D] mlp.fc1.bias[4D] mlp.fc2.weight[D*4D] mlp.fc2.bias }
 *   norm.weight[D] norm.bias[D]                                                                              */
typedef struct SpliceVitDesc {
    int patch, dim, heads, depth, n_pos;
    float ln_eps;
} SpliceVitDesc;
SPLICE_API size_t splice_vit_packed_floats(const SpliceVitDesc* desc);
SPLICE_API int splice_vit_create(void** ctx, const SpliceVitDesc* desc, const void* packed_weights, size_t n_floats,
                                 void* stream);
SPLICE_API int splice_vit_destroy(void* ctx);

typedef struct SpliceImage {
    void* data; /* fp32 [3,h,w], pixel range [0,1] (gradient: same shape) */
    int h, w;
} SpliceImage;

/* Batched forward of n_images images that share one ViT input size (out_h, out_w).
 * ref: get_feature_from_input / get_qkv_feature_from_input / get_keys_from_input, models/extractor.py:81-95,153-156 */
typedef struct SpliceVitForwardArgs {
    const SpliceImage* images; /* host array */
    int n_images;
    int out_h, out_w;      /* resize target = ViT input size; trailing out % patch pixels are dropped like the patch conv does */
    const void* pos;       /* fp32 [1 + gh*gw, D] interpolated pos_embed, NULL = trained grid */
    int n_grad;            /* activations of the first n_grad images are kept for splice_vit_backward */
    int slot;              /* 0..7: activation slot; passes in different slots may be alive (or in flight on different streams) at once */
    void* keys32;          /* out fp32 [n_images*t, D]: layer-11 keys (head h at columns h*64..), or NULL */
    void* cls32;           /* out fp32 [n_images, D]: block-11 output token 0 (pre final norm), or NULL */
    void* qkv32_all;       /* out fp32 [depth, n_images*t, 3D] (compat taps), or NULL */
    void* block32_all;     /* out fp32 [depth, n_images*t, D]  (compat taps), or NULL */
    int gemm_impl;         /* 0 = tcgen05 */
    int pre_normalized;    /* 1 = images are already ImageNet-normalised (VitExtractor API), skip (x-mean)/std */
    int use_graph;         /* 1 = the caller reuses the same output buffers every call: capture + replay a CUDA graph */
    int n_full;            /* > 0: only the first n_full images need the last block's OUTPUT (cls32 rows 0..n_full-1); the others
                            * are "keys-only" (ref: calculate_global_ssim_loss / calculate_global_id_loss, util/losses.py:74-83,96-105,
                            * read nothing past block 11's qkv) and stop after the last layer's qkv projection, forward and
                            * backward. 0 = every image runs the full depth, -1 = none does (all keys-only). Ignored when block32_all is requested. */
} SpliceVitForwardArgs;
SPLICE_API int splice_vit_forward(void* ctx, const SpliceVitForwardArgs* args, void* stream);

/* dgrad-only backward of the first n_grad images of the last forward in `slot`.
 * ref: loss_G.backward() train.py:78 restricted to the path util/losses.py builds */
typedef struct SpliceVitBackwardArgs {
    int slot;
    const void* dkeys32;       /* fp32 [n_grad*t, D] or NULL */
    const void* dcls32;        /* fp32 [n_grad, D] or NULL */
    const SpliceImage* grads;  /* host array, n_grad entries; data = fp32 [3,h,w] out (NULL entry = skip) */
    int gemm_impl;
    int use_graph;             /* 1 = dkeys32 / dcls32 are the same buffers every call: CUDA graph replay */
    /* gradients w.r.t. the all-layer taps of the forward (qkv32_all / block32_all), for callers that differentiate through
     * VitExtractor.get_feature_from_input / get_qkv_feature_from_input / get_keys_from_input at ANY layer (ref: inversion.py:33-39,
     * args.layer 0..11). Host arrays of `depth` device pointers, entries NULL where no gradient flows; the arrays may be NULL.
     * dblock32_layers[l]: fp32 [n_grad*t, D] = d loss / d (block l output); dqkv32_layers[l]: fp32 [n_grad*t, 3D]. Layers above the
     * highest one that receives a gradient are skipped. Not graph-replayed. */
    const void* const* dblock32_layers;
    const void* const* dqkv32_layers;
} SpliceVitBackwardArgs;
SPLICE_API int splice_vit_backward(void* ctx, const SpliceVitBackwardArgs* args, void* stream);

/* Per-kernel-class CUDA-event timing inside the engine (bench.py's roofline leg; off by default).
 * classes: 0 GEMM (tcgen05), 1 attention fwd, 2 attention bwd, 3 row-wise (LayerNorm), 4 preprocess */
typedef struct SpliceProfileEntry {
    long long count;   /* launches */
    double ms;         /* summed device time */
    double flops;      /* summed algorithmic FLOPs */
    double bytes;      /* summed algorithmic bytes */
} SpliceProfileEntry;
SPLICE_API int splice_vit_profile_enable(void* ctx, int on);
SPLICE_API int splice_vit_profile_read(void* ctx, SpliceProfileEntry* out, int n);

/* ---- losses ----------------------------------------------------------------------------------------- */
/* loss[0] = mean((S(keys_x) - S(keys_a))^2), S = cosine self-similarity of the rows of a [t,D] key matrix;
 * dkeys_x = coef * dloss/dkeys_x (fp32 [t,D]) or NULL.  ref: attn_cosine_sim models/extractor.py:4-9,
 * get_keys_self_sim_from_input :158-163, calculate_global_ssim_loss util/losses.py:74-83 */
SPLICE_API int splice_loss_ssim(void* ctx, const void* keys_x, const void* keys_a, int t, float coef, void* dkeys_x,
                                void* loss, int gemm_impl, void* stream);
/* loss[0] = mean((a - b)^2) over [rows, cols]; grad = coef * dloss/da or NULL.
 * ref: calculate_crop_cls_loss util/losses.py:85-94, calculate_global_id_loss :96-105 */
SPLICE_API int splice_loss_mse(void* ctx, const void* a, const void* b, int rows, int cols, float coef, void* grad,
                               void* loss, void* stream);
/* cosine self-similarity matrix fp32 [t,t] of keys fp32 [t,D].  ref: models/extractor.py:158-163 */
SPLICE_API int splice_keys_self_sim(void* ctx, const void* keys, int t, void* out_tt, int gemm_impl, void* stream);
/* total[0] = sum_i weights_host[i] * terms[i], n <= 8.  ref: LossG.forward util/losses.py:46-72 */
SPLICE_API int splice_weighted_total(const void* terms, const float* weights_host, int n, void* total, void* stream);
/* measurement aid: occupy `stream` for ~us microseconds (lets the host enqueue ahead of the device while bench.py
 * brackets individual kernels with events). No reference counterpart; not on the product path. */
SPLICE_API int splice_debug_spin(float us, void* stream);

/* ---- generator ------------------------------------------------------------------------------------- */
/* The default-argument skip() U-Net (ref: models/unet/skip.py:4-102, models/unet/common.py:11-124, called from
 * Model.forward models/model.py:12-25). Parameters are borrowed per call in netG.parameters() order (112 fp32
 * tensors), BatchNorm buffers in module order (30 layers). BatchNorm always uses batch statistics (the reference
 * never calls .eval()); running statistics are updated like nn.BatchNorm2d(momentum=0.1) when update_running != 0. */
#define SPLICE_GEN_PARAMS 112
#define SPLICE_GEN_BN 30
typedef struct SpliceGenPointers {
    void* param[SPLICE_GEN_PARAMS];
    void* grad[SPLICE_GEN_PARAMS];            /* fp32, ACCUMULATED into by splice_gen_backward; may be NULL for forward */
    void* running_mean[SPLICE_GEN_BN];
    void* running_var[SPLICE_GEN_BN];
    void* num_batches_tracked[SPLICE_GEN_BN]; /* int64 scalars */
} SpliceGenPointers;
SPLICE_API int splice_gen_create(void** ctx);
SPLICE_API int splice_gen_destroy(void* ctx);
/* out[N,3,H,W] = netG(x[N,3,H,W]); keep != 0 retains the activations in `slot` (0..3) for splice_gen_backward */
SPLICE_API int splice_gen_forward(void* ctx, const SpliceGenPointers* p, const void* x, int N, int H, int W, void* out,
                                  int slot, int keep, int update_running, void* stream);
/* Applies the batch statistics the last splice_gen_forward(update_running = 0) left in `slot` to the BatchNorm running
 * buffers (nn.BatchNorm2d momentum 0.1, unbiased variance, num_batches_tracked += 1). Callers that issue several forward
 * passes on parallel streams use this afterwards, in call order, to reproduce the reference's sequential updates. */
SPLICE_API int splice_gen_update_running(void* ctx, const SpliceGenPointers* p, int slot, void* stream);
/* 1 (default) = replay CUDA graphs keyed by (slot, shape, pointer table); 0 = always launch eagerly */
SPLICE_API int splice_gen_set_graphs(void* ctx, int on);
/* parameter gradients of the forward kept in `slot`, given dout[N,3,H,W] = d loss / d out (ref: loss.backward(), train.py:78).
 * accumulate != 0: p->grad += (autograd's accumulation over the 2-3 netG calls of a step); accumulate == 0: p->grad = (every
 * element is written). Calls on different slots may run concurrently on different streams when given disjoint grad tables. */
SPLICE_API int splice_gen_backward(void* ctx, const SpliceGenPointers* p, const void* dout, int slot, int accumulate, void* stream);
/* test hook, not on the product path: ONE stride-1 "same" convolution of the generator (K = 1 or 3, weights [Cout,Cin,K,K] as
 * nn.Conv2d stores them) or its data gradient, on the shared-memory-tiled kernel family (tiled != 0) or the direct one, without
 * producer BatchNorm / statistics. dgrad == 0: y[N,Cout,H,W] = conv(x[N,Cin,H,W]) + bias; dgrad != 0: x is d y [N,Cout,H,W] and
 * y receives d x [N,Cin,H,W]. ref: nn.Conv2d built by models/unet/common.py:99-124 */
SPLICE_API int splice_gen_debug_conv(const void* x, int N, int Cin, int H, int W, const void* w, int Cout, int K, const void* bias,
                                     void* y, int dgrad, int tiled, void* stream);
/* dst[i] += srcs[0][i] + ... + srcs[n_src-1][i] (fp32, fixed order, n_src <= 4): folds the per-call gradient buffers of
 * concurrently executed netG backward passes into .grad (ref: autograd gradient accumulation, train.py:56,78) */
SPLICE_API int splice_accumulate(void* dst, const void* const* srcs, int n_src, size_t n, void* stream);

/* ---- generator, other skip() configurations ------------------------------------------------------------ */
/* skip() with arguments other than the defaults (ref: models/unet/skip.py:4-102; the one the reference builds is inversion.py:21-25:
 * 6 scales, 32 input channels, 7x7 / 5x5 / 3x3 filters, pad = 'reflection' -> nn.ReflectionPad2d, models/unet/common.py:113-118).
 * The configuration is data; strided down-sampling, bilinear x2 up-sampling, LeakyReLU(0.2), BatchNorm2d in training mode,
 * need1x1_up and need_bias are fixed. Filter sizes 1, 3, 5 or 7; at most 160 channels per tensor. */
#define SPLICE_GENX_MAX_SCALES 8
typedef struct SpliceGenXConfig {
    int n_scales;                              /* len(num_channels_down) */
    int in_channels, out_channels;             /* num_input_channels, num_output_channels */
    int ch_down[SPLICE_GENX_MAX_SCALES];       /* num_channels_down */
    int ch_up[SPLICE_GENX_MAX_SCALES];         /* num_channels_up */
    int ch_skip[SPLICE_GENX_MAX_SCALES];       /* num_channels_skip (all > 0) */
    int k_down[SPLICE_GENX_MAX_SCALES];        /* filter_size_down */
    int k_up[SPLICE_GENX_MAX_SCALES];          /* filter_size_up */
    int k_skip;                                /* filter_skip_size */
    int reflect;                               /* pad: 0 = 'zero', 1 = 'reflection' */
    int sigmoid;                               /* need_sigmoid */
} SpliceGenXConfig;
SPLICE_API int splice_genx_create(const SpliceGenXConfig* cfg, void** ctx);
SPLICE_API int splice_genx_destroy(void* ctx);
/* number of parameter tensors (22 per scale + 2, netG.parameters() order) and of BatchNorm2d layers (6 per scale, module order) */
SPLICE_API int splice_genx_counts(void* ctx, int* n_params, int* n_bn);
/* Pointer tables (host arrays of device pointers, copied): fp32 parameters; fp32 gradients (NULL table = forward only); BatchNorm
 * running_mean / running_var (fp32) and num_batches_tracked (int64) (NULL tables = never update running statistics). */
SPLICE_API int splice_genx_bind(void* ctx, void* const* params, void* const* grads, void* const* running_mean,
                                void* const* running_var, void* const* num_batches_tracked);
/* out[N,out_channels,H,W] = net(x[N,in_channels,H,W]) (ref: net(net_input), inversion.py:65); keep != 0 retains the activations for
 * splice_genx_backward (one pass at a time; a later forward replaces it) */
SPLICE_API int splice_genx_forward(void* ctx, const void* x, int N, int H, int W, void* out, int keep, int update_running, void* stream);
/* parameter gradients of the kept forward given dout = d loss / d out (ref: loss.backward(), inversion.py:68); accumulate != 0: += */
SPLICE_API int splice_genx_backward(void* ctx, const void* dout, int accumulate, void* stream);
SPLICE_API int splice_genx_set_graphs(void* ctx, int on);

/* ---- optimiser -------------------------------------------------------------------------------------- */
/* One Adam step over n_tensors fp32 tensors (host arrays of device pointers / element counts).
 * `step` is the 1-based step count AFTER the increment.  ref: get_optimizer util/util.py:28-32 -> torch.optim.Adam */
SPLICE_API int splice_adam_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                                const int* numel, int n_tensors, int step, float lr, float beta1, float beta2, float eps,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPLICE_B200_H_ */
