// splice_b200 — native generator for skip() configurations other than the optimisation loop's default: see generator_x.cu
#pragma once
#ifdef SPLICE_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "graph.h"
#endif

namespace splice {

static constexpr int GENX_MAX_SCALES = 8;
static constexpr int GENX_MAX_PARAMS = 22 * GENX_MAX_SCALES + 2;   // netG.parameters(): 22 tensors per scale + the final conv
static constexpr int GENX_MAX_BN = 6 * GENX_MAX_SCALES;            // BatchNorm2d layers, module order
static constexpr int GENX_MAX_CH = 160;                            // widest tensor (ticket counters / statistics rows per layer)

struct GenXConfig {                  // the arguments of skip() that change the arithmetic (ref models/unet/skip.py:4-12)
    int n_scales;
    int in_channels, out_channels;
    int ch_down[GENX_MAX_SCALES], ch_up[GENX_MAX_SCALES], ch_skip[GENX_MAX_SCALES];
    int k_down[GENX_MAX_SCALES], k_up[GENX_MAX_SCALES];
    int k_skip;
    int reflect;                     // pad: 0 = 'zero', 1 = 'reflection'
    int sigmoid;                     // need_sigmoid
};

class GenXEngine {
public:
    static int create(const GenXConfig& cfg, GenXEngine** out);
    ~GenXEngine();
    int n_params() const { return 22 * ns_ + 2; }
    int n_bn() const { return 6 * ns_; }
    // tables in netG.parameters() / module order; grads may be null until backward() is used, the BatchNorm buffers until
    // forward(update_running = true) is used. The tables are copied.
    int bind(float* const* params, float* const* grads, float* const* running_mean, float* const* running_var,
             long long* const* num_batches_tracked);
    // x [N,Cin,H,W] -> out [N,Cout,H,W]. keep = the activations stay for backward() (one pass at a time).
    int forward(const float* x, int N, int H, int W, float* out, bool keep, bool update_running, cudaStream_t stream);
    // dout [N,Cout,H,W] -> parameter gradients (+= when accumulate, = otherwise: every element is written)
    int backward(const float* dout, bool accumulate, cudaStream_t stream);
    void set_graphs(bool on) { use_graphs_ = on; }

private:
    GenXEngine() = default;
    struct Conv { int cin, cout, k, stride, pw, pb; };         // pw/pb: parameter indices of weight / bias
    struct Bn { int c, pg, pb, idx; };                         // pg/pb: parameter indices of gamma / beta; idx: BN order
    struct Scale {
        Conv s, d1, d2, c1, c2;
        Bn bs, bd1, bd2, bcat, bc1, bc2;
        int cskip, cdeep;
    };
    struct ScaleBuf {
        int h, w, hd, wd;                                      // this scale's input size and the down-sampled size
        float *s_raw, *d1_raw, *d2_raw, *cat, *c1_raw, *c2_raw;
        float *dA_s, *dA_d1, *dA_d2, *dcat, *dA_c1, *dA_c2;
        float4 *k_s, *k_d1, *k_d2, *k_cat, *k_c1, *k_c2;       // per-channel (mean, invstd, a, b)
        float2 *m_s, *m_d1, *m_d2, *m_cat, *m_c1, *m_c2;       // per-channel (m1, m2) of the BN backward
    };
    int configure(int N, int H, int W);
    int forward_body(cudaStream_t st);
    int backward_body(bool accumulate, cudaStream_t st);
    int update_running(cudaStream_t st);

    GenXConfig cfg_{};
    int ns_ = 0, max_c_ = 0, max_cskip_ = 0;
    Scale sc_[GENX_MAX_SCALES];
    Conv final_{};
    float* param_[GENX_MAX_PARAMS] = {};
    float* grad_[GENX_MAX_PARAMS] = {};
    float* rmean_[GENX_MAX_BN] = {};
    float* rvar_[GENX_MAX_BN] = {};
    long long* nbt_[GENX_MAX_BN] = {};
    bool bound_ = false, have_grads_ = false, have_running_ = false;

    int N_ = 0, H_ = 0, W_ = 0;
    bool valid_ = false;                    // a kept forward pass is waiting for backward()
    void* pool_ = nullptr;
    size_t pool_bytes_ = 0;
    float *x_copy_ = nullptr, *out_ = nullptr, *dfin_ = nullptr, *dout_copy_ = nullptr;
    float2* bstat_ = nullptr;
    void* scratch_ = nullptr;               // [statistics partials | skip-branch statistics partials | wgrad partials | padded dgrad]
    size_t scratch_bytes_ = 0;
    size_t stats_floats_ = 0, skip_floats_ = 0, dpad_floats_ = 0;
    int* counters_ = nullptr;
    cudaStream_t side_ = nullptr;
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
    ScaleBuf sb_[GENX_MAX_SCALES];
    GraphCache graphs_;
    bool use_graphs_ = true;
};

}  // namespace splice
