// splice_b200 — native conv generator (the reference's default-argument skip() U-Net), see generator.cu
#pragma once
#include <vector>

#include "common.cuh"
#include "graph.h"

namespace splice {

static constexpr int GEN_SCALES = 5;
static constexpr int GEN_PARAMS = 112;     // netG.parameters() order (oracle/splice_ref.py generator_param_keys)
static constexpr int GEN_BN = 30;          // BatchNorm2d layers, module order
static constexpr int GEN_SLOTS = 4;
static constexpr int BSTAT_STRIDE = 160;   // >= the widest BatchNorm (132 channels)
static constexpr int GEN_COUNTERS_PER_BN = 160;   // ticket counters per BatchNorm layer (one per channel / channel tile)

struct GenPointers {
    float* param[GEN_PARAMS];
    float* grad[GEN_PARAMS];               // accumulated into (+=); may be null when only forward is used
    float* running_mean[GEN_BN];
    float* running_var[GEN_BN];
    long long* num_batches_tracked[GEN_BN];
};

int gen_debug_conv(const float* x, int N, int Cin, int H, int W, const float* Wt, int Cout, int K, const float* bias, float* y,
                   int dgrad, int tiled, cudaStream_t st);

class GenEngine {
public:
    GenEngine();
    ~GenEngine();
    // x [N,3,H,W] -> out [N,3,H,W] (sigmoid). keep = activations stay in `slot` for backward().
    int forward(const GenPointers& p, const float* x, int N, int H, int W, float* out, int slot, bool keep,
                bool update_running, cudaStream_t stream);
    // dout [N,3,H,W] -> parameter gradients: p.grad += (accumulate) or p.grad = (overwrite, every element is written).
    // Calls on different slots may be in flight on different streams at once (each slot owns its scratch); they must
    // then be given disjoint gradient tables (overwrite mode + splice_accumulate afterwards).
    int backward(const GenPointers& p, const float* dout, int slot, bool accumulate, cudaStream_t stream);
    // applies the batch statistics the last forward() left in `slot` to the BatchNorm running buffers (momentum 0.1);
    // forward(update_running = true) does this itself, parallel-stream callers pass false and call this in call order
    int update_running_stats(const GenPointers& p, int slot, cudaStream_t stream);
    void set_graphs(bool on) { use_graphs_ = on; }

private:
    struct Conv { int cin, cout, k, stride, pw, pb; };         // pw/pb: parameter indices of weight / bias
    struct Bn { int c, pg, pb, idx; };                         // pg/pb: parameter indices of gamma / beta; idx: BN order
    struct Scale {
        Conv s, d1, d2, c1, c2;
        Bn bs, bd1, bd2, bcat, bc1, bc2;
        int cdeep;
    };
    struct ScaleBuf {
        int h, w, hd, wd;                                      // this scale's input size and the down-sampled size
        float *s_raw, *d1_raw, *d2_raw, *cat, *c1_raw, *c2_raw;
        float *dA_s, *dA_d1, *dA_d2, *dcat, *dA_c1, *dA_c2;
        float4 *k_s, *k_d1, *k_d2, *k_cat, *k_c1, *k_c2;       // per-channel (mean, invstd, a, b)
        float2 *m_s, *m_d1, *m_d2, *m_cat, *m_c1, *m_c2;       // per-channel (m1, m2) of the BN backward
    };
    struct Slot {
        int N = 0, H = 0, W = 0;
        bool valid = false;
        void* pool = nullptr;
        size_t pool_bytes = 0;
        const float* x = nullptr;
        float* x_copy = nullptr;
        float* out = nullptr;
        float* dfin = nullptr;
        float* dout_copy = nullptr;
        float2* bstat = nullptr;
        bool stats_pending = false;
        size_t stats_floats = 0;
        void* scratch = nullptr;   // partial-reduction scratch of this slot: [BN statistics | split-K sums | wgrad partials]
        size_t scratch_bytes = 0;
        int* counters = nullptr;   // ticket counters of the in-kernel BatchNorm finalisation (zero between launches)
        cudaStream_t side = nullptr;            // weight-gradient branch of the backward pass (forked from / joined into
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // the caller's stream; becomes a parallel branch under capture)
        ScaleBuf sb[GEN_SCALES];
    };
    int configure(Slot& s, int N, int H, int W);
    int ensure_scratch(Slot& s, size_t bytes);

    Scale sc_[GEN_SCALES];
    Conv final_;
    Slot slots_[GEN_SLOTS];
    int forward_body(const GenPointers& p, Slot& s, cudaStream_t st);
    int backward_body(const GenPointers& p, Slot& s, bool accumulate, cudaStream_t st);
    GraphCache graphs_;
    bool use_graphs_ = true;
};

}  // namespace splice
