// splice_b200 — frozen DINO ViT engine (forward with feature taps + dgrad-only backward), see vit.cu
#pragma once
#include <vector>

#include "common.cuh"
#include "graph.h"

namespace splice {

struct VitDesc {
    int patch = 8, dim = 768, heads = 12, depth = 12, n_pos = 785;
    float ln_eps = 1e-6f;
};

static constexpr int VIT_SLOTS = 8;   // activation slots: forward passes that may be alive (or in flight on parallel streams) at once

struct ImageRef { const float* data; int h, w; };
struct ImageGradRef { float* data; int h, w; };

struct VitForwardArgs {
    const ImageRef* images = nullptr;
    int n_images = 0;
    int out_h = 0, out_w = 0;
    const float* pos = nullptr;       // [1 + gh*gw, D] or nullptr = native pos_embed
    int n_grad = 0;
    int n_full = 0;                   // > 0: images beyond the first n_full stop after the last layer's qkv projection
    int slot = 0;
    float* keys32 = nullptr;          // [n_images*t, D]
    float* cls32 = nullptr;           // [n_images, D]
    float* qkv32_all = nullptr;       // [depth, n_images*t, 3D]   (compat taps)
    float* block32_all = nullptr;     // [depth, n_images*t, D]    (compat taps)
    int gemm_impl = 0;                // 0 = tcgen05, 1 = SIMT cross-check
    bool pre_normalized = false;
    bool use_graph = false;
};

struct VitBackwardArgs {
    int slot = 0;
    const float* dkeys32 = nullptr;   // [n_grad*t, D]
    const float* dcls32 = nullptr;    // [n_grad, D]
    const ImageGradRef* grads = nullptr;  // n_grad entries
    int gemm_impl = 0;
    bool use_graph = false;
    const float* const* dblock32_layers = nullptr;   // [depth] -> [n_grad*t, D] or null
    const float* const* dqkv32_layers = nullptr;     // [depth] -> [n_grad*t, 3D] or null
};

enum ProfCat : int { PROF_GEMM = 0, PROF_ATTN_FWD = 1, PROF_ATTN_BWD = 2, PROF_ROWWISE = 3, PROF_PREPROC = 4, PROF_NCAT = 5 };
struct ProfTotals { long long count; double ms; double flops; double bytes; };

class VitEngine {
public:
    // per-kernel-class CUDA-event timing (off by default; bench.py turns it on for a separate measurement leg)
    void profile_enable(bool on);
    int profile_read(ProfTotals* out, int n);   // synchronises, accumulates and clears the pending events
    void prof_begin(int cat, double flops, double bytes, cudaStream_t st);
    void prof_end(cudaStream_t st);
    static int create(VitEngine** out, const VitDesc& d, const float* packed_dev, size_t n_floats, cudaStream_t stream);
    ~VitEngine();
    static size_t packed_size(const VitDesc& d);
    int forward(const VitForwardArgs& a, cudaStream_t stream);
    int backward(const VitBackwardArgs& a, cudaStream_t stream);
    const VitDesc& desc() const { return d_; }
    int slot_n_grad(int slot) const { return slots_[slot].pool ? slots_[slot].n_grad : 0; }
    // scratch for the loss kernels, sized for sequences of up to t tokens (grown on demand)
    int loss_scratch(int t, void** ptr, size_t* bytes);

private:
    struct LayerW {
        const float *ln1_g, *ln1_b, *qkv_b, *proj_b, *ln2_g, *ln2_b, *fc1_b, *fc2_b;
        const bf16 *qkv_w, *qkv_wT, *proj_w, *proj_wT, *fc1_w, *fc1_wT, *fc2_w, *fc2_wT;
        // LayerNorm folded into the consuming GEMM: W * gamma (and its transpose), column sums of the bf16 matrix, b + W beta
        const bf16 *qkv_wf, *qkv_wfT, *fc1_wf, *fc1_wfT;
        const float *qkv_cs, *qkv_bf, *fc1_cs, *fc1_bf;
    };
    struct Slot {
        int S = 0, t = 0, gh = 0, gw = 0, n_grad = 0, oh = 0, ow = 0, n_full = 0;
        bool pos_custom = false, pre_normalized = false;
        std::vector<ImageRef> imgs;
        void* pool = nullptr;
        size_t pool_bytes = 0;
        // carved from pool
        bf16* patches = nullptr;             // [S*(t-1), 3pp]
        std::vector<float*> x0, x1;          // x0[depth+1], x1[depth]   [M, D]
        std::vector<bf16*> qkv, o, hpre;     // per layer
        std::vector<float*> lse, st1, st2;   // per layer
        bf16 *a16 = nullptr, *h16 = nullptr;
        bf16 *xa16 = nullptr, *xb16 = nullptr;        // bf16 copies of the raw residual stream (x0 / x1 of the current layer)
        float2 *sp_a = nullptr, *sp_b = nullptr;      // their per-32-column (sum, M2) partials [M][D/32]
        // backward
        float *g = nullptr, *da = nullptr, *delta = nullptr, *dpatch = nullptr;
        bf16 *g16 = nullptr, *dh16 = nullptr, *do16 = nullptr, *dqkv16 = nullptr;
    };
    int configure(Slot& s, int S, int t, int n_grad);

    VitDesc d_;
    float* w32_ = nullptr;
    bf16* w16_ = nullptr;
    bf16* wf16_ = nullptr;      // folded matrices
    float* wf32_ = nullptr;     // their column sums and folded biases
    bool ln_fused_ = false;     // SPLICE_B200_LN_FUSED=1: LayerNorm folded into the qkv / fc1 GEMMs (measured slower: vit.cu create())
    const float *cls_ = nullptr, *pos_ = nullptr, *pe_b_ = nullptr, *norm_g_ = nullptr, *norm_b_ = nullptr;
    const bf16 *pe_w_ = nullptr, *pe_wT_ = nullptr;
    std::vector<LayerW> L_;
    Slot slots_[VIT_SLOTS];
    void* loss_ws_ = nullptr;
    size_t loss_ws_bytes_ = 0;
    struct ProfRec { int cat; double flops, bytes; cudaEvent_t e0, e1; };
    bool prof_on_ = false;
    std::vector<ProfRec> prof_pending_;
    std::vector<cudaEvent_t> prof_pool_;
    ProfTotals prof_tot_[PROF_NCAT] = {};
    GraphCache graphs_;
    int forward_body(const VitForwardArgs& a, Slot& s, int S, int t, const float* pos, cudaStream_t stream);
    int backward_body(const VitBackwardArgs& a, Slot& s, cudaStream_t stream);
};

}  // namespace splice
