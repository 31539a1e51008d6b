// splice_b200 — frozen DINO ViT engine.
//
// One object per process/GPU holds the ViT weights (bf16, each matrix also pre-transposed for the
// dgrad-only backward — the weights never change, train.py:43 only optimises netG) and the workspaces.
// forward(): batches ALL images of a step that share a token grid into one [S*t, D] problem (the reference
// runs 6 batch-1 forwards per step, 2 of them duplicates: util/losses.py:74-105), exports the two taps the
// loss needs (layer-11 keys, extractor.py:153-156; block-11 output token 0, losses.py:90) and keeps the
// activations of the first n_grad sequences. backward(): dgrad only, from d(keys) / d(cls) to d(image) —
// the reference's autograd also computes 133.6 GFLOP of unused ViT weight gradients per pass.
//
// Follows DINO VisionTransformer.forward as restated in oracle/dino_vit.py (third-party, not in the tree).
#include "vit.h"

#include "attention.h"
#include "elementwise.h"
#include "gemm.h"
#include "preprocess.h"

namespace splice {

// -------------------------------------------------------------------------------------------------
// weight packing: fp32 [rows, cols] -> bf16 same layout + bf16 transposed [cols, rows]
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convert_weight_kernel(const float* __restrict__ w, int rows, int cols,
                                                             bf16* __restrict__ w16, bf16* __restrict__ wT16) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int rr = r0 + r, cc = c0 + tx;
        const float v = (rr < rows && cc < cols) ? w[(size_t)rr * cols + cc] : 0.f;
        tile[r][tx] = v;
        if (rr < rows && cc < cols) w16[(size_t)rr * cols + cc] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int cc = c0 + r, rr = r0 + tx;
        if (rr < rows && cc < cols) wT16[(size_t)cc * rows + rr] = __float2bfloat16(tile[tx][r]);
    }
}

// LayerNorm folded into the GEMM that consumes it: y = LN(x) W^T + b = rstd * (x (W gamma)^T - mean * colsum) + (b + W beta).
// w16 / wT16 = bf16(W[n,k] * gamma[k]) and its transpose (the dgrad through LN's output then needs no gamma).
__global__ void __launch_bounds__(256) fold_weight_kernel(const float* __restrict__ w, const float* __restrict__ gamma, int rows,
                                                          int cols, bf16* __restrict__ w16, bf16* __restrict__ wT16) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int rr = r0 + r, cc = c0 + tx;
        const float v = (rr < rows && cc < cols) ? w[(size_t)rr * cols + cc] * gamma[cc] : 0.f;
        tile[r][tx] = v;
        if (rr < rows && cc < cols) w16[(size_t)rr * cols + cc] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int cc = c0 + r, rr = r0 + tx;
        if (rr < rows && cc < cols) wT16[(size_t)cc * rows + rr] = __float2bfloat16(tile[tx][r]);
    }
}
// one warp per output row n: colsum[n] = sum_k float(bf16(W[n,k] gamma[k])) (the matrix the tensor cores see),
// bias2[n] = bias[n] + sum_k W[n,k] beta[k]
__global__ void __launch_bounds__(128) fold_rowsum_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ bias, int rows,
                                                          int cols, float* __restrict__ colsum, float* __restrict__ bias2) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= rows) return;
    float cs = 0.f, bs = 0.f;
    for (int k = lane; k < cols; k += 32) {
        const float wv = w[(size_t)n * cols + k];
        cs += __bfloat162float(__float2bfloat16(wv * gamma[k]));
        bs += wv * beta[k];
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) { colsum[n] = cs; bias2[n] = bias[n] + bs; }
}

__global__ void gather_cls_kernel(const float* __restrict__ x, float* __restrict__ cls, int t, int D) {
    pdl_sync();
    const int s = blockIdx.x;
    for (int c = threadIdx.x; c < D; c += blockDim.x) cls[(size_t)s * D + c] = x[(size_t)s * t * D + c];
}

__global__ void scatter_cls_grad_kernel(float* __restrict__ g, bf16* __restrict__ g16, const float* __restrict__ dcls, int t,
                                        int D) {
    pdl_sync();
    const int s = blockIdx.x;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float v = dcls[(size_t)s * D + c];
        g[(size_t)s * t * D + c] = v;
        g16[(size_t)s * t * D + c] = __float2bfloat16(v);
    }
}

size_t VitEngine::packed_size(const VitDesc& d) {
    const size_t D = d.dim, pp3 = 3 * (size_t)d.patch * d.patch;
    size_t n = D + (size_t)d.n_pos * D + D * pp3 + D;
    n += (size_t)d.depth * (2 * D + 3 * D * D + 3 * D + D * D + D + 2 * D + 4 * D * D + 4 * D + 4 * D * D + D);
    n += 2 * D;
    return n;
}

int VitEngine::create(VitEngine** out, const VitDesc& d, const float* packed_dev, size_t n_floats, cudaStream_t stream) {
    SPLICE_REQUIRE(d.dim % 128 == 0 && d.dim == d.heads * 64, "vit: unsupported dims D=%d H=%d (head dim must be 64)", d.dim,
                   d.heads);
    SPLICE_REQUIRE((3 * d.patch * d.patch) % 64 == 0, "vit: 3*patch^2 = %d must be a multiple of 64", 3 * d.patch * d.patch);
    SPLICE_REQUIRE(n_floats == packed_size(d), "vit: packed weight buffer has %zu floats, expected %zu", n_floats, packed_size(d));
    VitEngine* e = new VitEngine();
    e->d_ = d;
    const size_t D = d.dim, pp3 = 3 * (size_t)d.patch * d.patch;
    SPLICE_CHECK_CUDA(cudaMalloc(&e->w32_, n_floats * sizeof(float)));
    SPLICE_CHECK_CUDA(cudaMemcpyAsync(e->w32_, packed_dev, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    // every matrix twice (W and W^T) in bf16
    const size_t n_mat = D * pp3 + (size_t)d.depth * (3 * D * D + D * D + 8 * D * D);
    SPLICE_CHECK_CUDA(cudaMalloc(&e->w16_, 2 * n_mat * sizeof(bf16)));
    {
        // measured (B200, config 2): 177.0 it/s fused against 184.1 unfused on the same box - the 60 LayerNorm launches it
        // removes (1.15 -> 0.64 ms of row-wise kernel time per step) cost less than the heavier GEMM epilogues it needs
        // (statistics partials, column sums, per-row merge: 6.42 -> 7.11 ms of GEMM time). Off unless asked for.
        const char* v = getenv("SPLICE_B200_LN_FUSED");
        e->ln_fused_ = (v && v[0] == '1');
    }
    if (e->ln_fused_) {   // the folded copies (~200 MB for ViT-B) exist only when the option is on
        SPLICE_CHECK_CUDA(cudaMalloc(&e->wf16_, (size_t)d.depth * 2 * (3 * D * D + 4 * D * D) * sizeof(bf16)));
        SPLICE_CHECK_CUDA(cudaMalloc(&e->wf32_, (size_t)d.depth * 14 * D * sizeof(float)));
    }
    bf16* qf = e->wf16_;
    float* pf = e->wf32_;
    const float* p = e->w32_;
    bf16* q = e->w16_;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    int rc = SPLICE_OK;
    // the LayerNorm (gamma, beta) in front of a Linear(cols -> rows) whose fp32 weight / bias are w / b: folded copies
    auto fold = [&](const float* w, const float* b, const float* g, const float* bt, int rows, int cols, const bf16** wf,
                    const bf16** wfT, const float** cs, const float** bf) {
        *wf = nullptr; *wfT = nullptr; *cs = nullptr; *bf = nullptr;
        if (!e->ln_fused_) return;
        bf16* a = qf; qf += (size_t)rows * cols;
        bf16* t2 = qf; qf += (size_t)rows * cols;
        float* c1 = pf; pf += rows;
        float* b2 = pf; pf += rows;
        dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
        fold_weight_kernel<<<grid, 256, 0, stream>>>(w, g, rows, cols, a, t2);
        count_launch();
        fold_rowsum_kernel<<<ceil_div(rows, 4), 128, 0, stream>>>(w, g, bt, b, rows, cols, c1, b2);
        count_launch();
        *wf = a; *wfT = t2; *cs = c1; *bf = b2;
    };
    auto mat = [&](int rows, int cols, const bf16** w, const bf16** wT) {
        const float* src = take((size_t)rows * cols);
        bf16* a = q; q += (size_t)rows * cols;
        bf16* b = q; q += (size_t)rows * cols;
        dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
        convert_weight_kernel<<<grid, 256, 0, stream>>>(src, rows, cols, a, b);
        count_launch();
        *w = a; *wT = b;
    };
    e->cls_ = take(D);
    e->pos_ = take((size_t)d.n_pos * D);
    mat((int)D, (int)pp3, &e->pe_w_, &e->pe_wT_);
    e->pe_b_ = take(D);
    e->L_.resize(d.depth);
    for (int l = 0; l < d.depth; ++l) {
        LayerW& L = e->L_[l];
        L.ln1_g = take(D); L.ln1_b = take(D);
        const float* qkv_w32 = p;
        mat(3 * (int)D, (int)D, &L.qkv_w, &L.qkv_wT); L.qkv_b = take(3 * D);
        fold(qkv_w32, L.qkv_b, L.ln1_g, L.ln1_b, 3 * (int)D, (int)D, &L.qkv_wf, &L.qkv_wfT, &L.qkv_cs, &L.qkv_bf);
        mat((int)D, (int)D, &L.proj_w, &L.proj_wT); L.proj_b = take(D);
        L.ln2_g = take(D); L.ln2_b = take(D);
        const float* fc1_w32 = p;
        mat(4 * (int)D, (int)D, &L.fc1_w, &L.fc1_wT); L.fc1_b = take(4 * D);
        fold(fc1_w32, L.fc1_b, L.ln2_g, L.ln2_b, 4 * (int)D, (int)D, &L.fc1_wf, &L.fc1_wfT, &L.fc1_cs, &L.fc1_bf);
        mat((int)D, 4 * (int)D, &L.fc2_w, &L.fc2_wT); L.fc2_b = take(D);
    }
    e->norm_g_ = take(D);
    e->norm_b_ = take(D);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        set_error("vit: weight conversion launch failed: %s", cudaGetErrorString(ce));
        delete e;
        return SPLICE_ERR_CUDA;
    }
    (void)rc;
    *out = e;
    return SPLICE_OK;
}

VitEngine::~VitEngine() {
    cudaFree(w32_);
    cudaFree(w16_);
    cudaFree(wf16_);
    cudaFree(wf32_);
    for (auto& s : slots_) cudaFree(s.pool);
    cudaFree(loss_ws_);
}

int VitEngine::loss_scratch(int t, void** ptr, size_t* bytes) {
    // generous: 4 Gram operands [tp, 3D] bf16, khat^T [D, tp] bf16, 2 x S [tp, tp] fp32, E [tp, tp] bf16, R [tp, D] fp32, vectors
    const size_t tp = (size_t)((t + 63) / 64) * 64, D = d_.dim;
    const size_t need = 4 * tp * 3 * D * 2 + D * tp * 2 + 2 * tp * tp * 4 + tp * tp * 2 + tp * D * 4 + 16 * tp * 4 + 4096;
    if (need > loss_ws_bytes_) {
        SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(loss_ws_);
        loss_ws_ = nullptr;
        SPLICE_CHECK_CUDA(cudaMalloc(&loss_ws_, need));
        SPLICE_CHECK_CUDA(cudaMemset(loss_ws_, 0, need));
        loss_ws_bytes_ = need;
    }
    *ptr = loss_ws_;
    *bytes = loss_ws_bytes_;
    return SPLICE_OK;
}

int VitEngine::configure(Slot& s, int S, int t, int n_grad) {
    if (s.S == S && s.t == t && s.n_grad == n_grad && s.pool) return SPLICE_OK;
    const size_t D = d_.dim, pp3 = 3 * (size_t)d_.patch * d_.patch, M = (size_t)S * t, Mg = (size_t)n_grad * t;
    const int depth = d_.depth, H = d_.heads;
    size_t off = 0;
    auto bump = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    // pass 1: sizes
    struct Item { size_t off; };
    std::vector<size_t> offs;
    auto plan = [&](size_t bytes) { offs.push_back(bump(bytes)); };
    plan((size_t)S * (t - 1) * pp3 * 2);                          // patches
    for (int l = 0; l <= depth; ++l) plan(M * D * 4);             // x0
    for (int l = 0; l < depth; ++l) plan(M * D * 4);              // x1
    for (int l = 0; l < depth; ++l) plan(M * 3 * D * 2);          // qkv
    for (int l = 0; l < depth; ++l) plan(M * D * 2);              // o
    for (int l = 0; l < depth; ++l) plan(M * 4 * D * 2);          // hpre
    for (int l = 0; l < depth; ++l) plan((size_t)S * H * t * 4);  // lse
    for (int l = 0; l < depth; ++l) plan(M * 2 * 4);              // st1
    for (int l = 0; l < depth; ++l) plan(M * 2 * 4);              // st2
    plan(M * D * 2);                                              // a16
    plan(M * D * 2);                                              // xa16
    plan(M * D * 2);                                              // xb16
    plan(M * (D / 32) * 8);                                       // sp_a
    plan(M * (D / 32) * 8);                                       // sp_b
    plan(M * 4 * D * 2);                                          // h16
    plan(Mg * D * 4);                                             // g
    plan(Mg * D * 4);                                             // da
    plan((size_t)(n_grad > 0 ? n_grad : 1) * H * t * 4);          // delta
    plan(Mg * pp3 * 4);                                           // dpatch
    plan(Mg * D * 2);                                             // g16
    plan(Mg * 4 * D * 2);                                         // dh16
    plan(Mg * D * 2);                                             // do16
    plan(Mg * 3 * D * 2);                                         // dqkv16
    if (off > s.pool_bytes) {
        SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(s.pool);
        s.pool = nullptr;
        s.pool_bytes = 0;
        SPLICE_CHECK_CUDA(cudaMalloc(&s.pool, off));
        s.pool_bytes = off;
    }
    uint8_t* base = static_cast<uint8_t*>(s.pool);
    size_t k = 0;
    auto nextp = [&]() { return base + offs[k++]; };
    s.patches = (bf16*)nextp();
    s.x0.resize(depth + 1); s.x1.resize(depth); s.qkv.resize(depth); s.o.resize(depth); s.hpre.resize(depth);
    s.lse.resize(depth); s.st1.resize(depth); s.st2.resize(depth);
    for (int l = 0; l <= depth; ++l) s.x0[l] = (float*)nextp();
    for (int l = 0; l < depth; ++l) s.x1[l] = (float*)nextp();
    for (int l = 0; l < depth; ++l) s.qkv[l] = (bf16*)nextp();
    for (int l = 0; l < depth; ++l) s.o[l] = (bf16*)nextp();
    for (int l = 0; l < depth; ++l) s.hpre[l] = (bf16*)nextp();
    for (int l = 0; l < depth; ++l) s.lse[l] = (float*)nextp();
    for (int l = 0; l < depth; ++l) s.st1[l] = (float*)nextp();
    for (int l = 0; l < depth; ++l) s.st2[l] = (float*)nextp();
    s.a16 = (bf16*)nextp();
    s.xa16 = (bf16*)nextp();
    s.xb16 = (bf16*)nextp();
    s.sp_a = (float2*)nextp();
    s.sp_b = (float2*)nextp();
    s.h16 = (bf16*)nextp();
    s.g = (float*)nextp();
    s.da = (float*)nextp();
    s.delta = (float*)nextp();
    s.dpatch = (float*)nextp();
    s.g16 = (bf16*)nextp();
    s.dh16 = (bf16*)nextp();
    s.do16 = (bf16*)nextp();
    s.dqkv16 = (bf16*)nextp();
    s.S = S; s.t = t; s.n_grad = n_grad;
    return SPLICE_OK;
}

void VitEngine::profile_enable(bool on) { prof_on_ = on; }
void VitEngine::prof_begin(int cat, double flops, double bytes, cudaStream_t st) {
    if (!prof_on_) return;
    ProfRec r;
    r.cat = cat; r.flops = flops; r.bytes = bytes;
    auto get = [&]() {
        cudaEvent_t e;
        if (!prof_pool_.empty()) { e = prof_pool_.back(); prof_pool_.pop_back(); } else cudaEventCreate(&e);
        return e;
    };
    r.e0 = get(); r.e1 = get();
    cudaEventRecord(r.e0, st);
    prof_pending_.push_back(r);
}
void VitEngine::prof_end(cudaStream_t st) {
    if (!prof_on_ || prof_pending_.empty()) return;
    cudaEventRecord(prof_pending_.back().e1, st);
}
int VitEngine::profile_read(ProfTotals* out, int n) {
    SPLICE_CHECK_CUDA(cudaDeviceSynchronize());
    for (auto& r : prof_pending_) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            ProfTotals& t = prof_tot_[r.cat];
            t.count += 1; t.ms += ms; t.flops += r.flops; t.bytes += r.bytes;
        }
        prof_pool_.push_back(r.e0); prof_pool_.push_back(r.e1);
    }
    prof_pending_.clear();
    for (int i = 0; i < n && i < PROF_NCAT; ++i) out[i] = prof_tot_[i];
    for (int i = 0; i < PROF_NCAT; ++i) prof_tot_[i] = ProfTotals{0, 0.0, 0.0, 0.0};
    return SPLICE_OK;
}

// profiled launch wrappers
#define PROF(cat, flops, bytes, expr)                   \
    do {                                                \
        prof_begin((cat), (flops), (bytes), stream);    \
        int _rc = (expr);                               \
        prof_end(stream);                               \
        if (_rc) return _rc;                            \
    } while (0)
#define GEMM_FLOPS(M, N, K) (2.0 * (double)(M) * (double)(N) * (double)(K))

#define RC(expr)                 \
    do {                         \
        int _rc = (expr);        \
        if (_rc) return _rc;     \
    } while (0)

int VitEngine::forward(const VitForwardArgs& a, cudaStream_t stream) {
    NvtxRange nvtx("splice_vit_forward");
    SPLICE_REQUIRE(a.images && a.n_images > 0, "vit_forward: no images");
    SPLICE_REQUIRE(a.slot >= 0 && a.slot < VIT_SLOTS, "vit_forward: slot must be in [0,%d)", VIT_SLOTS);
    SPLICE_REQUIRE(a.n_grad >= 0 && a.n_grad <= a.n_images, "vit_forward: n_grad out of range");
    const int p = d_.patch, D = d_.dim, H = d_.heads, depth = d_.depth, pp3 = 3 * p * p;
    SPLICE_REQUIRE(a.out_h >= p && a.out_w >= p, "vit_forward: ViT input %dx%d is smaller than one %d-pixel patch", a.out_h,
                   a.out_w, p);
    const int gh = a.out_h / p, gw = a.out_w / p, t = 1 + gh * gw, S = a.n_images, M = S * t;
    SPLICE_REQUIRE(a.pos || (t == d_.n_pos && gh == gw),
                   "vit_forward: token grid %dx%d differs from the trained grid; pass an interpolated pos_embed", gh, gw);
    Slot& s = slots_[a.slot];
    RC(configure(s, S, t, a.n_grad));
    s.gh = gh; s.gw = gw; s.oh = a.out_h; s.ow = a.out_w; s.pre_normalized = a.pre_normalized;
    SPLICE_REQUIRE(a.n_full >= -1 && a.n_full <= S, "vit_forward: n_full out of range");
    s.n_full = a.block32_all ? S : (a.n_full > 0 ? a.n_full : (a.n_full < 0 ? 0 : S));
    s.imgs.assign(a.images, a.images + S);
    const float* pos = a.pos ? a.pos : pos_;

    for (int i = 0; i < S; ++i)
        RC(preprocess_fwd(a.images[i].data, a.images[i].h, a.images[i].w, a.out_h, a.out_w, p, s.patches, i * (t - 1),
                          !a.pre_normalized, stream));
    if (prof_on_ || !a.use_graph) return forward_body(a, s, S, t, pos, stream);
    KeyHasher k;
    k.add((uint64_t)1).add((uint64_t)a.slot).add((uint64_t)S).add((uint64_t)t).add((uint64_t)a.n_grad).add((uint64_t)s.n_full).add(s.pool).add(pos)
        .add(a.keys32).add(a.cls32).add(a.qkv32_all).add(a.block32_all).add((uint64_t)a.gemm_impl);
    return graphs_.run(k.h, stream, [&](cudaStream_t st) { return forward_body(a, s, S, t, pos, st); });
}

// everything of the forward pass that depends only on (slot, S, t, output pointers): graph-captured
int VitEngine::forward_body(const VitForwardArgs& a, Slot& s, int S, int t, const float* pos, cudaStream_t stream) {
    const int p = d_.patch, D = d_.dim, H = d_.heads, depth = d_.depth, pp3 = 3 * p * p;
    const int M = S * t;
    (void)p;
    // LayerNorm folded into the GEMMs (ln_fused_): every producer of a residual-stream row (cls rows, patch embed, proj, fc2)
    // also writes its bf16 copy and per-32-column statistics partials; the qkv / fc1 GEMMs take the raw bf16 rows as A, the
    // gamma-folded weights as B and normalise in the epilogue. No LayerNorm launch on the forward chain.
    const bool lnf = ln_fused_;
    const int nparts = D / 32;
    RC(write_cls_rows(s.x0[0], cls_, pos, S, t, D, stream, lnf ? s.xa16 : nullptr, lnf ? s.sp_a : nullptr));
    {
        GemmEpilogue ep;
            ep.b_const = 1;
        ep.c32 = s.x0[0]; ep.ldc32 = D; ep.bias = pe_b_;
        ep.rows_per_seq = t - 1; ep.pos = pos; ep.ldpos = D;
        if (lnf) { ep.c16 = s.xa16; ep.ldc16 = D; ep.stat_part = s.sp_a; ep.stat_nparts = nparts; }
        PROF(PROF_GEMM, GEMM_FLOPS(S * (t - 1), D, pp3), 0.0, gemm_bf16_tn(s.patches, pp3, pe_w_, pp3, S * (t - 1), D, pp3, ep, a.gemm_impl, 0, stream));
    }
    for (int l = 0; l < depth; ++l) {
        const LayerW& L = L_[l];
        if (!lnf) PROF(PROF_ROWWISE, 0.0, 6.0 * M * D, layernorm_fwd(s.x0[l], L.ln1_g, L.ln1_b, s.a16, s.st1[l], M, D, d_.ln_eps, stream));
        {
            GemmEpilogue ep;
            ep.b_const = 1;
            ep.c16 = s.qkv[l]; ep.ldc16 = 3 * D; ep.bias = lnf ? L.qkv_bf : L.qkv_b;
            if (lnf) {
                ep.ln_part = s.sp_a; ep.ln_nparts = nparts; ep.ln_colsum = L.qkv_cs; ep.ln_eps = d_.ln_eps;
                ep.ln_stat_out = reinterpret_cast<float2*>(s.st1[l]);
            }
            if (a.qkv32_all) { ep.c32 = a.qkv32_all + (size_t)l * M * 3 * D; ep.ldc32 = 3 * D; }
            if (l == depth - 1 && a.keys32) { ep.slice32 = a.keys32; ep.slice_c0 = D; ep.slice_c1 = 2 * D; ep.ldslice = D; }
            PROF(PROF_GEMM, GEMM_FLOPS(M, 3 * D, D), 0.0,
                 gemm_bf16_tn(lnf ? s.xa16 : s.a16, D, lnf ? L.qkv_wf : L.qkv_w, D, M, 3 * D, D, ep, a.gemm_impl, 0, stream));
        }
        // the rest of the LAST layer only serves the block output ([CLS] row): keys-only sequences (the trailing S - n_full) stop here
        const int Sf = (l == depth - 1) ? s.n_full : S, Mf = Sf * t;
        if (Sf == 0) continue;
        PROF(PROF_ATTN_FWD, 4.0 * Sf * (double)t * t * D, 0.0, attention_fwd(s.qkv[l], s.o[l], s.lse[l], Sf, t, D, H, stream));
        {
            GemmEpilogue ep;
            ep.b_const = 1;
            ep.c32 = s.x1[l]; ep.ldc32 = D; ep.bias = L.proj_b; ep.residual = s.x0[l]; ep.ldr = D;
            if (lnf) { ep.c16 = s.xb16; ep.ldc16 = D; ep.stat_part = s.sp_b; ep.stat_nparts = nparts; }
            PROF(PROF_GEMM, GEMM_FLOPS(Mf, D, D), 0.0, gemm_bf16_tn(s.o[l], D, L.proj_w, D, Mf, D, D, ep, a.gemm_impl, 0, stream));
        }
        if (!lnf) PROF(PROF_ROWWISE, 0.0, 6.0 * Mf * D, layernorm_fwd(s.x1[l], L.ln2_g, L.ln2_b, s.a16, s.st2[l], Mf, D, d_.ln_eps, stream));
        {
            GemmEpilogue ep;
            ep.b_const = 1;
            ep.c16 = s.h16; ep.ldc16 = 4 * D; ep.bias = lnf ? L.fc1_bf : L.fc1_b; ep.act = GEMM_ACT_GELU; ep.aux16 = s.hpre[l]; ep.ldaux = 4 * D;
            if (lnf) {
                ep.ln_part = s.sp_b; ep.ln_nparts = nparts; ep.ln_colsum = L.fc1_cs; ep.ln_eps = d_.ln_eps;
                ep.ln_stat_out = reinterpret_cast<float2*>(s.st2[l]);
            }
            PROF(PROF_GEMM, GEMM_FLOPS(Mf, 4 * D, D), 0.0,
                 gemm_bf16_tn(lnf ? s.xb16 : s.a16, D, lnf ? L.fc1_wf : L.fc1_w, D, Mf, 4 * D, D, ep, a.gemm_impl, 0, stream));
        }
        {
            GemmEpilogue ep;
            ep.b_const = 1;
            ep.c32 = s.x0[l + 1]; ep.ldc32 = D; ep.bias = L.fc2_b; ep.residual = s.x1[l]; ep.ldr = D;
            if (lnf && l + 1 < depth) { ep.c16 = s.xa16; ep.ldc16 = D; ep.stat_part = s.sp_a; ep.stat_nparts = nparts; }
            PROF(PROF_GEMM, GEMM_FLOPS(Mf, D, 4 * D), 0.0, gemm_bf16_tn(s.h16, 4 * D, L.fc2_w, 4 * D, Mf, D, 4 * D, ep, a.gemm_impl, 0, stream));
        }
        if (a.block32_all)
            SPLICE_CHECK_CUDA(cudaMemcpyAsync(a.block32_all + (size_t)l * M * D, s.x0[l + 1], (size_t)M * D * sizeof(float),
                                              cudaMemcpyDeviceToDevice, stream));
    }
    if (a.cls32 && s.n_full > 0) {
        SPLICE_CHECK_CUDA(launch_pdl(gather_cls_kernel, dim3(s.n_full), dim3(256), 0, stream, (const float*)s.x0[depth], a.cls32, t, D));
        SPLICE_LAUNCH_CHECK();
    }
    return SPLICE_OK;
}

int VitEngine::backward(const VitBackwardArgs& a, cudaStream_t stream) {
    NvtxRange nvtx("splice_vit_backward");
    SPLICE_REQUIRE(a.slot >= 0 && a.slot < VIT_SLOTS, "vit_backward: slot must be in [0,%d)", VIT_SLOTS);
    Slot& s = slots_[a.slot];
    SPLICE_REQUIRE(s.pool && s.n_grad > 0, "vit_backward: slot %d holds no forward pass with n_grad > 0", a.slot);
    SPLICE_REQUIRE(a.grads, "vit_backward: no gradient outputs");
    const int p = d_.patch, D = d_.dim, H = d_.heads, depth = d_.depth, pp3 = 3 * p * p;
    const int t = s.t, Sg = s.n_grad, Mg = Sg * t;

    if (prof_on_ || !a.use_graph) {
        RC(backward_body(a, s, stream));
    } else {
        KeyHasher k;
        k.add((uint64_t)2).add((uint64_t)a.slot).add((uint64_t)s.S).add((uint64_t)Sg).add((uint64_t)s.n_full).add((uint64_t)t).add(s.pool).add(a.dkeys32).add(a.dcls32)
            .add((uint64_t)a.gemm_impl);
        RC(graphs_.run(k.h, stream, [&](cudaStream_t st) { return backward_body(a, s, st); }));
    }
    for (int i = 0; i < Sg; ++i) {
        SPLICE_REQUIRE(a.grads[i].h == s.imgs[i].h && a.grads[i].w == s.imgs[i].w,
                       "vit_backward: gradient %d is %dx%d but the forward image was %dx%d", i, a.grads[i].h, a.grads[i].w,
                       s.imgs[i].h, s.imgs[i].w);
        if (!a.grads[i].data) continue;
        RC(preprocess_bwd(s.dpatch, pp3, i * t + 1, s.imgs[i].h, s.imgs[i].w, s.oh, s.ow, p, a.grads[i].data,
                          !s.pre_normalized, stream));
    }
    return SPLICE_OK;
}


// g += src (fp32 residual-stream gradient), g16 = bf16(g): a tap gradient entering the stream between two blocks
__global__ void __launch_bounds__(256) add_stream_grad_kernel(float* __restrict__ g, bf16* __restrict__ g16, const float* __restrict__ src, size_t n) {
    pdl_sync();
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = g[i] + src[i];
        g[i] = v;
        g16[i] = __float2bfloat16(v);
    }
}

int VitEngine::backward_body(const VitBackwardArgs& a, Slot& s, cudaStream_t stream) {
    const int p = d_.patch, D = d_.dim, H = d_.heads, depth = d_.depth, pp3 = 3 * p * p;
    const int t = s.t, Sg = s.n_grad, Mg = Sg * t;
    (void)p;
    SPLICE_CHECK_CUDA(cudaMemsetAsync(s.g, 0, (size_t)Mg * D * sizeof(float), stream));
    SPLICE_CHECK_CUDA(cudaMemsetAsync(s.g16, 0, (size_t)Mg * D * sizeof(bf16), stream));
    bool have_g = false;
    if (a.dcls32) {
        SPLICE_CHECK_CUDA(launch_pdl(scatter_cls_grad_kernel, dim3(Sg), dim3(256), 0, stream, s.g, s.g16, (const float*)a.dcls32, t, D));
        SPLICE_LAUNCH_CHECK();
        have_g = true;
    }
    for (int l = depth - 1; l >= 0; --l) {
        const LayerW& L = L_[l];
        const float* dblock = a.dblock32_layers ? a.dblock32_layers[l] : nullptr;
        const float* dqkv = a.dqkv32_layers ? a.dqkv32_layers[l] : nullptr;
        if (dblock) {   // gradient w.r.t. block l's output (an all-layer tap): joins the residual-stream gradient here
            const size_t n = (size_t)Mg * D;
            const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
            SPLICE_CHECK_CUDA(launch_pdl(add_stream_grad_kernel, dim3(blocks), dim3(256), 0, stream, s.g, s.g16, dblock, n));
            SPLICE_LAUNCH_CHECK();
            have_g = true;
        }
        // last layer: only the first n_full sequences ran past the qkv projection (the others are keys-only)
        const int Sb = (l == depth - 1 && s.n_full < Sg) ? s.n_full : Sg, Mb = Sb * t;
        if (have_g && Sb > 0) {
            {   // d(gelu out) = g W2 ; d(pre) = . * gelu'(pre)
                GemmEpilogue ep;
            ep.b_const = 1;
                ep.c16 = s.dh16; ep.ldc16 = 4 * D; ep.act = GEMM_ACT_GELU_GRAD; ep.aux16 = s.hpre[l]; ep.ldaux = 4 * D;
                PROF(PROF_GEMM, GEMM_FLOPS(Mb, 4 * D, D), 0.0, gemm_bf16_tn(s.g16, D, L.fc2_wT, D, Mb, 4 * D, D, ep, a.gemm_impl, 0, stream));
            }
            {   // d(LN2 out) = d(pre) W1
                GemmEpilogue ep;
            ep.b_const = 1;
                ep.c32 = s.da; ep.ldc32 = D;
                PROF(PROF_GEMM, GEMM_FLOPS(Mb, D, 4 * D), 0.0, gemm_bf16_tn(s.dh16, 4 * D, ln_fused_ ? L.fc1_wfT : L.fc1_wT, 4 * D, Mb, D, 4 * D, ep, a.gemm_impl, 0, stream));
            }
            PROF(PROF_ROWWISE, 0.0, 18.0 * Mb * D, layernorm_bwd(s.da, s.x1[l], s.st2[l], ln_fused_ ? nullptr : L.ln2_g, s.g, s.g, s.g16, Mb, D, stream));
            {   // d(attn out) = g Wproj
                GemmEpilogue ep;
            ep.b_const = 1;
                ep.c16 = s.do16; ep.ldc16 = D;
                PROF(PROF_GEMM, GEMM_FLOPS(Mb, D, D), 0.0, gemm_bf16_tn(s.g16, D, L.proj_wT, D, Mb, D, D, ep, a.gemm_impl, 0, stream));
            }
            PROF(PROF_ATTN_BWD, 8.0 * Sb * (double)t * t * D, 0.0,
                 attention_bwd(s.qkv[l], s.o[l], s.do16, s.lse[l], s.delta, s.dqkv16, Sb, t, D, H, stream));
            if (Mb < Mg)   // keys-only sequences: d(qkv) starts at zero
                SPLICE_CHECK_CUDA(cudaMemsetAsync(s.dqkv16 + (size_t)Mb * 3 * D, 0, (size_t)(Mg - Mb) * 3 * D * sizeof(bf16), stream));
        } else {
            // nothing flows back from the block output (keys-only objective): d(qkv) starts at zero
            SPLICE_CHECK_CUDA(cudaMemsetAsync(s.dqkv16, 0, (size_t)Mg * 3 * D * sizeof(bf16), stream));
        }
        if (l == depth - 1 && a.dkeys32) RC(add_f32_into_bf16_cols(s.dqkv16, 3 * D, D, a.dkeys32, D, Mg, D, stream));
        if (dqkv) RC(add_f32_into_bf16_cols(s.dqkv16, 3 * D, 0, dqkv, 3 * D, Mg, 3 * D, stream));
        if (!have_g && !(l == depth - 1 && a.dkeys32) && !dqkv) continue;  // still all-zero
        {   // d(LN1 out) = d(qkv) Wqkv
            GemmEpilogue ep;
            ep.b_const = 1;
            ep.c32 = s.da; ep.ldc32 = D;
            PROF(PROF_GEMM, GEMM_FLOPS(Mg, D, 3 * D), 0.0, gemm_bf16_tn(s.dqkv16, 3 * D, ln_fused_ ? L.qkv_wfT : L.qkv_wT, 3 * D, Mg, D, 3 * D, ep, a.gemm_impl, 0, stream));
        }
        PROF(PROF_ROWWISE, 0.0, 18.0 * Mg * D, layernorm_bwd(s.da, s.x0[l], s.st1[l], ln_fused_ ? nullptr : L.ln1_g, s.g, s.g, s.g16, Mg, D, stream));
        have_g = true;
    }
    {   // d(patch pixels) = g Wpe  (cls rows produce rows that the adjoint resampler never reads)
        GemmEpilogue ep;
            ep.b_const = 1;
        ep.c32 = s.dpatch; ep.ldc32 = pp3;
        PROF(PROF_GEMM, GEMM_FLOPS(Mg, pp3, D), 0.0, gemm_bf16_tn(s.g16, D, pe_wT_, D, Mg, pp3, D, ep, a.gemm_impl, 0, stream));
    }
    return SPLICE_OK;
}

}  // namespace splice
