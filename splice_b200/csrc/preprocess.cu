// splice_b200 — fused resize + normalise + patchify (forward) and its adjoint (backward).
//
// Replaces LossG.global_transform (util/losses.py:19-24: Resize(dino_global_patch_size, max_size=480) then
// Normalize(ImageNet mean/std)), the unfold half of DINO's patch-embed Conv2d(k = stride = patch), and
// their autograd backward. The resize is torchvision's tensor path = ATen `_upsample_bilinear2d_aa`
// (align_corners=False): a separable triangle filter whose support widens by the scale factor when
// down-sampling and degenerates to plain bilinear when up-sampling.
// Per-axis tap tables (start index + weights, and their transposes for the adjoint) are built on the host
// once per (in, out) size and cached on the device; both kernels are pure gathers (no atomics), HBM-bound.
#include <map>
#include <mutex>
#include <vector>

#include "preprocess.h"

namespace splice {

void resized_hw(int h, int w, int size, int max_size, int* oh, int* ow) {
    const int shrt = (w <= h) ? w : h, lng = (w <= h) ? h : w;
    int new_short = size, new_long = (int)((long long)size * lng / shrt);
    if (max_size > 0 && new_long > max_size) {
        new_short = (int)((long long)max_size * new_short / new_long);
        new_long = max_size;
    }
    if (w <= h) { *ow = new_short; *oh = new_long; } else { *ow = new_long; *oh = new_short; }
}

struct AxisTable {
    int n_out = 0, taps = 0;   // weights [n_out, taps], start [n_out]
    int* start = nullptr;      // device
    float* w = nullptr;        // device
};

// forward table: out index i gathers in[start[i] + u] * w[i][u]; transposed: in index a gathers out[...]
static void build_tables_host(int n_in, int n_out, std::vector<int>& st, std::vector<float>& wt, int& taps,
                              std::vector<int>& st_t, std::vector<float>& wt_t, int& taps_t) {
    const double scale = (double)n_in / (double)n_out;
    const double support = scale >= 1.0 ? scale : 1.0;
    const double inv = scale >= 1.0 ? 1.0 / scale : 1.0;
    std::vector<std::vector<float>> rows(n_out);
    st.assign(n_out, 0);
    taps = 1;
    if (n_in == n_out) {
        for (int i = 0; i < n_out; ++i) { st[i] = i; rows[i] = {1.f}; }
    } else {
        for (int i = 0; i < n_out; ++i) {
            const double center = scale * (i + 0.5);
            int lo = (int)(center - support + 0.5); if (lo < 0) lo = 0;
            int hi = (int)(center + support + 0.5); if (hi > n_in) hi = n_in;
            std::vector<double> ww(hi - lo);
            double tot = 0.0;
            for (int j = lo; j < hi; ++j) {
                double x = ((double)j - center + 0.5) * inv; if (x < 0) x = -x;
                const double v = x < 1.0 ? 1.0 - x : 0.0;
                ww[j - lo] = v; tot += v;
            }
            st[i] = lo;
            rows[i].resize(hi - lo);
            for (int j = 0; j < hi - lo; ++j) rows[i][j] = (float)(ww[j] / tot);
            if (hi - lo > taps) taps = hi - lo;
        }
    }
    wt.assign((size_t)n_out * taps, 0.f);
    for (int i = 0; i < n_out; ++i)
        for (size_t j = 0; j < rows[i].size(); ++j) wt[(size_t)i * taps + j] = rows[i][j];
    // transpose: for input a, the contiguous range of outputs that touch it
    std::vector<int> first(n_in, n_out), last(n_in, -1);
    for (int i = 0; i < n_out; ++i)
        for (size_t j = 0; j < rows[i].size(); ++j) {
            const int a = st[i] + (int)j;
            if (i < first[a]) first[a] = i;
            if (i > last[a]) last[a] = i;
        }
    taps_t = 1;
    for (int a = 0; a < n_in; ++a)
        if (last[a] >= first[a] && last[a] - first[a] + 1 > taps_t) taps_t = last[a] - first[a] + 1;
    st_t.assign(n_in, 0);
    wt_t.assign((size_t)n_in * taps_t, 0.f);
    for (int a = 0; a < n_in; ++a) {
        if (last[a] < first[a]) continue;  // input pixel never sampled (cannot happen for this filter, kept for safety)
        st_t[a] = first[a];
        for (int i = first[a]; i <= last[a]; ++i) {
            const int j = a - st[i];
            if (j >= 0 && j < (int)rows[i].size()) wt_t[(size_t)a * taps_t + (i - first[a])] = rows[i][j];
        }
    }
}

struct TablePair { AxisTable fwd, bwd; };
static std::mutex g_tab_mu;
static std::map<std::pair<int, int>, TablePair> g_tabs;

static int upload(const std::vector<int>& st, const std::vector<float>& w, int n, int taps, AxisTable* t) {
    t->n_out = n; t->taps = taps;
    SPLICE_CHECK_CUDA(cudaMalloc(&t->start, st.size() * sizeof(int)));
    SPLICE_CHECK_CUDA(cudaMalloc(&t->w, w.size() * sizeof(float)));
    SPLICE_CHECK_CUDA(cudaMemcpy(t->start, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    SPLICE_CHECK_CUDA(cudaMemcpy(t->w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    return SPLICE_OK;
}

static int get_tables(int n_in, int n_out, TablePair* out) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_pair(n_in, n_out);
    auto it = g_tabs.find(key);
    if (it == g_tabs.end()) {
        std::vector<int> st, st_t; std::vector<float> w, w_t; int taps, taps_t;
        build_tables_host(n_in, n_out, st, w, taps, st_t, w_t, taps_t);
        TablePair tp;
        int rc = upload(st, w, n_out, taps, &tp.fwd); if (rc) return rc;
        rc = upload(st_t, w_t, n_in, taps_t, &tp.bwd); if (rc) return rc;
        it = g_tabs.emplace(key, tp).first;
    }
    *out = it->second;
    return SPLICE_OK;
}

__constant__ float c_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_istd[3] = {1.f / 0.229f, 1.f / 0.224f, 1.f / 0.225f};

__global__ void __launch_bounds__(256) preprocess_fwd_kernel(const float* __restrict__ img, int h, int w, int oh, int ow,
                                                             const int* __restrict__ ys, const float* __restrict__ wy, int ty,
                                                             const int* __restrict__ xs, const float* __restrict__ wx, int tx,
                                                             int patch, bf16* __restrict__ patches, int row0, bool normalize) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int c = blockIdx.z;
    // the patch conv (k = stride = patch, no padding) ignores the trailing oh % patch rows / ow % patch columns
    if (i >= (oh / patch) * patch || j >= (ow / patch) * patch) return;
    const float* src = img + (size_t)c * h * w;
    const int y0 = ys[i], x0 = xs[j];
    float acc = 0.f;
    for (int u = 0; u < ty; ++u) {
        const float a = wy[i * ty + u];
        if (a == 0.f) continue;
        const float* r = src + (size_t)(y0 + u) * w + x0;
        float racc = 0.f;
        for (int v = 0; v < tx; ++v) {
            const float b = wx[j * tx + v];
            if (b != 0.f) racc = fmaf(b, r[v], racc);
        }
        acc = fmaf(a, racc, acc);
    }
    const float val = normalize ? (acc - c_mean[c]) * c_istd[c] : acc;
    const int gw = ow / patch;
    const int row = row0 + (i / patch) * gw + (j / patch);
    const int col = c * patch * patch + (i % patch) * patch + (j % patch);
    patches[(size_t)row * (3 * patch * patch) + col] = __float2bfloat16(val);
}

// plain image output [3, oh, ow] (the standalone LossG.global_transform of the reference API)
__global__ void __launch_bounds__(256) resize_normalize_kernel(const float* __restrict__ img, int h, int w, int oh, int ow,
                                                               const int* __restrict__ ys, const float* __restrict__ wy, int ty,
                                                               const int* __restrict__ xs, const float* __restrict__ wx, int tx,
                                                               float* __restrict__ out, bool normalize) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int c = blockIdx.z;
    if (i >= oh || j >= ow) return;
    const float* src = img + (size_t)c * h * w;
    const int y0 = ys[i], x0 = xs[j];
    float acc = 0.f;
    for (int u = 0; u < ty; ++u) {
        const float a = wy[i * ty + u];
        if (a == 0.f) continue;
        const float* r = src + (size_t)(y0 + u) * w + x0;
        float racc = 0.f;
        for (int v = 0; v < tx; ++v) {
            const float b = wx[j * tx + v];
            if (b != 0.f) racc = fmaf(b, r[v], racc);
        }
        acc = fmaf(a, racc, acc);
    }
    out[((size_t)c * oh + i) * ow + j] = normalize ? (acc - c_mean[c]) * c_istd[c] : acc;
}

__global__ void __launch_bounds__(256) preprocess_bwd_kernel(const float* __restrict__ dpatch, int ldp, int row0, int h, int w,
                                                             int oh, int ow, const int* __restrict__ ys, const float* __restrict__ wy,
                                                             int ty, const int* __restrict__ xs, const float* __restrict__ wx, int tx,
                                                             int patch, float* __restrict__ dimg, bool normalize) {
    const int b = blockIdx.x * 32 + (threadIdx.x & 31);
    const int a = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int c = blockIdx.z;
    if (a >= h || b >= w) return;
    const int i0 = ys[a], j0 = xs[b];
    const int gw = ow / patch, pp = patch * patch;
    const int ch = (oh / patch) * patch, cw = gw * patch;  // region covered by whole patches
    float acc = 0.f;
    for (int u = 0; u < ty; ++u) {
        const float wa = wy[a * ty + u];
        const int i = i0 + u;
        if (wa == 0.f || i >= ch) continue;
        float racc = 0.f;
        for (int v = 0; v < tx; ++v) {
            const float wb = wx[b * tx + v];
            const int j = j0 + v;
            if (wb == 0.f || j >= cw) continue;
            const int row = row0 + (i / patch) * gw + (j / patch);
            const int col = c * pp + (i % patch) * patch + (j % patch);
            racc = fmaf(wb, dpatch[(size_t)row * ldp + col], racc);
        }
        acc = fmaf(wa, racc, acc);
    }
    dimg[((size_t)c * h + a) * w + b] = normalize ? acc * c_istd[c] : acc;
}

int preprocess_fwd(const float* img, int h, int w, int oh, int ow, int patch, bf16* patches, int row0, bool normalize,
                   cudaStream_t stream) {
    SPLICE_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0, "preprocess: empty image");
    SPLICE_REQUIRE(oh >= patch && ow >= patch, "preprocess: %dx%d is smaller than one %d-pixel patch", oh, ow, patch);
    TablePair ty, tx;
    int rc = get_tables(h, oh, &ty); if (rc) return rc;
    rc = get_tables(w, ow, &tx); if (rc) return rc;
    dim3 grid(ceil_div(ow, 32), ceil_div(oh, 8), 3);
    preprocess_fwd_kernel<<<grid, 256, 0, stream>>>(img, h, w, oh, ow, ty.fwd.start, ty.fwd.w, ty.fwd.taps, tx.fwd.start, tx.fwd.w,
                                                    tx.fwd.taps, patch, patches, row0, normalize);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int resize_normalize(const float* img, int h, int w, int oh, int ow, float* out, bool normalize, cudaStream_t stream) {
    SPLICE_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0, "resize_normalize: empty image");
    TablePair ty, tx;
    int rc = get_tables(h, oh, &ty); if (rc) return rc;
    rc = get_tables(w, ow, &tx); if (rc) return rc;
    dim3 grid(ceil_div(ow, 32), ceil_div(oh, 8), 3);
    resize_normalize_kernel<<<grid, 256, 0, stream>>>(img, h, w, oh, ow, ty.fwd.start, ty.fwd.w, ty.fwd.taps, tx.fwd.start,
                                                      tx.fwd.w, tx.fwd.taps, out, normalize);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

int preprocess_bwd(const float* dpatch, int ldp, int row0, int h, int w, int oh, int ow, int patch, float* dimg,
                   bool normalize, cudaStream_t stream) {
    SPLICE_REQUIRE(oh >= patch && ow >= patch, "preprocess_bwd: %dx%d is smaller than one %d-pixel patch", oh, ow, patch);
    TablePair ty, tx;
    int rc = get_tables(h, oh, &ty); if (rc) return rc;
    rc = get_tables(w, ow, &tx); if (rc) return rc;
    dim3 grid(ceil_div(w, 32), ceil_div(h, 8), 3);
    preprocess_bwd_kernel<<<grid, 256, 0, stream>>>(dpatch, ldp, row0, h, w, oh, ow, ty.bwd.start, ty.bwd.w, ty.bwd.taps,
                                                    tx.bwd.start, tx.bwd.w, tx.bwd.taps, patch, dimg, normalize);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

}  // namespace splice
