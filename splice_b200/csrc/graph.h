// splice_b200 — tiny CUDA-graph cache: a launch sequence that is a pure function of (pointers, shapes) is run eagerly
// the first time (warm-up: lazy allocations, function attributes), captured the second time and replayed afterwards.
// Capture happens on a private stream (the caller's stream may be the legacy default stream, which cannot be
// captured); the instantiated graph is launched on the caller's stream.
#pragma once
#include <functional>
#include <unordered_map>

#include "common.cuh"

namespace splice {

long long launch_count_now();   // capi.cu

struct KeyHasher {
    uint64_t h = 0xcbf29ce484222325ull;
    KeyHasher& add(uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return *this; }
    KeyHasher& add(const void* p) { return add((uint64_t)reinterpret_cast<uintptr_t>(p)); }
};

class GraphCache {
public:
    ~GraphCache() { clear(); if (cap_) cudaStreamDestroy(cap_); }
    void clear() {
        for (auto& kv : map_) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        map_.clear();
    }
    // body(stream) enqueues the launches and returns a SPLICE status
    int run(uint64_t key, cudaStream_t stream, const std::function<int(cudaStream_t)>& body) {
        if (!enabled()) return body(stream);
        Entry& e = map_[key];
        if (e.exec) {
            SPLICE_CHECK_CUDA(cudaGraphLaunch(e.exec, stream));
            count_launch(e.kernels);
            return SPLICE_OK;
        }
        if (e.seen == 0) {   // warm-up
            e.seen = 1;
            return body(stream);
        }
        if (map_.size() > 512) { set_error("graph cache overflow"); return body(stream); }
        if (!cap_) SPLICE_CHECK_CUDA(cudaStreamCreateWithFlags(&cap_, cudaStreamNonBlocking));
        SPLICE_CHECK_CUDA(cudaStreamBeginCapture(cap_, cudaStreamCaptureModeRelaxed));
        const long long before = launch_count_now();
        const int rc = body(cap_);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cap_, &g);
        e.kernels = (int)(launch_count_now() - before);
        count_launch(-e.kernels);   // nothing ran during capture
        if (rc != SPLICE_OK || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            map_.erase(key);
            if (rc != SPLICE_OK) return rc;
            return body(stream);   // capture unsupported here: stay eager
        }
        ce = cudaGraphInstantiate(&e.exec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            e.exec = nullptr;
            map_.erase(key);
            return body(stream);
        }
        SPLICE_CHECK_CUDA(cudaGraphLaunch(e.exec, stream));
        count_launch(e.kernels);
        return SPLICE_OK;
    }
    static bool enabled() {
        static int on = -1;
        if (on < 0) {
            const char* v = getenv("SPLICE_B200_GRAPHS");
            on = (v && v[0] == '0') ? 0 : 1;
        }
        return on == 1;
    }

private:
    struct Entry { int seen = 0; int kernels = 0; cudaGraphExec_t exec = nullptr; };
    std::unordered_map<uint64_t, Entry> map_;
    cudaStream_t cap_ = nullptr;
};

}  // namespace splice
