// splice_b200 — bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulator,
// TMA-staged 128B-swizzled operand tiles), with the ViT's elementwise work fused into the epilogue.
//
// Replaces, for the frozen DINO ViT of the reference, every cuBLAS call behind
//   models/extractor.py:83,91,99 (self.model(input_img)): qkv / proj / fc1 / fc2 linears, the patch-embed
//   conv (k = stride = patch, i.e. a GEMM over patchified pixels), and the dgrad halves of their backward.
//
// Kernel shape: one CTA per 128 x BN output tile, 6 warps:
//   warp 0      TMA producer   (one elected lane; STAGES-deep ring of {A 128x64, B BNx64} bf16 tiles)
//   warp 1      MMA issuer     (one elected lane; 4 x tcgen05.mma K=16 per stage; owns TMEM alloc/dealloc)
//   warps 2..5  epilogue       (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global)
// Shared memory per CTA is kept <= ~100 KB so that two CTAs are co-resident per SM: while one CTA drains
// its accumulator through the epilogue, the other one's mainloop keeps the tensor pipe busy.
#include <cudaTypedefs.h>
#include <stdio.h>

#include <mutex>
#include <unordered_map>

#include "gemm.h"

namespace splice {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr uint32_t A_BYTES = BM * BK * 2;
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TILES_BYTES = STAGES * STAGE_BYTES;
    static constexpr uint32_t BAR_BYTES = 256;
    static constexpr uint32_t TOTAL = TILES_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                         int K, GemmEpilogue ep) {
    using L = GemmSmem<BN, STAGES>;
    constexpr uint32_t TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + STAGES * L::A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::TILES_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int num_kb = K / BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
                tma_load_2d(smemA + s * L::A_BYTES, &tmA, &full_bar[s], kb * BK, m0);
                tma_load_2d(smemB + s * L::B_BYTES, &tmB, &full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smemA + s * L::A_BYTES));
                const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + s * L::B_BYTES));
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                    // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
                    umma_bf16_ss(tmem_base, adesc + 2u * j, bdesc + 2u * j, idesc, (kb | j) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
            }
            umma_commit(tmem_full_bar);  // accumulator complete
        }
    } else {
        // ---------------- epilogue: warps 2..5, TMEM lane quadrant = warp % 4 ----------------
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        const GemmRowCtx rc = row < M ? gemm_epilogue_row(ep, row) : GemmRowCtx{0.f, 1.f};
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            const int col = n0 + c * 32;
            if (col < N) {  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), r);
                tmem_ld_wait();
                if (row < M) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    gemm_epilogue_chunk(ep, row, col, v, rc);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant (product path): one CTA per SM loops over output tiles (n fastest, so co-scheduled CTAs share
// the A row block through L2). The accumulator is double-buffered in TMEM (2 x BN columns): while the 8 epilogue
// warps drain tile i (TMEM -> registers -> fused epilogue -> global), the MMA warp already accumulates tile i+1.
//   warp 0      TMA producer        warp 1   MMA issuer (+ TMEM alloc/dealloc)       warps 2..9   epilogue
// Epilogue warp e reads TMEM lane quadrant (warp_id % 4) and the column half (e / 4) of the tile.
//
// Thread-block clusters + TMA multicast (CM x CN CTAs per cluster, CM along M, CN along N): measured on B200 the
// 1-CTA kernel is bound by L2 -> shared-memory operand traffic (~7 TB/s for the whole chip: 128x256 tiles need
// 96 B/clk/SM, the fabric gives ~25), not by the tensor pipe. A cluster works on a (CM*128) x (CN*BN) super-tile:
// the CN CTAs of a row share their A tile (each loads 1/CN of its rows and multicasts it to the row), the CM CTAs of a
// column share their B tile the same way, so every operand byte crosses the L2 fabric once per cluster instead of
// once per CTA. Pipeline protocol: a stage of CTA X is written by all CTAs of X's row and column ("peers"), so
//   full_bar[s]  (count 1)            X's own producer arms it with the full stage size; peers' multicasts add bytes
//   empty_bar[s] (count CM + CN - 1)  every peer's MMA warp commits to it (tcgen05.commit multicast) once its MMAs
//                                     have read stage s; X's producer may then overwrite stage s in all its peers.
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int CM, int CN>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M,
                                    int N, int K, GemmEpilogue ep) {
    using L = GemmSmem<BN, STAGES>;
    constexpr uint32_t ACC_STRIDE = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;   // power-of-two column offset
    constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;                                                 // of the second accumulator
    static_assert(BN <= 256, "two accumulator stages must fit the 512 TMEM columns");
    constexpr int CS = CM * CN;                       // cluster size
    constexpr int A_ROWS = BM / CN, B_ROWS = BN / CM;   // rows of the A / B tile this CTA loads (and multicasts)
    static_assert(A_ROWS % 8 == 0 && B_ROWS % 8 == 0, "operand slices must be whole 8-row swizzle atoms");
    static_assert(CS <= 8, "portable cluster size");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + STAGES * L::A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::TILES_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;    // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = K / BK;
    // cluster geometry: rank r -> (mi, ni) = (r % CM, r / CM)
    const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
    const int mi = (int)(crank % CM), ni = (int)(crank / CM);
    const int cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
    const int stn = (N + CN * BN - 1) / (CN * BN);
    const int num_super = stn * ((M + CM * BM - 1) / (CM * BM));
    uint16_t mask_a = 0, mask_b = 0;                 // receivers of my A slice (my row) / my B slice (my column)
#pragma unroll
    for (int j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (j * CM + mi));
#pragma unroll
    for (int i = 0; i < CM; ++i) mask_b |= (uint16_t)(1u << (ni * CM + i));
    const uint16_t mask_peers = mask_a | mask_b;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CM + CN - 1);
        }
        mbar_init(&tmem_full_bar[0], 1);
        mbar_init(&tmem_full_bar[1], 1);
        mbar_init(&tmem_empty_bar[0], 8);   // one arrival per epilogue warp
        mbar_init(&tmem_empty_bar[1], 8);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();   // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail. The next kernel may be
    // scheduled from here on (it waits for this grid to complete before touching memory); each role waits for the
    // previous grid right before ITS first dependent global access - the producer only after it has put the first
    // weight (B) tiles in flight when the caller declared B constant, the epilogue before its first load / store.
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;   // running k-block counter across tiles (ring position)
            bool first = true;
            for (int st = cluster_id; st < num_super; st += num_clusters) {
                const int m0 = ((st / stn) * CM + mi) * BM, n0 = ((st % stn) * CN + ni) * BN;
                int kb0 = 0;
                if (first) {
                    first = false;
                    const int npre = ep.b_const ? (num_kb < STAGES ? num_kb : STAGES) : 0;
                    for (int kb = 0; kb < npre; ++kb) {    // fresh ring: every stage is free
                        mbar_arrive_expect_tx(&full_bar[kb], L::STAGE_BYTES);
                        uint8_t* dB = smemB + kb * L::B_BYTES + mi * (B_ROWS * BK * 2);
                        if constexpr (CM > 1) tma_load_2d_mc(dB, &tmB, &full_bar[kb], kb * BK, n0 + mi * B_ROWS, mask_b);
                        else tma_load_2d(dB, &tmB, &full_bar[kb], kb * BK, n0);
                    }
                    pdl_wait();
                    for (int kb = 0; kb < npre; ++kb) {
                        uint8_t* dA = smemA + kb * L::A_BYTES + ni * (A_ROWS * BK * 2);
                        if constexpr (CN > 1) tma_load_2d_mc(dA, &tmA, &full_bar[kb], kb * BK, m0 + ni * A_ROWS, mask_a);
                        else tma_load_2d(dA, &tmA, &full_bar[kb], kb * BK, m0);
                    }
                    it = npre;
                    kb0 = npre;
                }
                for (int kb = kb0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1u);      // all peers' MMAs have read stage s
                    mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
                    uint8_t* dA = smemA + s * L::A_BYTES + ni * (A_ROWS * BK * 2);
                    uint8_t* dB = smemB + s * L::B_BYTES + mi * (B_ROWS * BK * 2);
                    if constexpr (CN > 1) tma_load_2d_mc(dA, &tmA, &full_bar[s], kb * BK, m0 + ni * A_ROWS, mask_a);
                    else tma_load_2d(dA, &tmA, &full_bar[s], kb * BK, m0);
                    if constexpr (CM > 1) tma_load_2d_mc(dB, &tmB, &full_bar[s], kb * BK, n0 + mi * B_ROWS, mask_b);
                    else tma_load_2d(dB, &tmB, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            uint32_t it = 0, lt = 0;   // k-block counter, local tile counter
            for (int st = cluster_id; st < num_super; st += num_clusters, ++lt) {
                const uint32_t as = lt & 1u;
                mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * ACC_STRIDE;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smemA + s * L::A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + s * L::B_BYTES));
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j)
                        umma_bf16_ss(tacc, adesc + 2u * j, bdesc + 2u * j, idesc, (kb | j) != 0 ? 1u : 0u);
                    if constexpr (CS > 1) umma_commit_mc(&empty_bar[s], mask_peers);
                    else umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        const int e = warp - 2;            // 0..7
        const int quad = warp & 3;         // TMEM lane quadrant this warp may access
        const int half = e >> 2;           // column half of the tile
        constexpr int CHUNKS = BN / 32, CH_PER_HALF = (CHUNKS + 1) / 2;
        uint32_t lt = 0;
        pdl_wait();
        for (int st = cluster_id; st < num_super; st += num_clusters, ++lt) {
            const int m0 = ((st / stn) * CM + mi) * BM, n0 = ((st % stn) * CN + ni) * BN;
            const uint32_t as = lt & 1u;
            const int row = m0 + quad * 32 + lane;
            const GemmRowCtx rc = (row < M && m0 < M) ? gemm_epilogue_row(ep, row) : GemmRowCtx{0.f, 1.f};   // under the MMA wait
            mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < CH_PER_HALF; ++cc) {
                const int c = half * CH_PER_HALF + cc;
                const int col = n0 + c * 32;
                if (c < CHUNKS && col < N && m0 < M) {  // warp-uniform (a padded tile of the super-tile stores nothing)
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + as * ACC_STRIDE + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), r);
                    tmem_ld_wait();
                    if (row < M) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        gemm_epilogue_chunk(ep, row, col, v, rc);
                    }
                }
            }
            // all of this warp's TMEM reads of stage `as` are complete: hand the stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();   // no CTA exits while a peer may still signal its barriers
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check kernel (same operands, same epilogue). Not on the product path: selected only by
// impl == GEMM_IMPL_SIMT from the unit tests to separate "tcgen05 plumbing" from "epilogue logic".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gemm_bf16_simt_kernel(const bf16* __restrict__ A, int lda,
                                                             const bf16* __restrict__ B, int ldb, int M, int N, int K,
                                                             GemmEpilogue ep) {
    __shared__ float As[128][33];
    __shared__ float Bs[32][33];
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 32;
    const int tid = threadIdx.x;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = tid; i < 128 * 32; i += 128) {
            const int r = i >> 5, c = i & 31;
            As[r][c] = (m0 + r < M) ? __bfloat162float(A[(size_t)(m0 + r) * lda + k0 + c]) : 0.f;
        }
        for (int i = tid; i < 32 * 32; i += 128) {
            const int r = i >> 5, c = i & 31;
            Bs[r][c] = (n0 + r < N) ? __bfloat162float(B[(size_t)(n0 + r) * ldb + k0 + c]) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const float a = As[tid][k];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, Bs[j][k], acc[j]);
        }
        __syncthreads();
    }
    const int row = m0 + tid;
    if (row < M && n0 < N) gemm_epilogue_chunk(ep, row, n0, acc, gemm_epilogue_row(ep, row));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_tmap_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            return nullptr;
        }
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// Encoded tensor maps are cached: the engine's operands live at stable addresses (weights, per-slot workspaces), so
// after the first step every GEMM launch finds its two descriptors here instead of paying two driver calls.
struct TmapKey {
    const void* ptr; int rows, cols, ld, box_rows;
    bool operator==(const TmapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + ((size_t)k.cols << 20) + ((size_t)k.ld << 40) + (size_t)k.box_rows;
        return h;
    }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;

static int make_tmap_bf16_uncached(CUtensorMap* tm, const bf16* ptr, int rows, int cols, int ld, int box_rows);

// 2D bf16 row-major [rows, cols] with leading dimension ld (elements); box = box_rows x 64, 128B swizzle.
int make_tmap_bf16(CUtensorMap* tm, const bf16* ptr, int rows, int cols, int ld, int box_rows) {
    const TmapKey key{ptr, rows, cols, ld, box_rows};
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *tm = it->second; return SPLICE_OK; }
    int rc = make_tmap_bf16_uncached(tm, ptr, rows, cols, ld, box_rows);
    if (rc) return rc;
    if (g_tmaps.size() > 4096) g_tmaps.clear();   // unbounded growth guard (pointer churn from ad-hoc callers)
    g_tmaps.emplace(key, *tm);
    return SPLICE_OK;
}

static int make_tmap_bf16_uncached(CUtensorMap* tm, const bf16* ptr, int rows, int cols, int ld, int box_rows) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_tmap_encoder();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old or no GPU)");
        return SPLICE_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box_rows=%d ptr=%p", (int)r, rows, cols, ld,
                  box_rows, (const void*)ptr);
        return SPLICE_ERR_CUDA;
    }
    return SPLICE_OK;
}

template <int BN, int STAGES>
static int launch_tcgen05(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, const GemmEpilogue& ep,
                          cudaStream_t stream) {
    using L = GemmSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        SPLICE_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, STAGES>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    int rc = make_tmap_bf16(&tmA, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB, B, N, K, ldb, BN);
    if (rc) return rc;
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
    gemm_bf16_tcgen05_kernel<BN, STAGES><<<grid, 192, L::TOTAL, stream>>>(tmA, tmB, M, N, K, ep);
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

template <int BN, int STAGES, int CM, int CN>
static int launch_persistent(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, const GemmEpilogue& ep,
                             cudaStream_t stream) {
    using L = GemmSmem<BN, STAGES>;
    constexpr int CS = CM * CN;
    auto kernel = gemm_bf16_tcgen05_persistent_kernel<BN, STAGES, CM, CN>;
    static int max_clusters = 0;   // co-resident clusters of this kernel (persistent grid size / CS)
    if (max_clusters == 0) {
        SPLICE_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL));
        int dev = 0, sms = 148;
        SPLICE_CHECK_CUDA(cudaGetDevice(&dev));
        SPLICE_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int n = sms / CS;
        if (CS > 1) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(sms / CS * CS); q.blockDim = dim3(320); q.dynamicSmemBytes = L::TOTAL;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int nc = 0;
            SPLICE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel, &q));
            SPLICE_REQUIRE(nc > 0, "gemm: no %d-CTA cluster of the %dx%d kernel fits this device", CS, BM, BN);
            if (nc < n) n = nc;
        }
        max_clusters = n;
    }
    static int prefetch_b = -1;   // SPLICE_B200_GEMM_PREFETCH=0: no weight tiles ahead of the dependent-launch wait (A/B aid)
    if (prefetch_b < 0) { const char* v = getenv("SPLICE_B200_GEMM_PREFETCH"); prefetch_b = (v && v[0] == '0') ? 0 : 1; }
    GemmEpilogue epl = ep;
    if (!prefetch_b) epl.b_const = 0;
    CUtensorMap tmA, tmB;
    int rc = make_tmap_bf16(&tmA, A, M, K, lda, BM / CN);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB, B, N, K, ldb, BN / CM);
    if (rc) return rc;
    const int supers = ceil_div(N, CN * BN) * ceil_div(M, CM * BM);
    const int clusters = supers < max_clusters ? supers : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CS);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CS > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CS; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    SPLICE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, M, N, K, epl));
    SPLICE_LAUNCH_CHECK();
    return SPLICE_OK;
}

// cluster shape used when the caller does not ask for one (bn_hint < 1000): SPLICE_B200_GEMM_CLUSTER="<cm>x<cn>"
// overrides the built-in choice (tuning aid)
static void default_cluster(int bn, int M, int N, int* cm, int* cn) {
    static int env_cm = -1, env_cn = -1;
    if (env_cm < 0) {
        env_cm = 0; env_cn = 0;
        const char* v = getenv("SPLICE_B200_GEMM_CLUSTER");
        if (v && v[0] >= '1' && v[0] <= '8' && v[1] == 'x' && v[2] >= '1' && v[2] <= '8') { env_cm = v[0] - '0'; env_cn = v[2] - '0'; }
    }
    if (env_cm > 0) { *cm = env_cm; *cn = env_cn; return; }
    (void)bn; (void)M; (void)N;
    *cm = 1; *cn = 1;
}

int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, const GemmEpilogue& ep, int impl,
                 int bn_hint, cudaStream_t stream) {
    SPLICE_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    SPLICE_REQUIRE(K % BK == 0, "gemm: K=%d must be a multiple of %d", K, BK);
    SPLICE_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
    SPLICE_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda=%d / ldb=%d must be multiples of 8 (16-byte rows)", lda, ldb);
    SPLICE_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm: A/B must be 16-byte aligned");
    SPLICE_REQUIRE(!ep.c32 || (ep.ldc32 % 4 == 0 && ((uintptr_t)ep.c32 & 15) == 0), "gemm: c32 alignment");
    SPLICE_REQUIRE(!ep.c16 || (ep.ldc16 % 8 == 0 && ((uintptr_t)ep.c16 & 15) == 0), "gemm: c16 alignment");
    SPLICE_REQUIRE(!ep.residual || (ep.ldr % 4 == 0 && ((uintptr_t)ep.residual & 15) == 0), "gemm: residual alignment");
    SPLICE_REQUIRE(!ep.bias || ((uintptr_t)ep.bias & 15) == 0, "gemm: bias alignment");
    SPLICE_REQUIRE(ep.act == GEMM_ACT_NONE || ep.act == GEMM_ACT_GELU || ep.act == GEMM_ACT_GELU_GRAD, "gemm: bad act %d",
                   ep.act);
    SPLICE_REQUIRE(ep.act != GEMM_ACT_GELU_GRAD || ep.aux16, "gemm: GELU_GRAD needs aux16");
    SPLICE_REQUIRE(!ep.aux16 || (ep.ldaux % 8 == 0 && ((uintptr_t)ep.aux16 & 15) == 0), "gemm: aux16 alignment");
    SPLICE_REQUIRE(!ep.pos || (ep.ldpos % 4 == 0 && ((uintptr_t)ep.pos & 15) == 0), "gemm: pos alignment");
    SPLICE_REQUIRE(!ep.slice32 || (ep.slice_c0 % 32 == 0 && ep.slice_c1 % 32 == 0 && ep.ldslice % 4 == 0 &&
                                   ((uintptr_t)ep.slice32 & 15) == 0),
                   "gemm: slice32 alignment");
    SPLICE_REQUIRE(ep.c32 || ep.c16 || ep.slice32, "gemm: no output requested");

    if (impl == GEMM_IMPL_SIMT) {
        dim3 grid(ceil_div(N, 32), ceil_div(M, 128));
        gemm_bf16_simt_kernel<<<grid, 128, 0, stream>>>(A, lda, B, ldb, M, N, K, ep);
        SPLICE_LAUNCH_CHECK();
        return SPLICE_OK;
    }
    if (impl == GEMM_IMPL_TCGEN05_TILE) {   // first-generation one-tile-per-CTA kernel (kept for A/B comparison)
#ifdef SPLICE_B200_CROSSCHECK
        switch (bn_hint) {
            case 64:  return launch_tcgen05<64, 4>(A, lda, B, ldb, M, N, K, ep, stream);
            case 256: return launch_tcgen05<256, 4>(A, lda, B, ldb, M, N, K, ep, stream);
            default:  return launch_tcgen05<128, 3>(A, lda, B, ldb, M, N, K, ep, stream);
        }
#else
        set_error("gemm: the one-tile-per-CTA kernel is cross-check code, not in this build (SPLICE_B200_CROSSCHECK=1 python -m splice_b200.build)");
        return SPLICE_ERR_UNSUPPORTED;
#endif
    }
    SPLICE_REQUIRE(impl == GEMM_IMPL_TCGEN05, "gemm: unknown impl %d", impl);

    // bn_hint = bn + 1000 * cm + 10000 * cn; bn == 0 -> automatic tile width, cm == cn == 0 -> automatic cluster shape
    int bn = bn_hint % 1000, cm = (bn_hint / 1000) % 10, cn = (bn_hint / 10000) % 10;
    if (bn == 0) {
        // Measured on B200 (tools/gpu_checks.py gemm_tc_timing, ViT-B/8 shapes): the 128x256 tile wins whenever there are
        // enough tiles to occupy the 148 persistent CTAs (it halves the A re-reads through L2, which bound the 128-wide
        // tiles); the short-M backward GEMMs with N = 768 prefer 128x128; sub-128 widths only for narrow outputs.
        // (128x96 tiles - 104 instead of 78 CTAs at M = 1570, N = 768 - measured no faster than 128x128: not selected.)
        if (N % 256 == 0 && (M >= 2048 || N >= 2048)) bn = 256;
        else if (N >= 128) bn = 128;
        else bn = 64;
    }
    if (cm == 0 || cn == 0) default_cluster(bn, M, N, &cm, &cn);
#define SPLICE_GEMM_CASE(BN_, ST_, CM_, CN_) \
    if (bn == BN_ && cm == CM_ && cn == CN_) return launch_persistent<BN_, ST_, CM_, CN_>(A, lda, B, ldb, M, N, K, ep, stream)
    // the product build holds the shapes the dispatch above selects (one CTA per tile column, no cluster); the 96-wide tile
    // and the thread-block-cluster / TMA-multicast variants were measured no faster (profiles/ANALYSIS_r1.md) and are
    // cross-check code: SPLICE_B200_CROSSCHECK=1 python -m splice_b200.build
    SPLICE_GEMM_CASE(64, 6, 1, 1);
    SPLICE_GEMM_CASE(128, 5, 1, 1);
    SPLICE_GEMM_CASE(256, 4, 1, 1);
#ifdef SPLICE_B200_CROSSCHECK
    SPLICE_GEMM_CASE(96, 6, 1, 1);
    SPLICE_GEMM_CASE(128, 5, 2, 1); SPLICE_GEMM_CASE(128, 5, 1, 2);
    SPLICE_GEMM_CASE(128, 5, 2, 2); SPLICE_GEMM_CASE(128, 5, 4, 1); SPLICE_GEMM_CASE(128, 5, 4, 2);
    SPLICE_GEMM_CASE(256, 4, 2, 1); SPLICE_GEMM_CASE(256, 4, 1, 2);
    SPLICE_GEMM_CASE(256, 4, 2, 2); SPLICE_GEMM_CASE(256, 4, 4, 1); SPLICE_GEMM_CASE(256, 4, 4, 2);
#endif
#undef SPLICE_GEMM_CASE
    set_error("gemm: tile/cluster BN=%d cluster %dx%d is not in this build (cross-check shapes: SPLICE_B200_CROSSCHECK=1 python -m splice_b200.build)", bn, cm, cn);
    return SPLICE_ERR_UNSUPPORTED;
}

}  // namespace splice
