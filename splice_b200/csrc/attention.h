// splice_b200 — fused attention interface (see attention.cu)
#pragma once
#include "common.cuh"

namespace splice {
// qkv bf16 [S*t, 3D]; o bf16 [S*t, D]; lse fp32 [S,H,t] (log2 domain)
int attention_fwd(const bf16* qkv, bf16* o, float* lse, int S, int t, int D, int H, cudaStream_t stream);
// dout bf16 [S*t, D]; delta fp32 [S,H,t] scratch; dqkv bf16 [S*t, 3D] (fully overwritten)
int attention_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int S, int t,
                  int D, int H, cudaStream_t stream);
// tcgen05 / TMEM / TMA implementations (attention_tc.cu); same contracts
int attention_fwd_tc(const bf16* qkv, bf16* o, float* lse, int S, int t, int D, int H, cudaStream_t stream);
int attention_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int S, int t,
                     int D, int H, cudaStream_t stream);
}  // namespace splice
