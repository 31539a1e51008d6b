// splice_b200 — image -> ViT patch matrix (and back), see preprocess.cu
#pragma once
#include "common.cuh"

namespace splice {

// Output size of torchvision Resize(size:int, max_size) for an h x w input (util/losses.py:20).
void resized_hw(int h, int w, int size, int max_size, int* oh, int* ow);

// img fp32 [3,h,w] in [0,1]  --antialiased bilinear resize to (oh, ow), ImageNet normalise, patchify-->
// patches bf16 rows [row0 + gy*gw + gx, 3*p*p] (ld = 3*p*p), column index c*p*p + py*p + px.
int preprocess_fwd(const float* img, int h, int w, int oh, int ow, int patch, bf16* patches, int row0, bool normalize,
                   cudaStream_t stream);

// Adjoint: dpatch fp32 [*, ldp] in the same patch layout, row of patch (gy,gx) = row0 + gy*gw + gx
// -> dimg fp32 [3,h,w] (overwritten):  dimg = Resize^T( dpatch / std ).
int preprocess_bwd(const float* dpatch, int ldp, int row0, int h, int w, int oh, int ow, int patch, float* dimg,
                   bool normalize, cudaStream_t stream);

// out fp32 [3,oh,ow] = Normalize(Resize(img))  (standalone LossG.global_transform)
int resize_normalize(const float* img, int h, int w, int oh, int ow, float* out, bool normalize, cudaStream_t stream);

}  // namespace splice
