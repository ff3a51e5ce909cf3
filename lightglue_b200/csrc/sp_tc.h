// SuperPoint backbone on the tensor cores (sp_tc.cu): state owned by an SpHandle whose conf asks for it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "lg_handle.h"

struct SpTc {
  LgHandle lg;        // carrier for the shared tensor-core linear kernels: precision, packed weights (fp32 + bf16 hi / lo),
                      // tensor-map cache, debug words -- no matcher state
  size_t w_off[12];   // float offsets into lg.wpk: weights repacked to [256, k*k*Cin] ...
  size_t b_off[12];   // ... and biases padded to [256]
};

int sp_tc_create(SpTc** out, const float* wts_dev, cudaStream_t stream);
void sp_tc_destroy(SpTc* t);
size_t sp_tc_workspace_bytes(int B, int H, int W);
int sp_tc_backbone(SpTc* t, const float* wts_dev, const float* image, int B, int H, int W, void* workspace, float* logits_nchw,
                   float* dense_nchw, cudaStream_t stream);
