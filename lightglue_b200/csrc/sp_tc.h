// SuperPoint backbone on the tensor cores (sp_tc.cu): state owned by an SpHandle whose conf asks for it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "lg_handle.h"
#include "sp_pipeline.h"

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

// Stages policy of sp_run_post (sp_pipeline.h): warp / block kernels for candidate compaction, top-k and descriptor
// sampling; top-k beyond SP_SEL_KMAX (4096) falls back to the functor.
struct SpCudaStages {
  cudaStream_t stream;
  int compact_impl(const SpWorkspace& ws, int B, int H, int W, float thr, long cap) const;
  int select_impl(const SpWorkspace& ws, int B, int k, long cap, long out_cap) const;
  int sample_impl(const SpWorkspace& ws, float* kpts, float* kscores, float* desc, int B, int Hc, int Wc, long out_cap) const;
  template <class Exec>
  int compact(Exec&, const SpWorkspace& ws, int B, int H, int W, float thr, long cap) const { return compact_impl(ws, B, H, W, thr, cap); }
  template <class Exec>
  int select(Exec& exec, const SpWorkspace& ws, int B, int k, long cap, long out_cap) const {
    if (k > 4096) return SpFunctorStages().select(exec, ws, B, k, cap, out_cap);
    return select_impl(ws, B, k, cap, out_cap);
  }
  template <class Exec>
  int sample(Exec&, const SpWorkspace& ws, float* kpts, float* kscores, float* desc, int B, int Hc, int Wc, long out_cap) const {
    return sample_impl(ws, kpts, kscores, desc, B, Hc, Wc, out_cap);
  }
};
