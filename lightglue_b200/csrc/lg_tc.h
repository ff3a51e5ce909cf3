// Tensor-core (tcgen05 / TMA) path: LG_PREC_BF16 and LG_PREC_BF16X3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lg_internal.h"

struct LgHandle;

struct TcWeights {
  void* blob;  // device allocation owned by the handle
  size_t bytes;
};
struct TcBuffers {
  void* base;
};

int tc_pack_weights(LgHandle* h, cudaStream_t stream);
void tc_free_weights(TcWeights* w);
void tc_carve(size_t* off, char* base, size_t S, int Lp, const LgHandle* h, TcBuffers* out);
// x fp32 [S, Lp, 256] -> bf16 (hi / lo) shadow copies the linears consume
int tc_refresh_shadow(LgHandle* h, const TcBuffers& b, const float* x, const SeqState& st, cudaStream_t stream);
int tc_input_proj(LgHandle* h, const TcBuffers& b, const SeqState& st, const float* desc_packed, float* x, cudaStream_t stream);
int tc_block(LgHandle* h, const TcBuffers& b, const SeqState& st, int layer, int blk, float* x, const float* cs,
             cudaStream_t stream);
int tc_final_proj(LgHandle* h, const TcBuffers& b, const SeqState& st, float* p_out, cudaStream_t stream);
