// Attention for the tensor-core path.  (interim: CUDA-core kernel on the fp16 operands; the
// tcgen05 kernel replaces it behind the same entry point.)
#include "lg_handle.h"

namespace {
#define AT 64
#define ALD 65
__global__ void __launch_bounds__(256) attn_ref_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                       const __half* __restrict__ vt, __nv_bfloat16* __restrict__ ctxh,
                                                       __nv_bfloat16* __restrict__ ctxl, int kv_shift, SeqState st) {
  const int s = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * AT;
  const int len_q = st.len[s];
  if (r0 >= len_q || lg_pair_stopped(st, s)) return;
  const int skv = (s + kv_shift) % st.S;
  const int len_kv = st.len[skv];
  extern __shared__ float sm[];
  float* Qt = sm;
  float* Kt = Qt + AT * ALD;
  float* Vs = Kt + AT * ALD;  // [kv][d]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const __half* qb = q + (((long)s * LG_HEADS + h) * st.Lp + r0) * LG_HDIM;
  const __half* kb = k + ((long)skv * LG_HEADS + h) * st.Lp * LG_HDIM;
  const __half* vb = vt + ((long)skv * LG_HEADS + h) * LG_HDIM * st.Lp;
  for (int e = tid; e < AT * 64; e += 256) {
    const int row = e / 64, d = e % 64;
    Qt[d * ALD + row] = __half2float(qb[(long)row * LG_HDIM + d]);
  }
  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int c0 = 0; c0 < len_kv; c0 += AT) {
    __syncthreads();
    for (int e = tid; e < AT * 64; e += 256) {
      const int row = e / 64, d = e % 64;
      Kt[d * ALD + row] = __half2float(kb[(long)(c0 + row) * LG_HDIM + d]);
      const int d2 = e / 64, kv = e % 64;
      Vs[kv * 64 + d2] = __half2float(vb[(long)d2 * st.Lp + c0 + kv]);
    }
    __syncthreads();
    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Qt[d * ALD + ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Kt[d * ALD + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(af[i], bf[j], sc[i][j]);
    }
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[i][j] = (c0 + tx * 4 + j < len_kv) ? sc[i][j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, sc[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(mrow[i], mx);
      alpha[i] = expf(mrow[i] - mnew);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[i][j] = expf(sc[i][j] - mnew); rs += sc[i][j]; }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      lrow[i] = lrow[i] * alpha[i] + rs;
      mrow[i] = mnew;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Kt[(tx * 4 + j) * ALD + ty * 4 + i] = sc[i][j];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= alpha[i];
#pragma unroll 8
    for (int kv = 0; kv < 64; ++kv) {
      float af[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Kt[kv * ALD + ty * 4 + i];
      const float4 bv = *reinterpret_cast<const float4*>(Vs + kv * 64 + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i][0] = fmaf(af[i], bv.x, o[i][0]); o[i][1] = fmaf(af[i], bv.y, o[i][1]);
        o[i][2] = fmaf(af[i], bv.z, o[i][2]); o[i][3] = fmaf(af[i], bv.w, o[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= len_q) continue;
    const float inv = lrow[i] > 0.f ? 1.f / lrow[i] : 0.f;
    const long off = ((long)s * st.Lp + r) * LG_DIM + h * LG_HDIM + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float val = o[i][j] * inv;
      const __nv_bfloat16 hi = __float2bfloat16_rn(val);
      ctxh[off + j] = hi;
      if (ctxl) ctxl[off + j] = __float2bfloat16_rn(val - __bfloat162float(hi));
    }
  }
}
}  // namespace

int tc_attention(LgHandle* h, const TcBuffers& b, const SeqState& st, int kv_shift, const __half* kbuf, cudaStream_t stream) {
  const size_t smem = (2 * AT * ALD + AT * 64) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lg_set_cuda_error(e, __FILE__, __LINE__);
    attr_set = true;
  }
  dim3 grid(st.Lp / AT, LG_HEADS, st.S);
  attn_ref_kernel<<<grid, 256, smem, stream>>>(b.q, kbuf, b.vt, b.ctxh, b.ctxl, kv_shift, st);
  LG_CHECK_LAUNCH();
  h->launches += 1;
  return 0;
}
