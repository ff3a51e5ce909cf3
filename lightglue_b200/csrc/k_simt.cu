// fp32 CUDA-core kernels: the LG_PREC_FP32 path (reference-grade arithmetic, used for index-exact
// parity with the CPU oracle; lightglue.py line numbers in comments refer to /root/reference).
#include "lg_internal.h"

// ------------------------------------------------------------------------------------------------
// C[rows, Nout] = [A0 | A1][rows, K] * W[Nout, K]^T + bias   with fused epilogues.
// 64x64 output tile, BK = 16, 256 threads, 4x4 register micro-tile per thread.
// ------------------------------------------------------------------------------------------------
#define GB 64
#define GK 16

__global__ void __launch_bounds__(256) simt_gemm_kernel(GemmArgs a, SeqState st) {
  const int rows_per_seq_tiles = st.Lp / GB;
  const int s = blockIdx.y / rows_per_seq_tiles;
  const int r0 = (blockIdx.y % rows_per_seq_tiles) * GB;  // row inside the sequence
  if (r0 >= st.len[s]) return;
  const int pair = s >= st.B ? s - st.B : s;
  const int sl = st.stop_layer[pair];
  const float* __restrict__ W = a.W;
  const float* __restrict__ bias = a.bias;
  if (a.w_sel_stride > 0) {  // per-pair head (assignment uses log_assignment[stop - 1], lightglue.py:591)
    W += (long)(sl - 1) * a.w_sel_stride;
    if (bias) bias += (long)(sl - 1) * a.b_sel_stride;
  } else if (sl != 0) {
    return;  // pair already exited (lightglue.py:549-550)
  }
  const int n0 = blockIdx.x * GB;
  const long grow0 = (long)s * st.Lp + r0;

  __shared__ float As[GK][GB + 4];
  __shared__ float Bs[GK][GB + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int lr = tid / 4, lk = (tid % 4) * 4;  // loader: row lr, k offset lk..lk+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += GK) {
    const float* ap;
    if (k0 < a.K0) ap = a.A0 + (grow0 + lr) * a.lda0 + k0 + lk;
    else ap = a.A1 + (grow0 + lr) * a.lda1 + (k0 - a.K0) + lk;
    const float4 av = *reinterpret_cast<const float4*>(ap);
    const float4 wv = *reinterpret_cast<const float4*>(W + (long)(n0 + lr) * a.K + k0 + lk);
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk + 0][lr] = wv.x; Bs[lk + 1][lr] = wv.y; Bs[lk + 2][lr] = wv.z; Bs[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int len = st.len[s];
  const int c0 = n0 + tx * 4;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = bias ? bias[c0 + j] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= len) continue;  // padding rows are never written
    const long grow = (long)s * st.Lp + r;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (acc[i][j] + bv[j]) * a.scale;
    if (a.epi == EPI_STORE) {
      *reinterpret_cast<float4*>(a.out + grow * a.ldo + c0) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (a.epi == EPI_RESID) {
      float4* o = reinterpret_cast<float4*>(a.out + grow * a.ldo + c0);
      float4 x = *o;
      *o = make_float4(x.x + v[0], x.y + v[1], x.z + v[2], x.w + v[3]);
    } else if (a.epi == EPI_QKV_ROPE) {
      // columns were permuted at pack time to [q | k | v], each head-major (h*64 + d)
      const int which = c0 / LG_DIM, h = (c0 % LG_DIM) / LG_HDIM, d = c0 % LG_HDIM;
      float* dst = (which == 0 ? a.q : which == 1 ? a.k : a.v) + (((long)s * LG_HEADS + h) * st.Lp + r) * LG_HDIM + d;
      if (which < 2) {  // rotary embedding on q and k only (lightglue.py:168-169, 58-65)
        const float* csr = a.cs + grow * 64;
        const float ca = csr[d / 2], cb = csr[d / 2 + 1], sa = csr[32 + d / 2], sb = csr[32 + d / 2 + 1];
        const float o0 = v[0] * ca - v[1] * sa, o1 = v[1] * ca + v[0] * sa;
        const float o2 = v[2] * cb - v[3] * sb, o3 = v[3] * cb + v[2] * sb;
        *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {  // EPI_QK_V: cross block, [to_qk | to_v] (lightglue.py:204-209), no positional encoding
      const int which = c0 / LG_DIM, h = (c0 % LG_DIM) / LG_HDIM, d = c0 % LG_HDIM;
      float* dst = (which == 0 ? a.q : a.v) + (((long)s * LG_HEADS + h) * st.Lp + r) * LG_HDIM + d;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int simt_gemm(const GemmArgs& a, const SeqState& st, cudaStream_t stream) {
  dim3 grid(a.Nout / GB, st.S * (st.Lp / GB));
  simt_gemm_kernel<<<grid, 256, 0, stream>>>(a, st);
  LG_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Flash-style fp32 attention (lightglue.py:113-137): 64 queries x 64 keys per step, online softmax.
// ------------------------------------------------------------------------------------------------
#define AT 64
#define ALD 65

__global__ void __launch_bounds__(256) simt_attention_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             const float* __restrict__ v, float* __restrict__ ctx,
                                                             int kv_shift, SeqState st) {
  const int s = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * AT;
  const int len_q = st.len[s];
  if (r0 >= len_q || lg_pair_stopped(st, s)) return;
  const int skv = (s + kv_shift) % st.S;
  const int len_kv = st.len[skv];
  extern __shared__ float sm[];
  float* Qt = sm;                 // [64 d][ALD]   Q^T
  float* Kt = Qt + AT * ALD;      // [64 d][ALD]   K^T, later aliased by P^T [kv][row]
  float* Vs = Kt + AT * ALD;      // [64 kv][64 d]
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* qb = q + (((long)s * LG_HEADS + h) * st.Lp + r0) * LG_HDIM;
  const float* kb = k + ((long)skv * LG_HEADS + h) * st.Lp * LG_HDIM;
  const float* vb = v + ((long)skv * LG_HEADS + h) * st.Lp * LG_HDIM;
  for (int e = tid; e < AT * 16; e += 256) {
    const int row = e / 16, d4 = (e % 16) * 4;
    const float4 t = *reinterpret_cast<const float4*>(qb + (long)row * LG_HDIM + d4);
    Qt[(d4 + 0) * ALD + row] = t.x; Qt[(d4 + 1) * ALD + row] = t.y;
    Qt[(d4 + 2) * ALD + row] = t.z; Qt[(d4 + 3) * ALD + row] = t.w;
  }
  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int c0 = 0; c0 < len_kv; c0 += AT) {
    __syncthreads();  // previous P^T / V fully consumed
    for (int e = tid; e < AT * 16; e += 256) {
      const int row = e / 16, d4 = (e % 16) * 4;
      const float4 t = *reinterpret_cast<const float4*>(kb + (long)(c0 + row) * LG_HDIM + d4);
      Kt[(d4 + 0) * ALD + row] = t.x; Kt[(d4 + 1) * ALD + row] = t.y;
      Kt[(d4 + 2) * ALD + row] = t.z; Kt[(d4 + 3) * ALD + row] = t.w;
      *reinterpret_cast<float4*>(Vs + row * 64 + d4) =
          *reinterpret_cast<const float4*>(vb + (long)(c0 + row) * LG_HDIM + d4);
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
      const float mnew = fmaxf(mrow[i], mx);  // finite: every tile has >= 1 valid key
      alpha[i] = expf(mrow[i] - mnew);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[i][j] = expf(sc[i][j] - mnew);
        rs += sc[i][j];
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      lrow[i] = lrow[i] * alpha[i] + rs;
      mrow[i] = mnew;
    }
    __syncthreads();  // all S tiles computed: K^T may be overwritten by P^T
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
    const float inv = lrow[i] > 0.f ? 1.f / lrow[i] : 0.f;  // no keys -> zeros (lightglue.py:114-115)
    float* dst = ctx + ((long)s * st.Lp + r) * LG_DIM + h * LG_HDIM + tx * 4;  // heads concatenated h-major (171)
    *reinterpret_cast<float4*>(dst) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
  }
}

int simt_attention(const float* q, const float* k, const float* v, float* ctx, int kv_shift, const SeqState& st,
                   cudaStream_t stream) {
  const size_t smem = (2 * AT * ALD + AT * 64) * sizeof(float);
  if (int r = lg_func_smem_once((const void*)simt_attention_kernel, (int)smem)) return r;
  dim3 grid(st.Lp / AT, LG_HEADS, st.S);
  simt_attention_kernel<<<grid, 256, smem, stream>>>(q, k, v, ctx, kv_shift, st);
  LG_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm(512, eps 1e-5, affine) + exact GELU, in place, one warp per row (lightglue.py:154-155).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simt_ln_gelu_kernel(float* __restrict__ h, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, SeqState st) {
  const int row = blockIdx.x * 8 + threadIdx.x / 32;
  const int s = row / st.Lp, r = row % st.Lp;
  if (s >= st.S || r >= st.len[s] || lg_pair_stopped(st, s)) return;
  const int lane = threadIdx.x % 32;
  float* p = h + (long)row * LG_FFN;
  float v[16];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + i * 128 + lane * 4);
    v[i * 4 + 0] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    sum += t.x + t.y + t.z + t.w;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum * (1.f / LG_FFN);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; var = fmaf(d, d, var); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
  const float rstd = rsqrtf(var * (1.f / LG_FFN) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float y[4] = {(v[i * 4 + 0] - mean) * rstd * g.x + b.x, (v[i * 4 + 1] - mean) * rstd * g.y + b.y,
                  (v[i * 4 + 2] - mean) * rstd * g.z + b.z, (v[i * 4 + 3] - mean) * rstd * g.w + b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = 0.5f * y[j] * (1.f + erff(y[j] * 0.70710678118654752f));
    *reinterpret_cast<float4*>(p + c) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

int simt_layernorm_gelu(float* h, const float* gamma, const float* beta, const SeqState& st, cudaStream_t stream) {
  const int rows = st.S * st.Lp;
  simt_ln_gelu_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(h, gamma, beta, st);
  LG_CHECK_LAUNCH();
  return 0;
}
