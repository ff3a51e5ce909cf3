// Microbenchmark of the attention kernel's exponential phase on one SM: cycles per 32-key half block (32 MUFU.EX2 per
// thread) for 1 / 2 / 4 warps per scheduler, with and without the TMEM store of P, for the MUFU-only and polynomial mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../lightglue_b200/csrc micro_exp.cu -o micro_exp
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) { uint64_t r; asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void exp2_poly_x2(uint64_t x2, float& e0, float& e1) {
  float x0, x1; unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t xi = add2_rm(xc, pack2(12582912.f, 12582912.f));
  const uint64_t xf = add2(xi, pack2(-12582912.f, -12582912.f));
  const uint64_t f = fma2(xf, pack2(-1.f, -1.f), xc);
  uint64_t pq = fma2(pack2(0.07807237654924393f, 0.07807237654924393f), f, pack2(0.2259994000196457f, 0.2259994000196457f));
  pq = fma2(pq, f, pack2(0.6958566308021545f, 0.6958566308021545f));
  pq = fma2(pq, f, pack2(0.9999241232872009f, 0.9999241232872009f));
  float p0, p1, i0, i1; unpack2(pq, p0, p1); unpack2(xi, i0, i1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(i0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(i1) << 23));
}
// MODE bit0: store P to TMEM; bit1: no F2FP/sum (MUFU + FFMA2 only)
template <int POLY, int MODE>
__device__ __forceinline__ void exp_half(const uint32_t (&sv)[32], uint32_t tdst, uint64_t sc2, uint64_t nm2, uint64_t& la, uint64_t& lb, uint32_t& sink) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x2 = fma2(pack2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sc2, nm2);
    float e0, e1;
    if ((((i % 8) + 1) * POLY) / 8 != ((i % 8) * POLY) / 8) exp2_poly_x2(x2, e0, e1);
    else { float x0, x1; unpack2(x2, x0, x1); e0 = ex2(x0); e1 = ex2(x1); }
    if (MODE & 2) { pk[i] = __float_as_uint(e0) ^ __float_as_uint(e1); }
    else {
      if (i & 1) lb = add2(lb, pack2(e0, e1)); else la = add2(la, pack2(e0, e1));
      const __half2 hh = __floats2half2_rn(e0, e1);
      pk[i] = *reinterpret_cast<const uint32_t*>(&hh);
    }
  }
  if (MODE & 1) tmem_st16(tdst, pk);
  else {
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= pk[i];
  }
}

template <int POLY, int MODE>
__global__ void __launch_bounds__(512, 1) bench(const float* in, float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp % 4) * 32) << 16) + (warp / 4) * 16;
  uint32_t sv[32];
  for (int i = 0; i < 32; ++i) sv[i] = __float_as_uint(in[threadIdx.x * 32 + i]);
  uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
  uint32_t sink = 0;
  const uint64_t sc2 = pack2(0.18f, 0.18f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const float nmc = -0.001f * it;
    exp_half<POLY, MODE>(sv, tb, sc2, pack2(nmc, nmc), la, lb, sink);
    if (MODE & 1) { tmem_st_wait(); }
  }
  const long long t1 = clock64();
  float a0, a1; unpack2(add2(la, lb), a0, a1);
  out[threadIdx.x] = a0 + a1 + __uint_as_float(sink & 0x7fffff);
  if (threadIdx.x % 32 == 0) cyc[warp] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

template <int POLY, int MODE>
void run(const char* name, const float* in, float* out, long long* cyc) {
  for (int threads : {128, 256, 512}) {
    const int iters = 2000;
    bench<POLY, MODE><<<1, threads>>>(in, out, cyc, iters);
    bench<POLY, MODE><<<1, threads>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int w = 0; w < threads / 32; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_iter = mx / iters, wps = threads / 128.0;
    printf("%-28s warps/sched %d: %7.1f cyc per half-block iteration of all its warps, %5.2f cyc per MUFU-equivalent pair slot (XU floor 8 per MUFU: %5.1f)\n",
           name, threads / 128, per_iter, per_iter / (16 * wps), per_iter / ((32 - 4 * POLY) * wps));
  }
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 512 * 32 * 4); cudaMalloc(&out, 512 * 4); cudaMalloc(&cyc, 16 * 8);
  float h[512 * 32];
  for (int i = 0; i < 512 * 32; ++i) h[i] = -(float)(i % 97) * 0.37f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 0>("mufu, no tmem st", in, out, cyc);
  run<0, 1>("mufu, tmem st + wait", in, out, cyc);
  run<0, 2>("mufu + ffma2 only", in, out, cyc);
  run<2, 0>("poly 2/8, no tmem st", in, out, cyc);
  run<2, 1>("poly 2/8, tmem st + wait", in, out, cyc);
  run<4, 0>("poly 4/8, no tmem st", in, out, cyc);
  run<4, 1>("poly 4/8, tmem st + wait", in, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
