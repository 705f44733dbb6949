// Microbenchmark: tcgen05.ld throughput per SM (4 and 8 warps), FFMA vs FFMA2 (fma.rn.f32x2) throughput.
#include "../../chunkformer_b200/csrc/common.cuh"
#include <cstdio>
using namespace cf;

template <int NW, int SHAPE>
__global__ void tmem_ld_kernel(float* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t taddr = base + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (SHAPE == 32) {
        uint32_t r[32];
        tmem_ld32(taddr + ((c * 32 + (warp >> 2) * 256) & 511), r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 8) acc += __uint_as_float(r[k]);
      } else {
        uint32_t r[16];
        tmem_ld16(taddr + ((c * 16 + (warp >> 2) * 256) & 511), r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; k += 8) acc += __uint_as_float(r[k]);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

// two loads in flight before the wait
template <int NW>
__global__ void tmem_ld2_kernel(float* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t taddr = base + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32], s[32];
      tmem_ld32(taddr + ((c * 64) & 511), r);
      tmem_ld32(taddr + ((c * 64 + 32) & 511), s);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; k += 8) acc += __uint_as_float(r[k]) + __uint_as_float(s[k]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

template <int MODE>
__global__ void fma_kernel(float* out, long long* clk, int iters, float w) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float b = w, c = w * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long x, y, z;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c), "f"(c));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(y), "l"(z));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
      }
    } else if (MODE == 2) {   // fma with |a|
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(fabsf(a[i]), b, c);
    } else if (MODE == 3) {   // fma + max interleaved
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(fmaxf(a[i], 0.f), b, c);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&clk, 8 * 1024);
  long long h[8];
  const int iters = 2000;
#define RUN(name, kern, threads, bytes_per_iter)                                           \
  kern<<<1, threads>>>(out, clk, iters);                                                    \
  cudaDeviceSynchronize();                                                                  \
  cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);                                            \
  printf("%-28s clk %lld  -> %.1f B/clk/SM  (%s)\n", name, h[0], double(bytes_per_iter) * iters / h[0], cudaGetErrorString(cudaGetLastError()));
  RUN("ld x32, 4 warps", (tmem_ld_kernel<4, 32>), 128, 4 * 8 * 32 * 32 * 4);
  RUN("ld x32, 8 warps", (tmem_ld_kernel<8, 32>), 256, 8 * 8 * 32 * 32 * 4);
  RUN("ld x32, 16 warps", (tmem_ld_kernel<16, 32>), 512, 16 * 8 * 32 * 32 * 4);
  RUN("ld x16, 4 warps", (tmem_ld_kernel<4, 16>), 128, 4 * 8 * 16 * 32 * 4);
  RUN("ld x16, 8 warps", (tmem_ld_kernel<8, 16>), 256, 8 * 8 * 16 * 32 * 4);
  RUN("ld 2x32 inflight, 4 warps", (tmem_ld2_kernel<4>), 128, 4 * 4 * 64 * 32 * 4);
  RUN("ld 2x32 inflight, 8 warps", (tmem_ld2_kernel<8>), 256, 8 * 4 * 64 * 32 * 4);
#define RUNF(name, mode, threads)                                                          \
  fma_kernel<mode><<<1, threads>>>(out, clk, iters, 1.0001f);                               \
  cudaDeviceSynchronize();                                                                  \
  cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);                                            \
  printf("%-28s clk %lld  -> %.1f FMA/clk/SM (%s)\n", name, h[0], double(threads) * 16 * iters / h[0], cudaGetErrorString(cudaGetLastError()));
  RUNF("ffma 8 warps", 0, 256); RUNF("ffma 16 warps", 0, 512); RUNF("ffma 32 warps", 0, 1024);
  RUNF("ffma2 8 warps", 1, 256); RUNF("ffma2 16 warps", 1, 512); RUNF("ffma2 32 warps", 1, 1024);
  RUNF("ffma |a| 16 warps", 2, 512); RUNF("ffma+max 16 warps", 3, 512); RUNF("ffma+max 32 warps", 3, 1024);
  return 0;
}
