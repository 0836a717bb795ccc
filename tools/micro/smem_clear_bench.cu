// Microbenchmark (run on the GPU box): does zero-filling a warp's shared-memory accumulators with a TMA bulk copy
// (cp.async.bulk.shared::cluster.global, async proxy) relieve the LSU pipe compared with STS.128 stores?
// Each warp owns 3072 fp32 accumulators (as bm25_sweep_kernel does), runs PASSES read-modify-write passes of 4 random
// offsets per lane, then clears.  Variant 0: STS.128 clear.  Variant 1: bulk copy of zeros from global + mbarrier.
// Variant 2: no clear at all (lower bound).   nvcc -arch=sm_100a -O3 -o smem_clear_bench smem_clear_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int SLICE = 3072;
constexpr int WARPS = 8;
constexpr int WBYTES = SLICE * 4 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void __launch_bounds__(256, 2) bench(const float* __restrict__ zeros, int iters, int passes, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* acc = (float*)(sm + (size_t)w * WBYTES);
  uint64_t* bar = (uint64_t*)(sm + (size_t)w * WBYTES + SLICE * 4);
  for (int j = lane * 4; j < SLICE; j += 128) *reinterpret_cast<float4*>(acc + j) = make_float4(0, 0, 0, 0);
  if (VARIANT == 1 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t rng = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
  float sum = 0.f;
  uint32_t phase = 0;
  for (int it = 0; it < iters; ++it) {
    for (int p = 0; p < passes; ++p) {
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        rng = rng * 1664525u + 1013904223u;
        o[e] = (rng >> 8) % SLICE;
      }
      __syncwarp();
      float x0 = acc[o[0]], x1 = acc[o[1]], x2 = acc[o[2]], x3 = acc[o[3]];
      x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
      acc[o[0]] = x0; acc[o[1]] = x1; acc[o[2]] = x2; acc[o[3]] = x3;
      sum += x0;
    }
    __syncwarp();
    if (VARIANT == 0) {
#pragma unroll
      for (int j = lane * 4; j < SLICE; j += 128) *reinterpret_cast<float4*>(acc + j) = make_float4(0, 0, 0, 0);
    } else if (VARIANT == 1) {
      // generic-proxy accesses of acc are ordered before the async-proxy write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(SLICE * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(acc)), "l"(zeros), "r"(SLICE * 4), "r"(smem_u32(bar)) : "memory");
      }
      // (the real kernel would do the next slice's bookkeeping and global loads here)
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
    __syncwarp();
  }
  if (sum == 123456.f) out[0] = sum;
}

int main() {
  float* zeros; float* out;
  cudaMalloc(&zeros, SLICE * 4); cudaMemset(zeros, 0, SLICE * 4); cudaMalloc(&out, 4);
  const int smem = WARPS * WBYTES;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int passes : {6, 12, 24}) {
    for (int v = 0; v < 3; ++v) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (v == 0) bench<0><<<296, 256, smem>>>(zeros, iters, passes, out);
        if (v == 1) bench<1><<<296, 256, smem>>>(zeros, iters, passes, out);
        if (v == 2) bench<2><<<296, 256, smem>>>(zeros, iters, passes, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
      }
      cudaError_t err = cudaGetLastError();
      printf("passes %2d variant %d (%s): %.3f ms  = %.1f ns per (warp, slice)  %s\n", passes, v,
             v == 0 ? "STS.128 clear" : v == 1 ? "TMA bulk zero fill" : "no clear", best, best * 1e6 / iters,
             err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
