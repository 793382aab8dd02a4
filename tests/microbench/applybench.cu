// Streaming-pass experiment (not part of the library): the forward apply pass's memory pattern -- read bf16 z
// [1024][B], scale/shift per feature, ReLU, write bf16 h [1024][B] and 1 keep bit per element -- at different
// occupancies, to tell whether the persistent kernel's 8 warps per SM or the pattern itself caps the pass at
// ~2.5 TB/s.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a applybench.cu -o applybench && ./applybench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pk(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// one warp per (feature, 2048-entry group): lane owns 8 consecutive entries of each of U 256-entry chunks
template <int U>
__global__ void apply_kernel(const uint16_t* __restrict__ z, uint16_t* __restrict__ h, uint32_t* __restrict__ keep,
                             const float2* __restrict__ ss, int B, int warps_total) {
  const int lane = threadIdx.x & 31;
  const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int per = 256 * U;
  const int ngrp = B / per;
  const int items = 1024 * ngrp;
  for (int it = warp_g; it < items; it += warps_total) {
    const int f = it / ngrp, grp = it - f * ngrp;
    const size_t base = size_t(f) * B + size_t(grp) * per + 8 * lane;
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = __ldcg(reinterpret_cast<const uint4*>(z + base + 256 * u));
    const float2 k = __ldg(ss + f);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t w[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
      uint32_t o[4], bits = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = fmaf(lo(w[j]), k.x, k.y), b = fmaf(hi(w[j]), k.x, k.y);
        const bool ka = a > 0.f && ((it + j) & 1), kb = b > 0.f && ((it >> 1) & 1);
        bits |= (ka ? 1u : 0u) << (2 * j) | (kb ? 1u : 0u) << (2 * j + 1);
        o[j] = pk(ka ? 2.f * a : 0.f, kb ? 2.f * b : 0.f);
      }
      *reinterpret_cast<uint4*>(h + base + 256 * u) = make_uint4(o[0], o[1], o[2], o[3]);
      uint32_t kw = bits << (8 * (lane & 3));
      kw |= __shfl_xor_sync(0xffffffffu, kw, 1);
      kw |= __shfl_xor_sync(0xffffffffu, kw, 2);
      if ((lane & 3) == 0) keep[(base + 256 * u) >> 5] = kw;
    }
  }
}

int main() {
  const int B = 65536;
  const size_t n = size_t(1024) * B;
  uint16_t *z, *h; uint32_t* keep; float2* ss; float* flush;
  cudaMalloc(&z, n * 2); cudaMalloc(&h, n * 2); cudaMalloc(&keep, n / 8); cudaMalloc(&ss, 1024 * 8);
  cudaMalloc(&flush, 256 << 20);
  cudaMemset(z, 0x3f, n * 2); cudaMemset(ss, 0, 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = double(n) * 2 * 2 + double(n) / 8;
  auto run = [&](const char* name, int grid, int block, int u) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemsetAsync(flush, rep, 256 << 20);
      cudaEventRecord(e0);
      const int wt = grid * block / 32;
      if (u == 8) apply_kernel<8><<<grid, block>>>(z, h, keep, ss, B, wt);
      else if (u == 4) apply_kernel<4><<<grid, block>>>(z, h, keep, ss, B, wt);
      else if (u == 2) apply_kernel<2><<<grid, block>>>(z, h, keep, ss, B, wt);
      else apply_kernel<1><<<grid, block>>>(z, h, keep, ss, B, wt);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("%-44s grid %5d x %4d U=%d : %7.1f us  %6.0f GB/s  (%s)\n", name, grid, block, u, best * 1e3, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  };
  run("persistent, 8 warps/SM (megakernel-like)", 148, 256, 8);
  run("persistent, 8 warps/SM", 148, 256, 4);
  run("persistent, 16 warps/SM", 148, 512, 8);
  run("persistent, 16 warps/SM", 148, 512, 2);
  run("persistent, 32 warps/SM", 148, 1024, 8);
  run("persistent, 32 warps/SM", 148, 1024, 2);
  run("persistent, 64 warps/SM (2 x 1024)", 296, 1024, 2);
  run("persistent, 64 warps/SM (2 x 1024)", 296, 1024, 1);
  run("one item per warp, 64 warps/SM", 1024 * 32 / 8, 256, 8);
  run("one item per warp (U=2)", 1024 * 128 / 8, 256, 2);
  run("one item per warp (U=1)", 1024 * 256 / 8, 256, 1);
  // plain device-to-device copy of the same bytes for reference
  {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemsetAsync(flush, rep, 256 << 20);
      cudaEventRecord(e0);
      cudaMemcpyAsync(h, z, n * 2, cudaMemcpyDeviceToDevice);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("cudaMemcpy D2D 128 MB: %7.1f us  %6.0f GB/s (read + write)\n", best * 1e3, double(n) * 4 / best / 1e6);
  }
  return 0;
}
