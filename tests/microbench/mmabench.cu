// microbenchmark: cycles per tcgen05.mma (M=128, N=256, K=16, bf16) for K-major / MN-major operands
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../bilinear_b200/csrc/ptx.cuh"
using namespace b200;
__global__ void __launch_bounds__(256, 1) k(int iters, int a_mn, int b_mn, int n, long long* out, int mode, int rnd) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4], full[4], empty[4], never;
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 * 4 / 4; i += 256) { uint32_t h = (i + 1) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; reinterpret_cast<uint32_t*>(smem)[i] = rnd ? ((h & 0x807f807fu) | 0x3f003f00u) : 0x3f803f80u; }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); for (int i = 0; i < 4; ++i) { mbar_init(&dummy[i], 1); mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) { tmem_alloc<1>(&tb, 512); tmem_relinquish<1>(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_all();
  const uint32_t tmem = tb;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, n, a_mn, b_mn);
    const uint32_t a_lbo = a_mn ? 8192u : 16u, b_lbo = b_mn ? 8192u : 16u;
    const uint32_t a_kstep = a_mn ? 2048u : 32u, b_kstep = b_mn ? 2048u : 32u;
    long long t0 = clock64();
    if (mode == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t st = smem_u32(smem) + (it & 3) * 49152;
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc_sw128(st + kk * a_kstep, a_lbo, 1024u);
            const uint64_t db = umma_desc_sw128(st + 16384 + kk * b_kstep, b_lbo, 1024u);
            umma_bf16<1>(tmem, da, db, idesc, (it | kk) ? 1u : 0u);
          }
        }
        umma_commit<1>(&bar);
      }
    } else if (mode == 2) {
      // commit after every k-block (4 MMAs) to a barrier nobody waits on
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t st = smem_u32(smem) + (it & 3) * 49152;
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc_sw128(st + kk * a_kstep, a_lbo, 1024u);
            const uint64_t db = umma_desc_sw128(st + 16384 + kk * b_kstep, b_lbo, 1024u);
            umma_bf16<1>(tmem, da, db, idesc, (it | kk) ? 1u : 0u);
          }
          umma_commit<1>(&dummy[it & 3]);
        }
        umma_commit<1>(&bar);
      }
    } else if (mode >= 3 && mode <= 5) {
      // full ring protocol against a producer thread (warp 2) that only arrives
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
          const int sg = it & 3;
          mbar_wait(&full[sg], (it >> 2) & 1, 2);
          tc_fence_after();
          const uint32_t st = smem_u32(smem) + sg * 49152;
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc_sw128(st + kk * a_kstep, a_lbo, 1024u);
            const uint64_t db = umma_desc_sw128(st + 16384 + kk * b_kstep, b_lbo, 1024u);
            umma_bf16<1>(tmem, da, db, idesc, (it | kk) ? 1u : 0u);
          }
          umma_commit<1>(&empty[sg]);
        }
        umma_commit<1>(&bar);
        mbar_wait(&bar, 0, 1);
        mbar_arrive(&never);
      }
    } else {
      // warp-uniform flow, one elected lane issues
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = smem_u32(smem) + (it & 3) * 49152;
        const uint64_t da0 = umma_desc_sw128(st, a_lbo, 1024u);
        const uint64_t db0 = umma_desc_sw128(st + 16384, b_lbo, 1024u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (pred) umma_bf16<1>(tmem, da0 + uint64_t((kk * a_kstep) >> 4), db0 + uint64_t((kk * b_kstep) >> 4), idesc, (it | kk) ? 1u : 0u);
        }
      }
      if (pred) umma_commit<1>(&bar);
    }
    if (mode < 3 || mode > 5) mbar_wait(&bar, 0, 1);
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  if (warp >= 4 && mode >= 4 && (mode == 4 || lane == 0)) mbar_wait(&never, 0, 9);
  if (warp == 2 && lane == 0 && mode >= 3 && mode <= 5) {
    for (int it = 0; it < iters; ++it) {
      const int sg = it & 3;
      mbar_wait(&empty[sg], ((it >> 2) & 1) ^ 1, 3);
      mbar_arrive(&full[sg]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  for (int grid : {148})
  for (int rnd : {0, 1}) for (int mode : {0, 3})
    for (int n : {256})
      for (int a_mn = 0; a_mn < 1; ++a_mn)
        for (int b_mn = 0; b_mn < 2; ++b_mn) {
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            k<<<grid, 256, 200 * 1024>>>(iters, a_mn, b_mn, n, d, mode, rnd);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          }
          printf("rnd %d grid %3d mode %d N=%3d a_mn=%d b_mn=%d: %.1f cycles / MMA (floor %d)\n", rnd, grid, mode, n, a_mn, b_mn, double(h) / (iters * 4), n / 2);
        }
  return 0;
}
