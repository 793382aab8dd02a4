// microbenchmark: TMA operand delivery rate per SM for the forward GEMM's operand shapes (no MMA, no epilogue)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../bilinear_b200/csrc/ptx.cuh"
using namespace b200;

struct P { CUtensorMap a, b, b4, a64, a32; int reps; };
// mode 0: A box {64k,128m} + 4 B boxes {64n,64k}   (what the kernel does)
// mode 1: A only        mode 2: B only (4 boxes)    mode 3: A + one 4-D B box {64n,64k,4}
// mode 4: B only as one 4-D box
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p, int kblocks, int mode, int stages, long long* out, int n_tiles_total) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  const uint32_t bytesA = 16384, bytesB = 32768;
  const uint32_t tx = ((mode == 0 || mode == 1 || mode == 3 || mode >= 5) ? bytesA : 0) + ((mode != 1 && mode != 5 && mode != 6) ? bytesB : 0);
  const uint32_t stage_bytes = 49152;
  if (warp == 0) {
    const uint32_t el = elect_one();
    long long t0 = clock64();
    int it = 0, sg = 0; uint32_t par = 1;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
      const int mt = tile & 7, nt = tile >> 3;
      for (int kb = 0; kb < kblocks; ++kb, ++it, sg = (sg + 1 == stages) ? 0 : sg + 1, par ^= (sg == 0)) {
        mbar_wait(&empty[sg], par, 1);
        if (el) {
          uint8_t* st = smem + sg * stage_bytes;
          mbar_expect_tx(&full[sg], tx);
          if (mode == 0 || mode == 1 || mode == 3) tma_load_3d(st, &p.a, &full[sg], kb * 64, mt * 128, blockIdx.x % p.reps);
          if (mode == 5 || mode == 7) for (int i = 0; i < 2; ++i) tma_load_3d(st + i * 8192, &p.a64, &full[sg], kb * 64, mt * 128 + 64 * i, 0);
          if (mode == 6 || mode == 8) for (int i = 0; i < 4; ++i) tma_load_3d(st + i * 4096, &p.a32, &full[sg], kb * 64, mt * 128 + 32 * i, 0);
          if (mode == 0 || mode == 2 || mode == 7 || mode == 8)
            for (int i = 0; i < 4; ++i) tma_load_3d(st + 16384 + i * 8192, &p.b, &full[sg], nt * 256 + 64 * i, kb * 64, 0);
          if (mode == 3 || mode == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(st + 16384)), "l"(&p.b4), "r"(smem_u32(&full[sg])), "r"(0), "r"(kb * 64), "r"(nt * 4), "r"(0) : "memory");
        }
      }
    }
    long long t1 = clock64();
    if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = it; }
  } else if (warp == 1) {
    int it = 0, sg = 0; uint32_t par = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x)
      for (int kb = 0; kb < kblocks; ++kb, ++it, sg = (sg + 1 == stages) ? 0 : sg + 1, par ^= (sg == 0)) {
        mbar_wait(&full[sg], par, 2);
        if (lane == 0) mbar_arrive(&empty[sg]);
        __syncwarp();
      }
  }
}
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMapL2promotion g_prom = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
int run(int prom);
int main() { for (int pr : {2}) { g_prom = (CUtensorMapL2promotion)pr; printf("== L2 promotion enum %d\n", pr); if (run(pr)) return 1; } return 0; }
int run(int prom) {
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  EncFn enc = (EncFn)fn;
  const int64_t M = 1024, K = 1024, N = 65536;
  uint16_t *A, *B; cudaMalloc(&A, 148 * M * K * 2); cudaMalloc(&B, K * N * 2);
  cudaMemset(A, 0x3c, 148 * M * K * 2); cudaMemset(B, 0x3c, K * N * 2);
  P p; memset(&p, 0, sizeof p);
  { cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(M), 148}; cuuint64_t str[2] = {cuuint64_t(K) * 2, cuuint64_t(M * K) * 2}; cuuint32_t box[3] = {64, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, g_prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc a %d\n", r); return 1; } }
  { cuuint64_t dims[3] = {cuuint64_t(N), cuuint64_t(K), 1}; cuuint64_t str[2] = {cuuint64_t(N) * 2, cuuint64_t(N * K) * 2}; cuuint32_t box[3] = {64, 64, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, B, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, g_prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc b %d\n", r); return 1; } }
  { // [K][N] viewed as {n_in 64, k K, n_out N/64, 1}: strides k: N*2, n_out: 128 B
    cuuint64_t dims[4] = {64, cuuint64_t(K), cuuint64_t(N / 64), 1}; cuuint64_t str[3] = {cuuint64_t(N) * 2, 128, cuuint64_t(N * K) * 2}; cuuint32_t box[4] = {64, 64, 4, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.b4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, B, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, g_prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc b4 %d\n", r); return 1; } }
  { cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(M), 148}; cuuint64_t str[2] = {cuuint64_t(K) * 2, cuuint64_t(M * K) * 2}; cuuint32_t box[3] = {64, 64, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.a64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, g_prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc a64 %d\n", r); return 1; } }
  { cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(M), 148}; cuuint64_t str[2] = {cuuint64_t(K) * 2, cuuint64_t(M * K) * 2}; cuuint32_t box[3] = {64, 32, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.a32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, g_prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc a32 %d\n", r); return 1; } }
  long long* d; cudaMalloc(&d, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const char* names[9] = {"A + 4 B boxes", "A only", "4 B boxes", "A + one 4-D B box", "one 4-D B box", "A as 2 boxes", "A as 4 boxes", "A(2 boxes) + 4 B", "A(4 boxes) + 4 B"};
  for (int reps : {1})
  for (int stages : {4})
  for (int mode : {1, 5, 6, 0, 7, 8}) {
    float ms = 0; long long h[296]; p.reps = reps;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k<<<148, 128, 220 * 1024>>>(p, 16, mode, stages, d, 2048);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    }
    double cyc = 0, its = 0; for (int i = 0; i < 148; ++i) { cyc += h[2 * i]; its += h[2 * i + 1]; }
    printf("reps %3d stages %d  %-20s: %.1f us, %.0f cycles per k-block per SM\n", reps, stages, names[mode], ms * 1e3, cyc / its);
  }
  return 0;
}
