// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M=128) vs N, operands in smem.
// One CTA per SM, one thread issues `iters` groups of `per_commit` MMAs followed by a commit+wait.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../../multi-pass-gan_b200/csrc/ptx.cuh"
using namespace mpg;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int per_commit, int a_rows_shift, long long* cycles, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16kind(128, n, 0u);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < per_commit; ++j) {
        const uint64_t ad = umma_smem_desc(a0 + (j & 3) * 32 + ((j >> 2) & 1) * a_rows_shift, 1024, 2);
        const uint64_t bd = umma_smem_desc(b0 + (j & 3) * 32, 1024, 2);
        if (mode == 0) umma_bf16_ss(tb + ((j >> 2) & 1) * 128, ad, bd, idesc, 1u);
        else umma_ts(tb + ((j >> 2) & 1) * 128, tb + 256 + (j & 3) * 8, bd, idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int ns[] = {16, 32, 64, 128};
  for (int mode : {0, 1}) for (int grid : {148}) for (int n : ns) for (int pc : {64}) {
    int iters = 20000 / pc * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, 128, 100 * 1024>>>(n, 100, pc, 2048, d, mode);  // warm
    cudaEventRecord(e0);
    k<<<grid, 128, 100 * 1024>>>(n, iters, pc, 2048, d, mode);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    double mmas = (double)iters * pc;
    printf("mode %s grid %3d N %3d per_commit %2d : %.1f ns/MMA  %.1f clk/MMA  -> %.0f TFLOP/s chip-equivalent (x148)\n", mode ? "TS(A in TMEM)" : "SS", grid, n, pc,
           ms * 1e6 / mmas, (double)c / mmas, 2.0 * 128 * n * 16 * 148 / (ms * 1e6 / mmas) / 1e3);
  }
  return 0;
}
