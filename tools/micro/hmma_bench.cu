// Micro-benchmark: throughput of the legacy warp-level mma.sync.m16n8k16 (f16 x f16 -> f32) on sm_100a, per SM,
// as a function of warps per SM and independent accumulator chains per warp. Decides whether the thin head/tail
// layers of gen_resnet (3 % of the FLOPs) can run on register-level MMAs inside fused kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_bench hmma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(1024) k(int iters, float* out) {
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b[2] = {0x3c003c00u + threadIdx.x, 0x3c003c00u};
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.f) out[0] = s;
}

template <int CH>
void run(int warps) {
  float* d; cudaMalloc(&d, 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<148, warps * 32>>>(100, d);
  cudaEventRecord(e0);
  k<CH><<<148, warps * 32>>>(iters, d);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double macs = 148.0 * warps * (double)iters * CH * 16 * 8 * 16;
  printf("warps/SM %2d chains %d : %.3f ms  %.1f TFLOP/s  (%.0f MAC/ns/SM)\n", warps, CH, ms, 2 * macs / (ms * 1e-3) / 1e12,
         macs / 148 / (ms * 1e6));
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
