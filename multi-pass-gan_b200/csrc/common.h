// Shared host-side plumbing of libmpg_b200: handle, error reporting, driver entry points.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mpg.h"

namespace mpg {

void set_error(const char* fmt, ...);

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace mpg

struct mpg_handle_s {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  mpg::PFN_encodeTiled encode_tiled;
};

#define MPG_CHECK_ARG(cond, ...)  \
  do {                            \
    if (!(cond)) {                \
      mpg::set_error(__VA_ARGS__); \
      return MPG_EINVAL;          \
    }                             \
  } while (0)

#define MPG_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      mpg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                   \
                     cudaGetErrorString(e__));                                     \
      return static_cast<int>(e__);                                                \
    }                                                                              \
  } while (0)

namespace mpg {

// 16-bit activation element helpers; dtype is MPG_BF16 or MPG_F16 (both 2 bytes)
__device__ __forceinline__ float h16_to_float(uint16_t raw, int dtype) {
  if (dtype == MPG_F16) return __half2float(__ushort_as_half(raw));
  return __uint_as_float(static_cast<uint32_t>(raw) << 16);
}
__device__ __forceinline__ uint16_t float_to_h16(float v, int dtype) {
  if (dtype == MPG_F16) {
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);  // saturate instead of producing inf
    return __half_as_ushort(__float2half_rn(v));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
// two floats -> packed 16-bit pair, round-to-nearest-even, saturating at the largest finite value: ONE F2FP.SATFINITE
// instruction (the clamp + convert + permute sequence it replaces was 6 instructions per pair in every epilogue)
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi, int dtype) {
  uint32_t r;
  if (dtype == MPG_F16)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
inline bool is_h16(int dtype) { return dtype == MPG_BF16 || dtype == MPG_F16; }
inline bool is_dtype(int dtype) { return dtype == MPG_BF16 || dtype == MPG_F16 || dtype == MPG_F32; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Makes `device` current for the scope and restores the caller's device afterwards: every entry point that launches
// on a handle's device may be called while another device is current (one process can hold handles on several GPUs).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
constexpr int kMaxDevices = 64;

// Encode a tiled tensor map; returns 0 or a CUresult.
int encode_tmap(mpg_handle h, CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */,
                const uint32_t* box, CUtensorMapSwizzle swz);

}  // namespace mpg
