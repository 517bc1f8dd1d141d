// Handle lifecycle, error string, tensor-map encoding helper.
#include <string.h>

#include "common.h"

namespace mpg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int encode_tmap(mpg_handle h, CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swz) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = h->encode_tiled(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                               gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)",
              static_cast<int>(r), rank, (unsigned long long)dims[0],
              (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
              rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return static_cast<int>(r);
  }
  return 0;
}

}  // namespace mpg

extern "C" {

int mpg_version(void) { return 100; }

const char* mpg_last_error(void) { return mpg::g_err; }

int mpg_create(mpg_handle* out, int device) {
  MPG_CHECK_ARG(out != nullptr, "mpg_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    mpg::set_error("mpg_create: no CUDA device visible (%s); this library has no CPU fallback",
                   cudaGetErrorString(e));
    return e != cudaSuccess ? static_cast<int>(e) : MPG_EINVAL;
  }
  MPG_CHECK_ARG(device >= 0 && device < ndev, "mpg_create: device %d out of range [0,%d)", device,
                ndev);
  MPG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MPG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    mpg::set_error("mpg_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                   prop.major, prop.minor);
    return MPG_ENOSUP;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    mpg::set_error("mpg_create: cuTensorMapEncodeTiled entry point unavailable");
    return MPG_EDRIVER;
  }
  mpg_handle h = new mpg_handle_s();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->encode_tiled = reinterpret_cast<mpg::PFN_encodeTiled>(fn);
  *out = h;
  return MPG_OK;
}

int mpg_destroy(mpg_handle h) {
  delete h;
  return MPG_OK;
}

int mpg_sm_count(mpg_handle h) { return h ? h->sm_count : MPG_EINVAL; }

}  // extern "C"
