// Error state, device check and the TMA descriptor encoder (driver entry point resolved at run time, no -lcuda).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "tma.h"

namespace svl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_sms = 0;
static int g_checked = 0;   // 0 = not yet, 1 = ok, <0 = error code

int num_sms() { return g_sms > 0 ? g_sms : 148; }

int tma_encode(CUtensorMap* map, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
  if (!g_encode) {
    set_error("tma_encode: driver entry point not resolved (svl_check_device not called?)");
    return SVL_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(map, dtype == SVL_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u stride0 %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return SVL_ERR_CUDA;
  }
  return SVL_OK;
}

int tma_encode_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return tma_encode(map, base, SVL_BF16, 128, rank, dims, strides_bytes, box);
}

}  // namespace svl

using namespace svl;

extern "C" const char* svl_last_error(void) { return g_err; }
extern "C" int svl_version(void) { return 100; }

extern "C" int svl_check_device(void) {
  if (g_checked == 1) return SVL_OK;
  int dev = 0;
  SVL_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SVL_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("semivl_b200 needs an sm_100 (B200) device, found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
    return SVL_ERR_ARCH;
  }
  g_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SVL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return SVL_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  g_checked = 1;
  return SVL_OK;
}
