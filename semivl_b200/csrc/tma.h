// Host-side helpers shared by the kernels' launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace svl {
// Encode a tiled bf16 tensor map with SWIZZLE_128B and zero out-of-bounds fill.
// dims[0] is the contiguous dimension; strides_bytes has rank-1 entries (dimension 1..rank-1).
int tma_encode_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
// general form: dtype = SVL_F32 / SVL_BF16, swizzle_bytes = 0 / 32 / 64 / 128
int tma_encode(CUtensorMap* map, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box);
int num_sms();
// tcgen05 flash attention (attention_tc.cu), bf16 throughput mode only
size_t attention_bwd_tc_workspace(int b, int L, int heads);
int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* ws, const void* dv_add, int dv_add_dtype,
                     int64_t ld_dv_add, void* dqkv, int b, int L, int heads, float scale, cudaStream_t stream);
int attention_fwd_tc(const void* qkv, void* out, float* lse, int b, int L, int heads, float scale, cudaStream_t stream);
}  // namespace svl
