// Host-side helpers shared by the kernels' launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svl {
// Encode a tiled bf16 tensor map with SWIZZLE_128B and zero out-of-bounds fill.
// dims[0] is the contiguous dimension; strides_bytes has rank-1 entries (dimension 1..rank-1).
int tma_encode_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
int num_sms();
// tcgen05 flash attention (attention_tc.cu), bf16 throughput mode only
int attention_fwd_tc(const void* qkv, void* out, float* lse, int b, int L, int heads, float scale, cudaStream_t stream);
}  // namespace svl
