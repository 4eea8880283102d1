// Flash attention forward on tcgen05 (throughput mode, bf16 operands): S = Q K^T and O_tile = P V run on the 5th-gen tensor
// cores with accumulators in TMEM; the softmax warps own one query row per thread (no cross-thread reductions), read S with
// tcgen05.ld, write P as a K-major SWIZZLE_128B operand into shared memory and fold each O_tile into fp32 registers.
//
//   warp 0      : TMA producer (Q once; K/V tiles of 128 keys, 2-stage ring; 3-D tensor map -> rows >= L are zero-filled)
//   warp 1      : MMA issuer  (S_{j+1} is issued before P_j V_j so the tensor pipe works while the softmax of tile j runs)
//   warps 2..5  : softmax / output (thread = query row = TMEM lane)
//
// TMEM: S[2] (2 x 128 columns) | O_tile[2] (2 x 64 columns).  Shared memory: Q 16 KB, K/V 2 x 32 KB, P 2 x 32 KB.
// replaces the nn.MultiheadAttention core (maskclip_vit.py:77-84,141) for head_dim 64; the split-bf16 precise mode keeps the
// mma.sync kernels of attention.cu.
#include "common.cuh"
#include "ptx.cuh"
#include "tma.h"

namespace svl {
namespace {

constexpr int TQ = 128, TK = 128, D = 64;
constexpr int kThreads = 192;
constexpr uint32_t kTile = 128 * 128;        // bytes of a 128-row x 64-bf16 tile
constexpr float kLog2e = 1.4426950408889634f;

struct FwdParams {
  __nv_bfloat16* out; int64_t ldo;
  float* lse;
  int L, heads;
  float scale;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 64 TMEM columns of this thread's lane in one round trip (two x32 loads, one wait)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* v) {
  ptx::tmem_ld_32x32(taddr, v);
  ptx::tmem_ld_32x32(taddr + 32u, v + 32);
  ptx::tmem_ld_wait();
}

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ FwdParams p) {
  // Two CTAs per SM (112 KB shared memory, 256 TMEM columns each): while one CTA's softmax warps work, the other's MMAs run.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  if ((raw & 1023u) != 0) __trap();                      // SWIZZLE_128B tiles need 1024-byte alignment; the launch relies on it
  const uint32_t sQ = raw, sK = raw + kTile, sV = sK + 2 * kTile, sP = sV + 2 * kTile;     // P: 2 k-blocks x 16 KB
  const uint32_t bar = sP + 2 * kTile;
  const uint32_t q_full = bar, s_full = bar + 8, s_empty = bar + 16, p_full = bar + 24;
  auto kv_full = [&](int s) { return bar + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (6 + s); };
  auto o_full = [&](int s) { return bar + 8u * (8 + s); };
  auto o_empty = [&](int s) { return bar + 8u * (10 + s); };
  const uint32_t tmem_ptr_addr = bar + 8u * 12;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int E = p.heads * D;
  const int nt = (p.L + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 4);
    ptx::mbar_init(p_full, 4);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(kv_full(s), 1);
      ptx::mbar_init(kv_empty(s), 1);
      ptx::mbar_init(o_full(s), 1);
      ptx::mbar_init(o_empty(s), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_addr, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;            // S: columns [0,128), O_tile[2]: [128,256)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, kTile);
      ptx::tma_load_3d(sQ, &tmQKV, q_full, h * D, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1, k = j >> 1;
        ptx::mbar_wait(kv_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(kv_full(st), 2 * kTile);
        ptx::tma_load_3d(sK + st * kTile, &tmQKV, kv_full(st), E + h * D, j * TK, b);
        ptx::tma_load_3d(sV + st * kTile, &tmQKV, kv_full(st), 2 * E + h * D, j * TK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);
      const uint64_t tmpl = ptx::make_smem_desc(0, 8192, 1024);
      const uint64_t qd = tmpl + (uint64_t)(sQ >> 4), pa = tmpl + (uint64_t)(sP >> 4);
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1, k = j >> 1;
        ptx::mbar_wait(kv_full(st), (uint32_t)(k & 1));
        ptx::mbar_wait(s_empty, (uint32_t)((j & 1) ^ 1));
        ptx::tc_fence_after();
        const uint64_t kd = tmpl + (uint64_t)((sK + st * kTile) >> 4), vb = tmpl + (uint64_t)((sV + st * kTile) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tmem_base, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_s, kk > 0);
        ptx::umma_commit(s_full);
        ptx::mbar_wait(p_full, (uint32_t)(j & 1));
        ptx::mbar_wait(o_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          // A: P k-block s/4 (16 KB apart), 32 bytes per 16 keys inside the swizzled row; B: V rows (keys) are the K dimension, 16 rows = 2048 bytes
          ptx::umma_bf16(tmem_base + 128u + (uint32_t)(st * 64), pa + (uint64_t)((s >> 2) * (kTile >> 4) + (s & 3) * 2), vb + (uint64_t)(s * 128),
                         idesc_pv, s > 0);
        }
        ptx::umma_commit(o_full(st));
        ptx::umma_commit(kv_empty(st));
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    float m = -INFINITY, l = 0.f, c_prev = 1.f;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    uint8_t* prow = smem_raw + (sP - raw) + r * 128;
    auto fold_o = [&](int t) {
      const int pb = t & 1, pk = t >> 1;
      ptx::mbar_wait(o_full(pb), (uint32_t)(pk & 1));
      ptx::tc_fence_after();
      uint32_t v[64];
      tmem_ld64(tl + 128u + (uint32_t)(pb * 64), v);
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] = o[i] * c_prev + __uint_as_float(v[i]);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty(pb));
    };
    for (int j = 0; j < nt; ++j) {
      ptx::mbar_wait(s_full, (uint32_t)(j & 1));
      ptx::tc_fence_after();
      const int key0 = j * TK;
      const bool tail = key0 + TK > p.L;
      // pass A: row maximum of the raw scores (the scale is positive, so it is applied to the maximum only)
      float mr = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[64];
        tmem_ld64(tl + (uint32_t)(c * 64), v);
        if (tail) {
#pragma unroll
          for (int i = 0; i < 64; ++i) if (key0 + c * 64 + i < p.L) mr = fmaxf(mr, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) mr = fmaxf(mr, __uint_as_float(v[i]));
        }
      }
      const float mx = fmaxf(m, mr * sl2);
      const float corr = ex2(m - mx);                 // first tile: ex2(-inf) = 0
      // the previous tile's P V product must have left the P buffer (and its O_tile is folded in while we are at it)
      if (j > 0) fold_o(j - 1);
      // pass B: P = exp2(s * scale - mx) -> bf16 operand tile in shared memory (K-major, 128-byte rows, 16-byte chunks XOR-swizzled by row)
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[64];
        tmem_ld64(tl + (uint32_t)(c * 64), v);
        float pv[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) pv[i] = ex2(fmaf(__uint_as_float(v[i]), sl2, -mx));
        if (tail) {
#pragma unroll
          for (int i = 0; i < 64; ++i) if (key0 + c * 64 + i >= p.L) pv[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) sum += pv[i];
#pragma unroll
        for (int piece = 0; piece < 8; ++piece) *(uint4*)(prow + c * kTile + ((piece ^ (r & 7)) << 4)) = f32_to_bf16x8(pv + piece * 8);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
      l = l * corr + sum;
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      c_prev = corr;
      m = mx;
    }
    fold_o(nt - 1);
    const int row = q0 + r;
    if (row < p.L) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = p.out + ((int64_t)b * p.L + row) * p.ldo + h * D;
#pragma unroll
      for (int i = 0; i < D; ++i) o[i] *= inv;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) *(uint4*)(dst + i * 8) = f32_to_bf16x8(o + i * 8);
      if (p.lse) p.lse[((int64_t)b * p.heads + h) * p.L + row] = (m + log2f(l)) / kLog2e;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int attention_fwd_tc(const void* qkv, void* out, float* lse, int b, int L, int heads, float scale, cudaStream_t stream) {
  const int E = heads * D;
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)3 * E, (uint64_t)L, (uint64_t)b};
  uint64_t strides[2] = {(uint64_t)3 * E * 2, (uint64_t)L * 3 * E * 2};
  uint32_t box[3] = {64u, 128u, 1u};
  if (int rc = tma_encode_bf16(&tm, qkv, 3, dims, strides, box)) return rc;
  FwdParams p;
  p.out = (__nv_bfloat16*)out; p.ldo = E; p.lse = lse; p.L = L; p.heads = heads; p.scale = scale;
  const size_t smem = 7 * kTile + 8 * 13 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SVL_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((L + TQ - 1) / TQ, heads, b);
  attn_fwd_tc_kernel<<<grid, kThreads, smem, stream>>>(tm, p);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

}  // namespace svl
