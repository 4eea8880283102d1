// Multi-head self-attention for head_dim 64, flash style (no L x L tensor in HBM), forward and backward.
// replaces the nn.MultiheadAttention core called through mmcv's wrapper (maskclip_vit.py:77-84,141)
// and the class-attention of the VLG head's SemanticTransformer (vlg_head.py:39-67, sequence = classes).
//
// qkv is packed [b, L, 3E] (q | k | v, heads contiguous inside each third); out is [b, L, E]; lse is [b, heads, L].
// Tiles of 64 queries x 64 keys per CTA (4 warps x 16 rows), K/V (or Q/dO) tiles double-buffered in XOR-swizzled
// shared memory with cp.async, fp32 online softmax, bf16 mma.sync.m16n8k16 tensor-core products.
// SPLIT = precise mode: operands are bf16 (hi | lo) pairs and every product is issued as hi*hi + hi*lo + lo*hi.
// Backward is two kernels (dK/dV per key tile, dQ per query tile): no atomics, bitwise reproducible.
#include <stdlib.h>

#include "common.cuh"
#include "tma.h"

namespace svl {
namespace {

constexpr int D = 64;          // head dim
constexpr int TQ = 64;         // rows per CTA (16 per warp)
constexpr int TK = 64;         // columns per inner tile
constexpr int kThreads = 128;
constexpr uint32_t kTileBytes = TK * D * 2;   // 8 KB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t ptx_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *(uint32_t*)&h;
}
// hi/lo split of a pair
__device__ __forceinline__ void pack_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  float2 f = __bfloat1622float2(h);
  hi = *(uint32_t*)&h;
  lo = pack_bf16(a - f.x, b - f.y);
}

// 64 x 64 bf16 tile (rows row0.., 64 columns starting at g) -> swizzled smem; rows >= nrows are zero-filled
__device__ __forceinline__ void load_tile(uint32_t sbase, const __nv_bfloat16* g, int64_t ld, int row0, int nrows) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    const int r = idx >> 3, ch = idx & 7;
    const bool ok = row0 + r < nrows;
    const __nv_bfloat16* src = g + (int64_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(tile_addr(sbase, r, ch), src, ok ? 16 : 0);
  }
}

// A fragments (16 rows x 64 k) of this warp's rows from a swizzled tile: frag[kstep][4]
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int warp_row0, uint32_t (*frag)[4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm_x4(tile_addr(sbase, warp_row0 + (lane & 15), ks * 2 + (lane >> 4)), frag[ks][0], frag[ks][1], frag[ks][2], frag[ks][3]);
}

// acc[16 x 64] += A[16 x 64(k)] * T^T  where the tile T is stored [n rows][k contiguous]  (B fragment = plain ldmatrix)
__device__ __forceinline__ void mma_nt(float (*acc)[4], const uint32_t (*a)[4], uint32_t sbase) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {          // pairs of 8-wide n tiles
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile_addr(sbase, np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
      mma16816(acc[np * 2], a[ks], b0, b1);
      mma16816(acc[np * 2 + 1], a[ks], b2, b3);
    }
  }
}
// acc[16 x 64] += A[16 x 64(k)] * T  where the tile T is stored [k rows][n contiguous]  (B fragment = ldmatrix.trans)
__device__ __forceinline__ void mma_nn(float (*acc)[4], const uint32_t (*a)[4], uint32_t sbase) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr(sbase, ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816(acc[np * 2], a[ks], b0, b1);
      mma16816(acc[np * 2 + 1], a[ks], b2, b3);
    }
  }
}

template <bool SPLIT>
__device__ __forceinline__ void mma_nt_s(float (*acc)[4], const uint32_t (*ah)[4], const uint32_t (*al)[4], uint32_t sh, uint32_t sl) {
  mma_nt(acc, ah, sh);
  if (SPLIT) {
    mma_nt(acc, ah, sl);
    mma_nt(acc, al, sh);
  }
}
template <bool SPLIT>
__device__ __forceinline__ void mma_nn_s(float (*acc)[4], const uint32_t (*ah)[4], const uint32_t (*al)[4], uint32_t sh, uint32_t sl) {
  mma_nn(acc, ah, sh);
  if (SPLIT) {
    mma_nn(acc, ah, sl);
    mma_nn(acc, al, sh);
  }
}

// C-layout accumulator [8 n-tiles][4] of a 16 x 64 block -> A fragments [4 ksteps][4] (k = the 64 columns)
template <bool SPLIT>
__device__ __forceinline__ void acc_to_a(const float (*s)[4], uint32_t (*ah)[4], uint32_t (*al)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (SPLIT) {
      pack_split(s[2 * j][0], s[2 * j][1], ah[j][0], al[j][0]);
      pack_split(s[2 * j][2], s[2 * j][3], ah[j][1], al[j][1]);
      pack_split(s[2 * j + 1][0], s[2 * j + 1][1], ah[j][2], al[j][2]);
      pack_split(s[2 * j + 1][2], s[2 * j + 1][3], ah[j][3], al[j][3]);
    } else {
      ah[j][0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
      ah[j][1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
      ah[j][2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
      ah[j][3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (*a)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
}

// write a 16 x 64 C-layout block as bf16 (optionally split) to rows row0+..., 64 columns starting at g
template <bool SPLIT>
__device__ __forceinline__ void store_acc(const float (*acc)[4], __nv_bfloat16* g, int64_t ld, int64_t lo_off, int row0, int nrows) {
  const int lane = threadIdx.x & 31;
  const int r0 = row0 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + (lane & 3) * 2;
    uint32_t h, l;
    if (r0 < nrows) {
      if (SPLIT) {
        pack_split(acc[nt][0], acc[nt][1], h, l);
        *(uint32_t*)(g + (int64_t)r0 * ld + lo_off + col) = l;
      } else h = pack_bf16(acc[nt][0], acc[nt][1]);
      *(uint32_t*)(g + (int64_t)r0 * ld + col) = h;
    }
    if (r1 < nrows) {
      if (SPLIT) {
        pack_split(acc[nt][2], acc[nt][3], h, l);
        *(uint32_t*)(g + (int64_t)r1 * ld + lo_off + col) = l;
      } else h = pack_bf16(acc[nt][2], acc[nt][3]);
      *(uint32_t*)(g + (int64_t)r1 * ld + col) = h;
    }
  }
}

// ============================================================================================ forward
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, __nv_bfloat16* __restrict__ out, int64_t ldo, float* __restrict__ lse, int L,
                int heads, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = SPLIT ? 2 : 1;                      // tiles per operand (hi, lo)
  const uint32_t s0 = ptx_smem(smem);
  // layout: stage s: K tiles [NT], V tiles [NT]
  auto sK = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + t) * kTileBytes); };
  auto sV = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + NT + t) * kTileBytes); };
  const int E = heads * D;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t lo = ld / 2;
  const __nv_bfloat16* base = qkv + (int64_t)b * L * ld + h * D;
  const __nv_bfloat16 *gq = base, *gk = base + E, *gv = base + 2 * E;

  // Q -> smem (stage 1 buffers) -> A fragments
  load_tile(sK(1, 0), gq, ld, q0, L);
  if (SPLIT) load_tile(sK(1, NT - 1), gq + lo, ld, q0, L);
  cp_async_commit();
  // first K/V tile
  load_tile(sK(0, 0), gk, ld, 0, L);
  load_tile(sV(0, 0), gv, ld, 0, L);
  if (SPLIT) {
    load_tile(sK(0, NT - 1), gk + lo, ld, 0, L);
    load_tile(sV(0, NT - 1), gv + lo, ld, 0, L);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qh[4][4], ql[4][4];
  load_a_frags(sK(1, 0), warp * 16, qh);
  if (SPLIT) load_a_frags(sK(1, NT - 1), warp * 16, ql);
  __syncthreads();

  float o[8][4];
  zero_acc(o);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sl2 = scale * kLog2e;
  const int ntiles = (L + TK - 1) / TK;
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) {
      const int k1 = (t + 1) * TK;
      load_tile(sK(st ^ 1, 0), gk, ld, k1, L);
      load_tile(sV(st ^ 1, 0), gv, ld, k1, L);
      if (SPLIT) {
        load_tile(sK(st ^ 1, NT - 1), gk + lo, ld, k1, L);
        load_tile(sV(st ^ 1, NT - 1), gv + lo, ld, k1, L);
      }
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    float s[8][4];
    zero_acc(s);
    mma_nt_s<SPLIT>(s, qh, ql, sK(st, 0), sK(st, NT - 1));
    // mask keys beyond L, running max
    const int kbase = t * TK + (lane & 3) * 2;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + nt * 8 + (e & 1);
        float v = key < L ? s[nt][e] * sl2 : -INFINITY;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);     // first tile: exp2(-inf) = 0
    m0 = mx0; m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m0); s[nt][1] = exp2f(s[nt][1] - m0);
      s[nt][2] = exp2f(s[nt][2] - m1); s[nt][3] = exp2f(s[nt][3] - m1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
      o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
    uint32_t ph[4][4], pl[4][4];
    acc_to_a<SPLIT>(s, ph, pl);
    mma_nn_s<SPLIT>(o, ph, pl, sV(st, 0), sV(st, NT - 1));
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= i0; o[nt][1] *= i0; o[nt][2] *= i1; o[nt][3] *= i1; }
  store_acc<SPLIT>(o, out + (int64_t)b * L * ldo + h * D, ldo, ldo / 2, q0 + warp * 16, L);
  if ((lane & 3) == 0 && lse) {
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    float* lp = lse + ((int64_t)b * heads + h) * L;
    if (r0 < L) lp[r0] = (m0 + log2f(l0)) / kLog2e;
    if (r1 < L) lp[r1] = (m1 + log2f(l1)) / kLog2e;
  }
}

// ============================================================================================ backward
// delta[b, h, l] = sum_d out[b, l, h*64 + d] * dout[b, l, h*64 + d]; one warp per (b, l)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, int64_t ldo, const __nv_bfloat16* __restrict__ dout, int64_t lddo,
                                  int split, float* __restrict__ delta, int64_t rows, int L, int heads) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t b = row / L, l = row % L;
    for (int h = 0; h < heads; ++h) {
      const int col = h * D + lane * 2;
      float2 a = __bfloat1622float2(*(const __nv_bfloat162*)(out + row * ldo + col));
      float2 g = __bfloat1622float2(*(const __nv_bfloat162*)(dout + row * lddo + col));
      if (split) {
        float2 a2 = __bfloat1622float2(*(const __nv_bfloat162*)(out + row * ldo + ldo / 2 + col));
        float2 g2 = __bfloat1622float2(*(const __nv_bfloat162*)(dout + row * lddo + lddo / 2 + col));
        a.x += a2.x; a.y += a2.y; g.x += g2.x; g.y += g2.y;
      }
      float s = warp_sum(a.x * g.x + a.y * g.y);
      if (lane == 0) delta[(b * heads + h) * L + l] = s;
    }
  }
}

// dK, dV for one tile of 64 keys; loops over query tiles.  Works on the transposed score tile S^T[key, query].
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads)
attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ dout, int64_t lddo,
                   const float* __restrict__ lse, const float* __restrict__ delta, const void* __restrict__ dv_add, int dv_add_dtype,
                   int64_t ld_dv_add, __nv_bfloat16* __restrict__ dqkv, int64_t ldg, int L, int heads, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = SPLIT ? 2 : 1;
  const uint32_t s0 = ptx_smem(smem);
  auto sQ = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + t) * kTileBytes); };
  auto sG = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + NT + t) * kTileBytes); };
  float* s_lse = (float*)(smem + 4 * NT * kTileBytes);       // [2][64]
  float* s_delta = s_lse + 2 * TQ;                           // [2][64]
  const int E = heads * D;
  const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t lo = ld / 2, lodo = lddo / 2;
  const __nv_bfloat16* base = qkv + (int64_t)b * L * ld + h * D;
  const __nv_bfloat16 *gq = base, *gk = base + E, *gv = base + 2 * E;
  const __nv_bfloat16* gdo = dout + (int64_t)b * L * lddo + h * D;
  const float* glse = lse + ((int64_t)b * heads + h) * L;
  const float* gdelta = delta + ((int64_t)b * heads + h) * L;

  // K, V of this tile -> A fragments (through stage-1 buffers)
  load_tile(sQ(1, 0), gk, ld, k0, L);
  load_tile(sG(1, 0), gv, ld, k0, L);
  if (SPLIT) {
    load_tile(sQ(1, NT - 1), gk + lo, ld, k0, L);
    load_tile(sG(1, NT - 1), gv + lo, ld, k0, L);
  }
  cp_async_commit();
  auto load_q_tile = [&](int st, int q0) {
    load_tile(sQ(st, 0), gq, ld, q0, L);
    load_tile(sG(st, 0), gdo, lddo, q0, L);
    if (SPLIT) {
      load_tile(sQ(st, NT - 1), gq + lo, ld, q0, L);
      load_tile(sG(st, NT - 1), gdo + lodo, lddo, q0, L);
    }
    if (threadIdx.x < TQ) {
      const int q = q0 + threadIdx.x;
      s_lse[st * TQ + threadIdx.x] = q < L ? glse[q] * kLog2e : 0.f;
      s_delta[st * TQ + threadIdx.x] = q < L ? gdelta[q] : 0.f;
    }
  };
  load_q_tile(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t kh[4][4], kl[4][4], vh[4][4], vl[4][4];
  load_a_frags(sQ(1, 0), warp * 16, kh);
  load_a_frags(sG(1, 0), warp * 16, vh);
  if (SPLIT) {
    load_a_frags(sQ(1, NT - 1), warp * 16, kl);
    load_a_frags(sG(1, NT - 1), warp * 16, vl);
  }
  __syncthreads();

  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const float sl2 = scale * kLog2e;
  const int key0 = k0 + warp * 16 + (lane >> 2), key1 = key0 + 8;
  const int ntiles = (L + TQ - 1) / TQ;
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) load_q_tile(st ^ 1, (t + 1) * TQ);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_nt_s<SPLIT>(s, kh, kl, sQ(st, 0), sQ(st, NT - 1));        // S^T = K Q^T
    mma_nt_s<SPLIT>(dp, vh, vl, sG(st, 0), sG(st, NT - 1));       // dP^T = V dO^T
    const int qb = t * TQ + (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql_ = nt * 8 + (lane & 3) * 2 + (e & 1);
        const int q = qb + nt * 8 + (e & 1);
        const int key = e < 2 ? key0 : key1;
        float p = (q < L && key < L) ? exp2f(s[nt][e] * sl2 - s_lse[st * TQ + ql_]) : 0.f;
        s[nt][e] = p;
        dp[nt][e] = p * (dp[nt][e] - s_delta[st * TQ + ql_]);
      }
    }
    uint32_t ah[4][4], al[4][4];
    acc_to_a<SPLIT>(s, ah, al);
    mma_nn_s<SPLIT>(dv, ah, al, sG(st, 0), sG(st, NT - 1));       // dV += P^T dO
    acc_to_a<SPLIT>(dp, ah, al);
    mma_nn_s<SPLIT>(dk, ah, al, sQ(st, 0), sQ(st, NT - 1));       // dK += dS^T Q
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[nt][e] *= scale;
  }
  if (dv_add) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = e < 2 ? key0 : key1;
        const int col = h * D + nt * 8 + (lane & 3) * 2 + (e & 1);
        if (key < L) dv[nt][e] += load_as_f32(dv_add, dv_add_dtype, ((int64_t)b * L + key) * ld_dv_add + col, ld_dv_add / 2);
      }
    }
  }
  __nv_bfloat16* gout = dqkv + (int64_t)b * L * ldg + h * D;
  store_acc<SPLIT>(dk, gout + E, ldg, ldg / 2, k0 + warp * 16, L);
  store_acc<SPLIT>(dv, gout + 2 * E, ldg, ldg / 2, k0 + warp * 16, L);
}

// dQ for one tile of 64 queries; loops over key tiles.
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads)
attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ dout, int64_t lddo,
                  const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int64_t ldg, int L, int heads,
                  float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = SPLIT ? 2 : 1;
  const uint32_t s0 = ptx_smem(smem);
  auto sK = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + t) * kTileBytes); };
  auto sV = [&](int st, int t) { return s0 + (uint32_t)((st * 2 * NT + NT + t) * kTileBytes); };
  const int E = heads * D;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t lo = ld / 2, lodo = lddo / 2;
  const __nv_bfloat16* base = qkv + (int64_t)b * L * ld + h * D;
  const __nv_bfloat16 *gq = base, *gk = base + E, *gv = base + 2 * E;
  const __nv_bfloat16* gdo = dout + (int64_t)b * L * lddo + h * D;

  load_tile(sK(1, 0), gq, ld, q0, L);
  load_tile(sV(1, 0), gdo, lddo, q0, L);
  if (SPLIT) {
    load_tile(sK(1, NT - 1), gq + lo, ld, q0, L);
    load_tile(sV(1, NT - 1), gdo + lodo, lddo, q0, L);
  }
  cp_async_commit();
  auto load_kv = [&](int st, int k0) {
    load_tile(sK(st, 0), gk, ld, k0, L);
    load_tile(sV(st, 0), gv, ld, k0, L);
    if (SPLIT) {
      load_tile(sK(st, NT - 1), gk + lo, ld, k0, L);
      load_tile(sV(st, NT - 1), gv + lo, ld, k0, L);
    }
  };
  load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qh[4][4], ql[4][4], gh[4][4], gl[4][4];
  load_a_frags(sK(1, 0), warp * 16, qh);
  load_a_frags(sV(1, 0), warp * 16, gh);
  if (SPLIT) {
    load_a_frags(sK(1, NT - 1), warp * 16, ql);
    load_a_frags(sV(1, NT - 1), warp * 16, gl);
  }
  __syncthreads();

  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  const float* glse = lse + ((int64_t)b * heads + h) * L;
  const float* gdelta = delta + ((int64_t)b * heads + h) * L;
  const float lse0 = r0 < L ? glse[r0] * kLog2e : 0.f, lse1 = r1 < L ? glse[r1] * kLog2e : 0.f;
  const float dl0 = r0 < L ? gdelta[r0] : 0.f, dl1 = r1 < L ? gdelta[r1] : 0.f;

  float dq[8][4];
  zero_acc(dq);
  const float sl2 = scale * kLog2e;
  const int ntiles = (L + TK - 1) / TK;
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) load_kv(st ^ 1, (t + 1) * TK);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_nt_s<SPLIT>(s, qh, ql, sK(st, 0), sK(st, NT - 1));       // S = Q K^T
    mma_nt_s<SPLIT>(dp, gh, gl, sV(st, 0), sV(st, NT - 1));      // dP = dO V^T
    const int kb = t * TK + (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb + nt * 8 + (e & 1);
        const bool ok = key < L && (e < 2 ? r0 : r1) < L;
        const float p = ok ? exp2f(s[nt][e] * sl2 - (e < 2 ? lse0 : lse1)) : 0.f;
        dp[nt][e] = p * (dp[nt][e] - (e < 2 ? dl0 : dl1));
      }
    }
    uint32_t ah[4][4], al[4][4];
    acc_to_a<SPLIT>(dp, ah, al);
    mma_nn_s<SPLIT>(dq, ah, al, sK(st, 0), sK(st, NT - 1));      // dQ += dS K
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[nt][e] *= scale;
  }
  store_acc<SPLIT>(dq, dqkv + (int64_t)b * L * ldg + h * D, ldg, ldg / 2, q0 + warp * 16, L);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  SVL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SVL_OK;
}

}  // namespace
}  // namespace svl

using namespace svl;

static bool attention_uses_tc(int split, int L) {
  static int tc_mode = -1;
  if (tc_mode < 0) { const char* e = getenv("SVL_ATTN_TC"); tc_mode = e ? atoi(e) : 1; }
  return !split && tc_mode && L >= 64;
}

extern "C" size_t svl_attention_bwd_workspace(int split, int b, int L, int heads) {
  return attention_uses_tc(split, L) ? attention_bwd_tc_workspace(b, L, heads) : (size_t)b * heads * L;
}

extern "C" int svl_attention_fwd(const void* qkv, int split, void* out, float* lse, int b, int L, int heads, float scale, void* stream) {
  SVL_CHECK_ARG(qkv && out && b > 0 && L > 0 && heads > 0, "svl_attention_fwd: bad arguments");
  const int E = heads * D;
  const int64_t ld = split ? 6 * E : 3 * E, ldo = split ? 2 * E : E;
  if (attention_uses_tc(split, L)) {        // tcgen05 path (throughput mode); tiny sequences (class attention, L = #classes) stay on mma.sync
    if (int rc = svl_check_device()) return rc;
    return attention_fwd_tc(qkv, out, lse, b, L, heads, scale, (cudaStream_t)stream);
  }
  dim3 grid((L + TQ - 1) / TQ, heads, b);
  const size_t smem = (size_t)(split ? 8 : 4) * kTileBytes;
  if (split) {
    static bool once = false;
    if (!once) { if (int rc = set_smem(attn_fwd_kernel<true>, smem)) return rc; once = true; }
    attn_fwd_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)out, ldo, lse, L, heads, scale);
  } else {
    attn_fwd_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)out, ldo, lse, L, heads, scale);
  }
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

extern "C" int svl_attention_bwd(const void* qkv, const void* out, const void* dout, int split, const float* lse, float* delta_ws,
                                 const void* dv_add, int dv_add_dtype, int64_t ld_dv_add, void* dqkv, int b, int L, int heads, float scale,
                                 void* stream) {
  SVL_CHECK_ARG(qkv && out && dout && lse && delta_ws && dqkv && b > 0 && L > 0 && heads > 0, "svl_attention_bwd: bad arguments");
  const int E = heads * D;
  const int64_t ld = split ? 6 * E : 3 * E, ldo = split ? 2 * E : E;
  cudaStream_t st = (cudaStream_t)stream;
  if (attention_uses_tc(split, L)) {
    if (int rc = svl_check_device()) return rc;
    return attention_bwd_tc(qkv, out, dout, lse, delta_ws, dv_add, dv_add_dtype, ld_dv_add, dqkv, b, L, heads, scale, st);
  }
  const int64_t rows = (int64_t)b * L;
  int dgrid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
  attn_delta_kernel<<<dgrid, 256, 0, st>>>((const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)dout, ldo, split, delta_ws, rows, L, heads);
  SVL_LAUNCH_CHECK();
  dim3 grid((L + TQ - 1) / TQ, heads, b);
  const size_t smem_kv = (size_t)(split ? 8 : 4) * kTileBytes + 4 * TQ * sizeof(float);
  const size_t smem_q = (size_t)(split ? 8 : 4) * kTileBytes;
  if (split) {
    static bool once = false;
    if (!once) {
      if (int rc = set_smem(attn_bwd_kv_kernel<true>, smem_kv)) return rc;
      if (int rc = set_smem(attn_bwd_q_kernel<true>, smem_q)) return rc;
      once = true;
    }
    attn_bwd_kv_kernel<true><<<grid, kThreads, smem_kv, st>>>((const __nv_bfloat16*)qkv, ld, (const __nv_bfloat16*)dout, ldo, lse, delta_ws, dv_add,
                                                              dv_add_dtype, ld_dv_add, (__nv_bfloat16*)dqkv, ld, L, heads, scale);
    SVL_LAUNCH_CHECK();
    attn_bwd_q_kernel<true><<<grid, kThreads, smem_q, st>>>((const __nv_bfloat16*)qkv, ld, (const __nv_bfloat16*)dout, ldo, lse, delta_ws,
                                                            (__nv_bfloat16*)dqkv, ld, L, heads, scale);
  } else {
    attn_bwd_kv_kernel<false><<<grid, kThreads, smem_kv, st>>>((const __nv_bfloat16*)qkv, ld, (const __nv_bfloat16*)dout, ldo, lse, delta_ws, dv_add,
                                                               dv_add_dtype, ld_dv_add, (__nv_bfloat16*)dqkv, ld, L, heads, scale);
    SVL_LAUNCH_CHECK();
    attn_bwd_q_kernel<false><<<grid, kThreads, smem_q, st>>>((const __nv_bfloat16*)qkv, ld, (const __nv_bfloat16*)dout, ldo, lse, delta_ws,
                                                             (__nv_bfloat16*)dqkv, ld, L, heads, scale);
  }
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
