// Flash attention forward on tcgen05 (throughput mode, bf16 operands): S = Q K^T and O_tile = P V run on the 5th-gen tensor
// cores with accumulators in TMEM; the softmax warps own one query row per thread (no cross-thread reductions), read S with
// tcgen05.ld, write P as a K-major SWIZZLE_128B operand into shared memory and fold each O_tile into fp32 registers.
//
//   warp 0      : TMA producer (Q once; K/V tiles of 128 keys, 2-stage ring; 3-D tensor map -> rows >= L are zero-filled)
//   warp 1      : MMA issuer  (S_{j+1} is issued as soon as the softmax warps release S_j, BEFORE P_j V_j: the softmax of tile j+1 never
//                 waits behind a P V product; same order in the backward kernels: scores of tile j+1 before the accumulates of tile j)
//   warps 2..5  : softmax / output (thread = query row = TMEM lane)
//
// TMEM: S[2] (2 x 128 columns) | O_tile[2] (2 x 64 columns).  Shared memory: Q 16 KB, K/V 2 x 32 KB, P 2 x 32 KB.
// replaces the nn.MultiheadAttention core (maskclip_vit.py:77-84,141) for head_dim 64; the split-bf16 precise mode keeps the
// mma.sync kernels of attention.cu.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#ifdef SVL_GEMM_DIAG
#define ATT_T0() const long long att_t0 = clock64()
#define ATT_ADD(slot) do { if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) p.trace[slot] += clock64() - att_t0; } while (0)
#else
#define ATT_T0() do {} while (0)
#define ATT_ADD(slot) do {} while (0)
#endif
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
  int s_first;        // MMA issue order: 1 = S_{j+1} before P_j V_j (the softmax of tile j+1 does not wait behind the P V product)
  long long* trace;   // -DSVL_GEMM_DIAG: cycle totals of CTA (1, 0, 0): who waits for whom (SVL_ATTN_TRACE)
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
    // MMA issuer: the whole warp runs the loop, one elected lane issues, so the MMA operands stay in uniform registers
    // (a divergent single-thread loop costs ~30 instructions per tcgen05.mma, see gemm.cu)
    const bool leader = ptx::elect_one();
    const uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);
    const uint64_t tmpl = ptx::make_smem_desc(0, 8192, 1024);
    const uint64_t qd = tmpl + (uint64_t)(sQ >> 4), pa = tmpl + (uint64_t)(sP >> 4);
    ptx::mbar_wait(q_full, 0);
    auto issue_s = [&](int j) {              // S_j = Q K_j^T into the (single) S buffer once the softmax warps have released it
      const int st = j & 1, k = j >> 1;
      ptx::mbar_wait(kv_full(st), (uint32_t)(k & 1));
      ptx::mbar_wait(s_empty, (uint32_t)((j & 1) ^ 1));
      ptx::tc_fence_after();
      const uint64_t kd = tmpl + (uint64_t)((sK + st * kTile) >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tmem_base, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_s, kk > 0);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {             // O_tile[j & 1] = P_j V_j, then the K/V stage is free
      const int st = j & 1, k = j >> 1;
      const uint64_t vb = tmpl + (uint64_t)((sV + st * kTile) >> 4);
      ptx::mbar_wait(p_full, (uint32_t)(j & 1));
      ptx::mbar_wait(o_empty(st), (uint32_t)((k & 1) ^ 1));
      ptx::tc_fence_after();
      if (leader) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          // A: P k-block s/4 (16 KB apart), 32 bytes per 16 keys inside the swizzled row; B: V rows (keys) are the K dimension, 16 rows = 2048 bytes
          ptx::umma_bf16(tmem_base + 128u + (uint32_t)(st * 64), pa + (uint64_t)((s >> 2) * (kTile >> 4) + (s & 3) * 2), vb + (uint64_t)(s * 128),
                         idesc_pv, s > 0);
        }
        ptx::umma_commit(o_full(st));
        ptx::umma_commit(kv_empty(st));
      }
      __syncwarp();
    };
    if (p.s_first) {
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);      // waits for s_empty(j): the softmax warps are done reading S_j; P_j follows right behind
        issue_pv(j);
      }
    } else {
      for (int j = 0; j < nt; ++j) {
        issue_s(j);
        issue_pv(j);
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


// ================================================================================================ forward, second generation
// Same roles, different pipeline (profiles/README.md, round 2).  The first kernel above is SHARED-MEMORY-BANDWIDTH bound: per 128 x 128
// score tile it moves Q 16 KB + K 16 KB into the score product, writes and re-reads the 32 KB P tile, reads V 16 KB and takes 32 KB of
// TMA writes = 144 KB at 128 B/clk = 1152 clk per tile and CTA against 512 clk of tensor pipe, and it reads the scores twice from TMEM
// and folds every P V product into fp32 registers (168 registers, serial chain score -> softmax -> P -> P V).  Here
//   * both A operands live in TENSOR MEMORY (tcgen05.mma with a TMEM A operand): Q is copied there once (32 columns of packed bf16),
//     and P never touches shared memory -- a softmax thread packs its 64 exponentials to bf16 and stores them with tcgen05.st over the
//     first 32 columns of the score buffer it has just read.  Shared memory only carries the K / V tiles: 32 KB per 64 keys;
//   * keys advance in 64-wide steps with TWO score buffers: S_{j+1} / S_{j+2} are computed while the softmax threads work on S_j;
//     a thread reads its 64 scores ONCE and keeps them in registers for maximum, exponentials and the P row;
//   * O accumulates in TMEM across all key tiles; the reference maximum of a row is only moved when the running maximum exceeds it by
//     more than 8 (a factor 256 in the exponentials, harmless in bf16 P / fp32 sums), in which case the warp rescales its 32 TMEM lanes
//     of O in place (tcgen05.ld -> multiply -> tcgen05.st).  With attention logits this happens on the first tile and then almost never;
//   * K and V travel through separate rings (K is released by the score product, two tiles ahead of V);
//   * the tail: key columns >= L are neither exponentiated nor counted, warps whose 32 query rows are all >= L only keep the barrier
//     protocol going (L = 1025 = 8 * 128 + 1 leaves one valid row in the 9th query tile and one valid key in the 17th key tile);
//   (a variant with EIGHT softmax warps -- two threads per row exchanging partial row maxima through shared memory and a named barrier per
//    tile -- was measured slower, 111.5 vs 102.5 us per layer: the forward is not short of warps; profiles/r02_attention.md.)
// S_{j+2} overwrites score buffer j & 1 and the P_j stored inside it: the MMA warp issues it only after the commit of P_j V_j has arrived
// (issue order alone does not protect a TMEM A operand against a later accumulator write).
constexpr int TK2 = 64, KST = 4;
constexpr uint32_t kHalfTile = 64 * 128;       // bytes of a 64-row x 64-bf16 tile
constexpr float kRescaleThreshold = 8.f;       // log2 units

// D[tmem] (+)= A[tmem, packed bf16: 8 columns per K = 16] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 TMEM columns of this thread's lane; 16 packed columns stored
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  ptx::tmem_ld_32x32(taddr, v);
  ptx::tmem_ld_wait();
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
      "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  if ((raw & 1023u) != 0) __trap();
  const uint32_t sQ = raw, sK = sQ + kTile, sV = sK + KST * kHalfTile;
  const uint32_t bar = sV + KST * kHalfTile;
  const uint32_t q_full = bar, q_tmem = bar + 8, o_done = bar + 16;
  auto k_full = [&](int s) { return bar + 8u * (3 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (3 + KST + s); };
  auto v_full = [&](int s) { return bar + 8u * (3 + 2 * KST + s); };
  auto v_empty = [&](int s) { return bar + 8u * (3 + 3 * KST + s); };
  auto s_full = [&](int s) { return bar + 8u * (3 + 4 * KST + s); };
  auto p_full = [&](int s) { return bar + 8u * (5 + 4 * KST + s); };
  const uint32_t tmem_ptr_addr = bar + 8u * (7 + 4 * KST);
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int E = p.heads * D;
  const int nt = (p.L + TK2 - 1) / TK2;

  constexpr int kProducerWarp = 4, kMmaWarp = 5;       // softmax warps 0..3 (TMEM lane quarter = warp id); the issue warps get the highest ids
  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_tmem, 4);
    ptx::mbar_init(o_done, 1);
    for (int s = 0; s < KST; ++s) {
      ptx::mbar_init(k_full(s), 1);
      ptx::mbar_init(k_empty(s), 1);
      ptx::mbar_init(v_full(s), 1);
      ptx::mbar_init(v_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(s_full(s), 1);
      ptx::mbar_init(p_full(s), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(tmem_ptr_addr, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;            // S / P [2]: columns [0,64) [64,128); O: [128,192); Q (packed bf16): [192,224)

  if (warp == kProducerWarp) {
    // two producer lanes: the K ring runs two tiles ahead of the V ring and must not wait behind it
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, kTile);
      ptx::tma_load_3d(sQ, &tmQ, q_full, h * D, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int st = j % KST, k = j / KST;
        ptx::mbar_wait(k_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(k_full(st), kHalfTile);
        ptx::tma_load_3d(sK + st * kHalfTile, &tmKV, k_full(st), E + h * D, j * TK2, b);
      }
    } else if (lane == 16) {
      for (int j = 0; j < nt; ++j) {
        const int st = j % KST, k = j / KST;
        ptx::mbar_wait(v_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(v_full(st), kHalfTile);
        ptx::tma_load_3d(sV + st * kHalfTile, &tmKV, v_full(st), 2 * E + h * D, j * TK2, b);
      }
    }
  } else if (warp == kMmaWarp) {
    const bool leader = ptx::elect_one();
    const uint32_t idesc_s = ptx::make_idesc_bf16(128, 64, 0, 0);
    const uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);
    const uint64_t tmpl = ptx::make_smem_desc(0, 8192, 1024);
    ptx::mbar_wait(q_tmem, 0);
    ptx::tc_fence_after();
    auto issue_s = [&](int j) {              // S_j = Q K_j^T (A = Q in TMEM) into score buffer j & 1, then the K stage is free
      const int st = j % KST, sb = j & 1;
      { ATT_T0(); ptx::mbar_wait(k_full(st), (uint32_t)((j / KST) & 1)); ATT_ADD(1); }
      ptx::tc_fence_after();
      const uint64_t kd = tmpl + (uint64_t)((sK + st * kHalfTile) >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ts(tmem_base + (uint32_t)(sb * 64), tmem_base + 192u + (uint32_t)(kk * 8), kd + (uint64_t)(kk * 2), idesc_s, kk > 0);
        ptx::umma_commit(s_full(sb));
        ptx::umma_commit(k_empty(st));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {             // O += P_j V_j (A = P_j in the score buffer, K dimension = the 64 keys; V read MN-major)
      const int st = j % KST, sb = j & 1;
      { ATT_T0(); ptx::mbar_wait(v_full(st), (uint32_t)((j / KST) & 1)); ATT_ADD(2); }
      { ATT_T0(); ptx::mbar_wait(p_full(sb), (uint32_t)((j >> 1) & 1)); ATT_ADD(3); }
      ptx::tc_fence_after();
      const uint64_t vb = tmpl + (uint64_t)((sV + st * kHalfTile) >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ts(tmem_base + 128u, tmem_base + (uint32_t)(sb * 64 + kk * 8), vb + (uint64_t)(kk * 128), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(v_empty(st));
        ptx::umma_commit(o_done);
      }
      __syncwarp();
    };
    ATT_T0();
    issue_s(0);
    if (nt > 1) issue_s(1);
    for (int j = 0; j < nt; ++j) {
      issue_pv(j);
      if (j + 2 < nt) {
        // S_{j+2} overwrites score buffer j & 1 and the P_j stored inside it.  Issue order is NOT enough: the tensor pipe does not track a
        // later D write against an earlier TMEM A-operand read (run-to-run differences at 16 x 12 x 1025 proved it), so wait for P_j V_j.
        { const long long w0 = clock64(); ptx::mbar_wait(o_done, (uint32_t)(j & 1));
#ifdef SVL_GEMM_DIAG
          if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) p.trace[4] += clock64() - w0;
#endif
        }
        issue_s(j + 2);
      }
    }
    ATT_ADD(0);
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const bool warp_valid = q0 + q * 32 < p.L;            // warp-uniform: none of this warp's query rows exists otherwise
    {
      // Q row -> TMEM (the K-major SWIZZLE_128B tile holds row r at r * 128, 16-byte chunk c at (c ^ (r & 7)) * 16)
      ptx::mbar_wait(q_full, 0);
      const uint8_t* qrow = smem_raw + (sQ - raw) + r * 128;
      uint32_t qv[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 t = *(const uint4*)(qrow + ((c ^ (r & 7)) << 4));
        qv[c * 4] = t.x; qv[c * 4 + 1] = t.y; qv[c * 4 + 2] = t.z; qv[c * 4 + 3] = t.w;
      }
      ptx::tmem_st_32x32(tl + 192u, qv);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(q_tmem);
    }
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int sb = j & 1;
      { ATT_T0(); ptx::mbar_wait(s_full(sb), (uint32_t)((j >> 1) & 1)); if (warp == 0) ATT_ADD(5); }
      ptx::tc_fence_after();
      const int nvalid = p.L - j * TK2;                   // >= 64 except on the last tile
      if (warp_valid) {
        uint32_t v[64];
        tmem_ld64(tl + (uint32_t)(sb * 64), v);
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        if (nvalid >= 64) {
#pragma unroll
          for (int i = 0; i < 64; i += 8) {
            m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) if (i < nvalid) m0 = fmaxf(m0, __uint_as_float(v[i]));
        }
        const float ms = fmax3(m0, m1, fmaxf(m2, m3)) * sl2;
        const bool need = ms > m_ref + kRescaleThreshold;           // first tile: m_ref = -inf
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? ms : m_ref;
          const float corr = ex2(m_ref - m_new);                    // rows that keep their reference: ex2(0) = 1
          l *= corr;
          if (j > 0) {                                              // O holds the tiles before j: wait for P_{j-1} V_{j-1}, rescale in place
            ptx::mbar_wait(o_done, (uint32_t)((j - 1) & 1));
            ptx::tc_fence_after();
            uint32_t o[64];
            tmem_ld64(tl + 128u, o);
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            ptx::tmem_st_32x32(tl + 128u, o);
            ptx::tmem_st_32x32(tl + 160u, o + 32);
          }
          m_ref = m_new;
        }
        // P = exp2(s * scale - m_ref) as packed bf16 over the first 32 columns of the score buffer just read
        uint32_t pk[32];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (nvalid >= 64) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = ex2(fmaf(__uint_as_float(v[g * 8 + i]), sl2, -m_ref));
            s0 += pv[0] + pv[4]; s1 += pv[1] + pv[5]; s2 += pv[2] + pv[6]; s3 += pv[3] + pv[7];
            const uint4 t = f32_to_bf16x8(pv);
            pk[g * 4] = t.x; pk[g * 4 + 1] = t.y; pk[g * 4 + 2] = t.z; pk[g * 4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = 0.f;
            if (g * 8 < nvalid) {                         // warp-uniform: whole 8-key groups beyond L cost nothing
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (g * 8 + i < nvalid) {
                  pv[i] = ex2(fmaf(__uint_as_float(v[g * 8 + i]), sl2, -m_ref));
                  s0 += pv[i];
                }
              }
            }
            const uint4 t = f32_to_bf16x8(pv);
            pk[g * 4] = t.x; pk[g * 4 + 1] = t.y; pk[g * 4 + 2] = t.z; pk[g * 4 + 3] = t.w;
          }
        }
        l += (s0 + s1) + (s2 + s3);
        ptx::tmem_st_32x32(tl + (uint32_t)(sb * 64), pk);
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full(sb));
    }
    ptx::mbar_wait(o_done, (uint32_t)((nt - 1) & 1));
    ptx::tc_fence_after();
    const int row = q0 + r;
    if (warp_valid) {
      uint32_t o[64];
      tmem_ld64(tl + 128u, o);
      if (row < p.L) {
        const float inv = 1.f / l;
        float f[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) f[i] = __uint_as_float(o[i]) * inv;
        __nv_bfloat16* dst = p.out + ((int64_t)b * p.L + row) * p.ldo + h * D;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) *(uint4*)(dst + i * 8) = f32_to_bf16x8(f + i * 8);
        if (p.lse) p.lse[((int64_t)b * p.heads + h) * p.L + row] = (m_ref + log2f(l)) / kLog2e;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}


// ================================================================================================ backward (tcgen05)
// Workspace prepared by attn_prep_tc_kernel: delta_p / lse_p are [b*heads, Lp] (Lp = L rounded up to 64) so that 64-entry slices
// are 256-byte aligned bulk copies; lse_p is pre-multiplied by log2(e) and padded with +inf (=> P = 0 for queries >= L).
struct BwdParams {
  const float* lse_p; const float* delta_p;
  const void* dv_add; int dv_add_dtype; int64_t ld_dv_add;
  __nv_bfloat16* dqkv; int64_t ldg;
  int L, Lp, heads;
  float scale;
  int s_first;        // MMA issue order: 1 = the score products of tile j+1 are issued before the accumulate products of tile j
  long long* trace;   // -DSVL_GEMM_DIAG (SVL_ATTN_TRACE): cycle totals of CTA (1, 0, 0) of the kv kernel, slots 8..
};

__global__ void attn_prep_tc_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, int64_t ldo,
                                    const float* __restrict__ lse, float* __restrict__ lse_p, float* __restrict__ delta_p, int b, int L, int Lp,
                                    int heads) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)b * Lp;
  for (int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t bi = row / Lp, l = row % Lp;
    for (int h = 0; h < heads; ++h) {
      float s = 0.f, ls = INFINITY;
      if (l < L) {
        const int64_t src = (bi * L + l) * ldo + h * D + lane * 2;
        const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)(out + src));
        const float2 g = __bfloat1622float2(*(const __nv_bfloat162*)(dout + src));
        s = warp_sum(a.x * g.x + a.y * g.y);
        ls = lse[(bi * heads + h) * L + l] * kLog2e;
      }
      if (lane == 0) {
        delta_p[(bi * heads + h) * Lp + l] = s;
        lse_p[(bi * heads + h) * Lp + l] = ls;
      }
    }
  }
}

// heads % 4 == 0: 16-byte loads, 8 lanes per head (4 heads per 256-element chunk), all of a row's loads issued before the reductions
template <int CHUNKS>
__global__ void attn_prep_tc_vec_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, int64_t ldo,
                                        const float* __restrict__ lse, float* __restrict__ lse_p, float* __restrict__ delta_p, int b, int L, int Lp,
                                        int heads) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)b * Lp;
  for (int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t bi = row / Lp, l = row % Lp;
    uint4 a[CHUNKS], g[CHUNKS];
    if (l < L) {
      const int64_t src = (bi * L + l) * ldo + lane * 8;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        a[c] = __ldg((const uint4*)(out + src + c * 256));
        g[c] = __ldg((const uint4*)(dout + src + c * 256));
      }
    }
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int h = c * 4 + (lane >> 3);
      float s = 0.f, ls = INFINITY;
      if (l < L) {
        float fa[8], fg[8];
        bf16x8_to_f32(a[c], fa);
        bf16x8_to_f32(g[c], fg);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += fa[i] * fg[i];
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((lane & 7) == 0) {
        if (l < L) ls = lse[(bi * heads + h) * L + l] * kLog2e;
        delta_p[(bi * heads + h) * Lp + l] = s;
        lse_p[(bi * heads + h) * Lp + l] = ls;
      }
    }
  }
}

// write one 128-byte operand row (64 bf16) of a K-major SWIZZLE_128B tile
__device__ __forceinline__ void st_row64(uint8_t* row_base, int r, const float* v) {
#pragma unroll
  for (int piece = 0; piece < 8; ++piece) *(uint4*)(row_base + ((piece ^ (r & 7)) << 4)) = f32_to_bf16x8(v + piece * 8);
}

// dK, dV of one 128-key tile; streams 64-query tiles.  Thread = key row.
//   S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T = exp2(S^T*c - lse[q]), dS^T = P^T (dP^T - delta[q]) -> packed bf16 stored with tcgen05.st
//   over the first 32 columns of the S^T / dP^T buffers the thread has just read -> dV += P^T dO, dK += dS^T Q with the A operand in
//   TENSOR MEMORY (TMEM accumulators over all query tiles; Q / dO tiles are re-read MN-major from the same smem bytes).
// Round 2: the operand tiles used to go through shared memory (2 x 16 KB written, 2 x 16 KB read per query tile: 144 KB of shared-memory
// traffic per tile at 128 B/clk against 512 clk of tensor pipe); now 80 KB.  The score products of tile j+1 overwrite P^T_j / dS^T_j, so
// they are issued only after the accumulate products of tile j have completed; the second CTA of the SM fills the gap.
// Backward kernels: EIGHT softmax warps (two threads per row, 32 of the tile's 64 columns each -- P and dS are element-wise given the saved
// lse / delta, so the split needs no exchange): four warps per scheduler over the SM's two CTAs instead of two hide the TMEM-load and MUFU
// latencies the one-thread-per-row version exposed (XU pipe 30 % busy, long-scoreboard stalls dominant; profiles/r02_ncu_attention.md).
constexpr int kBwdThreads = 320;

__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_kv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKV64, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  if ((raw & 1023u) != 0) __trap();
  constexpr uint32_t kHalf = kTile / 2;                  // 64-row tile: 8 KB
  const uint32_t sK = raw, sV = raw + kTile, sQ = sV + kTile, sG = sQ + 2 * kHalf;
  const uint32_t sL = sG + 2 * kHalf;                    // [2][64] lse_p | [2][64] delta_p
  const uint32_t bar = sL + 1024;
  const uint32_t kv_full = bar, s_full = bar + 8, s_empty = bar + 16, p_full = bar + 24, p_empty = bar + 32, acc_full = bar + 40;
  auto q_full = [&](int s) { return bar + 8u * (6 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (8 + s); };
  const uint32_t tmem_ptr_addr = bar + 8u * 10;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int E = p.heads * D;
  const int nt = p.Lp / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV);
    ptx::prefetch_tmap(&tmQKV64);
    ptx::prefetch_tmap(&tmDO);
    ptx::mbar_init(kv_full, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 8);
    ptx::mbar_init(p_full, 8);
    ptx::mbar_init(p_empty, 1);
    ptx::mbar_init(acc_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(q_full(s), 1);
      ptx::mbar_init(q_empty(s), 1);
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
  const uint32_t tmem_base = *tmem_ptr_gen;      // S^T [0,64) | dP^T [64,128) | dV [128,192) | dK [192,256)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * kTile);
      ptx::tma_load_3d(sK, &tmQKV, kv_full, E + h * D, k0, b);
      ptx::tma_load_3d(sV, &tmQKV, kv_full, 2 * E + h * D, k0, b);
      const float* lsrc = p.lse_p + ((int64_t)b * p.heads + h) * p.Lp;
      const float* dsrc = p.delta_p + ((int64_t)b * p.heads + h) * p.Lp;
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1, k = j >> 1;
        ptx::mbar_wait(q_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(q_full(st), 2 * kHalf + 512);
        ptx::tma_load_3d(sQ + st * kHalf, &tmQKV64, q_full(st), h * D, j * 64, b);
        ptx::tma_load_3d(sG + st * kHalf, &tmDO, q_full(st), h * D, j * 64, b);
        ptx::bulk_load(sL + st * 256, lsrc + j * 64, 256, q_full(st));
        ptx::bulk_load(sL + 512 + st * 256, dsrc + j * 64, 256, q_full(st));
      }
    }
  } else if (warp == 1) {
    const bool leader = ptx::elect_one();                                 // warp-uniform issue loop (see the forward kernel)
    const uint32_t idesc_s = ptx::make_idesc_bf16(128, 64, 0, 0);       // scores: both operands K-major (d contiguous)
    const uint32_t idesc_a = ptx::make_idesc_bf16(128, 64, 0, 1);       // accumulates: B = Q / dO tile read MN-major
    const uint64_t tmpl = ptx::make_smem_desc(0, 8192, 1024);
    const uint64_t kd = tmpl + (uint64_t)(sK >> 4), vd = tmpl + (uint64_t)(sV >> 4);
    ptx::mbar_wait(kv_full, 0);
    auto scores = [&](int j) {               // S^T_j = K Q_j^T and dP^T_j = V dO_j^T (issued after the accumulates of tile j-1, which read the same columns)
      const int st = j & 1, k = j >> 1;
      { ATT_T0(); ptx::mbar_wait(q_full(st), (uint32_t)(k & 1)); ATT_ADD(9); }
      ptx::tc_fence_after();
      const uint64_t qd = tmpl + (uint64_t)((sQ + st * kHalf) >> 4), gd = tmpl + (uint64_t)((sG + st * kHalf) >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tmem_base, kd + (uint64_t)(kk * 2), qd + (uint64_t)(kk * 2), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tmem_base + 64u, vd + (uint64_t)(kk * 2), gd + (uint64_t)(kk * 2), idesc_s, kk > 0);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    auto accumulate = [&](int j) {           // dV += P^T_j dO_j, dK += dS^T_j Q_j, then the operand buffers and the Q / dO stage are free
      const int st = j & 1;
      const uint64_t qd = tmpl + (uint64_t)((sQ + st * kHalf) >> 4), gd = tmpl + (uint64_t)((sG + st * kHalf) >> 4);
      { ATT_T0(); ptx::mbar_wait(p_full, (uint32_t)(j & 1)); ATT_ADD(10); }
      ptx::tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)      // K = 64 queries: A = P^T in TMEM (8 columns per 16 queries; queries 32.. start at column 32), B 16 query rows = 2048 bytes
          umma_bf16_ts(tmem_base + 128u, tmem_base + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), gd + (uint64_t)(kk * 128), idesc_a, (j > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ts(tmem_base + 192u, tmem_base + 64u + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), qd + (uint64_t)(kk * 128), idesc_a, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(q_empty(st));
      }
      __syncwarp();
    };
    const long long att_tm0 = clock64();
    for (int j = 0; j < nt; ++j) {
      // the score products overwrite P^T_{j-1} / dS^T_{j-1}: wait until the accumulate products that read them have completed
      // (their commit on q_empty; issue order alone does not protect a TMEM A operand against a later accumulator write)
      if (j > 0) { ATT_T0(); ptx::mbar_wait(q_empty((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1)); ATT_ADD(11); }
      scores(j);
      accumulate(j);
    }
#ifdef SVL_GEMM_DIAG
    if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) p.trace[8] = clock64() - att_tm0;
#endif
    if (leader) ptx::umma_commit(acc_full);
    __syncwarp();
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;              // TMEM lane quarter (= warp % 4), column half of the 64-query tile
    const int r = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const float4* ls4 = (const float4*)(smem_raw + (sL - raw));
    const bool warp_valid = k0 + q * 32 < p.L;                  // warp-uniform: none of this warp's key rows exists otherwise
    for (int j = 0; j < nt; ++j) {
      const int st = j & 1, k = j >> 1;
      ptx::mbar_wait(q_full(st), (uint32_t)(k & 1));            // lse / delta slices of this query tile
      { ATT_T0(); ptx::mbar_wait(s_full, (uint32_t)(j & 1)); if (warp == 2) ATT_ADD(12); }
      ptx::tc_fence_after();
      if (warp_valid) {
        uint32_t sv[32], dv[32], pk[16], dk[16];
        ptx::tmem_ld_32x32(tl + (uint32_t)(half * 32), sv);
        ptx::tmem_ld_32x32(tl + 64u + (uint32_t)(half * 32), dv);
        ptx::tmem_ld_wait();
        const float4* lq4 = ls4 + st * 16 + half * 8;
        const float4* dq4 = ls4 + 32 + st * 16 + half * 8;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 la = lq4[2 * g], lb = lq4[2 * g + 1], da = dq4[2 * g], db = dq4[2 * g + 1];     // broadcast reads: 4 queries per LDS.128
          const float lqv[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
          const float dqv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
          float pt[8], ds[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pt[i] = ex2(fmaf(__uint_as_float(sv[g * 8 + i]), sl2, -lqv[i]));
            ds[i] = pt[i] * (__uint_as_float(dv[g * 8 + i]) - dqv[i]);
          }
          const uint4 a = f32_to_bf16x8(pt), c = f32_to_bf16x8(ds);
          pk[g * 4] = a.x; pk[g * 4 + 1] = a.y; pk[g * 4 + 2] = a.z; pk[g * 4 + 3] = a.w;
          dk[g * 4] = c.x; dk[g * 4 + 1] = c.y; dk[g * 4 + 2] = c.z; dk[g * 4 + 3] = c.w;
        }
        // packed bf16 operands over the first 16 of the 32 score columns THIS thread has just read (no other thread touches them):
        // P^T for queries [32 * half, +32) at S^T[32 * half, +16), dS^T likewise inside dP^T -- the MMA takes one A address per 16 queries
        tmem_st16(tl + (uint32_t)(half * 32), pk);
        tmem_st16(tl + 64u + (uint32_t)(half * 32), dk);
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const int key = k0 + r;
    uint32_t a[32];
    float f[32];
    // dV (+ v-path contribution): this thread's 32 of the 64 head dimensions
    tmem_ld32(tl + 128u + (uint32_t)(half * 32), a);
    if (key < p.L) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]);
      if (p.dv_add) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float t[8];
          ld8(p.dv_add, p.dv_add_dtype, ((int64_t)b * p.L + key) * p.ld_dv_add + h * D + half * 32 + g * 8, p.ld_dv_add / 2, 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[g * 8 + i] += t[i];
        }
      }
      __nv_bfloat16* dst = p.dqkv + ((int64_t)b * p.L + key) * p.ldg + 2 * E + h * D + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) *(uint4*)(dst + i * 8) = f32_to_bf16x8(f + i * 8);
    }
    tmem_ld32(tl + 192u + (uint32_t)(half * 32), a);
    if (key < p.L) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]) * p.scale;
      __nv_bfloat16* dst = p.dqkv + ((int64_t)b * p.L + key) * p.ldg + E + h * D + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) *(uint4*)(dst + i * 8) = f32_to_bf16x8(f + i * 8);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// dQ of one 128-query tile; streams 64-key tiles.  Thread = query row.
// Round 2: dO is copied to TENSOR MEMORY once (A operand of dP = dO V^T) and dS goes from the softmax threads to its own 32 TMEM columns
// with tcgen05.st (A operand of dQ += dS K): shared memory no longer carries the dO re-reads and the dS tile (104 -> 56 KB per key tile).
__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKV64, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  if ((raw & 1023u) != 0) __trap();
  constexpr uint32_t kHalf = kTile / 2;
  const uint32_t sQ = raw, sG = raw + kTile, sK = sG + kTile, sV = sK + 2 * kHalf;
  const uint32_t bar = sV + 2 * kHalf;
  const uint32_t q_full = bar, s_full = bar + 8, s_empty = bar + 16, p_full = bar + 24, p_empty = bar + 32, acc_full = bar + 40;
  auto kv_full = [&](int s) { return bar + 8u * (6 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (8 + s); };
  const uint32_t g_tmem = bar + 8u * 10;
  const uint32_t tmem_ptr_addr = bar + 8u * 11;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int E = p.heads * D;
  const int nt = (p.L + 63) / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV);
    ptx::prefetch_tmap(&tmQKV64);
    ptx::prefetch_tmap(&tmDO);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(g_tmem, 8);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 8);
    ptx::mbar_init(p_full, 8);
    ptx::mbar_init(p_empty, 1);
    ptx::mbar_init(acc_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(kv_full(s), 1);
      ptx::mbar_init(kv_empty(s), 1);
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
  const uint32_t tmem_base = *tmem_ptr_gen;      // S [0,64) | dP [64,128) | dQ [128,192) | dO packed bf16 [192,224) | dS packed bf16 [224,256)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, 2 * kTile);
      ptx::tma_load_3d(sQ, &tmQKV, q_full, h * D, q0, b);
      ptx::tma_load_3d(sG, &tmDO, q_full, h * D, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1, k = j >> 1;
        ptx::mbar_wait(kv_empty(st), (uint32_t)((k & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(kv_full(st), 2 * kHalf);
        ptx::tma_load_3d(sK + st * kHalf, &tmQKV64, kv_full(st), E + h * D, j * 64, b);
        ptx::tma_load_3d(sV + st * kHalf, &tmQKV64, kv_full(st), 2 * E + h * D, j * 64, b);
      }
    }
  } else if (warp == 1) {
    const bool leader = ptx::elect_one();                                 // warp-uniform issue loop (see the forward kernel)
    const uint32_t idesc_s = ptx::make_idesc_bf16(128, 64, 0, 0);
    const uint32_t idesc_a = ptx::make_idesc_bf16(128, 64, 0, 1);
    const uint64_t tmpl = ptx::make_smem_desc(0, 8192, 1024);
    const uint64_t qd = tmpl + (uint64_t)(sQ >> 4);
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(g_tmem, 0);
    ptx::tc_fence_after();
    auto scores = [&](int j) {               // S_j = Q K_j^T, dP_j = dO V_j^T (A = dO in TMEM)
      const int st = j & 1, k = j >> 1;
      ptx::mbar_wait(kv_full(st), (uint32_t)(k & 1));
      ptx::mbar_wait(s_empty, (uint32_t)((j & 1) ^ 1));
      ptx::tc_fence_after();
      const uint64_t kd = tmpl + (uint64_t)((sK + st * kHalf) >> 4), vd = tmpl + (uint64_t)((sV + st * kHalf) >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tmem_base, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ts(tmem_base + 64u, tmem_base + 192u + (uint32_t)(kk * 8), vd + (uint64_t)(kk * 2), idesc_s, kk > 0);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    auto accumulate = [&](int j) {           // dQ += dS_j K_j (A = dS in TMEM)
      const int st = j & 1;
      const uint64_t kd = tmpl + (uint64_t)((sK + st * kHalf) >> 4);
      ptx::mbar_wait(p_full, (uint32_t)(j & 1));
      ptx::tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)      // dQ += dS K : K dimension = the 64 keys (K tile read MN-major, 16 key rows = 2048 bytes)
          umma_bf16_ts(tmem_base + 128u, tmem_base + 224u + (uint32_t)(kk * 8), kd + (uint64_t)(kk * 128), idesc_a, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(p_empty);
        ptx::umma_commit(kv_empty(st));
      }
      __syncwarp();
    };
    if (p.s_first) {
      scores(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) scores(j + 1);
        accumulate(j);
      }
    } else {
      for (int j = 0; j < nt; ++j) {
        scores(j);
        accumulate(j);
      }
    }
    if (leader) ptx::umma_commit(acc_full);
    __syncwarp();
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;              // TMEM lane quarter (= warp % 4), column half of the 64-key tile
    const int r = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const int row = q0 + r;
    const bool warp_valid = q0 + q * 32 < p.L;                  // warp-uniform: none of this warp's query rows exists otherwise
    {
      // this thread's half of the dO row -> TMEM (K-major SWIZZLE_128B tile: row r at r * 128, 16-byte chunk c at (c ^ (r & 7)) * 16)
      ptx::mbar_wait(q_full, 0);
      const uint8_t* grow = smem_raw + (sG - raw) + r * 128;
      uint32_t gv[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 t = *(const uint4*)(grow + (((half * 4 + c) ^ (r & 7)) << 4));
        gv[c * 4] = t.x; gv[c * 4 + 1] = t.y; gv[c * 4 + 2] = t.z; gv[c * 4 + 3] = t.w;
      }
      tmem_st16(tl + 192u + (uint32_t)(half * 16), gv);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(g_tmem);
    }
    const int64_t li = ((int64_t)b * p.heads + h) * p.Lp + (row < p.Lp ? row : p.Lp - 1);
    const float lse_r = row < p.Lp ? p.lse_p[li] : INFINITY, delta_r = row < p.Lp ? p.delta_p[li] : 0.f;
    for (int j = 0; j < nt; ++j) {
      ptx::mbar_wait(s_full, (uint32_t)(j & 1));
      ptx::tc_fence_after();
      uint32_t sv[32], dv[32];
      if (warp_valid) {
        ptx::tmem_ld_32x32(tl + (uint32_t)(half * 32), sv);
        ptx::tmem_ld_32x32(tl + 64u + (uint32_t)(half * 32), dv);
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
      uint32_t dk[16];
      if (warp_valid) {
        const int nvalid = p.L - j * 64 - half * 32;            // >= 32 except on the last tile
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float ds[8];
          if (nvalid >= 32 || g * 8 < nvalid) {                 // warp-uniform: whole 8-key groups beyond L cost nothing
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float pr = ex2(fmaf(__uint_as_float(sv[g * 8 + i]), sl2, -lse_r));
              if (nvalid < 32 && g * 8 + i >= nvalid) pr = 0.f;
              ds[i] = pr * (__uint_as_float(dv[g * 8 + i]) - delta_r);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ds[i] = 0.f;
          }
          const uint4 c = f32_to_bf16x8(ds);
          dk[g * 4] = c.x; dk[g * 4 + 1] = c.y; dk[g * 4 + 2] = c.z; dk[g * 4 + 3] = c.w;
        }
      }
      ptx::mbar_wait(p_empty, (uint32_t)((j & 1) ^ 1));         // dQ += dS_{j-1} K_{j-1} has consumed the dS columns
      if (warp_valid) {
        ptx::tc_fence_after();
        tmem_st16(tl + 224u + (uint32_t)(half * 16), dk);
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    uint32_t a[32];
    tmem_ld32(tl + 128u + (uint32_t)(half * 32), a);
    if (row < p.L) {
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(a[i]) * p.scale;
      __nv_bfloat16* dst = p.dqkv + ((int64_t)b * p.L + row) * p.ldg + h * D + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) *(uint4*)(dst + i * 8) = f32_to_bf16x8(f + i * 8);
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
  p.trace = nullptr;
#ifdef SVL_GEMM_DIAG
  { const char* e = getenv("SVL_ATTN_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 10) : nullptr; }
#endif
  static int gen = -1;
  if (gen < 0) { const char* e = getenv("SVL_ATTN_FWD"); gen = e ? atoi(e) : 2; }      // 1: the first-generation kernel (kept for comparison)
  if (gen == 2) {
    CUtensorMap tmKV;
    uint32_t box64[3] = {64u, 64u, 1u};
    if (int rc = tma_encode_bf16(&tmKV, qkv, 3, dims, strides, box64)) return rc;
    p.s_first = 1;
    const size_t smem2 = kTile + 2 * KST * kHalfTile + 8 * (8 + 4 * KST) + 16;
    static bool attr2 = false;
    if (!attr2) {
      SVL_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      attr2 = true;
    }
    dim3 grid2((L + TQ - 1) / TQ, heads, b);
    attn_fwd_tc2_kernel<<<grid2, kThreads, smem2, stream>>>(tm, tmKV, p);
    SVL_LAUNCH_CHECK();
    return SVL_OK;
  }
  static int s_first = -1;
  if (s_first < 0) { const char* e = getenv("SVL_ATTN_S_FIRST"); s_first = e ? atoi(e) : 1; }   // measured: forward 168 -> 153 us per layer at 16 x 12 x 1025 (profiles/README.md)
  p.s_first = s_first;
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

size_t attention_bwd_tc_workspace(int b, int L, int heads) {
  const int Lp = (L + 63) / 64 * 64;
  return (size_t)2 * b * heads * Lp;
}

int attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, float* ws, const void* dv_add, int dv_add_dtype,
                     int64_t ld_dv_add, void* dqkv, int b, int L, int heads, float scale, cudaStream_t stream) {
  const int E = heads * D;
  const int Lp = (L + 63) / 64 * 64;
  float* lse_p = ws;
  float* delta_p = ws + (size_t)b * heads * Lp;
  const int64_t rows = (int64_t)b * Lp;
  int pgrid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
  if (heads == 12 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)dout & 15) == 0)
    attn_prep_tc_vec_kernel<3><<<pgrid, 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, E, lse, lse_p, delta_p, b, L, Lp, heads);
  else
    attn_prep_tc_kernel<<<pgrid, 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, E, lse, lse_p, delta_p, b, L, Lp, heads);
  SVL_LAUNCH_CHECK();
  CUtensorMap tmQKV64, tmQKV128, tmDO64, tmDO128;
  uint64_t dims[3] = {(uint64_t)3 * E, (uint64_t)L, (uint64_t)b};
  uint64_t strides[2] = {(uint64_t)3 * E * 2, (uint64_t)L * 3 * E * 2};
  uint64_t dimso[3] = {(uint64_t)E, (uint64_t)L, (uint64_t)b};
  uint64_t strideso[2] = {(uint64_t)E * 2, (uint64_t)L * E * 2};
  uint32_t box64[3] = {64u, 64u, 1u}, box128[3] = {64u, 128u, 1u};
  if (int rc = tma_encode_bf16(&tmQKV64, qkv, 3, dims, strides, box64)) return rc;
  if (int rc = tma_encode_bf16(&tmQKV128, qkv, 3, dims, strides, box128)) return rc;
  if (int rc = tma_encode_bf16(&tmDO64, dout, 3, dimso, strideso, box64)) return rc;
  if (int rc = tma_encode_bf16(&tmDO128, dout, 3, dimso, strideso, box128)) return rc;
  BwdParams p;
  p.lse_p = lse_p; p.delta_p = delta_p; p.dv_add = dv_add; p.dv_add_dtype = dv_add_dtype; p.ld_dv_add = ld_dv_add;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ldg = 3 * E; p.L = L; p.Lp = Lp; p.heads = heads; p.scale = scale;
  p.trace = nullptr;
#ifdef SVL_GEMM_DIAG
  { const char* e = getenv("SVL_ATTN_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 10) : nullptr; }
#endif
  static int s_first_b = -1;
  if (s_first_b < 0) { const char* e = getenv("SVL_ATTN_S_FIRST"); s_first_b = e ? atoi(e) : 1; }   // measured: backward 415 -> 377 us per layer
  p.s_first = s_first_b;
  const size_t smem_kv = 4 * (size_t)kTile + 1024 + 8 * 11 + 16, smem_q = 4 * (size_t)kTile + 8 * 12 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SVL_CUDA(cudaFuncSetAttribute(attn_bwd_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
    SVL_CUDA(cudaFuncSetAttribute(attn_bwd_q_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    attr_set = true;
  }
  dim3 grid((L + 127) / 128, heads, b);
  // the kv kernel keeps 128-row K/V tiles and streams 64-row Q/dO tiles, the q kernel the other way round
  attn_bwd_kv_tc_kernel<<<grid, kBwdThreads, smem_kv, stream>>>(tmQKV128, tmQKV64, tmDO64, p);
  SVL_LAUNCH_CHECK();
  attn_bwd_q_tc_kernel<<<grid, kBwdThreads, smem_q, stream>>>(tmQKV128, tmQKV64, tmDO128, p);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}

}  // namespace svl
