// Weight-gradient contraction on tcgen05 with MN-major operands:
//
//   DW[slot][i][j] += alpha * sum_{taps t of slot} sum_{rows r} DY[r, dy_koff_t + i] * X[shift_t(r), x_koff_t + j]
//
// Both operands are read straight from their forward (row = pixel/token, channels contiguous) layouts: a TMA box of
// 64 rows x 64 channels lands in shared memory as 64 K-rows of 128 bytes, which is exactly the canonical MN-major
// SWIZZLE_128B operand tile (8 K-rows per 1024-byte group, 64-channel chunks one box apart).  The contraction runs over
// rows; it is split across CTAs (split-K) and the partial tiles are reduced with red.global.add.f32.
// replaces: autograd's wgrad of nn.Linear / nn.Conv2d / nn.ConvTranspose2d on the hot path (semivl.py:327).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.h"

namespace svl {
namespace {

constexpr int BM = 128;       // dy channels per tile
constexpr int KB = 64;        // rows per K block
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;
constexpr uint32_t kBoxBytes = 64 * 128;

struct WgradParams {
  int conv;
  int64_t rows;
  int nb, h, w, bw, bh, bn, tiles_x, tiles_y;
  int64_t num_kblocks;            // K blocks over all rows
  int m, n, block_n, num_m_tiles, num_n_tiles, num_slots, splits;
  int num_taps;
  int tap_dy[SVL_MAX_TAPS], tap_dx[SVL_MAX_TAPS], tap_dy_koff[SVL_MAX_TAPS], tap_x_koff[SVL_MAX_TAPS];
  int slot_tap0[SVL_MAX_TAPS + 1];   // taps of slot s are [slot_tap0[s], slot_tap0[s+1])
  uint32_t k_tx_bytes;            // bytes one 64-channel box transfers (rows in the box * 128)
  int stages, tmem_cols;
  float* dw; int64_t ld_dw, slot_stride;
  float alpha;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  const uint32_t a_stage = 2 * kBoxBytes, b_stage = (uint32_t)(p.block_n / 64) * kBoxBytes;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + p.stages * a_stage;
  const uint32_t bar_base = smem_b + p.stages * b_stage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kMaxStages + 1);
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decomposition
  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int m_tile = bid % p.num_m_tiles; bid /= p.num_m_tiles;
  const int n_tile = bid % p.num_n_tiles; bid /= p.num_n_tiles;
  const int slot = bid;
  const int64_t kb_per = (p.num_kblocks + p.splits - 1) / p.splits;
  const int64_t kb0 = split * kb_per;
  const int64_t kb1 = kb0 + kb_per < p.num_kblocks ? kb0 + kb_per : p.num_kblocks;
  const int tap0 = p.slot_tap0[slot], tap1 = p.slot_tap0[slot + 1];
  const bool has_work = kb1 > kb0 && tap1 > tap0;

  // partial boxes (tiny images) leave K-rows of a stage untouched: they must read as zero
  if (p.k_tx_bytes < kBoxBytes) {
    uint4* z = (uint4*)(smem_raw + (smem_base - raw));
    const int n16 = (int)((p.stages * (a_stage + b_stage)) / 16);
    for (int i = threadIdx.x; i < n16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int nchunks_b = p.block_n / 64;

  if (has_work) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        // K-block -> box origin: one division at the start, then counters (a 64-bit division per K block made this lone thread,
        // not the tensor pipe, the bound of the narrow-channel convolutions)
        const int tx0 = p.conv ? (int)(kb0 % p.tiles_x) : 0;
        const int ty0 = p.conv ? (int)((kb0 / p.tiles_x) % p.tiles_y) : 0;
        const int tn0 = p.conv ? (int)(kb0 / ((int64_t)p.tiles_x * p.tiles_y)) : 0;
        for (int t = tap0; t < tap1; ++t) {
          int tx = tx0, ty = ty0, tn = tn0;
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
            ptx::mbar_arrive_expect_tx(full_bar(stage), p.k_tx_bytes * (2 + nchunks_b));
            const uint32_t sa = smem_a + stage * a_stage, sb = smem_b + stage * b_stage;
            const int ca = p.tap_dy_koff[t] + m_tile * BM, cb = p.tap_x_koff[t] + n_tile * p.block_n;
            if (p.conv) {
              const int cx = tx * p.bw, cy = ty * p.bh, cn = tn * p.bn;
              if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
              for (int c = 0; c < 2; ++c) ptx::tma_load_4d(sa + c * kBoxBytes, &tmDY, full_bar(stage), ca + c * 64, cx, cy, cn);
              for (int c = 0; c < nchunks_b; ++c)
                ptx::tma_load_4d(sb + c * kBoxBytes, &tmX, full_bar(stage), cb + c * 64, cx + p.tap_dx[t], cy + p.tap_dy[t], cn);
            } else {
              const int r0 = (int)(kb * KB);
              for (int c = 0; c < 2; ++c) ptx::tma_load_2d(sa + c * kBoxBytes, &tmDY, full_bar(stage), ca + c * 64, r0);
              for (int c = 0; c < nchunks_b; ++c) ptx::tma_load_2d(sb + c * kBoxBytes, &tmX, full_bar(stage), cb + c * 64, r0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      // whole warp in the loop, one elected lane issues: the MMA operands stay in uniform registers (see gemm.cu)
      const uint32_t idesc = ptx::make_idesc_bf16(BM, p.block_n, 1, 1);
      const bool leader = ptx::elect_one();
      const uint64_t tmpl = ptx::make_smem_desc(0, kBoxBytes, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int t = tap0; t < tap1; ++t) {
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t adesc = tmpl + (uint64_t)((smem_a + stage * a_stage) >> 4), bdesc = tmpl + (uint64_t)((smem_b + stage * b_stage) >> 4);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)   // 16 K-rows = 2048 bytes further into every chunk
              ptx::umma_bf16(tmem_base, adesc + (uint64_t)(kk * 128), bdesc + (uint64_t)(kk * 128), idesc, kk > 0 ? 1u : accum);
            ptx::umma_commit(empty_bar(stage));
          }
          accum = 1;
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (leader) ptx::umma_commit(tfull_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int i = m_tile * BM + q * 32 + lane;       // output row (dy channel)
      ptx::mbar_wait(tfull_bar, 0);
      ptx::tc_fence_after();
      float* row = p.dw + (int64_t)slot * p.slot_stride + (int64_t)i * p.ld_dw;
      const int n0 = n_tile * p.block_n;
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        if (n0 + c0 >= p.n) break;
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
        if (i >= p.m) continue;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int col = n0 + c0 + g * 4;
          float* dst = row + col;
          float f0 = __uint_as_float(v[g * 4]) * p.alpha, f1 = __uint_as_float(v[g * 4 + 1]) * p.alpha,
                f2 = __uint_as_float(v[g * 4 + 2]) * p.alpha, f3 = __uint_as_float(v[g * 4 + 3]) * p.alpha;
          if (col + 4 <= p.n && ((uintptr_t)dst & 15) == 0) {
            red_add_v4(dst, f0, f1, f2, f3);
          } else {
            if (col < p.n) atomicAdd(dst, f0);
            if (col + 1 < p.n) atomicAdd(dst + 1, f1);
            if (col + 2 < p.n) atomicAdd(dst + 2, f2);
            if (col + 3 < p.n) atomicAdd(dst + 3, f3);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------------------------------------
// Strip variant for k x k convolutions: the taps of one filter ROW (same dy, different dx) share one load of the dy tile and one
// load of an x strip that is (bw + span) pixels wide; each tap's operand is the same strip read at a different K-row offset
// (matrix-descriptor start address + base offset), so x is fetched 3x instead of 9x and dy 3x instead of 9x from L2.
constexpr int kMaxGroupTaps = 4;
struct StripParams {
  int nb, h, w, bw, bh, bn, tiles_x, tiles_y;
  int64_t num_kblocks;
  int m, n, block_n, num_m_tiles, num_n_tiles, ngroups, splits;
  int g_ntaps[8], g_dy[8], g_dxmin[8];
  int g_dx[8][kMaxGroupTaps], g_slot[8][kMaxGroupTaps];
  int dy_koff, x_koff;
  int strip_w, strip_rows;
  uint32_t strip_stride, a_chunks;
  int stages, tmem_cols, bo_mode;
  int npack, dxstep, g_nmma[8];      // npack: the taps of a group are 64-column chunks of ONE MMA (B chunk stride = dxstep pixels)
  float* dw; int64_t ld_dw, slot_stride;
  float alpha;
  uint32_t mma_boff[8][kMaxGroupTaps * (KB / 16)], mma_dcol[8][kMaxGroupTaps * (KB / 16)];   // per group: B offset (16-byte units) / TMEM column of each MMA
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_strip_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ StripParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  const int nchunks_b = p.block_n / 64;
  const uint32_t a_stage = p.a_chunks * kBoxBytes, b_stage = (uint32_t)nchunks_b * p.strip_stride;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + p.stages * a_stage;
  const uint32_t zero_buf = smem_b + p.stages * b_stage;             // 8 KB of zeros: the absent second dy chunk when m <= 64
  const uint32_t bar_base = zero_buf + kBoxBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kMaxStages + 1);
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int m_tile = bid % p.num_m_tiles; bid /= p.num_m_tiles;
  const int n_tile = bid % p.num_n_tiles; bid /= p.num_n_tiles;
  const int grp = bid;
  const int64_t kb_per = (p.num_kblocks + p.splits - 1) / p.splits;
  const int64_t kb0 = split * kb_per;
  const int64_t kb1 = kb0 + kb_per < p.num_kblocks ? kb0 + kb_per : p.num_kblocks;
  const int ntaps = p.g_ntaps[grp];
  const bool has_work = kb1 > kb0;

  if (p.a_chunks == 1) {
    uint4* z = (uint4*)(smem_raw + (zero_buf - raw));
    for (int i = threadIdx.x; i < (int)(kBoxBytes / 16); i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (has_work) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = p.a_chunks * kBoxBytes + (uint32_t)nchunks_b * (uint32_t)p.strip_rows * 128u;
        const int ca = p.dy_koff + m_tile * BM, cb = p.x_koff + n_tile * p.block_n;
        int tix = (int)(kb0 % p.tiles_x), tiy = (int)((kb0 / p.tiles_x) % p.tiles_y), tin = (int)(kb0 / ((int64_t)p.tiles_x * p.tiles_y));
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), tx);
          const uint32_t sa = smem_a + stage * a_stage, sb = smem_b + stage * b_stage;
          const int cx = tix * p.bw, cy = tiy * p.bh, cn = tin * p.bn;      // counters instead of a 64-bit division per K block
          if (++tix == p.tiles_x) { tix = 0; if (++tiy == p.tiles_y) { tiy = 0; ++tin; } }
          for (uint32_t c = 0; c < p.a_chunks; ++c) ptx::tma_load_4d(sa + c * kBoxBytes, &tmDY, full_bar(stage), ca + (int)c * 64, cx, cy, cn);
          for (int c = 0; c < nchunks_b; ++c)
            ptx::tma_load_4d(sb + c * p.strip_stride, &tmX, full_bar(stage), cb + c * 64, cx + p.g_dxmin[grp], cy + p.g_dy[grp], cn);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // whole warp in the loop, one elected lane issues: the MMA operands stay in uniform registers (see gemm.cu).
      // B descriptors: per (tap, K step) an offset inside the strip (constant bank, host-built) added to the stage base.
      // Tap packing (npack): the dx taps of the group are equally spaced, so tap t of the strip is the SAME strip read `t * dxstep`
      // K-rows (pixels) further.  An MN-major B operand is a sequence of 64-column chunks `LBO` bytes apart: with LBO = dxstep * 128
      // bytes chunk t IS tap t, and one MMA of N = 64 * ntaps accumulates all taps of the filter row (A is read once, not per tap).
      const uint32_t idesc = ptx::make_idesc_bf16(BM, p.npack ? 64 * ntaps : p.block_n, 1, 1);
      const bool leader = ptx::elect_one();
      const uint64_t tmpl_b = ptx::make_smem_desc(0, p.npack ? (uint32_t)p.dxstep * 128u : p.strip_stride, 1024);
      const uint64_t tmpl_a = ptx::make_smem_desc(0, kBoxBytes, 1024);
      const int nmma = p.g_nmma[grp];
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_a + stage * a_stage, sb = smem_b + stage * b_stage;
        uint64_t adesc[KB / 16];
#pragma unroll
        for (int kk = 0; kk < KB / 16; ++kk) {
          // the zero chunk (a_chunks == 1) is shared by all K steps: its offset must not advance with kk
          adesc[kk] = p.a_chunks == 2 ? tmpl_a + (uint64_t)((sa + kk * 2048) >> 4) : ptx::make_smem_desc(sa + kk * 2048, zero_buf - sa - kk * 2048, 1024);
        }
        const uint64_t bbase = tmpl_b + (uint64_t)(sb >> 4);
        if (leader) {
          for (int i = 0; i < nmma; i += KB / 16) {
            const uint32_t td = tmem_base + p.mma_dcol[grp][i];
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)
              ptx::umma_bf16(td, adesc[kk], bbase + (uint64_t)p.mma_boff[grp][i + kk], idesc, kk > 0 ? 1u : accum);
          }
          ptx::umma_commit(empty_bar(stage));
        }
        accum = 1;
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (leader) ptx::umma_commit(tfull_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int i = m_tile * BM + q * 32 + lane;
      ptx::mbar_wait(tfull_bar, 0);
      ptx::tc_fence_after();
      const int n0 = n_tile * p.block_n;
      for (int ti = 0; ti < ntaps; ++ti) {
        float* row = p.dw + (int64_t)p.g_slot[grp][ti] * p.slot_stride + (int64_t)i * p.ld_dw;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          if (n0 + c0 >= p.n) break;
          uint32_t v[32];
          const int tcol = p.npack ? ((c0 >> 6) * ntaps + ti) * 64 + (c0 & 63) : ti * p.block_n + c0;
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tcol, v);
          ptx::tmem_ld_wait();
          if (i >= p.m) continue;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int col = n0 + c0 + g * 4;
            float* dst = row + col;
            float f0 = __uint_as_float(v[g * 4]) * p.alpha, f1 = __uint_as_float(v[g * 4 + 1]) * p.alpha,
                  f2 = __uint_as_float(v[g * 4 + 2]) * p.alpha, f3 = __uint_as_float(v[g * 4 + 3]) * p.alpha;
            if (col + 4 <= p.n && ((uintptr_t)dst & 15) == 0) {
              red_add_v4(dst, f0, f1, f2, f3);
            } else {
              if (col < p.n) atomicAdd(dst, f0);
              if (col + 1 < p.n) atomicAdd(dst + 1, f1);
              if (col + 2 < p.n) atomicAdd(dst + 2, f2);
              if (col + 3 < p.n) atomicAdd(dst + 3, f3);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Split-K factor: one CTA per SM is resident, so the kernel runs in ceil(tiles * s / SMs) rounds; a round costs its K blocks plus the
// partial tile's red.global.add epilogue.  Fitted on the ViT shapes (scratch/wgrad_splits.py: 54 tiles x 257 K blocks, t(s) = rounds *
// (K / s * 0.46 us + 9 us) within 5 %): the epilogue is worth ~20 K blocks, and beyond 8 partial sums per tile the same-address reductions
// of the splits that finish together start to serialise (9 tiles: s = 8 27 us, s = 16 38 us).
int pick_splits(int64_t tiles, int64_t num_kblocks) {
  const int64_t sms = num_sms();
  int64_t cap = num_kblocks / 8 > 0 ? num_kblocks / 8 : 1;
  int64_t smax = (3 * sms + tiles - 1) / tiles;
  if (smax > cap) smax = cap;
  if (smax < 1) smax = 1;
  int64_t best = 1;
  double best_cost = 1e30;
  for (int64_t s = 1; s <= smax; ++s) {
    const int64_t rounds = (tiles * s + sms - 1) / sms;
    const double cost = (double)rounds * ((double)num_kblocks / (double)s + 20.0 + 2.5 * (double)(s > 8 ? s - 8 : 0));
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return (int)best;
}

// Returns 1 when the strip kernel was launched, 0 when the problem does not qualify, < 0 on error.
int try_launch_strip(const svl_wgrad_desc* d, int block_n, cudaStream_t stream) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SVL_WGRAD_STRIP");
    mode = e ? atoi(e) : 2;      // 2: descriptor start address only (the 128B swizzle is a function of the absolute smem address; verified on B200);
                                 // 1: additionally set the descriptor base_offset field (wrong on B200, kept for experiments); 0: disabled
  }
  if (mode == 0 || !d->conv || d->num_taps < 2 || d->num_taps > SVL_MAX_TAPS) return 0;
  // rows at least one K block wide: a K block is 64 pixels of ONE image row, the last block of a row may hang over the edge (TMA
  // zero-fills dy and x there: no contribution) -- 641^2 / 801^2 crops give 164- / 204-pixel rows.  Narrower rows are packed several per
  // K block and must then be a multiple of the 16-pixel MMA K step.
  if (d->x_map_w != 0 || (d->w < KB && d->w % 16 != 0)) return 0;
  StripParams p;
  memset(&p, 0, sizeof(p));
  p.bw = d->w < KB ? d->w : KB;
  if (p.bw % 16 != 0) return 0;
  p.bh = d->h < KB / p.bw ? d->h : KB / p.bw;
  p.bn = d->nb < KB / (p.bw * p.bh) ? d->nb : KB / (p.bw * p.bh);
  if (p.bw * p.bh * p.bn != KB) return 0;
  // group taps by dy; every tap must be its own slot and share the column offsets
  int span = 0;
  for (int t = 0; t < d->num_taps; ++t) {
    if (d->tap_dy_koff[t] != d->tap_dy_koff[0] || d->tap_x_koff[t] != d->tap_x_koff[0] || d->tap_slot[t] != t) return 0;
    int g = -1;
    for (int k = 0; k < p.ngroups; ++k) if (p.g_dy[k] == d->tap_dy[t]) g = k;
    if (g < 0) {
      if (p.ngroups == 8) return 0;
      g = p.ngroups++;
      p.g_dy[g] = d->tap_dy[t];
      p.g_dxmin[g] = d->tap_dx[t];
    }
    if (p.g_ntaps[g] == kMaxGroupTaps) return 0;
    p.g_dx[g][p.g_ntaps[g]] = d->tap_dx[t];
    p.g_slot[g][p.g_ntaps[g]] = t;
    p.g_ntaps[g]++;
    if (d->tap_dx[t] < p.g_dxmin[g]) p.g_dxmin[g] = d->tap_dx[t];
  }
  int max_taps = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    for (int k = 0; k < p.g_ntaps[g]; ++k) span = span > p.g_dx[g][k] - p.g_dxmin[g] ? span : p.g_dx[g][k] - p.g_dxmin[g];
    max_taps = max_taps > p.g_ntaps[g] ? max_taps : p.g_ntaps[g];
  }
  // tap packing needs every group to hold the same number of taps, ascending and equally spaced
  p.npack = max_taps >= 2 && (block_n / 64 > 0 ? block_n / 64 : 1) * max_taps * 64 <= 512 && block_n % 64 == 0 ? 1 : 0;
  p.dxstep = p.g_ntaps[0] >= 2 ? p.g_dx[0][1] - p.g_dx[0][0] : 1;
  for (int g = 0; g < p.ngroups && p.npack; ++g) {
    if (p.g_ntaps[g] != max_taps) p.npack = 0;
    for (int k = 0; k < p.g_ntaps[g] && p.npack; ++k)
      if (p.g_dx[g][k] != p.g_dxmin[g] + k * p.dxstep) p.npack = 0;
  }
  if (p.dxstep < 1 || p.dxstep * 128 >= (1 << 18)) p.npack = 0;
  if (const char* e = getenv("SVL_WGRAD_PACK")) if (atoi(e) == 0) p.npack = 0;
  if ((!p.npack && max_taps * block_n > 512) || p.bw + span > 256) return 0;
  p.nb = d->nb; p.h = d->h; p.w = d->w;
  p.tiles_x = (d->w + p.bw - 1) / p.bw;
  p.tiles_y = (d->h + p.bh - 1) / p.bh;
  p.num_kblocks = (int64_t)p.tiles_x * p.tiles_y * ((d->nb + p.bn - 1) / p.bn);
  p.m = d->m; p.n = d->n; p.block_n = block_n;
  p.num_m_tiles = (d->m + BM - 1) / BM;
  p.num_n_tiles = (d->n + block_n - 1) / block_n;
  p.dy_koff = d->tap_dy_koff[0]; p.x_koff = d->tap_x_koff[0];
  p.strip_w = p.bw + span;
  p.strip_rows = p.strip_w * p.bh * p.bn;
  p.strip_stride = ((uint32_t)p.strip_rows * 128u + 1023u) & ~1023u;
  p.a_chunks = d->m <= 64 ? 1u : 2u;
  p.bo_mode = mode;
  for (int g = 0; g < p.ngroups; ++g) {
    const int outer = p.npack ? block_n / 64 : p.g_ntaps[g];      // packed: one MMA group per 64-channel chunk of x; else one per tap
    for (int o = 0; o < outer; ++o)
      for (int kk = 0; kk < KB / 16; ++kk) {
        const int prow = kk * 16;                  // 16 K-rows (pixels) per MMA; a K block may span several image rows of the strip
        const int shift = p.npack ? 0 : p.g_dx[g][o] - p.g_dxmin[g];
        const int srow = (prow / p.bw) * p.strip_w + (prow % p.bw) + shift;
        p.mma_boff[g][o * (KB / 16) + kk] = ((uint32_t)srow * 128u + (p.npack ? (uint32_t)o * p.strip_stride : 0u)) >> 4;
        p.mma_dcol[g][o * (KB / 16) + kk] = (uint32_t)(p.npack ? o * p.g_ntaps[g] * 64 : o * block_n);
      }
    p.g_nmma[g] = outer * (KB / 16);
  }
  p.dw = d->dw; p.ld_dw = d->ld_dw; p.slot_stride = d->slot_stride;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  const uint32_t stage_bytes = p.a_chunks * kBoxBytes + (uint32_t)(block_n / 64) * p.strip_stride;
  p.stages = (int)((kSmemBudget - kBoxBytes) / stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) return 0;
  p.tmem_cols = p.npack ? pow2ceil((block_n / 64) * max_taps * 64) : pow2ceil(max_taps * block_n < 32 ? 32 : max_taps * block_n);

  const int64_t dy_cols = d->dy_cols > 0 ? d->dy_cols : d->ld_dy;
  const int64_t x_cols = d->x_cols > 0 ? d->x_cols : d->ld_x;
  CUtensorMap tmDY, tmX;
  {
    uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    uint64_t dims[4] = {(uint64_t)dy_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t st[3] = {(uint64_t)d->ld_dy * 2, (uint64_t)d->ld_dy * 2 * d->w, (uint64_t)d->ld_dy * 2 * d->w * d->h};
    if (int rc = tma_encode_bf16(&tmDY, d->dy, 4, dims, st, box)) return rc;
    uint32_t boxx[4] = {64u, (uint32_t)p.strip_w, (uint32_t)p.bh, (uint32_t)p.bn};
    uint64_t dimsx[4] = {(uint64_t)x_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t stx[3] = {(uint64_t)d->ld_x * 2, (uint64_t)d->ld_x * 2 * d->w, (uint64_t)d->ld_x * 2 * d->w * d->h};
    if (int rc = tma_encode_bf16(&tmX, d->x, 4, dimsx, stx, boxx)) return rc;
  }
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles * p.ngroups;
  int splits = d->splits;
  if (splits <= 0) splits = pick_splits(tiles, p.num_kblocks);
  if (splits > p.num_kblocks) splits = (int)p.num_kblocks;
  p.splits = splits;
  const size_t smem = 1024 + (size_t)p.stages * stage_bytes + kBoxBytes + 8 * (2 * kMaxStages + 2) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SVL_CUDA(cudaFuncSetAttribute(wgrad_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  wgrad_strip_kernel<<<(unsigned)(tiles * splits), kThreads, smem, stream>>>(tmDY, tmX, p);
  SVL_LAUNCH_CHECK();
  return 1;
}


// ------------------------------------------------------------------------------------------------------------------------------
// Row-stacked variant for 3 x 3 (dilation 1) convolutions with 32 output channels -- the 32-channel `Up` convolutions of the decode head on
// 128-pixel-wide maps, whose weight gradients ran at 12.5 % useful tensor work in the strip kernel (one CTA group per filter row, each
// loading dy and x again) and 5x over their HBM time.  Here ONE step = 64 pixels of one x row r against the dy rows r-1, r, r+1:
//   A (M = 128) = two 64-element MN-major chunks = two dy rows; a 32-channel row fills the first 64 bytes of each 128-byte operand line (TMA
//                 pads a 64-byte inner box to the swizzle span; the other half is stale shared memory and only feeds accumulator lanes
//                 32..63 / 96..127, which nobody reads).  MMA 1: rows r-1 | r (filter rows +1 | 0), MMA 2: row r+1 | unused (filter row -1)
//   B (N = 192) = three dx taps: the x strip (66 pixels) read 0 / 1 / 2 pixels further (LBO = 128 bytes), 64 channels per chunk
// so x is loaded once (not once per filter row and CTA group), dy three times from L2, all nine taps come out of two MMAs per 16 pixels, and the
// whole problem is one output tile split over K (pixels) across the SMs.  Rows / pixels outside the image are TMA zero fill.
// (A [row a | row b] packing of two 32-channel rows into one operand line is not reachable with TMA: measured, a 64-byte inner box lands
// on a 128-byte pitch; scratch/wgrad_dump.py.)
struct RowstackParams {
  int nb, h, w, tiles_x, cin;
  int64_t num_steps;
  int dy_koff, x_koff, splits, stages;
  uint32_t strip_stride, stage_bytes;
  int slot[3][3];                  // dW slot of filter position (dy + 1, dx + 1)
  float* dw; int64_t ld_dw, slot_stride;
  float alpha;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_rowstack_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ RowstackParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kMaxStages + 1);
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t per = (p.num_steps + p.splits - 1) / p.splits;
  const int64_t s0 = (int64_t)blockIdx.x * per;
  const int64_t s1 = s0 + per < p.num_steps ? s0 + per : p.num_steps;
  const bool has_work = s1 > s0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDY);
    ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_addr, 512u);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (has_work) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = 3u * 64u * 64u + 66u * (uint32_t)p.cin * 2u;      // bytes the four boxes really transfer
        int tix = (int)(s0 % p.tiles_x), r = (int)((s0 / p.tiles_x) % p.h), img = (int)(s0 / ((int64_t)p.tiles_x * p.h));
        for (int64_t s = s0; s < s1; ++s) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(stage), tx);
          const uint32_t sa = smem_base + stage * p.stage_bytes, sb = sa + 4 * kBoxBytes;
          const int cx = tix * 64, cr = r, ci = img;
          if (++tix == p.tiles_x) { tix = 0; if (++r == p.h) { r = 0; ++img; } }
          for (int c = 0; c < 3; ++c) ptx::tma_load_4d(sa + c * kBoxBytes, &tmDY, full_bar(stage), p.dy_koff, cx, cr - 1 + c, ci);
          ptx::tma_load_4d(sb, &tmX, full_bar(stage), p.x_koff, cx - 1, cr, ci);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = ptx::make_idesc_bf16(BM, 192, 1, 1);
      const bool leader = ptx::elect_one();
      const uint64_t tmpl_a = ptx::make_smem_desc(0, kBoxBytes, 1024);     // M chunks (dy rows) one box apart
      const uint64_t tmpl_b = ptx::make_smem_desc(0, 128, 1024);           // N chunk t = the strip read t pixels (128-byte K rows) further
      int stage = 0;
      uint32_t phase = 0, accum = 0;
      for (int64_t s = s0; s < s1; ++s) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * p.stage_bytes, sb = sa + 4 * kBoxBytes;
        if (leader) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)
              ptx::umma_bf16(tmem_base + (uint32_t)a * 256u, tmpl_a + (uint64_t)((sa + a * 2 * kBoxBytes + kk * 2048) >> 4),
                             tmpl_b + (uint64_t)((sb + kk * 2048) >> 4), idesc, kk > 0 ? 1u : accum);
          }
          ptx::umma_commit(empty_bar(stage));
        }
        accum = 1;
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (leader) ptx::umma_commit(tfull_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;                            // TMEM lane quarter: 0 = first dy row of an MMA (channels 0..31), 2 = its second row
      ptx::mbar_wait(tfull_bar, 0);
      ptx::tc_fence_after();
      if (q == 0 || q == 2) {
        for (int a = 0; a < (q == 0 ? 2 : 1); ++a) {
          const int t = a == 0 ? (q == 0 ? 1 : 0) : -1;  // filter row: MMA 1 = dy rows r-1 | r -> +1 | 0, MMA 2 = dy row r+1 -> -1
          for (int dxi = 0; dxi < 3; ++dxi)
            for (int c0 = 0; c0 < p.cin; c0 += 32) {
              uint32_t v[32];
              ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + dxi * 64 + c0), v);
              ptx::tmem_ld_wait();
              float* dst = p.dw + (int64_t)p.slot[t + 1][dxi] * p.slot_stride + (int64_t)lane * p.ld_dw + c0;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                red_add_v4(dst + g * 4, __uint_as_float(v[g * 4]) * p.alpha, __uint_as_float(v[g * 4 + 1]) * p.alpha,
                           __uint_as_float(v[g * 4 + 2]) * p.alpha, __uint_as_float(v[g * 4 + 3]) * p.alpha);
            }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512u);
  }
}

// Returns 1 when the row-stacked kernel was launched, 0 when the problem does not qualify, < 0 on error.
int try_launch_rowstack(const svl_wgrad_desc* d, cudaStream_t stream) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SVL_WGRAD_ROWSTACK"); on = e ? atoi(e) : 1; }
  if (!on || !d->conv || d->num_taps != 9 || d->m != 32 || (d->n != 32 && d->n != 64) || d->x_map_w != 0 || d->w < 64) return 0;
  if (((uintptr_t)d->dw & 15) != 0 || d->ld_dw % 4 != 0 || d->slot_stride % 4 != 0) return 0;
  RowstackParams p;
  memset(&p, 0, sizeof(p));
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) p.slot[a][b] = -1;
  for (int t = 0; t < 9; ++t) {
    const int fy = d->tap_dy[t], fx = d->tap_dx[t];
    if (fy < -1 || fy > 1 || fx < -1 || fx > 1 || d->tap_dy_koff[t] != d->tap_dy_koff[0] || d->tap_x_koff[t] != d->tap_x_koff[0]) return 0;
    p.slot[fy + 1][fx + 1] = d->tap_slot[t];
  }
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) if (p.slot[a][b] < 0) return 0;
  p.nb = d->nb; p.h = d->h; p.w = d->w; p.cin = d->n;
  p.tiles_x = (d->w + 63) / 64;
  p.num_steps = (int64_t)d->nb * d->h * p.tiles_x;
  p.dy_koff = d->tap_dy_koff[0]; p.x_koff = d->tap_x_koff[0];
  p.strip_stride = (66u * 128u + 1023u) & ~1023u;
  p.stage_bytes = 4 * kBoxBytes + p.strip_stride;
  p.stages = (int)(kSmemBudget / p.stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.dw = d->dw; p.ld_dw = d->ld_dw; p.slot_stride = d->slot_stride;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  const int64_t dy_cols = d->dy_cols > 0 ? d->dy_cols : d->ld_dy;
  const int64_t x_cols = d->x_cols > 0 ? d->x_cols : d->ld_x;
  CUtensorMap tmDY, tmX;
  {
    uint64_t dims[4] = {(uint64_t)dy_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t st[3] = {(uint64_t)d->ld_dy * 2, (uint64_t)d->ld_dy * 2 * d->w, (uint64_t)d->ld_dy * 2 * d->w * d->h};
    uint32_t box[4] = {32u, 64u, 1u, 1u};
    if (int rc = tma_encode_bf16(&tmDY, d->dy, 4, dims, st, box)) return rc;
    uint64_t dimsx[4] = {(uint64_t)x_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t stx[3] = {(uint64_t)d->ld_x * 2, (uint64_t)d->ld_x * 2 * d->w, (uint64_t)d->ld_x * 2 * d->w * d->h};
    uint32_t boxx[4] = {(uint32_t)p.cin, 66u, 1u, 1u};
    if (int rc = tma_encode_bf16(&tmX, d->x, 4, dimsx, stx, boxx)) return rc;
  }
  int splits = d->splits > 0 ? d->splits : num_sms();
  if (splits > p.num_steps) splits = (int)p.num_steps;
  p.splits = splits;
  const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 8 * (2 * kMaxStages + 2) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SVL_CUDA(cudaFuncSetAttribute(wgrad_rowstack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  wgrad_rowstack_kernel<<<(unsigned)splits, kThreads, smem, stream>>>(tmDY, tmX, p);
  SVL_LAUNCH_CHECK();
  return 1;
}

}  // namespace
}  // namespace svl

using namespace svl;

extern "C" int svl_wgrad(const svl_wgrad_desc* d, void* stream) {
  SVL_CHECK_ARG(d && d->dy && d->x && d->dw, "svl_wgrad: null operand");
  SVL_CHECK_ARG(d->rows > 0 && d->m > 0 && d->n > 0, "svl_wgrad: empty problem");
  SVL_CHECK_ARG(d->num_taps >= 1 && d->num_taps <= SVL_MAX_TAPS, "svl_wgrad: num_taps=%d out of range", d->num_taps);
  SVL_CHECK_ARG(d->ld_dy % 8 == 0 && d->ld_x % 8 == 0, "svl_wgrad: leading dimensions must be multiples of 8 elements");
  SVL_CHECK_ARG(((uintptr_t)d->dy & 15) == 0 && ((uintptr_t)d->x & 15) == 0, "svl_wgrad: operands must be 16-byte aligned");
  if (int rc = svl_check_device()) return rc;

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.conv = d->conv;
  p.rows = d->rows;
  p.m = d->m;
  p.n = d->n;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.num_taps = d->num_taps;
  int nslots = 0;
  for (int t = 0; t < d->num_taps; ++t) {
    p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t];
    p.tap_dy_koff[t] = d->tap_dy_koff[t]; p.tap_x_koff[t] = d->tap_x_koff[t];
    int s = d->tap_slot[t];
    SVL_CHECK_ARG(s == nslots || s == nslots - 1, "svl_wgrad: tap slots must be contiguous and ascending (tap %d slot %d)", t, s);
    if (s == nslots) { p.slot_tap0[nslots] = t; ++nslots; }
  }
  p.slot_tap0[nslots] = d->num_taps;
  p.num_slots = nslots;
  p.block_n = d->n >= 256 ? 256 : (d->n + 63) / 64 * 64;
  if (d->n > 256) {
    int best = 256, best_pad = 1 << 30;
    for (int bn : {256, 192, 128}) {
      int pad = (d->n + bn - 1) / bn * bn - d->n;
      if (pad < best_pad) { best = bn; best_pad = pad; }
    }
    p.block_n = best;
  }
  {
    int rc = try_launch_rowstack(d, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : SVL_OK;
    rc = try_launch_strip(d, p.block_n, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : SVL_OK;
  }
  p.num_m_tiles = (d->m + BM - 1) / BM;
  p.num_n_tiles = (d->n + p.block_n - 1) / p.block_n;
  const int64_t dy_cols = d->dy_cols > 0 ? d->dy_cols : d->ld_dy;
  const int64_t x_cols = d->x_cols > 0 ? d->x_cols : d->ld_x;

  CUtensorMap tmDY, tmX;
  if (d->conv) {
    SVL_CHECK_ARG(d->nb > 0 && d->h > 0 && d->w > 0 && (int64_t)d->nb * d->h * d->w == d->rows, "svl_wgrad: conv geometry does not match rows");
    p.nb = d->nb; p.h = d->h; p.w = d->w;
    p.bw = d->w < KB ? d->w : KB;
    p.bh = d->h < KB / p.bw ? d->h : KB / p.bw;
    p.bn = d->nb < KB / (p.bw * p.bh) ? d->nb : KB / (p.bw * p.bh);
    p.tiles_x = (d->w + p.bw - 1) / p.bw;
    p.tiles_y = (d->h + p.bh - 1) / p.bh;
    p.num_kblocks = (int64_t)p.tiles_x * p.tiles_y * ((d->nb + p.bn - 1) / p.bn);
    p.k_tx_bytes = (uint32_t)(p.bw * p.bh * p.bn) * 128u;
    uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    uint64_t dims[4] = {(uint64_t)dy_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t st[3] = {(uint64_t)d->ld_dy * 2, (uint64_t)d->ld_dy * 2 * d->w, (uint64_t)d->ld_dy * 2 * d->w * d->h};
    if (int rc = tma_encode_bf16(&tmDY, d->dy, 4, dims, st, box)) return rc;
    const uint64_t xw = d->x_map_w > 0 ? d->x_map_w : d->w;
    uint64_t dimsx[4] = {(uint64_t)x_cols, xw, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t stx[3] = {(uint64_t)d->ld_x * 2, (uint64_t)d->ld_x * 2 * xw, (uint64_t)d->ld_x * 2 * xw * d->h};
    if (int rc = tma_encode_bf16(&tmX, d->x, 4, dimsx, stx, box)) return rc;
  } else {
    p.num_kblocks = (d->rows + KB - 1) / KB;
    p.k_tx_bytes = kBoxBytes;
    uint32_t box[2] = {64u, (uint32_t)KB};
    uint64_t dims[2] = {(uint64_t)dy_cols, (uint64_t)d->rows};
    uint64_t st[1] = {(uint64_t)d->ld_dy * 2};
    if (int rc = tma_encode_bf16(&tmDY, d->dy, 2, dims, st, box)) return rc;
    uint64_t dimsx[2] = {(uint64_t)x_cols, (uint64_t)d->rows};
    uint64_t stx[1] = {(uint64_t)d->ld_x * 2};
    if (int rc = tma_encode_bf16(&tmX, d->x, 2, dimsx, stx, box)) return rc;
  }
  const uint32_t stage_bytes = 2 * kBoxBytes + (uint32_t)(p.block_n / 64) * kBoxBytes;
  p.stages = (int)(kSmemBudget / stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.tmem_cols = pow2ceil(p.block_n < 32 ? 32 : p.block_n);
  p.dw = d->dw; p.ld_dw = d->ld_dw; p.slot_stride = d->slot_stride;

  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles * p.num_slots;
  int splits = d->splits;
  if (splits <= 0) splits = pick_splits(tiles, p.num_kblocks);
  if (splits > p.num_kblocks) splits = (int)p.num_kblocks;
  p.splits = splits;

  const size_t smem = 1024 + (size_t)p.stages * stage_bytes + 8 * (2 * kMaxStages + 2) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SVL_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int64_t grid = tiles * splits;
  SVL_CHECK_ARG(grid < (1ll << 31), "svl_wgrad: grid too large");
  wgrad_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(tmDY, tmX, p);
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
