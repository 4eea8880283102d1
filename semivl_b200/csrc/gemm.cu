// Tensor-core contraction engine: persistent, warp-specialised tcgen05 GEMM with tap-based implicit convolution.
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> SWIZZLE_128B smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + MMA issuer (one lane issues tcgen05.mma, tcgen05.commit frees smem stages)
//   warps 2..17 : epilogue (tcgen05.ld 32 lanes x 32 columns -> fused bias/act/act'/residual -> global); warp w owns TMEM lane
//                 quarter w % 4 and the 32-column chunks (w - 2) / 4 and (w - 2) / 4 + 4: a QUARTET = the four warps of one chunk.
//                 Staged variants (EXT 5 / 6): a quartet owns one shared-memory buffer; the side tile (residual, saved gelu') arrives by
//                 TMA, is combined in place and leaves by TMA store (profiles/r02_gemm_epilogue.md)
//
// Measured anatomy of a K = 768 tile (profiles/r02_gemm_epilogue.md): the main loop alone needs 7.0k cycles (1500 TFLOP/s), bf16 output
// stores add ~1.4k cycles to the MAIN LOOP whatever their pattern (library parity), the epilogues that move a second tile were latency
// chains of 12-19k cycles before the staged variants.
//
// Tile order: M fastest (all concurrent CTAs share one B tile) unless the A operand is too large to stay in L2 between column
// blocks; then N fastest: the concurrent CTAs cover (#SMs / n_tiles) row blocks x all column blocks, so each A row block comes from
// HBM once and L2 serves its other column blocks (with M fastest the K = 3072 FFN GEMMs streamed the 100 MB A operand once per
// column block: 350 MB of DRAM reads per launch instead of 155 MB, profiles/r01_ncu_targets.md; measured -9..-14 % on those
// launches, while the L2-resident QKV / FFN1 shapes lost 6-9 % with N fastest and keep M fastest).
// Accumulators are double-buffered in TMEM (2 x block_n columns) so the epilogue of tile i overlaps the main
// loop of tile i+1.  Tile = 128 output rows x block_n columns, K step 64 (one 128-byte swizzle row).
// In conv mode the 128 rows of a tile are a (bn images x bh rows x bw columns) box of an NHWC tensor, fetched with a
// 4-D tensor map; a filter tap only shifts the box coordinates and TMA's out-of-bounds zero fill is the padding.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.h"

namespace svl {
namespace {

#ifdef SVL_GEMM_DIAG
#define SVL_DBG(bit) (p.dbg & (bit))
#define SVL_TRACE(seq, slot) do { if (p.trace && blockIdx.x == 0 && (seq) < 32) p.trace[(seq) * 16 + (slot)] = clock64(); } while (0)
#else
#define SVL_DBG(bit) false
#define SVL_TRACE(seq, slot) do {} while (0)
#endif

#ifdef SVL_GEMM_DIAG
__device__ __forceinline__ void st16_cs(void* p, int64_t off, const float* f) {      // diag: streaming (evict-first) 32-byte bf16 store
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]); r[i] = *(uint32_t*)&h; }
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"((__nv_bfloat16*)p + off), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
#else
__device__ __forceinline__ void st16_cs(void*, int64_t, const float*) {}
#endif
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiGroups = 4;          // epilogue warps per TMEM lane quarter (each takes every kEpiGroups-th 32-column chunk)
constexpr int kThreads = 64 + 128 * kEpiGroups;   // warp 0 TMA, warp 1 MMA, then 4 * kEpiGroups epilogue warps
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 8;          // TMEM accumulator ring: as many block_n-wide slots as fit in 512 columns, at most this
constexpr int kSmemBudget = 200 * 1024;

struct GemmParams {
  // geometry
  int a_conv;
  int64_t m;
  int nb, h, w;
  int bw, bh, bn;                  // conv tile box
  int tiles_x, tiles_y;            // conv tiles per image row / column-of-tiles
  int num_m_tiles, num_n_tiles;
  int n_fastest;                   // tile order (see the note at the top)
  int staged;                      // staged epilogue (EXT 5 / 6): per-quartet smem buffers, side tile loaded and output tile stored by TMA
  int cluster;                     // 1: CTA pairs (cluster of 2 along M) share every B tile through TMA multicast;
                                   // 2: CTA pairs run ONE cta_group::2 UMMA (M = 256), each CTA holds half of the B tile
  int n, block_n, k_per_tap, num_taps;
  int stages, tmem_cols, acc_stages;
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes;
  int tap_dy[SVL_MAX_TAPS], tap_dx[SVL_MAX_TAPS], tap_a_koff[SVL_MAX_TAPS], tap_b_row[SVL_MAX_TAPS], tap_b_col[SVL_MAX_TAPS];
  // epilogue
  void* out; int out_dtype; int64_t ldc; int out_mode; int out_h, out_w;
  float alpha; const float* bias; const float* row_bias; int64_t row_bias_div, row_bias_ld;
  int act; void* preact_out; int preact_dtype; int64_t ld_preact;
  const void* dact_src; int dact_dtype; int dact_kind; int64_t ld_dact;
  const void* residual; int res_dtype; int64_t ldres;
  int accumulate;
  long long* trace;                // diag: per-tile clock64 log of CTA 0 ([tile][16] slots)
  int dbg;                         // -DSVL_GEMM_DIAG builds only (SVL_GEMM_DBG bits: 1 no global stores, 2 no epilogue body, 4 no side loads)
  // row-strip convolution mode (tile = 128 pixels of one image row, k_per_tap <= 64, one N tile): the taps of a filter row share one
  // A strip of (128 + span) pixels, read by each tap at its own row offset; all tap weights stay resident in shared memory
  int strip, ng;
  int g_dy[8], g_dxmin[8], g_ntaps[8], g_tap[8][4];
  int g_nmma[4];                   // MMAs per strip group, and per MMA the A / B descriptor offsets (16-byte units)
  uint32_t mma_aoff[4][16], mma_boff[4][16];
  uint32_t strip_bytes, b_tile_bytes;
};

// Position of tile row r inside a conv tile box (tile-invariant: computed once per thread, the divisions are not repeated per tile)
struct RowInBox { int ix, iy, in; };
__device__ __forceinline__ RowInBox row_in_box(const GemmParams& p, int r) {
  RowInBox b = {0, 0, 0};
  if (p.a_conv) { b.ix = r % p.bw; b.iy = (r / p.bw) % p.bh; b.in = r / (p.bw * p.bh); }
  return b;
}
// global row index of tile row r, or -1 when the row is padding
__device__ __forceinline__ int64_t tile_row_to_global(const GemmParams& p, int m_tile, int r, const RowInBox& b) {
  if (!p.a_conv) {
    int64_t g = (int64_t)m_tile * BM + r;
    return g < p.m ? g : -1;
  }
  const int t2 = m_tile / p.tiles_x;
  const int tx = m_tile - t2 * p.tiles_x;
  const int tn = t2 / p.tiles_y;
  const int ty = t2 - tn * p.tiles_y;
  int x = tx * p.bw + b.ix, y = ty * p.bh + b.iy, img = tn * p.bn + b.in;
  if (b.in >= p.bn || x >= p.w || y >= p.h || img >= p.nb) return -1;
  return ((int64_t)img * p.h + y) * p.w + x;
}

// CONVT2X2: element offset (relative to output pixel (2y, 2x), channel 0) of column `col`: quadrant q = col / cq goes to pixel
// (2y + q / 2, 2x + q % 2), channel col % cq
__device__ __forceinline__ int64_t convt_quadrant_offset(const GemmParams& p, int col, int cq) {
  const int qq = col / cq, cc = col - qq * cq;
  return ((int64_t)(qq >> 1) * (2 * p.out_w) + (qq & 1)) * p.ldc + cc;
}

// Epilogue specialisations: the fully generic epilogue is ~8k instructions and starves the instruction cache (ncu: stall_no_inst
// dominated, profiles/r01_gemm_epilogue.md), so the hot combinations are compiled with everything else pruned.
//   OUT  : svl_dtype of the output            ACT : svl_act applied after bias
//   PRE  : dtype of the saved pre-activation or -1         DACT: 0 none, 1 GELU' from a bf16 pre-activation
//   RES  : dtype of the residual or -1        GEN : runtime-generic epilogue (all features, everything a runtime branch)
//   EXT  : 0 linear output; 1 ConvT 2x2 scatter; 2 per-row-block bias (row_bias); 3 accumulate into an F32 output;
//          5 / 6 staged epilogues (per-quartet shared-memory buffers, TMA side-tile load and TMA store): F32 + F32 residual / BF16
template <int OUT, int ACT, int PRE, int DACT, int RES, int EXT, bool GEN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBh,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A][stages x B][barriers][tmem ptr]
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + p.stages * p.a_stage_bytes;
  // strip mode: smem_b holds the resident weights of all taps (num_taps tiles), the A ring holds strips
  // staged epilogue: one buffer per epilogue quartet (the 4 warps that cover the 128 rows of one 32-column chunk): 128 rows x 32 columns,
  // f32 (16 KB, SWIZZLE_128B) or bf16 (8 KB, SWIZZLE_64B), after the operand ring
  const uint32_t smem_c = smem_b + (p.strip ? (uint32_t)p.num_taps * p.b_tile_bytes : p.stages * p.b_stage_bytes);   // 1024-aligned
  const uint32_t bar_base = smem_c + (uint32_t)p.staged * kEpiGroups;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + s); };
  const uint32_t wbar = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc + 1);
  auto side_bar = [&](int g) { return bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc + 2 + g); };
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks_per_tap = (p.k_per_tap + BK - 1) / BK;
  // Cluster mode (plain GEMMs): the two CTAs of a cluster take the row blocks 2s and 2s + 1 of the same column block in lockstep; each
  // loads its own A tile and HALF of the B tile, multicast into both CTAs, so a CTA pulls 32 KB instead of 48 KB per k-block through
  // L2 -> SMEM.  A stage is free when BOTH CTAs have
  // consumed it (the peer's multicast writes into this CTA's copy), hence the 2-arrival empty barriers and the multicast commits.
  const int crank = p.cluster ? (int)ptx::cluster_ctarank() : 0;
  constexpr bool pair = PAIR;                  // cta_group::2 code only exists in the PAIR instantiations (they must be launched as clusters)
  const int m_slots = p.cluster ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
  const int num_tiles = m_slots * p.num_n_tiles;
  const int tile0 = p.cluster ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstride = p.cluster ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), p.cluster == 1 ? 2 : 1);
    }
    for (int s = 0; s < p.acc_stages; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      // pair mode: the epilogue warps of BOTH CTAs release the accumulator stage on the leader's barrier
      ptx::mbar_init(tempty_bar(s), (pair ? 2 : 1) * 4 * (p.block_n < kEpiGroups * 32 ? (p.block_n + 31) / 32 : kEpiGroups));
    }
    ptx::mbar_init(wbar, 1);
    for (int g = 0; g < kEpiGroups; ++g) ptx::mbar_init(side_bar(g), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (pair) {
      ptx::tmem_alloc_pair(tmem_ptr_addr, (uint32_t)p.tmem_cols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster) ptx::cluster_sync();          // the peer's barriers are initialised before any multicast can reach them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.strip) {
        ptx::mbar_arrive_expect_tx(wbar, (uint32_t)p.num_taps * p.b_tile_bytes);
        for (int t = 0; t < p.num_taps; ++t) ptx::tma_load_2d(smem_b + t * p.b_tile_bytes, &tmB, wbar, p.tap_b_col[t], p.tap_b_row[t]);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int cx = (tile % p.tiles_x) * p.bw, cy = (tile / p.tiles_x) % p.tiles_y, cn = tile / (p.tiles_x * p.tiles_y);
          for (int g = 0; g < p.ng; ++g) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
            ptx::mbar_arrive_expect_tx(full_bar(stage), p.strip_bytes);
            ptx::tma_load_4d(smem_a + stage * p.a_stage_bytes, &tmA, full_bar(stage), p.tap_a_koff[0], cx + p.g_dxmin[g], cy + p.g_dy[g], cn);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      } else
      for (int tile = tile0; tile < num_tiles; tile += tstride) {
        const int n_tile = p.n_fastest ? tile % p.num_n_tiles : tile / m_slots, m_slot = p.n_fastest ? tile / p.num_n_tiles : tile % m_slots;
        const int m_tile = p.cluster ? 2 * m_slot + crank : m_slot;      // may be one past the last row block: TMA zero-fills, the epilogue skips
        const int n0 = n_tile * p.block_n;
        int cx = 0, cy = 0, cn = 0;
        if (p.a_conv) {
          cx = (m_tile % p.tiles_x) * p.bw;
          cy = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
          cn = (m_tile / (p.tiles_x * p.tiles_y)) * p.bn;
        }
        for (int t = 0; t < p.num_taps; ++t) {
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_a + stage * p.a_stage_bytes, sb = smem_b + stage * p.b_stage_bytes;
            if (pair) {
              // both CTAs' tiles complete on the LEADER's barrier (only its MMA warp waits); the leader arms it with the bytes of the pair
              if (crank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2u * (p.a_tx_bytes + p.b_stage_bytes));
              const uint32_t lead_bar = ptx::cluster_map(full_bar(stage), 0);
              ptx::tma_load_2d_pair(sa, &tmA, lead_bar, p.tap_a_koff[t] + kb * BK, m_tile * BM);
              ptx::tma_load_2d_pair(sb, &tmBh, lead_bar, p.tap_b_col[t] + kb * BK, p.tap_b_row[t] + n0 + crank * (p.block_n >> 1));
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
              continue;
            }
            ptx::mbar_arrive_expect_tx(full_bar(stage), p.a_tx_bytes + p.b_stage_bytes);
            if (p.a_conv)
              ptx::tma_load_4d(sa, &tmA, full_bar(stage), p.tap_a_koff[t] + kb * BK, cx + p.tap_dx[t], cy + p.tap_dy[t], cn);
            else
              ptx::tma_load_2d(sa, &tmA, full_bar(stage), p.tap_a_koff[t] + kb * BK, m_tile * BM);
            if (p.cluster == 1)
              ptx::tma_load_2d_multicast(sb + (uint32_t)crank * (p.b_stage_bytes >> 1), &tmBh, full_bar(stage), p.tap_b_col[t] + kb * BK,
                                         p.tap_b_row[t] + n0 + crank * (p.block_n >> 1), (uint16_t)3);
            else
              ptx::tma_load_2d(sb, &tmB, full_bar(stage), p.tap_b_col[t] + kb * BK, p.tap_b_row[t] + n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and one elected lane issues: every operand of tcgen05.mma (descriptors, TMEM address, instruction
    // descriptor, accumulate flag) then stays in UNIFORM registers.  Issued from a divergent `lane == 0` branch the same code needs
    // ~30 instructions per MMA (vector->uniform moves inside a waterfall loop) and small-N convolutions become bound by this one
    // thread (profiles/r01_up2_conv_issue_bound.md).  The strip mode reads its per-tile MMA list (A offset inside the strip,
    // B offset inside the resident weights) from a host-built table in the constant bank.
    const uint32_t idesc = ptx::make_idesc_bf16(pair ? 2 * BM : BM, p.block_n, 0, 0);
    const bool leader = ptx::elect_one();
    const uint64_t tmpl = ptx::make_smem_desc(0, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    if (p.strip) {
      ptx::mbar_wait(wbar, 0);
      ptx::tc_fence_after();
      const uint32_t b_lo = smem_b >> 4;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.block_n);
        uint32_t accum = 0;
        for (int g = 0; g < p.ng; ++g) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = (smem_a + stage * p.a_stage_bytes) >> 4;
          const int nm = p.g_nmma[g];
          if (leader) {
            for (int i = 0; i < nm; ++i) {
              ptx::umma_bf16(tmem_d, tmpl + (uint64_t)(a_lo + p.mma_aoff[g][i]), tmpl + (uint64_t)(b_lo + p.mma_boff[g][i]), idesc, accum);
              accum = 1;
            }
            ptx::umma_commit(empty_bar(stage));
          }
          accum = 1;
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) ptx::umma_commit(tfull_bar(as));
        __syncwarp();
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      }
    } else {
      int tseq = 0;
      for (int tile = (pair && crank != 0) ? num_tiles : tile0; tile < num_tiles; tile += tstride, ++tseq) {      // pair mode: the leader issues for both SMs
        if (leader) SVL_TRACE(tseq, 0);
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        if (leader) SVL_TRACE(tseq, 1);
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.block_n);
        uint32_t accum = 0;
#ifdef SVL_GEMM_DIAG
        long long waited = 0;
#endif
        for (int t = 0; t < p.num_taps; ++t) {
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
#ifdef SVL_GEMM_DIAG
            const long long w0 = clock64();
#endif
            ptx::mbar_wait(full_bar(stage), phase);
#ifdef SVL_GEMM_DIAG
            waited += clock64() - w0;
#endif
            ptx::tc_fence_after();
            const uint64_t adesc = tmpl + (uint64_t)((smem_a + stage * p.a_stage_bytes) >> 4);
            const uint64_t bdesc = tmpl + (uint64_t)((smem_b + stage * p.b_stage_bytes) >> 4);
            const int nk = min(BK, p.k_per_tap - kb * BK) / 16;
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < nk) {
                  if (pair) ptx::umma_bf16_pair(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk == 0 ? accum : 1u);
                  else ptx::umma_bf16(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, kk == 0 ? accum : 1u);
                }
              }
              if (pair) ptx::umma_commit_pair(empty_bar(stage), (uint16_t)3);
              else if (p.cluster) ptx::umma_commit_multicast(empty_bar(stage), (uint16_t)3);
              else ptx::umma_commit(empty_bar(stage));
            }
            accum = 1;
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
        if (leader) {
          if (pair) ptx::umma_commit_pair(tfull_bar(as), (uint16_t)3);      // both CTAs' epilogue warps read their half of the accumulator
          else ptx::umma_commit(tfull_bar(as));
          SVL_TRACE(tseq, 2);
#ifdef SVL_GEMM_DIAG
          if (p.trace && blockIdx.x == 0 && tseq < 32) p.trace[tseq * 16 + 15] = waited;
#endif
        }
        __syncwarp();
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int cgrp = (warp - 2) >> 2;       // which interleaved subset of the 32-column chunks
    const int r = q * 32 + lane;            // tile row of this thread
    const float alpha = p.alpha == 0.f ? 1.f : p.alpha;
    const int cq = p.out_mode == SVL_OUT_CONVT2X2 ? p.n / 4 : 0;
    // one column block (the Up blocks' transposed convolutions): the quadrant offsets of this thread's four 16-column groups do not depend
    // on the tile -- the two divisions per group were 40 % of the epilogue's instructions, and with one k-block per tile the epilogue IS the kernel
    int64_t ct_qoff[2][2] = {{0, 0}, {0, 0}};
    if (EXT == 1 && cq && p.num_n_tiles == 1) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = (((warp - 2) >> 2) + j * kEpiGroups) * 32 + g * 16;
          if (col < p.n) ct_qoff[j][g] = convt_quadrant_offset(p, col, cq);
        }
    }
    int as = 0;
    uint32_t aphase = 0;
    const RowInBox rib = row_in_box(p, r);
    // column groups without a 32-column chunk of this block_n (narrow conv outputs) take no part: the accumulator-free barrier counts
    // only the active groups
    const bool active = cgrp * 32 < p.block_n;
    // staged epilogue (EXT 5: f32 out + f32 residual; EXT 6: bf16 out, optional act' side tile / second output): the quartet = the four
    // warps (one per TMEM lane quarter) that share a 32-column chunk; its buffer holds that chunk for all 128 rows of the tile
    constexpr bool STAGED = !GEN && (EXT == 5 || EXT == 6);
    constexpr bool ST_F32 = EXT == 5;
    constexpr bool ST_SIDE = STAGED && (ST_F32 || DACT >= 1);
    constexpr uint32_t kStBuf = ST_F32 ? 16384u : 8192u, kStRow = ST_F32 ? 128u : 64u;
    const uint32_t st_buf = smem_c + (uint32_t)cgrp * kStBuf;
    uint8_t* const st_row = smem_raw + (st_buf - ptx::smem_u32(smem_raw)) + (uint32_t)r * kStRow;
    const uint32_t st_sw = ST_F32 ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);      // SWIZZLE_128B / SWIZZLE_64B chunk XOR of this row
    const bool qlead = q == 0 && lane == 0;                                          // the quartet's TMA thread
    uint32_t sphase = 0;
    auto quartet_bar = [&]() { __syncwarp(); asm volatile("bar.sync %0, 128;" ::"r"(1 + cgrp) : "memory"); };
    int tseq = -1;
    const int tslot = warp == 2 ? 4 : warp == 9 ? 8 : warp == 17 ? 12 : 100;
    for (int tile = active ? (p.strip ? (int)blockIdx.x : tile0) : num_tiles; tile < num_tiles; tile += (p.strip ? (int)gridDim.x : tstride)) {
      ++tseq;
      if (lane == 0 && tslot < 16) SVL_TRACE(tseq, tslot);
      const int n_tile = p.n_fastest ? tile % p.num_n_tiles : tile / m_slots, m_slot = p.n_fastest ? tile / p.num_n_tiles : tile % m_slots;
      const int m_tile = p.cluster ? 2 * m_slot + crank : m_slot;
      const int n0 = n_tile * p.block_n;
      const int64_t grow = tile_row_to_global(p, m_tile, r, rib);
      int64_t ct_base = 0;                  // CONVT2X2: element offset of output pixel (2y, 2x), channel 0
      if (cq && grow >= 0) {
        // 32-bit index arithmetic when the row index allows it: three 64-bit divisions per thread and tile were a third of the tile time of
        // the K = 64 transposed-convolution GEMMs (one k-block per tile: the epilogue IS the kernel)
        int64_t img;
        int y, x;
        if (p.m < (1ll << 31)) {
          const uint32_t hw = (uint32_t)(p.out_h * p.out_w), g32 = (uint32_t)grow;
          const uint32_t im = g32 / hw, rem = g32 - im * hw;
          y = (int)(rem / (uint32_t)p.out_w);
          x = (int)(rem - (uint32_t)y * (uint32_t)p.out_w);
          img = im;
        } else {
          const int64_t hw = (int64_t)p.out_h * p.out_w;
          img = grow / hw;
          y = (int)((grow % hw) / p.out_w);
          x = (int)(grow % p.out_w);
        }
        ct_base = ((img * 2 * p.out_h + 2 * y) * (2 * p.out_w) + 2 * x) * p.ldc;
      }
      // Side inputs of the specialised epilogues (act' source, residual) do not depend on the accumulator: fetch them while the
      // tile's MMAs are still running.  act' source: both 32-column chunks of this thread into registers (raw bf16);
      // residual (fp32, 128 B per chunk): an L2 prefetch of the line.
      uint4 sraw[2][4];
      if (!GEN && EXT < 5 && DACT >= 1 && grow >= 0 && !SVL_DBG(4)) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c0 = (cgrp + j * kEpiGroups) * 32;
          if (c0 < p.block_n && n0 + c0 < p.n) {       // launcher: n % 32 == 0 for these variants, so a started chunk is whole
            const uint4* src = (const uint4*)((const __nv_bfloat16*)p.dact_src + grow * p.ld_dact + n0 + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) sraw[j][i] = __ldg(src + i);
          }
        }
      }
      if (!GEN && EXT < 5 && RES == SVL_F32 && grow >= 0 && !SVL_DBG(4)) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c0 = (cgrp + j * kEpiGroups) * 32;
          if (c0 < p.block_n && n0 + c0 < p.n)
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const float*)p.residual + grow * p.ldres + n0 + c0));
        }
      }
      if (ST_SIDE && qlead && n0 + cgrp * 32 < p.n) {
        // side tile of the quartet's first chunk: the buffer is free once this thread's earlier stores have read it
        ptx::bulk_wait_group_read0();
        ptx::mbar_arrive_expect_tx(side_bar(cgrp), kStBuf);
        ptx::tma_load_2d(st_buf, &tmD, side_bar(cgrp), n0 + cgrp * 32, m_tile * BM);
        // the side tiles come from HBM (saved activations / the residual stream): 3-4k cycles of load latency per chunk sat on the quartet's
        // critical path; pull the NEXT tile's two chunks into L2 now (profiles/r02_gemm_epilogue.md)
        const int nt = tile + tstride;
        if (nt < num_tiles && !SVL_DBG(128)) {
          const int nn = p.n_fastest ? nt % p.num_n_tiles : nt / m_slots, ns = p.n_fastest ? nt / p.num_n_tiles : nt % m_slots;
          const int nm = p.cluster ? 2 * ns + crank : ns;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int cc = (cgrp + jj * kEpiGroups) * 32;
            if (cc < p.block_n && nn * p.block_n + cc < p.n) ptx::tma_prefetch_2d(&tmD, nn * p.block_n + cc, nm * BM);
          }
        }
      }
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      if (lane == 0 && tslot < 16) SVL_TRACE(tseq, tslot + 1);
      if (GEN)
      for (int c0 = cgrp * 32; c0 < p.block_n; c0 += kEpiGroups * 32) {
        if (n0 + c0 >= p.n) break;
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n + c0), v);
        ptx::tmem_ld_wait();
        if (GEN) {
          if (grow < 0) continue;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            const int col = n0 + c0 + g * 8;
            const int cnt = min(8, p.n - col);
            if (cnt <= 0) break;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]) * alpha;
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (i < cnt) f[i] += __ldg(p.bias + col + i);
            }
            if (p.row_bias) {
              const float* rb = p.row_bias + (grow / p.row_bias_div) * p.row_bias_ld + col;
#pragma unroll
              for (int i = 0; i < 8; ++i) if (i < cnt) f[i] += __ldg(rb + i);
            }
            if (p.act == SVL_ACT_GELU_DSAVE) {
              float dg[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { dg[i] = gelu_grad(f[i]); f[i] = gelu_exact(f[i]); }
              st8(p.preact_out, p.preact_dtype, grow * p.ld_preact + col, 0, cnt, dg);
            } else if (p.preact_out) st8(p.preact_out, p.preact_dtype, grow * p.ld_preact + col, 0, cnt, f);
            if (p.act == SVL_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = gelu_exact(f[i]);
            } else if (p.act == SVL_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.dact_src) {
              float s[8];
              ld8(p.dact_src, p.dact_dtype, grow * p.ld_dact + col, p.ld_dact / 2, cnt, s);
              if (p.dact_kind == SVL_ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= gelu_grad(s[i]);
              } else if (p.dact_kind == SVL_ACT_SAVED) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= s[i];
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = s[i] > 0.f ? f[i] : 0.f;
              }
            }
            if (p.residual) {
              float s[8];
              ld8(p.residual, p.res_dtype, grow * p.ldres + col, 0, cnt, s);
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] += s[i];
            }
            int64_t off;
            if (cq) {
              const int qq = col / cq, cc = col % cq;       // cq % 8 == 0: an 8-group never straddles a quadrant
              off = ct_base + ((int64_t)(qq >> 1) * (2 * p.out_w) + (qq & 1)) * p.ldc + cc;
            } else {
              off = grow * p.ldc + col;
            }
            if (p.accumulate) {
              float s[8];
              ld8(p.out, SVL_F32, off, 0, cnt, s);
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] += s[i];
            }
            st8(p.out, p.out_dtype, off, p.ldc / 2, cnt, f);
          }
        } else {
          // (unreachable: the specialised kernels use the slab loop below)
        }
      }
      bool released = false;
      auto release_acc = [&]() {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair && crank != 0) ptx::mbar_arrive_cluster(ptx::cluster_map(tempty_bar(as), 0));
          else ptx::mbar_arrive(tempty_bar(as));
        }
        released = true;
      };
      if (STAGED) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c0 = (cgrp + j * kEpiGroups) * 32;
          if (c0 >= p.block_n || n0 + c0 >= p.n) break;
          const int c1 = c0 + kEpiGroups * 32;
          const bool last = j == 1 || c1 >= p.block_n || n0 + c1 >= p.n;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n + c0), v);
          if (ST_SIDE) {
            ptx::mbar_wait(side_bar(cgrp), sphase);                  // the side tile has landed (and with it: the buffer was free)
            sphase ^= 1u;
          } else {
            if (qlead) ptx::bulk_wait_group_read0();                 // the previous store has read the buffer
            quartet_bar();
          }
          ptx::tmem_ld_wait();
          if (warp == 2 && lane == 0) SVL_TRACE(tseq, j == 0 ? 3 : 11);
          if (last) release_acc();                                   // the accumulator is in registers: the MMA warp may overwrite it now
          uint4 second[4];                                           // GELU_DSAVE: gelu' of the chunk, packed, stored after the first output
          if (ST_F32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4* cell = (float4*)(st_row + (((uint32_t)i ^ st_sw) << 4));
              float4 o = *cell;                                      // residual
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) bb = __ldg((const float4*)(p.bias + n0 + c0) + i);
              o.x += fmaf(__uint_as_float(v[4 * i]), alpha, bb.x);
              o.y += fmaf(__uint_as_float(v[4 * i + 1]), alpha, bb.y);
              o.z += fmaf(__uint_as_float(v[4 * i + 2]), alpha, bb.z);
              o.w += fmaf(__uint_as_float(v[4 * i + 3]), alpha, bb.w);
              *cell = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4* cell = (uint4*)(st_row + (((uint32_t)i ^ st_sw) << 4));
              float f[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * i + k]);
              if (alpha != 1.f) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] *= alpha;
              }
              if (p.bias) {
                const float4 b0 = __ldg((const float4*)(p.bias + n0 + c0) + 2 * i), b1 = __ldg((const float4*)(p.bias + n0 + c0) + 2 * i + 1);
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (ACT == SVL_ACT_GELU_DSAVE) {
                float dg[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) gelu_and_grad_fast(f[k], f[k], dg[k]);
                second[i] = f32_to_bf16x8(dg);
              } else if (ACT == SVL_ACT_GELU) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = gelu_fast(f[k]);
              }
              if (DACT >= 1) {
                const uint4 sd = *cell;
                const uint32_t w[4] = {sd.x, sd.y, sd.z, sd.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float s0 = __uint_as_float(w[k] << 16), s1 = __uint_as_float(w[k] & 0xffff0000u);
                  f[2 * k] *= DACT == 1 ? gelu_grad_fast(s0) : s0;
                  f[2 * k + 1] *= DACT == 1 ? gelu_grad_fast(s1) : s1;
                }
              }
              if (!SVL_DBG(64)) *cell = f32_to_bf16x8(f);
            }
          }
          ptx::fence_proxy_async();                                  // generic-proxy writes of the buffer -> visible to the TMA engine
          quartet_bar();
          if (warp == 2 && lane == 0 && j == 0) SVL_TRACE(tseq, 7);
          if (qlead && !SVL_DBG(1)) {
            ptx::tma_store_2d(&tmC, st_buf, n0 + c0, m_tile * BM);   // rows >= M are clipped by the tensor map
            ptx::bulk_commit_group();
          }
          if (ACT == SVL_ACT_GELU_DSAVE) {
            if (qlead) ptx::bulk_wait_group_read0();
            quartet_bar();
#pragma unroll
            for (int i = 0; i < 4; ++i) if (!SVL_DBG(64)) *(uint4*)(st_row + (((uint32_t)i ^ st_sw) << 4)) = second[i];
            ptx::fence_proxy_async();
            quartet_bar();
            if (qlead && !SVL_DBG(1)) {
              ptx::tma_store_2d(&tmD, st_buf, n0 + c0, m_tile * BM);
              ptx::bulk_commit_group();
            }
          }
          if (ST_SIDE && !last && qlead) {                           // side tile of the quartet's second chunk
            ptx::bulk_wait_group_read0();
            ptx::mbar_arrive_expect_tx(side_bar(cgrp), kStBuf);
            ptx::tma_load_2d(st_buf, &tmD, side_bar(cgrp), n0 + c1, m_tile * BM);
          }
        }
        __syncwarp();
      }
      if (!GEN && !STAGED && !SVL_DBG(2)) {
        // specialised epilogue: 32-column chunks, one output row per thread, 32-byte vector accesses (full sectors);
        // bias / act / act' / residual fused in registers
#pragma unroll
        for (int j = 0; j < 2; ++j) {                  // block_n <= 256: at most two chunks per warp
          const int c0 = (cgrp + j * kEpiGroups) * 32;
          if (c0 >= p.block_n || n0 + c0 >= p.n) break;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n + c0), v);
          float side[2][16];
          if (RES >= 0 && grow >= 0 && !SVL_DBG(4)) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = n0 + c0 + g * 16;
              const int cnt = min(16, p.n - col);
              if (cnt > 0) ld16(p.residual, RES, grow * p.ldres + col, 0, cnt, side[g]);
            }
          }
          if (DACT >= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t w = ((const uint32_t*)&sraw[j][0])[i];
              side[i >> 3][(2 * i) & 15] = __uint_as_float(w << 16);
              side[i >> 3][((2 * i) & 15) + 1] = __uint_as_float(w & 0xffff0000u);
            }
          }
          ptx::tmem_ld_wait();
          if (grow < 0) continue;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int col = n0 + c0 + g * 16;
            const int cnt = min(16, p.n - col);
            if (cnt <= 0) break;
            float f[16];
            if (alpha != 1.f) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[g * 16 + i]) * alpha;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[g * 16 + i]);
            }
            if (p.bias) {
              if (cnt == 16 && (((uintptr_t)(p.bias + col)) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 bb = __ldg((const float4*)(p.bias + col) + j);
                  f[4 * j] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (i < cnt) f[i] += __ldg(p.bias + col + i);
              }
            }
            if (EXT == 2) {
              const float* rb = p.row_bias + (grow / p.row_bias_div) * p.row_bias_ld + col;
#pragma unroll
              for (int i = 0; i < 16; ++i) if (i < cnt) f[i] += __ldg(rb + i);
            }
            if (ACT == SVL_ACT_GELU_DSAVE) {                 // out = gelu(z), preact_out = gelu'(z): the backward epilogue is one multiply
              float dg[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) gelu_and_grad_fast(f[i], f[i], dg[i]);
              if (!SVL_DBG(1)) { if (SVL_DBG(16)) st16_cs(p.preact_out, (SVL_DBG(8) ? (grow & 1023) : grow) * p.ld_preact + col, dg); else st16(p.preact_out, PRE, (SVL_DBG(8) ? (grow & 1023) : grow) * p.ld_preact + col, 0, cnt, dg); }
            } else if (PRE >= 0) st16(p.preact_out, PRE, grow * p.ld_preact + col, 0, cnt, f);
            if (ACT == SVL_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = gelu_fast(f[i]);
            } else if (ACT == SVL_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (DACT == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] *= gelu_grad_fast(side[g][i]);
            } else if (DACT == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] *= side[g][i];
            }
            if (RES >= 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += side[g][i];
            }
            int64_t off = (SVL_DBG(8) ? (grow & 1023) : grow) * p.ldc + col;
            if (SVL_DBG(32)) off = ((((int64_t)tile * 2 + j) * 2 + g) * 16 + (warp - 2)) * 512 + lane * 16;      // diag: every store instruction covers 1 KB contiguous
            if (EXT == 1) {                                  // cq % 16 == 0 is checked by the launcher for this variant
              off = ct_base + (p.num_n_tiles == 1 ? ct_qoff[j][g] : convt_quadrant_offset(p, col, cq));
            }
            if (EXT == 3) {
              float sv[16];
              ld16(p.out, SVL_F32, off, 0, cnt, sv);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += sv[i];
            }
            if (!SVL_DBG(1)) {
              if (OUT == SVL_BF16 && SVL_DBG(16)) st16_cs(p.out, off, f); else
              st16(p.out, OUT, off, p.ldc / 2, cnt, f);
            }
          }
        }
      }
      if (lane == 0 && tslot < 16) SVL_TRACE(tseq, tslot + 2);
      if (!released) release_acc();
      if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
    }
  }

  if (p.staged && warp >= 2 && (warp & 3) == 0 && lane == 0) ptx::bulk_wait_group0();      // the quartet leaders' last stores are complete before the CTA retires
  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster) ptx::cluster_sync();          // no CTA leaves while its peer may still signal into its shared memory
  if (warp == 1) {
    ptx::tc_fence_after();
    if (pair) ptx::tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int auto_block_n(int n) {
  if (n <= 256) return (n + 15) / 16 * 16;
  int best = 256, best_pad = 1 << 30;
  for (int bn : {256, 192, 128}) {
    int pad = (n + bn - 1) / bn * bn - n;
    if (pad < best_pad) { best = bn; best_pad = pad; }
  }
  return best;
}

}  // namespace
}  // namespace svl

namespace svl {
int try_launch_conv_roll(const svl_gemm_desc* d, cudaStream_t stream);      // conv_roll.cu: 3 x 3 convolutions with 32 / 64 output channels
int conv_roll_gn_splits(const svl_gemm_desc* d);
}

using namespace svl;

extern "C" int svl_conv_gn_splits(const svl_gemm_desc* d) { return d ? conv_roll_gn_splits(d) : 0; }

extern "C" int svl_gemm(const svl_gemm_desc* d, void* stream) {
  SVL_CHECK_ARG(d && d->a && d->b && d->out, "svl_gemm: null operand");
  SVL_CHECK_ARG(d->m > 0 && d->n > 0, "svl_gemm: empty problem m=%lld n=%d", (long long)d->m, d->n);
  SVL_CHECK_ARG(d->k_per_tap > 0 && d->k_per_tap % 16 == 0, "svl_gemm: k_per_tap=%d must be a positive multiple of 16", d->k_per_tap);
  SVL_CHECK_ARG(d->num_taps >= 1 && d->num_taps <= SVL_MAX_TAPS, "svl_gemm: num_taps=%d out of range", d->num_taps);
  SVL_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "svl_gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  SVL_CHECK_ARG(((uintptr_t)d->a & 15) == 0 && ((uintptr_t)d->b & 15) == 0, "svl_gemm: operands must be 16-byte aligned");
  SVL_CHECK_ARG(d->out_dtype == SVL_F32 || d->out_dtype == SVL_BF16 || d->out_dtype == SVL_BF16X2, "svl_gemm: bad out_dtype");
  SVL_CHECK_ARG(!d->accumulate || d->out_dtype == SVL_F32, "svl_gemm: accumulate needs an F32 output");
  if (d->out_mode == SVL_OUT_CONVT2X2)
    SVL_CHECK_ARG(d->n % 32 == 0 && d->out_h > 0 && d->out_w > 0 && d->m % ((int64_t)d->out_h * d->out_w) == 0,
                  "svl_gemm: CONVT2X2 needs n %% 32 == 0 and m a multiple of out_h*out_w");
  SVL_CHECK_ARG(d->act != SVL_ACT_GELU_DSAVE || d->preact_out, "svl_gemm: SVL_ACT_GELU_DSAVE needs preact_out (it receives gelu')");
  SVL_CHECK_ARG(d->act != SVL_ACT_SAVED && d->dact_kind != SVL_ACT_GELU_DSAVE, "svl_gemm: SVL_ACT_SAVED is a dact_kind, SVL_ACT_GELU_DSAVE an act");
  if (int rc = svl_check_device()) return rc;
  if (d->a_conv) {
    const int rc = try_launch_conv_roll(d, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : SVL_OK;
  }
  SVL_CHECK_ARG(!d->gn_part, "svl_gemm: gn_part is only served by the problems svl_conv_gn_splits() accepts");

  GemmParams p;
  p.cluster = 0;
  memset(&p, 0, sizeof(p));
  p.a_conv = d->a_conv;
  p.m = d->m;
  p.n = d->n;
  p.k_per_tap = d->k_per_tap;
  p.num_taps = d->num_taps;
  for (int t = 0; t < d->num_taps; ++t) {
    p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_a_koff[t] = d->tap_a_koff[t];
    p.tap_b_row[t] = d->tap_b_row[t]; p.tap_b_col[t] = d->tap_b_col[t];
  }
  p.block_n = d->block_n > 0 ? d->block_n : auto_block_n(d->n);
  SVL_CHECK_ARG(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "svl_gemm: block_n=%d invalid", p.block_n);
  p.num_n_tiles = (d->n + p.block_n - 1) / p.block_n;
  p.n_fastest = p.num_n_tiles > 1 && !d->a_conv && (int64_t)d->m * d->k_per_tap * d->num_taps * 2 > (48ll << 20);
  const int64_t a_cols = d->a_cols > 0 ? d->a_cols : d->lda;

  CUtensorMap tmA, tmB, tmBh, tmC, tmD;
  if (d->a_conv) {
    SVL_CHECK_ARG(d->nb > 0 && d->h > 0 && d->w > 0 && (int64_t)d->nb * d->h * d->w == d->m, "svl_gemm: conv geometry does not match m");
    p.nb = d->nb; p.h = d->h; p.w = d->w;
    p.bw = d->w < BM ? d->w : BM;
    p.bh = d->h < BM / p.bw ? d->h : BM / p.bw;
    p.bn = d->nb < BM / (p.bw * p.bh) ? d->nb : BM / (p.bw * p.bh);
    p.tiles_x = (d->w + p.bw - 1) / p.bw;
    p.tiles_y = (d->h + p.bh - 1) / p.bh;
    int tiles_n = (d->nb + p.bn - 1) / p.bn;
    int64_t mt = (int64_t)p.tiles_x * p.tiles_y * tiles_n;
    SVL_CHECK_ARG(mt < (1ll << 30), "svl_gemm: too many tiles");
    p.num_m_tiles = (int)mt;
    p.a_tx_bytes = (uint32_t)(p.bw * p.bh * p.bn) * 128u;
    const uint64_t mw = d->a_map_w > 0 ? d->a_map_w : d->w;
    uint64_t dims[4] = {(uint64_t)a_cols, mw, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->lda * 2 * mw, (uint64_t)d->lda * 2 * mw * d->h};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    // ---- row-strip mode?
    static int strip_enabled = -1;
    if (strip_enabled < 0) { const char* e = getenv("SVL_CONV_STRIP"); strip_enabled = e ? atoi(e) : 1; }
    bool strip = strip_enabled && d->a_map_w == 0 && p.bw == BM && p.bh == 1 && p.bn == 1 && d->k_per_tap <= BK && p.num_n_tiles == 1 &&
                 d->num_taps >= 2 && (size_t)d->num_taps * p.block_n * 128 <= 100 * 1024;
    int span = 0;
    if (strip) {
      for (int t = 0; t < d->num_taps && strip; ++t) {
        if (d->tap_a_koff[t] != d->tap_a_koff[0]) { strip = false; break; }
        int g = -1;
        for (int k = 0; k < p.ng; ++k) if (p.g_dy[k] == d->tap_dy[t]) g = k;
        if (g < 0) {
          if (p.ng == 4) { strip = false; break; }
          g = p.ng++;
          p.g_dy[g] = d->tap_dy[t];
          p.g_dxmin[g] = d->tap_dx[t];
        }
        if (p.g_ntaps[g] == 4) { strip = false; break; }
        p.g_tap[g][p.g_ntaps[g]++] = t;
        if (d->tap_dx[t] < p.g_dxmin[g]) p.g_dxmin[g] = d->tap_dx[t];
      }
      for (int g = 0; g < p.ng && strip; ++g)
        for (int j = 0; j < p.g_ntaps[g]; ++j) {
          int sh = d->tap_dx[p.g_tap[g][j]] - p.g_dxmin[g];
          span = span > sh ? span : sh;
        }
      if (BM + span > 256) strip = false;
    }
    if (strip) {
      p.strip = 1;
      p.strip_bytes = (uint32_t)(BM + span) * 128u;
      box[1] = (uint32_t)(BM + span);
    } else {
      p.ng = 0;
    }
    if (int rc = tma_encode_bf16(&tmA, d->a, 4, dims, strides, box)) return rc;
  } else {
    int64_t mt = (d->m + BM - 1) / BM;
    SVL_CHECK_ARG(mt < (1ll << 30), "svl_gemm: too many tiles");
    p.num_m_tiles = (int)mt;
    p.a_tx_bytes = BM * 128u;
    uint64_t dims[2] = {(uint64_t)a_cols, (uint64_t)d->m};
    uint64_t strides[1] = {(uint64_t)d->lda * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (int rc = tma_encode_bf16(&tmA, d->a, 2, dims, strides, box)) return rc;
  }
  // ---- staged epilogue (EXT 5 / 6)?  Plain GEMMs whose epilogue moves a second tile: fp32 output + fp32 residual (out-proj, FFN2), bf16 output
  // times a saved bf16 act' tile (FFN2 dgrad), bf16 output + saved gelu' (FFN1).  SVL_GEMM_STAGED: bit 0 f32 + residual, bit 1 act' side tile,
  // bit 2 GELU_DSAVE, bit 3 every other plain bf16 output.
  static int staged_mask = -1;
  if (staged_mask < 0) { const char* e = getenv("SVL_GEMM_STAGED"); staged_mask = e ? atoi(e) : 7; }
  int staged_kind = 0;                                   // 5 / 6 = EXT of the staged variants
  {
    const bool base_ok = !d->a_conv && d->out_mode == SVL_OUT_LINEAR && !d->row_bias && !d->accumulate && d->n % 32 == 0 &&
                         ((uintptr_t)d->out & 15) == 0 && (!d->bias || ((uintptr_t)d->bias & 15) == 0) && p.block_n % 32 == 0;
    if (base_ok && d->out_dtype == SVL_F32 && d->residual && d->res_dtype == SVL_F32 && !d->preact_out && !d->dact_src && d->act == SVL_ACT_NONE &&
        d->ldc % 4 == 0 && d->ldres % 4 == 0 && ((uintptr_t)d->residual & 15) == 0 && (staged_mask & 1))
      staged_kind = 5;
    else if (base_ok && d->out_dtype == SVL_BF16 && !d->residual && d->ldc % 8 == 0) {
      const bool side = d->dact_src && (d->dact_kind == SVL_ACT_GELU || d->dact_kind == SVL_ACT_SAVED) && d->dact_dtype == SVL_BF16 &&
                        d->ld_dact % 8 == 0 && ((uintptr_t)d->dact_src & 15) == 0 && !d->preact_out && d->act == SVL_ACT_NONE;
      const bool dsave = d->act == SVL_ACT_GELU_DSAVE && !d->dact_src && d->preact_dtype == SVL_BF16 && d->ld_preact % 8 == 0 &&
                         ((uintptr_t)d->preact_out & 15) == 0;
      const bool plain_out = !d->dact_src && !d->preact_out && (d->act == SVL_ACT_NONE || d->act == SVL_ACT_GELU);
      if ((side && (staged_mask & 2)) || (dsave && (staged_mask & 4)) || (plain_out && (staged_mask & 8))) staged_kind = 6;
    }
  }
  p.staged = staged_kind == 5 ? 16384 : staged_kind == 6 ? 8192 : 0;
  {
    uint64_t dims[2] = {(uint64_t)d->ldb, (uint64_t)d->b_rows};
    uint64_t strides[1] = {(uint64_t)d->ldb * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)p.block_n};
    if (int rc = tma_encode_bf16(&tmB, d->b, 2, dims, strides, box)) return rc;
    tmBh = tmB;
    static int cluster_on = -1;
    // SVL_GEMM_CLUSTER: 0 off, 1 multicast, 2 pair everywhere; default (3) = pair mode where it was measured to pay: long-K plain GEMMs
    // (K >= 2048: 71 % -> 83 % tensor-active, profiles/r01_ncu_targets_c.md); the K = 768 shapes are epilogue-paced and gain nothing
    if (cluster_on < 0) { const char* e = getenv("SVL_GEMM_CLUSTER"); cluster_on = e ? atoi(e) : 3; }
    const bool eligible = !d->a_conv && p.num_m_tiles >= 2 && p.block_n % 32 == 0;
    p.cluster = !eligible ? 0 : (cluster_on == 1 || cluster_on == 2) ? cluster_on
              : (cluster_on == 3 && (((int64_t)d->k_per_tap * d->num_taps >= 2048 && d->num_taps == 1) || staged_kind == 5)) ? 2 : 0;
    // (fp32 staging buffers take 64 KB: only the pair mode's 32 KB stages leave room for a deep enough operand ring)
    if (p.cluster) {                                        // half-height box: the B half a CTA loads (multicast to both, or kept, in pair mode)
      uint32_t boxh[2] = {(uint32_t)BK, (uint32_t)(p.block_n / 2)};
      if (int rc = tma_encode_bf16(&tmBh, d->b, 2, dims, strides, boxh)) return rc;
    }
  }
  p.a_stage_bytes = BM * 128u;
  p.b_stage_bytes = (uint32_t)p.block_n * 128u;      // block_n % 16 == 0 -> multiple of 2048, keeps every stage 1024-aligned
  if (p.cluster == 2) p.b_stage_bytes >>= 1;         // pair mode: a CTA stores only its half of the B tile (block_n % 32 == 0)
  p.b_tile_bytes = p.b_stage_bytes;
  size_t smem_data;
  if (p.strip) {
    p.a_stage_bytes = (p.strip_bytes + 1023u) & ~1023u;
    const size_t wbytes = (size_t)p.num_taps * p.b_tile_bytes;
    p.stages = (int)((kSmemBudget - wbytes) / p.a_stage_bytes);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    smem_data = (size_t)p.stages * p.a_stage_bytes + wbytes;
    for (int g = 0; g < p.ng; ++g) {
      int nm = 0;
      for (int j = 0; j < p.g_ntaps[g]; ++j) {
        const int t = p.g_tap[g][j];
        for (int kk = 0; kk < p.k_per_tap / 16; ++kk, ++nm) {
          p.mma_aoff[g][nm] = ((uint32_t)(p.tap_dx[t] - p.g_dxmin[g]) * 128u + (uint32_t)kk * 32u) >> 4;
          p.mma_boff[g][nm] = ((uint32_t)t * p.b_tile_bytes + (uint32_t)kk * 32u) >> 4;
        }
      }
      p.g_nmma[g] = nm;
    }
  } else {
    const size_t stage_c = (size_t)p.staged * kEpiGroups;
    const size_t budget = p.staged ? (size_t)(227 * 1024 - 1024 - 512) : (size_t)kSmemBudget;       // operand ring + staging buffers fill the SM
    p.stages = (int)((budget - stage_c) / (p.a_stage_bytes + p.b_stage_bytes));
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    smem_data = (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + stage_c;
    if (p.staged) {
      const int dt = staged_kind == 5 ? SVL_F32 : SVL_BF16, es = staged_kind == 5 ? 4 : 2, sw = staged_kind == 5 ? 128 : 64;
      uint32_t boxc[2] = {32u, (uint32_t)BM};
      uint64_t dimsc[2] = {(uint64_t)d->n, (uint64_t)d->m};
      uint64_t stridesc[1] = {(uint64_t)d->ldc * es};
      if (int rc = tma_encode(&tmC, d->out, dt, sw, 2, dimsc, stridesc, boxc)) return rc;
      const void* second = staged_kind == 5 ? d->residual : d->dact_src ? d->dact_src : d->preact_out;
      const int64_t ld2 = staged_kind == 5 ? d->ldres : d->dact_src ? d->ld_dact : d->ld_preact;
      if (second) {
        uint64_t stridesd[1] = {(uint64_t)ld2 * es};
        if (int rc = tma_encode(&tmD, second, dt, sw, 2, dimsc, stridesd, boxc)) return rc;
      }
    }
  }
  if (!p.staged) tmC = tmA;
  if (!p.staged || (staged_kind == 6 && !d->dact_src && !d->preact_out)) tmD = tmA;
  p.acc_stages = 512 / p.block_n < kMaxAcc ? 512 / p.block_n : kMaxAcc;
  if (const char* e = getenv("SVL_ACC_STAGES")) { int v = atoi(e); if (v >= 1 && v <= p.acc_stages) p.acc_stages = v; }
  p.tmem_cols = pow2ceil(p.acc_stages * p.block_n < 32 ? 32 : p.acc_stages * p.block_n);

  p.out = d->out; p.out_dtype = d->out_dtype; p.ldc = d->ldc; p.out_mode = d->out_mode; p.out_h = d->out_h; p.out_w = d->out_w;
  p.alpha = d->alpha; p.bias = d->bias; p.row_bias = d->row_bias; p.row_bias_div = d->row_bias_div > 0 ? d->row_bias_div : 1;
  p.row_bias_ld = d->row_bias_ld;
  p.act = d->act; p.preact_out = d->preact_out; p.preact_dtype = d->preact_dtype; p.ld_preact = d->ld_preact;
  p.dact_src = d->dact_src; p.dact_dtype = d->dact_dtype; p.dact_kind = d->dact_kind; p.ld_dact = d->ld_dact;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.ldres = d->ldres;
  p.accumulate = d->accumulate;
#ifdef SVL_GEMM_DIAG
  { const char* e = getenv("SVL_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("SVL_GEMM_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 10) : nullptr; }
#endif

  const size_t smem = 1024 + smem_data + 8 * (2 * kMaxStages + 2 * kMaxAcc + 2 + kEpiGroups) + 16;
  if (p.strip) p.cluster = 0;
  const int num_tiles = (p.cluster ? (p.num_m_tiles + 1) / 2 * 2 : p.num_m_tiles) * p.num_n_tiles;
  int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  if (p.cluster) grid &= ~1;                                // whole clusters
  const bool plain = !p.row_bias && !p.accumulate && p.out_mode == SVL_OUT_LINEAR;
  const bool no_extra = !p.preact_out && !p.dact_src && !p.residual && p.act == SVL_ACT_NONE;
#define SVL_LAUNCH_GEMM_AS(PAIRV, ...)                                                                                    \
  do {                                                                                                                    \
    static bool attr_set = false;                                                                                         \
    if (!attr_set) {                                                                                                      \
      SVL_CUDA(cudaFuncSetAttribute(gemm_kernel<__VA_ARGS__, PAIRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                                    \
    }                                                                                                                     \
    if (p.cluster) {                                                                                                      \
      cudaLaunchConfig_t cfg = {};                                                                                        \
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream; \
      cudaLaunchAttribute at[1];                                                                                          \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                                     \
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                                 \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                                   \
      SVL_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<__VA_ARGS__, PAIRV>, tmA, tmB, tmBh, tmC, tmD, p));                              \
    } else {                                                                                                              \
      gemm_kernel<__VA_ARGS__, PAIRV><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA, tmB, tmBh, tmC, tmD, p);          \
    }                                                                                                                     \
  } while (0)
#define SVL_LAUNCH_GEMM(...)                                          \
  do {                                                                \
    if (p.cluster == 2) SVL_LAUNCH_GEMM_AS(true, __VA_ARGS__);        \
    else SVL_LAUNCH_GEMM_AS(false, __VA_ARGS__);                      \
  } while (0)
  if (staged_kind == 5) {
    SVL_LAUNCH_GEMM(SVL_F32, SVL_ACT_NONE, -1, 0, SVL_F32, 5, false);
  } else if (staged_kind == 6) {
    if (p.dact_src && p.dact_kind == SVL_ACT_GELU) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 1, -1, 6, false);
    else if (p.dact_src) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 2, -1, 6, false);
    else if (p.act == SVL_ACT_GELU_DSAVE) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_GELU_DSAVE, SVL_BF16, 0, -1, 6, false);
    else if (p.act == SVL_ACT_GELU) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_GELU, -1, 0, -1, 6, false);
    else SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 0, -1, 6, false);
  } else if (plain && no_extra && p.out_dtype == SVL_BF16) {
    SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 0, -1, 0, false);
  } else if (plain && no_extra && p.out_dtype == SVL_F32) {
    SVL_LAUNCH_GEMM(SVL_F32, SVL_ACT_NONE, -1, 0, -1, 0, false);
  } else if (plain && p.out_dtype == SVL_F32 && p.residual && p.res_dtype == SVL_F32 && !p.preact_out && !p.dact_src && p.act == SVL_ACT_NONE) {
    SVL_LAUNCH_GEMM(SVL_F32, SVL_ACT_NONE, -1, 0, SVL_F32, 0, false);      // register epilogue (SVL_GEMM_STAGED bit 0 off, or an unaligned problem)
  } else if (plain && p.out_dtype == SVL_BF16 && p.act == SVL_ACT_GELU && !p.dact_src && !p.residual &&
             (!p.preact_out || p.preact_dtype == SVL_BF16)) {
    if (p.preact_out) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_GELU, SVL_BF16, 0, -1, 0, false);
    else SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_GELU, -1, 0, -1, 0, false);
  } else if (plain && p.out_dtype == SVL_BF16 && p.act == SVL_ACT_GELU_DSAVE && !p.dact_src && !p.residual && p.preact_dtype == SVL_BF16) {
    SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_GELU_DSAVE, SVL_BF16, 0, -1, 0, false);
  } else if (plain && p.out_dtype == SVL_BF16 && p.dact_src && (p.dact_kind == SVL_ACT_GELU || p.dact_kind == SVL_ACT_SAVED) &&
             p.dact_dtype == SVL_BF16 && !p.preact_out && !p.residual && p.act == SVL_ACT_NONE && p.n % 32 == 0 && p.ld_dact % 8 == 0 &&
             ((uintptr_t)p.dact_src & 15) == 0) {
    if (p.dact_kind == SVL_ACT_GELU) SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 1, -1, 0, false);
    else SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 2, -1, 0, false);
  } else if (no_extra && p.out_mode == SVL_OUT_CONVT2X2 && p.out_dtype == SVL_BF16 && !p.row_bias && !p.accumulate && (p.n / 4) % 16 == 0) {
    SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 0, -1, 1, false);
  } else if (no_extra && p.out_mode == SVL_OUT_LINEAR && p.out_dtype == SVL_BF16 && p.row_bias && !p.accumulate) {
    SVL_LAUNCH_GEMM(SVL_BF16, SVL_ACT_NONE, -1, 0, -1, 2, false);
  } else if (no_extra && p.out_mode == SVL_OUT_LINEAR && p.out_dtype == SVL_F32 && !p.row_bias && p.accumulate) {
    SVL_LAUNCH_GEMM(SVL_F32, SVL_ACT_NONE, -1, 0, -1, 3, false);
  } else {
    SVL_LAUNCH_GEMM(SVL_F32, SVL_ACT_NONE, -1, 0, -1, 0, true);
  }
#undef SVL_LAUNCH_GEMM
#undef SVL_LAUNCH_GEMM_AS
  SVL_LAUNCH_CHECK();
  return SVL_OK;
}
