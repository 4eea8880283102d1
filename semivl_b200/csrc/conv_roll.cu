// 3 x 3 (dilation 1) convolutions with 32 or 64 output channels on 128-pixel row tiles: the `Up` blocks of the decode head (vlg_head.py:116-137),
// forward and data gradient.  In the generic engine (gemm.cu) a 128-pixel tile of such a layer is 18-36 MMAs of N = 32 / 64: each reads a
// 4 KB A slab for 16-32 tensor-pipe cycles, and the single issuing thread spends ~45 cycles per MMA on descriptors, so those launches ran
// at 280-510 TFLOP/s, 3x over their HBM time (profiles/r02_gemm_epilogue.md).  Here the roles of rows and taps are swapped:
//
//   * ONE input strip (image row r, 130 pixels, TMA zero fill outside the image) is loaded ONCE and feeds the three OUTPUT rows r-1, r, r+1:
//     the MMA of dx tap j multiplies the strip (read j pixels further: descriptor start address) with B_j = [W(dy=+1,dx_j); W(dy=0,dx_j);
//     W(dy=-1,dx_j)], N = 3 * cout, and its three cout-wide column blocks are the accumulators of three consecutive output rows:
//     3 MMAs of N = 96 / 192 per 16 input channels instead of 9 of N = 32 / 64, a third of the A-operand shared-memory reads, one strip load
//     per tile instead of three;
//   * the accumulators form a RING of 512 / cout blocks in tensor memory.  A CTA walks down `roll_rows` image rows of one 128-pixel column
//     (a unit); strip s of its stream accumulates into blocks s, s+1, s+2 (always accumulate: the epilogue warps hand every block back zeroed)
//     and completes block s.  The first two and last two blocks of a unit belong to rows outside it: they are never read, only re-zeroed;
//   Measured (config-2 size, 336 maps of 128 x 128; scratch/conv_up.py, same box): 32 -> 32 channels 287 -> 147 us (689 TFLOP/s, 4.8 TB/s of
//   input + output bytes), 64 -> 32 399 -> 227 us, 32 -> 64 308 -> 243 us.  The first version of this kernel was no faster than the generic
//   engine: the per-role cycle counters (-DSVL_GEMM_DIAG, SVL_ROLL_TRACE) showed the ISSUING WARP as the critical path -- ~1450 cycles of
//   serial uniform-datapath code per strip (runtime loop bounds, per-MMA descriptor arithmetic, selects) against ~300 cycles of tensor time,
//   with the strip loads, the epilogue (idle 3/4 of the time), the TMEM read-modify-write of always-accumulating MMAs and the pixel-shifted
//   descriptors all ruled out by knock-out runs.  Templating on (cin / 16, cout), hoisting the B descriptors out of the strip loop and making
//   the non-wrapping case straight-line code halved the launch time.
//   * maps at most 64 pixels wide (the first Up block: 64 x 64 at the 512^2 crop) use the DUAL form: a tile is one row of TWO maps
//     (lanes 0..63 | 64..127), which cannot be one shifted window of a haloed strip (the second map would start 66, not 64, lines in), so
//     each dx tap gets its own box {64 channels, 64 pixels, 1 row, 2 maps} loaded at pixel offset dx - 1 and a pipeline stage of its own;
//   * more than 64 input channels = two 64-channel K chunks per stage;
//   * warp 0 = TMA producer (weights once, strips through a shared-memory ring), warp 1 = MMA issuer, warps 2..17 = four epilogue
//     quartets (tcgen05.ld -> bias / ReLU -> bf16 NHWC store -> tcgen05.st zeros -> block free); the blocks go to the quartets in turn: one
//     quartet's chain of TMEM load, store and re-zero latencies is ~4x the tensor time of a strip (measured: 4 warps alone left the
//     kernel at the generic engine's speed).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma.h"

namespace svl {
namespace {

constexpr int kRollQuartets = 4;                      // epilogue quartets (4 warps = 128 pixels each); accumulator block n belongs to quartet n % 4
constexpr int kRollThreads = 64 + 128 * kRollQuartets;
constexpr int kRollMaxStages = 8;
constexpr int kRollMaxBlocks = 16;

struct RollParams {
  int nb, h, w, tiles_x, roll_rows, chunks, units;   // tiles_x: 128-pixel columns of a map (single form) / 1 (dual form: a unit column is a pair of maps)
  int cin, cout, nblk;               // nblk = 512 / cout accumulator blocks in the ring
  int a_koff, stages;
  uint32_t stage_tx, chunk_stride, stage_stride, wgroup_bytes;   // wgroup = the three filter rows of one (dx tap, K chunk): 3 * cout rows of 128 bytes
  int wrow[3][3], wcol[3][3];        // weight-tensor coordinates (row, column) of filter position (dy + 1, dx + 1)
  __nv_bfloat16* out; int64_t ldc;
  const float* bias; int relu;
  float* gn_part; int gn_splits;     // fused GroupNorm statistics of the stored output (16 channels per group), see svl_gemm_desc.gn_part
  long long* trace;                  // -DSVL_GEMM_DIAG: cycle totals of CTA 0
};

__device__ __forceinline__ void tmem_zero_32(uint32_t taddr) {
  uint32_t z[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) z[i] = 0u;
  ptx::tmem_st_32x32(taddr, z);
}

template <int KS, int COUT, bool DUAL>
__global__ void __launch_bounds__(kRollThreads, 1)
conv_roll_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ RollParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t smem_w = (raw + 1023u) & ~1023u;
  constexpr int KC = (KS + 3) / 4;                 // 64-channel K chunks
  const uint32_t smem_a = smem_w + 3u * KC * p.wgroup_bytes;
  const uint32_t bar_base = smem_a + (uint32_t)p.stages * p.stage_stride;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kRollMaxStages + s); };
  auto bfull_bar = [&](int b) { return bar_base + 8u * (2 * kRollMaxStages + b); };
  auto bempty_bar = [&](int b) { return bar_base + 8u * (2 * kRollMaxStages + kRollMaxBlocks + b); };
  const uint32_t wbar = bar_base + 8u * (2 * kRollMaxStages + 2 * kRollMaxBlocks);
  const uint32_t tmem_ptr_addr = wbar + 8u;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < p.nblk; ++b) {
      ptx::mbar_init(bfull_bar(b), 1);
      ptx::mbar_init(bempty_bar(b), 4);          // the four epilogue warps hand a block back
    }
    ptx::mbar_init(wbar, 1);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wbar, 3u * KC * p.wgroup_bytes);
      for (int dxi = 0; dxi < 3; ++dxi)
        for (int kc = 0; kc < KC; ++kc)
          for (int k = 0; k < 3; ++k)            // block order inside a group: output rows r-1, r, r+1 <- filter rows +1, 0, -1
            ptx::tma_load_2d(smem_w + (uint32_t)(dxi * KC + kc) * p.wgroup_bytes + (uint32_t)(k * COUT) * 128u, &tmW, wbar,
                             p.wcol[2 - k][dxi] + kc * 64, p.wrow[2 - k][dxi]);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int c = u % p.chunks, tx = (u / p.chunks) % p.tiles_x, cn = u / (p.chunks * p.tiles_x);
        const int y0 = c * p.roll_rows, y1 = min(y0 + p.roll_rows, p.h);
        for (int rr = y0 - 1; rr <= y1; ++rr) {
          for (int dxi = 0; dxi < (DUAL ? 3 : 1); ++dxi) {
#ifdef SVL_GEMM_DIAG
            const long long tp = clock64();
#endif
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
#ifdef SVL_GEMM_DIAG
            if (p.trace && blockIdx.x == 0) p.trace[4] += clock64() - tp;
#endif
            ptx::mbar_arrive_expect_tx(full_bar(stage), p.stage_tx);
            for (int kc = 0; kc < KC; ++kc)
              ptx::tma_load_4d(smem_a + (uint32_t)stage * p.stage_stride + (uint32_t)kc * p.chunk_stride, &tmA, full_bar(stage), p.a_koff + kc * 64,
                               DUAL ? dxi - 1 : tx * 128 - 1, rr, DUAL ? 2 * cn : cn);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = ptx::elect_one();
    const uint64_t tmpl = ptx::make_smem_desc(0, 16, 1024);
    constexpr int NBLK = 512 / COUT;
    const uint32_t idesc1 = ptx::make_idesc_bf16(128, COUT, 0, 0), idesc2 = ptx::make_idesc_bf16(128, 2 * COUT, 0, 0),
                   idesc3 = ptx::make_idesc_bf16(128, 3 * COUT, 0, 0);
    ptx::mbar_wait(wbar, 0);
    ptx::tc_fence_after();
    // the issuing warp is the critical path (~1450 cycles of serial uniform code per strip in the first version against ~300 cycles of tensor
    // time): everything that does not depend on the strip is hoisted -- the B descriptors of the 3 * KS MMAs are loop constants, the A
    // descriptors are one 32-bit add away from the stage base, and the common case (the three blocks do not wrap) is straight-line code
    uint64_t bdesc[3][KS];
#pragma unroll
    for (int dxi = 0; dxi < 3; ++dxi)
#pragma unroll
      for (int kk = 0; kk < KS; ++kk)
        bdesc[dxi][kk] = tmpl + (uint64_t)((smem_w >> 4) + (uint32_t)((dxi * KC + (kk >> 2)) * 3 * COUT * 8 + (kk & 3) * 2));
    const uint32_t a_base = smem_a >> 4, a_step = p.stage_stride >> 4, a_chunk = p.chunk_stride >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int blk = 0;                               // ring block of the current strip's first output row
    uint32_t bphase = 0;                       // parity of the ring turn `blk` is in
    auto acquire = [&](int j) {                // block blk + j must have been handed back (zeroed) by the epilogue warps
      int b = blk + j;
      uint32_t ph = bphase;
      if (b >= NBLK) { b -= NBLK; ph ^= 1u; }
      ptx::mbar_wait(bempty_bar(b), ph);
    };
    auto complete = [&](int j) {
      int b = blk + j;
      if (b >= NBLK) b -= NBLK;
      ptx::umma_commit(bfull_bar(b));
    };
#ifdef SVL_GEMM_DIAG
    const long long tm0 = clock64();
#endif
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int c = u % p.chunks;
      const int y0 = c * p.roll_rows, nstrips = min(y0 + p.roll_rows, p.h) - y0 + 2;
      for (int s = 0; s < nstrips; ++s) {
#ifdef SVL_GEMM_DIAG
        const long long t0 = clock64();
#endif
        if (s == 0) { acquire(0); acquire(1); }
        acquire(2);
#ifdef SVL_GEMM_DIAG
        const long long t1 = clock64();
#endif
        const uint32_t dcol = tmem_base + (uint32_t)(blk * COUT);
        const bool last = s == nstrips - 1;
        const int first = NBLK - blk;              // blocks before the ring wraps (>= 3: no wrap inside this strip)
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi) {
          if (DUAL || dxi == 0) {                  // single form: one stage holds the strip of all three taps; dual form: a stage per tap
            ptx::mbar_wait(full_bar(stage), phase);
            ptx::tc_fence_after();
          }
          const uint32_t a_lo = a_base + (uint32_t)stage * a_step + (DUAL ? 0u : (uint32_t)(dxi * 8));
          if (first >= 3) {
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < KS; ++kk)
                ptx::umma_bf16(dcol, tmpl + (uint64_t)(a_lo + (uint32_t)((kk >> 2) * a_chunk + (kk & 3) * 2)), bdesc[dxi][kk], idesc3, 1u);
            }
          } else if (leader) {                     // the ring wraps inside the strip's three blocks: two MMAs per step
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              const uint64_t adesc = tmpl + (uint64_t)(a_lo + (uint32_t)((kk >> 2) * a_chunk + (kk & 3) * 2));
              ptx::umma_bf16(dcol, adesc, bdesc[dxi][kk], first == 2 ? idesc2 : idesc1, 1u);
              ptx::umma_bf16(tmem_base, adesc, bdesc[dxi][kk] + (uint64_t)(first * COUT * 8), first == 1 ? idesc2 : idesc1, 1u);
            }
          }
          if (DUAL || dxi == 2) {
            if (leader) ptx::umma_commit(empty_bar(stage));
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
#ifdef SVL_GEMM_DIAG
        if (p.trace && blockIdx.x == 0 && lane == 0) { p.trace[1] += t1 - t0; p.trace[2] += clock64() - t1; p.trace[3] += 1; }
#endif
        if (leader) {
          complete(0);
          if (last) { complete(1); complete(2); }
        }
        __syncwarp();
        blk += last ? 3 : 1;
        if (blk >= NBLK) { blk -= NBLK; bphase ^= 1u; }
      }
    }
#ifdef SVL_GEMM_DIAG
    if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[0] = clock64() - tm0;
#endif
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                    // TMEM lane quarter: pixels 32 q .. 32 q + 31 of the tile
    const int quartet = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    if (quartet == 0) {                        // every block starts zeroed and free
      for (int c0 = 0; c0 < 512; c0 += 32) tmem_zero_32(lane_base + (uint32_t)c0);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int b = 0; b < p.nblk; ++b) ptx::mbar_arrive(bempty_bar(b));
    }
    int blk = 0, turn = 0;                     // ring block and whose turn it is
    uint32_t bphase = 0;
    constexpr int G = COUT / 16;               // GroupNorm groups of 16 channels
    float gs1[G], gs2[G];                      // this warp's sums over the unit's rows it drains (fixed order: rows ascending)
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int c = u % p.chunks, tx = (u / p.chunks) % p.tiles_x, cn = u / (p.chunks * p.tiles_x);
      const int y0 = c * p.roll_rows, nrows = min(y0 + p.roll_rows, p.h) - y0;
      // single form: lane = pixel of the 128-pixel column tx; dual form: lanes 0..63 | 64..127 = pixels of the maps 2 cn | 2 cn + 1
      const int x = DUAL ? (q & 1) * 32 + lane : tx * 128 + q * 32 + lane;
      const int img = DUAL ? 2 * cn + (q >> 1) : cn;
      const bool valid = x < p.w && img < p.nb;
#pragma unroll
      for (int g = 0; g < G; ++g) gs1[g] = gs2[g] = 0.f;
      for (int j = 0; j < nrows + 4; ++j) {
        if (turn != quartet) {                 // another quartet's block
          if (++turn == kRollQuartets) turn = 0;
          if (++blk == p.nblk) { blk = 0; bphase ^= 1u; }
          continue;
        }
#ifdef SVL_GEMM_DIAG
        const long long te = clock64();
#endif
        ptx::mbar_wait(bfull_bar(blk), bphase);
        ptx::tc_fence_after();
#ifdef SVL_GEMM_DIAG
        if (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0) { p.trace[5] += clock64() - te; p.trace[6] += 1; }
#endif
        const uint32_t taddr = lane_base + (uint32_t)(blk * COUT);
        if (j >= 2 && j < nrows + 2) {         // a row of this unit: y0 + j - 2
          __nv_bfloat16* dst = p.out + (((int64_t)img * p.h + (y0 + j - 2)) * p.w + x) * p.ldc;
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[g * 16 + i]);
                if (p.bias) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 bb = __ldg((const float4*)(p.bias + c0 + g * 16) + i);
                    f[4 * i] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
                  }
                }
                if (p.relu) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                st16(dst, SVL_BF16, c0 + g * 16, 0, 16, f);
                if (p.gn_part) {                 // statistics of what was stored: the values rounded to bf16
                  float a = 0.f, b = 0.f;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float r = __bfloat162float(__float2bfloat16(f[i]));
                    a += r;
                    b += r * r;
                  }
                  gs1[c0 / 16 + g] += a;
                  gs2[c0 / 16 + g] += b;
                }
              }
            }
          }
        }
        for (int c0 = 0; c0 < COUT; c0 += 32) tmem_zero_32(taddr + (uint32_t)c0);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bempty_bar(blk));
        if (++turn == kRollQuartets) turn = 0;
        if (++blk == p.nblk) { blk = 0; bphase ^= 1u; }
      }
      if (p.gn_part) {
        // one partial per (map, unit, warp): slot order depends on the map geometry only, every slot is written exactly once
        const int slot = DUAL ? c * 8 + quartet * 2 + (q & 1) : (c * p.tiles_x + tx) * 16 + quartet * 4 + q;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float a = warp_sum(gs1[g]), b = warp_sum(gs2[g]);
          if (lane == 0 && img < p.nb) *(float2*)(p.gn_part + (((int64_t)img * p.gn_splits + slot) * G + g) * 2) = make_float2(a, b);
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

}  // namespace

namespace {
// Fills the launch plan; false when the problem does not qualify for the rolling kernel.
bool roll_plan(const svl_gemm_desc* d, RollParams& p, bool& dual) {
  static int roll_rows = -1;
  if (roll_rows < 0) { const char* e = getenv("SVL_CONV_ROLL"); roll_rows = e ? atoi(e) : 16; }
  if (roll_rows < 1 || !d->a_conv || d->num_taps != 9 || d->a_map_w != 0 || d->w < 33) return false;
  if ((d->n != 32 && d->n != 64) || (d->k_per_tap != 32 && d->k_per_tap != 64 && d->k_per_tap != 128)) return false;
  if (d->out_dtype != SVL_BF16 || d->out_mode != SVL_OUT_LINEAR || d->preact_out || d->dact_src || d->residual || d->row_bias || d->accumulate) return false;
  if (d->act != SVL_ACT_NONE && d->act != SVL_ACT_RELU) return false;
  if (d->alpha != 0.f && d->alpha != 1.f) return false;
  if (d->ldc % 16 != 0 || ((uintptr_t)d->out & 31) != 0 || (d->bias && ((uintptr_t)d->bias & 15) != 0)) return false;
  memset(&p, 0, sizeof(p));
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) p.wrow[a][b] = -1;
  for (int t = 0; t < 9; ++t) {
    const int fy = d->tap_dy[t], fx = d->tap_dx[t];
    if (fy < -1 || fy > 1 || fx < -1 || fx > 1 || d->tap_a_koff[t] != d->tap_a_koff[0]) return false;
    p.wrow[fy + 1][fx + 1] = d->tap_b_row[t];
    p.wcol[fy + 1][fx + 1] = d->tap_b_col[t];
  }
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) if (p.wrow[a][b] < 0) return false;
  dual = d->w <= 64;                              // two maps side by side in the 128 accumulator lanes
  if (!dual && d->w < 96) return false;           // 65..95-pixel rows would leave a third of a 128-pixel tile empty: generic engine
  p.nb = d->nb; p.h = d->h; p.w = d->w;
  p.cin = d->k_per_tap; p.cout = d->n; p.nblk = 512 / p.cout;
  const int kc = (p.cin + 63) / 64;
  p.tiles_x = dual ? 1 : (d->w + 127) / 128;
  p.roll_rows = roll_rows < d->h ? roll_rows : d->h;
  p.chunks = (d->h + p.roll_rows - 1) / p.roll_rows;
  p.units = (dual ? (d->nb + 1) / 2 : d->nb) * p.tiles_x * p.chunks;
  p.a_koff = d->tap_a_koff[0];
  const uint32_t box_bytes = (dual ? 128u : 130u) * 128u;
  p.chunk_stride = (box_bytes + 1023u) & ~1023u;
  p.stage_stride = (uint32_t)kc * p.chunk_stride;
  p.stage_tx = (uint32_t)kc * box_bytes;
  p.wgroup_bytes = (uint32_t)(3 * p.cout) * 128u;
  const size_t wbytes = 3 * (size_t)kc * p.wgroup_bytes;
  const size_t fixed = 1024 + wbytes + 8 * (2 * kRollMaxStages + 2 * kRollMaxBlocks + 2) + 16;
  if (fixed + 2 * (size_t)p.stage_stride > 227 * 1024) return false;
  p.stages = (int)((227 * 1024 - fixed) / p.stage_stride);
  if (p.stages > kRollMaxStages) p.stages = kRollMaxStages;
  p.out = (__nv_bfloat16*)d->out; p.ldc = d->ldc; p.bias = d->bias; p.relu = d->act == SVL_ACT_RELU;
  p.gn_splits = dual ? p.chunks * 8 : p.chunks * p.tiles_x * 16;
  p.gn_part = d->gn_part;
  return true;
}
}  // namespace

int conv_roll_gn_splits(const svl_gemm_desc* d) {
  RollParams p;
  bool dual;
  return roll_plan(d, p, dual) ? p.gn_splits : 0;
}

// Returns 1 when the rolling kernel was launched, 0 when the problem does not qualify, < 0 on error.
int try_launch_conv_roll(const svl_gemm_desc* d, cudaStream_t stream) {
  RollParams p;
  bool dual;
  if (!roll_plan(d, p, dual)) return 0;
  const int kc = (p.cin + 63) / 64;
  const size_t fixed = 1024 + 3 * (size_t)kc * p.wgroup_bytes + 8 * (2 * kRollMaxStages + 2 * kRollMaxBlocks + 2) + 16;
#ifdef SVL_GEMM_DIAG
  { const char* e = getenv("SVL_ROLL_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 10) : nullptr; }
#endif
  CUtensorMap tmA, tmW;
  {
    const int64_t a_cols = d->a_cols > 0 ? d->a_cols : d->lda;
    uint64_t dims[4] = {(uint64_t)a_cols, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->nb};
    uint64_t strides[3] = {(uint64_t)d->lda * 2, (uint64_t)d->lda * 2 * d->w, (uint64_t)d->lda * 2 * d->w * d->h};
    uint32_t box[4] = {64u, dual ? 64u : 130u, 1u, dual ? 2u : 1u};
    if (int rc = tma_encode_bf16(&tmA, d->a, 4, dims, strides, box)) return rc;
    uint64_t dimsw[2] = {(uint64_t)d->ldb, (uint64_t)d->b_rows};
    uint64_t stridesw[1] = {(uint64_t)d->ldb * 2};
    uint32_t boxw[2] = {64u, (uint32_t)p.cout};
    if (int rc = tma_encode_bf16(&tmW, d->b, 2, dimsw, stridesw, boxw)) return rc;
  }
  const size_t smem = fixed + (size_t)p.stages * p.stage_stride;
  const int grid = p.units < num_sms() ? p.units : num_sms();
#define SVL_LAUNCH_ROLL_AS(KS, COUT, DUALV)                                                                                       \
  do {                                                                                                                            \
    static bool attr_set = false;                                                                                                 \
    if (!attr_set) {                                                                                                              \
      SVL_CUDA(cudaFuncSetAttribute(conv_roll_kernel<KS, COUT, DUALV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                                            \
    }                                                                                                                             \
    conv_roll_kernel<KS, COUT, DUALV><<<grid, kRollThreads, smem, stream>>>(tmA, tmW, p);                                         \
  } while (0)
#define SVL_LAUNCH_ROLL(KS, COUT)                        \
  do {                                                   \
    if (dual) SVL_LAUNCH_ROLL_AS(KS, COUT, true);        \
    else SVL_LAUNCH_ROLL_AS(KS, COUT, false);            \
  } while (0)
  const int ks = p.cin / 16;
  if (ks == 2 && p.cout == 32) SVL_LAUNCH_ROLL(2, 32);
  else if (ks == 4 && p.cout == 32) SVL_LAUNCH_ROLL(4, 32);
  else if (ks == 8 && p.cout == 32) SVL_LAUNCH_ROLL(8, 32);
  else if (ks == 2 && p.cout == 64) SVL_LAUNCH_ROLL(2, 64);
  else if (ks == 4 && p.cout == 64) SVL_LAUNCH_ROLL(4, 64);
  else SVL_LAUNCH_ROLL(8, 64);
#undef SVL_LAUNCH_ROLL
#undef SVL_LAUNCH_ROLL_AS
  SVL_LAUNCH_CHECK();
  return 1;
}

}  // namespace svl
