// conv_kernel.cuh — INT8 implicit-GEMM convolution for sm_100a ("shift-GEMM" form).
//
// Replaces the reference's im2col_nchw + sgemm_tiled (+ bn_inference + add_inplace + relu_forward)
// sequence (reference: cpp/fp32/kernels/im2col.cu:6-58, sgemm_tiled.cu:6-46, runtime/infer_e2e.cu:102-136,
// :156-203) with ONE kernel: no col buffer, int8 x int8 -> int32 on tcgen05 tensor cores, fused epilogue.
//
// Formulation
//   Activations live in HBM as row-padded NHWC int8:  [PR zero rows][img0: H rows][PR zero rows][img1]...
//   Output positions are linearised in a *virtual* space  g = vrow * Wp + x  with vrow = n*Pv + r,
//   where Wp >= Wo + (taps-1) is the virtual row pitch and Pv the per-image virtual row pitch.
//   Positions with x >= Wo or r >= Ho are garbage and never stored.
//   A CTA handles a super-tile of MT*128 consecutive positions.  It TMA-loads the input patch (all
//   rows the super-tile touches, full virtual width, zero fill outside the tensor) ONCE into shared
//   memory, K-major with the hardware swizzle; filter tap (a,b) is then simply the same patch viewed
//   through a UMMA descriptor whose start address is advanced by (a*Wp + b) pixel rows.  Every input
//   byte crosses L2->SM once per super-tile instead of once per tap.
//   Stride-2 convolutions use up to four parity planes, each loaded by TMA with elementStrides = 2.
//   The 3-channel stem uses a 2x2 space-to-depth input stored as 32-byte "pixel pairs" [p | p+1] (SW32 rows),
//   so one K=32 MMA covers two horizontally adjacent taps.
//   (All three addressing tricks are validated by probe/umma_probe.cu on a B200.)
//
// CTA pairs (template TWO): for output-channel tiles of 128 the kernel runs as clusters of two CTAs on the two SMs
//   of a TPC and issues tcgen05.mma.cta_group::2 (M = 256): each SM feeds its own 128 positions and HALF of the
//   weight rows from its own shared memory, which halves the shared-memory operand traffic per MAC - the
//   resource that bounds the single-CTA form (probe/mma_rate.cu: one SS MMA reads 4 KB + 32 N bytes per N/2
//   cycles of tensor time against 128 B/clk of shared-memory bandwidth).  Pairs use row-aligned super-tiles (a
//   whole number of virtual rows) so both CTAs address their patches with identical descriptors.
//
// Warp roles (384 threads): warp 0 = activation-patch TMA producer, warps 1 and 3 = MMA issuers (warp 1 owns
//   TMEM), warp 2 = weight-step TMA producer, warps 4..11 = epilogue: tcgen05.ld -> alpha/beta/residual/ReLU/
//   requant -> staged 16-byte stores.  In a pair only the rank-0 CTA issues MMAs; "full" barriers live in the
//   rank-0 CTA (both producers arrive on them), "empty"/"accumulator full" barriers are signalled in both CTAs
//   by a multicast tcgen05.commit.
// Pipelines: A patch ring (a_full/a_empty), weight-step ring (b_full/b_empty), TMEM accumulator
//   stages (acc_full/acc_empty).
#pragma once
#include <cuda.h>
#include <cuda_fp8.h>
#include "sm100_ptx.cuh"

namespace dlq {

constexpr int kMaxSteps = 40;    // K steps (tap x channel-block) per conv
constexpr int kMaxPlanes = 4;
constexpr uint16_t kStepSecond = 0x8000;
constexpr int kTileM = 128;
constexpr int kEpiStageRow = 64;                       // bytes per staged row; 16-byte chunk q of row r lives at chunk
                                                       // q ^ ((r >> 1) & 3): conflict-free both row-per-lane and coalesced
constexpr int kEpiStageBytes = 32 * kEpiStageRow;      // per staging slot (one unit: 32 rows x 64 channels)
__device__ __forceinline__ int epi_stage_off(int row, int chunk) {
  return row * kEpiStageRow + ((chunk ^ ((row >> 1) & 3)) << 4);
}


// Per-role cycle counters (DLQ_DBG_TIMES=1, see conv_plan.cu) exist only in a library built with `make TIMING=1`:
// in the product build every clock read below is a compile-time 0 and the counter arithmetic disappears from the
// producer / issuer / epilogue loops.
__device__ __forceinline__ long long dbg_clock() {
#ifdef DLQ_TIMING
  return clock64();
#else
  return 0;
#endif
}

// tuning switches (ConvKernelParams::dbg) are likewise honoured only by a TIMING build
__device__ __forceinline__ int dbg_flags(int f) {
#ifdef DLQ_TIMING
  return f;
#else
  (void)f;
  return 0;
#endif
}

// Tile-level dependency on another conv launch of the same forward (instead of the grid-level griddepcontrol.wait): the
// producer launch counts, per work-unit position index (super-tile, or super-tile pair), the epilogue warps that have
// stored their part of it; a consumer item waits until the counters of every producer unit its patch (or residual rows)
// touches have reached `target`.  See the "dependency flags" section of the kernel.
struct ConvDep {
  const unsigned int* done;   // producer's counters, one per (super-tile | pair) index
  unsigned int target;        // producer n_tiles * epilogue warps * CTAs per unit
  int unit;                   // producer positions per counter: super_stride * (pair ? 2 : 1)
  int Pv, Wp, Wo, H;          // producer's virtual geometry: image pitch (rows), row pitch, valid width, valid rows
  int s, lo, hi;              // producer rows my output row r needs: [r*s - lo, r*s + hi] (clipped to the image)
};

struct ConvKernelParams {
  // dependency flags (all zero / null: grid-level dependencies through griddepcontrol.wait)
  int n_deps;
  ConvDep dep[2];
  unsigned int* done;         // this launch's own counters (null: none kept)
  unsigned int* dep_err;      // incremented when a dependency wait times out (~4 s): results are then invalid
  // virtual output space
  int Wp, Wo, Ho, Pv, N;
  int total_pos;          // N * Pv * Wp
  int num_super;          // number of super-tiles
  int super_stride;       // positions a super-tile owns (distance between consecutive super-tile origins):
                          // MT*128, or a whole number of virtual rows (<= MT*128) for CTA pairs
  int MT;                 // M tiles (128 positions) per super-tile
  int n_tile;             // UMMA N (output channels per work item)
  int n_tiles;            // OC / n_tile
  int OC;                 // total output channels
  int two;                // 1: CTA pairs (cta_group::2); a pair works on super-tiles 2*sp and 2*sp+1 together
  int n_items;            // work items = (super-tile or super-tile pair) x n-tile, dealt round-robin to CTAs / pairs
  // A patch: n_sub sub-patches (parity plane x channel block); each is ONE 3-D TMA load and one
  // stage of the A ring.  K steps are grouped by sub-patch: sub s owns steps [sub_step0[s], sub_step0[s+1]).
  int n_sub;
  int sub_bytes;          // bytes per sub-patch stage in smem (1024-aligned)
  int tma_bytes;          // bytes written per sub-patch by TMA (NR*Wp*ROWB)
  int row_mul;            // TMA row coordinate = row_mul * v0 + sub_row_off
  int16_t sub_c0[16];     // TMA coordinates per sub-patch: channel byte offset
  int16_t sub_col0[16];   //   start column (may be negative)
  int16_t sub_row_off[16];//   row offset
  int sub_plane_row[16];  //   + first row of the sub-patch's parity plane (plane-layout inputs, else 0)
  int16_t sub_step0[17];  //   first K step of each sub-patch (sub_step0[n_sub] = n_steps)
  // K steps
  int n_steps;
  uint16_t step_a16[kMaxSteps];   // A view offset of each step inside its sub-patch, in 16-byte units (tap shift * ROWB / 16);
                                  // bit 15 (kStepSecond): the step belongs to the fused second conv (see `fused`)
  int a_stages, b_stages, acc_stages;
  int b_resident;         // 1: all weight steps stay in smem (b_stages == n_steps), loaded once per CTA (needs n_tiles == 1)
  int w_rows;             // weight rows (output channels) each CTA loads per step: n_tile, or n_tile/2 in a pair
  uint32_t step_bytes;    // w_rows * ROWB
  // epilogue:  y = clamp(rne(fmaf(r, res_mul, fmaf(acc, alpha[oc], beta[oc]))), lo, 127)
  const float* alpha;     // [OC] requantisation multiplier (output scale folded in)
  const float* beta;      // [OC]
  const int8_t* residual; // row-padded NHWC int8 [.,Ho,Wo,OC] or nullptr
  int res_PR;
  float res_mul;
  int relu;               // lo = relu ? 0 : -128
  int8_t* out;            // row-padded NHWC int8
  int out_PR;
  int out_planes;         // 1: out is stored as four parity planes (see Act::planes); out_rows_half = total rows / 2
  int out_rows_half;
  int32_t* acc_out;       // optional dense NHWC int32 [N,Ho,Wo,OC] raw accumulators (debug / parity)
  int acc_f32;            // 1: acc_out receives fp32 t = fmaf(acc, alpha[oc], beta[oc]) (ReLU if `relu`) instead of the raw
                          // accumulators: the FC / logits form, R/infer_e2e.cu:206-219 (bias after the product)
  unsigned long long* stamps;   // optional [2]: min globaltimer at CTA entry, max at CTA exit (ns) - the launch's span inside
                                // a running step (bench.py roofline); one atomic per CTA at each end
  // Fused second conv (ResNet downsample blocks, R/infer_e2e.cu:181-196): the 1x1/s2 shortcut conv reads exactly the
  // A view of the 3x3/s2 conv's centre tap, so both run from ONE patch load: steps flagged kStepSecond multiply that
  // view with the shortcut's weights into a second accumulator block (TMEM columns + MT*n_tile), and the epilogue
  // writes it through its own alpha/beta/ReLU/output.  Requires MT == 1 and no residual.
  int fused;
  int first_second_step;  // index of the first kStepSecond step (overwrites the second accumulator)
  const float* alpha2;
  const float* beta2;
  int relu2;
  int8_t* out2;
  int out2_PR;
  uint32_t wp_magic, pv_magic;   // floor(2^32/d) for d = Wp, Pv (div_magic)
  int dbg;                // tuning experiments only: 1 = skip the epilogue, 2 = skip MMA issue, 4 = skip A loads
  long long* dbg_times;   // optional [gridDim.x][8] cycle counters (tuning): see conv_plan.cu
};

__device__ __forceinline__ uint32_t pack_sat_s8x4(int a, int b, int c, int d) {
  // cvt.pack.sat.s8.s32.b32 d, a, b, c :  d[7:0] = sat8(b), d[15:8] = sat8(a), d[31:16] = c[15:0]
  uint32_t hi, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;   // bytes (low->high): a, b, c, d
}

// ------------------------------------------------------------------------------------------------
// MMA issuer (one elected lane of an issuing warp).  Everything is a compile-time offset from (a_lo, b_lo, d0),
// the accumulate flag is constant (first step peeled) and nothing data-dependent sits between two tcgen05.mma:
// a branch or an address multiply there costs ~25 cycles of tensor-pipe idle time (probe/mma_rate.cu "morph").
// ------------------------------------------------------------------------------------------------
struct IssuerCtx {
  uint64_t *a_full, *a_empty, *b_full, *b_empty, *acc_full, *acc_empty, *k_first;
  const uint16_t* step_a16;     // shared-memory copy of the per-step A offsets
  const int16_t* sub_step0;
  uint32_t sA_u32, sB_u32, a_stage16, b_stage16;   // ring bases (shared window) and stage sizes in 16-byte units
  uint32_t tmem_base, acc_cols, n_tile;
  int n_sub, a_stages, b_stages, acc_stages, Wp, super_stride, n_tiles;
  int it_begin, it_end, it_stride;
  int fused, first_second_step;
  uint32_t second_off;          // accumulator column offset of the fused second conv
  uint32_t wp_magic;            // ConvKernelParams::wp_magic
  uint32_t tile_off16;          // this issuer's first tile, in 16-byte units down the patch
  uint32_t d_off;               // this issuer's first accumulator column offset
  bool leader;
  int dbg;
};

// FP8: E4M3 x E4M3 -> FP32 (kind::f8f6f4) instead of S8 x S8 -> S32 (kind::i8); same operand bytes, same descriptors
template <bool TWO, bool FP8>
__device__ __forceinline__ void umma_issue(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (FP8) { if (TWO) umma_f8_pair(d, a, b, idesc, acc); else umma_f8(d, a, b, idesc, acc); }
  else { if (TWO) umma_i8_pair(d, a, b, idesc, acc); else umma_i8(d, a, b, idesc, acc); }
}
template <bool TWO>
__device__ __forceinline__ void umma_done(uint64_t* bar) {
  if (TWO) umma_commit_pair(bar); else umma_commit(bar);
}

// KSEL: 0 = every K=32 slice of the step, 1 / 2 = first / second half of them (two issuers sharing ONE accumulator)
template <int ROWB, int MYMT, bool FIRST, bool TWO, int KSEL, bool FP8>
__device__ __forceinline__ void issue_step(uint32_t a_hi, uint32_t b_hi, uint32_t a_lo, uint32_t b_lo, uint32_t d0,
                                           uint32_t n_tile, uint32_t idesc) {
  constexpr int K32 = ROWB / 32;
  constexpr int KK0 = KSEL == 2 ? K32 / 2 : 0;
  constexpr int KK1 = KSEL == 1 ? K32 / 2 : K32;
  constexpr uint32_t TILE16 = kTileM * ROWB / 16;
#pragma unroll
  for (int kk = KK0; kk < KK1; ++kk) {
#pragma unroll
    for (int mt = 0; mt < MYMT; ++mt) {   // tile inner: consecutive MMAs hit different accumulators
      umma_issue<TWO, FP8>(d0 + (mt ? n_tile : 0u), (static_cast<uint64_t>(a_hi) << 32) | (a_lo + mt * TILE16 + 2u * kk),
                      (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2u * kk), idesc, (FIRST && kk == KK0) ? 0u : 1u);
    }
  }
}

// positions of one role in the three rings (patch stages, weight stages, accumulator stages); kept by the caller so
// that a kernel running several layers (conv_chain.cuh) carries them from one layer to the next
struct RingState { uint32_t as = 0, aph = 0, bs = 0, bph = 0, cs = 0, cph = 0; };

// The items  it, it + G, it + 2G, ...  a CTA (pair) walks, decomposed as it = sp * n_tiles + nt (super-tile index, n-tile
// index) without a division per item: every role used to pay one (~100 dependent cycles) per item or per pass.
struct ItemCursor {
  int it, sp, nt, G, dq, dr, n_tiles;
  __device__ __forceinline__ void init(int first_it, int G_, int n_tiles_) {
    it = first_it; G = G_; n_tiles = n_tiles_;
    sp = first_it / n_tiles_; nt = first_it - sp * n_tiles_;
    dq = G_ / n_tiles_; dr = G_ - dq * n_tiles_;
  }
  __device__ __forceinline__ void next() {
    it += G; sp += dq; nt += dr;
    if (nt >= n_tiles) { nt -= n_tiles; ++sp; }
  }
};

// floor(n / d) for 0 <= n < 2^31 without a hardware division: magic = floor(2^32 / d) (0xFFFFFFFF for d = 1, see
// conv_magic in conv_plan.cu) under-estimates the quotient by at most one, which the remainder test repairs
__device__ __forceinline__ int div_magic(int n, int d, uint32_t magic) {
  int q = static_cast<int>(__umulhi(static_cast<uint32_t>(n), magic));
  if (n - q * d >= d) ++q;
  return q;
}

template <int ROWB, int MYMT, bool RESIDENT, bool TWO, int KSEL, bool FP8, bool FUSED = false>
__device__ __forceinline__ void run_issuer(const IssuerCtx& c, long long* t_out, RingState& rs) {
  constexpr uint32_t LAYOUT = ROWB == 128 ? UMMA_SWZ_128B : ROWB == 64 ? UMMA_SWZ_64B : UMMA_SWZ_32B;
  constexpr uint32_t SBO = 8u * ROWB;
  const uint32_t idesc = FP8 ? umma_idesc_e4m3(TWO ? 2 * kTileM : kTileM, c.n_tile) : umma_idesc_s8(TWO ? 2 * kTileM : kTileM, c.n_tile);
  const uint32_t a_hi = static_cast<uint32_t>(umma_smem_desc(0, 0, SBO, LAYOUT) >> 32);
  const uint32_t b_hi = a_hi;
  const uint32_t a_flags = static_cast<uint32_t>(umma_smem_desc(0, 0, SBO, LAYOUT)) + (c.sA_u32 >> 4) + c.tile_off16;
  const uint32_t b_flags = static_cast<uint32_t>(umma_smem_desc(0, 0, SBO, LAYOUT)) + (c.sB_u32 >> 4);
  uint32_t as = rs.as, aph = rs.aph, bs = rs.bs, bph = rs.bph, cs = rs.cs, cph = rs.cph;
  long long t_acc = 0, t_a = 0, t_b = 0;
  bool first_pass = true;
  ItemCursor cur;
  cur.init(c.it_begin, c.it_stride, c.n_tiles);
  for (; cur.it < c.it_end; cur.next()) {
    // offset of the super-tile's first position inside its patch (pairs: super_stride % Wp == 0, so none)
    uint32_t in_patch16 = 0;
    if (!TWO) {
      const int g0 = cur.sp * c.super_stride;
      in_patch16 = static_cast<uint32_t>(g0 - div_magic(g0, c.Wp, c.wp_magic) * c.Wp) * (ROWB / 16);
    }
    long long tw = dbg_clock();
    if (TWO) mbar_wait_cluster(&c.acc_empty[cs], cph ^ 1u); else mbar_wait(&c.acc_empty[cs], cph ^ 1u);
    t_acc += dbg_clock() - tw;
    tc_fence_after();
    const uint32_t d0 = c.tmem_base + cs * c.acc_cols + c.d_off;
    for (int s = 0; s < c.n_sub; ++s) {
      tw = dbg_clock();
      mbar_wait(&c.a_full[as], aph);
      t_a += dbg_clock() - tw;
      tc_fence_after();
      const uint32_t a_base = a_flags + as * c.a_stage16 + in_patch16;
      int k = c.sub_step0[s];
      const int k_end = c.sub_step0[s + 1];
      if (RESIDENT && first_pass) {               // weights arrive once; afterwards no weight barrier at all
        for (int kw = k; kw < k_end; ++kw) {
          tw = dbg_clock();
          mbar_wait(&c.b_full[kw], 0u);
          t_b += dbg_clock() - tw;
        }
        tc_fence_after();
      }
      if (s == 0) {                               // peeled first step of the item: overwrites the accumulators
        if (!RESIDENT) {
          tw = dbg_clock();
          mbar_wait(&c.b_full[bs], bph);
          t_b += dbg_clock() - tw;
          tc_fence_after();
        }
        if (c.leader) {
          const uint32_t b_lo = b_flags + (RESIDENT ? static_cast<uint32_t>(k) : bs) * c.b_stage16;
          // K-split: the second issuer only accumulates, and only after the first one's overwriting MMA has
          // COMPLETED: tcgen05 orders the MMAs of one thread only, so the hand-off is a tcgen05.commit on k_first
          // (arrives when the first issuer's MMAs so far are done) plus tcgen05.fence::after_thread_sync here
          if (KSEL == 2) { mbar_wait(&c.k_first[cs], cph); tc_fence_after(); }
          if (!(dbg_flags(c.dbg) & 2)) issue_step<ROWB, MYMT, KSEL != 2, TWO, KSEL, FP8>(a_hi, b_hi, a_base + c.step_a16[k], b_lo, d0, c.n_tile, idesc);
          if (KSEL == 1) umma_done<TWO>(&c.k_first[cs]);
          if (!RESIDENT) umma_done<TWO>(&c.b_empty[bs]);
        }
        if (!RESIDENT) { if (++bs == static_cast<uint32_t>(c.b_stages)) { bs = 0; bph ^= 1u; } }
        ++k;
      }
      uint32_t so_next = c.step_a16[k];           // (the table has spare entries behind the last step)
      for (; k < k_end; ++k) {
        // next step's view offset: its load latency hides behind this step's MMAs.  Not for the stem (ROWB 32, two
        // short MMAs per step): there the faster issue measurably slows the kernel (107.5 -> 112.5 us), the epilogue
        // being what bounds it
        const uint32_t so = ROWB == 32 ? static_cast<uint32_t>(c.step_a16[k]) : so_next;
        if (ROWB != 32) so_next = c.step_a16[k + 1];
        if (!RESIDENT) {
          tw = dbg_clock();
          mbar_wait(&c.b_full[bs], bph);
          t_b += dbg_clock() - tw;
          tc_fence_after();
        }
        if (c.leader) {
          const uint32_t b_lo = b_flags + (RESIDENT ? static_cast<uint32_t>(k) : bs) * c.b_stage16;
          if (FUSED && (so & kStepSecond)) {
            // fused second conv: all K slices from the first issuer (no cross-issuer ordering on its accumulator),
            // the first such step overwrites
            if (KSEL != 2 && !(dbg_flags(c.dbg) & 2)) {
              constexpr int K32 = ROWB / 32;
#pragma unroll
              for (int kk = 0; kk < K32; ++kk)
                umma_issue<TWO, FP8>(d0 + c.second_off, (static_cast<uint64_t>(a_hi) << 32) | (a_base + (so & 0x7FFFu) + 2u * kk),
                                     (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2u * kk), idesc,
                                     (k == c.first_second_step && kk == 0) ? 0u : 1u);
            }
          } else if (!(dbg_flags(c.dbg) & 2)) {
            issue_step<ROWB, MYMT, false, TWO, KSEL, FP8>(a_hi, b_hi, a_base + so, b_lo, d0, c.n_tile, idesc);
          }
          if (!RESIDENT) umma_done<TWO>(&c.b_empty[bs]);
        }
        if (!RESIDENT) { if (++bs == static_cast<uint32_t>(c.b_stages)) { bs = 0; bph ^= 1u; } }
      }
      if (dbg_flags(c.dbg) & 8) { tw = dbg_clock(); }
      if (c.leader) umma_done<TWO>(&c.a_empty[as]);     // sub-patch stage free
      if (dbg_flags(c.dbg) & 8) { __syncwarp(); t_b += dbg_clock() - tw; }
      if (++as == static_cast<uint32_t>(c.a_stages)) { as = 0; aph ^= 1u; }
    }
    if (c.leader) umma_done<TWO>(&c.acc_full[cs]);      // accumulators ready for the epilogue
    __syncwarp();
    first_pass = false;
    if (++cs == static_cast<uint32_t>(c.acc_stages)) { cs = 0; cph ^= 1u; }
  }
  t_out[0] = t_acc; t_out[1] = t_a; t_out[2] = t_b;
  rs.as = as; rs.aph = aph; rs.bs = bs; rs.bph = bph; rs.cs = cs; rs.cph = cph;
}


// ------------------------------------------------------------------------------------------------
// Dependency flags.  Consecutive conv launches of one forward are chained with programmatic dependent launch and, in
// this mode, never call griddepcontrol.wait: CTAs of launch L+1 become resident as CTAs of L exit and start on whatever
// items already have their inputs, so the tail of L (partial last wave, last epilogues, grid teardown) overlaps the
// head of L+1.  No deadlock: L+1 is only scheduled once every CTA of L has started (launch_dependents at CTA entry),
// and L's CTAs never wait for L+1.
//   producer:  each epilogue warp, after its stores of an item, counts itself in shared memory; a signaller lane fences
//              and adds the item's 8 arrivals to done[unit] (see "Producer side" below).
//   consumer:  warp 0, before the item's first patch load: all lanes poll (ld.acquire.gpu) the counters of the producer
//              units its rows touch, then the elected lane fences (gpu scope + generic->async proxy) and issues the TMA.
//   The counters are zeroed by the LAST kernel of the forward (GAP+FC, which waits for the whole last conv grid), so a
//   forward always starts with zero counters and nothing of the previous forward can still increment them.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// producer units [lo, hi] whose rows the consumer positions [g0, g1] need (lo > hi: none)
__device__ __forceinline__ void dep_unit_range(const ConvKernelParams& p, const ConvDep& d, int g0, int g1, int& lo, int& hi) {
  int va = div_magic(g0, p.Wp, p.wp_magic), vb = div_magic(g1, p.Wp, p.wp_magic);
  int na = div_magic(va, p.Pv, p.pv_magic), ra = va - na * p.Pv;
  int nb = div_magic(vb, p.Pv, p.pv_magic), rb = vb - nb * p.Pv;
  if (ra >= p.Ho) { ++na; ra = 0; }                       // starts in the pad rows behind an image
  if (nb >= p.N) { nb = p.N - 1; rb = p.Ho - 1; }
  if (rb >= p.Ho) rb = p.Ho - 1;
  if (na > nb || (na == nb && ra > rb)) { lo = 1; hi = 0; return; }
  const int pra = max(0, ra * d.s - d.lo), prb = min(d.H - 1, rb * d.s + d.hi);
  lo = ((na * d.Pv + pra) * d.Wp) / d.unit;
  hi = ((nb * d.Pv + prb) * d.Wp + d.Wo - 1) / d.unit;
}

// whole warp: wait until every dependency of the positions [g0, g1] is satisfied.  The counters of all dependencies are
// polled by different lanes at once (one L2 round trip when everything is ready).
__device__ __forceinline__ void wait_deps(const ConvKernelParams& p, int g0, int g1, int lane) {
  int lo0 = 1, hi0 = 0, lo1 = 1, hi1 = 0;
  dep_unit_range(p, p.dep[0], g0, g1, lo0, hi0);
  if (p.n_deps > 1) dep_unit_range(p, p.dep[1], g0, g1, lo1, hi1);
  const int c0 = max(hi0 - lo0 + 1, 0), c1 = max(hi1 - lo1 + 1, 0);
  for (int i = lane; i < c0 + c1; i += 32) {
    const bool second = i >= c0;
    const unsigned int* ctr = second ? p.dep[1].done + lo1 + (i - c0) : p.dep[0].done + lo0 + i;
    const unsigned int target = second ? p.dep[1].target : p.dep[0].target;
    if (ld_acquire_gpu(ctr) >= target) continue;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned int spins = 0;
    while (ld_acquire_gpu(ctr) < target) {
      if ((++spins & 1023u) == 0) {                     // a lost producer must not hang the GPU: give up after ~4 s
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ULL) { atomicAdd(p.dep_err, 1u); break; }
      }
    }
  }
  __syncwarp();
}

// Producer side.  The gpu-scope fence in front of the counter update has to wait for outstanding stores (of the whole
// SM, in practice: ~0.8 us per item when issued by an epilogue warp, measured), so the epilogue warps do not issue it:
// each warp, after its stores of an item, adds 1 to a SHARED-memory count (red.release.cta, ring of 4 slots, counts are
// cumulative so nothing is ever reset), and a spare lane of the weight-producer warp - the "signaller" - waits for the
// eight arrivals of item k (acquire.cta), fences at gpu scope (cumulative: it orders the epilogue warps' stores it has
// observed through the shared count) and publishes the item with ONE atomicAdd of 8 on the global counter.  Epilogue
// warps are at most two items apart (two TMEM stages), so an arrival for item k+4 can only come after item k is complete.
constexpr int kStoredSlots = 4;

// ------------------------------------------------------------------------------------------------
// Epilogue (8 warps).  A "unit" is (M tile, 64-channel block): per unit a warp turns its 32 accumulator rows x 64
// int32 into 32 x 64 int8.  Units are processed in PAIRS so that two tcgen05.ld are in flight and - when the two
// units are the same channel block of two tiles - every alpha/beta shared-memory read serves both.
// Global traffic goes through per-warp staging slots (one per unit of the pair) so that every LDG/STG covers whole
// 64-byte pixel rows: the residual rows of the NEXT pair are prefetched into the slots with cp.async (no registers
// in flight, latency hidden behind the accumulator wait), read back row-per-lane, and the results are staged
// row-per-lane in the same slot, read back coalesced and stored.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kInvalidPix = 0xFFFFFFFFu;

struct EpiCtx {
  const float* s_alpha;
  const float* s_beta;
  uint8_t* slots;          // this warp's staging: [2][32 rows][kEpiStageRow]
  int lane, crow, cq;      // coalesced phase: row within a group of 8, 16-byte quarter
  int out_pitch, res_pitch, n0;
  int8_t* out;             // output tensor of the current unit (the fused second conv has its own)
  int out_PR;
  uint32_t relu_mask;      // 0xFFFFFFFF: ReLU (negative bytes -> 0), 0: none
  float res_mul;
};

__device__ __forceinline__ bool decode_pos(const ConvKernelParams& p, int g, int& n, int& r, int& x) {
  const int vrow = div_magic(g, p.Wp, p.wp_magic);
  n = div_magic(vrow, p.Pv, p.pv_magic);
  x = g - vrow * p.Wp;
  r = vrow - n * p.Pv;
  return (x < p.Wo) && (r < p.Ho) && (n < p.N);
}

// this lane's accumulator row at position g: pixel index of its output row in the (row-padded NHWC, or parity-plane) output
// tensor and of its residual row; kInvalidPix for garbage positions (pad columns / pad rows / behind the last image)
__device__ __forceinline__ void pos_to_pix(const ConvKernelParams& p, int g, int out_PR, uint32_t& opix, uint32_t& rpix) {
  int n, r, x;
  const bool valid = decode_pos(p, g, n, r, x);
  if (p.out_planes) {       // (never combined with the fused second conv)
    const int prow = out_PR + n * (p.Ho + out_PR) + r;
    const int pl = ((prow & 1) << 1) | (x & 1);
    opix = valid ? static_cast<uint32_t>((pl * p.out_rows_half + (prow >> 1)) * (p.Wo >> 1) + (x >> 1)) : kInvalidPix;
  } else {
    opix = valid ? static_cast<uint32_t>((out_PR + n * (p.Ho + out_PR) + r) * p.Wo + x) : kInvalidPix;
  }
  rpix = valid ? static_cast<uint32_t>((p.res_PR + n * (p.Ho + p.res_PR) + r) * p.Wo + x) : kInvalidPix;
}

// cp.async the 32 residual rows (64 B each) of one unit into a staging slot; rows outside the tensor are zero-filled.
// rpix = this lane's residual pixel (pos_to_pix), chan = first channel of the unit
__device__ __forceinline__ void epi_prefetch_res(const ConvKernelParams& p, const EpiCtx& e, int slot, uint32_t rpix, int chan) {
  uint8_t* base = e.slots + slot * kEpiStageBytes;
  const int8_t* col = p.residual + chan + e.cq * 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int srow = 8 * j + e.crow;
    const uint32_t rp = __shfl_sync(0xffffffffu, rpix, srow);
    const bool ok = rp != kInvalidPix;
    cp_async16_zfill(smem_u32(base + epi_stage_off(srow, e.cq)), col + (ok ? static_cast<size_t>(rp) * p.OC : 0), ok ? 16u : 0u);
  }
}

// 16-byte global store under a predicate (no branch around the address arithmetic)
__device__ __forceinline__ void st_global_v4_if(bool ok, void* ptr, const int4& v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v4.b32 [%1], {%2, %3, %4, %5};\n\t}"
               :: "r"(static_cast<int>(ok)), "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// byte K of w (value + 128, i.e. already XORed with 0x80) -> float(value), exact: 0x4B0000uu is 2^23 + uu
template <int K>
__device__ __forceinline__ float biased_byte_to_float(uint32_t w) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(0x7650u + K));
  return __int_as_float(static_cast<int>(r)) - 8388736.0f;
}

// two independent round-to-nearest fp32 FMAs in one instruction (sm_100 FFMA2); bit-identical to two fmaf
__device__ __forceinline__ void fma2_rn(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// E4M3 helpers (QUANT_SPEC section 6): float -> e4m3 is round-to-nearest-even, saturating to +-448
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
__device__ __forceinline__ float2 e4m3x2_to_float2(uint32_t two_bytes) {
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(two_bytes), __NV_E4M3);
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}

// opix = this lane's output pixel per unit (pos_to_pix, computed one pass ahead by the caller); g_own is only read by the
// raw-accumulator form
template <bool HAS_RES, int NU, bool SAME_CB, bool ACC_OUT, bool FP8>
__device__ __forceinline__ void epi_units(const ConvKernelParams& p, const EpiCtx& e, const uint32_t (&taddr)[2],
                                          const int (&c0)[2], const uint32_t (&opix)[2], const int (&g_own)[2],
                                          bool release_acc, uint32_t acc_empty_addr) {
  size_t dpix[NU];
  bool valid[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    valid[u] = opix[u] != kInvalidPix;
    dpix[u] = 0;
    if (ACC_OUT) {
      int n, r, x;
      decode_pos(p, g_own[u], n, r, x);
      dpix[u] = (static_cast<size_t>(n) * p.Ho + r) * p.Wo + x;
    }
  }
  uint8_t* my_row[NU];    // own staged row (row = lane); chunk q sits at my_row + ((q ^ my_swz) << 4)
  const int my_swz = (e.lane >> 1) & 3;
#pragma unroll
  for (int u = 0; u < NU; ++u) my_row[u] = e.slots + u * kEpiStageBytes + e.lane * kEpiStageRow;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    uint32_t v[NU][32];
#pragma unroll
    for (int u = 0; u < NU; ++u) tmem_ld_32x32b_x32(taddr[u] + h * 32, v[u]);
    int4 rv[NU][2];
    if (HAS_RES) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        rv[u][0] = *reinterpret_cast<const int4*>(my_row[u] + (((2 * h) ^ my_swz) << 4));
        rv[u][1] = *reinterpret_cast<const int4*>(my_row[u] + (((2 * h + 1) ^ my_swz) << 4));
      }
    }
    tmem_ld_wait();
    if (release_acc && h == 1) {      // accumulators are in registers: hand the TMEM stage back to the MMA warps
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) mbar_arrive_cluster(acc_empty_addr);     // (rank-0 CTA's barrier when running as a pair)
    }
    if (ACC_OUT) {                    // raw accumulators (parity / debug entry points), or the fp32 FC form
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (valid[u]) {
          int4* dst = reinterpret_cast<int4*>(p.acc_out + dpix[u] * p.OC + e.n0 + c0[u] + h * 32);
          if (p.acc_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 al = reinterpret_cast<const float4*>(e.s_alpha + c0[u] + h * 32)[j];
              const float4 be = reinterpret_cast<const float4*>(e.s_beta + c0[u] + h * 32)[j];
              float t[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float a = FP8 ? __uint_as_float(v[u][4 * j + i]) : static_cast<float>(static_cast<int32_t>(v[u][4 * j + i]));
                t[i] = __fmaf_rn(a, i == 0 ? al.x : i == 1 ? al.y : i == 2 ? al.z : al.w,
                                 i == 0 ? be.x : i == 1 ? be.y : i == 2 ? be.z : be.w);
                if (e.relu_mask && t[i] < 0.f) t[i] = 0.f;
              }
              dst[j] = make_int4(__float_as_int(t[0]), __float_as_int(t[1]), __float_as_int(t[2]), __float_as_int(t[3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_int4((int)v[u][4 * j], (int)v[u][4 * j + 1], (int)v[u][4 * j + 2], (int)v[u][4 * j + 3]);
          }
        }
    }
    uint32_t packed[NU][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 al[NU], be[NU];
      if (dbg_flags(p.dbg) & 16) {      // (tuning: no shared-memory reads for the per-channel constants; wrong results)
        al[0] = make_float4(0.01f, 0.02f, 0.03f, 0.04f); be[0] = make_float4(1.f, 2.f, 3.f, 4.f);
      } else {
        al[0] = reinterpret_cast<const float4*>(e.s_alpha + c0[0] + h * 32)[j];
        be[0] = reinterpret_cast<const float4*>(e.s_beta + c0[0] + h * 32)[j];
      }
      if (NU == 2) {
        if (SAME_CB) { al[NU - 1] = al[0]; be[NU - 1] = be[0]; }
        else {
          al[NU - 1] = reinterpret_cast<const float4*>(e.s_alpha + c0[NU - 1] + h * 32)[j];
          be[NU - 1] = reinterpret_cast<const float4*>(e.s_beta + c0[NU - 1] + h * 32)[j];
        }
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float t0, t1, t2, t3;
        if (FP8) {
          fma2_rn(t0, t1, __uint_as_float(v[u][4 * j + 0]), __uint_as_float(v[u][4 * j + 1]), al[u].x, al[u].y, be[u].x, be[u].y);
          fma2_rn(t2, t3, __uint_as_float(v[u][4 * j + 2]), __uint_as_float(v[u][4 * j + 3]), al[u].z, al[u].w, be[u].z, be[u].w);
        } else {
          fma2_rn(t0, t1, static_cast<float>(static_cast<int32_t>(v[u][4 * j + 0])),
                  static_cast<float>(static_cast<int32_t>(v[u][4 * j + 1])), al[u].x, al[u].y, be[u].x, be[u].y);
          fma2_rn(t2, t3, static_cast<float>(static_cast<int32_t>(v[u][4 * j + 2])),
                  static_cast<float>(static_cast<int32_t>(v[u][4 * j + 3])), al[u].z, al[u].w, be[u].z, be[u].w);
        }
        if (HAS_RES) {
          const uint32_t w = reinterpret_cast<const uint32_t*>(rv[u])[j];
          if (FP8) {
            const float2 r01 = e4m3x2_to_float2(w & 0xFFFFu), r23 = e4m3x2_to_float2(w >> 16);
            fma2_rn(t0, t1, r01.x, r01.y, e.res_mul, e.res_mul, t0, t1);
            fma2_rn(t2, t3, r23.x, r23.y, e.res_mul, e.res_mul, t2, t3);
          } else {
            // int8 -> float without the conversion pipe: byte + 128 dropped into the mantissa of 2^23 (PRMT), then
            // an exact subtraction of 2^23 + 128
            const uint32_t wb = w ^ 0x80808080u;
            fma2_rn(t0, t1, biased_byte_to_float<0>(wb), biased_byte_to_float<1>(wb), e.res_mul, e.res_mul, t0, t1);
            fma2_rn(t2, t3, biased_byte_to_float<2>(wb), biased_byte_to_float<3>(wb), e.res_mul, e.res_mul, t2, t3);
          }
        }
        const uint32_t q = FP8 ? pack_e4m3x4(t0, t1, t2, t3)
                               : pack_sat_s8x4(__float2int_rn(t0), __float2int_rn(t1), __float2int_rn(t2), __float2int_rn(t3));
        // ReLU on the packed bytes: replicate each byte's sign (PRMT mode 8+i) and clear the negative ones
        uint32_t sgn;
        asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(sgn) : "r"(q));
        asm("lop3.b32 %0, %1, %2, %3, 0x70;" : "=r"(packed[u][j]) : "r"(q), "r"(sgn), "r"(e.relu_mask));   // q & ~(sgn & mask)
      }
    }
    if (!(dbg_flags(p.dbg) & 32))       // (tuning: 32 = no staging stores / loads, no global stores; wrong results)
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      *reinterpret_cast<int4*>(my_row[u] + (((2 * h) ^ my_swz) << 4)) =
          make_int4((int)packed[u][0], (int)packed[u][1], (int)packed[u][2], (int)packed[u][3]);
      *reinterpret_cast<int4*>(my_row[u] + (((2 * h + 1) ^ my_swz) << 4)) =
          make_int4((int)packed[u][4], (int)packed[u][5], (int)packed[u][6], (int)packed[u][7]);
    }
  }
  __syncwarp();
  if (e.out && !(dbg_flags(p.dbg) & 32)) {
    // all shuffles, then all (unconditional) staged reads, then the predicated stores (no branches): no serialised
    // shuffle -> load -> store chains
    uint32_t op[NU][4];
    int4 val[NU][4];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) op[u][j] = __shfl_sync(0xffffffffu, opix[u], 8 * j + e.crow);
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        val[u][j] = *reinterpret_cast<const int4*>(e.slots + u * kEpiStageBytes + epi_stage_off(8 * j + e.crow, e.cq));
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      int8_t* col = e.out + e.n0 + c0[u] + e.cq * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_global_v4_if(op[u][j] != kInvalidPix, col + static_cast<size_t>(op[u][j]) * p.OC, val[u][j]);
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Epilogue warps: the items of ONE layer (the whole kernel for conv_i8_kernel; conv_chain.cuh calls it once per layer,
// with `er` carrying the accumulator-ring position and the CTA's item count across layers).
// The two warp groups take alternate units; the four warps of a group own the four TMEM lane quarters.
// Flag mode (p.n_deps != 0): the residual rows of an item are requested only once the patch producer has verified that
// item's dependencies (it publishes the number of verified items of the CTA in s_dep_seq); stores need no wait - every
// tensor of a forward has its own buffer, and the previous forward's readers finished before this forward began.
// ------------------------------------------------------------------------------------------------
struct EpiShared {
  float *s_alpha, *s_beta, *s_alpha2, *s_beta2;
  uint8_t* s_stage;
  uint64_t *acc_full, *acc_empty;
  uint32_t *s_dep_seq, *s_stored;
  uint32_t tmem_base, acc_cols;
  int acc_stages;
};
struct EpiRing { uint32_t cs = 0, cph = 0, items = 0; };

template <bool TWO, bool FP8>
__device__ __forceinline__ void run_epilogue_items(const ConvKernelParams& p, const EpiShared& sh, EpiRing& er, int warp, int lane,
                                                   int n_epi_warps, int first_it, int G, int rank, long long& t_wait) {
  const int ew = warp - 4;
  const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
  const int n_groups = n_epi_warps >> 2;
  const int grp = ew >> 2;
  const int row = quarter * 32 + lane;                // accumulator row within the tile
  const bool has_res = p.residual != nullptr;
  const int cshift = p.n_tile >> 7;                   // log2 of the 64-channel blocks per n-tile (n_tile is 64, 128 or 256)
  const int cmask = (1 << cshift) - 1;
  const int units1 = p.MT << cshift;                  // units of the (first) conv; the fused second conv adds as many
  const int n_units = units1 * (p.fused ? 2 : 1);
  const int upw = (n_units - grp + n_groups - 1) / n_groups;   // units of this warp per item: grp, grp + n_groups, ...
  const int upp = (p.acc_out || p.fused) ? 1 : 2;     // units per pass (raw-accumulator output / fused second conv: one at a time)
  const int n_pairs = (upw + upp - 1) / upp;
  const uint32_t tmem_base = sh.tmem_base, acc_cols = sh.acc_cols;
  EpiCtx e;
  e.s_alpha = sh.s_alpha; e.s_beta = sh.s_beta;
  e.slots = sh.s_stage + ew * 2 * kEpiStageBytes;
  e.lane = lane; e.crow = lane >> 2; e.cq = lane & 3;
  e.out_pitch = p.Ho + p.out_PR; e.res_pitch = p.Ho + p.res_PR; e.n0 = 0;
  e.out = p.out; e.out_PR = p.out_PR;
  e.relu_mask = p.relu ? 0xFFFFFFFFu : 0u;
  e.res_mul = p.res_mul;
  // global position of this lane's row in tile mt of super-tile st (rows past the super-tile's own positions
  // belong to the next super-tile: mapped to total_pos, which decodes as image N = invalid)
  auto own_pos = [&](int st, int mt) {
    const int local = mt * kTileM + row;
    return local < p.super_stride ? st * p.super_stride + local : p.total_pos;
  };
  // Pass descriptors are made ONE PASS AHEAD - after a pass's stores, i.e. while the warp would otherwise only wait for
  // the next accumulator: the output pixel of this lane's row in each unit of the pass (kept in opix_n) and, with a
  // residual, the cp.async prefetch of the units' residual rows into the staging slots.
  uint32_t opix_n[2] = {kInvalidPix, kInvalidPix};
  auto prepare = [&](int sp, int nt, int pi) {
    const int st = TWO ? 2 * sp + rank : sp;
    const bool second = grp + upp * pi * n_groups >= units1;       // (a fused pass never mixes the two convs)
    const int oPR = second ? p.out2_PR : p.out_PR;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = upp * pi + u;
      opix_n[u] = kInvalidPix;
      if (u < upp && k < upw) {
        int unit = grp + k * n_groups;
        if (second) unit -= units1;
        uint32_t rpix;
        pos_to_pix(p, own_pos(st, unit >> cshift), oPR, opix_n[u], rpix);
        if (has_res) epi_prefetch_res(p, e, u, rpix, nt * p.n_tile + ((unit & cmask) << 6));
      }
    }
    if (has_res) cp_async_commit();
  };
  uint32_t cs = er.cs, cph = er.cph;
  const bool flag_mode = p.n_deps != 0;
  auto wait_dep_seq = [&](uint32_t want) {             // until the patch producer has verified `want` items of this CTA
    if (flag_mode && has_res) {
      uint32_t v;
      do {
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(sh.s_dep_seq)) : "memory");
      } while (v < want);
    }
  };
  ItemCursor cur;
  cur.init(first_it, G, p.n_tiles);
  if (cur.it < p.n_items && n_pairs > 0) { wait_dep_seq(er.items + 1u); prepare(cur.sp, cur.nt, 0); }
  while (cur.it < p.n_items) {
    const int sp = cur.sp, nt = cur.nt;
    const int st = TWO ? 2 * sp + rank : sp;
    cur.next();                                        // (cur is now the CTA's next item)
    const uint32_t acc_empty_addr = TWO ? leader_cta_addr(&sh.acc_empty[cs]) : smem_u32(&sh.acc_empty[cs]);
    const long long tw = dbg_clock();
    mbar_wait(&sh.acc_full[cs], cph);
    t_wait += dbg_clock() - tw;
    tc_fence_after();
    const int np = (dbg_flags(p.dbg) & 1) ? 0 : n_pairs;
    e.n0 = nt * p.n_tile;
    for (int pi = 0; pi < np; ++pi) {
      uint32_t taddr[2];
      int c0[2], g_own[2];
      const uint32_t opix[2] = {opix_n[0], opix_n[1]};
      const int nu = (upp == 2 && 2 * pi + 1 < upw) ? 2 : 1;
      // (fused launches process one unit per pass, so a pass never mixes the two convs)
      const bool second = grp + upp * pi * n_groups >= units1;
      e.s_alpha = (second ? sh.s_alpha2 : sh.s_alpha) + e.n0;
      e.s_beta = (second ? sh.s_beta2 : sh.s_beta) + e.n0;
      e.out = second ? p.out2 : p.out;
      e.relu_mask = (second ? p.relu2 : p.relu) ? 0xFFFFFFFFu : 0u;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int unit = grp + (upp * pi + (u < nu ? u : 0)) * n_groups;
        if (second) unit -= units1;
        const int mt = unit >> cshift;
        c0[u] = (unit & cmask) << 6;
        g_own[u] = p.acc_out ? own_pos(st, mt) : 0;
        taddr[u] = tmem_base + cs * acc_cols + (second ? static_cast<uint32_t>(p.MT) * p.n_tile : 0u) +
                   static_cast<uint32_t>(mt) * p.n_tile + c0[u] + (static_cast<uint32_t>(quarter * 32) << 16);
      }
      const bool last = pi == np - 1;
      if (has_res) {
        cp_async_wait_all();
        __syncwarp();
      }
#define DLQ_EPI(RES, NUV, SAME, ACC) epi_units<RES, NUV, SAME, ACC, FP8>(p, e, taddr, c0, opix, g_own, last, acc_empty_addr)
#define DLQ_EPI_SHAPES(RES)                                       \
  do {                                                            \
    if (nu == 1) DLQ_EPI(RES, 1, true, false);                    \
    else if (c0[0] == c0[1]) DLQ_EPI(RES, 2, true, false);        \
    else DLQ_EPI(RES, 2, false, false);                           \
  } while (0)
      if (p.acc_out) { if (has_res) DLQ_EPI(true, 1, true, true); else DLQ_EPI(false, 1, true, true); }
      else if (has_res) DLQ_EPI_SHAPES(true);
      else DLQ_EPI_SHAPES(false);
#undef DLQ_EPI_SHAPES
#undef DLQ_EPI
      // the next pass's descriptors (and residual rows): of this item, or of the CTA's next item
      if (pi + 1 < n_pairs) prepare(sp, nt, pi + 1);
      else if (cur.it < p.n_items) { wait_dep_seq(er.items + 2u); prepare(cur.sp, cur.nt, 0); }
    }
    // this warp's part of the item is stored: tell the signaller (see "Producer side"; it publishes only if p.done)
    __syncwarp();
    if (lane == 0)
      asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(sh.s_stored + (er.items % kStoredSlots))) : "memory");
    ++er.items;
    if (np == 0) {                                    // (debug: epilogue skipped) still hand the stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_addr);
    }
    if (++cs == static_cast<uint32_t>(sh.acc_stages)) { cs = 0; cph ^= 1u; }
  }
  if (has_res) cp_async_wait_all();
  er.cs = cs; er.cph = cph;
}

// smem layout (dynamic, 1024-aligned base):
//   [A ring: a_stages * sub_bytes][B ring: b_stages * step_bytes(1024-aligned)][alpha, beta: 2*OC f32]
//   [epilogue staging: 8 warps * 2 slots * kEpiStageBytes][step offsets][barriers][tmem slot]
template <int ROWB, bool TWO, bool FP8>
__global__ void __launch_bounds__(384, 1)
conv_i8_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tmw, const ConvKernelParams p) {
  constexpr uint32_t TILE16 = kTileM * ROWB / 16;      // one M tile further down the patch, in 16-byte units

  const long long t_entry = dbg_clock();
  if (p.stamps && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    atomicMin(p.stamps, gt);
  }
  if (p.dbg_times && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg_times[static_cast<size_t>(blockIdx.x) * 16 + 12] = static_cast<long long>(gt);     // CTA entry, ns
  }
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B with pointer arithmetic on the __shared__ array (keeps the address space visible to
  // the compiler so alpha/beta reads become LDS, not generic loads)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_stage_bytes = static_cast<uint32_t>(p.sub_bytes);
  const uint32_t b_stage_bytes = (p.step_bytes + 1023u) & ~1023u;
  uint8_t* sA = smem;
  uint8_t* sB = sA + static_cast<size_t>(p.a_stages) * a_stage_bytes;
  float* s_alpha = reinterpret_cast<float*>(sB + static_cast<size_t>(p.b_stages) * b_stage_bytes);
  float* s_beta = s_alpha + p.OC;
  float* s_alpha2 = s_beta + p.OC;                                           // (fused second conv)
  float* s_beta2 = s_alpha2 + p.OC;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_beta2 + p.OC);             // [epilogue warps][2 slots][kEpiStageBytes]
  uint16_t* s_step_a16 = reinterpret_cast<uint16_t*>(s_stage + 16 * kEpiStageBytes);   // [kMaxSteps]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_step_a16 + kMaxSteps + 8);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint64_t* acc_empty = acc_full + p.acc_stages;
  uint64_t* k_first = acc_empty + p.acc_stages;       // K-split handshake, one per accumulator stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_first + p.acc_stages);
  volatile uint32_t* s_dep_seq = tmem_slot + 1;       // items of this CTA whose dependencies the patch producer has verified
  uint32_t* s_stored = tmem_slot + 2;                 // [kStoredSlots] cumulative epilogue-warp arrivals per item slot

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_epi_warps = (blockDim.x >> 5) - 4;
  // Two MMA-issuing warps (1 and 3): a single one cannot keep the tensor pipe fed (the per-step barrier wait and
  // descriptor arithmetic are exposed, probe/mma_rate.cu "pattern").  They split the M tiles of a super-tile when
  // there are at least two, else the K=32 slices of every step (same accumulator; integer accumulation commutes,
  // the overwriting first MMA is ordered by a tcgen05.commit hand-off).  FP32 accumulation does not commute: the
  // E4M3 instantiation keeps ONE issuer per accumulator, so its results do not depend on timing.
  constexpr int K32 = ROWB / 32;
  const bool k_split = !FP8 && p.MT == 1 && K32 >= 2;
  const int n_issuers = (p.MT >= 2 || k_split) ? 2 : 1;
  const int ncta = TWO ? 2 : 1;
  const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
  const int gid = static_cast<int>(blockIdx.x) / ncta;          // CTA (or pair) index
  const int G = static_cast<int>(gridDim.x) / ncta;             // CTAs (or pairs) in the grid
  const uint32_t acc_cols = static_cast<uint32_t>(p.MT) * p.n_tile * (p.fused ? 2u : 1u);   // TMEM columns per accumulator stage
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_cols * p.acc_stages) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    pdl_launch_dependents();   // the next kernel on the stream may begin its prologue / weight loads as SMs free up
    // "full" barriers: one arrive (+ the transaction bytes of every CTA of the pair); "empty" / acc_full ones one
    // commit per issuer; acc_empty one arrive per epilogue warp of every CTA of the pair
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], n_issuers); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_issuers); }
    for (int i = 0; i < p.acc_stages; ++i) {
      mbar_init(&acc_full[i], n_issuers);
      mbar_init(&acc_empty[i], n_epi_warps * ncta);
      mbar_init(&k_first[i], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tmw);
    *s_dep_seq = 0u;
    for (int i = 0; i < kStoredSlots; ++i) s_stored[i] = 0u;
  }
  if (warp == 1) {
    if (TWO) { tmem_alloc_pair(tmem_slot, tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i < p.n_steps; i += blockDim.x) s_step_a16[i] = p.step_a16[i];
  // (alpha/beta are fetched by the epilogue warps after the CTA-wide sync, off the producers' critical path)
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.dbg_times && threadIdx.x == 0) p.dbg_times[static_cast<size_t>(blockIdx.x) * 16 + 8] = dbg_clock() - t_entry;   // prologue

  if (warp == 0) {
    // ===================================================================== A (activation patch) producer
    // One elected lane issues the loads; with dependency flags the whole warp first polls the producer counters of the
    // item (lanes in parallel), so the other lanes walk the item loop too.
    const bool leader = elect_one();
    uint32_t as = 0, aph = 0;
    long long t_wait = 0;
    const long long t_begin = dbg_clock();
    if (p.n_deps == 0 && leader) pdl_wait();   // grid-level dependency: activations come from the previous kernel(s);
                                                // everything above (and the weight loads) does not
    uint32_t seq = 0;
    ItemCursor cur;
    cur.init(gid, G, p.n_tiles);
    for (; cur.it < p.n_items; cur.next()) {
      const int st = TWO ? 2 * cur.sp + rank : cur.sp;
      if (p.n_deps) {
        const int g0 = st * p.super_stride;
        wait_deps(p, g0, min(g0 + p.super_stride, p.total_pos) - 1, lane);
      }
      if (leader) {
        // (the polling lanes' acquires reach this lane through wait_deps' __syncwarp: causality order is transitive)
        if (p.n_deps) fence_proxy_async_all();   // producer's generic-proxy stores before this thread's async-proxy (TMA) reads
        ++seq;                                   // the epilogue warps may now prefetch this item's residual rows
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(s_dep_seq))), "r"(seq) : "memory");
        const int v0 = div_magic(st * p.super_stride, p.Wp, p.wp_magic);
        for (int s = 0; s < p.n_sub; ++s) {
          const long long tw = dbg_clock();
          mbar_wait(&a_empty[as], aph ^ 1u);
          t_wait += dbg_clock() - tw;
          uint8_t* dst = sA + static_cast<size_t>(as) * a_stage_bytes;
          if (dbg_flags(p.dbg) & 4) {
            if (rank == 0) mbar_arrive(&a_full[as]);
          } else if (TWO) {
            // the rank-0 producer alone arrives, expecting BOTH CTAs' bytes; the peer's TMA completes its
            // transaction bytes on the rank-0 barrier (a remote arrive per load would cost a cluster round trip)
            if (rank == 0) mbar_expect_tx(&a_full[as], 2u * static_cast<uint32_t>(p.tma_bytes));
            tma_load_3d_pair(dst, &tm0, leader_cta_addr(&a_full[as]), p.sub_c0[s], p.sub_col0[s],
                             p.row_mul * v0 + p.sub_row_off[s] + p.sub_plane_row[s]);
          } else {
            mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.tma_bytes));
            tma_load_3d(dst, &tm0, &a_full[as], p.sub_c0[s], p.sub_col0[s],
                        p.row_mul * v0 + p.sub_row_off[s] + p.sub_plane_row[s]);
          }
          if (++as == static_cast<uint32_t>(p.a_stages)) { as = 0; aph ^= 1u; }
        }
      }
    }
    if (p.dbg_times && leader) {
      long long* d = p.dbg_times + static_cast<size_t>(blockIdx.x) * 16;
      d[6] = dbg_clock() - t_begin; d[7] = t_wait;
    }
  } else if (warp == 2) {
    // ===================================================================== B (weight step) producer
    // weight image rows: [(n-tile * n_steps + step) * n_tile + channel], ROWB bytes each (pre-swizzled); a CTA of a
    // pair loads its half of the channels
    const bool b_leader = elect_one();
    const int sig_lane = (__ballot_sync(0xffffffffu, b_leader) & 0x80000000u) ? 30 : 31;     // a lane that is not the leader
    if (!b_leader && lane == sig_lane && p.done) {
      // ------------------------------------------------------------------- signaller (see "Producer side")
      uint32_t k = 0;
      ItemCursor cur;
      cur.init(gid, G, p.n_tiles);
      for (; cur.it < p.n_items; cur.next(), ++k) {
        const uint32_t want = static_cast<uint32_t>(n_epi_warps) * ((k / kStoredSlots) + 1u);
        const uint32_t addr = smem_u32(s_stored + (k % kStoredSlots));
        uint32_t v;
        do {
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        } while (v < want);
        __threadfence();
        atomicAdd(p.done + cur.sp, static_cast<unsigned int>(n_epi_warps));
      }
    }
    if (b_leader) {
      uint32_t bs = 0, bph = 0;
      bool first = true;
      ItemCursor cur;
      cur.init(gid, G, p.n_tiles);
      for (; cur.it < p.n_items; cur.next()) {
        if (p.b_resident && !first) break;   // resident weights: loaded once
        first = false;
        const int nt = cur.nt;
        for (int k = 0; k < p.n_steps; ++k) {
          mbar_wait(&b_empty[bs], bph ^ 1u);
          uint8_t* dst = sB + static_cast<size_t>(bs) * b_stage_bytes;
          const int row0 = (nt * p.n_steps + k) * p.n_tile + rank * p.w_rows;
          if (TWO) {
            if (rank == 0) mbar_expect_tx(&b_full[bs], 2u * p.step_bytes);
            tma_load_2d_pair(dst, &tmw, leader_cta_addr(&b_full[bs]), 0, row0);
          } else {
            mbar_expect_tx(&b_full[bs], p.step_bytes);
            tma_load_2d(dst, &tmw, &b_full[bs], 0, row0);
          }
          if (++bs == static_cast<uint32_t>(p.b_stages)) { bs = 0; bph ^= 1u; }
        }
      }
    }
  } else if ((warp == 1 || (warp == 3 && n_issuers == 2)) && rank == 0) {
    // ===================================================================== MMA issuers (rank-0 CTA of a pair only)
    // The whole warp walks the loop (so the address arithmetic stays warp-uniform); one elected lane
    // issues.  Descriptors are advanced by adding to their low word (start address >> 4).
    IssuerCtx c;
    c.a_full = a_full; c.a_empty = a_empty; c.b_full = b_full; c.b_empty = b_empty; c.acc_full = acc_full; c.acc_empty = acc_empty;
    c.k_first = k_first;
    c.step_a16 = s_step_a16; c.sub_step0 = p.sub_step0;
    c.sA_u32 = smem_u32(sA); c.sB_u32 = smem_u32(sB); c.a_stage16 = a_stage_bytes >> 4; c.b_stage16 = b_stage_bytes >> 4;
    c.tmem_base = tmem_base; c.acc_cols = acc_cols; c.n_tile = static_cast<uint32_t>(p.n_tile);
    c.n_sub = p.n_sub; c.a_stages = p.a_stages; c.b_stages = p.b_stages; c.acc_stages = p.acc_stages;
    c.Wp = p.Wp; c.super_stride = p.super_stride; c.n_tiles = p.n_tiles; c.wp_magic = p.wp_magic;
    c.it_begin = gid; c.it_end = p.n_items; c.it_stride = G;
    c.fused = p.fused; c.first_second_step = p.first_second_step;
    c.second_off = static_cast<uint32_t>(p.MT) * p.n_tile;
    const int issuer = (warp == 1) ? 0 : 1;
    const int my_mt = k_split ? 1 : p.MT / n_issuers;   // tiles this issuer owns: [issuer*my_mt, +my_mt)
    c.tile_off16 = k_split ? 0u : static_cast<uint32_t>(issuer * my_mt) * TILE16;
    c.d_off = k_split ? 0u : static_cast<uint32_t>(issuer * my_mt) * static_cast<uint32_t>(p.n_tile);
    c.leader = elect_one();
    c.dbg = p.dbg;
    long long tt[3] = {0, 0, 0};
    RingState rs;
    const long long t_begin = dbg_clock();
    if (k_split) {
      constexpr int KA = K32 >= 2 ? 1 : 0, KB = K32 >= 2 ? 2 : 0;   // (K32 == 1 never takes this branch)
      if (p.fused) {      // (the fused shortcut conv forces one tile per item, hence this branch)
        if (p.b_resident) { if (issuer == 0) run_issuer<ROWB, 1, true, TWO, KA, FP8, true>(c, tt, rs); else run_issuer<ROWB, 1, true, TWO, KB, FP8, true>(c, tt, rs); }
        else { if (issuer == 0) run_issuer<ROWB, 1, false, TWO, KA, FP8, true>(c, tt, rs); else run_issuer<ROWB, 1, false, TWO, KB, FP8, true>(c, tt, rs); }
      } else if (p.b_resident) { if (issuer == 0) run_issuer<ROWB, 1, true, TWO, KA, FP8>(c, tt, rs); else run_issuer<ROWB, 1, true, TWO, KB, FP8>(c, tt, rs); }
      else { if (issuer == 0) run_issuer<ROWB, 1, false, TWO, KA, FP8>(c, tt, rs); else run_issuer<ROWB, 1, false, TWO, KB, FP8>(c, tt, rs); }
    } else if (p.fused) {     // (one tile per item and no K split: the E4M3 network's fused shortcut, or the 32-byte stem rows)
      if (p.b_resident) run_issuer<ROWB, 1, true, TWO, 0, FP8, true>(c, tt, rs); else run_issuer<ROWB, 1, false, TWO, 0, FP8, true>(c, tt, rs);
    } else if (p.b_resident) {
      if (my_mt == 2) run_issuer<ROWB, 2, true, TWO, 0, FP8>(c, tt, rs); else run_issuer<ROWB, 1, true, TWO, 0, FP8>(c, tt, rs);
    } else {
      if (my_mt == 2) run_issuer<ROWB, 2, false, TWO, 0, FP8>(c, tt, rs); else run_issuer<ROWB, 1, false, TWO, 0, FP8>(c, tt, rs);
    }
    if (p.dbg_times && c.leader && issuer == 0) {
      long long* d = p.dbg_times + static_cast<size_t>(blockIdx.x) * 16;
      d[0] = dbg_clock() - t_begin; d[1] = tt[0]; d[2] = tt[1]; d[3] = tt[2];
      d[9] = dbg_clock() - t_entry;      // issuer done, since kernel entry
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue warps
    const long long t_begin = dbg_clock();
    long long t_wait = 0;
    // per-channel constants -> shared memory (kernel-lifetime constants, not produced by the previous kernel), while
    // the first patch / weight loads and MMAs are in flight; the epilogue warps meet on named barrier 1
    for (int i = static_cast<int>(threadIdx.x) - 128; i < p.OC; i += n_epi_warps * 32) {
      s_alpha[i] = p.alpha ? p.alpha[i] : 1.f;
      s_beta[i] = p.beta ? p.beta[i] : 0.f;
      if (p.fused) { s_alpha2[i] = p.alpha2[i]; s_beta2[i] = p.beta2[i]; }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(n_epi_warps * 32) : "memory");
    // grid-level mode: the residual is an earlier kernel's output, and our stores must not race its readers (flag mode:
    // see run_epilogue_items)
    if (p.n_deps == 0) pdl_wait();
    EpiShared sh;
    sh.s_alpha = s_alpha; sh.s_beta = s_beta; sh.s_alpha2 = s_alpha2; sh.s_beta2 = s_beta2; sh.s_stage = s_stage;
    sh.acc_full = acc_full; sh.acc_empty = acc_empty; sh.s_dep_seq = const_cast<uint32_t*>(s_dep_seq); sh.s_stored = s_stored;
    sh.tmem_base = tmem_base; sh.acc_cols = acc_cols; sh.acc_stages = p.acc_stages;
    EpiRing er;
    run_epilogue_items<TWO, FP8>(p, sh, er, warp, lane, n_epi_warps, gid, G, rank, t_wait);
    if (p.dbg_times && warp == 4 && lane == 0) {
      long long* d = p.dbg_times + static_cast<size_t>(blockIdx.x) * 16;
      d[4] = dbg_clock() - t_begin; d[5] = t_wait;
    }
  }
  if (p.dbg_times && threadIdx.x == 128) p.dbg_times[static_cast<size_t>(blockIdx.x) * 16 + 10] = dbg_clock() - t_entry;  // first epilogue warp done
  tc_fence_before();
  __syncthreads();
  if (p.dbg_times && threadIdx.x == 0) {
    p.dbg_times[static_cast<size_t>(blockIdx.x) * 16 + 11] = dbg_clock() - t_entry;    // CTA done
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg_times[static_cast<size_t>(blockIdx.x) * 16 + 13] = static_cast<long long>(gt);     // CTA done, ns
  }
  if (p.stamps && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    atomicMax(p.stamps + 1, gt);
  }
  if (TWO) cluster_sync_all();   // nobody leaves while the peer may still arrive on its barriers / read its smem
  if (warp == 1) {
    if (TWO) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace dlq
