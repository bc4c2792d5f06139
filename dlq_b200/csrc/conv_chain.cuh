// conv_chain.cuh — ONE persistent kernel running a sequence of convolution layers (ResNet-18: layer1's four convs as one
// chain of single CTAs, the fifteen convs of layer2 .. layer4 as one chain of CTA pairs).
//
// Why: with one launch per layer, every SM pays, per layer, the hand-over from one CTA to the next - the last item's
// epilogue drains with the tensor pipe idle, the CTA exits, the next one is scheduled, allocates TMEM, initialises
// barriers, loads its first patch - about 7 us of a 27 us layer at batch 256 (profiles/README.md), and the last, partial
// wave of items leaves most SMs idle on top of that.  Here a CTA pair never leaves: every role walks a LIST of layers,
// the shared-memory rings, TMEM accumulator stages and barrier phases simply continue from the last item of layer l to
// the first item of layer l+1, and the MMA issuers start layer l+1 while the epilogue warps still drain layer l.
// Cross-CTA dependencies (an item of layer l+1 reads rows that other CTAs produced in layer l) are the tile-level
// dependency flags of conv_kernel.cuh.  Items are dealt round-robin over the whole chain (layer l+1 continues with the
// CTA after the one that got layer l's last item), so the load is balanced to within one item over the chain.
//
// Because CTAs wait for other CTAs of the SAME grid, the grid must be co-resident: the host launches it as a cooperative
// kernel with one CTA pair per SM pair.
//
// Replaces the per-block launch sequence of basic_block_forward (reference runtime/infer_e2e.cu:156-203) for the blocks
// it covers; the arithmetic of every layer is conv_i8_kernel's (same issuer, same epilogue code).
//
// Static configuration shared by all layers of a chain (checked by the planner): one of two tile shapes - CTA pairs
// (cta_group::2) with two 128-position tiles per item and 128-channel n-tiles (layer2..4), or single CTAs with four tiles
// and 64-channel n-tiles (layer1) - i.e. two TMEM accumulator stages of 256 columns either way; streamed weights, no fused
// shortcut, no raw-accumulator output.  The K-row width (64 or 128 bytes) may differ from layer to layer.
#pragma once
#include "conv_kernel.cuh"

namespace dlq {

constexpr int kMaxChainLayers = 15;

struct ChainLayer {
  ConvKernelParams p;
  int item_shift;            // CTA (pair) g takes the items  it = (g - item_shift) mod G, + G, ...  of this layer
  int rowb;                  // bytes of K per A / B row of this layer: 64 or 128
  CUtensorMap tm0, tmw;      // activations (3-D), packed weight image (2-D); 64-byte aligned by their type
};

struct ChainParams {
  int n_layers;
  int a_stages, b_stages;              // ring depths (the same rings serve every layer)
  int a_stage_bytes, b_stage_bytes;    // bytes per stage: the maximum over the layers, 1024-aligned
  int oc_max;
  ChainLayer layer[kMaxChainLayers];
};
static_assert(sizeof(ChainParams) <= 32000, "kernel parameters are limited to 32764 bytes");

// smem layout (dynamic, 1024-aligned base):
//   [A ring][B ring][alpha, beta: 2 * oc_max f32][epilogue staging: 16 * kEpiStageBytes][step offsets: 2 issuers]
//   [barriers][tmem slot, dependency words]
template <bool TWO, int MT, int N_TILE, bool FP8>
__global__ void __launch_bounds__(384, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  constexpr int ACC_STAGES = 2;
  constexpr uint32_t ACC_COLS = MT * N_TILE;
  static_assert(ACC_COLS * ACC_STAGES == 512 && MT % 2 == 0, "two accumulator stages fill the 512 TMEM columns; one half of the tiles per issuer");
  constexpr int ncta = TWO ? 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_stage_bytes = static_cast<uint32_t>(cp.a_stage_bytes);
  const uint32_t b_stage_bytes = static_cast<uint32_t>(cp.b_stage_bytes);
  uint8_t* sA = smem;
  uint8_t* sB = sA + static_cast<size_t>(cp.a_stages) * a_stage_bytes;
  float* s_alpha = reinterpret_cast<float*>(sB + static_cast<size_t>(cp.b_stages) * b_stage_bytes);
  float* s_beta = s_alpha + cp.oc_max;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_beta + cp.oc_max);
  uint16_t* s_step_a16 = reinterpret_cast<uint16_t*>(s_stage + 16 * kEpiStageBytes);        // [2 issuers][kMaxSteps + 8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_step_a16 + 2 * (kMaxSteps + 8));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + cp.a_stages;
  uint64_t* b_full = a_empty + cp.a_stages;
  uint64_t* b_empty = b_full + cp.b_stages;
  uint64_t* acc_full = b_empty + cp.b_stages;
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
  uint32_t* s_dep_seq = tmem_slot + 1;
  uint32_t* s_stored = tmem_slot + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int n_epi_warps = 8;
  constexpr int n_issuers = 2;
  const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
  const int gid = static_cast<int>(blockIdx.x) / ncta;
  const int G = static_cast<int>(gridDim.x) / ncta;
  const bool first_grid_dep = cp.layer[0].p.n_deps == 0;     // the first layer waits for the previous kernel's grid

  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    for (int i = 0; i < cp.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], n_issuers); }
    for (int i = 0; i < cp.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_issuers); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], n_issuers); mbar_init(&acc_empty[i], n_epi_warps * ncta); }
    fence_mbar_init();
    *s_dep_seq = 0u;
    for (int i = 0; i < kStoredSlots; ++i) s_stored[i] = 0u;
  }
  if (warp == 1) {
    if (TWO) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto first_item = [&](const ChainLayer& L) { int f = (gid - L.item_shift) % G; return f < 0 ? f + G : f; };

  if (warp == 0) {
    // ===================================================================== A (activation patch) producer
    const bool leader = elect_one();
    uint32_t as = 0, aph = 0, seq = 0;
    if (first_grid_dep && leader) pdl_wait();
    for (int l = 0; l < cp.n_layers; ++l) {
      const ChainLayer& L = cp.layer[l];
      const ConvKernelParams& p = L.p;
      if (p.stamps && leader && rank == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMin(p.stamps, gt);
      }
      ItemCursor cur;
      cur.init(first_item(L), G, p.n_tiles);
      for (; cur.it < p.n_items; cur.next()) {
        const int st = TWO ? 2 * cur.sp + rank : cur.sp;
        if (p.n_deps) {
          const int g0 = st * p.super_stride;
          wait_deps(p, g0, min(g0 + p.super_stride, p.total_pos) - 1, lane);
        }
        if (leader) {
          if (p.n_deps) fence_proxy_async_all();
          ++seq;
          asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(s_dep_seq)), "r"(seq) : "memory");
          const int v0 = div_magic(st * p.super_stride, p.Wp, p.wp_magic);
          for (int s = 0; s < p.n_sub; ++s) {
            mbar_wait(&a_empty[as], aph ^ 1u);
            uint8_t* dst = sA + static_cast<size_t>(as) * a_stage_bytes;
            if (TWO) {
              if (rank == 0) mbar_expect_tx(&a_full[as], 2u * static_cast<uint32_t>(p.tma_bytes));
              tma_load_3d_pair(dst, &L.tm0, leader_cta_addr(&a_full[as]), p.sub_c0[s], p.sub_col0[s],
                               p.row_mul * v0 + p.sub_row_off[s] + p.sub_plane_row[s]);
            } else {
              mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.tma_bytes));
              tma_load_3d(dst, &L.tm0, &a_full[as], p.sub_c0[s], p.sub_col0[s],
                          p.row_mul * v0 + p.sub_row_off[s] + p.sub_plane_row[s]);
            }
            if (++as == static_cast<uint32_t>(cp.a_stages)) { as = 0; aph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== B (weight step) producer + signaller
    const bool b_leader = elect_one();
    const int sig_lane = (__ballot_sync(0xffffffffu, b_leader) & 0x80000000u) ? 30 : 31;
    if (!b_leader && lane == sig_lane) {
      uint32_t k = 0;       // (see "Producer side" in conv_kernel.cuh)
      for (int l = 0; l < cp.n_layers; ++l) {
        const ChainLayer& L = cp.layer[l];
        const ConvKernelParams& p = L.p;
        ItemCursor cur;
        cur.init(first_item(L), G, p.n_tiles);
        for (; cur.it < p.n_items; cur.next(), ++k) {
          const uint32_t want = static_cast<uint32_t>(n_epi_warps) * ((k / kStoredSlots) + 1u);
          const uint32_t addr = smem_u32(s_stored + (k % kStoredSlots));
          uint32_t v;
          do {
            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
          } while (v < want);
          if (p.done) {
            __threadfence();
            atomicAdd(p.done + cur.sp, static_cast<unsigned int>(n_epi_warps));
          }
        }
      }
    }
    if (b_leader) {
      uint32_t bs = 0, bph = 0;
      for (int l = 0; l < cp.n_layers; ++l) {
        const ChainLayer& L = cp.layer[l];
        const ConvKernelParams& p = L.p;
        ItemCursor cur;
        cur.init(first_item(L), G, p.n_tiles);
        for (; cur.it < p.n_items; cur.next()) {
          const int nt = cur.nt;
          for (int k = 0; k < p.n_steps; ++k) {
            mbar_wait(&b_empty[bs], bph ^ 1u);
            uint8_t* dst = sB + static_cast<size_t>(bs) * b_stage_bytes;
            const int row0 = (nt * p.n_steps + k) * p.n_tile + rank * p.w_rows;
            if (TWO) {
              if (rank == 0) mbar_expect_tx(&b_full[bs], 2u * p.step_bytes);
              tma_load_2d_pair(dst, &L.tmw, leader_cta_addr(&b_full[bs]), 0, row0);
            } else {
              mbar_expect_tx(&b_full[bs], p.step_bytes);
              tma_load_2d(dst, &L.tmw, &b_full[bs], 0, row0);
            }
            if (++bs == static_cast<uint32_t>(cp.b_stages)) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if ((warp == 1 || warp == 3) && rank == 0) {
    // ===================================================================== MMA issuers (rank-0 CTA): half of the tiles each
    const int issuer = (warp == 1) ? 0 : 1;
    constexpr int MY_MT = MT / 2;
    uint16_t* my_steps = s_step_a16 + issuer * (kMaxSteps + 8);
    IssuerCtx c;
    c.a_full = a_full; c.a_empty = a_empty; c.b_full = b_full; c.b_empty = b_empty; c.acc_full = acc_full; c.acc_empty = acc_empty;
    c.k_first = nullptr;
    c.step_a16 = my_steps;
    c.sA_u32 = smem_u32(sA); c.sB_u32 = smem_u32(sB); c.a_stage16 = a_stage_bytes >> 4; c.b_stage16 = b_stage_bytes >> 4;
    c.tmem_base = tmem_base; c.acc_cols = ACC_COLS; c.n_tile = N_TILE;
    c.a_stages = cp.a_stages; c.b_stages = cp.b_stages; c.acc_stages = ACC_STAGES;
    c.fused = 0; c.first_second_step = -1; c.second_off = 0;
    c.d_off = static_cast<uint32_t>(issuer * MY_MT) * N_TILE;
    c.leader = elect_one();
    c.dbg = 0;
    RingState rs;
    long long tt[3];
    for (int l = 0; l < cp.n_layers; ++l) {
      const ChainLayer& L = cp.layer[l];
      const ConvKernelParams& p = L.p;
      __syncwarp();       // the previous layer's reads of the step table are over
      for (int i = lane; i < p.n_steps; i += 32) my_steps[i] = p.step_a16[i];
      __syncwarp();
      c.sub_step0 = p.sub_step0;
      c.n_sub = p.n_sub;
      c.Wp = p.Wp; c.super_stride = p.super_stride; c.n_tiles = p.n_tiles; c.wp_magic = p.wp_magic;
      c.it_begin = first_item(L); c.it_end = p.n_items; c.it_stride = G;
      c.tile_off16 = static_cast<uint32_t>(issuer * MY_MT) * static_cast<uint32_t>(kTileM * L.rowb / 16);
      if (L.rowb == 128) run_issuer<128, MY_MT, false, TWO, 0, FP8>(c, tt, rs);
      else run_issuer<64, MY_MT, false, TWO, 0, FP8>(c, tt, rs);
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue warps
    if (first_grid_dep) pdl_wait();
    EpiShared sh;
    sh.s_alpha = s_alpha; sh.s_beta = s_beta; sh.s_alpha2 = s_alpha; sh.s_beta2 = s_beta; sh.s_stage = s_stage;
    sh.acc_full = acc_full; sh.acc_empty = acc_empty; sh.s_dep_seq = s_dep_seq; sh.s_stored = s_stored;
    sh.tmem_base = tmem_base; sh.acc_cols = ACC_COLS; sh.acc_stages = ACC_STAGES;
    EpiRing er;
    long long t_wait = 0;
    for (int l = 0; l < cp.n_layers; ++l) {
      const ChainLayer& L = cp.layer[l];
      const ConvKernelParams& p = L.p;
      // the eight warps switch layers together: nobody overwrites alpha / beta while another warp still reads them
      if (l > 0) asm volatile("bar.sync 1, %0;" ::"r"(n_epi_warps * 32) : "memory");
      for (int i = static_cast<int>(threadIdx.x) - 128; i < p.OC; i += n_epi_warps * 32) {
        s_alpha[i] = p.alpha[i];
        s_beta[i] = p.beta[i];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(n_epi_warps * 32) : "memory");
      run_epilogue_items<TWO, FP8>(p, sh, er, warp, lane, n_epi_warps, first_item(L), G, rank, t_wait);
      if (p.stamps && warp == 4 && lane == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        atomicMax(p.stamps + 1, gt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();
  if (warp == 1) {
    if (TWO) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dlq
