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
//   The 3-channel stem uses a 2x2 space-to-depth input (16 B per pixel, no swizzle) and pairs two
//   horizontally adjacent pixels into one K=32 MMA through the descriptor's leading-byte-offset.
//   (All three addressing tricks are validated by probe/umma_probe.cu on a B200.)
//
// Warp roles (256 or 384 threads): warp 0 = activation-patch TMA producer, warp 1 = MMA issuer
//   (+TMEM alloc), warp 2 = weight-step bulk-copy producer, warp 3 = spare, warps 4.. = epilogue
//   (4 or 8 warps): tcgen05.ld -> alpha/beta/residual/ReLU/requant -> 16-byte stores.
// Pipelines: A patch ring (a_full/a_empty), weight-step ring (b_full/b_empty), TMEM accumulator
//   stages (acc_full/acc_empty).
#pragma once
#include <cuda.h>
#include "sm100_ptx.cuh"

namespace dlq {

constexpr int kMaxSteps = 40;    // K steps (tap x channel-block) per conv
constexpr int kMaxPlanes = 4;
constexpr int kTileM = 128;

struct ConvStep {
  uint32_t a_off;    // byte offset of this step's A view inside its sub-patch (= tap shift * ROWB)
};

struct ConvKernelParams {
  // virtual output space
  int Wp, Wo, Ho, Pv, N;
  int num_super;          // number of super-tiles (each MT*128 positions)
  int MT;                 // M tiles per super-tile
  int n_tile;             // UMMA N (output channels per CTA column)
  int OC;                 // total output channels
  // A patch: n_sub sub-patches (parity plane x channel block); each is ONE 3-D TMA load and one
  // stage of the A ring.  K steps are grouped by sub-patch: sub s owns steps [sub_step0[s], sub_step0[s+1]).
  int n_sub;
  int sub_bytes;          // bytes per sub-patch stage in smem (1024-aligned)
  int tma_bytes;          // bytes written per sub-patch by TMA (NR*Wp*ROWB)
  int row_mul;            // TMA row coordinate = row_mul * v0 + sub_row_off
  int16_t sub_c0[16];     // TMA coordinates per sub-patch: channel byte offset
  int16_t sub_col0[16];   //   start column (may be negative)
  int16_t sub_row_off[16];//   row offset
  int16_t sub_step0[17];  //   first K step of each sub-patch (sub_step0[n_sub] = n_steps)
  // K steps
  int n_steps;
  int k32_per_step;       // MMAs (K=32) per step per tile
  ConvStep steps[kMaxSteps];
  int a_stages, b_stages, acc_stages;
  uint32_t step_bytes;    // weight image bytes per step (n_tile * ROWB, or n_tile*32 for the stem)
  const uint8_t* wimg;    // [n_tiles][n_steps][step_bytes] pre-swizzled smem images
  // epilogue
  const float* alpha;     // [OC]
  const float* beta;      // [OC]
  const int8_t* residual; // row-padded NHWC int8 [.,Ho,Wo,OC] or nullptr
  int res_PR;
  float res_scale;
  int relu;
  float inv_out_scale;
  int8_t* out;            // row-padded NHWC int8
  int out_PR;
  int32_t* acc_out;       // optional dense NHWC int32 [N,Ho,Wo,OC] raw accumulators (debug / parity)
};

// smem layout (dynamic, 1024-aligned base):
//   [A ring: a_stages * sub_bytes][B ring: b_stages * step_bytes(1024-aligned)][alpha,beta: 2*n_tile f32]
//   [barriers][tmem slot]
template <int ROWB>
__global__ void __launch_bounds__(384, 1)
conv_i8_kernel(const __grid_constant__ CUtensorMap tm0, const ConvKernelParams p) {
  constexpr uint32_t LAYOUT = ROWB == 128 ? UMMA_SWZ_128B : ROWB == 64 ? UMMA_SWZ_64B : UMMA_SWZ_NONE;
  // descriptor strides: swizzled K-major: SBO = 8 rows; no-swizzle 16B pixels: SBO = 128 B, LBO = 16 B (next pixel)
  constexpr uint32_t A_SBO = ROWB == 16 ? 128u : 8u * ROWB;
  constexpr uint32_t A_LBO = ROWB == 16 ? 16u : 0u;
  constexpr uint32_t B_SBO = ROWB == 16 ? 128u : 8u * ROWB;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_stage_bytes = static_cast<uint32_t>(p.sub_bytes);
  const uint32_t b_stage_bytes = (p.step_bytes + 1023u) & ~1023u;
  uint8_t* sA = smem;
  uint8_t* sB = sA + static_cast<size_t>(p.a_stages) * a_stage_bytes;
  float* s_alpha = reinterpret_cast<float*>(sB + static_cast<size_t>(p.b_stages) * b_stage_bytes);
  float* s_beta = s_alpha + p.n_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + p.n_tile);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint64_t* acc_empty = acc_full + p.acc_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + p.acc_stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_epi_warps = (blockDim.x >> 5) - 4;
  const int n_blk = blockIdx.y;                 // output-channel tile
  const int n0 = n_blk * p.n_tile;
  const uint32_t acc_cols = static_cast<uint32_t>(p.MT) * p.n_tile;   // TMEM columns per accumulator stage
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_cols * p.acc_stages) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < p.acc_stages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], n_epi_warps); }
    fence_mbar_init();
    tma_prefetch_desc(&tm0);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
    s_alpha[i] = p.alpha ? p.alpha[n0 + i] : 1.f;
    s_beta[i] = p.beta ? p.beta[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int super_pos = p.MT * kTileM;

  if (warp == 0) {
    // ===================================================================== A (activation patch) producer
    if (elect_one()) {
      uint32_t a_it = 0;
      for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
        const int g0 = st * super_pos;
        const int v0 = g0 / p.Wp;
        for (int s = 0; s < p.n_sub; ++s) {
          const uint32_t as = a_it % p.a_stages, aph = (a_it / p.a_stages) & 1u;
          mbar_wait(&a_empty[as], aph ^ 1u);
          mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.tma_bytes));
          tma_load_3d(sA + static_cast<size_t>(as) * a_stage_bytes, &tm0, &a_full[as], p.sub_c0[s], p.sub_col0[s],
                      p.row_mul * v0 + p.sub_row_off[s]);
          ++a_it;
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== B (weight step) producer
    if (elect_one()) {
      uint32_t b_it = 0;
      const uint8_t* wsrc = p.wimg + static_cast<size_t>(n_blk) * p.n_steps * p.step_bytes;
      for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
        for (int k = 0; k < p.n_steps; ++k) {
          const uint32_t bs = b_it % p.b_stages, bph = (b_it / p.b_stages) & 1u;
          mbar_wait(&b_empty[bs], bph ^ 1u);
          mbar_expect_tx(&b_full[bs], p.step_bytes);
          bulk_g2s(sB + static_cast<size_t>(bs) * b_stage_bytes, wsrc + static_cast<size_t>(k) * p.step_bytes,
                   p.step_bytes, &b_full[bs]);
          ++b_it;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_s8(kTileM, static_cast<uint32_t>(p.n_tile));
      uint32_t a_it = 0, b_it = 0, c_it = 0;
      for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
        const int g0 = st * super_pos;
        const int v0 = g0 / p.Wp;
        const uint32_t in_patch = static_cast<uint32_t>(g0 - v0 * p.Wp);    // first position's offset in the patch
        const uint32_t cs = c_it % p.acc_stages, cph = (c_it / p.acc_stages) & 1u;
        mbar_wait(&acc_empty[cs], cph ^ 1u);
        const uint32_t d_base = tmem_base + cs * acc_cols;
        for (int s = 0; s < p.n_sub; ++s) {
          const uint32_t as = a_it % p.a_stages, aph = (a_it / p.a_stages) & 1u;
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + static_cast<size_t>(as) * a_stage_bytes) + in_patch * ROWB;
          for (int k = p.sub_step0[s]; k < p.sub_step0[s + 1]; ++k) {
            const uint32_t bs = b_it % p.b_stages, bph = (b_it / p.b_stages) & 1u;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(sB + static_cast<size_t>(bs) * b_stage_bytes);
            const uint32_t a_step = a_base + p.steps[k].a_off;
            for (int mt = 0; mt < p.MT; ++mt) {
              const uint32_t a_tile = a_step + static_cast<uint32_t>(mt) * kTileM * ROWB;
              for (int kk = 0; kk < p.k32_per_step; ++kk) {
                const uint64_t ad = umma_smem_desc(a_tile + kk * 32u, A_LBO, A_SBO, LAYOUT);
                const uint64_t bd =
                    ROWB == 16 ? umma_smem_desc(b_addr, static_cast<uint32_t>(p.n_tile) * 16u, B_SBO, LAYOUT)
                               : umma_smem_desc(b_addr + kk * 32u, 0u, B_SBO, LAYOUT);
                umma_i8(d_base + static_cast<uint32_t>(mt) * p.n_tile, ad, bd, idesc, (k | kk) ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[bs]);   // weight stage free once these MMAs retire
            ++b_it;
          }
          umma_commit(&a_empty[as]);     // sub-patch stage free
          ++a_it;
        }
        umma_commit(&acc_full[cs]);      // accumulators ready for the epilogue
        ++c_it;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue warps
    const int ew = warp - 4;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int n_groups = n_epi_warps >> 2;              // 1 or 2 column groups
    const int grp = (ew >> 2);                          // which column group this warp handles
    const int cols_per_grp = p.n_tile / n_groups;
    const int col_lo = grp * cols_per_grp;
    const int row = quarter * 32 + lane;                // accumulator row within the tile
    const bool has_res = p.residual != nullptr;
    const int out_pitch = p.Ho + p.out_PR, res_pitch = p.Ho + p.res_PR;
    const int lo = p.relu ? 0 : -128;
    uint32_t c_it = 0;
    for (int st = blockIdx.x; st < p.num_super; st += gridDim.x) {
      const uint32_t cs = c_it % p.acc_stages, cph = (c_it / p.acc_stages) & 1u;
      mbar_wait(&acc_full[cs], cph);
      tc_fence_after();
      for (int mt = 0; mt < p.MT; ++mt) {
        const int g = st * super_pos + mt * kTileM + row;
        const int vrow = g / p.Wp, x = g - vrow * p.Wp;
        const int n = vrow / p.Pv, r = vrow - n * p.Pv;
        const bool valid = (x < p.Wo) && (r < p.Ho) && (n < p.N);
        const size_t opix = (static_cast<size_t>(p.out_PR + n * out_pitch + r) * p.Wo + x);
        const size_t rpix = (static_cast<size_t>(p.res_PR + n * res_pitch + r) * p.Wo + x);
        const size_t dpix = (static_cast<size_t>(n) * p.Ho + r) * p.Wo + x;
        const uint32_t taddr = tmem_base + cs * acc_cols + static_cast<uint32_t>(mt) * p.n_tile +
                               (static_cast<uint32_t>(quarter * 32) << 16);
        for (int c = col_lo; c < col_lo + cols_per_grp; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c, v);
          tmem_ld_wait();
          if (valid) {
            if (p.acc_out) {
              int4* dst = reinterpret_cast<int4*>(p.acc_out + dpix * p.OC + n0 + c);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
            }
            if (p.out) {
              int4 rv = make_int4(0, 0, 0, 0);
              if (has_res) rv = __ldg(reinterpret_cast<const int4*>(p.residual + rpix * p.OC + n0 + c));
              const int8_t* rb = reinterpret_cast<const int8_t*>(&rv);
              uint32_t packed[4];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float t = __fmaf_rn(static_cast<float>(static_cast<int32_t>(v[j])), s_alpha[c + j], s_beta[c + j]);
                if (has_res) t = __fmaf_rn(static_cast<float>(rb[j]), p.res_scale, t);
                if (p.relu && t < 0.f) t = 0.f;
                t = __fmul_rn(t, p.inv_out_scale);
                int q = __float2int_rn(t);
                q = max(lo, min(127, q));
                if ((j & 3) == 0) packed[j >> 2] = 0;
                packed[j >> 2] |= (static_cast<uint32_t>(q) & 0xFFu) << (8 * (j & 3));
              }
              *reinterpret_cast<int4*>(p.out + opix * p.OC + n0 + c) =
                  make_int4((int)packed[0], (int)packed[1], (int)packed[2], (int)packed[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[cs]);
      ++c_it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace dlq
