// conv_plan.cu — host side of the INT8 convolution: weight packing into shared-memory images,
// launch planning (virtual position space, sub-patches, K-step schedule, TMA descriptor) and launch.
//
// Replaces the per-call host work of conv2d_nchw_im2col_gemm (reference runtime/infer_e2e.cu:102-136:
// OIHW -> [OC, IC*kH*kW] repack, 3x cudaMalloc, weight upload, two launches) with a pack-once /
// plan-once / launch-many split.
#include "dlq_internal.h"
#include <algorithm>
#include <cstdlib>

namespace dlq {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct SubDesc {
  int c0;           // channel byte offset
  int col0;         // TMA start column
  int row_off_rel;  // TMA row = row_mul * v0 + PR_in + row_off_rel
};
struct StepDesc {
  int sub;
  int kh, kw;  // filter tap (stem: kh = row block a, kw = column pair bp)
  int kb;      // channel block
  int da, db;  // tap shift in patch rows / columns
  int second = 0;  // 1: step of the fused 1x1 shortcut conv (weights of w->second_q, A view of the centre tap)
};

void make_schedule(const dlq_conv_weights* w, std::vector<SubDesc>& subs, std::vector<StepDesc>& steps) {
  subs.clear();
  steps.clear();
  const int rowb = w->rowb;
  switch (w->kind) {
    case CONV_S1:
      for (int kb = 0; kb < w->kb; ++kb) {
        subs.push_back({kb * rowb, -w->pW, -w->pH});
        for (int kh = 0; kh < w->kH; ++kh)
          for (int kw = 0; kw < w->kW; ++kw) steps.push_back({kb, kh, kw, kb, kh, kw});
      }
      break;
    case CONV_S2_3x3: {
      // parity planes of the input (row parity, col parity) and the taps that read them.
      // input row 2*oh - 1 + kh: kh=0 -> odd plane index oh-1 (da 0), kh=1 -> even plane (da 0),
      // kh=2 -> odd plane index oh (da 1).  Same for columns.
      struct Tap { int kh, kw, da, db; };
      struct Plane { int row_off_rel, col0; std::vector<Tap> taps; };
      const Plane planes[4] = {
          {-1, -1, {{0, 0, 0, 0}, {0, 2, 0, 1}, {2, 0, 1, 0}, {2, 2, 1, 1}}},
          {-1, 0, {{0, 1, 0, 0}, {2, 1, 1, 0}}},
          {0, -1, {{1, 0, 0, 0}, {1, 2, 0, 1}}},
          {0, 0, {{1, 1, 0, 0}}},
      };
      for (const Plane& pl : planes)
        for (int kb = 0; kb < w->kb; ++kb) {
          const int s = static_cast<int>(subs.size());
          subs.push_back({kb * rowb, pl.col0, pl.row_off_rel});
          for (const Tap& t : pl.taps) steps.push_back({s, t.kh, t.kw, kb, t.da, t.db});
          // fused 1x1/s2 shortcut: input (2 oh, 2 ow) is the (even, even) plane at shift 0 - the centre tap's view
          if (w->fused && pl.row_off_rel == 0 && pl.col0 == 0) steps.push_back({s, 0, 0, kb, 0, 0, 1});
        }
      break;
    }
    case CONV_S2_1x1:
      for (int kb = 0; kb < w->kb; ++kb) {
        subs.push_back({kb * rowb, 0, 0});
        steps.push_back({kb, 0, 0, kb, 0, 0});
      }
      break;
    case CONV_STEM:
      // 7x7/s2/p3 over 3 channels == 4x4/s1 over the 2x2 space-to-depth image (16 B per pixel):
      // input row 2*oh - 3 + kh = 2*(oh - 2) + (kh + 1)  ->  s2d row oh - 2 + a, a = (kh+1)/2, dy = (kh+1)%2.
      // the input tensor carries its horizontal padding physically (2 left, 1 right), so the column origin is 0
      subs.push_back({0, 0, -2});
      for (int a = 0; a < 4; ++a)
        for (int bp = 0; bp < 2; ++bp) steps.push_back({0, a, bp, 0, a, 2 * bp});
      break;
  }
}

inline uint32_t swz_off(uint32_t row, uint32_t kbyte, uint32_t row_bytes) {
  const uint32_t lin = row * row_bytes + kbyte;
  if (row_bytes == 128) return lin ^ (((lin >> 7) & 7u) << 4);
  if (row_bytes == 64) return lin ^ (((lin >> 7) & 3u) << 4);
  if (row_bytes == 32) return lin ^ (((lin >> 7) & 1u) << 4);
  return lin;
}

}  // namespace

void conv_out_dims(const dlq_conv_weights* w, int H, int W, int* OH, int* OW) {
  *OH = (H + 2 * w->pH - w->kH) / w->sH + 1;
  *OW = (W + 2 * w->pW - w->kW) / w->sW + 1;
}

int conv_required_in_pr(const dlq_conv_weights* w) {
  switch (w->kind) {
    case CONV_S1: return w->pH;
    case CONV_S2_3x3: return 2;   // >= 1 and H + PR even (H even)
    case CONV_S2_1x1: return 2;
    case CONV_STEM: return 2;
  }
  return 0;
}

int pack_conv_weights(dlq_ctx* ctx, const int8_t* wq, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                      int pW, dlq_conv_weights* out, const int8_t* second_wq) {
  DLQ_ARG(ctx, OC > 0 && OC % 64 == 0, "OC must be a multiple of 64");
  DLQ_ARG(ctx, kH == kW && sH == sW && pH == pW, "square kernels / strides / pads only");
  out->OC = OC; out->IC = IC; out->kH = kH; out->kW = kW; out->sH = sH; out->sW = sW; out->pH = pH; out->pW = pW;
  out->device = ctx->device;
  if (IC == 3 && kH == 7 && sH == 2 && pH == 3) {
    out->kind = CONV_STEM; out->rowb = 32; out->kb = 1;
  } else {
    DLQ_ARG(ctx, IC == 64 || (IC > 0 && IC % 128 == 0), "IC must be 64 or a multiple of 128 (or the 3-channel 7x7/s2/p3 stem)");
    out->rowb = IC == 64 ? 64 : 128;
    out->kb = IC == 64 ? 1 : IC / 128;
    if (sH == 1) {
      DLQ_ARG(ctx, kH >= 1 && kH <= 7 && (kH & 1) && pH <= kH / 2, "stride-1 conv: odd k <= 7, pad <= k/2");
      out->kind = CONV_S1;
    } else if (sH == 2 && kH == 3 && pH == 1) {
      out->kind = CONV_S2_3x3;
    } else if (sH == 2 && kH == 1 && pH == 0) {
      out->kind = CONV_S2_1x1;
    } else {
      DLQ_ARG(ctx, false, "unsupported (kernel, stride, pad) combination");
    }
  }
  // 128-channel tiles run as CTA pairs (cta_group::2, M = 256): per SM an MMA then reads 4 KB of A + 2 KB of B per
  // 64 tensor cycles, inside the 128 B/clk shared-memory budget; 64-channel layers stay single-CTA
  out->n_tile = (OC % 128 == 0) ? 128 : 64;
  if (const char* e = dlq_dbg_env("DLQ_DBG_NTILE")) { const int v = atoi(e); if (v >= 64 && OC % v == 0) out->n_tile = v; }
  out->fused = 0;
  if (second_wq) {
    DLQ_ARG(ctx, out->kind == CONV_S2_3x3, "a fused 1x1/s2 shortcut needs a 3x3/s2/p1 main conv");
    out->fused = 1;
    out->second_q.assign(second_wq, second_wq + static_cast<size_t>(OC) * IC);
  }
  std::vector<SubDesc> subs;
  std::vector<StepDesc> steps;
  make_schedule(out, subs, steps);
  DLQ_ARG(ctx, static_cast<int>(steps.size()) <= kMaxSteps && subs.size() <= 16, "too many K steps for one conv");
  out->n_steps = static_cast<int>(steps.size());
  out->step_bytes = static_cast<uint32_t>(out->n_tile) * out->rowb;
  out->q_oihw.assign(wq, wq + static_cast<size_t>(OC) * IC * kH * kW);

  const int n_tiles = OC / out->n_tile;
  std::vector<uint8_t> img(static_cast<size_t>(n_tiles) * out->n_steps * out->step_bytes, 0);
  auto W = [&](int o, int c, int kh, int kw) -> int8_t {
    return wq[((static_cast<size_t>(o) * IC + c) * kH + kh) * kW + kw];
  };
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int s = 0; s < out->n_steps; ++s) {
      uint8_t* dst = img.data() + (static_cast<size_t>(nt) * out->n_steps + s) * out->step_bytes;
      const StepDesc& sd = steps[s];
      for (int n = 0; n < out->n_tile; ++n) {
        const int o = nt * out->n_tile + n;
        if (out->kind == CONV_STEM) {
          // 32-byte row = pixel pair (b = 2*bp + j, j = 0,1); 16 B per pixel = [dy][dx][c4]
          for (int j = 0; j < 2; ++j)
            for (int t = 0; t < 16; ++t) {
              const int dy = t >> 3, dx = (t >> 2) & 1, c = t & 3;
              const int kh = 2 * sd.kh + dy - 1, kw = 2 * (2 * sd.kw + j) + dx - 1;
              int8_t v = 0;
              if (c < 3 && kh >= 0 && kh < 7 && kw >= 0 && kw < 7) v = W(o, c, kh, kw);
              dst[swz_off(n, j * 16 + t, 32)] = static_cast<uint8_t>(v);
            }
        } else if (sd.second) {
          for (int k = 0; k < out->rowb; ++k)      // shortcut weights: OC x IC (1x1)
            dst[swz_off(n, k, out->rowb)] = static_cast<uint8_t>(out->second_q[static_cast<size_t>(o) * IC + sd.kb * out->rowb + k]);
        } else {
          for (int k = 0; k < out->rowb; ++k)
            dst[swz_off(n, k, out->rowb)] = static_cast<uint8_t>(W(o, sd.kb * out->rowb + k, sd.kh, sd.kw));
        }
      }
    }
  DLQ_CUDA(ctx, cudaMalloc(&out->d_img, img.size()));
  DLQ_CUDA(ctx, cudaMemcpyAsync(out->d_img, img.data(), img.size(), cudaMemcpyHostToDevice, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DLQ_OK;
}

int plan_conv(dlq_ctx* ctx, const dlq_conv_weights* w, const Act& in, const Act& out, const float* alpha,
              const float* beta, const Act* residual, float res_mul, int relu, int32_t* acc_out, ConvLaunch* L,
              const SecondConv* second, bool allow_resident) {
  ConvKernelParams& p = L->p;
  memset(&p, 0, sizeof(p));
  L->fp8 = w->fp8;
  const int rowb = w->rowb;
  L->rowb = rowb;
  int es = 1;
  switch (w->kind) {
    case CONV_S1:
      DLQ_ARG(ctx, in.C == w->IC, "input channel mismatch");
      DLQ_ARG(ctx, in.PR >= w->pH, "input tensor needs PR >= pad");
      p.Ho = in.H + 2 * w->pH - w->kH + 1;
      p.Wo = in.W + 2 * w->pW - w->kW + 1;
      // virtual row pitch: W + pad, not W + 2 pad - the zero column left of row r+1 doubles as the zero column
      // right of row r (the patch is one linear run of pixels), exactly as the pad rows are shared between images
      p.Wp = in.W + (dlq_dbg_env("DLQ_DBG_WIDE_ROWS") ? 2 * w->pW : w->pW);
      if (p.Wp < p.Wo) p.Wp = p.Wo;
      p.Pv = in.H + in.PR;
      break;
    case CONV_S2_3x3:
    case CONV_S2_1x1:
      DLQ_ARG(ctx, in.C == w->IC, "input channel mismatch");
      DLQ_ARG(ctx, in.H % 2 == 0 && in.W % 2 == 0, "stride-2 conv needs even H, W");
      DLQ_ARG(ctx, in.PR >= 1 && (in.H + in.PR) % 2 == 0, "stride-2 conv needs PR >= 1 and even image pitch");
      es = 2;
      p.Ho = in.H / 2;
      p.Wo = in.W / 2;
      p.Wp = w->kind == CONV_S2_3x3 ? p.Wo + 1 : p.Wo;
      p.Pv = (in.H + in.PR) / 2;
      break;
    case CONV_STEM:
      DLQ_ARG(ctx, in.C == 32 && in.PR >= 2 && in.W > 3,
              "stem expects the paired 2x2 space-to-depth input (32 B/pixel, W/2+3 columns, PR >= 2)");
      p.Ho = in.H + 3 - 4 + 1;
      p.Wo = in.W - 3;
      p.Wp = in.W;
      p.Pv = in.H + in.PR;
      break;
  }
  DLQ_ARG(ctx, out.N == in.N && out.H == p.Ho && out.W == p.Wo && out.C == w->OC, "output tensor geometry mismatch");
  if (residual)
    DLQ_ARG(ctx, residual->N == in.N && residual->H == p.Ho && residual->W == p.Wo && residual->C == w->OC,
            "residual tensor geometry mismatch");
  p.N = in.N;
  p.OC = w->OC;
  p.n_tile = w->n_tile;
  p.row_mul = es;

  std::vector<SubDesc> subs;
  std::vector<StepDesc> steps;
  make_schedule(w, subs, steps);
  p.n_sub = static_cast<int>(subs.size());
  p.n_steps = static_cast<int>(steps.size());
  int maxshift = 0;
  p.fused = w->fused;
  p.first_second_step = -1;
  for (int k = 0; k < p.n_steps; ++k) {
    const int sh = steps[k].da * p.Wp + steps[k].db;
    p.step_a16[k] = static_cast<uint16_t>(sh * rowb / 16) | (steps[k].second ? kStepSecond : 0);
    if (steps[k].second && p.first_second_step < 0) p.first_second_step = k;
    maxshift = std::max(maxshift, sh);
  }
  const bool in_planes = in.planes && es == 2;
  DLQ_ARG(ctx, !in.planes || (es == 2 && in.rows() % 2 == 0 && in.W % 2 == 0), "plane layout is for stride-2 convs over even tensors");
  const int rows_half = in.planes ? in.plane_rows : in.rows() / 2;
  DLQ_ARG(ctx, !in.planes || in.plane_rows >= in.rows() / 2, "plane layout: plane_rows smaller than the tensor");
  for (int s = 0; s < p.n_sub; ++s) {
    p.sub_c0[s] = static_cast<int16_t>(subs[s].c0);
    if (in_planes) {
      // parity plane (pr, pc) of the physical start (row 2 v0 + PR + rel, col col0): dense box inside that plane
      const int srow = in.PR + subs[s].row_off_rel, scol = subs[s].col0;
      const int pr = srow & 1, pc = scol & 1;
      p.sub_col0[s] = static_cast<int16_t>((scol - pc) / 2);
      // (the plane's first row is a 32-bit field of its own; only the offset inside the plane is 16-bit)
      DLQ_ARG(ctx, (srow >> 1) < 32768, "sub-patch row offset does not fit 16 bits");
      p.sub_row_off[s] = static_cast<int16_t>(srow >> 1);
      p.sub_plane_row[s] = (pr * 2 + pc) * rows_half;
    } else {
      p.sub_col0[s] = static_cast<int16_t>(subs[s].col0);
      p.sub_row_off[s] = static_cast<int16_t>(in.PR + subs[s].row_off_rel);
      p.sub_plane_row[s] = 0;
    }
  }
  if (in_planes) p.row_mul = 1;
  {
    int k = 0;
    for (int s = 0; s < p.n_sub; ++s) {
      p.sub_step0[s] = static_cast<int16_t>(k);
      while (k < p.n_steps && steps[k].sub == s) ++k;
    }
    p.sub_step0[p.n_sub] = static_cast<int16_t>(p.n_steps);
  }

  // ---- tile shape.  CTA pairs: one 128-position tile per CTA, super-tile = a whole number of virtual rows (both
  // CTAs then view their patches through identical descriptors).  Single CTA: MT tiles per super-tile.
  const int n_tiles = w->OC / w->n_tile;
  p.n_tiles = n_tiles;
  p.two = (w->n_tile == 128 && p.Wp <= kTileM && !dlq_dbg_env("DLQ_DBG_NO_PAIR")) ? 1 : 0;
  if (dlq_dbg_env("DLQ_DBG_PAIR64") && w->n_tile == 64 && p.Wp <= 2 * kTileM && (w->kind != CONV_STEM || atoi(dlq_dbg_env("DLQ_DBG_PAIR64")) > 1)) p.two = 1;
  const int ncta = p.two ? 2 : 1;
  p.w_rows = w->n_tile / ncta;
  p.step_bytes = static_cast<uint32_t>(p.w_rows) * rowb;
  const size_t budget = ctx->smem_optin - 1024 /*alignment slack*/ - 1024 /*barriers, step table, flag words*/ -
                        16 * static_cast<size_t>(w->OC) /*alpha, beta (x2 when fused)*/ - 16 * kEpiStageBytes;
  const uint32_t b_stage_bytes = (p.step_bytes + 1023u) & ~1023u;
  int MT = std::max(1, 256 / p.n_tile);
  if (w->fused) MT = 1;      // two accumulator blocks per tile: 2 * n_tile columns per stage
  // Small problems (the latency configurations): fewer tiles per item while the items do not fill the SMs - an item's MMAs
  // and its epilogue are serial inside a CTA, so more and smaller items shorten the layer (batch 1: 8.5 -> ~6 us per
  // layer1 conv).  Chain members keep the chain's static tile shape.
  if (allow_resident && !w->fused) {
    const long long last = (static_cast<long long>(in.N - 1) * p.Pv + (p.Ho - 1)) * p.Wp + (p.Wo - 1);
    while (MT > 1) {
      const int ss = p.two ? ((MT * kTileM) / p.Wp) * p.Wp : MT * kTileM;
      const long long items = ((last / ss + 1 + ncta - 1) / ncta) * n_tiles;
      if (items >= ctx->num_sms / ncta) break;
      MT >>= 1;
    }
  }
  if (const char* e = dlq_dbg_env("DLQ_DBG_MT")) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && v * p.n_tile <= 512) MT = v; }
  int NR = 0;
  const size_t all_b = static_cast<size_t>(p.n_steps) * b_stage_bytes;
  for (;; MT >>= 1) {
    const int in_patch_max = p.two ? 0 : p.Wp - 1;
    NR = (in_patch_max + MT * kTileM - 1 + maxshift) / p.Wp + 1;
    p.tma_bytes = NR * p.Wp * rowb;
    p.sub_bytes = (p.tma_bytes + 1023) & ~1023;
    // weights resident in smem when they are small (stem, layer1, stride-2 / 1x1 convs): no per-item re-fetch
    p.b_resident = (allow_resident && n_tiles == 1 && all_b <= 80 * 1024 && all_b + 2 * static_cast<size_t>(p.sub_bytes) <= budget) ? 1 : 0;
    const int min_a = 2;
    if (p.b_resident) {
      p.b_stages = p.n_steps;
    } else {
      const size_t left = budget > static_cast<size_t>(min_a) * p.sub_bytes ? budget - static_cast<size_t>(min_a) * p.sub_bytes : 0;
      const int b_cap = dlq_dbg_env("DLQ_DBG_B_CAP") ? atoi(dlq_dbg_env("DLQ_DBG_B_CAP")) : 8;     // (tuning)
      p.b_stages = static_cast<int>(std::min<size_t>(std::min(p.n_steps, b_cap), left / b_stage_bytes));
    }
    const size_t b_bytes = static_cast<size_t>(std::max(p.b_stages, 0)) * b_stage_bytes;
    const int want_a = std::min(dlq_dbg_env("DLQ_DBG_A_CAP") ? atoi(dlq_dbg_env("DLQ_DBG_A_CAP")) : 4, p.n_sub + 2);
    const int fit_a = budget > b_bytes ? static_cast<int>((budget - b_bytes) / p.sub_bytes) : 0;
    p.a_stages = std::min(want_a, fit_a);
    if ((p.a_stages >= 2 && p.b_stages >= std::min(p.n_steps, 3) && es * NR <= 256) || MT == 1) break;
  }
  // debug / tuning overrides (environment; not used by tests or the benchmark)
  if (const char* e = dlq_dbg_env("DLQ_DBG_A_STAGES")) p.a_stages = atoi(e);
  if (const char* e = dlq_dbg_env("DLQ_DBG_B_STAGES")) { if (!p.b_resident) p.b_stages = atoi(e); }
  DLQ_ARG(ctx, p.a_stages >= 1 && p.b_stages >= 1 && es * NR <= 256 && es * p.Wp <= 256 && w->OC <= 2048,
          "conv patch does not fit shared memory / TMA box");
  p.MT = MT;
  p.acc_stages = (2 * MT * p.n_tile * (w->fused ? 2 : 1) <= 512) ? 2 : 1;
  p.super_stride = p.two ? ((MT * kTileM) / p.Wp) * p.Wp : MT * kTileM;
  const long long total_pos = static_cast<long long>(in.N) * p.Pv * p.Wp;
  DLQ_ARG(ctx, total_pos + 4LL * MT * kTileM < (1LL << 31), "batch too large for 32-bit position arithmetic");
  p.total_pos = static_cast<int>(total_pos);
  // super-tiles up to the one holding the last VALID position (last image, last row, last column): the pad rows behind
  // the last image would only make items whose every position is garbage (and, with dependency flags, units that no
  // consumer ever waits for)
  const long long last_valid = (static_cast<long long>(in.N - 1) * p.Pv + (p.Ho - 1)) * p.Wp + (p.Wo - 1);
  p.num_super = static_cast<int>(last_valid / p.super_stride) + 1;
  p.n_items = ((p.num_super + ncta - 1) / ncta) * n_tiles;
  // position decode by multiply-high instead of division (div_magic in conv_kernel.cuh): floor(2^32 / d)
  auto conv_magic = [](int d) { return d <= 1 ? 0xFFFFFFFFu : static_cast<uint32_t>((1ULL << 32) / static_cast<unsigned>(d)); };
  p.wp_magic = conv_magic(p.Wp);
  p.pv_magic = conv_magic(p.Pv);

  // ---- epilogue
  p.alpha = alpha;
  p.beta = beta;
  p.residual = residual ? residual->ptr : nullptr;
  p.res_PR = residual ? residual->PR : 0;
  p.res_mul = res_mul;
  p.relu = relu;
  p.out = out.ptr;
  p.out_PR = out.PR;
  p.out_planes = out.planes;
  p.out_rows_half = out.planes ? out.plane_rows : out.rows() / 2;
  DLQ_ARG(ctx, !out.planes || (out.rows() % 2 == 0 && out.W % 2 == 0 && !w->fused), "plane-layout output needs even rows / width");
  p.acc_out = acc_out;
  if (w->fused) {
    DLQ_ARG(ctx, second && second->out.ptr && second->alpha && second->beta && !residual && !acc_out && MT == 1,
            "fused shortcut conv: needs its own output / alpha / beta, no residual, one tile per item");
    DLQ_ARG(ctx, second->out.N == in.N && second->out.H == p.Ho && second->out.W == p.Wo && second->out.C == w->OC,
            "fused shortcut conv: output geometry mismatch");
    p.alpha2 = second->alpha; p.beta2 = second->beta; p.relu2 = second->relu;
    p.out2 = second->out.ptr; p.out2_PR = second->out.PR;
  }
  if (dlq_dbg_env("DLQ_DBG_NO_STORE")) { p.out = nullptr; p.acc_out = nullptr; }
  if (const char* e = dlq_dbg_env("DLQ_DBG_FLAGS")) p.dbg = atoi(e);
  // DLQ_DBG_TIMES=1: per-CTA cycle counters (MMA warp: total / wait acc_empty / wait a_full / wait b_full;
  // epilogue warp 0: total / wait acc_full; A producer: total / wait a_empty), printed by launch_conv
#ifdef DLQ_TIMING
  if (dlq_dbg_env("DLQ_DBG_TIMES")) {
    static long long* buf = nullptr;
    if (!buf) cudaMalloc(&buf, 16 * sizeof(long long) * 1024);
    cudaMemset(buf, 0, 16 * sizeof(long long) * 1024);
    p.dbg_times = buf;
  }
#endif

  // ---- TMA descriptor over the row-padded NHWC input: dims (C bytes, W, rows)
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    ctx->err = "cuTensorMapEncodeTiled entry point not available";
    return DLQ_ERR_CUDA;
  }
  const int tes = in_planes ? 1 : es;       // plane layout: dense boxes inside a [C][W/2][4 * rows/2] tensor
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(in.C), static_cast<cuuint64_t>(in_planes ? in.W / 2 : in.W),
                        static_cast<cuuint64_t>(in_planes ? 4 * rows_half : in.rows())};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(in.C), static_cast<cuuint64_t>(in.C) * (in_planes ? in.W / 2 : in.W)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(rowb), static_cast<cuuint32_t>(tes * p.Wp),
                       static_cast<cuuint32_t>(tes * NR)};
  cuuint32_t estr[3] = {1, static_cast<cuuint32_t>(tes), static_cast<cuuint32_t>(tes)};
  const CUtensorMapSwizzle swz = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(&L->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, in.ptr, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return DLQ_ERR_CUDA;
  }

  // ---- TMA descriptor over the packed weight image: rows of ROWB bytes (already swizzled), box = one CTA's rows of a step
  {
    cuuint64_t wdim[2] = {static_cast<cuuint64_t>(rowb), static_cast<cuuint64_t>(n_tiles) * p.n_steps * w->n_tile};
    cuuint64_t wstr[1] = {static_cast<cuuint64_t>(rowb)};
    cuuint32_t wbox[2] = {static_cast<cuuint32_t>(rowb), static_cast<cuuint32_t>(p.w_rows)};
    cuuint32_t west[2] = {1, 1};
    const CUresult rw = enc(&L->tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, w->d_img, wdim, wstr, wbox, west,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rw != CUDA_SUCCESS) {
      ctx->err = "cuTensorMapEncodeTiled (weights) failed with CUresult " + std::to_string(static_cast<int>(rw));
      return DLQ_ERR_CUDA;
    }
  }

  // persistent grid: one CTA (or CTA pair) per SM (pair of SMs), work items dealt round-robin
  const int G = std::max(1, std::min(p.n_items, ctx->num_sms / ncta));
  L->grid = dim3(static_cast<unsigned>(G * ncta), 1, 1);
  L->block = dim3(128 + 8 * 32, 1, 1);
  L->smem = 1024 + static_cast<size_t>(p.a_stages) * p.sub_bytes + static_cast<size_t>(p.b_stages) * b_stage_bytes +
            4 * sizeof(float) * w->OC + 16 * kEpiStageBytes + 2 * (kMaxSteps + 8) +
            8 * (2 * p.a_stages + 2 * p.b_stages + 3 * p.acc_stages) + 32;
  return DLQ_OK;
}

template <int ROWB, bool TWO, bool FP8>
static int launch_t(dlq_ctx* ctx, const ConvLaunch& L) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = L.grid;
  cfg.blockDim = L.block;
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2u : 1u;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: this kernel's prologue and weight loads overlap the tail of the previous kernel
  // on the stream; the kernel itself waits (griddepcontrol.wait) before it touches activations
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->no_pdl ? 1 : 2;
  DLQ_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_i8_kernel<ROWB, TWO, FP8>, L.tmap, L.tmap_w, L.p));
  if (L.p.dbg_times) {
    cudaStreamSynchronize(ctx->stream);
    const int nb = static_cast<int>(L.grid.x);
    std::vector<long long> h(static_cast<size_t>(nb) * 16);
    cudaMemcpy(h.data(), L.p.dbg_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double a[16] = {0}, mx[16] = {0};
    int cnt[16] = {0};
    for (int b = 0; b < nb; ++b)
      for (int j = 0; j < 16; ++j)
        if (h[static_cast<size_t>(b) * 16 + j]) {
          a[j] += static_cast<double>(h[static_cast<size_t>(b) * 16 + j]); ++cnt[j];
          mx[j] = std::max(mx[j], static_cast<double>(h[static_cast<size_t>(b) * 16 + j]));
        }
    for (int j = 0; j < 16; ++j) if (cnt[j]) a[j] /= cnt[j];
    {
      long long s0 = h[12], s1 = h[12], e0 = h[13], e1 = h[13];
      for (int b = 0; b < nb; ++b) {
        s0 = std::min(s0, h[static_cast<size_t>(b) * 16 + 12]); s1 = std::max(s1, h[static_cast<size_t>(b) * 16 + 12]);
        e0 = std::min(e0, h[static_cast<size_t>(b) * 16 + 13]); e1 = std::max(e1, h[static_cast<size_t>(b) * 16 + 13]);
      }
      fprintf(stderr, "[dbg_span] globaltimer ns relative to the first CTA entry: last entry %lld | first exit %lld | last exit %lld\n",
              s1 - s0, e0 - s0, e1 - s0);
    }
    fprintf(stderr, "[dbg_phases] since kernel entry (mean / max over CTAs): prologue %.0f / %.0f | issuer done %.0f / %.0f | epilogue done %.0f / %.0f | CTA done %.0f / %.0f\n",
            a[8], mx[8], a[9], mx[9], a[10], mx[10], a[11], mx[11]);
    fprintf(stderr, "[dbg_times] grid %ux%d MT=%d n_tile=%d trips=%d | mma: total %.0f wait_acc %.0f wait_a %.0f wait_b %.0f | "
                    "epi: total %.0f wait_acc_full %.0f | prodA: total %.0f wait_a_empty %.0f (cycles, mean over CTAs)\n",
            L.grid.x, L.p.two ? 2 : 1, L.p.MT, L.p.n_tile, (L.p.n_items + nb / (L.p.two ? 2 : 1) - 1) / (nb / (L.p.two ? 2 : 1)),
            a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
  }
  return DLQ_OK;
}

template <bool TWO, bool FP8>
static int launch_rowb(dlq_ctx* ctx, const ConvLaunch& L) {
  switch (L.rowb) {
    case 32: return launch_t<32, TWO, FP8>(ctx, L);
    case 64: return launch_t<64, TWO, FP8>(ctx, L);
    case 128: return launch_t<128, TWO, FP8>(ctx, L);
  }
  ctx->err = "internal: bad rowb";
  return DLQ_ERR_ARG;
}

// every instantiation's dynamic shared-memory limit, once per context (= per device): no function-static state, so
// contexts on different devices / threads are independent (dlq.h), and a cubin that does not load fails dlq_create
namespace {
typedef void (*ChainKernel)(const ChainParams);
ChainKernel chain_kernel(bool two, bool fp8) {
  if (two) return fp8 ? conv_chain_kernel<true, 2, 128, true> : conv_chain_kernel<true, 2, 128, false>;
  return fp8 ? conv_chain_kernel<false, 4, 64, true> : conv_chain_kernel<false, 4, 64, false>;
}
}  // namespace

template <int ROWB, bool TWO, bool FP8>
static cudaError_t configure_t(dlq_ctx* ctx) {
  return cudaFuncSetAttribute(conv_i8_kernel<ROWB, TWO, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(ctx->smem_optin));
}
int configure_conv_kernels(dlq_ctx* ctx) {
  DLQ_CUDA(ctx, (configure_t<32, false, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<64, false, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<128, false, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<32, true, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<64, true, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<128, true, false>(ctx)));
  DLQ_CUDA(ctx, (configure_t<32, false, true>(ctx)));
  DLQ_CUDA(ctx, (configure_t<64, false, true>(ctx)));
  DLQ_CUDA(ctx, (configure_t<128, false, true>(ctx)));
  DLQ_CUDA(ctx, (configure_t<32, true, true>(ctx)));
  DLQ_CUDA(ctx, (configure_t<64, true, true>(ctx)));
  DLQ_CUDA(ctx, (configure_t<128, true, true>(ctx)));
  for (int two = 0; two < 2; ++two)
    for (int fp8 = 0; fp8 < 2; ++fp8)
      DLQ_CUDA(ctx, cudaFuncSetAttribute(chain_kernel(two != 0, fp8 != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(ctx->smem_optin)));
  return DLQ_OK;
}

// ---- dependency flags (conv_kernel.cuh "dependency flags")
int conv_flag_units(const ConvLaunch& L) { return L.p.n_items / L.p.n_tiles; }
void conv_set_flags(ConvLaunch* L, unsigned int* done, unsigned int* dep_err) {
  L->p.done = done;
  L->p.dep_err = dep_err;
}
void conv_add_dep(ConvLaunch* consumer, const ConvLaunch& producer, int s, int lo, int hi) {
  ConvKernelParams& c = consumer->p;
  const ConvKernelParams& q = producer.p;
  if (c.n_deps >= 2 || !q.done) return;
  for (int i = 0; i < c.n_deps; ++i)
    if (c.dep[i].done == q.done && c.dep[i].s == s && c.dep[i].lo >= lo && c.dep[i].hi >= hi) return;   // already covered
  ConvDep& d = c.dep[c.n_deps++];
  const int ncta = q.two ? 2 : 1;
  d.done = q.done;
  d.target = static_cast<unsigned int>(q.n_tiles) * 8u * static_cast<unsigned int>(ncta);
  d.unit = q.super_stride * ncta;
  d.Pv = q.Pv; d.Wp = q.Wp; d.Wo = q.Wo; d.H = q.Ho;
  d.s = s; d.lo = lo; d.hi = hi;
}

// ---- conv chain (conv_chain.cuh)
int plan_chain(dlq_ctx* ctx, const ConvLaunch* const* layers, int n_layers, ChainLaunch* out) {
  DLQ_ARG(ctx, n_layers >= 1 && n_layers <= kMaxChainLayers, "conv chain: 1 .. kMaxChainLayers layers");
  ChainParams& cp = out->cp;
  memset(&cp, 0, sizeof(cp));
  out->fp8 = layers[0]->fp8;
  out->two = layers[0]->p.two;
  const int want_mt = out->two ? 2 : 4, want_nt = out->two ? 128 : 64;      // the two tile shapes the kernel is built for
  int a_stage = 0, b_stage = 0, oc_max = 0;
  for (int l = 0; l < n_layers; ++l) {
    const ConvLaunch& L = *layers[l];
    const ConvKernelParams& p = L.p;
    DLQ_ARG(ctx, (L.rowb == 128 || L.rowb == 64) && p.two == out->two && p.MT == want_mt && p.n_tile == want_nt && p.acc_stages == 2 &&
                     !p.fused && !p.b_resident && !p.acc_out && L.fp8 == out->fp8 && p.alpha && p.beta && p.n_steps <= kMaxSteps,
            "conv chain: a layer does not have the chain's static configuration");
    a_stage = std::max(a_stage, p.sub_bytes);
    b_stage = std::max(b_stage, static_cast<int>((p.step_bytes + 1023u) & ~1023u));
    oc_max = std::max(oc_max, p.OC);
  }
  const size_t fixed = 1024 /*alignment slack*/ + 2 * sizeof(float) * oc_max + 16 * kEpiStageBytes + 2 * 2 * (kMaxSteps + 8) + 1024;
  DLQ_ARG(ctx, ctx->smem_optin > fixed + 2 * static_cast<size_t>(a_stage) + 3 * static_cast<size_t>(b_stage), "conv chain: does not fit shared memory");
  const size_t budget = ctx->smem_optin - fixed;
  int b_stages = static_cast<int>(std::min<size_t>(8, (budget - 2 * static_cast<size_t>(a_stage)) / b_stage));
  int a_stages = static_cast<int>(std::min<size_t>(4, (budget - static_cast<size_t>(b_stages) * b_stage) / a_stage));
  DLQ_ARG(ctx, a_stages >= 2 && b_stages >= 3, "conv chain: rings too shallow");
  cp.n_layers = n_layers;
  cp.a_stages = a_stages; cp.b_stages = b_stages;
  cp.a_stage_bytes = a_stage; cp.b_stage_bytes = b_stage;
  cp.oc_max = oc_max;
  // one CTA (pair) per SM (pair), all co-resident (the CTAs wait for each other through the dependency flags)
  const int ncta = out->two ? 2 : 1;
  int G = ctx->num_sms / ncta;
  out->block = dim3(128 + 8 * 32, 1, 1);
  out->smem = 1024 + static_cast<size_t>(a_stages) * a_stage + static_cast<size_t>(b_stages) * b_stage + 2 * sizeof(float) * oc_max +
              16 * kEpiStageBytes + 2 * 2 * (kMaxSteps + 8) + 8 * (2 * a_stages + 2 * b_stages + 4) + 32;
  if (out->two) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * G), 1, 1);
    cfg.blockDim = out->block;
    cfg.dynamicSmemBytes = out->smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, chain_kernel(true, out->fp8 != 0), &cfg);
    if (e == cudaSuccess && max_clusters > 0) G = std::min(G, max_clusters);
    else cudaGetLastError();
  }
  out->grid = dim3(static_cast<unsigned>(ncta * G), 1, 1);
  long long shift = 0;
  for (int l = 0; l < n_layers; ++l) {
    ChainLayer& C = cp.layer[l];
    C.p = layers[l]->p;
    C.tm0 = layers[l]->tmap;
    C.tmw = layers[l]->tmap_w;
    C.rowb = layers[l]->rowb;
    C.item_shift = static_cast<int>(shift % G);       // the chain's items are dealt round-robin across layer boundaries
    shift += C.p.n_items;
  }
  return DLQ_OK;
}

int launch_chain(dlq_ctx* ctx, ChainLaunch& C) {
  // launch modes, tried in order from the last one that worked: 0 = cooperative + programmatic stream serialization,
  // 1 = cooperative, 2 = neither.  Mode 2 gives no co-residency guarantee and is taken only when the driver refuses a
  // cooperative launch (seen under the ncu profiler, which runs every kernel alone on the device anyway) and the grid is
  // one CTA per SM of an otherwise idle device.
  const int first_mode = C.mode;
  for (; C.mode <= 2; ++C.mode) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = C.grid;
    cfg.blockDim = C.block;
    cfg.dynamicSmemBytes = C.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[3];
    int na = 0;
    if (C.two) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (C.mode <= 1) {
      attr[na].id = cudaLaunchAttributeCooperative;       // every CTA pair resident at once
      attr[na].val.cooperative = 1;
      ++na;
    }
    if (C.mode == 0 && !ctx->no_pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, chain_kernel(C.two != 0, C.fp8 != 0), C.cp);
    if (e == cudaSuccess) {
      if (C.mode == 2 && first_mode != 2 && !C.warned) {
        fprintf(stderr, "[dlq] conv chain: cooperative launch refused by the driver; launched without the co-residency guarantee\n");
        C.warned = true;
      }
      return DLQ_OK;
    }
    if (C.mode == 2 || (C.mode == 1 && static_cast<int>(C.grid.x) != ctx->num_sms)) DLQ_CUDA(ctx, e);
    cudaGetLastError();
  }
  ctx->err = "conv chain: launch failed";
  return DLQ_ERR_CUDA;
}

int launch_conv(dlq_ctx* ctx, const ConvLaunch& L) {
  if (L.fp8) return L.p.two ? launch_rowb<true, true>(ctx, L) : launch_rowb<false, true>(ctx, L);
  return L.p.two ? launch_rowb<true, false>(ctx, L) : launch_rowb<false, false>(ctx, L);
}

}  // namespace dlq
