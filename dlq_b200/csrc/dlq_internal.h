// dlq_internal.h — private types shared by the translation units of libdlq_b200.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/dlq.h"
#include "conv_chain.cuh"

struct dlq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  size_t smem_optin = 0;
  std::string err;
  // workspace of the per-layer NCHW entry points (never used by the fused network path): ONE allocation, sized by
  // dlq_workspace_reserve() - after which those entry points never allocate - or grown on demand when the caller
  // never reserved (SURVEY 8b "Ownership"; the reference instead allocates per call, R/infer_e2e.cu:128-130)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  bool ws_reserved = false;
  void* small = nullptr;    // 256 B of device scratch for reductions (dlq_compare_f32), allocated at dlq_create
  bool no_pdl = false;      // (TIMING builds only: DLQ_DBG_NO_PDL, read once at dlq_create)
};

// Tuning / debug switches are honoured only by a library built with `make TIMING=1`; in the product build every
// lookup is a compile-time nullptr, so no environment variable can change what the library computes or stores.
static inline const char* dlq_dbg_env(const char* name) {
#ifdef DLQ_TIMING
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

#define DLQ_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
      return DLQ_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define DLQ_ARG(ctx, cond, msg)                   \
  do {                                            \
    if (!(cond)) {                                \
      (ctx)->err = std::string("bad argument: ") + (msg); \
      return DLQ_ERR_ARG;                         \
    }                                             \
  } while (0)

namespace dlq {

// Row-padded NHWC int8 activation tensor:
//   rows:  [PR zero rows][image 0: H rows][PR zero rows][image 1: H rows] ...   each row = W*C bytes
// The zero rows are the vertical conv padding; they are written once (memset) and never again.
struct Act {
  int8_t* ptr = nullptr;
  int N = 0, H = 0, W = 0, C = 0, PR = 0;
  // planes = 1: the same rows x W pixels stored as four parity planes [row&1][col&1][rows/2][W/2][C] (rows and W
  // even).  Tensors read only by stride-2 convs use it: each parity plane is then a dense TMA box instead of an
  // elementStrides = 2 gather (which moves 4x the bytes through the TMA unit and writes shared memory pixel-wise).
  int planes = 0;
  int plane_rows = 0;   // rows per plane = (rows of the ALLOCATED tensor) / 2: fixed at allocation so that the planes,
                        // and with them the zero pad rows, do not move when a smaller batch uses the same buffer
  int rows() const { return PR + N * (H + PR); }
  size_t bytes() const { return static_cast<size_t>(rows()) * W * C; }
  size_t row_index(int n, int h) const { return static_cast<size_t>(PR) + static_cast<size_t>(n) * (H + PR) + h; }
};

enum ConvKind { CONV_S1 = 0, CONV_S2_3x3 = 1, CONV_S2_1x1 = 2, CONV_STEM = 3 };

}  // namespace dlq

// Packed conv weights: int8 per-output-channel quantised, stored as the exact shared-memory images
// (swizzle applied) the conv kernel bulk-copies, one image per (n-tile, K step).
struct dlq_conv_weights {
  int OC = 0, IC = 0, kH = 0, kW = 0, sH = 1, sW = 1, pH = 0, pW = 0;
  dlq::ConvKind kind = dlq::CONV_S1;
  int rowb = 0;        // bytes of K per A row (16 stem / 64 / 128)
  int kb = 1;          // channel blocks
  int n_tile = 0;      // output channels per CTA
  int n_steps = 0;
  uint32_t step_bytes = 0;
  uint8_t* d_img = nullptr;   // device: [OC/n_tile][n_steps][step_bytes]
  int fp8 = 0;                // 1: the bytes are E4M3 (QUANT_SPEC section 6), else signed int8
  int fused = 0;              // 1: a 1x1/s2 shortcut conv (second_q, OC x IC) rides on this 3x3/s2 conv's patch loads
  std::vector<int8_t> second_q;
  std::vector<int8_t> q_oihw; // host copy of the quantised weights (tests / checkpoints)
  std::vector<float> scale;   // per-output-channel scale
  int device = 0;
  int fc_O = 0, fc_I = 0;     // packed by dlq_fc_weights_pack*: the logical [O, I] of the matrix (OC / IC are its padded sizes)
};

namespace dlq {

// conv_plan.cu
int pack_conv_weights(dlq_ctx* ctx, const int8_t* wq, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                      int pW, dlq_conv_weights* out, const int8_t* second_wq = nullptr);
// epilogue of the fused shortcut conv (its own folded constants, ReLU flag and output tensor)
struct SecondConv {
  const float* alpha = nullptr;
  const float* beta = nullptr;
  int relu = 0;
  Act out;
};
// PR the conv requires of its input tensor (and whether the image pitch must be even)
int conv_required_in_pr(const dlq_conv_weights* w);
struct ConvLaunch {
  CUtensorMap tmap;     // activations (row-padded NHWC), 3-D
  CUtensorMap tmap_w;   // packed weight image, 2-D
  ConvKernelParams p;
  dim3 grid, block;
  size_t smem = 0;
  int rowb = 0;
  int fp8 = 0;          // operands are E4M3 bytes, accumulators FP32 (else S8 / S32)
};
// Build the launch for conv `w` reading `in` and writing `out` (either of out.ptr / acc_out may be null).
int plan_conv(dlq_ctx* ctx, const dlq_conv_weights* w, const Act& in, const Act& out, const float* alpha,
              const float* beta, const Act* residual, float res_mul, int relu, int32_t* acc_out, ConvLaunch* L,
              const SecondConv* second = nullptr, bool allow_resident = true);   // (false: stream the weight steps - conv chain)
int launch_conv(dlq_ctx* ctx, const ConvLaunch& L);
// Several planned convs as ONE persistent cooperative launch (conv_chain.cuh).  plan_chain returns DLQ_ERR_ARG (and
// leaves ctx->err) when a layer does not have the chain's static configuration; the caller then launches them one by one.
struct ChainLaunch {
  ChainParams cp;
  dim3 grid, block;
  size_t smem = 0;
  int fp8 = 0;
  int two = 1;              // CTA pairs (layer2..4 shape) or single CTAs (layer1 shape)
  int mode = 0;             // launch attributes that worked last (launch_chain): 0 cooperative + PDL, 1 cooperative, 2 plain
  bool warned = false;
};
int plan_chain(dlq_ctx* ctx, const ConvLaunch* const* layers, int n_layers, ChainLaunch* out);
int launch_chain(dlq_ctx* ctx, ChainLaunch& C);
// dependency flags between the conv launches of one forward (conv_kernel.cuh "dependency flags"):
int conv_flag_units(const ConvLaunch& L);                                          // counters the launch needs
void conv_set_flags(ConvLaunch* L, unsigned int* done, unsigned int* dep_err);     // keep per-unit completion counters
// `consumer` waits, per item, for the units of `producer` holding rows [r*s - lo, r*s + hi] of its output rows r
void conv_add_dep(ConvLaunch* consumer, const ConvLaunch& producer, int s, int lo, int hi);
void conv_out_dims(const dlq_conv_weights* w, int H, int W, int* OH, int* OW);

// elementwise.cu (all enqueue on ctx->stream)
int nchw_to_act_i8(dlq_ctx* ctx, const int8_t* x, const Act& a);          // dense NCHW -> row-padded NHWC
int act_to_nchw_i8(dlq_ctx* ctx, const Act& a, int8_t* y);                // row-padded NHWC -> dense NCHW
int nhwc_to_nchw_i32(dlq_ctx* ctx, const int32_t* x, int N, int C, int HW, int32_t* y);
int nchw_i8_to_stem_s2d(dlq_ctx* ctx, const int8_t* x, int N, int H, int W, const Act& a);   // C=3 int8 NCHW -> s2d
int quantize_input_s2d(dlq_ctx* ctx, const float* x, int N, int H, int W, float inv_scale, const Act& a, int fp8 = 0,
                       unsigned long long* stamp = nullptr);
int preprocess_u8_s2d(dlq_ctx* ctx, const uint8_t* x_hwc, int N, int H, int W, const uint8_t* lut_dev, const Act& a,
                      unsigned long long* stamp = nullptr);
int maxpool_act(dlq_ctx* ctx, const Act& in, const Act& out, unsigned long long* stamp = nullptr);
// zero / n_zero: the forward's dependency counters, cleared by this last kernel of the forward once the last conv grid
// has completed (conv_kernel.cuh "dependency flags")
int gap_fc_act(dlq_ctx* ctx, const Act& in, float scale_over_hw, float inv_gap_scale, const int8_t* fc_w,
               const float* fc_scale, const float* fc_bias, int O, int8_t* gap_q, float* logits,
               unsigned long long* stamp = nullptr, unsigned int* zero = nullptr, int n_zero = 0);

int gap_fc_act_e4m3(dlq_ctx* ctx, const Act& in, float scale_over_hw, float inv_gap_scale, const int8_t* fc_w,
                    const float* fc_scale, const float* fc_bias, int O, int8_t* gap_q, float* logits, unsigned long long* stamp,
                    unsigned int* zero, int n_zero);

// dlq_api.cu: per-row symmetric quantisation of a [rows, K] fp32 matrix (QUANT_SPEC 1 / 6), host
void quantize_rows(const float* w, int rows, int K, std::vector<int8_t>& q, std::vector<float>& s);
void quantize_rows_e4m3(const float* w, int rows, int K, std::vector<int8_t>& q, std::vector<float>& s);

// workspace: make sure ctx->ws holds `bytes`; DLQ_OK, DLQ_ERR_ARG (reserved workspace too small) or DLQ_ERR_CUDA (OOM)
int ctx_workspace(dlq_ctx* ctx, size_t bytes);
float inv_scale(float s);
// per-device kernel attributes (max dynamic shared memory), set once per context in dlq_create; also proves that the
// sm_100a images load on this device
int configure_conv_kernels(dlq_ctx* ctx);
int configure_elementwise_kernels(dlq_ctx* ctx);

}  // namespace dlq
