/* dlq.h — C ABI of libdlq_b200.so: B200-native INT8 inference operators for the DLQ lab's ResNet-18
 * kernel stack.
 *
 * The reference (yeontachi/DLQ, CUDA/resnet18-kernel-lab) has no FFI layer: its boundary is a set of
 * `extern "C" __global__` FP32 kernels launched directly from host code plus a few host launcher
 * helpers.  Each entry point below is the host-callable replacement for one of those, with the same
 * argument order and meaning (device pointers, dense NCHW, dims as int), cited as
 *   K/ = CUDA/resnet18-kernel-lab/cpp/fp32/kernels/      R/ = .../cpp/fp32/runtime/
 *
 * Conventions
 *   - every function returns int: 0 ok, 1 bad argument, 3 CUDA launch/runtime error
 *     (the reference's exit codes, R/utils.hpp:34-45; 2 stays reserved for parity failures in tests);
 *     dlq_last_error_string() gives the text.  Nothing here ever calls exit().
 *   - all tensor pointers are DEVICE pointers unless the name says host; the caller owns every buffer.
 *   - calls enqueue work on the context's stream and return without synchronising.  The context's own stream is
 *     NON-BLOCKING: it has no implicit ordering with the legacy default stream or any other stream.  Buffers the caller
 *     fills or zeroes on another stream must be complete (event / synchronise) before a call that touches them, or the
 *     caller adopts its own stream with dlq_set_stream().
 *   - one context per device; a context is not thread-safe, different contexts are independent (no shared static
 *     state: kernel attributes are set per context in dlq_create).
 *   - no entry point allocates device memory after set-up: models allocate at create (host-staging buffers at the first
 *     host-buffer call), the per-layer NCHW entry points use the context workspace (dlq_workspace_reserve).
 *   - allocation failures and CUDA runtime errors return 3, argument errors 1.
 *   - quantisation arithmetic is defined by spec/QUANT_SPEC.md (the reference defines none).
 */
#ifndef DLQ_H
#define DLQ_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DLQ_OK 0
#define DLQ_ERR_ARG 1
#define DLQ_ERR_CUDA 3

typedef struct dlq_ctx dlq_ctx;
typedef struct dlq_conv_weights dlq_conv_weights;
typedef struct dlq_resnet18 dlq_resnet18;

/* ------------------------------------------------------------------ context (replaces the reference's
 * implicit default-stream / cudaMalloc-per-call runtime, R/utils.hpp:122-160) */
int dlq_create(int device, dlq_ctx** out);
void dlq_destroy(dlq_ctx* ctx);
const char* dlq_last_error_string(const dlq_ctx* ctx);
int dlq_sync(dlq_ctx* ctx);
void* dlq_stream(dlq_ctx* ctx);                 /* the cudaStream_t work is enqueued on */
int dlq_set_stream(dlq_ctx* ctx, void* stream); /* adopt a caller-owned cudaStream_t */
const char* dlq_version(void);
/* Workspace of the per-layer NCHW entry points (dlq_conv2d_i8 / _fp8, dlq_fc_forward_*_tc): the reference's
 * conv2d_nchw_im2col_gemm allocates and frees three buffers per call (R/infer_e2e.cu:128-130); here the caller reserves
 * once - max over its calls of dlq_conv2d_workspace_bytes() / dlq_fc_workspace_bytes() - and those entry points never
 * allocate (a call that needs more then fails with 1).  Without a reservation the workspace grows on demand
 * (synchronising), which is convenient for tests, not for a hot path. */
int dlq_workspace_reserve(dlq_ctx* ctx, size_t bytes);
size_t dlq_workspace_bytes(const dlq_ctx* ctx); /* current size */

/* ------------------------------------------------------------------ quantise / dequantise helpers
 * (absent from the reference — SURVEY §8 a13; arithmetic per QUANT_SPEC §2) */
/* q = clamp(rne(x * fp32(1/scale)), -128, 127) */
int dlq_quantize_f32_i8(dlq_ctx* ctx, const float* x, size_t n, float scale, int8_t* q);
int dlq_dequantize_i8_f32(dlq_ctx* ctx, const int8_t* q, size_t n, float scale, float* x);
/* E4M3: q = e4m3_rn_satfinite(x * fp32(1/scale)); x' = float(q) * scale  (QUANT_SPEC 6) */
int dlq_quantize_f32_e4m3(dlq_ctx* ctx, const float* x, size_t n, float scale, uint8_t* q);
int dlq_dequantize_e4m3_f32(dlq_ctx* ctx, const uint8_t* q, size_t n, float scale, float* x);
/* x[n,c,hw] = q * scale[c] */
int dlq_dequantize_i8_f32_per_channel(dlq_ctx* ctx, const int8_t* q, int N, int C, int HW, const float* scale,
                                      float* x);

/* ------------------------------------------------------------------ convolution
 * replaces conv2d_nchw_im2col_gemm (R/infer_e2e.cu:102-136) = im2col_nchw (K/im2col.cu:6-58) +
 * sgemm_tiled (K/sgemm_tiled.cu:6-46), with bn_inference / add_inplace / relu_forward folded into the
 * epilogue (K/bn_inference.cu:6-28, K/add.cu:3-8, K/relu.cu:5-10). */

/* Pack OIHW fp32 weights (HOST pointer; the reference passes std::vector<float>, R/infer_e2e.cu:105) once:
 * per-output-channel symmetric int8 (QUANT_SPEC §1) in the kernel's shared-memory image order.
 * The per-channel scales are returned through w_scale_host[OC] (HOST, may be NULL). */
int dlq_conv_weights_pack(dlq_ctx* ctx, const float* w_oihw_host, int OC, int IC, int kH, int kW, int sH, int sW,
                          int pH, int pW, float* w_scale_host, dlq_conv_weights** out);
/* same from already-quantised int8 OIHW weights (HOST) */
int dlq_conv_weights_pack_i8(dlq_ctx* ctx, const int8_t* wq_oihw_host, int OC, int IC, int kH, int kW, int sH,
                             int sW, int pH, int pW, dlq_conv_weights** out);
/* FP8 (E4M3) variants, QUANT_SPEC section 6: per-output-channel scale = absmax / 448, round-to-nearest-even,
 * saturate-to-finite; the conv runs on tcgen05 kind::f8f6f4 with FP32 accumulation. */
int dlq_conv_weights_pack_fp8(dlq_ctx* ctx, const float* w_oihw_host, int OC, int IC, int kH, int kW, int sH, int sW,
                              int pH, int pW, float* w_scale_host, dlq_conv_weights** out);
int dlq_conv_weights_pack_e4m3(dlq_ctx* ctx, const uint8_t* wq_oihw_host, int OC, int IC, int kH, int kW, int sH,
                               int sW, int pH, int pW, dlq_conv_weights** out);
uint8_t dlq_f32_to_e4m3(float f); /* HOST helper: the conversion above for one value */
void dlq_conv_weights_free(dlq_conv_weights* w);

/* Fused epilogue, QUANT_SPEC §3 (requantisation-multiplier form; each line one binary32 rounding):
 *   t = fmaf((float)acc, alpha[oc], beta[oc]);
 *   if (residual) t = fmaf((float)r, res_mul, t);
 *   y = clamp(rne(t), relu ? 0 : -128, 127)           (ReLU == the lower clamp at 0)
 * with the output scale already folded in by the caller:
 *   alpha[oc] = s_x*s_w[oc]*gamma/sqrt(var+eps)/s_y,  beta[oc] = (beta_bn - mean*gamma/sqrt(var+eps))/s_y,
 *   res_mul = s_r/s_y      -- dlq_fold_bn() computes alpha/beta exactly as the spec prescribes. */
typedef struct {
  const float* alpha;     /* [OC] device */
  const float* beta;      /* [OC] device */
  const int8_t* residual; /* NCHW int8 [N,OC,OH,OW] device, or NULL */
  float res_mul;
  int relu;
} dlq_epilogue;

/* HOST helper: folded BN + requantisation constants (double arithmetic, rounded once to fp32; QUANT_SPEC §3).
 * g/b/m/v: BN gamma/beta/running_mean/running_var [OC]; s_w: per-channel weight scales [OC];
 * s_x / s_y: input / output activation scales.  Replaces bn_launch's per-call uploads (R/infer_e2e.cu:83-97). */
void dlq_fold_bn(const float* g, const float* b, const float* m, const float* v, float eps, const float* s_w,
                 float s_x, float s_y, int OC, float* alpha_host, float* beta_host);
/* HOST helper: fp32(s_r / s_y) */
float dlq_res_mul(float s_r, float s_y);

/* x: int8 NCHW [N,C,H,W]; y: int8 NCHW [N,OC,OH,OW] (may be NULL); acc_out: int32 NCHW raw accumulators
 * (may be NULL; parity/debug).  OH/OW are returned like the reference's int& OH, int& OW. */
int dlq_conv2d_i8(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                  const dlq_epilogue* ep, int8_t* y, int32_t* acc_out, int* OH, int* OW);

/* workspace bytes the call above needs for this shape (see dlq_workspace_reserve) */
size_t dlq_conv2d_workspace_bytes(const dlq_conv_weights* w, int N, int H, int W, int has_residual, int want_acc);

/* Same operator on E4M3 bytes (weights from dlq_conv_weights_pack_fp8/_e4m3): t = fmaf(acc_f32, alpha, beta) [+ residual],
 * y = e4m3(relu ? max(t, 0) : t); acc_out receives the raw FP32 accumulators.  FP32 accumulation order inside the
 * tensor core is unspecified, so parity with the oracle is by tolerance (QUANT_SPEC 6), not bit-exact. */
int dlq_conv2d_fp8(dlq_ctx* ctx, const uint8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                   const dlq_epilogue* ep, uint8_t* y, float* acc_out, int* OH, int* OW);

/* Native-layout variant (no NCHW transposes): activations as the library keeps them internally —
 * row-padded NHWC int8:  [PR zero rows][image 0: H rows of W*C bytes][PR zero rows][image 1] ...
 * The zero rows are the vertical padding; the caller zero-fills the buffer once (dlq_act_bytes) and the
 * kernels never write them.  x->PR must be >= dlq_conv_required_pad_rows(w); for stride-2 convs H+PR must be
 * even; the 3-channel stem takes the paired 2x2 space-to-depth image (H/2 rows x (W/2+3) pixel pairs x 32 B,
 * produced by dlq_stem_pack_input_i8). */
typedef struct {
  int8_t* ptr; /* device */
  int N, H, W, C, PR;
} dlq_act;
size_t dlq_act_bytes(int N, int H, int W, int C, int PR);
int dlq_conv_required_pad_rows(const dlq_conv_weights* w);
/* residual (may be NULL) shares y's geometry except PR; acc_out: dense NHWC int32 [N,OH,OW,OC] or NULL */
int dlq_conv2d_i8_act(dlq_ctx* ctx, const dlq_act* x, const dlq_conv_weights* w, const dlq_epilogue* ep,
                      const dlq_act* residual, const dlq_act* y, int32_t* acc_out);
/* plan-once / launch-many form of the same call (no host planning or descriptor encoding on the hot path) */
typedef struct dlq_conv_plan dlq_conv_plan;
int dlq_conv_plan_create(dlq_ctx* ctx, const dlq_act* x, const dlq_conv_weights* w, const dlq_epilogue* ep,
                         const dlq_act* residual, const dlq_act* y, int32_t* acc_out, dlq_conv_plan** out);
int dlq_conv_plan_launch(dlq_ctx* ctx, const dlq_conv_plan* plan);
void dlq_conv_plan_destroy(dlq_conv_plan* plan);
/* layout helpers: dense NCHW int8 <-> row-padded NHWC; [N,3,H,W] int8 -> stem space-to-depth */
int dlq_act_from_nchw_i8(dlq_ctx* ctx, const int8_t* x_nchw, const dlq_act* a);
int dlq_act_to_nchw_i8(dlq_ctx* ctx, const dlq_act* a, int8_t* y_nchw);
int dlq_stem_pack_input_i8(dlq_ctx* ctx, const int8_t* x_nchw, int N, int H, int W, const dlq_act* a);

/* ------------------------------------------------------------------ element-wise / pooling (standalone,
 * kept for per-layer API parity; the fused network path does not launch them) */
/* K/bn_inference.cu:6-28 — in place, y = g*((x-m)/sqrtf(v+eps)) + b, c = (idx/(OH*OW)) % C */
int dlq_bn_inference_f32(dlq_ctx* ctx, float* x, const float* g, const float* b, const float* m, const float* v,
                         float eps, int N, int C, int OH, int OW);
/* K/relu.cu:5-10 — in place, if (x < 0) x = 0 */
int dlq_relu_forward_f32(dlq_ctx* ctx, float* x, size_t n);
int dlq_relu_forward_i8(dlq_ctx* ctx, int8_t* x, size_t n);
/* K/add.cu:3-8 — y += x */
int dlq_add_inplace_f32(dlq_ctx* ctx, float* y, const float* x, size_t n);
/* residual add of two int8 tensors with different scales, requantised: QUANT_SPEC §4 */
int dlq_add_requant_i8(dlq_ctx* ctx, int8_t* y, float y_scale, const int8_t* x, float x_scale, size_t n, int relu,
                       float out_scale);
/* K/maxpool2d.cu:5-41 — 3x3 / stride 2 / pad 1, OOB taps skipped; int8 NCHW */
int dlq_maxpool2d_3x3_s2p1_nchw_i8(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, int8_t* y);
/* K/gap_global.cu:3-33 — per-channel mean; int8 in: y_f32 = (float)sum * fp32(in_scale/HW),
 * y_i8 = quantise(y_f32, out_scale).  Either output may be NULL. */
int dlq_gap_global_i8(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, float in_scale, float out_scale,
                      float* y_f32, int8_t* y_i8);
/* R/infer_e2e.cu:206-219 fc_forward — int8 W[O,I] (device) x int8 g[N,I] -> int32 ->
 * logits = fmaf((float)acc, scale[o], bias[o]) */
int dlq_fc_forward_i8(dlq_ctx* ctx, const int8_t* g, const int8_t* w, const float* scale, const float* bias, int N,
                      int O, int I, float* logits);
/* The same FC on the tensor cores (SURVEY 8f-4 "INT8/FP8 FC with the same GEMM core"): the layer is the 1x1
 * convolution of an [N,1,1,I] tensor, so it runs on the conv kernel (TMA + tcgen05.mma + TMEM) with an fp32 epilogue
 * logits = fmaf((float)acc, scale[o], bias[o]) - bit-identical to dlq_fc_forward_i8.  Weights are packed once from the
 * reference's [O, I] row-major layout (HOST pointers); I is padded to 64 / a multiple of 128 and O to a multiple of
 * 64 internally (max I 2048).  dlq_fc_weights_pack quantises fp32 rows per QUANT_SPEC 1 (fp8 = 0) or 6 (fp8 = 1) and
 * returns the row scales.  Free with dlq_conv_weights_free. */
int dlq_fc_weights_pack(dlq_ctx* ctx, const float* w_host, int O, int I, int fp8, float* w_scale_host,
                        dlq_conv_weights** out);
int dlq_fc_weights_pack_i8(dlq_ctx* ctx, const int8_t* wq_host, int O, int I, dlq_conv_weights** out);
int dlq_fc_weights_pack_e4m3(dlq_ctx* ctx, const uint8_t* wq_host, int O, int I, dlq_conv_weights** out);
size_t dlq_fc_workspace_bytes(const dlq_conv_weights* w, int N);
int dlq_fc_forward_i8_tc(dlq_ctx* ctx, const int8_t* g, const dlq_conv_weights* w, const float* scale, const float* bias,
                         int N, float* logits);
/* E4M3 x E4M3 with FP32 accumulation (kind::f8f6f4); parity by tolerance (QUANT_SPEC 6-7) */
int dlq_fc_forward_fp8(dlq_ctx* ctx, const uint8_t* g, const dlq_conv_weights* w, const float* scale, const float* bias,
                       int N, float* logits);
/* K/softmax.cu:6-47 — row softmax over [N,K] fp32 logits (optional head; uses expf) */
int dlq_softmax_f32(dlq_ctx* ctx, const float* x, int N, int K, float* y);

/* ------------------------------------------------------------------ whole network
 * replaces main()'s wiring in R/infer_e2e.cu:254-433 (stem -> layer1..4 -> GAP -> FC) for a batch. */
#define DLQ_NUM_CONVS 25 /* [0]=stem; block b: [1+3b]=conv1, [2+3b]=conv2, [3+3b]=downsample (NULL if none) */
#define DLQ_NUM_ACTS 27  /* activation scales: [0]=input, [1]=stem, block b: [2+3b..4+3b]=conv1/ds/out, [26]=gap */
typedef struct {
  /* HOST pointers, fp32, the reference's export layout (<key>.bin, tools/export_resnet18.py:85-92) */
  const float* conv_w[DLQ_NUM_CONVS]; /* OIHW */
  const float* bn_gamma[DLQ_NUM_CONVS];
  const float* bn_beta[DLQ_NUM_CONVS];
  const float* bn_mean[DLQ_NUM_CONVS];
  const float* bn_var[DLQ_NUM_CONVS];
  const float* fc_w; /* [1000,512] */
  const float* fc_b; /* [1000] */
  float act_scale[DLQ_NUM_ACTS];
  int fp8; /* 0: INT8 network (QUANT_SPEC 1-5).  1: E4M3 weights / activations, FP32 accumulation (QUANT_SPEC 6);
            * act_scale then maps each tensor's absmax to 448 instead of 127 */
} dlq_resnet18_weights;

int dlq_resnet18_create(dlq_ctx* ctx, const dlq_resnet18_weights* w, int max_batch, dlq_resnet18** out);
void dlq_resnet18_destroy(dlq_resnet18* m);
/* x: fp32 NCHW [N,3,224,224] device; logits: fp32 [N,1000] device.  N <= max_batch. */
int dlq_resnet18_forward(dlq_resnet18* m, const float* x, int N, float* logits);
/* same with HOST buffers (pinned or pageable): H2D copy, forward, D2H copy, synchronise */
int dlq_resnet18_forward_host(dlq_resnet18* m, const float* x_host, int N, float* logits_host);
/* uint8 image input (the adjacent stage of the reference's pipeline, tools/preprocess_to_bin.py:24-33: u8/255,
 * (x-mean)/std in float32, HWC -> CHW): normalisation + input quantisation become a 3x256 byte table applied on the
 * device, built with that fp32 arithmetic, so logits equal dlq_resnet18_forward on the tensor the reference's
 * preprocessing produces.  x_hwc: uint8 [N,224,224,3] (RGB).  Host variant: 150 KB/image of H2D instead of 602 KB. */
int dlq_resnet18_set_preprocess(dlq_resnet18* m, const float* mean3, const float* std3);
int dlq_resnet18_forward_u8(dlq_resnet18* m, const uint8_t* x_hwc, int N, float* logits);
int dlq_resnet18_forward_host_u8(dlq_resnet18* m, const uint8_t* x_hwc_host, int N, float* logits_host);
/* int8 NCHW copy of an internal checkpoint of the LAST forward: "stem_pool","layer1".."layer4","gap"
 * (names as R/infer_e2e.cu:297-426 dumps them).  out is a device pointer. */
int dlq_resnet18_checkpoint(dlq_resnet18* m, const char* name, int8_t* out);
/* Latency path (BASELINE config "batch 1"): capture one forward of batch N (x -> logits, both device pointers that
 * stay valid) into a CUDA graph once, then replay it with a single launch.  The reference instead pays ~45 launches,
 * allocations and weight uploads per image (R/infer_e2e.cu:230-441).  dlq_resnet18_graph_launch enqueues on the
 * context's stream and does not synchronise. */
int dlq_resnet18_graph_capture(dlq_resnet18* m, const float* x, int N, float* logits);
int dlq_resnet18_graph_launch(dlq_resnet18* m);
/* Pipelined host-buffer form: submit enqueues H2D -> forward of the WHOLE batch -> D2H and returns; up to two submits may
 * be in flight (double-buffered staging), so the copy of batch k+1 overlaps the forward of batch k and the logits copy of
 * batch k-1.  dlq_resnet18_wait blocks until the OLDEST outstanding submit's logits are in host memory (no-op when
 * nothing is outstanding).  x_host / logits_host must stay valid until the matching wait; use pinned memory. */
int dlq_resnet18_submit_host(dlq_resnet18* m, const float* x_host, int N, float* logits_host);
int dlq_resnet18_submit_host_u8(dlq_resnet18* m, const uint8_t* x_hwc_host, int N, float* logits_host);
int dlq_resnet18_wait(dlq_resnet18* m);
/* profile slots of one forward (dlq_resnet18_profile / _read_stamps): 23 = quantise, 20 convs, max-pool, GAP+FC */
int dlq_resnet18_launches(const dlq_resnet18* m);
/* kernels a forward of batch N really launches: 6 from batch 40 on (quantise, stem conv, max-pool, the layer1 chain, the
 * layer2..4 chain, GAP+FC), 20 below (the three 1x1 shortcut convs ride on conv1's launch), 23 from batch 40 on with the
 * chains switched off */
int dlq_resnet18_launches_for_batch(const dlq_resnet18* m, int N);
/* Plan options (A/B measurements and tools; results are bit-identical whatever their values).  Synchronises, drops the
 * cached plans and any captured graph.
 *   "conv_chain" (1)        layer1's four convs and the fifteen convs of layer2..4 as two persistent cooperative launches
 *                           (csrc/conv_chain.cuh) for batches >= "chain_min_batch"; 0: one launch per conv
 *   "chain_min_batch" (40)  smallest batch planned with chains (measured per forward, chains vs separate launches: batch 32
 *                           235 vs 229 us, batch 48 260 vs 276 us)
 *   "chain_layer1" (1), "chain_start" (7) / "chain_first_block"   which convs the chains cover
 *   "chain_launch_mode" (0) 0 cooperative + programmatic stream serialization, 1 cooperative, 2 neither (profilers)
 *   "tile_flags" (0)        also between SEPARATE conv launches: consumers wait for the completion counters of the producer
 *                           units their rows touch instead of for the producer's grid (griddepcontrol.wait) */
int dlq_resnet18_set_option(dlq_resnet18* m, const char* key, int value);
/* "conv_chain" in detail: from batch "chain_min_batch" on, the convs from "chain_start" to the last one run as ONE persistent
 * cooperative kernel (csrc/conv_chain.cuh) whose CTA pairs walk the layers without leaving the SMs - no per-layer
 * hand-over, no partial last waves - with the tile-level dependency flags between them, and layer1's four convs as a
 * second one ("chain_layer1").  "chain_start" is the conv index (DLQ_NUM_CONVS numbering) of a block's conv1 or conv2;
 * default 7 = layer2.0.conv1, i.e. the fifteen convs of layer2..4; "chain_first_block" = b is shorthand for conv1 of block
 * b (1..7).  A chain whose layers do not share the kernel's static tile configuration falls back to one launch per
 * conv.  Bit-identical results. */
/* facts about the launch plan of batch N (planned on demand): "chain_layers" (0 = one launch per conv), "chain_launch_mode",
 * "chain_cta_pairs", "chain_a_stages", "chain_b_stages", "flag_units", "chains" */
int dlq_resnet18_plan_info(dlq_resnet18* m, int N, const char* key, int* value);
/* dependency waits that ran into the 4 s safety timeout since creation (must be 0).  Synchronises. */
int dlq_resnet18_dep_timeouts(dlq_resnet18* m, unsigned int* count);
/* Span stamps (measurement): with a ring of `ring_forwards` entries, every kernel of a forward records the globaltimer
 * (ns) of its first block entry and last block exit - its span INSIDE a running sequence of forwards, overlap with its
 * neighbours included, which an event between launches cannot see.  dlq_resnet18_read_stamps synchronises, copies
 * [ring][dlq_resnet18_launches()][2] = (entry, exit) to HOST memory (entry == UINT64_MAX: slot not written), reports how
 * many forwards were recorded and resets the ring.  ring_forwards = 0 switches the stamps off. */
int dlq_resnet18_enable_stamps(dlq_resnet18* m, int ring_forwards);
int dlq_resnet18_read_stamps(dlq_resnet18* m, unsigned long long* out_host, int* forwards_recorded);
/* one forward with a CUDA event between launches; ms[dlq_resnet18_launches()] receives each launch's device
 * time in order: quantise+s2d, stem conv, max-pool, per block conv1,[downsample],conv2, GAP+FC.  Synchronises.
 * (the reference brackets every launch with its cudaEvent Timer the same way, R/utils.hpp:85-92) */
int dlq_resnet18_profile(dlq_resnet18* m, const float* x, int N, float* logits, float* ms);

/* ------------------------------------------------------------------ weight directory, PTQ calibration, accuracy
 * (the stages either side of the hot path)
 *
 * Weight directory = the reference's export format (T/export_resnet18.py:85-92): one raw little-endian fp32 file
 * per state_dict key, "<key>.bin", names as R/infer_e2e.cu:262-330,428-429 reads them.  dlq_weight_dir_load checks
 * every size like load_bin_f32 (R/utils.hpp:48-60) but reports instead of exiting: returns 1 and fills err.
 * The returned weights point into the object (valid until dlq_weight_dir_free); act_scale is filled when the
 * directory carries "quant.act_scale.int8.bin" (written by dlq_weight_dir_save with with_scales = 1 - the "quant"
 * block RKL/reports/Step1.md:92 plans), else zero: calibrate first. */
typedef struct dlq_weight_dir dlq_weight_dir;
int dlq_weight_dir_load(const char* dir, dlq_weight_dir** out, char* err, size_t err_len);
const dlq_resnet18_weights* dlq_weight_dir_weights(const dlq_weight_dir* d);
void dlq_weight_dir_free(dlq_weight_dir* d);
/* writes every tensor of w as <key>.bin plus manifest.json into an existing directory (HOST pointers) */
int dlq_weight_dir_save(const char* dir, const dlq_resnet18_weights* w, int with_scales);

/* The reference's FP32 network on the GPU, operator for operator and bit for bit (sequential-FMA conv of
 * K/im2col.cu + K/sgemm_tiled.cu, K/bn_inference.cu:22-27, K/relu.cu:9, K/add.cu:7, K/maxpool2d.cu:14-40,
 * K/gap_global.cu:10-32, R/infer_e2e.cu:206-219), for batches.  Two uses: (1) PTQ calibration - every forward
 * folds the absmax of the DLQ_NUM_ACTS activation tensors into a running maximum; (2) the FP32 side of the accuracy
 * harness (checkpoints named as R/infer_e2e.cu:297-426 dumps them).  Not a hot path (CUDA cores, one thread per
 * 8 outputs).  x, logits and checkpoint outputs are DEVICE pointers; act_scale of w is ignored. */
typedef struct dlq_resnet18_f32 dlq_resnet18_f32;
int dlq_resnet18_f32_create(dlq_ctx* ctx, const dlq_resnet18_weights* w, int max_batch, dlq_resnet18_f32** out);
void dlq_resnet18_f32_destroy(dlq_resnet18_f32* m);
int dlq_resnet18_f32_forward(dlq_resnet18_f32* m, const float* x, int N, float* logits);
/* fp32 NCHW copy of a checkpoint of the LAST forward: "stem_pool","layer1".."layer4","gap" */
int dlq_resnet18_f32_checkpoint(dlq_resnet18_f32* m, const char* name, float* out);
/* running absmax since creation / the last reset -> HOST array [DLQ_NUM_ACTS]; synchronises */
int dlq_resnet18_f32_absmax(dlq_resnet18_f32* m, float* absmax);
int dlq_resnet18_f32_reset_absmax(dlq_resnet18_f32* m);
/* spec/QUANT_SPEC.md 2 / 6: act_scale[i] = fp32(absmax[i] / 127) (INT8) or / 448 (E4M3); absmax 0 counts as 1. HOST. */
void dlq_act_scales_from_absmax(const float* absmax, int fp8, float* act_scale);

/* top-k of every row of x[N,K] (k <= 32): idx[N,k] (and val[N,k] unless NULL), ties to the lower index - the
 * arg-max R/infer_e2e.cu:436-438 prints, for batches and k > 1.  DEVICE pointers. */
int dlq_topk_f32(dlq_ctx* ctx, const float* x, int N, int K, int k, int* idx, float* val);
/* out3 (HOST) = { max |a-b|, mean |a-b|, cosine(a,b) } accumulated in double: diff_max_mean (R/utils.hpp:163-177)
 * and T/diag_e2e_compare.py:12-24 in one pass over DEVICE tensors.  Synchronises. */
int dlq_compare_f32(dlq_ctx* ctx, const float* a, const float* b, size_t n, double* out3);

/* deterministic synthetic data: v[i] = (lo + splitmix64(seed,name) % (hi-lo+1)) * 2^-shift. HOST. */
void dlq_synth_fill_f32(float* v, size_t n, uint64_t seed, const char* name, int lo, int hi, int shift);

/* ------------------------------------------------------------------ batch-sharded multi-GPU driver
 * (no reference counterpart: the reference is single-GPU, N=1).  One context + model replica + host
 * thread per device; images split contiguously; logits gathered into one host array. */
typedef struct dlq_multi dlq_multi;
int dlq_multi_create(const int* devices, int n_devices, const dlq_resnet18_weights* w, int max_batch_per_device,
                     dlq_multi** out);
void dlq_multi_destroy(dlq_multi* m);
int dlq_multi_forward_host(dlq_multi* m, const float* x_host, int N, float* logits_host);
const char* dlq_multi_last_error_string(const dlq_multi* m);
int dlq_multi_n_devices(const dlq_multi* m);
/* uint8 HWC input (dlq_resnet18_set_preprocess on every replica, then 150 KB/image over the host link) */
int dlq_multi_set_preprocess(dlq_multi* m, const float* mean3, const float* std3);
int dlq_multi_forward_host_u8(dlq_multi* m, const uint8_t* x_hwc_host, int N, float* logits_host);
/* asynchronous form: submit splits the batch and returns once every worker has its share queued; up to two batches per
 * device are in flight.  dlq_multi_wait blocks until everything submitted so far has its logits in host memory. */
int dlq_multi_submit_host(dlq_multi* m, const float* x_host, int N, float* logits_host);
int dlq_multi_submit_host_u8(dlq_multi* m, const uint8_t* x_hwc_host, int N, float* logits_host);
int dlq_multi_wait(dlq_multi* m);
/* device-resident form: x_dev[g] / logits_dev[g] are DEVICE pointers on the g-th listed device (n_per_dev[g] fp32 NCHW
 * images, n_per_dev[g] x 1000 logits); returns when every replica has finished. */
int dlq_multi_forward_device(dlq_multi* m, const float* const* x_dev, const int* n_per_dev, float* const* logits_dev);

/* ------------------------------------------------------------------ MNIST MLP forward
 * The GPU counterpart of the reference's MLP forward - MN/v4.cu:255-302 forward_timed (matmul -> bias -> relu ->
 * matmul -> bias -> softmax), MN/v5.cu:127-157 forward_pass_only, MN/v3.c:177-215 on the CPU - as INT8 (or E4M3) FC
 * layers on the tensor-core GEMM core:
 *     x_q   = quantise(x, s_x)                                                       QUANT_SPEC 2 / 6
 *     h_q   = clamp(rne(fmaf(acc1, s_x s_w1[h] / s_h, b1[h] / s_h)), 0, 127)           bias + ReLU + requant fused (3)
 *     logit = fmaf(acc2, s_h s_w2[o], b2[o])                                         QUANT_SPEC 5 (FC)
 *     prob  = softmax(logit)                                                         K/softmax.cu form, no clamp
 * w1 / w2 are HOST pointers in the reference's [in, out] layout (matmul_a_b(A[m,n], B[n,k]), MN/v3.c:125-134);
 * s_x / s_h: activation scales (absmax / 127 of a calibration batch, or / 448 for fp8).  hid must be 64 or a
 * multiple of 128 (MNIST: 784 -> 256 -> 10). */
typedef struct dlq_mlp dlq_mlp;
int dlq_mlp_create(dlq_ctx* ctx, const float* w1_in_hid, const float* b1, const float* w2_hid_out, const float* b2, int in,
                   int hid, int out, float s_x, float s_h, int max_batch, int fp8, dlq_mlp** out_m);
void dlq_mlp_destroy(dlq_mlp* m);
/* x: fp32 [B, in] device; logits / probs: fp32 [B, out] device, either may be NULL */
int dlq_mlp_forward(dlq_mlp* m, const float* x, int B, float* logits, float* probs);
/* checkpoints of the LAST forward, DEVICE output: "input" = quantised x [B, in], "hidden" = h_q [B, hid] (bytes) */
int dlq_mlp_checkpoint(dlq_mlp* m, const char* name, uint8_t* out);
/* HOST: the per-row weight scales of layer 1 (hid values) or 2 (out values) */
int dlq_mlp_weight_scales(const dlq_mlp* m, int layer, float* scale_host);

#ifdef __cplusplus
}
#endif
#endif
