/* dlq_oracle.h — CPU oracle for the dlq_b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libdlq_b200.so) never links or calls it.
 *
 * Two halves:
 *   (1) FP32 restatement of the reference's ResNet-18 operator semantics
 *       (CUDA/resnet18-kernel-lab/cpp/fp32/{kernels,runtime}); each function cites the file:line
 *       it follows.  Pinned by the reference's own GAP+FC known-answer vectors
 *       (tests/golden/ref_gapfc_*.bin) and by torchvision (the third-party arithmetic the
 *       reference's tests compare against, atol 1e-4).
 *   (2) INT8 operators per spec/QUANT_SPEC.md.  PARITY UNPINNED: the reference contains no
 *       quantised code, so nothing in the reference constrains these; they are pinned only to the
 *       written spec and, through dequantise + tolerance, to half (1).
 *
 * All tensors are dense NCHW (reference convention), row-major, caller-owned.
 */
#ifndef DLQ_ORACLE_H
#define DLQ_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- synthetic data */
uint64_t orc_splitmix64(uint64_t* state);
uint64_t orc_name_hash(const char* name);                 /* FNV-1a 64 */
/* v[i] = (lo + (splitmix % (hi-lo+1))) * 2^-shift ; stream keyed by (seed, name) */
void orc_fill_lattice_f32(float* v, size_t n, uint64_t seed, const char* name, int lo, int hi, int shift);
void orc_fill_lattice_i8(int8_t* v, size_t n, uint64_t seed, const char* name, int lo, int hi);

/* ---------------------------------------------------------------- FP32 reference semantics */
void orc_conv2d_f32(const float* x, int N, int C, int H, int W, const float* w, int OC, int kH, int kW,
                    int sH, int sW, int pH, int pW, float* y);
void orc_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K);
void orc_bn_inference_f32(float* x, const float* g, const float* b, const float* m, const float* v, float eps,
                          int N, int C, int OH, int OW);
void orc_relu_f32(float* x, size_t n);
void orc_add_f32(float* y, const float* x, size_t n);
void orc_maxpool3x3s2p1_f32(const float* x, int N, int C, int H, int W, float* y);
void orc_gap_f32(const float* x, int N, int C, int H, int W, float* y);
void orc_fc_f32(const float* gap, const float* W, const float* B, float* out, int N, int O, int I);
void orc_softmax_f32(const float* x, int K, float* y);

/* One conv + BN parameter set (reference file naming: <prefix>.weight, bnX.weight/bias/running_*) */
typedef struct {
  int ic, oc, k, stride, pad;
  const float* w;     /* OIHW */
  const float* gamma; /* [oc] */
  const float* beta;
  const float* mean;
  const float* var;
} orc_convbn;

/* ResNet-18 as wired by runtime/infer_e2e.cu:254-433.
 * convs[0] = stem; block b (0..7): convs[1+3b] = conv1, [2+3b] = conv2, [3+3b] = downsample
 * (w == NULL when the block has no downsample). */
#define ORC_NUM_CONVS 25
#define ORC_NUM_BLOCKS 8
typedef struct {
  orc_convbn convs[ORC_NUM_CONVS];
  const float* fc_w; /* [1000,512] */
  const float* fc_b; /* [1000] */
} orc_resnet18_f32;

/* activation tensors whose absmax the calibrator records / whose scale the INT8 path needs */
enum {
  ORC_ACT_INPUT = 0,
  ORC_ACT_STEM = 1,        /* after conv1+bn+relu (max-pool keeps the scale) */
  ORC_ACT_BLOCK0 = 2,      /* per block: +0 conv1 out, +1 downsample out, +2 block out */
  ORC_ACT_GAP = 2 + 3 * ORC_NUM_BLOCKS,
  ORC_NUM_ACTS = 3 + 3 * ORC_NUM_BLOCKS
};

/* Optional checkpoints (NULL to skip), names as runtime/infer_e2e.cu:297-433 dumps them. */
typedef struct {
  float* stem_pool; /* [N,64,56,56]  */
  float* layer1;    /* [N,64,56,56]  */
  float* layer2;    /* [N,128,28,28] */
  float* layer3;    /* [N,256,14,14] */
  float* layer4;    /* [N,512,7,7]   */
  float* gap;       /* [N,512]       */
  float* absmax;    /* [ORC_NUM_ACTS] running max over the batch (caller zero-inits) */
} orc_checkpoints_f32;

void orc_resnet18_f32_forward(const orc_resnet18_f32* m, const float* x, int N, float* logits,
                              orc_checkpoints_f32* ck);

/* ---------------------------------------------------------------- INT8 spec operators */
float orc_inv_scale(float s); /* fp32(1.0 / (double)s) */
void orc_quantize_f32_i8(const float* x, size_t n, float inv_s, int lo, int hi, int8_t* q);
void orc_dequantize_i8_f32(const int8_t* q, size_t n, float s, float* x);
void orc_dequantize_i8_f32_per_channel(const int8_t* q, int N, int C, int HW, const float* s, float* x);
/* per-output-channel symmetric weight quantisation: s[oc] = absmax/127 (1.0 if the row is all zero) */
void orc_quantize_weights_per_channel(const float* w, int OC, int K, int8_t* q, float* s);
/* folded requantisation constants (QUANT_SPEC §3), a = gamma/sqrt(var+eps), all in double, rounded once:
 *   alpha = fp32(s_x*s_w*a/s_y), beta = fp32((beta_bn - mean*a)/s_y);  res_mul = fp32(s_r/s_y) */
void orc_fold_bn(const float* g, const float* b, const float* m, const float* v, float eps, const float* s_w,
                 float s_x, float s_y, int OC, float* alpha, float* beta);
float orc_res_mul(float s_r, float s_y);

typedef struct {
  const float* alpha;     /* [OC] */
  const float* beta;      /* [OC] */
  const int8_t* residual; /* NCHW int8 or NULL */
  float res_mul;
  int relu;
} orc_epilogue;

/* int8 conv; acc_out (int32 NCHW, may be NULL) receives raw accumulators, y (int8 NCHW) the epilogue result */
void orc_conv2d_i8(const int8_t* x, int N, int C, int H, int W, const int8_t* w, int OC, int kH, int kW, int sH,
                   int sW, int pH, int pW, const orc_epilogue* ep, int32_t* acc_out, int8_t* y);
void orc_add_requant_i8(int8_t* y, float y_scale, const int8_t* x, float x_scale, size_t n, int relu, float out_scale);
void orc_maxpool3x3s2p1_i8(const int8_t* x, int N, int C, int H, int W, int8_t* y);
/* GAP: int32 sum -> (float)sum * scale_over_hw -> optional requant */
void orc_gap_i8(const int8_t* x, int N, int C, int H, int W, float scale_over_hw, float inv_out_scale,
                float* y_f32, int8_t* y_i8);
void orc_fc_i8(const int8_t* g, const int8_t* w, const float* w_scale_times_g, const float* bias, int N, int O,
               int I, int32_t* acc_out, float* logits);

typedef struct {
  int ic, oc, k, stride, pad;
  const int8_t* w; /* OIHW int8 */
  const float* alpha;
  const float* beta;
} orc_convq;

typedef struct {
  orc_convq convs[ORC_NUM_CONVS];
  float act_scale[ORC_NUM_ACTS];
  const int8_t* fc_w;     /* [1000,512] */
  const float* fc_scale;  /* [1000] = fp32(s_gap * s_w[o]) */
  const float* fc_b;      /* [1000] */
} orc_resnet18_i8;

typedef struct {
  int8_t* stem_pool;
  int8_t* layer1;
  int8_t* layer2;
  int8_t* layer3;
  int8_t* layer4;
  int8_t* gap;
} orc_checkpoints_i8;

void orc_resnet18_i8_forward(const orc_resnet18_i8* m, const float* x, int N, float* logits,
                             orc_checkpoints_i8* ck);

/* ---------------------------------------------------------------- FP8 (E4M3) path, spec/QUANT_SPEC.md section 6
 * PARITY UNPINNED by the reference (it has no quantised code).  E4M3 = 1-4-3, bias 7, max 448, 0x7F/0xFF = NaN;
 * float -> E4M3 rounds to nearest even and saturates to +-448.  Products are accumulated in DOUBLE here (the
 * tensor core's fp32 accumulation order is unspecified), so GPU parity is by tolerance. */
uint8_t orc_f32_to_e4m3(float f);
float orc_e4m3_to_f32(uint8_t b);
void orc_quantize_f32_e4m3(const float* x, size_t n, float inv_s, uint8_t* q);
void orc_dequantize_e4m3_f32(const uint8_t* q, size_t n, float s, float* x);
void orc_quantize_weights_per_channel_e4m3(const float* w, int OC, int K, uint8_t* q, float* s);
/* residual in ep is read as E4M3 bytes; acc_out = (float)double_accumulator */
void orc_conv2d_e4m3(const uint8_t* x, int N, int C, int H, int W, const uint8_t* w, int OC, int kH, int kW, int sH,
                     int sW, int pH, int pW, const orc_epilogue* ep, float* acc_out, uint8_t* y);
void orc_gap_e4m3(const uint8_t* x, int N, int C, int H, int W, float scale_over_hw, float inv_out_scale, uint8_t* y);
void orc_fc_e4m3(const uint8_t* g, const uint8_t* w, const float* w_scale_times_g, const float* bias, int N, int O,
                 int I, float* logits);
/* same model struct as the INT8 network: weight bytes are E4M3 codes, act_scale maps absmax to 448 */
void orc_resnet18_fp8_forward(const orc_resnet18_i8* m, const float* x, int N, float* logits, orc_checkpoints_i8* ck);

/* ---------------------------------------------------------------- MNIST MLP forward (config #1) */
/* restatement of CUDA/MNIST_on_GPU/v3.c:108-215 (forward only) */
void orc_mnist_mlp_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                           int batch, int in_dim, int hid, int out_dim, float* hidden, float* out);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
