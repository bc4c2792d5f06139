// step8_e2e_dlq — the reference's whole-network driver (runtime/infer_e2e.cu:230-441, built as `step8_e2e`) on top of
// libdlq_b200.so: same command line, same weight directory, same input file, same checkpoint dumps, same
// "[E2E] top-1 class index = ..." line (so tools/bench_fp32_vs_torch_e2e.py:30,108-118 parses it unchanged), but the
// network runs as the INT8 (or FP8) tensor-core path, for every image in the input file at once.
//
//   step8_e2e_dlq --manifest <weight dir> --input <N x 3x224x224 fp32 .bin> [--dump_dir D]
//                 [--calib <M x 3x224x224 fp32 .bin>]   calibration images (default: the input itself) when the
//                                                       directory carries no quant.act_scale.int8.bin
//                 [--save_scales]                       write the calibrated scales back into the directory
//                 [--fp8]                               E4M3 network instead of INT8
//                 [--fp32]                              the reference's FP32 arithmetic (dlq_resnet18_f32_*)
//                 [--compare]                           run FP32 too and print per-checkpoint max_abs / mean_abs /
//                                                       cosine (tools/diag_e2e_compare.py) and top-1 agreement
//                 [--expect_dir E] [--atol 1e-4]        parity-test mode, like the reference's per-step drivers
//                                                       (runtime/infer_layer1.cu:243,287, infer_head.cu:125-132):
//                                                       compare every checkpoint this run dumped with E/<name>.bin
//                                                       (a --dump_dir of the reference binary or of another run),
//                                                       criterion max_abs / max(1, max|expected|) <= atol; exit 2 on
//                                                       mismatch
// Exit codes follow the reference (runtime/utils.hpp:23-45, infer_conv1_bn1_relu.cu:150-156): 1 = bad usage / IO,
// 2 = parity mismatch (--expect_dir), 3 = CUDA error.
// Host code only: plain C++ against include/dlq.h and the CUDA runtime (device buffers); no kernels here.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../include/dlq.h"

#define CK_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      fprintf(stderr, "CUDA error %s @ %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      return 3;                                                                                         \
    }                                                                                                   \
  } while (0)
#define CK_DLQ(ctx, call)                                                                               \
  do {                                                                                                  \
    int rc_ = (call);                                                                                   \
    if (rc_ != DLQ_OK) {                                                                                \
      fprintf(stderr, "dlq error %d: %s @ %s:%d\n", rc_, dlq_last_error_string(ctx), __FILE__, __LINE__); \
      return rc_;                                                                                       \
    }                                                                                                   \
  } while (0)

static const size_t kImg = 3 * 224 * 224;
struct Ck { const char* name; size_t elems; int act; };
// checkpoint -> activation-scale index (include/dlq.h DLQ_NUM_ACTS): stem = 1, block b output = 4 + 3 b, gap = 26
static const Ck kCk[6] = {{"stem_pool", 64 * 56 * 56, 1},  {"layer1", 64 * 56 * 56, 4 + 3 * 1}, {"layer2", 128 * 28 * 28, 4 + 3 * 3},
                          {"layer3", 256 * 14 * 14, 4 + 3 * 5}, {"layer4", 512 * 7 * 7, 4 + 3 * 7}, {"gap", 512, 26}};

static bool read_f32(const std::string& path, std::vector<float>& v, size_t multiple_of) {
  std::ifstream ifs(path, std::ios::binary);
  if (!ifs) { std::cerr << "open fail: " << path << "\n"; return false; }
  ifs.seekg(0, std::ios::end);
  const size_t bytes = static_cast<size_t>(ifs.tellg());
  ifs.seekg(0);
  if (bytes == 0 || bytes % (4 * multiple_of)) {
    std::cerr << "unexpected size: " << path << " got " << bytes / 4 << " floats, expected a multiple of " << multiple_of << "\n";
    return false;
  }
  v.resize(bytes / 4);
  ifs.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(bytes));
  return static_cast<bool>(ifs);
}
static bool write_f32(const std::string& path, const std::vector<float>& v) {
  std::ofstream ofs(path, std::ios::binary);
  if (!ofs) { std::cerr << "open fail (write): " << path << "\n"; return false; }
  ofs.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * 4));
  return static_cast<bool>(ofs);
}

static void usage() {
  std::cerr << "usage: step8_e2e_dlq --manifest <weight dir> --input <input.bin> [--dump_dir D] [--calib calib.bin]\n"
               "                     [--save_scales] [--fp8] [--fp32] [--compare] [--expect_dir E [--atol 1e-4]]\n";
}

int main(int argc, char** argv) {
  std::string mani, input_path, dump_dir, calib_path, expect_dir;
  double atol = 1e-4;
  bool fp8 = false, fp32 = false, compare = false, save_scales = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--manifest" && i + 1 < argc) mani = argv[++i];
    else if (a == "--input" && i + 1 < argc) input_path = argv[++i];
    else if (a == "--dump_dir" && i + 1 < argc) dump_dir = argv[++i];
    else if (a == "--calib" && i + 1 < argc) calib_path = argv[++i];
    else if (a == "--expect_dir" && i + 1 < argc) expect_dir = argv[++i];
    else if (a == "--atol" && i + 1 < argc) atol = std::atof(argv[++i]);
    else if (a == "--fp8") fp8 = true;
    else if (a == "--fp32") fp32 = true;
    else if (a == "--compare") compare = true;
    else if (a == "--save_scales") save_scales = true;
  }
  if (mani.empty() || input_path.empty()) { usage(); return 1; }
  if (!expect_dir.empty() && dump_dir.empty()) { std::cerr << "--expect_dir compares the files of --dump_dir: give both\n"; return 1; }
  // the reference takes the directory; accept a path to its manifest.json as well
  if (mani.size() > 13 && mani.substr(mani.size() - 13) == "manifest.json") mani = mani.substr(0, mani.size() - 14);

  std::vector<float> X;
  if (!read_f32(input_path, X, kImg)) return 1;
  const int N = static_cast<int>(X.size() / kImg);

  char err[512];
  dlq_weight_dir* wd = nullptr;
  if (dlq_weight_dir_load(mani.c_str(), &wd, err, sizeof err) != DLQ_OK) { std::cerr << err << "\n"; return 1; }
  dlq_resnet18_weights W = *dlq_weight_dir_weights(wd);

  dlq_ctx* ctx = nullptr;
  if (dlq_create(0, &ctx) != DLQ_OK) { std::cerr << "dlq_create failed (an sm_100 GPU is required; there is no CPU path)\n"; return 3; }

  float *dX = nullptr, *dLogits = nullptr, *dLogitsF = nullptr, *dCkF = nullptr, *dCkQ = nullptr;
  int8_t* dCk8 = nullptr;
  CK_CUDA(cudaMalloc(&dX, X.size() * 4));
  CK_CUDA(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK_CUDA(cudaMalloc(&dLogits, static_cast<size_t>(N) * 1000 * 4));
  CK_CUDA(cudaMalloc(&dLogitsF, static_cast<size_t>(N) * 1000 * 4));
  const size_t ck_max = static_cast<size_t>(N) * 64 * 56 * 56;
  CK_CUDA(cudaMalloc(&dCkF, ck_max * 4));
  CK_CUDA(cudaMalloc(&dCkQ, ck_max * 4));
  CK_CUDA(cudaMalloc(&dCk8, ck_max));

  // ---- FP32 network (reference arithmetic): calibration, --fp32 and --compare
  bool have_scales = false;
  for (int i = 0; i < DLQ_NUM_ACTS; ++i) have_scales |= W.act_scale[i] != 0.f;
  dlq_resnet18_f32* f32net = nullptr;
  const int chunk = N < 32 ? N : 32;
  if (fp32 || compare || !have_scales || fp8) {
    std::vector<float> Cal;
    if (!calib_path.empty() && !read_f32(calib_path, Cal, kImg)) return 1;
    const int M = static_cast<int>(Cal.size() / kImg);
    const int cap = chunk > 32 ? chunk : 32;
    CK_DLQ(ctx, dlq_resnet18_f32_create(ctx, &W, cap, &f32net));
    if (!have_scales || fp8) {
      // PTQ: run the calibration images through the FP32 network; scales = absmax / 127 (or / 448)
      float* dC = dX;
      int n_cal = N;
      if (M) {
        CK_CUDA(cudaMalloc(&dC, Cal.size() * 4));
        CK_CUDA(cudaMemcpy(dC, Cal.data(), Cal.size() * 4, cudaMemcpyHostToDevice));
        n_cal = M;
      }
      float* dTmp = nullptr;
      CK_CUDA(cudaMalloc(&dTmp, static_cast<size_t>(cap) * 1000 * 4));
      for (int i = 0; i < n_cal; i += cap)
        CK_DLQ(ctx, dlq_resnet18_f32_forward(f32net, dC + static_cast<size_t>(i) * kImg, (n_cal - i) < cap ? (n_cal - i) : cap, dTmp));
      float am[DLQ_NUM_ACTS];
      CK_DLQ(ctx, dlq_resnet18_f32_absmax(f32net, am));
      dlq_act_scales_from_absmax(am, fp8 ? 1 : 0, W.act_scale);
      std::cout << "[PTQ] calibrated " << DLQ_NUM_ACTS << " activation scales on " << n_cal << " image(s): input "
                << W.act_scale[0] << ", stem " << W.act_scale[1] << ", gap " << W.act_scale[DLQ_NUM_ACTS - 1] << "\n";
      if (save_scales && !fp8) {
        if (dlq_weight_dir_save(mani.c_str(), &W, 1) != DLQ_OK) { std::cerr << "cannot write scales into " << mani << "\n"; return 1; }
      }
      cudaFree(dTmp);
      if (M) cudaFree(dC);
    }
  }
  W.fp8 = fp8 ? 1 : 0;

  auto maybe_save = [&](const std::string& name, const float* dev, size_t n) -> int {
    if (dump_dir.empty()) return 0;
    std::vector<float> h(n);
    CK_CUDA(cudaMemcpy(h.data(), dev, n * 4, cudaMemcpyDeviceToHost));
    return write_f32(dump_dir + "/" + name, h) ? 0 : 1;
  };
  if (!dump_dir.empty()) { const std::string cmd = "mkdir -p '" + dump_dir + "'"; if (std::system(cmd.c_str()) != 0) return 1; }

  std::vector<float> logits(static_cast<size_t>(N) * 1000);
  if (fp32) {
    // ---- the reference's arithmetic, batched
    for (int i = 0; i < N; i += 32) {
      const int n = (N - i) < 32 ? (N - i) : 32;
      CK_DLQ(ctx, dlq_resnet18_f32_forward(f32net, dX + static_cast<size_t>(i) * kImg, n, dLogits + static_cast<size_t>(i) * 1000));
      if (N <= 32)
        for (const Ck& c : kCk) {
          CK_DLQ(ctx, dlq_resnet18_f32_checkpoint(f32net, c.name, dCkF));
          CK_DLQ(ctx, dlq_sync(ctx));
          if (int rc = maybe_save(std::string(c.name) + ".bin", dCkF, static_cast<size_t>(n) * c.elems)) return rc;
        }
    }
  } else {
    dlq_resnet18* net = nullptr;
    CK_DLQ(ctx, dlq_resnet18_create(ctx, &W, N, &net));
    CK_DLQ(ctx, dlq_resnet18_forward(net, dX, N, dLogits));
    const bool want_ck = !dump_dir.empty() || compare;
    if (want_ck && !fp8) {
      if (compare) {
        if (N > 32) { std::cerr << "--compare handles at most 32 images per run\n"; return 1; }
        CK_DLQ(ctx, dlq_resnet18_f32_forward(f32net, dX, N, dLogitsF));
        std::cout << "[COMPARE] FP32 (reference arithmetic) vs " << (fp8 ? "FP8" : "INT8") << ", " << N << " image(s)\n";
      }
      for (const Ck& c : kCk) {
        const size_t n = static_cast<size_t>(N) * c.elems;
        CK_DLQ(ctx, dlq_resnet18_checkpoint(net, c.name, dCk8));
        CK_DLQ(ctx, dlq_dequantize_i8_f32(ctx, dCk8, n, W.act_scale[c.act], dCkQ));
        CK_DLQ(ctx, dlq_sync(ctx));
        if (int rc = maybe_save(std::string(c.name) + ".bin", dCkQ, n)) return rc;
        if (compare) {
          double r[3];
          CK_DLQ(ctx, dlq_resnet18_f32_checkpoint(f32net, c.name, dCkF));
          CK_DLQ(ctx, dlq_compare_f32(ctx, dCkF, dCkQ, n, r));
          printf("%-14s  max_abs=%.6g  mean_abs=%.6g  cosine=%.6f\n", (std::string(c.name) + ".bin").c_str(), r[0], r[1], r[2]);
        }
      }
    }
    if (compare) {
      if (fp8) CK_DLQ(ctx, dlq_resnet18_f32_forward(f32net, dX, N, dLogitsF));
      double r[3];
      CK_DLQ(ctx, dlq_compare_f32(ctx, dLogitsF, dLogits, static_cast<size_t>(N) * 1000, r));
      printf("%-14s  max_abs=%.6g  mean_abs=%.6g  cosine=%.6f\n", "logits.bin", r[0], r[1], r[2]);
      int *dT = nullptr, *dTF = nullptr;
      CK_CUDA(cudaMalloc(&dT, N * sizeof(int)));
      CK_CUDA(cudaMalloc(&dTF, N * sizeof(int)));
      CK_DLQ(ctx, dlq_topk_f32(ctx, dLogits, N, 1000, 1, dT, nullptr));
      CK_DLQ(ctx, dlq_topk_f32(ctx, dLogitsF, N, 1000, 1, dTF, nullptr));
      CK_DLQ(ctx, dlq_sync(ctx));
      std::vector<int> t(N), tf(N);
      CK_CUDA(cudaMemcpy(t.data(), dT, N * sizeof(int), cudaMemcpyDeviceToHost));
      CK_CUDA(cudaMemcpy(tf.data(), dTF, N * sizeof(int), cudaMemcpyDeviceToHost));
      int agree = 0;
      for (int i = 0; i < N; ++i) agree += t[i] == tf[i];
      printf("agree_top1=%d (%.2f%%)\n", agree, 100.0 * agree / N);    // tools/bench_fp32_vs_torch_e2e.py:127
      cudaFree(dT); cudaFree(dTF);
    }
    CK_DLQ(ctx, dlq_sync(ctx));
    dlq_resnet18_destroy(net);
  }
  CK_DLQ(ctx, dlq_sync(ctx));
  CK_CUDA(cudaMemcpy(logits.data(), dLogits, logits.size() * 4, cudaMemcpyDeviceToHost));
  if (!dump_dir.empty() && !write_f32(dump_dir + "/logits.bin", logits)) return 1;

  // top-1 per image, the reference's scan and output line (runtime/infer_e2e.cu:436-439)
  for (int n = 0; n < N; ++n) {
    int top = -1;
    float best = -1e30f;
    for (int i = 0; i < 1000; ++i)
      if (logits[static_cast<size_t>(n) * 1000 + i] > best) { best = logits[static_cast<size_t>(n) * 1000 + i]; top = i; }
    std::cout << "[E2E] top-1 class index = " << top << ", logit=" << best << "\n";
  }

  // ---- parity-test mode: this run's dumps against an expected dump directory (host-side, diff_max_mean of
  // runtime/utils.hpp:163-177)
  int mismatches = 0;
  if (!expect_dir.empty()) {
    const char* names[7] = {"stem_pool", "layer1", "layer2", "layer3", "layer4", "gap", "logits"};
    for (const char* nm : names) {
      std::vector<float> got, want;
      std::ifstream probe(dump_dir + "/" + nm + ".bin", std::ios::binary);
      if (!probe) continue;                                   // (FP8 runs dump logits only)
      probe.close();
      if (!read_f32(dump_dir + "/" + nm + ".bin", got, 1) || !read_f32(expect_dir + "/" + nm + ".bin", want, 1)) return 1;
      if (got.size() != want.size()) {
        std::cerr << "unexpected size: " << nm << ".bin got " << got.size() << " expected " << want.size() << "\n";
        return 1;
      }
      double max_abs = 0, mean_abs = 0, ref_max = 0;
      for (size_t i = 0; i < got.size(); ++i) {
        const double d = std::abs(static_cast<double>(got[i]) - static_cast<double>(want[i]));
        if (d > max_abs) max_abs = d;
        mean_abs += d;
        if (std::abs(static_cast<double>(want[i])) > ref_max) ref_max = std::abs(static_cast<double>(want[i]));
      }
      mean_abs /= static_cast<double>(got.size());
      const bool ok = max_abs / (ref_max > 1.0 ? ref_max : 1.0) <= atol;
      printf("%s %-10s max_abs=%.6g mean_abs=%.6g (atol %.3g, scale %.4g)\n", ok ? "[OK]  " : "[FAIL]", nm, max_abs, mean_abs,
             atol, ref_max > 1.0 ? ref_max : 1.0);
      mismatches += ok ? 0 : 1;
    }
  }

  if (f32net) dlq_resnet18_f32_destroy(f32net);
  cudaFree(dX); cudaFree(dLogits); cudaFree(dLogitsF); cudaFree(dCkF); cudaFree(dCkQ); cudaFree(dCk8);
  dlq_destroy(ctx);
  dlq_weight_dir_free(wd);
  return mismatches ? 2 : 0;
}
