// The opaque handle behind the C ABI (include/kfnet_b200.h): loaded weights, packed operands, cached launch plans, and --
// when kfn_train_begin was called -- the training state (train.cu).
#pragma once
#include <array>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "conv_tc.h"

namespace kfn {
// One packed GEMM operand: a conv, one phase of a conv2d_transpose, a dense layer, or the im2col'ed first layer.
struct PackedOp {
  void* w = nullptr;       // fp32 [ntaps][Cin][CoutP] (FP32 mode) or bf16 [ntaps][CoutP][Cin] (BF16 mode)
  float* bias = nullptr;   // fp32 [CoutP], zero padded
  int Cin = 0, Cout = 0, CoutP = 0;
  int ntaps = 0, dy[9] = {0}, dx[9] = {0};
  int stride = 1, relu = 0;
  int phase_cout = 0;      // merged conv2d_transpose: Cout per phase (op.Cout = 4 * phase_cout)
  // taps that can never reach an in-range pixel were dropped for the geometry the op was packed for (m-grid pMH x pMW over
  // an input of pIH x pIW); `pruned` ops may only run on exactly that geometry
  int pMH = 0, pMW = 0, pIH = 0, pIW = 0, pruned = 0;
};


struct TrainState;                       // train.cu
void train_state_free(TrainState* t);    // train.cu
}  // namespace kfn

struct kfn_handle {
  int device = 0, mode = 0, H = 0, W = 0, GH = 0, GW = 0, num_sms = 148;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, kfn::PackedOp> ops;
  std::vector<void*> allocs;
  // launch plans (tensor maps + kernel parameters) keyed by op, buffers AND the full geometry: kfn_conv_layer may be
  // called with different shapes / channel offsets / output types on recycled device pointers
  typedef std::tuple<std::string, const void*, void*, void*, std::array<int, 18>> PlanKey;
  std::map<PlanKey, kfn::TcConvPlan> plans;
  std::map<std::tuple<const void*, void*, void*, int, int, int>, kfn::Conv1abPlan> fused_plans;   // (img, out, feat1, N, H, W)
  bool finalized = false;
  int64_t launches = 0;          // kernels launched through this handle's entry points (kfn_launch_count)
  // kfn_eval_sequences_host: internal copy stream and per-batch events (library-owned)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  cudaEvent_t compute_done = nullptr;
  kfn::TrainState* train = nullptr;   // kfn_train_begin (train.cu)
};
