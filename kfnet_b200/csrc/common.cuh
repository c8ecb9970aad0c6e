// Shared host/device declarations for libkfnet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include <nvtx3/nvToolsExt.h>      // header-only; ranges are no-ops unless a profiler attached its injection library

#include "../../include/kfnet_b200.h"

namespace kfn {

void set_error(const std::string& msg);
extern thread_local int64_t g_launches;   // kernels launched by this thread through the library (all handles + handle-less calls)
extern thread_local int64_t* g_handle_launches;   // counter of the handle whose entry point is running on this thread, or null

#define KFN_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      kfn::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
      return KFN_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define KFN_LAUNCH_OK()                                                                            \
  do {                                                                                             \
    ++kfn::g_launches;                                                                             \
    if (kfn::g_handle_launches) ++*kfn::g_handle_launches;                                         \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) {                                                                       \
      kfn::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));                     \
      return KFN_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define KFN_REQUIRE(cond, msg)                                                                     \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      kfn::set_error(std::string("argument check failed: ") + #cond + " -- " + msg);               \
      return KFN_ERR_ARG;                                                                          \
    }                                                                                              \
  } while (0)

// NVTX range over a stage of the path (shows up in Nsight Systems / ncu --nvtx): scoped push/pop
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Handle-less entry points run on the device that owns their buffers (not on whatever device happens to be current):
// makes `dev` the current device and returns it through *dev.  SM counts are cached per device.
int use_device_of(const void* device_ptr, int* dev);
int sm_count(int dev);

// ---------------------------------------------------------------------------------------------
// A convolution seen as an implicit GEMM over an "m-grid" of N x MH x MW rows.
//   input pixel of tap t for row (n,my,mx):  (iy, ix) = (my*istride + dy[t], mx*istride + dx[t]);  zero outside
//   output pixel:                            (oy, ox) = (my*omul + oaddy,  mx*omul + oaddx)
// Normal SAME conv: m-grid = output grid, omul=1.  conv2d_transpose phase: m-grid = input grid, omul=2,
// oadd = phase parity.  Channel stride / offset on the output implement concat-by-construction.
struct ConvGeom {
  int N, MH, MW;
  int IH, IW, in_cs, Cin;
  int istride;
  int ntaps;
  int dy[9], dx[9];
  int OH, OW, out_cs, out_coff, Cout;
  int omul, oaddy, oaddx;
  int relu;
  int epi;            // EPI_*
  // error-compensated bf16x3 mode (0 is read as 1): activations / weights are stored as bf16 "limbs"
  // x = x1 + x2 + x3 laid out [limb][channels] inside each pixel (limb stride in_ls / out_ls elements);
  // the GEMM accumulates the limb products a1w1 + a1w2 + a2w1 + a2w2 + a1w3 + a3w1 in fp32.
  int limbs_a, limbs_w, limbs_out;
  int in_ls, out_ls;
  // KFN_MODE_FP16X2: limbs are fp16, x = x1 + 2^-11 * x2' (two limbs; products a1w1 + 2^-11 (a1w2' + a2'w1))
  int half_limbs;
  // merged conv2d_transpose: output columns [ph*phase_cout, (ph+1)*phase_cout) belong to sub-pixel phase ph = 2*py+px and
  // are written to pixel (my*2 + py, mx*2 + px), channels [out_coff, out_coff + phase_cout).  0 = plain layer.
  int phase_cout;
  // dual output (tensor-core path): columns >= split_col go to a second NHWC tensor `out2` (channel stride out2_cs,
  // channels from 0) at the same pixel -- lets two layers that read the same input share one GEMM (conv1a + feat1).
  int split_col, out2_cs;
  // fused first layer (tensor-core path, BF16 mode): `in` is the u8 NHWC image [N, IH, IW, 3]; the K = 32 operand row
  // (27 window values - 128, zero padded) is built inside the kernel.  Cin = 32, one tap, stride 1.
  int im2col_u8;
  // fused difference volume (tensor-core path, BF16 mode, U-Net conv0 on 8x8 windows): `in` is f_prev and `diff_cur`
  // f_cur, both fp32 [pairs, dGH, dGW, 32]; m-grid row n of the conv = label pixel n over all pairs.
  const float* diff_cur;
  int dGH, dGW;
  // fp32 CUDA-core path only (training, train.cu): add the result to what `out` already holds instead of overwriting it
  // (a gradient that has two consumers, e.g. the U-Net's skip connections)
  int accum;
};
enum { EPI_NONE = 0, EPI_EXP_CH3 = 1, EPI_L2NORM = 2 };

// fp32 CUDA-core implicit GEMM (conv_simt.cu).  weights: [ntaps][Cin][CoutP] fp32, CoutP = roundup(Cout,4).
int launch_conv_simt(const ConvGeom& g, const float* in, const float* w, int CoutP, const float* bias, float* out,
                     cudaStream_t st);

// pointwise / gather kernels (pointwise.cu)
int launch_preprocess_f32(const uint8_t* img, float* out, int64_t n, cudaStream_t st);
int launch_im2col27_bf16(const uint8_t* img, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st, int f16 = 0);
int launch_l2norm32(float* x, int64_t rows, cudaStream_t st);
int launch_diff_volume(const float* f_prev, const float* f_cur, int NP, int GH, int GW, void* diff, int out_bf16,
                       cudaStream_t st);   // out_bf16: 0 fp32, 1 bf16, 3 three bf16 limbs [.., 3, 32], 2 two fp16 limbs [.., 2, 32] (second scaled by 2^11)
int launch_softmax_flow(const float* logits, const float* logsig, int logsig_stride, int64_t rows, float* flow,
                        float* sigma_w, int32_t* argmax, cudaStream_t st);
int launch_kalman_scan(const float* measure, const float* flow, const float* sigma_w, const float* state_in,
                       float* state_out, const float* T16_host, int S, int T, int GH, int GW, int first_index,
                       int reinit_period, int nis_gate, float* out_map, float* temp, float* kf, float* nis,
                       int32_t* taps, cudaStream_t st);
int launch_bilinear_sampler(const float* imgs, const float* coords, float* out, int N, int H, int W, int C, int Ht, int Wt,
                            cudaStream_t st);
int launch_apply_transform(const float* in, float* out, int64_t npix, const float* T16_host, cudaStream_t st);
int launch_kalman_update(const float* prior, const float* meas, int64_t npix, float* post, float* nis, cudaStream_t st);

}  // namespace kfn
