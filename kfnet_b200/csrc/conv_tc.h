// Host interface of the tcgen05/TMA implicit-GEMM convolution (conv_tc.cu).
#pragma once
#include "common.cuh"

namespace kfn {

// Opaque launch plan: two CUtensorMaps (activation view, packed weights) + kernel parameters.
struct TcConvPlan {
  alignas(64) uint8_t blob[768];
  int grid;
  int ctas_per_sm;
  int threads;                    // 320, 192 (three CTAs per SM) or 576 (WIDE)
  int lean;                       // plain bf16 layer: co-resident plans use the lean kernel body
  int pair;                       // CTA-pair plan: conv_tc_kernel_pair, launched as clusters of 2
  size_t smem;
};

bool tc_halo_enabled();           // KFN_TC_HALO=0 disables the halo-tile variant (A/B measurements)
bool tc_prefetch_enabled();       // KFN_TC_PREFETCH=1 enables the producer TMA L2 prefetch (measured: no gain, off by default)
bool tc_two_ctas_enabled();       // KFN_TC_TWO_CTAS=0: always one CTA per SM (A/B measurements)
bool tc_three_ctas_enabled();     // KFN_TC_THREE_CTAS=0: at most two CTAs per SM (A/B measurements)
bool tc_wide_enabled();           // KFN_TC_WIDE: 576-thread four-group epilogue variant for short-tile one-CTA layers
bool tc_lean_x2_enabled();        // KFN_TC_LEAN_X2: co-resident CTAs of plain bf16 layers run the lean kernel body
bool tc_diff_fused_enabled();      // KFN_TC_DIFF_FUSED=1 (opt-in, not yet validated): U-Net conv0 forms the difference volume in its producer warps
bool tc_im2col_fused_enabled();   // KFN_TC_IM2COL_FUSED: BF16 first layers build their im2col rows from the u8 image in-kernel
bool tc_group2_enabled();         // KFN_TC_GROUP2: two-CTA plans alternate tiles between epilogue halves instead of splitting columns
bool tc_pdl_enabled();            // KFN_TC_PDL=0: plain stream-ordered launches (A/B measurements)
int tc_pick_kc(int Cin);          // channels per k-block: 64 / 32 / 16 (128B / 64B / 32B swizzle)
int tc_cout_padded(int Cout);     // weights are packed [tap][CoutP][Cin] bf16 with CoutP = roundup(Cout, 16)

// in: bf16 NHWC activations; w_packed: bf16 [ntaps][CoutP][Cin]; bias_padded: fp32 [CoutP]; out: bf16 or fp32.
int prepare_conv_tc(const ConvGeom& g, const void* in, const void* w_packed, int CoutP, const float* bias_padded,
                    void* out, int out_bf16, int num_sms, TcConvPlan* plan, void* out2 = nullptr);
int launch_conv_tc(const TcConvPlan& plan, cudaStream_t st);

// conv1a (+ tower feat1) and conv1b of SCoordNet as one kernel (conv1ab_fused.cu; KFN_TC_FUSE_CONV1=1, not yet validated
// on hardware).  img: u8 [N,H,W,3]; w1_packed: bf16 [80][32] ("first#both"), bias1 [80]; w2_packed: bf16 [9][64][64]
// ("ScoreNet/conv1b"), bias2 [64]; out: bf16 [N,H,W,64] = conv1b's output; feat1: bf16 [N,H,W,16] or null.
struct Conv1abPlan {
  alignas(64) uint8_t blob[768];
  int grid;
  size_t smem;
};
bool tc_fuse_conv1_enabled();
bool conv1ab_shape_ok(int H, int W);
int prepare_conv1ab(const uint8_t* img, int N, int H, int W, const void* w1_packed, const float* bias1, const void* w2_packed,
                    const float* bias2, void* out, void* feat1, int num_sms, Conv1abPlan* plan);
int launch_conv1ab(const Conv1abPlan& plan, cudaStream_t st);

}  // namespace kfn
