/* kfnet_b200 -- C ABI of the B200-native KFNet hot path (libkfnet_b200.so).
 *
 * The reference (zlthinker/KFNet, TF1/Python) has no FFI of its own: its "operator interface" for
 * this path is the Python call surface of KFNet/KFNet.py + KFNet/eval.py.  Each entry point below
 * names the reference function(s) it replaces (paths relative to the reference root); the Python
 * host layer (kfnet_b200/) binds them with ctypes and mirrors the reference's names on top.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.  Every call returns 0 on success or
 *    a negative kfn_status; kfn_last_error() gives the thread-local message.  Nothing throws.
 *  - All tensors are dense NHWC.  Device pointers are caller-owned (torch-allocated); the library
 *    owns only the opaque handle (packed weights, TMA descriptors, scratch it allocates at finalize).
 *  - Calls enqueue work on `stream` (a cudaStream_t passed as void*) and return; no host sync,
 *    except kfn_eval_sequences_host, which takes HOST buffers and synchronises before returning.
 *  - One handle per device / rank; a handle is not thread-safe.
 *  - There is NO CPU fallback: without a CUDA device kfn_create fails with KFN_ERR_CUDA.
 */
#ifndef KFNET_B200_H_
#define KFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define KFN_API __attribute__((visibility("default")))
#else
#define KFN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kfn_handle kfn_handle;

enum kfn_status {
  KFN_OK = 0,
  KFN_ERR_ARG = -1,       /* bad argument / shape / missing weight */
  KFN_ERR_CUDA = -2,      /* CUDA runtime or driver error (message has the cudaError string) */
  KFN_ERR_STATE = -3,     /* call order (e.g. forward before kfn_finalize_weights) */
  KFN_ERR_WORKSPACE = -4  /* workspace too small */
};

/* Compute modes.
 *  KFN_MODE_FP32 : fp32 CUDA-core implicit-GEMM convolutions, fp32 activations.  Numerics contract of the
 *                  reference (all reference arithmetic is fp32); used for the 1e-4 parity configs.
 *  KFN_MODE_BF16 : tcgen05.mma (kind::f16, bf16 operands, fp32 TMEM accumulators) fed by TMA; bf16
 *                  activations between layers; fp32 for everything after the conv stacks.  The
 *                  "bf16 compute / fp32 accumulate" configuration of BASELINE.json configs[3..4].
 *  KFN_MODE_BF16X3: the same tcgen05 kernel, error-compensated: activations and weights are carried as three bf16
 *                  limbs (x = x1+x2+x3, 24 significant bits) and every layer accumulates the six limb products
 *                  a1w1+a1w2+a2w1+a2w2+a1w3+a3w1 in fp32 TMEM -- fp32-grade results (the 1e-4 parity configs) on
 *                  the tensor cores at 1/6 of the bf16 MMA rate.
 *  KFN_MODE_FP16X2: error-compensated with TWO fp16 limbs (11 significant bits each): x = x1 + 2^-11 * x2', the second
 *                  limb stored scaled by 2^11 so that it stays a normal fp16 number; every layer accumulates
 *                  a1w1 (three TMEM accumulators, round-robin by tap) and a1w2' + a2'w1 (fourth accumulator, scaled
 *                  by 2^-11 in the epilogue): three MMAs per operand pair instead of six, representation error
 *                  2^-22 per operand -- the 1e-4 parity configs at 1/3 of the fp16/bf16 MMA rate.  fp16's exponent
 *                  range bounds activations to |x| < 65504 (true of this network: inputs are (u8-128)*0.00625 and
 *                  every layer is followed by a ReLU of an O(1) He-scaled sum); BF16X3 is the range-safe mode.   */
enum kfn_mode { KFN_MODE_FP32 = 0, KFN_MODE_BF16 = 1, KFN_MODE_BF16X3 = 2, KFN_MODE_FP16X2 = 3 };

/* geometry of the path (KFNet/KFNet.py:7-16,70; eval.py:49-50) */
#define KFN_IMG_H 480
#define KFN_IMG_W 640
#define KFN_GRID_H 60
#define KFN_GRID_W 80
#define KFN_WINDOW 8      /* KFNet.py:369  64 // flow_sample_rate */
#define KFN_FEAT_C 32

KFN_API const char* kfn_version(void);
KFN_API const char* kfn_last_error(void);

/* image_h / image_w must be multiples of 8 (reference: 480x640, KFNetDataSpec.image_size KFNet.py:9). */
KFN_API int kfn_create(kfn_handle** out, int device, int mode, int image_h, int image_w);
KFN_API int kfn_destroy(kfn_handle* h);

/* Weights by TF variable name and TF layout (restore scopes eval.py:66-68):
 *   "<scope>/<layer>/kernel"  conv [kh,kw,Cin,Cout] | conv2d_transpose [kh,kw,Cout,Cin] | dense [in,out]
 *   "<scope>/<layer>/bias"    [Cout]
 * `data` is HOST fp32; dims/ndim must match the architecture.  Replaces RestoreFromScope (train.py:317-321). */
KFN_API int kfn_load_weight(kfn_handle* h, const char* name, const float* data, const int64_t* dims, int ndim);
/* Packs/uploads all weights for the handle's mode.  Fails if any variable is missing. */
KFN_API int kfn_finalize_weights(kfn_handle* h);

/* Bytes of caller-owned device workspace needed by the forward calls for up to `frames` frames per
 * kfn_scoordnet_forward / kfn_oflow_feat_forward call and `pairs` frame pairs per kfn_oflownet_forward call. */
KFN_API int64_t kfn_workspace_bytes(kfn_handle* h, int frames, int pairs);

/* SCoordNet (cnn_wrapper/SCoordNet.py:18-44, preprocess :34-37; KFNet.GetMeasureCoord KFNet.py:82-83).
 *   images : device uint8 [B,H,W,3]   (decode_png output, train.py:215-216, before the cast to float)
 *   measure: device fp32  [B,H/8,W/8,4] = (coord xyz, sigma_z = exp(prediction ch3))                          */
KFN_API int kfn_scoordnet_forward(kfn_handle* h, const uint8_t* images, int B, float* measure,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* OFlowNet feature tower + l2_normalize (KFNet.BuildOFlowFeat, KFNet.py:315-341).
 *   feat: device fp32 [B,H/8,W/8,32]                                                                        */
KFN_API int kfn_oflow_feat_forward(kfn_handle* h, const uint8_t* images, int B, float* feat,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* Feature-difference ("cost") volume (KFNet.BuildCoordVolume, KFNet.py:343-359 + reshape :372).
 *   diff[n, p, i, j, c] = f_cur[n,p,c] - f_prev[n, p + (j-4, i-4), c]   (0 where the shifted pixel is outside)
 *   f_prev, f_cur: device fp32 [NP,GH,GW,32]; diff: device [NP*GH*GW, 8, 8, 32], fp32 if out_bf16==0 else bf16 */
KFN_API int kfn_diff_volume(const float* f_prev, const float* f_cur, int NP, int GH, int GW, void* diff, int out_bf16,
                    void* stream);

/* OFlowNet on NP frame pairs (cnn_wrapper/OFlowNet.py:17-56 + KFNet.BuildOFlowNet KFNet.py:369-386):
 * diff volume -> U-Net -> softmax -> soft-argmax flow; conv3b -> fc1 -> fc2 -> uncertainty -> exp*1e-2.
 *   f_prev, f_cur : device fp32 [NP,GH,GW,32]
 *   flow          : device fp32 [NP,GH,GW,2]  (flow_x, flow_y) = prob . offsets
 *   sigma_w       : device fp32 [NP,GH,GW]    transition (process-noise) uncertainty
 *   logits        : optional device fp32 [NP*GH*GW, 64] (NULL to skip)  pre-softmax U-Net output, k = i*8+j
 *   argmax        : optional device int32 [NP,GH,GW]    argmax_k prob (diagnostic)                  */
KFN_API int kfn_oflownet_forward(kfn_handle* h, const float* f_prev, const float* f_cur, int NP, float* flow,
                         float* sigma_w, float* logits, int32_t* argmax, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* Flow-guided bilinear warp + process noise + Kalman update + NIS + rigid transform, as a serial scan
 * over T frames of S independent sequences (one CTA per sequence, filter state resident in shared memory).
 * Replaces KFNet.BuildOFlowNet tail (KFNet.py:386-401), tools/util.py:bilinear_sampler, KFNet.BuildKFCoord
 * (:148-162), KFNet.GetNIS (:174-184), KFNet/util.py:ApplyTransform and the state feedback / re-init of the
 * eval.py loop (:87-104, :121-126).
 *   measure  : fp32 [S,T,GH,GW,4] (xyz, sigma_z)           flow: fp32 [S,T,GH,GW,2]    sigma_w: fp32 [S,T,GH,GW]
 *              (flow/sigma_w entry t belongs to the pair (t-1, t); entries with (first_index+t) % reinit_period == 0 are ignored)
 *   state_in : fp32 [S,GH,GW,4] (last_coord, last_uncertainty; eval.py:49-50) or NULL when the scan starts on a re-init frame
 *   state_out: fp32 [S,GH,GW,4] or NULL
 *   transform: HOST fp32 [16] row-major 4x4, ALREADY inverted (train.py:51-52)
 *   reinit_period: frames with (first_index + t) % reinit_period == 0 take the measurement as state and output
 *                  (eval.py:94-101; reference value 500); <= 0 disables re-init (needs state_in)
 *   nis_gate : eval.py:87-92 (--NIS): pixels with sum_c NIS > 7.815 take the transformed measurement in out_map only
 *   out_map  : fp32 [S,T,GH,GW,4] = (T * kf_coord, 1 / kf_uncertainty)  -- the array eval.py:123 saves
 *   temp, kf : optional fp32 [S,T,GH,GW,4] (temp_coord, temp_uncertainty) / (KF_coord, KF_uncertainty), untransformed
 *   nis      : optional fp32 [S,T,GH,GW,3]
 *   taps     : optional int32 [S,T,GH,GW,2] (floor(x+flow_x), floor(y+flow_y)) of the sampler (tools/util.py:36-39) */
KFN_API int kfn_kalman_scan(const float* measure, const float* flow, const float* sigma_w, const float* state_in,
                    float* state_out, const float* transform, int S, int T, int GH, int GW, int first_index,
                    int reinit_period, int nis_gate, float* out_map, float* temp, float* kf, float* nis,
                    int32_t* taps, void* stream);

/* One recursive filter step for one frame (KFNet.GetKFCoordRecursive, KFNet.py:85-100): kfn_kalman_scan with
 * S = T = 1, no re-init.  last/measure/temp/kf: fp32 [GH,GW,4]; flow [GH,GW,2]; sigma_w [GH,GW]. */
KFN_API int kfn_flow_warp_kalman(const float* flow, const float* sigma_w, const float* last, const float* measure,
                         const float* transform, int GH, int GW, float* temp, float* kf, float* nis,
                         float* out_map, int32_t* taps, void* stream);

/* Stand-alone bilinear sampler (tools/util.py:3-94; fused into kfn_kalman_scan on the hot path): imgs device fp32
 * [N,H,W,C], coords device fp32 [N,Ht,Wt,2] = (x, y) in source pixels, out device fp32 [N,Ht,Wt,C]. */
KFN_API int kfn_bilinear_sampler(const float* imgs, const float* coords, float* out, int N, int H, int W, int C, int Ht, int Wt,
                         void* stream);

/* Stand-alone rigid transform of a coordinate map (KFNet/util.py:12-40, ApplyTransform): in/out device fp32 [npix,3],
 * transform HOST fp32 [16] row-major 4x4 (fused into kfn_kalman_scan's output on the hot path). */
KFN_API int kfn_apply_transform(const float* in, float* out, int64_t npix, const float* transform, void* stream);

/* Pointwise Kalman update and NIS on already-aligned maps (no warp): KFNet.BuildKFCoord (KFNet.py:148-162) and
 * KFNet.GetNIS (:164-184).  prior = (last/temp coord xyz, uncertainty), measure = (measure coord xyz, sigma_z):
 * device fp32 [npix,4]; posterior [npix,4] and/or nis [npix,3] (either may be NULL). */
KFN_API int kfn_kalman_update(const float* prior, const float* measure, int64_t npix, float* posterior, float* nis, void* stream);

/* Whole eval path for S sequences of T frames resident on the device (the batched, time-decoupled form of the
 * eval.py loop: all conv work frame-parallel, then the scan).  images: device uint8 [S,T,H,W,3];
 * out_map: device fp32 [S,T,GH,GW,4].  frames_per_batch / pairs_per_batch bound the workspace (see
 * kfn_workspace_bytes) -- the extra per-sequence buffers (measure, feat, flow, sigma_w) come out of the
 * same workspace: add kfn_sequence_extra_bytes(S,T). */
KFN_API int64_t kfn_sequence_extra_bytes(kfn_handle* h, int S, int T);
KFN_API int kfn_eval_sequences(kfn_handle* h, const uint8_t* images, int S, int T, const float* transform,
                       int reinit_period, int nis_gate, int frames_per_batch, int pairs_per_batch,
                       float* out_map, void* workspace, int64_t workspace_bytes, void* stream);

/* The same with HOST buffers -- the end-to-end entry point: `host_images` uint8 [S,T,H,W,3] (pinned memory recommended) are
 * streamed to the caller-owned device staging buffer `dev_images` batch by batch on an internal copy stream (each frame
 * batch's kernels wait only for that batch), the path runs on `stream`, the maps are copied to `host_out`
 * fp32 [S,T,GH,GW,4] through `dev_out`, and the call returns after synchronising `stream`. */
KFN_API int kfn_eval_sequences_host(kfn_handle* h, const uint8_t* host_images, int S, int T, const float* transform,
                            int reinit_period, int nis_gate, int frames_per_batch, int pairs_per_batch, float* host_out,
                            uint8_t* dev_images, float* dev_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Streaming form for ONE sequence delivered in chunks -- what the drop-in eval() (kfnet_b200/eval.py; reference
 * KFNet/eval.py:77-126) calls once per chunk of decoded frames: evaluates frames [first_index, first_index + T) from
 * `host_images` uint8 [T,H,W,3].  The filter state (last_coord, last_uncertainty; eval.py:49-50,102-104) and the feature
 * map of the chunk's last frame are carried between calls in `carry`: device fp32 [GH*GW*4 | GH*GW*32], caller-owned,
 * read iff first_index % reinit_period != 0, always written.  host_out fp32 [T,GH,GW,4] as kfn_eval_sequences_host.
 * host_diag / dev_diag (both or neither): fp32 planar [measure T*P*4 | temp T*P*4 | kf T*P*4 | nis T*P*3], the tensors
 * eval.py:78-83 fetches for its printed line (temp / kf / nis entries of re-init frames are the measurement / 0). */
KFN_API int kfn_eval_stream_host(kfn_handle* h, const uint8_t* host_images, int T, int first_index, const float* transform,
                         int reinit_period, int nis_gate, int frames_per_batch, int pairs_per_batch, float* carry,
                         float* host_out, float* host_diag, uint8_t* dev_images, float* dev_out, float* dev_diag,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* One packed layer of the path by name (per-layer parity tests and per-layer roofline measurements).
 *   op_key: "<scope>/<layer>" for conv / dense layers, "<scope>/<layer>#p<k>" (k = 2*py+px) for one sub-pixel phase
 *           of a conv2d_transpose, "<scope>/<layer>#all" for all four phases in one GEMM (cnn_wrapper/network.py:116-135, :418-437),
 *           "ScoreNet/conv1a#u8" / "Temporal/feat1#u8" (KFN_MODE_BF16) for the first layers fed with the u8 image itself.
 *   in    : device NHWC [N,IH,IW,in_cs] -- fp32 in KFN_MODE_FP32, bf16 in KFN_MODE_BF16 (the first Cin channels are read;
 *           for the Cin=3 first layers in BF16 mode `in` is the 32-wide im2col buffer of u8-128 values, or with "#u8" the
 *           uint8 image [N,IH,IW,3] (in_cs = 3): the kernel builds the same rows in shared memory -- the path the networks use)
 *   out   : device NHWC [N,OH,OW,out_cs], channels [out_coff, out_coff+Cout) written; bf16 unless out_fp32 (FP32 mode: always fp32)
 *   epi   : 0 none, 1 exp on channel 3 (SCoordNet.py:43), 2 l2-normalise 32 channels (KFNet.py:340; BF16 mode only) */
KFN_API int kfn_conv_layer(kfn_handle* h, const char* op_key, const void* in, int N, int IH, int IW, int in_cs, void* out,
                   int OH, int OW, int out_cs, int out_coff, int out_fp32, int epi, void* stream);

/* ---- training (reference: KFNet/train.py:268-315 `run` + `solver`, KFNet.py:102-146,192-298,430-467) ------------
 * KFN_MODE_FP32 handles only.  kfn_train_begin copies the loaded weights into the trainable parameter set (device,
 * TF variable layouts), zeroes the Adam moments and global_step and allocates the activation / gradient arena for windows of
 * `batch` consecutive frames (reference: 4).  train_scoordnet / train_oflownet mirror KFNet(..., train_scoordnet,
 * train_oflownet): variables of a frozen scope receive no gradient, no regulariser and no update (train.py:283-285).
 *
 * kfn_train_step runs one iteration: forward of SCoordNet / tower / OFlowNet / the unrolled filter (GetKFCoordBatch), the three
 * losses (CoordLossWithUncertainty + 50 SmoothLoss; 0.2 / 0.2 / 0.6), the whole backward pass, and -- if apply_update --
 * AdamOptimizer(lr = base_lr * gamma^(global_step / decay_steps)) on loss + l2_regularizer(weight_decay).
 *   images : device uint8 [batch,H,W,3]      gt8: device fp32 [batch,GH,GW,3] labels ALREADY on the label grid and moved into the
 *   network's frame (train.py:275-280)         mask8: device fp32 [batch,GH,GW,1] in {0,1}
 *   scalars_host: HOST fp32 [9] = loss, measure loss, measure accuracy, temporal loss, temporal accuracy, KF loss, KF accuracy,
 *                 regulariser, learning rate.   The call synchronises `stream`.
 * kfn_train_get copies one variable (what = 0), its last gradient (1; without the regulariser term) or its Adam m / v (2 / 3)
 * to host memory by TF variable name.  kfn_train_commit writes the trained parameters back into the handle's weight set and
 * re-packs the inference operands, so the forward entry points above run the trained network. */
KFN_API int kfn_train_begin(kfn_handle* h, int batch, int train_scoordnet, int train_oflownet);
KFN_API int kfn_train_step(kfn_handle* h, const uint8_t* images, const float* gt8, const float* mask8, float base_lr,
                   float weight_decay, int decay_steps, float gamma, int apply_update, float* scalars_host, void* stream);
KFN_API int kfn_train_get(kfn_handle* h, const char* name, int what, float* host_out, int64_t n);
KFN_API int kfn_train_commit(kfn_handle* h);
/* Sets global_step (train.py:361-362,375 `global_step.assign(step)`, --reset_step): the learning-rate decay exponent. */
KFN_API int kfn_train_set_step(kfn_handle* h, int step);

/* Number of kernels launched through this handle's entry points since the last reset (bench.py's gpu_launches);
 * h == NULL: every kernel the calling thread launched through the library, handle-less entry points included. */
KFN_API int64_t kfn_launch_count(kfn_handle* h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* KFNET_B200_H_ */
