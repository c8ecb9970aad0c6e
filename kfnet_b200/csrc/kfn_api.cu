// C ABI of libkfnet_b200.so: handle, weights, workspace arena and the layer schedules of SCoordNet,
// the OFlowNet feature tower, the OFlowNet U-Net + FC head, and the batched sequence evaluation.
// See include/kfnet_b200.h for the contract and the reference functions each entry point replaces.
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_tc.h"
#include "handle.h"
#include "layers.h"

namespace kfn {
thread_local std::string g_error;
thread_local int64_t g_launches = 0;
thread_local int64_t* g_handle_launches = nullptr;
void set_error(const std::string& msg) { g_error = msg; }

int use_device_of(const void* device_ptr, int* dev) {
  cudaPointerAttributes a;
  KFN_CUDA_OK(cudaPointerGetAttributes(&a, device_ptr));
  KFN_REQUIRE(a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged, "buffer is not device memory");
  KFN_CUDA_OK(cudaSetDevice(a.device));
  if (dev) *dev = a.device;
  return KFN_OK;
}

int sm_count(int dev) {
  static std::mutex mu;
  static int cache[64] = {0};
  std::lock_guard<std::mutex> lk(mu);
  int& c = cache[dev & 63];
  if (!c && cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) c = 148;
  return c;
}
}  // namespace kfn

using namespace kfn;

namespace {

struct Arena {
  uint8_t* base; int64_t size; int64_t off = 0; bool overflow = false;
  Arena(void* b, int64_t s) : base(reinterpret_cast<uint8_t*>(b)), size(s) {}
  void* alloc(int64_t bytes) {
    off = (off + 1023) & ~int64_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (base && off > size) overflow = true;
    return p;
  }
};

struct Act { void* p; int N, H, W, C; int limbs; };   // NHWC view, C = channels per limb of the buffer; limbs 0 = mode default

}  // namespace


namespace {

// bytes per LOGICAL activation element (allocation) and per STORED element (pointer arithmetic inside a limb)
int mode_limbs(const kfn_handle* h) { return h->mode == KFN_MODE_BF16X3 ? 3 : h->mode == KFN_MODE_FP16X2 ? 2 : 1; }   // stored limbs per activation / weight
bool half_limbs(const kfn_handle* h) { return h->mode == KFN_MODE_FP16X2; }
int esize(const kfn_handle* h) { return h->mode == KFN_MODE_FP32 ? 4 : 2 * mode_limbs(h); }
int elem_bytes(const kfn_handle* h) { return h->mode == KFN_MODE_FP32 ? 4 : 2; }
bool is_tc(const kfn_handle* h) { return h->mode != KFN_MODE_FP32; }

// tap valid iff some m-grid row reads an in-range input pixel through it
bool tap_live(int d, int M, int stride, int I) {
  for (int m = 0; m < M; ++m) { int i = m * stride + d; if (i >= 0 && i < I) return true; }
  return false;
}

// Builds the packed operand `key` from TF-layout host weights.
//   kind CONV   : taps (ky,kx) with (dy,dx) = (ky - pad, kx - pad), pad = (k==3 && stride==1) ? 1 : 0  (SAME, App. A.2)
//   kind DECONV : phase (py,px): ky in {0 (dy=0), 2 (dy=-1)} for py==0, ky=1 (dy=0) for py==1  (App. A.3)
//   first_layer : 3x3xCin=3 conv as one tap over the 32-wide im2col row, 0.00625 folded in (BF16 mode only)
int pack_op(kfn_handle* h, const std::string& key, const LayerSpec& L, int phase, bool first_layer_im2col, int MH,
            int MW, int IH, int IW) {
  const std::string base = std::string(L.scope) + "/" + L.name;
  auto itw = h->host_w.find(base + "/kernel");
  auto itb = h->host_w.find(base + "/bias");
  if (itw == h->host_w.end() || itb == h->host_w.end()) { set_error("missing weight " + base); return KFN_ERR_ARG; }
  const float* w = itw->second.data();
  const float* b = itb->second.data();
  PackedOp op;
  op.Cout = L.cout; op.relu = L.relu ? 1 : 0; op.stride = (L.kind == CONV) ? L.stride : 1;
  // (ky, kx, dy, dx) list
  struct Tap { int ky, kx, dy, dx; };
  std::vector<Tap> taps;
  if (first_layer_im2col) {
    op.Cin = 32; taps.push_back({0, 0, 0, 0});
  } else if (L.kind == DENSE || L.k == 1) {
    op.Cin = L.cin; taps.push_back({0, 0, 0, 0});
  } else if (L.kind == CONV) {
    op.Cin = L.cin;
    const int pad = (L.stride == 1) ? 1 : 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx)
        if (tap_live(ky - pad, MH, L.stride, IH) && tap_live(kx - pad, MW, L.stride, IW))
          taps.push_back({ky, kx, ky - pad, kx - pad});
  } else if (phase == 4) {   // DECONV, all four sub-pixel phases in one GEMM: taps (dy,dx) in {0,-1}^2, N = 4*Cout
    op.Cin = L.cin;
    op.Cout = 4 * L.cout;
    op.phase_cout = L.cout;
    for (int a = 0; a < 2; ++a)
      for (int c = 0; c < 2; ++c)
        if (tap_live(-a, MH, 1, IH) && tap_live(-c, MW, 1, IW)) taps.push_back({0, 0, -a, -c});
  } else {   // DECONV phase
    op.Cin = L.cin;
    const int py = phase >> 1, px = phase & 1;
    const int kys[2] = {py ? 1 : 0, 2}, dys[2] = {0, -1};
    const int kxs[2] = {px ? 1 : 0, 2}, dxs[2] = {0, -1};
    for (int a = 0; a < (py ? 1 : 2); ++a)
      for (int c = 0; c < (px ? 1 : 2); ++c)
        if (tap_live(dys[a], MH, 1, IH) && tap_live(dxs[c], MW, 1, IW)) taps.push_back({kys[a], kxs[c], dys[a], dxs[c]});
  }
  op.ntaps = (int)taps.size();
  if (op.ntaps == 0 || op.ntaps > 9) { set_error("bad tap count for " + key); return KFN_ERR_ARG; }
  {
    int full = 1;
    if (!first_layer_im2col && L.kind == CONV && L.k == 3) full = 9;
    else if (L.kind == DECONV) full = phase == 4 ? 4 : ((phase >> 1) ? 1 : 2) * ((phase & 1) ? 1 : 2);
    op.pMH = MH; op.pMW = MW; op.pIH = IH; op.pIW = IW; op.pruned = op.ntaps < full ? 1 : 0;
  }
  for (int t = 0; t < op.ntaps; ++t) { op.dy[t] = taps[t].dy; op.dx[t] = taps[t].dx; }
  const bool bf16 = is_tc(h);
  const int wl = mode_limbs(h);                          // weight limbs
  const bool hl = half_limbs(h);                         // fp16 limbs, limb 1 = residual * 2^11
  op.CoutP = bf16 ? tc_cout_padded(op.Cout) : (op.Cout + 3) / 4 * 4;

  // element (tap t, ci, co) of the logical [taps][Cin][Cout] operand
  auto W = [&](int t, int ci, int co) -> float {
    const Tap& tp = taps[t];
    if (first_layer_im2col) {
      if (ci >= 27) return 0.f;
      const int ky = ci / 9, kx = (ci / 3) % 3, c = ci % 3;
      return w[((ky * 3 + kx) * 3 + c) * L.cout + co] * 0.00625f;
    }
    if (L.kind == DENSE) return w[ci * L.cout + co];
    if (L.kind == CONV) return w[((tp.ky * L.k + tp.kx) * L.cin + ci) * L.cout + co];
    if (op.phase_cout) {                                           // merged phases (App. A.3: 2i + k = j)
      const int ph = co / L.cout, ch = co % L.cout, py = ph >> 1, px = ph & 1;
      const int ky = py ? (tp.dy == 0 ? 1 : -1) : (tp.dy == 0 ? 0 : 2);
      const int kx = px ? (tp.dx == 0 ? 1 : -1) : (tp.dx == 0 ? 0 : 2);
      if (ky < 0 || kx < 0) return 0.f;
      return w[((ky * 3 + kx) * L.cout + ch) * L.cin + ci];
    }
    return w[((tp.ky * 3 + tp.kx) * L.cout + co) * L.cin + ci];   // DECONV [kh,kw,Cout,Cin]
  };
  const size_t nelem = (size_t)op.ntaps * op.Cin * op.CoutP;
  std::vector<float> biasp(op.CoutP, 0.f);
  for (int co = 0; co < op.Cout; ++co) biasp[co] = b[co % L.cout];
  KFN_CUDA_OK(cudaMalloc(&op.bias, op.CoutP * sizeof(float)));
  h->allocs.push_back(op.bias);
  KFN_CUDA_OK(cudaMemcpy(op.bias, biasp.data(), op.CoutP * sizeof(float), cudaMemcpyHostToDevice));
  if (!bf16) {
    std::vector<float> pk(nelem, 0.f);
    for (int t = 0; t < op.ntaps; ++t)
      for (int ci = 0; ci < op.Cin; ++ci)
        for (int co = 0; co < op.Cout; ++co) pk[((size_t)t * op.Cin + ci) * op.CoutP + co] = W(t, ci, co);
    KFN_CUDA_OK(cudaMalloc(&op.w, nelem * sizeof(float)));
    h->allocs.push_back(op.w);
    KFN_CUDA_OK(cudaMemcpy(op.w, pk.data(), nelem * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    // [tap][CoutP][limb][Cin]; limb l = bf16 of the residual left by limbs < l
    std::vector<__nv_bfloat16> pk(nelem * wl, __float2bfloat16(0.f));
    for (int t = 0; t < op.ntaps; ++t)
      for (int co = 0; co < op.Cout; ++co)
        for (int ci = 0; ci < op.Cin; ++ci) {
          float r = W(t, ci, co);
          for (int l = 0; l < wl; ++l) {
            if (hl) {
              const __half hv = __float2half_rn(r);
              pk[(((size_t)t * op.CoutP + co) * wl + l) * op.Cin + ci] = *reinterpret_cast<const __nv_bfloat16*>(&hv);
              r = (r - __half2float(hv)) * 2048.0f;
            } else {
              const __nv_bfloat16 b = __float2bfloat16(r);
              pk[(((size_t)t * op.CoutP + co) * wl + l) * op.Cin + ci] = b;
              r -= __bfloat162float(b);
            }
          }
        }
    KFN_CUDA_OK(cudaMalloc(&op.w, nelem * wl * sizeof(__nv_bfloat16)));
    h->allocs.push_back(op.w);
    KFN_CUDA_OK(cudaMemcpy(op.w, pk.data(), nelem * wl * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  }
  h->ops[key] = op;
  return KFN_OK;
}

// conv1a (ScoreNet, 3->64) and feat1 (Temporal, 3->16) as one K=32 GEMM over the shared im2col rows (BF16 mode).
int pack_first_both(kfn_handle* h) {
  const float* w1 = h->host_w["ScoreNet/conv1a/kernel"].data();
  const float* b1 = h->host_w["ScoreNet/conv1a/bias"].data();
  const float* w2 = h->host_w["Temporal/feat1/kernel"].data();
  const float* b2 = h->host_w["Temporal/feat1/bias"].data();
  PackedOp op;
  op.Cin = 32; op.Cout = 80; op.CoutP = 80; op.ntaps = 1; op.relu = 1; op.stride = 1;
  std::vector<__nv_bfloat16> pk((size_t)80 * 32, __float2bfloat16(0.f));
  std::vector<float> bias(80, 0.f);
  for (int co = 0; co < 80; ++co) {
    const float* w = co < 64 ? w1 : w2;
    const int cw = co < 64 ? 64 : 16, c = co < 64 ? co : co - 64;
    bias[co] = co < 64 ? b1[co] : b2[co - 64];
    for (int k = 0; k < 27; ++k) pk[(size_t)co * 32 + k] = __float2bfloat16(w[k * cw + c] * 0.00625f);
  }
  KFN_CUDA_OK(cudaMalloc(&op.w, pk.size() * sizeof(__nv_bfloat16)));
  h->allocs.push_back(op.w);
  KFN_CUDA_OK(cudaMemcpy(op.w, pk.data(), pk.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  KFN_CUDA_OK(cudaMalloc(&op.bias, 80 * sizeof(float)));
  h->allocs.push_back(op.bias);
  KFN_CUDA_OK(cudaMemcpy(op.bias, bias.data(), 80 * sizeof(float), cudaMemcpyHostToDevice));
  h->ops["first#both"] = op;
  return KFN_OK;
}

// Runs packed op `key`: in (NHWC, `in.C` channel stride, op.Cin channels read from channel 0 of `in.p`)
// -> out channels [out_coff, out_coff + Cout) of `out` at pixel (my*omul + oaddy, mx*omul + oaddx).
// `u8_image`: `in.p` is the u8 image [N,H,W,3] and the op is one of the im2col-packed first layers (BF16 mode): the
// 32-wide operand row is built inside the kernel (conv_tc.cu, fused first layer)
int run_op(kfn_handle* h, const std::string& key, Act in, Act out, int out_coff, int MH, int MW, int omul, int oaddy,
           int oaddx, int epi, bool out_fp32, cudaStream_t st, void* out2 = nullptr, int split_col = 0, int out2_cs = 0,
           bool u8_image = false, const float* diff_cur = nullptr, int dGH = 0, int dGW = 0) {
  auto it = h->ops.find(key);
  if (it == h->ops.end()) { set_error("unknown op " + key); return KFN_ERR_STATE; }
  const PackedOp& op = it->second;
  if (op.pruned && (MH != op.pMH || MW != op.pMW || in.H != op.pIH || in.W != op.pIW)) {
    set_error(key + ": packed with taps pruned for a " + std::to_string(op.pIH) + "x" + std::to_string(op.pIW) + " input; it cannot run on " +
              std::to_string(in.H) + "x" + std::to_string(in.W));
    return KFN_ERR_ARG;
  }
  if (omul == 1 && (MH != out.H || MW != out.W || in.H != MH * op.stride || in.W != MW * op.stride) && !u8_image && !diff_cur) {
    set_error(key + ": output grid must be input / stride");
    return KFN_ERR_ARG;
  }
  if (omul == 2 && (out.H != 2 * MH || out.W != 2 * MW || in.H != MH || in.W != MW)) { set_error(key + ": conv2d_transpose output must be 2x the input"); return KFN_ERR_ARG; }
  {
    const int cout_px = op.phase_cout ? op.phase_cout : (split_col ? split_col : op.Cout);
    if (out_coff < 0 || out_coff + cout_px > out.C) { set_error(key + ": out_coff + Cout exceeds the output channel stride"); return KFN_ERR_ARG; }
    if (!u8_image && !diff_cur && in.C < op.Cin) { set_error(key + ": input channel stride smaller than Cin"); return KFN_ERR_ARG; }
  }
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = in.N; g.MH = MH; g.MW = MW;
  const int ml = mode_limbs(h);
  const bool x3 = ml > 1;
  const int la = x3 ? (in.limbs ? in.limbs : ml) : 1;
  const int lo = (x3 && !out_fp32) ? ml : 1;
  g.IH = in.H; g.IW = in.W; g.in_cs = in.C * la; g.Cin = op.Cin;
  g.limbs_a = la; g.limbs_w = ml; g.limbs_out = lo; g.in_ls = in.C; g.out_ls = out.C;
  g.half_limbs = half_limbs(h) ? 1 : 0;
  g.istride = op.stride;
  g.ntaps = op.ntaps;
  for (int t = 0; t < op.ntaps; ++t) { g.dy[t] = op.dy[t]; g.dx[t] = op.dx[t]; }
  g.OH = out.H; g.OW = out.W; g.out_cs = out.C * lo; g.out_coff = out_coff; g.Cout = op.Cout;
  g.omul = omul; g.oaddy = oaddy; g.oaddx = oaddx; g.relu = op.relu; g.epi = epi;
  g.phase_cout = op.phase_cout;
  g.split_col = split_col; g.out2_cs = out2_cs;
  g.im2col_u8 = u8_image ? 1 : 0;
  g.diff_cur = diff_cur; g.dGH = dGH; g.dGW = dGW;   // fused difference volume: in.p is f_prev (fp32)
  if (diff_cur && h->mode != KFN_MODE_BF16) { set_error(key + ": fused difference volume is BF16-mode only"); return KFN_ERR_ARG; }
  if (u8_image && (h->mode != KFN_MODE_BF16 || op.Cin != 32 || op.ntaps != 1)) {
    set_error(key + ": u8 image input is only valid for the BF16-mode first layers");
    return KFN_ERR_ARG;
  }
  if (h->mode == KFN_MODE_FP32) {
    return launch_conv_simt(g, reinterpret_cast<const float*>(in.p), reinterpret_cast<const float*>(op.w), op.CoutP,
                            op.bias, reinterpret_cast<float*>(out.p), st);
  }
  const kfn_handle::PlanKey pk(key, (const void*)in.p, out.p, diff_cur ? (void*)diff_cur : out2,   // (never both)
                               std::array<int, 18>{in.N, in.H, in.W, g.in_cs, MH, MW, out.H, out.W, g.out_cs, out_coff, omul, oaddy,
                                                   oaddx, epi, out_fp32 ? 1 : 0, split_col, out2_cs, g.im2col_u8});
  auto pit = h->plans.find(pk);
  if (pit == h->plans.end()) {
    TcConvPlan plan;
    int rc = prepare_conv_tc(g, in.p, op.w, op.CoutP, op.bias, out.p, out_fp32 ? 0 : 1, h->num_sms, &plan, out2);
    if (rc != KFN_OK) { set_error(key + ": " + g_error); return rc; }
    if (h->plans.size() > 4096) h->plans.erase(h->plans.begin());   // bounded cache: drop one entry, not the working set
    pit = h->plans.emplace(pk, plan).first;
  }
  return launch_conv_tc(pit->second, st);
}

// plain conv: m-grid == output grid
int conv(kfn_handle* h, const std::string& key, Act in, Act out, int out_coff, int epi, bool out_fp32, cudaStream_t st) {
  return run_op(h, key, in, out, out_coff, out.H, out.W, 1, 0, 0, epi, out_fp32, st);
}
// conv2d_transpose k3 s2 SAME = 4 sub-pixel phase convs over the input grid, run as ONE GEMM with N = 4*Cout
// (the per-phase ops "<key>#p<k>" stay available through kfn_conv_layer)
int deconv(kfn_handle* h, const std::string& key, Act in, Act out, int out_coff, cudaStream_t st) {
  return run_op(h, key + "#all", in, out, out_coff, in.H, in.W, 2, 0, 0, EPI_NONE, false, st);
}

Act act(void* p, int N, int H, int W, int C, int limbs = 0) { return Act{p, N, H, W, C, limbs}; }
void* offs(void* p, int64_t elems, int es) { return p ? reinterpret_cast<uint8_t*>(p) + elems * es : nullptr; }

// BF16 mode: first layers read the u8 image directly (KFN_TC_IM2COL_FUSED; otherwise im2col kernel + TMA path)
bool first_fused(const kfn_handle* h) { return tc_im2col_fused_enabled() && h->mode == KFN_MODE_BF16; }

// conv1a (+feat1) -> conv1b in one kernel (conv1ab_fused.cu): `out` = conv1b's output [N,H,W,64] bf16, `feat1` optional
int run_conv1ab(kfn_handle* h, const uint8_t* images, int N, int H, int W, void* out, void* feat1, cudaStream_t st) {
  auto a = h->ops.find("first#both"), b = h->ops.find("ScoreNet/conv1b");
  if (h->mode != KFN_MODE_BF16 || a == h->ops.end() || b == h->ops.end() || b->second.ntaps != 9) {
    set_error("fused conv1a+conv1b: BF16 mode with all nine conv1b taps only");
    return KFN_ERR_STATE;
  }
  auto key = std::make_tuple((const void*)images, out, feat1, N, H, W);
  auto it = h->fused_plans.find(key);
  if (it == h->fused_plans.end()) {
    Conv1abPlan plan;
    int rc = prepare_conv1ab(images, N, H, W, a->second.w, a->second.bias, b->second.w, b->second.bias, out, feat1, h->num_sms, &plan);
    if (rc != KFN_OK) return rc;
    if (h->fused_plans.size() > 1024) h->fused_plans.clear();
    it = h->fused_plans.emplace(key, plan).first;
  }
  return launch_conv1ab(it->second, st);
}
bool fuse_conv1(const kfn_handle* h) { return first_fused(h) && tc_fuse_conv1_enabled() && conv1ab_shape_ok(h->H, h->W); }

// ---------------------------------------------------------------------------------------------
// `shared_pre`: optional already-preprocessed input (im2col / fp32) shared with the feature tower by kfn_eval_sequences
// `feat1_out`: when given (BF16 mode, kfn_eval_sequences) conv1a also produces the tower's feat1 into this buffer
int scoordnet_impl(kfn_handle* h, const uint8_t* images, int B, float* measure, Arena& ar, cudaStream_t st, bool dry,
                   void* shared_pre = nullptr, void* feat1_out = nullptr) {
  const int H = h->H, W = h->W, es = esize(h);
  const int64_t px = (int64_t)B * H * W;
  const bool fused = first_fused(h);
  void* pre = (shared_pre || fused) ? shared_pre : ar.alloc(h->mode == KFN_MODE_FP32 ? px * 3 * 4 : px * 32 * 2);
  void* bufA = ar.alloc(px * 64 * es);
  void* bufB = ar.alloc(px * 64 * es);
  if (dry) return KFN_OK;
  if (ar.overflow) { set_error("workspace too small for scoordnet_forward"); return KFN_ERR_WORKSPACE; }
  int rc;
  if (h->mode == KFN_MODE_FP32) {
    if (!shared_pre && (rc = launch_preprocess_f32(images, (float*)pre, px * 3, st))) return rc;
    if ((rc = conv(h, "ScoreNet/conv1a", act(pre, B, H, W, 3), act(bufA, B, H, W, 64), 0, EPI_NONE, false, st))) return rc;
  } else if (fuse_conv1(h)) {
    if ((rc = run_conv1ab(h, images, B, H, W, bufB, feat1_out, st))) return rc;   // conv1a never materialised
  } else if (fused) {
    Act img = act(const_cast<uint8_t*>(images), B, H, W, 32, 1);
    if (feat1_out) rc = run_op(h, "first#both", img, act(bufA, B, H, W, 64), 0, H, W, 1, 0, 0, EPI_NONE, false, st, feat1_out, 64, 16, true);
    else rc = run_op(h, "ScoreNet/conv1a", img, act(bufA, B, H, W, 64), 0, H, W, 1, 0, 0, EPI_NONE, false, st, nullptr, 0, 0, true);
    if (rc) return rc;
  } else {
    if (!shared_pre && (rc = launch_im2col27_bf16(images, (__nv_bfloat16*)pre, B, H, W, st, half_limbs(h)))) return rc;
    if (feat1_out && h->mode == KFN_MODE_BF16) {
      if ((rc = run_op(h, "first#both", act(pre, B, H, W, 32, 1), act(bufA, B, H, W, 64), 0, H, W, 1, 0, 0, EPI_NONE, false, st,
                       feat1_out, 64, 16)))
        return rc;
    } else if ((rc = conv(h, "ScoreNet/conv1a", act(pre, B, H, W, 32, 1), act(bufA, B, H, W, 64), 0, EPI_NONE, false, st))) return rc;
  }
  struct Step { const char* name; int div; int c; };
  const Step steps[] = {{"conv1b", 1, 64},  {"conv2a", 2, 256}, {"conv2b", 2, 256}, {"conv3a", 4, 512}, {"conv3b", 4, 512},
                        {"conv4a", 8, 1024}, {"conv4b", 8, 1024}, {"conv5", 8, 512}, {"conv6", 8, 256}, {"conv7", 8, 128}};
  const bool conv1b_done = h->mode != KFN_MODE_FP32 && fuse_conv1(h);
  Act cur = act(conv1b_done ? bufB : bufA, B, H, W, 64);
  void* nextbuf = conv1b_done ? bufA : bufB;
  for (const Step& s : steps) {
    if (conv1b_done && &s == &steps[0]) continue;
    Act nxt = act(nextbuf, B, H / s.div, W / s.div, s.c);
    if ((rc = conv(h, std::string("ScoreNet/") + s.name, cur, nxt, 0, EPI_NONE, false, st))) return rc;
    nextbuf = cur.p; cur = nxt;
  }
  // prediction (no ReLU) + exp on channel 3 -> (coord xyz, sigma_z)   SCoordNet.py:32,39-44
  return conv(h, "ScoreNet/prediction", cur, act(measure, B, H / 8, W / 8, 4), 0, EPI_EXP_CH3, true, st);
}

// `feat1_done`: feat1's output buffer already filled by the fused first layer (see scoordnet_impl)
int tower_impl(kfn_handle* h, const uint8_t* images, int B, float* feat, Arena& ar, cudaStream_t st, bool dry,
               void* shared_pre = nullptr, void* feat1_done = nullptr) {
  const int H = h->H, W = h->W, es = esize(h);
  const int64_t px = (int64_t)B * H * W;
  const bool fused = first_fused(h);
  void* pre = (shared_pre || fused) ? shared_pre : ar.alloc(h->mode == KFN_MODE_FP32 ? px * 3 * 4 : px * 32 * 2);
  void* bufA = feat1_done ? feat1_done : ar.alloc(px * 16 * es);
  void* bufB = ar.alloc(px * 16 * es);
  if (dry) return KFN_OK;
  if (ar.overflow) { set_error("workspace too small for oflow_feat_forward"); return KFN_ERR_WORKSPACE; }
  int rc;
  if (h->mode == KFN_MODE_FP32) {
    if (!shared_pre && (rc = launch_preprocess_f32(images, (float*)pre, px * 3, st))) return rc;
    if ((rc = conv(h, "Temporal/feat1", act(pre, B, H, W, 3), act(bufA, B, H, W, 16), 0, EPI_NONE, false, st))) return rc;
  } else if (fused) {
    if (!feat1_done && (rc = run_op(h, "Temporal/feat1", act(const_cast<uint8_t*>(images), B, H, W, 32, 1), act(bufA, B, H, W, 16), 0,
                                    H, W, 1, 0, 0, EPI_NONE, false, st, nullptr, 0, 0, true)))
      return rc;
  } else {
    if (!shared_pre && (rc = launch_im2col27_bf16(images, (__nv_bfloat16*)pre, B, H, W, st, half_limbs(h)))) return rc;
    if (!feat1_done && (rc = conv(h, "Temporal/feat1", act(pre, B, H, W, 32, 1), act(bufA, B, H, W, 16), 0, EPI_NONE, false, st))) return rc;
  }
  struct Step { const char* name; int div; int c; };
  const Step steps[] = {{"feat2", 2, 32}, {"feat3", 2, 32}, {"feat4", 4, 64}, {"feat5", 4, 64}, {"feat6", 8, 128}};
  Act cur = act(bufA, B, H, W, 16);
  void* nextbuf = bufB;
  for (const Step& s : steps) {
    Act nxt = act(nextbuf, B, H / s.div, W / s.div, s.c);
    if ((rc = conv(h, std::string("Temporal/") + s.name, cur, nxt, 0, EPI_NONE, false, st))) return rc;
    nextbuf = cur.p; cur = nxt;
  }
  Act out = act(feat, B, H / 8, W / 8, 32);
  if (h->mode == KFN_MODE_FP32) {
    if ((rc = conv(h, "Temporal/feat7", cur, out, 0, EPI_NONE, true, st))) return rc;
    return launch_l2norm32(feat, (int64_t)B * (H / 8) * (W / 8), st);
  }
  return conv(h, "Temporal/feat7", cur, out, 0, EPI_L2NORM, true, st);
}

int oflownet_impl(kfn_handle* h, const float* f_prev, const float* f_cur, int NP, float* flow, float* sigma_w,
                  float* logits_out, int32_t* argmax, Arena& ar, cudaStream_t st, bool dry) {
  const int es = esize(h);
  const int64_t R64 = (int64_t)NP * h->GH * h->GW;
  const int R = (int)R64;
  void* diff = ar.alloc(R64 * 64 * 32 * es);
  void* cat0 = ar.alloc(R64 * 64 * 48 * es);   // [upconv0 (16) | conv0 (32)]     OFlowNet.py:38-39
  void* c1a = ar.alloc(R64 * 16 * 32 * es);
  void* cat1 = ar.alloc(R64 * 16 * 64 * es);   // [upconv1 (32) | conv1b (32)]    OFlowNet.py:33-34
  void* c2a = ar.alloc(R64 * 4 * 64 * es);
  void* cat2 = ar.alloc(R64 * 4 * 128 * es);   // [upconv2 (64) | conv2b (64)]    OFlowNet.py:28-29
  void* c3a = ar.alloc(R64 * 128 * es);
  void* c3b = ar.alloc(R64 * 128 * es);
  void* c4 = ar.alloc(R64 * 4 * 64 * es);
  void* c5 = ar.alloc(R64 * 16 * 32 * es);
  void* c6 = ar.alloc(R64 * 64 * 16 * es);
  void* fc1 = ar.alloc(R64 * 64 * es);
  void* fc2 = ar.alloc(R64 * 32 * es);
  float* unc = (float*)ar.alloc(R64 * 4);
  float* logits = logits_out ? logits_out : (float*)ar.alloc(R64 * 64 * 4);
  if (dry) return KFN_OK;
  if (ar.overflow) { set_error("workspace too small for oflownet_forward"); return KFN_ERR_WORKSPACE; }
  int rc;
  const int vs = elem_bytes(h);   // bytes per stored element, for channel-offset views
  const std::string T = "Temporal/";
#define RUN(x) if ((rc = (x))) return rc
  if (h->mode == KFN_MODE_BF16 && tc_diff_fused_enabled()) {
    // conv0 builds its operand rows f_cur - shifted f_prev itself: the 64x32 volume per label pixel never reaches HBM
    RUN(run_op(h, T + "conv0", act(const_cast<float*>(f_prev), R, 8, 8, 32, 1), act(cat0, R, 8, 8, 48), 16, 8, 8, 1, 0, 0, EPI_NONE,
               false, st, nullptr, 0, 0, false, f_cur, h->GH, h->GW));
  } else {
    if ((rc = launch_diff_volume(f_prev, f_cur, NP, h->GH, h->GW, diff,
                                 h->mode == KFN_MODE_BF16 ? 1 : h->mode == KFN_MODE_BF16X3 ? 3 : h->mode == KFN_MODE_FP16X2 ? 2 : 0, st)))
      return rc;
    RUN(conv(h, T + "conv0", act(diff, R, 8, 8, 32), act(cat0, R, 8, 8, 48), 16, EPI_NONE, false, st));
  }
  RUN(conv(h, T + "conv1a", act(offs(cat0, 16, vs), R, 8, 8, 48), act(c1a, R, 4, 4, 32), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "conv1b", act(c1a, R, 4, 4, 32), act(cat1, R, 4, 4, 64), 32, EPI_NONE, false, st));
  RUN(conv(h, T + "conv2a", act(offs(cat1, 32, vs), R, 4, 4, 64), act(c2a, R, 2, 2, 64), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "conv2b", act(c2a, R, 2, 2, 64), act(cat2, R, 2, 2, 128), 64, EPI_NONE, false, st));
  RUN(conv(h, T + "conv3a", act(offs(cat2, 64, vs), R, 2, 2, 128), act(c3a, R, 1, 1, 128), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "conv3b", act(c3a, R, 1, 1, 128), act(c3b, R, 1, 1, 128), 0, EPI_NONE, false, st));
  RUN(deconv(h, T + "upconv2", act(c3b, R, 1, 1, 128), act(cat2, R, 2, 2, 128), 0, st));
  RUN(conv(h, T + "conv4", act(cat2, R, 2, 2, 128), act(c4, R, 2, 2, 64), 0, EPI_NONE, false, st));
  RUN(deconv(h, T + "upconv1", act(c4, R, 2, 2, 64), act(cat1, R, 4, 4, 64), 0, st));
  RUN(conv(h, T + "conv5", act(cat1, R, 4, 4, 64), act(c5, R, 4, 4, 32), 0, EPI_NONE, false, st));
  RUN(deconv(h, T + "upconv0", act(c5, R, 4, 4, 32), act(cat0, R, 8, 8, 48), 0, st));
  RUN(conv(h, T + "conv6", act(cat0, R, 8, 8, 48), act(c6, R, 8, 8, 16), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "prediction", act(c6, R, 8, 8, 16), act(logits, R, 8, 8, 1), 0, EPI_NONE, true, st));
  RUN(conv(h, T + "fc1", act(c3b, R, 1, 1, 128), act(fc1, R, 1, 1, 64), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "fc2", act(fc1, R, 1, 1, 64), act(fc2, R, 1, 1, 32), 0, EPI_NONE, false, st));
  RUN(conv(h, T + "uncertainty", act(fc2, R, 1, 1, 32), act(unc, R, 1, 1, 1), 0, EPI_NONE, true, st));
#undef RUN
  return launch_softmax_flow(logits, unc, 1, R64, flow, sigma_w, argmax, st);
}

}  // namespace

// =============================================================================================
extern "C" {

const char* kfn_version(void) { return "kfnet_b200 0.1 (sm_100a)"; }
const char* kfn_last_error(void) { return kfn::g_error.c_str(); }

int kfn_create(kfn_handle** out, int device, int mode, int image_h, int image_w) {
  KFN_REQUIRE(out != nullptr, "null out");
  KFN_REQUIRE(mode == KFN_MODE_FP32 || mode == KFN_MODE_BF16 || mode == KFN_MODE_BF16X3 || mode == KFN_MODE_FP16X2, "unknown mode");
  KFN_REQUIRE(image_h > 0 && image_w > 0 && image_h % 8 == 0 && image_w % 8 == 0, "image size must be a multiple of 8");
  int ndev = 0;
  KFN_CUDA_OK(cudaGetDeviceCount(&ndev));
  KFN_REQUIRE(device >= 0 && device < ndev, "no such CUDA device (this library has no CPU fallback)");
  KFN_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  KFN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { set_error("libkfnet_b200 is built for sm_100a only; found sm_" + std::to_string(prop.major * 10 + prop.minor)); return KFN_ERR_CUDA; }
  kfn_handle* h = new kfn_handle();
  h->device = device; h->mode = mode; h->H = image_h; h->W = image_w; h->GH = image_h / 8; h->GW = image_w / 8;
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return KFN_OK;
}

int kfn_destroy(kfn_handle* h) {
  if (!h) return KFN_OK;
  if (kfn::g_handle_launches == &h->launches) kfn::g_handle_launches = nullptr;
  cudaSetDevice(h->device);
  if (h->train) { kfn::train_state_free(h->train); h->train = nullptr; }
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->copy_events) cudaEventDestroy(e);
  if (h->compute_done) cudaEventDestroy(h->compute_done);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return KFN_OK;
}

int kfn_load_weight(kfn_handle* h, const char* name, const float* data, const int64_t* dims, int ndim) {
  KFN_REQUIRE(h && name && data && dims, "null argument");
  std::string nm(name);
  size_t cut = nm.rfind('/');
  KFN_REQUIRE(cut != std::string::npos, "weight name must be <scope>/<layer>/{kernel,bias}");
  const LayerSpec* L = find_layer(nm.substr(0, cut));
  const std::string what = nm.substr(cut + 1);
  KFN_REQUIRE(L != nullptr, ("unknown layer in " + nm).c_str());
  int64_t expect[4]; int en = 0;
  if (what == "bias") { expect[0] = L->cout; en = 1; }
  else if (what == "kernel") {
    if (L->kind == CONV) { expect[0] = L->k; expect[1] = L->k; expect[2] = L->cin; expect[3] = L->cout; en = 4; }
    else if (L->kind == DECONV) { expect[0] = L->k; expect[1] = L->k; expect[2] = L->cout; expect[3] = L->cin; en = 4; }
    else { expect[0] = L->cin; expect[1] = L->cout; en = 2; }
  } else { set_error("weight name must end in /kernel or /bias: " + nm); return KFN_ERR_ARG; }
  KFN_REQUIRE(ndim == en, ("rank mismatch for " + nm).c_str());
  int64_t n = 1;
  for (int i = 0; i < en; ++i) { KFN_REQUIRE(dims[i] == expect[i], ("shape mismatch for " + nm).c_str()); n *= dims[i]; }
  h->host_w[nm].assign(data, data + n);
  h->finalized = false;
  return KFN_OK;
}

int kfn_finalize_weights(kfn_handle* h) {
  KFN_REQUIRE(h, "null handle");
  KFN_CUDA_OK(cudaSetDevice(h->device));
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear(); h->ops.clear(); h->plans.clear(); h->fused_plans.clear();
  const bool bf16 = is_tc(h);
  // spatial dims (output m-grid MH x MW, input IH x IW) of every layer, for tap pruning
  int rc;
  int sh = h->H, sw = h->W;   // running SCoordNet / tower resolution
  for (int i = 0; i < kNumLayers; ++i) {
    const LayerSpec& L = kLayers[i];
    const std::string key = std::string(L.scope) + "/" + L.name;
    const std::string nm(L.name);
    if (std::string(L.scope) == "ScoreNet" || nm.rfind("feat", 0) == 0) {
      if (nm == "conv1a" || nm == "feat1") { sh = h->H; sw = h->W; }
      const int ih = sh, iw = sw;
      if (L.stride == 2) { sh /= 2; sw /= 2; }
      const bool first = bf16 && L.cin == 3;
      if ((rc = pack_op(h, key, L, 0, first, sh, sw, ih, iw))) return rc;
      continue;
    }
    // U-Net: fixed 8x8 windows
    int ih = 8, oh = 8;
    if (nm == "conv0" || nm == "conv6" || nm == "prediction") { ih = 8; oh = 8; }
    else if (nm == "conv1a") { ih = 8; oh = 4; }
    else if (nm == "conv1b" || nm == "conv5") { ih = 4; oh = 4; }
    else if (nm == "conv2a") { ih = 4; oh = 2; }
    else if (nm == "conv2b" || nm == "conv4") { ih = 2; oh = 2; }
    else if (nm == "conv3a") { ih = 2; oh = 1; }
    else if (nm == "conv3b") { ih = 1; oh = 1; }
    else if (nm == "upconv2") { ih = 1; oh = 1; }   // phases run over the INPUT grid
    else if (nm == "upconv1") { ih = 2; oh = 2; }
    else if (nm == "upconv0") { ih = 4; oh = 4; }
    else { ih = 1; oh = 1; }                        // fc1, fc2, uncertainty
    if (L.kind == DECONV) {
      for (int ph = 0; ph < 4; ++ph)
        if ((rc = pack_op(h, key + "#p" + std::to_string(ph), L, ph, false, oh, oh, ih, ih))) return rc;
      if ((rc = pack_op(h, key + "#all", L, 4, false, oh, oh, ih, ih))) return rc;
    } else {
      if ((rc = pack_op(h, key, L, 0, false, oh, oh, ih, ih))) return rc;
    }
  }
  if (h->mode == KFN_MODE_BF16) {
    // "first#both": conv1a (64) and feat1 (16) read the same im2col rows -> one GEMM with N = 80, dual output
    if ((rc = pack_first_both(h))) return rc;
  }
  h->finalized = true;
  return KFN_OK;
}

int64_t kfn_workspace_bytes(kfn_handle* h, int frames, int pairs) {
  if (!h || frames < 0 || pairs < 0) return KFN_ERR_ARG;
  int64_t need = 0;
  if (frames > 0) {
    Arena a(nullptr, 0); scoordnet_impl(h, nullptr, frames, nullptr, a, nullptr, true); need = a.off > need ? a.off : need;
    Arena b(nullptr, 0); tower_impl(h, nullptr, frames, nullptr, b, nullptr, true); need = b.off > need ? b.off : need;
  }
  if (pairs > 0) {
    Arena c(nullptr, 0); oflownet_impl(h, nullptr, nullptr, pairs, nullptr, nullptr, nullptr, nullptr, c, nullptr, true);
    need = c.off > need ? c.off : need;
  }
  if (frames > 0) need += (int64_t)frames * h->H * h->W * 16 * 2 + 1024;   // fused first layer's feat1 buffer (kfn_eval_sequences)
  return need + 4096;
}

#define KFN_READY(h) do { KFN_REQUIRE(h, "null handle"); if (!h->finalized) { set_error("call kfn_finalize_weights first"); return KFN_ERR_STATE; } KFN_CUDA_OK(cudaSetDevice(h->device)); kfn::g_handle_launches = &h->launches; } while (0)

int kfn_scoordnet_forward(kfn_handle* h, const uint8_t* images, int B, float* measure, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(images && measure && workspace && B > 0, "null/empty argument");
  Arena ar(workspace, workspace_bytes);
  return scoordnet_impl(h, images, B, measure, ar, (cudaStream_t)stream, false);
}

int kfn_oflow_feat_forward(kfn_handle* h, const uint8_t* images, int B, float* feat, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(images && feat && workspace && B > 0, "null/empty argument");
  Arena ar(workspace, workspace_bytes);
  return tower_impl(h, images, B, feat, ar, (cudaStream_t)stream, false);
}

int kfn_diff_volume(const float* f_prev, const float* f_cur, int NP, int GH, int GW, void* diff, int out_bf16,
                    void* stream) {
  KFN_REQUIRE(f_prev && f_cur && diff && NP > 0 && GH > 0 && GW > 0, "null/empty argument");
  if (int rc = use_device_of(diff, nullptr)) return rc;
  return launch_diff_volume(f_prev, f_cur, NP, GH, GW, diff, out_bf16, (cudaStream_t)stream);
}

int kfn_oflownet_forward(kfn_handle* h, const float* f_prev, const float* f_cur, int NP, float* flow, float* sigma_w,
                         float* logits, int32_t* argmax, void* workspace, int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(f_prev && f_cur && flow && sigma_w && workspace && NP > 0, "null/empty argument");
  Arena ar(workspace, workspace_bytes);
  return oflownet_impl(h, f_prev, f_cur, NP, flow, sigma_w, logits, argmax, ar, (cudaStream_t)stream, false);
}

int kfn_kalman_scan(const float* measure, const float* flow, const float* sigma_w, const float* state_in,
                    float* state_out, const float* transform, int S, int T, int GH, int GW, int first_index,
                    int reinit_period, int nis_gate, float* out_map, float* temp, float* kf, float* nis,
                    int32_t* taps, void* stream) {
  KFN_REQUIRE(measure && flow && sigma_w && transform && out_map, "null argument");
  if (int rc = use_device_of(out_map, nullptr)) return rc;
  return launch_kalman_scan(measure, flow, sigma_w, state_in, state_out, transform, S, T, GH, GW, first_index,
                            reinit_period, nis_gate, out_map, temp, kf, nis, taps, (cudaStream_t)stream);
}

int kfn_flow_warp_kalman(const float* flow, const float* sigma_w, const float* last, const float* measure,
                         const float* transform, int GH, int GW, float* temp, float* kf, float* nis, float* out_map,
                         int32_t* taps, void* stream) {
  KFN_REQUIRE(flow && sigma_w && last && measure && transform && out_map, "null argument");
  if (int rc = use_device_of(out_map, nullptr)) return rc;
  return launch_kalman_scan(measure, flow, sigma_w, last, nullptr, transform, 1, 1, GH, GW, 1, 0, 0, out_map, temp, kf,
                            nis, taps, (cudaStream_t)stream);
}

int kfn_bilinear_sampler(const float* imgs, const float* coords, float* out, int N, int H, int W, int C, int Ht, int Wt,
                         void* stream) {
  KFN_REQUIRE(imgs && coords && out && N > 0 && H > 0 && W > 0 && C > 0 && Ht > 0 && Wt > 0, "null/empty argument");
  if (int rc = use_device_of(out, nullptr)) return rc;
  return launch_bilinear_sampler(imgs, coords, out, N, H, W, C, Ht, Wt, (cudaStream_t)stream);
}

int kfn_apply_transform(const float* in, float* out, int64_t npix, const float* transform, void* stream) {
  KFN_REQUIRE(in && out && transform && npix > 0, "null/empty argument");
  if (int rc = use_device_of(out, nullptr)) return rc;
  return launch_apply_transform(in, out, npix, transform, (cudaStream_t)stream);
}

int kfn_kalman_update(const float* prior, const float* measure, int64_t npix, float* posterior, float* nis, void* stream) {
  KFN_REQUIRE(prior && measure && npix > 0 && (posterior || nis), "null/empty argument");
  if (int rc = use_device_of(prior, nullptr)) return rc;
  return launch_kalman_update(prior, measure, npix, posterior, nis, (cudaStream_t)stream);
}

int64_t kfn_sequence_extra_bytes(kfn_handle* h, int S, int T) {
  if (!h || S <= 0 || T <= 0) return KFN_ERR_ARG;
  const int64_t F = (int64_t)S * T, P = (int64_t)h->GH * h->GW;
  return F * P * (4 + 32 + 2 + 1) * 4 + P * 32 * 4 + 8 * 1024;   // + the carried feature slot
}

}  // extern "C"

namespace {
// Optional extras of the streaming form (kfn_eval_stream_host): one sequence delivered in chunks.
struct StreamExtras {
  int first_index = 0;          // global index of the chunk's first frame (re-init when (first_index + t) % period == 0)
  float* carry = nullptr;       // device fp32 [P*4 filter state | P*32 features of the previous chunk's last frame], in/out
  float* measure_out = nullptr; // device fp32 [T,P,4]: use as the measurement buffer (diagnostics) instead of the workspace
  float* temp = nullptr;        // device fp32 [T,P,4]
  float* kf = nullptr;          // device fp32 [T,P,4]
  float* nis = nullptr;         // device fp32 [T,P,3]
};

// `batch_ready`: optional events, one per frame batch, the compute stream waits on before touching that batch's images
int eval_sequences_impl(kfn_handle* h, const uint8_t* images, int S, int T, const float* transform, int reinit_period,
                        int nis_gate, int frames_per_batch, int pairs_per_batch, float* out_map, void* workspace,
                        int64_t workspace_bytes, cudaStream_t st, const cudaEvent_t* batch_ready, const StreamExtras* ex = nullptr) {
  const int64_t F = (int64_t)S * T, P = (int64_t)h->GH * h->GW;
  const int period = reinit_period > 0 ? reinit_period : (1 << 30);
  const int first_index = ex ? ex->first_index : 0;
  // streaming: the chunk continues a sequence iff its first frame is not a re-initialisation frame
  const bool continues = ex && ex->carry && (first_index % period != 0);
  Arena top(workspace, workspace_bytes);
  float* measure = (ex && ex->measure_out) ? ex->measure_out : (float*)top.alloc(F * P * 4 * 4);
  float* featbuf = (float*)top.alloc((F + 1) * P * 32 * 4);       // slot 0: features of the frame before this chunk
  float* feat = featbuf + P * 32;
  float* flow = (float*)top.alloc(F * P * 2 * 4);
  float* sigw = (float*)top.alloc(F * P * 4);
  top.alloc(0);
  if (top.overflow) { set_error("workspace too small for eval_sequences buffers"); return KFN_ERR_WORKSPACE; }
  uint8_t* sub = reinterpret_cast<uint8_t*>(workspace) + top.off;
  const int64_t sub_bytes = workspace_bytes - top.off;
  const int64_t img_bytes = (int64_t)h->H * h->W * 3;
  int rc;
  if (continues) KFN_CUDA_OK(cudaMemcpyAsync(featbuf, ex->carry + P * 4, (size_t)(P * 32 * 4), cudaMemcpyDeviceToDevice, st));
  NvtxRange nv_all("kfn_eval_sequences");
  // 1. measurement system + feature tower: frame-parallel over all S*T frames
  for (int64_t f0 = 0; f0 < F; f0 += frames_per_batch) {
    NvtxRange nv("scoordnet+tower batch");
    const int B = (int)((F - f0) < frames_per_batch ? (F - f0) : frames_per_batch);
    if (batch_ready) KFN_CUDA_OK(cudaStreamWaitEvent(st, batch_ready[f0 / frames_per_batch], 0));
    // both networks start from the same preprocessed input (im2col rows / fp32 image): build it once per batch
    Arena a(sub, sub_bytes);
    const int64_t px = (int64_t)B * h->H * h->W;
    const bool fused = first_fused(h);   // first layers read the u8 frames directly: no preprocessed copy at all
    void* pre = fused ? nullptr : a.alloc(h->mode == KFN_MODE_FP32 ? px * 3 * 4 : px * 32 * 2);
    if (a.overflow) { set_error("workspace too small for eval_sequences"); return KFN_ERR_WORKSPACE; }
    if (h->mode == KFN_MODE_FP32) { if ((rc = launch_preprocess_f32(images + f0 * img_bytes, (float*)pre, px * 3, st))) return rc; }
    else if (!fused) { if ((rc = launch_im2col27_bf16(images + f0 * img_bytes, (__nv_bfloat16*)pre, B, h->H, h->W, st, half_limbs(h)))) return rc; }
    static const bool fuse_first = [] { const char* e = getenv("KFN_FUSE_FIRST"); return !(e && e[0] == '0'); }();   // A/B switch
    void* f1 = (h->mode == KFN_MODE_BF16 && fuse_first) ? a.alloc(px * 16 * 2) : nullptr;   // feat1 output, produced by the fused first layer
    if (a.overflow) { set_error("workspace too small for eval_sequences"); return KFN_ERR_WORKSPACE; }
    const int64_t mark = a.off;
    if ((rc = scoordnet_impl(h, images + f0 * img_bytes, B, measure + f0 * P * 4, a, st, false, pre, f1))) return rc;
    // The tower runs in sub-batches of <= 16 frames of the same batch: its small layers are fastest when a sub-batch's activations
    // stay L2-resident, while SCoordNet wants LARGE batches (the 60x80-grid layers have 37.5 M-tiles per frame: on 148 SMs a batch of 16
    // ends in a 20 %-filled last wave for conv6, 10 % for conv5; 71 frames = 2662.5 tiles = 17.99 waves)
    const int TB = 16;
    const int64_t pre_frame = (int64_t)h->H * h->W * (h->mode == KFN_MODE_FP32 ? 3 * 4 : 32 * 2);   // bytes of `pre` per frame
    for (int b0 = 0; b0 < B; b0 += TB) {
      const int b = (B - b0) < TB ? (B - b0) : TB;
      a.off = mark;
      void* pre_sub = pre ? reinterpret_cast<uint8_t*>(pre) + b0 * pre_frame : nullptr;
      void* f1_sub = f1 ? reinterpret_cast<uint8_t*>(f1) + (int64_t)b0 * h->H * h->W * 16 * 2 : nullptr;
      if ((rc = tower_impl(h, images + (f0 + b0) * img_bytes, b, feat + (f0 + b0) * P * 32, a, st, false, pre_sub, f1_sub))) return rc;
    }
  }
  // 2. process system: pair (f-1, f) for every flat frame f >= 1 (f >= 0 when the chunk continues a sequence: slot 0 of
  //    featbuf then holds frame -1).  Pairs that straddle two sequences are computed and ignored (frame 0 of every
  //    sequence re-initialises the filter, eval.py:94-101).
  for (int64_t f0 = continues ? 0 : 1; f0 < F; f0 += pairs_per_batch) {
    const int NP = (int)((F - f0) < pairs_per_batch ? (F - f0) : pairs_per_batch);
    NvtxRange nv("oflownet pair batch");
    Arena a(sub, sub_bytes);
    if ((rc = oflownet_impl(h, featbuf + f0 * P * 32, feat + f0 * P * 32, NP, flow + f0 * P * 2, sigw + f0 * P, nullptr,
                            nullptr, a, st, false)))
      return rc;
  }
  // 3. serial scan: warp + Kalman update, one CTA per sequence
  NvtxRange nv_scan("warp+kalman scan");
  if ((rc = launch_kalman_scan(measure, flow, sigw, continues ? ex->carry : nullptr, (ex && ex->carry) ? ex->carry : nullptr, transform,
                               S, T, h->GH, h->GW, first_index, period, nis_gate, out_map, ex ? ex->temp : nullptr,
                               ex ? ex->kf : nullptr, ex ? ex->nis : nullptr, nullptr, st)))
    return rc;
  if (ex && ex->carry)   // features of the chunk's last frame, for the first pair of the next chunk
    KFN_CUDA_OK(cudaMemcpyAsync(ex->carry + P * 4, feat + (F - 1) * P * 32, (size_t)(P * 32 * 4), cudaMemcpyDeviceToDevice, st));
  return KFN_OK;
}

// H2D of the frames batch by batch on the handle's copy stream; fills h->copy_events (one per frame batch)
int stream_frames_in(kfn_handle* h, const uint8_t* host_images, uint8_t* dev_images, int64_t F, int frames_per_batch, cudaStream_t st) {
  const int64_t img_bytes = (int64_t)h->H * h->W * 3;
  const int nb = (int)((F + frames_per_batch - 1) / frames_per_batch);
  if (!h->copy_stream) KFN_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  if (!h->compute_done) KFN_CUDA_OK(cudaEventCreateWithFlags(&h->compute_done, cudaEventDisableTiming));
  while ((int)h->copy_events.size() < nb) {
    cudaEvent_t e;
    KFN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->copy_events.push_back(e);
  }
  // the staging buffer may still be read by work already queued on `st` (previous call): order the copies after it
  KFN_CUDA_OK(cudaEventRecord(h->compute_done, st));
  KFN_CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->compute_done, 0));
  for (int b = 0; b < nb; ++b) {
    const int64_t f0 = (int64_t)b * frames_per_batch;
    const int64_t nfr = (F - f0) < frames_per_batch ? (F - f0) : frames_per_batch;
    KFN_CUDA_OK(cudaMemcpyAsync(dev_images + f0 * img_bytes, host_images + f0 * img_bytes, (size_t)(nfr * img_bytes),
                                cudaMemcpyHostToDevice, h->copy_stream));
    KFN_CUDA_OK(cudaEventRecord(h->copy_events[b], h->copy_stream));
  }
  return KFN_OK;
}
}  // namespace

extern "C" {

int kfn_eval_sequences(kfn_handle* h, const uint8_t* images, int S, int T, const float* transform, int reinit_period,
                       int nis_gate, int frames_per_batch, int pairs_per_batch, float* out_map, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(images && transform && out_map && workspace, "null argument");
  KFN_REQUIRE(S > 0 && T > 0 && frames_per_batch > 0 && pairs_per_batch > 0, "empty problem");
  return eval_sequences_impl(h, images, S, T, transform, reinit_period, nis_gate, frames_per_batch, pairs_per_batch, out_map,
                             workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}

int kfn_eval_sequences_host(kfn_handle* h, const uint8_t* host_images, int S, int T, const float* transform,
                            int reinit_period, int nis_gate, int frames_per_batch, int pairs_per_batch, float* host_out,
                            uint8_t* dev_images, float* dev_out, void* workspace, int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(host_images && host_out && dev_images && dev_out && transform && workspace, "null argument");
  KFN_REQUIRE(S > 0 && T > 0 && frames_per_batch > 0 && pairs_per_batch > 0, "empty problem");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t F = (int64_t)S * T;
  int rc = stream_frames_in(h, host_images, dev_images, F, frames_per_batch, st);
  if (rc != KFN_OK) return rc;
  rc = eval_sequences_impl(h, dev_images, S, T, transform, reinit_period, nis_gate, frames_per_batch, pairs_per_batch,
                           dev_out, workspace, workspace_bytes, st, h->copy_events.data());
  if (rc != KFN_OK) return rc;
  KFN_CUDA_OK(cudaMemcpyAsync(host_out, dev_out, (size_t)(F * h->GH * h->GW * 4 * sizeof(float)), cudaMemcpyDeviceToHost, st));
  KFN_CUDA_OK(cudaStreamSynchronize(st));
  return KFN_OK;
}

int kfn_eval_stream_host(kfn_handle* h, const uint8_t* host_images, int T, int first_index, const float* transform,
                         int reinit_period, int nis_gate, int frames_per_batch, int pairs_per_batch, float* carry,
                         float* host_out, float* host_diag, uint8_t* dev_images, float* dev_out, float* dev_diag,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(host_images && host_out && dev_images && dev_out && transform && workspace && carry, "null argument");
  KFN_REQUIRE(T > 0 && first_index >= 0 && frames_per_batch > 0 && pairs_per_batch > 0, "empty problem");
  KFN_REQUIRE((host_diag == nullptr) == (dev_diag == nullptr), "host_diag and dev_diag go together");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)h->GH * h->GW;
  int rc = stream_frames_in(h, host_images, dev_images, T, frames_per_batch, st);
  if (rc != KFN_OK) return rc;
  StreamExtras ex;
  ex.first_index = first_index; ex.carry = carry;
  if (dev_diag) {   // planar [measure T*P*4 | temp T*P*4 | kf T*P*4 | nis T*P*3]
    ex.measure_out = dev_diag; ex.temp = dev_diag + (int64_t)T * P * 4; ex.kf = dev_diag + (int64_t)T * P * 8;
    ex.nis = dev_diag + (int64_t)T * P * 12;
  }
  rc = eval_sequences_impl(h, dev_images, 1, T, transform, reinit_period, nis_gate, frames_per_batch, pairs_per_batch, dev_out,
                           workspace, workspace_bytes, st, h->copy_events.data(), &ex);
  if (rc != KFN_OK) return rc;
  KFN_CUDA_OK(cudaMemcpyAsync(host_out, dev_out, (size_t)(T * P * 4 * sizeof(float)), cudaMemcpyDeviceToHost, st));
  if (dev_diag) KFN_CUDA_OK(cudaMemcpyAsync(host_diag, dev_diag, (size_t)(T * P * 15 * sizeof(float)), cudaMemcpyDeviceToHost, st));
  KFN_CUDA_OK(cudaStreamSynchronize(st));
  return KFN_OK;
}

int kfn_conv_layer(kfn_handle* h, const char* op_key, const void* in, int N, int IH, int IW, int in_cs, void* out, int OH,
                   int OW, int out_cs, int out_coff, int out_fp32, int epi, void* stream) {
  KFN_READY(h);
  KFN_REQUIRE(op_key && in && out && N > 0, "null/empty argument");
  std::string key(op_key);
  if (key == "ScoreNet/conv1ab#u8") {   // conv1a -> conv1b fused (conv1ab_fused.cu); out = conv1b's output, bf16 [N,IH,IW,64]
    KFN_REQUIRE(in_cs == 3 && OH == IH && OW == IW && out_cs == 64 && out_coff == 0 && !out_fp32 && epi == 0 && conv1ab_shape_ok(IH, IW),
                "fused conv1a+conv1b: u8 [N,IH,IW,3] in, bf16 [N,IH,IW,64] out, IH % 16 == 0, IW % 8 == 0");
    return run_conv1ab(h, reinterpret_cast<const uint8_t*>(in), N, IH, IW, out, nullptr, (cudaStream_t)stream);
  }
  if (key.size() > 3 && key.compare(key.size() - 3, 3, "#u8") == 0) {   // first layer straight from the u8 image [N,IH,IW,3]
    key.resize(key.size() - 3);
    KFN_REQUIRE(in_cs == 3 && OH == IH && OW == IW, "u8 first layer: in is the [N,IH,IW,3] image, same-size output");
    return run_op(h, key, act(const_cast<void*>(in), N, IH, IW, 32, 1), act(out, N, OH, OW, out_cs), out_coff, OH, OW, 1, 0, 0, epi,
                  out_fp32 != 0, (cudaStream_t)stream, nullptr, 0, 0, true);
  }
  Act a = act(const_cast<void*>(in), N, IH, IW, in_cs);
  Act o = act(out, N, OH, OW, out_cs);
  if (key.size() > 4 && key.compare(key.size() - 4, 4, "#all") == 0) {   // merged conv2d_transpose (all four phases)
    KFN_REQUIRE(OH == 2 * IH && OW == 2 * IW, "bad conv2d_transpose shape");
    return run_op(h, key, a, o, out_coff, IH, IW, 2, 0, 0, epi, out_fp32 != 0, (cudaStream_t)stream);
  }
  const size_t hash = key.find("#p");
  if (hash != std::string::npos) {
    const int ph = key[hash + 2] - '0';
    KFN_REQUIRE(ph >= 0 && ph < 4 && OH == 2 * IH && OW == 2 * IW, "bad conv2d_transpose phase / shape");
    return run_op(h, key, a, o, out_coff, IH, IW, 2, ph >> 1, ph & 1, epi, out_fp32 != 0, (cudaStream_t)stream);
  }
  return run_op(h, key, a, o, out_coff, OH, OW, 1, 0, 0, epi, out_fp32 != 0, (cudaStream_t)stream);
}

int64_t kfn_launch_count(kfn_handle* h, int reset) {
  // with a handle: kernels launched by that handle's entry points; with NULL: everything this thread launched through
  // the library, handle-less entry points included
  int64_t& c = h ? h->launches : kfn::g_launches;
  const int64_t v = c;
  if (reset) c = 0;
  return v;
}

}  // extern "C"
