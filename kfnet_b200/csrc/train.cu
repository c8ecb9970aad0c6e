// Training path of the KFNet hot path in fp32 on CUDA cores: the 4-frame unrolled filter of
// KFNet.GetKFCoordBatch (KFNet/KFNet.py:102-146), the three losses of KFNet/train.py:287-296 (MeasureCoordLoss /
// TemporalCoordLoss / KFCoordLoss = CoordLossWithUncertainty + 50 x SmoothLoss, KFNet.py:192-298,430-467, mixed
// 0.2 / 0.2 / 0.6), the backward pass of every stage, the L2 regulariser over the trainable variables and one
// tf.train.AdamOptimizer step with exponential learning-rate decay (train.py:300-315).
//
//   forward   : the same conv_simt_kernel as KFN_MODE_FP32 inference, every layer output kept
//   loss      : train_loss_kernel (forward scalars + gradients w.r.t. the three (coord, uncertainty) map sets)
//   filter    : filter_backward_kernel, one launch per unrolled step t = 3, 2, 1: Kalman update, process noise, bilinear
//               warp (scatter into the previous posterior's gradient), flow
//   soft-argmax / sigma_w : softmax_flow_backward_kernel;  difference volume: diff_volume_backward_kernel
//   conv layers: relu mask -> bias gradient (column sums) -> weight gradient (wgrad_kernel: split-K over pixels,
//               deterministic two-pass reduction) -> data gradient = conv_simt_kernel on re-packed weights:
//               stride-1 conv  -> conv with mirrored taps and transposed weights,
//               stride-2 conv  -> the four sub-pixel phases of a conv2d_transpose (App. A.3) over the same kernel tensor,
//               conv2d_transpose -> stride-2 conv over the same kernel tensor
//   optimiser : adam_kernel (g + weight_decay * w, TF-1.13 "epsilon hat" form)
// Parameters live on the device in the TF variable layouts (HWIO / [kh,kw,Cout,Cin] / [in,out]); the GEMM operand layouts
// of the forward and data-gradient ops are rebuilt from them by repack_kernel at the start of every step.
// Parity tests: tests/test_gpu_training.py (every gradient against fp64 autograd of the CPU restatement).
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "handle.h"
#include "layers.h"

namespace kfn {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// operand re-packing: dst[(t * A + a) * BP + b] = src[tapoff[t] + a * sa + b * sb]   (b >= B: 0)
struct PackDesc {
  int64_t dst_off;      // floats, into TrainState::packed
  int64_t src_off;      // floats, into TrainState::params (the variable's kernel tensor)
  int ntaps, A, B, BP;
  int tapoff[9];
  int sa, sb;
};
struct GemmOp {         // one conv_simt launch
  PackDesc pk;
  int dy[9], dx[9];
  int istride;          // input pixel = row * istride + tap offset
  int omul, oaddy, oaddx;   // output pixel = row * omul + oadd (phases)
};
struct LayerOps {
  std::vector<GemmOp> fwd;    // 1, or 4 (conv2d_transpose phases)
  std::vector<GemmOp> dgrad;  // 1, or 4 (stride-2 conv: phases over the input grid)
};

__global__ void repack_kernel(const float* __restrict__ params, float* __restrict__ packed, const PackDesc d) {
  const int64_t n = (int64_t)d.ntaps * d.A * d.BP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % d.BP);
    const int64_t r = i / d.BP;
    const int a = (int)(r % d.A), t = (int)(r / d.A);
    packed[d.dst_off + i] = b < d.B ? params[d.src_off + d.tapoff[t] + (int64_t)a * d.sa + (int64_t)b * d.sb] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// tensor views: NHWC fp32, `cs` channel stride, channels [coff, coff + C) of the buffer
struct View { float* p; int N, H, W, cs, coff, C; };
View view(float* p, int N, int H, int W, int cs, int coff, int C) { return View{p, N, H, W, cs, coff, C}; }
int64_t rows_of(const View& v) { return (int64_t)v.N * v.H * v.W; }

// ------------------------------------------------------------------------------------------------------------------
// small elementwise kernels
__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ act, int64_t rows, int cs, int coff, int C) {
  const int64_t n = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C; const int c = (int)(i - r * C);
    const int64_t o = r * cs + coff + c;
    if (!(act[o] > 0.f)) d[o] = 0.f;
  }
}
__global__ void zero_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}
int zero(float* p, int64_t n, cudaStream_t st) {
  if (n <= 0) return KFN_OK;
  int64_t b = (n + 255) / 256; if (b > 148 * 32) b = 148 * 32;
  zero_kernel<<<(unsigned)b, 256, 0, st>>>(p, n);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// column sums of d[rows, cs] over channels [coff, coff+C): partial[split][c]
__global__ void colsum_kernel(const float* __restrict__ d, int64_t rows, int cs, int coff, int C, int64_t rows_per_split,
                              float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += d[r * cs + coff + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}
// out[i] = sum_s partial[s][i]  (fixed order: deterministic)
__global__ void reduce_splits_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient:  dW[t][a][b] = sum_rows IN[pix(row) * istride + d_t][a] * DOUT[row][b]
// CTA = 64 x 64 tile of one tap over one split of the rows; 256 threads, 4x4 register tile; rows in chunks of 16.
struct WgradParams {
  const float* in; int in_cs;          // IN[N, IH, IW, in_cs] (pointer already offset to the first channel)
  const float* dout; int dout_cs;      // DOUT[N, MH, MW, dout_cs]
  int N, MH, MW, IH, IW, istride;
  int A, B, ntaps;
  int dy[9], dx[9];
  int64_t rows, rows_per_split;
  float* partial;                      // [split][tap][A][B]
};
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradParams p) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tiles_b = (p.B + 63) / 64;
  const int a0 = (blockIdx.x / tiles_b) * 64, b0 = (blockIdx.x % tiles_b) * 64;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int dy = p.dy[tap], dx = p.dx[tap];
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = r_begin + p.rows_per_split < p.rows ? r_begin + p.rows_per_split : p.rows;
  const int tid = threadIdx.x;
  const int lr = tid >> 4;            // row of the chunk this thread stages
  const int lc = (tid & 15) * 4;      // first of its 4 channels
  const int ta = (tid >> 4) * 4, tb = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r = r_begin; r < r_end; r += 16) {
    const int64_t row = r + lr;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < r_end) {
      const int mx = (int)(row % p.MW);
      const int my = (int)((row / p.MW) % p.MH);
      const int n = (int)(row / ((int64_t)p.MW * p.MH));
      const int iy = my * p.istride + dy, ix = mx * p.istride + dx;
      if (iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW) {
        const float* src = p.in + (((int64_t)n * p.IH + iy) * p.IW + ix) * p.in_cs + a0 + lc;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (a0 + lc + q < p.A) av[q] = __ldg(src + q);
      }
      const float* dsrc = p.dout + row * p.dout_cs + b0 + lc;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (b0 + lc + q < p.B) bv[q] = __ldg(dsrc + q);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[lr][lc + q] = av[q]; Bs[lr][lc + q] = bv[q]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ta + i]; b[i] = Bs[k][tb + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = p.partial + ((int64_t)split * p.ntaps + tap) * p.A * p.B;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (a0 + ta + i < p.A && b0 + tb + j < p.B) dst[(int64_t)(a0 + ta + i) * p.B + b0 + tb + j] = acc[i][j];
}

// ------------------------------------------------------------------------------------------------------------------
// tf.nn.l2_normalize backward (KFNet.py:340): y = x * r, r = rsqrt(max(|x|^2, 1e-12))  ->  dx = r * (dy - y (y . dy))
__global__ void l2norm_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t rows) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float xv[32], ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { xv[i] = x[r * 32 + i]; ss += xv[i] * xv[i]; }
  const bool clamped = !(ss > 1e-12f);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
  for (int i = 0; i < 32; ++i) dot += xv[i] * inv * dy[r * 32 + i];
#pragma unroll
  for (int i = 0; i < 32; ++i) dx[r * 32 + i] = clamped ? inv * dy[r * 32 + i] : inv * (dy[r * 32 + i] - xv[i] * inv * dot);
}

// difference volume backward (KFNet.py:343-359): diff[j, p, k, c] = f[j+1, p, c] - f[j, p + off_k, c], pairs j = 0..B-2
//   dfeat[f, q, c] = sum_k ddiff[f-1, q, k, c]  -  sum_k ddiff[f, q - off_k, k, c]
__global__ void diff_volume_backward_kernel(const float* __restrict__ ddiff, float* __restrict__ dfeat, int B, int GH, int GW) {
  const int P = GH * GW;
  const int64_t total = (int64_t)B * P * 32;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i & 31);
  const int q = (int)((i >> 5) % P);
  const int f = (int)((i >> 5) / P);
  const int qy = q / GW, qx = q - qy * GW;
  float s = 0.f;
  if (f >= 1) {
    const float* base = ddiff + (((int64_t)(f - 1) * P + q) * 64) * 32 + c;
    for (int k = 0; k < 64; ++k) s += base[k * 32];
  }
  if (f + 1 < B) {
    for (int k = 0; k < 64; ++k) {
      const int py = qy - ((k >> 3) - 4), px = qx - ((k & 7) - 4);     // p + off_k = q, off_k = (ox, oy) = (k%8 - 4, k/8 - 4)
      if (py >= 0 && py < GH && px >= 0 && px < GW) s -= ddiff[((((int64_t)f * P + py * GW + px) * 64) + k) * 32 + c];
    }
  }
  dfeat[i] = s;
}

// soft-argmax / sigma_w backward (OFlowNet.py:45-56, KFNet.py:381-385): one warp per window row
//   prob = softmax(logits), flow = sum_k prob_k off_k, sigma_w = 0.01 exp(u)
__global__ void softmax_flow_backward_kernel(const float* __restrict__ logits, const float* __restrict__ dflow,
                                             const float* __restrict__ dsigw, const float* __restrict__ sigw, int64_t rows,
                                             float* __restrict__ dlogits, float* __restrict__ dunc) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float2 l = reinterpret_cast<const float2*>(logits + row * 64)[lane];
  float m = fmaxf(l.x, l.y);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e0 = expf(l.x - m), e1 = expf(l.y - m);
  float s = e0 + e1;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float p0 = e0 / s, p1 = e1 / s;
  const int k0 = 2 * lane, k1 = k0 + 1;
  const float gx = dflow[row * 2], gy = dflow[row * 2 + 1];
  const float dp0 = gx * (float)((k0 & 7) - 4) + gy * (float)((k0 >> 3) - 4);
  const float dp1 = gx * (float)((k1 & 7) - 4) + gy * (float)((k1 >> 3) - 4);
  float dot = p0 * dp0 + p1 * dp1;
#pragma unroll
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  reinterpret_cast<float2*>(dlogits + row * 64)[lane] = make_float2(p0 * (dp0 - dot), p1 * (dp1 - dot));
  if (lane == 0) dunc[row] = dsigw[row] * sigw[row];
}

// ------------------------------------------------------------------------------------------------------------------
// Losses (KFNet.py:192-232, 430-467; train.py:287-296).  One CTA per loss kind k (0 measure, 1 temporal, 2 KF); fixed
// reduction order.  est[k] = [B, P, 4] maps (xyz, sigma).  Writes G[k] = mix_k * d(loss_k)/d(est[k]) and the scalars
// scal[k*2] = loss_k (coord + 50 smooth), scal[k*2+1] = accuracy_k.
struct LossParams {
  const float* est[3];
  float* G[3];
  const float* gt;        // [B, P, 3]
  const float* mask;      // [B, P, 1] in {0,1}
  const uint8_t* images;  // [B, H, W, 3] (resize_nearest_neighbor /8 taken on the fly, train.py:274)
  int B, GH, GW, H, W;
  float mix[3];
  float* scal;            // [6]
};
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__global__ void __launch_bounds__(1024) train_loss_kernel(const LossParams p) {
  __shared__ float red[33];
  const int k = blockIdx.x;
  const float4* est = reinterpret_cast<const float4*>(p.est[k]);
  float4* G = reinterpret_cast<float4*>(p.G[k]);
  const int P = p.GH * p.GW, n = p.B * P;
  float ms = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ms += p.mask[i];
  const float valid = block_sum(ms, red) + 1.0f;
  const int sy = p.H / p.GH, sx = p.W / p.GW;
  auto img_w = [&](int b, int y0, int x0, int y1, int x1) -> float {     // exp(-0.625 mean_c |I(a) - I(b)|)
    const uint8_t* a = p.images + (((int64_t)b * p.H + y0 * sy) * p.W + x0 * sx) * 3;
    const uint8_t* c = p.images + (((int64_t)b * p.H + y1 * sy) * p.W + x1 * sx) * 3;
    const float d = (fabsf((float)a[0] - (float)c[0]) + fabsf((float)a[1] - (float)c[1]) + fabsf((float)a[2] - (float)c[2])) / 3.0f;
    return expf(-0.625f * d);
  };
  float loss_s = 0.f, smooth_s = 0.f, nz = 0.f;
  const float gscale = p.mix[k] / valid;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / P, q = i - b * P, y = q / p.GW, x = q - y * p.GW;
    const float4 e = est[i];
    const float m = p.mask[i];
    const float dxv = e.x - p.gt[i * 3], dyv = e.y - p.gt[i * 3 + 1], dzv = e.z - p.gt[i * 3 + 2];
    const float d = dxv * dxv + dyv * dyv + dzv * dzv;
    const float u = fmaxf(e.w, 1e-5f);
    const float raw = 3.0f * logf(u) + d / (2.0f * u * u);
    loss_s += m * fminf(raw, -2.0f);                       // sic: tf.minimum (KFNet.py:217)
    nz += (m * d - 0.05f * 0.05f > 0.f) ? 1.f : 0.f;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (raw <= -2.0f) {                                    // only pixels whose unclamped loss is below -2 carry gradient
      const float iu2 = 1.0f / (u * u);
      g.x = m * dxv * iu2; g.y = m * dyv * iu2; g.z = m * dzv * iu2;
      g.w = (e.w >= 1e-5f) ? m * (3.0f / u - d * iu2 / u) : 0.f;
    }
    // SmoothLoss: sx[x] = mean_c (c[x] - c[x+1])^2 * wx[x] * mask[x] (x < GW-1), same in y; weight 50
    float hx = 0.f, hy = 0.f, hz = 0.f;
    if (x + 1 < p.GW) {
      const float4 r = est[i + 1];
      const float w = img_w(b, y, x, y, x + 1) * m;
      const float ax = e.x - r.x, ay = e.y - r.y, az = e.z - r.z;
      smooth_s += (ax * ax + ay * ay + az * az) / 3.0f * w;
      hx += ax * w; hy += ay * w; hz += az * w;
    }
    if (x > 0) {
      const float4 l = est[i - 1];
      const float w = img_w(b, y, x - 1, y, x) * p.mask[i - 1];
      hx -= (l.x - e.x) * w; hy -= (l.y - e.y) * w; hz -= (l.z - e.z) * w;
    }
    if (y + 1 < p.GH) {
      const float4 r = est[i + p.GW];
      const float w = img_w(b, y, x, y + 1, x) * m;
      const float ax = e.x - r.x, ay = e.y - r.y, az = e.z - r.z;
      smooth_s += (ax * ax + ay * ay + az * az) / 3.0f * w;
      hx += ax * w; hy += ay * w; hz += az * w;
    }
    if (y > 0) {
      const float4 l = est[i - p.GW];
      const float w = img_w(b, y - 1, x, y, x) * p.mask[i - p.GW];
      hx -= (l.x - e.x) * w; hy -= (l.y - e.y) * w; hz -= (l.z - e.z) * w;
    }
    const float sm = 50.0f * 2.0f / 3.0f;
    g.x = (g.x + sm * hx) * gscale; g.y = (g.y + sm * hy) * gscale; g.z = (g.z + sm * hz) * gscale; g.w *= gscale;
    G[i] = g;
  }
  const float L = block_sum(loss_s, red);
  const float S = block_sum(smooth_s, red);
  const float NZ = block_sum(nz, red);
  if (threadIdx.x == 0) {
    p.scal[k * 2] = L / valid + 50.0f * S / valid;
    p.scal[k * 2 + 1] = (valid - NZ) / valid;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// One unrolled filter step backwards (KFNet.py:127-143 -> BuildOFlowNet tail :386-401, BuildKFCoord :148-162).
//   prev  = posterior of step t-1 (xyz, sigma) [P,4];  z = measurement t;  G_temp / G_kf = gradients w.r.t. (temp, KF) of step t
//   out: G_prev += (scatter through the bilinear taps, atomics), dflow [P,2], dsigw [P], G_meas_t += (z, sigma_z)
__global__ void filter_backward_kernel(const float4* __restrict__ prev, const float2* __restrict__ flow, const float* __restrict__ sigw,
                                       const float4* __restrict__ z4, const float4* __restrict__ G_temp, const float4* __restrict__ G_kf,
                                       float* __restrict__ G_prev, float2* __restrict__ dflow, float* __restrict__ dsigw,
                                       float4* __restrict__ G_meas, int GH, int GW) {
  const int P = GH * GW;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float minvar = 1e-5f * 1e-5f;
  const float xmax = (float)(GW - 1), ymax = (float)(GH - 1);
  const float2 fl = flow[p];
  const int y = p / GW, x = p - y * GW;
  const float cx = (float)x + fl.x, cy = (float)y + fl.y;
  const float x0 = floorf(cx), y0 = floorf(cy);
  const float x0s = fminf(fmaxf(x0, 0.f), xmax), x1s = fminf(fmaxf(x0 + 1.0f, 0.f), xmax);
  const float y0s = fminf(fmaxf(y0, 0.f), ymax), y1s = fminf(fmaxf(y0 + 1.0f, 0.f), ymax);
  const float wx0 = x1s - cx, wx1 = cx - x0s, wy0 = y1s - cy, wy1 = cy - y0s;
  const float w[4] = {wx0 * wy0, wx0 * wy1, wx1 * wy0, wx1 * wy1};
  const int idx[4] = {(int)y0s * GW + (int)x0s, (int)y1s * GW + (int)x0s, (int)y0s * GW + (int)x1s, (int)y1s * GW + (int)x1s};
  float4 tap[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tap[i] = prev[idx[i]];
  float xm[3] = {0.f, 0.f, 0.f}, sl = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { xm[0] += w[i] * tap[i].x; xm[1] += w[i] * tap[i].y; xm[2] += w[i] * tap[i].z; sl += w[i] * tap[i].w; }
  const float sw = sigw[p];
  const float sl2 = sl * sl, sw2 = sw * sw;
  const float last_var = fmaxf(sl2, minvar), trans_var = fmaxf(sw2, minvar);
  const float tu = sqrtf(trans_var + last_var);
  const float4 z = z4[p];
  const float lv = tu * tu, mv = z.w * z.w;
  const float den = lv + mv;
  const float K = lv / den;
  const float omk_raw = 1.0f - K;
  const float omk = fmaxf(omk_raw, 0.f);
  const float ku = sqrtf(omk * lv);
  const float4 gk = G_kf[p], gt = G_temp[p];
  // kc = omk * xm + K * z
  float d_xm[3] = {omk * gk.x + gt.x, omk * gk.y + gt.y, omk * gk.z + gt.z};
  float d_z[3] = {K * gk.x, K * gk.y, K * gk.z};
  float d_omk = gk.x * xm[0] + gk.y * xm[1] + gk.z * xm[2];
  float d_K = gk.x * z.x + gk.y * z.y + gk.z * z.z;
  // ku = sqrt(omk * lv)
  float d_lv = 0.f;
  if (ku > 0.f) { const float h = gk.w / (2.0f * ku); d_omk += h * lv; d_lv += h * omk; }
  if (omk_raw >= 0.f) d_K -= d_omk;
  // K = lv / (lv + mv)
  d_lv += d_K * mv / (den * den);
  const float d_mv = -d_K * lv / (den * den);
  const float d_sz = 2.0f * z.w * d_mv;
  // lv = tu^2, tu also feeds the temporal loss
  const float d_tu = gt.w + 2.0f * tu * d_lv;
  const float h2 = d_tu / (2.0f * tu);
  const float d_sw = (sw2 >= minvar) ? 2.0f * sw * h2 : 0.f;
  const float d_sl = (sl2 >= minvar) ? 2.0f * sl * h2 : 0.f;
  // warp: scatter into the previous posterior, and the weights' dependence on the sampling point
  float dw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dw[i] = d_xm[0] * tap[i].x + d_xm[1] * tap[i].y + d_xm[2] * tap[i].z + d_sl * tap[i].w;
    float* g = G_prev + (int64_t)idx[i] * 4;
    atomicAdd(g + 0, w[i] * d_xm[0]); atomicAdd(g + 1, w[i] * d_xm[1]); atomicAdd(g + 2, w[i] * d_xm[2]); atomicAdd(g + 3, w[i] * d_sl);
  }
  const float d_wx0 = dw[0] * wy0 + dw[1] * wy1, d_wx1 = dw[2] * wy0 + dw[3] * wy1;
  const float d_wy0 = dw[0] * wx0 + dw[2] * wx1, d_wy1 = dw[1] * wx0 + dw[3] * wx1;
  dflow[p] = make_float2(d_wx1 - d_wx0, d_wy1 - d_wy0);
  dsigw[p] = d_sw;
  float4 gm = G_meas[p];
  gm.x += d_z[0]; gm.y += d_z[1]; gm.z += d_z[2]; gm.w += d_sz;
  G_meas[p] = gm;
}
// frame 0 of the window IS the measurement (KFNet.py:112-118): its temp / KF gradients flow straight into the measurement;
// also turns d/d(sigma_z) into d/d(logit) (sigma_z = exp(logit), SCoordNet.py:43) for all frames
__global__ void measure_grad_finish_kernel(float4* __restrict__ G_meas, const float4* __restrict__ G_temp0, const float4* __restrict__ G_kf0,
                                           const float4* __restrict__ measure, int P, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 g = G_meas[i];
  if (i < P) {
    const float4 a = G_temp0[i], b = G_kf0[i];
    g.x += a.x + b.x; g.y += a.y + b.y; g.z += a.z + b.z; g.w += a.w + b.w;
  }
  g.w *= measure[i].w;
  G_meas[i] = g;
}

// ------------------------------------------------------------------------------------------------------------------
// tf.train.AdamOptimizer (TF-1.13): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps)
// with g = dLoss/dp + weight_decay * p  (l2_regularizer(scale) = scale * sum(p^2) / 2, train.py:301-303)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float wd, float lr_t, float b1, float b2, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] + wd * p[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void sumsq_kernel(const float* __restrict__ p, int64_t n, float* __restrict__ partial) {
  __shared__ float red[33];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += p[i] * p[i];
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void preprocess_train_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ((float)img[i] - 128.0f) * 0.00625f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
struct TrainState {
  int B = 4, H = 0, W = 0, GH = 0, GW = 0;
  bool train_scoordnet = true, train_oflownet = true;
  int64_t n_params = 0, n_scorenet = 0;           // ScoreNet variables come first in the flat buffers
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr, *packed = nullptr, *zeros = nullptr;
  int64_t w_off[kNumLayers], b_off[kNumLayers];
  LayerOps ops[kNumLayers];
  int64_t packed_elems = 0;
  int step = 0;                                    // global_step (train.py:305): learning-rate decay exponent
  int adam_t = 0;                                  // updates applied since kfn_train_begin (AdamOptimizer's beta^t accumulators start fresh)
  // activations / gradients / scratch (one cudaMalloc)
  float* arena = nullptr;
  int64_t arena_elems = 0;
  float* partial = nullptr;                        // wgrad / colsum split partials
  int64_t partial_elems = 0;
  float* scal = nullptr;                           // device [8]
  float* sumsq_partial = nullptr;                  // [2][1024]
  std::vector<void*> allocs;
};

void train_state_free(TrainState* t) {
  if (!t) return;
  for (void* p : t->allocs) cudaFree(p);
  delete t;
}

namespace {

int dmalloc(TrainState* t, float** p, int64_t elems) {
  KFN_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), (size_t)(elems > 0 ? elems : 1) * sizeof(float)));
  t->allocs.push_back(*p);
  return KFN_OK;
}

// builds the forward / data-gradient GEMM ops of layer li and reserves their packed operand storage
void build_ops(TrainState* t, int li) {
  const LayerSpec& L = kLayers[li];
  LayerOps& lo = t->ops[li];
  auto reserve = [&](GemmOp& op) {
    op.pk.dst_off = t->packed_elems;
    t->packed_elems += (int64_t)op.pk.ntaps * op.pk.A * op.pk.BP;
    t->packed_elems = (t->packed_elems + 3) & ~int64_t(3);      // keep every operand 16-byte aligned (float4 loads)
  };
  auto base = [&](GemmOp& op, int A, int B, int sa, int sb) {
    memset(&op, 0, sizeof(op));
    op.pk.src_off = t->w_off[li]; op.pk.A = A; op.pk.B = B; op.pk.BP = (B + 3) / 4 * 4; op.pk.sa = sa; op.pk.sb = sb;
    op.istride = 1; op.omul = 1;
  };
  const int cin = L.cin, cout = L.cout;
  if (L.kind == DENSE || (L.kind == CONV && L.k == 1)) {
    GemmOp f; base(f, cin, cout, cout, 1); f.pk.ntaps = 1; reserve(f); lo.fwd.push_back(f);
    GemmOp d; base(d, cout, cin, 1, cout); d.pk.ntaps = 1; reserve(d); lo.dgrad.push_back(d);
    return;
  }
  if (L.kind == CONV) {
    const int pad = L.stride == 1 ? 1 : 0;
    GemmOp f; base(f, cin, cout, cout, 1); f.pk.ntaps = 9; f.istride = L.stride;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int tix = ky * 3 + kx;
        f.pk.tapoff[tix] = tix * cin * cout; f.dy[tix] = ky - pad; f.dx[tix] = kx - pad;
      }
    reserve(f); lo.fwd.push_back(f);
    if (L.stride == 1) {     // dx[i] = sum_t dy[i - d_t] W_t^T
      GemmOp d; base(d, cout, cin, 1, cout); d.pk.ntaps = 9;
      for (int tix = 0; tix < 9; ++tix) { d.pk.tapoff[tix] = tix * cin * cout; d.dy[tix] = -(tix / 3 - 1); d.dx[tix] = -(tix % 3 - 1); }
      reserve(d); lo.dgrad.push_back(d);
    } else {                 // input pixel 2m+p collects dy[m - (k-p)/2] W_k for k = p (mod 2): conv2d_transpose phases
      for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        GemmOp d; base(d, cout, cin, 1, cout); d.omul = 2; d.oaddy = py; d.oaddx = px;
        const int kys[2] = {py ? 1 : 0, 2}, dys[2] = {0, -1}, kxs[2] = {px ? 1 : 0, 2}, dxs[2] = {0, -1};
        int nt = 0;
        for (int a = 0; a < (py ? 1 : 2); ++a)
          for (int c = 0; c < (px ? 1 : 2); ++c) {
            d.pk.tapoff[nt] = (kys[a] * 3 + kxs[c]) * cin * cout; d.dy[nt] = dys[a]; d.dx[nt] = dxs[c]; ++nt;
          }
        d.pk.ntaps = nt; reserve(d); lo.dgrad.push_back(d);
      }
    }
    return;
  }
  // DECONV (conv2d_transpose k3 s2 SAME, kernel [kh,kw,Cout,Cin]; App. A.3: out[2i + k] += x[i] W_k, cropped)
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    GemmOp f; base(f, cin, cout, 1, cin); f.omul = 2; f.oaddy = py; f.oaddx = px;
    const int kys[2] = {py ? 1 : 0, 2}, dys[2] = {0, -1}, kxs[2] = {px ? 1 : 0, 2}, dxs[2] = {0, -1};
    int nt = 0;
    for (int a = 0; a < (py ? 1 : 2); ++a)
      for (int c = 0; c < (px ? 1 : 2); ++c) {
        f.pk.tapoff[nt] = (kys[a] * 3 + kxs[c]) * cin * cout; f.dy[nt] = dys[a]; f.dx[nt] = dxs[c]; ++nt;
      }
    f.pk.ntaps = nt; reserve(f); lo.fwd.push_back(f);
  }
  GemmOp d; base(d, cout, cin, cin, 1); d.pk.ntaps = 9; d.istride = 2;      // dx[i] = sum_k dy[2i + k] W_k
  for (int tix = 0; tix < 9; ++tix) { d.pk.tapoff[tix] = tix * cin * cout; d.dy[tix] = tix / 3; d.dx[tix] = tix % 3; }
  reserve(d); lo.dgrad.push_back(d);
}

struct Ctx {
  kfn_handle* h; TrainState* t; cudaStream_t st;
};

// one conv_simt launch of `op`: in view (channels [coff, coff + A)) -> out view (channels [coff, coff + B)) over the row grid MH x MW
int run_gemm(Ctx& c, const GemmOp& op, const View& in, const View& out, int MH, int MW, const float* bias, bool relu, int epi, bool accum) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = in.N; g.MH = MH; g.MW = MW;
  g.IH = in.H; g.IW = in.W; g.in_cs = in.cs; g.Cin = op.pk.A;
  g.istride = op.istride; g.ntaps = op.pk.ntaps;
  for (int i = 0; i < op.pk.ntaps; ++i) { g.dy[i] = op.dy[i]; g.dx[i] = op.dx[i]; }
  g.OH = out.H; g.OW = out.W; g.out_cs = out.cs; g.out_coff = out.coff; g.Cout = op.pk.B;
  g.omul = op.omul; g.oaddy = op.oaddy; g.oaddx = op.oaddx; g.relu = relu ? 1 : 0; g.epi = epi; g.accum = accum ? 1 : 0;
  return launch_conv_simt(g, in.p + in.coff, c.t->packed + op.pk.dst_off, op.pk.BP, bias ? bias : c.t->zeros, out.p, c.st);
}

int layer_forward(Ctx& c, int li, const View& in, const View& out, int epi = EPI_NONE) {
  const LayerSpec& L = kLayers[li];
  const float* bias = c.t->params + c.t->b_off[li];
  int rc;
  if (L.kind == DECONV) {
    for (const GemmOp& op : c.t->ops[li].fwd)
      if ((rc = run_gemm(c, op, in, out, in.H, in.W, bias, L.relu, epi, false))) return rc;
    return KFN_OK;
  }
  return run_gemm(c, c.t->ops[li].fwd[0], in, out, out.H, out.W, bias, L.relu, epi, false);
}

int launch_wgrad(Ctx& c, const float* in, int in_cs, int N, int IH, int IW, const float* dout, int dout_cs, int MH, int MW, int istride,
                 int A, int B, int ntaps, const int* dy, const int* dx, float* dW) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.in = in; p.in_cs = in_cs; p.dout = dout; p.dout_cs = dout_cs; p.N = N; p.MH = MH; p.MW = MW; p.IH = IH; p.IW = IW; p.istride = istride;
  p.A = A; p.B = B; p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) { p.dy[i] = dy[i]; p.dx[i] = dx[i]; }
  p.rows = (int64_t)N * MH * MW;
  const int tiles = ((A + 63) / 64) * ((B + 63) / 64);
  int64_t splits = (148 * 4 + (int64_t)tiles * ntaps - 1) / ((int64_t)tiles * ntaps);
  const int64_t max_by_rows = (p.rows + 511) / 512;                      // at least ~512 rows per split
  if (splits > max_by_rows) splits = max_by_rows;
  const int64_t max_by_mem = c.t->partial_elems / ((int64_t)ntaps * A * B);
  if (splits > max_by_mem) splits = max_by_mem;
  if (splits < 1) splits = 1;
  p.rows_per_split = ((p.rows + splits - 1) / splits + 15) / 16 * 16;
  splits = (p.rows + p.rows_per_split - 1) / p.rows_per_split;
  KFN_REQUIRE((int64_t)ntaps * A * B * splits <= c.t->partial_elems, "wgrad partial buffer too small");
  p.partial = c.t->partial;
  wgrad_kernel<<<dim3((unsigned)tiles, (unsigned)ntaps, (unsigned)splits), 256, 0, c.st>>>(p);
  KFN_LAUNCH_OK();
  const int64_t n = (int64_t)ntaps * A * B;
  int64_t blocks = (n + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
  reduce_splits_kernel<<<(unsigned)blocks, 256, 0, c.st>>>(c.t->partial, (int)splits, n, dW);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

int launch_colsum(Ctx& c, const View& d, float* out) {
  const int64_t rows = rows_of(d);
  int64_t splits = (rows + 2047) / 2048;
  if (splits > 512) splits = 512;
  const int64_t rps = (rows + splits - 1) / splits;
  splits = (rows + rps - 1) / rps;
  KFN_REQUIRE(splits * d.C <= c.t->partial_elems, "colsum partial buffer too small");
  colsum_kernel<<<dim3((unsigned)((d.C + 31) / 32), (unsigned)splits), dim3(32, 8), 0, c.st>>>(d.p, rows, d.cs, d.coff, d.C, rps, c.t->partial);
  KFN_LAUNCH_OK();
  reduce_splits_kernel<<<(unsigned)((d.C + 255) / 256), 256, 0, c.st>>>(c.t->partial, (int)splits, d.C, out);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// backward of layer li:  in (the layer's input activations), out (its output activations), dout (gradient w.r.t. the
// output, same layout as out; masked in place), din (gradient w.r.t. the input, or p == nullptr), accumulate into din?
int layer_backward(Ctx& c, int li, const View& in, const View& out, const View& dout, const View* din, bool accum) {
  const LayerSpec& L = kLayers[li];
  TrainState* t = c.t;
  int rc;
  if (L.relu) {
    const int64_t n = rows_of(dout) * dout.C;
    int64_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
    relu_mask_kernel<<<(unsigned)blocks, 256, 0, c.st>>>(dout.p, out.p, rows_of(dout), dout.cs, dout.coff, dout.C);
    KFN_LAUNCH_OK();
  }
  const bool trainable = (li < 12) ? t->train_scoordnet : t->train_oflownet;
  if (trainable) {
    if ((rc = launch_colsum(c, dout, t->grads + t->b_off[li]))) return rc;
    if (L.kind == DECONV) {     // dW[k][co][ci] = sum_i dy[2i + k][co] x[i][ci]
      const GemmOp& d = t->ops[li].dgrad[0];
      if ((rc = launch_wgrad(c, dout.p + dout.coff, dout.cs, dout.N, dout.H, dout.W, in.p + in.coff, in.cs, in.H, in.W, 2, L.cout, L.cin, 9,
                             d.dy, d.dx, t->grads + t->w_off[li])))
        return rc;
    } else {
      const GemmOp& f = t->ops[li].fwd[0];
      if ((rc = launch_wgrad(c, in.p + in.coff, in.cs, in.N, in.H, in.W, dout.p + dout.coff, dout.cs, dout.H, dout.W, f.istride, L.cin, L.cout,
                             f.pk.ntaps, f.dy, f.dx, t->grads + t->w_off[li])))
        return rc;
    }
  }
  if (din && din->p) {
    const std::vector<GemmOp>& dg = t->ops[li].dgrad;
    if (L.kind == CONV && L.stride == 2) {
      for (const GemmOp& op : dg)
        if ((rc = run_gemm(c, op, dout, *din, dout.H, dout.W, nullptr, false, EPI_NONE, accum))) return rc;
    } else {
      if ((rc = run_gemm(c, dg[0], dout, *din, din->H, din->W, nullptr, false, EPI_NONE, accum))) return rc;
    }
  }
  return KFN_OK;
}

struct Bump {
  float* base; int64_t off = 0, cap;
  Bump(float* b, int64_t c) : base(b), cap(c) {}
  float* get(int64_t n) { off = (off + 63) & ~int64_t(63); float* p = base ? base + off : nullptr; off += n; return p; }
};

// every activation / gradient buffer of one step; `dry` (base == nullptr) only measures
struct Bufs {
  float *x0, *s[12], *ds[12], *tw[7], *dtw[7], *feat, *dfeat;
  float *diff, *ddiff, *cat0, *dcat0, *c1a, *dc1a, *cat1, *dcat1, *c2a, *dc2a, *cat2, *dcat2, *c3a, *dc3a, *c3b, *dc3b, *c4, *dc4, *c5, *dc5,
      *c6, *dc6, *logits, *dlogits, *fc1, *dfc1, *fc2, *dfc2, *unc, *dunc;
  float *flow, *sigw, *dflow, *dsigw, *temp, *kf, *outmap, *G[3];
};
const int kSDiv[12] = {1, 1, 2, 2, 4, 4, 8, 8, 8, 8, 8, 8};
const int kTDiv[7] = {1, 2, 2, 4, 4, 8, 8};

void plan(TrainState* t, Bump& a, Bufs& b) {
  const int B = t->B, H = t->H, W = t->W;
  const int64_t P = (int64_t)t->GH * t->GW, R = (int64_t)(B - 1) * P;
  b.x0 = a.get((int64_t)B * H * W * 3);
  for (int i = 0; i < 12; ++i) {
    const int64_t n = (int64_t)B * (H / kSDiv[i]) * (W / kSDiv[i]) * kLayers[i].cout;
    b.s[i] = a.get(n); b.ds[i] = a.get(n);
  }
  for (int i = 0; i < 7; ++i) {
    const int64_t n = (int64_t)B * (H / kTDiv[i]) * (W / kTDiv[i]) * kLayers[12 + i].cout;
    b.tw[i] = a.get(n); b.dtw[i] = a.get(n);
  }
  b.feat = a.get(B * P * 32); b.dfeat = a.get(B * P * 32);
  auto pair = [&](float*& x, float*& dx, int64_t n) { x = a.get(n); dx = a.get(n); };
  pair(b.diff, b.ddiff, R * 64 * 32); pair(b.cat0, b.dcat0, R * 64 * 48); pair(b.c1a, b.dc1a, R * 16 * 32);
  pair(b.cat1, b.dcat1, R * 16 * 64); pair(b.c2a, b.dc2a, R * 4 * 64); pair(b.cat2, b.dcat2, R * 4 * 128);
  pair(b.c3a, b.dc3a, R * 128); pair(b.c3b, b.dc3b, R * 128); pair(b.c4, b.dc4, R * 4 * 64); pair(b.c5, b.dc5, R * 16 * 32);
  pair(b.c6, b.dc6, R * 64 * 16); pair(b.logits, b.dlogits, R * 64); pair(b.fc1, b.dfc1, R * 64); pair(b.fc2, b.dfc2, R * 32);
  pair(b.unc, b.dunc, R);
  pair(b.flow, b.dflow, B * P * 2); pair(b.sigw, b.dsigw, B * P);
  b.temp = a.get(B * P * 4); b.kf = a.get(B * P * 4); b.outmap = a.get(B * P * 4);
  for (int k = 0; k < 3; ++k) b.G[k] = a.get(B * P * 4);
}

int L_(const char* scope, const char* name) { return layer_index(std::string(scope) + "/" + name); }

}  // namespace
}  // namespace kfn

using namespace kfn;

// ====================================================================================================================
extern "C" {

static int train_begin_impl(kfn_handle* h, int batch, int train_scoordnet, int train_oflownet);

// a failed begin leaves NO training state behind (kfn_train_step on a half-built state would read unallocated buffers)
int kfn_train_begin(kfn_handle* h, int batch, int train_scoordnet, int train_oflownet) {
  const int rc = train_begin_impl(h, batch, train_scoordnet, train_oflownet);
  if (rc != KFN_OK && h && h->train) { train_state_free(h->train); h->train = nullptr; }
  return rc;
}

static int train_begin_impl(kfn_handle* h, int batch, int train_scoordnet, int train_oflownet) {
  KFN_REQUIRE(h, "null handle");
  KFN_REQUIRE(h->mode == KFN_MODE_FP32, "training runs in KFN_MODE_FP32 (the reference's arithmetic)");
  KFN_REQUIRE(batch >= 2 && batch <= 16, "training window of 2..16 frames (reference: 4)");
  KFN_CUDA_OK(cudaSetDevice(h->device));
  if (h->train) { train_state_free(h->train); h->train = nullptr; }
  TrainState* t = new TrainState();
  h->train = t;
  t->B = batch; t->H = h->H; t->W = h->W; t->GH = h->GH; t->GW = h->GW;
  t->train_scoordnet = train_scoordnet != 0; t->train_oflownet = train_oflownet != 0;
  int64_t off = 0;
  for (int i = 0; i < kNumLayers; ++i) {
    if (i == 12) t->n_scorenet = off;
    t->w_off[i] = off; off += layer_kernel_elems(kLayers[i]);
    t->b_off[i] = off; off += kLayers[i].cout;
    off = (off + 3) & ~int64_t(3);
  }
  t->n_params = off;
  for (int i = 0; i < kNumLayers; ++i) build_ops(t, i);
  int rc;
  if ((rc = dmalloc(t, &t->params, off)) || (rc = dmalloc(t, &t->grads, off)) || (rc = dmalloc(t, &t->adam_m, off)) ||
      (rc = dmalloc(t, &t->adam_v, off)) || (rc = dmalloc(t, &t->packed, t->packed_elems)) || (rc = dmalloc(t, &t->zeros, 4096)) ||
      (rc = dmalloc(t, &t->scal, 16)) || (rc = dmalloc(t, &t->sumsq_partial, 2048)))
    return rc;
  KFN_CUDA_OK(cudaMemset(t->params, 0, off * sizeof(float)));
  KFN_CUDA_OK(cudaMemset(t->grads, 0, off * sizeof(float)));
  KFN_CUDA_OK(cudaMemset(t->adam_m, 0, off * sizeof(float)));
  KFN_CUDA_OK(cudaMemset(t->adam_v, 0, off * sizeof(float)));
  KFN_CUDA_OK(cudaMemset(t->zeros, 0, 4096 * sizeof(float)));
  for (int i = 0; i < kNumLayers; ++i) {
    const std::string base = std::string(kLayers[i].scope) + "/" + kLayers[i].name;
    auto w = h->host_w.find(base + "/kernel"), b = h->host_w.find(base + "/bias");
    if (w == h->host_w.end() || b == h->host_w.end()) { set_error("kfn_train_begin: missing weight " + base); return KFN_ERR_STATE; }
    KFN_CUDA_OK(cudaMemcpy(t->params + t->w_off[i], w->second.data(), w->second.size() * sizeof(float), cudaMemcpyHostToDevice));
    KFN_CUDA_OK(cudaMemcpy(t->params + t->b_off[i], b->second.data(), b->second.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  Bump dry(nullptr, 0);
  Bufs tmp;
  plan(t, dry, tmp);
  t->arena_elems = dry.off + 64;
  if ((rc = dmalloc(t, &t->arena, t->arena_elems))) return rc;
  t->partial_elems = 48ll << 20;                  // 192 MB of split partials: >= 4 splits of the largest kernel (conv4b, 9.4 M)
  if ((rc = dmalloc(t, &t->partial, t->partial_elems))) return rc;
  return KFN_OK;
}

int kfn_train_step(kfn_handle* h, const uint8_t* images, const float* gt8, const float* mask8, float base_lr, float weight_decay,
                   int decay_steps, float gamma, int apply_update, float* scalars_host, void* stream) {
  KFN_REQUIRE(h && h->train, "call kfn_train_begin first");
  KFN_REQUIRE(images && gt8 && mask8, "null argument");
  KFN_CUDA_OK(cudaSetDevice(h->device));
  kfn::g_handle_launches = &h->launches;
  TrainState* t = h->train;
  Ctx c{h, t, (cudaStream_t)stream};
  cudaStream_t st = c.st;
  const int B = t->B, H = t->H, W = t->W, GH = t->GH, GW = t->GW;
  const int P = GH * GW;
  const int R = (B - 1) * P;
  Bump arena(t->arena, t->arena_elems);
  Bufs b;
  plan(t, arena, b);
  int rc;
  NvtxRange nv_step("kfn_train_step");
#define RUN(x) do { if ((rc = (x))) return rc; } while (0)
  // ---- operands of this step from the current parameters
  for (int li = 0; li < kNumLayers; ++li)
    for (const std::vector<GemmOp>* v : {&t->ops[li].fwd, &t->ops[li].dgrad})
      for (const GemmOp& op : *v) {
        const int64_t n = (int64_t)op.pk.ntaps * op.pk.A * op.pk.BP;
        int64_t blocks = (n + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
        repack_kernel<<<(unsigned)blocks, 256, 0, st>>>(t->params, t->packed, op.pk);
        KFN_LAUNCH_OK();
      }
  // ---- forward: SCoordNet (cnn_wrapper/SCoordNet.py:18-44) and the feature tower (KFNet.py:315-341) on all B frames
  {
    const int64_t n = (int64_t)B * H * W * 3;
    preprocess_train_kernel<<<148 * 16, 256, 0, st>>>(images, b.x0, n);
    KFN_LAUNCH_OK();
  }
  auto sview = [&](int i, float* p) { return view(p, B, H / kSDiv[i], W / kSDiv[i], kLayers[i].cout, 0, kLayers[i].cout); };
  auto tview = [&](int i, float* p) { return view(p, B, H / kTDiv[i], W / kTDiv[i], kLayers[12 + i].cout, 0, kLayers[12 + i].cout); };
  const View vx0 = view(b.x0, B, H, W, 3, 0, 3);
  for (int i = 0; i < 12; ++i)
    RUN(layer_forward(c, i, i ? sview(i - 1, b.s[i - 1]) : vx0, sview(i, b.s[i]), i == 11 ? (int)EPI_EXP_CH3 : (int)EPI_NONE));
  for (int i = 0; i < 7; ++i) RUN(layer_forward(c, 12 + i, i ? tview(i - 1, b.tw[i - 1]) : vx0, tview(i, b.tw[i])));
  KFN_CUDA_OK(cudaMemcpyAsync(b.feat, b.tw[6], (size_t)B * P * 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RUN(launch_l2norm32(b.feat, (int64_t)B * P, st));
  // ---- forward: OFlowNet on the B-1 pairs (OFlowNet.py:17-56)
  RUN(launch_diff_volume(b.feat, b.feat + (int64_t)P * 32, B - 1, GH, GW, b.diff, 0, st));
  const int l_c0 = L_("Temporal", "conv0"), l_c1a = L_("Temporal", "conv1a"), l_c1b = L_("Temporal", "conv1b"), l_c2a = L_("Temporal", "conv2a"),
            l_c2b = L_("Temporal", "conv2b"), l_c3a = L_("Temporal", "conv3a"), l_c3b = L_("Temporal", "conv3b"), l_u2 = L_("Temporal", "upconv2"),
            l_c4 = L_("Temporal", "conv4"), l_u1 = L_("Temporal", "upconv1"), l_c5 = L_("Temporal", "conv5"), l_u0 = L_("Temporal", "upconv0"),
            l_c6 = L_("Temporal", "conv6"), l_pr = L_("Temporal", "prediction"), l_f1 = L_("Temporal", "fc1"), l_f2 = L_("Temporal", "fc2"),
            l_un = L_("Temporal", "uncertainty");
  struct UV { View a, d; };   // activation view and its gradient twin
  auto uv = [&](float* a, float* d, int hw, int cs, int coff, int C) { return UV{view(a, R, hw, hw, cs, coff, C), view(d, R, hw, hw, cs, coff, C)}; };
  const UV v_diff = uv(b.diff, b.ddiff, 8, 32, 0, 32), v_conv0 = uv(b.cat0, b.dcat0, 8, 48, 16, 32), v_up0 = uv(b.cat0, b.dcat0, 8, 48, 0, 16),
           v_cat0 = uv(b.cat0, b.dcat0, 8, 48, 0, 48), v_c1a = uv(b.c1a, b.dc1a, 4, 32, 0, 32), v_conv1b = uv(b.cat1, b.dcat1, 4, 64, 32, 32),
           v_up1 = uv(b.cat1, b.dcat1, 4, 64, 0, 32), v_cat1 = uv(b.cat1, b.dcat1, 4, 64, 0, 64), v_c2a = uv(b.c2a, b.dc2a, 2, 64, 0, 64),
           v_conv2b = uv(b.cat2, b.dcat2, 2, 128, 64, 64), v_up2 = uv(b.cat2, b.dcat2, 2, 128, 0, 64), v_cat2 = uv(b.cat2, b.dcat2, 2, 128, 0, 128),
           v_c3a = uv(b.c3a, b.dc3a, 1, 128, 0, 128), v_c3b = uv(b.c3b, b.dc3b, 1, 128, 0, 128), v_c4 = uv(b.c4, b.dc4, 2, 64, 0, 64),
           v_c5 = uv(b.c5, b.dc5, 4, 32, 0, 32), v_c6 = uv(b.c6, b.dc6, 8, 16, 0, 16), v_log = uv(b.logits, b.dlogits, 8, 1, 0, 1),
           v_fc1 = uv(b.fc1, b.dfc1, 1, 64, 0, 64), v_fc2 = uv(b.fc2, b.dfc2, 1, 32, 0, 32), v_unc = uv(b.unc, b.dunc, 1, 1, 0, 1);
  RUN(layer_forward(c, l_c0, v_diff.a, v_conv0.a));
  RUN(layer_forward(c, l_c1a, v_conv0.a, v_c1a.a));
  RUN(layer_forward(c, l_c1b, v_c1a.a, v_conv1b.a));
  RUN(layer_forward(c, l_c2a, v_conv1b.a, v_c2a.a));
  RUN(layer_forward(c, l_c2b, v_c2a.a, v_conv2b.a));
  RUN(layer_forward(c, l_c3a, v_conv2b.a, v_c3a.a));
  RUN(layer_forward(c, l_c3b, v_c3a.a, v_c3b.a));
  RUN(layer_forward(c, l_u2, v_c3b.a, v_up2.a));
  RUN(layer_forward(c, l_c4, v_cat2.a, v_c4.a));
  RUN(layer_forward(c, l_u1, v_c4.a, v_up1.a));
  RUN(layer_forward(c, l_c5, v_cat1.a, v_c5.a));
  RUN(layer_forward(c, l_u0, v_c5.a, v_up0.a));
  RUN(layer_forward(c, l_c6, v_cat0.a, v_c6.a));
  RUN(layer_forward(c, l_pr, v_c6.a, v_log.a));
  RUN(layer_forward(c, l_f1, v_c3b.a, v_fc1.a));
  RUN(layer_forward(c, l_f2, v_fc1.a, v_fc2.a));
  RUN(layer_forward(c, l_un, v_fc2.a, v_unc.a));
  // flow / sigma_w of pair (t-1, t) are entry t (entry 0 unused: frame 0 is the measurement, KFNet.py:112-118)
  RUN(zero(b.flow, (int64_t)P * 2, st));
  RUN(zero(b.sigw, P, st));
  RUN(launch_softmax_flow(b.logits, b.unc, 1, R, b.flow + (int64_t)P * 2, b.sigw + P, nullptr, st));
  // ---- forward: the unrolled filter = the scan with a re-init on frame 0 only (GetKFCoordBatch, KFNet.py:102-146)
  const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  RUN(launch_kalman_scan(b.s[11], b.flow, b.sigw, nullptr, nullptr, I16, 1, B, GH, GW, 0, 1 << 30, 0, b.outmap, b.temp, b.kf, nullptr, nullptr, st));
  // ---- losses and their gradients w.r.t. the three map sets (train.py:287-296)
  {
    LossParams lp;
    lp.est[0] = b.s[11]; lp.est[1] = b.temp; lp.est[2] = b.kf;
    for (int k = 0; k < 3; ++k) lp.G[k] = b.G[k];
    lp.gt = gt8; lp.mask = mask8; lp.images = images; lp.B = B; lp.GH = GH; lp.GW = GW; lp.H = H; lp.W = W;
    lp.mix[0] = 0.2f; lp.mix[1] = 0.2f; lp.mix[2] = 0.6f; lp.scal = t->scal;
    train_loss_kernel<<<3, 1024, 0, st>>>(lp);
    KFN_LAUNCH_OK();
  }
  // ---- backward through the filter, newest step first: G[2][t-1] receives the warp's scatter of step t
  RUN(zero(b.dflow, (int64_t)P * 2, st));
  RUN(zero(b.dsigw, P, st));
  for (int tt = B - 1; tt >= 1; --tt) {
    const int64_t o = (int64_t)tt * P;
    filter_backward_kernel<<<(P + 127) / 128, 128, 0, st>>>(
        reinterpret_cast<const float4*>(b.kf) + o - P, reinterpret_cast<const float2*>(b.flow) + o, b.sigw + o,
        reinterpret_cast<const float4*>(b.s[11]) + o, reinterpret_cast<const float4*>(b.G[1]) + o, reinterpret_cast<const float4*>(b.G[2]) + o,
        b.G[2] + (o - P) * 4, reinterpret_cast<float2*>(b.dflow) + o, b.dsigw + o, reinterpret_cast<float4*>(b.G[0]) + o, GH, GW);
    KFN_LAUNCH_OK();
  }
  measure_grad_finish_kernel<<<(B * P + 255) / 256, 256, 0, st>>>(reinterpret_cast<float4*>(b.G[0]), reinterpret_cast<const float4*>(b.G[1]),
                                                                  reinterpret_cast<const float4*>(b.G[2]), reinterpret_cast<const float4*>(b.s[11]), P, B * P);
  KFN_LAUNCH_OK();
  // ---- backward: SCoordNet
  KFN_CUDA_OK(cudaMemcpyAsync(b.ds[11], b.G[0], (size_t)B * P * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (t->train_scoordnet) {
    for (int i = 11; i >= 0; --i) {
      const View in = i ? sview(i - 1, b.s[i - 1]) : vx0;
      const View din = i ? sview(i - 1, b.ds[i - 1]) : View{nullptr, 0, 0, 0, 0, 0, 0};
      RUN(layer_backward(c, i, in, sview(i, b.s[i]), sview(i, b.ds[i]), i ? &din : nullptr, false));
    }
  }
  // ---- backward: soft-argmax flow / sigma_w, U-Net, difference volume, tower
  if (t->train_oflownet) {
    softmax_flow_backward_kernel<<<(unsigned)(((int64_t)R * 32 + 255) / 256), 256, 0, st>>>(b.logits, b.dflow + (int64_t)P * 2, b.dsigw + P, b.sigw + P, R,
                                                                                            b.dlogits, b.dunc);
    KFN_LAUNCH_OK();
    RUN(layer_backward(c, l_pr, v_c6.a, v_log.a, v_log.d, &v_c6.d, false));
    RUN(layer_backward(c, l_c6, v_cat0.a, v_c6.a, v_c6.d, &v_cat0.d, false));          // writes d(upconv0 out) and a first part of d(conv0 out)
    RUN(layer_backward(c, l_u0, v_c5.a, v_up0.a, v_up0.d, &v_c5.d, false));
    RUN(layer_backward(c, l_c5, v_cat1.a, v_c5.a, v_c5.d, &v_cat1.d, false));
    RUN(layer_backward(c, l_u1, v_c4.a, v_up1.a, v_up1.d, &v_c4.d, false));
    RUN(layer_backward(c, l_c4, v_cat2.a, v_c4.a, v_c4.d, &v_cat2.d, false));
    RUN(layer_backward(c, l_u2, v_c3b.a, v_up2.a, v_up2.d, &v_c3b.d, false));
    RUN(layer_backward(c, l_un, v_fc2.a, v_unc.a, v_unc.d, &v_fc2.d, false));
    RUN(layer_backward(c, l_f2, v_fc1.a, v_fc2.a, v_fc2.d, &v_fc1.d, false));
    RUN(layer_backward(c, l_f1, v_c3b.a, v_fc1.a, v_fc1.d, &v_c3b.d, true));            // conv3b feeds upconv2 AND fc1
    RUN(layer_backward(c, l_c3b, v_c3a.a, v_c3b.a, v_c3b.d, &v_c3a.d, false));
    RUN(layer_backward(c, l_c3a, v_conv2b.a, v_c3a.a, v_c3a.d, &v_conv2b.d, true));     // skip connection: conv2b also feeds conv4
    RUN(layer_backward(c, l_c2b, v_c2a.a, v_conv2b.a, v_conv2b.d, &v_c2a.d, false));
    RUN(layer_backward(c, l_c2a, v_conv1b.a, v_c2a.a, v_c2a.d, &v_conv1b.d, true));     // conv1b also feeds conv5
    RUN(layer_backward(c, l_c1b, v_c1a.a, v_conv1b.a, v_conv1b.d, &v_c1a.d, false));
    RUN(layer_backward(c, l_c1a, v_conv0.a, v_c1a.a, v_c1a.d, &v_conv0.d, true));       // conv0 also feeds conv6
    RUN(layer_backward(c, l_c0, v_diff.a, v_conv0.a, v_conv0.d, &v_diff.d, false));
    {
      const int64_t n = (int64_t)B * P * 32;
      diff_volume_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.ddiff, b.dfeat, B, GH, GW);
      KFN_LAUNCH_OK();
      l2norm_backward_kernel<<<(unsigned)(((int64_t)B * P + 127) / 128), 128, 0, st>>>(b.tw[6], b.dfeat, b.dtw[6], (int64_t)B * P);
      KFN_LAUNCH_OK();
    }
    for (int i = 6; i >= 0; --i) {
      const View in = i ? tview(i - 1, b.tw[i - 1]) : vx0;
      const View din = i ? tview(i - 1, b.dtw[i - 1]) : View{nullptr, 0, 0, 0, 0, 0, 0};
      RUN(layer_backward(c, 12 + i, in, tview(i, b.tw[i]), tview(i, b.dtw[i]), i ? &din : nullptr, false));
    }
  }
#undef RUN
  // ---- regulariser value (scalars only) and the optimiser step
  struct Range { int64_t off, n; bool on; };
  const Range ranges[2] = {{0, t->n_scorenet, t->train_scoordnet}, {t->n_scorenet, t->n_params - t->n_scorenet, t->train_oflownet}};
  for (int r = 0; r < 2; ++r) {
    KFN_CUDA_OK(cudaMemsetAsync(t->sumsq_partial + r * 1024, 0, 1024 * sizeof(float), st));
    if (!ranges[r].on) continue;
    sumsq_kernel<<<592, 256, 0, st>>>(t->params + ranges[r].off, ranges[r].n, t->sumsq_partial + r * 1024);
    KFN_LAUNCH_OK();
  }
  float scal[8] = {0};
  std::vector<float> ssq(2048);
  KFN_CUDA_OK(cudaMemcpyAsync(scal, t->scal, 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
  KFN_CUDA_OK(cudaMemcpyAsync(ssq.data(), t->sumsq_partial, 2048 * sizeof(float), cudaMemcpyDeviceToHost, st));
  const double lr = (double)base_lr * std::pow((double)gamma, (double)t->step / (double)(decay_steps > 0 ? decay_steps : 1));   // exponential_decay
  if (apply_update) {
    const int tstep = t->adam_t + 1;
    const double b1 = 0.9, b2 = 0.999;
    const float lr_t = (float)(lr * std::sqrt(1.0 - std::pow(b2, tstep)) / (1.0 - std::pow(b1, tstep)));
    for (int r = 0; r < 2; ++r) {
      if (!ranges[r].on) continue;
      adam_kernel<<<592, 256, 0, st>>>(t->params + ranges[r].off, t->grads + ranges[r].off, t->adam_m + ranges[r].off, t->adam_v + ranges[r].off,
                                       ranges[r].n, weight_decay, lr_t, (float)b1, (float)b2, 1e-8f);
      KFN_LAUNCH_OK();
    }
    t->adam_t = tstep;
    t->step += 1;
  }
  KFN_CUDA_OK(cudaStreamSynchronize(st));
  if (scalars_host) {
    double reg = 0.0;
    for (float v : ssq) reg += v;
    reg *= 0.5 * (double)weight_decay;
    // [0] loss (0.2 / 0.2 / 0.6 mix, train.py:294-296), [1..6] measure / temporal / KF (loss, accuracy), [7] regulariser, [8] learning rate
    scalars_host[0] = 0.2f * scal[0] + 0.2f * scal[2] + 0.6f * scal[4];
    for (int i = 0; i < 6; ++i) scalars_host[1 + i] = scal[i];
    scalars_host[7] = (float)reg;
    scalars_host[8] = (float)lr;
  }
  return KFN_OK;
}

// what: 0 parameter, 1 gradient of the last step (without the regulariser term), 2 Adam m, 3 Adam v -- by TF variable name
int kfn_train_get(kfn_handle* h, const char* name, int what, float* host_out, int64_t n) {
  KFN_REQUIRE(h && h->train && name && host_out, "null argument / kfn_train_begin not called");
  KFN_CUDA_OK(cudaSetDevice(h->device));
  TrainState* t = h->train;
  const std::string nm(name);
  const size_t cut = nm.rfind('/');
  KFN_REQUIRE(cut != std::string::npos, "variable name must be <scope>/<layer>/{kernel,bias}");
  const int li = layer_index(nm.substr(0, cut));
  KFN_REQUIRE(li >= 0, "unknown layer");
  const bool bias = nm.substr(cut + 1) == "bias";
  const int64_t count = bias ? kLayers[li].cout : layer_kernel_elems(kLayers[li]);
  KFN_REQUIRE(n == count, "element count mismatch");
  const float* src = what == 0 ? t->params : what == 1 ? t->grads : what == 2 ? t->adam_m : what == 3 ? t->adam_v : nullptr;
  KFN_REQUIRE(src, "what must be 0..3");
  KFN_CUDA_OK(cudaMemcpy(host_out, src + (bias ? t->b_off[li] : t->w_off[li]), (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  return KFN_OK;
}

// global_step of the optimiser (train.py:361-362,375: `init_step = global_step.assign(step)`; --reset_step): drives the learning-rate decay
// only -- Adam's bias correction counts the updates applied since kfn_train_begin, like a freshly built AdamOptimizer
int kfn_train_set_step(kfn_handle* h, int step) {
  KFN_REQUIRE(h && h->train && step >= 0, "kfn_train_begin not called / negative step");
  h->train->step = step;
  return KFN_OK;
}

// copies the trained parameters back into the handle's weight set and re-packs the inference operands
int kfn_train_commit(kfn_handle* h) {
  KFN_REQUIRE(h && h->train, "kfn_train_begin not called");
  KFN_CUDA_OK(cudaSetDevice(h->device));
  TrainState* t = h->train;
  for (int i = 0; i < kNumLayers; ++i) {
    const std::string base = std::string(kLayers[i].scope) + "/" + kLayers[i].name;
    std::vector<float>& w = h->host_w[base + "/kernel"];
    std::vector<float>& b = h->host_w[base + "/bias"];
    KFN_CUDA_OK(cudaMemcpy(w.data(), t->params + t->w_off[i], w.size() * sizeof(float), cudaMemcpyDeviceToHost));
    KFN_CUDA_OK(cudaMemcpy(b.data(), t->params + t->b_off[i], b.size() * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return kfn_finalize_weights(h);
}

}  // extern "C"
