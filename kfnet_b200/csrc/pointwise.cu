// Memory-bound kernels of the KFNet path: input preprocessing, feature-difference volume,
// softmax -> soft-argmax flow, and the serial warp + Kalman scan.  Compiled with --fmad=false so the
// few flops here round exactly like the reference's separate fp32 TF ops (mul, add, div, sqrt).
#include <cooperative_groups.h>
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"

namespace kfn {

// ---------------------------------------------------------------------------------------------
// (img - 128) * 0.00625  (cnn_wrapper/SCoordNet.py:34-37, KFNet/KFNet.py:317)
__global__ void preprocess_f32_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = ((float)img[i] - 128.0f) * 0.00625f;
}

int launch_preprocess_f32(const uint8_t* img, float* out, int64_t n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  preprocess_f32_kernel<<<blocks, 256, 0, st>>>(img, out, n);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// 3x3 SAME im2col of the raw image for the two Cin=3 first layers (SCoordNet conv1a, tower feat1):
// out[b,y,x, (ky*3+kx)*3 + c] = img[b, y+ky-1, x+kx-1, c] - 128  (0 outside; k = 27..31 zero).
// u8 - 128 is exact in bf16; the 0.00625 preprocess scale is folded into the packed weights.
// bf16 or (KFN_MODE_FP16X2) fp16 bit pattern of one stored activation value; -128..127 is exact in both
template <bool kHalf>
__device__ __forceinline__ __nv_bfloat16 to_limb(float v) {
  if constexpr (kHalf) { const __half h = __float2half_rn(v); return *reinterpret_cast<const __nv_bfloat16*>(&h); }
  else return __float2bfloat16(v);
}
template <bool kHalf>
__global__ void im2col27_bf16_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H,
                                     int W) {
  int64_t total = (int64_t)B * H * W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int x = (int)(i % W);
  int y = (int)((i / W) % H);
  const uint8_t* base = img + (i - (int64_t)y * W - x) * 3;   // start of image b
  __align__(16) __nv_bfloat16 v[32];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int yy = y + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int xx = x + kx - 1;
      bool inb = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
      const uint8_t* p = base + ((int64_t)yy * W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[(ky * 3 + kx) * 3 + c] = to_limb<kHalf>(inb ? (float)p[c] - 128.0f : 0.0f);
    }
  }
#pragma unroll
  for (int k = 27; k < 32; ++k) v[k] = to_limb<kHalf>(0.0f);
  uint4* dst = reinterpret_cast<uint4*>(out + i * 32);
  const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = src[q];
}

int launch_im2col27_bf16(const uint8_t* img, __nv_bfloat16* out, int B, int H, int W, cudaStream_t st, int f16) {
  int64_t total = (int64_t)B * H * W;
  if (f16) im2col27_bf16_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, out, B, H, W);
  else im2col27_bf16_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, out, B, H, W);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// tf.nn.l2_normalize over 32 channels: x * rsqrt(max(sum x^2, 1e-12))  (KFNet/KFNet.py:340)
__global__ void l2norm32_kernel(float* __restrict__ x, int64_t rows) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float4* p = reinterpret_cast<float4*>(x + r * 32);
  float4 v[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = p[i];
    ss += v[i].x * v[i].x;
    ss += v[i].y * v[i].y;
    ss += v[i].z * v[i].z;
    ss += v[i].w * v[i].w;
  }
  float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i].x *= inv; v[i].y *= inv; v[i].z *= inv; v[i].w *= inv;
    p[i] = v[i];
  }
}

int launch_l2norm32(float* x, int64_t rows, cudaStream_t st) {
  l2norm32_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(x, rows);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// Feature-difference volume (KFNet.BuildCoordVolume, KFNet/KFNet.py:343-359, reshape :372):
//   diff[n, p, i, j, c] = f_cur[n,p,c] - f_prev[n, p + (ox,oy), c],  (ox,oy) = (j-4, i-4), 0 outside.
// One warp per label pixel.  A lane owns a channel group (CH channels = 16 output bytes) and a sub-offset; f_cur stays
// in registers, the loop over the remaining offsets is fully unrolled (all neighbour loads in flight at once), and
// every store instruction of the warp writes one contiguous 512-byte run.  ~15 instructions per 16 bytes stored
// (the first version decoded (n,p,k,g) per thread with div/mod: ~100 instructions, instruction bound at 45 % of HBM).
template <bool kBf16, int LIMBS>
__global__ void __launch_bounds__(256) diff_volume_kernel(const float* __restrict__ f_prev, const float* __restrict__ f_cur,
                                                          uint4* __restrict__ out, int64_t npix, int GH, int GW) {
  constexpr int CH = kBf16 ? 8 : 4;            // channels per lane
  constexpr int GROUPS = 32 / CH;              // lanes per (pixel, offset)
  constexpr int KSUB = 32 / GROUPS;            // offsets covered by one warp store
  constexpr int ITERS = 64 / KSUB;
  const int lane = threadIdx.x & 31;
  const int g = lane % GROUPS, ksub = lane / GROUPS;
  const int P = GH * GW;
  int64_t np = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; np < npix; np += wstride) {
    const int p = (int)(np % P);
    const int64_t nbase = np - p;                // n * P
    const int y = p / GW, x = p - y * GW;
    const float4* cur = reinterpret_cast<const float4*>(f_cur + np * 32 + g * CH);
    const float4 a0 = __ldg(cur);
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kBf16) a1 = __ldg(cur + 1);
    float4 b0[ITERS], b1[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int k = it * KSUB + ksub;
      const int yy = y + (k >> 3) - 4, xx = x + (k & 7) - 4;
      const bool inb = (yy >= 0) && (yy < GH) && (xx >= 0) && (xx < GW);
      const float4* prv = reinterpret_cast<const float4*>(f_prev + (nbase + (inb ? yy * GW + xx : 0)) * 32 + g * CH);
      b0[it] = inb ? __ldg(prv) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (kBf16) b1[it] = inb ? __ldg(prv + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint4* dst = out + np * 64 * GROUPS + lane;  // + it * 32 : consecutive lanes -> consecutive 16-byte words
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      uint4 o;
      if constexpr (kBf16 && LIMBS == 2) {
        // fp16x2: two fp16 limbs per (pixel, offset): [limb][32 channels], the second = residual * 2^11
        float d[8] = {a0.x - b0[it].x, a0.y - b0[it].y, a0.z - b0[it].z, a0.w - b0[it].w,
                      a1.x - b1[it].x, a1.y - b1[it].y, a1.z - b1[it].z, a1.w - b1[it].w};
        uint4* base = out + ((np * 64 + it * KSUB + ksub) * 2) * GROUPS + g;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          __half2 h0 = __floats2half2_rn(d[0], d[1]), h1 = __floats2half2_rn(d[2], d[3]);
          __half2 h2 = __floats2half2_rn(d[4], d[5]), h3 = __floats2half2_rn(d[6], d[7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
          base[l * GROUPS] = o;
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = (d[i] - __half2float(__float2half_rn(d[i]))) * 2048.0f;
        }
        continue;
      } else if constexpr (kBf16 && LIMBS == 3) {
        // bf16x3: three limbs per (pixel, offset): [limb][32 channels]
        float d[8] = {a0.x - b0[it].x, a0.y - b0[it].y, a0.z - b0[it].z, a0.w - b0[it].w,
                      a1.x - b1[it].x, a1.y - b1[it].y, a1.z - b1[it].z, a1.w - b1[it].w};
        uint4* base = out + ((np * 64 + it * KSUB + ksub) * 3) * GROUPS + g;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(d[0], d[1]), h1 = __floats2bfloat162_rn(d[2], d[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(d[4], d[5]), h3 = __floats2bfloat162_rn(d[6], d[7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
          base[l * GROUPS] = o;
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] -= __bfloat162float(__float2bfloat16(d[i]));
        }
        continue;
      } else if constexpr (kBf16) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a0.x - b0[it].x, a0.y - b0[it].y);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(a0.z - b0[it].z, a0.w - b0[it].w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(a1.x - b1[it].x, a1.y - b1[it].y);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(a1.z - b1[it].z, a1.w - b1[it].w);
        o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
      } else {
        o.x = __float_as_uint(a0.x - b0[it].x); o.y = __float_as_uint(a0.y - b0[it].y);
        o.z = __float_as_uint(a0.z - b0[it].z); o.w = __float_as_uint(a0.w - b0[it].w);
      }
      __stcs(dst + it * 32, o);      // streaming store: the volume is written once and read once by conv0
    }
  }
}

int launch_diff_volume(const float* f_prev, const float* f_cur, int NP, int GH, int GW, void* diff, int out_bf16,
                       cudaStream_t st) {
  const int64_t npix = (int64_t)NP * GH * GW;     // one warp each
  int64_t blocks = (npix + 7) / 8;
  if (blocks > 148 * 64) blocks = 148 * 64;       // grid-stride beyond that; multiple of the SM count
  if (out_bf16 == 2)
    diff_volume_kernel<true, 2><<<(unsigned)blocks, 256, 0, st>>>(f_prev, f_cur, reinterpret_cast<uint4*>(diff), npix, GH, GW);
  else if (out_bf16 == 3)
    diff_volume_kernel<true, 3><<<(unsigned)blocks, 256, 0, st>>>(f_prev, f_cur, reinterpret_cast<uint4*>(diff), npix, GH, GW);
  else if (out_bf16)
    diff_volume_kernel<true, 1><<<(unsigned)blocks, 256, 0, st>>>(f_prev, f_cur, reinterpret_cast<uint4*>(diff), npix, GH, GW);
  else
    diff_volume_kernel<false, 1><<<(unsigned)blocks, 256, 0, st>>>(f_prev, f_cur, reinterpret_cast<uint4*>(diff), npix, GH, GW);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// softmax over the 64 window logits -> soft-argmax flow (cnn_wrapper/OFlowNet.py:45-47,
// KFNet/KFNet.py:381-385), argmax_k (diagnostic) and sigma_w = exp(u) * 1e-2 (OFlowNet.py:56).
// One warp per label pixel: lane l owns logits k = 2l, 2l+1 (one coalesced 256-byte row per warp).
__global__ void softmax_flow_kernel(const float* __restrict__ logits, const float* __restrict__ logsig,
                                    int logsig_stride, int64_t rows, float* __restrict__ flow,
                                    float* __restrict__ sigma_w, int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < rows; row += wstride) {
    float2 l = reinterpret_cast<const float2*>(logits + row * 64)[lane];
    int k0 = 2 * lane;
    float m = fmaxf(l.x, l.y);
    int mi = (l.y > l.x) ? k0 + 1 : k0;          // lowest index wins ties (np.argmax)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, m, o);
      int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    float e0 = expf(l.x - m), e1 = expf(l.y - m);
    float s = e0 + e1;
    // offsets[k] = (ox, oy) = (k%8 - 4, k/8 - 4)   (KFNet.py:348-350)
    float ox0 = (float)((k0 & 7) - 4), oy0 = (float)((k0 >> 3) - 4);
    float fx = e0 * ox0 + e1 * (ox0 + 1.0f);
    float fy = (e0 + e1) * oy0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      fx += __shfl_xor_sync(0xffffffffu, fx, o);
      fy += __shfl_xor_sync(0xffffffffu, fy, o);
    }
    if (lane == 0) {
      flow[row * 2 + 0] = fx / s;
      flow[row * 2 + 1] = fy / s;
      sigma_w[row] = expf(logsig[row * logsig_stride]) * 1e-2f;
      if (argmax) argmax[row] = mi;
    }
  }
}

int launch_softmax_flow(const float* logits, const float* logsig, int logsig_stride, int64_t rows, float* flow,
                        float* sigma_w, int32_t* argmax, cudaStream_t st) {
  int64_t blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  softmax_flow_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, logsig, logsig_stride, rows, flow, sigma_w, argmax);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// Serial scan over frames: flow-guided bilinear warp of the filter state (tools/util.py:3-94 with
// the clipped-neighbour weights of :56-59 and the add_n order of :90-93), process noise
// (KFNet/KFNet.py:393-401), Kalman gain / posterior (KFNet.py:154-160), NIS (:164-184), rigid
// transform (KFNet/util.py:31-37) and the state feedback / re-init of eval.py:87-104.
// One CTA per sequence; the state (xyz, sigma) of all GH*GW pixels lives in shared memory
// (double-buffered: the gather of frame t reads buffer t&1, the posterior goes to the other one).
struct Xform { float m[16]; };

__global__ void __launch_bounds__(1024, 1)
kalman_scan_kernel(const float4* __restrict__ measure, const float2* __restrict__ flow, const float* __restrict__ sigma_w,
                   const float4* __restrict__ state_in, float4* __restrict__ state_out, Xform X, int T, int GH, int GW,
                   int first_index, int reinit_period, int nis_gate, float4* __restrict__ out_map,
                   float4* __restrict__ temp, float4* __restrict__ kf, float* __restrict__ nis, int2* __restrict__ taps) {
  extern __shared__ float4 s_state[];          // [2][P]
  const int P = GH * GW;
  const int s = blockIdx.x;
  const int64_t seq = (int64_t)s * T * P;
  float4* cur = s_state;
  float4* nxt = s_state + P;
  if (state_in) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) cur[p] = state_in[(int64_t)s * P + p];
  }
  __syncthreads();
  const float xmax = (float)(GW - 1), ymax = (float)(GH - 1);
  const float minvar = 1e-5f * 1e-5f;          // KFNet.py:71,396,398 (min_uncertainty^2)
  const bool need_nis = (nis != nullptr) || (nis_gate != 0);   // the NIS divisions are skipped when nobody reads them
  for (int t = 0; t < T; ++t) {
    const int64_t fr = seq + (int64_t)t * P;
    const int idx = first_index + t;
    const bool reinit = (reinit_period > 0) && (idx % reinit_period == 0);
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const float4 z = measure[fr + p];        // (xyz, sigma_z)
      float4 xm, xp;                           // prior (temp) and posterior (KF), w = sigma
      float n0 = 0.f, n1 = 0.f, n2 = 0.f;
      int2 tp = make_int2(0, 0);
      if (reinit) {
        xm = z; xp = z;                        // eval.py:94-101
      } else {
        const float2 fl = flow[fr + p];
        const int y = p / GW, x = p - y * GW;
        const float cx = (float)x + fl.x, cy = (float)y + fl.y;          // KFNet.py:386
        const float x0 = floorf(cx), y0 = floorf(cy);
        const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
        const float x0s = fminf(fmaxf(x0, 0.f), xmax), x1s = fminf(fmaxf(x1, 0.f), xmax);
        const float y0s = fminf(fmaxf(y0, 0.f), ymax), y1s = fminf(fmaxf(y1, 0.f), ymax);
        const float wx0 = x1s - cx, wx1 = cx - x0s, wy0 = y1s - cy, wy1 = cy - y0s;
        const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
        const int ix0 = (int)x0s, ix1 = (int)x1s, iy0 = (int)y0s, iy1 = (int)y1s;
        const float4 i00 = cur[iy0 * GW + ix0], i01 = cur[iy1 * GW + ix0];
        const float4 i10 = cur[iy0 * GW + ix1], i11 = cur[iy1 * GW + ix1];
        xm.x = ((w00 * i00.x + w01 * i01.x) + w10 * i10.x) + w11 * i11.x;
        xm.y = ((w00 * i00.y + w01 * i01.y) + w10 * i10.y) + w11 * i11.y;
        xm.z = ((w00 * i00.z + w01 * i01.z) + w10 * i10.z) + w11 * i11.z;
        const float sl = ((w00 * i00.w + w01 * i01.w) + w10 * i10.w) + w11 * i11.w;
        const float sw = sigma_w[fr + p];
        const float last_var = fmaxf(sl * sl, minvar);
        const float trans_var = fmaxf(sw * sw, minvar);
        xm.w = sqrtf(trans_var + last_var);                               // temp_uncertainty
        // BuildKFCoord (KFNet.py:154-160) -- squares temp_uncertainty again, like the reference
        const float lv = xm.w * xm.w, mv = z.w * z.w;
        const float K = lv / (lv + mv);
        const float omk = fmaxf(1.0f - K, 0.0f);
        xp.x = omk * xm.x + K * z.x;
        xp.y = omk * xm.y + K * z.y;
        xp.z = omk * xm.z + K * z.z;
        xp.w = sqrtf(omk * lv);
        if (need_nis) {
          // GetNIS (KFNet.py:164-184): inno_variance = square(sqrt(temp_var + measure_var))
          const float iu = sqrtf(lv + mv);
          const float iv = iu * iu;
          const float d0 = z.x - xm.x, d1 = z.y - xm.y, d2 = z.z - xm.z;
          n0 = (d0 * d0) / iv; n1 = (d1 * d1) / iv; n2 = (d2 * d2) / iv;
        }
        tp = make_int2((int)x0, (int)y0);
      }
      nxt[p] = xp;
      // output map: (T * x+, 1/sigma+), NIS-gated to the transformed measurement if asked (eval.py:87-92,123)
      float4 src = xp;
      if (nis_gate && !reinit && ((n0 + n1) + n2 > 7.815f)) { src.x = z.x; src.y = z.y; src.z = z.z; }
      float4 o;
      o.x = ((X.m[0] * src.x + X.m[1] * src.y) + X.m[2] * src.z) + X.m[3];
      o.y = ((X.m[4] * src.x + X.m[5] * src.y) + X.m[6] * src.z) + X.m[7];
      o.z = ((X.m[8] * src.x + X.m[9] * src.y) + X.m[10] * src.z) + X.m[11];
      o.w = 1.0f / xp.w;
      out_map[fr + p] = o;
      if (temp) temp[fr + p] = xm;
      if (kf) kf[fr + p] = xp;
      if (nis) { nis[(fr + p) * 3 + 0] = n0; nis[(fr + p) * 3 + 1] = n1; nis[(fr + p) * 3 + 2] = n2; }
      if (taps) taps[fr + p] = tp;
    }
    __syncthreads();
    float4* sw_ = cur; cur = nxt; nxt = sw_;
  }
  if (state_out) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) state_out[(int64_t)s * P + p] = cur[p];
  }
}

// Cluster variant: the single-CTA scan above is instruction-issue bound on its one SM (~180 instructions per pixel),
// so a sequence is spread over a thread-block cluster of scan_cluster_size() CTAs.  Each CTA owns a band of label-grid rows
// and keeps that band's state (double-buffered) in its own shared memory; the bilinear taps of the warp read whichever
// CTA owns the sampled row through distributed shared memory (cluster.map_shared_rank), and one cluster.sync() per
// frame replaces the __syncthreads().  Same arithmetic, same order: results are bit-identical to the single-CTA kernel.
namespace cg = cooperative_groups;
constexpr int kScanClusterMax = 8;              // portable cluster limit; the size itself is a launch attribute (scan_cluster_size)

__global__ void __launch_bounds__(1024, 1)
kalman_scan_cluster_kernel(const float4* __restrict__ measure, const float2* __restrict__ flow, const float* __restrict__ sigma_w,
                           const float4* __restrict__ state_in, float4* __restrict__ state_out, Xform X, int T, int GH, int GW,
                           int first_index, int reinit_period, int nis_gate, float4* __restrict__ out_map,
                           float4* __restrict__ temp, float4* __restrict__ kf, float* __restrict__ nis, int2* __restrict__ taps) {
  extern __shared__ float4 s_band[];           // [2][cap]  this CTA's rows of the state
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  const int s = blockIdx.x / csize;
  const int P = GH * GW;
  const int rows_per = (GH + csize - 1) / csize;
  const int cap = rows_per * GW;
  const int y_lo = rank * rows_per;
  const int y_hi = min(GH, y_lo + rows_per);
  const int nloc = max(0, y_hi - y_lo) * GW;   // pixels owned by this CTA
  const int p_lo = y_lo * GW;
  const int64_t seq = (int64_t)s * T * P;
  int cur = 0;
  if (state_in) {
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) s_band[i] = state_in[(int64_t)s * P + p_lo + i];
  }
  cluster.sync();
  const float xmax = (float)(GW - 1), ymax = (float)(GH - 1);
  const float minvar = 1e-5f * 1e-5f;
  const bool need_nis = (nis != nullptr) || (nis_gate != 0);
  auto gather = [&](int iy, int ix) -> float4 {
    const int r = iy / rows_per;                // owner of the sampled row: its band through distributed shared memory (mapa)
    return cluster.map_shared_rank(s_band, r)[cur * cap + (iy - r * rows_per) * GW + ix];
  };
  // inputs of this thread's first pixel of the NEXT frame are loaded while the current frame is still being computed and the
  // cluster synchronises: the L2 round trip of the three loads leaves the per-frame critical path
  float4 z_n = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 fl_n = make_float2(0.f, 0.f);
  float sw_n = 0.f;
  const bool own_first = (int)threadIdx.x < nloc;
  if (own_first && T > 0) {
    const int64_t q = seq + p_lo + threadIdx.x;
    z_n = measure[q]; fl_n = flow[q]; sw_n = sigma_w[q];
  }
  for (int t = 0; t < T; ++t) {
    const int64_t fr = seq + (int64_t)t * P;
    const int idx = first_index + t;
    const bool reinit = (reinit_period > 0) && (idx % reinit_period == 0);
    float4* nxt = s_band + (cur ^ 1) * cap;
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
      const int p = p_lo + i;
      const bool pre = i == (int)threadIdx.x;   // the prefetched pixel
      const float4 z = pre ? z_n : measure[fr + p];
      const float2 fl_in = pre ? fl_n : flow[fr + p];
      const float sw_in = pre ? sw_n : sigma_w[fr + p];
      if (pre && t + 1 < T) {
        const int64_t q = fr + P + p;
        z_n = measure[q]; fl_n = flow[q]; sw_n = sigma_w[q];
      }
      float4 xm, xp;
      float n0 = 0.f, n1 = 0.f, n2 = 0.f;
      int2 tp = make_int2(0, 0);
      if (reinit) {
        xm = z; xp = z;
      } else {
        const float2 fl = fl_in;
        const int y = p / GW, x = p - y * GW;
        const float cx = (float)x + fl.x, cy = (float)y + fl.y;
        const float x0 = floorf(cx), y0 = floorf(cy);
        const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
        const float x0s = fminf(fmaxf(x0, 0.f), xmax), x1s = fminf(fmaxf(x1, 0.f), xmax);
        const float y0s = fminf(fmaxf(y0, 0.f), ymax), y1s = fminf(fmaxf(y1, 0.f), ymax);
        const float wx0 = x1s - cx, wx1 = cx - x0s, wy0 = y1s - cy, wy1 = cy - y0s;
        const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
        const int ix0 = (int)x0s, ix1 = (int)x1s, iy0 = (int)y0s, iy1 = (int)y1s;
        const float4 i00 = gather(iy0, ix0), i01 = gather(iy1, ix0);
        const float4 i10 = gather(iy0, ix1), i11 = gather(iy1, ix1);
        xm.x = ((w00 * i00.x + w01 * i01.x) + w10 * i10.x) + w11 * i11.x;
        xm.y = ((w00 * i00.y + w01 * i01.y) + w10 * i10.y) + w11 * i11.y;
        xm.z = ((w00 * i00.z + w01 * i01.z) + w10 * i10.z) + w11 * i11.z;
        const float sl = ((w00 * i00.w + w01 * i01.w) + w10 * i10.w) + w11 * i11.w;
        const float sw = sw_in;
        const float last_var = fmaxf(sl * sl, minvar);
        const float trans_var = fmaxf(sw * sw, minvar);
        xm.w = sqrtf(trans_var + last_var);
        const float lv = xm.w * xm.w, mv = z.w * z.w;
        const float K = lv / (lv + mv);
        const float omk = fmaxf(1.0f - K, 0.0f);
        xp.x = omk * xm.x + K * z.x;
        xp.y = omk * xm.y + K * z.y;
        xp.z = omk * xm.z + K * z.z;
        xp.w = sqrtf(omk * lv);
        if (need_nis) {
          const float iu = sqrtf(lv + mv);
          const float iv = iu * iu;
          const float d0 = z.x - xm.x, d1 = z.y - xm.y, d2 = z.z - xm.z;
          n0 = (d0 * d0) / iv; n1 = (d1 * d1) / iv; n2 = (d2 * d2) / iv;
        }
        tp = make_int2((int)x0, (int)y0);
      }
      nxt[i] = xp;
      float4 src = xp;
      if (nis_gate && !reinit && ((n0 + n1) + n2 > 7.815f)) { src.x = z.x; src.y = z.y; src.z = z.z; }
      float4 o;
      o.x = ((X.m[0] * src.x + X.m[1] * src.y) + X.m[2] * src.z) + X.m[3];
      o.y = ((X.m[4] * src.x + X.m[5] * src.y) + X.m[6] * src.z) + X.m[7];
      o.z = ((X.m[8] * src.x + X.m[9] * src.y) + X.m[10] * src.z) + X.m[11];
      o.w = 1.0f / xp.w;
      out_map[fr + p] = o;
      if (temp) temp[fr + p] = xm;
      if (kf) kf[fr + p] = xp;
      if (nis) { nis[(fr + p) * 3 + 0] = n0; nis[(fr + p) * 3 + 1] = n1; nis[(fr + p) * 3 + 2] = n2; }
      if (taps) taps[fr + p] = tp;
    }
    cluster.sync();                            // all bands of frame t written before anyone gathers for frame t+1
    cur ^= 1;
  }
  if (state_out) {
    for (int i = threadIdx.x; i < nloc; i += blockDim.x) state_out[(int64_t)s * P + p_lo + i] = s_band[cur * cap + i];
  }
  cluster.sync();                              // no CTA may exit while a peer can still read its shared memory
}

int launch_kalman_scan(const float* measure, const float* flow, const float* sigma_w, const float* state_in,
                       float* state_out, const float* T16_host, int S, int T, int GH, int GW, int first_index,
                       int reinit_period, int nis_gate, float* out_map, float* temp, float* kf, float* nis,
                       int32_t* taps, cudaStream_t st) {
  KFN_REQUIRE(S > 0 && T > 0 && GH > 0 && GW > 0, "empty scan");
  // without an incoming state the first frame must re-initialise the filter from the measurement (eval.py:94-101);
  // otherwise frame 0 would warp an uninitialised shared-memory state
  KFN_REQUIRE(state_in || (reinit_period > 0 && first_index % reinit_period == 0),
              "state_in is required unless the first frame is a re-initialisation frame (first_index % reinit_period == 0)");
  static const bool use_cluster = [] { const char* e = getenv("KFN_SCAN_CLUSTER"); return !(e && e[0] == '0'); }();
  int dev0 = 0;
  KFN_CUDA_OK(cudaGetDevice(&dev0));
  const int num_sms = sm_count(dev0);
  // latency regime only: with many sequences one CTA per sequence already fills the SMs (measured: 148 sequences run
  // 8.2 us/frame-step single-CTA vs 25 us as 592 cluster CTAs; 8 sequences run 7.2 vs 4.8 us/frame)
  // cluster size: the largest of 8..2 that leaves every CTA >= 2 grid rows and the clusters half of the SMs (measured on a B200, one
  // 256-frame sequence of 60x80: 8 CTAs 2.78 us per frame, 6: 2.97, 5: 3.16, 4: 4.11 -- 1200 pixels per CTA are two passes of its
  // 1024 threads --, 3: 4.44, single CTA 6.33; eight sequences: 3.11 / 3.33 / 3.49 / 4.97 / 5.16 / 7.35); KFN_SCAN_CLUSTER_SIZE pins it
  static const int forced = [] { const char* e = getenv("KFN_SCAN_CLUSTER_SIZE"); const int v = e ? atoi(e) : 0;
                                 return (v >= 2 && v <= kScanClusterMax) ? v : 0; }();
  int kScanCluster = forced ? forced : kScanClusterMax;
  while (!forced && kScanCluster > 2 && !(GH >= 2 * kScanCluster && 2 * S * kScanCluster <= num_sms)) --kScanCluster;
  if (use_cluster && GH >= 2 * kScanCluster && 2 * S * kScanCluster <= num_sms) {
    Xform Xc;
    for (int i = 0; i < 16; ++i) Xc.m[i] = T16_host[i];
    const int rows_per = (GH + kScanCluster - 1) / kScanCluster;
    const size_t smem_c = (size_t)2 * rows_per * GW * sizeof(float4);
    KFN_REQUIRE(smem_c <= 200 * 1024, "label grid too large for the cluster scan");
    static bool attr_dev[64] = {false};
    int dev_c = 0;
    KFN_CUDA_OK(cudaGetDevice(&dev_c));
    if (!attr_dev[dev_c & 63]) {
      KFN_CUDA_OK(cudaFuncSetAttribute(kalman_scan_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_dev[dev_c & 63] = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(S * kScanCluster));
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem_c;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)kScanCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, kalman_scan_cluster_kernel,
        reinterpret_cast<const float4*>(measure), reinterpret_cast<const float2*>(flow), sigma_w,
        reinterpret_cast<const float4*>(state_in), reinterpret_cast<float4*>(state_out), Xc, T, GH, GW, first_index,
        reinit_period, nis_gate, reinterpret_cast<float4*>(out_map), reinterpret_cast<float4*>(temp),
        reinterpret_cast<float4*>(kf), nis, reinterpret_cast<int2*>(taps)));
    KFN_LAUNCH_OK();
    return KFN_OK;
  }
  const size_t smem = (size_t)2 * GH * GW * sizeof(float4);
  KFN_REQUIRE(smem <= 200 * 1024, "label grid too large for the shared-memory resident filter state");
  Xform X;
  for (int i = 0; i < 16; ++i) X.m[i] = T16_host[i];
  static bool attr_set_dev[64] = {false};
  int dev_ = 0;
  KFN_CUDA_OK(cudaGetDevice(&dev_));
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    KFN_CUDA_OK(cudaFuncSetAttribute(kalman_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  kalman_scan_kernel<<<S, 1024, smem, st>>>(
      reinterpret_cast<const float4*>(measure), reinterpret_cast<const float2*>(flow), sigma_w,
      reinterpret_cast<const float4*>(state_in), reinterpret_cast<float4*>(state_out), X, T, GH, GW, first_index,
      reinit_period, nis_gate, reinterpret_cast<float4*>(out_map), reinterpret_cast<float4*>(temp),
      reinterpret_cast<float4*>(kf), nis, reinterpret_cast<int2*>(taps));
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// Pointwise Kalman update + NIS without the warp (KFNet.BuildKFCoord, KFNet/KFNet.py:148-162; KFNet.GetNIS, :164-184):
// prior (x-, sigma-) and measurement (z, sigma_z) as [npix,4] maps -> posterior [npix,4] and NIS [npix,3].
__global__ void kalman_update_kernel(const float4* __restrict__ prior, const float4* __restrict__ meas, int64_t npix,
                                     float4* __restrict__ post, float* __restrict__ nis) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float4 xm = prior[i], z = meas[i];
  const float lv = xm.w * xm.w, mv = z.w * z.w;
  if (post) {
    const float K = lv / (lv + mv);
    const float omk = fmaxf(1.0f - K, 0.0f);
    post[i] = make_float4(omk * xm.x + K * z.x, omk * xm.y + K * z.y, omk * xm.z + K * z.z, sqrtf(omk * lv));
  }
  if (nis) {
    const float iu = sqrtf(lv + mv);
    const float iv = iu * iu;
    const float d0 = z.x - xm.x, d1 = z.y - xm.y, d2 = z.z - xm.z;
    nis[i * 3 + 0] = (d0 * d0) / iv; nis[i * 3 + 1] = (d1 * d1) / iv; nis[i * 3 + 2] = (d2 * d2) / iv;
  }
}

int launch_kalman_update(const float* prior, const float* meas, int64_t npix, float* post, float* nis, cudaStream_t st) {
  kalman_update_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(prior), reinterpret_cast<const float4*>(meas), npix, reinterpret_cast<float4*>(post), nis);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone tools/util.py:bilinear_sampler (3-94): imgs [N,H,W,C] sampled at coords [N,Ht,Wt,2] (x,y) -> [N,Ht,Wt,C],
// with the reference's clipped-neighbour weights (:56-59) and add_n order (:90-93).  One thread per output element.
__global__ void bilinear_sampler_kernel(const float* __restrict__ imgs, const float* __restrict__ coords,
                                        float* __restrict__ out, int N, int H, int W, int C, int Ht, int Wt) {
  const int64_t total = (int64_t)N * Ht * Wt * C;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t pt = i / C;                       // (n, yt, xt)
  const int n = (int)(pt / ((int64_t)Ht * Wt));
  const float cx = coords[pt * 2 + 0], cy = coords[pt * 2 + 1];
  const float xmax = (float)(W - 1), ymax = (float)(H - 1);
  const float x0 = floorf(cx), y0 = floorf(cy);
  const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
  const float x0s = fminf(fmaxf(x0, 0.f), xmax), x1s = fminf(fmaxf(x1, 0.f), xmax);
  const float y0s = fminf(fmaxf(y0, 0.f), ymax), y1s = fminf(fmaxf(y1, 0.f), ymax);
  const float wx0 = x1s - cx, wx1 = cx - x0s, wy0 = y1s - cy, wy1 = cy - y0s;
  const float* base = imgs + (int64_t)n * H * W * C + c;
  const float i00 = base[((int64_t)y0s * W + (int64_t)x0s) * C], i01 = base[((int64_t)y1s * W + (int64_t)x0s) * C];
  const float i10 = base[((int64_t)y0s * W + (int64_t)x1s) * C], i11 = base[((int64_t)y1s * W + (int64_t)x1s) * C];
  out[i] = (((wx0 * wy0) * i00 + (wx0 * wy1) * i01) + (wx1 * wy0) * i10) + (wx1 * wy1) * i11;
}

int launch_bilinear_sampler(const float* imgs, const float* coords, float* out, int N, int H, int W, int C, int Ht, int Wt,
                            cudaStream_t st) {
  const int64_t total = (int64_t)N * Ht * Wt * C;
  bilinear_sampler_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(imgs, coords, out, N, H, W, C, Ht, Wt);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

// Stand-alone KFNet/util.py:ApplyTransform (12-40): (T @ [x,y,z,1])[:3] for every pixel of a [npix,3] map.
__global__ void apply_transform_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t npix, Xform X) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float x = in[i * 3 + 0], y = in[i * 3 + 1], z = in[i * 3 + 2];
  out[i * 3 + 0] = ((X.m[0] * x + X.m[1] * y) + X.m[2] * z) + X.m[3];
  out[i * 3 + 1] = ((X.m[4] * x + X.m[5] * y) + X.m[6] * z) + X.m[7];
  out[i * 3 + 2] = ((X.m[8] * x + X.m[9] * y) + X.m[10] * z) + X.m[11];
}

int launch_apply_transform(const float* in, float* out, int64_t npix, const float* T16_host, cudaStream_t st) {
  Xform X;
  for (int i = 0; i < 16; ++i) X.m[i] = T16_host[i];
  apply_transform_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(in, out, npix, X);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

}  // namespace kfn
