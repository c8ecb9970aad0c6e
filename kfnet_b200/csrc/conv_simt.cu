// fp32 CUDA-core implicit-GEMM convolution (KFN_MODE_FP32): the numerics-contract path.
// Every product and sum is fp32 like the reference's TF1 kernels; only the summation order over
// K = taps x Cin differs.  Covers every conv / conv2d_transpose phase / dense layer of the path
// through the ConvGeom description (common.cuh).
//
// Tiling: CTA = 128 output rows (m-grid pixels) x 64 output channels, 256 threads, 8x4 register
// tile per thread, K consumed in (tap, 16-channel) chunks through shared memory with register
// prefetch of the next chunk.
#include "common.cuh"

namespace kfn {

namespace {
constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

struct SimtParams {
  ConvGeom g;
  const float* in;
  const float* w;      // [ntaps][Cin][CoutP]
  int CoutP;
  const float* bias;   // [Cout]
  float* out;
  int64_t M;           // N*MH*MW
};

__global__ void __launch_bounds__(NT) conv_simt_kernel(const SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  __shared__ int s_ipix[BM];      // input pixel index of tap (0,0), -1 if row >= M
  __shared__ short s_iy[BM], s_ix[BM];
  __shared__ int64_t s_opix[BM];  // output pixel index, -1 if row >= M

  const ConvGeom& g = p.g;
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  if (tid < BM) {
    int64_t m = m0 + tid;
    if (m < p.M) {
      int mx = (int)(m % g.MW);
      int my = (int)((m / g.MW) % g.MH);
      int n = (int)(m / ((int64_t)g.MW * g.MH));
      s_iy[tid] = (short)(my * g.istride);
      s_ix[tid] = (short)(mx * g.istride);
      s_ipix[tid] = (n * g.IH + my * g.istride) * g.IW + mx * g.istride;
      s_opix[tid] = ((int64_t)n * g.OH + my * g.omul + g.oaddy) * g.OW + mx * g.omul + g.oaddx;
    } else {
      s_iy[tid] = 0; s_ix[tid] = 0; s_ipix[tid] = -1; s_opix[tid] = -1;
    }
  }
  __syncthreads();

  // global -> register staging assignments
  const int a_k = tid & 15;          // channel within the chunk
  const int a_m = tid >> 4;          // rows a_m + 16*i
  const int b_k = tid >> 4;          // 0..15
  const int b_n = (tid & 15) * 4;    // 0..60

  const int tm = (tid >> 4) * 8;     // thread's 8 rows
  const int tn = (tid & 15) * 4;     // thread's 4 channels

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int chunks_per_tap = (g.Cin + BK - 1) / BK;
  const int nchunks = g.ntaps * chunks_per_tap;

  float a_reg[8];
  float4 b_reg;

  auto load_chunk = [&](int chunk) {
    const int tap = chunk / chunks_per_tap;
    const int c0 = (chunk - tap * chunks_per_tap) * BK;
    const int dy = g.dy[tap], dx = g.dx[tap];
    const int c = c0 + a_k;
    const bool cok = c < g.Cin;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = a_m + 16 * i;
      const int ip = s_ipix[r];
      const int iy = s_iy[r] + dy, ix = s_ix[r] + dx;
      const bool ok = cok && (ip >= 0) && (iy >= 0) && (iy < g.IH) && (ix >= 0) && (ix < g.IW);
      a_reg[i] = ok ? __ldg(p.in + ((int64_t)ip + dy * g.IW + dx) * g.in_cs + c) : 0.f;
    }
    const int kc = c0 + b_k;
    const int n = n0 + b_n;
    if (kc < g.Cin && n < p.CoutP)
      b_reg = __ldg(reinterpret_cast<const float4*>(p.w + ((int64_t)tap * g.Cin + kc) * p.CoutP + n));
    else
      b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
  };

  load_chunk(0);
  for (int chunk = 0; chunk < nchunks; ++chunk) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_k][a_m + 16 * i] = a_reg[i];
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = b_reg;
    __syncthreads();
    if (chunk + 1 < nchunks) load_chunk(chunk + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: bias, ReLU, optional exp on channel 3, store with the output pixel / channel mapping
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int n = n0 + tn + j;
    bv[j] = (n < g.Cout) ? __ldg(p.bias + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t op = s_opix[tm + i];
    if (op < 0) continue;
    int cbase = n0 + tn;
    int64_t ph_off = 0;
    if (g.phase_cout) {                     // merged conv2d_transpose phases (4-channel groups never straddle a phase)
      const int ph = cbase / g.phase_cout;
      cbase -= ph * g.phase_cout;
      ph_off = ((int64_t)(ph >> 1) * g.OW + (ph & 1)) * g.out_cs;
    }
    float* dst = p.out + op * g.out_cs + ph_off + g.out_coff + cbase;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j] + bv[j];
      if (g.relu) x = fmaxf(x, 0.f);
      if (g.epi == EPI_EXP_CH3 && (n0 + tn + j) == 3) x = expf(x);
      v[j] = x;
    }
    if (g.accum) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tn + j < g.Cout) dst[j] += v[j];
    } else if (n0 + tn + 3 < g.Cout && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tn + j < g.Cout) dst[j] = v[j];
    }
  }
}
}  // namespace

int launch_conv_simt(const ConvGeom& g, const float* in, const float* w, int CoutP, const float* bias, float* out,
                     cudaStream_t st) {
  SimtParams p;
  p.g = g; p.in = in; p.w = w; p.CoutP = CoutP; p.bias = bias; p.out = out;
  p.M = (int64_t)g.N * g.MH * g.MW;
  KFN_REQUIRE(p.M > 0 && g.ntaps > 0 && g.ntaps <= 9, "empty conv");
  KFN_REQUIRE((int64_t)g.N * g.IH * g.IW < (1ll << 31), "input pixel index overflows int32");
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((g.Cout + BN - 1) / BN));
  conv_simt_kernel<<<grid, NT, 0, st>>>(p);
  KFN_LAUNCH_OK();
  return KFN_OK;
}

}  // namespace kfn
