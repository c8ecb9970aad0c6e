// SCoordNet conv1a (+ tower feat1) and conv1b as ONE tcgen05 kernel (KFN_MODE_BF16, opt-in: KFN_TC_FUSE_CONV1=1).
//
// Why: after r01j the two full-resolution layers cost 322 us + 374 us per 16 frames; conv1a's 64-channel output is
// written to HBM (629 MB) only to be read back by conv1b (706 MB), and conv1b's own tensor time is ~250 us.  Here
// conv1a's output never leaves the SM (cnn_wrapper/SCoordNet.py:18-21: conv1a 3->64, conv1b 64->64, both 3x3 SAME + ReLU):
//
//   tile = 8 (w) x 16 (h) output pixels of conv1b  ->  needs conv1a on the 10 x 18 halo block (180 pixels)
//   producers (6 warps)   u8 image -> 32-wide im2col rows (u8 - 128, bf16) of the 180 halo pixels -> smem A1 (64B swizzle)
//   MMA-1                 [256 x 32] x [32 x 80] -> TMEM acc1 (two 128-row tiles; cols 0..63 conv1a, 64..79 feat1)
//   epilogue-1 (6 warps)  acc1 -> +bias, ReLU, bf16; pixels outside the image -> 0 (that IS conv1b's SAME padding)
//                         -> smem A2 rows [y][x] (10 per image row, 128B swizzle): exactly the haloed block the
//                         halo-tile MMA path of conv_tc.cu reads; feat1 columns of the 128 interior pixels -> HBM
//   MMA-2                 nine shifted-descriptor taps x 4 k-steps on A2 against resident conv1b weights -> TMEM acc2
//   epilogue-2 (8 warps, two groups alternating tiles)  acc2 -> +bias, ReLU, bf16 -> HBM
//
// One MMA-issuing thread interleaves MMA-1 of tile j with MMA-2 of tile j-1, so epilogue-1 of a tile overlaps the
// 36 MMAs of the previous one.  All 512 TMEM columns: acc1 2 stages x (2 x 96), acc2 2 stages x 64.
// STATUS: written at the end of round 1 without GPU time left -- compiles for sm_100a, NOT yet run; the layer test
// (tests/test_gpu_conv_layers.py::test_fused_conv1ab) is skipped unless KFN_TC_FUSE_CONV1=1.
#include <cstring>

#include "common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace kfn {

namespace {

constexpr int kFThreads = 704;            // 22 warps: 0-5 producers, 6 MMA, 7 spare, 8-13 epilogue-1, 14-21 epilogue-2
constexpr int kHaloW = 10, kHaloH = 18, kHaloRows = kHaloW * kHaloH;   // 180
constexpr uint32_t kW2Tile = 64 * 128;                                  // one tap of conv1b: [64 cout][64 cin] bf16
constexpr uint32_t kW2Bytes = 9 * kW2Tile;                              // 73,728
constexpr uint32_t kW1Bytes = 80 * 64;                                  // [80 cout][32 k] bf16 = 5,120
constexpr uint32_t kA1Stage = 256 * 64;                                 // 16,384 (192 rows written)
constexpr uint32_t kA2Stage = 23 * 1024;                                // 180 rows x 128 B = 23,040 -> 23,552
constexpr uint32_t kOffW1 = kW2Bytes, kOffA1 = kOffW1 + kW1Bytes, kOffA2 = kOffA1 + 2 * kA1Stage;
constexpr uint32_t kFusedSmem = kOffA2 + 3 * kA2Stage + 1024;           // 183,296 incl. alignment slack
constexpr uint32_t kAcc1Stride = 192, kAcc1M = 96, kAcc2Base = 384, kAcc2Stride = 64;
static_assert(kOffW1 % 1024 == 0 && kOffA1 % 1024 == 0 && kOffA2 % 1024 == 0, "swizzled regions must be 1024-aligned");

struct FusedParams {
  const uint8_t* img;       // [N,H,W,3]
  void* out;                // bf16 [N,H,W,64]  (conv1b output)
  void* feat1;              // bf16 [N,H,W,16]  or null
  const float* bias1;       // [80]: conv1a (64) | feat1 (16)
  const float* bias2;       // [64]
  int N, H, W, tiles_w, tiles_h, total_tiles;
  uint32_t idesc1, idesc2;
  uint32_t hi_sw64, hi_sw128, hi_sw128_halo;   // descriptor upper words
  uint32_t tap_off16[9];
};

struct FTile {   // contiguous balanced tile range per CTA, tw fastest
  int tile, end, tw, th, n;
  __device__ __forceinline__ void init(const FusedParams& p) {
    const int base = p.total_tiles / (int)gridDim.x, rem = p.total_tiles % (int)gridDim.x, b = (int)blockIdx.x;
    tile = b * base + (b < rem ? b : rem);
    end = tile + base + (b < rem ? 1 : 0);
    tw = tile % p.tiles_w;
    const int r = tile / p.tiles_w;
    th = r % p.tiles_h;
    n = r / p.tiles_h;
  }
  __device__ __forceinline__ bool valid() const { return tile < end; }
  __device__ __forceinline__ void next(const FusedParams& p) {
    ++tile;
    if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++n; } }
  }
};

__global__ void __launch_bounds__(kFThreads, 1)
conv1ab_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a1_full[2], a1_empty[2], t1_full[2], t1_empty[2];
  __shared__ __align__(8) uint64_t a2_full[3], a2_empty[3], t2_full[2], t2_empty[2];
  __shared__ __align__(8) uint64_t wbar;
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&a1_full[s]), 6);   // producer warps
      mbar_init(smem_u32(&a1_empty[s]), 1);
      mbar_init(smem_u32(&t1_full[s]), 1);
      mbar_init(smem_u32(&t1_empty[s]), 6);  // epilogue-1 warps
      mbar_init(smem_u32(&t2_full[s]), 1);
      mbar_init(smem_u32(&t2_empty[s]), 4);  // one epilogue-2 group
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(smem_u32(&a2_full[s]), 6);   // epilogue-1 warps
      mbar_init(smem_u32(&a2_empty[s]), 1);
    }
    mbar_init(smem_u32(&wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // after the TMEM allocation (see conv_tc.cu)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < 6) {
    // ================================================================== im2col producers: halo pixel r = warp*32 + lane
    const int r = warp * 32 + lane;
    const int hy = r / kHaloW, hx = r - hy * kHaloW;
    uint8_t* const row_ptr = smem_gen + kOffA1 + (uint32_t)r * 64u;
    const int sw = (r >> 1) & 3;
    const int H = p.H, W = p.W;
    int j = 0;
    FTile it;
    for (it.init(p); it.valid(); it.next(p), ++j) {
      const int y = it.th * 16 - 1 + hy, x = it.tw * 8 - 1 + hx;
      const bool inside = r < kHaloRows && y >= 0 && y < H && x >= 0 && x < W;
      float e[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) e[k] = 0.f;
      if (inside) {
        const uint8_t* src0 = p.img + (((int64_t)it.n * H + (y - 1)) * W + (x - 1)) * 3;
        const bool xl = x > 0, xr = x + 1 < W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1;
          const bool rowv = yy >= 0 && yy < H;
          const uint8_t* src = src0 + (int64_t)ky * W * 3;
#pragma unroll
          for (int jj = 0; jj < 9; ++jj) {
            const bool v = rowv && (jj < 3 ? xl : (jj >= 6 ? xr : true));
            e[ky * 9 + jj] = v ? (float)__ldg(src + jj) - 128.0f : 0.f;
          }
        }
      }
      uint4 c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        c[q] = make_uint4(pack_bf16(e[8 * q], e[8 * q + 1]), pack_bf16(e[8 * q + 2], e[8 * q + 3]),
                          pack_bf16(e[8 * q + 4], e[8 * q + 5]), pack_bf16(e[8 * q + 6], e[8 * q + 7]));
      const int s1 = j & 1;
      mbar_wait(smem_u32(&a1_empty[s1]), (((uint32_t)j >> 1) & 1u) ^ 1u);
      uint8_t* dst = row_ptr + (uint32_t)s1 * kA1Stage;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = c[q];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&a1_full[s1]));
    }
  } else if (warp == 6) {
    // ================================================================== MMA issuer
    if (elect_one(lane)) {
      const uint32_t wb = smem_u32(&wbar);
      mbar_expect_tx(wb, kW2Bytes + kW1Bytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_3d(smem_base + (uint32_t)tap * kW2Tile, &tmW2, wb, 0, 0, tap);
      tma_load_3d(smem_base + kOffW1, &tmW1, wb, 0, 0, 0);
      mbar_wait(wb, 0);
      const uint64_t fx64 = ((uint64_t)p.hi_sw64 << 32) | (1ull << 16);
      const uint64_t fx128 = ((uint64_t)p.hi_sw128 << 32) | (1ull << 16);
      const uint64_t fx128h = ((uint64_t)p.hi_sw128_halo << 32) | (1ull << 16);
      const uint64_t b1desc = fx64 | (uint64_t)((smem_base + kOffW1) >> 4);
      const uint64_t b2desc0 = fx128 | (uint64_t)(smem_base >> 4);
      int n_my = 0;
      { FTile t; t.init(p); n_my = t.end - t.tile; }
      for (int j = 0; j <= n_my; ++j) {
        if (j < n_my) {   // MMA-1 of tile j: conv1a (+feat1) on the 256-row (180 used) halo block
          const int s1 = j & 1;
          const uint32_t p1 = ((uint32_t)j >> 1) & 1u;
          mbar_wait(smem_u32(&t1_empty[s1]), p1 ^ 1u);
          mbar_wait(smem_u32(&a1_full[s1]), p1);
          tc_fence_after();
          const uint64_t a1desc = fx64 | (uint64_t)((smem_base + kOffA1 + (uint32_t)s1 * kA1Stage) >> 4);
          const uint32_t d1 = tmem_base + (uint32_t)s1 * kAcc1Stride;
          umma_kblock<2>(d1, a1desc, b1desc, p.idesc1, 0u);
          umma_kblock<2>(d1 + kAcc1M, a1desc + (uint64_t)((128u * 64u) >> 4), b1desc, p.idesc1, 0u);
          umma_commit(smem_u32(&a1_empty[s1]));
          umma_commit(smem_u32(&t1_full[s1]));
        }
        if (j >= 1) {     // MMA-2 of tile j-1: conv1b, nine shifted taps on the bf16 halo block epilogue-1 wrote
          const int i = j - 1;
          const int s2 = i % 3, t = i & 1;
          const uint32_t p2 = (uint32_t)(i / 3) & 1u, pt = ((uint32_t)i >> 1) & 1u;
          mbar_wait(smem_u32(&t2_empty[t]), pt ^ 1u);
          mbar_wait(smem_u32(&a2_full[s2]), p2);
          tc_fence_after();
          const uint64_t a2desc = fx128h | (uint64_t)((smem_base + kOffA2 + (uint32_t)s2 * kA2Stage) >> 4);
          const uint32_t d2 = tmem_base + kAcc2Base + (uint32_t)t * kAcc2Stride;
          uint64_t bdesc = b2desc0;
          umma_kblock<4>(d2, a2desc + (uint64_t)p.tap_off16[0], bdesc, p.idesc2, 0u);
#pragma unroll
          for (int tap = 1; tap < 9; ++tap) {
            bdesc += (uint64_t)(kW2Tile >> 4);
            umma_kblock<4>(d2, a2desc + (uint64_t)p.tap_off16[tap], bdesc, p.idesc2, 1u);
          }
          umma_commit(smem_u32(&a2_empty[s2]));
          umma_commit(smem_u32(&t2_full[t]));
        }
      }
    }
  } else if (warp >= 8 && warp < 14) {
    // ================================================================== epilogue-1: acc1 -> bf16 halo block in smem
    const int q = warp & 3, m = (warp - 8) >> 2;          // warps 8..11: rows 0..127, 12..13: rows 128..191
    const int r = m * 128 + q * 32 + lane;
    const bool rvalid = r < kHaloRows;
    const int hy = r / kHaloW, hx = r - hy * kHaloW;
    const bool interior = rvalid && hy >= 1 && hy <= 16 && hx >= 1 && hx <= 8;   // the tile's own 8x16 pixels
    const uint32_t row_off = kOffA2 + (uint32_t)r * 128u;
    const int sw7 = r & 7;
    int j = 0;
    FTile it;
    for (it.init(p); it.valid(); it.next(p), ++j) {
      const int s1 = j & 1, s2 = j % 3;
      const uint32_t p1 = ((uint32_t)j >> 1) & 1u, p2 = (uint32_t)(j / 3) & 1u;
      const int y = it.th * 16 - 1 + hy, x = it.tw * 8 - 1 + hx;
      const bool inside = rvalid && y >= 0 && y < p.H && x >= 0 && x < p.W;
      mbar_wait(smem_u32(&t1_full[s1]), p1);
      tc_fence_after();
      mbar_wait(smem_u32(&a2_empty[s2]), p2 ^ 1u);          // MMA-2 of tile j-3 has finished reading this stage
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)s1 * kAcc1Stride + (uint32_t)m * kAcc1M;
      uint8_t* const dst = smem_gen + row_off + (uint32_t)s2 * kA2Stage;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t rr[32];
        tmem_ld32(taddr + blk * 32, rr);
        const float4* b4 = reinterpret_cast<const float4*>(p.bias1 + blk * 32);
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b = __ldg(b4 + g);
          v[4 * g + 0] = __uint_as_float(rr[4 * g + 0]) + b.x;      // the ReLU rides on the bf16 conversion (F2FP.RELU)
          v[4 * g + 1] = __uint_as_float(rr[4 * g + 1]) + b.y;
          v[4 * g + 2] = __uint_as_float(rr[4 * g + 2]) + b.z;
          v[4 * g + 3] = __uint_as_float(rr[4 * g + 3]) + b.w;
        }
        if (rvalid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o = make_uint4(pack_bf16_relu(v[8 * g + 0], v[8 * g + 1]), pack_bf16_relu(v[8 * g + 2], v[8 * g + 3]),
                                 pack_bf16_relu(v[8 * g + 4], v[8 * g + 5]), pack_bf16_relu(v[8 * g + 6], v[8 * g + 7]));
            if (!inside) o = make_uint4(0u, 0u, 0u, 0u);   // zero padding of conv1b's input (tf.layers.conv2d SAME)
            *reinterpret_cast<uint4*>(dst + (((blk * 4 + g) ^ sw7) << 4)) = o;
          }
        }
      }
      {   // feat1 = columns 64..79, wanted only for the tile's own pixels
        float f[16];
        tmem_ld16(taddr + 64, f);
        if (p.feat1 != nullptr && interior) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias1 + 64);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b = __ldg(b4 + g);
            f[4 * g + 0] = fmaxf(f[4 * g + 0] + b.x, 0.f); f[4 * g + 1] = fmaxf(f[4 * g + 1] + b.y, 0.f);
            f[4 * g + 2] = fmaxf(f[4 * g + 2] + b.z, 0.f); f[4 * g + 3] = fmaxf(f[4 * g + 3] + b.w, 0.f);
          }
          uint4* fo = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.feat1) + (((int64_t)it.n * p.H + y) * p.W + x) * 32);
          fo[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          fo[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&t1_empty[s1]));
        mbar_arrive(smem_u32(&a2_full[s2]));
      }
    }
  } else if (warp >= 14) {
    // ================================================================== epilogue-2: acc2 -> conv1b output in HBM
    const int q = warp & 3, grp = (warp - 14) >> 2;        // two groups of four warps alternate tiles
    const int row = q * 32 + lane;                         // tile rows ordered [h][w], w = 8
    const int rw = row & 7, rh = row >> 3;
    int j = 0;
    FTile it;
    for (it.init(p); it.valid(); it.next(p), ++j) {
      if ((j & 1) != grp) continue;
      const int t = j & 1;
      const uint32_t pt = ((uint32_t)j >> 1) & 1u;
      mbar_wait(smem_u32(&t2_full[t]), pt);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kAcc2Base + (uint32_t)t * kAcc2Stride;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) +
                      ((((int64_t)it.n * p.H + it.th * 16 + rh) * p.W) + it.tw * 8 + rw) * 128;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t rr[32];
        tmem_ld32(taddr + blk * 32, rr);
        const float4* b4 = reinterpret_cast<const float4*>(p.bias2 + blk * 32);
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b = __ldg(b4 + g);
          v[4 * g + 0] = __uint_as_float(rr[4 * g + 0]) + b.x;      // the ReLU rides on the bf16 conversion (F2FP.RELU)
          v[4 * g + 1] = __uint_as_float(rr[4 * g + 1]) + b.y;
          v[4 * g + 2] = __uint_as_float(rr[4 * g + 2]) + b.z;
          v[4 * g + 3] = __uint_as_float(rr[4 * g + 3]) + b.w;
        }
        uint4* o = reinterpret_cast<uint4*>(orow + blk * 64);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          o[g] = make_uint4(pack_bf16_relu(v[8 * g + 0], v[8 * g + 1]), pack_bf16_relu(v[8 * g + 2], v[8 * g + 3]),
                            pack_bf16_relu(v[8 * g + 4], v[8 * g + 5]), pack_bf16_relu(v[8 * g + 6], v[8 * g + 7]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&t2_empty[t]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

bool tc_fuse_conv1_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_FUSE_CONV1"); return e && e[0] && e[0] != '0'; }();   // opt-in until validated
  return on;
}

bool conv1ab_shape_ok(int H, int W) { return H % 16 == 0 && W % 8 == 0 && H >= 16 && W >= 8; }

int prepare_conv1ab(const uint8_t* img, int N, int H, int W, const void* w1_packed, const float* bias1, const void* w2_packed,
                    const float* bias2, void* out, void* feat1, int num_sms, Conv1abPlan* plan) {
  PFN_encodeTiled encode = get_encode();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)"); return KFN_ERR_CUDA; }
  KFN_REQUIRE(conv1ab_shape_ok(H, W) && N > 0, "fused conv1a+conv1b needs H % 16 == 0 and W % 8 == 0");
  KFN_REQUIRE((reinterpret_cast<uintptr_t>(w1_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2_packed) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(feat1) & 15) == 0,
              "16-byte aligned buffers");
  static_assert(sizeof(FusedParams) + 2 * sizeof(CUtensorMap) <= sizeof(Conv1abPlan::blob), "Conv1abPlan blob too small");
  memset(plan->blob, 0, sizeof(plan->blob));
  CUtensorMap* tmW1 = reinterpret_cast<CUtensorMap*>(plan->blob);
  CUtensorMap* tmW2 = tmW1 + 1;
  FusedParams& p = *reinterpret_cast<FusedParams*>(tmW2 + 1);
  cuuint32_t ones[3] = {1, 1, 1};
  {   // conv1a|feat1 weights [1 tap][80][32]
    cuuint64_t dim[3] = {32, 80, 1}, str[2] = {64, 64 * 80};
    cuuint32_t box[3] = {32, 80, 1};
    CUresult r = encode(tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w1_packed), dim, str, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(W1) failed: " + std::to_string((int)r)); return KFN_ERR_CUDA; }
  }
  {   // conv1b weights [9 taps][64][64]
    cuuint64_t dim[3] = {64, 64, 9}, str[2] = {128, 128 * 64};
    cuuint32_t box[3] = {64, 64, 1};
    CUresult r = encode(tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w2_packed), dim, str, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(W2) failed: " + std::to_string((int)r)); return KFN_ERR_CUDA; }
  }
  p.img = img; p.out = out; p.feat1 = feat1; p.bias1 = bias1; p.bias2 = bias2;
  p.N = N; p.H = H; p.W = W; p.tiles_w = W / 8; p.tiles_h = H / 16; p.total_tiles = p.tiles_w * p.tiles_h * N;
  auto idesc = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  p.idesc1 = idesc(80); p.idesc2 = idesc(64);
  auto hi = [](uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); };
  p.hi_sw64 = hi(8 * 64, 4u);              // 64B swizzle, 8-row groups of 64-byte rows
  p.hi_sw128 = hi(8 * 128, 2u);            // 128B swizzle, canonical
  p.hi_sw128_halo = hi(kHaloW * 128, 2u);  // 8-row groups pitched by one haloed image row (10 pixels)
  for (int t = 0; t < 9; ++t) p.tap_off16[t] = (uint32_t)(((t / 3) * kHaloW + (t % 3)) * 128) >> 4;   // (1+dy, 1+dx) = (ky, kx)
  plan->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  plan->smem = kFusedSmem;
  return KFN_OK;
}

int launch_conv1ab(const Conv1abPlan& plan, cudaStream_t st) {
  static bool attr_set_dev[64] = {false};
  int dev_ = 0;
  KFN_CUDA_OK(cudaGetDevice(&dev_));
  if (!attr_set_dev[dev_ & 63]) {
    KFN_CUDA_OK(cudaFuncSetAttribute(conv1ab_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
    attr_set_dev[dev_ & 63] = true;
  }
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(plan.blob);
  const FusedParams& p = *reinterpret_cast<const FusedParams*>(tm + 2);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv1ab_fused_kernel, tm[0], tm[1], p));
  KFN_LAUNCH_OK();
  return KFN_OK;
}

}  // namespace kfn
