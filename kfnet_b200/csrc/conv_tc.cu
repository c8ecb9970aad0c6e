// tcgen05 / TMA implicit-GEMM convolution for sm_100a (KFN_MODE_BF16 and KFN_MODE_BF16X3).
//
// GEMM view of one conv (ConvGeom, common.cuh):  D[M, Cout] = sum_{tap, c} A_tap[M, c] * W[tap][Cout][c]
//   M tile  = 128 m-grid rows = a (bw x bh x bn) box of pixels, bw*bh*bn = 128, chosen per layer on the host;
//   A tiles = one TMA tiled load per (tap, KC-channel chunk): a 5-D box of the NHWC activation tensor
//             starting at the tap-shifted pixel; out-of-range coordinates are zero-filled by TMA, which IS the
//             SAME padding of the reference's tf.layers.conv2d.  Stride-2 convs read a
//             parity-split 5-D view (c|wpar, w/2, hpar, h/2, n) of the same buffer, so no element strides are needed;
//             in HALO mode the tile's haloed block is loaded once and the taps are shifted-descriptor MMAs (TcParams);
//   B tiles = [BN x KC] K-major slices of the packed weights [tap][CoutP][Cin] (resident in smem in halo mode);
//   both land in shared memory in the canonical K-major swizzled layout (128B/64B/32B for KC = 64/32/16 bf16)
//   that tcgen05.mma consumes through shared-memory descriptors;
//   accumulators live in TMEM (2 stages, allocated by need), drained by 8 epilogue warps with tcgen05.ld while the
//   next tile's MMAs run; bias + ReLU (+ exp / l2-normalise / limb split / conv2d_transpose phase scatter) fused.
//
// One persistent CTA (or two / three co-resident ones for small layers) per SM, 320 threads: warp 0 = TMA producer,
// warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4; the two warps of a
// quadrant split the column blocks).  WIDE variant (576 threads) for one-CTA layers whose tiles carry few MMAs
// (conv1a, conv1b): four accumulator stages and four epilogue groups of 4 warps, group g draining every 4th tile
// whole -- the per-tile epilogue is a ~350-instruction dependent chain (profiles/r01h: 3100 cycles per tile against
// 1700 cycles of MMAs), so four tiles are drained concurrently instead of one.
// DESIGN.md section 4.1 has the measurements behind each design choice.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"

namespace kfn {

namespace {

constexpr int kBlockM = 128;
constexpr int kMaxStages = 12;
constexpr int kThreads = 320;                           // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kThreadsWide = 576;                       // WIDE: warps 2..17 = four epilogue groups
constexpr int kEpiWarps = 8;
constexpr int kMaxAcc = 4;                              // TMEM accumulator stages (2, or 4 in WIDE mode)
constexpr uint32_t kEpiStageBytes = kEpiWarps * 2048;     // per epilogue warp: 32 rows x 64 B staging for coalesced stores

struct __align__(8) TcParams {
  int tiles_w, tiles_h, tiles_n, n_tiles;
  int lbw, lbh;                 // log2 of the pixel box (bn = 128 >> (lbw + lbh))
  int N, MH, MW;
  int istride, in_cs;
  int ntaps;
  int dy[9], dx[9];
  int kchunks, KC, BN;
  uint32_t desc_hi;             // upper 32 bits of the UMMA smem descriptor (SBO, version, swizzle mode)
  uint32_t idesc;
  int stages;
  uint32_t stage_bytes, a_bytes, b_bytes;
  int OH, OW, out_cs, out_coff, Cout, omul, oaddy, oaddx, relu, epi, out_bf16;
  void* out;
  const float* bias;            // [CoutP] zero padded
  // halo mode (stride-1 3x3 layers whose weights fit in smem): the (8 x bh x bn) pixel tile's haloed input block
  // is loaded ONCE per channel chunk (rows ordered [y][n][x], 10 pixels per image row) and every tap is an MMA
  // whose A descriptor starts at a shifted row with a 10-row group pitch (tools/umma_probe.cu: the 128B swizzle
  // is a function of absolute smem address bits, so shifted starts / non-canonical SBO read the right rows).
  int halo, lbn, row_bytes;
  uint32_t desc_hi_halo;        // A descriptor upper word with SBO = 10 rows
  uint32_t w_tile_bytes, w_bytes_total;
  int pf_dist;                  // L2 prefetch distance in tiles (0 = off)
  // bf16x3: k-chunks = n_terms x cpk; term t multiplies A limb la[t] with W limb lw[t]
  int n_terms, cpk, in_ls, w_cin, out_limbs, out_ls;
  int epi_warps;                // 8 (320-thread CTA), 4 (192-thread CTA, three CTAs per SM) or 16 (576-thread WIDE CTA)
  int lg_groups;                // WIDE: log2 of the number of 4-warp epilogue groups (tile j -> group j % G); -1 = column split
  int lg_acc;                   // log2 of the number of TMEM accumulator stages (1, or 2 in WIDE mode)
  uint32_t tmem_cols, acc_stride;   // TMEM columns allocated by this CTA (power of 2 >= 32) and columns per accumulator stage
  int phase_cout;               // merged conv2d_transpose phases (see ConvGeom)
  int split_col, out2_cs;       // dual output: columns >= split_col -> out2 (bf16, generic 16-column path)
  void* out2;
  int x3acc;                    // 1: four 128-column TMEM accumulators (3 main-term splits + corrections), single-buffered
  int la[6], lw[6];
  // halo mode with limbs (bf16x3 / fp16x2): the haloed block is loaded once per (A limb, channel chunk) and every tap x
  // every W limb that pairs with that A limb is a shifted-descriptor MMA on the resident weights
  int a_chunks, w_chunks, n_limbs_a, n_limbs_w;
  uint32_t term_mask[3];        // bit lw of term_mask[la]: the product (A limb la) x (W limb lw) is accumulated
  int strided;                  // interleaved tile assignment (see TileIter)
  int pair;                     // CTA-pair plan (conv_tc_kernel_pair): 256-row tiles, each CTA loads its 128 rows of A and 128 of the 256 rows of B
  int x3_fused;                 // non-halo limb modes: one stage holds ALL limbs of A and W for a (tap, chunk) -- every product that uses a
                                // tile reads it from the same stage (4 tile loads per 3 products in fp16x2 instead of 6)
  int f16;                      // limbs are fp16 (KFN_MODE_FP16X2): limb 1 holds the residual scaled by 2^11
  float corr_scale;             // factor applied to the correction accumulator in the epilogue (1, or 2^-11 for fp16 limbs)
  uint32_t epi_stage_off;       // byte offset (from the 1024-aligned smem base) of the epilogue staging area
  // fused first layer (BF16 mode): the A tile is the 32-wide im2col row of the 3x3x3 window, built in shared memory by
  // four producer warps straight from the u8 image (SCoordNet.py:34-37 preprocess: u8 - 128 is exact in bf16, the
  // 0.00625 scale is folded into the packed weights; taps outside the image are 0 = SAME padding of the preprocessed
  // tensor); the [BN x 32] weight tile is resident.  No im2col tensor ever exists in HBM.
  int im2col;
  int ph16;                     // merged deconv with four 16-channel phases: one-wait epilogue (upconv0)
  int wres;                     // non-halo layers: every (tap, chunk) weight tile resident at the smem base, stages hold only A
  int par_issue;                // > 0: warp-parallel TMA issue, this many k-blocks (lanes) per pass (<= stages)
  const uint8_t* img;
  uint32_t tap_off16[9];        // halo: ((1+dy) * 10*bn + (1+dx)) * row_bytes / 16 per tap
  // fused difference volume (U-Net conv0, BF16 mode): the halo tile's rows are diff[p,i,j,:] = f_cur[p,:] - f_prev[p + (j-4, i-4),:]
  // (KFNet.BuildCoordVolume, KFNet.py:343-359; 0 outside the 60x80 map, 0 in the window's halo ring = SAME padding),
  // rounded to bf16 exactly like diff_volume_kernel, built in shared memory by seven producer warps.
  int diffp;
  const float* f_prev;
  const float* f_cur;
  int dGH, dGW;
};

// ------------------------------------------------------------------------------------------ tile iteration
// Each persistent CTA owns a contiguous, balanced range of tiles (n-tile fastest, then w, h, n) and walks it with
// incremental coordinate updates: the integer divisions of the decode happen once per CTA, not once per tile per
// role (profiles/r01c: the per-tile div/mod + pointer arithmetic was ~1900 cycles on the epilogue's critical path).
struct TileIter {
  int tile, end, nt, tw, th, tn, step;
  __device__ __forceinline__ void decode(const TcParams& p) {
    nt = tile % p.n_tiles;
    int mt = tile / p.n_tiles;
    tw = mt % p.tiles_w; mt /= p.tiles_w;
    th = mt % p.tiles_h;
    tn = mt / p.tiles_h;
  }
  __device__ __forceinline__ void init(const TcParams& p, int total) {
    if (p.strided) {
      // interleaved assignment: CTA b walks tiles b, b + G, b + 2G, ...  At any moment the G CTAs of the grid work inside a window of
      // G consecutive tiles (neighbouring pixels, all N tiles of the same pixels together), so a tile's halo and the operand tiles its
      // neighbours need are fetched from DRAM once and hit in L2 for everyone else; with contiguous per-CTA ranges the grid touches
      // 148 distant regions at once.  Costs a few integer divisions per tile: used for the layers whose tiles carry thousands of MMA cycles.
      tile = (int)blockIdx.x; end = total; step = (int)gridDim.x;
      if (tile < end) decode(p);
      return;
    }
    const int base = total / (int)gridDim.x, rem = total % (int)gridDim.x, b = (int)blockIdx.x;
    tile = b * base + (b < rem ? b : rem);
    end = tile + base + (b < rem ? 1 : 0);
    step = 1;
    decode(p);
  }
  __device__ __forceinline__ bool valid() const { return tile < end; }
  __device__ __forceinline__ void next(const TcParams& p) {
    if (step != 1) {
      tile += step;
      if (tile < end) decode(p);
      return;
    }
    ++tile;
    if (++nt == p.n_tiles) { nt = 0; if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } } }
  }
};

// ------------------------------------------------------------------------------------------ MMA issuer roles
// Descriptors are 64-bit values advanced with adds of precomputed constants (start address field is in 16-byte
// units and never overflows its 14 bits for addresses < 256 KB), so one issued MMA costs a couple of uniform
// instructions instead of a descriptor rebuild.
template <int KSTEPS>
__device__ __forceinline__ void mma_role(const TcParams& p, uint32_t smem_base, uint32_t tmem_base, int total_tiles, int kblocks,
                                         uint64_t* s_full, uint64_t* s_empty, uint64_t* s_tfull, uint64_t* s_tempty, uint64_t* s_wbar) {
  const uint64_t desc_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
  const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
  // resident weights (TcParams::wres): tile of k-block kb at smem_base + kb * w_tile_bytes, the A stages behind them
  const uint32_t wres = (uint32_t)p.wres;
  if (wres) mbar_wait(smem_u32(s_wbar), 0);
  const uint64_t a0 = desc_fixed | (uint64_t)((smem_base + (wres ? p.w_bytes_total : 0u)) >> 4);
  const uint64_t bres0 = desc_fixed | (uint64_t)(smem_base >> 4);
  const uint64_t wtile16 = (uint64_t)(p.w_tile_bytes >> 4);
  const uint64_t b_off = (uint64_t)(p.a_bytes >> 4);
  int s = 0, j = 0;
  uint32_t ph = 0;
  uint64_t adesc = a0;
  TileIter it;
  for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
    const int a = j & ((1 << p.lg_acc) - 1);
    const uint32_t aph = (uint32_t)(j >> p.lg_acc) & 1u;
    mbar_wait(smem_u32(&s_tempty[a]), aph ^ 1u);
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + (uint32_t)a * p.acc_stride;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(smem_u32(&s_full[s]), ph);
      tc_fence_after();
      umma_kblock<KSTEPS>(tmem_d, adesc, wres ? bres0 + (uint64_t)kb * wtile16 : adesc + b_off, p.idesc, kb ? 1u : 0u);
      umma_commit(smem_u32(&s_empty[s]));       // frees the smem stage once these MMAs retire
      adesc += stage16;
      if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
    }
    umma_commit(smem_u32(&s_tfull[a]));          // accumulator ready for the epilogue
  }
}

template <int KSTEPS>
__device__ __forceinline__ void mma_role_halo(const TcParams& p, uint32_t smem_base, uint32_t tmem_base, int total_tiles,
                                              uint64_t* s_full, uint64_t* s_empty, uint64_t* s_tfull, uint64_t* s_tempty,
                                              uint64_t* s_wbar) {
  mbar_wait(smem_u32(s_wbar), 0);               // resident weights have landed
  const uint64_t a_fixed = ((uint64_t)p.desc_hi_halo << 32) | (1ull << 16);
  const uint64_t b_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
  const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
  const uint64_t a0 = a_fixed | (uint64_t)((smem_base + p.w_bytes_total) >> 4);
  const uint64_t b0 = b_fixed | (uint64_t)(smem_base >> 4);
  const uint64_t wtile16 = (uint64_t)(p.w_tile_bytes >> 4);
  const uint64_t wtap16 = wtile16 * (uint64_t)p.kchunks;
  int s = 0, j = 0;
  uint32_t ph = 0;
  uint64_t adesc = a0;
  TileIter it;
  for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
    const int a = j & ((1 << p.lg_acc) - 1);
    const uint32_t aph = (uint32_t)(j >> p.lg_acc) & 1u;
    mbar_wait(smem_u32(&s_tempty[a]), aph ^ 1u);
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + (uint32_t)a * p.acc_stride;
    uint64_t bchunk = b0;
    for (int kc = 0; kc < p.kchunks; ++kc) {
      mbar_wait(smem_u32(&s_full[s]), ph);
      tc_fence_after();
      uint64_t bdesc = bchunk;
      umma_kblock<KSTEPS>(tmem_d, adesc + (uint64_t)p.tap_off16[0], bdesc, p.idesc, kc ? 1u : 0u);
      for (int tap = 1; tap < p.ntaps; ++tap) {
        bdesc += wtap16;
        umma_kblock<KSTEPS>(tmem_d, adesc + (uint64_t)p.tap_off16[tap], bdesc, p.idesc, 1u);
      }
      umma_commit(smem_u32(&s_empty[s]));
      bchunk += wtile16;
      adesc += stage16;
      if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
    }
    umma_commit(smem_u32(&s_tfull[a]));
  }
}

// fused first layer: one k-block per tile, A written by the im2col producer warps, B resident at the smem base
template <int KSTEPS>
__device__ __forceinline__ void mma_role_im2col(const TcParams& p, uint32_t smem_base, uint32_t tmem_base, int total_tiles,
                                                uint64_t* s_full, uint64_t* s_empty, uint64_t* s_tfull, uint64_t* s_tempty,
                                                uint64_t* s_wbar) {
  mbar_wait(smem_u32(s_wbar), 0);               // resident weights have landed
  const uint64_t desc_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
  const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
  const uint64_t a0 = desc_fixed | (uint64_t)((smem_base + p.w_bytes_total) >> 4);
  const uint64_t bdesc = desc_fixed | (uint64_t)(smem_base >> 4);
  int s = 0, j = 0;
  uint32_t ph = 0;
  uint64_t adesc = a0;
  TileIter it;
  for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
    const int a = j & ((1 << p.lg_acc) - 1);
    const uint32_t aph = (uint32_t)(j >> p.lg_acc) & 1u;
    mbar_wait(smem_u32(&s_tempty[a]), aph ^ 1u);
    mbar_wait(smem_u32(&s_full[s]), ph);
    tc_fence_after();
    umma_kblock<KSTEPS>(tmem_base + (uint32_t)a * p.acc_stride, adesc, bdesc, p.idesc, 0u);
    umma_commit(smem_u32(&s_empty[s]));
    umma_commit(smem_u32(&s_tfull[a]));
    adesc += stage16;
    if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
  }
}

// bf16x3: the tensor core truncates (does not round) on every accumulation, a bias that grows with the length of
// the accumulation chain (~n_mma * 2^-24).  Keep chains short and small terms separate: the main term a1*w1 goes
// round-robin by tap into accumulators 0..2, all correction terms into accumulator 3; the epilogue adds the four
// with round-to-nearest fp32.  Single-buffered (4 x 128 columns = all of TMEM).
template <int KSTEPS>
__device__ __forceinline__ void mma_role_x3(const TcParams& p, uint32_t smem_base, uint32_t tmem_base, int total_tiles,
                                            uint64_t* s_full, uint64_t* s_empty, uint64_t* s_tfull, uint64_t* s_tempty) {
  const uint64_t desc_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
  const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
  const uint64_t a0 = desc_fixed | (uint64_t)(smem_base >> 4);
  const uint64_t b_off = (uint64_t)(p.a_bytes >> 4);
  int s = 0, j = 0;
  uint32_t ph = 0;
  uint64_t adesc = a0;
  TileIter it;
  for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
    mbar_wait(smem_u32(&s_tempty[0]), ((uint32_t)j & 1u) ^ 1u);
    tc_fence_after();
    uint32_t used = 0;
    if (p.x3_fused) {
      const uint64_t a16 = (uint64_t)(p.a_bytes >> 4), b16 = (uint64_t)(p.b_bytes >> 4);
      const uint64_t w_off = (uint64_t)p.n_limbs_a * a16;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        for (int kc = 0; kc < p.cpk; ++kc) {
          mbar_wait(smem_u32(&s_full[s]), ph);
          tc_fence_after();
          for (int t = 0; t < p.n_terms; ++t) {        // smallest products first, the main term last (as in the unfused order)
            const int acc = (p.la[t] == 0 && p.lw[t] == 0) ? (tap % 3) : 3;
            umma_kblock<KSTEPS>(tmem_base + (uint32_t)acc * p.acc_stride, adesc + (uint64_t)p.la[t] * a16, adesc + w_off + (uint64_t)p.lw[t] * b16,
                                p.idesc, (used >> acc) & 1u);
            used |= 1u << acc;
          }
          umma_commit(smem_u32(&s_empty[s]));
          adesc += stage16;
          if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
        }
      }
      umma_commit(smem_u32(&s_tfull[0]));
      continue;
    }
    for (int tap = 0; tap < p.ntaps; ++tap) {
      for (int t = 0; t < p.n_terms; ++t) {
        const int acc = (p.la[t] == 0 && p.lw[t] == 0) ? (tap % 3) : 3;
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * p.acc_stride;
        for (int kc = 0; kc < p.cpk; ++kc) {
          mbar_wait(smem_u32(&s_full[s]), ph);
          tc_fence_after();
          umma_kblock<KSTEPS>(tmem_d, adesc, adesc + b_off, p.idesc, (used >> acc) & 1u);
          used |= 1u << acc;
          umma_commit(smem_u32(&s_empty[s]));
          adesc += stage16;
          if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
        }
      }
    }
    umma_commit(smem_u32(&s_tfull[0]));
  }
}

// halo mode + limbs: accumulator assignment as mma_role_x3 (main term round-robin by tap over accumulators 0..2, every
// correction term into accumulator 3), operand walk as mma_role_halo
template <int KSTEPS>
__device__ __forceinline__ void mma_role_halo_x3(const TcParams& p, uint32_t smem_base, uint32_t tmem_base, int total_tiles,
                                                 uint64_t* s_full, uint64_t* s_empty, uint64_t* s_tfull, uint64_t* s_tempty,
                                                 uint64_t* s_wbar) {
  mbar_wait(smem_u32(s_wbar), 0);
  const uint64_t a_fixed = ((uint64_t)p.desc_hi_halo << 32) | (1ull << 16);
  const uint64_t b_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
  const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
  const uint64_t a0 = a_fixed | (uint64_t)((smem_base + p.w_bytes_total) >> 4);
  const uint64_t b0 = b_fixed | (uint64_t)(smem_base >> 4);
  const uint64_t wtile16 = (uint64_t)(p.w_tile_bytes >> 4);
  int s = 0, j = 0;
  uint32_t ph = 0;
  uint64_t adesc = a0;
  TileIter it;
  for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
    mbar_wait(smem_u32(&s_tempty[0]), ((uint32_t)j & 1u) ^ 1u);
    tc_fence_after();
    uint32_t used = 0;
    for (int la = 0; la < p.n_limbs_a; ++la) {
      for (int kc = 0; kc < p.cpk; ++kc) {
        mbar_wait(smem_u32(&s_full[s]), ph);
        tc_fence_after();
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int lw = 0; lw < p.n_limbs_w; ++lw) {
            if (!((p.term_mask[la] >> lw) & 1u)) continue;
            const int acc = (la == 0 && lw == 0) ? (tap % 3) : 3;
            const uint64_t bdesc = b0 + wtile16 * (uint64_t)((tap * p.n_limbs_w + lw) * p.cpk + kc);
            umma_kblock<KSTEPS>(tmem_base + (uint32_t)acc * p.acc_stride, adesc + (uint64_t)p.tap_off16[tap], bdesc, p.idesc, (used >> acc) & 1u);
            used |= 1u << acc;
          }
        }
        umma_commit(smem_u32(&s_empty[s]));
        adesc += stage16;
        if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
      }
    }
    umma_commit(smem_u32(&s_tfull[0]));
  }
}

// ------------------------------------------------------------------------------------------ the kernel
// kLean (WIDE / fused-first-layer kernels): plain bf16 layers only -- the bf16x3 accumulator splits, limb outputs,
// conv2d_transpose phase scatter and special epilogues are compiled out, which shortens the per-tile epilogue chain
// these layers are bound by and frees registers for the 576 / 672-thread CTAs.
template <bool kLean, int kProd, bool kPh16 = false>
__device__ __forceinline__ void conv_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p) {
  const int f_x3acc = kLean ? 0 : p.x3acc;
  const int f_out_limbs = kLean ? 1 : p.out_limbs;
  const int f_out_ls = kLean ? 0 : p.out_ls;
  const int f_phase_cout = kProd != 0 ? 0 : p.phase_cout;   // merged conv2d_transpose phases (in the lean body too, not in the producer-warp kernels)
  const int f_epi = kLean ? (int)EPI_NONE : p.epi;
  const bool f_f16 = kLean ? false : (p.f16 != 0);
  const float f_cs = kLean ? 1.0f : p.corr_scale;
  const bool f_im2col = kProd == 1 && p.im2col;   // the im2col producer exists only in conv_tc_kernel_first
  const bool f_diff = kProd == 2 && p.diffp;      // the difference-volume producer only in conv_tc_kernel_diff
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_tfull[kMaxAcc];
  __shared__ __align__(8) uint64_t s_tempty[kMaxAcc];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(8) uint64_t s_wbar;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), f_im2col ? 4u : f_diff ? 7u : 1u);   // one arrival per producer warp
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    for (int a = 0; a < kMaxAcc; ++a) {
      mbar_init(smem_u32(&s_tfull[a]), 1);
      mbar_init(smem_u32(&s_tempty[a]), p.lg_groups >= 0 ? 4u : (uint32_t)p.epi_warps);   // the warps draining one tile
    }
    mbar_init(smem_u32(&s_wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {   // whole warp: TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // programmatic dependent launch: let the next kernel's CTAs be scheduled as SMs free up (they run their own prologue
  // and then block in griddepcontrol.wait until this grid has completed and flushed).  Issued only AFTER this CTA owns
  // its TMEM columns: a dependent CTA that became co-resident earlier could take the free columns and then sit in
  // griddepcontrol.wait while this CTA spins in tcgen05.alloc.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on we touch global memory, so wait for the previous grid (no-op without the PDL launch attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  const int kblocks = p.ntaps * p.kchunks;

  const bool extra_producer = f_im2col && warp >= 2 + p.epi_warps;   // im2col: three more producer warps after the epilogue's
  if (f_im2col && (warp == 0 || extra_producer)) {
    // ===================================================================== im2col producers (4 warps, one pixel per lane)
    const int pq = warp == 0 ? 0 : warp - (1 + p.epi_warps);
    if (pq == 0 && lane == 0) {
      const uint32_t wbar = smem_u32(&s_wbar);
      mbar_expect_tx(wbar, p.b_bytes);
      tma_load_3d(smem_base, &tmB, wbar, 0, 0, 0);
    }
    const int r = pq * 32 + lane;                               // A-tile row = pixel of the tile, ordered [n][h][w]
    const int r_wi = r & ((1 << p.lbw) - 1);
    const int r_hi = (r >> p.lbw) & ((1 << p.lbh) - 1);
    const int r_ni = r >> (p.lbw + p.lbh);
    // 64-byte rows in the 64B-swizzled K-major layout TMA would produce: 16-byte chunk q of row r sits at chunk
    // q ^ ((address >> 7) & 3) = q ^ ((r >> 1) & 3)  (stage bases are 1024-aligned)
    uint8_t* const row_ptr = smem_raw + (smem_base - smem_u32(smem_raw)) + p.w_bytes_total + (uint32_t)r * 64u;
    const int sw = (r >> 1) & 3;
    const int H = p.MH, W = p.MW;
    int s = 0;
    uint32_t ph = 0;
    TileIter it;
    for (it.init(p, total_tiles); it.valid(); it.next(p)) {
      const int x = (it.tw << p.lbw) + r_wi, y = (it.th << p.lbh) + r_hi, n = (it.tn << (7 - p.lbw - p.lbh)) + r_ni;
      float e[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) e[k] = 0.f;
      const uint8_t* src0 = p.img + (((int64_t)n * H + (y - 1)) * W + (x - 1)) * 3;   // window row 0, column x-1
      const bool xl = x > 0, xr = x + 1 < W;
      if (n < p.N && y > 0 && y + 1 < H) {
        // interior rows (all but 2 of H; warp-uniform for row-segment tiles): 27 byte loads, only the window's
        // left / right column predicated -- the fully predicated form below costs ~550 instructions per pixel row and
        // made these four warps the layer's bottleneck (profiles/r01h)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint8_t* src = src0 + (int64_t)ky * W * 3;
#pragma unroll
          for (int jj = 0; jj < 9; ++jj) {
            const bool v = jj < 3 ? xl : (jj >= 6 ? xr : true);
            e[ky * 9 + jj] = v ? (float)__ldg(src + jj) - 128.0f : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1;
          const bool rowv = n < p.N && yy >= 0 && yy < H;
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
      mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);             // the MMAs that read this stage have retired
      uint8_t* dst = row_ptr + (uint32_t)s * p.stage_bytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = c[q];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (f_diff && (warp == 0 || warp >= 2 + p.epi_warps)) {
    // ===================================================================== difference-volume producers (7 warps)
    // halo tile of two 8x8 windows: rows ordered [y][n][x], y, x in 0..9 (ring = zero padding), n in 0..1
    const int pq = warp == 0 ? 0 : warp - (1 + p.epi_warps);
    if (pq == 0 && lane == 0) {
      const uint32_t wbar = smem_u32(&s_wbar);
      mbar_expect_tx(wbar, (uint32_t)(p.ntaps * p.kchunks) * p.b_bytes);
      for (int tap = 0; tap < p.ntaps; ++tap)
        tma_load_3d(smem_base + (uint32_t)tap * p.w_tile_bytes, &tmB, wbar, 0, 0, tap);
    }
    const int r = pq * 32 + lane;
    const bool rvalid = r < 200;
    const int y = r / 20, rem = r - y * 20, wn = rem / 10, x = rem - wn * 10;
    const bool ring = !(y >= 1 && y <= 8 && x >= 1 && x <= 8);
    uint8_t* const row_ptr = smem_raw + (smem_base - smem_u32(smem_raw)) + p.w_bytes_total + (uint32_t)r * 64u;
    const int sw = (r >> 1) & 3;
    const int P = p.dGH * p.dGW;
    int s = 0;
    uint32_t ph = 0;
    TileIter it;
    for (it.init(p, total_tiles); it.valid(); it.next(p)) {
      const int window = (it.tn << 1) + wn;                   // label pixel index over all pairs
      float d[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) d[k] = 0.f;
      if (rvalid && !ring && window < p.N) {
        const int pair = window / P, pix = window - pair * P;
        const int py = pix / p.dGW, px = pix - py * p.dGW;
        const int sy = py + (y - 1) - 4, sx = px + (x - 1) - 4;    // (i, j) = (y-1, x-1); offsets (ox, oy) = (j-4, i-4)
        const bool inb = sy >= 0 && sy < p.dGH && sx >= 0 && sx < p.dGW;
        const float4* cur = reinterpret_cast<const float4*>(p.f_cur + (int64_t)window * 32);
        const float4* prv = reinterpret_cast<const float4*>(p.f_prev + ((int64_t)pair * P + (inb ? sy * p.dGW + sx : 0)) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a = __ldg(cur + q);
          const float4 b = inb ? __ldg(prv + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          d[4 * q + 0] = a.x - b.x; d[4 * q + 1] = a.y - b.y; d[4 * q + 2] = a.z - b.z; d[4 * q + 3] = a.w - b.w;
        }
      }
      uint4 c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        c[q] = make_uint4(pack_bf16(d[8 * q], d[8 * q + 1]), pack_bf16(d[8 * q + 2], d[8 * q + 3]),
                          pack_bf16(d[8 * q + 4], d[8 * q + 5]), pack_bf16(d[8 * q + 6], d[8 * q + 7]));
      mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
      if (rvalid) {
        uint8_t* dst = row_ptr + (uint32_t)s * p.stage_bytes;
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = c[q];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_full[s]));
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 0) {
    // ===================================================================== TMA producer (one lane)
    const bool producer_lane = elect_one(lane);
    if (producer_lane && p.halo) {
      // resident weights: every (tap, chunk) tile once, one barrier
      const uint32_t wbar = smem_u32(&s_wbar);
      mbar_expect_tx(wbar, (uint32_t)(p.ntaps * p.w_chunks) * p.b_bytes);
      for (int tap = 0; tap < p.ntaps; ++tap)
        for (int wc = 0; wc < p.w_chunks; ++wc)     // wc = lw * cpk + kc; channel coordinate lw * w_cin + kc * KC
          tma_load_3d(smem_base + (uint32_t)(tap * p.w_chunks + wc) * p.w_tile_bytes, &tmB, wbar,
                      (wc / p.cpk) * p.w_cin + (wc % p.cpk) * p.KC, 0, tap);
      const uint32_t stage0 = smem_base + p.w_bytes_total;
      int s = 0;
      uint32_t ph = 0;
      TileIter it, pf;
      pf.init(p, total_tiles);
      for (int d = 0; d < p.pf_dist && pf.valid(); ++d, pf.next(p))
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_prefetch_5d(&tmA, kc * p.KC, (pf.tw << 3) - 1, pf.tn << p.lbn, (pf.th << p.lbh) - 1, 0);
      for (it.init(p, total_tiles); it.valid(); it.next(p)) {
        const int w0 = it.tw << 3, h0 = it.th << p.lbh, n0 = it.tn << p.lbn;
        if (p.pf_dist && pf.valid()) {
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_prefetch_5d(&tmA, kc * p.KC, (pf.tw << 3) - 1, pf.tn << p.lbn, (pf.th << p.lbh) - 1, 0);
          pf.next(p);
        }
        for (int ac = 0; ac < p.a_chunks; ++ac) {     // ac = la * cpk + kc; channel coordinate la * in_ls + kc * KC
          mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
          const uint32_t full = smem_u32(&s_full[s]);
          mbar_expect_tx(full, p.a_bytes);
          tma_load_5d(stage0 + (uint32_t)s * p.stage_bytes, &tmA, full, (ac / p.cpk) * p.in_ls + (ac % p.cpk) * p.KC, w0 - 1, n0, h0 - 1, 0);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    } else if (!p.halo && p.par_issue) {
      // ------------------------------------------------------------- warp-parallel issue (small-K-block layers)
      // One elected lane walking the k-blocks spends ~70 dependent instructions (coordinate arithmetic, barrier wait, expect_tx, two
      // TMA instructions) = 550-650 cycles on EACH k-block; for the 8-10 KB k-blocks of the U-Net / tower layers that chain, not the
      // copy engine or the tensor core, was the layer's clock (ncu source view, profiles/r02s: the MMA lane waits for operands, the
      // producer lane never waits for a free stage, the epilogue warps idle on the accumulator barrier half of the time).  Here lane l
      // of a pass owns k-block base + l: up to `par_issue` (<= stages) k-blocks are issued by one pass through the same chain.  Lanes
      // of one pass wait on DISTINCT stages whose previous fills were issued by earlier passes, so no lane's wait depends on
      // another lane of the pass leaving its wait loop.
      const int G = p.par_issue;
      const int kpt = p.ntaps * p.cpk;                    // k-blocks per tile: (tap, chunk); a fused-limb stage holds every limb
      const uint32_t stage0 = smem_base + (p.wres ? p.w_bytes_total : 0u);
      if (p.wres && lane < kpt) {                         // resident weights: lane l fetches the tile of k-block l (kpt <= 32 here)
        const uint32_t wbar = smem_u32(&s_wbar);
        if (lane == 0) mbar_expect_tx(wbar, (uint32_t)kpt * p.b_bytes);
        __syncwarp(__activemask());
        const int wt = lane / p.cpk, wk = lane - wt * p.cpk;
        tma_load_3d(smem_base + (uint32_t)lane * p.w_tile_bytes, &tmB, wbar, wk * p.KC, 0, wt);
      }
      __syncwarp();
      int s0 = 0;
      uint32_t ph0 = 0;
      TileIter it;
      for (it.init(p, total_tiles); it.valid(); it.next(p)) {
        const int nt = it.nt;
        const int w0 = it.tw << p.lbw, h0 = it.th << p.lbh, n0 = it.tn << (7 - p.lbw - p.lbh);
        for (int base = 0; base < kpt; base += G) {
          const int n = kpt - base < G ? kpt - base : G;
          if (lane < n) {
            const int kb = base + lane;
            const int tap = kb / p.cpk, kc = kb - tap * p.cpk;
            int s = s0 + lane;
            uint32_t ph = ph0;
            if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
            const int dy = p.dy[tap], dx = p.dx[tap];
            int cw, cpar, chh, coff;
            if (p.istride == 1) { cw = w0 + dx; cpar = 0; chh = h0 + dy; coff = 0; }
            else { cw = w0 + (dx >> 1); cpar = dy & 1; chh = h0 + (dy >> 1); coff = (dx & 1) * p.in_cs; }
            mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
            const uint32_t full = smem_u32(&s_full[s]);
            const uint32_t dst = stage0 + (uint32_t)s * p.stage_bytes;
            if (p.wres) {
              mbar_expect_tx(full, p.a_bytes);
              tma_load_5d(dst, &tmA, full, coff + kc * p.KC, cw, cpar, chh, n0);
            } else if (p.x3_fused) {
              mbar_expect_tx(full, (uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)p.n_limbs_w * p.b_bytes);
              for (int la = 0; la < p.n_limbs_a; ++la)
                tma_load_5d(dst + (uint32_t)la * p.a_bytes, &tmA, full, coff + la * p.in_ls + kc * p.KC, cw, cpar, chh, n0);
              for (int lw = 0; lw < p.n_limbs_w; ++lw)
                tma_load_3d(dst + (uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)lw * p.b_bytes, &tmB, full, lw * p.w_cin + kc * p.KC, nt * p.BN, tap);
            } else {
              mbar_expect_tx(full, p.a_bytes + p.b_bytes);
              tma_load_5d(dst, &tmA, full, coff + kc * p.KC, cw, cpar, chh, n0);
              tma_load_3d(dst + p.a_bytes, &tmB, full, kc * p.KC, nt * p.BN, tap);
            }
          }
          __syncwarp();
          s0 += n;
          if (s0 >= p.stages) { s0 -= p.stages; ph0 ^= 1u; }
        }
      }
    } else if (producer_lane) {
      int s = 0;
      uint32_t ph = 0;
      TileIter it, pf;
      pf.init(p, total_tiles);
      if (p.wres) {   // resident weights: every (tap, chunk) tile once, one barrier (as in halo mode)
        const uint32_t wbar = smem_u32(&s_wbar);
        mbar_expect_tx(wbar, (uint32_t)(p.ntaps * p.cpk) * p.b_bytes);
        for (int tap = 0; tap < p.ntaps; ++tap)
          for (int kc = 0; kc < p.cpk; ++kc)
            tma_load_3d(smem_base + (uint32_t)(tap * p.cpk + kc) * p.w_tile_bytes, &tmB, wbar, kc * p.KC, 0, tap);
      }
      // prefetch the centre-tap box of future tiles (the other taps overlap it except for a 1-pixel rim)
      auto prefetch_tile = [&](const TileIter& t) {
        if (t.nt != 0) return;
        const int pw = t.tw << p.lbw, phh = t.th << p.lbh, pn = t.tn << (7 - p.lbw - p.lbh);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (p.istride == 1) tma_prefetch_5d(&tmA, kc * p.KC, pw, 0, phh, pn);
          else { tma_prefetch_5d(&tmA, kc * p.KC, pw, 0, phh, pn); tma_prefetch_5d(&tmA, p.in_cs + kc * p.KC, pw, 1, phh, pn); }
        }
      };
      for (int d = 0; d < p.pf_dist && pf.valid(); ++d, pf.next(p)) prefetch_tile(pf);
      for (it.init(p, total_tiles); it.valid(); it.next(p)) {
        const int nt = it.nt;
        const int w0 = it.tw << p.lbw, h0 = it.th << p.lbh, n0 = it.tn << (7 - p.lbw - p.lbh);
        if (p.pf_dist && pf.valid()) { prefetch_tile(pf); pf.next(p); }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dy = p.dy[tap], dx = p.dx[tap];
          int cw, cpar, chh, coff;
          if (p.istride == 1) { cw = w0 + dx; cpar = 0; chh = h0 + dy; coff = 0; }
          else { cw = w0 + (dx >> 1); cpar = dy & 1; chh = h0 + (dy >> 1); coff = (dx & 1) * p.in_cs; }
          if (p.x3_fused) {
            // stage = [A limb 0 | .. | A limb la-1 | W limb 0 | .. | W limb lw-1] of this (tap, chunk)
            for (int kc = 0; kc < p.cpk; ++kc) {
              mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
              const uint32_t full = smem_u32(&s_full[s]);
              const uint32_t dst = smem_base + (uint32_t)s * p.stage_bytes;
              mbar_expect_tx(full, (uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)p.n_limbs_w * p.b_bytes);
              for (int la = 0; la < p.n_limbs_a; ++la)
                tma_load_5d(dst + (uint32_t)la * p.a_bytes, &tmA, full, coff + la * p.in_ls + kc * p.KC, cw, cpar, chh, n0);
              for (int lw = 0; lw < p.n_limbs_w; ++lw)
                tma_load_3d(dst + (uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)lw * p.b_bytes, &tmB, full, lw * p.w_cin + kc * p.KC, nt * p.BN, tap);
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            continue;
          }
          for (int t = 0; t < p.n_terms; ++t) {
            const int a_c0 = coff + p.la[t] * p.in_ls, w_c0 = p.lw[t] * p.w_cin;
            for (int kc = 0; kc < p.cpk; ++kc) {
              mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);
              const uint32_t full = smem_u32(&s_full[s]);
              const uint32_t a_dst = smem_base + (p.wres ? p.w_bytes_total : 0u) + (uint32_t)s * p.stage_bytes;
              mbar_expect_tx(full, p.wres ? p.a_bytes : p.a_bytes + p.b_bytes);
              tma_load_5d(a_dst, &tmA, full, a_c0 + kc * p.KC, cw, cpar, chh, n0);
              if (!p.wres) tma_load_3d(a_dst + p.a_bytes, &tmB, full, w_c0 + kc * p.KC, nt * p.BN, tap);
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one lane)
    if (elect_one(lane)) {
      const int ksteps = p.KC >> 4;
      if (f_im2col) {
        mma_role_im2col<2>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
      } else if (f_x3acc && p.halo) {
        if (ksteps == 4) mma_role_halo_x3<4>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else if (ksteps == 2) mma_role_halo_x3<2>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else mma_role_halo_x3<1>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
      } else if (f_x3acc) {
        if (ksteps == 4) mma_role_x3<4>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty);
        else if (ksteps == 2) mma_role_x3<2>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty);
        else mma_role_x3<1>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty);
      } else if (p.halo) {
        if (ksteps == 4) mma_role_halo<4>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else if (ksteps == 2) mma_role_halo<2>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else mma_role_halo<1>(p, smem_base, tmem_base, total_tiles, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
      } else {
        if (ksteps == 4) mma_role<4>(p, smem_base, tmem_base, total_tiles, kblocks, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else if (ksteps == 2) mma_role<2>(p, smem_base, tmem_base, total_tiles, kblocks, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
        else mma_role<1>(p, smem_base, tmem_base, total_tiles, kblocks, s_full, s_empty, s_tfull, s_tempty, &s_wbar);
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps)
    // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the column blocks.
    const int quad = warp & 3;
    const bool grouped = p.lg_groups >= 0;                       // WIDE: 4-warp groups, each draining whole tiles
    const int grp = grouped ? (warp - 2) >> 2 : 0;
    const int hsel = grouped ? 0 : (warp - 2) >> 2;              // column split: which half of the column blocks
    const bool col_split = !grouped && p.epi_warps == 8;
    const int row = quad * 32 + lane;
    const int bn_shift = p.halo ? p.lbn : (7 - p.lbw - p.lbh);
    const int esize = p.out_bf16 ? 2 : 4;
    // coalesced 32-column path: bf16 output, no special epilogue, 16-byte aligned rows / channel offsets
    const bool fast32 = p.out_bf16 && f_epi == EPI_NONE && ((p.out_cs * 2) & 15) == 0 && ((p.out_coff * 2) & 15) == 0 &&
                        ((f_out_ls * 2) & 15) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    uint8_t* stg = smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_stage_off + (uint32_t)(warp - 2) * 2048u;
    // tile-invariant part of every row's output offset (bytes): computed once per thread.
    // normal mode: rows ordered [n][h][w]; halo mode: [h][n][w] with w = 8
    auto row_off = [&](int r, int& r_ni) -> int64_t {
      const int r_wi = r & ((1 << p.lbw) - 1);
      const int r_hi = p.halo ? (r >> (3 + p.lbn)) : ((r >> p.lbw) & ((1 << p.lbh) - 1));
      r_ni = p.halo ? ((r >> 3) & ((1 << p.lbn) - 1)) : (r >> (p.lbw + p.lbh));
      return (((int64_t)r_ni * p.OH + r_hi * p.omul) * p.OW + r_wi * p.omul) * p.out_cs * esize;
    };
    int my_ni;
    const int64_t my_off = row_off(row, my_ni);
    const int64_t my_off2 = p.split_col ? my_off / ((int64_t)p.out_cs * esize) * p.out2_cs * 2 : 0;   // same pixel, other tensor
    int64_t wr_off[4];      // rows this lane writes in the coalesced pass: quad*32 + i*8 + lane/4
    int wr_ni[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wr_off[i] = row_off(quad * 32 + i * 8 + (lane >> 2), wr_ni[i]);
    const int sw_st = (lane >> 1) & 3;        // staging swizzle of this lane's own row (64-byte rows, 2 per bank row)
    const int cq = lane & 3;
    int j = 0;
    TileIter it;
    for (it.init(p, total_tiles); it.valid(); it.next(p), ++j) {
      if (grouped && (j & ((1 << p.lg_groups) - 1)) != grp) continue;
      const int n0 = it.tn << bn_shift;
      const int64_t tile_off = ((((int64_t)n0 * p.OH + (it.th << p.lbh) * p.omul + p.oaddy) * p.OW +
                                 (it.tw << p.lbw) * p.omul + p.oaddx) * p.out_cs + p.out_coff) * esize;
      uint8_t* const tile_ptr = reinterpret_cast<uint8_t*>(p.out) + tile_off;
      const bool valid = n0 + my_ni < p.N;
      uint8_t* const orow = tile_ptr + my_off;
      uint8_t* orow2 = nullptr;
      if (p.split_col) {
        const int64_t tile_pix = ((int64_t)n0 * p.OH + (it.th << p.lbh) * p.omul + p.oaddy) * p.OW + (it.tw << p.lbw) * p.omul + p.oaddx;
        orow2 = reinterpret_cast<uint8_t*>(p.out2) + tile_pix * p.out2_cs * 2 + my_off2;
      }
      // accumulator stage of tile j = j % stages (WIDE: stages == groups, so group g always drains stage g)
      const int a = f_x3acc ? 0 : (j & ((1 << p.lg_acc) - 1));
      const uint32_t aph = f_x3acc ? ((uint32_t)j & 1u) : ((uint32_t)(j >> p.lg_acc) & 1u);
      mbar_wait(smem_u32(&s_tfull[a]), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (f_x3acc ? 0u : (uint32_t)a * p.acc_stride);
      const int col0 = it.nt * p.BN;
      const int n_main = p.ntaps < 3 ? p.ntaps : 3;     // x3acc: main-term accumulators in use

      if (f_epi == EPI_L2NORM) {
        if (hsel == 0) {
          // feat7: BN == Cout == 32, fp32 output, x * rsqrt(max(sum x^2, 1e-12))  (KFNet/KFNet.py:340)
          uint32_t r[32];
          tmem_ld32(taddr, r);
          float v[32];
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (f_x3acc) {
            for (int acc = 1; acc < 4; ++acc) {
              if (acc < 3 && acc >= n_main) continue;
              tmem_ld32(taddr + (uint32_t)acc * p.acc_stride, r);
              const float sc = acc == 3 ? f_cs : 1.0f;
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), sc, v[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] += __ldg(p.bias + i); ss += v[i] * v[i]; }
          const float inv = rsqrtf(fmaxf(ss, 1e-12f));
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(orow);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(v[4 * q] * inv, v[4 * q + 1] * inv, v[4 * q + 2] * inv, v[4 * q + 3] * inv);
          }
        }
      } else {
        int nc = 0;
        // merged conv2d_transpose: running (phase, channel-in-phase) of the current column block -- a runtime division per block
        // sat on the epilogue's dependent chain (upconv0: four 16-column blocks per tile)
        int ph_c = 0, cc_c = 0;
        if (f_phase_cout && col0) { ph_c = col0 / f_phase_cout; cc_c = col0 - ph_c * f_phase_cout; }
        auto adv_phase = [&](int wdt) {
          if (f_phase_cout) { cc_c += wdt; if (cc_c >= f_phase_cout) { cc_c -= f_phase_cout; ++ph_c; } }
        };
        if (kPh16 && p.ph16) {
          // merged conv2d_transpose with four 16-channel phases (U-Net upconv0, N = 64): both TMEM loads, ONE wait, then per phase
          // bias + ReLU on the conversion + 32 contiguous bytes to pixel (2y + py, 2x + px).  The generic 16-column loop paid a TMEM
          // round trip, a division and a parameter reload per phase: ~2300 cycles x 4 per tile, the layer's clock (profiles/r02s)
          uint32_t ra[32], rb[32];
          tmem_ld32_nowait(taddr, ra);
          tmem_ld32_nowait(taddr + 32, rb);
          tmem_wait_ld2(ra, rb);
          const float4* b4 = reinterpret_cast<const float4*>(p.bias);
          const int64_t px_b = (int64_t)p.out_cs * 2, py_b = (int64_t)p.OW * p.out_cs * 2;
#pragma unroll
          for (int ph = 0; ph < 4; ++ph) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bq = __ldg(b4 + ph * 4 + q);
              const uint32_t* r = ph < 2 ? ra : rb;
              const int o = (ph & 1) * 16 + 4 * q;
              v[4 * q + 0] = __uint_as_float(r[o + 0]) + bq.x; v[4 * q + 1] = __uint_as_float(r[o + 1]) + bq.y;
              v[4 * q + 2] = __uint_as_float(r[o + 2]) + bq.z; v[4 * q + 3] = __uint_as_float(r[o + 3]) + bq.w;
            }
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = p.relu ? pack_bf16_relu(v[2 * i], v[2 * i + 1]) : pack_bf16(v[2 * i], v[2 * i + 1]);
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(orow + (ph >> 1) * py_b + (ph & 1) * px_b);
              dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
              dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
          nc = p.BN;
        }
        if (fast32 && !(kPh16 && p.ph16)) {
          uint8_t* wrow[4];
          bool wvalid[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { wrow[i] = tile_ptr + wr_off[i]; wvalid[i] = n0 + wr_ni[i] < p.N; }
          const int fast_end = p.split_col ? p.split_col : p.Cout;
          for (; nc + 32 <= p.BN && col0 + nc + 32 <= fast_end && (f_phase_cout == 0 || (f_phase_cout & 31) == 0); nc += 32, adv_phase(32)) {
            if (col_split && ((nc >> 5) & 1) != hsel) continue;
            uint32_t r[32];
            tmem_ld32(taddr + nc, r);
            const int c = col0 + nc;
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
            // merged conv2d_transpose: this 32-column block belongs to phase ph -> pixel offset (py, px), channel cc
            int cc = c;
            int64_t ph_off = 0;
            if (f_phase_cout) {
              cc = cc_c;
              ph_off = ((int64_t)(ph_c >> 1) * p.OW + (ph_c & 1)) * p.out_cs * 2;
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (f_x3acc) {   // main splits 1, 2 and the correction accumulator, added with RN fp32
              for (int acc = 1; acc < 4; ++acc) {
                if (acc < 3 && acc >= n_main) continue;
                tmem_ld32(taddr + (uint32_t)acc * p.acc_stride + nc, r);
                const float sc = acc == 3 ? f_cs : 1.0f;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), sc, v[i]);
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bq = __ldg(b4 + q);
              v[4 * q + 0] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
            }
            // single-limb bf16 outputs: the ReLU rides on the conversion (F2FP.RELU)
            const bool relu_cvt = p.relu && f_out_limbs == 1 && !f_f16;
            if (p.relu && !relu_cvt) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            for (int l = 0; l < f_out_limbs; ++l) {
              // stage: row = lane (64 B), 16-byte chunk q at slot q ^ ((lane >> 1) & 3): conflict-free both ways
              if (relu_cvt) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t w0 = pack_bf16_relu(v[8 * q + 0], v[8 * q + 1]), w1 = pack_bf16_relu(v[8 * q + 2], v[8 * q + 3]);
                  const uint32_t w2 = pack_bf16_relu(v[8 * q + 4], v[8 * q + 5]), w3 = pack_bf16_relu(v[8 * q + 6], v[8 * q + 7]);
                  *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw_st) << 4)) = make_uint4(w0, w1, w2, w3);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t w0 = pack_limb(v[8 * q + 0], v[8 * q + 1], f_f16), w1 = pack_limb(v[8 * q + 2], v[8 * q + 3], f_f16);
                  const uint32_t w2 = pack_limb(v[8 * q + 4], v[8 * q + 5], f_f16), w3 = pack_limb(v[8 * q + 6], v[8 * q + 7], f_f16);
                  *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw_st) << 4)) = make_uint4(w0, w1, w2, w3);
                }
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = i * 8 + (lane >> 2);
                const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((cq ^ ((rr >> 1) & 3)) << 4));
                if (wvalid[i]) *reinterpret_cast<uint4*>(wrow[i] + ph_off + ((int64_t)cc + (int64_t)l * f_out_ls) * 2 + cq * 16) = val;
              }
              __syncwarp();
              if (l + 1 < f_out_limbs) {     // next limb = residual after removing this limb's bf16 value
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = limb_residual(v[i], f_f16);
              }
            }
          }
        }
        for (; nc < p.BN; nc += 16, adv_phase(16)) {
          if (col_split && ((nc >> 4) & 1) != hsel) continue;
          float v[16];
          tmem_ld16(taddr + nc, v);
          if (f_x3acc) {
            for (int acc = 1; acc < 4; ++acc) {
              if (acc < 3 && acc >= n_main) continue;
              float w[16];
              tmem_ld16(taddr + (uint32_t)acc * p.acc_stride + nc, w);
              const float sc = acc == 3 ? f_cs : 1.0f;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaf(w[i], sc, v[i]);
            }
          }
          const int c = col0 + nc;
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = __ldg(b4 + q);
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (f_epi == EPI_EXP_CH3 && c == 0) v[3] = expf(v[3]);
          if (!valid) continue;
          const int nvalid = p.Cout - c;
          int cc = c;
          int64_t ph_off = 0;
          if (f_phase_cout) {
            cc = cc_c;
            ph_off = ((int64_t)(ph_c >> 1) * p.OW + (ph_c & 1)) * p.out_cs * esize;
          }
          if (p.out_bf16) {
            uint8_t* rowbase = orow;
            if (p.split_col && c >= p.split_col) { rowbase = orow2; cc = c - p.split_col; }
            for (int l = 0; l < f_out_limbs; ++l) {
              uint8_t* dst = rowbase + ph_off + ((int64_t)cc + (int64_t)l * f_out_ls) * 2;
              if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                uint4 o0 = make_uint4(pack_limb(v[0], v[1], f_f16), pack_limb(v[2], v[3], f_f16), pack_limb(v[4], v[5], f_f16), pack_limb(v[6], v[7], f_f16));
                uint4 o1 = make_uint4(pack_limb(v[8], v[9], f_f16), pack_limb(v[10], v[11], f_f16), pack_limb(v[12], v[13], f_f16), pack_limb(v[14], v[15], f_f16));
                reinterpret_cast<uint4*>(dst)[0] = o0;
                reinterpret_cast<uint4*>(dst)[1] = o1;
              } else {
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) { const uint32_t b = pack_limb(v[i], 0.f, f_f16); d[i] = *reinterpret_cast<const __nv_bfloat16*>(&b); }
              }
              if (l + 1 < f_out_limbs) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = limb_residual(v[i], f_f16);
              }
            }
          } else {
            float* d = reinterpret_cast<float*>(orow + ph_off) + cc;
            if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                reinterpret_cast<float4*>(d)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) d[i] = v[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_tempty[a]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// The N = 256-column layers of SCoordNet (conv2a .. conv6).  Two CTAs of a cluster (neighbouring SMs) execute ONE tcgen05.mma of
// M = 256 x N = 256: each CTA stages only ITS 128 rows of A and ITS 128 of the 256 rows of the weight tile (32 KB per k-block instead
// of 48 KB), the tensor core of each SM reads both halves of B, each CTA's TMEM receives its own 128 rows of D.  What it buys: the
// one-CTA kernel keeps 4 x 48 KB in flight against a consumption of 96 B/clk per SM -- the ncu capture of conv2b shows the MMA issuer
// waiting for operands a quarter of the time, the producer waiting for a free stage 40 % of the time; here 6 x 32 KB are in flight
// against 64 B/clk.  Leader = cluster rank 0: it alone issues the MMAs; every full barrier lives in the leader (both CTAs' TMA loads
// complete on it), tcgen05.commit multicasts the stage-free / accumulator-ready arrivals to both CTAs, the epilogue warps of both
// CTAs release the accumulator on the leader's barrier.
struct PairIter {
  int pt, end, step, nt, tw, th, tn;
  __device__ __forceinline__ void decode(const TcParams& p, int rank) {
    nt = pt % p.n_tiles;
    int mt = (pt / p.n_tiles) * 2 + rank;          // this CTA's 128-row tile of the pair's 256 rows (may lie past the end: all rows invalid)
    tw = mt % p.tiles_w; mt /= p.tiles_w;
    th = mt % p.tiles_h;
    tn = mt / p.tiles_h;
  }
  __device__ __forceinline__ void init(const TcParams& p, int total_pt, int rank) {
    pt = (int)(blockIdx.x >> 1); end = total_pt; step = (int)(gridDim.x >> 1);
    if (pt < end) decode(p, rank);
  }
  __device__ __forceinline__ bool valid() const { return pt < end; }
  __device__ __forceinline__ void next(const TcParams& p, int rank) { pt += step; if (pt < end) decode(p, rank); }
};

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ __align__(8) uint64_t s_empty[kMaxStages];
  __shared__ __align__(8) uint64_t s_tfull[2];
  __shared__ __align__(8) uint64_t s_tempty[2];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&s_tfull[a]), 1); mbar_init(smem_u32(&s_tempty[a]), 2u * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {   // one warp of EACH CTA of the pair: the allocation is collective
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_pt = ((m_tiles + 1) >> 1) * p.n_tiles;
  const uint32_t b_half = p.b_bytes >> 1;

  if (warp == 0) {
    // ===================================================================== TMA producer (one lane, both CTAs)
    if (elect_one(lane)) {
      int s = 0;
      uint32_t ph = 0;
      PairIter it;
      for (it.init(p, total_pt, rank); it.valid(); it.next(p, rank)) {
        const int w0 = it.tw << p.lbw, h0 = it.th << p.lbh, n0 = it.tn << (7 - p.lbw - p.lbh);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dy = p.dy[tap], dx = p.dx[tap];
          int cw, cpar, chh, coff;
          if (p.istride == 1) { cw = w0 + dx; cpar = 0; chh = h0 + dy; coff = 0; }
          else { cw = w0 + (dx >> 1); cpar = dy & 1; chh = h0 + (dy >> 1); coff = (dx & 1) * p.in_cs; }
          for (int kc = 0; kc < p.cpk; ++kc) {
            mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);                       // the pair's MMAs that read this stage (in BOTH CTAs) have retired
            const uint32_t full_leader = mapa_cluster(smem_u32(&s_full[s]), 0u);
            const uint32_t a_dst = smem_base + (uint32_t)s * p.stage_bytes;
            if (p.x3acc) {
              // limb modes: the stage holds every limb of this CTA's A rows and of its half of the weight tile (fused-limb stage)
              if (leader) mbar_expect_tx(smem_u32(&s_full[s]), 2u * ((uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)p.n_limbs_w * b_half));
              for (int la = 0; la < p.n_limbs_a; ++la)
                tma_load_5d_2cta(a_dst + (uint32_t)la * p.a_bytes, &tmA, full_leader, coff + la * p.in_ls + kc * p.KC, cw, cpar, chh, n0);
              for (int lw = 0; lw < p.n_limbs_w; ++lw)
                tma_load_3d_2cta(a_dst + (uint32_t)p.n_limbs_a * p.a_bytes + (uint32_t)lw * b_half, &tmB, full_leader, lw * p.w_cin + kc * p.KC,
                                 it.nt * p.BN + rank * (p.BN >> 1), tap);
            } else {
              if (leader) mbar_expect_tx(smem_u32(&s_full[s]), 2u * (p.a_bytes + b_half));   // both CTAs' loads land on the leader's barrier
              tma_load_5d_2cta(a_dst, &tmA, full_leader, coff + kc * p.KC, cw, cpar, chh, n0);
              tma_load_3d_2cta(a_dst + p.a_bytes, &tmB, full_leader, kc * p.KC, it.nt * p.BN + rank * (p.BN >> 1), tap);
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one lane of the LEADER)
    if (leader && elect_one(lane)) {
      const uint64_t desc_fixed = ((uint64_t)p.desc_hi << 32) | (1ull << 16);
      const uint64_t stage16 = (uint64_t)(p.stage_bytes >> 4);
      const uint64_t a0 = desc_fixed | (uint64_t)(smem_base >> 4);
      const uint64_t b_off = (uint64_t)(p.a_bytes >> 4);
      const int kblocks = p.ntaps * p.cpk;
      int s = 0, j = 0;
      uint32_t ph = 0;
      uint64_t adesc = a0;
      PairIter it;
      if (p.x3acc) {
        // limb modes: four 128-column accumulators (main term round-robin by tap over 0..2, correction terms in 3), single-buffered
        const uint64_t a16 = (uint64_t)(p.a_bytes >> 4), bh16 = (uint64_t)(b_half >> 4);
        const uint64_t w_off = (uint64_t)p.n_limbs_a * a16;
        for (it.init(p, total_pt, rank); it.valid(); it.next(p, rank), ++j) {
          mbar_wait(smem_u32(&s_tempty[0]), ((uint32_t)j & 1u) ^ 1u);
          tc_fence_after();
          uint32_t used = 0;
          for (int tap = 0; tap < p.ntaps; ++tap) {
            for (int kc = 0; kc < p.cpk; ++kc) {
              mbar_wait(smem_u32(&s_full[s]), ph);
              tc_fence_after();
              for (int t = 0; t < p.n_terms; ++t) {
                const int acc = (p.la[t] == 0 && p.lw[t] == 0) ? (tap % 3) : 3;
                umma_kblock_2cta<4>(tmem_base + (uint32_t)acc * p.acc_stride, adesc + (uint64_t)p.la[t] * a16, adesc + w_off + (uint64_t)p.lw[t] * bh16,
                                    p.idesc, (used >> acc) & 1u);
                used |= 1u << acc;
              }
              umma_commit_2cta(smem_u32(&s_empty[s]));
              adesc += stage16;
              if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
            }
          }
          umma_commit_2cta(smem_u32(&s_tfull[0]));
        }
      } else
      for (it.init(p, total_pt, rank); it.valid(); it.next(p, rank), ++j) {
        const int a = j & 1;
        mbar_wait(smem_u32(&s_tempty[a]), (((uint32_t)j >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)a * 256u;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(&s_full[s]), ph);
          tc_fence_after();
          umma_kblock_2cta<4>(tmem_d, adesc, adesc + b_off, p.idesc, kb ? 1u : 0u);
          umma_commit_2cta(smem_u32(&s_empty[s]));
          adesc += stage16;
          if (++s == p.stages) { s = 0; ph ^= 1u; adesc = a0; }
        }
        umma_commit_2cta(smem_u32(&s_tfull[a]));
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps per CTA: own 128 rows, 256 columns)
    const int quad = warp & 3, hsel = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    uint8_t* stg = smem_raw + (smem_base - smem_u32(smem_raw)) + p.epi_stage_off + (uint32_t)(warp - 2) * 2048u;
    auto row_off = [&](int r, int& r_ni) -> int64_t {
      const int r_wi = r & ((1 << p.lbw) - 1);
      const int r_hi = (r >> p.lbw) & ((1 << p.lbh) - 1);
      r_ni = r >> (p.lbw + p.lbh);
      return (((int64_t)r_ni * p.OH + r_hi) * p.OW + r_wi) * p.out_cs * 2;
    };
    int64_t wr_off[4];
    int wr_ni[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wr_off[i] = row_off(quad * 32 + i * 8 + (lane >> 2), wr_ni[i]);
    (void)row;
    const int sw_st = (lane >> 1) & 3, cq = lane & 3;
    const uint32_t tempty_leader0 = mapa_cluster(smem_u32(&s_tempty[0]), 0u), tempty_leader1 = mapa_cluster(smem_u32(&s_tempty[1]), 0u);
    int j = 0;
    PairIter it;
    if (p.x3acc) {
      // limb modes: BN = 128 columns per tile, two warps per lane quadrant split the four 32-column blocks; the accumulators are added with
      // round-to-nearest fp32 (the correction accumulator scaled by corr_scale), then split back into limbs
      const bool f16 = p.f16 != 0;
      const int n_main = p.ntaps < 3 ? p.ntaps : 3;
      for (it.init(p, total_pt, rank); it.valid(); it.next(p, rank), ++j) {
        const int n0 = it.tn << (7 - p.lbw - p.lbh);
        const int64_t tile_off = ((((int64_t)n0 * p.OH + (it.th << p.lbh)) * p.OW + (it.tw << p.lbw)) * p.out_cs + p.out_coff) * 2;
        uint8_t* const tile_ptr = reinterpret_cast<uint8_t*>(p.out) + tile_off;
        uint8_t* wrow[4];
        bool wvalid[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { wrow[i] = tile_ptr + wr_off[i]; wvalid[i] = n0 + wr_ni[i] < p.N; }
        mbar_wait(smem_u32(&s_tfull[0]), (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int col0 = it.nt * p.BN;
        for (int nc = 0; nc < p.BN; nc += 32) {
          if (((nc >> 5) & 1) != hsel) continue;
          uint32_t r[32];
          tmem_ld32(taddr + nc, r);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          for (int acc = 1; acc < 4; ++acc) {
            if (acc < 3 && acc >= n_main) continue;
            tmem_ld32(taddr + (uint32_t)acc * p.acc_stride + nc, r);
            const float sc = acc == 3 ? p.corr_scale : 1.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), sc, v[i]);
          }
          const int c = col0 + nc;
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = __ldg(b4 + q);
            v[4 * q + 0] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          for (int l = 0; l < p.out_limbs; ++l) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw_st) << 4)) =
                  make_uint4(pack_limb(v[8 * q + 0], v[8 * q + 1], f16), pack_limb(v[8 * q + 2], v[8 * q + 3], f16),
                             pack_limb(v[8 * q + 4], v[8 * q + 5], f16), pack_limb(v[8 * q + 6], v[8 * q + 7], f16));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + (lane >> 2);
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((cq ^ ((rr >> 1) & 3)) << 4));
              if (wvalid[i]) *reinterpret_cast<uint4*>(wrow[i] + ((int64_t)c + (int64_t)l * p.out_ls) * 2 + cq * 16) = val;
            }
            __syncwarp();
            if (l + 1 < p.out_limbs) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = limb_residual(v[i], f16);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader0);
      }
    } else
    for (it.init(p, total_pt, rank); it.valid(); it.next(p, rank), ++j) {
      const int a = j & 1;
      const int n0 = it.tn << (7 - p.lbw - p.lbh);
      const int64_t tile_off = ((((int64_t)n0 * p.OH + (it.th << p.lbh)) * p.OW + (it.tw << p.lbw)) * p.out_cs + p.out_coff) * 2;
      uint8_t* const tile_ptr = reinterpret_cast<uint8_t*>(p.out) + tile_off;
      uint8_t* wrow[4];
      bool wvalid[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { wrow[i] = tile_ptr + wr_off[i]; wvalid[i] = n0 + wr_ni[i] < p.N; }
      mbar_wait(smem_u32(&s_tfull[a]), ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)a * 256u;
      const int col0 = it.nt * p.BN;
      for (int nc = 0; nc < 256; nc += 32) {
        if (((nc >> 5) & 1) != hsel) continue;
        uint32_t r[32];
        tmem_ld32(taddr + nc, r);
        const int c = col0 + nc;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = __ldg(b4 + q);
          v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
          v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
        }
        if (p.relu) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw_st) << 4)) =
                make_uint4(pack_bf16_relu(v[8 * q + 0], v[8 * q + 1]), pack_bf16_relu(v[8 * q + 2], v[8 * q + 3]),
                           pack_bf16_relu(v[8 * q + 4], v[8 * q + 5]), pack_bf16_relu(v[8 * q + 6], v[8 * q + 7]));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw_st) << 4)) =
                make_uint4(pack_bf16(v[8 * q + 0], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                           pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2);
          const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((cq ^ ((rr >> 1) & 3)) << 4));
          if (wvalid[i]) *reinterpret_cast<uint4*>(wrow[i] + (int64_t)c * 2 + cq * 16) = val;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a ? tempty_leader1 : tempty_leader0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while its peer may still read its shared memory / signal it
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// One CTA per SM for the big compute-bound layers; two co-resident CTAs per SM (half the smem budget, TMEM by need)
// for the small latency-bound ones, so one CTA's barrier / TMA / TMEM latencies are covered by the other's work.
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<false, 0>(tmA, tmB, p);
}
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_x2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<false, 0>(tmA, tmB, p);
}
// co-resident CTAs for plain bf16 layers (most of the U-Net and the tower): the lean body
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_x2_lean(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<true, 0>(tmA, tmB, p);
}
// the same with the one-wait epilogue of the 4 x 16-channel merged deconv (U-Net upconv0) compiled in: its 64 live accumulator
// registers would cost the other lean kernels spills
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel_x2_lean_ph(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<true, 0, true>(tmA, tmB, p);
}
// WIDE: 16 epilogue warps (four groups draining four tiles concurrently) for one-CTA layers with short tiles.
__global__ void __launch_bounds__(kThreadsWide, 1)
conv_tc_kernel_wide(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<true, 0>(tmA, tmB, p);
}

// Fused first layer: WIDE (or the 320-thread layout) plus three more im2col producer warps.
__global__ void __launch_bounds__(kThreadsWide + 96, 1)
conv_tc_kernel_first(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<true, 1>(tmA, tmB, p);
}

// U-Net conv0 with the difference volume formed by its own producer warps: WIDE plus six more producer warps.
__global__ void __launch_bounds__(kThreadsWide + 192, 1)
conv_tc_kernel_diff(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  conv_tc_body<true, 2>(tmA, tmB, p);
}

// ------------------------------------------------------------------------------------------ host side
int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }
int pow2_divisor(int v, int cap) { int d = 1; while (d * 2 <= cap && v % (d * 2) == 0) d *= 2; return d; }

}  // namespace

bool tc_halo_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_HALO"); return !(e && e[0] == '0'); }();
  return on;
}
bool tc_two_ctas_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_TWO_CTAS"); return !(e && e[0] == '0'); }();
  return on;
}
bool tc_three_ctas_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_THREE_CTAS"); return !(e && e[0] == '0'); }();
  return on;
}
// kernel-variant switches for A/B measurements: NAME=0 disables, NAME=1 enables, unset = default
bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? e[0] != '0' : dflt;
}
bool tc_wide_enabled() {
  static const bool on = env_flag("KFN_TC_WIDE", true);
  return on;
}
bool tc_group2_enabled() {
  static const bool on = env_flag("KFN_TC_GROUP2", true);
  return on;
}
bool tc_lean_x2_enabled() {
  static const bool on = env_flag("KFN_TC_LEAN_X2", true);
  return on;
}
bool tc_diff_fused_enabled() {   // opt-in until validated on hardware
  static const bool on = env_flag("KFN_TC_DIFF_FUSED", false);
  return on;
}
bool tc_im2col_fused_enabled() {
  static const bool on = env_flag("KFN_TC_IM2COL_FUSED", true);
  return on;
}
bool tc_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
bool tc_prefetch_enabled() {
  static const bool on = [] { const char* e = getenv("KFN_TC_PREFETCH"); return e && e[0] == '1'; }();
  return on;
}
int tc_pick_kc(int Cin) { return (Cin % 64 == 0) ? 64 : (Cin % 32 == 0) ? 32 : 16; }
int tc_cout_padded(int Cout) { return (Cout + 15) / 16 * 16; }

int prepare_conv_tc(const ConvGeom& g, const void* in, const void* w_packed, int CoutP, const float* bias_padded,
                    void* out, int out_bf16, int num_sms, TcConvPlan* plan, void* out2) {
  PFN_encodeTiled encode = get_encode();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)"); return KFN_ERR_CUDA; }
  KFN_REQUIRE(g.Cin % 16 == 0, "tensor-core conv needs Cin % 16 == 0");
  KFN_REQUIRE(g.in_cs % 8 == 0, "input channel stride must keep pixels 16-byte aligned");
  KFN_REQUIRE(CoutP % 16 == 0 && (CoutP <= 128 || CoutP % 128 == 0), "CoutP must be 16..128 (mult of 16) or k*128");
  KFN_REQUIRE(g.istride == 1 || (g.istride == 2 && g.IH % 2 == 0 && g.IW % 2 == 0), "stride must be 1 or 2 (even input)");
  KFN_REQUIRE((g.im2col_u8 || (reinterpret_cast<uintptr_t>(in) & 15) == 0) && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
              "TMA needs 16-byte aligned global addresses");
  if (g.diff_cur)
    KFN_REQUIRE(g.Cin == 32 && g.ntaps == 9 && g.istride == 1 && g.omul == 1 && g.limbs_w <= 1 && g.limbs_a <= 1 && CoutP <= 128 &&
                g.IH == 8 && g.IW == 8 && g.MH == 8 && g.MW == 8 && g.dGH > 0 && g.dGW > 0 && !g.im2col_u8 &&
                (reinterpret_cast<uintptr_t>(g.diff_cur) & 15) == 0, "fused difference volume: conv0 on 8x8 windows");
  if (g.im2col_u8)
    KFN_REQUIRE(g.Cin == 32 && g.ntaps == 1 && g.istride == 1 && g.omul == 1 && g.limbs_w <= 1 && g.limbs_a <= 1 &&
                CoutP <= 128 && g.IH == g.MH && g.IW == g.MW, "fused first layer shape");
  static_assert(sizeof(TcParams) + 2 * sizeof(CUtensorMap) <= sizeof(TcConvPlan::blob), "TcConvPlan blob too small");

  CUtensorMap* tmA = reinterpret_cast<CUtensorMap*>(plan->blob);
  CUtensorMap* tmB = tmA + 1;
  TcParams& p = *reinterpret_cast<TcParams*>(tmB + 1);
  memset(plan->blob, 0, sizeof(plan->blob));

  const int KC = tc_pick_kc(g.Cin);
  const bool x3 = g.limbs_w > 1;
  const int bn_cap = x3 ? 128 : 256;                    // bf16x3: four 128-column accumulators fill TMEM
  const int BN = CoutP < bn_cap ? CoutP : bn_cap;
  const int row_bytes = KC * 2;                       // == swizzle span
  const uint32_t layout_type = (KC == 64) ? 2u : (KC == 32) ? 4u : 6u;   // SWIZZLE_128B / 64B / 32B
  const CUtensorMapSwizzle swz = (KC == 64) ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (KC == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const cuuint64_t cs = (cuuint64_t)g.in_cs * 2;      // bytes per pixel

  // ---- halo mode eligibility: stride-1 3x3, 8-wide tiles, one N tile, weights resident in smem
  bool halo = tc_halo_enabled() && g.istride == 1 && g.omul == 1 && g.ntaps >= 2 && CoutP <= bn_cap && g.MW % 8 == 0 &&
              g.IH == g.MH && g.IW == g.MW && g.epi != EPI_L2NORM;
  for (int t = 0; t < g.ntaps; ++t) halo = halo && g.dy[t] >= -1 && g.dy[t] <= 1 && g.dx[t] >= -1 && g.dx[t] <= 1;
  int hbh = 0, hbn = 0;
  if (halo) {
    if (g.MH % 16 == 0) { hbh = 16; hbn = 1; }
    else if (g.MH % 8 == 0) { hbh = 8; hbn = 2; }
    else halo = false;
  }
  const uint32_t w_tile_bytes = ((uint32_t)(BN * row_bytes) + 1023u) & ~1023u;
  const int lw_n = g.limbs_w > 1 ? g.limbs_w : 1;
  const uint32_t w_total = (uint32_t)(g.ntaps * (g.Cin / KC) * lw_n) * w_tile_bytes;   // every (tap, W limb, chunk) tile resident
  if (halo && w_total > (x3 ? 150u : 100u) * 1024u) halo = false;   // (two-CTA plans re-check against their smaller budget below)
  static const bool halo_x3_on = env_flag("KFN_TC_HALO_X3", true);   // A/B switch: halo tiles for the limb modes
  if (x3 && !halo_x3_on) halo = false;

  p.N = g.N; p.MH = g.MH; p.MW = g.MW;
  p.istride = g.istride; p.in_cs = g.in_cs;
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) { p.dy[t] = g.dy[t]; p.dx[t] = g.dx[t]; }
  // limb terms (bf16x3 error-compensated mode): which A limb multiplies which W limb, largest products last
  const int limbs_a = g.limbs_a > 1 ? g.limbs_a : 1, limbs_w = g.limbs_w > 1 ? g.limbs_w : 1;
  p.n_terms = 0;
  auto term = [&](int la_, int lw_) { p.la[p.n_terms] = la_; p.lw[p.n_terms] = lw_; ++p.n_terms; };
  if (limbs_a == 3 && limbs_w == 3) { term(0, 2); term(2, 0); term(1, 1); term(0, 1); term(1, 0); term(0, 0); }
  else if (limbs_a == 1 && limbs_w == 3) { term(0, 2); term(0, 1); term(0, 0); }
  else if (limbs_a == 2 && limbs_w == 2) { term(0, 1); term(1, 0); term(0, 0); }     // fp16x2: a1*w2' + a2'*w1 (x 2^-11) + a1*w1
  else if (limbs_a == 1 && limbs_w == 2) { term(0, 1); term(0, 0); }
  else if (limbs_a == 1 && limbs_w == 1) { term(0, 0); }
  else { set_error("unsupported limb combination"); return KFN_ERR_ARG; }
  p.cpk = g.Cin / KC;
  p.kchunks = p.n_terms * p.cpk; p.KC = KC; p.BN = BN;
  p.in_ls = limbs_a > 1 ? g.in_ls : 0;
  p.w_cin = g.Cin;
  p.out_limbs = g.limbs_out > 1 ? g.limbs_out : 1;
  p.out_ls = p.out_limbs > 1 ? g.out_ls : 0;
  p.n_limbs_a = limbs_a; p.n_limbs_w = limbs_w;
  p.a_chunks = limbs_a * p.cpk; p.w_chunks = limbs_w * p.cpk;
  for (int t = 0; t < p.n_terms; ++t) p.term_mask[p.la[t]] |= 1u << p.lw[t];
  p.x3acc = x3 ? 1 : 0;
  p.phase_cout = g.phase_cout;
  p.split_col = out2 ? g.split_col : 0; p.out2_cs = g.out2_cs; p.out2 = out2;
  if (out2) KFN_REQUIRE(g.split_col % 32 == 0 && g.split_col > 0 && g.split_col < g.Cout && out_bf16 && g.limbs_out <= 1 && g.phase_cout == 0,
                        "dual-output layer shape");
  p.n_tiles = CoutP / BN;
  const uint32_t sbo16 = (uint32_t)(8 * row_bytes) >> 4;
  p.desc_hi = sbo16 | (1u << 14) | (layout_type << 29);   // bits 32-45 SBO, 46 version=1, 61-63 layout
  p.desc_hi_halo = ((uint32_t)(10 * row_bytes) >> 4) | (1u << 14) | (layout_type << 29);
  // instruction descriptor: D = f32, A / B format 1 = bf16 (0 = fp16: KFN_MODE_FP16X2), N >> 3, M >> 4
  const uint32_t ab_fmt = g.half_limbs ? 0u : 1u;
  p.idesc = (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  p.f16 = g.half_limbs ? 1 : 0;
  p.corr_scale = g.half_limbs ? 1.0f / 2048.0f : 1.0f;
  if (g.half_limbs) KFN_REQUIRE(limbs_w == 2 && limbs_a <= 2 && (g.limbs_out <= 2), "fp16 limbs: two-limb weights / activations only");
  p.b_bytes = (uint32_t)(BN * row_bytes);
  p.row_bytes = row_bytes;
  p.OH = g.OH; p.OW = g.OW; p.out_cs = g.out_cs; p.out_coff = g.out_coff; p.Cout = g.Cout;
  p.omul = g.omul; p.oaddy = g.oaddy; p.oaddx = g.oaddx; p.relu = g.relu; p.epi = g.epi; p.out_bf16 = out_bf16;
  p.out = out; p.bias = bias_padded;
  if (g.epi == EPI_L2NORM) KFN_REQUIRE(BN == 32 && g.Cout == 32 && !out_bf16 && g.out_cs % 4 == 0, "l2norm epilogue shape");

  // occupancy: 2 CTAs/SM when the accumulators need <= 256 TMEM columns and an 88 KB smem budget still gives the
  // pipeline its full depth (or the layer is small anyway); big layers keep the whole SM
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)BN) tmem_cols <<= 1;
  if (x3) {   // four accumulators (3 main-term splits + corrections) of pow2(BN) columns each, single-buffered
    uint32_t per = 32;
    while (per < (uint32_t)BN) per <<= 1;
    tmem_cols = 4 * per;
    p.acc_stride = per;
  } else {
    p.acc_stride = tmem_cols / 2;
  }
  p.tmem_cols = tmem_cols;
  const uint32_t stage_guess = ((uint32_t)(kBlockM + BN) * row_bytes + 1023u) & ~1023u;
  const bool streaming = g.ntaps == 1 && g.Cin / KC == 1 && (int64_t)g.N * g.MH * g.MW > (1 << 20);   // K=KC 1-tap GEMMs over big maps
  const bool two_ctas = tc_two_ctas_enabled() && !streaming && !g.im2col_u8 && !g.diff_cur && tmem_cols <= 256 && (12u * stage_guess <= 88u * 1024u || BN <= 64);
  // three CTAs/SM (192 threads, 4 epilogue warps, 56 KB pipeline) for the narrowest layers
  const bool three_ctas = two_ctas && tc_three_ctas_enabled() && tmem_cols <= 128 && 6u * stage_guess <= 56u * 1024u;
  plan->ctas_per_sm = three_ctas ? 3 : two_ctas ? 2 : 1;
  p.epi_warps = three_ctas ? 4 : 8;
  uint32_t budget = three_ctas ? 56 * 1024 : two_ctas ? 88 * 1024 : 200 * 1024;   // n x (budget + staging + 1 KB + static) fits the 228 KB SM
  const uint32_t halo_stage = halo ? (((uint32_t)(10 * hbn * (hbh + 2) * row_bytes) + 1023u) & ~1023u) : 0u;
  if (halo && w_total + 3 * halo_stage > budget) {   // a co-resident budget cannot hold resident weights + pipeline
    plan->ctas_per_sm = 1; p.epi_warps = 8; budget = 200 * 1024;
  }
  // WIDE: a one-CTA layer whose tile carries so few MMA cycles that the ~3000-cycle epilogue chain of one tile would
  // bound it (tensor time per MMA ~ max(BN/2, 48) cycles: N <= 64 is bound by the A-operand fetch from smem).
  // Four accumulator stages of roundup(BN, 32) columns, four epilogue groups, 32 KB of store staging.
  const int64_t mma_cycles = (int64_t)g.ntaps * (g.Cin / KC) * (KC / 16) * (BN / 2 > 48 ? BN / 2 : 48);
  const bool lean_ok = !x3 && !g.half_limbs && g.limbs_out <= 1 && g.epi == EPI_NONE;   // what conv_tc_body<true> keeps
  if (g.im2col_u8 || g.diff_cur) KFN_REQUIRE(g.phase_cout == 0, "producer-warp kernels: no conv2d_transpose phases");
  if (g.im2col_u8) KFN_REQUIRE(lean_ok && plan->ctas_per_sm == 1, "fused first layer: plain bf16 layer, one CTA per SM");
  if (g.diff_cur) KFN_REQUIRE(lean_ok && plan->ctas_per_sm == 1 && halo && hbn == 2 && KC == 32, "fused difference volume: halo-tile plan");
  const bool wide = tc_wide_enabled() && lean_ok && g.phase_cout == 0 && plan->ctas_per_sm == 1 && BN <= 128 && mma_cycles <= 2500 &&
                    (int64_t)g.N * g.MH * g.MW >= (int64_t)kBlockM * 8 * num_sms;
  p.lg_groups = -1; p.lg_acc = 1;
  plan->threads = plan->ctas_per_sm == 3 ? 192 : kThreads;
  plan->lean = (lean_ok && tc_lean_x2_enabled()) ? 1 : 0;
  static const bool ph16_on = env_flag("KFN_TC_PH16", true);   // A/B switch: one-wait epilogue of the 4 x 16-channel deconv
  p.ph16 = (ph16_on && plan->lean && plan->ctas_per_sm >= 2 && !wide && g.phase_cout == 16 && CoutP == 64 && g.Cout == 64 && out_bf16 &&
            !out2 && ((g.out_cs * 2) & 15) == 0 && ((g.out_coff * 2) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  if (wide) {
    p.lg_groups = 2; p.lg_acc = 2; p.epi_warps = 16;
    p.acc_stride = (uint32_t)(BN + 31) & ~31u;
    p.tmem_cols = 32;
    while (p.tmem_cols < 4u * p.acc_stride) p.tmem_cols <<= 1;
    budget = 184 * 1024;
    plan->threads = kThreadsWide;
  } else if (tc_group2_enabled() && !x3 && plan->ctas_per_sm == 2 && g.epi != EPI_L2NORM) {
    // co-resident pairs: the two 4-warp halves of the epilogue alternate tiles (each draining its own accumulator
    // stage whole) instead of splitting one tile's column blocks -- the per-tile fixed cost is paid once per two tiles
    // and layers with a single 32-column block keep all eight warps busy
    p.lg_groups = 1;
  }
  cuuint64_t dimA[5], strA[4];
  cuuint32_t boxA[5];
  if (halo) {
    p.halo = 1;
    p.lbw = 3; p.lbh = ilog2(hbh); p.lbn = ilog2(hbn);
    p.tiles_w = g.MW / 8; p.tiles_h = g.MH / hbh; p.tiles_n = (g.N + hbn - 1) / hbn;
    p.a_bytes = (uint32_t)(10 * hbn * (hbh + 2) * row_bytes);
    p.stage_bytes = (p.a_bytes + 1023u) & ~1023u;
    p.w_tile_bytes = w_tile_bytes; p.w_bytes_total = w_total;
    for (int t = 0; t < g.ntaps; ++t)
      p.tap_off16[t] = (uint32_t)(((1 + g.dy[t]) * 10 * hbn + (1 + g.dx[t])) * row_bytes) >> 4;
    int stages = (int)((budget - w_total) / p.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    KFN_REQUIRE(stages >= 2, "halo pipeline needs at least 2 stages");
    p.stages = stages;
    p.epi_stage_off = w_total + (uint32_t)stages * p.stage_bytes;
    plan->smem = (size_t)p.epi_stage_off + (size_t)p.epi_warps * 2048 + 1024;
    if (g.diff_cur) {   // rows come from the kernel's own producer warps, not from TMA
      p.diffp = 1; p.f_prev = reinterpret_cast<const float*>(in); p.f_cur = g.diff_cur; p.dGH = g.dGH; p.dGW = g.dGW;
      plan->threads += 192;
    }
    // halo view: dims (c, w, n, h, 1) so that the box lands in smem as [y][n][x] rows
    dimA[0] = (cuuint64_t)(limbs_a - 1) * p.in_ls + g.Cin; dimA[1] = g.IW; dimA[2] = g.N; dimA[3] = g.IH; dimA[4] = 1;
    strA[0] = cs; strA[1] = cs * g.IW * g.IH; strA[2] = cs * g.IW; strA[3] = cs * g.IW * g.IH * g.N;
    boxA[0] = (cuuint32_t)KC; boxA[1] = 10; boxA[2] = (cuuint32_t)hbn; boxA[3] = (cuuint32_t)(hbh + 2); boxA[4] = 1;
  } else {
    const int bw = pow2_divisor(g.MW, 128);
    const int bh = pow2_divisor(g.MH, 128 / bw);
    const int bn = 128 / (bw * bh);
    p.lbw = ilog2(bw); p.lbh = ilog2(bh);
    p.tiles_w = g.MW / bw; p.tiles_h = g.MH / bh; p.tiles_n = (g.N + bn - 1) / bn;
    p.a_bytes = (uint32_t)(kBlockM * row_bytes);
    p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
    static const bool x3_fused_on = env_flag("KFN_TC_X3_FUSED", true);   // A/B switch: all limbs of a (tap, chunk) in one stage
    const uint32_t fused_stage = ((uint32_t)limbs_a * p.a_bytes + (uint32_t)limbs_w * p.b_bytes + 1023u) & ~1023u;
    if (x3 && x3_fused_on && !g.im2col_u8 && 2u * fused_stage <= budget) {
      p.x3_fused = 1;
      p.stage_bytes = fused_stage;
    }
    // CTA-pair plan (conv_tc_kernel_pair): plain bf16 layers with a 256-column tile and 64-channel chunks, on their own SM
    static const bool pair_on = env_flag("KFN_TC_PAIR", true);    // KFN_TC_PAIR=0: one-CTA tiles (A/B measurements)
    static const bool pair_x3_on = env_flag("KFN_TC_PAIR_X3", true);   // the limb modes' pair plan (A/B switch)
    const bool pair_common = pair_on && KC == 64 && plan->ctas_per_sm == 1 && !wide && !g.im2col_u8 && !g.diff_cur && out_bf16 && g.phase_cout == 0 &&
                             g.epi == EPI_NONE && !out2 && g.omul == 1 && (num_sms & 1) == 0 && ((g.out_cs * 2) & 15) == 0 &&
                             ((g.out_coff * 2) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                             (int64_t)p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles >= 2 * (int64_t)num_sms;
    const uint32_t pair_x3_stage = (uint32_t)limbs_a * p.a_bytes + (uint32_t)limbs_w * (p.b_bytes >> 1);
    const bool pair = pair_common && !x3 && !g.half_limbs && BN == 256 && g.limbs_out <= 1;
    const bool pair_x3 = pair_common && pair_x3_on && x3 && BN == 128 && ((p.out_ls * 2) & 15) == 0 && 2u * pair_x3_stage <= budget;
    if (pair || pair_x3) {
      p.pair = 1;
      p.x3_fused = 0;
      // 128 rows of A + this CTA's half of the weight tile (every limb of both in the limb modes)
      p.stage_bytes = pair_x3 ? pair_x3_stage : p.a_bytes + (p.b_bytes >> 1);
      const uint32_t ab_fmt2 = g.half_limbs ? 0u : 1u;
      p.idesc = (1u << 4) | (ab_fmt2 << 7) | (ab_fmt2 << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
    }
    uint32_t resident = 0;
    if (g.im2col_u8) {   // resident weight tile at the base, stages hold only the A tile the producer warps write
      p.im2col = 1; p.img = reinterpret_cast<const uint8_t*>(in);
      p.stage_bytes = p.a_bytes;
      p.w_tile_bytes = w_tile_bytes; p.w_bytes_total = resident = w_tile_bytes;
      plan->threads += 96;
    }
    // resident weights for the small non-halo layers (plain bf16, one N tile): a third to a fifth of a tile's TMA lines are weight rows
    // that every tile re-fetches; with at most a third of the pipeline budget they stay at the smem base for the kernel's lifetime
    static const bool wres_on = env_flag("KFN_TC_WRES", true);
    if (wres_on && !g.im2col_u8 && !g.diff_cur && !p.pair && !x3 && !g.half_limbs && p.n_terms == 1 && p.n_tiles == 1 &&
        g.ntaps * p.cpk <= 32 && 3u * w_total <= budget) {
      p.wres = 1;
      p.stage_bytes = (p.a_bytes + 1023u) & ~1023u;
      p.w_tile_bytes = w_tile_bytes; p.w_bytes_total = resident = w_total;
    }
    int stages = (int)((budget - resident) / p.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    KFN_REQUIRE(stages >= 2, "pipeline needs at least 2 stages");
    p.stages = stages;
    p.epi_stage_off = resident + (uint32_t)p.stages * p.stage_bytes;
    plan->smem = (size_t)p.epi_stage_off + (size_t)p.epi_warps * 2048 + 1024;
    boxA[0] = (cuuint32_t)KC; boxA[1] = (cuuint32_t)bw; boxA[2] = 1u; boxA[3] = (cuuint32_t)bh; boxA[4] = (cuuint32_t)bn;
    if (g.istride == 1) {
      dimA[0] = (cuuint64_t)(limbs_a - 1) * p.in_ls + g.Cin; dimA[1] = g.IW; dimA[2] = 1; dimA[3] = g.IH; dimA[4] = g.N;
      strA[0] = cs; strA[1] = cs * g.IW; strA[2] = cs * g.IW; strA[3] = cs * g.IW * g.IH;
    } else {
      dimA[0] = (cuuint64_t)g.in_cs + (cuuint64_t)(limbs_a - 1) * p.in_ls + g.Cin; dimA[1] = g.IW / 2; dimA[2] = 2; dimA[3] = g.IH / 2; dimA[4] = g.N;
      strA[0] = cs * 2; strA[1] = cs * g.IW; strA[2] = cs * g.IW * 2; strA[3] = cs * g.IW * g.IH;
    }
  }
  CUresult r = CUDA_SUCCESS;
  if (!g.im2col_u8 && !g.diff_cur)   // (producer-built A tiles have no activation tensor map: tmA stays a copy of tmB, never used)
    r = encode(tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), dimA, strA, boxA, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r)); return KFN_ERR_CUDA; }
  // ---- B: packed weights [tap][CoutP][Cin]
  cuuint64_t dimB[3] = {(cuuint64_t)g.Cin * limbs_w, (cuuint64_t)CoutP, (cuuint64_t)g.ntaps};
  cuuint64_t strB[2] = {(cuuint64_t)g.Cin * limbs_w * 2, (cuuint64_t)g.Cin * limbs_w * 2 * CoutP};
  cuuint32_t boxB[3] = {(cuuint32_t)KC, (cuuint32_t)(p.pair ? BN / 2 : BN), 1u};
  r = encode(tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dimB, strB, boxB, ones,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r)); return KFN_ERR_CUDA; }
  if (g.im2col_u8 || g.diff_cur) memcpy(tmA, tmB, sizeof(CUtensorMap));

  {
    // prefetch only for memory/latency-bound shapes (few k-blocks per tile); ~384 KB ahead per SM
    const int kblocks = g.ntaps * p.kchunks;
    const uint32_t bytes_per_tile = p.a_bytes * (uint32_t)p.kchunks;
    int dist = (int)((384u * 1024u) / (bytes_per_tile ? bytes_per_tile : 1u));
    if (dist > 64) dist = 64;
    if (dist < 2) dist = 2;
    p.pf_dist = (tc_prefetch_enabled() && (halo || kblocks <= 18)) ? dist : 0;
  }
  {
    // warp-parallel TMA issue for layers whose k-blocks are small and many per tile (see the producer): plain bf16 and fused-limb stages.
    // Measured per layer (profiles/r02t): 8-9 lanes per pass: feat2 -32 %, U-Net conv1b -20 %, conv1a / conv2a / feat4 -11..13 %;
    // 3-4 lanes per pass (24 KB stages, 3-4 of them): conv4 +23 %, conv2b +14 %, conv5 +6 % -- the pass costs about two serial k-blocks
    static const bool par_on = env_flag("KFN_TC_PAR_ISSUE", true);
    const int kpt = g.ntaps * p.cpk;
    int G = p.stages < 32 ? p.stages : 32;
    if (kpt < G) G = kpt;
    p.par_issue = (par_on && !halo && !g.im2col_u8 && !g.diff_cur && !p.pair && p.pf_dist == 0 && (p.n_terms == 1 || p.x3_fused) && G >= 8) ? G : 0;
  }
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  const int max_ctas = num_sms * plan->ctas_per_sm;
  {
    // big one-CTA layers only: >= 2000 tensor cycles per tile amortise the per-tile index decode
    static const bool strided_on = env_flag("KFN_TC_STRIDED", true);
    const int64_t tile_cycles = (int64_t)g.ntaps * (g.Cin / KC) * (KC / 16) * (BN / 2 > 48 ? BN / 2 : 48) * (p.n_terms > 0 ? p.n_terms : 1);
    p.strided = (strided_on && plan->ctas_per_sm == 1 && !g.im2col_u8 && !g.diff_cur && tile_cycles >= 2000 && total_tiles > max_ctas) ? 1 : 0;
  }
  plan->grid = total_tiles < max_ctas ? total_tiles : max_ctas;
  plan->pair = p.pair;
  if (p.pair) {
    const int pair_tiles = ((p.tiles_w * p.tiles_h * p.tiles_n + 1) / 2) * p.n_tiles;
    const int pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
    plan->grid = 2 * pairs;
    plan->threads = kThreads;
    p.strided = 0;
  }
  return KFN_OK;
}

int launch_conv_tc(const TcConvPlan& plan, cudaStream_t st) {
  static bool attr_set_dev[64] = {false};
  int dev_ = 0;
  KFN_CUDA_OK(cudaGetDevice(&dev_));
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    const int max_smem = 201 * 1024 + 1024 + (int)kEpiStageBytes;
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, 106 * 1024));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_x2_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 106 * 1024));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_x2_lean_ph, cudaFuncAttributeMaxDynamicSharedMemorySize, 106 * 1024));
    // WIDE: 184 KB pipeline + 32 KB staging + 1 KB alignment; first layer in the 320-thread layout: 200 + 16 + 1 KB
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 219 * 1024));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_first, cudaFuncAttributeMaxDynamicSharedMemorySize, 219 * 1024));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, 219 * 1024));
    KFN_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  const CUtensorMap* tmA = reinterpret_cast<const CUtensorMap*>(plan.blob);
  const TcParams& p = *reinterpret_cast<const TcParams*>(tmA + 2);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)plan.grid);
  cfg.blockDim = dim3((unsigned)plan.threads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  if (plan.pair) {
    cudaLaunchAttribute attr2[2];
    attr2[0].id = cudaLaunchAttributeClusterDimension;
    attr2[0].val.clusterDim.x = 2; attr2[0].val.clusterDim.y = 1; attr2[0].val.clusterDim.z = 1;
    attr2[1] = attr[0];
    cfg.attrs = attr2;
    cfg.numAttrs = tc_pdl_enabled() ? 2 : 1;
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_pair, tmA[0], tmA[1], p));
  } else if (p.diffp)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_diff, tmA[0], tmA[1], p));
  else if (p.im2col)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_first, tmA[0], tmA[1], p));
  else if (plan.threads == kThreadsWide)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_wide, tmA[0], tmA[1], p));
  else if (plan.ctas_per_sm >= 2 && plan.lean && p.ph16)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_x2_lean_ph, tmA[0], tmA[1], p));
  else if (plan.ctas_per_sm >= 2 && plan.lean)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_x2_lean, tmA[0], tmA[1], p));
  else if (plan.ctas_per_sm >= 2)
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel_x2, tmA[0], tmA[1], p));
  else
    KFN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel, tmA[0], tmA[1], p));
  KFN_LAUNCH_OK();
  return KFN_OK;
}

}  // namespace kfn
