// Hardware probe: which shared-memory matrix descriptors does tcgen05.mma accept for the A operand?
// Decides how halo tiles (one load, nine shifted taps) can be addressed.  Each variant issues ONE
// M128 x N16 x K16 bf16 MMA with B = [I16 | 0] so that D[m][n] = A[m][n]; A's smem image encodes the
// source row in elements k=0..2 (row & 15, (row>>4) & 15, row>>8) and k in elements 3..15, so D tells which
// smem row / chunk every MMA row actually read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/umma_probe tools/umma_probe.cu && gpurun_out/umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

struct Variant {
  const char* name;
  int layout;        // 2 = SWIZZLE_128B, 0 = SWIZZLE_NONE
  int start_bytes;   // offset of the descriptor start from the (1024-aligned) A region
  int lbo, sbo;      // bytes
  int base_offset;   // descriptor bits 49-51
  int kadv_bytes;    // extra start offset for a K advance (expect k values shifted)
  int row_pitch;     // expected: src(m) = row0 + (m/8)*group_pitch + m%8
  int row0, group_pitch, k0;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int layout, int start_bytes, int lbo, int sbo, int base_offset, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const int R = 320;                       // data rows
  const int PLANE = R * 16;                // bytes per k-chunk plane in the no-swizzle image
  // region 0 (SW128 image): R rows x 128 B.  region 1 (no-swizzle image): 8 planes x R rows x 16 B.  region 2: B (SW128).
  uint8_t* regA = gen;
  uint8_t* regN = gen + 48 * 1024;
  uint8_t* regB = gen + 96 * 1024;
  for (int i = threadIdx.x; i < R * 64; i += blockDim.x) {
    const int r = i / 64, kk = i % 64;
    float v = (kk == 0) ? (float)(r & 15) : (kk == 1) ? (float)((r >> 4) & 15) : (kk == 2) ? (float)(r >> 8) : (float)kk;
    __nv_bfloat16 b = __float2bfloat16(v);
    const int chunk = kk / 8, e = kk % 8;
    *reinterpret_cast<__nv_bfloat16*>(regA + r * 128 + ((chunk ^ (r & 7)) * 16) + e * 2) = b;   // TMA-style absolute swizzle
    *reinterpret_cast<__nv_bfloat16*>(regN + chunk * PLANE + r * 16 + e * 2) = b;               // [chunk][row][16B]
  }
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int n = i / 64, kk = i % 64;
    __nv_bfloat16 b = __float2bfloat16((kk == n) ? 1.0f : 0.0f);
    const int chunk = kk / 8, e = kk % 8;
    *reinterpret_cast<__nv_bfloat16*>(regB + n * 128 + ((chunk ^ (n & 7)) * 16) + e * 2) = b;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = (layout == 0 ? base + 48 * 1024 : base) + start_bytes;
    const uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) |
                           ((uint64_t)layout << 61);
    const uint32_t b_addr = base + 96 * 1024;
    const uint64_t bdesc = (uint64_t)((b_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (!done && clock64() - t0 > 2000000000ll) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  const int PLANE = 320 * 16;
  Variant vs[] = {
      {"sw128 base            sbo1024 bo0", 2, 0, 16, 1024, 0, 0, 0, 0, 8, 0},
      {"sw128 +1 row          sbo1024 bo0", 2, 128, 16, 1024, 0, 0, 0, 1, 8, 0},
      {"sw128 +1 row          sbo1024 bo1", 2, 128, 16, 1024, 1, 0, 0, 1, 8, 0},
      {"sw128 +3 rows         sbo1024 bo0", 2, 384, 16, 1024, 0, 0, 0, 3, 8, 0},
      {"sw128 +3 rows         sbo1024 bo3", 2, 384, 16, 1024, 3, 0, 0, 3, 8, 0},
      {"sw128 +8 rows         sbo1024 bo0", 2, 1024, 16, 1024, 0, 0, 0, 8, 8, 0},
      {"sw128 base            sbo1280 bo0", 2, 0, 16, 1280, 0, 0, 0, 0, 10, 0},
      {"sw128 +11 rows        sbo1280 bo0", 2, 1408, 16, 1280, 0, 0, 0, 11, 10, 0},
      {"sw128 +11 rows        sbo1280 bo3", 2, 1408, 16, 1280, 3, 0, 0, 11, 10, 0},
      {"sw128 base +32B (k16) sbo1024 bo0", 2, 32, 16, 1024, 0, 0, 0, 0, 8, 16},
      {"sw128 +11 rows +32B   sbo1280 bo0", 2, 1408 + 32, 16, 1280, 0, 0, 0, 11, 10, 16},
      {"none  base   lbo=plane sbo128    ", 0, 0, PLANE, 128, 0, 0, 0, 0, 8, 0},
      {"none  +1 row lbo=plane sbo128    ", 0, 16, PLANE, 128, 0, 0, 0, 1, 8, 0},
      {"none  base   lbo=plane sbo160    ", 0, 0, PLANE, 160, 0, 0, 0, 0, 10, 0},
      {"none  +11    lbo=plane sbo160    ", 0, 176, PLANE, 160, 0, 0, 0, 11, 10, 0},
      {"none  base   lbo=128 sbo=plane (roles swapped?)", 0, 0, 128, PLANE, 0, 0, 0, 0, 8, 0},
      {"none  +2 planes (k16) lbo=plane sbo128", 0, 2 * PLANE, PLANE, 128, 0, 0, 0, 0, 8, 16},
  };
  float* d_out; float h[128 * 16];
  cudaMalloc(&d_out, sizeof(h));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 2048);
  for (const Variant& v : vs) {
    cudaMemset(d_out, 0, sizeof(h));
    probe<<<1, 128, 100 * 1024 + 2048>>>(v.layout, v.start_bytes, v.lbo, v.sbo, v.base_offset, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-52s : CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    int ok_rows = 0, ok_k = 0;
    for (int m = 0; m < 128; ++m) {
      const int src = (int)h[m * 16 + 0] + 16 * (int)h[m * 16 + 1] + 256 * (int)h[m * 16 + 2];
      const int expect = v.row0 + (m / 8) * v.group_pitch + (m % 8);
      bool kok = true;
      for (int n = 3; n < 16; ++n) kok &= (h[m * 16 + n] == (float)(v.k0 + n));
      if (v.k0 == 0) { ok_rows += (src == expect); } else { ok_rows += 1; kok &= (h[m * 16] == (float)v.k0 && h[m * 16 + 1] == (float)(v.k0 + 1)); }
      ok_k += kok;
    }
    printf("%-52s : rows as expected %3d/128, k-chunks consistent %3d/128   first rows:", v.name, ok_rows, ok_k);
    for (int m : {0, 1, 7, 8, 9, 16, 127})
      printf(" %d->%d", m, (int)h[m * 16 + 0] + 16 * (int)h[m * 16 + 1] + 256 * (int)h[m * 16 + 2]);
    printf("  | k[3..5]=%g,%g,%g\n", h[3], h[4], h[5]);
  }
  return 0;
}
