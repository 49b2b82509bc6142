// Probe for tcgen05 block-scaled FP8 (kind::mxf8f6f4.block_scale): verifies, on the hardware,
//   (A) where tcgen05.cp.32x128b.warpx4 puts the bytes of a 512-byte scale-factor block in TMEM, and
//   (B) which scale a block-scaled MMA applies to every output row / column.
// Stand-alone (no torch): nvcc -gencode arch=compute_100a,code=sm_100a -o build/mxfp8_probe csrc/probe/mxfp8_probe.cu
#include "../common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace mg;

__device__ __forceinline__ void tc_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_mxf8_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n\t}\n"
      ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}
// no-swizzle smem descriptor (layout type 0), explicit LBO / SBO in bytes
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_mxf8(uint32_t M, uint32_t N, uint32_t a_sf, uint32_t b_sf) {
  return (b_sf << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | (1u << 23) /*scale = UE8M0*/ | ((M >> 4) << 24) | (a_sf << 29);
}

// ---------------- probe A: SF block -> TMEM via tcgen05.cp, read back with tcgen05.ld ----------------
__global__ void __launch_bounds__(128) probe_cp(uint32_t* out, uint32_t lbo, uint32_t sbo) {
  __shared__ __align__(1024) uint8_t sf[1024];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 1024; i += 128) sf[i] = (uint8_t)(i & 0xFF);        // byte i of the block holds (i mod 256)
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<32>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    tc_cp_32x128b_warpx4(tm, desc_noswz(smem_u32(sf), lbo, sbo));
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld_x16(tm + ((threadIdx.x & ~31u) << 16), r);
  tmem_ld_wait();
  for (int c = 0; c < 8; ++c) out[threadIdx.x * 8 + c] = r[c];
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<32>(tm);
}

// ---------------- probe B: one block-scaled MMA, operands written by hand (no swizzle, K-major, 8x16B core matrices) ------------
// A: 128 x 32 e4m3 (all 1.0), B: 256 x 32 e4m3 (all 1.0).  Core matrix = 8 rows x 16 bytes; K=32 bytes = 2 core columns.
// smem layout: [k-chunk (16 B)][row] i.e. byte(row, kbyte) = (kbyte/16) * rows*16 + row*16 + kbyte%16  => LBO = rows*16, SBO = 128.
__global__ void __launch_bounds__(128) probe_mma(float* out, int mode, uint32_t a_sf_id, uint32_t b_sf_id, uint32_t lbo_cp, uint32_t sbo_cp) {
  __shared__ __align__(1024) uint8_t sA[128 * 32];
  __shared__ __align__(1024) uint8_t sB[256 * 32];
  __shared__ __align__(1024) uint8_t sfa[512];
  __shared__ __align__(1024) uint8_t sfb[1024];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 128 * 32; i += 128) sA[i] = 0x38;   // e4m3 1.0
  for (int i = threadIdx.x; i < 256 * 32; i += 128) sB[i] = 0x38;
  // scale blocks in the "(32,4),(.,4) : (16,4),(.,1)" order: byte offset = (m%32)*16 + (m/32)*4 + kblock
  for (int i = threadIdx.x; i < 512; i += 128) {
    const int m0 = i / 16, m1 = (i % 16) / 4, kb = i % 4, m = m0 + 32 * m1;
    uint8_t v = 127;
    if (mode == 0) v = 127 + (m % 4);                 // SFA varies with the row
    if (mode == 2) v = 127 + kb;                      // SFA varies with the k-block id
    sfa[i] = v;
  }
  for (int i = threadIdx.x; i < 1024; i += 128) {
    const int blk = i / 512, j = i % 512, n0 = j / 16, n1 = (j % 16) / 4, kb = j % 4, n = blk * 128 + n0 + 32 * n1;
    uint8_t v = 127;
    if (mode == 1) v = 127 + (n % 8);                 // SFB varies with the column
    if (mode == 3) v = 127 + kb;
    sfb[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t D = tm, SFA = tm + 256, SFB = tm + 264;
  if (threadIdx.x == 0) {
    tc_cp_32x128b_warpx4(SFA, desc_noswz(smem_u32(sfa), lbo_cp, sbo_cp));
    tc_cp_32x128b_warpx4(SFB, desc_noswz(smem_u32(sfb), lbo_cp, sbo_cp));
    tc_cp_32x128b_warpx4(SFB + 4, desc_noswz(smem_u32(sfb + 512), lbo_cp, sbo_cp));
    const uint64_t da = desc_noswz(smem_u32(sA), 128 * 16, 128);
    const uint64_t db = desc_noswz(smem_u32(sB), 256 * 16, 128);
    umma_mxf8_ss(D, da, db, idesc_mxf8(128, 256, a_sf_id, b_sf_id), SFA, SFB, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // row = lane: dump columns 0..15 and 128..143 of every row
  uint32_t r[16], r2[16];
  tmem_ld_x16(D + ((threadIdx.x & ~31u) << 16), r);
  tmem_ld_x16(D + ((threadIdx.x & ~31u) << 16) + 128, r2);
  tmem_ld_wait();
  for (int c = 0; c < 16; ++c) { out[threadIdx.x * 32 + c] = __uint_as_float(r[c]); out[threadIdx.x * 32 + 16 + c] = __uint_as_float(r2[c]); }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  uint32_t* d; float* f;
  CK(cudaMalloc(&d, 128 * 8 * 4)); CK(cudaMalloc(&f, 128 * 32 * 4));
  const uint32_t cand[3][2] = {{16, 128}, {128, 16}, {512, 128}};
  for (auto& c : cand) {
    CK(cudaMemset(d, 0, 128 * 8 * 4));
    probe_cp<<<1, 128>>>(d, c[0], c[1]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== probe_cp lbo=%u sbo=%u : %s\n", c[0], c[1], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint32_t> h(128 * 8);
    CK(cudaMemcpy(h.data(), d, 128 * 8 * 4, cudaMemcpyDeviceToHost));
    for (int lane : {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 96, 127}) {
      printf("  lane %3d:", lane);
      for (int cc = 0; cc < 6; ++cc) printf(" %08x", h[lane * 8 + cc]);
      printf("\n");
    }
  }
  for (int mode = 0; mode < 4; ++mode) {
    for (uint32_t id = 0; id < (mode >= 2 ? 4u : 1u); ++id) {
      CK(cudaMemset(f, 0, 128 * 32 * 4));
      probe_mma<<<1, 128>>>(f, mode, mode == 2 ? id : 0, mode == 3 ? id : 0, 16, 128);
      cudaError_t e = cudaDeviceSynchronize();
      printf("== probe_mma mode=%d sf_id=%u : %s\n", mode, id, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      std::vector<float> h(128 * 32);
      CK(cudaMemcpy(h.data(), f, 128 * 32 * 4, cudaMemcpyDeviceToHost));
      for (int row : {0, 1, 2, 3, 4, 32, 33, 64, 127}) {
        printf("  row %3d: cols0-9:", row);
        for (int cc = 0; cc < 10; ++cc) printf(" %g", h[row * 32 + cc]);
        printf(" | cols128-131:");
        for (int cc = 16; cc < 20; ++cc) printf(" %g", h[row * 32 + cc]);
        printf("\n");
      }
    }
  }
  printf("probe done\n");
  return 0;
}
