// Probe for the 2-CTA (cta_group::2) tcgen05 GEMM building blocks: one 256 x 256 x (64*KB) bf16 tile computed by a
// CTA pair (cluster 2x1x1).  Each CTA TMA-loads its 128 rows of A and its 128 rows (half) of B into its own smem; all loads
// signal the LEADER's full barrier; the leader issues tcgen05.mma.cta_group::2 (M = 256), commits with .multicast::cluster to
// both CTAs' barriers; each CTA's warps read their 128 x 256 accumulator half from their own TMEM and write fp32 to global.
// Stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a -o cta2_probe.bin csrc/probe/cta2_probe.cu -lcuda
#include "../common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace mg;

MG_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
MG_DEVICE void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
MG_DEVICE uint32_t mapa_rank0(uint32_t smem_addr) {       // same offset in the leader CTA's shared window
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(0));
  return r;
}
template <uint32_t kCols>
MG_DEVICE void tmem_alloc2(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
MG_DEVICE void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load executed by either CTA of the pair; completion bytes are credited to the mbarrier at cluster address `bar_cluster`
MG_DEVICE void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
MG_DEVICE void umma_bf16_ss_cg2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
MG_DEVICE void umma_commit_mc2(uint64_t* bar) {            // arrive on the barrier at this smem offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

constexpr int BK = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
cta2_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                    // 128 x 64 bf16 = 16 KB
  uint8_t* sB = smem + 16384;            // 128 x 64 bf16 = 16 KB (this CTA's half of the 256-row B tile)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint64_t* empty_bar = full_bar + 1;
  uint64_t* done_bar = full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 3);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = K / BK;
  if (threadIdx.x == 0) {
    mbar_init(full_bar, 1);              // leader: its own arrive.expect_tx covers the bytes of both CTAs
    mbar_init(empty_bar, 1);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // producer (both CTAs): own A rows, own half of B; everything signals the leader's full barrier
    const uint32_t full_leader = mapa_rank0(smem_u32(full_bar));
    for (int kb = 0; kb < num_k; ++kb) {
      mbar_wait(empty_bar, (kb & 1) ^ 1);
      if (rank == 0) mbar_expect_tx(full_bar, 4 * 16384);
      tma_load_2d_cg2(sA, &tmA, full_leader, kb * BK, rank * 128);
      tma_load_2d_cg2(sB, &tmB, full_leader, kb * BK, rank * 128);
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // MMA issuer: leader CTA only
    const uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
    for (int kb = 0; kb < num_k; ++kb) {
      mbar_wait(full_bar, kb & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      for (int k = 0; k < BK / 16; ++k)
        umma_bf16_ss_cg2(tmem, make_smem_desc(a_base + k * 32, 16, 1024), make_smem_desc(b_base + k * 32, 16, 1024), idesc, (kb | k) != 0);
      umma_commit_mc2(empty_bar);        // both CTAs' smem slots are free again
    }
    umma_commit_mc2(done_bar);           // accumulators complete (both CTAs)
  }
  if (warp >= 2) {
    // epilogue in both CTAs: rows rank*128 + q*32 + lane, all 256 columns (fp32)
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = rank * 128 + q * 32 + lane;
    for (int c = 0; c < 256; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int e = 0; e < 32; ++e) out[row * 256 + c + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2<256>(tmem);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap tmap(void* p, int rows, int cols, int box_rows) {
  cudaDriverEntryPointQueryResult q; void* fnp = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = ((PFN_enc)fnp)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  const int M = 256, N = 256, K = 512;
  std::vector<__nv_bfloat16> a(M * K), b(N * K);
  std::vector<float> af(M * K), bf(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.f; a[i] = __float2bfloat16(v); af[i] = __bfloat162float(a[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.f; b[i] = __float2bfloat16(v); bf[i] = __bfloat162float(b[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  CK(cudaMalloc(&da, M * K * 2)); CK(cudaMalloc(&db, N * K * 2)); CK(cudaMalloc(&dout, M * N * 4));
  CK(cudaMemcpy(da, a.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, M * N * 4));
  CUtensorMap tA = tmap(da, M, K, 128), tB = tmap(db, N, K, 128);
  CK(cudaFuncSetAttribute(cta2_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34000));
  cta2_probe<<<2, 192, 34000>>>(tA, tB, dout, K, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("cta2_probe: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> o(M * N);
  CK(cudaMemcpy(o.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)af[m * K + k] * bf[n * K + k];
    double d = fabs(ref - o[m * N + n]); if (d > maxerr) maxerr = d; if (d > 1e-2) { if (bad < 8) printf("  mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, o[m * N + n]); ++bad; }
  }
  printf("max abs err %.3g, mismatches %d / %d\n", maxerr, bad, M * N);
  printf("probe done\n");
  return 0;
}
