// MXFP8 block-scaled GEMM for sm_100a:  D[M,N] = (A ∘ SFA) · (B ∘ SFB)ᵀ  with A [M,K], B [N,K] in e4m3 and one UE8M0
// scale per 32 consecutive K-elements (OCP MX format), computed by tcgen05.mma.kind::mxf8f6f4.block_scale with the scale
// factors resident in tensor memory.
//
// Verified on hardware with csrc/probe/mxfp8_probe.cu (profiles/mxfp8_probe_r1.txt):
//   * a 512-byte scale block (128 rows x 4 k-blocks, byte offset (r%32)*16 + (r/32)*4 + kb) is moved to TMEM with
//     tcgen05.cp.32x128b.warpx4 through a no-swizzle descriptor (SBO = 128 B) and occupies 4 TMEM columns;
//   * the MMA picks the scale of row r from lane r%32(+32q), column r/32, byte = sf_id of the instruction descriptor.
//
// Structure = csrc/gemm.cu (TMA producer warp, single-thread MMA issuer, 8 epilogue warps, persistent M-grouped tiles) with:
//   * 128 x 256 tiles, BK = 128 fp8 elements (128-byte swizzled rows), 4 stages of A 16 KB + B 32 KB + SFA 512 B + SFB 1 KB;
//   * TMEM: accumulator 256 columns + 4 (SFA) + 8 (SFB) columns — the accumulator is single buffered (2 x 256 + 12 > 512),
//     so the epilogue first drains its 128 columns into registers, releases the accumulator at once, and then converts and
//     writes bf16 straight to global memory while the next tile's MMAs already run.
#include "common.cuh"

#include <stdexcept>
#include <string>

namespace mg {

CUtensorMap make_tmap_2d(const void* ptr, int elem_bytes, long long rows, long long cols, long long ld, int box_cols, int box_rows);

namespace mx {

constexpr int BM = 128, BN = 256, BKB = 128, STAGES = 4;     // BKB: bytes (= fp8 elements) of K per stage
constexpr int A_BYTES = BM * BKB, B_BYTES = BN * BKB, SFA_BYTES = 512, SFB_BYTES = 1024;
constexpr int AB_BYTES = A_BYTES + B_BYTES;                    // 48 KB, keeps every A/B tile 1024-byte aligned
constexpr int SF_BYTES = SFA_BYTES + SFB_BYTES;
constexpr int SMEM_BYTES = STAGES * (AB_BYTES + SF_BYTES) + 512;
constexpr int GROUP_M = 8;
constexpr int NUM_THREADS = 320, EPI_THREADS = 256;
constexpr uint32_t D_COL = 0, SFA_COL = 256, SFB_COL = 260;

enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_GELU = 2 };

struct Params {
  const uint8_t* sfa;            // [ceil(M/128)][K/128][512]
  const uint8_t* sfb;            // [ceil(N/128)][K/128][512]
  __nv_bfloat16* d;
  __nv_bfloat16* d2;
  const __nv_bfloat16* aux;
  long long ldd, ldd2, ld_aux;
  int M, N, K;
};

MG_DEVICE void tc_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
MG_DEVICE void umma_mxf8_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n\t}\n"
      ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}
MG_DEVICE uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_mxf8(uint32_t M, uint32_t N) {
  return (0u << 7) | (0u << 10) | ((N >> 3) << 17) | (1u << 23) /* UE8M0 scales */ | ((M >> 4) << 24);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_mxfp8_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sSF = smem + STAGES * AB_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sSF + STAGES * SF_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = p.K / BKB;                                // K % 128 == 0 (checked on the host)
  const int n_blocks128 = (p.N + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = GROUP_M * num_n;
    const int g = tile / per_group;
    const int first_m = g * GROUP_M;
    const int gsize = min(num_m - first_m, GROUP_M);
    const int r = tile - g * per_group;
    m_blk = first_m + r % gsize;
    n_blk = r / gsize;
  };

  if (warp == 0) {
    // ===================== producer: TMA tiles + bulk copies of the scale blocks =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int nb0 = 2 * n_blk, nb1 = 2 * n_blk + 1;
        const bool has_b1 = nb1 < n_blocks128;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], AB_BYTES + SFA_BYTES + (has_b1 ? 1024 : 512));
          uint8_t* sA = smem + stage * AB_BYTES;
          uint8_t* sB = sA + A_BYTES;
          uint8_t* sfa = sSF + stage * SF_BYTES;
          uint8_t* sfb = sfa + SFA_BYTES;
          tma_load_2d(sA, &tmA, &full_bar[stage], kb * BKB, m_blk * BM);
          tma_load_2d(sB, &tmB, &full_bar[stage], kb * BKB, n_blk * BN);
          bulk_load_1d(sfa, p.sfa + (static_cast<long long>(m_blk) * num_k + kb) * 512, 512, &full_bar[stage]);
          bulk_load_1d(sfb, p.sfb + (static_cast<long long>(nb0) * num_k + kb) * 512, 512, &full_bar[stage]);
          if (has_b1) bulk_load_1d(sfb + 512, p.sfb + (static_cast<long long>(nb1) * num_k + kb) * 512, 512, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp converged (descriptors stay in uniform registers), one elected lane issues each stage's tcgen05 instructions
    {
      constexpr uint32_t idesc0 = idesc_mxf8(BM, BN);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), 16, 1024), b_desc0 = make_smem_desc(smem_u32(smem) + A_BYTES, 16, 1024);
      const uint64_t sf_desc0 = desc_noswz(smem_u32(sSF), 16, 128);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1) ^ 1);                  // epilogue of the previous tile has drained the accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = desc_shift(a_desc0, stage * AB_BYTES), b_desc = desc_shift(b_desc0, stage * AB_BYTES);
          const uint64_t sf_desc = desc_shift(sf_desc0, stage * SF_BYTES);
          if (elect_one()) {
            // scale factors smem -> TMEM (tcgen05.cp and tcgen05.mma execute in issue order: no hazard with the previous stage)
            tc_cp_32x128b_warpx4(tmem + SFA_COL, sf_desc);
            tc_cp_32x128b_warpx4(tmem + SFB_COL, desc_shift(sf_desc, 512));
            tc_cp_32x128b_warpx4(tmem + SFB_COL + 4, desc_shift(sf_desc, 1024));
#pragma unroll
            for (int k = 0; k < 4; ++k) {                     // 4 x K=32 per stage; sf_id = k selects the k-block's scale byte
              const uint32_t idesc = idesc0 | (static_cast<uint32_t>(k) << 4) | (static_cast<uint32_t>(k) << 29);
              umma_mxf8_ss(tmem + D_COL, desc_shift(a_desc, k * 32), desc_shift(b_desc, k * 32), idesc, tmem + SFA_COL, tmem + SFB_COL, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (kb + 1 == num_k) umma_commit(tfull_bar);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> bf16 -> global (direct 16-byte stores) =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t t_row = tmem + (static_cast<uint32_t>(q * 32) << 16) + D_COL;
      const long long grow = static_cast<long long>(m_blk) * BM + row;
      const bool row_ok = grow < p.M;
      // Register-staged epilogue: this thread's 128 accumulator columns are pulled out of TMEM in one go and the (single)
      // accumulator buffer is handed back to the MMA warp immediately; conversion and stores then overlap the next tile's MMAs.
      uint32_t racc[BN / 64][32];
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) tmem_ld_x32(t_row + c * 64 + half * 32, racc[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty_bar);
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) {
        const uint32_t(&r)[32] = racc[c];
        const int col = n_blk * BN + c * 64 + half * 32;
        if (row_ok && col < p.N) {                             // N % 32 == 0 (host check): a 32-column slab is all-in or all-out
          __nv_bfloat16* dst = p.d + grow * p.ldd + col;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(r[8 * j + e]);
            if (EPI == EPI_RESIDUAL) {
              float a[8];
              unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(p.aux + grow * p.ld_aux + col) + j), a);
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] += a[e];
            }
            const uint4 o = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            reinterpret_cast<uint4*>(dst)[j] = o;
            if (EPI == EPI_GELU) {
              const uint32_t ou[4] = {o.x, o.y, o.z, o.w};
              uint32_t g[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xr = unpack_bf16(ou[e]);          // GELU of the rounded pre-activation (what backward sees)
                g[e] = pack_bf16(gelu_tanh(xr.x), gelu_tanh(xr.y));
              }
              reinterpret_cast<uint4*>(p.d2 + grow * p.ldd2 + col)[j] = make_uint4(g[0], g[1], g[2], g[3]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int EPI>
static void launch(const CUtensorMap& tA, const CUtensorMap& tB, const Params& p, int num_sms, cudaStream_t stream) {
  auto kern = gemm_mxfp8_tcgen05<EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(gemm_mx): ") + cudaGetErrorString(e));
    configured = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tA, tB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_mx launch: ") + cudaGetErrorString(e));
}

}  // namespace mx

void gemm_mx_launch(const void* a, const void* sfa, const void* b, const void* sfb, void* d, long long ldd, void* d2, long long ldd2,
                    const void* aux, long long ld_aux, int M, int N, int K, int epi, int num_sms, cudaStream_t stream) {
  if (K % 128 || N % 32) throw std::runtime_error("gemm_mx: K must be a multiple of 128 and N of 32");
  CUtensorMap tA = make_tmap_2d(a, 1, M, K, K, mx::BKB, mx::BM);
  CUtensorMap tB = make_tmap_2d(b, 1, N, K, K, mx::BKB, mx::BN);
  mx::Params p{static_cast<const uint8_t*>(sfa), static_cast<const uint8_t*>(sfb), static_cast<__nv_bfloat16*>(d),
               static_cast<__nv_bfloat16*>(d2), static_cast<const __nv_bfloat16*>(aux), ldd, ldd2, ld_aux, M, N, K};
  if (epi == mx::EPI_STORE) return mx::launch<mx::EPI_STORE>(tA, tB, p, num_sms, stream);
  if (epi == mx::EPI_RESIDUAL) return mx::launch<mx::EPI_RESIDUAL>(tA, tB, p, num_sms, stream);
  if (epi == mx::EPI_GELU) return mx::launch<mx::EPI_GELU>(tA, tB, p, num_sms, stream);
  throw std::runtime_error("gemm_mx: unsupported epilogue");
}

}  // namespace mg
