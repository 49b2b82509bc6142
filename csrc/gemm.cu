// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (fp32 accumulators in TMEM, double buffered) -> tcgen05.ld epilogue -> swizzled smem -> TMA store.
//
//   D[M,N] (+)= A[M,K] · B[N,K]^T          A, B bf16; each operand either K-major or MN-major in memory
//
// One kernel family serves every linear layer of the model:
//   forward   y  = x · Wᵀ           A K-major,  B K-major   (+ residual | + tanh-GELU dual output)
//   dgrad     dx = dy · W           A K-major,  B MN-major  (· gelu'(pre) for the MLP up-projection)
//   wgrad     dW += dyᵀ · x         A MN-major, B MN-major  fp32 output, TMA reduce-add accumulation
//
// Roles (320 threads, pair kernels 384; 1 CTA / SM, grid = #SMs, static persistent tile schedule with M-grouping for L2 reuse):
//   warp 0   TMA producer   (one elected lane)
//   warp 1   TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warp 10-11 (pair kernels) copy warps: the TMA bulk-copy rings of a carried all-gather (see GatherParams); idle otherwise
//   warp 2-9 epilogue: TMEM -> registers -> epilogue math -> smem staging -> TMA store / reduce-add.
//            Two warps share each TMEM lane quarter and take alternate column chunks; every warp owns a
//            32-row staging box and issues its own TMA stores -- no barrier couples the epilogue warps.
// Pipelines: smem ring (full/empty mbarriers, 4 stages of 128x64 A + 256x64 B), TMEM ring (2 x 256 columns).
#include "common.cuh"

#include <map>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <vector>
#include <string>

namespace mg {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;             // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;             // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int CSTAGE_BYTES = BM * 128;                 // 128 rows x 128 B (64 bf16 or 32 fp32 per row)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * CSTAGE_BYTES + 512 + 1024;  // + barriers + align slack
constexpr int GROUP_M = 8;
constexpr int NUM_THREADS = 320;     // 2 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;

enum Epilogue : int {
  EPI_STORE = 0,      // D = acc                       (bf16)
  EPI_RESIDUAL = 1,   // D = acc + AUX                 (bf16)
  EPI_GELU = 2,       // D = acc, D2 = gelu(acc)       (bf16, bf16)
  EPI_DGELU = 3,      // D = acc * gelu'(AUX)          (bf16)
  EPI_F32 = 4,        // D (+)= acc                    (fp32; accumulate via TMA reduce-add)
};

// Fused all-gather (all-gather ⊕ GEMM).  The weights of an FSDP unit live in a flat buffer whose `world` equal
// chunks are owned by different ranks.  A pair-kernel GEMM (CARRY instantiations) can CARRY the gather of a whole unit: two extra
// warps in EVERY CTA each drive a ring of four 4 KB bulk copies -- peer shard -> shared memory -> local unit buffer, the bytes never
// touch registers -- over the remote 64 KB segments of the unit, in address order starting at the carrier's own B operand, and
// publish one release-flag per segment.  No SM is taken away from the GEMM (round 1 dedicated ~24 copy-engine CTAs and lost their
// tiles).  If the carrier's B operand is itself part of the unit (forward QKV / LM-head GEMM: K-major rows; dgrad: MN-major k-rows)
// its TMA producers acquire the flags of the segments below the rows they are about to load (a watermark: segments land in address
// order and every CTA's tiles ascend through B).  Later kernels of the stream need no flags at all: the carrier does not retire
// before its copy warps have finished -- which is how the engine uses it most of the time: the gather of unit l+1 is carried by a
// GEMM of unit l (parallel/fsdp.py), because NVLink traffic issued next to a saturating tcgen05 mainloop is capped near 140 GB/s.
struct GatherParams {
  void* peer[8];                    // peer[c] = rank c's bf16 shard (symmetric memory), length chunk_elems
  __nv_bfloat16* local;             // local gathered unit buffer, world * chunk_elems elements
  unsigned int* flags;              // one word per SEG_ELEMS-element segment of the unit; == target (epoch) once the segment landed
  long long chunk_elems;            // elements per chunk (multiple of 8)
  long long b_offset;               // element offset of this GEMM's B matrix inside the unit buffer; < 0: B is not in the unit (no polling)
  long long first_seg;              // the pull starts at this segment (the carrier's own B) and wraps around
  unsigned int target;              // epoch of this call (monotonically increasing per flags buffer)
  int world, rank;
  int n_seg;                        // segments of the unit; 0 = this GEMM carries no gather
};
constexpr long long SEG_ELEMS = 32768;   // 64 KB of bf16: flag granularity of the fused gather (a 256-row weight tile at K=2048 is 16 segments)
constexpr int COPY_THREADS = 64;         // two copy warps per CTA (pair kernels only)

struct GemmParams {
  const __nv_bfloat16* aux;
  long long ld_aux;
  int M, N, K;
  int accumulate;
  GatherParams ag;
  int group_m;                      // rasterisation: m-blocks per group (groups sweep all n-blocks), chosen on the host
  int split_k;                      // > 0 only for the reduce-add fp32 epilogue (wgrad): K-slices of the leftover tiles (see WorkIter)
  __nv_bfloat16* d2;                // EPI_GELU: second output (gelu(pre)), written with direct 16-byte stores
  long long ldd2;
  const float* row_scale_a;         // FP8 only: de-quantisation multipliers, one per row of A (token) …
  const float* row_scale_b;         // … and one per row of B (output feature); acc[m,n] *= sa[m] * sb[n]
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// FP8 = true: A and B hold e4m3 bytes (row-wise scaled, see quant_rows_e4m3 in elementwise.cu); the smem/TMA byte geometry
// is unchanged (128-byte rows = 128 fp8 elements = one BK), tcgen05.mma runs kind::f8f6f4 with K = 32 per instruction.
// CG2 = true: CTA-pair kernel (cluster 2x1x1, tcgen05 cta_group::2).  A pair owns a 256 x 256 output tile: each CTA loads its
// 128 rows of A and HALF of the B tile (128 rows), the leader issues M = 256 MMAs that read both CTAs' smem, each CTA's TMEM
// receives its 128 accumulator rows.  Per-SM smem traffic for B and the L2 -> SM traffic for B halve; 6 stages instead of 4.
// Work list of one worker (a CTA, or a CTA pair).  Plain mode (split == 0): whole tiles worker, worker + n, ... (grouped
// rasterisation keeps the tiles of a wave next to each other in L2).  Hybrid split-K (accumulating fp32 epilogue = wgrad, where
// the output has only 64-256 tiles but K = #tokens is huge): the full waves are done the same way, and only the LEFTOVER
// tiles (fewer than one wave) are cut into `split` K-slices, handed out slice-major so that the workers running at any moment
// share one k-range of neighbouring tiles in L2; every piece adds into D with the TMA reduce-add the epilogue uses anyway.
// (Measured alternatives: plain split-K of ALL tiles reduce-adds every tile 3-10 times; equal contiguous k-ranges per worker
// -- "stream-K" -- balance perfectly but share nothing in L2 and turned the 64-tile attn_proj wgrad HBM-bound, +44 %.)
struct WorkIter {
  int n_workers, num_k, full_tiles, left, kps;
  int t;          // next whole tile
  int i, n_items; // leftover items (slice-major): slice = i / left, tile = full_tiles + i % left
  __host__ MG_DEVICE WorkIter(int worker, int n_workers_, int tiles_mn, int num_k_, int split) : n_workers(n_workers_), num_k(num_k_) {
    full_tiles = split > 0 ? (tiles_mn / n_workers) * n_workers : tiles_mn;
    left = tiles_mn - full_tiles;
    kps = split > 0 ? (num_k + split - 1) / split : num_k;
    n_items = left * (split > 0 ? (num_k + kps - 1) / kps : 0);       // only the non-empty slices
    t = worker < 0 ? full_tiles : worker;                             // (worker < 0: owns no tiles)
    i = worker < 0 ? n_items : worker;
  }
  __host__ MG_DEVICE bool next(int& tile, int& kb0, int& kb1) {
    if (t < full_tiles) { tile = t; kb0 = 0; kb1 = num_k; t += n_workers; return true; }
    if (i < n_items) {
      const int slice = i / left;
      tile = full_tiles + (i - slice * left);
      kb0 = slice * kps;
      kb1 = kb0 + kps < num_k ? kb0 + kps : num_k;
      i += n_workers;
      return true;
    }
    return false;
  }
};

template <bool A_MN, bool B_MN, int EPI, bool AG = false, bool FP8 = false, bool CG2 = false, bool AUXB = false, bool CARRY = false>
__global__ void __launch_bounds__(CG2 ? NUM_THREADS + COPY_THREADS : NUM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
                  const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Pair kernels whose epilogue reads a second [M, N] operand (residual / pre-activation) or writes a second output (gelu)
  // move that tile by TMA through a second 4 KB box per epilogue warp (AUXBUF) and give up one of the six operand stages for it:
  // 16-byte-per-lane global accesses with one row per lane cost 32 L1 wavefronts per instruction, and those wavefronts compete
  // with the tensor core's operand reads for the same shared-memory / L1 bandwidth.
  constexpr bool HAS_AUX = (EPI == EPI_RESIDUAL || EPI == EPI_DGELU);
  constexpr bool AUXBUF = CG2 && AUXB && !CARRY && (HAS_AUX || EPI == EPI_GELU);
  // CARRY: this instantiation can carry the all-gather of an FSDP unit (see GatherParams): the sixth operand stage becomes the
  // 2 x 16 KB bulk-copy rings of the two copy warps.
  static_assert(!CARRY || CG2, "only pair kernels carry a gather");
  constexpr int NST = CG2 ? ((AUXBUF || CARRY) ? 5 : 6) : STAGES;   // pipeline depth
  constexpr int STG = CG2 ? 2 * A_STAGE_BYTES : STAGE_BYTES;     // bytes per stage in this CTA (A 16 KB + B 16 KB (half) | 32 KB)
  static_assert(NST * STG + ((AUXBUF || CARRY) ? 8 * 4096 : 0) == STAGES * STAGE_BYTES, "operand ring (+ aux boxes / copy rings) must end where the barriers begin");
  uint8_t* sC = smem + NST * STG;                                // 8 x 4 KB staging boxes
  uint8_t* sAux = sC + 8 * 4096;                                 // 8 x 4 KB aux boxes (AUXBUF) | 2 x 4 x 4 KB copy rings (CARRY)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 2 * CSTAGE_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tfull_bar = empty_bar + NST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* aux_bar = full_bar + 24;                             // [8] one per epilogue warp (AUXBUF only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;
  const int num_m = CG2 ? (p.M + 2 * BM - 1) / (2 * BM) : (p.M + BM - 1) / BM;   // CG2: tiles are 256 rows tall
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles_mn = num_m * num_n;
  constexpr int BKE = FP8 ? 2 * BK : BK;            // elements per k-block (a k-block is always 128 bytes per row)
  const int num_k_total = (p.K + BKE - 1) / BKE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (AUXBUF) tma_prefetch_desc(&tmD2);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    if (AUXBUF || CARRY)
      for (int w = 0; w < 8; ++w) mbar_init(&aux_bar[w], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], CG2 ? 2 * EPI_THREADS : EPI_THREADS);     // CG2: the leader waits for both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG2) tmem_alloc2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool carry = CARRY && p.ag.n_seg > 0;             // this launch carries the gather of an FSDP unit (copy warps 10-11)
  const bool poll = carry && p.ag.b_offset >= 0;           // ... and its own B operand is part of that unit
  const int tile0 = CG2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_stride = CG2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  // tile -> (m_blk, n_blk)
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = p.group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * p.group_m;
    const int gsize = min(num_m - first_m, p.group_m);
    const int r = tile - g * per_group;
    m_blk = first_m + r % gsize;
    n_blk = r / gsize;
    if (CG2) m_blk = 2 * m_blk + static_cast<int>(cta_rank);      // this CTA's 128-row block inside the pair's 256-row tile
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // fused gather: every segment below `watermark` is known to have landed (segments arrive in address order starting at
      // this GEMM's B, and this CTA's tiles / k-blocks ascend through B, so each flag is acquired at most once per CTA)
      long long watermark = poll ? p.ag.b_offset / SEG_ELEMS : 0;
      auto acquire_upto = [&](long long e_end) {
        if (watermark * SEG_ELEMS >= e_end) return;
        const long long loc0 = static_cast<long long>(p.ag.rank) * p.ag.chunk_elems, loc1 = loc0 + p.ag.chunk_elems;
        while (watermark * SEG_ELEMS < e_end) {
          const long long s0 = watermark * SEG_ELEMS, s1 = s0 + SEG_ELEMS;
          if (!(s0 >= loc0 && s1 <= loc1))                    // (locally owned segments are in place before the launch)
            while (static_cast<int>(ld_acquire_gpu_u32(p.ag.flags + watermark) - p.ag.target) < 0) {
            }
          ++watermark;
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");   // the copy warps' generic-proxy stores -> this thread's TMA loads
      };
      WorkIter work(tile0, tile_stride, tiles_mn, num_k_total, p.split_k);
      for (int tile, kb0, kb1; work.next(tile, kb0, kb1);) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        if (poll && !B_MN) {       // K-major B: rows [.., b_row_end) of the weight, whole rows
          const int b_row_end = min(p.N, n_blk * BN + (static_cast<int>(cta_rank) + 1) * (BN / 2));
          acquire_upto(p.ag.b_offset + static_cast<long long>(b_row_end) * p.K);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          if (poll && B_MN)          // MN-major B (dgrad): k-rows [.., (kb+1)*BK) of the weight, whole rows of N elements
            acquire_upto(p.ag.b_offset + static_cast<long long>(min(p.K, (kb + 1) * BK)) * p.N);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STG;
          uint8_t* sB = sA + A_STAGE_BYTES;
          if (CG2) {
            // both CTAs load; every byte is credited to the LEADER's full barrier, which the leader arms for the pair's 64 KB
            const uint32_t full_leader = mapa_rank0(smem_u32(&full_bar[stage]));
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STG);
            const int b_row0 = n_blk * BN + static_cast<int>(cta_rank) * (BN / 2);
            if (!A_MN) {
              tma_load_2d_cg2(sA, &tmA, full_leader, kb * BKE, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sA + j * (BK * 128), &tmA, full_leader, m_blk * BM + j * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d_cg2(sB, &tmB, full_leader, kb * BKE, b_row0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_load_2d_cg2(sB + j * (BK * 128), &tmB, full_leader, b_row0 + j * 64, kb * BK);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(sA, &tmA, &full_bar[stage], kb * BKE, m_blk * BM);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(sB, &tmB, &full_bar[stage], kb * BKE, n_blk * BN);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
            }
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (converged) control flow so that descriptors and addresses live in uniform registers; one
    // elected lane issues each k-block's tcgen05 instructions.  (With `if (lane == 0)` around the loop the compiler wraps
    // every MMA in a register->uniform-register waterfall and rebuilds both descriptors: ~20 instructions per MMA.)
    if (!CG2 || cta_rank == 0) {                          // CG2: only the leader CTA issues MMAs
      constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(CG2 ? 2 * BM : BM, BN) : make_idesc_bf16(CG2 ? 2 * BM : BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      const uint64_t a_desc0 = A_MN ? make_smem_desc(smem_u32(smem), BK * 128, 1024) : make_smem_desc(smem_u32(smem), 16, 1024);
      const uint64_t b_desc0 = B_MN ? make_smem_desc(smem_u32(smem) + A_STAGE_BYTES, BK * 128, 1024)
                                    : make_smem_desc(smem_u32(smem) + A_STAGE_BYTES, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      WorkIter work(tile0, tile_stride, tiles_mn, num_k_total, p.split_k);
      for (int tile, kb0, kb1; work.next(tile, kb0, kb1); ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], tphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = desc_shift(a_desc0, stage * STG), b_desc = desc_shift(b_desc0, stage * STG);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = desc_shift(a_desc, A_MN ? k * 2048 : k * 32);
              const uint64_t db = desc_shift(b_desc, B_MN ? k * 2048 : k * 32);
              const uint32_t acc = ((kb - kb0) | k) != 0 ? 1u : 0u;
              if (FP8 && CG2) umma_fp8_ss_cg2(d_tmem, da, db, idesc, acc);
              else if (FP8) umma_fp8_ss(d_tmem, da, db, idesc, acc);
              else if (CG2) umma_bf16_ss_cg2(d_tmem, da, db, idesc, acc);
              else umma_bf16_ss(d_tmem, da, db, idesc, acc);
            }
            if (CG2) umma_commit_mc2(&empty_bar[stage]);   // frees the smem slot in BOTH CTAs when these MMAs retire
            else umma_commit(&empty_bar[stage]);
            if (kb + 1 == kb1) {
              if (CG2) umma_commit_mc2(&tfull_bar[buf]);   // accumulator complete -> both CTAs' epilogues
              else umma_commit(&tfull_bar[buf]);
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 2 + EPI_THREADS / 32) {
    // ===================== copy warps (pair kernels): the all-gather this GEMM carries =====================
    // One elected thread per copy warp drives a ring of four 4 KB bulk copies: peer shard --cp.async.bulk--> shared memory
    // --cp.async.bulk--> local unit buffer.  The bytes never touch registers or the LSU: measured, ordinary global loads issued
    // next to a running tcgen05 GEMM see ~17 us of latency (its TMA operand traffic owns the SM's path to L2), which capped every
    // register-based copy loop -- pull or multicast push, 24 dedicated CTAs or 148 x 64 threads -- at ~140 GB/s; the TMA path is
    // the one the GEMM's own operands use and is served in about a microsecond.
    if (CARRY && carry && lane == 0) {
      const int w = warp - (2 + EPI_THREADS / 32);                          // 0 | 1
      uint8_t* ring = sAux + w * (4 * 4096);
      uint64_t* full = aux_bar + w * 4;
      const long long unit_elems = p.ag.chunk_elems * p.ag.world;
      const long long loc0 = static_cast<long long>(p.ag.rank) * p.ag.chunk_elems, loc1 = loc0 + p.ag.chunk_elems;
      uint32_t head = 0, tail = 0;                                          // pieces loaded / pieces stored
      long long q_dst[4];
      uint32_t q_bytes[4];
      auto drain_one = [&]() {                                              // oldest piece: landed in shared memory -> store it
        const uint32_t sl = tail & 3;
        mbar_wait(&full[sl], (tail >> 2) & 1);
        bulk_store_1d(p.ag.local + q_dst[sl], ring + sl * 4096, q_bytes[sl]);
        tma_commit_group();
        ++tail;
      };
      for (long long i = static_cast<long long>(blockIdx.x) * 2 + w; i < p.ag.n_seg; i += 2LL * gridDim.x) {
        long long seg = p.ag.first_seg + i;
        if (seg >= p.ag.n_seg) seg -= p.ag.n_seg;
        const long long e0 = seg * SEG_ELEMS, e1 = (e0 + SEG_ELEMS < unit_elems) ? e0 + SEG_ELEMS : unit_elems;
        if (e0 >= loc0 && e1 <= loc1) continue;                              // owned by this rank: already in place
        for (long long cur = e0; cur < e1;) {
          const long long owner = cur / p.ag.chunk_elems;
          const long long own_end = (owner + 1) * p.ag.chunk_elems;
          const long long lim = e1 < own_end ? e1 : own_end;
          if (owner == p.ag.rank) { cur = lim; continue; }
          const long long left = (lim - cur) * 2;
          const uint32_t bytes = left < 4096 ? static_cast<uint32_t>(left) : 4096u;
          if (head - tail == 4) drain_one();
          if (head >= 4) tma_wait_group_read<0>();                           // the slot's previous store has read its bytes
          const uint32_t sl = head & 3;
          mbar_expect_tx(&full[sl], bytes);
          bulk_load_1d(ring + sl * 4096, reinterpret_cast<const __nv_bfloat16*>(p.ag.peer[owner]) + (cur - owner * p.ag.chunk_elems), bytes, &full[sl]);
          q_dst[sl] = cur;
          q_bytes[sl] = bytes;
          ++head;
          cur += bytes / 2;
        }
        while (tail < head) drain_one();
        tma_wait_group<0>();                                                 // the segment's bytes are in the local buffer
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.ag.flags + seg), "r"(p.ag.target) : "memory");
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // Every warp is its own pipeline: it owns a 32-row x 128-byte (4 KB) staging box and issues its own TMA stores, so no
    // barrier ever couples the eight warps (round 1 synchronised all 256 threads twice per 64-column chunk and a single thread
    // issued every store: with K = 768 -- 12 k-blocks per tile -- the GELU / DGELU tiles were epilogue-bound at 0.7 PFLOP/s).
    // Warp (q, half) drains the 64-column (fp32: 32-column) chunks of its TMEM lane quarter whose index has parity `half`;
    // the operand of the fused element-wise op (residual / pre-activation) is fetched into registers BEFORE the accumulator
    // wait, and the TMEM buffer is handed back right after this thread's last tcgen05.ld of the tile.
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // chunk parity this warp handles
    const int row = q * 32 + lane;                // row inside the 128-row tile
    uint8_t* sbuf = sC + (warp - 2) * 4096;       // this warp's staging box (1024-byte aligned: swizzle-128B atom)
    uint8_t* rowp = sbuf + lane * 128;
    const uint32_t sw = lane & 7;
    uint8_t* abuf = sAux + (warp - 2) * 4096;     // AUXBUF: this warp's aux box (TMA-loaded operand, or the gelu output tile)
    uint64_t* abar = &aux_bar[warp - 2];
    uint32_t aphase = 0;
    auto request_aux = [&](int col0, int row0) {  // warp-uniform; boxes entirely outside the matrix are neither requested nor awaited
      if (lane == 0 && col0 < p.N && row0 < p.M) {
        mbar_expect_tx(abar, 4096);
        tma_load_2d(abuf, &tmD2, abar, col0, row0);
      }
    };
    auto load_aux = [&](uint4 (&a)[8], long long grow, int col0) {
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux + grow * p.ld_aux + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (grow < p.M && col0 + 8 * j < p.N) ? __ldg(ap + j) : make_uint4(0, 0, 0, 0);
    };
    int it = 0;
    WorkIter work(tile0, tile_stride, tiles_mn, num_k_total, p.split_k);
    for (int tile, kb0_, kb1_; work.next(tile, kb0_, kb1_); ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      const long long grow = static_cast<long long>(m_blk) * BM + row;
      const int row0 = m_blk * BM + q * 32;       // first row of this warp's box
      uint4 aux_next[8];
      if (HAS_AUX && !AUXBUF) load_aux(aux_next, grow, n_blk * BN + half * 64);
      if (HAS_AUX && AUXBUF) request_aux(n_blk * BN + half * 64, row0);
      mbar_wait(&tfull_bar[buf], tphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;

      if (EPI == EPI_F32) {
        // 8 chunks of 32 fp32 columns (128 B per row); this warp moves the 4 of its parity
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (c + 2 >= BN / 32) {               // last TMEM read of this tile by this thread
            tc_fence_before();
            if (CG2) mbar_arrive_cluster(mapa_rank0(smem_u32(&tempty_bar[buf])));
            else mbar_arrive(&tempty_bar[buf]);
          }
          if (lane == 0) tma_wait_group_read<0>();      // the previous store has finished reading the staging box
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int c0 = n_blk * BN + c * 32;
            if (c0 < p.N && row0 < p.M) {
              if (p.accumulate) tma_reduce_add_2d(&tmD, sbuf, c0, row0);
              else tma_store_2d(&tmD, sbuf, c0, row0);
            }
            tma_commit_group();
          }
        }
      } else {
        // 4 chunks of 64 bf16 columns (128 B per row); this warp converts the 2 of its parity
#pragma unroll 1
        for (int c = half; c < BN / 64; c += 2) {
          const int col0 = n_blk * BN + c * 64;
          uint4 aux[8];
          if (HAS_AUX && !AUXBUF) {
#pragma unroll
            for (int j = 0; j < 8; ++j) aux[j] = aux_next[j];
          }
          uint32_t r[64];
          {
            uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
            uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
            tmem_ld_x32(t_row + c * 64, r0);
            tmem_ld_x32(t_row + c * 64 + 32, r1);
          }
          if (HAS_AUX && !AUXBUF && c + 2 < BN / 64) load_aux(aux_next, grow, col0 + 128);     // next chunk's operand while TMEM answers
          const bool have_box = col0 < p.N && row0 < p.M;       // warp-uniform
          if (HAS_AUX && AUXBUF && have_box) {
            mbar_wait(abar, aphase);
            aphase ^= 1;
          }
          tmem_ld_wait();
          if (c + 2 >= BN / 64) {               // last TMEM read of this tile by this thread
            tc_fence_before();
            if (CG2) mbar_arrive_cluster(mapa_rank0(smem_u32(&tempty_bar[buf])));
            else mbar_arrive(&tempty_bar[buf]);
          }
          if (FP8) {
            const float sa = grow < p.M ? __ldg(p.row_scale_a + grow) : 0.f;
            const float4* sb4 = reinterpret_cast<const float4*>(p.row_scale_b + col0);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 sb = (col0 + 4 * j < p.N) ? __ldg(sb4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * sa * sb.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * sa * sb.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * sa * sb.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * sa * sb.w);
            }
          }
          if (lane == 0) tma_wait_group_read<0>();      // the previous store(s) have finished reading the staging box(es)
          __syncwarp();
          // two halves of 32 columns: operand -> math -> staging, so that only the accumulator chunk stays live in registers
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[16];
            if (HAS_AUX) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 a4;
                if (AUXBUF) a4 = have_box ? *reinterpret_cast<const uint4*>(abuf + lane * 128 + (((4 * hh + j) ^ sw) << 4)) : make_uint4(0, 0, 0, 0);
                else a4 = aux[4 * hh + j];
                const uint32_t au[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 av = unpack_bf16(au[e]);
                  float x0 = __uint_as_float(r[32 * hh + 8 * j + 2 * e]), x1 = __uint_as_float(r[32 * hh + 8 * j + 2 * e + 1]);
                  if (EPI == EPI_RESIDUAL) { x0 += av.x; x1 += av.y; }
                  else { x0 *= gelu_tanh_grad(av.x); x1 *= gelu_tanh_grad(av.y); }
                  o[4 * j + e] = pack_bf16(x0, x1);
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = pack_bf16(__uint_as_float(r[32 * hh + 2 * e]), __uint_as_float(r[32 * hh + 2 * e + 1]));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(rowp + (((4 * hh + j) ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            if (EPI == EPI_GELU) {
              // GELU of the *rounded* pre-activation so that backward (which only sees bf16 pre) is consistent
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float2 xr = unpack_bf16(o[e]);
                o[e] = pack_bf16(gelu_tanh(xr.x), gelu_tanh(xr.y));
              }
              if (AUXBUF) {                 // second staging box -> TMA store
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(abuf + lane * 128 + (((4 * hh + j) ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              } else {                      // 1-CTA kernel: straight from registers to global memory (one 128-byte line per thread)
                __nv_bfloat16* d2row = p.d2 + grow * p.ldd2 + col0 + 32 * hh;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (grow < p.M && col0 + 32 * hh + 8 * j < p.N)
                    reinterpret_cast<uint4*>(d2row)[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              }
            }
          }
          if (HAS_AUX && AUXBUF) {
            __syncwarp();                                         // every lane has consumed its operand: the box may be refilled
            if (c + 2 < BN / 64) request_aux(col0 + 128, row0);   // (the next tile's first box is requested at the top of the loop)
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (col0 < p.N && row0 < p.M) {
              tma_store_2d(&tmD, sbuf, col0, row0);
              if (EPI == EPI_GELU && AUXBUF) tma_store_2d(&tmD2, abuf, col0, row0);
            }
            tma_commit_group();
          }
        }
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync();
  if (warp == 1) {
    if (CG2) tmem_dealloc2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* ptr = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || ptr == nullptr) throw std::runtime_error("cuTensorMapEncodeTiled entry point not found");
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, leading dimension ld elements), 128B swizzle.
CUtensorMap make_tmap_2d(const void* ptr, int elem_bytes, long long rows, long long cols, long long ld, int box_cols,
                         int box_rows) {
  using Key = std::tuple<const void*, int, long long, long long, long long, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{ptr, elem_bytes, rows, cols, ld, box_cols, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : (elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ") rows=" +
                             std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld));
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

// One rasterisation group (group_m m-blocks x all n-blocks) should span about two waves of CTAs: every operand panel is then
// shared by many concurrently running CTAs (L2 hits) and re-read from HBM as rarely as possible.  With few n-blocks (wgrad of
// a narrow layer: N/256 = 8) the fixed group of 8 re-read the wide operand once per 64 tiles (ncu: 2x the ideal DRAM traffic).
static int pick_group_m(int N, int num_sms) {
  const int num_n = (N + BN - 1) / BN;
  int g = (2 * num_sms + num_n - 1) / num_n;
  if (g < GROUP_M) g = GROUP_M;
  if (g > 64) g = 64;
  return g;
}

// K-slices for the leftover tiles of the accumulating fp32 epilogue (see WorkIter); 0 = plain whole-tile schedule.
// Cost model in k-block units: waves x (slice length + one epilogue ~ 8 k-blocks); slices keep >= 16 k-blocks.
static int pick_split_k(int tiles, int workers, int num_k, int epi, int accumulate) {
  if (epi != EPI_F32 || !accumulate) return 0;
  const int left = tiles % workers;
  if (left == 0) return 1;
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 16 && (s == 1 || num_k / s >= 16); ++s) {
    const int kps = (num_k + s - 1) / s, slices = (num_k + kps - 1) / kps;
    const long long waves = (static_cast<long long>(left) * slices + workers - 1) / workers;
    const long long cost = waves * (kps + 8);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

static int g_use_cg2 = -1;      // CTA-pair kernels for the plain bf16 GEMMs (MIDGPT_GEMM_2CTA=0 disables)
static bool use_cg2() {
  if (g_use_cg2 < 0) {
    const char* e = getenv("MIDGPT_GEMM_2CTA");
    g_use_cg2 = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return g_use_cg2 == 1;
}

// MIDGPT_GEMM_AUXBUF = 0 | 1 forces the register / TMA-box path of the epilogue's second operand; default: boxes iff K <= 1024
static bool use_auxbuf(int K) {
  static const int mode = [] { const char* e = getenv("MIDGPT_GEMM_AUXBUF"); return e ? atoi(e) : -1; }();
  return mode < 0 ? K <= 1024 : mode != 0;
}

template <bool A_MN, bool B_MN, int EPI, bool FP8 = false, bool AUXB = false, bool CARRY = false>
static void launch_pair(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tD, const CUtensorMap& tD2,
                        const GemmParams& p, int num_sms, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05<A_MN, B_MN, EPI, false, FP8, true, AUXB, CARRY>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(pair): ") + cudaGetErrorString(e));
    configured = true;
  }
  const int num_tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
  int pairs = num_sms / 2;
  if (num_tiles * (p.split_k > 0 ? p.split_k : 1) < pairs && p.ag.n_seg == 0) pairs = num_tiles * (p.split_k > 0 ? p.split_k : 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);                       // (a carried gather keeps every SM's copy warps: idle pairs own no tiles)
  cfg.blockDim = dim3(NUM_THREADS + COPY_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tA, tB, tD, tD2, p);
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm pair launch: ") + cudaGetErrorString(e));
}

template <bool A_MN, bool B_MN, int EPI, bool AG = false, bool FP8 = false>
static void launch_one(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tD, const CUtensorMap& tD2,
                       const GemmParams& p, int num_sms, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05<A_MN, B_MN, EPI, AG, FP8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int grid = num_tiles * (p.split_k > 0 ? p.split_k : 1) < num_sms ? num_tiles * (p.split_k > 0 ? p.split_k : 1) : num_sms;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tA, tB, tD, tD2, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm launch: ") + cudaGetErrorString(e));
}

// a: [M,K] row-major (a_mn=0) or [K,M] row-major (a_mn=1);  b: [N,K] (b_mn=0) or [K,N] (b_mn=1)
// d: [M,N] bf16 (fp32 for EPI_F32); d2: second bf16 output for EPI_GELU; aux: [M,N] bf16 for RESIDUAL / DGELU.
void gemm_launch_ag(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d,
                    long long ldd, void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K,
                    int epi, int accumulate, int num_sms, cudaStream_t stream, const GatherParams* ag);

void gemm_launch(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d,
                 long long ldd, void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K,
                 int epi, int accumulate, int num_sms, cudaStream_t stream) {
  gemm_launch_ag(a, lda, a_mn, b, ldb, b_mn, d, ldd, d2, ldd2, aux, ld_aux, M, N, K, epi, accumulate, num_sms, stream, nullptr);
}

void ag_pull(const long long* lp_ptrs, void* local, long long shard_elems, int rank, int world, unsigned* flags, int n_ctas, cudaStream_t s);

// Host entry for all-gather ⊕ GEMM: any GEMM plus the description of the FSDP unit whose gather it carries.  If b lies inside the
// unit buffer (a weight of this very unit) the kernel's TMA producers synchronise with the copy warps segment by segment and the
// pull starts at b; otherwise the GEMM just hides the pull.  GEMMs that cannot take the pair kernel (M < 256) fall back to the
// stand-alone NVLink pull kernel followed by the plain GEMM.
void gemm_launch_ag_raw(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d, long long ldd, void* d2,
                        long long ldd2, const void* aux, long long ld_aux, int M, int N, int K, int epi, int accumulate, int num_sms,
                        cudaStream_t stream, const long long* peer_ptrs, void* local, unsigned int* flags, long long chunk_elems,
                        unsigned int target, int world, int rank) {
  const bool has_variant = (!a_mn && !b_mn && (epi == EPI_STORE || epi == EPI_GELU)) || (a_mn && b_mn && epi == EPI_F32) || (!a_mn && b_mn && epi == EPI_DGELU);
  if (!(use_cg2() && M >= 2 * BM && has_variant)) {
    ag_pull(peer_ptrs, local, chunk_elems, rank, world, flags + (chunk_elems * world + SEG_ELEMS - 1) / SEG_ELEMS, 32, stream);
    return gemm_launch_ag(a, lda, a_mn, b, ldb, b_mn, d, ldd, d2, ldd2, aux, ld_aux, M, N, K, epi, accumulate, num_sms, stream, nullptr);
  }
  GatherParams ag{};
  for (int i = 0; i < 8; ++i) ag.peer[i] = i < world ? reinterpret_cast<void*>(peer_ptrs[i]) : nullptr;
  ag.local = static_cast<__nv_bfloat16*>(local);
  ag.flags = flags; ag.chunk_elems = chunk_elems; ag.target = target; ag.world = world; ag.rank = rank;
  const long long unit_elems = chunk_elems * world;
  ag.n_seg = static_cast<int>((unit_elems + SEG_ELEMS - 1) / SEG_ELEMS);
  const long long off = static_cast<const __nv_bfloat16*>(b) - static_cast<const __nv_bfloat16*>(local);
  const bool b_in_unit = off >= 0 && off + static_cast<long long>(N) * K <= unit_elems && ldb == (b_mn ? N : K);
  ag.b_offset = b_in_unit ? off : -1;
  ag.first_seg = b_in_unit ? off / SEG_ELEMS : 0;
  gemm_launch_ag(a, lda, a_mn, b, ldb, b_mn, d, ldd, d2, ldd2, aux, ld_aux, M, N, K, epi, accumulate, num_sms, stream, &ag);
}

void gemm_launch_ag(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d,
                    long long ldd, void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K,
                    int epi, int accumulate, int num_sms, cudaStream_t stream, const GatherParams* ag) {
  CUtensorMap tA = a_mn ? make_tmap_2d(a, 2, K, M, lda, 64, BK) : make_tmap_2d(a, 2, M, K, lda, BK, BM);
  CUtensorMap tB = b_mn ? make_tmap_2d(b, 2, K, N, ldb, 64, BK) : make_tmap_2d(b, 2, N, K, ldb, BK, BN);
  CUtensorMap tD = (epi == EPI_F32) ? make_tmap_2d(d, 4, M, N, ldd, 32, 32) : make_tmap_2d(d, 2, M, N, ldd, 64, 32);   // one box per epilogue warp
  // second [M, N] bf16 tile of the epilogue: gelu output (stored) or residual / pre-activation operand (loaded); pair kernels only
  CUtensorMap tD2 = (epi == EPI_GELU) ? make_tmap_2d(d2, 2, M, N, ldd2, 64, 32)
                    : ((epi == EPI_RESIDUAL || epi == EPI_DGELU) ? make_tmap_2d(aux, 2, M, N, ld_aux, 64, 32) : tD);
  GemmParams p{reinterpret_cast<const __nv_bfloat16*>(aux), ld_aux, M, N, K, accumulate, GatherParams{}, pick_group_m(N, num_sms), 1, static_cast<__nv_bfloat16*>(d2), ldd2, nullptr, nullptr};
  if (ag) {
    p.ag = *ag;
    // weight rows arrive in address order: with a K-major B that is part of the gathered unit, sweep all m-blocks of one n-block
    // before the next n-block, so the first wave needs only the first 256 weight rows and later waves consume rows long after
    // they landed (A stays L2-resident)
    if (ag->b_offset >= 0 && !b_mn) p.group_m = (M + 2 * BM - 1) / (2 * BM);
    if (!(use_cg2() && M >= 2 * BM)) throw std::runtime_error("all-gather ⊕ GEMM needs the pair kernel (M >= 256)");
  }
  if (use_cg2() && M >= 2 * BM) {
    // pair kernels: each CTA loads half of the B tile -> K-major B needs a 128-row box (MN-major boxes are 64 wide anyway)
    CUtensorMap tB2 = b_mn ? tB : make_tmap_2d(b, 2, N, K, ldb, BK, BN / 2);
    if (!(ag && ag->b_offset >= 0 && !b_mn)) p.group_m = pick_group_m(N, num_sms / 2);
    p.split_k = pick_split_k(((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN), num_sms / 2, (K + BK - 1) / BK, epi, accumulate);
    // second [M, N] tile of the epilogue through TMA boxes (5 operand stages) or through registers (6 stages): the boxes win
    // when the k-loop is short and the epilogue sets the pace (K = 768), the deeper ring when it is long
    if (ag) {      // instantiations that can carry a gather: forward store (QKV / LM head), MLP-in (gelu), wgrad, the first dgrad of a block
      if (!a_mn && !b_mn && epi == EPI_STORE) return launch_pair<false, false, EPI_STORE, false, false, true>(tA, tB2, tD, tD2, p, num_sms, stream);
      if (!a_mn && !b_mn && epi == EPI_GELU) return launch_pair<false, false, EPI_GELU, false, false, true>(tA, tB2, tD, tD2, p, num_sms, stream);
      if (a_mn && b_mn && epi == EPI_F32) return launch_pair<true, true, EPI_F32, false, false, true>(tA, tB2, tD, tD2, p, num_sms, stream);
      if (!a_mn && b_mn && epi == EPI_DGELU) return launch_pair<false, true, EPI_DGELU, false, false, true>(tA, tB2, tD, tD2, p, num_sms, stream);
      throw std::runtime_error("all-gather ⊕ GEMM: this (layout, epilogue) has no carrying instantiation");
    }
    const bool auxb = use_auxbuf(K);
#define MG_PAIR(AM, BM_, E) \
    if (a_mn == AM && b_mn == BM_ && epi == E) return launch_pair<AM, BM_, E>(tA, tB2, tD, tD2, p, num_sms, stream);
#define MG_PAIRX(AM, BM_, E) \
    if (a_mn == AM && b_mn == BM_ && epi == E) { \
      if (auxb) return launch_pair<AM, BM_, E, false, true>(tA, tB2, tD, tD2, p, num_sms, stream); \
      return launch_pair<AM, BM_, E>(tA, tB2, tD, tD2, p, num_sms, stream); }
    MG_PAIR(false, false, EPI_STORE)
    MG_PAIRX(false, false, EPI_RESIDUAL)
    MG_PAIRX(false, false, EPI_GELU)
    MG_PAIR(false, false, EPI_F32)
    MG_PAIR(false, true, EPI_STORE)
    MG_PAIRX(false, true, EPI_DGELU)
    MG_PAIRX(false, true, EPI_RESIDUAL)
    MG_PAIR(true, true, EPI_F32)
    MG_PAIR(true, true, EPI_STORE)
    MG_PAIR(true, false, EPI_F32)
#undef MG_PAIR
#undef MG_PAIRX
  }
  p.split_k = pick_split_k(((M + BM - 1) / BM) * ((N + BN - 1) / BN), num_sms, (K + BK - 1) / BK, epi, accumulate);
#define MG_CASE(AM, BM_, E) \
  if (a_mn == AM && b_mn == BM_ && epi == E) return launch_one<AM, BM_, E>(tA, tB, tD, tD2, p, num_sms, stream);
  MG_CASE(false, false, EPI_STORE)
  MG_CASE(false, false, EPI_RESIDUAL)
  MG_CASE(false, false, EPI_GELU)
  MG_CASE(false, false, EPI_F32)
  MG_CASE(false, true, EPI_STORE)
  MG_CASE(false, true, EPI_DGELU)
  MG_CASE(false, true, EPI_RESIDUAL)
  MG_CASE(true, true, EPI_F32)
  MG_CASE(true, true, EPI_STORE)
  MG_CASE(true, false, EPI_F32)
#undef MG_CASE
  throw std::runtime_error("gemm_launch: unsupported (a_mn, b_mn, epilogue) combination");
}

// FP8 forward GEMM: a [M,K] and b [N,K] hold e4m3 bytes with per-row de-quantisation scales sa [M], sb [N]; d (bf16) = (a·bᵀ)∘(sa⊗sb).
void gemm_fp8_launch(const void* a, long long lda, const void* b, long long ldb, const float* sa, const float* sb, void* d, long long ldd,
                     void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K, int epi, int num_sms,
                     cudaStream_t stream) {
  CUtensorMap tA = make_tmap_2d(a, 1, M, K, lda, 2 * BK, BM);
  CUtensorMap tB = make_tmap_2d(b, 1, N, K, ldb, 2 * BK, BN);
  CUtensorMap tD = make_tmap_2d(d, 2, M, N, ldd, 64, 32);
  CUtensorMap tD2 = (epi == EPI_GELU) ? make_tmap_2d(d2, 2, M, N, ldd2, 64, 32)
                    : (epi == EPI_RESIDUAL ? make_tmap_2d(aux, 2, M, N, ld_aux, 64, 32) : tD);
  GemmParams p{reinterpret_cast<const __nv_bfloat16*>(aux), ld_aux, M, N, K, 0, GatherParams{}, pick_group_m(N, num_sms), 1, static_cast<__nv_bfloat16*>(d2), ldd2, sa, sb};
  if (use_cg2() && M >= 2 * BM) {
    CUtensorMap tB2 = make_tmap_2d(b, 1, N, K, ldb, 2 * BK, BN / 2);
    p.group_m = pick_group_m(N, num_sms / 2);
    if (epi == EPI_STORE) return launch_pair<false, false, EPI_STORE, true>(tA, tB2, tD, tD2, p, num_sms, stream);
    if (epi == EPI_GELU) return launch_pair<false, false, EPI_GELU, true>(tA, tB2, tD, tD2, p, num_sms, stream);
    if (epi == EPI_RESIDUAL) return launch_pair<false, false, EPI_RESIDUAL, true>(tA, tB2, tD, tD2, p, num_sms, stream);
  }
  if (epi == EPI_STORE) return launch_one<false, false, EPI_STORE, false, true>(tA, tB, tD, tD2, p, num_sms, stream);
  if (epi == EPI_GELU) return launch_one<false, false, EPI_GELU, false, true>(tA, tB, tD, tD2, p, num_sms, stream);
  if (epi == EPI_RESIDUAL) return launch_one<false, false, EPI_RESIDUAL, false, true>(tA, tB, tD, tD2, p, num_sms, stream);
  throw std::runtime_error("gemm_fp8: unsupported epilogue");
}

// The work plan of a GEMM exactly as the kernel will walk it (pick_split_k + WorkIter run on the host): one row
// (worker, tile, kb0, kb1) per piece.  Used by the CPU tests to check that every (tile, k-block) is covered exactly once.
std::vector<int> gemm_work_plan(int M, int N, int K, int epi, int accumulate, int num_sms, bool pairs) {
  const int tiles = pairs ? ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN) : ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int num_k = (K + BK - 1) / BK;
  const int max_workers = pairs ? num_sms / 2 : num_sms;
  const int split = pick_split_k(tiles, max_workers, num_k, epi, accumulate);
  const int items = tiles * (split > 0 ? split : 1);
  const int workers = items < max_workers ? items : max_workers;      // same rule as the launchers' grid size
  std::vector<int> out;
  for (int w = 0; w < workers; ++w) {
    WorkIter it(w, workers, tiles, num_k, split);
    for (int tile, kb0, kb1; it.next(tile, kb0, kb1);) {
      out.push_back(w); out.push_back(tile); out.push_back(kb0); out.push_back(kb1);
    }
  }
  return out;
}

}  // namespace mg
