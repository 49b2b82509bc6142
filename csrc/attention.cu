// Causal flash attention for sm_100a on tcgen05 / TMEM / TMA — forward and backward.
//
// Inputs are the already QK-LayerNorm'ed + RoPE'd q, k and the raw v, as [B, T, H, C] views with an
// arbitrary token stride (so q/k come from the packed [M, 2D] buffer and v straight from the
// projection output — no copies).  Scores are scaled by 1/sqrt(C) inside an fp32 softmax
// (reference src/model.py:71-77).  C ∈ {64, 128}.
//
// Forward  (one CTA = one (batch, head, 128-query tile); 2 CTAs co-reside per SM and fill each other's bubbles)
//   warps 0-3  softmax: thread = one query row; S row read from TMEM, online max/sum with lazy (threshold)
//              rescaling of the O accumulator *in TMEM*, P written back as packed bf16 over its own S columns (tcgen05.st)
//   warp 4     TMA producer (Q once, K/V double buffered)
//   warp 5     tcgen05.mma issuer:  S = Q·Kᵀ (SS, fp32 in TMEM),  O += P·V (P from tensor memory, V MN-major smem)
// Backward: work item = one (batch, head, 128-key tile), looping over the query tiles at or below the diagonal
//   S = Q·Kᵀ, dP = dO·Vᵀ  →  P = exp2(S·c − lse), dS = P∘(dP − δ)/√C  →  dV += Pᵀ·dO, dK += dSᵀ·Q (TMEM-resident
//   accumulators), dQ = dS·K per tile in TMEM, reduced into an fp32 buffer with TMA reduce-add.
//   v3 (head_dim 128, T % 64 == 0; see the block comment at attn_bwd3_kernel): transposed scores, Pᵀ/dSᵀ as tensor-memory
//      A operands, 64-row query tiles, dedicated output warps, persistent CTAs with dynamic work distribution.
//   v2 (attn_bwd_kernel; head_dim 64, ragged T, MIDGPT_ATTN_BWD=v2): 128-row query tiles, one CTA per item, dQ aliases dP.
// In every kernel the MMA-issuing warp runs converged with one elected lane per MMA group and hoisted descriptors
// (common.cuh: desc_shift) -- otherwise the issuing thread, not the tensor core, sets the pace.
#include "common.cuh"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace mg {

typedef CUresult (*PFN_encodeTiled4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_tmap_bthc(const void* ptr, int B, int T, int H, int C, long long row_stride, int box_rows) {
  static PFN_encodeTiled4 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled entry point not found");
    fn = reinterpret_cast<PFN_encodeTiled4>(p);
  }
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)row_stride * 2, (cuuint64_t)row_stride * 2 * T};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("attention: cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return m;
}

// fp32 dQ accumulator [B, T, H*C] as a 3-D map so that query-tile rows past T are clipped instead of spilling into the next batch
static CUtensorMap make_tmap_dq(const void* ptr, int B, int T, int D, int box_rows = 128) {
  cudaDriverEntryPointQueryResult q;
  void* p = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
    throw std::runtime_error("cuTensorMapEncodeTiled entry point not found");
  PFN_encodeTiled4 fn = reinterpret_cast<PFN_encodeTiled4>(p);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)D * 4 * T};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("attention: dq tensor map failed: " + std::to_string((int)r));
  return m;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Write 8 bf16 (one 16-byte chunk) of row r, logical column chunk cc (8-element units) into a K-major,
// 128B-swizzled tile made of 64-column blocks of `rows` rows each.
MG_DEVICE void st_swizzled_chunk(uint8_t* tile, int rows, int r, int cc, uint4 v) {
  const int blk = cc >> 3, c = cc & 7;
  *reinterpret_cast<uint4*>(tile + blk * (rows * 128) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

// ================================================================================================
// Forward
// ================================================================================================
template <int C>
struct FwdCfg {
  static constexpr int BQ = 128;
  static constexpr int BKV = 64;
  static constexpr int Q_BYTES = BQ * C * 2;
  static constexpr int K_BYTES = BKV * C * 2;
  static constexpr int V_BYTES = BKV * C * 2;
  static constexpr int SMEM = Q_BYTES + 2 * K_BYTES + 2 * V_BYTES + 256;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, O_COL = 2 * BKV;      // two S buffers (QK of tile j+1 overlaps the softmax of tile j), then O;
                                                        // P_j (bf16, two per column) overwrites the first BKV/2 columns of S buffer j & 1
  static_assert(2 * BKV + C <= TMEM_COLS, "TMEM budget");
  static_assert(BKV == 64, "P is packed over an S row that must already sit in registers (KEEP path)");
};

// DROP: attention-probability dropout (reference src/model.py:78) on P in registers, after the row sum l has been taken
// (dropout acts on the *normalised* probabilities: O = (mask ∘ P / (1-p)) · V, the softmax statistics are untouched).
template <int C, bool DROP>
__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out,
                int T, int H, float scale, uint32_t thr16, float inv_keep, uint2 seed, int eager) {
  // eager (MIDGPT_ATTN_DEBUG & 8, for compute-sanitizer synccheck only): await every phase of pv_done.  By design the softmax warps
  // await that barrier lazily, so phases complete that nobody waits for -- which synccheck reports as "missing wait" and aborts.
  using Cfg = FwdCfg<C>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B operands need 1024-byte aligned tiles
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + 2 * Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * Cfg::V_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // 2 (one per S buffer)
  uint64_t* p_full = bars + 11;       // 2 (128 arrivals each; one per S/P buffer -- a softmax warp may run a tile ahead of the others, and its
                                      //    arrival for tile j+1 must not be counted towards tile j: tile j+2 cannot start before p_full(j) completed)
  uint64_t* pv_done = bars + 13;      // 1: completes once per P·V (awaited LAZILY: only when a row maximum forces a rescale of O)
  uint64_t* o_done = bars + 14;       // 1: the LAST P·V has retired.  The epilogue must not use a parity wait on pv_done: a warp that
                                      //    never rescaled has skipped phases, and if it runs two tiles ahead of the slowest warp the
                                      //    phase it asks for aliases one that completed long ago (it then normalises an O that lacks the
                                      //    last two key tiles -- an intermittent few-percent error on 32 rows, caught by the row-wise test)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;          // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BQ;
  const int n_kv = min((q0 + BQ + BKV - 1) / BKV, (T + BKV - 1) / BKV);   // causal: keys < q0 + BQ

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&p_full[0], 128); mbar_init(&p_full[1], 128); mbar_init(pv_done, 1); mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sQ + cb * (BQ * 128), &tmQ, q_full, cb * 64, h, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::K_BYTES);
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sK + st * Cfg::K_BYTES + cb * (BKV * 128), &tmK, &k_full[st], cb * 64, h, j * BKV, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], Cfg::V_BYTES);
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sV + st * Cfg::V_BYTES + cb * (BKV * 128), &tmV, &v_full[st], cb * 64, h, j * BKV, b);
      }
    }
  } else if (warp == 5) {
    // ------------------------------- MMA issuer -------------------------------
    // whole warp converged (descriptors stay in uniform registers), one elected lane per group of tcgen05 instructions
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, C, 0, 1);
      // descriptors are built once; per MMA only a 64-bit add of a compile-time offset (see desc_shift)
      const uint64_t q_desc = make_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = make_smem_desc(smem_u32(sK), 16, 1024), v_desc0 = make_smem_desc(smem_u32(sV), BKV * 128, 1024);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t k_desc = desc_shift(k_desc0, st * Cfg::K_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            const uint32_t blk = (k * 16) / 64, in = (k * 16) % 64;
            umma_bf16_ss(tmem + Cfg::S_COL + st * BKV, desc_shift(q_desc, blk * (BQ * 128) + in * 2),
                         desc_shift(k_desc, blk * (BKV * 128) + in * 2), idesc_s, k != 0);
          }
          umma_commit(&k_empty[st]);
          umma_commit(&s_full[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (n_kv > 1) issue_qk(1);           // S is double buffered: the next QKᵀ never waits for the softmax
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);   // P_j is in tensor memory (over the S_j columns every thread has already read)
        tc_fence_after();
        const int st = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t v_desc = desc_shift(v_desc0, st * Cfg::V_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)       // A = P_j straight from TMEM: 16 keys = 8 packed columns per k-step
            umma_bf16_ts(tmem + Cfg::O_COL, tmem + Cfg::S_COL + st * BKV + k * 8, desc_shift(v_desc, k * 2048), idesc_o, (j | k) != 0);
          umma_commit(&v_empty[st]);
          if (j + 1 < n_kv) umma_commit(pv_done);   // (the last P·V is awaited through o_done only)
          else umma_commit(o_done);
        }
        __syncwarp();
        if (j + 2 < n_kv) issue_qk(j + 2);   // reuses S buffer j&1, which the softmax warps released at p_full(j)
      }
    }
  } else {
    // ------------------------------- softmax warps (thread = query row) -------------------------------
    const int row = threadIdx.x;                 // 0..127, also the TMEM lane
    const int q_idx = q0 + row;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = scale * kLog2e;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      const uint32_t s_col = Cfg::S_COL + (j & 1) * BKV;
      const bool need_mask = (kv0 + BKV - 1 > q0);          // tile touches or crosses the diagonal
      constexpr int NCH = BKV / 32;
      constexpr bool KEEP = (BKV == 64);                     // S row stays in registers between the two passes
      uint32_t sreg[KEEP ? NCH : 1][32];
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t tmp[32];
        uint32_t(&r)[32] = KEEP ? sreg[KEEP ? ch : 0] : tmp;
        tmem_ld_x32(t_lane + s_col + ch * 32, r);
        if (!KEEP) {
          tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (kv0 + ch * 32 + e > q_idx) ? -INFINITY : __uint_as_float(r[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
          }
        }
      }
      if (KEEP) {
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (kv0 + ch * 32 + e > q_idx) ? -INFINITY : __uint_as_float(sreg[KEEP ? ch : 0][e]));
        } else {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(sreg[KEEP ? ch : 0][e]));
        }
      }
      mx *= sl2;                                             // sl2 > 0: max commutes with the scaling
      if (j > 0) {
        const bool need = mx > m_used + 8.f;       // lazy rescale: only when the max grew by > 2^8
        if (__any_sync(0xffffffffu, need) || eager) {
          mbar_wait(pv_done, (j - 1) & 1);         // O_{j-1} complete (P_{j-1} lives in the other S buffer: nothing else to wait for)
          tc_fence_after();
          const float alpha = need ? exp2f(m_used - mx) : 1.f;
          if (need) m_used = mx;
          l *= alpha;
#pragma unroll
          for (int ch = 0; ch < C / 16; ++ch) {
            uint32_t r[16];
            tmem_ld_x16(t_lane + Cfg::O_COL + ch * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            tmem_st_x16(t_lane + Cfg::O_COL + ch * 16, r);
          }
          tmem_st_wait();
        }
      } else {
        m_used = mx;
      }
      // pass 2: P = exp2(s*c - m_used) -> bf16, packed two per column, back into this thread's lane of S buffer j & 1
      //         (tensor-memory A operand of P·V; the whole S row is already in registers)
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t tmp[32];
        uint32_t(&r)[32] = KEEP ? sreg[KEEP ? ch : 0] : tmp;
        if (!KEEP) {
          tmem_ld_x32(t_lane + s_col + ch * 32, r);
          tmem_ld_wait();
        }
        float p[32];
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float pe = exp2f(fmaf(__uint_as_float(r[e]), sl2, -m_used));
            if (kv0 + ch * 32 + e > q_idx) pe = 0.f;
            p[e] = pe;
            l += pe;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            p[e] = exp2f(fmaf(__uint_as_float(r[e]), sl2, -m_used));
            l += p[e];
          }
        }
        if (DROP) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const uint32_t keep = dropout_keep8(seed, b * H + h, q_idx, (kv0 + ch * 32) / 8 + g8, thr16);
#pragma unroll
            for (int i = 0; i < 8; ++i) p[8 * g8 + i] = (keep >> i) & 1u ? p[8 * g8 + i] * inv_keep : 0.f;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(p[2 * e], p[2 * e + 1]);
        tmem_st_x16(t_lane + s_col + ch * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }
    // epilogue: O / l -> bf16 -> global ; lse
    mbar_wait(o_done, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const bool valid = q_idx < T;
    __nv_bfloat16* orow = out + ((static_cast<long long>(b) * T + q_idx) * H + h) * C;
#pragma unroll
    for (int ch = 0; ch < C / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_x32(t_lane + Cfg::O_COL + ch * 32, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v4;
          v4.x = pack_bf16(__uint_as_float(r[8 * g]) * inv_l, __uint_as_float(r[8 * g + 1]) * inv_l);
          v4.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * inv_l, __uint_as_float(r[8 * g + 3]) * inv_l);
          v4.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * inv_l, __uint_as_float(r[8 * g + 5]) * inv_l);
          v4.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * inv_l, __uint_as_float(r[8 * g + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + ch * 32 + g * 8) = v4;
        }
      }
    }
    if (valid) lse_out[(static_cast<long long>(b) * H + h) * T + q_idx] = (m_used + log2f(l)) * kLn2;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ================================================================================================
// Backward
// ================================================================================================
// delta[b,h,t] = sum_c dO·O.  C/8 lanes per (token, head) vector with 16-byte loads; every lane keeps U vectors in
// flight (2·U loads issued before any arithmetic) so the kernel runs at HBM speed instead of load latency.
template <int C, int U>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                            float* __restrict__ delta, const float* __restrict__ lse, float scale,
                                                            size_t plane, unsigned nvec, int T, int H) {
  // delta has three planes of `plane` = B*H*T floats: raw delta (v2 kernel) | delta*scale | lse*log2(e) (v3 kernel)
  constexpr int LPV = C / 8, VPW = 32 / LPV;
  const int lane = threadIdx.x & 31, sub = lane % LPV, slot = lane / LPV;
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned long long base = static_cast<unsigned long long>(warp0) * (VPW * U); base < nvec;
       base += static_cast<unsigned long long>(nwarps) * (VPW * U)) {
    uint4 a[U], d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long vec = base + u * VPW + slot;
      const size_t v = vec < nvec ? vec : nvec - 1;
      a[u] = __ldg(reinterpret_cast<const uint4*>(o + v * C) + sub);
      d[u] = __ldg(reinterpret_cast<const uint4*>(dout + v * C) + sub);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t au[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, du[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16(au[e]), y = unpack_bf16(du[e]);
        acc += x.x * y.x + x.y * y.y;
      }
#pragma unroll
      for (int off = LPV / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      const unsigned long long vec = base + u * VPW + slot;
      if (sub == 0 && vec < nvec) {
        const unsigned v = static_cast<unsigned>(vec);
        const unsigned bt = v / H, hh = v - bt * H, bb = bt / T, t = bt - bb * T;
        const size_t di = (static_cast<size_t>(bb) * H + hh) * T + t;
        delta[di] = acc;
        delta[plane + di] = acc * scale;
        delta[2 * plane + di] = lse[di] * 1.4426950408889634f;
      }
    }
  }
}

// dq (bf16, strided) = dq_accum (fp32 [M, D])
__global__ void dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, long long M, int D, long long ld) {
  const int nv = D >> 2;
  const long long total = M * nv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / nv;
    const int j = i - m * nv;
    float4 v = __ldg(reinterpret_cast<const float4*>(acc + m * D) + j);
    *reinterpret_cast<uint2*>(dq + m * ld + j * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

template <int C>
struct BwdCfg {
  static constexpr int TILE = 128 * C * 2;             // one [128 x C] bf16 operand tile
  static constexpr int PS = 128 * 128 * 2;             // P, then dS, then the dQ staging area (one buffer, reused in turn)
  static constexpr int SMEM = 6 * TILE + PS + 256;     // K, V, 2 x Q, 2 x dO, P/dS
  static constexpr int S_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 256 + C;   // dQ aliases dP
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 320;                  // 8 compute warps + TMA warp + MMA warp
};

// Backward.  One CTA = one (batch, head, 128-key tile); loop over query tiles i >= j.
//   compute warps 0-7: thread = query row (TMEM lane) x column half (warps 0-3: keys 0-63, warps 4-7: keys 64-127)
//   warp 8: TMA producer (K,V once; Q_i/dO_i double buffered)       warp 9: tcgen05.mma issuer
// Tensor-pipe order per tile:  dV(i) dK(i) dQ(i) | S(i+1) dP(i+1)  — S(i+1) is issued before dQ(i) is drained
// (dQ aliases dP, not S), so the P phase of the next tile never waits for the tensor core.
// DROP: the forward's keep mask is regenerated from the same Philox counters: dV += (mask ∘ P / (1-p))ᵀ · dO and
// dS = P ∘ (mask / (1-p) ∘ dP − δ) / √C, with δ = Σ dO ∘ O unchanged (O already contains the dropped probabilities).
template <int C, bool DROP>
__global__ void __launch_bounds__(BwdCfg<C>::THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const float* __restrict__ lse, const float* __restrict__ delta,
                __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, long long dkv_row_stride, int T, int H,
                float scale, int dbg, uint32_t thr16, float inv_keep, uint2 seed) {
  // dbg (MIDGPT_ATTN_DEBUG, profiling experiments only; results are wrong when set): 1 = skip the dQ reduce-add,
  // 2 = skip the exponentials, 4 = skip the dQ staging stores as well
  using Cfg = BwdCfg<C>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B operands need 1024-byte aligned tiles
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::TILE;
  uint8_t* sQ = sV + Cfg::TILE;                  // 2 stages
  uint8_t* sdO = sQ + 2 * Cfg::TILE;             // 2 stages
  uint8_t* sPS = sdO + 2 * Cfg::TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPS + Cfg::PS);
  uint64_t* kv_full = bars;          // K and V landed
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* do_full = bars + 3;      // [2]
  uint64_t* qdo_empty = bars + 5;    // [2] Q_i / dO_i no longer read by any MMA
  uint64_t* s_full = bars + 7;
  uint64_t* dp_full = bars + 8;
  uint64_t* p_ready = bars + 9;      // 256 arrivals: P in smem, S fully read
  uint64_t* dv_done = bars + 10;     // dV MMA retired: P buffer may be overwritten with dS
  uint64_t* ds_ready = bars + 11;    // 256 arrivals: dS in smem, dP fully read
  uint64_t* dq_full = bars + 12;     // dK / dQ MMAs retired
  uint64_t* dq_drained = bars + 13;  // 256 arrivals: dQ read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x;                     // key tile (small index = most work, scheduled first)
  const int h = blockIdx.y, b = blockIdx.z;
  const int kv0 = jt * 128;
  const int nq = (T + 127) / 128;
  const int n_it = nq - jt;                      // query tiles jt .. nq-1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&do_full[s], 1); mbar_init(&qdo_empty[s], 2); }   // MMA commit + dQ drain
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_ready, 256); mbar_init(dv_done, 1); mbar_init(ds_ready, 256);
    mbar_init(dq_full, 1); mbar_init(dq_drained, 256);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::TILE);
#pragma unroll
      for (int cb = 0; cb < C / 64; ++cb) {
        tma_load_4d(sK + cb * (128 * 128), &tmK, kv_full, cb * 64, h, kv0, b);
        tma_load_4d(sV + cb * (128 * 128), &tmV, kv_full, cb * 64, h, kv0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const int q0 = (jt + it) * 128;
        mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], Cfg::TILE);
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sQ + st * Cfg::TILE + cb * (128 * 128), &tmQ, &q_full[st], cb * 64, h, q0, b);
        mbar_expect_tx(&do_full[st], Cfg::TILE);
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sdO + st * Cfg::TILE + cb * (128 * 128), &tmdO, &do_full[st], cb * 64, h, q0, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------- MMA issuer -------------------------------
    // whole warp converged (descriptors stay in uniform registers), one elected lane per group of tcgen05 instructions
    {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);   // S, dP : both operands K-major, N = 128
      constexpr uint32_t idesc_mm = make_idesc_bf16(128, C, 1, 1);     // dV, dK: both MN-major, N = C
      constexpr uint32_t idesc_km = make_idesc_bf16(128, C, 0, 1);     // dQ    : A K-major (dS), B MN-major (K)
      // descriptors are built once; per MMA only a 64-bit add of a compile-time offset (see desc_shift)
      const uint64_t k_k = make_smem_desc(smem_u32(sK), 16, 1024), v_k = make_smem_desc(smem_u32(sV), 16, 1024);       // K-major views
      const uint64_t q_k0 = make_smem_desc(smem_u32(sQ), 16, 1024), do_k0 = make_smem_desc(smem_u32(sdO), 16, 1024);
      const uint64_t k_mn = make_smem_desc(smem_u32(sK), 128 * 128, 1024), ps_mn = make_smem_desc(smem_u32(sPS), 128 * 128, 1024);   // MN-major views
      const uint64_t q_mn0 = make_smem_desc(smem_u32(sQ), 128 * 128, 1024), do_mn0 = make_smem_desc(smem_u32(sdO), 128 * 128, 1024);
      const uint64_t ps_k = make_smem_desc(smem_u32(sPS), 16, 1024);
      auto issue_s = [&](int it) {
        const int st = it & 1;
        mbar_wait(&q_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t q_k = desc_shift(q_k0, st * Cfg::TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            const uint32_t off = ((k * 16) / 64) * (128 * 128) + ((k * 16) % 64) * 2;
            umma_bf16_ss(tmem + Cfg::S_COL, desc_shift(q_k, off), desc_shift(k_k, off), idesc_kk, k != 0);
          }
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int it) {
        const int st = it & 1;
        mbar_wait(&do_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t do_k = desc_shift(do_k0, st * Cfg::TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            const uint32_t off = ((k * 16) / 64) * (128 * 128) + ((k * 16) % 64) * 2;
            umma_bf16_ss(tmem + Cfg::DP_COL, desc_shift(do_k, off), desc_shift(v_k, off), idesc_kk, k != 0);
          }
          umma_commit(dp_full);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int it = 0; it < n_it; ++it) {
        const uint32_t ph = it & 1;
        const int st = it & 1;
        const uint64_t q_mn = desc_shift(q_mn0, st * Cfg::TILE), do_mn = desc_shift(do_mn0, st * Cfg::TILE);
        // dV += Pᵀ · dO        (A = P as MN-major [k = q rows][m = kv], B = dO as MN-major [k = q rows][n = c])
        mbar_wait(p_ready, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)
            umma_bf16_ss(tmem + Cfg::DV_COL, desc_shift(ps_mn, k * 2048), desc_shift(do_mn, k * 2048), idesc_mm, (it | k) != 0);
          umma_commit(dv_done);
        }
        __syncwarp();
        // dK += dSᵀ · Q ; dQ = dS · K   (dS overwrote P in the same buffer)
        mbar_wait(ds_ready, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)
            umma_bf16_ss(tmem + Cfg::DK_COL, desc_shift(ps_mn, k * 2048), desc_shift(q_mn, k * 2048), idesc_mm, (it | k) != 0);
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k) {
            const uint32_t off = ((k * 16) / 64) * (128 * 128) + ((k * 16) % 64) * 2;
            umma_bf16_ss(tmem + Cfg::DP_COL, desc_shift(ps_k, off), desc_shift(k_mn, k * 2048), idesc_km, k != 0);
          }
          umma_commit(&qdo_empty[st]);
          umma_commit(dq_full);
        }
        __syncwarp();
        if (it + 1 < n_it) {
          issue_s(it + 1);                       // S columns were released at p_ready(it)
          mbar_wait(dq_drained, ph);             // dQ(it) read out -> dP columns reusable
          tc_fence_after();
          issue_dp(it + 1);
        }
      }
    }
  } else {
    // ------------------------------- compute warps -------------------------------
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;             // query row inside the tile == TMEM lane
    const int cbase = half * 64;                     // this thread's 64 key columns
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = scale * kLog2e;
    const bool issuer = (threadIdx.x == 0);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = it & 1;
      const int q0 = (jt + it) * 128;
      const int q_idx = q0 + row;
      const bool qvalid = q_idx < T;
      const long long sidx = (static_cast<long long>(b) * H + h) * T + (qvalid ? q_idx : 0);
      const float lse2 = lse[sidx] * kLog2e;
      const float dsc = delta[sidx] * scale;
      const bool masked = (it == 0) || (q0 + 128 > T) || (kv0 + 128 > T);
      float p[64];
      // ---- P = exp2(S*c - lse)
      mbar_wait(s_full, ph);
      tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(t_lane + Cfg::S_COL + cbase, r0);
        tmem_ld_x32(t_lane + Cfg::S_COL + cbase + 32, r1);
        tmem_ld_wait();
        if (dbg & 2) {
#pragma unroll
          for (int e = 0; e < 32; ++e) { p[e] = __uint_as_float(r0[e]); p[32 + e] = __uint_as_float(r1[e]); }
        } else if (!masked) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            p[e] = exp2f(fmaf(__uint_as_float(r0[e]), sl2, -lse2));
            p[32 + e] = exp2f(fmaf(__uint_as_float(r1[e]), sl2, -lse2));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            const int kv = kv0 + cbase + e;
            const float sv = __uint_as_float(e < 32 ? r0[e & 31] : r1[e & 31]);
            const bool dead = !qvalid || kv >= T || kv > q_idx;
            p[e] = dead ? 0.f : exp2f(fmaf(sv, sl2, -lse2));
          }
        }
      }
      uint32_t keep[DROP ? 8 : 1];                  // keep[g] bit i: key kv0 + cbase + 8 g + i of this query row survives
      if (DROP) {
#pragma unroll
        for (int g = 0; g < 8; ++g) keep[DROP ? g : 0] = dropout_keep8(seed, b * H + h, q_idx, (kv0 + cbase) / 8 + g, thr16);
      }
      if (it > 0) {
        // the previous tile's dQ reduce-add may still be reading its staging areas (the P/dS buffer and that tile's Q stage):
        // wait for it only now, after this tile's exponentials have been computed, then release the Q/dO stage for refill
        if (issuer) {
          tma_wait_group_read<0>();
          mbar_arrive(&qdo_empty[(it - 1) & 1]);     // second arrival (first: the MMA commit)
        }
        named_bar_sync(2, 256);
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float pd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pd[i] = DROP ? ((keep[DROP ? g : 0] >> i) & 1u ? p[8 * g + i] * inv_keep : 0.f) : p[8 * g + i];
        uint4 v4 = make_uint4(pack_bf16(pd[0], pd[1]), pack_bf16(pd[2], pd[3]), pack_bf16(pd[4], pd[5]), pack_bf16(pd[6], pd[7]));
        st_swizzled_chunk(sPS, 128, row, half * 8 + g, v4);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      // ---- dS = P ∘ (dP·scale − δ·scale)   (kept in registers until the dV MMA has finished reading P)
      mbar_wait(dp_full, ph);
      tc_fence_after();
      {
        uint32_t d0[32], d1[32];
        tmem_ld_x32(t_lane + Cfg::DP_COL + cbase, d0);
        tmem_ld_x32(t_lane + Cfg::DP_COL + cbase + 32, d1);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float s0 = scale, s1 = scale;
          if (DROP) {
            s0 = (keep[DROP ? e / 8 : 0] >> (e & 7)) & 1u ? scale * inv_keep : 0.f;
            s1 = (keep[DROP ? 4 + e / 8 : 0] >> (e & 7)) & 1u ? scale * inv_keep : 0.f;
          }
          p[e] *= fmaf(__uint_as_float(d0[e]), s0, -dsc);
          p[32 + e] *= fmaf(__uint_as_float(d1[e]), s1, -dsc);
        }
      }
      mbar_wait(dv_done, ph);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 v4 = make_uint4(pack_bf16(p[8 * g], p[8 * g + 1]), pack_bf16(p[8 * g + 2], p[8 * g + 3]),
                              pack_bf16(p[8 * g + 4], p[8 * g + 5]), pack_bf16(p[8 * g + 6], p[8 * g + 7]));
        st_swizzled_chunk(sPS, 128, row, half * 8 + g, v4);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(ds_ready);
      // ---- drain dQ (TMEM, aliasing dP) -> fp32 staging in the P/dS buffer -> TMA reduce-add into dq_accum
      mbar_wait(dq_full, ph);
      tc_fence_after();
      constexpr int ROUNDS = C / 64;               // every round moves 32 columns per column-half
      uint32_t dq[ROUNDS][32];
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) tmem_ld_x32(t_lane + Cfg::DP_COL + half * (C / 2) + rd * 32, dq[rd]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_drained);                     // TMEM side is done; the MMA warp may issue dP(it+1)
      // staging: 32 fp32 columns per (column-half, round) = one [128 x 128 B] box; round 0 lives in the P/dS buffer, round 1
      // (C = 128 only) in this tile's Q stage, which the TMA producer may not refill before the drain has released it.
      uint8_t* stage_q = sQ + (it & 1) * Cfg::TILE;
      if (!(dbg & 4))
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        uint8_t* stg = (rd == 0 ? sPS : stage_q) + half * (128 * 128) + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(stg + ((g ^ (row & 7)) << 4)) = make_uint4(dq[rd][4 * g], dq[rd][4 * g + 1], dq[rd][4 * g + 2], dq[rd][4 * g + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (issuer) {
        if (!(dbg & 1))
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            tma_reduce_add_3d(&tmdQ, (rd == 0 ? sPS : stage_q) + hh * (128 * 128), h * C + hh * (C / 2) + rd * 32, q0, b);
        tma_commit_group();                        // completion is awaited at the top of the next tile (overlaps its S wait + exps)
      }
    }
    if (issuer) tma_wait_group_read<0>();
    // ---- epilogue: dV, dK (thread = kv row x column half).  The last dq_full wait also covers the final dV/dK MMAs.
    const int kv_idx = kv0 + row;
    const bool kvalid = kv_idx < T;
    const long long krow = kvalid ? (static_cast<long long>(b) * T + kv_idx) : 0;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* dst = (which ? dk : dv) + krow * dkv_row_stride + static_cast<long long>(h) * C + half * (C / 2);
      const uint32_t col = (which ? Cfg::DK_COL : Cfg::DV_COL) + half * (C / 2);
#pragma unroll
      for (int ch = 0; ch < C / 64; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(t_lane + col + ch * 32, r);     // warp-collective: executed by every lane, stores are guarded
        tmem_ld_wait();
        if (kvalid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v4;
            v4.x = pack_bf16(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
            v4.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
            v4.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
            v4.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
            *reinterpret_cast<uint4*>(dst + ch * 32 + g * 8) = v4;
          }
        }
      }
    }
    if (issuer) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ================================================================================================
// Backward, version 3 (head_dim 128, T % 64 == 0): transposed scores, P/dS kept in tensor memory, 64-row query tiles,
// persistent CTAs.
//
//   Work item = one (batch, head, 128-key tile); inside it a loop over the 64-row query tiles at or below the diagonal.
//   Sᵀ = K·Qᵀ and dPᵀ = V·dOᵀ put KEYS on the TMEM lanes, so Pᵀ and dSᵀ (bf16, packed two per column, written back over
//   the fp32 columns they came from with tcgen05.st) are directly the A operands of dV += Pᵀ·dO and dK += dSᵀ·Q
//   (tcgen05.mma with A in TMEM): no P tile in shared memory, half the operand traffic for those two products.
//   dQᵀ = Kᵀ·dS (M = head_dim, N = 64 queries) has its OWN 64 TMEM columns and dPᵀ is double buffered, so nothing in a tile
//   aliases the next one.  Tensor-pipe order:  dV(i) | S(i+1) dP(i+1) | dK(i) dQ(i): the next tile's scores and dP are in
//   TMEM before the softmax warps finish dS(i), so their loop P(i) → dS(i) → P(i+1) never waits for the tensor core.
//   Four dedicated warps move results out one tile behind: dQ(i) (transposing staging store + TMA reduce-add) and, at the
//   end of an item, dV/dK (bf16 tiles + TMA store).
//   One CTA per SM walks a static list of items (heaviest key tiles first, round-robin over CTAs); every pipeline runs
//   on a tile counter that keeps counting across items, so the K/V/Q loads and the first Sᵀ/dPᵀ of the next item overlap
//   the dQ drain and the dV/dK epilogue of the current one -- with one CTA per item ~30 % of the kernel at T = 1024 was
//   pipeline fill and drain.  Version 2 (above in the file, still used for head_dim 64 / ragged T) had to walk each tile
//   in lock step because dQ aliased dP and P, dS and the dQ staging shared one buffer.
//
//   TMEM columns:  Sᵀ 0..63 (Pᵀ packed in the first 16 columns of each 32-column half) | dPᵀ[2] 64..191 (dSᵀ packed
//   likewise) | dQᵀ 192..255 | dV 256..383 | dK 384..511
//   warps 0-7 softmax (thread = key row × 32-query half), 8-11 output (dQ drain, dV/dK), 12 TMA producer, 13 MMA issuer.
// ================================================================================================
template <int C_>
struct Bwd3Cfg {
  static constexpr int C = C_, BQ = 64;
  static constexpr int KV_TILE = 128 * C * 2;          // K or V   [128 x C] bf16
  static constexpr int Q_TILE = BQ * C * 2;            // Q or dO  [64 x C] bf16
  static constexpr int NQ = 3, NDO = 2;                // pipeline depth of the Q and dO stages
  static constexpr int DS = 128 * BQ * 2;              // dSᵀ [128 keys x 64 queries] bf16 (B operand of the dQᵀ product)
  static constexpr int STG = BQ * C * 4;               // dQ staging [64 queries x C] fp32, 2 buffers; they double as the dV / dK output tiles
  static constexpr int STAT = 2 * BQ * 4;              // lse·log2e | delta·scale of one query tile
  static constexpr int SMEM = 2 * KV_TILE + (NQ + NDO) * Q_TILE + DS + 2 * STG + NQ * STAT + 512;   // + barriers, item ring
  static constexpr int ST_COL = 0, DPT_COL = 64, DQT_COL = 192, DV_COL = 256, DK_COL = 256 + C;   // dPᵀ: 2 buffers of 64 columns
  static constexpr int NSW = 8, NDW = 4;               // softmax warps (P, dS), output warps (dQ drain, dV/dK epilogue)
  static constexpr int NST = 32 * NSW, NDT = 32 * NDW;
  static constexpr int THREADS = NST + NDT + 64;       // + TMA warp + MMA warp
  static constexpr int CPT = 64 * 4 / NSW;             // query columns per softmax thread: 4 lane quarters x 2 column groups
};

template <int C_>
__global__ void __launch_bounds__(Bwd3Cfg<C_>::THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
                 const __grid_constant__ CUtensorMap tmdV, const float* __restrict__ lse2, const float* __restrict__ dsc,
                 int T, int H, int n_items, int* __restrict__ sched, float scale, int dbg) {
  using Cfg = Bwd3Cfg<C_>;
  constexpr int C = Cfg::C, CPT = Cfg::CPT;
  constexpr int W_OUT = Cfg::NSW, W_TMA = Cfg::NSW + Cfg::NDW, W_MMA = W_TMA + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::KV_TILE;
  uint8_t* sQ = sV + Cfg::KV_TILE;                       // NQ stages
  uint8_t* sdO = sQ + Cfg::NQ * Cfg::Q_TILE;             // NDO stages
  uint8_t* sDS = sdO + Cfg::NDO * Cfg::Q_TILE;
  uint8_t* sStg = sDS + Cfg::DS;                         // 2 buffers
  uint8_t* sStat = sStg + 2 * Cfg::STG;                  // NQ x (lse2[64] | dsc[64])
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + Cfg::NQ * Cfg::STAT);
  uint64_t* k_full = bars;           // per item
  uint64_t* q_full = bars + 1;       // [3] Q tile + its statistics landed
  uint64_t* q_empty = bars + 4;      // [3] dK(i) retired: Q stage (and its statistics) reusable
  uint64_t* do_full = bars + 7;      // [2]
  uint64_t* do_empty = bars + 9;     // [2] dV(i) retired
  uint64_t* s_full = bars + 11;
  uint64_t* dp_full = bars + 12;     // [2] dPᵀ(i) retired (buffer i & 1)
  uint64_t* p_ready = bars + 14;     // NST arrivals: Pᵀ(i) in TMEM
  uint64_t* ds_ready = bars + 15;    // NST arrivals: dSᵀ(i) in TMEM and in shared memory
  uint64_t* dq_full = bars + 16;     // dQᵀ(i) (and everything issued before it) retired
  uint64_t* dq_drained = bars + 17;  // NDT arrivals: dQᵀ(i) read out of TMEM
  uint64_t* v_full = bars + 18;      // per item
  uint64_t* kv_free = bars + 19;     // per item: every MMA that reads this item's K / V has retired
  uint64_t* acc_free = bars + 20;    // per item, NDT arrivals: dV / dK accumulators read out of TMEM
  uint64_t* item_full = bars + 21;   // [8] work-item ring: slot i holds the id of this CTA's (8k+i)-th item (-1 = no more work)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  volatile int* item_q = reinterpret_cast<volatile int*>(bars + 30);   // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq64 = T / 64;
  // Work items are handed out dynamically (one global counter, fetched by the TMA producer and passed to the other roles
  // through an 8-slot shared-memory ring: no role can be more than a few tiles -- at most three items -- behind it).
  // item w -> (batch*head, key tile): the key tiles of one (batch, head) are consecutive and start with the longest loop, so
  // the CTAs running at any moment share that head's Q / dO / dq_accum in L2, exactly like a plain grid would.
  const int njt = (T + 127) / 128;
  auto decode = [&](int w, int& jt, int& h, int& b) {
    const int bh = w / njt;
    jt = w - bh * njt;
    b = bh / H;
    h = bh - b * H;
  };
  auto next_item = [&](int ic) {                     // consumer side of the ring
    mbar_wait(&item_full[ic & 7], (ic >> 3) & 1);
    return item_q[ic & 7];
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQ);
    tma_prefetch_desc(&tmdK); tma_prefetch_desc(&tmdV);
    mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(kv_free, 1); mbar_init(acc_free, Cfg::NDT);
    for (int i = 0; i < 8; ++i) mbar_init(&item_full[i], 1);
    for (int s = 0; s < 3; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&do_full[s], 1); mbar_init(&do_empty[s], 1); mbar_init(&dp_full[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, Cfg::NST); mbar_init(ds_ready, Cfg::NST); mbar_init(dq_full, 1); mbar_init(dq_drained, Cfg::NDT);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int s = 0, g = 0;
      uint32_t qph = 0;
      for (int ic = 0;; ++ic) {
        int w = atomicAdd(sched, 1);
        if (w >= n_items) w = -1;
        item_q[ic & 7] = w;
        mbar_arrive(&item_full[ic & 7]);              // release: the id is visible to whoever completes the wait
        if (w < 0) break;
        int jt, h, b;
        decode(w, jt, h, b);
        const int kv0 = jt * 128, n_it = nq64 - 2 * jt;
        const long long stat_row = (static_cast<long long>(b) * H + h) * T;
        // issue order = order of first use: Q(0) and K feed Sᵀ(0); dO(0) and V feed dPᵀ(0)
        for (int it = 0; it < n_it; ++it, ++g) {
          const int q0 = kv0 + it * 64;
          mbar_wait(&q_empty[s], qph ^ 1);
          mbar_expect_tx(&q_full[s], Cfg::Q_TILE + Cfg::STAT);
#pragma unroll
          for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sQ + s * Cfg::Q_TILE + cb * (64 * 128), &tmQ, &q_full[s], cb * 64, h, q0, b);
          bulk_load_1d(sStat + s * Cfg::STAT, lse2 + stat_row + q0, 256, &q_full[s]);
          bulk_load_1d(sStat + s * Cfg::STAT + 256, dsc + stat_row + q0, 256, &q_full[s]);
          if (it == 0) {
            mbar_wait(kv_free, (ic & 1) ^ 1);          // previous item's last MMA has retired
            mbar_expect_tx(k_full, Cfg::KV_TILE);
#pragma unroll
            for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sK + cb * (128 * 128), &tmK, k_full, cb * 64, h, kv0, b);
          }
          const int d = g & 1;
          mbar_wait(&do_empty[d], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&do_full[d], Cfg::Q_TILE);
#pragma unroll
          for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sdO + d * Cfg::Q_TILE + cb * (64 * 128), &tmdO, &do_full[d], cb * 64, h, q0, b);
          if (it == 0) {
            mbar_expect_tx(v_full, Cfg::KV_TILE);
#pragma unroll
            for (int cb = 0; cb < C / 64; ++cb) tma_load_4d(sV + cb * (128 * 128), &tmV, v_full, cb * 64, h, kv0, b);
          }
          if (++s == 3) { s = 0; qph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    // The whole warp runs this (converged) control flow so that descriptors and addresses live in uniform registers; one
    // elected lane issues each group of tcgen05 instructions.  With `if (lane == 0)` around everything the compiler wraps
    // every MMA in a register->uniform-register waterfall loop and the issuing thread, not the tensor core, sets the pace.
    {
      constexpr uint32_t idesc_st = make_idesc_bf16(128, 64, 0, 0);    // Sᵀ, dPᵀ: A = K / V (K-major), B = Q / dO (K-major), N = 64
      constexpr uint32_t idesc_ts = make_idesc_bf16(128, C, 0, 1);     // dV, dK : A in TMEM, B = dO / Q MN-major, N = C
      constexpr uint32_t idesc_dq = make_idesc_bf16(C, 64, 1, 1);      // dQᵀ    : A = Kᵀ (MN-major, M = head_dim), B = dSᵀ tile (MN-major), N = 64
      // descriptors are built once; per MMA only a 64-bit add of a compile-time offset (see desc_shift)
      const uint64_t k_k = make_smem_desc(smem_u32(sK), 16, 1024), v_k = make_smem_desc(smem_u32(sV), 16, 1024);       // K-major views
      const uint64_t q_k0 = make_smem_desc(smem_u32(sQ), 16, 1024), do_k0 = make_smem_desc(smem_u32(sdO), 16, 1024);
      const uint64_t k_mn = make_smem_desc(smem_u32(sK), 128 * 128, 1024), ds_mn = make_smem_desc(smem_u32(sDS), 64 * 128, 1024);   // MN-major views
      const uint64_t q_mn0 = make_smem_desc(smem_u32(sQ), 64 * 128, 1024), do_mn0 = make_smem_desc(smem_u32(sdO), 64 * 128, 1024);
      // column of k-step ks (queries 16·ks ..) of a packed Pᵀ/dSᵀ operand: every thread packs its CPT columns into the first
      // CPT/2 columns of its own CPT-column group
      auto a_col = [](int ks) { return ((16 * ks) / CPT) * CPT + ((16 * ks) % CPT) / 2; };
      auto issue_st = [&](uint32_t d_col, uint64_t a_desc, uint64_t b_desc, uint64_t* bar) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            const uint32_t offa = (k / 4) * (128 * 128) + (k % 4) * 32, offb = (k / 4) * (64 * 128) + (k % 4) * 32;
            umma_bf16_ss(tmem + d_col, desc_shift(a_desc, offa), desc_shift(b_desc, offb), idesc_st, k != 0);
          }
          umma_commit(bar);
        }
        __syncwarp();
      };
      auto issue_s = [&](int s, uint32_t qph) {                 // Sᵀ of the tile whose Q sits in stage s
        mbar_wait(&q_full[s], qph);
        tc_fence_after();
        issue_st(Cfg::ST_COL, k_k, desc_shift(q_k0, s * Cfg::Q_TILE), s_full);
      };
      auto issue_dp = [&](int g) {                               // dPᵀ of global tile g
        mbar_wait(&do_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        issue_st(Cfg::DPT_COL + (g & 1) * 64, v_k, desc_shift(do_k0, (g & 1) * Cfg::Q_TILE), &dp_full[g & 1]);
      };
      int s_cur = 0, g = 0;
      uint32_t ph_cur = 0;
      for (int ic = 0;; ++ic) {
        const int w = next_item(ic);
        if (w < 0) break;
        int jt, h_, b_;
        decode(w, jt, h_, b_);
        const int n_it = nq64 - 2 * jt;
        mbar_wait(k_full, ic & 1);
        issue_s(s_cur, ph_cur);
        mbar_wait(v_full, ic & 1);
        issue_dp(g);
        for (int it = 0; it < n_it; ++it, ++g) {
          const uint32_t ph = g & 1;
          const int s_nxt = (s_cur == 2) ? 0 : s_cur + 1;
          const uint32_t ph_nxt = (s_nxt == 0) ? (ph_cur ^ 1) : ph_cur;
          // dV += Pᵀ(g) · dO(g)          A: packed bf16 in the Sᵀ columns;  B: dO as MN-major [k = query][n = c]
          mbar_wait(p_ready, ph);
          tc_fence_after();
          if (it == 0 && ic > 0) {
            mbar_wait(acc_free, (ic - 1) & 1);       // the output warps have read the previous item's dV / dK out of TMEM
            tc_fence_after();
          }
          const uint64_t do_mn = desc_shift(do_mn0, (g & 1) * Cfg::Q_TILE), q_mn = desc_shift(q_mn0, s_cur * Cfg::Q_TILE);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_ts(tmem + Cfg::DV_COL, tmem + Cfg::ST_COL + a_col(ks), desc_shift(do_mn, ks * 2048), idesc_ts, (it | ks) != 0);
            umma_commit(&do_empty[g & 1]);
          }
          __syncwarp();
          // next tile's Sᵀ (its columns: Pᵀ(g) was consumed by dV(g), just ahead in the pipe) and dPᵀ (other buffer: dSᵀ(g-1)
          // was consumed by dK(g-1)) -- both are in TMEM before the softmax warps finish dS(g), so they never wait for them
          if (it + 1 < n_it) {
            issue_s(s_nxt, ph_nxt);
            issue_dp(g + 1);
          }
          // dK += dSᵀ(g) · Q(g) ;  dQᵀ(g) = Kᵀ · dS(g)
          mbar_wait(ds_ready, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_ts(tmem + Cfg::DK_COL, tmem + Cfg::DPT_COL + (g & 1) * 64 + a_col(ks), desc_shift(q_mn, ks * 2048), idesc_ts, (it | ks) != 0);
            umma_commit(&q_empty[s_cur]);
          }
          __syncwarp();
          if (g > 0) {
            mbar_wait(dq_drained, ph ^ 1);           // dQᵀ(g-1) read out of its columns by the output warps
            tc_fence_after();
          }
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_ss(tmem + Cfg::DQT_COL, desc_shift(k_mn, ks * 2048), desc_shift(ds_mn, ks * 2048), idesc_dq, ks != 0);
            umma_commit(dq_full);
            if (it + 1 == n_it) umma_commit(kv_free);
          }
          __syncwarp();
          s_cur = s_nxt;
          ph_cur = ph_nxt;
        }
      }
    }
  } else if (warp >= W_OUT) {
    // ------------------------------- output warps -------------------------------
    // dQᵀ(g): lane = head-dim index, 64 query columns  ->  transposed into fp32 staging rows of 32 head-dim values (one
    // 128-byte swizzled row per query and 32-column box)  ->  TMA reduce-add into dq_accum.  Runs a tile behind the softmax
    // warps and never blocks them: the only consumer of dq_drained is the MMA warp (before it overwrites the columns).
    // After an item's last tile: dV and dK (lane = key row) -> bf16 swizzled tiles -> TMA store (clips key rows past T).
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool issuer = (warp == W_OUT && lane == 0);
    int g = 0;
    for (int ic = 0;; ++ic) {
      const int w = next_item(ic);
      if (w < 0) break;
      int jt, h, b;
      decode(w, jt, h, b);
      const int kv0 = jt * 128, n_it = nq64 - 2 * jt;
      for (int j = 0; j < n_it; ++j, ++g) {
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        uint32_t dq0[32], dq1[32];
        tmem_ld_x32(t_lane + Cfg::DQT_COL, dq0);
        tmem_ld_x32(t_lane + Cfg::DQT_COL + 32, dq1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_drained);
        if (issuer) {                              // staging buffer g & 1: its previous user is two bulk groups back
          if (j == 0) tma_wait_group_read<0>();    // (first tile of an item: the previous item's dV / dK stores used both buffers)
          else tma_wait_group_read<1>();
        }
        named_bar_sync(1, Cfg::NDT);
        // head-dim index of this lane's accumulator row.  M = 128: lane i of the CTA holds row i.  M = 64 (head_dim 64): the
        // accumulator occupies the first 16 lanes of every 32-lane quarter, row = 16 * quarter + lane (cute tmem_frg_1sm, M_MMA = 64).
        const int d = (C == 128) ? row : quarter * 16 + (lane & 15);
        const bool d_valid = (C == 128) || lane < 16;
        const int d4 = (d & 31) >> 2;              // 16-byte chunk of the 32-column box
        uint8_t* stg = sStg + (g & 1) * Cfg::STG + (d >> 5) * (64 * 128) + (d & 3) * 4;
        if (d_valid) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            *reinterpret_cast<uint32_t*>(stg + e * 128 + ((d4 ^ (e & 7)) << 4)) = dq0[e];
            *reinterpret_cast<uint32_t*>(stg + (32 + e) * 128 + ((d4 ^ (e & 7)) << 4)) = dq1[e];
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, Cfg::NDT);
        if (issuer) {
          if (!(dbg & 1))                          // MIDGPT_ATTN_DEBUG=1: what-if without the reduce-add (wrong dQ)
#pragma unroll
          for (int bx = 0; bx < C / 32; ++bx) tma_reduce_add_3d(&tmdQ, sStg + (g & 1) * Cfg::STG + bx * (64 * 128), h * C + bx * 32, kv0 + j * 64, b);
          tma_commit_group();
        }
      }
      // ---- item epilogue.  dq_full of the last tile (observed above) covers every MMA of the item.  dV goes out through the
      //      staging buffer of the second-to-last tile, dK is packed in registers (so the accumulators are released before
      //      waiting for the last reduce-add) and then goes out through the other buffer.
      uint8_t* tile_v = sStg + (g & 1) * Cfg::STG;          // g already counts the next tile: (g & 1) == buffer of tile g-2
      uint8_t* tile_k = sStg + ((g & 1) ^ 1) * Cfg::STG;
      if (issuer) tma_wait_group_read<1>();
      named_bar_sync(1, Cfg::NDT);
#pragma unroll
      for (int ch = 0; ch < C / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(t_lane + Cfg::DV_COL + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          st_swizzled_chunk(tile_v, 128, row, ch * 4 + q4,
                            make_uint4(pack_bf16(__uint_as_float(r[8 * q4]), __uint_as_float(r[8 * q4 + 1])),
                                       pack_bf16(__uint_as_float(r[8 * q4 + 2]), __uint_as_float(r[8 * q4 + 3])),
                                       pack_bf16(__uint_as_float(r[8 * q4 + 4]), __uint_as_float(r[8 * q4 + 5])),
                                       pack_bf16(__uint_as_float(r[8 * q4 + 6]), __uint_as_float(r[8 * q4 + 7]))));
      }
      uint32_t kk[C / 2];
#pragma unroll
      for (int ch = 0; ch < C / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(t_lane + Cfg::DK_COL + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) kk[ch * 16 + e] = pack_bf16(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
      }
      tc_fence_before();
      mbar_arrive(acc_free);                       // the next item's first dV / dK MMAs may overwrite the accumulators
      fence_proxy_async_smem();
      named_bar_sync(1, Cfg::NDT);
      if (issuer) {
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_store_4d(&tmdV, tile_v + cb * (128 * 128), cb * 64, h, kv0, b);
        tma_commit_group();
        tma_wait_group_read<1>();                  // everything but the dV store: the last reduce-add has left tile_k
      }
      named_bar_sync(1, Cfg::NDT);
#pragma unroll
      for (int cc = 0; cc < C / 8; ++cc)
        st_swizzled_chunk(tile_k, 128, row, cc, make_uint4(kk[4 * cc], kk[4 * cc + 1], kk[4 * cc + 2], kk[4 * cc + 3]));
      fence_proxy_async_smem();
      named_bar_sync(1, Cfg::NDT);
      if (issuer) {
#pragma unroll
        for (int cb = 0; cb < C / 64; ++cb) tma_store_4d(&tmdK, tile_k + cb * (128 * 128), cb * 64, h, kv0, b);
        tma_commit_group();
      }
    }
    if (issuer) tma_wait_group<0>();
  } else {
    // ------------------------------- softmax warps -------------------------------
    const int quarter = warp & 3, grp = warp >> 2;   // TMEM lane quarter, query-column group
    const int row = quarter * 32 + lane;             // key row inside the tile == TMEM lane
    const int c0 = grp * CPT;                        // this thread's first query column
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = scale * kLog2e;

    int s = 0, g = 0;
    uint32_t qph = 0;
    for (int ic = 0;; ++ic) {
      const int w = next_item(ic);
      if (w < 0) break;
      int jt, h_, b_;
      decode(w, jt, h_, b_);
      const int kv0 = jt * 128, n_it = nq64 - 2 * jt;
      const int kv_idx = kv0 + row;
      for (int it = 0; it < n_it; ++it, ++g) {
        const uint32_t ph = g & 1;
        const int q0 = kv0 + it * 64;
        mbar_wait(&q_full[s], qph);                    // statistics of this tile are in shared memory
        const float4* st_lse = reinterpret_cast<const float4*>(sStat + s * Cfg::STAT) + c0 / 4;
        const float4* st_del = st_lse + 16;
        float p[CPT];
        // ---- Pᵀ = exp2(Sᵀ·c − lse·log2e)
        mbar_wait(s_full, ph);
        tc_fence_after();
        {
          uint32_t r[CPT];
          tmem_ld_x32(t_lane + Cfg::ST_COL + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < CPT / 4; ++q4) {
            const float4 l4 = st_lse[q4];
            p[4 * q4] = exp2f(fmaf(__uint_as_float(r[4 * q4]), sl2, -l4.x));
            p[4 * q4 + 1] = exp2f(fmaf(__uint_as_float(r[4 * q4 + 1]), sl2, -l4.y));
            p[4 * q4 + 2] = exp2f(fmaf(__uint_as_float(r[4 * q4 + 2]), sl2, -l4.z));
            p[4 * q4 + 3] = exp2f(fmaf(__uint_as_float(r[4 * q4 + 3]), sl2, -l4.w));
          }
          if (it < 2) {                                // the two 64-row tiles on the diagonal of this 128-key tile
#pragma unroll
            for (int e = 0; e < CPT; ++e)
              if (kv_idx > q0 + c0 + e || kv_idx >= T) p[e] = 0.f;
          }
          uint32_t pk[CPT / 2];
#pragma unroll
          for (int e = 0; e < CPT / 2; ++e) pk[e] = pack_bf16(p[2 * e], p[2 * e + 1]);
          tmem_st_x16(t_lane + Cfg::ST_COL + c0, pk);     // over this thread's own (already read) Sᵀ columns
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(p_ready);
        // ---- dSᵀ = Pᵀ ∘ (dPᵀ·scale − δ·scale)
        mbar_wait(&dp_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        {
          uint32_t d[CPT];
          tmem_ld_x32(t_lane + Cfg::DPT_COL + (g & 1) * 64 + c0, d);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < CPT / 4; ++q4) {
            const float4 d4 = st_del[q4];
            p[4 * q4] *= fmaf(__uint_as_float(d[4 * q4]), scale, -d4.x);
            p[4 * q4 + 1] *= fmaf(__uint_as_float(d[4 * q4 + 1]), scale, -d4.y);
            p[4 * q4 + 2] *= fmaf(__uint_as_float(d[4 * q4 + 2]), scale, -d4.z);
            p[4 * q4 + 3] *= fmaf(__uint_as_float(d[4 * q4 + 3]), scale, -d4.w);
          }
          uint32_t pk[CPT / 2];
#pragma unroll
          for (int e = 0; e < CPT / 2; ++e) pk[e] = pack_bf16(p[2 * e], p[2 * e + 1]);
          tmem_st_x16(t_lane + Cfg::DPT_COL + (g & 1) * 64 + c0, pk);
          if (g > 0) mbar_wait(dq_full, ph ^ 1);       // dQᵀ(g-1) has retired: the (single) dSᵀ shared-memory tile may be overwritten
          // the same values as the MN-major B tile of the dQᵀ product: row = key, 64 queries (128 bytes, swizzled) per row
          uint8_t* rowp = sDS + row * 128;
#pragma unroll
          for (int q8 = 0; q8 < CPT / 8; ++q8)
            *reinterpret_cast<uint4*>(rowp + (((c0 / 8 + q8) ^ (row & 7)) << 4)) = make_uint4(pk[4 * q8], pk[4 * q8 + 1], pk[4 * q8 + 2], pk[4 * q8 + 3]);
          tmem_st_wait();
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(ds_ready);
        if (++s == 3) { s = 0; qph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<512>(tmem);
}

// ================================================================================================
// launchers
// ================================================================================================
template <typename K>
static void set_smem(K kern, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(attention): ") + cudaGetErrorString(e));
}
static void check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

struct DropArgs {
  uint32_t thr16;      // round(p * 65536); 0 = no dropout
  float inv_keep;
  uint2 seed;
};
static DropArgs drop_args(float p_drop, unsigned long long seed) {
  if (p_drop < 0.f || p_drop >= 1.f) throw std::runtime_error("attention: dropout probability must be in [0, 1)");
  DropArgs d;
  d.thr16 = static_cast<uint32_t>(lrintf(p_drop * 65536.f));
  d.inv_keep = 1.f / (1.f - static_cast<float>(d.thr16) / 65536.f);
  d.seed = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return d;
}

template <int C, bool DROP>
static void launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int T, int H, long long q_row_stride,
                       long long k_row_stride, long long v_row_stride, const DropArgs& d, cudaStream_t s) {
  using Cfg = FwdCfg<C>;
  const float scale = 1.f / sqrtf(static_cast<float>(C));
  dim3 grid((T + 127) / 128, H, B);
  CUtensorMap tq = make_tmap_bthc(q, B, T, H, C, q_row_stride, Cfg::BQ), tk = make_tmap_bthc(k, B, T, H, C, k_row_stride, Cfg::BKV),
              tv = make_tmap_bthc(v, B, T, H, C, v_row_stride, Cfg::BKV);
  static bool once = (set_smem(attn_fwd_kernel<C, DROP>, Cfg::SMEM), true);
  (void)once;
  static const int eager = [] { const char* e = getenv("MIDGPT_ATTN_DEBUG"); return e ? (atoi(e) & 8) : 0; }();
  attn_fwd_kernel<C, DROP><<<grid, 192, Cfg::SMEM, s>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(o), lse, T, H, scale, d.thr16, d.inv_keep, d.seed, eager);
}

void attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int T, int H, int C, long long q_row_stride,
              long long k_row_stride, long long v_row_stride, float p_drop, unsigned long long seed, int num_sms, cudaStream_t s) {
  const DropArgs d = drop_args(p_drop, seed);
  if (C != 64 && C != 128) throw std::runtime_error("attention: head_dim must be 64 or 128");
  if (C == 64) {
    if (d.thr16) launch_fwd<64, true>(q, k, v, o, lse, B, T, H, q_row_stride, k_row_stride, v_row_stride, d, s);
    else launch_fwd<64, false>(q, k, v, o, lse, B, T, H, q_row_stride, k_row_stride, v_row_stride, d, s);
  } else {
    if (d.thr16) launch_fwd<128, true>(q, k, v, o, lse, B, T, H, q_row_stride, k_row_stride, v_row_stride, d, s);
    else launch_fwd<128, false>(q, k, v, o, lse, B, T, H, q_row_stride, k_row_stride, v_row_stride, d, s);
  }
  check("attn_fwd");
}

// Work counter of the backward kernel (one int per device, zeroed on the launch stream before every launch; launches of one
// device are stream-ordered in this engine, so one counter per device is enough).
static int* sched_counter(cudaStream_t s) {
  static int* ptrs[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ptrs[dev] && cudaMalloc(&ptrs[dev], sizeof(int)) != cudaSuccess) throw std::runtime_error("attention: cannot allocate the work counter");
  if (cudaMemsetAsync(ptrs[dev], 0, sizeof(int), s) != cudaSuccess) throw std::runtime_error("attention: cannot reset the work counter");
  return ptrs[dev];
}

void attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, float* delta, void* dq,
              void* dk, void* dv, float* dq_accum, int B, int T, int H, int C, long long q_row_stride, long long k_row_stride,
              long long v_row_stride, long long dqkv_row_stride, float p_drop, unsigned long long seed, int num_sms, cudaStream_t s) {
  const float scale = 1.f / sqrtf(static_cast<float>(C));
  const DropArgs d = drop_args(p_drop, seed);
  const long long nvec = static_cast<long long>(B) * T * H;
  if (nvec >= (1ll << 31)) throw std::runtime_error("attention: too many (token, head) vectors for 32-bit indexing");
  if (C == 64)
    attn_bwd_prep_kernel<64, 4><<<num_sms * 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(dout), delta, lse, scale, static_cast<size_t>(B) * H * T, static_cast<unsigned>(nvec), T, H);
  else if (C == 128)
    attn_bwd_prep_kernel<128, 4><<<num_sms * 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(dout), delta, lse, scale, static_cast<size_t>(B) * H * T, static_cast<unsigned>(nvec), T, H);
  else
    throw std::runtime_error("attention: head_dim must be 64 or 128");
  check("attn_bwd_prep");
  dim3 grid((T + 127) / 128, H, B);
  CUtensorMap tq = make_tmap_bthc(q, B, T, H, C, q_row_stride, 128), tk = make_tmap_bthc(k, B, T, H, C, k_row_stride, 128),
              tv = make_tmap_bthc(v, B, T, H, C, v_row_stride, 128), tdo = make_tmap_bthc(dout, B, T, H, C, static_cast<long long>(H) * C, 128);
  CUtensorMap tdq = make_tmap_dq(dq_accum, B, T, H * C);
  static const int dbg = [] { const char* e = getenv("MIDGPT_ATTN_DEBUG"); return e ? atoi(e) : 0; }();
  static const bool use_v2 = [] { const char* e = getenv("MIDGPT_ATTN_BWD"); return e && std::string(e) == "v2"; }();
  // dropout: the transposed (v3) kernel's threads own one key and 32 queries, i.e. one Philox call per element instead of
  // one per eight; the row-major v2 kernel regenerates the forward's mask at the forward's cost, so it serves p > 0.
  if (T % 64 == 0 && !use_v2 && d.thr16 == 0) {
    CUtensorMap tq64 = make_tmap_bthc(q, B, T, H, C, q_row_stride, 64), tdo64 = make_tmap_bthc(dout, B, T, H, C, static_cast<long long>(H) * C, 64);
    CUtensorMap tdq64 = make_tmap_dq(dq_accum, B, T, H * C, 64);
    const size_t plane = static_cast<size_t>(B) * H * T;          // delta planes: raw | delta*scale | lse*log2e (attn_bwd_prep_kernel)
    CUtensorMap tdk = make_tmap_bthc(dk, B, T, H, C, dqkv_row_stride, 128), tdv = make_tmap_bthc(dv, B, T, H, C, dqkv_row_stride, 128);
    const int n_items = static_cast<int>(grid.x) * H * B;         // persistent: one CTA per SM pulls items off a global counter
    if (C == 128) {
      static bool once = (set_smem(attn_bwd3_kernel<128>, Bwd3Cfg<128>::SMEM), true);
      (void)once;
      attn_bwd3_kernel<128><<<std::min(num_sms, n_items), Bwd3Cfg<128>::THREADS, Bwd3Cfg<128>::SMEM, s>>>(
          tq64, tk, tv, tdo64, tdq64, tdk, tdv, delta + 2 * plane, delta + plane, T, H, n_items, sched_counter(s), scale, dbg);
    } else {
      static bool once = (set_smem(attn_bwd3_kernel<64>, Bwd3Cfg<64>::SMEM), true);
      (void)once;
      attn_bwd3_kernel<64><<<std::min(num_sms, n_items), Bwd3Cfg<64>::THREADS, Bwd3Cfg<64>::SMEM, s>>>(
          tq64, tk, tv, tdo64, tdq64, tdk, tdv, delta + 2 * plane, delta + plane, T, H, n_items, sched_counter(s), scale, dbg);
    }
  } else {
#define MG_BWD2(CC, DR)                                                                                                              \
    {                                                                                                                                \
      static bool once = (set_smem(attn_bwd_kernel<CC, DR>, BwdCfg<CC>::SMEM), true);                                                \
      (void)once;                                                                                                                    \
      attn_bwd_kernel<CC, DR><<<grid, BwdCfg<CC>::THREADS, BwdCfg<CC>::SMEM, s>>>(tq, tk, tv, tdo, tdq, lse, delta,                  \
          static_cast<__nv_bfloat16*>(dk), static_cast<__nv_bfloat16*>(dv), dqkv_row_stride, T, H, scale, dbg, d.thr16, d.inv_keep, d.seed); \
    }
    if (C == 64 && d.thr16) MG_BWD2(64, true)
    else if (C == 64) MG_BWD2(64, false)
    else if (d.thr16) MG_BWD2(128, true)
    else MG_BWD2(128, false)
#undef MG_BWD2
  }
  check("attn_bwd");
  if (dq != nullptr) {      // dq == nullptr: the caller consumes the fp32 accumulator directly (qkln_rope_bwd reads and clears it)
    dq_convert_kernel<<<num_sms * 8, 256, 0, s>>>(dq_accum, static_cast<__nv_bfloat16*>(dq), static_cast<long long>(B) * T, H * C, dqkv_row_stride);
    check("dq_convert");
  }
}

}  // namespace mg
