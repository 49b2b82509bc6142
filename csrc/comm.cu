// In-kernel NVLink / NVSwitch communication over symmetric memory (peer-mapped buffers + signal pads), sm_100a.
//
//   xrank_barrier          all-ranks device barrier (system-scope release/acquire flags in each rank's signal pad)
//   rs_sumsq               reduce-scatter ⊕ 1/world scale ⊕ Σg² : every rank PULLS its shard of every peer's fp32
//                          gradient accumulator (P2P ld over NVLink, or one multimem.ld_reduce that adds inside the
//                          switch), writes the reduced fp32 shard locally and accumulates the clip statistic (X3+K14+K15)
//   adamw_allgather        clip · Adam · independent-WD · LR (K15-K18) on the local fp32 shard, and the bf16 result is
//                          PUSHED straight into every rank's gathered-weights buffer (st over NVLink / multimem.st):
//                          the optimizer step IS the parameter all-gather (X1/X4 + K13)
//   ag_pull                gather peers' bf16 shards into the local buffer chunk by chunk, publishing per-chunk ready
//                          flags that the tcgen05 GEMM's TMA producer polls (all-gather ⊕ GEMM, csrc/gemm.cu)
//
// Flag protocol: slot s of rank r's signal pad holds, per peer p, the last epoch p has reached at that slot.
// Epochs increase monotonically (passed from the host), so pads never need clearing.
#include "common.cuh"

#include <stdexcept>
#include <string>

namespace mg {

constexpr int MAX_RANKS = 8;

struct PeerPtrs {
  void* p[MAX_RANKS];
};

MG_DEVICE void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
MG_DEVICE uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
MG_DEVICE float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
MG_DEVICE float4 multimem_ld_reduce_f4(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
MG_DEVICE void multimem_st_u2(uint2* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// barrier: thread t < world signals peer t, then waits for peer t's signal
// ------------------------------------------------------------------------------------------------
__global__ void xrank_barrier_kernel(PeerPtrs pads, int rank, int world, int slot, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    uint32_t* theirs = reinterpret_cast<uint32_t*>(pads.p[t]) + slot * MAX_RANKS + rank;
    st_release_sys(theirs, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[rank]) + slot * MAX_RANKS + t;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reduce-scatter ⊕ scale ⊕ Σg²   (pull).  grads[p] = peer p's full fp32 accumulator (same layout on every rank).
// ------------------------------------------------------------------------------------------------
template <bool MULTIMEM>
__global__ void __launch_bounds__(512) rs_sumsq_kernel(PeerPtrs grads, const float* __restrict__ mc_grads, float* __restrict__ out_shard,
                                                       float* __restrict__ sumsq, long long shard_off4, long long n4, int rank, int world,
                                                       float inv_world, int accumulate, int want_sumsq) {
  // accumulate: out_shard += mean (sharded gradient accumulation over micro-steps: the reduce-scatter runs every micro-step and
  // only a shard-sized fp32 accumulator persists, like the reference's per-microstep sharding constraint, src/train.py:87);
  // want_sumsq: add Σ out² of the FINAL value to *sumsq (last micro-step only).
  __shared__ float part[16];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 s;
    if (MULTIMEM) {
      s = multimem_ld_reduce_f4(reinterpret_cast<const float4*>(mc_grads) + shard_off4 + i);
    } else {
      s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int k = 0; k < world; ++k) {
        const int p = (rank + k) % world;          // start with the local copy, then walk the ring: spreads link load
        const float4 v = ld_peer_f4(reinterpret_cast<const float4*>(grads.p[p]) + shard_off4 + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    s.x *= inv_world; s.y *= inv_world; s.z *= inv_world; s.w *= inv_world;
    if (accumulate) {
      const float4 o = reinterpret_cast<const float4*>(out_shard)[i];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    reinterpret_cast<float4*>(out_shard)[i] = s;
    acc += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
  }
  if (!want_sumsq) return;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sumsq, v);
  }
}

// ------------------------------------------------------------------------------------------------
// AdamW ⊕ all-gather (push).  lp[p] = rank p's gathered bf16 buffer; this rank owns elements [shard_off, shard_off+n).
// ------------------------------------------------------------------------------------------------
struct AdamArgsC {
  float lr, b1, b2, eps, wd_over_lr_peak, inv_bc1, inv_bc2;
};

template <bool MULTIMEM>
__global__ void __launch_bounds__(256) adamw_allgather_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                              const float* __restrict__ g, PeerPtrs lp, __nv_bfloat16* __restrict__ mc_lp,
                                                              long long shard_off4, long long n4, const float* __restrict__ clip_ptr,
                                                              AdamArgsC a, int world) {
  const float clip = *clip_ptr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, v2[4] = {vv.x, vv.y, vv.z, vv.w};
    const float gg[4] = {gv.x * clip, gv.y * clip, gv.z * clip, gv.w * clip};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = a.b1 * mm[e] + (1.f - a.b1) * gg[e];
      v2[e] = a.b2 * v2[e] + (1.f - a.b2) * gg[e] * gg[e];
      const float u = (mm[e] * a.inv_bc1) / (sqrtf(v2[e] * a.inv_bc2) + a.eps) + a.wd_over_lr_peak * pp[e];
      pp[e] -= a.lr * u;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(v2[0], v2[1], v2[2], v2[3]);
    const uint2 packed = make_uint2(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]));
    if (MULTIMEM) {
      multimem_st_u2(reinterpret_cast<uint2*>(mc_lp) + shard_off4 + i, packed);     // one store, delivered to every rank by the switch
    } else {
#pragma unroll 1
      for (int k = 0; k < world; ++k) reinterpret_cast<uint2*>(lp.p[k])[shard_off4 + i] = packed;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// all-gather pull with per-chunk ready flags (consumed by the GEMM's TMA producer: all-gather ⊕ GEMM)
//   local[c*shard .. (c+1)*shard) <- peer c's same range, for every c != rank; flags[c] += 1 per finished CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) ag_pull_kernel(PeerPtrs lp, __nv_bfloat16* __restrict__ local, long long shard16, int rank, int world,
                                                      uint32_t* __restrict__ flags) {
  for (int k = 1; k < world; ++k) {
    const int c = (rank + k) % world;
    const uint4* src = reinterpret_cast<const uint4*>(lp.p[c]);          // peer c's shard of this unit
    uint4* dst = reinterpret_cast<uint4*>(local) + static_cast<long long>(c) * shard16;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < shard16; i0 += stride * 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * stride;
        if (i < shard16)
          asm volatile("ld.global.relaxed.sys.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * stride;
        if (i < shard16) dst[i] = v[u];
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.global;" ::: "memory");   // later read by TMA (async proxy) in the GEMM
      atomicAdd(flags + c, 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------------
static void check_launch_c(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
static PeerPtrs to_peers(const long long* ptrs, int world) {
  if (world > MAX_RANKS) throw std::runtime_error("comm: at most 8 ranks per shard group");
  PeerPtrs pp;
  for (int i = 0; i < MAX_RANKS; ++i) pp.p[i] = i < world ? reinterpret_cast<void*>(ptrs[i]) : nullptr;
  return pp;
}

void xrank_barrier(const long long* pad_ptrs, int rank, int world, int slot, unsigned epoch, cudaStream_t s) {
  xrank_barrier_kernel<<<1, 32, 0, s>>>(to_peers(pad_ptrs, world), rank, world, slot, epoch);
  check_launch_c("xrank_barrier");
}

void rs_sumsq(const long long* grad_ptrs, long long mc_ptr, float* out_shard, float* sumsq, long long shard_off, long long n, int rank,
              int world, int accumulate, int want_sumsq, int num_sms, cudaStream_t s) {
  if (n % 4 || shard_off % 4) throw std::runtime_error("rs_sumsq: offsets must be multiples of 4 elements");
  PeerPtrs pp = to_peers(grad_ptrs, world);
  const int grid = num_sms * 2;
  if (mc_ptr)
    rs_sumsq_kernel<true><<<grid, 512, 0, s>>>(pp, reinterpret_cast<const float*>(mc_ptr), out_shard, sumsq, shard_off / 4, n / 4, rank, world, 1.f / world, accumulate, want_sumsq);
  else
    rs_sumsq_kernel<false><<<grid, 512, 0, s>>>(pp, nullptr, out_shard, sumsq, shard_off / 4, n / 4, rank, world, 1.f / world, accumulate, want_sumsq);
  check_launch_c("rs_sumsq");
}

void adamw_allgather(float* p, float* m, float* v, const float* g, const long long* lp_ptrs, long long mc_ptr, long long shard_off,
                     long long n, const float* clip_ptr, float lr, float b1, float b2, float eps, float wd_over_lr_peak, float bc1,
                     float bc2, int world, int num_sms, cudaStream_t s) {
  if (n % 4 || shard_off % 4) throw std::runtime_error("adamw_allgather: offsets must be multiples of 4 elements");
  AdamArgsC a{lr, b1, b2, eps, wd_over_lr_peak, 1.f / bc1, 1.f / bc2};
  PeerPtrs pp = to_peers(lp_ptrs, world);
  if (mc_ptr)
    adamw_allgather_kernel<true><<<num_sms * 4, 256, 0, s>>>(p, m, v, g, pp, reinterpret_cast<__nv_bfloat16*>(mc_ptr), shard_off / 4, n / 4, clip_ptr, a, world);
  else
    adamw_allgather_kernel<false><<<num_sms * 4, 256, 0, s>>>(p, m, v, g, pp, nullptr, shard_off / 4, n / 4, clip_ptr, a, world);
  check_launch_c("adamw_allgather");
}

void ag_pull(const long long* lp_ptrs, void* local, long long shard_elems, int rank, int world, unsigned* flags, int n_ctas, cudaStream_t s) {
  if (shard_elems % 8) throw std::runtime_error("ag_pull: shard must be a multiple of 8 bf16 elements");
  ag_pull_kernel<<<n_ctas, 512, 0, s>>>(to_peers(lp_ptrs, world), static_cast<__nv_bfloat16*>(local), shard_elems / 8, rank, world, flags);
  check_launch_c("ag_pull");
}

}  // namespace mg
