// Memory-bound kernels of the model (everything that is not a GEMM or attention), sm_100a.
//   rmsnorm fwd/bwd (K2), QK-LayerNorm + RoPE fwd/bwd (K4,K5), embedding gather / scatter-add (K1),
//   cross-entropy over a logits chunk (K11), Philox dropout (K12), fused clip+Adam+WD+LR+bf16-cast (K13-K18),
//   sum-of-squares (K15).
// All of them are single-pass over HBM with 16-byte vector accesses and fp32 math.
#include "common.cuh"

#include <curand_philox4x32_x.h>
#include <stdexcept>
#include <string>

namespace mg {

static inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------
// RMSNorm (weightless).  One warp per row, the row is kept in registers (D <= 32*8*MAXV elements).
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ rstd_out, int M, int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = D >> 3;  // uint4 = 8 bf16
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * D);
  uint4 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      v[i] = __ldg(xr + j);
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = unpack_bf16(u[e]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / D + eps);
  if (lane == 0) rstd_out[warp] = r;
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(warp) * D);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = unpack_bf16(u[e]);
        o[e] = pack_bf16(f.x * r, f.y * r);
      }
      yr[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// dx = rstd * (dy - yhat * mean(dy*yhat)) (+ dres),  yhat = x * rstd
template <int MAXV>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dres,
                                                          __nv_bfloat16* __restrict__ dx, int M, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = D >> 3;
  const size_t off = static_cast<size_t>(warp) * D;
  const uint4* xr = reinterpret_cast<const uint4*>(x + off);
  const uint4* gr = reinterpret_cast<const uint4*>(dy + off);
  const float r = rstd[warp];
  uint4 xv[MAXV], gv[MAXV];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      xv[i] = __ldg(xr + j);
      gv[i] = __ldg(gr + j);
      const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      const uint32_t gu[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 a = unpack_bf16(xu[e]), g = unpack_bf16(gu[e]);
        dot += g.x * a.x * r + g.y * a.y * r;
      }
    }
  }
  dot = warp_sum(dot) / D;
  uint4* outr = reinterpret_cast<uint4*>(dx + off);
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + off) : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      const uint32_t gu[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
      uint4 rv = rr ? __ldg(rr + j) : make_uint4(0, 0, 0, 0);
      const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 a = unpack_bf16(xu[e]), g = unpack_bf16(gu[e]), d = unpack_bf16(ru[e]);
        float o0 = r * (g.x - a.x * r * dot) + d.x;
        float o1 = r * (g.y - a.y * r * dot) + d.y;
        o[e] = pack_bf16(o0, o1);
      }
      outr[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename F>
static void dispatch_maxv(int D, F&& f) {
  const int nvec = D / 8;
  if (D % 8) throw std::runtime_error("rmsnorm: D must be a multiple of 8");
  if (nvec <= 32) f(std::integral_constant<int, 1>());
  else if (nvec <= 64) f(std::integral_constant<int, 2>());
  else if (nvec <= 128) f(std::integral_constant<int, 4>());
  else if (nvec <= 256) f(std::integral_constant<int, 8>());
  else if (nvec <= 512) f(std::integral_constant<int, 16>());
  else throw std::runtime_error("rmsnorm: D > 4096 not supported");
}

void rmsnorm_fwd(const void* x, void* y, float* rstd, int M, int D, float eps, cudaStream_t s) {
  dispatch_maxv(D, [&](auto mv) {
    constexpr int MAXV = decltype(mv)::value;
    rmsnorm_fwd_kernel<MAXV><<<(M + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), rstd, M, D, eps);
  });
  check_launch("rmsnorm_fwd");
}
void rmsnorm_bwd(const void* dy, const void* x, const float* rstd, const void* dres, void* dx, int M, int D, cudaStream_t s) {
  dispatch_maxv(D, [&](auto mv) {
    constexpr int MAXV = decltype(mv)::value;
    rmsnorm_bwd_kernel<MAXV><<<(M + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), rstd,
                                                          static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), M, D);
  });
  check_launch("rmsnorm_bwd");
}

// ------------------------------------------------------------------------------------------------
// QK-LayerNorm (gain only, shared across heads) + interleaved RoPE.
// One warp per (token, q|k, head) vector of C elements; lane holds EPL = C/32 consecutive elements
// (EPL even => RoPE pairs never straddle lanes).  sin/cos come from a [T, C/2] fp32 table.
// Input  qkv  [M, 3, H, C] bf16 (raw projection);  output qk [M, 2, H, C] bf16;  stats [M, 2, H, 2] fp32.
// ------------------------------------------------------------------------------------------------
// (QK-LayerNorm + RoPE kernels)
// Each vector (one head of q or k for one token, C elements) is handled by LPV = C/8 lanes, 8 elements (16 B) per
// lane, so a warp processes 32/LPV vectors per pass with 16-byte loads/stores; reductions are sub-warp shuffles.
template <int LPV>
MG_DEVICE float group_sum(float v) {
#pragma unroll
  for (int o = LPV / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
MG_DEVICE void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
MG_DEVICE uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// Index algebra: vector v = m*2H + r with r = which*H + h, so the raw-projection offset is m*3HC + r*C and the output /
// in-place offset math needs one 32-bit division per vector.  Every lane keeps U independent vectors in flight (all
// global loads of a pass are issued before any arithmetic) -- with one vector per pass the kernels ran at ~45% of HBM
// bandwidth purely on memory-level parallelism.
template <int C, int U>
__global__ void __launch_bounds__(256, 2) qkln_rope_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ q_gain,
                                                               const float* __restrict__ k_gain, const float* __restrict__ sin_t,
                                                               const float* __restrict__ cos_t, __nv_bfloat16* __restrict__ qk,
                                                               float* __restrict__ stats, unsigned nvec, int T, int H, float eps) {
  constexpr int LPV = C / 8, VPW = 32 / LPV;
  const int lane = threadIdx.x & 31, sub = lane % LPV, slot = lane / LPV;
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned twoH = 2u * H;
  const int c0 = sub * 8;
  float gq[8], gk[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { gq[e] = q_gain[c0 + e]; gk[e] = k_gain[c0 + e]; }
  for (unsigned long long base = static_cast<unsigned long long>(warp0) * (VPW * U); base < nvec;
       base += static_cast<unsigned long long>(nwarps) * (VPW * U)) {
    uint4 raw[U];
    unsigned vm[U], vr[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long vec = base + u * VPW + slot;
      ok[u] = vec < nvec;
      const unsigned v = ok[u] ? static_cast<unsigned>(vec) : nvec - 1;
      vm[u] = v / twoH;
      vr[u] = v - vm[u] * twoH;
      raw[u] = __ldg(reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(vm[u]) * 3 * H + vr[u]) * C + c0));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool which = vr[u] >= static_cast<unsigned>(H);   // false = q, true = k
      const int pos = vm[u] % T;
      float x[8];
      unpack8(raw[u], x);
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += x[e];
      const float mu = group_sum<LPV>(s) / C;
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] -= mu; ss += x[e] * x[e]; }
      const float rs = rsqrtf(group_sum<LPV>(ss) / C + eps);
      const float4 sn = __ldg(reinterpret_cast<const float4*>(sin_t + pos * (C / 2) + c0 / 2));
      const float4 cs = __ldg(reinterpret_cast<const float4*>(cos_t + pos * (C / 2) + c0 / 2));
      const float sv[4] = {sn.x, sn.y, sn.z, sn.w}, cv[4] = {cs.x, cs.y, cs.z, cs.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float a = x[e] * rs * (which ? gk[e] : gq[e]), b = x[e + 1] * rs * (which ? gk[e + 1] : gq[e + 1]);
        o[e] = a * cv[e / 2] - b * sv[e / 2];
        o[e + 1] = b * cv[e / 2] + a * sv[e / 2];
      }
      if (ok[u]) {
        const size_t vec = static_cast<size_t>(vm[u]) * twoH + vr[u];
        *reinterpret_cast<uint4*>(qk + vec * C + c0) = pack8(o);
        if (sub == 0) *reinterpret_cast<float2*>(stats + vec * 2) = make_float2(mu, rs);
      }
    }
  }
}

// In place on dqkv [M, 3, H, C] (q and k thirds): d(roped, normalised) -> d(raw projection output).
// Gain gradients are accumulated per lane over the grid-stride loop, reduced through shared memory, then atomically added.
template <int C, int U>
__global__ void __launch_bounds__(256, 2) qkln_rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, const __nv_bfloat16* __restrict__ qkv_raw,
                                                               const float* __restrict__ stats, const float* __restrict__ q_gain,
                                                               const float* __restrict__ k_gain, const float* __restrict__ sin_t,
                                                               const float* __restrict__ cos_t, float* __restrict__ dq_gain,
                                                               float* __restrict__ dk_gain, unsigned nvec, int T, int H,
                                                               float* __restrict__ dq_acc) {
  // dq_acc != nullptr: the q gradients are read straight from the attention backward's fp32 accumulator [M, H*C] (and the
  // accumulator is cleared behind the read, ready for the next layer) instead of from the bf16 dq third of dqkv -- removes the
  // fp32->bf16 conversion pass and the zero-fill pass between the two kernels.
  constexpr int LPV = C / 8, VPW = 32 / LPV;
  __shared__ float red[2][C];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane % LPV, slot = lane / LPV;
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned twoH = 2u * H;
  const int c0 = sub * 8;
  float gq[8], gk[8], gacc[2][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { gq[e] = q_gain[c0 + e]; gk[e] = k_gain[c0 + e]; gacc[0][e] = gacc[1][e] = 0.f; }
  for (unsigned long long base = static_cast<unsigned long long>(warp0) * (VPW * U); base < nvec;
       base += static_cast<unsigned long long>(nwarps) * (VPW * U)) {
    uint4 graw[U], xraw[U];
    float2 st[U];
    unsigned vm[U], vr[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long vec = base + u * VPW + slot;
      ok[u] = vec < nvec;
      const unsigned v = ok[u] ? static_cast<unsigned>(vec) : nvec - 1;
      vm[u] = v / twoH;
      vr[u] = v - vm[u] * twoH;
      const size_t off = (static_cast<size_t>(vm[u]) * 3 * H + vr[u]) * C + c0;
      if (dq_acc != nullptr && vr[u] < static_cast<unsigned>(H)) {
        float4* ap = reinterpret_cast<float4*>(dq_acc + (static_cast<size_t>(vm[u]) * H + vr[u]) * C + c0);
        const float4 a0 = ap[0], a1 = ap[1];
        graw[u] = make_uint4(pack_bf16(a0.x, a0.y), pack_bf16(a0.z, a0.w), pack_bf16(a1.x, a1.y), pack_bf16(a1.z, a1.w));
        if (ok[u]) { ap[0] = make_float4(0.f, 0.f, 0.f, 0.f); ap[1] = make_float4(0.f, 0.f, 0.f, 0.f); }
      } else {
        graw[u] = *reinterpret_cast<const uint4*>(dqkv + off);
      }
      xraw[u] = __ldg(reinterpret_cast<const uint4*>(qkv_raw + off));
      st[u] = __ldg(reinterpret_cast<const float2*>(stats + static_cast<size_t>(v) * 2));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool which = vr[u] >= static_cast<unsigned>(H);
      const int pos = vm[u] % T;
      float g[8], x[8];
      unpack8(graw[u], g);
      unpack8(xraw[u], x);
      const float mu = st[u].x, rs = st[u].y;
      const float4 sn = __ldg(reinterpret_cast<const float4*>(sin_t + pos * (C / 2) + c0 / 2));
      const float4 cs = __ldg(reinterpret_cast<const float4*>(cos_t + pos * (C / 2) + c0 / 2));
      const float sv[4] = {sn.x, sn.y, sn.z, sn.w}, cv[4] = {cs.x, cs.y, cs.z, cs.w};
      float dxh[8], xh[8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float dn0 = g[e] * cv[e / 2] + g[e + 1] * sv[e / 2];        // inverse rotation
        const float dn1 = g[e + 1] * cv[e / 2] - g[e] * sv[e / 2];
        xh[e] = (x[e] - mu) * rs;
        xh[e + 1] = (x[e + 1] - mu) * rs;
        if (ok[u]) {
          if (which) { gacc[1][e] += dn0 * xh[e]; gacc[1][e + 1] += dn1 * xh[e + 1]; }
          else { gacc[0][e] += dn0 * xh[e]; gacc[0][e + 1] += dn1 * xh[e + 1]; }
        }
        dxh[e] = dn0 * (which ? gk[e] : gq[e]);
        dxh[e + 1] = dn1 * (which ? gk[e + 1] : gq[e + 1]);
        s1 += dxh[e] + dxh[e + 1];
        s2 += dxh[e] * xh[e] + dxh[e + 1] * xh[e + 1];
      }
      s1 = group_sum<LPV>(s1) / C;
      s2 = group_sum<LPV>(s2) / C;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs * (dxh[e] - s1 - xh[e] * s2);
      if (ok[u]) *reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(vm[u]) * 3 * H + vr[u]) * C + c0) = pack8(o);
    }
  }
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&red[w][c0 + e], gacc[w][e]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dq_gain + i, red[0][i]);
    atomicAdd(dk_gain + i, red[1][i]);
  }
}

void qkln_rope_fwd(const void* qkv, const float* qg, const float* kg, const float* sin_t, const float* cos_t, void* qk,
                   float* stats, long long M, int T, int H, int C, float eps, cudaStream_t s) {
  const long long nvec64 = M * 2 * H;
  if (nvec64 >= (1ll << 31)) throw std::runtime_error("qkln_rope: too many (token, head) vectors for 32-bit indexing");
  const unsigned nvec = static_cast<unsigned>(nvec64);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = static_cast<unsigned>(sms * 4);
  if (C == 64)
    qkln_rope_fwd_kernel<64, 4><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), qg, kg, sin_t, cos_t, static_cast<__nv_bfloat16*>(qk), stats, nvec, T, H, eps);
  else if (C == 128)
    qkln_rope_fwd_kernel<128, 4><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), qg, kg, sin_t, cos_t, static_cast<__nv_bfloat16*>(qk), stats, nvec, T, H, eps);
  else throw std::runtime_error("qkln_rope: head_dim must be 64 or 128");
  check_launch("qkln_rope_fwd");
}
void qkln_rope_bwd(void* dqkv, const void* qkv_raw, const float* stats, const float* qg, const float* kg, const float* sin_t,
                   const float* cos_t, float* dqg, float* dkg, long long M, int T, int H, int C, int num_sms, cudaStream_t s, float* dq_acc) {
  const long long nvec64 = M * 2 * H;
  if (nvec64 >= (1ll << 31)) throw std::runtime_error("qkln_rope: too many (token, head) vectors for 32-bit indexing");
  const unsigned nvec = static_cast<unsigned>(nvec64);
  const unsigned grid = static_cast<unsigned>(num_sms * 4);
  if (C == 64)
    qkln_rope_bwd_kernel<64, 4><<<grid, 256, 0, s>>>(static_cast<__nv_bfloat16*>(dqkv), static_cast<const __nv_bfloat16*>(qkv_raw), stats, qg, kg, sin_t, cos_t, dqg, dkg, nvec, T, H, dq_acc);
  else if (C == 128)
    qkln_rope_bwd_kernel<128, 4><<<grid, 256, 0, s>>>(static_cast<__nv_bfloat16*>(dqkv), static_cast<const __nv_bfloat16*>(qkv_raw), stats, qg, kg, sin_t, cos_t, dqg, dkg, nvec, T, H, dq_acc);
  else throw std::runtime_error("qkln_rope: head_dim must be 64 or 128");
  check_launch("qkln_rope_bwd");
}

// ------------------------------------------------------------------------------------------------
// Embedding gather and scatter-add
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const __nv_bfloat16* __restrict__ wte, const int* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                 long long M, int D) {
  const int nvec = D >> 3;
  const long long total = M * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / nvec;
    const int j = i - m * nvec;
    reinterpret_cast<uint4*>(out + m * D)[j] = __ldg(reinterpret_cast<const uint4*>(wte + static_cast<long long>(idx[m]) * D) + j);
  }
}
__global__ void embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ idx, float* __restrict__ dwte,
                                 long long M, int D) {
  const int nvec = D >> 2;  // 4 elements per thread-iteration
  const long long total = M * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / nvec;
    const int j = i - m * nvec;
    uint2 u = __ldg(reinterpret_cast<const uint2*>(dx + m * D) + j);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    float4* dst = reinterpret_cast<float4*>(dwte + static_cast<long long>(idx[m]) * D) + j;
    atomicAdd(dst, make_float4(a.x, a.y, b.x, b.y));
  }
}
void embed_fwd(const void* wte, const int* idx, void* out, long long M, int D, int num_sms, cudaStream_t s) {
  embed_fwd_kernel<<<num_sms * 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(wte), idx, static_cast<__nv_bfloat16*>(out), M, D);
  check_launch("embed_fwd");
}
void embed_bwd(const void* dx, const int* idx, float* dwte, long long M, int D, int num_sms, cudaStream_t s) {
  embed_bwd_kernel<<<num_sms * 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dx), idx, dwte, M, D);
  check_launch("embed_bwd");
}

// ------------------------------------------------------------------------------------------------
// Cross-entropy over a chunk of bf16 logits [rows, V] (row stride ldl): loss_sum += Σ(lse - logit[target]);
// optionally overwrite the logits in place with grad_scale * (softmax - onehot) (bf16) for the backward GEMMs.
// One 512-thread block per row; online (max, sum) in a single read, second pass served from L2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) ce_chunk_kernel(__nv_bfloat16* __restrict__ logits, long long ldl, const int* __restrict__ targets,
                                                       float* __restrict__ loss_sum, int V, float grad_scale, int write_grad) {
  __shared__ float s_max[16], s_sum[16];
  __shared__ float s_bcast[2];
  const long long row = blockIdx.x;
  __nv_bfloat16* lr = logits + row * ldl;
  const int nvec = V >> 3;
  float mx = -INFINITY, sm = 0.f;
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    uint4 u = *reinterpret_cast<const uint4*>(lr + j * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 t = unpack_bf16(w[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
    float lm = f[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) lm = fmaxf(lm, f[e]);
    const float nm = fmaxf(mx, lm);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += __expf(f[e] - nm);
    sm = sm * __expf(mx - nm) + acc;
    mx = nm;
  }
  for (int j = nvec * 8 + threadIdx.x; j < V; j += blockDim.x) {   // tail (V % 8)
    const float f = __bfloat162float(lr[j]);
    const float nm = fmaxf(mx, f);
    sm = sm * __expf(mx - nm) + __expf(f - nm);
    mx = nm;
  }
  // block reduce (max, sum)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wm = warp_max(mx);
  float ws = warp_sum(sm * ((mx == -INFINITY) ? 0.f : __expf(mx - wm)));
  if (lane == 0) { s_max[warp] = wm; s_sum[warp] = ws; }
  __syncthreads();
  if (warp == 0) {
    float m2 = lane < (blockDim.x >> 5) ? s_max[lane] : -INFINITY;
    float s2 = lane < (blockDim.x >> 5) ? s_sum[lane] : 0.f;
    const float bm = warp_max(m2);
    const float bs = warp_sum(s2 * ((m2 == -INFINITY) ? 0.f : __expf(m2 - bm)));
    if (lane == 0) { s_bcast[0] = bm; s_bcast[1] = bs; }
  }
  __syncthreads();
  const float bm = s_bcast[0], bs = s_bcast[1];
  const float lse = bm + __logf(bs);
  const int tgt = targets[row];
  if (threadIdx.x == 0) atomicAdd(loss_sum, lse - __bfloat162float(lr[tgt]));
  if (!write_grad) return;
  const float inv = 1.f / bs;
  __syncthreads();   // the target logit has been read by thread 0 before anyone overwrites it
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    uint4 u = *reinterpret_cast<const uint4*>(lr + j * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 t = unpack_bf16(w[e]);
      const int c = j * 8 + 2 * e;
      float g0 = __expf(t.x - bm) * inv - (c == tgt ? 1.f : 0.f);
      float g1 = __expf(t.y - bm) * inv - (c + 1 == tgt ? 1.f : 0.f);
      o[e] = pack_bf16(g0 * grad_scale, g1 * grad_scale);
    }
    *reinterpret_cast<uint4*>(lr + j * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (int j = nvec * 8 + threadIdx.x; j < V; j += blockDim.x) {
    const float f = __bfloat162float(lr[j]);
    lr[j] = __float2bfloat16((__expf(f - bm) * inv - (j == tgt ? 1.f : 0.f)) * grad_scale);
  }
}
void ce_chunk(void* logits, long long ldl, const int* targets, float* loss_sum, int rows, int V, float grad_scale,
              int write_grad, cudaStream_t s) {
  ce_chunk_kernel<<<rows, 512, 0, s>>>(static_cast<__nv_bfloat16*>(logits), ldl, targets, loss_sum, V, grad_scale, write_grad);
  check_launch("ce_chunk");
}

// ------------------------------------------------------------------------------------------------
// Dropout: y = x * keep / (1-p), keep ~ Bernoulli(1-p) from Philox4x32-10 keyed by (seed, element/4)
// ------------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4, float p,
                               unsigned long long seed) {
  const float scale = 1.f / (1.f - p);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 ctr = make_uint4(static_cast<unsigned>(i), static_cast<unsigned>(i >> 32), 0u, 0u);
    uint2 key = make_uint2(static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32));
    uint4 rnd = curand_Philox4x32_10(ctr, key);
    uint2 u = __ldg(reinterpret_cast<const uint2*>(x) + i);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    const float th = p * 4294967296.f;
    a.x = (static_cast<float>(rnd.x) >= th) ? a.x * scale : 0.f;
    a.y = (static_cast<float>(rnd.y) >= th) ? a.y * scale : 0.f;
    b.x = (static_cast<float>(rnd.z) >= th) ? b.x * scale : 0.f;
    b.y = (static_cast<float>(rnd.w) >= th) ? b.y * scale : 0.f;
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(a.x, a.y), pack_bf16(b.x, b.y));
  }
}
void dropout(const void* x, void* y, long long n, float p, unsigned long long seed, int num_sms, cudaStream_t s) {
  if (n % 4) throw std::runtime_error("dropout: numel must be a multiple of 4");
  dropout_kernel<<<num_sms * 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n / 4, p, seed);
  check_launch("dropout");
}

// ------------------------------------------------------------------------------------------------
// Row-wise e4m3 quantisation for the FP8 forward GEMMs: q[r,:] = round_e4m3(x[r,:] * 448/amax_r), dequant[r] = amax_r/448.
// One warp per row, the row stays in registers between the amax reduction and the conversion (single pass over HBM).
// ------------------------------------------------------------------------------------------------
MG_DEVICE uint32_t pack4_e4m3(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
template <int MAXV>
__global__ void __launch_bounds__(256) quant_rows_e4m3_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                              float* __restrict__ dequant, int R, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const int nvec = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * C);
  uint4 v[MAXV];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      v[i] = __ldg(xr + j);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
    }
  }
  amax = fmaxf(warp_max(amax), 1e-12f);
  const float scale = 448.f / amax;
  if (lane == 0) dequant[warp] = amax / 448.f;
  uint2* qr = reinterpret_cast<uint2*>(q + static_cast<size_t>(warp) * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      float f[8];
      unpack8(v[i], f);
      qr[j] = make_uint2(pack4_e4m3(f[0] * scale, f[1] * scale, f[2] * scale, f[3] * scale),
                         pack4_e4m3(f[4] * scale, f[5] * scale, f[6] * scale, f[7] * scale));
    }
  }
}
void quant_rows_e4m3(const void* x, void* q, float* dequant, int R, int C, cudaStream_t s) {
  if (C % 8) throw std::runtime_error("quant_rows_e4m3: C must be a multiple of 8");
  const int nvec = C / 8;
  auto launch = [&](auto mv) {
    constexpr int MAXV = decltype(mv)::value;
    quant_rows_e4m3_kernel<MAXV><<<(R + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(q), dequant, R, C);
  };
  if (nvec <= 32) launch(std::integral_constant<int, 1>());
  else if (nvec <= 64) launch(std::integral_constant<int, 2>());
  else if (nvec <= 128) launch(std::integral_constant<int, 4>());
  else if (nvec <= 256) launch(std::integral_constant<int, 8>());
  else if (nvec <= 512) launch(std::integral_constant<int, 16>());
  else if (nvec <= 1024) launch(std::integral_constant<int, 32>());
  else throw std::runtime_error("quant_rows_e4m3: row longer than 8192 elements");
  check_launch("quant_rows_e4m3");
}

// ------------------------------------------------------------------------------------------------
// MX (OCP micro-scaling) e4m3 quantisation: one UE8M0 scale per 32 consecutive elements of a row,
//   scale = 2^ceil(log2(amax_block / 448)),  q = round_e4m3(x / scale).
// Scales are written in the tcgen05 block-scale atom order: for every (128-row block, group of 4 k-blocks) a 512-byte
// tile with byte offset (row%32)*16 + (row/32 %4)*4 + (kblock%4)  — the order tcgen05.cp.32x128b.warpx4 expects.
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) quant_mx_e4m3_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                            uint8_t* __restrict__ sf, int R, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const int nvec = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * C);
  uint2* qr = reinterpret_cast<uint2*>(q + static_cast<size_t>(warp) * C);
  const int rb = warp >> 7, m = warp & 127, m0 = m & 31, m1 = m >> 5;
  const size_t kgroups = static_cast<size_t>(C) >> 7;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + i * 32;                 // 8-element vector index; 4 consecutive lanes = one 32-element block
    float f[8];
    float amax = 0.f;
    if (j < nvec) {
      unpack8(__ldg(xr + j), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    // exact ceil(log2(amax / 448)) from the float's exponent / mantissa (no transcendental approximations)
    const uint32_t vb = __float_as_uint(fmaxf(amax, 1e-30f) * (1.f / 448.f));
    int e8 = static_cast<int>((vb >> 23) & 0xFF) - 127 + ((vb & 0x7FFFFF) ? 1 : 0);
    e8 = max(-127, min(127, e8));
    const float inv = exp2f(static_cast<float>(-e8));
    if (j < nvec) {
      qr[j] = make_uint2(pack4_e4m3(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv), pack4_e4m3(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv));
      if ((lane & 3) == 0) {
        const int kblk = j >> 2;                 // 32-element block index along the row
        sf[(static_cast<size_t>(rb) * kgroups + (kblk >> 2)) * 512 + m0 * 16 + m1 * 4 + (kblk & 3)] = static_cast<uint8_t>(e8 + 127);
      }
    }
  }
}
void quant_mx_e4m3(const void* x, void* q, void* sf, int R, int C, cudaStream_t s) {
  if (C % 128) throw std::runtime_error("quant_mx_e4m3: C must be a multiple of 128");
  const int nvec = C / 8;
  auto launch = [&](auto mv) {
    constexpr int MAXV = decltype(mv)::value;
    quant_mx_e4m3_kernel<MAXV><<<(R + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(q), static_cast<uint8_t*>(sf), R, C);
  };
  if (nvec <= 32) launch(std::integral_constant<int, 1>());
  else if (nvec <= 64) launch(std::integral_constant<int, 2>());
  else if (nvec <= 128) launch(std::integral_constant<int, 4>());
  else if (nvec <= 256) launch(std::integral_constant<int, 8>());
  else if (nvec <= 512) launch(std::integral_constant<int, 16>());
  else if (nvec <= 1024) launch(std::integral_constant<int, 32>());
  else throw std::runtime_error("quant_mx_e4m3: row longer than 8192 elements");
  check_launch("quant_mx_e4m3");
}

// ------------------------------------------------------------------------------------------------
// Optimizer: Σ g² and the fused clip · Adam · independent-WD · LR · bf16 cast sweep over one fp32 shard
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  __shared__ float part[16];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}
void sumsq_accumulate(const float* g, long long n, float* out, int num_sms, cudaStream_t s) {
  if (n % 4) throw std::runtime_error("sumsq: numel must be a multiple of 4");
  sumsq_kernel<<<num_sms * 2, 512, 0, s>>>(g, n / 4, out);
  check_launch("sumsq");
}

struct AdamArgs {
  float lr, b1, b2, eps, wd_over_lr_peak, inv_bc1, inv_bc2;
};

// clip = *clip_ptr (device scalar: min(1, 1/||g||)), so the host never waits for the norm.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                    float* __restrict__ g, __nv_bfloat16* __restrict__ p_lp, long long n4,
                                                    const float* __restrict__ clip_ptr, AdamArgs a, int zero_grad) {
  const float clip = *clip_ptr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, vv_[4] = {vv.x, vv.y, vv.z, vv.w};
    const float gg[4] = {gv.x * clip, gv.y * clip, gv.z * clip, gv.w * clip};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = a.b1 * mm[e] + (1.f - a.b1) * gg[e];
      vv_[e] = a.b2 * vv_[e] + (1.f - a.b2) * gg[e] * gg[e];
      const float u = (mm[e] * a.inv_bc1) / (sqrtf(vv_[e] * a.inv_bc2) + a.eps) + a.wd_over_lr_peak * pp[e];
      pp[e] -= a.lr * u;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv_[0], vv_[1], vv_[2], vv_[3]);
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p_lp) reinterpret_cast<uint2*>(p_lp)[i] = make_uint2(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]));
  }
}
void adamw_step(float* p, float* m, float* v, float* g, void* p_lp, long long n, const float* clip_ptr, float lr, float b1,
                float b2, float eps, float wd_over_lr_peak, float bc1, float bc2, int zero_grad, int num_sms, cudaStream_t s) {
  if (n % 4) throw std::runtime_error("adamw: numel must be a multiple of 4");
  AdamArgs a{lr, b1, b2, eps, wd_over_lr_peak, 1.f / bc1, 1.f / bc2};
  adamw_kernel<<<num_sms * 8, 256, 0, s>>>(p, m, v, g, static_cast<__nv_bfloat16*>(p_lp), n / 4, clip_ptr, a, zero_grad);
  check_launch("adamw");
}

}  // namespace mg
