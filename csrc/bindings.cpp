// Python bindings (torch extension).  Thin: argument checking + stream lookup; kernels live in the .cu files.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

namespace mg {
std::vector<int> gemm_work_plan(int M, int N, int K, int epi, int accumulate, int num_sms, bool pairs);
void gemm_launch(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d, long long ldd,
                 void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K, int epi, int accumulate,
                 int num_sms, cudaStream_t stream);
void gemm_launch_ag_raw(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, void* d, long long ldd, void* d2,
                        long long ldd2, const void* aux, long long ld_aux, int M, int N, int K, int epi, int accumulate, int num_sms,
                        cudaStream_t stream, const long long* peer_ptrs, void* local, unsigned int* flags, long long chunk_elems,
                        unsigned int target, int world, int rank);
void gemm_fp8_launch(const void* a, long long lda, const void* b, long long ldb, const float* sa, const float* sb, void* d, long long ldd,
                     void* d2, long long ldd2, const void* aux, long long ld_aux, int M, int N, int K, int epi, int num_sms,
                     cudaStream_t stream);
void gemm_mx_launch(const void* a, const void* sfa, const void* b, const void* sfb, void* d, long long ldd, void* d2, long long ldd2,
                    const void* aux, long long ld_aux, int M, int N, int K, int epi, int num_sms, cudaStream_t stream);
void quant_rows_e4m3(const void* x, void* q, float* dequant, int R, int C, cudaStream_t s);
void quant_mx_e4m3(const void* x, void* q, void* sf, int R, int C, cudaStream_t s);
void rmsnorm_fwd(const void* x, void* y, float* rstd, int M, int D, float eps, cudaStream_t s);
void rmsnorm_bwd(const void* dy, const void* x, const float* rstd, const void* dres, void* dx, int M, int D, cudaStream_t s);
void qkln_rope_fwd(const void* qkv, const float* qg, const float* kg, const float* sin_t, const float* cos_t, void* qk,
                   float* stats, long long M, int T, int H, int C, float eps, cudaStream_t s);
void qkln_rope_bwd(void* dqkv, const void* qkv_raw, const float* stats, const float* qg, const float* kg, const float* sin_t,
                   const float* cos_t, float* dqg, float* dkg, long long M, int T, int H, int C, int num_sms, cudaStream_t s, float* dq_acc);
void embed_fwd(const void* wte, const int* idx, void* out, long long M, int D, int num_sms, cudaStream_t s);
void embed_bwd(const void* dx, const int* idx, float* dwte, long long M, int D, int num_sms, cudaStream_t s);
void ce_chunk(void* logits, long long ldl, const int* targets, float* loss_sum, int rows, int V, float grad_scale,
              int write_grad, cudaStream_t s);
void dropout(const void* x, void* y, long long n, float p, unsigned long long seed, int num_sms, cudaStream_t s);
void sumsq_accumulate(const float* g, long long n, float* out, int num_sms, cudaStream_t s);
void adamw_step(float* p, float* m, float* v, float* g, void* p_lp, long long n, const float* clip_ptr, float lr, float b1,
                float b2, float eps, float wd_over_lr_peak, float bc1, float bc2, int zero_grad, int num_sms, cudaStream_t s);
void attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int T, int H, int C, long long q_row_stride,
              long long kv_row_stride_k, long long kv_row_stride_v, float p_drop, unsigned long long seed, int num_sms, cudaStream_t s);
void attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, float* delta,
              void* dq, void* dk, void* dv, float* dq_accum, int B, int T, int H, int C, long long q_row_stride,
              long long k_row_stride, long long v_row_stride, long long dqkv_row_stride, float p_drop, unsigned long long seed,
              int num_sms, cudaStream_t s);
void xrank_barrier(const long long* pad_ptrs, int rank, int world, int slot, unsigned epoch, cudaStream_t s);
void rs_sumsq(const long long* grad_ptrs, long long mc_ptr, float* out_shard, float* sumsq, long long shard_off, long long n, int rank,
              int world, int accumulate, int want_sumsq, int num_sms, cudaStream_t s);
void adamw_allgather(float* p, float* m, float* v, const float* g, const long long* lp_ptrs, long long mc_ptr, long long shard_off,
                     long long n, const float* clip_ptr, float lr, float b1, float b2, float eps, float wd_over_lr_peak, float bc1,
                     float bc2, int world, int num_sms, cudaStream_t s);
void ag_pull(const long long* lp_ptrs, void* local, long long shard_elems, int rank, int world, unsigned* flags, int n_ctas, cudaStream_t s);
}  // namespace mg

namespace {

int num_sms_of(const at::Tensor& t) {
  static int cached[16] = {0};
  const int dev = t.get_device();
  if (!cached[dev]) cached[dev] = at::cuda::getDeviceProperties(dev)->multiProcessorCount;
  return cached[dev];
}
cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }

#define CHECK_BF16(t) TORCH_CHECK((t).is_cuda() && (t).scalar_type() == at::kBFloat16, #t " must be a CUDA bf16 tensor")
#define CHECK_F32(t) TORCH_CHECK((t).is_cuda() && (t).scalar_type() == at::kFloat, #t " must be a CUDA fp32 tensor")
#define CHECK_2D_ROWMAJOR(t) TORCH_CHECK((t).dim() == 2 && (t).stride(1) == 1, #t " must be 2-D with unit inner stride")

// d (+)= op(a) · op(b)ᵀ ; see csrc/gemm.cu for operand conventions.  With `gather` the same GEMM also carries the all-gather of an
// FSDP unit (all-gather ⊕ GEMM): the remote chunks of `local` are pulled from the owners' symmetric shards by copy warps inside
// the GEMM's own CTAs; if b is a view into `local` the kernel synchronises with them per 64 KB segment.
struct GatherArgs {
  std::vector<long long> peers;
  at::Tensor local, flags;
  int64_t chunk_elems, target, rank;
};

static void gemm_impl(const at::Tensor& a, bool a_mn, const at::Tensor& b, bool b_mn, at::Tensor& d, int64_t epi,
                      const c10::optional<at::Tensor>& d2, const c10::optional<at::Tensor>& aux, bool accumulate, const GatherArgs* ga) {
  c10::cuda::CUDAGuard guard(a.device());
  CHECK_BF16(a); CHECK_BF16(b); CHECK_2D_ROWMAJOR(a); CHECK_2D_ROWMAJOR(b); CHECK_2D_ROWMAJOR(d);
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb, "gemm: inner dimensions differ: ", K, " vs ", Kb);
  TORCH_CHECK(d.size(0) == M && d.size(1) == N, "gemm: output must be [", M, ", ", N, "]");
  TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "gemm: leading dimensions must be multiples of 8 elements (16 B)");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) % 16 == 0) && (reinterpret_cast<uintptr_t>(b.data_ptr()) % 16 == 0) &&
              (reinterpret_cast<uintptr_t>(d.data_ptr()) % 16 == 0), "gemm: base pointers must be 16-byte aligned");
  if (epi == 4) { CHECK_F32(d); TORCH_CHECK(d.stride(0) % 4 == 0); }
  else { CHECK_BF16(d); TORCH_CHECK(d.stride(0) % 8 == 0); }
  const void* auxp = nullptr; long long ld_aux = 0;
  if (epi == 1 || epi == 3) {
    TORCH_CHECK(aux.has_value(), "gemm: this epilogue needs aux");
    CHECK_BF16(*aux); CHECK_2D_ROWMAJOR(*aux);
    TORCH_CHECK(aux->size(0) == M && aux->size(1) == N && N % 64 == 0 && aux->stride(0) % 8 == 0);
    auxp = aux->data_ptr(); ld_aux = aux->stride(0);
  }
  void* d2p = nullptr; long long ldd2 = 0;
  if (epi == 2) {
    TORCH_CHECK(d2.has_value(), "gemm: GELU epilogue needs d2");
    CHECK_BF16(*d2); CHECK_2D_ROWMAJOR(*d2);
    TORCH_CHECK(d2->size(0) == M && d2->size(1) == N);
    d2p = d2->data_ptr(); ldd2 = d2->stride(0);
  }
  if (ga == nullptr) {
    mg::gemm_launch(a.data_ptr(), a.stride(0), a_mn, b.data_ptr(), b.stride(0), b_mn, d.data_ptr(), d.stride(0), d2p, ldd2, auxp,
                    ld_aux, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), static_cast<int>(epi), accumulate ? 1 : 0,
                    num_sms_of(a), stream());
    return;
  }
  CHECK_BF16(ga->local);
  const int64_t world = static_cast<int64_t>(ga->peers.size());
  TORCH_CHECK(ga->local.is_contiguous() && ga->flags.scalar_type() == at::kInt && ga->flags.is_contiguous());
  TORCH_CHECK(ga->chunk_elems % 8 == 0 && world * ga->chunk_elems <= ga->local.numel(), "gemm_ag: unit buffer too small");
  TORCH_CHECK(ga->flags.numel() >= (world * ga->chunk_elems + 32767) / 32768 + 16,
              "gemm_ag: flags needs one word per 32768-element segment (+16 for the stand-alone pull)");
  mg::gemm_launch_ag_raw(a.data_ptr(), a.stride(0), a_mn, b.data_ptr(), b.stride(0), b_mn, d.data_ptr(), d.stride(0), d2p, ldd2, auxp,
                         ld_aux, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), static_cast<int>(epi), accumulate ? 1 : 0,
                         num_sms_of(a), stream(), ga->peers.data(), ga->local.data_ptr(),
                         reinterpret_cast<unsigned int*>(ga->flags.data_ptr()), ga->chunk_elems, static_cast<unsigned int>(ga->target),
                         static_cast<int>(world), static_cast<int>(ga->rank));
}

void gemm(const at::Tensor& a, bool a_mn, const at::Tensor& b, bool b_mn, at::Tensor& d, int64_t epi,
          const c10::optional<at::Tensor>& d2, const c10::optional<at::Tensor>& aux, bool accumulate) {
  gemm_impl(a, a_mn, b, b_mn, d, epi, d2, aux, accumulate, nullptr);
}

void gemm_ag(const at::Tensor& a, bool a_mn, const at::Tensor& b, bool b_mn, at::Tensor& d, int64_t epi,
             const c10::optional<at::Tensor>& d2, const c10::optional<at::Tensor>& aux, bool accumulate,
             std::vector<int64_t> peer_ptrs, at::Tensor& local, at::Tensor& flags, int64_t chunk_elems, int64_t target, int64_t rank) {
  GatherArgs ga{std::vector<long long>(peer_ptrs.begin(), peer_ptrs.end()), local, flags, chunk_elems, target, rank};
  gemm_impl(a, a_mn, b, b_mn, d, epi, d2, aux, accumulate, &ga);
}

// row-wise e4m3 quantisation: returns (q uint8 [R,C], dequant fp32 [R])
std::vector<at::Tensor> quant_rows_e4m3(const at::Tensor& x) {
  c10::cuda::CUDAGuard guard(x.device());
  CHECK_BF16(x); TORCH_CHECK(x.dim() == 2 && x.is_contiguous());
  auto q = at::empty(x.sizes(), x.options().dtype(at::kByte));
  auto dq = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  mg::quant_rows_e4m3(x.data_ptr(), q.data_ptr(), dq.data_ptr<float>(), x.size(0), x.size(1), stream());
  return {q, dq};
}
// MX e4m3: returns (q uint8 [R,C], sf uint8 [ceil(R/128), C/128, 512]) — UE8M0 scale per 32 elements, tcgen05 atom order
std::vector<at::Tensor> quant_mx_e4m3(const at::Tensor& x) {
  c10::cuda::CUDAGuard guard(x.device());
  CHECK_BF16(x); TORCH_CHECK(x.dim() == 2 && x.is_contiguous() && x.size(1) % 128 == 0);
  auto q = at::empty(x.sizes(), x.options().dtype(at::kByte));
  auto sf = at::full({(x.size(0) + 127) / 128, x.size(1) / 128, 512}, 127, x.options().dtype(at::kByte));
  mg::quant_mx_e4m3(x.data_ptr(), q.data_ptr(), sf.data_ptr(), x.size(0), x.size(1), stream());
  return {q, sf};
}
// MXFP8 block-scaled GEMM: d = dequant(a_q, sfa) · dequant(b_q, sfb)ᵀ, scales applied by tcgen05 block_scale MMAs
void gemm_mx(const at::Tensor& a, const at::Tensor& sfa, const at::Tensor& b, const at::Tensor& sfb, at::Tensor& d, int64_t epi,
             const c10::optional<at::Tensor>& d2, const c10::optional<at::Tensor>& aux) {
  c10::cuda::CUDAGuard guard(a.device());
  TORCH_CHECK(a.scalar_type() == at::kByte && b.scalar_type() == at::kByte && a.is_contiguous() && b.is_contiguous());
  TORCH_CHECK(sfa.scalar_type() == at::kByte && sfb.scalar_type() == at::kByte && sfa.is_contiguous() && sfb.is_contiguous());
  CHECK_BF16(d); CHECK_2D_ROWMAJOR(d);
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K && d.size(0) == M && d.size(1) == N && d.stride(0) % 8 == 0);
  TORCH_CHECK(sfa.numel() == ((M + 127) / 128) * (K / 128) * 512 && sfb.numel() == ((N + 127) / 128) * (K / 128) * 512, "gemm_mx: scale tensor size");
  const void* auxp = nullptr; long long ld_aux = 0;
  if (epi == 1) { TORCH_CHECK(aux.has_value()); CHECK_BF16(*aux); auxp = aux->data_ptr(); ld_aux = aux->stride(0); }
  void* d2p = nullptr; long long ldd2 = 0;
  if (epi == 2) { TORCH_CHECK(d2.has_value()); CHECK_BF16(*d2); d2p = d2->data_ptr(); ldd2 = d2->stride(0); }
  mg::gemm_mx_launch(a.data_ptr(), sfa.data_ptr(), b.data_ptr(), sfb.data_ptr(), d.data_ptr(), d.stride(0), d2p, ldd2, auxp, ld_aux,
                     static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), static_cast<int>(epi), num_sms_of(a), stream());
}
// d = (a_q · b_qᵀ) ∘ (sa ⊗ sb) (+ epilogue), a_q [M,K] / b_q [N,K] e4m3 bytes
void gemm_fp8(const at::Tensor& a, const at::Tensor& sa, const at::Tensor& b, const at::Tensor& sb, at::Tensor& d, int64_t epi,
              const c10::optional<at::Tensor>& d2, const c10::optional<at::Tensor>& aux) {
  c10::cuda::CUDAGuard guard(a.device());
  TORCH_CHECK(a.scalar_type() == at::kByte && b.scalar_type() == at::kByte && a.is_contiguous() && b.is_contiguous());
  CHECK_F32(sa); CHECK_F32(sb); CHECK_BF16(d); CHECK_2D_ROWMAJOR(d);
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K && d.size(0) == M && d.size(1) == N && sa.numel() == M && sb.numel() == N);
  TORCH_CHECK(K % 16 == 0 && N % 64 == 0, "gemm_fp8: K must be a multiple of 16 and N of 64");
  const void* auxp = nullptr; long long ld_aux = 0;
  if (epi == 1) { TORCH_CHECK(aux.has_value()); CHECK_BF16(*aux); auxp = aux->data_ptr(); ld_aux = aux->stride(0); }
  void* d2p = nullptr; long long ldd2 = 0;
  if (epi == 2) { TORCH_CHECK(d2.has_value()); CHECK_BF16(*d2); d2p = d2->data_ptr(); ldd2 = d2->stride(0); }
  mg::gemm_fp8_launch(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), sa.data_ptr<float>(), sb.data_ptr<float>(), d.data_ptr(),
                      d.stride(0), d2p, ldd2, auxp, ld_aux, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K),
                      static_cast<int>(epi), num_sms_of(a), stream());
}

std::vector<at::Tensor> rmsnorm_fwd(const at::Tensor& x, double eps) {
  c10::cuda::CUDAGuard guard(x.device());
  CHECK_BF16(x); TORCH_CHECK(x.dim() == 2 && x.is_contiguous());
  auto y = at::empty_like(x);
  auto rstd = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  mg::rmsnorm_fwd(x.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(), x.size(0), x.size(1), static_cast<float>(eps), stream());
  return {y, rstd};
}
at::Tensor rmsnorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& rstd, const c10::optional<at::Tensor>& dres) {
  c10::cuda::CUDAGuard guard(x.device());
  CHECK_BF16(dy); CHECK_BF16(x); CHECK_F32(rstd);
  TORCH_CHECK(dy.is_contiguous() && x.is_contiguous() && dy.sizes() == x.sizes());
  if (dres.has_value()) { CHECK_BF16(*dres); TORCH_CHECK(dres->is_contiguous() && dres->sizes() == x.sizes()); }
  auto dx = at::empty_like(x);
  mg::rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), rstd.data_ptr<float>(), dres.has_value() ? dres->data_ptr() : nullptr, dx.data_ptr(),
                  x.size(0), x.size(1), stream());
  return dx;
}

std::vector<at::Tensor> qkln_rope_fwd(const at::Tensor& qkv, const at::Tensor& qg, const at::Tensor& kg, const at::Tensor& sin_t,
                                      const at::Tensor& cos_t, int64_t T, int64_t H, int64_t C, double eps) {
  c10::cuda::CUDAGuard guard(qkv.device());
  CHECK_BF16(qkv); CHECK_F32(qg); CHECK_F32(kg); CHECK_F32(sin_t); CHECK_F32(cos_t);
  TORCH_CHECK(qkv.is_contiguous() && qkv.dim() == 2 && qkv.size(1) == 3 * H * C);
  TORCH_CHECK(sin_t.is_contiguous() && sin_t.size(0) >= T && sin_t.size(1) == C / 2);
  const int64_t M = qkv.size(0);
  auto qk = at::empty({M, 2 * H * C}, qkv.options());
  auto stats = at::empty({M, 2, H, 2}, qkv.options().dtype(at::kFloat));
  mg::qkln_rope_fwd(qkv.data_ptr(), qg.data_ptr<float>(), kg.data_ptr<float>(), sin_t.data_ptr<float>(), cos_t.data_ptr<float>(),
                    qk.data_ptr(), stats.data_ptr<float>(), M, T, H, C, static_cast<float>(eps), stream());
  return {qk, stats};
}
void qkln_rope_bwd(at::Tensor& dqkv, const at::Tensor& qkv_raw, const at::Tensor& stats, const at::Tensor& qg, const at::Tensor& kg,
                   const at::Tensor& sin_t, const at::Tensor& cos_t, at::Tensor& dqg, at::Tensor& dkg, int64_t T, int64_t H, int64_t C,
                   const c10::optional<at::Tensor>& dq_acc) {
  c10::cuda::CUDAGuard guard(dqkv.device());
  float* accp = nullptr;
  if (dq_acc.has_value()) {
    CHECK_F32(*dq_acc);
    TORCH_CHECK(dq_acc->is_contiguous() && dq_acc->numel() == dqkv.size(0) * H * C, "qkln_rope_bwd: dq_acc must be [M, H*C] fp32");
    accp = dq_acc->data_ptr<float>();
  }
  CHECK_BF16(dqkv); CHECK_BF16(qkv_raw); CHECK_F32(stats); CHECK_F32(qg); CHECK_F32(kg); CHECK_F32(dqg); CHECK_F32(dkg);
  TORCH_CHECK(dqkv.is_contiguous() && qkv_raw.is_contiguous() && dqkv.sizes() == qkv_raw.sizes());
  mg::qkln_rope_bwd(dqkv.data_ptr(), qkv_raw.data_ptr(), stats.data_ptr<float>(), qg.data_ptr<float>(), kg.data_ptr<float>(),
                    sin_t.data_ptr<float>(), cos_t.data_ptr<float>(), dqg.data_ptr<float>(), dkg.data_ptr<float>(), dqkv.size(0), T, H, C,
                    num_sms_of(dqkv), stream(), accp);
}

at::Tensor embed_fwd(const at::Tensor& wte, const at::Tensor& idx) {
  c10::cuda::CUDAGuard guard(wte.device());
  CHECK_BF16(wte); TORCH_CHECK(idx.is_cuda() && idx.scalar_type() == at::kInt && idx.is_contiguous());
  TORCH_CHECK(wte.is_contiguous() && wte.size(1) % 8 == 0);
  auto out = at::empty({idx.numel(), wte.size(1)}, wte.options());
  mg::embed_fwd(wte.data_ptr(), idx.data_ptr<int>(), out.data_ptr(), idx.numel(), wte.size(1), num_sms_of(wte), stream());
  return out;
}
void embed_bwd(const at::Tensor& dx, const at::Tensor& idx, at::Tensor& dwte) {
  c10::cuda::CUDAGuard guard(dx.device());
  CHECK_BF16(dx); CHECK_F32(dwte); TORCH_CHECK(idx.scalar_type() == at::kInt && idx.is_contiguous() && dx.is_contiguous() && dwte.is_contiguous());
  mg::embed_bwd(dx.data_ptr(), idx.data_ptr<int>(), dwte.data_ptr<float>(), idx.numel(), dx.size(1), num_sms_of(dx), stream());
}

void ce_chunk(at::Tensor& logits, const at::Tensor& targets, at::Tensor& loss_sum, int64_t V, double grad_scale, bool write_grad) {
  c10::cuda::CUDAGuard guard(logits.device());
  CHECK_BF16(logits); CHECK_F32(loss_sum); CHECK_2D_ROWMAJOR(logits);
  TORCH_CHECK(targets.scalar_type() == at::kInt && targets.is_contiguous() && targets.numel() == logits.size(0));
  TORCH_CHECK(logits.stride(0) % 8 == 0 && V <= logits.size(1));
  mg::ce_chunk(logits.data_ptr(), logits.stride(0), targets.data_ptr<int>(), loss_sum.data_ptr<float>(), logits.size(0), V,
               static_cast<float>(grad_scale), write_grad ? 1 : 0, stream());
}

at::Tensor dropout(const at::Tensor& x, double p, int64_t seed) {
  c10::cuda::CUDAGuard guard(x.device());
  CHECK_BF16(x); TORCH_CHECK(x.is_contiguous());
  auto y = at::empty_like(x);
  mg::dropout(x.data_ptr(), y.data_ptr(), x.numel(), static_cast<float>(p), static_cast<unsigned long long>(seed), num_sms_of(x), stream());
  return y;
}

void sumsq_accumulate(const at::Tensor& g, at::Tensor& out) {
  c10::cuda::CUDAGuard guard(g.device());
  CHECK_F32(g); CHECK_F32(out); TORCH_CHECK(g.is_contiguous());
  mg::sumsq_accumulate(g.data_ptr<float>(), g.numel(), out.data_ptr<float>(), num_sms_of(g), stream());
}

void adamw_step(at::Tensor& p, at::Tensor& m, at::Tensor& v, at::Tensor& g, const c10::optional<at::Tensor>& p_lp, const at::Tensor& clip,
                double lr, double b1, double b2, double eps, double wd_over_lr_peak, double bc1, double bc2, bool zero_grad) {
  c10::cuda::CUDAGuard guard(p.device());
  CHECK_F32(p); CHECK_F32(m); CHECK_F32(v); CHECK_F32(g); CHECK_F32(clip);
  TORCH_CHECK(p.is_contiguous() && m.is_contiguous() && v.is_contiguous() && g.is_contiguous());
  TORCH_CHECK(p.numel() == m.numel() && p.numel() == v.numel() && p.numel() == g.numel());
  void* lp = nullptr;
  if (p_lp.has_value()) { CHECK_BF16(*p_lp); TORCH_CHECK(p_lp->is_contiguous() && p_lp->numel() == p.numel()); lp = p_lp->data_ptr(); }
  mg::adamw_step(p.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), g.data_ptr<float>(), lp, p.numel(), clip.data_ptr<float>(),
                 static_cast<float>(lr), static_cast<float>(b1), static_cast<float>(b2), static_cast<float>(eps),
                 static_cast<float>(wd_over_lr_peak), static_cast<float>(bc1), static_cast<float>(bc2), zero_grad ? 1 : 0, num_sms_of(p), stream());
}

// q,k: [B,T,H,C] views with arbitrary row (token) stride; v likewise; returns (o [B*T, H*C], lse [B,H,T]).
// dropout_p > 0: attention-probability dropout inside the kernel (Philox keyed by seed; the backward regenerates the mask).
std::vector<at::Tensor> attn_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, double dropout_p, int64_t seed) {
  c10::cuda::CUDAGuard guard(q.device());
  CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v);
  TORCH_CHECK(q.dim() == 4 && q.stride(3) == 1 && q.stride(2) == q.size(3), "q must be [B,T,H,C] with packed heads");
  TORCH_CHECK(k.stride(3) == 1 && k.stride(2) == k.size(3) && v.stride(3) == 1 && v.stride(2) == v.size(3));
  const int64_t B = q.size(0), T = q.size(1), H = q.size(2), C = q.size(3);
  TORCH_CHECK(q.stride(0) == T * q.stride(1) && k.stride(0) == T * k.stride(1) && v.stride(0) == T * v.stride(1));
  auto o = at::empty({B * T, H * C}, q.options());
  auto lse = at::empty({B, H, T}, q.options().dtype(at::kFloat));
  mg::attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr<float>(), B, T, H, C, q.stride(1), k.stride(1),
               v.stride(1), static_cast<float>(dropout_p), static_cast<unsigned long long>(seed), num_sms_of(q), stream());
  return {o, lse};
}

// dqkv: [B*T, 3*H*C] bf16 output (dq | dk | dv thirds).
// dq_acc (optional): a ZEROED fp32 [B*T, H*C] accumulator owned by the caller.  With it the kernel leaves dq there (the dq third of
// the returned tensor is not written) and launches neither the zero-fill nor the fp32->bf16 conversion: qkln_rope_bwd consumes and
// re-zeroes the accumulator.
at::Tensor attn_bwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& o, const at::Tensor& dout,
                    const at::Tensor& lse, double dropout_p, int64_t seed, const c10::optional<at::Tensor>& dq_acc) {
  c10::cuda::CUDAGuard guard(q.device());
  CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(o); CHECK_BF16(dout); CHECK_F32(lse);
  const int64_t B = q.size(0), T = q.size(1), H = q.size(2), C = q.size(3);
  TORCH_CHECK(o.is_contiguous() && dout.is_contiguous());
  auto dqkv = at::empty({B * T, 3 * H * C}, q.options());
  auto delta = at::empty({3, B, H, T}, lse.options());   // delta | delta*scale | lse*log2(e)
  at::Tensor dq_accum;
  if (dq_acc.has_value()) {
    CHECK_F32(*dq_acc);
    TORCH_CHECK(dq_acc->is_contiguous() && dq_acc->numel() == B * T * H * C, "attn_bwd: dq_acc must be [B*T, H*C] fp32");
    dq_accum = *dq_acc;
  } else {
    dq_accum = at::zeros({B * T, H * C}, lse.options());
  }
  char* base = static_cast<char*>(dqkv.data_ptr());
  mg::attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(), lse.data_ptr<float>(), delta.data_ptr<float>(),
               dq_acc.has_value() ? nullptr : base, base + H * C * 2, base + 2 * H * C * 2, dq_accum.data_ptr<float>(), B, T, H, C, q.stride(1), k.stride(1), v.stride(1),
               3 * H * C, static_cast<float>(dropout_p), static_cast<unsigned long long>(seed), num_sms_of(q), stream());
  return dqkv;
}

// ---- in-kernel NVLink communication (csrc/comm.cu).  Pointer lists are raw device addresses of the peers' symmetric buffers.
void xrank_barrier(std::vector<int64_t> pads, int64_t rank, int64_t slot, int64_t epoch) {
  std::vector<long long> p(pads.begin(), pads.end());
  mg::xrank_barrier(p.data(), rank, static_cast<int>(p.size()), slot, static_cast<unsigned>(epoch), stream());
}
void rs_sumsq(std::vector<int64_t> grad_ptrs, int64_t mc_ptr, at::Tensor& out_shard, at::Tensor& sumsq, int64_t shard_off, int64_t rank,
              bool accumulate, bool want_sumsq) {
  c10::cuda::CUDAGuard guard(out_shard.device());
  CHECK_F32(out_shard); CHECK_F32(sumsq);
  std::vector<long long> p(grad_ptrs.begin(), grad_ptrs.end());
  mg::rs_sumsq(p.data(), mc_ptr, out_shard.data_ptr<float>(), sumsq.data_ptr<float>(), shard_off, out_shard.numel(), rank,
               static_cast<int>(p.size()), accumulate ? 1 : 0, want_sumsq ? 1 : 0, num_sms_of(out_shard), stream());
}
void adamw_allgather(at::Tensor& p, at::Tensor& m, at::Tensor& v, const at::Tensor& g, std::vector<int64_t> lp_ptrs, int64_t mc_ptr,
                     int64_t shard_off, const at::Tensor& clip, double lr, double b1, double b2, double eps, double wd_over_lr_peak,
                     double bc1, double bc2) {
  c10::cuda::CUDAGuard guard(p.device());
  CHECK_F32(p); CHECK_F32(m); CHECK_F32(v); CHECK_F32(g); CHECK_F32(clip);
  std::vector<long long> pp(lp_ptrs.begin(), lp_ptrs.end());
  mg::adamw_allgather(p.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), g.data_ptr<float>(), pp.data(), mc_ptr, shard_off,
                      p.numel(), clip.data_ptr<float>(), static_cast<float>(lr), static_cast<float>(b1), static_cast<float>(b2),
                      static_cast<float>(eps), static_cast<float>(wd_over_lr_peak), static_cast<float>(bc1), static_cast<float>(bc2),
                      static_cast<int>(pp.size()), num_sms_of(p), stream());
}
void ag_pull(std::vector<int64_t> lp_ptrs, at::Tensor& local, int64_t shard_elems, int64_t rank, at::Tensor& flags, int64_t n_ctas) {
  c10::cuda::CUDAGuard guard(local.device());
  CHECK_BF16(local);
  std::vector<long long> pp(lp_ptrs.begin(), lp_ptrs.end());
  mg::ag_pull(pp.data(), local.data_ptr(), shard_elems, rank, static_cast<int>(pp.size()), reinterpret_cast<unsigned*>(flags.data_ptr()),
              n_ctas, stream());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm, "tcgen05 bf16 GEMM with fused epilogues");
  m.def("gemm_work_plan", [](int M, int N, int K, int epi, int accumulate, int num_sms, bool pairs) { return mg::gemm_work_plan(M, N, K, epi, accumulate, num_sms, pairs); },
        "host replay of the GEMM tile / K-slice schedule: flat list of (worker, tile, kb0, kb1)");
  m.def("gemm_ag", &gemm_ag, "tcgen05 GEMM that carries the NVLink all-gather of an FSDP unit (copy warps in every CTA)");
  m.def("quant_rows_e4m3", &quant_rows_e4m3);
  m.def("quant_mx_e4m3", &quant_mx_e4m3);
  m.def("gemm_mx", &gemm_mx, "MXFP8 block-scaled GEMM (tcgen05 kind::mxf8f6f4.block_scale)");
  m.def("gemm_fp8", &gemm_fp8, "tcgen05 kind::f8f6f4 GEMM on row-scaled e4m3 operands");
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("qkln_rope_fwd", &qkln_rope_fwd);
  m.def("qkln_rope_bwd", &qkln_rope_bwd);
  m.def("embed_fwd", &embed_fwd);
  m.def("embed_bwd", &embed_bwd);
  m.def("ce_chunk", &ce_chunk);
  m.def("dropout", &dropout);
  m.def("sumsq_accumulate", &sumsq_accumulate);
  m.def("adamw_step", &adamw_step);
  m.def("attn_fwd", &attn_fwd);
  m.def("attn_bwd", &attn_bwd);
  m.def("xrank_barrier", &xrank_barrier);
  m.def("rs_sumsq", &rs_sumsq);
  m.def("adamw_allgather", &adamw_allgather);
  m.def("ag_pull", &ag_pull);
}
