// TORCH_LIBRARY registration of the forward-path ops: `torch.ops.sdpnet_b200.{gemm, layernorm_rows, ln_dwconv, attention}`
// as thin C++ wrappers over the C-ABI (include/sdpnet_b200.h) -- tensors in, raw device pointers + sizes + the
// tensors' current CUDA stream out, TORCH_CHECK on a non-zero status.  No compute happens here.  The reference has no
// operator layer of its own (its ops are nn.Linear / nn.Conv2d / nn.LayerNorm / F.scaled_dot_product_attention calls
// inside layers.py); these are the custom ops its forward bodies call instead.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <cstring>

#include "../../include/sdpnet_b200.h"

namespace {

int dt(const at::Tensor &t) {
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kBFloat16, "sdpnet_b200: float32 / bfloat16 tensors only");
  return t.scalar_type() == at::kBFloat16 ? SDP_BF16 : SDP_F32;
}
const float *f32(const c10::optional<at::Tensor> &t, const char *name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous(), name, " must be a contiguous float32 CUDA tensor");
  return t->data_ptr<float>();
}
void *stream_of(const at::Tensor &t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
void check(int rc, const char *what) { TORCH_CHECK(rc == 0, what, " failed (rc=", rc, "): ", sdp_last_error()); }

// layers.py:34-42, 79-92, 282-284, 301, 308, 443-460: out = act(A @ W^T + bias) + residual
void gemm(const at::Tensor &A, const at::Tensor &W, at::Tensor out, const c10::optional<at::Tensor> &bias,
          const c10::optional<at::Tensor> &residual, int64_t act) {
  TORCH_CHECK(A.is_cuda() && W.is_cuda() && out.is_cuda(), "sdpnet_b200::gemm: CUDA tensors only (there is no CPU fallback)");
  TORCH_CHECK(A.dim() == 2 && W.dim() == 2 && out.dim() == 2 && A.stride(1) == 1 && W.stride(1) == 1 && out.stride(1) == 1,
              "sdpnet_b200::gemm: 2-D, K/N-contiguous operands");
  TORCH_CHECK(A.scalar_type() == W.scalar_type() && A.size(1) == W.size(1) && out.size(0) == A.size(0) && out.size(1) == W.size(0),
              "sdpnet_b200::gemm: shape / dtype mismatch");
  c10::cuda::CUDAGuard guard(A.device());
  sdp_gemm_args a;
  std::memset(&a, 0, sizeof(a));
  a.A = A.data_ptr(); a.lda = A.stride(0);
  a.W = W.data_ptr(); a.ldw = W.stride(0);
  a.bias = f32(bias, "bias");
  a.out = out.data_ptr(); a.ldo = out.stride(0);
  a.M = (int32_t)A.size(0); a.N = (int32_t)W.size(0); a.K = (int32_t)A.size(1);
  a.dtype = dt(A); a.out_dtype = dt(out); a.act = (int32_t)act;
  if (residual.has_value() && residual->defined()) {
    TORCH_CHECK(residual->is_cuda() && residual->dim() == 2 && residual->stride(1) == 1, "sdpnet_b200::gemm: bad residual");
    a.residual = residual->data_ptr(); a.ldr = residual->stride(0); a.res_dtype = dt(*residual);
  }
  check(sdp_gemm(&a, stream_of(A)), "sdp_gemm");
}

// layers.py:280,307 (nn.LayerNorm) / :12-24 seen token-major
void layernorm_rows(const at::Tensor &x, const c10::optional<at::Tensor> &w, const c10::optional<at::Tensor> &b, at::Tensor out,
                    double eps) {
  TORCH_CHECK(x.is_cuda() && out.is_cuda() && x.dim() == 2 && out.dim() == 2 && x.stride(1) == 1 && out.stride(1) == 1 &&
                  x.sizes() == out.sizes() && x.scalar_type() == out.scalar_type(), "sdpnet_b200::layernorm_rows: bad arguments");
  c10::cuda::CUDAGuard guard(x.device());
  check(sdp_layernorm_rows(x.data_ptr(), x.stride(0), f32(w, "w"), f32(b, "b"), out.data_ptr(), out.stride(0), (int)x.size(0),
                           (int)x.size(1), (float)eps, dt(x), stream_of(x)), "sdp_layernorm_rows");
}

// layers.py:102: channel LayerNorm + depthwise k x k 'same' conv on the patch rows of [B, R + Gh*Gw, C]
void ln_dwconv(const at::Tensor &act, const at::Tensor &gamma, const at::Tensor &beta, const at::Tensor &wdw,
               const c10::optional<at::Tensor> &bdw, at::Tensor out, int64_t Gh, int64_t Gw, int64_t R, double eps) {
  TORCH_CHECK(act.is_cuda() && out.is_cuda() && act.dim() == 3 && act.is_contiguous() && out.is_contiguous() &&
                  act.sizes() == out.sizes() && act.size(1) == R + Gh * Gw, "sdpnet_b200::ln_dwconv: act / out must be contiguous [B, R + Gh*Gw, C]");
  int k = 1;
  while (k * k < wdw.size(0)) ++k;
  TORCH_CHECK(wdw.dim() == 2 && k * k == wdw.size(0) && wdw.size(1) == act.size(2), "sdpnet_b200::ln_dwconv: wdw must be tap-major [k*k, C]");
  c10::cuda::CUDAGuard guard(act.device());
  check(sdp_ln_dwconv(act.data_ptr(), f32(gamma, "gamma"), f32(beta, "beta"), f32(wdw, "wdw"), f32(bdw, "bdw"), out.data_ptr(),
                      (int)act.size(0), (int)Gh, (int)Gw, (int)act.size(2), k, (int)R, (float)eps, dt(act), stream_of(act)), "sdp_ln_dwconv");
}

// layers.py:286-300: per-head q/k LayerNorm (optional) + softmax(q k^T / sqrt(d)) v on qkv [B, S, 3C]
void attention(const at::Tensor &qkv, at::Tensor out, int64_t n_head, const c10::optional<at::Tensor> &qn_w,
               const c10::optional<at::Tensor> &qn_b, const c10::optional<at::Tensor> &kn_w, const c10::optional<at::Tensor> &kn_b,
               double eps, double score_bound) {
  TORCH_CHECK(qkv.is_cuda() && out.is_cuda() && qkv.dim() == 3 && qkv.is_contiguous() && out.is_contiguous() &&
                  qkv.size(2) == 3 * out.size(2) && qkv.size(0) == out.size(0) && qkv.size(1) == out.size(1) && out.size(2) % n_head == 0,
              "sdpnet_b200::attention: qkv [B,S,3C] and out [B,S,C] must be contiguous");
  c10::cuda::CUDAGuard guard(qkv.device());
  if (score_bound > 0.0 && !qn_w.has_value()) {   // q, k already normalised and their scores bounded: one-pass softmax
    check(sdp_attention_bounded(qkv.data_ptr(), out.data_ptr(), (int)qkv.size(0), (int)qkv.size(1), (int)n_head,
                                (int)(out.size(2) / n_head), (float)score_bound, dt(qkv), stream_of(qkv)),
          "sdp_attention_bounded");
    return;
  }
  check(sdp_attention(qkv.data_ptr(), f32(qn_w, "qn_w"), f32(qn_b, "qn_b"), f32(kn_w, "kn_w"), f32(kn_b, "kn_b"), out.data_ptr(),
                      (int)qkv.size(0), (int)qkv.size(1), (int)n_head, (int)(out.size(2) / n_head), (float)eps, dt(qkv), stream_of(qkv)),
        "sdp_attention");
}

}  // namespace

TORCH_LIBRARY(sdpnet_b200, m) {
  m.def("gemm(Tensor A, Tensor W, Tensor(a!) out, Tensor? bias, Tensor? residual, int act) -> ()");
  m.def("layernorm_rows(Tensor x, Tensor? w, Tensor? b, Tensor(a!) out, float eps) -> ()");
  m.def("ln_dwconv(Tensor act, Tensor gamma, Tensor beta, Tensor wdw, Tensor? bdw, Tensor(a!) out, int Gh, int Gw, int R, float eps) -> ()");
  m.def("attention(Tensor qkv, Tensor(a!) out, int n_head, Tensor? qn_w, Tensor? qn_b, Tensor? kn_w, Tensor? kn_b, float eps, float score_bound=0.0) -> ()");
}

TORCH_LIBRARY_IMPL(sdpnet_b200, CUDA, m) {
  m.impl("gemm", &gemm);
  m.impl("layernorm_rows", &layernorm_rows);
  m.impl("ln_dwconv", &ln_dwconv);
  m.impl("attention", &attention);
}
