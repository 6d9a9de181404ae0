// fp32 verification-mode GEMM on the CUDA cores (FFMA, no TF32) + the sdp_gemm C-ABI dispatcher.
// The fp32 mode exists for the 1e-4 logits check of the north star; the product path is bf16 on
// tcgen05 (gemm_tc.cu).  Same epilogue code (common.cuh) as the tensor-core kernel.
#include "common.cuh"

namespace sdp {

int gemm_bf16_tc(const sdp_gemm_args &a, const Epilogue &e, cudaStream_t st);
int gemm_stats_parts(int N);

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

// 256 threads, each a 4 x 4 micro-tile (rows strided by 16 so that a thread's 4 columns are
// contiguous for the shared epilogue).  A: [M,K], W: [N,K], both K-contiguous.
__global__ void __launch_bounds__(256)
gemm_f32_simt_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ W, long long ldw,
                     const Epilogue epi, int K) {
  __shared__ float sA[SB_K][SB_M + 4];
  __shared__ float sW[SB_K][SB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * SB_M, n0 = blockIdx.y * SB_N;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += SB_K) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; k fastest for coalescing
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      sA[kk][r] = (gm < epi.M && gk < K) ? A[(long long)gm * lda + gk] : 0.0f;
      sW[kk][r] = (gn < epi.N && gk < K) ? W[(long long)gn * ldw + gk] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SB_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
      const float4 bv = *reinterpret_cast<const float4 *>(&sW[kk][tx * 4]);
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const RowMap rm = map_row(epi, m0 + ty + 16 * i);
    const int c0 = n0 + tx * 4;
    if (c0 >= epi.N) continue;
    // the shared epilogue works on groups of 8; finish 4 columns scalar-wise (vec=false)
    float v[8] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3], 0.f, 0.f, 0.f, 0.f};
    Epilogue e4 = epi;
    e4.N = min(epi.N, c0 + 4);
    epilogue_row<8, true, -1>(e4, rm, c0, v, false);
  }
}

}  // namespace sdp

using namespace sdp;

extern "C" int sdp_gemm_headnorm_ok(int d, int N, int dtype) {
  return dtype == SDP_BF16 && (d == 32 || d == 64 || d == 96 || d == 128) && N % d == 0 ? 1 : 0;
}

extern "C" int sdp_gemm_stats_parts(int N, int dtype) { return dtype == SDP_BF16 && N > 0 ? gemm_stats_parts(N) : 0; }

extern "C" int sdp_gemm(const sdp_gemm_args *a, void *stream) {
  SDP_CHECK(a != nullptr, "sdp_gemm: null args");
  SDP_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "sdp_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  SDP_CHECK(a->A && a->W && a->out, "sdp_gemm: null operand");
  SDP_CHECK(a->pass_seq == 0 || a->out == a->residual,
            "sdp_gemm: pass-through rows need an in-place residual (out == residual)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Epilogue e;
  e.bias = a->bias;
  e.residual = a->residual;
  e.out = a->out;
  e.res_lo = a->residual_lo;
  e.out_lo = a->out_lo;
  e.ldr = a->ldr;
  e.ldo = a->ldo;
  e.M = a->M;
  e.N = a->N;
  e.out_dtype = a->out_dtype;
  e.res_dtype = a->res_dtype;
  e.act = a->act;
  e.res_first = a->res_first;
  e.res_mod = a->res_mod;
  e.seq_in = a->seq_in;
  e.seq_out = a->seq_out;
  e.seq_off = a->seq_off;
  e.pass_seq = a->pass_seq;
  e.pass_rows = a->pass_rows;
  e.hn_d = a->headnorm_d;
  e.hn_C = a->headnorm_C;
  e.hn_eps = a->headnorm_eps;
  e.hn_qw = a->hn_q_w; e.hn_qb = a->hn_q_b; e.hn_kw = a->hn_k_w; e.hn_kb = a->hn_k_b;
  e.stats_out = a->stats_out; e.stats_parts = a->stats_parts;
  e.ln_stats = a->ln_stats; e.ln_parts = a->ln_parts; e.ln_K = a->K; e.ln_eps = a->ln_eps;
  e.ln_s = a->ln_s; e.ln_t = a->ln_t;
  if (a->residual_lo || a->out_lo) {
    SDP_CHECK(a->dtype == SDP_BF16 && a->out_dtype == SDP_BF16 && (a->residual_lo == nullptr || (a->residual && a->res_dtype == SDP_BF16)),
              "sdp_gemm: a split (hi + lo) stream needs bf16 operands, bf16 output and a bf16 residual");
  }
  if (a->stats_out) {
    SDP_CHECK(a->stats_parts > 0 && a->stats_parts == sdp_gemm_stats_parts(a->N, a->dtype) && a->out_dtype == SDP_BF16 &&
                  a->seq_in == 0,
              "sdp_gemm: stats_out needs a bf16 tensor-core GEMM with stats_parts == sdp_gemm_stats_parts(N)");
  }
  if (a->ln_stats) {
    SDP_CHECK(a->dtype == SDP_BF16 && a->ln_parts > 0 && a->ln_parts % 2 == 0 && a->ln_s && a->ln_t && a->bias == nullptr &&
                  (reinterpret_cast<uintptr_t>(a->ln_stats) & 15) == 0,
              "sdp_gemm: LN folding needs bf16, ln_s/ln_t, an even ln_parts and no bias");
  }
  if (a->headnorm_d) {
    SDP_CHECK(sdp_gemm_headnorm_ok(a->headnorm_d, a->N, a->dtype), "sdp_gemm: head-norm unsupported for d=%d N=%d dtype=%d",
              a->headnorm_d, a->N, a->dtype);
    SDP_CHECK(a->bias == nullptr && a->hn_q_w && a->hn_q_b && a->hn_k_w && a->hn_k_b && a->headnorm_C > 0 &&
                  a->headnorm_C % a->headnorm_d == 0,
              "sdp_gemm: head-norm needs q/k affine parameters, no bias, and C %% d == 0");
  }
  if (a->dtype == SDP_BF16) {
    SDP_CHECK(sdp_device_ok(), "sdp_gemm: bf16 path needs an sm_100 device (tcgen05/TMEM); none found");
    return gemm_bf16_tc(*a, e, st);
  }
  SDP_CHECK(a->dtype == SDP_F32, "sdp_gemm: unknown dtype %d", a->dtype);
  dim3 grid((a->M + SB_M - 1) / SB_M, (a->N + SB_N - 1) / SB_N);
  SDP_CHECK(grid.y <= 65535, "sdp_gemm(fp32): N=%d too large for the verification kernel", a->N);
  gemm_f32_simt_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float *>(a->A), a->lda,
                                              reinterpret_cast<const float *>(a->W), a->ldw, e, a->K);
  SDP_LAUNCH_OK();
  return 0;
}
