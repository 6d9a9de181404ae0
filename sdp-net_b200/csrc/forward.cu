// ABI housekeeping + the whole-model forward (reference model.py:129-149) sequenced behind the
// C-ABI: one call enqueues every kernel of the forward on the caller's stream.  No allocation,
// no synchronisation, no host<->device copies happen here.
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace sdp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace sdp

using namespace sdp;

extern "C" int sdp_abi_version(void) { return SDPNET_B200_ABI_VERSION; }
extern "C" const char *sdp_last_error(void) { return g_err; }
extern "C" int64_t sdp_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}
extern "C" int sdp_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

namespace {

struct Ctx {
  const sdp_model_desc *m;
  const sdp_workspace *ws;
  int B, S, T, R, Gh, Gw;
  int parts;       // column parts of the row-statistics buffer (ln_fold, or producer statistics for the mixers)
  bool emit;       // mixer-feeding GEMMs write (sum, sumsq) parts for the tensor-core depthwise kernel
  mutable bool stats_fresh;   // ws->stats currently describes ws->act
  void *st;
};

struct Fold {                 // LayerNorm folded into the consumer GEMM (bf16 only)
  const float *s = nullptr, *t = nullptr;
  float eps = 0.0f;
};

int gemm(const Ctx &c, const void *A, long long lda, const void *W, long long ldw, const float *bias, int M, int N,
         int K, int act, const void *res, void *out, long long ldo, int out_dtype, bool mask_regs,
         const Fold *fold = nullptr, bool emit_stats = false, const sdp_encoder_weights *qk = nullptr) {
  sdp_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.bias = bias;
  a.residual = res; a.ldr = ldo; a.out = out; a.ldo = ldo;
  a.M = M; a.N = N; a.K = K;
  a.dtype = c.m->dtype; a.out_dtype = out_dtype; a.res_dtype = c.m->dtype;
  a.act = act;
  if (out == c.ws->act && c.ws->act_lo != nullptr) {   // the split residual stream: both planes, in place
    a.out_lo = c.ws->act_lo;
    if (res == c.ws->act) a.residual_lo = c.ws->act_lo;
  }
  if (mask_regs && c.R > 0) { a.pass_seq = c.S; a.pass_rows = c.R; }
  if (fold) {
    a.ln_stats = c.ws->stats; a.ln_parts = c.parts; a.ln_eps = fold->eps; a.ln_s = fold->s; a.ln_t = fold->t;
  }
  if (emit_stats) { a.stats_out = c.ws->stats; a.stats_parts = c.parts; }
  if (qk) {     // per-head q/k LayerNorm in the QKV epilogue
    a.headnorm_d = c.m->C / c.m->n_head; a.headnorm_C = c.m->C; a.headnorm_eps = 1e-5f;
    a.hn_q_w = qk->qn_w; a.hn_q_b = qk->qn_b; a.hn_k_w = qk->kn_w; a.hn_k_b = qk->kn_b;
  }
  return sdp_gemm(&a, c.st);
}

// layers.py:259-316
int encoder(const Ctx &c, const sdp_encoder_weights &w) {
  const sdp_model_desc &m = *c.m;
  const int C = m.C, M = c.B * c.S, dt = m.dtype, F = m.ff_mult * C;
  const bool fold = m.ln_fold != 0;
  // q/k LayerNorm (layers.py:286): fused into the QKV GEMM epilogue when the tile holds whole heads,
  // otherwise applied by the attention kernel while it loads q and k
  const int d = C / m.n_head;
  const bool fuse_qk = w.qn_w != nullptr && sdp_gemm_headnorm_ok(d, 3 * C, dt);
  const void *xin = c.ws->act;                  // with folding the GEMMs read the residual stream itself
  if (!fold) {
    if (int rc = sdp_layernorm_rows(c.ws->act, C, w.norm1_w, w.norm1_b, c.ws->norm, C, M, C, 1e-5f, dt, c.st)) return rc;
    xin = c.ws->norm;
  }
  Fold f1; f1.s = w.s_qkv; f1.t = w.t_qkv; f1.eps = 1e-5f;
  if (int rc = gemm(c, xin, C, w.w_qkv, C, nullptr, M, 3 * C, C, SDP_ACT_NONE, nullptr, c.ws->qkv, 3 * C, dt, false,
                    fold ? &f1 : nullptr, false, fuse_qk ? &w : nullptr)) return rc;
  if (fuse_qk && w.qk_score_bound > 0.0f) {     // q, k normalised by the epilogue above, their scores bounded by its parameters
    if (int rc = sdp_attention_bounded(c.ws->qkv, c.ws->attn, c.B, c.S, m.n_head, d, w.qk_score_bound, dt, c.st)) return rc;
  } else if (int rc = sdp_attention(c.ws->qkv, fuse_qk ? nullptr : w.qn_w, fuse_qk ? nullptr : w.qn_b,
                                    fuse_qk ? nullptr : w.kn_w, fuse_qk ? nullptr : w.kn_b, c.ws->attn, c.B, c.S, m.n_head, d,
                                    1e-5f, dt, c.st)) {
    return rc;
  }
  if (int rc = gemm(c, c.ws->attn, C, w.w_o, C, nullptr, M, C, C, SDP_ACT_NONE, c.ws->act, c.ws->act, C, dt, false,
                    nullptr, fold)) return rc;
  if (!fold) {
    if (int rc = sdp_layernorm_rows(c.ws->act, C, w.norm2_w, w.norm2_b, c.ws->norm, C, M, C, 1e-5f, dt, c.st)) return rc;
  }
  Fold f2; f2.s = w.s_ff1; f2.t = w.t_ff1; f2.eps = 1e-5f;
  if (int rc = gemm(c, xin, C, w.w_ff1, C, fold ? nullptr : w.b_ff1, M, F, C, m.act, nullptr, c.ws->hidden, F, dt, false,
                    fold ? &f2 : nullptr)) return rc;
  c.stats_fresh = fold || c.emit;
  return gemm(c, c.ws->hidden, F, w.w_ff2, F, w.b_ff2, M, C, F, SDP_ACT_NONE, c.ws->act, c.ws->act, C, dt, false, nullptr,
              fold || c.emit);
}

// layers.py:101-104
int mixer(const Ctx &c, const sdp_mixer_weights &w) {
  const sdp_model_desc &m = *c.m;
  const int C = m.C, M = c.B * c.S, dt = m.dtype;
  const bool fold = m.ln_fold != 0;
  const bool have = (fold || c.emit) && c.stats_fresh;      // ws->stats holds the producer GEMM's (sum, sumsq) parts of act
  if (c.ws->stats != nullptr && sdp_ln_dwconv_slab_ok(c.Gh, c.Gw, C, m.conv_k, dt)) {
    // channel-stationary tensor-core kernel.  Token statistics: the producer GEMM's parts when it wrote them, else a
    // pass over act.  Its (mean, rstd) scratch is the head of the QKV buffer (dead between two encoders) when the
    // statistics workspace is in use, else the statistics workspace itself.
    float *scratch = have ? reinterpret_cast<float *>(c.ws->qkv) : c.ws->stats;
    if (int rc = sdp_ln_dwconv_slab_stats(c.ws->act, have ? c.ws->stats : nullptr, have ? c.parts : 0, scratch, w.ln1_g,
                                          w.ln1_b, w.w_dw, w.b_dw, c.ws->norm, c.B, c.Gh, c.Gw, C, m.conv_k, c.R, 1e-6f,
                                          c.st)) return rc;
  } else {
    const float *dw_stats = have ? c.ws->stats : nullptr;
    int dw_parts = c.parts;
    if (!have && c.ws->stats != nullptr && sdp_ln_dwconv_wants_stats(c.Gh, c.Gw, C, m.conv_k, c.R, dt)) {
      if (int rc = sdp_row_stats(c.ws->act, C, c.ws->stats, 1, M, C, dt, c.st)) return rc;
      dw_stats = c.ws->stats;
      dw_parts = 1;
    }
    if (int rc = sdp_ln_dwconv_stats(c.ws->act, dw_stats, dw_parts, w.ln1_g, w.ln1_b, w.w_dw, w.b_dw, c.ws->norm, c.B,
                                     c.Gh, c.Gw, C, m.conv_k, c.R, 1e-6f, dt, c.st)) return rc;
  }
  if (int rc = gemm(c, c.ws->norm, C, w.w_pw, C, w.b_pw, M, C, C, m.act, c.ws->act, c.ws->act, C, dt, true, nullptr, fold)) return rc;
  c.stats_fresh = fold;
  const void *xin = c.ws->act;
  if (!fold) {
    if (int rc = sdp_layernorm_rows(c.ws->act, C, w.ln2_g, w.ln2_b, c.ws->norm, C, M, C, 1e-6f, dt, c.st)) return rc;
    xin = c.ws->norm;
  }
  Fold f; f.s = w.s_mlp1; f.t = w.t_mlp1; f.eps = 1e-6f;
  if (int rc = gemm(c, xin, C, w.w_mlp1, C, fold ? nullptr : w.b_mlp1, M, 4 * C, C, m.act, nullptr, c.ws->hidden, 4 * C, dt,
                    false, fold ? &f : nullptr)) return rc;
  c.stats_fresh = fold || c.emit;
  return gemm(c, c.ws->hidden, 4 * C, w.w_mlp2, 4 * C, w.b_mlp2, M, C, 4 * C, SDP_ACT_NONE, c.ws->act, c.ws->act, C, dt, true,
              nullptr, fold || c.emit);
}

}  // namespace

extern "C" int sdp_forward(const sdp_model_desc *m, const sdp_workspace *ws, const void *x, int x_dtype, int B,
                           int H, int W, int R, float *logits, void *stream) {
  SDP_CHECK(m && ws && x && logits, "sdp_forward: null argument");
  SDP_CHECK(B > 0 && R >= 1, "sdp_forward: need B > 0 and R >= 1 (got B=%d R=%d)", B, R);
  SDP_CHECK(m->patch > 0 && H % m->patch == 0 && W % m->patch == 0,
            "sdp_forward: image %dx%d not divisible by patch %d", H, W, m->patch);
  SDP_CHECK(m->C % m->n_head == 0, "sdp_forward: embedding_dim %d not divisible by n_head %d", m->C, m->n_head);
  Ctx c;
  c.m = m; c.ws = ws; c.B = B; c.R = R; c.st = stream;
  c.Gh = H / m->patch; c.Gw = W / m->patch;
  c.T = c.Gh * c.Gw; c.S = c.T + R;
  const int C = m->C, dt = m->dtype, Kc3 = 3 * m->patch * m->patch;
  const int sparts = ws->stats != nullptr ? sdp_gemm_stats_parts(C, dt) : 0;
  c.emit = !m->ln_fold && sparts > 0 && sparts % 2 == 0 && sparts <= 16 && m->conv_block_num > 0 &&
           sdp_ln_dwconv_slab_ok(c.Gh, c.Gw, C, m->conv_k, dt);
  c.stats_fresh = false;
  c.parts = (m->ln_fold || c.emit) ? sparts : 0;
  SDP_CHECK(!m->ln_fold || (dt == SDP_BF16 && c.parts > 0 && c.parts % 2 == 0 && ws->stats != nullptr),
            "sdp_forward: ln_fold needs bf16 and a statistics workspace");
  SDP_CHECK(ws->act_lo == nullptr || dt == SDP_BF16, "sdp_forward: the split (hi + lo) residual stream is a bf16 feature");

  // patcher + position table + embedding activation, scattered behind the register rows
  // (layers.py:34-42, :152-168 / :202-209)
  if (int rc = sdp_im2col_patches(x, x_dtype, ws->im2col, dt, m->Kp, B, H, W, m->patch, stream)) return rc;
  {
    sdp_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = ws->im2col; a.lda = m->Kp; a.W = m->w_patch; a.ldw = m->Kp;
    a.residual = m->pos_table; a.ldr = C; a.res_dtype = SDP_F32; a.res_first = 1; a.res_mod = c.T;
    a.out = ws->act; a.out_lo = ws->act_lo; a.ldo = C; a.M = B * c.T; a.N = C; a.K = Kc3;
    a.dtype = dt; a.out_dtype = dt; a.act = m->embed_act;
    a.seq_in = c.T; a.seq_out = c.S; a.seq_off = R;
    if (int rc = sdp_gemm(&a, stream)) return rc;
  }
  if (int rc = sdp_fill_registers(ws->act, ws->act_lo, dt, m->reg_table, B, c.S, R, C, stream)) return rc;
  if (m->ln_fold) {   // first statistics of the residual stream; every later producer GEMM refreshes its rows
    if (int rc = sdp_row_stats(ws->act, C, ws->stats, c.parts, B * c.S, C, dt, stream)) return rc;
    c.stats_fresh = true;
  }

  for (int i = 0; i < m->num_blocks; ++i) {            // model.py:139-140, layers.py:377-386
    if (m->conv_first)
      for (int j = 0; j < m->conv_block_num; ++j)
        if (int rc = mixer(c, m->mix[i * m->conv_block_num + j])) return rc;
    if (int rc = encoder(c, m->enc[i])) return rc;
    if (!m->conv_first)
      for (int j = 0; j < m->conv_block_num; ++j)
        if (int rc = mixer(c, m->mix[i * m->conv_block_num + j])) return rc;
  }
  if (int rc = encoder(c, m->enc[m->num_blocks])) return rc;   // model.py:143

  // classification head (layers.py:443-465)
  const int K = m->classes;
  if (m->head_from_register) {
    if (int rc = sdp_pool_ln(ws->act, ws->act_lo, dt, B, c.S, C, 0, R, m->head_ln_w, m->head_ln_b, 1e-5f, ws->pooled, dt, C, stream)) return rc;
  } else {
    if (int rc = sdp_pool_ln(ws->act, ws->act_lo, dt, B, c.S, C, R, c.T, nullptr, nullptr, 0.0f, ws->pooled, dt, C, stream)) return rc;
  }
  const bool two = m->head_from_register && !m->head_simple;
  {
    sdp_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = ws->pooled; a.lda = C; a.W = m->w_head1; a.ldw = C; a.bias = m->b_head1;
    a.M = B; a.N = K; a.K = C; a.dtype = dt;
    if (two) { a.out = ws->head_h; a.ldo = m->Kc; a.out_dtype = dt; a.act = SDP_ACT_TANH; }
    else { a.out = logits; a.ldo = K; a.out_dtype = SDP_F32; a.act = SDP_ACT_NONE; }
    if (int rc = sdp_gemm(&a, stream)) return rc;
  }
  if (two) {
    sdp_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = ws->head_h; a.lda = m->Kc; a.W = m->w_head2; a.ldw = m->Kc; a.bias = m->b_head2;
    a.M = B; a.N = K; a.K = K; a.dtype = dt;
    a.out = logits; a.ldo = K; a.out_dtype = SDP_F32; a.act = SDP_ACT_NONE;
    if (int rc = sdp_gemm(&a, stream)) return rc;
  }
  return 0;
}
