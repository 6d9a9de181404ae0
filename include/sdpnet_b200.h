/*
 * sdpnet_b200.h -- C-ABI of the B200-native (sm_100a) SdP-Net forward engine.
 *
 * The reference (y-akbal/SdP-Net) has no FFI layer: its hot path is `MainModel.forward`
 * (reference model.py:129-149) calling PyTorch modules (reference layers.py).  The drop-in
 * boundary is therefore the nn.Module API, mirrored by the Python package `sdp-net_b200/`,
 * whose `forward` bodies call ONLY the entry points below (via ctypes, registered as
 * `torch.ops.sdpnet_b200.*`).  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless stated otherwise;
 *     no ownership transfer: the caller allocates inputs, outputs and workspaces.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - return value: 0 = ok, non-zero = error; `sdp_last_error()` gives the message
 *     (thread-local).  Nothing here ever computes on the CPU.
 *   - dtypes: SDP_F32 / SDP_BF16.  Activations are token-major `[B, S = R + T, C]`:
 *     rows 0..R-1 of every image are the register (CLS) tokens, rows R..S-1 the patches in
 *     row-major (i*Gw + j) order -- the reference's `cat([register, x_flat])`
 *     (layers.py:271-275) made permanent, so no transpose/cat/split ever runs.
 *   - parameters (LN affine, biases, depthwise taps, tables) are fp32; GEMM weights are in
 *     the compute dtype, `[N, K]` row-major (nn.Linear layout; 1x1 convs reshaped).
 */
#ifndef SDPNET_B200_H
#define SDPNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDPNET_B200_ABI_VERSION 7

typedef enum { SDP_F32 = 0, SDP_BF16 = 1 } sdp_dtype;

/* reference model.py:13-24 (`activations` table) + training_utilities.py:91-92 (KeLu) */
typedef enum {
  SDP_ACT_NONE = 0,
  SDP_ACT_RELU = 1,
  SDP_ACT_GELU = 2,       /* exact erf form, nn.GELU() */
  SDP_ACT_GELU_TANH = 3,
  SDP_ACT_TANH = 4,
  SDP_ACT_SIGMOID = 5,
  SDP_ACT_LEAKY_RELU = 6, /* slope 0.01 */
  SDP_ACT_SELU = 7,
  SDP_ACT_KELU = 8        /* a = 3.5 */
} sdp_act;

int sdp_abi_version(void);
const char *sdp_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM/TMA path usable). */
int sdp_device_ok(void);

/* ---------------------------------------------------------------------------------------
 * GEMM with fused epilogue:  out = epi(A[M,K] . W[N,K]^T)
 * Replaces every nn.Linear / 1x1 nn.Conv2d / the patch conv on the path:
 *   layers.py:34-42 (patcher), :79-92 (mixer pointwise + MLP), :282-284,301 (q/k/v/o),
 *   :308 (FFN), :443-460 (head).
 * bf16: TMA-fed tcgen05.mma, fp32 accumulators in TMEM, persistent warp-specialised kernel.
 * fp32: CUDA-core FFMA kernel (verification mode, 1e-4 logits).
 *
 * Epilogue, per element (r, c), v = acc:
 *   v += bias[c]                                   (bias != NULL)
 *   if res_first:  v = act(v + res[rr, c])         (patch embedding: act(x + pos))
 *   else:          v = act(v) + res[ro, c]         (residual != NULL)
 *   out[ro, c] = v
 * where  ro = seq_in ? (r / seq_in) * seq_out + seq_off + r % seq_in : r   (row remap, used to
 * scatter the B*T patch rows into the [B, S, C] activation behind the R register rows),
 * rr = res_mod ? r % res_mod : ro   (res_mod = T: broadcast the [T, C] position table), and
 * rows with (ro % pass_seq) < pass_rows are left untouched when pass_seq != 0 (the mixers
 * must not modify register rows; requires out == residual, i.e. in-place).
 *
 * Head-norm (bf16 only, the QKV projection): when headnorm_d != 0, before anything else every
 * group of headnorm_d consecutive columns below 2*headnorm_C gets a LayerNorm over the group
 * (eps headnorm_eps, affine hn_q_* for columns < headnorm_C, hn_k_* above) -- the per-head
 * nn.LayerNorm on q and k (layers.py:236-237,286) fused into the projection's epilogue.
 * Supported for headnorm_d in {32, 64, 96, 128} (see sdp_gemm_headnorm_ok); needs bias == NULL.
 *
 * LayerNorm folding (bf16 only; removes the stand-alone token-LayerNorm pass of layers.py:280,307 and
 * :103).  A producer GEMM whose output is the residual stream (N == C) also emits, per output row and
 * per column part, the partial (sum, sum of squares) of the bf16 values it stores:
 *     stats_out[(ro * stats_parts + part) * 2 + {0,1}],  stats_parts == sdp_gemm_stats_parts(N)
 * A consumer GEMM reading that stream as A applies LN(x) @ W^T without ever materialising LN(x):
 * with W' = W * diag(gamma) (given as `W`), ln_s[n] = sum_k W'[n,k], ln_t[n] = sum_k beta[k] W[n,k] + b[n]:
 *     v = rstd_r * (acc - mean_r * ln_s[c]) + ln_t[c]
 * where (mean_r, rstd_r) come from ln_stats (ln_parts partials per row, eps ln_eps, dim K).  Applied
 * before the head-norm / activation / residual steps; needs bias == NULL.
 *
 * Split residual stream (bf16 only).  The reference keeps its residual stream in fp32 even under autocast
 * (model.py:129-149 adds bf16 GEMM outputs to an fp32 `x`); a bf16 stream rounds it at each of the 6 residual
 * adds per block, which is 3-4 x the reference's own bf16 noise at XL depth.  With residual_lo / out_lo the
 * stream is two bf16 planes of the same shape and pitch:  value = hi + lo,  hi = bf16(v),  lo = bf16(v - hi)
 * (~16 mantissa bits).  `residual` / `out` are the hi planes -- which is also what every consumer GEMM, LayerNorm
 * and the depthwise kernel read as their input, exactly the rounding the reference's autocast applies there --
 * and only the residual epilogues, the head pooling and the raw-output bridge touch the lo plane.
 * --------------------------------------------------------------------------------------- */
typedef struct {
  const void *A;        int64_t lda;   /* [M, K], row pitch in elements */
  const void *W;        int64_t ldw;   /* [N, K] */
  const float *bias;                   /* [N] or NULL */
  const void *residual; int64_t ldr;   /* or NULL */
  void *out;            int64_t ldo;
  int32_t M, N, K;
  int32_t dtype;        /* of A and W */
  int32_t out_dtype;
  int32_t res_dtype;
  int32_t act;
  int32_t res_first;
  int32_t res_mod;
  int32_t seq_in, seq_out, seq_off;
  int32_t pass_seq, pass_rows;
  int32_t headnorm_d, headnorm_C;
  float headnorm_eps;
  const float *hn_q_w, *hn_q_b, *hn_k_w, *hn_k_b;
  float *stats_out;     int32_t stats_parts;      /* producer side, or NULL */
  const float *ln_stats; int32_t ln_parts;        /* consumer side, or NULL */
  float ln_eps;
  const float *ln_s, *ln_t;                       /* [N] each */
  const void *residual_lo;                        /* lo plane of the residual (pitch ldr), or NULL */
  void *out_lo;                                   /* lo plane of the output (pitch ldo), or NULL */
} sdp_gemm_args;

int sdp_gemm(const sdp_gemm_args *args, void *stream);
/* 1 if sdp_gemm can fuse the per-head LayerNorm for this head_dim / N / dtype. */
int sdp_gemm_headnorm_ok(int head_dim, int N, int dtype);
/* Number of per-row column parts a producer GEMM of width N emits into stats_out (0: unsupported). */
int sdp_gemm_stats_parts(int N, int dtype);
/* Stand-alone producer of the same statistics layout (after the patch embedding / register fill):
 * part 0 holds the row's full (sum, sum of squares), the other parts are zero.  x: [M, C]. */
int sdp_row_stats(const void *x, int64_t ldx, float *stats, int parts, int M, int C, int dtype, void *stream);

/* im2col for kernel == stride patches (layers.py:34-42): x NCHW [B,3,H,W] (fp32 or bf16) ->
 * A [B*T, ldA] with A[b*T + i*Gw + j, c*p*p + dy*p + dx] = x[b, c, i*p+dy, j*p+dx]; columns
 * 3*p*p .. ldA-1 are zero-filled. */
int sdp_im2col_patches(const void *x, int x_dtype, void *A, int a_dtype, int64_t ldA,
                       int B, int H, int W, int p, void *stream);

/* Register rows (layers.py:157,166 / :206-208): act[b, r, :] = table[r, :] for r < R.  act_lo: the lo plane of a
 * split bf16 stream (see sdp_gemm), or NULL. */
int sdp_fill_registers(void *act, void *act_lo, int dtype, const float *table, int B, int S, int R, int C,
                       void *stream);

/* Token LayerNorm over the last dim (nn.LayerNorm eps 1e-5: layers.py:280,307; channel-first
 * LayerNorm eps 1e-6 seen token-major: layers.py:12-24,103). x,out: [M, C]. */
int sdp_layernorm_rows(const void *x, int64_t ldx, const float *w, const float *b, void *out,
                       int64_t ldo, int M, int C, float eps, int dtype, void *stream);

/* Mixer front half (layers.py:102 up to the depthwise conv): per image, channel-LayerNorm
 * (eps 1e-6, gamma/beta) of the patch rows, then depthwise kxk 'same' conv (zero halo applied
 * AFTER the norm), taps wdw TAP-MAJOR [k*k, C] fp32 (wdw[(dy*k+dx)*C + c] = conv.weight[c,0,dy,dx]),
 * optional bias bdw [C].  act,out: [B, S, C];
 * register rows of `out` are written as zeros. */
int sdp_ln_dwconv(const void *act, const float *gamma, const float *beta, const float *wdw,
                  const float *bdw, void *out, int B, int Gh, int Gw, int C, int k, int R,
                  float eps, int dtype, void *stream);
/* Same, with the token statistics supplied by the producer GEMM (stats layout as above; row index =
 * b * S + R + t) instead of being recomputed from `act`.  stats == NULL behaves like sdp_ln_dwconv. */
/* 1 if the kernel chosen for this shape REQUIRES caller-supplied row statistics (currently never: every
 * kernel computes them itself when stats == NULL). */
int sdp_ln_dwconv_wants_stats(int Gh, int Gw, int C, int k, int R, int dtype);
int sdp_ln_dwconv_stats(const void *act, const float *stats, int parts, const float *gamma, const float *beta,
                        const float *wdw, const float *bdw, void *out, int B, int Gh, int Gw, int C, int k,
                        int R, float eps, int dtype, void *stream);

/* The same operator as a channel-stationary tensor-core kernel (dwconv_slab.cu): a CTA keeps the Toeplitz
 * fragments of 32 channels' taps in registers and streams images past them (14 mma.sync per channel and image
 * for k = 7 instead of 12544 FMAs).  Needs `token_stats`: caller-owned scratch of 2 * B * Gh * Gw floats that
 * the call fills with (mean, rstd) of every spatial token and then consumes.  bf16 only;
 * sdp_ln_dwconv_slab_ok says whether the shape is covered (Gh, Gw <= 16 with an even token count, C % 32 == 0,
 * C <= 2048, k in {3,5,7}). */
int sdp_ln_dwconv_slab_ok(int Gh, int Gw, int C, int k, int dtype);
int sdp_ln_dwconv_slab(const void *act, float *token_stats, const float *gamma, const float *beta,
                       const float *wdw, const float *bdw, void *out, int B, int Gh, int Gw, int C, int k,
                       int R, float eps, void *stream);
/* Same, with the statistics taken from the GEMM that produced `act` instead of a pass over `act`: producer_stats =
 * that GEMM's stats_out buffer ((sum, sumsq) column parts, row index b * S + R + t, parts ==
 * sdp_gemm_stats_parts(C)); a small kernel turns them into (mean, rstd) in `token_stats` (scratch as above, must
 * not alias producer_stats).  producer_stats == NULL behaves like sdp_ln_dwconv_slab. */
int sdp_ln_dwconv_slab_stats(const void *act, const float *producer_stats, int parts, float *token_stats,
                             const float *gamma, const float *beta, const float *wdw, const float *bdw,
                             void *out, int B, int Gh, int Gw, int C, int k, int R, float eps, void *stream);

/* Fused QK-LayerNorm + softmax attention (layers.py:282-300): qkv [B, S, 3C] with column
 * blocks q | k | v, each [h, d]; q/k get a per-head LayerNorm(d) (eps, affine) when
 * qn_w != NULL; out[b, s, head*d + :] = softmax(q k^T / sqrt(d)) v.  No mask, no dropout
 * (eval).  bf16 with d % 16 == 0, d <= 128 runs on tensor cores; otherwise CUDA cores. */
int sdp_attention(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                  const float *kn_b, void *out, int B, int S, int h, int d, float eps, int dtype,
                  void *stream);

/* The same attention (layers.py:289-291) for q / k whose per-head LayerNorm (layers.py:286) has already been applied
 * (the QKV GEMM's head-norm epilogue), with a caller-supplied bound: |q_i . k_j| / sqrt(d) <= score_bound (nats) for
 * every pair.  LayerNorm outputs have a known norm -- |LN(x)|_2 <= max|gamma| sqrt(d) + |beta|_2 -- so the bound follows
 * from the four parameter vectors alone.  With it the softmax needs no row maximum (it is invariant to the exponent
 * reference; exp(s) cannot overflow) and the tcgen05 kernel reads every score from TMEM once instead of twice.
 * score_bound == 0 (unknown) or above 60 nats, or a shape the tcgen05 kernel does not cover: behaves exactly like
 * sdp_attention with qn_w == NULL.  A bound that does not hold makes the result undefined (inf / NaN), never out of bounds. */
int sdp_attention_bounded(const void *qkv, void *out, int B, int S, int h, int d, float score_bound, int dtype,
                          void *stream);

/* Head front (layers.py:464 `registers.mean(-2)` + LN at :445/:449, or AdaptiveAvgPool at
 * :457): out[b, :] = LN(mean over rows [row0, row0+nrows) of image b), LN skipped when
 * ln_w == NULL.  act [B, S, C] (+ act_lo, the lo plane of a split bf16 stream, or NULL) -> out [B, ldo]. */
int sdp_pool_ln(const void *act, const void *act_lo, int dtype, int B, int S, int C, int row0, int nrows,
                const float *ln_w, const float *ln_b, float eps, void *out, int out_dtype,
                int64_t ldo, void *stream);

/* Layout bridges for the reference's NCHW module API (layers.py:271,314) and for
 * `return_raw_outputs` (model.py:147-149).  x NCHW [B,C,Gh,Gw], reg [B,R,C], both fp32. */
int sdp_tokens_from_nchw(const float *x, const float *reg, void *act, int dtype, int B, int C,
                         int T, int R, void *stream);
int sdp_tokens_to_nchw(const void *act, const void *act_lo, int dtype, float *x, float *reg, int B, int C, int T,
                       int R, void *stream);   /* act_lo: lo plane of a split bf16 stream, or NULL */

/* Position-embedding add on token-major data (layers.py:162-168 / :205):
 * act[b, R + t, :] = actfn(act[b, R + t, :] + pos[t, :]); pos is [T, C] fp32.  (MainModel fuses
 * this into the patch GEMM epilogue; this entry serves the stand-alone EmbeddingLayer module.) */
int sdp_embed_tokens(void *act, int dtype, const float *pos, int B, int T, int R, int C, int act_id,
                     void *stream);

/* Elementwise activation (embedding activation on a [n] buffer; test hook for epilogues). */
int sdp_activation(const void *x, void *y, int64_t n, int act, int dtype, void *stream);

/* Evaluation metrics right behind the forward, without per-batch host syncs (model_test.py:76-85,
 * training_utilities.py:50-88, 95-107): for logits [B, K] (fp32) and integer labels [B], ADD into acc[5]
 * (device doubles): sum over rows of cross-entropy, sum over all B*K elements of
 * binary_cross_entropy_with_logits against the smoothed one-hot target
 * (one_hot*(1-ls) + ls/K), number of rows whose argmax equals the label, number of rows, and -- acc[4] -- the
 * number of rows whose label lies outside [0, K) (nn.CrossEntropyLoss raises on those; such rows contribute to
 * nothing else; label -100, its ignore_index, is skipped silently). */
int sdp_eval_metrics(const float *logits, int64_t ldl, const int64_t *labels, int B, int K, float label_smoothing,
                     double *acc, void *stream);

/* ---------------------------------------------------------------------------------------
 * Validation preprocessing right in front of the forward (val_transforms, hf_dataset_generator.py:27-41:
 * RGB() -> Resize(resize, BICUBIC) -> CenterCrop(crop) -> ToImage() -> ToDtype(float32, scale=True) ->
 * Normalize(mean, std), applied to PIL images).  Input: decoded 8-bit RGB images of any sizes, packed in one device
 * buffer of `pixels_bytes` bytes (4-byte aligned; 3 bytes per pixel, rows contiguous, any image offsets); output: [B, 3, crop_h, crop_w] float32 (bit-identical to the
 * reference transform: Pillow's fixed-point bicubic resample with its two uint8 roundings, then the float32
 * scale / subtract / divide in torchvision's order) or the same values rounded to bf16 for the patcher.
 * `images` is a HOST array (sizes drive the launch geometry); it is copied into the workspace on `stream`.
 * `mean` / `std` are host arrays of 3 floats.  The workspace (device, 16-byte aligned) holds the descriptors, the
 * per-image tap tables and the uint8 intermediate between the two passes; size it with
 * sdp_val_preprocess_workspace_bytes (returns -1 and sets sdp_last_error on bad arguments).
 * --------------------------------------------------------------------------------------- */
typedef struct {
  int64_t offset;          /* byte offset of the image's first pixel in `pixels` */
  int32_t height, width;
} sdp_image_desc;

/* path: 0 = automatic (one fused band kernel whenever the batch's tap tables fit shared memory), 1 = the three-kernel
 * path with the uint8 intermediate in the workspace (what oversized batches take anyway); same results bit for bit.
 * Pass the same value to both calls. */
int64_t sdp_val_preprocess_workspace_bytes(const sdp_image_desc *images, int B, int resize_h, int resize_w,
                                           int crop_h, int crop_w, int path);
int sdp_val_preprocess(const uint8_t *pixels, int64_t pixels_bytes, const sdp_image_desc *images, int B, int resize_h, int resize_w,
                       int crop_h, int crop_w, const float *mean, const float *std, void *workspace,
                       int64_t workspace_bytes, void *out, int out_dtype, int path, void *stream);

/* ---------------------------------------------------------------------------------------
 * Whole-model forward (model.py:129-149) sequenced on the device side of the ABI: one call
 * enqueues every kernel of the forward on `stream`.
 * --------------------------------------------------------------------------------------- */
typedef struct {          /* one EncoderLayer, layers.py:216-257 */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b;
  const float *qn_w, *qn_b, *kn_w, *kn_b;          /* NULL when normalize_qv=False */
  const void *w_qkv;                               /* [3C, C] = cat(q_proj, k_proj, v_proj); * diag(norm1_w) when ln_fold */
  const void *w_o;                                 /* [C, C] */
  const void *w_ff1; const float *b_ff1;           /* [mC, C], [mC] */
  const void *w_ff2; const float *b_ff2;           /* [C, mC], [C] */
  const float *s_qkv, *t_qkv, *s_ff1, *t_ff1;      /* LN-fold vectors ([3C], [3C], [mC], [mC]) or NULL */
  float qk_score_bound;                            /* bound on |q.k|/sqrt(d) after q_norm / k_norm (see sdp_attention_bounded); 0 = unknown */
} sdp_encoder_weights;

typedef struct {          /* one ConvMixer, layers.py:63-99 */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float *w_dw, *b_dw;                        /* tap-major [k*k, C], [C] or NULL */
  const void *w_pw;  const float *b_pw;            /* [C, C] */
  const void *w_mlp1; const float *b_mlp1;         /* [4C, C] */
  const void *w_mlp2; const float *b_mlp2;         /* [C, 4C] */
  const float *s_mlp1, *t_mlp1;                    /* LN-fold vectors [4C] or NULL */
} sdp_mixer_weights;

typedef struct {
  int32_t dtype;                 /* compute dtype of activations and GEMM weights */
  int32_t C, n_head, num_blocks, conv_block_num, ff_mult, conv_k, patch, classes;
  int32_t act, embed_act, conv_first;
  int32_t head_from_register, head_simple;
  int32_t Kp;                    /* padded 3*p*p (pitch of w_patch and of the im2col buffer) */
  int32_t Kc;                    /* padded `classes` (pitch of w_head2 and of the hidden buffer) */
  int32_t ln_fold;               /* 1: token LayerNorms are folded into their consumer GEMMs (bf16 only) */
  const void *w_patch;           /* [C, Kp] */
  const float *pos_table;        /* [T, C] fp32, precomputed for this grid (layers.py:158-163 / :205) */
  const float *reg_table;        /* [R, C] fp32, the R selected register rows */
  const sdp_encoder_weights *enc;   /* HOST array, num_blocks + 1 entries (last = final_block) */
  const sdp_mixer_weights *mix;     /* HOST array, num_blocks * conv_block_num entries */
  const float *head_ln_w, *head_ln_b;
  const void *w_head1; const float *b_head1;       /* [classes, C] */
  const void *w_head2; const float *b_head2;       /* [classes, Kc] or NULL */
} sdp_model_desc;

typedef struct {                 /* caller-allocated device workspaces */
  void *act;      /* [B, S, C] the residual stream (bf16: its hi plane) */
  void *act_lo;   /* [B, S, C] lo plane of the split bf16 residual stream (see sdp_gemm), or NULL: plain bf16 stream */
  void *norm;     /* [B, S, C] */
  void *qkv;      /* [B, S, 3C] */
  void *attn;     /* [B, S, C] */
  void *hidden;   /* [B, S, max(ff_mult,4)*C] */
  void *im2col;   /* [B*T, Kp] */
  void *pooled;   /* [B, C] */
  void *head_h;   /* [B, Kc] */
  float *stats;   /* [B*S, max(1, sdp_gemm_stats_parts(C)), 2] fp32 row statistics (mixers; LN folding) */
} sdp_workspace;

/* x: NCHW [B,3,H,W] in x_dtype; logits: fp32 [B, classes]. */
int sdp_forward(const sdp_model_desc *m, const sdp_workspace *ws, const void *x, int x_dtype,
                int B, int H, int W, int R, float *logits, void *stream);

/* Number of kernels the engine launched since the counter was last reset (all streams). */
int64_t sdp_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SDPNET_B200_H */
