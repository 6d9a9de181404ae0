// Fused QK-LayerNorm + softmax attention over the short (R + T <= ~300 token) sequence.
// Reference: layers.py:282-300 -- q/k/v views of the projections, per-head nn.LayerNorm(d) on q and
// k (layers.py:236-237,286), F.scaled_dot_product_attention without mask or dropout (eval).
//
// Input is the raw QKV GEMM output [B, S, 3C] (q | k | v column blocks, head-major inside each);
// output is token-major [B, S, C], i.e. already the A operand of the o_proj GEMM -- the reference's
// view/transpose/contiguous round trips (layers.py:282-284,300) never materialise.
//
// bf16 path (round 1): one CTA per (image, head, query half); K (normalised on load) and V of the
// whole sequence live in shared memory, each warp owns one 16-query tile and runs flash-style
// online softmax over 32-key blocks with mma.sync.m16n8k16 (fp32 accumulate), exp2 with the
// 1/sqrt(d)*log2(e) scale folded in.  The tcgen05 version of this kernel is the planned upgrade;
// attention is 2.3 % of the model FLOPs.
// fp32 / odd head_dim path: CUDA-core kernel, one warp per query row (verification mode).
#include "common.cuh"

namespace sdp {

// ---------------------------------------------------------------------------------------
// Tensor-core path
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2,
                                                  uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

constexpr int KVB = 32;   // keys per online-softmax block

template <int D>
__global__ void __launch_bounds__(320)
attention_bf16_mma_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ qn_w,
                          const float *__restrict__ qn_b, const float *__restrict__ kn_w,
                          const float *__restrict__ kn_b, bf16 *__restrict__ out, int S, int h, float eps,
                          float scale_log2, int tiles_per_cta) {
  constexpr int PITCH = D + 8;                 // bf16 elements; +16 B keeps ldmatrix rows conflict-free
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int S_pad = ((S + KVB - 1) / KVB) * KVB;
  bf16 *sK = reinterpret_cast<bf16 *>(smem_attn);
  bf16 *sV = sK + (size_t)S_pad * PITCH;
  const int C = h * D;
  const int bh = blockIdx.x;
  const int b = bh / h, head = bh % h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x;
  const bf16 *base = qkv + (long long)b * S * 3 * C + head * D;

  // ---- stage K (LayerNorm over d applied on the way in) and V; 16 lanes per row, 16 B per lane ----
  {
    constexpr int LPR = 16;                    // lanes per row (D <= 128 -> <= 16 vectors of 8)
    constexpr int NV = D / 8;
    const int sub = tid % LPR;
    for (int r = tid / LPR; r < S_pad; r += nthreads / LPR) {
      float kf[8];
      uint4 vraw = make_uint4(0, 0, 0, 0);
      const bool live = r < S && sub < NV;
      if (live) {
        const bf16 *rowp = base + (long long)r * 3 * C;
        const uint4 kraw = *reinterpret_cast<const uint4 *>(rowp + C + sub * 8);
        vraw = *reinterpret_cast<const uint4 *>(rowp + 2 * C + sub * 8);
        float2 f;
        f = unpack_bf16x2(kraw.x); kf[0] = f.x; kf[1] = f.y;
        f = unpack_bf16x2(kraw.y); kf[2] = f.x; kf[3] = f.y;
        f = unpack_bf16x2(kraw.z); kf[4] = f.x; kf[5] = f.y;
        f = unpack_bf16x2(kraw.w); kf[6] = f.x; kf[7] = f.y;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) kf[j] = 0.0f;
      }
      if (kn_w != nullptr) {                   // warp-uniform
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += kf[j];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)D;
        float q = 0.0f;
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float dlt = kf[j] - mean;
            q = fmaf(dlt, dlt, q);
          }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.0f / sqrtf(q / (float)D + eps);
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            kf[j] = (kf[j] - mean) * rstd * __ldg(kn_w + sub * 8 + j) + __ldg(kn_b + sub * 8 + j);
        }
      }
      if (sub < NV) {
        uint4 kout;
        kout.x = pack_bf16x2(kf[0], kf[1]); kout.y = pack_bf16x2(kf[2], kf[3]);
        kout.z = pack_bf16x2(kf[4], kf[5]); kout.w = pack_bf16x2(kf[6], kf[7]);
        if (r >= S) kout = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(sK + (size_t)r * PITCH + sub * 8) = kout;
        *reinterpret_cast<uint4 *>(sV + (size_t)r * PITCH + sub * 8) = vraw;   // zero rows past S
      }
    }
  }
  __syncthreads();

  // ---- this warp's 16-query tile ----
  const int qtile = blockIdx.y * tiles_per_cta + warp;
  const int q0 = qtile * 16;
  if (warp >= tiles_per_cta || q0 >= S) return;      // no further block-level syncs below
  const int g = lane >> 2, qd = lane & 3;
  const int row_lo = min(q0 + g, S - 1), row_hi = min(q0 + g + 8, S - 1);

  // Q fragments (A operand, 16 x D), LayerNorm over d in registers
  uint32_t qa[D / 16][4];
  {
    float ql[D / 16][4], qh[D / 16][4];
    const bf16 *plo = base + (long long)row_lo * 3 * C;
    const bf16 *phi = base + (long long)row_hi * 3 * C;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      const int c0 = kk * 16 + qd * 2;
      float2 f;
      f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(plo + c0));     ql[kk][0] = f.x; ql[kk][1] = f.y;
      f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(plo + c0 + 8)); ql[kk][2] = f.x; ql[kk][3] = f.y;
      f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(phi + c0));     qh[kk][0] = f.x; qh[kk][1] = f.y;
      f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(phi + c0 + 8)); qh[kk][2] = f.x; qh[kk][3] = f.y;
    }
    if (qn_w != nullptr) {
      float sl = 0.0f, sh = 0.0f;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
#pragma unroll
        for (int j = 0; j < 4; ++j) { sl += ql[kk][j]; sh += qh[kk][j]; }
      sl += __shfl_xor_sync(0xffffffffu, sl, 1); sl += __shfl_xor_sync(0xffffffffu, sl, 2);
      sh += __shfl_xor_sync(0xffffffffu, sh, 1); sh += __shfl_xor_sync(0xffffffffu, sh, 2);
      const float ml = sl / (float)D, mh = sh / (float)D;
      float vl = 0.0f, vh = 0.0f;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = ql[kk][j] - ml, c = qh[kk][j] - mh;
          vl = fmaf(a, a, vl);
          vh = fmaf(c, c, vh);
        }
      vl += __shfl_xor_sync(0xffffffffu, vl, 1); vl += __shfl_xor_sync(0xffffffffu, vl, 2);
      vh += __shfl_xor_sync(0xffffffffu, vh, 1); vh += __shfl_xor_sync(0xffffffffu, vh, 2);
      const float rl = 1.0f / sqrtf(vl / (float)D + eps), rh = 1.0f / sqrtf(vh / (float)D + eps);
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = kk * 16 + qd * 2 + (j & 1) + (j >> 1) * 8;
          const float w = __ldg(qn_w + c), bb = __ldg(qn_b + c);
          ql[kk][j] = (ql[kk][j] - ml) * rl * w + bb;
          qh[kk][j] = (qh[kk][j] - mh) * rh * w + bb;
        }
    }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      qa[kk][0] = pack_bf16x2(ql[kk][0], ql[kk][1]);   // (row g   , k 2q..2q+1)
      qa[kk][1] = pack_bf16x2(qh[kk][0], qh[kk][1]);   // (row g+8 , k 2q..2q+1)
      qa[kk][2] = pack_bf16x2(ql[kk][2], ql[kk][3]);   // (row g   , k 2q+8..)
      qa[kk][3] = pack_bf16x2(qh[kk][2], qh[kk][3]);   // (row g+8 , k 2q+8..)
    }
  }

  float o[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[n][j] = 0.0f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.0f, l_hi = 0.0f;

  const uint32_t sK_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t sV_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const int lm = lane >> 3, lr = lane & 7;           // ldmatrix: matrix index / row inside it

  for (int kb = 0; kb < S_pad; kb += KVB) {
    // ---- scores: 16 queries x 32 keys ----
    float sc[KVB / 8][4];
#pragma unroll
    for (int n = 0; n < KVB / 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[n][j] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < KVB / 16; ++np) {
        // matrices: (keys np*16 + 0..7, d kk*16 + 0..7) (same keys, d + 8) (keys + 8, d) (keys + 8, d + 8)
        const int key = kb + np * 16 + (lm >> 1) * 8 + lr;
        const int col = kk * 16 + (lm & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sK_addr + (uint32_t)(key * PITCH + col) * 2, b0, b1, b2, b3);
        mma_bf16_16816(sc[np * 2], qa[kk], b0, b1);
        mma_bf16_16816(sc[np * 2 + 1], qa[kk], b2, b3);
      }
    }
    // ---- scale, mask the padded keys, online softmax ----
    float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
    for (int n = 0; n < KVB / 8; ++n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kb + n * 8 + qd * 2 + (j & 1);
        const float v = key < S ? sc[n][j] * scale_log2 : -INFINITY;
        sc[n][j] = v;
        if (j < 2) bm_lo = fmaxf(bm_lo, v); else bm_hi = fmaxf(bm_hi, v);
      }
    }
    bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1));
    bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
    bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1));
    bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
    const float mn_lo = fmaxf(m_lo, bm_lo), mn_hi = fmaxf(m_hi, bm_hi);   // finite: every block has >= 1 live key
    const float cr_lo = exp2f(m_lo - mn_lo), cr_hi = exp2f(m_hi - mn_hi);
    m_lo = mn_lo; m_hi = mn_hi;
    float ps_lo = 0.0f, ps_hi = 0.0f;
    uint32_t pa[KVB / 16][4];
#pragma unroll
    for (int n = 0; n < KVB / 8; ++n) {
      const float p0 = exp2f(sc[n][0] - mn_lo), p1 = exp2f(sc[n][1] - mn_lo);
      const float p2 = exp2f(sc[n][2] - mn_hi), p3 = exp2f(sc[n][3] - mn_hi);
      ps_lo += p0 + p1;
      ps_hi += p2 + p3;
      // C-fragment of score tile n -> A-fragment of k16 step n/2 (a0,a1 from even tile; a2,a3 from odd)
      pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l_lo = l_lo * cr_lo + ps_lo;
    l_hi = l_hi * cr_hi + ps_hi;
#pragma unroll
    for (int n = 0; n < D / 8; ++n) {
      o[n][0] *= cr_lo; o[n][1] *= cr_lo;
      o[n][2] *= cr_hi; o[n][3] *= cr_hi;
    }
    // ---- O += P . V ----
#pragma unroll
    for (int t = 0; t < KVB / 16; ++t) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        // transposed matrices: (keys t*16 + 0..7, d np*16 + 0..7) (keys + 8, same d) (keys, d + 8) (keys + 8, d + 8)
        const int key = kb + t * 16 + (lm & 1) * 8 + lr;
        const int col = np * 16 + (lm >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(sV_addr + (uint32_t)(key * PITCH + col) * 2, b0, b1, b2, b3);
        mma_bf16_16816(o[np * 2], pa[t], b0, b1);
        mma_bf16_16816(o[np * 2 + 1], pa[t], b2, b3);
      }
    }
  }
  // the per-lane partial row sums cover this lane's key columns only: reduce over the quad
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float il = 1.0f / l_lo, ih = 1.0f / l_hi;
  bf16 *olo = out + ((long long)b * S + q0 + g) * C + head * D;
  bf16 *ohi = olo + 8LL * C;
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int c = n * 8 + qd * 2;
    if (q0 + g < S) *reinterpret_cast<uint32_t *>(olo + c) = pack_bf16x2(o[n][0] * il, o[n][1] * il);
    if (q0 + g + 8 < S) *reinterpret_cast<uint32_t *>(ohi + c) = pack_bf16x2(o[n][2] * ih, o[n][3] * ih);
  }
}

// ---------------------------------------------------------------------------------------
// CUDA-core path: fp32 verification mode and head dims the mma path does not cover.
// grid (S, B*h) is too many tiny CTAs; use one CTA per (b, head) with 8 warps striding the queries.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T *__restrict__ qkv, const float *__restrict__ qn_w, const float *__restrict__ qn_b,
                      const float *__restrict__ kn_w, const float *__restrict__ kn_b, T *__restrict__ out, int S,
                      int h, int d, float eps, float scale) {
  extern __shared__ float sm[];
  // per warp: q[d] and p[S]; shared: normalised K [S][d+1]
  const int C = h * d;
  const int b = blockIdx.x / h, head = blockIdx.x % h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int kp = d + 1;
  float *sK = sm;
  float *sQ = sK + (size_t)S * kp + (size_t)warp * (d + S);
  float *sP = sQ + d;
  const T *base = qkv + (long long)b * S * 3 * C + head * d;

  auto norm_row = [&](const T *src, float *dst, const float *w, const float *bb) {
    float s = 0.0f;
    for (int c = lane; c < d; c += 32) s += to_f(src[c]);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.0f;
    for (int c = lane; c < d; c += 32) {
      const float t = to_f(src[c]) - mean;
      q = fmaf(t, t, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + eps);
    for (int c = lane; c < d; c += 32) {
      const float v = to_f(src[c]);
      dst[c] = w ? (v - mean) * rstd * __ldg(w + c) + __ldg(bb + c) : v;
    }
  };
  for (int r = warp; r < S; r += nw) norm_row(base + (long long)r * 3 * C + C, sK + (size_t)r * kp, kn_w, kn_b);
  __syncthreads();
  for (int qi = warp; qi < S; qi += nw) {
    norm_row(base + (long long)qi * 3 * C, sQ, qn_w, qn_b);
    __syncwarp();
    float mx = -INFINITY;
    for (int key = lane; key < S; key += 32) {
      float acc = 0.0f;
      const float *kr = sK + (size_t)key * kp;
      for (int c = 0; c < d; ++c) acc = fmaf(sQ[c], kr[c], acc);
      acc *= scale;
      sP[key] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int key = lane; key < S; key += 32) {
      const float p = expf(sP[key] - mx);
      sP[key] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int c = lane; c < d; c += 32) {
      float acc = 0.0f;
      const T *vp = base + 2 * C + c;
      for (int key = 0; key < S; ++key) acc = fmaf(sP[key], to_f(vp[(long long)key * 3 * C]), acc);
      out[((long long)b * S + qi) * C + head * d + c] = from_f<T>(acc * inv);
    }
    __syncwarp();
  }
}

template <int D>
static int launch_attn_mma(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                           const float *kn_b, void *out, int B, int S, int h, float eps, cudaStream_t st) {
  const int tiles = (S + 15) / 16;
  const int ny = (tiles + 9) / 10;                    // <= 10 warps (query tiles) per CTA
  const int tpc = (tiles + ny - 1) / ny;
  const int S_pad = ((S + KVB - 1) / KVB) * KVB;
  const size_t smem = (size_t)2 * S_pad * (D + 8) * sizeof(bf16);
  SDP_CHECK(smem <= 220 * 1024, "sdp_attention: S=%d d=%d needs %zu B of shared memory", S, D, smem);
  auto kern = attention_bf16_mma_kernel<D>;
  static size_t configured = 0;
  if (smem > configured) {
    SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  dim3 grid(B * h, ny);
  kern<<<grid, 32 * tpc, smem, st>>>((const bf16 *)qkv, qn_w, qn_b, kn_w, kn_b, (bf16 *)out, S, h, eps,
                                     scale_log2, tpc);
  SDP_LAUNCH_OK();
  return 0;
}

template <typename T>
static int launch_attn_simt(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                            const float *kn_b, void *out, int B, int S, int h, int d, float eps, cudaStream_t st) {
  const int nw = 8;
  const size_t smem = ((size_t)S * (d + 1) + (size_t)nw * (d + S)) * sizeof(float);
  SDP_CHECK(smem <= 220 * 1024, "sdp_attention(simt): S=%d d=%d needs %zu B of shared memory", S, d, smem);
  auto kern = attention_simt_kernel<T>;
  static size_t configured = 0;
  if (smem > configured) {
    SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  kern<<<B * h, 32 * nw, smem, st>>>((const T *)qkv, qn_w, qn_b, kn_w, kn_b, (T *)out, S, h, d, eps,
                                     1.0f / sqrtf((float)d));
  SDP_LAUNCH_OK();
  return 0;
}

}  // namespace sdp

using namespace sdp;

extern "C" int sdp_attention(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                             const float *kn_b, void *out, int B, int S, int h, int d, float eps, int dtype,
                             void *stream) {
  SDP_CHECK(qkv && out && B > 0 && S > 0 && h > 0 && d > 0, "sdp_attention: bad arguments");
  SDP_CHECK((qn_w == nullptr) == (kn_w == nullptr) && (qn_w == nullptr) == (qn_b == nullptr) &&
                (kn_w == nullptr) == (kn_b == nullptr),
            "sdp_attention: q/k norm parameters must be all present or all NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SDP_BF16) {
    const bool aligned = (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ((long long)h * d) % 8 == 0;
    if (aligned) {
      switch (d) {
        case 16: return launch_attn_mma<16>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st);
        case 32: return launch_attn_mma<32>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st);
        case 64: return launch_attn_mma<64>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st);
        case 96: return launch_attn_mma<96>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st);
        case 128: return launch_attn_mma<128>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st);
        default: break;
      }
    }
    return launch_attn_simt<bf16>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, d, eps, st);
  }
  SDP_CHECK(dtype == SDP_F32, "sdp_attention: unknown dtype %d", dtype);
  return launch_attn_simt<float>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, d, eps, st);
}
