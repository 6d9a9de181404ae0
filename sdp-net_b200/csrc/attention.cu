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
#include <algorithm>

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
// Tensor-core path, version 2: q and k arrive already LayerNorm-ed (fused into the QKV GEMM
// epilogue, gemm_tc.cu head_layernorm), so K and V stream into shared memory with cp.async in
// three commit groups that the first KV blocks overlap with; one CTA per (image, head); every
// warp owns TPW adjacent 16-query tiles and reuses each K / V fragment for all of them, which
// halves the shared-memory traffic per MMA relative to version 1.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int ATT2_TAILW = 10;  // max warps cooperating on a tail tile (bounds the partial buffer)

// NT adjacent 16-query tiles starting at row q0 against KV blocks blk0, blk0+step, ...: flash-style
// online softmax; leaves un-normalised O, running max m and this lane's partial row sums l.
template <int D, int NT, bool GROUP_SYNC>
__device__ __forceinline__ void attn_tiles(const bf16 *__restrict__ base, int C, int S, int q0, int blk0, int step,
                                           int nblk, int gblk, uint32_t sK_addr, uint32_t sV_addr, float scale_log2,
                                           int lane, float (&o)[NT][D / 8][4], float (&m_)[NT][2],
                                           float (&l_)[NT][2]) {
  constexpr int PITCH = D + 8;
  const int g = lane >> 2, qd = lane & 3;
  const int lm = lane >> 3, lr = lane & 7;
  // per-lane ldmatrix row addresses; everything else is a compile-time offset
  // K: matrices (keys 0..7, d 0..7) (same keys, d + 8) (keys + 8, d) (keys + 8, d + 8)
  // V (transposed): (keys 0..7, d 0..7) (keys + 8, same d) (keys, d + 8) (keys + 8, d + 8)
  const uint32_t k_lane = sK_addr + (uint32_t)((((lm >> 1) * 8 + lr) * PITCH + (lm & 1) * 8) * 2);
  const uint32_t v_lane = sV_addr + (uint32_t)((((lm & 1) * 8 + lr) * PITCH + (lm >> 1) * 8) * 2);
  uint32_t qa[NT][D / 16][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int rl = min(q0 + t * 16 + g, S - 1), rh = min(q0 + t * 16 + g + 8, S - 1);
    const bf16 *plo = base + (long long)rl * 3 * C, *phi = base + (long long)rh * 3 * C;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      const int c0 = kk * 16 + qd * 2;
      qa[t][kk][0] = __ldg(reinterpret_cast<const uint32_t *>(plo + c0));
      qa[t][kk][1] = __ldg(reinterpret_cast<const uint32_t *>(phi + c0));
      qa[t][kk][2] = __ldg(reinterpret_cast<const uint32_t *>(plo + c0 + 8));
      qa[t][kk][3] = __ldg(reinterpret_cast<const uint32_t *>(phi + c0 + 8));
    }
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    m_[t][0] = m_[t][1] = -INFINITY;
    l_[t][0] = l_[t][1] = 0.0f;
#pragma unroll
    for (int n = 0; n < D / 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[t][n][j] = 0.0f;
  }
  for (int blk = blk0; blk < nblk; blk += step) {
    if (GROUP_SYNC) {                            // block-uniform: every warp walks all KV blocks in this mode
      if (blk == 0) { cp_async_wait<2>(); __syncthreads(); }
      else if (blk == gblk) { cp_async_wait<1>(); __syncthreads(); }
      else if (blk == 2 * gblk) { cp_async_wait<0>(); __syncthreads(); }
    }
    const int kb = blk * KVB;
    const uint32_t k_blk = k_lane + (uint32_t)blk * (KVB * PITCH * 2);
    const uint32_t v_blk = v_lane + (uint32_t)blk * (KVB * PITCH * 2);
    float sc[NT][KVB / 8][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int n = 0; n < KVB / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[t][n][j] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < KVB / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(k_blk + (np * 16 * PITCH + kk * 16) * 2, b0, b1, b2, b3);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mma_bf16_16816(sc[t][np * 2], qa[t][kk], b0, b1);
          mma_bf16_16816(sc[t][np * 2 + 1], qa[t][kk], b2, b3);
        }
      }
    }
    uint32_t pa[NT][KVB / 16][4];
    const bool last = kb + KVB > S;              // only the final KV block holds padded keys (block-uniform)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (last) {
#pragma unroll
        for (int n = 0; n < KVB / 8; ++n)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kb + n * 8 + qd * 2 + (j & 1) >= S) sc[t][n][j] = -INFINITY;
      }
      float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
      for (int n = 0; n < KVB / 8; ++n) {
        bm_lo = fmaxf(bm_lo, fmaxf(sc[t][n][0], sc[t][n][1]));
        bm_hi = fmaxf(bm_hi, fmaxf(sc[t][n][2], sc[t][n][3]));
      }
      bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1));
      bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
      bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1));
      bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
      bm_lo *= scale_log2;                       // maxima live in the exp2 domain; every block has a live key
      bm_hi *= scale_log2;
      // Lazy rescale: the reference max only moves when a row's block max exceeds it by more than
      // 2^8 (p then stays <= 256, exact enough in fp32 / bf16); O and l always share one reference,
      // so the final O / l is unchanged.  After the first block this almost never fires.
      const bool need = bm_lo > m_[t][0] + 8.0f || bm_hi > m_[t][1] + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float mn_lo = fmaxf(m_[t][0], bm_lo), mn_hi = fmaxf(m_[t][1], bm_hi);
        const float cr_lo = fast_ex2(m_[t][0] - mn_lo), cr_hi = fast_ex2(m_[t][1] - mn_hi);
        m_[t][0] = mn_lo; m_[t][1] = mn_hi;
        l_[t][0] *= cr_lo;
        l_[t][1] *= cr_hi;
#pragma unroll
        for (int n = 0; n < D / 8; ++n) {
          o[t][n][0] *= cr_lo; o[t][n][1] *= cr_lo;
          o[t][n][2] *= cr_hi; o[t][n][3] *= cr_hi;
        }
      }
      const float nm_lo = -m_[t][0], nm_hi = -m_[t][1];
      float ps_lo = 0.0f, ps_hi = 0.0f;
#pragma unroll
      for (int n = 0; n < KVB / 8; ++n) {
        const float p0 = fast_ex2(fmaf(sc[t][n][0], scale_log2, nm_lo)), p1 = fast_ex2(fmaf(sc[t][n][1], scale_log2, nm_lo));
        const float p2 = fast_ex2(fmaf(sc[t][n][2], scale_log2, nm_hi)), p3 = fast_ex2(fmaf(sc[t][n][3], scale_log2, nm_hi));
        ps_lo += p0 + p1;
        ps_hi += p2 + p3;
        // C-fragment of score tile n -> A-fragment of k16 step n/2 (a0,a1 from the even tile; a2,a3 from the odd)
        pa[t][n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[t][n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      l_[t][0] += ps_lo;
      l_[t][1] += ps_hi;
    }
#pragma unroll
    for (int kt = 0; kt < KVB / 16; ++kt) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(v_blk + (kt * 16 * PITCH + np * 16) * 2, b0, b1, b2, b3);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mma_bf16_16816(o[t][np * 2], pa[t][kt], b0, b1);
          mma_bf16_16816(o[t][np * 2 + 1], pa[t][kt], b2, b3);
        }
      }
    }
  }
}

// MAXW warps of TPW tiles each: <96, 1, 16> keeps 16 warps (128 registers) resident per SM, which hides
// the mma.sync / ldmatrix / MUFU latencies far better than 8 fat warps (measured on B200).
template <int D, int TPW, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
attention_bf16_mma2_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int S, int h, float scale_log2) {
  constexpr int PITCH = D + 8;
  constexpr int NV = D / 8;                     // 16-byte vectors per row
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int S_pad = ((S + KVB - 1) / KVB) * KVB;
  const int nblk = S_pad / KVB;
  bf16 *sK = reinterpret_cast<bf16 *>(smem_attn);
  bf16 *sV = sK + (size_t)S_pad * PITCH;
  float *part = reinterpret_cast<float *>(sV + (size_t)S_pad * PITCH);   // tail partials (only if needed)
  const int C = h * D;
  const int b = blockIdx.x / h, head = blockIdx.x % h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nw = nthreads >> 5;
  const bf16 *base = qkv + (long long)b * S * 3 * C + head * D;
  const uint32_t sK_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t sV_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sV));

  // ---- K / V: three cp.async commit groups of whole KV blocks ----
  const int gblk = (nblk + 2) / 3;              // KV blocks per group
  for (int grp = 0; grp < 3; ++grp) {
    const int r0 = min(S_pad, grp * gblk * KVB), r1 = min(S_pad, (grp + 1) * gblk * KVB);
    const int per = (r1 - r0) * NV;
    for (int i = tid; i < 2 * per; i += nthreads) {
      const int which = i >= per;               // 0 = K, 1 = V
      const int j = i - which * per;
      const int r = r0 + j / NV, cv = j % NV;
      const uint32_t off = (uint32_t)(r * PITCH + cv * 8) * 2;
      if (r < S) cp_async16((which ? sV_addr : sK_addr) + off, base + (long long)r * 3 * C + (which + 1) * C + cv * 8);
      else *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(which ? sV : sK) + off) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
  }

  const int g = lane >> 2, qd = lane & 3;
  const int tiles = (S + 15) / 16;
  const int units = (tiles + TPW - 1) / TPW;
  // full rounds: every warp owns one unit of TPW tiles per round (nw <= units by construction)
  const int full_units = (units / nw) * nw;
  bool first = true;
  for (int unit = warp; unit < full_units; unit += nw) {
    float o[TPW][D / 8][4], m_[TPW][2], l_[TPW][2];
    const int q0 = unit * TPW * 16;
    if (first)
      attn_tiles<D, TPW, true>(base, C, S, q0, 0, 1, nblk, gblk, sK_addr, sV_addr, scale_log2, lane, o, m_, l_);
    else
      attn_tiles<D, TPW, false>(base, C, S, q0, 0, 1, nblk, gblk, sK_addr, sV_addr, scale_log2, lane, o, m_, l_);
    first = false;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float ll = l_[t][0], lh = l_[t][1];
      ll += __shfl_xor_sync(0xffffffffu, ll, 1); ll += __shfl_xor_sync(0xffffffffu, ll, 2);
      lh += __shfl_xor_sync(0xffffffffu, lh, 1); lh += __shfl_xor_sync(0xffffffffu, lh, 2);
      const float il = 1.0f / ll, ih = 1.0f / lh;
      const int r_lo = q0 + t * 16 + g, r_hi = r_lo + 8;
      bf16 *olo = out + ((long long)b * S + r_lo) * C + head * D;
      bf16 *ohi = olo + 8LL * C;
#pragma unroll
      for (int n = 0; n < D / 8; ++n) {
        const int c = n * 8 + qd * 2;
        if (r_lo < S) *reinterpret_cast<uint32_t *>(olo + c) = pack_bf16x2(o[t][n][0] * il, o[t][n][1] * il);
        if (r_hi < S) *reinterpret_cast<uint32_t *>(ohi + c) = pack_bf16x2(o[t][n][2] * ih, o[t][n][3] * ih);
      }
    }
  }
  // ---- tail tiles (e.g. the 17th tile of S = 261): all warps split the KV blocks of one tile,
  //      partial (m, l, O) meet in shared memory ----
  const int tw = min(min(nw, nblk), ATT2_TAILW);   // ideally one KV block per cooperating warp
  for (int tile = full_units * TPW; tile < tiles; ++tile) {
    constexpr int PW_ = (D / 8) * 4 * 32 + 4 * 32;        // floats per warp partial
    if (warp < tw) {
      float o[1][D / 8][4], m_[1][2], l_[1][2];
      attn_tiles<D, 1, false>(base, C, S, tile * 16, warp, tw, nblk, gblk, sK_addr, sV_addr, scale_log2, lane, o, m_, l_);
      float ll = l_[0][0], lh = l_[0][1];
      ll += __shfl_xor_sync(0xffffffffu, ll, 1); ll += __shfl_xor_sync(0xffffffffu, ll, 2);
      lh += __shfl_xor_sync(0xffffffffu, lh, 1); lh += __shfl_xor_sync(0xffffffffu, lh, 2);
      float *mine = part + (size_t)warp * PW_;
#pragma unroll
      for (int n = 0; n < D / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) mine[(n * 4 + j) * 32 + lane] = o[0][n][j];
      float *ml = mine + (D / 8) * 4 * 32;
      ml[lane] = m_[0][0]; ml[32 + lane] = m_[0][1]; ml[64 + lane] = ll; ml[96 + lane] = lh;
    }
    __syncthreads();
    float M_lo = -INFINITY, M_hi = -INFINITY;
    for (int w = 0; w < tw; ++w) {
      const float *q = part + (size_t)w * PW_ + (D / 8) * 4 * 32;
      M_lo = fmaxf(M_lo, q[lane]);
      M_hi = fmaxf(M_hi, q[32 + lane]);
    }
    float L_lo = 0.0f, L_hi = 0.0f;
    for (int w = 0; w < tw; ++w) {
      const float *q = part + (size_t)w * PW_ + (D / 8) * 4 * 32;
      L_lo += q[64 + lane] * fast_ex2(q[lane] - M_lo);
      L_hi += q[96 + lane] * fast_ex2(q[32 + lane] - M_hi);
    }
    const float il = 1.0f / L_lo, ih = 1.0f / L_hi;
    const int r_lo = tile * 16 + g, r_hi = r_lo + 8;
    bf16 *olo = out + ((long long)b * S + r_lo) * C + head * D;
    bf16 *ohi = olo + 8LL * C;
    for (int n = warp; n < D / 8; n += nw) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int w = 0; w < tw; ++w) {
        const float *pw = part + (size_t)w * PW_;
        const float *q = pw + (D / 8) * 4 * 32;
        const float f_lo = fast_ex2(q[lane] - M_lo), f_hi = fast_ex2(q[32 + lane] - M_hi);
        a0 = fmaf(pw[(n * 4 + 0) * 32 + lane], f_lo, a0);
        a1 = fmaf(pw[(n * 4 + 1) * 32 + lane], f_lo, a1);
        a2 = fmaf(pw[(n * 4 + 2) * 32 + lane], f_hi, a2);
        a3 = fmaf(pw[(n * 4 + 3) * 32 + lane], f_hi, a3);
      }
      const int c = n * 8 + qd * 2;
      if (r_lo < S) *reinterpret_cast<uint32_t *>(olo + c) = pack_bf16x2(a0 * il, a1 * il);
      if (r_hi < S) *reinterpret_cast<uint32_t *>(ohi + c) = pack_bf16x2(a2 * ih, a3 * ih);
    }
    __syncthreads();
  }
}

template <int D, int TPW, int MAXW>
static int launch_attn_mma2(const void *qkv, void *out, int B, int S, int h, cudaStream_t st) {
  const int tiles = (S + 15) / 16;
  const int units = (tiles + TPW - 1) / TPW;
  const int nw = units < MAXW ? units : MAXW;
  const int S_pad = ((S + KVB - 1) / KVB) * KVB;
  const bool tail = (units / nw) * nw * TPW < tiles;
  const size_t smem = (size_t)2 * S_pad * (D + 8) * sizeof(bf16) +
                      (tail ? (size_t)std::min(std::min(nw, S_pad / KVB), ATT2_TAILW) * ((D / 8) * 4 * 32 + 4 * 32) * sizeof(float) : 0);
  if (smem > 220 * 1024) return -1;             // does not fit: the dispatcher tries the next kernel
  auto kern = attention_bf16_mma2_kernel<D, TPW, MAXW>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  kern<<<B * h, 32 * nw, smem, st>>>((const bf16 *)qkv, (bf16 *)out, S, h, scale_log2);
  SDP_LAUNCH_OK();
  return 0;
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
  if (smem > 220 * 1024) return -1;             // does not fit: the dispatcher falls back to the CUDA-core kernel
  auto kern = attention_bf16_mma_kernel<D>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
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
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
  kern<<<B * h, 32 * nw, smem, st>>>((const T *)qkv, qn_w, qn_b, kn_w, kn_b, (T *)out, S, h, d, eps,
                                     1.0f / sqrtf((float)d));
  SDP_LAUNCH_OK();
  return 0;
}

int attention_tc5(const void *qkv, void *out, int B, int S, int h, int d, float score_bound, cudaStream_t st);   // attention_tc.cu

}  // namespace sdp

using namespace sdp;

static int attention_dispatch(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                              const float *kn_b, void *out, int B, int S, int h, int d, float eps, int dtype,
                              float score_bound, void *stream) {
  SDP_CHECK(qkv && out && B > 0 && S > 0 && h > 0 && d > 0, "sdp_attention: bad arguments");
  SDP_CHECK((qn_w == nullptr) == (kn_w == nullptr) && (qn_w == nullptr) == (qn_b == nullptr) &&
                (kn_w == nullptr) == (kn_b == nullptr),
            "sdp_attention: q/k norm parameters must be all present or all NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SDP_BF16) {
    const bool aligned = (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && ((long long)h * d) % 8 == 0;
    int rc = -1;
    if (aligned && qn_w == nullptr) {
      rc = attention_tc5(qkv, out, B, S, h, d, score_bound, st);      // tcgen05 path: d in {64, 96, 128}, S <= 288
      if (rc >= 0) return rc;
    }
    if (aligned && qn_w == nullptr) {
      switch (d) {
        case 16: rc = launch_attn_mma2<16, 1, 16>(qkv, out, B, S, h, st); break;
        case 32: rc = launch_attn_mma2<32, 1, 16>(qkv, out, B, S, h, st); break;
        case 64: rc = launch_attn_mma2<64, 1, 16>(qkv, out, B, S, h, st); break;
        case 96: rc = launch_attn_mma2<96, 1, 16>(qkv, out, B, S, h, st); break;
        case 128: rc = launch_attn_mma2<128, 1, 8>(qkv, out, B, S, h, st); break;
        default: break;
      }
      if (rc >= 0) return rc;
    }
    if (aligned) {
      switch (d) {
        case 16: rc = launch_attn_mma<16>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st); break;
        case 32: rc = launch_attn_mma<32>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st); break;
        case 64: rc = launch_attn_mma<64>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st); break;
        case 96: rc = launch_attn_mma<96>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st); break;
        case 128: rc = launch_attn_mma<128>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, eps, st); break;
        default: break;
      }
      if (rc >= 0) return rc;
    }
    return launch_attn_simt<bf16>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, d, eps, st);
  }
  SDP_CHECK(dtype == SDP_F32, "sdp_attention: unknown dtype %d", dtype);
  return launch_attn_simt<float>(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, d, eps, st);
}

extern "C" int sdp_attention(const void *qkv, const float *qn_w, const float *qn_b, const float *kn_w,
                             const float *kn_b, void *out, int B, int S, int h, int d, float eps, int dtype,
                             void *stream) {
  return attention_dispatch(qkv, qn_w, qn_b, kn_w, kn_b, out, B, S, h, d, eps, dtype, 0.0f, stream);
}

extern "C" int sdp_attention_bounded(const void *qkv, void *out, int B, int S, int h, int d, float score_bound,
                                     int dtype, void *stream) {
  SDP_CHECK(score_bound >= 0.0f && score_bound == score_bound, "sdp_attention_bounded: score_bound must be >= 0 (0 = unknown)");
  return attention_dispatch(qkv, nullptr, nullptr, nullptr, nullptr, out, B, S, h, d, 0.0f, dtype, score_bound, stream);
}

