// Shared device/host helpers for the sdpnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sdpnet_b200.h"

namespace sdp {

typedef __nv_bfloat16 bf16;

// ---- host-side error plumbing (thread-local message behind sdp_last_error) -------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define SDP_CHECK(cond, ...)          \
  do {                                \
    if (!(cond)) {                    \
      sdp::set_error(__VA_ARGS__);    \
      return 1;                       \
    }                                 \
  } while (0)

#define SDP_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      sdp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                         \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

#define SDP_LAUNCH_OK()                                                              \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      sdp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                            \
      return 3;                                                                      \
    }                                                                                \
    sdp::count_launch();                                                             \
  } while (0)

static inline size_t dtype_size(int dt) { return dt == SDP_BF16 ? 2 : 4; }

// ---- element load/store in either dtype --------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162 *>(&u);
  return __bfloat1622float2(p);
}

// ---- activations (reference model.py:13-24, training_utilities.py:91-92) -----------------
// EXACT selects libdevice erff/tanhf (fp32 verification mode); otherwise the fast forms whose
// error is far below one bf16 ulp of the result.
__device__ __forceinline__ float gelu_erf_exact(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact-erf GELU with ONE special-function op:  gelu(x) = relu(x) - 0.5|x| * erfc(|x|/sqrt2), and
// log2(erfc(t/sqrt2)) is smooth on t >= 0 with value 0 at 0, so erfc = 2^(t*Q(t)) with a degree-4 Q
// (minimax fit on [0, 8] weighted by the sensitivity 0.5 t^2 erfc ln2 of the result: |gelu err| <= 5.4e-7,
// far below bf16 resolution; checked against erff in tests/test_gpu_kernels.py::test_fast_gelu_close_to_erf_gelu).
// 6 FMA-pipe ops + 1 MUFU.EX2 + 2 min/max: the GELU epilogue of the C->4C GEMMs is the hottest non-MMA code of the
// forward.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fabsf(x), 8.0f);
  float q = fmaf(-4.88102407e-04f, t, 7.19871929e-03f);
  q = fmaf(q, t, -5.21466316e-02f);
  q = fmaf(q, t, -4.59595847e-01f);
  q = fmaf(q, t, -1.15100054e+00f);
  const float e = fast_ex2(fmaf(q, t, -1.0f));      // 0.5 erfc(|x|/sqrt2): the 1/2 rides in the exponent
  return fmaf(-t, e, fmaxf(x, 0.0f));
}

// Degree-3 variant for epilogues that round the result to bf16 anyway: |gelu err| <= 8.6e-6 (below one bf16 ulp
// of every output larger than 2e-3 in magnitude; exact at 0).  One FMA fewer per element in the hottest epilogue.
__device__ __forceinline__ float gelu_erf_fast3(float x) {
  const float t = fminf(fabsf(x), 8.0f);
  float q = fmaf(4.16165e-03f, t, -4.573538e-02f);
  q = fmaf(q, t, -4.6493058e-01f);
  q = fmaf(q, t, -1.14956692e+00f);
  const float e = fast_ex2(fmaf(q, t, -1.0f));
  return fmaf(-t, e, fmaxf(x, 0.0f));
}
// ---- packed fp32 pairs (sm_100 FADD2 / FFMA2: one issue slot for two lanes of fp32 math) ----------
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf_fast3 on two values at once, bit-identical to the scalar form: the polynomial runs in s = -t so that the
// last step is a plain packed FMA (relu(x) + s * e); negating t flips the sign of the odd Horner steps only, every
// intermediate has the same magnitude as in gelu_erf_fast3.  4.5 FMA-pipe slots per pair instead of 10.
__device__ __forceinline__ void gelu_erf_fast3_x2(float &x0, float &x1) {
  const uint64_t s = pack_f32x2(fmaxf(-fabsf(x0), -8.0f), fmaxf(-fabsf(x1), -8.0f));
  uint64_t q = fma_f32x2(pack_f32x2(4.16165e-03f, 4.16165e-03f), s, pack_f32x2(4.573538e-02f, 4.573538e-02f));
  q = fma_f32x2(q, s, pack_f32x2(-4.6493058e-01f, -4.6493058e-01f));
  q = fma_f32x2(q, s, pack_f32x2(1.14956692e+00f, 1.14956692e+00f));
  float a0, a1;
  unpack_f32x2(fma_f32x2(q, s, pack_f32x2(-1.0f, -1.0f)), a0, a1);
  const uint64_t r = fma_f32x2(s, pack_f32x2(fast_ex2(a0), fast_ex2(a1)), pack_f32x2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  unpack_f32x2(r, x0, x1);
}

#ifndef SDP_GELU_FAST_FN          // a translation unit whose outputs are bf16 may select gelu_erf_fast3
#define SDP_GELU_FAST_FN gelu_erf_fast
#endif

__device__ __forceinline__ float kelu_f(float x) {
  const float a = 3.5f;
  if (x < -a) return 0.0f;
  if (x > a) return x;
  return 0.5f * x * (1.0f + x / a + 0.31830988618379067154f * sinf(x * (3.14159265358979323846f / a)));
}

template <bool EXACT>
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case SDP_ACT_NONE: return x;
    case SDP_ACT_RELU: return fmaxf(x, 0.0f);
    case SDP_ACT_GELU: return EXACT ? gelu_erf_exact(x) : SDP_GELU_FAST_FN(x);
    case SDP_ACT_GELU_TANH: {
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(u));
    }
    case SDP_ACT_TANH: return tanhf(x);
    case SDP_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case SDP_ACT_LEAKY_RELU: return x > 0.0f ? x : 0.01f * x;
    case SDP_ACT_SELU: {
      const float alpha = 1.6732632423543772848170429916717f;
      const float scale = 1.0507009873554804934193349852946f;
      return scale * (x > 0.0f ? x : alpha * expm1f(x));
    }
    case SDP_ACT_KELU: return kelu_f(x);
    default: return x;
  }
}

// NV values at once: the switch sits outside the (unrolled) element loop, so a run-time activation costs one
// branch per chunk instead of a jump table per element.
template <int NV, bool EXACT>
__device__ __forceinline__ void apply_act_vec(float *v, int act) {
  switch (act) {
    case SDP_ACT_NONE: break;
#define SDP_ACT_CASE(ID)                                                \
    case ID:                                                            \
      _Pragma("unroll") for (int j = 0; j < NV; ++j) v[j] = apply_act<EXACT>(v[j], ID); \
      break;
    SDP_ACT_CASE(SDP_ACT_RELU)
#ifdef SDP_GELU_FAST_X2           // bf16-output epilogues: the packed-pair form of gelu_erf_fast3 (same bits)
    case SDP_ACT_GELU:
      if constexpr (!EXACT && NV % 2 == 0) {
        _Pragma("unroll") for (int j = 0; j < NV; j += 2) gelu_erf_fast3_x2(v[j], v[j + 1]);
      } else {
        _Pragma("unroll") for (int j = 0; j < NV; ++j) v[j] = apply_act<EXACT>(v[j], SDP_ACT_GELU);
      }
      break;
#else
    SDP_ACT_CASE(SDP_ACT_GELU)
#endif
    SDP_ACT_CASE(SDP_ACT_GELU_TANH)
    SDP_ACT_CASE(SDP_ACT_TANH)
    SDP_ACT_CASE(SDP_ACT_SIGMOID)
    SDP_ACT_CASE(SDP_ACT_LEAKY_RELU)
    SDP_ACT_CASE(SDP_ACT_SELU)
    SDP_ACT_CASE(SDP_ACT_KELU)
#undef SDP_ACT_CASE
    default: break;
  }
}

// ---- GEMM epilogue description (shared by the tcgen05 and the CUDA-core GEMM) -----------
struct Epilogue {
  const float *bias;
  const void *residual;
  void *out;
  const void *res_lo;        // split residual stream (bf16 only): value = residual + res_lo, same pitch
  void *out_lo;              // ... and out = bf16(v), out_lo = bf16(v - out)
  long long ldr, ldo;
  int M, N;
  int out_dtype, res_dtype;
  int act, res_first, res_mod;
  int seq_in, seq_out, seq_off;
  int pass_seq, pass_rows;
  int hn_d, hn_C;
  float hn_eps;
  const float *hn_qw, *hn_qb, *hn_kw, *hn_kb;
  float *stats_out;          // producer: per-row column-part (sum, sumsq) of the stored bf16 values
  int stats_parts;
  const float *ln_stats;     // consumer: LayerNorm folded into this GEMM
  int ln_parts, ln_K;
  float ln_eps;
  const float *ln_s, *ln_t;
};

struct RowMap {
  long long ro, rr;   // output row, residual row
  bool live;          // row < M and not a pass-through (register) row
  bool pass;          // row < M and a pass-through row (keeps its residual value)
};

__device__ __forceinline__ RowMap map_row(const Epilogue &e, int r) {
  RowMap m;
  m.live = r < e.M;
  const int rc = m.live ? r : 0;
  m.ro = e.seq_in ? (long long)(rc / e.seq_in) * e.seq_out + e.seq_off + rc % e.seq_in : rc;
  m.rr = e.res_mod ? rc % e.res_mod : m.ro;
  m.pass = false;
  if (m.live && e.pass_seq && (int)(m.ro % e.pass_seq) < e.pass_rows) {
    m.live = false;
    m.pass = true;
  }
  return m;
}

// Finish NV consecutive columns [c0, c0+NV) of one row held in v[] (fp32 accumulators).
// NV is a multiple of 8.  VEC = the pointers/pitches allow 16-byte accesses.
template <int NV, bool EXACT, int ACT>   // ACT < 0: runtime e.act
__device__ __forceinline__ void epilogue_math(const Epilogue &e, const RowMap &m, int c0, float *v, bool vec) {
  const int act = ACT < 0 ? e.act : ACT;
  const bool full = vec && (c0 + NV <= e.N);
  if (e.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(e.bias + c0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c0 + j < e.N) v[j] += __ldg(e.bias + c0 + j);
    }
  }
  if (e.residual == nullptr || !e.res_first) {
    apply_act_vec<NV, EXACT>(v, act);
  }
  if (e.residual) {
    if (e.res_dtype == SDP_BF16) {
      const bf16 *rp = reinterpret_cast<const bf16 *>(e.residual) + m.rr * e.ldr + c0;
      const bf16 *lp = e.res_lo ? reinterpret_cast<const bf16 *>(e.res_lo) + m.rr * e.ldr + c0 : nullptr;
      if (full) {
#pragma unroll
        for (int j = 0; j < NV; j += 8) {
          const uint4 u = *reinterpret_cast<const uint4 *>(rp + j);
          float2 f;
          float r[8];
          f = unpack_bf16x2(u.x); r[0] = f.x; r[1] = f.y;
          f = unpack_bf16x2(u.y); r[2] = f.x; r[3] = f.y;
          f = unpack_bf16x2(u.z); r[4] = f.x; r[5] = f.y;
          f = unpack_bf16x2(u.w); r[6] = f.x; r[7] = f.y;
          if (lp) {
            const uint4 w = *reinterpret_cast<const uint4 *>(lp + j);
            f = unpack_bf16x2(w.x); r[0] += f.x; r[1] += f.y;
            f = unpack_bf16x2(w.y); r[2] += f.x; r[3] += f.y;
            f = unpack_bf16x2(w.z); r[4] += f.x; r[5] += f.y;
            f = unpack_bf16x2(w.w); r[6] += f.x; r[7] += f.y;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) v[j + q] += r[q];
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (c0 + j < e.N) v[j] += __bfloat162float(rp[j]) + (lp ? __bfloat162float(lp[j]) : 0.0f);
      }
    } else {
      const float *rp = reinterpret_cast<const float *>(e.residual) + m.rr * e.ldr + c0;
      if (full) {
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
          const float4 u = *reinterpret_cast<const float4 *>(rp + j);
          v[j] += u.x; v[j + 1] += u.y; v[j + 2] += u.z; v[j + 3] += u.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (c0 + j < e.N) v[j] += rp[j];
      }
    }
    if (e.res_first) {
      apply_act_vec<NV, EXACT>(v, act);
    }
  }
}

template <int NV>
__device__ __forceinline__ void epilogue_store(const Epilogue &e, const RowMap &m, int c0, const float *v, bool vec) {
  const bool full = vec && (c0 + NV <= e.N);
  if (e.out_dtype == SDP_BF16) {
    bf16 *op = reinterpret_cast<bf16 *>(e.out) + m.ro * e.ldo + c0;
    bf16 *lp = e.out_lo ? reinterpret_cast<bf16 *>(e.out_lo) + m.ro * e.ldo + c0 : nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[j], v[j + 1]);
        u.y = pack_bf16x2(v[j + 2], v[j + 3]);
        u.z = pack_bf16x2(v[j + 4], v[j + 5]);
        u.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4 *>(op + j) = u;
        if (lp) {                            // what the bf16 rounding dropped, itself rounded to bf16
          const uint32_t hw[4] = {u.x, u.y, u.z, u.w};
          uint32_t lw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            lw[q] = pack_bf16x2(v[j + 2 * q] - __uint_as_float(hw[q] << 16),
                                v[j + 2 * q + 1] - __uint_as_float(hw[q] & 0xffff0000u));
          *reinterpret_cast<uint4 *>(lp + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c0 + j < e.N) {
          const bf16 h = __float2bfloat16_rn(v[j]);
          op[j] = h;
          if (lp) lp[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
        }
    }
  } else {
    float *op = reinterpret_cast<float *>(e.out) + m.ro * e.ldo + c0;
    if (full) {
#pragma unroll
      for (int j = 0; j < NV; j += 4)
        *reinterpret_cast<float4 *>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c0 + j < e.N) op[j] = v[j];
    }
  }
}

// bias / activation / residual, then a direct global store (row-per-thread layout)
template <int NV, bool EXACT, int ACT>
__device__ __forceinline__ void epilogue_row(const Epilogue &e, const RowMap &m, int c0, float *v, bool vec) {
  if (!m.live) return;
  epilogue_math<NV, EXACT, ACT>(e, m, c0, v, vec);
  epilogue_store<NV>(e, m, c0, v, vec);
}

// Pass-through row in a staged (TMA-store) epilogue: the tile box is written as a whole, so the
// row re-emits its residual (== current output, in place) bit for bit; rows past M emit zeros
// (clipped by the TMA store anyway).
template <int NV>
__device__ __forceinline__ void epilogue_passthrough(const Epilogue &e, const RowMap &m, int c0, float *v, bool vec) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = 0.0f;
  if (!m.pass || e.residual == nullptr) return;
  if (e.res_dtype == SDP_BF16) {
    const bf16 *rp = reinterpret_cast<const bf16 *>(e.residual) + m.rr * e.ldr + c0;
    if (vec && c0 + NV <= e.N) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        const uint4 u = *reinterpret_cast<const uint4 *>(rp + j);
        float2 f;
        f = unpack_bf16x2(u.x); v[j] = f.x; v[j + 1] = f.y;
        f = unpack_bf16x2(u.y); v[j + 2] = f.x; v[j + 3] = f.y;
        f = unpack_bf16x2(u.z); v[j + 4] = f.x; v[j + 5] = f.y;
        f = unpack_bf16x2(u.w); v[j + 6] = f.x; v[j + 7] = f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (c0 + j < e.N) v[j] = __bfloat162float(rp[j]);
    }
  } else {
    const float *rp = reinterpret_cast<const float *>(e.residual) + m.rr * e.ldr + c0;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (c0 + j < e.N) v[j] = rp[j];
  }
}

// Host: can the epilogue use 16-byte accesses for every row?
static inline bool epilogue_vec_ok(const Epilogue &e) {
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const size_t os = dtype_size(e.out_dtype), rs = dtype_size(e.res_dtype);
  bool ok = al(e.out) && (e.ldo * os) % 16 == 0;
  if (e.residual) ok = ok && al(e.residual) && (e.ldr * rs) % 16 == 0;
  if (e.res_lo) ok = ok && al(e.res_lo);
  if (e.out_lo) ok = ok && al(e.out_lo);
  if (e.bias) ok = ok && al(e.bias);
  return ok;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace sdp
