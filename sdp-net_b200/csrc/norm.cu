// Bandwidth-bound row kernels: token LayerNorm, register fill, pooled head front, layout bridges,
// patch im2col, elementwise activation.  One pass over HBM each, 128-bit accesses where the shape
// allows.  Reference call sites are cited per entry point in include/sdpnet_b200.h.
#include "common.cuh"

namespace sdp {

// ---------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, row cached in registers (single HBM read).
// ---------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  using Raw = float4;
  __device__ static void cvt(const float4 &u, float *v) { v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
  __device__ static void load(const float *p, float *v) {
    const float4 u = *reinterpret_cast<const float4 *>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
  __device__ static void store(float *p, const float *v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  using Raw = uint4;
  __device__ static void cvt(const uint4 &u, float *v) {
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
  }
  __device__ static void load(const bf16 *p, float *v) {
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
  }
  __device__ static void store(bf16 *p, const float *v) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(p) = u;
  }
};

// Persistent: a warp walks rows with a grid stride and has the NEXT row's loads in flight while it reduces,
// normalises and stores the current one (two rows of memory-level parallelism per warp, no block churn).
// Rows are walked from the LAST to the first: the producer GEMM wrote them in ascending order, so its most recent
// ~100 MB are still in L2 when this kernel starts, and the consumer GEMM starts at row 0, which this kernel wrote last.
template <typename T, int VPL>
__global__ void __launch_bounds__(256)
ln_rows_vec_kernel(const T *__restrict__ x, long long ldx, const float *__restrict__ w,
                   const float *__restrict__ b, T *__restrict__ out, long long ldo, int M, int C, float eps) {
  constexpr int EPV = Vec<T>::N;
  using Raw = typename Vec<T>::Raw;
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C / EPV;
  Raw cur[VPL], nxt[VPL];
  auto load_row = [&](int r, Raw *dst) {
    const Raw *xr = reinterpret_cast<const Raw *>(x + (long long)(M - 1 - r) * ldx);
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + 32 * i < nvec) dst[i] = xr[lane + 32 * i];
  };
  load_row(row, cur);
  while (true) {
    const int next = row + nwarps;
    if (next < M) load_row(next, nxt);
    float v[VPL][EPV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < nvec) {
        Vec<T>::cvt(cur[i], v[i]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) s += v[i][j];
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < EPV; ++j) {
          const float d = v[i][j] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
    T *orow = out + (long long)(M - 1 - row) * ldo;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int iv = lane + 32 * i;
      if (iv < nvec) {
        float y[EPV];
        const int c0 = iv * EPV;
#pragma unroll
        for (int j = 0; j < EPV; j += 4) {          // affine parameters as 128-bit loads (L1/L2 resident)
          float4 wv = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (w) wv = __ldg(reinterpret_cast<const float4 *>(w + c0 + j));
          if (b) bv = __ldg(reinterpret_cast<const float4 *>(b + c0 + j));
          y[j] = (v[i][j] - mean) * rstd * wv.x + bv.x;
          y[j + 1] = (v[i][j + 1] - mean) * rstd * wv.y + bv.y;
          y[j + 2] = (v[i][j + 2] - mean) * rstd * wv.z + bv.z;
          y[j + 3] = (v[i][j + 3] - mean) * rstd * wv.w + bv.w;
        }
        Vec<T>::store(orow + iv * EPV, y);
      }
    }
    if (next >= M) break;
    row = next;
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
  }
}

// any C / alignment: three passes over the (L1-resident) row
template <typename T>
__global__ void __launch_bounds__(256)
ln_rows_generic_kernel(const T *__restrict__ x, long long ldx, const float *__restrict__ w,
                       const float *__restrict__ b, T *__restrict__ out, long long ldo, int M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T *xr = x + (long long)row * ldx;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f(xr[c]) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
  T *orow = out + (long long)row * ldo;
  for (int c = lane; c < C; c += 32) {
    float t = (to_f(xr[c]) - mean) * rstd;
    if (w) t *= __ldg(w + c);
    if (b) t += __ldg(b + c);
    orow[c] = from_f<T>(t);
  }
}

template <typename T>
static int launch_ln_rows(const void *x, long long ldx, const float *w, const float *b, void *out, long long ldo,
                          int M, int C, float eps, cudaStream_t st) {
  constexpr int EPV = Vec<T>::N;
  const T *xp = reinterpret_cast<const T *>(x);
  T *op = reinterpret_cast<T *>(out);
  const int wpb = 8;
  int dev = 0, sms = 0;                   // per call: a process may drive several devices
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int full = (M + wpb - 1) / wpb;
  const dim3 grid(full), block(32 * wpb);
  auto persistent_grid = [&](const void *kern) {         // one resident wave of the vector kernel
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * wpb, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    const int g = per_sm * sms;
    return dim3(full < g ? full : g);
  };
  const bool vec = C % EPV == 0 && (ldx * sizeof(T)) % 16 == 0 && (ldo * sizeof(T)) % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
  const int need = vec ? (C / EPV + 31) / 32 : 99;
#define LN_CASE(V)                                                                             \
  if (need <= V) {                                                                             \
    const dim3 pg = persistent_grid(reinterpret_cast<const void *>(&ln_rows_vec_kernel<T, V>));   \
    ln_rows_vec_kernel<T, V><<<full < (int)pg.x ? dim3(full) : pg, block, 0, st>>>(xp, ldx, w, b, op, ldo, M, C, eps); \
    SDP_LAUNCH_OK();                                                                           \
    return 0;                                                                                  \
  }
  LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8)
#undef LN_CASE
  ln_rows_generic_kernel<T><<<grid, block, 0, st>>>(xp, ldx, w, b, op, ldo, M, C, eps);
  SDP_LAUNCH_OK();
  return 0;
}

// Row (sum, sum of squares) in the producer-GEMM statistics layout: part 0 = full sums, others 0.
template <typename T>
__global__ void __launch_bounds__(256)
row_stats_kernel(const T *__restrict__ x, long long ldx, float *__restrict__ stats, int parts, int M, int C) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T *xr = x + (long long)row * ldx;
  float s1 = 0.0f, s2 = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float v = to_f(xr[c]);
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  float *o = stats + (long long)row * parts * 2;
  for (int p = lane; p < parts; p += 32) {
    o[2 * p] = p == 0 ? s1 : 0.0f;
    o[2 * p + 1] = p == 0 ? s2 : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void fill_registers_kernel(T *act, T *act_lo, const float *__restrict__ table, int B, int S, int R, int C) {
  const long long n = (long long)B * R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int r = (i / C) % R;
    const long long b = i / ((long long)C * R);
    const float v = __ldg(table + r * C + c);
    const T h = from_f<T>(v);
    act[(b * S + r) * C + c] = h;
    if (act_lo) act_lo[(b * S + r) * C + c] = from_f<T>(v - to_f(h));     // split stream: what the rounding dropped
  }
}

// out[b, :] = LN(mean_{rows}(act[b, row0:row0+nrows, :])); one CTA per image
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
pool_ln_kernel(const T *__restrict__ act, const T *__restrict__ act_lo, int S, int C, int row0, int nrows,
               const float *__restrict__ lw, const float *__restrict__ lb, float eps, TO *__restrict__ out, long long ldo) {
  extern __shared__ float pooled[];   // [C] + 32 scratch
  float *red = pooled + C;
  const int b = blockIdx.x;
  const T *base = act + ((long long)b * S + row0) * C;
  const T *base_lo = act_lo ? act_lo + ((long long)b * S + row0) * C : nullptr;
  float s = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.0f;
    for (int r = 0; r < nrows; ++r)
      a += to_f(base[(long long)r * C + c]) + (base_lo ? to_f(base_lo[(long long)r * C + c]) : 0.0f);
    a /= (float)nrows;
    pooled[c] = a;
    s += a;
  }
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
  };
  float mean = 0.0f, rstd = 1.0f;
  if (lw) {
    mean = block_sum(s) / (float)C;
    float q = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float d = pooled[c] - mean;
      q = fmaf(d, d, q);
    }
    rstd = 1.0f / sqrtf(block_sum(q) / (float)C + eps);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = pooled[c];
    if (lw) {
      t = (t - mean) * rstd * __ldg(lw + c);
      if (lb) t += __ldg(lb + c);
    }
    out[(long long)b * ldo + c] = from_f<TO>(t);
  }
}

// ---------------------------------------------------------------------------------------
// NCHW fp32 <-> token-major bridges (32x32 smem transpose per image)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void tokens_from_nchw_kernel(const float *__restrict__ x, T *__restrict__ act, int C, int T_, int R) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T_) ? x[((long long)b * C + c) * T_ + t] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T_ && c < C) act[((long long)b * (R + T_) + R + t) * C + c] = from_f<T>(tile[threadIdx.x][i]);
  }
}
template <typename T>
__global__ void tokens_to_nchw_kernel(const T *__restrict__ act, const T *__restrict__ act_lo, float *__restrict__ x,
                                      int C, int T_, int R) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    const long long at = ((long long)b * (R + T_) + R + t) * C + c;
    tile[i][threadIdx.x] = (t < T_ && c < C) ? to_f(act[at]) + (act_lo ? to_f(act_lo[at]) : 0.0f) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T_) x[((long long)b * C + c) * T_ + t] = tile[threadIdx.x][i];
  }
}
template <typename T, bool TO_ACT>
__global__ void registers_copy_kernel(T *act, const T *act_lo, float *reg, int B, int S, int R, int C) {
  const long long n = (long long)B * R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int r = (i / C) % R;
    const long long b = i / ((long long)C * R);
    T *a = act + (b * S + r) * C + c;
    if (TO_ACT) *a = from_f<T>(reg[i]);
    else reg[i] = to_f(*a) + (act_lo ? to_f(act_lo[(b * S + r) * C + c]) : 0.0f);
  }
}

// ---------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void im2col_kernel(const TI *__restrict__ x, TO *__restrict__ A, long long ldA, int B, int H, int W,
                              int p, int Gh, int Gw) {
  const int Kc = 3 * p * p;
  const long long n = (long long)B * Gh * Gw * ldA;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = i % ldA;
    const long long row = i / ldA;
    float v = 0.0f;
    if (col < Kc) {
      const int dx = col % p, dy = (col / p) % p, c = col / (p * p);
      const int j = row % Gw, ii = (row / Gw) % Gh;
      const long long b = row / ((long long)Gw * Gh);
      v = to_f(x[((b * 3 + c) * H + (ii * p + dy)) * (long long)W + (j * p + dx)]);
    }
    A[i] = from_f<TO>(v);
  }
}

// Even patch sizes: one thread per PAIR of neighbouring input pixels.  A warp reads 32 consecutive pairs of an image
// row (coalesced) and writes them as p-element runs into the rows of the patches they belong to (a pair never
// straddles two patches); the thread that owns a patch's first pair also zeroes the row's padding columns.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
im2col_pairs_kernel(const TI *__restrict__ x, TO *__restrict__ A, int ldA, int B, int H, int W, int p, int Gh, int Gw) {
  const int Kc = 3 * p * p;
  const int wp = W >> 1;                                     // pairs per image row
  const long long n = (long long)B * 3 * H * wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % wp);
    const long long r = i / wp;                              // (b * 3 + c) * H + y
    const int y = (int)(r % H);
    const int bc = (int)(r / H);
    const int c = bc % 3, b = bc / 3;
    const int xx = 2 * w, j = xx / p, dx = xx - j * p, ii = y / p, dy = y - ii * p;
    if (ii >= Gh || j >= Gw) continue;                       // pixels behind the last whole patch
    float v0, v1;
    if constexpr (sizeof(TI) == 4) {
      const float2 t = *reinterpret_cast<const float2 *>(x + r * W + xx);
      v0 = t.x; v1 = t.y;
    } else {
      const float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(x + r * W + xx));
      v0 = t.x; v1 = t.y;
    }
    TO *row = A + ((long long)(b * Gh + ii) * Gw + j) * ldA;
    TO *dst = row + c * p * p + dy * p + dx;
    if constexpr (sizeof(TO) == 4) *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
    else *reinterpret_cast<uint32_t *>(dst) = pack_bf16x2(v0, v1);
    if (c == 0 && dy == 0 && dx == 0)
      for (int k = Kc; k < ldA; ++k) row[k] = from_f<TO>(0.0f);
  }
}

// One CTA per row of patches (b, ii): the 3 x p image rows of that patch row arrive in shared memory with coalesced
// loads (converted to the output type), then every patch's im2col row [3 p p | zero padding up to ldA] leaves as one
// contiguous run of pairs.  The pair kernel above writes each patch row in p-element pieces (28 bytes at p = 14:
// partial, misaligned sectors); this one moves the same bytes in full lines.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
im2col_rows_kernel(const TI *__restrict__ x, TO *__restrict__ A, int ldA, int H, int W, int p, int Gh, int Gw) {
  extern __shared__ __align__(16) uint8_t i2c_raw[];
  TO *tile = reinterpret_cast<TO *>(i2c_raw);                // [3][p][Wc], Wc = Gw * p
  const int b = blockIdx.x / Gh, ii = blockIdx.x - b * Gh;
  const int Wc = Gw * p, wp = Wc >> 1, rows = 3 * p;
#pragma unroll 4
  for (int i = threadIdx.x; i < rows * wp; i += blockDim.x) {
    const int r = i / wp, w2 = (i - r * wp) * 2;             // r = c * p + dy
    const int c = r / p, dy = r - c * p;
    const TI *src = x + ((long long)(b * 3 + c) * H + ii * p + dy) * W + w2;
    float v0, v1;
    if constexpr (sizeof(TI) == 4) {
      const float2 t = *reinterpret_cast<const float2 *>(src);
      v0 = t.x; v1 = t.y;
    } else {
      const float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(src));
      v0 = t.x; v1 = t.y;
    }
    if constexpr (sizeof(TO) == 4) *reinterpret_cast<float2 *>(tile + r * Wc + w2) = make_float2(v0, v1);
    else *reinterpret_cast<uint32_t *>(tile + r * Wc + w2) = pack_bf16x2(v0, v1);
  }
  __syncthreads();
  const int Kc = 3 * p * p, lp = ldA >> 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Gw; j += blockDim.x >> 5) {         // one warp per patch: consecutive lanes, consecutive pairs
    TO *row = A + ((long long)(b * Gh + ii) * Gw + j) * ldA;
#pragma unroll 4
    for (int k2 = lane; k2 < lp; k2 += 32) {
      const int k = 2 * k2;
      if (k < Kc) {
        const int r = k / p, dx = k - r * p;                 // r = c * p + dy (p even: a pair never leaves its tap row)
        const TO *src = tile + r * Wc + j * p + dx;
        if constexpr (sizeof(TO) == 4) *reinterpret_cast<float2 *>(row + k) = *reinterpret_cast<const float2 *>(src);
        else *reinterpret_cast<uint32_t *>(row + k) = *reinterpret_cast<const uint32_t *>(src);
      } else {
        if constexpr (sizeof(TO) == 4) *reinterpret_cast<float2 *>(row + k) = make_float2(0.0f, 0.0f);
        else *reinterpret_cast<uint32_t *>(row + k) = 0u;
      }
    }
  }
}

// one warp per row: log-sum-exp cross-entropy, smoothed-target BCE-with-logits, first-max argmax
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float *__restrict__ logits, long long ldl, const long long *__restrict__ labels, int B, int K,
                    float ls, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float *x = logits + (long long)row * ldl;
  const long long label64 = labels[row];
  // nn.CrossEntropyLoss (model_test.py:66,80) ignores its ignore_index (-100) and raises on any other label outside
  // [0, K): the former rows contribute nothing, the latter are counted in acc[4] for the host side to raise on
  if (label64 < 0 || label64 >= K) {
    if (lane == 0 && label64 != -100) atomicAdd(acc + 4, 1.0);
    return;
  }
  const int label = (int)label64;
  float mx = -INFINITY;
  int arg = K;
  for (int c = lane; c < K; c += 32) {
    const float v = x[c];
    if (v > mx) { mx = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  const float t_on = 1.0f - ls + ls / (float)K, t_off = ls / (float)K;
  float se = 0.0f, bce = 0.0f;
  for (int c = lane; c < K; c += 32) {
    const float v = x[c];
    se += expf(v - mx);
    const float t = c == label ? t_on : t_off;
    bce += fmaxf(v, 0.0f) - v * t + log1pf(expf(-fabsf(v)));
  }
  se = warp_sum(se);
  bce = warp_sum(bce);
  if (lane == 0) {
    const float ce = logf(se) + mx - x[label];
    atomicAdd(acc + 0, (double)ce);
    atomicAdd(acc + 1, (double)bce);
    atomicAdd(acc + 2, arg == label ? 1.0 : 0.0);
    atomicAdd(acc + 3, 1.0);
  }
}

template <typename T, bool EXACT>
__global__ void embed_tokens_kernel(T *__restrict__ act, const float *__restrict__ pos, int B, int T_, int R, int C,
                                    int act_id) {
  const long long n = (long long)B * T_ * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int t = (i / C) % T_;
    const long long b = i / ((long long)C * T_);
    T *a = act + ((b * (R + T_) + R + t) * C + c);
    *a = from_f<T>(apply_act<EXACT>(to_f(*a) + __ldg(pos + (long long)t * C + c), act_id));
  }
}

template <typename T, bool EXACT>
__global__ void activation_kernel(const T *__restrict__ x, T *__restrict__ y, long long n, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(apply_act<EXACT>(to_f(x[i]), act));
}

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

}  // namespace sdp

using namespace sdp;

extern "C" int sdp_layernorm_rows(const void *x, int64_t ldx, const float *w, const float *b, void *out,
                                  int64_t ldo, int M, int C, float eps, int dtype, void *stream) {
  SDP_CHECK(x && out && M > 0 && C > 0, "sdp_layernorm_rows: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SDP_BF16) return launch_ln_rows<bf16>(x, ldx, w, b, out, ldo, M, C, eps, st);
  SDP_CHECK(dtype == SDP_F32, "sdp_layernorm_rows: unknown dtype %d", dtype);
  return launch_ln_rows<float>(x, ldx, w, b, out, ldo, M, C, eps, st);
}

extern "C" int sdp_row_stats(const void *x, int64_t ldx, float *stats, int parts, int M, int C, int dtype,
                             void *stream) {
  SDP_CHECK(x && stats && parts > 0 && M > 0 && C > 0, "sdp_row_stats: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((M + 7) / 8), block(256);
  if (dtype == SDP_BF16) row_stats_kernel<bf16><<<grid, block, 0, st>>>((const bf16 *)x, ldx, stats, parts, M, C);
  else row_stats_kernel<float><<<grid, block, 0, st>>>((const float *)x, ldx, stats, parts, M, C);
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_fill_registers(void *act, void *act_lo, int dtype, const float *table, int B, int S, int R, int C,
                                  void *stream) {
  SDP_CHECK(act && table && B > 0 && R > 0 && R <= S && C > 0, "sdp_fill_registers: bad arguments");
  SDP_CHECK(act_lo == nullptr || dtype == SDP_BF16, "sdp_fill_registers: a lo plane needs a bf16 stream");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int g = grid_for((long long)B * R * C);
  if (dtype == SDP_BF16) fill_registers_kernel<bf16><<<g, 256, 0, st>>>((bf16 *)act, (bf16 *)act_lo, table, B, S, R, C);
  else fill_registers_kernel<float><<<g, 256, 0, st>>>((float *)act, nullptr, table, B, S, R, C);
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_pool_ln(const void *act, const void *act_lo, int dtype, int B, int S, int C, int row0, int nrows,
                           const float *ln_w, const float *ln_b, float eps, void *out, int out_dtype,
                           int64_t ldo, void *stream) {
  SDP_CHECK(act && out && B > 0 && nrows > 0 && row0 >= 0 && row0 + nrows <= S, "sdp_pool_ln: bad arguments");
  SDP_CHECK(act_lo == nullptr || dtype == SDP_BF16, "sdp_pool_ln: a lo plane needs a bf16 stream");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t sm = (C + 32) * sizeof(float);
  SDP_CHECK(sm <= 48 * 1024, "sdp_pool_ln: C=%d too large", C);
#define POOL(TI, TO) \
  pool_ln_kernel<TI, TO><<<B, 256, sm, st>>>((const TI *)act, (const TI *)act_lo, S, C, row0, nrows, ln_w, ln_b, eps, (TO *)out, ldo)
  if (dtype == SDP_BF16 && out_dtype == SDP_BF16) POOL(bf16, bf16);
  else if (dtype == SDP_BF16) POOL(bf16, float);
  else if (out_dtype == SDP_BF16) POOL(float, bf16);
  else POOL(float, float);
#undef POOL
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_tokens_from_nchw(const float *x, const float *reg, void *act, int dtype, int B, int C, int T,
                                    int R, void *stream) {
  SDP_CHECK(x && act && B > 0 && C > 0 && T > 0 && R >= 0, "sdp_tokens_from_nchw: bad arguments");
  SDP_CHECK(B <= 65535, "sdp_tokens_from_nchw: B too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  if (dtype == SDP_BF16) tokens_from_nchw_kernel<bf16><<<grid, block, 0, st>>>(x, (bf16 *)act, C, T, R);
  else tokens_from_nchw_kernel<float><<<grid, block, 0, st>>>(x, (float *)act, C, T, R);
  SDP_LAUNCH_OK();
  if (R > 0 && reg) {
    const int g = grid_for((long long)B * R * C);
    if (dtype == SDP_BF16)
      registers_copy_kernel<bf16, true><<<g, 256, 0, st>>>((bf16 *)act, nullptr, const_cast<float *>(reg), B, R + T, R, C);
    else
      registers_copy_kernel<float, true><<<g, 256, 0, st>>>((float *)act, nullptr, const_cast<float *>(reg), B, R + T, R, C);
    SDP_LAUNCH_OK();
  }
  return 0;
}

extern "C" int sdp_tokens_to_nchw(const void *act, const void *act_lo, int dtype, float *x, float *reg, int B, int C,
                                  int T, int R, void *stream) {
  SDP_CHECK(act && B > 0 && C > 0 && T > 0 && R >= 0, "sdp_tokens_to_nchw: bad arguments");
  SDP_CHECK(act_lo == nullptr || dtype == SDP_BF16, "sdp_tokens_to_nchw: a lo plane needs a bf16 stream");
  SDP_CHECK(B <= 65535, "sdp_tokens_to_nchw: B too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x) {
    dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
    if (dtype == SDP_BF16) tokens_to_nchw_kernel<bf16><<<grid, block, 0, st>>>((const bf16 *)act, (const bf16 *)act_lo, x, C, T, R);
    else tokens_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float *)act, nullptr, x, C, T, R);
    SDP_LAUNCH_OK();
  }
  if (R > 0 && reg) {
    const int g = grid_for((long long)B * R * C);
    if (dtype == SDP_BF16)
      registers_copy_kernel<bf16, false><<<g, 256, 0, st>>>((bf16 *)const_cast<void *>(act), (const bf16 *)act_lo, reg, B, R + T, R, C);
    else
      registers_copy_kernel<float, false><<<g, 256, 0, st>>>((float *)const_cast<void *>(act), nullptr, reg, B, R + T, R, C);
    SDP_LAUNCH_OK();
  }
  return 0;
}

extern "C" int sdp_im2col_patches(const void *x, int x_dtype, void *A, int a_dtype, int64_t ldA, int B, int H,
                                  int W, int p, void *stream) {
  SDP_CHECK(x && A && B > 0 && p > 0, "sdp_im2col_patches: bad arguments");
  SDP_CHECK(H % p == 0 && W % p == 0, "sdp_im2col_patches: image %dx%d not divisible by patch %d", H, W, p);
  SDP_CHECK(ldA >= 3 * p * p, "sdp_im2col_patches: ldA=%lld < 3*p*p", (long long)ldA);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Gh = H / p, Gw = W / p;
  const int g = grid_for((long long)B * Gh * Gw * ldA);
  // pair kernel: even patch size and width, pair-aligned buffers and row pitch, 32-bit sized problem
  const size_t xs = dtype_size(x_dtype), as = dtype_size(a_dtype);
  const bool pairs = p % 2 == 0 && W % 2 == 0 && ldA % 2 == 0 && ldA < (1 << 24) && (long long)B * Gh * Gw < (1ll << 31) &&
                     (reinterpret_cast<uintptr_t>(x) % (2 * xs)) == 0 && (reinterpret_cast<uintptr_t>(A) % (2 * as)) == 0;
  // whole patch rows staged in shared memory when they fit (every reference geometry does: 3 * 14 * 224 values)
  const size_t tile_bytes = (size_t)3 * p * (Gw * p) * as;
  if (pairs && tile_bytes <= 48 * 1024 && (long long)B * Gh < (1ll << 31)) {
#define I2R(TI, TO) im2col_rows_kernel<TI, TO><<<B * Gh, 256, tile_bytes, st>>>((const TI *)x, (TO *)A, (int)ldA, H, W, p, Gh, Gw)
    if (x_dtype == SDP_F32 && a_dtype == SDP_BF16) I2R(float, bf16);
    else if (x_dtype == SDP_F32) I2R(float, float);
    else if (a_dtype == SDP_BF16) I2R(bf16, bf16);
    else I2R(bf16, float);
#undef I2R
    SDP_LAUNCH_OK();
    return 0;
  }
  if (pairs) {
    const int gp = grid_for((long long)B * 3 * H * (W / 2));
#define I2P(TI, TO) im2col_pairs_kernel<TI, TO><<<gp, 256, 0, st>>>((const TI *)x, (TO *)A, (int)ldA, B, H, W, p, Gh, Gw)
    if (x_dtype == SDP_F32 && a_dtype == SDP_BF16) I2P(float, bf16);
    else if (x_dtype == SDP_F32) I2P(float, float);
    else if (a_dtype == SDP_BF16) I2P(bf16, bf16);
    else I2P(bf16, float);
#undef I2P
    SDP_LAUNCH_OK();
    return 0;
  }
#define I2C(TI, TO) im2col_kernel<TI, TO><<<g, 256, 0, st>>>((const TI *)x, (TO *)A, ldA, B, H, W, p, Gh, Gw)
  if (x_dtype == SDP_F32 && a_dtype == SDP_BF16) I2C(float, bf16);
  else if (x_dtype == SDP_F32) I2C(float, float);
  else if (a_dtype == SDP_BF16) I2C(bf16, bf16);
  else I2C(bf16, float);
#undef I2C
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_eval_metrics(const float *logits, int64_t ldl, const int64_t *labels, int B, int K,
                                float label_smoothing, double *acc, void *stream) {
  SDP_CHECK(logits && labels && acc && B > 0 && K > 0 && ldl >= K, "sdp_eval_metrics: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  eval_metrics_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, ldl, reinterpret_cast<const long long *>(labels), B, K,
                                                   label_smoothing, acc);
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_embed_tokens(void *act, int dtype, const float *pos, int B, int T, int R, int C, int act_id,
                                void *stream) {
  SDP_CHECK(act && pos && B > 0 && T > 0 && R >= 0 && C > 0, "sdp_embed_tokens: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int g = grid_for((long long)B * T * C);
  if (dtype == SDP_BF16) embed_tokens_kernel<bf16, false><<<g, 256, 0, st>>>((bf16 *)act, pos, B, T, R, C, act_id);
  else embed_tokens_kernel<float, true><<<g, 256, 0, st>>>((float *)act, pos, B, T, R, C, act_id);
  SDP_LAUNCH_OK();
  return 0;
}

extern "C" int sdp_activation(const void *x, void *y, int64_t n, int act, int dtype, void *stream) {
  SDP_CHECK(x && y && n > 0, "sdp_activation: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool force_fast = (act & 0x100) != 0;   // test hook: fast forms on fp32 I/O
  act &= 0xff;
  const int g = grid_for(n);
  if (dtype == SDP_BF16) activation_kernel<bf16, false><<<g, 256, 0, st>>>((const bf16 *)x, (bf16 *)y, n, act);
  else if (force_fast) activation_kernel<float, false><<<g, 256, 0, st>>>((const float *)x, (float *)y, n, act);
  else activation_kernel<float, true><<<g, 256, 0, st>>>((const float *)x, (float *)y, n, act);
  SDP_LAUNCH_OK();
  return 0;
}
