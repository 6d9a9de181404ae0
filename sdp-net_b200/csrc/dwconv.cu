// Mixer front half: channel LayerNorm (eps 1e-6) fused with the depthwise k x k 'same' conv.
// Reference: layers.py:102 `conv2d[0](layer_norm_1(x))` = layers.py:12-24 + :73-78.
//
// One CTA per image.  Phase 1: per-token mean / rstd over C (one warp per token, coalesced
// 128-bit row reads) into shared memory.  Phase 2: for each slab of CH channels, stage the
// normalised G x G plane with its zero halo in shared memory (the halo is literal 0 -- padding is
// applied after the norm), then every thread owns one channel and slides an
// XB-wide output window along image rows (k*(XB+k-1) shared loads per k*k*XB FMAs).
// HBM traffic: the image's [T, C] slab is read once from HBM (phase 1), re-read from L2 (phase 2)
// and written once: algorithmic bytes = 2 * T * C * sizeof(T) per image.
#include <cstdlib>

#include "common.cuh"

namespace sdp {

constexpr int DW_THREADS = 256;
constexpr int DW_XB = 8;

template <typename T>
__device__ __forceinline__ void token_stats(const T *__restrict__ row, int C, int lane, float eps, float &mean,
                                            float &rstd) {
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += to_f(row[c]);
  mean = warp_sum(s) / (float)C;
  float q = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f(row[c]) - mean;
    q = fmaf(d, d, q);
  }
  rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
}
template <>
__device__ __forceinline__ void token_stats<bf16>(const bf16 *__restrict__ row, int C, int lane, float eps,
                                                  float &mean, float &rstd) {
  float s = 0.0f, q = 0.0f;
  if ((C & 7) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    const uint4 *rv = reinterpret_cast<const uint4 *>(row);
    const int nv = C >> 3;
    for (int i = lane; i < nv; i += 32) {
      const uint4 u = rv[i];
      float2 f;
      f = unpack_bf16x2(u.x); s += f.x + f.y;
      f = unpack_bf16x2(u.y); s += f.x + f.y;
      f = unpack_bf16x2(u.z); s += f.x + f.y;
      f = unpack_bf16x2(u.w); s += f.x + f.y;
    }
    mean = warp_sum(s) / (float)C;
    for (int i = lane; i < nv; i += 32) {
      const uint4 u = rv[i];
      float2 f;
      float d;
      f = unpack_bf16x2(u.x); d = f.x - mean; q = fmaf(d, d, q); d = f.y - mean; q = fmaf(d, d, q);
      f = unpack_bf16x2(u.y); d = f.x - mean; q = fmaf(d, d, q); d = f.y - mean; q = fmaf(d, d, q);
      f = unpack_bf16x2(u.z); d = f.x - mean; q = fmaf(d, d, q); d = f.y - mean; q = fmaf(d, d, q);
      f = unpack_bf16x2(u.w); d = f.x - mean; q = fmaf(d, d, q); d = f.y - mean; q = fmaf(d, d, q);
    }
  } else {
    for (int c = lane; c < C; c += 32) s += to_f(row[c]);
    mean = warp_sum(s) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float d = to_f(row[c]) - mean;
      q = fmaf(d, d, q);
    }
  }
  rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
}

// KS > 0: compile-time kernel size with the sliding window; KS == 0: runtime k, one output at a time.
template <typename T, int KS, int CH>
__global__ void __launch_bounds__(DW_THREADS)
ln_dwconv_kernel(const T *__restrict__ act, const float *__restrict__ gamma, const float *__restrict__ beta,
                 const float *__restrict__ wdw, const float *__restrict__ bdw, T *__restrict__ out, int Gh, int Gw,
                 int C, int krt, int R, float eps, int PW) {
  extern __shared__ float smem[];
  const int k = KS > 0 ? KS : krt;
  const int lo = (k - 1) / 2;                       // 'same': left/top pad (k-1)/2, rest on the right/bottom
  const int Tn = Gh * Gw, S = R + Tn;
  const int PH = Gh + k - 1;
  float *s_mean = smem;
  float *s_rstd = smem + Tn;
  float *tile = smem + 2 * Tn;                      // [PH][PW][CH]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T *xin = act + (long long)b * S * C;
  T *xout = out + (long long)b * S * C;

  // register rows of the output: zeros (the following GEMM passes registers through untouched)
  for (int i = tid; i < R * C; i += DW_THREADS) xout[i] = from_f<T>(0.0f);

  // ---- phase 1: token statistics ----
  for (int t = warp; t < Tn; t += DW_THREADS / 32) {
    float mean, rstd;
    token_stats<T>(xin + (long long)(R + t) * C, C, lane, eps, mean, rstd);
    if (lane == 0) { s_mean[t] = mean; s_rstd[t] = rstd; }
  }
  __syncthreads();

  // ---- phase 2: channel slabs ----
  constexpr int GROUPS = DW_THREADS / CH;
  const int cl = tid % CH, grp = tid / CH;
  for (int c0 = 0; c0 < C; c0 += CH) {
    const int c = c0 + cl;
    const bool cok = c < C;
    const float g = cok ? __ldg(gamma + c) : 0.0f;
    const float be = cok ? __ldg(beta + c) : 0.0f;
    // stage the normalised plane with zero halo
    for (int pix = grp; pix < PH * PW; pix += GROUPS) {
      const int py = pix / PW, px = pix % PW;
      const int iy = py - lo, ix = px - lo;
      float v = 0.0f;
      if (cok && iy >= 0 && iy < Gh && ix >= 0 && ix < Gw) {
        const int t = iy * Gw + ix;
        v = (to_f(xin[(long long)(R + t) * C + c]) - s_mean[t]) * s_rstd[t] * g + be;
      }
      tile[pix * CH + cl] = v;
    }
    __syncthreads();
    const float bias = (cok && bdw) ? __ldg(bdw + c) : 0.0f;
    if (KS > 0) {
      float w[KS > 0 ? KS * KS : 1];
#pragma unroll
      for (int i = 0; i < KS * KS; ++i) w[i] = cok ? __ldg(wdw + (long long)i * C + c) : 0.0f;
      const int xchunks = (Gw + DW_XB - 1) / DW_XB;
      for (int item = grp; item < Gh * xchunks; item += GROUPS) {
        const int y = item / xchunks, x0 = (item % xchunks) * DW_XB;
        float acc[DW_XB];
#pragma unroll
        for (int i = 0; i < DW_XB; ++i) acc[i] = bias;
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          const float *trow = tile + ((y + dy) * PW + x0) * CH + cl;
          float win[DW_XB + KS - 1];
#pragma unroll
          for (int xx = 0; xx < DW_XB + KS - 1; ++xx) win[xx] = trow[xx * CH];
#pragma unroll
          for (int dx = 0; dx < KS; ++dx)
#pragma unroll
            for (int i = 0; i < DW_XB; ++i) acc[i] = fmaf(win[i + dx], w[dy * KS + dx], acc[i]);
        }
        if (cok) {
#pragma unroll
          for (int i = 0; i < DW_XB; ++i)
            if (x0 + i < Gw) xout[(long long)(R + y * Gw + x0 + i) * C + c] = from_f<T>(acc[i]);
        }
      }
    } else {
      for (int t = grp; t < Tn; t += GROUPS) {
        const int y = t / Gw, x = t % Gw;
        float acc = bias;
        for (int dy = 0; dy < k; ++dy)
          for (int dx = 0; dx < k; ++dx)
            acc = fmaf(tile[((y + dy) * PW + x + dx) * CH + cl],
                       cok ? __ldg(wdw + (long long)(dy * k + dx) * C + c) : 0.0f, acc);
        if (cok) xout[(long long)(R + t) * C + c] = from_f<T>(acc);
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------
// bf16 fast path.  Same structure, but a slab is 64 channels held as packed bf16x2 words:
// lane = channel pair, so every global access is one 128-byte line per warp (token-major rows),
// every shared access is a conflict-free LDS.32/STS.32, and each loaded word feeds two FMA
// chains.  Depthwise taps are tap-major fp32 ([k*k, C]) so a warp reads them as one coalesced
// LDG.64 per tap.  FMA-bound by design: per output pair 2*k*k FMAs against k*(XB+k-1)/XB LDS.
// ---------------------------------------------------------------------------------------
constexpr int DWF_THREADS = 512;   // fast path: one 16-warp CTA per SM

__device__ __forceinline__ void dw_cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Shared-memory carve-up of the fast kernel (floats); shared with the launcher.
struct DwLayout {
  int stats, off, w, raw, tile, total;
  __host__ __device__ DwLayout(int Tn, int PH, int PW, int KS) {
    stats = 0;                                 // mean[Tn], rstd[Tn]
    off = 2 * Tn;                              // int[Tn]
    w = (3 * Tn + 3) & ~3;                     // 2 x [KS*KS][32] float2 (16-byte aligned)
    raw = w + 2 * KS * KS * 64;                // [Tn][32] raw bf16x2 words of the NEXT slab
    tile = (raw + Tn * 32 + 1) & ~1;           // [PH][PW][32] normalised float2 (channel pair), zero halo
    total = tile + PH * PW * 64;
  }
};

template <int KS>
__global__ void __launch_bounds__(DWF_THREADS, 1)
ln_dwconv_bf16_kernel(const bf16 *__restrict__ act, const float *__restrict__ stats, int parts,
                      const float *__restrict__ gamma, const float *__restrict__ beta,
                      const float *__restrict__ wdw, const float *__restrict__ bdw, bf16 *__restrict__ out, int Gh,
                      int Gw, int C, int R, float eps, int PW) {
  extern __shared__ __align__(16) float smem[];
  constexpr int lo = (KS - 1) / 2;
  constexpr int NW = DWF_THREADS / 32;
  const int Tn = Gh * Gw, S = R + Tn;
  const int PH = Gh + KS - 1;
  const DwLayout L(Tn, PH, PW, KS);
  float *s_mean = smem + L.stats;
  float *s_rstd = s_mean + Tn;
  int *s_off = reinterpret_cast<int *>(smem + L.off);
  float2 *s_w = reinterpret_cast<float2 *>(smem + L.w);
  uint32_t *s_raw = reinterpret_cast<uint32_t *>(smem + L.raw);
  float2 *tile = reinterpret_cast<float2 *>(smem + L.tile);
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bf16 *xin = act + ((long long)b * S + R) * C;       // first patch row of this image
  bf16 *xout = out + (long long)b * S * C;
  const uint32_t raw_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_raw));
  const uint32_t w_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_w));

  // async prefetch of one slab: raw token rows (128 B each) and the slab's taps (256 B per tap)
  auto prefetch = [&](int c0, int buf) {
    const int chunks = min(8, (C - c0) / 8);               // 16-byte chunks of live channels per token
    for (int i = tid; i < Tn * 8; i += DWF_THREADS) {
      const int t = i >> 3, ch = i & 7;
      if (ch < chunks) dw_cp_async16(raw_addr + (uint32_t)(t * 128 + ch * 16), xin + (long long)t * C + c0 + ch * 8);
    }
    const int wchunks = min(16, (C - c0) / 4);
    for (int i = tid; i < KS * KS * 16; i += DWF_THREADS) {
      const int tap = i >> 4, ch = i & 15;
      if (ch < wchunks)
        dw_cp_async16(w_addr + (uint32_t)((buf * KS * KS + tap) * 256 + ch * 16), wdw + (long long)tap * C + c0 + ch * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // gridDim.y CTAs share one image: each takes a contiguous range of channel slabs (and recomputes the cheap
  // token statistics), which halves the bytes in flight per SM -- the L2 holds every image being worked on
  const int nslab = (C + 63) / 64;
  const int slab_lo = (int)(((long long)nslab * blockIdx.y) / gridDim.y), slab_hi = (int)(((long long)nslab * (blockIdx.y + 1)) / gridDim.y);
  const int c_lo = slab_lo * 64, c_hi = min(C, slab_hi * 64);
  prefetch(c_lo, 0);

  if (blockIdx.y == 0)
    for (int i = tid; i < R * C / 2; i += DWF_THREADS) reinterpret_cast<uint32_t *>(xout)[i] = 0u;
  // zero the whole tile once: the halo cells are never written again
  for (int i = tid; i < PH * PW * 32; i += DWF_THREADS) tile[i] = make_float2(0.f, 0.f);
  for (int t = tid; t < Tn; t += DWF_THREADS) s_off[t] = ((t / Gw + lo) * PW + (t % Gw + lo)) * 32;

  // ---- phase 1: token statistics.  Preferred: the (sum, sumsq) column parts the producer GEMM emitted
  //      for exactly these bf16 values; otherwise one pass over the rows with the first element as
  //      shift (no cancellation: the shift is within a few sigma of the mean), 4 tokens per warp ----
  if (stats != nullptr) {
    const float *sp = stats + ((long long)b * S + R) * parts * 2;
    for (int t = tid; t < Tn; t += DWF_THREADS) {
      float s1 = 0.0f, s2 = 0.0f;
      for (int p = 0; p < parts; ++p) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(sp + ((long long)t * parts + p) * 2));
        s1 += v.x;
        s2 += v.y;
      }
      const float mean = s1 / (float)C;
      s_mean[t] = mean;
      s_rstd[t] = 1.0f / sqrtf(fmaxf(s2 / (float)C - mean * mean, 0.0f) + eps);
    }
  } else {
    const int nv = C >> 3;
    for (int t0 = warp * 4; t0 < Tn; t0 += NW * 4) {
      float s1[4], s2[4], sh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = min(t0 + u, Tn - 1);
        const uint4 *rv = reinterpret_cast<const uint4 *>(xin + (long long)t * C);
        sh[u] = __bfloat162float(xin[(long long)t * C]);
        s1[u] = s2[u] = 0.0f;
        for (int i = lane; i < nv; i += 32) {
          const uint4 v = __ldg(rv + i);
          const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = __uint_as_float(wd[j] << 16) - sh[u], c2 = __uint_as_float(wd[j] & 0xffff0000u) - sh[u];
            s1[u] += a + c2;
            s2[u] = fmaf(a, a, fmaf(c2, c2, s2[u]));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float a = warp_sum(s1[u]), q = warp_sum(s2[u]);
        const float dm = a / (float)C;
        const float var = fmaxf(q / (float)C - dm * dm, 0.0f);
        if (lane == 0 && t0 + u < Tn) {
          s_mean[t0 + u] = sh[u] + dm;
          s_rstd[t0 + u] = 1.0f / sqrtf(var + eps);
        }
      }
    }
  }

  const int xchunks = (Gw + DW_XB - 1) / DW_XB;
  int buf = 0;
  for (int c0 = c_lo; c0 < c_hi; c0 += 64, buf ^= 1) {
    const int c = c0 + 2 * lane;
    const bool cok = c < C;                       // C % 8 == 0 on this path
    const float2 g = cok ? __ldg(reinterpret_cast<const float2 *>(gamma + c)) : make_float2(0.f, 0.f);
    const float2 be = cok ? __ldg(reinterpret_cast<const float2 *>(beta + c)) : make_float2(0.f, 0.f);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                               // this slab's raw rows + taps (and the stats) are visible
    // ---- transform: raw -> normalised tile ----
    for (int t = warp; t < Tn; t += NW) {
      const uint32_t u = s_raw[t * 32 + lane];
      const float m = s_mean[t], r = s_rstd[t];
      const float lo_v = __uint_as_float(u << 16), hi_v = __uint_as_float(u & 0xffff0000u);
      // kept in fp32: no unpack in the FMA loop (the normalised value is not re-rounded either)
      tile[s_off[t] + lane] = cok ? make_float2((lo_v - m) * r * g.x + be.x, (hi_v - m) * r * g.y + be.y)
                                  : make_float2(0.f, 0.f);
    }
    __syncthreads();                               // tile ready; s_raw free
    if (c0 + 64 < c_hi) prefetch(c0 + 64, buf ^ 1);   // next slab streams in under this slab's FMAs
    const float2 bias = (cok && bdw) ? __ldg(reinterpret_cast<const float2 *>(bdw + c)) : make_float2(0.f, 0.f);
    const float2 *wslab = s_w + buf * KS * KS * 32 + lane;
    for (int item = warp; item < Gh * xchunks; item += NW) {
      const int y = item / xchunks, x0 = (item - y * xchunks) * DW_XB;
      float2 acc[DW_XB];
#pragma unroll
      for (int i = 0; i < DW_XB; ++i) acc[i] = bias;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        float2 w[KS];
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) w[dx] = cok ? wslab[(dy * KS + dx) * 32] : make_float2(0.f, 0.f);
        const float2 *trow = tile + ((y + dy) * PW + x0) * 32 + lane;
        float2 win[DW_XB + KS - 1];
#pragma unroll
        for (int xx = 0; xx < DW_XB + KS - 1; ++xx) win[xx] = trow[xx * 32];
#pragma unroll
        for (int dx = 0; dx < KS; ++dx)
#pragma unroll
          for (int i = 0; i < DW_XB; ++i) {
            acc[i].x = fmaf(win[i + dx].x, w[dx].x, acc[i].x);
            acc[i].y = fmaf(win[i + dx].y, w[dx].y, acc[i].y);
          }
      }
      if (cok) {
        bf16 *op = xout + (long long)(R + y * Gw + x0) * C + c;
#pragma unroll
        for (int i = 0; i < DW_XB; ++i)
          if (x0 + i < Gw) *reinterpret_cast<uint32_t *>(op + (long long)i * C) = pack_bf16x2(acc[i].x, acc[i].y);
      }
    }
    // the next iteration's first __syncthreads orders this slab's tile reads before its overwrite
  }
}

template <int KS>
static int launch_dw_bf16(const void *act, const float *stats, int parts, const float *gamma, const float *beta,
                          const float *wdw, const float *bdw, void *out, int B, int Gh, int Gw, int C, int R,
                          float eps, cudaStream_t st) {
  const int PW = ((Gw + DW_XB - 1) / DW_XB) * DW_XB + KS - 1;
  const int PH = Gh + KS - 1;
  const DwLayout L(Gh * Gw, PH, PW, KS);
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 220 * 1024) return -1;             // caller falls back to the generic kernel
  auto kern = ln_dwconv_bf16_kernel<KS>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
  // one CTA per image: splitting an image's channel slabs over several CTAs was measured slower on B200 (every extra
  // CTA pays the statistics pass again)
  kern<<<dim3(B, 1), DWF_THREADS, smem, st>>>((const bf16 *)act, stats, parts, gamma, beta, wdw, bdw, (bf16 *)out, Gh, Gw, C, R, eps,
                                     PW);
  SDP_LAUNCH_OK();
  return 0;
}


template <typename T, int KS>
static int launch_dw(const void *act, const float *gamma, const float *beta, const float *wdw, const float *bdw,
                     void *out, int B, int Gh, int Gw, int C, int k, int R, float eps, cudaStream_t st) {
  constexpr int CH = 32;
  const int PW = ((Gw + DW_XB - 1) / DW_XB) * DW_XB + k - 1;
  const int PH = Gh + k - 1;
  const size_t smem = (size_t)(2 * Gh * Gw + (size_t)PH * PW * CH) * sizeof(float);
  SDP_CHECK(smem <= 220 * 1024, "sdp_ln_dwconv: grid %dx%d with k=%d needs %zu B of shared memory", Gh, Gw, k,
            smem);
  auto kern = ln_dwconv_kernel<T, KS, CH>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
  kern<<<B, DW_THREADS, smem, st>>>((const T *)act, gamma, beta, wdw, bdw, (T *)out, Gh, Gw, C, k, R, eps, PW);
  SDP_LAUNCH_OK();
  return 0;
}

template <typename T>
static int dispatch_dw(const void *act, const float *gamma, const float *beta, const float *wdw, const float *bdw,
                       void *out, int B, int Gh, int Gw, int C, int k, int R, float eps, cudaStream_t st) {
  switch (k) {
    case 3: return launch_dw<T, 3>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
    case 5: return launch_dw<T, 5>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
    case 7: return launch_dw<T, 7>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
    default: return launch_dw<T, 0>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
  }
}

}  // namespace sdp

using namespace sdp;

// Every kernel here computes the token statistics itself when none are supplied, so nobody needs to run
// sdp_row_stats for it; kept in the ABI for callers that want to know whether supplied statistics are required.
extern "C" int sdp_ln_dwconv_wants_stats(int Gh, int Gw, int C, int k, int R, int dtype) {
  (void)Gh; (void)Gw; (void)C; (void)k; (void)R; (void)dtype;
  return 0;
}

extern "C" int sdp_ln_dwconv(const void *act, const float *gamma, const float *beta, const float *wdw,
                             const float *bdw, void *out, int B, int Gh, int Gw, int C, int k, int R, float eps,
                             int dtype, void *stream) {
  return sdp_ln_dwconv_stats(act, nullptr, 0, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, dtype, stream);
}

extern "C" int sdp_ln_dwconv_stats(const void *act, const float *stats, int parts, const float *gamma,
                                   const float *beta, const float *wdw, const float *bdw, void *out, int B, int Gh,
                                   int Gw, int C, int k, int R, float eps, int dtype, void *stream) {
  SDP_CHECK(act && gamma && beta && wdw && out, "sdp_ln_dwconv: null pointer");
  SDP_CHECK(stats == nullptr || (parts > 0 && dtype == SDP_BF16 && (k == 3 || k == 5 || k == 7) && C % 8 == 0),
            "sdp_ln_dwconv_stats: producer statistics are consumed by the bf16 k in {3,5,7} kernel only");
  SDP_CHECK(B > 0 && Gh > 0 && Gw > 0 && C > 0 && k > 0 && R >= 0, "sdp_ln_dwconv: bad sizes");
  SDP_CHECK(act != out, "sdp_ln_dwconv: must not run in place (spatial neighbours are read)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == SDP_BF16) {
    const bool fast = C % 8 == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (reinterpret_cast<uintptr_t>(wdw) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(gamma) & 7) == 0 && (reinterpret_cast<uintptr_t>(beta) & 7) == 0 &&
                      (bdw == nullptr || (reinterpret_cast<uintptr_t>(bdw) & 7) == 0);
    int rc = -1;
    if (fast && k == 7) rc = launch_dw_bf16<7>(act, stats, parts, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
    if (fast && k == 5) rc = launch_dw_bf16<5>(act, stats, parts, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
    if (fast && k == 3) rc = launch_dw_bf16<3>(act, stats, parts, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
    if (rc >= 0) return rc;
    SDP_CHECK(stats == nullptr, "sdp_ln_dwconv_stats: the fast kernel does not fit this shape; pass stats == NULL");
    return dispatch_dw<bf16>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
  }
  SDP_CHECK(dtype == SDP_F32, "sdp_ln_dwconv: unknown dtype %d", dtype);
  return dispatch_dw<float>(act, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, st);
}
