// Validation preprocessing in front of the forward, on the GPU, bit-exact with the reference's CPU pipeline:
//   RGB uint8 image of any size -> Pillow bicubic resize -> centre crop -> float32 * (1/255) -> (x - mean) / std
// Reference: val_transforms, hf_dataset_generator.py:27-41 (torchvision v2 transforms on PIL images; the resize is
// Pillow's ImagingResample, src/libImaging/Resample.c, 8-bit fixed-point path).  Only the crop window is computed.
//
// prep_coeffs_kernel, then ONE fused kernel (prep_fused_kernel: both passes with the uint8 intermediate in shared
// memory) when a band of output rows fits shared memory for every image of the batch; otherwise the general pair
// prep_horizontal_kernel / prep_vertical_kernel with the intermediate in global memory.  All byte/integer work:
//   prep_coeffs_kernel      per image, axis and output index of the crop window: tap range + 22-bit fixed-point Keys
//                           (a = -0.5) weights, evaluated in double precision with explicitly rounded operations (no
//                           FMA contraction) in Pillow's operation order, so the integers equal Pillow's;
//   prep_horizontal_kernel  source row (only the columns and rows the window needs) staged in shared memory with
//                           32-bit loads -> [rows, crop_w, 3] uint8 intermediate (Pillow rounds to uint8 between passes);
//   prep_vertical_kernel    vertical taps on the intermediate -> uint8 -> 3 x 256 normalisation table (built per CTA with
//                           the reference's float32 operation order) -> planar NCHW float32 / bf16, coalesced stores.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sdp {

constexpr int PREP_PRECISION_BITS = 32 - 8 - 2;     // Resample.c PRECISION_BITS
constexpr int PREP_THREADS = 256;

// ---- host + device: tap geometry of one axis (Resample.c precompute_coeffs) ----
struct PrepAxis {
  double scale, support, ss;
  int ksize;
};
static inline PrepAxis prep_axis_host(int in_size, int out_size) {
  PrepAxis a;
  a.scale = (double)in_size / out_size;
  const double fs = a.scale < 1.0 ? 1.0 : a.scale;
  a.support = 2.0 * fs;
  a.ss = 1.0 / fs;
  a.ksize = (int)ceil(a.support) * 2 + 1;
  return a;
}
static inline void prep_bounds_host(const PrepAxis &a, int in_size, int xx, int *xmin, int *cnt) {
  const double center = 0.0 + (xx + 0.5) * a.scale;
  int lo = (int)(center - a.support + 0.5);
  if (lo < 0) lo = 0;
  int hi = (int)(center + a.support + 0.5);
  if (hi > in_size) hi = in_size;
  *xmin = lo;
  *cnt = hi - lo;
}
static inline int crop_anchor(int resized, int crop) { return (int)nearbyint((resized - crop) / 2.0); }   // Python round(): half to even

__device__ __forceinline__ double prep_bicubic(double x) {      // Resample.c bicubic_filter, a = -0.5
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// coefficient records: per image (crop_h + crop_w) records of (kmax + 2) ints: first tap, tap count, weights.
// Records 0..crop_h-1 are the window's rows (vertical pass), the rest its columns (horizontal pass).
__global__ void __launch_bounds__(128)
prep_coeffs_kernel(const sdp_image_desc *__restrict__ img, int rh, int rw, int ch, int cw, int top, int left, int kmax,
                   int *__restrict__ coef, float *__restrict__ lut, float m0, float m1, float m2, float s0, float s1,
                   float s2) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  // normalisation table [3][256]: float32(v) * float32(1/255), minus mean, divided by std, each step rounded to
  // float32 (torchvision to_dtype_image + normalize_image); explicit _rn operations: nothing contracts into an FMA
  if (b == 0)
    for (int i = j; i < 768; i += gridDim.x * blockDim.x) {
      const int c = i >> 8;
      const float mean = c == 0 ? m0 : c == 1 ? m1 : m2, sd = c == 0 ? s0 : c == 1 ? s1 : s2;
      lut[i] = __fdiv_rn(__fsub_rn(__fmul_rn((float)(i & 255), (float)(1.0 / 255.0)), mean), sd);
    }
  if (j >= ch + cw) return;
  const bool vert = j < ch;
  const int in_size = vert ? img[b].height : img[b].width;
  const int out_size = vert ? rh : rw;
  const int xx = vert ? top + j : left + (j - ch);
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, fs);
  const double ss = __ddiv_rn(1.0, fs);
  const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  int cnt = xmax - xmin;
  if (cnt > kmax) cnt = kmax;                       // cannot happen when kmax comes from sdp_val_preprocess itself
  int *rec = coef + ((long long)b * (ch + cw) + j) * (kmax + 2);
  rec[0] = xmin;
  rec[1] = cnt;
  auto weight = [&](int x) { return prep_bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)); };
  double ww = 0.0;
  for (int x = 0; x < cnt; ++x) ww = __dadd_rn(ww, weight(x));
  for (int x = 0; x < cnt; ++x) {
    double k = weight(x);
    if (ww != 0.0) k = __ddiv_rn(k, ww);
    const double scaled = __dmul_rn(k, (double)(1 << PREP_PRECISION_BITS));
    rec[2 + x] = k < 0.0 ? __double2int_rz(__dadd_rn(-0.5, scaled)) : __double2int_rz(__dadd_rn(0.5, scaled));
  }
  for (int x = cnt; x < kmax; ++x) rec[2 + x] = 0;
}

__device__ __forceinline__ uint32_t prep_clip8(int acc) {       // Resample.c clip8: arithmetic shift, clamp to a byte
  return (uint32_t)min(max(acc >> PREP_PRECISION_BITS, 0), 255);
}

// One CTA walks source rows of one image: row segment -> shared memory -> crop_w x 3 weighted sums.
__global__ void __launch_bounds__(PREP_THREADS)
prep_horizontal_kernel(const uint8_t *__restrict__ pixels, const sdp_image_desc *__restrict__ img,
                       const int *__restrict__ coef, uint8_t *__restrict__ temp, int ch, int cw, int kmax, int temp_rows) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int rl = kmax + 2;
  const int *rec_v = coef + (long long)b * (ch + cw) * rl;
  const int *rec_h = rec_v + (long long)ch * rl;
  const int y0 = rec_v[0];
  const int nrows = rec_v[(ch - 1) * rl] + rec_v[(ch - 1) * rl + 1] - y0;
  const int x0 = rec_h[0];
  const int x1 = rec_h[(cw - 1) * rl] + rec_h[(cw - 1) * rl + 1];
  const int W = img[b].width;
  const int n = (x1 - x0) * 3;                       // bytes of a row segment
  const uint8_t *base = pixels + img[b].offset;
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const uint8_t *src = base + ((long long)(y0 + r) * W + x0) * 3;
    // 32-bit loads for the aligned middle of the segment, bytes for its head and tail; the shared copy keeps the
    // global word alignment (srow[j] <-> src[j])
    const int head = min((int)((4 - (reinterpret_cast<uintptr_t>(src) & 3)) & 3), n);
    uint8_t *srow = prep_smem + ((4 - head) & 3);
    const int nw = (n - head) >> 2;
    const uint32_t *gw = reinterpret_cast<const uint32_t *>(src + head);
    uint32_t *sw = reinterpret_cast<uint32_t *>(srow + head);
    for (int i = tid; i < nw; i += PREP_THREADS) sw[i] = __ldg(gw + i);
    if (tid < head) srow[tid] = src[tid];
    const int tail0 = head + 4 * nw;
    if (tid < n - tail0) srow[tail0 + tid] = src[tail0 + tid];
    __syncthreads();
    uint8_t *dst = temp + ((long long)b * temp_rows + r) * cw * 3;
    for (int idx = tid; idx < cw * 3; idx += PREP_THREADS) {
      const int x = idx / 3, c = idx - 3 * x;
      const int *rec = rec_h + (long long)x * rl;
      const int cnt = rec[1];
      const uint8_t *p = srow + (rec[0] - x0) * 3 + c;
      int acc = 1 << (PREP_PRECISION_BITS - 1);
      for (int k = 0; k < cnt; ++k) acc += (int)p[3 * k] * __ldg(rec + 2 + k);
      dst[idx] = (uint8_t)prep_clip8(acc);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
prep_vertical_kernel(const uint8_t *__restrict__ temp, const int *__restrict__ coef, const float *__restrict__ lut_g,
                     T *__restrict__ out, int ch, int cw, int kmax, int temp_rows) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  float *lut = reinterpret_cast<float *>(prep_smem);             // [3][256]
  uint8_t *urow = prep_smem + 3 * 256 * sizeof(float);          // [cw * 3]
  const int b = blockIdx.y, y = blockIdx.x, tid = threadIdx.x;
  const int rl = kmax + 2;
  const int *rec_v = coef + (long long)b * (ch + cw) * rl;
  const int *rec = rec_v + (long long)y * rl;
  const int first = rec[0] - rec_v[0], cnt = rec[1];
  for (int i = tid; i < 768; i += PREP_THREADS) lut[i] = __ldg(lut_g + i);
  const int rowb = cw * 3;
  const uint8_t *src = temp + ((long long)b * temp_rows + first) * rowb;
  if ((rowb & 3) == 0) {                             // four bytes per thread (rows of the intermediate are 4-byte aligned)
    for (int i4 = tid; i4 < rowb / 4; i4 += PREP_THREADS) {
      int a0, a1, a2, a3;
      a0 = a1 = a2 = a3 = 1 << (PREP_PRECISION_BITS - 1);
      for (int k = 0; k < cnt; ++k) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src + (long long)k * rowb) + i4);
        const int kk = __ldg(rec + 2 + k);
        a0 += (int)(w & 255u) * kk;
        a1 += (int)((w >> 8) & 255u) * kk;
        a2 += (int)((w >> 16) & 255u) * kk;
        a3 += (int)(w >> 24) * kk;
      }
      reinterpret_cast<uint32_t *>(urow)[i4] = prep_clip8(a0) | (prep_clip8(a1) << 8) | (prep_clip8(a2) << 16) | (prep_clip8(a3) << 24);
    }
  } else {
    for (int idx = tid; idx < rowb; idx += PREP_THREADS) {
      int acc = 1 << (PREP_PRECISION_BITS - 1);
      for (int k = 0; k < cnt; ++k) acc += (int)src[(long long)k * rowb + idx] * __ldg(rec + 2 + k);
      urow[idx] = (uint8_t)prep_clip8(acc);
    }
  }
  __syncthreads();
  for (int i = tid; i < rowb; i += PREP_THREADS) {   // planar output: consecutive threads, consecutive x
    const int c = i / cw, x = i - c * cw;
    out[(((long long)b * 3 + c) * ch + y) * cw + x] = from_f<T>(lut[c * 256 + urow[x * 3 + c]]);
  }
}


// Fast path: both passes in ONE kernel, the uint8 intermediate never leaves shared memory.  A CTA owns a band of
// `band` output rows of one image: it stages the source rows the band needs (RS rows at a time, 32-bit loads, one warp
// per row), runs the horizontal taps into a [source rows, crop_w * 3] uint8 tile, then the vertical taps four bytes at a
// time, and writes the band planar through the normalisation table.  Neighbouring bands recompute the few source rows
// they share.  The host sizes the band so that everything fits; batches it cannot fit take the three-kernel path.
struct PrepFused {
  int band, rs, seg_stride, rowb_pad, src_rows;       // geometry
  int coefh_off, coefv_off, lut_off, temp_off, stage_off, ubuf_off, total;   // shared-memory layout (bytes)
};

template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
prep_fused_kernel(const uint8_t *__restrict__ pixels, long long pixels_bytes, const sdp_image_desc *__restrict__ img,
                  const int *__restrict__ coef, const float *__restrict__ lut_g, T *__restrict__ out, int ch, int cw, int kmax,
                  const PrepFused g) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = kmax + 2;
  const int yb0 = blockIdx.x * g.band, nb = min(g.band, ch - yb0);
  int *coefh = reinterpret_cast<int *>(prep_smem + g.coefh_off);
  int *coefv = reinterpret_cast<int *>(prep_smem + g.coefv_off);
  float *lut = reinterpret_cast<float *>(prep_smem + g.lut_off);
  uint8_t *temp = prep_smem + g.temp_off, *stage = prep_smem + g.stage_off, *ubuf = prep_smem + g.ubuf_off;
  const int *rec_v = coef + ((long long)b * (ch + cw) + yb0) * rl;
  const int *rec_h = coef + ((long long)b * (ch + cw) + ch) * rl;
  for (int i = tid; i < cw * rl; i += PREP_THREADS) coefh[i] = __ldg(rec_h + i);
  for (int i = tid; i < nb * rl; i += PREP_THREADS) coefv[i] = __ldg(rec_v + i);
  for (int i = tid; i < 768; i += PREP_THREADS) lut[i] = __ldg(lut_g + i);
  __syncthreads();
  const int s0 = coefv[0], s1 = coefv[(nb - 1) * rl] + coefv[(nb - 1) * rl + 1];   // source rows of this band
  const int x0 = coefh[0], x1 = coefh[(cw - 1) * rl] + coefh[(cw - 1) * rl + 1];   // source columns of the window
  const int W = img[b].width, n = (x1 - x0) * 3;
  const uint8_t *base = pixels + img[b].offset;

  // ---- horizontal pass: row segments arrive with cp.async (aligned 32-bit words covering the segment), one chunk
  //      of RS rows ahead of the chunk being filtered ----
  const uintptr_t pix_end = reinterpret_cast<uintptr_t>(pixels) + pixels_bytes;
  auto stage_chunk = [&](int r0, int buf) {
    const int nr = min(g.rs, s1 - r0);
    for (int r = warp; r < nr; r += PREP_THREADS / 32) {         // one warp stages one row segment
      const uintptr_t src = reinterpret_cast<uintptr_t>(base) + ((long long)(r0 + r) * W + x0) * 3;
      const uintptr_t w0 = src & ~uintptr_t(3);                  // byte j of the segment lands at row base + (src & 3) + j
      const int nw = (int)((src + n + 3 - w0) >> 2);
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(stage + (buf * g.rs + r) * g.seg_stride));
      for (int i = lane; i < nw; i += 32) {
        const uintptr_t ga = w0 + 4 * (uintptr_t)i;
        if (ga + 4 <= pix_end) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4 * i), "l"(ga) : "memory");
        } else {                                                 // the buffer's last, partial word: byte by byte
          for (int e = 0; e < 4; ++e)
            if (ga + e < pix_end) stage[(buf * g.rs + r) * g.seg_stride + 4 * i + e] = *reinterpret_cast<const uint8_t *>(ga + e);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_chunk(s0, 0);
  int cbuf = 0;
  for (int r0 = s0; r0 < s1; r0 += g.rs, cbuf ^= 1) {
    const int nr = min(g.rs, s1 - r0);
    if (r0 + g.rs < s1) {
      stage_chunk(r0 + g.rs, cbuf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *sbuf = stage + cbuf * g.rs * g.seg_stride;
    // thread <-> output column: the column's taps are read once per tap and applied to all staged rows
    for (int x = tid; x < cw; x += PREP_THREADS) {
      const int *rec = coefh + x * rl;
      const int cnt = rec[1];
      const int off = (rec[0] - x0) * 3;
      int a[PREP_THREADS / 32][3];
      int po[PREP_THREADS / 32];                                   // byte offset of the window in each staged row
#pragma unroll
      for (int r = 0; r < PREP_THREADS / 32; ++r) {
        a[r][0] = a[r][1] = a[r][2] = 1 << (PREP_PRECISION_BITS - 1);
        const int rr = min(r, nr - 1);                             // rows past the chunk repeat the last one (discarded)
        const uintptr_t src = reinterpret_cast<uintptr_t>(base) + ((long long)(r0 + rr) * W + x0) * 3;
        po[r] = rr * g.seg_stride + off + (int)(src & 3);
      }
      for (int k = 0; k < cnt; ++k) {
        const int kk = rec[2 + k];
#pragma unroll
        for (int r = 0; r < PREP_THREADS / 32; ++r) {
          const uint8_t *p = sbuf + po[r] + 3 * k;
          a[r][0] += (int)p[0] * kk;
          a[r][1] += (int)p[1] * kk;
          a[r][2] += (int)p[2] * kk;
        }
      }
#pragma unroll
      for (int r = 0; r < PREP_THREADS / 32; ++r)
        if (r < nr) {
          uint8_t *d = temp + (r0 - s0 + r) * g.rowb_pad + 3 * x;
          d[0] = (uint8_t)prep_clip8(a[r][0]);
          d[1] = (uint8_t)prep_clip8(a[r][1]);
          d[2] = (uint8_t)prep_clip8(a[r][2]);
        }
    }
    __syncthreads();
  }

  // ---- vertical pass: four bytes of an output row per item ----
  const int q = g.rowb_pad >> 2;
  for (int it = tid; it < nb * q; it += PREP_THREADS) {
    const int y = it / q, i4 = it - y * q;
    const int *rec = coefv + y * rl;
    const int cnt = rec[1];
    const uint32_t *src = reinterpret_cast<const uint32_t *>(temp + (rec[0] - s0) * g.rowb_pad) + i4;
    int a0, a1, a2, a3;
    a0 = a1 = a2 = a3 = 1 << (PREP_PRECISION_BITS - 1);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t w = src[k * q];
      const int kk = rec[2 + k];
      a0 += (int)(w & 255u) * kk;
      a1 += (int)((w >> 8) & 255u) * kk;
      a2 += (int)((w >> 16) & 255u) * kk;
      a3 += (int)(w >> 24) * kk;
    }
    reinterpret_cast<uint32_t *>(ubuf)[it] = prep_clip8(a0) | (prep_clip8(a1) << 8) | (prep_clip8(a2) << 16) | (prep_clip8(a3) << 24);
  }
  __syncthreads();

  // ---- planar output through the table: a warp writes one (row, channel) line, consecutive lanes consecutive x ----
  for (int yc = warp; yc < nb * 3; yc += PREP_THREADS / 32) {
    const int y = yc / 3, c = yc - 3 * y;
    const uint8_t *u = ubuf + y * g.rowb_pad + c;
    const float *l = lut + c * 256;
    T *o = out + (((long long)b * 3 + c) * ch + yb0 + y) * cw;
    for (int x = lane; x < cw; x += 32) o[x] = from_f<T>(l[u[3 * x]]);
  }
}

struct PrepPlan {
  int kmax, temp_rows, max_seg_bytes, top, left;
  size_t desc_off, coef_off, lut_off, temp_off, total;
  PrepFused fused;          // fused.band == 0: the batch takes the three-kernel path
};

// Largest band (<= 32 output rows) whose working set fits `budget` bytes of shared memory for EVERY image of the batch.
static void prep_fused_geometry(const sdp_image_desc *images, int B, int rh, int ch, int cw, const PrepPlan &p, int budget,
                                PrepFused *g) {
  const int rl = p.kmax + 2;
  g->band = 0;
  g->rowb_pad = (cw * 3 + 3) & ~3;
  g->seg_stride = (p.max_seg_bytes + 8 + 15) / 16 * 16;
  g->rs = PREP_THREADS / 32;     // rows staged per chunk = accumulator rows per thread in the horizontal pass
  for (int band = min(32, ch); band >= 1; band = band > 8 ? band - 8 : band - 1) {
    int src_rows = 1;
    for (int b = 0; b < B; ++b) {
      const PrepAxis av = prep_axis_host(images[b].height, rh);
      for (int y0 = 0; y0 < ch; y0 += band) {
        int lo, cnt, lo2, cnt2;
        prep_bounds_host(av, images[b].height, p.top + y0, &lo, &cnt);
        prep_bounds_host(av, images[b].height, p.top + min(y0 + band, ch) - 1, &lo2, &cnt2);
        src_rows = max(src_rows, lo2 + cnt2 - lo);
      }
    }
    auto up = [](int v) { return (v + 15) / 16 * 16; };
    PrepFused t = *g;
    t.band = band;
    t.src_rows = src_rows;
    t.coefh_off = 0;
    t.coefv_off = t.coefh_off + up(cw * rl * 4);
    t.lut_off = t.coefv_off + up(band * rl * 4);
    t.temp_off = t.lut_off + 768 * 4;
    t.stage_off = t.temp_off + up(src_rows * t.rowb_pad);
    t.ubuf_off = t.stage_off + up(2 * t.rs * t.seg_stride);     // double-buffered
    t.total = t.ubuf_off + up(band * t.rowb_pad);
    if (t.total <= budget) {
      *g = t;
      return;
    }
  }
}

static int prep_plan(const sdp_image_desc *images, int B, int rh, int rw, int ch, int cw, int path, PrepPlan *p) {
  SDP_CHECK(images && B > 0 && rh > 0 && rw > 0 && ch > 0 && cw > 0, "sdp_val_preprocess: bad arguments");
  SDP_CHECK(ch <= rh && cw <= rw, "sdp_val_preprocess: crop %dx%d larger than the resized image %dx%d (the reference would pad)",
            ch, cw, rh, rw);
  p->top = crop_anchor(rh, ch);
  p->left = crop_anchor(rw, cw);
  p->kmax = 1;
  p->temp_rows = 1;
  p->max_seg_bytes = 4;
  for (int b = 0; b < B; ++b) {
    const int H = images[b].height, W = images[b].width;
    SDP_CHECK(H > 0 && W > 0 && images[b].offset >= 0, "sdp_val_preprocess: image %d has size %dx%d, offset %lld", b, H, W,
              (long long)images[b].offset);
    const PrepAxis av = prep_axis_host(H, rh), ah = prep_axis_host(W, rw);
    p->kmax = max(p->kmax, max(av.ksize, ah.ksize));
    int lo, cnt, lo2, cnt2;
    prep_bounds_host(av, H, p->top, &lo, &cnt);
    prep_bounds_host(av, H, p->top + ch - 1, &lo2, &cnt2);
    p->temp_rows = max(p->temp_rows, lo2 + cnt2 - lo);
    prep_bounds_host(ah, W, p->left, &lo, &cnt);
    prep_bounds_host(ah, W, p->left + cw - 1, &lo2, &cnt2);
    p->max_seg_bytes = max(p->max_seg_bytes, (lo2 + cnt2 - lo) * 3);
  }
  // fused path: two CTAs per SM when the band fits 100 KB, else one CTA with up to 200 KB; path == 1 asks for the
  // three-kernel path (the one batches with oversized tap tables take anyway)
  p->fused.band = 0;
  if (path != 1) {
    prep_fused_geometry(images, B, rh, ch, cw, *p, 100 * 1024, &p->fused);
    if (p->fused.band < 8) prep_fused_geometry(images, B, rh, ch, cw, *p, 200 * 1024, &p->fused);
  }
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  p->desc_off = 0;
  p->coef_off = up((size_t)B * sizeof(sdp_image_desc));
  p->lut_off = p->coef_off + up((size_t)B * (ch + cw) * (p->kmax + 2) * sizeof(int));
  p->temp_off = p->lut_off + up(768 * sizeof(float));
  // the intermediate of the three-kernel path is always part of the size, so that the answer does not depend on the path
  p->total = p->temp_off + up((size_t)B * p->temp_rows * cw * 3);
  return 0;
}

}  // namespace sdp

using namespace sdp;

extern "C" int64_t sdp_val_preprocess_workspace_bytes(const sdp_image_desc *images, int B, int resize_h, int resize_w,
                                                      int crop_h, int crop_w, int path) {
  PrepPlan p;
  if (prep_plan(images, B, resize_h, resize_w, crop_h, crop_w, path, &p) != 0) return -1;
  return (int64_t)p.total;
}

extern "C" int sdp_val_preprocess(const uint8_t *pixels, int64_t pixels_bytes, const sdp_image_desc *images, int B, int resize_h, int resize_w,
                                  int crop_h, int crop_w, const float *mean, const float *std_, void *workspace,
                                  int64_t workspace_bytes, void *out, int out_dtype, int path, void *stream) {
  PrepPlan p;
  if (int rc = prep_plan(images, B, resize_h, resize_w, crop_h, crop_w, path, &p)) return rc;
  SDP_CHECK(pixels && mean && std_ && workspace && out, "sdp_val_preprocess: null pointer");
  SDP_CHECK(out_dtype == SDP_F32 || out_dtype == SDP_BF16, "sdp_val_preprocess: out_dtype %d", out_dtype);
  SDP_CHECK(workspace_bytes >= (int64_t)p.total, "sdp_val_preprocess: workspace of %lld bytes, %lld needed",
            (long long)workspace_bytes, (long long)p.total);
  SDP_CHECK((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sdp_val_preprocess: workspace not 16-byte aligned");
  SDP_CHECK(B <= 65535, "sdp_val_preprocess: at most 65535 images per call");
  SDP_CHECK((reinterpret_cast<uintptr_t>(pixels) & 3) == 0, "sdp_val_preprocess: pixel buffer not 4-byte aligned");
  for (int b = 0; b < B; ++b)
    SDP_CHECK(images[b].offset + 3ll * images[b].height * images[b].width <= pixels_bytes,
              "sdp_val_preprocess: image %d (%dx%d at byte %lld) reaches past the %lld-byte pixel buffer", b, images[b].height,
              images[b].width, (long long)images[b].offset, (long long)pixels_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  sdp_image_desc *d_img = reinterpret_cast<sdp_image_desc *>(ws + p.desc_off);
  int *d_coef = reinterpret_cast<int *>(ws + p.coef_off);
  float *d_lut = reinterpret_cast<float *>(ws + p.lut_off);
  uint8_t *d_temp = ws + p.temp_off;
  SDP_CUDA(cudaMemcpyAsync(d_img, images, (size_t)B * sizeof(sdp_image_desc), cudaMemcpyHostToDevice, st));
  prep_coeffs_kernel<<<dim3((crop_h + crop_w + 127) / 128, B), 128, 0, st>>>(d_img, resize_h, resize_w, crop_h, crop_w, p.top,
                                                                             p.left, p.kmax, d_coef, d_lut, mean[0], mean[1],
                                                                             mean[2], std_[0], std_[1], std_[2]);
  SDP_LAUNCH_OK();
  if (p.fused.band > 0) {
    const dim3 grid((crop_h + p.fused.band - 1) / p.fused.band, B);
    if (out_dtype == SDP_F32) {
      if (p.fused.total > 48 * 1024)      // per device, cheap
        SDP_CUDA(cudaFuncSetAttribute(prep_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      prep_fused_kernel<float><<<grid, PREP_THREADS, p.fused.total, st>>>(pixels, pixels_bytes, d_img, d_coef, d_lut,
                                                                          reinterpret_cast<float *>(out), crop_h, crop_w, p.kmax,
                                                                          p.fused);
    } else {
      if (p.fused.total > 48 * 1024)
        SDP_CUDA(cudaFuncSetAttribute(prep_fused_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      prep_fused_kernel<bf16><<<grid, PREP_THREADS, p.fused.total, st>>>(pixels, pixels_bytes, d_img, d_coef, d_lut,
                                                                         reinterpret_cast<bf16 *>(out), crop_h, crop_w, p.kmax,
                                                                         p.fused);
    }
    SDP_LAUNCH_OK();
    return 0;
  }
  // ---- general path: intermediate in global memory ----
  const int seg_smem = p.max_seg_bytes + 16;
  SDP_CHECK(seg_smem <= 200 * 1024, "sdp_val_preprocess: a source row segment of %d bytes does not fit shared memory",
            p.max_seg_bytes);
  if (seg_smem > 48 * 1024)
    SDP_CUDA(cudaFuncSetAttribute(prep_horizontal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int row_ctas = min(p.temp_rows, max(1, (148 * 8 + B - 1) / B));   // >= 8 CTAs per SM over the batch
  prep_horizontal_kernel<<<dim3(row_ctas, B), PREP_THREADS, seg_smem, st>>>(pixels, d_img, d_coef, d_temp, crop_h, crop_w,
                                                                            p.kmax, p.temp_rows);
  SDP_LAUNCH_OK();
  const int v_smem = 3 * 256 * (int)sizeof(float) + (crop_w * 3 + 15) / 16 * 16;
  SDP_CHECK(v_smem <= 48 * 1024, "sdp_val_preprocess: crop width %d too large", crop_w);
  if (out_dtype == SDP_F32)
    prep_vertical_kernel<float><<<dim3(crop_h, B), PREP_THREADS, v_smem, st>>>(d_temp, d_coef, d_lut, reinterpret_cast<float *>(out),
                                                                               crop_h, crop_w, p.kmax, p.temp_rows);
  else
    prep_vertical_kernel<bf16><<<dim3(crop_h, B), PREP_THREADS, v_smem, st>>>(d_temp, d_coef, d_lut, reinterpret_cast<bf16 *>(out),
                                                                              crop_h, crop_w, p.kmax, p.temp_rows);
  SDP_LAUNCH_OK();
  return 0;
}
