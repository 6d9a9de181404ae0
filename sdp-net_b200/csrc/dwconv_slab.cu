// Channel-stationary tensor-core depthwise conv + channel LayerNorm (bf16, grids up to 16 x 16 with an even number
// of tokens, k in {3,5,7}).  Reference: layer_norm_1 + conv2d[0] of ConvMixer, layers.py:102 (:12-24 + :73-78).
//
// For one channel the k x k 'same' conv of the zero-haloed plane P is, per tap row dy and per block of eight
// output columns x0..x0+7, one 16 x 8 x 16 matrix product
//      Out[y, x0 + n] += sum_k P[y + dy, x0 + k] * T_dy[k, n],     T_dy[k, n] = w[dy][k - n - 1]  (k = 7)
// (a banded Toeplitz block of the row's taps; shift-invariant, so ONE block serves every x0).  That makes the
// whole weight operand of a channel 2 registers per tap row and lane -- 14 for k = 7 -- small enough to stay in
// registers for the lifetime of the CTA.  So the kernel is organised around the weights, not around the image:
//   * a CTA owns 32 channels (each of its 8 warps owns 4, B fragments loaded once) and loops over images;
//   * the [grid rows, 16 column slots, 32 channels] slice of an image (64 B per token = two full sectors) arrives by
//     TMA through a 4-D view [image][grid row][grid column][channel] of the token-major stream -- 64B-swizzled, column
//     slots past a narrow row's end zero-filled by the TMA unit -- into a three-deep mbarrier ring, together with the
//     image's token statistics (one bulk copy); one elected thread issues both, nobody computes addresses;
//   * the warps normalise and transpose the slice into per-channel planes with ldmatrix.trans (48-byte plane rows:
//     conflict-free), run 14 mma.sync.m16n8k16 per channel on ldmatrix'ed plane rows, and write the [tokens, 32]
//     result back into the ring slot the slice came in (64B-swizzled, free once the planes are built), which leaves
//     as ONE TMA store through the same 4-D view of `out`; the slot is refilled once that store has read it: two
//     CTA barriers per image (planes complete; result tile complete);
//   * the token statistics (over all C channels) come from the producer GEMM's epilogue (or a warp-per-row pass),
//     because a slab CTA never sees the whole row.
// Two CTAs per SM overlap one's transposes with the other's MMAs.  12544 scalar FMAs per channel and image
// become 14 MMAs; the per-image instruction count is what bounds the kernel, hence no index arithmetic in the loop.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc5.cuh"

namespace sdp {

constexpr int DS_THREADS = 256;
constexpr int DS_CH = 32;                 // channels per CTA
constexpr int DS_HL = 4;                  // left halo columns (keeps 8-column blocks 16-byte aligned)
constexpr int DS_PROW = 48;               // bytes per plane row: 24 bf16 = 4 halo + 16 + 4 halo
constexpr int DS_RAWP = 64;               // bytes per token in the raw / out tiles; 16-byte chunk c of slot t sits at c ^ ((t >> 1) & 3)
                                          // (the TMA unit's 64B swizzle): conflict-free for the ldmatrix reads (8 tokens x 16 B)
constexpr int DS_STAGES = 4;              // ring slots per CTA: a slot holds an image's slice, then (in place) its result

__device__ __forceinline__ void ds_mma(float *c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ds_ldmatrix_x4(uint32_t addr, uint32_t *r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ds_ldmatrix_x2(uint32_t addr, uint32_t *r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ds_ldmatrix_x4_trans(uint32_t addr, uint32_t *r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// (mean, rstd) of every spatial token row over all C channels, compact [B][Tn]: warp per row, the row held in
// registers (exact two-pass).
template <int NI>                          // 16-byte chunks per lane: C <= 256 * NI
__global__ void __launch_bounds__(256)
token_stats_kernel(const bf16 *__restrict__ act, float2 *__restrict__ stats, long long rows, int Tn, int R, int C,
                   float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long src_row = (row / Tn) * (R + Tn) + R + row % Tn;
  const uint4 *rv = reinterpret_cast<const uint4 *>(act + src_row * C);
  const int nv = C >> 3;
  float v[NI][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const uint4 u = __ldg(rv + idx);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][2 * j] = __uint_as_float(w[j] << 16);
        v[i][2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        s += v[i][2 * j] + v[i][2 * j + 1];
      }
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NI; ++i)
    if (lane + 32 * i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  q = warp_sum(q);
  if (lane == 0) stats[row] = make_float2(mean, 1.0f / sqrtf(q / (float)C + eps));
}

// (mean, rstd) of the spatial tokens from the (sum, sumsq) column parts a producer GEMM wrote for every row of the
// [B*S, C] activation (row b*S + R + t  ->  compact index b*Tn + t).
__global__ void __launch_bounds__(256)
token_stats_from_parts_kernel(const float2 *__restrict__ parts_in, float2 *__restrict__ stats, long long rows, int Tn,
                              int R, int parts, int C, float eps) {
  const long long row = (long long)blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  const float2 *p = parts_in + ((row / Tn) * (R + Tn) + R + row % Tn) * parts;
  float s1 = 0.0f, s2 = 0.0f;
  for (int i = 0; i < parts; ++i) {
    const float2 v = __ldg(p + i);
    s1 += v.x;
    s2 += v.y;
  }
  const float mean = s1 / (float)C;
  stats[row] = make_float2(mean, 1.0f / sqrtf(fmaxf(s2 / (float)C - mean * mean, 0.0f) + eps));
}

template <int KS>
struct DsLayout {
  static constexpr int ROWS = 16 + KS - 1;
  int raw, otile, stat, planes, bars, total;
  int plane_bytes, tile_bytes, stat_bytes;
  __host__ __device__ explicit DsLayout(int Gh, int Tn) {
    const int TP = Gh * 16;                                 // token slots of the raw / out tiles: grid rows padded to 16
    // plane stride: a multiple of 16 bytes whose word count is 12 (mod 32): the eight channels a transposing
    // store touches at once then land in eight different bank quads
    int pb = (ROWS * DS_PROW + 15) / 16 * 16;
    while ((pb / 4) % 32 != 12) pb += 16;
    plane_bytes = pb;
    tile_bytes = (TP * DS_RAWP + 1023) / 1024 * 1024;       // swizzled TMA boxes: 1 KB-aligned
    stat_bytes = (Tn * 8 + 15) / 16 * 16;
    raw = 0;                                                // [DS_STAGES][TP][64 B]: input slice, later the output tile
    otile = raw;
    stat = raw + DS_STAGES * tile_bytes;                    // [DS_STAGES][Tn] float2 (mean, rstd)
    planes = (stat + DS_STAGES * stat_bytes + 15) / 16 * 16;   // [32][plane_bytes], zero halo
    bars = planes + DS_CH * pb;                             // DS_STAGES mbarriers
    total = bars + 64 + 1024;                               // + slack for the 1 KB alignment of the base
  }
};

template <int KS>
__global__ void __launch_bounds__(DS_THREADS, 2)
ln_dwconv_slab_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                      const float2 *__restrict__ stats, const float *__restrict__ gamma,
                      const float *__restrict__ beta, const float *__restrict__ wdw, const float *__restrict__ bdw,
                      bf16 *__restrict__ out, int B, int Gh, int Gw, int C, int R, int img_per_cta) {
  constexpr int lo = (KS - 1) / 2;
  extern __shared__ uint8_t ds_raw[];
  uint8_t *ds_smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(ds_raw) + 1023) & ~uintptr_t(1023));
  const int Tn = Gh * Gw, S = R + Tn;
  const DsLayout<KS> L(Gh, Tn);
  const int PB = L.plane_bytes;
  const int TP = Gh * 16;
  const uint32_t sbase = smem_u32(ds_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int c0 = blockIdx.x * DS_CH;
  const int b_begin = blockIdx.y * img_per_cta, b_end = min(B, b_begin + img_per_cta);
  if (b_begin >= b_end) return;
  auto full_bar = [&](int st) { return sbase + L.bars + 8u * st; };
  // images are walked from the last to the first (loop index b -> image B - 1 - b): the producer GEMM wrote the
  // activations in ascending order, so the tail of the batch is still in L2, and the consumer GEMM starts at image 0
  auto issue = [&](int bi) {               // one thread: the [Gh, 16, 32-channel] slice of an image + its statistics
    const int st = (bi - b_begin) % DS_STAGES, b = B - 1 - bi;
    const uint32_t bar = full_bar(st);
    mbar_expect_tx(bar, (uint32_t)(TP * DS_RAWP + Tn * 8));
    tma_load_4d(sbase + L.raw + st * L.tile_bytes, &tmIn, bar, c0, 0, 0, b);
    bulk_load(sbase + L.stat + st * L.stat_bytes, stats + (long long)b * Tn, (uint32_t)(Tn * 8), bar);   // Tn is even: 16-byte units
  };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    for (int st = 0; st < DS_STAGES; ++st) mbar_init(full_bar(st), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int bi = b_begin; bi < b_end && bi < b_begin + DS_STAGES; ++bi) issue(bi);
  }

  // zero the planes once: halo cells are never written again
  for (int i = tid; i < DS_CH * PB / 16; i += DS_THREADS)
    reinterpret_cast<uint4 *>(ds_smem + L.planes)[i] = make_uint4(0, 0, 0, 0);

  // ---- the weights of this warp's four channels as Toeplitz B fragments, for the lifetime of the CTA ----
  // lane (g, q) holds (k = 2q, 2q+1 | n = g) and (k = 2q+8, 2q+9 | n = g); tap column kx = k - n + lo - HL
  uint32_t bfr[4][KS][2];
  float bias[4];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int ch = c0 + warp * 4 + cc;
    bias[cc] = bdw ? __ldg(bdw + ch) : 0.0f;
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      auto tap = [&](int k) {
        const int kx = k - g + lo - DS_HL;
        return (kx >= 0 && kx < KS) ? __ldg(wdw + (long long)(dy * KS + kx) * C + ch) : 0.0f;
      };
      bfr[cc][dy][0] = pack_bf16x2(tap(2 * q), tap(2 * q + 1));
      bfr[cc][dy][1] = pack_bf16x2(tap(2 * q + 8), tap(2 * q + 9));
    }
  }
  // transform constants: after ldmatrix.trans matrix m hands this lane channel 8m + g of tokens 2q, 2q + 1
  float gm[4], bt[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    gm[m] = __ldg(gamma + c0 + 8 * m + g);
    bt[m] = __ldg(beta + c0 + 8 * m + g);
  }
  // ldmatrix lane addressing inside a plane: matrices (rows 0-7 | 8-15) x (cols 0-7 | 8-15), then cols 16-23
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * DS_PROW + (lane >> 4) * 16);
  const uint32_t a_lane2 = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * DS_PROW + 32);
  const uint32_t planes_addr = sbase + L.planes;
  const bool full_rows = Gw == 16;         // no padded column slots: every slot is a token
  __syncthreads();                         // barriers initialised, planes zeroed

  for (int b = b_begin; b < b_end; ++b) {
    const int it = b - b_begin, st = it % DS_STAGES;
    mbar_wait(full_bar(st), (uint32_t)((it / DS_STAGES) & 1));       // slice b and its statistics have landed

    // ---- normalise + transpose: [token][channel] rows -> per-channel planes (two tokens per 32-bit store) ----
    const uint32_t raw_addr = sbase + L.raw + st * L.tile_bytes;
    const float2 *s_stat = reinterpret_cast<const float2 *>(ds_smem + L.stat + st * L.stat_bytes);
    for (int t0 = warp * 8; t0 < TP; t0 += 8 * (DS_THREADS / 32)) {      // eight slots of one grid row
      const int y = t0 >> 4, x = (t0 & 15) + 2 * q;
      if (!full_rows && (t0 & 15) >= Gw) continue;            // the padded half of a narrow row (warp-uniform)
      uint32_t r4[4];
      ds_ldmatrix_x4_trans(raw_addr + (t0 + (lane & 7)) * DS_RAWP + (((lane >> 3) ^ ((lane >> 1) & 3)) << 4), r4);
      const bool in0 = full_rows || x < Gw, in1 = full_rows || x + 1 < Gw;   // slots past the row end are zero-filled
      const float4 st01 = full_rows ? *reinterpret_cast<const float4 *>(s_stat + t0 + 2 * q)
                                    : make_float4(s_stat[in0 ? y * Gw + x : 0].x, s_stat[in0 ? y * Gw + x : 0].y,
                                                  s_stat[in1 ? y * Gw + x + 1 : 0].x, s_stat[in1 ? y * Gw + x + 1 : 0].y);
      const uint32_t cell = planes_addr + (y + lo) * DS_PROW + (DS_HL + x) * 2;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float v0 = (__uint_as_float(r4[m] << 16) - st01.x) * st01.y;
        const float v1 = (__uint_as_float(r4[m] & 0xffff0000u) - st01.z) * st01.w;
        const uint32_t pk = pack_bf16x2(in0 ? fmaf(v0, gm[m], bt[m]) : 0.0f, in1 ? fmaf(v1, gm[m], bt[m]) : 0.0f);
        if (in0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(cell + (8 * m + g) * PB), "r"(pk) : "memory");   // halo stays 0
      }
    }
    __syncthreads();                       // planes of image b complete: its ring slot now takes the result

    // ---- MMA: four channels per warp, 2 * KS MMAs each ----
    float acc[4][2][4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[cc][n][j] = bias[cc];
      const uint32_t pbase = planes_addr + (warp * 4 + cc) * PB;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        uint32_t a[4], a2[2];
        ds_ldmatrix_x4(pbase + dy * DS_PROW + a_lane, a);        // (r0-7,c0-7) (r8-15,c0-7) (r0-7,c8-15) (r8-15,c8-15)
        ds_ldmatrix_x2(pbase + dy * DS_PROW + a_lane2, a2);      // (r0-7,c16-23) (r8-15,c16-23)
        ds_mma(acc[cc][0], a[0], a[1], a[2], a[3], bfr[cc][dy][0], bfr[cc][dy][1]);
        ds_mma(acc[cc][1], a[2], a[3], a2[0], a2[1], bfr[cc][dy][0], bfr[cc][dy][1]);
      }
    }
    const uint32_t otile_addr = raw_addr;
    // C fragments: (y = g | g + 8, x = 8n + 2q, +1); the warp's four channels of a token are 8 bytes of the token's
    // 64-byte row: 16-byte chunk warp / 2 (64B swizzle: ^ (slot >> 1) & 3 = q), half warp & 1
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int y = g + 8 * hh, x = 8 * n + 2 * q;
        if (y < Gh) {
          // the bank of a store depends on the token's parity, q and the warp only -- not on y -- so lanes of odd g take
          // the two tokens of their pair in the other order: eight bank positions per instruction instead of four
          const uint32_t lo_e[2] = {pack_bf16x2(acc[0][n][2 * hh], acc[1][n][2 * hh]), pack_bf16x2(acc[0][n][2 * hh + 1], acc[1][n][2 * hh + 1])};
          const uint32_t hi_e[2] = {pack_bf16x2(acc[2][n][2 * hh], acc[3][n][2 * hh]), pack_bf16x2(acc[2][n][2 * hh + 1], acc[3][n][2 * hh + 1])};
          const bool odd = g & 1;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ee = e ^ (int)odd;
            const uint32_t lo2 = odd ? lo_e[e ^ 1] : lo_e[e], hi2 = odd ? hi_e[e ^ 1] : hi_e[e];
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(otile_addr + (y * 16 + x + ee) * DS_RAWP +
                                                                  ((((warp >> 1) ^ q) << 4) | ((warp & 1) << 3))),
                         "r"(lo2), "r"(hi2)
                         : "memory");
          }
        }
      }
    fence_proxy_async_smem();
    __syncthreads();                       // result tile complete, planes free for the next image
    // the [Gh, Gw, 32] result leaves as one TMA store (column slots past Gw are clipped); once the PREVIOUS image's
    // store has read its slot, that slot takes the slice of image b - 1 + DS_STAGES.  The R register rows of the slab
    // are zero (the consumer GEMM passes those rows through).
    if (tid == 0) {
      tma_store_4d(&tmOut, otile_addr, c0, 0, 0, B - 1 - b);
      bulk_commit();
      if (b > b_begin && b - 1 + DS_STAGES < b_end) {
        bulk_wait_read<1>();
        issue(b - 1 + DS_STAGES);
      }
    }
    bf16 *dst = out + (long long)(B - 1 - b) * S * C + c0;
    for (int i = tid - 32; i >= 0 && i < R * 4; i += DS_THREADS - 32)
      *reinterpret_cast<uint4 *>(dst + (long long)(i >> 2) * C + (i & 3) * 8) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) bulk_wait_read<0>();       // shared memory outlives the last store's read
}

static int num_sms_ds() {                 // per call: a process may drive several devices
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

template <int KS>
static int launch_slab(const void *act, const float *producer, int parts, float *token_stats, const float *gamma,
                       const float *beta, const float *wdw, const float *bdw, void *out, int B, int Gh, int Gw, int C, int R,
                       float eps, cudaStream_t st) {
  const int Tn = Gh * Gw, S = R + Tn;
  const long long rows = (long long)B * Tn;
  const unsigned sblocks = (unsigned)((rows + 7) / 8);
  float2 *ts = reinterpret_cast<float2 *>(token_stats);
  if (producer != nullptr && parts > 0)
    token_stats_from_parts_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2 *>(producer), ts, rows,
                                                                                    Tn, R, parts, C, eps);
  else if (C <= 256) token_stats_kernel<1><<<sblocks, 256, 0, st>>>((const bf16 *)act, ts, rows, Tn, R, C, eps);
  else if (C <= 512) token_stats_kernel<2><<<sblocks, 256, 0, st>>>((const bf16 *)act, ts, rows, Tn, R, C, eps);
  else if (C <= 768) token_stats_kernel<3><<<sblocks, 256, 0, st>>>((const bf16 *)act, ts, rows, Tn, R, C, eps);
  else if (C <= 1024) token_stats_kernel<4><<<sblocks, 256, 0, st>>>((const bf16 *)act, ts, rows, Tn, R, C, eps);
  else token_stats_kernel<8><<<sblocks, 256, 0, st>>>((const bf16 *)act, ts, rows, Tn, R, C, eps);
  SDP_LAUNCH_OK();
  const DsLayout<KS> L(Gh, Tn);
  auto kern = ln_dwconv_slab_kernel<KS>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));   // per device, cheap
  // [image][grid row][grid column][channel] views of the patch rows of act / out (behind the R register rows)
  CUtensorMap tin, tout;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Gw, (uint64_t)Gh, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)C, (uint64_t)Gw * C, (uint64_t)S * C};
    const uint32_t box[4] = {DS_CH, 16, (uint32_t)Gh, 1};
    if (int rc = make_tensor_map_bf16_4d(reinterpret_cast<const bf16 *>(act) + (long long)R * C, dims, str, box, &tin)) return rc;
    if (int rc = make_tensor_map_bf16_4d(reinterpret_cast<bf16 *>(out) + (long long)R * C, dims, str, box, &tout)) return rc;
  }
  const int slabs = C / DS_CH;
  // two CTAs per SM; image groups sized so that the grid is (just under) one wave
  int groups = (2 * num_sms_ds()) / slabs;
  if (groups < 1) groups = 1;
  if (groups > B) groups = B;
  const int per = (B + groups - 1) / groups;
  groups = (B + per - 1) / per;
  kern<<<dim3(slabs, groups), DS_THREADS, L.total, st>>>(tin, tout, reinterpret_cast<const float2 *>(token_stats),
                                                         gamma, beta, wdw, bdw, (bf16 *)out, B, Gh, Gw, C, R, per);
  SDP_LAUNCH_OK();
  return 0;
}

}  // namespace sdp

using namespace sdp;

extern "C" int sdp_ln_dwconv_slab_ok(int Gh, int Gw, int C, int k, int dtype) {
  return dtype == SDP_BF16 && (k == 3 || k == 5 || k == 7) && Gh >= 1 && Gh <= 16 && Gw >= 1 && Gw <= 16 &&
         (Gh * Gw) % 2 == 0 && C % DS_CH == 0 && C <= 2048;
}

extern "C" int sdp_ln_dwconv_slab_stats(const void *act, const float *producer_stats, int parts, float *token_stats,
                                        const float *gamma, const float *beta, const float *wdw, const float *bdw,
                                        void *out, int B, int Gh, int Gw, int C, int k, int R, float eps, void *stream) {
  SDP_CHECK(act && token_stats && gamma && beta && wdw && out, "sdp_ln_dwconv_slab: null pointer");
  SDP_CHECK(sdp_ln_dwconv_slab_ok(Gh, Gw, C, k, SDP_BF16), "sdp_ln_dwconv_slab: shape not covered (ask sdp_ln_dwconv_slab_ok)");
  SDP_CHECK(B > 0 && R >= 0, "sdp_ln_dwconv_slab: bad sizes");
  SDP_CHECK(producer_stats == nullptr || parts > 0, "sdp_ln_dwconv_slab: producer statistics need parts > 0");
  SDP_CHECK(producer_stats != token_stats, "sdp_ln_dwconv_slab: producer_stats and token_stats must not alias");
  SDP_CHECK(act != out, "sdp_ln_dwconv_slab: must not run in place (spatial neighbours are read)");
  SDP_CHECK((reinterpret_cast<uintptr_t>(act) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(token_stats) & 15) == 0 && (reinterpret_cast<uintptr_t>(producer_stats) & 7) == 0,
            "sdp_ln_dwconv_slab: act, out and token_stats must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (k == 7) return launch_slab<7>(act, producer_stats, parts, token_stats, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
  if (k == 5) return launch_slab<5>(act, producer_stats, parts, token_stats, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
  return launch_slab<3>(act, producer_stats, parts, token_stats, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, R, eps, st);
}

extern "C" int sdp_ln_dwconv_slab(const void *act, float *token_stats, const float *gamma, const float *beta,
                                  const float *wdw, const float *bdw, void *out, int B, int Gh, int Gw, int C, int k,
                                  int R, float eps, void *stream) {
  return sdp_ln_dwconv_slab_stats(act, nullptr, 0, token_stats, gamma, beta, wdw, bdw, out, B, Gh, Gw, C, k, R, eps, stream);
}
