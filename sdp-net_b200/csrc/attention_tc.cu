// Attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for the short SdP-Net sequence.
// Reference: F.scaled_dot_product_attention at layers.py:289-291 (q, k already LayerNorm-ed by the QKV
// GEMM epilogue, layers.py:286).
//
// Persistent: one CTA per SM walks (image, head) items; barriers, TMEM and tensor maps are set up once, and the next
// item's Q tile 0 and K arrive (as soon as the last score MMA of the current item has retired) under the current item's
// last softmax, its V under the next item's first scores.  The whole K and V of the sequence (<= 288 keys) sit in shared memory as
// 64B-swizzled 32-column tiles loaded by TMA straight out of the [B*S, 3C] QKV buffer.  Per 128-query tile:
//   S = Q K^T   tcgen05.mma, K-major A (Q) and B (K), fp32 scores in TMEM columns [0, KEYS)
//   softmax     four warps, one thread per query row: tcgen05.ld the row, max, exp2, row sum; P (bf16) goes
//               back to shared memory in the 64B-swizzled K-major layout the next MMA reads
//   O = P V     tcgen05.mma, A = P (K-major), B = V read as an MN-major operand (V is [keys, d] in memory,
//               i.e. d-contiguous: no transpose anywhere), fp32 O in TMEM columns [384, 384 + d)
//   epilogue    tcgen05.ld O, scale by 1 / row sum, bf16, 16-byte stores into the token-major output
// Warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..9 = softmax / epilogue.
// No online rescaling is needed: a full score row fits in TMEM.
#include "tc5.cuh"
#if AT_TRACE
#define AT_TDECL unsigned char tcode[200]; long long tclk[200]; int tn = 0;
#define AT_T(code) do { if (blockIdx.x == 0 && lane == 0 && tn < 200) { tcode[tn] = (unsigned char)(code); tclk[tn++] = clock64(); } } while (0)
#define AT_DUMP(name) do { if (blockIdx.x == 0 && lane == 0) for (int i = 0; i < tn; ++i) printf("%s w%d %3d %lld\n", name, warp, (int)tcode[i], tclk[i]); } while (0)
#else
#define AT_TDECL
#define AT_T(code)
#define AT_DUMP(name)
#endif

namespace sdp {


// The kernel's hot code has to stay inside the 32 KB instruction cache (it ran at an 87 % hit rate when every
// wait carried its own copy of the watchdog loop and the MMA issue loops were unrolled): waits are one try_wait
// with the bounded poll loop out of line.
__device__ __noinline__ void at_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void at_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) at_wait_slow(bar, parity);
}

constexpr int AT_MT = 128;          // query rows per tile
constexpr int AT_THREADS = 320;        // TMA warp, MMA warp, 8 softmax warps
constexpr float AT_ONE_PASS_BOUND = 60.0f;   // largest |score| (nats) the one-pass softmax accepts

// 64B-swizzled tiles: rows of 64 bytes (32 bf16), groups of 8 rows = 512 bytes.
// K-major operand descriptor (A, and B = K): layout type 4 (SWIZZLE_64B), SBO = 512.
__device__ __forceinline__ uint64_t at_desc_kmajor(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// MN-major B operand (V): 32 d-columns contiguous per 64-byte row, 8 key rows per 512-byte group (SBO), the
// next 32 d-columns `lbo` bytes further (LBO).
__device__ __forceinline__ uint64_t at_desc_mnmajor(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// Descriptor words.  Low: [0,14) address >> 4, [16,30) leading byte offset >> 4.  High: [0,14) stride byte offset >> 4
// (one 8-row group), bit 14 = descriptor version 1, [29,32) swizzle mode (2 = 128B, 4 = 64B).
constexpr uint32_t AT_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t AT_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t at_lo(uint32_t saddr, uint32_t lbo) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo >> 4) & 0x3FFF) << 16);
}
// The MMA warp runs CONVERGED: all 32 lanes follow the control flow and compute the (warp-uniform) descriptors, so
// they live in uniform registers; only the tcgen05 instruction itself is predicated on the elected lane.  Issued from
// inside an `if (lane == 0)` region every MMA went through an ELECT / 5 x R2UR.BROADCAST waterfall loop (~160 cycles
// per issue against 48-72 tensor-core cycles per MMA).
__device__ __forceinline__ uint32_t at_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void at_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void at_commit(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(leader)
      : "memory");
}
__host__ __device__ constexpr uint32_t at_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// One chunk of scores: 32 columns (16 for a tail) of this thread's TMEM lane.
__device__ __forceinline__ void at_ld_chunk(uint32_t taddr, int n, float (&v)[32]) {
  if (n >= 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}
// Columns that were not loaded (a 16-wide tail) or lie past the end of the sequence become -inf: they lose the
// maximum and exponentiate to 0.  Only the last chunk of a row needs it; the empty asm keeps the block a real
// (warp-uniform) branch instead of 32 predicated selects on every chunk.
__device__ __forceinline__ void at_mask_chunk(float (&v)[32], int lim) {
  if (lim < 32) {
    asm volatile("" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j < lim ? v[j] : -INFINITY;
  }
}
__device__ __forceinline__ void at_max_chunk(float (&v)[32], int lim, float (&mx)[4]) {
  at_mask_chunk(v, lim);
#pragma unroll
  for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], v[j]);
}
// RS: the row sum is taken over the bf16-ROUNDED probabilities, the values the P V MMA multiplies: numerator and
// denominator then carry the same rounding (a row dominated by one key comes out exact whatever the exponent
// reference; with the row maximum as reference the dominant p is 1.0 and exact anyway).
template <bool RS>
__device__ __forceinline__ void at_exp_chunk(float (&v)[32], int lim, float scale_log2, float nm, float (&sum)[4],
                                             uint32_t prow, uint32_t ch0, uint32_t sw) {
  at_mask_chunk(v, lim);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = fast_ex2(fmaf(v[j], scale_log2, nm));
    if constexpr (!RS) sum[j & 3] += v[j];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {                               // 64 B of the row's P atom at swizzled 16-byte chunks
    const uint32_t pw[4] = {pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                            pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7])};
    if constexpr (RS) {                                       // one packed add per rounded pair
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint64_t pr = pack_f32x2(__uint_as_float(pw[q] << 16), __uint_as_float(pw[q] & 0xffff0000u));
        uint64_t acc = pack_f32x2(sum[2 * (q & 1)], sum[2 * (q & 1) + 1]);
        acc = add_f32x2(acc, pr);
        unpack_f32x2(acc, sum[2 * (q & 1)], sum[2 * (q & 1) + 1]);
      }
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (((ch0 + c) ^ sw) << 4)), "r"(pw[0]), "r"(pw[1]), "r"(pw[2]),
                 "r"(pw[3])
                 : "memory");
  }
}

// Score columns are split in two key halves X = [0, KA) and Y = [KA, KEYS) that the tensor core and the
// softmax warps hand back and forth (a full second score tile does not fit next to O in the 512 TMEM columns):
//   MMA    : S.X(t+1) is issued as soon as pass 2 has drained X(t), S.Y(t+1) after Y(t); P V(t) runs in the
//            same two instalments, so the tensor core works underneath the exponentials
//   softmax: max over X, [O(t-1) -> global], max over Y, exp/sum/P over X, exp/sum/P over Y
struct AtBars {
  uint32_t kv_full, q_full, q_free, sx_full, sy_full, px_full, py_full, o_full, o_free, v_full, k1_full, v_free, sx_full1, px_free;   // o_*: two barriers each; sx_full1: second X buffer; px_free: one-pass kernels only
};

template <int D, int SQ, bool OP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ32,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmO,
                     bf16 *__restrict__ out, int S, int h, int KEYS, int kv_box_rows, float scale_log2, int tail_mode,
                     int n_items) {
  // SQ > 0: the sequence length is a compile-time constant (the reference's 224^2 geometries: 196 + 5 and 256 + 5
  // tokens).  Key count, X / Y split, TMEM layout and the tail mode fold into constants: a third less code (the hot
  // loops of the three roles then fit the 32 KB instruction cache) and 134 instead of 168 registers; 0.54 -> 0.49 ms
  // per XL launch.  SQ = 0 keeps every shape parameter at run time.
  if constexpr (SQ > 0) {
    S = SQ;
    KEYS = (SQ + 15) & ~15;
    kv_box_rows = KEYS / ((KEYS + 255) / 256);
    tail_mode = (SQ > AT_MT && SQ - (SQ - 1) / AT_MT * AT_MT <= 8) ? 1 : 0;
  }
  constexpr int NA = D / 32;                   // 32-column (64B-swizzled) atoms of V along the head dim
  // Q, K and P are K-major MMA operands read in 32-byte k-slices: only the 128B swizzle spreads the eight rows of
  // a slice over all banks (64B-swizzled rows collide two by two), so they use 64-column atoms wherever 64
  // columns are left and a 32-column atom for the rest (d = 96: one of each).
  constexpr int N128 = D / 64, N64 = (D % 64) / 32;
  extern __shared__ uint8_t at_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP128 = KEYS / 64, NP64 = (KEYS % 64 + 31) / 32;  // P: 64-key atoms, then 32-key atoms
  int KA = (KEYS / 2) & ~63;                   // keys in the X half: whole P atoms (64 keys, else 32, else all)
  if (KA == 0) KA = (KEYS / 2) & ~31;
  if (KA == 0) KA = KEYS;
  const int KB = KEYS - KA;                    // keys in the Y half (may be 0)
  // TMEM columns.  When they fit, the X half has TWO score buffers (tile G uses buffer G & 1), so the tensor core
  // computes S.X of the next tile -- across item boundaries too -- while the softmax warps are still reading this
  // one, and a single O accumulator (the softmax warps drain O(t - 1) before they publish P.X(t), so the second one
  // bought nothing):  [X0 | X1 | Y | O]  = 128 + 128 + 144 (-> 416) + 96 = 512 columns at S = 261, d = 96.
  const int sw2 = (2 * KA + KB + 31) & ~31;
  const int XB = (KB > 0 && sw2 + D <= 512) ? 2 : 1;
  const uint32_t y_col = XB * KA;              // column of key KA
  const uint32_t o_base = XB == 2 ? sw2 : ((KEYS + 31) & ~31);
  const int OB = o_base + 2 * D <= 512 ? 2 : 1;          // O accumulators
  // tail tile P^T: compact atoms [NP128][16 rows][128 B] [NP64][16 rows][64 B].  First in the layout: the P V MMAs of
  // the tail tile are M = 128 wide and read 112 rows past each 16-row atom (their results land in TMEM lanes nobody
  // reads), which must still be shared memory of this CTA -- the Q / K tiles behind it.
  const uint32_t pt_off = 0;
  const uint32_t q_off = NP128 * 2048 + NP64 * 1024;          // [N128][128 rows][128 B] [N64][128 rows][64 B]
  const uint32_t q64_off = q_off + N128 * AT_MT * 128;
  const uint32_t k_off = q_off + NA * AT_MT * 64;             // [N128][KEYS][128 B] [N64][KEYS][64 B]
  const uint32_t k64_off = k_off + N128 * KEYS * 128;
  const uint32_t v_off = k_off + NA * KEYS * 64;              // [NA][KEYS][64 B]
  const uint32_t p_off = v_off + NA * KEYS * 64;              // [NP128][128 rows][128 B] [NP64][128 rows][64 B]
  const uint32_t p64_off = p_off + NP128 * AT_MT * 128;
  const uint32_t bar_off = p64_off + NP64 * AT_MT * 64;
  const uint32_t bar = sbase + bar_off;
  AtBars B;
  B.kv_full = bar; B.q_full = bar + 8; B.q_free = bar + 16; B.sx_full = bar + 24; B.sy_full = bar + 32;
  B.px_full = bar + 40; B.py_full = bar + 48; B.o_full = bar + 56; B.o_free = bar + 72; B.v_full = bar + 88; B.k1_full = bar + 96; B.v_free = bar + 112; B.sx_full1 = bar + 120;
  B.px_free = bar + 128;                       // (one-pass kernels: in the unused row-maximum exchange area)
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + bar_off + 104);

  // (the shuffle makes the warp index provably warp-uniform: role branches are then uniform branches and the code
  // under them may use the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  AT_TDECL
  const int C = h * D;
  // Items (image, head) are walked with a grid stride, images from the last to the first: the QKV GEMM wrote its rows
  // in ascending order (the tail of the batch is still in L2), and the output projection starts at image 0, which
  // this grid then writes last.
  const int n_img = n_items / h;
  auto item_image = [&](int item) { return n_img - 1 - item / h; };
  const int tiles = (S + AT_MT - 1) / AT_MT;
  // Per-barrier completion counts.  Per tile (global tile index G = item iteration * tiles + t): q_full, q_free,
  // sx_full, px_full; o_full / o_free alternate between the OB accumulators with G.  sy_full / py_full skip the
  // transposed tail tile (HPT completions per item).  Per item: kv_full, k1_full, v_full.

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    mbar_init(B.kv_full, 1);
    mbar_init(B.v_full, 1);
    mbar_init(B.k1_full, 1);
    mbar_init(B.v_free, 1);
    mbar_init(B.q_full, 1);
    mbar_init(B.q_free, 1);
    mbar_init(B.sx_full, 1);
    mbar_init(B.sx_full1, 1);
    if constexpr (OP) mbar_init(B.px_free, 1);
    mbar_init(B.sy_full, 1);
    mbar_init(B.px_full, 8);
    mbar_init(B.py_full, 8);
    mbar_init(B.o_full, 1);
    mbar_init(B.o_full + 8, 1);
    mbar_init(B.o_free, 8);
    mbar_init(B.o_free + 8, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t *>(tmem_slot))),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    // One loop body issues every load (the code of this warp is as hot as the softmax: it must stay small).  Tile 0 of
    // an item brings K and V along: Q and K as soon as the last score MMA of the previous item has retired, V once its
    // last P V has (its own once-per-item barrier: this thread does not follow the O-full barriers tile by tile, and a
    // parity wait that skips completions aliases).
    if (lane == 0) {
      int G = 0, itn = 0;                                     // global tile index, item count of this CTA
#pragma unroll 1
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++itn) {
        const int img = item_image(item), head = item - (item / h) * h;
        const int row0 = img * S, qcol = head * D, kcol = C + head * D, vcol = 2 * C + head * D;
#pragma unroll 1
        for (int t = 0; t < tiles; ++t, ++G) {
          AT_T(1);
          if (G > 0) at_wait(B.q_free, (G - 1) & 1);          // the scores of the previous tile have been issued and retired
          AT_T(2);
          mbar_expect_tx(B.q_full, NA * AT_MT * 64);
#pragma unroll 1
          for (int a = 0; a < N128; ++a)
            tma_load_2d(sbase + q_off + a * AT_MT * 128, &tmQ, B.q_full, qcol + 64 * a, row0 + t * AT_MT);
          if (N64) tma_load_2d(sbase + q64_off, &tmQ32, B.q_full, qcol + 64 * N128, row0 + t * AT_MT);
          if (t > 0) continue;
#pragma unroll 1
          for (int r = 0; r < KEYS; r += kv_box_rows) {       // the first box of K rows is all the X scores need
            const uint32_t kbar = r == 0 ? B.kv_full : B.k1_full;
            mbar_expect_tx(kbar, NA * kv_box_rows * 64);
#pragma unroll 1
            for (int a = 0; a < N128; ++a)
              tma_load_2d(sbase + k_off + (a * KEYS + r) * 128, &tmK, kbar, kcol + 64 * a, row0 + r);
            if (N64) tma_load_2d(sbase + k64_off + r * 64, &tmKV, kbar, kcol + 64 * N128, row0 + r);
          }
          AT_T(3);
          if (itn > 0) at_wait(B.v_free, (itn - 1) & 1);
          AT_T(4);
          mbar_expect_tx(B.v_full, NA * KEYS * 64);
#pragma unroll 1
          for (int a = 0; a < NA; ++a)
#pragma unroll 1
            for (int r = 0; r < KEYS; r += kv_box_rows)
              tma_load_2d(sbase + v_off + (a * KEYS + r) * 64, &tmKV, B.v_full, vcol + 32 * a, row0 + r);
        }
      }
      AT_DUMP("prd");
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    {
      const uint32_t leader = at_elect();
      // (shuffles: the shared-memory window and TMEM base become provably uniform, so the descriptor arithmetic below
      // runs on the uniform datapath instead of going through R2UR for every MMA)
      const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0), tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_x = at_idesc(AT_MT, KA, 0), idesc_y = at_idesc(AT_MT, KB > 0 ? KB : 16, 0);
      const uint32_t idesc_o = at_idesc(AT_MT, D, 1);
      // The MMAs here are short (48-72 tensor-core cycles), so the issue path is kept to a couple of integer adds
      // per instruction: descriptor low words (address >> 4 | LBO) are prepared once, high words are constants.
      const uint32_t q128_lo = at_lo(sb + q_off, 16), q64_lo = at_lo(sb + q64_off, 16);
      const uint32_t k128_lo = at_lo(sb + k_off, 16), k64_lo = at_lo(sb + k64_off, 16);
      const uint32_t p128_lo = at_lo(sb + p_off, 16), p64_lo = at_lo(sb + p64_off, 16);
      const uint32_t v_lo = at_lo(sb + v_off, KEYS * 64);
      // Issue loops are rolled (a dozen instructions per MMA against 48-72 tensor-core cycles each): unrolled they were
      // a fifth of the kernel's code.
      auto issue_qk = [&](uint32_t dcol, int key0, uint32_t idesc) {   // scores of keys [key0, key0 + N) -> TMEM column dcol
        const uint32_t kl128 = k128_lo + key0 * (128 >> 4), kl64 = k64_lo + key0 * (64 >> 4);
#pragma unroll
        for (int ks = 0; ks < 4 * N128; ++ks)                   // 16 columns of a 64-column atom
          at_mma(tm + dcol, q128_lo + (ks >> 2) * (AT_MT * 128 >> 4) + (ks & 3) * 2, AT_HI_SW128,
                 kl128 + (ks >> 2) * (KEYS * 128 >> 4) + (ks & 3) * 2, AT_HI_SW128, idesc, ks != 0, leader);
        if (N64) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)                        // ... of the 32-column atom
            at_mma(tm + dcol, q64_lo + ks * 2, AT_HI_SW64, kl64 + ks * 2, AT_HI_SW64, idesc, (4 * N128 + ks) != 0, leader);
        }
      };
      // 16-key steps [j0, j1) of O += P V; P from the tile's atoms (64-key, then 32-key) or, for the transposed tail
      // tile, from the compact 16-row atoms
      const uint32_t pt128_lo = at_lo(sb + pt_off, 16), pt64_lo = at_lo(sb + pt_off + NP128 * 2048, 16);
      auto issue_pv = [&](int j0, int j1, uint32_t o_col, bool tail) {
        const uint32_t a128 = tail ? pt128_lo : p128_lo, a64 = tail ? pt64_lo : p64_lo;
        const uint32_t s128 = tail ? (2048 >> 4) : (AT_MT * 128 >> 4), s64 = tail ? (1024 >> 4) : (AT_MT * 64 >> 4);
        const int jw = j1 < 4 * NP128 ? j1 : 4 * NP128;        // steps below jw read P from 64-key atoms
#pragma unroll 4
        for (int j = j0; j < jw; ++j)
          at_mma(tm + o_col, a128 + (j >> 2) * s128 + ((j & 3) << 1), AT_HI_SW128, v_lo + j * (16 * 64 >> 4), AT_HI_SW64, idesc_o,
                 j != 0, leader);
#pragma unroll 1
        for (int j = j0 > jw ? j0 : jw; j < j1; ++j) {
          const int j2 = j - 4 * NP128;
          at_mma(tm + o_col, a64 + (j2 >> 1) * s64 + ((j2 & 1) << 1), AT_HI_SW64, v_lo + j * (16 * 64 >> 4), AT_HI_SW64, idesc_o,
                 j != 0, leader);
        }
      };
      // Tail tile (tail_mode: the last query tile holds <= 8 rows, e.g. the 5 register tokens of S = 261): scores are
      // computed TRANSPOSED, S^T = K Q_tail^T (keys on the TMEM lanes, 16 query columns per 128-key tile), so the
      // softmax is spread over all lanes by key instead of 5 live lanes doing a 272-wide row each; the warps write
      // P back as rows 0..15 of compact 2 KB atoms and the usual P V follows.
      const uint32_t idesc_t = at_idesc(AT_MT, 16, 0);
      auto issue_st = [&](uint32_t dcol) {
#pragma unroll 1
        for (int kt = 0; kt * 128 < KEYS; ++kt) {
#pragma unroll
          for (int ks = 0; ks < 4 * N128; ++ks)
            at_mma(tm + dcol + 16 * kt, k128_lo + (ks >> 2) * (KEYS * 128 >> 4) + kt * (128 * 128 >> 4) + (ks & 3) * 2, AT_HI_SW128,
                   q128_lo + (ks >> 2) * (AT_MT * 128 >> 4) + (ks & 3) * 2, AT_HI_SW128, idesc_t, ks != 0, leader);
          if (N64) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              at_mma(tm + dcol + 16 * kt, k64_lo + kt * (128 * 64 >> 4) + ks * 2, AT_HI_SW64, q64_lo + ks * 2, AT_HI_SW64, idesc_t,
                     (4 * N128 + ks) != 0, leader);
          }
        }
      };
      // Scores of tile Gn.  `first`: the tile opens an item (its K arrives with kv_full / k1_full of parity kpar);
      // `tail`: transposed tail tile (reads all of K and its Q rows in one go, has no Y half).
      auto issue_scores_x = [&](int Gn, bool first, bool tail, uint32_t kpar) {
        AT_T(10);
        at_wait(B.q_full, Gn & 1);
        AT_T(11);
        if (first) at_wait(B.kv_full, kpar);
        if (tail && kv_box_rows < KEYS) at_wait(B.k1_full, kpar);
        tc_fence_after();
        const uint32_t xc = (XB == 2 && (Gn & 1)) ? KA : 0;
        if (tail) issue_st(xc);
        else issue_qk(xc, 0, idesc_x);
        at_commit((XB == 2 && (Gn & 1)) ? B.sx_full1 : B.sx_full, leader);     // (one barrier per X buffer)
        if (KB == 0 || tail) at_commit(B.q_free, leader);
      };
      auto issue_scores_y = [&](bool first, uint32_t kpar) {
        if (first && kv_box_rows < KEYS) at_wait(B.k1_full, kpar);  // the K rows past the first box
        tc_fence_after();
        issue_qk(y_col, KA, idesc_y);
        at_commit(B.sy_full, leader);
        at_commit(B.q_free, leader);
      };
      // One flat sequence of tiles over the items of this CTA.  Step G: [X scores of tile G + 1 when they have their own
      // buffer] - P.X(G) ready -> P V over the X keys [- X scores of G + 1 into the buffer just released] - P.Y(G)
      // ready -> Y scores of G + 1, P V over the Y keys.  The scores of an item's first tile are issued during the
      // previous item's last step.
      int G = 0, itn = 0, yc = 0;                             // global tile index, item count, Y halves so far
      if (blockIdx.x < n_items) {
        issue_scores_x(0, true, false, 0);
        if (KB > 0) issue_scores_y(true, 0);
      }
#pragma unroll 1
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++itn) {
        const uint32_t itp = itn & 1;
        const bool more = item + (int)gridDim.x < n_items;
#pragma unroll 1
        for (int t = 0; t < tiles; ++t, ++G) {
          const bool last = t == tiles - 1, cur_tail = tail_mode && last, next_tail = tail_mode && t + 2 == tiles;
          const bool has_next = !last || more;
          const uint32_t npar = last ? itp ^ 1 : itp;         // (the next tile opens the next item)
          if (XB == 2 && has_next) issue_scores_x(G + 1, last, next_tail, npar);
          AT_T(13);
          at_wait(B.px_full, G & 1);                          // P.X(t) in shared memory, S.X(t) consumed
          AT_T(14);
          const int ob = OB == 2 ? (G & 1) : 0;
          const uint32_t o_col = o_base + ob * D;
          if (G >= OB) at_wait(B.o_free + 8 * ob, ((OB == 2 ? (G >> 1) : G) - 1) & 1);    // O(G - OB) has been read out
          if (t == 0) at_wait(B.v_full, itp);
          tc_fence_after();
          if (cur_tail) {                                     // the whole transposed P arrives with one barrier
            issue_pv(0, KEYS / 16, o_col, true);
            at_commit(B.o_full + 8 * ob, leader);
          } else {
            issue_pv(0, KA / 16, o_col, false);
            if (KB == 0) at_commit(B.o_full + 8 * ob, leader);
          }
          // (one-pass softmax: P.X(G + 1) is written before O(G) has been waited for -- its own signal that these MMAs
          // have read P.X(G))
          if constexpr (OP) at_commit(B.px_free, leader);
          if (XB == 1 && has_next) issue_scores_x(G + 1, last, next_tail, npar);
          const bool cur_y = KB > 0 && !cur_tail;
          if (cur_y) {
            AT_T(15);
            at_wait(B.py_full, yc & 1);
            ++yc;
            AT_T(16);
          }
          // (the Y columns are free: the P.Y barrier of the last tile that had a Y half has been seen)
          if (KB > 0 && has_next && !next_tail) issue_scores_y(last, npar);
          if (cur_y) {
            tc_fence_after();
            issue_pv(KA / 16, KEYS / 16, o_col, false);
            at_commit(B.o_full + 8 * ob, leader);
          }
          if (last) {
            AT_T(19);
            at_commit(B.v_free, leader);                      // every MMA of this item has been issued: V may be replaced once they retire
          }
        }
      }
    }
    AT_DUMP("mma");
    __syncwarp();
  } else {
    // ================= softmax + epilogue =================
    // Eight warps: TMEM lane quadrant = warp % 4 (a warp only reaches its own 32 lanes), so two warps share every
    // query row; warp group g = 0/1 takes the first / second part of each key half and of the O columns.  The
    // two partial maxima and sums of a row meet in shared memory across a 64-thread named barrier.
    const int quad = warp & 3;
    const int g = (warp - 2) >> 2;
    const int r = quad * 32 + lane;                           // row inside the tile = TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    // P row of a 32-key chunk at key kk: address of the row in its atom, first 16-byte chunk, swizzle XOR
    auto p_place = [&](int kk, uint32_t &prow, uint32_t &ch0, uint32_t &sw) {
      if (kk < 64 * NP128) {
        prow = sbase + p_off + (kk >> 6) * AT_MT * 128 + r * 128;
        ch0 = (kk & 63) >> 3;
        sw = r & 7;                                           // 128B swizzle: chunk ^ address bits [7:9]
      } else {
        prow = sbase + p64_off + ((kk - 64 * NP128) >> 5) * AT_MT * 64 + r * 64;
        ch0 = 0;
        sw = (r >> 1) & 3;                                    // 64B swizzle: chunk ^ address bits [7:8]
      }
    };
    float *xch = reinterpret_cast<float *>(smem + bar_off + 128);     // [max|sum][tile parity][group][128 rows]
    auto first_part = [](int n) { return n < 64 ? (n < 32 ? n : 32) : ((n / 2) & ~31); };
    const int xa = first_part(KA), ya = first_part(KB);
    const int xk0 = g ? xa : 0, xn = g ? KA - xa : xa;        // this group's keys of the X half
    const int yk0 = KA + (g ? ya : 0), yn = g ? KB - ya : ya; // ... and of the Y half (up to 96 = three chunks)
    // Output staging: when this warp's Y keys start on a 64-key P atom of their own, its 32 rows of that atom
    // (4 KB nobody else touches) are free from the end of P V(t) until the warp's own pass 2 over Y(t + 1):
    // O goes there and leaves as one TMA store instead of 16-byte stores to 32 different rows.
    const bool stage_ok = (KA & 63) == 0 && ya == 64 && KB >= 128 && KA / 64 + 2 <= NP128;
    const uint32_t stage = sbase + p_off + (KA / 64 + g) * AT_MT * 128 + quad * 32 * 128 + lane * (D / 2) * 2;

    // (key kk of the X half sits in TMEM column xc + kk, xc = this tile's X buffer; key kk of the Y half in column
    // kk + (XB - 1) KA)
    auto pass1 = [&](int key0, int nkeys, float m, uint32_t coff) {          // row maximum of this group's keys
      float mx[4] = {m, m, m, m};
#pragma unroll 1
      for (int k = 0; k < nkeys; k += 32) {                   // one chunk per TMEM round trip: two in flight were measured
        const int n = nkeys - k < 32 ? nkeys - k : 32, kk = key0 + k;   // at 1.09 vs 0.79 ms (spills at the 168-register cap)
        float v0[32];
        at_ld_chunk(t_lane + coff + kk, n, v0);
        tmem_ld_wait();
        at_max_chunk(v0, min(n, S - kk), mx);
      }
      return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    };
    auto pass2 = [&](int key0, int nkeys, float nm, uint32_t coff) {         // p = 2^(s*scale - max*scale) -> P (bf16), row sum
      float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
      for (int k = 0; k < nkeys; k += 32) {
        const int n = nkeys - k < 32 ? nkeys - k : 32, kk = key0 + k;
        float v0[32];
        at_ld_chunk(t_lane + coff + kk, n, v0);
        tmem_ld_wait();
        uint32_t prow, ch0, sw;
        p_place(kk, prow, ch0, sw);
        at_exp_chunk<OP>(v0, min(n, S - kk), scale_log2, nm, sum, prow, ch0, sw);
      }
      return (sum[0] + sum[1]) + (sum[2] + sum[3]);
    };
    auto publish = [&](uint32_t barrier) {                    // P part visible to the tensor core, S part free
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(barrier);
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };
    // this group's half of O / row sum -> bf16 -> global, for tile G = tile t of image b, head `head`
    auto epilogue = [&](int G, int b, int head, int t, bool tail) {
      constexpr int HALF = D / 2;
      const int ob = OB == 2 ? (G & 1) : 0;
      const float *xs = xch + 512 + (G & 1) * 256;
      const float den = tail ? xs[r & 7] : xs[r] + xs[128 + r];    // (transposed tail tile: one sum per query row)
      const float inv = 1.0f / den;
      const int row = t * AT_MT + r;
      bf16 *orow = out + ((long long)b * S + row) * C + head * D + g * HALF;
      const bool warp_live = t * AT_MT + quad * 32 < S;
      if (warp_live) {
        at_wait(B.o_full + 8 * ob, (OB == 2 ? (G >> 1) : G) & 1);
        tc_fence_after();
        AT_T(40);
      }
      auto store8 = [&](const float *w, int c) {              // 8 columns -> 16 bytes
        const uint32_t p0 = pack_bf16x2(w[0] * inv, w[1] * inv), p1 = pack_bf16x2(w[2] * inv, w[3] * inv);
        const uint32_t p2 = pack_bf16x2(w[4] * inv, w[5] * inv), p3 = pack_bf16x2(w[6] * inv, w[7] * inv);
        if (stage_ok) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + 2 * c), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                       : "memory");
        } else if (row < S) {
          *reinterpret_cast<uint4 *>(orow + c) = make_uint4(p0, p1, p2, p3);
        }
      };
      constexpr int REST = HALF - 32;                         // 0, 16 or 32 columns after the first 32
      if (warp_live) {
        float v[32];
        tmem_ld32(t_lane + o_base + ob * D + g * HALF, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) store8(&v[8 * c], 8 * c);
        if constexpr (REST > 0) {
          if constexpr (REST == 32) tmem_ld32(t_lane + o_base + ob * D + g * HALF + 32, v);
          else tmem_ld16(t_lane + o_base + ob * D + g * HALF + 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < REST / 8; ++c) store8(&v[8 * c], 32 + 8 * c);
        }
        tc_fence_before();
        if (stage_ok) fence_proxy_async_smem();
        AT_T(41);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(B.o_free + 8 * ob);
        if (warp_live && stage_ok) {                          // rows past the end of the image are clipped by the map
          tma_store_3d(&tmO, stage, head * D + g * HALF, t * AT_MT + quad * 32, b);
          bulk_commit();
        }
      }
    };

    // ---- tail tile, transposed.  Scores: this thread owns key 128 kt + 32 quad + lane of key tiles kt = g, g + 2 and
    // drops its 8 query columns into an [8 queries][KEYS] fp32 table in shared memory (keys past the sequence as -inf);
    // softmax: warp w then owns query row w -- two keys per lane and 64-key step, one max and one sum reduction per
    // row -- and writes P^T (bf16 pairs) into the compact atoms.
    float *tsc = xch + 1024;
    auto tail_scores = [&](uint32_t sx_bar, uint32_t sx_par, uint32_t xc) {
      AT_T(49);
      at_wait(sx_bar, sx_par);
      tc_fence_after();
      AT_T(50);
#pragma unroll 1
      for (int kt = g; kt * 128 < KEYS; kt += 2) {
        float v[8];
        tmem_ld8(t_lane + xc + 16 * kt, v);
        tmem_ld_wait();
        const int key = 128 * kt + 32 * quad + lane;
        if (key < KEYS) {
#pragma unroll
          for (int qi = 0; qi < 8; ++qi) tsc[qi * KEYS + key] = key < S ? v[qi] : -INFINITY;
        }
      }
      AT_T(51);
    };
    auto tail_softmax = [&](int G) {
      const int qi = warp - 2;
      if (qi < S - (tiles - 1) * AT_MT) {
        const float *row = tsc + qi * KEYS;
        float x[10];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 5; ++i) {                         // KEYS <= 288 < 5 * 64
          const int k = 64 * i + 2 * lane;
          float2 p2 = make_float2(-INFINITY, -INFINITY);
          if (k < KEYS) p2 = *reinterpret_cast<const float2 *>(row + k);
          x[2 * i] = p2.x;
          x[2 * i + 1] = p2.y;
          m = fmaxf(m, fmaxf(p2.x, p2.y));
        }
        if constexpr (!OP) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        }
        const float nm = OP ? 0.0f : -m * scale_log2;
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int k = 64 * i + 2 * lane;
          const float p0 = fast_ex2(fmaf(x[2 * i], scale_log2, nm)), p1 = fast_ex2(fmaf(x[2 * i + 1], scale_log2, nm));
          if constexpr (OP) {                                 // (sum of the rounded values, as in at_exp_chunk)
            const uint32_t pw = pack_bf16x2(p0, p1);
            sum += __uint_as_float(pw << 16) + __uint_as_float(pw & 0xffff0000u);
          } else {
            sum += p0 + p1;
          }
          if (k < KEYS) {                                     // P[qi][k], P[qi][k + 1] (zeros past the sequence)
            const int k2 = k - 64 * NP128;
            const uint32_t addr = k2 < 0 ? sbase + pt_off + (k >> 6) * 2048 + qi * 128 + (((((k & 63) >> 3)) ^ (qi & 7)) << 4) + (k & 7) * 2
                                         : sbase + pt_off + NP128 * 2048 + (k2 >> 5) * 1024 + qi * 64 +
                                               (((((k2 & 31) >> 3)) ^ ((qi >> 1) & 3)) << 4) + (k2 & 7) * 2;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16x2(p0, p1)) : "memory");
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) xch[512 + (G & 1) * 256 + qi] = sum;
      }
      AT_T(54);
    };

    // One flat sequence of tiles over the items of this CTA; iteration G: scores of tile G (row maxima), output of tile
    // G - 1 -- across item boundaries too, so that nobody waits for the P V of an item's last tile (the transposed tail,
    // whose few live rows belong to the two warps of one lane quadrant) -- then the probabilities of tile G.
    int item = blockIdx.x, t = 0, b = 0, head = 0, yc = 0;    // current tile; Y halves so far
    int pb = 0, phead = 0, pt = 0;                            // previous tile
    bool ptail = false;
#pragma unroll 1
    for (int G = 0;; ++G) {
      const bool cur = item < n_items;
      if (!cur && G == 0) break;
      if (cur && t == 0) {
        b = item_image(item);
        head = item - (item / h) * h;
      }
      const bool is_tail = cur && tail_mode && t == tiles - 1;
      const bool live = t * AT_MT + quad * 32 < S;            // warps whose rows are all past the sequence idle
      float *xm = xch + (G & 1) * 256, *xs = xch + 512 + (G & 1) * 256;
      float m = -INFINITY;
      const uint32_t xc = (XB == 2 && (G & 1)) ? KA : 0;      // this tile's X score buffer
      // (one barrier per X buffer: the tensor core may commit the scores of tile G + 1 before these warps have looked
      // at the barrier of tile G -- on a single barrier that parity wait aliases and the CTA hangs)
      const uint32_t sx_bar = (XB == 2 && (G & 1)) ? B.sx_full1 : B.sx_full, sx_par = (XB == 2 ? G >> 1 : G) & 1;
      if constexpr (OP) {
        // ONE pass over the scores: the caller bounds |q k^T| / sqrt(d) (q and k come out of a LayerNorm whose scale
        // and shift are known), so p = exp(s) cannot overflow and needs no row maximum; softmax is invariant to the
        // reference.  Every score is read from TMEM once instead of twice, and nothing waits for the Y scores before
        // the X exponentials start.  Order per tile: P.X(G) - output of tile G - 1 - P.Y(G).
        const bool full = cur && !is_tail;
        float sum = 0.0f;
        bool done = false;
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          if (full && (hf == 0 || KB > 0)) {
            AT_T(20 + hf);
            at_wait(hf ? B.sy_full : sx_bar, hf ? yc & 1 : sx_par);
            tc_fence_after();
            if (hf && stage_ok) {                             // the staged O(G - 1) must have left shared memory
              if (lane == 0) bulk_wait_read<0>();
              __syncwarp();
            }
            // the P V MMAs of tile G - 1 have read P.X (a transposed tail reads its own P^T region, and the tile in
            // front of it was waited for through its O).  After an item's first tile this wait is exposed (~800 cycles
            // per item: those MMAs have themselves waited for the item's V, which loads once the previous item's last
            // P V has retired); waiting behind the first chunk's exponentials instead measured slower.
            if (hf == 0 && G > 0 && !ptail) at_wait(B.px_free, (G - 1) & 1);
            AT_T(22 + hf);
            if (live) sum += pass2(hf ? yk0 : xk0, hf ? yn : xn, 0.0f, hf ? y_col - KA : xc);
            AT_T(28 + hf);
            publish(hf ? B.py_full : B.px_full);
          }
          if (hf == 0) {
            if (is_tail) tail_scores(sx_bar, sx_par, xc);
            // partial row sums of tile G - 1: shared by the two warps of a row -- except around a transposed tail tile,
            // whose score table and row sums are shared by all eight softmax warps
            if (is_tail || ptail) {
              asm volatile("bar.sync 9, 256;" ::: "memory");
              if (ptail && stage_ok) {                        // O(G - 2) was staged in the tail iteration with no P.Y pass since
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
              }
            } else {
              pair_sync();
            }
            AT_T(26);
            if (G > 0) epilogue(G - 1, pb, phead, pt, ptail);
            AT_T(27);
            if (!full) {
              done = !cur;
              break;
            }
          }
        }
        if (done) break;
        if (is_tail) {
          tail_softmax(G);
          publish(B.px_full);
        } else {
          xs[g * 128 + r] = sum;
          if (KB > 0) ++yc;
        }
      } else {
        if (is_tail) {
          tail_scores(sx_bar, sx_par, xc);
        } else if (cur) {
  #pragma unroll 1
          for (int hf = 0; hf < (KB > 0 ? 2 : 1); ++hf) {       // X then Y: the Y scores land while X is scanned
            AT_T(20 + hf);
            at_wait(hf ? B.sy_full : sx_bar, hf ? yc & 1 : sx_par);
            tc_fence_after();
            AT_T(22 + hf);
            if (live) m = pass1(hf ? yk0 : xk0, hf ? yn : xn, m, hf ? y_col - KA : xc);
            AT_T(24 + hf);
          }
          xm[g * 128 + r] = m;
        }
        // partial maxima of tile G, partial sums of tile G - 1: shared by the two warps of a row -- except around a
        // transposed tail tile, whose score table and row sums are shared by all eight softmax warps
        if (is_tail || ptail) {
          asm volatile("bar.sync 9, 256;" ::: "memory");
          if (ptail && stage_ok) {                              // O(G - 2) was staged in the tail iteration with no P pass since:
            if (lane == 0) bulk_wait_read<0>();                 // its TMA store must have read the staging rows
            __syncwarp();
          }
        } else {
          pair_sync();
        }
        AT_T(26);
        if (G > 0) epilogue(G - 1, pb, phead, pt, ptail);
        AT_T(27);
        if (!cur) break;
        if (is_tail) {
          tail_softmax(G);
          publish(B.px_full);
          AT_T(55);
        } else {
          m = fmaxf(m, xm[(g ^ 1) * 128 + r]);
          const float nm = -m * scale_log2;
          float sum = 0.0f;
  #pragma unroll 1
          for (int hf = 0; hf < (KB > 0 ? 2 : 1); ++hf) {
            if (hf && stage_ok) {                               // the staged O(G - 1) must have left shared memory
              if (lane == 0) bulk_wait_read<0>();
              __syncwarp();
            }
            if (live) sum += pass2(hf ? yk0 : xk0, hf ? yn : xn, nm, hf ? y_col - KA : xc);
            AT_T(28 + hf);
            publish(hf ? B.py_full : B.px_full);
            AT_T(30 + hf);
          }
          xs[g * 128 + r] = sum;
          if (KB > 0) ++yc;
        }
      }
      pb = b;
      phead = head;
      pt = t;
      ptail = is_tail;
      if (++t == tiles) {
        t = 0;
        item += gridDim.x;
      }
    }
    if (stage_ok && lane == 0) bulk_wait_read<0>();          // shared memory outlives the last output store
    AT_DUMP("smx");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int D, int SQ, bool OP>
static int launch_attn_tc5(const void *qkv, void *out, int B, int S, int h, cudaStream_t st) {
  const int C = h * D;
  const int KEYS = (S + 15) & ~15;
  const int nbox = (KEYS + 255) / 256;
  if (KEYS > 288 || KEYS % nbox != 0 || (KEYS / nbox) % 8 != 0) return -1;
  const int NA = D / 32, NKA = (KEYS + 31) / 32;
  const int NP128 = KEYS / 64, NP64 = (KEYS % 64 + 31) / 32;
  // the last query tile is transposed when it holds at most 8 rows
  const int tiles = (S + AT_MT - 1) / AT_MT;
  const int tail_mode = tiles >= 2 && S - (tiles - 1) * AT_MT <= 8 ? 1 : 0;
  // Q, K + V, P, tail P^T, barriers, row max / sum exchange, (tail) score table, alignment slack
  const size_t smem = (size_t)NA * AT_MT * 64 + 2 * (size_t)NA * KEYS * 64 + (size_t)NKA * AT_MT * 64 +
                      (size_t)NP128 * 2048 + (size_t)NP64 * 1024 + 128 + 4096 + (tail_mode ? 8 * (size_t)KEYS * 4 : 0) + 1024;
  if (smem > 227 * 1024) return -1;
  CUtensorMap tq, tq32, tk, tkv;          // 64-column boxes are 128B-swizzled, 32-column boxes 64B-swizzled
  if (int rc = make_tensor_map_bf16(qkv, (uint64_t)B * S, 3 * C, 3 * C, AT_MT, 64, &tq)) return rc;
  if (int rc = make_tensor_map_bf16(qkv, (uint64_t)B * S, 3 * C, 3 * C, AT_MT, 32, &tq32)) return rc;
  if (int rc = make_tensor_map_bf16(qkv, (uint64_t)B * S, 3 * C, 3 * C, KEYS / nbox, 64, &tk)) return rc;
  if (int rc = make_tensor_map_bf16(qkv, (uint64_t)B * S, 3 * C, 3 * C, KEYS / nbox, 32, &tkv)) return rc;
  CUtensorMap to;                          // [B][S][C] output, 32-row x D/2-column boxes
  if (int rc = make_tensor_map_bf16_3d(out, B, S, C, 32, D / 2, &to)) return rc;
  auto kern = attention_tc5_kernel<D, SQ, OP>;
  SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device, cheap
  int dev = 0, sms = 0;
  SDP_CUDA(cudaGetDevice(&dev));
  SDP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  const int n_items = B * h;               // persistent: one CTA per SM (512 TMEM columns, ~210 KB of shared memory)
  kern<<<n_items < sms ? n_items : sms, AT_THREADS, smem, st>>>(tq, tq32, tk, tkv, to, (bf16 *)out, S, h, KEYS, KEYS / nbox,
                                                                scale_log2, tail_mode, n_items);
  SDP_LAUNCH_OK();
  return 0;
}

template <int D, int SQ>
static int launch_attn_tc5_sq(const void *qkv, void *out, int B, int S, int h, bool one_pass, cudaStream_t st) {
  return one_pass ? launch_attn_tc5<D, SQ, true>(qkv, out, B, S, h, st) : launch_attn_tc5<D, SQ, false>(qkv, out, B, S, h, st);
}

// -1: shape not covered (the caller falls back to the mma.sync kernel).  score_bound > 0: the caller guarantees
// |q k^T| / sqrt(d) <= score_bound (nats) for every query / key pair; up to AT_ONE_PASS_BOUND the one-pass kernel runs
// (p = exp(s) stays inside [e^-60, e^60], its row sums and P V inside fp32 with 38 decades to spare).
int attention_tc5(const void *qkv, void *out, int B, int S, int h, int d, float score_bound, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (h * d) % 8) return -1;
  const bool op = score_bound > 0.0f && score_bound <= AT_ONE_PASS_BOUND;
  switch (d) {
    case 64: return S == 201 ? launch_attn_tc5_sq<64, 201>(qkv, out, B, S, h, op, st) : launch_attn_tc5_sq<64, 0>(qkv, out, B, S, h, op, st);
    case 96:
      return S == 261   ? launch_attn_tc5_sq<96, 261>(qkv, out, B, S, h, op, st)
             : S == 201 ? launch_attn_tc5_sq<96, 201>(qkv, out, B, S, h, op, st)
                        : launch_attn_tc5_sq<96, 0>(qkv, out, B, S, h, op, st);
    case 128: return launch_attn_tc5_sq<128, 0>(qkv, out, B, S, h, op, st);
    default: return -1;
  }
}

}  // namespace sdp
