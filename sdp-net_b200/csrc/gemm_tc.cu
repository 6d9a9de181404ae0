// bf16 GEMM on the 5th-gen tensor cores:  out = epi(A[M,K] . W[N,K]^T), fp32 accumulate in TMEM.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, STAGES-deep ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma 128 x BLOCK_N x 16; owns TMEM alloc)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> bias/act/residual -> global), double-buffered
//                              TMEM accumulators so tile i's epilogue overlaps tile i+1's MMAs.
// Replaces every nn.Linear / 1x1 conv of the reference forward (see include/sdpnet_b200.h).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

// this kernel only runs on bf16 operands and (bar the tiny head GEMMs) writes bf16: the degree-3 GELU is enough
#define SDP_GELU_FAST_FN gelu_erf_fast3
#define SDP_GELU_FAST_X2          // ... evaluated on packed fp32 pairs (FFMA2), bit-identical to the scalar form
#include "tc5.cuh"

namespace sdp {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;      // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;          // default: warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (320 threads)
constexpr int EPI_WARPS_WIDE = 16;    // heavy epilogues (GELU): four warps per TMEM lane quadrant (576 threads)

// RSM (in-place residual epilogue): slots per epilogue warp, each [32 rows x 64 B] per plane (hi, or hi + lo)
__host__ __device__ constexpr int rsm_slots(int cg) { return cg == 2 ? 3 : 2; }
__host__ __device__ constexpr int out_bytes_for(bool lean, int rsm, int cg) {
  return lean ? 65536 : rsm > 0 ? EPI_WARPS * rsm_slots(cg) * 2048 * rsm : 32768;
}

template <int BN, int STAGES, int CG = 1, int OUTB = 32768>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BN / CG) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_OFF = STAGES * STAGE_BYTES;              // epilogue staging: 2 KB [32 rows x 64 B] slots
  static constexpr int OUT_BYTES = OUTB;                            // 8 warps x 2 (32 KB), 16 warps x 2 (64 KB), or the RSM slots
  static constexpr int BAR_OFF = OUT_OFF + OUT_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;   // barriers + slack for 1024B alignment
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// pipeline depth that fits next to `out_bytes` of epilogue staging
__host__ __device__ constexpr int stages_for(int bn, int cg, int out_bytes, int want) {
  const int stage = BLOCK_M * BLOCK_K * 2 + (bn / cg) * BLOCK_K * 2;
  const int fit = (232448 - 512 - 1024 - out_bytes) / stage;
  return fit < want ? fit : want;
}

__host__ __device__ constexpr uint32_t tmem_cols_for(int bn) {
  return 2 * bn <= 32 ? 32u : 2 * bn <= 64 ? 64u : 2 * bn <= 128 ? 128u : 2 * bn <= 256 ? 256u : 512u;
}

// STAGED: bf16 output tiles leave through shared memory and TMA stores (one 32 x 32 box per warp and
// 32-column chunk, 64B-swizzled, double-buffered per warp) instead of 16-byte-per-row global stores.
// CG = 1: one CTA per 128 x BN tile.  CG = 2: launched as clusters of 2 (one TPC); the pair owns a
// 256 x BN tile, B traffic from L2 halves and the smem ring gets deeper for the same capacity.
// LEAN (16 epilogue warps, four per scheduler): bias + activation (LEAN = 1) or the folded LayerNorm + activation
// (LEAN = 2: v = rstd * (acc - mean * s[c]) + t[c], row statistics from the producer GEMM), N a multiple of BN,
// every chunk staged and TMA-stored; none of the residual / pass-through / statistics / head-norm paths are
// compiled in, which is what lets the epilogue fit the 96 registers a 18-warp CTA leaves per thread.
// RSM (in-place residual GEMMs, out == residual: the four GEMMs per block that write the residual stream): the
// residual tile travels like the operands -- each epilogue warp TMA-loads the [32 x 32] box it is going to
// overwrite into its own shared-memory slot one to two chunks ahead (across tile boundaries), adds the accumulator
// in place, and TMA-stores the slot back.  No residual registers, no exposed global-load latency.  RSM = 2: the
// stream is split in two bf16 planes (hi = bf16(v), lo = bf16(v - hi), ~16 mantissa bits), both planes ride in
// the slot; the consumers' A operand is the hi plane.  Pass-through rows (register tokens in the mixers) simply
// leave their slot rows untouched.
template <int BN, int STAGES, int ACT, int HN, bool STAGED, int CG = 1, int EW = EPI_WARPS, int LEAN = 0, int RSM = 0>
__global__ void __launch_bounds__((2 + EW) * 32, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                    const Epilogue epi, const int K, const int flags) {
  // flags: bit 0 = 16-byte epilogue accesses are legal; bit 2 = clusters of two CTA pairs that share the weight tile
  // (the pairs take neighbouring 256-row tiles of the same column block; each CTA loads a quarter of the block and
  // multicasts it to its counterpart in the other pair: a quarter less L2 -> SM traffic, which is what bounds the GEMMs)
  const int vec_ok = flags & 1;
  const int mc = (CG == 2 && (flags & 4)) ? 2 : 1;
#ifdef SDP_DIAG
  const bool nofeed = (flags & 2) != 0;   // diagnostic build only: no TMA loads, MMAs run on stale shared memory
#else
  constexpr bool nofeed = false;
#endif
  using L = SmemLayout<BN, STAGES, CG, out_bytes_for(LEAN != 0, RSM, CG)>;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const uint32_t rank = crank & 1u;                 // position inside the CTA pair (0 = leader, issues the MMAs)
  const uint32_t pairid = crank >> 1, leader = crank & ~1u;
  extern __shared__ uint8_t smem_raw[];
  // swizzle-128B tiles need 1024-byte alignment
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + L::BAR_OFF + 8 * (2 * STAGES + 4));
  constexpr int NSLOT = rsm_slots(CG);
  auto res_bar = [&](int ew, int sl) { return bar_base + 8u * (2 * STAGES + 5 + ew * NSLOT + sl); };   // RSM only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BN);   // power of two >= 2 * BN

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (STAGED) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    if (RSM == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO2) : "memory");
    if constexpr (RSM > 0)
      for (int i = 0; i < EW * NSLOT; ++i) mbar_init(res_bar(0, i), 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), mc);           // one commit per pair that reads this slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CG * EW);   // one arrive per epilogue warp (of both CTAs when paired)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t *>(tmem_slot))),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {     // issued by the same warp of BOTH CTAs of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t *>(tmem_slot))),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr int TILE_M = BLOCK_M * CG;
  const int UNIT_M = TILE_M * mc;                    // rows one cluster covers per tile
  const int m_tiles = (epi.M + UNIT_M - 1) / UNIT_M;
  const int n_tiles = (epi.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int unit0 = blockIdx.x / (CG * mc), unit_stride = gridDim.x / (CG * mc);
  const int row_off = pairid * TILE_M + rank * BLOCK_M;   // this CTA's rows inside the cluster's tile

  if (warp == 0) {
    // ================= TMA producer =================
    // (the whole warp walks the loop so that the stage / phase counters stay warp-uniform; lane 0 issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
        const int m0 = (tile / n_tiles) * UNIT_M + row_off;
        const int n0 = (tile % n_tiles) * BN + rank * (BN / CG);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          if (lane == 0) {
            if (nofeed) {
              if (rank == 0) mbar_arrive(full_bar(stage));
            } else if constexpr (CG == 1) {
              mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
              tma_load_2d(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0);
              tma_load_2d(sa + L::A_BYTES, &tmW, full_bar(stage), kb * BLOCK_K, n0);
            } else {
              // both CTAs' bytes complete on the LEADER's full barrier; only the leader arms it
              if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
              tma_load_2d_2sm(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0);
              if (mc == 1) {
                tma_load_2d_2sm(sa + L::A_BYTES, &tmW, full_bar(stage), kb * BLOCK_K, n0);
              } else {       // our quarter of the column block, to us and to the same-rank CTA of the other pair
                tma_load_2d_2sm_mc(sa + L::A_BYTES + pairid * (L::B_BYTES / 2), &tmW, full_bar(stage), kb * BLOCK_K,
                                   n0 + pairid * (BN / 4), static_cast<uint16_t>(5u << rank));
              }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the loops (stage / phase / tile counters stay warp-uniform, so descriptors are built on
    // the uniform datapath); one lane issues.  Descriptor low words carry the address, high words are constant.
    constexpr uint32_t idesc = make_idesc(TILE_M, BN);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, SWIZZLE_128B
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (rank == 0) {
      for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);         // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint32_t a_lo = ((sa & 0x3FFFF) >> 4) | (1u << 16);
          const uint32_t b_lo = (((sa + L::A_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 32 bytes (16 bf16) inside the 128B swizzle atom: +2 in the >>4 address field
              if constexpr (CG == 1) tc_mma_bf16_w(d_tmem, a_lo + 2 * k, b_lo + 2 * k, DESC_HI, idesc, (kb | k) != 0);
              else tc_mma_bf16_2sm_w(d_tmem, a_lo + 2 * k, b_lo + 2 * k, DESC_HI, idesc, (kb | k) != 0);
            }
            if constexpr (CG == 1) {
              tc_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
              if (kb == num_kb - 1) tc_commit(tfull_bar(acc));
            } else {                                       // same barriers in both CTAs of the pair
              tc_commit_2sm(empty_bar(stage), mc == 2 ? 15 : 3);         // every CTA whose slot these MMAs read or feed
              if (kb == num_kb - 1) tc_commit_2sm(tfull_bar(acc), static_cast<uint16_t>(3u << leader));
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if constexpr (LEAN != 0) {
    // ================= lean epilogue: bias (or folded LayerNorm) + activation, four warps per TMEM lane quadrant =================
    static_assert(LEAN == 0 || (EW == 16 && HN == 0 && STAGED && BN % 128 == 0 && RSM == 0), "lean epilogue: 16 warps, staged, no head-norm");
    constexpr int WCOLS = BN / 4;                  // columns per warp
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const uint32_t stage_base = smem_base + L::OUT_OFF + static_cast<uint32_t>(warp - 2) * 4096;   // two 2 KB slots
    const uint32_t rowp = lane * 64, sw = (lane >> 1) & 3;
    uint32_t cc = 0;
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * UNIT_M + row_off;
      const int n0 = (tile % n_tiles) * BN + part * WCOLS;
      uint64_t nmean2 = 0ull, rstd2 = 0ull;
      if constexpr (LEAN == 2) {                   // this row's (mean, rstd) from the producer's column-part sums
        const int row = m0 + quad * 32 + lane;
        float s1 = 0.0f, s2 = 0.0f;
        if (row < epi.M) {
          const float4 *sp = reinterpret_cast<const float4 *>(epi.ln_stats + (long long)row * epi.ln_parts * 2);
          for (int p2 = 0; p2 < epi.ln_parts / 2; ++p2) {
            const float4 t = __ldg(sp + p2);
            s1 += t.x + t.z;
            s2 += t.y + t.w;
          }
        }
        const float inv = 1.0f / (float)epi.ln_K;
        const float mean = s1 * inv;
        const float rstd = rsqrtf(fmaxf(s2 * inv - mean * mean, 0.0f) + epi.ln_eps);
        nmean2 = pack_f32x2(-mean, -mean);
        rstd2 = pack_f32x2(rstd, rstd);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + part * WCOLS;
#pragma unroll 1
      for (int c = 0; c < WCOLS; c += 32) {
        float v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (c + 32 == WCOLS) {                     // this warp's share of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cta(tempty_bar(acc), leader);
          }
        }
        const int col = n0 + c;
        if constexpr (LEAN == 2) {                 // v = rstd * (acc - mean * s[c]) + t[c]: two packed FMAs per pair
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 sv = __ldg(reinterpret_cast<const float4 *>(epi.ln_s + col + j));
            const float4 tv = __ldg(reinterpret_cast<const float4 *>(epi.ln_t + col + j));
            const uint64_t a0 = fma_f32x2(pack_f32x2(sv.x, sv.y), nmean2, pack_f32x2(v[j], v[j + 1]));
            const uint64_t a1 = fma_f32x2(pack_f32x2(sv.z, sv.w), nmean2, pack_f32x2(v[j + 2], v[j + 3]));
            unpack_f32x2(fma_f32x2(a0, rstd2, pack_f32x2(tv.x, tv.y)), v[j], v[j + 1]);
            unpack_f32x2(fma_f32x2(a1, rstd2, pack_f32x2(tv.z, tv.w)), v[j + 2], v[j + 3]);
          }
        } else if (epi.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(epi.bias + col + j));
            unpack_f32x2(add_f32x2(pack_f32x2(v[j], v[j + 1]), pack_f32x2(b.x, b.y)), v[j], v[j + 1]);
            unpack_f32x2(add_f32x2(pack_f32x2(v[j + 2], v[j + 3]), pack_f32x2(b.z, b.w)), v[j + 2], v[j + 3]);
          }
        }
        apply_act_vec<32, false>(v, ACT < 0 ? epi.act : ACT);
        const uint32_t slot = stage_base + ((cc & 1) << 11);
        if (cc >= 2) {                             // the store issued from this slot two chunks ago has read it
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t p0 = pack_bf16x2(v[8 * j], v[8 * j + 1]), p1 = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          const uint32_t p2 = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), p3 = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + rowp + ((j ^ sw) << 4)), "r"(p0), "r"(p1),
                       "r"(p2), "r"(p3)
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, slot, col, m0 + quad * 32);
          bulk_commit();
        }
        ++cc;
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  } else if constexpr (RSM > 0) {
    // ================= in-place residual epilogue (see the kernel comment) =================
    static_assert(RSM == 0 || (EW == 8 && HN == 0 && STAGED && BN % 64 == 0), "RSM epilogue: 8 warps, staged, no head-norm");
    constexpr uint32_t SLOT_BYTES = 2048u * RSM;             // hi plane [32 x 64 B] (+ lo plane behind it)
    constexpr int HCOLS = BN / 2;                            // columns per warp (two warps per TMEM lane quadrant)
    constexpr int CPW = HCOLS / 32;                          // 32-column chunks per warp and tile
    const int quad = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;
    const uint32_t slot_base = smem_base + L::OUT_OFF + static_cast<uint32_t>(ew) * (NSLOT * SLOT_BYTES);
    const uint32_t rowoff = lane * 64, sw = (lane >> 1) & 3;
    // prefetch cursor: walks this warp's chunks (tile-major, then 32-column chunks, columns past N skipped) ahead
    // of the compute loop; both enumerate the same sequence, so chunk number i always lives in slot i % NSLOT
    int pf_tile = unit0, pf_k = 0, issued = 0;
    auto pf_col = [&]() { return (pf_tile % n_tiles) * BN + half * HCOLS + pf_k * 32; };
    auto pf_skip = [&]() {                                   // settle on the next chunk that has columns < N
      while (pf_tile < num_tiles && pf_col() >= epi.N) {
        if (++pf_k == CPW) { pf_k = 0; pf_tile += unit_stride; }
      }
    };
    auto issue_load = [&]() {                                // warp-uniform bookkeeping, lane 0 issues
      pf_skip();
      if (pf_tile < num_tiles) {
        const int sl = issued % NSLOT;
        if (lane == 0) {
          const uint32_t dst = slot_base + sl * SLOT_BYTES, bar = res_bar(ew, sl);
          const int col = pf_col(), row = (pf_tile / n_tiles) * UNIT_M + row_off + quad * 32;
          mbar_expect_tx(bar, SLOT_BYTES);
          tma_load_2d(dst, &tmO, bar, col, row);
          if constexpr (RSM == 2) tma_load_2d(dst + 2048, &tmO2, bar, col, row);
        }
        ++issued;
        if (++pf_k == CPW) { pf_k = 0; pf_tile += unit_stride; }
      }
    };
#pragma unroll
    for (int d = 0; d < NSLOT - 1; ++d) issue_load();
    int done = 0, it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * UNIT_M + row_off;
      const int n0 = (tile % n_tiles) * BN + half * HCOLS;
      const RowMap rm = map_row(epi, m0 + quad * 32 + lane);
      const int nchunks = n0 >= epi.N ? 0 : (epi.N - n0 >= HCOLS ? CPW : (epi.N - n0 + 31) / 32);
      uint64_t s1p = 0ull, s2p = 0ull;                       // producer statistics of the hi values this row stores
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * HCOLS;
      if (nchunks == 0) {                                    // nothing of this tile is ours: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cta(tempty_bar(acc), leader);
        }
      }
#pragma unroll 1
      for (int k = 0; k < nchunks; ++k, ++done) {
        const int sl = done % NSLOT;
        const uint32_t slot = slot_base + sl * SLOT_BYTES;
        float v[32];
        tmem_ld32(taddr + k * 32, v);
        mbar_wait(res_bar(ew, sl), (done / NSLOT) & 1);      // this chunk's residual box has landed
        tmem_ld_wait();
        if (k == nchunks - 1) {                              // this warp's share of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cta(tempty_bar(acc), leader);
          }
        }
        const int col = n0 + k * 32;
        if (rm.live) {
          if (epi.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4 *>(epi.bias + col + j));
              unpack_f32x2(add_f32x2(pack_f32x2(v[j], v[j + 1]), pack_f32x2(b.x, b.y)), v[j], v[j + 1]);
              unpack_f32x2(add_f32x2(pack_f32x2(v[j + 2], v[j + 3]), pack_f32x2(b.z, b.w)), v[j + 2], v[j + 3]);
            }
          }
          apply_act_vec<32, false>(v, ACT < 0 ? epi.act : ACT);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = slot + rowoff + ((j ^ sw) << 4);
            uint32_t h[4], l[4] = {0u, 0u, 0u, 0u};
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(a));
            if constexpr (RSM == 2)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]) : "r"(a + 2048));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // residual = hi + lo (exact in fp32), then v + residual; bf16 halves widen by a shift / mask
              uint64_t r = pack_f32x2(__uint_as_float(h[q] << 16), __uint_as_float(h[q] & 0xffff0000u));
              if constexpr (RSM == 2)
                r = add_f32x2(r, pack_f32x2(__uint_as_float(l[q] << 16), __uint_as_float(l[q] & 0xffff0000u)));
              const uint64_t x = add_f32x2(pack_f32x2(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]), r);
              float x0, x1;
              unpack_f32x2(x, x0, x1);
              h[q] = pack_bf16x2(x0, x1);
              const uint64_t hf = pack_f32x2(__uint_as_float(h[q] << 16), __uint_as_float(h[q] & 0xffff0000u));
              if constexpr (RSM == 2) {                      // what the rounding dropped: x - hi (exact), rounded to bf16
                float d0, d1;
                unpack_f32x2(fma_f32x2(hf, pack_f32x2(-1.0f, -1.0f), x), d0, d1);
                l[q] = pack_bf16x2(d0, d1);
              }
              s1p = add_f32x2(s1p, hf);
              s2p = fma_f32x2(hf, hf, s2p);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
            if constexpr (RSM == 2)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        // the box that last left from the slot we are about to refill has been read (it was committed a whole
        // chunk ago); then: next residual box in, this chunk out
        if (lane == 0) bulk_wait_read<0>();
        issue_load();
        if (lane == 0) {
          tma_store_2d(&tmO, slot, col, m0 + quad * 32);
          if constexpr (RSM == 2) tma_store_2d(&tmO2, slot + 2048, col, m0 + quad * 32);
          bulk_commit();
        }
      }
      if (epi.stats_out != nullptr && rm.live) {          // a column half past N still owns its (zero) part
        float a0, a1, b0, b1;
        unpack_f32x2(s1p, a0, a1);
        unpack_f32x2(s2p, b0, b1);
        const int part = (tile % n_tiles) * 2 + half;
        *reinterpret_cast<float2 *>(epi.stats_out + (rm.ro * epi.stats_parts + part) * 2) = make_float2(a0 + a1, b0 + b1);
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  } else {
    // ================= epilogue =================
    // Two warps per TMEM lane quadrant: warps 2..5 take the low half of the tile's columns,
    // warps 6..9 the high half.  One thread = one output row.
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which share of the tile's columns (EW / 4 shares)
    constexpr int PARTS = EW / 4;
    constexpr int UNIT = HN > 0 ? HN : 32;     // columns finished together (whole heads with head-norm)
    constexpr int UNITS = BN / UNIT;
    static_assert(PARTS == 2 || (HN == 0 && UNITS % PARTS == 0), "wide epilogue: plain 32-column units only");
    const int u_begin = (half * UNITS + PARTS - 1) / PARTS, u_end = ((half + 1) * UNITS + PARTS - 1) / PARTS;
    constexpr int SLOTS = 16 / EW;             // 2 KB staging slots per warp (8 warps: double-buffered)
    const uint32_t stage_base = smem_base + L::OUT_OFF + static_cast<uint32_t>(warp - 2) * (SLOTS * 2048);
    uint32_t cc = 0;                           // chunks this warp has staged (slot = cc % SLOTS)
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * UNIT_M + row_off;
      const int n0 = (tile % n_tiles) * BN;
      const RowMap rm = map_row(epi, m0 + quad * 32 + lane);
      // LayerNorm folded into this GEMM: the row's (mean, rstd) from the producer's column-part sums
      const bool lnf = epi.ln_stats != nullptr;
      float ln_mean = 0.0f, ln_rstd = 1.0f;
      if (lnf && rm.live) {
        const float4 *sp = reinterpret_cast<const float4 *>(epi.ln_stats + (long long)(m0 + quad * 32 + lane) * epi.ln_parts * 2);
        float s1 = 0.0f, s2 = 0.0f;
        for (int p2 = 0; p2 < epi.ln_parts / 2; ++p2) {
          const float4 t = __ldg(sp + p2);
          s1 += t.x + t.z;
          s2 += t.y + t.w;
        }
        const float inv = 1.0f / (float)epi.ln_K;
        ln_mean = s1 * inv;
        ln_rstd = rsqrtf(fmaxf(s2 * inv - ln_mean * ln_mean, 0.0f) + epi.ln_eps);
      }
      auto ln_fold = [&](float *v, int col) {      // v = rstd * (acc - mean * s[col..]) + t[col..], 32 columns
        const uint64_t nmean2 = pack_f32x2(-ln_mean, -ln_mean), rstd2 = pack_f32x2(ln_rstd, ln_rstd);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {          // two packed FMAs per pair
          const float4 sv = __ldg(reinterpret_cast<const float4 *>(epi.ln_s + col + j));
          const float4 tv = __ldg(reinterpret_cast<const float4 *>(epi.ln_t + col + j));
          const uint64_t a0 = fma_f32x2(pack_f32x2(sv.x, sv.y), nmean2, pack_f32x2(v[j], v[j + 1]));
          const uint64_t a1 = fma_f32x2(pack_f32x2(sv.z, sv.w), nmean2, pack_f32x2(v[j + 2], v[j + 3]));
          unpack_f32x2(fma_f32x2(a0, rstd2, pack_f32x2(tv.x, tv.y)), v[j], v[j + 1]);
          unpack_f32x2(fma_f32x2(a1, rstd2, pack_f32x2(tv.z, tv.w)), v[j + 2], v[j + 3]);
        }
      };
      float st1 = 0.0f, st2 = 0.0f;              // producer side: sums of the bf16 values this thread stores
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int u = u_begin; u < u_end; ++u) {
        const int c = u * UNIT;
        const int col0 = n0 + c;
        if (col0 >= epi.N) break;
        float mean = 0.0f, rstd = 1.0f;
        const float *hw = nullptr, *hb = nullptr;
        if constexpr (HN > 0) {
          // per-head LayerNorm over this thread's HN accumulator columns: TMEM is read twice (statistics, then
          // normalise) instead of holding HN values in registers
          if (col0 < 2 * epi.hn_C) {
            // one pass: sums of (v - shift) and (v - shift)^2 with the head's first column as the shift (no
            // cancellation even when |mean| >> std), four accumulators each so the adds are not one serial chain
            // (packed fp32 pairs: accumulator j & 3 is lane j & 1 of pair (j >> 1) & 1, the same sums as four scalars)
            uint64_t s1p[2] = {0ull, 0ull}, s2p[2] = {0ull, 0ull};
            float shift = 0.0f;
#pragma unroll 1
            for (int i = 0; i < HN; i += 32) {
              float v[32];
              tmem_ld32(taddr + c + i, v);
              tmem_ld_wait();
              if (lnf) {                           // folded once: the normalise pass reads the folded values back
                ln_fold(v, col0 + i);
                tmem_st32(taddr + c + i, v);
              }
              if (i == 0) shift = v[0];
              const uint64_t nshift = pack_f32x2(-shift, -shift);
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint64_t dlt = add_f32x2(pack_f32x2(v[j], v[j + 1]), nshift);
                s1p[(j >> 1) & 1] = add_f32x2(s1p[(j >> 1) & 1], dlt);
                s2p[(j >> 1) & 1] = fma_f32x2(dlt, dlt, s2p[(j >> 1) & 1]);
              }
            }
            float s1[4], s2[4];
            unpack_f32x2(s1p[0], s1[0], s1[1]);
            unpack_f32x2(s1p[1], s1[2], s1[3]);
            unpack_f32x2(s2p[0], s2[0], s2[1]);
            unpack_f32x2(s2p[1], s2[2], s2[3]);
            if (lnf) tmem_st_wait();
            const float dm = ((s1[0] + s1[1]) + (s1[2] + s1[3])) * (1.0f / HN);
            mean = shift + dm;
            const float q = fmaxf(((s2[0] + s2[1]) + (s2[2] + s2[3])) * (1.0f / HN) - dm * dm, 0.0f) * HN;
            rstd = rsqrtf(q * (1.0f / HN) + epi.hn_eps);
            const bool is_q = col0 < epi.hn_C;
            hw = is_q ? epi.hn_qw : epi.hn_kw;
            hb = is_q ? epi.hn_qb : epi.hn_kb;
          }
        }
#pragma unroll 1
        for (int i = 0; i < UNIT; i += 32) {
          float v[32];
          tmem_ld32(taddr + c + i, v);
          tmem_ld_wait();
          if (lnf && hw == nullptr && col0 + i + 32 <= epi.N) ln_fold(v, col0 + i);   // (head-norm columns: folded above)
          if constexpr (HN > 0) {
            if (hw != nullptr) {
              const uint64_t nmean = pack_f32x2(-mean, -mean), rstd2 = pack_f32x2(rstd, rstd);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {   // ((v - mean) * rstd) * w + b on packed pairs: 1.5 slots per value
                const float4 wv = __ldg(reinterpret_cast<const float4 *>(hw + i + j));
                const float4 bv = __ldg(reinterpret_cast<const float4 *>(hb + i + j));
                const uint64_t a0 = mul_f32x2(add_f32x2(pack_f32x2(v[j], v[j + 1]), nmean), rstd2);
                const uint64_t a1 = mul_f32x2(add_f32x2(pack_f32x2(v[j + 2], v[j + 3]), nmean), rstd2);
                unpack_f32x2(fma_f32x2(a0, pack_f32x2(wv.x, wv.y), pack_f32x2(bv.x, bv.y)), v[j], v[j + 1]);
                unpack_f32x2(fma_f32x2(a1, pack_f32x2(wv.z, wv.w), pack_f32x2(bv.z, bv.w)), v[j + 2], v[j + 3]);
              }
            }
          }
          if (col0 + i < epi.N) {
            if constexpr (!STAGED) {
              epilogue_row<32, false, ACT>(epi, rm, col0 + i, v, vec_ok != 0);
            } else {
              const int col = col0 + i;
              if (rm.live) {
                epilogue_math<32, false, ACT>(epi, rm, col, v, vec_ok != 0);
              } else {
                epilogue_passthrough<32>(epi, rm, col, v, vec_ok != 0);
              }
              const uint32_t slot = stage_base + ((cc & (SLOTS - 1)) << 11);
              if (cc >= SLOTS) {                 // the store last issued from this slot has read it
                if (lane == 0) bulk_wait_read<SLOTS - 1>();
                __syncwarp();
              }
              const uint32_t rowp = slot + lane * 64;
              const uint32_t sw = (lane >> 1) & 3;   // 64-byte swizzle: 16B chunk index ^ address bits [7:8]
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t p0 = pack_bf16x2(v[8 * j], v[8 * j + 1]), p1 = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                const uint32_t p2 = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), p3 = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((j ^ sw) << 4)), "r"(p0), "r"(p1),
                             "r"(p2), "r"(p3)
                             : "memory");
                if (epi.stats_out != nullptr) {  // statistics of exactly the values consumers will read
                  const uint32_t pw[4] = {p0, p1, p2, p3};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float lo = __uint_as_float(pw[q] << 16), hi = __uint_as_float(pw[q] & 0xffff0000u);
                    const bool in_lo = col + 8 * j + 2 * q < epi.N, in_hi = col + 8 * j + 2 * q + 1 < epi.N;
                    if (in_lo) { st1 += lo; st2 = fmaf(lo, lo, st2); }
                    if (in_hi) { st1 += hi; st2 = fmaf(hi, hi, st2); }
                  }
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmO, slot, col0 + i, m0 + quad * 32);
                bulk_commit();
              }
              ++cc;
            }
          }
        }
      }
      if (STAGED && epi.stats_out != nullptr && rm.live) {
        const int part = (tile % n_tiles) * PARTS + half;
        *reinterpret_cast<float2 *>(epi.stats_out + (rm.ro * epi.stats_parts + part) * 2) = make_float2(st1, st2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cta(tempty_bar(acc), leader);   // the leader's MMA warp waits for both CTAs' epilogues
      }
    }
    if (STAGED && lane == 0) bulk_wait_read<0>();   // staging slots must outlive their last store's read
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody may still signal or read a peer
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void *ptr;
  uint64_t rows, cols, pitch;
  uint32_t box_rows, box_cols;
  bool operator==(const MapKey &o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey &k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1315423911u ^ k.rows;
    h = h * 1315423911u ^ k.cols;
    h = h * 1315423911u ^ k.pitch;
    h = h * 1315423911u ^ k.box_rows;
    h = h * 1315423911u ^ k.box_cols;
    return h;
  }
};

// [rows, cols] bf16 row-major with `pitch` elements per row; box = box_rows x box_cols with the
// swizzle span equal to the box row (64 cols -> 128B operand tiles, 32 cols -> 64B output boxes).
static int get_tensor_map(const void *ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
                          CUtensorMap *out, uint32_t box_cols = BLOCK_K) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, cols, pitch, box_rows, box_cols};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  SDP_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  SDP_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand not 16-byte aligned");
  SDP_CHECK((pitch * 2) % 16 == 0, "GEMM operand row pitch (%llu elements) not a multiple of 8",
            (unsigned long long)pitch);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {pitch * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return 0;
}

static int num_sms() {            // per device: a process may drive several GPUs
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

int make_tensor_map_bf16(const void *ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
                         uint32_t box_cols, CUtensorMap *out) {
  return get_tensor_map(ptr, rows, cols, pitch, box_rows, out, box_cols);
}

// Dense [n2, n1, n0] bf16 tensor, box = 1 x box1 x box0, no swizzle: per-image boxes whose rows past n1 are
// clipped (stores) or zero-filled (loads).  Cached under the same key type (rows = n1, cols = n0, pitch = n2).
int make_tensor_map_bf16_3d(const void *ptr, uint64_t n2, uint64_t n1, uint64_t n0, uint32_t box1, uint32_t box0,
                            CUtensorMap *out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, n1, n0, n2, box1, box0};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  SDP_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  SDP_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (n0 * 2) % 16 == 0 && (box0 * 2) % 16 == 0,
            "3-D tensor map: base, row pitch and box row must be multiples of 16 bytes");
  cuuint64_t gdim[3] = {n0, n1, n2};
  cuuint64_t gstr[2] = {n0 * 2, n1 * n0 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return 0;
}

// [n3][n2][n1][n0] bf16 tensor with element strides (1, s1, s2, s3) and a (b0 x b1 x b2 x 1) box, 64B-swizzled when the
// box row is 64 bytes: the depthwise kernel's [image][grid row][grid column][channel] view of the token-major stream
// (box columns past n1 are zero-filled on loads and clipped on stores).  Not cached: callers keep the map.
int make_tensor_map_bf16_4d(const void *ptr, const uint64_t dims[4], const uint64_t strides_elems[3], const uint32_t box[4],
                            CUtensorMap *out) {
  EncodeTiledFn enc = get_encode_fn();
  SDP_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  SDP_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "4-D tensor map: base not 16-byte aligned");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_elems[0] * 2, strides_elems[1] * 2, strides_elems[2] * 2};
  for (int i = 0; i < 3; ++i) SDP_CHECK(gstr[i] % 16 == 0, "4-D tensor map: stride %d not a multiple of 16 bytes", i);
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = box[0] * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : box[0] * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
  return 0;
}

static int pick_bn(int N) {   // BLOCK_N minimising padded columns; ties -> the larger tile
  int bn = 256;
  long best = -1;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const long padded = (long)((N + cand[i] - 1) / cand[i]) * cand[i];
    if (best < 0 || padded < best) { best = padded; bn = cand[i]; }
  }
  return bn;
}

int gemm_stats_parts(int N) { const int bn = pick_bn(N); return 2 * ((N + bn - 1) / bn); }

// Clusters of two CTA pairs with the weight tile multicast between them: measured on B200 (XL step, batch 1024) at
// +4 % GEMM time -- a quarter less L2 -> SM traffic does not pay for the 16 SMs a cluster of four cannot reach (33
// clusters = 132 of 148 SMs).  Correct (the GEMM tests pass with it on) and kept for diagnostic builds only.
static bool gemm_multicast_enabled() {
#ifdef SDP_DIAG
  static const bool on = [] { const char *v = getenv("SDP_GEMM_MC"); return v && v[0] == '1'; }();
  return on;
#else
  return false;
#endif
}

static bool staged_ok(const Epilogue &e) {
  return e.out_dtype == SDP_BF16 && e.seq_in == 0 && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
         (e.ldo * 2) % 16 == 0 && (e.pass_seq == 0 || e.residual != nullptr);
}

// In-place residual GEMM the RSM epilogue covers: out == residual (and out_lo == res_lo), bf16, plain row mapping,
// whole 32-column chunks.  Returns the number of planes (1: hi only, 2: hi + lo), 0 if not applicable.
static int rsm_planes(const Epilogue &e) {
  if (!staged_ok(e) || e.residual == nullptr || e.residual != e.out || e.res_dtype != SDP_BF16 || e.res_first ||
      e.res_mod != 0 || e.ldr != e.ldo || e.N % 32 != 0 || e.hn_d != 0 || e.ln_stats != nullptr || !epilogue_vec_ok(e))
    return 0;
  if (e.out_lo == nullptr && e.res_lo == nullptr) return 1;
  return (e.out_lo != nullptr && e.out_lo == e.res_lo) ? 2 : 0;
}

template <int BN, int STAGES, int ACT, int HN, bool STAGED, int CG = 1, int EW = EPI_WARPS, int LEAN = 0, int RSM = 0>
static int launch_tc2(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &to, const CUtensorMap &to2,
                      const Epilogue &e, int K, cudaStream_t st, const sdp_gemm_args *wsrc = nullptr) {
  using L = SmemLayout<BN, STAGES, CG, out_bytes_for(LEAN != 0, RSM, CG)>;
  static_assert(STAGES >= 2, "pipeline too shallow");
  auto kern = gemm_bf16_tc_kernel<BN, STAGES, ACT, HN, STAGED, CG, EW, LEAN, RSM>;
  static unsigned long long configured = 0;     // one bit per device ordinal
  static int max_quads[64] = {};                // co-resident clusters of two pairs, per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3((2 + EW) * 32);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG > 1 ? 1 : 0;
  if (!((configured >> dev) & 1ull)) {
    SDP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (CG == 2) {
      cfg.gridDim = dim3(4 * num_sms());
      attr[0].val.clusterDim.x = 4;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
      max_quads[dev] = n;
      attr[0].val.clusterDim.x = CG;
    }
    configured |= 1ull << dev;
  }
  // Clusters of two pairs sharing the weight tile (multicast): worth it once there is more than one wave of tiles,
  // so that the SMs a 4-cluster cannot use (132 of 148 are reachable) cost less than the L2 traffic saved.
  CUtensorMap tw_used = tw;
  int mc = 1;
  if (CG == 2 && wsrc != nullptr && max_quads[dev] > 0 && (BN / 4) % 8 == 0) {
    const long units2 = (long)((e.M + 4 * BLOCK_M - 1) / (4 * BLOCK_M)) * ((e.N + BN - 1) / BN);
    if (units2 >= 2L * max_quads[dev] && gemm_multicast_enabled()) {
      if (int rc = get_tensor_map(wsrc->W, wsrc->N, wsrc->K, wsrc->ldw, BN / 4, &tw_used)) return rc;
      mc = 2;
    }
  }
  const int units = ((e.M + BLOCK_M * CG * mc - 1) / (BLOCK_M * CG * mc)) * ((e.N + BN - 1) / BN);
  const int slots = mc == 2 ? max_quads[dev] : num_sms() / CG;
  cfg.gridDim = dim3(CG * mc * (units < slots ? units : slots));
  attr[0].val.clusterDim.x = CG * mc;
  int flags = (epilogue_vec_ok(e) ? 1 : 0) | (mc == 2 ? 4 : 0);
#ifdef SDP_DIAG
  static const int nofeed = [] { const char *v = getenv("SDP_GEMM_NOFEED"); return (v && v[0] == '1') ? 2 : 0; }();
  flags |= nofeed;
#endif
  SDP_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw_used, to, to2, e, K, flags));
  SDP_LAUNCH_OK();
  return 0;
}

// pairs (cta_group::2, 256 x BN tiles) whenever the staged epilogue applies and M spans enough pairs
static bool use_pairs(const Epilogue &e) {
#ifdef SDP_DIAG
  static const int pref = [] { const char *s = getenv("SDP_GEMM_CTA_GROUP"); return s ? atoi(s) : 0; }();
  if (pref == 1) return false;
  if (pref == 2) return e.M > BLOCK_M;
#endif
  return e.M >= 8 * BLOCK_M;
}

template <int BN, int STAGES, int ACT, int HN = 0>
static int launch_tc(const sdp_gemm_args &a, const CUtensorMap &ta, const Epilogue &e, cudaStream_t st) {
  CUtensorMap tw;
  const int rsm = rsm_planes(e);
  const bool lo = e.out_lo != nullptr || e.res_lo != nullptr;
  if (staged_ok(e) && (!lo || rsm == 2)) {
    CUtensorMap to, to2;
    if (int rc = get_tensor_map(e.out, e.M, e.N, e.ldo, 32, &to, 32)) return rc;
    to2 = to;
    if (rsm == 2)
      if (int rc = get_tensor_map(e.out_lo, e.M, e.N, e.ldo, 32, &to2, 32)) return rc;
    if constexpr (BN >= 128) {
      if (use_pairs(e)) {
        constexpr int ST2 = (STAGES * (BLOCK_M + BN)) / (BLOCK_M + BN / 2);   // same bytes, deeper ring
        if (int rc = get_tensor_map(a.W, a.N, a.K, a.ldw, BN / 2, &tw)) return rc;
        if constexpr (HN == 0) {
          if (rsm == 2) return launch_tc2<BN, stages_for(BN, 2, out_bytes_for(false, 2, 2), ST2), ACT, 0, true, 2, EPI_WARPS, 0, 2>(ta, tw, to, to2, e, a.K, st, &a);
          if (rsm == 1) return launch_tc2<BN, stages_for(BN, 2, out_bytes_for(false, 1, 2), ST2), ACT, 0, true, 2, EPI_WARPS, 0, 1>(ta, tw, to, to2, e, a.K, st, &a);
        }
        if constexpr (ACT == SDP_ACT_GELU && HN == 0 && BN == 256) {
          // bias (or folded LayerNorm) + GELU only (the C -> 4C GEMMs of the mixers and encoders): the lean 16-warp
          // epilogue, one smem stage traded for 64 KB of output staging
          if (e.residual == nullptr && e.stats_out == nullptr && e.pass_seq == 0 && e.seq_out == 0 && e.res_mod == 0 &&
              a.N % BN == 0 && epilogue_vec_ok(e)) {
            if (e.ln_stats != nullptr)
              return launch_tc2<BN, ST2 - 1, ACT, HN, true, 2, EPI_WARPS_WIDE, 2>(ta, tw, to, to2, e, a.K, st, &a);
            return launch_tc2<BN, ST2 - 1, ACT, HN, true, 2, EPI_WARPS_WIDE, 1>(ta, tw, to, to2, e, a.K, st, &a);
          }
        }
        return launch_tc2<BN, ST2, ACT, HN, true, 2>(ta, tw, to, to2, e, a.K, st, &a);
      }
    }
    if (int rc = get_tensor_map(a.W, a.N, a.K, a.ldw, BN, &tw)) return rc;
    if constexpr (HN == 0) {
      if (rsm == 2) return launch_tc2<BN, stages_for(BN, 1, out_bytes_for(false, 2, 1), STAGES), ACT, 0, true, 1, EPI_WARPS, 0, 2>(ta, tw, to, to2, e, a.K, st);
      if (rsm == 1) return launch_tc2<BN, stages_for(BN, 1, out_bytes_for(false, 1, 1), STAGES), ACT, 0, true, 1, EPI_WARPS, 0, 1>(ta, tw, to, to2, e, a.K, st);
    }
    if (!lo) return launch_tc2<BN, STAGES, ACT, HN, true>(ta, tw, to, to2, e, a.K, st);
  }
  // direct global stores: row remaps (patch embedding), fp32 outputs, split-stream shapes the RSM epilogue does not cover
  SDP_CHECK(e.stats_out == nullptr, "sdp_gemm: stats_out needs an epilogue that stages its output (bf16, no row remap; with "
                                    "a split stream: in place, N %% 32 == 0)");
  if (int rc = get_tensor_map(a.W, a.N, a.K, a.ldw, BN, &tw)) return rc;
  return launch_tc2<BN, STAGES, ACT, HN, false>(ta, tw, ta, ta, e, a.K, st);
}

template <int BN, int STAGES>
static int launch_tc_act(const sdp_gemm_args &a, const CUtensorMap &ta, const Epilogue &e, cudaStream_t st) {
  if (e.act == SDP_ACT_NONE) return launch_tc<BN, STAGES, SDP_ACT_NONE>(a, ta, e, st);
  if (e.act == SDP_ACT_GELU) return launch_tc<BN, STAGES, SDP_ACT_GELU>(a, ta, e, st);
  return launch_tc<BN, STAGES, -1>(a, ta, e, st);
}

int gemm_bf16_tc(const sdp_gemm_args &a, const Epilogue &e, cudaStream_t st) {
  CUtensorMap ta;
  if (int rc = get_tensor_map(a.A, a.M, a.K, a.lda, BLOCK_M, &ta)) return rc;
  if (e.hn_d) {
    // QKV projection with the per-head LayerNorm fused: the N block must hold whole heads
    SDP_CHECK(e.act == SDP_ACT_NONE, "sdp_gemm: head-norm epilogue has no activation");
    SDP_CHECK(e.stats_out == nullptr, "sdp_gemm: head-norm GEMMs do not emit row statistics");
    SDP_CHECK(e.ln_stats == nullptr || (staged_ok(e) && epilogue_vec_ok(e)), "sdp_gemm: LN folding needs the staged bf16 epilogue");
    if (e.hn_d == 96) return launch_tc<192, 4, SDP_ACT_NONE, 96>(a, ta, e, st);
    if (a.N % 256 == 0) {
      if (e.hn_d == 32) return launch_tc<256, 4, SDP_ACT_NONE, 32>(a, ta, e, st);
      if (e.hn_d == 64) return launch_tc<256, 4, SDP_ACT_NONE, 64>(a, ta, e, st);
      return launch_tc<256, 4, SDP_ACT_NONE, 128>(a, ta, e, st);
    }
    if (e.hn_d == 32) return launch_tc<128, 6, SDP_ACT_NONE, 32>(a, ta, e, st);
    if (e.hn_d == 64) return launch_tc<128, 6, SDP_ACT_NONE, 64>(a, ta, e, st);
    return launch_tc<128, 6, SDP_ACT_NONE, 128>(a, ta, e, st);
  }
  SDP_CHECK(e.stats_out == nullptr || staged_ok(e), "sdp_gemm: stats_out needs the staged bf16 epilogue");
  SDP_CHECK(e.ln_stats == nullptr || (staged_ok(e) && a.N % 32 == 0 && epilogue_vec_ok(e)),
            "sdp_gemm: LN folding needs the staged bf16 epilogue and N %% 32 == 0");
  const int bn = pick_bn(a.N);
  switch (bn) {
    case 256: return launch_tc_act<256, 4>(a, ta, e, st);
    case 128: return launch_tc_act<128, 6>(a, ta, e, st);
    default: return launch_tc_act<64, 8>(a, ta, e, st);
  }
}

}  // namespace sdp
