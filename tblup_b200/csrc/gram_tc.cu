// Masked Gram kernel: C_w = Z_w Z_w^T (uncentred integer cross-products) for a wave of individuals,
// where Z_w = panel[w] is the K-major int8 gather of the individual's markers.
//
// This is the `np.matmul(W, W^T)` of tblup/utils.py:17 (called from tblup/evaluator.py:275) with the
// centring factored out: the uncentred cross-products are exact integers (bit-for-bit equal to the
// oracle), the centring is applied afterwards as exact rank-1 integer corrections (scale.cu).
//
// sm_100a design: persistent, warp-specialised, one CTA per SM.
//   warp 0      TMA producer   cp.async.bulk.tensor 2-D loads (128-byte swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer     one thread issues tcgen05.mma.kind::i8 (M=128, N=256, K=32), s32 accumulators
//                              in TMEM, two accumulator stages (2 x 256 columns) so the epilogue of tile t
//                              overlaps the main loop of tile t+1
//   warps 2..9  epilogue       tcgen05.ld 32x32b -> registers -> 16-byte global stores of the int32 tile
//                              (optionally also the scaled fp32 matrix, see GramFuse)
// Only tiles that touch the lower triangle are scheduled (tile list built on the host per row set).
#include "tb_internal.h"
#include "tb_ptx.cuh"

namespace {

using namespace tbptx;

constexpr int BM = TB_GRAM_BM, BN = TB_GRAM_BN, BK = TB_GRAM_BK;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK;                 // 16 KiB
constexpr int B_BYTES = BN * BK;                 // 32 KiB (int8 tiles; the fp4 variant uses 224 rows = 28 KiB)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KiB (stage slots are sized for the wider variant)
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;                   // int8: 2 x 256 accumulator columns; fp4: 2 x 224 + scale factors
// fp4 variant (kind::mxf4, E2M1 operands, K = 64 per instruction): dosages 0/1/2 are exact in E2M1 and every partial
// sum is an integer below 2^24, so the fp32 accumulators hold the same integers the int8 path produces -- at twice
// the MMA rate and half the operand bytes.  The block scale factors (UE8M0, one per 32 elements) are all 2^0: their
// TMEM columns are filled once with 0x7f bytes, whatever the layout.  TMEM has 512 columns in total, so the tile is
// 224 wide: 2 x 224 accumulator columns + 32 columns of scale factors.
constexpr int BN_I8 = BN;
constexpr int BN4 = TB_GRAM_BN_FP4;              // 224
constexpr int ACC_STRIDE4 = 240;
constexpr int SF_COL = 480;
constexpr int EPI_WARPS = 8;                     // two per TMEM lane quarter, each owning half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int COLTERM_BYTES = 2 * BN * 8;         // fused scaling: two buffers of per-column terms (fp32; sized generously)
constexpr int STAGING_BYTES = EPI_WARPS * 2048;   // per-warp transposition buffers of the coalescing epilogue
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + COLTERM_BYTES +
                           STAGING_BYTES;

// Fused scaling (mixed-precision path, one contiguous row set): besides the raw cross-products the epilogue writes
// the fp32 matrix the tensor-core Cholesky factors, A = G_tt + lambda I, straight from the accumulator:
//   A_ab = (2 N^2/den) C_ab - (2 N/den) s_a + (2/den)(Q - N s_b)  (+ lambda on the diagonal),  identity padding,
// i.e. what scale32_kernel (solve_mixed.cu) computes, without re-reading C from HBM (coefficients from the exact
// integers in fp64, evaluation in fp32: this matrix only feeds the 10-bit preconditioner).
struct GramFuse {
  const TbScaleJob* jobs;   // [W] (one row set): s, SQ, N, n_t, ntp, lambda
  float* L32;               // [W][ntp_all][ntp_all]
  int ntp_all;
};
constexpr uint32_t IDESC = umma_idesc_s8(BM, BN);

constexpr int MAX_STAGES = 8;
// ring depth: a CTA of a tcgen05 pair keeps only its half of the B tile, so its k-block is 30 KB (fp4) / 32 KB (int8)
// instead of 44 / 48 KB and six stages fit where four did -- the ring is what hides the L2 / HBM latency of the operand
// stream (r02: single-CTA, multicast-pair and CTA-pair kernels all ran at the same ~60 % of the MMA rate with four)
template <int PAIR, bool FP4> struct RingCfg {
  static constexpr int STAGE = PAIR == 2 ? A_BYTES + ((FP4 ? TB_GRAM_BN_FP4 : BN) / 2) * BK : STAGE_BYTES;
  static constexpr int N = PAIR == 2 ? 6 : STAGES;
};
struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t acc_full[ACC_STAGES];
  uint64_t acc_empty[ACC_STAGES];
  uint32_t tmem_base;
};

// C16: the cross-products are stored as int16 (the caller guarantees 4 k <= 32 767, so C_ab <= 4 k fits): half the
// bytes for every later pass over C (scaling, refinement mat-vecs, predictions).  Row stride stays rpad ELEMENTS.
// PAIR: the kernel runs as clusters of two CTAs that take tiles (I, J) and (I + 1, J) of the same genome together: they
// need the same B rows, so each CTA fetches HALF of the B tile and TMA multicasts it into both shared memories -- the
// L2 -> SM operand traffic per MMA drops by a third (r01 ncu: 15.5 TB/s of L2 reads held the tensor pipe at 57 %).
// The MMAs stay cta_group::1; a ring slot is released to both producers by a multicast commit.
// PAIR == 2: the pair runs as ONE tcgen05 CTA pair (cta_group::2): M = 256 (128 rows from each CTA), each CTA fetches
// and keeps only its HALF of the B tile (no multicast), one thread of the leader CTA issues the MMAs for both tensor
// cores.  Per k-block a CTA's shared memory then takes 30 KB of TMA writes and 30 KB of operand reads instead of
// 44 + 44 KB -- at the mxf4 rate the single-CTA form needs ~164 B/clk of a 128 B/clk shared-memory port.
template <bool FUSE, bool C16, bool FP4, int PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b,
               const int* __restrict__ tiles, int n_tiles,
               const int* __restrict__ kblocks, int W, int rpad, int32_t* __restrict__ C, const GramFuse fz) {
  constexpr int BN = FP4 ? BN4 : BN_I8;                     // tile columns of this variant
  constexpr int ACC_STRIDE = FP4 ? ACC_STRIDE4 : BN_I8;     // TMEM columns between the two accumulators
  constexpr uint32_t STAGE_TX = A_BYTES + BN * BK;
  constexpr int STAGES = RingCfg<PAIR, FP4>::N;              // (shadows the file-scope constants: this variant's ring)
  constexpr int STAGE_BYTES = RingCfg<PAIR, FP4>::STAGE;
  constexpr int RING_BYTES = 4 * 49152;                      // the ring region is the same size for every variant
  static_assert(STAGES * STAGE_BYTES <= RING_BYTES && STAGE_BYTES % 1024 == 0, "ring layout");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + RING_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = W * n_tiles;
  const int crank = PAIR ? (int)cluster_ctarank() : 0;          // which tile of the pair / which half of B this CTA loads
  constexpr bool CG2 = PAIR == 2;
  constexpr uint32_t STAGE_TX2 = A_BYTES + (BN / 2) * BK;       // what ONE CTA of a tcgen05 pair loads per k-block
  const int worker = PAIR ? blockIdx.x >> 1 : blockIdx.x, n_workers = PAIR ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], PAIR == 1 ? 2 : 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&bars->acc_full[s], 1);
      mbar_init(&bars->acc_empty[s], CG2 ? 2 * EPI_WARPS : EPI_WARPS);   // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG2) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base);
    else tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (PAIR) cluster_sync_all();       // the partner's barriers are initialised before anything is multicast to them
  if (FP4) {
    // scale factors: every byte 0x7f = 2^0 (UE8M0); columns SF_COL .. SF_COL + 31 of all 128 lanes
    if (warp >= 2 && warp < 6) tmem_fill_32x32(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + SF_COL, 0x7f7f7f7fu);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < n_items; item += n_workers) {
        const int w = item / n_tiles, t = tiles[item - w * n_tiles];
        const int row_a = w * rpad + ((t >> 16) + crank) * BM;
        const int row_b = w * rpad + (t & 0xffff) * BN;
        const int nkb = kblocks[w];
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (CG2) {
            // both CTAs' loads are counted on the LEADER's ring barrier (its MMA thread is the only consumer)
            if (crank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2 * STAGE_TX2);
            const uint32_t lead_full = mapa_u32(smem_u32(&bars->full[stage]), 0);
            tma_load_2d_pair(sa, &tmap, lead_full, kb * BK, row_a);
            tma_load_2d_pair(sa + A_BYTES, &tmap_b, lead_full, kb * BK, row_b + crank * (BN / 2));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&bars->full[stage], STAGE_TX);
          tma_load_2d(sa, &tmap, &bars->full[stage], kb * BK, row_a);
          if (PAIR) {
            tma_load_2d_multicast(sa + A_BYTES + crank * (BN / 2) * BK, &tmap_b, &bars->full[stage], kb * BK,
                                  row_b + crank * (BN / 2), (uint16_t)3);
          } else {
            tma_load_2d(sa + A_BYTES, &tmap_b, &bars->full[stage], kb * BK, row_b);
            tma_load_2d(sa + A_BYTES + (BN / 2) * BK, &tmap_b, &bars->full[stage], kb * BK, row_b + BN / 2);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (with PAIR == 2: the leader CTA's thread drives both tensor cores) ============
    if (lane == 0 && !(CG2 && crank != 0)) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = worker; item < n_items; item += n_workers) {
        const int w = item / n_tiles;
        const int nkb = kblocks[w];
        mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 32; ++k) {
            // advance 32 bytes along K inside the 128-byte swizzle span: +2 in 16-byte units
            if (CG2) {
              if (FP4)
                umma_mxf4_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, umma_idesc_mxf4(2 * BM, BN), (kb | k) != 0,
                               tmem_base + SF_COL, tmem_base + SF_COL + 8);
              else
                umma_s8_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, umma_idesc_s8(2 * BM, BN), (kb | k) != 0);
            } else if (FP4)
              umma_mxf4(tmem_d, adesc + 2 * k, bdesc + 2 * k, umma_idesc_mxf4(BM, BN), (kb | k) != 0,
                        tmem_base + SF_COL, tmem_base + SF_COL + 8);
            else
              umma_s8(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
          }
          if (CG2) umma_commit_pair(&bars->empty[stage], (uint16_t)3);
          else if (PAIR) umma_commit_multicast(&bars->empty[stage], (uint16_t)3);
          else umma_commit(&bars->empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CG2) umma_commit_pair(&bars->acc_full[acc], (uint16_t)3);
        else umma_commit(&bars->acc_full[acc]);
        if (++acc == ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // which half of the tile's 32-column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t colterm = smem_u32(smem + RING_BYTES + 256);     // [2][BN] doubles
    const uint32_t stg = smem_u32(smem + RING_BYTES + 256 + COLTERM_BYTES) + (warp - 2) * 2048;
    int fbuf = 0;
    for (int item = worker; item < n_items; item += n_workers) {
      const int w = item / n_tiles, t = tiles[item - w * n_tiles];
      const int ti = (t >> 16) + crank, tj = t & 0xffff;
      const int row = ti * BM + q * 32 + lane;
      const int row_hi = ti * BM + BM - 1;
      // ---- fused scaling: per-tile terms, prepared while the tile's MMAs run
      bool a_tile = false;
      int f_nt = 0, f_ntp = 0;
      float f_scale = 0.f, f_lam = 0.f;
      float f_rowterm[4] = {0.f, 0.f, 0.f, 0.f};      // rows 16 hh + 8 rr + (lane >> 2) of this warp, index 2 hh + rr
      uint32_t ct = 0;
      if (FUSE) {
        const TbScaleJob& jb = fz.jobs[w];
        f_ntp = jb.ntp;
        f_nt = jb.n_t;
        a_tile = ti * BM < f_ntp && tj * BN < f_ntp;      // uniform over the four epilogue warps
        if (a_tile) {
          // G_ab + [a = b] lambda = (2 N^2 / den) C_ab - (2 N / den) s_a + (2 / den) (Q - N s_b): the three
          // coefficients are formed in fp64 from the exact integers and rounded once, so the fp32 evaluation below only
          // combines O(1) quantities (absolute error ~1e-7 of the diagonal; the factorisation rounds to 10 bits, and
          // the refinement in solve_mixed.cu works from the exact integers, not from this matrix)
          const long long N = jb.N, S = jb.SQ[0], Q = jb.SQ[1];
          const double inv_d = 2.0 / (double)(2 * N * S - Q);
          f_scale = (float)(inv_d * (double)(N * N));
          f_lam = (float)jb.lambda;
          const uint32_t cw = colterm + fbuf * BN_I8 * 4;
          for (int e = threadIdx.x - 64; e < BN; e += 32 * EPI_WARPS) {
            const int b = tj * BN + e;
            const float term = b < f_nt ? (float)(inv_d * (double)(Q - N * jb.s[b])) : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(cw + e * 4), "f"(term) : "memory");
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int grow = ti * BM + q * 32 + 16 * (i >> 1) + 8 * (i & 1) + (lane >> 2);
            f_rowterm[i] = grow < f_nt ? (float)(inv_d * (double)(-N * jb.s[grow])) : 0.f;
          }
          ct = cw;
          fbuf ^= 1;
          // the buffer written two A-tiles ago is free again: every warp passed this barrier once since
          asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        }
      }
      mbar_wait(&bars->acc_full[acc], acc_phase);
      tc_fence_after();
      // the warp's 32 rows; stores go through the staging buffer so that every instruction writes whole sectors
      const size_t cwarp_off = ((size_t)w * rpad + ti * BM + q * 32) * rpad;
      int32_t* cwarp = C + cwarp_off;
      int16_t* cwarp16 = reinterpret_cast<int16_t*>(C) + cwarp_off;
      float* awarp = (FUSE && a_tile) ? fz.L32 + ((size_t)w * fz.ntp_all + ti * BM + q * 32) * fz.ntp_all : nullptr;
      const int rl = lane >> 2, gl = 4 * (lane & 3);          // read-back role: row 8 it + rl, words gl .. gl + 3
#pragma unroll 1
      for (int c = chalf * 4; c < (chalf + 1) * 4 && c < BN / 32; ++c) {
        const int col0 = tj * BN + c * 32;
        if (col0 >= rpad || col0 > row_hi || ti * BM >= rpad) continue;   // outside the matrix / strictly above the diagonal
        if (!FUSE && fz.ntp_all == -2) continue;                         // experiment: no epilogue work at all
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_STRIDE + c * 32, v);
        tmem_ld_wait();
        float vf[32];                                  // the cross-products as floats (exact: integers below 2^24)
#pragma unroll
        for (int e = 0; e < 32; ++e) vf[e] = FP4 ? __uint_as_float(v[e]) : 0.f;
        if (FP4) {
          // fp32 accumulators -> integers: adding 2^23 leaves the integer in the low mantissa bits (values < 2^23;
          // the int16 layout only needs 15 of them); the int32 layout goes through the converter
#pragma unroll
          for (int e = 0; e < 32; ++e)
            v[e] = C16 ? (__float_as_uint(vf[e] + 8388608.f) & 0xffffu) : (uint32_t)__float2int_rn(vf[e]);
        }
        if (C16) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = (v[2 * e] & 0xffffu) | (v[2 * e + 1] << 16);
          stage_write16(stg, lane, pk);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const uint4 u = stage_read16(stg, lane, it);
            if (FUSE || fz.ntp_all != -1)                                  // (experiment -1: everything but the global stores)
              *reinterpret_cast<uint4*>(cwarp16 + (size_t)(8 * it + rl) * rpad + col0 + 2 * gl) = u;
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            stage_write16(stg, lane, v + 16 * h);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const uint4 u = stage_read16(stg, lane, it);
              *reinterpret_cast<uint4*>(cwarp + (size_t)(8 * it + rl) * rpad + col0 + 16 * h + gl) = u;
            }
            __syncwarp();
          }
        }
        if (FUSE && a_tile && col0 < f_ntp) {
          // second read of the chunk in the accumulator-fragment layout: four lanes own 32 contiguous bytes of a row,
          // so the scaled matrix goes out with 8-byte stores of whole sectors and no shared-memory transposition
          // (the staging traffic of this, the larger, output competed with the MMA operand reads)
          const int t0 = lane & 3, t1 = lane >> 2;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t w[16];
            tmem_ld_16x256b_x4(tmem_base + (static_cast<uint32_t>(q * 32 + 16 * hh) << 16) + acc * ACC_STRIDE + c * 32, w);
            tmem_ld_wait();
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int rloc = 16 * hh + 8 * rr + t1;
              const int grow = ti * BM + q * 32 + rloc;
              if (grow >= f_ntp) continue;
              const float rt = f_rowterm[2 * hh + rr];
              float* drow = awarp + (size_t)rloc * fz.ntp_all + col0 + 2 * t0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float c0f, c1f;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(c0f), "=f"(c1f) : "r"(ct + (c * 32 + 8 * j + 2 * t0) * 4));
                const uint32_t w0 = w[4 * j + 2 * rr], w1 = w[4 * j + 2 * rr + 1];
                const float x0 = FP4 ? __uint_as_float(w0) : (float)(int)w0;
                const float x1 = FP4 ? __uint_as_float(w1) : (float)(int)w1;
                const int b0 = col0 + 8 * j + 2 * t0;
                // columns b >= n_t only occur above the diagonal of a training row (never read)
                float g0 = fmaf(x0, f_scale, rt + c0f), g1 = fmaf(x1, f_scale, rt + c1f);
                if (b0 == grow) g0 += f_lam;
                if (b0 + 1 == grow) g1 += f_lam;
                if (grow >= f_nt) {                          // identity padding rows
                  g0 = b0 == grow ? 1.f : 0.f;
                  g1 = b0 + 1 == grow ? 1.f : 0.f;
                }
                *reinterpret_cast<float2*>(drow + 8 * j) = make_float2(g0, g1);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->acc_empty[acc]), 0));   // the leader's MMA thread waits for both CTAs
        else mbar_arrive(&bars->acc_empty[acc]);
      }
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();       // nobody leaves while its partner may still multicast into its ring or barriers
  if (warp == 1) {
    tc_fence_after();
    if (CG2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

cudaError_t tb_gram_tc_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* fn) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  };
  set((const void*)gram_tc_kernel<false, false, false, 0>);
  set((const void*)gram_tc_kernel<true, false, false, 0>);
  set((const void*)gram_tc_kernel<false, true, false, 0>);
  set((const void*)gram_tc_kernel<true, true, false, 0>);
  set((const void*)gram_tc_kernel<false, false, true, 0>);
  set((const void*)gram_tc_kernel<true, false, true, 0>);
  set((const void*)gram_tc_kernel<false, true, true, 0>);
  set((const void*)gram_tc_kernel<true, true, true, 0>);
  set((const void*)gram_tc_kernel<false, false, false, 1>);
  set((const void*)gram_tc_kernel<false, true, false, 1>);
  set((const void*)gram_tc_kernel<false, false, true, 1>);
  set((const void*)gram_tc_kernel<false, true, true, 1>);
  set((const void*)gram_tc_kernel<false, false, false, 2>);
  set((const void*)gram_tc_kernel<false, true, false, 2>);
  set((const void*)gram_tc_kernel<false, false, true, 2>);
  set((const void*)gram_tc_kernel<false, true, true, 2>);
  return e;
}

// d_panel must have (W * rpad + 128) rows of kstride BYTES allocated (slack for the last B half-tile).
// fp4 != 0: the panel holds E2M1 nibbles (two markers per byte, dosage d stored as 2 d), kstride = padded k / 2, one
// k-block = 256 markers, tiles are TB_GRAM_BN_FP4 columns wide (the tile list must be built for that width).
cudaError_t tb_launch_gram_tc(const int8_t* d_panel, int W, int rpad, int kstride, const int* d_kblocks,
                              const int* d_tiles, int n_tiles, int32_t* d_C, int n_sm, cudaStream_t st,
                              std::string* err, const TbScaleJob* d_fuse_jobs, float* d_L32, int ntp_all, int c16,
                              int fp4, int pair) {
  CUtensorMap tmap, tmap_b;
  const cuuint64_t dims[2] = {(cuuint64_t)kstride, (cuuint64_t)W * rpad + 128};
  const cuuint64_t strides[1] = {(cuuint64_t)kstride};
  const cuuint32_t estr[2] = {1, 1};
  for (int which = 0; which < 2; ++which) {
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(which == 0 ? 128 : (fp4 ? BN4 / 2 : BN / 2))};
    CUresult r = g_encode(which == 0 ? &tmap : &tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(d_panel),
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
  }
  const int n_items = W * n_tiles;
  const GramFuse fz{d_fuse_jobs, d_L32, ntp_all};
  if (pair) {
    // clusters of two CTAs; d_tiles holds PAIR items (first row block of the pair << 16 | column block)
    if (d_fuse_jobs) {
      if (err) *err = "gram: the paired kernel does not write the scaled matrix";
      return cudaErrorInvalidValue;
    }
    const int clusters = std::min(n_items, n_sm / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int which = (pair == 2 ? 4 : 0) | (c16 ? 2 : 0) | (fp4 ? 1 : 0);
#define TB_GRAM_PAIR(S, P4, PR) \
  return cudaLaunchKernelEx(&cfg, gram_tc_kernel<false, S, P4, PR>, tmap, tmap_b, d_tiles, n_tiles, d_kblocks, W, rpad, d_C, fz)
    switch (which) {
      case 0: TB_GRAM_PAIR(false, false, 1);
      case 1: TB_GRAM_PAIR(false, true, 1);
      case 2: TB_GRAM_PAIR(true, false, 1);
      case 3: TB_GRAM_PAIR(true, true, 1);
      case 4: TB_GRAM_PAIR(false, false, 2);
      case 5: TB_GRAM_PAIR(false, true, 2);
      case 6: TB_GRAM_PAIR(true, false, 2);
      default: TB_GRAM_PAIR(true, true, 2);
    }
#undef TB_GRAM_PAIR
  }
  const int grid = n_items < n_sm ? n_items : n_sm;
  const int which = (d_fuse_jobs ? 4 : 0) | (c16 ? 2 : 0) | (fp4 ? 1 : 0);
#define TB_GRAM_LAUNCH(F, S, P4) \
  gram_tc_kernel<F, S, P4, 0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmap, tmap_b, d_tiles, n_tiles, d_kblocks, W, rpad, d_C, fz)
  switch (which) {
    case 0: TB_GRAM_LAUNCH(false, false, false); break;
    case 1: TB_GRAM_LAUNCH(false, false, true); break;
    case 2: TB_GRAM_LAUNCH(false, true, false); break;
    case 3: TB_GRAM_LAUNCH(false, true, true); break;
    case 4: TB_GRAM_LAUNCH(true, false, false); break;
    case 5: TB_GRAM_LAUNCH(true, false, true); break;
    case 6: TB_GRAM_LAUNCH(true, true, false); break;
    default: TB_GRAM_LAUNCH(true, true, true); break;
  }
#undef TB_GRAM_LAUNCH
  return cudaGetLastError();
}
