// Marker gather for one wave of individuals, and the integer ingredients of the exact centring.
//
//   panel[w][a][j] = x[idx_w[j]][a]          (a: universe animal position, j: position in the genome)
//
// i.e. the reference's `data[:, indices]` (tblup/evaluator.py:275 and :298) restricted to the animals the
// fitness depends on, written K-major (markers contiguous) so the tcgen05 Gram kernel can stream it with
// TMA.  A marker listed twice is copied twice (numpy fancy-indexing semantics).  HBM-bound: every
// selected marker row is read once (coalesced, 128 B per warp) and every panel byte written once.
#include "tb_internal.h"

namespace {

// 128 markers x 128 animals per block, transposed through shared memory as 32-bit words.
// Word (jl, c) holds animals 4c..4c+3 of marker jl and is stored at column (c + (jl >> 2)) & 31,
// which makes both the row-wise fill and the 4x4 byte-transposing drain bank-conflict free.
__global__ void __launch_bounds__(256) gather_kernel(const int8_t* __restrict__ x, int ldn,
                                                     const int* __restrict__ idx, const long long* __restrict__ off,
                                                     int w0, int rpad, int kstride, int8_t* __restrict__ panel) {
  __shared__ uint32_t tile[128][32];
  const int w = blockIdx.z;
  const long long o0 = off[w0 + w];
  const int k = (int)(off[w0 + w + 1] - o0);
  const int j0 = blockIdx.x * 128;
  if (j0 >= tb_round_up(k, TB_GRAM_BK)) return;  // beyond this individual's (padded) K extent: never read
  const int a0 = blockIdx.y * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int jl = warp; jl < 128; jl += 8) {
    const int j = j0 + jl;
    uint32_t v = 0;
    if (j < k) {
      const int src = idx[o0 + j];
      v = *reinterpret_cast<const uint32_t*>(x + (size_t)src * ldn + a0 + lane * 4);
    }
    tile[jl][(lane + (jl >> 2)) & 31] = v;
  }
  __syncthreads();

  int8_t* out = panel + ((size_t)w * rpad + a0) * kstride + j0;
  // each warp drains 4 animals (one word column c) per iteration; lane q covers markers 4q..4q+3
  for (int c = warp; c < 32; c += 8) {
    const int q = lane;
    const int col = (c + q) & 31;
    const uint32_t r0 = tile[4 * q + 0][col], r1 = tile[4 * q + 1][col], r2 = tile[4 * q + 2][col],
                   r3 = tile[4 * q + 3][col];
    // 4x4 byte transpose: output word for animal 4c+i = bytes i of (r0, r1, r2, r3)
    const uint32_t t0 = __byte_perm(r0, r1, 0x5140);  // r0.b0 r1.b0 r0.b1 r1.b1
    const uint32_t t1 = __byte_perm(r2, r3, 0x5140);
    const uint32_t t2 = __byte_perm(r0, r1, 0x7362);  // r0.b2 r1.b2 r0.b3 r1.b3
    const uint32_t t3 = __byte_perm(r2, r3, 0x7362);
    const uint32_t o_0 = __byte_perm(t0, t1, 0x5410);
    const uint32_t o_1 = __byte_perm(t0, t1, 0x7632);
    const uint32_t o_2 = __byte_perm(t2, t3, 0x5410);
    const uint32_t o_3 = __byte_perm(t2, t3, 0x7632);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)(4 * c) * kstride) + q;
    const size_t rs = kstride / 4;
    dst[0] = o_0;
    dst[rs] = o_1;
    dst[2 * rs] = o_2;
    dst[3 * rs] = o_3;
  }
}

// Gather from the 2-bit packed matrix, tile = (32 * MPL) markers x 512 animals.  The tile stays PACKED in shared
// memory (128 B per marker = 512 animals), so a block moves four times the animals of gather_kernel per byte of
// shared memory: every selected marker contributes one full 128-byte line per block (read side) and every animal row
// receives 32 * MPL contiguous bytes per block (write side, 8- or 16-byte stores, a warp writes 256 / 512 B
// contiguous) -- the large-granule access pattern HBM wants, at a quarter of the read traffic.
// Word (jl, c) = animals 16c .. 16c+15 of marker jl sits at column (c + jl / MPL) & 31: the fill (lane = c) and the
// drain (lane = marker group, fixed c) are both bank-conflict free.
template <int MPL>
__global__ void __launch_bounds__(256) gather_packed_kernel(const uint8_t* __restrict__ x2, int ld4,
                                                            const int* __restrict__ idx, const long long* __restrict__ off,
                                                            int w0, int rpad, int kstride, int8_t* __restrict__ panel) {
  constexpr int TM = 32 * MPL;
  extern __shared__ uint32_t ptile[];            // [TM][32]
  const int w = blockIdx.z;
  const long long o0 = off[w0 + w];
  const int k = (int)(off[w0 + w + 1] - o0);
  const int j0 = blockIdx.x * TM;
  const int kpad = tb_round_up(k, TB_GRAM_BK);   // the Gram reads [0, kpad): zeros beyond k
  if (j0 >= kpad) return;
  const int a0 = blockIdx.y * 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool lane_in = (a0 >> 2) + 4 * lane < ld4;

  for (int jb = warp * 8; jb < TM; jb += 64) {   // 8 rows per warp per round, loads issued together
    int src[8];
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) src[u] = (j0 + jb + u < k) ? idx[o0 + j0 + jb + u] : -1;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = (src[u] >= 0 && lane_in)
                 ? *reinterpret_cast<const uint32_t*>(x2 + (size_t)src[u] * ld4 + (a0 >> 2) + 4 * lane) : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) ptile[(jb + u) * 32 + ((lane + (jb + u) / MPL) & 31)] = v[u];
  }
  __syncthreads();

  const int jl0 = lane * MPL;                    // this lane's markers inside the tile
  if (j0 + jl0 >= kpad) return;                  // (kpad is a multiple of 128 >= MPL: whole lanes)
  for (int c = warp; c < 32; c += 8) {
    if (a0 + 16 * c >= rpad) break;              // rpad is a multiple of 128: whole 16-animal groups
    uint32_t wv[MPL];
#pragma unroll
    for (int t = 0; t < MPL; ++t) wv[t] = ptile[(jl0 + t) * 32 + ((c + lane) & 31)];
    int8_t* dst = panel + ((size_t)w * rpad + a0 + 16 * c) * kstride + j0 + jl0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t o[MPL / 4];
#pragma unroll
      for (int g = 0; g < MPL / 4; ++g)
        o[g] = ((wv[4 * g] >> (2 * i)) & 3u) | (((wv[4 * g + 1] >> (2 * i)) & 3u) << 8) |
               (((wv[4 * g + 2] >> (2 * i)) & 3u) << 16) | (((wv[4 * g + 3] >> (2 * i)) & 3u) << 24);
      if (MPL == 8)
        *reinterpret_cast<uint2*>(dst + (size_t)i * kstride) = make_uint2(o[0], o[1]);
      else
        *reinterpret_cast<uint4*>(dst + (size_t)i * kstride) = make_uint4(o[0], o[1], o[2 % (MPL / 4)], o[3 % (MPL / 4)]);
    }
  }
}

// The same gather for the fp4 Gram (gram_tc_kernel<.., FP4>): the panel holds E2M1 nibbles, two markers per byte,
// marker 2q in the low nibble of byte q; dosage d is the nibble 2 d (E2M1: 0b0010 = 1.0, 0b0100 = 2.0).  Tile = 32 MPL
// markers x 512 animals kept packed in shared memory, a lane owns MPL markers = MPL / 2 output bytes per animal.
// Half the panel bytes of the int8 layout.
template <int MPL>
__global__ void __launch_bounds__(256) gather_fp4_kernel(const uint8_t* __restrict__ x2, int ld4,
                                                         const int* __restrict__ idx, const long long* __restrict__ off,
                                                         int w0, int rpad, int kstride_b, int8_t* __restrict__ panel,
                                                         const int* __restrict__ rowmap, int rows_univ) {
  constexpr int TM = 32 * MPL;
  extern __shared__ uint32_t ptile[];            // [TM][32]
  const int w = blockIdx.z;
  const long long o0 = off[w0 + w];
  const int k = (int)(off[w0 + w + 1] - o0);
  const int j0 = blockIdx.x * TM;
  const int kpad = tb_round_up(k, TB_GRAM_BK_FP4);   // the Gram reads [0, kpad): zeros beyond k
  if (j0 >= kpad) return;
  const int a0 = blockIdx.y * 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool lane_in = (a0 >> 2) + 4 * lane < ld4;

  for (int jb = warp * 8; jb < TM; jb += 64) {
    int src[8];
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) src[u] = (j0 + jb + u < k) ? idx[o0 + j0 + jb + u] : -1;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = (src[u] >= 0 && lane_in)
                 ? *reinterpret_cast<const uint32_t*>(x2 + (size_t)src[u] * ld4 + (a0 >> 2) + 4 * lane) : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) ptile[(jb + u) * 32 + ((lane + (jb + u) / MPL) & 31)] = v[u];
  }
  __syncthreads();

  const int jl0 = lane * MPL;
  if (j0 + jl0 >= kpad) return;                  // kpad is a multiple of 256 >= MPL: whole lanes
  for (int c = warp; c < 32; c += 8) {
    if (a0 + 16 * c >= (rowmap ? rows_univ : rpad)) break;
    uint32_t wv[MPL];
#pragma unroll
    for (int t = 0; t < MPL; ++t) wv[t] = ptile[(jl0 + t) * 32 + ((c + lane) & 31)];
    int8_t* dst0 = panel + (size_t)w * rpad * kstride_b + ((j0 + jl0) >> 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      // panel row of this animal: its universe position, or wherever the row set's permutation sends it
      int prow = a0 + 16 * c + i;
      if (rowmap) prow = prow < rows_univ ? rowmap[prow] : -1;
      if (prow < 0) continue;
      int8_t* dst = dst0 + (size_t)prow * kstride_b;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) {                // (bfe/bfi inline PTX was tried here: slower than shift + LOP3)
        lo |= ((wv[t] >> (2 * i)) & 3u) << (4 * t + 1);
        if (MPL == 16) hi |= ((wv[(t + 8) % MPL] >> (2 * i)) & 3u) << (4 * t + 1);
      }
      if (MPL == 16) *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
      else *reinterpret_cast<uint32_t*>(dst) = lo;
    }
  }
}

// csg[job][j] = colsum_job[idx[j]] (zero padded to kstride) and SQ[job] = { sum_j csg, sum_j csg^2 }.
// SPLIT8 (fp4 panel, column sums below 65 536): instead of ints, csg receives the sums as two unsigned bytes per
// marker, c = lo + 256 hi, grouped per 8 markers (= one 32-bit panel word) as
//   [lo of markers 0,2,4,6 | lo of 1,3,5,7 | hi of 0,2,4,6 | hi of 1,3,5,7]      (16 bytes)
// so that centre_rows_fp4_kernel can use dp4a on the even / odd nibbles of the word.
template <bool SPLIT8>
__global__ void __launch_bounds__(256) centre_sq_kernel(const int* __restrict__ idx, const long long* __restrict__ off,
                                                        int w0, int n_slots, int kstride,
                                                        const int* const* __restrict__ colsum_of,
                                                        int* __restrict__ csg, long long* __restrict__ SQ) {
  __shared__ long long sh[2][8];
  const int job = blockIdx.x, w = job / n_slots;
  const long long o0 = off[w0 + w];
  const int k = (int)(off[w0 + w + 1] - o0);
  const int* cs = colsum_of[job];
  int* out = csg + (size_t)job * kstride;
  long long S = 0, Q = 0;
  for (int j = threadIdx.x; j < kstride; j += blockDim.x) {
    int c = 0;
    if (j < k) c = cs[idx[o0 + j]];
    if (SPLIT8) {
      unsigned char* o8 = reinterpret_cast<unsigned char*>(out) + 16 * (j >> 3) + ((j & 1) << 2) + ((j & 7) >> 1);
      o8[0] = (unsigned char)(c & 255);
      o8[8] = (unsigned char)(c >> 8);
    } else {
      out[j] = c;
    }
    S += c;
    Q += (long long)c * c;
  }
  for (int o = 16; o; o >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, o);
    Q += __shfl_xor_sync(0xffffffffu, Q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = S;
    sh[1][threadIdx.x >> 5] = Q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    S = Q = 0;
    for (int i = 0; i < 8; ++i) {
      S += sh[0][i];
      Q += sh[1][i];
    }
    SQ[2 * job] = S;
    SQ[2 * job + 1] = Q;
  }
}

// s[a] = sum_j panel[a][j] * csg[j]  (exact, int64).  One warp per FOUR animal rows: the gathered column sums are
// four times the bytes of the panel row they multiply, so sharing each csg load between four rows takes the kernel
// from L1-bandwidth-bound to HBM-bound (panel streamed once from HBM, csg from L1/L2).
__global__ void __launch_bounds__(256) centre_rows_kernel(const int8_t* __restrict__ panel, int rpad, int kstride,
                                                          const int* __restrict__ kblocks, int n_slots,
                                                          const int* __restrict__ csg, long long* __restrict__ s) {
  const int job = blockIdx.y;  // w * n_slots + slot
  const int w = job / n_slots;
  const int a0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4, lane = threadIdx.x & 31;
  if (a0 >= rpad) return;      // rpad is a multiple of 128: rows a0 .. a0 + 3 are all inside
  const int n16 = kblocks[w] * TB_GRAM_BK / 16;     // panel and csg are zero beyond k
  const uint4* row = reinterpret_cast<const uint4*>(panel + ((size_t)w * rpad + a0) * kstride);
  const size_t rs = kstride / 16;
  const int4* cg = reinterpret_cast<const int4*>(csg + (size_t)job * kstride);
  long long acc[4] = {0, 0, 0, 0};
  auto dot16 = [](const uint4 v, const int4 c0, const int4 c1, const int4 c2, const int4 c3) -> int {
    int t = 0;
    t += (int)(v.x & 0xff) * c0.x + (int)((v.x >> 8) & 0xff) * c0.y + (int)((v.x >> 16) & 0xff) * c0.z + (int)(v.x >> 24) * c0.w;
    t += (int)(v.y & 0xff) * c1.x + (int)((v.y >> 8) & 0xff) * c1.y + (int)((v.y >> 16) & 0xff) * c1.z + (int)(v.y >> 24) * c1.w;
    t += (int)(v.z & 0xff) * c2.x + (int)((v.z >> 8) & 0xff) * c2.y + (int)((v.z >> 16) & 0xff) * c2.z + (int)(v.z >> 24) * c2.w;
    t += (int)(v.w & 0xff) * c3.x + (int)((v.w >> 8) & 0xff) * c3.y + (int)((v.w >> 16) & 0xff) * c3.z + (int)(v.w >> 24) * c3.w;
    return t;   // 16 products of (dosage <= 2) x (column sum <= 2n < 2^26) fit an int
  };
  for (int j = lane; j < n16; j += 32) {
    const uint4 v0 = row[j], v1 = row[rs + j], v2 = row[2 * rs + j], v3 = row[3 * rs + j];
    const int4 c0 = cg[4 * j], c1 = cg[4 * j + 1], c2 = cg[4 * j + 2], c3 = cg[4 * j + 3];
    acc[0] += dot16(v0, c0, c1, c2, c3);
    acc[1] += dot16(v1, c0, c1, c2, c3);
    acc[2] += dot16(v2, c0, c1, c2, c3);
    acc[3] += dot16(v3, c0, c1, c2, c3);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    long long v = acc[r];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s[(size_t)job * rpad + a0 + r] = v;
  }
}

// centre_rows_kernel for the fp4 panel: 16 bytes = 32 markers (nibble = 2 x dosage), kstride counts markers.
__global__ void __launch_bounds__(256) centre_rows_fp4_kernel(const int8_t* __restrict__ panel, int rpad, int kstride,
                                                              const int* __restrict__ kblocks, int n_slots,
                                                              const int* __restrict__ csg, long long* __restrict__ s) {
  const int job = blockIdx.y;
  const int w = job / n_slots;
  const int a0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4, lane = threadIdx.x & 31;
  if (a0 >= rpad) return;
  const int kstride_b = kstride / 2;
  const int n16 = kblocks[w] * TB_GRAM_BK / 16;          // a k-block is 128 bytes in either layout
  const uint4* row = reinterpret_cast<const uint4*>(panel + ((size_t)w * rpad + a0) * kstride_b);
  const size_t rs = kstride_b / 16;
  const int4* cg = reinterpret_cast<const int4*>(csg + (size_t)job * kstride);
  long long acc[4] = {0, 0, 0, 0};
  auto dot8 = [](const uint32_t v, const int4 c0, const int4 c1) -> int {
    return (int)((v >> 1) & 3u) * c0.x + (int)((v >> 5) & 3u) * c0.y + (int)((v >> 9) & 3u) * c0.z +
           (int)((v >> 13) & 3u) * c0.w + (int)((v >> 17) & 3u) * c1.x + (int)((v >> 21) & 3u) * c1.y +
           (int)((v >> 25) & 3u) * c1.z + (int)((v >> 29) & 3u) * c1.w;
  };
  for (int j = lane; j < n16; j += 32) {
    const uint4 v0 = row[j], v1 = row[rs + j], v2 = row[2 * rs + j], v3 = row[3 * rs + j];
    const uint32_t a[4][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w},
                              {v3.x, v3.y, v3.z, v3.w}};
#pragma unroll
    for (int q = 0; q < 4; ++q) {                        // 8 markers per 32-bit word
      const int4 c0 = cg[8 * j + 2 * q], c1 = cg[8 * j + 2 * q + 1];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += dot8(a[r][q], c0, c1);   // 8 x (<= 2) x (< 2^26) fits an int
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    long long v = acc[r];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s[(size_t)job * rpad + a0 + r] = v;
  }
}

// The same with dp4a: the column sums arrive as split bytes (centre_sq_kernel<true>), the even / odd nibbles of a panel
// word become two words of dosage bytes with one shift + mask each, and 8 markers cost 4 dp4a instead of 24 ALU ops.
__global__ void __launch_bounds__(256) centre_rows_fp4_dp4a_kernel(const int8_t* __restrict__ panel, int rpad, int kstride,
                                                                   const int* __restrict__ kblocks, int n_slots,
                                                                   const int* __restrict__ csg, long long* __restrict__ s) {
  const int job = blockIdx.y;
  const int w = job / n_slots;
  const int a0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4, lane = threadIdx.x & 31;
  if (a0 >= rpad) return;
  const int kstride_b = kstride / 2;
  const int n16 = kblocks[w] * TB_GRAM_BK / 16;
  const uint4* row = reinterpret_cast<const uint4*>(panel + ((size_t)w * rpad + a0) * kstride_b);
  const size_t rs = kstride_b / 16;
  const uint4* cg = reinterpret_cast<const uint4*>(csg + (size_t)job * kstride);      // 16 bytes per 8 markers
  unsigned int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
  for (int j = lane; j < n16; j += 32) {
    const uint4 v0 = row[j], v1 = row[rs + j], v2 = row[2 * rs + j], v3 = row[3 * rs + j];
    const uint32_t a[4][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w},
                              {v3.x, v3.y, v3.z, v3.w}};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 g = cg[4 * j + q];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const unsigned int ev = (a[r][q] >> 1) & 0x03030303u, od = (a[r][q] >> 5) & 0x03030303u;
        lo[r] = __dp4a(ev, g.x, lo[r]);
        lo[r] = __dp4a(od, g.y, lo[r]);
        hi[r] = __dp4a(ev, g.z, hi[r]);
        hi[r] = __dp4a(od, g.w, hi[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    long long v = (long long)lo[r] + 256LL * (long long)hi[r];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s[(size_t)job * rpad + a0 + r] = v;
  }
}

}  // namespace

cudaError_t tb_gather_init() { return cudaSuccess; }

cudaError_t tb_launch_gather_fp4(const TbGeno& g, const int* d_idx, const long long* d_off, int w0, int W, int rpad,
                                 int kstride_b, int8_t* d_panel, cudaStream_t st, const int* d_rowmap, int rows_univ) {
  if (!g.x2) return cudaErrorInvalidValue;       // written for the packed matrix
  // 256-marker tiles (32 KiB, more blocks in flight) measured 3.5 ms against 4.3 ms for 512-marker tiles (64 KiB)
  dim3 grid((2 * kstride_b + 255) / 256, ((d_rowmap ? rows_univ : rpad) + 511) / 512, W);
  gather_fp4_kernel<8><<<grid, 256, 256 * 128, st>>>(g.x2, g.ld4, d_idx, d_off, w0, rpad, kstride_b, d_panel, d_rowmap,
                                                     rows_univ);
  return cudaGetLastError();
}

cudaError_t tb_launch_gather(const TbGeno& g, const int* d_idx, const long long* d_off, int w0, int W,
                             int rpad, int kstride, int8_t* d_panel, cudaStream_t st) {
  dim3 grid(kstride / 128, rpad / 128, W);
  if (g.x2) {
    constexpr int MPL = 8, TM = 32 * MPL;
    dim3 pgrid((kstride + TM - 1) / TM, (rpad + 511) / 512, W);
    gather_packed_kernel<MPL><<<pgrid, 256, TM * 128, st>>>(g.x2, g.ld4, d_idx, d_off, w0, rpad, kstride, d_panel);
  } else
    gather_kernel<<<grid, 256, 0, st>>>(g.x, g.ldn, d_idx, d_off, w0, rpad, kstride, d_panel);
  return cudaGetLastError();
}

cudaError_t tb_launch_centre_terms(const int8_t* d_panel, int rpad, int kstride, const int* d_idx,
                                   const long long* d_off, int w0, int W, int n_slots, const int* d_kblocks,
                                   const int* const* d_colsum_of, int* d_csg, long long* d_s, long long* d_SQ,
                                   cudaStream_t st, int fp4) {
  // fp4 == 2: column sums fit 16 bits (at most 32 767 animals behind the frequencies) -> split bytes + dp4a
  if (fp4 == 2)
    centre_sq_kernel<true><<<W * n_slots, 256, 0, st>>>(d_idx, d_off, w0, n_slots, kstride, d_colsum_of, d_csg, d_SQ);
  else
    centre_sq_kernel<false><<<W * n_slots, 256, 0, st>>>(d_idx, d_off, w0, n_slots, kstride, d_colsum_of, d_csg, d_SQ);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid((rpad + 31) / 32, W * n_slots);
  if (fp4 == 2)
    centre_rows_fp4_dp4a_kernel<<<grid, 256, 0, st>>>(d_panel, rpad, kstride, d_kblocks, n_slots, d_csg, d_s);
  else if (fp4)
    centre_rows_fp4_kernel<<<grid, 256, 0, st>>>(d_panel, rpad, kstride, d_kblocks, n_slots, d_csg, d_s);
  else
    centre_rows_kernel<<<grid, 256, 0, st>>>(d_panel, rpad, kstride, d_kblocks, n_slots, d_csg, d_s);
  return cudaGetLastError();
}
