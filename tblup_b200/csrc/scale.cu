// Exact centring + scaling: integer cross-products -> VanRaden relationship blocks in fp64.
//
//   G_ab = 2 (N^2 C_ab - N (s_a + s_b) + Q) / (2 N S - Q)
//
// which is W W^T / (2 sum p(1-p)) of tblup/utils.py:14-18 with W = X - 2p and 2p_j = colsum_j / N:
// numerator and denominator are exact int64, the only rounding is the final division (the oracle's
// exact_grm_block does the same operations in the same order, so A is bit-identical to it).
// Output per (individual, row set): M = [ A ; G_vt ] with A = G_tt + lambda I (lower triangle only,
// tblup/evaluator.py:280-281) padded to a multiple of the Cholesky block with an identity block.
// HBM-bound: reads 4 B and writes 8 B per matrix entry.
#include "tb_internal.h"

namespace {

__device__ __forceinline__ double grm_entry(long long cv, long long s_sum, long long N, long long Q, double den) {
  const long long num = N * N * cv - N * s_sum + Q;
  return 2.0 * (double)num / den;
}

// One thread = one row x 4 consecutive columns; a block covers 16 rows x 128 columns (2 rows per thread).
// Fast path: the four column animals sit at consecutive, 16-byte aligned universe positions strictly below the
// row animal's position -> one 16-byte load of C; otherwise element-wise with the (max, min) lookup.
__global__ void __launch_bounds__(256) scale_kernel(const TbScaleJob* __restrict__ jobs) {
  const TbScaleJob jb = jobs[blockIdx.z];
  const int ntp = jb.ntp, n_t = jb.n_t, n_v = jb.n_v, rpad = jb.rpad;
  const int r0 = blockIdx.y * 16, c0 = blockIdx.x * 128;
  if (r0 >= ntp + n_v || c0 >= ntp) return;
  if (r0 < ntp && c0 > r0 + 15) return;            // strictly above the diagonal of A
  const int c = c0 + (threadIdx.x & 31) * 4;
  if (c >= ntp) return;
  const long long N = jb.N, S = jb.SQ[0], Q = jb.SQ[1];
  const double den = (double)(2 * N * S - Q);
  int pc[4];
  long long sc[4];
  bool creal[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    creal[i] = c + i < n_t;
    pc[i] = creal[i] ? jb.tpos[c + i] : 0;
    sc[i] = creal[i] ? jb.s[pc[i]] : 0;
  }
  const bool run = creal[3] && pc[1] == pc[0] + 1 && pc[2] == pc[0] + 2 && pc[3] == pc[0] + 3 && (pc[0] & 3) == 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + (threadIdx.x >> 5) + 8 * h;
    if (r >= ntp + n_v) break;
    const bool in_a = r < ntp;
    if (in_a && c > r) continue;
    double out[4];
    const bool r_real = in_a ? (r < n_t) : true;
    if (!r_real) {
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = (r == c + i) ? 1.0 : 0.0;
    } else {
      const int pr = in_a ? jb.tpos[r] : jb.vpos[r - ntp];
      const long long sr = jb.s[pr];
      const int32_t* crow = jb.C + (size_t)pr * rpad;
      if (run && pc[3] < pr) {
        const int4 cv = *reinterpret_cast<const int4*>(crow + pc[0]);
        out[0] = grm_entry(cv.x, sr + sc[0], N, Q, den);
        out[1] = grm_entry(cv.y, sr + sc[1], N, Q, den);
        out[2] = grm_entry(cv.z, sr + sc[2], N, Q, den);
        out[3] = grm_entry(cv.w, sr + sc[3], N, Q, den);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!creal[i]) {
            out[i] = 0.0;
          } else {
            const int hi = pr > pc[i] ? pr : pc[i], lo = pr > pc[i] ? pc[i] : pr;
            out[i] = grm_entry(jb.C[(size_t)hi * rpad + lo], sr + sc[i], N, Q, den);
          }
        }
      }
      if (in_a) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (r == c + i) out[i] += jb.lambda;
      }
    }
    double2* dst = reinterpret_cast<double2*>(jb.M + (size_t)r * ntp + c);
    dst[0] = make_double2(out[0], out[1]);
    dst[1] = make_double2(out[2], out[3]);
  }
}

}  // namespace

cudaError_t tb_launch_scale(const TbScaleJob* d_jobs, int n_jobs, int max_rows, int max_ntp, cudaStream_t st) {
  dim3 grid((max_ntp + 127) / 128, (max_rows + 15) / 16, n_jobs);
  scale_kernel<<<grid, 256, 0, st>>>(d_jobs);
  return cudaGetLastError();
}
