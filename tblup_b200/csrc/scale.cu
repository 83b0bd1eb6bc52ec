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

__global__ void __launch_bounds__(256) scale_kernel(const TbScaleJob* __restrict__ jobs) {
  const TbScaleJob jb = jobs[blockIdx.z];
  const int ntp = jb.ntp, n_t = jb.n_t, n_v = jb.n_v;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (r0 >= ntp + n_v || c0 >= ntp) return;
  if (r0 < ntp && c0 > r0 + 31) return;            // strictly above the diagonal of A
  const int c = c0 + (threadIdx.x & 31);
  const long long N = jb.N, S = jb.SQ[0], Q = jb.SQ[1];
  const double den = (double)(2 * N * S - Q);
  const bool c_real = c < n_t;
  int pc = 0;
  long long sc = 0;
  if (c_real) {
    pc = jb.tpos[c];
    sc = jb.s[pc];
  }
  for (int rr = threadIdx.x >> 5; rr < 32; rr += 8) {
    const int r = r0 + rr;
    if (r >= ntp + n_v) break;
    double out;
    if (r < ntp) {
      if (c > r) continue;
      if (r >= n_t || !c_real) {
        out = (r == c) ? 1.0 : 0.0;
      } else {
        const int pr = jb.tpos[r];
        const int hi = pr > pc ? pr : pc, lo = pr > pc ? pc : pr;
        const long long cv = jb.C[(size_t)hi * jb.rpad + lo];
        const long long num = N * N * cv - N * (jb.s[pr] + sc) + Q;
        out = 2.0 * (double)num / den;
        if (r == c) out += jb.lambda;
      }
    } else {
      if (!c_real) {
        out = 0.0;
      } else {
        const int pr = jb.vpos[r - ntp];
        const int hi = pr > pc ? pr : pc, lo = pr > pc ? pc : pr;
        const long long cv = jb.C[(size_t)hi * jb.rpad + lo];
        const long long num = N * N * cv - N * (jb.s[pr] + sc) + Q;
        out = 2.0 * (double)num / den;
      }
    }
    jb.M[(size_t)r * ntp + c] = out;
  }
}

}  // namespace

cudaError_t tb_launch_scale(const TbScaleJob* d_jobs, int n_jobs, int max_rows, int max_ntp, cudaStream_t st) {
  dim3 grid((max_ntp + 31) / 32, (max_rows + 31) / 32, n_jobs);
  scale_kernel<<<grid, 256, 0, st>>>(d_jobs);
  return cudaGetLastError();
}
