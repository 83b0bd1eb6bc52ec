// Triangular solves, validation predictions and the accuracy reduction, fused: one CTA per
// (individual, row set).
//
//   z = L^-1 y_t,  alpha = L^-T z            (alpha = (G_tt + lambda I)^-1 y_t, tblup/evaluator.py:282-284)
//   pred_v = G_vt alpha                      (only the validation rows of evaluator.py:284 are ever used)
//   fitness = | pearson(y_v, pred_v) |       (evaluator.py:286 / :314; scipy conventions: clip to [-1, 1],
//                                             NaN when either vector is constant)
// Blocked substitution with the stored inverses of the diagonal blocks, so each step is two small
// mat-vecs.  HBM-bound: streams L twice and G_vt once with 16-byte coalesced loads.
#include "tb_internal.h"

namespace {

constexpr int NB = TB_NB;
constexpr int ST = 512;   // threads

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < ST / 32; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(ST) solve_kernel(const TbSolveJob* __restrict__ jobs) {
  extern __shared__ double ssm[];
  const TbSolveJob jb = jobs[blockIdx.x];
  const int ntp = jb.ntp, n_v = jb.n_v, nb = ntp / NB;
  double* z = ssm;                 // [ntp]   z, then alpha in place
  double* rvec = z + ntp;          // [NB]
  double* part = rvec + NB;        // [8][NB]
  double* red = part + 8 * NB;     // [ST / 32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* M = jb.M;

  // ---------------- forward: L z = y ----------------
  for (int b = 0; b < nb; ++b) {
    const int kc = b * NB;
    for (int i = warp; i < NB; i += ST / 32) {
      const double2* row = reinterpret_cast<const double2*>(M + (size_t)(kc + i) * ntp);
      const double2* zz = reinterpret_cast<const double2*>(z);
      double s0 = 0.0, s1 = 0.0;
      int c = lane;
      for (; c + 32 < kc / 2; c += 64) {
        const double2 l0 = row[c], l1 = row[c + 32];
        const double2 z0 = zz[c], z1 = zz[c + 32];
        s0 += l0.x * z0.x + l0.y * z0.y;
        s1 += l1.x * z1.x + l1.y * z1.y;
      }
      for (; c < kc / 2; c += 32) {
        const double2 l0 = row[c];
        const double2 z0 = zz[c];
        s0 += l0.x * z0.x + l0.y * z0.y;
      }
      const double s = warp_sum(s0 + s1);
      if (lane == 0) rvec[i] = jb.y_t[kc + i] - s;
    }
    __syncthreads();
    {
      // z_b = Linv_b rvec   (8 threads per row, 8 columns each)
      const int i = tid >> 3, sub = tid & 7;
      const double* li = jb.Linv + (size_t)b * NB * NB + i * NB + sub * 8;
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < 8; ++p) s += li[p] * rvec[sub * 8 + p];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (sub == 0) z[kc + i] = s;
    }
    __syncthreads();
  }

  // ---------------- backward: L^T alpha = z ----------------
  for (int b = nb - 1; b >= 0; --b) {
    const int kc = b * NB;
    {
      const int c = tid & 63, grp = tid >> 6;
      double s0 = 0.0, s1 = 0.0;
      int i = kc + NB + grp;
      for (; i + 8 < ntp; i += 16) {
        s0 += M[(size_t)i * ntp + kc + c] * z[i];
        s1 += M[(size_t)(i + 8) * ntp + kc + c] * z[i + 8];
      }
      for (; i < ntp; i += 8) s0 += M[(size_t)i * ntp + kc + c] * z[i];
      part[grp * NB + c] = s0 + s1;
    }
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int gI = 0; gI < 8; ++gI) s += part[gI * NB + tid];
      rvec[tid] = z[kc + tid] - s;
    }
    __syncthreads();
    {
      // alpha_b = Linv_b^T rvec : alpha[i] = sum_{p >= i} Linv[p][i] rvec[p]
      const int i = tid & 63, grp = tid >> 6;
      const double* li = jb.Linv + (size_t)b * NB * NB;
      double s = 0.0;
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const int p = grp * 8 + pp;
        s += li[p * NB + i] * rvec[p];
      }
      part[grp * NB + i] = s;
    }
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int gI = 0; gI < 8; ++gI) s += part[gI * NB + tid];
      z[kc + tid] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < ntp; i += ST) jb.alpha[i] = z[i];

  // ---------------- predictions on the validation animals ----------------
  const double* V = M + (size_t)ntp * ntp;
  for (int v = warp; v < n_v; v += ST / 32) {
    const double2* row = reinterpret_cast<const double2*>(V + (size_t)v * ntp);
    const double2* aa = reinterpret_cast<const double2*>(z);
    double s0 = 0.0, s1 = 0.0;
    int c = lane;
    for (; c + 32 < ntp / 2; c += 64) {
      const double2 l0 = row[c], l1 = row[c + 32];
      const double2 a0 = aa[c], a1 = aa[c + 32];
      s0 += l0.x * a0.x + l0.y * a0.y;
      s1 += l1.x * a1.x + l1.y * a1.y;
    }
    for (; c < ntp / 2; c += 32) {
      const double2 l0 = row[c];
      const double2 a0 = aa[c];
      s0 += l0.x * a0.x + l0.y * a0.y;
    }
    const double s = warp_sum(s0 + s1);
    if (lane == 0) jb.pred[v] = s;
  }
  __syncthreads();

  // ---------------- |Pearson r| ----------------
  double sy = 0.0, sp = 0.0;
  for (int v = tid; v < n_v; v += ST) {
    sy += jb.y_v[v];
    sp += jb.pred[v];
  }
  const double my = block_sum(sy, red) / n_v;
  const double mp = block_sum(sp, red) / n_v;
  double sxy = 0.0, sxx = 0.0, syy = 0.0;
  for (int v = tid; v < n_v; v += ST) {
    const double dy = jb.y_v[v] - my, dp = jb.pred[v] - mp;
    sxy += dy * dp;
    sxx += dy * dy;
    syy += dp * dp;
  }
  sxy = block_sum(sxy, red);
  sxx = block_sum(sxx, red);
  syy = block_sum(syy, red);
  if (tid == 0) {
    double r;
    if (*jb.status != 0 || !(sxx > 0.0) || !(syy > 0.0)) {
      r = __longlong_as_double(0x7ff8000000000000LL);
    } else {
      r = sxy / (sqrt(sxx) * sqrt(syy));
      // fmin / fmax drop NaN operands: a non-finite r (syy = Inf) must not be clipped to 1
      r = (fabs(r) < 1e300) ? fabs(fmax(fmin(r, 1.0), -1.0)) : __longlong_as_double(0x7ff8000000000000LL);
    }
    *jb.fitness = r;
  }
}

int g_solve_smem_max = 0;

}  // namespace

static inline int solve_smem_bytes(int ntp) { return (ntp + NB + 8 * NB + ST / 32) * (int)sizeof(double); }

cudaError_t tb_solve_init() {
  g_solve_smem_max = 200 * 1024;
  return cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_solve_smem_max);
}

cudaError_t tb_launch_solve(const TbSolveJob* d_jobs, int n_jobs, int max_ntp, cudaStream_t st) {
  const int smem = solve_smem_bytes(max_ntp);
  if (smem > g_solve_smem_max) return cudaErrorInvalidConfiguration;
  solve_kernel<<<n_jobs, ST, smem, st>>>(d_jobs);
  return cudaGetLastError();
}
