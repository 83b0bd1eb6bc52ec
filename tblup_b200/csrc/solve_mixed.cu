// Solve + predict + accuracy for the mixed-precision path: the TF32 Cholesky factor of chol_tc.cu is used as a
// preconditioner and the solution is refined in fp64 against the EXACT operator, applied straight from the
// integer cross-products (no fp64 copy of A is ever formed):
//
//   (A alpha)_a = lambda alpha_a + (2/den) [ N^2 (C alpha)_a - N s_a (sum alpha) - N (s . alpha) + Q (sum alpha) ]
//
// with C, s, S, Q the integers of scale.cu / DESIGN.md §1 and den = 2 N S - Q.  Iteration:
//   alpha <- M^-1 y;  repeat { r = y - A alpha;  d = M^-1 r;  alpha += d } until max|d| <= 1e-11 max|alpha|
// where M^-1 = (L L^T)^-1 by blocked substitution (fp32 factor, fp64 accumulation).  Then
//   pred_v = (G_vt alpha)_v from the integer rows of the validation animals, fitness = |pearson(y_v, pred_v)|.
// Same reference lines as solve.cu (tblup/evaluator.py:282-286, :311-314).  One CTA per (individual, row set);
// HBM-bound: per sweep the factor is streamed twice (fp32) and the lower triangle of C twice (int32).
#include "tb_internal.h"

namespace {

constexpr int NB = TB_NB;
constexpr int ST = 512;
constexpr int MAX_SWEEPS = 8;

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
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
__device__ double block_max(double v, double* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < ST / 32; ++i) t = fmax(t, red[i]);
  return t;
}

// work <- (L L^T)^-1 work, in place.
__device__ void apply_minv(const float* __restrict__ L, const float* __restrict__ Linv, int ntp, double* work,
                           double* rvec, double* part) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = ntp / NB;
  for (int b = 0; b < nb; ++b) {                      // forward: L z = work
    const int kc = b * NB;
    for (int i = warp; i < NB; i += ST / 32) {
      const float4* row = reinterpret_cast<const float4*>(L + (size_t)(kc + i) * ntp);
      double s0 = 0.0, s1 = 0.0;
      int c = lane;
      for (; c + 32 < kc / 4; c += 64) {
        const float4 l0 = row[c], l1 = row[c + 32];
        const double* z0 = work + 4 * c;
        const double* z1 = work + 4 * (c + 32);
        s0 += (double)l0.x * z0[0] + (double)l0.y * z0[1] + (double)l0.z * z0[2] + (double)l0.w * z0[3];
        s1 += (double)l1.x * z1[0] + (double)l1.y * z1[1] + (double)l1.z * z1[2] + (double)l1.w * z1[3];
      }
      for (; c < kc / 4; c += 32) {
        const float4 l0 = row[c];
        const double* z0 = work + 4 * c;
        s0 += (double)l0.x * z0[0] + (double)l0.y * z0[1] + (double)l0.z * z0[2] + (double)l0.w * z0[3];
      }
      const double s = warp_sum(s0 + s1);
      if (lane == 0) rvec[i] = work[kc + i] - s;
    }
    __syncthreads();
    {
      const int i = tid >> 3, sub = tid & 7;
      const float* li = Linv + ((size_t)kc + i) * NB + sub * 8;
      double s = 0.0;
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) s += (double)li[pp] * rvec[sub * 8 + pp];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (sub == 0) work[kc + i] = s;
    }
    __syncthreads();
  }
  for (int b = nb - 1; b >= 0; --b) {                 // backward: L^T d = z
    const int kc = b * NB;
    {
      const int c = tid & 63, grp = tid >> 6;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int i = kc + NB + grp;
      for (; i + 24 < ntp; i += 32) {
        s0 += (double)L[(size_t)i * ntp + kc + c] * work[i];
        s1 += (double)L[(size_t)(i + 8) * ntp + kc + c] * work[i + 8];
        s2 += (double)L[(size_t)(i + 16) * ntp + kc + c] * work[i + 16];
        s3 += (double)L[(size_t)(i + 24) * ntp + kc + c] * work[i + 24];
      }
      for (; i < ntp; i += 8) s0 += (double)L[(size_t)i * ntp + kc + c] * work[i];
      part[grp * NB + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int gI = 0; gI < 8; ++gI) s += part[gI * NB + tid];
      rvec[tid] = work[kc + tid] - s;
    }
    __syncthreads();
    {
      const int i = tid & 63, grp = tid >> 6;
      const float* li = Linv + (size_t)kc * NB;
      double s = 0.0;
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const int q = grp * 8 + pp;
        s += (double)li[q * NB + i] * rvec[q];
      }
      part[grp * NB + i] = s;
    }
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int gI = 0; gI < 8; ++gI) s += part[gI * NB + tid];
      work[kc + tid] = s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ST) solve_mixed_kernel(const TbSolveMixedJob* __restrict__ jobs) {
  extern __shared__ double msm[];
  const TbSolveMixedJob jb = jobs[blockIdx.x];
  const int ntp = jb.ntp, n_t = jb.n_t, n_v = jb.n_v, rpad = jb.rpad;
  double* alpha = msm;               // [ntp]
  double* work = alpha + ntp;        // [ntp]
  double* sT = work + ntp;           // [ntp] s at the training positions
  double* rvec = sT + ntp;           // [NB]
  double* part = rvec + NB;          // [8][NB]
  double* red = part + 8 * NB;       // [ST/32]
  int* tp = reinterpret_cast<int*>(red + ST / 32);   // [ntp]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double Nd = (double)jb.N, Sd = (double)jb.SQ[0], Qd = (double)jb.SQ[1];
  const double coef = 2.0 / (2.0 * Nd * Sd - Qd);
  const int32_t* C = jb.C;

  for (int a = tid; a < ntp; a += ST) {
    const bool real = a < n_t;
    const int pa = real ? jb.tpos[a] : 0;
    tp[a] = pa;
    sT[a] = real ? (double)jb.s[pa] : 0.0;
    work[a] = jb.y_t[a];
  }
  __syncthreads();
  apply_minv(jb.L32, jb.Linv32, ntp, work, rvec, part);
  for (int a = tid; a < ntp; a += ST) alpha[a] = work[a];
  __syncthreads();

  int sweeps = 0;
  double sa = 0.0, ssa = 0.0;
  for (;;) {
    double l0 = 0.0, l1 = 0.0;
    for (int a = tid; a < n_t; a += ST) {
      l0 += alpha[a];
      l1 += sT[a] * alpha[a];
    }
    sa = block_sum(l0, red);
    ssa = block_sum(l1, red);
    if (sweeps == MAX_SWEEPS) break;
    // ---- r = y - A alpha : rows (b <= a) ...
    for (int a = warp; a < n_t; a += ST / 32) {
      const int pa = tp[a];
      double d0 = 0.0, d1 = 0.0;
      int b = lane;
      for (; b + 32 <= a; b += 64) {
        const int p0 = tp[b], p1 = tp[b + 32];
        const int c0 = C[(size_t)(pa > p0 ? pa : p0) * rpad + (pa > p0 ? p0 : pa)];
        const int c1 = C[(size_t)(pa > p1 ? pa : p1) * rpad + (pa > p1 ? p1 : pa)];
        d0 += (double)c0 * alpha[b];
        d1 += (double)c1 * alpha[b + 32];
      }
      for (; b <= a; b += 32) {
        const int p0 = tp[b];
        d0 += (double)C[(size_t)(pa > p0 ? pa : p0) * rpad + (pa > p0 ? p0 : pa)] * alpha[b];
      }
      const double d = warp_sum(d0 + d1);
      if (lane == 0) work[a] = d;
    }
    __syncthreads();
    // ---- ... and columns (b > a): consecutive threads own consecutive columns, rows are streamed
    for (int a0 = 0; a0 < n_t; a0 += ST) {
      const int a = a0 + tid;
      if (a < n_t) {
        const int pa = tp[a];
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        int b = a + 1;
        for (; b + 3 < n_t; b += 4) {
          const int p0 = tp[b], p1 = tp[b + 1], p2 = tp[b + 2], p3 = tp[b + 3];
          const int c0 = C[(size_t)(pa > p0 ? pa : p0) * rpad + (pa > p0 ? p0 : pa)];
          const int c1 = C[(size_t)(pa > p1 ? pa : p1) * rpad + (pa > p1 ? p1 : pa)];
          const int c2 = C[(size_t)(pa > p2 ? pa : p2) * rpad + (pa > p2 ? p2 : pa)];
          const int c3 = C[(size_t)(pa > p3 ? pa : p3) * rpad + (pa > p3 ? p3 : pa)];
          d0 += (double)c0 * alpha[b];
          d1 += (double)c1 * alpha[b + 1];
          d2 += (double)c2 * alpha[b + 2];
          d3 += (double)c3 * alpha[b + 3];
        }
        for (; b < n_t; ++b) {
          const int p0 = tp[b];
          d0 += (double)C[(size_t)(pa > p0 ? pa : p0) * rpad + (pa > p0 ? p0 : pa)] * alpha[b];
        }
        const double ca = work[a] + (d0 + d1) + (d2 + d3);
        const double Aa = jb.lambda * alpha[a] + coef * (Nd * Nd * ca - Nd * sT[a] * sa - Nd * ssa + Qd * sa);
        work[a] = jb.y_t[a] - Aa;
      }
    }
    for (int a = n_t + tid; a < ntp; a += ST) work[a] = 0.0;
    __syncthreads();
    apply_minv(jb.L32, jb.Linv32, ntp, work, rvec, part);
    double dmax = 0.0, amax = 0.0;
    for (int a = tid; a < ntp; a += ST) {
      const double d = work[a];
      const double v = alpha[a] + d;
      alpha[a] = v;
      dmax = fmax(dmax, fabs(d));
      amax = fmax(amax, fabs(v));
    }
    dmax = block_max(dmax, red);
    amax = block_max(amax, red);
    ++sweeps;
    if (!(dmax > 1e-11 * amax)) {
      // converged: one more pass through the loop head refreshes sa / ssa for the prediction, then leave
      double m0 = 0.0, m1 = 0.0;
      for (int a = tid; a < n_t; a += ST) {
        m0 += alpha[a];
        m1 += sT[a] * alpha[a];
      }
      sa = block_sum(m0, red);
      ssa = block_sum(m1, red);
      break;
    }
  }
  for (int a = tid; a < ntp; a += ST) jb.alpha[a] = alpha[a];
  if (tid == 0 && jb.sweeps) *jb.sweeps = sweeps;

  // ---- predictions on the validation animals
  for (int v = warp; v < n_v; v += ST / 32) {
    const int pv = jb.vpos[v];
    double d0 = 0.0, d1 = 0.0;
    int b = lane;
    for (; b + 32 < n_t; b += 64) {
      const int p0 = tp[b], p1 = tp[b + 32];
      const int c0 = C[(size_t)(pv > p0 ? pv : p0) * rpad + (pv > p0 ? p0 : pv)];
      const int c1 = C[(size_t)(pv > p1 ? pv : p1) * rpad + (pv > p1 ? p1 : pv)];
      d0 += (double)c0 * alpha[b];
      d1 += (double)c1 * alpha[b + 32];
    }
    for (; b < n_t; b += 32) {
      const int p0 = tp[b];
      d0 += (double)C[(size_t)(pv > p0 ? pv : p0) * rpad + (pv > p0 ? p0 : pv)] * alpha[b];
    }
    const double d = warp_sum(d0 + d1);
    if (lane == 0) jb.pred[v] = coef * (Nd * Nd * d - Nd * (double)jb.s[pv] * sa - Nd * ssa + Qd * sa);
  }
  __syncthreads();

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
      r = fabs(fmax(fmin(r, 1.0), -1.0));
    }
    *jb.fitness = r;
  }
}

// fp32 copy of A = G_tt + lambda I (lower triangle, identity padding) for the tensor-core factorisation.
__global__ void __launch_bounds__(256) scale32_kernel(const TbScaleJob* __restrict__ jobs, float* __restrict__ L32,
                                                      int ntp_all) {
  const TbScaleJob jb = jobs[blockIdx.z];
  const int ntp = jb.ntp, n_t = jb.n_t, rpad = jb.rpad;
  const int r0 = blockIdx.y * 16, c0 = blockIdx.x * 128;
  if (r0 >= ntp || c0 >= ntp || c0 > r0 + 15) return;
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
  float* out_base = L32 + (size_t)blockIdx.z * ntp_all * ntp_all;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + (threadIdx.x >> 5) + 8 * h;
    if (r >= ntp) break;
    if (c > r) continue;
    float out[4];
    if (r >= n_t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = (r == c + i) ? 1.f : 0.f;
    } else {
      const int pr = jb.tpos[r];
      const long long sr = jb.s[pr];
      int cv[4];
      if (run && pc[3] < pr) {
        const int4 q4 = *reinterpret_cast<const int4*>(jb.C + (size_t)pr * rpad + pc[0]);
        cv[0] = q4.x; cv[1] = q4.y; cv[2] = q4.z; cv[3] = q4.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int hi = pr > pc[i] ? pr : pc[i], lo = pr > pc[i] ? pc[i] : pr;
          cv[i] = creal[i] ? jb.C[(size_t)hi * rpad + lo] : 0;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double g = 0.0;
        if (creal[i]) {
          const long long num = N * N * (long long)cv[i] - N * (sr + sc[i]) + Q;
          g = 2.0 * (double)num / den;
          if (r == c + i) g += jb.lambda;
        }
        out[i] = (float)g;
      }
    }
    *reinterpret_cast<float4*>(out_base + (size_t)r * ntp_all + c) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

int g_solve_mixed_smem_max = 0;

}  // namespace

static inline int solve_mixed_smem_bytes(int ntp) {
  return (3 * ntp + NB + 8 * NB + ST / 32) * (int)sizeof(double) + ntp * (int)sizeof(int);
}

cudaError_t tb_solve_mixed_init() {
  g_solve_mixed_smem_max = 220 * 1024;
  return cudaFuncSetAttribute(solve_mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_solve_mixed_smem_max);
}

bool tb_solve_mixed_fits(int ntp) { return solve_mixed_smem_bytes(ntp) <= 220 * 1024; }

cudaError_t tb_launch_solve_mixed(const TbSolveMixedJob* d_jobs, int n_jobs, int ntp, cudaStream_t st) {
  const int smem = solve_mixed_smem_bytes(ntp);
  if (smem > g_solve_mixed_smem_max) return cudaErrorInvalidConfiguration;
  solve_mixed_kernel<<<n_jobs, ST, smem, st>>>(d_jobs);
  return cudaGetLastError();
}

cudaError_t tb_launch_scale32(const TbScaleJob* d_jobs, int n_jobs, int ntp, float* L32, cudaStream_t st) {
  dim3 grid((ntp + 127) / 128, (ntp + 15) / 16, n_jobs);
  scale32_kernel<<<grid, 256, 0, st>>>(d_jobs, L32, ntp);
  return cudaGetLastError();
}
