// Batched blocked Cholesky of A = G_tt + lambda I (fp64), one matrix per (individual, row set), all
// matrices of a wave advanced in lock-step.  Replaces the explicit `np.linalg.inv` of
// tblup/evaluator.py:282 (and the normal-equation solve inside sklearn's Ridge, evaluator.py:311-312):
// only the action of the inverse on y is needed, so we factor A = L L^T and substitute (solve.cu).
//
// Left-looking, block size NB = 64.  Step j:
//   update  T[i][j] = A[i][j] - sum_{k<j} L[i][k] L[j][k]^T  for row blocks i >= j   (chol_gemm_kernel<0>)
//   diag    L[j][j] = chol(T[j][j]),  Linv_j = L[j][j]^-1                          (chol_diag_kernel)
//   panel   L[i][j] = T[i][j] Linv_j^T                          for i > j          (chol_gemm_kernel<1>)
// Left-looking keeps the long dimension in K: every L entry is written once and the n^3/3 flops run in a
// double-precision tensor-core GEMM ("NT": both operands K-contiguous) built on mma.sync m8n8k4 f64 (DMMA)
// with a 3-stage cp.async pipeline.  Roofline: FP64 pipe.
#include "tb_internal.h"

namespace {

constexpr int NB = TB_NB;
constexpr int GBM = 128, GBN = 64, GBK = 16, GSTAGES = 3;
constexpr int LDK = GBK + 4;                       // padded K stride (doubles): conflict-free LDS.64 fragments
constexpr int A_TILE = GBM * LDK, B_TILE = GBN * LDK;
constexpr int GEMM_SMEM = GSTAGES * (A_TILE + B_TILE) * (int)sizeof(double);
constexpr int DIAG_SMEM = 2 * NB * (NB + 1) * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// MODE 0: out = M[r][jc] - sum_{k < j NB} M[r][k] M[j NB + c][k]       rows r >= j NB
// MODE 1: out = sum_{p < NB} M[r][j NB + p] Linv_j[c][p]                rows r >= (j+1) NB
template <int MODE>
__global__ void __launch_bounds__(256, 2) chol_gemm_kernel(const TbCholJob* __restrict__ jobs, int j) {
  extern __shared__ double gsm[];
  const TbCholJob jb = jobs[blockIdx.y];
  const int ntp = jb.ntp;
  const int row0 = (MODE == 0 ? j : j + 1) * NB + blockIdx.x * GBM;
  if (row0 >= ntp) return;
  const int ktot = MODE == 0 ? j * NB : NB;
  const int nk = ktot / GBK;
  double* M = jb.M;
  const double* Aop = MODE == 0 ? M : M + (size_t)j * NB;          // + r * ntp + k
  const double* Bop = MODE == 0 ? M + (size_t)j * NB * ntp : jb.Linv + (size_t)j * NB * NB;
  const int ldb = MODE == 0 ? ntp : NB;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;           // 4 x 2 warps, 32 x 32 each
  const int g = lane >> 2, t4 = lane & 3;

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  auto load_stage = [&](int kt, int slot) {
    double* As = gsm + slot * (A_TILE + B_TILE);
    double* Bs = As + A_TILE;
    const int kbase = kt * GBK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {                    // A: 128 rows x 8 chunks of 16 B
      const int ch = tid + i * 256, r = ch >> 3, c = ch & 7;
      int gr = row0 + r;
      gr = gr < ntp ? gr : ntp - 1;
      cp_async16(As + r * LDK + c * 2, Aop + (size_t)gr * ntp + kbase + c * 2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {                    // B: 64 rows x 8 chunks
      const int ch = tid + i * 256, r = ch >> 3, c = ch & 7;
      cp_async16(Bs + r * LDK + c * 2, Bop + (size_t)r * ldb + kbase + c * 2);
    }
  };

  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    const int pf = kt + GSTAGES - 1;
    if (pf < nk) load_stage(pf, pf % GSTAGES);
    cp_async_commit();
    const double* As = gsm + (kt % GSTAGES) * (A_TILE + B_TILE);
    const double* Bs = As + A_TILE;
#pragma unroll
    for (int kk = 0; kk < GBK / 4; ++kk) {
      double af[4], bf[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) af[mi] = As[(wm * 32 + mi * 8 + g) * LDK + kk * 4 + t4];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) bf[ni] = Bs[(wn * 32 + ni * 8 + g) * LDK + kk * 4 + t4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  if (MODE == 1) __syncthreads();   // every warp has consumed T before anyone overwrites it in place

#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int r = row0 + wm * 32 + mi * 8 + g;
    if (r >= ntp) continue;
    double2* dst = reinterpret_cast<double2*>(M + (size_t)r * ntp + j * NB + wn * 32 + 2 * t4);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2 v;
      if (MODE == 0) {
        v = dst[ni * 4];
        v.x -= acc[mi][ni][0];
        v.y -= acc[mi][ni][1];
      } else {
        v.x = acc[mi][ni][0];
        v.y = acc[mi][ni][1];
      }
      dst[ni * 4] = v;
    }
  }
}

// One CTA per matrix: factor the 64 x 64 diagonal block and invert the factor.
// potrf: thread (r = tid / 4, q = tid % 4) keeps row r, columns q, q+4, ... in registers; per column one
// barrier: the owners publish the raw column through a double-buffered smem vector, every thread rescales it
// by 1/sqrt(pivot) itself and applies the rank-1 update to its registers.
// inverse: forward substitution row by row, the dot products of a row split over 4 thread groups.
__global__ void __launch_bounds__(256) chol_diag_kernel(const TbCholJob* __restrict__ jobs, int j) {
  extern __shared__ double dsm[];
  double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm);
  double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + NB * (NB + 1));
  __shared__ double colbuf[2][NB];
  __shared__ double part[4][NB];
  __shared__ int bad;
  const TbCholJob jb = jobs[blockIdx.x];
  const int ntp = jb.ntp;
  if (j * NB >= ntp) return;
  double* D = jb.M + (size_t)j * NB * ntp + j * NB;
  const int tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;
  if (tid == 0) bad = 0;
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = q + 4 * i;
    a[i] = c <= r ? D[(size_t)r * ntp + c] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    const int qc = c & 3, ic = c >> 2;
    if (q == qc && r >= c) colbuf[c & 1][r] = a[ic];
    __syncthreads();
    const double d = colbuf[c & 1][c];
    if (!(d > 0.0) && tid == 0) bad = 1;
    const double inv = 1.0 / sqrt(d);
    if (r >= c) {
      const double lr = colbuf[c & 1][r] * inv;
      if (q == qc) a[ic] = lr;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int cc = q + 4 * i;
        if (cc > c && cc <= r) a[i] -= lr * (colbuf[c & 1][cc] * inv);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = q + 4 * i;
    Ls[r][c] = c <= r ? a[i] : 0.0;
    Xs[r][c] = 0.0;
  }
  __syncthreads();
  // X = L^-1, row by row: X[rr][c] = (delta - sum_{p<rr} L[rr][p] X[p][c]) / L[rr][rr]
  {
    const int c = tid & 63, h = tid >> 6;
    for (int rr = 0; rr < NB; ++rr) {
      double sacc = 0.0;
      if (c <= rr) {
        for (int p = c + h; p < rr; p += 4) sacc += Ls[rr][p] * Xs[p][c];   // X[p][c] = 0 for p < c
      }
      part[h][c] = sacc;
      __syncthreads();
      if (h == 0 && c <= rr) {
        const double tot = (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
        Xs[rr][c] = ((c == rr ? 1.0 : 0.0) - tot) / Ls[rr][rr];
      }
      __syncthreads();
    }
  }
  double* Li = jb.Linv + (size_t)j * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    const int rr = e >> 6, c = e & 63;
    D[(size_t)rr * ntp + c] = Ls[rr][c];
    Li[e] = Xs[rr][c];
  }
  if (tid == 0 && bad) *jb.status = 1;
}

}  // namespace

cudaError_t tb_chol_init() {
  cudaError_t e = cudaFuncSetAttribute(chol_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(chol_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
}

cudaError_t tb_launch_chol_update(const TbCholJob* d_jobs, int n_jobs, int max_ntp, int j, cudaStream_t st) {
  const int rows = max_ntp - j * NB;
  if (j == 0 || rows <= 0) return cudaSuccess;
  dim3 grid((rows + GBM - 1) / GBM, n_jobs);
  chol_gemm_kernel<0><<<grid, 256, GEMM_SMEM, st>>>(d_jobs, j);
  return cudaGetLastError();
}

cudaError_t tb_launch_chol_diag(const TbCholJob* d_jobs, int n_jobs, int max_ntp, int j, cudaStream_t st) {
  if (j * NB >= max_ntp) return cudaSuccess;
  chol_diag_kernel<<<n_jobs, 256, DIAG_SMEM, st>>>(d_jobs, j);
  return cudaGetLastError();
}

cudaError_t tb_launch_chol_panel(const TbCholJob* d_jobs, int n_jobs, int max_ntp, int j, cudaStream_t st) {
  const int rows = max_ntp - (j + 1) * NB;
  if (rows <= 0) return cudaSuccess;
  dim3 grid((rows + GBM - 1) / GBM, n_jobs);
  chol_gemm_kernel<1><<<grid, 256, GEMM_SMEM, st>>>(d_jobs, j);
  return cudaGetLastError();
}
