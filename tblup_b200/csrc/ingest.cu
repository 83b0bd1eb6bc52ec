// Genotype ingest: animal-major int8 rows (host order already permuted into "universe" order) ->
// SNP-major device matrix, plus per-marker dosage sums over a set of animals.
// HBM-bound byte work; runs once per data set / per row set, never in the per-generation loop.
//
// Replaces the per-worker `np.load(data_path)` of tblup/evaluator.py:215 and the column means of
// tblup/utils.py:14 and tblup/evaluator.py:304 (kept as exact integer sums).
#include "tb_internal.h"

namespace {

// rows: [n_rows][m] (m contiguous).  x: [m][ldn], writes x[j][pos0 + r].
__global__ void transpose_rows_kernel(const int8_t* __restrict__ rows, int n_rows, int m, int8_t* __restrict__ x,
                                      int ldn, int pos0) {
  __shared__ int8_t tile[64][65];
  const int j0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 64 x 4
  for (int rr = ty; rr < 64; rr += 4) {
    int r = r0 + rr, j = j0 + tx;
    tile[rr][tx] = (r < n_rows && j < m) ? rows[(size_t)r * m + j] : (int8_t)0;
  }
  __syncthreads();
  for (int jj = ty; jj < 64; jj += 4) {
    int j = j0 + jj, r = r0 + tx;
    if (j < m && r < n_rows) x[(size_t)j * ldn + pos0 + r] = tile[tx][jj];
  }
}

// One warp per marker: colsum[j] = sum_i x[j][pos[i]].
__global__ void colsum_kernel(const int8_t* __restrict__ x, int ldn, int m, const int* __restrict__ pos, int n_pos,
                              int* __restrict__ colsum) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int8_t* row = x + (size_t)warp * ldn;
  int acc = 0;
  for (int i = lane; i < n_pos; i += 32) acc += row[pos[i]];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) colsum[warp] = acc;
}

}  // namespace

cudaError_t tb_launch_transpose_rows(const int8_t* d_rows, int n_rows, int m, int8_t* d_x, int ldn, int pos0,
                                     cudaStream_t st) {
  dim3 grid((m + 63) / 64, (n_rows + 63) / 64), block(64, 4);
  transpose_rows_kernel<<<grid, block, 0, st>>>(d_rows, n_rows, m, d_x, ldn, pos0);
  return cudaGetLastError();
}

cudaError_t tb_launch_colsum(const int8_t* d_x, int ldn, int m, const int* d_pos, int n_pos, int* d_colsum,
                             cudaStream_t st) {
  const int warps_per_block = 8;
  colsum_kernel<<<(m + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(d_x, ldn, m, d_pos,
                                                                                             n_pos, d_colsum);
  return cudaGetLastError();
}
