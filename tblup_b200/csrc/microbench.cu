// Peak probes used as roofline denominators where MEASURED_PEAKS.json has no entry:
//   fp64 DMMA issue rate (register-resident operands, no memory traffic) for the Cholesky update kernel.
//   tcgen05 kind::i8 and kind::mxf4 issue rate with shared-memory-resident operands (no TMA, no epilogue): the
//   ceiling of the Gram kernel's main loop (SURVEY.md 8d asks for a measured int8 tcgen05 peak; MEASURED_PEAKS.json
//   only holds bf16).
#include "tb_internal.h"
#include "tb_ptx.cuh"

namespace {

using namespace tbptx;

// One CTA per SM; A (128 rows) and B (256 rows) K-major tiles of one 128-byte swizzle span sit in shared memory
// (pseudo-random bytes: every E2M1 nibble / int8 byte is a finite value) and ONE thread issues
// M128 x N256 MMAs back to back into one TMEM accumulator: K = 32 per instruction for kind::i8, K = 64 for
// kind::mxf4 (block scales 2^0).  A commit every 16 k-blocks, at most two batches in flight.
// dosage != 0: the operands hold genotype-like values (dosage 0 / 1 / 2 with probabilities ~ .55 / .37 / .08, the E2M1
// nibble 2 d or the int8 byte d) instead of random bits -- the tensor cores' power draw, and with it the clock the
// board sustains under its power cap, depends on how many operand bits toggle.
template <bool FP4>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters, uint32_t seed, int dosage) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);      // [2]
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    if (dosage) {
      uint32_t out = 0, g = h;
      const int per = FP4 ? 8 : 4, bits = FP4 ? 4 : 8;
      for (int e = 0; e < per; ++e) {
        g = g * 1664525u + 1013904223u;
        const uint32_t u = g >> 24;                     // 0..255
        const uint32_t d = u < 141 ? 0u : (u < 236 ? 1u : 2u);
        out |= (FP4 ? 2u * d : d) << (bits * e);
      }
      h = out;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = h;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (FP4) {
    tmem_fill_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 480, 0x7f7f7f7fu);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + A_BYTES);
    int commits = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (FP4) umma_mxf4(tmem_base, adesc + 2 * k, bdesc + 2 * k, umma_idesc_mxf4(128, 256), (it | k) != 0,
                           tmem_base + 480, tmem_base + 488);
        else umma_s8(tmem_base, adesc + 2 * k, bdesc + 2 * k, umma_idesc_s8(128, 256), (it | k) != 0);
      }
      if ((it & 15) == 15 || it == iters - 1) {
        umma_commit(&bars[commits & 1]);
        if (commits >= 1) mbar_wait(&bars[(commits - 1) & 1], ((commits - 1) >> 1) & 1);
        ++commits;
      }
    }
    if (commits >= 1) mbar_wait(&bars[(commits - 1) & 1], ((commits - 1) >> 1) & 1);
    tc_fence_after();
  }
  (void)lane;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

__global__ void __launch_bounds__(256, 2) dmma_peak_kernel(double* sink, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    b[i] = 1.0 - 1e-9 * (threadIdx.x + 2 * i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(acc[mi * 4 + ni][0]), "+d"(acc[mi * 4 + ni][1])
                     : "d"(a[mi]), "d"(b[ni]));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;   // keep the loop alive
}

}  // namespace

cudaError_t tb_microbench_dmma(int n_sm, cudaStream_t st, double* tflops) {
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, 8);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = n_sm * 2, iters = 20000;
  dmma_peak_kernel<<<blocks, 256, 0, st>>>(sink, 2000);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, st);
    dmma_peak_kernel<<<blocks, 256, 0, st>>>(sink, iters);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)blocks * 8.0 * iters * 16.0 * 512.0;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return e;
}


// which: 1 = tcgen05 kind::i8 (TOP/s, 2 ops per multiply-accumulate), 2 = kind::mxf4 on E2M1 operands: best of three
// ~20 ms launches on random operand bits (burst).  3 / 4 = the same two instructions SUSTAINED: launches back to back
// for ~2.5 s, rate over the last ~1.5 s (the clock has settled under the power cap by then), random operand bits;
// 5 / 6 = sustained with genotype-like operands.
cudaError_t tb_microbench_umma(int which, int n_sm, cudaStream_t st, double* tops) {
  const int smem = 128 * 128 + 256 * 128 + 1024 + 64;
  const bool fp4 = which == 2 || which == 4 || which == 6;
  const bool sustained = which >= 3;
  const int dosage = which >= 5 ? 1 : 0;
  cudaError_t e = fp4 ? cudaFuncSetAttribute((const void*)umma_peak_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                      : cudaFuncSetAttribute((const void*)umma_peak_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto launch = [&](int iters) {
    if (fp4) umma_peak_kernel<true><<<n_sm, 128, smem, st>>>(iters, 12345u, dosage);
    else umma_peak_kernel<false><<<n_sm, 128, smem, st>>>(iters, 12345u, dosage);
  };
  launch(2048);                                    // warm-up
  const int iters = fp4 ? 80000 : 40000;           // ~20 ms per launch
  double best = 0.0;
  if (sustained) {
    const double ops1 = (double)n_sm * iters * 4.0 * 2.0 * 128.0 * 256.0 * (fp4 ? 64.0 : 32.0);
    const int n_pre = 50, n_meas = 75;             // ~1 s to settle, ~1.5 s measured
    for (int i = 0; i < n_pre; ++i) launch(iters);
    cudaEventRecord(e0, st);
    for (int i = 0; i < n_meas; ++i) launch(iters);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops = ms > 0.f ? ops1 * n_meas / (ms * 1e-3) / 1e12 : 0.0;
    return e;
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, st);
    launch(iters);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)n_sm * iters * 4.0 * 2.0 * 128.0 * 256.0 * (fp4 ? 64.0 : 32.0);
    best = std::max(best, ops / (ms * 1e-3) / 1e12);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tops = best;
  return e;
}
