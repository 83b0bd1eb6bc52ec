// Verification kernel for the Gram stage: the same integer cross-products as gram_tc.cu computed with
// plain dp4a dot products, no TMA / tensor cores / swizzles.  It exists so that the GPU tests can tell
// "the tcgen05 data path is wrong" from "the gather is wrong"; it is exported only through
// tb_gram_debug(impl=1) and is never used by tb_eval*.
#include "tb_internal.h"

namespace {

__global__ void __launch_bounds__(256) gram_simt_kernel(const int8_t* __restrict__ panel, int rpad, int kstride,
                                                        const int* __restrict__ kblocks, int32_t* __restrict__ C) {
  const int w = blockIdx.z;
  const int b = blockIdx.x * 16 + (threadIdx.x & 15);
  const int a = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (a >= rpad || b >= rpad || b > a) return;
  const int kbytes = kblocks[w] * TB_GRAM_BK;
  const int4* ra = reinterpret_cast<const int4*>(panel + ((size_t)w * rpad + a) * kstride);
  const int4* rb = reinterpret_cast<const int4*>(panel + ((size_t)w * rpad + b) * kstride);
  int acc = 0;
  for (int j = 0; j < kbytes / 16; ++j) {
    const int4 x = ra[j], y = rb[j];
    acc = __dp4a(x.x, y.x, acc);
    acc = __dp4a(x.y, y.y, acc);
    acc = __dp4a(x.z, y.z, acc);
    acc = __dp4a(x.w, y.w, acc);
  }
  C[((size_t)w * rpad + a) * rpad + b] = acc;
}

}  // namespace

cudaError_t tb_launch_gram_simt(const int8_t* d_panel, int W, int rpad, int kstride, const int* d_kblocks,
                                int32_t* d_C, cudaStream_t st) {
  dim3 grid((rpad + 15) / 16, (rpad + 15) / 16, W);
  gram_simt_kernel<<<grid, 256, 0, st>>>(d_panel, rpad, kstride, d_kblocks, d_C);
  return cudaGetLastError();
}
