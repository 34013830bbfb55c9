// smem_peak.cu — measured shared-memory bandwidth of the device: the denominator of the fixpoint kernel's roofline
// (SURVEY.md 8d states it as 128 B/clk/SM; this measures what an LDS.64 stream really gets, so that
// `roofline.frac` is a ratio of two measurements).
//
// Every thread streams conflict-free 8-byte loads (the access the fixpoint loop issues: one {lb, ub} pair per lane)
// from a 32 KB shared array, 16 independent loads in flight per thread; the loaded values are folded into a
// checksum so that nothing is optimised away. One launch reads `iters * 16 * 8` bytes per thread.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/turbo_b200.h"

void tb_set_error_internal(const char* s);

namespace {

constexpr int kThreads = 1024;
constexpr int kSlots = 4096;          // 32 KB of {lb, ub} pairs

__global__ void __launch_bounds__(kThreads) smem_stream_kernel(int iters, unsigned long long* out) {
  __shared__ int2 store[kSlots + 256];
  for (int i = threadIdx.x; i < kSlots + 256; i += kThreads) store[i] = make_int2(i, ~i);
  __syncthreads();
  // lane l of a warp reads 32 consecutive 8-byte words: two conflict-free wavefronts per LDS.64; the 16 loads of one
  // iteration are independent (fixed addresses 2 KB apart, immediate offsets: no address arithmetic in the loop)
  const unsigned a = (unsigned)__cvta_generic_to_shared(store) + (threadIdx.x & 31) * 8u + ((threadIdx.x >> 5) & 7) * 256u;
  int acc0 = 0, acc1 = 0;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#define TB_LD(K) { int x, y; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2+" #K "];" : "=r"(x), "=r"(y) : "r"(a)); acc0 ^= x; acc1 += y; }
    TB_LD(0) TB_LD(2048) TB_LD(4096) TB_LD(6144) TB_LD(8192) TB_LD(10240) TB_LD(12288) TB_LD(14336)
    TB_LD(16384) TB_LD(18432) TB_LD(20480) TB_LD(22528) TB_LD(24576) TB_LD(26624) TB_LD(28672) TB_LD(30720)
#undef TB_LD
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (unsigned long long)(c1 - c0);   // SM cycles of the stream, one CTA's view
  if ((acc0 ^ acc1) == 0x7fffffff) out[0] = (unsigned long long)acc0;     // never true: keeps the loads alive
}

}  // namespace

// Measured shared-memory read bandwidth of `device` in GB/s (all SMs, LDS.64 stream, best of `reps` launches), and
// the bytes per clock per SM it corresponds to at the SM clock the device reports as its maximum.
extern "C" tb_status tb_measure_smem_peak(int32_t device, double* gb_per_s, double* bytes_per_clk_per_sm) {
  if (!gb_per_s) { tb_set_error_internal("null argument"); return TB_ERR_INVALID; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); tb_set_error_internal("no such CUDA device"); return TB_ERR_NO_DEVICE; }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaDeviceProp dp;
  cudaGetDeviceProperties(&dp, device);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smem_stream_kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const int grid = dp.multiProcessorCount * per_sm;
  unsigned long long* d = nullptr;
  cudaEvent_t e0, e1;
  tb_status rc = TB_OK;
  if (cudaMalloc(&d, 64) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    cudaGetLastError(); tb_set_error_internal("smem peak: allocation failed"); cudaSetDevice(prev); return TB_ERR_CUDA;
  }
  const int iters = 4096;
  double best_ms = 1e30;
  for (int r = 0; r < 6; ++r) {          // the first launches warm the clocks up
    cudaEventRecord(e0);
    smem_stream_kernel<<<grid, kThreads>>>(iters, d);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = TB_ERR_CUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2 && ms < best_ms) best_ms = ms;
  }
  if (rc == TB_OK) {
    const double bytes = (double)grid * kThreads * (double)iters * 16.0 * 8.0;
    *gb_per_s = bytes / (best_ms * 1e-3) / 1e9;
    // bytes per clock per SM from the SM's own cycle counter (no assumption about the clock the launch ran at): the
    // CTAs resident on block 0's SM moved per_sm * kThreads * iters * 128 B in the cycles block 0 counted
    unsigned long long h[2] = {0, 0};
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (bytes_per_clk_per_sm) *bytes_per_clk_per_sm = h[1] ? (double)per_sm * kThreads * (double)iters * 16.0 * 8.0 / (double)h[1] : 0.0;
  } else {
    tb_set_error_internal("smem peak: kernel failed"); cudaGetLastError();
  }
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaSetDevice(prev);
  return rc;
}
