// smem_peak.cu — measured shared-memory bandwidth of the device: the denominator of the fixpoint kernel's roofline
// (SURVEY.md 8d states it as 128 B/clk/SM; this measures what an LDS.64 stream really gets, so that
// `roofline.frac` is a ratio of two measurements).
//
// Every thread streams conflict-free 8-byte loads (the access the fixpoint loop issues: one {lb, ub} pair per lane)
// from a 32 KB shared array, 16 independent loads in flight per thread; the loaded values are folded into a
// checksum so that nothing is optimised away. One launch reads `iters * 16 * 8` bytes per thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/turbo_b200.h"

void tb_set_error_internal(const char* s);

namespace {

constexpr int kThreads = 1024;
constexpr int kBytes = 136 * 1024;    // 17 K {lb, ub} pairs of dynamic shared memory (MODE 1: every warp streams its own 4 KB window and the next)

// (The "memory" clobber of the load matters: without it the compiler merges the identical loads of the iterations it
// unrolls, one LDS per four counted - the first version of this benchmark reported 478 B/clk/SM that way; ncu's
// instruction count gave it away.)
// MODE 0: every load of a lane hits the same 16 addresses in every iteration;
// MODE 1: warp w reads its own 4 KB window, lane l the l-th pair of a 256-byte row, the row advancing every load:
//         32 warps x 16 rows of distinct addresses (a contiguous row stream: the hardware serves a 256-byte row wide);
// MODE 2: a GATHER: the 16 lanes of a half-warp read 16 different 8-byte banks in 16 different 128-byte rows, the rows
//         changing with every load - conflict free, nothing contiguous: the fixpoint loop's access shape at its best
//         (two wavefronts per LDS.64). This is the measured denominator of the roofline.
template <int MODE>
__global__ void __launch_bounds__(kThreads) smem_stream_kernel(int iters, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char dyn[];
  int2* store = (int2*)dyn;
  for (int i = threadIdx.x; i < kBytes / 8; i += kThreads) store[i] = make_int2(i, ~i);
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(dyn);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // MODE 2: bank = lane mod 16 (8-byte banks), row = a per-lane pseudo-random 128-byte row; the 16 immediate offsets
  // below move every lane by whole rows, so the banks stay distinct within a half-warp
  const unsigned a = MODE == 0 ? base + lane * 8u + (warp & 7) * 256u
                   : MODE == 1 ? base + warp * 4096u + lane * 8u + (blockIdx.x & 0u)
                               : base + (lane & 15) * 8u + (((lane * 37u + warp * 101u) & 511u) << 7);
  int acc0 = 0, acc1 = 0;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // whole rows more every iteration (same banks): the compiler cannot merge the loads of the iterations it unrolls
    const unsigned ai = MODE == 1 ? a + (((unsigned)it * 256u) & 0xF00u) : a + (((unsigned)it * 128u) & 0x780u);
#define TB_LD(K) { int x, y; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2+" #K "];" : "=r"(x), "=r"(y) : "r"(ai) : "memory"); acc0 ^= x; acc1 += y; }
    if (MODE == 0) {
      TB_LD(0) TB_LD(2048) TB_LD(4096) TB_LD(6144) TB_LD(8192) TB_LD(10240) TB_LD(12288) TB_LD(14336)
      TB_LD(16384) TB_LD(18432) TB_LD(20480) TB_LD(22528) TB_LD(24576) TB_LD(26624) TB_LD(28672) TB_LD(30720)
    } else if (MODE == 1) {
      TB_LD(0) TB_LD(256) TB_LD(512) TB_LD(768) TB_LD(1024) TB_LD(1280) TB_LD(1536) TB_LD(1792)
      TB_LD(2048) TB_LD(2304) TB_LD(2560) TB_LD(2816) TB_LD(3072) TB_LD(3328) TB_LD(3584) TB_LD(3840)
    } else {      // rows 0..511 from `a`, plus up to 496 rows here: inside the 1024 rows of the array
      TB_LD(0) TB_LD(4224) TB_LD(8448) TB_LD(12672) TB_LD(16896) TB_LD(21120) TB_LD(25344) TB_LD(29568)
      TB_LD(33792) TB_LD(38016) TB_LD(42240) TB_LD(46464) TB_LD(50688) TB_LD(54912) TB_LD(59136) TB_LD(63360)
    }
#undef TB_LD
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (unsigned long long)(c1 - c0);   // SM cycles of the stream, one CTA's view
  if ((acc0 ^ acc1) == 0x7fffffff) out[0] = (unsigned long long)acc0;     // never true: keeps the loads alive
}

template <int MODE>
tb_status run_stream(int device, double* gb_per_s, double* bytes_per_clk_per_sm) {
  cudaDeviceProp dp;
  cudaGetDeviceProperties(&dp, device);
  if (cudaFuncSetAttribute(smem_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes) != cudaSuccess) { cudaGetLastError(); return TB_ERR_CUDA; }
  const int grid = dp.multiProcessorCount;          // one CTA of 1024 threads per SM
  unsigned long long* d = nullptr;
  cudaEvent_t e0, e1;
  tb_status rc = TB_OK;
  if (cudaMalloc(&d, 64) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    cudaGetLastError(); tb_set_error_internal("smem peak: allocation failed"); return TB_ERR_CUDA;
  }
  const int iters = 4096;
  double best_ms = 1e30;
  unsigned long long best_cycles = 0;
  for (int r = 0; r < 6; ++r) {          // the first launches warm the clocks up
    cudaEventRecord(e0);
    smem_stream_kernel<MODE><<<grid, kThreads, kBytes>>>(iters, d);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = TB_ERR_CUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2] = {0, 0};
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (r >= 2 && ms < best_ms) { best_ms = ms; best_cycles = h[1]; }
  }
  if (rc == TB_OK) {
    const double per_sm = (double)kThreads * (double)iters * 16.0 * 8.0;
    *gb_per_s = per_sm * grid / (best_ms * 1e-3) / 1e9;
    // bytes per clock per SM from the SM's own cycle counter: no assumption about the clock the launch ran at
    if (bytes_per_clk_per_sm) *bytes_per_clk_per_sm = best_cycles ? per_sm / (double)best_cycles : 0.0;
  } else {
    tb_set_error_internal("smem peak: kernel failed"); cudaGetLastError();
  }
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

}  // namespace

// Measured shared-memory read bandwidth of `device` in GB/s (all SMs, LDS.64 stream, best of four launches), and the
// bytes per clock per SM by the SM's own cycle counter. Default: the conflict-free gather (MODE 2);
// TB_SMEM_PEAK_MODE=1 the contiguous row stream, =0 the repeated-address variant.
extern "C" tb_status tb_measure_smem_peak(int32_t device, double* gb_per_s, double* bytes_per_clk_per_sm) {
  if (!gb_per_s) { tb_set_error_internal("null argument"); return TB_ERR_INVALID; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); tb_set_error_internal("no such CUDA device"); return TB_ERR_NO_DEVICE; }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  const char* m = getenv("TB_SMEM_PEAK_MODE");
  const tb_status rc = (m && m[0] == '0') ? run_stream<0>(device, gb_per_s, bytes_per_clk_per_sm)
                     : (m && m[0] == '1') ? run_stream<1>(device, gb_per_s, bytes_per_clk_per_sm)
                                          : run_stream<2>(device, gb_per_s, bytes_per_clk_per_sm);
  cudaSetDevice(prev);
  return rc;
}
