// FMA-pipe micro-benchmark: the MEASURED denominator of the "FP32-pipe utilisation" the metric asks for family O
// ("the harness must measure an FMA micro-benchmark").  MEASURED_PEAKS.json holds HBM bandwidth and
// tensor-core throughput only; the CUDA-core FMA rate depends on the clock the part sustains under this very load.
// Each thread runs 16 independent dependent-FMA chains (enough to cover the pipe latency at full occupancy); the
// result is stored so nothing is eliminated.
#include "common.cuh"

namespace b200ctl {

template <typename T>
__global__ void __launch_bounds__(1024) fma_peak_kernel(T* out, int iters, T b, T c) {
  T a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = (T)(threadIdx.x + k) * (T)1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = a[k] * b + c;      // contracted to one FMA per chain and iteration
  }
  T s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = s;
}

template <typename T>
static int run_fma_peak(int dev, int launches, double* tflops_best, double* tflops_median) {
  const int block = 1024, iters = 4096;
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fma_peak_kernel<T>, block, 0));
  const int grid = sm_count(dev) * (per_sm > 0 ? per_sm : 1) * 4;
  T* out = nullptr;
  B200_CUDA(cudaMalloc(&out, sizeof(T) * (size_t)grid * block));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const double flop = 2.0 * 16.0 * iters * (double)grid * block;
  double t[64];
  if (launches > 64) launches = 64;
  for (int w = 0; w < 3; ++w) fma_peak_kernel<T><<<grid, block>>>(out, iters, (T)0.999, (T)1e-4);
  int rc = 0;
  for (int l = 0; l < launches && rc == 0; ++l) {
    cudaEventRecord(e0);
    fma_peak_kernel<T><<<grid, block>>>(out, iters, (T)0.999, (T)1e-4);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = 1; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    t[l] = flop / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (rc) B200_FAIL((int)cudaGetLastError(), "fma_peak_kernel failed");
  for (int i = 1; i < launches; ++i)
    for (int j = i; j > 0 && t[j] < t[j - 1]; --j) { const double x = t[j]; t[j] = t[j - 1]; t[j - 1] = x; }
  *tflops_best = t[launches - 1];
  *tflops_median = t[launches / 2];
  g_launch_count.fetch_add(launches + 3, std::memory_order_relaxed);
  return 0;
}

}  // namespace b200ctl

using namespace b200ctl;

extern "C" int b200ctl_measure_fma_peak(int32_t dtype, int32_t device, int32_t launches, double* tflops_best,
                                        double* tflops_median) {
  if (!tflops_best || !tflops_median) B200_FAIL(B200CTL_E_NULL, "output pointer is NULL");
  if (dtype != 0 && dtype != 1) B200_FAIL(B200CTL_E_VALUE, "dtype must be 0 (fp32) or 1 (fp64)");
  if (launches < 1) B200_FAIL(B200CTL_E_VALUE, "launches must be >= 1");
  DeviceGuard g;
  B200_TRY(g.enter(device));
  return dtype == 0 ? run_fma_peak<float>(device, launches, tflops_best, tflops_median)
                    : run_fma_peak<double>(device, launches, tflops_best, tflops_median);
}
