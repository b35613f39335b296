// Accuracy of the fp64 reciprocal square roots used by the factorisation chains, against 1 / sqrt(x) (both IEEE).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rsqrt_ulp rsqrt_ulp.cu && ./rsqrt_ulp
// Prints the worst relative error (in units of 2^-53) over 2^24 log-uniform samples in [1e-12, 1e12] per variant.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double newton2(double x) {       // two Newton steps (quadratic each)
  double y = seed(x);
  const double h = 0.5 * x;
  y = fma(y, fma(-h, y * y, 0.5), y);
  y = fma(y, fma(-h, y * y, 0.5), y);
  return y;
}
__device__ __forceinline__ double halley1(double x) {       // one cubic step: y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2
  const double y = seed(x);
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(e, 0.375, 0.5), y);
}
__global__ void k(double* worst, double* worst_seed) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  // log-uniform samples: x = 10^(-12 + 24 u), u from a Weyl sequence
  const double u = fmod((double)i * 0.6180339887498949, 1.0);
  const double x = exp10(-12.0 + 24.0 * u);
  const double ref = 1.0 / sqrt(x);
  const double v[3] = {::rsqrt(x), newton2(x), halley1(x)};
  for (int j = 0; j < 3; ++j) {
    const double err = fabs(v[j] - ref) / ref * 9007199254740992.0;
    // atomicMax on the bit pattern of a non-negative double
    atomicMax(reinterpret_cast<unsigned long long*>(worst + j), (unsigned long long)__double_as_longlong(err));
  }
  const double es = fabs(seed(x) - ref) / ref;
  atomicMax(reinterpret_cast<unsigned long long*>(worst_seed), (unsigned long long)__double_as_longlong(es));
}
int main() {
  double *d, h[4] = {0, 0, 0, 0};
  cudaMalloc(&d, sizeof(h));
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  k<<<(1 << 24) / 256, 256>>>(d, d + 3);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("worst relative error in units of 2^-53: libdevice rsqrt %.2f, seed + 2 Newton %.2f, seed + 1 cubic step %.2f; "
         "MUFU.RSQ64H seed alone: %.3g (2^%.1f)\n", h[0], h[1], h[2], h[3], log2(h[3]));
  return cudaDeviceSynchronize() != cudaSuccess;
}
