// How fast can B200 HBM deliver a small contiguous slot out of every wide row?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_slot_bw strided_slot_bw.cu && ./strided_slot_bw
// The O family reads j_eef = jacobian[:, hand, :, :7]: a 216 B window (256 B of sectors) out of every 2,376 B row
// of Isaac Gym's (N, 11, 6, 9) jacobian tensor.  This measures the ceiling of that access pattern with plain,
// fully coalesced 16 B loads at full occupancy, independent of our staging engine.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void slot_read(const float4* __restrict__ src, float* __restrict__ out, long n, int row_f4, int slot_f4, int off_f4) {
  // `slot_f4` consecutive lanes read one env's slot; a warp covers 32 / slot_f4 envs per step
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long i = gid; i < n * slot_f4; i += stride) {
    const long env = i / slot_f4;
    const int k = (int)(i - env * slot_f4);
    const float4 v = __ldcs(src + env * row_f4 + off_f4 + k);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[gid % 32] = acc;   // keep the loads alive
}

int main() {
  const long n = 262144;
  const int row_bytes = 2376;                 // 11 * 6 * 9 * 4
  // rows are 8 B aligned only in general; use a 16 B multiple close to the real stride for the float4 kernel
  const int row_f4 = (row_bytes + 8) / 16;    // 2,384 B
  float4* src; float* out;
  cudaMalloc(&src, n * row_f4 * 16L);
  cudaMalloc(&out, 4096);
  cudaMemset(src, 0, n * row_f4 * 16L);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int slots[] = {1, 2, 4, 8, 16, 32, 64, row_f4};
  printf("row stride %d B, %ld rows (%.0f MB)\n", row_f4 * 16, n, n * row_f4 * 16.0 / 1e6);
  for (int slot_f4 : slots) {
    const int grid = 148 * 16;
    for (int w = 0; w < 3; ++w) slot_read<<<grid, 256>>>(src, out, n, row_f4, slot_f4, slot_f4 == row_f4 ? 0 : 30);
    cudaEventRecord(a);
    const int iters = 20;
    for (int it = 0; it < iters; ++it) slot_read<<<grid, 256>>>(src, out, n, row_f4, slot_f4, slot_f4 == row_f4 ? 0 : 30);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= iters;
    const double useful = n * slot_f4 * 16.0;
    printf("slot %5d B: %8.2f us  useful %7.1f GB/s\n", slot_f4 * 16, ms * 1e3, useful / (ms * 1e-3) / 1e9);
  }
  if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 1; }
  return 0;
}
