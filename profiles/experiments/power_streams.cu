// What does a streaming kernel's POWER depend on?  The sustained PD loop sits at the 1,000 W cap (SM clock 1.62 GHz) while
// torch's copy kernel moves the same bytes per second at 780 W and 1.965 GHz (pd_power_probe.py).  Variants of one
// grid-stride float4 kernel over the PD step's buffers (1M x 12: 100.7 MB state, 50.3 MB target, 50.3 MB out; 4 rotating
// sets), each run back to back for ~1.5 s with NVML power / SM clock sampled during the second half:
//   copy2      out[v] = state[2v] (+ state[2v+1] ignored)            1 read stream (50 MB) + 1 write stream  -- light
//   read3      out[v] = f(state[2v], state[2v+1], tgt[v]) trivial add  the PD step's exact traffic, ~no math
//   pdmath     the PD law (sub, mul, sub, mul, add, clamp) on registers, parameters from shared memory (LDS.128 x3)
//   pdmath_c   the same with parameters as compile-time constants (no shared-memory reads)
//   each in two load flavours: ld.global.nc.L1::no_allocate / st.global.L1::no_allocate  vs  plain ld / st
//   and grids of 6 / 3 / 2 CTAs of 256 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o power_streams power_streams.cu -ldl
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <bool NC> __device__ __forceinline__ float4 ld4(const float4* p) {
  float4 v;
  if (NC) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <bool NC> __device__ __forceinline__ void st4(float4* p, float4 v) {
  if (NC) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else *p = v;
}

// MODE 0 copy2, 1 read3, 2 pdmath (smem params), 3 pdmath (constant params)
template <int MODE, bool NC>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ state, const float4* __restrict__ tgt, float4* __restrict__ out,
                                         unsigned nvec, const float* __restrict__ par) {
  __shared__ __align__(16) float s_par[36];
  if (MODE == 2) { if (threadIdx.x < 36) s_par[threadIdx.x] = par[threadIdx.x]; __syncthreads(); }
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  int d0 = (int)((4ull * v) % 12u);
  const int dstep = (int)((4ull * stride) % 12u);
  for (; v < nvec; v += stride) {
    float4 o;
    if (MODE == 0) {
      o = ld4<NC>(state + 2 * v);
    } else {
      const float4 s0 = ld4<NC>(state + 2 * v), s1 = ld4<NC>(state + 2 * v + 1), tg = ld4<NC>(tgt + v);
      if (MODE == 1) {
        o = make_float4(s0.x + s1.x + tg.x, s0.y + s1.y + tg.y, s0.z + s1.z + tg.z, s0.w + s1.w + tg.w);
      } else {
        float4 kp, kd, tm;
        if (MODE == 2) {
          kp = *reinterpret_cast<const float4*>(s_par + d0);
          kd = *reinterpret_cast<const float4*>(s_par + 12 + d0);
          tm = *reinterpret_cast<const float4*>(s_par + 24 + d0);
          d0 += dstep; if (d0 >= 12) d0 -= 12;
        } else {
          kp = make_float4(400.f, 400.f, 400.f, 400.f); kd = make_float4(40.f, 40.f, 40.f, 40.f); tm = make_float4(80.f, 80.f, 80.f, 80.f);
        }
        auto law = [](float q, float qd, float t, float kp, float kd, float tm) {
          float tau = __fadd_rn(__fmul_rn(kp, __fsub_rn(t, q)), __fmul_rn(kd, -qd));
          tau = tau > tm ? tm : tau;
          return tau < -tm ? -tm : tau;
        };
        o = make_float4(law(s0.x, s0.y, tg.x, kp.x, kd.x, tm.x), law(s0.z, s0.w, tg.y, kp.y, kd.y, tm.y),
                        law(s1.x, s1.y, tg.z, kp.z, kd.z, tm.z), law(s1.z, s1.w, tg.w, kp.w, kd.w, tm.w));
      }
    }
    st4<NC>(out + v, o);
  }
}

typedef int (*nvmlInit_t)(void);
typedef int (*nvmlHandle_t)(unsigned, void**);
typedef int (*nvmlPower_t)(void*, unsigned*);
typedef int (*nvmlClock_t)(void*, int, unsigned*);
static void* g_dev;
static nvmlPower_t g_power;
static nvmlClock_t g_clock;

int main() {
  void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!h) { printf("no NVML\n"); return 1; }
  ((nvmlInit_t)dlsym(h, "nvmlInit_v2"))();
  ((nvmlHandle_t)dlsym(h, "nvmlDeviceGetHandleByIndex_v2"))(0, &g_dev);
  g_power = (nvmlPower_t)dlsym(h, "nvmlDeviceGetPowerUsage");
  g_clock = (nvmlClock_t)dlsym(h, "nvmlDeviceGetClockInfo");
  const unsigned n = 1u << 20, D = 12, nvec = n * D / 4;
  const int SETS = 4;
  float4 *state[SETS], *tgt[SETS], *out[SETS];
  for (int i = 0; i < SETS; ++i) {
    CK(cudaMalloc(&state[i], (size_t)nvec * 32)); CK(cudaMalloc(&tgt[i], (size_t)nvec * 16)); CK(cudaMalloc(&out[i], (size_t)nvec * 16));
    CK(cudaMemset(state[i], 0x3c, (size_t)nvec * 32)); CK(cudaMemset(tgt[i], 0x3d, (size_t)nvec * 16));
  }
  float hp[36]; for (int i = 0; i < 36; ++i) hp[i] = i < 12 ? 400.f : i < 24 ? 40.f : 80.f;
  float* par; CK(cudaMalloc(&par, sizeof hp)); CK(cudaMemcpy(par, hp, sizeof hp, cudaMemcpyHostToDevice));
  struct V { const char* name; void (*fn)(const float4*, const float4*, float4*, unsigned, const float*); double bytes; };
  const double b3 = (double)nvec * 64, b2 = (double)nvec * 32;
  V vs[] = {{"copy2  nc", k<0, true>, b2}, {"copy2  ld", k<0, false>, b2}, {"read3  nc", k<1, true>, b3}, {"read3  ld", k<1, false>, b3},
            {"pdmath nc", k<2, true>, b3}, {"pdmath ld", k<2, false>, b3}, {"pdmat_c nc", k<3, true>, b3}};
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int per_sm : {6, 3, 2}) {
    for (auto& v : vs) {
      std::atomic<bool> stop{false};
      std::vector<unsigned> pw, ck;
      std::thread smp([&] {
        while (!stop) { unsigned p = 0, c = 0; g_power(g_dev, &p); g_clock(g_dev, 1, &c); pw.push_back(p); ck.push_back(c);
                        std::this_thread::sleep_for(std::chrono::milliseconds(20)); }
      });
      const int grid = 148 * per_sm;
      auto run = [&](double seconds) {
        auto t0 = std::chrono::steady_clock::now();
        long reps = 0;
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
          for (int i = 0; i < 64; ++i) { const int s = (int)((reps + i) % SETS); v.fn<<<grid, 256>>>(state[s], tgt[s], out[s], nvec, par); }
          reps += 64;
          CK(cudaDeviceSynchronize());
        }
        return reps;
      };
      run(0.6);
      pw.clear(); ck.clear();
      CK(cudaEventRecord(e0));
      const long reps = run(0.8);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      stop = true; smp.join();
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      std::sort(pw.begin(), pw.end()); std::sort(ck.begin(), ck.end());
      printf("%d CTAs/SM  %-10s %7.2f us/launch  %7.1f GB/s   %4u W  sm %4u MHz\n", per_sm, v.name, ms * 1e3 / reps, v.bytes * reps / ms / 1e6,
             pw.empty() ? 0 : pw[pw.size() / 2] / 1000, ck.empty() ? 0 : ck[ck.size() / 2]);
      fflush(stdout);
    }
  }
  return 0;
}
