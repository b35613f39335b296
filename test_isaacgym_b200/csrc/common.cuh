// Shared host/device plumbing for libb200ctl: error reporting, DLTensor
// validation, strided tensor views, launch helpers.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/b200ctl.h"

namespace b200ctl {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

#define B200_FAIL(code, ...)             \
  do {                                   \
    ::b200ctl::set_error(__VA_ARGS__);   \
    return (code);                       \
  } while (0)

#define B200_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ::b200ctl::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));     \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

#define B200_TRY(expr)              \
  do {                              \
    int rc__ = (expr);              \
    if (rc__ != 0) return rc__;     \
  } while (0)

// ---------------------------------------------------------------- tensor views
enum DType : int { F32 = 0, F64 = 1, I64 = 2, U8 = 3 };

// POD view passed to kernels by value: base pointer + element strides.
struct TView {
  const void* p;
  int64_t n[4];
  int64_t s[4];
  int ndim;
  int dtype;
};

constexpr unsigned M_F32 = 1u << F32, M_F64 = 1u << F64, M_I64 = 1u << I64, M_U8 = 1u << U8;

// Validates a DLTensor and converts it to a TView.  `dev` is in/out: -1 accepts
// any CUDA device and records it, otherwise the tensor must live on that device.
int view_of(const DLTensor* t, const char* name, unsigned dtype_mask, int min_ndim, int max_ndim,
            int* dev, TView* out);

// Drops a trailing extent-1 dimension (the reference passes (N,6,1) / (N,9,1) columns).
inline void squeeze_last(TView& v) {
  if (v.ndim >= 2 && v.n[v.ndim - 1] == 1) v.ndim -= 1;
}

inline bool is_compact(const TView& v) {
  int64_t expect = 1;
  for (int i = v.ndim - 1; i >= 0; --i) {
    if (v.n[i] != 1 && v.s[i] != expect) return false;
    expect *= v.n[i];
  }
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count(int dev);
// CTA slots a persistent grid leaves free on `dev` (b200ctl_reserve_cta_slots): room for a co-resident kernel of another
// stream -- the statistics all-reduce -- so that it does not push one CTA of a one-wave grid into a second wave.
int reserved_slots(int dev);
inline int usable_slots(int dev, int per_sm) {
  const int all = sm_count(dev) * (per_sm > 0 ? per_sm : 1), r = reserved_slots(dev);
  return r < all ? all - r : 1;
}

// Raw device pointers of the ABI (`stats`, `aux`: float64 accumulators that carry no DLTensor descriptor) are checked
// against the driver's own record before a kernel may atomicAdd / store through them: device (or managed) memory of
// device `dev`, 8-byte aligned.  NULL is accepted (the argument is optional).  Verdicts are cached per pointer.
int check_f64_device_ptr(const void* p, const char* name, int dev);

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int enter(int dev) {
    B200_CUDA(cudaGetDevice(&prev));
    if (prev != dev) {
      B200_CUDA(cudaSetDevice(dev));
      switched = true;
    }
    return 0;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

inline int post_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", kernel, cudaGetErrorString(e));
    return (int)e;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Programmatic dependent launch (PDL).  Every kernel of the library starts with pdl_prologue(): it waits for the
// previous kernel of the stream to complete and flush (so stream-order semantics are unchanged -- nothing is
// read before it) and immediately allows the NEXT launch to be scheduled, so that launch's CTAs take SM slots as
// this grid's CTAs retire instead of after a launch bubble.  Measured on the 1M-env PD step: 35.0 -> 33.4 us.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// L2 prefetches issued AHEAD of the dependency wait (the CTAs of a programmatically dependent launch are resident
// while the previous kernel drains): a prefetch consumes nothing, and L2 is the coherence point -- a line the previous
// kernel still writes is updated in place -- so this is always safe; it starts the step's first DRAM reads early, which
// is most of what a 16K-64K-env launch waits for.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifdef B200_NO_PDL      // A/B knob: plain stream-ordered launches
  cfg.numAttrs = 0;
#endif
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_pdl(kernel, grid, dim3((unsigned)block), smem, stream, args...);
}

// dtype-dispatched scalar load / store of a strided element (f32 or f64 storage).
template <typename T>
__device__ __forceinline__ T ld_as(const TView& v, int64_t off) {
  return v.dtype == F32 ? (T) reinterpret_cast<const float*>(v.p)[off]
                        : (T) reinterpret_cast<const double*>(v.p)[off];
}
template <typename T>
__device__ __forceinline__ void st_as(const TView& v, int64_t off, T x) {
  if (v.dtype == F32) reinterpret_cast<float*>(const_cast<void*>(v.p))[off] = (float)x;
  else reinterpret_cast<double*>(const_cast<void*>(v.p))[off] = (double)x;
}

// Streaming 128-bit global accesses: data is touched exactly once per step.
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- async-proxy (TMA) plumbing shared by the tile kernels: mbarrier + 1-D bulk copies
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA 1-D bulk copy shared -> global.  Shared memory written by ordinary stores must be made visible to the
// async proxy first (fence_proxy_async_smem by the writers, then a CTA barrier); the issuing thread commits the
// group and waits until the source has been READ before the CTA may retire its shared memory.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

// End-of-CTA commit of a kernel's statistics: ND fp64 sums and NI integer counts per thread -> stats[slots[...]]
// (slots lists the ND sums first, then the NI counts).  The commit is exposed time -- a persistent grid ends with
// every CTA in it at once -- so it is built for latency: the fp64 sums ride ONE interleaved shuffle butterfly, the
// counts take the integer REDUX unit, and after the cross-warp stage lane j of warp 0 owns entry j, so the CTA issues
// ONE predicated RED instruction.  (The first version ran one butterfly per entry, twice, and then up to five
// branch-separated atomics: 2.9 us of a 6.9 us PD step at 65,536 envs.)  Every thread of the CTA must call it.
// Tried and rejected: folding eight CTAs' sums through distributed shared memory of a thread-block cluster before
// the atomics -- the release/acquire cluster barrier costs a GPU-scope fence per CTA (osc 51.6 -> 67.2 us per
// 262,144 envs, servo 82 -> 93 us per 1M envs); what fixes a tile kernel is committing once per PERSISTENT CTA.
template <int ND, int NI>
__device__ __forceinline__ void block_stats_commit(double (&d)[ND], unsigned (&u)[NI], double* stats,
                                                   const int (&slots)[ND + NI]) {
  static_assert(ND + NI <= 8, "statistics vector has 8 entries");
  __shared__ double s_d[ND][32];
  __shared__ unsigned s_u[NI][32];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;      // blocks are 1-D except the (env, role) blocks of osc_pair_kernel
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < ND; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
  }
#pragma unroll
  for (int k = 0; k < NI; ++k) u[k] = __reduce_add_sync(0xffffffffu, u[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < ND; ++k) s_d[k][warp] = d[k];
#pragma unroll
    for (int k = 0; k < NI; ++k) s_u[k][warp] = u[k];
  }
  __syncthreads();
  if (warp != 0) return;
  // cross-warp stage as a second butterfly over the per-warp partials (lane w holds warp w's): straight-line code.
  // The commit runs once per CTA at the cold tail of a kernel whose loop body fills the instruction cache -- a
  // persistent servo grid of four waves pays it four times per CTA slot (profiles/r02_ab_servo_stats.txt) -- so its
  // SIZE counts: the first version looped over the warps (unrolled sixteen-fold by the compiler: ~250 instructions).
  const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
  for (int k = 0; k < ND; ++k) d[k] = lane < nwarp ? s_d[k][lane] : 0.0;
#pragma unroll
  for (int k = 0; k < NI; ++k) u[k] = lane < nwarp ? s_u[k][lane] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < ND; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
  }
#pragma unroll
  for (int k = 0; k < NI; ++k) u[k] = __reduce_add_sync(0xffffffffu, u[k]);
  // every lane holds the CTA totals; lane j owns entry j: ONE predicated RED instruction per CTA
  double v = 0.0;
  int slot = 0;
#pragma unroll
  for (int j = 0; j < ND; ++j) { v = (lane == j) ? d[j] : v; slot = (lane == j) ? slots[j] : slot; }
#pragma unroll
  for (int j = 0; j < NI; ++j) { v = (lane == ND + j) ? (double)u[j] : v; slot = (lane == ND + j) ? slots[ND + j] : slot; }
  if (lane < ND + NI && v != 0.0) atomicAdd(stats + slot, v);
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The same commit for kernels whose per-thread partial sums are fp32 already (the servo step): a persistent servo grid of
// four waves runs it once per 1.7 tiles, so its instruction count is paid per tile -- the fp64 form above was ~95 warp
// instructions per CTA-warp, a third of what a statistics-carrying 1M-env step executes on top of the plain one
// (ncu smsp__inst_executed 29.7 M vs 25.7 M).  Here: one fp32 butterfly per warp (one SHFL + one FADD per level and
// sum), REDUX for the counts, lane j keeps entry j and stores it with ONE predicated STS; after the barrier lane j of
// warp 0 adds the per-warp partials of ITS entry in fp64 (no second butterfly) and issues the CTA's one RED instruction.
// Counts stay exact (a warp's count is far below 2^24).  Every thread of the CTA must call it; at most 32 warps.
template <int NF, int NI>
__device__ __forceinline__ void block_stats_commit_f32(float (&f)[NF], unsigned (&u)[NI], double* stats,
                                                       const int (&slots)[NF + NI]) {
  static_assert(NF + NI <= 8, "statistics vector has 8 entries");
  __shared__ float s_p[NF + NI][32];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], o);
  }
#pragma unroll
  for (int k = 0; k < NI; ++k) u[k] = __reduce_add_sync(0xffffffffu, u[k]);
  float v = 0.f;
  int slot = 0;
#pragma unroll
  for (int j = 0; j < NF; ++j) { v = (lane == j) ? f[j] : v; slot = (lane == j) ? slots[j] : slot; }
#pragma unroll
  for (int j = 0; j < NI; ++j) { v = (lane == NF + j) ? (float)u[j] : v; slot = (lane == NF + j) ? slots[NF + j] : slot; }
  if (lane < NF + NI) s_p[lane][warp] = v;
  __syncthreads();
  if (warp != 0 || lane >= NF + NI) return;
  const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
  double tot = 0.0;
#pragma unroll 1      // two to eight trips at the cold tail of the kernel: size counts, not trip overhead
  for (int w = 0; w < nwarp; ++w) tot += (double)s_p[lane][w];
  if (tot != 0.0) atomicAdd(stats + slot, tot);
}

#endif  // __CUDACC__

}  // namespace b200ctl
