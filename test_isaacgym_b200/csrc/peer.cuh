// Mailbox layout and the publish / consume step of the statistics all-reduce over NVLink peer memory, shared by the
// stand-alone all-reduce kernel (peer.cu) and the control kernels that publish from their own last CTA (pd_torque.cu).
#pragma once
#include "common.cuh"

namespace b200ctl {

constexpr int kRing = 4;
constexpr int kMaxWorld = 16;
struct __align__(128) MailSlot {
  double v[B200CTL_STATS_LEN];
  unsigned long long stamp;        // window + 1 once v[] is complete (0 = never written)
  unsigned long long pad[7];
};
struct Mailbox {
  MailSlot slot[kRing][kMaxWorld];
  unsigned long long timeouts;     // consume deadlines missed (the sum then holds NaN): a peer died or never published
  unsigned long long step;         // fused form: the window number, advanced by the publishing kernel's last CTA
  unsigned ticket;                 // fused form: CTAs of the current launch that have committed their statistics
};
struct PeerTable { Mailbox* box[kMaxWorld]; };

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Executed by (at least) the first max(count, world) threads of ONE CTA, all of which hold their entry `v` (thread t:
// entry t of the rank's partial vector, t < count).  Publishes the vector as `window` to every rank, then consumes
// `window` (lagged = 0) or `window - 1` (lagged = 1) from this rank's own mailbox into out[0..count).  Contains two
// __syncthreads(): every thread of the CTA must call it.
__device__ __forceinline__ void peer_publish_consume(const PeerTable& peers, int rank, int world, unsigned long long window,
                                                     int lagged, int count, double v, double* out, long long deadline_ns) {
  const int t = threadIdx.x;
  const int ring = (int)(window % kRing);
  if (t < count) {
    for (int p = 0; p < world; ++p) peers.box[p]->slot[ring][rank].v[t] = v;
    __threadfence_system();
  }
  __syncthreads();
  if (t < world) st_release_sys(&peers.box[t]->slot[ring][rank].stamp, window + 1);
  __shared__ int s_ok;
  if (t == 0) s_ok = 1;
  if (lagged && window == 0) {      // nothing older to consume yet: the reduced vector of "window -1" is zero
    __syncthreads();
    if (t < count) out[t] = 0.0;
    return;
  }
  const unsigned long long cw = lagged ? window - 1 : window;
  const int cring = (int)(cw % kRing);
  Mailbox* mine = peers.box[rank];
  __syncthreads();
  if (t < world) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(&mine->slot[cring][t].stamp) != cw + 1) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if ((long long)(now - t0) > deadline_ns) {      // never hang the GPU on a dead peer
        s_ok = 0;
        atomicAdd(&mine->timeouts, 1ull);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (t < count) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += *reinterpret_cast<volatile double*>(&mine->slot[cring][p].v[t]);
    out[t] = s_ok ? s : __longlong_as_double(0x7ff8000000000000ll);
  }
}
#endif  // __CUDACC__

}  // namespace b200ctl
