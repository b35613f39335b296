// Mailbox layout and the publish / consume step of the statistics all-reduce over NVLink peer memory, shared by the
// stand-alone all-reduce kernel (peer.cu) and the control kernels that publish from their own last CTA (pd_torque.cu).
#pragma once
#include "common.cuh"

namespace b200ctl {

constexpr int kRing = 4;
constexpr int kMaxWorld = 16;
constexpr int kSlotWords = 2 * B200CTL_STATS_LEN;
// One (window, source rank) row of a mailbox: every 64-bit word carries 32 bits of payload (the low or the high half of
// one fp64 entry) and, in its upper half, the 32-bit tag of the window it belongs to.  An aligned 64-bit store is
// single-copy atomic, so a word whose tag matches IS its payload: the exchange needs no fence and no separate flag
// (the protocol NCCL calls LL).  A system-scope fence in the publishing CTA stalled the SM it shares with the control
// kernel's working CTAs -- and one late SM is a late kernel for a statically partitioned persistent grid.
struct __align__(128) MailSlot {
  unsigned long long w[kSlotWords];      // w[2 j] = tag : lo32(v[j]), w[2 j + 1] = tag : hi32(v[j]); tag = (u32)(window + 1), 0 = never written
};
struct Mailbox {
  MailSlot slot[kRing][kMaxWorld];
  unsigned long long timeouts;     // consume deadlines missed (the sum then holds NaN): a peer died or never published
  unsigned long long step;         // in-kernel form: the window number, advanced by the publisher CTA of every launch
};
struct PeerTable { Mailbox* box[kMaxWorld]; };

#ifdef __CUDACC__
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Executed by ALL threads of ONE CTA (blockDim.x >= 32; contains __syncthreads()); thread t < count holds entry t of the
// rank's partial vector in `v`.  Publishes the vector as `window` to every rank, then consumes `window` (lagged = 0) or
// `window - 1` (lagged = 1) from this rank's own mailbox into out[0..count): the rows are added in rank order, the same
// order on every rank, so all ranks hold the same bits.
__device__ __forceinline__ void peer_publish_consume(const PeerTable& peers, int rank, int world, unsigned long long window,
                                                     int lagged, int count, double v, double* out, long long deadline_ns) {
  const int t = threadIdx.x;
  __shared__ unsigned s_half[kMaxWorld * kSlotWords];
  __shared__ int s_ok;
  // ---- publish: thread t < 2 count owns word t of this rank's row (entry t / 2 comes from lane t / 2 of warp 0)
  if (t < 32) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long src = __shfl_sync(0xffffffffu, bits, (t >> 1) & 31);
    if (t < 2 * count) {
      const unsigned half = (t & 1) ? (unsigned)(src >> 32) : (unsigned)src;
      const unsigned long long word = ((unsigned long long)(unsigned)(window + 1) << 32) | half;
      const int ring = (int)(window % kRing);
      for (int p = 0; p < world; ++p) st_relaxed_sys(&peers.box[p]->slot[ring][rank].w[t], word);
    }
  }
  if (t == 0) s_ok = 1;
  if (lagged && window == 0) {      // nothing older to consume yet: the reduced vector of "window -1" is zero
    if (t < count) out[t] = 0.0;
    return;
  }
  // ---- consume
  const unsigned long long cw = lagged ? window - 1 : window;
  const unsigned tag = (unsigned)(cw + 1);
  const Mailbox* mine = peers.box[rank];
  const int cring = (int)(cw % kRing);
  __syncthreads();
  for (int i = t; i < world * 2 * count; i += blockDim.x) {
    const int p = i / (2 * count), k = i - p * 2 * count;
    const unsigned long long* src = &mine->slot[cring][p].w[k];
    unsigned long long word = ld_relaxed_sys(src);
    if ((unsigned)(word >> 32) != tag) {
      unsigned long long t0, now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do {
        __nanosleep(64);
        word = ld_relaxed_sys(src);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      } while ((unsigned)(word >> 32) != tag && (long long)(now - t0) <= deadline_ns);
      if ((unsigned)(word >> 32) != tag) {      // never hang the GPU on a dead peer
        s_ok = 0;
        if (k == 0) atomicAdd(const_cast<unsigned long long*>(&mine->timeouts), 1ull);
      }
    }
    s_half[p * kSlotWords + k] = (unsigned)word;
  }
  __syncthreads();
  if (t < count) {
    double s = 0.0;
    for (int p = 0; p < world; ++p)
      s += __hiloint2double((int)s_half[p * kSlotWords + 2 * t + 1], (int)s_half[p * kSlotWords + 2 * t]);
    out[t] = s_ok ? s : __longlong_as_double(0x7ff8000000000000ll);
  }
}
#endif  // __CUDACC__

}  // namespace b200ctl
