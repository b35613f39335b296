// Statistics all-reduce over NVLink peer memory: the library's own collective for the one exchange step of the path.
//
// Every rank (one process per GPU of one NVSwitch box) owns a MAILBOX in its device memory and maps every peer's
// mailbox through CUDA IPC.  One tiny kernel per window, launched IN ORDER on the control stream (programmatic dependent
// launch like every other kernel of the library: no stream hop, no second process-group stream):
//   1. publish   the rank's partial statistics vector goes into row [window % 4][rank] of EVERY rank's mailbox as sixteen
//                self-validating 64-bit words (32 bits of payload + the 32-bit window tag, peer.cuh): relaxed stores over
//                NVLink, no fence, no separate flag;
//   2. consume   it polls (with a deadline) the `world` rows of the window it consumes in ITS OWN mailbox until every
//                word carries that window's tag, and sums the rows in rank order -- the same order on every rank, so
//                all ranks hold bit-identical sums.
// `lagged` = 0 consumes the window it just published (a classic all-reduce: one NVLink round trip on the critical
// path); `lagged` = 1 consumes the PREVIOUS window, whose rows arrived a whole window ago -- the collective then never
// waits for a peer and its cost is the launch alone, which is what a per-step statistics exchange needs.
// A ring of four slots per rank: a rank can publish window w + 1 only after it consumed w - 1 (lagged) from everyone,
// i.e. after every peer published w - 1 and therefore finished reading w - 2 or older; slots w + 1 and w - 3 coincide.
// NCCL's all-reduce of the same 64 bytes costs ~34 us per call on the control stream of this box (two stream hops and a
// 640-thread kernel that needs an empty SM); this one costs ~3 us.
#include "peer.cuh"

#include <string.h>

namespace b200ctl {

__global__ void __launch_bounds__(64)
stats_allreduce_peer_kernel(PeerTable peers, int rank, int world, unsigned long long window, int lagged, int count,
                            double* stats, double* __restrict__ zero_after, double* out, int clear_source,
                            long long deadline_ns) {
  pdl_prologue();
  const int t = threadIdx.x;
  // the accumulator of the NEXT window is cleared here, so that the step loop needs no separate fill kernel (an
  // ordinary launch that would break the chain of programmatically dependent launches)
  if (zero_after && t < B200CTL_STATS_LEN) zero_after[t] = 0.0;
  double v = 0.0;
  if (t < count) {
    v = stats[t];
    // out-of-place form: the source accumulator is complete (its kernels finished), so it can be recycled right here
    // -- the step loop then alternates two accumulators while this kernel runs NEXT to the following step
    if (clear_source) stats[t] = 0.0;
  }
  peer_publish_consume(peers, rank, world, window, lagged, count, v, out, deadline_ns);
}

}  // namespace b200ctl

using namespace b200ctl;

extern "C" int b200ctl_peer_mailbox_create(int32_t device, void** mailbox_out, void* ipc_handle_out_64_bytes) {
  if (!mailbox_out || !ipc_handle_out_64_bytes) B200_FAIL(B200CTL_E_NULL, "output pointer is NULL");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard g;
  B200_TRY(g.enter(device));
  void* p = nullptr;
  B200_CUDA(cudaMalloc(&p, sizeof(Mailbox)));
  B200_CUDA(cudaMemset(p, 0, sizeof(Mailbox)));
  B200_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  B200_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_out_64_bytes, &h, sizeof(h));
  *mailbox_out = p;
  return 0;
}

extern "C" int b200ctl_peer_mailbox_open(int32_t device, const void* ipc_handle_64_bytes, void** peer_out) {
  if (!ipc_handle_64_bytes || !peer_out) B200_FAIL(B200CTL_E_NULL, "handle / output pointer is NULL");
  DeviceGuard g;
  B200_TRY(g.enter(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64_bytes, sizeof(h));
  B200_CUDA(cudaIpcOpenMemHandle(peer_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int b200ctl_peer_mailbox_close(int32_t device, void* mailbox, int32_t is_peer) {
  if (!mailbox) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(device));
  if (is_peer) B200_CUDA(cudaIpcCloseMemHandle(mailbox));
  else B200_CUDA(cudaFree(mailbox));
  return 0;
}

extern "C" int b200ctl_peer_mailbox_timeouts(int32_t device, const void* mailbox, uint64_t* count_out) {
  if (!mailbox || !count_out) B200_FAIL(B200CTL_E_NULL, "mailbox / output pointer is NULL");
  DeviceGuard g;
  B200_TRY(g.enter(device));
  B200_CUDA(cudaMemcpy(count_out, &static_cast<const Mailbox*>(mailbox)->timeouts, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int b200ctl_stats_allreduce_peer(void* const* mailboxes, int32_t rank, int32_t world, uint64_t window,
                                            int32_t lagged, double* stats, int32_t count, double* zero_after,
                                            double* out, double timeout_s, int32_t device, b200ctl_stream_t stream) {
  if (!mailboxes || !stats) B200_FAIL(B200CTL_E_NULL, "mailboxes / stats is NULL");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) B200_FAIL(B200CTL_E_VALUE, "bad rank %d / world %d (max %d)", rank, world, kMaxWorld);
  if (count < 1 || count > B200CTL_STATS_LEN) B200_FAIL(B200CTL_E_VALUE, "count must be in [1, %d]", B200CTL_STATS_LEN);
  B200_TRY(check_f64_device_ptr(stats, "stats", device));
  B200_TRY(check_f64_device_ptr(zero_after, "zero_after", device));
  B200_TRY(check_f64_device_ptr(out, "out", device));
  if (zero_after == stats || (zero_after && zero_after == out)) B200_FAIL(B200CTL_E_ALIAS, "zero_after must not be the vector being reduced / written");
  const int clear_source = out != nullptr && out != stats;
  if (!out) out = stats;
  PeerTable T{};
  for (int p = 0; p < world; ++p) {
    if (!mailboxes[p]) B200_FAIL(B200CTL_E_NULL, "mailbox of rank %d is NULL", p);
    T.box[p] = static_cast<Mailbox*>(mailboxes[p]);
  }
  DeviceGuard g;
  B200_TRY(g.enter(device));
  const long long deadline_ns = (long long)((timeout_s > 0 ? timeout_s : 2.0) * 1e9);
  launch_pdl(stats_allreduce_peer_kernel, 1, 64, 0, (cudaStream_t)stream, T, (int)rank, (int)world,
             (unsigned long long)window, lagged ? 1 : 0, (int)count, stats, zero_after, out, clear_source, deadline_ns);
  return post_launch("stats_allreduce_peer_kernel");
}
