// libb200ctl runtime: error channel, DLTensor validation, device info, row
// gather, and the NCCL statistics all-reduce (NCCL resolved with dlsym so the
// library has no link-time dependency on libnccl).
#include "common.cuh"

#include <dlfcn.h>
#include <string.h>
#include <mutex>

namespace b200ctl {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int view_of(const DLTensor* t, const char* name, unsigned dtype_mask, int min_ndim, int max_ndim,
            int* dev, TView* out) {
  if (!t) B200_FAIL(B200CTL_E_NULL, "%s: required tensor is NULL", name);
  if (t->device.device_type != kDLCUDA)
    B200_FAIL(B200CTL_E_DEVICE, "%s: tensor must be on a CUDA device (device_type=%d); there is no CPU path",
              name, (int)t->device.device_type);
  if (*dev < 0) *dev = t->device.device_id;
  else if (*dev != t->device.device_id)
    B200_FAIL(B200CTL_E_DEVICE, "%s: on cuda:%d, expected cuda:%d", name, t->device.device_id, *dev);
  int dt = -1;
  if (t->dtype.lanes == 1 && t->dtype.code == kDLFloat && t->dtype.bits == 32) dt = F32;
  else if (t->dtype.lanes == 1 && t->dtype.code == kDLFloat && t->dtype.bits == 64) dt = F64;
  else if (t->dtype.lanes == 1 && t->dtype.code == kDLInt && t->dtype.bits == 64) dt = I64;
  else if (t->dtype.lanes == 1 && (t->dtype.code == kDLUInt || t->dtype.code == 6 /* kDLBool */) && t->dtype.bits == 8) dt = U8;
  if (dt < 0 || !((1u << dt) & dtype_mask))
    B200_FAIL(B200CTL_E_DTYPE, "%s: unsupported dtype (code=%d bits=%d)", name, (int)t->dtype.code, (int)t->dtype.bits);
  if (t->ndim < min_ndim || t->ndim > max_ndim || t->ndim > 4)
    B200_FAIL(B200CTL_E_SHAPE, "%s: rank %d not in [%d,%d]", name, t->ndim, min_ndim, max_ndim);
  if (t->ndim > 0 && !t->shape) B200_FAIL(B200CTL_E_NULL, "%s: shape is NULL", name);
  out->ndim = t->ndim;
  out->dtype = dt;
  int64_t compact = 1;
  for (int i = 0; i < 4; ++i) { out->n[i] = 1; out->s[i] = 0; }
  for (int i = t->ndim - 1; i >= 0; --i) {
    out->n[i] = t->shape[i];
    if (t->shape[i] < 0) B200_FAIL(B200CTL_E_SHAPE, "%s: negative extent", name);
    out->s[i] = t->strides ? t->strides[i] : compact;
    compact *= t->shape[i];
  }
  const size_t esz = dt == F32 ? 4 : (dt == U8 ? 1 : 8);
  if (t->byte_offset % esz) B200_FAIL(B200CTL_E_LAYOUT, "%s: byte_offset not element aligned", name);
  out->p = static_cast<const char*>(t->data) + t->byte_offset;
  if (!t->data && compact > 0) B200_FAIL(B200CTL_E_NULL, "%s: data pointer is NULL", name);
  if (reinterpret_cast<uintptr_t>(out->p) % esz) B200_FAIL(B200CTL_E_LAYOUT, "%s: data pointer not element aligned", name);
  return 0;
}

int sm_count(int dev) {
  static int cache[64];
  static std::once_flag once;
  std::call_once(once, [] { for (int& c : cache) c = 0; });
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

static std::atomic<int> g_reserved[64];
int reserved_slots(int dev) { return (dev >= 0 && dev < 64) ? g_reserved[dev].load(std::memory_order_relaxed) : 0; }

int check_f64_device_ptr(const void* p, const char* name, int dev) {
  if (!p) return 0;
  if (reinterpret_cast<uintptr_t>(p) & 7u) B200_FAIL(B200CTL_E_LAYOUT, "%s: pointer is not 8-byte aligned (float64 expected)", name);
  // small direct-mapped cache of pointers already vouched for: the step loop passes the same two or three buffers
  // forever, and a hit also keeps the driver query out of CUDA-graph capture (the warm-up call populates it)
  struct Slot { std::atomic<uintptr_t> key{0}; };
  static Slot cache[64];
  const uintptr_t key = reinterpret_cast<uintptr_t>(p) ^ ((uintptr_t)(dev + 1) << 56);
  Slot& slot = cache[(reinterpret_cast<uintptr_t>(p) >> 6) & 63];
  if (slot.key.load(std::memory_order_relaxed) == key) return 0;
  cudaPointerAttributes a{};
  const cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    B200_FAIL(B200CTL_E_DEVICE, "%s: not a CUDA pointer (%s); b200ctl has no CPU path", name, cudaGetErrorString(e));
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    B200_FAIL(B200CTL_E_DEVICE, "%s: expected device memory (a float64 CUDA tensor), got %s memory", name,
              a.type == cudaMemoryTypeHost ? "pinned host" : "unregistered host");
  if (a.type == cudaMemoryTypeDevice && a.device != dev)
    B200_FAIL(B200CTL_E_DEVICE, "%s: on cuda:%d, expected cuda:%d", name, a.device, dev);
  slot.key.store(key, std::memory_order_relaxed);
  return 0;
}

// ---------------------------------------------------------------- gather_rows
__global__ void gather_rows_kernel(TView src, TView idx, int col0, int ncols, TView dst, int64_t n) {
  const int64_t total = n * ncols;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ncols;
    const int j = (int)(e - i * ncols);
    const int64_t row = reinterpret_cast<const int64_t*>(idx.p)[i * idx.s[0]];
    // bit-exact copy: move the 32-bit pattern, no arithmetic.  A row outside the source is not dereferenced (NaN).
    const uint32_t bits = (row >= 0 && row < src.n[0])
        ? reinterpret_cast<const uint32_t*>(src.p)[row * src.s[0] + (col0 + j) * src.s[1]] : 0x7fc00000u;
    reinterpret_cast<uint32_t*>(const_cast<void*>(dst.p))[i * dst.s[0] + j * dst.s[1]] = bits;
  }
}

}  // namespace b200ctl

using namespace b200ctl;

extern "C" int b200ctl_version(void) { return B200CTL_VERSION; }
extern "C" const char* b200ctl_last_error(void) { return t_error; }
extern "C" uint64_t b200ctl_launch_count(void) { return g_launch_count.load(); }

extern "C" int b200ctl_reserve_cta_slots(int32_t device, int32_t slots) {
  if (device < 0 || device >= 64) B200_FAIL(B200CTL_E_DEVICE, "device %d out of range", device);
  if (slots < 0 || slots > 4096) B200_FAIL(B200CTL_E_VALUE, "slots must be in [0, 4096]");
  g_reserved[device].store(slots, std::memory_order_relaxed);
  return 0;
}

extern "C" int b200ctl_gather_rows(const DLTensor* src, const DLTensor* index, int32_t col0, int32_t ncols,
                                   DLTensor* dst, b200ctl_stream_t stream) {
  int dev = -1;
  TView s, i, d;
  B200_TRY(view_of(src, "src", M_F32, 2, 2, &dev, &s));
  B200_TRY(view_of(index, "index", M_I64, 1, 1, &dev, &i));
  B200_TRY(view_of(dst, "dst", M_F32, 2, 2, &dev, &d));
  const int64_t n = i.n[0];
  if (d.n[0] != n || d.n[1] != ncols) B200_FAIL(B200CTL_E_SHAPE, "dst must be (%lld,%d)", (long long)n, ncols);
  if (col0 < 0 || ncols <= 0 || col0 + ncols > s.n[1]) B200_FAIL(B200CTL_E_VALUE, "column window out of range");
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  const int64_t total = n * ncols;
  const int block = 256;
  const int grid = (int)((total + block - 1) / block < (int64_t)sm_count(dev) * 16 ? (total + block - 1) / block
                                                                                   : (int64_t)sm_count(dev) * 16);
  gather_rows_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(s, i, col0, ncols, d, n);
  return post_launch("gather_rows_kernel");
}

// ---------------------------------------------------------------- NCCL (dlsym)
namespace {
struct NcclId { char b[128]; };   // ncclUniqueId: 128 opaque bytes, passed by value
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    // Prefer the copy already mapped into the process (torch's), then the usual sonames.
    const char* names[] = {nullptr, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      void* h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL | (nm ? 0 : RTLD_NOLOAD));
      if (!h) continue;
      void* sym = dlsym(h, "ncclAllReduce");
      if (!sym) continue;
      n.h = h;
      n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      n.CommInitRank = (decltype(n.CommInitRank))dlsym(h, "ncclCommInitRank");
      n.CommDestroy = (decltype(n.CommDestroy))dlsym(h, "ncclCommDestroy");
      n.AllReduce = (decltype(n.AllReduce))sym;
      n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
      n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce;
      if (n.ok) break;
    }
  });
  return n;
}
int nccl_fail(const char* what, int rc) {
  Nccl& n = nccl();
  set_error("%s: %s", what, n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
  return B200CTL_E_NCCL;
}
}  // namespace

extern "C" int b200ctl_nccl_unique_id(void* id_out) {
  if (!id_out) B200_FAIL(B200CTL_E_NULL, "id_out is NULL");
  Nccl& n = nccl();
  if (!n.ok) B200_FAIL(B200CTL_E_NCCL, "libnccl not found in the process or on the loader path");
  int rc = n.GetUniqueId(id_out);
  return rc ? nccl_fail("ncclGetUniqueId", rc) : 0;
}

extern "C" int b200ctl_nccl_comm_init(void** comm_out, int32_t world_size, const void* id, int32_t rank) {
  if (!comm_out || !id) B200_FAIL(B200CTL_E_NULL, "comm_out / id is NULL");
  Nccl& n = nccl();
  if (!n.ok) B200_FAIL(B200CTL_E_NCCL, "libnccl not found in the process or on the loader path");
  NcclId uid;
  memcpy(uid.b, id, 128);
  int rc = n.CommInitRank(comm_out, world_size, uid, rank);
  return rc ? nccl_fail("ncclCommInitRank", rc) : 0;
}

extern "C" int b200ctl_nccl_comm_destroy(void* comm) {
  Nccl& n = nccl();
  if (!n.ok || !comm) return 0;
  int rc = n.CommDestroy(comm);
  return rc ? nccl_fail("ncclCommDestroy", rc) : 0;
}

extern "C" int b200ctl_stats_allreduce(void* comm, double* stats, int32_t count, b200ctl_stream_t stream) {
  if (!comm || !stats) B200_FAIL(B200CTL_E_NULL, "comm / stats is NULL");
  Nccl& n = nccl();
  if (!n.ok) B200_FAIL(B200CTL_E_NCCL, "libnccl not found in the process or on the loader path");
  // ncclDouble = 8 (ncclFloat64), ncclSum = 0 in nccl.h (stable since NCCL 2.0)
  int rc = n.AllReduce(stats, stats, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
  return rc ? nccl_fail("ncclAllReduce", rc) : 0;
}
