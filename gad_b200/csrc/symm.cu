// symm.cu — symmetric memory for the data-parallel exchange (north-star subsystem 5): one region per rank, mapped
// into every process of the communicator, and — where the fabric has it — bound to ONE multicast object, so that a
// kernel can sum an address over all ranks inside the NVSwitch (multimem.ld_reduce) and store to all ranks with one
// instruction (multimem.st). The reference has nothing of the kind (single process: SURVEY §0); the seam it serves is the
// per-example gradient sum of DiffNet::apply_gradient_step (gad/src/net_ext.rs:62-70).
//
// Two ways to get there, tried in this order, every rank taking the same branch (each step ends in a min-all-reduce of
// "did it work here"):
//   1. CUDA virtual memory management: cuMemCreate (POSIX file-descriptor handles) -> the descriptors travel between the
//      processes over abstract unix sockets (SCM_RIGHTS) -> cuMemImportFromShareableHandle + cuMemMap on every peer.
//      Multicast on top: rank 0 creates the object (cuMulticastCreate), everybody adds its device, binds its own
//      allocation at offset 0 and maps the object.
//   2. cudaMalloc + CUDA IPC memory handles (what round 1 used): peer mappings only, no multicast.
// The driver entry points are looked up at run time (cudaGetDriverEntryPoint), like cuTensorMapEncodeTiled in gemm_tc.cu:
// the library does not link libcuda.
#include "symm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "core.h"

namespace gadcu {
namespace {

struct Driver {
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  bool vmm = false, multicast = false;
};

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name, void* fn_ptr_ptr) {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p) {
        *reinterpret_cast<void**>(fn_ptr_ptr) = p;
        return true;
      }
      cudaGetLastError();
      return false;
    };
    bool v = true, m = true;
    v &= get("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity);
    v &= get("cuMemCreate", &d.MemCreate);
    v &= get("cuMemRelease", &d.MemRelease);
    v &= get("cuMemExportToShareableHandle", &d.MemExportToShareableHandle);
    v &= get("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle);
    v &= get("cuMemAddressReserve", &d.MemAddressReserve);
    v &= get("cuMemAddressFree", &d.MemAddressFree);
    v &= get("cuMemMap", &d.MemMap);
    v &= get("cuMemUnmap", &d.MemUnmap);
    v &= get("cuMemSetAccess", &d.MemSetAccess);
    v &= get("cuDeviceGet", &d.DeviceGet);
    v &= get("cuDeviceGetAttribute", &d.DeviceGetAttribute);
    m &= get("cuMulticastCreate", &d.MulticastCreate);
    m &= get("cuMulticastAddDevice", &d.MulticastAddDevice);
    m &= get("cuMulticastBindMem", &d.MulticastBindMem);
    m &= get("cuMulticastGetGranularity", &d.MulticastGetGranularity);
    d.vmm = v;
    d.multicast = v && m;
  });
  return d;
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

// ---- file descriptors between processes: abstract unix sockets + SCM_RIGHTS ---------------------------------------
sockaddr_un abstract_addr(const std::string& key, int rank, socklen_t* len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  const std::string name = "gadcu-" + key + "-" + std::to_string(rank);
  const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
  std::memcpy(a.sun_path + 1, name.data(), n);  // sun_path[0] == 0: abstract namespace, no file system entry
  *len = socklen_t(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}

bool send_fds(int sock, int rank, const int* fds, int nfds) {
  int header[2] = {rank, nfds};
  iovec iov{header, sizeof(header)};
  char ctrl[CMSG_SPACE(sizeof(int) * 2)];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = CMSG_SPACE(sizeof(int) * size_t(nfds));
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int) * size_t(nfds));
  std::memcpy(CMSG_DATA(c), fds, sizeof(int) * size_t(nfds));
  return sendmsg(sock, &msg, 0) == ssize_t(sizeof(header));
}

bool recv_fds(int sock, int* rank, int* fds, int* nfds) {
  int header[2] = {-1, 0};
  iovec iov{header, sizeof(header)};
  char ctrl[CMSG_SPACE(sizeof(int) * 2)];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, MSG_WAITALL) != ssize_t(sizeof(header))) return false;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return false;
  const int got = int((c->cmsg_len - CMSG_LEN(0)) / sizeof(int));
  if (got != header[1] || got < 1 || got > 2) return false;
  std::memcpy(fds, CMSG_DATA(c), sizeof(int) * size_t(got));
  *rank = header[0];
  *nfds = got;
  return true;
}

// every rank hands `mine` (1 or 2 descriptors; rank 0's second one is the multicast object) to every other rank
bool exchange_fds(const std::string& key, int nranks, int rank, const int* mine, int n_mine, int* peer_fd /*[nranks]*/, int* mc_fd,
                  const SymmCollectives& coll) {
  int listener = socket(AF_UNIX, SOCK_STREAM, 0);
  bool ok = listener >= 0;
  if (ok) {
    socklen_t len;
    sockaddr_un a = abstract_addr(key, rank, &len);
    ok = bind(listener, reinterpret_cast<sockaddr*>(&a), len) == 0 && listen(listener, 2 * kMaxRanks) == 0;
  }
  if (coll.allreduce_min(ok ? 1 : 0) == 0) {  // (also the barrier: everybody is listening before anybody connects)
    if (listener >= 0) close(listener);
    return false;
  }
  for (int p = 0; p < nranks && ok; p++) {
    if (p == rank) continue;
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len;
    sockaddr_un a = abstract_addr(key, p, &len);
    ok = s >= 0 && connect(s, reinterpret_cast<sockaddr*>(&a), len) == 0 && send_fds(s, rank, mine, n_mine);
    if (s >= 0) close(s);
  }
  for (int i = 0; i + 1 < nranks && ok; i++) {
    pollfd pf{listener, POLLIN, 0};
    if (poll(&pf, 1, 30000) <= 0) { ok = false; break; }  // a peer that never shows up must not hang the job
    int c = accept(listener, nullptr, nullptr);
    int src = -1, fds[2] = {-1, -1}, n = 0;
    ok = c >= 0 && recv_fds(c, &src, fds, &n) && src >= 0 && src < nranks && src != rank;
    if (ok) {
      peer_fd[src] = fds[0];
      if (n == 2) {
        if (src == 0) *mc_fd = fds[1];
        else close(fds[1]);
      }
    }
    if (c >= 0) close(c);
  }
  close(listener);
  return coll.allreduce_min(ok ? 1 : 0) == 1;
}

void unmap_all(SymmRegion& r) {
  Driver& d = driver();
  if (!r.vmm) return;
  for (int p = 0; p < r.nranks; p++) {
    if (r.peer[p]) {
      d.MemUnmap(CUdeviceptr(r.peer[p]), r.mapped_bytes);
      d.MemAddressFree(CUdeviceptr(r.peer[p]), r.mapped_bytes);
      r.peer[p] = nullptr;
    }
    if (r.mem_handle[p]) {
      d.MemRelease(CUmemGenericAllocationHandle(r.mem_handle[p]));
      r.mem_handle[p] = 0;
    }
  }
  if (r.mc) {
    d.MemUnmap(CUdeviceptr(r.mc), r.mapped_bytes);
    d.MemAddressFree(CUdeviceptr(r.mc), r.mapped_bytes);
    r.mc = nullptr;
  }
  if (r.mc_handle) {
    d.MemRelease(CUmemGenericAllocationHandle(r.mc_handle));
    r.mc_handle = 0;
  }
  r.local = nullptr;
}

bool map_handle(Driver& d, CUmemGenericAllocationHandle h, size_t size, size_t gran, int device, void** out) {
  CUdeviceptr va = 0;
  if (d.MemAddressReserve(&va, size, gran, 0, 0) != CUDA_SUCCESS) return false;
  if (d.MemMap(va, size, 0, h, 0) != CUDA_SUCCESS) {
    d.MemAddressFree(va, size);
    return false;
  }
  CUmemAccessDesc acc;
  std::memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) {
    d.MemUnmap(va, size);
    d.MemAddressFree(va, size);
    return false;
  }
  *out = reinterpret_cast<void*>(va);
  return true;
}

// 1 = region with multicast, 2 = region without, 0 = the VMM route does not work here (every rank returns the same)
int alloc_vmm(SymmRegion& r, int device, int nranks, int rank, size_t bytes, bool want_multicast, const std::string& key,
              const SymmCollectives& coll) {
  Driver& d = driver();
  static const bool vmm_off = [] { const char* e = getenv("GADCU_SYMM"); return e && std::string(e) == "ipc"; }();
  bool ok = d.vmm && !vmm_off;
  CUdevice dev = 0;
  bool mc = ok && want_multicast && d.multicast;
  if (ok) ok = d.DeviceGet(&dev, device) == CUDA_SUCCESS;
  if (ok && mc) {
    int supported = 0;
    mc = d.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && supported != 0;
  }
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  if (ok) ok = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran > 0;
  // multicast wanted by everybody or by nobody (one rank without support switches it off for all)
  mc = coll.allreduce_min(ok && mc ? 1 : 0) == 1;
  CUmulticastObjectProp mprop;
  std::memset(&mprop, 0, sizeof(mprop));
  if (ok && mc) {
    mprop.numDevices = unsigned(nranks);
    mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mprop.size = round_up(bytes, gran);
    size_t mgran = 0;
    if (d.MulticastGetGranularity(&mgran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mgran > 0) gran = std::max(gran, mgran);
    else mc = false;
  }
  mc = coll.allreduce_min(ok && mc ? 1 : 0) == 1;
  const size_t size = ok ? round_up(bytes, gran) : 0;
  mprop.size = size;
  CUmemGenericAllocationHandle mine = 0, mch = 0;
  int fds[2] = {-1, -1};
  if (ok) ok = d.MemCreate(&mine, size, &prop, 0) == CUDA_SUCCESS;
  if (ok) ok = d.MemExportToShareableHandle(&fds[0], mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
  bool mc_here = mc;
  if (ok && mc && rank == 0)
    mc_here = d.MulticastCreate(&mch, &mprop) == CUDA_SUCCESS &&
              d.MemExportToShareableHandle(&fds[1], mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
  mc = coll.allreduce_min(mc_here ? 1 : 0) == 1;
  if (!mc && mch) { d.MemRelease(mch); mch = 0; if (fds[1] >= 0) { close(fds[1]); fds[1] = -1; } }
  auto give_up = [&] {
    if (fds[0] >= 0) close(fds[0]);
    if (fds[1] >= 0) close(fds[1]);
    if (mch) d.MemRelease(mch);
    if (mine) d.MemRelease(mine);
    return 0;
  };
  if (coll.allreduce_min(ok ? 1 : 0) == 0) return give_up();
  // the descriptors travel
  int peer_fd[kMaxRanks], mc_fd = -1;
  for (int& f : peer_fd) f = -1;
  if (!exchange_fds(key, nranks, rank, fds, (rank == 0 && mc) ? 2 : 1, peer_fd, &mc_fd, coll)) {
    for (int f : peer_fd) if (f >= 0) close(f);
    if (mc_fd >= 0) close(mc_fd);
    return give_up();
  }
  close(fds[0]);
  fds[0] = -1;
  if (fds[1] >= 0) { close(fds[1]); fds[1] = -1; }
  // import + map every rank's memory
  r.vmm = true;
  r.nranks = nranks;
  r.rank = rank;
  r.bytes = bytes;
  r.mapped_bytes = size;
  r.mem_handle[rank] = mine;
  for (int p = 0; p < nranks && ok; p++) {
    CUmemGenericAllocationHandle h = mine;
    if (p != rank) {
      ok = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(intptr_t(peer_fd[p])), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
      if (ok) r.mem_handle[p] = h;
    }
    if (ok) ok = map_handle(d, h, size, gran, device, &r.peer[p]);
  }
  for (int f : peer_fd) if (f >= 0) close(f);
  r.local = r.peer[rank];
  if (coll.allreduce_min(ok ? 1 : 0) == 0) {
    if (mc_fd >= 0) close(mc_fd);
    if (mch) d.MemRelease(mch);
    unmap_all(r);
    r = SymmRegion{};
    return 0;
  }
  if (!mc) return 2;
  // multicast: import, add every device, THEN bind (all devices must have been added before the first bind), then map
  bool m = true;
  if (rank != 0) m = mc_fd >= 0 && d.MemImportFromShareableHandle(&mch, reinterpret_cast<void*>(intptr_t(mc_fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
  if (mc_fd >= 0) close(mc_fd);
  if (m) m = d.MulticastAddDevice(mch, dev) == CUDA_SUCCESS;
  bool all = coll.allreduce_min(m ? 1 : 0) == 1;
  if (all) {
    m = d.MulticastBindMem(mch, 0, mine, 0, size, 0) == CUDA_SUCCESS;
    all = coll.allreduce_min(m ? 1 : 0) == 1;
  }
  if (all) {
    m = map_handle(d, mch, size, gran, device, &r.mc);
    all = coll.allreduce_min(m ? 1 : 0) == 1;
    if (!all && r.mc) {
      d.MemUnmap(CUdeviceptr(r.mc), size);
      d.MemAddressFree(CUdeviceptr(r.mc), size);
      r.mc = nullptr;
    }
  }
  if (!all) {
    if (mch) d.MemRelease(mch);
    return 2;
  }
  r.mc_handle = mch;
  return 1;
}

void alloc_ipc(SymmRegion& r, int device, int nranks, int rank, size_t bytes, const SymmCollectives& coll) {
  (void)device;
  void* local = nullptr;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  bool ok = cudaMalloc(&local, bytes) == cudaSuccess && cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
  if (coll.allreduce_min(ok ? 1 : 0) == 0) {
    if (local) cudaFree(local);
    cudaGetLastError();
    fail(GADCU_ERR_UNSUPPORTED, "symmetric memory: neither CUDA VMM handles nor CUDA IPC work between these processes");
  }
  std::vector<cudaIpcMemHandle_t> all{static_cast<size_t>(nranks)};
  coll.allgather(&mine, all.data(), sizeof(mine));
  r = SymmRegion{};
  r.bytes = bytes;
  r.nranks = nranks;
  r.rank = rank;
  r.local = local;
  for (int p = 0; p < nranks && ok; p++) {
    if (p == rank) { r.peer[p] = local; continue; }
    ok = cudaIpcOpenMemHandle(&r.peer[p], all[size_t(p)], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  }
  if (coll.allreduce_min(ok ? 1 : 0) == 0) {
    cudaGetLastError();
    symm_free(r);
    fail(GADCU_ERR_UNSUPPORTED, "symmetric memory: cudaIpcOpenMemHandle failed (no peer access between the GPUs?)");
  }
}

}  // namespace

void symm_alloc(SymmRegion& r, int device, int nranks, int rank, size_t bytes, bool want_multicast, const std::string& key,
                const SymmCollectives& coll, std::string* how) {
  if (nranks > kMaxRanks) fail(GADCU_ERR_UNSUPPORTED, "symmetric memory: at most 8 ranks");
  r = SymmRegion{};
  const int got = alloc_vmm(r, device, nranks, rank, bytes, want_multicast, key, coll);
  if (got == 1) { if (how) *how = "cuMem* allocations, peer-mapped, bound to one multicast object (NVLS)"; return; }
  if (got == 2) { if (how) *how = "cuMem* allocations, peer-mapped (no multicast on this system)"; return; }
  alloc_ipc(r, device, nranks, rank, bytes, coll);
  if (how) *how = "cudaMalloc + CUDA IPC handles, peer-mapped (no multicast)";
}

void symm_free(SymmRegion& r) {
  if (r.vmm) {
    unmap_all(r);
  } else {
    for (int p = 0; p < r.nranks; p++)
      if (r.peer[p] && p != r.rank) cudaIpcCloseMemHandle(r.peer[p]);
    if (r.local) cudaFree(r.local);
  }
  r = SymmRegion{};
}

}  // namespace gadcu
