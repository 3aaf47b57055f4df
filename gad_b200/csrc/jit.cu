// jit.cu — runtime specialisation of lane programs through NVRTC.
//
// The reference's array backend (ArrayFire) JIT-fuses chains of elementwise operations; the batched scalar
// tapes and the deferred elementwise regions of array tapes (lanes.cu) are the same thing here: a straight-line
// SSA program over "one value per lane". The program text is generated from hand-written per-op snippets and a
// hand-written kernel skeleton (128-bit vector I/O, grid-stride, all loads first), compiled once per distinct
// program structure for sm_100a with --fmad=false (one rounding per tape operation, exactly like the eager
// kernels of ew.cu, so fused and unfused sweeps still compare with ==), and cached by structure hash.
// NVRTC is dlopen()ed from the CUDA toolkit the library was built with so libdevice matches nvcc's.
#include "jit.h"

#include <dlfcn.h>

#include <cstdio>

#include <mutex>
#include <unordered_map>

namespace gadcu {
namespace jit {
namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char* e = getenv("GADCU_JIT"))
      if (std::string(e) == "off") return;
    // the toolkit's own copy first: a different NVRTC (e.g. a pip wheel earlier on the search path) may carry a
    // different libdevice, and the compiled math functions must match the nvcc-built eager kernels bit for bit
    const char* names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12", "libnvrtc.so"};
    for (auto name : names) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
#define GADCU_SYM(field, sym) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, sym))
    GADCU_SYM(CreateProgram, "nvrtcCreateProgram");
    GADCU_SYM(CompileProgram, "nvrtcCompileProgram");
    GADCU_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    GADCU_SYM(GetCUBIN, "nvrtcGetCUBIN");
    GADCU_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    GADCU_SYM(GetProgramLog, "nvrtcGetProgramLog");
    GADCU_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef GADCU_SYM
    n.ok = n.CreateProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN && n.GetProgramLogSize && n.GetProgramLog &&
           n.DestroyProgram;
  });
  return n;
}

struct Cache {
  std::mutex mu;
  std::unordered_map<std::string, Kernel> kernels;
};
Cache& cache() {
  static Cache c;
  return c;
}

}  // namespace

bool available() { return nvrtc().ok; }

const Kernel* find(const std::string& key) {
  Cache& c = cache();
  std::lock_guard<std::mutex> lk(c.mu);
  auto it = c.kernels.find(key);
  return it == c.kernels.end() ? nullptr : &it->second;
}

const Kernel* compile(const std::string& key, const std::string& source, const char* entry) {
  Nvrtc& n = nvrtc();
  if (!n.ok) fail(GADCU_ERR_UNSUPPORTED, "NVRTC is not available");
  nvrtcProgram prog = nullptr;
  if (n.CreateProgram(&prog, source.c_str(), "gadcu_lane_program.cu", 0, nullptr, nullptr) != 0)
    fail(GADCU_ERR_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-std=c++17", "-lineinfo", "-default-device"};
  const int rc = n.CompileProgram(prog, int(sizeof(opts) / sizeof(opts[0])), opts);
  if (rc != 0) {
    size_t log_size = 0;
    n.GetProgramLogSize(prog, &log_size);
    std::string log(log_size, '\0');
    if (log_size) n.GetProgramLog(prog, log.data());
    n.DestroyProgram(&prog);
    fail(GADCU_ERR_CUDA, "NVRTC compilation of a lane program failed:\n" + log);
  }
  size_t size = 0;
  n.GetCUBINSize(prog, &size);
  std::vector<char> cubin(size);
  n.GetCUBIN(prog, cubin.data());
  n.DestroyProgram(&prog);

  Kernel k;
  GADCU_CUDA(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  GADCU_CUDA(cudaLibraryGetKernel(&k.fn, k.lib, entry));
  cudaFuncAttributes attr;
  if (cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(k.fn)) == cudaSuccess) {
    k.local_bytes = attr.localSizeBytes;
    if (const char* dump = getenv("GADCU_JIT_DUMP")) {
      if (FILE* f = fopen(dump, "a")) {
        fprintf(f, "// compiled: %d registers, %zu bytes local (spills), %zu bytes cubin\n", attr.numRegs, attr.localSizeBytes, cubin.size());
        fclose(f);
      }
    }
  }
  Cache& c = cache();
  std::lock_guard<std::mutex> lk(c.mu);
  auto res = c.kernels.emplace(key, k);
  return &res.first->second;
}

void launch(gadcu_ctx* ctx, const Kernel* k, unsigned grid, unsigned block, void** args, size_t smem_bytes) {
  if (smem_bytes > k->smem_allowed) {
    GADCU_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(k->fn), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    k->smem_allowed = smem_bytes;
  }
  GADCU_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(k->fn), dim3(grid), dim3(block), args, smem_bytes, ctx->stream));
  ctx->count_launch();
}

}  // namespace jit
}  // namespace gadcu
