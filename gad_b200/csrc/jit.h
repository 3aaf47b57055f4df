// jit.h — NVRTC-compiled lane programs (see jit.cu).
#pragma once
#include <string>

#include "core.h"

namespace gadcu {
namespace jit {

struct Kernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
  size_t local_bytes = 0;  // per-thread local memory: register spills (and any indexed local array)
  mutable size_t smem_allowed = 48 * 1024;  // dynamic shared memory the kernel has been opted in to
};

bool available();  // NVRTC could be loaded (GADCU_JIT=off disables it)
const Kernel* find(const std::string& key);
// Compiles `source` (one extern "C" __global__ function `entry`) for sm_100a with --fmad=false and caches it under `key`.
const Kernel* compile(const std::string& key, const std::string& source, const char* entry);
void launch(gadcu_ctx* ctx, const Kernel* k, unsigned grid, unsigned block, void** args, size_t smem_bytes = 0);

}  // namespace jit
}  // namespace gadcu
