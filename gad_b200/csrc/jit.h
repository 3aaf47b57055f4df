// jit.h — NVRTC-compiled lane programs (see jit.cu).
#pragma once
#include <string>

#include "core.h"

namespace gadcu {
namespace jit {

struct Kernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
};

bool available();  // NVRTC could be loaded (GADCU_JIT=off disables it)
const Kernel* find(const std::string& key);
// Compiles `source` (one extern "C" __global__ function `entry`) for sm_100a with --fmad=false and caches it under `key`.
const Kernel* compile(const std::string& key, const std::string& source, const char* entry);
void launch(gadcu_ctx* ctx, const Kernel* k, unsigned grid, unsigned block, void** args);

}  // namespace jit
}  // namespace gadcu
