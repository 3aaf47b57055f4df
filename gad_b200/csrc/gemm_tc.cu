// gemm_tc.cu — tensor-core GEMM (tcgen05 + TMA, 3xTF32). Placeholder until the kernel lands.
#include "kernels.h"

namespace gadcu {
namespace k {

bool gemm_tc_supported(const GemmArgs&) { return false; }
void launch_gemm_tc(gadcu_ctx*, const GemmArgs&) { fail(GADCU_ERR_UNSUPPORTED, "tensor-core gemm not built"); }

}  // namespace k
}  // namespace gadcu
