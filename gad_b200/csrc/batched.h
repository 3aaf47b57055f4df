// batched.h — batched scalar tapes: thousands of independent gad scalar tapes, one per lane, in
// structure-of-arrays layout (north-star subsystem 4).
#pragma once
#include "algebra.h"

namespace gadcu {

// A GraphAlgebra whose eval algebra is the symbolic per-lane algebra (kind GADCU_BATCHED1).
std::unique_ptr<GraphAlgebra> make_batched(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t lanes);
// Graph1::evaluate_gradients over lanes: symbolic sweep, then one kernel computes the gradients of
// all variables. Other store entries stay symbolic and are computed when read.
std::unique_ptr<Store> batched_evaluate_gradients(GraphAlgebra& g, Id id, ArrayRef seed, bool once);
std::vector<ArrayRef> batched_rerun(Store& store, const std::vector<ArrayRef>& columns);


// ---- deferred elementwise regions of array tapes (batched.cu) ---------------------------------------
bool lazy_enabled(gadcu_ctx* ctx, uint64_t n, bool is_scalar);
ArrayRef lazy_elementwise(gadcu_ctx* ctx, k::EwOp op, const ArrayRef& acc, const std::vector<const ArrayRef*>& ins, double c, double c2,
                          gadcu_dtype dt, const Dim4& dims, bool is_scalar);
// a matmul recorded as a leaf of the region of its result (launched when something needs it, with the elementwise tail
// behind it in the GEMM's epilogue where that is one of the known ones: GemmEpilogue)
ArrayRef lazy_matmul(gadcu_ctx* ctx, const ArrayRef& a, const ArrayRef& b, MatProp p1, MatProp p2, const Dim4& rd, bool fusable);
// sum / max of a still-deferred value over contiguous segments ([reduce, outer] view), fused into the kernel that computes
// it, bit-identical to materialise + k::launch_reduce; null when not applicable (the caller takes the ordinary path)
ArrayRef lazy_reduce(gadcu_ctx* ctx, k::RedOp op, const ArrayRef& v, uint64_t reduce, uint64_t outer, const Dim4& out_dims);
void lazy_before_write(gadcu_ctx* ctx, const void* buffer);
void lazy_flush(const std::vector<ArrayRef>& values);

}  // namespace gadcu
