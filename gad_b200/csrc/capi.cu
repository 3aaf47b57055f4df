// capi.cu — the extern "C" surface declared in include/gadcu.h (algebra, sweep and Check parts).
#include "algebra.h"
#include "batched.h"

using namespace gadcu;

namespace {

Value in_value(const gadcu_value* v) {
  if (!v) fail(GADCU_ERR_INVALID, "null value");
  return Value(ArrayRef(v->data, true), Id{v->arena, v->index});
}
void out_value(gadcu_value* out, Value v) {
  out->arena = v.id.arena;
  out->index = v.id.index;
  out->data = v.data.detach();
}
Dim4 in_dims(const gadcu_dim4* d) {
  if (!d) fail(GADCU_ERR_INVALID, "null dims");
  return Dim4(*d);
}

}  // namespace

extern "C" {

// ---- Check --------------------------------------------------------------------------------------
gadcu_status gadcu_check_equal(const gadcu_dim4* const* dims, size_t n, gadcu_dim4* out) {
  return guard([&] {
    std::vector<Dim4> ds;
    for (size_t i = 0; i < n; i++) ds.push_back(Dim4(*dims[i]));
    std::vector<const Dim4*> ps;
    for (auto& d : ds) ps.push_back(&d);
    Dim4 r = check::equal("check_equal", ps);
    if (out) *out = r.c();
  });
}
gadcu_status gadcu_check_select_argmax(const gadcu_dim4* v0, const gadcu_dim4* v1, const gadcu_dim4* r0, const gadcu_dim4* r1,
                                       gadcu_dim4* out) {
  return guard([&] {
    Dim4 a(*v0), b(*v1), c, d;
    if (r0) c = Dim4(*r0);
    if (r1) d = Dim4(*r1);
    Dim4 r = check::select_argmax(a, b, r0 ? &c : nullptr, r1 ? &d : nullptr);
    if (out) *out = r.c();
  });
}
gadcu_status gadcu_check_flat(const gadcu_dim4* v, gadcu_dim4* out) {
  return guard([&] { *out = check::flat(Dim4(*v)).c(); });
}
gadcu_status gadcu_check_moddims(const gadcu_dim4* v, const gadcu_dim4* d, gadcu_dim4* out) {
  return guard([&] { *out = check::moddims(Dim4(*v), Dim4(*d)).c(); });
}
gadcu_status gadcu_check_tile_as(const gadcu_dim4* v, const gadcu_dim4* r, gadcu_dim4* out) {
  return guard([&] { *out = check::tile_as(Dim4(*v), Dim4(*r)).c(); });
}
gadcu_status gadcu_check_sum_as(const gadcu_dim4* v, const gadcu_dim4* r, gadcu_dim4* out) {
  return guard([&] { *out = check::sum_as(Dim4(*v), Dim4(*r)).c(); });
}
gadcu_status gadcu_check_as_scalar(const gadcu_dim4* v) {
  return guard([&] { check::as_scalar(Dim4(*v)); });
}
gadcu_status gadcu_check_dot(const gadcu_dim4* a, const gadcu_dim4* b) {
  return guard([&] { check::dot(Dim4(*a), Dim4(*b)); });
}
gadcu_status gadcu_check_reduced(const gadcu_dim4* v, const gadcu_dim4* r, int keeps_shape, gadcu_dim4* out) {
  return guard([&] {
    Dim4 rd = check::reduced("reduced", Dim4(*v), Dim4(*r));
    *out = keeps_shape ? *v : rd.c();
  });
}
gadcu_status gadcu_check_transpose(const gadcu_dim4* v, gadcu_dim4* out) {
  return guard([&] { *out = check::transpose(Dim4(*v)).c(); });
}
gadcu_status gadcu_check_matmul(const gadcu_dim4* a, const gadcu_dim4* b, gadcu_matprop p1, gadcu_matprop p2, gadcu_dim4* out) {
  return guard([&] {
    *out = check::matmul(Dim4(*a), Dim4(*b), MatProp{p1.transposed != 0, p1.conjugated != 0},
                         MatProp{p2.transposed != 0, p2.conjugated != 0}).c();
  });
}

// ---- algebras -----------------------------------------------------------------------------------
gadcu_status gadcu_algebra_create(gadcu_ctx* ctx, gadcu_algebra_kind kind, gadcu_algebra** out) {
  return guard([&] {
    switch (kind) {
      case GADCU_EVAL: *out = new EvalAlgebra(ctx); break;
      case GADCU_GRAPH1: *out = new GraphAlgebra(ctx, false); break;
      case GADCU_GRAPHN: *out = new GraphAlgebra(ctx, true); break;
      default: fail(GADCU_ERR_INVALID, "algebra_create: use gadcu_batched_create for batched tapes");
    }
  });
}
gadcu_status gadcu_algebra_clone(const gadcu_algebra* alg, gadcu_algebra** out) {
  return guard([&] { *out = alg->clone(); });
}
gadcu_status gadcu_algebra_destroy(gadcu_algebra* alg) {
  return guard([&] { delete alg; });
}
gadcu_algebra_kind gadcu_algebra_get_kind(const gadcu_algebra* alg) { return alg->kind(); }
size_t gadcu_algebra_num_nodes(const gadcu_algebra* alg) { return alg->num_nodes(); }

// The algebra takes a CLONE of the caller's handle (core.rs:162-170 `variable(data: D)` takes the array by value; gad
// programs pass `w.clone()`): what the tape and the returned Value hold is their own handle on the same buffer, so a
// later `+=` on the caller's handle (WeightOps::add_assign, copy-on-write) cannot change what the tape recorded.
gadcu_status gadcu_variable(gadcu_algebra* g, gadcu_array* data, gadcu_value* out) {
  return guard([&] { out_value(out, g->variable(clone_handle(ArrayRef(data, true)))); });
}
gadcu_status gadcu_constant(gadcu_algebra* g, gadcu_array* data, gadcu_value* out) {
  return guard([&] { out_value(out, g->constant(clone_handle(ArrayRef(data, true)))); });
}
gadcu_status gadcu_add_all(gadcu_algebra* g, const gadcu_value* const* vs, size_t n, gadcu_value* out) {
  return guard([&] {
    std::vector<Value> vals;
    for (size_t i = 0; i < n; i++) vals.push_back(in_value(vs[i]));
    std::vector<const Value*> ps;
    for (auto& v : vals) ps.push_back(&v);
    out_value(out, g->add_all(ps));
  });
}

#define GADCU_C_UNARY(NAME)                                                                   \
  gadcu_status gadcu_##NAME(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out) {        \
    return guard([&] { out_value(out, g->NAME(in_value(v))); });                               \
  }
#define GADCU_C_BINARY(NAME)                                                                                       \
  gadcu_status gadcu_##NAME(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out) {        \
    return guard([&] { out_value(out, g->NAME(in_value(a), in_value(b))); });                                        \
  }
#define GADCU_C_CONST(NAME)                                                                                         \
  gadcu_status gadcu_##NAME(gadcu_algebra* g, const gadcu_value* v, double c, int c_is_int, gadcu_value* out) {      \
    return guard([&] { out_value(out, g->NAME(in_value(v), c, c_is_int != 0)); });                                   \
  }
#define GADCU_C_DIMS(NAME)                                                                                          \
  gadcu_status gadcu_##NAME(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* d, gadcu_value* out) {         \
    return guard([&] { out_value(out, g->NAME(in_value(v), in_dims(d))); });                                         \
  }

GADCU_C_BINARY(add)
GADCU_C_BINARY(sub)
GADCU_C_BINARY(mul)
GADCU_C_UNARY(zeros)
GADCU_C_UNARY(ones)
GADCU_C_UNARY(neg)
GADCU_C_CONST(setc)
GADCU_C_CONST(addc)
GADCU_C_CONST(mulc)
GADCU_C_CONST(powc)
GADCU_C_UNARY(exp)
GADCU_C_UNARY(log)
GADCU_C_UNARY(log1p)
GADCU_C_UNARY(sin)
GADCU_C_UNARY(cos)
GADCU_C_UNARY(tanh)
GADCU_C_UNARY(sigmoid)
GADCU_C_UNARY(reciprocal)
GADCU_C_UNARY(sqrt)
GADCU_C_BINARY(div)
GADCU_C_BINARY(pow)
GADCU_C_BINARY(min)
GADCU_C_BINARY(max)
GADCU_C_UNARY(abs)
GADCU_C_UNARY(sign)
GADCU_C_UNARY(relu)
GADCU_C_UNARY(flat)
GADCU_C_DIMS(moddims)
GADCU_C_DIMS(tile_as)
GADCU_C_DIMS(sum_as)
GADCU_C_DIMS(constant_as)
GADCU_C_UNARY(as_scalar)
GADCU_C_BINARY(scale)
GADCU_C_BINARY(dot)
GADCU_C_UNARY(norm2)
GADCU_C_BINARY(matmul_nn)
GADCU_C_DIMS(max_as)
GADCU_C_DIMS(argmax_as)
GADCU_C_DIMS(softmax_as)

gadcu_status gadcu_select_argmax(gadcu_algebra* g, const gadcu_value* v0, const gadcu_value* v1, const gadcu_value* r0,
                                 const gadcu_value* r1, gadcu_value* out) {
  return guard([&] {
    Value a = in_value(v0), b = in_value(v1), c, d;
    if (r0) c = in_value(r0);
    if (r1) d = in_value(r1);
    out_value(out, g->select_argmax(a, b, r0 ? &c : nullptr, r1 ? &d : nullptr));
  });
}
gadcu_status gadcu_matmul(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_matprop p1, gadcu_matprop p2,
                          gadcu_value* out) {
  return guard([&] {
    out_value(out, g->matmul(in_value(a), in_value(b), MatProp{p1.transposed != 0, p1.conjugated != 0},
                             MatProp{p2.transposed != 0, p2.conjugated != 0}));
  });
}
gadcu_status gadcu_transpose(gadcu_algebra* g, const gadcu_value* v, int conjugate, gadcu_value* out) {
  return guard([&] { out_value(out, g->transpose(in_value(v), conjugate != 0)); });
}

// ---- user-defined operators ---------------------------------------------------------------------
gadcu_status gadcu_make_node(gadcu_algebra* g, gadcu_array* data, const gadcu_value* const* inputs, size_t n, gadcu_vjp_fn vjp,
                             void* user, void (*drop_user)(void*), gadcu_value* out) {
  return guard([&] {
    auto custom = std::make_shared<CustomVjp>();
    custom->fn = vjp;
    custom->user = user;
    custom->drop = drop_user;
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph) {  // Eval: a node-less value
      out_value(out, Value(ArrayRef(data, true)));
      return;
    }
    std::vector<Id> ids;
    for (size_t i = 0; i < n; i++) ids.push_back(Id{inputs[i]->arena, inputs[i]->index});
    out_value(out, graph->make_custom(ArrayRef(data, true), std::move(ids), std::move(custom)));
  });
}
gadcu_status gadcu_store_add_gradient(gadcu_store* store, gadcu_algebra* grad_algebra, uint32_t arena, uint32_t index,
                                      const gadcu_value* value) {
  return guard([&] { store->add_gradient(*grad_algebra, Id{arena, index}, in_value(value)); });
}

// ---- sweep --------------------------------------------------------------------------------------
gadcu_status gadcu_evaluate_gradients(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed, int once,
                                      gadcu_store** out) {
  return guard([&] {
    if (index == 0) fail(GADCU_ERR_MISSING_ID, "Trying to obtain `id` of a constant value.");
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph) fail(GADCU_ERR_INVALID, "evaluate_gradients needs a Graph1 or batched algebra");
    if (graph->kind() == GADCU_BATCHED1)
      *out = batched_evaluate_gradients(*graph, Id{arena, index}, ArrayRef(seed, true), once != 0).release();
    else
      *out = graph->evaluate_gradients(Id{arena, index}, ArrayRef(seed, true), once != 0).release();
    g->ctx->segment_flush();  // auto replay: a finished sweep is where a step's recorded launches run (as one graph)
  });
}
gadcu_status gadcu_compute_gradients(gadcu_algebra* g, uint32_t arena, uint32_t index, const gadcu_value* seed, gadcu_store** out) {
  return guard([&] {
    if (index == 0) fail(GADCU_ERR_MISSING_ID, "Trying to obtain `id` of a constant value.");
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph) fail(GADCU_ERR_INVALID, "compute_gradients needs a GraphN algebra");
    *out = graph->compute_gradients(Id{arena, index}, in_value(seed)).release();
    g->ctx->segment_flush();
  });
}
gadcu_status gadcu_store_get(const gadcu_store* s, uint32_t arena, uint32_t index, gadcu_value* out, int* found) {
  return guard([&] {
    const Value* v = s->get(Id{arena, index});
    *found = v ? 1 : 0;
    if (v) out_value(out, *v);
    else { out->data = nullptr; out->arena = 0; out->index = 0; }
  });
}
size_t gadcu_store_ids(const gadcu_store* s, uint32_t* out, size_t cap) {
  std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
  size_t k = 0;
  for (const Id& id : s->ids()) {
    if (k < cap) out[k] = id.index;
    k++;
  }
  return k;
}
gadcu_status gadcu_store_destroy(gadcu_store* s) {
  return guard([&] { delete s; });
}

}  // extern "C"
