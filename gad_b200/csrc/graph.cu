// graph.cu — the tape (gad/src/graph.rs), the gradient store (gad/src/store.rs) and every VJP body
// (SURVEY App. A), over device arrays.
//
// Each VJP exists in two forms that compute the same rounded values:
//   * generic: the reference's exact sequence of algebra calls, written against the abstract
//     algebra. With the graph itself as gradient algebra this RECORDS the backward pass (GraphN);
//     with Eval it is the reference's one-kernel-per-primitive sweep (ctx fusion off).
//   * fused  : (Graph1, fusion on) one kernel per contribution that evaluates the whole VJP per
//     element and folds in `current + new` of add_gradient, in place when the store owns the
//     buffer alone — no intermediate array ever reaches HBM.
#include "algebra.h"
#include "batched.h"

#include <algorithm>
#include <cstring>
#include <functional>

using namespace gadcu;

// ---- store --------------------------------------------------------------------------------------
void gadcu_store::add_gradient(Algebra& graph, Id id, const Value& value) {  // store.rs:44-55
  realize(id);
  auto it = values.find(id);
  if (it == values.end()) {
    values.emplace(id, value);  // first contribution is stored by clone, not added to zero
  } else {
    Value cur = std::move(it->second);
    it->second = graph.accumulate(std::move(cur), value);
  }
}

void gadcu_store::add_fused(EvalAlgebra& ev, Id id, k::EwOp op, std::initializer_list<const ArrayRef*> ins, double c, double c2) {
  realize(id);
  auto it = values.find(id);
  if (it == values.end()) {
    values.emplace(id, Value(ev.fused(op, ArrayRef(), ins, c, c2)));
  } else {
    ArrayRef cur = std::move(it->second.data);
    it->second.data = ev.fused(op, std::move(cur), ins, c, c2);
  }
}

namespace gadcu {

static std::atomic<uint32_t> g_arena_counter{1};
// nodes of the tape swept last: a program records the same tape every step, so the next one starts with room for it
// (four reallocations, each moving every node recorded so far, per 569-node tape otherwise)
static std::atomic<uint32_t> g_tape_nodes_hint{0};

GraphAlgebra::GraphAlgebra(gadcu_ctx* c, bool higher_order, std::shared_ptr<Algebra> eval)
    : Algebra(c), higher_order_(higher_order), arena_id_(g_arena_counter.fetch_add(1)),
      eval_(eval ? std::move(eval) : std::make_shared<EvalAlgebra>(c)) {
  if (higher_order)
    if (auto* ev = dynamic_cast<EvalAlgebra*>(eval_.get())) ev->fuse_tails = false;
  c->epoch.fetch_add(1, std::memory_order_relaxed);  // a new tape: TF32 splits cached by earlier steps expire
  const uint32_t hint = g_tape_nodes_hint.load(std::memory_order_relaxed);
  if (hint) nodes_.reserve(std::min<uint32_t>(hint, 1u << 16));
}

// graph.rs:94-101
Value GraphAlgebra::variable(ArrayRef data) {
  if (kind_override_ == GADCU_BATCHED1) variable_data_.push_back(data);  // gadcu_batched_rerun swaps these columns
  nodes_.emplace_back();
  if (data) nodes_.back().fwd_dims = data->dims;
  return Value(std::move(data), Id{arena_id_, uint32_t(nodes_.size())});
}

// graph.rs:127-165: no tracked input => constant (no node); else record dims and the update.
Value GraphAlgebra::make_node(ArrayRef data, Node node) {
  bool all_none = true;
  for (auto& i : node.inputs) all_none = all_none && !i.some();
  if (all_none) return Value(std::move(data));
  node.has_update = true;
  node.fwd_dims = data->dims;
  node.fwd_scalar = data->is_scalar;
  node.fwd_dtype = data->dtype;
  nodes_.push_back(std::move(node));
  return Value(std::move(data), Id{arena_id_, uint32_t(nodes_.size())});
}
Value GraphAlgebra::make_custom(ArrayRef data, std::vector<Id> inputs, std::shared_ptr<CustomVjp> vjp) {
  Node n;
  n.op = Op::CUSTOM;
  n.inputs = std::move(inputs);
  n.custom = std::move(vjp);
  return make_node(std::move(data), std::move(n));
}

namespace {
Node node1(Op op, const Value& v) {
  Node n;
  n.op = op;
  n.inputs = {v.id};
  return n;
}
Node node2(Op op, const Value& a, const Value& b) {
  Node n;
  n.op = op;
  n.inputs = {a.id, b.id};
  return n;
}
}  // namespace

// ---- recording (forward through Eval, then a node) ---------------------------------------------
Value GraphAlgebra::add(const Value& a, const Value& b) {  // core.rs:219-237
  Value r = eval_->add(a, b);
  return make_node(r.data, node2(Op::ADD, a, b));
}
Value GraphAlgebra::add_all(const std::vector<const Value*>& vs) {  // core.rs:239-256
  Value r = eval_->add_all(vs);
  Node n;
  n.op = Op::ADD_ALL;
  for (auto v : vs) n.inputs.push_back(v->id);
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::neg(const Value& v) {  // arith.rs:175-187
  Value r = eval_->neg(v);
  return make_node(r.data, node1(Op::NEG, v));
}
Value GraphAlgebra::sub(const Value& a, const Value& b) {  // arith.rs:189-206
  Value r = eval_->sub(a, b);
  return make_node(r.data, node2(Op::SUB, a, b));
}
Value GraphAlgebra::mul(const Value& a, const Value& b) {  // arith.rs:208-228
  Value r = eval_->mul(a, b);
  Node n = node2(Op::MUL, a, b);
  n.saved = {a, b};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::addc(const Value& v, double c, bool is_int) {  // const_arith.rs:139-150
  Value r = eval_->addc(v, c, is_int);
  return make_node(r.data, node1(Op::ADDC, v));
}
Value GraphAlgebra::mulc(const Value& v, double c, bool is_int) {  // const_arith.rs:152-164
  Value r = eval_->mulc(v, c, is_int);
  Node n = node1(Op::MULC, v);
  n.c = c;
  n.c_is_int = is_int;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::powc(const Value& v, double c, bool is_int) {  // const_arith.rs:166-181
  Value r = eval_->powc(v, c, is_int);
  Node n = node1(Op::POWC, v);
  n.c = c;
  n.c_is_int = is_int;
  n.saved = {v};
  return make_node(r.data, std::move(n));
}
#define GADCU_RECORD_UNARY(NAME, OP)               \
  Value GraphAlgebra::NAME(const Value& v) {       \
    Value r = eval_->NAME(v);                       \
    Node n = node1(Op::OP, v);                     \
    n.saved = {v};                                 \
    return make_node(r.data, std::move(n));        \
  }
GADCU_RECORD_UNARY(exp, EXP)                // analytic.rs:301-315
GADCU_RECORD_UNARY(log, LOG)                // :317-330
GADCU_RECORD_UNARY(log1p, LOG1P)            // :332-346
GADCU_RECORD_UNARY(sin, SIN)                // :348-362
GADCU_RECORD_UNARY(cos, COS)                // :364-379
GADCU_RECORD_UNARY(tanh, TANH)              // :381-398
GADCU_RECORD_UNARY(sigmoid, SIGMOID)        // :400-417
GADCU_RECORD_UNARY(reciprocal, RECIPROCAL)  // :419-435
GADCU_RECORD_UNARY(sqrt, SQRT)              // :437-453
#undef GADCU_RECORD_UNARY
Value GraphAlgebra::div(const Value& a, const Value& b) {  // analytic.rs:455-478
  Value r = eval_->div(a, b);
  Node n = node2(Op::DIV, a, b);
  n.saved = {a, b};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) {  // compare.rs:202-245
  Value r = eval_->select_argmax(v0, v1, r0, r1);
  Node n;
  n.op = Op::SELECT_ARGMAX;
  if (r0) n.inputs.push_back(r0->id);  // node inputs are r0, r1 only: v0, v1 get no gradient
  if (r1) n.inputs.push_back(r1->id);
  n.saved = {v0, v1, Value(ArrayRef(), r0 ? r0->id : Id{}), Value(ArrayRef(), r1 ? r1->id : Id{})};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::flat(const Value& v) {  // array.rs:214-227
  Value r = eval_->flat(v);
  Node n = node1(Op::MODDIMS, v);
  n.d = v.data->dims;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::moddims(const Value& v, const Dim4& d) {  // array.rs:229-243
  Value r = eval_->moddims(v, d);
  Node n = node1(Op::MODDIMS, v);
  n.d = v.data->dims;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::tile_as(const Value& v, const Dim4& d) {  // array.rs:245-259
  Value r = eval_->tile_as(v, d);
  Node n = node1(Op::TILE_AS, v);
  n.d = v.data->dims;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::sum_as(const Value& v, const Dim4& d) {  // array.rs:261-275
  Value r = eval_->sum_as(v, d);
  Node n = node1(Op::SUM_AS, v);
  n.d = v.data->dims;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::constant_as(const Value& s, const Dim4& d) {  // array.rs:277-291
  Value r = eval_->constant_as(s, d);
  return make_node(r.data, node1(Op::CONSTANT_AS, s));
}
Value GraphAlgebra::as_scalar(const Value& v) {  // array.rs:293-307
  Value r = eval_->as_scalar(v);
  Node n = node1(Op::AS_SCALAR, v);
  n.d = v.data->dims;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::scale(const Value& lambda, const Value& v) {  // array.rs:309-329
  Value r = eval_->scale(lambda, v);
  Node n = node2(Op::SCALE, lambda, v);
  n.saved = {lambda, v};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::dot(const Value& a, const Value& b) {  // array.rs:331-351
  Value r = eval_->dot(a, b);
  Node n = node2(Op::DOT, a, b);
  n.saved = {a, b};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::matmul(const Value& a, const Value& b, MatProp p1, MatProp p2) {  // matrix.rs:153-179
  Value r = eval_->matmul(a, b, p1, p2);
  Node n = node2(Op::MATMUL, a, b);
  n.saved = {a, b};
  n.p1 = p1;
  n.p2 = p2;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::transpose(const Value& v, bool conjugate) {  // matrix.rs:181-194
  Value r = eval_->transpose(v, conjugate);
  Node n = node1(Op::TRANSPOSE, v);
  n.flag = conjugate;
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::max_as(const Value& v, const Dim4& d) {  // array_compare.rs:123-139
  Value r = eval_->max_as(v, d);
  Node n = node1(Op::MAX_AS, v);
  n.d = d;
  n.saved = {v};
  return make_node(r.data, std::move(n));
}
Value GraphAlgebra::softmax_as(const Value& v, const Dim4& d) {  // array_compare.rs:146-168
  Value r = eval_->softmax_as(v, d);
  Node n = node1(Op::SOFTMAX_AS, v);
  n.d = d;
  n.saved = {v};
  return make_node(r.data, std::move(n));
}

// The product about to be computed is the whole gradient of variable `id` (no earlier contribution, and the node
// being differentiated is the last one that feeds it): offer it to the sink.
ArrayRef GraphAlgebra::try_gemm_sink(Id id, const ArrayRef& lhs, const ArrayRef& rhs, MatProp pl, MatProp pr) {
  if (!gemm_sink_ || id.arena != arena_id_ || id.index >= waiting_.size() || waiting_[id.index] != 1 || !is_variable(id)) return ArrayRef();
  // is any other large variable still waiting for contributions? if not, this product ends the sweep's heavy work
  bool last = true;
  for (size_t v = 1; v < waiting_.size() && last; v++)
    if (v != id.index && waiting_[v] > 0 && !served_[v] && nodes_[v - 1].fwd_dims.elements() >= (1u << 20)) last = false;
  ArrayRef out = gemm_sink_(id, lhs, rhs, pl, pr, last);
  if (out) served_[id.index] = 1;
  return out;
}

// ---- VJPs -----------------------------------------------------------------------
void GraphAlgebra::vjp(const Node& n, Algebra& ga, Store& store, const Value& g) {
  // Fused kernels apply when gradients are plain data (Graph1) and fusion is on.
  EvalAlgebra* ev = (ga.is_eval() && ctx->fuse) ? static_cast<EvalAlgebra*>(&ga) : nullptr;
  auto in = [&](size_t i) { return n.inputs[i]; };
  switch (n.op) {
    case Op::VARIABLE:
      return;
    case Op::ADD:  // core.rs:221-237
      if (in(0).some()) store.add_gradient(ga, in(0), g);
      if (in(1).some()) store.add_gradient(ga, in(1), g);
      return;
    case Op::ADD_ALL:  // core.rs:247-251
      for (auto& id : n.inputs)
        if (id.some()) store.add_gradient(ga, id, g);
      return;
    case Op::NEG:  // arith.rs:179-185
      if (in(0).some()) {
        if (ev) { store.add_fused(*ev, in(0), k::EwOp::VJP_NEG, {&g.data}); return; }
        Value x = ga.neg(g);
        store.add_gradient(ga, in(0), x);
      }
      return;
    case Op::SUB:  // arith.rs:194-203
      if (in(0).some()) store.add_gradient(ga, in(0), g);
      if (in(1).some()) {
        if (ev) { store.add_fused(*ev, in(1), k::EwOp::VJP_NEG, {&g.data}); return; }
        Value x = ga.neg(g);
        store.add_gradient(ga, in(1), x);
      }
      return;
    case Op::MUL: {  // arith.rs:213-225
      const Value &v0 = n.saved[0], &v1 = n.saved[1];
      if (v0.id.some()) {
        if (ev) {
          store.add_fused(*ev, v0.id, k::EwOp::VJP_MUL, {&g.data, &v1.data});
        } else {
          Value c1 = ga.link(v1);
          Value grad = ga.mul(g, c1);
          store.add_gradient(ga, v0.id, grad);
        }
      }
      if (v1.id.some()) {
        if (ev) {
          // mul(c0, g): the product commutes bit-for-bit, so the same kernel serves both sides
          store.add_fused(*ev, v1.id, k::EwOp::VJP_MUL, {&g.data, &v0.data});
        } else {
          Value c0 = ga.link(v0);
          Value grad = ga.mul(c0, g);
          store.add_gradient(ga, v1.id, grad);
        }
      }
      return;
    }
    case Op::ADDC:  // const_arith.rs:143-147
      if (in(0).some()) store.add_gradient(ga, in(0), g);
      return;
    case Op::MULC:  // const_arith.rs:156-161
      if (in(0).some()) {
        if (ev) { store.add_fused(*ev, in(0), k::EwOp::VJP_MULC, {&g.data}, n.c_is_int ? double(int16_t(n.c)) : n.c); return; }
        Value grad = ga.mulc(g, n.c, n.c_is_int);
        store.add_gradient(ga, in(0), grad);
      }
      return;
    case Op::POWC: {  // const_arith.rs:170-178
      const Value& v = n.saved[0];
      if (v.id.some()) {
        // c - C::one() is computed in the constant's own type C (i16 or T)
        const double cm1 = n.c_is_int ? double(int16_t(int16_t(n.c) - 1)) : (g.data->dtype == GADCU_F32 ? double(float(n.c) - 1.0f) : n.c - 1.0);
        if (ev) {
          store.add_fused(*ev, v.id, k::EwOp::VJP_POWC, {&g.data, &v.data}, n.c_is_int ? double(int16_t(n.c)) : n.c, cm1);
          return;
        }
        Value lv = ga.link(v);
        Value e = ga.powc(lv, cm1, n.c_is_int);
        Value f = ga.mulc(e, n.c, n.c_is_int);
        Value grad = ga.mul(f, g);
        store.add_gradient(ga, v.id, grad);
      }
      return;
    }
    case Op::EXP: {  // analytic.rs:305-312
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_EXP, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value kk = ga.exp(lv);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::LOG: {  // analytic.rs:321-327
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_LOG, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value grad = ga.div(g, lv);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::LOG1P: {  // analytic.rs:336-343
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_LOG1P, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value v1p = ga.addc(lv, 1, true);
      Value grad = ga.div(g, v1p);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::SIN: {  // analytic.rs:352-359
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_SIN, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value kk = ga.cos(lv);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::COS: {  // analytic.rs:368-376
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_COS, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value c = ga.sin(lv);
      Value kk = ga.neg(c);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::TANH: {  // analytic.rs:385-395
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_TANH, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value t = ga.tanh(lv);
      Value c = ga.mul(t, t);
      Value c2 = ga.neg(c);
      Value kk = ga.addc(c2, 1, true);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::SIGMOID: {  // analytic.rs:404-414
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_SIGMOID, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value c = ga.sigmoid(lv);
      Value d = ga.neg(c);
      Value d2 = ga.addc(d, 1, true);
      Value kk = ga.mul(c, d2);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::RECIPROCAL: {  // analytic.rs:423-432
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_RECIPROCAL, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value c = ga.mul(lv, lv);
      Value c2 = ga.neg(c);
      Value kk = ga.reciprocal(c2);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::SQRT: {  // analytic.rs:441-450
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      if (ev) { store.add_fused(*ev, v.id, k::EwOp::VJP_SQRT, {&g.data, &v.data}); return; }
      Value lv = ga.link(v);
      Value c = ga.sqrt(lv);
      Value c2 = ga.mulc(c, 2, true);
      Value kk = ga.reciprocal(c2);
      Value grad = ga.mul(g, kk);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::DIV: {  // analytic.rs:460-475
      const Value &v0 = n.saved[0], &v1 = n.saved[1];
      if (ev) {
        if (v0.id.some()) store.add_fused(*ev, v0.id, k::EwOp::VJP_DIV0, {&g.data, &v1.data});
        if (v1.id.some()) store.add_fused(*ev, v1.id, k::EwOp::VJP_DIV1, {&g.data, &v1.data, &v0.data});
        return;
      }
      Value c1 = ga.link(v1);
      Value r1 = ga.reciprocal(c1);
      Value g0 = ga.mul(g, r1);
      if (v0.id.some()) store.add_gradient(ga, v0.id, g0);
      if (v1.id.some()) {
        Value c0 = ga.link(v0);
        Value c = ga.mul(g0, r1);
        Value c2 = ga.mul(c, c0);
        Value g1 = ga.neg(c2);
        store.add_gradient(ga, v1.id, g1);
      }
      return;
    }
    case Op::SELECT_ARGMAX: {  // compare.rs:229-242
      const Value &v0 = n.saved[0], &v1 = n.saved[1];
      const Id id0 = n.saved[2].id, id1 = n.saved[3].id;
      if (ev) {
        if (id0.some()) store.add_fused(*ev, id0, k::EwOp::VJP_SELECT0, {&g.data, &v0.data, &v1.data});
        if (id1.some()) store.add_fused(*ev, id1, k::EwOp::VJP_SELECT1, {&g.data, &v0.data, &v1.data});
        return;
      }
      Value c0 = ga.link(v0), c1 = ga.link(v1);
      if (id0.some()) {
        Value grad = ga.select_argmax(c0, c1, &g, nullptr);
        store.add_gradient(ga, id0, grad);
      }
      if (id1.some()) {
        Value grad = ga.select_argmax(c0, c1, nullptr, &g);
        store.add_gradient(ga, id1, grad);
      }
      return;
    }
    case Op::MODDIMS:  // array.rs:219-224, 234-239
      if (in(0).some()) {
        Value x = ga.moddims(g, n.d);
        store.add_gradient(ga, in(0), x);
      }
      return;
    case Op::TILE_AS:  // array.rs:250-255
      if (in(0).some()) {
        if (ev) {
          // sum_as(g, vdims) with the accumulation folded into the final reduction pass
          auto it = store.values.find(in(0));
          ArrayRef cur;
          if (it != store.values.end()) cur = std::move(it->second.data);
          ArrayRef r = ev->sum_as_acc(std::move(cur), g.data, n.d);
          store.values[in(0)] = Value(r);
          return;
        }
        Value x = ga.sum_as(g, n.d);
        store.add_gradient(ga, in(0), x);
      }
      return;
    case Op::SUM_AS:  // array.rs:266-271
      if (in(0).some()) {
        Value x = ga.tile_as(g, n.d);
        store.add_gradient(ga, in(0), x);
      }
      return;
    case Op::CONSTANT_AS:  // array.rs:282-287
      if (in(0).some()) {
        Value x = ga.sum_as(g, Dim4());
        Value y = ga.as_scalar(x);
        store.add_gradient(ga, in(0), y);
      }
      return;
    case Op::AS_SCALAR:  // array.rs:298-303
      if (in(0).some()) {
        Value x = ga.constant_as(g, n.d);
        store.add_gradient(ga, in(0), x);
      }
      return;
    case Op::SCALE: {  // array.rs:315-326
      const Value &v1 = n.saved[0], &v2 = n.saved[1];
      if (v1.id.some()) {
        Value c2 = ga.link(v2);
        Value grad = ga.dot(g, c2);
        store.add_gradient(ga, v1.id, grad);
      }
      if (v2.id.some()) {
        Value c1 = ga.link(v1);
        Value grad = ga.scale(c1, g);
        store.add_gradient(ga, v2.id, grad);
      }
      return;
    }
    case Op::DOT: {  // array.rs:337-348
      const Value &v1 = n.saved[0], &v2 = n.saved[1];
      if (v1.id.some()) {
        Value c2 = ga.link(v2);
        if (ev) {
          ArrayRef lam = view_array(g.data, c2.data->dims, Dim4(), c2.data->is_scalar);
          store.add_fused(*ev, v1.id, k::EwOp::SCALE, {&c2.data, &lam});
        } else {
          Value grad = ga.scale(g, c2);
          store.add_gradient(ga, v1.id, grad);
        }
      }
      if (v2.id.some()) {
        Value c1 = ga.link(v1);
        if (ev) {
          ArrayRef lam = view_array(g.data, c1.data->dims, Dim4(), c1.data->is_scalar);
          store.add_fused(*ev, v2.id, k::EwOp::SCALE, {&c1.data, &lam});
        } else {
          Value grad = ga.scale(g, c1);
          store.add_gradient(ga, v2.id, grad);
        }
      }
      return;
    }
    case Op::MATMUL: {  // matrix.rs:164-176
      // Reference: dA = matmul(G, B, p1, p2^T), dB = matmul(A, G, p1^T, p2). These are the correct
      // adjoints exactly when the differentiated operand is itself not transposed;
      // otherwise the reference returns a Dimensions error (or, for square matrices, the transpose
      // of the adjoint). Unless ctx->strict_reference is set, a transposed operand gets the correct
      // adjoint instead — dA = op2(B) G^T, dB = G^T op1(A) — which is what lets GraphN differentiate
      // through the recorded backward GEMMs (BASELINE config C5). Identical where the reference is valid.
      const Value &v1 = n.saved[0], &v2 = n.saved[1];
      const bool fix1 = n.p1.transposed && !ctx->strict_reference;
      const bool fix2 = n.p2.transposed && !ctx->strict_reference;
      const MatProp T{true, false};
      // matmul(W, W): both products feed the SAME variable. Neither is its whole gradient, so neither may be handed to
      // the data-parallel sink (the second would be accumulated, unreduced, into a buffer the exchange is still writing)
      const bool same_operand = v1.id.some() && v1.id == v2.id;
      if (v1.id.some()) {
        // operands and props of the product that yields dA
        const Value& lhs = fix1 ? v2 : g;
        const Value& rhs = fix1 ? g : v2;
        const MatProp pl = fix1 ? n.p2 : n.p1, pr = fix1 ? T : n.p2.transpose();
        if (ev) {
          auto it = store.values.find(v1.id);
          ArrayRef cur;
          if (it != store.values.end()) cur = std::move(it->second.data);
          ArrayRef sunk = (cur || same_operand) ? ArrayRef() : try_gemm_sink(v1.id, lhs.data, rhs.data, pl, pr);
          store.values[v1.id] = Value(sunk ? sunk : ev->matmul_acc(std::move(cur), lhs.data, rhs.data, pl, pr, /*defer=*/!is_variable(v1.id)));
        } else {
          Value l = fix1 ? ga.link(lhs) : lhs, r = fix1 ? rhs : ga.link(rhs);
          Value grad = ga.matmul(l, r, pl, pr);
          store.add_gradient(ga, v1.id, grad);
        }
      }
      if (v2.id.some()) {
        const Value& lhs = fix2 ? g : v1;
        const Value& rhs = fix2 ? v1 : g;
        const MatProp pl = fix2 ? T : n.p1.transpose(), pr = fix2 ? n.p1 : n.p2;
        if (ev) {
          auto it = store.values.find(v2.id);
          ArrayRef cur;
          if (it != store.values.end()) cur = std::move(it->second.data);
          ArrayRef sunk = (cur || same_operand) ? ArrayRef() : try_gemm_sink(v2.id, lhs.data, rhs.data, pl, pr);
          store.values[v2.id] = Value(sunk ? sunk : ev->matmul_acc(std::move(cur), lhs.data, rhs.data, pl, pr, /*defer=*/!is_variable(v2.id)));
        } else {
          Value l = fix2 ? lhs : ga.link(lhs), r = fix2 ? ga.link(rhs) : rhs;
          Value grad = ga.matmul(l, r, pl, pr);
          store.add_gradient(ga, v2.id, grad);
        }
      }
      return;
    }
    case Op::TRANSPOSE:  // matrix.rs:186-190
      if (in(0).some()) {
        Value grad = ga.transpose(g, n.flag);
        store.add_gradient(ga, in(0), grad);
      }
      return;
    case Op::MAX_AS: {  // array_compare.rs:127-136
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      Value lv = ga.link(v);
      Value mask = ga.argmax_as(lv, n.d);
      Value tiled = ga.tile_as(g, lv.data->dims);
      Value grad = ga.mul(tiled, mask);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::SOFTMAX_AS: {  // array_compare.rs:151-165
      const Value& v = n.saved[0];
      if (!v.id.some()) return;
      Value lv = ga.link(v);
      Value res = ga.softmax_as(lv, n.d);
      Value g1 = ga.mul(g, res);
      Value rg = ga.sum_as(g1, n.d);
      Value gt = ga.tile_as(rg, v.data->dims);
      Value g2 = ga.mul(gt, res);
      Value grad = ga.sub(g1, g2);
      store.add_gradient(ga, v.id, grad);
      return;
    }
    case Op::CUSTOM: {
      gadcu_value gv{g.data.get(), g.id.arena, g.id.index};
      gadcu_status st = n.custom->fn(n.custom->user, &ga, &store, &gv);
      if (st != GADCU_OK) fail(st, last_error());
      return;
    }
  }
}

// ---- reverse sweep: graph.rs:172-246 -----------------------------------------------------------
std::unique_ptr<Store> GraphAlgebra::sweep(Algebra& ga, Id gid, Value gradient, bool once) {
  auto store = std::make_unique<Store>();
  store->insert(gid, std::move(gradient));
  std::priority_queue<Id> heap;  // max-heap: strictly descending ids = reverse topological order
  heap.push(gid);
  Id guard = gid.next_id();
  // final-gradient hook: how many nodes still have to run before a variable's gradient is complete
  // (member state is touched only by hooked — data-parallel — sweeps; plain evaluate_gradients(&self) stays re-entrant
  // from several threads, graph.rs:263-274)
  // (Graph1: gradients are data. GraphN — the data-parallel compute_gradients — too: the hook sees the data of the
  // recorded gradient value and answers with the array that replaces it in the store, as a constant)
  g_tape_nodes_hint.store(uint32_t(nodes_.size()), std::memory_order_relaxed);
  const bool hooked = bool(final_hook_);
  std::vector<uint32_t> unhooked;
  std::vector<uint32_t>& waiting = hooked ? waiting_ : unhooked;  // by node index; only variables are counted
  if (hooked) {
    waiting.clear();
    served_.clear();
  }
  auto distinct_variable_inputs = [&](const Node& nd, std::vector<uint32_t>& out) {
    out.clear();
    for (auto& in : nd.inputs) {
      if (!in.some() || in.arena != arena_id_ || in.index > nodes_.size()) continue;
      const Node& src = nodes_[in.index - 1];
      if (src.op != Op::VARIABLE || src.has_update) continue;
      if (std::find(out.begin(), out.end(), in.index) == out.end()) out.push_back(in.index);
    }
  };
  std::vector<uint32_t> scratch;
  if (hooked) {
    waiting.assign(nodes_.size() + 1, 0);
    for (auto& nd : nodes_) {
      if (!nd.has_update) continue;
      distinct_variable_inputs(nd, scratch);
      for (uint32_t v : scratch) waiting[v]++;
    }
  }
  std::vector<char> fired(waiting.size(), 0), was_variable(waiting.size(), 0);  // (a consuming sweep resets the nodes it passes)
  for (size_t i = 1; i < was_variable.size(); i++) was_variable[i] = nodes_[i - 1].op == Op::VARIABLE && !nodes_[i - 1].has_update;
  if (hooked) served_.assign(waiting.size(), 0);
  while (!heap.empty()) {
    Id id = heap.top();
    heap.pop();
    if (!(id < guard)) continue;  // duplicate push
    guard = id;
    if (id.arena != arena_id_ || id.index == 0 || id.index > nodes_.size())
      fail(GADCU_ERR_MISSING_NODE, "Trying to obtain a node from an incorrect `id`.");
    Node& node = nodes_[id.index - 1];
    if (node.has_update) {
      // wrapper of make_generic_node (graph.rs:150-158): fetch own gradient, check its dims
      const Value* own = store->get(id);
      if (!own) fail(GADCU_ERR_MISSING_GRADIENT, "No gradient of the expected type could be found in gradient store.");
      Value g = *own;  // clone out of the store
      if (g.data->dtype != node.fwd_dtype || g.data->is_scalar != node.fwd_scalar)
        fail(GADCU_ERR_TYPE, "gradient type differs from the node's forward type");
      if (!node.fwd_scalar && g.data->dims != node.fwd_dims)
        fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for make_generic_node: " + g.data->dims.str() + " " + node.fwd_dims.str());
      if (once) {
        // the node is cleared after use (graph.rs:242); run the update from a moved-out copy so
        // the captured forward buffers are released as soon as the VJP has been issued
        Node taken = std::move(node);
        node = Node{};
        node.inputs = taken.inputs;
        vjp(taken, ga, *store, g);
      } else {
        vjp(node, ga, *store, g);
      }
    }
    if (!waiting.empty()) {
      distinct_variable_inputs(node, scratch);
      for (uint32_t v : scratch)
        if (--waiting[v] == 0) {
          const Value* gv = store->get(Id{arena_id_, v});
          if (gv && gv->data) {
            fired[v] = 1;
            if (!served_[v]) {
              ArrayRef replacement = final_hook_(Id{arena_id_, v}, gv->data);
              if (replacement) store->values[Id{arena_id_, v}] = Value(replacement);
            }
          }
        }
    }
    for (auto& in : node.inputs)
      if (in.some()) heap.push(in);
    if (once) node.clear();
  }
  if (!waiting.empty())  // variables some unreached node also feeds (or the root itself): complete as far as this sweep goes
    for (auto& kv : store->values) {
      const Id vid = kv.first;
      if (vid.arena == arena_id_ && vid.index < fired.size() && was_variable[vid.index] && !fired[vid.index] && !served_[vid.index] && kv.second.data) {
        ArrayRef replacement = final_hook_(vid, kv.second.data);
        if (replacement) kv.second = Value(replacement);
      }
    }
  if (hooked) waiting_.clear();
  return store;
}

bool GraphAlgebra::structure_hash(uint64_t out[2], const std::function<int64_t(const gadcu_array&)>& register_of) const {
  uint64_t h1 = 1469598103934665603ull, h2 = 0x9E3779B97F4A7C15ull;
  auto mix = [&](uint64_t w) {
    h1 = (h1 ^ w) * 1099511628211ull;
    h2 = (h2 + w) * 0xff51afd7ed558ccdull;
    h2 ^= h2 >> 33;
  };
  if (final_hook_ || higher_order_) return false;
  g_tape_nodes_hint.store(uint32_t(nodes_.size()), std::memory_order_relaxed);  // (a replayed sweep never reaches sweep())
  mix(nodes_.size());
  for (const Node& n : nodes_) {
    if (n.op == Op::CUSTOM || n.custom) return false;
    uint64_t cbits;
    std::memcpy(&cbits, &n.c, 8);
    mix(uint64_t(n.op) | uint64_t(n.has_update) << 8 | uint64_t(n.c_is_int) << 9 | uint64_t(n.flag) << 10 | uint64_t(n.p1.transposed) << 11 |
        uint64_t(n.p2.transposed) << 12 | uint64_t(n.fwd_scalar) << 13 | uint64_t(n.fwd_dtype) << 14 | uint64_t(n.inputs.size()) << 16 |
        uint64_t(n.saved.size()) << 40);
    mix(cbits);
    for (int i = 0; i < 4; i++) mix(n.d[i] * 31 + n.fwd_dims[i]);
    for (const Id& in : n.inputs) mix(in.index);
    for (const Value& v : n.saved) {
      int64_t r = -1;
      if (v.data) r = v.data->sym_reg >= 0 && v.data->sym ? v.data->sym_reg : register_of(*v.data);
      if (v.data && r < 0) return false;  // (a captured value the lane program does not know by register yet)
      mix(uint64_t(uint32_t(r)) | uint64_t(v.id.index) << 32);
    }
  }
  out[0] = h1;
  out[1] = h2;
  return true;
}

void GraphAlgebra::consume(const std::vector<uint32_t>& node_indices) {
  for (uint32_t i : node_indices)
    if (i >= 1 && i <= nodes_.size()) nodes_[i - 1].clear();
}

// graph.rs:263-290: the gradient algebra is a clone of the eval algebra; gradients are pure data.
std::unique_ptr<Store> GraphAlgebra::evaluate_gradients(Id id, ArrayRef seed, bool once) {
  if (higher_order_) fail(GADCU_ERR_INVALID, "evaluate_gradients is defined for Graph1; use compute_gradients on GraphN");
  // the variables' gradients are what the caller came for: their deferred elementwise chains are computed when the
  // sweep ends, all of one region in one kernel (asked before a consuming sweep clears the nodes)
  std::vector<Id> var_ids;
  if (ctx->lazy && ctx->fuse && eval_->is_eval())
    for (uint32_t i = 1; i <= nodes_.size(); i++)
      if (is_variable(Id{arena_id_, i})) var_ids.push_back(Id{arena_id_, i});
  std::unique_ptr<Store> store = sweep(*eval_, id, Value(std::move(seed)), once);
  if (!var_ids.empty()) {
    std::vector<ArrayRef> grads;
    for (Id v : var_ids) {
      auto it = store->values.find(v);
      if (it != store->values.end() && it->second.data) grads.push_back(it->second.data);
    }
    lazy_flush(grads);
  }
  return store;
}

// graph.rs:307-318: walk (and consume) a clone while `self` records the VJP operations.
std::unique_ptr<Store> GraphAlgebra::compute_gradients(Id id, const Value& seed) {
  if (!higher_order_) fail(GADCU_ERR_INVALID, "compute_gradients is defined for GraphN; use evaluate_gradients on Graph1");
  GraphAlgebra current(*this);
  return current.sweep(*this, id, seed, true);
}

}  // namespace gadcu
