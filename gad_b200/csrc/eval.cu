// eval.cu — Check (dimension algebra) and Eval (forward primitives) over device arrays.
// Mirrors the `impl ... for Check` / `impl ... for Eval` blocks over af::Array in
// gad/src/{core,arith,const_arith,analytic,compare,array,array_compare,matrix}.rs: every fallible op
// runs the Check rule on the host first and launches only if it passes.
#include "algebra.h"
#include "batched.h"

namespace gadcu {

// =================================================================================================
// Check
// =================================================================================================
namespace check {

static std::string join(const std::vector<const Dim4*>& dims) {
  std::string s;
  for (auto d : dims) s += d->str() + " ";
  return s;
}
Dim4 equal(const char* name, const std::vector<const Dim4*>& dims) {  // error.rs:139-153
  if (dims.empty()) fail(GADCU_ERR_EMPTY, std::string("Unexpected empty input for ") + name);
  for (auto d : dims)
    if (*d != *dims[0]) fail(GADCU_ERR_DIMENSIONS, std::string("Incompatible dimensions for ") + name + ": " + join(dims));
  return *dims[0];
}
Dim4 equal(const char* name, std::initializer_list<const Dim4*> dims) { return equal(name, std::vector<const Dim4*>(dims)); }
Dim4 select_argmax(const Dim4& v0, const Dim4& v1, const Dim4* r0, const Dim4* r1) {  // compare.rs:128-144
  equal("select_argmax", {&v0, &v1});
  if (r0) equal("select_argmax", {&v0, r0});
  if (r1) equal("select_argmax", {&v1, r1});
  return v0;
}
Dim4 flat(const Dim4& v) { return Dim4(v.elements()); }  // array.rs:133-136
Dim4 moddims(const Dim4& v, const Dim4& d) {             // array.rs:138-145
  if (v.elements() != d.elements()) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for moddims: " + v.str() + " " + d.str());
  return d;
}
Dim4 tile_as(const Dim4& v, const Dim4& r) {  // array.rs:147-157
  for (int i = 0; i < 4; i++)
    if (v[i] == 0 || r[i] % v[i] != 0) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for tile_as: " + v.str() + " " + r.str());
  return r;
}
Dim4 sum_as(const Dim4& v, const Dim4& r) {  // array.rs:159-170
  for (int i = 0; i < 4; i++) {
    if (r[i] == v[i]) continue;
    if (r[i] != 1) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for sum_as: " + v.str() + " " + r.str());
  }
  return r;
}
void as_scalar(const Dim4& v) {  // array.rs:177-181
  Dim4 one;
  equal("as_scalar", {&v, &one});
}
void dot(const Dim4& a, const Dim4& b) { equal("dot", {&a, &b}); }  // array.rs:188-192
Dim4 reduced(const char* name, const Dim4& v, const Dim4& r) {      // error.rs:175-189
  for (int i = 0; i < 4; i++) {
    if (r[i] == v[i]) continue;
    if (r[i] != 1)
      fail(GADCU_ERR_REDUCED_DIMENSIONS, std::string("Incorrect reduction of dimensions for ") + name + ": " + v.str() + " " + r.str());
  }
  return r;
}
Dim4 transpose(const Dim4& v) {  // matrix.rs:118-125
  if (!(v[2] == 1 && v[3] == 1)) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for transpose: " + v.str());
  return Dim4(v[1], v[0], 1, 1);
}
Dim4 matmul(const Dim4& v1, const Dim4& v2, MatProp p1, MatProp p2) {  // matrix.rs:88-116
  Dim4 tv1 = p1.transposed ? transpose(v1) : v1;
  Dim4 tv2 = p2.transposed ? transpose(v2) : v2;
  if (tv1[1] != tv2[0]) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for matmul: " + v1.str() + " " + v2.str());
  uint64_t a, b;
  if (tv1[2] == 1 && tv1[3] == 1) { a = tv2[2]; b = tv2[3]; }
  else if (tv2[2] == 1 && tv2[3] == 1) { a = tv1[2]; b = tv1[3]; }
  else if (tv1[2] == tv2[2] && tv1[3] == tv2[3]) { a = tv1[2]; b = tv1[3]; }
  else fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for matmul: " + v1.str() + " " + v2.str());
  return Dim4(tv1[0], tv2[1], a, b);
}

}  // namespace check

// =================================================================================================
// Algebra defaults
// =================================================================================================
}  // namespace gadcu

using namespace gadcu;

Value gadcu_algebra::add_all(const std::vector<const Value*>& vs) {  // core.rs:27-37
  if (vs.empty()) fail(GADCU_ERR_EMPTY, "Unexpected empty input for add_all");
  Value result = *vs[0];
  for (size_t i = 1; i < vs.size(); i++) result = add(result, *vs[i]);
  return result;
}
Value gadcu_algebra::pow(const Value& v, const Value& p) {  // analytic.rs:48-55
  Value l = log(v);
  Value e = mul(p, l);
  return exp(e);
}
Value gadcu_algebra::abs(const Value& v) {  // compare.rs:27-34
  Value n = neg(v);
  return select_argmax(v, n, &v, &n);
}
Value gadcu_algebra::sign(const Value& v) {  // compare.rs:37-46
  Value n = neg(v);
  Value one = ones(v);
  Value mone = neg(one);
  return select_argmax(v, n, &one, &mone);
}
Value gadcu_algebra::relu(const Value& v) {  // compare.rs:49-55
  Value z = zeros(v);
  return max(z, v);
}

namespace gadcu {

// =================================================================================================
// Eval
// =================================================================================================
namespace {

void require_data(const Value& v, const char* name) {
  if (!v.data) fail(GADCU_ERR_INVALID, std::string(name) + ": value has no data");
}
void same_kind(const char* name, const ArrayRef& a, const ArrayRef& b) {
  if (a->dtype != b->dtype) fail(GADCU_ERR_TYPE, std::string(name) + ": dtype mismatch");
  if (a->is_scalar != b->is_scalar) fail(GADCU_ERR_TYPE, std::string(name) + ": scalar/array mismatch");
}
// operand as the elementwise kernels take it: dense pointer, or one broadcast element
struct Operand {
  ArrayRef keep;
  const void* ptr;
  bool bcast;
};
Operand operand(const ArrayRef& a) {
  if (a->symbolic()) { ArrayRef m = materialize(a); return {m, m->ptr(), false}; }  // a deferred elementwise value: compute it
  if (a->bcast1()) return {a, a->ptr(), true};
  if (a->lazy()) { ArrayRef m = materialize(a); return {m, m->ptr(), false}; }
  return {a, a->ptr(), false};
}
// a constant array of `dims`: one stored element, broadcast lazily (af::constant, arith.rs:49-57)
ArrayRef constant_array(gadcu_ctx* ctx, gadcu_dtype dt, const Dim4& dims, bool is_scalar, double value) {
  ArrayRef one = new_array(ctx, dt, Dim4(), is_scalar);
  k::launch_fill(ctx, dt, one->ptr(), 1, value);
  if (dims.elements() == 1) { one->dims = dims; one->phys = dims; return one; }
  return view_array(one, dims, Dim4(), is_scalar);
}
// T::from(c) for the i16 instantiation, identity for the T one
double const_value(double c, bool is_int) { return is_int ? double(int16_t(c)) : c; }

}  // namespace

ArrayRef EvalAlgebra::fused(k::EwOp op, ArrayRef acc, std::initializer_list<const ArrayRef*> ins, double c, double c2) {
  const ArrayRef& first = **ins.begin();
  gadcu_dtype dt = first->dtype;
  Dim4 dims = first->dims;
  bool is_scalar = first->is_scalar;
  for (auto p : ins) {
    if ((*p)->dtype != dt) fail(GADCU_ERR_TYPE, "fused: dtype mismatch");
    if ((*p)->dims != dims) {  // a broadcast operand adopts the dims of the dense ones
      if (first->bcast1() || first->elements() == 1) { dims = (*p)->dims; is_scalar = (*p)->is_scalar; }
    }
  }
  if (acc) {
    if (acc->dtype != dt) fail(GADCU_ERR_TYPE, "fused: accumulator dtype mismatch");
    if (acc->dims != dims) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for add: " + acc->dims.str() + " " + dims.str());
  }
  {
    // large elementwise work is recorded into the tape's lane program instead of launched (batched.cu): the
    // chain it belongs to becomes one kernel when its bytes are needed
    bool every_bcast = !acc || acc->bcast1();
    for (auto p : ins) every_bcast = every_bcast && (*p)->bcast1();
    if (!every_bcast && lazy_enabled(ctx, dims.elements(), is_scalar))
      return lazy_elementwise(ctx, op, acc, std::vector<const ArrayRef*>(ins), c, c2, dt, dims, is_scalar);
  }
  // decided before any further handle copies are made: is `acc` ours alone?
  const bool inplace = acc && ctx->fuse && acc->unique();
  k::EwArgs e;
  e.dt = dt;
  e.c = c;
  e.c2 = c2;
  std::vector<Operand> keep;
  bool all_bcast = true;
  for (auto p : ins) {
    Operand o = operand(*p);
    e.in[e.nin] = o.ptr;
    e.bcast[e.nin] = o.bcast;
    e.nin++;
    all_bcast = all_bcast && o.bcast;
    keep.push_back(std::move(o));
  }
  Operand acc_op{};
  if (acc) {
    if (acc->dtype != dt) fail(GADCU_ERR_TYPE, "fused: accumulator dtype mismatch");
    if (acc->dims != dims) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for add: " + acc->dims.str() + " " + dims.str());
    acc_op = operand(acc);
    e.acc = acc_op.ptr;
    e.acc_bcast = acc_op.bcast;
    all_bcast = all_bcast && acc_op.bcast;
  }
  // everything is one broadcast element: compute one element and keep the result lazy
  if (all_bcast && dims.elements() > 1) {
    ArrayRef one = new_array(ctx, dt, Dim4(), is_scalar);
    e.out = one->ptr();
    e.n = 1;
    k::launch_ew(ctx, op, e);
    return view_array(one, dims, Dim4(), is_scalar);
  }
  ArrayRef out;
  if (inplace) {
    out = acc;  // update in place: nobody else can observe this buffer
    out->buf->mutated();
  } else {
    out = new_array(ctx, dt, dims, is_scalar);
  }
  e.out = out->ptr();
  e.n = dims.elements();
  k::launch_ew(ctx, op, e);
  return out;
}

Value EvalAlgebra::unary(k::EwOp op, const Value& v, double c) {
  require_data(v, "unary");
  return Value(fused(op, ArrayRef(), {&v.data}, c));
}
Value EvalAlgebra::binary(k::EwOp op, const char* name, const Value& a, const Value& b) {
  require_data(a, name);
  require_data(b, name);
  same_kind(name, a.data, b.data);
  check::equal(name, {&a.data->dims, &b.data->dims});
  return Value(fused(op, ArrayRef(), {&a.data, &b.data}));
}

Value EvalAlgebra::add(const Value& a, const Value& b) { return binary(k::EwOp::ADD, "add", a, b); }  // core.rs:179-182
Value EvalAlgebra::sub(const Value& a, const Value& b) { return binary(k::EwOp::SUB, "sub", a, b); }  // arith.rs:65-68
Value EvalAlgebra::mul(const Value& a, const Value& b) { return binary(k::EwOp::MUL, "mul", a, b); }  // arith.rs:71-74
Value EvalAlgebra::div(const Value& a, const Value& b) { return binary(k::EwOp::DIV, "div", a, b); }  // analytic.rs:119-122
Value EvalAlgebra::pow(const Value& a, const Value& b) { return binary(k::EwOp::POW, "pow", a, b); }  // analytic.rs:124-127
Value EvalAlgebra::min(const Value& a, const Value& b) { return binary(k::EwOp::MIN, "min", a, b); }
Value EvalAlgebra::max(const Value& a, const Value& b) { return binary(k::EwOp::MAX, "max", a, b); }

Value EvalAlgebra::zeros(const Value& v) {  // arith.rs:49-52
  require_data(v, "zeros");
  return Value(constant_array(ctx, v.data->dtype, v.data->dims, v.data->is_scalar, 0.0));
}
Value EvalAlgebra::ones(const Value& v) {  // arith.rs:54-57
  require_data(v, "ones");
  return Value(constant_array(ctx, v.data->dtype, v.data->dims, v.data->is_scalar, 1.0));
}
Value EvalAlgebra::neg(const Value& v) { return unary(k::EwOp::NEG, v); }  // arith.rs:59-62
Value EvalAlgebra::setc(const Value& v, double c, bool is_int) {            // const_arith.rs:41-43
  require_data(v, "setc");
  return Value(constant_array(ctx, v.data->dtype, v.data->dims, v.data->is_scalar, const_value(c, is_int)));
}
Value EvalAlgebra::addc(const Value& v, double c, bool is_int) { return unary(k::EwOp::ADDC, v, const_value(c, is_int)); }
Value EvalAlgebra::mulc(const Value& v, double c, bool is_int) { return unary(k::EwOp::MULC, v, const_value(c, is_int)); }
Value EvalAlgebra::powc(const Value& v, double c, bool is_int) { return unary(k::EwOp::POWC, v, const_value(c, is_int)); }
Value EvalAlgebra::exp(const Value& v) { return unary(k::EwOp::EXP, v); }
Value EvalAlgebra::log(const Value& v) { return unary(k::EwOp::LOG, v); }
Value EvalAlgebra::log1p(const Value& v) { return unary(k::EwOp::LOG1P, v); }
Value EvalAlgebra::sin(const Value& v) { return unary(k::EwOp::SIN, v); }
Value EvalAlgebra::cos(const Value& v) { return unary(k::EwOp::COS, v); }
Value EvalAlgebra::tanh(const Value& v) { return unary(k::EwOp::TANH, v); }
Value EvalAlgebra::sigmoid(const Value& v) { return unary(k::EwOp::SIGMOID, v); }
Value EvalAlgebra::reciprocal(const Value& v) { return unary(k::EwOp::RECIPROCAL, v); }
Value EvalAlgebra::sqrt(const Value& v) { return unary(k::EwOp::SQRT, v); }

Value EvalAlgebra::select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) {  // compare.rs:94-114
  require_data(v0, "select_argmax");
  require_data(v1, "select_argmax");
  same_kind("select_argmax", v0.data, v1.data);
  if (r0) same_kind("select_argmax", v0.data, r0->data);
  if (r1) same_kind("select_argmax", v0.data, r1->data);
  check::select_argmax(v0.data->dims, v1.data->dims, r0 ? &r0->data->dims : nullptr, r1 ? &r1->data->dims : nullptr);
  if (r0 && r1) return Value(fused(k::EwOp::SELECT_BOTH, ArrayRef(), {&v0.data, &v1.data, &r0->data, &r1->data}));
  if (r0) return Value(fused(k::EwOp::SELECT_R0, ArrayRef(), {&v0.data, &v1.data, &r0->data}));
  if (r1) return Value(fused(k::EwOp::SELECT_R1, ArrayRef(), {&v0.data, &v1.data, &r1->data}));
  return zeros(v0);
}

Value EvalAlgebra::flat(const Value& v) {  // array.rs:65-67
  require_data(v, "flat");
  Dim4 d = check::flat(v.data->dims);
  if (v.data->bcast1()) return Value(view_array(v.data, d, Dim4(), false));
  ArrayRef src = materialize(v.data);
  return Value(view_array(src, d, d, false));
}
Value EvalAlgebra::moddims(const Value& v, const Dim4& d) {  // array.rs:70-73
  require_data(v, "moddims");
  check::moddims(v.data->dims, d);
  if (v.data->bcast1()) return Value(view_array(v.data, d, Dim4(), false));
  ArrayRef src = materialize(v.data);
  return Value(view_array(src, d, d, false));
}
Value EvalAlgebra::tile_as(const Value& v, const Dim4& rd) {  // array.rs:76-84
  require_data(v, "tile_as");
  check::tile_as(v.data->dims, rd);
  if (v.data->dims == rd) return Value(v.data);
  if (v.data->phys.elements() == 1) return Value(view_array(v.data, rd, Dim4(), false));  // lazy scalar broadcast
  ArrayRef src = materialize(v.data);
  if (lazy_enabled(ctx, rd.elements(), false)) {
    // a row or a column repeated over the other dimension (the bias of a layer): stays a view; the elementwise
    // kernel that consumes it indexes the small array
    const Dim4& sd = src->dims;
    const bool col = sd[0] == 1 && sd[1] == rd[1] && sd[2] == 1 && sd[3] == 1 && rd[2] == 1 && rd[3] == 1;
    const bool row = sd[0] == rd[0] && sd[1] == 1 && sd[2] == 1 && sd[3] == 1;
    if (col || row) return Value(view_array(src, rd, sd, false));
  }
  ArrayRef out = new_array(ctx, src->dtype, rd);
  k::launch_tile(ctx, src->dtype, src->ptr(), src->dims, out->ptr(), rd);
  return Value(out);
}

// Reduce every axis where rd differs from v (rd[i] == 1). Consecutive reduced axes form one pass
// (the reference issues one af::sum / af::max per axis: array.rs:91-96, array_compare.rs:42-47).
ArrayRef EvalAlgebra::reduce_as(k::RedOp op, const ArrayRef& vin, const Dim4& rd, ArrayRef acc) {
  if (!acc && vin->symbolic() && !vin->lazy() && rd != vin->dims) {
    // a deferred elementwise value reduced over its leading axes only (one pass, contiguous segments): computed and
    // reduced by one kernel, never written
    const Dim4& d = vin->dims;
    int last = -1;
    bool leading = true;
    for (int i = 0; i < 4; i++)
      if (rd[i] != d[i]) { leading = leading && i == last + 1; last = i; }
    if (leading) {
      uint64_t red = 1, outer = 1;
      for (int i = 0; i <= last; i++) red *= d[i];
      for (int i = last + 1; i < 4; i++) outer *= d[i];
      Dim4 nd = d;
      for (int i = 0; i <= last; i++) nd[i] = 1;
      ArrayRef r = lazy_reduce(ctx, op, vin, red, outer, nd);
      if (r) return r;
    }
  }
  ArrayRef cur = materialize(vin);
  Dim4 cd = cur->dims;
  std::vector<std::pair<int, int>> runs;  // [first, last] axis of each run of reduced axes
  for (int i = 0; i < 4;) {
    if (rd[i] != cd[i]) {
      int j = i;
      while (j + 1 < 4 && rd[j + 1] != cd[j + 1]) j++;
      runs.push_back({i, j});
      i = j + 1;
    } else {
      i++;
    }
  }
  if (runs.empty()) {
    if (!acc) return cur;
    return fused(k::EwOp::COPY, acc, {&cur});
  }
  for (size_t r = 0; r < runs.size(); r++) {
    uint64_t inner = 1, red = 1, outer = 1;
    for (int i = 0; i < runs[r].first; i++) inner *= cd[i];
    for (int i = runs[r].first; i <= runs[r].second; i++) red *= cd[i];
    for (int i = runs[r].second + 1; i < 4; i++) outer *= cd[i];
    Dim4 nd = cd;
    for (int i = runs[r].first; i <= runs[r].second; i++) nd[i] = 1;
    const bool last = r + 1 == runs.size();
    ArrayRef out;
    const void* acc_ptr = nullptr;
    ArrayRef acc_keep;
    if (last && acc) {
      if (acc->dims != nd) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for add: " + acc->dims.str() + " " + nd.str());
      const bool inplace = ctx->fuse && acc->unique();
      acc_keep = materialize(acc);
      acc_ptr = acc_keep->ptr();
      if (inplace) { out = acc; out->buf->mutated(); } else { out = new_array(ctx, cur->dtype, nd); }
    } else {
      out = new_array(ctx, cur->dtype, nd);
    }
    k::launch_reduce(ctx, cur->dtype, op, cur->ptr(), out->ptr(), inner, red, outer, acc_ptr);
    cur = out;
    cd = nd;
  }
  return cur;
}
Value EvalAlgebra::sum_as(const Value& v, const Dim4& rd) {  // array.rs:87-98
  require_data(v, "sum_as");
  check::sum_as(v.data->dims, rd);
  return Value(reduce_as(k::RedOp::SUM, v.data, rd, ArrayRef()));
}
ArrayRef EvalAlgebra::sum_as_acc(ArrayRef acc, const ArrayRef& v, const Dim4& rd) {
  check::sum_as(v->dims, rd);
  return reduce_as(k::RedOp::SUM, v, rd, std::move(acc));
}
Value EvalAlgebra::max_as(const Value& v, const Dim4& rd) {  // array_compare.rs:38-49
  require_data(v, "max_as");
  check::reduced("max_as", v.data->dims, rd);
  return Value(reduce_as(k::RedOp::MAX, v.data, rd, ArrayRef()));
}
Value EvalAlgebra::argmax_as(const Value& v, const Dim4& rd) {  // array_compare.rs:51-65
  require_data(v, "argmax_as");
  check::reduced("argmax_as", v.data->dims, rd);
  Dim4 dims = v.data->dims;
  Value mx = tile_as(max_as(v, rd), dims);
  Value one = ones(v);
  Value arg = select_argmax(v, mx, &one, nullptr);
  Value sum = tile_as(sum_as(arg, rd), dims);
  return div(arg, sum);
}
Value EvalAlgebra::softmax_as(const Value& v, const Dim4& rd) {  // array_compare.rs:67-82
  require_data(v, "softmax_as");
  check::reduced("softmax_as", v.data->dims, rd);
  Dim4 dims = v.data->dims;
  Value mx = tile_as(max_as(v, rd), dims);
  Value e = exp(sub(v, mx));
  Value sum = tile_as(sum_as(e, rd), dims);
  return div(e, sum);
}
Value EvalAlgebra::constant_as(const Value& s, const Dim4& d) {  // array.rs:101-103
  require_data(s, "constant_as");
  if (s.data->phys.elements() != 1) fail(GADCU_ERR_TYPE, "constant_as: expected a scalar");
  if (d.elements() == 1) return Value(view_array(s.data, d, d, false));
  return Value(view_array(s.data, d, Dim4(), false));
}
Value EvalAlgebra::as_scalar(const Value& v) {  // array.rs:106-111 (the D2H read is deferred to scalar_to_host)
  require_data(v, "as_scalar");
  check::as_scalar(v.data->dims);
  return Value(view_array(v.data, Dim4(), Dim4(), true));
}
Value EvalAlgebra::scale(const Value& lambda, const Value& v) {  // array.rs:114-116: v * lambda
  require_data(lambda, "scale");
  require_data(v, "scale");
  if (lambda.data->phys.elements() != 1) fail(GADCU_ERR_TYPE, "scale: lambda must be a scalar");
  if (lambda.data->dtype != v.data->dtype) fail(GADCU_ERR_TYPE, "scale: dtype mismatch");
  ArrayRef lam = lambda.data->dims.elements() == 1 && v.data->elements() != 1
                     ? view_array(lambda.data, v.data->dims, Dim4(), v.data->is_scalar)
                     : lambda.data;
  return Value(fused(k::EwOp::SCALE, ArrayRef(), {&v.data, &lam}));
}
Value EvalAlgebra::dot(const Value& a, const Value& b) {  // array.rs:119-126
  require_data(a, "dot");
  require_data(b, "dot");
  same_kind("dot", a.data, b.data);
  check::dot(a.data->dims, b.data->dims);
  ArrayRef x = materialize(a.data), y = materialize(b.data);
  ArrayRef out = new_array(ctx, x->dtype, Dim4(), true);
  k::launch_dot(ctx, x->dtype, x->ptr(), y->ptr(), false, false, out->ptr(), x->elements());
  return Value(out);
}

ArrayRef EvalAlgebra::matmul_acc(ArrayRef acc, const ArrayRef& a_in, const ArrayRef& b_in, MatProp p1, MatProp p2, bool defer) {
  if (a_in->dtype != b_in->dtype) fail(GADCU_ERR_TYPE, "matmul: dtype mismatch");
  Dim4 rd = check::matmul(a_in->dims, b_in->dims, p1, p2);
  ArrayRef a = materialize(a_in), b = materialize(b_in);
  const Dim4 &ad = a->dims, &bd = b->dims;
  k::GemmArgs g;
  g.dt = a->dtype;
  g.A = a->ptr();
  g.B = b->ptr();
  g.M = rd[0];
  g.N = rd[1];
  g.K = p1.transposed ? ad[0] : ad[1];
  g.ta = p1.transposed;
  g.tb = p2.transposed;
  g.lda = ad[0];
  g.ldb = bd[0];
  g.batch = rd[2] * rd[3];
  g.strideA = (ad[2] == 1 && ad[3] == 1) ? 0 : ad[0] * ad[1];
  g.strideB = (bd[2] == 1 && bd[3] == 1) ? 0 : bd[0] * bd[1];
  g.strideC = rd[0] * rd[1];
  if (a->ptr() == a->buf->ptr) g.a_buf = a->buf.get();
  if (b->ptr() == b->buf->ptr) g.b_buf = b->buf.get();
  // A product nothing is accumulated into, whose [M, N] result would be deferred elementwise territory anyway, is recorded
  // rather than launched: what the layer does next to it (+ bias, tanh; the tanh VJP on an incoming dA; the residual of
  // the square loss) then runs in the GEMM's epilogue (batched.cu: lazy_matmul / resolve_pending).
  // Worth it when the product's output tiles keep at least half of the chip's SM pairs busy. With several waves of tiles
  // the tail of tile i runs while the tensor cores work on tile i + 1 (double-buffered accumulators); with one well-filled
  // wave it adds a few microseconds to the launch but saves the separate elementwise and split kernels (measured inside
  // one call, 8 GPUs, 64-tile products of the 1024-row shards: 636.6 steps/s fused, 592.1 not). With a handful of tiles it
  // would sit on the critical path of a few SMs while the rest of the chip idles — there the chip-wide elementwise kernel
  // is faster (a 512 x 256 step, 2 tiles per product: 0.31 ms unfused, 0.39 fused).
  // gadcu_ctx_set_gemm_epilogue(ctx, 2) fuses regardless (the parity tests do, at small shapes).
  const uint64_t tiles = ((g.M + 255) / 256) * ((g.N + 255) / 256);
  const bool worth_it = ctx->gemm_epilogue == 2 || (ctx->gemm_epilogue == 1 && tiles * 2 >= uint64_t(ctx->sm_count / 2));
  if (!acc && defer && worth_it && g.batch == 1 && g.dt == GADCU_F32 && lazy_enabled(ctx, rd.elements(), false) &&
      k::gemm_epilogue_supported(g))
    return lazy_matmul(ctx, a, b, p1, p2, rd, fuse_tails);
  ArrayRef out;
  if (acc) {
    if (acc->dims != rd) fail(GADCU_ERR_DIMENSIONS, "Incompatible dimensions for add: " + acc->dims.str() + " " + rd.str());
    if (acc->unique() && ctx->fuse) {
      out = acc;
      out->buf->mutated();
      g.accumulate = true;
    } else {
      // current + new into a fresh buffer: product first, then one add pass
      ArrayRef prod = new_array(ctx, a->dtype, rd);
      g.C = prod->ptr();
      k::launch_gemm(ctx, g);
      return fused(k::EwOp::COPY, acc, {&prod});
    }
  } else {
    out = new_array(ctx, a->dtype, rd);
  }
  g.C = out->ptr();
  k::launch_gemm(ctx, g);
  return out;
}
Value EvalAlgebra::matmul(const Value& a, const Value& b, MatProp p1, MatProp p2) {  // matrix.rs:44-54
  require_data(a, "matmul");
  require_data(b, "matmul");
  return Value(matmul_acc(ArrayRef(), a.data, b.data, p1, p2, /*defer=*/true));
}
Value EvalAlgebra::transpose(const Value& v, bool /*conjugate: identity on real types*/) {  // matrix.rs:57-60
  require_data(v, "transpose");
  Dim4 rd = check::transpose(v.data->dims);
  ArrayRef src = materialize(v.data);
  ArrayRef out = new_array(ctx, src->dtype, rd);
  k::launch_transpose(ctx, src->dtype, src->ptr(), out->ptr(), src->dims[0], src->dims[1]);
  return Value(out);
}

// store.rs:52 `*current = graph.add(current, value)`; `current` is consumed, so when nobody else
// holds its buffer the sum is written in place (same values, no fresh allocation, one pass less).
Value EvalAlgebra::accumulate(Value&& current, const Value& value) {
  same_kind("add", current.data, value.data);
  check::equal("add", {&current.data->dims, &value.data->dims});
  ArrayRef cur = std::move(current.data);
  return Value(fused(k::EwOp::COPY, std::move(cur), {&value.data}));
}

}  // namespace gadcu
