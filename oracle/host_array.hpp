// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by
// the product (gad_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and there only as the checker / reported CPU baseline.
//
// host_array.hpp: a host, column-major, Dim4-shaped array with the primitive operations that
// gad's `Eval for af::Array<T>` impls call into ArrayFire for (third-party crate `arrayfire`
// 3.8.0, Cargo.lock:25-28, NOT vendored under /root/reference, so its published semantics are
// restated here: column-major storage, dim 0 fastest, Dim4 with trailing ones).
//
// PARITY STATUS of this file: "parity unpinned" at the bit level for floating-point arrays — the
// reference's own array tests use unseeded af::randu inputs and finite differences at 1e-3
// (gad/src/arrayfire.rs:120-163), so there are no golden vectors for array arithmetic. What IS
// pinned (exactly) are the Check dimension rules and the tape semantics, see gad_oracle.hpp.
// Reduction order used here (and documented for the tolerance in tests): plain sequential
// accumulation in index order, in type T; an f64-accumulated variant is offered as a second
// reference for the tolerance bound.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace gad_oracle {

// ---------------------------------------------------------------------------------------------
// Errors — gad/src/error.rs:9-40 (variants), :139-167 (check_equal_*), :170-190 (reduced dims).
// ---------------------------------------------------------------------------------------------
enum class ErrKind : int {
  Dimensions = 1,
  ReducedDimensions = 2,
  Lengths = 3,
  Empty = 4,
  MissingId = 5,
  MissingGradient = 6,
  MissingNode = 7,
};

struct Error : std::exception {
  ErrKind kind;
  std::string name;    // reference: Rust type_name of the function (compiler dependent; not compared)
  std::string detail;  // debug print of the offending dims / lengths
  std::string text;
  Error(ErrKind k, std::string n, std::string d = {}) : kind(k), name(std::move(n)), detail(std::move(d)) {
    text = "gad error kind=" + std::to_string(int(kind)) + " in " + name + ": " + detail;
  }
  const char* what() const noexcept override { return text.c_str(); }
};

// Dims of a scalar: Rust `()` (gad/src/core.rs:58-63).
struct Unit {
  bool operator==(const Unit&) const { return true; }
  bool operator!=(const Unit&) const { return false; }
};
inline std::string debug(const Unit&) { return "()"; }

// af::Dim4: four u64, trailing ones; Default is [1,1,1,1] (used as "reduce to scalar",
// gad/src/array.rs:283).
struct Dim4 {
  std::array<uint64_t, 4> d{1, 1, 1, 1};
  Dim4() = default;
  Dim4(uint64_t a, uint64_t b = 1, uint64_t c = 1, uint64_t e = 1) : d{a, b, c, e} {}
  uint64_t operator[](int i) const { return d[i]; }
  uint64_t& operator[](int i) { return d[i]; }
  uint64_t elements() const { return d[0] * d[1] * d[2] * d[3]; }
  bool operator==(const Dim4& o) const { return d == o.d; }
  bool operator!=(const Dim4& o) const { return !(d == o.d); }
};
inline std::string debug(const Dim4& v) {
  return "[" + std::to_string(v[0]) + " " + std::to_string(v[1]) + " " + std::to_string(v[2]) + " " +
         std::to_string(v[3]) + "]";
}

// error.rs:139-153
template <class D>
D check_equal_dimensions(const char* name, std::initializer_list<const D*> dims) {
  auto it = dims.begin();
  if (it == dims.end()) throw Error(ErrKind::Empty, name);
  const D* first = *it;
  for (++it; it != dims.end(); ++it) {
    if (!(**it == *first)) {
      std::string s;
      for (auto p : dims) s += debug(*p) + " ";
      throw Error(ErrKind::Dimensions, name, s);
    }
  }
  return *first;
}
template <class D>
D check_equal_dimensions(const char* name, const std::vector<const D*>& dims) {
  if (dims.empty()) throw Error(ErrKind::Empty, name);
  for (auto p : dims)
    if (!(*p == *dims[0])) {
      std::string s;
      for (auto q : dims) s += debug(*q) + " ";
      throw Error(ErrKind::Dimensions, name, s);
    }
  return *dims[0];
}

// error.rs:156-167
inline size_t check_equal_lengths(const char* name, const std::vector<size_t>& lengths) {
  if (lengths.empty()) throw Error(ErrKind::Empty, name);
  for (auto l : lengths)
    if (l != lengths[0]) throw Error(ErrKind::Lengths, name);
  return lengths[0];
}

// error.rs:175-189
inline Dim4 check_reduced_dimensions(const char* name, const Dim4& dims, const Dim4& rdims) {
  for (int i = 0; i < 4; i++) {
    if (rdims[i] == dims[i]) continue;
    if (rdims[i] != 1) throw Error(ErrKind::ReducedDimensions, name, debug(dims) + " " + debug(rdims));
  }
  return rdims;
}

// MatProp — gad/src/matrix.rs:13-17, 202-221.
struct MatProp {
  bool transposed = false;
  bool conjugated = false;
  MatProp transpose() const { return {!transposed, conjugated}; }
  MatProp conjugate() const { return {transposed, !conjugated}; }
  bool operator==(const MatProp& o) const { return transposed == o.transposed && conjugated == o.conjugated; }
};

// ---------------------------------------------------------------------------------------------
// HostArray<T>: immutable value with O(1) clone (shared buffer), like af::Array.
// ---------------------------------------------------------------------------------------------
template <class T>
struct HostArray {
  Dim4 dims_;
  std::shared_ptr<std::vector<T>> buf;

  HostArray() : buf(std::make_shared<std::vector<T>>(1, T(0))) {}
  explicit HostArray(Dim4 d) : dims_(d), buf(std::make_shared<std::vector<T>>(d.elements())) {}
  HostArray(Dim4 d, const T* src) : dims_(d), buf(std::make_shared<std::vector<T>>(src, src + d.elements())) {}
  HostArray(Dim4 d, T fill) : dims_(d), buf(std::make_shared<std::vector<T>>(d.elements(), fill)) {}

  const Dim4& dims() const { return dims_; }
  size_t size() const { return buf->size(); }
  const T* data() const { return buf->data(); }
  T* data() { return buf->data(); }
  const T& operator[](size_t i) const { return (*buf)[i]; }
  T& operator[](size_t i) { return (*buf)[i]; }
};

namespace ha {

template <class T, class F>
HostArray<T> map1(const HostArray<T>& a, F f) {
  HostArray<T> r(a.dims());
  for (size_t i = 0; i < a.size(); i++) r[i] = f(a[i]);
  return r;
}
template <class T, class F>
HostArray<T> map2(const HostArray<T>& a, const HostArray<T>& b, F f) {
  HostArray<T> r(a.dims());
  for (size_t i = 0; i < a.size(); i++) r[i] = f(a[i], b[i]);
  return r;
}

// af::tile(v, rdims / vdims)
template <class T>
HostArray<T> tile(const HostArray<T>& v, const Dim4& rd) {
  const Dim4& vd = v.dims();
  HostArray<T> r(rd);
  size_t o = 0;
  for (uint64_t l = 0; l < rd[3]; l++)
    for (uint64_t k = 0; k < rd[2]; k++)
      for (uint64_t j = 0; j < rd[1]; j++)
        for (uint64_t i = 0; i < rd[0]; i++) {
          size_t src = (i % vd[0]) + vd[0] * ((j % vd[1]) + vd[1] * ((k % vd[2]) + vd[2] * (l % vd[3])));
          r[o++] = v[src];
        }
  return r;
}

// af::sum(v, axis): one axis, sequential accumulation in Acc along that axis.
template <class T, class Acc = T>
HostArray<T> sum_axis(const HostArray<T>& v, int axis) {
  Dim4 vd = v.dims(), rd = v.dims();
  rd[axis] = 1;
  HostArray<T> r(rd);
  uint64_t stride = 1;
  for (int i = 0; i < axis; i++) stride *= vd[i];
  uint64_t n = vd[axis];
  uint64_t outer = v.size() / (stride * n);
  for (uint64_t o = 0; o < outer; o++)
    for (uint64_t s = 0; s < stride; s++) {
      Acc acc = Acc(0);
      for (uint64_t k = 0; k < n; k++) acc += Acc(v[s + stride * (k + n * o)]);
      r[s + stride * o] = T(acc);
    }
  return r;
}

// af::max(v, axis)
template <class T>
HostArray<T> max_axis(const HostArray<T>& v, int axis) {
  Dim4 vd = v.dims(), rd = v.dims();
  rd[axis] = 1;
  HostArray<T> r(rd);
  uint64_t stride = 1;
  for (int i = 0; i < axis; i++) stride *= vd[i];
  uint64_t n = vd[axis];
  uint64_t outer = v.size() / (stride * n);
  for (uint64_t o = 0; o < outer; o++)
    for (uint64_t s = 0; s < stride; s++) {
      T acc = v[s + stride * (n * o)];
      for (uint64_t k = 1; k < n; k++) {
        T x = v[s + stride * (k + n * o)];
        if (x > acc) acc = x;
      }
      r[s + stride * o] = acc;
    }
  return r;
}

// af::dot(flat(v1), flat(v2), CONJ, NONE): for real T, plain dot; sequential in Acc.
template <class T, class Acc = T>
T dot(const HostArray<T>& a, const HostArray<T>& b) {
  Acc acc = Acc(0);
  for (size_t i = 0; i < a.size(); i++) acc += Acc(a[i]) * Acc(b[i]);
  return T(acc);
}

// af::transpose (2-D)
template <class T>
HostArray<T> transpose(const HostArray<T>& v) {
  const Dim4& d = v.dims();
  HostArray<T> r(Dim4(d[1], d[0], 1, 1));
  for (uint64_t j = 0; j < d[1]; j++)
    for (uint64_t i = 0; i < d[0]; i++) r[j + d[1] * i] = v[i + d[0] * j];
  return r;
}

// af::matmul(v1, v2, p1, p2) with batch broadcast over dims 2,3 (gad/src/matrix.rs:108-114 gives
// the accepted batch shapes). Accumulates over k in increasing order in Acc.
template <class T, class Acc = T>
HostArray<T> matmul(const HostArray<T>& a, const HostArray<T>& b, bool ta, bool tb, const Dim4& rd) {
  const Dim4 &ad = a.dims(), &bd = b.dims();
  uint64_t M = rd[0], N = rd[1];
  uint64_t K = ta ? ad[0] : ad[1];
  HostArray<T> r(rd);
  uint64_t a_mat = ad[0] * ad[1], b_mat = bd[0] * bd[1];
  for (uint64_t l = 0; l < rd[3]; l++)
    for (uint64_t kk = 0; kk < rd[2]; kk++) {
      const T* A = a.data() + a_mat * ((kk % ad[2]) + ad[2] * (l % ad[3]));
      const T* B = b.data() + b_mat * ((kk % bd[2]) + bd[2] * (l % bd[3]));
      T* C = r.data() + M * N * (kk + rd[2] * l);
      for (uint64_t n = 0; n < N; n++)
        for (uint64_t m = 0; m < M; m++) {
          Acc acc = Acc(0);
          for (uint64_t k = 0; k < K; k++) {
            T x = ta ? A[k + ad[0] * m] : A[m + ad[0] * k];
            T y = tb ? B[n + bd[0] * k] : B[k + bd[0] * n];
            acc += Acc(x) * Acc(y);
          }
          C[m + M * n] = T(acc);
        }
    }
  return r;
}

}  // namespace ha
}  // namespace gad_oracle
