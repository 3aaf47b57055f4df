// CPU test of gadcu::SmallVec (gad_b200/csrc/algebra.h): the container of a tape node's inputs and captured values.
// Element type with counted constructions / destructions so that leaks, double destruction and lost moves show.
#include <cstdio>
#include <utility>
#include <vector>

#include "algebra.h"

// (gadcu_array's destructor refers to it; no array is ever created here)
void gadcu::Buffer::release() {}

static long alive = 0;
struct Counted {
  int v = -1;
  Counted() { alive++; }
  explicit Counted(int x) : v(x) { alive++; }
  Counted(const Counted& o) : v(o.v) { alive++; }
  Counted(Counted&& o) noexcept : v(o.v) { o.v = -2; alive++; }
  Counted& operator=(const Counted& o) { v = o.v; return *this; }
  Counted& operator=(Counted&& o) noexcept { v = o.v; o.v = -2; return *this; }
  ~Counted() { alive--; }
};

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

using Vec = gadcu::SmallVec<Counted, 2>;

static bool holds(const Vec& v, std::initializer_list<int> want) {
  if (v.size() != want.size()) return false;
  size_t i = 0;
  for (int w : want)
    if (v[i++].v != w) return false;
  return true;
}

int main() {
  {
    Vec a;
    CHECK(a.empty() && a.size() == 0 && a.begin() == a.end());
    a.push_back(Counted(1));
    a.push_back(Counted(2));
    CHECK(holds(a, {1, 2}) && alive == 2);  // inline
    a.push_back(Counted(3));                // spills to the heap, elements moved
    a.push_back(Counted(4));
    a.push_back(Counted(5));
    CHECK(holds(a, {1, 2, 3, 4, 5}) && alive == 5);
    Vec b = a;                              // copy of a heap-backed vector
    CHECK(holds(b, {1, 2, 3, 4, 5}) && holds(a, {1, 2, 3, 4, 5}) && alive == 10);
    Vec c = std::move(a);                   // steals the heap block
    CHECK(holds(c, {1, 2, 3, 4, 5}) && a.empty() && alive == 10);
    a.push_back(Counted(9));                // a moved-from vector is an empty, usable one
    CHECK(holds(a, {9}) && alive == 11);
    b = Vec{Counted(7), Counted(8)};        // move-assign from an inline vector over a heap-backed one
    CHECK(holds(b, {7, 8}) && alive == 8);
    c = b;                                  // copy-assign inline over heap
    CHECK(holds(c, {7, 8}) && alive == 5);
    c = c;                                  // self assignment
    CHECK(holds(c, {7, 8}) && alive == 5);
    Vec d = std::move(b);                   // move of an inline vector: element-wise
    CHECK(holds(d, {7, 8}) && b.empty() && alive == 5);
    d.clear();
    CHECK(d.empty() && alive == 3);
    d = {Counted(4), Counted(5), Counted(6)};  // initializer list longer than the inline capacity
    CHECK(holds(d, {4, 5, 6}) && alive == 6);
    std::vector<Counted> sv;
    sv.emplace_back(11);
    sv.emplace_back(12);
    sv.emplace_back(13);
    d = sv;                                 // from a std::vector (gadcu_make_node's input ids)
    CHECK(holds(d, {11, 12, 13}) && alive == 9);
    int sum = 0;
    for (const Counted& x : d) sum += x.v;
    CHECK(sum == 36);
    // a vector of nodes-worth of SmallVecs growing and being moved, as the tape does
    std::vector<Vec> many;
    for (int i = 0; i < 100; i++) {
      Vec v;
      for (int j = 0; j <= i % 4; j++) v.push_back(Counted(i * 10 + j));
      many.push_back(std::move(v));
    }
    bool ok = true;
    for (int i = 0; i < 100; i++) {
      ok = ok && many[size_t(i)].size() == size_t(i % 4 + 1);
      for (int j = 0; j <= i % 4; j++) ok = ok && many[size_t(i)][size_t(j)].v == i * 10 + j;
    }
    CHECK(ok);
  }
  CHECK(alive == 0);  // everything constructed was destroyed exactly once
  // gadcu::Node: clear() empties both vectors; a moved node leaves an empty one behind
  {
    gadcu::Node n;
    n.inputs = {gadcu::Id{1, 2}, gadcu::Id{1, 3}};
    n.saved = {gadcu::Value(), gadcu::Value(), gadcu::Value()};
    n.has_update = true;
    gadcu::Node m = std::move(n);
    CHECK(m.inputs.size() == 2 && m.saved.size() == 3 && m.inputs[1].index == 3);
    CHECK(n.inputs.empty() && n.saved.empty());
    m.clear();
    CHECK(m.inputs.empty() && m.saved.empty() && !m.has_update);
  }
  std::printf("%s\n", failures ? "FAILED" : "ok");
  return failures ? 1 : 0;
}
