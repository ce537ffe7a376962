// TEST INFRASTRUCTURE (oracle). GL-free stand-in for the reference's
// source/vertexpool.h (OpenGL persistent-mapped vertex pool, out of scope).
// It keeps the two global using-directives that file provides
// (vertexpool.h:4 `using namespace glm;`, :65 `using namespace std;`) because
// cellpool.h / world.h / water.h only compile in their presence, and it offers
// the members cellpool.h:252-336 calls (section/resize/index/update/fill and
// the `indices` vector).  The vertex records are captured so that the N1 row
// (updatenode vertex fill) can be checked against the reference as well.
#pragma once
#include <sys/types.h>
#include <algorithm>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <vector>

using namespace glm;

struct Vertex {
  float position[3], normal[3], tangent[3], bitangent[3];
  Vertex() : position{0, 0, 0}, normal{0, 0, 0}, tangent{0, 0, 0}, bitangent{0, 0, 0} {}
  Vertex(vec3 p, vec3 n, vec3 t, vec3 b)
      : position{p.x, p.y, p.z}, normal{n.x, n.y, n.z}, tangent{t.x, t.y, t.z}, bitangent{b.x, b.y, b.z} {}
};

using namespace std;

template <typename T>
struct Vertexpool {
  vector<unsigned> indices;
  vector<T> store;      // captured vertices, section-major
  vector<uint> starts;  // one entry per section
  size_t per_section = 0;

  Vertexpool() { starts.reserve(4096); }  // node.vertex pointers must stay valid

  void reserve(const int k, const int n) {
    per_section = (size_t)k;
    store.assign((size_t)k * (size_t)n, T());
  }
  uint* section(const int size, const int = 0, vec3 = vec3(0), const int = 0) {
    if (per_section == 0) per_section = (size_t)size;
    starts.push_back((uint)(starts.size() * per_section));
    if (store.size() < starts.size() * per_section) store.resize(starts.size() * per_section);
    return &starts.back();
  }
  void resize(const uint*, const int) {}
  void index() {}
  void update() {}
  template <typename... A>
  void fill(uint* sec, const int k, A&&... a) {
    store[(size_t)(*sec) + (size_t)k] = T(a...);
  }
};
