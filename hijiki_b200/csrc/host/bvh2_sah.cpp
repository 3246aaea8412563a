// Reference-layout binary BVH (the optional `bvh` binding).
//
// The reference calls `bvh::bvh::BVH::build` (crate bvh 0.3.1 over nalgebra 0.17.3; source
// NOT in the reference tree — src/main.rs:199) and flattens the result to a preorder
// skip-pointer array (src/main.rs:203-244).  This file restates that crate's published
// build rule in our own code:
//   * one shape per leaf;
//   * split axis = largest axis of the CENTROID bounds;
//   * if that extent < 1e-5: split the index list in half;
//   * else 6 equal-width centroid buckets, bucket = floor(rel * (6 - 0.01)), and the
//     5 candidate splits are scored (n_l*A_l + n_r*A_r) / A_parent; first minimum wins;
//   * children keep bucket order (stable).
// Topology only influences tie outcomes — parity unpinned, see DESIGN.md.
#include <cmath>
#include <limits>

#include "host_scene.h"

namespace hjk {
namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

Aabb aabb_empty() { return Aabb{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}}; }
Aabb aabb_join(const Aabb& a, const Aabb& b) {
  Aabb r;
  for (int k = 0; k < 3; k++) {
    r.min[k] = std::fmin(a.min[k], b.min[k]);
    r.max[k] = std::fmax(a.max[k], b.max[k]);
  }
  return r;
}
Aabb aabb_grow(const Aabb& a, const float p[3]) {
  Aabb r;
  for (int k = 0; k < 3; k++) {
    r.min[k] = std::fmin(a.min[k], p[k]);
    r.max[k] = std::fmax(a.max[k], p[k]);
  }
  return r;
}
void aabb_center(const Aabb& a, float c[3]) {
  for (int k = 0; k < 3; k++) c[k] = a.min[k] + (a.max[k] - a.min[k]) / 2.0f;
}
float aabb_area(const Aabb& a) {
  float sx = a.max[0] - a.min[0], sy = a.max[1] - a.min[1], sz = a.max[2] - a.min[2];
  return 2.0f * (sx * sy + sx * sz + sy * sz);
}
int aabb_largest_axis(const Aabb& a) {
  float sx = a.max[0] - a.min[0], sy = a.max[1] - a.min[1], sz = a.max[2] - a.min[2];
  if (sx > sy && sx > sz) return 0;
  if (sy > sz) return 1;
  return 2;
}

struct Node2 {
  bool leaf = false;
  uint32_t shape = 0;
  uint32_t child_l = 0, child_r = 0;
  Aabb aabb_l{}, aabb_r{};
};

struct Builder {
  const std::vector<Aabb>& shapes;
  std::vector<Node2> nodes;

  uint32_t build(const std::vector<uint32_t>& indices) {
    Aabb bounds = aabb_empty(), cbounds = aabb_empty();
    for (uint32_t i : indices) {
      float c[3];
      aabb_center(shapes[i], c);
      bounds = aabb_join(bounds, shapes[i]);
      cbounds = aabb_grow(cbounds, c);
    }
    if (indices.size() == 1) {
      Node2 n;
      n.leaf = true;
      n.shape = indices[0];
      nodes.push_back(n);
      return (uint32_t)nodes.size() - 1;
    }
    uint32_t me = (uint32_t)nodes.size();
    nodes.emplace_back();

    int axis = aabb_largest_axis(cbounds);
    float axis_size = cbounds.max[axis] - cbounds.min[axis];
    std::vector<uint32_t> left, right;
    Aabb la = aabb_empty(), ra = aabb_empty();
    if (axis_size < 0.00001f) {
      size_t half = indices.size() / 2;
      left.assign(indices.begin(), indices.begin() + half);
      right.assign(indices.begin() + half, indices.end());
      for (uint32_t i : left) la = aabb_join(la, shapes[i]);
      for (uint32_t i : right) ra = aabb_join(ra, shapes[i]);
    } else {
      constexpr int NB = 6;
      Aabb baabb[NB];
      size_t bsize[NB];
      std::vector<uint32_t> assign[NB];
      for (int b = 0; b < NB; b++) {
        baabb[b] = aabb_empty();
        bsize[b] = 0;
      }
      for (uint32_t i : indices) {
        float c[3];
        aabb_center(shapes[i], c);
        float rel = (c[axis] - cbounds.min[axis]) / axis_size;
        int b = (int)(rel * ((float)NB - 0.01f));
        if (b < 0) b = 0;
        if (b >= NB) b = NB - 1;
        baabb[b] = aabb_join(baabb[b], shapes[i]);
        bsize[b]++;
        assign[b].push_back(i);
      }
      int min_bucket = 0;
      float min_cost = kInf;
      float parent_area = aabb_area(bounds);
      for (int s = 0; s < NB - 1; s++) {
        Aabb l = aabb_empty(), r = aabb_empty();
        size_t nl = 0, nr = 0;
        for (int b = 0; b <= s; b++) {
          l = aabb_join(l, baabb[b]);
          nl += bsize[b];
        }
        for (int b = s + 1; b < NB; b++) {
          r = aabb_join(r, baabb[b]);
          nr += bsize[b];
        }
        float cost = ((float)nl * aabb_area(l) + (float)nr * aabb_area(r)) / parent_area;
        if (cost < min_cost) {
          min_bucket = s;
          min_cost = cost;
          la = l;
          ra = r;
        }
      }
      for (int b = 0; b <= min_bucket; b++) left.insert(left.end(), assign[b].begin(), assign[b].end());
      for (int b = min_bucket + 1; b < NB; b++)
        right.insert(right.end(), assign[b].begin(), assign[b].end());
      if (left.empty() || right.empty()) {  // all costs NaN/inf: fall back to the median split
        size_t half = indices.size() / 2;
        left.assign(indices.begin(), indices.begin() + half);
        right.assign(indices.begin() + half, indices.end());
        la = aabb_empty();
        ra = aabb_empty();
        for (uint32_t i : left) la = aabb_join(la, shapes[i]);
        for (uint32_t i : right) ra = aabb_join(ra, shapes[i]);
      }
    }
    uint32_t l = build(left);
    uint32_t r = build(right);
    Node2& n = nodes[me];
    n.child_l = l;
    n.child_r = r;
    n.aabb_l = la;
    n.aabb_r = ra;
    return me;
  }
};

}  // namespace

void build_flat_bvh2(const std::vector<Aabb>& shape_aabbs, std::vector<HjkBvh2Node>& flat) {
  flat.clear();
  if (shape_aabbs.empty()) return;
  Builder b{shape_aabbs, {}};
  std::vector<uint32_t> all(shape_aabbs.size());
  for (size_t i = 0; i < all.size(); i++) all[i] = (uint32_t)i;
  b.build(all);
  const std::vector<Node2>& nodes = b.nodes;

  // preorder numbering (calculate_indices, src/main.rs:203-213) — iterative to spare the stack
  std::vector<uint32_t> pre(nodes.size(), 0xFFFFFFFFu);
  {
    uint32_t counter = 0;
    std::vector<uint32_t> st{0};
    while (!st.empty()) {
      uint32_t cur = st.back();
      st.pop_back();
      pre[cur] = counter++;
      if (!nodes[cur].leaf) {
        st.push_back(nodes[cur].child_r);
        st.push_back(nodes[cur].child_l);
      }
    }
  }
  // flatten with skip pointers (flatten_bvh, src/main.rs:214-231); root exit = 1000000
  struct Item {
    uint32_t node;
    Aabb aabb;
    uint32_t skip;
  };
  Aabb root;
  if (nodes[0].leaf) {
    root = shape_aabbs[nodes[0].shape];  // single-shape scene: the reference would panic here
  } else {
    root = aabb_join(nodes[0].aabb_l, nodes[0].aabb_r);
  }
  std::vector<Item> st{{0, root, 1000000u}};
  flat.reserve(nodes.size());
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    const Node2& n = nodes[it.node];
    HjkBvh2Node out;
    for (int k = 0; k < 3; k++) {
      out.aabb_min[k] = it.aabb.min[k];
      out.aabb_max[k] = it.aabb.max[k];
    }
    out.shape_index = n.leaf ? n.shape : 0xFFFFFFFFu;
    out.exit_index = it.skip;
    flat.push_back(out);
    if (!n.leaf) {
      st.push_back({n.child_r, n.aabb_r, it.skip});
      st.push_back({n.child_l, n.aabb_l, pre[n.child_r]});
    }
  }
}

}  // namespace hjk
