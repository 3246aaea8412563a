// Host builder of the 8-wide compressed BVH (cwbvh.h) that stands in for the reference's
// `bvh::BVH::build` + flatten (reference src/main.rs:199-244).
//
//   1. per-primitive fp32 boxes (same Bounded rules as src/main.rs:69-82, src/shape.rs:13-20,
//      46-53), inflated by an absolute pad so that the wide tree never culls a primitive the
//      reference's exact-arithmetic test (shapes/triangle.glsl:15-52) would accept;
//   2. binary BVH by binned SAH (16 bins, 3 axes), one primitive per leaf;
//   3. SAH-optimal collapse of the binary tree into 8-wide nodes with <= 3 primitives per leaf
//      (dynamic programme over "forest of at most i roots" costs, node cost 1, primitive 0.3);
//   4. children assigned to slots so that slot ^ ray-octant gives a near-to-far order;
//   5. child boxes quantised to 8 bits per plane, rounded outward; the children of a node are contiguous;
//      the top of the tree is emitted breadth-first, each subtree below it into one contiguous run of
//      nodes and primitives (steps 3 and 5 run on all host threads; the result does not depend on their number).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include "wide_bvh_host.h"

namespace hjk {
namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kNodeCost = 1.0f;
// cost of one primitive test relative to one node step (~70 against ~230 instructions); HJK_BVH_PRIM_COST overrides it
// for tuning sweeps
static const float kPrimCost = [] {
  const char* e = std::getenv("HJK_BVH_PRIM_COST");
  return e ? (float)std::atof(e) : 0.3f;
}();
constexpr int kBins = 16;
#ifndef HJK_BVH_SWEEP_MAX
#define HJK_BVH_SWEEP_MAX 2048
#endif

std::atomic<int> g_builder_threads{0};  // 0 = all
unsigned builder_threads() {
  const unsigned hw = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  const int cap = g_builder_threads.load();
  return cap > 0 ? std::min(hw, (unsigned)cap) : hw;
}

// Runs fn(i) for i in [0, n) on the host's threads (dynamic schedule; results must not depend on it).
template <class Fn>
void parallel_for(size_t n, Fn fn) {
  const size_t nt = std::min<size_t>(builder_threads(), n);
  if (nt <= 1) {
    for (size_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (size_t t = 0; t < nt; t++)
    pool.emplace_back([&]() {
      for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    });
  for (auto& th : pool) th.join();
}

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) {
      lo[k] = kInf;
      hi[k] = -kInf;
    }
  }
  void grow(const Box& b) {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], b.lo[k]);
      hi[k] = std::max(hi[k], b.hi[k]);
    }
  }
  void grow(const float p[3]) {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], p[k]);
      hi[k] = std::max(hi[k], p[k]);
    }
  }
  float half_area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct Node2 {
  Box box;
  uint32_t left = 0, right = 0;  // children (inner)
  uint32_t first = 0, count = 0; // range in the ordered primitive index array (count of the SUBTREE)
  bool leaf = false;
};

struct Binary {
  std::vector<Node2> nodes;
  std::vector<uint32_t> order;  // primitive indices, subtree ranges are contiguous
};

// Binned-SAH binary build.  Nodes are numbered in preorder — a subtree over c primitives owns the
// 2c-1 consecutive indices starting at its root — so subtrees can be built by independent threads
// into disjoint index ranges, and the result does not depend on the thread count.
struct BinaryBuilder {
  const std::vector<Box>& boxes;
  Binary& out;
  uint32_t sweep_max;  // nodes of at most this many primitives get the exact SAH sweep (0: binned everywhere)

  // Occupied bins only: a node over c primitives touches at most c bins per axis, and most of the 2n-1
  // nodes are tiny, so nothing here costs O(kBins) per node (no reset, no sweep over empty bins).
  struct Bins {
    Box bb[3][kBins];
    uint32_t bc[3][kBins];
    uint32_t mask[3];
    void reset() { mask[0] = mask[1] = mask[2] = 0; }
    void add(int a, int b, const Box& box) {
      if ((mask[a] >> b) & 1u) {
        bb[a][b].grow(box);
        bc[a][b]++;
      } else {
        bb[a][b] = box;
        bc[a][b] = 1;
        mask[a] |= 1u << b;
      }
    }
    void merge(const Bins& o) {
      for (int a = 0; a < 3; a++)
        for (uint32_t m = o.mask[a]; m; m &= m - 1) {
          const int b = __builtin_ctz(m);
          if ((mask[a] >> b) & 1u) {
            bb[a][b].grow(o.bb[a][b]);
            bc[a][b] += o.bc[a][b];
          } else {
            bb[a][b] = o.bb[a][b];
            bc[a][b] = o.bc[a][b];
            mask[a] |= 1u << b;
          }
        }
    }
  };
  // Nodes over at least this many primitives spread their two passes over the host's threads (min / max
  // and counts: the result does not depend on the chunking).
  static constexpr uint32_t kParallelNode = 1u << 20, kChunk = 1u << 16;

  // Fills node `ni` (first/count already set); returns false for a leaf, else sets its children.
  bool split(uint32_t ni) {
    const uint32_t first = out.nodes[ni].first, count = out.nodes[ni].count;
    uint32_t* idx = out.order.data() + first;
    // pass 1: bounds of the boxes and of the centroids
    Box nb, cb;
    nb.reset();
    cb.reset();
    auto bounds = [&](uint32_t i0, uint32_t i1, Box& nbox, Box& cbox) {
      for (uint32_t i = i0; i < i1; i++) {
        const Box& pb = boxes[idx[i]];
        const float pc[3] = {0.5f * (pb.lo[0] + pb.hi[0]), 0.5f * (pb.lo[1] + pb.hi[1]), 0.5f * (pb.lo[2] + pb.hi[2])};
        nbox.grow(pb);
        cbox.grow(pc);
      }
    };
    const uint32_t n_chunks = (count + kChunk - 1) / kChunk;
    if (count >= kParallelNode) {
      std::vector<Box> part(2 * (size_t)n_chunks);
      parallel_for(n_chunks, [&](size_t c) {
        part[2 * c].reset(), part[2 * c + 1].reset();
        bounds((uint32_t)c * kChunk, std::min(count, ((uint32_t)c + 1) * kChunk), part[2 * c], part[2 * c + 1]);
      });
      for (uint32_t c = 0; c < n_chunks; c++) nb.grow(part[2 * c]), cb.grow(part[2 * c + 1]);
    } else {
      bounds(0, count, nb, cb);
    }
    out.nodes[ni].box = nb;
    if (count == 1) {
      out.nodes[ni].leaf = true;
      return false;
    }
    // small nodes of small scenes: exact SAH sweep — every split position of the centroid order on each axis, not
    // just the kBins - 1 bin boundaries (cbox: SAH cost 3.95 -> 3.63).  Scenes of millions of primitives keep the
    // binned sweep everywhere: the terrain's cost does not move and its build would take 70 % longer.
    if (count <= sweep_max) {
      thread_local std::vector<uint32_t> order[3];
      thread_local std::vector<float> right_area;
      int best_axis = -1;
      uint32_t best_pos = 0;
      float best_cost = kInf;
      right_area.resize(count);
      for (int axis = 0; axis < 3; axis++) {
        if (!(cb.hi[axis] - cb.lo[axis] > 0.f)) continue;
        std::vector<uint32_t>& o = order[axis];
        o.assign(idx, idx + count);
        std::sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) {
          const float ca = boxes[a].lo[axis] + boxes[a].hi[axis], cb2 = boxes[b].lo[axis] + boxes[b].hi[axis];
          return ca < cb2 || (ca == cb2 && a < b);
        });
        Box acc;
        acc.reset();
        for (uint32_t j = count; j-- > 1;) {
          acc.grow(boxes[o[j]]);
          right_area[j] = acc.half_area();
        }
        acc.reset();
        for (uint32_t j = 0; j + 1 < count; j++) {  // split after position j
          acc.grow(boxes[o[j]]);
          const float cost = acc.half_area() * (float)(j + 1) + right_area[j + 1] * (float)(count - j - 1);
          if (cost < best_cost) best_cost = cost, best_axis = axis, best_pos = j + 1;
        }
      }
      uint32_t mid = count / 2;
      if (best_axis >= 0) {
        std::copy(order[best_axis].begin(), order[best_axis].end(), idx);
        mid = best_pos;
      }
      const uint32_t l = ni + 1, r = ni + 2 * mid;
      out.nodes[ni].left = l;
      out.nodes[ni].right = r;
      out.nodes[l].first = first;
      out.nodes[l].count = mid;
      out.nodes[r].first = first + mid;
      out.nodes[r].count = count - mid;
      return true;
    }
    // pass 2: binned SAH over the three axes, all three binned in one sweep over the primitives
    float c0[3], scale[3];
    bool usable[3];
    for (int axis = 0; axis < 3; axis++) {
      c0[axis] = cb.lo[axis];
      const float ext = cb.hi[axis] - cb.lo[axis];
      usable[axis] = ext > 0.f;
      scale[axis] = usable[axis] ? (float)kBins / ext : 0.f;
    }
    auto fill = [&](uint32_t i0, uint32_t i1, Bins& bins) {
      for (uint32_t i = i0; i < i1; i++) {
        const Box& pb = boxes[idx[i]];
        const float pc[3] = {0.5f * (pb.lo[0] + pb.hi[0]), 0.5f * (pb.lo[1] + pb.hi[1]), 0.5f * (pb.lo[2] + pb.hi[2])};
        for (int axis = 0; axis < 3; axis++) {
          if (!usable[axis]) continue;
          int b = (int)((pc[axis] - c0[axis]) * scale[axis]);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          bins.add(axis, b, pb);
        }
      }
    };
    Bins bins;
    bins.reset();
    if (count >= kParallelNode) {
      std::vector<Bins> part(n_chunks);
      parallel_for(n_chunks, [&](size_t c) {
        part[c].reset();
        fill((uint32_t)c * kChunk, std::min(count, ((uint32_t)c + 1) * kChunk), part[c]);
      });
      for (uint32_t c = 0; c < n_chunks; c++) bins.merge(part[c]);
    } else {
      fill(0, count, bins);
    }
    // A split after bin b puts bins <= b left and bins > b right.  Splits after an empty bin cost the same
    // as the split after the occupied bin below it and never win the strict comparison, so only occupied
    // bins are visited (same winner as a sweep over all kBins - 1 splits, axis 0 first, lowest bin first).
    int best_axis = -1, best_split = -1;
    float best_cost = kInf;
    for (int axis = 0; axis < 3; axis++) {
      if (!usable[axis]) continue;
      const Box* bb = bins.bb[axis];
      const uint32_t* bc = bins.bc[axis];
      int occ[kBins], n_occ = 0;
      for (uint32_t m = bins.mask[axis]; m; m &= m - 1) occ[n_occ++] = __builtin_ctz(m);
      float right_area[kBins];  // union of occupied bins occ[j..]
      uint32_t right_cnt[kBins];
      Box acc;
      acc.reset();
      uint32_t cnt = 0;
      for (int j = n_occ - 1; j > 0; j--) {
        acc.grow(bb[occ[j]]);
        cnt += bc[occ[j]];
        right_area[j] = acc.half_area();
        right_cnt[j] = cnt;
      }
      acc.reset();
      cnt = 0;
      for (int j = 0; j + 1 < n_occ; j++) {
        acc.grow(bb[occ[j]]);
        cnt += bc[occ[j]];
        float cost = acc.half_area() * (float)cnt + right_area[j + 1] * (float)right_cnt[j + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = axis;
          best_split = occ[j];
        }
      }
    }
    uint32_t mid;
    if (best_axis < 0) {
      mid = count / 2;  // coincident centroids: split the list in half
    } else {
      const float pc0 = c0[best_axis], pscale = scale[best_axis];
      uint32_t* m = std::partition(idx, idx + count, [&](uint32_t p) {
        const float pc = 0.5f * (boxes[p].lo[best_axis] + boxes[p].hi[best_axis]);
        int b = (int)((pc - pc0) * pscale);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        return b <= best_split;
      });
      mid = (uint32_t)(m - idx);
      if (mid == 0 || mid == count) mid = count / 2;
    }
    const uint32_t l = ni + 1, r = ni + 2 * mid;  // preorder: the left subtree owns 2*mid-1 nodes
    out.nodes[ni].left = l;
    out.nodes[ni].right = r;
    out.nodes[l].first = first;
    out.nodes[l].count = mid;
    out.nodes[r].first = first + mid;
    out.nodes[r].count = count - mid;
    return true;
  }

  void build_serial(uint32_t root) {
    std::vector<uint32_t> stack{root};
    while (!stack.empty()) {
      const uint32_t ni = stack.back();
      stack.pop_back();
      if (split(ni)) {
        stack.push_back(out.nodes[ni].right);
        stack.push_back(out.nodes[ni].left);
      }
    }
  }

  void build_parallel(uint32_t root, int depth) {
    if (depth >= 5 || out.nodes[root].count < 65536u || builder_threads() == 1) {  // at most 32 concurrent subtrees
      build_serial(root);
      return;
    }
    if (!split(root)) return;
    const uint32_t l = out.nodes[root].left, r = out.nodes[root].right;
    std::thread other([this, l, depth]() { build_parallel(l, depth + 1); });
    build_parallel(r, depth + 1);
    other.join();
  }
};

void build_binary(const std::vector<Box>& boxes, uint32_t sweep_max, Binary& out) {
  const uint32_t n = (uint32_t)boxes.size();
  out.order.resize(n);
  for (uint32_t i = 0; i < n; i++) out.order[i] = i;
  BinaryBuilder bb{boxes, out, sweep_max};
  out.nodes.assign((size_t)2 * n - 1, Node2());
  out.nodes[0].first = 0;
  out.nodes[0].count = n;
  bb.build_parallel(0, 0);
}

// Insertion-based optimisation of a binary tree (Bittner, Hapala, Havran 2013), for small scenes: every node in
// turn — largest boxes first — is cut out with its subtree (its parent disappears, the sibling moves up) and put
// back where it raises the tree's surface-area cost least: next to the node X that minimises
//     area(X + N) + sum over the ancestors A of X of (area(A + N) - area(A)),
// found by a branch-and-bound descent from the root.  The greedy top-down build cannot undo an early split; this
// can.  Deterministic (fixed order, ties by node index); the result is re-linearised into the preorder layout the
// collapse and the emission rely on (a subtree over c primitives owns 2c - 1 consecutive nodes and c consecutive
// entries of `order`).
void optimise_by_reinsertion(const std::vector<Box>& boxes, Binary& bin, int passes) {
  const int n_nodes = (int)bin.nodes.size();
  if (n_nodes < 7) return;
  struct R {
    Box box;
    int parent, left, right;  // left < 0: leaf
    uint32_t prim;
  };
  std::vector<R> t((size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    const Node2& nd = bin.nodes[(size_t)i];
    t[(size_t)i].box = nd.box;
    t[(size_t)i].parent = -1;
    if (nd.leaf) {
      t[(size_t)i].left = t[(size_t)i].right = -1;
      t[(size_t)i].prim = bin.order[nd.first];
    } else {
      t[(size_t)i].left = (int)nd.left, t[(size_t)i].right = (int)nd.right;
      t[(size_t)i].prim = 0;
    }
  }
  for (int i = 0; i < n_nodes; i++)
    if (t[(size_t)i].left >= 0) t[(size_t)t[(size_t)i].left].parent = i, t[(size_t)t[(size_t)i].right].parent = i;
  const int root = 0;
  auto refit_up = [&](int a) {
    for (; a >= 0; a = t[(size_t)a].parent) {
      Box b = t[(size_t)t[(size_t)a].left].box;
      b.grow(t[(size_t)t[(size_t)a].right].box);
      t[(size_t)a].box = b;
    }
  };
  auto merged_area = [](const Box& a, const Box& b) {
    Box u = a;
    u.grow(b);
    return u.half_area();
  };
  std::vector<int> cand;
  std::vector<std::pair<int, float>> stack;
  for (int pass = 0; pass < passes; pass++) {
    cand.clear();
    for (int i = 1; i < n_nodes; i++)
      if (t[(size_t)i].parent != root) cand.push_back(i);  // the root's children stay: their parent cannot vanish
    std::sort(cand.begin(), cand.end(), [&](int a, int b) {
      const float aa = t[(size_t)a].box.half_area(), ab = t[(size_t)b].box.half_area();
      return aa > ab || (aa == ab && a < b);
    });
    bool changed = false;
    for (int nidx : cand) {
      const int P = t[(size_t)nidx].parent;
      if (P < 0 || P == root) continue;  // (earlier moves of this pass may have brought it under the root)
      const int G = t[(size_t)P].parent;
      const int S = t[(size_t)P].left == nidx ? t[(size_t)P].right : t[(size_t)P].left;
      // cut out: S takes P's place under G
      (t[(size_t)G].left == P ? t[(size_t)G].left : t[(size_t)G].right) = S;
      t[(size_t)S].parent = G;
      refit_up(G);
      // best place for nidx in what is left
      const Box& nb = t[(size_t)nidx].box;
      const float n_area = nb.half_area();
      float best = kInf;
      int best_x = -1;
      stack.clear();
      stack.emplace_back(root, 0.f);
      while (!stack.empty()) {
        const auto [x, induced] = stack.back();
        stack.pop_back();
        if (induced + n_area >= best) continue;
        const float direct = merged_area(t[(size_t)x].box, nb);
        const float total = induced + direct;
        if (total < best || (total == best && x < best_x)) best = total, best_x = x;
        if (t[(size_t)x].left >= 0) {
          const float below = induced + direct - t[(size_t)x].box.half_area();
          if (below + n_area < best) {
            stack.emplace_back(t[(size_t)x].right, below);
            stack.emplace_back(t[(size_t)x].left, below);
          }
        }
      }
      if (best_x == root) best_x = S;  // P cannot become the root (node 0 must stay): put the pair back
      // insert: P becomes the parent of best_x and nidx, in best_x's place
      const int X = best_x, XP = t[(size_t)X].parent;
      if (X != S) changed = true;
      (t[(size_t)XP].left == X ? t[(size_t)XP].left : t[(size_t)XP].right) = P;
      t[(size_t)P].parent = XP;
      t[(size_t)P].left = X, t[(size_t)P].right = nidx;
      t[(size_t)X].parent = P, t[(size_t)nidx].parent = P;
      refit_up(P);
    }
    if (!changed) break;
  }
  // back to the preorder layout
  Binary out;
  out.nodes.assign((size_t)n_nodes, Node2());
  out.order.resize(bin.order.size());
  struct Job {
    int src;
    uint32_t dst;
  };
  // subtree sizes first (leaves below each node)
  std::vector<uint32_t> leaves((size_t)n_nodes, 0);
  {
    std::vector<int> order_dfs;
    std::vector<int> st{root};
    while (!st.empty()) {
      const int x = st.back();
      st.pop_back();
      order_dfs.push_back(x);
      if (t[(size_t)x].left >= 0) st.push_back(t[(size_t)x].left), st.push_back(t[(size_t)x].right);
    }
    for (size_t k = order_dfs.size(); k-- > 0;) {
      const int x = order_dfs[k];
      leaves[(size_t)x] = t[(size_t)x].left < 0 ? 1u : leaves[(size_t)t[(size_t)x].left] + leaves[(size_t)t[(size_t)x].right];
    }
  }
  std::vector<std::pair<Job, uint32_t>> st2;  // (source node -> preorder index, first entry of `order`)
  st2.push_back({{root, 0u}, 0u});
  while (!st2.empty()) {
    const auto [job, first] = st2.back();
    st2.pop_back();
    Node2& nd = out.nodes[job.dst];
    nd.box = t[(size_t)job.src].box;
    nd.first = first;
    nd.count = leaves[(size_t)job.src];
    if (t[(size_t)job.src].left < 0) {
      nd.leaf = true;
      out.order[first] = t[(size_t)job.src].prim;
      continue;
    }
    const int l = t[(size_t)job.src].left, r = t[(size_t)job.src].right;
    const uint32_t cl = leaves[(size_t)l];
    nd.left = job.dst + 1;
    nd.right = job.dst + 2 * cl;
    st2.push_back({{r, nd.right}, first + cl});
    st2.push_back({{l, nd.left}, first});
  }
  (void)boxes;
  bin = std::move(out);
}

// Dynamic programme of the wide-tree collapse.  cost[n][i-1] = cheapest way to represent the
// subtree of binary node n as a forest of at most i wide-tree children (i = 1..7).
struct Collapse {  // plain arrays: every entry is written by the sweep, first touched by the thread that fills it
  std::unique_ptr<float[]> cost;      // 7 per node
  std::unique_ptr<uint8_t[]> choice;  // 7 per node: i = 1: 0 leaf / 1 inner; i >= 2: k = roots given to
                                      // the left child, 0 = "same as i-1"
  std::unique_ptr<uint8_t[]> root8;   // per node: left share when the node becomes an inner wide node
  std::unique_ptr<uint8_t[]> has_sphere;  // per node: bit 0 = a sphere, bit 1 = something else among the primitives
                                          // of its subtree
};

void collapse_costs(const Binary& bin, uint32_t n_spheres, Collapse& c) {
  const size_t n = bin.nodes.size();
  c.cost.reset(new float[n * 7]);
  c.choice.reset(new uint8_t[n * 7]);
  c.root8.reset(new uint8_t[n]);
  c.has_sphere.reset(new uint8_t[n]);
  auto process = [&](size_t ni) {
    const Node2& nd = bin.nodes[ni];
    const float area = nd.box.half_area();
    float* cn = &c.cost[ni * 7];
    uint8_t* ch = &c.choice[ni * 7];
    if (nd.leaf) {
      for (int i = 0; i < 7; i++) {
        cn[i] = area * kPrimCost;
        ch[i] = 0;
      }
      c.root8[ni] = 0;
      c.has_sphere[ni] = bin.order[nd.first] < n_spheres ? 1 : 2;  // shape index space: spheres first
      return;
    }
    c.has_sphere[ni] = c.has_sphere[nd.left] | c.has_sphere[nd.right];
    const float* cl = &c.cost[(size_t)nd.left * 7];
    const float* cr = &c.cost[(size_t)nd.right * 7];
    auto distribute = [&](int j, uint8_t& kbest) {  // best split of j roots between the children
      float best = kInf;
      kbest = 1;
      for (int k = 1; k < j; k++) {
        float v = cl[k - 1] + cr[j - k - 1];
        if (v < best) {
          best = v;
          kbest = (uint8_t)k;
        }
      }
      return best;
    };
    uint8_t k8;
    const float inner = distribute(8, k8) + area * kNodeCost;
    c.root8[ni] = k8;
    const float leaf = nd.count <= kWideMaxLeafPrims ? area * kPrimCost * (float)nd.count : kInf;
    if (leaf <= inner) {
      cn[0] = leaf;
      ch[0] = 0;
    } else {
      cn[0] = inner;
      ch[0] = 1;
    }
    for (int i = 2; i <= 7; i++) {
      uint8_t k;
      float d = distribute(i, k);
      if (d < cn[i - 2]) {
        cn[i - 1] = d;
        ch[i - 1] = k;
      } else {
        cn[i - 1] = cn[i - 2];
        ch[i - 1] = 0;
      }
    }
  };
  // Preorder numbering: the subtree over `count` primitives rooted at r owns nodes [r, r + 2 count - 1),
  // children after their parent.  Subtrees below ~n/256 primitives are swept bottom-up by independent
  // threads, the few nodes above them afterwards.
  const uint32_t n_prims = bin.nodes[0].count;
  const uint32_t task_prims = std::max<uint32_t>(4096u, n_prims / 256u);
  std::vector<uint32_t> top, roots, stack{0};
  while (!stack.empty()) {
    const uint32_t ni = stack.back();
    stack.pop_back();
    const Node2& nd = bin.nodes[ni];
    if (nd.leaf || nd.count <= task_prims) {
      roots.push_back(ni);
    } else {
      top.push_back(ni);
      stack.push_back(nd.left);
      stack.push_back(nd.right);
    }
  }
  parallel_for(roots.size(), [&](size_t t) {
    const size_t r = roots[t], end = r + 2 * (size_t)bin.nodes[r].count - 1;
    for (size_t ni = end; ni-- > r;) process(ni);
  });
  std::sort(top.begin(), top.end());
  for (size_t i = top.size(); i-- > 0;) process(top[i]);
}

struct ChildRef {
  uint32_t node2;  // binary node that becomes this wide child
  bool leaf;
};

// expands "binary node n as a forest of at most i roots" into the list of wide children
void gather_children(const Binary& bin, const Collapse& c, uint32_t n, int i, std::vector<ChildRef>& out) {
  struct Item {
    uint32_t n;
    int i;
  };
  std::vector<Item> st{{n, i}};
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    const Node2& nd = bin.nodes[it.n];
    if (nd.leaf) {
      out.push_back({it.n, true});
      continue;
    }
    int ii = it.i;
    while (ii >= 2 && c.choice[(size_t)it.n * 7 + ii - 1] == 0) ii--;
    if (ii == 1) {
      out.push_back({it.n, c.choice[(size_t)it.n * 7] == 0});
      continue;
    }
    const int k = c.choice[(size_t)it.n * 7 + ii - 1];
    st.push_back({nd.right, ii - k});
    st.push_back({nd.left, k});
  }
}

void make_prim(const HjkScene& s, uint32_t shape, WidePrim& p) {
  const uint32_t S = (uint32_t)s.spheres.count, Q = (uint32_t)s.quads.count;
  std::memset(&p, 0, sizeof(p));
  uint32_t id = shape;
  if (shape < S) {
    const HjkSphere& sp = ((const HjkSphere*)s.spheres.ptr)[shape];
    for (int k = 0; k < 3; k++) p.r0[k] = sp.position[k];
    p.r1[0] = sp.radius;
  } else if (shape < S + Q) {
    const HjkQuad& q = ((const HjkQuad*)s.quads.ptr)[shape - S];
    for (int k = 0; k < 3; k++) {
      p.r0[k] = q.origin[k];
      p.r1[k] = q.edge1[k];
      p.r2[k] = q.edge2[k];
    }
  } else {
    const uint32_t* tri = (const uint32_t*)s.triangles.ptr + (size_t)3 * (shape - S - Q);
    const HjkVertex* v = (const HjkVertex*)s.vertices.ptr;
    for (int k = 0; k < 3; k++) {
      // separately rounded fp32 differences, exactly what shapes/triangle.glsl:19-20 computes
      volatile float ab = v[tri[1]].pos[k] - v[tri[0]].pos[k];
      volatile float ac = v[tri[2]].pos[k] - v[tri[0]].pos[k];
      p.r0[k] = v[tri[0]].pos[k];
      p.r1[k] = ab;
      p.r2[k] = ac;
    }
  }
  std::memcpy(&p.r0[3], &id, 4);
#if HJK_PRIM_STRIDE == 4
  if (shape >= S) {  // n = cross(e1, e2), each product and difference rounded to fp32 on its own
    const float* a = p.r1;
    const float* b = p.r2;
    volatile float m0 = a[1] * b[2], m1 = b[1] * a[2], m2 = a[2] * b[0], m3 = b[2] * a[0], m4 = a[0] * b[1],
                   m5 = b[0] * a[1];
    volatile float n0 = m0 - m1, n1 = m2 - m3, n2 = m4 - m5;
    p.r3[0] = n0, p.r3[1] = n1, p.r3[2] = n2;
  }
#endif
}

Box shape_box(const HjkScene& s, uint32_t shape) {
  const uint32_t S = (uint32_t)s.spheres.count, Q = (uint32_t)s.quads.count;
  Box b;
  b.reset();
  if (shape < S) {
    const HjkSphere& sp = ((const HjkSphere*)s.spheres.ptr)[shape];
    const float r = std::fabs(sp.radius);
    for (int k = 0; k < 3; k++) {
      b.lo[k] = sp.position[k] - r;
      b.hi[k] = sp.position[k] + r;
    }
  } else if (shape < S + Q) {
    const HjkQuad& q = ((const HjkQuad*)s.quads.ptr)[shape - S];
    float p[3];
    b.grow(q.origin);
    for (int k = 0; k < 3; k++) p[k] = q.origin[k] + q.edge1[k];
    b.grow(p);
    for (int k = 0; k < 3; k++) p[k] = q.origin[k] + q.edge2[k];
    b.grow(p);
    for (int k = 0; k < 3; k++) p[k] = q.origin[k] + q.edge1[k] + q.edge2[k];
    b.grow(p);
  } else {
    const uint32_t* tri = (const uint32_t*)s.triangles.ptr + (size_t)3 * (shape - S - Q);
    const HjkVertex* v = (const HjkVertex*)s.vertices.ptr;
    for (int c = 0; c < 3; c++) b.grow(v[tri[c]].pos);
  }
  return b;
}

}  // namespace

// bounding ball of the sphere centres + radius range: inputs of the traversal's sphere guard
void sphere_guard_bounds(const HjkScene& s, WideBvh& out) {
  const uint64_t S = s.spheres.count;
  for (int k = 0; k < 4; k++) out.sph_centre[k] = 0.f;
  out.sph_rmin = out.sph_rmax = 0.f;
  if (!S) return;
  const HjkSphere* sp = (const HjkSphere*)s.spheres.ptr;
  Box cb;
  cb.reset();
  float rmin = kInf, rmax = 0.f;
  for (uint64_t i = 0; i < S; i++) {
    cb.grow(sp[i].position);
    rmin = std::min(rmin, std::fabs(sp[i].radius));
    rmax = std::max(rmax, std::fabs(sp[i].radius));
  }
  float rad2 = 0.f;
  for (int k = 0; k < 3; k++) {
    out.sph_centre[k] = 0.5f * (cb.lo[k] + cb.hi[k]);
    const float h = 0.5f * (cb.hi[k] - cb.lo[k]);
    rad2 += h * h;
  }
  out.sph_centre[3] = std::sqrt(rad2) * 1.0001f;
  out.sph_rmin = rmin;
  out.sph_rmax = rmax;
}

void set_builder_threads(int n) { g_builder_threads.store(n < 0 ? 0 : n); }

static bool build_wide_bvh_impl(const HjkScene& s, float pad_rel, WideBvh& out, std::string& err, bool plain) {
  out = WideBvh();
  const uint64_t S = s.spheres.count, Q = s.quads.count, T = s.triangles.count;
  const uint64_t n64 = S + Q + T;
  if (n64 == 0) {
    err = "scene has no shapes";
    return false;
  }
  if (n64 >= 0x7FFFFFFFull) {
    err = "too many shapes for 31-bit shape ids";
    return false;
  }
  const uint32_t n = (uint32_t)n64;
  if (T) {
    const uint32_t* tri = (const uint32_t*)s.triangles.ptr;
    for (uint64_t i = 0; i < 3 * T; i++)
      if (tri[i] >= s.vertices.count) {
        err = "triangle vertex index out of range";
        return false;
      }
  }
  std::vector<Box> boxes(n);
  Box scene;
  scene.reset();
  for (uint32_t i = 0; i < n; i++) {
    boxes[i] = shape_box(s, i);
    for (int k = 0; k < 3; k++)
      if (!std::isfinite(boxes[i].lo[k]) || !std::isfinite(boxes[i].hi[k])) {
        err = "non-finite shape bounds";
        return false;
      }
    scene.grow(boxes[i]);
  }
  float ext = 0.f, mag = 0.f;
  for (int k = 0; k < 3; k++) {
    ext = std::max(ext, scene.hi[k] - scene.lo[k]);
    mag = std::max(mag, std::max(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));
  }
  const float pad = pad_rel * std::max(ext, mag);
  for (uint32_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      boxes[i].lo[k] -= pad;
      boxes[i].hi[k] += pad;
    }
  for (int k = 0; k < 3; k++) {
    out.scene_min[k] = scene.lo[k];
    out.scene_max[k] = scene.hi[k];
  }
  out.pad = pad;

  const bool verbose = std::getenv("HJK_BVH_VERBOSE") != nullptr;
  auto tick = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    auto now = std::chrono::steady_clock::now();
    if (verbose) std::fprintf(stderr, "[hjk bvh] %-18s %.2f s\n", what, std::chrono::duration<double>(now - tick).count());
    tick = now;
  };
  lap("primitive boxes");
  // Small scenes: the top-down build is greedy, so which nodes get the exact sweep changes the tree in ways that are
  // not monotone in quality — build the candidates (milliseconds each) and keep the one whose collapsed wide tree
  // is cheapest.  Scenes of more than kSweepSceneMax primitives keep the binned build everywhere: the terrain's
  // cost does not move with the sweep and its build would take 70 % longer.
  constexpr size_t kSweepSceneMax = 200000;
  std::vector<uint32_t> candidates{0u};
  if (n <= kSweepSceneMax && !plain) candidates = {(uint32_t)HJK_BVH_SWEEP_MAX, 0xFFFFFFFFu, 0u};
  Binary bin;
  Collapse col;
  bool have = false;
  auto consider = [&](Binary& b2) {
    Collapse c2;
    collapse_costs(b2, (uint32_t)s.spheres.count, c2);
    if (!have || c2.cost[0] < col.cost[0]) {
      bin = b2;
      col = std::move(c2);
      have = true;
    }
  };
  for (size_t k = 0; k < candidates.size(); k++) {
    Binary b2;
    build_binary(boxes, candidates[k], b2);
    consider(b2);
  }
  // ... and the winner once more after insertion-based optimisation (HJK_BVH_REINSERT passes; the default of one
  // takes as long as the three builds together and lowers the cost of cbox by another 1 %)
  if (n <= kSweepSceneMax && !plain) {
    static const int passes = [] {
      const char* e = std::getenv("HJK_BVH_REINSERT");
      return e ? std::atoi(e) : 1;
    }();
    if (passes > 0) {
      Binary b2 = bin;
      optimise_by_reinsertion(boxes, b2, passes);
      consider(b2);
    }
  }
  lap("binary SAH build + collapse costs");
  out.sah_cost = col.cost[0] / std::max(bin.nodes[0].box.half_area(), 1e-30f);

  // Emission.  One wide node at a time: gather its (at most 8) children from the collapse choices,
  // assign slots, quantise, append its primitives and reserve the contiguous block of its inner children.
  // `nodes` / `prims` / `queue` are the arrays being appended to (the top of the tree, or one subtree's arena).
  struct Pending {
    uint32_t node2;  // binary node whose subtree this wide node covers
    uint32_t wide;   // index in `nodes`
    uint32_t depth;
  };
  auto emit_one = [&](const Pending cur, std::vector<WideNode>& nodes, std::vector<WidePrim>& prims,
                      std::vector<Pending>& queue, std::vector<ChildRef>& kids) -> bool {
    const Node2& nd = bin.nodes[cur.node2];
    kids.clear();
    if (nd.leaf) {
      kids.push_back({cur.node2, true});  // single-primitive scene: root with one leaf child
    } else {
      const int k = col.root8[cur.node2];
      gather_children(bin, col, nd.left, k, kids);
      gather_children(bin, col, nd.right, 8 - k, kids);
    }
    const int nk = (int)kids.size();
    // ---- slot assignment: greedy maximisation of sum dot(child centre - node centre, slot signs)
    Box nb;
    nb.reset();
    for (int c = 0; c < nk; c++) nb.grow(bin.nodes[kids[c].node2].box);
    float score[8][8];
    for (int c = 0; c < nk; c++) {
      const Box& cb = bin.nodes[kids[c].node2].box;
      float d[3];
      for (int k = 0; k < 3; k++) d[k] = 0.5f * (cb.lo[k] + cb.hi[k]) - 0.5f * (nb.lo[k] + nb.hi[k]);
      for (int sl = 0; sl < 8; sl++)
        score[c][sl] = ((sl & 1) ? d[0] : -d[0]) + ((sl & 2) ? d[1] : -d[1]) + ((sl & 4) ? d[2] : -d[2]);
    }
    int slot_of[8];
    bool child_done[8] = {false}, slot_used[8] = {false};
    for (int it = 0; it < nk; it++) {
      int bc = -1, bs = -1;
      float best = -kInf;
      for (int c = 0; c < nk; c++) {
        if (child_done[c]) continue;
        for (int sl = 0; sl < 8; sl++) {
          if (slot_used[sl]) continue;
          if (score[c][sl] > best) {
            best = score[c][sl];
            bc = c;
            bs = sl;
          }
        }
      }
      slot_of[bc] = bs;
      child_done[bc] = true;
      slot_used[bs] = true;
    }
    int child_in_slot[8];
    for (int sl = 0; sl < 8; sl++) child_in_slot[sl] = -1;
    for (int c = 0; c < nk; c++) child_in_slot[slot_of[c]] = c;

    // ---- quantisation frame
    WideNode wn;
    std::memset(&wn, 0, sizeof(wn));
    double scale[3];
    for (int k = 0; k < 3; k++) {
      wn.origin[k] = nb.lo[k];
      const double extent = (double)nb.hi[k] - (double)nb.lo[k];
      int e = extent > 0.0 ? (int)std::ceil(std::log2(extent / 255.0)) : -126;
      e = std::max(e, -126);
      while (std::ceil(extent / std::ldexp(1.0, e)) > 255.0) e++;
      if (e > 127) {
        return false;
      }
      wn.e[k] = (uint8_t)(e + 127);
      scale[k] = std::ldexp(1.0, e);
    }
    wn.child_base = (uint32_t)nodes.size();
    wn.prim_base = (uint32_t)prims.size() | ((col.has_sphere[cur.node2] & 1) ? kWideHasSpheres : 0u) |
                   (col.has_sphere[cur.node2] == 1 ? kWideOnlySpheres : 0u);
    uint32_t prim_off = 0, n_inner = 0;
    for (int sl = 0; sl < 8; sl++) {
      const int c = child_in_slot[sl];
      if (c < 0) {  // empty: inverted box never passes the slab test
        for (int k = 0; k < 3; k++) {
          wn.qlo[k][sl] = 255;
          wn.qhi[k][sl] = 0;
        }
        continue;
      }
      const Node2& cn = bin.nodes[kids[c].node2];
      for (int k = 0; k < 3; k++) {
        double lo = std::floor(((double)cn.box.lo[k] - (double)wn.origin[k]) / scale[k]);
        double hi = std::ceil(((double)cn.box.hi[k] - (double)wn.origin[k]) / scale[k]);
        lo = std::min(std::max(lo, 0.0), 255.0);
        hi = std::min(std::max(hi, 0.0), 255.0);
        wn.qlo[k][sl] = (uint8_t)lo;
        wn.qhi[k][sl] = (uint8_t)hi;
      }
      if (kids[c].leaf) {
        const uint32_t cnt = cn.count;  // 1..3 primitives of this binary subtree
        wn.meta[sl] = (uint8_t)((((1u << cnt) - 1u) << 5) | prim_off);
        for (uint32_t i = 0; i < cnt; i++) {
          WidePrim wp;
          make_prim(s, bin.order[cn.first + i], wp);
          prims.push_back(wp);
        }
        prim_off += cnt;
      } else {
        wn.meta[sl] = (uint8_t)((1u << 5) | (24u + (uint32_t)sl));
        wn.imask |= (uint8_t)(1u << sl);
        n_inner++;
      }
    }
    // inner children are stored in slot order starting at child_base
    for (int sl = 0; sl < 8; sl++) {
      const int c = child_in_slot[sl];
      if (c < 0 || kids[c].leaf) continue;
      queue.push_back({kids[c].node2, (uint32_t)nodes.size(), cur.depth + 1});
      nodes.emplace_back();
    }
    (void)n_inner;
    nodes[cur.wide] = wn;
      return true;
  };

  // The top of the tree is emitted breadth-first by one thread until enough subtrees are pending; each of
  // those is then emitted into its own arena by the host's threads, and the arenas are appended in task
  // order (indices rebased), so the layout does not depend on the thread count.
  std::vector<Pending> queue;
  std::vector<ChildRef> kids;
  out.nodes.emplace_back();
  queue.push_back({0, 0, 1});
  size_t qi = 0;
  const size_t want_tasks = 2048;
  for (; qi < queue.size() && queue.size() - qi < want_tasks; qi++) {
    out.depth = std::max(out.depth, queue[qi].depth);
    if (!emit_one(queue[qi], out.nodes, out.prims, queue, kids)) {
      err = "scene extent too large to quantise";
      return false;
    }
  }
  struct Arena {
    std::vector<WideNode> nodes;  // [0] = the task's root (it lives at Pending::wide of the top array)
    std::vector<WidePrim> prims;
    uint32_t depth = 0;
    bool ok = true;
  };
  const size_t n_tasks = queue.size() - qi;
  std::vector<Arena> arenas(n_tasks);
  parallel_for(n_tasks, [&](size_t t) {
    Arena& ar = arenas[t];
    std::vector<Pending> q;
    std::vector<ChildRef> k;
    ar.nodes.emplace_back();
    q.push_back({queue[qi + t].node2, 0, queue[qi + t].depth});
    for (size_t i = 0; i < q.size() && ar.ok; i++) {
      ar.depth = std::max(ar.depth, q[i].depth);
      ar.ok = emit_one(q[i], ar.nodes, ar.prims, q, k);
    }
  });
  std::vector<uint32_t> node_off(n_tasks + 1), prim_off(n_tasks + 1);
  node_off[0] = (uint32_t)out.nodes.size(), prim_off[0] = (uint32_t)out.prims.size();
  for (size_t t = 0; t < n_tasks; t++) {
    if (!arenas[t].ok) {
      err = "scene extent too large to quantise";
      return false;
    }
    out.depth = std::max(out.depth, arenas[t].depth);
    node_off[t + 1] = node_off[t] + (uint32_t)arenas[t].nodes.size() - 1u;
    prim_off[t + 1] = prim_off[t] + (uint32_t)arenas[t].prims.size();
  }
  out.nodes.resize(node_off[n_tasks]);
  out.prims.resize(prim_off[n_tasks]);
  parallel_for(n_tasks, [&](size_t t) {
    Arena& ar = arenas[t];
    // arena node i >= 1 moves to node_off + i - 1, arena primitive j to prim_off + j
    for (WideNode& wn : ar.nodes) {
      wn.child_base = node_off[t] + wn.child_base - 1u;
      wn.prim_base += prim_off[t];
    }
    out.nodes[queue[qi + t].wide] = ar.nodes[0];
    std::copy(ar.nodes.begin() + 1, ar.nodes.end(), out.nodes.begin() + node_off[t]);
    std::copy(ar.prims.begin(), ar.prims.end(), out.prims.begin() + prim_off[t]);
    std::vector<WideNode>().swap(ar.nodes);
    std::vector<WidePrim>().swap(ar.prims);
  });
  lap("emit wide nodes");
  out.n_shapes = n;
  sphere_guard_bounds(s, out);
  return true;
}

// Structural check used by the not-gpu tests: every shape is referenced exactly once, every
// child box (decoded the way the traversal decodes it) contains the boxes below it.

// The exact sweep can peel an adversarial scene (sizes falling geometrically towards a point) one primitive per
// split: a tree deeper than the traversal supports (32 levels) is rebuilt with the binned splits only, which halve
// such a scene's extent per level.
bool build_wide_bvh(const HjkScene& s, float pad_rel, WideBvh& out, std::string& err) {
  if (!build_wide_bvh_impl(s, pad_rel, out, err, false)) return false;
  if (out.depth > 32u) return build_wide_bvh_impl(s, pad_rel, out, err, true);
  return true;
}

bool validate_wide_bvh(const HjkScene& s, const WideBvh& bvh, std::string& err) {
  const uint32_t n = bvh.n_shapes;
  std::vector<uint8_t> seen(n, 0);
  struct Item {
    uint32_t node;
    float lo[3], hi[3];
    bool has_box;
  };
  std::vector<Item> st;
  st.push_back({0, {0, 0, 0}, {0, 0, 0}, false});
  uint64_t visited = 0;
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    if (it.node >= bvh.nodes.size()) {
      err = "child index out of range";
      return false;
    }
    visited++;
    const WideNode& wn = bvh.nodes[it.node];
    uint32_t inner_rank = 0;
    for (int sl = 0; sl < 8; sl++) {
      const uint8_t m = wn.meta[sl];
      if (m == 0) continue;
      float lo[3], hi[3];
      for (int k = 0; k < 3; k++) {
        const float sc = std::ldexp(1.0f, (int)wn.e[k] - 127);
        lo[k] = wn.origin[k] + (float)wn.qlo[k][sl] * sc;
        hi[k] = wn.origin[k] + (float)wn.qhi[k][sl] * sc;
        if (it.has_box && (lo[k] < it.lo[k] - 1e-3f * std::fabs(it.lo[k]) - 1e-6f ||
                           hi[k] > it.hi[k] + 1e-3f * std::fabs(it.hi[k]) + 1e-6f)) {
          // a child may exceed its parent's QUANTISED box only by rounding; flag gross errors
          const float tol = 2.f * sc + 1e-5f;
          if (lo[k] < it.lo[k] - tol || hi[k] > it.hi[k] + tol) {
            err = "child box escapes its parent";
            return false;
          }
        }
      }
      const bool inner = (m & 0x18) == 0x18 && (m >> 5) == 1;
      if (inner) {
        if (!((wn.imask >> sl) & 1)) {
          err = "imask disagrees with meta";
          return false;
        }
        Item ch;
        ch.node = wn.child_base + inner_rank++;
        ch.has_box = true;
        for (int k = 0; k < 3; k++) {
          ch.lo[k] = lo[k];
          ch.hi[k] = hi[k];
        }
        st.push_back(ch);
      } else {
        const uint32_t bits = m >> 5, off = m & 31u;
        const uint32_t cnt = bits == 1 ? 1 : (bits == 3 ? 2 : (bits == 7 ? 3 : 0));
        if (cnt == 0 || off + cnt > kWideMaxNodePrims) {
          err = "bad leaf meta";
          return false;
        }
        for (uint32_t i = 0; i < cnt; i++) {
          const size_t pi = (size_t)(wn.prim_base & kWidePrimBaseMask) + off + i;
          if (pi >= bvh.prims.size()) {
            err = "primitive index out of range";
            return false;
          }
          uint32_t id;
          std::memcpy(&id, &bvh.prims[pi].r0[3], 4);
          if (id >= n || seen[id]) {
            err = "shape referenced twice or out of range";
            return false;
          }
          seen[id] = 1;
          Box b = shape_box(s, id);
          for (int k = 0; k < 3; k++)
            if (b.lo[k] < lo[k] || b.hi[k] > hi[k]) {
              err = "leaf box does not contain its primitive";
              return false;
            }
        }
      }
    }
  }
  for (uint32_t i = 0; i < n; i++)
    if (!seen[i]) {
      err = "shape missing from the tree";
      return false;
    }
  if (visited != bvh.nodes.size()) {
    err = "unreachable nodes";
    return false;
  }
  return true;
}

}  // namespace hjk
