// GPU builder of the 8-wide compressed BVH (cwbvh.h) — SURVEY.md §8(f)-1: stands in for the
// reference's host-side `bvh::BVH::build` (src/main.rs:199) when setup time matters (10 M triangles:
// tens of milliseconds instead of seconds on the host).
//
//   1. k_shape_boxes   per-shape fp32 boxes (same Bounded rules as host/cwbvh_build.cpp), scene bounds
//   2. k_morton        outward pad, 63-bit Morton code of the box centre
//   3. cub::DeviceRadixSort::SortPairs
//   4. k_ploc_*        binary tree by locally-ordered clustering over the Morton order (PLOC, radius 16): SAH-grade
//                      (option "bvh_gpu_tree" = 0 selects the older pair instead: k_radix_tree, the binary radix
//                      tree over the sorted codes (Karras 2012), + k_fit_boxes, bottom-up box fitting)
//   6. k_collapse      level by level: a wide node pulls up to 8 children out of its binary subtree by
//                      repeatedly opening the child of largest area; subtrees of <= 3 primitives become
//                      leaves; slots by octant order; 8-bit quantisation rounded outward; primitive
//                      records written with the reference's separately rounded differences
//
// Select with hjk_set_option("bvh_builder", 1); by default scenes of more than a million shapes use it (the host
// SAH builder needs seconds there).  Hit results do not depend on the tree (ties excepted), so every parity test
// also runs against this builder.
#pragma once
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../cwbvh.h"
#include "scene_dev.cuh"

namespace hjk {
namespace gpubvh {

struct BuildScene {
  const f4* spheres;
  const f4* quads;
  const uint32_t* triangles;
  const f4* vertices;
  uint32_t S, Q, T;
};

constexpr uint32_t kLeafFlag = 0x80000000u;
// primitive counts of subtrees carry two flag bits on top: a sphere / something other than a sphere lives below
constexpr uint32_t kCountMask = 0x3FFFFFFFu, kKindSphere = 0x40000000u, kKindOther = 0x80000000u;
__device__ __forceinline__ uint32_t merge_counts(uint32_t a, uint32_t b) {
  return ((a & kCountMask) + (b & kCountMask)) | ((a | b) & ~kCountMask);
}

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ void shape_box(const BuildScene& s, uint32_t shape, float lo[3], float hi[3]) {
  if (shape < s.S) {  // src/shape.rs:13-20
    const f4 sp = s.spheres[shape];
    const float r = fabsf(sp.w);
    lo[0] = sp.x - r, lo[1] = sp.y - r, lo[2] = sp.z - r;
    hi[0] = sp.x + r, hi[1] = sp.y + r, hi[2] = sp.z + r;
  } else if (shape < s.S + s.Q) {  // src/shape.rs:46-53
    const f4* q = s.quads + 3 * (size_t)(shape - s.S);
    const f4 o = q[0], e1 = q[1], e2 = q[2];
    const float ox[3] = {o.x, o.y, o.z}, a[3] = {e1.x, e1.y, e1.z}, b[3] = {e2.x, e2.y, e2.z};
    for (int k = 0; k < 3; k++) {
      const float p0 = ox[k], p1 = __fadd_rn(ox[k], a[k]), p2 = __fadd_rn(ox[k], b[k]),
                  p3 = __fadd_rn(__fadd_rn(ox[k], a[k]), b[k]);
      lo[k] = fminf(fminf(p0, p1), fminf(p2, p3));
      hi[k] = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
    }
  } else {  // src/main.rs:74-79
    const uint32_t* t = s.triangles + 3 * (size_t)(shape - s.S - s.Q);
    const f4 a = s.vertices[2 * (size_t)t[0]], b = s.vertices[2 * (size_t)t[1]], c = s.vertices[2 * (size_t)t[2]];
    lo[0] = fminf(a.x, fminf(b.x, c.x)), lo[1] = fminf(a.y, fminf(b.y, c.y)), lo[2] = fminf(a.z, fminf(b.z, c.z));
    hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)), hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)), hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
  }
}

// bounds: 6 ordered uints (min xyz, max xyz), initialised to 0xFFFFFFFF x3, 0 x3
__global__ void k_shape_boxes(BuildScene s, uint32_t n, f4* blo, f4* bhi, uint32_t* bounds, uint32_t* bad) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float lo[3], hi[3];
    shape_box(s, i, lo, hi);
    blo[i] = F4(lo[0], lo[1], lo[2], 0.f);
    bhi[i] = F4(hi[0], hi[1], hi[2], 0.f);
    for (int k = 0; k < 3; k++) {
      if (!isfinite(lo[k]) || !isfinite(hi[k])) atomicExch(bad, 1u);
      mn[k] = fminf(mn[k], lo[k]);
      mx[k] = fmaxf(mx[k], hi[k]);
    }
  }
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], o));
    }
    if ((threadIdx.x & 31) == 0 && mn[k] <= mx[k]) {
      atomicMin(bounds + k, float_to_ordered(mn[k]));
      atomicMax(bounds + 3 + k, float_to_ordered(mx[k]));
    }
  }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1FFFFFu;
  x = (x | (x << 32)) & 0x1F00000000FFFFull;
  x = (x | (x << 16)) & 0x1F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void k_morton(uint32_t n, const uint32_t* bounds, float pad_rel, f4* blo, f4* bhi, uint64_t* keys,
                         uint32_t* vals, float* pad_out) {
  float smin[3], smax[3], ext = 0.f, mag = 0.f;
  for (int k = 0; k < 3; k++) {
    smin[k] = ordered_to_float(bounds[k]);
    smax[k] = ordered_to_float(bounds[3 + k]);
    ext = fmaxf(ext, smax[k] - smin[k]);
    mag = fmaxf(mag, fmaxf(fabsf(smin[k]), fabsf(smax[k])));
  }
  const float pad = pad_rel * fmaxf(ext, mag);
  if (blockIdx.x == 0 && threadIdx.x == 0) *pad_out = pad;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    f4 lo = blo[i], hi = bhi[i];
    lo.x -= pad, lo.y -= pad, lo.z -= pad;
    hi.x += pad, hi.y += pad, hi.z += pad;
    blo[i] = lo;
    bhi[i] = hi;
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t q[3];
    for (int k = 0; k < 3; k++) {
      const float e = smax[k] - smin[k];
      float u = e > 0.f ? (c[k] - smin[k]) / e : 0.5f;
      u = fminf(fmaxf(u, 0.f), 1.f);
      q[k] = (uint32_t)fminf(u * 2097152.0f, 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
  }
}

// common-prefix length of sorted keys i and j (ties broken by position), -1 outside [0, n)
__device__ __forceinline__ int delta(const uint64_t* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
  return __clzll((long long)(a ^ b));
}

// Karras 2012: inner node i covers a range of sorted leaves; children carry kLeafFlag when leaves.
__global__ void k_radix_tree(int n, const uint64_t* keys, uint32_t* child_l, uint32_t* child_r, uint32_t* parent_inner,
                             uint32_t* parent_leaf) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
      if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
      t = (t + 1) >> 1;
      if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t cl = (lo == gamma) ? ((uint32_t)gamma | kLeafFlag) : (uint32_t)gamma;
    const uint32_t cr = (hi == gamma + 1) ? ((uint32_t)(gamma + 1) | kLeafFlag) : (uint32_t)(gamma + 1);
    child_l[i] = cl;
    child_r[i] = cr;
    if (cl & kLeafFlag) parent_leaf[gamma] = (uint32_t)i; else parent_inner[gamma] = (uint32_t)i;
    if (cr & kLeafFlag) parent_leaf[gamma + 1] = (uint32_t)i; else parent_inner[gamma + 1] = (uint32_t)i;
    if (i == 0) parent_inner[0] = 0xFFFFFFFFu;
  }
}

// One thread per leaf walks up; the second thread to reach an inner node fits its box and goes on.
__global__ void k_fit_boxes(int n, uint32_t n_spheres, const uint32_t* vals, const f4* blo, const f4* bhi, const uint32_t* child_l,
                            const uint32_t* child_r, const uint32_t* parent_inner, const uint32_t* parent_leaf,
                            uint32_t* visits, f4* ilo, f4* ihi, uint32_t* icount) {
  for (int leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n; leaf += gridDim.x * blockDim.x) {
    uint32_t node = parent_leaf[leaf];
    while (node != 0xFFFFFFFFu) {
      __threadfence();
      if (atomicAdd(visits + node, 1u) == 0u) break;  // first arriver: the sibling subtree is not done yet
      __threadfence();
      const uint32_t cl = child_l[node], cr = child_r[node];
      f4 llo, lhi, rlo, rhi;
      uint32_t lc, rc;
      // inner children were written by other threads: read them around the caches (volatile), after
      // the fence that follows the atomic
      auto load_inner = [&](uint32_t r, f4& o_lo, f4& o_hi, uint32_t& o_cnt) {
        const volatile float* pl = reinterpret_cast<const volatile float*>(ilo + r);
        const volatile float* ph = reinterpret_cast<const volatile float*>(ihi + r);
        o_lo = F4(pl[0], pl[1], pl[2], 0.f);
        o_hi = F4(ph[0], ph[1], ph[2], 0.f);
        o_cnt = *reinterpret_cast<const volatile uint32_t*>(icount + r);
      };
      if (cl & kLeafFlag) {
        const uint32_t p = vals[cl & ~kLeafFlag];
        llo = blo[p], lhi = bhi[p], lc = 1u | (p < n_spheres ? kKindSphere : kKindOther);
      } else {
        load_inner(cl, llo, lhi, lc);
      }
      if (cr & kLeafFlag) {
        const uint32_t p = vals[cr & ~kLeafFlag];
        rlo = blo[p], rhi = bhi[p], rc = 1u | (p < n_spheres ? kKindSphere : kKindOther);
      } else {
        load_inner(cr, rlo, rhi, rc);
      }
      ilo[node] = F4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
      ihi[node] = F4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
      icount[node] = merge_counts(lc, rc);
      node = parent_inner[node];
    }
  }
}

// ---------------------------------------------------------------- PLOC (parallel locally-ordered clustering)
// Binary tree by bottom-up agglomeration over the Morton order (Meister & Bittner 2018): every cluster looks for
// the neighbour within kPlocRadius positions whose union with it has the smallest surface area; mutual nearest
// neighbours merge; the survivors are compacted (order kept) and the search repeats until one cluster is left.
// The trees are SAH-grade — unlike the radix tree, whose splits follow the bits of the codes — and the inner
// nodes come out with their boxes and primitive counts, so no fitting pass is needed.  Inner node ids are handed
// out so that the root, created last, is node 0 (what k_collapse starts from); everything is deterministic.
constexpr int kPlocRadius = 16, kPlocMaxRadius = 64;  // default and largest search radius (option "bvh_ploc_radius")
constexpr int kPlocThreads = 256;

__global__ void k_ploc_init(uint32_t n, uint32_t n_spheres, const uint32_t* vals, const f4* blo, const f4* bhi,
                            uint32_t* ref, f4* clo, f4* chi, uint32_t* ccnt) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = vals[i];
    ref[i] = i | kLeafFlag;
    clo[i] = blo[p];
    chi[i] = bhi[p];
    ccnt[i] = 1u | (p < n_spheres ? kKindSphere : kKindOther);
  }
}

// nn[i] = the cluster within `radius` positions of i whose union with i has the smallest area (ties: lower index)
__global__ void __launch_bounds__(kPlocThreads) k_ploc_nn(uint32_t m, int radius, const f4* clo, const f4* chi, uint32_t* nn) {
  __shared__ f4 slo[kPlocThreads + 2 * kPlocMaxRadius], shi[kPlocThreads + 2 * kPlocMaxRadius];
  const uint32_t n_tiles = (m + kPlocThreads - 1) / kPlocThreads;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int first = (int)(tile * kPlocThreads) - radius;
    for (int k = threadIdx.x; k < kPlocThreads + 2 * radius; k += kPlocThreads) {
      const int g = first + k;
      if (g >= 0 && g < (int)m) slo[k] = clo[g], shi[k] = chi[g];
    }
    __syncthreads();
    const int i = (int)(tile * kPlocThreads + threadIdx.x);
    if (i < (int)m) {
      const f4 a = slo[threadIdx.x + radius], b = shi[threadIdx.x + radius];
      float best = INFINITY;
      int best_j = -1;
      const int j0 = max(i - radius, 0), j1 = min(i + radius, (int)m - 1);
      for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const f4 c = slo[j - first], d = shi[j - first];
        const float dx = fmaxf(b.x, d.x) - fminf(a.x, c.x), dy = fmaxf(b.y, d.y) - fminf(a.y, c.y),
                    dz = fmaxf(b.z, d.z) - fminf(a.z, c.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best) best = area, best_j = j;
      }
      nn[i] = (uint32_t)best_j;
    }
    __syncthreads();
  }
}

// valid[i] = 0 for the member of a mutual pair with the higher index (it is absorbed by its partner)
__global__ void k_ploc_flags(uint32_t m, const uint32_t* nn, uint32_t* valid) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const uint32_t j = nn[i];
    valid[i] = (nn[j] == i && j < i) ? 0u : 1u;
  }
}

// state[0] = clusters absorbed so far (= inner nodes created), state[1] = clusters after this round
__global__ void k_ploc_merge(uint32_t m, uint32_t n, const uint32_t* nn, const uint32_t* valid, const uint32_t* pos,
                             const uint32_t* ref_in, const f4* lo_in, const f4* hi_in, const uint32_t* cnt_in,
                             uint32_t* ref_out, f4* lo_out, f4* hi_out, uint32_t* cnt_out, uint32_t* child_l,
                             uint32_t* child_r, f4* ilo, f4* ihi, uint32_t* icount, uint32_t* parent_inner,
                             uint32_t* parent_leaf, const uint32_t* state) {
  const uint32_t base = state[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    const uint32_t j = nn[i];
    uint32_t r = ref_in[i], c = cnt_in[i];
    f4 lo = lo_in[i], hi = hi_in[i];
    if (nn[j] == i && j > i) {  // mutual nearest neighbours: one new inner node
      const uint32_t rank = base + (j - pos[j]);  // j is the rank-th cluster ever absorbed
      const uint32_t id = (n - 2u) - rank;        // the last merge creates node 0, the root
      const f4 lo2 = lo_in[j], hi2 = hi_in[j];
      lo = F4(fminf(lo.x, lo2.x), fminf(lo.y, lo2.y), fminf(lo.z, lo2.z), 0.f);
      hi = F4(fmaxf(hi.x, hi2.x), fmaxf(hi.y, hi2.y), fmaxf(hi.z, hi2.z), 0.f);
      c = merge_counts(c, cnt_in[j]);
      const uint32_t r2 = ref_in[j];
      child_l[id] = r;
      child_r[id] = r2;
      ilo[id] = lo, ihi[id] = hi, icount[id] = c;
      if (r & kLeafFlag) parent_leaf[r & ~kLeafFlag] = id; else parent_inner[r] = id;
      if (r2 & kLeafFlag) parent_leaf[r2 & ~kLeafFlag] = id; else parent_inner[r2] = id;
      if (id == 0u) parent_inner[0] = 0xFFFFFFFFu;
      r = id;
    }
    const uint32_t o = pos[i];
    ref_out[o] = r, lo_out[o] = lo, hi_out[o] = hi, cnt_out[o] = c;
  }
}
__global__ void k_ploc_advance(uint32_t m, const uint32_t* valid, const uint32_t* pos, uint32_t* state) {
  const uint32_t m_new = pos[m - 1] + valid[m - 1];
  state[0] += m - m_new;
  state[1] = m_new;
}

struct TreeDev {
  const uint32_t* vals;     // sorted position -> shape id
  const f4* blo;            // padded shape boxes, by shape id
  const f4* bhi;
  const uint32_t* child_l;  // inner nodes
  const uint32_t* child_r;
  const f4* ilo;
  const f4* ihi;
  const uint32_t* icount;   // subtree primitive count | kind bits
  const uint8_t* choice;    // collapse choices of k_collapse_costs, 8 per inner node (nullptr: greedy collapse)
};

__device__ __forceinline__ void ref_box(const TreeDev& t, uint32_t ref, float lo[3], float hi[3], uint32_t& count) {
  f4 a, b;
  if (ref & kLeafFlag) {
    const uint32_t p = t.vals[ref & ~kLeafFlag];
    a = t.blo[p], b = t.bhi[p], count = 1;
  } else {
    a = t.ilo[ref], b = t.ihi[ref], count = t.icount[ref] & kCountMask;
  }
  lo[0] = a.x, lo[1] = a.y, lo[2] = a.z;
  hi[0] = b.x, hi[1] = b.y, hi[2] = b.z;
}

__device__ __forceinline__ void write_prim(const BuildScene& s, uint32_t shape, WidePrim* out) {
  WidePrim p;
  for (int k = 0; k < 4; k++) p.r0[k] = p.r1[k] = p.r2[k] = 0.f;
#if HJK_PRIM_STRIDE == 4
  for (int k = 0; k < 4; k++) p.r3[k] = 0.f;
#endif
  if (shape < s.S) {
    const f4 sp = s.spheres[shape];
    p.r0[0] = sp.x, p.r0[1] = sp.y, p.r0[2] = sp.z;
    p.r1[0] = sp.w;
  } else if (shape < s.S + s.Q) {
    const f4* q = s.quads + 3 * (size_t)(shape - s.S);
    p.r0[0] = q[0].x, p.r0[1] = q[0].y, p.r0[2] = q[0].z;
    p.r1[0] = q[1].x, p.r1[1] = q[1].y, p.r1[2] = q[1].z;
    p.r2[0] = q[2].x, p.r2[1] = q[2].y, p.r2[2] = q[2].z;
  } else {
    const uint32_t* t = s.triangles + 3 * (size_t)(shape - s.S - s.Q);
    const f4 a = s.vertices[2 * (size_t)t[0]], b = s.vertices[2 * (size_t)t[1]], c = s.vertices[2 * (size_t)t[2]];
    p.r0[0] = a.x, p.r0[1] = a.y, p.r0[2] = a.z;
    // separately rounded fp32 differences, exactly what shapes/triangle.glsl:19-20 computes
    p.r1[0] = __fsub_rn(b.x, a.x), p.r1[1] = __fsub_rn(b.y, a.y), p.r1[2] = __fsub_rn(b.z, a.z);
    p.r2[0] = __fsub_rn(c.x, a.x), p.r2[1] = __fsub_rn(c.y, a.y), p.r2[2] = __fsub_rn(c.z, a.z);
  }
#if HJK_PRIM_STRIDE == 4
  if (shape >= s.S) {
    p.r3[0] = __fsub_rn(__fmul_rn(p.r1[1], p.r2[2]), __fmul_rn(p.r2[1], p.r1[2]));
    p.r3[1] = __fsub_rn(__fmul_rn(p.r1[2], p.r2[0]), __fmul_rn(p.r2[2], p.r1[0]));
    p.r3[2] = __fsub_rn(__fmul_rn(p.r1[0], p.r2[1]), __fmul_rn(p.r2[0], p.r1[1]));
  }
#endif
  p.r0[3] = __uint_as_float(shape);
  *out = p;
}

// ---------------------------------------------------------------- collapse costs
// The dynamic programme of the host builder (cwbvh_build.cpp collapse_costs), bottom-up on the device: cost[n][i-1]
// = cheapest way to represent the subtree of binary node n as at most i children of a wide node (i = 1..7), with the
// choice that achieves it; choice[n][7] = how many of a wide node's 8 children go to n's left subtree when n itself
// becomes an inner wide node.  One thread per leaf walks up; the second thread to reach a node has both children's
// costs and goes on (as k_fit_boxes).
constexpr float kCollapseNodeCost = 1.0f, kCollapsePrimCost = 0.3f;

__global__ void k_collapse_costs(int n, const uint32_t* vals, const f4* blo, const f4* bhi, const uint32_t* child_l,
                                 const uint32_t* child_r, const uint32_t* parent_inner, const uint32_t* parent_leaf,
                                 const f4* ilo, const f4* ihi, const uint32_t* icount, uint32_t* visits, float* cost,
                                 uint8_t* choice) {
  auto half_area = [](f4 lo, f4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
  };
  for (int leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n; leaf += gridDim.x * blockDim.x) {
    uint32_t node = parent_leaf[leaf];
    while (node != 0xFFFFFFFFu) {
      __threadfence();
      if (atomicAdd(visits + node, 1u) == 0u) break;
      __threadfence();
      float c[2][7];
      const uint32_t refs[2] = {child_l[node], child_r[node]};
      for (int k = 0; k < 2; k++) {
        if (refs[k] & kLeafFlag) {
          const uint32_t p = vals[refs[k] & ~kLeafFlag];
          const float a = half_area(blo[p], bhi[p]) * kCollapsePrimCost;
          for (int i = 0; i < 7; i++) c[k][i] = a;
        } else {
          const volatile float* src = cost + (size_t)refs[k] * 7;
          for (int i = 0; i < 7; i++) c[k][i] = src[i];
        }
      }
      const float area = half_area(ilo[node], ihi[node]);
      const uint32_t count = icount[node] & kCountMask;
      auto distribute = [&](int j, uint8_t& kbest) {
        float best = INFINITY;
        kbest = 1;
        for (int k = 1; k < j; k++) {
          const float v = c[0][k - 1] + c[1][j - k - 1];
          if (v < best) best = v, kbest = (uint8_t)k;
        }
        return best;
      };
      float cn[7];
      uint8_t ch[8];
      const float inner = distribute(8, ch[7]) + area * kCollapseNodeCost;
      const float as_leaf = count <= kWideMaxLeafPrims ? area * kCollapsePrimCost * (float)count : INFINITY;
      if (as_leaf <= inner) cn[0] = as_leaf, ch[0] = 0; else cn[0] = inner, ch[0] = 1;
      for (int i = 2; i <= 7; i++) {
        uint8_t k;
        const float d = distribute(i, k);
        if (d < cn[i - 2]) cn[i - 1] = d, ch[i - 1] = k; else cn[i - 1] = cn[i - 2], ch[i - 1] = 0;
      }
      for (int i = 0; i < 7; i++) cost[(size_t)node * 7 + i] = cn[i];
      for (int i = 0; i < 8; i++) choice[(size_t)node * 8 + i] = ch[i];
      node = parent_inner[node];
    }
  }
}

// Between two levels of k_collapse: the tasks just emitted become the next level's input; counts the levels
// that had work (no host round trip per level).  lvl[0] = n_in, lvl[1] = n_out, lvl[2] = depth.
__global__ void k_next_level(uint32_t* lvl) {
  const uint32_t n = lvl[1];
  lvl[0] = n;
  lvl[1] = 0;
  if (n) lvl[2] += 1;
}

// counters: [0] wide nodes allocated, [1] primitive records allocated, [2] overflow flag
// One thread per wide node of the current level.
__global__ void k_collapse(BuildScene s, TreeDev t, const uint2* tasks_in, const uint32_t* n_in, uint2* tasks_out,
                           uint32_t* n_out, WideNode* nodes, WidePrim* prims, uint32_t* counters, uint32_t node_capacity) {
  const uint32_t n_tasks = *n_in;
  for (uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x; ti < n_tasks; ti += gridDim.x * blockDim.x) {
    const uint32_t wide = tasks_in[ti].x, root = tasks_in[ti].y;
    uint32_t ref[8], cnt[8];
    bool as_leaf[8];
    float lo[8][3], hi[8][3];
    int ne = 0;
    if (t.choice) {  // the children the collapse costs chose (cwbvh_build.cpp gather_children)
      uint32_t st_ref[16];
      int st_i[16], sp = 0;
      const int k8 = t.choice[(size_t)root * 8 + 7];
      st_ref[sp] = t.child_r[root], st_i[sp++] = 8 - k8;
      st_ref[sp] = t.child_l[root], st_i[sp++] = k8;
      while (sp > 0 && ne < 8) {
        const uint32_t r = st_ref[--sp];
        int ii = st_i[sp];
        if (r & kLeafFlag) {
          ref[ne] = r, as_leaf[ne++] = true;
          continue;
        }
        while (ii >= 2 && t.choice[(size_t)r * 8 + ii - 1] == 0) ii--;
        if (ii == 1) {
          ref[ne] = r, as_leaf[ne++] = t.choice[(size_t)r * 8] == 0;
          continue;
        }
        const int k = t.choice[(size_t)r * 8 + ii - 1];
        st_ref[sp] = t.child_r[r], st_i[sp++] = ii - k;
        st_ref[sp] = t.child_l[r], st_i[sp++] = k;
      }
      for (int k = 0; k < ne; k++) ref_box(t, ref[k], lo[k], hi[k], cnt[k]);
    } else {  // greedy: repeatedly open the child of largest area that is still a subtree of more than 3 primitives
      float area[8];
      ne = 2;
      ref[0] = t.child_l[root], ref[1] = t.child_r[root];
      for (int k = 0; k < 2; k++) {
        ref_box(t, ref[k], lo[k], hi[k], cnt[k]);
        const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
        area[k] = dx * dy + dy * dz + dz * dx;
      }
      while (ne < 8) {
        int best = -1;
        for (int k = 0; k < ne; k++)
          if (cnt[k] > kWideMaxLeafPrims && (best < 0 || area[k] > area[best])) best = k;
        if (best < 0) break;
        const uint32_t r = ref[best];
        const uint32_t pair[2] = {t.child_l[r], t.child_r[r]};
        const int at[2] = {best, ne};
        for (int c = 0; c < 2; c++) {
          const int k = at[c];
          ref[k] = pair[c];
          ref_box(t, ref[k], lo[k], hi[k], cnt[k]);
          const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
          area[k] = dx * dy + dy * dz + dz * dx;
        }
        ne++;
      }
      for (int k = 0; k < ne; k++) as_leaf[k] = cnt[k] <= kWideMaxLeafPrims;
    }
    // node box, slot assignment (greedy on sum of dot(child centre - node centre, slot signs))
    float nlo[3] = {INFINITY, INFINITY, INFINITY}, nhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < ne; k++)
      for (int a = 0; a < 3; a++) nlo[a] = fminf(nlo[a], lo[k][a]), nhi[a] = fmaxf(nhi[a], hi[k][a]);
    int child_in_slot[8];
    for (int sl = 0; sl < 8; sl++) child_in_slot[sl] = -1;
    uint32_t child_done = 0, slot_used = 0;
    for (int it = 0; it < ne; it++) {
      int bc = -1, bs = -1;
      float bestv = -INFINITY;
      for (int c = 0; c < ne; c++) {
        if (child_done & (1u << c)) continue;
        float d[3];
        for (int a = 0; a < 3; a++) d[a] = 0.5f * (lo[c][a] + hi[c][a]) - 0.5f * (nlo[a] + nhi[a]);
        for (int sl = 0; sl < 8; sl++) {
          if (slot_used & (1u << sl)) continue;
          const float v = ((sl & 1) ? d[0] : -d[0]) + ((sl & 2) ? d[1] : -d[1]) + ((sl & 4) ? d[2] : -d[2]);
          if (v > bestv) bestv = v, bc = c, bs = sl;
        }
      }
      child_in_slot[bs] = bc;
      child_done |= 1u << bc;
      slot_used |= 1u << bs;
    }
    // allocation: inner children contiguous (slot order), primitive records contiguous
    uint32_t n_inner = 0, n_prims = 0;
    for (int k = 0; k < ne; k++) {
      if (!as_leaf[k]) n_inner++; else n_prims += cnt[k];
    }
    const uint32_t child_base = n_inner ? atomicAdd(counters + 0, n_inner) : 0u;
    const uint32_t prim_base = n_prims ? atomicAdd(counters + 1, n_prims) : 0u;
    if (child_base + n_inner > node_capacity) {
      atomicExch(counters + 2, 1u);
      continue;
    }
    const uint32_t task_base = n_inner ? atomicAdd(n_out, n_inner) : 0u;
    WideNode wn;
    double scale[3];
    for (int a = 0; a < 3; a++) {
      wn.origin[a] = nlo[a];
      const double extent = (double)nhi[a] - (double)nlo[a];
      int e = extent > 0.0 ? (int)ceil(log2(extent / 255.0)) : -126;
      e = max(e, -126);
      while (ceil(extent / ldexp(1.0, e)) > 255.0) e++;
      e = min(e, 127);
      wn.e[a] = (uint8_t)(e + 127);
      scale[a] = ldexp(1.0, e);
    }
    wn.imask = 0;
    wn.child_base = child_base;
    const uint32_t kind = t.icount[root] & ~kCountMask;  // what lives below this node
    wn.prim_base = prim_base | ((kind & kKindSphere) ? kWideHasSpheres : 0u) | (kind == kKindSphere ? kWideOnlySpheres : 0u);
    uint32_t inner_rank = 0, prim_off = 0;
    for (int sl = 0; sl < 8; sl++) {
      const int c = child_in_slot[sl];
      if (c < 0) {
        wn.meta[sl] = 0;
        for (int a = 0; a < 3; a++) wn.qlo[a][sl] = 255, wn.qhi[a][sl] = 0;
        continue;
      }
      for (int a = 0; a < 3; a++) {
        double ql = floor(((double)lo[c][a] - (double)wn.origin[a]) / scale[a]);
        double qh = ceil(((double)hi[c][a] - (double)wn.origin[a]) / scale[a]);
        ql = fmin(fmax(ql, 0.0), 255.0);
        qh = fmin(fmax(qh, 0.0), 255.0);
        wn.qlo[a][sl] = (uint8_t)ql;
        wn.qhi[a][sl] = (uint8_t)qh;
      }
      if (!as_leaf[c]) {
        wn.meta[sl] = (uint8_t)((1u << 5) | (24u + (uint32_t)sl));
        wn.imask |= (uint8_t)(1u << sl);
        tasks_out[task_base + inner_rank] = make_uint2(child_base + inner_rank, ref[c]);
        inner_rank++;
      } else {
        wn.meta[sl] = (uint8_t)((((1u << cnt[c]) - 1u) << 5) | prim_off);
        // the <= 3 primitives of this small subtree, left to right
        uint32_t stack[4];
        int sp = 0;
        stack[sp++] = ref[c];
        uint32_t written = 0;
        while (sp > 0) {
          const uint32_t r = stack[--sp];
          if (r & kLeafFlag) {
            write_prim(s, t.vals[r & ~kLeafFlag], prims + prim_base + prim_off + written);
            written++;
          } else {
            stack[sp++] = t.child_r[r];
            stack[sp++] = t.child_l[r];
          }
        }
        prim_off += cnt[c];
      }
    }
    nodes[wide] = wn;
  }
}

}  // namespace gpubvh
}  // namespace hjk
