// Compressed wide BVH (8-ary, quantised child boxes) — the layout that replaces the
// reference's threaded binary BVH (reference src/main.rs:92-99, shader/scene.glsl:10-17,
// 99-133).  The reference tree is one-primitive-per-leaf, walked in storage order; this one
// is walked near-child-first with 16-byte vector loads.  POD shared by the host builder
// (host/cwbvh_build.cpp) and the device traversal (device/traverse.cuh).
#pragma once
#include <stdint.h>

namespace hjk {

// 80-byte node = five 16-byte loads:
//   q0: origin.xyz, {ex, ey, ez, imask}
//   q1: child_base, prim_base | kWideHasSpheres | kWideOnlySpheres, meta[0..3], meta[4..7]
//   q2: qlo_x[0..7], qlo_y[0..7]      q3: qlo_z[0..7], qhi_x[0..7]      q4: qhi_y[0..7], qhi_z[0..7]
// Child box i on axis a spans origin[a] + q{lo,hi}_a[i] * 2^(e[a]-127).
// meta[i]: 0 = empty slot; inner child = 0b001'11sss (sss = slot, bit index 24+slot);
//          leaf = (unary prim count 1/3/7) << 5 | first prim offset within the node (0..23).
struct WideNode {
  float origin[3];
  uint8_t e[3];
  uint8_t imask;  // bit s set = slot s holds an inner node
  uint32_t child_base;
  uint32_t prim_base;
  uint8_t meta[8];
  uint8_t qlo[3][8];
  uint8_t qhi[3][8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be five 16-byte words");

// 48-byte primitive record = three 16-byte loads.  The global shape id (reference index
// space [spheres | quads | triangles], src/main.rs:233-243) rides in r0.w.
//   triangle: r0 = (a.xyz, id)      r1 = (b-a, 0)        r2 = (c-a, 0)   (fp32 differences, as
//             shapes/triangle.glsl:19-20 computes them per ray)
//   sphere  : r0 = (centre.xyz, id) r1 = (radius, 0,0,0) r2 = 0
//   quad    : r0 = (origin.xyz, id) r1 = (edge1, 0)      r2 = (edge2, 0)
// HJK_PRIM_STRIDE = 4 adds r3 = (cross(r1, r2), 0): the fp32 cross product the reference evaluates
// per ray (shapes/triangle.glsl:22), precomputed with the same separately rounded operations.
#ifndef HJK_PRIM_STRIDE
#define HJK_PRIM_STRIDE 3
#endif
struct WidePrim {
  float r0[4];
  float r1[4];
  float r2[4];
#if HJK_PRIM_STRIDE == 4
  float r3[4];
#endif
};
static_assert(sizeof(WidePrim) == 16 * HJK_PRIM_STRIDE, "WidePrim must be whole 16-byte words");

enum : uint32_t { kWideMaxLeafPrims = 3, kWideMaxNodePrims = 24 };
// Top bit of WideNode::prim_base: a sphere lives somewhere below this node (in one of its leaf children or in
// a descendant).  Only such nodes need the traversal's sphere guard (traverse.cuh): triangle and quad tests
// are geometric for any direction length, so the rest of the tree is culled with the plain slab test.
// Second bit: EVERY primitive below the node is a sphere.  For directions shorter than unit length the reference's
// sphere test can only accept spheres close to the ray origin (traverse.cuh, "far clip"), so such subtrees are
// walked with a finite far distance.
enum : uint32_t { kWideHasSpheres = 0x80000000u, kWideOnlySpheres = 0x40000000u, kWidePrimBaseMask = 0x3FFFFFFFu };

}  // namespace hjk
