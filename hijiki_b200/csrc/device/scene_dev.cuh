// Device-side view of the compiled scene (reference bindings, src/main.rs:314-327) plus the
// wide BVH, and the 16-byte load helpers every kernel uses.
#pragma once
#include "../../../include/hijiki_b200.h"
#include "hjk_math.cuh"

namespace hjk {

struct alignas(16) f4 {
  float x, y, z, w;
};
struct alignas(16) u4 {
  uint32_t x, y, z, w;
};
HJK_HD f4 F4(float a, float b, float c, float d) {
  f4 r;
  r.x = a;
  r.y = b;
  r.z = c;
  r.w = d;
  return r;
}

// 16-byte read-only load (LDG.E.128.CONSTANT on the device)
HJK_HD f4 ld16(const f4* p) {
#if defined(__CUDA_ARCH__)
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4(v.x, v.y, v.z, v.w);
#else
  return *p;
#endif
}
HJK_HD uint32_t ld4(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
HJK_HD vec3 xyz(const f4& v) { return V3(v.x, v.y, v.z); }

struct SceneDev {
  // wide BVH (cwbvh.h): 5 x f4 per node, 3 x f4 per primitive record
  const f4* nodes;
  const f4* prims;
  // reference bindings
  const f4* spheres;          // (centre, radius)
  const f4* quads;            // 3 per quad
  const uint32_t* triangles;  // 3 per triangle
  const f4* vertices;         // 2 per vertex: (pos, u) (normal, v)
  const uint32_t* materials;  // one word per shape
  const f4* emitters;         // (shape bits, pdf, cdf, pad)
  const f4* diffuse;
  const f4* diffusecb;        // 2 per material
  const f4* dielectric;
  const f4* emissive;
  uint32_t num_spheres, num_quads, num_triangles, num_emitters;
  HjkCamera camera;
  // bounding ball of the sphere CENTRES (xyz, radius) and the sphere radius range: inputs of the
  // traversal's sphere guard (traverse.cuh)
  float sph_centre[4];
  float sph_rmin, sph_rmax;
};

}  // namespace hjk
