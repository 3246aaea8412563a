// Scene::compile (reference src/main.rs:172-358): partition shapes by type, assign the
// global shape index space [spheres | quads | triangles], pack one material word per
// shape, derive the emitter table, and (optionally) build the reference-layout BVH2.
#include <cmath>

#include "host_scene.h"

namespace hjk {

namespace {

Aabb shape_aabb(const Scene& scene, const Shape& s) {
  const float inf = INFINITY;
  Aabb a{{inf, inf, inf}, {-inf, -inf, -inf}};
  auto grow = [&](const float p[3]) {
    for (int k = 0; k < 3; k++) {
      a.min[k] = std::fmin(a.min[k], p[k]);
      a.max[k] = std::fmax(a.max[k], p[k]);
    }
  };
  switch (s.kind) {
    case ShapeKind::Sphere:  // src/shape.rs:13-20
      for (int k = 0; k < 3; k++) {
        a.min[k] = s.sphere.position[k] - s.sphere.radius;
        a.max[k] = s.sphere.position[k] + s.sphere.radius;
      }
      break;
    case ShapeKind::Quad: {  // src/shape.rs:46-53
      float p[3];
      grow(s.quad.origin);
      for (int k = 0; k < 3; k++) p[k] = s.quad.origin[k] + s.quad.edge1[k];
      grow(p);
      for (int k = 0; k < 3; k++) p[k] = s.quad.origin[k] + s.quad.edge2[k];
      grow(p);
      for (int k = 0; k < 3; k++) p[k] = s.quad.origin[k] + s.quad.edge1[k] + s.quad.edge2[k];
      grow(p);
      break;
    }
    case ShapeKind::Triangle:  // src/main.rs:74-79
      for (int v = 0; v < 3; v++) grow(scene.vertices[s.tri[v]].pos);
      break;
  }
  return a;
}

}  // namespace

void compile_scene(const Scene& scene, bool with_bvh2, CompiledScene& out) {
  out = CompiledScene();
  std::vector<size_t> sphere_mat, quad_mat, tri_mat;
  std::vector<size_t> shape_indices;  // index of each object within its own type array
  shape_indices.reserve(scene.objects.size());
  for (const auto& obj : scene.objects) {
    const Shape& s = obj.first;
    switch (s.kind) {
      case ShapeKind::Sphere:
        shape_indices.push_back(out.spheres.size());
        out.spheres.push_back(s.sphere);
        sphere_mat.push_back(obj.second);
        break;
      case ShapeKind::Quad:
        shape_indices.push_back(out.quads.size());
        out.quads.push_back(s.quad);
        quad_mat.push_back(obj.second);
        break;
      case ShapeKind::Triangle:
        shape_indices.push_back(out.triangles.size());
        out.triangles.push_back(s.tri);
        tri_mat.push_back(obj.second);
        break;
    }
  }

  if (with_bvh2 && !scene.objects.empty()) {
    std::vector<Aabb> aabbs;
    aabbs.reserve(scene.objects.size());
    for (const auto& obj : scene.objects) aabbs.push_back(shape_aabb(scene, obj.first));
    build_flat_bvh2(aabbs, out.bvh);
    // transform shape indices (src/main.rs:232-243)
    for (HjkBvh2Node& n : out.bvh) {
      if (n.shape_index == 0xFFFFFFFFu) continue;
      size_t offset = 0;
      switch (scene.objects[n.shape_index].first.kind) {
        case ShapeKind::Sphere: offset = 0; break;
        case ShapeKind::Quad: offset = out.spheres.size(); break;
        case ShapeKind::Triangle: offset = out.spheres.size() + out.quads.size(); break;
      }
      n.shape_index = (uint32_t)(shape_indices[n.shape_index] + offset);
    }
  }

  // material words (src/main.rs:246-287)
  std::vector<uint32_t> reprs;
  for (const Material& m : scene.materials) {
    size_t ix = 0;
    switch (m.tag) {
      case HJK_MAT_DIFFUSE:
        out.diffuse.push_back(m.color);
        ix = out.diffuse.size() - 1;
        break;
      case HJK_MAT_DIFFUSECBOARD:
        out.diffusecb.push_back(m.cboard);
        ix = out.diffusecb.size() - 1;
        break;
      case HJK_MAT_MIRROR: ix = 0; break;
      case HJK_MAT_DIELECTRIC:
        out.dielectric.push_back(m.dielectric);
        ix = out.dielectric.size() - 1;
        break;
      case HJK_MAT_EMISSIVE:
        out.emissive.push_back(m.color);
        ix = out.emissive.size() - 1;
        break;
    }
    reprs.push_back(((uint32_t)m.tag << HJK_MATERIAL_TAG_SHIFT) + (uint32_t)ix);
  }
  for (size_t m : sphere_mat) out.materials.push_back(reprs[m]);
  for (size_t m : quad_mat) out.materials.push_back(reprs[m]);
  for (size_t m : tri_mat) out.materials.push_back(reprs[m]);

  // emitter table (src/main.rs:289-307): uniform pdf, running cdf
  for (size_t ix = 0; ix < out.materials.size(); ix++) {
    if ((out.materials[ix] >> HJK_MATERIAL_TAG_SHIFT) == (uint32_t)HJK_MAT_EMISSIVE)
      out.emitters.push_back(HjkEmitter{(uint32_t)ix, 0.f, 0.f, 0.f});
  }
  float emitter_pdf = 1.0f / (float)out.emitters.size();
  float cdf = 0.f;
  for (HjkEmitter& e : out.emitters) {
    cdf += emitter_pdf;
    e.pdf = emitter_pdf;
    e.cdf = cdf;
  }

  out.vertices = scene.vertices;
  out.info.camera = scene.camera;
  out.info.num_spheres = (uint32_t)out.spheres.size();
  out.info.num_quads = (uint32_t)out.quads.size();
  out.info.num_triangles = (uint32_t)out.triangles.size();
  out.info.num_emitters = (uint32_t)out.emitters.size();
}

}  // namespace hjk
