// Host-side container + builder entry of the 8-wide compressed BVH (see cwbvh.h).
#pragma once
#include <string>
#include <vector>

#include "../../../include/hijiki_b200.h"
#include "../cwbvh.h"

namespace hjk {

struct WideBvh {
  std::vector<WideNode> nodes;  // nodes[0] = root
  std::vector<WidePrim> prims;
  float scene_min[3] = {0, 0, 0}, scene_max[3] = {0, 0, 0};
  float pad = 0.f;       // absolute outward pad applied to every primitive box
  float sah_cost = 0.f;  // collapse cost of the root, relative to the root area
  uint32_t depth = 0;
  uint32_t n_shapes = 0;
  // bounding ball of the sphere centres and radius range (sphere guard of the traversal)
  float sph_centre[4] = {0, 0, 0, 0};
  float sph_rmin = 0.f, sph_rmax = 0.f;
};

// default relative pad (times the larger of scene extent and |coordinate|)
constexpr float kDefaultBvhPadRel = 1e-5f;

bool build_wide_bvh(const HjkScene& scene, float pad_rel, WideBvh& out, std::string& err);
// cap on the host threads the builder uses (0 = all of them, the default)
void set_builder_threads(int n);
bool validate_wide_bvh(const HjkScene& scene, const WideBvh& bvh, std::string& err);
void sphere_guard_bounds(const HjkScene& scene, WideBvh& out);

}  // namespace hjk
