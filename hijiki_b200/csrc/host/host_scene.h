// Host-side mirror of the reference's Rust front-end types (no CUDA here).
//   Scene / Shape / Material      reference src/main.rs:34-170
//   CompiledScene                 reference src/main.rs:376-397
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/hijiki_b200.h"

namespace hjk {

static_assert(sizeof(HjkCamera) == 48, "Camera layout (render.glsl:12-16)");
static_assert(sizeof(HjkSceneInfo) == 64, "SceneBufferInfo layout (src/main.rs:400-408)");
static_assert(sizeof(HjkBvh2Node) == 32, "CompiledBVHNode layout (src/main.rs:92-99)");
static_assert(sizeof(HjkSphere) == 16, "Sphere layout (src/shape.rs:6-11)");
static_assert(sizeof(HjkQuad) == 48, "Quad layout (src/shape.rs:22-31)");
static_assert(sizeof(HjkVertex) == 32, "Vertex layout (shapes/triangle.glsl:1-4)");
static_assert(sizeof(HjkEmitter) == 16, "Emitter layout (src/main.rs:368-374)");
static_assert(sizeof(HjkColor16) == 16, "vec3 material padded to 16");
static_assert(sizeof(HjkDiffuseCB) == 32, "DiffuseCheckerboardMaterial (src/main.rs:108-115)");
static_assert(sizeof(HjkDielectric) == 16, "DielectricMaterial (src/main.rs:122-126)");
static_assert(sizeof(HjkImageBlock) == 40, "ImageBlock layout (src/main.rs:608-617)");
static_assert(sizeof(HjkRay) == 32, "ray batch layout");

enum class ShapeKind : uint8_t { Sphere, Quad, Triangle };

// reference `enum Shape` (src/main.rs:47-52)
struct Shape {
  ShapeKind kind;
  HjkSphere sphere{};
  HjkQuad quad{};
  std::array<uint32_t, 3> tri{};
};

// reference `enum Material` (src/main.rs:34-44); tag == enum discriminant
struct Material {
  uint8_t tag = HJK_MAT_DIFFUSE;
  HjkColor16 color{};       // Diffuse.color / Emissive.power
  HjkDiffuseCB cboard{};    // DiffuseCBoard
  HjkDielectric dielectric{};
};

// reference `struct Scene` (src/main.rs:162-170)
struct Scene {
  HjkCamera camera{};
  std::vector<std::pair<Shape, size_t>> objects;  // (shape, material index)
  std::vector<HjkVertex> vertices;
  std::vector<Material> materials;
};

// reference `struct CompiledScene` (src/main.rs:376-397)
struct CompiledScene {
  HjkSceneInfo info{};
  std::vector<HjkBvh2Node> bvh;
  std::vector<HjkSphere> spheres;
  std::vector<HjkQuad> quads;
  std::vector<std::array<uint32_t, 3>> triangles;
  std::vector<HjkVertex> vertices;
  std::vector<uint32_t> materials;
  std::vector<HjkEmitter> emitters;
  std::vector<HjkColor16> diffuse;
  std::vector<HjkDiffuseCB> diffusecb;
  std::vector<HjkDielectric> dielectric;
  std::vector<HjkColor16> emissive;
};

// Scene::from_obj, src/main.rs:413-530 (tobj 0.1.11 semantics restated in obj_loader.cpp)
bool scene_from_obj(const std::string& path, Scene& out, std::string& err);
// main()'s --put-cbox-spheres block, src/main.rs:1463-1483
void put_cbox_spheres(Scene& scene);
// Scene::compile, src/main.rs:172-358
void compile_scene(const Scene& scene, bool with_bvh2, CompiledScene& out);

// bvh 0.3.1 `BVH::build` + the flatten of src/main.rs:203-244
struct Aabb {
  float min[3];
  float max[3];
};
void build_flat_bvh2(const std::vector<Aabb>& shape_aabbs, std::vector<HjkBvh2Node>& flat);

// synthetic scenes of BASELINE.json configs 3 / 4
void make_terrain_scene(uint32_t grid_n, uint64_t seed, Scene& out);
void make_spheres_scene(uint32_t lattice_n, uint64_t seed, Scene& out);

// splitmix64 — the recorded stream that replaces rand::random() (src/main.rs:643,670,675)
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t next_u32() { return (uint32_t)(next() >> 32); }
  // uniform in [0,1) with 24 bits, like rand 0.7's Standard f32
  float next_f32() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
};

}  // namespace hjk
