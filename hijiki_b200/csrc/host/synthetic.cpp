// Synthetic scenes named by BASELINE.json configs 3 and 4 (definitions: SURVEY.md §8d).
// They are expressed in the reference's own Scene vocabulary (triangles, spheres,
// Diffuse / DiffuseCBoard / Mirror / Dielectric / Emissive materials, src/main.rs:34-52)
// so that the same compile + upload path as cbox consumes them.
#include <cmath>

#include "host_scene.h"

namespace hjk {
namespace {

uint32_t hash2(int32_t x, int32_t y, uint32_t seed) {
  uint32_t h = seed ^ 0x9E3779B9u;
  h ^= (uint32_t)x * 0x85EBCA6Bu;
  h = (h << 13) | (h >> 19);
  h ^= (uint32_t)y * 0xC2B2AE35u;
  h *= 0x27D4EB2Du;
  h ^= h >> 15;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  return h;
}

// lattice value noise in [-1,1] with smoothstep interpolation
double value_noise(double x, double y, uint32_t seed) {
  double fx = std::floor(x), fy = std::floor(y);
  int32_t ix = (int32_t)fx, iy = (int32_t)fy;
  double tx = x - fx, ty = y - fy;
  double sx = tx * tx * (3.0 - 2.0 * tx), sy = ty * ty * (3.0 - 2.0 * ty);
  auto v = [&](int32_t a, int32_t b) { return (double)hash2(a, b, seed) / 2147483647.5 - 1.0; };
  double a = v(ix, iy), b = v(ix + 1, iy), c = v(ix, iy + 1), d = v(ix + 1, iy + 1);
  double top = a + (b - a) * sx, bot = c + (d - c) * sx;
  return top + (bot - top) * sy;
}

// 5 octaves, amplitude 2, lacunarity 2, gain 0.5
double terrain_height(double x, double z, uint32_t seed) {
  double amp = 2.0, freq = 0.15, h = 0.0;
  for (int o = 0; o < 5; o++) {
    h += amp * value_noise(x * freq, z * freq, seed + 131u * (uint32_t)o);
    amp *= 0.5;
    freq *= 2.0;
  }
  return h;
}

void set_camera(Scene& s, float px, float py, float pz, float pitch_deg, float hfov_deg) {
  const double half = 0.5 * (double)pitch_deg * M_PI / 180.0;
  s.camera = HjkCamera{};
  s.camera.position[0] = px;
  s.camera.position[1] = py;
  s.camera.position[2] = pz;
  s.camera.rotation[0] = (float)std::sin(half);
  s.camera.rotation[3] = (float)std::cos(half);
  s.camera.fov = hfov_deg;
}

// a horizontal rectangle at height y as two triangles; normal (0, ny, 0)
void add_rect(Scene& s, float x0, float z0, float x1, float z1, float y, float ny, size_t mat) {
  uint32_t base = (uint32_t)s.vertices.size();
  const float xs[4] = {x0, x1, x1, x0};
  const float zs[4] = {z0, z0, z1, z1};
  for (int k = 0; k < 4; k++) {
    HjkVertex v{};
    v.pos[0] = xs[k];
    v.pos[1] = y;
    v.pos[2] = zs[k];
    v.u = (xs[k] - x0) / (x1 - x0);
    v.v = (zs[k] - z0) / (z1 - z0);
    v.normal[1] = ny;
    s.vertices.push_back(v);
  }
  Shape a;
  a.kind = ShapeKind::Triangle;
  a.tri = {base, base + 1, base + 2};
  Shape b;
  b.kind = ShapeKind::Triangle;
  b.tri = {base, base + 2, base + 3};
  s.objects.emplace_back(a, mat);
  s.objects.emplace_back(b, mat);
}

}  // namespace

void make_terrain_scene(uint32_t n, uint64_t seed, Scene& s) {
  s = Scene();
  set_camera(s, 0.f, 4.f, 14.f, -15.f, 40.f);
  Material terrain;
  terrain.tag = HJK_MAT_DIFFUSECBOARD;  // the "textured" material of the north star
  terrain.cboard = HjkDiffuseCB{{0.75f, 0.72f, 0.65f}, 0.05f, {0.25f, 0.35f, 0.22f}, 0.05f};
  s.materials.push_back(terrain);
  Material light;
  light.tag = HJK_MAT_EMISSIVE;
  light.color = HjkColor16{{20.f, 20.f, 20.f}, 0.f};
  s.materials.push_back(light);

  const uint32_t nv = n + 1;
  const uint32_t sd = (uint32_t)(seed ^ (seed >> 32));
  const double step = 20.0 / (double)n;
  s.vertices.resize((size_t)nv * nv);
  for (uint32_t j = 0; j < nv; j++) {
    for (uint32_t i = 0; i < nv; i++) {
      double x = -10.0 + step * i, z = -10.0 + step * j;
      double h = terrain_height(x, z, sd);
      const double d = 1e-3;
      double hx = terrain_height(x + d, z, sd) - terrain_height(x - d, z, sd);
      double hz = terrain_height(x, z + d, sd) - terrain_height(x, z - d, sd);
      double nx = -hx / (2 * d), ny = 1.0, nz = -hz / (2 * d);
      double inv = 1.0 / std::sqrt(nx * nx + ny * ny + nz * nz);
      HjkVertex v{};
      v.pos[0] = (float)x;
      v.pos[1] = (float)h;
      v.pos[2] = (float)z;
      v.u = (float)(x / 20.0);
      v.v = (float)(z / 20.0);
      v.normal[0] = (float)(nx * inv);
      v.normal[1] = (float)(ny * inv);
      v.normal[2] = (float)(nz * inv);
      s.vertices[(size_t)j * nv + i] = v;
    }
  }
  s.objects.reserve((size_t)2 * n * n + 32);
  for (uint32_t j = 0; j < n; j++) {
    for (uint32_t i = 0; i < n; i++) {
      uint32_t a = j * nv + i, b = a + 1, c = a + nv, d = c + 1;
      Shape t0;
      t0.kind = ShapeKind::Triangle;
      t0.tri = {a, c, b};
      Shape t1;
      t1.kind = ShapeKind::Triangle;
      t1.tri = {b, c, d};
      s.objects.emplace_back(t0, 0);
      s.objects.emplace_back(t1, 0);
    }
  }
  // 16 emissive 1x1 quads (32 triangles, power 20) in a 4x4 grid at y = 6, facing down
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) {
      float cx = -7.5f + 5.f * i, cz = -7.5f + 5.f * j;
      add_rect(s, cx - 0.5f, cz - 0.5f, cx + 0.5f, cz + 0.5f, 6.f, -1.f, 1);
    }
}

void make_spheres_scene(uint32_t n, uint64_t seed, Scene& s) {
  s = Scene();
  const float half = 0.5f * (float)(n - 1);
  set_camera(s, 0.f, 0.5f * (float)n, 2.0f * (float)n, 0.f, 40.f);
  Material floor;
  floor.tag = HJK_MAT_DIFFUSE;
  floor.color = HjkColor16{{0.7f, 0.7f, 0.7f}, 0.f};
  s.materials.push_back(floor);  // 0
  Material light;
  light.tag = HJK_MAT_EMISSIVE;
  light.color = HjkColor16{{30.f, 30.f, 30.f}, 0.f};
  s.materials.push_back(light);  // 1
  Material glass;
  glass.tag = HJK_MAT_DIELECTRIC;
  glass.dielectric = HjkDielectric{{0.f, 0.f, 0.f, 1.5f}};  // DielectricMaterial::clear(1.5)
  s.materials.push_back(glass);  // 2
  Material mirror;
  mirror.tag = HJK_MAT_MIRROR;
  s.materials.push_back(mirror);  // 3

  SplitMix64 rng(seed);
  for (uint32_t k = 0; k < n; k++)
    for (uint32_t j = 0; j < n; j++)
      for (uint32_t i = 0; i < n; i++) {
        Shape sp;
        sp.kind = ShapeKind::Sphere;
        sp.sphere.position[0] = (float)i - half;
        sp.sphere.position[1] = (float)j + 0.5f;
        sp.sphere.position[2] = (float)k - half;
        sp.sphere.radius = 0.25f + 0.17f * rng.next_f32();
        s.objects.emplace_back(sp, ((i + j + k) & 1u) ? 3 : 2);
      }
  const float ext = 1.5f * (float)n;
  add_rect(s, -ext, -ext, ext, ext, 0.f, 1.f, 0);
  const float le = 0.375f * (float)n;
  add_rect(s, -le, -le, le, le, (float)n + 2.f, -1.f, 1);
}

}  // namespace hjk
