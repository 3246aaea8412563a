// CPU ORACLE — TEST INFRASTRUCTURE ONLY (see hijiki_oracle.h).  PARITY UNPINNED by the
// reference (no tests / fixtures upstream); pinned by SURVEY §8c known-answer vectors.
//
// Literal restatement of the reference GLSL; transcendentals per orc_math.h.  Every function cites the shader lines it
// follows.  Build: g++ -O2 -ffp-contract=off -fno-fast-math (see Makefile) so that every
// fp32 operation is a separate IEEE operation in source order.
#include "hijiki_oracle.h"
#include "orc_math.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------ math (GLSL built-ins)
struct vec3 {
  float x, y, z;
};
inline vec3 V(float x, float y, float z) { return vec3{x, y, z}; }
inline vec3 V(float s) { return vec3{s, s, s}; }
inline vec3 operator+(vec3 a, vec3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return V(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return V(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(vec3 a, vec3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
  return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) {  // convention fixed by the oracle: v * inversesqrt(dot(v,v))
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
inline vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }
inline vec3 vexp(vec3 a) { return V(orc_expf(a.x), orc_expf(a.y), orc_expf(a.z)); }
inline float fract(float x) { return x - floorf(x); }
inline float fminf_(float a, float b) { return a < b ? a : b; }  // GLSL min(x,y): y<x ? y : x
inline float glsl_min(float x, float y) { return y < x ? y : x; }
inline float glsl_max(float x, float y) { return x < y ? y : x; }

struct mat3 {
  vec3 c0, c1, c2;
};
inline vec3 mul(const mat3& m, vec3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }

// math.glsl:1-2
constexpr float M_PI_F = 3.1415926535897932384626433832795f;

// ------------------------------------------------------------------ scene views
struct Camera {  // render.glsl:12-16
  float position[4];
  float rotation[4];
  float fov;
  float pad[3];
};
struct SceneInfo {  // scene.glsl:1-8
  Camera camera;
  int32_t numSpheres, numQuads, numTriangles, numEmitters;
};
struct BVHNode {  // scene.glsl:10-13
  float aabbMin[3];
  uint32_t shapeIndex;
  float aabbMax[3];
  uint32_t exitIndex;
};
struct Sphere {  // shapes/sphere.glsl:1-3
  float positionRadius[4];
};
struct Quad {  // shapes/quad.glsl:1-5 (std430 vec3 = 16 B stride)
  float origin[4], edge1[4], edge2[4];
};
struct Vertex {  // shapes/triangle.glsl:1-4
  float pos_u[4];
  float norm_v[4];
};
struct Emitter {  // scene.glsl:33-38
  uint32_t shape;
  float pdf, cdf, pad;
};
struct Vec4 {
  float v[4];
};

struct SceneView {
  const SceneInfo* info;
  const BVHNode* bvh;
  uint64_t bvhLength;
  const Sphere* spheres;
  const Quad* quads;
  const uint32_t* triangles;
  const Vertex* vertices;
  const uint32_t* materials;
  const Emitter* emitters;
  const Vec4* diffuseMaterials;     // vec3 color (16 B stride)
  const Vec4* diffuseCBMaterials;   // 2 x vec4 per material
  const Vec4* dielectricMaterials;  // extinction_etaRatio
  const Vec4* emissiveMaterials;    // vec3 power (16 B stride)
  int numSpheres, numQuads, numTriangles, numEmitters;
  const struct CullAccel* accel = nullptr;  // mode 3 only (see CullAccel)
};

SceneView make_view(const OrcScene* s) {
  SceneView v{};
  v.info = (const SceneInfo*)s->scene.ptr;
  v.bvh = (const BVHNode*)s->bvh.ptr;
  v.bvhLength = s->bvh.count;
  v.spheres = (const Sphere*)s->spheres.ptr;
  v.quads = (const Quad*)s->quads.ptr;
  v.triangles = (const uint32_t*)s->triangles.ptr;
  v.vertices = (const Vertex*)s->vertices.ptr;
  v.materials = (const uint32_t*)s->materials.ptr;
  v.emitters = (const Emitter*)s->emitters.ptr;
  v.diffuseMaterials = (const Vec4*)s->diffuse.ptr;
  v.diffuseCBMaterials = (const Vec4*)s->diffusecb.ptr;
  v.dielectricMaterials = (const Vec4*)s->dielectric.ptr;
  v.emissiveMaterials = (const Vec4*)s->emissive.ptr;
  v.numSpheres = v.info->numSpheres;
  v.numQuads = v.info->numQuads;
  v.numTriangles = v.info->numTriangles;
  v.numEmitters = v.info->numEmitters;
  return v;
}

constexpr uint32_t MATERIAL_TAG_SHIFT = 24;  // src/main.rs:45
enum : uint32_t { TAG_DIFFUSE = 0, TAG_DIFFUSECBOARD = 1, TAG_MIRROR = 2, TAG_DIELECTRIC = 3, TAG_EMISSIVE = 4 };

// ------------------------------------------------------------------ rand.glsl
struct Rng {
  uint32_t rngState;
  uint32_t randUint() {  // rand.glsl:2-7
    rngState ^= (rngState << 13);
    rngState ^= (rngState >> 17);
    rngState ^= (rngState << 5);
    return rngState;
  }
  void seedRng(uint32_t seed) {  // rand.glsl:9-16 (Wang hash)
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    rngState = seed;
  }
  float randUniformFloat() {  // rand.glsl:18-20
    return (float)randUint() * (1.0f / 4294967296.0f);
  }
  vec3 randCosHemisphere() {  // rand.glsl:22-30
    float u = randUniformFloat();
    float v = randUniformFloat();
    float r = sqrtf(u);
    float theta = (2.0f * M_PI_F) * v;
    float x = r * orc_cosf(theta);
    float y = r * orc_sinf(theta);
    return V(x, y, sqrtf(glsl_max(0.0f, 1.0f - u)));
  }
  vec3 randUniformSphere() {  // rand.glsl:32-40
    float u = randUniformFloat();
    float v = randUniformFloat();
    float z = 2.0f * u - 1.0f;
    float theta = (2.0f * M_PI_F) * v;
    float r = sqrtf(1.0f - z * z);
    return V(r * orc_cosf(theta), r * orc_sinf(theta), z);
  }
  vec3 randBarycentric() {  // rand.glsl:42-50 (the fold is kept as written)
    float u = randUniformFloat();
    float v = randUniformFloat();
    if (u + v > 1.0f) {
      u = 1.0f - v;
      v = 1.0f - u;
    }
    return V(u, v, 1.0f - u - v);
  }
};

// ------------------------------------------------------------------ quaternion.glsl
struct vec4 {
  float x, y, z, w;
};
inline vec3 xyz(vec4 q) { return V(q.x, q.y, q.z); }
vec4 quaternionMult(vec4 qa, vec4 qb) {  // quaternion.glsl:1-6
  vec4 r;
  r.w = qa.w * qb.w - dot(xyz(qa), xyz(qb));
  vec3 v = cross(xyz(qa), xyz(qb)) + xyz(qa) * qb.w + xyz(qb) * qa.w;
  r.x = v.x;
  r.y = v.y;
  r.z = v.z;
  return r;
}
vec3 quaternionRotate(vec3 v, vec4 r) {  // quaternion.glsl:15-19
  vec4 tmp = quaternionMult(r, vec4{v.x, v.y, v.z, 0.0f});
  r.x = -r.x;
  r.y = -r.y;
  r.z = -r.z;
  return xyz(quaternionMult(tmp, r));
}

// ------------------------------------------------------------------ render.glsl structs
struct Ray {  // render.glsl:19-24
  vec3 origin;
  vec3 direction;
  float tMin;
  float tMax;
};
struct Intersection {  // render.glsl:39-46
  int objectID;
  float t;
  vec3 p;
  vec3 n;
  float uvx, uvy;
  mat3 frame;
};
struct ShapeQueryRecord {  // render.glsl:48-52
  vec3 p;
  vec3 n;
  float pdf;
};

struct Ctx {
  SceneView s;
  float M_EPS;
  int useBvh;
  Rng rng;
  uint64_t nExt = 0, nShadow = 0;
};

Ray getCameraRayAt(const Camera& c, float xx, float xy, float dimx, float dimy, float M_EPS) {
  // render.glsl:26-36
  xx = xx - 0.5f * dimx;
  xy = xy - 0.5f * dimy;
  float radians = (0.5f * c.fov) * (M_PI_F / 180.0f);
  float tn = orc_tanf(radians);
  xx = xx * tn / (0.5f * dimx);
  xy = xy * tn / (0.5f * dimx);
  Ray res;
  res.origin = V(c.position[0], c.position[1], c.position[2]);
  vec4 rot{c.rotation[0], c.rotation[1], c.rotation[2], c.rotation[3]};
  res.direction = normalize(quaternionRotate(V(xx, -xy, -1.0f), rot));
  res.tMin = M_EPS;
  res.tMax = (float)1e100;  // +inf in fp32
  return res;
}

// ------------------------------------------------------------------ shapes/triangle.glsl
inline vec3 vpos(const Vertex& v) { return V(v.pos_u[0], v.pos_u[1], v.pos_u[2]); }
inline vec3 vnrm(const Vertex& v) { return V(v.norm_v[0], v.norm_v[1], v.norm_v[2]); }

bool intersectTriangle(const SceneView& s, const Ray& ray, uint32_t ix, Intersection& its) {
  // shapes/triangle.glsl:15-52
  const Vertex& a = s.vertices[s.triangles[3 * ix + 0]];
  const Vertex& b = s.vertices[s.triangles[3 * ix + 1]];
  const Vertex& c = s.vertices[s.triangles[3 * ix + 2]];
  vec3 ab = vpos(b) - vpos(a);
  vec3 ac = vpos(c) - vpos(a);
  vec3 n = cross(ab, ac);
  vec3 ro = ray.origin - vpos(a);
  vec3 q = cross(ro, ray.direction);
  float d = 1.0f / dot(ray.direction, n);
  float u = d * dot(-q, ac);
  float v = d * dot(q, ab);
  if (u < 0.0f || v < 0.0f || u + v > 1.0f) return false;
  float t = d * dot(-n, ro);
  if (ray.tMin <= t && t <= ray.tMax) {
    its.t = t;
    its.uvx = u;
    its.uvy = v;
    its.n = normalize(n);
    return true;
  }
  return false;
}

void populateTriangleIntersection(const SceneView& s, uint32_t ix, Intersection& its) {
  // shapes/triangle.glsl:54-78
  float l0 = 1.0f - its.uvx - its.uvy, l1 = its.uvx, l2 = its.uvy;
  const Vertex& a = s.vertices[s.triangles[3 * ix + 0]];
  const Vertex& b = s.vertices[s.triangles[3 * ix + 1]];
  const Vertex& c = s.vertices[s.triangles[3 * ix + 2]];
  its.n = normalize(vnrm(a) * l0 + vnrm(b) * l1 + vnrm(c) * l2);
  float uvx = a.pos_u[3] * l0 + b.pos_u[3] * l1 + c.pos_u[3] * l2;
  float uvy = a.norm_v[3] * l0 + b.norm_v[3] * l1 + c.norm_v[3] * l2;
  its.uvx = uvx;
  its.uvy = uvy;
  vec3 t, bt;
  if (fabsf(its.n.x) > fabsf(its.n.y)) {
    bt = V(0.f, 1.f, 0.f);
  } else {
    bt = V(1.f, 0.f, 0.f);
  }
  t = normalize(cross(its.n, bt));
  bt = cross(its.n, t);
  its.frame = mat3{t, bt, its.n};
}

void sampleTriangle(Ctx& c, uint32_t ix, ShapeQueryRecord& sRec) {
  // shapes/triangle.glsl:81-102
  const SceneView& s = c.s;
  const Vertex& a = s.vertices[s.triangles[3 * ix + 0]];
  const Vertex& b = s.vertices[s.triangles[3 * ix + 1]];
  const Vertex& cc = s.vertices[s.triangles[3 * ix + 2]];
  vec3 ab = vpos(b) - vpos(a);
  vec3 ac = vpos(cc) - vpos(a);
  vec3 n = cross(ab, ac);
  float area = length(n) / 2.0f;
  vec3 lambda = c.rng.randBarycentric();
  sRec.n = normalize(vnrm(a) * lambda.x + vnrm(b) * lambda.y + vnrm(cc) * lambda.z);
  sRec.p = vpos(a) * lambda.x + vpos(b) * lambda.y + vpos(cc) * lambda.z;
  sRec.pdf = 1.0f / area;
}

// ------------------------------------------------------------------ shapes/sphere.glsl
bool intersectSphere(const Ray& ray, const Sphere& sphere, Intersection& its) {
  // shapes/sphere.glsl:18-41
  vec3 pos = V(sphere.positionRadius[0], sphere.positionRadius[1], sphere.positionRadius[2]);
  float r = sphere.positionRadius[3];
  vec3 l = ray.origin - pos;
  float b = 2.0f * dot(ray.direction, l);
  float c = dot(l, l) - r * r;
  float d = b * b - 4.0f * c;
  if (d < 0.0f) return false;
  d = sqrtf(d);
  float t0 = -0.5f * (b + d);
  if (ray.tMin <= t0 && t0 <= ray.tMax) {
    its.t = t0;
    return true;
  }
  float t1 = -0.5f * (b - d);
  if (ray.tMin <= t1 && t1 <= ray.tMax) {
    its.t = t1;
    return true;
  }
  return false;
}

void populateSphereIntersection(const Sphere& sphere, Intersection& its) {
  // shapes/sphere.glsl:43-52
  vec3 pos = V(sphere.positionRadius[0], sphere.positionRadius[1], sphere.positionRadius[2]);
  vec3 n = its.n = (its.p - pos) / sphere.positionRadius[3];
  vec3 t = normalize(V(-n.z, 0.f, n.x));
  vec3 b = cross(n, t);
  its.frame = mat3{t, b, n};
  its.uvx = 0.5f + orc_atan2f(n.z, n.x) / (2.0f * M_PI_F);
  its.uvy = 0.5f + orc_asinf(glsl_min(glsl_max(n.y, -1.0f), 1.0f)) / M_PI_F;
  if (std::isnan(its.uvx)) its.uvx = 0.f;
}

void sampleSphere(Ctx& c, const Sphere& sphere, ShapeQueryRecord& sRec) {
  // shapes/sphere.glsl:54-58
  vec3 pos = V(sphere.positionRadius[0], sphere.positionRadius[1], sphere.positionRadius[2]);
  float r = sphere.positionRadius[3];
  sRec.n = c.rng.randUniformSphere();
  sRec.p = pos + r * sRec.n;
  sRec.pdf = 1.0f / (r * r * 4.0f * M_PI_F);
}

// ------------------------------------------------------------------ shapes/quad.glsl
inline vec3 q3(const float* p) { return V(p[0], p[1], p[2]); }
bool intersectQuad(const Ray& ray, const Quad& quad, Intersection& its) {
  // shapes/quad.glsl:7-25
  vec3 n = cross(q3(quad.edge1), q3(quad.edge2));
  vec3 ro = ray.origin - q3(quad.origin);
  vec3 q = cross(ro, ray.direction);
  float d = 1.0f / dot(ray.direction, n);
  float u = d * dot(-q, q3(quad.edge2));
  float v = d * dot(q, q3(quad.edge1));
  if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return false;
  float t = d * dot(-n, ro);
  if (ray.tMin <= t && t <= ray.tMax) {
    its.t = t;
    its.uvx = u;
    its.uvy = v;
    return true;
  }
  return false;
}
void populateQuadIntersection(const Quad& quad, Intersection& its) {
  // shapes/quad.glsl:27-32
  vec3 t = normalize(q3(quad.edge1));
  vec3 b = normalize(q3(quad.edge2));
  vec3 n = its.n = cross(t, b);
  its.frame = mat3{t, b, n};
}
void sampleQuad(Ctx& c, const Quad& quad, ShapeQueryRecord& sRec) {
  // shapes/quad.glsl:34-45
  vec3 n = cross(q3(quad.edge1), q3(quad.edge2));
  float area = length(n);
  n = n / area;
  sRec.n = n;
  float u = c.rng.randUniformFloat();
  float v = c.rng.randUniformFloat();
  sRec.p = q3(quad.origin) + u * q3(quad.edge1) + v * q3(quad.edge2);
  sRec.pdf = 1.0f / area;
}

// ------------------------------------------------------------------ scene.glsl
bool intersectShape(const SceneView& s, const Ray& ray, uint32_t shapeIndex, Intersection& its) {
  // scene.glsl:107-114
  if ((int)shapeIndex < s.numSpheres) return intersectSphere(ray, s.spheres[shapeIndex], its);
  if ((int)shapeIndex < s.numSpheres + s.numQuads)
    return intersectQuad(ray, s.quads[shapeIndex - s.numSpheres], its);
  return intersectTriangle(s, ray, shapeIndex - s.numSpheres - s.numQuads, its);
}

// ------------------------------------------------------------------ mode 3: culled linear scan (oracle extension)
// The reference's linear scan (scene.glsl:134-157, mode 2 = without the >100 failsafe) costs one primitive test
// per primitive per ray: minutes for one 1080p pass of cbox, hours for the 10 M-triangle terrain.  Mode 3 runs
// THE SAME SCAN — ascending shape index, tMax = t - M_EPS after every accepted hit — over a subset of the
// primitives: those whose bounding box, padded, the ray's supporting half-line pierces.  A primitive outside
// that subset cannot be accepted by its test (an accepted hit lies on the primitive, hence inside its box), so
// dropping it changes nothing: same winner, same t, same order dependence among ties.  The subset comes from
// an oracle-private median-split box tree over the primitives' own boxes; the slab test runs in double
// precision against boxes padded by 1e-4 of the scene scale, and ignores the ray's tMin/tMax altogether.
// Two cases are kept out of the culling because the reference's arithmetic is not geometric there:
//   * spheres are only culled for directions of unit length within 1e-4 and origins inside the padded scene
//     bounds; their boxes carry the extra radius R' - r of shapes/sphere.glsl's non-unit-direction behaviour
//     (see traverse.cuh "SPHERE GUARD" for the algebra); any other ray tests every sphere;
//   * a direction exactly perpendicular to a triangle's normal makes triangle.glsl:24 divide by zero and can
//     report t = +inf "hits": such artefact hits (never produced by integrator rays) are not reproduced.
// tests/test_oracle_cull.py holds mode 3 identical to modes 0/2 (ids, t, uv, tie flags, whole frames).
struct CullAccel {
  struct Node {
    double lo[3], hi[3];
    uint32_t left, right;   // children (inner) ...
    uint32_t first, count;  // ... or a range of `order` (leaf: count > 0)
  };
  std::vector<Node> nodes;
  std::vector<uint32_t> order;  // global shape ids
  double lo[3], hi[3];          // padded bounds of every primitive and the camera
  int numSpheres = 0;
};

struct PrimBox {
  double lo[3], hi[3], c[3];
};

inline void box_grow(double* lo, double* hi, double x, double y, double z) {
  const double p[3] = {x, y, z};
  for (int k = 0; k < 3; k++) {
    if (p[k] < lo[k]) lo[k] = p[k];
    if (p[k] > hi[k]) hi[k] = p[k];
  }
}

std::shared_ptr<CullAccel> build_cull_accel(const SceneView& s) {
  auto acc = std::make_shared<CullAccel>();
  const int total = s.numSpheres + s.numQuads + s.numTriangles;
  acc->numSpheres = s.numSpheres;
  std::vector<PrimBox> boxes((size_t)total);
  double slo[3] = {1e300, 1e300, 1e300}, shi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < total; i++) {
    PrimBox& b = boxes[(size_t)i];
    for (int k = 0; k < 3; k++) b.lo[k] = 1e300, b.hi[k] = -1e300;
    if (i < s.numSpheres) {
      const Sphere& sp = s.spheres[i];
      const double r = std::fabs((double)sp.positionRadius[3]);
      box_grow(b.lo, b.hi, sp.positionRadius[0] - r, sp.positionRadius[1] - r, sp.positionRadius[2] - r);
      box_grow(b.lo, b.hi, sp.positionRadius[0] + r, sp.positionRadius[1] + r, sp.positionRadius[2] + r);
    } else if (i < s.numSpheres + s.numQuads) {
      const Quad& q = s.quads[i - s.numSpheres];
      for (int a = 0; a < 2; a++)
        for (int c = 0; c < 2; c++)
          box_grow(b.lo, b.hi, (double)q.origin[0] + a * (double)q.edge1[0] + c * (double)q.edge2[0],
                   (double)q.origin[1] + a * (double)q.edge1[1] + c * (double)q.edge2[1],
                   (double)q.origin[2] + a * (double)q.edge1[2] + c * (double)q.edge2[2]);
    } else {
      const uint32_t t = (uint32_t)(i - s.numSpheres - s.numQuads);
      for (int v = 0; v < 3; v++) {
        const Vertex& vx = s.vertices[s.triangles[3 * t + v]];
        box_grow(b.lo, b.hi, vx.pos_u[0], vx.pos_u[1], vx.pos_u[2]);
      }
    }
    for (int k = 0; k < 3; k++) {
      b.c[k] = 0.5 * (b.lo[k] + b.hi[k]);
      if (b.lo[k] < slo[k]) slo[k] = b.lo[k];
      if (b.hi[k] > shi[k]) shi[k] = b.hi[k];
    }
  }
  box_grow(slo, shi, s.info->camera.position[0], s.info->camera.position[1], s.info->camera.position[2]);
  double scale = 0.0;
  for (int k = 0; k < 3; k++) scale = std::max(scale, std::max(std::fabs(slo[k]), std::fabs(shi[k])));
  const double pad = 1e-4 * scale + 1e-7;
  double diag2 = 0.0;
  for (int k = 0; k < 3; k++) {
    acc->lo[k] = slo[k] - 2 * pad, acc->hi[k] = shi[k] + 2 * pad;
    diag2 += (acc->hi[k] - acc->lo[k]) * (acc->hi[k] - acc->lo[k]);
  }
  for (int i = 0; i < total; i++) {
    double extra = pad;
    if (i < s.numSpheres) {  // R' - r <= eps (r^2 + L^2) / (2 r) for |s^2 - 1| <= eps = 1e-4, L <= scene diagonal; x 2
      const double r = std::max(std::fabs((double)s.spheres[i].positionRadius[3]), 1e-30);
      extra += 2.0 * 1e-4 * (r * r + diag2) / (2.0 * r);
    }
    for (int k = 0; k < 3; k++) boxes[(size_t)i].lo[k] -= extra, boxes[(size_t)i].hi[k] += extra;
  }
  acc->order.resize((size_t)total);
  for (int i = 0; i < total; i++) acc->order[(size_t)i] = (uint32_t)i;
  // median split on the longest axis of the centroid bounds, leaves of <= 4 primitives; the big ranges near the
  // root are split on their own threads (node slots come from an atomic counter, so the tree does not depend on
  // scheduling in anything but node numbering, which nothing reads)
  acc->nodes.resize((size_t)std::max(total, 1) * 2);
  std::atomic<uint32_t> next_node{1};
  std::function<void(uint32_t, uint32_t, uint32_t, int)> build = [&](uint32_t node, uint32_t first, uint32_t count,
                                                                      int par_depth) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t k = first; k < first + count; k++) {
      const PrimBox& b = boxes[acc->order[k]];
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], b.lo[a]), hi[a] = std::max(hi[a], b.hi[a]);
        clo[a] = std::min(clo[a], b.c[a]), chi[a] = std::max(chi[a], b.c[a]);
      }
    }
    CullAccel::Node n{};
    for (int a = 0; a < 3; a++) n.lo[a] = lo[a], n.hi[a] = hi[a];
    if (count <= 4) {
      n.first = first, n.count = count;
      acc->nodes[node] = n;
      return;
    }
    int axis = 0;
    for (int a = 1; a < 3; a++)
      if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
    const uint32_t half = count / 2;
    std::nth_element(acc->order.begin() + first, acc->order.begin() + first + half, acc->order.begin() + first + count,
                     [&](uint32_t x, uint32_t y) {
                       const double cx = boxes[x].c[axis], cy = boxes[y].c[axis];
                       return cx < cy || (cx == cy && x < y);
                     });
    n.left = next_node.fetch_add(2);
    n.right = n.left + 1;
    n.count = 0;
    acc->nodes[node] = n;
    if (par_depth > 0 && count > (1u << 16)) {
      std::thread t(build, n.left, first, half, par_depth - 1);
      build(n.right, first + half, count - half, par_depth - 1);
      t.join();
    } else {
      build(n.left, first, half, 0);
      build(n.right, first + half, count - half, 0);
    }
  };
  if (total > 0) build(0u, 0u, (uint32_t)total, 4);
  acc->nodes.resize(next_node.load());
  return acc;
}

// shape ids (ascending) the mode-3 scan has to test for this ray
void cull_candidates(const CullAccel& a, const Ray& ray, std::vector<uint32_t>& out) {
  out.clear();
  const double o[3] = {ray.origin.x, ray.origin.y, ray.origin.z};
  const double d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
  const double s2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  bool cull_spheres = std::fabs(s2 - 1.0) <= 1e-4;
  for (int k = 0; k < 3; k++)
    if (!(o[k] >= a.lo[k] && o[k] <= a.hi[k])) cull_spheres = false;
  if (!cull_spheres)
    for (int i = 0; i < a.numSpheres; i++) out.push_back((uint32_t)i);
  const double t_lo = ray.tMin < 0.f ? -1e300 : 0.0;  // the half-line ahead of the origin (the whole line if tMin < 0)
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const CullAccel::Node& n = a.nodes[stack[--sp]];
    double t0 = t_lo, t1 = 1e300;
    bool miss = false;
    for (int k = 0; k < 3 && !miss; k++) {
      if (d[k] == 0.0) {
        if (o[k] < n.lo[k] || o[k] > n.hi[k]) miss = true;
      } else {
        double ta = (n.lo[k] - o[k]) / d[k], tb = (n.hi[k] - o[k]) / d[k];
        if (ta > tb) std::swap(ta, tb);
        // one part in 1e9 of slack: the divisions round
        ta -= 1e-9 * std::fabs(ta), tb += 1e-9 * std::fabs(tb);
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) miss = true;
      }
    }
    if (miss) continue;
    if (n.count) {
      for (uint32_t k = n.first; k < n.first + n.count; k++) {
        const uint32_t id = a.order[k];
        if (!cull_spheres && (int)id < a.numSpheres) continue;  // already listed
        out.push_back(id);
      }
    } else if (sp + 2 <= 128) {
      stack[sp++] = n.left;
      stack[sp++] = n.right;
    }
  }
  std::sort(out.begin(), out.end());
}

// mode: 0 = USE_BVH 0 (linear scan incl. the >100 failsafe), 1 = USE_BVH 1, 2 = linear scan
// without the failsafe (oracle extension for scenes the reference refuses), 3 = the scan of mode 2 over
// the primitives whose box the ray pierces (CullAccel above; same results, tractable at BASELINE sizes)
bool intersectScene(const SceneView& s, int useBvh, float M_EPS, Ray ray, Intersection& its) {
  // scene.glsl:97-175
  its.objectID = -1;
  if (useBvh == 1) {
    // scene.glsl:100-133 (USE_BVH == 1)
    vec3 invRayDir = V(1.0f) / ray.direction;
    vec3 timeOffset = -ray.origin * invRayDir;
    for (uint32_t currentNode = 0; currentNode < s.bvhLength;) {
      const BVHNode& node = s.bvh[currentNode];
      uint32_t shapeIndex = node.shapeIndex;
      uint32_t exitIndex = node.exitIndex;
      if (shapeIndex != 0xFFFFFFFFu) {
        if (intersectShape(s, ray, shapeIndex, its)) {
          ray.tMax = its.t - M_EPS;
          its.objectID = (int)shapeIndex;
        }
        currentNode = exitIndex;
      } else {
        vec3 tNegative = q3(node.aabbMin) * invRayDir + timeOffset;
        vec3 tPositive = q3(node.aabbMax) * invRayDir + timeOffset;
        vec3 tMin = V(glsl_min(tNegative.x, tPositive.x), glsl_min(tNegative.y, tPositive.y),
                      glsl_min(tNegative.z, tPositive.z));
        vec3 tMax = V(glsl_max(tNegative.x, tPositive.x), glsl_max(tNegative.y, tPositive.y),
                      glsl_max(tNegative.z, tPositive.z));
        float t0 = glsl_max(glsl_max(tMin.x, tMin.y), tMin.z);
        float t1 = glsl_min(glsl_min(tMax.x, tMax.y), tMax.z);
        if (t0 < t1 + M_EPS && t0 < ray.tMax && t1 > ray.tMin) {
          currentNode = currentNode + 1;
        } else {
          currentNode = exitIndex;
        }
      }
    }
  } else if (useBvh == 3) {
    // scene.glsl:139-157 over the culled subset, in the same (ascending shape index) order
    thread_local std::vector<uint32_t> cands;
    cull_candidates(*s.accel, ray, cands);
    for (uint32_t id : cands) {
      if (intersectShape(s, ray, id, its)) {
        ray.tMax = its.t - M_EPS;
        its.objectID = (int)id;
      }
    }
  } else {
    // scene.glsl:134-157 (USE_BVH == 0, the reference default)
    if (useBvh == 0 && (s.numSpheres > 100 || s.numQuads > 100)) return false;  // "failsafe", scene.glsl:135-138
    for (int i = 0; i < s.numSpheres; i++) {
      if (intersectSphere(ray, s.spheres[i], its)) {
        ray.tMax = its.t - M_EPS;
        its.objectID = i;
      }
    }
    for (int i = 0; i < s.numQuads; i++) {
      if (intersectQuad(ray, s.quads[i], its)) {
        ray.tMax = its.t - M_EPS;
        its.objectID = s.numSpheres + i;
      }
    }
    for (int i = 0; i < s.numTriangles; i++) {
      if (intersectTriangle(s, ray, (uint32_t)i, its)) {
        ray.tMax = its.t - M_EPS;
        its.objectID = s.numSpheres + s.numQuads + i;
      }
    }
  }
  if (its.objectID == -1) return false;
  its.p = ray.origin + its.t * ray.direction;  // scene.glsl:164
  if (its.objectID < s.numSpheres) {
    populateSphereIntersection(s.spheres[its.objectID], its);
  } else if (its.objectID < s.numSpheres + s.numQuads) {
    populateQuadIntersection(s.quads[its.objectID - s.numSpheres], its);
  } else {
    populateTriangleIntersection(s, (uint32_t)(its.objectID - s.numSpheres - s.numQuads), its);
  }
  return true;
}

void sampleShape(Ctx& c, uint32_t shape, ShapeQueryRecord& sRec) {
  // scene.glsl:44-52
  const SceneView& s = c.s;
  if ((int)shape < s.numSpheres) {
    sampleSphere(c, s.spheres[shape], sRec);
  } else if ((int)shape < s.numSpheres + s.numQuads) {
    sampleQuad(c, s.quads[shape - s.numSpheres], sRec);
  } else {
    sampleTriangle(c, shape - s.numSpheres - s.numQuads, sRec);
  }
}

vec3 sampleEmitter(Ctx& c, vec3 ref, Ray& shadowRay) {
  // scene.glsl:54-89
  const SceneView& s = c.s;
  float emitterSample = c.rng.randUniformFloat();
  int emitter = 0;
  for (int i = 0; i < s.numEmitters; i++) {
    emitterSample -= s.emitters[i].pdf;
    if (emitterSample < 0.f) {
      emitter = i;
      break;
    }
  }
  ShapeQueryRecord sRec;
  sampleShape(c, s.emitters[emitter].shape, sRec);
  uint32_t mat = s.materials[s.emitters[emitter].shape];
  const float* pw = s.emissiveMaterials[mat & ((1u << MATERIAL_TAG_SHIFT) - 1u)].v;
  vec3 power = V(pw[0], pw[1], pw[2]);
  vec3 dir = sRec.p - ref;
  float dist = length(dir);
  dir = dir / dist;
  shadowRay.origin = ref;
  shadowRay.direction = dir;
  shadowRay.tMin = 2.0f * c.M_EPS;
  shadowRay.tMax = dist - c.M_EPS;
  float cosTheta = -dot(dir, sRec.n);
  if (cosTheta < 0.f) return V(0.f);
  float pdf = s.emitters[emitter].pdf * sRec.pdf * dist * dist / cosTheta;
  return power / pdf;
}

// ------------------------------------------------------------------ material.glsl
vec3 getCheckerboardTexture(const Vec4* mat2, float uvx, float uvy) {
  // materials/diffusecb.glsl:6-13
  const float* a = mat2[0].v;  // color_a_scale_u
  const float* b = mat2[1].v;  // color_b_scale_v
  float ux = fract(0.5f * uvx / a[3]);
  float uy = fract(0.5f * uvy / b[3]);
  if ((ux < 0.5f) != (uy < 0.5f)) return V(b[0], b[1], b[2]);
  return V(a[0], a[1], a[2]);
}

vec3 evalBSDF(const SceneView& s, uint32_t material, vec3 wi, const Intersection& its) {
  // material.glsl:18-30
  uint32_t tag = material >> MATERIAL_TAG_SHIFT;
  uint32_t idx = material & ((1u << MATERIAL_TAG_SHIFT) - 1u);
  if (tag == TAG_DIFFUSE) {
    const float* cl = s.diffuseMaterials[idx].v;
    return dot(its.n, wi) * V(cl[0], cl[1], cl[2]) / M_PI_F;
  } else if (tag == TAG_DIFFUSECBOARD) {
    vec3 color = getCheckerboardTexture(&s.diffuseCBMaterials[2 * idx], its.uvx, its.uvy);
    return dot(its.n, wi) * color / M_PI_F;
  }
  return V(0.f);
}

// returns false when `wo` is left unwritten (emissive, material.glsl:88-89)
bool sampleBSDF(Ctx& c, uint32_t material, vec3 wi, const Intersection& its, vec3& wo,
                vec3& extinction, vec3& weight) {
  // material.glsl:33-91
  const SceneView& s = c.s;
  uint32_t tag = material >> MATERIAL_TAG_SHIFT;
  uint32_t idx = material & ((1u << MATERIAL_TAG_SHIFT) - 1u);
  switch (tag) {
    case TAG_DIFFUSE: {
      vec3 wo_local = c.rng.randCosHemisphere();
      wo = mul(its.frame, wo_local);
      const float* cl = s.diffuseMaterials[idx].v;
      weight = V(cl[0], cl[1], cl[2]);
      return true;
    }
    case TAG_DIFFUSECBOARD: {
      vec3 wo_local = c.rng.randCosHemisphere();
      wo = mul(its.frame, wo_local);
      weight = getCheckerboardTexture(&s.diffuseCBMaterials[2 * idx], its.uvx, its.uvy);
      return true;
    }
    case TAG_MIRROR:
      wo = reflect(wi, its.n);
      weight = V(1.f);
      return true;
    case TAG_DIELECTRIC: {
      const float* de = s.dielectricMaterials[idx].v;
      float eta = de[3];
      float etaInv = 1.0f / eta;
      float cosThetaI = -dot(its.n, wi);
      vec3 normal = its.n;
      bool isInsideDielectric = cosThetaI > 0.f;
      if (cosThetaI < 0.f) {
        eta = etaInv;
        etaInv = 1.0f / eta;
        normal = -normal;
        cosThetaI = -cosThetaI;
      }
      float k = 1.0f - etaInv * etaInv * (1.0f - cosThetaI * cosThetaI);
      if (k <= 0.f) {
        wo = reflect(wi, normal);
      } else {
        float cosThetaO = sqrtf(k);
        float rho_par = (eta * cosThetaI - cosThetaO) / (eta * cosThetaI + cosThetaO);
        float rho_orth = (cosThetaI - eta * cosThetaO) / (cosThetaI + eta * cosThetaO);
        float f_r = 0.5f * (rho_par * rho_par + rho_orth * rho_orth);
        if (c.rng.randUniformFloat() < f_r) {
          wo = reflect(wi, normal);
        } else {
          isInsideDielectric = !isInsideDielectric;
          vec3 parallel = wi - dot(wi, normal) * normal;
          wo = etaInv * parallel - sqrtf(k) * normal;
        }
      }
      if (isInsideDielectric) extinction = V(de[0], de[1], de[2]);
      weight = V(1.f);
      return true;
    }
    case TAG_EMISSIVE:
    default:
      weight = V(0.f);
      return false;
  }
}

// ------------------------------------------------------------------ render.glsl integrateRay
struct PathOut {
  vec3 total, albedo, normal;
  float depth;
};

void integrateRay(Ctx& c, Ray ray, uint32_t maxBounces, uint32_t rrStart, PathOut& o,
                  OrcPathVertex* log, int logCap, int* logN) {
  // render.glsl:81-147
  vec3 currentExtinction = V(0.f);
  o.total = V(0.f);
  o.albedo = V(0.f);
  o.depth = 0.f;
  o.normal = V(0.f);
  vec3 throughput = V(1.f);
  bool wasDiscrete = true;
  Intersection its;
  for (uint32_t bounce = 0; bounce < maxBounces; bounce++) {
    c.nExt++;
    const float dirLen2 = dot(ray.direction, ray.direction);
    if (!intersectScene(c.s, c.useBvh, c.M_EPS, ray, its)) {
      if (log && *logN < logCap) {
        OrcPathVertex& pv = log[(*logN)++];
        pv.dir_len2 = dirLen2;
        pv.shape_id = -1;
        pv.t = 0.f;
        pv.rng_after = c.rng.rngState;
        pv.throughput[0] = throughput.x, pv.throughput[1] = throughput.y, pv.throughput[2] = throughput.z;
        pv.total[0] = o.total.x, pv.total[1] = o.total.y, pv.total[2] = o.total.z;
        pv.shadow_state = 0;
      }
      return;
    }
    if (bounce == 0) {
      o.depth = its.t;
      o.normal = its.n;
    }
    uint32_t mat = c.s.materials[its.objectID];
    uint32_t material_tag = mat >> MATERIAL_TAG_SHIFT;
    uint32_t material_idx = mat & ((1u << MATERIAL_TAG_SHIFT) - 1u);

    float dist = length(ray.origin - its.p);
    throughput = throughput * vexp(-currentExtinction * dist);

    if (material_tag == TAG_EMISSIVE && wasDiscrete) {
      const float* pw = c.s.emissiveMaterials[material_idx].v;
      o.total = o.total + throughput * V(pw[0], pw[1], pw[2]);
    }
    int shadowState = 0;
    if (material_tag == TAG_DIFFUSE || material_tag == TAG_DIFFUSECBOARD) {
      Ray shadowRay;
      vec3 importance = sampleEmitter(c, its.p, shadowRay);
      if (length(importance) > c.M_EPS && dot(shadowRay.direction, its.n) > 0.f) {
        c.nShadow++;
        Intersection dummy;
        if (!intersectScene(c.s, c.useBvh, c.M_EPS, shadowRay, dummy)) {
          o.total = o.total + throughput * evalBSDF(c.s, mat, shadowRay.direction, its) * importance;
          shadowState = 2;
        } else {
          shadowState = 1;
        }
      }
    }
    vec3 wo = V(0.f), weight;
    bool woWritten = sampleBSDF(c, mat, ray.direction, its, wo, currentExtinction, weight);
    throughput = throughput * weight;
    ray.direction = wo;
    ray.origin = its.p;
    ray.tMin = 2.0f * c.M_EPS;
    ray.tMax = (float)1e100;
    wasDiscrete = material_tag != TAG_DIFFUSE && material_tag != TAG_DIFFUSECBOARD;

    bool terminate = false;
    if (bounce > rrStart) {
      float q = glsl_min(0.99f, glsl_max(throughput.x, glsl_max(throughput.y, throughput.z)));
      if (c.rng.randUniformFloat() > q) {
        terminate = true;
      } else {
        throughput = throughput / q;
      }
    }
    if (log && *logN < logCap) {
      OrcPathVertex& pv = log[(*logN)++];
      pv.shape_id = its.objectID;
      pv.t = its.t;
      pv.rng_after = c.rng.rngState;
      pv.throughput[0] = throughput.x, pv.throughput[1] = throughput.y, pv.throughput[2] = throughput.z;
      pv.total[0] = o.total.x, pv.total[1] = o.total.y, pv.total[2] = o.total.z;
      pv.shadow_state = shadowState;
      pv.dir_len2 = dirLen2;
    }
    if (terminate) break;
    // after an emissive hit `wo` is undefined and throughput is 0; the reference
    // keeps tracing garbage rays that can never contribute.  Terminating is result-equivalent.
    if (!woWritten) break;
  }
}

// render.glsl:149-175 for one pixel of one block
bool renderPixel(Ctx& c, const OrcBlock& blk, uint32_t lx, uint32_t ly, const OrcParams& p,
                 PathOut& o, OrcPathVertex* log = nullptr, int logCap = 0, int* logN = nullptr) {
  if (lx >= blk.original_dimension[0] || ly >= blk.original_dimension[1]) return false;  // :152
  uint32_t gx = lx + blk.origin[0], gy = ly + blk.origin[1];
  uint32_t seed = blk.seed + lx + ly * blk.dimension[0];  // :156
  c.rng.seedRng(seed);
  Ray ray = getCameraRayAt(c.s.info->camera, (float)gx + blk.sample_offset[0],
                           (float)gy + blk.sample_offset[1], (float)blk.original_dimension[0],
                           (float)blk.original_dimension[1], c.M_EPS);
  integrateRay(c, ray, p.max_bounces, p.rr_start, o, log, logCap, logN);
  return true;
}

void parallel_for(uint64_t n, int n_threads, const std::function<void(uint64_t, uint64_t, int)>& fn) {
  if (n_threads <= 1 || n < 2) {
    fn(0, n, 0);
    return;
  }
  std::atomic<uint64_t> next{0};
  uint64_t chunk = n / ((uint64_t)n_threads * 16) + 1;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    th.emplace_back([&, t]() {
      for (;;) {
        uint64_t b = next.fetch_add(chunk);
        if (b >= n) break;
        uint64_t e = b + chunk < n ? b + chunk : n;
        fn(b, e, t);
      }
    });
  }
  for (auto& x : th) x.join();
}

inline float* px4(float* base, uint32_t w, uint32_t x, uint32_t y) { return base + 4 * ((uint64_t)y * w + x); }

// reconstruction.glsl:22-66 for ONE block.  `sample(layer, sx, sy, out)` reads the block's
// intermediate texture (zeros when out of bounds of the block_size^2 texture).
template <class SampleFn>
void reconstructBlock(const OrcBlock& blk, const OrcParams& p, SampleFn sample, float* acc,
                      int n_threads) {
  const int R = (int)p.recon_radius;
  const float gaussFac = -1.0f / (2.0f * p.recon_stddev * p.recon_stddev);
  const float curveOffset = orc_expf(gaussFac * (float)R * (float)R);
  const uint32_t W = blk.original_dimension[0], H = blk.original_dimension[1];
  const uint32_t ext_x = blk.dimension[0] + 2 * R, ext_y = blk.dimension[1] + 2 * R;
  parallel_for(ext_y, n_threads, [&](uint64_t y0, uint64_t y1, int) {
    for (uint32_t gidy = (uint32_t)y0; gidy < (uint32_t)y1; gidy++) {
      for (uint32_t gidx = 0; gidx < ext_x; gidx++) {
        uint32_t localx = gidx - (uint32_t)R, localy = gidy - (uint32_t)R;  // uvec2, wraps (:23)
        uint32_t globalx = localx + blk.origin[0], globaly = localy + blk.origin[1];
        bool in_image = (int32_t)globalx >= 0 && (int32_t)globaly >= 0 && globalx < W && globaly < H;
        float outv[4] = {0.f, 0.f, 0.f, 0.f};
        if (in_image) memcpy(outv, px4(acc, W, globalx, globaly), 16);  // imageLoad(outputImage) :26
        float nc[4], ac[4];
        sample(1, (int32_t)localx, (int32_t)localy, nc);  // normalCenter :32
        sample(2, (int32_t)localx, (int32_t)localy, ac);  // albedoCenter :33
        for (int dx = -R; dx <= R; dx++) {
          if (localx + (uint32_t)dx >= blk.dimension[0]) continue;  // :36 (uint compare)
          for (int dy = -R; dy <= R; dy++) {
            if (localy + (uint32_t)dy >= blk.dimension[1]) continue;  // :39
            float sox = (float)dx + blk.sample_offset[0] - 0.5f;
            float soy = (float)dy + blk.sample_offset[1] - 0.5f;
            float weight = orc_expf(gaussFac * (sox * sox + soy * soy)) - curveOffset;  // :44
            if (weight < 0.f) continue;
            float cw[4], nd[4], al[4];
            int32_t sx = (int32_t)(localx + (uint32_t)dx), sy = (int32_t)(localy + (uint32_t)dy);
            sample(0, sx, sy, cw);
            sample(1, sx, sy, nd);
            sample(2, sx, sy, al);
            vec3 nO = V(nd[0] - nc[0], nd[1] - nc[1], nd[2] - nc[2]);
            vec3 aO = V(al[0] - ac[0], al[1] - ac[1], al[2] - ac[2]);
            weight *= orc_exp_bilateral(dot(nO, nO) * 2.0f + dot(aO, aO));  // :54, exp(-(..)) in the FMA specification
            float wv[4] = {weight * cw[0], weight * cw[1], weight * cw[2], weight * cw[3]};
            if (std::isnan(wv[0]) || std::isnan(wv[1]) || std::isnan(wv[2]) || std::isnan(wv[3])) continue;
            for (int k = 0; k < 4; k++) outv[k] += wv[k];
          }
        }
        if (in_image) memcpy(px4(acc, W, globalx, globaly), outv, 16);  // imageStore :65
      }
    }
  });
}

int default_threads(int n) {
  if (n > 0) return n;
  int h = (int)std::thread::hardware_concurrency();
  return h > 0 ? h : 1;
}

}  // namespace

extern "C" {

int orc_hardware_threads(void) { return default_threads(0); }

void orc_math_eval(int fn, const float* a, const float* b, float* out, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) {
    switch (fn) {
      case 0: out[i] = orc_sinf(a[i]); break;
      case 1: out[i] = orc_cosf(a[i]); break;
      case 2: out[i] = orc_tanf(a[i]); break;
      case 3: out[i] = orc_expf(a[i]); break;
      case 4: out[i] = orc_atan2f(a[i], b[i]); break;
      case 6: out[i] = orc_exp_bilateral(a[i]); break;
      default: out[i] = orc_asinf(a[i]); break;
    }
  }
}

uint32_t orc_seed_rng(uint32_t seed) {
  Rng r;
  r.seedRng(seed);
  return r.rngState;
}
uint32_t orc_rand_uint(uint32_t* state) {
  Rng r{*state};
  uint32_t v = r.randUint();
  *state = r.rngState;
  return v;
}
float orc_rand_uniform_float(uint32_t* state) {
  Rng r{*state};
  float v = r.randUniformFloat();
  *state = r.rngState;
  return v;
}

void orc_camera_ray(const void* scene_info64, float px, float py, float dim_x, float dim_y,
                    float eps, OrcRay* out) {
  const SceneInfo* info = (const SceneInfo*)scene_info64;
  Ray r = getCameraRayAt(info->camera, px, py, dim_x, dim_y, eps);
  out->origin[0] = r.origin.x, out->origin[1] = r.origin.y, out->origin[2] = r.origin.z;
  out->direction[0] = r.direction.x, out->direction[1] = r.direction.y, out->direction[2] = r.direction.z;
  out->t_min = r.tMin;
  out->t_max = r.tMax;
}

void orc_recon_spatial_weights(uint32_t radius, float stddev, float so_x, float so_y, float* out) {
  const int R = (int)radius;
  const float gaussFac = -1.0f / (2.0f * stddev * stddev);
  const float curveOffset = orc_expf(gaussFac * (float)R * (float)R);
  int k = 0;
  for (int dx = -R; dx <= R; dx++)
    for (int dy = -R; dy <= R; dy++) {
      float sox = (float)dx + so_x - 0.5f, soy = (float)dy + so_y - 0.5f;
      float w = orc_expf(gaussFac * (sox * sox + soy * soy)) - curveOffset;
      out[k++] = w < 0.f ? -1.f : w;
    }
}

int orc_trace(const OrcScene* scene, const OrcRay* rays, uint64_t n, int use_bvh, float eps,
              int32_t* shape_id, float* t, float* uv, uint8_t* tie, int n_threads) {
  if (!scene || !rays || !shape_id) return -1;
  SceneView s = make_view(scene);
  if (use_bvh == 1 && (!s.bvh || s.bvhLength == 0)) return -2;
  std::shared_ptr<CullAccel> accel;
  if (use_bvh == 3) accel = build_cull_accel(s), s.accel = accel.get();
  n_threads = default_threads(n_threads);
  parallel_for(n, n_threads, [&](uint64_t b, uint64_t e, int) {
    std::vector<uint32_t> subset;
    for (uint64_t i = b; i < e; i++) {
      Ray ray{V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
              V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].t_min,
              rays[i].t_max};
      Intersection its{};
      // closest hit exactly as intersectScene does, minus the populate step
      its.objectID = -1;
      bool hit = intersectScene(s, use_bvh, eps, ray, its);
      shape_id[i] = hit ? its.objectID : -1;
      if (t) t[i] = hit ? its.t : 0.f;
      if (tie) {
        tie[i] = 0;
        if (hit) {
          // candidates of every primitive in the ORIGINAL interval; tie if the winner is
          // test-order dependent
          int total = s.numSpheres + s.numQuads + s.numTriangles;
          if (use_bvh == 3) {  // only primitives of the culled subset can be hit at all
            cull_candidates(*s.accel, ray, subset);
            total = (int)subset.size();
          }
          float tP = INFINITY;
          int P = -1;
          std::vector<std::pair<int, float>> cands;
          for (int kk = 0; kk < total; kk++) {
            const int k = use_bvh == 3 ? (int)subset[(size_t)kk] : kk;
            Intersection tmp{};
            if (intersectShape(s, ray, (uint32_t)k, tmp)) {
              cands.emplace_back(k, tmp.t);
              if (tmp.t < tP) {
                tP = tmp.t;
                P = k;
              }
            }
          }
          for (auto& c : cands) {
            if (c.first == P) continue;
            float tQ = c.second;
            if ((tQ - eps) < tP || tQ <= (tP - eps)) tie[i] = 1;
          }
        }
      }
      if (uv) {
        // raw barycentrics / sphere: re-run the winning primitive's test for its.uv
        uv[2 * i] = 0.f;
        uv[2 * i + 1] = 0.f;
        if (hit && its.objectID >= s.numSpheres) {
          Intersection tmp{};
          Ray r2 = ray;
          if (intersectShape(s, r2, (uint32_t)its.objectID, tmp)) {
            uv[2 * i] = tmp.uvx;
            uv[2 * i + 1] = tmp.uvy;
          }
        }
      }
    }
  });
  return 0;
}

int orc_occluded(const OrcScene* scene, const OrcRay* rays, uint64_t n, int use_bvh, float eps,
                 uint8_t* occluded, int n_threads) {
  if (!scene || !rays || !occluded) return -1;
  SceneView s = make_view(scene);
  if (use_bvh == 1 && (!s.bvh || s.bvhLength == 0)) return -2;
  std::shared_ptr<CullAccel> accel;
  if (use_bvh == 3) accel = build_cull_accel(s), s.accel = accel.get();
  n_threads = default_threads(n_threads);
  parallel_for(n, n_threads, [&](uint64_t b, uint64_t e, int) {
    for (uint64_t i = b; i < e; i++) {
      Ray ray{V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
              V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].t_min,
              rays[i].t_max};
      Intersection its{};
      occluded[i] = intersectScene(s, use_bvh, eps, ray, its) ? 1 : 0;
    }
  });
  return 0;
}

int orc_integrate_frame(const OrcScene* scene, const OrcBlock* blocks, uint64_t n_blocks,
                        const OrcParams* p, float* layers, OrcStats* stats, int n_threads) {
  if (!scene || !blocks || !p || !layers || n_blocks == 0) return -1;
  SceneView s = make_view(scene);
  if (p->use_bvh == 1 && (!s.bvh || s.bvhLength == 0)) return -2;
  std::shared_ptr<CullAccel> accel;
  if (p->use_bvh == 3) accel = build_cull_accel(s), s.accel = accel.get();
  n_threads = default_threads(n_threads);
  const uint32_t W = blocks[0].original_dimension[0], H = blocks[0].original_dimension[1];
  const uint64_t plane = (uint64_t)W * H * 4;
  std::vector<uint64_t> ext(n_threads, 0), sh(n_threads, 0), paths(n_threads, 0);
  auto t0 = std::chrono::steady_clock::now();
  for (uint64_t bi = 0; bi < n_blocks; bi++) {
    const OrcBlock& blk = blocks[bi];
    parallel_for(blk.dimension[1], n_threads, [&](uint64_t y0, uint64_t y1, int tid) {
      Ctx c;
      c.s = s;
      c.M_EPS = p->eps;
      c.useBvh = (int)p->use_bvh;
      for (uint32_t ly = (uint32_t)y0; ly < (uint32_t)y1; ly++)
        for (uint32_t lx = 0; lx < blk.dimension[0]; lx++) {
          uint32_t gx = lx + blk.origin[0], gy = ly + blk.origin[1];
          if (gx >= W || gy >= H) continue;
          PathOut o;
          if (!renderPixel(c, blk, lx, ly, *p, o)) continue;
          paths[tid]++;
          float* l0 = px4(layers, W, gx, gy);
          float* l1 = px4(layers + plane, W, gx, gy);
          float* l2 = px4(layers + 2 * plane, W, gx, gy);
          l0[0] = o.total.x, l0[1] = o.total.y, l0[2] = o.total.z, l0[3] = 1.0f;
          l1[0] = o.normal.x, l1[1] = o.normal.y, l1[2] = o.normal.z, l1[3] = o.depth;
          l2[0] = o.albedo.x, l2[1] = o.albedo.y, l2[2] = o.albedo.z, l2[3] = 0.f;
        }
      ext[tid] += c.nExt;
      sh[tid] += c.nShadow;
    });
  }
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    stats->n_paths = stats->n_extension_rays = stats->n_shadow_rays = 0;
    for (int t = 0; t < n_threads; t++) {
      stats->n_paths += paths[t];
      stats->n_extension_rays += ext[t];
      stats->n_shadow_rays += sh[t];
    }
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

int orc_reconstruct_frame(const OrcBlock* blocks, uint64_t n_blocks, const OrcParams* p,
                          const float* radiance, const float* normal_depth, const float* albedo,
                          float* accumulator, int n_threads) {
  if (!blocks || !p || !radiance || !normal_depth || !accumulator) return -1;
  n_threads = default_threads(n_threads);
  const uint32_t bs = p->block_size ? p->block_size : 128;
  for (uint64_t bi = 0; bi < n_blocks; bi++) {
    const OrcBlock& blk = blocks[bi];
    const uint32_t W = blk.original_dimension[0], H = blk.original_dimension[1];
    // The per-block intermediate texture is bs x bs (src/main.rs:1197-1201).  Texels with
    // local coords inside it but outside this block's clipped `dimension` hold whatever an
    // earlier block left there; they are only ever read as centre features of apron pixels
    // whose output lies outside the image (dropped store), so zeros are result-equivalent.
    auto sample = [&](int layer, int32_t sx, int32_t sy, float* out) {
      out[0] = out[1] = out[2] = out[3] = 0.f;
      if (sx < 0 || sy < 0 || (uint32_t)sx >= bs || (uint32_t)sy >= bs) return;  // robust OOB load
      if ((uint32_t)sx >= blk.dimension[0] || (uint32_t)sy >= blk.dimension[1]) return;
      uint32_t gx = (uint32_t)sx + blk.origin[0], gy = (uint32_t)sy + blk.origin[1];
      if (gx >= W || gy >= H) return;
      const float* src = layer == 0 ? radiance : (layer == 1 ? normal_depth : albedo);
      if (!src) return;
      memcpy(out, src + 4 * ((uint64_t)gy * W + gx), 16);
    };
    reconstructBlock(blk, *p, sample, accumulator, n_threads);
  }
  return 0;
}

int orc_render(const OrcScene* scene, const OrcBlock* blocks, uint64_t n_blocks,
               const OrcParams* p, float* accumulator, OrcStats* stats, int n_threads) {
  if (!scene || !blocks || !p || !accumulator || n_blocks == 0) return -1;
  SceneView s = make_view(scene);
  if (p->use_bvh == 1 && (!s.bvh || s.bvhLength == 0)) return -2;
  std::shared_ptr<CullAccel> accel;
  if (p->use_bvh == 3) accel = build_cull_accel(s), s.accel = accel.get();
  n_threads = default_threads(n_threads);
  const uint32_t bs = p->block_size ? p->block_size : 128;
  // intermediate texture: bs x bs x 3 layers RGBA32F, persistent across blocks
  std::vector<float> inter((uint64_t)bs * bs * 4 * 3, 0.f);
  const uint64_t plane = (uint64_t)bs * bs * 4;
  std::vector<uint64_t> ext(n_threads, 0), sh(n_threads, 0), paths(n_threads, 0);
  auto t0 = std::chrono::steady_clock::now();
  for (uint64_t bi = 0; bi < n_blocks; bi++) {
    const OrcBlock& blk = blocks[bi];
    if (blk.dimension[0] > bs || blk.dimension[1] > bs) return -3;
    const uint32_t W = blk.original_dimension[0], H = blk.original_dimension[1];
    // integrator dispatch (src/main.rs:891-897).  Threads of the padded 16x16 groups whose
    // pixel lies outside the image are skipped: they never reach the accumulator (Q9).
    parallel_for(blk.dimension[1], n_threads, [&](uint64_t y0, uint64_t y1, int tid) {
      Ctx c;
      c.s = s;
      c.M_EPS = p->eps;
      c.useBvh = (int)p->use_bvh;
      for (uint32_t ly = (uint32_t)y0; ly < (uint32_t)y1; ly++)
        for (uint32_t lx = 0; lx < blk.dimension[0]; lx++) {
          if (lx + blk.origin[0] >= W || ly + blk.origin[1] >= H) continue;
          PathOut o;
          if (!renderPixel(c, blk, lx, ly, *p, o)) continue;
          paths[tid]++;
          float* l0 = px4(inter.data(), bs, lx, ly);
          float* l1 = px4(inter.data() + plane, bs, lx, ly);
          float* l2 = px4(inter.data() + 2 * plane, bs, lx, ly);
          l0[0] = o.total.x, l0[1] = o.total.y, l0[2] = o.total.z, l0[3] = 1.0f;
          l1[0] = o.normal.x, l1[1] = o.normal.y, l1[2] = o.normal.z, l1[3] = o.depth;
          l2[0] = o.albedo.x, l2[1] = o.albedo.y, l2[2] = o.albedo.z, l2[3] = 0.f;
        }
      ext[tid] += c.nExt;
      sh[tid] += c.nShadow;
    });
    if (p->skip_recon) continue;
    auto sample = [&](int layer, int32_t sx, int32_t sy, float* out) {
      if (sx < 0 || sy < 0 || (uint32_t)sx >= bs || (uint32_t)sy >= bs) {
        out[0] = out[1] = out[2] = out[3] = 0.f;  // robust out-of-bounds image load (Q7)
        return;
      }
      memcpy(out, px4(inter.data() + layer * plane, bs, (uint32_t)sx, (uint32_t)sy), 16);
    };
    reconstructBlock(blk, *p, sample, accumulator, n_threads);
  }
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    stats->n_paths = stats->n_extension_rays = stats->n_shadow_rays = 0;
    for (int t = 0; t < n_threads; t++) {
      stats->n_paths += paths[t];
      stats->n_extension_rays += ext[t];
      stats->n_shadow_rays += sh[t];
    }
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

int orc_trace_path(const OrcScene* scene, const OrcBlock* block, uint32_t lx, uint32_t ly,
                   const OrcParams* p, OrcPathVertex* out, int capacity) {
  if (!scene || !block || !p || !out) return -1;
  Ctx c;
  c.s = make_view(scene);
  c.M_EPS = p->eps;
  c.useBvh = (int)p->use_bvh;
  if (c.useBvh == 1 && (!c.s.bvh || c.s.bvhLength == 0)) return -2;
  std::shared_ptr<CullAccel> accel;
  if (c.useBvh == 3) accel = build_cull_accel(c.s), c.s.accel = accel.get();
  PathOut o;
  int n = 0;
  renderPixel(c, *block, lx, ly, *p, o, out, capacity, &n);
  return n;
}

}  // extern "C"
