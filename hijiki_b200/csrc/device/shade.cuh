// Per-path arithmetic of the integrator: RNG, camera, hit-record population, emitter
// sampling, BSDF evaluation/sampling and Russian roulette — the body of
// reference shader/render.glsl:81-147 and what it includes, restated with exact fp32
// operation order (hjk_math.cuh).  One call of shade_vertex() is one iteration of the
// reference's bounce loop between its two intersectScene() calls; the wavefront kernels
// (kernels.cu) only move its inputs and outputs through queues.
#pragma once
#include "scene_dev.cuh"

namespace hjk {

// ------------------------------------------------------------------ rand.glsl
struct Rng {
  uint32_t state;
  HJK_HD uint32_t next_uint() {  // rand.glsl:2-7
    state ^= state << 13;
    state ^= state >> 17;
    state ^= state << 5;
    return state;
  }
  HJK_HD float uniform() { return x::mul(x::u2f(next_uint()), 1.0f / 4294967296.0f); }  // :18-20
};
HJK_HD uint32_t seed_rng(uint32_t seed) {  // rand.glsl:9-16 (Wang hash)
  seed = (seed ^ 61u) ^ (seed >> 16);
  seed *= 9u;
  seed = seed ^ (seed >> 4);
  seed *= 0x27d4eb2du;
  seed = seed ^ (seed >> 15);
  return seed;
}
HJK_HD vec3 rand_cos_hemisphere(Rng& rng) {  // rand.glsl:22-30
  const float u = rng.uniform();
  const float v = rng.uniform();
  const float r = x::sqrt(u);
  const float theta = x::mul(2.0f * HJK_PI_F, v);
  float sn, cs;
  sincos_det(theta, &sn, &cs);
  return V3(x::mul(r, cs), x::mul(r, sn), x::sqrt(x::gmax(0.0f, x::sub(1.0f, u))));
}
HJK_HD vec3 rand_uniform_sphere(Rng& rng) {  // rand.glsl:32-40
  const float u = rng.uniform();
  const float v = rng.uniform();
  const float z = x::sub(x::mul(2.0f, u), 1.0f);
  const float theta = x::mul(2.0f * HJK_PI_F, v);
  const float r = x::sqrt(x::sub(1.0f, x::mul(z, z)));
  float sn, cs;
  sincos_det(theta, &sn, &cs);
  return V3(x::mul(r, cs), x::mul(r, sn), z);
}
HJK_HD vec3 rand_barycentric(Rng& rng) {  // rand.glsl:42-50, fold kept as written
  float u = rng.uniform();
  float v = rng.uniform();
  if (x::add(u, v) > 1.0f) {
    u = x::sub(1.0f, v);
    v = x::sub(1.0f, u);
  }
  return V3(u, v, x::sub(x::sub(1.0f, u), v));
}

// ------------------------------------------------------------------ quaternion.glsl / camera
struct quat {
  float x, y, z, w;
};
HJK_HD quat quaternion_mult(quat a, quat b) {  // quaternion.glsl:1-6
  const vec3 av = V3(a.x, a.y, a.z), bv = V3(b.x, b.y, b.z);
  quat r;
  r.w = x::sub(x::mul(a.w, b.w), dot(av, bv));
  const vec3 v = (cross(av, bv) + av * b.w) + bv * a.w;
  r.x = v.x, r.y = v.y, r.z = v.z;
  return r;
}
HJK_HD vec3 quaternion_rotate(vec3 v, quat r) {  // quaternion.glsl:15-19
  quat p;
  p.x = v.x, p.y = v.y, p.z = v.z, p.w = 0.0f;
  const quat tmp = quaternion_mult(r, p);
  r.x = -r.x, r.y = -r.y, r.z = -r.z;
  const quat o = quaternion_mult(tmp, r);
  return V3(o.x, o.y, o.z);
}
// render.glsl:26-36.  (px, py) = vec2(global) + sampleOffset, dim = originalDimension.
HJK_HD void camera_ray(const HjkCamera& c, float px, float py, float dimx, float dimy, float eps,
                       f4& o_tmin, f4& d_tmax) {
  px = x::sub(px, x::mul(0.5f, dimx));
  py = x::sub(py, x::mul(0.5f, dimy));
  const float radians = x::mul(x::mul(0.5f, c.fov), HJK_PI_F / 180.0f);
  const float tn = tan_det(radians);
  const float half_w = x::mul(0.5f, dimx);
  px = x::div(x::mul(px, tn), half_w);
  py = x::div(x::mul(py, tn), half_w);
  quat rot;
  rot.x = c.rotation[0], rot.y = c.rotation[1], rot.z = c.rotation[2], rot.w = c.rotation[3];
  const vec3 dir = normalize(quaternion_rotate(V3(px, -py, -1.0f), rot));
  o_tmin = F4(c.position[0], c.position[1], c.position[2], eps);
  d_tmax = F4(dir.x, dir.y, dir.z, x::as_float(0x7F800000u));  // 1e100 -> +inf in fp32
}

// ------------------------------------------------------------------ hit record
struct Intersection {  // render.glsl:39-46
  int id;
  float t;
  vec3 p, n;
  float uvx, uvy;
  vec3 ft, fb;  // frame columns 0 and 1; column 2 is n
};
HJK_HD vec3 frame_mul(const Intersection& its, vec3 v) { return (its.ft * v.x + its.fb * v.y) + its.n * v.z; }

struct TriVerts {
  f4 a0, a1, b0, b1, c0, c1;
};
HJK_HD TriVerts load_tri(const SceneDev& sc, uint32_t ix) {
  const uint32_t ia = ld4(sc.triangles + 3 * (size_t)ix), ib = ld4(sc.triangles + 3 * (size_t)ix + 1),
                 ic = ld4(sc.triangles + 3 * (size_t)ix + 2);
  TriVerts t;
  t.a0 = ld16(sc.vertices + 2 * (size_t)ia), t.a1 = ld16(sc.vertices + 2 * (size_t)ia + 1);
  t.b0 = ld16(sc.vertices + 2 * (size_t)ib), t.b1 = ld16(sc.vertices + 2 * (size_t)ib + 1);
  t.c0 = ld16(sc.vertices + 2 * (size_t)ic), t.c1 = ld16(sc.vertices + 2 * (size_t)ic + 1);
  return t;
}

// scene.glsl:164-172 + populate{Sphere,Quad,Triangle}Intersection
HJK_HD void populate(const SceneDev& sc, int id, float t, float u, float v, vec3 o, vec3 d,
                     Intersection& its) {
  its.id = id;
  its.t = t;
  its.p = o + t * d;
  const uint32_t S = sc.num_spheres, Q = sc.num_quads;
  if ((uint32_t)id < S) {  // shapes/sphere.glsl:43-52
    const f4 sp = ld16(sc.spheres + id);
    const vec3 n = (its.p - xyz(sp)) / sp.w;
    its.n = n;
    its.ft = normalize(V3(-n.z, 0.f, n.x));
    its.fb = cross(n, its.ft);
    its.uvx = x::add(0.5f, x::div(atan2_det(n.z, n.x), 2.0f * HJK_PI_F));
    its.uvy = x::add(0.5f, x::div(asin_det(x::gmin(x::gmax(n.y, -1.0f), 1.0f)), HJK_PI_F));
    if (x::is_nan(its.uvx)) its.uvx = 0.f;
  } else if ((uint32_t)id < S + Q) {  // shapes/quad.glsl:27-32
    const f4* qp = sc.quads + 3 * (size_t)((uint32_t)id - S);
    its.ft = normalize(xyz(ld16(qp + 1)));
    its.fb = normalize(xyz(ld16(qp + 2)));
    its.n = cross(its.ft, its.fb);
    its.uvx = u, its.uvy = v;
  } else {  // shapes/triangle.glsl:54-78
    const TriVerts tv = load_tri(sc, (uint32_t)id - S - Q);
    const float l0 = x::sub(x::sub(1.0f, u), v), l1 = u, l2 = v;
    its.n = normalize((xyz(tv.a1) * l0 + xyz(tv.b1) * l1) + xyz(tv.c1) * l2);
    its.uvx = x::add(x::add(x::mul(tv.a0.w, l0), x::mul(tv.b0.w, l1)), x::mul(tv.c0.w, l2));
    its.uvy = x::add(x::add(x::mul(tv.a1.w, l0), x::mul(tv.b1.w, l1)), x::mul(tv.c1.w, l2));
    vec3 bt;
    const float ax = its.n.x < 0.f ? -its.n.x : its.n.x, ay = its.n.y < 0.f ? -its.n.y : its.n.y;
    if (ax > ay) {
      bt = V3(0.f, 1.f, 0.f);
    } else {
      bt = V3(1.f, 0.f, 0.f);
    }
    its.ft = normalize(cross(its.n, bt));
    its.fb = cross(its.n, its.ft);
  }
}

// ------------------------------------------------------------------ emitters (scene.glsl:44-89)
struct ShapeSample {  // ShapeQueryRecord, render.glsl:48-52
  vec3 p, n;
  float pdf;
};
HJK_HD void sample_shape(const SceneDev& sc, uint32_t shape, Rng& rng, ShapeSample& r) {
  const uint32_t S = sc.num_spheres, Q = sc.num_quads;
  if (shape < S) {  // shapes/sphere.glsl:54-58
    const f4 sp = ld16(sc.spheres + shape);
    r.n = rand_uniform_sphere(rng);
    r.p = xyz(sp) + sp.w * r.n;
    r.pdf = x::div(1.0f, x::mul(x::mul(x::mul(sp.w, sp.w), 4.0f), HJK_PI_F));
  } else if (shape < S + Q) {  // shapes/quad.glsl:34-45
    const f4* qp = sc.quads + 3 * (size_t)(shape - S);
    const vec3 org = xyz(ld16(qp)), e1 = xyz(ld16(qp + 1)), e2 = xyz(ld16(qp + 2));
    vec3 n = cross(e1, e2);
    const float area = length(n);
    r.n = n / area;
    const float u = rng.uniform();
    const float v = rng.uniform();
    r.p = (org + u * e1) + v * e2;
    r.pdf = x::div(1.0f, area);
  } else {  // shapes/triangle.glsl:81-102
    const TriVerts tv = load_tri(sc, shape - S - Q);
    const vec3 ab = xyz(tv.b0) - xyz(tv.a0), ac = xyz(tv.c0) - xyz(tv.a0);
    const vec3 n = cross(ab, ac);
    const float area = x::div(length(n), 2.0f);
    const vec3 l = rand_barycentric(rng);
    r.n = normalize((xyz(tv.a1) * l.x + xyz(tv.b1) * l.y) + xyz(tv.c1) * l.z);
    r.p = (xyz(tv.a0) * l.x + xyz(tv.b0) * l.y) + xyz(tv.c0) * l.z;
    r.pdf = x::div(1.0f, area);
  }
}

// returns importance (power / pdf); fills the shadow ray
HJK_HD vec3 sample_emitter(const SceneDev& sc, vec3 ref, Rng& rng, float eps, f4& sh_o, f4& sh_d) {
  float emitter_sample = rng.uniform();
  uint32_t emitter = 0;
  for (uint32_t i = 0; i < sc.num_emitters; i++) {
    emitter_sample = x::sub(emitter_sample, ld16(sc.emitters + i).y);
    if (emitter_sample < 0.f) {
      emitter = i;
      break;
    }
  }
  const f4 em = ld16(sc.emitters + emitter);
  const uint32_t shape = x::as_uint(em.x);
  ShapeSample sr;
  sample_shape(sc, shape, rng, sr);
  const uint32_t mat = ld4(sc.materials + shape);
  const vec3 power = xyz(ld16(sc.emissive + (mat & ((1u << HJK_MATERIAL_TAG_SHIFT) - 1u))));
  vec3 dir = sr.p - ref;
  const float dist = length(dir);
  dir = dir / dist;
  sh_o = F4(ref.x, ref.y, ref.z, x::mul(2.0f, eps));
  sh_d = F4(dir.x, dir.y, dir.z, x::sub(dist, eps));
  const float cos_theta = -dot(dir, sr.n);
  if (cos_theta < 0.f) return V3(0.f);
  const float pdf = x::div(x::mul(x::mul(x::mul(em.y, sr.pdf), dist), dist), cos_theta);
  return power / pdf;
}

// ------------------------------------------------------------------ materials (material.glsl)
HJK_HD float fract_f(float v) { return x::sub(v, x::floor(v)); }
HJK_HD vec3 checkerboard(const SceneDev& sc, uint32_t idx, float uvx, float uvy) {  // diffusecb.glsl:6-13
  const f4 a = ld16(sc.diffusecb + 2 * (size_t)idx), b = ld16(sc.diffusecb + 2 * (size_t)idx + 1);
  const float ux = fract_f(x::div(x::mul(0.5f, uvx), a.w));
  const float uy = fract_f(x::div(x::mul(0.5f, uvy), b.w));
  if ((ux < 0.5f) != (uy < 0.5f)) return xyz(b);
  return xyz(a);
}
HJK_HD vec3 material_color(const SceneDev& sc, uint32_t tag, uint32_t idx, const Intersection& its) {
  if (tag == HJK_MAT_DIFFUSE) return xyz(ld16(sc.diffuse + idx));
  return checkerboard(sc, idx, its.uvx, its.uvy);
}

// ------------------------------------------------------------------ one bounce
struct VertexIn {
  f4 ray_o, ray_d;  // the ray that produced the hit
  int hit_id;
  float hit_t, hit_u, hit_v;
  vec3 throughput, extinction;
  uint32_t rng;
  bool was_discrete;
  uint32_t bounce;
};
struct VertexOut {
  vec3 normal;  // first-bounce features (render.glsl:102-105)
  float depth;
  bool add_emission;
  vec3 emission;  // throughput * power, to be added to `total`
  bool has_shadow;
  f4 sh_o, sh_d;
  vec3 contribution;  // throughput * evalBSDF * importance, added when the shadow ray is free
  bool continues;     // false: the path ends after this vertex
  f4 next_o, next_d;
  vec3 throughput, extinction;
  uint32_t rng;
  bool was_discrete;
};

HJK_HD void shade_vertex(const SceneDev& sc, const VertexIn& in, uint32_t max_bounces,
                         uint32_t rr_start, float eps, VertexOut& out) {
  const vec3 ro = xyz(in.ray_o), rd = xyz(in.ray_d);
  Intersection its;
  populate(sc, in.hit_id, in.hit_t, in.hit_u, in.hit_v, ro, rd, its);
  out.normal = its.n;
  out.depth = its.t;
  Rng rng;
  rng.state = in.rng;

  const uint32_t mat = ld4(sc.materials + in.hit_id);
  const uint32_t tag = mat >> HJK_MATERIAL_TAG_SHIFT;
  const uint32_t idx = mat & ((1u << HJK_MATERIAL_TAG_SHIFT) - 1u);

  // render.glsl:111-112
  const float dist = length(ro - its.p);
  const vec3 ea = -in.extinction * dist;
  vec3 throughput = in.throughput * V3(exp_det(ea.x), exp_det(ea.y), exp_det(ea.z));
  vec3 extinction = in.extinction;

  out.add_emission = false;
  out.has_shadow = false;
  if (tag == HJK_MAT_EMISSIVE && in.was_discrete) {  // :114-116
    out.add_emission = true;
    out.emission = throughput * xyz(ld16(sc.emissive + idx));
  }
  const bool diffuse_like = tag == HJK_MAT_DIFFUSE || tag == HJK_MAT_DIFFUSECBOARD;
  vec3 color = V3(0.f);
  if (diffuse_like) {  // :117-126
    color = material_color(sc, tag, idx, its);
    f4 sh_o, sh_d;
    const vec3 importance = sample_emitter(sc, its.p, rng, eps, sh_o, sh_d);
    const vec3 wi = xyz(sh_d);
    if (length(importance) > eps && dot(wi, its.n) > 0.f) {
      out.has_shadow = true;
      out.sh_o = sh_o;
      out.sh_d = sh_d;
      const vec3 bsdf = (dot(its.n, wi) * color) / HJK_PI_F;  // material.glsl:18-30
      out.contribution = (throughput * bsdf) * importance;
    }
  }

  // sampleBSDF, material.glsl:33-91
  vec3 wo = V3(0.f);
  bool wo_written = true;
  if (diffuse_like) {
    const vec3 wl = rand_cos_hemisphere(rng);
    wo = frame_mul(its, wl);
    throughput = throughput * color;
  } else if (tag == HJK_MAT_MIRROR) {
    wo = reflect(rd, its.n);
    throughput = throughput * V3(1.f);
  } else if (tag == HJK_MAT_DIELECTRIC) {
    const f4 de = ld16(sc.dielectric + idx);
    float eta = de.w;
    float eta_inv = x::div(1.0f, eta);
    float cos_i = -dot(its.n, rd);
    vec3 normal = its.n;
    bool inside = cos_i > 0.f;
    if (cos_i < 0.f) {
      eta = eta_inv;
      eta_inv = x::div(1.0f, eta);
      normal = -normal;
      cos_i = -cos_i;
    }
    const float k = x::sub(1.0f, x::mul(x::mul(eta_inv, eta_inv), x::sub(1.0f, x::mul(cos_i, cos_i))));
    if (k <= 0.f) {
      wo = reflect(rd, normal);
    } else {
      const float cos_o = x::sqrt(k);
      const float rho_par = x::div(x::sub(x::mul(eta, cos_i), cos_o), x::add(x::mul(eta, cos_i), cos_o));
      const float rho_orth = x::div(x::sub(cos_i, x::mul(eta, cos_o)), x::add(cos_i, x::mul(eta, cos_o)));
      const float f_r = x::mul(0.5f, x::add(x::mul(rho_par, rho_par), x::mul(rho_orth, rho_orth)));
      if (rng.uniform() < f_r) {
        wo = reflect(rd, normal);
      } else {
        inside = !inside;
        const vec3 parallel = rd - dot(rd, normal) * normal;
        wo = eta_inv * parallel - x::sqrt(k) * normal;
      }
    }
    if (inside) extinction = xyz(de);
    throughput = throughput * V3(1.f);
  } else {  // emissive: weight 0 and `wo` unwritten (material.glsl:88-89)
    throughput = throughput * V3(0.f);
    wo_written = false;
  }

  bool terminate = false;
  if (in.bounce > rr_start) {  // render.glsl:137-144
    const float q = x::gmin(0.99f, x::gmax(throughput.x, x::gmax(throughput.y, throughput.z)));
    if (rng.uniform() > q) {
      terminate = true;
    } else {
      throughput = throughput / q;
    }
  }
  out.continues = !terminate && wo_written && (in.bounce + 1u < max_bounces);
  out.next_o = F4(its.p.x, its.p.y, its.p.z, x::mul(2.0f, eps));
  out.next_d = F4(wo.x, wo.y, wo.z, x::as_float(0x7F800000u));
  out.throughput = throughput;
  out.extinction = extinction;
  out.rng = rng.state;
  out.was_discrete = !diffuse_like;
}

}  // namespace hjk
