// Software traversal of the 8-wide compressed BVH (cwbvh.h) — replaces the node loop of
// reference shader/scene.glsl:99-133 — and the primitive tests of shader/shapes/*.glsl,
// restated with exact fp32 operation order so that accepted hits carry the same t/u/v bits
// as the reference arithmetic.
//
// Closest hit, default mode: the hit of smallest t the reference's interval test accepts, exactly
// equal t resolved by the lower shape id (closer_hit below).  That is a function of the SET of
// primitives the ray hits, not of the order they are visited in, so results do not depend on warp
// composition, wave size or the tree.  The reference instead keeps "tMax = t - M_EPS after an
// accepted hit" (scene.glsl:116) along its linear scan, which makes its winner order-dependent
// among hits closer than M_EPS to each other (SURVEY §8-Q1 "ties"): outside such clusters both
// rules give the same hit; inside them this one reports the nearest member, the exact-tie mode
// (TieCands below) the member the reference's scan order picks.  Any hit returns as soon as one
// primitive is accepted, which is exactly when the reference's shadow overload
// (scene.glsl:92-96) returns true.
//
// Box culling uses FMA and conservative (outward-rounded, padded) boxes; it never decides a
// hit, it only skips primitives whose exact test would fail.
#pragma once
#include "../cwbvh.h"
#include "scene_dev.cuh"

namespace hjk {

HJK_HD int hi_bit(uint32_t v) {  // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}
HJK_HD int pop_count(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
// per byte: 0xFF when the byte's top bit is set, else 0
HJK_HD uint32_t sign_extend_s8x4(uint32_t v) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(r) : "r"(v));
  return r;
#else
  return ((v >> 7) & 0x01010101u) * 0xFFu;
#endif
}
// byte j of `word` as a float: shift/mask + I2F.  (Measured on B200: building 2^23 + byte with PRMT
// and subtracting 2^23 is slower — it moves the conversion from the otherwise idle XU pipe onto the
// ALU/FMA pipes the slab test already saturates.)
HJK_HD float byte_to_float(uint32_t word, int j) { return (float)((word >> (8 * j)) & 0xFFu); }

struct TravState {
  float ox, oy, oz, dx, dy, dz, tmin, tmax;
  float idx, idy, idz;   // reciprocal direction (zeros replaced by +-1e-20 before inversion)
  uint32_t octinv4;      // (7 - octant) replicated in 4 bytes
  uint32_t ng_x, ng_y;   // node group: first child index, (hit bits << 24) | imask
  uint32_t tg_x, tg_y;   // primitive group: first record index, pending hit bits
  int32_t hit_id;        // global shape id of the current winner, -1 = none
  float hit_t, hit_u, hit_v;
  uint32_t slot;         // caller payload (path slot / ray index); top bit = any-hit ray
  // sphere guard (see trav_init): per-ray constants of the per-node box inflation, box-test interval
  float guard_e, guard_lo;
  float box_tmin, box_tmax_scale;
  float sph_clip;  // far clip of sphere-only subtrees (directions shorter than unit length), else +inf
  // exact-tie mode (see TieCands): boxes are culled against t_cull, tmax stays the ray's own
  float t_cull;
};

HJK_HD float safe_rcp(float d) {
  const float a = d < 0.f ? -d : d;
  if (!(a > 1e-20f)) d = (x::as_uint(d) >> 31) ? -1e-20f : 1e-20f;
  return 1.0f / d;
}

// SPHERE GUARD.  The reference's sphere test (shapes/sphere.glsl:18-41) solves t^2 + b t + c = 0,
// i.e. it assumes |direction| = 1.  Directions drift off unit length along specular chains
// (n = (p - centre)/r is not exactly unit, reflect() amplifies it), and for s^2 = |d|^2 != 1 the
// test is no longer geometric: it accepts exactly the rays whose line passes within
//     R'^2 = r^2/s^2 + L^2 (1 - 1/s^2)          (L = |origin - centre|)
// of the centre, and reports t' = s^2 * (true entry parameter of that ball).  To stay a superset
// of what the reference accepts, the child boxes of a node are inflated by Delta >= R' - r, with L
// bounded by the distance from the ray origin to the farthest corner of the node (every sphere below
// the node has its centre inside it), and the box interval is [0, tMax * max(1, 1/s^2)]:
//     s^2 >= 1:  Delta = sqrt(r_min^2 + L^2 (1 - 1/s^2)) - r_min      s^2 < 1:  Delta = r_max (1/s - 1)
// For |s^2 - 1| of a few ulps this degenerates to a pad of ~1e-7 L^2 / r; for scenes without spheres
// it is compiled out (triangle and quad tests are homogeneous in d, hence geometric for any s).
// FAR CLIP.  For s^2 < 1 the discriminant of the reference's quadratic, 4 (d.l)^2 - 4 (L^2 - r^2), is negative
// whenever L^2 (1 - s^2) > r^2 (because (d.l)^2 <= s^2 L^2): a sphere farther than r / sqrt(1 - s^2) from the ray
// ORIGIN is never accepted, however the ray points.  For the others both roots obey |t| <= |b| + r <= 2 s L + r, so
// every accepted sphere hit has t <= (2 s / sqrt(1 - s^2) + 1) r_max.  Subtrees that hold nothing but spheres
// (kWideOnlySpheres) are therefore walked with that far distance — along specular chains the directions the
// reference arithmetic produces shrink far below unit length within a few bounces (median |s^2 - 1| = 0.3 at
// bounce 12 of the 512-sphere lattice), where the inflation alone would make every box of the scene pass.
// GUARD (a per-scene constant the kernels are instantiated for): 0 = the scene has no spheres, 1 = some
// subtrees hold spheres (each node says so: kWideHasSpheres), 2 = every node does (an all-sphere scene, or a
// tree from the GPU builder): the per-node flag test is compiled out.
template <int GUARD>
HJK_HD void trav_init(TravState& s, const SceneDev& sc, const f4& o_tmin, const f4& d_tmax) {
  s.ox = o_tmin.x, s.oy = o_tmin.y, s.oz = o_tmin.z, s.tmin = o_tmin.w;
  s.dx = d_tmax.x, s.dy = d_tmax.y, s.dz = d_tmax.z, s.tmax = d_tmax.w;
  s.idx = safe_rcp(s.dx), s.idy = safe_rcp(s.dy), s.idz = safe_rcp(s.dz);
  s.guard_e = s.guard_lo = 0.f;
  s.box_tmin = s.tmin;
  s.box_tmax_scale = 1.0f;
  s.sph_clip = x::as_float(0x7F800000u);
  if (GUARD) {
    const float s2 = s.dx * s.dx + s.dy * s.dy + s.dz * s.dz;
    const float inv_s2 = 1.0f / s2;
    if (s2 >= 1.0f) {
      s.guard_e = 1.0f - inv_s2;
    } else {
      s.guard_lo = sc.sph_rmax * (sqrtf(inv_s2) - 1.0f);
      // 1.01: the float evaluation of the discriminant may pass a sphere a few ulps beyond the exact bound
      const float clip = (2.0f * sqrtf(s2 / (1.0f - s2)) + 1.0f) * sc.sph_rmax * 1.01f;
      if (clip >= 0.f) s.sph_clip = clip;  // (NaN: no clip)
    }
    if (!(s.guard_e >= 0.f) || !(s.guard_lo >= 0.f))  // NaN / degenerate direction: no culling
      s.guard_lo = x::as_float(0x7F800000u);
    s.box_tmin = 0.f;
    s.box_tmax_scale = (inv_s2 > 1.0f ? inv_s2 : 1.0f) * 1.000001f;
    if (!(s.box_tmax_scale >= 1.0f)) s.box_tmax_scale = x::as_float(0x7F800000u);
  }
  const uint32_t oct = (s.idx < 0.f ? 1u : 0u) | (s.idy < 0.f ? 2u : 0u) | (s.idz < 0.f ? 4u : 0u);
  s.octinv4 = (7u - oct) * 0x01010101u;
  s.ng_x = 0;
  s.ng_y = 0x80000000u;  // "slot 7^octinv of a virtual parent with imask 0" = node 0
  // A direction of exactly (0,0,0) is outside the contract: the reference arithmetic turns it into
  // "hits" at t = +inf on every triangle facing the origin (1/dot(d,n) = inf, shapes/triangle.glsl:24)
  // — an artefact no integrator ray can produce (directions come out of normalize/reflect/frame
  // products).  Such a ray is reported as a miss: no traversal work is queued.
  if (s.dx == 0.f && s.dy == 0.f && s.dz == 0.f) s.ng_y = 0;
  s.tg_x = s.tg_y = 0;
  s.hit_id = -1;
  s.hit_t = 0.f, s.hit_u = 0.f, s.hit_v = 0.f;
  s.t_cull = x::as_float(0x7F800000u);
}

// 8 child boxes of one node against the ray interval; returns the 32-bit hit mask
// (bits 31..24: inner children in near-to-far priority order, bits 23..0: primitives).
template <int GUARD, bool EXACT>
HJK_HD uint32_t intersect_node(const SceneDev& sc, const TravState& s, const f4& q0, const f4& q1, const f4& q2,
                               const f4& q3, const f4& q4) {
  const uint32_t e_imask = x::as_uint(q0.w);
  const float scx = x::as_float((e_imask & 0xFFu) << 23), scy = x::as_float(((e_imask >> 8) & 0xFFu) << 23),
              scz = x::as_float(((e_imask >> 16) & 0xFFu) << 23);
  const float adjx = scx * s.idx, adjy = scy * s.idy, adjz = scz * s.idz;
  float infl_x = 0.f, infl_y = 0.f, infl_z = 0.f;
  // the sphere guard applies to nodes with a sphere below them; the rest of the tree holds triangles and
  // quads only, whose tests are geometric for any direction length: plain slab test, the ray's own interval
  const bool guarded = GUARD == 2 || (GUARD == 1 && (x::as_uint(q1.y) & kWideHasSpheres) != 0u);
  if (guarded) {  // inflation from the farthest corner of this node's frame (see trav_init)
    const float fx = fmaxf(fabsf(q0.x - s.ox), fabsf(fmaf(255.0f, scx, q0.x) - s.ox));
    const float fy = fmaxf(fabsf(q0.y - s.oy), fabsf(fmaf(255.0f, scy, q0.y) - s.oy));
    const float fz = fmaxf(fabsf(q0.z - s.oz), fabsf(fmaf(255.0f, scz, q0.z) - s.oz));
    const float L2 = fx * fx + fy * fy + fz * fz;
    float delta = fmaxf(sqrtf(fmaf(L2, s.guard_e, sc.sph_rmin * sc.sph_rmin)) - sc.sph_rmin, s.guard_lo);
    delta = fmaf(delta, 1.02f, 1e-6f * (fx + fy + fz + 1.0f));
    if (!(delta >= 0.f)) delta = x::as_float(0x7F800000u);
    infl_x = delta * fabsf(s.idx), infl_y = delta * fabsf(s.idy), infl_z = delta * fabsf(s.idz);
  }
  const float orgx = (q0.x - s.ox) * s.idx;
  const float orgy = (q0.y - s.oy) * s.idy;
  const float orgz = (q0.z - s.oz) * s.idz;
  // near planes move towards the origin, far planes away from it, by the sphere-guard inflation
  const float o0x = orgx - infl_x, o0y = orgy - infl_y, o0z = orgz - infl_z;
  const float o1x = orgx + infl_x, o1y = orgy + infl_y, o1z = orgz + infl_z;
  const float box_tmin = guarded ? s.box_tmin : s.tmin;
  const float far_t = EXACT ? fminf(s.tmax, s.t_cull) : s.tmax;
  float box_tmax = guarded ? far_t * s.box_tmax_scale : far_t;
  if (GUARD && guarded && (x::as_uint(q1.y) & kWideOnlySpheres) != 0u) box_tmax = fminf(box_tmax, s.sph_clip);
  const bool nx = s.idx < 0.f, ny = s.idy < 0.f, nz = s.idz < 0.f;
  uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int half = 0; half < 2; half++) {
    const uint32_t meta4 = x::as_uint(half ? q1.w : q1.z);
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (s.octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t lox = x::as_uint(half ? q2.y : q2.x), loy = x::as_uint(half ? q2.w : q2.z);
    const uint32_t loz = x::as_uint(half ? q3.y : q3.x), hix = x::as_uint(half ? q3.w : q3.z);
    const uint32_t hiy = x::as_uint(half ? q4.y : q4.x), hiz = x::as_uint(half ? q4.w : q4.z);
    const uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
    const uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
    const uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; j++) {
      const float t0x = fmaf(byte_to_float(nearx, j), adjx, o0x);
      const float t0y = fmaf(byte_to_float(neary, j), adjy, o0y);
      const float t0z = fmaf(byte_to_float(nearz, j), adjz, o0z);
      const float t1x = fmaf(byte_to_float(farx, j), adjx, o1x);
      const float t1y = fmaf(byte_to_float(fary, j), adjy, o1y);
      const float t1z = fmaf(byte_to_float(farz, j), adjz, o1z);
      const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, box_tmin));
      const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, box_tmax));
      if (cmin <= cmax) {
        const uint32_t bits = (child_bits4 >> (8 * j)) & 0xFFu;
        const uint32_t index = (bit_index4 >> (8 * j)) & 0xFFu;
        hitmask |= bits << index;
      }
    }
  }
  return hitmask;
}

// Default-mode replacement rule of the closest hit (see the head of this file): strictly closer, or
// exactly as far with a lower shape id.  After an accepted hit tmax = t, so intersect_prim's own
// interval test already rejects everything farther.
HJK_HD bool closer_hit(const TravState& s, float t, uint32_t id) {
  return s.hit_id < 0 || t < s.hit_t || (t == s.hit_t && id < (uint32_t)s.hit_id);
}

// One primitive record against the ray, in the reference's exact arithmetic.
//   triangle: shapes/triangle.glsl:15-52      sphere: shapes/sphere.glsl:18-41
//   quad    : shapes/quad.glsl:7-25
// On acceptance writes t (and u, v for triangles/quads) and returns true.
HJK_HD bool intersect_prim(const SceneDev& sc, const TravState& s, const f4& r0, const f4& r1,
                           const f4& r2, const f4& r3, float& t_out, float& u_out, float& v_out) {
  const uint32_t id = x::as_uint(r0.w);
  const vec3 o = V3(s.ox, s.oy, s.oz), d = V3(s.dx, s.dy, s.dz);
  if (id < sc.num_spheres) {
    const float r = r1.x;
    const vec3 l = o - xyz(r0);
    const float b = x::mul(2.0f, dot(d, l));
    const float c = x::sub(dot(l, l), x::mul(r, r));
    float disc = x::sub(x::mul(b, b), x::mul(4.0f, c));
    if (disc < 0.0f) return false;
    disc = x::sqrt(disc);
    const float t0 = x::mul(-0.5f, x::add(b, disc));
    if (s.tmin <= t0 && t0 <= s.tmax) {
      t_out = t0, u_out = 0.f, v_out = 0.f;
      return true;
    }
    const float t1 = x::mul(-0.5f, x::sub(b, disc));
    if (s.tmin <= t1 && t1 <= s.tmax) {
      t_out = t1, u_out = 0.f, v_out = 0.f;
      return true;
    }
    return false;
  }
  const vec3 e1 = xyz(r1), e2 = xyz(r2);
#if HJK_PRIM_STRIDE == 4
  const vec3 n = xyz(r3);  // precomputed cross(e1, e2), same bits
#else
  const vec3 n = cross(e1, e2);
#endif
  const vec3 ro = o - xyz(r0);
  const vec3 q = cross(ro, d);
  const float dd = x::div(1.0f, dot(d, n));
  if (id < sc.num_spheres + sc.num_quads) {
    const float u = x::mul(dd, dot(-q, e2));
    const float v = x::mul(dd, dot(q, e1));
    if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) return false;
    const float t = x::mul(dd, dot(-n, ro));
    if (s.tmin <= t && t <= s.tmax) {
      t_out = t, u_out = u, v_out = v;
      return true;
    }
    return false;
  }
  const float u = x::mul(dd, dot(-q, e2));
  const float v = x::mul(dd, dot(q, e1));
  if (u < 0.f || v < 0.f || x::add(u, v) > 1.f) return false;
  const float t = x::mul(dd, dot(-n, ro));
  if (s.tmin <= t && t <= s.tmax) {
    t_out = t, u_out = u, v_out = v;
    return true;
  }
  return false;
}

// EXACT-TIE MODE.  The reference reports, among hits closer than M_EPS to each other, whichever its
// linear scan (scene.glsl:134-157: spheres, quads, triangles in index order, tMax = t - M_EPS after
// every accepted hit) accepts last — a function of the primitive ORDER, not of geometry.  The
// default traversal visits near children first and therefore may pick another member of such a
// cluster ("ties: excluded and counted").  In exact mode the traversal never shrinks tMax; it
// records every hit within a window of the nearest one (boxes are culled against nearest + window),
// then sorts the cluster by shape id and replays the reference's acceptance rule on it.  Hits
// beyond the first gap >= M_EPS cannot influence the winner (they are accepted before and
// superseded, or rejected after), so the replay is exact unless the cluster is longer than the
// window or the candidate list — then the ray is counted as unresolved.
constexpr int kTieCands = 12;
constexpr float kTieWindowEps = 8.0f;  // window = 8 * M_EPS
struct TieCands {
  float t[kTieCands];
  uint32_t id[kTieCands];
  uint32_t prim[kTieCands];
  uint32_t n;
  uint32_t unresolved;
  HJK_HD void reset() { n = 0, unresolved = 0; }
  // drops the candidates the nearest hit has since left behind (t beyond its window)
  HJK_HD void prune(float t_cull) {
    uint32_t m = 0;
    for (uint32_t k = 0; k < n; k++)
      if (t[k] <= t_cull) {
        t[m] = t[k], id[m] = id[k], prim[m] = prim[k];
        m++;
      }
    n = m;
  }
  HJK_HD void add(float tt, uint32_t i, uint32_t p) {
    if (n < (uint32_t)kTieCands) {
      t[n] = tt, id[n] = i, prim[n] = p;
      n++;
    } else {
      unresolved = 1;
    }
  }
};
struct NoCands {  // default mode: nothing is recorded
  HJK_HD void reset() {}
  HJK_HD void prune(float) {}
  HJK_HD void add(float, uint32_t, uint32_t) {}
};
HJK_HD uint32_t cands_unresolved(const TieCands& c) { return c.unresolved; }
HJK_HD bool cands_full(const TieCands& c) { return c.n == (uint32_t)kTieCands; }
HJK_HD bool cands_full(const NoCands&) { return false; }
HJK_HD uint32_t cands_unresolved(const NoCands&) { return 0u; }

// Forward declaration (defined below with the primitive tests).
HJK_HD bool intersect_prim(const SceneDev& sc, const TravState& s, const f4& r0, const f4& r1,
                           const f4& r2, const f4& r3, float& t_out, float& u_out, float& v_out);

// Picks the reference's winner among the recorded candidates; s.hit_* hold the nearest one on entry.
HJK_HD void resolve_ties(const SceneDev& sc, TravState& s, TieCands& c, float eps) {
  if (c.n <= 1u) return;
  // selection sort by t (n <= 12)
  for (uint32_t i = 0; i + 1 < c.n; i++) {
    uint32_t m = i;
    for (uint32_t k = i + 1; k < c.n; k++)
      if (c.t[k] < c.t[m]) m = k;
    if (m != i) {
      const float tt = c.t[i];
      const uint32_t ii = c.id[i], pp = c.prim[i];
      c.t[i] = c.t[m], c.id[i] = c.id[m], c.prim[i] = c.prim[m];
      c.t[m] = tt, c.id[m] = ii, c.prim[m] = pp;
    }
  }
  // the cluster ends at the first gap: fl(t[k] - eps) >= t[k-1] means hit k, scanned first, cannot make the scan
  // reject anything before it — unless the two are exactly as far (an eps below one ulp of t leaves fl(t - eps) = t):
  // then whichever comes later in the scan is accepted over the other, so they still belong together
  uint32_t g = 1;
  while (g < c.n && (!(x::sub(c.t[g], eps) >= c.t[g - 1]) || c.t[g] == c.t[g - 1])) g++;
  const float window_end = x::add(c.t[0], x::mul(kTieWindowEps, eps));
  if (g == c.n && x::add(c.t[g - 1], eps) > window_end) c.unresolved = 1;  // chain may continue past the window
  if (g == 1u) return;  // the nearest hit stands alone
  // replay the linear scan over the cluster in shape-id order
  float cur = s.tmax;
  uint32_t winner = 0xFFFFFFFFu, last_id = 0;
  for (uint32_t step = 0; step < g; step++) {
    uint32_t m = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < g; k++)
      if ((step == 0 || c.id[k] > last_id) && (m == 0xFFFFFFFFu || c.id[k] < c.id[m])) m = k;
    if (m == 0xFFFFFFFFu) break;
    last_id = c.id[m];
    if (c.t[m] <= cur) {
      winner = m;
      cur = x::sub(c.t[m], eps);
    }
  }
  if (winner == 0xFFFFFFFFu || (int32_t)c.id[winner] == s.hit_id) return;
  const f4* pp = sc.prims + (size_t)c.prim[winner] * HJK_PRIM_STRIDE;
  const f4 r0 = ld16(pp), r1 = ld16(pp + 1), r2 = ld16(pp + 2);
#if HJK_PRIM_STRIDE == 4
  const f4 r3 = ld16(pp + 3);
#else
  const f4 r3 = r2;
#endif
  float t, u, v;
  if (intersect_prim(sc, s, r0, r1, r2, r3, t, u, v)) {
    s.hit_id = (int32_t)c.id[winner];
    s.hit_t = t, s.hit_u = u, s.hit_v = v;
  }
}

HJK_HD void resolve_ties(const SceneDev&, TravState&, NoCands&, float) {}

// Scheduling policy of the host-side harness and of single-ray callers: never yield, never postpone.
struct TravNoPolicy {
  HJK_HD bool yield() const { return false; }
  HJK_HD bool postpone() const { return false; }
};

// Runs the traversal until the ray is finished (returns true) or `policy.yield()` asks to stop
// (returns false; call again with the same state to resume).  `policy.postpone()` is asked inside
// the primitive loop: when it says yes and the lane still has inner-node work, the pending
// primitives go to the stack and are tested later, when more lanes of the warp have primitive work
// (only the visiting order changes, never what is accepted).
//   Stack: push(uint32_t, uint32_t), pop(uint32_t&, uint32_t&), empty().  A postponed primitive group takes a
//   stack entry on top of the one node-group entry per level: callers whose stack holds fewer than two entries
//   per tree level must use a policy that never postpones (hjk_scene_upload does: see postpone_lanes).
// Any-hit rays (top bit of s.slot set) return at the first accepted primitive.
// EXACT: closest-hit rays record their candidates in `cands` (TieCands) and resolve ties at the end.
template <int GUARD, bool EXACT, class Stack, class Policy, class Cands>
HJK_HD bool trav_run(const SceneDev& sc, TravState& s, Stack& stack, float eps, const Policy& policy, Cands& cands) {
  for (;;) {
    if (s.ng_y > 0x00FFFFFFu) {
      const uint32_t hits_imask = s.ng_y;
      const int bit = hi_bit(hits_imask);
      s.ng_y &= ~(1u << bit);
      if (s.ng_y > 0x00FFFFFFu) stack.push(s.ng_x, s.ng_y);
      const uint32_t slot = ((uint32_t)bit - 24u) ^ (s.octinv4 & 0xFFu);
      const uint32_t rel = (uint32_t)pop_count(hits_imask & ~(0xFFFFFFFFu << slot));
      const f4* np = sc.nodes + (size_t)(s.ng_x + rel) * 5;
      const f4 q0 = ld16(np), q1 = ld16(np + 1), q2 = ld16(np + 2), q3 = ld16(np + 3), q4 = ld16(np + 4);
      const uint32_t hitmask = intersect_node<GUARD, EXACT>(sc, s, q0, q1, q2, q3, q4);
      s.ng_x = x::as_uint(q1.x);
      s.ng_y = (hitmask & 0xFF000000u) | (x::as_uint(q0.w) >> 24);
      s.tg_x = x::as_uint(q1.y) & kWidePrimBaseMask;
      s.tg_y = hitmask & 0x00FFFFFFu;
    } else {
      s.tg_x = s.ng_x, s.tg_y = s.ng_y;
      s.ng_x = s.ng_y = 0;
    }
    while (s.tg_y) {
      if (s.ng_y > 0x00FFFFFFu && policy.postpone()) {
        stack.push(s.tg_x, s.tg_y);
        break;
      }
      const int i = hi_bit(s.tg_y);
      s.tg_y &= ~(1u << i);
      const uint32_t prim_index = s.tg_x + (uint32_t)i;
      const f4* pp = sc.prims + (size_t)prim_index * HJK_PRIM_STRIDE;
      const f4 r0 = ld16(pp), r1 = ld16(pp + 1), r2 = ld16(pp + 2);
#if HJK_PRIM_STRIDE == 4
      const f4 r3 = ld16(pp + 3);
#else
      const f4 r3 = r2;
#endif
      float t, u, v;
      if (intersect_prim(sc, s, r0, r1, r2, r3, t, u, v)) {
        if (s.slot >> 31) {
          s.hit_id = (int32_t)x::as_uint(r0.w);
          return true;
        }
        if (EXACT) {
          if (t <= s.t_cull) {
            if (s.hit_id < 0 || t < s.hit_t) {
              s.hit_id = (int32_t)x::as_uint(r0.w);
              s.hit_t = t, s.hit_u = u, s.hit_v = v;
              s.t_cull = x::add(t, x::mul(kTieWindowEps, eps));
            }
            if (cands_full(cands)) cands.prune(s.t_cull);  // far-to-near visiting order leaves stale entries
            cands.add(t, x::as_uint(r0.w), prim_index);
          }
        } else if (closer_hit(s, t, x::as_uint(r0.w))) {
          s.hit_id = (int32_t)x::as_uint(r0.w);
          s.hit_t = t, s.hit_u = u, s.hit_v = v;
          s.tmax = t;
        }
      }
    }
    if (s.ng_y <= 0x00FFFFFFu) {
      if (stack.empty()) {
        if (EXACT && !(s.slot >> 31)) resolve_ties(sc, s, cands, eps);
        return true;
      }
      stack.pop(s.ng_x, s.ng_y);
    }
    if (policy.yield()) return false;
  }
}

}  // namespace hjk
