// TEST HARNESS — compiles the DEVICE headers of the product (hijiki_b200/csrc/device/*.cuh)
// for the host with g++ -ffp-contract=off, so that the traversal, shading and
// reconstruction logic the CUDA kernels run can be unit-tested against the oracle in the
// CPU-only container.  It is not part of libhijiki_b200 and is never shipped or timed: the
// product has no CPU path.  The loops below replay the wavefront schedule of
// csrc/device/kernels.cu sequentially (paths are independent, so order does not matter).
#include <cstring>
#include <string>
#include <vector>

#include "../../hijiki_b200/csrc/device/recon.cuh"
#include "../../hijiki_b200/csrc/device/shade.cuh"
#include "../../hijiki_b200/csrc/device/traverse.cuh"
#include "../../hijiki_b200/csrc/host/pass_plan.h"
#include "../../hijiki_b200/csrc/host/wide_bvh_host.h"

using namespace hjk;

namespace {

struct HostStack {
  uint32_t data[2 * 64];
  int n = 0;
  int max_n = 0;
  void push(uint32_t a, uint32_t b) {
    data[2 * n] = a, data[2 * n + 1] = b;
    n++;
    if (n > max_n) max_n = n;
  }
  void pop(uint32_t& a, uint32_t& b) {
    n--;
    a = data[2 * n], b = data[2 * n + 1];
  }
  bool empty() const { return n == 0; }
};

// Scheduling stress: postpones and yields on a fixed pseudo-random pattern, so the resumable state
// machine and the primitive postponing of trav_run are exercised on the CPU (bit 1 of ht_trace's mode).
struct ChaosPolicy {
  mutable uint32_t state;
  bool next() const {
    state = state * 1664525u + 1013904223u;
    return (state >> 28) < 6u;
  }
  bool yield() const { return next(); }
  bool postpone() const { return next(); }
};

// the kernels are instantiated per scene for GUARD = "has spheres"; mirror that choice here
template <bool ANY_HIT, class Policy>
bool run_ray(const SceneDev& sc, TravState& s, struct HostStack& st, float eps, const Policy& p);
thread_local TieCands g_cands;  // exact-tie mode: candidates of the ray in flight
template <class F4>
void init_ray(const SceneDev& sc, TravState& s, const F4& o, const F4& d) {
  g_cands.reset();
  if (sc.num_spheres) trav_init<true>(s, sc, o, d); else trav_init<false>(s, sc, o, d);
}

struct Harness {
  WideBvh bvh;
  SceneDev sc{};
  std::vector<uint8_t> storage[12];
  int max_stack = 0;
};

template <class T>
const T* keep(std::vector<uint8_t>& store, const HjkArray& a, size_t elem) {
  store.assign((const uint8_t*)a.ptr, (const uint8_t*)a.ptr + a.count * elem);
  if (store.empty()) store.resize(16);
  return (const T*)store.data();
}

struct FrameLayers {
  const float* l0;
  const float* l1;
  const float* l2;
  uint32_t W;
  f4 at(const float* base, uint32_t gx, uint32_t gy) const {
    const float* p = base + 4 * ((size_t)gy * W + gx);
    return F4(p[0], p[1], p[2], p[3]);
  }
  f4 radiance(uint32_t gx, uint32_t gy) const { return at(l0, gx, gy); }
  f4 feature(uint32_t gx, uint32_t gy) const { return at(l1, gx, gy); }
  f4 albedo(uint32_t gx, uint32_t gy) const { return at(l2, gx, gy); }
};

bool g_exact_ties = false;     // exact-tie mode of the harness (ht_set_exact)
uint64_t g_unresolved = 0;

// Work statistics of a tree (ht_work_stats): the same visiting order and acceptance rule as trav_run's default mode,
// with the node steps and primitive tests counted — what the trace kernel's time is made of.
bool g_count_work = false;
uint64_t g_work[6];  // closest-hit rays, their node steps, their primitive tests; the same for any-hit rays
template <int GUARD>
void count_ray(const SceneDev& sc, TravState& s, HostStack& stack) {
  uint64_t* w = g_work + ((s.slot >> 31) ? 3 : 0);
  w[0]++;
  for (;;) {
    if (s.ng_y > 0x00FFFFFFu) {
      const uint32_t hits_imask = s.ng_y;
      const int bit = hi_bit(hits_imask);
      s.ng_y &= ~(1u << bit);
      if (s.ng_y > 0x00FFFFFFu) stack.push(s.ng_x, s.ng_y);
      const uint32_t slot = ((uint32_t)bit - 24u) ^ (s.octinv4 & 0xFFu);
      const uint32_t rel = (uint32_t)pop_count(hits_imask & ~(0xFFFFFFFFu << slot));
      const f4* np = sc.nodes + (size_t)(s.ng_x + rel) * 5;
      w[1]++;
      const uint32_t hitmask = intersect_node<GUARD, false>(sc, s, np[0], np[1], np[2], np[3], np[4]);
      s.ng_x = x::as_uint(np[1].x);
      s.ng_y = (hitmask & 0xFF000000u) | (x::as_uint(np[0].w) >> 24);
      s.tg_x = x::as_uint(np[1].y) & kWidePrimBaseMask;
      s.tg_y = hitmask & 0x00FFFFFFu;
    } else {
      s.tg_x = s.ng_x, s.tg_y = s.ng_y;
      s.ng_x = s.ng_y = 0;
    }
    while (s.tg_y) {
      const int i = hi_bit(s.tg_y);
      s.tg_y &= ~(1u << i);
      const f4* pp = sc.prims + (size_t)(s.tg_x + (uint32_t)i) * HJK_PRIM_STRIDE;
      float t, u, v;
      w[2]++;
      if (intersect_prim(sc, s, pp[0], pp[1], pp[2], pp[HJK_PRIM_STRIDE - 1], t, u, v)) {
        if (s.slot >> 31) {
          s.hit_id = (int32_t)x::as_uint(pp[0].w);
          return;
        }
        if (closer_hit(s, t, x::as_uint(pp[0].w))) {
          s.hit_id = (int32_t)x::as_uint(pp[0].w);
          s.hit_t = t, s.hit_u = u, s.hit_v = v;
          s.tmax = t;
        }
      }
    }
    if (s.ng_y <= 0x00FFFFFFu) {
      if (stack.empty()) return;
      stack.pop(s.ng_x, s.ng_y);
    }
  }
}

template <bool ANY_HIT, class Policy>
bool run_ray(const SceneDev& sc, TravState& s, HostStack& st, float eps, const Policy& p) {
  s.slot = ANY_HIT ? 0x80000000u : 0u;
  if (g_count_work) {
    if (sc.num_spheres) count_ray<1>(sc, s, st); else count_ray<0>(sc, s, st);
    return true;
  }
  if (g_exact_ties) {
    const bool done = sc.num_spheres ? trav_run<true, true>(sc, s, st, eps, p, g_cands)
                                     : trav_run<false, true>(sc, s, st, eps, p, g_cands);
    if (done && g_cands.unresolved) g_unresolved++;
    return done;
  }
  NoCands none;
  return sc.num_spheres ? trav_run<true, false>(sc, s, st, eps, p, none) : trav_run<false, false>(sc, s, st, eps, p, none);
}

}  // namespace

extern "C" {

void* ht_create(const HjkScene* s, float pad_rel, char* err_out, int err_cap) {
  Harness* h = new Harness();
  std::string err;
  if (!build_wide_bvh(*s, pad_rel, h->bvh, err) || !validate_wide_bvh(*s, h->bvh, err)) {
    if (err_out && err_cap > 0) {
      strncpy(err_out, err.c_str(), err_cap - 1);
      err_out[err_cap - 1] = 0;
    }
    delete h;
    return nullptr;
  }
  SceneDev& sc = h->sc;
  sc.nodes = (const f4*)h->bvh.nodes.data();
  sc.prims = (const f4*)h->bvh.prims.data();
  const HjkSceneInfo* info = (const HjkSceneInfo*)s->scene.ptr;
  sc.spheres = keep<f4>(h->storage[0], s->spheres, 16);
  sc.quads = keep<f4>(h->storage[1], s->quads, 48);
  sc.triangles = keep<uint32_t>(h->storage[2], s->triangles, 12);
  sc.vertices = keep<f4>(h->storage[3], s->vertices, 32);
  sc.materials = keep<uint32_t>(h->storage[4], s->materials, 4);
  sc.emitters = keep<f4>(h->storage[5], s->emitters, 16);
  sc.diffuse = keep<f4>(h->storage[6], s->diffuse, 16);
  sc.diffusecb = keep<f4>(h->storage[7], s->diffusecb, 32);
  sc.dielectric = keep<f4>(h->storage[8], s->dielectric, 16);
  sc.emissive = keep<f4>(h->storage[9], s->emissive, 16);
  sc.num_spheres = info->num_spheres;
  sc.num_quads = info->num_quads;
  sc.num_triangles = info->num_triangles;
  sc.num_emitters = info->num_emitters;
  sc.camera = info->camera;
  for (int k = 0; k < 4; k++) sc.sph_centre[k] = h->bvh.sph_centre[k];
  sc.sph_rmin = h->bvh.sph_rmin, sc.sph_rmax = h->bvh.sph_rmax;
  return h;
}
void ht_destroy(void* p) { delete (Harness*)p; }
// node steps and primitive tests of the rays of a render (single-threaded): out6 = g_work
int ht_render(void* p, const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm, float* accumulator,
              float* layers_out, uint64_t* counts);
int ht_work_stats(void* p, const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm, uint64_t* out6) {
  Harness* h = (Harness*)p;
  const uint64_t W = blocks[0].original_dimension[0], H = blocks[0].original_dimension[1];
  std::vector<float> acc(W * H * 4, 0.f);
  HjkParams q = *prm;
  q.flags |= HJK_RENDER_NO_RECON;
  memset(g_work, 0, sizeof g_work);
  g_count_work = true;
  const int rc = ht_render(h, blocks, n_blocks, &q, acc.data(), nullptr, nullptr);
  g_count_work = false;
  memcpy(out6, g_work, sizeof g_work);
  return rc;
}
void ht_set_exact(int on) {
  g_exact_ties = on != 0;
  g_unresolved = 0;
}
uint64_t ht_unresolved(void) { return g_unresolved; }

void ht_bvh_stats(void* p, uint64_t* n_nodes, uint64_t* n_prims, uint32_t* depth, float* sah, float* pad,
                  int* max_stack) {
  Harness* h = (Harness*)p;
  *n_nodes = h->bvh.nodes.size();
  *n_prims = h->bvh.prims.size();
  *depth = h->bvh.depth;
  *sah = h->bvh.sah_cost;
  *pad = h->bvh.pad;
  *max_stack = h->max_stack;
}

// hjk_trace_first_hit semantics
void ht_trace(void* p, const HjkRay* rays, uint64_t n, int any_hit, float eps, int32_t* shape_id,
              float* t, float* uv) {
  Harness* h = (Harness*)p;
  for (uint64_t i = 0; i < n; i++) {
    TravState s;
    init_ray(h->sc, s, F4(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2], rays[i].t_min),
              F4(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2], rays[i].t_max));
    HostStack st;
    const TravNoPolicy never;
    const ChaosPolicy chaos{(uint32_t)i * 2654435761u + 12345u};
    if (any_hit & 1) {
      if (any_hit & 2) {
        while (!run_ray<true>(h->sc, s, st, eps, chaos)) {
        }
      } else {
        run_ray<true>(h->sc, s, st, eps, never);
      }
      shape_id[i] = s.hit_id >= 0 ? 1 : 0;
    } else {
      if (any_hit & 2) {
        while (!run_ray<false>(h->sc, s, st, eps, chaos)) {
        }
      } else {
        run_ray<false>(h->sc, s, st, eps, never);
      }
      shape_id[i] = s.hit_id;
    }
    if (st.max_n > h->max_stack) h->max_stack = st.max_n;
    if (t) t[i] = s.hit_id >= 0 ? s.hit_t : 0.f;
    if (uv) {
      uv[2 * i] = s.hit_id >= 0 ? s.hit_u : 0.f;
      uv[2 * i + 1] = s.hit_id >= 0 ? s.hit_v : 0.f;
    }
  }
}

// hjk_render semantics: integrates every pass into full-frame layers, then reconstructs it.
// layers_out (optional): 3 x W*H float4 of the LAST pass.  counts: paths, extension, shadow.
int ht_render(void* p, const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm,
              float* accumulator, float* layers_out, uint64_t* counts) {
  Harness* h = (Harness*)p;
  PassPlan plan;
  std::string err;
  if (!plan_passes(blocks, n_blocks, plan, err)) return -1;
  const uint32_t W = plan.width, H = plan.height;
  const size_t npx = (size_t)W * H;
  std::vector<float> layers(npx * 12, 0.f);
  float* l0 = layers.data();
  float* l1 = l0 + npx * 4;
  float* l2 = l1 + npx * 4;
  const int R = (int)prm->recon_radius, taps = 2 * R + 1;
  std::vector<float> weights((size_t)n_blocks * taps * taps);
  std::vector<uint32_t> tap_lists((size_t)n_blocks * recon_tap_stride(R));
  for (uint64_t b = 0; b < n_blocks; b++)
    recon_fill_block_tables(blocks[b], R, prm->recon_stddev, weights.data() + b * taps * taps,
                            tap_lists.data() + b * recon_tap_stride(R));
  uint64_t n_paths = 0, n_ext = 0, n_sh = 0;
  const size_t tiles = (size_t)plan.tiles_x * plan.tiles_y;
  for (size_t pi = 0; pi < plan.passes.size(); pi++) {
    const int32_t* tile_block = plan.tile_block.data() + pi * tiles;
    std::fill(layers.begin(), layers.end(), 0.f);
    for (uint32_t gy = 0; gy < H; gy++)
      for (uint32_t gx = 0; gx < W; gx++) {
        const int32_t b = tile_block[(size_t)(gy / plan.tile_h) * plan.tiles_x + gx / plan.tile_w];
        if (b < 0) continue;
        const HjkImageBlock& blk = blocks[b];
        const uint32_t lx = gx - blk.origin[0], ly = gy - blk.origin[1];
        if (lx >= blk.dimension[0] || ly >= blk.dimension[1]) continue;
        n_paths++;
        // raygen (render.glsl:149-162)
        VertexIn in;
        in.rng = seed_rng(blk.seed + lx + ly * blk.dimension[0]);
        camera_ray(h->sc.camera, x::add((float)gx, blk.sample_offset[0]),
                   x::add((float)gy, blk.sample_offset[1]), (float)W, (float)H, prm->eps, in.ray_o, in.ray_d);
        in.throughput = V3(1.f);
        in.extinction = V3(0.f);
        in.was_discrete = true;
        float* r = l0 + 4 * ((size_t)gy * W + gx);
        float* f = l1 + 4 * ((size_t)gy * W + gx);
        r[0] = r[1] = r[2] = 0.f;
        r[3] = 1.f;
        for (uint32_t bounce = 0; bounce < prm->max_bounces; bounce++) {
          in.bounce = bounce;
          // extend
          TravState s;
          init_ray(h->sc, s, in.ray_o, in.ray_d);
          HostStack st;
          const TravNoPolicy never;
          n_ext++;
          run_ray<false>(h->sc, s, st, prm->eps, never);
          if (s.hit_id < 0) break;
          in.hit_id = s.hit_id, in.hit_t = s.hit_t, in.hit_u = s.hit_u, in.hit_v = s.hit_v;
          // shade
          VertexOut o;
          shade_vertex(h->sc, in, prm->max_bounces, prm->rr_start, prm->eps, o);
          if (bounce == 0) f[0] = o.normal.x, f[1] = o.normal.y, f[2] = o.normal.z, f[3] = o.depth;
          if (o.add_emission) {
            r[0] = x::add(r[0], o.emission.x), r[1] = x::add(r[1], o.emission.y), r[2] = x::add(r[2], o.emission.z);
          }
          // shadow
          if (o.has_shadow) {
            n_sh++;
            TravState ss;
            init_ray(h->sc, ss, o.sh_o, o.sh_d);
            HostStack st2;
            run_ray<true>(h->sc, ss, st2, prm->eps, never);
            if (ss.hit_id < 0) {
              r[0] = x::add(r[0], o.contribution.x), r[1] = x::add(r[1], o.contribution.y),
              r[2] = x::add(r[2], o.contribution.z);
            }
          }
          if (!o.continues) break;
          in.ray_o = o.next_o, in.ray_d = o.next_d;
          in.throughput = o.throughput, in.extinction = o.extinction;
          in.rng = o.rng, in.was_discrete = o.was_discrete;
        }
      }
    if (!(prm->flags & HJK_RENDER_NO_RECON)) {
      PassDev ps;
      ps.width = W, ps.height = H, ps.tile_w = plan.tile_w, ps.tile_h = plan.tile_h;
      ps.tiles_x = plan.tiles_x, ps.tiles_y = plan.tiles_y;
      ps.tile_block = tile_block;
      ps.blocks = blocks;
      ps.weights = weights.data();
      ps.taps = tap_lists.data();
      ps.radius = R;
      FrameLayers L{l0, l1, l2, W};
      for (uint32_t gy = 0; gy < H; gy++)
        for (uint32_t gx = 0; gx < W; gx++) {
          float* a = accumulator + 4 * ((size_t)gy * W + gx);
          f4 v = reconstruct_pixel<false>(ps, L, gx, gy, F4(a[0], a[1], a[2], a[3]));
          a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w;
        }
    }
  }
  if (layers_out) memcpy(layers_out, layers.data(), layers.size() * 4);
  if (counts) counts[0] = n_paths, counts[1] = n_ext, counts[2] = n_sh;
  return 0;
}

// One pixel's path, bounce by bounce (debugging aid, mirrors orc_trace_path): out = (shape id, t bits,
// rng after, shadow state) per vertex.
int ht_trace_path(void* p, const HjkImageBlock* blk, uint32_t lx, uint32_t ly, const HjkParams* prm, int32_t* out,
                  int capacity, float* rays_out) {
  Harness* h = (Harness*)p;
  VertexIn in;
  const uint32_t gx = lx + blk->origin[0], gy = ly + blk->origin[1];
  in.rng = seed_rng(blk->seed + lx + ly * blk->dimension[0]);
  camera_ray(h->sc.camera, x::add((float)gx, blk->sample_offset[0]), x::add((float)gy, blk->sample_offset[1]),
             (float)blk->original_dimension[0], (float)blk->original_dimension[1], prm->eps, in.ray_o, in.ray_d);
  in.throughput = V3(1.f);
  in.extinction = V3(0.f);
  in.was_discrete = true;
  int n = 0;
  const TravNoPolicy never;
  for (uint32_t bounce = 0; bounce < prm->max_bounces && n < capacity; bounce++) {
    in.bounce = bounce;
    if (rays_out) {
      memcpy(rays_out + 8 * n, &in.ray_o, 16);
      memcpy(rays_out + 8 * n + 4, &in.ray_d, 16);
    }
    TravState s;
    init_ray(h->sc, s, in.ray_o, in.ray_d);
    HostStack st;
    run_ray<false>(h->sc, s, st, prm->eps, never);
    int32_t* o = out + 4 * n++;
    o[0] = s.hit_id;
    memcpy(&o[1], &s.hit_t, 4);
    o[2] = 0, o[3] = 0;
    if (s.hit_id < 0) break;
    in.hit_id = s.hit_id, in.hit_t = s.hit_t, in.hit_u = s.hit_u, in.hit_v = s.hit_v;
    VertexOut vo;
    shade_vertex(h->sc, in, prm->max_bounces, prm->rr_start, prm->eps, vo);
    o[2] = (int32_t)vo.rng;
    if (vo.has_shadow) {
      TravState ss;
      init_ray(h->sc, ss, vo.sh_o, vo.sh_d);
      HostStack st2;
      run_ray<true>(h->sc, ss, st2, prm->eps, never);
      o[3] = ss.hit_id < 0 ? 2 : 1;
    }
    if (!vo.continues) break;
    in.ray_o = vo.next_o, in.ray_d = vo.next_d;
    in.throughput = vo.throughput, in.extinction = vo.extinction;
    in.rng = vo.rng, in.was_discrete = vo.was_discrete;
  }
  return n;
}

// reconstruction of caller-supplied layers (hjk_denoise_pass semantics)
int ht_denoise(const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm, const float* radiance,
               const float* normal_depth, const float* albedo, float* accumulator) {
  PassPlan plan;
  std::string err;
  if (!plan_passes(blocks, n_blocks, plan, err)) return -1;
  const int R = (int)prm->recon_radius, taps = 2 * R + 1;
  std::vector<float> weights((size_t)n_blocks * taps * taps);
  std::vector<uint32_t> tap_lists((size_t)n_blocks * recon_tap_stride(R));
  for (uint64_t b = 0; b < n_blocks; b++)
    recon_fill_block_tables(blocks[b], R, prm->recon_stddev, weights.data() + b * taps * taps,
                            tap_lists.data() + b * recon_tap_stride(R));
  const size_t tiles = (size_t)plan.tiles_x * plan.tiles_y;
  for (size_t pi = 0; pi < plan.passes.size(); pi++) {
    PassDev ps;
    ps.width = plan.width, ps.height = plan.height, ps.tile_w = plan.tile_w, ps.tile_h = plan.tile_h;
    ps.tiles_x = plan.tiles_x, ps.tiles_y = plan.tiles_y;
    ps.tile_block = plan.tile_block.data() + pi * tiles;
    ps.blocks = blocks;
    ps.weights = weights.data();
    ps.taps = tap_lists.data();
    ps.radius = R;
    FrameLayers L{radiance, normal_depth, albedo, plan.width};
    for (uint32_t gy = 0; gy < plan.height; gy++)
      for (uint32_t gx = 0; gx < plan.width; gx++) {
        float* a = accumulator + 4 * ((size_t)gy * plan.width + gx);
        f4 v = albedo ? reconstruct_pixel<true>(ps, L, gx, gy, F4(a[0], a[1], a[2], a[3]))
                      : reconstruct_pixel<false>(ps, L, gx, gy, F4(a[0], a[1], a[2], a[3]));
        a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w;
      }
  }
  return 0;
}

// hjk_math.cuh evaluated on the host (compared bit-for-bit with oracle/orc_math.h)
void ht_math_eval(int fn, const float* a, const float* b, float* out, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) {
    switch (fn) {
      case 0: out[i] = sin_det(a[i]); break;
      case 1: out[i] = cos_det(a[i]); break;
      case 2: out[i] = tan_det(a[i]); break;
      case 3: out[i] = exp_det(a[i]); break;
      case 4: out[i] = atan2_det(a[i], b[i]); break;
      case 6: out[i] = exp_fma_neg(a[i]); break;
      default: out[i] = asin_det(a[i]); break;
    }
  }
}

}  // extern "C"
