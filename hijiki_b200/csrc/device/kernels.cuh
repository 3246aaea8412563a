// Wavefront path-tracing kernels for sm_100a — the B200 replacement of the reference's single
// render.glsl megakernel (reference shader/render.glsl:149-175) and of its per-block
// reconstruction dispatch (shader/reconstruction.glsl:22-66).
//
//   k_raygen      one thread per path slot: RNG seed, camera ray, layer initialisation
//   k_trace_coop  persistent warps pull the extension rays of bounce b (closest hit = "extend") and
//                 the shadow rays of bounce b-1 (any hit) from one work range and walk the 8-wide
//                 BVH, stack in shared memory (sized to the tree's depth at launch); the primitive tests of a warp are pooled over
//                 its 32 lanes when the per-lane counts are skewed (the default trace kernel)
//   k_trace       the same with a per-lane primitive loop: exact-tie mode, hjk_trace_first_hit
//   k_shade       per tile of the queue: misses dropped, hits counting-sorted by material tag in
//                 shared memory (ballot / prefix sum), then emission, next-event estimation, BSDF
//                 sample, roulette; the surviving path and its shadow ray are appended to their
//                 queues by block-level compaction
//   k_recon       shared-memory tiled bilateral splat of one or more passes into the accumulator
//
// Path state (ray, throughput, RNG) and hit records are QUEUE-ORDERED: element e belongs to the
// path at position e of the bounce's extension queue, so every stream a kernel touches is dense.
// No host synchronisation inside a wave: every kernel reads its element count from device
// counters written by the previous stage.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the driver entry point is fetched at run time)
#include <cuda_runtime.h>

#include <type_traits>

#include "recon.cuh"
#include "shade.cuh"
#include "traverse.cuh"

namespace hjk {

// per-bounce device counters (uint32 each)
enum : uint32_t {
  CTR_EXT = 0,      // extension rays entering this bounce
  CTR_SHADOW = 6,   // shadow rays emitted by this bounce
  CTR_EXT_CURSOR = 7,  // work cursor of k_trace(bounce): shadow rays of bounce-1, then extension rays
  CTR_STRIDE = 16
};

struct WaveDev {
  SceneDev scene;
  uint32_t width, height, n_pixels;
  uint32_t n_wave_passes, n_slots;
  uint32_t tile_w, tile_h, tiles_x, tiles_y;
  const int32_t* tile_block;     // [n_wave_passes][tiles_y*tiles_x] -> index into blocks
  const HjkImageBlock* blocks;   // the whole block list of the call
  const float* weights;          // [(2R+1)^2] per block
  // path state, queue-ordered: element e belongs to the path at position e of the bounce's extension
  // queue (compacted every bounce by k_shade; [bounce & 1] is read, [(bounce + 1) & 1] written)
  f4* ray_o[2];                  // origin.xyz, tMin
  f4* ray_d[2];                  // direction.xyz, tMax
  f4* thr_rng[2];                // throughput.rgb, rng state bits
  f4* extinction[2];             // currentExtinction.rgb (only when the scene can set it)
  f4* hit;                       // shape id bits, t, u, v
  // per-pass intermediate layers of the wave, [n_wave_passes][n_pixels]
  f4* layer0;                    // (radiance, 1)      render.glsl:172
  f4* layer1;                    // (normal, depth)    render.glsl:173
  // queues
  uint32_t* ext_q[2];            // slot | wasDiscrete << 31
  f4* sh_o;                      // shadow rays, dense
  f4* sh_d;
  f4* sh_c;                      // contribution.rgb, slot bits
  uint32_t* counters;            // [bounce][CTR_STRIDE]
  f4* accumulator;               // full frame (sum w*rgb, sum w)
  uint32_t max_bounces, rr_start;
  int32_t recon_radius;
  float eps;
  uint32_t has_extinction;
  uint32_t fetch_threshold;      // refill a warp when fewer lanes than this are busy
  uint32_t postpone_lanes;       // postpone primitive tests that fewer lanes than this would run
  uint32_t coop_batch_cost;      // pooled primitive tests: assumed instructions per batch of 32 (0 = always pool)
  uint32_t stack_cap;            // entries per thread of the launch's shared-memory traversal stack
  uint32_t* unresolved;          // exact-tie mode: rays whose tie cluster outgrew the window/list
};

#ifndef HJK_TRACE_COOP_MIN_BLOCKS
// 8 CTAs = 64 registers: no spills and (the stack being shared memory) no local memory at all.  At 9 CTAs = 56
// registers the kernel spills 36-52 bytes, and whether that costs nothing or 33 % of the kernel (32.1 vs 42.8 ms per
// step on cbox: local-memory sectors missing L1, long-scoreboard stalls x5) flipped with unrelated one-line changes
// elsewhere in the file (profiles/README.md r02d).  Measured 8 / 9: 32.07 / 32.19-42.75 ms.
#define HJK_TRACE_COOP_MIN_BLOCKS 8
#endif
#ifndef HJK_TRACE_MIN_BLOCKS
#define HJK_TRACE_MIN_BLOCKS 9 /* 56 registers, no spills: 36 warps per SM (measured best of 8/9/10/12) */
#endif
#ifndef HJK_TRAV_THREADS
#define HJK_TRAV_THREADS 128
#endif
constexpr int kTravThreads = HJK_TRAV_THREADS;
// The traversal stack lives entirely in shared memory, sized at launch to the tree at hand (dynamic shared memory:
// stack_cap entries of 8 bytes per thread, thread-interleaved): one node-group entry per level, plus one postponed
// primitive group per level in the per-lane kernel.  kMaxStack bounds the tree depth hjk_scene_upload accepts.
constexpr int kMaxStack = 32;
constexpr int kFetchThreshold = 20;  // refill a warp when fewer lanes than this are busy
constexpr int kPostponeLanes = 8;    // postpone primitive tests that fewer lanes than this would run
#ifndef HJK_TILE_THREADS
#define HJK_TILE_THREADS 256
#endif
constexpr int kTileThreads = HJK_TILE_THREADS;  // k_raygen
#ifndef HJK_SHADE_THREADS
#define HJK_SHADE_THREADS 128  // measured 64 / 128 / 256 / 512: k_shade 12.6 / 12.9 / 13.3 / 14.5 ms per step on cbox (64 loses on the sphere lattice)
#endif
constexpr int kShadeThreads = HJK_SHADE_THREADS;  // k_shade: tile of the extension queue sorted and shaded by one CTA
#ifndef HJK_SHADE_MIN_BLOCKS
#define HJK_SHADE_MIN_BLOCKS (1024 / HJK_SHADE_THREADS)
#endif

// ---------------------------------------------------------------- block-level compaction
// Every thread of the block calls this (flag may be false).  Returns the global position
// of the thread's element in the queue whose length lives at *counter.  One atomic per block.
template <int NQ>
struct BlockAppend {
  uint32_t warp_total[NQ][(kTileThreads > kShadeThreads ? kTileThreads : kShadeThreads) / 32];
  uint32_t base[NQ];
};
// TRAILING_SYNC = false when the caller passes at least one other block barrier before it calls
// block_append again (the shared scratch is then already safe to overwrite).
template <int NQ, bool TRAILING_SYNC = true>
__device__ __forceinline__ void block_append(BlockAppend<NQ>& sm, const bool (&flag)[NQ],
                                             uint32_t* const (&counter)[NQ], uint32_t (&pos)[NQ]) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  uint32_t prefix[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, flag[q]);
    prefix[q] = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) sm.warp_total[q][warp] = __popc(b);
  }
  __syncthreads();
  if (threadIdx.x < NQ) {
    uint32_t total = 0;
    for (uint32_t w = 0; w < n_warps; w++) {
      const uint32_t c = sm.warp_total[threadIdx.x][w];
      sm.warp_total[threadIdx.x][w] = total;
      total += c;
    }
    sm.base[threadIdx.x] = total ? atomicAdd(counter[threadIdx.x], total) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; q++) pos[q] = sm.base[q] + sm.warp_total[q][warp] + prefix[q];
  if (TRAILING_SYNC) __syncthreads();  // sm is reused by the next tile
}

// ---------------------------------------------------------------- raygen
// render.glsl:149-162 for every pixel of every block of the wave's passes.
__global__ void __launch_bounds__(kTileThreads) k_raygen(WaveDev w) {
  __shared__ BlockAppend<1> sm;
  const uint32_t n_tiles = (w.n_slots + kTileThreads - 1) / kTileThreads;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t slot = tile * kTileThreads + threadIdx.x;
    bool valid = false;
    f4 o = F4(0.f, 0.f, 0.f, 0.f), d = o;
    uint32_t rng = 0;
    if (slot < w.n_slots) {
      const uint32_t wp = slot / w.n_pixels, pix = slot - wp * w.n_pixels;
      const uint32_t gy = pix / w.width, gx = pix - gy * w.width;
      const int32_t b = w.tile_block[(size_t)wp * w.tiles_x * w.tiles_y + (gy / w.tile_h) * w.tiles_x + gx / w.tile_w];
      if (b >= 0) {
        const HjkImageBlock blk = w.blocks[b];
        const uint32_t lx = gx - blk.origin[0], ly = gy - blk.origin[1];
        if (lx < blk.dimension[0] && ly < blk.dimension[1]) {
          valid = true;
          rng = seed_rng(blk.seed + lx + ly * blk.dimension[0]);  // render.glsl:156
          camera_ray(w.scene.camera, x::add((float)gx, blk.sample_offset[0]),
                     x::add((float)gy, blk.sample_offset[1]), (float)blk.original_dimension[0],
                     (float)blk.original_dimension[1], w.eps, o, d);
        }
      }
      w.layer0[slot] = F4(0.f, 0.f, 0.f, valid ? 1.f : 0.f);
      w.layer1[slot] = F4(0.f, 0.f, 0.f, 0.f);
    }
    const bool flag[1] = {valid};
    uint32_t* const ctr[1] = {w.counters + CTR_EXT};
    uint32_t pos[1];
    block_append<1>(sm, flag, ctr, pos);
    if (valid) {  // the path state lives at the path's queue position
      const uint32_t e = pos[0];
      w.ext_q[0][e] = slot | 0x80000000u;  // wasDiscrete = true (render.glsl:91)
      w.ray_o[0][e] = o;
      w.ray_d[0][e] = d;
      w.thr_rng[0][e] = F4(1.f, 1.f, 1.f, __uint_as_float(rng));
      if (w.has_extinction) w.extinction[0][e] = F4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---------------------------------------------------------------- traversal
// Set when a traversal-stack entry did not fit (a lost entry can lose a hit): the launch sizes the stack to the tree
// (depth + 1 entries, twice that where primitive groups are postponed), so this stays 0.  A module-level flag
// written by one predicated store; read by hjk_get_info("stack_overflows").
__device__ unsigned int g_stack_overflows;
struct DevStack {
  uint2* sm;  // this thread's column of the CTA's shared-memory stack (stride kTravThreads)
  int n, cap;
  __device__ __forceinline__ void push(uint32_t a, uint32_t b) {
    if (n < cap)
      sm[n * kTravThreads] = make_uint2(a, b);
    else
      g_stack_overflows = 1u;
    n++;
  }
  __device__ __forceinline__ void pop(uint32_t& a, uint32_t& b) {
    n--;
    const uint2 v = sm[(n < cap ? n : cap - 1) * kTravThreads];
    a = v.x, b = v.y;
  }
  __device__ __forceinline__ bool empty() const { return n == 0; }
};
extern __shared__ __align__(16) uint2 dyn_trav_stack[];  // [stack_cap][kTravThreads]

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Source of rays / sink of results of a traversal launch.  A ray's flavour rides in the top bit of
// TravState::slot: 1 = any hit (shadow ray), 0 = closest hit (extension ray).
constexpr uint32_t kAnyHitBit = 0x80000000u;

// One bounce step of the wave: the shadow rays emitted by bounce-1 and the extension rays of
// `bounce` are drawn from ONE work range [0, n_shadow + n_ext) by the same persistent warps, so a
// launch has one tail instead of two and shadow rays fill the lanes extension rays leave idle.
struct WaveIO {
  const WaveDev& w;
  const f4* ray_o;  // extension rays of this bounce, in queue order
  const f4* ray_d;
  uint32_t n_shadow;
  template <int GUARD>
  __device__ __forceinline__ void load(uint32_t i, TravState& s) const {
    if (i < n_shadow) {  // dense shadow queue
      s.slot = i | kAnyHitBit;
      trav_init<GUARD>(s, w.scene, w.sh_o[i], w.sh_d[i]);
    } else {
      const uint32_t e = i - n_shadow;  // queue position: ray and hit record live there (dense, no indirection)
      s.slot = e;
      trav_init<GUARD>(s, w.scene, ray_o[e], ray_d[e]);
    }
  }
  __device__ __forceinline__ void store(const TravState& s) const {
    if (s.slot & kAnyHitBit) {
      if (s.hit_id >= 0) return;  // occluded (render.glsl:122)
      const f4 c = w.sh_c[s.slot & ~kAnyHitBit];
      const uint32_t slot = __float_as_uint(c.w);
      f4 r = w.layer0[slot];      // total += throughput * evalBSDF * importance (render.glsl:123)
      r.x = x::add(r.x, c.x), r.y = x::add(r.y, c.y), r.z = x::add(r.z, c.z);
      w.layer0[slot] = r;
    } else {
      w.hit[s.slot] = F4(__int_as_float(s.hit_id), s.hit_t, s.hit_u, s.hit_v);
    }
  }
};
// Standalone ray batch (hjk_trace_first_hit): rays and results are indexed by ray number.
struct BatchIO {
  const SceneDev& sc;
  const f4* ray_o;
  const f4* ray_d;
  f4* hit;
  uint32_t flavour;  // kAnyHitBit or 0 for the whole batch
  template <int GUARD>
  __device__ __forceinline__ void load(uint32_t i, TravState& s) const {
    s.slot = i | flavour;
    trav_init<GUARD>(s, sc, ray_o[i], ray_d[i]);
  }
  __device__ __forceinline__ void store(const TravState& s) const {
    hit[s.slot & ~kAnyHitBit] = F4(__int_as_float(s.hit_id), s.hit_t, s.hit_u, s.hit_v);
  }
};

struct WarpPolicy {
  bool can_refill;
  int fetch_threshold, postpone_lanes;
  __device__ __forceinline__ bool yield() const { return can_refill && __popc(__activemask()) < fetch_threshold; }
  __device__ __forceinline__ bool postpone() const { return __popc(__activemask()) < postpone_lanes; }
};

// Persistent warps: each warp keeps its 32 lanes supplied with rays from the work range; a lane
// whose ray finishes is refilled as soon as fewer than `fetch_threshold` lanes are busy.
template <int GUARD, bool EXACT, class IO>
__device__ __forceinline__ void traverse_queue(const SceneDev& sc, const IO& io, uint32_t n,
                                               uint32_t* cursor, float eps, int fetch_threshold,
                                               int postpone_lanes, uint32_t* unresolved, int stack_cap) {
  const uint32_t lane = threadIdx.x & 31u;
  DevStack st;
  st.sm = dyn_trav_stack + threadIdx.x;
  st.n = 0;
  st.cap = stack_cap;
  TravState s;
  typename std::conditional<EXACT, TieCands, NoCands>::type cands;
  cands.reset();
  bool active = false, exhausted = false;
  for (;;) {
    if (!exhausted) {
      const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active);
      if (need) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(need));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (!active) {
          const uint32_t i = base + __popc(need & ((1u << lane) - 1u));
          if (i < n) {
            io.template load<GUARD>(i, s);
            cands.reset();
            st.n = 0;
            active = true;
          }
        }
        if (base + __popc(need) >= n) exhausted = true;
      }
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;
    if (active) {
      const WarpPolicy policy{!exhausted, fetch_threshold, postpone_lanes};
      const bool done = trav_run<GUARD, EXACT>(sc, s, st, eps, policy, cands);
      if (done) {
        if (EXACT && cands_unresolved(cands)) atomicAdd(unresolved, 1u);
        io.store(s);
        active = false;
      }
    }
    __syncwarp();
  }
}

// Warp-cooperative variant.  Node steps stay per lane (one ray per lane, 26 of 32 lanes busy), but
// the primitive tests of a step are pooled: the pending (ray, primitive) pairs of all 32 lanes are
// numbered with a warp prefix sum and tested 32 at a time, one pair per lane, by whichever lane is
// free — the owner's ray travels by shuffle, the nearest accepted hit comes back through a 64-bit
// shared-memory atomicMin keyed by (t, testing lane).  In the per-lane loop the same tests ran at ~9
// of 32 lanes (lanes of one warp reach leaves holding 0..24 primitives at the same step).
// Closest hit = closer_hit (traverse.cuh): the candidate of smallest t, equal t by the lower shape id —
// the same function of the hit set whether a primitive is tested by its own lane or by another one.
// The exact-tie mode keeps the per-lane loop (it must record every candidate).
template <int GUARD, class IO>
__device__ __forceinline__ void traverse_queue_coop(const SceneDev& sc, const IO& io, uint32_t n, uint32_t* cursor,
                                                    float eps, int fetch_threshold, uint32_t coop_batch_cost,
                                                    int stack_cap) {
  __shared__ unsigned long long sm_best[kTravThreads];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t FULL = 0xFFFFFFFFu;
  DevStack st;
  st.sm = dyn_trav_stack + threadIdx.x;
  st.n = 0;
  st.cap = stack_cap;
  TravState s;
  s.tg_y = 0, s.ng_y = 0;
  bool active = false, exhausted = false;
  for (;;) {
    // ---- refill
    const uint32_t busy = __ballot_sync(FULL, active);
    if (!exhausted && (busy == 0u || __popc(busy) < fetch_threshold)) {  // an idle warp always refills (threshold 0)
      const uint32_t need = ~busy;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(need));
      base = __shfl_sync(FULL, base, 0);
      if (!active) {
        const uint32_t i = base + __popc(need & ((1u << lane) - 1u));
        if (i < n) {
          io.template load<GUARD>(i, s);
          st.n = 0;
          active = true;
        }
      }
      if (base + __popc(need) >= n) exhausted = true;
    }
    if (__ballot_sync(FULL, active) == 0u) {
      if (exhausted) break;
      continue;
    }
    // ---- node step (per lane)
    if (active) {
      if (s.ng_y > 0x00FFFFFFu) {
        const uint32_t hits_imask = s.ng_y;
        const int bit = hi_bit(hits_imask);
        s.ng_y &= ~(1u << bit);
        if (s.ng_y > 0x00FFFFFFu) st.push(s.ng_x, s.ng_y);
        const uint32_t slot = ((uint32_t)bit - 24u) ^ (s.octinv4 & 0xFFu);
        const uint32_t rel = (uint32_t)__popc(hits_imask & ~(0xFFFFFFFFu << slot));
        const f4* np = sc.nodes + (size_t)(s.ng_x + rel) * 5;
        const f4 q0 = ld16(np), q1 = ld16(np + 1), q2 = ld16(np + 2), q3 = ld16(np + 3), q4 = ld16(np + 4);
        const uint32_t hitmask = intersect_node<GUARD, false>(sc, s, q0, q1, q2, q3, q4);
        s.ng_x = __float_as_uint(q1.x);
        s.ng_y = (hitmask & 0xFF000000u) | (__float_as_uint(q0.w) >> 24);
        s.tg_x = __float_as_uint(q1.y) & kWidePrimBaseMask;
        s.tg_y = hitmask & 0x00FFFFFFu;
      } else {
        s.tg_y = 0;
      }
    }
    // ---- primitive tests: pooled over the warp when the per-lane counts are skewed enough to pay for
    // the pooling overhead (~70 instructions per batch of 32), else each lane tests its own.  (Re-evaluating the
    // rule after every per-lane round, to pool only the tail of a skewed warp, measured slower: 33.4 vs 31.95 ms
    // per step on cbox, 38.5 vs 36.3 on the sphere lattice — two warp reductions per round cost more than the
    // idle lanes of the tail.)
    const uint32_t cnt = active ? (uint32_t)__popc(s.tg_y) : 0u;
    const uint32_t sum_cnt = __reduce_add_sync(FULL, cnt), max_cnt = __reduce_max_sync(FULL, cnt);
    if (max_cnt * 71u <= ((sum_cnt + 31u) >> 5) * coop_batch_cost) {
      while (s.tg_y) {
        const int i = hi_bit(s.tg_y);
        s.tg_y &= ~(1u << i);
        const f4* pp = sc.prims + (size_t)(s.tg_x + (uint32_t)i) * HJK_PRIM_STRIDE;
        const f4 r0 = ld16(pp), r1 = ld16(pp + 1), r2 = ld16(pp + 2);
#if HJK_PRIM_STRIDE == 4
        const f4 r3 = ld16(pp + 3);
#else
        const f4 r3 = r2;
#endif
        float t, u, v;
        if (intersect_prim(sc, s, r0, r1, r2, r3, t, u, v) && closer_hit(s, t, __float_as_uint(r0.w))) {
          s.hit_id = (int32_t)__float_as_uint(r0.w);
          s.hit_t = t, s.hit_u = u, s.hit_v = v;
          if (s.slot >> 31) break;
          s.tmax = t;
        }
      }
    } else {
      const uint32_t pooled = cnt;
      uint32_t incl = pooled;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if ((int)lane >= o) incl += v;
      }
      const uint32_t total = sum_cnt;
      const uint32_t excl = incl - pooled;
      // key = (t bits, low 27 bits of the shape id, testing lane): the minimum is the closest hit, equal t
      // by the lower id (closer_hit); the lane bits only tell the owner where u, v and the full id are
      unsigned long long cur_key = ~0ull;
      if (active && s.hit_id >= 0)
        cur_key = ((unsigned long long)__float_as_uint(s.hit_t) << 32) | (((uint32_t)s.hit_id & 0x07FFFFFFu) << 5) | 31u;
      sm_best[threadIdx.x] = ~0ull;
      __syncwarp();
      for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t j = base + lane;
        const uint32_t jj = j < total ? j : total - 1u;
        // owner = first lane whose inclusive count exceeds jj
        uint32_t owner = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const uint32_t v = __shfl_sync(FULL, incl, (int)(owner + step - 1u));
          if (v <= jj) owner += step;
        }
        const uint32_t o_excl = __shfl_sync(FULL, excl, (int)owner);
        uint32_t bits = __shfl_sync(FULL, s.tg_y, (int)owner);
        const uint32_t o_tgx = __shfl_sync(FULL, s.tg_x, (int)owner);
        TravState r;
        r.ox = __shfl_sync(FULL, s.ox, (int)owner), r.oy = __shfl_sync(FULL, s.oy, (int)owner);
        r.oz = __shfl_sync(FULL, s.oz, (int)owner), r.dx = __shfl_sync(FULL, s.dx, (int)owner);
        r.dy = __shfl_sync(FULL, s.dy, (int)owner), r.dz = __shfl_sync(FULL, s.dz, (int)owner);
        r.tmin = __shfl_sync(FULL, s.tmin, (int)owner), r.tmax = __shfl_sync(FULL, s.tmax, (int)owner);
        float t = 0.f, u = 0.f, v = 0.f;
        uint32_t id = 0;
        if (j < total) {
          for (uint32_t k = jj - o_excl; k > 0; k--) bits &= bits - 1u;  // k-th pending primitive of the owner
          const uint32_t prim_index = o_tgx + (uint32_t)(__ffs((int)bits) - 1);
          const f4* pp = sc.prims + (size_t)prim_index * HJK_PRIM_STRIDE;
          const f4 r0 = ld16(pp), r1 = ld16(pp + 1), r2 = ld16(pp + 2);
#if HJK_PRIM_STRIDE == 4
          const f4 r3 = ld16(pp + 3);
#else
          const f4 r3 = r2;
#endif
          if (intersect_prim(sc, r, r0, r1, r2, r3, t, u, v)) {
            id = __float_as_uint(r0.w);
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(t) << 32) | ((id & 0x07FFFFFFu) << 5) | lane;
            atomicMin(&sm_best[(threadIdx.x & ~31u) + owner], key);
          }
        }
        __syncwarp();
        // owners pick up an improvement made by this batch; the winner's u, v, id come by shuffle
        unsigned long long key = cur_key;
        if (pooled) key = sm_best[threadIdx.x];
        const bool improved = (key >> 5) < (cur_key >> 5);
        const int src = improved ? (int)(key & 31ull) : (int)lane;
        const float wu = __shfl_sync(FULL, u, src), wv = __shfl_sync(FULL, v, src);
        const uint32_t wid = __shfl_sync(FULL, id, src);
        if (improved) {
          cur_key = key;
          s.hit_id = (int32_t)wid;
          s.hit_t = __uint_as_float((uint32_t)(key >> 32));
          s.hit_u = wu, s.hit_v = wv;
          s.tmax = s.hit_t;
        }
        __syncwarp();
      }
    }
    // ---- advance / finish (per lane)
    if (active) {
      s.tg_y = 0;
      bool done = (s.slot >> 31) && s.hit_id >= 0;  // any-hit ray: first accepted primitive ends it
      if (!done && s.ng_y <= 0x00FFFFFFu) {
        if (st.empty()) done = true; else st.pop(s.ng_x, s.ng_y);
      }
      if (done) {
        io.store(s);
        active = false;
      }
    }
  }
}

// GUARD: the scene contains spheres (sphere guard of traverse.cuh compiled in).
// bounce in [0, max_bounces]: extension rays of `bounce` (none at max_bounces) + shadow rays of bounce-1.
// Without the sphere guard the kernel fits 56 registers (9 CTAs = 36 warps per SM, measured best of
// 8/9/10/12); the guard variants would spill at 56 and keep 64 registers (8 CTAs).
template <int GUARD, bool EXACT>
__global__ void __launch_bounds__(kTravThreads, GUARD ? 8 : HJK_TRACE_MIN_BLOCKS)
    k_trace(WaveDev w, uint32_t bounce, uint32_t last) {
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const uint32_t n_ext = bounce < last ? ctr[CTR_EXT] : 0u;
  const uint32_t n_shadow = bounce > 0 ? w.counters[(size_t)(bounce - 1) * CTR_STRIDE + CTR_SHADOW] : 0u;
  const WaveIO io{w, w.ray_o[bounce & 1u], w.ray_d[bounce & 1u], n_shadow};
  traverse_queue<GUARD, EXACT>(w.scene, io, n_shadow + n_ext, ctr + CTR_EXT_CURSOR, w.eps, (int)w.fetch_threshold,
                               (int)w.postpone_lanes, w.unresolved, (int)w.stack_cap);
}
// same work, pooled primitive tests (traverse_queue_coop)
template <int GUARD>
__global__ void __launch_bounds__(kTravThreads, GUARD ? 8 : HJK_TRACE_COOP_MIN_BLOCKS) k_trace_coop(WaveDev w, uint32_t bounce, uint32_t last) {
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const uint32_t n_ext = bounce < last ? ctr[CTR_EXT] : 0u;
  const uint32_t n_shadow = bounce > 0 ? w.counters[(size_t)(bounce - 1) * CTR_STRIDE + CTR_SHADOW] : 0u;
  const WaveIO io{w, w.ray_o[bounce & 1u], w.ray_d[bounce & 1u], n_shadow};
  traverse_queue_coop<GUARD>(w.scene, io, n_shadow + n_ext, ctr + CTR_EXT_CURSOR, w.eps, (int)w.fetch_threshold,
                             w.coop_batch_cost, (int)w.stack_cap);
}
// cursor[0] = work cursor, cursor[1] = unresolved-tie counter (exact mode)
template <int GUARD, bool EXACT>
__global__ void __launch_bounds__(kTravThreads) k_trace_batch(SceneDev sc, const f4* ray_o, const f4* ray_d,
                                                              f4* hit, uint32_t n, uint32_t* cursor, float eps,
                                                              uint32_t flavour, int postpone_lanes, int stack_cap) {
  const BatchIO io{sc, ray_o, ray_d, hit, flavour};
  traverse_queue<GUARD, EXACT>(sc, io, n, cursor, eps, kFetchThreshold, postpone_lanes, cursor + 1, stack_cap);
}

// ---------------------------------------------------------------- sort + shade
// Material-sorted shading of one bounce.  Each CTA takes a tile of the bounce's extension queue,
// drops the misses (render.glsl:94-96), counting-sorts the hits of the tile by material tag in
// shared memory (warp ballots + prefix sums) so that its warps shade one material each, then runs
// one bounce-loop iteration per hit and appends the next extension ray and the shadow ray to
// their queues by block-level compaction.
struct TileSort {
  uint32_t warp_count[5][kShadeThreads / 32];
  uint32_t n_hits;
  uint32_t entry[kShadeThreads];
  uint32_t src[kShadeThreads];  // position in the tile before the sort (the path state is read from there)
  f4 hit[kShadeThreads];
};

// Front part of a tile: the thread's queue entry, its hit record and material tag (0xFFFFFFFF = miss or
// past the end).  Issued one tile ahead, between the two barriers of the previous tile's queue append, so the
// loads travel while that tile waits for its append atomics.
__device__ __forceinline__ void shade_tile_front(const WaveDev& w, const uint32_t* q, uint32_t n, uint32_t tile,
                                                 uint32_t bounce, uint32_t& entry, f4& h, uint32_t& tag) {
  const uint32_t i = tile * kShadeThreads + threadIdx.x;
  entry = 0, tag = 0xFFFFFFFFu;
  h = F4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    entry = q[i];
    h = w.hit[i];  // k_trace wrote the hit records in queue order: both loads are dense and independent
    const int id = __float_as_int(h.x);
    if (id >= 0) tag = ld4(w.scene.materials + id) >> HJK_MATERIAL_TAG_SHIFT;
  }
  // start the tile after this one towards L1
  const uint32_t i_next = i + gridDim.x * kShadeThreads;
  if (i_next < n) {
    if ((threadIdx.x & 7u) == 0) prefetch_l1(q + i_next);  // 8 entries per 32-byte sector
    if ((threadIdx.x & 1u) == 0) {
      prefetch_l1(w.hit + i_next);
      prefetch_l1(w.ray_o[bounce & 1u] + i_next), prefetch_l1(w.ray_d[bounce & 1u] + i_next);
      prefetch_l1(w.thr_rng[bounce & 1u] + i_next);
    }
  }
}

// SORT: counting-sort the hits of a tile by material tag first, so that a warp shades one material (scenes that mix
// diffuse, mirror and dielectric surfaces).  Where every surface a path can continue from is diffuse-like (cbox,
// the terrain) the sort only costs its three barriers and the shared-memory round trip — 12.75 vs 11.82 ms per step
// on cbox — and each thread shades the hit of its own queue entry instead (hjk_scene_upload decides).
template <bool SORT>
__global__ void __launch_bounds__(kShadeThreads, HJK_SHADE_MIN_BLOCKS) k_shade(WaveDev w, uint32_t bounce) {
  __shared__ BlockAppend<2> sm;
  __shared__ TileSort ts;
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const uint32_t n = ctr[CTR_EXT];
  const uint32_t* q = w.ext_q[bounce & 1u];
  uint32_t* next_q = w.ext_q[(bounce + 1u) & 1u];
  uint32_t* const counter[2] = {ctr + CTR_STRIDE + CTR_EXT, ctr + CTR_SHADOW};
  const uint32_t n_tiles = (n + kShadeThreads - 1) / kShadeThreads;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, par = bounce & 1u;
  uint32_t entry = 0, tag = 0xFFFFFFFFu;
  f4 h = F4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x < n_tiles) shade_tile_front(w, q, n, blockIdx.x, bounce, entry, h, tag);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    bool want_next = false, want_shadow = false, shade_this = false;
    VertexOut out;
    uint32_t slot = 0, e = 0;
    if (SORT) {
    // ---- tile-local material sort
    uint32_t prefix = 0;
#pragma unroll
    for (uint32_t t = 0; t < 5; t++) {
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, tag == t);
      if (tag == t) prefix = __popc(b & ((1u << lane) - 1u));
      if (lane == 0) ts.warp_count[t][warp] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive scan of the 5 x 8 warp counts, tag-major
      uint32_t total = 0;
      for (uint32_t t = 0; t < 5; t++) {
        for (uint32_t wi = 0; wi < kShadeThreads / 32; wi++) {
          const uint32_t c = ts.warp_count[t][wi];
          ts.warp_count[t][wi] = total;
          total += c;
        }
      }
      ts.n_hits = total;
    }
    __syncthreads();
    if (tag < 5u) {
      const uint32_t pos = ts.warp_count[tag][warp] + prefix;
      ts.entry[pos] = entry;
      ts.src[pos] = threadIdx.x;
      ts.hit[pos] = h;
    }
    __syncthreads();
    const uint32_t n_hits = ts.n_hits;

    // ---- one bounce-loop iteration per hit
    shade_this = threadIdx.x < n_hits;
    if (shade_this) {
      entry = ts.entry[threadIdx.x];
      e = tile * kShadeThreads + ts.src[threadIdx.x];
      h = ts.hit[threadIdx.x];
    }
    } else {  // every thread shades the hit of its own queue entry (misses idle)
      shade_this = tag < 5u;
      e = tile * kShadeThreads + threadIdx.x;
    }
    if (shade_this) {
      slot = entry & 0x7FFFFFFFu;
      VertexIn in;
      in.ray_o = w.ray_o[par][e];
      in.ray_d = w.ray_d[par][e];
      in.hit_id = __float_as_int(h.x), in.hit_t = h.y, in.hit_u = h.z, in.hit_v = h.w;
      const f4 tr = w.thr_rng[par][e];
      in.throughput = xyz(tr);
      in.rng = __float_as_uint(tr.w);
      in.extinction = w.has_extinction ? xyz(w.extinction[par][e]) : V3(0.f);
      in.was_discrete = (entry >> 31) != 0u;
      in.bounce = bounce;
      shade_vertex(w.scene, in, w.max_bounces, w.rr_start, w.eps, out);
      if (bounce == 0) w.layer1[slot] = F4(out.normal.x, out.normal.y, out.normal.z, out.depth);
      if (out.add_emission) {
        f4 r = w.layer0[slot];
        r.x = x::add(r.x, out.emission.x), r.y = x::add(r.y, out.emission.y), r.z = x::add(r.z, out.emission.z);
        w.layer0[slot] = r;
      }
      want_next = out.continues;
      want_shadow = out.has_shadow;
    }

    // ---- append to the next extension queue and the shadow queue: one atomic per CTA per queue.  The next
    // tile's front loads are issued between the two barriers, while threads 0 and 1 wait for the atomics.
    const uint32_t b_next = __ballot_sync(0xFFFFFFFFu, want_next), b_sh = __ballot_sync(0xFFFFFFFFu, want_shadow);
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t pre_next = __popc(b_next & below), pre_sh = __popc(b_sh & below);
    if (lane == 0) sm.warp_total[0][warp] = __popc(b_next), sm.warp_total[1][warp] = __popc(b_sh);
    __syncthreads();
    if (threadIdx.x < 2) {
      uint32_t total = 0;
      for (uint32_t wi = 0; wi < kShadeThreads / 32; wi++) {
        const uint32_t c = sm.warp_total[threadIdx.x][wi];
        sm.warp_total[threadIdx.x][wi] = total;
        total += c;
      }
      sm.base[threadIdx.x] = total ? atomicAdd(counter[threadIdx.x], total) : 0u;
    }
    const uint32_t tile_next = tile + gridDim.x;
    const bool was_discrete = out.was_discrete;
    if (tile_next < n_tiles) shade_tile_front(w, q, n, tile_next, bounce, entry, h, tag);
    __syncthreads();
    if (want_next) {  // the surviving path moves to its position in the next queue
      const uint32_t e = sm.base[0] + sm.warp_total[0][warp] + pre_next;
      next_q[e] = slot | (was_discrete ? 0x80000000u : 0u);
      w.ray_o[par ^ 1u][e] = out.next_o;
      w.ray_d[par ^ 1u][e] = out.next_d;
      w.thr_rng[par ^ 1u][e] = F4(out.throughput.x, out.throughput.y, out.throughput.z, __uint_as_float(out.rng));
      if (w.has_extinction) w.extinction[par ^ 1u][e] = F4(out.extinction.x, out.extinction.y, out.extinction.z, 0.f);
    }
    if (want_shadow) {
      const uint32_t pos = sm.base[1] + sm.warp_total[1][warp] + pre_sh;
      w.sh_o[pos] = out.sh_o;
      w.sh_d[pos] = out.sh_d;
      w.sh_c[pos] = F4(out.contribution.x, out.contribution.y, out.contribution.z, __uint_as_float(slot));
    }
    // (sm and ts are next written after the following tile's sort barriers / read before them)
  }
}

// ---------------------------------------------------------------- reconstruction
__global__ void k_recon_weights(const HjkImageBlock* blocks, uint32_t n_blocks, int radius, float stddev,
                                float* weights, uint32_t* taps) {
  const int t2 = (2 * radius + 1) * (2 * radius + 1);
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += gridDim.x * blockDim.x)
    recon_fill_block_tables(blocks[b], radius, stddev, weights + (size_t)b * t2,
                            taps + (size_t)b * recon_tap_stride(radius));
}

#ifndef HJK_RECON_TILE_Y
#define HJK_RECON_TILE_Y 8
#endif
#ifndef HJK_RECON_PRODUCER_SLEEP
#define HJK_RECON_PRODUCER_SLEEP 400 /* ns between the producer's polls of an empty barrier */
#endif
#ifndef HJK_RECON_UNROLL
#define HJK_RECON_UNROLL 1 /* tap pairs in flight per thread in k_recon's paired loop */
#endif
constexpr int kReconUnroll = HJK_RECON_UNROLL;
#ifndef HJK_RECON_MIN_BLOCKS
#define HJK_RECON_MIN_BLOCKS 3
#endif
constexpr int kReconTileX = 32, kReconTileY = HJK_RECON_TILE_Y;  // recon_smem_pitch() assumes 32

struct SmemLayers {  // tile + halo staged in shared memory
  const f4* s0;
  const f4* s1;
  const f4* s2;
  int x0, y0, pitch;  // global coordinate of smem element (0,0)
  __device__ __forceinline__ f4 radiance(uint32_t gx, uint32_t gy) const {
    return s0[((int)gy - y0) * pitch + ((int)gx - x0)];
  }
  __device__ __forceinline__ f4 feature(uint32_t gx, uint32_t gy) const {
    return s1[((int)gy - y0) * pitch + ((int)gx - x0)];
  }
  __device__ __forceinline__ f4 albedo(uint32_t gx, uint32_t gy) const {
    return s2[((int)gy - y0) * pitch + ((int)gx - x0)];
  }
};

// Interior texels (at least R away from every edge of their block: 94 % of a 128 x 128 block at R = 2) see
// exactly one block and all of its taps: same taps, same order, same arithmetic as reconstruct_pixel, without
// the per-tap bounds tests, coordinate unpacking and tile lookups (the shared-memory offset of a tap comes
// precomputed in the tap table): 18 % fewer instructions.  The kernel is latency-bound, not issue-bound:
// what paid was occupancy — 32 x 8 tiles at 40 registers, 6 CTAs = 48 warps per SM (990 -> 1244 GB/s at 4K;
// measured 16-row tiles x 2/3/4 CTAs and 8-row tiles x 4/5/6/7/8 CTAs); keeping two taps in flight per
// thread changed nothing.
// 16-byte shared-memory load by 32-bit shared-space address (no generic-address arithmetic per tap).  volatile:
// keeps its place after the mbarrier wait that publishes the tile.
__device__ __forceinline__ f4 lds16(uint32_t addr) {
  f4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// One tap of an interior texel.  a0 = shared address of the centre texel in layer 0, layer_bytes = distance to
// the same texel in the next layer; the tap's byte offset comes precomputed in the top half of tw.y.
template <bool HAS_ALBEDO>
__device__ __forceinline__ f4 recon_tap_weighted(uint2 tw, uint32_t a0, uint32_t layer_bytes, vec3 nc, vec3 ac) {
  const uint32_t a = a0 + (uint32_t)((int)tw.y >> 16);
  float w = __uint_as_float(tw.x);
  const f4 cw = lds16(a);
  const vec3 no = xyz(lds16(a + layer_bytes)) - nc;
  float e = x::mul(dot(no, no), 2.0f);
  if (HAS_ALBEDO) {
    const vec3 ao = xyz(lds16(a + 2u * layer_bytes)) - ac;
    e = x::add(e, dot(ao, ao));
  }
  w = x::mul(w, exp_fma_neg(e));  // exp(-e) in the FMA specification (hjk_math.cuh); exp_fma_neg(0) == 1
  return F4(x::mul(w, cw.x), x::mul(w, cw.y), x::mul(w, cw.z), x::mul(w, cw.w));
}
__device__ __forceinline__ f4 recon_accumulate(f4 acc, f4 wv) {  // reconstruction.glsl:55-58: NaN samples are dropped
  // two unordered compares cover the four components (FSETP.NAN takes two operands)
  uint32_t bad;
  asm("{\n\t.reg .pred p;\n\tsetp.nan.f32 p, %1, %2;\n\tsetp.nan.or.f32 p, %3, %4, p;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(bad)
      : "f"(wv.x), "f"(wv.y), "f"(wv.z), "f"(wv.w));
  if (bad) return acc;
  return F4(x::add(acc.x, wv.x), x::add(acc.y, wv.y), x::add(acc.z, wv.z), x::add(acc.w, wv.w));
}
template <bool HAS_ALBEDO>
__device__ __forceinline__ f4 reconstruct_interior(const uint32_t* tl, uint32_t a0, uint32_t layer_bytes, f4 acc) {
  const vec3 nc = xyz(lds16(a0 + layer_bytes));
  const vec3 ac = HAS_ALBEDO ? xyz(lds16(a0 + 2u * layer_bytes)) : V3(0.f);
  const uint32_t n = tl[0];
  const uint2* tp = reinterpret_cast<const uint2*>(tl + 2);
#pragma unroll 4
  for (uint32_t k = 0; k < n; k++)
    acc = recon_accumulate(acc, recon_tap_weighted<HAS_ALBEDO>(__ldg(tp + k), a0, layer_bytes, nc, ac));
  return acc;
}

// ---- R = 2 (the reference's radius, src/main.rs:1284) on packed fp32 pairs.  sm_100 executes add/mul/fma on two
// independent fp32 lanes of a 64-bit register pair (FADD2 / FMUL2 / FFMA2): every lane is the same correctly
// rounded IEEE operation as its scalar form, so the arithmetic below is reconstruction.glsl's, operation for
// operation — it only takes fewer issue slots (the kernel is issue-bound: ~930 instructions per texel in scalar
// form).  Packed: (x, y) and (z, w) of a tap's radiance and normal, as they arrive from one 16-byte shared-memory
// load, and the bilateral exponential of TWO taps at a time.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 splat2(float v) { return pk2(v, v); }
// 16-byte shared-memory load as two register pairs: (x, y), (z, w)
__device__ __forceinline__ void lds16x2(uint32_t addr, f32x2& xy, f32x2& zw) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xy), "=l"(zw) : "r"(addr));
}
// exp_fma_neg (hjk_math.cuh) of two arguments at once; takes a = -e.  Same operations in the same order per lane.
__device__ __forceinline__ f32x2 exp_fma_neg2(f32x2 a) {
  const f32x2 t = fma2(a, splat2(1.44269504088896341f), splat2(12582912.0f));
  const f32x2 n = add2(t, splat2(-12582912.0f));
  f32x2 r = fma2(n, splat2(-0.693145751953125f), a);
  r = fma2(n, splat2(-1.42860682030941723212e-6f), r);
  f32x2 p = fma2(splat2(1.384070492e-03f), r, splat2(8.368702605e-03f));
  p = fma2(p, r, splat2(4.166791961e-02f));
  p = fma2(p, r, splat2(1.666652113e-01f));
  p = fma2(p, r, splat2(4.999999404e-01f));
  p = fma2(p, r, splat2(1.0f));
  const f32x2 y = fma2(p, r, splat2(1.0f));
  float t0, t1, a0, a1, v0, v1;
  upk2(t, t0, t1);
  upk2(a, a0, a1);
  const f32x2 v = mul2(y, pk2(__uint_as_float((__float_as_uint(t0) << 23) + 0x3F800000u),
                              __uint_as_float((__float_as_uint(t1) << 23) + 0x3F800000u)));
  upk2(v, v0, v1);
  return pk2(a0 < -87.3365402f ? 0.0f : v0, a1 < -87.3365402f ? 0.0f : v1);  // e > cutoff  <=>  -e < -cutoff
}

// One visit of a warp's 32 texels by ONE block (k_recon decides which: the block is warp-uniform, so its tap list
// is read through uniform addresses and the loop below does not diverge).  The taps of the list are taken two at a
// time on packed fp32 pairs.
//   BOUNDS = false: every texel of the warp lies at least R inside the block: all taps of the list count.
//   BOUNDS = true : a tap counts for a lane when its sample lies inside the block (reconstruction.glsl:41-46); the
//     lane's position relative to the block is (lx, ly), any value.  A tap that does not count is evaluated on the
//     tile's dummy texel (finite values) with the weight -0: the products are -0, and x + (-0) = x for every x, so
//     the lane's sum is untouched without a branch.  A pair no lane needs is skipped (warp vote).
// No NaN test per tap: a NaN product, which reconstruction.glsl:55-58 drops, is added instead and leaves a NaN in
// the sum — the caller looks at the sum afterwards and, if it finds one, repeats the texel from the saved
// accumulator on the careful scalar path (exact whatever the NaN's origin).
//   `one` = PassDev::one.  ptxas (12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 — the scalar
//   forms with .rn are never contracted, the packed ones are, -fmad=false or not — which would round once where the
//   reference rounds twice.  The sum is therefore written fma(product, one, acc) with a 1.0 that arrives as a
//   kernel parameter: x * 1 + acc is exactly acc + x, two multiplies cannot be contracted, and what the compiler
//   cannot see it cannot simplify back into an add.
template <bool HAS_ALBEDO, bool BOUNDS>
struct ReconPairs {
  uint32_t a0, layer_bytes;    // shared address of the lane's texel in layer 0; bytes between layers
  uint32_t dummy;              // shared address of the dummy texel of layer 0 (BOUNDS)
  int lx128, ly128;            // lx - 128, ly - 128 (the packed tap offset carries dx + 128, dy + 128) (BOUNDS)
  uint32_t dimx, dimy;         // the block's dimension (BOUNDS)
  f32x2 neg_nc_xy, neg_ac_xy;  // minus the centre features
  float nc_z, ac_z, one;
  f32x2 acc_xy, acc_zw;

  // squared distance between the features at `layer_addr` and the centre's
  __device__ __forceinline__ float dist(uint32_t layer_addr, f32x2 neg_c_xy, float c_z) const {
    f32x2 n_xy, n_zw;
    lds16x2(layer_addr, n_xy, n_zw);
    float nz, nw, sx, sy;
    upk2(n_zw, nz, nw);
    const f32x2 d_xy = add2(n_xy, neg_c_xy);  // n - nc = n + (-nc), exactly
    const float dz = x::sub(nz, c_z);
    upk2(mul2(d_xy, d_xy), sx, sy);
    return x::add(x::add(sx, sy), x::mul(dz, dz));
  }
  __device__ __forceinline__ void add_products(f32x2 c_xy, f32x2 c_zw, float wt) {
    const f32x2 ww = splat2(wt), o = splat2(one);
    acc_xy = fma2(mul2(ww, c_xy), o, acc_xy), acc_zw = fma2(mul2(ww, c_zw), o, acc_zw);
  }
  __device__ __forceinline__ bool counts(uint32_t packed) const {
    return (uint32_t)(lx128 + (int)(packed & 0xFFu)) < dimx && (uint32_t)(ly128 + (int)((packed >> 8) & 0xFFu)) < dimy;
  }
  // taps A and B of the list: {spatial weight bits, packed offset}; TWO = false: A alone (the odd one at the end)
  template <bool TWO>
  __device__ __forceinline__ void pair(uint2 A, uint2 B) {
    uint32_t pa = a0 + (uint32_t)((int)A.y >> 16), pb = a0 + (uint32_t)((int)B.y >> 16);
    float wsa = __uint_as_float(A.x), wsb = __uint_as_float(B.x);
    if (BOUNDS) {
      const bool va = counts(A.y), vb = TWO && counts(B.y);
      if (!__any_sync(0xFFFFFFFFu, va || vb)) return;
      if (!va) pa = dummy, wsa = -0.0f;
      if (!vb) pb = dummy, wsb = -0.0f;
    }
    f32x2 ca_xy, ca_zw, cb_xy, cb_zw;
    lds16x2(pa, ca_xy, ca_zw);
    if (TWO) lds16x2(pb, cb_xy, cb_zw);
    const float dna = dist(pa + layer_bytes, neg_nc_xy, nc_z), dnb = TWO ? dist(pb + layer_bytes, neg_nc_xy, nc_z) : dna;
    f32x2 a;  // -e of both taps
    if (HAS_ALBEDO) {
      const float daa = dist(pa + 2u * layer_bytes, neg_ac_xy, ac_z);
      const float dab = TWO ? dist(pb + 2u * layer_bytes, neg_ac_xy, ac_z) : daa;
      a = mul2(pk2(x::add(x::mul(dna, 2.0f), daa), x::add(x::mul(dnb, 2.0f), dab)), splat2(-1.0f));
    } else {
      a = mul2(pk2(dna, dnb), splat2(-2.0f));  // -(2 x) = (-2) x, exactly
    }
    float wa, wb;
    upk2(mul2(pk2(wsa, wsb), exp_fma_neg2(a)), wa, wb);
    add_products(ca_xy, ca_zw, wa);
    if (TWO) add_products(cb_xy, cb_zw, wb);
  }
  // tl = the block's tap list (PassDev::taps), n = its length (tl[0]).  (Fetching the next pair's entries before
  // evaluating the current one measured slower: 0.235 vs 0.218 ms per 4K pass.)
  __device__ __forceinline__ void run(const uint32_t* tl, uint32_t n) {
    const uint2* tp = reinterpret_cast<const uint2*>(tl + 2);
    if (n == 0u) return;
    uint32_t k = 0;
#pragma unroll kReconUnroll
    for (; k + 1 < n; k += 2) pair<true>(__ldg(tp + k), __ldg(tp + k + 1));
    if (k < n) {
      const uint2 A = __ldg(tp + k);
      pair<false>(A, A);
    }
  }
  __device__ __forceinline__ void run(const uint32_t* tl) { run(tl, __ldg(tl)); }
  // centre features: the block's own texel, or the zero a robust out-of-bounds load returns for apron texels
  __device__ __forceinline__ void set_centre(bool inside) {
    f32x2 n_xy, n_zw;
    float nx, ny, nz, nw;
    lds16x2(a0 + layer_bytes, n_xy, n_zw);
    upk2(n_xy, nx, ny);
    upk2(n_zw, nz, nw);
    neg_nc_xy = inside ? pk2(-nx, -ny) : pk2(-0.f, -0.f), nc_z = inside ? nz : 0.f;
    neg_ac_xy = pk2(-0.f, -0.f), ac_z = 0.f;
    if (HAS_ALBEDO) {
      lds16x2(a0 + 2u * layer_bytes, n_xy, n_zw);
      upk2(n_xy, nx, ny);
      upk2(n_zw, nz, nw);
      neg_ac_xy = inside ? pk2(-nx, -ny) : pk2(-0.f, -0.f), ac_z = inside ? nz : 0.f;
    }
  }
  __device__ __forceinline__ void set_acc(f4 acc) { acc_xy = pk2(acc.x, acc.y), acc_zw = pk2(acc.z, acc.w); }
  __device__ __forceinline__ f4 get_acc() const {
    f4 r;
    upk2(acc_xy, r.x, r.y);
    upk2(acc_zw, r.z, r.w);
    return r;
  }
};
__device__ __forceinline__ bool any_nan(f4 v) { return v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w; }

// ---- TMA plumbing (cp.async.bulk.tensor + mbarrier), sm_90+ PTX written out by hand
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {  // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// the producer's wait (an item's time, not a memory round trip): sleeps between polls so that it leaves the issue
// slots to the warps that filter
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(HJK_RECON_PRODUCER_SLEEP);
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// box of the rank-3 tensor (floats of a row, rows, passes) at coordinates (c0, c1, c2) -> shared memory; texels
// outside the image arrive as zeros (the tensor map's out-of-bounds fill), negative coordinates included
__device__ __forceinline__ void tma_load_box3(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_addr(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
      : "memory");
}

// One launch reconstructs `n_passes` consecutive passes (layers [pass][pixel]) of the whole frame.
// PERSISTENT, WARP-SPECIALISED CTAs: CTA c filters the 32 x 8 tiles c, c + gridDim.x, ... and, for each, the passes one
// after the other, the accumulator texel staying in a register across them.  Warp 8 is the PRODUCER: for every
// (tile, pass) item it has the TMA engine bring the tile + halo of the two (three) layers into one of kReconStages
// shared-memory stages — one box copy per layer, texels outside the image zero-filled by the copy itself — and
// writes the item's description next to it (tile origin, the block the tile lies in: everything the eight CONSUMER
// warps would otherwise each derive with integer divisions and two dependent global loads).  full/empty mbarriers per
// stage: the producer runs up to kReconStages items ahead, across passes AND across tiles, and a consumer warp that
// is done with an item moves on without waiting for the CTA's slowest warp (a warp on a block edge does 2-3 times
// the work of an interior one).
// layer_stride = f4 elements between the layers of a stage: the box, one dummy texel, rounded up to 128 bytes; the
// accumulator texels of the tile travel with the first pass of every tile (a fourth box copy into the stage).
// RT = the filter radius when it is a compile-time constant (2, the reference's: src/main.rs:1284), else -1.
#ifndef HJK_RECON_STAGES
#define HJK_RECON_STAGES 3
#endif
constexpr uint32_t kReconStages = HJK_RECON_STAGES;
constexpr int kReconConsumers = kReconTileX * kReconTileY;  // threads; one more warp produces
constexpr int kReconThreads = kReconConsumers + 32;
HJK_HD uint32_t recon_layer_stride(int radius) {
  return ((uint32_t)(recon_smem_pitch(radius) * (kReconTileY + 2 * radius)) + 1u + 7u) & ~7u;
}
// Launches that fold several passes keep kReconSlots tiles of a CTA in flight at once, their items interleaved
// (T0 p0, T1 p0, T2 p0, T3 p0, T0 p1, ...): the passes of ONE tile are a dependency chain through the accumulator, and
// on a block edge the chain of the slow rectangle (3 times the work, the same warp for all the tile's passes) would
// gate the CTA — interleaved, four different warps carry the slow rectangles of four tiles.  The accumulator texels of
// a tile in flight live in shared memory (one 32 x 8 slot per tile, read and written by the one thread that owns the
// texel), where the TMA engine delivers them on the tile's first pass.  More slots than stages, so that a slot's
// previous tile is done by the time its next one is issued.
constexpr uint32_t kReconSlots = 4;
static_assert(kReconSlots > kReconStages, "a slot is reissued only after its previous tile left the pipeline");
struct alignas(16) ReconItem {
  int32_t tox, toy;      // image coordinate of the tile's first texel
  int32_t cb;            // the block the whole tile lies in (block grid a multiple of the tile), else -1
  uint32_t pass;         // pass index | kReconFirst | kReconLast
  int32_t box, boy;      // that block's origin,
  uint32_t bdx, bdy;     //   dimension
  int32_t btx, bty;      //   position in the block grid
  uint32_t n_taps;       //   and length of its tap list
  uint32_t slot, rot;    // accumulator slot of the tile (launches of several passes); the tile's number in the CTA's
                         // sequence (rotates the rectangles over the warps)
};
constexpr uint32_t kReconFirst = 0x40000000u, kReconLast = 0x80000000u;
// FEAT: also sum the texel's own first-hit features over the passes (feature_sum += (normal, depth),
// sample_count += layer 0's w) — the averaged feature buffers a multi-GPU frame reduces with the accumulator.
template <bool HAS_ALBEDO, int RT, bool FEAT>
__global__ void __launch_bounds__(kReconThreads, HJK_RECON_MIN_BLOCKS)
    k_recon(PassDev ps, uint32_t n_passes, uint32_t tiles_gx, uint32_t n_tiles, const __grid_constant__ CUtensorMap tm0,
            const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
            const __grid_constant__ CUtensorMap tm_acc, f4* __restrict__ accumulator, f4* __restrict__ feature_sum,
            float* __restrict__ sample_count) {
  extern __shared__ __align__(128) f4 smem[];
  __shared__ uint64_t full[kReconStages], empty[kReconStages];
  __shared__ ReconItem items[kReconStages];
  constexpr uint32_t NL = HAS_ALBEDO ? 3u : 2u;
  const int R = RT >= 0 ? RT : ps.radius;
  const int pitch = recon_smem_pitch(R), rows = kReconTileY + 2 * R;
  const uint32_t layer_stride = recon_layer_stride(R);
  const uint32_t stage_stride = NL * layer_stride + kReconConsumers;  // f4 elements: the layers, then the tile's accumulator texels
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (uint32_t st = 0; st < kReconStages; st++) mbar_init(&full[st], 1), mbar_init(&empty[st], kReconConsumers / 32);
    mbar_init_fence();
  }
  if ((uint32_t)tid < kReconStages * NL)  // the dummy texel behind every layer's box (ReconPairs): finite features, radiance 1
    smem[(size_t)((uint32_t)tid / NL) * stage_stride + (size_t)((uint32_t)tid % NL) * layer_stride + (size_t)(pitch * rows)] = F4(1.f, 1.f, 1.f, 1.f);
  __syncthreads();
  const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  const uint32_t n_items = my_tiles * n_passes;
  const size_t pass_tiles = (size_t)ps.tiles_x * ps.tiles_y;
  const bool aligned = RT >= 0 && ps.tile_w % kReconTileX == 0 && ps.tile_h % kReconTileY == 0;
  // several passes (and no feature sums, which stay in registers): kReconSlots tiles in flight, accumulators in the slots
  const uint32_t G = !FEAT && n_passes > 1 ? kReconSlots : 1u;
  f4* const slots = smem + (size_t)kReconStages * stage_stride;  // [kReconSlots][32 x 8]

  if (warp == kReconConsumers / 32) {  // ---------------------------------------------------------------- producer
    // 32 items at a time: every lane derives one item's description (two dependent global loads: all in flight at
    // once), then the lanes take turns issuing their item, so that an item costs the producer no memory round trip
    const uint32_t tx_bytes = (uint32_t)(pitch * rows) * 16u * NL;
    uint32_t st = 0, use = 0;
    for (uint32_t base = 0; base < n_items; base += 32u) {
      const uint32_t j = base + (uint32_t)lane;
      // item j: groups of G tiles, inside a group pass-major
      const uint32_t group = j / (G * n_passes), r = j - group * G * n_passes;
      const uint32_t g_n = my_tiles - group * G < G ? my_tiles - group * G : G;  // (j >= n_items: never used)
      const uint32_t pass = g_n ? r / g_n : 0u, g = g_n ? r % g_n : 0u;
      const uint32_t seq = group * G + g;  // the tile's number in this CTA's sequence
      const uint32_t tile = blockIdx.x + seq * gridDim.x;
      // tiles are numbered column by column: the tiles c, c + gridDim.x, ... of a CTA then fall on every position
      // inside a block (row-major, with 4 tiles per block row and a grid that is a multiple of 4, half of the CTAs
      // would get nothing but tiles on a block's left or right edge: 20 % more work than the others)
      const uint32_t tiles_gy = n_tiles / tiles_gx;
      const uint32_t tbx = tile / tiles_gy, tby = tile % tiles_gy;
      const bool last = pass + 1 == n_passes;
      ReconItem it;
      it.tox = (int32_t)(tbx * kReconTileX), it.toy = (int32_t)(tby * kReconTileY);
      it.pass = pass | (pass == 0 ? kReconFirst : 0u) | (last ? kReconLast : 0u);
      it.cb = -1, it.box = it.boy = 0, it.bdx = it.bdy = 0, it.btx = it.bty = 0, it.n_taps = 0;
      it.slot = g, it.rot = seq;
      if (aligned && j < n_items) {
        it.btx = (int32_t)(tbx * kReconTileX / ps.tile_w), it.bty = (int32_t)(tby * kReconTileY / ps.tile_h);
        it.cb = ps.tile_block[(size_t)pass * pass_tiles + (size_t)it.bty * ps.tiles_x + it.btx];
        if (it.cb >= 0) {
          const HjkImageBlock& blk = ps.blocks[it.cb];
          it.box = (int32_t)blk.origin[0], it.boy = (int32_t)blk.origin[1];
          it.bdx = blk.dimension[0], it.bdy = blk.dimension[1];
          it.n_taps = ps.taps[(size_t)it.cb * recon_tap_stride(R)];
        }
      }
      const uint32_t n_here = n_items - base < 32u ? n_items - base : 32u;
      for (uint32_t l = 0; l < n_here; l++) {
        if ((uint32_t)lane == l) {
          if (use) mbar_wait_relaxed(&empty[st], (use - 1u) & 1u);  // every consumer warp is done with the stage's previous item
          const int x0 = it.tox - R, y0 = it.toy - R;
          f4* dst = smem + (size_t)st * stage_stride;
          tma_load_box3(dst, &tm0, &full[st], 4 * x0, y0, (int)pass);
          tma_load_box3(dst + layer_stride, &tm1, &full[st], 4 * x0, y0, (int)pass);
          if (HAS_ALBEDO) tma_load_box3(dst + 2 * layer_stride, &tm2, &full[st], 4 * x0, y0, (int)pass);
          if (pass == 0)  // the tile's accumulator texels: into its slot, or (one tile at a time) into the stage
            tma_load_box3(G > 1u ? slots + (size_t)g * kReconConsumers : dst + NL * layer_stride, &tm_acc, &full[st], 4 * it.tox, it.toy, 0);
          items[st] = it;
          // the stage's one arrival (release: the description is visible with it)
          mbar_expect_tx(&full[st], tx_bytes + (pass == 0 ? (uint32_t)kReconConsumers * 16u : 0u));
        }
        __syncwarp();
        if (++st == kReconStages) st = 0, use++;
      }
    }
    return;
  }

  // ---------------------------------------------------------------------------------------------------- consumers
  // a warp covers 4 x 8 texels, not 32 x 1: the fewest warps touch a block edge that way
  // Which rectangle a warp takes rotates from tile to tile: the rectangle on a block edge costs 2-3 times the
  // others and sits at the same place in every tile of a block column, so a fixed assignment would leave seven warps
  // waiting for the eighth; rotated, the pipeline's stages absorb it.
  static_assert(kReconTileY == 8 && kReconTileX == 32, "8 warp rectangles of 4 x 8 texels");
  uint32_t wx = 0, tx_in = 0;
  const uint32_t wy = 0, ty_in = (uint32_t)lane >> 2;
  int cidx = 0;
  f4 acc = F4(0.f, 0.f, 0.f, 0.f), feat = acc;
  float cnt = 0.f;
  uint32_t gx = 0, gy = 0, st = 0, use = 0;
  bool in_image = false;
  for (uint32_t j = 0; j < n_items; j++) {
    mbar_wait(&full[st], use & 1u);
    const ReconItem& it = items[st];  // (left in shared memory: fields are read where they are used — all of them before
                                      // this warp's arrival on empty[st] below, after which the producer may overwrite it)
    const uint32_t item_flags = it.pass;
    const uint32_t pass = item_flags & 0x3FFFFFFFu;
    wx = 4u * (((uint32_t)warp + it.rot) & 7u), tx_in = wx + ((uint32_t)lane & 3u);
    cidx = ((int)ty_in + R) * pitch + ((int)tx_in + R);
    gx = (uint32_t)it.tox + tx_in, gy = (uint32_t)it.toy + ty_in;
    in_image = gx < ps.width && gy < ps.height;
    f4* const my_slot = slots + (size_t)it.slot * kReconConsumers + ty_in * kReconTileX + tx_in;
    if (G > 1u) {
      acc = *my_slot;  // delivered by the TMA engine on the tile's first pass, else this thread's own sum so far
    } else if (item_flags & kReconFirst) {
      acc = smem[(size_t)st * stage_stride + NL * layer_stride + ty_in * kReconTileX + tx_in];  // came with the stage
      if (FEAT && in_image) feat = feature_sum[(size_t)gy * ps.width + gx], cnt = sample_count[(size_t)gy * ps.width + gx];
    }
    const f4* s0 = smem + (size_t)st * stage_stride;
    const f4* s1 = s0 + layer_stride;
    const f4* s2 = s1 + layer_stride;
    const int x0 = it.tox - R, y0 = it.toy - R;
    const int32_t* tile_block = ps.tile_block + (size_t)pass * pass_tiles;
    if (FEAT && in_image) {  // a texel without a sample in this pass holds zeros in both layers (k_raygen)
      const f4 f = s1[cidx];
      feat = F4(x::add(feat.x, f.x), x::add(feat.y, f.y), x::add(feat.z, f.z), x::add(feat.w, f.w));
      cnt = x::add(cnt, s0[cidx].w);
    }
    // it.cb = the block the tile lies in (block grid a multiple of the tile: the reference's 128 x 128 blocks).  All 32
    // lanes of a warp stay together: either every texel of the warp sits at least R inside that block (one visit,
    // every tap counts), or the warp visits, in block-list order, every block whose apron reaches its 4 x 8
    // rectangle, each lane counting the taps whose samples lie inside the visited block.
    if (RT >= 0 && it.cb >= 0) {
      const uint32_t a0 = smem_addr(s0) + (uint32_t)cidx * 16u, layer_bytes = layer_stride * 16u;
      const uint32_t ox = (uint32_t)(it.tox - it.box) + wx, oy = (uint32_t)(it.toy - it.boy) + wy;  // the rectangle in the block
      const bool interior = ox >= (uint32_t)R && oy >= (uint32_t)R && ox + 3u + (uint32_t)R < it.bdx && oy + 7u + (uint32_t)R < it.bdy;
      f4 r;
      if (interior) {
        ReconPairs<HAS_ALBEDO, false> v;
        v.a0 = a0, v.layer_bytes = layer_bytes, v.one = ps.one;
        v.set_centre(true);
        v.set_acc(acc);
        v.run(ps.taps + (size_t)it.cb * recon_tap_stride(R), it.n_taps);
        r = v.get_acc();
      } else {
        ReconPairs<HAS_ALBEDO, true> v;
        v.a0 = a0, v.layer_bytes = layer_bytes, v.one = ps.one;
        v.dummy = smem_addr(s0) + (uint32_t)(pitch * rows) * 16u;
        v.set_acc(acc);
        // blocks that can hold a sample within R of the rectangle: the own one and its neighbours in the block grid
        // (blocks sit on the grid: the own block's origin is (btx * tile_w, bty * tile_h))
        int tx0 = it.btx - (ox < (uint32_t)R ? 1 : 0), ty0 = it.bty - (oy < (uint32_t)R ? 1 : 0);
        int tx1 = it.btx + (ox + 3u + (uint32_t)R >= ps.tile_w ? 1 : 0), ty1 = it.bty + (oy + 7u + (uint32_t)R >= ps.tile_h ? 1 : 0);
        if (tx0 < 0) tx0 = 0;
        if (ty0 < 0) ty0 = 0;
        if (tx1 >= (int)ps.tiles_x) tx1 = (int)ps.tiles_x - 1;
        if (ty1 >= (int)ps.tiles_y) ty1 = (int)ps.tiles_y - 1;
        for (int by = ty0; by <= ty1; by++) {
          for (int bx = tx0; bx <= tx1; bx++) {
            const int32_t b = tile_block[by * (int)ps.tiles_x + bx];
            if (b < 0) continue;
            const HjkImageBlock& blk = ps.blocks[b];
            const int lx = (int)gx - (int)blk.origin[0], ly = (int)gy - (int)blk.origin[1];
            v.dimx = blk.dimension[0], v.dimy = blk.dimension[1];
            // a lane outside the image counts no tap
            v.lx128 = in_image ? lx - 128 : (int)0x80000000, v.ly128 = ly - 128;
            v.set_centre((uint32_t)lx < v.dimx && (uint32_t)ly < v.dimy);
            v.run(ps.taps + (size_t)b * recon_tap_stride(R));
          }
        }
        r = v.get_acc();
      }
      if (in_image) {
        if (any_nan(r)) {  // a NaN was added (or the accumulator held one): the careful path decides
          const SmemLayers L{s0, s1, s2, x0, y0, pitch};
          PassDev pp = ps;
          pp.tile_block = tile_block;
          r = reconstruct_pixel<HAS_ALBEDO>(pp, L, gx, gy, acc);
        }
        acc = r;
      }
    } else if (in_image) {
      const int32_t b = tile_block[(gy / ps.tile_h) * ps.tiles_x + gx / ps.tile_w];
      bool interior = false;
      if (b >= 0) {
        const HjkImageBlock& blk = ps.blocks[b];
        const uint32_t lx = gx - blk.origin[0], ly = gy - blk.origin[1];  // >= 0: block origins sit on the tile grid
        interior = lx >= (uint32_t)R && ly >= (uint32_t)R && lx + (uint32_t)R < blk.dimension[0] &&
                   ly + (uint32_t)R < blk.dimension[1];
      }
      if (interior) {
        acc = reconstruct_interior<HAS_ALBEDO>(ps.taps + (size_t)b * recon_tap_stride(R),
                                               smem_addr(s0) + (uint32_t)cidx * 16u, layer_stride * 16u, acc);
      } else {
        const SmemLayers L{s0, s1, s2, x0, y0, pitch};
        PassDev pp = ps;
        pp.tile_block = tile_block;
        acc = reconstruct_pixel<HAS_ALBEDO>(pp, L, gx, gy, acc);
      }
    }
    if (G > 1u && !(item_flags & kReconLast)) {
      *my_slot = acc;
      // the slot's next tile arrives through the async proxy: order this generic-proxy write before it
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done with the stage
    if ((item_flags & kReconLast) && in_image) {
      accumulator[(size_t)gy * ps.width + gx] = acc;
      if (FEAT) feature_sum[(size_t)gy * ps.width + gx] = feat, sample_count[(size_t)gy * ps.width + gx] = cnt;
    }
    if (++st == kReconStages) st = 0, use++;
  }
}

// averaged features: (sum normal / n, sum depth / n); texels never sampled read as zeros
__global__ void k_normalise_features(const f4* sum, const float* cnt, f4* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const f4 a = sum[i];
    const float c = cnt[i];
    out[i] = c > 0.f ? F4(x::div(a.x, c), x::div(a.y, c), x::div(a.z, c), x::div(a.w, c)) : F4(0.f, 0.f, 0.f, 0.f);
  }
}

// save_image's divide (reference src/main.rs:1399): (r/w, g/w, b/w, w)
__global__ void k_normalise(const f4* acc, f4* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const f4 a = acc[i];
    out[i] = F4(x::div(a.x, a.w), x::div(a.y, a.w), x::div(a.z, a.w), a.w);
  }
}

// paths / extension rays / shadow rays of one wave, added to the call's totals
__global__ void k_wave_totals(const uint32_t* counters, uint32_t n_rows, unsigned long long* totals) {
  unsigned long long ext = 0, sh = 0;
  for (uint32_t r = threadIdx.x; r < n_rows; r += 32) {
    ext += counters[(size_t)r * CTR_STRIDE + CTR_EXT];
    sh += counters[(size_t)r * CTR_STRIDE + CTR_SHADOW];
  }
  for (int o = 16; o > 0; o >>= 1) {
    ext += __shfl_down_sync(0xFFFFFFFFu, ext, o);
    sh += __shfl_down_sync(0xFFFFFFFFu, sh, o);
  }
  if (threadIdx.x == 0) {
    totals[0] += counters[CTR_EXT];
    totals[1] += ext;
    totals[2] += sh;
  }
}

__global__ void k_iota(uint32_t* q, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) q[i] = i;
}

}  // namespace hjk
