// Wavefront path-tracing kernels for sm_100a — the B200 replacement of the reference's single
// render.glsl megakernel (reference shader/render.glsl:149-175) and of its per-block
// reconstruction dispatch (shader/reconstruction.glsl:22-66).
//
//   k_raygen    one thread per path slot: RNG seed, camera ray, layer initialisation
//   k_traverse  persistent warps pull rays from a queue and walk the 8-wide BVH
//               (closest hit = "extend", any hit = "shadow"), short stack in shared memory
//   k_bin       ballot / prefix-sum compaction of the hits into one queue per material tag
//   k_shade     material-sorted: emission, next-event estimation, BSDF sample, roulette;
//               appends the next extension ray and the shadow ray by block-level compaction
//   k_recon     shared-memory tiled bilateral splat of one or more passes into the accumulator
//
// No host synchronisation inside a wave: every kernel reads its element count from device
// counters written by the previous stage.
#pragma once
#include <cuda_runtime.h>

#include "recon.cuh"
#include "shade.cuh"
#include "traverse.cuh"

namespace hjk {

// per-bounce device counters (uint32 each)
enum : uint32_t {
  CTR_EXT = 0,      // extension rays entering this bounce
  CTR_TAG0 = 1,     // .. CTR_TAG0+4: hits per material tag after extend
  CTR_SHADOW = 6,   // shadow rays emitted by this bounce
  CTR_EXT_CURSOR = 7,
  CTR_SH_CURSOR = 8,
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
  // per-slot path state
  f4* ray_o;                     // origin.xyz, tMin
  f4* ray_d;                     // direction.xyz, tMax
  f4* hit;                       // shape id bits, t, u, v
  f4* thr_rng;                   // throughput.rgb, rng state bits
  f4* extinction;                // currentExtinction.rgb (only when the scene can set it)
  // per-pass intermediate layers of the wave, [n_wave_passes][n_pixels]
  f4* layer0;                    // (radiance, 1)      render.glsl:172
  f4* layer1;                    // (normal, depth)    render.glsl:173
  // queues
  uint32_t* ext_q[2];            // slot | wasDiscrete << 31
  uint32_t* tag_q;               // [5][n_slots]
  f4* sh_o;                      // shadow rays, dense
  f4* sh_d;
  f4* sh_c;                      // contribution.rgb, slot bits
  uint32_t* counters;            // [bounce][CTR_STRIDE]
  f4* accumulator;               // full frame (sum w*rgb, sum w)
  uint32_t max_bounces, rr_start;
  int32_t recon_radius;
  float eps;
  uint32_t has_extinction;
};

constexpr int kTravThreads = 128;
constexpr int kSmStack = 8;       // stack entries kept in shared memory per thread
constexpr int kLocalStack = 24;   // overflow entries in local memory
constexpr int kMaxStack = kSmStack + kLocalStack;
constexpr int kFetchThreshold = 20;  // refill a warp when fewer lanes than this are busy
constexpr int kTileThreads = 256;

// ---------------------------------------------------------------- block-level compaction
// Every thread of the block calls this (flag may be false).  Returns the global position
// of the thread's element in the queue whose length lives at *counter.  One atomic per block.
template <int NQ>
struct BlockAppend {
  uint32_t warp_total[NQ][kTileThreads / 32];
  uint32_t base[NQ];
};
template <int NQ>
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
  __syncthreads();  // sm is reused by the next tile
}

// ---------------------------------------------------------------- raygen
// render.glsl:149-162 for every pixel of every block of the wave's passes.
__global__ void __launch_bounds__(kTileThreads) k_raygen(WaveDev w) {
  __shared__ BlockAppend<1> sm;
  const uint32_t n_tiles = (w.n_slots + kTileThreads - 1) / kTileThreads;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t slot = tile * kTileThreads + threadIdx.x;
    bool valid = false;
    if (slot < w.n_slots) {
      const uint32_t wp = slot / w.n_pixels, pix = slot - wp * w.n_pixels;
      const uint32_t gy = pix / w.width, gx = pix - gy * w.width;
      const int32_t b = w.tile_block[(size_t)wp * w.tiles_x * w.tiles_y + (gy / w.tile_h) * w.tiles_x + gx / w.tile_w];
      if (b >= 0) {
        const HjkImageBlock blk = w.blocks[b];
        const uint32_t lx = gx - blk.origin[0], ly = gy - blk.origin[1];
        if (lx < blk.dimension[0] && ly < blk.dimension[1]) {
          valid = true;
          const uint32_t rng = seed_rng(blk.seed + lx + ly * blk.dimension[0]);  // render.glsl:156
          f4 o, d;
          camera_ray(w.scene.camera, x::add((float)gx, blk.sample_offset[0]),
                     x::add((float)gy, blk.sample_offset[1]), (float)blk.original_dimension[0],
                     (float)blk.original_dimension[1], w.eps, o, d);
          w.ray_o[slot] = o;
          w.ray_d[slot] = d;
          w.thr_rng[slot] = F4(1.f, 1.f, 1.f, __uint_as_float(rng));
          if (w.has_extinction) w.extinction[slot] = F4(0.f, 0.f, 0.f, 0.f);
        }
      }
      w.layer0[slot] = F4(0.f, 0.f, 0.f, valid ? 1.f : 0.f);
      w.layer1[slot] = F4(0.f, 0.f, 0.f, 0.f);
    }
    const bool flag[1] = {valid};
    uint32_t* const ctr[1] = {w.counters + CTR_EXT};
    uint32_t pos[1];
    block_append<1>(sm, flag, ctr, pos);
    if (valid) w.ext_q[0][pos[0]] = slot | 0x80000000u;  // wasDiscrete = true (render.glsl:91)
  }
}

// ---------------------------------------------------------------- traversal
struct DevStack {
  uint2* sm;  // this thread's column of the shared-memory stack (stride kTravThreads)
  uint2 local[kLocalStack];
  int n;
  __device__ __forceinline__ void push(uint32_t a, uint32_t b) {
    if (n < kSmStack) {
      sm[n * kTravThreads] = make_uint2(a, b);
    } else if (n < kMaxStack) {
      local[n - kSmStack] = make_uint2(a, b);
    }
    n++;
  }
  __device__ __forceinline__ void pop(uint32_t& a, uint32_t& b) {
    n--;
    const uint2 v = n < kSmStack ? sm[n * kTravThreads] : local[(n < kMaxStack ? n : kMaxStack - 1) - kSmStack];
    a = v.x, b = v.y;
  }
  __device__ __forceinline__ bool empty() const { return n == 0; }
};

// Source of rays / sink of results for the two traversal flavours.
struct ExtendIO {  // closest hit over the extension queue of `bounce`
  const WaveDev& w;
  const uint32_t* queue;
  __device__ __forceinline__ void load(uint32_t i, TravState& s) const {
    const uint32_t slot = queue[i] & 0x7FFFFFFFu;
    s.slot = slot;
    trav_init(s, w.scene, w.ray_o[slot], w.ray_d[slot]);
  }
  __device__ __forceinline__ void store(const TravState& s) const {
    w.hit[s.slot] = F4(__int_as_float(s.hit_id), s.hit_t, s.hit_u, s.hit_v);
  }
};
struct ShadowIO {  // any hit over the dense shadow queue; visible -> add the contribution
  const WaveDev& w;
  __device__ __forceinline__ void load(uint32_t i, TravState& s) const {
    s.slot = i;
    trav_init(s, w.scene, w.sh_o[i], w.sh_d[i]);
  }
  __device__ __forceinline__ void store(const TravState& s) const {
    if (s.hit_id >= 0) return;  // occluded (render.glsl:122)
    const f4 c = w.sh_c[s.slot];
    const uint32_t slot = __float_as_uint(c.w);
    f4 r = w.layer0[slot];      // total += throughput * evalBSDF * importance (render.glsl:123)
    r.x = x::add(r.x, c.x), r.y = x::add(r.y, c.y), r.z = x::add(r.z, c.z);
    w.layer0[slot] = r;
  }
};
// Standalone ray batch (hjk_trace_first_hit): rays and results are indexed by ray number.
struct BatchIO {
  const SceneDev& sc;
  const f4* ray_o;
  const f4* ray_d;
  f4* hit;
  __device__ __forceinline__ void load(uint32_t i, TravState& s) const {
    s.slot = i;
    trav_init(s, sc, ray_o[i], ray_d[i]);
  }
  __device__ __forceinline__ void store(const TravState& s) const {
    hit[s.slot] = F4(__int_as_float(s.hit_id), s.hit_t, s.hit_u, s.hit_v);
  }
};

// Persistent warps: each warp keeps its 32 lanes supplied with rays from the queue; a lane
// whose ray finishes is refilled as soon as fewer than kFetchThreshold lanes are busy.
template <bool ANY_HIT, class IO>
__device__ __forceinline__ void traverse_queue(const SceneDev& sc, const IO& io, uint32_t n,
                                               uint32_t* cursor, float eps) {
  __shared__ uint2 sm_stack[kSmStack * kTravThreads];
  const uint32_t lane = threadIdx.x & 31u;
  DevStack st;
  st.sm = sm_stack + threadIdx.x;
  st.n = 0;
  TravState s;
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
            io.load(i, s);
            st.n = 0;
            active = true;
          }
        }
        if (base + __popc(need) >= n) exhausted = true;
      }
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) break;
    if (active) {
      const bool done = trav_run<ANY_HIT>(sc, s, st, eps, [&]() {
        return !exhausted && __popc(__activemask()) < kFetchThreshold;
      });
      if (done) {
        io.store(s);
        active = false;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kTravThreads) k_extend(WaveDev w, uint32_t bounce) {
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const ExtendIO io{w, w.ext_q[bounce & 1u]};
  traverse_queue<false>(w.scene, io, ctr[CTR_EXT], ctr + CTR_EXT_CURSOR, w.eps);
}
__global__ void __launch_bounds__(kTravThreads) k_shadow(WaveDev w, uint32_t bounce) {
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const ShadowIO io{w};
  traverse_queue<true>(w.scene, io, ctr[CTR_SHADOW], ctr + CTR_SH_CURSOR, w.eps);
}
// counters: [0] = cursor
template <bool ANY_HIT>
__global__ void __launch_bounds__(kTravThreads) k_trace_batch(SceneDev sc, const f4* ray_o, const f4* ray_d,
                                                              f4* hit, uint32_t n, uint32_t* cursor, float eps) {
  const BatchIO io{sc, ray_o, ray_d, hit};
  traverse_queue<ANY_HIT>(sc, io, n, cursor, eps);
}

// ---------------------------------------------------------------- material binning
// Splits the hits of one bounce into one queue per material tag (material-sorted shading);
// misses leave the pipeline here (render.glsl:94-96).
__global__ void __launch_bounds__(kTileThreads) k_bin(WaveDev w, uint32_t bounce) {
  __shared__ BlockAppend<5> sm;
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  const uint32_t n = ctr[CTR_EXT];
  const uint32_t* q = w.ext_q[bounce & 1u];
  const uint32_t n_tiles = (n + kTileThreads - 1) / kTileThreads;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t i = tile * kTileThreads + threadIdx.x;
    uint32_t entry = 0, tag = 0xFFFFFFFFu;
    if (i < n) {
      entry = q[i];
      const int id = __float_as_int(w.hit[entry & 0x7FFFFFFFu].x);
      if (id >= 0) tag = ld4(w.scene.materials + id) >> HJK_MATERIAL_TAG_SHIFT;
    }
    const bool flag[5] = {tag == 0u, tag == 1u, tag == 2u, tag == 3u, tag == 4u};
    uint32_t* const c[5] = {ctr + CTR_TAG0, ctr + CTR_TAG0 + 1, ctr + CTR_TAG0 + 2, ctr + CTR_TAG0 + 3,
                            ctr + CTR_TAG0 + 4};
    uint32_t pos[5];
    block_append<5>(sm, flag, c, pos);
    if (tag < 5u) w.tag_q[(size_t)tag * w.n_slots + pos[tag]] = entry;
  }
}

// ---------------------------------------------------------------- shade
__global__ void __launch_bounds__(kTileThreads) k_shade(WaveDev w, uint32_t bounce) {
  __shared__ BlockAppend<2> sm;
  uint32_t* ctr = w.counters + (size_t)bounce * CTR_STRIDE;
  uint32_t first[6];
  first[0] = 0;
#pragma unroll
  for (int t = 0; t < 5; t++) first[t + 1] = first[t] + ctr[CTR_TAG0 + t];
  const uint32_t n = first[5];
  const uint32_t n_tiles = (n + kTileThreads - 1) / kTileThreads;
  uint32_t* next_q = w.ext_q[(bounce + 1u) & 1u];
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t i = tile * kTileThreads + threadIdx.x;
    bool want_next = false, want_shadow = false;
    VertexOut out;
    uint32_t slot = 0;
    if (i < n) {
      uint32_t tag = 0;
#pragma unroll
      for (int t = 1; t < 5; t++) tag += (i >= first[t]) ? 1u : 0u;
      const uint32_t entry = w.tag_q[(size_t)tag * w.n_slots + (i - first[tag])];
      slot = entry & 0x7FFFFFFFu;
      VertexIn in;
      in.ray_o = w.ray_o[slot];
      in.ray_d = w.ray_d[slot];
      const f4 h = w.hit[slot];
      in.hit_id = __float_as_int(h.x), in.hit_t = h.y, in.hit_u = h.z, in.hit_v = h.w;
      const f4 tr = w.thr_rng[slot];
      in.throughput = xyz(tr);
      in.rng = __float_as_uint(tr.w);
      in.extinction = w.has_extinction ? xyz(w.extinction[slot]) : V3(0.f);
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
      if (want_next) {
        w.ray_o[slot] = out.next_o;
        w.ray_d[slot] = out.next_d;
        w.thr_rng[slot] = F4(out.throughput.x, out.throughput.y, out.throughput.z, __uint_as_float(out.rng));
        if (w.has_extinction) w.extinction[slot] = F4(out.extinction.x, out.extinction.y, out.extinction.z, 0.f);
      }
    }
    const bool flag[2] = {want_next, want_shadow};
    uint32_t* const c[2] = {ctr + CTR_STRIDE + CTR_EXT, ctr + CTR_SHADOW};
    uint32_t pos[2];
    block_append<2>(sm, flag, c, pos);
    if (want_next) next_q[pos[0]] = slot | (out.was_discrete ? 0x80000000u : 0u);
    if (want_shadow) {
      w.sh_o[pos[1]] = out.sh_o;
      w.sh_d[pos[1]] = out.sh_d;
      w.sh_c[pos[1]] = F4(out.contribution.x, out.contribution.y, out.contribution.z, __uint_as_float(slot));
    }
  }
}

// ---------------------------------------------------------------- reconstruction
__global__ void k_recon_weights(const HjkImageBlock* blocks, uint32_t n_blocks, int radius, float stddev,
                                float* weights) {
  const int taps = 2 * radius + 1;
  const uint32_t total = n_blocks * (uint32_t)(taps * taps);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t b = i / (uint32_t)(taps * taps);
    const int k = (int)(i - b * (uint32_t)(taps * taps));
    const int dx = k / taps - radius, dy = k % taps - radius;
    weights[i] = recon_spatial_weight(dx, dy, radius, stddev, blocks[b].sample_offset[0],
                                      blocks[b].sample_offset[1]);
  }
}

constexpr int kReconTileX = 32, kReconTileY = 16;

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

// One launch reconstructs `n_passes` consecutive passes (layers [pass][pixel]): the
// accumulator texel stays in a register across passes, each pass' tile is staged once.
// pass_tile_block advances by tiles_x*tiles_y per pass.  albedo may be null.
template <bool HAS_ALBEDO>
__global__ void __launch_bounds__(kReconTileX* kReconTileY)
    k_recon(PassDev ps, uint32_t n_passes, const f4* __restrict__ layer0, const f4* __restrict__ layer1,
            const f4* __restrict__ layer2, f4* __restrict__ accumulator) {
  extern __shared__ f4 smem[];
  const int R = ps.radius;
  const int pitch = kReconTileX + 2 * R, rows = kReconTileY + 2 * R;
  f4* s0 = smem;
  f4* s1 = s0 + pitch * rows;
  f4* s2 = s1 + pitch * rows;
  const int x0 = (int)blockIdx.x * kReconTileX - R, y0 = (int)blockIdx.y * kReconTileY - R;
  const uint32_t gx = blockIdx.x * kReconTileX + threadIdx.x, gy = blockIdx.y * kReconTileY + threadIdx.y;
  const bool in_image = gx < ps.width && gy < ps.height;
  const size_t n_pixels = (size_t)ps.width * ps.height;
  f4 acc = F4(0.f, 0.f, 0.f, 0.f);
  if (in_image) acc = accumulator[(size_t)gy * ps.width + gx];
  const int tid = threadIdx.y * kReconTileX + threadIdx.x;
  for (uint32_t p = 0; p < n_passes; p++) {
    const f4* l0 = layer0 + p * n_pixels;
    const f4* l1 = layer1 + p * n_pixels;
    const f4* l2 = HAS_ALBEDO ? layer2 + p * n_pixels : nullptr;
    for (int i = tid; i < pitch * rows; i += kReconTileX * kReconTileY) {
      const int sy = i / pitch, sx = i - sy * pitch;
      const int px = x0 + sx, py = y0 + sy;
      f4 a = F4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
      if (px >= 0 && py >= 0 && px < (int)ps.width && py < (int)ps.height) {
        const size_t o = (size_t)py * ps.width + px;
        a = l0[o];
        b = l1[o];
        if (HAS_ALBEDO) c = l2[o];
      }
      s0[i] = a;
      s1[i] = b;
      if (HAS_ALBEDO) s2[i] = c;
    }
    __syncthreads();
    if (in_image) {
      const SmemLayers L{s0, s1, s2, x0, y0, pitch};
      acc = reconstruct_pixel<HAS_ALBEDO>(ps, L, gx, gy, acc);
    }
    __syncthreads();
    ps.tile_block += (size_t)ps.tiles_x * ps.tiles_y;
  }
  if (in_image) accumulator[(size_t)gy * ps.width + gx] = acc;
}

// save_image's divide (reference src/main.rs:1399): (r/w, g/w, b/w, w)
__global__ void k_normalise(const f4* acc, f4* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const f4 a = acc[i];
    out[i] = F4(x::div(a.x, a.w), x::div(a.y, a.w), x::div(a.z, a.w), a.w);
  }
}

__global__ void k_iota(uint32_t* q, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) q[i] = i;
}

}  // namespace hjk
