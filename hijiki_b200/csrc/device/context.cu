// C ABI of the device side (include/hijiki_b200.h): context, scene upload, the wavefront
// render loop that replaces Renderer::render (reference src/main.rs:1316-1355), readback,
// the parity hook and the standalone reconstruction entry.
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../host/pass_plan.h"
#include "../host/wide_bvh_host.h"
#include "bvh_build_gpu.cuh"
#include "kernels.cuh"

using namespace hjk;
using gpubvh::kPlocMaxRadius;
using gpubvh::kPlocRadius;

namespace {

// ------------------------------------------------------------------ NCCL via dlopen
// The library has no link-time dependency on NCCL; multi-GPU hosts that do not bring their
// own communicator (see hjk_accumulator_device_ptr) get one through these entry points.
struct Id128 {  // ncclUniqueId, passed by value
  char bytes[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string& err) {
  if (g_nccl.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    err = "cannot dlopen libnccl.so.2";
    return false;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(h, "ncclReduce");
  g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(h, "ncclBroadcast");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommInitAll || !g_nccl.AllReduce || !g_nccl.Reduce ||
      !g_nccl.Broadcast || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.CommDestroy) {
    err = "libnccl is missing expected symbols";
    return false;
  }
  g_nccl.handle = h;
  return true;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t ensure(size_t count) {  // grow-only
    if (count <= n) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

thread_local std::string g_create_error;

// scenes with more shapes than this are built on the GPU unless the option "bvh_builder" says otherwise
constexpr uint64_t kGpuBuilderMinShapes = 1000000;

}  // namespace

struct HjkContext {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;  // per-stage timing (profiling on)
  std::vector<int> ev_slot;
  size_t ev_used = 0;
  int n_sms = 0;
  int blocks_trav = 0, blocks_tile = 0;  // resident CTAs per SM of the persistent kernels (traverse / shade)
  int blocks_trav_v[2][2] = {{0, 0}, {0, 0}};  // ... per k_trace variant [sphere guard][exact ties]
  int blocks_trav_override = 0;               // option blocks_per_sm_traverse
  int blocks_light = 0;                  // ... of the light tile kernels (raygen, bin)
  std::string error;
  bool profiling = false;
  uint64_t wave_paths = 64u << 20;  // target camera paths per wave (13 GB of path state; tails amortise)
  float bvh_pad_rel = kDefaultBvhPadRel;
  int coop_trace = 1;    // 1 = k_trace_coop: pooled primitive tests (default mode only); 0 = per-lane k_trace
  int coop_batch_cost = -1;  // option "coop_batch_cost"; -1 = by scene (trace_tuning)
  int blocks_coop[3] = {0, 0, 0};  // k_trace_coop<GUARD = 0, 1, 2>
  int blocks_batch = 0;            // k_trace_batch (hjk_trace_first_hit)
  uint32_t stack_cap_coop = 8, stack_cap_lane = 16;  // traversal-stack entries per thread (see trace_launch_shape)
  bool lane_postpones = true;
  int bvh_builder = 0;   // 0 = host SAH builder (default), 1 = GPU builder, -1 = GPU builder beyond a million shapes
  int bvh_gpu_tree = 1;  // GPU builder's binary tree: 1 = PLOC (default), 0 = radix tree (LBVH)
  int bvh_ploc_radius = kPlocRadius;
  int bvh_gpu_collapse = 1;  // GPU builder's wide-node collapse: 1 = collapse-cost dynamic programme, 0 = greedy
  int bvh_validate = 0;  // download the tree after a GPU build and run the host structural check
  int bvh_broadcast = 1;  // several ranks: rank 0 builds the wide BVH, the others receive it over ncclBroadcast
  float bvh_build_ms = 0.f;
  int fetch_threshold = -1;  // option "fetch_threshold"; -1 = by scene (trace_tuning)
  uint32_t postpone_lanes = kPostponeLanes;

  // scene
  bool has_scene = false;
  SceneDev scene{};
  bool has_extinction = false;
  bool scene_mixes_materials = false;  // more than one of {diffuse-like, mirror, dielectric} among the shapes
  int shade_sort = -1;                 // option: -1 = by scene_mixes_materials, 0 / 1 = never / always sort the tiles
  bool bvh_all_guarded = false;  // every node of the wide BVH has a sphere below it (k_trace_coop<2>)
  WideBvh bvh_host_stats;  // nodes/prims cleared after upload; keeps depth etc.
  uint64_t n_nodes = 0, n_prims = 0;
  DevBuf<f4> d_nodes, d_prims, d_spheres, d_quads, d_vertices, d_emitters, d_diffuse, d_diffusecb,
      d_dielectric, d_emissive;
  DevBuf<uint32_t> d_triangles, d_materials;

  // frame
  uint32_t width = 0, height = 0;
  // [accumulator: 4 floats per texel | feature sums (normal, depth): 4 per texel | sample counts: 1 per texel]
  DevBuf<float> d_frame;
  DevBuf<float> d_sum;   // the same, summed over the ranks / devices of a multi-GPU frame (never aliases d_frame)
  DevBuf<f4> d_norm;     // staging of a readback
  bool feature_buffers = false;  // option "feature_buffers": k_recon also sums the first-hit features
  bool reduced_valid = false;    // d_sum holds the reduction of the frame as it stands
  int reduced_root = -2;         // ... delivered to this rank (-1 = to every rank)
  f4* acc() const { return (f4*)d_frame.p; }
  f4* feat() const { return (f4*)(d_frame.p + 4 * (size_t)width * height); }
  float* cnt() const { return d_frame.p + 8 * (size_t)width * height; }
  size_t frame_floats() const { return (size_t)width * height * (feature_buffers ? 9 : 4); }
  // asynchronous readback (hjk_readback_begin / hjk_readback_wait): the staged frame is copied on a stream of its own
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_staged = nullptr, ev_copied = nullptr;
  bool copy_pending = false;
  // single-process multi-GPU (hjk_create with n_devices > 1): members[0] is this context, the others own one
  // further device each; every member holds its communicator of ncclCommInitAll in `comm`
  std::vector<HjkContext*> members;
  HjkContext* group_parent = nullptr;

  // wave buffers
  DevBuf<f4> d_ray_o[2], d_ray_d[2], d_thr[2], d_ext[2], d_hit, d_layer0, d_layer1, d_sh_o, d_sh_d, d_sh_c;
  DevBuf<uint32_t> d_ext_q0, d_ext_q1, d_counters;
  DevBuf<unsigned long long> d_totals;  // paths, extension rays, shadow rays of the current call
  DevBuf<uint32_t> d_unresolved;        // exact-tie mode: rays whose cluster outgrew the window/list
  uint64_t unresolved_last = 0;         // ... of the last render / trace call
  DevBuf<int32_t> d_tile_block;
  DevBuf<HjkImageBlock> d_blocks;
  DevBuf<float> d_weights;
  DevBuf<uint32_t> d_taps;
  std::vector<uint32_t> h_counters;
  uint32_t last_wave_passes = 0;  // for hjk_read_intermediate
  bool have_features = false;

  // hjk_trace_first_hit batches (grow-only)
  DevBuf<f4> d_batch_o, d_batch_d, d_batch_h;
  DevBuf<uint32_t> d_batch_cur;
  std::vector<f4> h_batch_o, h_batch_d;

  // resident block lists
  struct Resident {
    std::vector<HjkImageBlock> host;
    HjkImageBlock* dev = nullptr;
  };
  std::map<uint64_t, Resident> resident;
  uint64_t next_handle = 1;

  // standalone denoise
  DevBuf<f4> d_dn0, d_dn1, d_dn2;
  std::vector<HjkImageBlock> dn_blocks;
  bool dn_has_albedo = false;

  // NCCL
  void* comm = nullptr;
  int rank = 0, n_ranks = 1;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    error = buf;
    return code;
  }
};

#define HJK_CUDA(ctx, call)                                                                    \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return (ctx)->fail(e_ == cudaErrorMemoryAllocation ? HJK_ERR_OUT_OF_MEMORY : HJK_ERR_CUDA, \
                         "%s failed: %s", #call, cudaGetErrorString(e_));                      \
  } while (0)

namespace {

template <class T>
int upload(HjkContext* c, DevBuf<T>& buf, const HjkArray& a, size_t elem_bytes) {
  const size_t bytes = (size_t)a.count * elem_bytes;
  const size_t n = (bytes + sizeof(T) - 1) / sizeof(T);
  HJK_CUDA(c, buf.ensure(std::max<size_t>(n, 1)));
  if (bytes) {
    if (!a.ptr) return c->fail(HJK_ERR_INVALID_ARGUMENT, "scene array has a count but no pointer");
    HJK_CUDA(c, cudaMemcpyAsync(buf.p, a.ptr, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  return HJK_OK;
}

int grid_for(const HjkContext* c, int per_sm) { return c->n_sms * std::max(per_sm, 1); }

// Per-stage device timing without host synchronisation: when profiling is on, every stage is
// bracketed by a pair of events taken from a pool; the pairs are resolved after the call's
// final synchronise.  Costs two event records per launch, no bubbles.
struct KernelTimer {
  HjkContext* c;
  bool on;
  KernelTimer(HjkContext* c_, HjkStats* st, int slot) : c(c_), on(c_->profiling && st != nullptr) {
    if (!on) return;
    if (c->ev_used + 2 > c->ev_pool.size()) {
      const size_t grow = c->ev_pool.size() + 256;
      while (c->ev_pool.size() < grow) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
      }
    }
    c->ev_slot.push_back(slot);
    cudaEventRecord(c->ev_pool[c->ev_used++], c->stream);
  }
  ~KernelTimer() {
    if (on) cudaEventRecord(c->ev_pool[c->ev_used++], c->stream);
  }
};
void resolve_timers(HjkContext* c, HjkStats* st) {  // after the stream has been synchronised
  if (st)
    for (size_t i = 0; i < c->ev_slot.size(); i++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]) == cudaSuccess)
        st->kernel_ms[c->ev_slot[i]] += ms;
    }
  c->ev_slot.clear();
  c->ev_used = 0;
}

int ensure_frame(HjkContext* c, uint32_t w, uint32_t h, bool zero) {
  const size_t n = (size_t)w * h;
  const bool fresh = c->width != w || c->height != h || !c->d_frame.p;
  HJK_CUDA(c, c->d_frame.ensure(9 * n));
  c->width = w;
  c->height = h;
  if (fresh || zero) HJK_CUDA(c, cudaMemsetAsync(c->d_frame.p, 0, 9 * n * sizeof(float), c->stream));
  c->reduced_valid = false;
  return HJK_OK;
}

// cuTensorMapEncodeTiled, fetched from the driver at run time (the library links the static runtime only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor map of one intermediate layer for k_recon's TMA staging: the layer is [n_passes][height][width] float4,
// described as a rank-3 fp32 tensor (4 * width floats per row, height rows, n_passes images) with a box of one
// tile + halo; out-of-bounds elements read as zero.
int recon_tensor_map(HjkContext* c, CUtensorMap* tm, const f4* layer, uint32_t width, uint32_t height, uint32_t n_passes,
                     int radius, bool halo = true) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return c->fail(HJK_ERR_CUDA, "the driver does not export cuTensorMapEncodeTiled");
  const cuuint64_t dims[3] = {4ull * width, height, n_passes};
  const cuuint64_t strides[2] = {16ull * width, 16ull * width * height};  // bytes, dimensions 1 and 2
  // halo: tile + apron of an intermediate layer; else the bare tile (the accumulator)
  const cuuint32_t box[3] = {4u * (uint32_t)(halo ? recon_smem_pitch(radius) : kReconTileX),
                             (uint32_t)(halo ? kReconTileY + 2 * radius : kReconTileY), 1u};
  const cuuint32_t elem_strides[3] = {1u, 1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<f4*>(layer), dims, strides, box, elem_strides,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return c->fail(HJK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return HJK_OK;
}

int launch_recon(HjkContext* c, const PassDev& ps_in, uint32_t n_passes, const f4* l0, const f4* l1, const f4* l2,
                 f4* acc, bool features) {
  if (ps_in.radius < 0 || ps_in.radius > 8) return c->fail(HJK_ERR_UNSUPPORTED, "recon_radius must be in [0, 8]");
  PassDev ps = ps_in;
  ps.one = 1.0f;
  const uint32_t layer_stride = recon_layer_stride(ps.radius);  // float4 elements, a multiple of 128 bytes
  const uint32_t stages = kReconStages;
  const size_t smem = (((size_t)layer_stride * (l2 ? 3 : 2) + kReconConsumers) * stages + (size_t)kReconSlots * kReconConsumers) * sizeof(f4);
  CUtensorMap tm0, tm1, tm2, tm_acc;
  int rc;
  if ((rc = recon_tensor_map(c, &tm0, l0, ps.width, ps.height, n_passes, ps.radius))) return rc;
  if ((rc = recon_tensor_map(c, &tm1, l1, ps.width, ps.height, n_passes, ps.radius))) return rc;
  if ((rc = recon_tensor_map(c, &tm_acc, acc, ps.width, ps.height, 1, ps.radius, false))) return rc;
  tm2 = tm1;
  if (l2 && (rc = recon_tensor_map(c, &tm2, l2, ps.width, ps.height, n_passes, ps.radius))) return rc;
  const dim3 block(kReconThreads);
  const uint32_t tiles_gx = (ps.width + kReconTileX - 1) / kReconTileX;
  const uint32_t n_tiles = tiles_gx * ((ps.height + kReconTileY - 1) / kReconTileY);
  // persistent CTAs, as many as are resident at once (registers and this radius' shared memory decide), each taking
  // every grid-th tile
#define HJK_RECON(A, RT, F)                                                                                         \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      HJK_CUDA(c, cudaFuncSetAttribute(k_recon<A, RT, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 0;                                                                                                 \
    HJK_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_recon<A, RT, F>, kReconThreads, smem));    \
    if (per_sm < 1) return c->fail(HJK_ERR_UNSUPPORTED, "k_recon does not fit an SM at recon_radius %d", ps.radius); \
    const dim3 grid(std::min<uint32_t>(n_tiles, (uint32_t)(c->n_sms * per_sm)));                                    \
    k_recon<A, RT, F><<<grid, block, smem, c->stream>>>(ps, n_passes, tiles_gx, n_tiles, tm0, tm1, tm2, tm_acc, acc, \
                                                        c->feat(), c->cnt());                                       \
  } while (0)
  if (l2) {  // standalone denoise with an albedo layer: no feature sums
    if (ps.radius == 2) HJK_RECON(true, 2, false); else HJK_RECON(true, -1, false);
  } else if (features) {
    if (ps.radius == 2) HJK_RECON(false, 2, true); else HJK_RECON(false, -1, true);
  } else {
    if (ps.radius == 2) HJK_RECON(false, 2, false); else HJK_RECON(false, -1, false);
  }
#undef HJK_RECON
  c->reduced_valid = false;
  HJK_CUDA(c, cudaGetLastError());
  return HJK_OK;
}

// The render loop over a block list that already lives on the device (d_blocks) and on the
// host (blocks).  Replaces Renderer::render, src/main.rs:1316-1355.
int render_blocks(HjkContext* c, const HjkImageBlock* blocks, const HjkImageBlock* d_blocks, uint64_t n_blocks,
                  const HjkParams* prm, HjkStats* stats) {
  if (!c->has_scene) return c->fail(HJK_ERR_NO_SCENE, "hjk_render before hjk_scene_upload");
  if (!prm) return c->fail(HJK_ERR_INVALID_ARGUMENT, "params is null");
  if (prm->max_bounces == 0 || prm->max_bounces > (1u << 16))
    return c->fail(HJK_ERR_UNSUPPORTED, "max_bounces must be in [1, 65536]");
  if (!(prm->recon_stddev > 0.f)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "recon_stddev must be positive");
  PassPlan plan;
  std::string err;
  if (!plan_passes(blocks, n_blocks, plan, err)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "%s", err.c_str());
  if (c->width && (c->width != plan.width || c->height != plan.height) && c->d_frame.p)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "blocks are for a %ux%u image but the frame is %ux%u", plan.width,
                   plan.height, c->width, c->height);
  int rc = ensure_frame(c, plan.width, plan.height, false);
  if (rc) return rc;

  const size_t n_pixels = (size_t)plan.width * plan.height;
  const size_t tiles = (size_t)plan.tiles_x * plan.tiles_y;
  const uint32_t n_passes = (uint32_t)plan.passes.size();
  uint32_t wave_passes = (uint32_t)std::max<uint64_t>(1, (c->wave_paths + n_pixels / 2) / n_pixels);
  wave_passes = std::min(wave_passes, n_passes);
  while ((size_t)n_pixels * wave_passes > 0x7FFFFFFFull && wave_passes > 1) wave_passes /= 2;
  if (n_pixels * wave_passes > 0x7FFFFFFFull) return c->fail(HJK_ERR_UNSUPPORTED, "frame too large");
  const int R = (int)prm->recon_radius, taps = 2 * R + 1;
  const bool do_recon = !(prm->flags & HJK_RENDER_NO_RECON);
  if (do_recon && R > 8) return c->fail(HJK_ERR_UNSUPPORTED, "recon_radius must be in [0, 8]");

  // Path state of one wave (~200 B per path slot).  If the device cannot hold the configured wave (other
  // tenants, a smaller part), halve the wave until it fits: only launch shapes change, never results.
  size_t n_slots = 0;
  for (;;) {
    n_slots = n_pixels * wave_passes;
    cudaError_t e = cudaSuccess;
    auto want = [&](DevBuf<f4>& b) {
      if (e == cudaSuccess) e = b.ensure(n_slots);
    };
    for (int par = 0; par < 2; par++) {  // queue-ordered path state, ping-pong by bounce parity
      want(c->d_ray_o[par]), want(c->d_ray_d[par]), want(c->d_thr[par]);
      if (c->has_extinction) want(c->d_ext[par]);
    }
    want(c->d_hit), want(c->d_layer0), want(c->d_layer1), want(c->d_sh_o), want(c->d_sh_d), want(c->d_sh_c);
    if (e == cudaSuccess) e = c->d_ext_q0.ensure(n_slots);
    if (e == cudaSuccess) e = c->d_ext_q1.ensure(n_slots);
    if (e == cudaSuccess) break;
    if (e != cudaErrorMemoryAllocation || wave_passes == 1) HJK_CUDA(c, e);
    cudaGetLastError();  // clear the allocation error, give back what was taken and retry with half the wave
    for (int par = 0; par < 2; par++)
      c->d_ray_o[par].release(), c->d_ray_d[par].release(), c->d_thr[par].release(), c->d_ext[par].release();
    c->d_hit.release(), c->d_layer0.release(), c->d_layer1.release(), c->d_sh_o.release(), c->d_sh_d.release();
    c->d_sh_c.release(), c->d_ext_q0.release(), c->d_ext_q1.release();
    c->have_features = false;
    wave_passes = (wave_passes + 1) / 2;
  }
  const size_t n_ctr = ((size_t)prm->max_bounces + 1) * CTR_STRIDE;
  HJK_CUDA(c, c->d_counters.ensure(n_ctr));
  HJK_CUDA(c, c->d_unresolved.ensure(1));
  HJK_CUDA(c, cudaMemsetAsync(c->d_unresolved.p, 0, 4, c->stream));
  HJK_CUDA(c, c->d_totals.ensure(3));
  HJK_CUDA(c, cudaMemsetAsync(c->d_totals.p, 0, 3 * sizeof(unsigned long long), c->stream));
  HJK_CUDA(c, c->d_tile_block.ensure(plan.tile_block.size()));
  HJK_CUDA(c, c->d_weights.ensure((size_t)n_blocks * taps * taps));
  HJK_CUDA(c, c->d_taps.ensure((size_t)n_blocks * recon_tap_stride(R)));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_tile_block.p, plan.tile_block.data(), plan.tile_block.size() * 4,
                              cudaMemcpyHostToDevice, c->stream));

  HJK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  uint64_t launches = 0;
  if (stats) {
    const float keep_ms = 0.f;
    (void)keep_ms;
    memset(stats, 0, sizeof(*stats));
  }
  if (do_recon) {
    KernelTimer t(c, stats, HJK_K_OTHER);
    k_recon_weights<<<std::max<int>(1, (int)std::min<size_t>(1024, ((size_t)n_blocks + 127) / 128)), 128, 0,
                      c->stream>>>(d_blocks, (uint32_t)n_blocks, R, prm->recon_stddev, c->d_weights.p, c->d_taps.p);
    launches++;
  }

  WaveDev w{};
  w.scene = c->scene;
  w.width = plan.width, w.height = plan.height, w.n_pixels = (uint32_t)n_pixels;
  w.tile_w = plan.tile_w, w.tile_h = plan.tile_h, w.tiles_x = plan.tiles_x, w.tiles_y = plan.tiles_y;
  w.blocks = d_blocks;
  w.weights = c->d_weights.p;
  for (int par = 0; par < 2; par++) {
    w.ray_o[par] = c->d_ray_o[par].p, w.ray_d[par] = c->d_ray_d[par].p, w.thr_rng[par] = c->d_thr[par].p;
    w.extinction[par] = c->d_ext[par].p;
  }
  w.hit = c->d_hit.p;
  w.layer0 = c->d_layer0.p, w.layer1 = c->d_layer1.p;
  w.ext_q[0] = c->d_ext_q0.p, w.ext_q[1] = c->d_ext_q1.p;
  w.sh_o = c->d_sh_o.p, w.sh_d = c->d_sh_d.p, w.sh_c = c->d_sh_c.p;
  w.counters = c->d_counters.p;
  w.accumulator = c->acc();
  w.max_bounces = prm->max_bounces, w.rr_start = prm->rr_start;
  w.recon_radius = R;
  w.eps = prm->eps;
  w.has_extinction = c->has_extinction ? 1u : 0u;
  // Scheduling constants of k_trace_coop the caller left to the library.  Trees that hold nothing but spheres (the
  // 512-sphere lattice) want fuller warps before a refill and a higher bar for pooling — a sphere test is half a
  // triangle test, so the pooling overhead weighs more: 1958 -> 1984 (cost 260) / 1990 (threshold 24) Mrays/s at 64
  // bounces; on cbox 20 / 24 and 180 / 240 measure the same within 0.3 %.
  // A tree far larger than the L2 (the 10 M-triangle terrain: 560 MB) also wants the fuller warps: its node fetches are
  // DRAM round trips, and more rays in flight hide them: 2254 -> 2293 Mrays/s (cbox: 29.65 -> 29.71 ms, so not there).
  const bool sphere_tree = c->bvh_all_guarded;
  const bool large_tree = c->n_nodes * sizeof(WideNode) + c->n_prims * sizeof(WidePrim) > (size_t)64 << 20;
  w.fetch_threshold =
      c->fetch_threshold >= 0 ? (uint32_t)c->fetch_threshold : (sphere_tree || large_tree ? 24u : (uint32_t)kFetchThreshold);
  // a postponed primitive group takes a second stack entry on its level: only trees of at most kMaxStack / 2 levels
  // (8^16 leaves) leave room for that, deeper ones are walked without postponing
  w.postpone_lanes = c->lane_postpones ? c->postpone_lanes : 0u;
  w.unresolved = c->d_unresolved.p;
  w.coop_batch_cost = c->coop_batch_cost >= 0 ? (uint32_t)c->coop_batch_cost : (sphere_tree ? 260u : 180u);
  const bool exact = (prm->flags & HJK_RENDER_EXACT_TIES) != 0;
  const bool guard = c->scene.num_spheres != 0;
  const bool use_coop = c->coop_trace && !exact;
  const bool shade_sort = c->shade_sort < 0 ? c->scene_mixes_materials : c->shade_sort != 0;
  w.stack_cap = use_coop ? c->stack_cap_coop : c->stack_cap_lane;
  const size_t sm_trav = (size_t)w.stack_cap * kTravThreads * sizeof(uint2);
  const int g_trav = grid_for(c, c->blocks_trav_override ? c->blocks_trav_override : c->blocks_trav_v[guard][exact]);
  const int g_tile = grid_for(c, c->blocks_tile);
  const int g_light = grid_for(c, c->blocks_light);
  uint64_t n_ext = 0, n_sh = 0, n_paths = 0;
  c->h_counters.resize(n_ctr);
  // with the reference's bounce limit (1000) paths die by roulette long before the limit:
  // look at the live count every `check_every` bounces and stop when the wave is empty
  const uint32_t check_every = 8;

  for (uint32_t p0 = 0; p0 < n_passes; p0 += wave_passes) {
    const uint32_t wp = std::min(wave_passes, n_passes - p0);
    w.n_wave_passes = wp;
    w.n_slots = (uint32_t)(n_pixels * wp);
    w.tile_block = c->d_tile_block.p + (size_t)p0 * tiles;
    HJK_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, n_ctr * 4, c->stream));
    {
      KernelTimer t(c, stats, HJK_K_RAYGEN);
      k_raygen<<<g_light, kTileThreads, 0, c->stream>>>(w);
      launches++;
    }
    // Extension rays exist for bounces [0, last).  Launch b traces them together with the shadow
    // rays bounce b-1 emitted; one more launch after the last bounce drains its shadow rays.
    uint32_t last = prm->max_bounces;
    for (uint32_t b = 0;; b++) {
      {
        KernelTimer t(c, stats, HJK_K_EXTEND);
        if (use_coop) {
          const int gv = guard ? (c->bvh_all_guarded ? 2 : 1) : 0;
          const int g_coop = grid_for(c, c->blocks_trav_override ? c->blocks_trav_override : c->blocks_coop[gv]);
          if (gv == 2)
            k_trace_coop<2><<<g_coop, kTravThreads, sm_trav, c->stream>>>(w, b, last);
          else if (gv == 1)
            k_trace_coop<1><<<g_coop, kTravThreads, sm_trav, c->stream>>>(w, b, last);
          else
            k_trace_coop<0><<<g_coop, kTravThreads, sm_trav, c->stream>>>(w, b, last);
        } else if (guard && exact)
          k_trace<true, true><<<g_trav, kTravThreads, sm_trav, c->stream>>>(w, b, last);
        else if (guard)
          k_trace<true, false><<<g_trav, kTravThreads, sm_trav, c->stream>>>(w, b, last);
        else if (exact)
          k_trace<false, true><<<g_trav, kTravThreads, sm_trav, c->stream>>>(w, b, last);
        else
          k_trace<false, false><<<g_trav, kTravThreads, sm_trav, c->stream>>>(w, b, last);
        launches++;
      }
      if (b == last) break;
      {
        KernelTimer t(c, stats, HJK_K_SHADE);
        if (shade_sort)
          k_shade<true><<<g_tile, kShadeThreads, 0, c->stream>>>(w, b);
        else
          k_shade<false><<<g_tile, kShadeThreads, 0, c->stream>>>(w, b);
      }
      launches++;
      if (b + 1 < last && (b + 1) % check_every == 0) {
        uint32_t live = 0;
        HJK_CUDA(c, cudaMemcpyAsync(&live, c->d_counters.p + (size_t)(b + 1) * CTR_STRIDE + CTR_EXT, 4,
                                    cudaMemcpyDeviceToHost, c->stream));
        HJK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (live == 0) last = b + 1;
      }
    }
    // fold this wave's ray counts into the call's totals on the device (no host sync per wave)
    k_wave_totals<<<1, 32, 0, c->stream>>>(c->d_counters.p, prm->max_bounces + 1, c->d_totals.p);
    launches++;
    HJK_CUDA(c, cudaGetLastError());
    if (do_recon) {
      KernelTimer t(c, stats, HJK_K_RECON);
      PassDev ps{};
      ps.width = plan.width, ps.height = plan.height, ps.tile_w = plan.tile_w, ps.tile_h = plan.tile_h;
      ps.tiles_x = plan.tiles_x, ps.tiles_y = plan.tiles_y;
      ps.tile_block = w.tile_block;
      ps.blocks = d_blocks;
      ps.weights = c->d_weights.p;
      ps.taps = c->d_taps.p;
      ps.radius = R;
      rc = launch_recon(c, ps, wp, w.layer0, w.layer1, nullptr, c->acc(), c->feature_buffers);
      if (rc) return rc;
      launches++;
    }
    c->last_wave_passes = wp;
    c->have_features = true;
  }
  HJK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  if (!(prm->flags & HJK_RENDER_ASYNC) || stats) {
    HJK_CUDA(c, cudaEventSynchronize(c->ev1));
    resolve_timers(c, stats);
    if (stats) {
      unsigned long long totals[3] = {0, 0, 0};
      HJK_CUDA(c, cudaMemcpy(totals, c->d_totals.p, sizeof totals, cudaMemcpyDeviceToHost));
      uint32_t unresolved = 0;
      HJK_CUDA(c, cudaMemcpy(&unresolved, c->d_unresolved.p, 4, cudaMemcpyDeviceToHost));
      c->unresolved_last = unresolved;
      n_paths = totals[0], n_ext = totals[1], n_sh = totals[2];
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->ev0, c->ev1);
      stats->ms_total = ms;
      stats->n_paths = n_paths;
      stats->n_extension_rays = n_ext;
      stats->n_shadow_rays = n_sh;
      stats->n_launches = launches;
    }
  }
  return HJK_OK;
}


// GPU build of the wide BVH from the scene arrays already resident on the device (bvh_build_gpu.cuh).
// Returns HJK_OK and fills d_nodes/d_prims + meta, or HJK_ERR_UNSUPPORTED when the host builder must
// be used instead (tiny scenes, too deep a tree, capacity overflow).
int build_bvh_gpu(HjkContext* c, uint32_t S, uint32_t Q, uint32_t T, float pad_rel, WideBvh& meta) {
  using namespace gpubvh;
  const uint32_t n = S + Q + T;
  if (n < 16u) return HJK_ERR_UNSUPPORTED;
  BuildScene bs{c->d_spheres.p, c->d_quads.p, c->d_triangles.p, c->d_vertices.p, S, Q, T};
  cudaStream_t st = c->stream;
  const int grid = c->n_sms * 4, block = 256;
  DevBuf<f4> blo, bhi, ilo, ihi;
  DevBuf<uint32_t> small, vals, vals_sorted, child_l, child_r, parent_inner, parent_leaf, visits, icount;
  DevBuf<uint64_t> keys, keys_sorted;
  DevBuf<float> pad;
  DevBuf<uint8_t> sort_tmp;
  DevBuf<WideNode> tmp_nodes;
  DevBuf<uint2> tasks_a, tasks_b;
  HJK_CUDA(c, blo.ensure(n));
  HJK_CUDA(c, bhi.ensure(n));
  HJK_CUDA(c, small.ensure(16));  // [0..5] bounds, [6] bad, [8..10] counters, [12] n_in, [13] n_out
  HJK_CUDA(c, keys.ensure(n));
  HJK_CUDA(c, keys_sorted.ensure(n));
  HJK_CUDA(c, vals.ensure(n));
  HJK_CUDA(c, vals_sorted.ensure(n));
  HJK_CUDA(c, pad.ensure(1));
  const uint32_t init[16] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};  // [12] = one root task
  HJK_CUDA(c, cudaMemcpyAsync(small.p, init, sizeof init, cudaMemcpyHostToDevice, st));
  HJK_CUDA(c, cudaEventRecord(c->ev0, st));
  k_shape_boxes<<<grid, block, 0, st>>>(bs, n, blo.p, bhi.p, small.p, small.p + 6);
  k_morton<<<grid, block, 0, st>>>(n, small.p, pad_rel, blo.p, bhi.p, keys.p, vals.p, pad.p);
  size_t tmp_bytes = 0;
  HJK_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n, 0,
                                              63, st));
  HJK_CUDA(c, sort_tmp.ensure(tmp_bytes));
  HJK_CUDA(c, cub::DeviceRadixSort::SortPairs(sort_tmp.p, tmp_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n,
                                              0, 63, st));
  HJK_CUDA(c, child_l.ensure(n));
  HJK_CUDA(c, child_r.ensure(n));
  HJK_CUDA(c, icount.ensure(n));
  HJK_CUDA(c, ilo.ensure(n));
  HJK_CUDA(c, ihi.ensure(n));
  HJK_CUDA(c, parent_inner.ensure(n));
  HJK_CUDA(c, parent_leaf.ensure(n));
  HJK_CUDA(c, visits.ensure(n));
  if (c->bvh_gpu_tree == 0) {  // binary radix tree + bottom-up fitting
    HJK_CUDA(c, cudaMemsetAsync(visits.p, 0, (size_t)n * 4, st));
    k_radix_tree<<<grid, block, 0, st>>>((int)n, keys_sorted.p, child_l.p, child_r.p, parent_inner.p, parent_leaf.p);
    k_fit_boxes<<<grid, block, 0, st>>>((int)n, S, vals_sorted.p, blo.p, bhi.p, child_l.p, child_r.p, parent_inner.p,
                                        parent_leaf.p, visits.p, ilo.p, ihi.p, icount.p);
    HJK_CUDA(c, cudaGetLastError());
  } else {  // PLOC: clusters merge round by round until the root is left (one 4-byte read back per round)
    DevBuf<uint32_t> ref[2], ccnt[2], nn, valid, pos, state;
    DevBuf<f4> clo[2], chi[2];
    DevBuf<uint8_t> scan_tmp;
    for (int k = 0; k < 2; k++) {
      HJK_CUDA(c, ref[k].ensure(n));
      HJK_CUDA(c, ccnt[k].ensure(n));
      HJK_CUDA(c, clo[k].ensure(n));
      HJK_CUDA(c, chi[k].ensure(n));
    }
    HJK_CUDA(c, nn.ensure(n));
    HJK_CUDA(c, valid.ensure(n));
    HJK_CUDA(c, pos.ensure(n));
    HJK_CUDA(c, state.ensure(2));
    HJK_CUDA(c, cudaMemsetAsync(state.p, 0, 8, st));
    size_t scan_bytes = 0;
    HJK_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, valid.p, pos.p, (int)n, st));
    HJK_CUDA(c, scan_tmp.ensure(scan_bytes));
    k_ploc_init<<<grid, block, 0, st>>>(n, S, vals_sorted.p, blo.p, bhi.p, ref[0].p, clo[0].p, chi[0].p, ccnt[0].p);
    uint32_t m = n;
    int cur = 0;
    for (int round = 0; m > 1; round++) {
      if (round > 4096) return c->fail(HJK_ERR_CUDA, "PLOC did not converge");
      const int g_nn = (int)std::min<uint32_t>((m + kPlocThreads - 1) / kPlocThreads, (uint32_t)c->n_sms * 8u);
      k_ploc_nn<<<g_nn, kPlocThreads, 0, st>>>(m, c->bvh_ploc_radius, clo[cur].p, chi[cur].p, nn.p);
      k_ploc_flags<<<grid, block, 0, st>>>(m, nn.p, valid.p);
      HJK_CUDA(c, cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, valid.p, pos.p, (int)m, st));
      k_ploc_merge<<<grid, block, 0, st>>>(m, n, nn.p, valid.p, pos.p, ref[cur].p, clo[cur].p, chi[cur].p, ccnt[cur].p,
                                           ref[cur ^ 1].p, clo[cur ^ 1].p, chi[cur ^ 1].p, ccnt[cur ^ 1].p, child_l.p,
                                           child_r.p, ilo.p, ihi.p, icount.p, parent_inner.p, parent_leaf.p, state.p);
      k_ploc_advance<<<1, 1, 0, st>>>(m, valid.p, pos.p, state.p);
      uint32_t st2[2] = {0, 0};
      HJK_CUDA(c, cudaMemcpyAsync(st2, state.p, 8, cudaMemcpyDeviceToHost, st));
      HJK_CUDA(c, cudaStreamSynchronize(st));
      if (st2[1] >= m || st2[1] == 0) return c->fail(HJK_ERR_CUDA, "PLOC made no progress");
      m = st2[1];
      cur ^= 1;
    }
  }
  const uint32_t node_capacity = n;
  HJK_CUDA(c, tmp_nodes.ensure(node_capacity));
  HJK_CUDA(c, c->d_prims.ensure((size_t)n * HJK_PRIM_STRIDE));
  HJK_CUDA(c, tasks_a.ensure(n));
  HJK_CUDA(c, tasks_b.ensure(n));
  const uint2 root_task = make_uint2(0u, 0u);
  HJK_CUDA(c, cudaMemcpyAsync(tasks_a.p, &root_task, sizeof root_task, cudaMemcpyHostToDevice, st));
  // collapse costs (the host builder's dynamic programme), unless the greedy collapse was asked for
  DevBuf<float> dp_cost;
  DevBuf<uint8_t> dp_choice;
  if (c->bvh_gpu_collapse) {
    HJK_CUDA(c, dp_cost.ensure((size_t)n * 7));
    HJK_CUDA(c, dp_choice.ensure((size_t)n * 8));
    HJK_CUDA(c, cudaMemsetAsync(visits.p, 0, (size_t)n * 4, st));
    k_collapse_costs<<<grid, block, 0, st>>>((int)n, vals_sorted.p, blo.p, bhi.p, child_l.p, child_r.p, parent_inner.p,
                                             parent_leaf.p, ilo.p, ihi.p, icount.p, visits.p, dp_cost.p, dp_choice.p);
    HJK_CUDA(c, cudaGetLastError());
  }
  TreeDev tree{vals_sorted.p, blo.p, bhi.p, child_l.p, child_r.p, ilo.p, ihi.p, icount.p,
               c->bvh_gpu_collapse ? dp_choice.p : nullptr};
  uint2* t_in = tasks_a.p;
  uint2* t_out = tasks_b.p;
  // one launch per level, kMaxStack levels at most, no host round trip in between: a level without tasks is an
  // empty launch; small[12] = tasks in, [13] = tasks out, [14] = levels that had work (root level included)
  for (int level = 0; level <= kMaxStack; level++) {
    k_collapse<<<grid, block, 0, st>>>(bs, tree, t_in, small.p + 12, t_out, small.p + 13, tmp_nodes.p,
                                       (WidePrim*)c->d_prims.p, small.p + 8, node_capacity);
    k_next_level<<<1, 1, 0, st>>>(small.p + 12);
    std::swap(t_in, t_out);
  }
  HJK_CUDA(c, cudaGetLastError());
  uint32_t fin[16];
  float pad_h = 0.f;
  HJK_CUDA(c, cudaMemcpyAsync(fin, small.p, sizeof fin, cudaMemcpyDeviceToHost, st));
  HJK_CUDA(c, cudaMemcpyAsync(&pad_h, pad.p, 4, cudaMemcpyDeviceToHost, st));
  HJK_CUDA(c, cudaStreamSynchronize(st));
  if (fin[12] != 0) return HJK_ERR_UNSUPPORTED;  // deeper than the traversal stack: the host builder decides
  const uint32_t depth = fin[14] + 1;
  if (fin[6]) return c->fail(HJK_ERR_INVALID_ARGUMENT, "non-finite shape bounds");
  if (fin[10] || fin[9] != n) return HJK_ERR_UNSUPPORTED;  // capacity overflow / primitive count mismatch
  const uint32_t n_nodes = fin[8];
  HJK_CUDA(c, c->d_nodes.ensure((size_t)n_nodes * 5));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_nodes.p, tmp_nodes.p, (size_t)n_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
  HJK_CUDA(c, cudaEventRecord(c->ev1, st));
  HJK_CUDA(c, cudaEventSynchronize(c->ev1));
  cudaEventElapsedTime(&c->bvh_build_ms, c->ev0, c->ev1);
  meta.depth = depth;
  meta.n_shapes = n;
  meta.pad = pad_h;
  meta.nodes.clear();
  meta.prims.clear();
  c->n_nodes = n_nodes;
  c->n_prims = n;
  if (c->bvh_validate) {
    meta.nodes.resize(n_nodes);
    meta.prims.resize(n);
    HJK_CUDA(c, cudaMemcpy(meta.nodes.data(), c->d_nodes.p, (size_t)n_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost));
    HJK_CUDA(c, cudaMemcpy(meta.prims.data(), c->d_prims.p, (size_t)n * sizeof(WidePrim), cudaMemcpyDeviceToHost));
  }
  return HJK_OK;
}

}  // namespace

extern "C" {

const char* hjk_version(void) { return "hijiki_b200 0.1 (sm_100a)"; }

const char* hjk_last_error(const HjkContext* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

// Launch shape of the trace kernels for a tree of `depth` levels: the traversal stack is dynamic shared memory, one
// entry per level (+1) in k_trace_coop, two per level where the per-lane kernels postpone primitive groups (only
// when that still fits kMaxStack entries), and the resident CTAs per SM follow from it.
static void trace_launch_shape(HjkContext* c, uint32_t depth) {
  c->stack_cap_coop = std::max<uint32_t>(depth + 1, 2);
  c->lane_postpones = 2 * depth + 2 <= (uint32_t)kMaxStack;
  c->stack_cap_lane = c->lane_postpones ? 2 * depth + 2 : std::max<uint32_t>(depth + 1, 2);
  const size_t sm_coop = (size_t)c->stack_cap_coop * kTravThreads * sizeof(uint2);
  const size_t sm_lane = (size_t)c->stack_cap_lane * kTravThreads * sizeof(uint2);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<false, false>, kTravThreads, sm_lane);
  c->blocks_trav_v[0][0] = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<false, true>, kTravThreads, sm_lane);
  c->blocks_trav_v[0][1] = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<true, false>, kTravThreads, sm_lane);
  c->blocks_trav_v[1][0] = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace<true, true>, kTravThreads, sm_lane);
  c->blocks_trav_v[1][1] = std::max(occ, 1);
  c->blocks_trav = c->blocks_trav_v[0][0];
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_batch<0, false>, kTravThreads, sm_lane);
  c->blocks_batch = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_batch<1, true>, kTravThreads, sm_lane);
  c->blocks_batch = std::max(std::min(c->blocks_batch, occ), 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_coop<0>, kTravThreads, sm_coop);
  c->blocks_coop[0] = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_coop<1>, kTravThreads, sm_coop);
  c->blocks_coop[1] = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_coop<2>, kTravThreads, sm_coop);
  c->blocks_coop[2] = std::max(occ, 1);
}

static void destroy_one(HjkContext* c);
static int create_one(int device, HjkContext** out_ctx) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return HJK_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "hjk_create: device id out of range";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  HjkContext* c = new (std::nothrow) HjkContext();
  if (!c) return HJK_ERR_OUT_OF_MEMORY;
  c->device = device;
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    delete c;
    return HJK_ERR_CUDA;
  };
  if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, c->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (prop.major < 10) {
    g_create_error = "hijiki_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
    delete c;
    return HJK_ERR_UNSUPPORTED;
  }
  c->n_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  c->own_stream = true;
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  trace_launch_shape(c, 7);  // refreshed by every scene upload for the depth of its tree
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shade<true>, kShadeThreads, 0);
  c->blocks_tile = std::max(occ, 1);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_raygen, kTileThreads, 0);
  c->blocks_light = std::max(occ, 1);
  if ((e = cudaGetLastError()) != cudaSuccess) return bail("kernel image (built for sm_100a)", e);
  // tuning sweeps without touching the host program: HJK_OPTIONS="key=value,key=value" (hjk_set_option keys)
  if (const char* env = getenv("HJK_OPTIONS")) {
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
      const size_t end = std::min(all.find(',', pos), all.size());
      const std::string kv = all.substr(pos, end - pos);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos && hjk_set_option(c, kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1)) != HJK_OK) {
        g_create_error = "HJK_OPTIONS: " + c->error;
        destroy_one(c);
        return HJK_ERR_INVALID_ARGUMENT;
      }
      pos = end + 1;
    }
  }
  *out_ctx = c;
  return HJK_OK;
}

static void destroy_one(HjkContext* c) {
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->resident)
    if (kv.second.dev) cudaFree(kv.second.dev);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->ev_staged);
    cudaEventDestroy(c->ev_copied);
  }
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int hjk_create(const int* device_ids, int n_devices, HjkContext** out_ctx) {
  if (!out_ctx || n_devices < 1 || n_devices > 64 || !device_ids) {
    g_create_error = "hjk_create: needs 1..64 device ids";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (device_ids[i] == device_ids[j]) {
        g_create_error = "hjk_create: a device is listed twice";
        return HJK_ERR_INVALID_ARGUMENT;
      }
  std::vector<HjkContext*> all;
  auto undo = [&]() {
    for (HjkContext* m : all) destroy_one(m);
  };
  for (int i = 0; i < n_devices; i++) {
    HjkContext* c = nullptr;
    const int rc = create_one(device_ids[i], &c);
    if (rc) {
      undo();
      return rc;
    }
    all.push_back(c);
  }
  if (n_devices > 1) {  // one communicator per device, created together (single process, no id exchange)
    std::string err;
    if (!load_nccl(err)) {
      g_create_error = err;
      undo();
      return HJK_ERR_NCCL;
    }
    std::vector<void*> comms((size_t)n_devices, nullptr);
    const int r = g_nccl.CommInitAll(comms.data(), n_devices, device_ids);
    if (r != 0) {
      g_create_error = std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
      undo();
      return HJK_ERR_NCCL;
    }
    for (int i = 0; i < n_devices; i++) {
      all[i]->comm = comms[i];
      all[i]->rank = i, all[i]->n_ranks = n_devices;
      all[i]->group_parent = i ? all[0] : nullptr;
    }
    all[0]->members = all;
    cudaSetDevice(all[0]->device);
  }
  *out_ctx = all[0];
  return HJK_OK;
}

int hjk_destroy(HjkContext* c) {
  if (!c) return HJK_OK;
  for (size_t i = 1; i < c->members.size(); i++) destroy_one(c->members[i]);
  destroy_one(c);
  return HJK_OK;
}

int hjk_set_stream(HjkContext* c, void* cuda_stream) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  cudaStreamSynchronize(c->stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return HJK_OK;
}

// hjk_scene_upload for one device.  shared != nullptr: upload this host-built wide BVH instead of building one
// (the other devices of a single-process group).  keep != nullptr: hand the host-built tree back to the caller.
// With several ranks joined and the option "bvh_broadcast" on, rank 0 alone builds; the others receive nodes and
// primitive records over ncclBroadcast.
static int scene_upload_impl(HjkContext* c, const HjkScene* s, const WideBvh* shared, WideBvh* keep) {
  if (!s || !s->scene.ptr || s->scene.count != 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "scene info missing");
  HJK_CUDA(c, cudaSetDevice(c->device));
  const HjkSceneInfo* info = (const HjkSceneInfo*)s->scene.ptr;
  if (info->num_spheres != s->spheres.count || info->num_quads != s->quads.count ||
      info->num_triangles != s->triangles.count || info->num_emitters != s->emitters.count)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "SceneBufferInfo counts disagree with the array counts");
  const uint64_t n_shapes = s->spheres.count + s->quads.count + s->triangles.count;
  // sampleEmitter (scene.glsl:54-89) reads emitters[0] unconditionally at every diffuse hit: a scene without
  // emitters has no defined result in the reference either
  if (s->emitters.count == 0)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "the scene has no emitter (no shape with an emissive material): "
                                             "next-event estimation needs at least one");
  if (s->materials.count != n_shapes)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "materials must hold one word per shape");
  // validate material words and emitter table against the typed arrays they index
  const uint32_t* mats = (const uint32_t*)s->materials.ptr;
  bool has_ext = false;
  uint32_t classes = 0;  // bit 0 diffuse-like, 1 mirror, 2 dielectric
  for (uint64_t i = 0; i < n_shapes; i++) {
    const uint32_t tag = mats[i] >> HJK_MATERIAL_TAG_SHIFT, idx = mats[i] & ((1u << HJK_MATERIAL_TAG_SHIFT) - 1u);
    classes |= tag == HJK_MAT_MIRROR ? 2u : tag == HJK_MAT_DIELECTRIC ? 4u : tag == HJK_MAT_EMISSIVE ? 0u : 1u;
    uint64_t limit = 1;
    switch (tag) {
      case HJK_MAT_DIFFUSE: limit = s->diffuse.count; break;
      case HJK_MAT_DIFFUSECBOARD: limit = s->diffusecb.count; break;
      case HJK_MAT_MIRROR: limit = 0xFFFFFFFFull; break;
      case HJK_MAT_DIELECTRIC: limit = s->dielectric.count; break;
      case HJK_MAT_EMISSIVE: limit = s->emissive.count; break;
      default: return c->fail(HJK_ERR_INVALID_ARGUMENT, "shape %llu has unknown material tag %u", (unsigned long long)i, tag);
    }
    if (idx >= limit) return c->fail(HJK_ERR_INVALID_ARGUMENT, "shape %llu: material index out of range", (unsigned long long)i);
  }
  for (uint64_t i = 0; i < s->dielectric.count; i++) {
    const HjkDielectric& d = ((const HjkDielectric*)s->dielectric.ptr)[i];
    if (d.extinction_eta[0] != 0.f || d.extinction_eta[1] != 0.f || d.extinction_eta[2] != 0.f) has_ext = true;
  }
  for (uint64_t i = 0; i < s->emitters.count; i++) {
    const HjkEmitter& e = ((const HjkEmitter*)s->emitters.ptr)[i];
    if (e.shape >= n_shapes || (mats[e.shape] >> HJK_MATERIAL_TAG_SHIFT) != HJK_MAT_EMISSIVE)
      return c->fail(HJK_ERR_INVALID_ARGUMENT, "emitter %llu does not point at an emissive shape", (unsigned long long)i);
  }
  int rc;
  if ((rc = upload(c, c->d_spheres, s->spheres, 16))) return rc;
  if ((rc = upload(c, c->d_quads, s->quads, 48))) return rc;
  if ((rc = upload(c, c->d_triangles, s->triangles, 12))) return rc;
  if ((rc = upload(c, c->d_vertices, s->vertices, 32))) return rc;
  if ((rc = upload(c, c->d_materials, s->materials, 4))) return rc;
  if ((rc = upload(c, c->d_emitters, s->emitters, 16))) return rc;
  if ((rc = upload(c, c->d_diffuse, s->diffuse, 16))) return rc;
  if ((rc = upload(c, c->d_diffusecb, s->diffusecb, 32))) return rc;
  if ((rc = upload(c, c->d_dielectric, s->dielectric, 16))) return rc;
  if ((rc = upload(c, c->d_emissive, s->emissive, 16))) return rc;
  if (s->triangles.count) {
    const uint32_t* tri = (const uint32_t*)s->triangles.ptr;
    for (uint64_t i = 0; i < 3 * s->triangles.count; i++)
      if (tri[i] >= s->vertices.count) return c->fail(HJK_ERR_INVALID_ARGUMENT, "triangle vertex index out of range");
  }
  WideBvh bvh;
  std::string err;
  bool built_on_gpu = false;
  c->bvh_build_ms = 0.f;
  const bool bcast = c->comm && c->n_ranks > 1 && c->members.size() <= 1 && !c->group_parent && c->bvh_broadcast;
  auto all_guarded = [](const WideBvh& b) {
    if (b.nodes.empty()) return false;
    for (const WideNode& wn : b.nodes)
      if (!(wn.prim_base & kWideHasSpheres)) return false;
    return true;
  };
  auto upload_tree = [&](const WideBvh& b) {
    HjkArray a_nodes{b.nodes.data(), b.nodes.size()}, a_prims{b.prims.data(), b.prims.size()};
    int r;
    if ((r = upload(c, c->d_nodes, a_nodes, sizeof(WideNode)))) return r;
    if ((r = upload(c, c->d_prims, a_prims, sizeof(WidePrim)))) return r;
    c->n_nodes = b.nodes.size();
    c->n_prims = b.prims.size();
    c->bvh_all_guarded = all_guarded(b);
    return (int)HJK_OK;
  };
  if (shared) {
    if ((rc = upload_tree(*shared))) return rc;
    bvh.depth = shared->depth, bvh.n_shapes = shared->n_shapes, bvh.pad = shared->pad, bvh.sah_cost = shared->sah_cost;
    for (int k = 0; k < 4; k++) bvh.sph_centre[k] = shared->sph_centre[k];
    bvh.sph_rmin = shared->sph_rmin, bvh.sph_rmax = shared->sph_rmax;
  } else if (bcast && c->rank != 0) {
    // nothing to build: the tree arrives below
  } else {
    // by default the GPU builder takes the scenes the host builder needs seconds for
    const bool use_gpu_builder = c->bvh_builder == 1 || (c->bvh_builder < 0 && n_shapes > kGpuBuilderMinShapes);
    if (use_gpu_builder) {
      rc = build_bvh_gpu(c, info->num_spheres, info->num_quads, info->num_triangles, c->bvh_pad_rel, bvh);
      if (rc == HJK_OK) {
        built_on_gpu = true;
        // the GPU builder flags its nodes per subtree too; "every node guarded" only holds for all-sphere scenes
        c->bvh_all_guarded = info->num_spheres != 0 && info->num_quads == 0 && info->num_triangles == 0;
        sphere_guard_bounds(*s, bvh);
        if (c->bvh_validate && !validate_wide_bvh(*s, bvh, err))
          return c->fail(HJK_ERR_CUDA, "GPU-built BVH failed the structural check: %s", err.c_str());
        bvh.nodes.clear();
        bvh.prims.clear();
      } else if (rc != HJK_ERR_UNSUPPORTED) {
        return rc;
      }
    }
    if (!built_on_gpu) {
      if (!build_wide_bvh(*s, c->bvh_pad_rel, bvh, err)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "%s", err.c_str());
      if (bvh.depth > (uint32_t)kMaxStack)
        return c->fail(HJK_ERR_UNSUPPORTED, "wide BVH depth %u exceeds the traversal stack (%d)", bvh.depth, kMaxStack);
      if ((rc = upload_tree(bvh))) return rc;
    }
  }
  if (bcast) {  // header (counts, depth, sphere-guard constants), then the two arrays, root 0
    struct Header {
      uint32_t n_nodes, n_prims, depth, all_guarded;
      float sph_centre[4], sph_rmin, sph_rmax, pad, sah;
    } hd{};
    static_assert(sizeof(Header) == 48, "header layout");
    DevBuf<uint32_t> d_hd;
    HJK_CUDA(c, d_hd.ensure(12));
    if (c->rank == 0) {
      hd.n_nodes = (uint32_t)c->n_nodes, hd.n_prims = (uint32_t)c->n_prims, hd.depth = bvh.depth;
      hd.all_guarded = c->bvh_all_guarded ? 1u : 0u;
      for (int k = 0; k < 4; k++) hd.sph_centre[k] = bvh.sph_centre[k];
      hd.sph_rmin = bvh.sph_rmin, hd.sph_rmax = bvh.sph_rmax, hd.pad = bvh.pad, hd.sah = bvh.sah_cost;
      HJK_CUDA(c, cudaMemcpyAsync(d_hd.p, &hd, sizeof hd, cudaMemcpyHostToDevice, c->stream));
    }
    int r = g_nccl.Broadcast(d_hd.p, d_hd.p, sizeof hd, /*ncclChar*/ 0, 0, c->comm, c->stream);
    if (r != 0) return c->fail(HJK_ERR_NCCL, "ncclBroadcast: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    HJK_CUDA(c, cudaMemcpyAsync(&hd, d_hd.p, sizeof hd, cudaMemcpyDeviceToHost, c->stream));
    HJK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->rank != 0) {
      HJK_CUDA(c, c->d_nodes.ensure((size_t)hd.n_nodes * 5));
      HJK_CUDA(c, c->d_prims.ensure((size_t)hd.n_prims * HJK_PRIM_STRIDE));
      c->n_nodes = hd.n_nodes, c->n_prims = hd.n_prims;
      c->bvh_all_guarded = hd.all_guarded != 0;
      bvh.depth = hd.depth, bvh.pad = hd.pad, bvh.sah_cost = hd.sah, bvh.n_shapes = (uint32_t)n_shapes;
      for (int k = 0; k < 4; k++) bvh.sph_centre[k] = hd.sph_centre[k];
      bvh.sph_rmin = hd.sph_rmin, bvh.sph_rmax = hd.sph_rmax;
    }
    r = g_nccl.GroupStart();
    if (r == 0) r = g_nccl.Broadcast(c->d_nodes.p, c->d_nodes.p, (size_t)hd.n_nodes * sizeof(WideNode), 0, 0, c->comm, c->stream);
    if (r == 0) r = g_nccl.Broadcast(c->d_prims.p, c->d_prims.p, (size_t)hd.n_prims * sizeof(WidePrim), 0, 0, c->comm, c->stream);
    const int r2 = g_nccl.GroupEnd();
    if (r != 0 || r2 != 0)
      return c->fail(HJK_ERR_NCCL, "ncclBroadcast: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r ? r : r2) : "error");
  }
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope

  SceneDev& d = c->scene;
  d.nodes = c->d_nodes.p, d.prims = c->d_prims.p;
  d.spheres = c->d_spheres.p, d.quads = c->d_quads.p, d.triangles = c->d_triangles.p;
  d.vertices = c->d_vertices.p, d.materials = c->d_materials.p, d.emitters = c->d_emitters.p;
  d.diffuse = c->d_diffuse.p, d.diffusecb = c->d_diffusecb.p, d.dielectric = c->d_dielectric.p;
  d.emissive = c->d_emissive.p;
  d.num_spheres = info->num_spheres, d.num_quads = info->num_quads;
  d.num_triangles = info->num_triangles, d.num_emitters = info->num_emitters;
  d.camera = info->camera;
  for (int k = 0; k < 4; k++) d.sph_centre[k] = bvh.sph_centre[k];
  d.sph_rmin = bvh.sph_rmin, d.sph_rmax = bvh.sph_rmax;
  c->has_extinction = has_ext;
  c->scene_mixes_materials = (classes & (classes - 1u)) != 0u;
  if (keep) *keep = bvh;  // (empty vectors after a GPU build: the other devices of a group then build their own)
  bvh.nodes.clear();
  bvh.nodes.shrink_to_fit();
  bvh.prims.clear();
  bvh.prims.shrink_to_fit();
  c->bvh_host_stats = bvh;
  trace_launch_shape(c, bvh.depth);
  {
    const unsigned int zero = 0;
    HJK_CUDA(c, cudaMemcpyToSymbol(g_stack_overflows, &zero, sizeof zero));
  }
  c->has_scene = true;
  return HJK_OK;
}

int hjk_scene_upload(HjkContext* c, const HjkScene* s) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (c->members.size() <= 1) return scene_upload_impl(c, s, nullptr, nullptr);
  // single-process group: the host builder runs once, every device gets a copy of its tree
  WideBvh tree;
  int rc = scene_upload_impl(c, s, nullptr, &tree);
  if (rc) return rc;
  for (size_t i = 1; i < c->members.size(); i++) {
    HjkContext* m = c->members[i];
    m->bvh_builder = c->bvh_builder, m->bvh_pad_rel = c->bvh_pad_rel, m->bvh_validate = c->bvh_validate;
    m->bvh_gpu_tree = c->bvh_gpu_tree, m->bvh_gpu_collapse = c->bvh_gpu_collapse, m->bvh_ploc_radius = c->bvh_ploc_radius;
    rc = scene_upload_impl(m, s, tree.nodes.empty() ? nullptr : &tree, nullptr);
    if (rc) return c->fail(rc, "device %d: %s", m->device, m->error.c_str());
  }
  return HJK_OK;
}

int hjk_frame_begin(HjkContext* c, uint32_t width, uint32_t height) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (width == 0 || height == 0 || (uint64_t)width * height > 0x7FFFFFFFull)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "unsupported frame size");
  for (size_t i = 1; i < c->members.size(); i++) {
    HjkContext* m = c->members[i];
    m->feature_buffers = c->feature_buffers;
    const int rc = hjk_frame_begin(m, width, height);
    if (rc) return c->fail(rc, "device %d: %s", m->device, m->error.c_str());
  }
  HJK_CUDA(c, cudaSetDevice(c->device));
  c->have_features = false;
  return ensure_frame(c, width, height, true);
}

// Single-process group: sample pass p of the list goes to device p mod n; one host thread per device
// drives its wave loop, the frames meet in hjk_readback's reduction.  Devices without a pass keep their zeroed frame.
static int render_group(HjkContext* c, const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm,
                        HjkStats* stats) {
  PassPlan plan;
  std::string err;
  if (!plan_passes(blocks, n_blocks, plan, err)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "%s", err.c_str());
  const size_t n = c->members.size();
  std::vector<std::vector<HjkImageBlock>> lists(n);
  for (size_t p = 0; p < plan.passes.size(); p++) {
    const PassPlan::Pass& ps = plan.passes[p];
    lists[p % n].insert(lists[p % n].end(), blocks + ps.first_block, blocks + ps.first_block + ps.n_blocks);
  }
  std::vector<int> rcs(n, HJK_OK);
  std::vector<HjkStats> sts(n);
  std::vector<std::thread> threads;
  auto work = [&](size_t i) {
    HjkContext* m = c->members[i];
    memset(&sts[i], 0, sizeof(HjkStats));
    if (lists[i].empty()) return;
    m->profiling = c->profiling;
    if (cudaSetDevice(m->device) != cudaSuccess) {
      rcs[i] = m->fail(HJK_ERR_CUDA, "cudaSetDevice failed");
      return;
    }
    if (m->width != plan.width || m->height != plan.height || !m->d_frame.p) {
      rcs[i] = ensure_frame(m, plan.width, plan.height, true);
      if (rcs[i]) return;
    }
    cudaError_t e = m->d_blocks.ensure(lists[i].size());
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(m->d_blocks.p, lists[i].data(), lists[i].size() * sizeof(HjkImageBlock), cudaMemcpyHostToDevice,
                          m->stream);
    if (e != cudaSuccess) {
      rcs[i] = m->fail(HJK_ERR_CUDA, "block upload failed: %s", cudaGetErrorString(e));
      return;
    }
    HjkParams p = *prm;
    p.flags &= ~(uint32_t)HJK_RENDER_ASYNC;  // the block lists above live until the threads join
    rcs[i] = render_blocks(m, lists[i].data(), m->d_blocks.p, lists[i].size(), &p, stats ? &sts[i] : nullptr);
  };
  for (size_t i = 1; i < n; i++) {
    try {
      threads.emplace_back(work, i);
    } catch (...) {  // no thread to be had: this device's share runs on the calling thread (nothing may throw across the ABI)
      work(i);
    }
  }
  work(0);
  for (std::thread& t : threads) t.join();
  cudaSetDevice(c->device);
  c->reduced_valid = false;
  for (size_t i = 0; i < n; i++)
    if (rcs[i]) return i ? c->fail(rcs[i], "device %d: %s", c->members[i]->device, c->members[i]->error.c_str()) : rcs[i];
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (size_t i = 0; i < n; i++) {
      stats->n_paths += sts[i].n_paths;
      stats->n_extension_rays += sts[i].n_extension_rays;
      stats->n_shadow_rays += sts[i].n_shadow_rays;
      stats->n_launches += sts[i].n_launches;
      stats->ms_total = std::max(stats->ms_total, sts[i].ms_total);
      for (int k = 0; k < HJK_N_KERNEL_SLOTS; k++) stats->kernel_ms[k] = std::max(stats->kernel_ms[k], sts[i].kernel_ms[k]);
    }
  }
  return HJK_OK;
}

int hjk_render(HjkContext* c, const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm, HjkStats* stats) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!blocks || n_blocks == 0) return c->fail(HJK_ERR_INVALID_ARGUMENT, "empty block list");
  if (!prm) return c->fail(HJK_ERR_INVALID_ARGUMENT, "params is null");
  if (c->members.size() > 1) return render_group(c, blocks, n_blocks, prm, stats);
  HJK_CUDA(c, cudaSetDevice(c->device));
  HJK_CUDA(c, c->d_blocks.ensure(n_blocks));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_blocks.p, blocks, n_blocks * sizeof(HjkImageBlock), cudaMemcpyHostToDevice,
                              c->stream));
  return render_blocks(c, blocks, c->d_blocks.p, n_blocks, prm, stats);
}

int hjk_blocks_upload(HjkContext* c, const HjkImageBlock* blocks, uint64_t n_blocks, uint64_t* out_handle) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!blocks || n_blocks == 0 || !out_handle) return c->fail(HJK_ERR_INVALID_ARGUMENT, "empty block list");
  HJK_CUDA(c, cudaSetDevice(c->device));
  HjkContext::Resident r;
  r.host.assign(blocks, blocks + n_blocks);
  HJK_CUDA(c, cudaMalloc((void**)&r.dev, n_blocks * sizeof(HjkImageBlock)));
  cudaError_t e = cudaMemcpy(r.dev, blocks, n_blocks * sizeof(HjkImageBlock), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(r.dev);
    return c->fail(HJK_ERR_CUDA, "block upload failed: %s", cudaGetErrorString(e));
  }
  const uint64_t h = c->next_handle++;
  c->resident[h] = std::move(r);
  *out_handle = h;
  return HJK_OK;
}

int hjk_render_resident(HjkContext* c, uint64_t handle, uint64_t first_block, uint64_t n_blocks,
                        const HjkParams* prm, HjkStats* stats) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  auto it = c->resident.find(handle);
  if (it == c->resident.end()) return c->fail(HJK_ERR_INVALID_ARGUMENT, "unknown block-list handle");
  if (n_blocks == 0 || first_block + n_blocks > it->second.host.size())
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "block range out of bounds");
  if (c->members.size() > 1)  // the resident list lives on the first device only: split it like a host list
    return render_group(c, it->second.host.data() + first_block, n_blocks, prm, stats);
  HJK_CUDA(c, cudaSetDevice(c->device));
  return render_blocks(c, it->second.host.data() + first_block, it->second.dev + first_block, n_blocks, prm, stats);
}

int hjk_blocks_free(HjkContext* c, uint64_t handle) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  auto it = c->resident.find(handle);
  if (it == c->resident.end()) return c->fail(HJK_ERR_INVALID_ARGUMENT, "unknown block-list handle");
  cudaStreamSynchronize(c->stream);
  cudaFree(it->second.dev);
  c->resident.erase(it);
  return HJK_OK;
}

// Sum of the frame (accumulator, and the feature sums + sample counts when feature_buffers is on: one
// contiguous run of floats) over the ranks of the communicator, into d_sum — the rank's own d_frame is left as
// it is, so a frame can be reduced and read back any number of times and rendered into afterwards.
// root < 0: every rank receives the sum (ncclAllReduce); else only `root` does (ncclReduce).
static int reduce_frame(HjkContext* c, int root, float* out_ms) {
  if (out_ms) *out_ms = 0.f;
  if (!c->d_frame.p) return c->fail(HJK_ERR_NO_FRAME, "no frame");
  if (!c->comm || c->n_ranks == 1) return HJK_OK;
  if (root >= c->n_ranks) return c->fail(HJK_ERR_INVALID_ARGUMENT, "root rank out of range");
  // Never skipped, even when this rank's frame has not changed since the last reduction: whether to run a
  // collective cannot be decided from one rank's state (another rank may have rendered since), and the sum goes
  // to d_sum, so repeating it is harmless.
  HJK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = c->frame_floats();
  HJK_CUDA(c, c->d_sum.ensure(9 * (size_t)c->width * c->height));
  HJK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  int r;
  if (root < 0)
    r = g_nccl.AllReduce(c->d_frame.p, c->d_sum.p, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->comm, c->stream);
  else
    r = g_nccl.Reduce(c->d_frame.p, c->d_sum.p, n, 7, 0, root, c->comm, c->stream);
  if (r != 0)
    return c->fail(HJK_ERR_NCCL, "%s: %s", root < 0 ? "ncclAllReduce" : "ncclReduce",
                   g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  HJK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  if (out_ms) {
    HJK_CUDA(c, cudaEventSynchronize(c->ev1));
    cudaEventElapsedTime(out_ms, c->ev0, c->ev1);
  }
  c->reduced_valid = true;
  c->reduced_root = root;
  return HJK_OK;
}

// The devices of a single-process group: one ncclReduce per member inside one NCCL group, all to member 0.
static int reduce_group(HjkContext* c, float* out_ms) {
  if (out_ms) *out_ms = 0.f;
  if (c->reduced_valid) return HJK_OK;
  const size_t n = c->frame_floats();
  HJK_CUDA(c, cudaSetDevice(c->device));
  HJK_CUDA(c, c->d_sum.ensure(9 * (size_t)c->width * c->height));
  for (HjkContext* m : c->members)
    if (!m->d_frame.p || m->width != c->width || m->height != c->height)
      return c->fail(HJK_ERR_NO_FRAME, "device %d has no frame of this size", m->device);
  HJK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  int r = g_nccl.GroupStart();
  for (HjkContext* m : c->members) {
    if (r != 0) break;
    cudaSetDevice(m->device);
    r = g_nccl.Reduce(m->d_frame.p, m == c ? c->d_sum.p : m->d_frame.p, n, 7, 0, 0, m->comm, m->stream);
  }
  const int r2 = g_nccl.GroupEnd();
  cudaSetDevice(c->device);
  if (r != 0 || r2 != 0)
    return c->fail(HJK_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r ? r : r2) : "error");
  HJK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  if (out_ms) {
    HJK_CUDA(c, cudaEventSynchronize(c->ev1));
    cudaEventElapsedTime(out_ms, c->ev0, c->ev1);
  }
  c->reduced_valid = true;
  c->reduced_root = 0;
  return HJK_OK;
}

// what a readback on this context copies from: the reduced frame when there is more than one rank / device
static const float* frame_source(const HjkContext* c) {
  return (c->members.size() > 1 || (c->comm && c->n_ranks > 1)) ? c->d_sum.p : c->d_frame.p;
}

int hjk_reduce_frame(HjkContext* c, int root, float* out_ms) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (c->members.size() > 1) return reduce_group(c, out_ms);
  return reduce_frame(c, root, out_ms);
}

int hjk_allreduce_accumulator(HjkContext* c, float* out_ms) { return hjk_reduce_frame(c, -1, out_ms); }

// Waits (on the host) for the copy of an earlier hjk_readback_begin.
static int wait_pending_copy(HjkContext* c) {
  if (c->copy_pending) {
    HJK_CUDA(c, cudaEventSynchronize(c->ev_copied));
    c->copy_pending = false;
  }
  return HJK_OK;
}

// async: the frame is staged in d_norm (normalised, or copied as it is) on the render stream and copied to the host
// on the copy stream, so the render stream is free for the next frame at once; hjk_readback_wait completes it.
static int copy_frame_out(HjkContext* c, float* rgba, uint64_t pitch_bytes, int normalise, bool async) {
  const uint32_t n = c->width * c->height;
  const f4* src = (const f4*)frame_source(c);
  int rc = wait_pending_copy(c);  // d_norm is about to be rewritten
  if (rc) return rc;
  if (normalise) {
    HJK_CUDA(c, c->d_norm.ensure(n));
    k_normalise<<<grid_for(c, 4), 256, 0, c->stream>>>(src, c->d_norm.p, n);
    HJK_CUDA(c, cudaGetLastError());
    src = c->d_norm.p;
  } else if (async) {
    HJK_CUDA(c, c->d_norm.ensure(n));
    HJK_CUDA(c, cudaMemcpyAsync(c->d_norm.p, src, (size_t)n * sizeof(f4), cudaMemcpyDeviceToDevice, c->stream));
    src = c->d_norm.p;
  }
  if (!async) {
    HJK_CUDA(c, cudaMemcpy2DAsync(rgba, pitch_bytes, src, (size_t)c->width * 16, (size_t)c->width * 16, c->height,
                                  cudaMemcpyDeviceToHost, c->stream));
    HJK_CUDA(c, cudaStreamSynchronize(c->stream));
    return HJK_OK;
  }
  if (!c->copy_stream) {
    HJK_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    HJK_CUDA(c, cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
    HJK_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
  }
  HJK_CUDA(c, cudaEventRecord(c->ev_staged, c->stream));
  HJK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_staged, 0));
  HJK_CUDA(c, cudaMemcpy2DAsync(rgba, pitch_bytes, src, (size_t)c->width * 16, (size_t)c->width * 16, c->height,
                                cudaMemcpyDeviceToHost, c->copy_stream));
  HJK_CUDA(c, cudaEventRecord(c->ev_copied, c->copy_stream));
  c->copy_pending = true;
  return HJK_OK;
}

static int readback_impl(HjkContext* c, int root, float* rgba, uint64_t pitch_bytes, int normalise, bool async) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!c->d_frame.p) return c->fail(HJK_ERR_NO_FRAME, "hjk_readback before any frame");
  const bool multi_rank = c->comm && c->n_ranks > 1 && c->members.size() <= 1;
  const bool receives = !multi_rank || root < 0 || root == c->rank;
  if (receives && (!rgba || pitch_bytes < (uint64_t)c->width * 16))
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "bad destination");
  HJK_CUDA(c, cudaSetDevice(c->device));
  int rc = hjk_reduce_frame(c, root, nullptr);
  if (rc) return rc;
  if (!receives) return HJK_OK;  // the collective is enqueued; this rank's host gets nothing
  return copy_frame_out(c, rgba, pitch_bytes, normalise, async);
}

int hjk_readback_root(HjkContext* c, int root, float* rgba, uint64_t pitch_bytes, int normalise) {
  return readback_impl(c, root, rgba, pitch_bytes, normalise, false);
}

int hjk_readback_begin(HjkContext* c, int root, float* rgba, uint64_t pitch_bytes, int normalise) {
  return readback_impl(c, root, rgba, pitch_bytes, normalise, true);
}

int hjk_readback_wait(HjkContext* c) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  HJK_CUDA(c, cudaSetDevice(c->device));
  return wait_pending_copy(c);
}

int hjk_readback(HjkContext* c, float* rgba, uint64_t pitch_bytes, int normalise) {
  return hjk_readback_root(c, -1, rgba, pitch_bytes, normalise);
}

int hjk_read_features(HjkContext* c, int root, float* normal_depth, uint64_t pitch_bytes) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!c->d_frame.p) return c->fail(HJK_ERR_NO_FRAME, "hjk_read_features before any frame");
  if (!c->feature_buffers)
    return c->fail(HJK_ERR_UNSUPPORTED, "set the option \"feature_buffers\" to 1 before the frame is rendered");
  const bool multi_rank = c->comm && c->n_ranks > 1 && c->members.size() <= 1;
  const bool receives = !multi_rank || root < 0 || root == c->rank;
  if (receives && (!normal_depth || pitch_bytes < (uint64_t)c->width * 16))
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "bad destination");
  HJK_CUDA(c, cudaSetDevice(c->device));
  int rc = hjk_reduce_frame(c, root, nullptr);
  if (rc) return rc;
  if (!receives) return HJK_OK;
  const uint32_t n = c->width * c->height;
  const float* base = frame_source(c);
  HJK_CUDA(c, c->d_norm.ensure(n));
  k_normalise_features<<<grid_for(c, 4), 256, 0, c->stream>>>((const f4*)(base + 4 * (size_t)n), base + 8 * (size_t)n,
                                                              c->d_norm.p, n);
  HJK_CUDA(c, cudaGetLastError());
  HJK_CUDA(c, cudaMemcpy2DAsync(normal_depth, pitch_bytes, c->d_norm.p, (size_t)c->width * 16, (size_t)c->width * 16,
                                c->height, cudaMemcpyDeviceToHost, c->stream));
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));
  return HJK_OK;
}

int hjk_read_intermediate(HjkContext* c, int layer, float* rgba) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!c->have_features || !c->last_wave_passes) return c->fail(HJK_ERR_NO_FRAME, "no pass has been rendered");
  if (layer < 0 || layer > 2 || !rgba) return c->fail(HJK_ERR_INVALID_ARGUMENT, "layer must be 0, 1 or 2");
  HJK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)c->width * c->height;
  if (layer == 2) {  // albedo is never written by the integrator (render.glsl:84-85,174)
    memset(rgba, 0, n * 16);
    return HJK_OK;
  }
  const f4* src = (layer == 0 ? c->d_layer0.p : c->d_layer1.p) + (size_t)(c->last_wave_passes - 1) * n;
  HJK_CUDA(c, cudaMemcpyAsync(rgba, src, n * 16, cudaMemcpyDeviceToHost, c->stream));
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));
  return HJK_OK;
}

int hjk_trace_first_hit_eps(HjkContext* c, const HjkRay* rays, uint64_t n_rays, int any_hit, float eps,
                            int32_t* shape_id, float* t, float* uv) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!c->has_scene) return c->fail(HJK_ERR_NO_SCENE, "hjk_trace_first_hit before hjk_scene_upload");
  if (!rays || !shape_id || n_rays == 0 || n_rays > 0x7FFFFFFFull)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "bad ray batch");
  if (!(eps >= 0.f)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "eps must be non-negative");
  HJK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)n_rays;
  std::vector<f4>& ho = c->h_batch_o;
  std::vector<f4>& hd = c->h_batch_d;
  ho.resize(n), hd.resize(n);
  for (size_t i = 0; i < n; i++) {
    ho[i] = F4(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2], rays[i].t_min);
    hd[i] = F4(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2], rays[i].t_max);
  }
  // the batch buffers belong to the context and only grow: repeated calls do not allocate
  DevBuf<f4>&d_o = c->d_batch_o, &d_d = c->d_batch_d, &d_h = c->d_batch_h;
  DevBuf<uint32_t>& d_cur = c->d_batch_cur;
  HJK_CUDA(c, d_o.ensure(n));
  HJK_CUDA(c, d_d.ensure(n));
  HJK_CUDA(c, d_h.ensure(n));
  HJK_CUDA(c, d_cur.ensure(2));  // work cursor, unresolved tie clusters
  HJK_CUDA(c, cudaMemcpyAsync(d_o.p, ho.data(), n * 16, cudaMemcpyHostToDevice, c->stream));
  HJK_CUDA(c, cudaMemcpyAsync(d_d.p, hd.data(), n * 16, cudaMemcpyHostToDevice, c->stream));
  HJK_CUDA(c, cudaMemsetAsync(d_cur.p, 0, 8, c->stream));
  const int g = grid_for(c, c->blocks_batch);
  const bool guard = c->scene.num_spheres != 0;
  const uint32_t flavour = (any_hit & 1) ? kAnyHitBit : 0u;
  const bool exact = (any_hit & 2) != 0;
  const int postpone = c->lane_postpones ? kPostponeLanes : 0;
  const size_t sm_trav = (size_t)c->stack_cap_lane * kTravThreads * sizeof(uint2);
#define HJK_BATCH(G, E) \
  k_trace_batch<G, E><<<g, kTravThreads, sm_trav, c->stream>>>(c->scene, d_o.p, d_d.p, d_h.p, (uint32_t)n, d_cur.p, eps, \
                                                               flavour, postpone, (int)c->stack_cap_lane)
  if (guard && exact)
    HJK_BATCH(true, true);
  else if (guard)
    HJK_BATCH(true, false);
  else if (exact)
    HJK_BATCH(false, true);
  else
    HJK_BATCH(false, false);
#undef HJK_BATCH
  HJK_CUDA(c, cudaGetLastError());
  std::vector<f4>& hh = ho;  // the origins are on the device by now (same stream): reuse as the result buffer
  HJK_CUDA(c, cudaMemcpyAsync(hh.data(), d_h.p, n * 16, cudaMemcpyDeviceToHost, c->stream));
  uint32_t cur2[2] = {0, 0};
  HJK_CUDA(c, cudaMemcpyAsync(cur2, d_cur.p, 8, cudaMemcpyDeviceToHost, c->stream));
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->unresolved_last = cur2[1];
  for (size_t i = 0; i < n; i++) {
    int32_t id;
    memcpy(&id, &hh[i].x, 4);
    if (any_hit & 1) {
      shape_id[i] = id >= 0 ? 1 : 0;
    } else {
      shape_id[i] = id;
    }
    if (t) t[i] = id >= 0 ? hh[i].y : 0.f;
    if (uv) {
      uv[2 * i] = id >= 0 ? hh[i].z : 0.f;
      uv[2 * i + 1] = id >= 0 ? hh[i].w : 0.f;
    }
  }
  return HJK_OK;
}

int hjk_trace_first_hit(HjkContext* c, const HjkRay* rays, uint64_t n_rays, int any_hit, int32_t* shape_id,
                        float* t, float* uv) {
  return hjk_trace_first_hit_eps(c, rays, n_rays, any_hit, 1e-4f /* M_EPS, math.glsl:2 */, shape_id, t, uv);
}

static int denoise_apply(HjkContext* c, const HjkParams* prm, uint32_t repeat, float* out_ms) {
  PassPlan plan;
  std::string err;
  if (!plan_passes(c->dn_blocks.data(), c->dn_blocks.size(), plan, err))
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "%s", err.c_str());
  if (plan.passes.size() != 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "denoise blocks must form exactly one pass");
  if (!(prm->recon_stddev > 0.f)) return c->fail(HJK_ERR_INVALID_ARGUMENT, "recon_stddev must be positive");
  int rc = ensure_frame(c, plan.width, plan.height, false);
  if (rc) return rc;
  const int R = (int)prm->recon_radius, taps = 2 * R + 1;
  const size_t nb = c->dn_blocks.size();
  HJK_CUDA(c, c->d_blocks.ensure(nb));
  HJK_CUDA(c, c->d_weights.ensure(nb * taps * taps));
  HJK_CUDA(c, c->d_taps.ensure(nb * recon_tap_stride(R)));
  HJK_CUDA(c, c->d_tile_block.ensure(plan.tile_block.size()));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_blocks.p, c->dn_blocks.data(), nb * sizeof(HjkImageBlock), cudaMemcpyHostToDevice,
                              c->stream));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_tile_block.p, plan.tile_block.data(), plan.tile_block.size() * 4,
                              cudaMemcpyHostToDevice, c->stream));
  k_recon_weights<<<std::max<int>(1, (int)std::min<size_t>(1024, (nb + 127) / 128)), 128, 0, c->stream>>>(
      c->d_blocks.p, (uint32_t)nb, R, prm->recon_stddev, c->d_weights.p, c->d_taps.p);
  PassDev ps{};
  ps.width = plan.width, ps.height = plan.height, ps.tile_w = plan.tile_w, ps.tile_h = plan.tile_h;
  ps.tiles_x = plan.tiles_x, ps.tiles_y = plan.tiles_y;
  ps.tile_block = c->d_tile_block.p;
  ps.blocks = c->d_blocks.p;
  ps.weights = c->d_weights.p;
  ps.taps = c->d_taps.p;
  ps.radius = R;
  HJK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  for (uint32_t i = 0; i < repeat; i++) {
    rc = launch_recon(c, ps, 1, c->d_dn0.p, c->d_dn1.p, c->dn_has_albedo ? c->d_dn2.p : nullptr, c->acc(), false);
    if (rc) return rc;
  }
  HJK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  HJK_CUDA(c, cudaEventSynchronize(c->ev1));
  if (out_ms) cudaEventElapsedTime(out_ms, c->ev0, c->ev1);
  return HJK_OK;
}

static int denoise_upload(HjkContext* c, const float* radiance, const float* normal_depth, const float* albedo,
                          const HjkImageBlock* blocks, uint64_t n_blocks) {
  if (!radiance || !normal_depth || !blocks || n_blocks == 0)
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "denoise needs radiance, normal_depth and blocks");
  HJK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)blocks[0].original_dimension[0] * blocks[0].original_dimension[1];
  if (n == 0 || n > 0x7FFFFFFFull) return c->fail(HJK_ERR_INVALID_ARGUMENT, "unsupported image size");
  HJK_CUDA(c, c->d_dn0.ensure(n));
  HJK_CUDA(c, c->d_dn1.ensure(n));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_dn0.p, radiance, n * 16, cudaMemcpyHostToDevice, c->stream));
  HJK_CUDA(c, cudaMemcpyAsync(c->d_dn1.p, normal_depth, n * 16, cudaMemcpyHostToDevice, c->stream));
  c->dn_has_albedo = albedo != nullptr;
  if (albedo) {
    HJK_CUDA(c, c->d_dn2.ensure(n));
    HJK_CUDA(c, cudaMemcpyAsync(c->d_dn2.p, albedo, n * 16, cudaMemcpyHostToDevice, c->stream));
  }
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->dn_blocks.assign(blocks, blocks + n_blocks);
  return HJK_OK;
}

int hjk_denoise_pass(HjkContext* c, const float* radiance, const float* normal_depth, const float* albedo,
                     const HjkImageBlock* blocks, uint64_t n_blocks, const HjkParams* prm) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!prm) return c->fail(HJK_ERR_INVALID_ARGUMENT, "params is null");
  int rc = denoise_upload(c, radiance, normal_depth, albedo, blocks, n_blocks);
  if (rc) return rc;
  return denoise_apply(c, prm, 1, nullptr);
}

int hjk_denoise_upload(HjkContext* c, const float* radiance, const float* normal_depth, const HjkImageBlock* blocks,
                       uint64_t n_blocks) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  return denoise_upload(c, radiance, normal_depth, nullptr, blocks, n_blocks);
}

int hjk_denoise_resident(HjkContext* c, const HjkParams* prm, uint32_t repeat, float* out_ms) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!prm || repeat == 0) return c->fail(HJK_ERR_INVALID_ARGUMENT, "bad arguments");
  if (c->dn_blocks.empty()) return c->fail(HJK_ERR_NO_FRAME, "hjk_denoise_resident before hjk_denoise_upload");
  HJK_CUDA(c, cudaSetDevice(c->device));
  return denoise_apply(c, prm, repeat, out_ms);
}

int hjk_comm_unique_id(void* out_id128) {
  std::string err;
  if (!out_id128 || !load_nccl(err)) {
    g_create_error = err.empty() ? "null id buffer" : err;
    return HJK_ERR_NCCL;
  }
  int r = g_nccl.GetUniqueId(out_id128);
  if (r != 0) {
    g_create_error = "ncclGetUniqueId failed";
    return HJK_ERR_NCCL;
  }
  return HJK_OK;
}

int hjk_comm_init(HjkContext* c, const void* id128, int rank, int n_ranks) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return c->fail(HJK_ERR_INVALID_ARGUMENT, "bad rank/size");
  if (c->comm) return c->fail(HJK_ERR_INVALID_ARGUMENT, "the context already has a communicator");
  std::string err;
  if (!load_nccl(err)) return c->fail(HJK_ERR_NCCL, "%s", err.c_str());
  HJK_CUDA(c, cudaSetDevice(c->device));
  Id128 id;
  memcpy(id.bytes, id128, 128);
  int r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != 0) return c->fail(HJK_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  c->rank = rank;
  c->n_ranks = n_ranks;
  return HJK_OK;
}

int hjk_accumulator_device_ptr(HjkContext* c, uint64_t* out_ptr, uint64_t* out_n_floats) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  if (!c->d_frame.p) return c->fail(HJK_ERR_NO_FRAME, "no frame");
  if (out_ptr) *out_ptr = (uint64_t)(uintptr_t)c->acc();
  if (out_n_floats) *out_n_floats = (uint64_t)c->width * c->height * 4;
  return HJK_OK;
}

int hjk_synchronize(HjkContext* c) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  HJK_CUDA(c, cudaSetDevice(c->device));
  HJK_CUDA(c, cudaStreamSynchronize(c->stream));
  return HJK_OK;
}

int hjk_set_profiling(HjkContext* c, int enabled) {
  if (!c) return HJK_ERR_INVALID_ARGUMENT;
  c->profiling = enabled != 0;
  return HJK_OK;
}

int hjk_set_option(HjkContext* c, const char* key, int64_t value) {
  if (!c || !key) return HJK_ERR_INVALID_ARGUMENT;
  for (size_t i = 1; i < c->members.size(); i++) {  // a group's options apply to every device
    const int rc = hjk_set_option(c->members[i], key, value);
    if (rc) return c->fail(rc, "%s", c->members[i]->error.c_str());
  }
  const std::string k(key);
  if (k == "wave_paths") {
    if (value < 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "wave_paths must be positive");
    c->wave_paths = (uint64_t)value;
  } else if (k == "bvh_pad_rel_e9") {  // relative primitive-box pad in units of 1e-9 (next scene upload)
    if (value < 0) return c->fail(HJK_ERR_INVALID_ARGUMENT, "pad must be non-negative");
    c->bvh_pad_rel = (float)value * 1e-9f;
  } else if (k == "coop_trace") {
    c->coop_trace = value != 0;
  } else if (k == "coop_batch_cost") {
    if (value < -1 || value > 100000) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->coop_batch_cost = (int)value;
  } else if (k == "bvh_builder") {  // 0 host SAH (default), 1 GPU, -1 by scene size; takes effect at the next scene upload
    if (value < -1 || value > 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->bvh_builder = (int)value;
  } else if (k == "bvh_gpu_collapse") {  // 1 = the host builder's collapse-cost dynamic programme (default), 0 = greedy
    c->bvh_gpu_collapse = value != 0;
  } else if (k == "bvh_ploc_radius") {
    if (value < 1 || value > kPlocMaxRadius) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->bvh_ploc_radius = (int)value;
  } else if (k == "bvh_gpu_tree") {  // the GPU builder's binary tree: 1 PLOC (default), 0 radix tree
    if (value < 0 || value > 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->bvh_gpu_tree = (int)value;
  } else if (k == "shade_sort") {  // -1 = decided per scene (default), 0 = never, 1 = always sort a tile's hits by material
    if (value < -1 || value > 1) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->shade_sort = (int)value;
  } else if (k == "bvh_validate") {
    c->bvh_validate = value != 0;
  } else if (k == "bvh_broadcast") {  // several ranks: 1 = rank 0 builds and broadcasts the wide BVH (default)
    c->bvh_broadcast = value != 0;
  } else if (k == "feature_buffers") {  // sum the first-hit (normal, depth) per texel; takes effect at hjk_frame_begin
    c->feature_buffers = value != 0;
    c->reduced_valid = false;
  } else if (k == "fetch_threshold") {
    if (value < -1 || value > 32) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->fetch_threshold = (int)value;
  } else if (k == "postpone_lanes") {
    if (value < 0 || value > 32) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->postpone_lanes = (uint32_t)value;
  } else if (k == "blocks_per_sm_traverse") {  // 0 = the occupancy of each trace kernel
    if (value < 0 || value > 32) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->blocks_trav_override = (int)value;
  } else if (k == "blocks_per_sm_tile") {
    if (value < 1 || value > 32) return c->fail(HJK_ERR_INVALID_ARGUMENT, "out of range");
    c->blocks_tile = (int)value;
  } else {
    return c->fail(HJK_ERR_INVALID_ARGUMENT, "unknown option '%s'", key);
  }
  return HJK_OK;
}

int hjk_get_info(HjkContext* c, const char* key, int64_t* out) {
  if (!c || !key || !out) return HJK_ERR_INVALID_ARGUMENT;
  const std::string k(key);
  if (k == "n_sms") *out = c->n_sms;
  else if (k == "bvh_nodes") *out = (int64_t)c->n_nodes;
  else if (k == "bvh_prims") *out = (int64_t)c->n_prims;
  else if (k == "bvh_depth") *out = c->bvh_host_stats.depth;
  else if (k == "bvh_bytes") *out = (int64_t)(c->n_nodes * sizeof(WideNode) + c->n_prims * sizeof(WidePrim));
  else if (k == "blocks_per_sm_traverse") *out = c->blocks_trav_override ? c->blocks_trav_override : c->blocks_trav;
  else if (k == "blocks_per_sm_tile") *out = c->blocks_tile;
  else if (k == "wave_paths") *out = (int64_t)c->wave_paths;
  else if (k == "has_extinction") *out = c->has_extinction ? 1 : 0;
  else if (k == "shade_sort") *out = c->shade_sort < 0 ? (c->scene_mixes_materials ? 1 : 0) : c->shade_sort;
  else if (k == "sphere_guard") *out = c->scene.num_spheres ? (c->bvh_all_guarded ? 2 : 1) : 0;
  else if (k == "bvh_builder") *out = c->bvh_builder;
  else if (k == "bvh_build_us") *out = (int64_t)(c->bvh_build_ms * 1000.f);
  else if (k == "unresolved_ties") *out = (int64_t)c->unresolved_last;
  else if (k == "stack_overflows") {  // entries the traversal stack dropped on this device since the last scene upload
    unsigned int v = 0;
    HJK_CUDA(c, cudaSetDevice(c->device));
    HJK_CUDA(c, cudaStreamSynchronize(c->stream));
    HJK_CUDA(c, cudaMemcpyFromSymbol(&v, g_stack_overflows, sizeof v));
    *out = (int64_t)v;
  }
  else if (k == "device") *out = c->device;
  else if (k == "n_devices") *out = c->members.empty() ? 1 : (int64_t)c->members.size();
  else if (k == "rank") *out = c->rank;
  else if (k == "n_ranks") *out = c->n_ranks;
  else if (k == "feature_buffers") *out = c->feature_buffers ? 1 : 0;
  else if (k == "width") *out = c->width;
  else if (k == "height") *out = c->height;
  else return c->fail(HJK_ERR_INVALID_ARGUMENT, "unknown info key '%s'", key);
  return HJK_OK;
}

}  // extern "C"
