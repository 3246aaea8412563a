/*
 * hijiki_b200.h — C ABI of the B200-native Hijiki path-tracing hot path.
 *
 * This is the drop-in boundary for the wgpu/GLSL render loop of mad-s/hijiki.
 * The reference has no FFI; the seam this header replaces is internal to
 * reference `src/main.rs`:
 *
 *   IntegratorPipeline::{new,run}       src/main.rs:760-898   (render.glsl dispatch per block)
 *   ReconstructionPipeline::{new,run}   src/main.rs:907-1004  (reconstruction.glsl dispatch per block)
 *   Renderer::{new,render,save_image}   src/main.rs:1167-1423 (scene upload, block loop, readback)
 *
 * Everything crossing the boundary is plain pointers + sizes in the byte
 * layouts the reference hands to its shaders.  No C++/torch
 * types appear in any signature.  Every call returns an int: 0 = ok, <0 = an
 * HjkStatus error; nothing throws or aborts across the ABI.
 *
 * A context is bound to ONE host thread at a time (like the reference's single
 * queue) and owns all device memory, streams and (optional) NCCL state.
 */
#ifndef HIJIKI_B200_H
#define HIJIKI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HJK_API __attribute__((visibility("default")))
#else
#define HJK_API
#endif

/* ------------------------------------------------------------------ status */
typedef enum HjkStatus {
  HJK_OK = 0,
  HJK_ERR_INVALID_ARGUMENT = -1,
  HJK_ERR_CUDA = -2,          /* a CUDA runtime call failed; see hjk_last_error */
  HJK_ERR_NO_SCENE = -3,      /* render/trace before hjk_scene_upload */
  HJK_ERR_NO_FRAME = -4,      /* readback before any frame was allocated */
  HJK_ERR_UNSUPPORTED = -5,
  HJK_ERR_NCCL = -6,
  HJK_ERR_IO = -7,
  HJK_ERR_OUT_OF_MEMORY = -8
} HjkStatus;

/* ---------------------------------------------- POD layouts */

/* reference Camera, src/main.rs:154-160; render.glsl:12-16 (std430: 48 B) */
typedef struct HjkCamera {
  float position[4];
  float rotation[4]; /* quaternion xyzw */
  float fov;         /* horizontal, degrees */
  float _pad[3];
} HjkCamera;

/* reference SceneBufferInfo, src/main.rs:400-408; scene.glsl:1-8 (64 B) */
typedef struct HjkSceneInfo {
  HjkCamera camera;
  uint32_t num_spheres;
  uint32_t num_quads;
  uint32_t num_triangles;
  uint32_t num_emitters;
} HjkSceneInfo;

/* reference CompiledBVHNode, src/main.rs:92-99; scene.glsl:10-17 (32 B).
 * Preorder, skip-pointer ("threaded") binary BVH.  shape_index 0xFFFFFFFF = inner. */
typedef struct HjkBvh2Node {
  float aabb_min[3];
  uint32_t shape_index;
  float aabb_max[3];
  uint32_t exit_index;
} HjkBvh2Node;

/* reference Sphere, src/shape.rs:6-11 (16 B) */
typedef struct HjkSphere {
  float position[3];
  float radius;
} HjkSphere;

/* reference Quad, src/shape.rs:22-31 (48 B) */
typedef struct HjkQuad {
  float origin[3], _pad1;
  float edge1[3], _pad2;
  float edge2[3], _pad3;
} HjkQuad;

/* reference Vertex as the shader reads it, shapes/triangle.glsl:1-4 (32 B) */
typedef struct HjkVertex {
  float pos[3];
  float u;
  float normal[3];
  float v;
} HjkVertex;

/* reference Emitter, src/main.rs:368-374; scene.glsl:33-38 (16 B) */
typedef struct HjkEmitter {
  uint32_t shape;
  float pdf;
  float cdf;
  float pad;
} HjkEmitter;

/* reference DiffuseMaterial / EmitterMaterial (vec3 padded to 16 B), src/main.rs:102-106,142-146 */
typedef struct HjkColor16 {
  float rgb[3];
  float _pad;
} HjkColor16;

/* reference DiffuseCheckerboardMaterial, src/main.rs:108-115 (32 B) */
typedef struct HjkDiffuseCB {
  float color1[3];
  float scale_u;
  float color2[3];
  float scale_v;
} HjkDiffuseCB;

/* reference DielectricMaterial, src/main.rs:122-126 (16 B): xyz extinction, w eta ratio */
typedef struct HjkDielectric {
  float extinction_eta[4];
} HjkDielectric;

/* reference ImageBlock, src/main.rs:608-617; block.glsl:1-8 (40 B) */
typedef struct HjkImageBlock {
  uint32_t id;
  uint32_t seed;
  uint32_t origin[2];
  uint32_t dimension[2];
  uint32_t original_dimension[2];
  float sample_offset[2];
} HjkImageBlock;

/* reference Ray, render.glsl:19-24 (std430: vec3 origin, vec3 direction interleaved
 * with the two scalars -> 32 B) */
typedef struct HjkRay {
  float origin[3];
  float t_min;
  float direction[3];
  float t_max;
} HjkRay;

/* material word = (tag << 24) | index, src/main.rs:34-45,275 */
enum {
  HJK_MATERIAL_TAG_SHIFT = 24,
  HJK_MAT_DIFFUSE = 0,
  HJK_MAT_DIFFUSECBOARD = 1,
  HJK_MAT_MIRROR = 2,
  HJK_MAT_DIELECTRIC = 3,
  HJK_MAT_EMISSIVE = 4
};

typedef struct HjkArray {
  const void* ptr; /* host pointer; the library copies, caller keeps ownership */
  uint64_t count;  /* number of ELEMENTS (not bytes) */
} HjkArray;

/* The 12 CompiledScene arrays in binding order (src/main.rs:314-327,561-605).
 * Global shape index space: [0,S) spheres, [S,S+Q) quads, [S+Q,S+Q+T) triangles. */
typedef struct HjkScene {
  HjkArray scene;      /* 1 x HjkSceneInfo */
  HjkArray bvh;        /* HjkBvh2Node[], OPTIONAL (the library builds its own wide BVH) */
  HjkArray spheres;    /* HjkSphere[] */
  HjkArray quads;      /* HjkQuad[] */
  HjkArray triangles;  /* uint32[3] per element: global vertex indices */
  HjkArray vertices;   /* HjkVertex[] */
  HjkArray materials;  /* uint32 per shape in global shape order */
  HjkArray emitters;   /* HjkEmitter[] */
  HjkArray diffuse;    /* HjkColor16[] */
  HjkArray diffusecb;  /* HjkDiffuseCB[] */
  HjkArray dielectric; /* HjkDielectric[] */
  HjkArray emissive;   /* HjkColor16[] */
} HjkScene;

/* Constants the reference injects as shader macros or hard-codes
 * (src/main.rs:769-783,916-922,1284-1285; render.glsl:92,137; math.glsl:2). */
typedef struct HjkParams {
  uint32_t max_bounces;  /* reference: 1000 (render.glsl:92) */
  uint32_t rr_start;     /* Russian roulette when bounce > rr_start; reference: 3 */
  uint32_t recon_radius; /* reference: 2 */
  float recon_stddev;    /* reference: 0.5 */
  float eps;             /* M_EPS; reference: 1e-4 */
  uint32_t flags;        /* HJK_RENDER_* */
} HjkParams;

enum {
  HJK_RENDER_ASYNC = 1u << 0,       /* return after enqueue; default waits for the device */
  HJK_RENDER_NO_RECON = 1u << 1,    /* integrate only (debug / feature export) */
  HJK_RENDER_KEEP_FEATURES = 1u << 2, /* keep the last pass' intermediate layers for hjk_read_intermediate */
  HJK_RENDER_EXACT_TIES = 1u << 3     /* resolve hits closer than eps to each other exactly as the reference's
                                         linear scan does (scene.glsl:134-157).  The default reports the nearest
                                         member of such a cluster (equal t: the lower shape id) — deterministic,
                                         but not always the one the reference's scan order picks.  Slower; rays
                                         whose cluster cannot be resolved are counted in
                                         hjk_get_info("unresolved_ties") */
};

enum { HJK_N_KERNEL_SLOTS = 8 };
enum {
  HJK_K_RAYGEN = 0,
  HJK_K_EXTEND = 1,
  HJK_K_SHADE = 2,
  HJK_K_SHADOW = 3, /* unused since shadow rays are traced by the same launches as extension rays
                       (their time is in HJK_K_EXTEND) */
  HJK_K_RECON = 4,
  HJK_K_OTHER = 5,
  HJK_K_SORT = 6 /* unused since the material sort is tile-local inside the shade kernel */
};

typedef struct HjkStats {
  uint64_t n_paths;          /* camera paths started (= pixels x passes) */
  uint64_t n_extension_rays; /* closest-hit intersectScene calls, render.glsl:94 */
  uint64_t n_shadow_rays;    /* shadow intersectScene calls, render.glsl:122 */
  float ms_total;            /* device time of the whole call (CUDA events) */
  float kernel_ms[HJK_N_KERNEL_SLOTS]; /* per-stage device time, only when profiling is on */
  uint64_t n_launches;       /* kernels launched by this call */
} HjkStats;

typedef struct HjkContext HjkContext;

/* --------------------------------------------------------- context (GPU) */

/* Replaces GPU::new (src/main.rs:692-712).  n_devices == 1: one GPU per context; multi-GPU runs then use one
 * context per process joined by hjk_comm_init.  n_devices > 1: a single-process GROUP over the listed devices
 * (ncclCommInitAll): hjk_scene_upload builds the wide BVH once and copies it to every device, hjk_render sends
 * sample pass p of the list to device p mod n_devices (one host thread per device), hjk_readback sums the
 * frames onto the first device (one ncclReduce per frame) and copies from there.  hjk_trace_first_hit, the
 * standalone denoise entries and hjk_read_intermediate of a group run on its first device. */
HJK_API int hjk_create(const int* device_ids, int n_devices, HjkContext** out_ctx);
HJK_API int hjk_destroy(HjkContext* ctx);

/* Message for the last failing call on this context (ctx may be NULL for
 * creation failures).  Replaces panic!/unwrap (src/main.rs:415,434,697,705). */
HJK_API const char* hjk_last_error(const HjkContext* ctx);

/* Replaces the scene staging copy (src/main.rs:1187-1195,1238-1244) and builds
 * the 8-wide compressed BVH that stands in for scene.glsl's USE_BVH walk. */
HJK_API int hjk_scene_upload(HjkContext* ctx, const HjkScene* scene);

/* (Re)allocate and zero the full-frame RGBA32F accumulator
 * (final_texture, src/main.rs:1211-1267). */
HJK_API int hjk_frame_begin(HjkContext* ctx, uint32_t width, uint32_t height);

/* Replaces Renderer::render (src/main.rs:1316-1355): integrates every block of
 * `blocks` (HOST pointer, id order) and splats it into the accumulator with the
 * bilateral reconstruction filter.  ADDS to the accumulator.  `stats` optional. */
HJK_API int hjk_render(HjkContext* ctx, const HjkImageBlock* blocks, uint64_t n_blocks,
                       const HjkParams* params, HjkStats* stats);

/* Same, with the block list already resident in device memory (a pointer
 * previously returned by hjk_blocks_upload).  Used to time the device-resident path. */
HJK_API int hjk_blocks_upload(HjkContext* ctx, const HjkImageBlock* blocks, uint64_t n_blocks,
                              uint64_t* out_handle);
HJK_API int hjk_render_resident(HjkContext* ctx, uint64_t handle, uint64_t first_block,
                                uint64_t n_blocks, const HjkParams* params, HjkStats* stats);
HJK_API int hjk_blocks_free(HjkContext* ctx, uint64_t handle);

/* Replaces save_image's copy + map + divide (src/main.rs:1357-1400).
 * rgba: HOST destination, `pitch_bytes` per row (>= width*16).  normalise != 0
 * writes (r/w, g/w, b/w, w) like src/main.rs:1399, else the raw (sum w*rgb, sum w).
 * When hjk_comm_init joined several ranks this first sums the frame over the ranks (a collective: every rank
 * must call it).  The sum goes to a buffer of its own, the rank's frame is left untouched: reading back twice,
 * or rendering more passes and reading back again, gives the right frame each time. */
HJK_API int hjk_readback(HjkContext* ctx, float* rgba, uint64_t pitch_bytes, int normalise);
/* The same with the sum delivered to rank `root` only (ncclReduce instead of ncclAllReduce): only the root's
 * `rgba` is written, so W*H*16 bytes cross PCIe once per frame, not once per rank.  The other ranks may pass
 * NULL.  root < 0 = hjk_readback. */
HJK_API int hjk_readback_root(HjkContext* ctx, int root, float* rgba, uint64_t pitch_bytes, int normalise);
/* hjk_readback_root without the wait: the frame is summed and staged (normalised or raw) on the render stream, and
 * copied to `rgba` (which should be pinned) on a stream of its own — the call returns as soon as that is enqueued, and
 * the next hjk_frame_begin / hjk_render can be issued at once, so the copy of frame k overlaps the rendering of frame
 * k + 1.  `rgba` is complete after hjk_readback_wait (or the next readback call on the context). */
HJK_API int hjk_readback_begin(HjkContext* ctx, int root, float* rgba, uint64_t pitch_bytes, int normalise);
HJK_API int hjk_readback_wait(HjkContext* ctx);
/* Feature buffers of the frame (option "feature_buffers" = 1 before hjk_frame_begin): per texel the mean over
 * its samples of layer 1 = (first-hit normal, depth) (render.glsl:173) — summed by the reconstruction kernel,
 * reduced over ranks/devices together with the accumulator (one collective carries both), divided by the
 * sample count here.  Same `root` rule as hjk_readback_root. */
HJK_API int hjk_read_features(HjkContext* ctx, int root, float* normal_depth, uint64_t pitch_bytes);

/* Intermediate layers of the LAST pass rendered with HJK_RENDER_KEEP_FEATURES:
 * layer 0 = (radiance, 1), layer 1 = (normal, depth), layer 2 = (albedo == 0, 0)
 * (render.glsl:172-174), full-frame, width*height*4 floats into HOST memory. */
HJK_API int hjk_read_intermediate(HjkContext* ctx, int layer, float* rgba);

/* Parity hook (no reference analogue): closest hit of scene.glsl:97-175 on a
 * caller-supplied HOST ray batch.  shape_id = -1 on a miss (scene.glsl:98,160).
 * t/uv may be NULL.  `any_hit` is a bit set: bit 0 runs the shadow-ray (occlusion) traversal
 * instead and writes 0/1 into shape_id; bit 1 selects the exact-tie mode of HJK_RENDER_EXACT_TIES. */
HJK_API int hjk_trace_first_hit(HjkContext* ctx, const HjkRay* rays, uint64_t n_rays, int any_hit,
                                int32_t* shape_id, float* t, float* uv);
/* The same with M_EPS (math.glsl:2; 1e-4 above) as a parameter: the tie window of the exact mode. */
HJK_API int hjk_trace_first_hit_eps(HjkContext* ctx, const HjkRay* rays, uint64_t n_rays, int any_hit, float eps,
                                    int32_t* shape_id, float* t, float* uv);

/* Replaces ReconstructionPipeline::run (src/main.rs:992-1003) as a standalone
 * entry: splats caller-supplied full-frame intermediate layers (HOST pointers,
 * width*height float4 each; albedo may be NULL = zeros) of ONE pass described by
 * `blocks` into the accumulator. */
HJK_API int hjk_denoise_pass(HjkContext* ctx, const float* radiance, const float* normal_depth,
                             const float* albedo, const HjkImageBlock* blocks, uint64_t n_blocks,
                             const HjkParams* params);

/* Device-resident variant used for bandwidth measurement: uploads the layers
 * once (hjk_denoise_upload) and then applies them `repeat` times. */
HJK_API int hjk_denoise_upload(HjkContext* ctx, const float* radiance, const float* normal_depth,
                               const HjkImageBlock* blocks, uint64_t n_blocks);
HJK_API int hjk_denoise_resident(HjkContext* ctx, const HjkParams* params, uint32_t repeat,
                                 float* out_ms);

/* ------------------------------------------------------------ multi-GPU */

/* One process per GPU.  Rank 0 calls hjk_comm_unique_id (128 bytes), the host
 * program broadcasts it, every rank calls hjk_comm_init.  From then on
 *   - hjk_scene_upload is a collective: rank 0 builds the wide BVH and the others receive nodes and primitive
 *     records over ncclBroadcast (option "bvh_broadcast" = 0: every rank builds its own);
 *   - hjk_readback / hjk_readback_root / hjk_read_features sum the frame over the ranks, once per frame. */
HJK_API int hjk_comm_unique_id(void* out_id128);
HJK_API int hjk_comm_init(HjkContext* ctx, const void* id128, int rank, int n_ranks);
/* The frame's reduction on its own (what the readbacks do first).  A collective: every rank calls it, with the
 * same root, the same number of times.  root < 0: ncclAllReduce, else ncclReduce to `root`.  out_ms optional
 * (device time). */
HJK_API int hjk_reduce_frame(HjkContext* ctx, int root, float* out_ms);
HJK_API int hjk_allreduce_accumulator(HjkContext* ctx, float* out_ms); /* = hjk_reduce_frame(ctx, -1, out_ms) */

/* Device pointer of this rank's accumulator so a host framework that already owns a communicator (e.g.
 * torch.distributed) can reduce it in place (only without hjk_comm_init: with it the readbacks read the
 * library's own sum). */
HJK_API int hjk_accumulator_device_ptr(HjkContext* ctx, uint64_t* out_ptr, uint64_t* out_n_floats);
HJK_API int hjk_synchronize(HjkContext* ctx);
/* Run on a stream the host owns (a cudaStream_t, e.g. torch.cuda.Stream.cuda_stream) instead
 * of the context's private stream, so the host's own events and collectives order with it. */
HJK_API int hjk_set_stream(HjkContext* ctx, void* cuda_stream);

/* ----------------------------------------------------- misc / profiling */
HJK_API int hjk_set_profiling(HjkContext* ctx, int enabled); /* per-stage CUDA-event timing */
/* Tuning options (all have measured defaults; none changes results):
 *   "wave_paths"             camera paths rendered per wave (default 64 Mi = 13 GB of path state)
 *   "bvh_builder"            0 (default) = host SAH builder, 1 = GPU builder (Morton sort + PLOC clustering + collapse:
 *                            tens of milliseconds for 10 M triangles, traversal 3-16 % slower), -1 = the host builder up
 *                            to a million shapes and the GPU builder beyond; takes effect at the next hjk_scene_upload
 *   "bvh_gpu_tree"           the GPU builder's binary tree: 1 = PLOC (default), 0 = radix tree (LBVH)
 *   "bvh_validate"           1 = run the host structural check on a GPU-built tree
 *   "bvh_broadcast"          several ranks: 1 = rank 0 builds and broadcasts the wide BVH (default), 0 = every rank builds
 *   "feature_buffers"        1 = sum the first-hit (normal, depth) per texel for hjk_read_features (default 0)
 *   "bvh_pad_rel_e9"         outward pad of primitive boxes, in 1e-9 of the scene extent (default 10000)
 *   "fetch_threshold"        refill a traversal warp when fewer lanes than this are busy (default -1 = by scene: 20,
 *                            24 for trees that hold only spheres or exceed 64 MB; an idle warp always refills, so 0 is valid)
 *   "coop_trace"             1 = k_trace_coop (default): a warp pools the primitive tests of its leaves and
 *                            spreads them over all 32 lanes when that is cheaper; 0 = per-lane k_trace
 *   "coop_batch_cost"        assumed instructions per pooled batch of 32 tests (default -1 = by scene: 180, 260 for
 *                            trees that hold only spheres; 0 = always pool)
 *   "postpone_lanes"         per-lane k_trace: postpone primitive tests that fewer lanes than this would run
 *                            (default 8)
 *   "shade_sort"             1 = counting-sort the hits of a shading tile by material so a warp shades one material,
 *                            0 = shade in queue order, -1 (default) = sort only for scenes that mix diffuse-like,
 *                            mirror and dielectric surfaces
 *   "blocks_per_sm_traverse", "blocks_per_sm_tile"   persistent-grid sizes
 * Info keys: "n_sms", "bvh_nodes", "bvh_prims", "bvh_depth", "bvh_bytes", "bvh_builder", "bvh_build_us",
 *   "wave_paths", "has_extinction", "sphere_guard" (0 no spheres, 1 per-node flag, 2 every node), "unresolved_ties", "width", "height", "device",
 *   "blocks_per_sm_traverse", "blocks_per_sm_tile", "n_devices", "rank", "n_ranks", "feature_buffers", "shade_sort",
 *   "stack_overflows" (traversal-stack entries that did not fit since the scene upload: must read 0 — the
 *   upload refuses trees deeper than the 32-entry stack and primitive postponing stops short of it). */
HJK_API int hjk_set_option(HjkContext* ctx, const char* key, int64_t value);
HJK_API int hjk_get_info(HjkContext* ctx, const char* key, int64_t* out_value);
HJK_API const char* hjk_version(void);

/* ===================================================================== */
/* Host side (no GPU needed): mirror of the reference's Rust front-end.   */
/* ===================================================================== */

typedef struct HjkHostScene HjkHostScene;

/* Scene::from_obj + (optional) --put-cbox-spheres + Scene::compile
 * (src/main.rs:413-530,1463-1483,172-358).  with_bvh2 != 0 also builds the
 * reference-layout skip-pointer BVH2 (the `bvh` binding). */
HJK_API int hjk_host_scene_from_obj(const char* obj_path, int put_cbox_spheres, int with_bvh2,
                                    HjkHostScene** out_scene);
/* Synthetic scenes of BASELINE.json configs 3 and 4. */
HJK_API int hjk_host_scene_terrain(uint32_t grid_n, uint64_t seed, int with_bvh2,
                                   HjkHostScene** out_scene);
HJK_API int hjk_host_scene_spheres(uint32_t lattice_n, uint64_t seed, int with_bvh2,
                                   HjkHostScene** out_scene);
/* View of the compiled arrays; pointers stay valid until hjk_host_scene_free. */
HJK_API int hjk_host_scene_view(const HjkHostScene* scene, HjkScene* out_view);
HJK_API int hjk_host_scene_free(HjkHostScene* scene);
HJK_API const char* hjk_host_last_error(void);

/* Builds the 8-wide compressed BVH on the host (what hjk_scene_upload does) and reports
 * out[0] nodes, out[1] primitive records, out[2] depth, out[3] 1 if the structural check
 * passed (every shape reachable once, every box contains what is below it),
 * out[4] root SAH cost x 1000, out[5] bytes.  pad_rel < 0 selects the default pad. */
HJK_API int hjk_host_bvh_stats(const HjkScene* scene, float pad_rel, uint64_t* out6);
/* Same build on at most n_threads host threads (0 = all) and a 64-bit FNV-1a digest of the node and
 * primitive-record bytes: the builder runs its phases on every host thread, and the tree must not depend
 * on how many there are (checked by the tests with 1, 3 and all threads). */
HJK_API int hjk_host_bvh_digest(const HjkScene* scene, float pad_rel, int n_threads, uint64_t* out_digest);

/* ImageBlockGenerator (src/main.rs:619-682) with the OS-entropy draws replaced
 * by a recorded splitmix64 stream from `root_seed`.  Returns the block count
 * (tiles x spp); writes at most `capacity` blocks when out_blocks != NULL. */
HJK_API uint64_t hjk_host_generate_blocks(uint32_t width, uint32_t height, uint32_t block_size,
                                          uint32_t num_samples, uint64_t root_seed,
                                          HjkImageBlock* out_blocks, uint64_t capacity);

/* save_image's EXR write (src/main.rs:1402-1419): 3-channel float scanline EXR
 * from an RGBA float image already normalised by hjk_readback. */
HJK_API int hjk_host_write_exr(const char* path, const float* rgba, uint32_t width,
                               uint32_t height, uint64_t pitch_bytes);

#ifdef __cplusplus
}
#endif
#endif /* HIJIKI_B200_H */
