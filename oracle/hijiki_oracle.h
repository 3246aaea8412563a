/*
 * hijiki_oracle.h — CPU ORACLE for the Hijiki path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a literal CPU restatement of the reference
 * shaders (reference shader/{rand,math,quaternion,block,render,scene,material,
 * reconstruction}.glsl, shader/shapes/ and shader/materials/) used as the
 * checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (hijiki_b200/) never does.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
 * cannot be built in this image (no Rust, no Vulkan, needs a window).  The oracle is
 * pinned by the known-answer vectors derived from the reference arithmetic in
 * SURVEY.md §8c (tests/test_oracle_kat.py) and by a second, independent restatement of whole
 * paths and of a reconstruction block in numpy float32 (tests/test_oracle_crosscheck.py) —
 * two restatements agreeing bit for bit, not the reference itself.  Floating-point conventions the GLSL spec
 * leaves open are fixed here: fp32 everywhere, no FMA contraction, IEEE div/sqrt,
 * sin/cos/tan/exp/atan/asin = the fixed polynomial kernels specified in orc_math.h
 * (<= 2 ulp from libm on the ranges the path uses, tests/test_math_spec.py),
 * normalize(v) = v * (1/sqrt(dot(v,v))).
 */
#ifndef HIJIKI_ORACLE_H
#define HIJIKI_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcArray {
  const void* ptr;
  uint64_t count;
} OrcArray;

/* Same 12 arrays, same byte layouts as the reference's CompiledScene bindings
 * (reference src/main.rs:314-327,561-605). */
typedef struct OrcScene {
  OrcArray scene, bvh, spheres, quads, triangles, vertices, materials, emitters, diffuse,
      diffusecb, dielectric, emissive;
} OrcScene;

typedef struct OrcBlock { /* reference block.glsl:1-8 */
  uint32_t id, seed;
  uint32_t origin[2], dimension[2], original_dimension[2];
  float sample_offset[2];
} OrcBlock;

typedef struct OrcRay {
  float origin[3];
  float t_min;
  float direction[3];
  float t_max;
} OrcRay;

typedef struct OrcParams {
  uint32_t max_bounces; /* render.glsl:92 (1000) */
  uint32_t rr_start;    /* render.glsl:137 (3) */
  uint32_t recon_radius;
  float recon_stddev;
  float eps;            /* math.glsl:2 */
  uint32_t use_bvh;     /* scene.glsl USE_BVH: 0 = linear scan (reference default), 1 = threaded BVH2;
                           oracle extensions: 2 = linear scan without the >100-shape failsafe, 3 = the scan of
                           mode 2 over the primitives whose box the ray pierces (same results, tractable at
                           BASELINE sizes; hijiki_oracle.cpp "mode 3") */
  uint32_t block_size;  /* intermediate texture edge, src/main.rs:1197-1201 (128) */
  uint32_t skip_recon;
} OrcParams;

typedef struct OrcStats {
  uint64_t n_paths, n_extension_rays, n_shadow_rays;
  double seconds;
} OrcStats;

typedef struct OrcPathVertex { /* debugging log of one path, one entry per bounce */
  int32_t shape_id;
  float t;
  uint32_t rng_after; /* rngState at the end of the bounce */
  float throughput[3];
  float total[3];
  int32_t shadow_state; /* 0 = no shadow ray, 1 = occluded, 2 = unoccluded */
  float dir_len2;       /* dot(d, d) of the ray that produced this vertex (directions drift off unit length) */
} OrcPathVertex;

/* rand.glsl:1-20 */
uint32_t orc_seed_rng(uint32_t seed);
uint32_t orc_rand_uint(uint32_t* state);
float orc_rand_uniform_float(uint32_t* state);
/* render.glsl:26-36; out_ray = OrcRay */
void orc_camera_ray(const void* scene_info64, float px, float py, float dim_x, float dim_y,
                    float eps, OrcRay* out_ray);
/* reconstruction.glsl:29-46: the (2R+1)^2 spatial weights, dx-major, negative ones clamped to -1 */
void orc_recon_spatial_weights(uint32_t radius, float stddev, float so_x, float so_y, float* out);

/* scene.glsl:97-175 on a ray batch.  shape_id -1 on miss.  tie (optional, linear mode only):
 * 1 when the reported primitive depends on test order. uv optional. */
int orc_trace(const OrcScene* scene, const OrcRay* rays, uint64_t n, int use_bvh, float eps,
              int32_t* shape_id, float* t, float* uv, uint8_t* tie, int n_threads);
/* shadow overload scene.glsl:92-96: occluded[i] = intersectScene(ray) */
int orc_occluded(const OrcScene* scene, const OrcRay* rays, uint64_t n, int use_bvh, float eps,
                 uint8_t* occluded, int n_threads);

/* Renderer::render (src/main.rs:1316-1355): per block, render.glsl main then
 * reconstruction.glsl main; accumulator = width*height float4, ADDED to. */
int orc_render(const OrcScene* scene, const OrcBlock* blocks, uint64_t n_blocks,
               const OrcParams* params, float* accumulator, OrcStats* stats, int n_threads);

/* render.glsl main for a list of blocks into FULL-FRAME intermediate layers
 * (3 x width*height float4; layer 0 radiance,1 / layer 1 normal,depth / layer 2 albedo,0);
 * blocks must not overlap. */
int orc_integrate_frame(const OrcScene* scene, const OrcBlock* blocks, uint64_t n_blocks,
                        const OrcParams* params, float* layers, OrcStats* stats, int n_threads);

/* reconstruction.glsl main for a list of blocks whose samples live in full-frame layers
 * (layer 0, layer 1, optional layer 2) — emulates the per-block 128^2 intermediate
 * texture, including zero centre features for apron pixels. */
int orc_reconstruct_frame(const OrcBlock* blocks, uint64_t n_blocks, const OrcParams* params,
                          const float* radiance, const float* normal_depth, const float* albedo,
                          float* accumulator, int n_threads);

/* One pixel's path, bounce by bounce (debug). Returns number of vertices written. */
int orc_trace_path(const OrcScene* scene, const OrcBlock* block, uint32_t lx, uint32_t ly,
                   const OrcParams* params, OrcPathVertex* out, int capacity);

int orc_hardware_threads(void);

/* orc_math.h evaluated over an array (tests).  fn: 0 sin, 1 cos, 2 tan, 3 exp, 4 atan2(a,b), 5 asin,
 * 6 exp_bilateral(a) = exp(-a) in the FMA specification of the reconstruction weight */
void orc_math_eval(int fn, const float* a, const float* b, float* out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
