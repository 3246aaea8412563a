/* The reference's main() (reference src/main.rs:1459-1494) as a plain C99 host over the C ABI of
 * include/hijiki_b200.h — no Python, no C++ types: load the OBJ, enumerate the ImageBlocks, render,
 * read back, write the EXR.
 *
 *   cc -std=c99 -Iinclude examples/render_c.c -Lhijiki_b200/lib -lhijiki_b200 -Wl,-rpath,$PWD/hijiki_b200/lib -o render_c
 *   ./render_c scenes/cbox/cbox.obj 800 600 64 /tmp/output.exr
 */
#include <stdio.h>
#include <stdlib.h>

#include "hijiki_b200.h"

static void die(const char* what, const char* why) {
  fprintf(stderr, "render_c: %s: %s\n", what, why ? why : "?");
  exit(1);
}

int main(int argc, char** argv) {
  const char* obj = argc > 1 ? argv[1] : "scenes/cbox/cbox.obj";
  const uint32_t width = argc > 2 ? (uint32_t)atoi(argv[2]) : 800u;   /* src/main.rs:1437-1449 defaults */
  const uint32_t height = argc > 3 ? (uint32_t)atoi(argv[3]) : 600u;
  const uint32_t spp = argc > 4 ? (uint32_t)atoi(argv[4]) : 64u;
  const char* out = argc > 5 ? argv[5] : "/tmp/output.exr";

  /* Scene::from_obj + compile (src/main.rs:413-530,172-358) */
  HjkHostScene* host_scene = NULL;
  if (hjk_host_scene_from_obj(obj, 0, 0, &host_scene) != HJK_OK) die("scene", hjk_host_last_error());
  HjkScene scene;
  if (hjk_host_scene_view(host_scene, &scene) != HJK_OK) die("scene view", hjk_host_last_error());

  /* ImageBlockGenerator::new(width, height, 128, spp).collect() (src/main.rs:1485,1175) */
  const uint64_t n_blocks = hjk_host_generate_blocks(width, height, 128, spp, 0x48494A494B49ull, NULL, 0);
  HjkImageBlock* blocks = (HjkImageBlock*)malloc((size_t)n_blocks * sizeof(HjkImageBlock));
  if (!blocks) die("blocks", "out of memory");
  hjk_host_generate_blocks(width, height, 128, spp, 0x48494A494B49ull, blocks, n_blocks);

  /* GPU::new + Renderer::new (src/main.rs:692-712,1167-1314) */
  HjkContext* ctx = NULL;
  const int device = 0;
  if (hjk_create(&device, 1, &ctx) != HJK_OK) die("hjk_create", hjk_last_error(NULL));
  if (hjk_scene_upload(ctx, &scene) != HJK_OK) die("hjk_scene_upload", hjk_last_error(ctx));
  hjk_host_scene_free(host_scene); /* the library copied what it needs */
  if (hjk_frame_begin(ctx, width, height) != HJK_OK) die("hjk_frame_begin", hjk_last_error(ctx));

  /* Renderer::render (src/main.rs:1316-1355) */
  HjkParams params;
  params.max_bounces = 1000, params.rr_start = 3, params.recon_radius = 2, params.recon_stddev = 0.5f;
  params.eps = 1e-4f, params.flags = 0;
  HjkStats st;
  printf("Starting rendering\n");
  if (hjk_render(ctx, blocks, n_blocks, &params, &st) != HJK_OK) die("hjk_render", hjk_last_error(ctx));
  const double rays = (double)(st.n_extension_rays + st.n_shadow_rays);
  printf("Integrated %llu paths = %.0f rays over all bounces in %.1f ms (%.0f Mrays/s, %llu launches)\n",
         (unsigned long long)st.n_paths, rays, st.ms_total, rays / (st.ms_total * 1e3), (unsigned long long)st.n_launches);

  /* Renderer::save_image (src/main.rs:1357-1423) */
  float* rgba = (float*)malloc((size_t)width * height * 4 * sizeof(float));
  if (!rgba) die("frame", "out of memory");
  if (hjk_readback(ctx, rgba, (uint64_t)width * 16, 1) != HJK_OK) die("hjk_readback", hjk_last_error(ctx));
  if (hjk_host_write_exr(out, rgba, width, height, (uint64_t)width * 16) != HJK_OK) die("exr", hjk_host_last_error());
  printf("wrote %s\n", out);

  free(rgba);
  free(blocks);
  hjk_destroy(ctx);
  return 0;
}
