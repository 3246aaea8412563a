// C ABI of the host-side front-end (no CUDA): scene loading/compiling, the block
// generator and the EXR writer.  Declared in include/hijiki_b200.h.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "host_scene.h"
#include "wide_bvh_host.h"

struct HjkHostScene {
  hjk::CompiledScene compiled;
};

namespace {
thread_local std::string g_host_error;

template <class T>
HjkArray view_of(const std::vector<T>& v) {
  return HjkArray{v.empty() ? nullptr : (const void*)v.data(), (uint64_t)v.size()};
}
}  // namespace

extern "C" {

const char* hjk_host_last_error(void) { return g_host_error.c_str(); }

int hjk_host_scene_from_obj(const char* obj_path, int put_spheres, int with_bvh2,
                            HjkHostScene** out_scene) {
  if (!obj_path || !out_scene) {
    g_host_error = "hjk_host_scene_from_obj: null argument";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  try {
    hjk::Scene scene;
    std::string err;
    if (!hjk::scene_from_obj(obj_path, scene, err)) {
      g_host_error = err;
      return HJK_ERR_IO;
    }
    if (put_spheres) hjk::put_cbox_spheres(scene);
    HjkHostScene* hs = new HjkHostScene();
    hjk::compile_scene(scene, with_bvh2 != 0, hs->compiled);
    *out_scene = hs;
    return HJK_OK;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return HJK_ERR_OUT_OF_MEMORY;
  }
}

int hjk_host_scene_terrain(uint32_t grid_n, uint64_t seed, int with_bvh2, HjkHostScene** out) {
  if (!out || grid_n == 0 || grid_n > 8192) {
    g_host_error = "hjk_host_scene_terrain: grid_n must be in [1, 8192]";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  try {
    hjk::Scene scene;
    hjk::make_terrain_scene(grid_n, seed, scene);
    HjkHostScene* hs = new HjkHostScene();
    hjk::compile_scene(scene, with_bvh2 != 0, hs->compiled);
    *out = hs;
    return HJK_OK;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return HJK_ERR_OUT_OF_MEMORY;
  }
}

int hjk_host_scene_spheres(uint32_t lattice_n, uint64_t seed, int with_bvh2, HjkHostScene** out) {
  if (!out || lattice_n == 0 || lattice_n > 256) {
    g_host_error = "hjk_host_scene_spheres: lattice_n must be in [1, 256]";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  try {
    hjk::Scene scene;
    hjk::make_spheres_scene(lattice_n, seed, scene);
    HjkHostScene* hs = new HjkHostScene();
    hjk::compile_scene(scene, with_bvh2 != 0, hs->compiled);
    *out = hs;
    return HJK_OK;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return HJK_ERR_OUT_OF_MEMORY;
  }
}

int hjk_host_scene_view(const HjkHostScene* hs, HjkScene* v) {
  if (!hs || !v) {
    g_host_error = "hjk_host_scene_view: null argument";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  const hjk::CompiledScene& c = hs->compiled;
  v->scene = HjkArray{&c.info, 1};
  v->bvh = view_of(c.bvh);
  v->spheres = view_of(c.spheres);
  v->quads = view_of(c.quads);
  v->triangles = view_of(c.triangles);
  v->vertices = view_of(c.vertices);
  v->materials = view_of(c.materials);
  v->emitters = view_of(c.emitters);
  v->diffuse = view_of(c.diffuse);
  v->diffusecb = view_of(c.diffusecb);
  v->dielectric = view_of(c.dielectric);
  v->emissive = view_of(c.emissive);
  return HJK_OK;
}

int hjk_host_bvh_stats(const HjkScene* scene, float pad_rel, uint64_t* out6) {
  if (!scene || !out6) {
    g_host_error = "hjk_host_bvh_stats: null argument";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  try {
    hjk::WideBvh bvh;
    std::string err;
    if (!hjk::build_wide_bvh(*scene, pad_rel < 0.f ? hjk::kDefaultBvhPadRel : pad_rel, bvh, err)) {
      g_host_error = err;
      return HJK_ERR_INVALID_ARGUMENT;
    }
    const bool ok = hjk::validate_wide_bvh(*scene, bvh, err);
    if (!ok) g_host_error = err;
    out6[0] = bvh.nodes.size();
    out6[1] = bvh.prims.size();
    out6[2] = bvh.depth;
    out6[3] = ok ? 1 : 0;
    out6[4] = (uint64_t)(bvh.sah_cost * 1000.f);
    out6[5] = bvh.nodes.size() * sizeof(hjk::WideNode) + bvh.prims.size() * sizeof(hjk::WidePrim);
    return HJK_OK;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return HJK_ERR_OUT_OF_MEMORY;
  }
}

int hjk_host_bvh_digest(const HjkScene* scene, float pad_rel, int n_threads, uint64_t* out_digest) {
  if (!scene || !out_digest) {
    g_host_error = "hjk_host_bvh_digest: null argument";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  try {
    hjk::WideBvh bvh;
    std::string err;
    hjk::set_builder_threads(n_threads);
    const bool built = hjk::build_wide_bvh(*scene, pad_rel < 0.f ? hjk::kDefaultBvhPadRel : pad_rel, bvh, err);
    hjk::set_builder_threads(0);
    if (!built) {
      g_host_error = err;
      return HJK_ERR_INVALID_ARGUMENT;
    }
    uint64_t h = 1469598103934665603ull;  // FNV-1a
    auto eat = [&h](const void* p, size_t n) {
      const unsigned char* b = static_cast<const unsigned char*>(p);
      for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    };
    eat(bvh.nodes.data(), bvh.nodes.size() * sizeof(hjk::WideNode));
    eat(bvh.prims.data(), bvh.prims.size() * sizeof(hjk::WidePrim));
    *out_digest = h;
    return HJK_OK;
  } catch (const std::exception& e) {
    hjk::set_builder_threads(0);
    g_host_error = e.what();
    return HJK_ERR_OUT_OF_MEMORY;
  }
}

int hjk_host_scene_free(HjkHostScene* hs) {
  delete hs;
  return HJK_OK;
}

// ImageBlockGenerator::next (src/main.rs:648-682).  Literal restatement, including the
// quirk that `sample_offset` is re-drawn BEFORE the last block of a pass is emitted
// (src/main.rs:664-680), so that block already carries the next pass' offset.
uint64_t hjk_host_generate_blocks(uint32_t width, uint32_t height, uint32_t block_size,
                                  uint32_t num_samples, uint64_t root_seed, HjkImageBlock* out,
                                  uint64_t capacity) {
  if (width == 0 || height == 0 || block_size == 0 || (block_size & 63u) != 0) return 0;
  const uint64_t tiles_x = (width + block_size - 1) / block_size;
  const uint64_t tiles_y = (height + block_size - 1) / block_size;
  const uint64_t total = tiles_x * tiles_y * (uint64_t)num_samples;
  if (!out) return total;
  hjk::SplitMix64 rng(root_seed);
  uint32_t id = 0, x = 0, y = 0, remaining = num_samples;
  float so[2] = {rng.next_f32(), rng.next_f32()};  // ImageBlockGenerator::new, :643
  uint64_t n = 0;
  while (remaining != 0 && n < capacity) {
    uint32_t bx = x, by = y;
    uint32_t w = block_size < width - bx ? block_size : width - bx;
    uint32_t h = block_size < height - by ? block_size : height - by;
    uint32_t bid = id++;
    x += block_size;
    if (x >= width) {
      x = 0;
      y += block_size;
      if (y >= height) {
        y = 0;
        remaining--;
        so[0] = rng.next_f32();
        so[1] = rng.next_f32();
      }
    }
    HjkImageBlock b;
    b.id = bid;
    b.seed = rng.next_u32();
    b.origin[0] = bx;
    b.origin[1] = by;
    b.dimension[0] = w;
    b.dimension[1] = h;
    b.original_dimension[0] = width;
    b.original_dimension[1] = height;
    b.sample_offset[0] = so[0];
    b.sample_offset[1] = so[1];
    out[n++] = b;
  }
  return total;
}

// Minimal scanline OpenEXR (uncompressed, FLOAT channels B,G,R) — what the reference
// writes through the openexr crate in save_image (src/main.rs:1402-1419).
int hjk_host_write_exr(const char* path, const float* rgba, uint32_t width, uint32_t height,
                       uint64_t pitch_bytes) {
  if (!path || !rgba || width == 0 || height == 0 || pitch_bytes < (uint64_t)width * 16) {
    g_host_error = "hjk_host_write_exr: invalid argument";
    return HJK_ERR_INVALID_ARGUMENT;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    g_host_error = std::string("cannot create ") + path;
    return HJK_ERR_IO;
  }
  std::string hdr;
  auto put_u32 = [&](uint32_t v) { hdr.append((const char*)&v, 4); };
  auto put_f32 = [&](float v) { hdr.append((const char*)&v, 4); };
  auto put_str = [&](const char* s) { hdr.append(s, strlen(s) + 1); };
  auto attr = [&](const char* name, const char* type, uint32_t size) {
    put_str(name);
    put_str(type);
    put_u32(size);
  };
  put_u32(20000630u);  // magic
  put_u32(2u);         // version 2, single-part scanline
  attr("channels", "chlist", 3 * 18 + 1);
  for (const char* ch : {"B", "G", "R"}) {
    put_str(ch);
    put_u32(2);  // FLOAT
    put_u32(0);  // pLinear + reserved
    put_u32(1);  // xSampling
    put_u32(1);  // ySampling
  }
  hdr.push_back('\0');
  attr("compression", "compression", 1);
  hdr.push_back('\0');  // NO_COMPRESSION
  for (const char* win : {"dataWindow", "displayWindow"}) {
    attr(win, "box2i", 16);
    put_u32(0);
    put_u32(0);
    put_u32(width - 1);
    put_u32(height - 1);
  }
  attr("lineOrder", "lineOrder", 1);
  hdr.push_back('\0');  // INCREASING_Y
  attr("pixelAspectRatio", "float", 4);
  put_f32(1.f);
  attr("screenWindowCenter", "v2f", 8);
  put_f32(0.f);
  put_f32(0.f);
  attr("screenWindowWidth", "float", 4);
  put_f32(1.f);
  hdr.push_back('\0');  // end of header

  const uint64_t line_bytes = (uint64_t)width * 3 * 4;
  const uint64_t table_pos = hdr.size();
  const uint64_t first = table_pos + (uint64_t)height * 8;
  bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
  for (uint32_t y = 0; ok && y < height; y++) {
    uint64_t off = first + (uint64_t)y * (8 + line_bytes);
    ok = fwrite(&off, 8, 1, f) == 1;
  }
  std::string line(line_bytes, '\0');
  for (uint32_t y = 0; ok && y < height; y++) {
    const float* row = (const float*)((const char*)rgba + (uint64_t)y * pitch_bytes);
    float* dst = (float*)&line[0];
    for (int c = 0; c < 3; c++) {
      const int src_c = 2 - c;  // B, G, R planes
      for (uint32_t x = 0; x < width; x++) dst[(uint64_t)c * width + x] = row[4 * x + src_c];
    }
    int32_t yy = (int32_t)y, sz = (int32_t)line_bytes;
    ok = fwrite(&yy, 4, 1, f) == 1 && fwrite(&sz, 4, 1, f) == 1 &&
         fwrite(line.data(), 1, line_bytes, f) == line_bytes;
  }
  fclose(f);
  if (!ok) {
    g_host_error = std::string("short write to ") + path;
    return HJK_ERR_IO;
  }
  return HJK_OK;
}

}  // extern "C"
