// Accumulation + bilateral reconstruction: reference shader/reconstruction.glsl:22-66,
// restated as a full-frame gather.
//
// The reference runs one dispatch per 128x128 block over the block plus an apron of R
// pixels and read-modify-writes the shared accumulator, block after block.  Here one pass
// (one set of non-overlapping blocks covering the frame) is reconstructed by one launch:
// every output pixel visits the (at most nine) blocks whose apron reaches it, in block-list
// order, and replays that block's taps in the reference's loop order — so the sequence of
// fp32 additions into each accumulator texel is the same as the reference's.
// Centre features of apron pixels (outside the sample's block) are zero, the robust
// out-of-bounds image load the oracle assumes.
#pragma once
#include "scene_dev.cuh"

namespace hjk {

struct PassDev {
  uint32_t width, height;
  uint32_t tile_w, tile_h;     // block grid pitch (reference: 128 x 128)
  uint32_t tiles_x, tiles_y;
  const int32_t* tile_block;   // [tiles_y * tiles_x] index into `blocks`, -1 = no block
  const HjkImageBlock* blocks;
  const float* weights;        // [(2R+1)^2] spatial weights per block, dx-major; < 0 = skipped tap
  int32_t radius;
};

// spatial weight of tap (dx, dy) for a block's sample offset — reconstruction.glsl:29-30,43-46
HJK_HD float recon_spatial_weight(int dx, int dy, int radius, float stddev, float so_x, float so_y) {
  const float gauss_fac = x::div(-1.0f, x::mul(x::mul(2.0f, stddev), stddev));
  const float curve_offset = exp_det(x::mul(x::mul(gauss_fac, (float)radius), (float)radius));
  const float sx = x::sub(x::add((float)dx, so_x), 0.5f);
  const float sy = x::sub(x::add((float)dy, so_y), 0.5f);
  const float w = x::sub(exp_det(x::mul(gauss_fac, x::add(x::mul(sx, sx), x::mul(sy, sy)))), curve_offset);
  return w < 0.f ? -1.0f : w;
}

// Layers: radiance(gx,gy) -> f4 (rgb, 1), feature(gx,gy) -> f4 (normal, depth).
// HAS_ALBEDO adds albedo(gx,gy) (the reference's layer 2, always zero on the render path).
template <bool HAS_ALBEDO, class Layers>
HJK_HD f4 reconstruct_pixel(const PassDev& ps, const Layers& L, uint32_t gx, uint32_t gy, f4 acc) {
  const int R = ps.radius;
  const int taps = 2 * R + 1;
  const int tx = (int)(gx / ps.tile_w), ty = (int)(gy / ps.tile_h);
  const int reach_x = (R + (int)ps.tile_w - 1) / (int)ps.tile_w;  // 1 unless R > tile
  const int reach_y = (R + (int)ps.tile_h - 1) / (int)ps.tile_h;
  for (int by = ty - reach_y; by <= ty + reach_y; by++) {
    if (by < 0 || by >= (int)ps.tiles_y) continue;
    for (int bx = tx - reach_x; bx <= tx + reach_x; bx++) {
      if (bx < 0 || bx >= (int)ps.tiles_x) continue;
      const int32_t b = ps.tile_block[by * (int)ps.tiles_x + bx];
      if (b < 0) continue;
      const HjkImageBlock& blk = ps.blocks[b];
      const int lx = (int)gx - (int)blk.origin[0], ly = (int)gy - (int)blk.origin[1];
      const int dimx = (int)blk.dimension[0], dimy = (int)blk.dimension[1];
      if (lx < -R || ly < -R || lx >= dimx + R || ly >= dimy + R) continue;
      const bool inside = lx >= 0 && ly >= 0 && lx < dimx && ly < dimy;
      vec3 nc = V3(0.f), ac = V3(0.f);
      if (inside) {
        nc = xyz(L.feature(gx, gy));
        if (HAS_ALBEDO) ac = xyz(L.albedo(gx, gy));
      }
      const float* wt = ps.weights + (size_t)b * taps * taps;
      for (int dx = -R; dx <= R; dx++) {
        const int sx = lx + dx;
        if (sx < 0 || sx >= dimx) continue;
        for (int dy = -R; dy <= R; dy++) {
          const int sy = ly + dy;
          if (sy < 0 || sy >= dimy) continue;
          float w = wt[(dx + R) * taps + (dy + R)];
          if (w < 0.f) continue;
          const uint32_t px = (uint32_t)((int)gx + dx), py = (uint32_t)((int)gy + dy);
          const f4 cw = L.radiance(px, py);
          const vec3 no = xyz(L.feature(px, py)) - nc;
          float e = x::mul(dot(no, no), 2.0f);
          if (HAS_ALBEDO) {
            const vec3 ao = xyz(L.albedo(px, py)) - ac;
            e = x::add(e, dot(ao, ao));
          }
          w = x::mul(w, exp_det(-e));
          const f4 wv = F4(x::mul(w, cw.x), x::mul(w, cw.y), x::mul(w, cw.z), x::mul(w, cw.w));
          if (x::is_nan(wv.x) || x::is_nan(wv.y) || x::is_nan(wv.z) || x::is_nan(wv.w)) continue;
          acc = F4(x::add(acc.x, wv.x), x::add(acc.y, wv.y), x::add(acc.z, wv.z), x::add(acc.w, wv.w));
        }
      }
    }
  }
  return acc;
}

}  // namespace hjk
