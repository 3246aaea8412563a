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
  const uint32_t* taps;        // per block: {count, 0}, then (2R+1)^2 x {weight bits, packed offset} of the
                               // taps with weight >= 0 in loop order (recon_tap_stride words per block,
                               // 8-byte aligned pairs); packed offset = (dx+128) | (dy+128) << 8 |
                               // int16(16 * (dy * recon_smem_pitch(R) + dx)) << 16: the tap's BYTE offset in
                               // k_recon's shared-memory tile of float4 texels (|.| <= 16 * 392)
  int32_t radius;
  float one;                   // 1.0f (launch_recon sets it), opaque to the compiler: see ReconPairs (kernels.cuh)
};
// pitch of k_recon's shared-memory tile (32 texels + halo); baked into the tap table
HJK_HD int recon_smem_pitch(int radius) { return 32 + 2 * radius; }
HJK_HD uint32_t recon_tap_stride(int radius) { return 2u + 2u * (uint32_t)((2 * radius + 1) * (2 * radius + 1)); }

// spatial weight of tap (dx, dy) for a block's sample offset — reconstruction.glsl:29-30,43-46
HJK_HD float recon_spatial_weight(int dx, int dy, int radius, float stddev, float so_x, float so_y) {
  const float gauss_fac = x::div(-1.0f, x::mul(x::mul(2.0f, stddev), stddev));
  const float curve_offset = exp_det(x::mul(x::mul(gauss_fac, (float)radius), (float)radius));
  const float sx = x::sub(x::add((float)dx, so_x), 0.5f);
  const float sy = x::sub(x::add((float)dy, so_y), 0.5f);
  const float w = x::sub(exp_det(x::mul(gauss_fac, x::add(x::mul(sx, sx), x::mul(sy, sy)))), curve_offset);
  return w < 0.f ? -1.0f : w;
}

// Fills one block's weight table and its compact tap list.
HJK_HD void recon_fill_block_tables(const HjkImageBlock& blk, int radius, float stddev, float* weights,
                                    uint32_t* taps) {
  const int t = 2 * radius + 1;
  uint32_t n = 0;
  for (int dx = -radius; dx <= radius; dx++)
    for (int dy = -radius; dy <= radius; dy++) {
      const float w = recon_spatial_weight(dx, dy, radius, stddev, blk.sample_offset[0], blk.sample_offset[1]);
      weights[(dx + radius) * t + (dy + radius)] = w;
      if (w < 0.f) continue;
      taps[2 + 2 * n] = x::as_uint(w);
      taps[3 + 2 * n] = (uint32_t)(dx + 128) | ((uint32_t)(dy + 128) << 8) |
                        ((uint32_t)(uint16_t)(int16_t)(16 * (dy * recon_smem_pitch(radius) + dx)) << 16);
      n++;
    }
  taps[0] = n;
  taps[1] = 0;
}

// one tap of reconstruction.glsl:47-59: bilateral factor, NaN rejection, accumulate
template <bool HAS_ALBEDO, class Layers>
HJK_HD void recon_tap(const Layers& L, uint32_t px, uint32_t py, float w, vec3 nc, vec3 ac, f4& acc) {
  const f4 cw = L.radiance(px, py);
  const vec3 no = xyz(L.feature(px, py)) - nc;
  float e = x::mul(dot(no, no), 2.0f);
  if (HAS_ALBEDO) {
    const vec3 ao = xyz(L.albedo(px, py)) - ac;
    e = x::add(e, dot(ao, ao));
  }
  w = x::mul(w, exp_fma_neg(e));  // exp(-e) in the FMA specification (hjk_math.cuh); exp_fma_neg(0) == 1
  const f4 wv = F4(x::mul(w, cw.x), x::mul(w, cw.y), x::mul(w, cw.z), x::mul(w, cw.w));
  if (x::is_nan(wv.x) || x::is_nan(wv.y) || x::is_nan(wv.z) || x::is_nan(wv.w)) return;  // :55-58
  acc = F4(x::add(acc.x, wv.x), x::add(acc.y, wv.y), x::add(acc.z, wv.z), x::add(acc.w, wv.w));
}

// Layers: radiance(gx,gy) -> f4 (rgb, 1), feature(gx,gy) -> f4 (normal, depth).
// HAS_ALBEDO adds albedo(gx,gy) (the reference's layer 2, always zero on the render path).
//
// Every texel runs the same loop: the blocks whose apron reaches it (1 in the interior, 2 at an
// edge, 4 at a corner — visited in block-list order, like the reference's block-after-block
// dispatches) x that block's compact tap list (the taps whose spatial weight is >= 0, in the
// reference's dx-major loop order), skipping taps whose sample lies outside the block.
template <bool HAS_ALBEDO, class Layers>
HJK_HD f4 reconstruct_pixel(const PassDev& ps, const Layers& L, uint32_t gx, uint32_t gy, f4 acc) {
  const int R = ps.radius;
  const uint32_t stride = recon_tap_stride(R);
  // tiles that can hold a sample within R of this texel
  const int tx0 = (int)gx - R < 0 ? 0 : ((int)gx - R) / (int)ps.tile_w;
  const int ty0 = (int)gy - R < 0 ? 0 : ((int)gy - R) / (int)ps.tile_h;
  int tx1 = ((int)gx + R) / (int)ps.tile_w, ty1 = ((int)gy + R) / (int)ps.tile_h;
  if (tx1 >= (int)ps.tiles_x) tx1 = (int)ps.tiles_x - 1;
  if (ty1 >= (int)ps.tiles_y) ty1 = (int)ps.tiles_y - 1;
  for (int by = ty0; by <= ty1; by++) {
    for (int bx = tx0; bx <= tx1; bx++) {
      const int32_t b = ps.tile_block[by * (int)ps.tiles_x + bx];
      if (b < 0) continue;
      const HjkImageBlock& blk = ps.blocks[b];
      const int lx = (int)gx - (int)blk.origin[0], ly = (int)gy - (int)blk.origin[1];
      const int dimx = (int)blk.dimension[0], dimy = (int)blk.dimension[1];
      if (lx < -R || ly < -R || lx >= dimx + R || ly >= dimy + R) continue;
      // centre features: the block's own texel, or the zero a robust out-of-bounds load returns
      // for apron texels
      const bool inside = lx >= 0 && ly >= 0 && lx < dimx && ly < dimy;
      vec3 nc = V3(0.f), ac = V3(0.f);
      if (inside) {
        nc = xyz(L.feature(gx, gy));
        if (HAS_ALBEDO) ac = xyz(L.albedo(gx, gy));
      }
      const uint32_t* tl = ps.taps + (size_t)b * stride;
      const uint32_t n = tl[0];
      for (uint32_t k = 0; k < n; k++) {
        const uint32_t wbits = tl[2 + 2 * k], o = tl[3 + 2 * k];  // one 8-byte load
        const int dx = (int)(o & 0xFFu) - 128, dy = (int)((o >> 8) & 0xFFu) - 128;
        const int sx = lx + dx, sy = ly + dy;
        if (sx < 0 || sy < 0 || sx >= dimx || sy >= dimy) continue;
        recon_tap<HAS_ALBEDO>(L, (uint32_t)((int)gx + dx), (uint32_t)((int)gy + dy), x::as_float(wbits), nc, ac,
                              acc);
      }
    }
  }
  return acc;
}

}  // namespace hjk
