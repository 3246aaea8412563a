#include "pass_plan.h"

#include <algorithm>

namespace hjk {

bool plan_passes(const HjkImageBlock* blocks, uint64_t n, PassPlan& plan, std::string& err) {
  plan = PassPlan();
  if (!blocks || n == 0) {
    err = "empty block list";
    return false;
  }
  if (n > 0x7FFFFFFFull) {
    err = "more than 2^31 blocks in one call";
    return false;
  }
  plan.width = blocks[0].original_dimension[0];
  plan.height = blocks[0].original_dimension[1];
  if (plan.width == 0 || plan.height == 0 || (uint64_t)plan.width * plan.height > 0x7FFFFFFFull) {
    err = "unsupported image size";
    return false;
  }
  for (uint64_t i = 0; i < n; i++) {
    const HjkImageBlock& b = blocks[i];
    if (b.original_dimension[0] != plan.width || b.original_dimension[1] != plan.height) {
      err = "all blocks of one call must share original_dimension";
      return false;
    }
    if (b.dimension[0] == 0 || b.dimension[1] == 0 ||
        (uint64_t)b.origin[0] + b.dimension[0] > plan.width ||
        (uint64_t)b.origin[1] + b.dimension[1] > plan.height) {
      err = "block does not lie inside the image";
      return false;
    }
    plan.tile_w = std::max(plan.tile_w, b.dimension[0]);
    plan.tile_h = std::max(plan.tile_h, b.dimension[1]);
  }
  for (uint64_t i = 0; i < n; i++) {
    if (blocks[i].origin[0] % plan.tile_w || blocks[i].origin[1] % plan.tile_h) {
      err = "block origins must lie on the grid of the largest block dimension";
      return false;
    }
  }
  plan.tiles_x = (plan.width + plan.tile_w - 1) / plan.tile_w;
  plan.tiles_y = (plan.height + plan.tile_h - 1) / plan.tile_h;
  const size_t tiles = (size_t)plan.tiles_x * plan.tiles_y;
  PassPlan::Pass cur;
  std::vector<int32_t> map(tiles, -1);
  auto flush = [&]() {
    plan.passes.push_back(cur);
    plan.tile_block.insert(plan.tile_block.end(), map.begin(), map.end());
    std::fill(map.begin(), map.end(), -1);
  };
  for (uint64_t i = 0; i < n; i++) {
    const HjkImageBlock& b = blocks[i];
    const size_t t = (size_t)(b.origin[1] / plan.tile_h) * plan.tiles_x + b.origin[0] / plan.tile_w;
    if (map[t] >= 0) {  // tile already sampled in this pass: the next pass starts here
      flush();
      cur = PassPlan::Pass();
      cur.first_block = i;
    }
    map[t] = (int32_t)i;
    cur.n_blocks++;
    cur.n_pixels += (uint64_t)b.dimension[0] * b.dimension[1];
  }
  flush();
  return true;
}

}  // namespace hjk
