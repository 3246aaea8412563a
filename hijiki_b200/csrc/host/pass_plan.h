// Groups an ImageBlock list (reference src/main.rs:608-682) into passes: maximal runs of
// consecutive blocks that do not overlap, laid out on a regular tile grid.  The reference
// renders block after block (src/main.rs:1316-1355); the wavefront pipeline renders a whole
// pass (or several) per launch, which is result-identical because blocks of one pass touch
// disjoint samples and the reconstruction replays them in list order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/hijiki_b200.h"

namespace hjk {

struct PassPlan {
  uint32_t width = 0, height = 0;
  uint32_t tile_w = 0, tile_h = 0;  // grid pitch = largest block dimension (reference: 128)
  uint32_t tiles_x = 0, tiles_y = 0;
  struct Pass {
    uint64_t first_block = 0, n_blocks = 0;  // range in the caller's list
    uint64_t n_pixels = 0;                   // camera paths of this pass
  };
  std::vector<Pass> passes;
  std::vector<int32_t> tile_block;  // [pass][tiles_y][tiles_x] -> index in the caller's list or -1
};

bool plan_passes(const HjkImageBlock* blocks, uint64_t n_blocks, PassPlan& plan, std::string& err);

}  // namespace hjk
