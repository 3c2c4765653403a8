// place_records.cuh — the per-cell candidate records place_prep_kernel (greedy.cu) writes and the placement chains read
// (greedy.cu: one consumer warp behind a TMA ring; greedy_wave.cu: a window of cells iterated to its fixed point).
#pragma once
#include "common.cuh"

namespace mosaic {

constexpr uint32_t CH_IDMASK = 0x03FFFFFFu;  // material id inside a record word (the lane sits in the six bits above)

struct __align__(16) HRec {  // what every cell needs
  uint2 wm[32];  // .x = lane << 26 | material id (the dummy id `n`, usage pinned at the cap, for lanes >= k)
                 // .y = threat window mask
};
struct __align__(16) XRec {  // what only contended cells need
  float d2[32], l[32], a[32], b[32];
  float d2_32;          // d2 of candidate 32 (+inf if k <= 32)
  uint32_t win32;       // window mask of candidate 32 (lane 0 of HRec carries all ones in the first column)
  uint32_t pad[2];
};
static_assert(sizeof(HRec) == 256 && sizeof(XRec) == 528, "record layout");

}  // namespace mosaic
