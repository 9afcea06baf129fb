// tests/emul/simt/cooperative_groups.h — TEST HARNESS ONLY: grid.sync() of the SIMT emulator (see cuda_runtime.h).
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group {
  void sync() const { emu::grid_sync(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
