// Instantiates only the float / VEC 4 / z + t kernels of the two tile-kernel forms (scripts/sass_probe.sh).
#include "kernels_tile.cuh"
#ifndef SCH
#define SCH 3
#endif
namespace pytvb {
void* probe_ptrs[2] = {(void*)tv_tile2_kernel<float, 4, SCH, true, true, 4, 0, false>, (void*)tv_tile_kernel<float, 4, SCH, true, true, 4, 0, false>};
}
