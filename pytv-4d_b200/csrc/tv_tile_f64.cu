// Single-sweep tv_<scheme> kernels without the norms output, double.
#define PYTVB_TILE_NORMS false
#define PYTVB_TILE_ENTRY run_tv_tile_plain
#define PYTVB_TILE_T double
#include "tv_tile_impl.cuh"
