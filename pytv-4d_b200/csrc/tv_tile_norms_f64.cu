// Single-sweep tv_<scheme> kernels with the norms output, double.
#define PYTVB_TILE_NORMS true
#define PYTVB_TILE_ENTRY run_tv_tile_norms
#define PYTVB_TILE_T double
#include "tv_tile_impl.cuh"
