// Single-sweep tv_<scheme> kernels with the norms output (return_grad_norms=True, tv_GPU.py:128-131), float.
#define PYTVB_TILE_NORMS true
#define PYTVB_TILE_ENTRY run_tv_tile_norms
#define PYTVB_TILE_T float
#include "tv_tile_impl.cuh"
