// Single-sweep tv_<scheme> kernels without the norms output (the hot path), float; and the entry that picks between the sets.
#define PYTVB_TILE_NORMS false
#define PYTVB_TILE_ENTRY run_tv_tile_plain
#define PYTVB_TILE_T float
#include "tv_tile_impl.cuh"

namespace pytvb {
template <typename T> int run_tv_tile(int vec, int scheme, bool z_on, bool t_on, const TvArgs<T>& a) {
    return a.norms ? run_tv_tile_norms<T>(vec, scheme, z_on, t_on, a) : run_tv_tile_plain<T>(vec, scheme, z_on, t_on, a);
}
template int run_tv_tile<float>(int, int, bool, bool, const TvArgs<float>&);
template int run_tv_tile<double>(int, int, bool, bool, const TvArgs<double>&);
}  // namespace pytvb
