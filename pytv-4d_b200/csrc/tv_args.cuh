// Arguments of one pytvb_tv call, shared by tv.cu (entry point, two-sweep fallback) and tv_tile.cu (single-sweep kernel).
#pragma once
#include "host_common.cuh"

namespace pytvb {

template <typename T> struct TvArgs {
    ImgView<T> X; ImgView<T> W; T* Wz0; T* G; T* norms; double* partial; Params<T> P; int z_lo, nz; cudaStream_t st;
    long long* nblocks_out;
    ImgView<T> TS;      // time-scale map with its one-plane z halos (tile kernel)
    unsigned* counter; double* d_out;      // tile kernel: arrival counter and destination of the TV value (the last CTA finishes the sum)
};

// tv_tile.cu
template <typename T> int run_tv_tile(int vec, int scheme, bool z_on, bool t_on, const TvArgs<T>& a);
// tv_tile.cu / tv_tile_norms.cu: the kernel sets without / with the norms output
template <typename T> int run_tv_tile_plain(int vec, int scheme, bool z_on, bool t_on, const TvArgs<T>& a);
template <typename T> int run_tv_tile_norms(int vec, int scheme, bool z_on, bool t_on, const TvArgs<T>& a);

}  // namespace pytvb
