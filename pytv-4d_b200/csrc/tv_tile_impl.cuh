// tv_<scheme>, single-sweep form: launch of the z-marching tile kernel (kernels_tile.cuh / tile_core.cuh).  Included by four translation
// units that compile in parallel (the kernels are the bulk of the library's build time): tv_tile.cu / tv_tile_f64.cu (kernels
// without the norms output, float / double) and tv_tile_norms.cu / tv_tile_norms_f64.cu (with it).
#pragma once
#include "host_common.cuh"
#include "tv_path.cuh"
#include "tv_args.cuh"
#include "tmap.cuh"

using namespace pytvb;

namespace {

// ---- single-sweep tile kernel (kernels_tile.cuh): one launch, x read once, G written once
template <typename T, int VEC, int SCHEME, bool Z, bool TT, int TSMODE, int FORM>
auto tile_kernel_ptr() {
    if constexpr (FORM == 2) return tv_tile2_kernel<T, VEC, SCHEME, Z, TT, PYTVB_TILE_R, TSMODE, PYTVB_TILE_NORMS>;
    else return tv_tile_kernel<T, VEC, SCHEME, Z, TT, PYTVB_TILE_R, TSMODE, PYTVB_TILE_NORMS>;
}
template <typename T, int VEC, int SCHEME, bool Z, bool TT, int TSMODE, int FORM>
int launch_tile(const TvArgs<T>& a, const TileGeom& g, size_t smem) {
    auto kern = tile_kernel_ptr<T, VEC, SCHEME, Z, TT, TSMODE, FORM>();
    static bool attr_set = false;      // per instantiation; the attribute is sticky for the function
    if (!attr_set) {
        PYTVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_LIMIT));
        attr_set = true;
    }
    // tensor maps of the image and of its two halo-plane buffers (vector path: staged by TMA; the scalar path copies per thread)
    CUtensorMap mx, mlo, mhi;
    const bool tma = VEC > 1;
    const int rc0 = make_image_tmap<T>(&mx, tma ? a.X.base : nullptr, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, g.FC, g.rowsX, g.pitchX);
    if (rc0 != PYTVB_OK) return rc0;
    const int rc1 = make_image_tmap<T>(&mlo, tma ? a.X.lo : nullptr, a.X.depth, a.P.M, a.P.Ni, a.P.Nj, g.FC, g.rowsX, g.pitchX);
    if (rc1 != PYTVB_OK) return rc1;
    const int rc2 = make_image_tmap<T>(&mhi, tma ? a.X.hi : nullptr, a.X.depth, a.P.M, a.P.Ni, a.P.Nj, g.FC, g.rowsX, g.pitchX);
    if (rc2 != PYTVB_OK) return rc2;
    kern<<<(unsigned)g.nblocks, g.nthreads, smem, a.st>>>(a.X, a.TS, a.G, a.norms, a.partial, a.counter, a.d_out, a.P, g, mx, mlo, mhi);
    count_launches(1);
    PYTVB_CUDA(cudaGetLastError());
    *a.nblocks_out = g.nblocks;
    return PYTVB_OK;
}
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchTvTile {
    static int run(const TvArgs<T>& a) {
        TileGeom g;
        const bool mask = TT && a.P.mask_static;
        const int form = pick_tile_form<T, VEC, PYTVB_TILE_R>(g, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, TT, mask, tile_form_forced());
        PYTVB_REQUIRE(form != 0, "internal: the tile kernel does not take this problem");
        PYTVB_REQUIRE(g.nblocks > 0 && g.nblocks < 2147483647LL, "grid of %lld CTAs is out of range", g.nblocks);
        const size_t smem = tile_smem_bytes<T>(g, mask);
        if (form == 2) {
            if (TT && a.P.tscale) return launch_tile<T, VEC, SCHEME, Z, TT, TT ? 2 : 0, 2>(a, g, smem);
            if (mask) return launch_tile<T, VEC, SCHEME, Z, TT, TT ? 1 : 0, 2>(a, g, smem);
            return launch_tile<T, VEC, SCHEME, Z, TT, 0, 2>(a, g, smem);
        }
        if (TT && a.P.tscale) return launch_tile<T, VEC, SCHEME, Z, TT, TT ? 2 : 0, 1>(a, g, smem);
        if (mask) return launch_tile<T, VEC, SCHEME, Z, TT, TT ? 1 : 0, 1>(a, g, smem);
        return launch_tile<T, VEC, SCHEME, Z, TT, 0, 1>(a, g, smem);
    }
};

}  // namespace

namespace pytvb {
template <typename T> int PYTVB_TILE_ENTRY(int vec, int scheme, bool z_on, bool t_on, const TvArgs<T>& a) {
    return dispatch<LaunchTvTile, T>(vec, scheme, z_on, t_on, a);
}
template int PYTVB_TILE_ENTRY<PYTVB_TILE_T>(int, int, bool, bool, const TvArgs<PYTVB_TILE_T>&);
}  // namespace pytvb
