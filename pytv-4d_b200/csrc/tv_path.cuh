// Which implementation pytvb_tv runs for a problem: the single-sweep tile kernel (kernels_tile.cuh) or the two-sweep
// fallback (inverse norms through a workspace; kernels2.cuh).  Shared by tv.cu (launch) and api.cu (workspace size).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "host_common.cuh"
#include "kernels_tile.cuh"

namespace pytvb {

// PYTVB_TV_PATH=sweeps forces the fallback (A/B measurements); read once per process.
inline bool tv_force_sweeps() {
    static const bool v = [] { const char* e = getenv("PYTVB_TV_PATH"); return e && strcmp(e, "sweeps") == 0; }();
    return v;
}

inline bool tv_uses_tile(const pytvb_problem* pb) {
    if (tv_force_sweeps()) return false;
    const Axes ax = axes_of(pb);
    // the centred scheme on a length-2 axis degrades to forward differences there (tv_operators_CPU.py:339,347): fallback only
    if (pb->scheme == PYTVB_CENTRAL && ((ax.z_on && pb->Nz_global == 2) || (ax.t_on && pb->M == 2))) return false;
    TileGeom g;
    const bool mask = ax.t_on && pb->mask_static;
    if (pb->dtype == PYTVB_F32) return make_tile_geom<float, 4, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask);
    return make_tile_geom<double, 2, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask);
}

// Most CTAs (= TV partial sums) the tile kernel can launch for this problem, over its vector and scalar forms.
inline long long tile_max_blocks(const pytvb_problem* pb) {
    const Axes ax = axes_of(pb);
    const bool mask = ax.t_on && pb->mask_static;
    const int Nz = (int)pb->Nz, M = (int)pb->M, Ni = (int)pb->Ni, Nj = (int)pb->Nj;
    TileGeom g;
    long long n = 0;
    if (pb->dtype == PYTVB_F32) {
        if (make_tile_geom<float, 4, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask) && g.nblocks > n) n = g.nblocks;
        if (make_tile_geom<float, 1, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask) && g.nblocks > n) n = g.nblocks;
    } else {
        if (make_tile_geom<double, 2, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask) && g.nblocks > n) n = g.nblocks;
        if (make_tile_geom<double, 1, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask) && g.nblocks > n) n = g.nblocks;
    }
    return n;
}

}  // namespace pytvb
