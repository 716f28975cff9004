// Which implementation pytvb_tv runs for a problem: the single-sweep tile kernel (kernels_tile.cuh) or the two-sweep
// fallback (inverse norms through a workspace; kernels2.cuh).  Shared by tv.cu (launch) and api.cu (workspace size).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "host_common.cuh"
#include "kernels_tile.cuh"

namespace pytvb {

// PYTVB_TV_PATH = auto (default) | tile | sweeps; read once per process.
enum TvPathMode { TV_AUTO = 0, TV_TILE = 1, TV_SWEEPS = 2 };
inline TvPathMode tv_path_mode() {
    static const TvPathMode v = [] {
        const char* e = getenv("PYTVB_TV_PATH");
        if (e && strcmp(e, "sweeps") == 0) return TV_SWEEPS;
        if (e && strcmp(e, "tile") == 0) return TV_TILE;
        return TV_AUTO;
    }();
    return v;
}

// PYTVB_TILE_FORM = 1 | 2 forces the two-phase / one-phase form of the tile kernel where it can take the problem (A/B
// measurements); default 0: pick_tile_form chooses.  Read once per process.
inline int tile_form_forced() {
    static const int v = [] { const char* e = getenv("PYTVB_TILE_FORM"); return e ? atoi(e) : 0; }();
    return (v == 1 || v == 2) ? v : 0;
}

// Can the tile kernel take the problem at all?
inline bool tv_tile_possible(const pytvb_problem* pb) {
    const Axes ax = axes_of(pb);
    // the centred scheme on a length-2 axis degrades to forward differences there (tv_operators_CPU.py:339,347): two-sweep form only
    if (pb->scheme == PYTVB_CENTRAL && ((ax.z_on && pb->Nz_global == 2) || (ax.t_on && pb->M == 2))) return false;
    TileGeom g;
    const bool mask = ax.t_on && pb->mask_static;
    if (pb->dtype == PYTVB_F32) return pick_tile_form<float, 4, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask) != 0;
    return pick_tile_form<double, 2, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask) != 0;
}

// Which implementation runs.  Both are parity-green on every golden; the choice is measured speed (DESIGN.md 3.3,
// profiles/r02z_tv_times.txt; C4 slab, ms, tile / sweeps: hybrid 1.83 / 2.71, upwind 1.59 / 2.31, centred 1.91 / 3.54; C5 slab hybrid
// 7.5 / 11.4; 512^3 hybrid 0.49 / 0.64, upwind 0.44 / 0.62): the tile kernel wherever it can take the problem.  It moves 1.03 x the
// algorithmic 8 B/voxel (the sweeps 2.5 x), needs no workspace and is one launch.
inline bool tv_uses_tile(const pytvb_problem* pb) {
    const TvPathMode m = tv_path_mode();
    if (m == TV_SWEEPS || !tv_tile_possible(pb)) return false;
    return true;
}

// Most CTAs (= TV partial sums) the tile kernel can launch for this problem, over its vector and scalar forms.
inline long long tile_max_blocks(const pytvb_problem* pb) {
    const Axes ax = axes_of(pb);
    const bool mask = ax.t_on && pb->mask_static;
    const int Nz = (int)pb->Nz, M = (int)pb->M, Ni = (int)pb->Ni, Nj = (int)pb->Nj;
    TileGeom g;
    long long n = 0;
    for (int form = 1; form <= 2; ++form) {
        if (pb->dtype == PYTVB_F32) {
            if (make_tile_geom<float, 4, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask, form) && g.nblocks > n) n = g.nblocks;
            if (make_tile_geom<float, 1, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask, form) && g.nblocks > n) n = g.nblocks;
        } else {
            if (make_tile_geom<double, 2, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask, form) && g.nblocks > n) n = g.nblocks;
            if (make_tile_geom<double, 1, PYTVB_TILE_R>(g, Nz, M, Ni, Nj, ax.t_on, mask, form) && g.nblocks > n) n = g.nblocks;
        }
    }
    return n;
}

}  // namespace pytvb
