// TEST INFRASTRUCTURE ONLY - never linked into libpytv_b200.so.
// Runs the per-quad device functions of pytv-4d_b200/csrc/strip_core.cuh (and the retired generation-1 code kept in
// gen1_quad.cuh as a second statement of the arithmetic) as HOST code, quad by quad, so that
// the index / boundary / halo / weighting logic of the CUDA kernels can be checked against the oracle in a
// container that has no GPU.  Pointers in the problem descriptor and all arrays are HOST pointers here.
#include <stdarg.h>
#include <vector>

#include "../../pytv-4d_b200/csrc/host_common.cuh"
#include "../../pytv-4d_b200/csrc/strip_core.cuh"
#include "gen1_quad.cuh"
#include "../../pytv-4d_b200/csrc/tv_path.cuh"

namespace pytvb {
static char g_err[512];
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pytvb
using namespace pytvb;

namespace {

template <typename F>
void for_each_quad(int z_lo, int nz, int M, int Ni, int Nj, int vec, F f) {
    for (int z = z_lo; z < z_lo + nz; ++z)
        for (int t = 0; t < M; ++t)
            for (int i = 0; i < Ni; ++i)
                for (int j0 = 0; j0 < Nj; j0 += vec) f(z, t, i, j0);
}

template <typename T> struct EArgs {
    Params<T> P; ImgView<T> X; ImgView<T> W; FieldView<T> F; T* out; T* out2; T* aux; const T* x0; T c0, c1; int z_lo, nz; int variant; double* sum;
};

template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct ED {
    static int run(const EArgs<T>& a) {
        typedef Comp<SCHEME, Z, TT> C;
        for_each_quad(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            T d[C::ND][VEC];
            quad_D<T, VEC, SCHEME, Z, TT>(d, a.X, a.P, z, t, i, j0);
            for (int k = 0; k < C::ND; ++k)
                for (int e = 0; e < VEC; ++e)
                    a.out[(long long)z * a.P.sZf + (long long)k * a.P.sC + (long long)t * a.P.sT + (long long)i * a.P.Nj + j0 + e] = d[k][e];
        });
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDT {
    static int run(const EArgs<T>& a) {
        for_each_quad(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            T o[VEC];
            quad_DT<T, VEC, SCHEME, Z, TT>(o, a.F, a.P, z, t, i, j0);
            for (int e = 0; e < VEC; ++e) a.out[(long long)z * a.P.sZ + (long long)t * a.P.sT + (long long)i * a.P.Nj + j0 + e] = o[e];
        });
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct ETV {
    static int run(const EArgs<T>& a) {
        typedef Comp<SCHEME, Z, TT> C;
        T* Wz0 = const_cast<T*>(a.W.base);
        double tv = 0;
        for_each_quad(a.z_lo, a.nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            T d[C::ND][VEC], nr[VEC];
            quad_D<T, VEC, SCHEME, Z, TT>(d, a.X, a.P, z, t, i, j0);
            quad_norm<T, VEC, C::ND>(nr, d);
            const long long off = (long long)z * a.P.sZ + (long long)t * a.P.sT + (long long)i * a.P.Nj + j0;
            for (int e = 0; e < VEC; ++e) {
                Wz0[off + e] = nr[e] > T(0) ? T(1) / nr[e] : T(0);
                if (z >= 0 && z < a.P.Nz) {
                    tv += (double)nr[e];
                    if (a.out2) a.out2[off + e] = nr[e] > T(0) ? nr[e] : T(INFINITY);
                }
            }
        });
        for_each_quad(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            T g[VEC];
            quad_G<T, VEC, SCHEME, Z, TT>(g, a.X, a.W, a.P, z, t, i, j0);
            for (int e = 0; e < VEC; ++e) a.out[(long long)z * a.P.sZ + (long long)t * a.P.sT + (long long)i * a.P.Nj + j0 + e] = g[e];
        });
        *a.sum = tv;
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDual {
    static int run(const EArgs<T>& a) {
        double s = 0;
        for_each_quad(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            s += (double)quad_cp_dual<T, VEC, SCHEME, Z, TT>(a.out, a.X, a.P, a.c0, a.c1, z, t, i, j0);
        });
        *a.sum = s;
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EPrimal {
    static int run(const EArgs<T>& a) {
        double s = 0;
        for_each_quad(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0) {
            if (a.variant == 0)
                s += (double)quad_cp_primal_rof<T, VEC, SCHEME, Z, TT>(a.out, a.aux, a.x0, a.F, a.P, a.c0, a.c1, z, t, i, j0);
            else
                s += (double)quad_cp_primal_readme<T, VEC, SCHEME, Z, TT>(a.out, a.aux, a.x0, a.F, a.P, a.c0, a.c1, z, t, i, j0);
        });
        *a.sum = s;
        return 0;
    }
};

// generation-2 strip code, walked row by row exactly like cp_dual_strip_kernel / cp_primal_strip_kernel
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDual2 {
    static int run(const EArgs<T>& a) {
        double s = 0;
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const DualPlane<T> pl = make_dual_plane<T, SCHEME>(a.X, a.out, a.P, z, t);
                for (int i = 0; i < a.P.Ni; ++i)
                    for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                        const int o = i * a.P.Nj + j0, o_up = i > 0 ? o - a.P.Nj : o, o_dn = i < a.P.Ni - 1 ? o + a.P.Nj : o;
                        s += (double)(a.P.tscale ? strip_quad_cp_dual<T, VEC, SCHEME, Z, TT, T, true>(pl, a.P, i, j0, o, o_up, o_dn, a.c0 * a.P.inv_div, T(1) / a.c1) : strip_quad_cp_dual<T, VEC, SCHEME, Z, TT, T, false>(pl, a.P, i, j0, o, o_up, o_dn, a.c0 * a.P.inv_div, T(1) / a.c1));
                    }
            }
        *a.sum = s * (double)a.P.inv_div;
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EPrimal2 {
    static int run(const EArgs<T>& a) {
        double s = 0;
        const T c1 = T(1) / (T(1) + (a.variant == 0 ? a.c0 : a.c1));
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z, TT>(a.F, a.P, z, t);
                for (int i = 0; i < a.P.Ni; ++i)
                    for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                        const int o = i * a.P.Nj + j0, o_up = i > 0 ? o - a.P.Nj : o, o_dn = i < a.P.Ni - 1 ? o + a.P.Nj : o;
                        if (a.variant == 0)
                            s += (double)(a.P.tscale ? strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 0, false, T, true>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1) : strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 0, false, T, false>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1));
                        else
                            s += (double)(a.P.tscale ? strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 1, false, T, true>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1) : strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 1, false, T, false>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1));
                    }
            }
        *a.sum = s;
        return 0;
    }
};

template <typename F>
void for_each_quad_strip(int z_lo, int nz, int M, int Ni, int Nj, int vec, F f) {
    for (int z = z_lo; z < z_lo + nz; ++z)
        for (int t = 0; t < M; ++t)
            for (int i = 0; i < Ni; ++i)
                for (int j0 = 0; j0 < Nj; j0 += vec) {
                    const int o = i * Nj + j0;
                    f(z, t, i, j0, o, i > 0 ? o - Nj : o, i < Ni - 1 ? o + Nj : o);
                }
}
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct ED2 {
    static int run(const EArgs<T>& a) {
        for_each_quad_strip(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0, int o, int o_up, int o_dn) {
            const DualPlane<T> pl = make_dual_plane<T, SCHEME>(a.X, a.out, a.P, z, t);
            if (a.P.tscale) strip_quad_D<T, VEC, SCHEME, Z, TT, true>(pl.y, pl, a.P, i, j0, o, o_up, o_dn); else strip_quad_D<T, VEC, SCHEME, Z, TT, false>(pl.y, pl, a.P, i, j0, o, o_up, o_dn);
        });
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDT2 {
    static int run(const EArgs<T>& a) {
        for_each_quad_strip(0, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, VEC, [&](int z, int t, int i, int j0, int o, int o_up, int o_dn) {
            const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z, TT>(a.F, a.P, z, t);
            T v[VEC];
            if (a.P.tscale) strip_quad_DT<T, VEC, SCHEME, Z, TT, false, T, true>(v, pl, a.P, i, j0, o, o_up, o_dn); else strip_quad_DT<T, VEC, SCHEME, Z, TT, false, T, false>(v, pl, a.P, i, j0, o, o_up, o_dn);
            for (int e = 0; e < VEC; ++e) a.out[pl.img + o + e] = v[e];
        });
        return 0;
    }
};
// TV sweeps as the strip kernels run them: the centred scheme row by row, the other schemes in row-marching strips of
// EMUL_R rows per thread (strip_rows_tv_norm / strip_rows_G), incl. a last strip that overhangs the image.  The library
// runs R = 8 (tv.cu: PYTVB_STRIP_R); pytvb_emulate_set_rows switches the emulation between 4 and 8.
static int g_emul_rows = 8;
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct ETV2 {
    static int run(const EArgs<T>& a) { return g_emul_rows == 4 ? run_r<4>(a) : run_r<8>(a); }
    template <int EMUL_R> static int run_r(const EArgs<T>& a) {
        T* Wz0 = const_cast<T*>(a.W.base);
        double tv = 0;
        if constexpr (SCHEME == CENTRAL) {
            for (int z = a.z_lo; z < a.z_lo + a.nz; ++z)
                for (int t = 0; t < a.P.M; ++t) {
                    const DualPlane<T> pl = make_dual_plane<T, SCHEME>(a.X, Wz0, a.P, z, t);
                    const long long img = (long long)z * a.P.sZ + (long long)t * a.P.sT;
                    const bool own = z >= 0 && z < a.P.Nz;
                    T* np = (a.out2 && own) ? a.out2 + img : nullptr;
                    for (int i0 = 0; i0 < a.P.Ni; i0 += EMUL_R)
                        for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                            const T v = a.P.tscale ? strip_rows_tv_norm_central<T, VEC, Z, TT, EMUL_R, TT, TT>(Wz0 + img, np, pl, a.P, i0, j0)
                                                   : strip_rows_tv_norm_central<T, VEC, Z, TT, EMUL_R, false, TT>(Wz0 + img, np, pl, a.P, i0, j0);
                            if (own) tv += (double)v;
                        }
                }
            for (int z = 0; z < a.P.Nz; ++z)
                for (int t = 0; t < a.P.M; ++t) {
                    const GradPlane<T> pl = make_grad_plane<T, SCHEME>(a.X, a.W, a.P, z, t);
                    T* gp = a.out + (long long)z * a.P.sZ + (long long)t * a.P.sT;
                    for (int i0 = 0; i0 < a.P.Ni; i0 += EMUL_R)
                        for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                            if (a.P.tscale) strip_rows_G_central<T, VEC, Z, TT, EMUL_R, true>(gp, pl, a.P, i0, j0);
                            else strip_rows_G_central<T, VEC, Z, TT, EMUL_R, false>(gp, pl, a.P, i0, j0);
                        }
                }
        } else {
            for (int z = a.z_lo; z < a.z_lo + a.nz; ++z)
                for (int t = 0; t < a.P.M; ++t) {
                    const DualPlane<T> pl = make_dual_plane<T, SCHEME>(a.X, Wz0, a.P, z, t);
                    const long long img = (long long)z * a.P.sZ + (long long)t * a.P.sT;
                    const bool own = z >= 0 && z < a.P.Nz;
                    T* np = (a.out2 && own) ? a.out2 + img : nullptr;
                    for (int i0 = 0; i0 < a.P.Ni; i0 += EMUL_R)
                        for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                            const T v = a.P.tscale ? strip_rows_tv_norm<T, VEC, SCHEME, Z, TT, EMUL_R, true, true>(Wz0 + img, np, pl, a.P, i0, j0)
                                        : a.P.mask_static ? strip_rows_tv_norm<T, VEC, SCHEME, Z, TT, EMUL_R, false, true>(Wz0 + img, np, pl, a.P, i0, j0)
                                                          : strip_rows_tv_norm<T, VEC, SCHEME, Z, TT, EMUL_R, false, false>(Wz0 + img, np, pl, a.P, i0, j0);
                            if (own) tv += (double)v;
                        }
                }
            for (int z = 0; z < a.P.Nz; ++z)
                for (int t = 0; t < a.P.M; ++t) {
                    const GradPlane<T> pl = make_grad_plane<T, SCHEME>(a.X, a.W, a.P, z, t);
                    T* gp = a.out + (long long)z * a.P.sZ + (long long)t * a.P.sT;
                    for (int i0 = 0; i0 < a.P.Ni; i0 += EMUL_R)
                        for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                            if (a.P.tscale) strip_rows_G<T, VEC, SCHEME, Z, TT, EMUL_R, true, true>(gp, pl, a.P, i0, j0);
                            else if (a.P.mask_static) strip_rows_G<T, VEC, SCHEME, Z, TT, EMUL_R, false, true>(gp, pl, a.P, i0, j0);
                            else strip_rows_G<T, VEC, SCHEME, Z, TT, EMUL_R, false, false>(gp, pl, a.P, i0, j0);
                        }
                }
        }
        *a.sum = tv;
        return 0;
    }
};

// Single-sweep tile kernel (kernels_tile.cuh), CTA by CTA; inside a CTA the phases between two barriers are run for all
// threads one after the other, which is one valid execution of the kernel.
static int g_tile_strips = 0, g_tile_Lz = 0, g_tile_form = 0;      // test overrides of the geometry / the form (0 = what the library chooses)
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct ETile {
    static int run(const EArgs<T>& a) {
        constexpr int R = PYTVB_TILE_R;
        TileGeom g;
        const bool mask = TT && a.P.mask_static;
        const int form = pick_tile_form<T, VEC, R>(g, a.P.Nz, a.P.M, a.P.Ni, a.P.Nj, TT, mask, g_tile_form);
        if (!form) return -20;
        if (g_tile_strips > 0 && g_tile_strips < g.strips) {
            tile_set_strips<T, R>(g, g_tile_strips);
            g.nti = (a.P.Ni + g.TI - 1) / g.TI;
            g.nblocks = (long long)g.nti * g.ntj * g.nfg * g.nzc;
        }
        if (g_tile_Lz > 0) { g.Lz = g_tile_Lz; g.nzc = (a.P.Nz + g.Lz - 1) / g.Lz; g.nblocks = (long long)g.nti * g.ntj * g.nfg * g.nzc; }
        const int tsmode = (TT && a.P.tscale) ? 2 : (mask ? 1 : 0);
        if (form == 2) {
            if (tsmode == 2) return run_mode2<2>(a, g, mask);
            if (tsmode == 1) return run_mode2<1>(a, g, mask);
            return run_mode2<0>(a, g, mask);
        }
        if (tsmode == 2) return run_mode<2>(a, g, mask);
        if (tsmode == 1) return run_mode<1>(a, g, mask);
        return run_mode<0>(a, g, mask);
    }
    // one-phase form (tile2_core.cuh): threads one after the other within a step is a valid execution (a step reads the windows of
    // the planes p-1, p, p+1 and the w window of plane p-1, and writes only the other w window)
    template <int TSMODE> static int run_mode2(const EArgs<T>& a, const TileGeom& g, bool mask) {
        constexpr int R = PYTVB_TILE_R;
        constexpr int TSM = TT ? TSMODE : 0;
        std::vector<unsigned char> smem(tile_smem_bytes<T>(g, mask) + 16);
        auto stage = [&](const TileCtx<T>& c, int q, int tid) {
            const int ql = tile2_plane<T, Z>(a.P, q), sl = tile2_slot(q);
            if (VEC > 1) tile_stage_plane_zfill<T, VEC>(c, g, a.X, a.P, ql, sl, tid); else tile_stage_plane<T, VEC>(c, g, a.X, a.P, ql, sl, tid);
        };
        auto land = [&](const TileCtx<T>& c, int q, int tid) { if (VEC > 1 && c.fix) tile_fixup_plane<T, VEC>(c, g, a.P, tile2_slot(q), tid); };
        std::vector<Tile2Thread<T, VEC, R>> st(g.nthreads);
        double tv = 0;
        for (long long b = 0; b < g.nblocks; ++b) {
            std::fill(smem.begin(), smem.end(), (unsigned char)0xFF);      // NaN pattern: an unstaged element shows up in the results
            const TileCtx<T> c = tile_ctx<T, VEC>(g, b, a.P, smem.data(), mask);
            for (auto& s : st) memset(&s, 0, sizeof(s));
            std::vector<TilePos> tps(g.nthreads);
            for (int tid = 0; tid < g.nthreads; ++tid) tps[tid] = tile_pos<T, VEC, R>(c, g, a.P, tid);
            auto all = [&](auto f) { for (int tid = 0; tid < g.nthreads; ++tid) f(tid); };
            if (mask) all([&](int tid) { tile_stage_mask<T, VEC>(c, g, a.P, tid); });
            all([&](int tid) { tile_stage_tables<T>(c, g, a.P, tid); });
            const int p0 = Z ? c.zc0 - 1 : c.zc0, p1 = c.zc1;
            for (int q = p0 - 1; q <= p0 + 1; ++q) { all([&](int tid) { stage(c, q, tid); }); all([&](int tid) { land(c, q, tid); }); }
            for (int p = p0; p <= p1; ++p) {
                const bool more = p + 2 <= p1 + 1, early = (p & 1) != 0;      // the asynchronous load lands before or after the step's work
                if (more && early) all([&](int tid) { stage(c, p + 2, tid); });
                all([&](int tid) {
                    if (a.out2) tile2_step<T, VEC, SCHEME, Z, TT, R, TSM, true>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]);
                    else tile2_step<T, VEC, SCHEME, Z, TT, R, TSM, false>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]);
                });
                if (more && !early) all([&](int tid) { stage(c, p + 2, tid); });
                if (more) all([&](int tid) { land(c, p + 2, tid); });
            }
            for (auto& s : st) tv += (double)s.tv;
        }
        *a.sum = tv;
        return 0;
    }
    template <int TSMODE> static int run_mode(const EArgs<T>& a, const TileGeom& g, bool mask) {
        constexpr int R = PYTVB_TILE_R;
        constexpr int TSM = TT ? TSMODE : 0;
        std::vector<unsigned char> smem(tile_smem_bytes<T>(g, mask) + 16);
        // staging as on the device: the vector path by TMA (zero fill outside the image, then the repair of the border CTAs), the scalar
        // path by clamped per-thread copies
        auto stage = [&](const TileCtx<T>& c, int q, int tid) {
            const int ql = Z ? tile_clamp_plane(a.P, q) : q, sl = tile_slot(q);
            if (VEC > 1) tile_stage_plane_zfill<T, VEC>(c, g, a.X, a.P, ql, sl, tid); else tile_stage_plane<T, VEC>(c, g, a.X, a.P, ql, sl, tid);
        };
        auto land = [&](const TileCtx<T>& c, int q, int tid) { if (VEC > 1 && c.fix) tile_fixup_plane<T, VEC>(c, g, a.P, tile_slot(q), tid); };
        std::vector<TileThread<T, VEC, R>> st(g.nthreads);
        double tv = 0;
        for (long long b = 0; b < g.nblocks; ++b) {
            std::fill(smem.begin(), smem.end(), (unsigned char)0xFF);      // NaN pattern: an unstaged element shows up in the results
            const TileCtx<T> c = tile_ctx<T, VEC>(g, b, a.P, smem.data(), mask);
            for (auto& s : st) memset(&s, 0, sizeof(s));
            std::vector<TilePos> tps(g.nthreads);
            for (int tid = 0; tid < g.nthreads; ++tid) tps[tid] = tile_pos<T, VEC, R>(c, g, a.P, tid);
            auto all = [&](auto f) { for (int tid = 0; tid < g.nthreads; ++tid) f(tid); };
            if (mask) all([&](int tid) { tile_stage_mask<T, VEC>(c, g, a.P, tid); });
            all([&](int tid) { tile_stage_tables<T>(c, g, a.P, tid); });
            if (Z) {
                const int p0 = c.zc0 - 1, p1 = c.zc1;
                all([&](int tid) { stage(c, p0, tid); stage(c, p0 + 1, tid); });
                all([&](int tid) { land(c, p0, tid); land(c, p0 + 1, tid); });
                all([&](int tid) { tile_init_z<T, VEC, SCHEME, R>(st[tid], c, g, a.X, a.P, p0, tps[tid]); });
                for (int p = p0; p <= p1; ++p) {
                    // the staging of plane p+2 is asynchronous on the device: it may land at any time before the end of the step.
                    // Emulate the two extremes on alternating steps: before the w-phase / after the G-phase.
                    const bool early = (p & 1) != 0;
                    if (early && p + 2 <= p1 + 1) all([&](int tid) { stage(c, p + 2, tid); });
                    all([&](int tid) { if (a.out2) tile_phase_w<T, VEC, SCHEME, Z, TT, R, TSM, true>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]); else tile_phase_w<T, VEC, SCHEME, Z, TT, R, TSM, false>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]); });
                    if (p >= c.zc0 && p < c.zc1) all([&](int tid) { tile_phase_g<T, VEC, SCHEME, Z, TT, R, TSM>(st[tid], c, g, a.P, a.W, a.out, p, tps[tid]); });
                    if (!early && p + 2 <= p1 + 1) all([&](int tid) { stage(c, p + 2, tid); });
                    if (p + 2 <= p1 + 1) all([&](int tid) { land(c, p + 2, tid); });
                }
            } else {
                all([&](int tid) { stage(c, c.zc0, tid); });
                all([&](int tid) { land(c, c.zc0, tid); });
                for (int p = c.zc0; p < c.zc1; ++p) {
                    const bool early = (p & 1) != 0;
                    if (early && p + 1 < c.zc1) all([&](int tid) { stage(c, p + 1, tid); });
                    all([&](int tid) { if (a.out2) tile_phase_w<T, VEC, SCHEME, Z, TT, R, TSM, true>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]); else tile_phase_w<T, VEC, SCHEME, Z, TT, R, TSM, false>(st[tid], c, g, a.P, a.W, a.out, a.out2, p, tps[tid]); });
                    all([&](int tid) { tile_phase_g<T, VEC, SCHEME, Z, TT, R, TSM>(st[tid], c, g, a.P, a.W, a.out, p, tps[tid]); });
                    if (!early && p + 1 < c.zc1) all([&](int tid) { stage(c, p + 1, tid); });
                    if (p + 1 < c.zc1) all([&](int tid) { land(c, p + 1, tid); });
                }
            }
            for (auto& s : st) tv += (double)s.tv;
        }
        *a.sum = tv;
        return 0;
    }
};

template <typename T> int vec_for(const pytvb_problem* pb, int force_scalar) {
    return (force_scalar || pb->Nj % VecOf<T>::value) ? 1 : VecOf<T>::value;
}

template <typename T>
int run(int op, const pytvb_problem* pb, const void* in, void* out, void* out2, void* aux, const void* x0, const void* lo, const void* hi,
        double c0, double c1, int variant, int force_scalar, double* sum) {
    const Axes ax = axes_of(pb);
    EArgs<T> a;
    a.P = make_params<T>(pb);
    a.out = (T*)out; a.out2 = (T*)out2; a.aux = (T*)aux; a.x0 = (const T*)x0; a.c0 = (T)c0; a.c1 = (T)c1; a.variant = variant; a.sum = sum;
    a.z_lo = 0; a.nz = a.P.Nz;
    const int vec = vec_for<T>(pb, force_scalar);
    std::vector<T> wbuf;
    switch (op) {
        case 0: a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 1}; return dispatch<ED, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 1: a.F = FieldView<T>{(const T*)in, (const T*)lo, (const T*)hi}; return dispatch<EDT, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 7: a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 1}; return dispatch<ED2, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 8: a.F = FieldView<T>{(const T*)in, (const T*)lo, (const T*)hi}; return dispatch<EDT2, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 10:
            if (tv_tile_possible(pb)) {   // tile kernel: in = x, out = G, out2 = norms | NULL; lo / hi = two halo planes each; a.W carries the time-scale view
                a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 2};
                a.W = ImgView<T>{a.P.tscale, (const T*)pb->time_scale_lo, (const T*)pb->time_scale_hi, 1};
                return dispatch<ETile, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
            }
            op = 9;   // the problems the library sends to the two-sweep fallback
            // fall through
        case 2:
        case 9: {
            const bool has_lo = ax.z_on && pb->z_offset > 0, has_hi = ax.z_on && pb->z_offset + pb->Nz < pb->Nz_global;
            a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 2};
            wbuf.assign((size_t)(a.P.Nz + 2) * a.P.sZ, T(0));
            T* Wz0 = wbuf.data() + a.P.sZ;
            a.W = ImgView<T>{Wz0, wbuf.data(), Wz0 + (long long)a.P.Nz * a.P.sZ, 1};
            a.z_lo = has_lo ? -1 : 0;
            a.nz = a.P.Nz + (has_lo ? 1 : 0) + (has_hi ? 1 : 0);
            return op == 2 ? dispatch<ETV, T>(vec, pb->scheme, ax.z_on, ax.t_on, a) : dispatch<ETV2, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        }
        case 3: a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 1}; return dispatch<EDual, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 4: a.F = FieldView<T>{(const T*)in, (const T*)lo, (const T*)hi}; return dispatch<EPrimal, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 5: a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 1}; return dispatch<EDual2, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
        case 6: a.F = FieldView<T>{(const T*)in, (const T*)lo, (const T*)hi}; return dispatch<EPrimal2, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
    }
    return -1;
}

// peer-memory halo push: the dual / primal strip code with mirror stores (MIR = true), walked like the kernels
template <typename T> struct MArgs { EArgs<T> e; MirrorBufs<T> mb; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDualM {
    static int run(const MArgs<T>& m) {
        const EArgs<T>& a = m.e;
        double s = 0;
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const DualPlane<T> pl = make_dual_plane<T, SCHEME>(a.X, a.out, a.P, z, t);
                const MirrorPlanes<T> mir = mirror_planes<T>(m.mb, a.P, z, t);
                for_each_quad_strip(0, 1, 1, a.P.Ni, a.P.Nj, VEC, [&](int, int, int i, int j0, int o, int o_up, int o_dn) {
                    s += (double)strip_quad_cp_dual<T, VEC, SCHEME, Z, TT, T, false, true>(pl, a.P, i, j0, o, o_up, o_dn, a.c0 * a.P.inv_div, T(1) / a.c1, mir);
                });
            }
        *a.sum = s * (double)a.P.inv_div;
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EPrimalM {
    static int run(const MArgs<T>& m) {
        const EArgs<T>& a = m.e;
        double s = 0;
        const T c1 = T(1) / (T(1) + (a.variant == 0 ? a.c0 : a.c1));
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z, TT>(a.F, a.P, z, t);
                const MirrorPlanes<T> mir = mirror_planes<T>(m.mb, a.P, z, t);
                for_each_quad_strip(0, 1, 1, a.P.Ni, a.P.Nj, VEC, [&](int, int, int i, int j0, int o, int o_up, int o_dn) {
                    if (a.variant == 0)
                        s += (double)strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 0, false, T, false, true>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1, T(-1), mir);
                    else
                        s += (double)strip_quad_cp_primal<T, VEC, SCHEME, Z, TT, 1, false, T, false, true>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn, a.c0, c1, a.c1, T(-1), mir);
                });
            }
        *a.sum = s;
        return 0;
    }
};
template <typename T>
int run_mirror(int op, const pytvb_problem* pb, const void* in, void* out, void* aux, const void* x0, const void* lo, const void* hi, void* mp,
               void* mn, double c0, double c1, int variant, int force_scalar, double* sum) {
    const Axes ax = axes_of(pb);
    MArgs<T> m;
    EArgs<T>& a = m.e;
    a.P = make_params<T>(pb);
    a.out = (T*)out; a.out2 = nullptr; a.aux = (T*)aux; a.x0 = (const T*)x0; a.c0 = (T)c0; a.c1 = (T)c1; a.variant = variant; a.sum = sum;
    a.z_lo = 0; a.nz = a.P.Nz;
    m.mb = MirrorBufs<T>{(T*)mp, (T*)mn};
    const int vec = vec_for<T>(pb, force_scalar);
    if (op == 0) { a.X = ImgView<T>{(const T*)in, (const T*)lo, (const T*)hi, 1}; return dispatch<EDualM, T>(vec, pb->scheme, ax.z_on, ax.t_on, m); }
    a.F = FieldView<T>{(const T*)in, (const T*)lo, (const T*)hi};
    return dispatch<EPrimalM, T>(vec, pb->scheme, ax.z_on, ax.t_on, m);
}

}  // namespace

// op 0: dual pass (in=xbar, out=y), op 1: primal pass (in=y, out=x); mp / mn = the neighbours' halo planes or NULL
extern "C" int pytvb_emulate_mirror(int op, const pytvb_problem* pb, const void* in, void* out, void* aux, const void* x0, const void* lo,
                                    const void* hi, void* mp, void* mn, double c0, double c1, int variant, int force_scalar, double* sum) {
    if (check_problem(pb)) return -1;
    double dummy = 0;
    if (!sum) sum = &dummy;
    return pb->dtype == PYTVB_F32 ? run_mirror<float>(op, pb, in, out, aux, x0, lo, hi, mp, mn, c0, c1, variant, force_scalar, sum)
                                  : run_mirror<double>(op, pb, in, out, aux, x0, lo, hi, mp, mn, c0, c1, variant, force_scalar, sum);
}

// op: 0 D (in=x, out=D) | 1 DT (in=p, out=img) | 2 tv (in=x, out=G, out2=norms|NULL, sum=tv)
//     5 / 6: generation-2 (strip) code of 3 / 4, same arguments;  7 / 8 / 9: generation-2 code of 0 / 1 / 2
//     3 cp_dual (in=xbar, out=y in place, c0=sigma, c1=1/lam, sum=l21) | 4 cp_primal (in=y, out=x, aux, x0, c0=tau, c1=theta|sigma_A, variant)
extern "C" int pytvb_emulate(int op, const pytvb_problem* pb, const void* in, void* out, void* out2, void* aux, const void* x0, const void* lo,
                             const void* hi, double c0, double c1, int variant, int force_scalar, double* sum) {
    if (check_problem(pb)) return -1;
    double dummy = 0;
    if (!sum) sum = &dummy;
    return pb->dtype == PYTVB_F32 ? run<float>(op, pb, in, out, out2, aux, x0, lo, hi, c0, c1, variant, force_scalar, sum)
                                  : run<double>(op, pb, in, out, out2, aux, x0, lo, hi, c0, c1, variant, force_scalar, sum);
}
// half-precision dual storage (float images), walked like the strip kernels
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EDualH {
    static int run(const EArgs<float>& a) {
        double s = 0;
        __half* y = (__half*)a.out;
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const DualPlane<float, __half> pl = make_dual_plane<float, SCHEME, __half>(a.X, y, a.P, z, t);
                for (int i = 0; i < a.P.Ni; ++i)
                    for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                        const int o = i * a.P.Nj + j0, o_up = i > 0 ? o - a.P.Nj : o, o_dn = i < a.P.Ni - 1 ? o + a.P.Nj : o;
                        // c0 = sigma, c1 = 1/lam
                        s += (double)strip_quad_cp_dual<float, VEC, SCHEME, Z, TT, __half>(pl, a.P, i, j0, o, o_up, o_dn, a.c0 * a.c1 * a.P.inv_div, 1.0f);
                    }
            }
        *a.sum = s * (double)a.P.inv_div;
        return 0;
    }
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct EPrimalH {
    static int run(const EArgs<float>& a) {
        double s = 0;
        // c0 = tau, c1 = theta, variant field carries nothing; lam is passed through a.z_lo as bits (see run_h)
        const float lam = *(const float*)&a.z_lo;
        const float c1 = 1.0f / (1.0f + a.c0);
        FieldView<__half> Y{(const __half*)a.F.base, (const __half*)a.F.lo, (const __half*)a.F.hi};
        for (int z = 0; z < a.P.Nz; ++z)
            for (int t = 0; t < a.P.M; ++t) {
                const PrimalPlane<float, __half> pl = make_primal_plane<float, SCHEME, Z, TT, __half>(Y, a.P, z, t);
                for (int i = 0; i < a.P.Ni; ++i)
                    for (int j0 = 0; j0 < a.P.Nj; j0 += VEC) {
                        const int o = i * a.P.Nj + j0, o_up = i > 0 ? o - a.P.Nj : o, o_dn = i < a.P.Ni - 1 ? o + a.P.Nj : o;
                        s += (double)strip_quad_cp_primal<float, VEC, SCHEME, Z, TT, 0, false, __half>(a.out, a.aux, a.x0, pl, a.P, i, j0, o, o_up, o_dn,
                                                                                                        a.c0 * lam, c1, a.c1, a.c0);
                    }
            }
        *a.sum = s;
        return 0;
    }
};
// op 0: dual (in = xbar, y_half in place, c0 = sigma, c1 = 1/lam);  op 1: primal (in = y_half, out = x, aux = xbar, x0, c0 = tau, c1 = theta, lam)
extern "C" int pytvb_emulate_f16y(int op, const pytvb_problem* pb, const void* in, void* out, void* aux, const void* x0, double c0, double c1, double lam,
                                  int force_scalar, double* sum) {
    if (check_problem(pb) || pb->dtype != PYTVB_F32) return -1;
    const Axes ax = axes_of(pb);
    EArgs<float> a;
    a.P = make_params<float>(pb);
    a.out = (float*)out; a.aux = (float*)aux; a.x0 = (const float*)x0; a.c0 = (float)c0; a.c1 = (float)c1; a.sum = sum;
    const float lamf = (float)lam;
    a.z_lo = *(const int*)&lamf;
    const int vec = vec_for<float>(pb, force_scalar);
    if (op == 0) { a.X = ImgView<float>{(const float*)in, nullptr, nullptr, 1}; return dispatch<EDualH, float>(vec, pb->scheme, ax.z_on, ax.t_on, a); }
    a.F = FieldView<float>{(const float*)in, nullptr, nullptr};
    return dispatch<EPrimalH, float>(vec, pb->scheme, ax.z_on, ax.t_on, a);
}

extern "C" void pytvb_emulate_set_rows(int r) { g_emul_rows = (r == 4) ? 4 : 8; }
extern "C" void pytvb_emulate_set_tile(int strips, int Lz, int form) { g_tile_strips = strips; g_tile_Lz = Lz; g_tile_form = form; }
// Geometry the library picks for a float32 (VEC 4) / float64 (VEC 2) problem: out = {form, strips, RPF, TI, TJ, FC, nthreads, Lz, nzc,
// nblocks, shared-memory bytes, shared-memory limit}.  Host code only (kernels_tile.cuh::pick_tile_form).
extern "C" int pytvb_emulate_tile_geom(const pytvb_problem* pb, long long* out) {
    const Axes ax = axes_of(pb);
    const bool mask = ax.t_on && pb->mask_static;
    TileGeom g;
    int form;
    size_t smem;
    if (pb->dtype == PYTVB_F32) {
        form = pick_tile_form<float, 4, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask);
        smem = form ? tile_smem_bytes<float>(g, mask) : 0;
    } else {
        form = pick_tile_form<double, 2, PYTVB_TILE_R>(g, (int)pb->Nz, (int)pb->M, (int)pb->Ni, (int)pb->Nj, ax.t_on, mask);
        smem = form ? tile_smem_bytes<double>(g, mask) : 0;
    }
    if (!form) { out[0] = 0; return 0; }
    const long long v[12] = {form, g.strips, g.RPF, g.TI, g.TJ, g.FC, g.nthreads, g.Lz, g.nzc, g.nblocks, (long long)smem, (long long)TILE_SMEM_LIMIT};
    for (int k = 0; k < 12; ++k) out[k] = v[k];
    return 0;
}
extern "C" const char* pytvb_emulate_error(void) { return g_err; }
