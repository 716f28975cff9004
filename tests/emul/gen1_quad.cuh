// TEST INFRASTRUCTURE ONLY.  Generation-1 per-quad code (one quad per thread, exact IEEE division / sqrt, a boundary
// predicate per neighbour): the first implementation of the hot path, retired from libpytv_b200.so in round 2 and kept
// here as an independently written second statement of the arithmetic for the host emulation tests (tests/emul/emul.cu).
#pragma once
#include "../../pytv-4d_b200/csrc/core.cuh"

namespace pytvb {

// ------------------------------------------------------------------------------------------------
// Neighbourhood of a quad in an image.
template <typename T, int VEC>
struct Nbhd {
    T c[VEC + 2];   // c[0] = x[j0-1], c[1..VEC] = the quad, c[VEC+1] = x[j0+VEC]  (0 where out of range)
    T up[VEC], dn[VEC], zm[VEC], zp[VEC], tm[VEC], tp[VEC];
    bool v_up, v_dn, v_zm, v_zp, v_tm, v_tp;   // neighbour exists in the global domain
};

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
PYTVB_HD void load_nbhd(Nbhd<T, VEC>& n, const ImgView<T>& X, const Params<T>& P, int z, int t, int i, int j0) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    const T* r = X.row(P, z, t, i);
    ld_into<T, VEC>(n.c + 1, r + j0);
    n.c[0] = (C::NEED_BWD && j0 > 0) ? r[j0 - 1] : T(0);
    n.c[VEC + 1] = (C::NEED_FWD && j0 + VEC < P.Nj) ? r[j0 + VEC] : T(0);
    const long long zg = P.zg0 + z;
    n.v_up = i > 0;
    n.v_dn = i < P.Ni - 1;
    n.v_zm = zg > 0;
    n.v_zp = zg < P.NzG - 1;
    n.v_tm = t > 0;
    n.v_tp = t < P.M - 1;
    if (C::NEED_BWD && n.v_up) ld_into<T, VEC>(n.up, X.row(P, z, t, i - 1) + j0); else zero_into<T, VEC>(n.up);
    if (C::NEED_FWD && n.v_dn) ld_into<T, VEC>(n.dn, X.row(P, z, t, i + 1) + j0); else zero_into<T, VEC>(n.dn);
    if (Z_ON) {
        if (C::NEED_BWD && n.v_zm) ld_into<T, VEC>(n.zm, X.row(P, z - 1, t, i) + j0); else zero_into<T, VEC>(n.zm);
        if (C::NEED_FWD && n.v_zp) ld_into<T, VEC>(n.zp, X.row(P, z + 1, t, i) + j0); else zero_into<T, VEC>(n.zp);
    }
    if (T_ON) {
        if (C::NEED_BWD && n.v_tm) ld_into<T, VEC>(n.tm, X.row(P, z, t - 1, i) + j0); else zero_into<T, VEC>(n.tm);
        if (C::NEED_FWD && n.v_tp) ld_into<T, VEC>(n.tp, X.row(P, z, t + 1, i) + j0); else zero_into<T, VEC>(n.tp);
    }
}

// One axis of a single-component scheme given the two neighbours; `fallback` = central on a length-2 axis.
template <typename T, int SCHEME>
PYTVB_HD T axis_diff(T c, T minus, T plus, bool v_m, bool v_p, bool fallback) {
    if (SCHEME == UPWIND) return v_p ? plus - c : T(0);
    if (SCHEME == DOWNWIND) return v_m ? c - minus : T(0);
    if (fallback) return v_p ? plus - c : T(0);
    return (v_m && v_p) ? plus - minus : T(0);
}

// Global divisor of a scheme.  EXACT: a true division, to stay closest to the reference's
// `D_img/np.sqrt(2.0)` (operator kernels); otherwise a multiply by the reciprocal (fused kernels, where
// eight IEEE divisions per voxel would be a third of the instruction stream).
template <typename T, bool EXACT>
PYTVB_HD T apply_div(T v, const Params<T>& P) { return EXACT ? v / P.div : v * P.inv_div; }

// D_scheme at a quad: d[comp][e], the reference's operator output (weights, mask_static and the global
// divisor applied).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, bool EXACT = true>
PYTVB_HD void diffs_from_nbhd(T (*d)[VEC], const Nbhd<T, VEC>& n, const Params<T>& P, int i, int j0) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    T fac[VEC];
    if (T_ON) static_factor<T, VEC>(fac, P, i, j0);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const int j = j0 + e;
        const T c = n.c[e + 1];
        const bool v_l = j > 0, v_r = j < P.Nj - 1;
        if (SCHEME == HYBRID) {
            d[C::I_F][e] = apply_div<T, EXACT>((n.v_dn ? n.dn[e] - c : T(0)), P);
            d[C::J_F][e] = apply_div<T, EXACT>((v_r ? n.c[e + 2] - c : T(0)), P);
            d[C::I_B][e] = apply_div<T, EXACT>((n.v_up ? c - n.up[e] : T(0)), P);
            d[C::J_B][e] = apply_div<T, EXACT>((v_l ? c - n.c[e] : T(0)), P);
            if (Z_ON) {
                d[C::Z_F][e] = apply_div<T, EXACT>((n.v_zp ? P.srz * (n.zp[e] - c) : T(0)), P);
                d[C::Z_B][e] = apply_div<T, EXACT>((n.v_zm ? P.srz * (c - n.zm[e]) : T(0)), P);
            }
            if (T_ON) {
                d[C::T_F][e] = apply_div<T, EXACT>((n.v_tp ? P.srt * (n.tp[e] - c) * fac[e] : T(0)), P);
                d[C::T_B][e] = apply_div<T, EXACT>((n.v_tm ? P.srt * (c - n.tm[e]) * fac[e] : T(0)), P);
            }
        } else {
            d[C::I_F][e] = apply_div<T, EXACT>(axis_diff<T, SCHEME>(c, n.up[e], n.dn[e], n.v_up, n.v_dn, false), P);
            d[C::J_F][e] = apply_div<T, EXACT>(axis_diff<T, SCHEME>(c, n.c[e], n.c[e + 2], v_l, v_r, false), P);
            if (Z_ON)
                d[C::Z_F][e] = apply_div<T, EXACT>(P.srz * axis_diff<T, SCHEME>(c, n.zm[e], n.zp[e], n.v_zm, n.v_zp, P.z_fwd_fallback != 0), P);
            if (T_ON)
                d[C::T_F][e] = apply_div<T, EXACT>(P.srt * axis_diff<T, SCHEME>(c, n.tm[e], n.tp[e], n.v_tm, n.v_tp, P.t_fwd_fallback != 0) * fac[e], P);
        }
    }
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, bool EXACT = true>
PYTVB_HD void quad_D(T (*d)[VEC], const ImgView<T>& X, const Params<T>& P, int z, int t, int i, int j0) {
    Nbhd<T, VEC> n;
    load_nbhd<T, VEC, SCHEME, Z_ON, T_ON>(n, X, P, z, t, i, j0);
    diffs_from_nbhd<T, VEC, SCHEME, Z_ON, T_ON, EXACT>(d, n, P, i, j0);
}

// 2-norm over the components at each voxel of the quad.
template <typename T, int VEC, int ND>
PYTVB_HD void quad_norm(T* nrm, const T (*d)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < ND; ++k) s += d[k][e] * d[k][e];
        nrm[e] = pytvb_sqrt(s);
    }
}

// ------------------------------------------------------------------------------------------------
// Adjoint.  KIND: 0 forward-type component, 1 backward-type, 2 centred (tv_operators_CPU.py:555-560,
// :488-493, :623-628).  `k` is the index along the axis, `L` its global length; pm / pc / pp are the
// component at k-1, k, k+1 (only the ones the kind needs are read by the callers).
template <typename T>
PYTVB_HD T adj_fwd(long long k, long long L, T pm, T pc) { return (k > 0 ? pm : T(0)) - (k < L - 1 ? pc : T(0)); }
template <typename T>
PYTVB_HD T adj_bwd(long long k, long long L, T pc, T pp) { return (k > 0 ? pc : T(0)) - (k < L - 1 ? pp : T(0)); }
template <typename T>
PYTVB_HD T adj_ctr(long long k, long long L, T pm, T pp) {
    return ((k >= 2) ? pm : T(0)) - ((k <= L - 3) ? pp : T(0));   // p[k-1] counts iff 1<=k-1<=L-2 (k<=L-1 always)
}

// Adjoint of one in-plane / z / t axis for a pack: the callers pass row pointers of the component at the
// neighbouring index along that axis (null when not needed / out of the domain).
template <typename T, int VEC, int SCHEME>
PYTVB_HD void adj_axis_rows(T* acc, T w, long long k, long long L, bool fallback,
                            const T* rf_m, const T* rf_c,      // forward-type slot at k-1, k
                            const T* rb_c, const T* rb_p) {    // backward-type slot at k, k+1
    // For the single-component schemes rf_* and rb_* point into the same component.
    if (SCHEME == UPWIND || SCHEME == HYBRID || (SCHEME == CENTRAL && fallback)) {
        if (k > 0) {
            const Pack<T, VEC> a = ld_pack<T, VEC>(rf_m);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += w * a.v[e];
        }
        if (k < L - 1) {
            const Pack<T, VEC> b = ld_pack<T, VEC>(rf_c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] -= w * b.v[e];
        }
    }
    if (SCHEME == DOWNWIND || SCHEME == HYBRID) {
        if (k > 0) {
            const Pack<T, VEC> a = ld_pack<T, VEC>(rb_c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += w * a.v[e];
        }
        if (k < L - 1) {
            const Pack<T, VEC> b = ld_pack<T, VEC>(rb_p);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] -= w * b.v[e];
        }
    }
    if (SCHEME == CENTRAL && !fallback) {
        if (k >= 2) {
            const Pack<T, VEC> a = ld_pack<T, VEC>(rf_m);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += w * a.v[e];
        }
        if (k <= L - 3) {
            const Pack<T, VEC> b = ld_pack<T, VEC>(rb_p);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] -= w * b.v[e];
        }
    }
}

// D_T_scheme at a quad: out[e] (weights, mask_static on the time part, global divisor).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, bool EXACT = true>
PYTVB_HD void quad_DT(T* out, const FieldView<T>& Pf, const Params<T>& P, int z, int t, int i, int j0) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    T acc[VEC];
    zero_into<T, VEC>(acc);
    // rows
    adj_axis_rows<T, VEC, SCHEME>(acc, T(1), i, P.Ni, false,
                                  Pf.row(P, z, C::I_F, t, i - 1) + j0, Pf.row(P, z, C::I_F, t, i) + j0,
                                  Pf.row(P, z, C::I_B, t, i) + j0, Pf.row(P, z, C::I_B, t, i + 1) + j0);
    // columns: one pack per slot plus the element on each side
    {
        const T* rf = Pf.row(P, z, C::J_F, t, i);
        const T* rb = Pf.row(P, z, C::J_B, t, i);
        T f[VEC + 2], b[VEC + 2];
        ld_into<T, VEC>(f + 1, rf + j0);
        f[0] = (j0 > 0) ? rf[j0 - 1] : T(0);
        f[VEC + 1] = (j0 + VEC < P.Nj) ? rf[j0 + VEC] : T(0);
        if (SCHEME == HYBRID) {
            ld_into<T, VEC>(b + 1, rb + j0);
            b[0] = T(0);
            b[VEC + 1] = (j0 + VEC < P.Nj) ? rb[j0 + VEC] : T(0);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int j = j0 + e;
            if (SCHEME == UPWIND) acc[e] += adj_fwd<T>(j, P.Nj, f[e], f[e + 1]);
            if (SCHEME == DOWNWIND) acc[e] += adj_bwd<T>(j, P.Nj, f[e + 1], f[e + 2]);
            if (SCHEME == CENTRAL) acc[e] += adj_ctr<T>(j, P.Nj, f[e], f[e + 2]);
            if (SCHEME == HYBRID) acc[e] += adj_fwd<T>(j, P.Nj, f[e], f[e + 1]) + adj_bwd<T>(j, P.Nj, b[e + 1], b[e + 2]);
        }
    }
    if (Z_ON) {
        const long long zg = P.zg0 + z;
        adj_axis_rows<T, VEC, SCHEME>(acc, P.srz, zg, P.NzG, P.z_fwd_fallback != 0,
                                      Pf.row(P, z - 1, C::Z_F, t, i) + j0, Pf.row(P, z, C::Z_F, t, i) + j0,
                                      Pf.row(P, z, C::Z_B, t, i) + j0, Pf.row(P, z + 1, C::Z_B, t, i) + j0);
    }
    if (T_ON) {
        T tacc[VEC], fac[VEC];
        zero_into<T, VEC>(tacc);
        adj_axis_rows<T, VEC, SCHEME>(tacc, P.srt, t, P.M, P.t_fwd_fallback != 0,
                                      Pf.row(P, z, C::T_F, t - 1, i) + j0, Pf.row(P, z, C::T_F, t, i) + j0,
                                      Pf.row(P, z, C::T_B, t, i) + j0, Pf.row(P, z, C::T_B, t + 1, i) + j0);
        static_factor<T, VEC>(fac, P, i, j0);   // tv_operators_CPU.py:577-581
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += tacc[e] * fac[e];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) out[e] = apply_div<T, EXACT>(acc[e], P);
}

// ------------------------------------------------------------------------------------------------
// Sub-gradient from the image and the inverse-norm field w = 1/|D x| (0 where the norm is 0):
//   G = D_T_unit( D_w(x) * w )            (tv_CPU.py:92-124, :176-188, :239-251, :302-328)
// Along one axis with weight a (the weight D applied; the adjoint applies none):
//   forward slot :  a*[(x[k]-x[k-1]) w[k-1] [k>0]  -  (x[k+1]-x[k]) w[k]   [k<L-1]]
//   backward slot:  a*[(x[k]-x[k-1]) w[k]   [k>0]  -  (x[k+1]-x[k]) w[k+1] [k<L-1]]
//   centred      :  a*[(x[k]-x[k-2]) w[k-1] [k>=2] -  (x[k+2]-x[k]) w[k+1] [k<=L-3]]
// then divided by div (from D) and by div again (tv_CPU.py:124 `G /= sqrt(2)`, :328 `G /= 2`).
template <typename T, int SCHEME>
PYTVB_HD T g_axis(T a, long long k, long long L, bool fallback, T xm2, T xm, T xc, T xp, T xp2, T wm, T wc, T wp) {
    T g = T(0);
    if (SCHEME == UPWIND || (SCHEME == CENTRAL && fallback)) {
        if (k > 0) g += (xc - xm) * wm;
        if (k < L - 1) g -= (xp - xc) * wc;
    } else if (SCHEME == DOWNWIND) {
        if (k > 0) g += (xc - xm) * wc;
        if (k < L - 1) g -= (xp - xc) * wp;
    } else if (SCHEME == HYBRID) {
        if (k > 0) g += (xc - xm) * (wm + wc);
        if (k < L - 1) g -= (xp - xc) * (wc + wp);
    } else {
        if (k >= 2) g += (xc - xm2) * wm;
        if (k <= L - 3) g -= (xp2 - xc) * wp;
    }
    return a * g;
}

// Load a pack at (z,t,i) or zeros when `valid` is false.
template <typename T, int VEC>
PYTVB_HD void ld_row_or_zero(T* dst, const ImgView<T>& V, const Params<T>& P, bool valid, int z, int t, int i, int j0) {
    if (valid) ld_into<T, VEC>(dst, V.row(P, z, t, i) + j0); else zero_into<T, VEC>(dst);
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
PYTVB_HD void quad_G(T* g, const ImgView<T>& X, const ImgView<T>& W, const Params<T>& P, int z, int t, int i, int j0) {
    constexpr bool CEN = (SCHEME == CENTRAL);
    constexpr int H = CEN ? 2 : 1;   // reach of x along the axis
    T fac[VEC];
    if (T_ON) static_factor<T, VEC>(fac, P, i, j0);
    // ---- columns: x and w with H elements on each side
    T xr[VEC + 4], wr[VEC + 4];   // index e+2 is element e
    {
        const T* xrow = X.row(P, z, t, i);
        const T* wrow = W.row(P, z, t, i);
        ld_into<T, VEC>(xr + 2, xrow + j0);
        ld_into<T, VEC>(wr + 2, wrow + j0);
#pragma unroll
        for (int h = 1; h <= 2; ++h) {
            const bool vl = (h <= H) && (j0 - h >= 0), vr = (h <= H) && (j0 + VEC - 1 + h < P.Nj);
            xr[2 - h] = vl ? xrow[j0 - h] : T(0);
            wr[2 - h] = vl ? wrow[j0 - h] : T(0);
            xr[VEC + 1 + h] = vr ? xrow[j0 + VEC - 1 + h] : T(0);
            wr[VEC + 1 + h] = vr ? wrow[j0 + VEC - 1 + h] : T(0);
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        g[e] = g_axis<T, SCHEME>(T(1), j0 + e, P.Nj, false, xr[e], xr[e + 1], xr[e + 2], xr[e + 3], xr[e + 4], wr[e + 1], wr[e + 2], wr[e + 3]);
    // ---- rows, z, t: packs at the neighbouring index along the axis
    T xm2[VEC], xm[VEC], xp[VEC], xp2[VEC], wm[VEC], wp[VEC];
    {
        ld_row_or_zero<T, VEC>(xm, X, P, i >= 1, z, t, i - 1, j0);
        ld_row_or_zero<T, VEC>(xp, X, P, i <= P.Ni - 2, z, t, i + 1, j0);
        ld_row_or_zero<T, VEC>(wm, W, P, i >= 1, z, t, i - 1, j0);
        ld_row_or_zero<T, VEC>(wp, W, P, i <= P.Ni - 2, z, t, i + 1, j0);
        ld_row_or_zero<T, VEC>(xm2, X, P, CEN && i >= 2, z, t, i - 2, j0);
        ld_row_or_zero<T, VEC>(xp2, X, P, CEN && i <= P.Ni - 3, z, t, i + 2, j0);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            g[e] += g_axis<T, SCHEME>(T(1), i, P.Ni, false, xm2[e], xm[e], xr[e + 2], xp[e], xp2[e], wm[e], wr[e + 2], wp[e]);
    }
    if (Z_ON) {
        const long long zg = P.zg0 + z;
        const bool fb = P.z_fwd_fallback != 0;
        ld_row_or_zero<T, VEC>(xm, X, P, zg >= 1, z - 1, t, i, j0);
        ld_row_or_zero<T, VEC>(xp, X, P, zg <= P.NzG - 2, z + 1, t, i, j0);
        ld_row_or_zero<T, VEC>(wm, W, P, zg >= 1, z - 1, t, i, j0);
        ld_row_or_zero<T, VEC>(wp, W, P, zg <= P.NzG - 2, z + 1, t, i, j0);
        ld_row_or_zero<T, VEC>(xm2, X, P, CEN && !fb && zg >= 2, z - 2, t, i, j0);
        ld_row_or_zero<T, VEC>(xp2, X, P, CEN && !fb && zg <= P.NzG - 3, z + 2, t, i, j0);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            g[e] += g_axis<T, SCHEME>(P.srz, zg, P.NzG, fb, xm2[e], xm[e], xr[e + 2], xp[e], xp2[e], wm[e], wr[e + 2], wp[e]);
    }
    if (T_ON) {
        const bool fb = P.t_fwd_fallback != 0;
        ld_row_or_zero<T, VEC>(xm, X, P, t >= 1, z, t - 1, i, j0);
        ld_row_or_zero<T, VEC>(xp, X, P, t <= P.M - 2, z, t + 1, i, j0);
        ld_row_or_zero<T, VEC>(wm, W, P, t >= 1, z, t - 1, i, j0);
        ld_row_or_zero<T, VEC>(wp, W, P, t <= P.M - 2, z, t + 1, i, j0);
        ld_row_or_zero<T, VEC>(xm2, X, P, CEN && !fb && t >= 2, z, t - 2, i, j0);
        ld_row_or_zero<T, VEC>(xp2, X, P, CEN && !fb && t <= P.M - 3, z, t + 2, i, j0);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            g[e] += g_axis<T, SCHEME>(P.srt * fac[e], t, P.M, fb, xm2[e], xm[e], xr[e + 2], xp[e], xp2[e], wm[e], wr[e + 2], wp[e]);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) g[e] = g[e] * P.inv_div * P.inv_div;
}

// ------------------------------------------------------------------------------------------------
// Chambolle-Pock, dual pass at a quad:  y <- proj_{|.|_2 <= lam}( y + sigma * D(xbar) )   (README.md:149-151)
// y is updated in place (each quad touches only its own entries).  Returns sum over the quad of |D xbar|_2.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
PYTVB_HD T quad_cp_dual(T* y, const ImgView<T>& Xb, const Params<T>& P, T sigma, T inv_lam, int z, int t, int i, int j0) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    constexpr int ND = C::ND;
    T d[ND][VEC], yn[ND][VEC];
    T* yq = y + (long long)z * P.sZf + (long long)t * P.sT + (long long)i * P.Nj + j0;
#pragma unroll
    for (int k = 0; k < ND; ++k) ld_into<T, VEC>(yn[k], yq + (long long)k * P.sC);
    quad_D<T, VEC, SCHEME, Z_ON, T_ON, false>(d, Xb, P, z, t, i, j0);
    T l21 = T(0);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        T s = T(0), sd = T(0);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            sd += d[k][e] * d[k][e];
            yn[k][e] += sigma * d[k][e];
            s += yn[k][e] * yn[k][e];
        }
        l21 += pytvb_sqrt(sd);
        const T nrm = pytvb_sqrt(s) * inv_lam;
        const T scale = T(1) / (nrm > T(1) ? nrm : T(1));
#pragma unroll
        for (int k = 0; k < ND; ++k) yn[k][e] *= scale;
    }
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        Pack<T, VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.v[e] = yn[k][e];
        st_pack<T, VEC>(yq + (long long)k * P.sC, o);
    }
    return l21;
}

// Primal pass, ROF form:  x+ = (x - tau D^T y + tau x0) / (1 + tau);  xbar = x+ + theta (x+ - x).
// Returns sum over the quad of (x+ - x0)^2.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
PYTVB_HD T quad_cp_primal_rof(T* x, T* xbar, const T* x0, const FieldView<T>& Y, const Params<T>& P, T tau, T theta,
                              int z, int t, int i, int j0) {
    T dty[VEC];
    quad_DT<T, VEC, SCHEME, Z_ON, T_ON, false>(dty, Y, P, z, t, i, j0);
    const long long off = (long long)z * P.sZ + (long long)t * P.sT + (long long)i * P.Nj + j0;
    const Pack<T, VEC> xo = ld_pack<T, VEC>(x + off), x0q = ld_pack<T, VEC>(x0 + off);
    Pack<T, VEC> xn, xb;
    T fid = T(0);
    const T inv = T(1) / (T(1) + tau);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        xn.v[e] = (xo.v[e] - tau * dty[e] + tau * x0q.v[e]) * inv;
        xb.v[e] = xn.v[e] + theta * (xn.v[e] - xo.v[e]);
        const T r = xn.v[e] - x0q.v[e];
        fid += r * r;
    }
    st_pack<T, VEC>(x + off, xn);
    st_pack<T, VEC>(xbar + off, xb);
    return fid;
}

// Primal pass, README form (README.md:148,154):  y_f <- (y_f + sigma_A (x - x0)) / (1 + sigma_A);
// x <- x - tau y_f - tau D^T y_tv.   Returns sum over the quad of (x+ - x0)^2.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
PYTVB_HD T quad_cp_primal_readme(T* x, T* y_f, const T* x0, const FieldView<T>& Y, const Params<T>& P, T tau, T sigma_A,
                                 int z, int t, int i, int j0) {
    T dty[VEC];
    quad_DT<T, VEC, SCHEME, Z_ON, T_ON, false>(dty, Y, P, z, t, i, j0);
    const long long off = (long long)z * P.sZ + (long long)t * P.sT + (long long)i * P.Nj + j0;
    const Pack<T, VEC> xo = ld_pack<T, VEC>(x + off), x0q = ld_pack<T, VEC>(x0 + off), yfo = ld_pack<T, VEC>(y_f + off);
    Pack<T, VEC> xn, yfn;
    T fid = T(0);
    const T inv = T(1) / (T(1) + sigma_A);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        yfn.v[e] = (yfo.v[e] + sigma_A * (xo.v[e] - x0q.v[e])) * inv;
        xn.v[e] = xo.v[e] - tau * yfn.v[e] - tau * dty[e];
        const T r = xn.v[e] - x0q.v[e];
        fid += r * r;
    }
    st_pack<T, VEC>(x + off, xn);
    st_pack<T, VEC>(y_f + off, yfn);
    return fid;
}

}  // namespace pytvb
