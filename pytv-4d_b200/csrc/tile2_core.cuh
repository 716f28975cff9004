// Single-sweep TV value + sub-gradient, ONE-PHASE form of the z-marching tile kernel (kernels_tile.cuh): per-thread code.
//
// The two-phase form (tile_core.cuh) runs, per z-plane p, a w-phase, a barrier, a G-phase and a second barrier, and hands
// w(p) and the z term of every owned voxel from the first phase to the second in registers (124 registers per thread, 16
// warps per SM, issue slots 54 % busy: profiles/r02l_*).  Here the sub-gradient lags the norms by one plane: in step p a thread
//   * computes n(p), w(p) of its R rows (tile + 1 halo ring) and publishes w(p) in one of TWO w windows,
//   * forms the in-plane and time edge terms of plane p-1 from the x window of plane p-1 and the OTHER w window (complete
//     since the barrier that ended step p-1), adds the two z terms - (p-2 -> p-1) carried in registers, (p-1 -> p) from
//     the w just computed and this thread's own w(p-1) read back from the window - and stores G(p-1).
// One barrier per step; one straight-line block per thread and step (the halo rows run the G part too and drop the result),
// so loads, shuffles, the MUFU and the arithmetic of the two parts interleave; the state carried across steps shrinks to
// the incoming z term (one value per voxel; the centred scheme two) plus the side elements of the x rows (the raw z
// difference is recomputed from the two x windows).  Price: a fourth plane slot in the x window ring (plane p-1 is read
// while plane p+2 lands) and the second w window - 222 KB for four coupled frames, which is why more than four coupled
// frames stay with the two-phase form (tv_path.cuh).
//
// Spec and edge-term algebra: tile_core.cuh (reference tv_GPU.py:84-126, :176-188, :239-251, :302-328).
#pragma once
#include "tile_core.cuh"

namespace pytvb {

// Per-thread state carried from one z step to the next.
template <typename T, int VEC, int R>
struct Tile2Thread {
    T f[R][VEC];     // one-sided / hybrid: srz * term(p-2 -> p-1), the z term coming into plane p-1;  centred: srz * Cz(p-2)
    T f1[R][VEC];    // centred only: srz * Cz(p-1)
    T xl[R], xr[R];  // one-sided / hybrid: the elements left / right of the quad in the rows of x(p-1) (exchanged in the step before)
    double tv;
};

PYTVB_HD int tile2_slot(int q) { return q & 3; }      // two's complement: -1 -> 3, -2 -> 2

// Plane that step q stages: clamped to the global volume when planes couple, to the slab otherwise (no halo buffers then).
template <typename T, bool Z_ON>
PYTVB_HD int tile2_plane(const Params<T>& P, int q) {
    return Z_ON ? tile_clamp_plane(P, q) : clampi(q, 0, P.Nz - 1);
}

// Geometry-dependent addresses of the windows of step p.
template <typename T>
struct Tile2Win {
    const T* Xm;     // x(p-1), own quad, work row 0
    const T* Xc;     // x(p)
    const T* Xn;     // x(p+1)
    const T* Wo;     // w(p-1)
    T* Wn;           // w(p)
};
// PH >= 0: p & 3 as a compile-time constant (four copies of the step, one per phase of the slot ring; every window is then a
// loop-invariant offset from two per-thread base addresses) - an experiment, see tile2_step.
template <typename T, int PH>
PYTVB_HD Tile2Win<T> tile2_windows(const TileCtx<T>& c, const TileGeom& g, const TilePos& tp, int p) {
    const int ph = PH >= 0 ? PH : p;        // PH < 0: the phase is taken from the plane index at run time
    Tile2Win<T> w;
    w.Xm = c.Xs + (long long)tile2_slot(ph - 1) * g.xslot + tp.xo;
    w.Xc = c.Xs + (long long)tile2_slot(ph) * g.xslot + tp.xo;
    w.Xn = c.Xs + (long long)tile2_slot(ph + 1) * g.xslot + tp.xo;
    w.Wo = c.Ws + (long long)((ph + 1) & 1) * g.wbuf + tp.wo;
    w.Wn = c.Ws + (long long)(ph & 1) * g.wbuf + tp.wo;
    return w;
}

// One step: norms of plane p, sub-gradient of plane p-1.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS, int PH>
PYTVB_HD void tile2_step_ph(Tile2Thread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, T* G, T* norms, int p,
                         const TilePos& tp) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    typedef VOp<T, VEC> V;
    constexpr bool FWD = C::NEED_FWD, BWD = C::NEED_BWD, CEN = SCHEME == CENTRAL;
    constexpr int PX = TileC<VEC>::PX, WJ = TileC<VEC>::WJ;
    const Tile2Win<T> win = tile2_windows<T, PH>(c, g, tp, p);
    const int t = tp.t, lane = tp.lane;
    const int ql = tile2_plane<T, Z_ON>(P, p);                                         // plane of the w part
    const int qg = clampi(p - 1, c.zc0, c.zc1 - 1);                                     // plane of the G part (clamped: steps without output)
    const long long zg = P.zg0 + p;
    const bool plane_out = p >= c.zc0 && p < c.zc1, prev_out = p - 1 >= c.zc0 && p - 1 < c.zc1;
    const T rz2 = P.srz * P.srz, rt2 = P.srt * P.srt;
    const T fz = CEN ? cen_exists<T>(zg, P.NzG) : T(1), ft = CEN ? cen_exists<T>(t, P.M) : T(1);
    T* Gq = G + (long long)(p - 1) * P.sZ + tp.goff;            // G(p-1) at the thread's quad, work row 0
    keep_in_registers(Gq);
    const T colf = tp.col_out ? T(1) : T(0);
    T tvq[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) tvq[e] = T(0);
    // rows carried down the strip: x(p) rows rr-1, rr; x(p-1) rows rr-1 (centred: rr-2 too), rr, rr+1; w(p-1) rows rr-1, rr
    T xu[VEC], xc[VEC], dip[VEC];
    ld_into<T, VEC>(xu, win.Xc - PX);
    ld_into<T, VEC>(xc, win.Xc);
    if (FWD && BWD) V::sub(dip, xc, xu);           // x(rr) - x(rr-1): the backward row difference of row rr
    T yc[VEC], yd[VEC], wc[VEC];
    T tup[VEC];            // one-sided / hybrid: row term (rr-1 -> rr)
    T yu[VEC], cu[VEC], cc[VEC];    // centred: x(p-1) row rr-1, C_i(rr-1), C_i(rr)
    ld_into<T, VEC>(yc, win.Xm);
    ld_into<T, VEC>(yd, win.Xm + PX);
    ld_into<T, VEC>(wc, win.Wo);
    {
        T wu[VEC], dx[VEC];
        ld_into<T, VEC>(yu, win.Xm - PX);
        ld_into<T, VEC>(wu, win.Wo - WJ);
        if (!CEN) {
            T sw[VEC];
            V::sub(dx, yc, yu);
            vpair_w<T, VEC, SCHEME>(sw, wu, wc);
            V::mul(tup, dx, sw);
        } else {
            // row rr0 - 2 of the window; the first strip (rr0 = -1) would leave the window into the slot that is being staged
            // (its C_i(-2) only feeds the dropped halo row): it reads row rr0 - 1 instead
            T yu2[VEC];
            ld_into<T, VEC>(yu2, win.Xm - (tp.rr0 >= 0 ? 2 : 1) * PX);
            const int gi0 = c.i0 + tp.rr0;
            V::sub(dx, yc, yu2);
            V::mul(cu, dx, wu);
            V::muls(cu, cu, cen_exists<T>(gi0 - 1, P.Ni));
            V::sub(dx, yd, yu);
            V::mul(cc, dx, wc);
            V::muls(cc, cc, cen_exists<T>(gi0, P.Ni));
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r, gi = c.i0 + rr;
        const bool row_out = rr >= 0 && rr < g.TI && gi < P.Ni;        // warp-uniform
        const T* yrow = win.Xm + r * PX;
        const T* wrow = win.Wo + r * WJ;
        const T* xrow = win.Xc + r * PX;
        // =========================== G part: in-plane and time edge terms of plane p-1 ===========================
        T gq[VEC], wd[VEC], y2[VEC];       // y2: x(p-1) row rr+2 (the centred row term needs it now, the other schemes in the next row)
        ld_into<T, VEC>(wd, wrow + WJ);
        if (CEN || r + 1 < R) ld_into<T, VEC>(y2, yrow + 2 * PX);
        if (!CEN) {
            T dx[VEC], sw[VEC], tdn[VEC];
            V::sub(dx, yd, yc);
            vpair_w<T, VEC, SCHEME>(sw, wc, wd);
            V::mul(tdn, dx, sw);
            V::sub(gq, tup, tdn);
#pragma unroll
            for (int e = 0; e < VEC; ++e) tup[e] = tdn[e];
            // columns (element-shifted within the quad: scalar code); the side elements of the x row were exchanged when this
            // plane went through the w part (previous step), those of w come from the neighbouring lanes now
            T wlv, wrv;
            quad_sides<T, VEC, 1>(&wlv, &wrv, wc, wrow, lane, false);
            T tj = (yc[0] - st.xl[r]) * pair_w<T, SCHEME>(wlv, wc[0]);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T xe = e + 1 < VEC ? yc[e + 1 < VEC ? e + 1 : e] : st.xr[r], we = e + 1 < VEC ? wc[e + 1 < VEC ? e + 1 : e] : wrv;
                const T tn = (xe - yc[e]) * pair_w<T, SCHEME>(wc[e], we);
                gq[e] += tj - tn;
                tj = tn;
            }
        } else {
            // C_i(m) = exists(m) * (x(m+1) - x(m-1)) * w(m);  G_i(rr) = C_i(rr-1) - C_i(rr+1)
            T cd[VEC], dx[VEC];
            V::sub(dx, y2, yc);
            V::mul(cd, dx, wd);
            V::muls(cd, cd, cen_exists<T>(gi + 1, P.Ni));
            V::sub(gq, cu, cd);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { cu[e] = cc[e]; cc[e] = cd[e]; }
            // x with 2, w with 1 element on each side; for the halo lanes the outer elements are clamped (results unused)
            T xw[VEC + 4], ww[VEC + 2];
#pragma unroll
            for (int e = 0; e < VEC; ++e) { xw[e + 2] = yc[e]; ww[e + 1] = wc[e]; }
            T xs_l[2], xs_r[2];
            quad_sides<T, VEC, 2>(xs_l, xs_r, yc, yrow, lane, VEC == 1);
            xw[0] = xs_l[1]; xw[1] = xs_l[0]; xw[VEC + 2] = xs_r[0]; xw[VEC + 3] = xs_r[1];
            quad_sides<T, VEC, 1>(&ww[0], &ww[VEC + 1], wc, wrow, lane, false);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int gj = c.j0 + tp.cj + e;
                const T cm = cen_exists<T>(gj - 1, P.Nj) * ((xw[e + 2] - xw[e]) * ww[e]);
                const T cp = cen_exists<T>(gj + 1, P.Nj) * ((xw[e + 4] - xw[e + 2]) * ww[e + 2]);
                gq[e] += cm - cp;
            }
        }
        if (T_ON) {
            const int dm = tp.fl > 0 ? -1 : 0, dp = tp.fl < g.FC - 1 ? 1 : 0;          // neighbouring frames, clamped
            T ym[VEC], yp[VEC], wm[VEC], wp[VEC], wq[VEC], v[VEC], v2[VEC], dx[VEC], sw[VEC];
            ld_into<T, VEC>(wm, wrow + tp.dwm);
            ld_into<T, VEC>(wp, wrow + tp.dwp);
#pragma unroll
            for (int e = 0; e < VEC; ++e) wq[e] = wc[e];
            if constexpr (TSMODE == 2) {   // along t the inverse norms travel with their voxel's scale (strip_quad_G_impl)
                T f[VEC];
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, qg, t + dm);
                V::mul(wm, wm, f);
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, qg, t);
                V::mul(wq, wq, f);
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, qg, t + dp);
                V::mul(wp, wp, f);
            }
            if (!CEN) {
                ld_into<T, VEC>(ym, yrow + tp.dxm);
                ld_into<T, VEC>(yp, yrow + tp.dxp);
                V::sub(dx, yc, ym);
                vpair_w<T, VEC, SCHEME>(sw, wm, wq);
                V::mul(v, dx, sw);
                V::sub(dx, yp, yc);
                vpair_w<T, VEC, SCHEME>(sw, wq, wp);
                V::mul(v2, dx, sw);
                V::sub(v, v, v2);
                V::muls(v, v, P.srt);
            } else {
                // C_t(t-1) = exists(t-1) (x(t) - x(t-2)) w(t-1);  C_t(t+1) = exists(t+1) (x(t+2) - x(t)) w(t+1)
                const int dm2 = tp.fl > 1 ? -2 : -tp.fl, dp2 = tp.fl < g.FC - 2 ? 2 : g.FC - 1 - tp.fl;
                ld_into<T, VEC>(ym, yrow + dm2 * g.slotX);
                ld_into<T, VEC>(yp, yrow + dp2 * g.slotX);
                const T am = P.srt * cen_exists<T>(t - 1, P.M), ap = P.srt * cen_exists<T>(t + 1, P.M);
                V::sub(dx, yc, ym);
                V::mul(v, dx, wm);
                V::muls(v, v, am);
                V::sub(dx, yp, yc);
                V::mul(v2, dx, wp);
                V::muls(v2, v2, ap);
                V::sub(v, v, v2);
            }
            if constexpr (TSMODE >= 1) {
                if (c.Ms) {
                    T f[VEC];
                    ld_into<T, VEC>(f, c.Ms + (rr + 1) * WJ + tp.cj + VEC);
                    V::mul(v, v, f);
                }
            }
            V::add(gq, gq, v);
        }
        // =========================== w part: norms of plane p ===========================
        T xd[VEC], s[VEC];
        ld_into<T, VEC>(xd, xrow + PX);
        T cl, cr;
        quad_sides<T, VEC, 1>(&cl, &cr, xc, xrow, lane, VEC == 1);
        if (!CEN) {
            st.xl[r] = cl;
            st.xr[r] = cr;
            // column differences: VEC + 1 of them serve the forward and the backward component of the quad
            T djf[VEC], djb[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) djf[e] = (e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - xc[e];
#pragma unroll
            for (int e = 0; e < VEC; ++e) djb[e] = e > 0 ? (FWD ? djf[e > 0 ? e - 1 : 0] : xc[e] - xc[e > 0 ? e - 1 : 0]) : xc[0] - cl;
            T di[VEC];
            if (FWD) {
                V::sub(di, xd, xc);
                V::mul(s, di, di);
                V::fma(s, djf, djf, s);
            }
            if (BWD) {
                if (FWD) {
                    V::fma(s, dip, dip, s);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dip[e] = di[e];
                } else {
                    V::sub(di, xc, xu);
                    V::mul(s, di, di);
                }
                V::fma(s, djb, djb, s);
            }
        } else {
            const T fi = cen_exists<T>(gi, P.Ni);
            T di[VEC], dj[VEC];
            V::sub(di, xd, xu);
            V::muls(di, di, fi);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T fj = cen_exists<T>(c.j0 + tp.cj + e, P.Nj);
                dj[e] = fj * ((e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - (e > 0 ? xc[e > 0 ? e - 1 : 0] : cl));
            }
            V::mul(s, di, di);
            V::fma(s, dj, dj, s);
        }
        T az[VEC];       // one-sided / hybrid: the raw z difference x(p) - x(p-1);  centred: fz * (x(p+1) - x(p-1))
        if (Z_ON) {
            T xn[VEC], q[VEC];
            ld_into<T, VEC>(xn, win.Xn + r * PX);
            if (!CEN) {
                T d[VEC];
                V::sub(az, xc, yc);
                V::sub(d, xn, xc);
                if (FWD) V::mul(q, d, d);
                if (BWD) { if (FWD) V::fma(q, az, az, q); else V::mul(q, az, az); }
            } else {
                V::sub(az, xn, yc);
                V::muls(az, az, fz);
                V::mul(q, az, az);
            }
            V::fmas(s, q, rz2, s);
        }
        if (T_ON) {
            T xm[VEC], xp[VEC], q[VEC], d[VEC];
            ld_into<T, VEC>(xm, xrow + tp.dxm);
            ld_into<T, VEC>(xp, xrow + tp.dxp);
            if (!CEN) {
                if (FWD) { V::sub(d, xp, xc); V::mul(q, d, d); }
                if (BWD) { V::sub(d, xc, xm); if (FWD) V::fma(q, d, d, q); else V::mul(q, d, d); }
            } else {
                V::sub(d, xp, xm);
                V::muls(d, d, ft);
                V::mul(q, d, d);
            }
            if constexpr (TSMODE == 0) {
                V::fmas(s, q, rt2, s);
            } else {
                T fac[VEC];
                tile_time_factor<T, VEC, TSMODE>(fac, c, g, P, TS, rr + 1, tp.cj, ql, t);
                V::muls(fac, fac, P.srt);
                V::mul(fac, fac, fac);
                V::fma(s, fac, q, s);
            }
        }
        Pack<T, VEC> wq;
        T nrv[VEC];
        bool posv[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) tile_norm<T, NORMS>(s[e], P, wq.v[e], posv[e]);
        V::mul(nrv, s, wq.v);             // |D x| = sqrt(s) / div = s * w (0 by itself where s = 0)
        V::fmas(tvq, nrv, row_out ? colf : T(0), tvq);
        st_pack<T, VEC>(win.Wn + r * WJ, wq);
        if constexpr (NORMS) {
            if (plane_out && row_out && tp.col_out) {
                Pack<T, VEC> nq;
#pragma unroll
                for (int e = 0; e < VEC; ++e) nq.v[e] = posv[e] ? nrv[e] : T(INFINITY);
                st_pack<T, VEC>(norms + (long long)p * P.sZ + tp.goff + (long long)r * P.Nj, nq);
            }
        }
        // =========================== the z terms complete G(p-1) ===========================
        Pack<T, VEC> go;
        if (Z_ON) {
            if (!CEN) {
                T sw[VEC];
                vpair_w<T, VEC, SCHEME>(sw, wc, wq.v);
                V::add(gq, gq, st.f[r]);
                V::mul(sw, az, sw);
                V::muls(st.f[r], sw, P.srz);          // srz * term(p-1 -> p): leaves plane p-1, enters plane p in the next step
                V::sub(gq, gq, st.f[r]);
            } else {
                T en[VEC];
                V::mul(en, az, wq.v);
                V::muls(en, en, P.srz);               // srz * Cz(p)
                V::add(gq, gq, st.f[r]);
                V::sub(gq, gq, en);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { st.f[r][e] = st.f1[r][e]; st.f1[r][e] = en[e]; }
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) go.v[e] = gq[e];
        if (prev_out && row_out && tp.col_out) st_pack<T, VEC>(Gq + (long long)r * P.Nj, go);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { xu[e] = xc[e]; xc[e] = xd[e]; wc[e] = wd[e]; yc[e] = yd[e]; }
        if (CEN || r + 1 < R) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) yd[e] = y2[e];
        }
    }
    if (plane_out) {
        T sum = tvq[0];
#pragma unroll
        for (int e = 1; e < VEC; ++e) sum += tvq[e];
        st.tv += (double)sum;
    }
}

// The step of plane p.  PYTVB_T2_PHASES=1 dispatches on the phase of the slot ring (four copies of the step whose window
// offsets are loop-invariant: 10 % fewer instructions, but measured SLOWER - 2.38 vs 2.05 ms on the C4 slab,
// profiles/r02r_tv_times.txt: four times the code and 24 bytes of spills); the default computes the window addresses per step.
#ifndef PYTVB_T2_PHASES
#define PYTVB_T2_PHASES 0
#endif
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS>
PYTVB_HD void tile2_step(Tile2Thread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, T* G, T* norms, int p,
                         const TilePos& tp) {
#if PYTVB_T2_PHASES
    switch (p & 3) {
        case 0: tile2_step_ph<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS, 0>(st, c, g, P, TS, G, norms, p, tp); break;
        case 1: tile2_step_ph<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS, 1>(st, c, g, P, TS, G, norms, p, tp); break;
        case 2: tile2_step_ph<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS, 2>(st, c, g, P, TS, G, norms, p, tp); break;
        default: tile2_step_ph<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS, 3>(st, c, g, P, TS, G, norms, p, tp); break;
    }
#else
    tile2_step_ph<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS, -1>(st, c, g, P, TS, G, norms, p, tp);
#endif
}

}  // namespace pytvb
