// Single-sweep TV value + sub-gradient (tv_<scheme>, reference pytv/tv_GPU.py:47,142,217,290): per-thread code of the
// z-marching tile kernel (kernels_tile.cuh).
//
// Spec (tv_GPU.py:84-126 hybrid, :176-188, :239-251, :302-328): D = D_s(x); n = |D|_2 per voxel; tv = sum n;
// G = D_T_s^unit(D / n) with 0/0 := 0.  With w = 1/n (0 where n = 0) the sub-gradient is a sum of EDGE
// TERMS: along one axis, term(a -> b) = (x_b - x_a) * S(w_a, w_b) for neighbouring voxels a, b (S = w_a upwind, w_b
// downwind, w_a + w_b hybrid) and G(v) = sum_axes [term(v-e -> v) - term(v -> v+e)] * weight; the centred scheme has
// C(m) = (x_{m+1} - x_{m-1}) * w_m and G(v) = C(v-e) - C(v+e).  So G(v) needs w at distance 1 and x at distance 2.
//
// One launch, x read once, G written once (8 B/voxel): a CTA owns an in-plane tile (TI x TJ output voxels, all M time
// frames when the time axis is on) and marches along z.  Per z-plane p ("step"):
//   * x(p+2) arrives in shared memory (cp.async issued one step ahead; the window holds the tile plus 2 halo rows / columns
//     with indices CLAMPED at the volume boundary, so every out-of-range difference is exactly 0 - the reference's rule,
//     tv_operators_CPU.py:118 - without predicates; zero fill would be wrong);
//   * w-phase: every thread computes n(p), w(p) for its R rows x one quad column (tile + 1 halo ring), publishes w(p) to
//     shared memory, forms the z edge term between planes p-1 and p from registers and with it completes and stores G(p-1);
//   * barrier; G-phase: the in-plane and time edge terms of plane p from the x / w tiles -> kept in registers until the
//     next step supplies the z term.
// z neighbours live in registers (each thread keeps, per owned voxel: the raw z difference, w and the running G: 3 values),
// in-plane and time neighbours in shared memory (3 x-planes x M frames, 1 w-plane x M frames).
// Everything here is __host__ __device__ and index-pure: tests/emul runs the same code thread by thread on the CPU.
#pragma once
#include "strip_core.cuh"

namespace pytvb {

// Geometry of one launch (host-computed, warp-uniform).
struct TileGeom {
    int strips;        // R-row strips per frame handled by one CTA
    int FC;            // frames per CTA: M when the time axis is on, else 1
    int RPF;           // work rows per frame = strips * R  (tile rows -1 .. TI)
    int TI, TJ;        // output tile
    int WJ;            // work columns = 32 * VEC (tile columns -VEC .. TJ+VEC-1)
    int pitchX, rowsX; // x window: rows -2 .. TI+1, columns -VEC-1 .. TJ+VEC  (pitch = WJ + 2 VEC)
    int slotX, slotW;  // elements of one (plane, frame) x window / w window
    int nti, ntj, nfg, nzc, Lz;   // tiles along i, j; frame groups; z chunks and their length
    int nthreads;
    long long nblocks;
};

// Compile-time part of the geometry (a function of the vector width only): with these as constants the window rows of a
// thread are immediate offsets from one base register instead of a multiply-add per access.
template <int VEC>
struct TileC {
    static constexpr int WJ = 32 * VEC;        // work columns
    static constexpr int TJ = 30 * VEC;        // output columns
    static constexpr int PX = 34 * VEC;        // pitch of the x window
};

template <typename T>
struct TileCtx {
    T* Xs;             // [3][FC][slotX]
    T* Ws;             // [FC][slotW]
    uint8_t* Ms;       // [RPF][WJ] static-mask bytes of the work region, or null
    long long* rowg;   // [FC * rowsX] staging table: element offset of window row r inside a z-plane group (frame and clamped row)
    int* rowd;         // [FC * rowsX] staging table: element offset of window row r inside a slot group (frame, row)
    int i0, j0, t0;    // global row / column of tile (0, 0); first frame
    int zc0, zc1;      // output planes [zc0, zc1), slab-local
};

// Per-thread state carried from one z step to the next.
template <typename T, int VEC, int R>
struct TileThread {
    T a[R][VEC];    // one-sided / hybrid: raw z difference x(p) - x(p-1);  centred: x(p-1)
    T w[R][VEC];    // w of the plane of the last w-phase
    T e[R][VEC];    // centred only: srz * Cz(p-1)
    T g[R][VEC];    // running sub-gradient: in-plane + time part of the last G-phase plus the incoming z term
    double tv;      // running sum of the norms of this thread's output voxels (one double add per step: the float part spans R rows only,
                    // so the value does not depend on how the volume is cut into z chunks or slabs beyond 1e-16)
};

PYTVB_HD int tile_slot(int q) { return (q + 3) % 3; }   // q >= -3

// ---- staging: global -> shared, indices clamped to the volume.
template <int BYTES>
PYTVB_HD void stage_copy(void* dst, const void* src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
#else
    memcpy(dst, src, BYTES);
#endif
}
PYTVB_HD void stage_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

PYTVB_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Slab-local index of plane q clamped to the GLOBAL volume (planes outside the slab but inside the volume come from the
// halo buffers of the view).
template <typename T>
PYTVB_HD int tile_clamp_plane(const Params<T>& P, int q) {
    long long zg = P.zg0 + q;
    if (zg < 0) zg = 0;
    if (zg > P.NzG - 1) zg = P.NzG - 1;
    return (int)(zg - P.zg0);
}

// Staging tables, once per CTA: everything about a window row that does not depend on the plane.
template <typename T>
PYTVB_HD void tile_stage_tables(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    for (int r = tid; r < g.FC * g.rowsX; r += g.nthreads) {
        const int fl = r / g.rowsX, xr = r - fl * g.rowsX;
        c.rowg[r] = (long long)(c.t0 + fl) * P.sT + (long long)clampi(c.i0 + xr - 2, 0, P.Ni - 1) * P.Nj;
        c.rowd[r] = fl * g.slotX + xr * g.pitchX;
    }
}

// Stage plane q of every frame of the CTA into its slot.  Work is split by warps over (frame, window row); a lane copies
// one quad of the row, lanes 0 / 1 also the scalar column left / right of the work columns.  Per row: one table look-up,
// one add and the copy (the first version recomputed frame / row / clamps per copy and spent 40 % of the kernel's
// instructions here, profiles/r02b_tv_tile_first_ncu_full.txt).
template <typename T, int VEC>
PYTVB_HD void tile_stage_plane(const TileCtx<T>& c, const TileGeom& g, const ImgView<T>& X, const Params<T>& P, int q, int tid) {
    const int lane = tid & 31, wid = tid >> 5, nwarps = g.nthreads >> 5;
    const T* plane = X.row(P, tile_clamp_plane(P, q), 0, 0);
    T* slot = c.Xs + (long long)tile_slot(q) * g.FC * g.slotX;
    const int cj = -VEC + lane * VEC;                 // tile column of this lane's quad
    const int gj0 = c.j0 + cj;
    const bool vec_ok = VEC > 1 && gj0 >= 0 && gj0 + VEC <= P.Nj;
    const int nrows = g.FC * g.rowsX;
    if (vec_ok) {
        for (int rq = wid; rq < nrows; rq += nwarps) stage_copy<sizeof(T) * VEC>(slot + c.rowd[rq] + cj + 2 * VEC, plane + c.rowg[rq] + gj0);
    } else {
        for (int rq = wid; rq < nrows; rq += nwarps) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) stage_copy<sizeof(T)>(slot + c.rowd[rq] + cj + 2 * VEC + e, plane + c.rowg[rq] + clampi(gj0 + e, 0, P.Nj - 1));
        }
    }
    if (lane < 2) {
        const int gcol = lane == 0 ? clampi(c.j0 - VEC - 1, 0, P.Nj - 1) : clampi(c.j0 + g.TJ + VEC, 0, P.Nj - 1);
        const int dcol = lane == 0 ? VEC - 1 : g.TJ + 3 * VEC;
        for (int rq = wid; rq < nrows; rq += nwarps) stage_copy<sizeof(T)>(slot + c.rowd[rq] + dcol, plane + c.rowg[rq] + gcol);
    }
}

// Static-mask bytes of the work region (clamped), once per CTA.
template <typename T, int VEC>
PYTVB_HD void tile_stage_mask(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    for (int k = tid; k < g.RPF * g.WJ; k += g.nthreads) {
        const int r = k / g.WJ, cc = k - r * g.WJ;
        const int gi = clampi(c.i0 + r - 1, 0, P.Ni - 1), gj = clampi(c.j0 + cc - VEC, 0, P.Nj - 1);
        c.Ms[k] = P.mask_static[(long long)gi * P.Nj + gj];
    }
}

// ---- thread coordinates
struct TilePos {
    int fl, rr0, cj;     // frame within the CTA, first work row (tile coordinates, >= -1), tile column of the quad
    int t;               // global frame
    bool col_out;        // the quad lies in the output tile and inside the image
    long long goff;      // element offset of (frame t, row i0 + rr0, column j0 + cj) inside a z-plane group (outputs)
};
// Computed once per thread (the division by the strip count and the 64-bit products stay out of the z loop).
template <typename T, int VEC, int R>
PYTVB_HD TilePos tile_pos(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    TilePos p;
    const int lane = tid & 31, wid = tid >> 5;
    p.fl = wid / g.strips;
    p.rr0 = -1 + (wid - p.fl * g.strips) * R;
    p.cj = -VEC + lane * VEC;
    p.t = c.t0 + p.fl;
    p.col_out = p.cj >= 0 && p.cj < g.TJ && c.j0 + p.cj < P.Nj;     // vector path: Nj % VEC == 0, so a quad is in or out as a whole
    p.goff = (long long)p.t * P.sT + (long long)(c.i0 + p.rr0) * P.Nj + c.j0 + p.cj;
    return p;
}

// Factor of the time component(s) at the thread's quad in work row wr (0-based work row): sqrt(factor_reg_static) on
// static pixels (tv_operators_CPU.py:148-150) times the per-voxel time scale (extension, README.md:258) at plane ql / frame t.
// TSMODE 0: none (uniform weight), 1: mask_static, 2: time-scale map (and mask_static when present).
template <typename T, int VEC, int TSMODE>
PYTVB_HD void tile_time_factor(T* f, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, int wr, int cj, int ql, int t) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = T(1);
    if constexpr (TSMODE >= 1) {
        if (c.Ms) {
            const uint8_t* m = c.Ms + wr * TileC<VEC>::WJ + cj + VEC;
#pragma unroll
            for (int e = 0; e < VEC; ++e) f[e] = m[e] ? P.sfac : T(1);
        }
    }
    if constexpr (TSMODE == 2) {
        const int gi = clampi(c.i0 + wr - 1, 0, P.Ni - 1);
        const T* row = TS.row(P, ql, t, gi);
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] *= row[clampi(c.j0 + cj + e, 0, P.Nj - 1)];
    }
}
// The time scale alone at (plane ql, frame t) (the inverse norms travel with their voxel's scale along t, strip_quad_G_impl).
template <typename T, int VEC>
PYTVB_HD void tile_time_scale(T* f, const TileCtx<T>& c, const Params<T>& P, const ImgView<T>& TS, int wr, int cj, int ql, int t) {
    const int gi = clampi(c.i0 + wr - 1, 0, P.Ni - 1);
    const T* row = TS.row(P, ql, t, gi);
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = row[clampi(c.j0 + cj + e, 0, P.Nj - 1)];
}

// Existence factor of a centred difference at index k of an axis of length L (tv_operators_CPU.py:331-358).
template <typename T>
PYTVB_HD T cen_exists(long long k, long long L) { return (k >= 1 && k <= L - 2) ? T(1) : T(0); }

// w = 1/|D x| (0 where the norm is 0) and the norm itself from the sum of squares: one MUFU, one select.
// nr = sqrt(s) / div = s * w * k2 with k2 = 1/div^2 (w = div / sqrt(s)), which is 0 by itself where w was forced to 0.
template <typename T>
PYTVB_HD void tile_norm(T s, const Params<T>& P, T k2, T& nr, T& w, bool& pos) {
    const T rs = fast_rsqrt(s);
    pos = s > T(0);
    w = pos ? rs * P.div : T(0);
    nr = s * w * k2;
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ void tile_norm<float>(float s, const Params<float>& P, float k2, float& nr, float& w, bool& pos) {
    float rs;     // a sum of squares below FLT_MIN (all differences < 1.1e-19) counts as zero, as in norm_finish
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(s, 1.17549435e-38f)));
    pos = s >= 1.17549435e-38f;
    w = pos ? rs * P.div : 0.0f;
    nr = s * w * k2;
}
#endif

// Initial z state of a chunk whose first step is plane p: needs the thread's own x(p-1) (straight from global memory, once).
template <typename T, int VEC, int SCHEME, int R>
PYTVB_HD void tile_init_z(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const ImgView<T>& X, const Params<T>& P, int p,
                          const TilePos& tp) {
    constexpr int PX = TileC<VEC>::PX;
    const int qm = tile_clamp_plane(P, p - 1);
    const T* Xp = c.Xs + ((long long)tile_slot(p) * g.FC + tp.fl) * g.slotX;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r;
        const int gi = clampi(c.i0 + rr, 0, P.Ni - 1);
        const T* src = X.row(P, qm, c.t0 + tp.fl, gi);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T xm = src[clampi(c.j0 + tp.cj + e, 0, P.Nj - 1)];
            st.a[r][e] = SCHEME == CENTRAL ? xm : Xp[(rr + 2) * PX + tp.cj + 2 * VEC + e] - xm;
            st.w[r][e] = st.e[r][e] = st.g[r][e] = T(0);
        }
    }
}

// ---- w-phase of plane p.  Outputs: w(p) to shared memory, norms (optional) and the TV partial sum for output voxels, and -
// with the z axis on - the finished G(p-1).
// Register diet (the kernel is instruction-issue bound, 128 registers per thread at 16 warps per SM): the z edge term is
// folded into the running sub-gradient - st.g holds  Gp(p-1) + srz * term(p-2 -> p-1)  on entry, G(p-1) = st.g - srz * term(p-1 -> p)
// is stored, and st.g restarts as  srz * term(p-1 -> p)  for the G-phase to add Gp(p) to - so a thread carries three values
// per voxel (raw z difference, w, running G), the centred scheme four (x(p-1), Cz(p-1), running G; w within the step).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE>
PYTVB_HD void tile_phase_w(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS,
                           T* G, T* norms, int p, const TilePos& tp) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    constexpr bool FWD = C::NEED_FWD, BWD = C::NEED_BWD, CEN = SCHEME == CENTRAL;
    constexpr int PX = TileC<VEC>::PX, WJ = TileC<VEC>::WJ;
    const int sp = tile_slot(p);
    const int xo = (tp.rr0 + 2) * PX + tp.cj + 2 * VEC;        // own quad in an x window, work row 0
    const T* Xp = c.Xs + ((long long)sp * g.FC + tp.fl) * g.slotX + xo;
    const T* Xn = c.Xs + ((long long)tile_slot(p + 1) * g.FC + tp.fl) * g.slotX + xo;
    const int flm = tp.fl > 0 ? tp.fl - 1 : tp.fl, flp = tp.fl < g.FC - 1 ? tp.fl + 1 : tp.fl;
    const T* Xtm = c.Xs + ((long long)sp * g.FC + flm) * g.slotX + xo;
    const T* Xtp = c.Xs + ((long long)sp * g.FC + flp) * g.slotX + xo;
    T* Wp = c.Ws + (long long)tp.fl * g.slotW + (tp.rr0 + 1) * WJ + tp.cj + VEC;
    const int t = tp.t;
    const int ql = tile_clamp_plane(P, p);
    const long long zg = P.zg0 + p;
    const bool plane_out = p >= c.zc0 && p < c.zc1, prev_out = Z_ON && p - 1 >= c.zc0 && p - 1 < c.zc1;
    const T rz2 = P.srz * P.srz, rt2 = P.srt * P.srt, k2 = P.inv_div * P.inv_div;
    const T fz = CEN ? cen_exists<T>(zg, P.NzG) : T(1), ft = CEN ? cen_exists<T>(t, P.M) : T(1);
    T* Gq = G + (long long)(p - 1) * P.sZ + tp.goff;            // G(p-1) at the thread's quad, work row 0
    T tvstep = T(0);
    T xu[VEC], xc[VEC];
    ld_into<T, VEC>(xu, Xp - PX);
    ld_into<T, VEC>(xc, Xp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r, gi = c.i0 + rr;
        T xd[VEC], s[VEC];
        ld_into<T, VEC>(xd, Xp + (r + 1) * PX);
        const T cl = Xp[r * PX - 1], cr = Xp[r * PX + VEC];
        if (!CEN) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                s[e] = T(0);
                if (FWD) {
                    const T di = xd[e] - xc[e], dj = (e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - xc[e];
                    s[e] = di * di;
                    s[e] += dj * dj;
                }
                if (BWD) {
                    const T di = xc[e] - xu[e], dj = xc[e] - (e > 0 ? xc[e > 0 ? e - 1 : 0] : cl);
                    s[e] += di * di;
                    s[e] += dj * dj;
                }
            }
        } else {
            const T fi = cen_exists<T>(gi, P.Ni);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T fj = cen_exists<T>(c.j0 + tp.cj + e, P.Nj);
                const T di = fi * (xd[e] - xu[e]);
                const T dj = fj * ((e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - (e > 0 ? xc[e > 0 ? e - 1 : 0] : cl));
                s[e] = di * di;
                s[e] += dj * dj;
            }
        }
        T dzf[VEC];       // one-sided / hybrid: x(p+1) - x(p);  centred: x(p+1) - x(p-1), 0 where it does not exist
        if (Z_ON) {
            T xn[VEC];
            ld_into<T, VEC>(xn, Xn + r * PX);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (!CEN) {
                    dzf[e] = xn[e] - xc[e];
                    T q = T(0);
                    if (FWD) q = dzf[e] * dzf[e];
                    if (BWD) q += st.a[r][e] * st.a[r][e];
                    s[e] += rz2 * q;
                } else {
                    dzf[e] = fz * (xn[e] - st.a[r][e]);
                    s[e] += rz2 * (dzf[e] * dzf[e]);
                }
            }
        }
        if (T_ON) {
            T xm[VEC], xp[VEC], q[VEC];
            ld_into<T, VEC>(xm, Xtm + r * PX);
            ld_into<T, VEC>(xp, Xtp + r * PX);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (!CEN) {
                    q[e] = T(0);
                    if (FWD) { const T d = xp[e] - xc[e]; q[e] = d * d; }
                    if (BWD) { const T d = xc[e] - xm[e]; q[e] += d * d; }
                } else {
                    const T d = ft * (xp[e] - xm[e]);
                    q[e] = d * d;
                }
            }
            if constexpr (TSMODE == 0) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) s[e] += rt2 * q[e];
            } else {
                T fac[VEC];
                tile_time_factor<T, VEC, TSMODE>(fac, c, g, P, TS, rr + 1, tp.cj, ql, t);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { const T wt = P.srt * fac[e]; s[e] += (wt * wt) * q[e]; }
            }
        }
        Pack<T, VEC> wq;
        T nrv[VEC];
        bool posv[VEC];
        T rowsum = T(0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            tile_norm<T>(s[e], P, k2, nrv[e], wq.v[e], posv[e]);
            rowsum += nrv[e];
        }
        st_pack<T, VEC>(Wp + r * WJ, wq);
        const bool row_out = rr >= 0 && rr < g.TI && gi < P.Ni;        // warp-uniform
        if (plane_out && row_out && tp.col_out) {
            tvstep += rowsum;
            if (norms) {
                Pack<T, VEC> nq;
#pragma unroll
                for (int e = 0; e < VEC; ++e) nq.v[e] = posv[e] ? nrv[e] : T(INFINITY);
                st_pack<T, VEC>(norms + (long long)p * P.sZ + tp.goff + (long long)r * P.Nj, nq);
            }
        }
        if (Z_ON) {
            // z edge term between planes p-1 and p (centred: Cz(p)); it completes G(p-1)
            Pack<T, VEC> gq;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (!CEN) {
                    const T en = P.srz * (st.a[r][e] * pair_w<T, SCHEME>(st.w[r][e], wq.v[e]));
                    gq.v[e] = (st.g[r][e] - en) * k2;
                    st.g[r][e] = en;
                    st.a[r][e] = dzf[e];
                } else {
                    const T cz = P.srz * (dzf[e] * wq.v[e]);
                    gq.v[e] = (st.g[r][e] - cz) * k2;
                    st.g[r][e] = st.e[r][e];        // srz * Cz(p-1): the incoming z term of G(p)
                    st.e[r][e] = cz;
                    st.a[r][e] = xc[e];
                }
            }
            if (prev_out && row_out && tp.col_out) st_pack<T, VEC>(Gq + (long long)r * P.Nj, gq);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) { st.w[r][e] = wq.v[e]; xu[e] = xc[e]; xc[e] = xd[e]; }
    }
    if (plane_out) st.tv += (double)tvstep;
}

// ---- G-phase of plane p: in-plane and time edge terms, added to st.g (z axis on) or stored as the finished G(p) (z axis off).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE>
PYTVB_HD void tile_phase_g(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, T* G, int p,
                           const TilePos& tp) {
    constexpr bool CEN = SCHEME == CENTRAL;
    constexpr int PX = TileC<VEC>::PX, WJ = TileC<VEC>::WJ;
    const int sp = tile_slot(p);
    const int xcol = tp.cj + 2 * VEC, wcol = tp.cj + VEC;
    const int xo0 = (tp.rr0 + 2) * PX + xcol, wo0 = (tp.rr0 + 1) * WJ + wcol;       // own quad, work row 0
    const T* Xp = c.Xs + ((long long)sp * g.FC + tp.fl) * g.slotX + xo0;
    const T* Wp = c.Ws + (long long)tp.fl * g.slotW + wo0;
    const int t = tp.t;
    const T k2 = P.inv_div * P.inv_div;
    // halo lanes (first / last quad of the work columns) have no left / right neighbour in the w window: clamp (results unused)
    const int wl = wcol > 0 ? -1 : 0, wr = wcol + VEC < WJ ? VEC : VEC - 1;
    bool have = false;
    T tdn[VEC];          // one-sided / hybrid: row term (rr -> rr+1) of the previous row
    T cu[VEC], cc[VEC];  // centred: C_i(rr-1), C_i(rr)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r, gi = c.i0 + rr;
        const bool row_out = rr >= 0 && rr < g.TI;          // warp-uniform: halo rows carry no output
        if (!row_out) {
            have = false;
            continue;
        }
        const T* xrow = Xp + r * PX;
        const T* wrow = Wp + r * WJ;
        T xc[VEC], wc[VEC], gq[VEC];
        ld_into<T, VEC>(xc, xrow);
#pragma unroll
        for (int e = 0; e < VEC; ++e) wc[e] = st.w[r][e];
        // ---- rows
        if (!CEN) {
            T xd[VEC], wd[VEC], tup[VEC];
            if (!have) {
                T xu[VEC], wu[VEC];
                ld_into<T, VEC>(xu, xrow - PX);
                ld_into<T, VEC>(wu, wrow - WJ);
#pragma unroll
                for (int e = 0; e < VEC; ++e) tup[e] = (xc[e] - xu[e]) * pair_w<T, SCHEME>(wu[e], wc[e]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) tup[e] = tdn[e];
            }
            ld_into<T, VEC>(xd, xrow + PX);
            if (r + 1 < R) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wd[e] = st.w[r + 1 < R ? r + 1 : r][e];
            } else ld_into<T, VEC>(wd, wrow + WJ);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                tdn[e] = (xd[e] - xc[e]) * pair_w<T, SCHEME>(wc[e], wd[e]);
                gq[e] = tup[e] - tdn[e];
            }
        } else {
            // C_i(m) = exists(m) * (x(m+1) - x(m-1)) * w(m);  G_i(rr) = C_i(rr-1) - C_i(rr+1)
            T xu[VEC], xd[VEC], xd2[VEC], wd[VEC], cd[VEC];
            ld_into<T, VEC>(xd, xrow + PX);
            ld_into<T, VEC>(xd2, xrow + 2 * PX);
            if (!have) {
                T xu2[VEC], wu[VEC];
                ld_into<T, VEC>(xu, xrow - PX);
                ld_into<T, VEC>(xu2, xrow - 2 * PX);
                ld_into<T, VEC>(wu, wrow - WJ);
                const T fu = cen_exists<T>(gi - 1, P.Ni), fc = cen_exists<T>(gi, P.Ni);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    cu[e] = fu * ((xc[e] - xu2[e]) * wu[e]);
                    cc[e] = fc * ((xd[e] - xu[e]) * wc[e]);
                }
            }
            if (r + 1 < R) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wd[e] = st.w[r + 1 < R ? r + 1 : r][e];
            } else ld_into<T, VEC>(wd, wrow + WJ);
            const T fd = cen_exists<T>(gi + 1, P.Ni);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                cd[e] = fd * ((xd2[e] - xc[e]) * wd[e]);
                gq[e] = cu[e] - cd[e];
                cu[e] = cc[e];
                cc[e] = cd[e];
            }
        }
        have = true;
        // ---- columns
        if (!CEN) {
            const T xl = xrow[-1], xr = xrow[VEC], wlv = wrow[wl], wrv = wrow[wr];
            T tj = (xc[0] - xl) * pair_w<T, SCHEME>(wlv, wc[0]);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T xe = e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : xr, we = e + 1 < VEC ? wc[e + 1 < VEC ? e + 1 : e] : wrv;
                const T tn = (xe - xc[e]) * pair_w<T, SCHEME>(wc[e], we);
                gq[e] += tj - tn;
                tj = tn;
            }
        } else {
            // x with 2, w with 1 element on each side; for the halo lanes the outer elements are clamped (results unused)
            T xw[VEC + 4], ww[VEC + 2];
#pragma unroll
            for (int e = 0; e < VEC; ++e) { xw[e + 2] = xc[e]; ww[e + 1] = wc[e]; }
            const int xl2 = xcol - 2 >= 0 ? -2 : -xcol, xr2 = xcol + VEC + 1 < PX ? VEC + 1 : PX - 1 - xcol;
            xw[0] = xrow[xl2]; xw[1] = xrow[-1]; xw[VEC + 2] = xrow[VEC]; xw[VEC + 3] = xrow[xr2];
            ww[0] = wrow[wl]; ww[VEC + 1] = wrow[wr];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int gj = c.j0 + tp.cj + e;
                const T cm = cen_exists<T>(gj - 1, P.Nj) * ((xw[e + 2] - xw[e]) * ww[e]);
                const T cp = cen_exists<T>(gj + 1, P.Nj) * ((xw[e + 4] - xw[e + 2]) * ww[e + 2]);
                gq[e] += cm - cp;
            }
        }
        // ---- time
        if (T_ON) {
            const int dm = tp.fl > 0 ? -1 : 0, dp = tp.fl < g.FC - 1 ? 1 : 0;          // neighbouring frames, clamped
            T xm[VEC], xp[VEC], wm[VEC], wp[VEC], wq[VEC], v[VEC];
            ld_into<T, VEC>(wm, wrow + dm * g.slotW);
            ld_into<T, VEC>(wp, wrow + dp * g.slotW);
#pragma unroll
            for (int e = 0; e < VEC; ++e) wq[e] = wc[e];
            if constexpr (TSMODE == 2) {   // along t the inverse norms travel with their voxel's scale (strip_quad_G_impl)
                T f[VEC];
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t + dm);
#pragma unroll
                for (int e = 0; e < VEC; ++e) wm[e] *= f[e];
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t);
#pragma unroll
                for (int e = 0; e < VEC; ++e) wq[e] *= f[e];
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t + dp);
#pragma unroll
                for (int e = 0; e < VEC; ++e) wp[e] *= f[e];
            }
            if (!CEN) {
                ld_into<T, VEC>(xm, xrow + dm * g.slotX);
                ld_into<T, VEC>(xp, xrow + dp * g.slotX);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    v[e] = P.srt * ((xc[e] - xm[e]) * pair_w<T, SCHEME>(wm[e], wq[e]) - (xp[e] - xc[e]) * pair_w<T, SCHEME>(wq[e], wp[e]));
            } else {
                // C_t(t-1) = exists(t-1) (x(t) - x(t-2)) w(t-1);  C_t(t+1) = exists(t+1) (x(t+2) - x(t)) w(t+1)
                const int dm2 = tp.fl > 1 ? -2 : -tp.fl, dp2 = tp.fl < g.FC - 2 ? 2 : g.FC - 1 - tp.fl;
                ld_into<T, VEC>(xm, xrow + dm2 * g.slotX);
                ld_into<T, VEC>(xp, xrow + dp2 * g.slotX);
                const T am = P.srt * cen_exists<T>(t - 1, P.M), ap = P.srt * cen_exists<T>(t + 1, P.M);
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = am * ((xc[e] - xm[e]) * wm[e]) - ap * ((xp[e] - xc[e]) * wp[e]);
            }
            if constexpr (TSMODE >= 1) {
                if (c.Ms) {
                    const uint8_t* m = c.Ms + (rr + 1) * WJ + wcol;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[e] *= m[e] ? P.sfac : T(1);
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) gq[e] += v[e];
        }
        if (Z_ON) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) st.g[r][e] += gq[e];
        } else {
            if (tp.col_out && gi < P.Ni) {
                Pack<T, VEC> pk;
#pragma unroll
                for (int e = 0; e < VEC; ++e) pk.v[e] = gq[e] * k2;
                st_pack<T, VEC>(G + (long long)p * P.sZ + tp.goff + (long long)r * P.Nj, pk);
            }
        }
    }
}

}  // namespace pytvb
