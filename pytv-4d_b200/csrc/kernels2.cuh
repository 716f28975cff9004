// Generation-2 launch machinery for the Chambolle-Pock passes (see strip_core.cuh for the why).
//
// CTA tile: TW quads wide x (TRT thread-rows x R rows per thread) tall, one (z, t) plane.  CTAs are numbered
// band-major like generation 1 (all (z, t) planes of a 64-row band before the next band) so that the z / t
// neighbour planes stay in L2.  The block decode is warp-uniform; per thread there is only
// (row, quad column) and a 32-bit element offset.
#pragma once
#include "kernels.cuh"
#include "strip_core.cuh"

#ifndef PYTVB_DUAL_MINB
#define PYTVB_DUAL_MINB 2     // min resident CTAs per SM requested for the dual strip kernel (register cap)
#endif
#ifndef PYTVB_PRIMAL_MINB
#define PYTVB_PRIMAL_MINB 3
#endif
#ifndef PYTVB_TVNORM_MINB
#define PYTVB_TVNORM_MINB 4   // TV sweeps: 64 registers, 4 CTAs per SM (measured against 3 / uncapped, profiles/r01zd_*)
#endif
#ifndef PYTVB_TVGRAD_MINB
#define PYTVB_TVGRAD_MINB 4
#endif

namespace pytvb {

// Base tiling over strips: Ni is replaced by the number of R-row strips.
template <int R>
inline Tiling make_strip_tiling(int Nj, int Ni, int M, int nz, int vec, int z_lo = 0) {
    const int nstrips = (Ni + R - 1) / R;
    Tiling t = make_tiling(Nj, nstrips, M, z_lo, nz, vec);
    // keep a band at about BAND_ROWS image rows
    int rb = BAND_ROWS / (t.TR * R);
    if (rb < 1) rb = 1;
    const int nrb = (nstrips + t.TR - 1) / t.TR;
    if (rb > nrb) rb = nrb;
    t.RB = rb;
    t.nbands = (nrb + rb - 1) / rb;
    t.nblocks = (long long)t.ncb * t.RB * M * nz * t.nbands;
    t.d_RB = make_fastdiv((unsigned)t.RB);
    return t;
}


template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, typename YT = T, bool TS = false, bool MIR = false>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_DUAL_MINB) cp_dual_strip_kernel(ImgView<T> Xb, YT* __restrict__ y, double* __restrict__ partial, Params<T> P, T sig,
                                                                    T lam, Tiling tl, MirrorBufs<YT> mb = MirrorBufs<YT>{nullptr, nullptr}) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);   // q.i = strip index
    T l21 = T(0);
    if (q.active) {
        const DualPlane<T, YT> pl = make_dual_plane<T, SCHEME, YT>(Xb, y, P, q.z, q.t);
        MirrorPlanes<YT> mir{nullptr, nullptr};
        if constexpr (MIR) mir = mirror_planes<YT>(mb, P, q.z, q.t);
        const int i0 = q.i * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            if (i < P.Ni) {
                const int o = i * P.Nj + q.j0;
                const int o_up = i > 0 ? o - P.Nj : o;
                const int o_dn = i < P.Ni - 1 ? o + P.Nj : o;
                l21 += strip_quad_cp_dual<T, VEC, SCHEME, Z_ON, T_ON, YT, TS, MIR>(pl, P, i, q.j0, o, o_up, o_dn, sig, lam, mir);
            }
        }
    }
    if (partial) {
        const double bs = block_sum((double)l21 * (double)P.inv_div);
        store_or_finish(bs, partial, P);
    }
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int VARIANT, int R, typename YT = T, bool TS = false, bool MIR = false>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_PRIMAL_MINB) cp_primal_strip_kernel(FieldView<YT> Y, T* __restrict__ x, T* __restrict__ aux, const T* __restrict__ x0,
                                                                      double* __restrict__ partial, Params<T> P, T tau, T c1, T c2, Tiling tl, T tau_x0 = T(-1),
                                                                      MirrorBufs<T> mb = MirrorBufs<T>{nullptr, nullptr}) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    T fid = T(0);
    if (q.active) {
        const PrimalPlane<T, YT> pl = make_primal_plane<T, SCHEME, Z_ON, T_ON, YT>(Y, P, q.z, q.t);
        MirrorPlanes<T> mir{nullptr, nullptr};
        if constexpr (MIR) mir = mirror_planes<T>(mb, P, q.z, q.t);
        const int i0 = q.i * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            if (i < P.Ni) {
                const int o = i * P.Nj + q.j0;
                const int o_up = i > 0 ? o - P.Nj : o;
                const int o_dn = i < P.Ni - 1 ? o + P.Nj : o;
                fid += strip_quad_cp_primal<T, VEC, SCHEME, Z_ON, T_ON, VARIANT, false, YT, TS, MIR>(x, aux, x0, pl, P, i, q.j0, o, o_up, o_dn, tau, c1, c2, tau_x0, mir);
            }
        }
    }
    if (partial) {
        const double bs = block_sum((double)fid);
        store_or_finish(bs, partial, P);
    }
}

// Row loop shared by every strip kernel: calls f(i, o, o_up, o_dn) for the R rows of this thread's strip.
template <int R, typename F>
__device__ __forceinline__ void for_strip_rows(const QuadIdx& q, int Ni, int Nj, F f) {
    const int i0 = q.i * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r;
        if (i < Ni) {
            const int o = i * Nj + q.j0;
            f(i, o, i > 0 ? o - Nj : o, i < Ni - 1 ? o + Nj : o);
        }
    }
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, bool TS = false>
__global__ void __launch_bounds__(CTA_THREADS) D_strip_kernel(ImgView<T> X, T* __restrict__ D, Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    if (!q.active) return;
    const DualPlane<T> pl = make_dual_plane<T, SCHEME>(X, D, P, q.z, q.t);   // pl.y = component 0 of D at plane (z, t)
    for_strip_rows<R>(q, P.Ni, P.Nj, [&](int i, int o, int o_up, int o_dn) {
        strip_quad_D<T, VEC, SCHEME, Z_ON, T_ON, TS>(pl.y, pl, P, i, q.j0, o, o_up, o_dn);
    });
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, bool TS = false>
__global__ void __launch_bounds__(CTA_THREADS) DT_strip_kernel(FieldView<T> Pf, T* __restrict__ out, Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    if (!q.active) return;
    const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z_ON, T_ON>(Pf, P, q.z, q.t);
    for_strip_rows<R>(q, P.Ni, P.Nj, [&](int i, int o, int o_up, int o_dn) {
        T v[VEC];
        strip_quad_DT<T, VEC, SCHEME, Z_ON, T_ON, false, T, TS>(v, pl, P, i, q.j0, o, o_up, o_dn);
        Pack<T, VEC> pk;
#pragma unroll
        for (int e = 0; e < VEC; ++e) pk.v[e] = v[e];
        st_pack<T, VEC>(out + pl.img + o, pk);
    });
}

// ND: the number of components as a compile-time constant (the schemes' 2, 3, 4, 6, 8: the loads of a row are then issued together
// instead of one per trip of a dependent loop - C3 one-sided, Nd = 3: 0.66 of the copy peak in the runtime form), or 0 = runtime.
template <typename T, int VEC, int R, int ND = 0>
__global__ void __launch_bounds__(CTA_THREADS) l21_strip_kernel(const T* __restrict__ D, int Nd_rt, T* __restrict__ norms, double* __restrict__ partial,
                                                                Params<T> P, Tiling tl) {
    const int Nd = ND > 0 ? ND : Nd_rt;
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    T sum = T(0);
    if (q.active) {
        const T* base = D + (long long)q.z * P.sZf + (long long)q.t * P.sT;
        T* nbase = norms ? norms + (long long)q.z * P.sZ + (long long)q.t * P.sT : nullptr;
        for_strip_rows<R>(q, P.Ni, P.Nj, [&](int, int o, int, int) {
            T s[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) s[e] = T(0);
            if constexpr (ND > 0) {
                Pack<T, VEC> v[ND];
#pragma unroll
                for (int k = 0; k < ND; ++k) v[k] = ld_pack<T, VEC>(base + (long long)k * P.sC + o);
#pragma unroll
                for (int k = 0; k < ND; ++k)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) s[e] += v[k].v[e] * v[k].v[e];
            } else {
                for (int k = 0; k < Nd; ++k) {
                    const Pack<T, VEC> v = ld_pack<T, VEC>(base + (long long)k * P.sC + o);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) s[e] += v.v[e] * v.v[e];
                }
            }
            Pack<T, VEC> nr;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                nr.v[e] = pytvb_sqrt(s[e]);
                sum += nr.v[e];
            }
            if (nbase) st_pack<T, VEC>(nbase + o, nr);
        });
    }
    const double bs = block_sum((double)sum);
    store_or_finish(bs, partial, P);
}

// TV sweep 1 (strip form): z range tl.z_lo .. tl.z_lo+tl.nz-1 includes one halo plane per side when present.
// FAC: the time component carries a per-voxel factor (mask_static / weight map); the launcher picks the variant.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, bool TS = false, bool FAC = true>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_TVNORM_MINB) tv_norm_strip_kernel(ImgView<T> X, T* __restrict__ Wz0, T* __restrict__ norms, double* __restrict__ partial,
                                                                    Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    T sum = T(0);
    if (q.active) {
        const DualPlane<T> pl = make_dual_plane<T, SCHEME>(X, Wz0, P, q.z, q.t);
        const long long img = (long long)q.z * P.sZ + (long long)q.t * P.sT;
        const bool own = q.z >= 0 && q.z < P.Nz;
        T* np = (norms && own) ? norms + img : nullptr;
        if constexpr (SCHEME == CENTRAL) {
            const T v = strip_rows_tv_norm_central<T, VEC, Z_ON, T_ON, R, TS && T_ON, T_ON>(Wz0 ? Wz0 + img : nullptr, np, pl, P, q.i * R, q.j0);
            if (own) sum = v;
        } else {
            const T v = strip_rows_tv_norm<T, VEC, SCHEME, Z_ON, T_ON, R, TS, FAC>(Wz0 ? Wz0 + img : nullptr, np, pl, P, q.i * R, q.j0);
            if (own) sum = v;
        }
    }
    const double bs = block_sum((double)sum);
    store_or_finish(bs, partial, P);
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, bool TS = false, bool FAC = true>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_TVGRAD_MINB) tv_grad_strip_kernel(ImgView<T> X, ImgView<T> W, T* __restrict__ G, Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    if (!q.active) return;
    const GradPlane<T> pl = make_grad_plane<T, SCHEME>(X, W, P, q.z, q.t);
    T* gp = G + (long long)q.z * P.sZ + (long long)q.t * P.sT;
    if constexpr (SCHEME == CENTRAL)
        strip_rows_G_central<T, VEC, Z_ON, T_ON, R, TS>(gp, pl, P, q.i * R, q.j0);
    else
        strip_rows_G<T, VEC, SCHEME, Z_ON, T_ON, R, TS, FAC>(gp, pl, P, q.i * R, q.j0, tl.TW >= 32);
}

}  // namespace pytvb
