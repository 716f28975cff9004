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

namespace pytvb {

// Base tiling over strips: Ni is replaced by the number of R-row strips.
template <int R>
inline Tiling make_strip_tiling(int Nj, int Ni, int M, int nz, int vec) {
    const int nstrips = (Ni + R - 1) / R;
    Tiling t = make_tiling(Nj, nstrips, M, 0, nz, vec);
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

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_DUAL_MINB) cp_dual_strip_kernel(ImgView<T> Xb, T* __restrict__ y, double* __restrict__ partial, Params<T> P, T sig,
                                                                    T lam, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);   // q.i = strip index
    T l21 = T(0);
    if (q.active) {
        const DualPlane<T> pl = make_dual_plane<T, SCHEME>(Xb, y, P, q.z, q.t);
        const int i0 = q.i * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            if (i < P.Ni) {
                const int o = i * P.Nj + q.j0;
                const int o_up = i > 0 ? o - P.Nj : o;
                const int o_dn = i < P.Ni - 1 ? o + P.Nj : o;
                l21 += strip_quad_cp_dual<T, VEC, SCHEME, Z_ON, T_ON>(pl, P, i, q.j0, o, o_up, o_dn, sig, lam);
            }
        }
    }
    if (partial) {
        const double bs = block_sum((double)l21 * (double)P.inv_div);
        if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    }
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int VARIANT, int R>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_PRIMAL_MINB) cp_primal_strip_kernel(FieldView<T> Y, T* __restrict__ x, T* __restrict__ aux, const T* __restrict__ x0,
                                                                      double* __restrict__ partial, Params<T> P, T tau, T c1, T c2, Tiling tl) {
    const QuadIdx q = decode_quad(tl, (P.Ni + R - 1) / R, VEC);
    T fid = T(0);
    if (q.active) {
        const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z_ON, T_ON>(Y, P, q.z, q.t);
        const int i0 = q.i * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            if (i < P.Ni) {
                const int o = i * P.Nj + q.j0;
                const int o_up = i > 0 ? o - P.Nj : o;
                const int o_dn = i < P.Ni - 1 ? o + P.Nj : o;
                fid += strip_quad_cp_primal<T, VEC, SCHEME, Z_ON, T_ON, VARIANT>(x, aux, x0, pl, P, i, q.j0, o, o_up, o_dn, tau, c1, c2);
            }
        }
    }
    if (partial) {
        const double bs = block_sum((double)fid);
        if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    }
}

}  // namespace pytvb
