// tv_<scheme>: TV value + sub-gradient (+ gradient norms).
#include <stdlib.h>

#include "host_common.cuh"
#include "kernels2.cuh"
#include "tv_path.cuh"
#include "tv_args.cuh"

#ifndef PYTVB_STRIP_R
#define PYTVB_STRIP_R 8
#endif

using namespace pytvb;

namespace {
// (the single-sweep tile kernel is launched from tv_tile.cu - its own translation unit, so that its kernel variants compile in
// parallel with the two-sweep fallback's: run_tv_tile, declared in tv_args.cuh)
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchTv {
    static int run(const TvArgs<T>& a) {
        {
            {
                constexpr int R = PYTVB_STRIP_R;
                const Tiling t1 = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.nz, VEC, a.z_lo);
                if (int rc = check_grid(t1)) return rc;
                // kernel variants: weight map (TS) | mask_static only (FAC) | uniform time weight; the centred scheme has no FAC form
                constexpr bool CEN = SCHEME == CENTRAL;
                const bool fac = CEN || (TT && a.P.mask_static);
                if (TT && a.P.tscale) tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, TT, true><<<(unsigned)t1.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.Wz0, a.norms, a.partial, a.P, t1);
                else if (fac) tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, true><<<(unsigned)t1.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.Wz0, a.norms, a.partial, a.P, t1);
                else tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, CEN><<<(unsigned)t1.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.Wz0, a.norms, a.partial, a.P, t1);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                *a.nblocks_out = t1.nblocks;
                const Tiling t2 = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
                if (TT && a.P.tscale) tv_grad_strip_kernel<T, VEC, SCHEME, Z, TT, R, TT, true><<<(unsigned)t2.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.W, a.G, a.P, t2);
                else if (fac) tv_grad_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, true><<<(unsigned)t2.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.W, a.G, a.P, t2);
                else tv_grad_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, CEN><<<(unsigned)t2.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.W, a.G, a.P, t2);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                return PYTVB_OK;
            }
        }
    }
};

template <typename T>
int run_tv(const pytvb_problem* pb, const void* x, void* G, void* norms, double* d_tv, const void* lo2, const void* hi2, void* ws_reduce,
           void* ws_tv, cudaStream_t st) {
    const Axes ax = axes_of(pb);
    const bool has_lo = ax.z_on && pb->z_offset > 0, has_hi = ax.z_on && pb->z_offset + pb->Nz < pb->Nz_global;
    TvArgs<T> a;
    a.P = make_params<T>(pb);
    a.X = ImgView<T>{(const T*)x, (const T*)lo2, (const T*)hi2, 2};
    // inverse-norm workspace: (Nz+2) planes, plane z=-1 first; 256-byte aligned
    uintptr_t wp = (reinterpret_cast<uintptr_t>(ws_tv) + 255) & ~uintptr_t(255);
    T* wbuf = reinterpret_cast<T*>(wp);
    a.Wz0 = wbuf + a.P.sZ;
    a.W = ImgView<T>{a.Wz0, wbuf, a.Wz0 + (long long)a.P.Nz * a.P.sZ, 1};
    a.G = (T*)G;
    a.norms = (T*)norms;
    a.partial = reduce_partials(ws_reduce);
    a.z_lo = has_lo ? -1 : 0;
    a.nz = (int)pb->Nz + (has_lo ? 1 : 0) + (has_hi ? 1 : 0);
    a.st = st;
    long long nblocks = 0;
    a.nblocks_out = &nblocks;
    a.TS = ImgView<T>{a.P.tscale, (const T*)pb->time_scale_lo, (const T*)pb->time_scale_hi, 1};
    const int vec = pick_vec<T>(pb, {x, G, norms, lo2, hi2, pb->time_scale_lo, pb->time_scale_hi});
    if (tv_uses_tile(pb)) {
        a.counter = reduce_counter(ws_reduce);
        a.d_out = d_tv;
        return run_tv_tile<T>(vec, pb->scheme, ax.z_on, ax.t_on, a);       // one launch: the last CTA writes d_tv
    } else {
        PYTVB_REQUIRE(!(pb->time_scale && (lo2 || hi2)), "time_scale on slabs with z halos needs the single-sweep kernel (not this problem: > 16 coupled frames or a centred length-2 axis)");
        if (int rc = dispatch<LaunchTv, T>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    }
    return finalize_sum(a.partial, nblocks, d_tv, st);
}

}  // namespace

namespace {
// TV value only: sweep 1 without the inverse-norm output (primal energy / duality gap evaluations).
template <typename T> struct TvValArgs { ImgView<T> X; double* partial; Params<T> P; cudaStream_t st; long long* nb; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchTvVal {
    static int run(const TvValArgs<T>& a) {
        constexpr int R = PYTVB_STRIP_R;
        const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
        if (int rc = check_grid(tl)) return rc;
        constexpr bool CEN = SCHEME == CENTRAL;
        const bool fac = CEN || (TT && a.P.mask_static);
        if (TT && a.P.tscale) tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, TT, true><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.X, (T*)nullptr, (T*)nullptr, a.partial, a.P, tl);
        else if (fac) tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, true><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.X, (T*)nullptr, (T*)nullptr, a.partial, a.P, tl);
        else tv_norm_strip_kernel<T, VEC, SCHEME, Z, TT, R, false, CEN><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.X, (T*)nullptr, (T*)nullptr, a.partial, a.P, tl);
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        *a.nb = tl.nblocks;
        return PYTVB_OK;
    }
};
template <typename T>
int run_tv_value(const pytvb_problem* pb, const void* x, double* d_tv, const void* lo, const void* hi, void* ws, cudaStream_t st) {
    const Axes ax = axes_of(pb);
    TvValArgs<T> a;
    a.X = ImgView<T>{(const T*)x, (const T*)lo, (const T*)hi, 1};
    a.partial = reduce_partials(ws);
    a.P = make_params<T>(pb);
    arm_reduction(a.P, ws, d_tv);
    a.st = st;
    long long nb = 0;
    a.nb = &nb;
    const int vec = pick_vec<T>(pb, {x, lo, hi});
    if (int rc = dispatch<LaunchTvVal, T>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    return finish_reduction(a.partial, nb, d_tv, st);
}
}  // namespace

extern "C" int pytvb_tv_value(const pytvb_problem* pb, const void* x, double* d_tv, const void* halo_lo, const void* halo_hi, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x && d_tv && ws, "x, d_tv and ws must not be NULL");
    if (int rc = check_halos(pb, axes_of(pb).z_on, false, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_tv_value<float>(pb, x, d_tv, halo_lo, halo_hi, ws, st) : run_tv_value<double>(pb, x, d_tv, halo_lo, halo_hi, ws, st);
}

extern "C" int pytvb_tv(const pytvb_problem* pb, const void* x, void* G, void* norms_or_null, double* d_tv, const void* halo_lo2,
                        const void* halo_hi2, void* ws_reduce, void* ws_tv, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x && G && d_tv && ws_reduce && ws_tv, "x, G, d_tv and the workspaces must not be NULL");
    if (pb->time_scale && axes_of(pb).z_on && axes_of(pb).t_on) {
        PYTVB_REQUIRE(!(halo_lo2 && !pb->time_scale_lo), "slab starts inside the volume: time_scale_lo (the scale's plane z = -1) is required");
        PYTVB_REQUIRE(!(halo_hi2 && !pb->time_scale_hi), "slab ends inside the volume: time_scale_hi (the scale's plane z = Nz) is required");
    }
    if (axes_of(pb).z_on) {
        PYTVB_REQUIRE(!(pb->z_offset > 0 && !halo_lo2), "slab starts inside the volume: halo_lo2 (2 planes) is required");
        PYTVB_REQUIRE(!(pb->z_offset + pb->Nz < pb->Nz_global && !halo_hi2), "slab ends inside the volume: halo_hi2 (2 planes) is required");
    }
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_tv<float>(pb, x, G, norms_or_null, d_tv, halo_lo2, halo_hi2, ws_reduce, ws_tv, st)
                                  : run_tv<double>(pb, x, G, norms_or_null, d_tv, halo_lo2, halo_hi2, ws_reduce, ws_tv, st);
}
