// Operator entry points: D, D_T, L21, mask.
#include <stdlib.h>

#include "host_common.cuh"
#include "kernels2.cuh"

#ifndef PYTVB_STRIP_R
#define PYTVB_STRIP_R 4
#endif

using namespace pytvb;

namespace {

template <typename T> struct DArgs { ImgView<T> X; T* D; Params<T> P; int vec; cudaStream_t st; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchD {
    static int run(const DArgs<T>& a) {
        {
            {
                constexpr int R = PYTVB_STRIP_R;
                const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
                if (int rc = check_grid(tl)) return rc;
                if (TT && a.P.tscale) D_strip_kernel<T, VEC, SCHEME, Z, TT, R, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.D, a.P, tl);
                else D_strip_kernel<T, VEC, SCHEME, Z, TT, R, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.X, a.D, a.P, tl);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                return PYTVB_OK;
            }
        }
    }
};

template <typename T> struct DTArgs { FieldView<T> F; T* out; Params<T> P; cudaStream_t st; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchDT {
    static int run(const DTArgs<T>& a) {
        {
            {
                constexpr int R = PYTVB_STRIP_R;
                const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
                if (int rc = check_grid(tl)) return rc;
                if (TT && a.P.tscale) DT_strip_kernel<T, VEC, SCHEME, Z, TT, R, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.F, a.out, a.P, tl);
                else DT_strip_kernel<T, VEC, SCHEME, Z, TT, R, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.F, a.out, a.P, tl);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                return PYTVB_OK;
            }
        }
    }
};

template <typename T>
int run_D(const pytvb_problem* pb, const void* x, void* D, const void* lo, const void* hi, cudaStream_t st) {
    const Axes ax = axes_of(pb);
    DArgs<T> a;
    a.X = ImgView<T>{(const T*)x, (const T*)lo, (const T*)hi, 1};
    a.D = (T*)D;
    a.P = make_params<T>(pb);
    a.st = st;
    const int vec = pick_vec<T>(pb, {x, D, lo, hi});
    return dispatch<LaunchD, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
}

template <typename T>
int run_DT(const pytvb_problem* pb, const void* p, void* out, const void* lo, const void* hi, cudaStream_t st) {
    const Axes ax = axes_of(pb);
    DTArgs<T> a;
    a.F = FieldView<T>{(const T*)p, (const T*)lo, (const T*)hi};
    a.out = (T*)out;
    a.P = make_params<T>(pb);
    a.st = st;
    const int vec = pick_vec<T>(pb, {p, out, lo, hi});
    return dispatch<LaunchDT, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
}

template <typename T>
int run_l21(const pytvb_problem* pb, const void* D, int Nd, void* norms, double* d_sum, void* ws, cudaStream_t st) {
    Params<T> P = make_params<T>(pb);
    P.sZf = P.sC * Nd;
    arm_reduction(P, ws, d_sum);
    const int vec = pick_vec<T>(pb, {D, norms});
    double* partial = reduce_partials(ws);
    {
        constexpr int R = PYTVB_STRIP_R;
        const Tiling tl = make_strip_tiling<R>(P.Nj, P.Ni, P.M, P.Nz, vec);
        if (int rc = check_grid(tl)) return rc;
#define PYTVB_L21(ND) l21_strip_kernel<T, VecOf<T>::value, R, ND><<<(unsigned)tl.nblocks, CTA_THREADS, 0, st>>>((const T*)D, Nd, (T*)norms, partial, P, tl)
        if (vec > 1) {
            switch (Nd) {       // the schemes' component counts get their own instantiation, any other field the runtime loop
                case 2: PYTVB_L21(2); break;
                case 3: PYTVB_L21(3); break;
                case 4: PYTVB_L21(4); break;
                case 6: PYTVB_L21(6); break;
                case 8: PYTVB_L21(8); break;
                default: PYTVB_L21(0); break;
            }
        } else {
            l21_strip_kernel<T, 1, R><<<(unsigned)tl.nblocks, CTA_THREADS, 0, st>>>((const T*)D, Nd, (T*)norms, partial, P, tl);
        }
#undef PYTVB_L21
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        return finish_reduction(partial, tl.nblocks, d_sum, st);
    }
}

}  // namespace

extern "C" {

int pytvb_D(const pytvb_problem* pb, const void* x, void* D, const void* halo_lo, const void* halo_hi, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x && D, "x and D must not be NULL");
    if (int rc = check_halos(pb, axes_of(pb).z_on, false, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_D<float>(pb, x, D, halo_lo, halo_hi, st) : run_D<double>(pb, x, D, halo_lo, halo_hi, st);
}

int pytvb_DT(const pytvb_problem* pb, const void* p, void* out, const void* halo_lo, const void* halo_hi, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(p && out, "p and out must not be NULL");
    if (int rc = check_halos(pb, axes_of(pb).z_on, true, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_DT<float>(pb, p, out, halo_lo, halo_hi, st) : run_DT<double>(pb, p, out, halo_lo, halo_hi, st);
}

int pytvb_l21(const pytvb_problem* pb, const void* D, int64_t Nd, void* norms_or_null, double* d_sum, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(D && d_sum && ws, "D, d_sum and ws must not be NULL");
    PYTVB_REQUIRE(Nd >= 1 && Nd <= 4096, "Nd = %lld out of range", (long long)Nd);
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_l21<float>(pb, D, (int)Nd, norms_or_null, d_sum, ws, st)
                                  : run_l21<double>(pb, D, (int)Nd, norms_or_null, d_sum, ws, st);
}

int pytvb_gd_update(const pytvb_problem* pb, void* x, const void* x0, const void* G, double step, double lam, double* d_fid_or_null, void* ws,
                    void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x && x0 && G, "x, x0 and G must not be NULL");
    PYTVB_REQUIRE(!d_fid_or_null || ws, "a reduction workspace is required when d_fid is requested");
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)pb->Ni * pb->Nj * pb->M * pb->Nz;
    long long nb = (V + CTA_THREADS - 1) / CTA_THREADS;
    if (nb > 148 * 8) nb = 148 * 8;
    double* partial = d_fid_or_null ? reduce_partials(ws) : nullptr;
    unsigned* counter = d_fid_or_null ? reduce_counter(ws) : nullptr;       // the sum is finished by the last CTA: one launch
    if (pb->dtype == PYTVB_F32)
        gd_update_kernel<float><<<(unsigned)nb, CTA_THREADS, 0, st>>>((float*)x, (const float*)x0, (const float*)G, V, (float)step, (float)lam, partial, counter,
                                                                      d_fid_or_null);
    else
        gd_update_kernel<double><<<(unsigned)nb, CTA_THREADS, 0, st>>>((double*)x, (const double*)x0, (const double*)G, V, step, lam, partial, counter, d_fid_or_null);
    count_launches(1);
    PYTVB_CUDA(cudaGetLastError());
    return PYTVB_OK;
}

int pytvb_apply_mask(const pytvb_problem* pb, void* x, const uint8_t* mask, int mask_is_plane, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x && mask, "x and mask must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const long long plane = (long long)pb->Ni * pb->Nj, planes = (long long)pb->M * pb->Nz, V = plane * planes;
    const size_t es = pb->dtype == PYTVB_F32 ? 4 : 8;
    const bool quads = pb->Nj % 4 == 0 && reinterpret_cast<uintptr_t>(x) % (4 * es < 16 ? 4 * es : 16) == 0 && reinterpret_cast<uintptr_t>(mask) % 4 == 0;
    if (quads) {
        const long long nq = plane / 4;
        const dim3 grid((unsigned)((nq + CTA_THREADS - 1) / CTA_THREADS), (unsigned)(planes < 64 ? planes : 64));
        if (pb->dtype == PYTVB_F32) apply_mask_quad_kernel<float><<<grid, CTA_THREADS, 0, st>>>((float*)x, mask, nq, planes, mask_is_plane);
        else apply_mask_quad_kernel<double><<<grid, CTA_THREADS, 0, st>>>((double*)x, mask, nq, planes, mask_is_plane);
    } else {
        long long nb = (V + CTA_THREADS - 1) / CTA_THREADS;
        if (nb > 148 * 32) nb = 148 * 32;
        if (pb->dtype == PYTVB_F32)
            apply_mask_kernel<float><<<(unsigned)nb, CTA_THREADS, 0, st>>>((float*)x, mask, V, plane, mask_is_plane);
        else
            apply_mask_kernel<double><<<(unsigned)nb, CTA_THREADS, 0, st>>>((double*)x, mask, V, plane, mask_is_plane);
    }
    count_launches(1);
    PYTVB_CUDA(cudaGetLastError());
    return PYTVB_OK;
}

}  // extern "C"
