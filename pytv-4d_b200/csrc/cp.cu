// Fused Chambolle-Pock iteration: dual pass (A) and primal pass (B).
#include <stdlib.h>

#include "host_common.cuh"
#include "kernels2.cuh"

#ifndef PYTVB_STRIP_R
#define PYTVB_STRIP_R 4   // image rows walked by one thread in the generation-2 kernels
#endif

using namespace pytvb;

namespace {

template <typename T> struct DualArgs { ImgView<T> Xb; T* y; double* partial; Params<T> P; T sigma, inv_lam, lam; cudaStream_t st; long long* nb; T* mir_prev; T* mir_next; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchDual {
    static int run(const DualArgs<T>& a) {
        {
            {
                constexpr int R = PYTVB_STRIP_R;
                const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
                if (int rc = check_grid(tl)) return rc;
                if (Z && (a.mir_prev || a.mir_next)) {
                    if (TT && a.P.tscale)
                        cp_dual_strip_kernel<T, VEC, SCHEME, Z, TT, R, T, TT, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Xb, a.y, a.partial, a.P, a.sigma * a.P.inv_div, a.lam, tl, MirrorBufs<T>{a.mir_prev, a.mir_next});
                    else
                        cp_dual_strip_kernel<T, VEC, SCHEME, Z, TT, R, T, false, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Xb, a.y, a.partial, a.P, a.sigma * a.P.inv_div, a.lam, tl, MirrorBufs<T>{a.mir_prev, a.mir_next});
                } else if (TT && a.P.tscale) cp_dual_strip_kernel<T, VEC, SCHEME, Z, TT, R, T, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                    a.Xb, a.y, a.partial, a.P, a.sigma * a.P.inv_div, a.lam, tl);
                else cp_dual_strip_kernel<T, VEC, SCHEME, Z, TT, R, T, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                    a.Xb, a.y, a.partial, a.P, a.sigma * a.P.inv_div, a.lam, tl);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                *a.nb = tl.nblocks;
                return PYTVB_OK;
            }
        }
    }
};

template <typename T> struct PrimalArgs {
    FieldView<T> Y; T* x; T* aux; const T* x0; double* partial; Params<T> P; T tau, c2; int variant; cudaStream_t st; long long* nb; T* mir_prev; T* mir_next;
};
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchPrimal {
    static int run(const PrimalArgs<T>& a) {
        {
            {
                constexpr int R = PYTVB_STRIP_R;
                const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
                if (int rc = check_grid(tl)) return rc;
                const T c1 = T(1) / (T(1) + (a.variant == 0 ? a.tau : a.c2));
                if (Z && (a.mir_prev || a.mir_next)) {
                    const MirrorBufs<T> mb{a.mir_prev, a.mir_next};
                    const bool ts = TT && a.P.tscale;
                    if (a.variant == 0 && ts)
                        cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 0, R, T, TT, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl, T(-1), mb);
                    else if (a.variant == 0)
                        cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 0, R, T, false, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl, T(-1), mb);
                    else if (ts)
                        cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 1, R, T, TT, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl, T(-1), mb);
                    else
                        cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 1, R, T, false, Z><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                            a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl, T(-1), mb);
                } else if (a.variant == 0)
                    if (TT && a.P.tscale) cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 0, R, T, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                        a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl);
                    else cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 0, R, T, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                        a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl);
                else
                    if (TT && a.P.tscale) cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 1, R, T, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                        a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl);
                    else cp_primal_strip_kernel<T, VEC, SCHEME, Z, TT, 1, R, T, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
                        a.Y, a.x, a.aux, a.x0, a.partial, a.P, a.tau, c1, a.c2, tl);
                count_launches(1);
                PYTVB_CUDA(cudaGetLastError());
                *a.nb = tl.nblocks;
                return PYTVB_OK;
            }
        }
    }
};

template <typename T>
int run_dual(const pytvb_problem* pb, const void* xbar, void* y, double lam, double sigma, double* d_l21, const void* lo, const void* hi, void* ws,
             cudaStream_t st, void* mir_prev = nullptr, void* mir_next = nullptr) {
    const Axes ax = axes_of(pb);
    DualArgs<T> a;
    a.Xb = ImgView<T>{(const T*)xbar, (const T*)lo, (const T*)hi, 1};
    a.y = (T*)y;
    a.partial = d_l21 ? reduce_partials(ws) : nullptr;
    a.P = make_params<T>(pb);
    arm_reduction(a.P, ws, d_l21);
    a.sigma = (T)sigma;
    a.inv_lam = (T)(1.0 / lam);
    a.lam = (T)lam;
    a.mir_prev = (T*)mir_prev; a.mir_next = (T*)mir_next;
    a.st = st;
    long long nb = 0;
    a.nb = &nb;
    const int vec = pick_vec<T>(pb, {xbar, y, lo, hi, mir_prev, mir_next});
    if (int rc = dispatch<LaunchDual, T>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    return d_l21 ? finish_reduction(a.partial, nb, d_l21, st) : PYTVB_OK;
}

template <typename T>
int run_primal(const pytvb_problem* pb, int variant, const void* y, void* x, void* aux, const void* x0, double tau, double c2, double* d_fid,
               const void* lo, const void* hi, void* ws, cudaStream_t st, void* mir_prev = nullptr, void* mir_next = nullptr) {
    const Axes ax = axes_of(pb);
    PrimalArgs<T> a;
    a.Y = FieldView<T>{(const T*)y, (const T*)lo, (const T*)hi};
    a.x = (T*)x; a.aux = (T*)aux; a.x0 = (const T*)x0;
    a.partial = d_fid ? reduce_partials(ws) : nullptr;
    a.P = make_params<T>(pb);
    arm_reduction(a.P, ws, d_fid);
    a.tau = (T)tau; a.c2 = (T)c2;
    a.variant = variant;
    a.mir_prev = (T*)mir_prev; a.mir_next = (T*)mir_next;
    a.st = st;
    long long nb = 0;
    a.nb = &nb;
    const int vec = pick_vec<T>(pb, {y, x, aux, x0, lo, hi, mir_prev, mir_next});
    if (int rc = dispatch<LaunchPrimal, T>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    return d_fid ? finish_reduction(a.partial, nb, d_fid, st) : PYTVB_OK;
}

}  // namespace

extern "C" {

int pytvb_cp_dual(const pytvb_problem* pb, const void* xbar, void* y, double lam, double sigma, double* d_l21_or_null, const void* halo_lo,
                  const void* halo_hi, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(xbar && y, "xbar and y must not be NULL");
    PYTVB_REQUIRE(!d_l21_or_null || ws, "a reduction workspace is required when d_l21 is requested");
    PYTVB_REQUIRE(lam >= 0, "lam must be >= 0");
    if (int rc = check_halos(pb, axes_of(pb).z_on, false, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_dual<float>(pb, xbar, y, lam, sigma, d_l21_or_null, halo_lo, halo_hi, ws, st)
                                  : run_dual<double>(pb, xbar, y, lam, sigma, d_l21_or_null, halo_lo, halo_hi, ws, st);
}

int pytvb_cp_dual_p2p(const pytvb_problem* pb, const void* xbar, void* y, double lam, double sigma, double* d_l21_or_null, const void* halo_lo,
                      const void* halo_hi, void* mirror_prev, void* mirror_next, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(xbar && y, "xbar and y must not be NULL");
    PYTVB_REQUIRE(!d_l21_or_null || ws, "a reduction workspace is required when d_l21 is requested");
    PYTVB_REQUIRE(lam >= 0, "lam must be >= 0");
    if (int rc = check_halos(pb, axes_of(pb).z_on, false, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_dual<float>(pb, xbar, y, lam, sigma, d_l21_or_null, halo_lo, halo_hi, ws, st, mirror_prev, mirror_next)
                                  : run_dual<double>(pb, xbar, y, lam, sigma, d_l21_or_null, halo_lo, halo_hi, ws, st, mirror_prev, mirror_next);
}

int pytvb_cp_primal_p2p(const pytvb_problem* pb, int variant, const void* y, void* x, void* aux, const void* x0, double tau, double c2,
                        double* d_fid_or_null, const void* halo_lo, const void* halo_hi, void* mirror_prev, void* mirror_next, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (rof) or 1 (readme)");
    PYTVB_REQUIRE(y && x && aux && x0, "y, x, x0 and the auxiliary image must not be NULL");
    PYTVB_REQUIRE(!d_fid_or_null || ws, "a reduction workspace is required when d_fid is requested");
    if (int rc = check_halos(pb, axes_of(pb).z_on, true, halo_lo, halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_primal<float>(pb, variant, y, x, aux, x0, tau, c2, d_fid_or_null, halo_lo, halo_hi, ws, st, mirror_prev, mirror_next)
                                  : run_primal<double>(pb, variant, y, x, aux, x0, tau, c2, d_fid_or_null, halo_lo, halo_hi, ws, st, mirror_prev, mirror_next);
}

static int primal_common(const pytvb_problem* pb, int variant, const void* y, void* x, void* aux, const void* x0, double tau, double c2,
                         double* d_fid, const void* lo, const void* hi, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(y && x && aux && x0, "y, x, x0 and the auxiliary image must not be NULL");
    PYTVB_REQUIRE(!d_fid || ws, "a reduction workspace is required when d_fid is requested");
    if (int rc = check_halos(pb, axes_of(pb).z_on, true, lo, hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32 ? run_primal<float>(pb, variant, y, x, aux, x0, tau, c2, d_fid, lo, hi, ws, st)
                                  : run_primal<double>(pb, variant, y, x, aux, x0, tau, c2, d_fid, lo, hi, ws, st);
}

int pytvb_cp_primal_rof(const pytvb_problem* pb, const void* y, void* x, void* xbar, const void* x0, double tau, double theta,
                        double* d_fid_or_null, const void* halo_lo, const void* halo_hi, void* ws, void* stream) {
    return primal_common(pb, 0, y, x, xbar, x0, tau, theta, d_fid_or_null, halo_lo, halo_hi, ws, stream);
}

int pytvb_cp_primal_readme(const pytvb_problem* pb, const void* y_tv, void* x, void* y_f, const void* x0, double tau, double sigma_A,
                           double* d_fid_or_null, const void* halo_lo, const void* halo_hi, void* ws, void* stream) {
    return primal_common(pb, 1, y_tv, x, y_f, x0, tau, sigma_A, d_fid_or_null, halo_lo, halo_hi, ws, stream);
}

}  // extern "C"
