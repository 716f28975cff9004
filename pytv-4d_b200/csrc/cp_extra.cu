// Chambolle-Pock iteration, optional form: half-precision storage of the dual field.  Split from cp.cu to keep the
// translation units compiling in parallel.
#include <stdlib.h>

#include "host_common.cuh"
#include "kernels2.cuh"

#ifndef PYTVB_STRIP_R
#define PYTVB_STRIP_R 4   // image rows walked by one thread in the generation-2 kernels
#endif

using namespace pytvb;

namespace {

// ---- half-precision storage of the dual field (float32 images; the field holds y / lam) ---------------------
struct DualHArgs { ImgView<float> Xb; __half* y; double* partial; Params<float> P; float sig, lam_proj; cudaStream_t st; long long* nb; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchDualH {
    static int run(const DualHArgs& a) {
        constexpr int R = PYTVB_STRIP_R;
        const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
        if (int rc = check_grid(tl)) return rc;
        if (TT && a.P.tscale) cp_dual_strip_kernel<float, VEC, SCHEME, Z, TT, R, __half, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.Xb, a.y, a.partial, a.P, a.sig, a.lam_proj, tl);
        else cp_dual_strip_kernel<float, VEC, SCHEME, Z, TT, R, __half, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(a.Xb, a.y, a.partial, a.P, a.sig, a.lam_proj, tl);
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        *a.nb = tl.nblocks;
        return PYTVB_OK;
    }
};
struct PrimalHArgs { FieldView<__half> Y; float* x; float* xbar; const float* x0; double* partial; Params<float> P; float tau_y, tau, c1, theta; cudaStream_t st; long long* nb; };
template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchPrimalH {
    static int run(const PrimalHArgs& a) {
        constexpr int R = PYTVB_STRIP_R;
        const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
        if (int rc = check_grid(tl)) return rc;
        if (TT && a.P.tscale) cp_primal_strip_kernel<float, VEC, SCHEME, Z, TT, 0, R, __half, TT><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
            a.Y, a.x, a.xbar, a.x0, a.partial, a.P, a.tau_y, a.c1, a.theta, tl, a.tau);
        else cp_primal_strip_kernel<float, VEC, SCHEME, Z, TT, 0, R, __half, false><<<(unsigned)tl.nblocks, CTA_THREADS, 0, a.st>>>(
            a.Y, a.x, a.xbar, a.x0, a.partial, a.P, a.tau_y, a.c1, a.theta, tl, a.tau);
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        *a.nb = tl.nblocks;
        return PYTVB_OK;
    }
};
// vector width: 4 floats / 4 halves per quad need 16-byte images and 8-byte field pointers
inline int pick_vec_h(const pytvb_problem* pb, std::initializer_list<const void*> f32, std::initializer_list<const void*> f16) {
    if (pb->Nj % 4 != 0) return 1;
    if (pb->time_scale && (reinterpret_cast<uintptr_t>(pb->time_scale) % 16) != 0) return 1;
    for (const void* p : f32) if (p && (reinterpret_cast<uintptr_t>(p) % 16) != 0) return 1;
    for (const void* p : f16) if (p && (reinterpret_cast<uintptr_t>(p) % 8) != 0) return 1;
    return 4;
}

}  // namespace

extern "C" {

int pytvb_cp_dual_f16y(const pytvb_problem* pb, const void* xbar, void* y_half, double lam, double sigma, double* d_l21_or_null,
                       const void* halo_lo, const void* halo_hi, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(pb->dtype == PYTVB_F32, "half-precision dual storage needs float32 images");
    PYTVB_REQUIRE(xbar && y_half, "xbar and y must not be NULL");
    PYTVB_REQUIRE(!d_l21_or_null || ws, "a reduction workspace is required when d_l21 is requested");
    PYTVB_REQUIRE(lam > 0, "lam must be > 0 (the field stores y / lam)");
    const Axes ax = axes_of(pb);
    if (int rc = check_halos(pb, ax.z_on, false, halo_lo, halo_hi)) return rc;
    DualHArgs a;
    a.Xb = ImgView<float>{(const float*)xbar, (const float*)halo_lo, (const float*)halo_hi, 1};
    a.y = (__half*)y_half;
    a.partial = d_l21_or_null ? reduce_partials(ws) : nullptr;
    a.P = make_params<float>(pb);
    a.sig = (float)(sigma / lam) * a.P.inv_div;      // (y + sigma D)/lam = y/lam + (sigma/lam) D
    a.lam_proj = 1.0f;                               // projection onto the unit ball
    a.st = (cudaStream_t)stream;
    long long nb = 0;
    a.nb = &nb;
    const int vec = pick_vec_h(pb, {xbar, halo_lo, halo_hi}, {y_half});
    if (int rc = dispatch<LaunchDualH, float>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    return d_l21_or_null ? finalize_sum(a.partial, nb, d_l21_or_null, a.st) : PYTVB_OK;
}

int pytvb_cp_primal_rof_f16y(const pytvb_problem* pb, const void* y_half, void* x, void* xbar, const void* x0, double lam, double tau, double theta,
                             double* d_fid_or_null, const void* halo_lo_half, const void* halo_hi_half, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(pb->dtype == PYTVB_F32, "half-precision dual storage needs float32 images");
    PYTVB_REQUIRE(y_half && x && xbar && x0, "y, x, xbar and x0 must not be NULL");
    PYTVB_REQUIRE(!d_fid_or_null || ws, "a reduction workspace is required when d_fid is requested");
    const Axes ax = axes_of(pb);
    if (int rc = check_halos(pb, ax.z_on, true, halo_lo_half, halo_hi_half)) return rc;
    PrimalHArgs a;
    a.Y = FieldView<__half>{(const __half*)y_half, (const __half*)halo_lo_half, (const __half*)halo_hi_half};
    a.x = (float*)x; a.xbar = (float*)xbar; a.x0 = (const float*)x0;
    a.partial = d_fid_or_null ? reduce_partials(ws) : nullptr;
    a.P = make_params<float>(pb);
    a.tau_y = (float)(tau * lam);                    // D^T y = lam D^T (y / lam)
    a.tau = (float)tau;
    a.c1 = (float)(1.0 / (1.0 + tau));
    a.theta = (float)theta;
    a.st = (cudaStream_t)stream;
    long long nb = 0;
    a.nb = &nb;
    const int vec = pick_vec_h(pb, {x, xbar, x0}, {y_half, halo_lo_half, halo_hi_half});
    if (int rc = dispatch<LaunchPrimalH, float>(vec, pb->scheme, ax.z_on, ax.t_on, a)) return rc;
    return d_fid_or_null ? finalize_sum(a.partial, nb, d_fid_or_null, a.st) : PYTVB_OK;
}

}  // extern "C"
