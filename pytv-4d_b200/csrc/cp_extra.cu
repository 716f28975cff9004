// Chambolle-Pock iteration, optional forms: the single-launch iteration (generation 3) and half-precision storage of the
// dual field.  Split from cp.cu to keep the translation units compiling in parallel.
#include <stdlib.h>

#include "host_common.cuh"
#include "kernels2.cuh"
#include "kernels3.cuh"

#ifndef PYTVB_STRIP_R
#define PYTVB_STRIP_R 4   // image rows walked by one thread in the generation-2 kernels
#endif

using namespace pytvb;

namespace {

// ---- generation 3: one launch per iteration -----------------------------------------------------------
template <typename T> struct FusedArgs {
    ImgView<T> Xin; FieldView<T> Y; T* y; T* x; T* aux; const T* x0; Params<T> P; T sigma, lam, tau, c2; int variant; int lag;
    double* d_l21; double* d_fid; void* ws; cudaStream_t st;
};

inline int fused_lag() {
    const char* e = getenv("PYTVB_FUSED_LAG");
    const int v = e ? atoi(e) : 3;
    return v < 1 ? 1 : v;
}

// workspace: [partials A][partials B][stage 2][ticket, error, done[nbands*Nz]]
template <int R>
inline size_t fused_ws_bytes(const pytvb_problem* pb, int vec, FusedSched* out = nullptr) {
    const Tiling tl = make_strip_tiling<R>((int)pb->Nj, (int)pb->Ni, (int)pb->M, (int)pb->Nz, vec);
    const FusedSched s = make_fused_sched(tl, (int)pb->Nz, ((int)pb->Ni + R - 1) / R, fused_lag());
    if (out) *out = s;
    const size_t per_phase = (size_t)(s.total / 2);
    return (2 * per_phase + REDUCE_STAGE2) * sizeof(double) + (2 + (size_t)s.nbands * s.Nz) * sizeof(unsigned) + 64;
}

template <typename T, int VEC, int SCHEME, bool Z, bool TT> struct LaunchFused {
    static int run(const FusedArgs<T>& a) {
        constexpr int R = PYTVB_STRIP_R;
        const Tiling tl = make_strip_tiling<R>(a.P.Nj, a.P.Ni, a.P.M, a.P.Nz, VEC);
        const FusedSched s = make_fused_sched(tl, a.P.Nz, (a.P.Ni + R - 1) / R, a.lag);
        PYTVB_REQUIRE(s.total > 0 && s.total < 2147483647LL, "grid of %lld CTAs is out of range", s.total);
        const size_t per_phase = (size_t)(s.total / 2);
        double* pA = (double*)a.ws;
        double* pB = pA + per_phase;
        unsigned* ctl_mem = (unsigned*)(pB + per_phase + REDUCE_STAGE2);
        FusedCtl ctl{ctl_mem, ctl_mem + 2, ctl_mem + 1};
        PYTVB_CUDA(cudaMemsetAsync(ctl_mem, 0, (2 + (size_t)s.nbands * s.Nz) * sizeof(unsigned), a.st));
        const T c1 = T(1) / (T(1) + (a.variant == 0 ? a.tau : a.c2));
        const T sig = a.sigma * a.P.inv_div;
        if (a.variant == 0)
            cp_fused_kernel<T, VEC, SCHEME, Z, TT, 0, R, false><<<(unsigned)s.total, CTA_THREADS, 0, a.st>>>(
                a.Xin, a.Y, a.y, a.x, a.aux, a.x0, a.d_l21 ? pA : nullptr, a.d_fid ? pB : nullptr, a.P, sig, a.lam, a.tau, c1, a.c2, s, ctl);
        else
            cp_fused_kernel<T, VEC, SCHEME, Z, TT, 1, R, false><<<(unsigned)s.total, CTA_THREADS, 0, a.st>>>(
                a.Xin, a.Y, a.y, a.x, a.aux, a.x0, a.d_l21 ? pA : nullptr, a.d_fid ? pB : nullptr, a.P, sig, a.lam, a.tau, c1, a.c2, s, ctl);
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        // reductions (stage-2 scratch sits behind the two partial arrays; the two finalisations run one after the other)
        if (a.d_l21) {
            if (int rc = finalize_sum_at(pA, (long long)per_phase, pB + per_phase, a.d_l21, a.st)) return rc;
        }
        if (a.d_fid) {
            if (int rc = finalize_sum_at(pB, (long long)per_phase, pB + per_phase, a.d_fid, a.st)) return rc;
        }
        fused_check_kernel<<<1, 1, 0, a.st>>>(ctl.error, a.d_l21, a.d_fid);
        count_launches(1);
        PYTVB_CUDA(cudaGetLastError());
        return PYTVB_OK;
    }
};

template <typename T>
int run_fused(const pytvb_problem* pb, int variant, const void* xin, void* y, void* x, void* aux, const void* x0, double lam, double sigma, double tau,
              double c2, double* d_l21, double* d_fid, const void* img_lo, const void* img_hi, const void* fld_lo, const void* fld_hi, void* ws,
              cudaStream_t st) {
    const Axes ax = axes_of(pb);
    FusedArgs<T> a;
    a.Xin = ImgView<T>{(const T*)xin, (const T*)img_lo, (const T*)img_hi, 1};
    a.Y = FieldView<T>{(const T*)y, (const T*)fld_lo, (const T*)fld_hi};
    a.y = (T*)y; a.x = (T*)x; a.aux = (T*)aux; a.x0 = (const T*)x0;
    a.P = make_params<T>(pb);
    a.sigma = (T)sigma; a.lam = (T)lam; a.tau = (T)tau; a.c2 = (T)c2;
    a.variant = variant;
    a.lag = fused_lag();
    a.d_l21 = d_l21; a.d_fid = d_fid; a.ws = ws; a.st = st;
    const int vec = pick_vec<T>(pb, {xin, y, x, aux, x0, img_lo, img_hi, fld_lo, fld_hi});
    return dispatch<LaunchFused, T>(vec, pb->scheme, ax.z_on, ax.t_on, a);
}

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

size_t pytvb_fused_workspace_bytes(const pytvb_problem* pb) {
    if (check_problem(pb) != PYTVB_OK) return 0;
    // the scalar tiling has the most tiles
    return fused_ws_bytes<PYTVB_STRIP_R>(pb, 1);
}

int pytvb_cp_iter_fused(const pytvb_problem* pb, int variant, const void* u, void* y, void* x, void* aux, const void* x0, double lam, double sigma,
                        double tau, double c2, double* d_l21_or_null, double* d_fid_or_null, const void* img_halo_lo, const void* img_halo_hi,
                        const void* fld_halo_lo, const void* fld_halo_hi, void* ws, void* stream) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(u && y && x && aux && x0 && ws, "u, y, x, aux, x0 and ws must not be NULL");
    PYTVB_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (rof) or 1 (readme)");
    PYTVB_REQUIRE(!pb->time_scale, "time_scale is not supported by the single-launch iteration");
    PYTVB_REQUIRE(lam >= 0, "lam must be >= 0");
    if (int rc = check_halos(pb, axes_of(pb).z_on, false, img_halo_lo, img_halo_hi)) return rc;
    if (int rc = check_halos(pb, axes_of(pb).z_on, true, fld_halo_lo, fld_halo_hi)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return pb->dtype == PYTVB_F32
               ? run_fused<float>(pb, variant, u, y, x, aux, x0, lam, sigma, tau, c2, d_l21_or_null, d_fid_or_null, img_halo_lo, img_halo_hi, fld_halo_lo,
                                  fld_halo_hi, ws, st)
               : run_fused<double>(pb, variant, u, y, x, aux, x0, lam, sigma, tau, c2, d_l21_or_null, d_fid_or_null, img_halo_lo, img_halo_hi, fld_halo_lo,
                                   fld_halo_hi, ws, st);
}

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
    a.partial = d_l21_or_null ? (double*)ws : nullptr;
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
    a.partial = d_fid_or_null ? (double*)ws : nullptr;
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
