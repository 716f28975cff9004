// Host-side plumbing shared by the translation units of libpytv_b200.so: error reporting, translation of
// the C-ABI problem descriptor into kernel parameters, template dispatch, reduction finalisation.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pytv_b200.h"
#include "kernels.cuh"

namespace pytvb {

void set_error(const char* fmt, ...);   // api.cu
void count_launches(int n);             // api.cu: kernels launched by this library (pytvb_launch_count)

#define PYTVB_CUDA(call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return PYTVB_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

#define PYTVB_REQUIRE(cond, ...)     \
    do {                             \
        if (!(cond)) {               \
            set_error(__VA_ARGS__);  \
            return PYTVB_ERR_ARG;    \
        }                            \
    } while (0)

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int value = 4; };
template <> struct VecOf<double> { static constexpr int value = 2; };

struct Axes {
    bool z_on, t_on;
    int Nd;
};

inline Axes axes_of(const pytvb_problem* pb) {
    Axes a;
    a.z_on = pb->Nz_global > 1 && pb->reg_z_over_reg > 0;   // NaN fails `> 0` like the reference
    a.t_on = pb->M > 1 && pb->reg_time > 0;
    a.Nd = pb->scheme == PYTVB_HYBRID ? 4 + 2 * a.z_on + 2 * a.t_on : 2 + a.z_on + a.t_on;
    return a;
}

inline int check_problem(const pytvb_problem* pb) {
    PYTVB_REQUIRE(pb != nullptr, "problem descriptor is NULL");
    PYTVB_REQUIRE(pb->scheme >= 0 && pb->scheme <= 3, "unknown scheme %d", pb->scheme);
    PYTVB_REQUIRE(pb->dtype == PYTVB_F32 || pb->dtype == PYTVB_F64, "unknown dtype %d", pb->dtype);
    PYTVB_REQUIRE(pb->Nz >= 1 && pb->M >= 1 && pb->Ni >= 1 && pb->Nj >= 1, "empty volume %lld x %lld x %lld x %lld",
                  (long long)pb->Nz, (long long)pb->M, (long long)pb->Ni, (long long)pb->Nj);
    PYTVB_REQUIRE(pb->Nz < (1LL << 30) && pb->M < (1LL << 30) && pb->Ni < (1LL << 30) && pb->Nj < (1LL << 30), "extent too large");
    PYTVB_REQUIRE(pb->Ni * pb->Nj < (1LL << 31) - 8, "one image plane must have fewer than 2^31 pixels (offsets inside a plane are 32-bit)");
    PYTVB_REQUIRE(pb->z_offset >= 0 && pb->z_offset + pb->Nz <= pb->Nz_global, "slab [%lld, %lld) outside 0..%lld",
                  (long long)pb->z_offset, (long long)(pb->z_offset + pb->Nz), (long long)pb->Nz_global);
    PYTVB_REQUIRE(!(pb->factor_reg_static < 0), "factor_reg_static must be >= 0");
    return PYTVB_OK;
}

template <typename T>
inline Params<T> make_params(const pytvb_problem* pb) {
    const Axes a = axes_of(pb);
    Params<T> P;
    P.Nz = (int)pb->Nz; P.M = (int)pb->M; P.Ni = (int)pb->Ni; P.Nj = (int)pb->Nj;
    P.zg0 = pb->z_offset; P.NzG = pb->Nz_global;
    P.z_fwd_fallback = (pb->scheme == PYTVB_CENTRAL && pb->Nz_global == 2) ? 1 : 0;
    P.t_fwd_fallback = (pb->scheme == PYTVB_CENTRAL && pb->M == 2) ? 1 : 0;
    P.srz = a.z_on ? (T)sqrt(pb->reg_z_over_reg) : T(0);
    P.srt = a.t_on ? (T)sqrt(pb->reg_time) : T(0);
    P.sfac = (T)sqrt(pb->factor_reg_static);
    P.div = pb->scheme == PYTVB_HYBRID ? (T)sqrt(2.0) : (pb->scheme == PYTVB_CENTRAL ? T(2) : T(1));
    P.inv_div = T(1) / P.div;
    P.mask_static = a.t_on ? pb->mask_static : nullptr;
    P.tscale = a.t_on ? (const T*)pb->time_scale : nullptr;
    P.sT = (long long)pb->Ni * pb->Nj;
    P.sZ = P.sT * pb->M;
    P.sC = P.sZ;
    P.sZf = P.sC * a.Nd;
    P.red_counter = nullptr;
    P.red_out = nullptr;
    return P;
}

// Widest vector width every pointer and the row length allow.
template <typename T>
inline int pick_vec(const pytvb_problem* pb, std::initializer_list<const void*> ptrs) {
    constexpr int VM = VecOf<T>::value;
    if (pb->Nj % VM != 0) return 1;
    if (pb->time_scale && (reinterpret_cast<uintptr_t>(pb->time_scale) % (VM * sizeof(T))) != 0) return 1;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) % (VM * sizeof(T))) != 0) return 1;
    return VM;
}

inline bool needs_lo(int scheme) { return scheme != PYTVB_UPWIND; }     // image plane z-1 (backward / centred)
inline bool needs_hi(int scheme) { return scheme != PYTVB_DOWNWIND; }   // image plane z+1 (forward / centred)

// Halo presence check for an image-type halo (D, cp_dual, tv) or a field-type halo (DT, cp_primal: roles swap).
inline int check_halos(const pytvb_problem* pb, bool z_on, bool field_type, const void* lo, const void* hi) {
    if (!z_on) return PYTVB_OK;
    const bool interior_lo = pb->z_offset > 0, interior_hi = pb->z_offset + pb->Nz < pb->Nz_global;
    // image: lo needed by backward/centred differences; field adjoint: lo needed by forward/centred components
    const bool need_lo = field_type ? (pb->scheme != PYTVB_DOWNWIND) : needs_lo(pb->scheme);
    const bool need_hi = field_type ? (pb->scheme != PYTVB_UPWIND) : needs_hi(pb->scheme);
    PYTVB_REQUIRE(!(interior_lo && need_lo && !lo), "slab starts at z=%lld inside the volume: halo_lo is required", (long long)pb->z_offset);
    PYTVB_REQUIRE(!(interior_hi && need_hi && !hi), "slab ends at z=%lld inside the volume: halo_hi is required",
                  (long long)(pb->z_offset + pb->Nz));
    return PYTVB_OK;
}

// L<T, VEC, SCHEME, Z_ON, T_ON>::run(args) for the runtime (vec, scheme, z_on, t_on).
template <template <typename, int, int, bool, bool> class L, typename T, typename Args>
inline int dispatch(int vec, int scheme, bool z, bool t, const Args& a) {
    constexpr int VM = VecOf<T>::value;
#define PYTVB_ZT(V, S)                                          \
    (z ? (t ? L<T, V, S, true, true>::run(a) : L<T, V, S, true, false>::run(a)) \
       : (t ? L<T, V, S, false, true>::run(a) : L<T, V, S, false, false>::run(a)))
#define PYTVB_SCH(V)                                 \
    switch (scheme) {                                \
        case PYTVB_UPWIND: return PYTVB_ZT(V, UPWIND);     \
        case PYTVB_DOWNWIND: return PYTVB_ZT(V, DOWNWIND); \
        case PYTVB_CENTRAL: return PYTVB_ZT(V, CENTRAL);   \
        default: return PYTVB_ZT(V, HYBRID);         \
    }
    if (vec == VM) { PYTVB_SCH(VM) } else { PYTVB_SCH(1) }
#undef PYTVB_SCH
#undef PYTVB_ZT
}

inline int check_grid(const Tiling& tl) {
    PYTVB_REQUIRE(tl.nblocks > 0 && tl.nblocks < 2147483647LL, "grid of %lld CTAs is out of range", tl.nblocks);
    return PYTVB_OK;
}

// Workspace layout for reductions: [partials: max CTAs][stage 2: 256 doubles]
constexpr int REDUCE_STAGE2 = 256;
constexpr int REDUCE_HEAD = 8;      // doubles in FRONT of the partials: the arrival counter of finish_partials (at a fixed place, whatever the problem)
inline long long max_partials(const pytvb_problem* pb) {
    // worst case: scalar path with one row per thread, one extra halo plane on each side (tv sweep 1)
    const Tiling tl = make_tiling((int)pb->Nj, (int)pb->Ni, (int)pb->M, 0, (int)pb->Nz + 2, 1);
    return tl.nblocks;
}

// partials[0..n) -> d_out[0]; fixed summation order.  stage2: REDUCE_STAGE2 doubles of scratch.
inline int finalize_sum_at(double* partials, long long n, double* stage2, double* d_out, cudaStream_t st) {
    if (n <= 8192) {
        reduce_chunks_kernel<<<1, CTA_THREADS, 0, st>>>(partials, n, d_out, 1.0);
        count_launches(1);
    } else {
        reduce_chunks_kernel<<<REDUCE_STAGE2, CTA_THREADS, 0, st>>>(partials, n, stage2, 1.0);
        reduce_chunks_kernel<<<1, CTA_THREADS, 0, st>>>(stage2, REDUCE_STAGE2, d_out, 1.0);
        count_launches(2);
    }
    PYTVB_CUDA(cudaGetLastError());
    return PYTVB_OK;
}
// Layout of a reduce workspace: [arrival counter of the in-kernel reductions, REDUCE_HEAD doubles][per-CTA partials][second stage].
// No kernel writes the head except finish_partials, which leaves it zero.
inline unsigned* reduce_counter(void* ws) { return static_cast<unsigned*>(ws); }
inline double* reduce_partials(void* ws) { return ws ? static_cast<double*>(ws) + REDUCE_HEAD : nullptr; }
// Strip kernels: grids of up to REDUCE_IN_KERNEL_MAX CTAs finish their sum inside the kernel (the launch-bound sizes: a second launch
// costs as much as the kernel), larger ones through the two-stage reduction.  arm_reduction before the launch, finish_reduction after.
template <typename T>
inline void arm_reduction(Params<T>& P, void* ws, double* d_out) {
    P.red_counter = d_out ? reduce_counter(ws) : nullptr;
    P.red_out = d_out;
}
inline int finalize_sum(double* partials, long long n, double* d_out, cudaStream_t st);
inline int finish_reduction(double* partials, long long n, double* d_out, cudaStream_t st) {
    return n <= REDUCE_IN_KERNEL_MAX ? PYTVB_OK : finalize_sum(partials, n, d_out, st);
}
inline int finalize_sum(double* partials, long long n, double* d_out, cudaStream_t st) {
    return finalize_sum_at(partials, n, partials + n, d_out, st);
}

}  // namespace pytvb
