// Types shared by every kernel of the TV hot path: packs, component layout, problem parameters, image / field views.
//
// A "quad" is VEC consecutive voxels of one image row (z, t, i, j0 .. j0+VEC-1).  The per-thread code built on these
// types (strip_core.cuh, tile_core.cuh) uses plain pointer loads and is __host__ __device__, so the same code compiles
// as __device__ code for the sm_100a kernels and as host code for the CPU emulation harness under tests/emul/ (test
// infrastructure: it lets the index / boundary / halo logic be checked against the oracle in a container without a
// GPU; it is never linked into the product library).
//
// Layouts (reference: README.md:235, tv_operators_CPU.py:115): images (Nz, M, Ni, Nj), gradient fields
// (Nz, Nd, M, Ni, Nj), both dense C order; j is the fastest axis.
//
// Boundary rule (tv_operators_CPU.py:118,121 write only [:-1]): an out-of-range DIFFERENCE is zero - not
// an out-of-range value - so every edge is an index predicate, never a zero-filled load.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PYTVB_HD __host__ __device__ __forceinline__
#else
#define PYTVB_HD inline
#endif

namespace pytvb {

enum : int { UPWIND = 0, DOWNWIND = 1, CENTRAL = 2, HYBRID = 3 };

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
    T v[VEC];
};

template <typename T, int VEC>
PYTVB_HD Pack<T, VEC> ld_pack(const T* p) {
    return *reinterpret_cast<const Pack<T, VEC>*>(p);
}
template <typename T, int VEC>
PYTVB_HD void st_pack(T* p, const Pack<T, VEC>& v) {
    *reinterpret_cast<Pack<T, VEC>*>(p) = v;
}

PYTVB_HD float pytvb_sqrt(float a) { return sqrtf(a); }
PYTVB_HD double pytvb_sqrt(double a) { return sqrt(a); }

// Compile-time component layout of a scheme (tv_operators_CPU.py:117-152 hybrid, :264-284 others).
template <int SCHEME, bool Z_ON, bool T_ON>
struct Comp {
    static constexpr bool HYB = (SCHEME == HYBRID);
    static constexpr int ND = HYB ? 4 + 2 * Z_ON + 2 * T_ON : 2 + Z_ON + T_ON;
    // forward-type slot and backward-type slot of every axis; for the single-component schemes both
    // names refer to the one slot of that axis
    static constexpr int I_F = 0, J_F = 1;
    static constexpr int I_B = HYB ? 2 : 0, J_B = HYB ? 3 : 1;
    static constexpr int Z_F = HYB ? 4 : 2, Z_B = HYB ? 5 : 2;
    static constexpr int T_F = HYB ? 4 + 2 * Z_ON : 2 + Z_ON, T_B = HYB ? 5 + 2 * Z_ON : 2 + Z_ON;
    static constexpr bool NEED_FWD = (SCHEME != DOWNWIND);   // reads x[k+1]
    static constexpr bool NEED_BWD = (SCHEME != UPWIND);     // reads x[k-1]
};

// Geometry and weights of one call (a whole volume, or one z-slab of it).
template <typename T>
struct Params {
    int Nz, M, Ni, Nj;            // local extents
    long long zg0, NzG;           // global z index of local plane 0, global number of planes
    int z_fwd_fallback;           // central scheme on a z axis of global length 2 -> forward difference
    int t_fwd_fallback;           // same for M == 2 (tv_operators_CPU.py:339, :347)
    T srz, srt, sfac;             // sqrt(reg_z_over_reg), sqrt(reg_time), sqrt(factor_reg_static)
    T div, inv_div;               // global divisor: sqrt(2) hybrid, 2 central, 1 otherwise
    const uint8_t* mask_static;   // (Ni, Nj) bytes, nonzero = static pixel; or null
    const T* tscale;              // (Nz, M, Ni, Nj) per-voxel factor of the time component(s) (sqrt of a weight map); or null
    long long sT, sZ;             // image strides (elements): Ni*Nj, M*Ni*Nj
    long long sC, sZf;            // field strides: component = M*Ni*Nj, plane group = Nd*M*Ni*Nj
    unsigned* red_counter;        // strip kernels with a scalar output: arrival counter and destination of the sum when a small grid
    double* red_out;              // finishes its reduction inside the kernel (kernels.cuh::store_or_finish); null otherwise
};

// Image with optional z-halo planes (multi-GPU slabs).  `lo` holds `depth` planes z = -depth .. -1 in
// increasing z, `hi` holds planes z = Nz .. Nz+depth-1; each plane is (M, Ni, Nj).
template <typename T>
struct ImgView {
    const T* base;
    const T* lo;
    const T* hi;
    int depth;
    PYTVB_HD const T* row(const Params<T>& P, int z, int t, int i) const {
        const long long off = (long long)t * P.sT + (long long)i * P.Nj;
        if (z < 0) return lo + (long long)(z + depth) * P.sZ + off;
        if (z >= P.Nz) return hi + (long long)(z - P.Nz) * P.sZ + off;
        return base + (long long)z * P.sZ + off;
    }
};

// Gradient field with optional one-plane halos for the adjoint: `lo` is the (M, Ni, Nj) plane of the
// z-component that the adjoint reads at z = -1, `hi` the one it reads at z = Nz.
template <typename T>
struct FieldView {
    const T* base;
    const T* lo;
    const T* hi;
    template <typename PT>
    PYTVB_HD const T* row(const PT& P, int z, int comp, int t, int i) const {
        const long long off = (long long)t * P.sT + (long long)i * P.Nj;
        if (z < 0) return lo + off;
        if (z >= P.Nz) return hi + off;
        return base + (long long)z * P.sZf + (long long)comp * P.sC + off;
    }
};
template <typename T, int VEC>
PYTVB_HD void ld_into(T* dst, const T* src) {
    const Pack<T, VEC> p = ld_pack<T, VEC>(src);
#pragma unroll
    for (int e = 0; e < VEC; ++e) dst[e] = p.v[e];
}
template <typename T, int VEC>
PYTVB_HD void zero_into(T* dst) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) dst[e] = T(0);
}

// Factor applied to the time component(s) at static pixels (tv_operators_CPU.py:148-150).
template <typename T, int VEC>
PYTVB_HD void static_factor(T* f, const Params<T>& P, int i, int j0) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = T(1);
    if (P.mask_static) {
        const uint8_t* m = P.mask_static + (long long)i * P.Nj + j0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = m[e] ? P.sfac : T(1);
    }
}

}  // namespace pytvb
