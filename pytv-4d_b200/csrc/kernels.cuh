// sm_100a kernels of the TV hot path (generation 1: one thread per quad, direct neighbour loads).
//
// Work decomposition shared by every kernel here:
//   * a thread owns one quad (VEC consecutive voxels of a row, one 128-bit access when VEC*sizeof(T)=16);
//   * a CTA of 256 threads owns TR rows x TW quads of one (z, t) plane;
//   * CTAs are numbered so that, for a band of BAND_ROWS image rows, ALL (z, t) planes are visited before
//     the next band starts.  The z+-1 / t+-1 neighbour rows of a quad (and, for the adjoint, the
//     neighbouring planes of the dual field) were therefore touched a few hundred CTAs earlier and are
//     still in the 126 MB L2 even when one z-plane group is larger than L2 (M=8, N=2048: 134 MB).
//   * reductions (TV value, L21, data-fidelity) are accumulated per thread in T, per CTA in double, written
//     to a per-CTA slot and summed by a fixed-order second stage: deterministic, no float atomics.
#pragma once
#include <cuda_runtime.h>
#include "tv_core.cuh"

namespace pytvb {

constexpr int CTA_THREADS = 256;
constexpr int BAND_ROWS = 64;

// Division by a runtime constant as multiply-high + shift (valid for 0 <= n < 2^31): the block-index decode
// below would otherwise spend ~30 instructions per integer division.
struct FastDiv {
    unsigned d, mul, shr;
    __host__ __device__ __forceinline__ void divmod(unsigned n, unsigned& q, unsigned& r) const {
#if defined(__CUDA_ARCH__)
        q = (d != 1) ? (__umulhi(n, mul) >> shr) : n;
#else
        q = (d != 1) ? (unsigned)(((unsigned long long)n * mul) >> 32) >> shr : n;
#endif
        r = n - q * d;
    }
};
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d; f.mul = 0; f.shr = 0;
    if (d > 1) {
        unsigned lg = 0;
        while ((1ull << lg) < d) ++lg;              // ceil(log2 d)
        const unsigned p = 31 + lg;
        f.mul = (unsigned)(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
    }
    return f;
}

// Thread -> quad mapping.
struct Tiling {
    int W;          // quads per row
    int TW, TR;     // CTA tile: TW quads x TR rows (TW*TR == CTA_THREADS)
    int ncb;        // column blocks per row
    int RB;         // row blocks per band
    int nbands;
    int z_lo, nz;   // z range covered: z_lo .. z_lo+nz-1 (z_lo = -1 when a halo plane is processed too)
    int M;
    long long nblocks;
    FastDiv d_ncb, d_RB, d_M, d_nz;
    int tw_shift;   // TW == 1 << tw_shift
};

inline Tiling make_tiling(int Nj, int Ni, int M, int z_lo, int nz, int vec) {
    Tiling t;
    t.W = (Nj + vec - 1) / vec;
    int tw = 8;
    while (tw < t.W && tw < CTA_THREADS) tw <<= 1;
    t.TW = tw;
    t.TR = CTA_THREADS / tw;
    t.ncb = (t.W + t.TW - 1) / t.TW;
    int rb = BAND_ROWS / t.TR;
    if (rb < 1) rb = 1;
    const int nrb = (Ni + t.TR - 1) / t.TR;
    if (rb > nrb) rb = nrb;
    t.RB = rb;
    t.nbands = (nrb + rb - 1) / rb;
    t.z_lo = z_lo;
    t.nz = nz;
    t.M = M;
    t.nblocks = (long long)t.ncb * t.RB * M * nz * t.nbands;
    t.d_ncb = make_fastdiv((unsigned)t.ncb);
    t.d_RB = make_fastdiv((unsigned)t.RB);
    t.d_M = make_fastdiv((unsigned)M);
    t.d_nz = make_fastdiv((unsigned)nz);
    t.tw_shift = 0;
    while ((1 << t.tw_shift) < t.TW) ++t.tw_shift;
    return t;
}

struct QuadIdx {
    int z, t, i, j0;
    bool active;
};

__device__ __forceinline__ QuadIdx decode_quad(const Tiling& tl, int Ni, int vec) {
    QuadIdx q;
    const int tq = threadIdx.x & (tl.TW - 1), tr = threadIdx.x >> tl.tw_shift;
    unsigned b = blockIdx.x, cb, rbi, t, zi;
    tl.d_ncb.divmod(b, b, cb);
    tl.d_RB.divmod(b, b, rbi);
    tl.d_M.divmod(b, b, t);
    tl.d_nz.divmod(b, b, zi);
    q.t = (int)t;
    q.z = tl.z_lo + (int)zi;
    const int band = (int)b;
    q.i = (band * tl.RB + (int)rbi) * tl.TR + tr;
    const int qi = (int)cb * tl.TW + tq;
    q.j0 = qi * vec;
    q.active = (q.i < Ni) && (qi < tl.W);
    return q;
}

// CTA-wide sum in double; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[CTA_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();   // protects warp_part when block_sum is called twice
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < CTA_THREADS / 32; ++w) s += warp_part[w];
    }
    return s;
}

// Second stage of the reductions: `n` partials -> `nout` partials (contiguous chunks, fixed order).
static __global__ void __launch_bounds__(CTA_THREADS) reduce_chunks_kernel(const double* __restrict__ in, long long n, double* __restrict__ out, double scale) {
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n) hi = n;
    double s = 0.0;
    for (long long k = lo + threadIdx.x; k < hi; k += CTA_THREADS) s += in[k];
    s = block_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s * scale;
}

// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
__global__ void __launch_bounds__(CTA_THREADS) D_kernel(ImgView<T> X, T* __restrict__ D, Params<T> P, Tiling tl) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    if (!q.active) return;
    T d[C::ND][VEC];
    quad_D<T, VEC, SCHEME, Z_ON, T_ON>(d, X, P, q.z, q.t, q.i, q.j0);
    T* o = D + (long long)q.z * P.sZf + (long long)q.t * P.sT + (long long)q.i * P.Nj + q.j0;
#pragma unroll
    for (int k = 0; k < C::ND; ++k) {
        Pack<T, VEC> pk;
#pragma unroll
        for (int e = 0; e < VEC; ++e) pk.v[e] = d[k][e];
        st_pack<T, VEC>(o + (long long)k * P.sC, pk);
    }
}

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
__global__ void __launch_bounds__(CTA_THREADS) DT_kernel(FieldView<T> Pf, T* __restrict__ out, Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    if (!q.active) return;
    T o[VEC];
    quad_DT<T, VEC, SCHEME, Z_ON, T_ON>(o, Pf, P, q.z, q.t, q.i, q.j0);
    Pack<T, VEC> pk;
#pragma unroll
    for (int e = 0; e < VEC; ++e) pk.v[e] = o[e];
    st_pack<T, VEC>(out + (long long)q.z * P.sZ + (long long)q.t * P.sT + (long long)q.i * P.Nj + q.j0, pk);
}

// L2,1 norm of a field with a runtime number of components (compute_L21_norm, tv_operators_GPU.py:46).
template <typename T, int VEC>
__global__ void __launch_bounds__(CTA_THREADS) l21_kernel(const T* __restrict__ D, int Nd, T* __restrict__ norms, double* __restrict__ partial,
                                                          Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    T sum = T(0);
    if (q.active) {
        const long long off = (long long)q.t * P.sT + (long long)q.i * P.Nj + q.j0;
        const T* p = D + (long long)q.z * P.sZf + off;
        T s[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) s[e] = T(0);
        for (int k = 0; k < Nd; ++k) {
            const Pack<T, VEC> v = ld_pack<T, VEC>(p + (long long)k * P.sC);
#pragma unroll
            for (int e = 0; e < VEC; ++e) s[e] += v.v[e] * v.v[e];
        }
        Pack<T, VEC> nr;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            nr.v[e] = pytvb_sqrt(s[e]);
            sum += nr.v[e];
        }
        if (norms) st_pack<T, VEC>(norms + (long long)q.z * P.sZ + off, nr);
    }
    const double bs = block_sum((double)sum);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
}

// img[~mask] = 0 in place (tv_GPU.py:79-80).  mask: one byte per voxel, either the full (Nz,M,Ni,Nj)
// volume or a single (Ni,Nj) plane broadcast over z and t.
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) apply_mask_kernel(T* __restrict__ x, const uint8_t* __restrict__ mask, long long V,
                                                                 long long plane, int mask_is_plane) {
    const long long stride = (long long)gridDim.x * CTA_THREADS;
    for (long long k = (long long)blockIdx.x * CTA_THREADS + threadIdx.x; k < V; k += stride) {
        const uint8_t m = mask_is_plane ? mask[k % plane] : mask[k];
        if (!m) x[k] = T(0);
    }
}

// TV, sweep 1: inverse gradient norm w (0 where the norm is 0) for every plane the sub-gradient sweep
// reads (the slab plus one plane each side when halos exist), the TV partial sums and, on request, the
// norm array with the reference's infs (tv_GPU.py:85-88).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
__global__ void __launch_bounds__(CTA_THREADS) tv_norm_kernel(ImgView<T> X, T* __restrict__ Wbase /* plane z=0 */, T* __restrict__ norms,
                                                              double* __restrict__ partial, Params<T> P, Tiling tl) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    T sum = T(0);
    if (q.active) {
        T d[C::ND][VEC], nr[VEC];
        quad_D<T, VEC, SCHEME, Z_ON, T_ON>(d, X, P, q.z, q.t, q.i, q.j0);
        quad_norm<T, VEC, C::ND>(nr, d);
        const long long off = (long long)q.z * P.sZ + (long long)q.t * P.sT + (long long)q.i * P.Nj + q.j0;
        Pack<T, VEC> w, no;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            w.v[e] = nr[e] > T(0) ? T(1) / nr[e] : T(0);
            no.v[e] = nr[e] > T(0) ? nr[e] : T(INFINITY);
        }
        st_pack<T, VEC>(Wbase + off, w);
        if (q.z >= 0 && q.z < P.Nz) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) sum += nr[e];
            if (norms) st_pack<T, VEC>(norms + off, no);
        }
    }
    const double bs = block_sum((double)sum);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
}

// TV, sweep 2: sub-gradient from x and w.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
__global__ void __launch_bounds__(CTA_THREADS) tv_grad_kernel(ImgView<T> X, ImgView<T> W, T* __restrict__ G, Params<T> P, Tiling tl) {
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    if (!q.active) return;
    T g[VEC];
    quad_G<T, VEC, SCHEME, Z_ON, T_ON>(g, X, W, P, q.z, q.t, q.i, q.j0);
    Pack<T, VEC> pk;
#pragma unroll
    for (int e = 0; e < VEC; ++e) pk.v[e] = g[e];
    st_pack<T, VEC>(G + (long long)q.z * P.sZ + (long long)q.t * P.sT + (long long)q.i * P.Nj + q.j0, pk);
}

// Chambolle-Pock dual pass (pass A of the iteration).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON>
__global__ void __launch_bounds__(CTA_THREADS) cp_dual_kernel(ImgView<T> Xb, T* __restrict__ y, double* __restrict__ partial, Params<T> P, T sigma,
                                                              T inv_lam, Tiling tl) {
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    T l21 = T(0);
    if (q.active) l21 = quad_cp_dual<T, VEC, SCHEME, Z_ON, T_ON>(y, Xb, P, sigma, inv_lam, q.z, q.t, q.i, q.j0);
    if (partial) {
        const double bs = block_sum((double)l21);
        if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    }
}

// Chambolle-Pock primal pass (pass B).  VARIANT 0: ROF prox + over-relaxation (aux = xbar, out);
// VARIANT 1: README loop (aux = y_f, in/out).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int VARIANT>
__global__ void __launch_bounds__(CTA_THREADS) cp_primal_kernel(FieldView<T> Y, T* __restrict__ x, T* __restrict__ aux, const T* __restrict__ x0,
                                                                double* __restrict__ partial, Params<T> P, T tau, T c2, Tiling tl) {
    const QuadIdx q = decode_quad(tl, P.Ni, VEC);
    T fid = T(0);
    if (q.active) {
        if (VARIANT == 0)
            fid = quad_cp_primal_rof<T, VEC, SCHEME, Z_ON, T_ON>(x, aux, x0, Y, P, tau, c2, q.z, q.t, q.i, q.j0);
        else
            fid = quad_cp_primal_readme<T, VEC, SCHEME, Z_ON, T_ON>(x, aux, x0, Y, P, tau, c2, q.z, q.t, q.i, q.j0);
    }
    if (partial) {
        const double bs = block_sum((double)fid);
        if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    }
}

}  // namespace pytvb
