// Launch machinery shared by the sm_100a kernels of the TV hot path: tiling, block-index decode, reductions.
//
// Work decomposition of the strip kernels (kernels2.cuh):
//   * a thread owns one quad column (VEC consecutive voxels of a row, one 128-bit access when VEC*sizeof(T)=16) and
//     walks R consecutive rows of it;
//   * a CTA of 256 threads owns TR thread-rows x TW quads of one (z, t) plane;
//   * CTAs are numbered so that, for a band of BAND_ROWS image rows, ALL (z, t) planes are visited before
//     the next band starts.  The z+-1 / t+-1 neighbour rows of a quad (and, for the adjoint, the
//     neighbouring planes of the dual field) were therefore touched a few hundred CTAs earlier and are
//     still in the 126 MB L2 even when one z-plane group is larger than L2 (M=8, N=2048: 134 MB).
//   * reductions (TV value, L21, data-fidelity) are accumulated per thread in T, per CTA in double, written
//     to a per-CTA slot and summed by a fixed-order second stage: deterministic, no float atomics.
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"

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

// Reduction finished inside the producing kernel: every CTA stores its partial; the CTA that arrives last at the counter sums all
// partials in a fixed order (thread k takes partials k, k + blockDim, ...; then the warps in order) and writes the scalar - no
// second launch, which is what a launch-bound volume pays for (a 256^2 sub-gradient descent iteration: 19 -> 12 us).  The value
// does not depend on which CTA is last.  `counter` must be 0 on entry; atomicInc wraps it back to 0 with the last arrival, so a
// workspace zeroed once stays usable.  block_value: this CTA's partial, valid in thread 0.  counter == nullptr: store the partial only.
__device__ __forceinline__ void finish_partials(double block_value, double* __restrict__ partial, unsigned* counter, double* __restrict__ d_out) {
    __shared__ bool is_last;
    __shared__ double warp_sum[32];
    const int tid = threadIdx.x;
    if (tid == 0) {
        partial[blockIdx.x] = block_value;
        bool last = false;
        if (counter) {
            __threadfence();
            last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
        }
        is_last = last;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double v = 0.0;
    for (unsigned k = tid; k < gridDim.x; k += blockDim.x) v += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) s += warp_sum[w];
        *d_out = s;
    }
}

// Strip kernels: a small grid finishes its sum in the kernel, a large one leaves the partials to the two-stage reduction
// (host_common.cuh::finish_reduction applies the same rule).  Called by all threads of the CTA; block_value valid in thread 0.
constexpr unsigned REDUCE_IN_KERNEL_MAX = 8192;
template <typename PT>
__device__ __forceinline__ void store_or_finish(double block_value, double* __restrict__ partial, const PT& P) {
    if (P.red_out && gridDim.x <= REDUCE_IN_KERNEL_MAX) finish_partials(block_value, partial, P.red_counter, P.red_out);
    else if (threadIdx.x == 0) partial[blockIdx.x] = block_value;
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

// img[~mask] = 0 in place (tv_GPU.py:79-80).  mask: one byte per voxel, either the full (Nz,M,Ni,Nj)
// volume or a single (Ni,Nj) plane broadcast over z and t.  Scalar form (any shape / alignment).
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) apply_mask_kernel(T* __restrict__ x, const uint8_t* __restrict__ mask, long long V,
                                                                 long long plane, int mask_is_plane) {
    const long long stride = (long long)gridDim.x * CTA_THREADS;
    for (long long k = (long long)blockIdx.x * CTA_THREADS + threadIdx.x; k < V; k += stride) {
        const uint8_t m = mask_is_plane ? mask[k % plane] : mask[k];
        if (!m) x[k] = T(0);
    }
}

// The same for row lengths divisible by 4 and aligned pointers: a thread owns four consecutive pixels of the (Ni, Nj) plane and
// walks the (z, t) planes (blockIdx.y strided).  The image is never READ - zeros are stored where the mask is 0 - and with a plane
// mask a thread whose four pixels are all kept leaves after one 4-byte load: the DRAM traffic is the zeros written (C5, disc of
// radius 0.48 N: 28 % of 8.6 GB instead of a byte read per voxel through one-byte loads; 2.77 -> 0.4 ms, profiles/r02z_launches_c5.csv).
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) apply_mask_quad_kernel(T* __restrict__ x, const uint8_t* __restrict__ mask, long long nquads,
                                                                      long long planes, int mask_is_plane) {
    const long long q = (long long)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (q >= nquads) return;
    const long long plane = nquads * 4;
    uchar4 m = *reinterpret_cast<const uchar4*>(mask + q * 4);
    if (mask_is_plane && m.x && m.y && m.z && m.w) return;
    for (long long pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        if (!mask_is_plane) m = *reinterpret_cast<const uchar4*>(mask + pl * plane + q * 4);
        T* p = x + pl * plane + q * 4;
        if (!(m.x | m.y | m.z | m.w)) {
            Pack<T, 4> z;
            z.v[0] = z.v[1] = z.v[2] = z.v[3] = T(0);
            if (sizeof(T) == 4) st_pack<T, 4>(p, z);
            else { st_pack<T, 2>(p, Pack<T, 2>{{T(0), T(0)}}); st_pack<T, 2>(p + 2, Pack<T, 2>{{T(0), T(0)}}); }
        } else {
            if (!m.x) p[0] = T(0);
            if (!m.y) p[1] = T(0);
            if (!m.z) p[2] = T(0);
            if (!m.w) p[3] = T(0);
        }
    }
}

// Sub-gradient descent update of the README loop (README.md:120-123), fused: x <- x - step * ((x - x0) + lam * G) and the
// fidelity partial sums sum (x_new - x0)^2, one read of x, x0, G and one write of x.
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) gd_update_kernel(T* __restrict__ x, const T* __restrict__ x0, const T* __restrict__ G, long long V, T step,
                                                                T lam, double* __restrict__ partial, unsigned* counter, double* __restrict__ d_out) {
    const long long stride = (long long)gridDim.x * CTA_THREADS;
    T fid = T(0);
    for (long long k = (long long)blockIdx.x * CTA_THREADS + threadIdx.x; k < V; k += stride) {
        const T xo = x[k], d0 = x0[k];
        const T xn = xo - step * ((xo - d0) + lam * G[k]);
        x[k] = xn;
        const T r = xn - d0;
        fid += r * r;
    }
    if (partial) {
        const double bs = block_sum((double)fid);
        finish_partials(bs, partial, counter, d_out);
    }
}

}  // namespace pytvb
