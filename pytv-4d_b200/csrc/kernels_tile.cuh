// Single-sweep tv_<scheme> kernel: z-marching CTA tiles (per-thread code and the design in tile_core.cuh).
#pragma once
#include <cuda.h>       // CUtensorMap (type only: the encoder is looked up at run time, tmap.cuh)
#include "kernels.cuh"
#include "tile_core.cuh"

#ifndef PYTVB_TILE_R
#define PYTVB_TILE_R 4          // rows per thread
#endif
#ifndef PYTVB_TILE_WARPS
#define PYTVB_TILE_WARPS 16     // warps per CTA aimed at (frames x strips)
#endif
#ifndef PYTVB_TILE_MINB
#define PYTVB_TILE_MINB 1
#endif

namespace pytvb {

#ifndef PYTVB_TILE_MAXT
#define PYTVB_TILE_MAXT 512
#endif
constexpr int TILE_MAX_THREADS = PYTVB_TILE_MAXT;
constexpr size_t TILE_SMEM_LIMIT = 227 * 1024 - 256;     // opt-in maximum per CTA on sm_100 minus the static reduction scratch

template <typename T>
inline size_t tile_smem_bytes(const TileGeom& g, bool mask) {
    size_t b = 128 + ((size_t)3 * g.xslot + (size_t)g.FC * g.slotW) * sizeof(T);      // 128: the windows start on a 128-byte boundary
    b = (b + 15) & ~size_t(15);
    b += (size_t)g.FC * g.rowsX * (sizeof(long long) + sizeof(int));       // staging tables
    if (mask) b += (size_t)g.RPF * g.WJ;
    return (b + 15) & ~size_t(15);
}

// Geometry for a problem; returns false when the tile kernel cannot take it (too many coupled frames).
template <typename T, int VEC, int R>
inline bool make_tile_geom(TileGeom& g, int Nz, int M, int Ni, int Nj, bool t_on, bool mask, int sm_count = 148) {
    g.FC = t_on ? M : 1;
    g.WJ = 32 * VEC;
    g.TJ = g.WJ - 2 * VEC;
    g.pitchX = g.WJ + 2 * VEC;
    if (g.FC * 32 > TILE_MAX_THREADS) return false;
    int strips = PYTVB_TILE_WARPS / g.FC;
    if (strips < 1) strips = 1;
    const int need = (Ni + 2 + R - 1) / R;          // strips that cover the whole image height plus the halo rows
    if (strips > need) strips = need;
    for (;; --strips) {
        g.strips = strips;
        g.RPF = strips * R;
        g.TI = g.RPF - 2;
        g.rowsX = g.RPF + 2;
        g.slotX = g.rowsX * g.pitchX;
        g.slotW = g.RPF * g.WJ;
        g.xslot = (int)((((size_t)g.FC * g.slotX * sizeof(T) + 127) & ~size_t(127)) / sizeof(T));
        if (tile_smem_bytes<T>(g, mask) <= TILE_SMEM_LIMIT) break;
        if (strips == 1) return false;
    }
    g.nthreads = 32 * g.FC * g.strips;
    g.nti = (Ni + g.TI - 1) / g.TI;
    g.ntj = (Nj + g.TJ - 1) / g.TJ;
    g.nfg = t_on ? 1 : M;
    // z chunks: enough CTAs to fill the machine in whole waves, against the 2-3 warm-up planes every chunk pays
    const long long base = (long long)g.nti * g.ntj * g.nfg;
    size_t occ = TILE_SMEM_LIMIT / tile_smem_bytes<T>(g, mask);
    if (occ * g.nthreads > 2048) occ = 2048 / g.nthreads;
    if (occ < 1) occ = 1;
    const long long resident = (long long)sm_count * (long long)occ;
    long long best = -1;
    int best_n = 1;
    for (int n = 1; n <= Nz; ++n) {
        const int L = (Nz + n - 1) / n;
        if (n > 1 && L < 4) break;
        const long long ctas = base * ((Nz + L - 1) / L);
        const long long waves = (ctas + resident - 1) / resident;
        const long long cost = waves * (L + 3);
        if (best < 0 || cost < best) { best = cost; best_n = n; }
    }
    g.Lz = (Nz + best_n - 1) / best_n;
    g.nzc = (Nz + g.Lz - 1) / g.Lz;
    g.nblocks = base * g.nzc;
    return true;
}

// Block index -> tile context (j tiles fastest: neighbouring tiles run at the same time and share their halos through L2).
template <typename T, int VEC>
PYTVB_HD TileCtx<T> tile_ctx(const TileGeom& g, long long b, const Params<T>& P, unsigned char* smem, bool mask) {
    TileCtx<T> c;
    const int Nz = P.Nz;
    smem += (128 - (int)(reinterpret_cast<uintptr_t>(smem) & 127)) & 127;     // pointer + offset: the compiler keeps the shared address space
    const int tj = (int)(b % g.ntj); b /= g.ntj;
    const int ti = (int)(b % g.nti); b /= g.nti;
    const int fg = (int)(b % g.nfg); b /= g.nfg;
    const int zc = (int)b;
    c.i0 = ti * g.TI;
    c.j0 = tj * g.TJ;
    c.t0 = fg;
    c.zc0 = zc * g.Lz;
    c.zc1 = c.zc0 + g.Lz < Nz ? c.zc0 + g.Lz : Nz;
    c.fix = c.i0 - 2 < 0 || c.i0 - 2 + g.rowsX > P.Ni || c.j0 - 2 * VEC < 0 || c.j0 - 2 * VEC + g.pitchX > P.Nj;
    c.Xs = reinterpret_cast<T*>(smem);
    c.Ws = c.Xs + (size_t)3 * g.xslot;
    size_t off = (((size_t)3 * g.xslot + (size_t)g.FC * g.slotW) * sizeof(T) + 15) & ~size_t(15);
    c.rowg = reinterpret_cast<long long*>(smem + off);
    c.rowd = reinterpret_cast<int*>(c.rowg + (size_t)g.FC * g.rowsX);
    c.Ms = mask ? reinterpret_cast<uint8_t*>(c.rowd + (size_t)g.FC * g.rowsX) : nullptr;
    return c;
}

#if defined(__CUDACC__)
// Staging of one plane of the x window: the vector path (VEC > 1) by TMA - thread 0 issues one box load per plane, everyone
// waits on the slot's mbarrier and, in CTAs at the image border, repairs the zero-filled cells - the scalar path (row lengths
// not divisible by the vector width, unaligned pointers: no tensor map possible) by per-thread cp.async with clamped indices.
template <typename T, int VEC>
struct TileStager {
    static constexpr bool TMA = VEC > 1;
    const TileCtx<T>& c; const TileGeom& g; const ImgView<T>& X; const Params<T>& P;
    const CUtensorMap* mapX; const CUtensorMap* mapLo; const CUtensorMap* mapHi;
    unsigned long long* bar;      // [3], one per slot
    unsigned parity;              // bit s: parity of the next completion of slot s
    int tid;
    __device__ __forceinline__ void init() {
        parity = 0;
        if (TMA) {
            if (tid == 0) {
                for (int s = 0; s < 3; ++s) mbar_init(bar + s, 1);
                mbar_init_fence();
            }
        }
    }
    __device__ __forceinline__ void issue(int q) {
        if constexpr (TMA) {
            if (tid == 0) {
                const int s = tile_slot(q), ql = tile_clamp_plane(P, q);
                const CUtensorMap* m = ql < 0 ? mapLo : (ql >= P.Nz ? mapHi : mapX);
                const int zi = ql < 0 ? ql + X.depth : (ql >= P.Nz ? ql - P.Nz : ql);
                fence_proxy_async();
                mbar_expect_tx(bar + s, (unsigned)((size_t)g.FC * g.slotX * sizeof(T)));
                tma_load_4d(c.Xs + (size_t)s * g.xslot, m, bar + s, c.j0 - 2 * VEC, c.i0 - 2, c.t0, zi);
            }
        } else {
            tile_stage_plane<T, VEC>(c, g, X, P, q, tid);
        }
    }
    // plane q has landed (and is repaired); a __syncthreads must follow before other threads' cells are read
    __device__ __forceinline__ void land(int q) {
        if constexpr (TMA) {
            const int s = tile_slot(q);
            mbar_wait(bar + s, (parity >> s) & 1u);
            parity ^= 1u << s;
            if (c.fix) tile_fixup_plane<T, VEC>(c, g, P, q, tid);
        } else {
            stage_wait_all();
        }
    }
};

template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS>
__global__ void __launch_bounds__(TILE_MAX_THREADS, PYTVB_TILE_MINB)
tv_tile_kernel(ImgView<T> X, ImgView<T> TS, T* __restrict__ G, T* __restrict__ norms, double* __restrict__ partial, Params<T> P, TileGeom g,
               const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapLo, const __grid_constant__ CUtensorMap mapHi) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long stage_bar[3];
    const int tid = threadIdx.x;
    const bool mask = TSMODE >= 1 && P.mask_static != nullptr;
    const TileCtx<T> c = tile_ctx<T, VEC>(g, blockIdx.x, P, tile_smem, mask);
    TileStager<T, VEC> sg{c, g, X, P, &mapX, &mapLo, &mapHi, stage_bar, 0u, tid};
    sg.init();
    TileThread<T, VEC, R> st;
    st.tv = 0.0;
    if (mask) tile_stage_mask<T, VEC>(c, g, P, tid);
    if (!TileStager<T, VEC>::TMA) tile_stage_tables<T>(c, g, P, tid);
    const TilePos tp = tile_pos<T, VEC, R>(c, g, P, tid);
    __syncthreads();
    if (Z_ON) {
        const int p0 = c.zc0 - 1, p1 = c.zc1;
        sg.issue(p0);
        sg.issue(p0 + 1);
        sg.land(p0);
        sg.land(p0 + 1);
        __syncthreads();
        tile_init_z<T, VEC, SCHEME, R>(st, c, g, X, P, p0, tp);
        for (int p = p0; p <= p1; ++p) {
            const bool more = p + 2 <= p1 + 1;
            if (more) sg.issue(p + 2);
            tile_phase_w<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS>(st, c, g, P, TS, G, norms, p, tp);
            __syncthreads();
            if (p >= c.zc0 && p < c.zc1) tile_phase_g<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE>(st, c, g, P, TS, G, p, tp);
            if (more) sg.land(p + 2);
            __syncthreads();
        }
    } else {
        sg.issue(c.zc0);
        sg.land(c.zc0);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int e = 0; e < VEC; ++e) st.a[r][e] = st.w[r][e] = st.e[r][e] = st.g[r][e] = T(0);
        for (int p = c.zc0; p < c.zc1; ++p) {
            const bool more = p + 1 < c.zc1;
            if (more) sg.issue(p + 1);
            tile_phase_w<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS>(st, c, g, P, TS, G, norms, p, tp);
            __syncthreads();
            tile_phase_g<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE>(st, c, g, P, TS, G, p, tp);
            if (more) sg.land(p + 1);
            __syncthreads();
        }
    }
    // TV partial of this CTA (threads beyond the geometry's count do not exist: blockDim == g.nthreads)
    __shared__ double warp_part[TILE_MAX_THREADS / 32];
    double v = st.tv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_part[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < (g.nthreads >> 5); ++w) s += warp_part[w];
        partial[blockIdx.x] = s;
    }
}
#endif

}  // namespace pytvb
