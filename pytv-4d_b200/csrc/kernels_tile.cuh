// Single-sweep tv_<scheme> kernels: z-marching CTA tiles.  Two forms share the geometry, the staging and the launch:
//   form 1 (tile_core.cuh):  per plane a w-phase and a G-phase, two barriers, 3 x slots + 1 w window;
//   form 2 (tile2_core.cuh): one phase per plane (norms of plane p, sub-gradient of plane p-1), one barrier, 4 x slots + 2 w windows.
#pragma once
#include <cuda.h>       // CUtensorMap (type only: the encoder is looked up at run time, tmap.cuh)
#include "kernels.cuh"
#include "tile_core.cuh"
#include "tile2_core.cuh"

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
constexpr size_t TILE_SMEM_LIMIT = 227 * 1024 - 1024;    // opt-in maximum per CTA on sm_100 minus the static scratch (reductions, mbarriers)

template <typename T>
inline size_t tile_smem_bytes(const TileGeom& g, bool mask) {
    // 128: the windows start on a 128-byte boundary (TMA destination)
    size_t b = 128 + ((size_t)2 * g.padf + (size_t)g.nslots * g.xslot + (size_t)g.nwbuf * g.wbuf) * sizeof(T);
    b = (b + 15) & ~size_t(15);
    b += (size_t)g.FC * g.rowsX * (sizeof(long long) + sizeof(int));       // staging tables
    b = (b + 15) & ~size_t(15);
    if (mask) b += (size_t)g.RPF * g.WJ * sizeof(T);                        // static-mask factors of the work region
    return (b + 15) & ~size_t(15);
}

// The tile of a geometry with `strips` strips per frame (everything that follows from the strip count).
template <typename T, int R>
inline void tile_set_strips(TileGeom& g, int strips) {
    g.strips = strips;
    g.RPF = strips * R;
    g.TI = g.RPF - 2;
    g.rowsX = g.RPF + 2;
    g.slotX = g.rowsX * g.pitchX;
    g.slotW = g.RPF * g.WJ;
    g.xslot = (int)((((size_t)g.FC * g.slotX * sizeof(T) + 127) & ~size_t(127)) / sizeof(T));
    g.wbuf = g.FC * g.slotW + g.padb;      // form 2: one unused row behind EACH w window (tile2_core.cuh: the halo rows' reads one row outside)
    g.nthreads = 32 * g.FC * g.strips;
}

// Geometry for a problem in the given form (1 | 2); returns false when the tile kernel cannot take it (too many coupled frames).
template <typename T, int VEC, int R>
inline bool make_tile_geom(TileGeom& g, int Nz, int M, int Ni, int Nj, bool t_on, bool mask, int form = 1, int sm_count = 148) {
    g.FC = t_on ? M : 1;
    g.WJ = 32 * VEC;
    g.TJ = g.WJ - 2 * VEC;
    g.pitchX = g.WJ + 2 * VEC;
    g.nslots = form == 2 ? 4 : 3;
    g.nwbuf = form == 2 ? 2 : 1;
    g.padf = form == 2 ? (int)((((size_t)2 * g.pitchX * sizeof(T) + 127) & ~size_t(127)) / sizeof(T)) : 0;
    g.padb = form == 2 ? g.WJ : 0;
    if (g.FC * 32 > TILE_MAX_THREADS) return false;
    int strips = PYTVB_TILE_WARPS / g.FC;
    if (strips < 1) strips = 1;
    const int need = (Ni + 2 + R - 1) / R;          // strips that cover the whole image height plus the halo rows
    if (strips > need) strips = need;
    for (;; --strips) {
        tile_set_strips<T, R>(g, strips);
        if (tile_smem_bytes<T>(g, mask) <= TILE_SMEM_LIMIT) break;
        if (strips == 1) return false;
    }
    g.nti = (Ni + g.TI - 1) / g.TI;
    g.ntj = (Nj + g.TJ - 1) / g.TJ;
    g.nfg = t_on ? 1 : M;
    // z chunks: enough CTAs to fill the machine in whole waves, against the 2-3 warm-up planes every chunk pays.  A chunk may be
    // as short as one plane: a volume with fewer tiles than SMs (the README volume: 8 tiles x 20 planes) is bound by the number of
    // steps a CTA runs one after the other, not by the redundant warm-up work.
    const long long base = (long long)g.nti * g.ntj * g.nfg;
    size_t occ = TILE_SMEM_LIMIT / tile_smem_bytes<T>(g, mask);
    if (occ * g.nthreads > 2048) occ = 2048 / g.nthreads;
    if (occ < 1) occ = 1;
    const long long resident = (long long)sm_count * (long long)occ;
    long long best = -1;
    int best_n = 1;
    for (int n = 1; n <= Nz; ++n) {
        const int L = (Nz + n - 1) / n;
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

// Which form runs a problem: the one-phase form when it keeps the tile of the two-phase form (its larger shared-memory
// footprint costs tile rows - redundant norm evaluations - beyond four coupled frames); 0 when neither can take it.
// force: 1 | 2 = that form if it can take the problem at all.
template <typename T, int VEC, int R>
inline int pick_tile_form(TileGeom& g, int Nz, int M, int Ni, int Nj, bool t_on, bool mask, int force = 0) {
    TileGeom g1, g2;
    const bool ok1 = make_tile_geom<T, VEC, R>(g1, Nz, M, Ni, Nj, t_on, mask, 1);
    const bool ok2 = make_tile_geom<T, VEC, R>(g2, Nz, M, Ni, Nj, t_on, mask, 2);
    if (force == 1 && ok1) { g = g1; return 1; }
    if (force == 2 && ok2) { g = g2; return 2; }
    if (ok2 && (!ok1 || g2.strips >= g1.strips)) { g = g2; return 2; }
    if (ok1) { g = g1; return 1; }
    return 0;
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
    c.Xs = reinterpret_cast<T*>(smem) + g.padf;
    c.Ws = c.Xs + (size_t)g.nslots * g.xslot + g.padf;      // padf elements unused on both sides of the x windows
    size_t off = (((size_t)2 * g.padf + (size_t)g.nslots * g.xslot + (size_t)g.nwbuf * g.wbuf) * sizeof(T) + 15) & ~size_t(15);
    c.rowg = reinterpret_cast<long long*>(smem + off);
    c.rowd = reinterpret_cast<int*>(c.rowg + (size_t)g.FC * g.rowsX);
    off = (off + (size_t)g.FC * g.rowsX * (sizeof(long long) + sizeof(int)) + 15) & ~size_t(15);
    c.Ms = mask ? reinterpret_cast<T*>(smem + off) : nullptr;
    return c;
}

#if defined(__CUDACC__)
// Staging of one plane of the x window: the vector path (VEC > 1) by TMA - thread 0 issues one box load per plane, everyone
// waits on the slot's mbarrier and, in CTAs at the image border, repairs the zero-filled cells - the scalar path (row lengths
// not divisible by the vector width, unaligned pointers: no tensor map possible) by per-thread cp.async with clamped indices.
// `ql` = slab-local plane (already clamped), `s` = slot of the ring.
template <typename T, int VEC>
struct TileStager {
    static constexpr bool TMA = VEC > 1;
    const TileCtx<T>& c; const TileGeom& g; const ImgView<T>& X; const Params<T>& P;
    const CUtensorMap* mapX; const CUtensorMap* mapLo; const CUtensorMap* mapHi;
    unsigned long long* bar;      // one per slot
    unsigned parity;              // bit s: parity of the next completion of slot s
    int tid;
    __device__ __forceinline__ void init() {
        parity = 0;
        if (TMA) {
            if (tid == 0) {
                for (int s = 0; s < g.nslots; ++s) mbar_init(bar + s, 1);
                mbar_init_fence();
            }
        }
    }
    __device__ __forceinline__ void issue(int ql, int s) {
        if constexpr (TMA) {
            if (tid == 0) {
                const CUtensorMap* m = ql < 0 ? mapLo : (ql >= P.Nz ? mapHi : mapX);
                const int zi = ql < 0 ? ql + X.depth : (ql >= P.Nz ? ql - P.Nz : ql);
                fence_proxy_async();
                mbar_expect_tx(bar + s, (unsigned)((size_t)g.FC * g.slotX * sizeof(T)));
                tma_load_4d(c.Xs + (size_t)s * g.xslot, m, bar + s, c.j0 - 2 * VEC, c.i0 - 2, c.t0, zi);
            }
        } else {
            tile_stage_plane<T, VEC>(c, g, X, P, ql, s, tid);
        }
    }
    // L2 prefetch of plane ql (vector path; the scalar path has no tensor map)
    __device__ __forceinline__ void prefetch(int ql) {
        if constexpr (TMA) {
            if (tid == 0) {
                const CUtensorMap* m = ql < 0 ? mapLo : (ql >= P.Nz ? mapHi : mapX);
                const int zi = ql < 0 ? ql + X.depth : (ql >= P.Nz ? ql - P.Nz : ql);
                tma_prefetch_4d(m, c.j0 - 2 * VEC, c.i0 - 2, c.t0, zi);
            }
        }
    }
    // the plane of slot s has landed (and is repaired); a __syncthreads must follow before other threads' cells are read
    __device__ __forceinline__ void land(int s) {
        if constexpr (TMA) {
            mbar_wait(bar + s, (parity >> s) & 1u);
            parity ^= 1u << s;
            if (c.fix) tile_fixup_plane<T, VEC>(c, g, P, s, tid);
        } else {
            stage_wait_all();
        }
    }
};

// CTA-wide sum of the threads' TV partials -> partial[blockIdx.x]; the CTA that finishes last adds the partials up (finish_partials).
__device__ __forceinline__ void tile_store_partial(double v, double* __restrict__ partial, unsigned* counter, double* __restrict__ d_out, int nthreads) {
    __shared__ double warp_part[TILE_MAX_THREADS / 32];
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_part[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (tid == 0)
        for (int w = 0; w < (nthreads >> 5); ++w) s += warp_part[w];
    finish_partials(s, partial, counter, d_out);
}

// ---- form 1: two phases per plane
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS>
__global__ void __launch_bounds__(TILE_MAX_THREADS, PYTVB_TILE_MINB)
tv_tile_kernel(ImgView<T> X, ImgView<T> TS, T* __restrict__ G, T* __restrict__ norms, double* __restrict__ partial, unsigned* counter,
               double* __restrict__ d_tv, Params<T> P, TileGeom g,
               const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapLo, const __grid_constant__ CUtensorMap mapHi) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long stage_bar[4];
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
        sg.issue(tile_clamp_plane(P, p0), tile_slot(p0));
        sg.issue(tile_clamp_plane(P, p0 + 1), tile_slot(p0 + 1));
        sg.land(tile_slot(p0));
        sg.land(tile_slot(p0 + 1));
        __syncthreads();
        tile_init_z<T, VEC, SCHEME, R>(st, c, g, X, P, p0, tp);
        for (int p = p0; p <= p1; ++p) {
            const bool more = p + 2 <= p1 + 1;
            if (more) sg.issue(tile_clamp_plane(P, p + 2), tile_slot(p + 2));
            if (p + 3 <= p1 + 1) sg.prefetch(tile_clamp_plane(P, p + 3));
            tile_phase_w<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS>(st, c, g, P, TS, G, norms, p, tp);
            __syncthreads();
            if (p >= c.zc0 && p < c.zc1) tile_phase_g<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE>(st, c, g, P, TS, G, p, tp);
            if (more) sg.land(tile_slot(p + 2));
            __syncthreads();
        }
    } else {
        sg.issue(c.zc0, tile_slot(c.zc0));
        sg.land(tile_slot(c.zc0));
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int e = 0; e < VEC; ++e) st.a[r][e] = st.w[r][e] = st.e[r][e] = st.g[r][e] = T(0);
        for (int p = c.zc0; p < c.zc1; ++p) {
            const bool more = p + 1 < c.zc1;
            if (more) sg.issue(p + 1, tile_slot(p + 1));
            tile_phase_w<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS>(st, c, g, P, TS, G, norms, p, tp);
            __syncthreads();
            tile_phase_g<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE>(st, c, g, P, TS, G, p, tp);
            if (more) sg.land(tile_slot(p + 1));
            __syncthreads();
        }
    }
    tile_store_partial(st.tv, partial, counter, d_tv, g.nthreads);
}

// ---- form 2: one phase per plane (tile2_core.cuh)
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS>
__global__ void __launch_bounds__(TILE_MAX_THREADS, PYTVB_TILE_MINB)
tv_tile2_kernel(ImgView<T> X, ImgView<T> TS, T* __restrict__ G, T* __restrict__ norms, double* __restrict__ partial, unsigned* counter,
                double* __restrict__ d_tv, Params<T> P, TileGeom g,
                const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapLo, const __grid_constant__ CUtensorMap mapHi) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long stage_bar[4];
    const int tid = threadIdx.x;
    const bool mask = TSMODE >= 1 && P.mask_static != nullptr;
    const TileCtx<T> c = tile_ctx<T, VEC>(g, blockIdx.x, P, tile_smem, mask);
    TileStager<T, VEC> sg{c, g, X, P, &mapX, &mapLo, &mapHi, stage_bar, 0u, tid};
    sg.init();
    Tile2Thread<T, VEC, R> st;
    st.tv = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        st.xl[r] = st.xr[r] = T(0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) st.f[r][e] = st.f1[r][e] = T(0);
    }
    if (mask) tile_stage_mask<T, VEC>(c, g, P, tid);
    if (!TileStager<T, VEC>::TMA) tile_stage_tables<T>(c, g, P, tid);
    const TilePos tp = tile_pos<T, VEC, R>(c, g, P, tid);
    __syncthreads();
    // steps p0 .. p1: step p needs the planes p-1, p and (z axis on) p+1; plane p+2 lands during step p
    const int p0 = Z_ON ? c.zc0 - 1 : c.zc0, p1 = c.zc1;
    for (int q = p0 - 1; q <= p0 + 1; ++q) sg.issue(tile2_plane<T, Z_ON>(P, q), tile2_slot(q));
    if (p0 + 2 <= p1 + 1) sg.prefetch(tile2_plane<T, Z_ON>(P, p0 + 2));
    for (int q = p0 - 1; q <= p0 + 1; ++q) sg.land(tile2_slot(q));
    __syncthreads();
    for (int p = p0; p <= p1; ++p) {
        const bool more = p + 2 <= p1 + 1;
        if (more) sg.issue(tile2_plane<T, Z_ON>(P, p + 2), tile2_slot(p + 2));
        if (p + 3 <= p1 + 1) sg.prefetch(tile2_plane<T, Z_ON>(P, p + 3));      // one plane further ahead, into L2 only
        tile2_step<T, VEC, SCHEME, Z_ON, T_ON, R, TSMODE, NORMS>(st, c, g, P, TS, G, norms, p, tp);
        if (more) sg.land(tile2_slot(p + 2));
        __syncthreads();
    }
    tile_store_partial(st.tv, partial, counter, d_tv, g.nthreads);
}
#endif

}  // namespace pytvb
