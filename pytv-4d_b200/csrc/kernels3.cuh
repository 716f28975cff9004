// Generation 3: the whole Chambolle-Pock iteration in ONE launch, pass B lagging pass A by a few z-planes.
//
// Generations 1-2 run pass A (dual) over the volume, then pass B (primal): y is written by A (Nd*4 B/voxel),
// and read again from DRAM by B (another Nd*4 B/voxel) because a z-plane group of y (128 MB for the C4 slab)
// is larger than the L2.  Both kernels sit at the HBM roofline, so the only way to go faster is to move
// fewer bytes: 4(3Nd+5) -> 4(2Nd+5) = 116 -> 84 B/voxel for Nd = 8 if B finds y in L2.
//
// Schedule.  Tiles are the strip tiles of generation 2.  For every band of 64 image rows the tile sequence is
//     A(b,0) ... A(b,L-1) | A(b,L) B(b,0) | A(b,L+1) B(b,1) | ... | B(b,Nz-L) ... B(b,Nz-1)
// where X(b,z) is the group of all tiles of band b, plane z (all t, all column blocks).  A B tile reads y at its
// own rows +-1, planes z-1..z+1, all t: it must wait for the A groups {b-1,b} x {z-1,z,z+1}.  The same set
// covers the write-after-read hazard on xbar / x (B overwrites what those A groups read).  To make the row
// dependency point backwards only, B tiles are shifted up by one row: band b's B tiles own rows
// [64b-1, 64(b+1)-1) (the last band also takes the final row).
//
// Mechanics.  A CTA takes a ticket (atomicAdd) and decodes it into a tile, so a CTA only ever waits for
// tickets smaller than its own, all of which have started: no deadlock, whatever order the hardware
// dispatches blockIdx in.  A tiles publish completion with  stores -> __threadfence -> barrier -> atomicAdd(done);
// B tiles spin (with back-off and a bail-out that raises an error flag instead of hanging) on the six counters
// and then read y with ld.global.cg, because the L1 of their SM may still hold pre-update lines of y.
// Only the distance between A and B in the sequence (the lag L, in z-plane groups of ~8 MB) decides whether y
// is still in L2 when B reads it.
//
// Status (measured on B200, profiles/r01m_gen3_fused_ncu_full.txt, C4 slab): with L = 3 the kernel moves 51.3 GB
// instead of 62.7 GB (pass B finds two thirds of y in L2) but runs at 5.2 TB/s instead of 6.5 TB/s, so an
// iteration takes 9.69 ms against 9.60 ms for the two separate passes: a tie.  L = 1-2 wait too long for their
// producers, L >= 4 loses the L2 hits; createpolicy evict_last / evict_first hints on y and finer tiles (R = 2)
// did not help, and prefetching pass B's x / x0 before its wait made it slower (10.1 ms).  The kernel is therefore an option (PYTVB_FUSED=1 / CPSolver(fused=True)), not the default.
#pragma once
#include "kernels2.cuh"

namespace pytvb {

struct FusedSched {
    int Nz, M, ncb, RB, nbands, TW, tw_shift, TR, W, nstrips;
    int L;                       // lag in z-plane groups, 1 <= L <= Nz
    unsigned G;                  // tiles per group = ncb * RB * M
    unsigned per_band;           // tickets per band = 2 * Nz * G
    long long total;             // nbands * per_band
    FastDiv d_per_band, d_G, d_2G, d_ncb, d_RB;
};

struct FusedTile {
    int phase;                   // 0 = pass A (dual), 1 = pass B (primal)
    int band, z, cb, rbi, t;
};

inline FusedSched make_fused_sched(const Tiling& tl, int Nz, int nstrips, int lag) {
    FusedSched s;
    s.Nz = Nz; s.M = tl.M; s.ncb = tl.ncb; s.RB = tl.RB; s.nbands = tl.nbands; s.TW = tl.TW; s.tw_shift = tl.tw_shift;
    s.TR = CTA_THREADS / tl.TW; s.W = tl.W; s.nstrips = nstrips;
    s.L = lag < 1 ? 1 : (lag > Nz ? Nz : lag);
    s.G = (unsigned)(tl.ncb * tl.RB * tl.M);
    s.per_band = 2u * (unsigned)Nz * s.G;
    s.total = (long long)s.nbands * s.per_band;
    s.d_per_band = make_fastdiv(s.per_band);
    s.d_G = make_fastdiv(s.G);
    s.d_2G = make_fastdiv(2u * s.G);
    s.d_ncb = make_fastdiv((unsigned)s.ncb);
    s.d_RB = make_fastdiv((unsigned)s.RB);
    return s;
}

PYTVB_HD FusedTile fused_decode(unsigned ticket, const FusedSched& s) {
    FusedTile f;
    unsigned band, r, q, idx;
    s.d_per_band.divmod(ticket, band, r);
    f.band = (int)band;
    const unsigned headA = (unsigned)s.L * s.G;                          // A(b,0..L-1)
    const unsigned mid = (unsigned)(s.Nz - s.L) * 2u * s.G;              // pairs A(b,z) B(b,z-L)
    if (r < headA) {
        s.d_G.divmod(r, q, idx);
        f.phase = 0; f.z = (int)q;
    } else if (r < headA + mid) {
        unsigned pair, within;
        s.d_2G.divmod(r - headA, pair, within);
        if (within < s.G) { f.phase = 0; f.z = s.L + (int)pair; idx = within; }
        else { f.phase = 1; f.z = (int)pair; idx = within - s.G; }
    } else {
        s.d_G.divmod(r - headA - mid, q, idx);
        f.phase = 1; f.z = s.Nz - s.L + (int)q;
    }
    unsigned cb, rbi, t;
    s.d_ncb.divmod(idx, idx, cb);
    s.d_RB.divmod(idx, t, rbi);
    f.cb = (int)cb; f.rbi = (int)rbi; f.t = (int)t;
    return f;
}

struct FusedCtl {
    unsigned* ticket;       // [1]
    unsigned* done;         // [nbands * Nz] completed A tiles per (band, z) group
    unsigned* error;        // [1] set when a wait gave up
};

constexpr unsigned FUSED_SPIN_LIMIT = 1u << 22;   // x ~200 ns back-off: ~1 s before bailing out

__device__ __forceinline__ void fused_wait(const FusedCtl& c, const FusedSched& s, int band, int z) {
    if (band < 0 || z < 0 || z >= s.Nz) return;
    const volatile unsigned* flag = c.done + (long long)band * s.Nz + z;
    unsigned spins = 0;
    while (*flag < s.G) {
        if (*(const volatile unsigned*)c.error) break;            // somebody already gave up: drain quickly
        __nanosleep(200);
        if (++spins > FUSED_SPIN_LIMIT) { atomicExch(c.error, 1u); break; }
    }
}

// Poison the energy outputs when a wait timed out (the iteration's results are then not trustworthy).
static __global__ void fused_check_kernel(const unsigned* error, double* d_l21, double* d_fid) {
    if (*error) {
        if (d_l21) *d_l21 = NAN;
        if (d_fid) *d_fid = NAN;
    }
}

#ifndef PYTVB_FUSED_MINB
#define PYTVB_FUSED_MINB PYTVB_DUAL_MINB
#endif
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int VARIANT, int R, bool TS = false>
__global__ void __launch_bounds__(CTA_THREADS, PYTVB_FUSED_MINB)
cp_fused_kernel(ImgView<T> Xin, FieldView<T> Y, T* __restrict__ y, T* __restrict__ x, T* __restrict__ aux, const T* __restrict__ x0,
                double* __restrict__ partialA, double* __restrict__ partialB, Params<T> P, T sig, T lam, T tau, T c1, T c2, FusedSched s, FusedCtl ctl) {
    __shared__ unsigned s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ctl.ticket, 1u);
    __syncthreads();
    const unsigned ticket = s_ticket;
    const FusedTile f = fused_decode(ticket, s);
    const int tq = threadIdx.x & (s.TW - 1), tr = threadIdx.x >> s.tw_shift;
    const int strip = (f.band * s.RB + f.rbi) * s.TR + tr;
    const int qi = f.cb * s.TW + tq;
    const int j0 = qi * VEC;
    const bool active = strip < s.nstrips && qi < s.W;
    // one reduction slot per tile and phase: slot = position of the tile inside its phase
    const long long slot = ((long long)f.band * s.Nz + f.z) * s.G + ((f.t * s.RB + f.rbi) * s.ncb + f.cb);
    if (f.phase == 0) {
        T l21 = T(0);
        if (active) {
            const DualPlane<T> pl = make_dual_plane<T, SCHEME>(Xin, y, P, f.z, f.t);
            const int i0 = strip * R;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r;
                if (i < P.Ni) {
                    const int o = i * P.Nj + j0;
                    l21 += strip_quad_cp_dual<T, VEC, SCHEME, Z_ON, T_ON, T, TS>(pl, P, i, j0, o, i > 0 ? o - P.Nj : o, i < P.Ni - 1 ? o + P.Nj : o, sig, lam);
                }
            }
        }
        __threadfence();                       // this thread's y stores are visible device-wide ...
        const double bs = block_sum((double)l21 * (double)P.inv_div);   // (contains the CTA barrier)
        if (threadIdx.x == 0) {
            if (partialA) partialA[slot] = bs;
            __threadfence();
            atomicAdd(ctl.done + (long long)f.band * s.Nz + f.z, 1u);   // ... before the group counter moves
        }
    } else {
        if (threadIdx.x == 0) {
            for (int db = -1; db <= 0; ++db)
                for (int dz = -1; dz <= 1; ++dz) fused_wait(ctl, s, f.band + db, f.z + dz);
            __threadfence();
        }
        __syncthreads();
        T fid = T(0);
        if (active) {
            const PrimalPlane<T> pl = make_primal_plane<T, SCHEME, Z_ON, T_ON>(Y, P, f.z, f.t);
            const int i0 = strip * R - 1;                              // B tiles sit one row higher than A tiles
            const int nrows = (strip == s.nstrips - 1) ? R + 1 : R;    // the last strip also takes the final row
#pragma unroll
            for (int r = 0; r < R + 1; ++r) {
                const int i = i0 + r;
                if (r < nrows && i >= 0 && i < P.Ni) {
                    const int o = i * P.Nj + j0;
                    fid += strip_quad_cp_primal<T, VEC, SCHEME, Z_ON, T_ON, VARIANT, true, T, TS>(x, aux, x0, pl, P, i, j0, o, i > 0 ? o - P.Nj : o,
                                                                                           i < P.Ni - 1 ? o + P.Nj : o, tau, c1, c2);
                }
            }
        }
        if (partialB) {
            const double bs = block_sum((double)fid);
            if (threadIdx.x == 0) partialB[slot] = bs;
        }
    }
}

}  // namespace pytvb
