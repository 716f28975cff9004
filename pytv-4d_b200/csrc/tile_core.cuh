// Single-sweep TV value + sub-gradient (tv_<scheme>, reference pytv/tv_GPU.py:47,142,217,290): per-thread code of the
// z-marching tile kernel (kernels_tile.cuh).
//
// Spec (tv_GPU.py:84-126 hybrid, :176-188, :239-251, :302-328): D = D_s(x); n = |D|_2 per voxel; tv = sum n;
// G = D_T_s^unit(D / n) with 0/0 := 0.  With w = 1/n (0 where n = 0) the sub-gradient is a sum of EDGE
// TERMS: along one axis, term(a -> b) = (x_b - x_a) * S(w_a, w_b) for neighbouring voxels a, b (S = w_a upwind, w_b
// downwind, w_a + w_b hybrid) and G(v) = sum_axes [term(v-e -> v) - term(v -> v+e)] * weight; the centred scheme has
// C(m) = (x_{m+1} - x_{m-1}) * w_m and G(v) = C(v-e) - C(v+e).  So G(v) needs w at distance 1 and x at distance 2.
//
// One launch, x read once, G written once (8 B/voxel): a CTA owns an in-plane tile (TI x TJ output voxels, all M time
// frames when the time axis is on) and marches along z.  This file holds what both forms of the kernel share - geometry, the
// staging of the x windows (TMA on the vector path, per-thread cp.async on the scalar path), the thread coordinates, the norm -
// and the per-thread code of FORM 1, two phases per z-plane p ("step"):
//   * x(p+2) arrives in shared memory one step ahead.  The window holds the tile plus 2 halo rows / columns; at the volume
//     boundary the cell just outside takes the boundary value, so every out-of-range difference is exactly 0 - the reference's
//     rule, tv_operators_CPU.py:118 - without predicates (TMA's zero fill alone would be wrong: tile_fixup_plane);
//   * w-phase: every thread computes n(p), w(p) for its R rows x one quad column (tile + 1 halo ring), publishes w(p) to
//     shared memory, forms the z edge term between planes p-1 and p and with it completes and stores G(p-1);
//   * barrier; G-phase: the in-plane and time edge terms of plane p from the x / w windows -> kept in registers until the
//     next step supplies the z term; barrier.
// z neighbours live in registers (per owned voxel: the raw z difference, the running G and - within a step - w and the z term),
// in-plane and time neighbours in shared memory (3 x-planes x M frames, 1 w-plane x M frames).  FORM 2 (tile2_core.cuh) runs one
// phase and one barrier per plane with a 4-slot x ring and two w windows; kernels_tile.cuh picks the form and holds the kernels.
// Everything here is __host__ __device__ and index-pure: tests/emul runs the same code thread by thread on the CPU.
#pragma once
#include "strip_core.cuh"

namespace pytvb {

// ---- element-wise arithmetic on the VEC values of a quad.  The kernel is instruction-issue bound (profiles/r02d_*: 63 %
// issue utilisation, FP32 40 % of the executed instructions), and sm_100 has packed FP32 instructions - FADD2 / FMUL2 / FFMA2
// (PTX add/mul/fma.rn.f32x2) - that do two IEEE operations in one issue slot (measured: the same FMA rate as the scalar form,
// half the issue slots, profiles/r02a_ffma2.txt).  On the device, for float quads, every operation below is two packed
// instructions instead of four scalar ones; the values are the same as those of the scalar code with the same fusion.  On the
// host (tests/emul) and for double the plain loops run.
template <typename T, int VEC>
struct VOp {
    static constexpr bool PACKED = false;
    static PYTVB_HD void sub(T* d, const T* a, const T* b) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] - b[e];
    }
    static PYTVB_HD void add(T* d, const T* a, const T* b) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] + b[e];
    }
    static PYTVB_HD void mul(T* d, const T* a, const T* b) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] * b[e];
    }
    static PYTVB_HD void fma(T* d, const T* a, const T* b, const T* c) {     // d = a * b + c (fused where the target fuses)
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] * b[e] + c[e];
    }
    static PYTVB_HD void muls(T* d, const T* a, T k) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] * k;
    }
    static PYTVB_HD void fmas(T* d, const T* a, T k, const T* c) {           // d = a * k + c
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = a[e] * k + c[e];
    }
};
#if defined(__CUDA_ARCH__)
#define PYTVB_F2_BIN(NAME, INS)                                                                                        \
    static __device__ __forceinline__ void NAME(float* d, const float* a, const float* b) {                           \
        _Pragma("unroll") for (int e = 0; e < 4; e += 2)                                                               \
            asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n " INS " rd, ra, rb;\n mov.b64 {%0, %1}, rd; }" \
                : "=f"(d[e]), "=f"(d[e + 1]) : "f"(a[e]), "f"(a[e + 1]), "f"(b[e]), "f"(b[e + 1]));                    \
    }
template <>
struct VOp<float, 4> {
    static constexpr bool PACKED = true;
    PYTVB_F2_BIN(sub, "sub.rn.f32x2")
    PYTVB_F2_BIN(add, "add.rn.f32x2")
    PYTVB_F2_BIN(mul, "mul.rn.f32x2")
    static __device__ __forceinline__ void fma(float* d, const float* a, const float* b, const float* c) {
#pragma unroll
        for (int e = 0; e < 4; e += 2)
            asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }"
                : "=f"(d[e]), "=f"(d[e + 1]) : "f"(a[e]), "f"(a[e + 1]), "f"(b[e]), "f"(b[e + 1]), "f"(c[e]), "f"(c[e + 1]));
    }
    static __device__ __forceinline__ void muls(float* d, const float* a, float k) {
#pragma unroll
        for (int e = 0; e < 4; e += 2)
            asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %4};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }"
                : "=f"(d[e]), "=f"(d[e + 1]) : "f"(a[e]), "f"(a[e + 1]), "f"(k));
    }
    static __device__ __forceinline__ void fmas(float* d, const float* a, float k, const float* c) {
#pragma unroll
        for (int e = 0; e < 4; e += 2)
            asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %4};\n mov.b64 rc, {%5, %6};\n fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }"
                : "=f"(d[e]), "=f"(d[e + 1]) : "f"(a[e]), "f"(a[e + 1]), "f"(k), "f"(c[e]), "f"(c[e + 1]));
    }
};
#undef PYTVB_F2_BIN
#endif

// Geometry of one launch (host-computed, warp-uniform).
struct TileGeom {
    int strips;        // R-row strips per frame handled by one CTA
    int FC;            // frames per CTA: M when the time axis is on, else 1
    int RPF;           // work rows per frame = strips * R  (tile rows -1 .. TI)
    int TI, TJ;        // output tile
    int WJ;            // work columns = 32 * VEC (tile columns -VEC .. TJ+VEC-1)
    int pitchX, rowsX; // x window: rows -2 .. TI+1, columns -VEC-1 .. TJ+VEC  (pitch = WJ + 2 VEC)
    int slotX, slotW;  // elements of one (plane, frame) x window / w window
    int xslot;         // elements from one plane slot of the x window to the next (FC * slotX rounded up to 128 bytes: TMA destination alignment)
    int nslots, nwbuf; // plane slots of the x window ring / w windows: 3 / 1 two-phase form, 4 / 2 one-phase form (tile2_core.cuh)
    int wbuf;          // elements from one w window to the next (FC * slotW + padb)
    int padf, padb;    // elements kept free in front of and behind the x windows / behind each w window (the one-phase form's halo rows read one
                       // or two rows outside their window and drop the result; the rows they hit are never written during a step)
    int nti, ntj, nfg, nzc, Lz;   // tiles along i, j; frame groups; z chunks and their length
    int nthreads;
    long long nblocks;
};

// Compile-time part of the geometry (a function of the vector width only): with these as constants the window rows of a
// thread are immediate offsets from one base register instead of a multiply-add per access.
template <int VEC>
struct TileC {
    static constexpr int WJ = 32 * VEC;        // work columns
    static constexpr int TJ = 30 * VEC;        // output columns
    static constexpr int PX = 34 * VEC;        // pitch of the x window
};

template <typename T>
struct TileCtx {
    T* Xs;             // [nslots] x windows, xslot elements apart, each [FC][slotX]
    T* Ws;             // [nwbuf] w windows, wbuf elements apart, each [FC][slotW]
    T* Ms;             // [RPF][WJ] factor of the time component(s) from the static mask over the work region (sqrt(factor_reg_static) on
                       // static pixels, 1 elsewhere), or null.  As values, not mask bytes: one 128-bit load per row and part instead of a
                       // 32-bit load and four selects (mask_static cost 19-21 % of the kernel on the C5 slab, profiles/r02z_launches_c5.csv)
    long long* rowg;   // [FC * rowsX] staging table: element offset of window row r inside a z-plane group (frame and clamped row)
    int* rowd;         // [FC * rowsX] staging table: element offset of window row r inside a slot group (frame, row)
    int i0, j0, t0;    // global row / column of tile (0, 0); first frame
    int zc0, zc1;      // output planes [zc0, zc1), slab-local
    bool fix;          // the x window reaches outside the image: zero-filled (TMA) windows need tile_fixup_plane
};

// Per-thread state carried from one z step to the next.
template <typename T, int VEC, int R>
struct TileThread {
    T a[R][VEC];    // one-sided / hybrid: raw z difference x(p) - x(p-1);  centred: x(p-1)
    T w[R][VEC];    // w of the plane of the last w-phase
    T e[R][VEC];    // one-sided / hybrid: srz * term(p-1 -> p) of the last w-phase (the z term coming into plane p);  centred: srz * Cz(p-1)
    T g[R][VEC];    // running sub-gradient: in-plane + time part of the last G-phase plus the incoming z term
    double tv;      // running sum of the norms of this thread's output voxels (one double add per step: the float part spans R rows only,
                    // so the value does not depend on how the volume is cut into z chunks or slabs beyond 1e-16)
};

PYTVB_HD int tile_slot(int q) { return (q + 3) % 3; }   // q >= -3

// ---- staging: global -> shared, indices clamped to the volume.
template <int BYTES>
PYTVB_HD void stage_copy(void* dst, const void* src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
#else
    memcpy(dst, src, BYTES);
#endif
}
PYTVB_HD void stage_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

// The compiler otherwise rebuilds an output address from the kernel parameters inside every predicated store (27 instructions
// per row in profiles/r02m_*): an empty asm makes the pointer an opaque value that has to stay in its two registers.  (The store
// becomes a generic-address ST; the alternative - the opaque value on the offset, STG kept - costs two registers more and measured
// slower, profiles/r02r_tv_times.txt.)
template <typename T>
PYTVB_HD void keep_in_registers(T*& p) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(p));
#else
    (void)p;
#endif
}

PYTVB_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Slab-local index of plane q clamped to the GLOBAL volume (planes outside the slab but inside the volume come from the
// halo buffers of the view).
template <typename T>
PYTVB_HD int tile_clamp_plane(const Params<T>& P, int q) {
    long long zg = P.zg0 + q;
    if (zg < 0) zg = 0;
    if (zg > P.NzG - 1) zg = P.NzG - 1;
    return (int)(zg - P.zg0);
}

// Staging tables, once per CTA: everything about a window row that does not depend on the plane.
template <typename T>
PYTVB_HD void tile_stage_tables(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    for (int r = tid; r < g.FC * g.rowsX; r += g.nthreads) {
        const int fl = r / g.rowsX, xr = r - fl * g.rowsX;
        c.rowg[r] = (long long)(c.t0 + fl) * P.sT + (long long)clampi(c.i0 + xr - 2, 0, P.Ni - 1) * P.Nj;
        c.rowd[r] = fl * g.slotX + xr * g.pitchX;
    }
}

// Stage plane q of every frame of the CTA into its slot.  Work is split by warps over (frame, window row); a lane copies
// one quad of the row, lanes 0 / 1 also the scalar column left / right of the work columns.  Per row: one table look-up,
// one add and the copy (the first version recomputed frame / row / clamps per copy and spent 40 % of the kernel's
// instructions here, profiles/r02b_tv_tile_first_ncu_full.txt).
template <typename T, int VEC>
PYTVB_HD void tile_stage_plane(const TileCtx<T>& c, const TileGeom& g, const ImgView<T>& X, const Params<T>& P, int ql, int sl, int tid) {
    const int lane = tid & 31, wid = tid >> 5, nwarps = g.nthreads >> 5;
    const T* plane = X.row(P, ql, 0, 0);
    T* slot = c.Xs + (long long)sl * g.xslot;
    const int cj = -VEC + lane * VEC;                 // tile column of this lane's quad
    const int gj0 = c.j0 + cj;
    const bool vec_ok = VEC > 1 && gj0 >= 0 && gj0 + VEC <= P.Nj;
    const int nrows = g.FC * g.rowsX;
    if (vec_ok) {
        for (int rq = wid; rq < nrows; rq += nwarps) stage_copy<sizeof(T) * VEC>(slot + c.rowd[rq] + cj + 2 * VEC, plane + c.rowg[rq] + gj0);
    } else {
        for (int rq = wid; rq < nrows; rq += nwarps) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) stage_copy<sizeof(T)>(slot + c.rowd[rq] + cj + 2 * VEC + e, plane + c.rowg[rq] + clampi(gj0 + e, 0, P.Nj - 1));
        }
    }
    if (lane < 2) {
        const int gcol = lane == 0 ? clampi(c.j0 - VEC - 1, 0, P.Nj - 1) : clampi(c.j0 + g.TJ + VEC, 0, P.Nj - 1);
        const int dcol = lane == 0 ? VEC - 1 : g.TJ + 3 * VEC;
        for (int rq = wid; rq < nrows; rq += nwarps) stage_copy<sizeof(T)>(slot + c.rowd[rq] + dcol, plane + c.rowg[rq] + gcol);
    }
}

// ---- staging by TMA (vector path).  The per-thread copies above cost 37 % of the kernel's instructions and of its stall
// samples (table look-ups, 64-bit addresses, one LDGSTS per quad; profiles/r02f_tv_tile_sass_mix.txt).  With a tensor map of
// the (planes, M, Ni, Nj) image one thread issues ONE cp.async.bulk.tensor per plane for the whole (FC, rowsX, pitchX) box
// and the other threads spend nothing.  TMA fills cells outside the tensor with ZEROS, which is the wrong boundary rule
// here (core.cuh), so the CTAs whose window reaches outside the image (c.fix) repair the window once it has landed.  Only the
// FIRST row / column outside the image needs the clamped value: it makes the difference across the image border exactly 0.
// Everything further out feeds only the w of ring voxels outside the image, and those w meet nothing but that zero difference
// (edge terms) or a zero existence factor (centred scheme) - any finite value, so the zeros stay.  (The first version clamped
// every out-of-range cell: up to 16 quads per row in the last column of tiles, 6 % of the kernel on the C4 slab.)  The repair
// reads only in-image cells and writes only out-of-image cells, rows and columns disjoint (the corners stay zero): no ordering.
template <typename T, int VEC>
PYTVB_HD void tile_fixup_plane(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int sl, int tid) {
    T* slot = c.Xs + (long long)sl * g.xslot;
    const int iw = c.i0 - 2, jw = c.j0 - 2 * VEC;           // image row / column of window cell (0, 0)
    const int rT = -iw, rB = P.Ni - 1 - iw, cL = -jw, cR = P.Nj - 1 - jw;      // window row / column of the image's first / last
    const int nq = g.pitchX / VEC;
    // rows: the row just above the image <- row 0, the row just below <- the last row (in-image columns).  (Loops over the frames
    // with the thread index as the item: no integer divisions in a piece of code the border CTAs run every step.)
    for (int fl = 0; fl < g.FC; ++fl) {
        T* frame = slot + (long long)fl * g.slotX;
        for (int idx = tid; idx < 2 * nq; idx += g.nthreads) {
            const int side = idx >= nq ? 1 : 0, q = idx - side * nq;
            const int rs = side ? rB : rT, rd = side ? rB + 1 : rT - 1;
            if (rd < 0 || rd >= g.rowsX || q * VEC < cL || q * VEC + VEC - 1 > cR) continue;
            st_pack<T, VEC>(frame + rd * g.pitchX + q * VEC, ld_pack<T, VEC>(frame + rs * g.pitchX + q * VEC));
        }
        // columns: the cell left of the image <- column 0, the cell right of it <- the last column (in-image rows)
        for (int idx = tid; idx < 2 * g.rowsX; idx += g.nthreads) {
            const int side = idx & 1, xr = idx >> 1;
            const int cs = side ? cR : cL, cd = side ? cR + 1 : cL - 1;
            if (cd < 0 || cd >= g.pitchX || xr < rT || xr > rB) continue;
            T* row = frame + xr * g.pitchX;
            row[cd] = row[cs];
        }
    }
}

// What the TMA load leaves in the window, as plain code (host emulation of the vector path: tests/emul).
template <typename T, int VEC>
PYTVB_HD void tile_stage_plane_zfill(const TileCtx<T>& c, const TileGeom& g, const ImgView<T>& X, const Params<T>& P, int ql, int sl, int tid) {
    T* slot = c.Xs + (long long)sl * g.xslot;
    const int iw = c.i0 - 2, jw = c.j0 - 2 * VEC;
    const int ncell = g.FC * g.rowsX * g.pitchX;
    for (int k = tid; k < ncell; k += g.nthreads) {
        const int fl = k / g.slotX, rem = k - fl * g.slotX, xr = rem / g.pitchX, xc = rem - xr * g.pitchX;
        const int gi = iw + xr, gj = jw + xc;
        const bool in = gi >= 0 && gi < P.Ni && gj >= 0 && gj < P.Nj;
        slot[(long long)fl * g.slotX + rem] = in ? X.row(P, ql, c.t0 + fl, gi)[gj] : T(0);
    }
}

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Generic-proxy accesses to shared memory (the phases' reads, the repair's writes) before this point are ordered before
// the async-proxy writes of a TMA load issued after it.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// One (FC, rowsX, pitchX) box of plane `zi` of the tensor behind `map` -> dst; completion is counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_4d(void* dst, const void* map, unsigned long long* bar, int cj, int ci, int ct, int cz) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(cj), "r"(ci), "r"(ct), "r"(cz)
                 : "memory");
}
// The same box into L2 only (no shared memory, no completion): issued one plane further ahead than the load, so that the load finds
// its lines in L2 whatever the DRAM / far-partition latency of the moment is.
__device__ __forceinline__ void tma_prefetch_4d(const void* map, int cj, int ci, int ct, int cz) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(cj), "r"(ci), "r"(ct), "r"(cz) : "memory");
}
#endif

// Static-mask factors of the work region (clamped), once per CTA.
template <typename T, int VEC>
PYTVB_HD void tile_stage_mask(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    for (int k = tid; k < g.RPF * g.WJ; k += g.nthreads) {
        const int r = k / g.WJ, cc = k - r * g.WJ;
        const int gi = clampi(c.i0 + r - 1, 0, P.Ni - 1), gj = clampi(c.j0 + cc - VEC, 0, P.Nj - 1);
        c.Ms[k] = P.mask_static[(long long)gi * P.Nj + gj] ? P.sfac : T(1);
    }
}

// ---- thread coordinates
struct TilePos {
    int fl, rr0, cj;     // frame within the CTA, first work row (tile coordinates, >= -1), tile column of the quad
    int t;               // global frame
    int lane;            // index of the quad in its work row (0 and 31 are the halo lanes)
    bool col_out;        // the quad lies in the output tile and inside the image
    int xo, wo;          // element offset of the quad in work row 0 inside an x slot / the w window (frame included)
    int dxm, dxp;        // element distance to the same quad of the previous / next frame in an x slot (0 at the ends: clamped)
    int dwm, dwp;        // the same in the w window
    long long goff;      // element offset of (frame t, row i0 + rr0, column j0 + cj) inside a z-plane group (outputs)
};
// Computed once per thread (the division by the strip count and the 64-bit products stay out of the z loop).  On the device
// the warp index is read through a shuffle: the compiler then knows that everything derived from it (frame, rows, the row
// predicates) is warp-uniform and keeps it in uniform registers / uniform branches.
template <typename T, int VEC, int R>
PYTVB_HD TilePos tile_pos(const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, int tid) {
    TilePos p;
    const int lane = tid & 31;
#if defined(__CUDA_ARCH__)
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
#else
    const int wid = tid >> 5;
#endif
    p.fl = wid / g.strips;
    p.rr0 = -1 + (wid - p.fl * g.strips) * R;
    p.cj = -VEC + lane * VEC;
    p.lane = lane;
    p.t = c.t0 + p.fl;
    p.col_out = p.cj >= 0 && p.cj < g.TJ && c.j0 + p.cj < P.Nj;     // vector path: Nj % VEC == 0, so a quad is in or out as a whole
    p.xo = p.fl * g.slotX + (p.rr0 + 2) * TileC<VEC>::PX + p.cj + 2 * VEC;
    p.wo = p.fl * g.slotW + (p.rr0 + 1) * TileC<VEC>::WJ + p.cj + VEC;
    p.dxm = p.fl > 0 ? -g.slotX : 0;
    p.dxp = p.fl < g.FC - 1 ? g.slotX : 0;
    p.dwm = p.fl > 0 ? -g.slotW : 0;
    p.dwp = p.fl < g.FC - 1 ? g.slotW : 0;
    p.goff = (long long)p.t * P.sT + (long long)(c.i0 + p.rr0) * P.Nj + c.j0 + p.cj;
    return p;
}

// Factor of the time component(s) at the thread's quad in work row wr (0-based work row): sqrt(factor_reg_static) on
// static pixels (tv_operators_CPU.py:148-150) times the per-voxel time scale (extension, README.md:258) at plane ql / frame t.
// TSMODE 0: none (uniform weight), 1: mask_static, 2: time-scale map (and mask_static when present).
template <typename T, int VEC, int TSMODE>
PYTVB_HD void tile_time_factor(T* f, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, int wr, int cj, int ql, int t) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = T(1);
    if constexpr (TSMODE >= 1) {
        if (c.Ms) ld_into<T, VEC>(f, c.Ms + wr * TileC<VEC>::WJ + cj + VEC);
    }
    if constexpr (TSMODE == 2) {
        const int gi = clampi(c.i0 + wr - 1, 0, P.Ni - 1);
        const T* row = TS.row(P, ql, t, gi);
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] *= row[clampi(c.j0 + cj + e, 0, P.Nj - 1)];
    }
}
// The time scale alone at (plane ql, frame t) (the inverse norms travel with their voxel's scale along t, strip_quad_G_impl).
template <typename T, int VEC>
PYTVB_HD void tile_time_scale(T* f, const TileCtx<T>& c, const Params<T>& P, const ImgView<T>& TS, int wr, int cj, int ql, int t) {
    const int gi = clampi(c.i0 + wr - 1, 0, P.Ni - 1);
    const T* row = TS.row(P, ql, t, gi);
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] = row[clampi(c.j0 + cj + e, 0, P.Nj - 1)];
}

// Existence factor of a centred difference at index k of an axis of length L (tv_operators_CPU.py:331-358).
template <typename T>
PYTVB_HD T cen_exists(long long k, long long L) { return (k >= 1 && k <= L - 2) ? T(1) : T(0); }

// w = 1 / (div^2 |D x|) = rsqrt(s) / div from the sum s of the squared raw differences (|D x| = sqrt(s) / div; div = the scheme's
// divisor): with the 1 / div^2 of the adjoint folded into w the edge terms sum to G directly, and |D x| = s * w.  The reference sets 0/0 := 0 (tv_GPU.py:87-88: the norm 0 becomes inf before the
// division).  Every product the sub-gradient forms with w(v) has a difference that is one of the terms of s(v) as its other
// factor - the edge terms (x_b - x_a) S(w_a, w_b) pair each w with a difference of its own voxel, the centred C(m) likewise -
// so where s = 0 that factor is exactly 0 and ANY FINITE w gives the reference's 0.  EXACT = false (the kernel's fast path)
// therefore skips the select: float on the device is max + MUFU.RSQ + one multiply (a sum of squares below FLT_MIN - all
// differences < 1.1e-19 - is raised to FLT_MIN, which keeps w finite).  EXACT = true also reports s > 0 (the norms output
// needs the inf) and returns w = 0 there.
template <typename T, bool EXACT>
PYTVB_HD void tile_norm(T s, const Params<T>& P, T& w, bool& pos) {
    pos = s > T(0);
    w = pos ? fast_rsqrt(s) * P.inv_div : T(0);
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ void tile_norm<float, true>(float s, const Params<float>& P, float& w, bool& pos) {
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(s, 1.17549435e-38f)));
    pos = s >= 1.17549435e-38f;
    w = pos ? rs * P.inv_div : 0.0f;
}
template <>
__device__ __forceinline__ void tile_norm<float, false>(float s, const Params<float>& P, float& w, bool& pos) {
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(s, 1.17549435e-38f)));
    pos = true;
    w = rs * P.inv_div;
}
#endif

// Initial z state of a chunk whose first step is plane p: needs the thread's own x(p-1) (straight from global memory, once).
template <typename T, int VEC, int SCHEME, int R>
PYTVB_HD void tile_init_z(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const ImgView<T>& X, const Params<T>& P, int p,
                          const TilePos& tp) {
    constexpr int PX = TileC<VEC>::PX;
    const int qm = tile_clamp_plane(P, p - 1);
    const T* Xp = c.Xs + (long long)tile_slot(p) * g.xslot + (long long)tp.fl * g.slotX;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r;
        const int gi = clampi(c.i0 + rr, 0, P.Ni - 1);
        const T* src = X.row(P, qm, c.t0 + tp.fl, gi);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const T xm = src[clampi(c.j0 + tp.cj + e, 0, P.Nj - 1)];
            st.a[r][e] = SCHEME == CENTRAL ? xm : Xp[(rr + 2) * PX + tp.cj + 2 * VEC + e] - xm;
            st.w[r][e] = st.e[r][e] = st.g[r][e] = T(0);
        }
    }
}

// Elements just left / right of a thread's quad in its row.  A warp covers the 32 consecutive quads of one window row, so
// on the device they come from the neighbouring lanes' registers (two shuffles).  Lanes 0 / 31 have no such neighbour: with
// `edge` they read the element from the window (the scalar halo columns of x), else the quad's own end element stands in.
// Only the scalar path (VEC == 1) needs `edge`: of a halo lane's results only the w of the element next to the tile is ever
// used, and for VEC >= 2 that element's differences stay inside the quad and the tile-side neighbour lane.  The scalar loads this replaces had 4-way bank conflicts (lane stride 4 words): 24
// of the 64 shared-memory wavefronts per row and warp (profiles/r02f_*).  N elements per side (1, or 2 for the centred
// column terms): l[0] = element -1, l[1] = element -2; r[0] = element VEC, r[1] = element VEC+1.  `lane` = the quad's index
// in the row.  The host emulation reads the window where the device shuffles.
template <typename T, int VEC, int N>
PYTVB_HD void quad_sides(T* l, T* r, const T* q, const T* row, int lane, bool edge) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int dl = k / VEC + 1, ix = k % VEC;       // element -1-k is element VEC-1-ix of the quad dl lanes to the left
#if defined(__CUDA_ARCH__)
        T sl = __shfl_up_sync(0xffffffffu, q[VEC - 1 - ix], dl), sr = __shfl_down_sync(0xffffffffu, q[ix], dl);
#else
        T sl = lane >= dl ? row[-1 - k] : T(0), sr = lane + dl <= 31 ? row[VEC + k] : T(0);
#endif
        // lanes without that neighbour (written as independent predicated assignments: no branches in the row loop): the window
        // has the element when it is not further out than the scalar halo column, else the quad's own end element stands in
        if (edge) {
            if (lane < dl && lane * VEC >= k) sl = row[-1 - k];
            if (lane + dl > 31 && (31 - lane) * VEC >= k) sr = row[VEC + k];
            if (lane < dl && lane * VEC < k) sl = q[0];
            if (lane + dl > 31 && (31 - lane) * VEC < k) sr = q[VEC - 1];
        } else {
            if (lane < dl) sl = q[0];
            if (lane + dl > 31) sr = q[VEC - 1];
        }
        l[k] = sl;
        r[k] = sr;
    }
}

// S(w_k, w_{k+1}) of an edge term for a whole quad (pair_w element-wise).
template <typename T, int VEC, int SCHEME>
PYTVB_HD void vpair_w(T* d, const T* wk, const T* wk1) {
    if (SCHEME == HYBRID) VOp<T, VEC>::add(d, wk, wk1);
    else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) d[e] = SCHEME == UPWIND ? wk[e] : wk1[e];
    }
}

// ---- w-phase of plane p.  Outputs: w(p) to shared memory, norms (optional) and the TV partial sum for output voxels, and -
// with the z axis on - the finished G(p-1).
// Instruction diet (the kernel is instruction-issue bound: profiles/r02k_tv_tile_tma_ncu_full.txt, 371 instructions per quad of
// which 157 floating point):
//   * the loop-carried state is three values per voxel - the raw z difference a, the running sub-gradient g (in-plane + time
//     part of plane p-1 plus its incoming z term) and, within a step, the z term e = srz * term(p-1 -> p) handed from the
//     w-phase to the G-phase - and every carried value is assigned exactly once per step by the arithmetic instruction that
//     produces it (a register-to-register copy per value and step otherwise: 12 MOVs per quad row): a is recomputed from the
//     rows still in registers, g = e + Gp in the G-phase, and the previous plane's own w comes back from the w window (it is
//     still there until this phase overwrites it) instead of living in registers across the step;
//   * no select on the norm (tile_norm, EXACT only for the norms output), the TV sum as two packed FMAs per row with a 0/1
//     factor instead of a predicate and a scalar sum per row;
//   * window offsets per thread computed once (TilePos), the hybrid scheme's backward row difference is the forward one of
//     the row above.
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE, bool NORMS>
PYTVB_HD void tile_phase_w(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS,
                           T* G, T* norms, int p, const TilePos& tp) {
    typedef Comp<SCHEME, Z_ON, T_ON> C;
    typedef VOp<T, VEC> V;
    constexpr bool FWD = C::NEED_FWD, BWD = C::NEED_BWD, CEN = SCHEME == CENTRAL;
    constexpr int PX = TileC<VEC>::PX, WJ = TileC<VEC>::WJ;
    const T* Xp = c.Xs + (long long)tile_slot(p) * g.xslot + tp.xo;            // own quad in the x window of plane p, work row 0
    const T* Xn = c.Xs + (long long)tile_slot(p + 1) * g.xslot + tp.xo;
    T* Wp = c.Ws + tp.wo;
    const int t = tp.t;
    const int ql = tile_clamp_plane(P, p);
    const long long zg = P.zg0 + p;
    const bool plane_out = p >= c.zc0 && p < c.zc1, prev_out = Z_ON && p - 1 >= c.zc0 && p - 1 < c.zc1;
    const T rz2 = P.srz * P.srz, rt2 = P.srt * P.srt;
    const T fz = CEN ? cen_exists<T>(zg, P.NzG) : T(1), ft = CEN ? cen_exists<T>(t, P.M) : T(1);
    T* Gq = G + (long long)(p - 1) * P.sZ + tp.goff;            // G(p-1) at the thread's quad, work row 0
    keep_in_registers(Gq);
    const T colf = tp.col_out ? T(1) : T(0);
    T tvq[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) tvq[e] = T(0);
    T xu[VEC], xc[VEC], dip[VEC];
    ld_into<T, VEC>(xu, Xp - PX);
    ld_into<T, VEC>(xc, Xp);
    if (FWD && BWD) V::sub(dip, xc, xu);           // x(rr) - x(rr-1): the backward row difference of row rr
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r, gi = c.i0 + rr;
        T xd[VEC], s[VEC];
        ld_into<T, VEC>(xd, Xp + (r + 1) * PX);
        T cl, cr;
        quad_sides<T, VEC, 1>(&cl, &cr, xc, Xp + r * PX, tp.lane, VEC == 1);
        if (!CEN) {
            // column differences: VEC + 1 of them serve the forward and the backward component of the quad
            T djf[VEC], djb[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) djf[e] = (e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - xc[e];
#pragma unroll
            for (int e = 0; e < VEC; ++e) djb[e] = e > 0 ? (FWD ? djf[e > 0 ? e - 1 : 0] : xc[e] - xc[e > 0 ? e - 1 : 0]) : xc[0] - cl;
            T di[VEC];
            if (FWD) {
                V::sub(di, xd, xc);
                V::mul(s, di, di);
                V::fma(s, djf, djf, s);
            }
            if (BWD) {
                if (FWD) {
                    V::fma(s, dip, dip, s);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dip[e] = di[e];
                } else {
                    V::sub(di, xc, xu);
                    V::mul(s, di, di);
                }
                V::fma(s, djb, djb, s);
            }
        } else {
            const T fi = cen_exists<T>(gi, P.Ni);
            T di[VEC], dj[VEC];
            V::sub(di, xd, xu);
            V::muls(di, di, fi);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T fj = cen_exists<T>(c.j0 + tp.cj + e, P.Nj);
                dj[e] = fj * ((e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : cr) - (e > 0 ? xc[e > 0 ? e - 1 : 0] : cl));
            }
            V::mul(s, di, di);
            V::fma(s, dj, dj, s);
        }
        T xn[VEC], dzf[VEC];       // dzf, centred only: x(p+1) - x(p-1), 0 where it does not exist
        if (Z_ON) {
            T q[VEC];
            ld_into<T, VEC>(xn, Xn + r * PX);
            if (!CEN) {
                T d[VEC];
                V::sub(d, xn, xc);
                if (FWD) V::mul(q, d, d);
                if (BWD) { if (FWD) V::fma(q, st.a[r], st.a[r], q); else V::mul(q, st.a[r], st.a[r]); }
            } else {
                V::sub(dzf, xn, st.a[r]);
                V::muls(dzf, dzf, fz);
                V::mul(q, dzf, dzf);
            }
            V::fmas(s, q, rz2, s);
        }
        if (T_ON) {
            T xm[VEC], xp[VEC], q[VEC], d[VEC];
            ld_into<T, VEC>(xm, Xp + r * PX + tp.dxm);
            ld_into<T, VEC>(xp, Xp + r * PX + tp.dxp);
            if (!CEN) {
                if (FWD) { V::sub(d, xp, xc); V::mul(q, d, d); }
                if (BWD) { V::sub(d, xc, xm); if (FWD) V::fma(q, d, d, q); else V::mul(q, d, d); }
            } else {
                V::sub(d, xp, xm);
                V::muls(d, d, ft);
                V::mul(q, d, d);
            }
            if constexpr (TSMODE == 0) {
                V::fmas(s, q, rt2, s);
            } else {
                T fac[VEC];
                tile_time_factor<T, VEC, TSMODE>(fac, c, g, P, TS, rr + 1, tp.cj, ql, t);
                V::muls(fac, fac, P.srt);
                V::mul(fac, fac, fac);
                V::fma(s, fac, q, s);
            }
        }
        Pack<T, VEC> wq;
        T nrv[VEC];
        bool posv[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) tile_norm<T, NORMS>(s[e], P, wq.v[e], posv[e]);
        V::mul(nrv, s, wq.v);             // |D x| = sqrt(s) / div = s * w (0 by itself where s = 0)
        const bool row_out = rr >= 0 && rr < g.TI && gi < P.Ni;        // warp-uniform
        V::fmas(tvq, nrv, row_out ? colf : T(0), tvq);
        T wold[VEC];                      // this thread's w(p-1): still in the w window until the store below
        if (Z_ON && !CEN && FWD) ld_into<T, VEC>(wold, Wp + r * WJ);
        st_pack<T, VEC>(Wp + r * WJ, wq);
        if constexpr (NORMS) {
            if (plane_out && row_out && tp.col_out) {
                Pack<T, VEC> nq;
#pragma unroll
                for (int e = 0; e < VEC; ++e) nq.v[e] = posv[e] ? nrv[e] : T(INFINITY);
                st_pack<T, VEC>(norms + (long long)p * P.sZ + tp.goff + (long long)r * P.Nj, nq);
            }
        }
        if (Z_ON) {
            // z edge term between planes p-1 and p (centred: Cz(p)); it completes G(p-1)
            Pack<T, VEC> gq;
            if (!CEN) {
                T sw[VEC];
                vpair_w<T, VEC, SCHEME>(sw, wold, wq.v);
                V::mul(sw, st.a[r], sw);
                V::muls(st.e[r], sw, P.srz);
                V::sub(gq.v, st.g[r], st.e[r]);
                V::sub(st.a[r], xn, xc);
            } else {
                T en[VEC];
                V::mul(en, dzf, wq.v);
                V::muls(en, en, P.srz);
                V::sub(gq.v, st.g[r], en);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    st.g[r][e] = st.e[r][e];        // srz * Cz(p-1): the incoming z term of G(p)
                    st.e[r][e] = en[e];
                    st.a[r][e] = xc[e];
                }
            }
            if (prev_out && row_out && tp.col_out) st_pack<T, VEC>(Gq + (long long)r * P.Nj, gq);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) { st.w[r][e] = wq.v[e]; xu[e] = xc[e]; xc[e] = xd[e]; }
    }
    if (plane_out) {
        T sum = tvq[0];
#pragma unroll
        for (int e = 1; e < VEC; ++e) sum += tvq[e];
        st.tv += (double)sum;
    }
}

// ---- G-phase of plane p: in-plane and time edge terms; with the z axis on they start the running sub-gradient of plane p
// (st.g = st.e + Gp for the one-sided / hybrid schemes, st.g += Gp for the centred one), else they are the finished G(p).
template <typename T, int VEC, int SCHEME, bool Z_ON, bool T_ON, int R, int TSMODE>
PYTVB_HD void tile_phase_g(TileThread<T, VEC, R>& st, const TileCtx<T>& c, const TileGeom& g, const Params<T>& P, const ImgView<T>& TS, T* G, int p,
                           const TilePos& tp) {
    typedef VOp<T, VEC> V;
    constexpr bool CEN = SCHEME == CENTRAL;
    constexpr int PX = TileC<VEC>::PX, WJ = TileC<VEC>::WJ;
    const int wcol = tp.cj + VEC;
    const T* Xp = c.Xs + (long long)tile_slot(p) * g.xslot + tp.xo;
    const T* Wp = c.Ws + tp.wo;
    const int t = tp.t;
    const int lane = tp.lane;              // halo lanes (0 and 31) have no outer neighbour in the w window: their results are unused
    bool have = false;
    T tdn[VEC];          // one-sided / hybrid: row term (rr -> rr+1) of the previous row
    T cu[VEC], cc[VEC];  // centred: C_i(rr-1), C_i(rr)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int rr = tp.rr0 + r, gi = c.i0 + rr;
        const bool row_out = rr >= 0 && rr < g.TI;          // warp-uniform: halo rows carry no output
        if (!row_out) {
            have = false;
            continue;
        }
        const T* xrow = Xp + r * PX;
        const T* wrow = Wp + r * WJ;
        T xc[VEC], wc[VEC], gq[VEC];
        ld_into<T, VEC>(xc, xrow);
#pragma unroll
        for (int e = 0; e < VEC; ++e) wc[e] = st.w[r][e];
        // ---- rows
        if (!CEN) {
            T xd[VEC], wd[VEC], tup[VEC], dx[VEC], sw[VEC];
            if (!have) {
                T xu[VEC], wu[VEC];
                ld_into<T, VEC>(xu, xrow - PX);
                ld_into<T, VEC>(wu, wrow - WJ);
                V::sub(dx, xc, xu);
                vpair_w<T, VEC, SCHEME>(sw, wu, wc);
                V::mul(tup, dx, sw);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) tup[e] = tdn[e];
            }
            ld_into<T, VEC>(xd, xrow + PX);
            if (r + 1 < R) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wd[e] = st.w[r + 1 < R ? r + 1 : r][e];
            } else ld_into<T, VEC>(wd, wrow + WJ);
            V::sub(dx, xd, xc);
            vpair_w<T, VEC, SCHEME>(sw, wc, wd);
            V::mul(tdn, dx, sw);
            V::sub(gq, tup, tdn);
        } else {
            // C_i(m) = exists(m) * (x(m+1) - x(m-1)) * w(m);  G_i(rr) = C_i(rr-1) - C_i(rr+1)
            T xu[VEC], xd[VEC], xd2[VEC], wd[VEC], cd[VEC], dx[VEC];
            ld_into<T, VEC>(xd, xrow + PX);
            ld_into<T, VEC>(xd2, xrow + 2 * PX);
            if (!have) {
                T xu2[VEC], wu[VEC];
                ld_into<T, VEC>(xu, xrow - PX);
                ld_into<T, VEC>(xu2, xrow - 2 * PX);
                ld_into<T, VEC>(wu, wrow - WJ);
                const T fu = cen_exists<T>(gi - 1, P.Ni), fc = cen_exists<T>(gi, P.Ni);
                V::sub(dx, xc, xu2);
                V::mul(cu, dx, wu);
                V::muls(cu, cu, fu);
                V::sub(dx, xd, xu);
                V::mul(cc, dx, wc);
                V::muls(cc, cc, fc);
            }
            if (r + 1 < R) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) wd[e] = st.w[r + 1 < R ? r + 1 : r][e];
            } else ld_into<T, VEC>(wd, wrow + WJ);
            const T fd = cen_exists<T>(gi + 1, P.Ni);
            V::sub(dx, xd2, xc);
            V::mul(cd, dx, wd);
            V::muls(cd, cd, fd);
            V::sub(gq, cu, cd);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { cu[e] = cc[e]; cc[e] = cd[e]; }
        }
        have = true;
        // ---- columns (element-shifted within the quad: scalar code)
        if (!CEN) {
            T xl, xr, wlv, wrv;
            quad_sides<T, VEC, 1>(&xl, &xr, xc, xrow, lane, VEC == 1);
            quad_sides<T, VEC, 1>(&wlv, &wrv, wc, wrow, lane, false);
            T tj = (xc[0] - xl) * pair_w<T, SCHEME>(wlv, wc[0]);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const T xe = e + 1 < VEC ? xc[e + 1 < VEC ? e + 1 : e] : xr, we = e + 1 < VEC ? wc[e + 1 < VEC ? e + 1 : e] : wrv;
                const T tn = (xe - xc[e]) * pair_w<T, SCHEME>(wc[e], we);
                gq[e] += tj - tn;
                tj = tn;
            }
        } else {
            // x with 2, w with 1 element on each side; for the halo lanes the outer elements are clamped (results unused)
            T xw[VEC + 4], ww[VEC + 2];
#pragma unroll
            for (int e = 0; e < VEC; ++e) { xw[e + 2] = xc[e]; ww[e + 1] = wc[e]; }
            T xs_l[2], xs_r[2];
            quad_sides<T, VEC, 2>(xs_l, xs_r, xc, xrow, lane, VEC == 1);
            xw[0] = xs_l[1]; xw[1] = xs_l[0]; xw[VEC + 2] = xs_r[0]; xw[VEC + 3] = xs_r[1];
            quad_sides<T, VEC, 1>(&ww[0], &ww[VEC + 1], wc, wrow, lane, false);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int gj = c.j0 + tp.cj + e;
                const T cm = cen_exists<T>(gj - 1, P.Nj) * ((xw[e + 2] - xw[e]) * ww[e]);
                const T cp = cen_exists<T>(gj + 1, P.Nj) * ((xw[e + 4] - xw[e + 2]) * ww[e + 2]);
                gq[e] += cm - cp;
            }
        }
        // ---- time
        if (T_ON) {
            const int dm = tp.fl > 0 ? -1 : 0, dp = tp.fl < g.FC - 1 ? 1 : 0;          // neighbouring frames, clamped
            T xm[VEC], xp[VEC], wm[VEC], wp[VEC], wq[VEC], v[VEC], v2[VEC], dx[VEC], sw[VEC];
            ld_into<T, VEC>(wm, wrow + tp.dwm);
            ld_into<T, VEC>(wp, wrow + tp.dwp);
#pragma unroll
            for (int e = 0; e < VEC; ++e) wq[e] = wc[e];
            if constexpr (TSMODE == 2) {   // along t the inverse norms travel with their voxel's scale (strip_quad_G_impl)
                T f[VEC];
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t + dm);
                V::mul(wm, wm, f);
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t);
                V::mul(wq, wq, f);
                tile_time_scale<T, VEC>(f, c, P, TS, rr + 1, tp.cj, p, t + dp);
                V::mul(wp, wp, f);
            }
            if (!CEN) {
                ld_into<T, VEC>(xm, xrow + tp.dxm);
                ld_into<T, VEC>(xp, xrow + tp.dxp);
                V::sub(dx, xc, xm);
                vpair_w<T, VEC, SCHEME>(sw, wm, wq);
                V::mul(v, dx, sw);
                V::sub(dx, xp, xc);
                vpair_w<T, VEC, SCHEME>(sw, wq, wp);
                V::mul(v2, dx, sw);
                V::sub(v, v, v2);
                V::muls(v, v, P.srt);
            } else {
                // C_t(t-1) = exists(t-1) (x(t) - x(t-2)) w(t-1);  C_t(t+1) = exists(t+1) (x(t+2) - x(t)) w(t+1)
                const int dm2 = tp.fl > 1 ? -2 : -tp.fl, dp2 = tp.fl < g.FC - 2 ? 2 : g.FC - 1 - tp.fl;
                ld_into<T, VEC>(xm, xrow + dm2 * g.slotX);
                ld_into<T, VEC>(xp, xrow + dp2 * g.slotX);
                const T am = P.srt * cen_exists<T>(t - 1, P.M), ap = P.srt * cen_exists<T>(t + 1, P.M);
                V::sub(dx, xc, xm);
                V::mul(v, dx, wm);
                V::muls(v, v, am);
                V::sub(dx, xp, xc);
                V::mul(v2, dx, wp);
                V::muls(v2, v2, ap);
                V::sub(v, v, v2);
            }
            if constexpr (TSMODE >= 1) {
                if (c.Ms) {
                    T f[VEC];
                    ld_into<T, VEC>(f, c.Ms + (rr + 1) * WJ + wcol);
                    V::mul(v, v, f);
                }
            }
            V::add(gq, gq, v);
        }
        if (Z_ON) {
            if (!CEN) V::add(st.g[r], st.e[r], gq);
            else V::add(st.g[r], st.g[r], gq);
        } else {
            if (tp.col_out && gi < P.Ni) {
                Pack<T, VEC> pk;
#pragma unroll
                for (int e = 0; e < VEC; ++e) pk.v[e] = gq[e];
                st_pack<T, VEC>(G + (long long)p * P.sZ + tp.goff + (long long)r * P.Nj, pk);
            }
        }
    }
}

}  // namespace pytvb
