"""The peer-memory halo push schedule run inside ONE process: a volume cut into z-slabs ("ranks") that hold four halo
planes each and store their boundary planes straight into their neighbours' planes through the `_p2p` passes - what
CPSolver(comm="p2p") does across GPUs, minus the symmetric-memory mapping and the barrier (the slabs run one after the
other here).  Shared by the CPU test (host emulation of the kernel code) and the single-GPU test (the real kernels
through the C ABI).  Test infrastructure."""
import torch

from pytv_b200 import _lib

IMG_LO, IMG_HI, FLD_LO, FLD_HI = 0, 1, 2, 3


def run(ops, device, scheme, variant, dtype, shape=(6, 2, 8, 8), bounds=((0, 1), (1, 4), (4, 6)), iterations=3, rz=0.5, rt=0.25,
        lam=0.1, sigma=0.5, tau=0.07):
    """Returns ((x, aux, y) of the slab run concatenated, (x, aux, y) of the whole-volume run)."""
    Nz, M, Ni, Nj = shape
    g = torch.Generator().manual_seed(11)
    x0 = torch.rand(shape, generator=g, dtype=torch.float64).to(dtype).to(device)
    code = _lib.F32 if dtype == torch.float32 else _lib.F64
    c2 = 0.9 if variant == "rof" else 1.0
    pb_whole = _lib.make_problem(scheme, code, shape, rz, rt, 0.0, None, 0, Nz)
    z_on, t_on = Nz > 1 and rz > 0, M > 1 and rt > 0
    Nd = (4 + 2 * z_on + 2 * t_on) if scheme == "hybrid" else (2 + z_on + t_on)
    ws = ops.workspace(pb_whole, device)

    def fresh():
        x = x0.clone()
        aux = x0.clone() if variant == "rof" else torch.zeros_like(x0)
        y = torch.zeros((Nz, Nd, M, Ni, Nj), dtype=dtype, device=device)
        return x, aux, y

    # ---- whole volume
    xw, auxw, yw = fresh()
    for _ in range(iterations):
        ops.cp_dual(pb_whole, auxw if variant == "rof" else xw, yw, lam, sigma, None, None, None, ws)
        ops.cp_primal(variant, pb_whole, yw, xw, auxw, x0, tau, c2, None, None, None, ws)

    # ---- slabs with four halo planes each
    x, aux, y = fresh()
    nr = len(bounds)
    X = [x[a:b].contiguous() for a, b in bounds]
    A = [aux[a:b].contiguous() for a, b in bounds]
    Y = [y[a:b].contiguous() for a, b in bounds]
    X0 = [x0[a:b].contiguous() for a, b in bounds]
    H = [torch.full((4, M, Ni, Nj), float("nan"), dtype=dtype, device=device) for _ in bounds]
    PB = [_lib.make_problem(scheme, code, (b - a, M, Ni, Nj), rz, rt, 0.0, None, a, Nz) for a, b in bounds]
    need_img_lo, need_img_hi = scheme != "upwind", scheme != "downwind"
    need_fld_lo, need_fld_hi = scheme != "downwind", scheme != "upwind"
    U = A if variant == "rof" else X
    for r in range(nr):                        # the one start-up exchange of the image halos
        if r > 0:
            H[r][IMG_LO].copy_(U[r - 1][-1])
        if r < nr - 1:
            H[r][IMG_HI].copy_(U[r + 1][0])

    def ptr(r, slot, needed):
        return H[r][slot].data_ptr() if (0 <= r < nr and needed) else None

    def plane(r, slot, needed):
        return H[r][slot] if needed else None

    for _ in range(iterations):
        for r in range(nr):
            # my backward-type z slot -> the previous rank's fld_hi; my forward-type z slot -> the next rank's fld_lo
            ops.cp_dual_p2p(PB[r], U[r], Y[r], lam, sigma, None, plane(r, IMG_LO, r > 0 and need_img_lo),
                            plane(r, IMG_HI, r < nr - 1 and need_img_hi), ptr(r - 1, FLD_HI, need_fld_hi), ptr(r + 1, FLD_LO, need_fld_lo), ws)
        for r in range(nr):
            # my first plane -> the previous rank's img_hi; my last plane -> the next rank's img_lo
            ops.cp_primal_p2p(variant, PB[r], Y[r], X[r], A[r], X0[r], tau, c2, None, plane(r, FLD_LO, r > 0 and need_fld_lo),
                              plane(r, FLD_HI, r < nr - 1 and need_fld_hi), ptr(r - 1, IMG_HI, need_img_hi), ptr(r + 1, IMG_LO, need_img_lo), ws)
    return (torch.cat(X), torch.cat(A), torch.cat(Y)), (xw, auxw, yw)
