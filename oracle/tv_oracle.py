"""CPU oracle for the PyTV-4D total-variation hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the algorithm of the reference's CPU path
(`pytv/tv_operators_CPU.py`, `pytv/tv_CPU.py`, and the Chambolle-Pock loop of `README.md:139-158`).
It is the checker the CUDA kernels are compared against.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
(`pytv-4d_b200/`) never does and has no CPU fallback.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference from
`/root/reference` (possible only in the build container) and stores its outputs as fixtures under
`tests/golden/`; `tests/test_oracle_golden.py` checks every function below against those fixtures and
against the known answers the reference publishes (README.md:91, the 5x5 delta images of
examples/b_TV_discretizations_math.ipynb).  One case is unpinned: the central scheme with Nz == 2 and
z-regularisation on, where the reference itself raises; the documented intent
("use upwind instead", reference README.md:236) is implemented here.

Design note: the reference spells every scheme out as explicit slice assignments.  The restatement is
organised differently on purpose: a scheme is a *list of components*, each `(axis, kind)` with kind in
{forward, backward, centred}, and three axis-generic primitives (difference, adjoint, subgradient
scatter) do the work.  Layouts are the reference's: images `(Nz, M, N, N)`, gradient fields
`(Nz, Nd, M, N, N)`.

Deliberate deviation: the output dtype always equals the input dtype (float32 stays
float32); the reference's CPU path upcasts float32 to float64 in the final `/np.sqrt(2.0)` under
numpy >= 2, an accident of NEP 50.
"""
import numpy as np

SCHEMES = ("upwind", "downwind", "central", "hybrid")

# axes of a (Nz, M, N, N) image
_AX_Z, _AX_T, _AX_I, _AX_J = 0, 1, 2, 3


# ----------------------------------------------------------------------------- scheme description
def _axes_on(Nz, M, reg_z_over_reg, reg_time):
    """Which optional axes carry a component (tv_operators_CPU.py:110-114, :190-194, :256-260).
    A NaN weight fails `> 0` and so disables the axis (the `== np.nan` test at :100 is dead code)."""
    z_on = bool(Nz > 1 and reg_z_over_reg > 0)
    t_on = bool(M > 1 and reg_time > 0)
    return z_on, t_on


def components(scheme, Nz, M, reg_z_over_reg=1.0, reg_time=0.0):
    """Ordered component list [(axis, kind)], kind in 'f' (x[k+1]-x[k]), 'b' (x[k]-x[k-1]), 'c'
    (x[k+1]-x[k-1]).  Order is part of the contract: row, col, [z], [t]; hybrid interleaves forward and
    backward as row_f, col_f, row_b, col_b, [z_f, z_b], [t_f, t_b] (tv_operators_CPU.py:117-152)."""
    z_on, t_on = _axes_on(Nz, M, reg_z_over_reg, reg_time)
    if scheme == "hybrid":
        comps = [(_AX_I, "f"), (_AX_J, "f"), (_AX_I, "b"), (_AX_J, "b")]
        if z_on:
            comps += [(_AX_Z, "f"), (_AX_Z, "b")]
        if t_on:
            comps += [(_AX_T, "f"), (_AX_T, "b")]
        return comps
    kind = {"upwind": "f", "downwind": "b", "central": "c"}[scheme]
    comps = [(_AX_I, kind), (_AX_J, kind)]
    if z_on:
        comps.append((_AX_Z, kind))
    if t_on:
        comps.append((_AX_T, kind))
    return comps


def num_components(scheme, Nz, M, reg_z_over_reg=1.0, reg_time=0.0):
    return len(components(scheme, Nz, M, reg_z_over_reg, reg_time))


def _global_divisor(scheme, dtype):
    """hybrid: whole field / sqrt(2) (tv_operators_CPU.py:154); central: / 2 (:358)."""
    if scheme == "hybrid":
        return dtype.type(np.sqrt(2.0))
    if scheme == "central":
        return dtype.type(2.0)
    return None


def _sl(axis, s, ndim=4):
    idx = [slice(None)] * ndim
    idx[axis] = s
    return tuple(idx)


def _effective_kind(kind, L, axis):
    """Centred differences on a z or time axis of length 2 fall back to the forward difference
    (tv_operators_CPU.py:339-340 for z, :347-348 for time).  Rows and columns have no such fallback:
    a 2x2 image simply has zero centred differences (:331-334)."""
    return "f" if (kind == "c" and L == 2 and axis in (_AX_Z, _AX_T)) else kind


# ----------------------------------------------------------------------------- axis primitives
def _difference(x, axis, kind):
    """One difference component; out-of-range *differences* are zero (the arrays start as zeros:
    tv_operators_CPU.py:115, :196, :262, :328)."""
    L = x.shape[axis]
    out = np.zeros_like(x)
    kind = _effective_kind(kind, L, axis)
    if L < 2:
        return out
    if kind == "f":      # :265-268
        out[_sl(axis, slice(0, L - 1))] = x[_sl(axis, slice(1, L))] - x[_sl(axis, slice(0, L - 1))]
    elif kind == "b":    # :199-202
        out[_sl(axis, slice(1, L))] = x[_sl(axis, slice(1, L))] - x[_sl(axis, slice(0, L - 1))]
    else:                # :331-334
        if L > 2:
            out[_sl(axis, slice(1, L - 1))] = x[_sl(axis, slice(2, L))] - x[_sl(axis, slice(0, L - 2))]
    return out


def _adjoint_accumulate(out, p, axis, kind, weight=None):
    """Exact transpose of `_difference`: entries of p where the difference is structurally zero are
    ignored (tv_operators_CPU.py:555-560 forward, :488-493 backward, :623-628 centred)."""
    L = p.shape[axis]
    kind = _effective_kind(kind, L, axis)
    if L < 2:
        return
    if kind == "f":
        src, plus, minus = slice(0, L - 1), slice(1, L), slice(0, L - 1)
    elif kind == "b":
        src, plus, minus = slice(1, L), slice(1, L), slice(0, L - 1)
    else:
        if L < 3:
            return
        src, plus, minus = slice(1, L - 1), slice(2, L), slice(0, L - 2)
    v = p[_sl(axis, src)]
    if weight is not None:
        v = weight * v
    out[_sl(axis, plus)] += v
    out[_sl(axis, minus)] -= v


def _mask_plane(mask_static, N_i, N_j):
    """mask_static is a boolean (1,1,N,N) array, or False (tv_operators_CPU.py:148)."""
    if isinstance(mask_static, bool):
        return None
    m = np.asarray(mask_static).astype(bool)
    return np.broadcast_to(m.reshape(m.shape[-2], m.shape[-1]), (N_i, N_j))


# ----------------------------------------------------------------------------- operators
def _time_scale(time_weight, shape, dt):
    """EXTENSION (the reference's TODO, README.md:258: "replace mask_static, factor_reg_static with a weight matrix of
    size Nz x M x N x N"): per-voxel weight of the time regularisation; like reg_time it enters as its square root,
    applied to the time component(s) at the voxel where the component lives.  No reference implementation exists:
    pinned only by adjointness and by its reduction to mask_static for weights in {1, factor} (tests)."""
    if time_weight is None:
        return None
    w = np.asarray(time_weight, dtype=np.float64)
    return np.sqrt(np.broadcast_to(w, shape)).astype(dt)


def D(img, scheme, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0.0, time_weight=None):
    """Forward operator of `scheme`: (Nz,M,N,N) -> (Nz,Nd,M,N,N).
    Restates D_upwind / D_downwind / D_central / D_hybrid (tv_operators_CPU.py:222, :156, :288, :76)."""
    img = np.asarray(img)
    Nz, M, Ni, Nj = img.shape
    dt = img.dtype
    comps = components(scheme, Nz, M, reg_z_over_reg, reg_time)
    out = np.zeros((Nz, len(comps), M, Ni, Nj), dtype=dt)
    s_z = dt.type(np.sqrt(reg_z_over_reg)) if reg_z_over_reg > 0 else dt.type(0)
    s_t = dt.type(np.sqrt(reg_time)) if reg_time > 0 else dt.type(0)
    s_f = dt.type(np.sqrt(factor_reg_static))
    mplane = _mask_plane(mask_static, Ni, Nj)
    tsc = _time_scale(time_weight, img.shape, dt)
    for d, (axis, kind) in enumerate(comps):
        c = _difference(img, axis, kind)
        if axis == _AX_Z:
            c = s_z * c                                    # :273
        elif axis == _AX_T:
            c = s_t * c                                    # :278
            if mplane is not None:                         # :280-282
                c = np.where(mplane[None, None], c * s_f, c)
            if tsc is not None:
                c = c * tsc
        out[:, d] = c
    div = _global_divisor(scheme, dt)
    if div is not None:
        out = out / div
    return out


def D_T(p, scheme, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0.0, time_weight=None):
    """Adjoint operator: (Nz,Nd,M,N,N) -> (Nz,M,N,N).
    Restates D_T_upwind / D_T_downwind / D_T_central / D_T_hybrid (tv_operators_CPU.py:518, :450, :585, :360)."""
    p = np.asarray(p)
    Nz, Nd, M, Ni, Nj = p.shape
    dt = p.dtype
    comps = components(scheme, Nz, M, reg_z_over_reg, reg_time)
    if len(comps) != Nd:
        raise IndexError("field has %d components, scheme %s expects %d" % (Nd, scheme, len(comps)))
    s_z = dt.type(np.sqrt(reg_z_over_reg)) if reg_z_over_reg > 0 else dt.type(0)
    s_t = dt.type(np.sqrt(reg_time)) if reg_time > 0 else dt.type(0)
    s_f = dt.type(np.sqrt(factor_reg_static))
    out = np.zeros((Nz, M, Ni, Nj), dtype=dt)
    time_part = None
    tsc = _time_scale(time_weight, out.shape, dt)
    for d, (axis, kind) in enumerate(comps):
        if axis == _AX_Z:
            _adjoint_accumulate(out, p[:, d], axis, kind, s_z)          # :565-566
        elif axis == _AX_T:
            if time_part is None:
                time_part = np.zeros_like(out)                          # :571
            pd = p[:, d] if tsc is None else p[:, d] * tsc              # exact adjoint: the scale sits where the component lives
            _adjoint_accumulate(time_part, pd, axis, kind, s_t)         # :573-574
        else:
            _adjoint_accumulate(out, p[:, d], axis, kind)
    if time_part is not None:
        mplane = _mask_plane(mask_static, Ni, Nj)
        if mplane is not None:                                          # :577-579
            time_part = np.where(mplane[None, None], time_part * s_f, time_part)
        out += time_part
    div = _global_divisor(scheme, dt)
    if div is not None:
        out = out / div
    return out


def l21(D_img, return_array=False):
    """sum over voxels of the 2-norm over axis 1 (compute_L21_norm, tv_operators_CPU.py:45-74)."""
    D_img = np.asarray(D_img)
    norms = np.sqrt(np.sum(np.square(D_img), axis=1))
    total = np.sum(norms)
    return (total, norms) if return_array else total


# ----------------------------------------------------------------------------- direct API
def tv(img, scheme, mask=None, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False,
       factor_reg_static=0.0, return_grad_norms=False, time_weight=None):
    """TV value and the reference's subgradient (tv_hybrid/tv_downwind/tv_upwind/tv_central,
    tv_CPU.py:47, :131, :195, :258).

    `mask` zeroes `img` outside the mask IN PLACE, like the reference (tv_CPU.py:77-78).  Zero norms are
    replaced by inf (tv_CPU.py:86) so that 0/0 := 0, and the returned norms keep those infs.  The
    subgradient is assembled per component exactly like tv_CPU.py:92-124: every component divided by the
    norm is scattered with unit weight (the sqrt(reg) chain-rule factors are *not* applied: SURVEY B2).
    """
    if mask is not None and not (isinstance(mask, list) and len(mask) == 0):
        m = np.broadcast_to(np.asarray(mask).astype(bool), img.shape)
        img[~m] = 0
    Nz, M, Ni, Nj = img.shape
    field = D(img, scheme, reg_z_over_reg, reg_time, mask_static, factor_reg_static, time_weight)
    value, norms = l21(field, return_array=True)
    norms[norms == 0] = np.inf
    comps = components(scheme, Nz, M, reg_z_over_reg, reg_time)
    G = np.zeros_like(img, dtype=field.dtype)
    for d, (axis, kind) in enumerate(comps):
        q = field[:, d] / norms
        _adjoint_accumulate(G, q, axis, kind)   # unit weights: tv_CPU.py:239-251 / :176-188 / :302-326
    div = _global_divisor(scheme, field.dtype)
    if div is not None:
        G /= div                                # tv_CPU.py:124, :328
    if return_grad_norms:
        return value, G, norms
    return value, G


# ----------------------------------------------------------------------------- Chambolle-Pock
def project_l2_ball(p, radius):
    """Pointwise projection of the Nd-vector at every voxel onto the 2-ball of given radius
    (README.md:151, with keepdims so that it also works for Nz > 1: SURVEY B10)."""
    n = np.sqrt(np.sum(p * p, axis=1, keepdims=True))
    return p / np.maximum(p.dtype.type(1.0), n / p.dtype.type(radius))


def cp_readme_step(x, x0, y_f, y_tv, scheme="hybrid", lam=25.0, sigma_D=0.5, sigma_A=1.0, tau=1.0 / 9.0,
                   **weights):
    """One iteration of the README's "simple" Chambolle-Pock loop (README.md:145-157):
    dual fidelity variable y_f, dual TV variable y_tv, no over-relaxation.
    Returns (x, y_f, y_tv, loss)."""
    dt = x.dtype.type
    y_f = (y_f + dt(sigma_A) * (x - x0)) / dt(1.0 + sigma_A)             # :148
    Dx = D(x, scheme, **weights)                                         # :149
    y_tv = project_l2_ball(y_tv + dt(sigma_D) * Dx, lam)                 # :150-151
    x = x - dt(tau) * y_f - dt(tau) * D_T(y_tv, scheme, **weights)       # :154
    loss = 0.5 * np.sum(np.square(x - x0), dtype=np.float64) + lam * np.float64(l21(Dx))   # :157
    return x, y_f, y_tv, float(loss)


def cp_rof_step(x, xbar, x0, y, scheme="hybrid", lam=0.1, sigma=0.5, tau=1.0 / 13.0, theta=1.0, **weights):
    """One iteration of the standard Chambolle-Pock algorithm for 0.5||x-x0||^2 + lam*TV(x)
    (Chambolle & Pock 2011, Alg. 1, cited at README.md:140): dual ascent on D(xbar) with projection,
    primal descent with the exact prox of the data term, then over-relaxation.
    Returns (x, xbar, y, primal_energy) with primal_energy = 0.5||x_new-x0||^2 + lam*L21(D xbar_old)."""
    dt = x.dtype.type
    Dxb = D(xbar, scheme, **weights)
    y = project_l2_ball(y + dt(sigma) * Dxb, lam)
    x_new = (x - dt(tau) * D_T(y, scheme, **weights) + dt(tau) * x0) / dt(1.0 + tau)
    xbar = x_new + dt(theta) * (x_new - x)
    energy = 0.5 * np.sum(np.square(x_new - x0), dtype=np.float64) + lam * np.float64(l21(Dxb))
    return x_new, xbar, y, float(energy)
