"""Host-side bookkeeping of the x-slab decomposition and a NumPy restatement of the partitioned (SPIKE-style)
line solve that the CUDA kernels implement along the decomposed axis (csrc/kernels_fast.cu MODE 1 / k_x_interface /
MODE 2).  TEST MODEL ONLY: used by tests/test_dist_gloo.py, which runs the same algebra over `torch.distributed`
(gloo, world size 2) - no GPU involved; the product (cmc_fluid_solver_b200/) never imports it.

Reference counterparts: GPUplan::splitEven1D (src/Common/GPUplan.cpp:122-141) for the split; the pipelined
distributed Thomas it replaces is LaunchSolveSegments_X (src/FluidSolver3D/AdiSolver3D.cu:524-640).
"""
from __future__ import annotations

import numpy as np


def split_even(dimx: int, nslabs: int):
    """[(x0, nx)] of every slab: dimx // n planes each, the remainder spread over the first slabs."""
    base, rem = divmod(dimx, nslabs)
    out, x0 = [], 0
    for r in range(nslabs):
        nx = base + (1 if r < rem else 0)
        out.append((x0, nx))
        x0 += nx
    return out


def lines_per_owner(ny: int, nz: int, nslabs: int) -> int:
    """x-lines (one per (j, k)) are dealt to the ranks in contiguous blocks for the interface solve."""
    return (ny * nz + nslabs - 1) // nslabs


def thomas(a, b, c, d):
    """Common::SolveTridiagonal (src/Common/Algorithms.h:21-38) for a batch: arrays (..., n)."""
    a, b, c, d = (np.array(v, dtype=np.float64) for v in (a, b, c, d))
    n = a.shape[-1]
    c[..., -1] = 0
    c[..., 0] /= b[..., 0]
    d[..., 0] /= b[..., 0]
    for i in range(1, n):
        den = b[..., i] - a[..., i] * c[..., i - 1]
        c[..., i] /= den
        d[..., i] = (d[..., i] - d[..., i - 1] * a[..., i]) / den
    x = np.empty_like(d)
    x[..., -1] = d[..., -1]
    for i in range(n - 2, -1, -1):
        x[..., i] = d[..., i] - c[..., i] * x[..., i + 1]
    return x


def slab_spike(a, b, c, d):
    """Spike pass of one slab (kernel MODE 1).  a, b, c, d: the slab's rows of every line, shape (lines, nx); a[:, 0]
    couples to the previous slab's last row (x_left), c[:, -1] to the next slab's first row (x_right).
    Returns f, pf, qf, l, pl, ql with   x_first = f - pf*x_left - qf*x_right ,  x_last = l - pl*x_left - ql*x_right."""
    a0 = a.copy(); a0[:, 0] = 0
    c0 = c.copy(); c0[:, -1] = 0
    e_first = np.zeros_like(d); e_first[:, 0] = a[:, 0]
    e_last = np.zeros_like(d); e_last[:, -1] = c[:, -1]
    y = thomas(a0, b, c0, d)
    p = thomas(a0, b, c0, e_first)
    q = thomas(a0, b, c0, e_last)
    return y[:, 0], p[:, 0], q[:, 0], y[:, -1], p[:, -1], q[:, -1]


def interface_solve(f, pf, qf, l, pl, ql):
    """Interface system (kernel k_x_interface).  Inputs have shape (P, lines): the spike coefficients of every slab.
    Unknowns F_r (first row of slab r) and L_r (last row):
        F_r + pf_r L_{r-1} + qf_r F_{r+1} = f_r ,   L_r + pl_r L_{r-1} + ql_r F_{r+1} = l_r
    solved by block Thomas over Z_r = (L_r, F_{r+1}).  Returns x_left, x_right of shape (P, lines): for each slab the
    solution of its neighbours' adjacent rows (0 where there is no neighbour)."""
    P, n = f.shape
    xl, xr = np.zeros((P, n)), np.zeros((P, n))
    if P == 1:
        return xl, xr
    i00 = np.zeros((P - 1, n)); i01 = np.zeros_like(i00); i10 = np.zeros_like(i00); i11 = np.zeros_like(i00)
    r0 = np.zeros_like(i00); r1 = np.zeros_like(i00)
    p_i01 = np.zeros(n); p_s = np.zeros(n)
    for r in range(P - 1):
        m01 = ql[r] - pl[r] * p_i01 * qf[r]
        idet = 1.0 / (1.0 - m01 * pf[r + 1])
        i00[r], i01[r], i10[r], i11[r] = idet, -m01 * idet, -pf[r + 1] * idet, idet
        r0[r] = l[r] - pl[r] * p_s
        r1[r] = f[r + 1]
        p_s = i00[r] * r0[r] + i01[r] * r1[r]
        p_i01 = i01[r]
    nextF = np.zeros(n)
    for r in range(P - 2, -1, -1):
        b0, b1 = r0[r], r1[r] - qf[r + 1] * nextF
        Lr = i00[r] * b0 + i01[r] * b1
        Fr1 = i10[r] * b0 + i11[r] * b1
        xl[r + 1] = Lr
        xr[r] = Fr1
        nextF = Fr1
    return xl, xr


def slab_coupled_solve(a, b, c, d, x_left, x_right):
    """Coupled sweep of one slab (kernel MODE 2): the neighbours' adjacent rows are known, so their terms move to
    the right-hand side and the slab's lines decouple."""
    d = d.copy()
    d[:, 0] -= a[:, 0] * x_left
    d[:, -1] -= c[:, -1] * x_right
    a0 = a.copy(); a0[:, 0] = 0
    c0 = c.copy(); c0[:, -1] = 0
    return thomas(a0, b, c0, d)


def slab_open_solve(a, b, c, d):
    """One-pass form (kernel k_tma_sweep XS / CTA pair): the slab's rows solved ONCE with two extra right-hand sides, so that
    every row is known up to the two neighbour values:   x = y - p * x_left - q * x_right   (shapes (lines, nx)).
    Row 0 / row nx-1 of (y, p, q) are the coefficients (f, pf, qf) / (l, pl, ql) that go to the other slabs."""
    a0 = a.copy(); a0[:, 0] = 0
    c0 = c.copy(); c0[:, -1] = 0
    e_first = np.zeros_like(d); e_first[:, 0] = a[:, 0]
    e_last = np.zeros_like(d); e_last[:, -1] = c[:, -1]
    return thomas(a0, b, c0, d), thomas(a0, b, c0, e_first), thomas(a0, b, c0, e_last)


def pair_interface(l0, ql0, f1, pf1):
    """Two parts of a line (the CTA pair of k_tma_sweep, or two slabs): L_0 + ql_0 F_1 = l_0 , F_1 + pf_1 L_0 = f_1 in closed
    form.  Returns (x_last of the lower part, x_first of the upper part)."""
    L0 = (l0 - ql0 * f1) / (1.0 - ql0 * pf1)
    return L0, f1 - pf1 * L0


def thomas_chain(a, b, c, d, slabs):
    """CMC_MODE_EXACT on a decomposed grid (exact_x_chain, kernels_exact.cu): the same recurrence, slab by slab - the forward
    elimination hands (c', d') of a slab's last row to the next slab, the back substitution hands the first row's solution back.
    Same operations in the same order as thomas(): the result is bit-identical."""
    a, b, c, d = (np.array(v, dtype=np.float64) for v in (a, b, c, d))
    c[..., -1] = 0
    cp, dp = np.empty_like(c), np.empty_like(d)
    carry = None
    for x0, nx in slabs:                                   # forward, first slab -> last
        for i in range(x0, x0 + nx):
            if i == 0:
                cp[..., 0] = c[..., 0] / b[..., 0]; dp[..., 0] = d[..., 0] / b[..., 0]
            else:
                pc, pd = (cp[..., i - 1], dp[..., i - 1]) if i > x0 else carry
                den = b[..., i] - a[..., i] * pc
                cp[..., i] = c[..., i] / den
                dp[..., i] = (d[..., i] - pd * a[..., i]) / den
        carry = (cp[..., x0 + nx - 1].copy(), dp[..., x0 + nx - 1].copy())       # the plane that goes to the next slab
    x = np.empty_like(d)
    up = None
    for x0, nx in reversed(slabs):                         # back substitution, last slab -> first
        for i in range(x0 + nx - 1, x0 - 1, -1):
            if i == d.shape[-1] - 1:
                x[..., i] = dp[..., i]
            else:
                nxt = x[..., i + 1] if i + 1 < x0 + nx else up
                x[..., i] = dp[..., i] - cp[..., i] * nxt
        up = x[..., x0].copy()                             # the plane that goes back to the previous slab
    return x
