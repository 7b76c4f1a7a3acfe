"""The N > 1 path on CPU: world_size-2 (and 3) `gloo` process groups run the host-side restatement of the
slab-decomposed x-solve (tests/partition_model.py - the algebra of the CUDA kernels MODE 1, k_x_interface,
MODE 2) with the same exchange pattern as the NCCL transport: all-to-all of the spike coefficients by line
ownership, interface solve on the owner, all-to-all of the neighbours' row values back."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import partition_model as P


def test_split_even_matches_reference_rule():
    assert P.split_even(512, 8) == [(64 * r, 64) for r in range(8)]
    assert P.split_even(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]          # remainder goes to the first slabs
    assert sum(nx for _, nx in P.split_even(1000, 7)) == 1000
    assert P.lines_per_owner(512, 512, 8) == 32768 and P.lines_per_owner(5, 7, 4) == 9


def _system(lines, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1.2, -0.2, (lines, n)); c = rng.uniform(-1.2, -0.2, (lines, n))
    b = 2.6 + rng.uniform(0, 1, (lines, n)); d = rng.normal(size=(lines, n))
    a[:, 0] = 0; c[:, -1] = 0
    # a few decoupling boundary rows like ApplyBC0/1 produce (a = 0 or c = 0 inside the line)
    a[::3, n // 3] = 0; c[1::4, n // 2] = 0
    return a, b, c, d


def test_partitioned_solve_single_process():
    a, b, c, d = _system(37, 96, 1)
    ref = P.thomas(a, b, c, d)
    for nsl in (2, 3, 4):
        slabs = P.split_even(96, nsl)
        co = [P.slab_spike(a[:, x0:x0 + nx], b[:, x0:x0 + nx], c[:, x0:x0 + nx], d[:, x0:x0 + nx]) for x0, nx in slabs]
        f, pf, qf, l, pl, ql = (np.stack([cc[i] for cc in co]) for i in range(6))
        xl, xr = P.interface_solve(f, pf, qf, l, pl, ql)
        for r, (x0, nx) in enumerate(slabs):
            x = P.slab_coupled_solve(a[:, x0:x0 + nx], b[:, x0:x0 + nx], c[:, x0:x0 + nx], d[:, x0:x0 + nx], xl[r], xr[r])
            assert np.abs(x - ref[:, x0:x0 + nx]).max() < 1e-12


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines, n = 40, 64 * world
        a, b, c, d = _system(lines, n, 7)                       # every rank builds the same global system ...
        x0, nx = P.split_even(n, world)[rank]                   # ... and keeps only its slab
        sl = slice(x0, x0 + nx)
        co = np.stack(P.slab_spike(a[:, sl], b[:, sl], c[:, sl], d[:, sl]))          # (6, lines)
        lpo = (lines + world - 1) // world
        pad = lpo * world
        send = np.zeros((world, 6, lpo)); tmp = np.zeros((6, pad)); tmp[:, :lines] = co
        for o in range(world):
            send[o] = tmp[:, o * lpo:(o + 1) * lpo]             # block o -> owner o
        recv = [torch.zeros(6, lpo, dtype=torch.float64) for _ in range(world)]
        for o in range(world):                                  # all-to-all by line ownership (gloo: one gather per owner)
            dist.gather(torch.from_numpy(send[o].copy()), recv if rank == o else None, dst=o)
        coef = np.stack([r.numpy() for r in recv])              # (src, 6, lpo) on the owner
        xl, xr = P.interface_solve(*(coef[:, i, :] for i in range(6)))          # (P, lpo) each
        back = [torch.zeros(2, lpo, dtype=torch.float64) for _ in range(world)]
        for o in range(world):                                  # owner o scatters (x_left, x_right) of its lines to every slab
            out = [torch.from_numpy(np.stack([xl[r], xr[r]]).copy()) for r in range(world)] if rank == o else None
            dist.scatter(back[o], out, src=o)
        bl = np.concatenate([t.numpy()[0] for t in back])[:lines]
        br = np.concatenate([t.numpy()[1] for t in back])[:lines]
        x = P.slab_coupled_solve(a[:, sl], b[:, sl], c[:, sl], d[:, sl], bl, br)
        ref = P.thomas(a, b, c, d)[:, sl]
        err = torch.tensor([float(np.abs(x - ref).max())], dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        # distributed residual-style reduction (sum, count) like the all-reduce of EvalDivError partials
        part = torch.tensor([float(np.abs(x).sum()), float(x.size)], dtype=torch.float64)
        dist.all_reduce(part)
        if rank == 0:
            ret["err"] = float(err.item())
            ret["mean_ok"] = abs(part[0].item() / part[1].item() - np.abs(P.thomas(a, b, c, d)).mean()) < 1e-12
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_solve_over_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["err"] < 1e-12 and ret["mean_ok"]


def test_one_pass_completion_single_process():
    """The algebra of the one-pass slab-coupled sweep and of the CTA-pair tiles: every slab solves its rows once with two
    spike columns, all slabs see all first / last row coefficients, each solves the interface for itself and completes
    x = y - p x_left - q x_right - equal to the undivided Thomas solve; two parts: the closed-form 2x2 interface."""
    a, b, c, d = _system(29, 128, 3)
    ref = P.thomas(a, b, c, d)
    for nsl in (2, 3, 8):
        slabs = P.split_even(128, nsl)
        parts = [P.slab_open_solve(a[:, x0:x0 + nx], b[:, x0:x0 + nx], c[:, x0:x0 + nx], d[:, x0:x0 + nx]) for x0, nx in slabs]
        f, pf, qf = (np.stack([pt[i][:, 0] for pt in parts]) for i in range(3))
        l, pl, ql = (np.stack([pt[i][:, -1] for pt in parts]) for i in range(3))
        xl, xr = P.interface_solve(f, pf, qf, l, pl, ql)
        for r, (x0, nx) in enumerate(slabs):
            y, p, q = parts[r]
            x = y - p * xl[r][:, None] - q * xr[r][:, None]
            assert np.abs(x - ref[:, x0:x0 + nx]).max() < 1e-12
        if nsl == 2:
            L0, F1 = P.pair_interface(l[0], ql[0], f[1], pf[1])
            assert np.abs(L0 - xl[1]).max() < 1e-13 and np.abs(F1 - xr[0]).max() < 1e-13


def _worker_one_pass(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines, n = 33, 48 * world
        a, b, c, d = _system(lines, n, 11)
        x0, nx = P.split_even(n, world)[rank]
        sl = slice(x0, x0 + nx)
        y, p, q = P.slab_open_solve(a[:, sl], b[:, sl], c[:, sl], d[:, sl])
        mine = torch.from_numpy(np.stack([y[:, 0], p[:, 0], q[:, 0], y[:, -1], p[:, -1], q[:, -1]]).copy())       # (6, lines)
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)                         # every slab's coefficients go to every slab (the kernel: peer stores)
        coef = np.stack([t.numpy() for t in everyone])          # (P, 6, lines)
        xl, xr = P.interface_solve(*(coef[:, i, :] for i in range(6)))
        x = y - p * xl[rank][:, None] - q * xr[rank][:, None]   # finished from what the slab already holds: nothing is read twice
        err = torch.tensor([float(np.abs(x - P.thomas(a, b, c, d)[:, sl]).max())], dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            ret["err"] = float(err.item())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_one_pass_completion_over_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_one_pass, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["err"] < 1e-12


def test_exact_chain_is_bit_identical_with_the_undivided_recurrence():
    a, b, c, d = _system(23, 90, 5)
    ref = P.thomas(a, b, c, d)
    for slabs in (P.split_even(90, 2), P.split_even(90, 4), [(0, 8), (8, 50), (58, 32)]):
        assert np.array_equal(P.thomas_chain(a, b, c, d, slabs), ref)
