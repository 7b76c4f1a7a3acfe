#!/usr/bin/env python
"""Multi-GPU (one process per GPU, NCCL) check of the slab-decomposed ADI step:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank also runs the SAME case on its own GPU as a single slab and compares its planes (tolerance of the fast
mode: fp64 1e-10 / fp32 1e-5), the residual and the GetLayer output on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from cmc_fluid_solver_b200 import AdiSolver3D  # noqa: E402
from cmc_fluid_solver_b200.cases import channel_case  # noqa: E402
from cmc_fluid_solver_b200.solver import nccl_unique_id  # noqa: E402
from conftest import layer_errors  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    def fresh_id():          # one ncclUniqueId per communicator
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    ok = True
    for fp, tol in ((8, 1e-10), (4, 1e-5)):
        nid = fresh_id()
        case = channel_case(int(os.environ.get("DIST_CHECK_PLANES", "32")) * world, 40, 48, fp_bytes=fp, depth_var=0.25)
        case.outdims = (9, 7, 5)
        one = AdiSolver3D().Init(case, device=lr, mode="fast"); one.CreateSegments()
        many = AdiSolver3D().Init(case, device=lr, mode="fast", rank=rank, nranks=world, nccl_id=nid); many.CreateSegments()
        assert [many.numSegs(d) for d in range(3)] == [one.numSegs(d) for d in range(3)], "segment counts differ"
        if rank == 0:
            print(f"exchange {many.exchange_kind()} kernel_x {many.get_option('kernel_x')}", flush=True)
        for i in range(4):
            one.UpdateBoundaries(); many.UpdateBoundaries()
            e1 = one.TimeStep(case.dt, 4, 2, True)
            e = many.TimeStep(case.dt, 4, 2, True)
            assert abs(e - e1) <= (1e-5 if fp == 4 else 1e-9) * abs(e1), (e, e1)
        v1, T1 = one.GetLayer(*case.outdims)
        v, T = many.GetLayer(*case.outdims)
        if rank == 0:
            assert np.allclose(v, v1, rtol=0, atol=tol * 1e5) and np.allclose(T, T1, rtol=0, atol=tol * 1e5), "GetLayer differs"
        # distributed output: every rank receives the output rows of its own slab, compactly
        many.set_option("local_output", 1)
        lo, hi = many.output_rows(case.outdims[0])
        rown = case.outdims[1] * case.outdims[2]
        vl = np.full((max(hi - lo, 1) * rown, 3), -7.0, dtype=v1.dtype); Tl = np.full(max(hi - lo, 1) * rown, -7.0)
        many.GetLayer(*case.outdims, vel=vl, T=Tl)
        many.set_option("local_output", 0)
        if hi > lo:
            assert np.allclose(vl[: (hi - lo) * rown], v1[lo * rown:hi * rown], rtol=0, atol=tol * 1e5), "local GetLayer (velocity) differs"
            assert np.allclose(Tl[: (hi - lo) * rown], T1[lo * rown:hi * rown], rtol=0, atol=tol * 1e5), "local GetLayer (T) differs"
        rows = torch.tensor([hi - lo], device="cuda"); dist.all_reduce(rows)
        assert int(rows.item()) == case.outdims[0], "the ranks' output rows do not cover the output"
        s1, sN = one.field_sums(0), many.field_sums(0)
        for n in "uvwT":
            assert abs(s1[n][0] - sN[n][0]) <= tol * max(abs(s1[n][1]) ** 0.5, 1.0) * 1e3, ("checksum", n, s1[n], sN[n])
        sl = slice(many.x0, many.x0 + many.nx)
        errs = layer_errors([one.read_field(0, q)[sl] for q in range(4)], [many.read_field(0, q) for q in range(4)])
        good = max(errs) <= tol
        ok &= good
        print(f"rank {rank}/{world} fp{fp * 8}: planes [{many.x0},{many.x0 + many.nx}) err {e:.6e} (1 GPU {e1:.6e}) "
              f"linf_vel {errs[0]:.2e} l2_vel {errs[1]:.2e} linf_T {errs[2]:.2e} l2_T {errs[3]:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        # the same run from slab-local node arrays (cmc_adi3d_set_nodes_slab): no rank builds the whole grid
        from cmc_fluid_solver_b200.solver import default_split, slab_window
        x0, nx = default_split(case.dimx, case.dimy, case.dimz, world)[rank]
        local = channel_case(case.dimx, 40, 48, fp_bytes=fp, depth_var=0.25, x_range=slab_window(case.dimx, x0, nx))
        loc = AdiSolver3D().Init(local, device=lr, mode="fast", rank=rank, nranks=world, nccl_id=fresh_id()); loc.CreateSegments()
        assert [loc.numSegs(d) for d in range(3)] == [one.numSegs(d) for d in range(3)], "segment counts differ (slab-local)"
        for i in range(4):
            loc.UpdateBoundaries()
            e2 = loc.TimeStep(case.dt, 4, 2, True)
        assert e2 == e, ("slab-local run differs from the whole-grid one", e2, e)
        for q in range(4):
            assert np.array_equal(loc.read_field(0, q), many.read_field(0, q)), f"slab-local field {q} differs"
        loc.close()
        one.close(); many.close()
        # bit-exact mode across ranks: the Thomas recurrence of an x-line runs through the ranks as a chain - every plane equal,
        # bit for bit, to the same case solved on one GPU (which equals the reference CPU solver)
        one_x = AdiSolver3D().Init(case, device=lr, mode="exact"); one_x.CreateSegments()
        many_x = AdiSolver3D().Init(case, device=lr, mode="exact", rank=rank, nranks=world, nccl_id=fresh_id()); many_x.CreateSegments()
        for i in range(3):
            one_x.UpdateBoundaries(); many_x.UpdateBoundaries()
            ex1 = one_x.TimeStep(case.dt, 4, 2, True)
            exn = many_x.TimeStep(case.dt, 4, 2, True)
        same = all(np.array_equal(one_x.read_field(0, q)[sl], many_x.read_field(0, q)) for q in range(4))
        ok &= same and abs(exn - ex1) <= (1e-12 if fp == 8 else 1e-6) * abs(ex1)
        print(f"rank {rank}/{world} fp{fp * 8}: exact mode, planes bit-identical with the one-GPU run: {same}; residual {exn:.15e} vs {ex1:.15e}", flush=True)
        one_x.close(); many_x.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST CHECK", "PASSED" if int(t.item()) else "FAILED")
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
