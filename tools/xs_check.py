#!/usr/bin/env python
"""One-pass slab-coupled x-sweep (kernels_tma.cu XS) on ONE GPU: several slabs, each with its own stream, share device 0
(CMC_SHARE_DEVICE=1: the kernels of the slabs wait for each other's flags, so they are launched with a share of the SMs
each), compared with the CPU oracle and with the same case on one slab.  A development vehicle: the real configurations -
one process per GPU, one process driving several GPUs - are tests/test_gpu_dist.py and tools/dist_check.py.

    CMC_SHARE_DEVICE=1 python tools/xs_check.py [fp_bytes]
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("CMC_SHARE_DEVICE", "1")
os.environ.setdefault("CMC_XS", "1")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from conftest import component_errors, layer_errors  # noqa: E402
from cmc_fluid_solver_b200 import AdiSolver3D  # noqa: E402
from cmc_fluid_solver_b200.cases import channel_case  # noqa: E402
from cmc_fluid_solver_b200.solver import LAYER_CUR  # noqa: E402
from oracle import oracle as O  # noqa: E402

fp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tol = 1e-10 if fp == 8 else 1e-5
ok = True
for dims, nslabs in (((256, 40, 44), 2), ((512, 24, 24), 2), ((384, 40, 28), 3), ((256, 24, 40), 4), ((512, 40, 20), 8)):
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    xs = AdiSolver3D().Init(case, mode="fast", devices=[0] * nslabs); xs.CreateSegments()
    kind = xs.get_option("kernel_x")
    assert kind == 5, f"the one-pass x-sweep is not selected (kernel kind {kind})"
    for i in range(3):
        ora.update_boundaries(); one.UpdateBoundaries(); xs.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e1 = one.TimeStep(case.dt, case.num_global, case.num_local, True)
        e = xs.TimeStep(case.dt, case.num_global, case.num_local, True)
        ref = [ora.field(O.LAYER_CUR, q) for q in range(4)]
        got = [xs.read_field(LAYER_CUR, q) for q in range(4)]
        le = layer_errors(ref, got); ce = component_errors(ref, got)
        good = max(le) <= tol and max(max(c) for c in ce) <= 2 * tol and abs(e - e_ref) <= 10 * tol * abs(e_ref)
        ok &= good
        print(f"{dims} x{nslabs} fp{fp * 8} kernel_x {kind} step {i}: err {e:.12e} (oracle {e_ref:.12e}, one slab {e1:.12e}) "
              f"linf/l2 vel,T {tuple(f'{v:.1e}' for v in le)} {'ok' if good else 'FAILED'}", flush=True)
    xs.close(); one.close()
print("XS CHECK", "PASSED" if ok else "FAILED")
sys.exit(0 if ok else 1)
