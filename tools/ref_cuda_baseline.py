#!/usr/bin/env python
"""Second baseline ("the reference's own CUDA path rebuilt for the GPU at hand"): the reference's CUDA backend -
AdiSolver3D.cu / TimeLayer3D.cu, unmodified, compiled by oracle/build_ref.sh with `nvcc -arch=sm_100`, its stock
configuration (MGPU_EMU 1, `GPU 1`, no transpose / decompose options) - timed on the same B200 on the masked channel at
128^3, 256^3 and 512^3 in fp64, next to the reference's CPU/OpenMP solver (residual cross-check at 128^3) and this repo's
solver behind the same driver.  Timing and residual only.

    python tools/ref_cuda_baseline.py [sizes...]   > profiles/rNN_reference_cuda_backend.log
"""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case  # noqa: E402

REF = ROOT / "oracle" / "_ref"
sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512]


def run(binary, data, cfg, steps, solver, td, timeout=1500):
    try:
        r = subprocess.run([str(binary), str(data), str(cfg), "-", str(steps), "align", "dump=none", f"solver={solver}"],
                           capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, "timeout"
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("probe:")]
    if r.returncode != 0:
        return None, (r.stdout[-800:] + r.stderr[-800:])
    return lines, ""


for n in sizes:
    with tempfile.TemporaryDirectory() as td:
        data, cfg = write_shape2d_case(td, f"c{n}", outline=BAFFLE_OUTLINE, grid_d=1.1 / n, depth=1.09, depth_var=0.2, time_steps=100,
                                       num_global=4, num_local=2, out_grid=(32, 32, 32))
        steps = 12 if n <= 256 else 4
        print(f"=== {n}^3 masked channel, fp64, num_global 4 num_local 2, {steps} steps ===")
        todo = [("reference CUDA backend (AdiSolver3D.cu, sm_100)", REF / "ref_probe3d_f64", "refgpu"),
                ("this repo, fast mode, behind the same driver", REF / "dropin3d_f64", "b200")]
        if n <= 128:
            todo.append(("reference CPU/OpenMP", REF / "ref_probe3d_f64", "cpu"))
        for name, binary, solver in todo:
            lines, err = run(binary, data, cfg, steps, solver, td)
            if lines is None:
                print(f"{name}: FAILED: {err}")
                continue
            print(f"{name}:")
            for ln in (lines if len(lines) <= 7 else lines[:2] + lines[-5:]):
                print("   ", ln)
