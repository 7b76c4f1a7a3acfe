#!/usr/bin/env python
"""BASELINE config 4 against the reference itself, at the stated size: the 512^3 masked channel (case files in the
reference's own format, read by its own loader, `align`) stepped by

  * the reference CPU/OpenMP solver (oracle/_ref/ref_probe3d_f64, unmodified sources) on the GPU box's host cores, and
  * the reference loader + Solver3D adapter + libcmcadi.so (oracle/_ref/dropin3d_f64), fast mode and exact mode,

and compared through the probe's compact records (per-field sums / sums of squares and the strided subsample
[::32, ::32, ::32] of the layer after every step).  Needs ~35 GB of host memory and about a minute per reference step on
16 cores, so it is run once per round under gpurun and its output is committed under profiles/ instead of being a test:

    python tools/config4_check.py [steps]   > profiles/rNN_config4_512_vs_reference.log
"""
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case  # noqa: E402
from conftest import component_errors, layer_errors  # noqa: E402
from oracle import oracle as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
REF = ROOT / "oracle" / "_ref"
with tempfile.TemporaryDirectory() as td:
    data, cfg = write_shape2d_case(td, "c4", outline=BAFFLE_OUTLINE, grid_d=0.00215, depth=1.09, depth_var=0.2, time_steps=100,
                                   num_global=4, num_local=2, out_grid=(32, 32, 32))
    extra = ["align", "dump=list:" + ",".join(str(i) for i in range(steps)), "stats=32"]
    res = {}
    for name, binary, solver in (("reference CPU", REF / "ref_probe3d_f64", "cpu"), ("b200 fast", REF / "dropin3d_f64", "b200"),
                                 ("b200 exact", REF / "dropin3d_f64", "b200exact")):
        out = Path(td) / (solver + ".bin")
        t0 = time.time()
        r = subprocess.run([str(binary), str(data), str(cfg), str(out), str(steps)] + extra + [f"solver={solver}"], capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED", r.stdout[-1500:], r.stderr[-1500:]); sys.exit(1)
        case = O.read_probe(out)
        res[name] = [s for s in case.snapshots if s["kind"] == 5]
        head = [ln for ln in r.stdout.splitlines() if ln.startswith("probe:")][:1]
        tail = [ln for ln in r.stdout.splitlines() if "seconds" in ln][-1:]
        print(f"{name}: {time.time() - t0:.0f} s wall; {head} {tail}")
    ok = True
    for i in range(steps):
        ref = res["reference CPU"][i]
        for name, tol in (("b200 exact", 0.0), ("b200 fast", 1e-10)):
            got = res[name][i]
            if tol == 0.0:
                same = all(np.array_equal(a, b) for a, b in zip(ref["sample"], got["sample"])) and ref["sums"] == got["sums"] and ref["sumsq"] == got["sumsq"]
                print(f"step {i} {name}: subsample and per-field sums bit-identical with the reference: {same}; residual {got['err']:.12e} vs {ref['err']:.12e}")
                ok &= same
            else:
                le = layer_errors(ref["sample"], got["sample"]); ce = component_errors(ref["sample"], got["sample"])
                ds = [abs(a - b) / max(c, 1e-300) for a, b, c in zip(ref["sums"], got["sums"], ref["sumabs"])]
                good = max(le) <= tol and max(max(c) for c in ce) <= tol and max(ds) <= tol
                print(f"step {i} {name}: (linf_vel, l2_vel, linf_T, l2_T) = {tuple(f'{e:.2e}' for e in le)}; per-field (linf, l2) = "
                      f"{[tuple(f'{e:.1e}' for e in c) for c in ce]}; |sum - sum_ref| / sum|ref| = {[f'{d:.1e}' for d in ds]}; within 1e-10: {good}")
                ok &= good
    print("CONFIG 4 CHECK", "PASSED" if ok else "FAILED")
    sys.exit(0 if ok else 1)
