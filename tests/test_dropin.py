"""Drop-in boundary end to end (GPU): the reference's OWN loader / Config / Grid3D objects + the Solver3D adapter
(cmc_fluid_solver_b200/host/B200AdiSolver3D.*) + libcmcadi.so, built by oracle/build_ref.sh as oracle/_ref/dropin3d_*,
against the reference's CPU solver (oracle/_ref/ref_probe3d_*) on the same case files.  Both binaries are the
driver of oracle/ref_probe3d.cpp, which follows FluidSolver3D.cpp:53-286 and dumps raw time layers.
exact mode: bit-identical; fast mode: within the stated tolerance (fp64 1e-10, fp32 1e-5)."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT, layer_errors
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case

REF = ROOT / "oracle" / "_ref"


def _run(binary, data, cfg, out, steps, solver, align=True):
    cmd = [str(binary), str(data), str(cfg), str(out), str(steps)] + (["align"] if align else []) + ["dump=every", "getlayer", f"solver={solver}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("fp", [8, 4])
def test_reference_driver_with_b200_solver(oracle_mod, tmp_path, fp):
    O = oracle_mod
    tag = "f32" if fp == 4 else "f64"
    dropin, probe = REF / f"dropin3d_{tag}", REF / f"ref_probe3d_{tag}"
    if not (dropin.exists() and probe.exists()):
        pytest.skip("oracle/_ref/dropin3d_* not built (needs /root/reference at build time)")
    data, cfg = write_shape2d_case(tmp_path, "case", outline=BAFFLE_OUTLINE, grid_d=0.02, depth_var=0.2, time_steps=300,
                                   out_grid=(20, 21, 19))
    steps = 12
    _run(probe, data, cfg, tmp_path / "cpu.bin", steps, "cpu")
    ref = O.read_probe(tmp_path / "cpu.bin")
    rs = {(s["step"], s["kind"]): s for s in ref.snapshots}
    for solver, tol in (("b200exact", 0.0), ("b200", 1e-10 if fp == 8 else 1e-5)):
        _run(dropin, data, cfg, tmp_path / f"{solver}.bin", steps, solver)
        got = O.read_probe(tmp_path / f"{solver}.bin")
        assert got.shape == ref.shape and np.array_equal(got.type, ref.type)
        gs = {(s["step"], s["kind"]): s for s in got.snapshots}
        assert set(gs) == set(rs)
        for key, r in rs.items():
            g = gs[key]
            if key[1] == 0:          # raw current layer after the step
                if tol == 0.0:
                    # the fields are bit-identical; the residual is a sum over all cells (serial on the CPU, a block
                    # reduction on the device): equal up to summation order
                    assert abs(g["err"] - r["err"]) <= 1e-12 * abs(r["err"])
                    for n in "uvwT":
                        assert np.array_equal(g[n], r[n]), f"{solver} fp{fp * 8} step {key[0]} field {n}"
                else:
                    assert abs(g["err"] - r["err"]) <= 1e-6 * abs(r["err"]) + 1e-12
                    errs = layer_errors([r[n] for n in "uvwT"], [g[n] for n in "uvwT"])
                    assert max(errs) <= tol, (solver, key, errs)
            else:                    # GetLayer output (previous layer, OUT cells = 99999, downsampled)
                if tol == 0.0:
                    assert np.array_equal(g["vel"], r["vel"]) and np.array_equal(g["T"], r["T"])
                else:
                    assert np.array_equal(g["vel"] == 99999, r["vel"] == 99999)
                    m = r["T"] != 99999
                    assert np.allclose(g["T"][m], r["T"][m], rtol=0, atol=tol * max(1.0, float(np.abs(r["T"][m]).max())) * 10)
