"""2D ADI on the GPU through the C ABI (cmc_adi2d_*): bit-for-bit against the golden vectors of the
reference's 2D solver on its own data/2D/box_pipe case (all 49 steps: fields, residual, iteration counts via the
residual, GetLayer output), and against the CPU oracle in fp64 and on a synthetic case with a free outflow."""
import numpy as np
import pytest

from test_oracle2d import golden_grid, load_golden2d


@pytest.mark.gpu
def test_box_pipe_2d_matches_reference_bitwise():
    from cmc_fluid_solver_b200 import AdiSolver2D
    z = load_golden2d()
    dimx, dimy = (int(v) for v in z["dims"])
    s = AdiSolver2D().Init(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
    for q in range(3):
        s.write_field(0, q, z["layer_init"][q])
    keep = {int(k): i for i, k in enumerate(z["keep"])}
    outs = {int(k): i for i, k in enumerate(z["out_steps"])}
    for step in range(int(z["steps"])):
        s.set_grid(*golden_grid(z, step))
        s.UpdateBoundaries()
        err = s.TimeStep(float(z["dt"]), int(z["iters"][0]), int(z["iters"][1]))
        assert err == z["err"][step], f"residual differs at step {step}"
        for q in range(3):
            f = s.read_field(0, q)
            assert float(np.sum(f.astype(np.float64))) == z["sums"][step, q], f"step {step} field {q}"
            if step in keep:
                assert np.array_equal(f, z["layers"][keep[step], q])
        if step in outs:
            vel, T = s.GetLayer(*[int(v) for v in z["outdims"]])
            assert np.array_equal(vel, z["out_vel"][outs[step]]) and np.array_equal(T, z["out_T"][outs[step]])
    assert s.launch_count() >= 2 * int(z["steps"])
    s.close()


def synthetic_2d(dimx, dimy, free_outflow=True):
    """Channel with an inflow valve on the left, (free) outflow on the right, a wall-attached block; OUT rim."""
    ty = np.ones((dimx, dimy), dtype=np.int32)          # NODE_OUT
    ty[2:dimx - 2, 2:dimy - 2] = 0                      # NODE_IN
    ty[1, 1:dimy - 1] = 3; ty[dimx - 2, 1:dimy - 1] = 3                 # valves
    ty[1:dimx - 1, 1] = 2; ty[1:dimx - 1, dimy - 2] = 2                 # walls
    ty[dimx // 3:dimx // 3 + 6, 2:dimy // 2] = 2                        # block attached to the lower wall
    bc = np.zeros((dimx, dimy), dtype=np.int32)
    vx = np.zeros((dimx, dimy)); vy = np.zeros((dimx, dimy)); T = np.ones((dimx, dimy))
    vx[1, 2:dimy - 2] = 1.0
    if free_outflow:
        bc[dimx - 2, 2:dimy - 2] = 1
    return ty.ravel(), bc.ravel(), vx.ravel(), vy.ravel(), T.ravel()


@pytest.mark.gpu
@pytest.mark.parametrize("fp", [4, 8])
def test_synthetic_2d_against_oracle(oracle_mod, fp):
    from cmc_fluid_solver_b200 import AdiSolver2D
    O = oracle_mod
    dimx, dimy, h = 61, 47, 0.02
    par = (1.0, 0.05, 0.07, 0.002)
    ft = np.float32 if fp == 4 else np.float64
    g = synthetic_2d(dimx, dimy)
    o = O.Oracle2D(dimx, dimy, h, h, *par, 1.0, fp)
    s = AdiSolver2D().Init(dimx, dimy, h, h, *par, 1.0, fp)
    for solver in (o, s):
        solver.set_grid(*g)
        solver.init_layer()
    for step in range(6):
        e_ref = (o.update_boundaries(), o.time_step(0.05, 3, 2))[1]
        e = (s.UpdateBoundaries(), s.TimeStep(0.05, 3, 2))[1]
        assert e == e_ref and s.iters == o.iters()
        for q in range(3):
            assert np.array_equal(s.read_field(0, q), o.field(0, q).astype(ft)), f"fp{fp * 8} step {step} field {q}"
    v, T = s.GetLayer(10, 9)
    vo, To = o.get_layer(10, 9)
    assert np.array_equal(v, vo) and np.array_equal(T, To)
    o.close(); s.close()


@pytest.mark.gpu
def test_2d_argument_errors():
    from cmc_fluid_solver_b200 import AdiSolver2D, CmcError
    with pytest.raises(CmcError):
        AdiSolver2D().Init(2, 10, 1, 1, 1, 1, 1, 1, 1)
    s = AdiSolver2D().Init(8, 8, 1, 1, 1, 1, 1, 1, 1)
    with pytest.raises(CmcError):
        s.TimeStep(0.1, 1, 1)           # no grid yet
    s.close()


def _write_2d_case(directory):
    """A 2D case in the reference's own file formats (Grid2D::LoadFromFile, src/FluidSolver2D/Grid2D.cpp:268-372;
    Config.h:203-245): the masked channel outline at the scale of the reference's data/2D/box_pipe."""
    from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE
    lines = ["1", "0.035", str(len(BAFFLE_OUTLINE))]
    for kind, pts, vel in BAFFLE_OUTLINE:
        lines.append(str(len(pts)))
        lines += [f"{x * 0.08 + 90} {y * 0.08 + 150}" for x, y in pts]
        lines.append(kind)
        if kind == "Motion":
            lines.append(f"{vel[0] * 0.1} {vel[1] * 0.1}")
    (directory / "data.txt").write_text("\n".join(lines) + "\n")
    (directory / "config.txt").write_text(
        "dimension\t2D\nviscosity\t0.05\ndensity\t1000.0\nbc_type\tNoSlip\nbc_strenght\t0.5\ngrid_dx\t0.0008\ngrid_dy\t0.0008\n"
        "cycles\t1\ntime_steps\t40\nout_time_steps\t10\nout_gridx\t30\nout_gridy\t25\nout_fmt\tNetCDF\nsolver\tADI\nnum_global\t2\nnum_local\t1\n")
    return directory / "data.txt", directory / "config.txt"


@pytest.mark.gpu
def test_reference_2d_driver_with_b200_solver(oracle_mod, tmp_path):
    """Drop-in boundary of the 2D solver end to end: the reference's own Grid2D / Config + the Solver2D adapter
    (cmc_fluid_solver_b200/host/B200AdiSolver2D.*) + libcmcadi.so (oracle/_ref/dropin2d_f32, built by
    oracle/build_ref.sh) against the reference CPU solver (oracle/_ref/ref_probe2d_f32) on the same case files:
    bit-identical layers, residuals and GetLayer outputs over all steps."""
    import subprocess
    from conftest import ROOT
    O = oracle_mod
    ref_bin, drop_bin = ROOT / "oracle" / "_ref" / "ref_probe2d_f32", ROOT / "oracle" / "_ref" / "dropin2d_f32"
    if not (ref_bin.exists() and drop_bin.exists()):
        pytest.skip("oracle/_ref/dropin2d_f32 not built (needs /root/reference at build time)")
    data, cfg = _write_2d_case(tmp_path)
    for binary, out, solver in ((ref_bin, "cpu.bin", "cpu"), (drop_bin, "gpu.bin", "b200")):
        r = subprocess.run([str(binary), str(data), str(cfg), str(tmp_path / out), "0", "dump=every", f"solver={solver}"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    a, b = O.read_probe2d(tmp_path / "cpu.bin"), O.read_probe2d(tmp_path / "gpu.bin")
    assert (a["dimx"], a["dimy"]) == (b["dimx"], b["dimy"]) and len(a["layers"]) == len(b["layers"]) == 40     # initial + 39 steps
    for s in a["layers"]:
        assert a["errs"][s] == b["errs"][s], f"residual differs at step {s}"
        for q in range(3):
            assert np.array_equal(a["layers"][s][q], b["layers"][s][q]), f"step {s} field {q}"
    assert set(a["outputs"]) == set(b["outputs"]) and len(a["outputs"]) >= 3
    for s in a["outputs"]:
        assert np.array_equal(a["outputs"][s][0], b["outputs"][s][0]) and np.array_equal(a["outputs"][s][1], b["outputs"][s][1])
    for s in a["grids"]:        # SetGridBoundaries fed the same velocities back into the reference's grid
        for k in ("type", "bc", "vx", "vy", "T"):
            assert np.array_equal(a["grids"][s][k], b["grids"][s][k]), f"grid {k} differs at step {s}"


@pytest.mark.gpu
def test_moving_boundaries_2d_matches_reference_bitwise():
    """Golden vectors of the reference's moving-boundary case data/2D/heart_MR (node types change every step: the
    segments are rebuilt on the device inside every TimeStep, like AdiSolver2D::CreateSegments)."""
    from cmc_fluid_solver_b200 import AdiSolver2D
    from conftest import GOLDEN
    from test_oracle2d import drive_dynamic
    z = np.load(GOLDEN / "heart_mr2d_f32.npz")
    dimx, dimy = (int(v) for v in z["dims"])
    s = AdiSolver2D().Init(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
    for q in range(3):
        s.write_field(0, q, z["layer_init"][q])
    drive_dynamic(s, z, lambda q: s.read_field(0, q))
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fp", [4, 8])
def test_batched_2d_cases_equal_single_steps(oracle_mod, fp):
    """cmc_adi2d_time_step_batch: several independent cases (different inflow speeds, so different outer-iteration counts)
    advance in one launch - bit-identical with stepping every case on its own, and with the oracle."""
    from cmc_fluid_solver_b200 import AdiSolver2D
    from cmc_fluid_solver_b200.solver import time_step_batch_2d
    O = oracle_mod
    dimx, dimy, h = 61, 47, 0.02
    par = (1.0, 0.05, 0.07, 0.002)
    ft = np.float32 if fp == 4 else np.float64
    speeds = [0.5, 1.0, 1.5, 0.8, 1.2, 0.3, 0.9]
    batch, single, oracles = [], [], []
    for v in speeds:
        ty, bc, vx, vy, T = synthetic_2d(dimx, dimy)
        vx = vx * v
        for lst in (batch, single):
            s = AdiSolver2D().Init(dimx, dimy, h, h, *par, 1.0, fp)
            s.set_grid(ty, bc, vx, vy, T); s.init_layer()
            lst.append(s)
        o = O.Oracle2D(dimx, dimy, h, h, *par, 1.0, fp)
        o.set_grid(ty, bc, vx, vy, T); o.init_layer()
        oracles.append(o)
    for step in range(4):
        errs, iters = time_step_batch_2d(batch, 0.05, 3, 2, update_boundaries=True)
        for k, (s, o) in enumerate(zip(single, oracles)):
            s.UpdateBoundaries(); e = s.TimeStep(0.05, 3, 2)
            o.update_boundaries(); e_ref = o.time_step(0.05, 3, 2)
            assert errs[k] == e == e_ref and iters[k] == s.iters == o.iters(), (step, k)
            for q in range(3):
                assert np.array_equal(batch[k].read_field(0, q), s.read_field(0, q))
                assert np.array_equal(batch[k].read_field(0, q), o.field(0, q).astype(ft))
    assert len(set(iters)) >= 1
    for s in batch + single:
        s.close()
    for o in oracles:
        o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fp", [4, 8])
def test_step_host_round_trip_equals_separate_calls(fp):
    """cmc_adi2d_step_host (what the Solver2D adapter calls every step): grid + both host layers up in one copy, TimeStep, both
    layers down in one copy - bit-identical with set_grid + write_field x6 + TimeStep + read_field x6."""
    from cmc_fluid_solver_b200 import AdiSolver2D
    dimx, dimy, h = 61, 47, 0.02
    par = (1.0, 0.05, 0.07, 0.002)
    ft = np.float32 if fp == 4 else np.float64
    g = synthetic_2d(dimx, dimy)
    a = AdiSolver2D().Init(dimx, dimy, h, h, *par, 1.0, fp)
    b = AdiSolver2D().Init(dimx, dimy, h, h, *par, 1.0, fp)
    a.set_grid(*g); a.init_layer()
    cur = [np.ascontiguousarray(np.asarray(g[2 + q], dtype=ft).reshape(-1)) for q in range(3)]
    nxt = [np.zeros(dimx * dimy, dtype=ft) for _ in range(3)]
    for step in range(4):
        vx = np.asarray(g[2], dtype=ft) * (1.0 + 0.1 * step)          # the driver changes the grid between steps
        a.set_grid(g[0], g[1], vx, g[3], g[4])
        for q in range(3):                                            # a host-side edit of both layers, as Solver2D::UpdateBoundaries does
            cur[q][7 * dimy + 5] += ft(0.01); nxt[q][9 * dimy + 3] -= ft(0.02)
            a.write_field(0, q, cur[q]); a.write_field(2, q, nxt[q])
        e_a = a.TimeStep(0.05, 3, 2)
        e_b = b.StepHost(g[0], g[1], vx, g[3], g[4], cur, nxt, 0.05, 3, 2)
        assert e_a == e_b and a.iters == b.iters
        for q in range(3):
            assert np.array_equal(a.read_field(0, q).reshape(-1), cur[q]), (step, q)
            assert np.array_equal(a.read_field(2, q).reshape(-1), nxt[q]), (step, q)
    a.close(); b.close()
