"""2D ADI (BASELINE config 1 = data/2D/box_pipe on the reference CPU solver): the C restatement
oracle/adi2d_oracle.c pinned against the REAL reference 2D solver.
 * golden vectors produced by oracle/_ref/ref_probe2d_f32 on the reference's data/2D/box_pipe case
   (tests/golden/box_pipe2d_f32.npz, generator tests/golden/make_golden2d.py): every step bit-for-bit;
 * the reference binary itself, when it and /root/reference are present (this container), all 49 steps;
 * the informal known answers of SURVEY 8(c): grid 120 x 135, dt 0.0007, 49 steps.
CPU only."""
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN


def load_golden2d():
    z = np.load(GOLDEN / "box_pipe2d_f32.npz")
    return {k: z[k] for k in z.files}


def golden_grid(z, step):
    """Grid2D::GetType / GetData arrays of one step (the solver reads the data only outside NODE_IN)."""
    n = int(z["dims"][0]) * int(z["dims"][1])
    out = []
    for q in range(3):
        a = np.zeros(n, dtype=np.float32)
        a[z["nonin"]] = z["grid_sparse"][step, q]
        out.append(a)
    return z["type"].astype(np.int32), z["bc"].astype(np.int32), out[0], out[1], out[2]


def drive2d(solver, z, steps, check=None):
    """FluidSolver2D.cpp:96-152: per step grid.Prepare(t) [= the stored grid arrays], UpdateBoundaries, TimeStep."""
    errs = []
    for s in range(steps):
        solver.set_grid(*golden_grid(z, s))
        solver.update_boundaries()
        errs.append(solver.time_step(float(z["dt"]), int(z["iters"][0]), int(z["iters"][1])))
        if check:
            check(s)
    return errs


def test_oracle2d_matches_golden_bitwise(oracle_mod):
    O = oracle_mod
    z = load_golden2d()
    dimx, dimy = (int(v) for v in z["dims"])
    assert (dimx, dimy) == (120, 135) and abs(float(z["dt"]) - 0.0007) < 1e-9 and int(z["steps"]) == 49
    o = O.Oracle2D(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
    for q in range(3):
        o.field(0, q)[:] = z["layer_init"][q]
    keep = {int(s): i for i, s in enumerate(z["keep"])}
    outs = {int(s): i for i, s in enumerate(z["out_steps"])}

    def check(s):
        for q in range(3):
            f = o.field(0, q)
            assert float(np.sum(f.astype(np.float64))) == z["sums"][s, q]
            if s in keep:
                assert np.array_equal(f, z["layers"][keep[s], q]), f"step {s} field {q}"
        if s in outs:
            vel, T = o.get_layer(*[int(v) for v in z["outdims"]])
            assert np.array_equal(vel, z["out_vel"][outs[s]]) and np.array_equal(T, z["out_T"][outs[s]])

    errs = drive2d(o, z, int(z["steps"]), check)
    assert np.array_equal(np.array(errs), z["err"])
    o.close()


def test_oracle2d_matches_reference_binary(oracle_mod, tmp_path):
    O = oracle_mod
    ref = Path("/root/reference/data/2D/box_pipe")
    if not (O.ref2d_binary().exists() and ref.is_dir()):
        pytest.skip("oracle/_ref/ref_probe2d_f32 or /root/reference not present")
    (tmp_path / "data.txt").write_bytes((ref / "box_pipe_data.txt").read_bytes().replace(b"\r", b""))
    cfg = (ref / "box_pipe_config.txt").read_bytes().replace(b"\r", b"").decode()
    (tmp_path / "config.txt").write_text("\n".join("solver\t\tADI" if ln.startswith("solver") else ln for ln in cfg.splitlines()) + "\n")
    O.run_ref2d(tmp_path / "data.txt", tmp_path / "config.txt", tmp_path / "out.bin", 0, "every")
    d = O.read_probe2d(tmp_path / "out.bin")
    o = O.Oracle2D(d["dimx"], d["dimy"], d["dx"], d["dy"], d["v_T"], d["v_vis"], d["t_vis"], d["t_phi"], d["startT"], d["fp_bytes"])
    for q in range(3):
        o.field(0, q)[:] = d["layers"][-1][q]
    assert len(d["grids"]) == 49
    for s in range(len(d["grids"])):
        g = d["grids"][s]
        o.set_grid(g["type"], g["bc"], g["vx"], g["vy"], g["T"])
        o.update_boundaries()
        assert o.time_step(d["dt"], d["num_global"], d["num_local"]) == d["errs"][s]
        for q in range(3):
            assert np.array_equal(o.field(0, q), d["layers"][s][q]), f"step {s} field {q}"
        if s in d["outputs"]:
            vel, T = o.get_layer(*d["outdims"])
            assert np.array_equal(vel, d["outputs"][s][0]) and np.array_equal(T, d["outputs"][s][1])
    o.close()


def test_segments_2d(oracle_mod):
    """AdiSolver2D::CreateSegments (AdiSolver2D.cpp:228-277): one segment per row / column that has fluid."""
    O = oracle_mod
    z = load_golden2d()
    dimx, dimy = (int(v) for v in z["dims"])
    o = O.Oracle2D(dimx, dimy, 1, 1, 1, 1, 1, 1, 1, 4)
    o.set_grid(*golden_grid(z, 0))
    f = o._fn("oracle2d_num_segments")
    import ctypes as C
    f.argtypes = [C.c_void_p, C.c_int]
    f.restype = C.c_int
    ty = z["type"].reshape(dimx, dimy)
    assert f(o.h, 0) == int(np.sum((ty == 0).any(axis=1))) and f(o.h, 1) == int(np.sum((ty == 0).any(axis=0)))
    o.close()


def drive_dynamic(solver, z, read):
    """The moving-boundary golden (reference data/2D/heart_MR): the full grid arrays change every step."""
    errs = []
    for s in range(int(z["steps"])):
        solver.set_grid(z["type"][s].astype(np.int32), z["bc"][s].astype(np.int32), z["gvx"][s], z["gvy"][s], z["gT"][s])
        solver.update_boundaries()
        errs.append(solver.time_step(float(z["dt"]), int(z["iters"][0]), int(z["iters"][1])))
        for q in range(3):
            assert float(np.sum(read(q).astype(np.float64))) == z["sums"][s, q], f"step {s} field {q}"
    assert np.array_equal(np.array(errs), z["err"])
    for q in range(3):
        assert np.array_equal(read(q), z["layer_last"][q])


def test_oracle2d_moving_boundaries_golden(oracle_mod):
    O = oracle_mod
    z = np.load(GOLDEN / "heart_mr2d_f32.npz")
    dimx, dimy = (int(v) for v in z["dims"])
    assert any(not np.array_equal(z["type"][s], z["type"][s - 1]) for s in range(1, int(z["steps"])))     # the mask really moves
    o = O.Oracle2D(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
    for q in range(3):
        o.field(0, q)[:] = z["layer_init"][q]
    drive_dynamic(o, z, lambda q: o.field(0, q))
    o.close()


@pytest.mark.parametrize("name", ["heart_MR", "heart_US"])
def test_oracle2d_moving_boundaries_reference_binary(oracle_mod, tmp_path, name):
    """120 steps of the reference's moving-boundary cases (25 / 80 frames): every layer and residual bit-for-bit."""
    O = oracle_mod
    ref = Path("/root/reference/data/2D") / name
    if not (O.ref2d_binary().exists() and ref.is_dir()):
        pytest.skip("oracle/_ref/ref_probe2d_f32 or /root/reference not present")
    (tmp_path / "data.txt").write_bytes((ref / f"{name}_data.txt").read_bytes().replace(b"\r", b""))
    cfg = (ref / f"{name}_config.txt").read_bytes().replace(b"\r", b"").decode()
    (tmp_path / "config.txt").write_text("\n".join("solver\t\tADI" if ln.startswith("solver") else ln for ln in cfg.splitlines()) + "\n")
    O.run_ref2d(tmp_path / "data.txt", tmp_path / "config.txt", tmp_path / "out.bin", 120, "every")
    d = O.read_probe2d(tmp_path / "out.bin")
    o = O.Oracle2D(d["dimx"], d["dimy"], d["dx"], d["dy"], d["v_T"], d["v_vis"], d["t_vis"], d["t_phi"], d["startT"], d["fp_bytes"])
    for q in range(3):
        o.field(0, q)[:] = d["layers"][-1][q]
    changed = 0
    for s in range(len(d["grids"])):
        g = d["grids"][s]
        changed += s > 0 and not np.array_equal(g["type"], d["grids"][s - 1]["type"])
        o.set_grid(g["type"], g["bc"], g["vx"], g["vy"], g["T"])
        o.update_boundaries()
        assert o.time_step(d["dt"], d["num_global"], d["num_local"]) == d["errs"][s]
        for q in range(3):
            assert np.array_equal(o.field(0, q), d["layers"][s][q]), f"{name} step {s} field {q}"
    assert changed > 50
    o.close()
