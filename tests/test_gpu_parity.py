"""Parity of the CUDA path (through the C ABI) with the oracle / the reference's golden vectors.

 * exact mode: bit-for-bit (every cell, fp32 and fp64);
 * fast mode : BASELINE.json north_star tolerances - fp64 1e-10, fp32 1e-5 (relative L-inf and L2 per field).
"""
import numpy as np
import pytest

from conftest import assert_fields_close, drive, layer_errors, load_golden, max_rel
from cmc_fluid_solver_b200 import AdiSolver3D, CmcError
from cmc_fluid_solver_b200.cases import Case, channel_case
from cmc_fluid_solver_b200.solver import (DIR_X, DIR_Y, DIR_Z, LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP,
                                          solve_tridiagonal_batch)

pytestmark = pytest.mark.gpu

TOL = {8: 1e-10, 4: 1e-5}


def _assert_close(ref4, got4, fp, what=""):
    """fp64 1e-10 / fp32 1e-5: velocity as a vector field, T, and every component u, v, w, T on its own (conftest)."""
    assert_fields_close(ref4, got4, fp, what)


def _check_fields(O, ora, sol, case, mode, what=""):
    ref = [ora.field(O.LAYER_CUR, q) for q in range(4)]
    got = [sol.read_field(LAYER_CUR, q) for q in range(4)]
    if mode == "exact":
        for q in range(4):
            assert np.array_equal(ref[q], got[q]), f"{what}: exact mode differs from the oracle in field {q}"
    else:
        _assert_close(ref, got, case.fp_bytes, what)


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("name", ["box32_f64", "box32_f32", "baffle32_f64", "baffle32_f32", "nupipe_f64", "nupipe_f32"])
def test_golden_vectors_from_the_reference(name, mode):
    case, exp = load_golden(name)
    s = AdiSolver3D().Init(case, mode=mode)
    s.CreateSegments()
    errs, layers = drive(s, case, exp["steps"])
    got = [s.read_field(LAYER_CUR, q).ravel() for q in range(4)]
    if mode == "exact":
        for q in range(4):
            assert np.array_equal(got[q], exp["last"][q])
    else:
        _assert_close(exp["last"], got, case.fp_bytes, name)
    assert np.allclose(errs, exp["err"], rtol=1e-6 if case.fp_bytes == 4 else 1e-9, atol=0)
    vel, T = layers[0]
    assert np.array_equal(vel, exp["layer0_vel"]) and np.array_equal(T, exp["layer0_T"])   # layer 0 = initial condition
    s.close()


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(40, 36, 32), (37, 29, 23), (64, 48, 80), (24, 130, 17)])
def test_steps_against_oracle(oracle_mod, dims, fp, mode):
    """Masked channel (baffle + bottom perturbation), ragged dims, 5 steps with GetLayer mutation in between."""
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    case.outdims = (7, 5, 6)
    ora = O.Oracle3D(case); ora.create_segments()
    s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
    assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(5):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        if i in (0, 3):
            v_ref, T_ref = ora.get_layer(*case.outdims)
            v, T = s.GetLayer(*case.outdims)
            if mode == "exact":
                assert np.array_equal(v, v_ref) and np.array_equal(T, T_ref)
            else:
                assert np.allclose(v, v_ref, rtol=0, atol=TOL[fp] * 1e5) and np.allclose(T, T_ref, rtol=0, atol=TOL[fp] * 1e5)
        _check_fields(O, ora, s, case, mode, f"step {i}")
    s.close()


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("d", [DIR_X, DIR_Y, DIR_Z])
def test_single_sweep_against_oracle(oracle_mod, d, fp, mode):
    """Component level: one SolveDirection (num_local x {all lines, merge}) from a developed flow state."""
    O = oracle_mod
    case = channel_case(48, 40, 56, fp_bytes=fp)
    ora = O.Oracle3D(case); ora.create_segments()
    for _ in range(2):
        ora.update_boundaries(); ora.time_step(case.dt, 2, 1, False)
    s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
    for slot in (LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP):
        for q in range(4):
            s.write_field(slot, q, ora.field(slot, q))
    ora.update_boundaries(); s.UpdateBoundaries()
    ora.step_prologue(); s.step_prologue()
    ora.solve_direction(d, case.dt, 2, O.LAYER_CUR, O.LAYER_TEMP, O.LAYER_NEXT)
    s.SolveDirection(d, case.dt, 2, LAYER_CUR, LAYER_NEXT)
    for slot in (LAYER_NEXT, LAYER_TEMP):
        ref = [ora.field(slot, q) for q in range(4)]
        got = [s.read_field(slot, q) for q in range(4)]
        if mode == "exact":
            for q in range(4):
                assert np.array_equal(ref[q], got[q]), (slot, q)
        else:
            _assert_close(ref, got, fp, f"layer {slot}")
    s.close()


def _line_case(line_types, bc_free_cells=(), fp=8):
    """3 x 3 x n grid whose centre z-line has the given node types."""
    n = len(line_types)
    t = np.full((3, 3, n), 1, dtype=np.int32)
    t[1, 1, :] = line_types
    bcv = np.zeros(t.shape, np.int32); bct = np.zeros(t.shape, np.int32)
    for k in bc_free_cells:
        bcv[1, 1, k] = 1; bct[1, 1, k] = 1
    ft = np.float32 if fp == 4 else np.float64
    vx = np.zeros(t.shape, ft); T = np.ones(t.shape, ft)
    vx[1, 1, :] = np.linspace(0.2, 1.0, n)
    z = np.zeros(t.size, ft)
    return Case(3, 3, n, .05, .05, .05, 1.0, 0.005, 0.007, 0.0014, 0.05, 2, 2, fp, type=t.ravel(), bc_vel=bcv.ravel(),
                bc_temp=bct.ravel(), vx=vx.ravel(), vy=z, vz=vx.ravel().copy(), T=T.ravel())


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_edge_case_lines(oracle_mod, mode):
    """A cell shared by two segments (NOSLIP and FREE boundary rows; FREE makes the fast solver fall back to the
    exact kernels), segments of the minimum length 3, valves inside the line."""
    O = oracle_mod
    line = [1, 2, 0, 2, 0, 0, 2, 1, 2, 0, 2, 1, 1, 2, 0, 0, 0, 3, 0, 0, 0, 0, 3, 0, 0, 2, 1, 1]
    O.set_threads(1)     # the reference's segment loop races on a shared FREE cell; one thread = list order
    for free in ((), (22,), (3, 22)):
        case = _line_case(line, free)
        ora = O.Oracle3D(case); ora.create_segments()
        s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
        assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
        for i in range(3):
            ora.update_boundaries(); s.UpdateBoundaries()
            ora.time_step(case.dt, 2, 2, False); s.TimeStep(case.dt, 2, 2, False)
        _check_fields(O, ora, s, case, mode, f"free={free}")
        s.close()
    O.set_threads(0)


def test_segment_counts_with_dropped_runs(oracle_mod):
    """GenerateListSegments corner cases that make the reference read out of bounds when stepped (IN cells on a
    domain face): only the line descriptors are compared - unterminated run dropped, IN cell at index 0 = start."""
    O = oracle_mod
    line = [0, 0, 2, 0, 0, 2, 1, 2, 0, 2, 1, 1, 2, 0, 0, 0]
    case = _line_case(line)
    ora = O.Oracle3D(case); ora.create_segments()
    s = AdiSolver3D().Init(case); s.CreateSegments()
    assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    assert s.numSegs(DIR_Z) == 3
    s.close()


def test_get_layer_returns_previous_layer_and_mutates_it(oracle_mod):
    O = oracle_mod
    case = channel_case(20, 18, 16, fp_bytes=8, baffle=False)
    s = AdiSolver3D().Init(case, mode="exact"); s.CreateSegments()
    s.UpdateBoundaries(); s.TimeStep(case.dt, 1, 1, True)
    vel, T = s.GetLayer()                                   # full resolution
    vel = vel.reshape(*case.shape, 3)
    out = case.type.reshape(case.shape) == 1
    assert np.all(vel[out] == 99999.0) and np.all(T.reshape(case.shape)[out] == 99999.0)
    # previous layer = initial condition: inflow valve u = 1, T = 1 in the fluid
    init_u = case.vx.reshape(case.shape)
    assert np.array_equal(vel[..., 0][~out], init_u[~out])
    assert np.all(s.read_field(LAYER_NEXT, 0)[out] == 99999.0)     # mutation is kept in `next`
    s.close()


def test_divergence_guard(oracle_mod):
    """dt far beyond the diagonal-dominance limit: residual > 0.01 -> CMC_ERR_DIVERGED, layers not swapped."""
    from cmc_fluid_solver_b200 import DivergedError
    case = channel_case(20, 18, 16, fp_bytes=8, baffle=False, inflow=400.0)
    s = AdiSolver3D().Init(case, mode="exact"); s.CreateSegments()
    s.UpdateBoundaries()
    before = s.read_field(LAYER_CUR, 0)
    with pytest.raises(DivergedError) as ei:
        for _ in range(20):
            s.TimeStep(5.0, 2, 1, True)
            before = s.read_field(LAYER_CUR, 0)
    assert "Error is too big!" in str(ei.value)
    assert np.array_equal(before, s.read_field(LAYER_CUR, 0))
    s.close()


def test_call_order_and_argument_errors():
    case = channel_case(12, 12, 12, baffle=False)
    s = AdiSolver3D().Init(case)
    with pytest.raises(CmcError):
        s.TimeStep(0.1, 1, 1, True)          # CreateSegments not called
    s.CreateSegments()
    with pytest.raises(CmcError):
        s.SolveDirection(5, 0.1, 1, LAYER_CUR, LAYER_NEXT)
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("n", [3, 8, 64, 77, 200, 512])
def test_batched_line_solver_against_thomas(oracle_mod, n, fp):
    """Unit level: GPU Thomas (exact, bit-for-bit) and partition+PCR (fast, tolerance) vs Common::SolveTridiagonal."""
    O = oracle_mod
    ft = np.float32 if fp == 4 else np.float64
    rng = np.random.default_rng(n)
    nsys = 37
    a = rng.uniform(-1.2, -0.2, (nsys, n)).astype(ft); c = rng.uniform(-1.2, -0.2, (nsys, n)).astype(ft)
    b = (2.6 + rng.uniform(0, 1, (nsys, n))).astype(ft); d = rng.normal(size=(nsys, n)).astype(ft)
    ref = np.stack([O.solve_tridiagonal(a[i], b[i], c[i], d[i]) for i in range(nsys)])
    assert np.array_equal(solve_tridiagonal_batch(a, b, c, d, "exact"), ref)
    got = solve_tridiagonal_batch(a, b, c, d, "fast")
    assert np.abs(got - ref).max() <= (2e-5 if fp == 4 else 1e-12) * np.abs(ref).max()


@pytest.mark.parametrize("fp", [8, 4])
def test_full_size_properties(fp):
    """BASELINE config sizes (256^3 masked) are too slow for the CPU oracle inside the test budget: check the fast
    path against the bit-exact GPU path (itself pinned to the oracle above) and invariants of the scheme."""
    dims = (256, 256, 256)
    case = channel_case(*dims, fp_bytes=fp)
    sols = {m: AdiSolver3D().Init(case, mode=m) for m in ("exact", "fast")}
    for s in sols.values():
        s.CreateSegments()
    assert sols["exact"].numSegs(0) == sols["fast"].numSegs(0) > 0
    errs = {}
    for m, s in sols.items():
        for i in range(2):
            s.UpdateBoundaries()
            errs[m] = s.TimeStep(case.dt, case.num_global, case.num_local, True)
    assert abs(errs["fast"] - errs["exact"]) <= 1e-6 * errs["exact"]
    out = (case.type == 1).reshape(dims)
    bnd = ((case.type == 2) | (case.type == 3)).reshape(dims)
    noslip = bnd & (case.bc_vel.reshape(dims) == 0)
    _assert_close([sols["exact"].read_field(LAYER_CUR, q) for q in range(4)], [sols["fast"].read_field(LAYER_CUR, q) for q in range(4)], fp, "256^3")
    for q in range(4):
        got = sols["fast"].read_field(LAYER_CUR, q)
        init = (case.vx, case.vy, case.vz, case.T)[q].reshape(dims)
        assert np.array_equal(got[out], init[out])                 # OUT cells are never written by a sweep
        if q < 3:
            assert np.array_equal(got[noslip], init[noslip])       # no-slip rows reproduce the node value exactly
        assert np.isfinite(got).all()
    for s in sols.values():
        s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("nslabs", [2, 4])
def test_slab_decomposition_emulated(oracle_mod, nslabs, fp):
    """The multi-GPU code path (x-slabs, halo exchange, partitioned x-sweep with spike pass / interface solve /
    coupled sweep, distributed residual and readback) with all slabs on ONE device, against the oracle and against
    the single-slab run."""
    O = oracle_mod
    case = channel_case(64, 40, 48, fp_bytes=fp, depth_var=0.25)
    case.outdims = (9, 7, 5)
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    many = AdiSolver3D().Init(case, mode="fast", emulate_slabs=nslabs); many.CreateSegments()
    assert [many.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(4):
        ora.update_boundaries(); one.UpdateBoundaries(); many.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e1 = one.TimeStep(case.dt, case.num_global, case.num_local, True)
        e = many.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref) and abs(e - e1) <= (1e-5 if fp == 4 else 1e-9) * abs(e1)
        if i == 1:
            v_ref, T_ref = ora.get_layer(*case.outdims)
            v, T = many.GetLayer(*case.outdims)
            assert np.allclose(v, v_ref, rtol=0, atol=TOL[fp] * 1e5) and np.allclose(T, T_ref, rtol=0, atol=TOL[fp] * 1e5)
            one.GetLayer(*case.outdims)
        _check_fields(O, ora, many, case, "fast", f"{nslabs} slabs, step {i}")
        _assert_close([one.read_field(LAYER_CUR, q) for q in range(4)], [many.read_field(LAYER_CUR, q) for q in range(4)], fp, "1 slab vs many")
    one.close(); many.close()


def test_slab_decomposition_rejects_unsupported_shapes():
    case = channel_case(12, 24, 24, baffle=False)          # 2 slabs of 8 + 4 planes: fewer than 8 planes in a slab
    s = AdiSolver3D().Init(case, emulate_slabs=2)
    with pytest.raises(CmcError) as ei:
        s.CreateSegments()
    assert ei.value.code == -4
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("planes", [None, [16, 32, 21], [40, 8, 21], "segments", "volume"])
def test_unequal_slabs_against_oracle(oracle_mod, planes, fp):
    """Slabs of different sizes (every slab but the last a multiple of 8 planes, the last one ragged): default split of a
    69-plane grid, hand-made splits, and the reference's load-balancing policies (Grid3D::SplitSegments_X: EVEN_SEGMENTS,
    EVEN_VOLUME) - fused guard-plane stores into neighbours of a different size, partitioned x-solve with a ragged last
    chunk, distributed residual and readback - against the oracle and the single-slab run."""
    from cmc_fluid_solver_b200.solver import SPLIT_EVEN_SEGMENTS, SPLIT_EVEN_VOLUME, split_planes
    O = oracle_mod
    case = channel_case(69, 40, 44, fp_bytes=fp, depth_var=0.25)
    case.outdims = (9, 7, 5)
    if planes == "segments":
        planes = split_planes(case, 3, SPLIT_EVEN_SEGMENTS)
    elif planes == "volume":
        planes = split_planes(case, 3, SPLIT_EVEN_VOLUME)
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    many = AdiSolver3D().Init(case, mode="fast", emulate_slabs=3, planes=planes) if planes else AdiSolver3D().Init(case, mode="fast", emulate_slabs=3)
    many.CreateSegments()
    assert [many.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(4):
        ora.update_boundaries(); one.UpdateBoundaries(); many.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e1 = one.TimeStep(case.dt, case.num_global, case.num_local, True)
        e = many.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref) and abs(e - e1) <= (1e-5 if fp == 4 else 1e-9) * abs(e1)
        if i == 1:
            v_ref, T_ref = ora.get_layer(*case.outdims)
            v, T = many.GetLayer(*case.outdims)
            assert np.allclose(v, v_ref, rtol=0, atol=TOL[fp] * 1e5) and np.allclose(T, T_ref, rtol=0, atol=TOL[fp] * 1e5)
            one.GetLayer(*case.outdims)
        _check_fields(O, ora, many, case, "fast", f"3 unequal slabs {planes}, step {i}")
        _assert_close([one.read_field(LAYER_CUR, q) for q in range(4)], [many.read_field(LAYER_CUR, q) for q in range(4)], fp, "1 slab vs 3")
    one.close(); many.close()


@pytest.mark.parametrize("name", ["nupipe_f64", "nupipe_f32"])
def test_reference_case_on_three_slabs(name):
    """The reference's own shipped case data/3D/example_tests/non_uniform_pipe (53 x 53 x 52 / 54 x 54 x 52: not divisible
    by anything) cut into three slabs, against the golden vector written by the reference CPU solver."""
    case, exp = load_golden(name)
    s = AdiSolver3D().Init(case, mode="fast", emulate_slabs=3)
    s.CreateSegments()
    errs, layers = drive(s, case, exp["steps"])
    got = [s.read_field(LAYER_CUR, q).ravel() for q in range(4)]
    _assert_close(exp["last"], got, case.fp_bytes, name + " on 3 slabs")
    assert np.allclose(errs, exp["err"], rtol=1e-5 if case.fp_bytes == 4 else 1e-9, atol=0)
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
def test_free_boundary_cells_on_a_slab_plane(oracle_mod, fp):
    """BC_FREE valve cells on the two planes either side of a slab boundary (they end / start x-segments there): after
    UpdateBoundaries the neighbour's guard-plane copy of such a cell must hold the node value, not the value the last
    x-sweep solved for it - otherwise the first sweep of the next step reads a different x-neighbour than the
    single-slab run does."""
    O = oracle_mod
    case = channel_case(32, 24, 24, fp_bytes=fp, baffle=False, depth_var=0.0)
    shp = case.shape
    t, bv, bt, T = (a.reshape(shp) for a in (case.type, case.bc_vel, case.bc_temp, case.T))
    sel = np.zeros(shp, bool)
    sel[15:17, 8:13, :] = True
    sel &= (t == 0)
    t[sel] = 3; bv[sel] = 1; bt[sel] = 1; T[sel] = case.baseT
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    two = AdiSolver3D().Init(case, mode="fast", emulate_slabs=2); two.CreateSegments()
    for i in range(4):
        ora.update_boundaries(); ora.time_step(case.dt, case.num_global, case.num_local, True)
        for s in (one, two):
            s.UpdateBoundaries(); s.TimeStep(case.dt, case.num_global, case.num_local, True)
        _check_fields(O, ora, two, case, "fast", f"2 slabs, step {i}")
        _assert_close([one.read_field(LAYER_CUR, q) for q in range(4)], [two.read_field(LAYER_CUR, q) for q in range(4)], fp, "1 slab vs 2")
    one.close(); two.close()


def test_in_node_carrying_free_flags(oracle_mod):
    """The ABI allows bc_vel / bc_temp = BC_FREE on a NODE_IN node; the reference never reads them there."""
    O = oracle_mod
    case = channel_case(24, 20, 18, fp_bytes=8, depth_var=0.2)
    plain = channel_case(24, 20, 18, fp_bytes=8, depth_var=0.2)
    inside = case.type == 0
    case.bc_vel[inside] = 1; case.bc_temp[inside] = 1
    for mode in ("exact", "fast"):
        s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
        o = O.Oracle3D(plain); o.create_segments()
        for i in range(2):
            o.update_boundaries(); o.time_step(case.dt, 2, 2, False)
            s.UpdateBoundaries(); s.TimeStep(case.dt, 2, 2, False)
        _check_fields(O, o, s, case, mode, "IN nodes with FREE flags")
        s.close()


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("fp", [8, 4])
def test_moving_boundary_update_nodes(oracle_mod, mode, fp):
    """3D dynamic grid (Grid3D::Prepare(t) between steps): a baffle that moves one cell along x per step.  The node arrays
    are replaced, the time layers are kept, the line descriptors are rebuilt on the device - against the oracle doing the
    same (update_nodes + create_segments)."""
    O = oracle_mod
    from cmc_fluid_solver_b200.cases import moving_baffle_case
    case = moving_baffle_case(40, 32, 24, fp_bytes=fp, shift=0)
    ora = O.Oracle3D(case); ora.create_segments()
    s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
    for step in range(5):
        if step:
            case = moving_baffle_case(40, 32, 24, fp_bytes=fp, shift=step)
            ora.update_nodes(case); ora.create_segments()
            s.UpdateNodes(case)
            assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        _check_fields(O, ora, s, case, mode, f"moving baffle, step {step}")
    s.close()


def _xline_case(line_types, bc_free_cells=(), fp=8):
    """n x 3 x 3 grid whose centre x-line has the given node types (the x twin of _line_case)."""
    n = len(line_types)
    t = np.full((n, 3, 3), 1, dtype=np.int32)
    t[:, 1, 1] = line_types
    bcv = np.zeros(t.shape, np.int32); bct = np.zeros(t.shape, np.int32)
    for i in bc_free_cells:
        bcv[i, 1, 1] = 1; bct[i, 1, 1] = 1
    ft = np.float32 if fp == 4 else np.float64
    vx = np.zeros(t.shape, ft); T = np.ones(t.shape, ft)
    vx[:, 1, 1] = np.linspace(0.2, 1.0, n)
    z = np.zeros(t.size, ft)
    c = Case(n, 3, 3, .05, .05, .05, 1.0, 0.005, 0.007, 0.0014, 0.05, 2, 2, fp, type=t.ravel(), bc_vel=bcv.ravel(),
             bc_temp=bct.ravel(), vx=vx.ravel(), vy=z, vz=vx.ravel().copy(), T=T.ravel())
    return c


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_slab_local_node_arrays_single_slab(oracle_mod, mode):
    """cmc_adi3d_set_nodes_slab on a one-slab handle (the window is the whole grid): the x-line descriptors come from the
    local rule (k_build_roles_x_local) instead of the scan - same segment counts and fields as the oracle, on the masked
    channel and on an x-line with shared cells, minimum-length segments and valves inside the line."""
    O = oracle_mod
    O.set_threads(1)
    cases = [channel_case(40, 36, 32, fp_bytes=8, depth_var=0.25)]
    line = [1, 2, 0, 2, 0, 0, 2, 1, 2, 0, 2, 1, 1, 2, 0, 0, 0, 3, 0, 0, 0, 0, 3, 0, 0, 2, 1, 1]
    cases += [_xline_case(line, free) for free in ((), (22,), (3, 22))]
    for case in cases:
        ora = O.Oracle3D(case); ora.create_segments()
        case.x_lo, case.x_hi = 0, case.dimx
        s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
        assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
        for i in range(3):
            ora.update_boundaries(); s.UpdateBoundaries()
            ora.time_step(case.dt, 2, 2, False); s.TimeStep(case.dt, 2, 2, False)
        _check_fields(O, ora, s, case, mode, "slab-local node arrays")
        s.close()
    O.set_threads(0)
    # a fluid cell on an x-face of the grid: the local rule cannot know whether its run is ever closed
    bad = channel_case(16, 12, 12, baffle=False)
    bad.type.reshape(bad.shape)[0, 5, 5] = 0
    bad.x_lo, bad.x_hi = 0, 16
    s = AdiSolver3D().Init(bad)
    with pytest.raises(CmcError) as ei:
        s.CreateSegments()
    assert ei.value.code == -4
    s.close()


@pytest.mark.parametrize("emulate", [0, 3])
def test_overlapped_readback_and_upload(oracle_mod, emulate):
    """cmc_adi3d_get_layer_async / _wait and cmc_adi3d_write_layer_async / _commit: the readback of step n lands while step
    n + 1 runs and equals the synchronous GetLayer; an asynchronously uploaded layer equals write_field x 4."""
    O = oracle_mod
    case = channel_case(48, 40, 36, fp_bytes=8, depth_var=0.25)
    case.outdims = (12, 10, 9)
    kw = dict(emulate_slabs=emulate) if emulate else {}
    a = AdiSolver3D().Init(case, mode="fast", **kw); a.CreateSegments()
    b = AdiSolver3D().Init(case, mode="fast", **kw); b.CreateSegments()
    n = 12 * 10 * 9
    bufs = [(np.empty((n, 3)), np.empty(n)) for _ in range(3)]
    full = (np.empty((case.ncells, 3)), np.empty(case.ncells))
    sync_out = []
    for i in range(3):
        for s in (a, b):
            s.UpdateBoundaries(); s.TimeStep(case.dt, case.num_global, case.num_local, False)
        sync_out.append(a.GetLayer(*case.outdims))
        b.GetLayerAsync(bufs[i][0], bufs[i][1], *case.outdims)         # three readbacks in flight over two staging sets
    b.GetLayerAsync(full[0], full[1])
    b.GetLayerWait()
    for i in range(3):
        assert np.array_equal(bufs[i][0], sync_out[i][0]) and np.array_equal(bufs[i][1], sync_out[i][1]), i
    vf, Tf = a.GetLayer()
    assert np.array_equal(full[0], vf) and np.array_equal(full[1], Tf)
    # asynchronous layer upload against write_field
    rng = np.random.default_rng(3)
    fields = [rng.standard_normal(case.shape) for _ in range(4)]
    for q in range(4):
        a.write_field(LAYER_CUR, q, fields[q])
    b.write_layer_async(*fields)
    b.UpdateBoundaries()                       # (work on the solver's stream while the upload runs)
    b.write_layer_commit(LAYER_CUR)
    a.UpdateBoundaries()
    for q in range(4):
        got, ref = b.read_field(LAYER_CUR, q), a.read_field(LAYER_CUR, q)
        bv = (case.type.reshape(case.shape) >= 2)
        assert np.array_equal(got[~bv], ref[~bv]) and np.array_equal(got[~bv], fields[q][~bv])
    a.close(); b.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims,planes", [((64, 40, 44), None), ((69, 40, 44), [16, 32, 21]), ((48, 24, 37), [8, 16, 8, 16])])
def test_exact_mode_on_slabs_is_bit_identical(oracle_mod, dims, planes, fp):
    """CMC_MODE_EXACT on a slab-decomposed grid: the Thomas recurrence of an x-line runs through the slabs as a chain (forward
    elimination first slab -> last, back substitution last -> first, one plane of c' / d' / carried solution between
    neighbours), y and z sweeps are slab-local.  Same operations in the same order as the undivided line: bit-identical with the
    oracle (= the reference CPU solver) in every field - the bit-exact anchor of the multi-slab path."""
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    case.outdims = (9, 7, 5)
    ora = O.Oracle3D(case); ora.create_segments()
    kw = dict(emulate_slabs=len(planes), planes=planes) if planes else dict(emulate_slabs=2)
    s = AdiSolver3D().Init(case, mode="exact", **kw); s.CreateSegments()
    for i in range(4):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        _check_fields(O, ora, s, case, "exact", f"{dims} {planes} step {i}")
        assert abs(e - e_ref) <= (1e-12 if fp == 8 else 1e-6) * abs(e_ref), (e, e_ref)       # (the residual is summed slab by slab)
    v_ref, T_ref = ora.get_layer(*case.outdims)
    v, T = s.GetLayer(*case.outdims)
    assert np.array_equal(v, v_ref) and np.array_equal(T, T_ref)
    s.close()
