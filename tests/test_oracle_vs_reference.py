"""The oracle restatement (oracle/adi3d_oracle.c) pinned against the REAL reference:
 * golden vectors produced by the reference's compiled CPU solver (tests/golden/, make_golden.py);
 * the reference binary itself (oracle/_ref/ref_probe3d_*, travels to the GPU box) on case files written in
   the reference's formats - bit-for-bit, fp32 and fp64, every cell including NODE_OUT;
 * the informal known answers recorded in SURVEY.md 8(c).
CPU only."""
import numpy as np
import pytest

from conftest import drive, load_golden
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, BOX_OUTLINE, write_shape2d_case


@pytest.mark.parametrize("name", ["box32_f64", "box32_f32", "baffle32_f64", "baffle32_f32", "nupipe_f64", "nupipe_f32"])
def test_oracle_matches_golden_bitwise(oracle_mod, name):
    O = oracle_mod
    case, exp = load_golden(name)
    o = O.Oracle3D(case)
    o.create_segments()
    errs, layers = drive(o, case, exp["steps"])
    assert np.array_equal(np.array(errs), exp["err"])
    for q in range(4):
        assert np.array_equal(o.field(O.LAYER_CUR, q).ravel(), exp["last"][q]), f"field {q} differs from the reference"
    vel, T = layers[0]
    assert np.array_equal(vel, exp["layer0_vel"]) and np.array_equal(T, exp["layer0_T"])


def _ref_case(O, tmp_path, fp, outline, steps, align, **kw):
    data, cfg = write_shape2d_case(tmp_path, "case", outline=outline, **kw)
    out = tmp_path / "dump.bin"
    O.run_ref(data, cfg, out, steps, fp_bytes=fp, align=align, dump="every", getlayer=True)
    return O.read_probe(out)


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("geom", ["box64_aligned", "baffle_unaligned"])
def test_oracle_matches_reference_binary_bitwise(oracle_mod, tmp_path, fp, geom):
    O = oracle_mod
    if not O.have_ref(fp):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    if geom == "box64_aligned":      # the reference's own box_pipe geometry (data/3D/example_tests/box_pipe)
        case = _ref_case(O, tmp_path, fp, BOX_OUTLINE, 11, True, grid_d=0.02, time_steps=100, out_grid=(54, 54, 52))
        assert case.shape == (64, 64, 64) and case.n_in == 115248          # SURVEY 8(c) known answers
    else:                            # masked, depth-perturbed, dims not multiples of anything (54 x 54 x 52)
        case = _ref_case(O, tmp_path, fp, BAFFLE_OUTLINE, 3, False, grid_d=0.02, depth_var=0.2, time_steps=300, out_grid=(20, 21, 19))
        assert case.shape in ((54, 54, 52), (53, 53, 52))   # FTYPE-dependent rounding of the dims
    o = O.Oracle3D(case)
    o.create_segments()
    snaps = {(s["step"], s["kind"]): s for s in case.snapshots}
    nsteps = max(s["step"] for s in case.snapshots) + 1
    for i in range(nsteps):
        o.update_boundaries()
        err = o.time_step(case.dt, case.num_global, case.num_local, (i % 10 == 0) or i == nsteps - 1)
        if i % 10 == 0:
            vel, T = o.get_layer(*case.outdims)
            assert np.array_equal(vel, snaps[(i, 1)]["vel"]) and np.array_equal(T, snaps[(i, 1)]["T"])
        s = snaps[(i, 0)]
        assert err == s["err"]
        for q, n in enumerate("uvwT"):
            assert np.array_equal(o.field(O.LAYER_CUR, q).ravel(), s[n]), f"step {i} field {n}"
    if geom == "box64_aligned":
        # err printed at steps 0 and 10 of the fp32 run = 0.00001252 / 0.00003099
        e0, e10 = snaps[(0, 0)]["err"], snaps[(10, 0)]["err"]
        assert f"{e0:.8f}" == "0.00001252" and f"{e10:.8f}" == "0.00003099"


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("d", ["X", "Y", "Z"])
def test_single_sweep_matches_reference_binary(oracle_mod, tmp_path, fp, d):
    """Component level: TimeStep prologue + ONE SolveDirection (AdiSolver3D.cpp:564-666)."""
    O = oracle_mod
    if not O.have_ref(fp):
        pytest.skip("oracle/_ref not built")
    data, cfg = write_shape2d_case(tmp_path, "case", outline=BAFFLE_OUTLINE, grid_d=0.045, depth_var=0.2, time_steps=300, rim=True)
    out = tmp_path / "dump.bin"
    O.run_ref(data, cfg, out, 0, fp_bytes=fp, align=True, sweep=d)
    case = O.read_probe(out)
    kinds = {s["kind"]: s for s in case.snapshots}
    o = O.Oracle3D(case)
    o.create_segments()
    o.update_boundaries()
    o.step_prologue()
    o.solve_direction("XYZ".index(d), case.dt, case.num_local, O.LAYER_CUR, O.LAYER_TEMP, O.LAYER_NEXT)
    for slot, kind in ((O.LAYER_CUR, 2), (O.LAYER_NEXT, 3), (O.LAYER_TEMP, 4)):
        for q, n in enumerate("uvwT"):
            assert np.array_equal(o.field(slot, q).ravel(), kinds[kind][n]), (slot, n)


def test_thomas_against_dense_solve(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(7)
    for n in (2, 3, 8, 65, 512):
        a = rng.uniform(-1, 0, n); c = rng.uniform(-1, 0, n); b = 2.5 + rng.uniform(0, 1, n); d = rng.normal(size=n)
        x = O.solve_tridiagonal(a, b, c, d)
        A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
        assert np.allclose(A @ x, d, rtol=0, atol=1e-12)


def test_segment_generation_edge_cases(oracle_mod):
    """Grid3D::GenerateListSegments semantics on a hand-made line: a run that reaches the domain edge is dropped,
    a cell may end one segment and start the next, an IN cell at index 0 becomes a segment start."""
    O = oracle_mod
    from cmc_fluid_solver_b200.cases import Case
    nx, ny, nz = 3, 3, 16
    t = np.full((nx, ny, nz), 1, dtype=np.int32)
    #            k: 0 1 2 3 4 5 6 7 8 9 ...
    t[1, 1, :] = [0, 0, 2, 0, 0, 2, 1, 2, 0, 2, 1, 1, 2, 0, 0, 0]
    z = np.zeros(t.size)
    case = Case(nx, ny, nz, .1, .1, .1, 1, .005, .007, .001, .1, 1, 1, 8, type=t.ravel(), bc_vel=np.zeros(t.size, np.int32),
                bc_temp=np.zeros(t.size, np.int32), vx=z, vy=z, vz=z, T=z)
    o = O.Oracle3D(case)
    o.create_segments()
    segs = o.segments(O.DIR_Z)
    got = [(int(s[2]), int(s[5]), int(s[6])) for s in segs if s[0] == 1 and s[1] == 1]
    # [0..2] (starts ON an IN cell at index 0), [2..5] shares cell 2, [7..9]; the run 13..15 is unterminated -> dropped
    assert got == [(0, 2, 3), (2, 5, 4), (7, 9, 3)]


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("iters", [(1, 1), (2, 3), (3, 1)])
def test_iteration_counts_match_reference_binary(oracle_mod, tmp_path, fp, iters):
    """num_global / num_local other than the shipped 4 / 2: the outer (global) and inner (local) iteration structure of
    AdiSolver3D::TimeStep / SolveDirection (AdiSolver3D.cpp:335-358, 589-655) - merges after every local iteration,
    one more merge per global iteration - bit-for-bit against the reference binary."""
    O = oracle_mod
    if not O.have_ref(fp):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ng, nl = iters
    case = _ref_case(O, tmp_path, fp, BAFFLE_OUTLINE, 3, True, grid_d=0.045, depth_var=0.2, time_steps=300, num_global=ng, num_local=nl,
                     out_grid=(8, 9, 7), rim=True)
    assert (case.num_global, case.num_local) == (ng, nl)
    o = O.Oracle3D(case)
    o.create_segments()
    snaps = {(s["step"], s["kind"]): s for s in case.snapshots}
    for i in range(3):
        o.update_boundaries()
        err = o.time_step(case.dt, ng, nl, (i % 10 == 0) or i == 2)      # the driver's computeError cadence (FluidSolver3D.cpp:242)
        if i % 10 == 0:          # the probe's driver loop reads a layer here (GetLayer mutates the previous layer)
            vel, T = o.get_layer(*case.outdims)
            assert np.array_equal(vel, snaps[(i, 1)]["vel"]) and np.array_equal(T, snaps[(i, 1)]["T"])
        s = snaps[(i, 0)]
        assert err == s["err"]
        for q, n in enumerate("uvwT"):
            assert np.array_equal(o.field(O.LAYER_CUR, q).ravel(), s[n]), f"({ng},{nl}) step {i} field {n}"


def test_oracle_matches_config2_golden_statistics(oracle_mod, tmp_path):
    """BASELINE config 2 at its stated size (data/3D box_pipe at 128^3, 20 steps, fp64): the oracle restatement against the
    statistics the reference CPU solver wrote (tests/golden/c2_box128_f64.npz, make_golden_configs.py).  The grid comes
    from the reference's own loader (a 0-step run of the probe dumps its Node[] array)."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, str(GOLDEN))
    import make_golden_configs as MG
    O = oracle_mod
    if not O.have_ref(8):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    z = np.load(GOLDEN / "c2_box128_f64.npz")
    data, cfg = MG.write_case("c2_box128_f64", tmp_path)
    O.run_ref(data, cfg, tmp_path / "nodes.bin", 0, fp_bytes=8, align=True, dump="none")
    case = O.read_probe(tmp_path / "nodes.bin")
    assert case.shape == (128, 128, 128) and case.n_in == int(z["n_in"])
    o = O.Oracle3D(case)
    o.create_segments()
    stride = int(z["stride"])
    want = {int(s): k for k, s in enumerate(z["steps"])}
    for i in range(int(z["steps"][-1]) + 1):
        o.update_boundaries()
        err = o.time_step(case.dt, case.num_global, case.num_local, (i % 10 == 0) or i == 19)
        if i in want:
            k = want[i]
            assert err == z["err"][k]
            for q in range(4):
                f = o.field(O.LAYER_CUR, q)
                assert np.array_equal(f[::stride, ::stride, ::stride], z["sample"][k][q]), f"step {i} field {q}"
                # the probe accumulates serially in index order; numpy sums pairwise: a different rounding of a 2-million-term sum
                assert abs(float(f.sum(dtype=np.float64)) - z["sums"][k][q]) <= 1e-10 * max(z["sumabs"][k][q], 1e-300)
