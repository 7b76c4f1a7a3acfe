"""Parity of what bench.py actually runs, and of BASELINE.json's configurations at their stated sizes.

 * 512-row lines in every direction (the kernel instantiations of the 512^3 benchmark: k_fast_sweep<.,DIR,64,.>,
   CR(3 levels)+PCR reduced solve, the 64-row y-blocks, the TMA-staged x / y tiles) on thin grids the CPU oracle
   finishes in seconds: exact mode bit-for-bit, fast mode within tolerance (fp64 1e-10, fp32 1e-5, per field);
 * a grid whose x-plane exceeds 512 KB, so that the DEFAULT y-blocking rule (cmc_adi.cu Engine::init) is what runs;
 * config 2 (data/3D box_pipe at 128^3, 20 steps, fp64) and config 3 (masked channel at 256^3, 10 steps, fp32 + fp64)
   through the reference's own loader + the Solver3D adapter + libcmcadi.so (oracle/_ref/dropin3d_*), against golden
   statistics written by the reference CPU solver (tests/golden/make_golden_configs.py);
 * config 4 (512^3 masked, fp64): fast mode against the bit-exact GPU mode (itself pinned to the oracle) for one step.
"""
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, assert_fields_close
from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
from cmc_fluid_solver_b200.solver import LAYER_CUR

sys.path.insert(0, str(GOLDEN))
import make_golden_configs as MG  # noqa: E402

pytestmark = pytest.mark.gpu
REF = ROOT / "oracle" / "_ref"
TOL = {8: 1e-10, 4: 1e-5}


def _steps_vs_oracle(O, case, steps, what):
    ora = O.Oracle3D(case); ora.create_segments()
    sols = {m: AdiSolver3D().Init(case, mode=m) for m in ("exact", "fast")}
    for s in sols.values():
        s.CreateSegments()
        assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(steps):
        ora.update_boundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        ref = [ora.field(O.LAYER_CUR, q) for q in range(4)]
        for m, s in sols.items():
            s.UpdateBoundaries()
            e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
            assert abs(e - e_ref) <= (1e-5 if case.fp_bytes == 4 else 1e-9) * abs(e_ref), (m, e, e_ref)
            got = [s.read_field(LAYER_CUR, q) for q in range(4)]
            if m == "exact":
                for q in range(4):
                    assert np.array_equal(ref[q], got[q]), f"{what}: exact mode differs from the oracle, field {q}, step {i}"
            else:
                assert_fields_close(ref, got, case.fp_bytes, f"{what} step {i}")
    return sols


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(512, 24, 24), (24, 512, 24), (24, 24, 512), (512, 40, 40), (40, 512, 512 // 8)])
def test_512_row_lines_against_oracle(oracle_mod, dims, fp):
    """Lines of 512 rows (64 chunks: CR down to 8 rows + PCR) along x, y and z; the last two shapes carry the baffle."""
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    for s in _steps_vs_oracle(oracle_mod, case, 3, f"{dims} fp{fp * 8}").values():
        s.close()


@pytest.mark.parametrize("fp", [8, 4])
def test_default_y_blocking_against_oracle(oracle_mod, fp):
    """An x-plane of 512 x 272 values (> 512 KB in both precisions): the library's own blocking rule picks the block
    height (no CMC_JB override), y-lines of 512 rows cross several blocks."""
    case = channel_case(24, 512, 264, fp_bytes=fp, depth_var=0.25)
    sols = _steps_vs_oracle(oracle_mod, case, 2, f"default blocking fp{fp * 8}")
    for s in sols.values():
        assert s.storage_block_rows() in (32, 64, 128), s.storage_block_rows()
        s.close()


# ---- BASELINE configs through the reference's loader + adapter --------------------------------------------------
def _run_dropin(name, tmp_path, solver):
    O = __import__("oracle.oracle", fromlist=["x"])
    fp = MG.CONFIGS[name][0]
    tag = "f32" if fp == 4 else "f64"
    binary = REF / f"dropin3d_{tag}"
    if not binary.exists():
        pytest.skip("oracle/_ref/dropin3d_* not built (needs /root/reference at build time)")
    data, cfg = MG.write_case(name, tmp_path)
    steps, extra = MG.stats_args(name)
    out = tmp_path / f"{solver}.bin"
    cmd = [str(binary), str(data), str(cfg), str(out), str(steps), "align", f"solver={solver}"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return MG.pack(O.read_probe(out))


def _compare_with_golden(name, got, exact):
    z = np.load(GOLDEN / f"{name}.npz")
    fp = int(z["fp_bytes"])
    assert tuple(got["dims"]) == tuple(z["dims"]) and got["n_in"] == int(z["n_in"]) and list(got["steps"]) == list(z["steps"])
    for si, step in enumerate(z["steps"]):
        ref_s, got_s = z["sample"][si], got["sample"][si]
        if exact:
            # identical fields => identical host-side sums (same accumulation code) and identical samples
            assert np.array_equal(ref_s, got_s), f"{name} step {step}: exact mode differs from the reference"
            assert np.array_equal(z["sums"][si], got["sums"][si]) and np.array_equal(z["sumsq"][si], got["sumsq"][si])
            # the residual is a sum over all cells (serial on the CPU, a block reduction on the device)
            assert abs(got["err"][si] - z["err"][si]) <= 1e-10 * abs(z["err"][si])
        else:
            assert_fields_close(list(ref_s), list(got_s), fp, f"{name} step {step} (strided subsample)")
            tol = TOL[fp]
            for q in range(4):
                assert abs(got["sums"][si][q] - z["sums"][si][q]) <= tol * max(z["sumabs"][si][q], 1e-300), (name, step, q)
                assert abs(np.sqrt(got["sumsq"][si][q]) - np.sqrt(z["sumsq"][si][q])) <= tol * max(np.sqrt(z["sumsq"][si][q]), 1e-300), (name, step, q)
            assert abs(got["err"][si] - z["err"][si]) <= (1e-5 if fp == 4 else 1e-9) * abs(z["err"][si])


@pytest.mark.parametrize("solver", ["b200exact", "b200"])
def test_config2_box_pipe_128_fp64(tmp_path, solver):
    """BASELINE config 2: the reference's data/3D box_pipe case at 128^3, 20 steps, fp64, tolerance check vs CPU."""
    got = _run_dropin("c2_box128_f64", tmp_path, solver)
    assert tuple(got["dims"]) == (128, 128, 128)
    _compare_with_golden("c2_box128_f64", got, exact=solver == "b200exact")


@pytest.mark.parametrize("name", ["c3_baffle256_f64", "c3_baffle256_f32"])
@pytest.mark.parametrize("solver", ["b200exact", "b200"])
def test_config3_masked_channel_256(tmp_path, name, solver):
    """BASELINE config 3: masked 256^3 channel, 10 steps, fp32 and fp64."""
    got = _run_dropin(name, tmp_path, solver)
    assert tuple(got["dims"]) == (256, 256, 256)
    _compare_with_golden(name, got, exact=solver == "b200exact")


def test_config4_512_fast_against_exact_fp64():
    """BASELINE config 4 (the benchmarked case): one step of the 512^3 masked channel, fast mode against the bit-exact
    mode on the device (the CPU reference needs ~35 GB and minutes per step: its run on the GPU box's host is kept as a
    log under profiles/)."""
    dims = (512, 512, 512)
    case = channel_case(*dims, fp_bytes=8, depth_var=0.2)
    fields = {}
    errs = {}
    for m in ("exact", "fast"):
        s = AdiSolver3D().Init(case, mode=m); s.CreateSegments()
        s.UpdateBoundaries()
        errs[m] = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        fields[m] = [s.read_field(LAYER_CUR, q) for q in range(4)]
        if m == "fast":
            assert s.storage_block_rows() == 64
        s.close()
    assert abs(errs["fast"] - errs["exact"]) <= 1e-9 * abs(errs["exact"])
    assert_fields_close(fields["exact"], fields["fast"], 8, "512^3 fast vs exact")
    out = (case.type == 1).reshape(dims)
    for q in range(4):
        init = (case.vx, case.vy, case.vz, case.T)[q].reshape(dims)
        assert np.array_equal(fields["fast"][q][out], init[out])       # OUT cells are never written by a sweep
        assert np.isfinite(fields["fast"][q]).all()
