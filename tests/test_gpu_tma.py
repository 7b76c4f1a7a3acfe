"""The TMA-staged x / y sweep kernel (kernels_tma.cu, option "tma") against the CPU oracle and against the direct-load
kernel: 512- and 256-row lines (64 / 32 chunks), y-blocked and plain storage, ragged k-tiles (nz not a multiple of 8),
fp32 and fp64, whole steps and single sweeps.  Fast-mode tolerance: fp64 1e-10, fp32 1e-5 per field."""
import numpy as np
import pytest

from conftest import assert_fields_close
from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
from cmc_fluid_solver_b200.solver import DIR_X, DIR_Y, LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP

pytestmark = pytest.mark.gpu


def _tma(case, mask=3, shape=0):
    s = AdiSolver3D().Init(case, mode="fast")
    s.set_option("tma", mask)
    if shape:
        s.set_option("tma_shape", shape)
    s.CreateSegments()
    return s


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(512, 24, 24), (24, 512, 24), (512, 40, 44), (40, 512, 60), (256, 48, 20), (48, 256, 37), (200, 176, 30)])
def test_tma_steps_against_oracle(oracle_mod, dims, fp):
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    s = _tma(case)
    kinds = [s.get_option("kernel_x"), s.get_option("kernel_y")]
    assert 3 in kinds, f"the TMA kernel is not selected for {dims}: kernel kinds {kinds}"
    for i in range(3):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        assert_fields_close([ora.field(O.LAYER_CUR, q) for q in range(4)], [s.read_field(LAYER_CUR, q) for q in range(4)], fp, f"{dims} step {i}")
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("jb", [0, 8, 16])
def test_tma_blocked_storage_and_single_sweeps(oracle_mod, monkeypatch, jb, fp):
    """Forced y-blocks of 8 / 16 rows (x-tiles jump between blocks, y-lines cross them) and one SolveDirection at a time,
    TMA kernel against the oracle."""
    O = oracle_mod
    if jb:
        monkeypatch.setenv("CMC_JB", str(jb))
    case = channel_case(144, 160, 28, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    for _ in range(2):
        ora.update_boundaries(); ora.time_step(case.dt, 2, 1, False)
    s = _tma(case)
    assert s.get_option("kernel_x") == 3 and s.get_option("kernel_y") == 3
    assert s.storage_block_rows() == jb
    for d in (DIR_Y, DIR_X):
        for slot in (LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP):
            for q in range(4):
                s.write_field(slot, q, ora.field(slot, q))
        ora.update_boundaries(); s.UpdateBoundaries()
        ora.step_prologue(); s.step_prologue()
        ora.solve_direction(d, case.dt, 2, O.LAYER_CUR, O.LAYER_TEMP, O.LAYER_NEXT)
        s.SolveDirection(d, case.dt, 2, LAYER_CUR, LAYER_NEXT)
        for slot in (LAYER_NEXT, LAYER_TEMP):
            assert_fields_close([ora.field(slot, q) for q in range(4)], [s.read_field(slot, q) for q in range(4)], fp, f"dir {d} layer {slot} jb {jb}")
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
def test_tma_against_direct_kernel_256(fp):
    """256^3 masked channel (BASELINE config 3 size): TMA kernel against the direct-load kernel, 2 steps; the two run the
    same arithmetic in the same order, so they agree far inside the tolerance."""
    case = channel_case(256, 256, 256, fp_bytes=fp)
    a = AdiSolver3D().Init(case, mode="fast"); a.set_option("tma", 0); a.CreateSegments()
    b = _tma(case)
    assert b.get_option("kernel_x") == 3 and b.get_option("kernel_y") == 3 and a.get_option("kernel_x") == 1
    for i in range(2):
        for s in (a, b):
            s.UpdateBoundaries(); s.TimeStep(case.dt, case.num_global, case.num_local, True)
    assert_fields_close([a.read_field(LAYER_CUR, q) for q in range(4)], [b.read_field(LAYER_CUR, q) for q in range(4)], fp, "256^3 TMA vs direct")
    a.close(); b.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("shape", ["8x1", "16x1", "8x2", "16x2"])
@pytest.mark.parametrize("dims", [(512, 40, 44), (40, 512, 60), (256, 48, 20), (48, 256, 37), (400, 144, 24)])
def test_tma_tile_shapes_against_oracle(oracle_mod, dims, shape, fp):
    """Every tile shape of the kernel (option "tma_shape" = lines per tile + 256 * CTAs per tile): 8 or 16 lines, whole lines
    in one CTA or a cluster of two CTAs that split every line in halves and couple them through distributed shared memory
    (spike column + 2x2 interface per line) - all against the oracle, ragged k-tiles included."""
    O = oracle_mod
    nl, cl = (int(v) for v in shape.split("x"))
    if max(dims[0], dims[1]) // 8 // cl > 32 and nl == 16:
        pytest.skip("16 lines x 64 chunks would need 1024 threads")
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    s = _tma(case, shape=nl + 256 * cl)
    assert 3 in [s.get_option("kernel_x"), s.get_option("kernel_y")]
    for i in range(2):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        assert_fields_close([ora.field(O.LAYER_CUR, q) for q in range(4)], [s.read_field(LAYER_CUR, q) for q in range(4)], fp, f"{dims} {shape} step {i}")
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(1024, 24, 24), (24, 1024, 24), (1008, 40, 28), (40, 1024, 60), (528, 144, 20)])
def test_lines_up_to_1024_rows_in_fast_mode(oracle_mod, dims, fp):
    """x / y lines of 520 .. 1024 rows (a multiple of 16): a CTA pair holds the line (64 chunks each, the halves coupled through
    distributed shared memory), so fast mode no longer stops at 512 rows along the strided axes.  Against the oracle."""
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    s = _tma(case)
    long_dir = 0 if dims[0] > 512 else 1
    # (fp32 keeps the exact kernels above 512 rows: the partition solve's rounding can exceed 1e-5 of the field there)
    assert s.get_option(("kernel_x", "kernel_y")[long_dir]) == (3 if fp == 8 else 0)
    for i in range(2):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        assert_fields_close([ora.field(O.LAYER_CUR, q) for q in range(4)], [s.read_field(LAYER_CUR, q) for q in range(4)], fp, f"{dims} step {i}")
    s.close()


@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(24, 24, 1024), (40, 28, 1000), (1024, 24, 1024)])
def test_z_lines_up_to_1024_rows_in_fast_mode(oracle_mod, dims, fp):
    """z lines of 520 .. 1024 rows: the direct-load kernel with 128 chunks per line (one more reduction level); together with the
    CTA-pair kernel along x / y, fast mode covers 1024^3-shaped grids.  Against the oracle."""
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    s = AdiSolver3D().Init(case, mode="fast"); s.CreateSegments()
    # (fp32 keeps the exact kernels along z above 512 rows: the partition solve's rounding would exceed 1e-5 there)
    assert s.get_option("kernel_z") == (1 if fp == 8 else 0)
    for i in range(2):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert_fields_close([ora.field(O.LAYER_CUR, q) for q in range(4)], [s.read_field(LAYER_CUR, q) for q in range(4)], fp, f"{dims} step {i}")
        # (the residual is a sum of differences of neighbouring values: in fp32, with 1024 cells per line, it carries the fields'
        # 1e-6 relative rounding amplified by the cancellation)
        assert abs(e - e_ref) <= (5e-3 if fp == 4 else 1e-9) * abs(e_ref)
    s.close()
