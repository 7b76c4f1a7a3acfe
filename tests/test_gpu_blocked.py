"""y-blocked field storage (Layout: [j / jb][i][j % jb][k]) on grids small enough for the oracle.
The default picks blocks only once an x-plane exceeds 512 KB, so the small parity grids of test_gpu_parity.py all run in
the one-block layout; here CMC_JB forces 8- and 16-row blocks (several blocks, a ragged last block, z-tiles and chunks
that end exactly at block edges) through the same oracle comparisons: exact mode bit-for-bit, fast mode within
tolerance, the slab-decomposed path (fused halo / interface stores across blocks) and the GetLayer readback."""
import numpy as np
import pytest

from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
from test_gpu_parity import TOL, _assert_close, _check_fields, LAYER_CUR

pytestmark = pytest.mark.gpu


@pytest.fixture
def jb(request, monkeypatch):
    monkeypatch.setenv("CMC_JB", str(request.param))
    return request.param


@pytest.mark.parametrize("jb", [8, 16], indirect=True)
@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("fp", [8, 4])
@pytest.mark.parametrize("dims", [(40, 36, 32), (24, 130, 17), (37, 29, 23)])
def test_blocked_steps_against_oracle(oracle_mod, jb, dims, fp, mode):
    O = oracle_mod
    case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
    case.outdims = (7, 5, 6)
    ora = O.Oracle3D(case); ora.create_segments()
    s = AdiSolver3D().Init(case, mode=mode); s.CreateSegments()
    assert [s.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(4):
        ora.update_boundaries(); s.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e = s.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref)
        if i == 1:
            v_ref, T_ref = ora.get_layer(*case.outdims)
            v, T = s.GetLayer(*case.outdims)
            if mode == "exact":
                assert np.array_equal(v, v_ref) and np.array_equal(T, T_ref)
            else:
                assert np.allclose(v, v_ref, rtol=0, atol=TOL[fp] * 1e5) and np.allclose(T, T_ref, rtol=0, atol=TOL[fp] * 1e5)
        _check_fields(O, ora, s, case, mode, f"jb {jb} step {i}")
    s.close()


@pytest.mark.parametrize("jb", [8, 16], indirect=True)
@pytest.mark.parametrize("fp", [8, 4])
def test_blocked_slab_decomposition(oracle_mod, jb, fp):
    O = oracle_mod
    case = channel_case(64, 40, 48, fp_bytes=fp, depth_var=0.25)
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    many = AdiSolver3D().Init(case, mode="fast", emulate_slabs=4); many.CreateSegments()
    assert many.exchange_kind() == "fused-stores"
    for i in range(3):
        ora.update_boundaries(); one.UpdateBoundaries(); many.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e1 = one.TimeStep(case.dt, case.num_global, case.num_local, True)
        e = many.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref) and abs(e - e1) <= (1e-5 if fp == 4 else 1e-9) * abs(e1)
        _check_fields(O, ora, many, case, "fast", f"jb {jb}, 4 slabs, step {i}")
        _assert_close([one.read_field(LAYER_CUR, q) for q in range(4)], [many.read_field(LAYER_CUR, q) for q in range(4)], fp, "1 slab vs 4")
    one.close(); many.close()


@pytest.mark.parametrize("jb", [8], indirect=True)
def test_blocked_field_roundtrip(jb):
    """write_field / read_field (one 3-D copy per y-block) are inverse for every layer buffer."""
    case = channel_case(24, 37, 19, fp_bytes=8)
    s = AdiSolver3D().Init(case); s.CreateSegments()
    rng = np.random.default_rng(5)
    for layer in range(4):
        a = rng.standard_normal((24, 37, 19))
        s.write_field(layer, layer, a)
        assert np.array_equal(s.read_field(layer, layer), a)
    s.close()
