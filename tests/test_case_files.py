"""The case writer reproduces the reference's own shipped case files (so that configs written by the tests ARE the
reference's cases): data/3D/example_tests/box_pipe/box_pipe_2D_data.txt token for token.  Needs /root/reference
(present in the build container, not on the GPU box)."""
from pathlib import Path

import pytest

from cmc_fluid_solver_b200.cases import BOX_OUTLINE, write_shape2d_case

REF = Path("/root/reference/data/3D/example_tests/box_pipe")


@pytest.mark.skipif(not REF.is_dir(), reason="/root/reference not present")
def test_box_pipe_case_file_matches_the_shipped_one(tmp_path):
    data, cfg = write_shape2d_case(tmp_path, "box", outline=BOX_OUTLINE, grid_d=0.02, depth=1.0, time_steps=100, out_grid=(54, 54, 52))
    ours = [float(t) if t.replace(".", "", 1).replace("-", "", 1).isdigit() else t for t in data.read_text().split()]
    theirs = [float(t) if t.replace(".", "", 1).replace("-", "", 1).isdigit() else t for t in (REF / "box_pipe_2D_data.txt").read_text().split()]
    assert ours == theirs
    kv = lambda text: {ln.split()[0]: ln.split()[1:] for ln in text.splitlines() if ln.split()}
    ours_cfg, their_cfg = kv(cfg.read_text()), kv((REF / "box_pipe_2D_config.txt").read_text())
    for key, val in their_cfg.items():
        if key in ours_cfg and key not in ("depth_var",):
            a = [float(v) if v.replace(".", "", 1).isdigit() else v for v in val]
            b = [float(v) if v.replace(".", "", 1).isdigit() else v for v in ours_cfg[key]]
            assert a == b, key
