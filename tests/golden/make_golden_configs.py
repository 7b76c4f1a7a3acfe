"""Golden statistics of the BASELINE.json configurations that are too large for whole-field fixtures, produced by
the REAL reference CPU solver (oracle/_ref/ref_probe3d_*, built from /root/reference/src by oracle/build_ref.sh):

  c2_box128_f64     config 2: data/3D/example_tests/box_pipe (Shape2D outline of the shipped case, written by
                    cases.write_shape2d_case - byte-identical with box_pipe_2D_data.txt, checked in tests/test_case_files.py)
                    with grid_d* = 0.01 + `align` => 128^3, shipped time_steps 100 (dt 0.1), 20 steps, fp64
  c3_baffle256_f64  config 3: the masked channel (wall-attached baffle + depth_var 0.2 bottom) at 256^3
  c3_baffle256_f32  (grid_d* 0.0045, depth 1.14, `align`), 10 steps, fp64 and fp32

Run in the build container (needs /root/reference; about 4 minutes on 8 cores):

    python tests/golden/make_golden_configs.py

Each .npz holds, for the selected steps, the residual and per field (u, v, w, T) the sum, the sum of squares, the sum of
absolute values (accumulated in double over the dense array in index order) and the strided subsample
[::stride, ::stride, ::stride] of the layer `cur` - written by the probe's `stats=` records (oracle/ref_probe3d.cpp).
The GPU tests (tests/test_gpu_configs.py) run the SAME case files through the reference's loader + our Solver3D
adapter + libcmcadi.so (oracle/_ref/dropin3d_*) and compare the same records.
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, BOX_OUTLINE, write_shape2d_case  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent

# name -> (fp_bytes, outline, case-writer kwargs, steps, recorded steps, sample stride)
CONFIGS = {
    "c2_box128_f64": (8, BOX_OUTLINE, dict(grid_d=0.01, depth=1.0, time_steps=100, out_grid=(54, 54, 52)), 20, (0, 9, 19), 8),
    "c3_baffle256_f64": (8, BAFFLE_OUTLINE, dict(grid_d=0.0045, depth=1.14, depth_var=0.2, time_steps=100, out_grid=(32, 32, 32)), 10, (0, 4, 9), 16),
    "c3_baffle256_f32": (4, BAFFLE_OUTLINE, dict(grid_d=0.0045, depth=1.14, depth_var=0.2, time_steps=100, out_grid=(32, 32, 32)), 10, (0, 4, 9), 16),
}


def write_case(name, directory):
    fp, outline, kw, steps, rec, stride = CONFIGS[name]
    return write_shape2d_case(directory, name, outline=outline, **kw)


def stats_args(name):
    fp, outline, kw, steps, rec, stride = CONFIGS[name]
    return steps, ["dump=list:" + ",".join(str(r) for r in rec), f"stats={stride}"]


def pack(case):
    snaps = [s for s in case.snapshots if s["kind"] == 5]
    return dict(dims=np.array(case.shape, dtype=np.int32), n_in=case.n_in, steps=np.array([s["step"] for s in snaps], dtype=np.int32),
                err=np.array([s["err"] for s in snaps]), stride=snaps[0]["stride"],
                sums=np.array([s["sums"] for s in snaps]), sumsq=np.array([s["sumsq"] for s in snaps]),
                sumabs=np.array([s["sumabs"] for s in snaps]), sample=np.stack([np.stack(s["sample"]) for s in snaps]))


def main():
    for name, (fp, outline, kw, steps, rec, stride) in CONFIGS.items():
        with tempfile.TemporaryDirectory() as td:
            data, cfg = write_case(name, td)
            out = Path(td) / "stats.bin"
            t0 = time.time()
            log = O.run_ref(data, cfg, out, steps, fp_bytes=fp, align=True, dump="list:" + ",".join(str(r) for r in rec), stats=stride)
            case = O.read_probe(out)
        g = pack(case)
        np.savez_compressed(HERE / f"{name}.npz", fp_bytes=fp, **g)
        print(name, case.shape, "NODE_IN", case.n_in, "err", [f"{e:.6e}" for e in g["err"]], f"{time.time() - t0:.0f} s", log.splitlines()[0])


if __name__ == "__main__":
    main()
