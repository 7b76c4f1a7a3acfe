"""Generates the golden vectors of tests/golden/*.npz by running the REAL reference CPU solver
(oracle/_ref/ref_probe3d_*, built from /root/reference/src by oracle/build_ref.sh) on small case files
written in the reference's own formats.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each .npz holds the case (grid, params, Node[] arrays), the layer `cur` after the last step, the residual
history and the GetLayer outputs, so tests can check the oracle restatement AND the CUDA path against the
reference without the reference being present (it does not exist on the GPU box).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, BOX_OUTLINE, write_shape2d_case  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent

CASES = [
    # name, fp_bytes, outline, kwargs for the case writer, align, steps
    ("box32_f64", 8, BOX_OUTLINE, dict(grid_d=0.05, depth=1.0, time_steps=400, out_grid=(10, 11, 9), rim=True), True, 4),
    ("box32_f32", 4, BOX_OUTLINE, dict(grid_d=0.05, depth=1.0, time_steps=400, out_grid=(10, 11, 9), rim=True), True, 4),
    ("baffle32_f64", 8, BAFFLE_OUTLINE, dict(grid_d=0.045, depth=1.0, depth_var=0.2, time_steps=300, out_grid=(12, 12, 12), rim=True), True, 3),
    ("baffle32_f32", 4, BAFFLE_OUTLINE, dict(grid_d=0.045, depth=1.0, depth_var=0.2, time_steps=300, out_grid=(12, 12, 12), rim=True), True, 3),
]


# the reference's OWN shipped case data/3D/example_tests/non_uniform_pipe (Shape2D outline + depth_var bottom, no `align`:
# 53 x 53 x 52 in fp32; SURVEY 8(c): NODE_IN = 99959 of 146068, first residual 0.00001366), read by its own loader
REF_CASES = [
    ("nupipe_f64", 8, "non_uniform_pipe/non_uniform_pipe_2D", False, 3),
    ("nupipe_f32", 4, "non_uniform_pipe/non_uniform_pipe_2D", False, 3),
]
REF_DATA = Path("/root/reference/data/3D/example_tests")


def main():
    jobs = [(name, fp, outline, kw, align, steps, None) for name, fp, outline, kw, align, steps in CASES]
    jobs += [(name, fp, None, None, align, steps, stem) for name, fp, stem, align, steps in REF_CASES]
    for name, fp, outline, kw, align, steps, stem in jobs:
        with tempfile.TemporaryDirectory() as td:
            if stem is None:
                data, cfg = write_shape2d_case(td, name, outline=outline, **kw)
            else:       # the reference's files as shipped, CR stripped (bin/Release/run_examples_CPU.sh:12-16)
                data, cfg = Path(td) / "data.txt", Path(td) / "config.txt"
                data.write_bytes((REF_DATA / f"{stem}_data.txt").read_bytes().replace(b"\r", b""))
                cfg.write_bytes((REF_DATA / f"{stem}_config.txt").read_bytes().replace(b"\r", b""))
            out = Path(td) / "dump.bin"
            log = O.run_ref(data, cfg, out, steps, fp_bytes=fp, align=align, dump="every", getlayer=True)
            case = O.read_probe(out)
        cur = [s for s in case.snapshots if s["kind"] == 0]
        lay = [s for s in case.snapshots if s["kind"] == 1]
        last = cur[-1]
        np.savez_compressed(
            HERE / f"{name}.npz",
            dims=np.array(case.shape, dtype=np.int32), spacing=np.array([case.dx, case.dy, case.dz]),
            params=np.array([case.v_T, case.v_vis, case.t_vis, case.t_phi]), dt=case.dt,
            iters=np.array([case.num_global, case.num_local], dtype=np.int32), fp_bytes=fp, steps=steps,
            outdims=np.array(case.outdims, dtype=np.int32),
            type=case.type.astype(np.int8), bc_vel=case.bc_vel.astype(np.int8), bc_temp=case.bc_temp.astype(np.int8),
            vx=case.vx, vy=case.vy, vz=case.vz, T=case.T,
            err=np.array([s["err"] for s in cur]),
            u_last=last["u"], v_last=last["v"], w_last=last["w"], T_last=last["T"],
            layer0_vel=lay[0]["vel"], layer0_T=lay[0]["T"],
        )
        t = case.type.reshape(case.shape)
        assert not any(((t == 0)[sl]).any() for sl in [np.s_[0], np.s_[-1], np.s_[:, 0], np.s_[:, -1], np.s_[:, :, 0], np.s_[:, :, -1]]), "IN cell on a face"
        if name == "nupipe_f32":
            assert case.shape == (53, 53, 52) and case.n_in == 99959 and f"{cur[0]['err']:.8f}" == "0.00001366", "SURVEY 8(c) known answers"
        print(name, case.shape, "NODE_IN", case.n_in, "err", [f"{s['err']:.6e}" for s in cur], log.splitlines()[0])


if __name__ == "__main__":
    main()
