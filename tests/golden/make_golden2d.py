#!/usr/bin/env python
"""Golden vectors of the reference's 2D ADI solver (BASELINE config 1).

Runs the UNMODIFIED reference 2D solver (oracle/_ref/ref_probe2d_f32, built by oracle/build_ref.sh from
/root/reference/src) on the reference's own case data/2D/box_pipe (`solver ADI` instead of the shipped `solver Stable`,
which never terminates in a Linux build - SURVEY 8(c)) and stores what crosses the Solver2D interface:

    python tests/golden/make_golden2d.py          # needs /root/reference; writes tests/golden/box_pipe2d_f32.npz

The solver reads Grid2D data only in non-NODE_IN cells (boundary rows, UpdateBoundaries), so the per-step grid arrays
are stored sparsely at those cells; the full layers are stored for a few steps, the residual and a float64 checksum of
every field for all steps."""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

REF = Path("/root/reference/data/2D/box_pipe")
KEEP = (0, 1, 9, 24, 48)


def main():
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "data.txt").write_bytes((REF / "box_pipe_data.txt").read_bytes().replace(b"\r", b""))
        cfg = (REF / "box_pipe_config.txt").read_bytes().replace(b"\r", b"").decode()
        cfg = "\n".join("solver\t\tADI" if ln.startswith("solver") else ln for ln in cfg.splitlines()) + "\n"
        (td / "config.txt").write_text(cfg)
        out = O.run_ref2d(td / "data.txt", td / "config.txt", td / "out.bin", 0, "every")
        d = O.read_probe2d(td / "out.bin")
    steps = len(d["grids"])
    g0 = d["grids"][0]
    assert all(np.array_equal(g0["type"], d["grids"][s]["type"]) and np.array_equal(g0["bc"], d["grids"][s]["bc"]) for s in range(steps))
    nonin = np.flatnonzero(g0["type"] != 0)
    sparse = np.stack([np.stack([d["grids"][s][k][nonin] for k in ("vx", "vy", "T")]) for s in range(steps)])      # [steps][3][n]
    sums = np.array([[float(np.sum(d["layers"][s][q].astype(np.float64))) for q in range(3)] for s in range(steps)])
    l2 = np.array([[float(np.sqrt(np.sum(d["layers"][s][q].astype(np.float64) ** 2))) for q in range(3)] for s in range(steps)])
    np.savez_compressed(
        Path(__file__).resolve().parent / "box_pipe2d_f32.npz",
        dims=np.array([d["dimx"], d["dimy"]]), spacing=np.array([d["dx"], d["dy"]]), dt=d["dt"],
        params=np.array([d["v_T"], d["v_vis"], d["t_vis"], d["t_phi"]]), startT=d["startT"],
        iters=np.array([d["num_global"], d["num_local"]]), outdims=np.array(d["outdims"]), steps=steps,
        type=g0["type"].astype(np.int8), bc=g0["bc"].astype(np.int8), nonin=nonin.astype(np.int32), grid_sparse=sparse,
        layer_init=np.stack(d["layers"][-1]), keep=np.array(KEEP), layers=np.stack([np.stack(d["layers"][s]) for s in KEEP]),
        err=np.array([d["errs"][s] for s in range(steps)]), sums=sums, l2=l2,
        out_steps=np.array(sorted(d["outputs"])), out_vel=np.stack([d["outputs"][s][0] for s in sorted(d["outputs"])]),
        out_T=np.stack([d["outputs"][s][1] for s in sorted(d["outputs"])]))
    print(out.splitlines()[0])
    print("steps", steps, "non-IN cells", nonin.size, "err[0], err[-1] =", d["errs"][0], d["errs"][steps - 1])


def main_dynamic(name="heart_MR", steps=10):
    """A moving-boundary case of the reference (data/2D/heart_MR: 25 frames, the node types change almost every step):
    the full grid arrays of the first `steps` steps, every layer's checksum and the last layer."""
    ref = Path("/root/reference/data/2D") / name
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "data.txt").write_bytes((ref / f"{name}_data.txt").read_bytes().replace(b"\r", b""))
        cfg = (ref / f"{name}_config.txt").read_bytes().replace(b"\r", b"").decode()
        (td / "config.txt").write_text("\n".join("solver\t\tADI" if ln.startswith("solver") else ln for ln in cfg.splitlines()) + "\n")
        O.run_ref2d(td / "data.txt", td / "config.txt", td / "out.bin", steps, "every")
        d = O.read_probe2d(td / "out.bin")
    g = d["grids"]
    changed = sum(not np.array_equal(g[s]["type"], g[s - 1]["type"]) for s in range(1, steps))
    np.savez_compressed(
        Path(__file__).resolve().parent / f"{name.lower()}2d_f32.npz",
        dims=np.array([d["dimx"], d["dimy"]]), spacing=np.array([d["dx"], d["dy"]]), dt=d["dt"],
        params=np.array([d["v_T"], d["v_vis"], d["t_vis"], d["t_phi"]]), startT=d["startT"],
        iters=np.array([d["num_global"], d["num_local"]]), steps=steps,
        type=np.stack([g[s]["type"] for s in range(steps)]).astype(np.int8), bc=np.stack([g[s]["bc"] for s in range(steps)]).astype(np.int8),
        gvx=np.stack([g[s]["vx"] for s in range(steps)]), gvy=np.stack([g[s]["vy"] for s in range(steps)]), gT=np.stack([g[s]["T"] for s in range(steps)]),
        layer_init=np.stack(d["layers"][-1]), layer_last=np.stack(d["layers"][steps - 1]),
        err=np.array([d["errs"][s] for s in range(steps)]),
        sums=np.array([[float(np.sum(d["layers"][s][q].astype(np.float64))) for q in range(3)] for s in range(steps)]))
    print(name, d["dimx"], d["dimy"], "steps", steps, "node types changed in", changed, "of them")


if __name__ == "__main__":
    main()
    main_dynamic()
