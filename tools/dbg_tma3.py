import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
from cmc_fluid_solver_b200.solver import DIR_X, DIR_Y, LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP
dims = tuple(int(v) for v in sys.argv[1].split(","))
fp, d = int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
nbad = 0
case = channel_case(*dims, fp_bytes=fp)
a = AdiSolver3D().Init(case, mode="fast"); a.set_option("tma", 0); a.CreateSegments()
b = AdiSolver3D().Init(case, mode="fast"); b.set_option("tma", 3); b.CreateSegments()
a.UpdateBoundaries(); a.TimeStep(case.dt, 2, 1, False)
for slot in range(4):
    for q in range(4):
        b.write_field(slot, q, a.read_field(slot, q))
for rep in range(reps):
    for s in (a, b):
        s.UpdateBoundaries(); s.step_prologue(); s.SolveDirection(d, case.dt, 1, LAYER_CUR, LAYER_NEXT)
    for slot, name in ((LAYER_NEXT, "next"), (LAYER_TEMP, "temp")):
        for q in range(4):
            fa, fb = a.read_field(slot, q), b.read_field(slot, q)
            bad = np.argwhere(fa != fb)
            if len(bad):
                nbad += 1
                print(f"rep {rep} {name} field {q}: {len(bad)} cells differ; max {np.abs(fa.astype(np.float64) - fb).max():.3e}; "
                      f"i%8 hist {np.bincount(bad[:, 0] % 8, minlength=8).tolist()} i//8 hist {np.bincount(bad[:, 0] // 8, minlength=dims[0] // 8).tolist()[:8]}.. "
                      f"k%8 hist {np.bincount(bad[:, 2] % 8, minlength=8).tolist()} j range {bad[:, 1].min()}-{bad[:, 1].max()} first {bad[:5].tolist()}")
    # keep both in the same state for the next repetition
    for slot in range(4):
        for q in range(4):
            b.write_field(slot, q, a.read_field(slot, q))
print("done", dims, fp, d, "reps", reps, "fields with mismatches", nbad)
