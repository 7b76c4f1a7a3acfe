import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from conftest import layer_errors
from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
dims = tuple(int(v) for v in sys.argv[1].split(","))
fp, mask, steps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
case = channel_case(*dims, fp_bytes=fp)
a = AdiSolver3D().Init(case, mode="fast"); a.set_option("tma", 0); a.CreateSegments()
b = AdiSolver3D().Init(case, mode="fast"); b.set_option("tma", mask); b.CreateSegments()
for i in range(steps):
    for s in (a, b):
        s.UpdateBoundaries(); s.TimeStep(case.dt, 4, 2, True)
    fa = [a.read_field(0, q) for q in range(4)]; fb = [b.read_field(0, q) for q in range(4)]
    errs = layer_errors(fa, fb)
    d = np.abs(fa[0].astype(np.float64) - fb[0])
    w = np.unravel_index(np.argmax(d), d.shape)
    print(dims, fp, mask, "step", i, "kinds", b.get_option("kernel_x"), b.get_option("kernel_y"), "errs", ["%.2e" % e for e in errs], "worst u at", w, "n bad", int((d > 1e-4).sum()))
