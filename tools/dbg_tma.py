import sys, os
sys.path.insert(0, "/root/repo")
from cmc_fluid_solver_b200 import AdiSolver3D
from cmc_fluid_solver_b200.cases import channel_case
dims = tuple(int(v) for v in sys.argv[1].split(","))
fp, mask = int(sys.argv[2]), int(sys.argv[3])
case = channel_case(*dims, fp_bytes=fp, depth_var=0.25)
s = AdiSolver3D().Init(case, mode="fast"); s.set_option("tma", mask); s.CreateSegments()
try:
    for i in range(2):
        s.UpdateBoundaries(); e = s.TimeStep(case.dt, 4, 2, True)
    print(dims, fp, mask, "kinds", s.get_option("kernel_x"), s.get_option("kernel_y"), "OK err", e)
except Exception as ex:
    print(dims, fp, mask, "FAILED", ex)
