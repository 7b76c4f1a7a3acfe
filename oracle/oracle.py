"""TEST INFRASTRUCTURE - ctypes front-end of the CPU oracle (oracle/adi3d_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.

Also holds the reader for the dump files written by oracle/_ref/ref_probe3d_* (the real
reference, see oracle/ref_probe3d.cpp) and helpers to run that binary.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import sys
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from cmc_fluid_solver_b200.cases import Case  # noqa: E402  (plain data container, no CUDA)
LIB_PATH = HERE / "_build" / "liboracle_adi.so"
REF_DIR = HERE / "_ref"

LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP = 0, 1, 2, 3
DIR_X, DIR_Y, DIR_Z = 0, 1, 2
NODE_IN, NODE_OUT, NODE_BOUND, NODE_VALVE = 0, 1, 2, 3
BC_NOSLIP, BC_FREE = 0, 1


def build(force: bool = False) -> Path:
    """Compile the oracle restatement (gcc) if needed and return the library path."""
    src = HERE / "adi3d_oracle.c"
    srcs = [src] + ([HERE / "adi2d_oracle.c"] if (HERE / "adi2d_oracle.c").exists() else [])
    if force or not LIB_PATH.exists() or any(LIB_PATH.stat().st_mtime < s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _np_ft(fp_bytes: int):
    return np.float32 if fp_bytes == 4 else np.float64


def read_probe(path) -> Case:
    """Parse a dump written by oracle/_ref/ref_probe3d_* (format: oracle/ref_probe3d.cpp)."""
    data = Path(path).read_bytes()
    assert data[:8] == b"CMCPROBE", "not a probe dump"
    off = 8
    ints = struct.unpack_from("<12i", data, off)
    off += 48
    ver, fpb, dimx, dimy, dimz, ng, nl, nsteps, ox, oy, oz, _ = ints
    dbl = struct.unpack_from("<9d", data, off)
    off += 72
    dx, dy, dz, dt, v_T, v_vis, t_vis, t_phi, baseT = dbl
    N = dimx * dimy * dimz
    ft = _np_ft(fpb)

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(data, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    case = Case(dimx, dimy, dimz, dx, dy, dz, v_T, v_vis, t_vis, t_phi, dt, ng, nl, fpb, baseT=baseT, outdims=(ox, oy, oz))
    case.type = take(np.int32, N)
    case.bc_vel = take(np.int32, N)
    case.bc_temp = take(np.int32, N)
    case.vx, case.vy, case.vz, case.T = (take(ft, N) for _ in range(4))
    outN = ox * oy * oz
    while off < len(data):
        step, kind = struct.unpack_from("<2i", data, off)
        off += 8
        (err,) = struct.unpack_from("<d", data, off)
        off += 8
        if kind == 5:  # compact record: stride, sample dims, per field (sum, sum of squares, sum |.|, strided subsample)
            stride, sx, sy, sz = struct.unpack_from("<4i", data, off)
            off += 16
            rec = dict(step=step, kind=kind, err=err, stride=stride, sums=[], sumsq=[], sumabs=[], sample=[])
            for _ in range(4):
                s1, s2, sa = struct.unpack_from("<3d", data, off)
                off += 24
                rec["sums"].append(s1); rec["sumsq"].append(s2); rec["sumabs"].append(sa)
                rec["sample"].append(take(ft, sx * sy * sz).reshape(sx, sy, sz))
            case.snapshots.append(rec)
        elif kind == 1:  # GetLayer output: Vec3D[outN] + double[outN]
            vel = take(ft, 3 * outN).reshape(outN, 3)
            T = take(np.float64, outN)
            case.snapshots.append(dict(step=step, kind=kind, err=err, vel=vel, T=T))
        else:
            u, v, w, T = (take(ft, N) for _ in range(4))
            case.snapshots.append(dict(step=step, kind=kind, err=err, u=u, v=v, w=w, T=T))
    return case


def ref_binary(fp_bytes: int) -> Path:
    return REF_DIR / ("ref_probe3d_f32" if fp_bytes == 4 else "ref_probe3d_f64")


def have_ref(fp_bytes: int = 8) -> bool:
    p = ref_binary(fp_bytes)
    return p.exists() and os.access(p, os.X_OK)


def run_ref(data_file, config_file, out_file, nsteps, fp_bytes=8, align=True, dump="last",
            getlayer=False, dt=None, sweep=None, threads=None, timeout=3600, stats=0):
    """Run the real reference CPU solver through the probe driver; returns its stdout."""
    cmd = [str(ref_binary(fp_bytes)), str(data_file), str(config_file), str(out_file), str(int(nsteps))]
    if align:
        cmd.append("align")
    cmd.append(f"dump={dump}")
    if stats:
        cmd.append(f"stats={int(stats)}")
    if getlayer:
        cmd.append("getlayer")
    if dt is not None:
        cmd.append(f"dt={dt!r}")
    if sweep:
        cmd.append(f"sweep={sweep}")
    if threads:
        cmd.append(f"threads={int(threads)}")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference probe failed ({r.returncode}): {r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout


class Oracle3D:
    """The CPU restatement driven like the reference's AdiSolver3D (CPU backend)."""

    def __init__(self, case: Case):
        self.case = case
        self.fp = case.fp_bytes
        self.ft = _np_ft(self.fp)
        self.suf = "_f32" if self.fp == 4 else "_f64"
        L = lib()
        f = self._fn("oracle3d_create")
        f.restype = C.c_void_p
        FP = C.POINTER(C.c_float if self.fp == 4 else C.c_double)
        IP = C.POINTER(C.c_int)
        f.argtypes = [C.c_int] * 3 + [C.c_double] * 7 + [IP] * 3 + [FP] * 4
        arrs_i = [np.ascontiguousarray(a, dtype=np.int32) for a in (case.type, case.bc_vel, case.bc_temp)]
        arrs_f = [np.ascontiguousarray(a, dtype=self.ft) for a in (case.vx, case.vy, case.vz, case.T)]
        self.h = C.c_void_p(f(case.dimx, case.dimy, case.dimz, case.dx, case.dy, case.dz,
                              case.v_T, case.v_vis, case.t_vis, case.t_phi,
                              *[a.ctypes.data_as(IP) for a in arrs_i], *[a.ctypes.data_as(FP) for a in arrs_f]))
        self._FP = FP
        self.err = 0.0

    def _fn(self, name):
        return getattr(lib(), name + self.suf)

    def close(self):
        if self.h:
            f = self._fn("oracle3d_destroy")
            f.argtypes = [C.c_void_p]
            f(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_nodes(self, case: Case):
        """Grid3D::Prepare(t): new Node[] contents, layers untouched; call create_segments() afterwards."""
        f = self._fn("oracle3d_update_nodes")
        IP = C.POINTER(C.c_int)
        f.argtypes = [C.c_void_p] + [IP] * 3 + [self._FP] * 4
        arrs_i = [np.ascontiguousarray(a, dtype=np.int32) for a in (case.type, case.bc_vel, case.bc_temp)]
        arrs_f = [np.ascontiguousarray(a, dtype=self.ft) for a in (case.vx, case.vy, case.vz, case.T)]
        f(self.h, *[a.ctypes.data_as(IP) for a in arrs_i], *[a.ctypes.data_as(self._FP) for a in arrs_f])
        self.case = case

    def create_segments(self):
        f = self._fn("oracle3d_create_segments")
        f.argtypes = [C.c_void_p]
        f(self.h)

    def segments(self, d):
        f = self._fn("oracle3d_num_segments")
        f.argtypes = [C.c_void_p, C.c_int]
        f.restype = C.c_int
        n = f(self.h, d)
        out = np.zeros((max(n, 1), 8), dtype=np.int32)
        g = self._fn("oracle3d_get_segments")
        g.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        g(self.h, d, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out[:n]

    def update_boundaries(self):
        f = self._fn("oracle3d_update_boundaries")
        f.argtypes = [C.c_void_p]
        f(self.h)

    def time_step(self, dt, num_global, num_local, compute_error=True):
        f = self._fn("oracle3d_time_step")
        f.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        f.restype = C.c_int
        e = C.c_double(0)
        rc = f(self.h, float(dt), num_global, num_local, int(bool(compute_error)), C.byref(e))
        self.err = e.value
        if rc != 0:
            raise RuntimeError("Error is too big! %f" % e.value)
        return e.value

    def step_prologue(self):
        f = self._fn("oracle3d_step_prologue")
        f.argtypes = [C.c_void_p]
        f(self.h)

    def solve_direction(self, d, dt, num_local, cur_slot, temp_slot, next_slot):
        f = self._fn("oracle3d_solve_direction")
        f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        f(self.h, d, float(dt), num_local, cur_slot, temp_slot, next_slot)

    def eval_div_error(self, slot=LAYER_NEXT):
        f = self._fn("oracle3d_eval_div_error")
        f.argtypes = [C.c_void_p, C.c_int]
        f.restype = C.c_double
        return f(self.h, slot)

    def field(self, slot, var) -> np.ndarray:
        """Zero-copy view of one field of a layer, shape (dimx, dimy, dimz)."""
        f = self._fn("oracle3d_field")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int]
        f.restype = C.c_void_p
        p = f(self.h, slot, var)
        n = self.case.ncells
        buf = (C.c_float if self.fp == 4 else C.c_double) * n
        return np.frombuffer(buf.from_address(p), dtype=self.ft).reshape(self.case.shape)

    def layer(self, slot):
        return [self.field(slot, q) for q in range(4)]

    def get_layer(self, ox=0, oy=0, oz=0):
        c = self.case
        ox, oy, oz = ox or c.dimx, oy or c.dimy, oz or c.dimz
        vel = np.zeros((ox * oy * oz, 3), dtype=self.ft)
        T = np.zeros(ox * oy * oz, dtype=np.float64)
        f = self._fn("oracle3d_get_layer")
        f.argtypes = [C.c_void_p, self._FP, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]
        f(self.h, vel.ctypes.data_as(self._FP), T.ctypes.data_as(C.POINTER(C.c_double)), ox, oy, oz)
        return vel, T


def set_threads(n: int):
    """OpenMP threads of the oracle's segment loop (0 = all cores)."""
    f = lib().oracle_set_threads_f64
    f.argtypes = [C.c_int]
    f(int(n))


def solve_tridiagonal(a, b, c, d):
    """Common::SolveTridiagonal on copies of a,b,c,d (dtype decides fp32/fp64)."""
    ft = a.dtype
    suf = "_f32" if ft == np.float32 else "_f64"
    f = getattr(lib(), "oracle_solve_tridiagonal" + suf)
    P = C.POINTER(C.c_float if ft == np.float32 else C.c_double)
    f.argtypes = [P] * 5 + [C.c_int]
    a, b, c, d = (np.array(v, dtype=ft, copy=True) for v in (a, b, c, d))
    x = np.zeros_like(a)
    f(*[v.ctypes.data_as(P) for v in (a, b, c, d, x)], len(a))
    return x


# ----------------------------------------------------------------------------------------------- 2D ADI (A16)
def read_probe2d(path) -> dict:
    """Parse a dump written by oracle/_ref/ref_probe2d_f32 (format: oracle/ref_probe2d.cpp)."""
    data = Path(path).read_bytes()
    assert data[:8] == b"CMCPRB2D", "not a 2D probe dump"
    off = 8
    ver, fpb, dimx, dimy, ng, nl, nsteps, ox, oy, out_every = struct.unpack_from("<10i", data, off)
    off += 40
    dx, dy, dt, v_T, v_vis, t_vis, t_phi, startT, _ = struct.unpack_from("<9d", data, off)
    off += 72
    N, outN, ft = dimx * dimy, ox * oy, _np_ft(fpb)

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(data, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    out = dict(fp_bytes=fpb, dimx=dimx, dimy=dimy, dx=dx, dy=dy, dt=dt, v_T=v_T, v_vis=v_vis, t_vis=t_vis, t_phi=t_phi, startT=startT,
               num_global=ng, num_local=nl, outdims=(ox, oy), out_every=out_every, grids={}, layers={}, outputs={}, errs={})
    while off < len(data):
        step, kind = struct.unpack_from("<2i", data, off)
        off += 8
        if kind == 2:       # grid arrays seen by the solver in this step
            out["grids"][step] = dict(type=take(np.int32, N), bc=take(np.int32, N), vx=take(ft, N), vy=take(ft, N), T=take(ft, N))
            continue
        (err,) = struct.unpack_from("<d", data, off)
        off += 12
        if kind == 0:       # current layer after the step (step -1: the initial layer)
            out["layers"][step] = [take(ft, N) for _ in range(3)]
            out["errs"][step] = err
        else:               # GetLayer output
            out["outputs"][step] = (take(ft, 2 * outN).reshape(outN, 2), take(np.float64, outN))
    return out


def ref2d_binary() -> Path:
    return REF_DIR / "ref_probe2d_f32"


def run_ref2d(data_file, config_file, out_file, nsteps=0, dump="every", timeout=600):
    cmd = [str(ref2d_binary()), str(data_file), str(config_file), str(out_file), str(int(nsteps)), f"dump={dump}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference 2D probe failed ({r.returncode}): {r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout


class Oracle2D:
    """oracle/adi2d_oracle.c driven like the reference's AdiSolver2D behind Solver2D (Solver2D.h:24-45)."""

    def __init__(self, dimx, dimy, dx, dy, v_T, v_vis, t_vis, t_phi, startT, fp_bytes=4):
        self.fp, self.ft = fp_bytes, _np_ft(fp_bytes)
        self.suf = "_f32" if fp_bytes == 4 else "_f64"
        self.dimx, self.dimy = dimx, dimy
        f = self._fn("oracle2d_create")
        f.restype = C.c_void_p
        f.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7
        self.h = C.c_void_p(f(dimx, dimy, dx, dy, v_T, v_vis, t_vis, t_phi, startT))
        self._FP = C.POINTER(C.c_float if fp_bytes == 4 else C.c_double)

    def _fn(self, name):
        return getattr(lib(), name + self.suf)

    def close(self):
        if self.h:
            f = self._fn("oracle2d_destroy")
            f.argtypes = [C.c_void_p]
            f(self.h)
            self.h = None

    def set_grid(self, type_, bc, vx, vy, T):
        f = self._fn("oracle2d_set_grid")
        IP = C.POINTER(C.c_int)
        f.argtypes = [C.c_void_p, IP, IP, self._FP, self._FP, self._FP]
        ai = [np.ascontiguousarray(a, dtype=np.int32) for a in (type_, bc)]
        af = [np.ascontiguousarray(a, dtype=self.ft) for a in (vx, vy, T)]
        f(self.h, *[a.ctypes.data_as(IP) for a in ai], *[a.ctypes.data_as(self._FP) for a in af])

    def init_layer(self):
        f = self._fn("oracle2d_init_layer")
        f.argtypes = [C.c_void_p]
        f(self.h)

    def update_boundaries(self):
        f = self._fn("oracle2d_update_boundaries")
        f.argtypes = [C.c_void_p]
        f(self.h)

    def time_step(self, dt, num_global, num_local):
        f = self._fn("oracle2d_time_step")
        f.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
        f.restype = C.c_int
        e = C.c_double(0)
        rc = f(self.h, float(dt), num_global, num_local, C.byref(e))
        if rc:
            raise RuntimeError("Exceeded max number of iterations" if rc == 1 else "Error is too big!")
        return e.value

    def iters(self):
        f = self._fn("oracle2d_iters")
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
        return f(self.h)

    def field(self, layer, var):
        f = self._fn("oracle2d_field")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int]
        f.restype = C.c_void_p
        p = f(self.h, layer, var)
        n = self.dimx * self.dimy
        buf = (C.c_float if self.fp == 4 else C.c_double) * n
        return np.frombuffer(buf.from_address(p), dtype=self.ft)

    def get_layer(self, ox=0, oy=0):
        ox, oy = ox or self.dimx, oy or self.dimy
        vel = np.zeros((ox * oy, 2), dtype=self.ft)
        T = np.zeros(ox * oy, dtype=np.float64)
        f = self._fn("oracle2d_get_layer")
        f.argtypes = [C.c_void_p, self._FP, C.POINTER(C.c_double), C.c_int, C.c_int]
        f(self.h, vel.ctypes.data_as(self._FP), T.ctypes.data_as(C.POINTER(C.c_double)), ox, oy)
        return vel, T
