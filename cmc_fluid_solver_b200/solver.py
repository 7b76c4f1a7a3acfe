"""Host-side mirror of the reference's solver interface for the ADI path.

`AdiSolver3D` keeps the reference's method names, argument meaning, call order and error
behaviour (FluidSolver3D::Solver3D / AdiSolver3D: reference src/FluidSolver3D/Solver3D.h:24-49,
AdiSolver3D.h:52-61; driver call sequence FluidSolver3D.cpp:191-265):

    solver = AdiSolver3D()
    solver.Init(case)                 # Solver3D::Init(backend, csv, grid, params, ...)
    solver.CreateSegments()           # AdiSolver3D::CreateSegments
    loop: solver.UpdateBoundaries(); solver.TimeStep(dt, num_global, num_local, computeError)
          solver.GetLayer(outdimx, outdimy, outdimz)

Everything forwards to the C ABI (include/cmc_adi.h).  No torch types, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Decomp, FluidParams, GridDesc, LIB_PATH, load_library
from .cases import Case

CMC_OK, CMC_ERR_DIVERGED = 0, 1
MODE_FAST, MODE_EXACT = 0, 1
LAYER_CUR, LAYER_HALF, LAYER_NEXT, LAYER_TEMP = 0, 1, 2, 3
DIR_X, DIR_Y, DIR_Z = 0, 1, 2
ERR_THRESHOLD = 0.01


class CmcError(RuntimeError):
    """Any failure reported by the C ABI (carries the status code and cmc_last_error())."""

    def __init__(self, code, msg):
        super().__init__(f"[cmc {code}] {msg}")
        self.code = code


class DivergedError(CmcError):
    """The reference throws std::runtime_error("") after printing "Error is too big!" (AdiSolver3D.cpp:371-374)."""


def lib_path():
    return LIB_PATH


def _check(rc):
    if rc == CMC_OK:
        return
    msg = load_library().cmc_last_error().decode(errors="replace")
    if rc == CMC_ERR_DIVERGED:
        raise DivergedError(rc, msg)
    raise CmcError(rc, msg)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class AdiSolver3D:
    """B200 drop-in for the reference's AdiSolver3D (GPU only)."""

    def __init__(self):
        self._h = None
        self.case = None
        self.diffError = 0.0
        self.rank, self.nranks = 0, 1

    # -- Solver3D::Init ------------------------------------------------------------------------------------
    def Init(self, case: Case, device: int = 0, mode: str = "fast", rank: int = 0, nranks: int = 1, nccl_id: bytes = None,
             emulate_slabs: int = 0, devices=None, planes=None):
        """rank/nranks/nccl_id: one x-slab per process (NCCL).  emulate_slabs=N: all N slabs in this handle, on one GPU.
        devices=[d0, d1, ...]: all slabs in this handle, one per device (the reference's "GPU <n>" mode).
        planes=[n0, n1, ...]: x-planes per slab (e.g. from split_planes()); default = even split on multiples of 8."""
        lib = load_library()
        self.case = case
        self.fp = case.fp_bytes
        self.ft = np.float32 if self.fp == 4 else np.float64
        g = GridDesc(case.dimx, case.dimy, case.dimz, case.dx, case.dy, case.dz)
        p = FluidParams(case.v_T, case.v_vis, case.t_vis, case.t_phi)
        h = C.c_void_p()
        if planes is not None:
            nsl = len(planes)
            pl = (C.c_int32 * nsl)(*planes)
            d = Decomp()
            d.n_slabs, d.device, d.rank, d.planes = nsl, device, rank, pl
            if devices is not None and len(devices) > 1:
                dv = (C.c_int32 * len(devices))(*devices)
                d.kind, d.devices = 3, dv
            elif nranks > 1:
                buf = C.create_string_buffer(nccl_id, 128)
                d.kind, d.nccl_unique_id = 2, C.cast(buf, C.c_void_p)
            else:
                d.kind = 1
            _check(lib.cmc_adi3d_create_ex(C.byref(g), C.byref(p), self.fp, C.byref(d), C.byref(h)))
        elif devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib.cmc_adi3d_create_multi(C.byref(g), C.byref(p), self.fp, arr, len(devices), C.byref(h)))
        elif emulate_slabs > 1:
            _check(lib.cmc_adi3d_create_emulated(C.byref(g), C.byref(p), self.fp, device, emulate_slabs, C.byref(h)))
        elif nranks > 1:
            buf = C.create_string_buffer(nccl_id, 128)
            _check(lib.cmc_adi3d_create_dist(C.byref(g), C.byref(p), self.fp, device, rank, nranks, buf, C.byref(h)))
        else:
            _check(lib.cmc_adi3d_create(C.byref(g), C.byref(p), self.fp, device, C.byref(h)))
        self._h = h
        self.rank, self.nranks = rank, nranks
        x0, nx = C.c_int(0), C.c_int(0)
        _check(lib.cmc_adi3d_slab(h, C.byref(x0), C.byref(nx)))
        self.x0, self.nx = x0.value, nx.value
        self.set_mode(mode)
        arr_i = [np.ascontiguousarray(a, dtype=np.int32) for a in (case.type, case.bc_vel, case.bc_temp)]
        arr_f = [np.ascontiguousarray(a, dtype=self.ft) for a in (case.vx, case.vy, case.vz, case.T)]
        if case.x_lo is not None:
            # slab-local case: the arrays cover this rank's planes plus `halo` planes either side (clipped to the grid)
            halo = max(NODE_HALO, self.x0 - case.x_lo, case.x_hi - self.x0 - self.nx)
            want = slab_window(case.dimx, self.x0, self.nx, halo)
            if (case.x_lo, case.x_hi) != want:
                raise ValueError(f"slab-local case covers planes [{case.x_lo}, {case.x_hi}), the slab [{self.x0}, {self.x0 + self.nx}) needs {want}")
            _check(lib.cmc_adi3d_set_nodes_slab(h, *[_ptr(a) for a in arr_i], *[_ptr(a) for a in arr_f], halo))
        else:
            _check(lib.cmc_adi3d_set_nodes(h, *[_ptr(a) for a in arr_i], *[_ptr(a) for a in arr_f]))
        return self

    def UpdateNodes(self, case: Case):
        """Grid3D::Prepare(t): new node types / boundary kinds / boundary values, time layers untouched.  Rebuilds the line
        descriptors on the device (the 2D solver's per-step CreateSegments, AdiSolver2D.cpp:279-283)."""
        arr_i = [np.ascontiguousarray(a, dtype=np.int32) for a in (case.type, case.bc_vel, case.bc_temp)]
        arr_f = [np.ascontiguousarray(a, dtype=self.ft) for a in (case.vx, case.vy, case.vz, case.T)]
        _check(load_library().cmc_adi3d_update_nodes(self._h, *[_ptr(a) for a in arr_i], *[_ptr(a) for a in arr_f]))
        self.case = case
        self.CreateSegments()

    def set_mode(self, mode):
        m = {"fast": MODE_FAST, "exact": MODE_EXACT}[mode] if isinstance(mode, str) else int(mode)
        _check(load_library().cmc_adi3d_set_option(self._h, b"mode", m))
        self.mode = m

    def set_option(self, key: str, value: int):
        _check(load_library().cmc_adi3d_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64(0)
        _check(load_library().cmc_adi3d_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            load_library().cmc_adi3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- AdiSolver3D::CreateSegments -------------------------------------------------------------------------
    def CreateSegments(self):
        _check(load_library().cmc_adi3d_build_lines(self._h))

    def numSegs(self, d):
        n = C.c_int64(0)
        _check(load_library().cmc_adi3d_num_segments(self._h, d, C.byref(n)))
        return n.value

    # -- Solver3D::UpdateBoundaries --------------------------------------------------------------------------
    def UpdateBoundaries(self):
        _check(load_library().cmc_adi3d_update_boundaries(self._h))

    # -- Solver3D::TimeStep ------------------------------------------------------------------------------------
    def TimeStep(self, dt, num_global, num_local, computeError=True):
        e = C.c_double(self.diffError)
        rc = load_library().cmc_adi3d_time_step(self._h, float(dt), int(num_global), int(num_local), int(bool(computeError)), C.byref(e))
        self.diffError = e.value
        _check(rc)
        return self.diffError

    def TimeStepAsync(self, dt, num_global, num_local, computeError=False):
        _check(load_library().cmc_adi3d_time_step_async(self._h, float(dt), int(num_global), int(num_local), int(bool(computeError))))

    def Sync(self):
        e = C.c_double(0.0)
        rc = load_library().cmc_adi3d_sync(self._h, C.byref(e))
        self.diffError = e.value
        _check(rc)
        return self.diffError

    # -- Solver3D::GetLayer ------------------------------------------------------------------------------------
    def GetLayer(self, outdimx=0, outdimy=0, outdimz=0, vel=None, T=None):
        c = self.case
        ox, oy, oz = outdimx or c.dimx, outdimy or c.dimy, outdimz or c.dimz
        n = ox * oy * oz
        if vel is None:
            vel = np.empty((n, 3), dtype=self.ft)
        if T is None:
            T = np.empty(n, dtype=np.float64)
        _check(load_library().cmc_adi3d_get_layer(self._h, _ptr(vel), _ptr(T), ox, oy, oz))
        return vel, T

    def GetLayerAsync(self, vel, T, outdimx=0, outdimy=0, outdimz=0):
        """GetLayer without the wait: `vel` ((n, 3) FTYPE) and `T` (n doubles) are filled in the background (keep them alive -
        pinned memory keeps the copy asynchronous); collect with GetLayerWait()."""
        c = self.case
        ox, oy, oz = outdimx or c.dimx, outdimy or c.dimy, outdimz or c.dimz
        _check(load_library().cmc_adi3d_get_layer_async(self._h, _ptr(vel), _ptr(T), ox, oy, oz))

    def GetLayerWait(self):
        _check(load_library().cmc_adi3d_get_layer_wait(self._h))

    def output_rows(self, outdimx=0):
        """Rows [lo, hi) of the GetLayer output whose source planes this handle holds (option "local_output": what this rank
        receives)."""
        lo, hi = C.c_int32(0), C.c_int32(0)
        _check(load_library().cmc_adi3d_output_rows(self._h, int(outdimx or self.case.dimx), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def write_layer_async(self, u, v, w, T):
        """Start the upload of a whole layer (dense arrays of this handle's planes) on the copy stream."""
        arrs = [np.ascontiguousarray(a, dtype=self.ft) for a in (u, v, w, T)]
        for a in arrs:
            assert a.size == self.nx * self.case.dimy * self.case.dimz
        self._pending_upload = arrs            # keep the host arrays alive until the commit
        _check(load_library().cmc_adi3d_write_layer_async(self._h, *[_ptr(a) for a in arrs]))

    def write_layer_commit(self, layer=LAYER_CUR):
        _check(load_library().cmc_adi3d_write_layer_commit(self._h, layer))

    # -- hooks -------------------------------------------------------------------------------------------------
    def read_field(self, layer, var) -> np.ndarray:
        c = self.case
        out = np.empty((self.nx, c.dimy, c.dimz), dtype=self.ft)
        _check(load_library().cmc_adi3d_read_field(self._h, layer, var, _ptr(out)))
        return out

    def write_field(self, layer, var, a):
        a = np.ascontiguousarray(a, dtype=self.ft)
        assert a.size == self.nx * self.case.dimy * self.case.dimz
        _check(load_library().cmc_adi3d_write_field(self._h, layer, var, _ptr(a)))

    def read_layer(self, layer):
        return [self.read_field(layer, q) for q in range(4)]

    def step_prologue(self):
        _check(load_library().cmc_adi3d_step_prologue(self._h))

    def SolveDirection(self, d, dt, num_local, cur_layer, next_layer):
        _check(load_library().cmc_adi3d_solve_direction(self._h, d, float(dt), int(num_local), cur_layer, next_layer))

    def EvalDivError(self, layer=LAYER_NEXT):
        e = C.c_double(0.0)
        _check(load_library().cmc_adi3d_eval_div_error(self._h, layer, C.byref(e)))
        return e.value

    def field_sums(self, layer=LAYER_CUR):
        """(sum, sum of squares) of u, v, w, T over the non-OUT cells of the whole grid (all ranks call, all receive)."""
        a = (C.c_double * 8)()
        _check(load_library().cmc_adi3d_field_sums(self._h, layer, a))
        return {n: (a[q], a[4 + q]) for q, n in enumerate("uvwT")}

    def stream(self) -> int:
        s = C.c_void_p()
        _check(load_library().cmc_adi3d_stream(self._h, C.byref(s)))
        return s.value or 0

    def launch_count(self, reset=False) -> int:
        n = C.c_int64(0)
        _check(load_library().cmc_adi3d_launch_count(self._h, C.byref(n), int(reset)))
        return n.value

    TIMING_KINDS = ("sweep_x", "sweep_y", "sweep_z", "merge", "copy", "boundary", "residual", "readback", "comm", "x_spike", "x_interface")

    def set_profile(self, on=True, reset=False):
        _check(load_library().cmc_adi3d_set_option(self._h, b"profile", 2 if (on and reset) else int(bool(on))))

    def timings(self):
        """{kind: (total_ms, calls)} of device time per kernel family (needs set_profile(True))."""
        out = {}
        for k, name in enumerate(self.TIMING_KINDS):
            ms, n = C.c_double(0), C.c_int64(0)
            _check(load_library().cmc_adi3d_get_timing(self._h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def sweep_kernel_name(self, kind: str) -> str:
        """Name of the kernel behind a timing kind ("sweep_x" / "sweep_y" / "sweep_z")."""
        d = "xyz".index(kind[-1])
        v = C.c_int64(0)
        _check(load_library().cmc_adi3d_get_option(self._h, f"kernel_{kind[-1]}".encode(), C.byref(v)))
        ft = "double" if self.fp == 8 else "float"
        return {0: f"k_exact_forward/backward<{ft}> + k_merge", 1: f"k_fast_sweep<{ft},{d}>", 2: f"k_ring_sweep<{ft},{d}>",
                3: f"k_tma_sweep<{ft},{d}>", 4: f"k_fast_sweep<{ft},0,MODE 1> + k_x_interface + k_fast_sweep<{ft},0,MODE 2>",
                5: f"k_tma_sweep<{ft},0,XS> (one-pass slab-coupled)"}[v.value]

    def storage_block_rows(self) -> int:
        v = C.c_int64(0)
        _check(load_library().cmc_adi3d_get_option(self._h, b"jb", C.byref(v)))
        return v.value

    def exchange_kind(self) -> str:
        v = C.c_int64(0)
        _check(load_library().cmc_adi3d_get_option(self._h, b"exchange", C.byref(v)))
        return ("none", "nccl", "fused-stores", "fused-stores-peer-memory", "fused-stores-peer-access")[v.value]

    def device_bytes(self) -> int:
        n = C.c_int64(0)
        _check(load_library().cmc_adi3d_device_bytes(self._h, C.byref(n)))
        return n.value


class AdiSolver2D:
    """B200 drop-in for the reference's 2D AdiSolver2D behind Solver2D (reference src/FluidSolver2D/Solver2D.h:24-45,
    AdiSolver2D.h:40-64; driver call sequence FluidSolver2D.cpp:60-152):

        solver = AdiSolver2D().Init(dimx, dimy, dx, dy, params, startT)      # Solver2D::Init(grid, params)
        loop: solver.set_grid(...)            # what grid.Prepare(t) changed: Grid2D::GetType / GetData per cell
              solver.UpdateBoundaries(); solver.TimeStep(dt, num_global, num_local)
              solver.GetLayer(outdimx, outdimy)

    Results are bit-identical with the reference CPU solver (GPU only, no CPU fallback)."""

    def __init__(self):
        self._h = None
        self.err = 0.0
        self.iters = 0

    def Init(self, dimx, dimy, dx, dy, v_T, v_vis, t_vis, t_phi, startT, fp_bytes=4, device=0):
        lib = load_library()
        self.dimx, self.dimy, self.fp = dimx, dimy, fp_bytes
        self.ft = np.float32 if fp_bytes == 4 else np.float64
        p = FluidParams(v_T, v_vis, t_vis, t_phi)
        h = C.c_void_p()
        _check(lib.cmc_adi2d_create(dimx, dimy, float(dx), float(dy), C.byref(p), float(startT), fp_bytes, device, C.byref(h)))
        self._h = h
        return self

    def close(self):
        if self._h:
            load_library().cmc_adi2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_grid(self, type_, bc, vx, vy, T):
        ai = [np.ascontiguousarray(a, dtype=np.int32) for a in (type_, bc)]
        af = [np.ascontiguousarray(a, dtype=self.ft) for a in (vx, vy, T)]
        _check(load_library().cmc_adi2d_set_grid(self._h, *[_ptr(a) for a in ai], *[_ptr(a) for a in af]))

    def init_layer(self):
        _check(load_library().cmc_adi2d_init_layer(self._h))

    def UpdateBoundaries(self):
        _check(load_library().cmc_adi2d_update_boundaries(self._h))

    def TimeStep(self, dt, num_global, num_local):
        e, it = C.c_double(0.0), C.c_int(0)
        rc = load_library().cmc_adi2d_time_step(self._h, float(dt), int(num_global), int(num_local), C.byref(e), C.byref(it))
        self.err, self.iters = e.value, it.value
        _check(rc)
        return self.err

    def StepHost(self, type_, bc, vx, vy, T, cur, nxt, dt, num_global, num_local):
        """cmc_adi2d_step_host: grid arrays + the host's cur / next layers (3 arrays each, updated in place) up in one copy, one
        TimeStep, both layers down in one copy - what the Solver2D adapter does every step."""
        ai = [np.ascontiguousarray(a, dtype=np.int32) for a in (type_, bc)]
        af = [np.ascontiguousarray(a, dtype=self.ft) for a in (vx, vy, T)]
        for a in list(cur) + list(nxt):
            assert a.dtype == self.ft and a.flags.c_contiguous and a.size == self.dimx * self.dimy
        pc = (C.c_void_p * 3)(*[a.ctypes.data for a in cur]); pn = (C.c_void_p * 3)(*[a.ctypes.data for a in nxt])
        e, it = C.c_double(0.0), C.c_int(0)
        rc = load_library().cmc_adi2d_step_host(self._h, *[_ptr(a) for a in ai], *[_ptr(a) for a in af], pc, pn, float(dt), int(num_global),
                                                int(num_local), C.byref(e), C.byref(it))
        self.err, self.iters = e.value, it.value
        _check(rc)
        return self.err

    def GetLayer(self, outdimx=0, outdimy=0):
        ox, oy = outdimx or self.dimx, outdimy or self.dimy
        vel = np.empty((ox * oy, 2), dtype=self.ft)
        T = np.empty(ox * oy, dtype=np.float64)
        _check(load_library().cmc_adi2d_get_layer(self._h, _ptr(vel), _ptr(T), ox, oy))
        return vel, T

    def read_field(self, layer, var):
        out = np.empty(self.dimx * self.dimy, dtype=self.ft)
        _check(load_library().cmc_adi2d_read_field(self._h, layer, var, _ptr(out)))
        return out

    def write_field(self, layer, var, a):
        a = np.ascontiguousarray(a, dtype=self.ft)
        assert a.size == self.dimx * self.dimy
        _check(load_library().cmc_adi2d_write_field(self._h, layer, var, _ptr(a)))

    def launch_count(self):
        n = C.c_int64(0)
        _check(load_library().cmc_adi2d_launch_count(self._h, C.byref(n)))
        return n.value

    # lower-case aliases (the 2D test driver uses either spelling)
    update_boundaries = UpdateBoundaries
    time_step = TimeStep


def time_step_batch_2d(solvers, dt, num_global, num_local, update_boundaries=True):
    """Advance many independent AdiSolver2D cases (same grid size, precision, device) in ONE launch; returns
    (residuals, outer iterations) per case.  Raises for the first case that diverged."""
    n = len(solvers)
    hs = (C.c_void_p * n)(*[s._h for s in solvers])
    err, it, st = (C.c_double * n)(), (C.c_int * n)(), (C.c_int * n)()
    rc = load_library().cmc_adi2d_time_step_batch(hs, n, float(dt), int(num_global), int(num_local), int(bool(update_boundaries)), err, it, st)
    for s, e, i in zip(solvers, err, it):
        s.err, s.iters = e, i
    _check(rc)
    return list(err), list(it)


def solve_tridiagonal_batch(a, b, c, d, mode="exact"):
    """Batched line solve on the GPU: rows of a,b,c,d are independent systems (Common::SolveTridiagonal)."""
    ft = a.dtype
    assert ft in (np.float32, np.float64)
    a, b, c, d = (np.ascontiguousarray(v, dtype=ft) for v in (a, b, c, d))
    nsys, n = a.shape
    x = np.empty_like(a)
    m = {"fast": MODE_FAST, "exact": MODE_EXACT}[mode]
    _check(load_library().cmc_solve_tridiagonal_batch(a.itemsize, m, nsys, n, _ptr(a), _ptr(b), _ptr(c), _ptr(d), _ptr(x)))
    return x


SPLIT_EVEN_X, SPLIT_EVEN_SEGMENTS, SPLIT_EVEN_VOLUME = 0, 1, 2
NODE_HALO = 2          # planes a slab-local node array carries on either side (cmc_adi3d_set_nodes_slab)


def slab_window(dimx, x0, nx, halo=NODE_HALO):
    """Planes [lo, hi) a slab-local case of the slab [x0, x0 + nx) has to cover."""
    return max(x0 - halo, 0), min(x0 + nx + halo, dimx)


def default_split(dimx, dimy, dimz, n_slabs):
    """(x0, nx) of every slab of the default split (even, cuts on multiples of 8 planes) - what cmc_adi3d_create_dist uses."""
    g = GridDesc(dimx, dimy, dimz, 1.0, 1.0, 1.0)
    out = (C.c_int32 * n_slabs)()
    _check(load_library().cmc_split_planes(SPLIT_EVEN_X, C.byref(g), None, n_slabs, out))
    x0, res = 0, []
    for n in out:
        res.append((x0, n)); x0 += n
    return res



def split_planes(case: Case, n_slabs: int, policy: int = SPLIT_EVEN_X):
    """x-planes per slab for `n_slabs` slabs (Grid3D::Split: EVEN_X / EVEN_SEGMENTS / EVEN_VOLUME, cuts on multiples of 8)."""
    g = GridDesc(case.dimx, case.dimy, case.dimz, case.dx, case.dy, case.dz)
    t = np.ascontiguousarray(case.type, dtype=np.int32)
    out = (C.c_int32 * n_slabs)()
    _check(load_library().cmc_split_planes(policy, C.byref(g), _ptr(t), n_slabs, out))
    return list(out)


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId for cmc_adi3d_create_dist (call on rank 0, broadcast to the others)."""
    buf = C.create_string_buffer(128)
    _check(load_library().cmc_nccl_unique_id(buf))
    return buf.raw
