"""Cases for the ADI path: the `Case` container (grid + FluidParams + the reference's Node[] as
arrays), a synthetic masked channel generator for benchmark-sized grids, and a writer for case files
in the reference's own text formats (Shape2D data + config) so that the
reference's loader (Grid3D::LoadFromFile / Grid2D::LoadFromFile, reference src/FluidSolver3D/Grid3D.cpp:488-513,
src/FluidSolver2D/Grid2D.cpp:268-372) can read them.

Node conventions follow Grid3D::Prepare2D (reference src/FluidSolver3D/Grid3D.cpp:608-665).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

NODE_IN, NODE_OUT, NODE_BOUND, NODE_VALVE = 0, 1, 2, 3
BC_NOSLIP, BC_FREE = 0, 1


@dataclass
class Case:
    """A solver case: grid, fluid parameters and the reference's Node[] as structure of arrays."""
    dimx: int
    dimy: int
    dimz: int
    dx: float
    dy: float
    dz: float
    v_T: float
    v_vis: float
    t_vis: float
    t_phi: float
    dt: float
    num_global: int
    num_local: int
    fp_bytes: int
    type: np.ndarray = None
    bc_vel: np.ndarray = None
    bc_temp: np.ndarray = None
    vx: np.ndarray = None
    vy: np.ndarray = None
    vz: np.ndarray = None
    T: np.ndarray = None
    baseT: float = 1.0
    outdims: tuple = (0, 0, 0)
    snapshots: list = field(default_factory=list)
    # slab-local cases: the node arrays cover only the x-planes [x_lo, x_hi) of the dimx x dimy x dimz grid
    x_lo: int = None
    x_hi: int = None

    @property
    def shape(self):
        return (self.dimx, self.dimy, self.dimz)

    @property
    def ncells(self):
        return self.dimx * self.dimy * self.dimz

    @property
    def dtype(self):
        return np.float32 if self.fp_bytes == 4 else np.float64

    @property
    def n_in(self):
        return int((self.type == NODE_IN).sum())

    @property
    def planes(self):
        """x-planes the node arrays cover (dimx unless the case is slab-local)."""
        return self.dimx if self.x_lo is None else self.x_hi - self.x_lo


def fluid_params(Re=200.0, Pr=0.72, lam=1.4, fp_bytes=8):
    """Common::FluidParams(Re, Pr, lambda) (reference src/Common/Geometry.h:545-552), rounded to FTYPE."""
    ft = np.float32 if fp_bytes == 4 else np.float64
    return dict(v_T=1.0, v_vis=float(ft(1.0 / Re)), t_vis=float(ft(1.0 / (Re * Pr))),
                t_phi=float(ft((lam - 1) / (lam * Re))))


def channel_case(dimx, dimy, dimz, fp_bytes=8, baffle=True, depth_var=0.2, h=None, dt=0.1,
                 num_global=4, num_local=2, Re=200.0, Pr=0.72, lam=1.4, active_dimz=None,
                 inflow=1.0, baseT=1.0, x_range=None) -> Case:
    """Synthetic masked channel: inflow valve at low x, free outflow valve at high x, no-slip side walls,
    optional wall-attached baffle (obstacle; keeps <= 2 segments per row) and the reference's bottom
    perturbation (`depth_var`), extruded along z exactly like Grid3D::Prepare2D.  A NODE_OUT shell is kept
    on every domain face.  Deterministic - no RNG.
    x_range=(lo, hi): only the x-planes [lo, hi) of the same grid (a slab-local case for cmc_adi3d_set_nodes_slab: no
    process has to build the whole grid)."""
    ft = np.float32 if fp_bytes == 4 else np.float64
    if h is None:
        h = 1.1 / max(dimx, dimy, dimz)
    h = float(np.float32(h))       # Config parses every real through float (Config.h:116-121)
    az = active_dimz or dimz
    t2 = np.full((dimx, dimy), NODE_IN, dtype=np.int32)
    vel2 = np.zeros((dimx, dimy), dtype=ft)
    # valves first, walls overwrite the corners (Grid2D::Build raster order, Grid2D.cpp:231-266)
    t2[1, 1:dimy - 1] = NODE_VALVE
    vel2[1, 1:dimy - 1] = inflow
    t2[dimx - 2, 1:dimy - 1] = NODE_VALVE
    t2[1:dimx - 1, 1] = NODE_BOUND
    t2[1:dimx - 1, dimy - 2] = NODE_BOUND
    vel2[:, 1] = 0
    vel2[:, dimy - 2] = 0
    t2[0, :] = NODE_OUT
    t2[dimx - 1, :] = NODE_OUT
    t2[:, 0] = NODE_OUT
    t2[:, dimy - 1] = NODE_OUT
    if baffle and dimx >= 24 and dimy >= 16:
        i0, i1 = int(0.4 * dimx), int(0.6 * dimx)
        j1 = max(4, int(0.3 * dimy))
        t2[i0:i1 + 1, 1:j1 + 1] = NODE_OUT
        t2[i0, 1:j1 + 1] = NODE_BOUND
        t2[i1, 1:j1 + 1] = NODE_BOUND
        t2[i0:i1 + 1, j1] = NODE_BOUND
        t2[i0:i1 + 1, 1] = NODE_BOUND
    ii, jj = np.meshgrid(np.arange(dimx), np.arange(dimy), indexing="ij")
    x = -1 + 2 * ii / dimx
    y = -1 + 2 * jj / dimy
    z = 1.0 - (x * x + y * y) * 0.5
    height = max(az - 2 - 2, 0)
    bottom = 1 + (depth_var * z * height).astype(np.int64)
    # the 2D outline is cheap and always built whole; the extrusion along z covers the requested planes only
    lo, hi = x_range if x_range is not None else (0, dimx)
    t2, vel2, bottom = t2[lo:hi], vel2[lo:hi], bottom[lo:hi]
    px = hi - lo
    N = px * dimy * dimz
    typ = np.full((px, dimy, dimz), NODE_OUT, dtype=np.int32)
    bcv = np.zeros((px, dimy, dimz), dtype=np.int32)
    bct = np.zeros((px, dimy, dimz), dtype=np.int32)
    vx = np.zeros((px, dimy, dimz), dtype=ft)
    T = np.zeros((px, dimy, dimz), dtype=ft)
    kk = np.arange(dimz)[None, None, :]
    col = (t2 != NODE_OUT)[:, :, None]
    bot = bottom[:, :, None]
    is_bottom = col & (kk >= 1) & (kk <= bot)
    is_top = col & (kk == az - 2)
    is_mid = col & (kk > bot) & (kk < az - 2)
    # bottom / top plates: SetBound(BC_NOSLIP, BC_FREE, 0, baseT)
    plate = is_bottom | is_top
    typ[plate] = NODE_BOUND
    bct[plate] = BC_FREE
    T[plate] = baseT
    t3 = np.broadcast_to(t2[:, :, None], typ.shape)
    v3 = np.broadcast_to(vel2[:, :, None], typ.shape)
    m = is_mid & (t3 == NODE_BOUND)
    typ[m] = NODE_BOUND; bct[m] = BC_FREE; T[m] = baseT
    m = is_mid & (t3 == NODE_VALVE) & (v3 == 0)
    typ[m] = NODE_VALVE; bcv[m] = BC_FREE; bct[m] = BC_FREE; T[m] = baseT
    m = is_mid & (t3 == NODE_VALVE) & (v3 != 0)
    typ[m] = NODE_VALVE; vx[m] = v3[m]; T[m] = baseT
    m = is_mid & (t3 == NODE_IN)
    typ[m] = NODE_IN; T[m] = baseT
    p = fluid_params(Re, Pr, lam, fp_bytes)
    zero = np.zeros(N, dtype=ft)
    return Case(dimx, dimy, dimz, h, h, h, p["v_T"], p["v_vis"], p["t_vis"], p["t_phi"], float(ft(dt)),
                num_global, num_local, fp_bytes, type=typ.ravel(), bc_vel=bcv.ravel(), bc_temp=bct.ravel(),
                vx=vx.ravel(), vy=zero, vz=zero.copy(), T=T.ravel(), baseT=baseT,
                x_lo=None if x_range is None else lo, x_hi=None if x_range is None else hi)


def moving_baffle_case(dimx, dimy, dimz, fp_bytes=8, shift=0, **kw) -> Case:
    """The masked channel with its wall-attached baffle displaced by `shift` cells along x: successive shifts are the
    node arrays a moving-boundary case hands to the solver step after step (Grid3D::Prepare(t), reference
    src/FluidSolver3D/Grid3D.cpp:900-945).  Cells the baffle leaves become fluid (NODE_IN), cells it enters become
    NODE_OUT / NODE_BOUND."""
    base = channel_case(dimx, dimy, dimz, fp_bytes=fp_bytes, baffle=False, **kw)
    shp = base.shape
    t = base.type.reshape(shp); bv = base.bc_vel.reshape(shp); bt = base.bc_temp.reshape(shp)
    vx = base.vx.reshape(shp); T = base.T.reshape(shp)
    i0, i1 = int(0.3 * dimx) + shift, int(0.45 * dimx) + shift
    j1 = max(4, int(0.35 * dimy))
    assert i1 < dimx - 3
    fluid_col = (t[i0:i1 + 1, 2:j1 + 1, :] == NODE_IN)                      # fluid cells the baffle covers
    blk = np.zeros(shp, bool); blk[i0:i1 + 1, 2:j1 + 1, :] = True
    inner = np.zeros(shp, bool); inner[i0 + 1:i1, 2:j1, :] = True
    sel = blk & (t == NODE_IN)
    t[sel & inner] = NODE_OUT
    shell = sel & ~inner
    t[shell] = NODE_BOUND; bt[shell] = BC_FREE; T[shell] = base.baseT; vx[shell] = 0
    del fluid_col
    return base


# ---- case files in the reference's formats ---------------------------------------------------------------
BOX_OUTLINE = [  # unit box channel: two passive walls, inflow valve (1 m/s), free outflow valve
    ("Passive", [(0, 0), (1000, 0)], None),
    ("Passive", [(1000, 1000), (0, 1000)], None),
    ("Motion", [(0, 0), (0, 1000)], (1000.0, 0.0)),
    ("Motion", [(1000, 0), (1000, 1000)], (0.0, 0.0)),
]
BAFFLE_OUTLINE = [  # same channel with a wall-attached baffle
    ("Passive", [(0, 0), (400, 0), (400, 300), (600, 300), (600, 0), (1000, 0)], None),
    ("Passive", [(1000, 1000), (0, 1000)], None),
    ("Motion", [(0, 0), (0, 1000)], (1000.0, 0.0)),
    ("Motion", [(1000, 0), (1000, 1000)], (0.0, 0.0)),
]


# Four one-millimetre passive marks well outside the channel.  They only widen the bounding box so that a
# coarse grid (grid_d > 2 % of the extent, the loader's fixed padding - Geometry.h:471-478) still gets an
# OUT rim around the geometry; each mark rasterises to an isolated BOUND cell inside the OUT region.
RIM_MARKS = [
    ("Passive", [(-150, 500), (-150, 501)], None),
    ("Passive", [(1150, 500), (1150, 501)], None),
    ("Passive", [(500, -150), (501, -150)], None),
    ("Passive", [(500, 1150), (501, 1150)], None),
]


def write_shape2d_case(directory, name, outline=None, grid_d=0.02, depth=1.0, depth_var=0.0, duration=10.0,
                       time_steps=100, num_global=4, num_local=2, Re=200.0, Pr=0.72, lam=1.4,
                       out_grid=(16, 16, 16), out_time_steps=10, solver="ADI", dimension="3D", rim=False):
    """Write `<name>_data.txt` (Shape2D: frames / shapes / points in millimetres, `Passive` or `Motion vx vy`)
    and `<name>_config.txt` (whitespace `key value` pairs, reference src/Common/Config.h:203-245).
    Returns (data_path, config_path)."""
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    outline = list(outline or BOX_OUTLINE) + (RIM_MARKS if rim else [])
    lines = ["1", f"{duration}", f"{len(outline)}"]
    for kind, pts, vel in outline:
        lines.append(str(len(pts)))
        lines += [f"{float(x)} {float(y)}" for x, y in pts]
        lines.append(kind)
        if kind == "Motion":
            lines.append(f"{vel[0]} {vel[1]}")
    data = directory / f"{name}_data.txt"
    data.write_text("\n".join(lines) + "\n")
    cfg = [
        f"dimension\t{dimension}", "in_fmt\t\tShape2D", f"depth\t\t{depth}", f"depth_var\t{depth_var}",
        f"Re\t\t{Re}", f"Pr\t\t{Pr}", f"lambda\t\t{lam}", "bc_type\t\tNoSlip",
        f"grid_dx\t\t{grid_d}", f"grid_dy\t\t{grid_d}", f"grid_dz\t\t{grid_d}",
        "cycles\t\t1", f"time_steps\t{time_steps}", "out_fmt\t\tNetCDF", f"out_time_steps\t{out_time_steps}",
        f"out_gridx\t{out_grid[0]}", f"out_gridy\t{out_grid[1]}", f"out_gridz\t{out_grid[2]}",
        "out_vars\t4 u v w T", f"solver\t\t{solver}", f"num_global\t{num_global}", f"num_local\t{num_local}",
    ]
    config = directory / f"{name}_config.txt"
    config.write_text("\n".join(cfg) + "\n")
    return data, config
