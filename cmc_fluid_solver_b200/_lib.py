"""ctypes binding of include/cmc_adi.h (libcmcadi.so).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libcmcadi.so"

# every symbol include/cmc_adi.h declares (checked by tests/test_abi.py against the header text)
SYMBOLS = [
    "cmc_last_error", "cmc_abi_version", "cmc_device_count",
    "cmc_adi3d_create", "cmc_nccl_unique_id", "cmc_adi3d_create_dist", "cmc_adi3d_create_emulated", "cmc_adi3d_create_multi", "cmc_adi3d_create_ex", "cmc_split_planes", "cmc_adi3d_destroy", "cmc_adi3d_slab",
    "cmc_adi3d_set_nodes", "cmc_adi3d_set_nodes_aos", "cmc_adi3d_set_nodes_slab", "cmc_adi3d_update_nodes", "cmc_adi3d_update_nodes_aos", "cmc_adi3d_build_lines", "cmc_adi3d_num_segments",
    "cmc_adi3d_update_boundaries", "cmc_adi3d_time_step", "cmc_adi3d_get_layer", "cmc_adi3d_get_layer_async", "cmc_adi3d_get_layer_wait", "cmc_adi3d_output_rows",
    "cmc_adi3d_write_layer_async", "cmc_adi3d_write_layer_commit",
    "cmc_adi3d_set_option", "cmc_adi3d_get_option",
    "cmc_adi3d_read_field", "cmc_adi3d_write_field", "cmc_adi3d_step_prologue", "cmc_adi3d_solve_direction",
    "cmc_adi3d_eval_div_error", "cmc_adi3d_field_sums", "cmc_adi3d_time_step_async", "cmc_adi3d_sync", "cmc_adi3d_stream",
    "cmc_adi3d_launch_count", "cmc_adi3d_get_timing", "cmc_adi3d_device_bytes", "cmc_solve_tridiagonal_batch",
    "cmc_adi2d_create", "cmc_adi2d_destroy", "cmc_adi2d_set_grid", "cmc_adi2d_init_layer", "cmc_adi2d_update_boundaries",
    "cmc_adi2d_time_step", "cmc_adi2d_step_host", "cmc_adi2d_time_step_batch", "cmc_adi2d_get_layer", "cmc_adi2d_read_field", "cmc_adi2d_write_field", "cmc_adi2d_launch_count",
]


class GridDesc(C.Structure):
    _fields_ = [("dimx", C.c_int32), ("dimy", C.c_int32), ("dimz", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double)]


class Decomp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("device", C.c_int32), ("n_slabs", C.c_int32), ("rank", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("devices", C.POINTER(C.c_int32)), ("planes", C.POINTER(C.c_int32))]


class FluidParams(C.Structure):
    _fields_ = [("v_T", C.c_double), ("v_vis", C.c_double), ("t_vis", C.c_double), ("t_phi", C.c_double)]


_lib = None


def load_library() -> C.CDLL:
    """Load libcmcadi.so from the package directory (built by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C cmc_fluid_solver_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    lib.cmc_last_error.restype = C.c_char_p
    lib.cmc_last_error.argtypes = []
    sig = {
        "cmc_abi_version": [],
        "cmc_device_count": [],
        "cmc_adi3d_create": [P(GridDesc), P(FluidParams), i32, i32, P(vp)],
        "cmc_nccl_unique_id": [vp],
        "cmc_adi3d_create_dist": [P(GridDesc), P(FluidParams), i32, i32, i32, i32, vp, P(vp)],
        "cmc_adi3d_create_emulated": [P(GridDesc), P(FluidParams), i32, i32, i32, P(vp)],
        "cmc_adi3d_create_multi": [P(GridDesc), P(FluidParams), i32, P(i32), i32, P(vp)],
        "cmc_adi3d_create_ex": [P(GridDesc), P(FluidParams), i32, P(Decomp), P(vp)],
        "cmc_split_planes": [i32, P(GridDesc), vp, i32, P(C.c_int32)],
        "cmc_adi3d_destroy": [vp],
        "cmc_adi3d_slab": [vp, P(i32), P(i32)],
        "cmc_adi3d_set_nodes": [vp, vp, vp, vp, vp, vp, vp, vp],
        "cmc_adi3d_set_nodes_aos": [vp, vp, C.c_size_t],
        "cmc_adi3d_set_nodes_slab": [vp, vp, vp, vp, vp, vp, vp, vp, i32],
        "cmc_adi3d_update_nodes": [vp, vp, vp, vp, vp, vp, vp, vp],
        "cmc_adi3d_update_nodes_aos": [vp, vp, C.c_size_t],
        "cmc_adi3d_build_lines": [vp],
        "cmc_adi3d_num_segments": [vp, i32, P(i64)],
        "cmc_adi3d_update_boundaries": [vp],
        "cmc_adi3d_time_step": [vp, dbl, i32, i32, i32, P(dbl)],
        "cmc_adi3d_get_layer": [vp, vp, vp, i32, i32, i32],
        "cmc_adi3d_get_layer_async": [vp, vp, vp, i32, i32, i32],
        "cmc_adi3d_get_layer_wait": [vp],
        "cmc_adi3d_write_layer_async": [vp, vp, vp, vp, vp],
        "cmc_adi3d_write_layer_commit": [vp, i32],
        "cmc_adi3d_set_option": [vp, C.c_char_p, i64],
        "cmc_adi3d_get_option": [vp, C.c_char_p, P(i64)],
        "cmc_adi3d_read_field": [vp, i32, i32, vp],
        "cmc_adi3d_write_field": [vp, i32, i32, vp],
        "cmc_adi3d_step_prologue": [vp],
        "cmc_adi3d_solve_direction": [vp, i32, dbl, i32, i32, i32],
        "cmc_adi3d_eval_div_error": [vp, i32, P(dbl)],
        "cmc_adi3d_field_sums": [vp, i32, P(dbl)],
        "cmc_adi3d_time_step_async": [vp, dbl, i32, i32, i32],
        "cmc_adi3d_sync": [vp, P(dbl)],
        "cmc_adi3d_stream": [vp, P(vp)],
        "cmc_adi3d_launch_count": [vp, P(i64), i32],
        "cmc_adi3d_get_timing": [vp, i32, P(dbl), P(i64)],
        "cmc_adi3d_device_bytes": [vp, P(i64)],
        "cmc_adi3d_output_rows": [vp, i32, P(i32), P(i32)],
        "cmc_solve_tridiagonal_batch": [i32, i32, i32, i32, vp, vp, vp, vp, vp],
        "cmc_adi2d_create": [i32, i32, dbl, dbl, P(FluidParams), dbl, i32, i32, P(vp)],
        "cmc_adi2d_destroy": [vp],
        "cmc_adi2d_set_grid": [vp, vp, vp, vp, vp, vp],
        "cmc_adi2d_init_layer": [vp],
        "cmc_adi2d_update_boundaries": [vp],
        "cmc_adi2d_time_step": [vp, dbl, i32, i32, P(dbl), P(i32)],
        "cmc_adi2d_step_host": [vp, vp, vp, vp, vp, vp, P(vp), P(vp), dbl, i32, i32, P(dbl), P(i32)],
        "cmc_adi2d_time_step_batch": [P(vp), i32, dbl, i32, i32, i32, P(dbl), P(i32), P(i32)],
        "cmc_adi2d_get_layer": [vp, vp, vp, i32, i32],
        "cmc_adi2d_read_field": [vp, i32, i32, vp],
        "cmc_adi2d_write_field": [vp, i32, i32, vp],
        "cmc_adi2d_launch_count": [vp, P(i64)],
    }
    for name, argtypes in sig.items():
        f = getattr(lib, name)
        f.argtypes = argtypes
        f.restype = C.c_int
    _lib = lib
    return lib
