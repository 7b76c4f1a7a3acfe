"""The REAL multi-rank path (one process per GPU, NCCL rendezvous, CUDA-IPC peer memory, epoch flags) under pytest:
launches tools/dist_check.py with torchrun on 2 ranks, once with the fused peer-memory exchange and once with the NCCL
send/recv transport (CMC_P2P=0).  Every rank compares its planes, the residual, the checksums and the GetLayer output with
the same case solved as one slab on its own GPU.  Needs 2 visible GPUs (skipped otherwise: ranks that wait on each other's
flags must not share one device, B200_PROFILING.md)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_ranks_against_single_gpu(p2p):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, CMC_P2P=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if p2p == "1" else "29534", str(ROOT / "tools" / "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and "DIST CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    want = "fused-stores-peer-memory" if p2p == "1" else "nccl"
    assert f"exchange {want}" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("fp", [8, 4])
def test_one_process_two_devices_against_single_gpu(oracle_mod, fp):
    """cmc_adi3d_create_multi: ONE process drives two devices (the reference's "GPU 2" mode, FluidSolver3D.cpp:88-95) -
    a slab per device, own streams, peer-access stores, event ordering - against the oracle and the one-device run."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import numpy as np
    from conftest import assert_fields_close
    from cmc_fluid_solver_b200 import AdiSolver3D
    from cmc_fluid_solver_b200.cases import channel_case
    O = oracle_mod
    case = channel_case(64, 40, 48, fp_bytes=fp, depth_var=0.25)
    case.outdims = (9, 7, 5)
    ora = O.Oracle3D(case); ora.create_segments()
    one = AdiSolver3D().Init(case, mode="fast"); one.CreateSegments()
    two = AdiSolver3D().Init(case, mode="fast", devices=[0, 1]); two.CreateSegments()
    assert two.exchange_kind() == "fused-stores-peer-access"
    assert [two.numSegs(d) for d in range(3)] == [len(ora.segments(d)) for d in range(3)]
    for i in range(5):
        ora.update_boundaries(); one.UpdateBoundaries(); two.UpdateBoundaries()
        e_ref = ora.time_step(case.dt, case.num_global, case.num_local, True)
        e1 = one.TimeStep(case.dt, case.num_global, case.num_local, True)
        e = two.TimeStep(case.dt, case.num_global, case.num_local, True)
        assert abs(e - e_ref) <= (1e-5 if fp == 4 else 1e-9) * abs(e_ref) and abs(e - e1) <= (1e-5 if fp == 4 else 1e-9) * abs(e1)
        if i == 2:
            v_ref, T_ref = ora.get_layer(*case.outdims)
            v, T = two.GetLayer(*case.outdims)
            tol = 1e-5 if fp == 4 else 1e-10
            assert np.allclose(v, v_ref, rtol=0, atol=tol * 1e5) and np.allclose(T, T_ref, rtol=0, atol=tol * 1e5)
            one.GetLayer(*case.outdims)
        assert_fields_close([ora.field(O.LAYER_CUR, q) for q in range(4)], [two.read_field(0, q) for q in range(4)], fp, f"2 devices, step {i}")
        assert_fields_close([one.read_field(0, q) for q in range(4)], [two.read_field(0, q) for q in range(4)], fp, "1 device vs 2")
    s1, s2 = one.field_sums(0), two.field_sums(0)
    for n in "uvwT":
        assert abs(s1[n][0] - s2[n][0]) <= (1e-5 if fp == 4 else 1e-10) * max(abs(s1[n][1]) ** 0.5, 1.0) * 1e3
    one.close(); two.close()


def test_reference_driver_on_two_devices(oracle_mod, tmp_path):
    """The drop-in boundary with the reference's "GPU 2": reference loader + Solver3D adapter (n devices) + libcmcadi.so
    against the reference CPU solver on the same case files."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import numpy as np
    from conftest import assert_fields_close
    from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case
    O = oracle_mod
    REF = ROOT / "oracle" / "_ref"
    dropin, probe = REF / "dropin3d_f64", REF / "ref_probe3d_f64"
    if not (dropin.exists() and probe.exists()):
        pytest.skip("oracle/_ref/dropin3d_* not built (needs /root/reference at build time)")
    data, cfg = write_shape2d_case(tmp_path, "case", outline=BAFFLE_OUTLINE, grid_d=0.0187, depth_var=0.2, time_steps=300, out_grid=(20, 21, 19))
    outs = {}
    for name, binary, extra in (("cpu", probe, ["solver=cpu"]), ("gpu2", dropin, ["solver=b200", "gpus=2"])):
        out = tmp_path / f"{name}.bin"
        r = subprocess.run([str(binary), str(data), str(cfg), str(out), "6", "align", "dump=last"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = O.read_probe(out)
        if name == "gpu2":
            assert "2 devices" in r.stdout, r.stdout[-1500:]
    a, b = outs["cpu"].snapshots[-1], outs["gpu2"].snapshots[-1]
    assert outs["cpu"].shape[0] % 16 == 0, outs["cpu"].shape
    assert_fields_close([a[n] for n in "uvwT"], [b[n] for n in "uvwT"], 8, "reference CPU vs adapter on 2 devices")


@pytest.mark.parametrize("fp", ["8", "4"])
def test_one_pass_coupled_x_sweep_slabs_sharing_one_gpu(fp):
    """The one-pass slab-coupled x-sweep (kernels_tma.cu XS, option "xs" / CMC_XS=1, off by default): 2, 3, 4 and 8 slabs with
    their own streams on ONE device (CMC_SHARE_DEVICE=1: the slabs' kernels wait for each other's words, so each gets a share
    of the SMs) against the oracle.  In its own process: a peer that never arrives ends in a trap, not in a hang."""
    env = dict(os.environ, CMC_SHARE_DEVICE="1", CMC_XS="1")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "xs_check.py"), fp], capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    if r.returncode != 0 and "FAILED" not in r.stdout and "AssertionError" not in r.stderr:
        # the slabs' kernels wait for each other: when a device does not run them side by side (hardware queue mapping of the
        # streams), the wait ends in the kernel's trap after two seconds - an environment condition, not a wrong result
        pytest.skip("the slabs' kernels did not run concurrently on this device: " + r.stderr[-300:])
    assert r.returncode == 0 and "XS CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_one_pass_coupled_x_sweep_two_ranks():
    """The same kernel between two real GPUs (one process per GPU, words stored into the peer's table over NVLink)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, CMC_P2P="1", CMC_XS="1", DIST_CHECK_PLANES="64")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29536", str(ROOT / "tools" / "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and "DIST CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "kernel_x 5" in r.stdout, r.stdout[-2000:]
