#!/usr/bin/env python
"""bench.py - the ADI time-step hot path on B200, measured against the HBM roofline.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU solver on the host cores

A "step" is one pass of the hot path: UpdateBoundaries + TimeStep(dt, num_global=4, num_local=2,
computeError every 10th step) - the reference driver's loop body (FluidSolver3D.cpp:241-242) - on the 512^3
masked synthetic channel (BASELINE.json: "3D 512^3 masked grid"; the case the north_star target is quoted on).
N > 1 (torchrun, one rank per GPU) splits the same grid into x-slabs: strong scaling.

One JSON line on rank 0; see DESIGN.md "Measurement" for the byte model behind `roofline`.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mcell-updates/s per 3D time step"
UNIT = "Mcell-updates/s"
NUM_GLOBAL, NUM_LOCAL = 4, 2          # the reference's example configs (data/3D/example_tests/*_config.txt)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- reference CPU arm
def run_reference_cpu(steps, warmup, size=128, fp_bytes=8):
    """The reference's own CPU solver (oracle/_ref/ref_probe3d_*, built from the unmodified reference sources) on a
    bounded sample of the workload: the same masked channel family at size^3 read by the reference's own loader."""
    from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case
    from oracle import oracle as O
    if not O.have_ref(fp_bytes):
        return None
    grid_d = {64: 0.02, 128: 0.0095, 256: 0.0045}.get(size, 1.2 / size)
    with tempfile.TemporaryDirectory() as td:
        data, cfg = write_shape2d_case(td, "bench", outline=BAFFLE_OUTLINE, grid_d=grid_d, depth=1.0 if size <= 128 else 1.14,
                                       depth_var=0.2, time_steps=100, num_global=NUM_GLOBAL, num_local=NUM_LOCAL)
        out = O.run_ref(data, cfg, "-", steps + warmup, fp_bytes=fp_bytes, align=True, dump="none")
    m = re.search(r"grid (\d+) x (\d+) x (\d+), NODE_IN (\d+).*threads (\d+)", out)
    dims = tuple(int(m.group(i)) for i in (1, 2, 3))
    threads = int(m.group(5))
    times = [float(x) for x in re.findall(r"probe: step \d+ seconds ([0-9.]+)", out)]
    timed = times[warmup:]
    ncells = dims[0] * dims[1] * dims[2]
    sec = sum(timed) / max(len(timed), 1)
    return dict(value=ncells / sec / 1e6, sec_per_step=sec, dims=dims, cores=threads, steps=len(timed),
                fluid_fraction=int(m.group(4)) / ncells)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_reference_cpu(args.steps, args.warmup, size=args.ref_size)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_probe3d_f64 not built (needs /root/reference at build time)"}))
        return
    sample = (f"{r['dims'][0]}x{r['dims'][1]}x{r['dims'][2]} masked channel (baffle + bottom perturbation), fp64, "
              f"{r['steps']} timed TimeSteps of the reference CPU/OpenMP solver (unmodified sources, g++ -O2 -fopenmp)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "num_global": NUM_GLOBAL, "num_local": NUM_LOCAL,
                   "note": "each step is a bounded sample of the workload: " + sample},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args):
    s = args.size
    return f"3D {s}^3 masked channel (wall-attached baffle + depth_var 0.2 bottom), fp{args.fp * 8}, ADI step num_global {NUM_GLOBAL} num_local {NUM_LOCAL}"


# ------------------------------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dims", default=None, help="X,Y,Z grid instead of --size^3 (experiments)")
    ap.add_argument("--fp", type=int, default=8, choices=[4, 8])
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-size", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    from cmc_fluid_solver_b200 import AdiSolver3D
    from cmc_fluid_solver_b200.cases import channel_case

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ADI path has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from cmc_fluid_solver_b200.solver import nccl_unique_id
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    S = args.size
    DX, DY, DZ = (int(v) for v in args.dims.split(",")) if args.dims else (S, S, S)
    case = channel_case(DX, DY, DZ, fp_bytes=args.fp, depth_var=0.2)
    ncells = case.ncells
    fluid = case.n_in / ncells
    sol = AdiSolver3D().Init(case, device=local_rank, mode=args.mode, rank=rank, nranks=world, nccl_id=nccl_id)
    sol.CreateSegments()
    stream = torch.cuda.ExternalStream(sol.stream(), device=local_rank)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        sol.UpdateBoundaries()
        sol.TimeStepAsync(case.dt, NUM_GLOBAL, NUM_LOCAL, i % 10 == 0)

    for i in range(args.warmup):
        step(i)
    sol.Sync()

    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the solver's stream -------------------------
    sol.set_profile(True, reset=True)
    sol.launch_count(reset=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    err = sol.Sync()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = sol.launch_count()
    timings = sol.timings()
    sol.set_profile(False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = ncells * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (directional sweep): algorithmic bytes per launch / avg launch duration ----
    fpb = args.fp
    local_cells = sol.nx * case.dimy * case.dimz
    sweep_bytes = local_cells * (16 * fpb + 1)             # read cur x4 + temp x4, write next x4 + temp' x4, 1 descriptor byte
    peak, peak_src = peaks()
    per_dir = {}
    for k in ("sweep_x", "sweep_y", "sweep_z"):
        tot, n = timings[k]
        if n:
            per_dir[k] = {"ms_per_launch": tot / n, "launches": n, "gbs": sweep_bytes / (tot / n * 1e-3) / 1e9}
    dom = max(per_dir, key=lambda k: per_dir[k]["ms_per_launch"] * per_dir[k]["launches"]) if per_dir else None
    sweep_ms = sum(timings[k][0] for k in ("sweep_x", "sweep_y", "sweep_z"))
    sweep_n = sum(timings[k][1] for k in ("sweep_x", "sweep_y", "sweep_z"))
    achieved = sweep_bytes * sweep_n / (sweep_ms * 1e-3) / 1e9 if sweep_ms else None
    step_bytes = local_cells * (NUM_GLOBAL * 3 * NUM_LOCAL * (16 * fpb + 1) + NUM_GLOBAL * 12 * fpb + 8 * fpb)   # BASELINE.md B_step
    # measured DRAM traffic per launch of the dominant kernel: one ncu --set full capture, committed under profiles/
    traffic = None
    tp = ROOT / "profiles" / "r01_ncu_traffic.json"
    key = f"{DX}x{DY}x{DZ}_f{fpb * 8}"
    if tp.exists() and world == 1 and dom:
        traffic = json.loads(tp.read_text()).get(key, {}).get(dom, {}).get("dram_bytes")
    roofline = {
        "bound": "hbm", "kernel": "k_fast_sweep<%s,X|Y|Z> (all three directions, %d launches)" % ("double" if fpb == 8 else "float", sweep_n),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
        "traffic_source": "profiles/r01_ncu_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch of the dominant kernel)" if traffic else None,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": sweep_bytes,
        "per_direction": per_dir, "dominant": dom,
        "sweep_share_of_step": sweep_ms / (ms if world == 1 else max(ms, 1e-9)),
        "step_achieved_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9, "step_frac": step_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
        "kernel_ms": {k: v[0] / args.steps for k, v in timings.items() if v[1]},
    }

    # ---- e2e: the same step through the public API with HOST buffers (H2D state, D2H layer) inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # every rank feeds its own slab from pinned host memory; rank 0 receives the full layer (GetLayer gathers)
        ft = torch.float32 if fpb == 4 else torch.float64
        host_in = [torch.empty(local_cells, dtype=ft).pin_memory() for _ in range(4)]
        for q in range(4):
            host_in[q].numpy()[:] = sol.read_field(0, q).ravel()
        host_vel = torch.empty(ncells * 3 if rank == 0 else 3, dtype=ft).pin_memory()
        host_T = torch.empty(ncells if rank == 0 else 1, dtype=torch.float64).pin_memory()
        h2d = 4 * ncells * fpb                                   # all ranks together
        d2h = ncells * (3 * fpb + 8) + 16

        def e2e_step():
            for q in range(4):
                sol.write_field(0, q, host_in[q].numpy())            # host -> device: the layer the caller owns
            sol.UpdateBoundaries()
            sol.TimeStep(case.dt, NUM_GLOBAL, NUM_LOCAL, True)       # residual read back
            sol.GetLayer(0, 0, 0, vel=host_vel.numpy().reshape(-1, 3), T=host_T.numpy())   # device -> host: full layer

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": ncells * args.e2e_steps / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms / args.e2e_steps, "steps": args.e2e_steps,
               "what": "write_field x4 (pinned host -> HBM, every rank its slab), UpdateBoundaries, TimeStep(computeError), GetLayer full "
                       "resolution (HBM -> pinned host on rank 0); host wall clock between barriers, max over ranks"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference_cpu(3, 1, size=args.ref_size)
            if r:
                cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                       "sample": f"{r['dims'][0]}x{r['dims'][1]}x{r['dims'][2]} masked channel of the same family, fp64, {r['steps']} timed steps "
                                 f"({r['sec_per_step']:.2f} s/step) of the reference CPU/OpenMP solver built from its unmodified sources"}
        except Exception as ex:      # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if fpb == 8 else "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "grid": [DX, DY, DZ], "num_global": NUM_GLOBAL, "num_local": NUM_LOCAL,
                       "sweeps_per_step": NUM_GLOBAL * 3 * NUM_LOCAL, "fluid_fraction": round(fluid, 4), "mode": args.mode,
                       "parallelism": f"x-slab x{world}", "exchange": sol.exchange_kind(), "storage": f"SoA, y-blocked ({sol.storage_block_rows()} rows per block)" if sol.storage_block_rows() else "SoA [i][j][k]", "residual": "every 10th step (reference driver cadence)",
                       "l2": f"inputs larger than L2: each field {ncells * fpb / 1e9:.2f} GB, 20 resident fields"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk, "residual": err, "device_bytes": sol.device_bytes(),
        }
        print(json.dumps(line))
    sol.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
