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
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_mem_gb():
    try:
        for ln in Path("/proc/meminfo").read_text().splitlines():
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# the masked channel family in the reference's own case format (Shape2D outline + depth_var bottom, `align`), per size
REF_GRID = {64: dict(grid_d=0.02, depth=1.0), 128: dict(grid_d=0.0095, depth=1.0), 256: dict(grid_d=0.0045, depth=1.14),
            512: dict(grid_d=0.00215, depth=1.09)}


def pick_ref_size(want, budget_s, steps):
    """Largest grid of the family (<= the workload's) whose `steps` reference steps fit the time budget on this host:
    the reference CPU solver does about 0.27 Mcell-updates/s per core (measured: 4.4 at 16 cores, 6.4 at 32) and needs
    about 280 bytes of host memory per cell."""
    cores, mem = host_cores(), host_mem_gb()
    for size in (512, 256, 128, 64):
        if size > want:
            continue
        cells = size ** 3
        if cells * steps / (0.27e6 * cores) <= budget_s and cells * 280 / 1e9 <= 0.8 * mem:
            return size
    return 64


def run_reference_cpu(steps, warmup, size=128, fp_bytes=8, threads=None):
    """The reference's own CPU solver (oracle/_ref/ref_probe3d_*, built from the unmodified reference sources) on a
    bounded sample of the workload: the same masked channel family at size^3 read by the reference's own loader, on
    all host cores (the thread count is passed explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    from cmc_fluid_solver_b200.cases import BAFFLE_OUTLINE, write_shape2d_case
    from oracle import oracle as O
    if not O.have_ref(fp_bytes):
        return None
    threads = threads or host_cores()
    kw = REF_GRID.get(size, dict(grid_d=1.2 / size, depth=1.0))
    with tempfile.TemporaryDirectory() as td:
        data, cfg = write_shape2d_case(td, "bench", outline=BAFFLE_OUTLINE, depth_var=0.2, time_steps=100, num_global=NUM_GLOBAL,
                                       num_local=NUM_LOCAL, **kw)
        out = O.run_ref(data, cfg, "-", steps + warmup, fp_bytes=fp_bytes, align=True, dump="none", threads=threads)
    m = re.search(r"grid (\d+) x (\d+) x (\d+), NODE_IN (\d+).*threads (\d+)", out)
    dims = tuple(int(m.group(i)) for i in (1, 2, 3))
    used = int(m.group(5))
    times = [float(x) for x in re.findall(r"probe: step \d+ seconds ([0-9.]+)", out)]
    timed = times[warmup:]
    ncells = dims[0] * dims[1] * dims[2]
    sec = sum(timed) / max(len(timed), 1)
    return dict(value=ncells / sec / 1e6, sec_per_step=sec, dims=dims, cores=used, steps=len(timed),
                fluid_fraction=int(m.group(4)) / ncells)


def cpu_sample_text(r):
    return (f"{r['dims'][0]}x{r['dims'][1]}x{r['dims'][2]} masked channel (wall-attached baffle + depth_var 0.2 bottom, the workload's "
            f"family read by the reference's own loader), fp64, {r['steps']} timed TimeSteps ({r['sec_per_step']:.2f} s/step) of the "
            f"reference CPU/OpenMP solver built from its unmodified sources (g++ -O2 -fopenmp), {r['cores']} threads")


def reference_arm(args):
    """bench.py --impl reference: rank 0 alone runs the reference CPU solver on all host cores; the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" is a bounded sample: at most 2 timed steps (+1 warm-up) of the largest grid of the workload's family
    # that fits about two minutes on this host - the 512^3 workload itself on a 16-core box
    timed = max(1, min(args.steps, 2))
    size = args.ref_size or pick_ref_size(args.size, 150.0, timed + 1)
    r = run_reference_cpu(timed, 1, size=size)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_probe3d_f64 not built (needs /root/reference at build time)"}))
        return
    sample = cpu_sample_text(r)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world=args.gpus),
        "sample": sample, "sample_grid": list(r["dims"]), "steps_timed": r["steps"], "warmup_run": 1,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def grid_dims(args, world):
    if args.dims:
        return tuple(int(v) for v in args.dims.split(","))
    if args.scaling == "weak":      # BASELINE config 5: 128 x 512 x 512 per GPU => 1024 x 512 x 512 at 8 GPUs
        return (128 * world, 512, 512)
    return (args.size, args.size, args.size)


def workload_name(args, world=1):
    if args.scaling == "weak":
        return (f"3D weak scaling, 128x512x512 masked channel cells per GPU (1024x512x512 at 8 GPUs; wall-attached baffle + depth_var 0.2 "
                f"bottom), fp{args.fp * 8}, ADI step num_global {NUM_GLOBAL} num_local {NUM_LOCAL}")
    s = args.size
    return f"3D {s}^3 masked channel (wall-attached baffle + depth_var 0.2 bottom), fp{args.fp * 8}, ADI step num_global {NUM_GLOBAL} num_local {NUM_LOCAL}"


def bench_config(args, world):
    """The `config` object of the JSON line: identical for both arms (it names the workload, not the implementation)."""
    DX, DY, DZ = grid_dims(args, world)
    return {"workload": workload_name(args, world), "grid": [DX, DY, DZ], "num_global": NUM_GLOBAL, "num_local": NUM_LOCAL,
            "sweeps_per_step": NUM_GLOBAL * 3 * NUM_LOCAL, "parallelism": f"x-slab x{world}",
            "residual": "every 10th step (reference driver cadence)",
            "l2": f"inputs larger than L2: each field {DX * DY * DZ * args.fp / 1e9:.2f} GB, 20 resident fields"}


# --------------------------------------------------------------------------------- BASELINE config 1 (2D)
def bench_2d(args):
    """`bench.py --config 2d`: the reference's own 2D case data/2D/box_pipe (120 x 135, fp32, 49 steps - BASELINE config 1)
    through cmc_adi2d_*: `--batch` independent copies of the case advance per launch (one thread block per case; the
    2D solver is bit-identical with the reference, so every copy reproduces the golden residuals).  CPU baseline: the
    oracle's C restatement of AdiSolver2D (pinned bit-for-bit to the reference; the reference's 2D solver is serial)."""
    import numpy as np
    import torch
    sys.path.insert(0, str(ROOT / "tests"))
    from test_oracle2d import golden_grid, load_golden2d
    from cmc_fluid_solver_b200 import AdiSolver2D
    from cmc_fluid_solver_b200.solver import time_step_batch_2d
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ADI path has no CPU fallback")
    z = load_golden2d()
    dimx, dimy = (int(v) for v in z["dims"])
    steps, dt, ng, nl = int(z["steps"]), float(z["dt"]), int(z["iters"][0]), int(z["iters"][1])
    B = args.batch

    def make(n):
        out = []
        for _ in range(n):
            s = AdiSolver2D().Init(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
            for q in range(3):
                s.write_field(0, q, z["layer_init"][q])
            s.set_grid(*golden_grid(z, 0))
            out.append(s)
        return out

    def run(solvers):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        errs = None
        for _ in range(steps):
            errs, _it = time_step_batch_2d(solvers, dt, ng, nl, update_boundaries=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, errs

    run(make(2))                                   # warm-up (module load, first launches)
    one = make(1)
    t1, e1 = run(one)
    many = make(B)
    tb, eb = run(many)
    assert all(e == eb[0] for e in eb), "the copies of a batch diverged from each other"
    # the case with a static grid (Prepare(0) only) for all 49 steps, against the oracle doing the same
    from oracle import oracle as O
    O.build()
    o = O.Oracle2D(dimx, dimy, *[float(v) for v in z["spacing"]], *[float(v) for v in z["params"]], float(z["startT"]), 4)
    for q in range(3):
        o.field(0, q)[:] = z["layer_init"][q]
    o.set_grid(*golden_grid(z, 0))
    t0 = time.perf_counter()
    for _ in range(steps):
        o.update_boundaries(); e_ref = o.time_step(dt, ng, nl)
    tc = time.perf_counter() - t0
    cells = dimx * dimy
    line = {
        "metric": "Mcell-updates/s per 2D time step", "value": B * cells * steps / tb / 1e6, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": 1,
        "ms_per_step": tb / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "reference case data/2D/box_pipe (static grid)",
        "config": {"workload": f"2D ADI, data/2D/box_pipe {dimx}x{dimy}, fp32, {steps} steps, num_global {ng} num_local {nl} (BASELINE config 1), {B} independent copies per launch",
                   "batch": B},
        "single_case": {"value": cells * steps / t1 / 1e6, "ms_per_step": t1 / steps * 1e3},
        "parity": {"residual_last_step": eb[0], "oracle_residual_last_step": e_ref, "bit_identical": bool(eb[0] == e_ref and e1[0] == e_ref)},
        "cpu_baseline": {"value": cells * steps / tc / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"oracle/adi2d_oracle.c (C restatement of the reference's serial AdiSolver2D, pinned bit-for-bit), the same {steps} steps"},
        "gpu_launches": sum(s.launch_count() for s in many[:1]),
    }
    print(json.dumps(line))
    for s in one + many:
        s.close()
    o.close()


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
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-size", type=int, default=0, help="grid of the reference arm's sample (0 = the largest of 512/256/128 that fits ~2 minutes)")
    ap.add_argument("--cpu-size", type=int, default=128, help="grid of the cpu_baseline leg printed with our own arm")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="weak = BASELINE config 5: 128x512x512 cells per GPU")
    ap.add_argument("--config", default="3d", choices=["3d", "2d"], help="2d = BASELINE config 1 (data/2D/box_pipe through cmc_adi2d_*)")
    ap.add_argument("--batch", type=int, default=296, help="--config 2d: independent cases per launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.config == "2d":
        bench_2d(args)
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

    DX, DY, DZ = grid_dims(args, world)
    if world > 1:
        # every rank builds only its own x-planes (+ 2 halo planes either side): cmc_adi3d_set_nodes_slab
        from cmc_fluid_solver_b200.solver import default_split, slab_window
        x0, nxl = default_split(DX, DY, DZ, world)[rank]
        lo, hi = slab_window(DX, x0, nxl)
        case = channel_case(DX, DY, DZ, fp_bytes=args.fp, depth_var=0.2, x_range=(lo, hi))
        n_in = int((case.type.reshape(hi - lo, DY, DZ)[x0 - lo:x0 - lo + nxl] == 0).sum())
        import torch.distributed as dist
        tn = torch.tensor([n_in], dtype=torch.float64, device="cuda")
        dist.all_reduce(tn)
        n_in = int(tn.item())
    else:
        case = channel_case(DX, DY, DZ, fp_bytes=args.fp, depth_var=0.2)
        n_in = case.n_in
    ncells = case.ncells
    fluid = n_in / ncells
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

    if os.environ.get("CMC_XS_TRACE") and rank == 0:      # (instrumented kernel builds only)
        ph = [sol.get_option(f"tilectr{i}") for i in range(12, 23)]
        nt = sol.get_option("tilectr8")
        if nt:
            names = ["-", "load+eliminate uvw", "reduced solve uvw", "send uvw", "wait uvw", "interface uvw", "back-subst + stores uvw", "diss + eliminate T", "reduced solve T", "send+wait+interface T", "back-subst + stores T"]
            print("[phase trace x] cycles per tile: " + ", ".join(f"{n} {64 * v / nt:.0f}" for n, v in zip(names[1:], ph[1:])), file=sys.stderr)
        c = [sol.get_option(f"tilectr{i}") for i in range(2, 12)]
        for d, o in (("x", 4), ("y", 7)):
            if c[o + 2]:
                print(f"[tile trace] {d}: per tile, cycles: whole tile {64 * c[o] / c[o + 2]:.0f}, of which waiting for the tile's copies {64 * c[o + 1] / c[o + 2]:.0f} (tiles {c[o + 2]})", file=sys.stderr)
        if c[2]:
            print(f"[xs trace] per tile, cycles: wait u,v,w {64 * c[0] / c[2]:.0f}, interface u,v,w {64 * c[1] / c[2]:.0f}, wait T {64 * c[3] / c[2]:.0f} (tiles {c[2]})", file=sys.stderr)
    # ---- per-field checksums of the final layer (all-reduced): the same numbers at every N ----------------------------
    sums = sol.field_sums(0)
    checksums = {n: {"sum": v[0], "l2": v[1] ** 0.5} for n, v in sums.items()}

    # ---- roofline of the dominant kernel: algorithmic bytes per launch / average launch duration ---------------------
    # a directional sweep reads cur x4 + temp x4, writes next x4 + temp' x4 and reads 1 descriptor byte per
    # cell.  N > 1: the slab-coupled x-sweep is TWO passes over the slab - the spike pass reads 8 values + 1 byte and
    # writes nothing, the coupled pass moves the 16 + 1 - plus the interface solve; it is counted ONCE, with those bytes,
    # over the sum of its three spans.
    fpb = args.fp
    local_cells = sol.nx * case.dimy * case.dimz
    sweep_bytes = local_cells * (16 * fpb + 1)
    spike_bytes = local_cells * (8 * fpb + 1)
    peak, peak_src = peaks()
    per_dir = {}
    for k in ("sweep_x", "sweep_y", "sweep_z"):
        tot, n = timings[k]
        if not n:
            continue
        nbytes = sweep_bytes
        if k == "sweep_x" and timings["x_spike"][1]:
            tot += timings["x_spike"][0] + timings["x_interface"][0]
            nbytes += spike_bytes
        per_dir[k] = {"ms_per_launch": tot / n, "launches": n, "bytes_per_launch": nbytes, "gbs": nbytes / (tot / n * 1e-3) / 1e9,
                      "frac": nbytes / (tot / n * 1e-3) / 1e9 / peak, "kernel": sol.sweep_kernel_name(k)}
    dom = max(per_dir, key=lambda k: per_dir[k]["ms_per_launch"] * per_dir[k]["launches"]) if per_dir else None
    sweep_ms = sum(v["ms_per_launch"] * v["launches"] for v in per_dir.values())
    sweep_n = sum(v["launches"] for v in per_dir.values())
    step_bytes = local_cells * (NUM_GLOBAL * 3 * NUM_LOCAL * (16 * fpb + 1) + NUM_GLOBAL * 12 * fpb + 8 * fpb)   # BASELINE.md B_step
    # measured DRAM traffic per launch of the dominant kernel: one ncu --set full capture, committed under profiles/
    traffic, traffic_src = None, None
    key = f"{DX}x{DY}x{DZ}_f{fpb * 8}"
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        tp = ROOT / "profiles" / name
        if tp.exists() and world == 1 and dom:
            traffic = json.loads(tp.read_text()).get(key, {}).get(dom, {}).get("dram_bytes")
            if traffic:
                traffic_src = f"profiles/{name} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch of the dominant kernel)"
                break
    kname = {"sweep_x": "x", "sweep_y": "y", "sweep_z": "z"}
    roofline = {
        "bound": "hbm", "kernel": f"directional sweep along {kname[dom]} ({sol.sweep_kernel_name(dom)})" if dom else None,
        "achieved": per_dir[dom]["gbs"] if dom else None, "peak": peak, "unit": "GB/s",
        "frac": per_dir[dom]["frac"] if dom else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": per_dir[dom]["bytes_per_launch"] if dom else None,
        "per_direction": per_dir, "dominant": dom,
        "all_sweeps_gbs": sum(v["bytes_per_launch"] * v["launches"] for v in per_dir.values()) / (sweep_ms * 1e-3) / 1e9 if sweep_ms else None,
        "sweep_share_of_step": sweep_ms / max(ms, 1e-9),
        "step_achieved_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9, "step_frac": step_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
        "kernel_ms": {k: v[0] / args.steps for k, v in timings.items() if v[1]},
    }

    # ---- e2e: the same step through the public API with HOST buffers (H2D state, D2H layer) inside the timed region ----
    # Every step copies its inputs (the four fields of the layer the caller owns) from pinned host memory and reads its
    # result (residual + the full-resolution layer of GetLayer) back.  The copies are overlapped with the compute the way a
    # production driver would: the inputs of step n+1 are uploaded on the copy stream while step n runs
    # (cmc_adi3d_write_layer_async / _commit), the layer read back by step n lands while step n+1 runs
    # (cmc_adi3d_get_layer_async / _wait).  `serial` is the same loop with blocking calls (round 1's e2e).
    e2e = None
    if not args.no_e2e:
        ft = torch.float32 if fpb == 4 else torch.float64
        host_in = [torch.empty(local_cells, dtype=ft).pin_memory() for _ in range(4)]
        for q in range(4):
            host_in[q].numpy()[:] = sol.read_field(0, q).ravel()
        # N > 1: every rank receives the output rows of its own slab (option "local_output": N host links carry the result);
        # `gathered` below is the same loop with everything collected on rank 0 (one host link)
        lo_, hi_ = sol.output_rows(0)
        own = (hi_ - lo_) * DY * DZ
        host_vel = [torch.empty(max((ncells if rank == 0 else own) * 3, 3), dtype=ft).pin_memory() for _ in range(2)]
        host_T = [torch.empty(max(ncells if rank == 0 else own, 1), dtype=torch.float64).pin_memory() for _ in range(2)]
        h2d = 4 * ncells * fpb                                   # all ranks together
        d2h = ncells * (3 * fpb + 8) + 16
        ins = [h.numpy() for h in host_in]

        def serial_step(i):
            for q in range(4):
                sol.write_field(0, q, ins[q])                        # host -> device: the layer the caller owns
            sol.UpdateBoundaries()
            sol.TimeStep(case.dt, NUM_GLOBAL, NUM_LOCAL, True)       # residual read back
            sol.GetLayer(0, 0, 0, vel=host_vel[0].numpy().reshape(-1, 3), T=host_T[0].numpy())   # device -> host: full layer

        def timed(fn, nsteps, before=None, after=None):
            if before:
                before()
            barrier()
            t0 = time.perf_counter()
            for i in range(nsteps):
                fn(i)
            if after:
                after()
            barrier()
            ms_ = (time.perf_counter() - t0) * 1e3
            if world > 1:
                import torch.distributed as dist
                t = torch.tensor([ms_], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t.item())
            return ms_

        serial_step(0)
        serial_ms = timed(serial_step, 2)

        def piped_step(i):
            sol.write_layer_commit(0)                                # this step's inputs (upload started during the previous step)
            sol.UpdateBoundaries()
            sol.TimeStepAsync(case.dt, NUM_GLOBAL, NUM_LOCAL, True)
            sol.write_layer_async(*ins)                              # the next step's inputs: host -> device behind the compute
            k = i & 1
            sol.GetLayerWait()                                       # the previous readback has landed: its host buffer is free
            sol.GetLayerAsync(host_vel[k].numpy().reshape(-1, 3), host_T[k].numpy())     # device -> host behind the next step
            sol.Sync()                                               # the residual of this step (host-visible every step)

        def piped_start():
            sol.write_layer_async(*ins)

        def piped_end():
            sol.GetLayerWait()
            sol.write_layer_commit(0)                                # (consume the last upload)
            sol.Sync()

        gathered = None
        if world > 1:
            timed(piped_step, 2, piped_start, piped_end)
            g_ms = timed(piped_step, 3, piped_start, piped_end)
            gathered = {"value": ncells * 3 / (g_ms * 1e-3) / 1e6, "ms_per_step": g_ms / 3, "steps": 3,
                        "what": "the same loop with the whole output gathered on rank 0 (one host link carries the 4.29 GB result)"}
            sol.set_option("local_output", 1)
        timed(piped_step, 2, piped_start, piped_end)                 # warm-up: staging buffers, copy streams
        e2e_ms = timed(piped_step, args.e2e_steps, piped_start, piped_end)
        if world > 1:
            sol.set_option("local_output", 0)
        e2e = {"value": ncells * args.e2e_steps / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms / args.e2e_steps, "steps": args.e2e_steps,
               "serial": {"value": ncells * 2 / (serial_ms * 1e-3) / 1e6, "ms_per_step": serial_ms / 2, "steps": 2,
                          "what": "the same step with blocking calls: write_field x4, UpdateBoundaries, TimeStep(computeError), GetLayer"},
               "what": "per step: inputs of the step from pinned host memory (write_layer_async during the previous step, write_layer_commit), "
                       "UpdateBoundaries, TimeStep(computeError) with the residual read back, GetLayer at full resolution into pinned host memory "
                       "(get_layer_async, landing during the next step; " + ("every rank receives the output rows of its own slab" if world > 1 else "one GPU, one host buffer") +
                       "); host wall clock between barriers, max over ranks"}
        if gathered:
            e2e["gathered"] = gathered

    # the reported CPU baseline: rank 0 runs the reference solver on the host cores while the other ranks wait at a barrier
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            r = run_reference_cpu(3, 1, size=args.cpu_size)
            if r:
                cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": cpu_sample_text(r)}
        except Exception as ex:      # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference", "sample": f"failed: {ex}"}
    barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if fpb == 8 else "f32", "data": "synthetic",
            "config": bench_config(args, world),
            "implementation": {"mode": args.mode, "exchange": sol.exchange_kind(), "fluid_fraction": round(fluid, 4),
                               "node_arrays": "slab-local (own planes + 2 halo planes per rank)" if world > 1 else "whole grid",
                               "storage": f"SoA, y-blocked ({sol.storage_block_rows()} rows per block)" if sol.storage_block_rows() else "SoA [i][j][k]"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk, "residual": err, "checksums": checksums, "device_bytes": sol.device_bytes(),
        }
        print(json.dumps(line))
    sol.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
