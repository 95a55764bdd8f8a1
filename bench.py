#!/usr/bin/env python
"""bench.py — air-sea flux points/s (F64) for the atmosphere–surface interface step.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (BASELINE.json configs[3]): the 1/12° (4320x1680) exchange grid with a JRA55-shaped
synthetic atmosphere (640x320, Float32, 3-hourly), sharded in latitude bands over N GPUs (strong
scaling: the global grid is fixed).  One *step* = one pass of the hot path: radiation + atmosphere
interpolation -> atmosphere–ocean similarity-theory solve -> net ocean flux assembly -> radiative
flux application -> global flux diagnostics (local sum + all-reduce when N > 1).

`value`   points/s with all inputs resident in HBM (CUDA events, max over ranks).
`e2e`     the same through the public API with HOST buffers: every step copies the ocean surface
          state (T, S, u, v) from pinned host memory to the device and reads the diagnostics back.
`roofline` the dominant kernel (the a–o solve) against the measured FP64 DFMA peak;
`roofline_hbm` the interpolation kernel against the measured HBM copy bandwidth.
`cpu_baseline` the CPU oracle (a port of the reference algorithm; Julia is not installed) on the
          host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints exactly ONE line on stdout.  NCCL_DEBUG is left as the caller set it (INFO shows the communicator's
# nranks / NVLS lines): whatever a library writes to file descriptor 1 (NCCL's log and version banner) is sent to
# stderr, and the JSON line is written to a private duplicate of the original stdout
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(line + "\n")
    _REAL_STDOUT.flush()


import numpy as np  # noqa: E402

import ne_b200  # noqa: E402,F401  (registers the package as numericalearth_jl_b200)

# algorithmic FP64 work of the default a–o variant (oracle op census x SURVEY §8(d) weights)
F_ITER = 4112.5
F_EPI = 150.0
BYTES_INTERP_ATM = 2 * 4 + 7 * 8      # 2 Float32 fractional indices in, 7 Float64 fields out
BYTES_INTERP_RAD = 2 * 4 + 2 * 8
DT_STEP = 1200.0


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of one committed `ncu --set full` capture of
    this workload, or None when no capture of this configuration is on file (profiles/ncu_traffic.json names its source)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="C4")
    p.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    p.add_argument("--atm-dtype", default="f32", choices=["f64", "f32"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--e2e-chunks", type=int, default=None,
                   help="latitude chunks of the host-buffer step (default 8; 3 for --dtype f32, whose persistent work-queue "
                        "solve pays a fixed cost per band launch: 4.09 / 3.92 / 3.68 ms per step at 8 / 5 / 3 chunks on C4)")
    p.add_argument("--equal-bands", action="store_true", help="equal row counts per band instead of cost-balanced bands")
    p.add_argument("--no-extras", action="store_true", help="skip the one-number measurements of the other BASELINE configs")
    p.add_argument("--sync-allreduce", action="store_true", help="diagnostics all-reduce on the compute stream (not overlapped)")
    p.add_argument("--no-rebalance", action="store_true", help="keep the mask-based bands (no re-balancing from measured trip counts)")
    p.add_argument("--no-time-rebalance", action="store_true", help="skip the second re-balancing pass (measured step time per band)")
    p.add_argument("--sustained-seconds", type=float, default=1.0, help="second timed region of at least this length (0: skip)")
    p.add_argument("--no-parity", action="store_true", help="skip the oracle parity check on the cpu_baseline sample")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (port of the reference algorithm) on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_step_factory(args, sample_ny):
    import ne_b200
    import oracle
    from numericalearth_jl_b200 import synthetic
    lib = oracle.load()
    host = ne_b200.NumpyHostBackend()
    cfg = dict(synthetic.CONFIGS[args.config])
    cfg["ny"] = sample_ny       # same longitudes and latitude range, rows subsampled
    ci = synthetic.build_case(cfg, host, FT=args.dtype, atm_FT=args.atm_dtype, lib=lib)
    ci.initialize()
    state = {"t": 0.0}

    def step():
        ci.update_state(state["t"])
        state["t"] += DT_STEP
    return ci, step, lib


def run_cpu(args, seconds_per_step=1.5, steps=None, warmup=1):
    """Time the oracle interface step on a bounded latitude-subsampled sample of the workload."""
    import oracle
    from numericalearth_jl_b200 import synthetic
    lib = oracle.load()
    # all the host cores this process may use — torchrun exports OMP_NUM_THREADS=1, which would leave the oracle on one
    try:
        lib.dll.neo_set_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        lib.dll.neo_set_threads(os.cpu_count() or 1)
    threads = lib.dll.neo_max_threads()
    full_ny = synthetic.CONFIGS[args.config]["ny"]
    ci, step, _ = cpu_step_factory(args, 8)
    t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
    rate = ci.grid.launch_points() / max(dt, 1e-9)
    nx = ci.grid.nx
    ny = int(min(full_ny, max(8, seconds_per_step * rate / (nx + 2))))
    ci, step, _ = cpu_step_factory(args, ny)
    for _ in range(warmup):
        step()
    times = []
    n = steps if steps is not None else 5
    for _ in range(n):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    pts = ci.grid.launch_points()
    total = float(np.sum(times))
    return {"value": pts * n / total, "unit": "points/s", "cores": int(threads), "kind": "port",
            "sample": f"{args.config} subsampled in latitude to {nx}x{ny} ({pts} launch points/step), {n} steps, "
                      f"OpenMP oracle, {threads} threads", "ms_per_step": 1e3 * total / n, "points_per_step": pts, "sample_ny": ny}


def parity_on_sample(args, backend, sample_ny):
    """The oracle as the checker: the cpu_baseline sample (the workload subsampled in latitude) through the CUDA path and
    through the oracle, compared point by point.  |a - b| <= tol * max(|b|, 1e-6 max|b|), tol = 1e-10 (Float64) / 1e-5
    (Float32); fields: the nine a-o outputs, the six net ocean fluxes, the three radiative fluxes."""
    import ne_b200
    import oracle
    from numericalearth_jl_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.config])
    cfg["ny"] = int(sample_ny)
    t = 0.37 * 10800.0
    ref = synthetic.build_case(cfg, ne_b200.NumpyHostBackend(), FT=args.dtype, atm_FT=args.atm_dtype, lib=oracle.load(), with_iterations=True)
    dev = synthetic.build_case(cfg, backend, FT=args.dtype, atm_FT=args.atm_dtype, with_iterations=True)
    ref.initialize(); dev.initialize()
    ref.update_state(t)
    dev.fused_interface_step(t)
    dev.fused_interface_step(t)          # the second step runs in trip-count order
    backend.synchronize()
    g = ref.grid
    tol = 1e-10 if args.dtype == "f64" else 1e-5
    it_r = g.interior(ref.ao_iterations).astype(np.int64)
    it_d = g.interior(backend.to_numpy(dev.ao_iterations)).astype(np.int64)
    maxiter = 100
    ok = (it_r < maxiter) & (it_d < maxiter)
    worst, exceed, fields = 0.0, 0, {}
    bags = [("ao_fluxes", True), ("net_ocean", False), ("rad_fluxes_ocean", False)]
    for bag, ring in bags:
        for n in getattr(ref, bag).names():
            a, b = backend.to_numpy(getattr(getattr(dev, bag), n)), getattr(getattr(ref, bag), n)
            if ring:
                a, b, m = g.interior(a), g.interior(b), ok
            else:
                sl = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
                a, b, m = a[sl], b[sl], ok[1:-1, 1:-1]
            a, b = a.astype(np.float64)[m], b.astype(np.float64)[m]
            scale = float(np.max(np.abs(b))) or 1.0
            r = np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * scale)
            fields[f"{bag}.{n}"] = float(r.max())
            worst = max(worst, float(r.max()))
            exceed += int((r > tol).sum())
    interp_exact = all(np.array_equal(g.interior(backend.to_numpy(getattr(dev.atmos_state, n))), g.interior(getattr(ref.atmos_state, n)))
                       for n in ref.atmos_state.names())
    solved = it_r > 0
    return {"checker": "oracle (C++ restatement of the reference; Julia absent)", "sample": f"{args.config} subsampled in latitude to {g.nx}x{g.ny}",
            "criterion": "|a-b| <= tol*max(|b|, 1e-6*max|b|)", "tol": tol, "max_pointwise_rel": worst, "points_above_tol": exceed,
            "per_field_max_pointwise_rel": fields, "interpolation_bit_exact": bool(interp_exact),
            "trip_count_mismatch_rate": float((it_r != it_d)[solved].mean()) if solved.any() else 0.0,
            "trip_count_max_abs_diff": int(np.abs(it_r - it_d).max()),
            "maxiter_points": {"oracle": int((it_r >= maxiter).sum()), "device": int((it_d >= maxiter).sum())},
            "solved_points": int(solved.sum())}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args, seconds_per_step=2.0, steps=args.steps, warmup=max(args.warmup, 1))
    line = {"impl": "reference", "metric": "air-sea flux points/s", "value": r["value"], "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "Julia is not installed: the reference arm is the C++/OpenMP oracle restating the reference algorithm"}
    emit(json.dumps(line))


def workload_config(args, n):
    from numericalearth_jl_b200 import synthetic
    c = synthetic.CONFIGS[args.config]
    return {"workload": f"{args.config}: {c['nx']}x{c['ny']} exchange grid (lat {c['latitude']}), JRA55-shaped 640x320 "
                        f"{args.atm_dtype} atmosphere + radiation, SimilarityTheoryFluxes defaults, OceanOnlyModel interface step",
            "exchange_dtype": args.dtype, "atmosphere_dtype": args.atm_dtype,
            "partition": f"{n} latitude band(s)" + ("" if n == 1 else ", " + getattr(args, "partition_note", "rows balanced by active-point count")) +
                         ", one-ring overcompute, no data-path collective",
            "l2": "inputs larger than L2 (≈35 fields x 58 MB at N=1); no explicit flush",
            "points_per_step": (c["nx"] + 2) * (c["ny"] + 2)}


# -------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import ne_b200
    from numericalearth_jl_b200 import sharding, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    backend = ne_b200.TorchCudaBackend(f"cuda:{local_rank}")
    lib = ne_b200.get_library()
    cfg = synthetic.CONFIGS[args.config]
    # latitude bands balanced by row cost.  First guess: active-point count from the land mask (static information);
    # then, as a coupled run would do every so often, re-balanced once from the trip counts the solve itself reports
    # (`iterations` of a set-up step; setup is not timed).
    weights = synthetic.row_cost_weights(args.config, FT=args.dtype) if (world > 1 and not args.equal_bands) else None
    grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], rank, world, FT=args.dtype, weights=weights)
    ci = synthetic.build_case(args.config, backend, FT=args.dtype, atm_FT=args.atm_dtype, grid=grid, with_iterations=True)
    partition_note = "equal row counts" if args.equal_bands else "rows balanced by active-point count"
    if world > 1 and not args.equal_bands and not args.no_rebalance:
        ci.initialize()
        ci.update_state(0.37 * 10800.0)
        torch.cuda.synchronize()
        active_rows, trip_rows, warp_trip_rows = sharding.gather_row_statistics(backend.to_numpy(grid.interior(ci.ao_iterations)), grid)
        if args.dtype == "f64":   # trip-ordered solve: a warp's lanes leave the loop together, the cost is the plain trip sum
            weights = sharding.ordered_row_weights(cfg["nx"], active_rows, trip_rows)
        else:
            weights = sharding.measured_row_weights(cfg["nx"], warp_trip_rows)
        new_grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], rank, world, FT=args.dtype, weights=weights)
        moved = torch.tensor([int((new_grid.j_offset, new_grid.ny) != (grid.j_offset, grid.ny))], device=backend.device)
        dist.all_reduce(moved)
        if int(moved.item()) > 0:
            del ci
            torch.cuda.empty_cache()
            grid = new_grid
            ci = synthetic.build_case(args.config, backend, FT=args.dtype, atm_FT=args.atm_dtype, grid=grid, with_iterations=True)
        partition_note = "rows balanced by measured trip counts of a set-up step"
        # second pass, on the clock: the cost model above is a fit; what is left after it (±5 % between ranks at N = 8) is
        # measured — a few untimed steps per rank — and each rank's rows are re-weighted by (its time / the mean time)
        if not args.no_time_rebalance:
            ci.initialize()
            fd = ci.fused_step_desc(0.37 * 10800.0)
            for _ in range(3):
                lib.call("fused_interface_step", args.dtype, fd, backend.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(10):
                lib.call("fused_interface_step", args.dtype, fd, backend.stream())
            e1.record()
            torch.cuda.synchronize()
            mine = torch.tensor([e0.elapsed_time(e1) / 10, float(grid.j_offset), float(grid.ny)], device=backend.device, dtype=torch.float64)
            allt = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allt, mine)
            times = np.array([float(x[0]) for x in allt])
            if times.max() > 1.02 * times.mean():
                w2 = np.array(weights, dtype=np.float64, copy=True)
                for x in allt:
                    j0, nrows = int(x[1]), int(x[2])
                    w2[j0:j0 + nrows] *= float(x[0]) / times.mean()
                new_grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], rank, world, FT=args.dtype, weights=w2)
                moved = torch.tensor([int((new_grid.j_offset, new_grid.ny) != (grid.j_offset, grid.ny))], device=backend.device)
                dist.all_reduce(moved)
                if int(moved.item()) > 0:
                    del ci, fd
                    torch.cuda.empty_cache()
                    grid = new_grid
                    ci = synthetic.build_case(args.config, backend, FT=args.dtype, atm_FT=args.atm_dtype, grid=grid, with_iterations=True)
                partition_note += f", then re-weighted once by the measured step time of each band (spread before: {times.min():.4f}-{times.max():.4f} ms)"
    args.partition_note = partition_note
    ci.initialize()
    f = ci.ao_fluxes
    diag = sharding.FluxDiagnostics(ci, [f.latent_heat, f.sensible_heat, f.water_vapor, f.x_momentum, f.y_momentum,
                                         ci.net_ocean.T, ci.net_ocean.eta])
    stream = backend.stream()
    global_points = (cfg["nx"] + 2) * (cfg["ny"] + 2)
    local_points = grid.launch_points()

    # cached descriptor: only the time interpolator changes from step to step
    fused = ci.fused_step_desc(0.0, diagnostics=diag)   # the diagnostics' local sums ride on the assembly kernel
    atm_times = ci.atmosphere.times

    def set_time(t):
        nt, n1, n2 = ne_b200.interpolating_time_indices(atm_times, t, "cyclical")
        for d in (fused.atmosphere, fused.radiation):
            d.time.frac, d.time.m1, d.time.m2, d.time.same = nt, n1, n2, int(n1 == n2)

    state = {"t": 0.37 * 10800.0}

    def step():
        set_time(state["t"])
        diag.flip(fused)                 # the previous step's all-reduce may still be in flight on the other result vector
        lib.call("fused_interface_step", args.dtype, fused, stream)
        diag.all_reduce(async_op=not args.sync_allreduce)   # on the communicator's stream, overlapped by the next step's kernels
        state["t"] += DT_STEP

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        diag.wait()                      # the timed region ends after the last step's all-reduce
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=backend.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput ---------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = global_points / (ms_per_step * 1e-3)
    # a second, long timed region: the K-step figure above is a ~40 ms burst at boost clock, this one is >= 1 s
    sustained = None
    if args.sustained_seconds > 0:
        n_long = int(max(args.steps, min(5000, np.ceil(args.sustained_seconds * 1e3 / max(ms_per_step, 1e-3)))))
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        ms_long = timed(step, n_long)
        c2 = sampler2.stop() if rank == 0 else None
        sustained = {"steps": n_long, "ms_per_step": ms_long / n_long, "value": global_points / (ms_long / n_long * 1e-3),
                     "seconds": ms_long * 1e-3, "clocks": c2}

    # ---- per-kernel timings for the rooflines (same inputs, kernel timed alone on the launch stream) ---
    ao_desc = ci.atmosphere_ocean_desc()     # as the step calls it (iterations array included: trip-count ordered launch)
    atm_desc, rad_desc = fused.atmosphere, fused.radiation
    reps = max(5, min(args.steps, 20))
    ms_ao = timed(lambda: lib.call("atmosphere_ocean_fluxes", args.dtype, ao_desc, stream), reps) / reps
    # this rank's own solve time (`timed` returns the max over ranks): how well the bands are balanced
    per_rank_ao = None
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(reps):
            lib.call("atmosphere_ocean_fluxes", args.dtype, ao_desc, stream)
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / reps], device=backend.device, dtype=torch.float64)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        per_rank_ao = [round(float(x.item()), 4) for x in allt]
    ms_ia = timed(lambda: lib.call("interp_state", args.dtype, atm_desc, stream), reps) / reps
    ms_ir = timed(lambda: lib.call("interp_state", args.dtype, rad_desc, stream), reps) / reps
    ms_as = timed(lambda: lib.call("assemble_net_ocean_fluxes", args.dtype, fused.assemble, stream), reps) / reps
    ms_ap = timed(lambda: lib.call("apply_radiative_fluxes", args.dtype, fused.apply_radiation, stream), reps) / reps
    ms_dg = timed(lambda: lib.call("diag_reduce", args.dtype, diag.desc, stream), reps) / reps
    # what the step actually runs after the solve: ONE kernel (assembly + radiation + diagnostics partial sums) + the
    # final diagnostics stage; the three component kernels above are timed for reference only
    os.environ["NE_B200_NO_POST_SOLVE_FUSION"] = "1"
    ms_unfused_step = timed(step, reps) / reps
    os.environ.pop("NE_B200_NO_POST_SOLVE_FUSION")
    it = backend.to_numpy(grid.interior(ci.ao_iterations))
    iters_sum = float(it.sum())
    if world > 1:
        t = torch.tensor([iters_sum, float(local_points)], device=backend.device, dtype=torch.float64)
        dist.all_reduce(t)
        iters_sum_g, pts_g = float(t[0]), float(t[1])
    else:
        iters_sum_g, pts_g = iters_sum, float(local_points)
    active = it > 0
    flops_local = iters_sum * F_ITER + float(active.sum()) * F_EPI
    fp64_peak, _ = lib.measure_fp64_peak()
    peaks, peak_src = load_peaks()
    algorithmic_tf = flops_local / (ms_ao * 1e-3) / 1e12
    # executed FP64 work of ONE launch of the solve on this rank's band, counted by the counting instantiation of the
    # same kernel source (ne_count_solve_ops_f64; no profiler): thread-level DFMA (2 flop), DMUL, DADD
    counts = None
    if args.dtype == "f64":
        try:
            cd = ci.atmosphere_ocean_desc()     # with the iterations array: the launch runs in trip-count order, as in the step
            counts = lib.count_solve_ops(cd, stream)
        except Exception as e:   # noqa: BLE001
            counts = {"error": repr(e)[:200]}
    executed_tf = counts["flop"] / (ms_ao * 1e-3) / 1e12 if counts and "flop" in counts else None
    wbytes = 8 if args.dtype == "f64" else 4
    ibytes = 4 if args.atm_dtype == "f32" else 8
    interp_bytes = local_points * (2 * ibytes + 7 * wbytes)
    hbm_achieved = interp_bytes / (ms_ia * 1e-3) / 1e9

    # ---- end to end with host buffers -----------------------------------------------------------------
    # public API: ne_b200.HostPipelinedStep — chunked H2D of the ocean surface state (pinned host memory) on a
    # copy stream overlapped with the band-restricted kernels, diagnostics read back to the host every step
    o = ci._host_inputs["ocean"]
    pinned = {k: torch.from_numpy(np.ascontiguousarray(o[k])).pin_memory() for k in ("T", "S", "u", "v")}
    if args.e2e_chunks is None:
        args.e2e_chunks = 8 if args.dtype == "f64" else 3
    pipe = ne_b200.HostPipelinedStep(ci, n_chunks=args.e2e_chunks, diagnostics=diag)
    h2d = pipe.h2d_bytes_per_step()
    result_host = torch.empty(diag.result.shape, dtype=torch.float64).pin_memory()
    d2h = result_host.numel() * 8

    def e2e_step():
        pipe.step(state["t"], pinned)
        result_host.copy_(diag.result, non_blocking=True)
        state["t"] += DT_STEP

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = global_points / (ms_e2e * 1e-3)
    torch.cuda.synchronize()
    diag_values = [float(x) for x in result_host.tolist()]
    # round trip: the same step, and the six net ocean fluxes come back to pinned host memory band by band behind the
    # kernels (the result an ocean model on the host needs), next to the 7 diagnostics sums
    net = ci.net_ocean
    back = {n: (getattr(net, n), torch.empty(getattr(net, n).shape, dtype=getattr(net, n).dtype).pin_memory()) for n in net.names()}
    pipe_rt = ne_b200.HostPipelinedStep(ci, n_chunks=args.e2e_chunks, diagnostics=diag, return_fields=back)
    d2h_rt = pipe_rt.d2h_bytes_per_step() + d2h

    def e2e_rt_step():
        pipe_rt.step(state["t"], pinned)
        result_host.copy_(diag.result, non_blocking=True)
        state["t"] += DT_STEP

    for _ in range(2):
        e2e_rt_step()
    ms_e2e_rt = timed(e2e_rt_step, args.steps) / args.steps
    torch.cuda.synchronize()
    rt_ok = bool(torch.equal(back["T"][1], net.T.cpu()))

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu(args, seconds_per_step=args.cpu_seconds / 6.0, steps=5, warmup=1)
        cpu = {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if not args.no_parity:
            try:
                parity = parity_on_sample(args, backend, r["sample_ny"])
            except Exception as e:   # noqa: BLE001
                parity = {"error": repr(e)[:300]}

    # ---- the other BASELINE configurations, one number each (N = 1 only; never allowed to break the headline line) ----
    extras = None
    if not args.no_extras and args.config != "C5":
        # BASELINE config 5 (1/48 degree) in both precisions at THIS GPU count: each rank builds its own latitude band
        # (band-local random surface state: only the throughput matters here), fused step incl. diagnostics + all-reduce
        c5 = {}
        for ft in ("f64", "f32"):
            try:
                c5cfg = synthetic.CONFIGS["C5"]
                # equal rows: the band-local mask below is ~30 % inactive in every band, so the bands cost the same
                g5 = sharding.band_grid(c5cfg["nx"], c5cfg["ny"], c5cfg["latitude"], rank, world, FT=ft)
                ci5 = synthetic.build_case("C5", backend, FT=ft, atm_FT=args.atm_dtype, grid=g5, with_iterations=True, local_surface=True)
                ci5.initialize()
                f5 = ci5.ao_fluxes
                dg5 = sharding.FluxDiagnostics(ci5, [f5.latent_heat, f5.sensible_heat, f5.water_vapor, f5.x_momentum, f5.y_momentum,
                                                     ci5.net_ocean.T, ci5.net_ocean.eta])
                fd5 = ci5.fused_step_desc(0.37 * 10800.0, diagnostics=dg5)

                def step5():
                    dg5.flip(fd5)
                    lib.call("fused_interface_step", ft, fd5, stream)
                    dg5.all_reduce(async_op=not args.sync_allreduce)

                for _ in range(3):
                    step5()
                dg5.wait()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record()
                for _ in range(8):
                    step5()
                dg5.wait()
                e1.record()
                barrier()
                ms5 = e0.elapsed_time(e1) / 8
                if world > 1:
                    t5 = torch.tensor([ms5], device=backend.device, dtype=torch.float64)
                    dist.all_reduce(t5, op=dist.ReduceOp.MAX)
                    ms5 = float(t5.item())
                pts5 = (c5cfg["nx"] + 2) * (c5cfg["ny"] + 2)
                c5[f"C5_{ft}"] = {"ms_per_step": ms5, "points_per_s": pts5 / (ms5 * 1e-3), "n_gpus": world, "steps": 8,
                                  "points_per_step": pts5}
                del ci5, dg5, fd5, f5
                torch.cuda.empty_cache()
            except Exception as e:   # noqa: BLE001
                c5[f"C5_{ft}"] = {"error": repr(e)[:200]}
        extras = dict(c5)
        extras["C5_note"] = ("BASELINE config 5, 17280x6720, device-resident fused step incl. diagnostics, latitude bands balanced by "
                             "active-point count, max over ranks; band-local random surface state")
    if rank == 0 and world == 1 and not args.no_extras:
        extras = extras or {}
        try:
            other = "f32" if args.dtype == "f64" else "f64"
            c2 = synthetic.build_case(args.config, backend, FT=other, atm_FT=args.atm_dtype)
            c2.initialize()
            f2 = c2.ao_fluxes
            dg2 = sharding.FluxDiagnostics(c2, [f2.latent_heat, f2.sensible_heat, f2.water_vapor, f2.x_momentum, f2.y_momentum,
                                                c2.net_ocean.T, c2.net_ocean.eta])
            fd2 = c2.fused_step_desc(0.37 * 10800.0, diagnostics=dg2)
            for _ in range(3):   # the first call builds and uploads the solver table of this parameter set
                lib.call("fused_interface_step", other, fd2, stream)
            extras[f"{args.config}_{other}_model_ms_per_step"] = timed(lambda: lib.call("fused_interface_step", other, fd2, stream), reps) / reps
            del c2, dg2, fd2
            torch.cuda.empty_cache()
            c3 = synthetic.build_case("C3", backend, FT=args.dtype, atm_FT=args.atm_dtype, sea_ice=True)
            c3.initialize()
            col = synthetic.ocean_column(c3.grid, backend, nz=10) + (600.0, 10)
            c3.update_state(0.37 * 10800.0, ocean_column=col)
            extras[f"C3_ocean_sea_ice_{args.dtype}_ms_per_step"] = timed(lambda: c3.update_state(0.37 * 10800.0, ocean_column=col), reps) / reps
            extras["note"] = ("other BASELINE configs, device-resident: the same grid in the other exchange precision (fused step incl. "
                              "diagnostics); config 3 = 1/4 degree OceanSeaIceModel update_state (a-o + a-si + si-o kernels, 10-level frazil "
                              "column, net fluxes, both radiation kernels; Python call sequence, 9 launches)")
        except Exception as e:   # noqa: BLE001
            extras["error"] = repr(e)[:200]

    cfg_key = f"{args.config}_{args.dtype}_atm{args.atm_dtype}" if world == 1 else None
    traffic = ncu_traffic(cfg_key) if cfg_key else None
    if rank == 0:
        n_active = int(active.sum())
        have_counts = bool(counts) and "flop" in counts
        line = {
            "metric": "air-sea flux points/s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e, "api": "ne_b200.HostPipelinedStep", "chunks": int(args.e2e_chunks),
                    "gpu_launches_per_step": int(pipe.launches_per_step()),
                    "round_trip": {"value": global_points / (ms_e2e_rt * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e_rt,
                                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_rt),
                                   "returns": "the six net ocean flux fields (band by band, behind the kernels) + 7 diagnostics sums",
                                   "host_copy_equals_device": rt_ok}},
            "gpu_launches": int((5 if ci.shared_frac else 6) * args.steps),
            "clocks": clocks,
            "sustained": sustained,
            "roofline": {"bound": "fp64", "kernel": "ao_flux_tab2_kernel (+ trip_order_kernel)",
                         "achieved": executed_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (executed_tf / fp64_peak) if (executed_tf and fp64_peak) else None,
                         "what": "EXECUTED thread-level FP64 flop of one launch (DFMA = 2, DMUL = DADD = 1; counted in this run by "
                                 "the counting instantiation of the same kernel source, ne_count_solve_ops_f64) / CUDA-event time of the launch",
                         "executed_counts": counts,
                         # the attainable flop rate of THIS instruction mix: every FP64 instruction occupies the pipe like a DFMA
                         # but DMUL / DADD carry one flop
                         "frac_of_mix_peak": ((counts["dfma"] + counts["dmul"] + counts["dadd"]) * 2.0 / (ms_ao * 1e-3) / 1e12 / fp64_peak)
                         if have_counts and fp64_peak else None,
                         "frac_of_mix_peak_note": "FP64 instructions x 2 / time / DFMA peak = share of the FP64 pipe's issue rate the solve uses "
                                                  "(ncu sm__pipe_fp64_cycles_active of the same kernel: profiles/r02_notes.md)",
                         "lane_efficiency": (counts["thread_trips"] / (32.0 * counts["warp_trips"])) if have_counts and counts["warp_trips"] else None,
                         "traffic": (traffic or {}).get("ao_flux_bytes"), "traffic_source": (traffic or {}).get("source"),
                         "peak_source": "measured in this run by ne_measure_fp64_peak (DFMA chains); MEASURED_PEAKS.json has no FP64 figure",
                         "ms_per_launch": ms_ao,
                         "algorithmic": {"achieved": algorithmic_tf, "frac_of_peak": algorithmic_tf / fp64_peak if fp64_peak else None,
                                         "flop_per_iteration": F_ITER, "flop_epilogue": F_EPI,
                                         "note": "as-written census of the reference's iteration (SURVEY 8(d) weights): the table-driven "
                                                 "kernel executes ~20x fewer flops, so this is NOT a fraction of peak; side note only"},
                         "mean_iterations_active": iters_sum / max(n_active, 1), "max_iterations": int(it.max()),
                         "active_points": n_active, "points_per_launch": int(local_points)},
            "roofline_hbm": {"bound": "hbm", "kernel": "interp_staged_kernel(atmosphere)", "achieved": hbm_achieved,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_achieved / peaks["hbm_gbs"],
                             "peak_source": peak_src, "ms_per_launch": ms_ia,
                             "traffic": (traffic or {}).get("interp_bytes")},
            "per_rank_solve_ms": per_rank_ao,
            "kernel_ms": {"interp_radiation": ms_ir, "interp_atmosphere": ms_ia, "atmosphere_ocean_fluxes": ms_ao,
                          "assemble_net_ocean_fluxes": ms_as, "apply_radiative_fluxes": ms_ap, "diag_reduce": ms_dg,
                          "step_with_unfused_post_solve_kernels": ms_unfused_step,
                          "note": "the step runs ONE interpolation launch (atmosphere + radiation: 9 series, shared fractional "
                                  "indices; the two launches above are timed alone for reference), the trip-order pass + the solve, ONE post-solve kernel (assembly + radiation + "
                                  "diagnostics partial sums) and the diagnostics final stage; the three post-solve "
                                  "component kernels are timed alone for reference"},
            "diagnostics": diag_values,
        }
        if parity is not None:
            line["parity"] = parity
        if extras:
            line["other_configs"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
