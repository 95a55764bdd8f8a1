"""Per-step device time of the C4 interface step when the atmosphere + radiation series are device rings
(series_window.SeriesWindow) and the clock crosses 3-hourly intervals: prefetching window (n_slots 4), a window with
nowhere to prefetch into (n_slots 2: one demand load per interval) and, for the reference's behaviour, a whole-window
synchronous reload (every slot re-loaded, host waits) when the interpolating indices leave the window.
Development tool; prints one JSON line and writes gpurun_out/time_window.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def windowed_case(cfg, backend, lib, nt, n_slots):
    ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", nt=nt)
    src = ci.atmosphere.grid
    a = ci._host_inputs["atmosphere"]
    raw = {k: np.ascontiguousarray(v[:, src.hy:src.hy + src.ny, src.hx:src.hx + src.nx]) for k, v in a.items()}
    w = None
    if n_slots:
        w = ne_b200.SeriesWindow(backend, lib, src, ci.atmosphere.times, raw, n_slots=n_slots)
        atm, rad, s = ci.atmosphere, ci.radiation, w.series
        atm.u, atm.v, atm.T, atm.q, atm.p = s["u"], s["v"], s["T"], s["q"], s["p"]
        atm.rain, atm.snow = (s["rain"],), (s["snow"],)
        rad.downwelling_shortwave, rad.downwelling_longwave = s["sw"], s["lw"]
        atm.window = rad.window = w
    ci.initialize()
    return ci, w


def run(ci, w, steps, dt, reload_whole_window=False):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    for k in range(3):
        ci.fused_interface_step(k * dt)
    torch.cuda.synchronize()
    host_ms = []
    import time
    ev[0].record()
    for k in range(steps):
        t = (3 + k) * dt
        t0 = time.perf_counter()
        if reload_whole_window and w is not None:
            _, n1, n2 = ne_b200.interpolating_time_indices(w.times, t, w.time_indexing)
            if n1 not in w.policy.where or n2 not in w.policy.where:
                # Oceananigans' update_field_time_series!: new window starting at n1, set!(fts) of every slot, then go on
                torch.cuda.synchronize()
                w.policy.resident = [None] * w.n_slots
                w.policy.where = {}
                for s in range(w.n_slots):
                    n = (n1 - 1 + s) % len(w.times) + 1
                    w.policy._assign(n, s)
                    w._load(n, s)
                torch.cuda.synchronize()
        ci.fused_interface_step(t)
        host_ms.append((time.perf_counter() - t0) * 1e3)
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)])
    return ms, np.array(host_ms)


def main():
    cfg = os.environ.get("NE_CFG", "C4")
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    nt, steps, dt = 16, 96, 10800.0 / 4
    out = {}
    for name, n_slots, reload_ in (("in_memory", 0, False), ("ring_4_slots_prefetch", 4, False), ("ring_2_slots_demand", 2, False),
                                   ("whole_window_reload_4_slots", 4, True)):
        ci, w = windowed_case(cfg, backend, lib, nt, n_slots)
        if reload_:
            w.policy.lookahead = 0
        ms, host = run(ci, w, steps, dt, reload_)
        out[name] = {"mean_ms": float(ms.mean()), "median_ms": float(np.median(ms)), "max_ms": float(ms.max()),
                     "p95_ms": float(np.percentile(ms, 95)), "host_mean_ms": float(host.mean()), "host_max_ms": float(host.max()),
                     "demand_loads": getattr(w, "demand_loads", 0), "prefetched": getattr(w, "prefetched", 0)}
        if w is not None:
            w.close()
        del ci, w
        torch.cuda.empty_cache()
    out["config"] = {"workload": cfg, "steps": steps, "steps_per_interval": 4, "series": 9, "slice": "640x320 f32"}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/time_window.json", "w"), indent=1)


if __name__ == "__main__":
    main()
