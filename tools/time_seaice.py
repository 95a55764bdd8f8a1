"""Per-kernel device times of the OceanSeaIceModel interface step (BASELINE config 3: a-o, a-si, si-o kernels
together) on a chosen grid (development tool; numbers go to gpurun_out/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
    FT = sys.argv[2] if len(sys.argv) > 2 else "f64"
    backend = ne_b200.TorchCudaBackend("cuda:0")
    ci = synthetic.build_case(cfg, backend, FT=FT, atm_FT="f32", sea_ice=True, with_iterations=True)
    ci.initialize()
    col = synthetic.ocean_column(ci.grid, backend, nz=10)
    T3, S3, dz = col
    t = 0.37 * 10800.0
    ci.update_state(t, ocean_column=(T3, S3, dz, 600.0, 10))
    torch.cuda.synchronize()
    phases = {
        "interpolate_state": lambda: ci.interpolate_state(t),
        "atmosphere_ocean_fluxes": ci.compute_atmosphere_ocean_fluxes,
        "atmosphere_sea_ice_fluxes": ci.compute_atmosphere_sea_ice_fluxes,
        "sea_ice_ocean_fluxes": lambda: ci.lib.call("sea_ice_ocean_fluxes", FT, ci.sea_ice_ocean_desc(T3, S3, dz, 600.0, 10), backend.stream()),
        "update_net_fluxes": ci.update_net_fluxes,
        "apply_air_sea_radiative_fluxes": ci.apply_air_sea_radiative_fluxes,
        "apply_air_sea_ice_radiative_fluxes": ci.apply_air_sea_ice_radiative_fluxes,
        "update_state(all)": lambda: ci.update_state(t, ocean_column=(T3, S3, dz, 600.0, 10)),
    }
    out = {"config": cfg, "dtype": FT, "points": ci.grid.launch_points()}
    for name, fn in phases.items():
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 10
        print(cfg, FT, name, f"{out[name]:.3f} ms", flush=True)
    # the same phases in update_state order, one event after each call (device time) + host wall time of the sequence
    import time as _t
    seq = ["interpolate_state", "atmosphere_ocean_fluxes", "atmosphere_sea_ice_fluxes", "sea_ice_ocean_fluxes",
           "update_net_fluxes", "apply_air_sea_radiative_fluxes", "apply_air_sea_ice_radiative_fluxes"]
    for rep in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)]
        torch.cuda.synchronize()
        w0 = _t.perf_counter()
        ev[0].record()
        for k, name in enumerate(seq):
            phases[name]()
            ev[k + 1].record()
        w1 = _t.perf_counter()
        torch.cuda.synchronize()
        w2 = _t.perf_counter()
        print("sequence", rep, "host enqueue %.3f ms, until done %.3f ms;" % (1e3 * (w1 - w0), 1e3 * (w2 - w0)),
              " ".join(f"{n.split('_')[0]}..={ev[k].elapsed_time(ev[k + 1]):.3f}" for k, n in enumerate(seq)), flush=True)
    it = backend.to_numpy(ci.asi_iterations) if getattr(ci, "asi_iterations", None) is not None else None
    if it is not None:
        act = it[it > 0]
        out["asi_active"] = int(act.size)
        out["asi_mean_iterations"] = float(act.mean()) if act.size else 0.0
        out["asi_max_iterations"] = int(act.max()) if act.size else 0
        print("a-si active", act.size, "mean it", out["asi_mean_iterations"], "max", out["asi_max_iterations"])
    json.dump(out, open(f"gpurun_out/time_seaice_{cfg}_{FT}.json", "w"))


if __name__ == "__main__":
    main()
