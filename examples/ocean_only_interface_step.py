"""The interface step of an OceanOnlyModel on one B200, written against the host-side mirror of the reference's
operator interface (what `update_state!(model)` does in src/EarthSystemModels/time_step_earth_system_model.jl:38-83).

    python examples/ocean_only_interface_step.py            # 1/4 degree grid, 6 coupled steps

The prescribed series are synthetic and JRA55-shaped (640x320 Float32, 3-hourly); they live on the device as rings of 4
time levels fed from pinned host memory (series_window.SeriesWindow), the way a `JRA55NetCDFBackend(start, 4)` series is
partly in memory in the reference — except that a window slide is one asynchronous slice load behind the step's kernels
instead of a synchronous reload of the whole window.  Needs the CUDA extension and a GPU: there is no CPU fallback.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import sharding, synthetic  # noqa: E402


def main():
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    nt = 16                                                      # 2 days of 3-hourly forcing
    ci = synthetic.build_case("C3", backend, FT="f64", atm_FT="f32", nt=nt)   # exchange grid, ocean surface state, land mask

    # hand the prescribed atmosphere + radiation over to one ring of 4 time levels (9 series, one time axis)
    src = ci.atmosphere.grid
    a = ci._host_inputs["atmosphere"]
    raw = {k: np.ascontiguousarray(v[:, src.hy:src.hy + src.ny, src.hx:src.hx + src.nx]) for k, v in a.items()}
    ring = ne_b200.SeriesWindow(backend, lib, src, ci.atmosphere.times, raw, n_slots=4)
    atm, rad = ci.atmosphere, ci.radiation
    atm.u, atm.v, atm.T, atm.q, atm.p = ring["u"], ring["v"], ring["T"], ring["q"], ring["p"]
    atm.rain, atm.snow = (ring["rain"],), (ring["snow"],)
    rad.downwelling_shortwave, rad.downwelling_longwave = ring["sw"], ring["lw"]
    atm.window = rad.window = ring

    ci.initialize()                                              # fractional indices of the exchange nodes (initialize!)
    f = ci.ao_fluxes
    diag = sharding.FluxDiagnostics(ci, [f.latent_heat, f.sensible_heat, f.water_vapor, ci.net_ocean.T])
    dt = 3600.0
    for k in range(6):
        t = k * dt
        ci.fused_interface_step(t, diagnostics=diag)             # interpolation -> solve -> net fluxes + radiation + sums
        backend.synchronize()
        sums = backend.to_numpy(diag.result)
        area = float(backend.to_numpy(diag.area)[ci.grid.hy:-ci.grid.hy, ci.grid.hx:-ci.grid.hx].sum())
        print(f"t = {t / 3600:4.1f} h   <latent> = {sums[0] / area:8.2f} W/m2   <sensible> = {sums[1] / area:7.2f} W/m2   "
              f"ring: {ring.demand_loads} demand loads, {ring.prefetched} prefetched")
    ring.close()


if __name__ == "__main__":
    main()
