"""Solve-kernel time on one latitude band of C4 (rank r of N) for a list of env settings (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import sharding, synthetic  # noqa: E402

N, r = int(sys.argv[1]), int(sys.argv[2])
settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[3:]] or [{}]
backend = ne_b200.TorchCudaBackend("cuda:0")
lib = ne_b200.get_library()
cfg = synthetic.CONFIGS["C4"]
w = synthetic.row_cost_weights("C4")
grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], r, N, weights=w)
ci = synthetic.build_case("C4", backend, FT="f64", atm_FT="f32", grid=grid, with_iterations=True)
ci.initialize()
ci.interpolate_state(0.37 * 10800.0)
d = ci.atmosphere_ocean_desc()
stream = backend.stream()
lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
torch.cuda.synchronize()
it = backend.to_numpy(grid.interior(ci.ao_iterations))
lane_trips = float(it.sum())
for env in settings:
    os.environ.update(env)
    for _ in range(3):
        lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"band {r}/{N} rows {grid.ny} active {(it > 0).sum()} {env} {ms:.4f} ms  {lane_trips / ms / 1e6:.1f} G lane-trips/s", flush=True)
    for k in env:
        os.environ.pop(k)
