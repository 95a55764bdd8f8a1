"""Generate tests/golden/*.npz: oracle outputs for the seeded `tiny` case (48x20 exchange grid).

Julia is not installed, so these are NOT outputs of the reference itself; they freeze the oracle
(which is pinned against the reference's own known-answer tests in tests/test_oracle_reference_kats.py)
so that any later change to the oracle or to the synthetic generator is caught, and they give the GPU
tests a fixture that does not depend on the oracle having been rebuilt.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ne_b200  # noqa: E402
import oracle  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

T_STEP = 0.37 * 10800.0
CASES = {
    "tiny_f64_atm64": dict(FT="f64", atm_FT="f64"),
    "tiny_f64_atm32": dict(FT="f64", atm_FT="f32"),
    "tiny_f32_atm32": dict(FT="f32", atm_FT="f32"),
    "tiny_f64_seaice": dict(FT="f64", atm_FT="f64", sea_ice=True),
}


def run_case(lib, backend, kw):
    ci = synthetic.build_case("tiny", backend, lib=lib, with_iterations=True, **kw)
    ci.initialize()
    col = None
    if kw.get("sea_ice"):
        col = synthetic.ocean_column(ci.grid, backend, nz=4) + (1200.0, 4, 0)
    ci.update_state(T_STEP, ocean_column=col)
    return ci, col


def collect(ci, col, to_numpy):
    out = {}
    bags = {"frac": ci.frac, "atmos": ci.atmos_state, "rad": ci.rad_state, "ao": ci.ao_fluxes, "net_ocean": ci.net_ocean,
            "rad_ocean": ci.rad_fluxes_ocean}
    if ci.has_sea_ice:
        bags.update({"asi": ci.asi_fluxes, "sio": ci.sio_fluxes, "net_sea_ice": ci.net_sea_ice})
    for b, bag in bags.items():
        for n in bag.names():
            out[f"{b}.{n}"] = to_numpy(getattr(bag, n))
    out["ao.interface_temperature"] = to_numpy(ci.ao_temperature)
    out["ao.iterations"] = to_numpy(ci.ao_iterations)
    if ci.has_sea_ice:
        out["asi.top_temperature"] = to_numpy(ci.sea_ice_state.top_temperature)
        out["asi.iterations"] = to_numpy(ci.asi_iterations)
        out["column.T"] = to_numpy(col[0])
    return out


def main():
    lib = oracle.load()
    backend = ne_b200.NumpyHostBackend()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in CASES.items():
        ci, col = run_case(lib, backend, kw)
        np.savez_compressed(os.path.join(here, name + ".npz"), **collect(ci, col, lambda a: np.asarray(a)))
        print("wrote", name)


if __name__ == "__main__":
    main()
