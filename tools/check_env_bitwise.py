"""Run the a-o solve on a config under two environment settings and compare every output bit for bit
(development tool).  usage: check_env_bitwise.py C4 "ENV=a" "ENV=b" """
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def run(ci, lib, d, stream, env, backend):
    for kv in env.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v
    for _ in range(2):   # the second launch is trip-ordered
        lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
    torch.cuda.synchronize()
    out = {n: backend.to_numpy(getattr(ci.ao_fluxes, n)).copy() for n in ci.ao_fluxes.names()}
    out["iterations"] = backend.to_numpy(ci.ao_iterations).copy()
    out["T"] = backend.to_numpy(ci.ao_temperature).copy()
    for kv in env.split(","):
        if kv:
            os.environ.pop(kv.split("=")[0], None)
    return out


def main():
    cfg, a, b = sys.argv[1:4]
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", with_iterations=True)
    ci.initialize()
    ci.interpolate_state(0.37 * 10800.0)
    d = ci.atmosphere_ocean_desc()
    ra = run(ci, lib, d, backend.stream(), a, backend)
    for n in ci.ao_fluxes.names():
        getattr(ci.ao_fluxes, n).zero_()
    rb = run(ci, lib, d, backend.stream(), b, backend)
    bad = [n for n in ra if not np.array_equal(ra[n], rb[n], equal_nan=True)]
    print(f"bitwise {cfg} [{a}] vs [{b}]:", "IDENTICAL" if not bad else f"DIFFERENT in {bad}", flush=True)
    for n in bad:
        x, y = ra[n].astype(np.float64), rb[n].astype(np.float64)
        m = ~(np.isnan(x) & np.isnan(y))
        print("  ", n, "points differing", int((x != y)[m].sum()), "max abs", float(np.nanmax(np.abs(x - y))))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
