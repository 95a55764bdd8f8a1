"""Write the RAW inputs of the golden cases (tests/golden/make_golden.py) as .npy files + a JSON manifest, for
julia/dump_reference.jl: the real NumericalEarth.jl, on a machine that has Julia, builds the same case from these bits
and writes the same .npz keys, which tests/test_golden.py then compares against (REFERENCE goldens instead of oracle ones).

    python tests/golden/export_inputs.py [outdir]        # default tests/golden/inputs/

Layout of every array: exactly what this repository keeps in memory, i.e. parents with halos,
  exchange fields   (ny + 2hy, nx + 2hx)            C order == Oceananigans' column-major (nx + 2hx, ny + 2hy)
  atmosphere series (nt, ny_a + 2hy_a, nx_a + 2hx_a) C order == FieldTimeSeries parent (nx_a + 2hx_a, ny_a + 2hy_a, 1, nt)
so Julia reads them with `permutedims` reversed (NPZ.jl returns column-major arrays of the reversed shape)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402


class _Recorder(ne_b200.NumpyHostBackend):
    """Host back-end that needs no compute library: only the inputs are generated."""


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs")
    os.makedirs(out, exist_ok=True)
    import oracle
    lib = oracle.load()
    manifest = {"T_STEP": make_golden.T_STEP, "cases": {}}
    for name, kw in make_golden.CASES.items():
        ci = synthetic.build_case("tiny", _Recorder(), lib=lib, with_iterations=True, **kw)
        g, src = ci.grid, ci.atmosphere.grid
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        for k, a in ci._host_inputs["atmosphere"].items():
            np.save(os.path.join(d, f"atm_{k}.npy"), a)
        for k, a in ci._host_inputs["ocean"].items():
            np.save(os.path.join(d, f"ocean_{k}.npy"), a)
        col = None
        if kw.get("sea_ice"):
            T3, S3, dz = synthetic.ocean_column(g, _Recorder(), nz=4)
            np.save(os.path.join(d, "column_T.npy"), T3); np.save(os.path.join(d, "column_S.npy"), S3); np.save(os.path.join(d, "column_dz.npy"), dz)
            col = {"nz": 4, "dt": 1200.0}
        manifest["cases"][name] = {
            "exchange": {"nx": g.nx, "ny": g.ny, "hx": g.hx, "hy": g.hy, "longitude": list(g.longitude), "latitude": list(g.latitude), "FT": g.FT},
            "atmosphere": {"nx": src.nx, "ny": src.ny, "hx": src.hx, "hy": src.hy, "FT": src.FT, "times": [float(t) for t in ci.atmosphere.times],
                           "surface_layer_height": ci.atmosphere.surface_layer_height, "boundary_layer_height": ci.atmosphere.boundary_layer_height,
                           "time_indexing": "Cyclical"},
            "sea_ice": bool(kw.get("sea_ice")), "column": col,
            "ocean_properties": {"reference_density": ci.ocean_properties.reference_density, "heat_capacity": ci.ocean_properties.heat_capacity},
            "outputs": f"{name}.npz"}
    with open(os.path.join(out, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
