#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_seaice.py > gpurun_out/j37_seaice.log 2>&1; grep -E "C3 f64 (sea_ice_ocean_fluxes|update_state\(all\)|atmosphere_sea_ice)" gpurun_out/j37_seaice.log | head -4
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "sea_ice or C3 or freez or heat_flux or ocean_only or paths" > gpurun_out/j37_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j37_pytest.log
tail -4 gpurun_out/j37_pytest.log
