#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_seaice.py > gpurun_out/j35_seaice.log 2>&1; grep -E "atmosphere_sea_ice_fluxes|update_state\(all\)" gpurun_out/j35_seaice.log | head -3
timeout 300 python tools/time_seaice.py C4 >> gpurun_out/j35_seaice.log 2>&1; grep -E "C4 f64 atmosphere_sea_ice_fluxes" gpurun_out/j35_seaice.log | head -2
rm -f gpurun_out/parity_r02.jsonl
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "sea_ice or C3 or seaice or golden or queue or albedo or ice" > gpurun_out/j35_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j35_pytest.log
tail -8 gpurun_out/j35_pytest.log
grep C3 gpurun_out/parity_r02.jsonl | cut -c1-600
