#!/bin/bash
# staged (cp.async) group inputs of the a-o solve: bitwise check against the non-staged launch, timing, new GPU tests
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_NO_STAGING=1" > gpurun_out/j21_bitwise.log 2>&1
python tools/check_env_bitwise.py C2 "" "NE_B200_TAB2_NO_STAGING=1" >> gpurun_out/j21_bitwise.log 2>&1
tail -4 gpurun_out/j21_bitwise.log
python tools/time_ao.py --out j21_time_ao "" NE_B200_TAB2_NO_STAGING=1 "" NE_B200_TAB2_NO_STAGING=1 > gpurun_out/j21_time_ao.log 2>&1
cat gpurun_out/j21_time_ao.log
python tools/time_ao.py --config C2 --out j21_time_ao_C2 "" NE_B200_TAB2_NO_STAGING=1 >> gpurun_out/j21_time_ao.log 2>&1
tail -2 gpurun_out/j21_time_ao.log
timeout 900 python -m pytest tests/test_series_window.py tests/test_land_fluxes.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j21_pytest.log
tail -5 gpurun_out/j21_pytest.log
