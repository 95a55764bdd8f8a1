#!/bin/bash
# packed-pair (FMUL2 / FFMA2) staged interpolation + REDUX window search: timing, bitwise against the direct-gather kernel, interpolation tests
mkdir -p gpurun_out
python tools/time_interp.py C4 NE_B200_INTERP_DIRECT=1 "" "" > gpurun_out/j23_interp.log 2>&1
cat gpurun_out/j23_interp.log
python tools/time_interp.py C2 NE_B200_INTERP_DIRECT=1 "" >> gpurun_out/j23_interp.log 2>&1
tail -2 gpurun_out/j23_interp.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py tests/test_land_and_rotation.py tests/test_series_window.py -m gpu -q --no-header -rf -p no:cacheprovider -k "interp or staged or step or window or potential or rotat or runoff or C4" > gpurun_out/j23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j23_pytest.log
tail -5 gpurun_out/j23_pytest.log
