#!/bin/bash
mkdir -p gpurun_out
python tools/time_interp.py C4 NE_B200_INTERP_STAGED_V1=1 "" NE_B200_INTERP_DIRECT=1 "" > gpurun_out/j17_interp.log 2>&1
python tools/time_interp.py C2 NE_B200_INTERP_STAGED_V1=1 "" >> gpurun_out/j17_interp.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x -k "interp or staged or fused or full_size or land_and_rotation or series_window or C4 or C5" > gpurun_out/j17_pytest.log 2>&1
cat gpurun_out/j17_interp.log; tail -4 gpurun_out/j17_pytest.log
