#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -k "interp or staged or window or potential or rotat or runoff or C4 or C5 or step or golden or sharding or pipeline" > gpurun_out/j42_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j42_pytest.log
tail -4 gpurun_out/j42_pytest.log
python tools/time_interp.py C5 NE_B200_INTERP_ROWS=1 "" 2>&1 | tail -2
