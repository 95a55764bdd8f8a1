#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
python tools/check_env_bitwise.py C4 NE_B200_TAB2_NO_ORDER=1 "" > gpurun_out/j10_bitwise.log 2>&1
timeout 900 python tools/time_ao.py --config C4 --out j10_time_ao NE_B200_TAB_V1=1 "" NE_B200_TAB2_PREFETCH=1 "" NE_B200_TAB2_PREFETCH=1 > gpurun_out/j10_time_ao.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/j10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j10_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j10_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j10_ncu.log 2>&1
cat gpurun_out/j10_bitwise.log gpurun_out/j10_time_ao.log
tail -5 gpurun_out/j10_pytest.log
grep summary gpurun_out/parity_r02.jsonl | cut -c1-300
