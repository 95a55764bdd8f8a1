#!/bin/bash
# CTA shapes 28 warps @ 72 registers and 25 warps @ 80 against the shipped 24 @ 80; full GPU suite
mkdir -p gpurun_out
python tools/time_ao.py --out j22_time_ao "" NE_B200_TAB2_SHAPE=3 NE_B200_TAB2_SHAPE=4 "" NE_B200_TAB2_SHAPE=3 NE_B200_TAB2_SHAPE=4 > gpurun_out/j22_time_ao.log 2>&1
cat gpurun_out/j22_time_ao.log
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_SHAPE=3" > gpurun_out/j22_bitwise.log 2>&1
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_SHAPE=4" >> gpurun_out/j22_bitwise.log 2>&1
tail -2 gpurun_out/j22_bitwise.log
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j22_pytest.log
tail -6 gpurun_out/j22_pytest.log
