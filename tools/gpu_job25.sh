#!/bin/bash
# counting-sort trip ordering against the bitonic (trips, record) ordering: bitwise, solve time, step time
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_BITONIC=1" > gpurun_out/j25_bitwise.log 2>&1
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_NO_ORDER=1" >> gpurun_out/j25_bitwise.log 2>&1
python tools/check_env_bitwise.py C1 "" "NE_B200_TAB2_NO_ORDER=1" >> gpurun_out/j25_bitwise.log 2>&1
grep bitwise gpurun_out/j25_bitwise.log
python tools/time_ao.py --out j25_time_ao "" NE_B200_TAB2_BITONIC=1 NE_B200_TAB2_NO_ORDER=1 "" NE_B200_TAB2_BITONIC=1 > gpurun_out/j25_time_ao.log 2>&1
cat gpurun_out/j25_time_ao.log
python tools/time_ao.py --config C2 --out j25_time_ao_C2 "" NE_B200_TAB2_BITONIC=1 >> gpurun_out/j25_time_ao.log 2>&1
tail -2 gpurun_out/j25_time_ao.log
for c in C4; do NE_CFG=$c python tools/time_step.py "" NE_B200_TAB2_BITONIC=1 > gpurun_out/j25_step_$c.log 2>&1; tail -3 gpurun_out/j25_step_$c.log; done
