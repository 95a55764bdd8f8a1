#!/bin/bash
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_ORDER=0" > gpurun_out/j27_bitwise.log 2>&1
python tools/check_env_bitwise.py C1 "" "NE_B200_TAB2_NO_ORDER=1" >> gpurun_out/j27_bitwise.log 2>&1
grep bitwise gpurun_out/j27_bitwise.log
python tools/time_ao.py --out j27_time_ao "" NE_B200_TAB2_ORDER=1 NE_B200_TAB2_ORDER=0 "" NE_B200_TAB2_ORDER=1 NE_B200_TAB2_ORDER=0 > gpurun_out/j27_time_ao.log 2>&1
cat gpurun_out/j27_time_ao.log
python tools/time_ao.py --config C2 --out j27_time_ao_C2 "" NE_B200_TAB2_ORDER=1 NE_B200_TAB2_ORDER=0 >> gpurun_out/j27_time_ao.log 2>&1
tail -3 gpurun_out/j27_time_ao.log
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"ao_flux_tab2|trip_order" --csv --log-file gpurun_out/j27_hist.csv python tools/prof_ao.py C4 f64 > gpurun_out/j27_a.log 2>&1
grep -E "gpu__time_duration|bank_conflicts" gpurun_out/j27_hist.csv | tail -4 | sed 's/.*"\(void [a-z_0-9]*\).*Command line profiler metrics",/\1 /'
