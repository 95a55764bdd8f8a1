#!/bin/bash
# per-kernel durations (ncu, cold cache) of the ordering pass and the solve under both orderings; full-set capture of the solve with the counting-sort order
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"ao_flux_tab2|trip_order" --csv --log-file gpurun_out/j26_counting.csv python tools/prof_ao.py C4 f64 > gpurun_out/j26_a.log 2>&1
NE_B200_TAB2_BITONIC=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"ao_flux_tab2|trip_order" --csv --log-file gpurun_out/j26_bitonic.csv python tools/prof_ao.py C4 f64 > gpurun_out/j26_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 3 -c 1 -f -o gpurun_out/j26_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j26_ncu.log 2>&1
ls -la gpurun_out/j26_tab2.ncu-rep
tail -n 12 gpurun_out/j26_counting.csv | cut -c1-400
