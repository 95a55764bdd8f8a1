#!/bin/bash
# round-2 GPU job 4: bitonic ordering pass, full GPU suite, bench, ncu of the solve, launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 900 python tools/time_ao.py --config C4 --out j4_time_ao NE_B200_TAB_V1=1 "" NE_B200_TAB2_WINDOW=512 NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB_WAVES=8 NE_B200_TAB_WAVES=16 "" > gpurun_out/j4_time_ao.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j4_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j4_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j4_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-seconds 0 > gpurun_out/j4_launches_bench.log 2>&1
tail -12 gpurun_out/j4_pytest.log
cat gpurun_out/j4_time_ao.log
