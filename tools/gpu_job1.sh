#!/bin/bash
# round-2 GPU job 1: full GPU test-suite, solve-kernel variants, bench line, one full-set ncu capture, sanitizer
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/j1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
timeout 600 python tools/time_ao.py --config C4 --out j1_time_ao NE_B200_TAB_V1=1 "" NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB2_LIBM_PROLOGUE=1 NE_B200_TAB2_NO_ORDER=1,NE_B200_TAB2_LIBM_PROLOGUE=1 NE_B200_TAB_WAVES=2 NE_B200_TAB_WAVES=8 NE_B200_TAB_WAVES=16 NE_B200_TAB_V1=1 "" > gpurun_out/j1_time_ao.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j1_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j1_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python __graft_entry__.py smoke > gpurun_out/j1_memcheck_smoke.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/j1_memcheck_smoke.log
ls -la gpurun_out | tail -20
tail -30 gpurun_out/j1_pytest.log
cat gpurun_out/j1_time_ao.log
