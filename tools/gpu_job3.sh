#!/bin/bash
# round-2 GPU job 3: (trips, record)-ordered solve, bench, ncu, sanitizer on the queue kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 900 python tools/time_ao.py --config C4 --out j3_time_ao NE_B200_TAB_V1=1 "" NE_B200_TAB2_WINDOW=512 NE_B200_TAB2_WINDOW=256 NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB_WAVES=2 NE_B200_TAB_WAVES=8 NE_B200_TAB2_WARPS=7 "" > gpurun_out/j3_time_ao.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j3_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j3_bench.json 2> gpurun_out/j3_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j3_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j3_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-seconds 0 > gpurun_out/j3_launches_bench.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "test_float32_work_queue_edge_cases and (odd_size or fixed_iterations) or test_ocean_sea_ice_model_step and conductive" > gpurun_out/j3_racecheck_queue.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/j3_racecheck_queue.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "test_float32_work_queue_edge_cases and odd_size or test_ocean_sea_ice_model_step and conductive or test_atmosphere_ocean_fluxes_f64 and C1" > gpurun_out/j3_memcheck_kernels.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/j3_memcheck_kernels.log
tail -12 gpurun_out/j3_pytest.log
cat gpurun_out/j3_time_ao.log
tail -4 gpurun_out/j3_racecheck_queue.log gpurun_out/j3_memcheck_kernels.log
