#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j29_bench.json 2> gpurun_out/j29_bench.err
tail -2 gpurun_out/j29_bench.err
timeout 600 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j29_launches_nvtx.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-seconds 0 --no-parity > gpurun_out/j29_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 3 -c 1 -f -o gpurun_out/j29_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j29_ncu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j29_pytest.log
tail -3 gpurun_out/j29_pytest.log
