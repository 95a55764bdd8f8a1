#!/bin/bash
# final-tree evidence: full GPU suite, memcheck of the new kernels' tests, bench line, launch list with the NVTX ranges
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j24_pytest.log
tail -4 gpurun_out/j24_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_series_window.py tests/test_land_fluxes.py -m gpu -q --no-header -p no:cacheprovider -k "column or per_cell or mangling" > gpurun_out/j24_memcheck_new_kernels.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/j24_memcheck_new_kernels.log
tail -4 gpurun_out/j24_memcheck_new_kernels.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j24_bench.json 2> gpurun_out/j24_bench.err
tail -2 gpurun_out/j24_bench.err
timeout 600 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j24_launches_nvtx.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-seconds 0 --no-parity > gpurun_out/j24_launches_bench.log 2>&1
head -30 gpurun_out/j24_launches_nvtx.csv | cut -c1-220
