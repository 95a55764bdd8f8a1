#!/bin/bash
# round-2 GPU job 2: LDS microbenchmark, GPU test-suite, solve-kernel variants after the L1TEX-pipe changes, bench, ncu
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
./tools/mb/mb_lds > gpurun_out/j2_mb_lds.json 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
timeout 900 python tools/time_ao.py --config C4 --out j2_time_ao NE_B200_TAB_V1=1 "" NE_B200_TAB2_WARPS=7 NE_B200_TAB2_WINDOW=512 NE_B200_TAB2_WINDOW=256 NE_B200_TAB2_WINDOW=256,NE_B200_TAB2_WARPS=7 NE_B200_TAB2_WINDOW=512,NE_B200_TAB2_WARPS=7 NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB_WAVES=4 NE_B200_TAB_WAVES=16 NE_B200_TAB_WAVES=32 NE_B200_TAB2_WARPS=7,NE_B200_TAB_WAVES=16 "" > gpurun_out/j2_time_ao.log 2>&1
timeout 300 python tools/time_seaice.py C3 f64 > gpurun_out/j2_time_seaice.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j2_tab2 python tools/prof_ao.py C4 f64 > gpurun_out/j2_ncu.log 2>&1
tail -25 gpurun_out/j2_pytest.log
cat gpurun_out/j2_time_ao.log gpurun_out/j2_mb_lds.json
tail -12 gpurun_out/j2_time_seaice.log
