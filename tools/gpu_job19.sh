#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j19_pytest.log
tail -6 gpurun_out/j19_pytest.log
