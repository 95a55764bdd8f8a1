#!/bin/bash
# final tree: smoke, full GPU suite, bench line, graph-replay check
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j33_smoke.log 2>&1; tail -1 gpurun_out/j33_smoke.log
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j33_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j33_pytest.log
tail -4 gpurun_out/j33_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j33_bench.json 2> gpurun_out/j33_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/j33_bench.json').read().strip().split('\n')[-1])
print('ms/step',j['ms_per_step'],'value',j['value'],'e2e',j['e2e']['ms_per_step'],'roof',j['roofline']['frac'],j['roofline']['traffic'],'solve',j['roofline']['ms_per_launch'],'parity',j['parity']['points_above_tol'],j['parity']['trip_count_mismatch_rate'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j33_ref.json 2> gpurun_out/j33_ref.err; cut -c1-300 gpurun_out/j33_ref.json
