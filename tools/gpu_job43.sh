#!/bin/bash
# the very last pass: full GPU suite + default bench line on the final tree
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j43_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j43_pytest.log
tail -3 gpurun_out/j43_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/j43_bench.json 2> gpurun_out/j43_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/j43_bench.json').read().strip().split('\n')[-1])
print('ms/step',round(j['ms_per_step'],4),'value',round(j['value']/1e9,3),'e2e',round(j['e2e']['ms_per_step'],3),'roof',round(j['roofline']['frac'],3),'hbm',round(j['roofline_hbm']['frac'],3),'solve',round(j['roofline']['ms_per_launch'],4),'parity',j['parity']['points_above_tol'],j['parity']['trip_count_mismatch_rate'], j['other_configs'].get('C3_ocean_sea_ice_f64_ms_per_step'))
PY
