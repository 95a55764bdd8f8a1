#!/bin/bash
# last pass on the final tree: full GPU suite, default bench at N = 1 and N = 2 (as the driver launches them)
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/j40_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j40_pytest.log
tail -3 gpurun_out/j40_pytest.log
timeout 900 python bench.py > gpurun_out/j40_bench_n1.json 2> gpurun_out/j40_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/j40_bench_n2.json 2> gpurun_out/j40_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29734 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/j40_ref_n2.json 2> gpurun_out/j40_ref_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        j=json.loads(open(f'gpurun_out/j40_bench_n{n}.json').read().strip().split('\n')[-1])
        print(n,'ms/step',round(j['ms_per_step'],4),'e2e',round(j['e2e']['ms_per_step'],3),'launches',j['gpu_launches'],'roof',j['roofline']['frac'],'parity',(j.get('parity') or {}).get('points_above_tol'),'part',j['config']['partition'][-90:])
    except Exception as e:
        print(n,'ERR',e)
print(open('gpurun_out/j40_ref_n2.json').read()[:200])
PY
