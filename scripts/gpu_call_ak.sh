#!/bin/bash
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'whole', d['roofline']['whole_step'])
print({k:v['avg_ms'] for k,v in d['roofline']['kernels'].items()})
e=d['e2e']; print('e2e', e['value'], e['frac_of_pcie'], 'plain', e['plain_arrays']['value'])
for k,v in d['extras']['configs'].items(): print(k, {x:v[x] for x in v if x in ('ms','ct_per_s','frac_hbm','matches_int64_column_sums')})
PY
