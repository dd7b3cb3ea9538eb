#!/bin/bash
# default bench at HEAD (one GPU)
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo rc=$?
tail -c 600 gpurun_out/r2_bench_1gpu.err
python3 - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','roofline','e2e','gpu_launches') if k in d})
for k,v in d.get('extras',{}).get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in ('ct_per_s','frac_hbm','ms','value','rows','mode')})
PY
