#!/bin/bash
# TUNE=8 (probe inside the warp-uniform issue loop): correctness stress, then timings against TUNE=0
for cfg in "167 128" "509 2048" "701 8192" "821 4096"; do
  TUNE=$1 timeout 300 python scripts/sanitize_target.py $cfg 40000 4 2>&1 | tail -1
done
bash scripts/gpu_call_f.sh 0 $1
TUNE=$1 timeout 300 python scripts/bench_configs.py c1 c2 2>> gpurun_out/tune_err.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TUNE', d.get('tune'), d['config'], 'enc %.3f dec %.3f'%(d['enc_ms'],d['dec_ms']), d['kernel_ms'], '%.1f M ct/s'%(d['ct_per_s']/1e6))
"
