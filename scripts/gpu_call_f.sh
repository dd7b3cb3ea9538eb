#!/bin/bash
# tuning-flag experiment at N >= 640 (same-key timings), TUNE values given as arguments
for t in "$@"; do
  TUNE=$t timeout 600 python scripts/bench_configs.py c3s c4 c5s 2> gpurun_out/tune_err.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TUNE', d.get('tune'), d['config'], 'enc %.3f dec %.3f'%(d['enc_ms'],d['dec_ms']), d['kernel_ms'], '%.1f M ct/s'%(d['ct_per_s']/1e6), d['roundtrip_equals_message'])
"
done
tail -3 gpurun_out/tune_err.log
