#!/bin/bash
nvidia-smi -L
for L in 1 0 1 0; do
LEAN=$L timeout 600 python scripts/bench_configs.py c2 c4 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('LEAN=$L', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6), d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_lean_issue.txt
done
