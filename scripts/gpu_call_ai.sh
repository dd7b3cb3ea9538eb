#!/bin/bash
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for E in 0 1 0 1; do
EPILOGUE=$E timeout 600 python scripts/bench_configs.py c1 c2 c5s 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('EPILOGUE=$E', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6))
" | tee -a gpurun_out/r2_one_group_auto.txt
done
