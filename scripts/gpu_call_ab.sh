#!/bin/bash
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp16_form or golden or outside_baseline" 2>&1 | tail -3
for F in 1 2 1 2; do
DEC1_FORM=$F timeout 300 python scripts/bench_configs.py c2 c3s 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('form $F:', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6), d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_dec1f_b_staged.txt
done
