#!/bin/bash
# fat ring stages (two slices per stage and issue trip): parity with short timeouts, then timing against thin stages
nvidia-smi -L
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "golden failed or hung"; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or tile_boundaries or outside_baseline or pad_columns or fp16_form or random_batch" 2>&1 | tail -3
for S in 0 1 0 1; do
STAGES=$S timeout 600 python scripts/bench_configs.py c1 c2 c3s c4 c5s 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('STAGES=$S', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6), 'frac %.3f' % d['frac_hbm'], d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_fat_stages.txt
done
