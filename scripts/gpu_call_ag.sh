#!/bin/bash
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" 2>&1 | tail -2
python scripts/trace_summary.py --build 2>/dev/null || exit 1
for L in 0 1; do
echo "=== LEAN=$L (light profile, trace build)" | tee -a gpurun_out/r2_lean2.txt
LEAN=$L timeout 300 python scripts/trace_light.py hps509 hps821 2>&1 | tail -6 | tee -a gpurun_out/r2_lean2.txt
done
for L in 1 0 1 0; do
LEAN=$L timeout 600 python scripts/bench_configs.py c2 c4 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('production build LEAN=$L', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6), d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_lean2.txt
done
