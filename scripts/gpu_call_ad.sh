#!/bin/bash
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or outside_baseline" 2>&1 | tail -3
for cfg in hps509 hps821; do
timeout 600 python scripts/ab_lib.py $cfg experiments/ab/libntru_b200_head.so experiments/ab/libntru_b200_relax0.so ntru-circom_b200/libntru_b200.so experiments/ab/libntru_b200_relax100.so experiments/ab/libntru_b200_relax400.so experiments/ab/libntru_b200_head.so 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'].split('/')[-1], d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6))
" | tee -a gpurun_out/r2_relaxed_wait.txt
done
