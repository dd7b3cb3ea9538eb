#!/bin/bash
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp16_form" 2>&1 | tail -4
for F in 2 3; do
DEC1_FORM=$F timeout 300 python scripts/bench_configs.py c2 c3s 2>&1 | grep '^{' | tee -a gpurun_out/r2_dec1_form_pu1.jsonl
done
