#!/bin/bash
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or outside_baseline" 2>&1 | tail -3
for cfg in hps821 hps509 hps677; do
timeout 600 python scripts/ab_lib.py $cfg experiments/ab/libntru_b200_head.so ntru-circom_b200/libntru_b200.so experiments/ab/libntru_b200_head.so ntru-circom_b200/libntru_b200.so 2>&1 | tee -a gpurun_out/r2_ab_uniform_wait.jsonl
done
