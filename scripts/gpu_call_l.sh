#!/bin/bash
# A/B of the library at 3c800f1 (before the fp16 form and the wire format) against HEAD, device-resident kernels
nvidia-smi -L
for cfg in hps509 hps821; do
timeout 600 python scripts/ab_lib.py $cfg experiments/ab/libntru_b200_3c800f1.so ntru-circom_b200/libntru_b200.so experiments/ab/libntru_b200_3c800f1.so ntru-circom_b200/libntru_b200.so 2>&1 | tee -a gpurun_out/r2_ab_head_vs_3c800f1.jsonl
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_element" 2>&1 | tail -3
