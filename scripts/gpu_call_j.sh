#!/bin/bash
# DEC1 on fp16 tiles (q <= 2048): parity of the new form, then both forms timed
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp16_form" 2>&1 | tail -8
for F in 0 1; do
DEC1_FORM=$F timeout 300 python scripts/bench_configs.py c2 c3s 2>&1 | tail -2 | tee -a gpurun_out/r2_dec1_fp16_form.jsonl
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_batch or many_tiles or tile_boundaries or outside_baseline or pad_columns" 2>&1 | tail -5
