#!/bin/bash
# two slices per trip of the MMA issue loop: parity, then timing against the committed build
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" 2>&1 | tail -3 || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or tile_boundaries or outside_baseline or pad_columns or fp16_form or random_batch" 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py c1 c2 c3s c4 c5s 2>&1 | grep '^{' | tee -a gpurun_out/r2_two_slice_issue.jsonl
