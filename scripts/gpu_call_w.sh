#!/bin/bash
# two MMA issuer warps on alternating slices: parity (short timeouts: a protocol error hangs), then timing against one issuer
nvidia-smi -L
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "golden failed or hung"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or tile_boundaries or outside_baseline or pad_columns or fp16_form or random_batch" 2>&1 | tail -3
for I in 0 1; do
ISSUERS=$I timeout 600 python scripts/bench_configs.py c1 c2 c3s c4 c5s 2>&1 | grep '^{' | tee -a gpurun_out/r2_two_issuers.jsonl
done
