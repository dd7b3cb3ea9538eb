#!/bin/bash
# the whole GPU suite, then the distinct-key configs
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_gpu_tests_full.txt
bash scripts/gpu_distinct_key.sh 2>&1 | tail -4
