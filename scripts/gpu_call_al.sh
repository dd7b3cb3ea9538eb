#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
timeout 300 python scripts/trace_light.py hps509 2>&1 | tail -3 | tee gpurun_out/r2_light_head.txt
TRACE_CFG=hps509 timeout 300 python scripts/trace_summary.py dec1 2>&1 | tail -16 | tee -a gpurun_out/r2_light_head.txt
