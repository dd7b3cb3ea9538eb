#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
timeout 300 python scripts/trace_light.py hps509 hps677 hps821 2>&1 | tail -12 | tee gpurun_out/r2_issuer_light_profile.txt
