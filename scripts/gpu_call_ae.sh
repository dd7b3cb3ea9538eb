#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
echo "=== shipped" | tee -a gpurun_out/r2_double_mma.txt
timeout 300 python scripts/trace_light.py hps509 hps821 2>&1 | tail -6 | tee -a gpurun_out/r2_double_mma.txt
echo "=== NTRU_DEBUG_DOUBLE_MMA=1 (eight MMAs per slice)" | tee -a gpurun_out/r2_double_mma.txt
NTRU_DEBUG_DOUBLE_MMA=1 timeout 300 python scripts/trace_light.py hps509 hps821 2>&1 | tail -6 | tee -a gpurun_out/r2_double_mma.txt
