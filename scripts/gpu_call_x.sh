#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
TRACE_CFG=hps821 timeout 300 python scripts/trace_slices.py dec2 110 60 2>&1 | tail -34 | tee -a gpurun_out/r2_trace_slices_dual.txt
