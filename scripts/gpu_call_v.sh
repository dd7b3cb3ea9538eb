#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
TRACE_CFG=hps821 timeout 300 python scripts/trace_slices.py dec2 110 45 2>&1 | tail -48 | tee -a gpurun_out/r2_trace_slices.txt
NTRU_DEBUG_NOB=1 NTRU_DEBUG_ONE_MMA=1 TRACE_CFG=hps821 timeout 300 python scripts/trace_slices.py dec2 110 45 2>&1 | tail -48 | tee -a gpurun_out/r2_trace_slices.txt
