#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build || exit 1
for cfg in hps821 hps509; do
for mode in enc dec1 dec2; do
TRACE_CFG=$cfg timeout 300 python scripts/trace_summary.py $mode 2>&1 | tail -22 | tee -a gpurun_out/r2_trace_summary.txt
done
done
