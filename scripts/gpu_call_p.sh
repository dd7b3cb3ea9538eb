#!/bin/bash
nvidia-smi -L
for F in 1 2; do
DEC1_FORM=$F timeout 300 python scripts/bench_configs.py c2 c3s 2>&1 | grep '^{' | tee -a gpurun_out/r2_dec1_form_lohi.jsonl
done
TRACE_CFG=hps509 DEC1_FORM=2 timeout 300 python scripts/trace_summary.py dec1 2>&1 | tail -22 | tee -a gpurun_out/r2_trace_summary_f16.txt
