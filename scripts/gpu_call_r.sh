#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
for V in "" "NTRU_DEBUG_ONE_MMA=1" "NTRU_DEBUG_NOB=1" "NTRU_DEBUG_NOB=1 NTRU_DEBUG_ONE_MMA=1"; do
echo "=== variant: ${V:-shipped}" | tee -a gpurun_out/r2_issuer_cost.txt
env $V TRACE_CFG=hps821 timeout 300 python scripts/trace_summary.py dec2 2>&1 | grep -A5 "phases per tile" | tee -a gpurun_out/r2_issuer_cost.txt
env $V TRACE_CFG=hps821 timeout 300 python scripts/trace_summary.py enc 2>&1 | grep -A5 "phases per tile" | tee -a gpurun_out/r2_issuer_cost.txt
done
