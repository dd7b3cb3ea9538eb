#!/bin/bash
nvidia-smi -L
python scripts/trace_summary.py --build 2>/dev/null || exit 1
export NTRU_B200_LIB=$PWD/ntru-circom_b200/libntru_trace.so
for V in "" "NTRU_DEBUG_NO_B_STORE=1"; do for F in 1 2; do
env $V DEC1_FORM=$F timeout 300 python scripts/bench_configs.py c2 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('trace build, form $F, ${V:-b stored}:', d['kernel_ms'], d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_dec1f_b_store_cost.txt
done; done
