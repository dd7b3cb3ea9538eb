#!/bin/bash
# cheaper lift in the DEC1 epilogue (odd log2 q): parity on the tcgen05 path, then A/B against the previous build
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or random or pad or many_tiles or general or full_size or fp16 or schedules" 2>&1 | tail -4 | tee gpurun_out/r2_lift_tests.txt
for cfg in hps509 hps677 hrss701 hps821 default167; do
  timeout 300 python scripts/ab_lib.py $cfg experiments/ab/libntru_b200_head.so experiments/ab/libntru_b200_lift.so experiments/ab/libntru_b200_head.so experiments/ab/libntru_b200_lift.so 2>&1 | python3 -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'],d['lib'][-12:],d['kernel_ms'],'%.1f M ct/s'%(d['ct_per_s']/1e6),d['checksum'])
" | tee -a gpurun_out/r2_lift_ab.txt
done
