#!/bin/bash
# A/B of library builds on the distinct-key configs: gpu_ab_distinct.sh "<configs>" variant...
cfgs=$1; shift
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/r2_imma_ab2.txt
  NTRU_B200_LIB=$PWD/experiments/ab/libntru_b200_$v.so timeout 300 python scripts/bench_configs.py $cfgs 2>&1 | python3 -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'],'enc %.4f dec %.4f ms  %.1f M ct/s'%(d['enc_ms'],d['dec_ms'],d['ct_per_s']/1e6),'ok' if d['matches_cuda_core_schedule'] else 'MISMATCH')
" | tee -a gpurun_out/r2_imma_ab2.txt
done
