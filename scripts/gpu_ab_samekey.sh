#!/bin/bash
# A/B of library builds on the same-key tcgen05 kernels: gpu_ab_samekey.sh "<cfgs>" variant variant ... (experiments/ab/libntru_b200_<variant>.so)
cfgs=$1; shift
for cfg in $cfgs; do
  libs=""
  for v in "$@" "$@"; do libs="$libs experiments/ab/libntru_b200_$v.so"; done
  timeout 600 python scripts/ab_lib.py $cfg $libs 2>&1 | python3 -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['config'],d['lib'].split('_')[-1],d['kernel_ms'],'%.1f M ct/s'%(d['ct_per_s']/1e6),d['checksum'])
"
done
