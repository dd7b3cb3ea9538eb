#!/bin/bash
# parity of everything that runs on the IMMA schedule, then the distinct-key configs with the production library
timeout 900 python -m pytest tests -m gpu -x -q -k "distinct or general or schedules or golden or random or keys or small or ragged or string or class or pad or many_tiles or verify or keygen" 2>&1 | tail -5 | tee gpurun_out/r2_imma_v2_tests.txt
timeout 300 python scripts/bench_configs.py c3 c3_509 c3_821 c3_167 2>&1 | python3 -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'],'enc %.4f dec %.4f ms  %.1f M ct/s  frac_hbm %.3f'%(d['enc_ms'],d['dec_ms'],d['ct_per_s']/1e6,d['frac_hbm']),'ok' if d['matches_cuda_core_schedule'] else 'MISMATCH')
" | tee gpurun_out/r2_imma_v2_bench.txt
