#!/bin/bash
nvidia-smi -L
for V in "DEC1_FORM=1 EPILOGUE=0" "DEC1_FORM=2 EPILOGUE=1" "DEC1_FORM=2 EPILOGUE=2" "DEC1_FORM=1 EPILOGUE=0" "DEC1_FORM=2 EPILOGUE=2"; do
env $V timeout 600 python scripts/bench_configs.py c2 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$V', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6), d['roundtrip_equals_message'])
" | tee -a gpurun_out/r2_dec1f_one_group.txt
done
